"""CPU checks of the landmark bundle-adjustment oracle (oracle/landmark_oracle.py) -- the checker of tests/test_gpu_landmark_ba.py."""
import numpy as np

from oracle import landmark_oracle as lo
from rgbdslam_v2_b200 import synth


def _problem(d, edges=True):
    kw = dict(ij=d["ij"], meas=d["meas"], info=d["info"]) if edges and "ij" in d else {}
    return lo.Problem(d["poses"], d["fixed"], d["points"], d["obs_cam"], d["obs_point"], d["obs_uvd"], d["obs_info3"], d["K4"], **kw)


def test_noise_free_problem_is_recovered_exactly():
    d = synth.make_ba_problem(n_cams=4, n_points=25, seed=1, pix_noise=0.0, depth_sigma=0.0, edge_noise=0.0)
    d["obs_info3"] = synth.landmark_information(d["obs_uvd"][:, 2], 0.0005)
    P = _problem(d)
    c0 = P.chi2()
    c1 = P.optimize(iterations=20)
    assert c0 > 1e3 and c1 < 1e-9 * c0
    assert np.abs(P.poses[:, :3] - d["gt_poses"][:, :3]).max() < 1e-6
    assert np.abs(P.points - d["gt_points"]).max() < 1e-6


def test_obs_error_matches_projection_and_edge_error_is_zero_at_measurement():
    d = synth.make_ba_problem(n_cams=3, n_points=10, seed=2, pix_noise=0.0, depth_sigma=0.0)
    for o in range(len(d["obs_cam"])):
        e = lo.obs_error(d["gt_poses"][d["obs_cam"][o]], d["gt_points"][d["obs_point"][o]], d["obs_uvd"][o], d["K4"])
        assert np.abs(e).max() < 1e-9
    rel = synth.pose_compose(synth.pose_inverse(d["gt_poses"][0]), d["gt_poses"][1])
    assert np.abs(lo.edge_error(d["gt_poses"][0], d["gt_poses"][1], rel)).max() < 1e-12


def test_static_information_quirk():
    w = synth.landmark_information(np.array([2.0, 3.0, 4.0]), 0.01, static_first=True)
    assert np.allclose(w[:, 2], 1.0 / (0.01 * 4.0) ** 2) and np.all(w[:, :2] == 1.0)
    w = synth.landmark_information(np.array([2.0, 3.0]), 0.01)
    assert np.allclose(w[:, 2], [1.0 / (0.01 * 4.0) ** 2, 1.0 / (0.01 * 9.0) ** 2])


def test_landmarks_beat_pose_edges_alone():
    """with noisy odometry edges the landmark observations pull the cameras closer to ground truth than the edges alone would"""
    d = synth.make_ba_problem(n_cams=5, n_points=40, seed=3, edge_noise=0.03)
    P = _problem(d)
    P.optimize(iterations=15)
    err_ba = np.linalg.norm(P.poses[:, :3] - d["gt_poses"][:, :3], axis=1).max()
    chain = [d["gt_poses"][0]]
    for k in range(len(d["ij"])):
        chain.append(synth.pose_compose(chain[-1], d["meas"][k]))
    err_odo = np.linalg.norm(np.array(chain)[:, :3] - d["gt_poses"][:, :3], axis=1).max()
    assert err_ba < 0.5 * err_odo
