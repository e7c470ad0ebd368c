"""rgbdslam_b200_landmark_ba (Schur-complement BA on the GPU) against the dense full-system oracle (oracle/landmark_oracle.py)."""
import numpy as np
import pytest

from oracle import landmark_oracle as lo
from rgbdslam_v2_b200 import _capi, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fe():
    f = _capi.Frontend(0)
    yield f
    f.close()


def _oracle(d, edges, iters):
    kw = dict(ij=d["ij"], meas=d["meas"], info=d["info"]) if edges else {}
    P = lo.Problem(d["poses"], d["fixed"], d["points"], d["obs_cam"], d["obs_point"], d["obs_uvd"], d["obs_info3"], d["K4"], **kw)
    c0 = P.chi2()
    c1 = P.optimize(iterations=iters)
    return P, c0, c1


def _gpu(fe, d, edges, iters):
    kw = dict(ij=d["ij"], meas=d["meas"], info=d["info"]) if edges else {}
    return fe.landmark_ba(d["poses"], d["fixed"], d["points"], d["obs_cam"], d["obs_point"], d["obs_uvd"], d["obs_info3"], d["K4"],
                          iterations=iters, **kw)


@pytest.mark.parametrize("edges", [False, True])
def test_same_optimum_as_full_system_oracle(fe, edges):
    d = synth.make_ba_problem(n_cams=6, n_points=60, seed=5)
    P, c0, c1 = _oracle(d, edges, 12)
    x, pts, g0, g1, it, cg = _gpu(fe, d, edges, 12)
    assert abs(g0 - c0) <= 1e-9 * c0          # same cost function
    assert abs(g1 - c1) <= 1e-6 * c1          # same optimum (eliminating the landmarks changes the algebra, not the step)
    assert np.abs(x[:, :3] - P.poses[:, :3]).max() < 1e-6
    dq = np.minimum(np.abs(x[:, 3:] - P.poses[:, 3:]).max(1), np.abs(x[:, 3:] + P.poses[:, 3:]).max(1))
    assert dq.max() < 1e-6
    assert np.abs(pts - P.points).max() < 1e-5
    assert it >= 2 and cg > 0


def test_first_iteration_matches_step_by_step(fe):
    """one LM iteration: identical lambda_0 and step => identical state (tight tolerance, no accumulated path differences)"""
    d = synth.make_ba_problem(n_cams=5, n_points=40, seed=9)
    P, c0, c1 = _oracle(d, True, 1)
    x, pts, g0, g1, it, cg = _gpu(fe, d, True, 1)
    assert it == 1
    assert abs(g1 - c1) <= 1e-7 * c1
    assert np.abs(x[:, :3] - P.poses[:, :3]).max() < 1e-7 and np.abs(pts - P.points).max() < 1e-6


def test_noise_free_recovers_ground_truth(fe):
    d = synth.make_ba_problem(n_cams=8, n_points=300, seed=2, pix_noise=0.0, depth_sigma=0.0, edge_noise=0.0)
    x, pts, g0, g1, it, cg = _gpu(fe, d, True, 25)
    assert g1 < 1e-9 * g0
    assert np.abs(x[:, :3] - d["gt_poses"][:, :3]).max() < 1e-6
    assert np.abs(pts - d["gt_points"]).max() < 1e-6


def test_fixed_cameras_and_unobserved_points_stay(fe):
    d = synth.make_ba_problem(n_cams=5, n_points=50, seed=4)
    d["fixed"][3] = 1
    pts0 = np.vstack([d["points"], [[9.0, 9.0, 9.0]]])  # a landmark without observations
    x, pts, *_ = fe.landmark_ba(d["poses"], d["fixed"], pts0, d["obs_cam"], d["obs_point"], d["obs_uvd"], d["obs_info3"], d["K4"],
                                iterations=8)
    assert np.array_equal(x[0], d["poses"][0]) and np.array_equal(x[3], d["poses"][3])
    assert np.array_equal(pts[-1], [9.0, 9.0, 9.0])
    assert np.abs(x[1] - d["poses"][1]).max() > 1e-4


def test_larger_problem_improves_on_odometry(fe):
    d = synth.make_ba_problem(n_cams=40, n_points=3000, seed=7, edge_noise=0.02)
    x, pts, g0, g1, it, cg = _gpu(fe, d, True, 15)
    assert g1 < 0.05 * g0
    err = np.linalg.norm(x[:, :3] - d["gt_poses"][:, :3], axis=1).max()
    chain = [d["gt_poses"][0]]
    for k in range(len(d["ij"])):
        chain.append(synth.pose_compose(chain[-1], d["meas"][k]))
    err_odo = np.linalg.norm(np.array(chain)[:, :3] - d["gt_poses"][:, :3], axis=1).max()
    assert err < 0.2 * err_odo


def test_argument_errors(fe):
    d = synth.make_ba_problem(n_cams=3, n_points=10, seed=1)
    bad = d["obs_cam"].copy(); bad[0] = 99
    with pytest.raises(_capi.B200Error):
        fe.landmark_ba(d["poses"], d["fixed"], d["points"], bad, d["obs_point"], d["obs_uvd"], d["obs_info3"], d["K4"])
    w = d["obs_info3"].copy(); w[0, 2] = np.inf
    with pytest.raises(_capi.B200Error):
        fe.landmark_ba(d["poses"], d["fixed"], d["points"], d["obs_cam"], d["obs_point"], d["obs_uvd"], w, d["K4"])
