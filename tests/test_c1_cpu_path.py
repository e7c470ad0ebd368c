"""BASELINE config C1 -- "single 640x480 synthetic RGB-D frame pair, ORB 500 kp, BF-match + RANSAC on CPU (reference path, no
GPU)": the whole oracle chain on one rendered pair, checked against the pose the frames were rendered from.  CPU only; this
is the chain the GPU tests compare the CUDA path with stage by stage."""
import numpy as np


def test_c1_single_pair_on_the_cpu_reference_path(oracle_mod):
    from oracle import orb_oracle
    from rgbdslam_v2_b200 import pipeline, synth
    poses = synth.trajectory(240)
    a, b = 0, 6                                              # about 5 cm / 2 degrees apart
    (g0, d0), (g1, d1) = synth.render_frame(poses[a], seed=a)[:2], synth.render_frame(poses[b], seed=b)[:2]
    K4 = (synth.FX, synth.FY, synth.CX, synth.CY)
    st = orb_oracle.DetectorState()
    older = orb_oracle.node_construct(g0, d0, orb_oracle.depth_to_mask(d0), K4, st, max_keypoints=500)
    newer = orb_oracle.node_construct(g1, d1, orb_oracle.depth_to_mask(d1), K4, st, max_keypoints=500)
    assert 300 <= len(older[1]) <= 500 and 300 <= len(newer[1]) <= 500
    assert older[1].shape[1] == 32 and older[2].shape[1] == 4 and np.all(older[2][:, 3] == 1.0)
    prm = oracle_mod.make_params(depth_cov_z0=2.0)
    res, allm, inl = oracle_mod.match_pairs(prm, newer[1], newer[2], [len(newer[1])], older[1], older[2], [len(older[1])], [1], [0],
                                            seed=4)
    r = res[0]
    assert r["id1"] == 0 and r["id2"] == 1 and r["n_inliers"] >= 60 and r["n_all_matches"] <= 300
    T = r["ransac_trafo"].reshape(4, 4).T
    T_true = np.linalg.inv(poses[a]) @ poses[b]              # newer -> older
    assert np.abs(T[:3, 3] - T_true[:3, 3]).max() < 0.01 and np.abs(T[:3, :3] - T_true[:3, :3]).max() < 5e-3
    assert r["info_scale"] == np.float32(r["n_inliers"]) / (np.float32(r["rmse"]) * np.float32(r["rmse"]))
    # the environment measurement model agrees with the estimate and rejects a wrong one
    z_old, z_new = oracle_mod.create_cloud_z(d0), oracle_mod.create_cloud_z(d1)
    c = oracle_mod.pairwise_observation(prm, T, z_new, K4, z_old, K4)
    assert oracle_mod.observation_criterion_met(c[0], c[1], c[2], 0.75)[0]
    Tb = T.copy(); Tb[2, 3] += 0.4
    cb = oracle_mod.pairwise_observation(prm, Tb, z_new, K4, z_old, K4)
    assert not oracle_mod.observation_criterion_met(cb[0], cb[1], cb[2], 0.75)[0]
