"""End-to-end sequence on the GPU (frames -> Node ctor -> pair matching -> pose graph) vs the same pipeline on the CPU
oracle, and vs ground truth: the ATE part of the north-star metric (BASELINE config C4 at reduced length)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


from oracle.backend import OracleBackend  # noqa: E402  (CPU twin of pipeline.GpuBackend)


def test_sequence_ate_within_1mm_of_the_oracle(built, oracle_mod):
    from oracle import orb_oracle
    from rgbdslam_v2_b200 import Frontend, pipeline, synth
    from rgbdslam_v2_b200._capi import default_params
    n = 36
    poses = synth.trajectory(240)[:n]
    frames = [synth.render_frame(poses[k], seed=k) for k in range(n)]
    gray = np.stack([f[0] for f in frames]); depth = np.stack([f[1] for f in frames])
    mask = np.stack([orb_oracle.depth_to_mask(d) for d in depth])
    K4 = (synth.FX, synth.FY, synth.CX, synth.CY)
    p = default_params(); p.depth_cov_z0 = 2.0; p.max_keypoints = 600
    fe = Frontend(0, p)
    out = pipeline.run_sequence(pipeline.GpuBackend(fe), gray, depth, mask, K4, seed=5)
    ref = pipeline.run_sequence(OracleBackend(oracle_mod, 600), gray, depth, mask, K4, seed=5)
    fe.close()
    gt = np.stack([pipeline.mat_to_pose7(np.linalg.inv(poses[0]) @ P) for P in poses])
    ate, ate_ref = synth.ate_rmse(out["traj"][:, :3], gt[:, :3]), synth.ate_rmse(ref["traj"][:, :3], gt[:, :3])
    valid, valid_ref = out["results"]["id1"] >= 0, ref["results"]["id1"] >= 0
    assert (valid == valid_ref).mean() > 0.98 and valid.sum() > 0.8 * len(valid)
    assert out["graph"]["n_const_edges"] == ref["graph"]["n_const_edges"]
    assert ate < 0.02, ate                                        # the synthetic room is tracked to < 2 cm
    assert abs(ate - ate_ref) < 1e-3, (ate, ate_ref)              # north star: ATE within 1 mm of the reference path
    assert synth.ate_rmse(out["traj"][:, :3], ref["traj"][:, :3]) < 1e-3


def test_online_graph_manager_sequence(built, oracle_mod):
    """The live front (addNode -> nodeComparisons with Dijkstra / keyframe candidates -> addEdgeToG2O -> optimizeGraph per
    node, graph_manager.cpp:204-324, 421-782) driven by the CUDA library vs driven by the CPU oracle: same comparisons,
    same accepted edges, trajectories within 1 mm."""
    from oracle import orb_oracle
    from oracle import graph_manager_oracle as graph_manager
    from rgbdslam_v2_b200 import Frontend, pipeline, synth
    from rgbdslam_v2_b200._capi import default_params
    n = 30
    poses = synth.trajectory(240)[:n]
    frames = [synth.render_frame(poses[k], seed=k) for k in range(n)]
    gray = np.stack([f[0] for f in frames]); depth = np.stack([f[1] for f in frames])
    mask = np.stack([orb_oracle.depth_to_mask(d) for d in depth])
    K4 = (synth.FX, synth.FY, synth.CX, synth.CY)
    p = default_params(); p.depth_cov_z0 = 2.0; p.max_keypoints = 600
    fe = Frontend(0, p)
    gm = graph_manager.run_online(pipeline.GpuBackend(fe), gray, depth, mask, K4, seed=9)
    ref = graph_manager.run_online(OracleBackend(oracle_mod, 600), gray, depth, mask, K4, seed=9)
    fe.close()
    assert gm.comparisons == ref.comparisons                      # same candidate sets (they depend on accepted edges)
    assert len(gm.comparisons[-1][1]) >= 8 and gm.comparisons[-1][1][-1] == n - 2
    same = len(set(gm.edges) & set(ref.edges)) / max(len(ref.edges), 1)
    assert same > 0.98 and gm.n_const_edges == ref.n_const_edges and gm.keyframe_ids == ref.keyframe_ids
    ids, traj = gm.trajectory(); rids, rtraj = ref.trajectory()
    assert list(ids) == list(rids) == list(range(n))
    gt = np.stack([pipeline.mat_to_pose7(np.linalg.inv(poses[0]) @ P) for P in poses])
    ate, ate_ref = synth.ate_rmse(traj[:, :3], gt[:, :3]), synth.ate_rmse(rtraj[:, :3], gt[:, :3])
    assert ate < 0.03 and abs(ate - ate_ref) < 1e-3, (ate, ate_ref)
    assert synth.ate_rmse(traj[:, :3], rtraj[:, :3]) < 1e-3


def test_prune_and_reoptimize_sequence(built, oracle_mod):
    """pruneEdgesWithErrorAbove(5 / 1 / 0.25) + re-optimisation (openni_listener.cpp:431-466) on a graph with
    outlier loop closures: the GPU back-end and the oracle back-end prune the same edges and land on the same
    trajectories."""
    from rgbdslam_v2_b200 import Frontend, pipeline, synth
    g = synth.make_pose_graph(300, 1500, seed=21, outlier_frac=0.05)
    fe = Frontend(0)
    gpu_levels = pipeline.evaluation_sequence(pipeline.GpuBackend(fe), g)
    ref_levels = pipeline.evaluation_sequence(OracleBackend(oracle_mod, 600), g)
    fe.close()
    for (x, chi2, n), (ox, ochi2, on) in zip(gpu_levels, ref_levels):
        assert n == on
        assert chi2 == pytest.approx(ochi2, rel=1e-4, abs=1e-6)
        assert synth.ate_rmse(x[:, :3], ox[:, :3]) < 1e-3
    assert gpu_levels[1][2] > 0  # the outlier edges are found at threshold 5
    ate = [synth.ate_rmse(x[:, :3], g["gt"][:, :3]) for x, _, _ in gpu_levels]
    assert ate[1] <= ate[0] + 1e-3  # pruning outliers does not hurt


def test_tum_trajectory_format(tmp_path):
    from rgbdslam_v2_b200 import pipeline
    p = np.array([[1, 2, 3, 0, 0, 0, 1.0], [0.5, 0, 0, 0, 0, np.sin(0.1), np.cos(0.1)]])
    f = tmp_path / "traj_estimate.txt"
    pipeline.save_trajectory(str(f), p, np.array([10.0, 10.033333]))
    lines = f.read_text().splitlines()
    assert lines[0].startswith("#") and len(lines) == 3
    v = [float(t) for t in lines[2].split()]
    assert len(v) == 8 and v[0] == pytest.approx(10.033333) and v[7] == pytest.approx(np.cos(0.1), abs=1e-6)
