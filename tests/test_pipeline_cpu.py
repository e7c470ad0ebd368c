"""Host glue of the batch back-end (rgbdslam_v2_b200/pipeline.py: candidate pairs, addEdgeToG2O rules, pruneEdgesWithErrorAbove,
the evaluation() sequence, TUM export) driven by the CPU oracle instead of the CUDA library -- no GPU needed.  The GPU suite runs
the same code with `pipeline.GpuBackend` and compares trajectories (tests/test_gpu_sequence.py)."""
import numpy as np
import pytest


class OracleGraphBackend:
    def __init__(self, o):
        self.o = o

    def optimize(self, graph, stop):
        x, chi2, _, _ = self.o.posegraph_optimize(graph["init"], graph["fixed"], graph["ij"], graph["meas"], graph["info"], stop=stop)
        return x, chi2

    def edge_chi2(self, poses, graph):
        out = np.zeros(len(graph["ij"]))
        for k, (i, j) in enumerate(graph["ij"]):
            e = self.o.edge_se3(poses[i], poses[j], graph["meas"][k], False)[0]
            out[k] = e @ graph["info"][k].reshape(6, 6) @ e
        return out


def test_candidate_pairs_shape():
    from rgbdslam_v2_b200 import pipeline
    pairs = pipeline.candidate_pairs(40, seed=1)
    by_new = {}
    for a, b in pairs:
        assert 0 <= b < a < 40
        by_new.setdefault(a, []).append(b)
    assert by_new[1] == [0] and by_new[5][:3] == [4, 3, 2]
    assert all(len(v) == len(set(v)) and len(v) <= 3 + 4 + 4 for v in by_new.values())
    assert len(by_new[39]) == 11 and min(by_new[39]) < 39 - 8          # sampled loop-closure candidates from the far past


def test_build_graph_rules():
    from rgbdslam_v2_b200 import pipeline
    from rgbdslam_v2_b200._capi import PAIR_RESULT_DTYPE
    pairs = [(1, 0), (2, 1), (2, 0), (3, 2)]
    res = np.zeros(4, PAIR_RESULT_DTYPE)
    for k, (a, b) in enumerate(pairs):
        T = np.eye(4, dtype=np.float32); T[0, 3] = 0.1 * (a - b)
        res[k]["id1"], res[k]["id2"], res[k]["n_inliers"], res[k]["info_scale"] = b, a, 50 + 10 * (a - b), 100.0
        res[k]["ransac_trafo"] = T.T.reshape(-1)
    res[3]["id1"] = res[3]["id2"] = -1                                   # node 3 lost its predecessor
    g = pipeline.build_graph(pairs, res, 4, dt=1 / 30)
    assert g["n_valid_edges"] == 3 and g["n_const_edges"] == 1
    assert np.allclose(g["init"][:, 0], [0, 0.1, 0.2, 0.2])              # vertex 2 from the edge with more inliers, 3 = constant position
    assert np.allclose(g["info"][-1].reshape(6, 6), np.eye(6) * 30) and list(g["ij"][-1]) == [2, 3]
    assert g["fixed"][0] == 1 and g["fixed"][1:].sum() == 0


def test_prune_and_evaluation_sequence_with_the_oracle(oracle_mod):
    from rgbdslam_v2_b200 import pipeline, synth
    g = synth.make_pose_graph(60, 240, seed=3, outlier_frac=0.08)
    levels = pipeline.evaluation_sequence(OracleGraphBackend(oracle_mod), g)
    assert len(levels) == 4 and levels[1][2] > 0                         # outlier loop closures found at chi2 > 5
    ate = [synth.ate_rmse(x[:, :3], g["gt"][:, :3]) for x, _, _ in levels]
    assert ate[-1] <= ate[0] + 1e-3 and ate[-1] < 0.05
    # pruning rules: a non-consecutive edge between well-connected vertices leaves the active set, a consecutive one gets I
    gg = dict(g, meas=g["meas"].copy(), info=g["info"].copy())
    chi = np.zeros(len(g["ij"])); far = int(np.argmax(np.abs(g["ij"][:, 0] - g["ij"][:, 1]) > 1)); near = int(np.argmax(np.abs(g["ij"][:, 0] - g["ij"][:, 1]) == 1))
    chi[far] = chi[near] = 10.0
    assert pipeline.prune_edges(gg, chi, 5.0) == 2
    assert not gg["active"][far] and gg["active"][near]
    assert np.allclose(gg["info"][near].reshape(6, 6), np.eye(6)) and np.allclose(gg["meas"][near], [0, 0, 0, 0, 0, 0, 1])


def test_save_trajectory_roundtrip(tmp_path):
    from rgbdslam_v2_b200 import pipeline
    poses = np.array([[0, 0, 0, 0, 0, 0, 1.0], [0.1, -0.2, 0.3, 0, 0, np.sin(0.2), np.cos(0.2)]])
    f = tmp_path / "t.txt"
    pipeline.save_trajectory(str(f), poses, np.array([1.0, 1.5]))
    rows = np.loadtxt(str(f), comments="#")
    assert rows.shape == (2, 8) and np.allclose(rows[:, 1:], poses, atol=1e-6) and np.allclose(rows[:, 0], [1.0, 1.5])


def test_build_graph_fast_equals_build_graph():
    """the vectorised graph builder of the long-sequence path == the reference-order loop"""
    from rgbdslam_v2_b200 import pipeline
    from rgbdslam_v2_b200._capi import PAIR_RESULT_DTYPE
    rng = np.random.default_rng(4)
    n = 60
    pairs = pipeline.candidate_pairs(n, seed=2)
    res = np.zeros(len(pairs), PAIR_RESULT_DTYPE)
    for i, (a, b) in enumerate(pairs):
        ok = rng.random() < (0.9 if a - b <= 3 else 0.4)
        if a == 17 or a == 30:
            ok = False  # frames without any edge -> constant-position edge
        res[i]["id1"], res[i]["id2"] = (b, a) if ok else (-1, -1)
        ang = rng.normal() * 0.05
        T = np.eye(4); T[:3, :3] = [[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]]; T[:3, 3] = rng.normal(size=3) * 0.1
        res[i]["ransac_trafo"] = T.T.reshape(-1).astype(np.float32)
        res[i]["n_inliers"] = rng.integers(20, 40)  # ties are frequent: the FIRST maximum must win
        res[i]["info_scale"] = rng.uniform(10, 500)
    a = pipeline.build_graph(pairs, res, n)
    b = pipeline.build_graph_fast(np.array(pairs), res, n)
    for k in ("init", "fixed", "ij", "meas", "info"):
        assert np.array_equal(a[k], b[k]), k
    assert a["n_const_edges"] == b["n_const_edges"] and a["n_valid_edges"] == b["n_valid_edges"]
    # the C-ABI host glue (rgbdslam_b200_graph_from_pairs; no GPU needed) builds the same graph
    from rgbdslam_v2_b200._capi import graph_from_pairs
    c = graph_from_pairs(np.array(pairs), res, n)
    assert np.array_equal(a["ij"], c["ij"]) and np.array_equal(a["fixed"], c["fixed"]) and np.array_equal(a["info"], c["info"])
    assert np.abs(a["meas"] - c["meas"]).max() < 1e-15 and np.abs(a["init"] - c["init"]).max() < 1e-13
    assert a["n_const_edges"] == c["n_const_edges"] > 0
