"""GPU parity tests of the pose-graph solve (GraphManager::optimizeGraph) vs the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fe(built):
    from rgbdslam_v2_b200 import Frontend
    f = Frontend(0)
    yield f
    f.close()


def test_chi2_and_per_edge_chi2(fe, oracle_mod):
    from rgbdslam_v2_b200 import synth
    g = synth.make_pose_graph(300, 1500, seed=2, outlier_frac=0.03)
    chi, pe = fe.graph_chi2(g["init"], g["ij"], g["meas"], g["info"], per_edge=True)
    ochi, orob = oracle_mod.posegraph_chi2(g["init"], g["ij"], g["meas"], g["info"])
    assert chi == pytest.approx(ochi, rel=1e-10)
    assert pe.sum() == pytest.approx(ochi, rel=1e-10)
    e = oracle_mod.edge_se3(g["init"][g["ij"][7, 0]], g["init"][g["ij"][7, 1]], g["meas"][7], False)[0]
    assert pe[7] == pytest.approx(e @ g["info"][7].reshape(6, 6) @ e, rel=1e-9)


@pytest.mark.parametrize("nv,ne,stop,seed", [(200, 1000, 0.01, 1), (500, 3000, 0.01, 4), (500, 3000, 20.0, 4),
                                             (64, 63, 0.01, 7), (400, 2500, 1e-4, 9)])
def test_optimize_vs_oracle(fe, oracle_mod, nv, ne, stop, seed):
    from rgbdslam_v2_b200 import synth
    g = synth.make_pose_graph(nv, ne, seed=seed)
    x, chi2, it, cg = fe.optimize_graph(g["init"], g["fixed"], g["ij"], g["meas"], g["info"], stop=stop)
    ox, ochi2, oit, ocg = oracle_mod.posegraph_optimize(g["init"], g["fixed"], g["ij"], g["meas"], g["info"], stop=stop)
    c0, _ = oracle_mod.posegraph_chi2(g["init"], g["ij"], g["meas"], g["info"])
    assert chi2 < c0 or ne == nv - 1
    # tolerance: converged chi2 within 1e-6 relative, poses within 1e-6 m / 1e-6 (same LM/PCG, different summation order)
    assert chi2 == pytest.approx(ochi2, rel=1e-6, abs=1e-9)
    assert abs(it - oit) <= 2  # the final 'nothing left to do' iteration (PCG starts below tolerance) is borderline
    assert np.abs(x[:, :3] - ox[:, :3]).max() < 1e-6
    sgn = np.sign((x[:, 3:] * ox[:, 3:]).sum(1))[:, None]
    assert np.abs(x[:, 3:] - sgn * ox[:, 3:]).max() < 1e-6
    assert np.array_equal(x[0], g["init"][0])  # fixed vertex (graph_manager.cpp:911-937 "first")
    # the returned chi2 is the chi2 of the returned poses
    assert fe.graph_chi2(x, g["ij"], g["meas"], g["info"]) == pytest.approx(chi2, rel=1e-9)


def test_huber_outliers_and_ate(fe, oracle_mod):
    from rgbdslam_v2_b200 import synth
    g = synth.make_pose_graph(600, 3600, seed=11, outlier_frac=0.05)
    x, chi2, it, cg = fe.optimize_graph(g["init"], g["fixed"], g["ij"], g["meas"], g["info"], stop=0.01)
    ox, ochi2, oit, _ = oracle_mod.posegraph_optimize(g["init"], g["fixed"], g["ij"], g["meas"], g["info"], stop=0.01)
    ate, oate = synth.ate_rmse(x[:, :3], g["gt"][:, :3]), synth.ate_rmse(ox[:, :3], g["gt"][:, :3])
    assert abs(ate - oate) < 1e-3  # north star: trajectory ATE RMSE within 1 mm of the reference
    assert synth.ate_rmse(x[:, :3], ox[:, :3]) < 1e-3
    assert ate < 0.05 and chi2 == pytest.approx(ochi2, rel=1e-4)


def test_several_fixed_vertices_and_multi_edges(fe, oracle_mod):
    from rgbdslam_v2_b200 import synth
    g = synth.make_pose_graph(120, 500, seed=13)
    g["fixed"][[0, 40, 80]] = 1
    g["init"][[40, 80]] = g["gt"][[40, 80]]
    ij = np.concatenate([g["ij"], g["ij"][:50]])  # duplicate edges between the same vertices are separate constraints
    meas = np.concatenate([g["meas"], g["meas"][:50]]); info = np.concatenate([g["info"], g["info"][:50]])
    x, chi2, it, _ = fe.optimize_graph(g["init"], g["fixed"], ij, meas, info, stop=0.001)
    ox, ochi2, oit, _ = oracle_mod.posegraph_optimize(g["init"], g["fixed"], ij, meas, info, stop=0.001)
    assert np.array_equal(x[[0, 40, 80]], g["init"][[0, 40, 80]])
    assert chi2 == pytest.approx(ochi2, rel=1e-6) and np.abs(x[:, :3] - ox[:, :3]).max() < 1e-6


def test_c5_size_properties(fe, oracle_mod):
    """BASELINE config C5: 5000 vertices / 30 000 edges."""
    from rgbdslam_v2_b200 import synth
    g = synth.make_pose_graph(5000, 30000, seed=0)
    c0 = fe.graph_chi2(g["init"], g["ij"], g["meas"], g["info"])
    x, chi2, it, cg = fe.optimize_graph(g["init"], g["fixed"], g["ij"], g["meas"], g["info"], stop=0.01)
    assert chi2 < 1e-3 * c0 and it >= 5 and cg > 0
    assert np.array_equal(x[0], g["init"][0])
    assert np.abs(np.linalg.norm(x[:, 3:], axis=1) - 1).max() < 1e-12
    ate0, ate = synth.ate_rmse(g["init"][:, :3], g["gt"][:, :3]), synth.ate_rmse(x[:, :3], g["gt"][:, :3])
    assert ate < 0.05 and ate < 0.05 * ate0
    # the CPU oracle stops at the same point of the same stop rule (graph_manager.cpp:1006-1014)
    ox, ochi2, oit, _ = oracle_mod.posegraph_optimize(g["init"], g["fixed"], g["ij"], g["meas"], g["info"], stop=0.01)
    assert abs(ate - synth.ate_rmse(ox[:, :3], g["gt"][:, :3])) < 1e-3 and synth.ate_rmse(x[:, :3], ox[:, :3]) < 1e-3
    assert chi2 == pytest.approx(ochi2, rel=1e-4) and abs(it - oit) <= 2
    # idempotence: optimising the optimum again changes nothing measurable
    x2, chi2b, _, _ = fe.optimize_graph(x, g["fixed"], g["ij"], g["meas"], g["info"], stop=0.01)
    assert chi2b <= chi2 * (1 + 1e-9) and np.abs(x2[:, :3] - x[:, :3]).max() < 1e-4
