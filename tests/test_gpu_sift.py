"""GPU tests of the SIFT-128 float-descriptor path (BASELINE config C3): RootSIFT, exact 2-NN through the bf16
tensor-core score matrix + fp32 re-ranking, ratio / uniqueness matching, RANSAC."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fe(built):
    from rgbdslam_v2_b200 import Frontend
    from rgbdslam_v2_b200._capi import default_params
    p = default_params()
    p.depth_cov_z0 = 2.0
    f = Frontend(0, p)
    yield f
    f.close()


def _sift_like(rng, n):
    """Non-negative, sparse-ish 128-d vectors with SIFT-like statistics (values 0..255, many small)."""
    d = rng.gamma(0.6, 30.0, size=(n, 128)).astype(np.float32)
    return np.minimum(d, 255.0).astype(np.float32)


@pytest.mark.parametrize("nq,nt", [(1, 2), (100, 300), (2000, 2000), (1025, 700), (4096, 4096)])
def test_knn2_matches_exact_search(fe, nq, nt):
    from oracle import sift_oracle
    rng = np.random.default_rng(nq + nt)
    q, t = _sift_like(rng, nq), _sift_like(rng, nt)
    k = min(nq, nt) // 2
    if k:
        q[:k] = np.maximum(t[rng.permutation(nt)[:k]] + rng.normal(0, 6.0, (k, 128)).astype(np.float32), 0)  # true matches
    idx, d = fe.knn2_l2(q, t)
    idx_b, d_b = fe.knn2_l2(q, t)
    assert np.array_equal(idx, idx_b) and np.array_equal(d, d_b)  # deterministic
    qr, tr = sift_oracle.root_sift(q), sift_oracle.root_sift(t)
    oidx, od = sift_oracle.knn2_exact(qr, tr)
    # distances of the returned neighbours are exact fp32 evaluations: tolerance 1e-5 absolute on squared L2 <= 2
    ex = ((qr[:, None, :].astype(np.float64) - tr[idx].astype(np.float64)) ** 2).sum(2) if nq * 2 * 128 < 5e7 else None
    if ex is not None:
        assert np.abs(ex - d).max() < 2e-5
    same1 = (idx[:, 0] == oidx[:, 0]).mean()
    assert same1 > 0.995, same1  # bf16 candidate generation + exact re-ranking finds the true nearest neighbour
    # where the index differs the distance is (nearly) tied
    bad = idx[:, 0] != oidx[:, 0]
    assert (np.abs(d[bad, 0] - od[bad, 0]) < 2e-3 * np.maximum(od[bad, 0], 1e-3)).all()
    assert (d[:, 0] <= d[:, 1] + 1e-7).all()
    assert np.abs(d[~bad, 0] - od[~bad, 0]).max() < 2e-5


def test_sift_match_pairs_vs_oracle(fe, oracle_mod):
    from oracle import sift_oracle
    from rgbdslam_v2_b200 import synth
    rng = np.random.default_rng(5)
    prm = oracle_mod.make_params(depth_cov_z0=2.0)
    newer, older, exp = [], [], []
    for i in range(6):
        b = synth.make_pair(800 + i, 1500, overlap=0.5)
        do = _sift_like(rng, 1500)
        dn = _sift_like(rng, 1500)
        # reuse the generator's geometry: common points are those whose binary descriptors are near-duplicates
        hd, idx = oracle_mod.brute_force_orb(b["desc_newer"], b["desc_older"])
        good = hd < 60
        dn[good] = np.maximum(do[idx[good]] + rng.normal(0, 5.0, (good.sum(), 128)).astype(np.float32), 0)
        newer.append(fe.node_from_sift(100 + i, dn, b["xyz_newer"]))
        older.append(fe.node_from_sift(i, do, b["xyz_older"]))
        exp.append(sift_oracle.match_node_pair(prm, dn, b["xyz_newer"], 100 + i, do, b["xyz_older"], i, seed=17, pair=i))
    res, allm, inl = fe.match_node_pairs(newer, older, seed=17)
    for i, (ores, om, oinl) in enumerate(exp):
        n = int(res[i]["n_all_matches"])
        assert abs(n - len(om)) <= 2
        common = np.intersect1d(allm[i, :n]["queryIdx"], om["queryIdx"])
        assert len(common) >= 0.99 * len(om)
        assert res[i]["id1"] == ores["id1"] and res[i]["id2"] == ores["id2"] and res[i]["id1"] == i
        T, To = res[i]["ransac_trafo"].reshape(4, 4).T, ores["ransac_trafo"].reshape(4, 4).T
        assert np.abs(T[:3, 3] - To[:3, 3]).max() < 2e-3 and np.abs(T[:3, :3] - To[:3, :3]).max() < 2e-3
        assert abs(int(res[i]["n_inliers"]) - int(ores["n_inliers"])) <= max(3, 0.03 * ores["n_inliers"])
        assert (np.diff(allm[i, :n]["distance"]) >= 0).all() and (allm[i, :n]["distance"] < 0.95).all()
        assert len(np.unique(allm[i, :n]["trainIdx"])) == n  # uniqueness of trainIdx (node.cpp:655-658)
    for h in newer + older:
        fe.node_destroy(h)


def test_mixing_orb_and_sift_nodes_is_rejected(fe):
    from rgbdslam_v2_b200 import synth
    from rgbdslam_v2_b200._capi import B200Error
    b = synth.make_pair(1, 100)
    a = fe.node_from_features(1, b["desc_newer"], b["xyz_newer"])
    s = fe.node_from_sift(2, np.ones((100, 128), np.float32), b["xyz_older"])
    with pytest.raises(B200Error):
        fe.match_node_pairs([a], [s])
    fe.node_destroy(a); fe.node_destroy(s)


def _unit_sift(rng, n):
    """SiftGPU-style descriptors: non-negative, unit L2 norm, clipped at 0.2 and renormalised (values <= ~0.5)."""
    d = rng.gamma(0.6, 1.0, size=(n, 128)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d = np.minimum(d, 0.2)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return d.astype(np.float32)


@pytest.mark.parametrize("nq,nt", [(1, 1), (40, 33), (700, 1000), (2000, 2000), (1300, 257)])
def test_siftgpu_matcher_is_bit_exact(fe, nq, nt):
    """matcher_type SIFTGPU (node.cpp:553-557): u8 dot-product matrix on the tensor cores, acos distance / ratio tests,
    mutual best match, float L2 DMatch.distance -- all integer or order-pinned float work, so the match list must equal
    the restated SiftMatchGPU + SiftGPUWrapper::match exactly (indices and distance bits)."""
    from oracle import sift_oracle
    rng = np.random.default_rng(nq * 7 + nt)
    t = _unit_sift(rng, nt)
    q = _unit_sift(rng, nq)
    k = min(nq, nt) * 2 // 3
    if k:
        q[:k] = np.abs(t[rng.permutation(nt)[:k]] + rng.normal(0, 0.01, (k, 128)).astype(np.float32))
    if nt > 40 and nq > 40:
        t[37] = t[5]; t[9] = t[5]          # duplicated train rows: row-pass ties resolved by (col % 32, col)
        q[3] = t[5]
        q[20] = q[21]                      # duplicated query rows: column-pass ties resolved by the lowest row
        q[30, :] = 0.0                     # all-zero descriptor: no positive dot product, never matched
        t[12, 0] = 0.5                     # 512 * 0.5 + 0.5 = 256 wraps to 0 in the unsigned char
    xyz_q = np.concatenate([rng.uniform(0.5, 3, (nq, 3)), np.ones((nq, 1))], 1).astype(np.float32)
    xyz_t = np.concatenate([rng.uniform(0.5, 3, (nt, 3)), np.ones((nt, 1))], 1).astype(np.float32)
    fe.set_sift_matcher(1)
    try:
        a, b = fe.node_from_sift(1, q, xyz_q), fe.node_from_sift(0, t, xyz_t)
    finally:
        fe.set_sift_matcher(0)
    res, allm, _ = fe.match_node_pairs([a], [b], seed=3)
    exp = sift_oracle.siftgpu_feature_matching(q, t, max_matches=300)
    n = int(res[0]["n_all_matches"])
    assert n == len(exp)
    got = allm[0, :n]
    assert np.array_equal(got["queryIdx"], exp["queryIdx"]) and np.array_equal(got["trainIdx"], exp["trainIdx"])
    assert np.array_equal(got["distance"].view(np.uint32), exp["distance"].view(np.uint32))
    if nq > 40 and nt > 40:
        assert n > 20
    fe.node_destroy(a); fe.node_destroy(b)


def test_siftgpu_and_ratio_nodes_do_not_mix(fe):
    from rgbdslam_v2_b200._capi import B200Error
    rng = np.random.default_rng(1)
    d = _unit_sift(rng, 64)
    xyz = np.concatenate([rng.uniform(0.5, 3, (64, 3)), np.ones((64, 1))], 1).astype(np.float32)
    a = fe.node_from_sift(1, d, xyz)
    fe.set_sift_matcher(1)
    try:
        b = fe.node_from_sift(0, d, xyz)
    finally:
        fe.set_sift_matcher(0)
    with pytest.raises(B200Error):
        fe.match_node_pairs([a], [b])
    fe.node_destroy(a); fe.node_destroy(b)
