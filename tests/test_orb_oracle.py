"""The reference's own glue around OpenCV's ORB as restated in oracle/orb_oracle.py (feature_adjuster.cpp:85-317, node.cpp:101-240):
grid geometry, adaptive thresholds that persist across frames, per-cell / per-node feature budgets, the depth filter and the
order the extractor leaves behind.  cv2 itself is the arithmetic the reference runs, so it is not re-tested here.  CPU only."""
import numpy as np


def _frames(ks):
    from oracle import orb_oracle
    from rgbdslam_v2_b200 import synth
    poses = synth.trajectory(240)
    out = []
    for k in ks:
        g, d = synth.render_frame(poses[k], seed=k)[:2]
        out.append((g, d, orb_oracle.depth_to_mask(d)))
    return out


def test_grid_cells_overlap_by_the_edge_threshold():
    from oracle import orb_oracle
    cells = orb_oracle._cells(640, 480, 3)
    assert len(cells) == 9 and cells[0] == (0, 160 + 31, 0, 213 + 31) and cells[4] == (160 - 31, 320 + 31, 213 - 31, 426 + 31)
    assert cells[8][1] == 480 and cells[8][3] == 640
    assert orb_oracle._cells(640, 480, 1) == [(0, 480, 0, 640)]


def test_detector_thresholds_adapt_and_persist():
    from oracle import orb_oracle
    (g0, d0, m0), (g1, d1, m1) = _frames([0, 1])
    st = orb_oracle.DetectorState()
    rec0 = orb_oracle.grid_detect(g0, m0, st, max_keypoints=600)
    th0 = list(st.thresh[:9])
    assert any(t != 20.0 for t in th0) and all(2.0 <= t <= 10000.0 for t in th0)     # adjusted away from the initial 20 (features.cpp:92)
    rec1 = orb_oracle.grid_detect(g1, m1, st, max_keypoints=600)
    # a fresh state on frame 1 starts from 20 again: the persisted thresholds are what makes the two runs differ
    fresh = orb_oracle.DetectorState()
    orb_oracle.grid_detect(g1, m1, fresh, max_keypoints=600)
    assert st.thresh[:9] != [20.0] * 9
    per_cell = (600 * 3 // 2) // 9
    for rec in (rec0, rec1):
        counts = np.bincount([r["cell"] for r in rec], minlength=9)
        assert counts.max() <= per_cell and len(rec) <= 9 * per_cell                 # keepStrongest(1.5 K / 9) per cell
        for c in range(9):                                                            # canonical order inside a cell
            resp = [abs(float(r["response"])) for r in rec if r["cell"] == c]
            assert resp == sorted(resp, reverse=True)


def test_node_constructor_invariants():
    from oracle import orb_oracle
    from rgbdslam_v2_b200 import synth
    (g, d, m), = _frames([3])
    K4 = (synth.FX, synth.FY, synth.CX, synth.CY)
    kp, desc, xyz = orb_oracle.node_construct(g, d, m, K4, orb_oracle.DetectorState(), max_keypoints=400)
    n = len(kp)
    assert 0 < n <= 400 and desc.shape == (n, 32) and desc.dtype == np.uint8 and xyz.shape == (n, 4)
    assert not np.isnan(xyz).any() and np.all(xyz[:, 3] == 1.0)                      # removeDepthless: every feature has depth
    assert np.all(np.diff(kp["octave"]) >= 0)                                         # ORB::compute re-orders by octave
    assert kp["x"].min() >= 31 and kp["x"].max() <= 640 - 31 and kp["y"].min() >= 31 and kp["y"].max() <= 480 - 31  # edgeThreshold 31 of the extractor
    # back-projection (misc2.h:62-64) of the rounded pixel's depth
    for i in range(0, n, max(1, n // 20)):
        u, v = int(round(float(kp["x"][i]))), int(round(float(kp["y"][i])))
        z = d[v, u]
        assert xyz[i, 2] == z
        assert abs(xyz[i, 0] - (kp["x"][i] - K4[2]) * z / K4[0]) < 1e-5 and abs(xyz[i, 1] - (kp["y"][i] - K4[3]) * z / K4[1]) < 1e-5


def test_mask_from_depth_rule_matches_cv2():
    """RGBDSLAM_B200_MASK_FROM_DEPTH derives the detection mask on the device as `d * 100.f rounds (half to even) to an int in [1, 2^31)`;
    the same rule in numpy must agree with cv2's conversion (depthToCV8UC1, misc.cpp:414-418) on non-zero-ness."""
    from oracle import orb_oracle
    rng = np.random.default_rng(3)
    d = rng.uniform(0.0, 6.0, (120, 160)).astype(np.float32)
    d[rng.random(d.shape) < 0.1] = np.nan
    d[:2, :8] = np.array([0.0, 0.004, 0.005, 0.0050001, 0.0149, 0.015, 2.55, 100.0], np.float32)
    d[2, :6] = np.array([-0.5, -0.004, np.inf, 1e-30, 3e7, 2e7], np.float32)
    m = orb_oracle.depth_to_mask(d)
    with np.errstate(invalid="ignore"):
        v = d * np.float32(100.0)
        rule = ~np.isnan(v) & (v < np.float32(2147483648.0)) & (np.rint(v) >= 1)
    assert np.array_equal(m != 0, rule)
