"""Time the match-selection + RANSAC stage of every A/B library under build_variants/ (one subprocess per variant)."""
import os, subprocess, sys, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, str(ROOT))
    import numpy as np, torch, ctypes as C
    from rgbdslam_v2_b200 import Frontend, synth
    from rgbdslam_v2_b200._capi import default_params, PAIR_RESULT_DTYPE
    prm = default_params(); prm.depth_cov_z0 = 2.0; prm.max_keypoints = 1000
    fe = Frontend(0, prm)
    b = synth.make_batch(256, 1000, seed0=1234)
    newer = np.array([fe.node_from_features(int(b["id_newer"][k]), q["desc_newer"], q["xyz_newer"]) for k, q in enumerate(b["pairs"])], np.uint64)
    older = np.array([fe.node_from_features(int(b["id_older"][k]), q["desc_older"], q["xyz_older"]) for k, q in enumerate(b["pairs"])], np.uint64)
    r = np.zeros(256, PAIR_RESULT_DTYPE)
    out = {}
    for path in (1, 2):
        fe.set_hamming_path(path)
        st = []
        for k in range(13):
            fe.match_node_pairs(newer, older, seed=1, out=(r, None, None))
            if k >= 3: st.append(fe.stage_times(0))
        out[f"path{path}"] = {k: round(float(np.median([s[k] for s in st])) * 1e3, 1) for k in ("hamming", "select_ransac", "total")}
    # pipelined resident throughput, 3 slots x 3 distinct batches (what bench.py's `value` measures)
    import time
    sets = []
    for j in range(3):
        bb = synth.make_batch(256, 1000, seed0=5000 + 256 * j)
        nw = np.array([fe.node_from_features(int(bb["id_newer"][k]), q["desc_newer"], q["xyz_newer"]) for k, q in enumerate(bb["pairs"])], np.uint64)
        od = np.array([fe.node_from_features(int(bb["id_older"][k]), q["desc_older"], q["xyz_older"]) for k, q in enumerate(bb["pairs"])], np.uint64)
        sets.append((nw, od, np.zeros(256, PAIR_RESULT_DTYPE)))
    def pipe(K):
        for k in range(K):
            j = k % 3
            if k >= 3: fe.wait_slot(1 + j)
            fe.submit_node_pairs(1 + j, sets[j][0], sets[j][1], (sets[j][2], None, None), seed=1)
        for j in range(3): fe.wait_slot(1 + j)
    fe.set_hamming_path(1)
    pipe(9); torch.cuda.synchronize(); t0 = time.perf_counter(); pipe(60); torch.cuda.synchronize()
    out["pipelined_us_per_step"] = round((time.perf_counter() - t0) / 60 * 1e6, 1)
    out["valid"] = int((r["id1"] >= 0).sum()); out["inl_sum"] = int(r["n_inliers"].sum()); out["rmse_sum"] = float(r["rmse"].sum())
    if hasattr(fe.lib, "rb200_debug_ransac_profile"):
        buf = (C.c_ulonglong * 24)()
        fe.lib.rb200_debug_ransac_profile(buf, 1)
        fe.match_node_pairs(newer, older, seed=1, out=(r, None, None))
        fe.lib.rb200_debug_ransac_profile(buf, 0)
        a = np.array(buf[:], dtype=np.float64).reshape(3, 8)
        for ph in range(3):
            n = max(a[ph, 3], 1)
            out[f"prof_phase{ph}"] = dict(hyps=int(a[ph, 3]), rounds_avg=round(a[ph, 2] / n, 2), rounds_max=int(a[ph, 6]), fit_cyc_per_round=round(a[ph, 0] / max(a[ph, 2], 1)),
                                           score_cyc_per_round=round(a[ph, 1] / max(a[ph, 2] , 1)), loop_cyc_avg=round(a[ph, 4] / n), loop_cyc_max=int(a[ph, 5]))
    print("RESULT", json.dumps(out))
    sys.exit(0)
for lib in sorted((ROOT / "build_variants").glob("librgbdslam_b200.*.so")):
    env = dict(os.environ, RGBDSLAM_B200_LIB=str(lib))
    res = subprocess.run([sys.executable, __file__, "--child"], env=env, capture_output=True, text=True)
    line = [l for l in res.stdout.splitlines() if l.startswith("RESULT")]
    print(lib.name.split(".")[1], line[0][7:] if line else ("FAILED " + res.stderr[-800:]))
