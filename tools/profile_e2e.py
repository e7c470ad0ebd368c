"""Where does the end-to-end (host buffers) step spend its time?  (GPU box)"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from rgbdslam_v2_b200 import Frontend, synth
from rgbdslam_v2_b200._capi import default_params, PAIR_RESULT_DTYPE, DMATCH_DTYPE

p = default_params(); p.depth_cov_z0 = 2.0
fe = Frontend(0, p)
b = synth.make_batch(256, 1000, seed0=1)
pin = {k: torch.from_numpy(b[k]).pin_memory() for k in ("desc_newer", "xyz_newer", "desc_older", "xyz_older")}
res = torch.zeros(256 * PAIR_RESULT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
allm = torch.zeros(256 * 300 * 16, dtype=torch.uint8).pin_memory(); inl = torch.zeros_like(allm).pin_memory()
r = res.numpy().view(PAIR_RESULT_DTYPE); a = allm.numpy().view(DMATCH_DTYPE).reshape(256, 300); i = inl.numpy().view(DMATCH_DTYPE).reshape(256, 300)

def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3

def host(out): return lambda: fe.match_pairs_host(pin["desc_newer"], pin["xyz_newer"], b["n_newer"], pin["desc_older"], pin["xyz_older"], b["n_older"], b["id_newer"], b["id_older"], seed=1, out=out)
print("e2e sync, results+matches  ms", t(host((r, a, i))))
print("e2e sync, results only     ms", t(host((r, None, None))))
print("e2e sync, pageable inputs  ms", t(lambda: fe.match_pairs_host(b["desc_newer"], b["xyz_newer"], b["n_newer"], b["desc_older"], b["xyz_older"], b["n_older"], b["id_newer"], b["id_older"], seed=1, out=(r, None, None))))
dev = {k: torch.empty_like(v, device="cuda") for k, v in pin.items()}
print("torch H2D of the 4 buffers ms", t(lambda: [dev[k].copy_(pin[k], non_blocking=True) for k in pin]))
newer = np.array([fe.node_from_features(int(b["id_newer"][k]), q["desc_newer"], q["xyz_newer"]) for k, q in enumerate(b["pairs"])], np.uint64)
older = np.array([fe.node_from_features(int(b["id_older"][k]), q["desc_older"], q["xyz_older"]) for k, q in enumerate(b["pairs"])], np.uint64)
print("resident sync (256-query tc kernel) ms", t(lambda: fe.match_node_pairs(newer, older, seed=1, out=(r, None, None))), "device", fe.last_timing())
fe.set_hamming_path(1)
print("resident sync, 128-query tc kernel ms", t(lambda: fe.match_node_pairs(newer, older, seed=1, out=(r, None, None))), "device", fe.last_timing())
fe.set_hamming_path(0)
print("e2e sync SIMT hamming (no int8 expansion) ms", t(host((r, None, None))))
fe.set_hamming_path(1)
# pipelined e2e, depth 3
outs = []
for j in range(3):
    rr = torch.zeros(256 * PAIR_RESULT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    outs.append((rr, rr.numpy().view(PAIR_RESULT_DTYPE)))
def pipe(K=30):
    for k in range(K):
        fe.submit_pairs_host(1 + k % 3, pin["desc_newer"], pin["xyz_newer"], b["n_newer"], pin["desc_older"], pin["xyz_older"], b["n_older"], b["id_newer"], b["id_older"], (outs[k % 3][1], None, None), seed=1)
    for j in range(3): fe.wait_slot(1 + j)
pipe(6); torch.cuda.synchronize(); t0 = time.perf_counter(); pipe(30); torch.cuda.synchronize()
print("e2e pipelined depth 3, results only ms/step", (time.perf_counter() - t0) / 30 * 1e3)
for j in range(3): print("   slot", 1 + j, fe.stage_times(1 + j))
fe.match_pairs_host(pin["desc_newer"], pin["xyz_newer"], b["n_newer"], pin["desc_older"], pin["xyz_older"], b["n_older"], b["id_newer"], b["id_older"], seed=1, out=(r, None, None))
print("   sync call stage times", fe.stage_times(0))
t0 = time.perf_counter()
for k in range(30):
    fe.submit_pairs_host(1 + k % 3, pin["desc_newer"], pin["xyz_newer"], b["n_newer"], pin["desc_older"], pin["xyz_older"], b["n_older"], b["id_newer"], b["id_older"], (outs[k % 3][1], None, None), seed=1)
host_ms = (time.perf_counter() - t0) / 30 * 1e3
for j in range(3): fe.wait_slot(1 + j)
print("host time per submit (incl. implicit drain) ms", host_ms)
# steady-state timeline (device ms since epoch): submit, h2d done, expand done, match start, match end, ransac end, d2h done
fe.timeline_epoch(); h0 = time.perf_counter()
rows = []
for k in range(24):
    s = 1 + k % 3
    tw0 = time.perf_counter()
    if k >= 3:
        fe.wait_slot(s); rows.append((k - 3, fe.slot_timeline(s).copy()))
    tw1 = time.perf_counter()
    fe.submit_pairs_host(s, pin["desc_newer"], pin["xyz_newer"], b["n_newer"], pin["desc_older"], pin["xyz_older"], b["n_older"], b["id_newer"], b["id_older"], (outs[k % 3][1], None, None), seed=1)
    tw2 = time.perf_counter()
    if 12 <= k < 20: print(f"   host k={k} t={1e3*(tw0-h0):7.3f} wait {1e3*(tw1-tw0):6.3f} submit {1e3*(tw2-tw1):6.3f}")
for j in range(3): fe.wait_slot(1 + j)
print("timeline  k  submit  h2d_done  expand_done  match_start  match_end  ransac_end  d2h_done")
for k, tl in rows[8:18]: print("   ", k, " ".join(f"{v:8.3f}" for v in tl))
# depth-1 split: host time inside submit vs inside wait (slot 0 = library stream, slot 1 = own stream)
for s in (0, 1):
    ts, tw = [], []
    for k in range(12):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        fe.submit_pairs_host(s, pin["desc_newer"], pin["xyz_newer"], b["n_newer"], pin["desc_older"], pin["xyz_older"], b["n_older"], b["id_newer"], b["id_older"], (r, None, None), seed=1)
        t1 = time.perf_counter(); fe.wait_slot(s); t2 = time.perf_counter()
        if k >= 2: ts.append(t1 - t0); tw.append(t2 - t1)
    print(f"depth-1 slot {s}: submit {1e3*np.median(ts):.3f} ms, wait {1e3*np.median(tw):.3f} ms, stages", fe.stage_times(s))
# pipelined with the match lists downloaded too, and with three distinct batches (cold L2) like bench.py
bs = [synth.make_batch(256, 1000, seed0=1 + 256 * j) for j in range(3)]
pins = [{k: torch.from_numpy(bb[k]).pin_memory() for k in ("desc_newer", "xyz_newer", "desc_older", "xyz_older")} for bb in bs]
mm = [(torch.zeros(256 * 300 * 16, dtype=torch.uint8).pin_memory(), torch.zeros(256 * 300 * 16, dtype=torch.uint8).pin_memory()) for _ in range(3)]
def pipe2(K, matches, distinct):
    for k in range(K):
        j = k % 3; bb, pp = (bs[j], pins[j]) if distinct else (b, pin)
        o = (outs[j][1], mm[j][0].numpy().view(DMATCH_DTYPE).reshape(256, 300), mm[j][1].numpy().view(DMATCH_DTYPE).reshape(256, 300)) if matches else (outs[j][1], None, None)
        if k >= 3: fe.wait_slot(1 + j)
        fe.submit_pairs_host(1 + j, pp["desc_newer"], pp["xyz_newer"], bb["n_newer"], pp["desc_older"], pp["xyz_older"], bb["n_older"], bb["id_newer"], bb["id_older"], o, seed=1)
    for j in range(3): fe.wait_slot(1 + j)
for matches in (False, True):
    for distinct in (False, True):
        for K in (20, 200):
            pipe2(6, matches, distinct); torch.cuda.synchronize(); t0 = time.perf_counter(); pipe2(K, matches, distinct); torch.cuda.synchronize()
            print(f"pipelined e2e matches={matches} distinct_batches={distinct} K={K}: {(time.perf_counter() - t0) / K * 1e3:.3f} ms/step")
