"""Host cost of one pipelined C2 step: wall time the calling thread spends inside submit (table staging + launches) and inside
wait (blocking on the slot's events + result copy), DEPTH batches in flight, resident nodes.  If submit + everything else the
host does per step approaches the device time per step (0.137 ms), several ranks sharing a few host cores scale badly."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from rgbdslam_v2_b200 import Frontend, synth
from rgbdslam_v2_b200._capi import default_params, PAIR_RESULT_DTYPE

DEPTH = 5
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 400
prm = default_params(); prm.depth_cov_z0 = 2.0
fe = Frontend(0, prm)
sets = []
for j in range(DEPTH):
    b = synth.make_batch(256, 1000, seed0=1234 + 256 * j)
    newer = np.array([fe.node_from_features(int(b["id_newer"][i]), p["desc_newer"], p["xyz_newer"]) for i, p in enumerate(b["pairs"])], np.uint64)
    older = np.array([fe.node_from_features(int(b["id_older"][i]), p["desc_older"], p["xyz_older"]) for i, p in enumerate(b["pairs"])], np.uint64)
    import torch
    keep = torch.zeros(256 * PAIR_RESULT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()  # pinned: the result copy stays asynchronous
    sets.append((newer, older, keep.numpy().view(PAIR_RESULT_DTYPE), keep))
def run(K):
    t_sub = t_wait = 0.0
    t0 = time.perf_counter()
    for k in range(K):
        if k >= DEPTH:
            a = time.perf_counter(); fe.wait_slot(1 + (k - DEPTH) % DEPTH); t_wait += time.perf_counter() - a
        n, o, r, _ = sets[k % DEPTH]
        a = time.perf_counter(); fe.submit_node_pairs(1 + k % DEPTH, n, o, (r, None, None), seed=1, first_pair_index=256 * (k % DEPTH)); t_sub += time.perf_counter() - a
    for k in range(max(0, K - DEPTH), K):
        a = time.perf_counter(); fe.wait_slot(1 + k % DEPTH); t_wait += time.perf_counter() - a
    tot = time.perf_counter() - t0
    return tot, t_sub, t_wait
run(50)
tot, ts, tw = run(steps)
import os
print(f"{steps} steps: {1e3 * tot / steps:.4f} ms/step wall; inside submit {1e6 * ts / steps:.1f} us/step, inside wait {1e6 * tw / steps:.1f} us/step; "
      f"process CPU time {1e6 * (os.times().user + os.times().system) / 1:.0f} us total")
c0 = os.times(); tot, ts, tw = run(steps); c1 = os.times()
print(f"second run: {1e3 * tot / steps:.4f} ms/step wall, CPU (user+sys, all threads) {1e6 * ((c1.user - c0.user) + (c1.system - c0.system)) / steps:.1f} us/step")
