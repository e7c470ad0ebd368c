"""A few passes of the C3 workload (SIFT-128, 2000 kp, 32 pairs, both matchers) -- target of ncu captures."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from rgbdslam_v2_b200 import Frontend, synth
from rgbdslam_v2_b200._capi import default_params
p = default_params(); p.depth_cov_z0 = 2.0
fe = Frontend(0, p)
for matcher, kind in ((0, "sift"), (1, "siftgpu")):
    fe.set_sift_matcher(matcher)
    pairs = [synth.make_pair_sift(9000 + k, 2000, overlap=0.5, kind=kind) for k in range(32)]
    newer = [fe.node_from_sift(2 * k + 1, q["desc_newer"], q["xyz_newer"]) for k, q in enumerate(pairs)]
    older = [fe.node_from_sift(2 * k, q["desc_older"], q["xyz_older"]) for k, q in enumerate(pairs)]
    for it in range(3):
        res, _, _ = fe.match_node_pairs(newer, older, seed=5, want_matches=False)
    print(matcher, fe.stage_times(0), int((res["id1"] >= 0).sum()))
