"""world_size-2 gloo test (CPU) of the multi-GPU host logic: pair sharding + edge all-gather reproduce the
single-process edge list (the per-pair work itself is the GPU path, tested with -m gpu)."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent

WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, os.environ["REPO_ROOT"])
import torch.distributed as dist
from rgbdslam_v2_b200 import sharding
from rgbdslam_v2_b200._capi import PAIR_RESULT_DTYPE
dist.init_process_group("gloo", init_method="env://")
rank, world = dist.get_rank(), dist.get_world_size()
n = 37
allrec = np.zeros(n, PAIR_RESULT_DTYPE)
rng = np.random.default_rng(0)
allrec["id1"] = np.arange(n); allrec["id2"] = np.arange(n) + 1
allrec["rmse"] = rng.random(n).astype(np.float32); allrec["n_inliers"] = rng.integers(0, 300, n)
allrec["ransac_trafo"] = rng.random((n, 16)).astype(np.float32)
rg = sharding.shard_range(n, rank, world)
merged = sharding.allgather_edges_torch(allrec[rg.start:rg.stop], n, world)
assert merged.tobytes() == allrec.tobytes(), "gathered edges differ from the single-process list"
dist.barrier()
if rank == 0:
    print("GLOO_SHARDING_OK", len(rg))
dist.destroy_process_group()
'''


def test_shard_ranges_cover_and_balance():
    from rgbdslam_v2_b200 import sharding
    for n in (0, 1, 7, 256, 1001):
        for w in (1, 2, 3, 8):
            rs = [sharding.shard_range(n, r, w) for r in range(w)]
            assert sum(len(r) for r in rs) == n
            assert [i for r in rs for i in r] == list(range(n))
            assert max(len(r) for r in rs) - min(len(r) for r in rs) <= 1


def test_frame_shards_are_contiguous_blocks():
    """rgbdslam_b200_nodes_create_sharded: rank r owns frames [r * per, min((r + 1) * per, total)), per = ceil(total / world)"""
    from rgbdslam_v2_b200 import sharding
    for n in (0, 1, 7, 250, 2000, 2001):
        for w in (1, 2, 3, 4, 8):
            rs = [sharding.frame_shard(n, r, w) for r in range(w)]
            assert [i for r in rs for i in r] == list(range(n))
            per = -(-n // w) if n else 0
            assert all(len(r) <= per for r in rs) and all(r.start == min(k * per, n) for k, r in enumerate(rs))


def test_pad_and_merge_roundtrip():
    from rgbdslam_v2_b200 import sharding
    from rgbdslam_v2_b200._capi import PAIR_RESULT_DTYPE
    n, w = 11, 4
    rec = np.zeros(n, PAIR_RESULT_DTYPE)
    rec["id1"] = np.arange(n)
    size = sharding.padded_shard_size(n, w)
    gathered = np.concatenate([sharding.pad_edges(rec[r.start:r.stop], size) for r in (sharding.shard_range(n, k, w) for k in range(w))])
    assert (gathered["id1"] == -1).sum() == w * size - n
    assert sharding.merge_gathered(gathered, n, w).tobytes() == rec.tobytes()


def test_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, REPO_ROOT=str(ROOT))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "GLOO_SHARDING_OK" in r.stdout
