"""Oracle of the float-descriptor (SIFT-128) branch -- TEST INFRASTRUCTURE, not product code.

  root_sift            squareroot_descriptor_space (node.cpp:1557-1571)
  knn2_exact           exact 2-NN by squared L2 (what cv::flann::Index::knnSearch returns, node.cpp:1573-1581, when the
                       kd-tree search is exhaustive).  The reference's FLANN kd-tree (4 trees, 16 checks) is approximate;
                       the CUDA path replaces it with an exact search, so the oracle is the exact 2-NN ("semantic superset",
                       SURVEY.md 8a-a10).  float64 arithmetic on the float32 RootSIFT rows.
  feature_matching     node.cpp:638-667: ratio = d1/d2 < nn_distance_ratio, first-come unique trainIdx, distance = ratio,
                       then keepStrongestMatches (node.cpp:674) + the sort of node.cpp:1127.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import oracle as co


def root_sift(desc: np.ndarray) -> np.ndarray:
    d = np.abs(np.asarray(desc, np.float32))
    s = d.sum(1, dtype=np.float32)
    out = d.copy()
    nz = s != 0
    out[nz] = np.sqrt(d[nz] / s[nz, None]).astype(np.float32)
    return out


def knn2_exact(q: np.ndarray, t: np.ndarray):
    q64, t64 = q.astype(np.float64), t.astype(np.float64)
    D = (q64 * q64).sum(1)[:, None] + (t64 * t64).sum(1)[None, :] - 2 * q64 @ t64.T
    D = np.maximum(D, 0)
    idx = np.argsort(D, axis=1, kind="stable")[:, :2]
    d = np.take_along_axis(D, idx, 1)
    return idx.astype(np.int32), d


def feature_matching(q_root, t_root, nn_ratio=0.95, max_matches=300):
    idx, d = knn2_exact(q_root, t_root)
    out = []
    seen = set()
    for i in range(len(q_root)):
        ratio = np.float32(np.float32(d[i, 0]) / np.float32(d[i, 1]))
        if nn_ratio > ratio:
            t = int(idx[i, 0])
            if t in seen:
                continue
            seen.add(t)
            out.append((i, t, -1, ratio))
    m = np.array(out, dtype=co.DMATCH_DTYPE) if out else np.zeros(0, co.DMATCH_DTYPE)
    order = np.lexsort((m["queryIdx"], m["distance"]))
    return m[order][:max_matches]


def match_node_pair(params, desc_newer, xyz_newer, id_newer, desc_older, xyz_older, id_older, seed, pair, nn_ratio=0.95,
                    use_root_sift=True):
    qn = root_sift(desc_newer) if use_root_sift else np.asarray(desc_newer, np.float32)
    tn = root_sift(desc_older) if use_root_sift else np.asarray(desc_older, np.float32)
    m = np.ascontiguousarray(feature_matching(qn, tn, nn_ratio, params.max_matches))
    res = np.zeros(1, co.RESULT_DTYPE)
    inl = np.zeros(max(params.max_matches, 1), co.DMATCH_DTYPE)
    x1 = np.ascontiguousarray(xyz_newer, np.float32); x2 = np.ascontiguousarray(xyz_older, np.float32)
    co.lib().oracle_match_node_pair_from_matches(C.byref(params), co._p(x1), C.c_int(id_newer), co._p(x2), C.c_int(id_older),
                                                 co._p(m) if len(m) else None, C.c_int(len(m)), C.c_uint64(seed), C.c_uint64(pair),
                                                 co._p(res), co._p(inl))
    return res[0], m, inl[: res[0]["n_inliers"]]
