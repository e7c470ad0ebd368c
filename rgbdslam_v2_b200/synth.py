"""Seeded synthetic inputs for the frame-pair hot path (numpy only; shared by tests and bench.py).

Feature-level generator: what ``Node`` holds after construction (node.h:167-174) -- ORB descriptors
(N x 32 B) and back-projected points (N x (x,y,z,1), camera frame, metres) -- for a pair of frames that
see an overlapping set of scene points under a small rigid motion, with descriptor bit noise, depth
noise and non-overlapping (outlier) features.  Pinhole model = the reference defaults
fx=fy=525, cx=319.5, cy=239.5 (graph_manager.cpp:189-192), 640x480.
"""
from __future__ import annotations

import numpy as np

FX = FY = 525.0
CX, CY = 319.5, 239.5
W, H = 640, 480


def random_rigid(rng: np.random.Generator, max_trans: float, max_rot_deg: float) -> np.ndarray:
    axis = rng.normal(size=3)
    axis /= np.linalg.norm(axis)
    ang = np.deg2rad(rng.uniform(0.2, 1.0) * max_rot_deg)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    R = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)
    t = rng.normal(size=3)
    t *= rng.uniform(0.2, 1.0) * max_trans / np.linalg.norm(t)
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = t
    return T


def _random_points(rng, n):
    u = rng.uniform(31, W - 31, n)
    v = rng.uniform(31, H - 31, n)
    z = rng.uniform(0.8, 4.0, n)
    return np.stack([(u - CX) * z / FX, (v - CY) * z / FY, z], 1)


def make_pair(seed: int, n_kp: int = 1000, overlap: float | None = None, max_trans: float = 0.05,
              max_rot_deg: float = 2.0, depth_noise: float = 0.0015):
    """One frame pair.  Returns dict with desc_newer/xyz_newer/desc_older/xyz_older and T_true
    (maps newer-frame points into the older frame, the direction of MatchingResult::ransac_trafo)."""
    rng = np.random.default_rng(seed)
    if overlap is None:
        overlap = rng.uniform(0.08, 0.8)
    n_common = int(round(overlap * n_kp))
    T = random_rigid(rng, max_trans, max_rot_deg)  # newer -> older
    # older frame
    p_old = _random_points(rng, n_kp)
    d_old = rng.integers(0, 256, size=(n_kp, 32), dtype=np.uint8)
    # common points, seen from the newer frame: p_new = T^-1 p_old
    Tinv = np.linalg.inv(T)
    common_idx = rng.permutation(n_kp)[:n_common]
    p_new_common = p_old[common_idx] @ Tinv[:3, :3].T + Tinv[:3, 3]
    # descriptor bit noise: per-feature flip probability in [0.02, 0.2]
    flip_p = rng.uniform(0.02, 0.2, n_common)
    flips = rng.random((n_common, 256)) < flip_p[:, None]
    d_new_common = d_old[common_idx] ^ np.packbits(flips, axis=1, bitorder="little")
    # unrelated features
    n_rest = n_kp - n_common
    p_new_rest = _random_points(rng, n_rest)
    d_new_rest = rng.integers(0, 256, size=(n_rest, 32), dtype=np.uint8)
    p_new = np.concatenate([p_new_common, p_new_rest])
    d_new = np.concatenate([d_new_common, d_new_rest])
    perm = rng.permutation(n_kp)
    p_new, d_new = p_new[perm], d_new[perm]
    # sensor noise: sigma_z = depth_noise * z^2 along the ray (both frames)
    for p in (p_old, p_new):
        z = p[:, 2].copy()
        zn = z + rng.normal(size=len(z)) * depth_noise * z * z
        p *= (zn / z)[:, None]
    to4 = lambda p: np.concatenate([p, np.ones((len(p), 1))], 1).astype(np.float32)
    return dict(desc_newer=np.ascontiguousarray(d_new), xyz_newer=to4(p_new), desc_older=np.ascontiguousarray(d_old),
                xyz_older=to4(p_old), T_true=T, n_common=n_common)


def make_batch(npairs: int, n_kp: int = 1000, seed0: int = 0, **kw):
    """Concatenated host buffers for match_pairs_host / the oracle batch driver."""
    pairs = [make_pair(seed0 + i, n_kp, **kw) for i in range(npairs)]
    cat = lambda k: np.ascontiguousarray(np.concatenate([p[k] for p in pairs]))
    return dict(
        desc_newer=cat("desc_newer"), xyz_newer=cat("xyz_newer"), desc_older=cat("desc_older"), xyz_older=cat("xyz_older"),
        n_newer=np.full(npairs, n_kp, np.int32), n_older=np.full(npairs, n_kp, np.int32),
        id_newer=np.arange(npairs, dtype=np.int32) + 1, id_older=np.arange(npairs, dtype=np.int32),
        T_true=np.stack([p["T_true"] for p in pairs]), n_common=np.array([p["n_common"] for p in pairs]),
        pairs=pairs)
