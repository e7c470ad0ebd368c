"""oracle/emm_oracle.c (environment measurement model, misc.cpp:814-969) against an independent, vectorised numpy restatement
of the same formulas -- two implementations, one in C with loops and one with arrays, must give the same counts.  CPU only."""
import math

import numpy as np


def _numpy_direction(T, src_z, srcK, dst_z, dstK, cov, cloud_step=2, skip=8):
    ch, cw = src_z.shape
    ys, xs = np.meshgrid(np.arange(0, ch, skip), np.arange(0, cw, skip), indexing="ij")
    all_pts = xs.size
    Z = src_z[ys, xs].astype(np.float32)
    u = (xs * cloud_step).astype(np.float32); v = (ys * cloud_step).astype(np.float32)
    fxinv = np.float32(1.0 / srcK[0]); fyinv = np.float32(1.0 / srcK[1])
    x = (u - np.float32(srcK[2])) * Z * fxinv; y = (v - np.float32(srcK[3])) * Z * fyinv
    T = T.astype(np.float32)
    px = T[0, 0] * x + T[0, 1] * y + T[0, 2] * Z + T[0, 3]
    py = T[1, 0] * x + T[1, 1] * y + T[1, 2] * Z + T[1, 3]
    pz = T[2, 0] * x + T[2, 1] * y + T[2, 2] * Z + T[2, 3]
    fx, fy, cx, cy = (np.float32(k) / np.float32(cloud_step) for k in dstK)
    good = bad = occl = 0
    och, ocw = dst_z.shape
    for a, b, c in zip(px.ravel(), py.ravel(), pz.ravel()):
        if not (c == c) or c < 0:
            continue
        rx = int(math.floor(float(np.float32(np.float32(a / c) * fx + cx)) + 0.5)); ry = int(math.floor(float(np.float32(np.float32(b / c) * fy + cy)) + 0.5))
        if rx >= ocw or rx < 0 or ry >= och or ry < 0:
            continue
        nb = dst_z[max(0, ry - 2):min(och, ry + 3):2, max(0, rx - 2):min(ocw, rx + 3):2].ravel().astype(np.float64)
        nb = nb[~np.isnan(nb)]
        if len(nb) == 0:
            continue
        sigma = math.sqrt(cloud_step * cov + cloud_step * cov)
        p = np.array([0.5 * (1 + math.erf((oz - float(c)) / (sigma * 1.41421))) for oz in nb])
        if np.any((p >= 0.001) & (p < 0.999)):
            good += 1
        elif np.any(p < 0.001):
            occl += 1
        else:
            bad += 1
    return np.array([good, bad, occl, all_pts])


def test_c_oracle_equals_numpy_restatement(oracle_mod):
    from rgbdslam_v2_b200 import synth
    poses = synth.trajectory(240)
    d0, d1 = synth.render_frame(poses[0], seed=0)[1], synth.render_frame(poses[30], seed=30)[1]
    K = (synth.FX, synth.FY, synth.CX, synth.CY)
    z_old, z_new = oracle_mod.create_cloud_z(d0), oracle_mod.create_cloud_z(d1)
    assert z_old.shape == (240, 320) and np.isnan(z_old).any()
    sub = d0[::2, ::2].copy(); sub[~(sub >= 0.1)] = np.nan
    assert np.array_equal(np.isnan(z_old), np.isnan(sub)) and np.array_equal(z_old[~np.isnan(z_old)], sub[~np.isnan(sub)])
    prm = oracle_mod.make_params(depth_cov_z0=2.0)
    cov = (0.01 * 2.0 * 2.0) ** 2
    T = np.linalg.inv(poses[0]) @ poses[30]
    for dz in (0.0, 0.3, -0.5):
        Tb = T.copy(); Tb[2, 3] += dz
        got = oracle_mod.pairwise_observation(prm, Tb, z_new, K, z_old, K)
        Ti = np.linalg.inv(Tb.astype(np.float32).astype(np.float64))
        exp = _numpy_direction(Tb, z_new, K, z_old, K, cov) + _numpy_direction(Ti, z_old, K, z_new, K, cov)
        assert got[3] == exp[3] == 2400
        assert np.abs(got[:3].astype(int) - exp[:3]).max() <= 2, (dz, got, exp)   # float rounding of the inverse / projection
