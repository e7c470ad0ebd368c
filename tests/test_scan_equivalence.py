"""The 32-records-at-a-time replay of the best-hypothesis bookkeeping used by ransac_select_kernel (`ransac_scan_warp`,
csrc/frontend_kernels.cu) is equivalent to the sequential loop of the reference (node.cpp:1130, 1170-1190: best by error /
inliers, `n += 10` at > 50 % and > 75 % inliers, break at > 80 %).  Both are restated here in Python and compared on random
hypothesis records, ties and early breaks included.  CPU only: this pins the ALGORITHM; the CUDA code is covered by the
bit-exact GPU tests."""
import numpy as np


def scan_sequential(cnt, err, M, min_thr, n_limit):
    rmse, best_cnt, best_n, valid, done = np.float32(1e6), 0, -1, 0, False
    n = 0
    while n < n_limit:
        c = int(cnt[n])
        if c > 0:
            valid += 1
            e = float(err[n])
            if e <= float(rmse) and c >= best_cnt and c >= min_thr:
                rmse, best_cnt, best_n = np.float32(e), c, n
                if c > M * 0.5:
                    n += 10
                if c > M * 0.75:
                    n += 10
                if c > M * 0.8:
                    done = True
                    break
        n += 1
    return float(rmse), best_cnt, best_n, valid, done


def scan_chunks(cnt, err, M, min_thr, n_limit, width=32):
    rmse, best_cnt, best_n, valid, done = np.float32(1e6), 0, -1, 0, False
    n = 0
    while n < n_limit:
        idx = [n + lane for lane in range(width)]
        val = [i < n_limit and cnt[i] > 0 for i in idx]
        imp = [val[k] and float(err[idx[k]]) <= float(rmse) and cnt[idx[k]] >= best_cnt and cnt[idx[k]] >= min_thr
               for k in range(width)]
        if not any(imp):
            valid += sum(val)
            n += width
            continue
        f = imp.index(True)                       # __ffs(ballot) - 1
        valid += sum(val[: f + 1])
        i = idx[f]
        c = int(cnt[i])
        rmse, best_cnt, best_n = np.float32(err[i]), c, i
        nn = i + (10 if c > M * 0.5 else 0) + (10 if c > M * 0.75 else 0)
        if c > M * 0.8:
            done = True
            break
        n = nn + 1
    return float(rmse), best_cnt, best_n, valid, done


def test_chunked_scan_equals_the_sequential_bookkeeping():
    rng = np.random.default_rng(0)
    for trial in range(4000):
        H = int(rng.integers(1, 201))
        M = int(rng.integers(21, 301))
        min_thr = min(20, int(0.75 * M))
        cnt = np.where(rng.random(H) < 0.7, rng.integers(0, M + 1, H), 0)
        err = rng.random(H) * 3
        if trial % 3 == 0:
            err = np.round(err, 1)                # ties on the error
        if trial % 5 == 0:
            cnt = np.minimum(cnt, int(0.6 * M))   # never reaches the 80 % break
        assert scan_sequential(cnt, err, M, min_thr, H) == scan_chunks(cnt, err, M, min_thr, H), trial
