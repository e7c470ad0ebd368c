"""oracle/sift_oracle.py (float-descriptor branch, node.cpp:610-667 + 1557-1581) pinned against OpenCV -- the library the
reference calls: the exact 2-NN equals cv2.BFMatcher(NORM_L2).knnMatch, RootSIFT equals the cv::Mat arithmetic of
squareroot_descriptor_space, and the ratio / uniqueness rules are checked on a hand-made case.  CPU only."""
import numpy as np


def _sift_like(rng, n):
    d = rng.gamma(0.6, 30.0, size=(n, 128)).astype(np.float32)
    return np.minimum(d, 255.0).astype(np.float32)


def test_root_sift_matches_the_cv_mat_arithmetic():
    import cv2
    from oracle import sift_oracle
    rng = np.random.default_rng(0)
    d = _sift_like(rng, 200)
    d[7] = 0                                                  # zero rows are left alone (node.cpp:1565)
    exp = np.abs(d).copy()
    for i in range(len(exp)):                                 # node.cpp:1560-1569 with cv::Mat ops
        s = float(cv2.sumElems(exp[i:i + 1])[0])
        if s != 0:
            exp[i] = cv2.sqrt(exp[i:i + 1] / np.float32(s))[0]
    got = sift_oracle.root_sift(d)
    assert np.abs(got - exp).max() < 1e-6 and np.all(got[7] == 0)
    assert np.allclose((got[:7] ** 2).sum(1), 1.0, atol=1e-5)   # RootSIFT rows have unit L2 norm


def test_exact_2nn_equals_cv2_bfmatcher():
    import cv2
    from oracle import sift_oracle
    rng = np.random.default_rng(1)
    q, t = sift_oracle.root_sift(_sift_like(rng, 300)), sift_oracle.root_sift(_sift_like(rng, 500))
    q[:100] = t[rng.permutation(500)[:100]] + rng.normal(0, 0.002, (100, 128)).astype(np.float32)
    idx, d = sift_oracle.knn2_exact(q, t)
    knn = cv2.BFMatcher(cv2.NORM_L2).knnMatch(q, t, k=2)
    cv_idx = np.array([[m[0].trainIdx, m[1].trainIdx] for m in knn])
    cv_d = np.array([[m[0].distance, m[1].distance] for m in knn])
    assert (idx[:, 0] == cv_idx[:, 0]).mean() > 0.995            # float32 vs float64 distance ties only
    same = idx[:, 0] == cv_idx[:, 0]
    assert np.abs(np.sqrt(d[same, 0]) - cv_d[same, 0]).max() < 1e-4   # the oracle returns SQUARED distances like cv::flann


def test_ratio_and_uniqueness_rules():
    from oracle import sift_oracle
    t = np.zeros((4, 128), np.float32); q = np.zeros((3, 128), np.float32)
    t[0, 0] = 1.0; t[1, 1] = 1.0; t[2, 2] = 1.0; t[3, 0] = 0.98; t[3, 3] = 0.2
    q[0, 0] = 1.0                     # nearest t0 (d=0), runner-up t3: ratio 0 -> accepted
    q[1, 0] = 0.995; q[1, 3] = 0.1    # nearest t3 / t0 almost tied -> ratio close to 1 -> rejected at 0.95
    q[2, 0] = 0.999                   # nearest t0 again -> trainIdx already taken by query 0 (first come)
    m = sift_oracle.feature_matching(q, t, nn_ratio=0.95, max_matches=300)
    assert [(int(a), int(b)) for a, b in zip(m["queryIdx"], m["trainIdx"])] == [(0, 0)]
    assert m["distance"][0] == 0.0    # distance = ratio of the squared distances
