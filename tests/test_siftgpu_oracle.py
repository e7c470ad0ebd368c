"""The restated SiftMatchGPU (oracle/sift_oracle.py, external/SiftGPU/src/SiftGPU/ProgramCU.cu:1405-1478, 1689-1784) against
hand-checked cases of its tie rules, thresholds and quantisation (CPU only)."""
import numpy as np

from oracle import sift_oracle as so


def test_quantisation_wraps_like_unsigned_char():
    q = so.siftgpu_quantise(np.array([[0.0, 0.001, 0.25, 0.4999, 0.5, 0.00097]], np.float32))[0]
    assert list(q) == [0, 1, 128, 0, 0, 0]  # int(512 * 0.4999 + 0.5) = 256 -> 0; int(0.4966 + 0.5) = 0


def test_row_pass_tie_rule_is_thread_major():
    # equal maxima at columns 2 and 33: thread 1 (col 33) beats thread 2 (col 2) in the tree reduction
    dot = np.zeros((1, 40), np.int64)
    dot[0, 2] = dot[0, 33] = 250000
    assert so.siftgpu_row_match(dot, distmax=2.0, ratiomax=2.0)[0] == 33
    # same thread (cols 1 and 33): the first one met wins
    dot[:] = 0; dot[0, 1] = dot[0, 33] = 250000
    assert so.siftgpu_row_match(dot, distmax=2.0, ratiomax=2.0)[0] == 1


def test_col_pass_takes_the_lowest_row_and_thresholds_apply():
    dot = np.zeros((20, 3), np.int64)
    dot[4, 0] = dot[11, 0] = 260000
    dot[7, 1] = 262144                     # dist = acos(1) = 0 < 0.9; runner-up 0 -> distn = pi/2
    dot[9, 2] = 100000                     # acos(0.381) = 1.18 > 0.9 -> rejected
    c = so.siftgpu_col_match(dot)
    assert c[0] == -1                      # runner-up equals the maximum: ratio test fails (dist < distn * 0.9 is false)
    assert c[1] == 7 and c[2] == -1
    assert so.siftgpu_col_match(dot, distmax=2.0, ratiomax=2.0)[0] == 4


def test_mutual_best_and_index_zero_heuristic():
    rng = np.random.default_rng(0)
    base = np.abs(rng.normal(0, 1, (50, 128))).astype(np.float32)
    base /= np.linalg.norm(base, axis=1, keepdims=True)
    q = base[[7, 3, 20, 41]].copy()
    m = so.siftgpu_match(q, base)
    assert [(int(a), int(b)) for a, b in zip(m["queryIdx"], m["trainIdx"])] == [(0, 7), (1, 3), (2, 20), (3, 41)]
    assert np.all(m["distance"] == 0)
    # a single match that involves index 0 is thrown away ("context error", sift_gpu_wrapper.cpp:204-213)
    assert len(so.siftgpu_match(base[[0]], base)) == 0
    # two query rows identical: the column pass gives the train row to the LOWER query index
    q2 = base[[5, 5, 9]]
    m2 = so.siftgpu_match(q2, base)
    assert [(int(a), int(b)) for a, b in zip(m2["queryIdx"], m2["trainIdx"])] == [(2, 9)]  # row 0/1 tie -> ratio test fails for col 5
