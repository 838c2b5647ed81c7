"""Pin the oracle against OpenCV itself (TEST INFRASTRUCTURE ONLY).

cv2 4.13.0's BFMatcher is the same C++ code the reference links
(cv::DescriptorMatcher::knnMatch, src/Slam.cpp:1149, src/LoopCloser.cpp:51).
This script checks, on seeded inputs, that oracle/vsm_oracle.c returns the same
indices AND the same fp32 distance bits, including ties (lowest trainIdx first),
nt in {0,1,2}, and the 1x256 single-row cases.

Run here (cv2 importable); exits non-zero on any difference.
"""
import os
import sys
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import gen, oracle  # noqa: E402


def cv2_knn(q, t, k=2):
    import cv2
    nq = q.shape[0]
    idx = -np.ones((nq, k), np.int64)
    dist = np.full((nq, k), np.finfo(np.float32).max, np.float32)
    if nq == 0 or t.shape[0] == 0:
        return idx, dist
    res = cv2.BFMatcher(cv2.NORM_L2).knnMatch(q, t, k=k)
    for i, ms in enumerate(res):
        for p, m in enumerate(ms):
            assert m.queryIdx == i and m.imgIdx == 0
            idx[i, p] = m.trainIdx
            dist[i, p] = m.distance
    return idx, dist


def same_bits(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32))


def main():
    bad = 0
    cases = []
    for seed in range(3):
        cases.append(("random", gen.rows(seed, 0, 0, 257), gen.rows(seed, 1, 0, 513)))
        q, t, _ = gen.planted(seed, 400, 400, 0.6, 0.08)
        cases.append(("planted400", q, t))
    q, t, _ = gen.planted(0, 2000, 2000, 0.6, 0.08)
    cases.append(("planted2000", q, t))
    # duplicates: exact ties must resolve to the lowest train index
    t = gen.rows(5, 1, 0, 64).copy()
    t[40] = t[7]; t[45] = t[7]; t[3] = t[60]
    cases.append(("dups", np.concatenate([t[7:8], t[60:61], gen.rows(5, 0, 0, 30)]), t))
    for nt in (1, 2, 3):
        cases.append((f"nt{nt}", gen.rows(9, 0, 0, 17), gen.rows(9, 1, 0, nt)))
    cases.append(("nq1", gen.rows(9, 0, 0, 1), gen.rows(9, 1, 0, 100)))
    # non-unit-norm rows and large magnitudes
    cases.append(("scaled", gen.rows(11, 0, 0, 50) * np.float32(3.7), gen.rows(11, 1, 0, 300) * np.float32(0.01)))
    for name, q, t in cases:
        for k in (1, 2):
            ci, cd = cv2_knn(q, t, k)
            oi, od = oracle.knn(q, t, k)
            ok = np.array_equal(ci, oi) and same_bits(cd, od)
            print(f"{name:12s} k={k} nq={q.shape[0]:5d} nt={t.shape[0]:5d} "
                  f"idx_equal={np.array_equal(ci, oi)} dist_bits_equal={same_bits(cd, od)}")
            bad += (not ok)
    print("OK" if not bad else f"FAILED: {bad} cases differ")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
