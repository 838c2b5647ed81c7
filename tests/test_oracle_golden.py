"""The CPU oracle against the committed cv2 (OpenCV 4.13.0) golden answers.

The reference has no tests for this path, so these fixtures --
OpenCV's own BFMatcher output followed by the reference's filter loops
(src/Slam.cpp:1151-1158, src/LoopCloser.cpp:54-62) -- are what pins the oracle.
Bar: indices identical, fp32 distances bit-identical.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import cases, gen, oracle


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("name", list(cases.PAIR_CASES))
def test_knn_matches_cv2_golden(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    q, t = cases.PAIR_CASES[name]()
    idx, dist = oracle.knn(q, t, 2)
    assert np.array_equal(idx, g["idx"])
    assert np.array_equal(bits(dist), bits(g["dist"]))
    back, _ = oracle.knn(t, q, 1)
    assert np.array_equal(back[:, 0], g["back"])


@pytest.mark.parametrize("name", list(cases.PAIR_CASES))
@pytest.mark.parametrize("ratio", cases.RATIOS)
def test_match_features_matches_golden(name, ratio):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    q, t = cases.PAIR_CASES[name]()
    for mutual, key in ((False, "good"), (True, "mutual")):
        good, raw = oracle.match_features(q, t, ratio, mutual=mutual)
        want = g[f"{key}_{int(ratio * 100)}"]
        assert np.array_equal(good["queryIdx"], want)
        assert np.array_equal(good["trainIdx"], g["idx"][want, 0])
        assert np.array_equal(bits(good["distance"]), bits(g["dist"][want, 0]))
        assert np.all(good["imgIdx"] == 0)
        # raw = every m[0] with m.size() >= 2 (src/Slam.cpp:1152-1153)
        has2 = np.nonzero(g["idx"][:, 1] >= 0)[0]
        assert np.array_equal(raw["queryIdx"], has2)
        assert np.array_equal(raw["trainIdx"], g["idx"][has2, 0])


def test_segmented_and_global_db_match_golden():
    g = np.load(os.path.join(GOLDEN, "db_small.npz"))
    q, db, seg_off = cases.db_case()
    assert np.array_equal(seg_off, g["seg_off"])
    gi, gd = oracle.knn(q, db, 2)
    assert np.array_equal(gi, g["gidx"]) and np.array_equal(bits(gd), bits(g["gdist"]))
    for col, ratio in enumerate(cases.RATIOS):
        counts, lists = oracle.segmented(q, db, seg_off, ratio)
        assert np.array_equal(counts, g["counts"][:, col])
        for s, m in enumerate(lists):
            assert np.array_equal(m["trainIdx"], g["seg_idx"][s][m["queryIdx"], 0])
            assert np.all(m["imgIdx"] == s)


def test_shard_merge_equals_whole():
    """Top-2 over a row-partitioned DB merged by (distance, global index) equals one pass."""
    q, db, _ = cases.db_case()
    whole_i, whole_d = oracle.knn(q, db, 2)
    for nshard in (2, 3, 8):
        cuts = np.linspace(0, db.shape[0], nshard + 1).astype(np.int64)
        ii, dd = [], []
        for s in range(nshard):
            i, d = oracle.knn(q, db[cuts[s]:cuts[s + 1]], 2)
            ii.append(np.where(i >= 0, i + cuts[s], -1))
            dd.append(d)
        mi, md = oracle.merge_top2(np.stack(ii), np.stack(dd))
        assert np.array_equal(mi, whole_i) and np.array_equal(bits(md), bits(whole_d))


def test_merge_tie_break_lowest_global_index():
    q, t = cases.PAIR_CASES["dups"]()
    whole_i, _ = oracle.knn(q, t, 2)
    # split so that the duplicate rows 7 / 40 / 45 land in different shards
    cuts = [0, 20, 42, 64]
    ii, dd = [], []
    for s in range(3):
        i, d = oracle.knn(q, t[cuts[s]:cuts[s + 1]], 2)
        ii.append(np.where(i >= 0, i + cuts[s], -1))
        dd.append(d)
    # feed shards in reverse order: the merge must not depend on arrival order
    mi, _ = oracle.merge_top2(np.stack(ii[::-1]), np.stack(dd[::-1]))
    assert np.array_equal(mi, whole_i)
    assert list(mi[0]) == [7, 40]


def test_generator_twin_and_unit_norm():
    a = gen.rows(3, 2, 100, 64)
    b = oracle.gen_rows(3, 2, 100, 64)
    assert np.array_equal(bits(a), bits(b))
    assert np.allclose(np.linalg.norm(a.astype(np.float64), axis=1), 1.0, atol=2e-7)
    # counter-based: any row range reproduces the same bytes
    assert np.array_equal(bits(gen.rows(3, 2, 0, 164)[100:]), bits(a))


def test_l2sqr_order_is_opencv_baseline():
    """4 accumulators x 4 lanes, mul then add (no FMA), (x0+x2)+(x1+x3) -- numpy restatement."""
    a = gen.rows(1, 0, 0, 8)
    b = gen.rows(1, 1, 0, 8)
    for i in range(8):
        d = (a[i] - b[i]).astype(np.float32)
        acc = np.zeros((4, 4), np.float32)
        for j in range(0, 256, 16):
            for k in range(4):
                seg = d[j + 4 * k: j + 4 * k + 4]
                acc[k] = (seg * seg).astype(np.float32) + acc[k]
        s = ((acc[0] + acc[1]) + acc[2]) + acc[3]
        want = np.float32(np.float32(s[0] + s[2]) + np.float32(s[1] + s[3]))
        assert bits(oracle.l2sqr(a[i], b[i])) == bits(want)


def test_empty_inputs():
    z = np.zeros((0, 256), np.float32)
    t = gen.rows(0, 0, 0, 10)
    good, raw = oracle.match_features(z, t)
    assert len(good) == 0 and len(raw) == 0          # src/Slam.cpp:1143
    good, raw = oracle.match_features(t, z)
    assert len(good) == 0 and len(raw) == 0
