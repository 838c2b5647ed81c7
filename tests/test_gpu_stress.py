"""Randomised soak of the C ABI on one context: entry points, sizes, pinned / pageable inputs and
profiling on / off are drawn at random, every answer is compared with the CPU oracle bit for bit.
Shakes the per-call machinery that carries state from one call to the next: the plan cache, the
alternating statistics slots, buffer growth, the zero-copy prologue, programmatic dependent launch."""
import numpy as np
import pytest

from oracle import gen, oracle
import vsm_b200

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _maybe_pinned(a, rng):
    if rng.random() < 0.5:
        return a
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()


def _scaled(t, rng):
    """Some calls get rows that are NOT unit length (the norm-spread term of the margin)."""
    if rng.random() < 0.8:
        return t
    s = rng.uniform(0.5, 1.5, size=(t.shape[0], 1)).astype(np.float32)
    return np.ascontiguousarray(t * s)


def test_random_call_sequence():
    import os
    # VSM_SOAK_SEED / VSM_SOAK_ITERS: a longer soak with another seed (run by hand on the GPU box)
    rng = np.random.default_rng(int(os.environ.get("VSM_SOAK_SEED", "2026")))
    iters = int(os.environ.get("VSM_SOAK_ITERS", "160"))
    m = vsm_b200.Matcher(engine=vsm_b200.ENGINE_TENSOR, scratch_rows=256)
    # model of the store: keyframes in Map::get_keyframes() order as (handle, matrix); the reference frame
    kf, prev_handle, prev_frame, prev_is_kf = [], -1, None, False
    counts = {}

    def layout():
        """concatenated keyframe matrix, its offsets, and each keyframe's first store row"""
        order = m.keyframes()
        assert [int(h) for h in order] == [h for h, _ in kf]
        db = np.concatenate([d for _, d in kf])
        seg = np.concatenate([[0], np.cumsum([len(d) for _, d in kf])]).astype(np.int64)
        row0 = np.array([m.frame_info(h)[3] for h, _ in kf], np.int64)
        return db, seg, row0

    def to_store_rows(oi, seg, row0):
        s_ = np.clip(np.searchsorted(seg, np.maximum(oi, 0), side="right") - 1, 0, len(row0) - 1)
        return np.where(oi >= 0, row0[s_] + (oi - seg[s_]), -1)

    for it in range(iters):
        if rng.random() < 0.15:
            m.set_profiling(bool(rng.integers(0, 2)))
        op = rng.choice(["knn", "match", "match", "track", "track", "add", "global", "segmented", "masked", "batch",
                         "repeat", "remove", "compact", "compact"])
        counts[op] = counts.get(op, 0) + 1
        nq, nt = int(rng.integers(1, 700)), int(rng.integers(1, 1100))
        q, t, _ = gen.planted(1000 + it, nq, nt, 0.5, 0.09)
        t = _scaled(t, rng)
        if op == "knn":
            gi, gd = m.knn_match(_maybe_pinned(q, rng), _maybe_pinned(t, rng))
            oi, od = oracle.knn(q, t, 2)
            assert np.array_equal(gi, oi) and np.array_equal(bits(gd), bits(od)), (it, op)
        elif op in ("match", "repeat"):
            mutual = bool(rng.integers(0, 2))
            ratio = float(rng.choice([0.7, 0.75, 0.8]))
            a, b = _maybe_pinned(q, rng), _maybe_pinned(t, rng)
            og, orw = oracle.match_features(q, t, ratio, mutual=mutual)
            want_raw = bool(rng.integers(0, 2))                     # without a raw list: the tile top-2 records
            for _ in range(3 if op == "repeat" else 1):             # repeats hit the plan cache
                good, raw = m.match_features(a, b, ratio, mutual=mutual, want_raw=want_raw)
                assert good.tobytes() == og.tobytes(), (it, op, want_raw)
                assert not want_raw or raw.tobytes() == orw.tobytes(), (it, op)
        elif op == "track":
            d = gen.rows(5000 + it, 0, 0, nq) if prev_frame is None else q
            mutual = bool(rng.integers(0, 2))
            want_raw = bool(rng.integers(0, 2))
            good, raw, h = m.track(prev_handle, it, _maybe_pinned(d, rng), 0.75, mutual=mutual, want_raw=want_raw)
            if prev_frame is not None:
                og, orw = oracle.match_features(prev_frame, d, 0.75, mutual=mutual)
                assert good.tobytes() == og.tobytes(), (it, op, want_raw)
                assert not want_raw or raw.tobytes() == orw.tobytes(), (it, op)
            # the frame is a plain frame until the "keyframe decision" (src/Slam.cpp:1065/:1076)
            prev_is_kf = bool(rng.random() < 0.4)
            if prev_is_kf:
                m.promote(h)
                kf.append((h, d))
            prev_handle, prev_frame = h, d
        elif op == "add":
            h = m.add_keyframe(it, t)
            kf.append((h, t))
            prev_handle, prev_frame, prev_is_kf = h, t, True
        elif op == "remove" and len(kf) > 2:
            k = int(rng.integers(0, len(kf)))
            h, _ = kf.pop(k)
            m.remove_frame(h)
            if h == prev_handle:
                prev_handle, prev_frame = -1, None
        elif kf and op == "global":
            db, seg, row0 = layout()
            gi, gd = m.search_map_points(_maybe_pinned(q, rng))
            oi, od = oracle.knn(q, db, 2)
            assert np.array_equal(gi, to_store_rows(oi, seg, row0)) and np.array_equal(bits(gd), bits(od)), (it, op)
        elif kf and op == "segmented":
            db, seg, _ = layout()
            c, lists = m.detect_candidates(_maybe_pinned(q, rng), 0.75)
            oc, ol = oracle.segmented(q, db, seg, 0.75)
            assert np.array_equal(c, oc), (it, op)
            for s in range(len(kf)):
                assert lists[s].tobytes() == ol[s].tobytes(), (it, op, s)
        elif kf and op == "compact":
            db, seg, _ = layout()
            ids = np.array([m.frame_info(h)[1] for h, _ in kf], np.int32)
            every, mm, gap = int(rng.integers(1, 4)), int(rng.choice([0, 1, 5, 30])), int(rng.choice([0, 40]))
            cur = it + 20
            st_c, lists_c, _ = m.loop_detect_compact(cur, _maybe_pinned(q, rng), 0.75, min_gap=gap, every=every, min_matches=mm)
            ost, ol = oracle.loop_detect(q, db, seg, ids, cur, 0.75, min_gap=gap, every=every)
            assert np.array_equal(st_c, ost), (it, op)
            assert set(lists_c) == {s for s in range(len(ost)) if ost[s] >= max(mm, 1)}, (it, op)
            for s in lists_c:
                assert lists_c[s].tobytes() == ol[s].tobytes(), (it, op, s)
        elif kf and op == "masked":
            # the mask runs over store rows; only rows of live keyframes are offered here
            db, seg, row0 = layout()
            nrows = m.store_info()[0]
            pick = rng.random(db.shape[0]) < 0.4
            rows_of_db = to_store_rows(np.arange(db.shape[0]), seg, row0)
            mask = np.zeros(nrows, np.uint8)
            mask[rows_of_db[pick]] = 1
            ids = np.nonzero(mask)[0]                                 # ascending store row = the compacted order
            phys = np.zeros((nrows, 256), np.float32)
            phys[rows_of_db] = db
            gi, gd = m.search_map_points_masked(q, mask)
            oi, od = oracle.knn(q, phys[ids], 2)
            want = np.where(oi >= 0, ids[np.maximum(oi, 0)] if len(ids) else -1, -1)
            assert np.array_equal(gi, want) and np.array_equal(bits(gd), bits(od)), (it, op)
        elif op == "batch":
            npair = int(rng.integers(1, 6))
            qs, ts = [], []
            for p in range(npair):
                a, b, _ = gen.planted(9000 + 10 * it + p, int(rng.integers(1, 300)), int(rng.integers(1, 400)), 0.5, 0.09)
                qs.append(a)
                ts.append(b)
            mutual = bool(rng.integers(0, 2))
            res = m.match_batch(qs, ts, 0.75, mutual=mutual)
            for p in range(npair):
                og, _ = oracle.match_features(qs[p], ts[p], 0.75, mutual=mutual)
                assert res[p].tobytes() == og.tobytes(), (it, op, p)
        if len(kf) > 14:                                            # keep the database (and the oracle's work) small
            m.clear_store()
            kf, prev_handle, prev_frame = [], -1, None
    m.close()
    assert len(counts) >= 9
