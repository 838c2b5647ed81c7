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
    rng = np.random.default_rng(2026)
    m = vsm_b200.Matcher(engine=vsm_b200.ENGINE_TENSOR, scratch_rows=256)
    kf, prev_handle, prev_frame = [], -1, None
    counts = {}
    for it in range(140):
        if rng.random() < 0.15:
            m.set_profiling(bool(rng.integers(0, 2)))
        op = rng.choice(["knn", "match", "match", "track", "add", "global", "segmented", "masked", "batch", "repeat"])
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
            for _ in range(3 if op == "repeat" else 1):             # repeats hit the plan cache
                good, raw = m.match_features(a, b, ratio, mutual=mutual)
                assert good.tobytes() == og.tobytes() and raw.tobytes() == orw.tobytes(), (it, op)
        elif op == "track":
            d = gen.rows(5000 + it, 0, 0, nq) if prev_frame is None else q
            mutual = bool(rng.integers(0, 2))
            good, raw, h = m.track(prev_handle, it, _maybe_pinned(d, rng), 0.75, mutual=mutual, want_raw=True)
            if prev_frame is not None:
                og, orw = oracle.match_features(prev_frame, d, 0.75, mutual=mutual)
                assert good.tobytes() == og.tobytes() and raw.tobytes() == orw.tobytes(), (it, op)
            kf.append(d)
            prev_handle, prev_frame = h, d
        elif op == "add":
            h = m.add_keyframe(it, t)
            kf.append(t)
            prev_handle, prev_frame = h, t
        elif kf and op == "global":
            db = np.concatenate(kf)
            gi, gd = m.search_map_points(_maybe_pinned(q, rng))
            oi, od = oracle.knn(q, db, 2)
            assert np.array_equal(gi, oi) and np.array_equal(bits(gd), bits(od)), (it, op)
        elif kf and op == "segmented":
            db = np.concatenate(kf)
            seg = np.concatenate([[0], np.cumsum([len(k) for k in kf])]).astype(np.int64)
            c, lists = m.detect_candidates(_maybe_pinned(q, rng), 0.75)
            oc, ol = oracle.segmented(q, db, seg, 0.75)
            assert np.array_equal(c, oc), (it, op)
            for s in range(len(kf)):
                assert lists[s].tobytes() == ol[s].tobytes(), (it, op, s)
        elif kf and op == "masked":
            db = np.concatenate(kf)
            mask = (rng.random(db.shape[0]) < 0.4).astype(np.uint8)
            ids = np.nonzero(mask)[0]
            gi, gd = m.search_map_points_masked(q, mask)
            oi, od = oracle.knn(q, db[ids], 2)
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
    assert len(counts) >= 8
