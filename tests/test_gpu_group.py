"""vsm_group: several GPUs behind ONE caller thread (the reference's slam_thread, src/main.cpp:1520).
On a 1-GPU box the group is built from several contexts on the same device (the same code path:
worker threads, gather buffer, stacked-row keys, event join, merge kernel); with more GPUs the
devices differ and the gather stores travel over NVLink.  Every answer is compared with the CPU
oracle over the whole database."""
import numpy as np
import pytest

from oracle import cases, gen, oracle
import vsm_b200

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def device_lists():
    import torch
    n = torch.cuda.device_count()
    lists = [[0], [0, 0, 0]]
    if n >= 2:
        lists.append(list(range(min(n, 8))))
    return lists


@pytest.mark.parametrize("devices", device_lists(), ids=lambda d: "dev" + "".join(map(str, d)))
def test_group_search_and_loop_detect_equal_oracle(devices):
    q, db, seg_off = cases.db_case(seed=3, nq=300, nkf=24)
    nkf = len(seg_off) - 1
    frame_ids = np.arange(nkf, dtype=np.int32) * 37
    with vsm_b200.Group(devices) as g:
        hs = [g.add_keyframe(int(frame_ids[s]), db[seg_off[s]:seg_off[s + 1]]) for s in range(nkf)]
        assert hs == list(range(nkf))
        rows, nk, per = g.store_info()
        assert rows == db.shape[0] and nk == nkf and per.sum() == rows
        if len(devices) > 1:
            assert per.min() > 0 and per.max() < 0.6 * rows          # dealt to every member, roughly evenly
        for rep in range(2):                                          # second call: plan cache, clean tables
            gi, gd, kh, kr = g.search_map_points(q, want_keyframes=True)
            oi, od = oracle.knn(q, db, 2)
            assert np.array_equal(gi, oi) and np.array_equal(bits(gd), bits(od))
            seg_of = np.searchsorted(seg_off, oi, side="right") - 1
            assert np.array_equal(kh, seg_of) and np.array_equal(kr, oi - seg_off[seg_of])
        cur_id = int(frame_ids[-1]) + 150
        st, lists = g.loop_detect(cur_id, q, 0.75, min_gap=200, every=3)
        ost, ol = oracle.loop_detect(q, db, seg_off, frame_ids, cur_id, 0.75, min_gap=200, every=3)
        assert np.array_equal(st, ost) and (ost >= 0).sum() >= 3
        for s in range(nkf):
            if ost[s] >= 0:
                assert lists[s].tobytes() == ol[s].tobytes(), s
        # the compact form: gate and packing on the members' devices, candidates ordered by list position
        for mm in (30, 1):
            cst, clists = g.loop_detect_compact(cur_id, q, 0.75, min_gap=200, every=3, min_matches=mm)
            assert np.array_equal(cst, ost)
            assert set(clists) == {s for s in range(nkf) if ost[s] >= mm}
            for s in clists:
                assert clists[s].tobytes() == ol[s].tobytes(), s
        # pageable and pinned query buffers give the same answer
        import torch
        pq = torch.from_numpy(q).pin_memory().numpy()
        gi2, gd2 = g.search_map_points(pq)
        assert np.array_equal(gi2, gi) and np.array_equal(bits(gd2), bits(gd))
        # remove a keyframe: its rows leave the search, the stacked numbering of the others stays
        g.remove_keyframe(7)
        keep = np.ones(db.shape[0], bool)
        keep[seg_off[7]:seg_off[8]] = False
        ids = np.nonzero(keep)[0]
        gi3, gd3 = g.search_map_points(q)
        oi3, od3 = oracle.knn(q, db[ids], 2)
        assert np.array_equal(gi3, ids[oi3]) and np.array_equal(bits(gd3), bits(od3))
        st3, _ = g.loop_detect(cur_id, q, 0.75, min_gap=200, every=3, want_matches=False)
        seg2 = np.concatenate([[0], np.cumsum(np.delete(np.diff(seg_off), 7))]).astype(np.int64)
        ost3, _ = oracle.loop_detect(q, db[ids], seg2, np.delete(frame_ids, 7), cur_id, 0.75, min_gap=200, every=3)
        assert np.array_equal(st3, ost3)
        # pair matching through a member context still works next to the group calls
        m0 = g.member(0)
        a, b, _ = gen.planted(5, 300, 280, 0.6, 0.08)
        good, _ = m0.match_features(a, b, 0.75, mutual=True)
        og, _ = oracle.match_features(a, b, 0.75, mutual=True)
        assert good.tobytes() == og.tobytes()
        g.clear_store()
        assert g.store_info()[:2] == (0, 0)


@pytest.mark.parametrize("devices", device_lists()[1:], ids=lambda d: "dev" + "".join(map(str, d)))
def test_group_with_adopted_shards_and_duplicates_across_members(devices):
    """Each member adopts its own device matrix (how a bulk database is loaded); exact duplicates of one
    row sit on different members: the merge must order them by stacked row like one pass over the whole."""
    import torch
    n = len(devices)
    db = gen.rows(41, 0, 0, 9000).copy()
    q = gen._normalize_int(1000 * gen.int_rows(41, 0, 100, 200) + 600 * gen.int_rows(42, 0, 0, 200))
    per = 9000 // n
    # the same row three times: in the first member, in the last one, and once more in the last one
    db[per - 5] = db[50]
    db[9000 - 7] = db[50]
    q[0] = db[50]
    with vsm_b200.Group(devices) as g:
        keep = []
        for r in range(n):
            lo, hi = r * per, (9000 if r == n - 1 else (r + 1) * per)
            with torch.cuda.device(devices[r]):
                t = torch.from_numpy(db[lo:hi]).cuda(devices[r])
            keep.append(t)
            seg = np.arange(0, hi - lo + 1, 500, dtype=np.int64)
            if seg[-1] != hi - lo:
                seg = np.append(seg, hi - lo)
            g.adopt_device_matrix(r, t.data_ptr(), hi - lo, seg)
        gi, gd = g.search_map_points(q)
        oi, od = oracle.knn(q, db, 2)
        assert np.array_equal(gi, oi) and np.array_equal(bits(gd), bits(od))
        assert list(gi[0]) == [50, per - 5] and gd[0, 0] == 0.0 and gd[0, 1] == 0.0


def test_group_with_more_members_than_keyframes_and_empty_keyframes():
    """Four contexts, two keyframes (one of them empty): members without rows answer "no neighbour" and still take
    part in the merge; the empty keyframe is skipped by the loop search (src/LoopCloser.cpp:45)."""
    q = gen.rows(61, 0, 0, 50)
    kf = gen.rows(61, 1, 0, 90)
    with vsm_b200.Group([0, 0, 0, 0]) as g:
        gi, gd = g.search_map_points(q)                          # nothing stored at all
        assert (gi == -1).all()
        g.add_keyframe(5, np.zeros((0, 256), np.float32))
        g.add_keyframe(9, kf)
        gi, gd, kh, kr = g.search_map_points(q, want_keyframes=True)
        oi, od = oracle.knn(q, kf, 2)
        assert np.array_equal(gi, oi) and np.array_equal(bits(gd), bits(od))
        assert (kh == 1).all() and np.array_equal(kr, oi)
        st, lists = g.loop_detect_compact(500, q, 0.75, min_gap=0, every=1, min_matches=0)
        og, _ = oracle.match_features(q, kf, 0.75)
        assert list(st) == [-1, len(og)]


@pytest.mark.parametrize("devices", device_lists()[1:], ids=lambda d: "dev" + "".join(map(str, d)))
def test_group_ragged_batch_equals_single_context(devices):
    """vsm_group_match_batch: the ragged batch cut into contiguous blocks of pairs, one per member (replicas)."""
    rng = np.random.default_rng(4)
    qs, ts = [], []
    for p in range(13):
        a, b, _ = gen.planted(800 + p, int(rng.integers(1, 500)), int(rng.integers(1, 600)), 0.6, 0.08)
        qs.append(a)
        ts.append(b)
    qs[5] = qs[5][:0]                                             # a pair without queries
    q_off = np.zeros(14, np.int32)
    t_off = np.zeros(14, np.int32)
    q_off[1:] = np.cumsum([len(a) for a in qs])
    t_off[1:] = np.cumsum([len(a) for a in ts])
    qa, ta = np.ascontiguousarray(np.concatenate(qs)), np.ascontiguousarray(np.concatenate(ts))
    with vsm_b200.Group(devices) as g:
        for mutual in (False, True):
            res = g.match_batch_packed(qa, q_off, ta, t_off, 0.75, mutual)
            for p in range(13):
                og, _ = oracle.match_features(qs[p], ts[p], 0.75, mutual=mutual)
                assert res[p].tobytes() == og.tobytes(), (p, mutual)
