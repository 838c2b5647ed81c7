"""GPU parity: the CUDA path (through the C ABI of libvsm.so) against the CPU oracle and the
committed cv2 golden answers.  Bar: identical indices, bit-identical fp32 distances, identical
ratio / mutual decisions -- no tolerance (the re-score reproduces OpenCV's fp32 arithmetic)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import cases, gen, oracle
import vsm_b200

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def tc():
    m = vsm_b200.Matcher(engine=vsm_b200.ENGINE_TENSOR)
    yield m
    m.close()


@pytest.fixture(scope="module")
def pair():
    """tensor-core pass on CTA pairs (tcgen05 cta_group::2, clusters of two)"""
    m = vsm_b200.Matcher(engine=vsm_b200.ENGINE_TENSOR_PAIR)
    yield m
    m.close()


@pytest.fixture(scope="module")
def appendm():
    """tensor-core pass with APPEND records for train sets of up to 16 tiles (vsm_opts.reserved[4]; off by default)"""
    m = vsm_b200.Matcher(engine=vsm_b200.ENGINE_TENSOR, append_tiles=16)
    yield m
    m.close()


@pytest.fixture(scope="module")
def top4m():
    """pair matching with the threshold-driven top-4 records (vsm_opts.reserved[5] = 1): the default for pair
    matching is the tile top-2 epilogue, this keeps the other one covered on the same cases"""
    m = vsm_b200.Matcher(engine=vsm_b200.ENGINE_TENSOR, tile_top2=False)
    yield m
    m.close()


@pytest.fixture(scope="module")
def simt():
    m = vsm_b200.Matcher(engine=vsm_b200.ENGINE_SIMT)
    yield m
    m.close()


def bf16_round(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).to(torch.bfloat16).to(torch.float32).numpy()


def test_tensor_core_tile_is_the_bf16_dot(tc):
    """Raw tcgen05 accumulators of the first 128x256 tile = dot products of the bf16-rounded rows
    (validates the TMA swizzle, the UMMA descriptors and the TMEM read-back)."""
    q, t = gen.rows(1, 0, 0, 128), gen.rows(1, 1, 0, 256)
    got = tc.debug_tile_scores(q, t)
    want = bf16_round(q).astype(np.float64) @ bf16_round(t).astype(np.float64).T
    assert np.abs(got - want).max() < 2e-5
    # ragged: fewer rows than the tile on both sides
    q, t = gen.rows(2, 0, 0, 77), gen.rows(2, 1, 0, 201)
    got = tc.debug_tile_scores(q, t)[:77, :201]
    want = bf16_round(q).astype(np.float64) @ bf16_round(t).astype(np.float64).T
    assert np.abs(got - want).max() < 2e-5


@pytest.mark.parametrize("name", list(cases.PAIR_CASES))
@pytest.mark.parametrize("engine", ["tc", "pair", "simt", "appendm"])
def test_knn_equals_oracle_and_golden(name, engine, request):
    m = request.getfixturevalue(engine)
    q, t = cases.PAIR_CASES[name]()
    idx, dist = m.knn_match(q, t)
    oi, od = oracle.knn(q, t, 2)
    assert np.array_equal(idx, oi)
    assert np.array_equal(bits(dist), bits(od))
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    assert np.array_equal(idx, g["idx"]) and np.array_equal(bits(dist), bits(g["dist"]))


@pytest.mark.parametrize("name", list(cases.PAIR_CASES))
@pytest.mark.parametrize("ratio", cases.RATIOS)
def test_match_features_equals_golden(name, ratio, tc):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    q, t = cases.PAIR_CASES[name]()
    for mutual, key in ((False, "good"), (True, "mutual")):
        good, raw = tc.match_features(q, t, ratio, mutual=mutual)
        want = g[f"{key}_{int(ratio * 100)}"]
        assert np.array_equal(good["queryIdx"], want)
        assert np.array_equal(good["trainIdx"], g["idx"][want, 0])
        assert np.array_equal(bits(good["distance"]), bits(g["dist"][want, 0]))
        assert np.all(good["imgIdx"] == 0)
        has2 = np.nonzero(g["idx"][:, 1] >= 0)[0]
        assert np.array_equal(raw["queryIdx"], has2)
        assert np.array_equal(raw["trainIdx"], g["idx"][has2, 0])
        og, orw = oracle.match_features(q, t, ratio, mutual=mutual)
        assert good.tobytes() == og.tobytes() and raw.tobytes() == orw.tobytes()
    # ratio-only form (no raw list, no mutual test): queries whose approximate top-2 already prove
    # that the ratio test fails are answered without an exact re-score -- same survivors
    good, _ = tc.match_features(q, t, ratio, mutual=False, want_raw=False)
    og, _ = oracle.match_features(q, t, ratio, mutual=False)
    assert good.tobytes() == og.tobytes()


def test_ratio_only_early_out_on_borderline_ratios(tc):
    """Planted pairs with noise levels that put d0/d1 on both sides of the ratio, scaled rows
    (non-unit norms) and extreme ratios: the ratio-only early-out never changes a decision."""
    rng = np.random.default_rng(3)
    for it, sigma in enumerate((0.07, 0.09, 0.11, 0.13)):
        q, t, _ = gen.planted(300 + it, 900, 1100, 0.7, sigma)
        if it & 1:
            t = np.ascontiguousarray(t * rng.uniform(0.7, 1.3, size=(t.shape[0], 1)).astype(np.float32))
        for ratio in (0.6, 0.75, 0.9, 0.999, 1.0, 1.2):
            good, _ = tc.match_features(q, t, ratio, mutual=False, want_raw=False)
            og, _ = oracle.match_features(q, t, ratio, mutual=False)
            assert good.tobytes() == og.tobytes(), (sigma, ratio)
            res = tc.match_batch([q, q[:100]], [t, t[:50]], ratio, mutual=False)
            assert res[0].tobytes() == og.tobytes(), (sigma, ratio)
            # with the mutual test on top (the forward problem still ends early, the reverse one never does)
            good, _ = tc.match_features(q, t, ratio, mutual=True, want_raw=False)
            ogm, _ = oracle.match_features(q, t, ratio, mutual=True)
            assert good.tobytes() == ogm.tobytes(), (sigma, ratio, "mutual")


def test_empty_inputs(tc):
    z = np.zeros((0, 256), np.float32)
    t = gen.rows(0, 0, 0, 10)
    for a, b in ((z, t), (t, z), (z, z)):
        good, raw = tc.match_features(a, b)
        assert len(good) == 0 and len(raw) == 0          # src/Slam.cpp:1143
    idx, dist = tc.knn_match(t, z)
    assert np.all(idx == -1) and np.all(dist == np.finfo(np.float32).max)


def test_keyframe_store_global_and_segmented(tc):
    g = np.load(os.path.join(GOLDEN, "db_small.npz"))
    q, db, seg_off = cases.db_case()
    tc.clear_store()
    handles = [tc.add_keyframe(s, db[seg_off[s]:seg_off[s + 1]]) for s in range(len(seg_off) - 1)]
    assert handles == list(range(len(seg_off) - 1))
    assert tc.store_info() == (db.shape[0], len(seg_off) - 1)
    # stacked-matrix search (src/Slam.cpp:546-574)
    idx, dist = tc.search_map_points(q)
    assert np.array_equal(idx, g["gidx"]) and np.array_equal(bits(dist), bits(g["gdist"]))
    idx2, _ = tc.search_map_points(q, row_offset=1000)
    assert np.array_equal(idx2, g["gidx"] + 1000)
    # LoopCloser::detect block (src/LoopCloser.cpp:43-62)
    for col, ratio in enumerate(cases.RATIOS):
        counts, lists = tc.detect_candidates(q, ratio)
        assert np.array_equal(counts, g["counts"][:, col])
        oc, ol = oracle.segmented(q, db, seg_off, ratio)
        for s in range(len(lists)):
            assert lists[s].tobytes() == ol[s].tobytes()
    # Slam::match_features(ref_kf, cur) with the keyframe resident (src/Slam.cpp:841)
    for s in (0, 3, 5, 7, 15):
        kf = db[seg_off[s]:seg_off[s + 1]]
        for mutual in (False, True):
            good, raw = tc.match_to_keyframe(handles[s], q, 0.75, mutual=mutual, want_raw=True)
            og, orw = oracle.match_features(kf, q, 0.75, mutual=mutual)
            assert good.tobytes() == og.tobytes() and raw.tobytes() == orw.tobytes()
    tc.clear_store()
    assert tc.store_info() == (0, 0)


def test_small_segments_force_overflow_path(simt):
    """seg_tiles=1 + a train set full of near-duplicates: more than three candidates per slice,
    so the select kernel must fall back to exact slice scans -- and still be exact."""
    m = vsm_b200.Matcher(engine=vsm_b200.ENGINE_TENSOR, seg_tiles=1)
    q, t = cases.PAIR_CASES["neardup_db"]()
    idx, dist = m.knn_match(q, t)
    oi, od = oracle.knn(q, t, 2)
    assert np.array_equal(idx, oi) and np.array_equal(bits(dist), bits(od))
    assert m.stats()["flagged_slices"] > 0
    m.close()


def test_ragged_batch(tc):
    sizes = [(200, 2048), (777, 1301), (1, 5), (300, 0), (0, 40), (2048, 200), (513, 511), (129, 257)]
    qs, ts = [], []
    for i, (nq, nt) in enumerate(sizes):
        if nq and nt:
            a, b, _ = gen.planted(40 + i, nq, nt, 0.6, 0.08)
        else:
            a, b = gen.rows(40 + i, 0, 0, nq), gen.rows(40 + i, 1, 0, nt)
        qs.append(a)
        ts.append(b)
    for mutual in (False, True):
        got = tc.match_batch(qs, ts, 0.75, mutual=mutual)
        for a, b, g in zip(qs, ts, got):
            og, _ = oracle.match_features(a, b, 0.75, mutual=mutual)
            assert g.tobytes() == og.tobytes()


def test_random_sizes_property(tc):
    """Random ragged sizes, planted pairs: TC engine == oracle, bit for bit."""
    rng = np.random.default_rng(7)
    for it in range(6):
        nq, nt = int(rng.integers(1, 1500)), int(rng.integers(1, 3000))
        q, t, _ = gen.planted(100 + it, nq, nt, 0.5, 0.09)
        idx, dist = tc.knn_match(q, t)
        oi, od = oracle.knn(q, t, 2)
        assert np.array_equal(idx, oi) and np.array_equal(bits(dist), bits(od))


def test_non_unit_norms_still_exact(tc):
    """Rows with very different norms defeat the dot-only ranking; the margin widens and
    the result stays exact."""
    rng = np.random.default_rng(3)
    q = gen.rows(5, 0, 0, 200) * rng.uniform(0.5, 2.0, (200, 1)).astype(np.float32)
    t = gen.rows(5, 1, 0, 900) * rng.uniform(0.5, 2.0, (900, 1)).astype(np.float32)
    idx, dist = tc.knn_match(q, t)
    oi, od = oracle.knn(q, t, 2)
    assert np.array_equal(idx, oi) and np.array_equal(bits(dist), bits(od))


def test_merge_kernel_matches_oracle(tc):
    import torch
    q, db, _ = cases.db_case()
    nshard = 3
    cuts = np.linspace(0, db.shape[0], nshard + 1).astype(np.int64)
    ii, dd = [], []
    for s in range(nshard):
        i, d = oracle.knn(q, db[cuts[s]:cuts[s + 1]], 2)
        ii.append(np.where(i >= 0, i + cuts[s], -1))
        dd.append(d)
    want_i, want_d = oracle.merge_top2(np.stack(ii), np.stack(dd))
    di = torch.from_numpy(np.stack(ii)).cuda()
    ddv = torch.from_numpy(np.stack(dd)).cuda()
    oi = torch.empty((q.shape[0], 2), dtype=torch.int64, device="cuda")
    od = torch.empty((q.shape[0], 2), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    tc.merge_top2_device(di.data_ptr(), ddv.data_ptr(), nshard, q.shape[0], oi.data_ptr(), od.data_ptr(), sync=True)
    assert np.array_equal(oi.cpu().numpy(), want_i) and np.array_equal(bits(od.cpu().numpy()), bits(want_d))


def test_device_resident_db_search(tc):
    import torch
    q, db, seg_off = cases.db_case()
    d_db = torch.from_numpy(db).cuda()
    d_q = torch.from_numpy(q).cuda()
    torch.cuda.synchronize()
    tc.adopt_device_matrix(d_db.data_ptr(), db.shape[0], seg_off)
    oi = torch.empty((q.shape[0], 2), dtype=torch.int64, device="cuda")
    od = torch.empty((q.shape[0], 2), dtype=torch.float32, device="cuda")
    tc.db_top2_device(d_q.data_ptr(), q.shape[0], 5000, oi.data_ptr(), od.data_ptr(), sync=True)
    wi, wd = oracle.knn(q, db, 2)
    assert np.array_equal(oi.cpu().numpy(), wi + 5000) and np.array_equal(bits(od.cpu().numpy()), bits(wd))
    counts, _ = tc.detect_candidates(q, 0.75, want_matches=False)
    oc, _ = oracle.segmented(q, db, seg_off, 0.75)
    assert np.array_equal(counts, oc)
    tc.clear_store()


def test_track_sequence(tc):
    """vsm_track: each frame is uploaded once and matched against the resident previous frame
    (src/Slam.cpp:838-842); results equal match_features on the same two frames."""
    frames = gen.video(3, 5, 400)
    tc.clear_store()
    _, _, h = tc.track(-1, 0, frames[0])
    for f in range(1, 5):
        for_mutual = (f % 2 == 0)
        good, raw, h2 = tc.track(h, f, frames[f], 0.75, mutual=for_mutual, want_raw=True)
        og, orw = oracle.match_features(frames[f - 1], frames[f], 0.75, mutual=for_mutual)
        assert good.tobytes() == og.tobytes() and raw.tobytes() == orw.tobytes()
        assert len(good) > 100
        h = h2
    # tracked frames are PLAIN frames (Frame::is_keyframe_ false): only the newest two stay resident
    assert tc.store_info()[1] == 0 and tc.store_info()[0] <= 3 * 400
    tc.clear_store()


def test_track_promote_remove_interleaved_with_loop_detect(tc):
    """One context used like the reference's Slam + LoopCloser together: every frame is tracked
    (vsm_track: plain frame, src/Slam.cpp:838-842), some are promoted afterwards (set_keyframe(true),
    :1065/:1076 -- one of them LATE, as the bridge keyframe of :851-863), one keyframe is removed, and
    the loop search (src/LoopCloser.cpp:43-62) and the stacked search must see Map::get_keyframes()
    only (src/Map.cpp:40-47): live keyframes in insertion order, never the plain frames."""
    nfr, n = 40, 220
    frames = gen.video(77, nfr, n)
    tc.clear_store()
    handles, kf_frames = {}, []
    prev, ref_f = -1, -1
    for f in range(nfr):
        good, _, h = tc.track(prev, 10 * f, frames[f], 0.75, mutual=False)
        if f > 0:
            og, _ = oracle.match_features(frames[ref_f], frames[f], 0.75)
            assert good.tobytes() == og.tobytes(), f
        handles[f] = h
        if f % 4 == 0:                                   # keyframe decision after matching
            tc.promote(h)
            kf_frames.append(f)
        if f == 10:                                      # bridge: last_frame_ (frame 9) is promoted after frame 10 was tracked
            tc.promote(handles[9])
            kf_frames.append(9)
            kf_frames.sort()
        prev, ref_f = h, f
    # only the two newest plain frames stay resident: the store holds the keyframes + 2 frames, not all 40
    assert tc.store_info() == ((len(kf_frames) + 2) * n, len(kf_frames))
    # drop a keyframe in the middle; its rows are reused by later frames
    tc.remove_frame(handles[12])
    kf_frames.remove(12)
    extra = gen.rows(79, 0, 0, 150)
    tc.add_keyframe(10 * nfr, extra)
    order = tc.keyframes()
    assert [tc.frame_info(int(h))[1] for h in order] == [10 * f for f in kf_frames] + [10 * nfr]
    assert all(tc.frame_info(int(h))[2] for h in order)
    kf_mats = [frames[f] for f in kf_frames] + [extra]
    db = np.concatenate(kf_mats)
    seg_off = np.concatenate([[0], np.cumsum([len(k) for k in kf_mats])]).astype(np.int64)
    ids = np.array([10 * f for f in kf_frames] + [10 * nfr], np.int32)
    # the query frame re-observes 150 rows of frame 0 (a keyframe far enough back): a loop candidate
    vq = gen.int_rows(80, 0, 0, n).copy()
    f0 = np.rint(frames[0][:150].astype(np.float64) * 3300).astype(np.int64)
    vq[:150] = 1000 * f0 + 900 * gen.int_rows(81, 0, 0, 150)
    q = gen._normalize_int(vq)
    cur_id = 10 * nfr + 50
    st, lists = tc.loop_detect(cur_id, q, 0.75, min_gap=200, every=2)
    ost, ol = oracle.loop_detect(q, db, seg_off, ids, cur_id, 0.75, min_gap=200, every=2)
    assert np.array_equal(st, ost)
    for s in range(len(ids)):
        if ost[s] >= 0:
            assert lists[s].tobytes() == ol[s].tobytes(), s
    # stacked search over the keyframes only: store rows map back through frame_info
    gi, gd = tc.search_map_points(q)
    oi, od = oracle.knn(q, db, 2)
    row0 = np.array([tc.frame_info(int(h))[3] for h in order], np.int64)
    seg_of = np.searchsorted(seg_off, oi, side="right") - 1
    want = row0[seg_of] + (oi - seg_off[seg_of])
    assert np.array_equal(gi, want) and np.array_equal(bits(gd), bits(od))
    # per-keyframe search without the eligibility rules sees the same list
    c, _ = tc.detect_candidates(q, 0.75, want_matches=False)
    oc, _ = oracle.segmented(q, db, seg_off, 0.75)
    assert np.array_equal(c, oc) and (oc >= 30).any()
    tc.clear_store()


def test_store_grows_in_place_without_copies(tc):
    """Keyframes added one by one past several growth steps: earlier rows keep their place and content
    (global top-2 equals the oracle over the concatenation) -- the arena maps new chunks, it never
    re-allocates (vsm_store_info's row count is the only thing that changes)."""
    tc.clear_store()
    mats = []
    for k in range(12):
        mats.append(gen.rows(300 + k, 0, 0, 30000))
        tc.add_keyframe(k, mats[-1])
    db = np.concatenate(mats)
    q = gen._normalize_int(1000 * gen.int_rows(300, 0, 100, 64) + 700 * gen.int_rows(299, 0, 0, 64))
    gi, gd = tc.search_map_points(q)
    oi, od = oracle.knn(q, db, 2)
    assert np.array_equal(gi, oi) and np.array_equal(bits(gd), bits(od))
    assert (gi[:, 0] == 100 + np.arange(64)).all()
    tc.clear_store()


def _pinned(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()


def test_pinned_inputs_take_the_zero_copy_path(tc):
    """Pinned host inputs of up to 2048 rows are read over PCIe by the call's prologue kernel (no
    DMA); larger or pageable ones are copied first.  Same answers either way, on every entry point."""
    q, t = cases.PAIR_CASES["pair_777x1301"]()
    pq, pt = _pinned(q), _pinned(t)
    oi, od = oracle.knn(q, t, 2)
    for a, b in ((pq, pt), (pq, t), (q, pt)):                  # pinned/pinned, pinned/pageable, pageable/pinned
        gi, gd = tc.knn_match(a, b)
        assert np.array_equal(gi, oi) and np.array_equal(bits(gd), bits(od))
        good, raw = tc.match_features(a, b, 0.75, mutual=True)
        og, orw = oracle.match_features(q, t, 0.75, mutual=True)
        assert good.tobytes() == og.tobytes() and raw.tobytes() == orw.tobytes()
    # above the zero-copy limit (2048 rows) a pinned buffer goes through the DMA path
    q2, t2 = gen.planted(8, 2100, 2500, 0.6, 0.08)[:2]
    gi, gd = tc.knn_match(_pinned(q2), _pinned(t2))
    oi, od = oracle.knn(q2, t2, 2)
    assert np.array_equal(gi, oi) and np.array_equal(bits(gd), bits(od))
    # tracking: pinned frames; the first call has no reference frame (stand-alone conversion)
    frames = gen.video(11, 4, 500)
    tc.clear_store()
    prev = -1
    for f, d in enumerate(frames):
        good, _, h = tc.track(prev, f, _pinned(d), 0.75, mutual=True)
        if f > 0:
            og, _ = oracle.match_features(frames[f - 1], d, 0.75, mutual=True)
            assert good.tobytes() == og.tobytes()
        prev = h
    # keyframe DB search with pinned queries (both the per-keyframe and the global form)
    qd, db, seg_off = cases.db_case()
    tc.clear_store()
    for s in range(len(seg_off) - 1):
        tc.add_keyframe(s, db[seg_off[s]:seg_off[s + 1]])
    gi, gd = tc.search_map_points(_pinned(qd))
    oi, od = oracle.knn(qd, db, 2)
    assert np.array_equal(gi, oi) and np.array_equal(bits(gd), bits(od))
    counts, lists = tc.detect_candidates(_pinned(qd), 0.75)
    oc, ol = oracle.segmented(qd, db, seg_off, 0.75)
    assert np.array_equal(counts, oc)
    tc.clear_store()


def test_batch_of_stored_keyframe_pairs(tc):
    """vsm_match_batch_stored: the ragged batch with every frame resident in the store (nothing
    uploaded) equals match_features on the same pairs, incl. an empty keyframe and a repeated one."""
    frames = gen.video(21, 7, 350)
    sizes = [350, 200, 0, 333, 350, 1, 129]
    frames = [f[:n] for f, n in zip(frames, sizes)]
    tc.clear_store()
    handles = [tc.add_keyframe(10 * i, f) for i, f in enumerate(frames)]
    pairs = [(0, 1), (1, 0), (3, 4), (2, 3), (3, 2), (5, 6), (6, 5), (4, 4), (0, 6)]
    for mutual in (False, True):
        res = tc.match_batch_stored([handles[a] for a, _ in pairs], [handles[b] for _, b in pairs], 0.75, mutual=mutual)
        assert len(res) == len(pairs)
        for (a, b), got in zip(pairs, res):
            og, _ = oracle.match_features(frames[a], frames[b], 0.75, mutual=mutual)
            assert got.tobytes() == og.tobytes(), (a, b, mutual)
    assert tc.match_batch_stored([], []) == []
    with pytest.raises(vsm_b200.VsmError):
        tc.match_batch_stored([0], [99])
    tc.clear_store()


def test_per_keyframe_search_adapts_its_epilogue():
    """vsm_db_segmented measures how many (query, keyframe) pairs its ratio-only test could not dismiss
    and picks the next call's tensor-core epilogue from it: maxima-only records while matches are rare
    (a pair that stays open is then re-scanned exactly), top-4 records otherwise.  Same answers."""
    q, db, seg_off = cases.db_case()                       # 30 % of the queries re-observe two keyframes
    nkf = len(seg_off) - 1
    q_none = gen.rows(77, 0, 0, 300)                       # a frame that matches nothing
    # (keyframes of up to 8192 rows keep tile top-2 records by default; the two adaptive record kinds are what
    # larger keyframes get, and what vsm_opts.reserved[5] = 1 selects here)
    m = vsm_b200.Matcher(tile_top2=False)
    for s in range(nkf):
        m.add_keyframe(s, db[seg_off[s]:seg_off[s + 1]])

    def run(qq):
        c, lists = m.detect_candidates(qq, 0.75)
        oc, ol = oracle.segmented(qq, db, seg_off, 0.75)
        assert np.array_equal(c, oc)
        for s in range(nkf):
            assert lists[s].tobytes() == ol[s].tobytes()
        return m.stats(), int(c.max())

    st, best = run(q)                                      # first call: pessimistic -> top-4 records
    assert best > 10 and st["candidates"] > 0
    st, best = run(q_none)                                 # nothing stays open ...
    assert best == 0
    st, best = run(q)                                      # ... so this one runs maxima-only: no candidates,
    assert best > 10 and st["candidates"] == 0 and st["flagged_slices"] > 0       # the open pairs are re-scanned
    st, best = run(q)                                      # many pairs stayed open -> back to top-4 records
    assert best > 10 and st["candidates"] > 0
    status, _ = m.loop_detect(900, q, 0.75, min_gap=200, every=1)
    ost, _ = oracle.loop_detect(q, db, seg_off, list(range(nkf)), 900, 0.75, 200, 1)
    assert np.array_equal(status, ost)
    top4_candidates = st["candidates"]
    m.close()
    # the default for keyframes of this size: tile top-2 records, whatever the previous search saw -- same answers,
    # one exact distance per match instead of ~4 per open pair
    m = vsm_b200.Matcher()
    for s in range(nkf):
        m.add_keyframe(s, db[seg_off[s]:seg_off[s + 1]])
    for qq in (q, q_none, q):
        st, best = run(qq)
        assert st["flagged_slices"] < 20
    assert 0 < st["candidates"] < top4_candidates
    m.close()


def test_loop_detect_eligibility_and_matches(tc):
    """vsm_loop_detect = LoopCloser::detect's loop incl. the gap >= 200 / every-5th rules."""
    q, db, seg_off = cases.db_case()
    nkf = len(seg_off) - 1
    frame_ids = [30 * s for s in range(nkf)]                 # keyframe s was frame 30*s
    tc.clear_store()
    for s in range(nkf):
        tc.add_keyframe(frame_ids[s], db[seg_off[s]:seg_off[s + 1]])
    for cur_id, gap, every in ((900, 200, 5), (900, 200, 1), (650, 100, 3), (100, 200, 5)):
        status, lists = tc.loop_detect(cur_id, q, 0.75, min_gap=gap, every=every)
        ost, ol = oracle.loop_detect(q, db, seg_off, frame_ids, cur_id, 0.75, gap, every)
        assert np.array_equal(status, ost)
        for s in range(nkf):
            if ost[s] >= 0:
                assert lists[s].tobytes() == ol[s].tobytes()
            else:
                assert lists[s] is None
    tc.clear_store()


def test_loop_detect_two_shards_on_one_gpu(tc):
    """vsm_loop_detect_shard: the keyframe list split over two contexts (as over two GPUs) gives the
    whole list's status and match lists; the every-5th counter carries across the shard boundary."""
    import torch
    sh = vsm_b200.load_sharded()
    q, db, seg_off = cases.db_case()
    nkf = len(seg_off) - 1
    frame_ids = [30 * s for s in range(nkf)]
    counts = np.diff(seg_off)
    parts = sh.partition_keyframes(seg_off, 2)
    shards = []
    for k0, k1, r0, r1 in parts:
        m = vsm_b200.Matcher()
        d = torch.from_numpy(db[r0:r1]).cuda().contiguous()
        m.adopt_device_matrix(d.data_ptr(), r1 - r0, np.asarray(seg_off[k0:k1 + 1] - r0, np.int64))
        m.set_frame_ids(frame_ids[k0:k1])
        shards.append((m, d, k0))
    for cur_id, gap, every in ((900, 200, 5), (650, 100, 3), (900, 200, 1)):
        ost, ol = oracle.loop_detect(q, db, seg_off, frame_ids, cur_id, 0.75, gap, every)
        got, after_prev = [], 0
        for m, _, k0 in shards:
            before = sh.loop_checked_before(cur_id, frame_ids, counts, gap, k0)
            assert before == after_prev                       # the table-derived count = the chained one
            st, lists, after_prev = m.loop_detect_shard(cur_id, q, before, 0.75, gap, every)
            for s in range(len(st)):
                g = k0 + s
                assert st[s] == ost[g]
                if ost[g] >= 0:
                    want = ol[g].copy()
                    want["imgIdx"] = s                        # imgIdx is the keyframe's index in ITS store
                    assert lists[s].tobytes() == want.tobytes()
            got.append(st)
        assert np.array_equal(np.concatenate(got), ost)
    for m, _, _ in shards:
        m.close()
    with pytest.raises(vsm_b200.VsmError):
        tc.set_frame_ids([1, 2, 3] * 1000)


def test_spcf_feature_cache_bulk_load(tc, tmp_path):
    """The reference's on-disk descriptor format (SPCF, src/FeatureExtractor.cpp:269-381) loads
    straight into the device store; float entries become keyframes, ORB (CV_8U) entries are skipped."""
    from oracle import spcf
    rng = np.random.default_rng(1)
    frames = gen.video(9, 4, 300)
    entries = {}
    for i, d in enumerate(frames):
        kps = np.zeros((len(d), 7), np.float32)
        kps[:, 0:2] = rng.uniform(0, 640, (len(d), 2))
        entries[3 * i] = (kps, d)                                   # FRAME_STEP = 3 (include/Config.h:123)
    entries[100] = (None, rng.integers(0, 255, (50, 32)).astype(np.uint8))    # an ORB-fallback entry
    entries[101] = (None, np.zeros((0, 0), np.float32))                      # no features
    path = str(tmp_path / "features.bin")
    spcf.write(path, entries)
    back = spcf.read(path)
    assert np.array_equal(back[3][1], frames[1])
    tc.clear_store()
    loaded, skipped, h0 = tc.load_feature_cache(path)
    assert (loaded, skipped, h0) == (4, 2, 0)
    assert tc.store_info() == (1200, 4)
    for f in range(1, 4):
        good, raw = tc.match_to_keyframe(h0 + f - 1, frames[f], 0.75, want_raw=True)
        og, orw = oracle.match_features(frames[f - 1], frames[f], 0.75)
        assert good.tobytes() == og.tobytes() and raw.tobytes() == orw.tobytes()
    # loop_detect uses the frame ids from the file
    status, _ = tc.loop_detect(500, frames[3], min_gap=200, every=1, want_matches=False)
    assert np.all(status >= 0)
    tc.clear_store()
    with pytest.raises(vsm_b200.VsmError):
        tc.load_feature_cache(str(tmp_path / "missing.bin"))


def test_masked_map_point_search(tc):
    """vsm_db_top2_masked = knnMatch(frame, stack(valid map-point descriptors), 2) with indices
    mapped back through mp_ids_vec (src/Slam.cpp:744-774)."""
    q, db, _ = cases.db_case()
    rng = np.random.default_rng(11)
    tc.clear_store()
    tc.add_keyframe(0, db)                                  # the map-point descriptors, one row per point
    for frac in (0.7, 0.1, 0.0003, 0.0):
        mask = (rng.random(db.shape[0]) < frac).astype(np.uint8)
        if frac == 0.0003:
            mask[:] = 0
            mask[[5, 9000 % db.shape[0]]] = 1
        ids = np.nonzero(mask)[0]                           # mp_ids_vec
        idx, dist = tc.search_map_points_masked(q, mask)
        oi, od = oracle.knn(q, db[ids], 2)
        want = np.where(oi >= 0, ids[np.maximum(oi, 0)] if len(ids) else -1, -1)
        assert np.array_equal(idx, want)
        assert np.array_equal(bits(dist), bits(od))
    with pytest.raises(vsm_b200.VsmError):
        tc.search_map_points_masked(q, np.ones(3, np.uint8))
    tc.clear_store()


def test_work_list_overflow_scans_inline():
    """A rescan work list that is too small: select falls back to scanning the slice itself."""
    m = vsm_b200.Matcher(engine=vsm_b200.ENGINE_TENSOR, seg_tiles=1, work_cap=3)
    q, t = cases.PAIR_CASES["neardup_db"]()
    idx, dist = m.knn_match(q, t)
    oi, od = oracle.knn(q, t, 2)
    assert np.array_equal(idx, oi) and np.array_equal(bits(dist), bits(od))
    assert m.stats()["flagged_slices"] > 3
    m.close()


def test_multi_unit_train_set(tc):
    """A train set longer than one work unit (64 tiles) and one slice: several units and slices
    per query tile, shared hints across them."""
    q, t, _ = gen.planted(77, 300, 40000, 0.6, 0.08)
    idx, dist = tc.knn_match(q, t)
    oi, od = oracle.knn(q, t, 2)
    assert np.array_equal(idx, oi) and np.array_equal(bits(dist), bits(od))


def test_many_queries(tc):
    q, t, _ = gen.planted(78, 5000, 3000, 0.5, 0.08)
    idx, dist = tc.knn_match(q, t)
    oi, od = oracle.knn(q, t, 2)
    assert np.array_equal(idx, oi) and np.array_equal(bits(dist), bits(od))
    good, raw = tc.match_features(q, t, 0.8, mutual=True)
    og, orw = oracle.match_features(q, t, 0.8, mutual=True)
    assert good.tobytes() == og.tobytes() and raw.tobytes() == orw.tobytes()


def test_alternating_shapes_and_growth():
    """Buffers grow, the descriptor-block cache is hit and missed, the store grows while in use."""
    m = vsm_b200.Matcher(engine=vsm_b200.ENGINE_TENSOR, scratch_rows=256)
    m.set_profiling(False)
    rng = np.random.default_rng(5)
    db_rows = []
    for it in range(12):
        nq, nt = int(rng.integers(1, 900)), int(rng.integers(2, 1200))
        q, t, _ = gen.planted(200 + it, nq, nt, 0.5, 0.09)
        for rep in range(2):                                   # second call: identical descriptor block
            good, raw = m.match_features(q, t, 0.75, mutual=bool(it & 1))
            og, orw = oracle.match_features(q, t, 0.75, mutual=bool(it & 1))
            assert good.tobytes() == og.tobytes() and raw.tobytes() == orw.tobytes()
        m.add_keyframe(it, t)
        db_rows.append(t)
        db = np.concatenate(db_rows)
        gi, gd = m.search_map_points(q)
        oi, od = oracle.knn(q, db, 2)
        assert np.array_equal(gi, oi) and np.array_equal(bits(gd), bits(od))
    m.close()


def test_batch_of_many_small_pairs(tc):
    rng = np.random.default_rng(8)
    qs, ts = [], []
    for i in range(150):
        nq, nt = int(rng.integers(0, 60)), int(rng.integers(0, 90))
        qs.append(gen.rows(300 + i, 0, 0, nq))
        ts.append(gen.rows(300 + i, 1, 0, nt))
    got = tc.match_batch(qs, ts, 0.9, mutual=True)
    for a, b, g in zip(qs, ts, got):
        og, _ = oracle.match_features(a, b, 0.9, mutual=True)
        assert g.tobytes() == og.tobytes()


def test_fused_exchange_single_rank(tc):
    """The peer-memory exchange kernel with world = 1 (publish to self, flag, merge) must equal
    the plain device search; the multi-rank path is exercised by bench.py --gpus N (checksum)."""
    import torch
    q, db, _ = cases.db_case()
    m = vsm_b200.Matcher(engine=vsm_b200.ENGINE_TENSOR)
    d_db = torch.from_numpy(db).cuda()
    d_q = torch.from_numpy(q).cuda()
    torch.cuda.synchronize()
    m.adopt_device_matrix(d_db.data_ptr(), db.shape[0])
    h = m.xchg_create(0, 1, 512)
    assert len(h) == 64
    m.xchg_connect([h])
    oi = torch.empty((q.shape[0], 2), dtype=torch.int64, device="cuda")
    od = torch.empty((q.shape[0], 2), dtype=torch.float32, device="cuda")
    wi, wd = oracle.knn(q, db, 2)
    for step in range(3):                                   # both parities
        oi.zero_(); od.zero_()
        torch.cuda.synchronize()
        m.db_top2_xchg_device(d_q.data_ptr(), q.shape[0], 1000, oi.data_ptr(), od.data_ptr(), sync=True)
        assert np.array_equal(oi.cpu().numpy(), wi + 1000) and np.array_equal(bits(od.cpu().numpy()), bits(wd))
    with pytest.raises(vsm_b200.VsmError):
        big = torch.zeros((600, 256), device="cuda")
        m.db_top2_xchg_device(big.data_ptr(), 600, 0, oi.data_ptr(), od.data_ptr(), sync=True)   # above nq_cap
    m.close()


def _unit_rows(rng, n):
    x = rng.standard_normal((n, 256), dtype=np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x


def test_baseline_config2_full_size(tc):
    """BASELINE configs[2] at full size: 1000 queries x 500 keyframes x 1000 descriptors (500K rows),
    global top-2, every index and every distance bit against the CPU oracle."""
    rng = np.random.default_rng(2024)
    db = _unit_rows(rng, 500_000)
    q = _unit_rows(rng, 1000)
    rows = rng.integers(0, db.shape[0], 200)
    v = db[rows] + 0.05 * rng.standard_normal((200, 256), dtype=np.float32)      # 20 % planted re-observations
    q[:200] = v / np.linalg.norm(v, axis=1, keepdims=True)
    tc.clear_store()
    for kf in range(0, 500, 50):                                   # 10 uploads of 50 keyframes' worth
        tc.add_keyframe(kf, db[kf * 1000:(kf + 50) * 1000])
    idx, dist = tc.search_map_points(q)
    oi, od = oracle.knn(q, db, 2)
    assert np.array_equal(idx, oi) and np.array_equal(bits(dist), bits(od))
    assert np.array_equal(idx[:200, 0], rows)
    tc.clear_store()


def test_baseline_config4_full_size(tc):
    """BASELINE configs[4] at full size: 64 ragged pairs, sizes U{200..2048}, mutual-NN + ratio 0.75."""
    rng = np.random.default_rng(0)
    sizes = rng.integers(200, 2049, size=(64, 2))
    qs, ts = [], []
    for nq, nt in sizes:
        a = _unit_rows(rng, int(nq))
        b = _unit_rows(rng, int(nt))
        k = int(0.6 * min(nq, nt))
        v = a[:k] + 0.08 * rng.standard_normal((k, 256), dtype=np.float32)
        b[:k] = v / np.linalg.norm(v, axis=1, keepdims=True)
        qs.append(a)
        ts.append(b)
    got = tc.match_batch(qs, ts, 0.75, mutual=True)
    total = 0
    for a, b, g in zip(qs, ts, got):
        og, _ = oracle.match_features(a, b, 0.75, mutual=True)
        assert g.tobytes() == og.tobytes()
        total += len(g)
    assert total > 64 * 50


def test_pair_engine_db_batch_and_mutual(pair):
    """The cta_group::2 engine through the other entry points: keyframe store (odd number of query
    tiles -> a dummy partner tile), segmented search, ragged batch with mutual filter."""
    q, db, seg_off = cases.db_case()                        # 300 queries = 3 query tiles
    pair.clear_store()
    for s in range(len(seg_off) - 1):
        pair.add_keyframe(s, db[seg_off[s]:seg_off[s + 1]])
    idx, dist = pair.search_map_points(q)
    oi, od = oracle.knn(q, db, 2)
    assert np.array_equal(idx, oi) and np.array_equal(bits(dist), bits(od))
    counts, lists = pair.detect_candidates(q, 0.75)
    oc, ol = oracle.segmented(q, db, seg_off, 0.75)
    assert np.array_equal(counts, oc)
    for a, b in zip(lists, ol):
        assert a.tobytes() == b.tobytes()
    pair.clear_store()
    qs = [gen.planted(500 + i, n, m_, 0.5, 0.08)[0] for i, (n, m_) in enumerate([(700, 900), (129, 300), (1, 1), (2048, 257)])]
    ts = [gen.planted(500 + i, n, m_, 0.5, 0.08)[1] for i, (n, m_) in enumerate([(700, 900), (129, 300), (1, 1), (2048, 257)])]
    got = pair.match_batch(qs, ts, 0.75, mutual=True)
    for a, b, g in zip(qs, ts, got):
        og, _ = oracle.match_features(a, b, 0.75, mutual=True)
        assert g.tobytes() == og.tobytes()
    q2, t2, _ = gen.planted(79, 1500, 40000, 0.5, 0.08)     # several 64-tile units per cluster
    idx, dist = pair.knn_match(q2, t2)
    oi, od = oracle.knn(q2, t2, 2)
    assert np.array_equal(idx, oi) and np.array_equal(bits(dist), bits(od))


def _check_compact(m, q, db, seg_off, frame_ids, cur_id, ratio, min_gap, every, min_matches):
    st, lists, _ = m.loop_detect_compact(cur_id, q, ratio, min_gap=min_gap, every=every, min_matches=min_matches)
    ost, ol = oracle.loop_detect(q, db, seg_off, frame_ids, cur_id, ratio, min_gap=min_gap, every=every)
    assert np.array_equal(st, ost)
    want = {s for s in range(len(ost)) if ost[s] >= max(min_matches, 1)}
    assert set(lists) == want
    for s in want:
        assert lists[s].tobytes() == ol[s].tobytes(), s
    return ost


def test_loop_detect_compact_equals_reference_loop(tc):
    """vsm_loop_detect_compact (fused ratio dismissal in the tensor-core epilogue, exact scans of the open
    pairs, >= MIN_MATCHES gate and packing on the device) against LoopCloser::detect's loop restated
    (src/LoopCloser.cpp:43-62): status of every keyframe, and the good_matches of every keyframe that
    passes the gate, byte for byte."""
    q, db, seg_off = cases.db_case()
    nkf = len(seg_off) - 1
    frame_ids = np.arange(nkf, dtype=np.int32) * 40
    tc.clear_store()
    for s in range(nkf):
        tc.add_keyframe(int(frame_ids[s]), db[seg_off[s]:seg_off[s + 1]])
    cur_id = int(frame_ids[-1]) + 100
    for every in (1, 2, 5):
        for min_matches in (30, 1, 0, 45, 1000):
            ost = _check_compact(tc, q, db, seg_off, frame_ids, cur_id, 0.75, 200, every, min_matches)
    assert (ost >= 30).any()
    for ratio in (0.6, 0.9, 1.0):
        _check_compact(tc, q, db, seg_off, frame_ids, cur_id, ratio, 0, 1, 10)
    # ragged query counts (not a multiple of 32 / 128), one query, scaled rows
    _check_compact(tc, q[:131], db, seg_off, frame_ids, cur_id, 0.75, 0, 1, 5)
    _check_compact(tc, q[:1], db, seg_off, frame_ids, cur_id, 0.75, 0, 1, 0)
    _check_compact(tc, np.ascontiguousarray(q * np.float32(1.7)), db, seg_off, frame_ids, cur_id, 0.75, 0, 1, 5)
    # repeated call (descriptor block reused), then a changed store
    _check_compact(tc, q, db, seg_off, frame_ids, cur_id, 0.75, 0, 1, 30)
    _check_compact(tc, q, db, seg_off, frame_ids, cur_id, 0.75, 0, 1, 30)
    tc.remove_frame(int(tc.keyframes()[2]))
    keep = np.ones(db.shape[0], bool)
    keep[seg_off[2]:seg_off[3]] = False
    seg2 = np.concatenate([[0], np.cumsum(np.delete(np.diff(seg_off), 2))]).astype(np.int64)
    _check_compact(tc, q, db[keep], seg2, np.delete(frame_ids, 2), cur_id, 0.75, 0, 1, 30)
    tc.clear_store()


def test_loop_detect_compact_overflow_takes_the_record_path():
    """More open pairs than the compact buffers hold (a scene full of matches; here the capacity is
    shrunk): the call falls back to the record-based search and returns the same answer."""
    q, db, seg_off = cases.db_case()
    nkf = len(seg_off) - 1
    frame_ids = np.arange(nkf, dtype=np.int32)
    with vsm_b200.Matcher(engine=vsm_b200.ENGINE_TENSOR, pair_cap=16) as m:
        for s in range(nkf):
            m.add_keyframe(int(frame_ids[s]), db[seg_off[s]:seg_off[s + 1]])
        _check_compact(m, q, db, seg_off, frame_ids, 1000, 0.75, 0, 1, 30)
    with vsm_b200.Matcher(engine=vsm_b200.ENGINE_TENSOR, work_cap=64) as m:
        for s in range(nkf):
            m.add_keyframe(int(frame_ids[s]), db[seg_off[s]:seg_off[s + 1]])
        _check_compact(m, q, db, seg_off, frame_ids, 1000, 0.75, 0, 1, 30)


def test_loop_detect_compact_at_baseline_config2_size(tc):
    """BASELINE configs[2] in LoopCloser's own form, FULL size: 1000 queries against 500 keyframes x 1000
    descriptors, every keyframe eligible; 3 keyframes are re-observed (loop candidates), a few more share
    a handful of rows.  Status of all 500 keyframes and every surviving list against the oracle."""
    import torch
    nkf, rows_kf, nq = 500, 1000, 1000
    db = np.concatenate([gen.rows(600 + k // 50, k % 50, 0, rows_kf) for k in range(nkf)])
    vq = gen.int_rows(611, 0, 0, nq).copy()
    vdb = lambda kf, n: np.rint(db[kf * rows_kf: kf * rows_kf + n].astype(np.float64) * 3300).astype(np.int64)
    vq[0:300] = 1000 * vdb(123, 300) + 1000 * gen.int_rows(612, 0, 0, 300)
    vq[300:500] = 1000 * vdb(124, 200) + 1100 * gen.int_rows(613, 0, 0, 200)
    vq[500:540] = 1000 * vdb(400, 40) + 900 * gen.int_rows(614, 0, 0, 40)
    vq[540:552] = 1000 * vdb(77, 12) + 900 * gen.int_rows(615, 0, 0, 12)
    q = gen._normalize_int(vq)
    seg_off = np.arange(nkf + 1, dtype=np.int64) * rows_kf
    d_db = torch.from_numpy(db).cuda()
    tc.clear_store()
    tc.adopt_device_matrix(d_db.data_ptr(), db.shape[0], seg_off)
    frame_ids = np.arange(nkf, dtype=np.int32)
    ost = _check_compact(tc, q, db, seg_off, frame_ids, 10_000, 0.75, 200, 1, 30)
    assert (ost >= 30).sum() >= 3 and ((ost > 0) & (ost < 30)).any()
    st = tc.stats()
    assert st["kernel_launches"] == 5
    # the reference's every-5th rule over the same list
    _check_compact(tc, q, db, seg_off, frame_ids, 10_000, 0.75, 200, 5, 30)
    tc.clear_store()


def test_resident_map_point_table(tc):
    """The map-point searches of Slam::try_pnp_recovery (src/Slam.cpp:546-574: every valid point) and
    Slam::handle_loop_closure (:744-774: valid points with an observation within 30 frames of the matched
    keyframe) against a point table that lives on the device: births from host descriptors and from rows
    of a stored frame (:1339, :1563), add_observation (:463), set_valid (:496, :1119-1123).  The model
    below IS the reference's loop: stack the selected descriptors in id order, knnMatch, mp_ids_vec."""
    rng = np.random.default_rng(11)
    tc.clear_store()
    tc.clear_map_points()
    desc, valid, obs = [], [], []                       # the model: per point descriptor, valid_, observation frames

    def check(q, near=-1, rng_=30):
        sel = [i for i in range(len(desc)) if valid[i] and (near < 0 or any(abs(f - near) < rng_ for f in obs[i]))]
        gi, gd, ns = tc.search_resident_map_points(q, near, rng_)
        assert ns == len(sel)
        if not sel:
            assert (gi == -1).all()
            return
        oi, od = oracle.knn(q, np.stack([desc[i] for i in sel]), 2)
        ids = np.array(sel, np.int64)
        want = np.where(oi >= 0, ids[np.maximum(oi, 0)], -1)
        assert np.array_equal(gi, want) and np.array_equal(bits(gd), bits(od))

    q0 = gen.rows(900, 9, 0, 150)
    check(q0)                                            # empty table
    # keyframes arrive; some of their keypoints become map points (depth points: one observation;
    # triangulated points: a second observation in the previous keyframe)
    frames = gen.video(901, 12, 300)
    handles = []
    for f, d in enumerate(frames):
        fid = 20 * f
        h = tc.add_keyframe(fid, d)
        handles.append(h)
        kp = np.sort(rng.choice(300, size=60, replace=False)).astype(np.int32)
        first = tc.add_map_points_from_frame(h, kp)
        assert first == len(desc)
        for k in kp:
            desc.append(d[k]); valid.append(True); obs.append([fid])
        if f > 0:
            tri = np.arange(first, first + 60, 3, dtype=np.int32)
            tc.observe_map_points(tri, 20 * (f - 1))
            for p in tri:
                obs[p].append(20 * (f - 1))
    # a batch born from host descriptors (no stored frame)
    extra = gen.rows(902, 0, 0, 77)
    first = tc.add_map_points(extra, 1000)
    assert first == len(desc)
    for r in extra:
        desc.append(r); valid.append(True); obs.append([1000])
    assert tc.map_point_info() == (len(desc), len(desc), sum(len(o) for o in obs))
    # queries: noisy re-observations of some points + fresh rows
    pts = rng.choice(len(desc), size=100, replace=False)
    vq = gen.int_rows(903, 0, 0, 160).copy()
    vq[:100] = 1000 * np.rint(np.stack([desc[p] for p in pts]).astype(np.float64) * 3300).astype(np.int64) + 700 * gen.int_rows(904, 0, 0, 100)
    q = gen._normalize_int(vq)
    check(q)
    check(q, near=100)                                   # points seen in frames 71..129: keyframes 80, 100, 120
    check(q, near=100, rng_=1)                           # exactly frame 100
    check(q, near=5000)                                  # nobody: empty selection
    # culling (src/Slam.cpp:1110-1126) and reprojection failures (:496) invalidate points; tracking re-observes others
    dead = rng.choice(len(desc), size=250, replace=False).astype(np.int32)
    tc.set_map_points_valid(dead, False)
    for p in dead:
        valid[p] = False
    tc.set_map_points_valid(dead[:10], False)            # idempotent
    seen = rng.choice(len(desc), size=120, replace=False).astype(np.int32)
    tc.observe_map_points(seen, 700)
    for p in seen:
        obs[p].append(700)
    assert tc.map_point_info() == (len(desc), sum(valid), sum(len(o) for o in obs))
    check(q)
    check(q, near=700, rng_=5)
    check(q, near=100)
    tc.set_map_points_valid(dead[:40], True)             # and back
    for p in dead[:40]:
        valid[p] = True
    check(q, near=90)
    # all but one point invalid: a one-row stacked matrix has no second neighbour (the size()>=2 guard, :570)
    tc.set_map_points_valid(np.arange(len(desc), dtype=np.int32), False)
    tc.set_map_points_valid(np.array([5], np.int32), True)
    valid[:] = [False] * len(desc)
    valid[5] = True
    check(q)
    tc.clear_map_points()
    assert tc.map_point_info() == (0, 0, 0)
    tc.clear_store()


def test_append_records_option_equals_oracle(appendm):
    """The append-record epilogue (an option, off by default) on match-heavy small problems: pair matching with
    ratio and mutual tests, a ragged batch, a tracking sequence, rows with unequal norms, and a train set whose
    near-duplicates overflow the record (exact re-scan of the slice)."""
    for name in ("pair_1000_video", "pair_777x1301", "pair_2048x200", "mutual_conflict", "neardup_db", "scaled", "nt2"):
        q, t = cases.PAIR_CASES[name]()
        for mutual in (False, True):
            good, raw = appendm.match_features(q, t, 0.75, mutual=mutual)
            og, orw = oracle.match_features(q, t, 0.75, mutual=mutual)
            assert good.tobytes() == og.tobytes() and raw.tobytes() == orw.tobytes(), (name, mutual)
    qs, ts = [], []
    for p in range(7):
        a, b, _ = gen.planted(700 + p, 150 + 137 * p, 260 + 211 * p, 0.6, 0.08)
        qs.append(a)
        ts.append(b)
    res = appendm.match_batch(qs, ts, 0.75, mutual=True)
    for p in range(7):
        og, _ = oracle.match_features(qs[p], ts[p], 0.75, mutual=True)
        assert res[p].tobytes() == og.tobytes(), p
    # 200 near-copies of one row: more values pass the threshold than a record holds
    base = gen.int_rows(55, 1, 0, 900).copy()
    base[300:500] = 1000 * base[10] + gen.int_rows(55, 2, 0, 200)
    t = gen._normalize_int(base)
    vq = gen.int_rows(55, 0, 0, 64).copy()
    vq[:32] = 1000 * gen.int_rows(55, 1, 10, 1) + 20 * gen.int_rows(55, 3, 0, 32)
    q = gen._normalize_int(vq)
    gi, gd = appendm.knn_match(q, t)
    oi, od = oracle.knn(q, t, 2)
    assert np.array_equal(gi, oi) and np.array_equal(bits(gd), bits(od))
    assert appendm.stats()["flagged_slices"] > 0


def test_strided_entry_points(tc):
    """cv::Mat ROIs: rows 1280 / 2048 bytes apart go through the *_strided calls (one strided DMA) and give
    the answers of the packed matrices; a stride below 1024 bytes is refused."""
    import ctypes as C
    q, t, _ = gen.planted(91, 333, 410, 0.6, 0.08)
    wq = np.zeros((q.shape[0], 320), np.float32)
    wq[:, :256] = q
    wt = np.zeros((t.shape[0], 512), np.float32)
    wt[:, 256:] = t                                              # the ROI starts in the middle of a wider row
    lib, h = tc.lib, tc.handle
    idx = np.empty((q.shape[0], 2), np.int32)
    dist = np.empty((q.shape[0], 2), np.float32)
    t_ptr = wt.ctypes.data + 256 * 4
    assert lib.vsm_knn2_strided(h, wq.ctypes.data, q.shape[0], 320 * 4, C.c_void_p(t_ptr), t.shape[0], 512 * 4,
                                idx.ctypes.data, dist.ctypes.data) == 0
    oi, od = oracle.knn(q, t, 2)
    assert np.array_equal(idx, oi) and np.array_equal(bits(dist), bits(od))
    good = np.zeros(q.shape[0], vsm_b200.DMATCH)
    ng = C.c_int32(0)
    assert lib.vsm_match_pair_strided(h, wq.ctypes.data, q.shape[0], 320 * 4, C.c_void_p(t_ptr), t.shape[0], 512 * 4,
                                      C.c_float(0.75), 1, good.ctypes.data, C.byref(ng), None, None) == 0
    og, _ = oracle.match_features(q, t, 0.75, mutual=True)
    assert good[:ng.value].tobytes() == og.tobytes()
    assert lib.vsm_knn2_strided(h, wq.ctypes.data, q.shape[0], 1000, C.c_void_p(t_ptr), t.shape[0], 512 * 4,
                                idx.ctypes.data, dist.ctypes.data) != 0
    # a keyframe added from a strided matrix
    tc.clear_store()
    hk = C.c_int32(-1)
    assert lib.vsm_store_add_strided(h, 7, C.c_void_p(t_ptr), t.shape[0], 512 * 4, C.byref(hk)) == 0
    g2, _ = tc.match_to_keyframe(hk.value, q, 0.75)
    o2, _ = oracle.match_features(t, q, 0.75)
    assert g2.tobytes() == o2.tobytes()
    tc.clear_store()


def test_store_edge_cases(tc):
    """Empty and one-row keyframes, removal of everything, handles of removed frames, searches on an empty store."""
    tc.clear_store()
    q = gen.rows(95, 0, 0, 40)
    gi, gd = tc.search_map_points(q)
    assert (gi == -1).all()
    st, lists, _ = tc.loop_detect_compact(1000, q, 0.75, min_gap=0, every=1, min_matches=1)
    assert len(st) == 0 and lists == {}
    z = np.zeros((0, 256), np.float32)
    h0 = tc.add_keyframe(0, z)
    h1 = tc.add_keyframe(1, gen.rows(95, 1, 0, 1))
    h2 = tc.add_keyframe(2, gen.rows(95, 2, 0, 300))
    st, lists, _ = tc.loop_detect_compact(1000, q, 0.75, min_gap=0, every=1, min_matches=0)
    assert list(st) == [-1, 0, 0] and lists == {}               # empty: skipped (LoopCloser.cpp:45); one row: no 2-entry list
    gi, gd = tc.search_map_points(q)
    db = np.concatenate([gen.rows(95, 1, 0, 1), gen.rows(95, 2, 0, 300)])
    oi, od = oracle.knn(q, db, 2)
    assert np.array_equal(gi, oi) and np.array_equal(bits(gd), bits(od))
    tc.remove_frame(h1)
    gi, gd = tc.search_map_points(q)                             # the run now starts at row 1
    oi, od = oracle.knn(q, db[1:], 2)
    assert np.array_equal(gi, oi + 1) and np.array_equal(bits(gd), bits(od))
    with pytest.raises(vsm_b200.VsmError):
        tc.remove_frame(h1)
    with pytest.raises(vsm_b200.VsmError):
        tc.promote(12345)
    tc.remove_frame(h2)
    tc.remove_frame(h0)
    assert tc.store_info() == (0, 0)
    assert (tc.search_map_points(q)[0] == -1).all()
    tc.clear_store()


def _band_case(seed, nq, nt, ratio):
    """Planted pairs whose noise is spread so that d0 / d1 covers [0.5 * ratio, 1.1]: plenty of queries land inside
    the narrow band around `ratio` where the tile top-2 path cannot decide from bounds and re-scores exactly."""
    vt = gen.int_rows(seed, 1, 0, nt)
    vq = gen.int_rows(seed, 0, 0, nq).copy()
    rows = gen._perm(seed, 7, nt)[np.arange(nq) % nt]
    noise = gen.int_rows(seed, 2, 0, nq)
    amp = (300 + (np.arange(nq) * 2500) // nq).astype(np.int64)[:, None]           # noise amplitude sweeps 0.3 .. 2.8
    vq[:] = 1000 * vt[rows] + (amp * noise) // 1
    return gen._normalize_int(vq), gen._normalize_int(vt)


def test_tile_top2_path_equals_oracle_and_top4_path(tc, top4m):
    """Pair matching without a raw list runs on tile top-2 records (small problems): same survivors as the
    oracle and as the top-4 path -- planted pairs, a noise sweep that fills the undecidable band around the ratio,
    near-duplicate clusters (slices whose second entry is a candidate: re-scan), exact duplicates (ties), train sets
    of 1 / 2 / 3 rows, scaled rows, partial tiles, a train set at the 32-tile limit and one above it."""
    named = ("pair_1000_video", "pair_777x1301", "pair_2048x200", "pair_129x257", "mutual_conflict", "dups", "neardup_db",
             "nt1", "nt2", "nt3", "nq1", "scaled")
    todo = [(n,) + tuple(cases.PAIR_CASES[n]()) for n in named]
    todo.append(("band_075",) + _band_case(91, 1500, 1700, 0.75))
    todo.append(("band_wide",) + _band_case(92, 700, 8192, 0.8))                    # 32 tiles: the largest tile top-2 train set
    todo.append(("above_limit",) + gen.planted(93, 300, 8192 + 256, 0.6, 0.08)[:2])  # 33 tiles: top-4 records
    for name, q, t in todo:
        for ratio in (0.7, 0.75, 0.8, 1.0):
            for mutual in (False, True):
                good, _ = tc.match_features(q, t, ratio, mutual=mutual, want_raw=False)
                og, _ = oracle.match_features(q, t, ratio, mutual=mutual)
                assert good.tobytes() == og.tobytes(), (name, ratio, mutual)
                if ratio == 0.75:
                    g4, _ = top4m.match_features(q, t, ratio, mutual=mutual, want_raw=False)
                    assert g4.tobytes() == og.tobytes(), (name, ratio, mutual, "top-4 records")
    # the point of it: a matching query costs one or two exact distances instead of ~4 per direction
    q, t = cases.PAIR_CASES["pair_1000_video"]()
    tc.match_features(q, t, 0.75, mutual=True, want_raw=False)
    c2 = tc.stats()["candidates"]
    top4m.match_features(q, t, 0.75, mutual=True, want_raw=False)
    c4 = top4m.stats()["candidates"]
    assert 0 < c2 < c4, (c2, c4)
    # near-duplicates: the band cases and the reverse problem's ties go through re-scans of single tile halves
    q, t = cases.PAIR_CASES["neardup_db"]()
    tc.match_features(q, t, 0.75, mutual=True, want_raw=False)
    assert tc.stats()["flagged_slices"] > 0
    # one call that mixes the record kinds: a per-keyframe search over keyframes of 300 rows (tile top-2 records)
    # and one of 9000 rows (above the 32-tile limit: maxima-only / top-4 records) -- both select kernels run
    sizes = [300, 9000, 300, 2, 700]
    seg_off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    vdb = gen.int_rows(95, 1, 0, int(seg_off[-1]))
    vq = gen.int_rows(95, 0, 0, 260).copy()
    vq[:60] = 1000 * vdb[seg_off[1] + 17 * np.arange(60)] + 900 * gen.int_rows(95, 2, 0, 60)      # matches in the big keyframe
    vq[60:120] = 1000 * vdb[seg_off[4] + 5 * np.arange(60)] + 900 * gen.int_rows(95, 3, 0, 60)    # ... and in a small one
    qd, db = gen._normalize_int(vq), gen._normalize_int(vdb)
    with vsm_b200.Matcher() as mm:
        for k in range(len(sizes)):
            mm.add_keyframe(k, db[seg_off[k]:seg_off[k + 1]])
        for _ in range(2):                                         # second call: the adaptive kind may have changed
            c, lists = mm.detect_candidates(qd, 0.75)
            oc, ol = oracle.segmented(qd, db, seg_off, 0.75)
            assert np.array_equal(c, oc) and c[1] > 30 and c[4] > 30
            for k in range(len(sizes)):
                assert lists[k].tobytes() == ol[k].tobytes(), k
    # ragged batch and stored pairs take the same path
    qs, ts = [], []
    for p in range(9):
        a, b = _band_case(300 + p, 100 + 211 * p, 90 + 283 * p, 0.75)
        qs.append(a)
        ts.append(b)
    for mutual in (False, True):
        res = tc.match_batch(qs, ts, 0.75, mutual=mutual)
        for p in range(9):
            og, _ = oracle.match_features(qs[p], ts[p], 0.75, mutual=mutual)
            assert res[p].tobytes() == og.tobytes(), (p, mutual)
