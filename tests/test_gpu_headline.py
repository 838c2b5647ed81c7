"""Parity in the regime of the headline benchmark (BASELINE configs[3]: 2000 queries against a 20M-row
keyframe database, one GPU's share of it here): at >= 1M rows the planner uses 64-tile slices and
64-tile work units, i.e. 8192 columns per slice half -- the 13 packed index bits at their limit.
The CUDA path (through the C ABI, host buffers in and out) is compared with the CPU oracle on sampled
queries: planted and un-planted ones, FIRST AND SECOND neighbour, indices and fp32 distance bits, and
queries aimed at a cluster of near-duplicate rows whose slice overflows the four recorded entries and is
re-scanned exactly."""
import numpy as np
import pytest

from oracle import oracle
import vsm_b200

pytestmark = pytest.mark.gpu

ROWS = 148 * 64 * 256          # 2,424,832 rows = 9472 tiles: 148 ranges of exactly 64 tiles
NQ = 2000


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _unit(torch, n, g):
    x = torch.randn((n, 256), generator=g, device="cuda", dtype=torch.float32)
    return x / x.norm(dim=1, keepdim=True)


@pytest.fixture(scope="module")
def big_db():
    import torch
    g = torch.Generator(device="cuda")
    g.manual_seed(4242)
    db = torch.empty((ROWS, 256), dtype=torch.float32, device="cuda")
    for r0 in range(0, ROWS, 1 << 19):
        n = min(1 << 19, ROWS - r0)
        db[r0:r0 + n] = _unit(torch, n, g)
    q = _unit(torch, NQ, g)
    # 400 planted queries: noisy re-observations of DB rows spread over the whole range
    planted_rows = torch.arange(400, device="cuda") * (ROWS // 400) + 17
    v = db[planted_rows] + 0.05 * torch.randn((400, 256), generator=g, device="cuda")
    q[:400] = v / v.norm(dim=1, keepdim=True)
    # a cluster of 400 near-duplicates of one row (distance ~1e-2 apart) in the middle of the database,
    # contiguous: both column halves of its slice see > 4 entries above any threshold
    centre = db[1_000_000].clone()
    v = centre[None, :] + 6e-4 * torch.randn((400, 256), generator=g, device="cuda")
    db[1_203_000:1_203_400] = v / v.norm(dim=1, keepdim=True)      # between two planted rows
    v = centre[None, :] + 1e-3 * torch.randn((32, 256), generator=g, device="cuda")
    q[400:432] = v / v.norm(dim=1, keepdim=True)
    torch.cuda.synchronize()
    host = db.cpu().numpy()
    yield db, host, q.cpu().numpy(), planted_rows.cpu().numpy()


SAMPLE = np.concatenate([np.arange(0, 400, 8), np.arange(400, 432), np.arange(432, 2000, 34)])   # 50 + 32 + 47


@pytest.mark.parametrize("seg_tiles", [0, 64])
def test_one_shard_of_the_20m_search_equals_oracle(big_db, seg_tiles):
    db, host, q, planted_rows = big_db
    with vsm_b200.Matcher(engine=vsm_b200.ENGINE_TENSOR, seg_tiles=seg_tiles) as m:
        m.adopt_device_matrix(db.data_ptr(), ROWS)
        gi, gd = m.search_map_points(q)                       # vsm_db_top2: host queries in, host top-2 out
        st = m.stats()
    assert st["flagged_slices"] > 0, "the near-duplicate cluster must overflow a slice's four entries"
    assert (gi[:400, 0] == planted_rows).all()
    qs = np.ascontiguousarray(q[SAMPLE])
    oi, od = oracle.knn(qs, host, 2)
    assert np.array_equal(gi[SAMPLE], oi), "indices (first and second neighbour) differ from the oracle"
    assert np.array_equal(bits(gd[SAMPLE]), bits(od)), "fp32 distance bits differ from the oracle"
    # the cluster queries' answers lie inside the cluster (or its centre row)
    c = gi[400:432]
    assert (((c >= 1_203_000) & (c < 1_203_400)) | (c == 1_000_000)).all()


def test_sharded_merge_of_two_half_databases_equals_oracle(big_db):
    """The same database as two shards on one GPU (two contexts, row offsets), merged by key: what two
    ranks of the sharded search compute -- identical to the oracle over the whole database."""
    import torch
    db, host, q, _ = big_db
    half = ROWS // 2 + 12345                                   # not tile aligned
    d_q = torch.from_numpy(q).cuda()
    keys = torch.zeros((2, NQ, 2), dtype=torch.int64, device="cuda")
    ms = []
    for r, (lo, hi) in enumerate(((0, half), (half, ROWS))):
        m = vsm_b200.Matcher(engine=vsm_b200.ENGINE_TENSOR)
        m.adopt_device_matrix(db[lo:hi].data_ptr(), hi - lo)
        m.db_top2_keys_device(d_q.data_ptr(), NQ, lo, keys[r].data_ptr(), sync=True)
        ms.append(m)
    oi_t = torch.empty((NQ, 2), dtype=torch.int64, device="cuda")
    od_t = torch.empty((NQ, 2), dtype=torch.float32, device="cuda")
    ms[0].merge_keys_device(keys.data_ptr(), 2, NQ, oi_t.data_ptr(), od_t.data_ptr(), sync=True)
    for m in ms:
        m.close()
    gi, gd = oi_t.cpu().numpy(), od_t.cpu().numpy()
    qs = np.ascontiguousarray(q[SAMPLE])
    oi, od = oracle.knn(qs, host, 2)
    assert np.array_equal(gi[SAMPLE], oi) and np.array_equal(bits(gd[SAMPLE]), bits(od))


def test_big_search_shortens_its_slices_on_clustered_data():
    """A database holding a large cluster of near-copies (keyframes of one place, stored next to each other):
    every query aimed at it overflows the four recorded entries of the cluster's slices.  A big search
    starts with 64-tile slices and re-scans the overflowing ones exactly; seeing the overflow count, the
    next search plans 16-tile slices (4x cheaper re-scans) -- and returns to 64 when queries stop
    overflowing.  The feedback also flows when the plan cache hits (same shapes, other queries).
    Exact answers in every regime."""
    import torch
    rows, nq = 1_200_000, 512
    g = torch.Generator(device="cuda")
    g.manual_seed(99)
    db = _unit(torch, rows, g)
    c0 = _unit(torch, 1, g)
    x = c0 + 0.05 * torch.randn((40_000, 256), generator=g, device="cuda")
    db[500_000:540_000] = x / x.norm(dim=1, keepdim=True)
    x = c0 + 0.05 * torch.randn((nq, 256), generator=g, device="cuda")
    q_hard = (x / x.norm(dim=1, keepdim=True)).cpu().numpy()            # aimed at the cluster
    q_easy = _unit(torch, nq, g).cpu().numpy()                          # anywhere
    host = db.cpu().numpy()
    sample = np.arange(0, nq, 8)
    want = {}
    for name, q in (("hard", q_hard), ("easy", q_easy)):
        want[name] = oracle.knn(np.ascontiguousarray(q[sample]), host, 2)

    def search(m, name, q):
        gi, gd = m.search_map_points(q)
        oi, od = want[name]
        assert np.array_equal(gi[sample], oi) and np.array_equal(bits(gd[sample]), bits(od)), name
        st = m.stats()
        return st["slice_tiles"], st["flagged_slices"]

    with vsm_b200.Matcher(engine=vsm_b200.ENGINE_TENSOR) as m:
        m.adopt_device_matrix(db.data_ptr(), rows)
        seen = [search(m, "hard", q_hard) for _ in range(3)]
        assert seen[0][0] == 64 and seen[0][1] * 8 > nq, seen
        assert seen[1][0] == 16 and seen[2][0] == 16, seen
        # other queries, same shapes: the plan cache hits, the overflow feedback must still flow
        seen = [search(m, "easy", q_easy) for _ in range(3)]
        assert [s_[0] for s_ in seen] == [16, 64, 64] and seen[0][1] * 64 < nq, seen
        seen = [search(m, "hard", q_hard) for _ in range(2)]
        assert [s_[0] for s_ in seen] == [64, 16], seen
