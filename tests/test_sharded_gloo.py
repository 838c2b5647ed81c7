"""The N > 1 path on CPU: two processes, gloo backend.  Each rank answers its database shard
with the CPU oracle (standing in for the per-rank CUDA search), then runs the product's own
exchange step (sharded.gather_and_merge) and must reproduce the whole-database top-2 exactly,
including the lowest-global-index tie-break across ranks."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import cases, oracle
import vsm_b200


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = vsm_b200.load_sharded()
    if case == "db":
        q, db, seg_off = cases.db_case()
    else:                                   # duplicates straddling the shard boundary
        q, db = cases.PAIR_CASES["dups"]()
        seg_off = np.array([0, 20, 42, 64], np.int64)
    parts = sh.partition_keyframes(seg_off, world)
    _, _, r0, r1 = parts[rank]
    li, ld = oracle.knn(q, db[r0:r1], 2)
    li = np.where(li >= 0, li + r0, -1)

    def merge(g_idx, g_dist):
        mi, md = oracle.merge_top2(g_idx.numpy(), g_dist.numpy())
        return torch.from_numpy(mi), torch.from_numpy(md)

    oi, od = sh.gather_and_merge(torch.from_numpy(li), torch.from_numpy(ld), world, None, merge)
    wi, wd = oracle.knn(q, db, 2)
    ok = np.array_equal(oi.numpy(), wi) and np.array_equal(od.numpy().view(np.uint32), wd.view(np.uint32))
    out[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def _loop_worker(rank, world, port, out):
    """LoopCloser::detect over a partitioned keyframe list: per-rank eligibility from the product's
    loop_checked_before, per-keyframe matching by the oracle (standing in for libvsm), status
    concatenation by the product's concat_keyframe_status."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = vsm_b200.load_sharded()
    q, db, seg_off = cases.db_case()
    nkf = len(seg_off) - 1
    frame_ids = [30 * s for s in range(nkf)]
    counts = np.diff(seg_off)
    parts = sh.partition_keyframes(seg_off, world)
    k0, k1, r0, r1 = parts[rank]
    ok = True
    for cur_id, gap, every in ((900, 200, 5), (900, 200, 1), (650, 100, 3), (100, 200, 5)):
        before = sh.loop_checked_before(cur_id, frame_ids, counts, gap, k0)
        st, _ = oracle.loop_detect(q, db[r0:r1], seg_off[k0:k1 + 1] - r0, frame_ids[k0:k1], cur_id, 0.75, gap, every,
                                   checked0=before)
        whole = sh.concat_keyframe_status(torch.from_numpy(st), [p[1] - p[0] for p in parts], world)
        want, _ = oracle.loop_detect(q, db, seg_off, frame_ids, cur_id, 0.75, gap, every)
        ok = ok and np.array_equal(whole.numpy(), want)
    out[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_loop_detect_equals_whole():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_loop_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def test_loop_checked_before_counts_gap_and_empty():
    sh = vsm_b200.load_sharded()
    ids = [0, 30, 60, 90, 120, 150]
    counts = [10, 0, 10, 10, 10, 10]
    # cur 200, gap 100: keyframes with id <= 100 pass the gap test -> 0, 30(empty), 60, 90
    assert sh.loop_checked_before(200, ids, counts, 100, 0) == 0
    assert sh.loop_checked_before(200, ids, counts, 100, 2) == 1
    assert sh.loop_checked_before(200, ids, counts, 100, 4) == 3
    assert sh.loop_checked_before(200, ids, counts, 100, 6) == 3


@pytest.mark.parametrize("case", ["db", "dups"])
def test_two_rank_search_equals_whole(case):
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), case, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def test_partition_keeps_keyframes_whole():
    sh = vsm_b200.load_sharded()
    _, _, seg_off = cases.db_case()
    for world in (1, 2, 3, 8):
        parts = sh.partition_keyframes(seg_off, world)
        assert parts[0][0] == 0 and parts[-1][1] == len(seg_off) - 1
        for (a0, a1, r0, r1), nxt in zip(parts, parts[1:] + [None]):
            assert r0 == seg_off[a0] and r1 == seg_off[a1] and a0 <= a1
            if nxt:
                assert nxt[0] == a1
        rows = [p[3] - p[2] for p in parts]
        assert sum(rows) == seg_off[-1]
        if world <= 3:
            assert max(rows) - min(rows) <= 2 * int(np.diff(seg_off).max())
