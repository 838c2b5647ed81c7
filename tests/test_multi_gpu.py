"""The partitioned keyframe database on REAL GPUs: one process per GPU, NCCL process group, both
exchange flavours (fused peer-memory kernel, NCCL all-gather + merge kernel) and the
per-keyframe LoopCloser search.  Needs >= 2 GPUs in the box (`gpurun --gpus 2`); skipped
otherwise.  Every rank compares with the CPU oracle over the WHOLE database, bit for bit."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:       # noqa: BLE001
        return 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, exchange, out):
    import torch
    import torch.distributed as dist
    from oracle import cases, oracle
    import vsm_b200

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    sh = vsm_b200.load_sharded()
    q, db, seg_off = cases.db_case()
    nkf = len(seg_off) - 1
    frame_ids = [30 * s for s in range(nkf)]
    parts = sh.partition_keyframes(seg_off, world)
    k0, k1, r0, r1 = parts[rank]
    sdb = sh.ShardedDB(rank, rank, world, exchange=exchange)
    shard = torch.from_numpy(db[r0:r1]).cuda().contiguous()
    sdb.adopt(shard, r0, np.asarray(seg_off[k0:k1 + 1] - r0, np.int64), frame_ids[k0:k1])
    sdb.set_keyframe_table(frame_ids, np.diff(seg_off), [p[0] for p in parts] + [nkf])
    fails = []
    if sdb.exchange != exchange:
        fails.append(f"exchange fell back to {sdb.exchange}: {sdb.exchange_note}")

    # global top-2 (stacked search, src/Slam.cpp:546-574) through host buffers, three times (the
    # peer-memory exchange alternates two buffer parities)
    wi, wd = oracle.knn(q, db, 2)
    hq = torch.from_numpy(q).pin_memory()
    for rep in range(3):
        hi, hd = sdb.search_host(hq)
        sdb.stream.synchronize()
        if not np.array_equal(hi.numpy(), wi.astype(np.int64)):
            fails.append(f"rep {rep}: indices differ")
        if not np.array_equal(hd.numpy().view(np.uint32), wd.view(np.uint32)):
            fails.append(f"rep {rep}: distance bits differ")
    # a smaller batch afterwards (buffers are re-sized; ragged last query tile)
    hi, hd = sdb.search_host(hq[:77].clone().pin_memory())
    sdb.stream.synchronize()
    if not (np.array_equal(hi.numpy(), wi[:77].astype(np.int64)) and
            np.array_equal(hd.numpy().view(np.uint32), wd[:77].view(np.uint32))):
        fails.append("77-query batch differs")

    # LoopCloser::detect over the partitioned list (src/LoopCloser.cpp:43-62)
    for cur_id, gap, every in ((900, 200, 5), (650, 100, 3), (900, 200, 1)):
        ost, ol = oracle.loop_detect(q, db, seg_off, frame_ids, cur_id, 0.75, gap, every)
        whole, mine = sdb.loop_detect(cur_id, q, 0.75, gap, every)
        if not np.array_equal(whole.numpy(), ost):
            fails.append(f"loop_detect status differs (cur {cur_id})")
        for g in range(k0, k1):
            if ost[g] >= 0:
                want = ol[g].copy()
                want["imgIdx"] = g - k0
                if g not in mine or mine[g].tobytes() != want.tobytes():
                    fails.append(f"loop_detect list of keyframe {g} differs")
            elif g in mine:
                fails.append(f"keyframe {g} should have been skipped")
        # the compact form (gate on the device) and the host-buffer collective search (one C-ABI call per rank)
        whole_c, mine_c = sdb.loop_detect_compact(cur_id, q, 0.75, gap, every, min_matches=20)
        if not np.array_equal(whole_c.numpy(), ost):
            fails.append(f"loop_detect_compact status differs (cur {cur_id})")
        want_c = {g for g in range(k0, k1) if ost[g] >= 20}
        if set(mine_c) != want_c:
            fails.append(f"loop_detect_compact candidates differ (cur {cur_id})")
        for g in want_c & set(mine_c):
            want = ol[g].copy()
            want["imgIdx"] = g
            if mine_c[g].tobytes() != want.tobytes():
                fails.append(f"loop_detect_compact list of keyframe {g} differs")
    ai, ad = sdb.search_host_abi(q)
    if not (np.array_equal(ai, wi.astype(np.int64)) and np.array_equal(ad.view(np.uint32), wd.view(np.uint32))):
        fails.append("search_host_abi (vsm_db_top2_xchg) differs")
    # more ranks than keyframes: one rank holds an empty shard and still takes part in the exchange
    one = np.array([0, int(seg_off[7 + 1] - seg_off[7])], np.int64)
    db1 = db[seg_off[7]:seg_off[8]]
    p1 = sh.partition_keyframes(one, world)
    a0, a1, b0, b1 = p1[rank]
    sdb.adopt(torch.from_numpy(db1[b0:b1]).cuda().contiguous(), b0, np.asarray(one[a0:a1 + 1] - b0, np.int64))
    w1i, w1d = oracle.knn(q, db1, 2)
    hi, hd = sdb.search_host(hq)
    sdb.stream.synchronize()
    if not (np.array_equal(hi.numpy(), w1i.astype(np.int64)) and np.array_equal(hd.numpy().view(np.uint32), w1d.view(np.uint32))):
        fails.append("single-keyframe database over two ranks differs")
    if sorted(b1 - b0 for _, _, b0, b1 in p1)[0] != 0:
        fails.append("expected one empty shard")
    out[rank] = fails
    dist.barrier()
    sdb.close()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs in one box")
@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
def test_two_gpu_sharded_search_and_loop_detect(exchange):
    import torch.multiprocessing as mp
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), exchange, out), nprocs=world, join=True)
        assert dict(out) == {0: [], 1: []}
