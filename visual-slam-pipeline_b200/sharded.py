"""Loop-closure search over a keyframe descriptor database partitioned across the GPUs of one box.

One process per GPU (torch.distributed, NCCL).  Each rank holds a contiguous block of whole
keyframes on its device and answers a query batch with its LOCAL exact top-2 (libvsm.so:
tcgen05 pass + fp32 re-score); a small all-gather moves every rank's [nq][2] (index, distance)
list to every rank and a merge kernel orders them by (distance, global index) -- the same
order one pass over the whole database gives (stacked-matrix search of src/Slam.cpp:546-574 /
744-774 of the reference, at the scale of BASELINE configs[2] and [3]).

torch is used here for device buffers, streams and the process group only.
"""
import torch

from .matcher import Matcher, ENGINE_TENSOR


def partition_keyframes(seg_off, world):
    """Split keyframes (row offsets seg_off[0..nkf]) into `world` contiguous blocks of WHOLE
    keyframes with near-equal row counts.  Returns [(kf_begin, kf_end, row_begin, row_end)] per rank.
    A keyframe is never split, so the per-keyframe (LoopCloser) epilogue stays rank-local."""
    seg_off = [int(x) for x in seg_off]
    nkf, total = len(seg_off) - 1, seg_off[-1]
    cuts = [0]
    for r in range(1, world):
        target = total * r // world
        k = cuts[-1]
        while k < nkf and seg_off[k] < target:          # first keyframe starting at/after the target
            k += 1
        if k > cuts[-1] and k <= nkf and (seg_off[k] - target) > (target - seg_off[k - 1]):
            k -= 1                                       # the previous boundary is closer
        cuts.append(max(k, cuts[-1]))
    cuts.append(nkf)
    return [(cuts[r], cuts[r + 1], seg_off[cuts[r]], seg_off[cuts[r + 1]]) for r in range(world)]


def gather_and_merge(l_idx, l_dist, world, group, merge, g_idx=None, g_dist=None):
    """The exchange step of the sharded search: all-gather every rank's [nq,2] (global index,
    distance) lists and merge them by (distance, index).  `merge(g_idx, g_dist)` -> (idx, dist).
    Works on CUDA tensors (NCCL) and on CPU tensors (gloo; used by the CPU tests)."""
    if world == 1:
        return l_idx, l_dist
    if g_idx is None:
        g_idx = torch.empty((world,) + tuple(l_idx.shape), dtype=l_idx.dtype, device=l_idx.device)
        g_dist = torch.empty((world,) + tuple(l_dist.shape), dtype=l_dist.dtype, device=l_dist.device)
    # concatenated-along-dim-0 form (accepted by both NCCL and gloo): [world*nq, 2] views
    torch.distributed.all_gather_into_tensor(g_idx.view(-1, *l_idx.shape[1:]), l_idx, group=group)
    torch.distributed.all_gather_into_tensor(g_dist.view(-1, *l_dist.shape[1:]), l_dist, group=group)
    return merge(g_idx, g_dist)


def loop_checked_before(cur_frame_id, frame_ids, counts, min_gap, kf_begin):
    """The `checked` counter of LoopCloser::detect (src/LoopCloser.cpp:43-48) on entry to the shard
    that starts at keyframe kf_begin: how many EARLIER keyframes pass the gap (:44) and non-empty
    (:45) tests.  frame_ids / counts describe the whole keyframe list (every rank holds this
    small table), so no rank waits for another."""
    n = 0
    for s in range(int(kf_begin)):
        if cur_frame_id - int(frame_ids[s]) < min_gap:
            continue
        if int(counts[s]) == 0:
            continue
        n += 1
    return n


def concat_keyframe_status(local_status, nkf_per_rank, world, group=None):
    """Host concatenation of the per-rank status arrays (the per-keyframe search has
    no data-path collective).  local_status: int32 tensor [nkf of this rank] (CPU with gloo, CUDA
    with NCCL); returns the whole list's status on every rank, as a CPU int32 tensor."""
    if world == 1:
        return local_status.cpu()
    cap = max(int(n) for n in nkf_per_rank)
    pad = torch.full((cap,), -1, dtype=torch.int32, device=local_status.device)
    pad[:local_status.shape[0]] = local_status
    out = torch.empty((world * cap,), dtype=torch.int32, device=local_status.device)
    torch.distributed.all_gather_into_tensor(out, pad, group=group)
    out = out.cpu().view(world, cap)
    return torch.cat([out[r, :int(nkf_per_rank[r])] for r in range(world)])


class ShardedDB:
    def __init__(self, device, rank=0, world=1, group=None, engine=ENGINE_TENSOR, exchange="p2p", nq_cap=4096):
        """exchange: "p2p"  = result keys stored straight into the peers' buffers over NVLink and merged
                              in the same kernel (libvsm's fused exchange; CUDA IPC handles are swapped
                              once through the process group);
                     "nccl" = all_gather_into_tensor of the keys + merge kernel."""
        self.rank, self.world, self.group = rank, world, group
        self.device = torch.device("cuda", device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.matcher = Matcher(device=device, engine=engine)
        self.matcher.set_stream(self.stream.cuda_stream)
        self.exchange = exchange if world > 1 else "none"
        self.exchange_note = None
        if self.exchange == "p2p":
            # CUDA IPC between the ranks' processes; if ANY rank cannot set it up (e.g. a container
            # without shared IPC namespaces) every rank switches to the NCCL exchange together
            ok, err = 1, ""
            try:
                mine = self.matcher.xchg_create(rank, world, nq_cap)
            except Exception as e:                          # noqa: BLE001
                ok, err, mine = 0, repr(e), b""
            handles = [None] * world
            torch.distributed.all_gather_object(handles, mine, group=group)
            if ok and all(len(h) == 64 for h in handles):
                try:
                    self.matcher.xchg_connect(handles)
                except Exception as e:                      # noqa: BLE001
                    ok, err = 0, repr(e)
            else:
                ok = 0
            flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
            torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN, group=group)
            if int(flag.item()) == 0:
                self.exchange = "nccl"
                self.exchange_note = f"peer-memory exchange unavailable ({err or 'another rank failed'}); using NCCL all-gather"
            torch.distributed.barrier(group=group)          # every buffer is zeroed and mapped before first use
        self.row_offset = 0
        self.rows = 0
        self._db = None
        self._nq = -1

    def adopt(self, db, row_offset, seg_off=None, frame_ids=None):
        """db: this rank's [rows, 256] fp32 CUDA tensor (kept alive here); row_offset: global
        index of its first row; seg_off: local row offsets of its keyframes; frame_ids: their
        Frame::id()s (for the gap rule of loop_detect)."""
        assert db.is_cuda and db.dtype == torch.float32 and db.is_contiguous() and db.shape[1] == 256
        torch.cuda.synchronize(self.device)
        self._db = db
        self.rows = db.shape[0]
        self.row_offset = int(row_offset)
        if self.rows == 0:
            # more ranks than keyframes: this rank holds nothing and answers "no neighbour" to every
            # query; it still takes part in the exchange
            self.matcher.clear_store()
            return
        self.matcher.adopt_device_matrix(db.data_ptr(), self.rows, seg_off)
        if frame_ids is not None:
            self.matcher.set_frame_ids(frame_ids)

    def set_keyframe_table(self, frame_ids, counts, kf_cuts):
        """The whole keyframe list's frame ids and descriptor counts plus the partition
        (kf_cuts[r] .. kf_cuts[r+1] = rank r's keyframes, e.g. from partition_keyframes)."""
        self._kf_ids = [int(x) for x in frame_ids]
        self._kf_counts = [int(x) for x in counts]
        self._kf_cuts = [int(x) for x in kf_cuts]
        assert len(self._kf_cuts) == self.world + 1 and self._kf_cuts[-1] == len(self._kf_ids)

    def loop_detect(self, cur_frame_id, h_q, ratio=0.75, min_gap=200, every=5, want_matches=True):
        """LoopCloser::detect's candidate loop (src/LoopCloser.cpp:43-62) over the partitioned
        keyframe list: each rank matches its own eligible keyframes, the status arrays are
        concatenated.  Returns (status of the WHOLE list [CPU int32], this rank's match lists
        keyed by global keyframe index)."""
        k0 = self._kf_cuts[self.rank]
        before = loop_checked_before(cur_frame_id, self._kf_ids, self._kf_counts, min_gap, k0)
        st, lists, _ = self.matcher.loop_detect_shard(cur_frame_id, h_q, before, ratio, min_gap, every, want_matches)
        nkf = [self._kf_cuts[r + 1] - self._kf_cuts[r] for r in range(self.world)]
        whole = concat_keyframe_status(torch.from_numpy(st).to(self.device if self.world > 1 else "cpu"), nkf,
                                       self.world, self.group)
        mine = {k0 + s: lists[s] for s in range(len(st)) if lists is not None and lists[s] is not None}
        return whole, mine

    def loop_detect_compact(self, cur_frame_id, h_q, ratio=0.75, min_gap=200, every=5, min_matches=30):
        """The same loop in compact form (vsm_loop_detect_compact): gate and packing on each rank's device.
        Returns (status of the WHOLE list [CPU int32], {global keyframe index: match list} of this rank's
        keyframes that pass the gate)."""
        k0 = self._kf_cuts[self.rank]
        before = loop_checked_before(cur_frame_id, self._kf_ids, self._kf_counts, min_gap, k0)
        if self.rows == 0:
            st, lists = [], {}
        else:
            st, lists, _ = self.matcher.loop_detect_compact(cur_frame_id, h_q, ratio, min_gap, every, min_matches,
                                                            checked_before=before)
        nkf = [self._kf_cuts[r + 1] - self._kf_cuts[r] for r in range(self.world)]
        import numpy as np
        st_t = torch.from_numpy(np.asarray(st, np.int32)).to(self.device if self.world > 1 else "cpu")
        whole = concat_keyframe_status(st_t, nkf, self.world, self.group)
        out = {}
        for s, lst in lists.items():
            lst = lst.copy()
            lst["imgIdx"] = k0 + s
            out[k0 + s] = lst
        return whole, out

    def _buffers(self, nq):
        """Per-batch-size device and pinned buffers.  They are only ever used on self.stream through raw
        pointers, so they are allocated under that stream (the caching allocator then orders any reuse
        of their memory after the kernels enqueued there) and kept per nq instead of being freed."""
        if nq != self._nq:
            if not hasattr(self, "_bufs"):
                self._bufs = {}
            if nq not in self._bufs:
                dev = self.device
                b = {}
                with torch.cuda.stream(self.stream):
                    b["l_idx"] = torch.empty((nq, 2), dtype=torch.int64, device=dev)
                    b["l_dist"] = torch.empty((nq, 2), dtype=torch.float32, device=dev)
                    if self.world > 1:
                        b["l_keys"] = torch.empty((nq, 2), dtype=torch.int64, device=dev)         # u64 bit patterns
                        b["g_keys"] = torch.empty((self.world * nq, 2), dtype=torch.int64, device=dev)
                        b["o_idx"] = torch.empty((nq, 2), dtype=torch.int64, device=dev)
                        b["o_dist"] = torch.empty((nq, 2), dtype=torch.float32, device=dev)
                    else:
                        b["o_idx"], b["o_dist"] = b["l_idx"], b["l_dist"]
                    b["d_q"] = torch.empty((nq, 256), dtype=torch.float32, device=dev)
                b["h_idx"] = torch.empty((nq, 2), dtype=torch.int64).pin_memory()
                b["h_dist"] = torch.empty((nq, 2), dtype=torch.float32).pin_memory()
                self._bufs[nq] = b
            for k, v in self._bufs[nq].items():
                setattr(self, k, v)
            self._nq = nq

    def search_device(self, d_q):
        """d_q: [nq,256] fp32 on this device.  Enqueues on self.stream; returns (idx, dist) device
        tensors holding the GLOBAL top-2 (valid once the stream has been synchronised)."""
        nq = d_q.shape[0]
        self._buffers(nq)
        with torch.cuda.stream(self.stream):
            if self.world == 1:
                self.matcher.db_top2_device(d_q.data_ptr(), nq, self.row_offset, self.l_idx.data_ptr(),
                                            self.l_dist.data_ptr(), sync=False)
            elif self.exchange == "p2p":
                self.matcher.db_top2_xchg_device(d_q.data_ptr(), nq, self.row_offset, self.o_idx.data_ptr(),
                                                 self.o_dist.data_ptr(), sync=False)
            else:
                # one 16 B x nq buffer per rank: keys ~((distance bits << 32) | global index)
                self.matcher.db_top2_keys_device(d_q.data_ptr(), nq, self.row_offset, self.l_keys.data_ptr(), sync=False)
                torch.distributed.all_gather_into_tensor(self.g_keys, self.l_keys, group=self.group)
                self.matcher.merge_keys_device(self.g_keys.data_ptr(), self.world, nq, self.o_idx.data_ptr(),
                                               self.o_dist.data_ptr(), sync=False)
        return self.o_idx, self.o_dist

    def search_host(self, h_q):
        """h_q: pinned [nq,256] fp32 host tensor -> pinned host (idx, dist); asynchronous on
        self.stream (synchronise it before reading)."""
        nq = h_q.shape[0]
        self._buffers(nq)
        with torch.cuda.stream(self.stream):
            self.d_q.copy_(h_q, non_blocking=True)
            oi, od = self.search_device(self.d_q)
            self.h_idx.copy_(oi, non_blocking=True)
            self.h_dist.copy_(od, non_blocking=True)
        return self.h_idx, self.h_dist

    def search_host_abi(self, q):
        """The reference-facing call: q is a host [nq,256] fp32 numpy array (pinned or pageable); returns
        host (idx, dist) numpy arrays holding the GLOBAL top-2.  One synchronous C-ABI call per rank
        (vsm_db_top2 on one GPU; vsm_db_top2_xchg, the collective form, with the fused exchange)."""
        if self.world == 1:
            return self.matcher.search_map_points(q, self.row_offset)
        if self.exchange == "p2p":
            return self.matcher.db_top2_xchg(q, self.row_offset)
        hi, hd = self.search_host(torch.from_numpy(q))
        self.stream.synchronize()
        return hi.numpy(), hd.numpy()

    def launches_per_search(self):
        return self.matcher.stats()["kernel_launches"] + (1 if self.exchange == "nccl" else 0)

    def close(self):
        self.matcher.close()
