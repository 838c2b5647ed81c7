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


class ShardedDB:
    def __init__(self, device, rank=0, world=1, group=None, engine=ENGINE_TENSOR):
        self.rank, self.world, self.group = rank, world, group
        self.device = torch.device("cuda", device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.matcher = Matcher(device=device, engine=engine)
        self.matcher.set_stream(self.stream.cuda_stream)
        self.row_offset = 0
        self.rows = 0
        self._db = None
        self._nq = -1

    def adopt(self, db, row_offset, seg_off=None):
        """db: this rank's [rows, 256] fp32 CUDA tensor (kept alive here); row_offset: global
        index of its first row."""
        assert db.is_cuda and db.dtype == torch.float32 and db.is_contiguous() and db.shape[1] == 256
        torch.cuda.synchronize(self.device)
        self._db = db
        self.rows = db.shape[0]
        self.row_offset = int(row_offset)
        self.matcher.adopt_device_matrix(db.data_ptr(), self.rows, seg_off)

    def _buffers(self, nq):
        if nq != self._nq:
            dev = self.device
            self.l_idx = torch.empty((nq, 2), dtype=torch.int64, device=dev)
            self.l_dist = torch.empty((nq, 2), dtype=torch.float32, device=dev)
            if self.world > 1:
                self.g_idx = torch.empty((self.world, nq, 2), dtype=torch.int64, device=dev)
                self.g_dist = torch.empty((self.world, nq, 2), dtype=torch.float32, device=dev)
                self.o_idx = torch.empty((nq, 2), dtype=torch.int64, device=dev)
                self.o_dist = torch.empty((nq, 2), dtype=torch.float32, device=dev)
            else:
                self.o_idx, self.o_dist = self.l_idx, self.l_dist
            self.h_idx = torch.empty((nq, 2), dtype=torch.int64).pin_memory()
            self.h_dist = torch.empty((nq, 2), dtype=torch.float32).pin_memory()
            self.d_q = torch.empty((nq, 256), dtype=torch.float32, device=dev)
            self._nq = nq

    def search_device(self, d_q):
        """d_q: [nq,256] fp32 on this device.  Enqueues on self.stream; returns (idx, dist) device
        tensors holding the GLOBAL top-2 (valid once the stream has been synchronised)."""
        nq = d_q.shape[0]
        self._buffers(nq)
        with torch.cuda.stream(self.stream):
            self.matcher.db_top2_device(d_q.data_ptr(), nq, self.row_offset, self.l_idx.data_ptr(),
                                        self.l_dist.data_ptr(), sync=False)
            if self.world > 1:
                torch.distributed.all_gather_into_tensor(self.g_idx, self.l_idx, group=self.group)
                torch.distributed.all_gather_into_tensor(self.g_dist, self.l_dist, group=self.group)
                self.matcher.merge_top2_device(self.g_idx.data_ptr(), self.g_dist.data_ptr(), self.world, nq,
                                               self.o_idx.data_ptr(), self.o_dist.data_ptr(), sync=False)
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

    def launches_per_search(self):
        return self.matcher.stats()["kernel_launches"] + (1 if self.world > 1 else 0)

    def close(self):
        self.matcher.close()
