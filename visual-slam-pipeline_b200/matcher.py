"""ctypes binding of libvsm.so and a thin Python mirror of the reference's matcher interface.

Names follow the reference (salah-dev-stu/visual-slam-pipeline):
  Matcher.match_features(desc1, desc2, ratio)   <- Slam::match_features      src/Slam.cpp:1140-1172
  Matcher.knn_match(query, train)               <- DescriptorMatcher::knnMatch src/Slam.cpp:1149
  Matcher.add_keyframe / match_to_keyframe      <- Frame::descriptors_ kept on the device (include/Frame.h:61)
  Matcher.search_map_points(frame_desc)         <- stacked-matrix search       src/Slam.cpp:546-574, 744-774
  Matcher.detect_candidates(frame_desc)         <- LoopCloser::detect block    src/LoopCloser.cpp:43-62

This module never computes a match on the CPU: if libvsm.so is missing or there is no
B200, construction raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
DMATCH = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")])
CAND = np.dtype([("keyframe", "<i4"), ("count", "<i4"), ("offset", "<i8")])       # vsm_loop_candidate
ENGINE_AUTO, ENGINE_TENSOR, ENGINE_SIMT, ENGINE_TENSOR_PAIR = 0, 1, 2, 3
DIM = 256


class VsmError(RuntimeError):
    pass


class _Opts(C.Structure):
    _fields_ = [("device", C.c_int32), ("engine", C.c_int32), ("scratch_rows", C.c_int64),
                ("store_rows", C.c_int64), ("reserved", C.c_int32 * 8)]


class TrackCfg(C.Structure):
    """vsm_track_cfg; defaults = the reference's Config.h constants."""
    _fields_ = [("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("width", C.c_int32), ("height", C.c_int32), ("cell_size", C.c_int32), ("reserved", C.c_int32),
                ("depth_min", C.c_double), ("depth_max", C.c_double), ("search_radius", C.c_double),
                ("desc_threshold", C.c_double)]

    def __init__(self, **kw):
        d = dict(fx=525.0, fy=525.0, cx=319.5, cy=239.5, width=640, height=480, cell_size=30, reserved=0,
                 depth_min=float(np.float32(0.1)), depth_max=50.0, search_radius=12.0, desc_threshold=0.5)
        d.update(kw)
        super().__init__(**d)


class _Stats(C.Structure):
    _fields_ = [("candidates", C.c_int64), ("flagged_slices", C.c_int64), ("kernel_launches", C.c_int64),
                ("device_ms", C.c_float), ("tc_ms", C.c_float), ("select_ms", C.c_float),
                ("slice_tiles", C.c_int32), ("reserved", C.c_int32)]


def lib_path():
    return os.path.join(_HERE, "lib", "libvsm.so")


_lib = None
SYMBOLS = {
    "vsm_default_opts": (None, [C.POINTER(_Opts)]),
    "vsm_create": (C.c_int, [C.POINTER(_Opts), C.POINTER(C.c_void_p)]),
    "vsm_destroy": (None, [C.c_void_p]),
    "vsm_last_error": (C.c_char_p, [C.c_void_p]),
    "vsm_version": (C.c_char_p, []),
    "vsm_get_stats": (C.c_int, [C.c_void_p, C.POINTER(_Stats)]),
    "vsm_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_int64]),
    "vsm_host_free": (None, [C.c_void_p]),
    "vsm_knn2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "vsm_knn2_strided": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p,
                                   C.c_void_p]),
    "vsm_match_pair_strided": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_int32, C.c_int64,
                                         C.c_float, C.c_int32, C.c_void_p, C.POINTER(C.c_int32), C.c_void_p,
                                         C.POINTER(C.c_int32)]),
    "vsm_store_add_strided": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.POINTER(C.c_int32)]),
    "vsm_match_pair": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_float, C.c_int32,
                                 C.c_void_p, C.POINTER(C.c_int32), C.c_void_p, C.POINTER(C.c_int32)]),
    "vsm_match_batch": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_float, C.c_int32, C.c_void_p, C.c_void_p]),
    "vsm_store_add": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]),
    "vsm_store_add_device": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.POINTER(C.c_int32)]),
    "vsm_store_adopt_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32]),
    "vsm_store_load_spcf": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                      C.POINTER(C.c_int32)]),
    "vsm_store_clear": (C.c_int, [C.c_void_p]),
    "vsm_store_promote": (C.c_int, [C.c_void_p, C.c_int32]),
    "vsm_store_remove": (C.c_int, [C.c_void_p, C.c_int32]),
    "vsm_store_frame_info": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int32),
                                       C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "vsm_store_keyframes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]),
    "vsm_store_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    "vsm_match_to_stored": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_float, C.c_int32,
                                      C.c_void_p, C.POINTER(C.c_int32), C.c_void_p, C.POINTER(C.c_int32)]),
    "vsm_track": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_float, C.c_int32,
                            C.c_void_p, C.POINTER(C.c_int32), C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "vsm_db_top2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]),
    "vsm_track_local_map": (C.c_int, [C.c_void_p, C.POINTER(TrackCfg), C.c_void_p, C.c_void_p, C.c_int32,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_void_p, C.c_void_p]),
    "vsm_db_top2_masked": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "vsm_points_add": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_int32)]),
    "vsm_points_add_from_frame": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]),
    "vsm_points_observe": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]),
    "vsm_points_set_valid": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]),
    "vsm_points_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "vsm_points_clear": (C.c_int, [C.c_void_p]),
    "vsm_points_top2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                  C.POINTER(C.c_int32)]),
    "vsm_db_segmented": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_float, C.c_void_p, C.c_void_p]),
    "vsm_loop_detect": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_float,
                                  C.c_void_p, C.c_void_p]),
    "vsm_loop_detect_shard": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                        C.c_float, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]),
    "vsm_loop_detect_compact": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                          C.c_float, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int32),
                                          C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    "vsm_store_set_frame_ids": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32]),
    "vsm_match_batch_stored": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_float, C.c_int32, C.c_void_p,
                                         C.c_int64, C.c_void_p, C.c_void_p]),
    "vsm_tc_history": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]),
    "vsm_db_top2_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32]),
    "vsm_db_top2_keys_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_int32]),
    "vsm_merge_keys_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32]),
    "vsm_xchg_create": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "vsm_xchg_connect": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vsm_db_top2_xchg_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32]),
    "vsm_db_top2_xchg": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]),
    "vsm_merge_top2_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                        C.c_void_p, C.c_int32]),
    "vsm_group_create": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(_Opts), C.POINTER(C.c_void_p)]),
    "vsm_group_destroy": (None, [C.c_void_p]),
    "vsm_group_last_error": (C.c_char_p, [C.c_void_p]),
    "vsm_group_size": (C.c_int, [C.c_void_p]),
    "vsm_group_ctx": (C.c_void_p, [C.c_void_p, C.c_int32]),
    "vsm_group_store_add": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]),
    "vsm_group_store_remove": (C.c_int, [C.c_void_p, C.c_int32]),
    "vsm_group_store_clear": (C.c_int, [C.c_void_p]),
    "vsm_group_store_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.c_void_p]),
    "vsm_group_adopt_device": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32]),
    "vsm_group_db_top2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vsm_group_loop_detect": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_float,
                                        C.c_void_p, C.c_void_p]),
    "vsm_group_loop_detect_compact": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_float,
                                                C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.c_void_p,
                                                C.c_int64, C.POINTER(C.c_int64)]),
    "vsm_group_match_batch": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float,
                                        C.c_int32, C.c_void_p, C.c_void_p]),
    "vsm_stream": (C.c_void_p, [C.c_void_p]),
    "vsm_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vsm_sync": (C.c_int, [C.c_void_p]),
    "vsm_set_profiling": (C.c_int, [C.c_void_p, C.c_int32]),
    "vsm_synth_rows_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_uint64]),
    "vsm_debug_fetch_dump": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "vsm_debug_tile_scores": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]),
}


def load_library():
    """dlopen libvsm.so and type every exported entry point (no CUDA call is made)."""
    global _lib
    if _lib is None:
        path = lib_path()
        if not os.path.exists(path):
            # not built yet (fresh checkout): compile it in-tree with nvcc; there is no other way to run
            try:
                from . import build as _build
                _build.build_lib()
            except Exception as e:
                raise VsmError(f"{path} is missing and could not be built ({e}): run "
                               "`python __graft_entry__.py build` (there is no CPU fallback)")
        lib = C.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _rows(a, name):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim != 2 or a.shape[1] != DIM:
        raise ValueError(f"{name}: expected an N x {DIM} float32 matrix, got {a.shape}")
    return a


class Matcher:
    """One matching context on one GPU (single caller, synchronous calls)."""

    @classmethod
    def _borrowed(cls, handle):
        """A view of a context owned by someone else (a group member): never destroyed here."""
        m = cls.__new__(cls)
        m._lib = load_library()
        m._h = C.c_void_p(handle)
        m._owned = False
        return m

    def __init__(self, device=0, engine=ENGINE_AUTO, scratch_rows=0, store_rows=0, seg_tiles=0, work_cap=0, ring=0,
                 pair_cap=0, append_tiles=0, tile_top2=True):
        self._lib = load_library()
        o = _Opts()
        self._lib.vsm_default_opts(C.byref(o))
        o.device, o.engine, o.scratch_rows, o.store_rows = device, engine, scratch_rows, store_rows
        o.reserved[0] = seg_tiles
        o.reserved[1] = work_cap
        o.reserved[2] = ring
        o.reserved[3] = pair_cap
        o.reserved[4] = append_tiles
        o.reserved[5] = 0 if tile_top2 else 1      # pair matching: tile top-2 records (default) or the top-4 records
        h = C.c_void_p()
        st = self._lib.vsm_create(C.byref(o), C.byref(h))
        if st != 0:
            raise VsmError(f"vsm_create failed ({st}): {self._lib.vsm_last_error(None).decode()}")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            if getattr(self, "_owned", True):
                self._lib.vsm_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, st):
        if st != 0:
            raise VsmError(f"libvsm error {st}: {self._lib.vsm_last_error(self._h).decode()}")

    @property
    def handle(self):
        return self._h

    @property
    def lib(self):
        return self._lib

    def stats(self):
        s = _Stats()
        self._ck(self._lib.vsm_get_stats(self._h, C.byref(s)))
        return {"candidates": s.candidates, "flagged_slices": s.flagged_slices,
                "kernel_launches": s.kernel_launches, "device_ms": s.device_ms,
                "tc_ms": s.tc_ms, "select_ms": s.select_ms, "slice_tiles": s.slice_tiles}

    # -- cv::DescriptorMatcher::knnMatch(query, train, knn, 2) ---------------------------
    def knn_match(self, query, train):
        """(idx[nq,2] int32, dist[nq,2] fp32); a missing neighbour is idx -1 / dist FLT_MAX."""
        q, t = _rows(query, "query"), _rows(train, "train")
        idx = np.empty((q.shape[0], 2), np.int32)
        dist = np.empty((q.shape[0], 2), np.float32)
        self._ck(self._lib.vsm_knn2(self._h, q.ctypes.data, q.shape[0], t.ctypes.data, t.shape[0],
                                    idx.ctypes.data, dist.ctypes.data))
        return idx, dist

    # -- Slam::match_features(desc1, desc2, raw_out) ----------------------------------------
    def match_features(self, desc1, desc2, ratio=0.75, mutual=False, want_raw=True):
        """Returns (good, raw) DMATCH arrays in query order; raw is None if not asked for."""
        q, t = _rows(desc1, "desc1"), _rows(desc2, "desc2")
        nq = q.shape[0]
        good = np.zeros(max(nq, 1), DMATCH)
        raw = np.zeros(max(nq, 1), DMATCH) if want_raw else None
        ng, nr = C.c_int32(0), C.c_int32(0)
        self._ck(self._lib.vsm_match_pair(self._h, q.ctypes.data, nq, t.ctypes.data, t.shape[0], ratio, int(mutual),
                                          good.ctypes.data, C.byref(ng),
                                          raw.ctypes.data if want_raw else None, C.byref(nr) if want_raw else None))
        return good[:ng.value], (raw[:nr.value] if want_raw else None)

    def match_batch(self, queries, trains, ratio=0.75, mutual=False):
        """Ragged batch: lists of N_i x 256 matrices -> list of DMATCH arrays (one launch sequence)."""
        n = len(queries)
        assert n == len(trains)
        q_off = np.zeros(n + 1, np.int32)
        t_off = np.zeros(n + 1, np.int32)
        q_off[1:] = np.cumsum([len(a) for a in queries])
        t_off[1:] = np.cumsum([len(a) for a in trains])
        qa = _rows(np.concatenate(queries) if n else np.zeros((0, DIM), np.float32), "queries")
        ta = _rows(np.concatenate(trains) if n else np.zeros((0, DIM), np.float32), "trains")
        return self.match_batch_packed(qa, q_off, ta, t_off, ratio, mutual)

    def match_batch_packed(self, qa, q_off, ta, t_off, ratio=0.75, mutual=False, out=None):
        n = len(q_off) - 1
        good = out if out is not None else np.zeros(max(int(q_off[-1]), 1), DMATCH)
        n_good = np.zeros(max(n, 1), np.int32)
        self._ck(self._lib.vsm_match_batch(self._h, n, qa.ctypes.data, q_off.ctypes.data, ta.ctypes.data,
                                           t_off.ctypes.data, ratio, int(mutual), good.ctypes.data,
                                           n_good.ctypes.data))
        return [good[q_off[p]:q_off[p] + n_good[p]] for p in range(n)]

    def match_batch_stored(self, q_handles, t_handles, ratio=0.75, mutual=False, capacity=None):
        """match_features for pairs of STORED keyframes (no upload).  Returns a list of DMATCH arrays.
        capacity: the sum of the query keyframes' rows if the caller knows it (saves the size query)."""
        qh = np.ascontiguousarray(q_handles, np.int32)
        th = np.ascontiguousarray(t_handles, np.int32)
        n = len(qh)
        assert len(th) == n
        n_good = np.zeros(max(n, 1), np.int32)
        off = np.zeros(n + 1, np.int64)
        if capacity is None:
            self._ck(self._lib.vsm_match_batch_stored(self._h, n, qh.ctypes.data, th.ctypes.data, ratio, int(mutual),
                                                      None, 0, n_good.ctypes.data, off.ctypes.data))
            capacity = int(off[n])
        cap = max(int(capacity), 1)
        good = np.empty(cap, DMATCH)        # a fresh buffer per call: the lists below are views into it (64 structured
                                            # copies cost 0.5 ms of Python time on a call that takes 0.2 ms on the device)
        self._ck(self._lib.vsm_match_batch_stored(self._h, n, qh.ctypes.data, th.ctypes.data, ratio, int(mutual),
                                                  good.ctypes.data, cap, n_good.ctypes.data, off.ctypes.data))
        return [good[off[p]:off[p] + n_good[p]] for p in range(n)]

    # -- device-resident keyframe store -----------------------------------------------------
    def add_keyframe(self, frame_id, desc):
        d = _rows(desc, "desc")
        h = C.c_int32(-1)
        self._ck(self._lib.vsm_store_add(self._h, frame_id, d.ctypes.data, d.shape[0], C.byref(h)))
        return h.value

    def add_keyframe_device(self, frame_id, dev_ptr, n):
        h = C.c_int32(-1)
        self._ck(self._lib.vsm_store_add_device(self._h, frame_id, C.c_void_p(dev_ptr), n, C.byref(h)))
        return h.value

    def adopt_device_matrix(self, dev_ptr, n_rows, seg_off=None):
        if seg_off is not None:
            seg_off = np.ascontiguousarray(seg_off, np.int64)
            self._ck(self._lib.vsm_store_adopt_device(self._h, C.c_void_p(dev_ptr), n_rows, seg_off.ctypes.data,
                                                      len(seg_off) - 1))
        else:
            self._ck(self._lib.vsm_store_adopt_device(self._h, C.c_void_p(dev_ptr), n_rows, None, 0))

    def load_feature_cache(self, path):
        """Bulk-load the reference's SPCF feature cache (src/FeatureExtractor.cpp:269-381) into the store.
        Returns (entries loaded as keyframes, entries skipped, handle of the first keyframe)."""
        a, b, h = C.c_int32(0), C.c_int32(0), C.c_int32(-1)
        self._ck(self._lib.vsm_store_load_spcf(self._h, os.fsencode(path), C.byref(a), C.byref(b), C.byref(h)))
        return a.value, b.value, h.value

    def clear_store(self):
        self._ck(self._lib.vsm_store_clear(self._h))

    def promote(self, handle):
        """Frame::set_keyframe(true) for a frame stored by track()."""
        self._ck(self._lib.vsm_store_promote(self._h, handle))

    def remove_frame(self, handle):
        self._ck(self._lib.vsm_store_remove(self._h, handle))

    def frame_info(self, handle):
        """(rows, frame id, is_keyframe, first store row) of a live handle."""
        r0, n, f, k = C.c_int64(0), C.c_int32(0), C.c_int32(0), C.c_int32(0)
        self._ck(self._lib.vsm_store_frame_info(self._h, handle, C.byref(r0), C.byref(n), C.byref(f), C.byref(k)))
        return n.value, f.value, bool(k.value), r0.value

    def keyframes(self):
        """Handles of the keyframes in Map::get_keyframes() order."""
        n = C.c_int32(0)
        self._ck(self._lib.vsm_store_keyframes(self._h, None, 0, C.byref(n)))
        h = np.zeros(max(n.value, 1), np.int32)
        self._ck(self._lib.vsm_store_keyframes(self._h, h.ctypes.data, n.value, C.byref(n)))
        return h[:n.value]

    def store_info(self):
        r, k = C.c_int64(0), C.c_int32(0)
        self._ck(self._lib.vsm_store_info(self._h, C.byref(r), C.byref(k)))
        return r.value, k.value

    def match_to_keyframe(self, handle, cur_desc, ratio=0.75, mutual=False, want_raw=False):
        """Slam::match_features(ref_kf->descriptors(), cur->descriptors()) with the keyframe resident."""
        t = _rows(cur_desc, "cur_desc")
        cap = max(self.frame_info(handle)[0], 1)
        good = np.zeros(cap, DMATCH)
        raw = np.zeros(cap, DMATCH) if want_raw else None
        ng, nr = C.c_int32(0), C.c_int32(0)
        self._ck(self._lib.vsm_match_to_stored(self._h, handle, t.ctypes.data, t.shape[0], ratio, int(mutual),
                                               good.ctypes.data, C.byref(ng),
                                               raw.ctypes.data if want_raw else None,
                                               C.byref(nr) if want_raw else None))
        return good[:ng.value].copy(), (raw[:nr.value].copy() if want_raw else None)

    def track(self, ref_handle, frame_id, cur_desc, ratio=0.75, mutual=False, want_raw=False, ref_rows=None):
        """Slam::process_frame's tracking match (src/Slam.cpp:838-842) for a sequence: the current
        frame is uploaded once, into the store; returns (good, raw, handle of the current frame)."""
        t = _rows(cur_desc, "cur_desc")
        if ref_rows is None:
            ref_rows = self.frame_info(ref_handle)[0] if ref_handle >= 0 else 0
        cap = max(ref_rows, 1)
        if not hasattr(self, "_trk") or len(self._trk[0]) < cap:
            self._trk = (np.zeros(cap, DMATCH), np.zeros(cap, DMATCH))
        good, raw = self._trk
        ng, nr, h = C.c_int32(0), C.c_int32(0), C.c_int32(-1)
        self._ck(self._lib.vsm_track(self._h, ref_handle, frame_id, t.ctypes.data, t.shape[0], ratio, int(mutual),
                                     good.ctypes.data, C.byref(ng), raw.ctypes.data if want_raw else None,
                                     C.byref(nr) if want_raw else None, C.byref(h)))
        return good[:ng.value].copy(), (raw[:nr.value].copy() if want_raw else None), h.value

    def search_map_points(self, frame_desc, row_offset=0):
        """Global top-2 of every query row over the whole store (src/Slam.cpp:546-574)."""
        q = _rows(frame_desc, "frame_desc")
        idx = np.empty((q.shape[0], 2), np.int64)
        dist = np.empty((q.shape[0], 2), np.float32)
        self._ck(self._lib.vsm_db_top2(self._h, q.ctypes.data, q.shape[0], row_offset, idx.ctypes.data,
                                       dist.ctypes.data))
        return idx, dist

    def search_map_points_masked(self, frame_desc, mask):
        """Global top-2 over the store rows with mask != 0 (valid / nearby map points only,
        src/Slam.cpp:553, :748-756).  idx = original store rows."""
        q = _rows(frame_desc, "frame_desc")
        mask = np.ascontiguousarray(mask, np.uint8)
        idx = np.empty((q.shape[0], 2), np.int64)
        dist = np.empty((q.shape[0], 2), np.float32)
        self._ck(self._lib.vsm_db_top2_masked(self._h, q.ctypes.data, q.shape[0], mask.ctypes.data, mask.shape[0],
                                              idx.ctypes.data, dist.ctypes.data))
        return idx, dist

    # -- resident map-point table (Map::map_points_) -------------------------------------------------
    def add_map_points(self, desc, frame_id):
        """MapPoint births from host descriptors (src/Slam.cpp:1337-1347); returns the first new point id."""
        d = _rows(desc, "desc")
        first = C.c_int32(-1)
        self._ck(self._lib.vsm_points_add(self._h, d.ctypes.data, d.shape[0], frame_id, C.byref(first)))
        return first.value

    def add_map_points_from_frame(self, handle, kp_idx):
        """MapPoint births whose descriptor is a row of a stored frame (row(i).clone(), src/Slam.cpp:1339, :1563)."""
        k = np.ascontiguousarray(kp_idx, np.int32)
        first = C.c_int32(-1)
        self._ck(self._lib.vsm_points_add_from_frame(self._h, handle, k.ctypes.data, len(k), C.byref(first)))
        return first.value

    def observe_map_points(self, point_ids, frame_id):
        p = np.ascontiguousarray(point_ids, np.int32)
        self._ck(self._lib.vsm_points_observe(self._h, p.ctypes.data, len(p), frame_id))

    def set_map_points_valid(self, point_ids, valid):
        p = np.ascontiguousarray(point_ids, np.int32)
        self._ck(self._lib.vsm_points_set_valid(self._h, p.ctypes.data, len(p), int(valid)))

    def map_point_info(self):
        a, b, c = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        self._ck(self._lib.vsm_points_info(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def clear_map_points(self):
        self._ck(self._lib.vsm_points_clear(self._h))

    def search_resident_map_points(self, frame_desc, near_frame_id=-1, frame_range=30):
        """knnMatch(frame, stacked descriptors of the selected map points, 2) with the selection made on the
        device: valid points (src/Slam.cpp:552-557), optionally only those observed within frame_range frames
        of near_frame_id (:744-759).  Returns (point ids [nq,2], dist [nq,2], rows of the stacked matrix)."""
        q = _rows(frame_desc, "frame_desc")
        idx = np.empty((q.shape[0], 2), np.int64)
        dist = np.empty((q.shape[0], 2), np.float32)
        ns = C.c_int32(0)
        self._ck(self._lib.vsm_points_top2(self._h, q.ctypes.data, q.shape[0], near_frame_id, frame_range, idx.ctypes.data,
                                           dist.ctypes.data, C.byref(ns)))
        return idx, dist, ns.value

    def track_local_map(self, kp_xy, desc, mp_pos, mp_desc, mp_valid, R_cam, t_cam, indices, cfg=None):
        """Slam::track_local_map (src/Slam.cpp:380-469).  indices is updated in place.
        Returns (tracked, observations [(mp, ki)], best_ki[nmp], best_dist[nmp])."""
        cfg = cfg or TrackCfg()
        kp = np.ascontiguousarray(kp_xy, np.float32).reshape(-1, 2)
        d = _rows(desc, "desc")
        pos = np.ascontiguousarray(mp_pos, np.float64).reshape(-1, 3)
        md = _rows(mp_desc, "mp_desc") if mp_desc is not None else None
        valid = np.ascontiguousarray(mp_valid, np.uint8) if mp_valid is not None else None
        R = np.ascontiguousarray(R_cam, np.float64).reshape(9)
        t = np.ascontiguousarray(t_cam, np.float64).reshape(3)
        nmp = pos.shape[0]
        assert indices.dtype == np.int32 and indices.flags.c_contiguous and len(indices) == kp.shape[0]
        obs_mp = np.zeros(max(nmp, 1), np.int32)
        obs_ki = np.zeros(max(nmp, 1), np.int32)
        bk = np.zeros(max(nmp, 1), np.int32)
        bd = np.zeros(max(nmp, 1), np.float64)
        n = C.c_int32(0)
        self._ck(self._lib.vsm_track_local_map(
            self._h, C.byref(cfg), kp.ctypes.data, d.ctypes.data, kp.shape[0], pos.ctypes.data,
            md.ctypes.data if md is not None else None, valid.ctypes.data if valid is not None else None, nmp,
            R.ctypes.data, t.ctypes.data, indices.ctypes.data, obs_mp.ctypes.data, obs_ki.ctypes.data, C.byref(n),
            bk.ctypes.data, bd.ctypes.data))
        return n.value, list(zip(obs_mp[:n.value].tolist(), obs_ki[:n.value].tolist())), bk[:nmp], bd[:nmp]

    def detect_candidates(self, frame_desc, ratio=0.75, want_matches=True):
        """LoopCloser::detect matching block: per stored keyframe, top-2 inside the keyframe +
        ratio test.  Returns (counts[nkf], [DMATCH array per keyframe] or None)."""
        q = _rows(frame_desc, "frame_desc")
        nkf = self.store_info()[1]
        counts = np.zeros(max(nkf, 1), np.int32)
        m = np.zeros((max(nkf, 1), max(q.shape[0], 1)), DMATCH) if want_matches else None
        self._ck(self._lib.vsm_db_segmented(self._h, q.ctypes.data, q.shape[0], ratio, counts.ctypes.data,
                                            m.ctypes.data if want_matches else None))
        counts = counts[:nkf]
        return counts, ([m[s, :counts[s]] for s in range(nkf)] if want_matches else None)

    def loop_detect(self, cur_frame_id, frame_desc, ratio=0.75, min_gap=200, every=5, want_matches=True):
        """LoopCloser::detect's candidate loop with its eligibility rules (src/LoopCloser.cpp:43-62).
        Returns (status[nkf]: -1 skipped / survivor count, [DMATCH array or None per keyframe])."""
        status, lists, _ = self.loop_detect_shard(cur_frame_id, frame_desc, 0, ratio, min_gap, every, want_matches)
        return status, lists

    def loop_detect_shard(self, cur_frame_id, frame_desc, checked_before, ratio=0.75, min_gap=200, every=5,
                          want_matches=True):
        """loop_detect for one shard of a partitioned keyframe list: checked_before = eligible
        (gap + non-empty) keyframes on earlier shards.  Returns (status, lists, checked_after)."""
        q = _rows(frame_desc, "frame_desc")
        nkf = self.store_info()[1]
        status = np.zeros(max(nkf, 1), np.int32)
        m = np.zeros((max(nkf, 1), max(q.shape[0], 1)), DMATCH) if want_matches else None
        after = C.c_int32(0)
        self._ck(self._lib.vsm_loop_detect_shard(self._h, cur_frame_id, min_gap, every, checked_before, q.ctypes.data,
                                                 q.shape[0], ratio, status.ctypes.data,
                                                 m.ctypes.data if want_matches else None, C.byref(after)))
        status = status[:nkf]
        lists = [m[s, :status[s]] if status[s] >= 0 else None for s in range(nkf)] if want_matches else None
        return status, lists, int(after.value)

    def loop_detect_compact(self, cur_frame_id, frame_desc, ratio=0.75, min_gap=200, every=5, min_matches=30,
                            checked_before=0):
        """LoopCloser::detect's candidate loop with the >= MIN_MATCHES gate on the device
        (src/LoopCloser.cpp:43-62).  Returns (status[nkf], {keyframe position: DMATCH list} for the
        keyframes that pass the gate, checked_after)."""
        q = _rows(frame_desc, "frame_desc")
        nkf = self.store_info()[1]
        status = np.zeros(max(nkf, 1), np.int32)
        cap_c, cap_m = 64, 64 * max(q.shape[0], 1)
        while True:
            cands = np.zeros(cap_c, CAND)
            matches = np.zeros(cap_m, DMATCH)
            nc, nm, after = C.c_int32(0), C.c_int64(0), C.c_int32(0)
            self._ck(self._lib.vsm_loop_detect_compact(self._h, cur_frame_id, min_gap, every, checked_before, q.ctypes.data,
                                                       q.shape[0], ratio, min_matches, status.ctypes.data, cands.ctypes.data,
                                                       cap_c, C.byref(nc), matches.ctypes.data, cap_m, C.byref(nm),
                                                       C.byref(after)))
            if nc.value <= cap_c and nm.value <= cap_m:
                break
            cap_c, cap_m = max(cap_c, nc.value), max(cap_m, nm.value)
        out = {int(c["keyframe"]): matches[c["offset"]:c["offset"] + c["count"]].copy() for c in cands[:nc.value]}
        return status[:nkf], out, int(after.value)

    def set_frame_ids(self, frame_ids):
        """Frame ids of the stored keyframes in store order (after adopt_device_matrix)."""
        ids = np.ascontiguousarray(frame_ids, np.int32)
        self._ck(self._lib.vsm_store_set_frame_ids(self._h, ids.ctypes.data, ids.shape[0]))

    # -- device-pointer plumbing (resident queries, sharded search) -------------------------
    def db_top2_device(self, d_query_ptr, nq, row_offset, d_idx_ptr, d_dist_ptr, sync=False):
        self._ck(self._lib.vsm_db_top2_device(self._h, C.c_void_p(d_query_ptr), nq, row_offset,
                                              C.c_void_p(d_idx_ptr), C.c_void_p(d_dist_ptr), int(sync)))

    def db_top2_keys_device(self, d_query_ptr, nq, row_offset, d_keys_ptr, sync=False):
        self._ck(self._lib.vsm_db_top2_keys_device(self._h, C.c_void_p(d_query_ptr), nq, row_offset,
                                                   C.c_void_p(d_keys_ptr), int(sync)))

    def merge_keys_device(self, d_keys_in, nshard, nq, d_idx_out, d_dist_out, sync=False):
        self._ck(self._lib.vsm_merge_keys_device(self._h, C.c_void_p(d_keys_in), nshard, nq, C.c_void_p(d_idx_out),
                                                 C.c_void_p(d_dist_out), int(sync)))

    def xchg_create(self, rank, world, nq_cap):
        """Allocate this rank's peer-exchange buffer; returns its 64-byte CUDA IPC handle."""
        h = (C.c_uint8 * 64)()
        self._ck(self._lib.vsm_xchg_create(self._h, rank, world, nq_cap, h))
        return bytes(h)

    def xchg_connect(self, handles):
        """handles: list of every rank's 64-byte IPC handle, in rank order."""
        blob = b"".join(handles)
        self._ck(self._lib.vsm_xchg_connect(self._h, blob))

    def db_top2_xchg_device(self, d_query_ptr, nq, row_offset, d_idx_ptr, d_dist_ptr, sync=False):
        self._ck(self._lib.vsm_db_top2_xchg_device(self._h, C.c_void_p(d_query_ptr), nq, row_offset,
                                                   C.c_void_p(d_idx_ptr), C.c_void_p(d_dist_ptr), int(sync)))

    def db_top2_xchg(self, query, row_offset):
        """One rank's call of the collective search with host buffers in and out (vsm_db_top2_xchg)."""
        q = _rows(query, "query")
        idx = np.empty((q.shape[0], 2), np.int64)
        dist = np.empty((q.shape[0], 2), np.float32)
        self._ck(self._lib.vsm_db_top2_xchg(self._h, q.ctypes.data, q.shape[0], row_offset, idx.ctypes.data,
                                            dist.ctypes.data))
        return idx, dist

    def merge_top2_device(self, d_idx_in, d_dist_in, nshard, nq, d_idx_out, d_dist_out, sync=False):
        self._ck(self._lib.vsm_merge_top2_device(self._h, C.c_void_p(d_idx_in), C.c_void_p(d_dist_in), nshard, nq,
                                                 C.c_void_p(d_idx_out), C.c_void_p(d_dist_out), int(sync)))

    def set_stream(self, cuda_stream):
        self._ck(self._lib.vsm_set_stream(self._h, C.c_void_p(cuda_stream)))

    def tc_history(self, n=64):
        """Device milliseconds of the tensor-core kernel in each of the last n calls (oldest first)."""
        ms = np.zeros(max(n, 1), np.float32)
        k = C.c_int32(0)
        self._ck(self._lib.vsm_tc_history(self._h, ms.ctypes.data, n, C.byref(k)))
        return ms[:k.value].copy()

    def set_profiling(self, on):
        self._ck(self._lib.vsm_set_profiling(self._h, int(on)))

    def stream(self):
        return self._lib.vsm_stream(self._h)

    def sync(self):
        self._ck(self._lib.vsm_sync(self._h))

    def debug_timeline(self):
        out = np.zeros(16384, np.int64)
        self._ck(self._lib.vsm_debug_fetch_dump(self._h, out.ctypes.data, out.nbytes))
        return out.reshape(4, 4096)

    def debug_tile_scores(self, query, train):
        q, t = _rows(query, "query"), _rows(train, "train")
        out = np.zeros((128, 256), np.float32)
        self._ck(self._lib.vsm_debug_tile_scores(self._h, q.ctypes.data, q.shape[0], t.ctypes.data, t.shape[0],
                                                 out.ctypes.data))
        return out


class Group:
    """Several GPUs behind one caller (vsm_group_*): the keyframe database is dealt to the devices by
    whole keyframes; a search fans out on the library's worker threads and merges on the first device.
    Mirrors what a single-threaded C++ caller (the reference's slam_thread) gets."""

    def __init__(self, devices, engine=ENGINE_AUTO, seg_tiles=0):
        self._lib = load_library()
        o = _Opts()
        self._lib.vsm_default_opts(C.byref(o))
        o.engine = engine
        o.reserved[0] = seg_tiles
        devs = np.ascontiguousarray(devices, np.int32)
        h = C.c_void_p()
        st = self._lib.vsm_group_create(devs.ctypes.data, len(devs), C.byref(o), C.byref(h))
        if st != 0:
            raise VsmError(f"vsm_group_create failed ({st}): {self._lib.vsm_group_last_error(None).decode()}")
        self._g = h
        self.n = len(devs)

    def close(self):
        if getattr(self, "_g", None):
            self._lib.vsm_group_destroy(self._g)
            self._g = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, st):
        if st != 0:
            raise VsmError(f"libvsm group error {st}: {self._lib.vsm_group_last_error(self._g).decode()}")

    def member(self, r):
        return Matcher._borrowed(self._lib.vsm_group_ctx(self._g, r))

    def add_keyframe(self, frame_id, desc):
        d = _rows(desc, "desc")
        h = C.c_int32(-1)
        self._ck(self._lib.vsm_group_store_add(self._g, frame_id, d.ctypes.data, d.shape[0], C.byref(h)))
        return h.value

    def remove_keyframe(self, handle):
        self._ck(self._lib.vsm_group_store_remove(self._g, handle))

    def clear_store(self):
        self._ck(self._lib.vsm_group_store_clear(self._g))

    def store_info(self):
        r, k = C.c_int64(0), C.c_int32(0)
        per = np.zeros(self.n, np.int64)
        self._ck(self._lib.vsm_group_store_info(self._g, C.byref(r), C.byref(k), per.ctypes.data))
        return r.value, k.value, per

    def adopt_device_matrix(self, member, dev_ptr, n_rows, seg_off=None):
        if seg_off is not None:
            seg_off = np.ascontiguousarray(seg_off, np.int64)
            self._ck(self._lib.vsm_group_adopt_device(self._g, member, C.c_void_p(dev_ptr), n_rows, seg_off.ctypes.data,
                                                      len(seg_off) - 1))
        else:
            self._ck(self._lib.vsm_group_adopt_device(self._g, member, C.c_void_p(dev_ptr), n_rows, None, 0))

    def search_map_points(self, frame_desc, want_keyframes=False):
        """Global top-2 over every member's keyframe rows: (idx [stacked rows], dist[, kf_handle, kf_row])."""
        q = _rows(frame_desc, "frame_desc")
        nq = q.shape[0]
        idx = np.empty((nq, 2), np.int64)
        dist = np.empty((nq, 2), np.float32)
        kh = np.empty((nq, 2), np.int32) if want_keyframes else None
        kr = np.empty((nq, 2), np.int32) if want_keyframes else None
        self._ck(self._lib.vsm_group_db_top2(self._g, q.ctypes.data, nq, idx.ctypes.data, dist.ctypes.data,
                                             kh.ctypes.data if want_keyframes else None,
                                             kr.ctypes.data if want_keyframes else None))
        return (idx, dist, kh, kr) if want_keyframes else (idx, dist)

    def match_batch_packed(self, qa, q_off, ta, t_off, ratio=0.75, mutual=False, out=None):
        """Ragged batch of pairs dealt to the members (vsm_group_match_batch); same result as Matcher.match_batch_packed."""
        n = len(q_off) - 1
        good = out if out is not None else np.zeros(max(int(q_off[-1]), 1), DMATCH)
        n_good = np.zeros(max(n, 1), np.int32)
        self._ck(self._lib.vsm_group_match_batch(self._g, n, qa.ctypes.data, q_off.ctypes.data, ta.ctypes.data,
                                                 t_off.ctypes.data, ratio, int(mutual), good.ctypes.data, n_good.ctypes.data))
        return [good[q_off[p]:q_off[p] + n_good[p]] for p in range(n)]

    def loop_detect_compact(self, cur_frame_id, frame_desc, ratio=0.75, min_gap=200, every=5, min_matches=30):
        """LoopCloser::detect's loop with the gate on the devices: (status[nkf], {keyframe position: list})."""
        q = _rows(frame_desc, "frame_desc")
        nkf = self.store_info()[1]
        status = np.zeros(max(nkf, 1), np.int32)
        cap_c, cap_m = 64, 64 * max(q.shape[0], 1)
        while True:
            cands = np.zeros(cap_c, CAND)
            matches = np.zeros(cap_m, DMATCH)
            nc, nm = C.c_int32(0), C.c_int64(0)
            self._ck(self._lib.vsm_group_loop_detect_compact(self._g, cur_frame_id, min_gap, every, q.ctypes.data, q.shape[0],
                                                             ratio, min_matches, status.ctypes.data, cands.ctypes.data, cap_c,
                                                             C.byref(nc), matches.ctypes.data, cap_m, C.byref(nm)))
            if nc.value <= cap_c and nm.value <= cap_m:
                break
            cap_c, cap_m = max(cap_c, nc.value), max(cap_m, nm.value)
        out = {int(c["keyframe"]): matches[c["offset"]:c["offset"] + c["count"]].copy() for c in cands[:nc.value]}
        return status[:nkf], out

    def loop_detect(self, cur_frame_id, frame_desc, ratio=0.75, min_gap=200, every=5, want_matches=True):
        q = _rows(frame_desc, "frame_desc")
        nkf = self.store_info()[1]
        status = np.zeros(max(nkf, 1), np.int32)
        m = np.zeros((max(nkf, 1), max(q.shape[0], 1)), DMATCH) if want_matches else None
        self._ck(self._lib.vsm_group_loop_detect(self._g, cur_frame_id, min_gap, every, q.ctypes.data, q.shape[0], ratio,
                                                 status.ctypes.data, m.ctypes.data if want_matches else None))
        status = status[:nkf]
        lists = [m[s, :status[s]] if status[s] >= 0 else None for s in range(nkf)] if want_matches else None
        return status, lists
