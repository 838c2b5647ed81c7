"""ctypes loader for libvsm_oracle.so (TEST INFRASTRUCTURE ONLY, see vsm_oracle.h).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libvsm_oracle.so")
DMATCH = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")])


def build(force=False):
    src = os.path.join(_HERE, "vsm_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libvsm_oracle.so"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        f32p, i64p, i32p = C.POINTER(C.c_float), C.POINTER(C.c_int64), C.POINTER(C.c_int32)
        _lib.vsm_oracle_l2sqr.restype = C.c_float
        _lib.vsm_oracle_l2sqr.argtypes = [f32p, f32p]
        _lib.vsm_oracle_knn.restype = None
        _lib.vsm_oracle_knn.argtypes = [f32p, C.c_int, C.c_int64, f32p, C.c_int64, C.c_int64,
                                        C.c_int, i64p, f32p, C.c_int]
        _lib.vsm_oracle_match_features.restype = C.c_int
        _lib.vsm_oracle_match_features.argtypes = [f32p, C.c_int, f32p, C.c_int, C.c_float, C.c_int,
                                                   C.c_void_p, i32p, C.c_void_p, i32p, C.c_int]
        _lib.vsm_oracle_segmented.restype = None
        _lib.vsm_oracle_segmented.argtypes = [f32p, C.c_int, f32p, i64p, C.c_int, C.c_float,
                                              i32p, C.c_void_p, C.c_int]
        _lib.vsm_oracle_merge_top2.restype = None
        _lib.vsm_oracle_merge_top2.argtypes = [i64p, f32p, C.c_int, C.c_int, i64p, f32p]
        _lib.vsm_oracle_gen_rows.restype = None
        _lib.vsm_oracle_gen_rows.argtypes = [C.c_uint64, C.c_uint64, C.c_int64, C.c_int64, f32p]
    return _lib


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


def l2sqr(a, b):
    a, pa = _f32(a)
    b, pb = _f32(b)
    return np.float32(lib().vsm_oracle_l2sqr(pa, pb))


def knn(q, t, k=2, threads=0):
    """(idx[nq,k] int64, dist[nq,k] fp32); missing neighbours: idx -1, dist FLT_MAX."""
    q, pq = _f32(q)
    t, pt = _f32(t)
    nq, nt = q.shape[0], t.shape[0]
    idx = np.empty((nq, k), np.int64)
    dist = np.empty((nq, k), np.float32)
    lib().vsm_oracle_knn(pq, nq, 256, pt, nt, 256, k,
                         idx.ctypes.data_as(C.POINTER(C.c_int64)),
                         dist.ctypes.data_as(C.POINTER(C.c_float)), threads)
    return idx, dist


def match_features(q, t, ratio=0.75, mutual=False, threads=0):
    """Slam::match_features (src/Slam.cpp:1140-1172).  Returns (good, raw) DMATCH arrays."""
    q, pq = _f32(q)
    t, pt = _f32(t)
    nq, nt = q.shape[0], t.shape[0]
    good = np.zeros(max(nq, 1), DMATCH)
    raw = np.zeros(max(nq, 1), DMATCH)
    ng, nr = C.c_int32(0), C.c_int32(0)
    lib().vsm_oracle_match_features(pq, nq, pt, nt, ratio, int(mutual),
                                    good.ctypes.data, C.byref(ng), raw.ctypes.data, C.byref(nr), threads)
    return good[:ng.value].copy(), raw[:nr.value].copy()


def segmented(q, db, seg_off, ratio=0.75, threads=0):
    """LoopCloser::detect matching block (src/LoopCloser.cpp:43-62).
    Returns (counts[nseg], list of DMATCH arrays per segment)."""
    q, pq = _f32(q)
    db, pdb = _f32(db)
    seg_off = np.ascontiguousarray(seg_off, np.int64)
    nseg, nq = len(seg_off) - 1, q.shape[0]
    counts = np.zeros(nseg, np.int32)
    m = np.zeros((nseg, max(nq, 1)), DMATCH)
    lib().vsm_oracle_segmented(pq, nq, pdb, seg_off.ctypes.data_as(C.POINTER(C.c_int64)), nseg, ratio,
                               counts.ctypes.data_as(C.POINTER(C.c_int32)), m.ctypes.data, threads)
    return counts, [m[s, :counts[s]].copy() for s in range(nseg)]


def merge_top2(idx_in, dist_in):
    idx_in = np.ascontiguousarray(idx_in, np.int64)
    dist_in = np.ascontiguousarray(dist_in, np.float32)
    ns, nq = idx_in.shape[0], idx_in.shape[1]
    io = np.empty((nq, 2), np.int64)
    do = np.empty((nq, 2), np.float32)
    lib().vsm_oracle_merge_top2(idx_in.ctypes.data_as(C.POINTER(C.c_int64)),
                                dist_in.ctypes.data_as(C.POINTER(C.c_float)), ns, nq,
                                io.ctypes.data_as(C.POINTER(C.c_int64)),
                                do.ctypes.data_as(C.POINTER(C.c_float)))
    return io, do


def gen_rows(seed, set_id, row0, n):
    out = np.empty((n, 256), np.float32)
    lib().vsm_oracle_gen_rows(seed, set_id, row0, n, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def loop_detect(q, db, seg_off, frame_ids, cur_frame_id, ratio=0.75, min_gap=200, every=5, threads=0, checked0=0):
    """LoopCloser::detect's candidate loop (src/LoopCloser.cpp:43-62): eligibility (:44-48) then the
    per-keyframe kNN + ratio test.  Returns status[nkf] (-1 = skipped, else survivors) and the lists.
    checked0: value of the reference's `checked` counter on entry (0 for the whole list; the tests of
    the partitioned search start a later shard where the earlier ones left off)."""
    nkf = len(seg_off) - 1
    status = -np.ones(nkf, np.int32)
    lists = [None] * nkf
    checked = checked0
    for s in range(nkf):
        if cur_frame_id - frame_ids[s] < min_gap:
            continue
        if seg_off[s + 1] == seg_off[s]:
            continue
        checked += 1
        if checked % every != 0:
            continue
        good, _ = match_features(q, db[seg_off[s]:seg_off[s + 1]], ratio, threads=threads)
        good = good.copy()
        good["imgIdx"] = s
        status[s] = len(good)
        lists[s] = good
    return status, lists


def track_local_map(kp_xy, desc, mp_pos, mp_desc, mp_valid, R_cam, t_cam, indices, cfg=None, norm_fn=None):
    """Slam::track_local_map (src/Slam.cpp:380-469) restated loop for loop: the 30-px cell grid
    (:391-401), fp64 projection (:417-428), the window of cells (:431-434), the radius test
    (:452-454), cv::norm in double (:456; here numpy float64, see include/vsm.h on its summation
    order -- or norm_fn(mp_desc_row, desc_row), e.g. cv2.norm itself, which is how the golden
    vectors are made) and the sequential assignment (:465-470).  indices is updated in place.
    Returns (tracked, observations, best_ki[nmp], best_dist[nmp])."""
    c = dict(fx=525.0, fy=525.0, cx=319.5, cy=239.5, width=640, height=480, cell_size=30,
             depth_min=float(np.float32(0.1)), depth_max=50.0, search_radius=12.0, desc_threshold=0.5)
    c.update(cfg or {})
    kp = np.asarray(kp_xy, np.float32).reshape(-1, 2)
    nkp, nmp = kp.shape[0], len(mp_pos)
    best_ki = -np.ones(nmp, np.int32)
    best_dist = np.full(nmp, c["desc_threshold"], np.float64)
    if nkp == 0 or nmp == 0:
        return 0, [], best_ki, best_dist
    CELL = c["cell_size"]
    GW, GH = (c["width"] + CELL - 1) // CELL, (c["height"] + CELL - 1) // CELL
    grid = [[] for _ in range(GW * GH)]
    for ki in range(nkp):
        gx = min(int(np.float32(kp[ki, 0]) / np.float32(CELL)), GW - 1)     # int() truncates like the C cast
        gy = min(int(np.float32(kp[ki, 1]) / np.float32(CELL)), GH - 1)
        if gx >= 0 and gy >= 0:
            grid[gy * GW + gx].append(ki)
    R = np.asarray(R_cam, np.float64).reshape(3, 3)
    t = np.asarray(t_cam, np.float64).reshape(3)
    RAD = c["search_radius"]
    d64 = np.asarray(desc, np.float32).astype(np.float64)
    best_desc_dist = np.full(nkp, 1e9)
    tracked, obs = 0, []
    for mp in range(nmp):
        if mp_valid is not None and not mp_valid[mp]:
            continue
        X, Y, Z = (float(v) for v in mp_pos[mp])
        px = R[0, 0] * X + R[0, 1] * Y + R[0, 2] * Z + t[0]
        py = R[1, 0] * X + R[1, 1] * Y + R[1, 2] * Z + t[1]
        pz = R[2, 0] * X + R[2, 1] * Y + R[2, 2] * Z + t[2]
        if pz < c["depth_min"] or pz > c["depth_max"]:
            continue
        u = c["fx"] * px / pz + c["cx"]
        v = c["fy"] * py / pz + c["cy"]
        if u < 0 or u >= c["width"] or v < 0 or v >= c["height"]:
            continue
        gx0, gy0 = max(0, int((u - RAD) / CELL)), max(0, int((v - RAD) / CELL))
        gx1, gy1 = min(GW - 1, int((u + RAD) / CELL)), min(GH - 1, int((v + RAD) / CELL))
        bk, bd = -1, c["desc_threshold"]
        m64 = np.asarray(mp_desc[mp], np.float32).astype(np.float64)
        for gy in range(gy0, gy1 + 1):
            for gx in range(gx0, gx1 + 1):
                for ki in grid[gy * GW + gx]:
                    dx, dy = u - float(kp[ki, 0]), v - float(kp[ki, 1])
                    if dx * dx + dy * dy > RAD * RAD:
                        continue
                    if norm_fn is not None:
                        dist = float(norm_fn(mp_desc[mp], desc[ki]))
                    else:
                        dd = m64 - d64[ki]
                        dist = float(np.sqrt(np.dot(dd, dd)))
                    if dist < bd:
                        bd, bk = dist, ki
        best_ki[mp], best_dist[mp] = bk, bd
        if bk >= 0 and bd < best_desc_dist[bk]:
            indices[bk] = mp
            best_desc_dist[bk] = bd
            obs.append((mp, bk))
            tracked += 1
    return tracked, obs, best_ki, best_dist
