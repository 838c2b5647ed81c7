"""Named, seeded parity cases shared by make_golden.py and tests/ (TEST INFRASTRUCTURE ONLY).

Each case regenerates its inputs from integers (oracle/gen.py), so only the
answers are committed under tests/golden/.
"""
import numpy as np
from . import gen


def _dups():
    t = gen.rows(5, 1, 0, 64).copy()
    t[40] = t[7]
    t[45] = t[7]
    t[3] = t[60]
    q = np.concatenate([t[7:8], t[60:61], gen.rows(5, 0, 0, 30)])
    return q, t


def _neardup_db():
    # many near-identical rows in the train set: stresses bf16 candidate overflow
    base = gen.int_rows(21, 1, 0, 600).copy()
    noise = gen.int_rows(21, 2, 0, 600)
    base[100:400] = 1000 * base[50] + noise[100:400]      # 300 rows within ~1e-3 of row 50
    t = gen._normalize_int(base)
    qv = gen.int_rows(21, 0, 0, 130).copy()
    qv[:40] = 1000 * gen.int_rows(21, 1, 50, 1) + 30 * gen.int_rows(21, 3, 0, 40)
    return gen._normalize_int(qv), t


def _mutual_conflict():
    # two query rows re-observe the SAME train row: both pass the ratio test, only the
    # closer one is the train row's own nearest neighbour, so the mutual filter bites.
    vt = gen.int_rows(31, 1, 0, 500)
    vq = gen.int_rows(31, 0, 0, 640).copy()
    rows = gen._perm(31, 5, 500)[:200]
    vq[0:200] = 1000 * vt[rows] + 800 * gen.int_rows(31, 2, 0, 200)
    vq[300:500] = 1000 * vt[rows] + 1000 * gen.int_rows(31, 3, 0, 200)
    return gen._normalize_int(vq), gen._normalize_int(vt)


PAIR_CASES = {
    # BASELINE configs[0]: 2000x2000, k=2, ratio 0.8 (also 0.75 / 0.70)
    "pair_2000": lambda: gen.planted(0, 2000, 2000, 0.6, 0.08)[:2],
    # BASELINE configs[1]: 1000x1000 consecutive frames
    "pair_1000_video": lambda: tuple(gen.video(0, 2, 1000)),
    # reference scale (SP_MAX_KEYPOINTS = 400, include/Config.h:42)
    "pair_400_s1": lambda: gen.planted(1, 400, 400, 0.6, 0.08)[:2],
    "pair_400_s2": lambda: gen.planted(2, 400, 400, 0.5, 0.10)[:2],
    # ragged, not multiples of any tile
    "pair_777x1301": lambda: gen.planted(3, 777, 1301, 0.6, 0.08)[:2],
    "pair_2048x200": lambda: gen.planted(4, 2048, 200, 0.6, 0.08)[:2],
    "pair_129x257": lambda: gen.planted(6, 129, 257, 0.6, 0.08)[:2],
    # edge cases the reference guards (src/Slam.cpp:1143,1152)
    "mutual_conflict": _mutual_conflict,
    "dups": _dups,
    "neardup_db": _neardup_db,
    "nt1": lambda: (gen.rows(9, 0, 0, 17), gen.rows(9, 1, 0, 1)),
    "nt2": lambda: (gen.rows(9, 0, 0, 17), gen.rows(9, 1, 0, 2)),
    "nt3": lambda: (gen.rows(9, 0, 0, 17), gen.rows(9, 1, 0, 3)),
    "nq1": lambda: (gen.rows(9, 0, 0, 1), gen.rows(9, 1, 0, 100)),
    "scaled": lambda: (gen.rows(11, 0, 0, 50) * np.float32(3.7), gen.rows(11, 1, 0, 300) * np.float32(0.01)),
}

RATIOS = (0.70, 0.75, 0.80)   # include/Config.h:53-56 and BASELINE configs[0]


def db_case(seed=0, nq=300, nkf=24, lo=40, hi=700, planted_frac=0.3):
    """Small keyframe DB: nkf keyframes of variable size; a share of the queries are noisy
    re-observations of rows of two of the keyframes (loop-closure candidates).
    Returns (q, db, seg_off)."""
    sizes = lo + (gen._perm(seed, 31, hi - lo)[:nkf])
    sizes[3] = 0                                  # an empty keyframe (LoopCloser.cpp:45)
    sizes[5] = 1                                  # fewer than 2 rows (size()>=2 guard)
    seg_off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    n = int(seg_off[-1])
    vdb = gen.int_rows(seed, 50, 0, n)
    vq = gen.int_rows(seed, 51, 0, nq).copy()
    k = int(nq * planted_frac)
    for kf, sl in ((7, slice(0, k // 2)), (15, slice(k // 2, k))):
        cnt = sl.stop - sl.start
        rows = seg_off[kf] + gen._perm(seed, 60 + kf, int(sizes[kf]))[:cnt]
        vq[sl.start:sl.start + len(rows)] = 1000 * vdb[rows] + 1100 * gen.int_rows(seed, 70 + kf, 0, len(rows))
    return gen._normalize_int(vq), gen._normalize_int(vdb), seg_off


# ---- Slam::track_local_map scenes (src/Slam.cpp:380-469), shared by make_golden.py and the tests ----
def track_scene(seed, nmp=3000, nkp=800, drop=0.3, dup=True):
    rng = np.random.default_rng(seed)
    # camera pose: small rotation about y + translation (world -> camera)
    a = 0.05
    R = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
    t = np.array([0.1, -0.05, 0.2])
    pos = np.stack([rng.uniform(-6, 6, nmp), rng.uniform(-4, 4, nmp), rng.uniform(-1, 12, nmp)], axis=1)
    mp_desc = gen.rows(seed, 0, 0, nmp)
    valid = (rng.random(nmp) > 0.1).astype(np.uint8)
    cam = (R @ pos.T).T + t
    z = cam[:, 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        u = 525.0 * cam[:, 0] / z + 319.5
        v = 525.0 * cam[:, 1] / z + 239.5
    vis = np.nonzero((z > 0.2) & (u >= 0) & (u < 640) & (v >= 0) & (v < 480))[0]
    pick = rng.permutation(vis)[:int(nkp * (1 - drop))]
    kp = np.zeros((nkp, 2), np.float32)
    desc = gen.rows(seed, 1, 0, nkp).copy()
    k = len(pick)
    kp[:k, 0] = (u[pick] + rng.normal(0, 3.0, k)).astype(np.float32)
    kp[:k, 1] = (v[pick] + rng.normal(0, 3.0, k)).astype(np.float32)
    noisy = mp_desc[pick] + 0.02 * rng.standard_normal((k, 256)).astype(np.float32)
    desc[:k] = noisy / np.linalg.norm(noisy, axis=1, keepdims=True)
    kp[k:, 0] = rng.uniform(0, 640, nkp - k)
    kp[k:, 1] = rng.uniform(0, 480, nkp - k)
    if dup and k > 40:
        # two map points at the same place with the same descriptor: the later one must NOT replace
        pos[pick[1]] = pos[pick[0]]
        mp_desc[pick[1]] = mp_desc[pick[0]]
        # a keypoint duplicated in a neighbouring cell position: visiting order decides the tie
        kp[k] = kp[5] + np.float32(0.25)
        desc[k] = desc[5]
    kp = np.clip(kp, 0, [639.5, 479.5]).astype(np.float32)
    return kp, desc, pos, mp_desc, valid, R, t
