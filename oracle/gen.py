"""Deterministic synthetic descriptor generator (numpy twin of vsm_oracle_gen_rows).

TEST INFRASTRUCTURE ONLY.  Integer arithmetic end to end (splitmix64 counters ->
Irwin-Hall(8 bytes) integers -> exact int64 norm -> one float64 scale -> fp32), so
the same (seed, set_id, row) gives the same bytes on every machine and in C.

Shape follows the reference's descriptor producer (src/FeatureExtractor.cpp:170-205):
N x 256 fp32, row-major, contiguous, unit L2 norm.

Purely random unit descriptors never pass a 0.7-0.8 ratio test, so
`planted` derives a second set that re-observes a subset of rows with noise, as
consecutive video frames do.
"""
import numpy as np

DIM = 256
_G = np.uint64(0x9E3779B97F4A7C15)


def _splitmix64(x):
    x = (x + _G).astype(np.uint64)
    x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)).astype(np.uint64)
    x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)).astype(np.uint64)
    return x ^ (x >> np.uint64(31))


def _key(seed, set_id):
    with np.errstate(over="ignore"):
        a = _splitmix64(np.array([seed], dtype=np.uint64))
        b = (np.array([set_id], dtype=np.uint64) * np.uint64(0xD1342543DE82EF95)).astype(np.uint64)
        return _splitmix64(a ^ b)[0]


def int_rows(seed, set_id, row0, n):
    """n x 256 int64 Irwin-Hall(8) integers in [-1020, 1020]."""
    with np.errstate(over="ignore"):
        key = _key(seed, set_id)
        ctr = (np.arange(row0 * DIM, (row0 + n) * DIM, dtype=np.uint64) * _G).astype(np.uint64)
        x = _splitmix64((key + ctr).astype(np.uint64))
    s = np.zeros(x.shape, dtype=np.int64)
    for b in range(8):
        s += ((x >> np.uint64(8 * b)) & np.uint64(0xFF)).astype(np.int64)
    return (s - 1020).reshape(n, DIM)


def _perm(seed, stream, n):
    """Deterministic permutation of range(n): stable argsort of splitmix64 keys."""
    with np.errstate(over="ignore"):
        key = _key(seed, stream)
        x = _splitmix64((key + np.arange(n, dtype=np.uint64) * _G).astype(np.uint64))
    return np.argsort(x, kind="stable").astype(np.int64)


def _normalize_int(v):
    n2 = (v * v).sum(axis=1, dtype=np.int64)
    inv = 1.0 / np.sqrt(n2.astype(np.float64))
    return (v.astype(np.float64) * inv[:, None]).astype(np.float32)


def rows(seed, set_id, row0, n):
    """n unit-norm fp32 descriptors; bit-identical to vsm_oracle_gen_rows."""
    if n == 0:
        return np.zeros((0, DIM), np.float32)
    return _normalize_int(int_rows(seed, set_id, row0, n))


def planted(seed, nq, nt, overlap=0.6, sigma=0.08, set_q=0, set_t=1):
    """(q, t, pairs): q = rows(seed,set_q); t = rows(seed,set_t) with
    round(overlap*min(nq,nt)) rows replaced by noisy re-observations of distinct q rows.
    pairs[:,0] = q row, pairs[:,1] = t row.  Noise std per component = sigma (relative
    to a unit-norm vector), i.e. T/S = 16*sigma in integer units."""
    vq = int_rows(seed, set_q, 0, nq)
    vt = int_rows(seed, set_t, 0, nt)
    k = int(round(overlap * min(nq, nt)))
    qi = _perm(seed, 7919, nq)[:k]                      # only picks WHICH rows pair up
    ti = _perm(seed, 7920, nt)[:k]
    S = 1000
    T = int(round(16000 * sigma))
    noise = int_rows(seed, 1000 + set_t, 0, nt)
    vt = vt.copy()
    vt[ti] = S * vq[qi] + T * noise[ti]
    pairs = np.stack([qi, ti], axis=1).astype(np.int64)
    return _normalize_int(vq), _normalize_int(vt), pairs


def video(seed, nframes, n, overlap=0.6, sigma=0.06):
    """Synthetic 'video': frame k+1 re-observes overlap*n rows of frame k with noise
    (BASELINE config 2).  Returns a list of n x 256 fp32 arrays."""
    cur = int_rows(seed, 0, 0, n)
    S, T = 1000, int(round(16000 * sigma))
    out = [_normalize_int(cur)]
    k = int(round(overlap * n))
    for f in range(1, nframes):
        src = _perm(seed, 200000 + 2 * f, n)[:k]
        dst = _perm(seed, 200001 + 2 * f, n)[:k]
        nxt = int_rows(seed, f, 0, n)
        noise = int_rows(seed, 100000 + f, 0, n)
        # keep magnitudes bounded: re-quantise the parent to the Irwin-Hall scale
        parent = cur[src]
        nxt[dst] = S * parent + T * noise[dst]
        scale = np.sqrt((nxt[dst] ** 2).sum(axis=1) / (256 * 209.0 ** 2))
        nxt[dst] = np.rint(nxt[dst] / scale[:, None]).astype(np.int64)
        cur = nxt
        out.append(_normalize_int(cur))
    return out
