"""The key arithmetic of the tile top-2 epilogue (visual-slam-pipeline_b200/csrc/vsm_common.cuh, t2_scale ff.;
vsm_tc.cuh, t2_step), restated in numpy and checked on the CPU: the three float operations that turn an
accumulator value into a key are exact after the first rounding, keys order like (quantised value, column), and the
value read back from a key is within 2^-17 / s of the accumulator value -- the bound DESIGN.md quotes against
dot_margin's packing allowance.  (The kernels themselves are checked against the oracle in the -m gpu tests.)"""
import numpy as np

F = np.float32


def fma32(a, b, c):
    # a * b is exact in float64 (24 + 24 significant bits), the sum of it and a float32 of this magnitude too
    return (a.astype(np.float64) * np.float64(b) + np.float64(c)).astype(F)


def t2_scale(qn2, tmax2):
    return F(0.48) / max(F(1.01) * np.sqrt(F(qn2)) * np.sqrt(F(tmax2)), F(1e-20))


def make_keys(v, s, col):
    q = fma32(v, s, 192.0)
    a = (q + F(-190.5)).astype(F)
    k = (a + (col.astype(F) * F(2.0 ** -23))).astype(F)
    # both additions must be exact
    assert np.array_equal(a.astype(np.float64), q.astype(np.float64) - 190.5)
    assert np.array_equal(k.astype(np.float64), a.astype(np.float64) + col.astype(np.float64) * 2.0 ** -23)
    return k


def decode(k, s):
    bits = k.view(np.uint32)
    n = ((bits & np.uint32(0x7FFFFF)) >> np.uint32(7)).astype(np.int64) - 32768
    return n.astype(np.float64) * 2.0 ** -16 / np.float64(s), (bits & np.uint32(127)).astype(np.int64)


def test_keys_are_exact_ordered_and_decodable():
    rng = np.random.default_rng(0)
    for qn, tn in ((1.0, 1.0), (3.7, 0.01), (1e-3, 250.0), (1.0, 1.003)):
        s = t2_scale(qn * qn, tn * tn)
        # accumulator values of bf16-rounded operands may exceed |q||t| by a hair: up to 1.008 |q||t| here
        v = (rng.uniform(-1.008, 1.008, size=20000) * qn * tn).astype(F)
        v[:200] = np.repeat(v[200:300], 2)                       # equal values in different columns
        col = rng.integers(0, 128, size=v.size)
        k = make_keys(v, s, col)
        assert np.all(k >= F(1.0)) and np.all(k < F(2.0))
        val, c = decode(k, s)
        assert np.array_equal(c, col)
        assert np.max(np.abs(val - v.astype(np.float64))) <= 2.0 ** -17 / np.float64(s) * (1 + 1e-6)
        assert 2.0 ** -17 / np.float64(s) < 1.7e-5 * qn * tn
        # order: by quantised value, then by column
        order = np.argsort(k, kind="stable")
        n = np.rint(v.astype(np.float64) * np.float64(s) * 2.0 ** 16)
        ref = np.lexsort((col, n))
        assert np.array_equal(k[order], k[ref])
        # a masked column (MASKED_VALUE) and an empty slot are below every real key
        masked = make_keys(np.array([-3.0e38], F), s, np.array([5]))
        assert not (masked[0] >= F(1.0))


def test_top2_tree_matches_sort():
    """The comparison tree of t2_step (sorted pairs, top-2 merges) returns the two largest keys of 128."""
    rng = np.random.default_rng(1)
    for _ in range(200):
        k = rng.uniform(1.0, 2.0, size=128).astype(F)
        if rng.random() < 0.3:
            k[rng.integers(0, 128, 5)] = k[0]
        H, L = F(-np.inf), F(-np.inf)
        for c0 in range(0, 128, 32):                              # chunks of 32, sub-blocks of 4, as in the kernel
            CH, CL = F(-np.inf), F(-np.inf)
            for b in range(c0, c0 + 32, 4):
                h0, l0 = max(k[b], k[b + 1]), min(k[b], k[b + 1])
                h1, l1 = max(k[b + 2], k[b + 3]), min(k[b + 2], k[b + 3])
                h, l = max(h0, h1), max(min(h0, h1), l0, l1)
                CH, CL = max(CH, h), max(min(CH, h), CL, l)
            H, L = max(H, CH), max(min(H, CH), L, CL)
        top = np.sort(k)[::-1]
        assert H == top[0] and L == top[1]
