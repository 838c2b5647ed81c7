"""Randomised parity fuzz of pair matching on tile top-2 records against the CPU oracle (test infrastructure, run by
hand on the GPU box: python scripts/t2_fuzz.py [cases] [seed]).  Sizes up to 2100 x 2100, ratios 0.55 .. 1.0, with and
without the mutual test; data = planted pairs with a noise sweep (fills the undecidable band), blocks of near-duplicates
and exact duplicates (ties, hidden third candidates), rows of unequal norms."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import vsm_b200
from oracle import gen, oracle

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
m = vsm_b200.Matcher()
bad = 0
for it in range(cases):
    nq, nt = int(rng.integers(1, 2100)), int(rng.integers(1, 2100))
    s = int(rng.integers(0, 1 << 30))
    vt = gen.int_rows(s, 1, 0, nt).copy()
    vq = gen.int_rows(s, 0, 0, nq).copy()
    k = int(rng.uniform(0, 0.8) * min(nq, nt))
    if k:
        rows = rng.permutation(nt)[:k]
        amp = rng.integers(200, 3000, size=(k, 1))
        vq[:k] = 1000 * vt[rows] + amp * gen.int_rows(s, 2, 0, k)
    if nt > 40 and rng.random() < 0.5:                      # a block of near-duplicates of one row, and exact copies
        b0, bl = int(rng.integers(0, nt - 30)), int(rng.integers(2, 30))
        centre = vt[b0].copy()                              # (entries <= 832: no int64 overflow in the squared norms below)
        vt[b0:b0 + bl] = 1000 * centre + int(rng.integers(0, 40)) * gen.int_rows(s, 3, 0, bl)
        if nq > 8:
            vq[-8:] = 1000 * centre + 30 * gen.int_rows(s, 4, 0, 8)
    q, t = gen._normalize_int(vq), gen._normalize_int(vt)
    if rng.random() < 0.25:
        t = np.ascontiguousarray(t * rng.uniform(0.6, 1.6, size=(nt, 1)).astype(np.float32))
    if rng.random() < 0.15:
        q = np.ascontiguousarray(q * np.float32(rng.uniform(0.01, 50)))
    ratio = float(rng.choice([0.55, 0.7, 0.75, 0.8, 0.9, 0.97, 1.0]))
    mutual = bool(rng.integers(0, 2))
    og, _ = oracle.match_features(q, t, ratio, mutual=mutual)
    good, _ = m.match_features(q, t, ratio, mutual=mutual, want_raw=False)
    ok = good.tobytes() == og.tobytes()
    if ok and it % 7 == 0:
        res = m.match_batch([q, q[: max(1, nq // 3)]], [t, t[: max(1, nt // 2)]], ratio, mutual=mutual)
        og2, _ = oracle.match_features(q[: max(1, nq // 3)], t[: max(1, nt // 2)], ratio, mutual=mutual)
        ok = res[0].tobytes() == og.tobytes() and res[1].tobytes() == og2.tobytes()
    if not ok:
        bad += 1
        print("MISMATCH case", it, "seed", s, nq, nt, ratio, mutual, len(good), len(og))
print("fuzz:", cases, "cases, seed", seed, "mismatches", bad)
sys.exit(1 if bad else 0)
