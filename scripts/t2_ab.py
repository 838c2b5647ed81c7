"""A/B of the two pair-matching epilogues (tile top-2 records, the default, against the threshold-driven top-4
records, vsm_opts.reserved[5] = 1) on the small-problem configs: one 1000 x 1000 mutual + ratio pair
(BASELINE configs[1]), 2000 x 2000 ratio 0.8 (configs[0]) and 64 ragged resident pairs (configs[4]).
Ad-hoc timing; results of the two paths are compared with each other (parity against the oracle is in tests/)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import vsm_b200
from oracle import gen          # input generator only (planted matches); nothing is checked against the oracle here


def med(f, n=30, warm=5):
    for _ in range(warm):
        f()
    t = []
    for _ in range(n):
        t0 = time.perf_counter(); f(); t.append(time.perf_counter() - t0)
    t.sort()
    return t[len(t) // 2] * 1e3


out = {}
rng = np.random.default_rng(0)
sizes = rng.integers(200, 2049, size=(64, 2))
pairs = [gen.planted(1000 + p, int(sizes[p, 0]), int(sizes[p, 1]), 0.6, 0.08)[:2] for p in range(64)]
q1, t1 = gen.planted(5, 1000, 1000, 0.6, 0.08)[:2]
q2, t2 = gen.planted(6, 2000, 2000, 0.6, 0.08)[:2]
res = {}
for name, on in (("tile_top2", True), ("top4", False)):
    m = vsm_b200.Matcher(tile_top2=on)
    r = {}
    ms = med(lambda: m.match_features(q1, t1, 0.75, mutual=True, want_raw=False))
    st = m.stats(); g1 = m.match_features(q1, t1, 0.75, mutual=True, want_raw=False)[0]
    r["pair_1000_mutual"] = dict(p50_ms=ms, device_ms=st["device_ms"], tc_ms=st["tc_ms"], select_ms=st["select_ms"],
                                 candidates=st["candidates"], flagged=st["flagged_slices"], matches=len(g1))
    ms = med(lambda: m.match_features(q2, t2, 0.8, mutual=False, want_raw=False))
    st = m.stats(); g2 = m.match_features(q2, t2, 0.8, mutual=False, want_raw=False)[0]
    r["pair_2000_ratio08"] = dict(p50_ms=ms, device_ms=st["device_ms"], tc_ms=st["tc_ms"], select_ms=st["select_ms"],
                                  candidates=st["candidates"], flagged=st["flagged_slices"], matches=len(g2))
    qh = [m.add_keyframe(2 * p, pairs[p][0]) for p in range(64)]
    th = [m.add_keyframe(2 * p + 1, pairs[p][1]) for p in range(64)]
    cap = int(sizes[:, 0].sum())
    ms = med(lambda: m.match_batch_stored(qh, th, 0.75, True, capacity=cap), n=20, warm=3)
    st = m.stats(); g3 = m.match_batch_stored(qh, th, 0.75, True, capacity=cap)
    flop = sum(2.0 * a.shape[0] * b.shape[0] * 256 for a, b in pairs)
    r["ragged64_resident_mutual"] = dict(p50_ms=ms, device_ms=st["device_ms"], tc_ms=st["tc_ms"], select_ms=st["select_ms"],
                                         candidates=st["candidates"], flagged=st["flagged_slices"],
                                         matches=int(sum(len(x) for x in g3)),
                                         tc_useful_tflops=flop / (st["tc_ms"] * 1e-3) / 1e12 if st["tc_ms"] else None)
    res[name] = (g1, g2, g3)
    out[name] = r
    m.close()
a, b = res["tile_top2"], res["top4"]
out["identical"] = bool(a[0].tobytes() == b[0].tobytes() and a[1].tobytes() == b[1].tobytes() and
                        all(x.tobytes() == y.tobytes() for x, y in zip(a[2], b[2])))
print(json.dumps(out, indent=1))
