"""Where does the host time of vsm_match_batch_stored go?  (ad-hoc timing, no oracle)
64 ragged resident pairs, random rows (no matches) and planted rows (~31K matches): Python wrapper against the raw
ctypes call, profiling on / off."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import vsm_b200

g = torch.Generator(device="cuda"); g.manual_seed(1)
def unit(n):
    x = torch.randn((n, 256), generator=g, device="cuda")
    return x / x.norm(dim=1, keepdim=True)
rng = np.random.default_rng(0)
sizes = rng.integers(200, 2049, size=(64, 2))
for planted in (False, True):
    m = vsm_b200.Matcher()
    qh, th = [], []
    for p in range(64):
        nq, nt = int(sizes[p, 0]), int(sizes[p, 1])
        base = unit(max(nq, nt))
        nxt = unit(max(nq, nt))
        if planted:
            k = int(0.6 * len(base))
            v = base[:k] + 0.08 * torch.randn((k, 256), generator=g, device="cuda")
            nxt[:k] = v / v.norm(dim=1, keepdim=True)
        qh.append(m.add_keyframe(2 * p, base[:nq].cpu().numpy()))
        th.append(m.add_keyframe(2 * p + 1, nxt[:nt].cpu().numpy()))
    cap = int(sizes[:, 0].sum())
    for prof in (True, False):
        m.set_profiling(prof)
        for _ in range(3):
            res = m.match_batch_stored(qh, th, 0.75, True, capacity=cap)
        t = []
        for _ in range(20):
            t0 = time.perf_counter(); m.match_batch_stored(qh, th, 0.75, True, capacity=cap); t.append(time.perf_counter() - t0)
        t.sort()
        print("planted", planted, "profiling", prof, "matches", sum(len(r) for r in res), "wrapper p50 ms", round(t[10] * 1e3, 3), m.stats())
        qa = np.ascontiguousarray(qh, np.int32); ta = np.ascontiguousarray(th, np.int32)
        good = np.zeros(cap, vsm_b200.DMATCH); ng = np.zeros(64, np.int32); off = np.zeros(65, np.int64)
        lib = m.lib
        t = []
        for _ in range(20):
            t0 = time.perf_counter()
            lib.vsm_match_batch_stored(m.handle, 64, qa.ctypes.data, ta.ctypes.data, 0.75, 1, good.ctypes.data, cap, ng.ctypes.data, off.ctypes.data)
            t.append(time.perf_counter() - t0)
        t.sort()
        print("   raw call p50 ms", round(t[10] * 1e3, 3))
    m.close()
