"""Where does the host time of vsm_match_batch_stored go?  (ad-hoc timing, no oracle)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C
import numpy as np
import torch
import vsm_b200

g = torch.Generator(device="cuda"); g.manual_seed(1)
def unit(n):
    x = torch.randn((n, 256), generator=g, device="cuda")
    return (x / x.norm(dim=1, keepdim=True)).cpu().numpy()
rng = np.random.default_rng(0)
sizes = rng.integers(200, 2049, size=(64, 2))
m = vsm_b200.Matcher()
qh = [m.add_keyframe(2 * p, unit(int(sizes[p, 0]))) for p in range(64)]
th = [m.add_keyframe(2 * p + 1, unit(int(sizes[p, 1]))) for p in range(64)]
cap = int(sizes[:, 0].sum())
for prof in (True, False):
    m.set_profiling(prof)
    for _ in range(3):
        m.match_batch_stored(qh, th, 0.75, True, capacity=cap)
    t = []
    for _ in range(20):
        t0 = time.perf_counter(); m.match_batch_stored(qh, th, 0.75, True, capacity=cap); t.append(time.perf_counter() - t0)
    t.sort()
    print("profiling", prof, "wrapper p50 ms", round(t[10] * 1e3, 3), m.stats())
    # raw ctypes call
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
