"""BASELINE configs[4] (64 ragged pairs, 200-2048 keypoints, mutual + ratio) from HOST buffers through one context and
through a group of all GPUs of the box (vsm_group_match_batch): the call is bound by the 150 MB upload, which the
group spreads over every member's own PCIe link.  python scripts/group_ragged_bench.py [n_gpus]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import vsm_b200

n_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
g = torch.Generator(device="cuda")
g.manual_seed(99)


def unit(n):
    x = torch.randn((n, 256), generator=g, device="cuda")
    return x / x.norm(dim=1, keepdim=True)


rng = np.random.default_rng(0)
sizes = rng.integers(200, 2049, size=(64, 2))
qs, ts = [], []
for nq, nt in sizes:
    base = unit(int(max(nq, nt)))
    nxt = unit(len(base))
    k = int(0.6 * len(base))
    v = base[:k] + 0.08 * torch.randn((k, 256), generator=g, device="cuda")
    nxt[:k] = v / v.norm(dim=1, keepdim=True)
    qs.append(base[:nq].cpu().numpy())
    ts.append(nxt[:nt].cpu().numpy())
q_off = np.zeros(65, np.int32)
t_off = np.zeros(65, np.int32)
q_off[1:] = np.cumsum(sizes[:, 0])
t_off[1:] = np.cumsum(sizes[:, 1])
qa = torch.from_numpy(np.concatenate(qs)).pin_memory().numpy()
ta = torch.from_numpy(np.concatenate(ts)).pin_memory().numpy()
out = {"bench": "ragged_batch_64_from_host", "h2d_mbytes": (qa.nbytes + ta.nbytes) / 1e6}


def timed(call):
    for _ in range(3):
        res = call()
    t = []
    for _ in range(15):
        t0 = time.perf_counter()
        res = call()
        t.append((time.perf_counter() - t0) * 1e3)
    t.sort()
    return t[len(t) // 2], res


with vsm_b200.Matcher() as m:
    ms, ref = timed(lambda: m.match_batch_packed(qa, q_off, ta, t_off, 0.75, True))
    ref = [r.copy() for r in ref]
    out["one_context_p50_ms"] = ms
with vsm_b200.Group(list(range(n_gpus))) as grp:
    ms, res = timed(lambda: grp.match_batch_packed(qa, q_off, ta, t_off, 0.75, True))
    out[f"group_{n_gpus}_gpus_p50_ms"] = ms
    out["same_matches"] = bool(all(a.tobytes() == b.tobytes() for a, b in zip(ref, res)))
print(json.dumps(out))
