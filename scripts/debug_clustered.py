import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import vsm_b200
from oracle import oracle

def unit(x): return x / x.norm(dim=1, keepdim=True)
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 256
g = torch.Generator(device="cuda"); g.manual_seed(77)
ncl = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
centres = unit(torch.randn((ncl, 256), generator=g, device="cuda"))
per = (rows + ncl - 1) // ncl
cl = (torch.arange(rows, device="cuda") // per).clamp(max=ncl - 1)
db = unit(centres[cl] + 0.05 * torch.randn((rows, 256), generator=g, device="cuda"))
qc = torch.randint(0, ncl, (nq,), generator=g, device="cuda")
q = unit(centres[qc] + 0.05 * torch.randn((nq, 256), generator=g, device="cuda"))
host = db.cpu().numpy(); hq = q.cpu().pin_memory().numpy() if os.environ.get('PINNED') else q.cpu().numpy()
sub = np.arange(0, nq, max(1, nq // 64))
oi_s, od_s = oracle.knn(np.ascontiguousarray(hq[sub]), host, 2)
oi = -np.ones((nq, 2), np.int64); od = np.zeros((nq, 2), np.float32)
oi[sub] = oi_s; od[sub] = od_s
segs = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0, 1, 2, 4, 8, 16, 32, 64]
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 1
for seg in segs:
    for eng in (vsm_b200.ENGINE_TENSOR,):
        with vsm_b200.Matcher(engine=eng, seg_tiles=seg) as m:
            m.adopt_device_matrix(db.data_ptr(), rows)
            for rep in range(reps):
                gi, gd = m.search_map_points(hq)
                st = m.stats()
                bad = sub[np.nonzero(~((gi[sub] == oi[sub]).all(axis=1)))[0]]
                print("seg", seg, "rep", rep, "mismatch", len(bad), "flagged", st["flagged_slices"], "cand", st["candidates"], flush=True)
        for b in bad[:3]:
            print("   q", b, "oracle", oi[b], od[b], "gpu", gi[b], gd[b], "tile of oracle rows", oi[b] // 256, "cluster rows", int(qc[b]) * per, (int(qc[b]) + 1) * per)
