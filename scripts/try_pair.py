import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vsm_b200
from oracle import cases, gen, oracle
m = vsm_b200.Matcher(engine=vsm_b200.ENGINE_TENSOR_PAIR)
for name in ["pair_400_s1", "pair_129x257", "nq1", "pair_2000", "pair_777x1301", "neardup_db", "dups", "nt2"]:
    q, t = cases.PAIR_CASES[name]()
    idx, dist = m.knn_match(q, t)
    oi, od = oracle.knn(q, t, 2)
    ok = np.array_equal(idx, oi) and np.array_equal(dist.view(np.uint32), od.view(np.uint32))
    print(name, q.shape, t.shape, "OK" if ok else "MISMATCH", m.stats(), flush=True)
    if not ok:
        bad = np.nonzero((idx != oi).any(axis=1))[0]
        print("  bad queries", bad[:10], idx[bad[:3]], oi[bad[:3]])
q, t, _ = gen.planted(77, 300, 40000, 0.6, 0.08)
idx, dist = m.knn_match(q, t)
oi, od = oracle.knn(q, t, 2)
print("multi-unit", np.array_equal(idx, oi) and np.array_equal(dist.view(np.uint32), od.view(np.uint32)), m.stats())
