import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vsm_b200
from oracle import gen, oracle
nq, nt = int(sys.argv[1]), int(sys.argv[2])
m = vsm_b200.Matcher(engine=int(os.environ.get("VSM_ENGINE", 1)), seg_tiles=int(os.environ.get("SEG", 0)))
q, t, _ = gen.planted(7, nq, nt, 0.5, 0.08)
t0 = time.time()
try:
    idx, dist = m.knn_match(q, t)
    oi, od = oracle.knn(q, t, 2)
    print(nq, nt, "OK" if np.array_equal(idx, oi) else "MISMATCH", round(time.time() - t0, 3), "s")
except Exception as e:
    print(nq, nt, "FAIL after", round(time.time() - t0, 3), "s", str(e)[:100])
