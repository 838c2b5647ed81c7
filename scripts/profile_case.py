"""Small driver for ncu / sweeps: runs one named case a few times and prints the library's own
event timings.  Usage: python scripts/profile_case.py <case> [reps]
  pair1000   1000x1000 match_features, mutual + ratio (BASELINE configs[1])
  pair2000   2000x2000 knn + ratio 0.8 (configs[0])
  ragged64   64 ragged pairs in one call (configs[4])
  seg:<nkf>:<rows>:<nq>  LoopCloser block: per-keyframe top-2 + ratio over nkf stored keyframes
  loop:<nkf>:<rows>:<nq>  LoopCloser::detect in compact form (fused dismissal, vsm_loop_detect_compact)
  db:<rows>:<nq>   global top-2 of nq queries over a <rows>-row resident DB (configs[2]/[3])
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import vsm_b200


def unit(n, g):
    x = torch.randn((n, 256), generator=g, device="cuda")
    return x / x.norm(dim=1, keepdim=True)


def main():
    case = sys.argv[1]
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    m = vsm_b200.Matcher(engine=int(os.environ.get("VSM_ENGINE", vsm_b200.ENGINE_TENSOR)))
    if case.startswith("pair"):
        n = int(case[4:])
        a = unit(n, g)
        b = unit(n, g)
        k = int(0.6 * n)
        v = a[:k] + 0.06 * torch.randn((k, 256), generator=g, device="cuda")
        b[:k] = v / v.norm(dim=1, keepdim=True)
        ha = a.cpu().pin_memory().numpy()
        hb = b.cpu().pin_memory().numpy()
        for r in range(reps + 3):
            t0 = time.perf_counter()
            good, _ = m.match_features(ha, hb, 0.75, mutual=True, want_raw=False)
            dt = time.perf_counter() - t0
            print(case, "wall_us", round(dt * 1e6, 1), "matches", len(good), m.stats())
    elif case == "ragged64":
        # BASELINE configs[4]: 64 pairs, sizes U{200..2048}, mutual + ratio, one launch sequence
        rng = np.random.default_rng(0)
        sizes = rng.integers(200, 2049, size=(64, 2))
        qs, tsets = [], []
        for nq, nt in sizes:
            base = unit(int(max(nq, nt)), g)
            k = int(0.6 * len(base))
            nxt = unit(len(base), g)
            v = base[:k] + 0.08 * torch.randn((k, 256), generator=g, device="cuda")
            nxt[:k] = v / v.norm(dim=1, keepdim=True)
            qs.append(base[:nq].cpu().numpy())
            tsets.append(nxt[:nt].cpu().numpy())
        q_off = np.zeros(65, np.int32); t_off = np.zeros(65, np.int32)
        q_off[1:] = np.cumsum(sizes[:, 0]); t_off[1:] = np.cumsum(sizes[:, 1])
        qa = torch.from_numpy(np.concatenate(qs)).pin_memory().numpy()
        ta = torch.from_numpy(np.concatenate(tsets)).pin_memory().numpy()
        for r in range(reps + 3):
            t0 = time.perf_counter()
            res = m.match_batch_packed(qa, q_off, ta, t_off, 0.75, True)
            dt = time.perf_counter() - t0
            print(case, "wall_ms", round(dt * 1e3, 3), "matches", int(sum(len(x) for x in res)), m.stats())
    elif case.startswith("loop"):
        _, nkf, rows, nq = case.split(":")
        nkf, rows, nq = int(nkf), int(rows), int(nq)
        db = torch.empty((nkf * rows, 256), device="cuda")
        for c0 in range(0, nkf * rows, 1 << 20):
            n = min(1 << 20, nkf * rows - c0)
            db[c0:c0 + n] = unit(n, g)
        q = unit(nq, g)
        src = (nkf // 2) * rows + torch.randperm(rows, generator=g, device="cuda")[:nq // 5]
        v = db[src] + 0.06 * torch.randn((nq // 5, 256), generator=g, device="cuda")
        q[:nq // 5] = v / v.norm(dim=1, keepdim=True)
        torch.cuda.synchronize()
        m.adopt_device_matrix(db.data_ptr(), nkf * rows, np.arange(nkf + 1, dtype=np.int64) * rows)
        hq = q.cpu().pin_memory().numpy()
        for r in range(reps + 3):
            t0 = time.perf_counter()
            st_, lists, _ = m.loop_detect_compact(10**6, hq, 0.75, min_gap=0, every=1, min_matches=30)
            dt = time.perf_counter() - t0
            st = m.stats()
            tf = 2.0 * nq * nkf * rows * 256 / (st["tc_ms"] * 1e-3) / 1e12
            print(case, "wall_ms", round(dt * 1e3, 3), "tc_TFLOPs", round(tf, 1), "candidates", sorted(lists), st)
    elif case.startswith("seg"):
        _, nkf, rows, nq = case.split(":")
        nkf, rows, nq = int(nkf), int(rows), int(nq)
        db = unit(nkf * rows, g)
        q = unit(nq, g)
        torch.cuda.synchronize()
        seg = np.arange(nkf + 1, dtype=np.int64) * rows
        m.adopt_device_matrix(db.data_ptr(), nkf * rows, seg)
        hq = q.cpu().pin_memory().numpy()
        for r in range(reps + 3):
            t0 = time.perf_counter()
            counts, _ = m.detect_candidates(hq, 0.75, want_matches=False)
            dt = time.perf_counter() - t0
            st = m.stats()
            tf = 2.0 * nq * nkf * rows * 256 / (st["tc_ms"] * 1e-3) / 1e12
            print(case, "wall_ms", round(dt * 1e3, 3), "tc_TFLOPs", round(tf, 1), "max_count", int(counts.max()), st)
    else:
        _, rows, nq = case.split(":")
        rows, nq = int(rows), int(nq)
        db = torch.empty((rows, 256), device="cuda")
        for c0 in range(0, rows, 1 << 20):
            n = min(1 << 20, rows - c0)
            db[c0:c0 + n] = unit(n, g)
        q = unit(nq, g)
        torch.cuda.synchronize()
        m.adopt_device_matrix(db.data_ptr(), rows)
        oi = torch.empty((nq, 2), dtype=torch.int64, device="cuda")
        od = torch.empty((nq, 2), dtype=torch.float32, device="cuda")
        for r in range(reps + 3):
            m.db_top2_device(q.data_ptr(), nq, 0, oi.data_ptr(), od.data_ptr(), sync=True)
            st = m.stats()
            tf = 2.0 * nq * rows * 256 / (st["tc_ms"] * 1e-3) / 1e12
            print(case, "tc_TFLOPs", round(tf, 1), st)
        if os.environ.get("VSM_DEBUG_TIMELINE"):
            tl = m.debug_timeline()
            nt = int((tl[0] > 0).sum())
            t0 = tl[1][0]
            print("tiles", nt)
            for n in list(range(0, min(nt, 24))) + list(range(24, nt, max(1, nt // 40))):
                e = int(tl[3][n])
                print(n, "load_issued", tl[1][n] - t0, "mma_issue", (int(tl[2][n]) & 0xFFFFFFFFFF) - (int(t0) & 0xFFFFFFFFFF), "mma_wait_full", int(tl[2][n]) >> 40, "acc_ready", tl[0][n] - t0,
                      "epi_done", (e & 0xFFFFFFFFFF) - (int(t0) & 0xFFFFFFFFFF), "slow_groups", (e >> 40) & 0xFF, "hint_seen", (e >> 48) & 1)
    m.close()


if __name__ == "__main__":
    main()
