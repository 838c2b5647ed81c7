#!/usr/bin/env python
"""bench.py -- the loop-closure descriptor search (BASELINE.json configs[3]) on N B200s.

Step = one query batch (2000 x 256-d fp32 SuperPoint-shaped descriptors) answered with its exact
global top-2 over a 20M-descriptor keyframe database partitioned across the N GPUs of the box:
per rank the tcgen05 bf16 pass + exact fp32 re-score (libvsm.so), then the exchange of the [nq][2]
result keys (stored straight into the peers' buffers over NVLink and merged in the same kernel; or
an NCCL all-gather + merge kernel with --exchange nccl).  Total work is fixed as N grows ("strong").

  python bench.py [--gpus N] [--steps K] [--warmup W]        one rank per GPU under torchrun for N > 1
  python bench.py --impl reference ...                        the reference's CPU matcher
                                                              (cv2.BFMatcher, the OpenCV code the
                                                              reference calls) on a bounded sample

Prints ONE JSON line (rank 0).  `value` is device-timed with inputs resident; `e2e` is the same
search through host buffers (H2D of the queries and D2H of the result inside the timed region).
`roofline` is the tensor-core kernel's own duration in the launches of the `value` loop (the
library's CUDA event pairs around it, read after the loop).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "DB query TFLOP/s (exact top-2 loop-closure search, 2*nq*nt*256 FLOP per query batch)"
NQ = 2000
TOTAL_ROWS = 20_000_000          # 10K keyframes x 2000 descriptors
KF_ROWS = 2000
N_PLANTED = 400                  # 20 % of the queries are noisy re-observations of DB rows
SIGMA = 0.05


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return (float(p["bf16_tflops_sustained"]), float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json, sustained)",
                float(p.get("bf16_tflops", p["bf16_tflops_sustained"])))        # "bf16_tflops" = best of 10 (burst)
    except Exception:
        return 1400.0, 6650.0, "fallback (B200_PROFILING.md)", 1400.0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def window(self, t0, t1):
        sm, mx, reasons = [], [], set()
        for ts, line in self.rows:
            if ts < t0 - 0.05 or ts > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}

    def stop(self):
        if self.proc:
            self.proc.terminate()


# ---- synthetic data ---------------------------------------------------------------------------
def make_queries(torch, device):
    g = torch.Generator(device=device)
    g.manual_seed(1234)
    q = torch.randn((NQ, 256), generator=g, device=device, dtype=torch.float32)
    q = q / q.norm(dim=1, keepdim=True)
    noise = torch.randn((N_PLANTED, 256), generator=g, device=device, dtype=torch.float32) * SIGMA
    return q, noise


def make_shard(torch, device, rank, world, q, noise, total_rows):
    rows = total_rows // world
    off = rank * rows
    if rank == world - 1:
        rows = total_rows - off
    db = torch.empty((rows, 256), dtype=torch.float32, device=device)
    g = torch.Generator(device=device)
    chunk = 1 << 20
    # rows are generated on a GLOBAL chunk grid (seed = chunk number), so the database -- and hence
    # the answer and its checksum -- is the same for every number of shards
    for gc in range(off // chunk, (off + rows + chunk - 1) // chunk):
        g.manual_seed(10_000 + gc)
        x = torch.randn((chunk, 256), generator=g, device=device, dtype=torch.float32)
        x = x / x.norm(dim=1, keepdim=True)
        lo, hi = max(off, gc * chunk), min(off + rows, (gc + 1) * chunk)
        db[lo - off:hi - off] = x[lo - gc * chunk:hi - gc * chunk]
    del x
    # plant: DB row r_i re-observes query i (noise renormalised), wherever r_i lives
    stride = total_rows // N_PLANTED
    planted = [i * stride + 17 for i in range(N_PLANTED)]
    for i, r in enumerate(planted):
        if off <= r < off + rows:
            v = q[i] + noise[i]
            db[r - off] = v / v.norm()
    return db, off, planted


# ---- reference arm ----------------------------------------------------------------------------
def cpu_matcher():
    """The reference's CPU path for float descriptors as BASELINE.json names it:
    cv::BFMatcher(NORM_L2).knnMatch(q, t, 2).  cv2 wraps the same OpenCV C++ code the reference
    links (src/Slam.cpp:1149); if cv2 is absent the CPU oracle port is timed instead."""
    try:
        import cv2
        cores = cv2.getNumThreads()
        bf = cv2.BFMatcher(cv2.NORM_L2)
        return (lambda q, t: bf.knnMatch(q, t, k=2)), cores, "reference", f"cv2 {cv2.__version__} BFMatcher.knnMatch k=2"
    except Exception:
        from oracle import oracle
        cores = os.cpu_count() or 1
        return (lambda q, t: oracle.knn(q, t, 2, threads=cores)), cores, "port", "oracle/vsm_oracle.c knn k=2"


def cpu_data(n_rows, seed=7):
    import numpy as np
    rng = np.random.default_rng(seed)
    q = rng.standard_normal((NQ, 256), dtype=np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    t = rng.standard_normal((n_rows, 256), dtype=np.float32)
    t /= np.linalg.norm(t, axis=1, keepdims=True)
    return q, t


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fn, cores, kind, what = cpu_matcher()
    n_rows = args.ref_rows
    q, t = cpu_data(n_rows)
    for _ in range(args.warmup):
        fn(q, t)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn(q, t)
    dt = (time.perf_counter() - t0) / args.steps
    flops = 2.0 * NQ * n_rows * 256
    val = flops / dt / 1e12
    sample = f"{what}; each step = {NQ} queries x {n_rows}-row sample of the {TOTAL_ROWS}-row DB (brute force is linear in rows)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "TFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus, args.rows),
        "cpu_baseline": {"value": val, "unit": "TFLOP/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def committed_traffic(shard_rows):
    """dram bytes per tc_top3_kernel launch from the committed `ncu --set full` capture, or None."""
    try:
        for name in ("r02_traffic.json", "r01_traffic.json"):
            with open(os.path.join(ROOT, "profiles", name)) as f:
                v = json.load(f)["tc_top3_kernel"].get(str(int(shard_rows)))
            if v is not None:
                return v
        return None
    except Exception:
        return None


def workload_config(n_gpus, total_rows):
    return {"workload": f"loop-closure search: {NQ} query descriptors x {total_rows}-descriptor keyframe DB "
                        f"({total_rows // KF_ROWS} keyframes x {KF_ROWS}), exact global top-2 per query "
                        "(BASELINE configs[3])",
            "nq": NQ, "db_rows": total_rows, "dim": 256, "sharding": f"db rows / {n_gpus} GPUs, per-rank exact top-2 then key exchange + merge",
            "l2": "inputs larger than L2 (bf16 shard >= 1.28 GB vs 126 MB L2); no flush needed"}


def parity_report(vsm_b200, device, q, t, cpu_result):
    """Indices / fp32 distance bits / ratio decisions of the CUDA path vs the CPU matcher's answer
    on the cpu_baseline sample (north star: identical except exact ties or |d0 - r*d1| <= 1e-5 r d1)."""
    import numpy as np
    nq = q.shape[0]
    ci = -np.ones((nq, 2), np.int64)
    cd = np.full((nq, 2), np.finfo(np.float32).max, np.float32)
    if len(cpu_result) == 2 and isinstance(cpu_result[0], np.ndarray):      # oracle port: (idx, dist)
        ci, cd = cpu_result[0].astype(np.int64), cpu_result[1]
    else:                                                  # cv2: list of DMatch lists
        for i, ms in enumerate(cpu_result):
            for p_, m_ in enumerate(ms):
                ci[i, p_], cd[i, p_] = m_.trainIdx, m_.distance
    with vsm_b200.Matcher(device=device, engine=vsm_b200.ENGINE_TENSOR) as m:
        gi, gd = m.knn_match(q, t)
        st = m.stats()
    same_idx = (gi == ci).all(axis=1)
    same_bits = (gd.view(np.uint32) == cd.view(np.uint32)).all(axis=1)
    rep = {"sample": f"{nq} x {t.shape[0]}", "identical_indices": int(same_idx.sum()),
           "identical_distance_bits": int(same_bits.sum()), "queries": nq,
           "candidates_rescored": st["candidates"], "rescanned_slices": st["flagged_slices"]}
    for r in (0.70, 0.75, 0.80):
        rr = np.float32(r)
        dec_g = gd[:, 0] < rr * gd[:, 1]
        dec_c = cd[:, 0] < rr * cd[:, 1]
        rep[f"ratio_{int(r * 100)}_decisions_identical"] = int((dec_g == dec_c).sum())
    bad = ~(same_idx & same_bits)
    ties = bad & (cd[:, 0] == cd[:, 1])
    rep["exempt_exact_ties"] = int(ties.sum())
    rep["unexplained_mismatches"] = int((bad & ~ties).sum())
    return rep


def oracle_parity(torch, dist, shard, off, q_host, sample, got_idx, got_dist, rank, world, device):
    """The answer of the timed search against the CPU oracle over the WHOLE database, at every N:
    each rank runs oracle.knn(q[sample], its own shard copied to the host chunk by chunk), the per-rank
    top-2 lists are gathered and rank 0 merges them with oracle.merge_top2 -- no GPU kernel of the
    product takes part -- and compares indices and fp32 distance bits (first and second neighbour)
    with what the GPU path returned for those queries."""
    import numpy as np
    from oracle import oracle
    t0 = time.perf_counter()
    qs = np.ascontiguousarray(q_host[sample])
    threads = max(1, (os.cpu_count() or 1) // world)
    S = len(sample)
    parts_i, parts_d = [], []
    chunk = 1 << 20
    for c0 in range(0, shard.shape[0], chunk):
        c1 = min(shard.shape[0], c0 + chunk)
        host = shard[c0:c1].cpu().numpy()
        oi, od = oracle.knn(qs, host, 2, threads=threads)
        parts_i.append(np.where(oi >= 0, oi + off + c0, -1))
        parts_d.append(od)
    if parts_i:
        li, ld = oracle.merge_top2(np.stack(parts_i), np.stack(parts_d))
    else:
        li = -np.ones((S, 2), np.int64)
        ld = np.full((S, 2), np.finfo(np.float32).max, np.float32)
    if world > 1:
        gi = torch.empty((world, S, 2), dtype=torch.int64, device=device)
        gd = torch.empty((world, S, 2), dtype=torch.float32, device=device)
        dist.all_gather_into_tensor(gi.view(-1, 2), torch.from_numpy(li).to(device))
        dist.all_gather_into_tensor(gd.view(-1, 2), torch.from_numpy(ld).to(device))
        gi, gd = gi.cpu().numpy(), gd.cpu().numpy()
    else:
        gi, gd = li[None], ld[None]
    if rank != 0:
        return None
    wi, wd = oracle.merge_top2(gi, gd)
    hi, hd = got_idx[sample], got_dist[sample]
    same_idx = (hi == wi).all(axis=1)
    same_bits = (hd.view(np.uint32) == wd.view(np.uint32)).all(axis=1)
    bad = ~(same_idx & same_bits)
    ties = bad & (wd[:, 0] == wd[:, 1])
    rep = {"oracle": "oracle/vsm_oracle.c knn (CPU) over every row of the database, per rank on its shard, merged by oracle.merge_top2",
           "queries": int(S), "planted_queries": int((sample < N_PLANTED).sum()),
           "identical_indices_both_neighbours": int(same_idx.sum()), "identical_distance_bits": int(same_bits.sum())}
    for r in (0.70, 0.75, 0.80):
        rr = np.float32(r)
        rep[f"ratio_{int(r * 100)}_decisions_identical"] = int(((hd[:, 0] < rr * hd[:, 1]) == (wd[:, 0] < rr * wd[:, 1])).sum())
    rep["exempt_exact_ties"] = int(ties.sum())
    rep["unexplained_mismatches"] = int((bad & ~ties).sum())
    rep["seconds"] = round(time.perf_counter() - t0, 1)
    return rep


def cpp_track_latency(vsm_b200):
    """Builds bench_cpp/track_latency.cpp against libvsm.so and runs it (2544 pairs)."""
    import tempfile
    try:
        libdir = os.path.dirname(vsm_b200.lib_path())
        exe = os.path.join(tempfile.mkdtemp(prefix="vsm_bench_"), "track_latency")
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"),
                               os.path.join(ROOT, "bench_cpp", "track_latency.cpp"), "-L", libdir, "-lvsm",
                               f"-Wl,-rpath,{libdir}", "-lpthread", "-ldl", "-o", exe], stderr=subprocess.DEVNULL)
        res = subprocess.run([exe, "2544"], capture_output=True, text=True, timeout=120)
        return json.loads(res.stdout.strip().splitlines()[-1])
    except Exception as e:            # a missing compiler only loses this informational number
        return {"unavailable": repr(e)[:200]}


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def cpu_pair_baselines():
    """The reference's CPU matchers on the pair-matching configs (BASELINE configs[0], [1], [4]), on this
    box's host: cv::BFMatcher(NORM_L2).knnMatch (the exact matcher BASELINE names) with all threads and
    with one, and cv::FlannBasedMatcher -- what the reference actually calls for float descriptors
    (src/Slam.cpp:23, src/LoopCloser.cpp:31), approximate and not reproducible run to run.
    Times are per pair (per batch for configs[4]), best of a few repetitions; mutual = both directions."""
    import numpy as np
    try:
        import cv2
    except Exception as e:
        return {"unavailable": repr(e)[:120]}
    rng = np.random.default_rng(5)

    def unit(n):
        x = rng.standard_normal((n, 256)).astype(np.float32)
        return x / np.linalg.norm(x, axis=1, keepdims=True)

    def best(fn, reps):
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        return min(ts) * 1e3

    nthreads = cv2.getNumThreads()
    out = {"cpu_model": cpu_model(), "host_threads": nthreads, "opencv": cv2.__version__,
           "flann_note": "cv::FlannBasedMatcher (default KD-tree index, rebuilt per call like the reference's 2-argument knnMatch): "
                         "approximate, shown for reference only"}
    sizes4 = np.random.default_rng(0).integers(200, 2049, size=(64, 2))
    cases = {"configs0_2000x2000_ratio": ([(2000, 2000)], False),
             "configs1_1000x1000_mutual_ratio": ([(1000, 1000)], True),
             "configs4_ragged64_mutual_ratio": ([(int(a), int(b)) for a, b in sizes4], True)}
    for name, (shapes, mutual) in cases.items():
        mats = [(unit(a), unit(b)) for a, b in shapes]

        def run(m):
            for a, b in mats:
                m.knnMatch(a, b, k=2)
                if mutual:
                    m.knnMatch(b, a, k=1)

        r = {}
        cv2.setNumThreads(nthreads)
        r["bfmatcher_all_threads_ms"] = best(lambda: run(cv2.BFMatcher(cv2.NORM_L2)), 3)
        r["flann_all_threads_ms"] = best(lambda: run(cv2.FlannBasedMatcher()), 2)
        cv2.setNumThreads(1)
        r["bfmatcher_1_thread_ms"] = best(lambda: run(cv2.BFMatcher(cv2.NORM_L2)), 2 if len(shapes) == 1 else 1)
        cv2.setNumThreads(nthreads)
        out[name] = r
    return out


# ---- extra: the pair-matching configs (rank 0, N = 1) ------------------------------------------------
def extra_pair_numbers(torch, vsm_b200, device):
    import numpy as np
    out = {}
    m = vsm_b200.Matcher(device=device, engine=vsm_b200.ENGINE_TENSOR)
    g = torch.Generator(device="cuda")
    g.manual_seed(99)

    def unit(n):
        x = torch.randn((n, 256), generator=g, device="cuda")
        return x / x.norm(dim=1, keepdim=True)

    def planted_from(prev, overlap=0.6, sigma=0.06):
        n = prev.shape[0]
        nxt = unit(n)
        k = int(overlap * n)
        src = torch.randperm(n, generator=g, device="cuda")[:k]
        dst = torch.randperm(n, generator=g, device="cuda")[:k]
        v = prev[src] + sigma * torch.randn((k, 256), generator=g, device="cuda")
        nxt[dst] = v / v.norm(dim=1, keepdim=True)
        return nxt

    # configs[1]: 2544 consecutive 1000x1000 frame pairs, mutual-NN + ratio 0.75, host buffers in and out
    npairs = 2544
    frames = torch.empty((npairs + 1, 1000, 256), dtype=torch.float32).pin_memory()
    cur = unit(1000)
    frames[0].copy_(cur)
    for f in range(1, npairs + 1):
        cur = planted_from(cur)
        frames[f].copy_(cur)
    torch.cuda.synchronize()
    fr = frames.numpy()

    def run_sequence(step):
        lat, nmatch = [], 0
        t_all = time.perf_counter()
        for f in range(npairs):
            t0 = time.perf_counter()
            nmatch += step(f)
            lat.append(time.perf_counter() - t0)
        t_all = time.perf_counter() - t_all
        lat.sort()
        return {"pairs": npairs, "p50_us": lat[len(lat) // 2] * 1e6, "p99_us": lat[int(len(lat) * 0.99)] * 1e6,
                "matches_per_s": nmatch / t_all, "pairs_per_s": npairs / t_all,
                "pair_distances_per_s": npairs * 1e6 / t_all}

    # (a) Slam::match_features as the reference calls it: both frames come from the host every call
    for f in range(20):
        m.match_features(fr[f], fr[f + 1], 0.75, mutual=True, want_raw=False)
    r = run_sequence(lambda f: len(m.match_features(fr[f], fr[f + 1], 0.75, mutual=True, want_raw=False)[0]))
    r.update({"timing": "host wall clock per call incl. H2D of BOTH frames (pinned) and D2H of the DMatch list",
              "device_ms_last": m.stats()["device_ms"], "launches_per_pair": m.stats()["kernel_launches"]})
    out["tracking_1000x1000_mutual_ratio_both_frames_from_host"] = r
    # (b) the tracking step with the reference frame resident (vsm_track): one frame uploaded per pair
    mt = vsm_b200.Matcher(device=device, engine=vsm_b200.ENGINE_TENSOR, store_rows=(npairs + 64) * 1000)
    state = {"h": mt.track(-1, 0, fr[0], ref_rows=1000)[2]}

    def step(f):
        good, _, h = mt.track(state["h"], f + 1, fr[f + 1], 0.75, mutual=True, ref_rows=1000)
        state["h"] = h
        return len(good)

    for f in range(20):
        step(f)
    mt.clear_store()
    state["h"] = mt.track(-1, 0, fr[0], ref_rows=1000)[2]
    mt.set_profiling(False)          # the per-kernel events cost a few microseconds per call
    r = run_sequence(step)
    mt.set_profiling(True)
    mt.clear_store()
    state["h"] = mt.track(-1, 0, fr[0], ref_rows=1000)[2]
    step(0)
    st = mt.stats()
    r.update({"timing": "host wall clock per call incl. H2D of the current frame (pinned) and D2H of the DMatch list",
              "device_ms_last": st["device_ms"], "tc_ms_last": st["tc_ms"], "select_ms_last": st["select_ms"],
              "launches_per_pair": st["kernel_launches"]})
    out["tracking_1000x1000_mutual_ratio"] = r
    mt.close()
    # (c) the same step timed from C++ (the reference's host language) through the C ABI
    out["tracking_1000x1000_mutual_ratio_cpp"] = cpp_track_latency(vsm_b200)
    # configs[0]: 2000 x 2000, k=2 + ratio 0.8
    a = unit(2000)
    b = planted_from(a, 0.6, 0.08)
    ha, hb = a.cpu().pin_memory().numpy(), b.cpu().pin_memory().numpy()     # pinned: DMA without a staging copy
    for _ in range(5):
        m.match_features(ha, hb, 0.8, want_raw=False)
    ts = []
    for _ in range(50):
        t0 = time.perf_counter()
        good, _ = m.match_features(ha, hb, 0.8, want_raw=False)
        ts.append(time.perf_counter() - t0)
    ts.sort()
    out["pair_2000x2000_ratio08"] = {"p50_us": ts[len(ts) // 2] * 1e6, "matches": int(len(good)),
                                     "device_ms": m.stats()["device_ms"], "tc_ms": m.stats()["tc_ms"]}
    # configs[4]: 64 ragged pairs, sizes U{200..2048}, mutual + ratio
    rng = np.random.default_rng(0)
    sizes = rng.integers(200, 2049, size=(64, 2))
    qs, tsets = [], []
    for nq, nt in sizes:
        base = unit(int(max(nq, nt)))
        nxt = planted_from(base, 0.6, 0.08)
        qs.append(base[:nq].cpu().numpy())
        tsets.append(nxt[:nt].cpu().numpy())
    q_off = np.zeros(65, np.int32); t_off = np.zeros(65, np.int32)
    q_off[1:] = np.cumsum(sizes[:, 0]); t_off[1:] = np.cumsum(sizes[:, 1])
    qa_t = torch.from_numpy(np.concatenate(qs)).pin_memory()
    ta_t = torch.from_numpy(np.concatenate(tsets)).pin_memory()
    qa, ta = qa_t.numpy(), ta_t.numpy()
    for _ in range(3):
        m.match_batch_packed(qa, q_off, ta, t_off, 0.75, True)
    tb = []
    for _ in range(10):
        t0 = time.perf_counter()
        res = m.match_batch_packed(qa, q_off, ta, t_off, 0.75, True)
        tb.append(time.perf_counter() - t0)
    tb.sort()
    flops = float(sum(2.0 * a_ * b_ * 256 for a_, b_ in sizes))
    st = m.stats()
    out["ragged_batch_64"] = {"p50_ms": tb[len(tb) // 2] * 1e3, "useful_gflop": flops / 1e9,
                              "matches": int(sum(len(r) for r in res)), "device_ms": st["device_ms"],
                              "h2d_mbytes": (qa.nbytes + ta.nbytes) / 1e6, "select_ms": st["select_ms"],
                              "tc_ms": st["tc_ms"], "tc_useful_tflops": flops / (st["tc_ms"] * 1e-3) / 1e12 if st["tc_ms"] else None}
    # the same 64 pairs with every frame already resident in the keyframe store (nothing uploaded):
    # what the batch costs inside a SLAM process that keeps its keyframes on the device
    m.clear_store()
    qh = [m.add_keyframe(2 * p, qs[p]) for p in range(64)]
    th = [m.add_keyframe(2 * p + 1, tsets[p]) for p in range(64)]
    cap = int(sizes[:, 0].sum())
    for _ in range(3):
        m.match_batch_stored(qh, th, 0.75, True, capacity=cap)
    tb = []
    for _ in range(20):
        t0 = time.perf_counter()
        res2 = m.match_batch_stored(qh, th, 0.75, True, capacity=cap)
        tb.append(time.perf_counter() - t0)
    tb.sort()
    st = m.stats()
    out["ragged_batch_64_resident"] = {"p50_ms": tb[len(tb) // 2] * 1e3, "useful_gflop": flops / 1e9,
                                       "matches": int(sum(len(r) for r in res2)), "same_matches_as_from_host": bool(
                                           all(a.tobytes() == b.tobytes() for a, b in zip(res, res2))),
                                       "device_ms": st["device_ms"], "tc_ms": st["tc_ms"], "select_ms": st["select_ms"],
                                       "useful_tflops_e2e": flops / tb[len(tb) // 2] / 1e12,
                                       # `useful_gflop` counts the forward direction only (BASELINE's 2*nq*nt*256 per pair); the
                                       # mutual test needs the reverse contraction as well, and the kernel computes both
                                       "tc_tflops_both_directions": 2 * flops / (st["tc_ms"] * 1e-3) / 1e12 if st["tc_ms"] else None}
    m.clear_store()
    # A/B of the pair-matching record kinds on the same 64 resident pairs: the threshold-driven top-4 records
    # (vsm_opts.reserved[5] = 1) against the tile top-2 records every number above was taken with
    m4 = vsm_b200.Matcher(device=device, engine=vsm_b200.ENGINE_TENSOR, tile_top2=False)
    try:
        qh4 = [m4.add_keyframe(2 * p, qs[p]) for p in range(64)]
        th4 = [m4.add_keyframe(2 * p + 1, tsets[p]) for p in range(64)]
        for _ in range(3):
            res4 = m4.match_batch_stored(qh4, th4, 0.75, True, capacity=cap)
        st4 = m4.stats()
        out["ragged_batch_64_resident_top4_records"] = {
            "device_ms": st4["device_ms"], "tc_ms": st4["tc_ms"], "select_ms": st4["select_ms"],
            "exact_distances": st4["candidates"], "same_matches": bool(all(a.tobytes() == b.tobytes() for a, b in zip(res2, res4))),
            "note": "tile top-2 records (default): device_ms / tc_ms / select_ms in ragged_batch_64_resident"}
        out["ragged_batch_64_resident"]["exact_distances"] = st["candidates"]
    finally:
        m4.close()
    # configs[2]: 1000 queries vs a 500-keyframe database (500K rows), both forms the reference uses:
    # the stacked global top-2 (src/Slam.cpp:546-574) and LoopCloser::detect's per-keyframe kNN +
    # ratio test (src/LoopCloser.cpp:43-62); queries come from pinned host memory, results go back
    nkf, rows_kf, nq2 = 500, 1000, 1000
    db = torch.empty((nkf * rows_kf, 256), device="cuda")
    for k0 in range(0, nkf * rows_kf, 100000):
        db[k0:k0 + 100000] = unit(100000)
    q2 = unit(nq2)
    src = 123 * rows_kf + torch.randperm(rows_kf, generator=g, device="cuda")[:200]      # re-observations of keyframe 123
    v = db[src] + 0.06 * torch.randn((200, 256), generator=g, device="cuda")
    q2[:200] = v / v.norm(dim=1, keepdim=True)
    hq2 = q2.cpu().pin_memory().numpy()
    torch.cuda.synchronize()
    m.adopt_device_matrix(db.data_ptr(), nkf * rows_kf, np.arange(nkf + 1, dtype=np.int64) * rows_kf)
    fl = 2.0 * nq2 * nkf * rows_kf * 256
    for name, call in (("loop_closure_500kf_global_top2", lambda: m.search_map_points(hq2)),
                       ("loop_closure_500kf_per_keyframe_ratio", lambda: m.detect_candidates(hq2, 0.75, want_matches=False))):
        for _ in range(3):
            call()
        tl = []
        for _ in range(20):
            t0 = time.perf_counter()
            res = call()
            tl.append(time.perf_counter() - t0)
        tl.sort()
        st = m.stats()
        out[name] = {"p50_ms": tl[len(tl) // 2] * 1e3, "device_ms": st["device_ms"], "tc_ms": st["tc_ms"],
                     "select_ms": st["select_ms"], "tflops_e2e": fl / tl[len(tl) // 2] / 1e12,
                     "tc_tflops": fl / (st["tc_ms"] * 1e-3) / 1e12 if st["tc_ms"] else None,
                     "candidates_rescored": st["candidates"], "rescanned_slices": st["flagged_slices"]}
    out["loop_closure_500kf_per_keyframe_ratio"]["keyframes_with_30_or_more_survivors"] = int((res[0] >= 30).sum())
    out["loop_closure_500kf_per_keyframe_ratio"]["api"] = "vsm_db_segmented (record-based, returns counts for every keyframe)"
    # the same search in LoopCloser::detect's own shape: gate (>= 30 survivors) and packing on the device,
    # only the surviving keyframes' lists come back (vsm_loop_detect_compact)
    call = lambda: m.loop_detect_compact(10**6, hq2, 0.75, min_gap=0, every=1, min_matches=30)
    for _ in range(3):
        call()
    tl = []
    for _ in range(30):
        t0 = time.perf_counter()
        res = call()
        tl.append(time.perf_counter() - t0)
    tl.sort()
    st = m.stats()
    out["loop_closure_500kf_compact"] = {"p50_ms": tl[len(tl) // 2] * 1e3, "device_ms": st["device_ms"], "tc_ms": st["tc_ms"],
                                         "after_tc_ms": st["select_ms"], "tflops_e2e": fl / tl[len(tl) // 2] / 1e12,
                                         "tc_tflops": fl / (st["tc_ms"] * 1e-3) / 1e12 if st["tc_ms"] else None,
                                         "launches": st["kernel_launches"], "candidate_keyframes": sorted(res[1]),
                                         "survivors_returned": int(sum(len(v) for v in res[1].values())),
                                         "api": "vsm_loop_detect_compact (fused dismissal, O(open pairs) after the tensor pass)"}
    m.clear_store()
    del db
    # BASELINE configs[3] in LoopCloser's form: 10K keyframes x 2000 descriptors, 2000 queries, the reference's
    # eligibility (gap >= 200 frames, every 5th keyframe): 2000 keyframes matched per search
    try:
        nkf3, rows3, nq3 = 10_000, 2000, 2000
        db3 = torch.empty((nkf3 * rows3, 256), device="cuda")
        for k0 in range(0, nkf3 * rows3, 500_000):
            db3[k0:k0 + 500_000] = unit(500_000)
        q3 = unit(nq3)
        src = 5004 * rows3 + torch.randperm(rows3, generator=g, device="cuda")[:400]   # keyframe 5004 is every-5th eligible
        v = db3[src] + 0.06 * torch.randn((400, 256), generator=g, device="cuda")
        q3[:400] = v / v.norm(dim=1, keepdim=True)
        hq3 = q3.cpu().pin_memory().numpy()
        torch.cuda.synchronize()
        m.adopt_device_matrix(db3.data_ptr(), nkf3 * rows3, np.arange(nkf3 + 1, dtype=np.int64) * rows3)
        m.set_frame_ids(np.arange(nkf3, dtype=np.int32))
        call = lambda: m.loop_detect_compact(nkf3 + 500, hq3, 0.75, min_gap=200, every=5, min_matches=30)
        for _ in range(3):
            call()
        tl = []
        for _ in range(10):
            t0 = time.perf_counter()
            res = call()
            tl.append(time.perf_counter() - t0)
        tl.sort()
        st = m.stats()
        matched = int((res[0] >= 0).sum())
        fl3 = 2.0 * nq3 * matched * rows3 * 256
        out["loop_closure_10k_kf_every5_compact"] = {
            "keyframes": nkf3, "keyframes_matched": matched, "p50_ms": tl[len(tl) // 2] * 1e3, "device_ms": st["device_ms"],
            "tc_ms": st["tc_ms"], "after_tc_ms": st["select_ms"], "tflops_e2e": fl3 / tl[len(tl) // 2] / 1e12,
            "tc_tflops": fl3 / (st["tc_ms"] * 1e-3) / 1e12 if st["tc_ms"] else None,
            "candidate_keyframes": sorted(res[1]), "survivors_returned": int(sum(len(v) for v in res[1].values()))}
        m.clear_store()
        del db3
    except Exception as e:                       # informational line: never lose the bench to it
        out["loop_closure_10k_kf_every5_compact"] = {"error": repr(e)[:300]}
    m.clear_store()
    # Slam::track_local_map (src/Slam.cpp:380-469): 3000 map points projected into a frame of 1000 keypoints,
    # 12 px window search, sequential assignment -- descriptors from the host per call, and from the
    # resident map-point table (only keypoints, positions and the pose travel)
    try:
        rngt = np.random.default_rng(3)
        nmp, nkp = 3000, 1000
        a = 0.05
        R = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
        tvec = np.array([0.1, -0.05, 0.2])
        pos = np.stack([rngt.uniform(-6, 6, nmp), rngt.uniform(-4, 4, nmp), rngt.uniform(1, 12, nmp)], axis=1)
        mp_desc = unit(nmp).cpu().numpy()
        cam = (R @ pos.T).T + tvec
        u_ = 525.0 * cam[:, 0] / cam[:, 2] + 319.5
        v_ = 525.0 * cam[:, 1] / cam[:, 2] + 239.5
        vis = np.nonzero((u_ >= 0) & (u_ < 640) & (v_ >= 0) & (v_ < 480))[0]
        pick = rngt.permutation(vis)[:700]
        kp = np.stack([rngt.uniform(0, 640, nkp), rngt.uniform(0, 480, nkp)], axis=1).astype(np.float32)
        kp[:len(pick), 0] = u_[pick] + rngt.normal(0, 3.0, len(pick))
        kp[:len(pick), 1] = v_[pick] + rngt.normal(0, 3.0, len(pick))
        kp = np.clip(kp, 0, [639.5, 479.5]).astype(np.float32)
        desc = unit(nkp).cpu().numpy()
        noisy = mp_desc[pick] + 0.02 * rngt.standard_normal((len(pick), 256)).astype(np.float32)
        desc[:len(pick)] = noisy / np.linalg.norm(noisy, axis=1, keepdims=True)
        valid = np.ones(nmp, np.uint8)
        res_t = {}
        m.add_map_points(mp_desc, 0)
        for name, md, vd in (("descriptors_from_host", mp_desc, valid), ("resident_point_table", None, None)):
            tl = []
            for it in range(23):
                ind = -np.ones(nkp, np.int32)
                t0 = time.perf_counter()
                tracked, _, _, _ = m.track_local_map(kp, desc, pos, md, vd, R, tvec, ind)
                tl.append(time.perf_counter() - t0)
            tl = sorted(tl[3:])
            res_t[name] = {"p50_us": tl[len(tl) // 2] * 1e6, "device_ms": m.stats()["device_ms"], "tracked": int(tracked)}
        m.clear_map_points()
        out["track_local_map_3000_points_x_1000_keypoints"] = res_t
    except Exception as e:
        out["track_local_map_3000_points_x_1000_keypoints"] = {"error": repr(e)[:300]}
    m.close()
    return out


def sharded_configs2(torch, dist, db, rank, world, device, barrier, steps=50):
    """BASELINE configs[2] at this N: 1000 query descriptors against a 500-keyframe database (500K rows)
    partitioned like the headline database; device-timed with resident queries, and end to end through
    the host-buffer C-ABI call.  Max over ranks."""
    import numpy as np
    nq, rows_total = 1000, 500_000
    g = torch.Generator(device=device)
    g.manual_seed(4321)
    full = torch.randn((rows_total, 256), generator=g, device=device)
    full = full / full.norm(dim=1, keepdim=True)
    q = torch.randn((nq, 256), generator=g, device=device)
    q = q / q.norm(dim=1, keepdim=True)
    v = full[123_000:123_200] + 0.06 * torch.randn((200, 256), generator=g, device=device)
    q[:200] = v / v.norm(dim=1, keepdim=True)
    rows = rows_total // world
    off = rank * rows
    if rank == world - 1:
        rows = rows_total - off
    shard = full[off:off + rows].contiguous()
    del full
    db.adopt(shard, off, None)
    hq = q.cpu().pin_memory().numpy()
    for _ in range(5):
        db.search_device(q)
        db.search_host_abi(hq)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(db.stream):
        e0.record()
    for _ in range(steps):
        db.search_device(q)
    with torch.cuda.stream(db.stream):
        e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / steps
    tc_ms = float(np.mean(db.matcher.tc_history(min(steps, 64))))
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        hi, hd = db.search_host_abi(hq)
    barrier()
    ms_e2e = (time.perf_counter() - t0) * 1e3 / steps
    if world > 1:
        t = torch.tensor([ms, ms_e2e, tc_ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e, tc_ms = (float(x) for x in t)
    fl = 2.0 * nq * rows_total * 256
    return {"workload": "1000 queries x 500K-row keyframe DB, exact global top-2, sharded over the GPUs (BASELINE configs[2])",
            "ms_per_search_device": ms, "tflops_device": fl / (ms * 1e-3) / 1e12,
            "ms_per_search_e2e_host_buffers": ms_e2e, "tflops_e2e": fl / (ms_e2e * 1e-3) / 1e12,
            "tc_kernel_ms_max_rank": tc_ms, "planted_recovered": int(((hi[:200, 0] >= 123_000) & (hi[:200, 0] < 123_200)).sum()),
            "note": "e2e timed by host wall clock between barriers (synchronous calls), max over ranks"}


# ---- main arm -----------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    import vsm_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    device = torch.device("cuda", local)
    total_rows = args.rows

    sh = vsm_b200.load_sharded()
    db = sh.ShardedDB(local, rank, world, exchange=args.exchange, engine=args.engine)
    q, noise = make_queries(torch, device)
    shard, off, planted = make_shard(torch, device, rank, world, q, noise, total_rows)
    seg = None   # one segment per shard for the global search
    db.adopt(shard, off, seg)
    hq = q.cpu().pin_memory()
    hq_np = hq.numpy()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up (also sizes every buffer and the NCCL communicator)
    for _ in range(max(args.warmup, 3)):
        db.search_device(q)
        db.search_host_abi(hq_np)
    db.stream.synchronize()
    launches_per_step = db.launches_per_search()

    sampler = ClockSampler(local) if rank == 0 else None
    # -- device-resident timing (no host synchronisation inside the timed region) ------------------
    # both timed loops start from the same state: the GPU sits at its power cap, so a loop that starts
    # hot (right after the database generation, or right after the other loop) runs at a lower clock --
    # idle for a second, then the warm-up steps of the loop's own kind
    barrier()
    time.sleep(1.0)
    for _ in range(max(args.warmup, 3)):
        db.search_device(q)
    db.stream.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    with torch.cuda.stream(db.stream):
        e0.record()
    for _ in range(args.steps):
        oi, od = db.search_device(q)
    with torch.cuda.stream(db.stream):
        e1.record()
    barrier()
    t_wall1 = time.time()
    ms = e0.elapsed_time(e1)
    # -- the dominant kernel inside that same timed loop: the library keeps a CUDA event pair around
    #    tc_top3_kernel for each of its last 64 calls (recorded on db.stream, read only now)
    tc_ms = [float(x) for x in db.matcher.tc_history(min(args.steps, 64))]
    # -- end to end through host buffers -------------------------------------------------------------
    # same starting conditions as the loop above
    barrier()
    time.sleep(1.0)
    for _ in range(max(args.warmup, 3)):
        db.search_host_abi(hq_np)
    db.stream.synchronize()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(db.stream):
        f0.record()
    e2e_dev, e2e_tc, e2e_wall = [], [], []
    for _ in range(args.steps):
        tw = time.perf_counter()
        hi, hd = db.search_host_abi(hq_np)             # ONE synchronous C-ABI call per rank: host queries in, host top-2 out
        e2e_wall.append((time.perf_counter() - tw) * 1e3)
        st = db.matcher.stats()                        # stream already idle: reads the call's own events
        e2e_dev.append(st["device_ms"])
        e2e_tc.append(st["tc_ms"])
    with torch.cuda.stream(db.stream):
        f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    if world > 1:
        t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
    t_wall2 = time.time()
    clocks = None
    if sampler:
        clocks = sampler.window(t_wall0, t_wall1)
        if clocks["samples"] < 2:        # short timed region: widen to the end-to-end loop (same load)
            clocks = sampler.window(t_wall0, t_wall2)
            clocks["note"] = "window widened to the device-timed + end-to-end loops"
    if sampler:
        sampler.stop()

    # parity of the timed search's answer against the CPU oracle over the whole database (all ranks take part)
    import numpy as _np
    parity = None
    if args.parity_queries > 0:
        half = max(1, args.parity_queries // 2)
        sample = _np.unique(_np.concatenate([_np.arange(0, N_PLANTED, max(1, N_PLANTED // half)),
                                             _np.arange(N_PLANTED, NQ, max(1, (NQ - N_PLANTED) // half))]))
        parity = oracle_parity(torch, dist, shard, off, hq_np, sample, hi, hd, rank, world, device)
    shard_rows_main = shard.shape[0]
    cfg2 = None
    if not args.no_extra:
        # BASELINE configs[2] (1000 queries x 500 keyframes x 1000 descriptors) sharded the same way, at this N
        del shard
        torch.cuda.empty_cache()
        cfg2 = sharded_configs2(torch, dist, db, rank, world, device, barrier)

    flops = 2.0 * NQ * total_rows * 256
    per_step = ms / args.steps
    per_step_e2e = ms_e2e / args.steps
    if rank == 0:
        import numpy as np
        peak_tf, peak_hbm, peak_src, peak_burst = peaks()
        # sanity: every planted query must find its DB row as the nearest neighbour
        got = hi[:N_PLANTED, 0]
        recovered = int((got == np.array(planted)).sum())
        # the same seeded database gives the same answer for every N: compare this across runs
        chk = (int((hi.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)).sum() & np.uint64(0xFFFFFFFFFFFFFFFF))
               ^ int(hd.view(np.uint32).astype(np.uint64).sum()))
        st = db.matcher.stats()
        tc_avg = sum(tc_ms) / len(tc_ms)
        shard_flops = 2.0 * NQ * shard_rows_main * 256
        achieved = shard_flops / (tc_avg * 1e-3) / 1e12
        # the sustained peak is what a long power-capped step can reach (N = 1: 16 ms steps at ~1.4 GHz);
        # a short shard step that runs at boost clocks is compared with the burst figure instead
        unthrottled = bool(clocks and clocks.get("sm_mhz") and clocks["sm_mhz"] >= 0.95 * clocks["sm_max_mhz"])
        if (unthrottled or achieved > peak_tf) and peak_burst > peak_tf:
            peak_tf, peak_src = peak_burst, "measured (MEASURED_PEAKS.json, burst: the kernel ran above the sustained-clock regime)"
        line = {
            "metric": METRIC, "value": flops / (per_step * 1e-3) / 1e12, "unit": "TFLOP/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(world, total_rows),
            "parity": parity,
            "check": {"planted_recovered": f"{recovered}/{N_PLANTED}", "result_checksum": f"{chk:016x}", "candidates_rescored": st["candidates"],
                      "flagged_slices": st["flagged_slices"], "select_ms": st["select_ms"]},
            "clocks": clocks,
            "e2e": {"value": flops / (per_step_e2e * 1e-3) / 1e12, "unit": "TFLOP/s", "ms_per_step": per_step_e2e,
                    "breakdown_ms_rank0_median": {"host_wall": float(np.median(e2e_wall)), "library_call_on_device": float(np.median(e2e_dev)),
                                                  "tc_top3_kernel": float(np.median(e2e_tc))},
                    "h2d_bytes_per_step": NQ * 1024, "d2h_bytes_per_step": NQ * 2 * 12,
                    "api": ("vsm_db_top2 (C ABI: host queries in, host top-2 out, synchronous)" if world == 1 else
                            ("vsm_db_top2_xchg (C ABI, one call per rank: host queries in, merged host top-2 out, fused peer-memory exchange)"
                             if db.exchange == "p2p" else "ShardedDB.search_host (torch H2D/D2H + NCCL all-gather + merge kernel)")),
                    "conditions": "each timed loop (this one and the `value` one) is preceded by 1 s idle + its own warm-up steps"},
            "gpu_launches": launches_per_step * args.steps, "exchange": db.exchange, "exchange_note": db.exchange_note, "engine": args.engine,
            "roofline": {"bound": "tensor", "kernel": "tc_top3_kernel", "achieved": achieved, "peak": peak_tf,
                         "unit": "TFLOP/s", "frac": achieved / peak_tf,
                         "traffic": args.traffic if args.traffic is not None else committed_traffic(shard_rows_main),
                         "peak_source": peak_src, "kernel_ms": tc_avg,
                         "algorithmic": "2*nq*shard_rows*256 FLOP per launch",
                         "hbm_gbs": (shard_rows_main * 512 + NQ * 512) / (tc_avg * 1e-3) / 1e9, "hbm_peak": peak_hbm},
            "matches_per_s": NQ / (per_step * 1e-3),
            "configs2_sharded": cfg2,
        }
        if world == 1 and not args.no_cpu:
            fn, cores, kind, what = cpu_matcher()
            n = args.cpu_rows
            cq, ct = cpu_data(n)
            fn(cq[:200], ct[:2000])
            t0 = time.perf_counter()
            reps = 0
            while True:
                fn(cq, ct)
                reps += 1
                if time.perf_counter() - t0 > 10.0 or reps >= 20:
                    break
            dt = (time.perf_counter() - t0) / reps
            # parity on the same sample: the CUDA path against what the CPU matcher just returned
            line["parity_vs_cpu_matcher_sample"] = parity_report(vsm_b200, local, cq, ct, fn(cq, ct))
            line["cpu_baseline"] = {"value": 2.0 * NQ * n * 256 / dt / 1e12, "unit": "TFLOP/s", "cores": cores,
                                    "kind": kind, "sample": f"{what}; {NQ} queries x {n}-row sample of the DB, {reps} repetitions"}
        if world == 1 and not args.no_extra:
            # free the big shard first: the pair configs need little memory
            try:
                line["extra"] = extra_pair_numbers(torch, vsm_b200, local)
            except Exception as e:                       # informational numbers: the headline line is printed regardless
                line["extra"] = {"error": repr(e)[:400]}
            if not args.no_cpu:
                line["extra"]["cpu_pair_baselines"] = cpu_pair_baselines()
        print(json.dumps(line))
    db.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="vsm", choices=["vsm", "reference"])
    ap.add_argument("--rows", type=int, default=TOTAL_ROWS, help="database rows in total (default: configs[3], 20M)")
    ap.add_argument("--ref-rows", type=int, default=100_000, help="reference arm: DB sample rows per step")
    ap.add_argument("--cpu-rows", type=int, default=100_000, help="cpu_baseline: DB sample rows")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: fused peer-memory exchange (default) or NCCL all-gather + merge")
    ap.add_argument("--engine", type=int, default=0, help="0 auto, 1 tensor cores (1 CTA/SM), 3 tensor cores on CTA pairs")
    ap.add_argument("--parity-queries", type=int, default=64,
                    help="queries whose answer is checked against the CPU oracle over the whole database (0 = skip)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--traffic", type=float, default=None,
                    help="dram bytes per launch (default: the committed ncu capture in profiles/, if one matches)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
