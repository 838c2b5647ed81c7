// vsm_kernels.cuh -- the CUDA-core kernels around the tensor-core pass:
//   convert_kernel  fp32 rows -> bf16 shadow + squared norms (+ min/max norm of the set)
//   prologue_kernel first kernel of a matching call: aux zeroing + descriptor upload + conversions
//   select_kernel   per query: threshold the approximate records, re-score the survivors
//                   with the canonical fp32 distance, keep the exact top-2
//   filter_kernel   Slam::match_features' loop (src/Slam.cpp:1151-1158) + mutual-NN,
//                   order-preserving compaction into DMatch lists
//   merge_kernel    per-shard top-2 lists -> global top-2 by (distance, index)
#pragma once

#include "vsm_common.cuh"

namespace vsm {

// One warp per row: 8 floats per lane.
__global__ void __launch_bounds__(256)
convert_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, float* __restrict__ n2,
               int64_t nrows, uint32_t* __restrict__ stats) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float lo = INFINITY, hi = 0.f;
    for (int64_t row = warp0; row < nrows; row += nwarps) {
        const float4* p = reinterpret_cast<const float4*>(src + row * VSM_DIM) + lane * 2;
        float4 a = __ldg(p), b = __ldg(p + 1);
        float s = a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w + b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        __nv_bfloat162 o0 = __floats2bfloat162_rn(a.x, a.y), o1 = __floats2bfloat162_rn(a.z, a.w);
        __nv_bfloat162 o2 = __floats2bfloat162_rn(b.x, b.y), o3 = __floats2bfloat162_rn(b.z, b.w);
        uint4 packed;
        packed.x = *reinterpret_cast<uint32_t*>(&o0);
        packed.y = *reinterpret_cast<uint32_t*>(&o1);
        packed.z = *reinterpret_cast<uint32_t*>(&o2);
        packed.w = *reinterpret_cast<uint32_t*>(&o3);
        reinterpret_cast<uint4*>(dst + row * VSM_DIM)[lane] = packed;
        if (lane == 0) n2[row] = s;
        lo = fminf(lo, s);
        hi = fmaxf(hi, s);
    }
    if (lane == 0 && lo <= hi) {
        atomicMax(stats, ~__float_as_uint(lo));            // see stats_read
        atomicMax(stats + 1, __float_as_uint(hi));
    }
}

// First kernel of every matching call.  One launch instead of a memset, a descriptor upload and
// one conversion per input:
//   * zeroes the per-call aux block (counters, unit queue head, hints, result keys) -- except the
//     scratch-statistics slot this call accumulates into (zeroed by the previous call, see
//     run_problems: the two slots alternate);
//   * copies the call's descriptor block from pinned host memory (read over PCIe);
//   * converts up to MAX_CONV row sets to bf16 + squared norms.  A set whose source is pinned HOST
//     memory (mapped into the device) is read straight over PCIe and its fp32 master copy is
//     written as well: upload + convert of a tracking frame without a DMA in front.
struct ConvJob {
    const float* src;          // fp32 rows: device memory, or pinned host memory mapped into the device
    float* dst_f32;            // fp32 master copy to write (zero-copy upload) or nullptr (src IS the master)
    __nv_bfloat16* dst_b16;
    float* n2;
    uint32_t* stats;           // norm statistics of the set the rows belong to (see stats_read)
    int64_t rows;
};
constexpr int MAX_CONV = 2;
struct Prologue {
    uint4* aux;
    const uint4* desc_src;
    uint4* desc_dst;
    uint32_t aux_vecs, desc_vecs;      // 16-byte vectors
    int32_t keep_slot, nconv;
    ConvJob conv[MAX_CONV];
};

__global__ void __launch_bounds__(256)
prologue_kernel(const Prologue pr) {
    pdl_launch_dependents();
    pdl_wait();
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (uint32_t v = tid; v < pr.aux_vecs; v += nth) {
        // vector 2 = bytes 32..47 = the two statistics slots of 8 bytes each
        if (v == 2) reinterpret_cast<uint2*>(pr.aux)[4 + (1 - pr.keep_slot)] = make_uint2(0u, 0u);
        else pr.aux[v] = make_uint4(0u, 0u, 0u, 0u);
    }
    for (uint32_t v = tid; v < pr.desc_vecs; v += nth) pr.desc_dst[v] = pr.desc_src[v];

    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)(tid >> 5), nwarps = (int64_t)(nth >> 5);
    for (int c = 0; c < pr.nconv; c++) {
        const ConvJob& J = pr.conv[c];
        float lo = INFINITY, hi = 0.f;
        for (int64_t row = warp0; row < J.rows; row += nwarps) {
            const float4* p = reinterpret_cast<const float4*>(J.src + row * VSM_DIM) + lane * 2;
            const float4 a = p[0], b = p[1];
            if (J.dst_f32) {
                float4* o = reinterpret_cast<float4*>(J.dst_f32 + row * VSM_DIM) + lane * 2;
                o[0] = a; o[1] = b;
            }
            float s = a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w + b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
#pragma unroll
            for (int o2 = 16; o2 > 0; o2 >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o2);
            __nv_bfloat162 o0 = __floats2bfloat162_rn(a.x, a.y), o1 = __floats2bfloat162_rn(a.z, a.w);
            __nv_bfloat162 o2b = __floats2bfloat162_rn(b.x, b.y), o3 = __floats2bfloat162_rn(b.z, b.w);
            uint4 packed;
            packed.x = *reinterpret_cast<uint32_t*>(&o0);
            packed.y = *reinterpret_cast<uint32_t*>(&o1);
            packed.z = *reinterpret_cast<uint32_t*>(&o2b);
            packed.w = *reinterpret_cast<uint32_t*>(&o3);
            reinterpret_cast<uint4*>(J.dst_b16 + row * VSM_DIM)[lane] = packed;
            if (lane == 0) J.n2[row] = s;
            lo = fminf(lo, s);
            hi = fmaxf(hi, s);
        }
        if (lane == 0 && lo <= hi) {
            atomicMax(J.stats, ~__float_as_uint(lo));          // see stats_read
            atomicMax(J.stats + 1, __float_as_uint(hi));
        }
    }
}

// ---- select / exact re-score ---------------------------------------------------
constexpr int SELECT_WARPS = 4;

struct Best2 {
    float d0, d1;
    int32_t i0, i1;
};

// Result slots are 64-bit keys ~((distance bits << 32) | index): for non-negative distances
// the integer order of the un-inverted key is the (distance, index) order, so "keep the two
// best" is a two-step atomicMax cascade on zero-initialised memory (0 = empty) -- no lock.
__device__ __forceinline__ unsigned long long result_key(float d, int32_t i) {
    return ~(((unsigned long long)__float_as_uint(d) << 32) | (uint32_t)i);
}
__device__ __forceinline__ void key_decode(unsigned long long k, int32_t& i, float& d) {
    if (k == 0ull) { i = -1; d = FLT_MAX; return; }
    k = ~k;
    i = (int32_t)(uint32_t)k;
    d = __uint_as_float((uint32_t)(k >> 32));
}

// Push a partial top-2 (k0 >= k1, k1 may be 0 = none) into a query's two slots: the larger of
// (old slot 0, k0) stays in slot 0, the loser competes with k1 (which can only ever be second)
// for slot 1.  Every key is either kept or re-offered one slot down, so concurrent pushes from
// several blocks leave the two largest keys -- no lock.
__device__ __forceinline__ void key_push2(unsigned long long* slot2, unsigned long long k0, unsigned long long k1) {
    const unsigned long long old = atomicMax(slot2, k0);
    const unsigned long long loser = old < k0 ? old : k0;
    const unsigned long long second = loser > k1 ? loser : k1;
    if (second != 0ull) atomicMax(slot2 + 1, second);
}

__device__ __forceinline__ int32_t slice_row(const SliceInfo& si, int r) {
    // r-th row of the slice as an offset into the slice's range, or -1 past the end
    int off = si.half < 0 ? r : (r / HALF_N) * TILE_N + si.half * HALF_N + (r % HALF_N);
    return off < si.t_count ? off : -1;
}
__device__ __forceinline__ int slice_span(const SliceInfo& si) {
    return si.half < 0 ? si.t_count : ((si.t_count + TILE_N - 1) / TILE_N) * HALF_N;
}

// Both half-warps score one train row each (j0 for lanes 0-15, j1 for 16-31; <0 = idle).
__device__ __forceinline__ void score_pair(const float (&qreg)[16], const float* __restrict__ t_f32,
                                           int32_t j, int l16, Best2& b) {
    const bool act = j >= 0;
    float d2 = canon_l2sqr_halfwarp(qreg, t_f32 + (size_t)(act ? j : 0) * VSM_DIM, l16);
    if (act && l16 == 0) insert2(__fsqrt_rn(d2), j, b.d0, b.i0, b.d1, b.i1);
}

// Each half-warp scores U train rows (j[u] < 0 = idle; j uniform over the half-warp): the loads of
// all U rows are issued before the first distance is reduced, so one round costs one memory
// latency instead of U.  Measured (B200): in select_kernel U = 1 (64 registers, 32 warps per SM)
// beats U = 2 (78) and U = 4 (127) by 6 % / 50 % when there are many queries (500 keyframes x 1000
// queries) and ties with them on a single tracking step, where a query has ~3 survivors -- so
// SELECT_U = 1; rescan_kernel, whose items are 128 rows long, uses RESCAN_U = 4.
constexpr int SELECT_U = 1;
constexpr int RESCAN_U = 4;
template <int U>
__device__ __forceinline__ void score_n(const float (&qreg)[16], const float* __restrict__ t_f32,
                                        const int32_t (&j)[U], int l16, Best2& b) {
    float tv[U][16];
#pragma unroll
    for (int u = 0; u < U; u++) {
        const float* p = t_f32 + (size_t)(j[u] >= 0 ? j[u] : 0) * VSM_DIM;
#pragma unroll
        for (int i = 0; i < 16; i++) tv[u][i] = __ldg(p + 16 * i + l16);
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
        const float d2 = canon_l2sqr_halfwarp_regs(qreg, tv[u]);
        if (l16 == 0 && j[u] >= 0) insert2(__fsqrt_rn(d2), j[u], b.d0, b.i0, b.d1, b.i1);
    }
}

// Exact scan of a whole slice by one warp, 2U rows per round.
template <int U>
__device__ __forceinline__ void scan_slice(const float (&qreg)[16], const float* __restrict__ t_f32,
                                           const SliceInfo& si, int h, int l16, Best2& b) {
    const int span = slice_span(si);
    for (int r0 = 0; r0 < span; r0 += 2 * U) {
        int32_t j[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int r = r0 + h * U + u;
            const int off = r < span ? slice_row(si, r) : -1;
            j[u] = off >= 0 ? si.t_index0 + off : -1;
        }
        score_n<U>(qreg, t_f32, j, l16, b);
    }
}

// An overflowing slice is re-scanned exactly by rescan_kernel, RESCAN_ROWS slice rows per
// work item, so that one unlucky query does not serialise thousands of rows in one warp.
constexpr int RESCAN_ROWS = 128;
struct WorkItem {                      // self-contained: no descriptor look-ups in rescan_kernel
    const float* q;                    // the query row
    const float* t;                    // first row of the train set
    unsigned long long* key;           // the query's two result slots
    SliceInfo si;
    int32_t r0, pad;
};

// ---- pair matching on tile top-2 records (Problem::exact bit 3, see t2_scale in vsm_common.cuh) -----------------
// Record layout: one PartialRec per (query, tile) = {H, L of column half 0, H, L of column half 1}: the exact two
// largest keys of each tile half.  Their union holds the exact two largest approximate dots a0 >= a1 of the whole
// train set, and every row not recorded lies below its slice's second entry (slice = tile * 2 + half).
//   forward problem of a ratio-only caller (skip_ratio2 > 0):
//     (1) dismissal from a0 / a1 alone, as for the other record kinds: no exact distance at all;
//     (2) a1 < a0 - 2*margin: the nearest neighbour is the row behind a0 and nobody else -- ONE exact distance e0;
//         e0 clearly above ratio * (upper bound on the second distance, from a1)  -> the test fails;
//         e0 clearly below ratio * (lower bound on every other row's distance)    -> it passes, and the second slot
//         gets a sentinel distance (FLT_MAX) that makes filter_kernel's fp32 test pass;
//     (3) otherwise (the ratio lands inside the bf16 band, ~2 % of the matching queries): t2_general_top2;
//   reverse problem of a mutual test (exact bit 4; filter_kernel reads only the nearest INDEX):
//     a1 < a0 - 2*margin: the index is the row behind a0, nothing is loaded or scored.  Otherwise the row is left
//     DEFERRED: filter_kernel resolves it (t2_resolve_top1) only if a surviving forward match points at it --
//     unmatched rows, whose two best are both noise and usually within the margin of each other, are hardly
//     ever asked for.
constexpr unsigned long long T2_DEFERRED = 1ull;          // decodes to index -2: never equal to a query

struct T2Top {                       // the two largest keys of a query over all its slices
    float K0, K1;
    int S0, S1;
};
__device__ __forceinline__ void t2_merge(T2Top& t, float hk, float lk, int s) {      // (hk >= lk) of slice s
    if (hk > t.K0) {
        if (t.K0 >= lk) { t.K1 = t.K0; t.S1 = t.S0; } else { t.K1 = lk; t.S1 = s; }
        t.K0 = hk; t.S0 = s;
    } else if (hk > t.K1) { t.K1 = hk; t.S1 = s; }
}
// Train index behind a key of slice s: a tile top-2 problem is one contiguous train range starting at index 0
// (the planner admits no run lists), slice s = (tile s / 2, column half s % 2) -- no SliceInfo load on the way to
// the exact distance.
__device__ __forceinline__ int32_t t2_row(const SliceInfo*, int s, float key) {
    return (s >> 1) * TILE_N + (s & 1) * HALF_N + t2_col(key);
}

// One warp, general case of one query: every recorded entry above thr is re-scored exactly (canonical distance),
// a slice whose SECOND entry is above thr may hide a third one and is re-scanned whole (128 rows): by rescan_kernel
// through `work` (key != nullptr: the result is pushed into key[0..1] by the caller, the re-scans by the atomic
// cascade), or right here (work == nullptr).  top1: thr = a0 - 2*margin, else a1 - 2*margin.
template <int CALLER>              // one copy per calling kernel: each is compiled under its caller's register budget
__device__ __noinline__ Best2 t2_general(const Problem& P, const PartialRec* __restrict__ recs, const SliceInfo* __restrict__ slices,
                                         int q, int lane, bool top1, unsigned long long* key, WorkItem* work, uint32_t work_cap,
                                         unsigned long long* counters) {
    const unsigned full = 0xffffffffu;
    const int h = lane >> 4, l16 = lane & 15;
    const int ntiles = P.nslices >> 1;
    const PartialRec* rq = recs + P.partial_off + (int64_t)q * ntiles;
    const SliceInfo* sl = slices + P.slice_off;
    // lane = slice (two rounds cover the 64 slices of the largest tile top-2 problem)
    float k[4];
    T2Top t = {0.f, 0.f, 0, 0};
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const int s = lane + 32 * j;
        float2 kv = make_float2(0.f, 0.f);
        if (s < P.nslices) kv = reinterpret_cast<const float2*>(rq)[s];
        k[2 * j] = kv.x;
        k[2 * j + 1] = kv.y;
        t2_merge(t, kv.x, kv.y, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float b0 = __shfl_xor_sync(full, t.K0, o), b1 = __shfl_xor_sync(full, t.K1, o);
        t2_merge(t, b0, b1, 0);
    }
    Best2 best = {FLT_MAX, FLT_MAX, -1, -1};
    const float Ktop = __shfl_sync(full, top1 ? t.K0 : t.K1, 0);
    float tmin2, tmax2;
    stats_read(P.t_stats, tmin2, tmax2);
    const float qn2 = __ldg(P.q_n2 + q);
    const float inv_s = 1.f / t2_scale(qn2, tmax2);
    const float thr = t2_valid(Ktop) ? t2_value(Ktop, inv_s) - 2.f * dot_margin(qn2, tmin2, tmax2) : -INFINITY;
    float qreg[16];
    load_qreg(qreg, P.q_f32 + (size_t)q * VSM_DIM, l16);
    unsigned long long n_cand = 0, n_flag = 0;
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const int s = lane + 32 * j;
        const bool have = s < P.nslices;
        const float v0 = have && t2_valid(k[2 * j]) ? t2_value(k[2 * j], inv_s) : -INFINITY;
        const float v1 = have && t2_valid(k[2 * j + 1]) ? t2_value(k[2 * j + 1], inv_s) : -INFINITY;
        const bool flagged = v1 > thr;                               // both entries above: a third may be hidden
        unsigned one = __ballot_sync(full, v0 > thr && !flagged);
        n_flag += __popc(__ballot_sync(full, flagged));
        n_cand += __popc(one);
        bool inline_scan = flagged && work == nullptr;
        if (flagged && work != nullptr) {
            const SliceInfo my = sl[s];
            const int span = slice_span(my);
            const uint32_t nitem = (uint32_t)((span + RESCAN_ROWS - 1) / RESCAN_ROWS);
            uint32_t* wcount = reinterpret_cast<uint32_t*>(counters + 2);
            const uint32_t base = atomicAdd(wcount, nitem);
            if (base + nitem <= work_cap) {
                for (uint32_t kk = 0; kk < nitem; kk++) {
                    WorkItem w = {P.q_f32 + (size_t)q * VSM_DIM, P.t_f32, key, my, (int32_t)(kk * RESCAN_ROWS), 0};
                    work[base + kk] = w;
                }
            } else {
                // list full: scan inline, and void the slots this reservation still owns
                inline_scan = true;
                for (uint32_t kk = base; kk < work_cap && kk < base + nitem; kk++) {
                    WorkItem w = {nullptr, nullptr, nullptr, my, 0, 0};
                    work[kk] = w;
                }
            }
        }
        unsigned scan = __ballot_sync(full, inline_scan);
        while (scan) {
            const int l0 = __ffs(scan) - 1; scan &= scan - 1;
            scan_slice<SELECT_U>(qreg, P.t_f32, sl[l0 + 32 * j], h, l16, best);
        }
        while (one) {                                                // two candidates per step, one per half-warp
            const int la = __ffs(one) - 1; one &= one - 1;
            int lb = -1;
            if (one) { lb = __ffs(one) - 1; one &= one - 1; }
            const int src = h ? lb : la;
            const float ks = __shfl_sync(full, k[2 * j], src < 0 ? 0 : src);
            int32_t jrow[1];
            jrow[0] = src >= 0 ? t2_row(sl, src + 32 * j, ks) : -1;
            score_n<1>(qreg, P.t_f32, jrow, l16, best);
        }
    }
    const float od0 = __shfl_sync(full, best.d0, 16), od1 = __shfl_sync(full, best.d1, 16);
    const int32_t oi0 = __shfl_sync(full, best.i0, 16), oi1 = __shfl_sync(full, best.i1, 16);
    if (oi0 >= 0) insert2(od0, oi0, best.d0, best.i0, best.d1, best.i1);
    if (oi1 >= 0) insert2(od1, oi1, best.d0, best.i0, best.d1, best.i1);
    best.d0 = __shfl_sync(full, best.d0, 0); best.d1 = __shfl_sync(full, best.d1, 0);
    best.i0 = __shfl_sync(full, best.i0, 0); best.i1 = __shfl_sync(full, best.i1, 0);
    if (lane == 0 && counters) {
        if (n_cand) atomicAdd(counters, n_cand);
        if (n_flag) atomicAdd(counters + 1, n_flag);
    }
    return best;
}

// filter_kernel's lazy resolution of a DEFERRED reverse row: exact nearest neighbour, slices scanned in place.
__device__ __forceinline__ unsigned long long t2_resolve_top1(const Problem& P, const PartialRec* __restrict__ recs,
                                                              const SliceInfo* __restrict__ slices, int q, int lane) {
    const Best2 b = t2_general<1>(P, recs, slices, q, lane, true, nullptr, nullptr, 0u, nullptr);
    return b.i0 >= 0 ? result_key(b.d0, b.i0) : 0ull;
}

// Lane = query (a warp takes 2^gshift <= 32 consecutive queries of one tile top-2 problem, Problem::gshift;
// groups[] = cumulative number of such warps per problem).  Reading the records, the dismissal and the uniqueness
// test are per-lane arithmetic; only the queries that need an exact distance take the warp's time: two at a time,
// one per half-warp.  The host picks the group size from the size of the call: 32 queries per warp when there are
// plenty (a ragged batch: 141K queries), 2 when the call is one frame pair and latency is what counts -- and a
// quarter of that for forward problems, whose warps walk through their exact distances one step after the other
// while the warps of reverse problems finish at once (ncu on the ragged batch: the kernel is bound by the latency
// of that walk, achieved occupancy 22 % of a possible 50 %).
constexpr int T2_SELECT_WARPS = 8;
__global__ void __launch_bounds__(T2_SELECT_WARPS * 32)
t2_select_kernel(const Problem* __restrict__ problems, int nproblems, const int32_t* __restrict__ groups,
                 const PartialRec* __restrict__ recs, const SliceInfo* __restrict__ slices,
                 unsigned long long* __restrict__ out_key, unsigned long long* __restrict__ counters,
                 WorkItem* __restrict__ work, uint32_t work_cap) {
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31, h = lane >> 4, l16 = lane & 15;
    const unsigned full = 0xffffffffu;
    const int g = (int)blockIdx.x * T2_SELECT_WARPS + (threadIdx.x >> 5);
    if (g >= groups[nproblems]) return;
    int lo = 0, hi = nproblems;                              // the problem with groups[p] <= g < groups[p + 1]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (groups[mid] <= g) lo = mid; else hi = mid;
    }
    const Problem P = problems[lo];
    const SliceInfo* sl = slices + P.slice_off;
    const int q = ((g - groups[lo]) << P.gshift) + lane;
    const bool live = lane < (1 << P.gshift) && q < P.nq;
    const int ntiles = P.nslices >> 1;
    const bool top1 = (P.exact & 16) != 0;

    // per lane: the two largest keys of the query, with their slices
    T2Top t = {0.f, 0.f, 0, 0};
    if (live) {
        const float4* rq = reinterpret_cast<const float4*>(recs + P.partial_off + (int64_t)q * ntiles);
        for (int n0 = 0; n0 < ntiles; n0 += 4) {             // four independent loads in flight: one latency per round
            float4 r[4];
#pragma unroll
            for (int j = 0; j < 4; j++) r[j] = n0 + j < ntiles ? rq[n0 + j] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                t2_merge(t, r[j].x, r[j].y, 2 * (n0 + j));
                t2_merge(t, r[j].z, r[j].w, 2 * (n0 + j) + 1);
            }
        }
    }
    float tmin2, tmax2;
    stats_read(P.t_stats, tmin2, tmax2);
    const float qn2 = live ? __ldg(P.q_n2 + q) : 1.f;
    const float margin = dot_margin(qn2, tmin2, tmax2);
    const float inv_s = 1.f / t2_scale(qn2, tmax2);
    unsigned long long* o = out_key + (P.out_off + (live ? q : 0)) * 2;
    const bool has0 = live && t2_valid(t.K0), has1 = has0 && t2_valid(t.K1);
    const float a0 = has0 ? t2_value(t.K0, inv_s) : 0.f;
    const float a1 = has1 ? t2_value(t.K1, inv_s) : -INFINITY;
    const int32_t i0c = has0 ? t2_row(sl, t.S0, t.K0) : -1;
    const bool unique = a1 < a0 - 2.f * margin;
    // 0 = answered here, 1 = one exact distance decides, 2 = general case
    int cls = 0;
    unsigned long long k0 = 0ull, k1 = 0ull;
    float lo1 = 0.f, hi1 = 0.f;
    if (has0) {
        if (top1) {
            k0 = unique ? result_key(0.f, i0c) : T2_DEFERRED;
        } else if (P.skip_ratio2 > 0.f && has1) {
            const float lo0 = qn2 + tmin2 - 2.f * (a0 + margin);
            hi1 = qn2 + tmax2 - 2.f * (a1 - margin);                 // the row behind a1 is at most this far (squared)
            lo1 = qn2 + tmin2 - 2.f * (a1 + margin);                 // every row but i0c is at least this far
            if (hi1 > 0.f && lo0 >= P.skip_ratio2 * hi1) cls = 0;    // the ratio test cannot pass
            else cls = unique ? 1 : 2;
        } else {
            cls = 2;
        }
    }
    // ratio-only queries the records could not dismiss: counted, so that the host can tell how rare matches are
    // (segmented_impl picks the next per-keyframe search's record kind from it)
    if (!top1 && P.skip_ratio2 > 0.f) {
        const unsigned open = __ballot_sync(full, has1 && cls != 0);     // (a train set of one row never matches)
        if (lane == 0 && open) atomicAdd(reinterpret_cast<uint32_t*>(counters + 2) + 1, (uint32_t)__popc(open));
    }
    // one exact distance: two queries per step, one per half-warp
    float e0 = 0.f;
    unsigned pend = __ballot_sync(full, cls == 1);
    unsigned n_cand = __popc(pend);
    while (pend) {
        const int la = __ffs(pend) - 1; pend &= pend - 1;
        int lb = -1;
        if (pend) { lb = __ffs(pend) - 1; pend &= pend - 1; }
        const int src = h ? lb : la;
        const int ssrc = src < 0 ? la : src;
        const int qq = __shfl_sync(full, q, ssrc);
        const int32_t row = __shfl_sync(full, i0c, ssrc);
        float qreg[16];
        load_qreg(qreg, P.q_f32 + (size_t)qq * VSM_DIM, l16);
        const float d2 = canon_l2sqr_halfwarp(qreg, P.t_f32 + (size_t)row * VSM_DIM, l16);
        const float e = __fsqrt_rn(d2);
        const float ea = __shfl_sync(full, e, 0), eb = __shfl_sync(full, e, 16);
        if (lane == la) e0 = ea;
        if (lane == lb) e0 = eb;
    }
    if (cls == 1) {
        const float r2 = P.ratio * P.ratio, e2 = e0 * e0;
        if (hi1 > 0.f && e2 >= r2 * hi1 * 1.002f) cls = 0;                                   // fails whatever the second is
        else if (lo1 > 0.f && e2 * 1.002f < r2 * lo1) { cls = 0; k0 = result_key(e0, i0c); k1 = result_key(FLT_MAX, 0); }
        else cls = 2;
    }
    if (live && cls == 0) { o[0] = k0; o[1] = k1; }
    // general case, one query at a time
    unsigned gen = __ballot_sync(full, cls == 2);
    while (gen) {
        const int l0 = __ffs(gen) - 1; gen &= gen - 1;
        const int qq = __shfl_sync(full, q, l0);
        unsigned long long* oq = out_key + (P.out_off + qq) * 2;
        const Best2 b = t2_general<0>(P, recs, slices, qq, lane, false, oq, work, work_cap, counters);
        // plain stores: this warp is the only writer of the slots until rescan_kernel runs
        if (lane == 0) {
            oq[0] = b.i0 >= 0 ? result_key(b.d0, b.i0) : 0ull;
            oq[1] = b.i1 >= 0 ? result_key(b.d1, b.i1) : 0ull;
        }
    }
    if (lane == 0 && n_cand) atomicAdd(counters, (unsigned long long)n_cand);
}

__global__ void __launch_bounds__(SELECT_WARPS * 32)
select_kernel(const Problem* __restrict__ problems, int problem0,
              const PartialRec* __restrict__ recs, const SliceInfo* __restrict__ slices,
              unsigned long long* __restrict__ out_key,
              unsigned long long* __restrict__ counters, WorkItem* __restrict__ work, uint32_t work_cap) {
    // grid: x = block of SELECT_WARPS queries, y = problem (ragged problems: surplus blocks exit)
    pdl_launch_dependents();
    pdl_wait();
    const Problem P = problems[problem0 + blockIdx.y];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = (int)blockIdx.x * SELECT_WARPS + warp;
    if (q >= P.nq || (P.exact & 8)) return;                 // tile top-2 problems: t2_select_kernel
    const int h = lane >> 4, l16 = lane & 15;
    const unsigned full = 0xffffffffu;

    // per-warp list of surviving train rows: < 2 x SELECT_U left over from earlier chunks + at most 32 x 4 new
    __shared__ int32_t s_list[SELECT_WARPS][2 * SELECT_U + 32 * VSM_TOPK];
    float qreg[16];
    Best2 best = {FLT_MAX, FLT_MAX, -1, -1};
    unsigned long long n_cand = 0, n_flag = 0;      // n_cand is per lane (summed at the end)
    const SliceInfo* sl = slices + P.slice_off;

    if (P.exact & 1) {
        load_qreg(qreg, P.q_f32 + (size_t)q * VSM_DIM, l16);
        for (int s = 0; s < P.nslices; s++) scan_slice<SELECT_U>(qreg, P.t_f32, sl[s], h, l16, best);
    } else if (P.exact & 4) {
        // ----- append records (small train sets): per slice a count and up to APPEND_CAP - 1 packed values that
        // passed the epilogue's running threshold.  Same two passes as below: the second-largest value over
        // everything recorded gives the final threshold; the survivors are re-scored exactly.
        const float* rq = reinterpret_cast<const float*>(recs + P.partial_off) + (int64_t)q * P.nslices * APPEND_CAP;
        load_qreg(qreg, P.q_f32 + (size_t)q * VSM_DIM, l16);          // needed by (almost) every query of a small problem
        float a0 = -INFINITY, a1 = -INFINITY;
        bool any_overflow = false;
        for (int s = 0; s < P.nslices; s++) {
            const float* rs = rq + (size_t)s * APPEND_CAP;
            const uint32_t cnt = __float_as_uint(rs[0]);
            any_overflow |= cnt > (uint32_t)(APPEND_CAP - 1);
            const uint32_t have = min(cnt, (uint32_t)(APPEND_CAP - 1));
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const uint32_t e = lane + 32 * k;
                if (e < have) {
                    const float v = rs[1 + e];
                    if (v > a0) { a1 = a0; a0 = v; } else if (v > a1) a1 = v;
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float b0 = __shfl_xor_sync(full, a0, o), b1 = __shfl_xor_sync(full, a1, o);
            if (b0 > a0) { a1 = fmaxf(a0, b1); a0 = b0; } else a1 = fmaxf(a1, b0);
        }
        float tmin2, tmax2;
        stats_read(P.t_stats, tmin2, tmax2);
        const float qn2 = __ldg(P.q_n2 + q);
        const float margin = dot_margin(qn2, tmin2, tmax2);
        const float thr = a1 - 2.f * margin;
        // ratio-only early-out (see below); a0 / a1 are the true largest values only if no slice overflowed
        if (P.skip_ratio2 > 0.f && a1 > VALID_FLOOR && !any_overflow) {
            const float lo0 = qn2 + tmin2 - 2.f * (a0 + margin);
            const float hi1 = qn2 + tmax2 - 2.f * (a1 - margin);
            if (hi1 > 0.f && lo0 >= P.skip_ratio2 * hi1) {
                if (lane == 0) {
                    unsigned long long* o = out_key + (P.out_off + q) * 2;
                    o[0] = 0ull;
                    o[1] = 0ull;
                }
                return;
            }
            if (lane == 0) atomicAdd(reinterpret_cast<uint32_t*>(counters + 2) + 1, 1u);
        }
        int32_t* list = s_list[warp];
        int cnt_l = 0;
        auto drain = [&]() {
            const int n = min(cnt_l, 2 * SELECT_U), base = cnt_l - n;
            int32_t j[SELECT_U];
#pragma unroll
            for (int u = 0; u < SELECT_U; u++) {
                const int slot = base + h * SELECT_U + u;
                j[u] = slot < cnt_l ? list[slot] : -1;
            }
            score_n<SELECT_U>(qreg, P.t_f32, j, l16, best);
            cnt_l = base;
            __syncwarp();
        };
        for (int s = 0; s < P.nslices; s++) {
            const float* rs = rq + (size_t)s * APPEND_CAP;
            const SliceInfo my = sl[s];
            const uint32_t cnt = __float_as_uint(rs[0]);
            if (cnt > (uint32_t)(APPEND_CAP - 1)) {
                // more values passed than the record holds: exact scan of the slice (rescan_kernel)
                n_flag++;
                int fits = 1;
                if (lane == 0) {
                    const int span = slice_span(my);
                    const uint32_t nitem = (uint32_t)((span + RESCAN_ROWS - 1) / RESCAN_ROWS);
                    uint32_t* wcount = reinterpret_cast<uint32_t*>(counters + 2);
                    const uint32_t base = atomicAdd(wcount, nitem);
                    fits = base + nitem <= work_cap;
                    for (uint32_t k = 0; k < nitem && base + k < work_cap; k++) {       // list full: the slots are voided
                        WorkItem w = {P.q_f32 + (size_t)q * VSM_DIM, P.t_f32, fits ? out_key + (P.out_off + q) * 2 : nullptr, my,
                                      (int32_t)(k * RESCAN_ROWS), 0};
                        work[base + k] = w;
                    }
                }
                fits = __shfl_sync(full, fits, 0);
                if (!fits) scan_slice<SELECT_U>(qreg, P.t_f32, my, h, l16, best);           // ... and the slice is scanned here
                continue;
            }
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const uint32_t e = lane + 32 * k;
                float v = -INFINITY;
                if (e < cnt) v = rs[1 + e];
                const bool on = v > VALID_FLOOR && v > thr;
                const uint32_t c = __float_as_uint(v) & ~PACK_MASK;
                const int32_t cand = my.t_index0 + (int32_t)(c / HALF_N) * TILE_N + my.half * HALF_N + (int32_t)(c % HALF_N);
                const unsigned bal = __ballot_sync(full, on);
                if (bal) {
                    if (on) list[cnt_l + __popc(bal & ((1u << lane) - 1u))] = cand;
                    cnt_l += __popc(bal);
                    n_cand += on ? 1 : 0;
                    __syncwarp();
                    while (cnt_l >= 2 * SELECT_U) drain();
                }
            }
        }
        while (cnt_l > 0) drain();
    } else {
        const PartialRec* rq = recs + P.partial_off + (int64_t)q * P.nslices;
        // pass 1: the second largest approximate dot over every record of the query
        float a0 = -INFINITY, a1 = -INFINITY;
#pragma unroll 4
        for (int s = lane; s < P.nslices; s += 32) {
            const float4 rv = *reinterpret_cast<const float4*>(rq + s);
            const float rs[VSM_TOPK] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
            for (int e = 0; e < 2; e++) {                // entries are sorted: only the first two matter
                float v = rs[e];
                if (v > a0) { a1 = a0; a0 = v; } else if (v > a1) a1 = v;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float b0 = __shfl_xor_sync(full, a0, o), b1 = __shfl_xor_sync(full, a1, o);
            if (b0 > a0) { a1 = fmaxf(a0, b1); a0 = b0; } else a1 = fmaxf(a1, b0);
        }
        float tmin2, tmax2;
        stats_read(P.t_stats, tmin2, tmax2);
        const float qn2 = __ldg(P.q_n2 + q);
        const float margin = dot_margin(qn2, tmin2, tmax2);
        const float thr = a1 - 2.f * margin;                                         // -inf if < 2 entries

        // Ratio-only callers (LoopCloser::detect's loop, match_features without raw list / mutual):
        // a0 is the largest approximate dot of the whole train set (a slice's maximum is always
        // recorded), so every exact squared distance is >= lo0; the row behind a1 exists, so the
        // exact SECOND-best squared distance is <= hi1.  If lo0 >= ratio^2 * hi1 (with 0.1 % slack for
        // the fp32 rounding of the distances and of the reference's `d0 < ratio * d1`), the ratio
        // test fails whatever the exact top-2 is: no match, no re-score.  Most (query, keyframe)
        // pairs of a loop-closure search end here.
        if (P.skip_ratio2 > 0.f && a1 > VALID_FLOOR) {       // a1 below the floor: fewer than two rows seen (or masked columns)
            const float lo0 = qn2 + tmin2 - 2.f * (a0 + margin);
            const float hi1 = qn2 + tmax2 - 2.f * (a1 - margin);
            if (hi1 > 0.f && lo0 >= P.skip_ratio2 * hi1) {
                if (lane == 0) {
                    unsigned long long* o = out_key + (P.out_off + q) * 2;
                    o[0] = 0ull;
                    o[1] = 0ull;
                }
                return;
            }
            // not dismissed: counted, so that the host can tell how rare matches are (see segmented_impl)
            if (lane == 0) atomicAdd(reinterpret_cast<uint32_t*>(counters + 2) + 1, 1u);
        }
        load_qreg(qreg, P.q_f32 + (size_t)q * VSM_DIM, l16);

        // pass 2: survivors -> this warp's candidate list -> exact distance, 2 x SELECT_U at a time;
        // overflowing slices -> exact scan.  The next chunk's records are loaded before the
        // current one is processed.
        int32_t* list = s_list[warp];
        int cnt = 0;
        auto drain = [&]() {                                 // score the last (up to) 2 x SELECT_U list entries
            const int n = min(cnt, 2 * SELECT_U), base = cnt - n;
            int32_t j[SELECT_U];
#pragma unroll
            for (int u = 0; u < SELECT_U; u++) {
                const int slot = base + h * SELECT_U + u;
                j[u] = slot < cnt ? list[slot] : -1;
            }
            score_n<SELECT_U>(qreg, P.t_f32, j, l16, best);
            cnt = base;
            __syncwarp();
        };
        const float4 none = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        float4 rv_next = none;
        SliceInfo my_next = {0, 0, 0, 0};
        if (lane < P.nslices) {
            rv_next = *reinterpret_cast<const float4*>(rq + lane);
            my_next = sl[lane];
        }
        for (int s0 = 0; s0 < P.nslices; s0 += 32) {
            const float4 rv = rv_next;
            const SliceInfo my = my_next;
            rv_next = none;
            if (s0 + 32 + lane < P.nslices) {
                rv_next = *reinterpret_cast<const float4*>(rq + s0 + 32 + lane);
                my_next = sl[s0 + 32 + lane];
            }
            const float rs[VSM_TOPK] = {rv.x, rv.y, rv.z, rv.w};
            // maxima-only records carry no candidates: every slice of a query that got here is re-scanned
            const bool flagged = (P.exact & 2) ? (s0 + lane < P.nslices)
                                               : (rs[VSM_TOPK - 1] > VALID_FLOOR && rs[VSM_TOPK - 1] > thr);
            // this lane's surviving entries (a bit per entry) and their logical train indices
            int32_t cand[VSM_TOPK];
            unsigned mine = 0;
#pragma unroll
            for (int e = 0; e < VSM_TOPK; e++) {
                const uint32_t c = __float_as_uint(rs[e]) & ~PACK_MASK;
                cand[e] = my.t_index0 + (int32_t)(c / HALF_N) * TILE_N + my.half * HALF_N + (int32_t)(c % HALF_N);
                if (!flagged && rs[e] > VALID_FLOOR && rs[e] > thr) mine |= 1u << e;
            }
            n_cand += __popc(mine);
            if (__ballot_sync(full, mine != 0)) {
#pragma unroll
                for (int e = 0; e < VSM_TOPK; e++) {
                    const bool on = (mine >> e) & 1u;
                    const unsigned bal = __ballot_sync(full, on);
                    if (on) list[cnt + __popc(bal & ((1u << lane) - 1u))] = cand[e];
                    cnt += __popc(bal);
                }
                __syncwarp();
                while (cnt >= 2 * SELECT_U) drain();
            }
            n_flag += __popc(__ballot_sync(full, flagged));
            // hand the overflowing slices to rescan_kernel; scan inline only if its list is full
            bool inline_scan = false;
            if (flagged) {
                const int span = slice_span(my);
                const uint32_t nitem = (uint32_t)((span + RESCAN_ROWS - 1) / RESCAN_ROWS);
                uint32_t* wcount = reinterpret_cast<uint32_t*>(counters + 2);
                const uint32_t base = atomicAdd(wcount, nitem);
                if (base + nitem <= work_cap) {
                    for (uint32_t k = 0; k < nitem; k++) {
                        WorkItem w = {P.q_f32 + (size_t)q * VSM_DIM, P.t_f32, out_key + (P.out_off + q) * 2, my,
                                      (int32_t)(k * RESCAN_ROWS), 0};
                        work[base + k] = w;
                    }
                } else {
                    // list full: scan inline, and void the slots this reservation still owns so
                    // that rescan_kernel never reads a stale item
                    inline_scan = true;
                    for (uint32_t k = base; k < work_cap && k < base + nitem; k++) {
                        WorkItem w = {nullptr, nullptr, nullptr, my, 0, 0};
                        work[k] = w;
                    }
                }
            }
            unsigned fm = __ballot_sync(full, inline_scan);
            while (fm) {
                int l0 = __ffs(fm) - 1; fm &= fm - 1;
                scan_slice<SELECT_U>(qreg, P.t_f32, sl[s0 + l0], h, l16, best);
            }
        }
        while (cnt > 0) drain();
    }
    unsigned long long n_cand_warp = n_cand;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n_cand_warp += __shfl_xor_sync(full, n_cand_warp, o);
    // merge the two half-warps (lane 16 -> lane 0)
    float od0 = __shfl_sync(full, best.d0, 16), od1 = __shfl_sync(full, best.d1, 16);
    int32_t oi0 = __shfl_sync(full, best.i0, 16), oi1 = __shfl_sync(full, best.i1, 16);
    if (lane == 0) {
        if (oi0 >= 0) insert2(od0, oi0, best.d0, best.i0, best.d1, best.i1);
        if (oi1 >= 0) insert2(od1, oi1, best.d0, best.i0, best.d1, best.i1);
        // plain stores: this warp is the only writer of the slots until rescan_kernel runs
        unsigned long long* o = out_key + (P.out_off + q) * 2;
        o[0] = best.i0 >= 0 ? result_key(best.d0, best.i0) : 0ull;
        o[1] = best.i1 >= 0 ? result_key(best.d1, best.i1) : 0ull;
        if (n_cand_warp) atomicAdd(counters, n_cand_warp);
        if (n_flag) atomicAdd(counters + 1, n_flag);
    }
}

// One block per work item: exact top-2 over RESCAN_ROWS rows of one slice for one query,
// pushed into the query's result slots with the atomicMax cascade.  16 half-warps, each
// scoring RESCAN_U rows per step (their loads in flight together).
__global__ void __launch_bounds__(256)
rescan_kernel(const WorkItem* __restrict__ work, const unsigned long long* __restrict__ counters, uint32_t work_cap) {
    pdl_launch_dependents();
    pdl_wait();
    const uint32_t nwork = min(*reinterpret_cast<const uint32_t*>(counters + 2), work_cap);
    __shared__ Best2 part[16];
    const int hw = threadIdx.x >> 4, l16 = threadIdx.x & 15;
    for (uint32_t it = blockIdx.x; it < nwork; it += gridDim.x) {
        const WorkItem w = work[it];
        if (w.key == nullptr) continue;                 // voided slot (block-uniform)
        const SliceInfo si = w.si;
        const int span = slice_span(si);
        const int r1 = min(span, w.r0 + RESCAN_ROWS);
        float qreg[16];
        load_qreg(qreg, w.q, l16);
        Best2 best = {FLT_MAX, FLT_MAX, -1, -1};
        // 16 half-warps x RESCAN_U rows per step, all loads of a step in flight together
        for (int k = 0; k < RESCAN_ROWS / (16 * RESCAN_U); k++) {
            int32_t j[RESCAN_U];
            float tv[RESCAN_U][16];
#pragma unroll
            for (int u = 0; u < RESCAN_U; u++) {
                const int r = w.r0 + k * 16 * RESCAN_U + u * 16 + hw;
                const int off = r < r1 ? slice_row(si, r) : -1;
                j[u] = off >= 0 ? si.t_index0 + off : -1;
                const float* p = w.t + (size_t)(j[u] >= 0 ? j[u] : 0) * VSM_DIM;
#pragma unroll
                for (int i = 0; i < 16; i++) tv[u][i] = __ldg(p + 16 * i + l16);
            }
#pragma unroll
            for (int u = 0; u < RESCAN_U; u++) {
                const float d2 = canon_l2sqr_halfwarp_regs(qreg, tv[u]);
                if (l16 == 0 && j[u] >= 0) insert2(__fsqrt_rn(d2), j[u], best.d0, best.i0, best.d1, best.i1);
            }
        }
        if (l16 == 0) part[hw] = best;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int k = 1; k < 16; k++) {
                if (part[k].i0 >= 0) insert2(part[k].d0, part[k].i0, best.d0, best.i0, best.d1, best.i1);
                if (part[k].i1 >= 0) insert2(part[k].d1, part[k].i1, best.d0, best.i0, best.d1, best.i1);
            }
            if (best.i0 >= 0) key_push2(w.key, result_key(best.d0, best.i0), best.i1 >= 0 ? result_key(best.d1, best.i1) : 0ull);
        }
        __syncthreads();
    }
}

// ---- LoopCloser::detect, compact form (src/LoopCloser.cpp:43-62) ---------------------------------
// The first tensor-core pass (fused units, TcUnit maps bit 3) leaves one bit per (eligible keyframe,
// query): "the ratio test could not be dismissed" = an OPEN pair.  Everything after it is O(open):
//   (tc_top3_kernel)       the epilogue that finds open pairs numbers them and pushes the unit back into the
//                          kernel's own queue as a top-4 unit (FusedArgs / RedoCtl in vsm_common.cuh): the
//                          second pass over the few (keyframe, query quarter) units that hold an open pair
//                          runs inside the same launch, ~4 candidates per open pair
//   loop_select_kernel     exact top-2 of every open pair from its candidates (canonical fp32 distance)
//   rescan_kernel          exact scans for the rare pair whose four recorded entries overflowed
//   loop_finish_kernel     ratio test per open pair (src/LoopCloser.cpp:55-60), survivors per keyframe;
//                          its last block applies the >= MIN_MATCHES gate (:62) and writes the surviving
//                          keyframes' lists, in query order, packed, straight into pinned host memory --
//                          nothing of size keyframes x queries ever exists
struct LoopSlot {                      // one eligible keyframe
    int64_t row0;                      // first store row
    int32_t count;                     // rows (>= 2)
    int32_t kf_pos;                    // position in Map::get_keyframes() order (DMatch::imgIdx)
};
struct LoopCand {                      // mirrors vsm_loop_candidate
    int32_t keyframe;
    int32_t count;
    int64_t offset;
};
struct LoopParams {
    const LoopSlot* slots;
    int32_t nslots, nq, words_per_slot;        // words_per_slot = 4 * query tiles
    const float* q_f32;                        // query rows (scratch arena)
    const float* q_n2;
    const float* store_f32;
    const uint32_t* t_stats;                   // norm statistics of the store
    const uint32_t* masks;                     // [nslots][words_per_slot]
    uint32_t* word_base;                       // [nslots][words_per_slot]: pair index of a word's first open bit
    unsigned long long* pair_keys;             // [pair_cap][2]
    PairRef* pair_ref;                         // [pair_cap]
    DMatch* stage;                             // [pair_cap]: the pair's match, trainIdx = -1 if it fails the ratio test
    uint32_t pair_cap;
    TcUnit* units2;                            // [unit2_cap] redo units, pushed by the first pass's epilogue
    uint32_t* hints2;                          // [unit2_cap][128]
    const PartialRec* recs2;                   // [unit2_cap][128][2]
    uint32_t unit2_cap;
    // aux block as uint32: [4] rescan work count, [5] open pairs, [7] overflow flag, [12..15] RedoCtl
    uint32_t* counters;
    WorkItem* work;
    uint32_t work_cap;
    int32_t* slot_good;                        // [nslots] survivors per eligible keyframe
    int64_t* slot_off;                         // [nslots] offset of its list in the output, -1 = below the gate
    float ratio;
    float skip_ratio2;
    int32_t min_matches;
    // outputs (pinned host memory)
    int32_t* out_head;                         // [0] candidates, [1] survivors emitted, [2] overflow, [3] open pairs
    int32_t* out_good;                         // [nslots] survivors per eligible keyframe (the caller's status array)
    LoopCand* out_cands;
    int32_t cand_cap;
    DMatch* out_matches;
    int64_t match_cap;
};

// One warp per open pair: candidates of the second pass's two records (one per column half of the
// keyframe) -> canonical fp32 distances -> the pair's exact top-2; a half whose four entries overflowed
// is re-scanned exactly by rescan_kernel.
__global__ void __launch_bounds__(SELECT_WARPS * 32)
loop_select_kernel(const LoopParams P) {
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31, h = lane >> 4, l16 = lane & 15;
    const unsigned full = 0xffffffffu;
    const uint32_t npairs = min(P.counters[5], P.pair_cap);
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < npairs; p += nwarps) {
        const PairRef r = P.pair_ref[p];
        unsigned long long* keys = P.pair_keys + 2 * (size_t)p;
        if (r.unit2 < 0) { if (lane == 0) { keys[0] = 0ull; keys[1] = 0ull; } continue; }       // overflow: the call is repeated on the record path
        const LoopSlot sl = P.slots[r.slot];
        const PartialRec* rq = P.recs2 + ((size_t)r.unit2 * TILE_M + (r.q % TILE_M)) * 2;
        // lanes 0..7 hold the eight recorded entries (two halves x top-4)
        float v = -INFINITY;
        if (lane < 8) v = rq[lane >> 2].s[lane & 3];
        // the two largest of the eight
        float a0 = v, a1 = -INFINITY;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            const float b0 = __shfl_xor_sync(full, a0, o), b1 = __shfl_xor_sync(full, a1, o);
            if (b0 > a0) { a1 = fmaxf(a0, b1); a0 = b0; } else a1 = fmaxf(a1, b0);
        }
        a0 = __shfl_sync(full, a0, 0);
        a1 = __shfl_sync(full, a1, 0);
        float tmin2, tmax2;
        stats_read(P.t_stats, tmin2, tmax2);
        const float qn2 = __ldg(P.q_n2 + r.q);
        const float margin = dot_margin(qn2, tmin2, tmax2);
        // with the true second-largest dot the ratio-only test is sharper than the first pass's
        if (P.skip_ratio2 > 0.f && a1 > VALID_FLOOR) {
            const float lo0 = qn2 + tmin2 - 2.f * (a0 + margin);
            const float hi1 = qn2 + tmax2 - 2.f * (a1 - margin);
            if (hi1 > 0.f && lo0 >= P.skip_ratio2 * hi1) {
                if (lane == 0) { keys[0] = 0ull; keys[1] = 0ull; }
                continue;
            }
        }
        const float thr = a1 - 2.f * margin;
        // a half is flagged when its 4th entry is still above the threshold: a needed row may have been displaced
        const float v3a = __shfl_sync(full, v, 3), v3b = __shfl_sync(full, v, 7);
        const bool flag_a = v3a > VALID_FLOOR && v3a > thr, flag_b = v3b > VALID_FLOOR && v3b > thr;
        const bool mine_flagged = (lane >> 2) ? flag_b : flag_a;
        const bool is_cand = lane < 8 && !mine_flagged && v > VALID_FLOOR && v > thr;
        const uint32_t c = __float_as_uint(v) & ~PACK_MASK;
        const int32_t row = (int32_t)(c / HALF_N) * TILE_N + (lane >> 2) * HALF_N + (int32_t)(c % HALF_N);
        unsigned cm = __ballot_sync(full, is_cand);
        float qreg[16];
        load_qreg(qreg, P.q_f32 + (size_t)r.q * VSM_DIM, l16);
        const float* t_f32 = P.store_f32 + (size_t)sl.row0 * VSM_DIM;
        Best2 best = {FLT_MAX, FLT_MAX, -1, -1};
        while (cm) {                                             // two candidates per round, one per half-warp
            const int l0 = __ffs(cm) - 1;
            cm &= cm - 1;
            int l1 = -1;
            if (cm) { l1 = __ffs(cm) - 1; cm &= cm - 1; }
            const int32_t j0 = __shfl_sync(full, row, l0);
            const int32_t j1 = __shfl_sync(full, row, l1 < 0 ? 0 : l1);
            score_pair(qreg, t_f32, h == 0 ? j0 : (l1 < 0 ? -1 : j1), l16, best);
        }
        const float od0 = __shfl_sync(full, best.d0, 16), od1 = __shfl_sync(full, best.d1, 16);
        const int32_t oi0 = __shfl_sync(full, best.i0, 16), oi1 = __shfl_sync(full, best.i1, 16);
        if (lane == 0) {
            if (oi0 >= 0) insert2(od0, oi0, best.d0, best.i0, best.d1, best.i1);
            if (oi1 >= 0) insert2(od1, oi1, best.d0, best.i0, best.d1, best.i1);
            keys[0] = best.i0 >= 0 ? result_key(best.d0, best.i0) : 0ull;      // plain stores: rescan_kernel runs after this kernel
            keys[1] = best.i1 >= 0 ? result_key(best.d1, best.i1) : 0ull;
            for (int hf = 0; hf < 2; hf++) {
                if (!(hf ? flag_b : flag_a)) continue;
                const SliceInfo si = {0, sl.count, hf, 0};
                const uint32_t nitem = (uint32_t)((slice_span(si) + RESCAN_ROWS - 1) / RESCAN_ROWS);
                const uint32_t wb = atomicAdd(P.counters + 4, nitem);
                const bool fits = wb + nitem <= P.work_cap;
                if (!fits) P.counters[7] = 1u;
                for (uint32_t k = 0; k < nitem && wb + k < P.work_cap; k++) {
                    WorkItem it = {P.q_f32 + (size_t)r.q * VSM_DIM, t_f32, fits ? keys : nullptr, si, (int32_t)(k * RESCAN_ROWS), 0};
                    P.work[wb + k] = it;
                }
            }
        }
    }
}

// The gate over the eligible keyframes (in list order) and the surviving keyframes' lists packed one
// after the other, each in query order.  Runs in ONE block (the last block of loop_finish_kernel).
__device__ void loop_emit_block(const LoopParams& P) {
    constexpr int LIST_CAP = 512;               // candidate keyframes remembered in shared memory (more: found again by scanning)
    __shared__ long long s_off;                 // running survivor offset
    __shared__ int s_cand;                      // running candidate count
    __shared__ int wcnt[32];
    __shared__ long long wsum[32];
    __shared__ int c_slot[LIST_CAP];
    __shared__ long long c_off[LIST_CAP];
    const int nthreads = blockDim.x, nw = nthreads >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { s_off = 0; s_cand = 0; }
    __syncthreads();
    for (int s0 = 0; s0 < P.nslots; s0 += nthreads) {
        const int s = s0 + threadIdx.x;
        const int g = s < P.nslots ? __ldcg(P.slot_good + s) : 0;        // written by other blocks' atomics: read at L2
        if (s < P.nslots) P.out_good[s] = g;
        const bool cand = s < P.nslots && g >= P.min_matches && g > 0;
        // block-wide exclusive scan of (cand, g) in slot order
        const unsigned bal = __ballot_sync(0xffffffffu, cand);
        long long v = cand ? g : 0;
        long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) { wsum[warp] = incl; wcnt[warp] = __popc(bal); }
        __syncthreads();
        long long off = s_off;
        int ci = s_cand;
        for (int w2 = 0; w2 < warp; w2++) { off += wsum[w2]; ci += wcnt[w2]; }
        off += incl - v;
        ci += __popc(bal & ((1u << lane) - 1u));
        if (cand) {
            if (ci < P.cand_cap) {
                LoopCand c = {P.slots[s].kf_pos, g, off};
                P.out_cands[ci] = c;
            }
            if (ci < LIST_CAP) { c_slot[ci] = s; c_off[ci] = off; P.slot_off[s] = -1; }
            else P.slot_off[s] = off;
        } else if (s < P.nslots) {
            P.slot_off[s] = -1;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            long long t = 0;
            int c = 0;
            for (int w2 = 0; w2 < nw; w2++) { t += wsum[w2]; c += wcnt[w2]; }
            s_off += t;
            s_cand += c;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        P.out_head[0] = s_cand;
        P.out_head[1] = (int32_t)(s_off < 0x7fffffffLL ? s_off : 0x7fffffffLL);
        P.out_head[2] = (int32_t)__ldcg(P.counters + 7);
        P.out_head[3] = (int32_t)__ldcg(P.counters + 5);
    }
    // lists: one warp per surviving keyframe, words in query order; a warp first loads 32 mask words
    // (and their pair bases) at once, then walks the non-empty ones
    auto emit = [&](int s, long long off) {
        for (int w0 = 0; w0 < P.words_per_slot; w0 += 32) {
            const int w = w0 + lane;
            const int64_t wi = (int64_t)s * P.words_per_slot + w;
            const uint32_t my_m = w < P.words_per_slot ? P.masks[wi] : 0u;
            const uint32_t my_b = my_m ? P.word_base[wi] : 0u;
            unsigned nz = __ballot_sync(0xffffffffu, my_m != 0u);
            while (nz) {
                const int k = __ffs(nz) - 1;
                nz &= nz - 1;
                const uint32_t m = __shfl_sync(0xffffffffu, my_m, k);
                const uint32_t base = __shfl_sync(0xffffffffu, my_b, k);
                DMatch dm = {0, -1, 0, 0.f};
                if ((m >> lane) & 1u) {
                    const uint32_t p = base + __popc(m & ((1u << lane) - 1u));
                    if (p < P.pair_cap) {                           // written by other blocks of this kernel: read at L2
                        const uint4 raw = __ldcg(reinterpret_cast<const uint4*>(P.stage + p));
                        dm.queryIdx = (int32_t)raw.x; dm.trainIdx = (int32_t)raw.y; dm.imgIdx = (int32_t)raw.z;
                        dm.distance = __uint_as_float(raw.w);
                    }
                }
                const unsigned gb = __ballot_sync(0xffffffffu, dm.trainIdx >= 0);
                const long long pos = off + __popc(gb & ((1u << lane) - 1u));
                if (dm.trainIdx >= 0 && pos < P.match_cap) P.out_matches[pos] = dm;
                off += __popc(gb);
            }
        }
    };
    const int ncand = s_cand;
    for (int c = warp; c < min(ncand, LIST_CAP); c += nw) emit(c_slot[c], c_off[c]);
    if (ncand > LIST_CAP)                                   // the rest were marked in slot_off
        for (int s = warp; s < P.nslots; s += nw) {
            const long long off = P.slot_off[s];
            if (off >= 0) emit(s, off);
        }
}

// One warp per mask word: the reference's test on the pair's exact top-2, survivors counted per keyframe.
// The block that finishes last (a ticket counter) applies the gate and emits the lists: one launch less.
__global__ void __launch_bounds__(256)
loop_finish_kernel(const LoopParams P) {
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int64_t nwords = (int64_t)P.nslots * P.words_per_slot;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < nwords; w += nwarps) {
        const uint32_t m = P.masks[w];
        if (m == 0u) continue;
        const int slot = (int)(w / P.words_per_slot);
        const int q = (int)(w % P.words_per_slot) * 32 + lane;
        const uint32_t base = P.word_base[w];
        bool good = false;
        if ((m >> lane) & 1u) {
            const uint32_t p = base + __popc(m & ((1u << lane) - 1u));
            if (p < P.pair_cap) {
                int32_t i0, i1;
                float d0, d1;
                key_decode(__ldcg(P.pair_keys + 2 * (size_t)p), i0, d0);
                key_decode(__ldcg(P.pair_keys + 2 * (size_t)p + 1), i1, d1);
                good = i1 >= 0 && d0 < __fmul_rn(P.ratio, d1);           // m.size() >= 2 && m[0].distance < ratio * m[1].distance
                DMatch dm = {q, good ? i0 : -1, P.slots[slot].kf_pos, d0};
                P.stage[p] = dm;
            }
        }
        const int n = __popc(__ballot_sync(0xffffffffu, good));
        if (lane == 0 && n) atomicAdd(P.slot_good + slot, n);
    }
    // last block done: every other block's stage[] / slot_good[] writes are visible after its fence + ticket
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(P.counters + 15, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    loop_emit_block(P);
}

// ---- match_features filter loop --------------------------------------------------
// One block per pair.  good = lists with two entries whose best passes the fp32 ratio
// test (and, if asked, the mutual-NN test); raw = every list with two entries.
// Query order is preserved (block-wide scan per 1024 queries).
constexpr int FILTER_THREADS = 1024;
__global__ void __launch_bounds__(FILTER_THREADS)
filter_kernel(const FilterJob* __restrict__ jobs, unsigned long long* __restrict__ out_key,
              DMatch* __restrict__ matches, int32_t* __restrict__ counts, const Problem* __restrict__ problems,
              const PartialRec* __restrict__ recs, const SliceInfo* __restrict__ slices) {
    pdl_launch_dependents();
    pdl_wait();
    const FilterJob J = jobs[blockIdx.x];
    __shared__ int wsum[2][FILTER_THREADS / 32];
    __shared__ int base[2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { base[0] = 0; base[1] = 0; }
    __syncthreads();
    for (int q0 = 0; q0 < J.nq; q0 += FILTER_THREADS) {
        const int q = q0 + threadIdx.x;
        bool is_raw = false, is_good = false;
        DMatch m = {q, -1, J.img_idx, 0.f};
        unsigned long long bk = 0ull;
        if (q < J.nq) {
            const int64_t o = (J.fwd_off + q) * 2;
            int32_t i0, i1;
            float d0, d1;
            key_decode(out_key[o], i0, d0);
            key_decode(out_key[o + 1], i1, d1);
            m.trainIdx = i0; m.distance = d0;
            is_raw = i1 >= 0;                                    // m.size() >= 2   (Slam.cpp:1152)
            is_good = is_raw && d0 < __fmul_rn(J.ratio, d1);     // fp32 product   (Slam.cpp:1154)
            if (is_good && J.back_off >= 0) bk = __ldcg(out_key + (J.back_off + i0) * 2);
        }
        if (J.back_prob >= 0) {
            // reverse rows the select pass left undecided, asked for by a surviving forward match: resolved here,
            // one warp per row (several matches may point at the same row: same answer, written twice)
            unsigned need = __ballot_sync(0xffffffffu, is_good && bk == T2_DEFERRED);
            while (need) {
                const int l0 = __ffs(need) - 1; need &= need - 1;
                const int t = __shfl_sync(0xffffffffu, m.trainIdx, l0);
                const unsigned long long r = t2_resolve_top1(problems[J.back_prob], recs, slices, t, lane);
                if (lane == l0) { bk = r; out_key[(J.back_off + t) * 2] = r; }
            }
        }
        if (is_good && J.back_off >= 0) {
            int32_t bi; float bd;
            key_decode(bk, bi, bd);
            is_good = bi == q;
        }
        const unsigned br = __ballot_sync(0xffffffffu, is_raw), bg = __ballot_sync(0xffffffffu, is_good);
        if (lane == 0) { wsum[0][warp] = __popc(bg); wsum[1][warp] = __popc(br); }
        __syncthreads();
        int pg = base[0], pr = base[1];
        for (int w = 0; w < warp; w++) { pg += wsum[0][w]; pr += wsum[1][w]; }
        const unsigned below = (1u << lane) - 1u;
        if (is_good) matches[J.good_off + pg + __popc(bg & below)] = m;
        if (is_raw && J.raw_off >= 0) matches[J.raw_off + pr + __popc(br & below)] = m;
        __syncthreads();
        if (threadIdx.x == 0) {
            int tg = 0, tr = 0;
            for (int w = 0; w < FILTER_THREADS / 32; w++) { tg += wsum[0][w]; tr += wsum[1][w]; }
            base[0] += tg; base[1] += tr;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { counts[2 * blockIdx.x] = base[0]; counts[2 * blockIdx.x + 1] = base[1]; }
}

// ---- shard merge ---------------------------------------------------------------------
// idx_in/dist_in: [nshard][nq][2] global indices (-1 = empty) -> [nq][2].
__global__ void merge_kernel(const int64_t* __restrict__ idx_in, const float* __restrict__ dist_in, int nshard,
                             int nq, int64_t* __restrict__ idx_out, float* __restrict__ dist_out) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    int64_t i0 = -1, i1 = -1;
    float d0 = FLT_MAX, d1 = FLT_MAX;
    for (int s = 0; s < nshard; s++) {
#pragma unroll
        for (int p = 0; p < 2; p++) {
            const int64_t j = idx_in[((int64_t)s * nq + q) * 2 + p];
            const float d = dist_in[((int64_t)s * nq + q) * 2 + p];
            if (j < 0) continue;
            const bool b1 = i1 < 0 || d < d1 || (d == d1 && j < i1);
            if (!b1) continue;
            const bool b0 = i0 < 0 || d < d0 || (d == d0 && j < i0);
            if (b0) { d1 = d0; i1 = i0; d0 = d; i0 = j; } else { d1 = d; i1 = j; }
        }
    }
    idx_out[2 * q] = i0; idx_out[2 * q + 1] = i1;
    dist_out[2 * q] = d0; dist_out[2 * q + 1] = d1;
}

// result keys -> int64 global index (row offset added) + distance, for the device-pointer DB search.
__global__ void widen_kernel(const unsigned long long* __restrict__ out_key, int n, int64_t row_offset,
                             int64_t* __restrict__ idx_out, float* __restrict__ dist_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int32_t j; float d;
    key_decode(out_key[i], j, d);
    idx_out[i] = j < 0 ? -1 : (int64_t)j + row_offset;
    dist_out[i] = d;
}

// dst row k <- src row sel[k] (fp32, 1 KB rows): one warp per row, two float4 per lane
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ src, const int32_t* __restrict__ sel, int64_t n, float* __restrict__ dst) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t k = warp0; k < n; k += nwarps) {
        const float4* p = reinterpret_cast<const float4*>(src + (int64_t)__ldg(sel + k) * VSM_DIM) + lane * 2;
        float4* o = reinterpret_cast<float4*>(dst + k * VSM_DIM) + lane * 2;
        const float4 a = __ldg(p), b = __ldg(p + 1);
        o[0] = a; o[1] = b;
    }
}

// ---- resident map-point table: selection of the rows a map-point search scans ------------------------
// The reference re-stacks the descriptors of the selected map points on the host for every search
// (src/Slam.cpp:552-557: valid points; :744-759: valid points with an observation within
// LC_NEARBY_FRAME_RANGE frames of the matched keyframe).  Here the points live on the device with their
// validity flags and their observation log; a search compacts the selection on the device, in
// ascending point id -- the order of the reference's loop, i.e. of its stacked matrix and its ties.
struct PointObs {
    int32_t point, frame;
};
// near[p] = 1 if point p has an observation with |frame - near_frame| < range  (:749-754)
__global__ void points_mark_near_kernel(const PointObs* __restrict__ log, int64_t nlog, int32_t near_frame, int32_t range,
                                        uint8_t* __restrict__ near) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nlog) return;
    const PointObs o = log[i];
    int32_t d = o.frame - near_frame;
    if (d < 0) d = -d;
    if (d < range) near[o.point] = 1;
}
constexpr int POINTS_PER_BLOCK = 2048;
// selected = valid (and near, if given); per-block counts
__global__ void __launch_bounds__(256)
points_count_kernel(const uint8_t* __restrict__ valid, const uint8_t* __restrict__ near, int64_t n, uint8_t* __restrict__ selflag,
                    int32_t* __restrict__ block_count) {
    const int64_t base = (int64_t)blockIdx.x * POINTS_PER_BLOCK;
    int c = 0;
    for (int k = threadIdx.x; k < POINTS_PER_BLOCK; k += 256) {
        const int64_t i = base + k;
        if (i >= n) break;
        const uint8_t f = (valid[i] != 0 && (near == nullptr || near[i] != 0)) ? 1 : 0;
        selflag[i] = f;
        c += f;
    }
    __shared__ int ws[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; w++) t += ws[w];
        block_count[blockIdx.x] = t;
    }
}
// one block: exclusive scan of the per-block counts; total[0] = number of selected points
__global__ void __launch_bounds__(1024)
points_scan_kernel(const int32_t* __restrict__ block_count, int nblocks, int32_t* __restrict__ block_off, int32_t* __restrict__ total) {
    __shared__ int wsum[32];
    __shared__ int carry;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int b0 = 0; b0 < nblocks; b0 += 1024) {
        const int b = b0 + threadIdx.x;
        const int v = b < nblocks ? block_count[b] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        int off = carry;
        for (int w = 0; w < warp; w++) off += wsum[w];
        if (b < nblocks) block_off[b] = off + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < 32; w++) t += wsum[w];
            carry += t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) total[0] = carry;
}
// sel[rank of point i among the selected] = i, ascending
__global__ void __launch_bounds__(256)
points_scatter_kernel(const uint8_t* __restrict__ selflag, int64_t n, const int32_t* __restrict__ block_off, int32_t* __restrict__ sel) {
    const int64_t base = (int64_t)blockIdx.x * POINTS_PER_BLOCK;
    __shared__ int ws[8];
    __shared__ int run;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) run = block_off[blockIdx.x];
    __syncthreads();
    for (int k0 = 0; k0 < POINTS_PER_BLOCK; k0 += 256) {
        const int64_t i = base + k0 + threadIdx.x;
        const bool f = i < n && selflag[i] != 0;
        const unsigned bal = __ballot_sync(0xffffffffu, f);
        if (lane == 0) ws[warp] = __popc(bal);
        __syncthreads();
        int off = run;
        for (int w = 0; w < warp; w++) off += ws[w];
        if (f) sel[off + __popc(bal & ((1u << lane) - 1u))] = (int32_t)i;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < 8; w++) t += ws[w];
            run += t;
        }
        __syncthreads();
    }
}
// result keys over the compacted rows -> (point id, distance); -1 / FLT_MAX = no such neighbour
__global__ void points_result_kernel(const unsigned long long* __restrict__ out_key, int n, const int32_t* __restrict__ sel,
                                     int64_t* __restrict__ idx_out, float* __restrict__ dist_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int32_t j; float d;
    key_decode(out_key[i], j, d);
    idx_out[i] = j < 0 ? -1 : (int64_t)sel[j];
    dist_out[i] = d;
}
// valid[ids[k]] = flag
__global__ void points_set_valid_kernel(const int32_t* __restrict__ ids, int n, uint8_t flag, uint8_t* __restrict__ valid) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) valid[ids[i]] = flag;
}

// local result keys -> keys carrying the GLOBAL index (for a single all-gather across shards)
__global__ void globalize_keys_kernel(const unsigned long long* __restrict__ out_key, int n, uint32_t row_offset,
                                      unsigned long long* __restrict__ keys) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long k = out_key[i];
    keys[i] = k == 0ull ? 0ull : ~((~k) + row_offset);        // index lives in the low 32 bits of ~k
}

// local result keys -> keys carrying the STACKED row index of a database whose keyframes are dealt to
// several devices (vsm_group): tab = (local row0, count, stacked row0) triples sorted by local row0;
// the key's row is looked up by binary search.  `keys` may point into a peer device's memory.
__global__ void stacked_keys_kernel(const unsigned long long* __restrict__ out_key, int n, const int64_t* __restrict__ tab,
                                    int ntab, unsigned long long* __restrict__ keys) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long k = out_key[i];
    unsigned long long g = 0ull;
    if (k != 0ull && ntab > 0) {
        const unsigned long long u = ~k;
        const int64_t row = (int64_t)(uint32_t)u;
        int lo = 0, hi = ntab;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (tab[3 * mid] <= row) lo = mid; else hi = mid;
        }
        const int64_t grow = tab[3 * lo + 2] + (row - tab[3 * lo]);
        g = ~((u & 0xFFFFFFFF00000000ull) | (unsigned long long)(uint32_t)grow);
    }
    keys[i] = g;
}

// gathered keys [nshard][nq][2] -> global top-2: the largest two keys per query
__global__ void merge_keys_kernel(const unsigned long long* __restrict__ keys, int nshard, int nq,
                                  int64_t* __restrict__ idx_out, float* __restrict__ dist_out) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    unsigned long long k0 = 0ull, k1 = 0ull;
    for (int s = 0; s < nshard; s++) {
#pragma unroll
        for (int p = 0; p < 2; p++) {
            const unsigned long long k = keys[((int64_t)s * nq + q) * 2 + p];
            if (k > k0) { k1 = k0; k0 = k; } else if (k > k1) k1 = k;
        }
    }
    const unsigned long long kk[2] = {k0, k1};
#pragma unroll
    for (int p = 0; p < 2; p++) {
        if (kk[p] == 0ull) { idx_out[2 * q + p] = -1; dist_out[2 * q + p] = FLT_MAX; continue; }
        const unsigned long long u = ~kk[p];
        idx_out[2 * q + p] = (int64_t)(uint32_t)u;
        dist_out[2 * q + p] = __uint_as_float((uint32_t)(u >> 32));
    }
}

// ---- synthetic descriptors (benchmarks written in C++ have no torch to make them) -------------------
// Row r of stream `seed`: 256 standard normals (counter-based hash + Box-Muller) scaled to unit length,
// the shape of FeatureExtractor's output (src/FeatureExtractor.cpp:170-205).  One warp per row.
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__global__ void __launch_bounds__(256)
synth_rows_kernel(float* __restrict__ dst, int64_t row0, int64_t n, uint64_t seed) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp0; r < n; r += nwarps) {
        float v[8];
        float s = 0.f;
#pragma unroll
        for (int p = 0; p < 4; p++) {
            const uint64_t h = mix64(mix64(seed) ^ (uint64_t)((row0 + r) * 128 + lane * 4 + p));
            const float u1 = ((uint32_t)(h >> 40) + 1u) * (1.0f / 16777217.0f);      // (0, 1]
            const float u2 = (uint32_t)(h & 0xFFFFFFu) * (1.0f / 16777216.0f);
            const float rad = sqrtf(-2.0f * __logf(u1));
            float sn, cs;
            __sincosf(6.28318530718f * u2, &sn, &cs);
            v[2 * p] = rad * cs;
            v[2 * p + 1] = rad * sn;
            s += v[2 * p] * v[2 * p] + v[2 * p + 1] * v[2 * p + 1];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float inv = rsqrtf(s);
        float4* o4 = reinterpret_cast<float4*>(dst + r * VSM_DIM) + lane * 2;
        o4[0] = make_float4(v[0] * inv, v[1] * inv, v[2] * inv, v[3] * inv);
        o4[1] = make_float4(v[4] * inv, v[5] * inv, v[6] * inv, v[7] * inv);
    }
}

// ---- Slam::track_local_map (src/Slam.cpp:380-469): projection + windowed descriptor search -------
struct TrackCfg {
    double fx, fy, cx, cy, depth_min, depth_max, radius_sq, desc_threshold;
    double R[9], t[3];
    int32_t width, height;
};

// One warp per map point.  kp_xy / kp_id are in the reference's visiting order (cell-major), so the
// first keypoint reaching the minimum wins, as in the reference's strict `<` (:457).
__global__ void __launch_bounds__(256)
track_local_map_kernel(TrackCfg cfg, const float2* __restrict__ kp_xy, const int32_t* __restrict__ kp_id, int nkp,
                       const float* __restrict__ frame_desc, const double* __restrict__ mp_pos,
                       const float* __restrict__ mp_desc, const uint8_t* __restrict__ mp_valid, int nmp,
                       int32_t* __restrict__ best_ki, double* __restrict__ best_dist) {
    const int lane = threadIdx.x & 31;
    const int mp = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (mp >= nmp) return;
    int32_t bk = -1;
    double bd = cfg.desc_threshold;
    bool live = mp_valid == nullptr || mp_valid[mp] != 0;
    double u = 0, v = 0;
    if (live) {
        const double X = mp_pos[3 * mp], Y = mp_pos[3 * mp + 1], Z = mp_pos[3 * mp + 2];
        // same operation order as :417-419 (no FMA contraction: explicit rounded ops)
        const double px = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(cfg.R[0], X), __dmul_rn(cfg.R[1], Y)), __dmul_rn(cfg.R[2], Z)), cfg.t[0]);
        const double py = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(cfg.R[3], X), __dmul_rn(cfg.R[4], Y)), __dmul_rn(cfg.R[5], Z)), cfg.t[1]);
        const double pz = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(cfg.R[6], X), __dmul_rn(cfg.R[7], Y)), __dmul_rn(cfg.R[8], Z)), cfg.t[2]);
        if (pz < cfg.depth_min || pz > cfg.depth_max) live = false;                       // :421
        else {
            u = __dadd_rn(__ddiv_rn(__dmul_rn(cfg.fx, px), pz), cfg.cx);                  // :423-424
            v = __dadd_rn(__ddiv_rn(__dmul_rn(cfg.fy, py), pz), cfg.cy);
            if (u < 0 || u >= cfg.width || v < 0 || v >= cfg.height) live = false;        // :426
        }
    }
    if (live) {
        // this map point's descriptor: 8 elements per lane, as doubles
        double md[8];
        const float4* mpd = reinterpret_cast<const float4*>(mp_desc + (size_t)mp * VSM_DIM) + lane * 2;
        const float4 m0 = __ldg(mpd), m1 = __ldg(mpd + 1);
        md[0] = m0.x; md[1] = m0.y; md[2] = m0.z; md[3] = m0.w; md[4] = m1.x; md[5] = m1.y; md[6] = m1.z; md[7] = m1.w;
        for (int k0 = 0; k0 < nkp; k0 += 32) {
            const int k = k0 + lane;
            bool in = false;
            if (k < nkp) {
                const float2 p = __ldg(kp_xy + k);
                const double dx = __dsub_rn(u, (double)p.x), dy = __dsub_rn(v, (double)p.y);
                in = !(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)) > cfg.radius_sq);   // :453-454
            }
            unsigned m = __ballot_sync(0xffffffffu, in);
            while (m) {
                const int l = __ffs(m) - 1;
                m &= m - 1;
                const int kk = k0 + l;
                const int32_t ki = __ldg(kp_id + kk);
                const float4* fd = reinterpret_cast<const float4*>(frame_desc + (size_t)ki * VSM_DIM) + lane * 2;
                const float4 f0 = __ldg(fd), f1 = __ldg(fd + 1);
                const double fv[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
                double s = 0.0;
#pragma unroll
                for (int e = 0; e < 8; e++) { const double d = md[e] - fv[e]; s = fma(d, d, s); }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                const double dist = sqrt(s);
                if (dist < bd) { bd = dist; bk = ki; }                                     // :456-460
            }
        }
    }
    if (lane == 0) { best_ki[mp] = bk; best_dist[mp] = bd; }
}

// ---- fused exchange over peer memory -----------------------------------------------------------
// Per rank one buffer: flags[2][XCHG_MAX_WORLD] (u32) then keys[2][world][nq_cap][2] (u64).
constexpr int XCHG_MAX_WORLD = 16;
constexpr size_t XCHG_FLAG_BYTES = 2 * XCHG_MAX_WORLD * sizeof(uint32_t);      // 128
struct XchgPeers {
    unsigned long long base[XCHG_MAX_WORLD];       // device addresses of every rank's buffer (own = local)
};
__device__ __forceinline__ uint32_t* xchg_flags(unsigned long long base, int parity) {
    return reinterpret_cast<uint32_t*>(base) + parity * XCHG_MAX_WORLD;
}
__device__ __forceinline__ unsigned long long* xchg_keys(unsigned long long base, int parity, int world, int nq_cap,
                                                         int src_rank) {
    return reinterpret_cast<unsigned long long*>(base + XCHG_FLAG_BYTES) +
           ((size_t)parity * world + src_rank) * (size_t)nq_cap * 2;
}

// One block.  (1) publish: this rank's keys (with global indices) are stored into EVERY rank's buffer
// -- peer stores travel over NVLink; (2) one flag per peer says "rank `rank`'s keys of step `step`
// are complete"; (3) wait for every peer's flag in the OWN buffer; (4) merge the world x 2 keys per
// query.  Two parities: a rank may publish step k+1 while a slower peer still merges step k.
// A peer that does not show up within `timeout_clocks` is a SOFT failure: the kernel writes
// XCHG_TIMEOUT (and the peer's rank) to the pinned status word, skips the merge and exits, so the
// host returns VSM_ERR_TIMEOUT instead of losing its CUDA context to a trap.
constexpr uint32_t XCHG_TIMEOUT = 0x7100u;
__global__ void __launch_bounds__(1024)
xchg_publish_merge_kernel(const unsigned long long* __restrict__ out_key, int nq, uint32_t row_offset, XchgPeers peers,
                          int rank, int world, int nq_cap, uint32_t step, int64_t* __restrict__ idx_out,
                          float* __restrict__ dist_out, long long timeout_clocks, volatile uint32_t* status) {
    const int parity = step & 1;
    const int n = nq * 2;
    __shared__ int s_timed_out;
    if (threadIdx.x == 0) s_timed_out = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const unsigned long long k = out_key[i];
        const unsigned long long g = k == 0ull ? 0ull : ~((~k) + row_offset);
        for (int p = 0; p < world; p++) {
            const int dst = (rank + p) % world;                         // spread the peers over time
            xchg_keys(peers.base[dst], parity, world, nq_cap, rank)[i] = g;
        }
    }
    __threadfence_system();                                             // my stores are visible system-wide ...
    __syncthreads();                                                    // ... and so are everybody's in this block
    if (threadIdx.x < world) {
        volatile uint32_t* f = xchg_flags(peers.base[threadIdx.x], parity) + rank;
        *f = step;                                                      // raise my flag at peer threadIdx.x
        __threadfence_system();
        // wait for peer threadIdx.x's flag in my own buffer (bounded: a dead peer must not hang the GPU)
        volatile uint32_t* w = xchg_flags(peers.base[rank], parity) + threadIdx.x;
        const long long t0 = clock64();
        while ((int32_t)(*w - step) < 0) {
            if (clock64() - t0 > timeout_clocks) {
                s_timed_out = 1;
                status[0] = XCHG_TIMEOUT;
                status[1] = (uint32_t)threadIdx.x;
                __threadfence_system();
                break;
            }
            __nanosleep(200);
        }
        __threadfence_system();
    }
    __syncthreads();
    if (s_timed_out) return;
    for (int q = threadIdx.x; q < nq; q += blockDim.x) {
        unsigned long long k0 = 0ull, k1 = 0ull;
        for (int s = 0; s < world; s++) {
            const volatile unsigned long long* ks = xchg_keys(peers.base[rank], parity, world, nq_cap, s) + 2 * q;
#pragma unroll
            for (int p = 0; p < 2; p++) {
                const unsigned long long k = ks[p];
                if (k > k0) { k1 = k0; k0 = k; } else if (k > k1) k1 = k;
            }
        }
        const unsigned long long kk[2] = {k0, k1};
#pragma unroll
        for (int p = 0; p < 2; p++) {
            if (kk[p] == 0ull) { idx_out[2 * q + p] = -1; dist_out[2 * q + p] = FLT_MAX; continue; }
            const unsigned long long u = ~kk[p];
            idx_out[2 * q + p] = (int64_t)(uint32_t)u;
            dist_out[2 * q + p] = __uint_as_float((uint32_t)(u >> 32));
        }
    }
}

}  // namespace vsm
