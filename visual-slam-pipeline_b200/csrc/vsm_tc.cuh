// vsm_tc.cuh -- the tensor-core pass: bf16 q.t tiles on tcgen05 with a fused
// per-query top-3 epilogue (sm_100a only).
//
// A TcUnit = 128 query rows x a contiguous range of train rows; persistent CTAs (one per SM)
// take units from an atomic counter.
//   warp 0      TMA producer: the 128x256 bf16 query tile once (4 boxes of 64 columns,
//               SWIZZLE_128B), then the train rows as a ring of 4 stages, each stage one
//               64-column K-chunk of a 256-row tile (2 boxes of 128 rows, 32 KB).
//   warp 1      MMA issuer (one thread): per 256-row tile 16 x tcgen05.mma
//               cta_group::1 kind::f16 M=128 N=256 K=16, fp32 accumulators in TMEM,
//               two accumulator stages (2 x 256 columns = all 512 TMEM columns).
//   warp 2      TMEM allocator.
//   warps 4-11  epilogue: two warp-groups; group h owns columns [128h, 128h+128) of
//               every tile.  Thread = one query row (TMEM lane).  tcgen05.ld 32 columns
//               at a time; per 8 values one max-tree and one compare against the
//               thread's threshold; the rare survivors go into a register top-3.
//
// What the epilogue keeps (see dot_margin in vsm_common.cuh for the bound): for each
// slice (seg_tiles tiles of one column half) the three largest approximate dots among
// the values v with v > g2 - 2*margin, g2 = the second largest value this thread has
// seen so far in the whole unit.  g2 never exceeds the final global second-best, so a
// value the select kernel needs (v > a2 - 2*margin) is never filtered here; it can only
// be displaced by three larger values of the same slice, which the select kernel
// detects (third entry above its threshold) and answers with an exact scan of the slice.
#pragma once

#include <cuda.h>
#include "vsm_common.cuh"

namespace vsm {
namespace tc {

constexpr int KCHUNK = 64;                       // bf16 per 128-byte swizzle row
constexpr int NCHUNK = VSM_DIM / KCHUNK;         // 4 K-chunks
constexpr int STAGES = 4;                        // train-chunk ring depth
constexpr uint32_t Q_SUB_BYTES = TILE_M * 128;   // 16 KB: 128 rows x 64 bf16
constexpr uint32_t T_STAGE_BYTES = TILE_N * 128; // 32 KB: 256 rows x 64 bf16
constexpr uint32_t SMEM_Q = 0;
constexpr uint32_t SMEM_T = NCHUNK * Q_SUB_BYTES;                 // 65536
constexpr uint32_t SMEM_BAR = SMEM_T + STAGES * T_STAGE_BYTES;    // 196608
constexpr uint32_t SMEM_XCHG = SMEM_BAR + 512;                    // [2][128] float2: column-half exchange of fused units
constexpr int APPEND_SLOTS = 16;                                  // chunks (32 values) a warp stages per round in append mode
constexpr uint32_t APPEND_WARP_BYTES = APPEND_SLOTS * (32 * 4 + 8);   // values + (threshold, source lane) per slot
constexpr uint32_t SMEM_APPEND = SMEM_XCHG + 2048;                // [8 epilogue warps][APPEND_WARP_BYTES]
constexpr uint32_t SMEM_BYTES = SMEM_APPEND + 8 * APPEND_WARP_BYTES + 1024;   // ... + alignment slack
constexpr int THREADS = 384;
constexpr int EPI_WARP0 = 4;
constexpr uint32_t TMEM_COLS = 512;
constexpr int L2_AHEAD = 4;                      // tiles of L2 prefetch distance

// ---- PTX wrappers -------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a protocol bug must surface as a launch failure, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    long long t0 = 0;
    for (uint32_t it = 0;; it++) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
        if ((it & 1023u) == 1023u) {
            long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000LL) __trap();     // ~2 s at 2 GHz
        }
    }
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int32_t c0, int32_t c1,
                                            uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* map, int32_t c0, int32_t c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, bf16 in, fp32 accumulate.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrives when all MMAs issued so far by this thread have completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                 : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (asynchronous).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    __syncwarp();                                       // .sync.aligned: the warp must be converged
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
// The same load under a warp-uniform predicate (off: the registers keep their values).  (ptxas turns it into a
// branch around the load.)
__device__ __forceinline__ void tmem_ld32_if(bool on, uint32_t taddr, uint32_t (&r)[32]) {
    __syncwarp();
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %33, 0;\n\t"
        "@p tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t}"
        : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
          "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
          "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
          "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
        : "r"(taddr), "r"((uint32_t)on)
        : "memory");
}
// Wait for this thread's outstanding tcgen05.ld; the registers pass through the
// statement so no use of them can be scheduled above it.
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
    __syncwarp();
    asm volatile(
        "tcgen05.wait::ld.sync.aligned;"
        : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
          "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
          "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
          "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
        :
        : "memory");
}

// Shared-memory matrix descriptor: K-major operand, SWIZZLE_128B, rows of 64 bf16
// (128 B), 8-row groups 1024 B apart (what the TMA boxes above produce).
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);      // start address, 16-byte units
    d |= (uint64_t)1 << 16;                             // leading byte offset (unused with swizzle)
    d |= (uint64_t)(1024u >> 4) << 32;                  // stride byte offset: 8 rows x 128 B
    d |= (uint64_t)1 << 46;                             // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                             // SWIZZLE_128B
    return d;
}
// Instruction descriptor: D fp32, A/B bf16, both K-major, N=256, M=128.
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TILE_N >> 3) << 17) |
                           ((uint32_t)(TILE_M >> 4) << 24);

// ---- epilogue state -------------------------------------------------------------
// Slice top-4 as PACKED floats: the low 13 mantissa bits of an accumulator value are
// replaced by its column number inside the slice (a slice is at most 64 tiles x 128
// columns = 2^13), so the whole (value, index) top-4 update is seven FMNMX and no branch.
// The packing moves a value by < 2^-10 relative; dot_margin() accounts for it.
struct Top3 {
    float b0, b1, b2, b3;      // slice top-4 (packed), descending; -inf = empty
    float G;                   // lower bound on the query's global second-best dot
    float margin2;             // 2 * dot_margin
    float thr;                 // max(b3, max(G, b1) - margin2): values <= thr are dropped
    float published;           // last bound pushed to the shared hint
};

__device__ __forceinline__ void top3_update_thr(Top3& s) {
    s.thr = fmaxf(s.b3, fmaxf(s.G, s.b1) - s.margin2);
}
__device__ __forceinline__ void top3_reset_slice(Top3& s) {
    s.G = fmaxf(s.G, s.b1);                     // second best of a subset <= global second best
    s.b0 = s.b1 = s.b2 = s.b3 = -INFINITY;
    top3_update_thr(s);
}
__device__ __forceinline__ void top3_push(Top3& s, float x) {
    float t0 = fmaxf(s.b0, x), x1 = fminf(s.b0, x);
    float t1 = fmaxf(s.b1, x1), x2 = fminf(s.b1, x1);
    float t2 = fmaxf(s.b2, x2), x3 = fminf(s.b2, x2);
    s.b0 = t0; s.b1 = t1; s.b2 = t2; s.b3 = fmaxf(s.b3, x3);
}
// monotone float <-> uint32 (for atomicMax on the shared bound); 0 = "nothing yet"
__device__ __forceinline__ uint32_t enc_ordered(float f) {
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float dec_ordered(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// The rare path lives OUT OF LINE (one copy for the whole kernel) and takes / returns the state
// by value, i.e. in registers: inlined, sixteen copies of it made the epilogue loop tens of KB
// of SASS and instruction fetch -- not the tensor pipe -- set the pace (ncu: icc hit rate 68 %,
// stall_no_instruction 5.8 per issue).
struct Top3Core {
    float b0, b1, b2, b3;
};
// compare-exchange: hi = max, lo = min (two FMNMX, independent of each other)
#define VSM_CE(hi, lo)                      \
    do {                                    \
        const float mx_ = fmaxf(hi, lo);    \
        lo = fminf(hi, lo);                 \
        hi = mx_;                           \
    } while (0)

__device__ __noinline__ Top3Core top3_push8(Top3Core c, float v0, float v1, float v2, float v3, float v4, float v5,
                                            float v6, float v7, uint32_t scol) {
    // All eight values enter, unconditionally (measured: an `if (v > thr)` around each insert, skipped by
    // the warp when no lane needs that position, is 50 % SLOWER on small problems).  They used to be
    // inserted one after the other -- 8 x 7 FMNMX in a dependent chain 32 deep, and with only two epilogue
    // warps per scheduler that latency, not the issue rate, set the pace of match-heavy small problems
    // (ncu: 0.66 eligible warps per scheduler).  Now: the eight are sorted as two quads by a network, their
    // top four are merged with the running top four by two bitonic steps -- 44 FMNMX, depth 9.
    // Same result: the four largest of the multiset {b0..b3} U {v0..v7}, descending.
    float a0 = __uint_as_float((__float_as_uint(v0) & PACK_MASK) | (scol + 0));
    float a1 = __uint_as_float((__float_as_uint(v1) & PACK_MASK) | (scol + 1));
    float a2 = __uint_as_float((__float_as_uint(v2) & PACK_MASK) | (scol + 2));
    float a3 = __uint_as_float((__float_as_uint(v3) & PACK_MASK) | (scol + 3));
    float e0 = __uint_as_float((__float_as_uint(v4) & PACK_MASK) | (scol + 4));
    float e1 = __uint_as_float((__float_as_uint(v5) & PACK_MASK) | (scol + 5));
    float e2 = __uint_as_float((__float_as_uint(v6) & PACK_MASK) | (scol + 6));
    float e3 = __uint_as_float((__float_as_uint(v7) & PACK_MASK) | (scol + 7));
    // sort each quad descending (5 compare-exchanges, depth 3)
    VSM_CE(a0, a1); VSM_CE(a2, a3); VSM_CE(e0, e1); VSM_CE(e2, e3);
    VSM_CE(a0, a2); VSM_CE(a1, a3); VSM_CE(e0, e2); VSM_CE(e1, e3);
    VSM_CE(a1, a2); VSM_CE(e1, e2);
    // top four of the eight: max(a_i, e_{3-i}) is a bitonic sequence holding them; sort it (depth 2)
    float t0 = fmaxf(a0, e3), t1 = fmaxf(a1, e2), t2 = fmaxf(a2, e1), t3 = fmaxf(a3, e0);
    VSM_CE(t0, t2); VSM_CE(t1, t3);
    VSM_CE(t0, t1); VSM_CE(t2, t3);
    // merge with the running top four (both descending): same two steps
    float r0 = fmaxf(c.b0, t3), r1 = fmaxf(c.b1, t2), r2 = fmaxf(c.b2, t1), r3 = fmaxf(c.b3, t0);
    VSM_CE(r0, r2); VSM_CE(r1, r3);
    VSM_CE(r0, r1); VSM_CE(r2, r3);
    c.b0 = r0; c.b1 = r1; c.b2 = r2; c.b3 = r3;
    return c;
}

// 32 accumulator values of one thread = slice columns [scol0, scol0+32).
// Fast path: four max-trees, ONE compare and ONE branch per 32 values.  The rare path re-checks
// the four groups and pushes only those that beat the threshold (packed top-4 insert).
__device__ __forceinline__ void scan32(Top3& s, const uint32_t (&r)[32], uint32_t scol0, int* slow = nullptr) {
    float m[4];
#pragma unroll
    for (int g = 0; g < 4; g++) {
        const float* v = reinterpret_cast<const float*>(&r[8 * g]);
        m[g] = fmaxf(fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3])), fmaxf(fmaxf(v[4], v[5]), fmaxf(v[6], v[7])));
    }
    if (slow && __any_sync(0xffffffffu, fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3])) > s.thr)) (*slow)++;   // debug only
    if (fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3])) > s.thr) {
#pragma unroll
        for (int g = 0; g < 4; g++) {
            if (m[g] > s.thr) {
                const float* v = reinterpret_cast<const float*>(&r[8 * g]);
                Top3Core c = {s.b0, s.b1, s.b2, s.b3};
                c = top3_push8(c, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], scol0 + 8 * g);
                s.b0 = c.b0; s.b1 = c.b1; s.b2 = c.b2; s.b3 = c.b3;
                top3_update_thr(s);
            }
        }
    }
}

// Maxima-only form (units with maps bit 2): the running top-2 of the GROUP MAXIMA of a slice --
// g0 is the slice's exact maximum, g1 a value of another row (a lower bound on the slice's second
// largest).  No insert path at all: 28 instructions per 32 values whatever the data.
__device__ __forceinline__ void scan32_max2(float& g0, float& g1, const uint32_t (&r)[32]) {
#pragma unroll
    for (int g = 0; g < 4; g++) {
        const float* v = reinterpret_cast<const float*>(&r[8 * g]);
        const float m = fmaxf(fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3])), fmaxf(fmaxf(v[4], v[5]), fmaxf(v[6], v[7])));
        g1 = fmaxf(g1, fminf(g0, m));
        g0 = fmaxf(g0, m);
    }
}

// Append form (units with maps bit 4).  Small train sets are match-heavy: with ~1000 rows every warp (32
// queries) has a value above the threshold in almost every 8-group, so the top-4 insert path ran for ~15
// of 16 groups per tile with 4 of 32 lanes active (ncu) and the epilogue, not the tensor pipe, set the
// pace (5500 cycles per tile against 2350 for the MMAs).  Here nothing is inserted and no state is
// carried from group to group: the running top-2 of the group maxima gives the threshold (a lower
// bound on the final second best minus the margin, exactly as in the top-4 form), and every value above
// it is stored straight to the query's record -- a predicated store per value, no dependency chain.
struct AppendState {
    float g0, g1;              // running top-2 of the group maxima (g1: a value of another row than g0's)
    float G;                   // lower bound on the query's global second-best dot (other warps / CTAs)
    float margin2;
    float published;
    uint32_t cnt;              // values that passed so far (stored only while below APPEND_CAP - 1)
};
// One 32-value chunk of every lane's row.  The lanes whose chunk holds a value above their threshold (a
// few per warp) stage the chunk in shared memory; then the WHOLE warp works on one staged chunk at a
// time, one value per lane: compare, ballot, and the passing lanes store to the owner row's record.  The
// per-value work of the sparse hits runs with all 32 lanes instead of one lane in eight.
//   stage: this warp's [APPEND_SLOTS][32] floats followed by [APPEND_SLOTS] (threshold, source lane) pairs
//   rec0:  record of the warp's row 0 (this column half); row r's record is rec0 + r * row_floats
__device__ __forceinline__ void append32(AppendState& s, const uint32_t (&r)[32], uint32_t scol0, float* __restrict__ stage,
                                         float* __restrict__ rec0, int64_t row_floats, int rows_valid, int lane) {
    float m[4];
#pragma unroll
    for (int g = 0; g < 4; g++) {
        const float* v = reinterpret_cast<const float*>(&r[8 * g]);
        m[g] = fmaxf(fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3])), fmaxf(fmaxf(v[4], v[5]), fmaxf(v[6], v[7])));
        s.g1 = fmaxf(s.g1, fminf(s.g0, m[g]));
        s.g0 = fmaxf(s.g0, m[g]);
    }
    // the chunk's own maxima already count: g1 is still a lower bound on the final second best
    const float thr = fmaxf(fmaxf(s.G, s.g1) - s.margin2, VALID_FLOOR);
    bool hit = fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3])) > thr;
    unsigned bal = __ballot_sync(0xffffffffu, hit);
    float2* meta = reinterpret_cast<float2*>(stage + APPEND_SLOTS * 32);
    const unsigned lt = (1u << lane) - 1u;
    while (bal) {                                            // warp-uniform: usually one round of ~4 chunks
        const int slot = __popc(bal & lt);
        if (hit && slot < APPEND_SLOTS) {
            uint4* dst = reinterpret_cast<uint4*>(stage + slot * 32);
#pragma unroll
            for (int j = 0; j < 8; j++) dst[j] = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
            meta[slot] = make_float2(thr, __int_as_float(lane));
            hit = false;
        }
        const int n = min(__popc(bal), APPEND_SLOTS);
        __syncwarp();
        for (int k = 0; k < n; k++) {
            const float v = stage[k * 32 + lane];
            const float2 mt = meta[k];
            const int src = __float_as_int(mt.y);
            const bool p = v > mt.x;
            const unsigned pb = __ballot_sync(0xffffffffu, p);
            const uint32_t base = __shfl_sync(0xffffffffu, s.cnt, src);
            const uint32_t idx = base + __popc(pb & lt);
            const uint32_t cap = src < rows_valid ? (uint32_t)(APPEND_CAP - 1) : 0u;
            if (p && idx < cap)
                rec0[(int64_t)src * row_floats + 1 + idx] = __uint_as_float((__float_as_uint(v) & PACK_MASK) | (scol0 + lane));
            if (lane == src) s.cnt += __popc(pb);
        }
        __syncwarp();
        bal = __ballot_sync(0xffffffffu, hit);
    }
}

// Tile top-2 form (units with maps bit 5, see t2_scale in vsm_common.cuh).  Per value: three FMA-pipe
// instructions make the key, ~2.5 ALU-pipe instructions find the exact top-2 (sorted pairs, top-2 merges of
// max, min and FMNMX3); no branch depends on the data.  The column inside the chunk (5 bits) goes into the
// keys, the chunk's first column (colbase = col0 * 2^-23) is added to the chunk's two winners only.
// Both pipes accept one warp instruction every other cycle and there are only two epilogue warps per scheduler.
// Written as "all keys, then the tree" the two kinds of work alternated, and the two warps -- released by the
// same barrier -- alternated in step (ncu: 0.43 IPC per scheduler, neither pipe above 42 %).  So the chunk is cut
// into eight sub-blocks of four values and the KEYS of sub-block i+1 are written between the TREE steps of
// sub-block i: consecutive instructions go to different pipes.
__device__ __forceinline__ void t2_step(float& H, float& L, const uint32_t (&r)[32], float s, float colbase) {
    float k[32];
    float CH = -INFINITY, CL = -INFINITY;
#pragma unroll
    for (int sb = 0; sb <= 8; sb++) {
        if (sb < 8) {
#pragma unroll
            for (int e = 4 * sb; e < 4 * sb + 4; e++)
                k[e] = __fadd_rn(__fadd_rn(__fmaf_rn(__uint_as_float(r[e]), s, 192.0f), -190.5f), (float)e * 1.1920928955078125e-7f);
        }
        if (sb > 0) {
            const int b = 4 * (sb - 1);
            const float h0 = fmaxf(k[b], k[b + 1]), l0 = fminf(k[b], k[b + 1]);
            const float h1 = fmaxf(k[b + 2], k[b + 3]), l1 = fminf(k[b + 2], k[b + 3]);
            const float h = fmaxf(h0, h1);
            const float l = fmaxf(fmaxf(fminf(h0, h1), l0), l1);
            const float nh = fmaxf(CH, h);
            CL = fmaxf(fmaxf(fminf(CH, h), CL), l);
            CH = nh;
        }
    }
    const float ch = __fadd_rn(CH, colbase), cl = __fadd_rn(CL, colbase);
    const float h = fmaxf(H, ch);
    L = fmaxf(fmaxf(fminf(H, ch), L), cl);
    H = h;
}

// A partial last tile: columns at or past the end of the train range become MASKED_VALUE.
__device__ __forceinline__ void mask32(uint32_t (&r)[32], int32_t ucol0, int32_t t_count) {
#pragma unroll
    for (int e = 0; e < 32; e++)
        if (ucol0 + e >= t_count) r[e] = __float_as_uint(MASKED_VALUE);
}

__device__ __forceinline__ uint32_t ld_relaxed(const uint32_t* p) {
    return *reinterpret_cast<const volatile uint32_t*>(p);
}
// A redo unit written by another SM (after its ready flag was seen and a fence): read at L2.
__device__ __forceinline__ TcUnit load_unit_cg(const TcUnit* p) {
    TcUnit u;
    static_assert(sizeof(TcUnit) % 16 == 0, "TcUnit is copied in 16-byte pieces");
    const uint4* src = reinterpret_cast<const uint4*>(p);
    uint4* dst = reinterpret_cast<uint4*>(&u);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(TcUnit) / 16); i++) dst[i] = __ldcg(src + i);
    return u;
}
// Blocking: the next redo unit of the compact loop search, or false once every first-pass epilogue has
// finished (nothing can be produced any more) and the queue is drained.  Bounded like the mbarrier waits.
__device__ __noinline__ bool take_redo(const FusedArgs* fa, TcUnit& out) {
    RedoCtl* ctl = fa->ctl;
    long long t0 = 0;
    for (uint32_t it = 0;; it++) {
        const uint32_t cnt = min(ld_relaxed(&ctl->produced), fa->unit2_cap);
        const uint32_t h = ld_relaxed(&ctl->head);
        if (h < cnt) {
            if (atomicCAS(&ctl->head, h, h + 1u) != h) continue;
            for (uint32_t w = 0; ld_relaxed(fa->ready2 + h) == 0u; w++)
                if ((w & 4095u) == 4095u) {
                    const long long now = clock64();
                    if (t0 == 0) t0 = now; else if (now - t0 > 4000000000LL) __trap();
                }
            __threadfence();
            out = load_unit_cg(fa->units2 + h);
            return true;
        }
        if (ld_relaxed(&ctl->main_done) >= 4u * fa->n_main) {
            __threadfence();
            const uint32_t cnt2 = min(ld_relaxed(&ctl->produced), fa->unit2_cap);
            if (ld_relaxed(&ctl->head) >= cnt2) return false;
            continue;
        }
        __nanosleep(200);
        if ((it & 1023u) == 1023u) {
            const long long now = clock64();
            if (t0 == 0) t0 = now; else if (now - t0 > 4000000000LL) __trap();     // ~2 s
        }
    }
}

template <bool DEBUG>
__global__ void __launch_bounds__(THREADS, 1)
tc_top3_kernel(const __grid_constant__ CUtensorMap map_scratch, const __grid_constant__ CUtensorMap map_store,
               const TcUnit* __restrict__ units, int nunits_host, const FusedArgs* __restrict__ fargs,
               uint32_t* __restrict__ work_counter, PartialRec* __restrict__ recs, float* __restrict__ dump) {
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B operands need 1024-byte alignment
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t bar_base = smem_base + SMEM_BAR;
    // mbarriers, 8 bytes each
    const uint32_t BAR_QFULL = bar_base;                    // [4] query K-chunk landed
    const uint32_t BAR_QEMPTY = bar_base + 32;              // [4] query K-chunk no longer read by any MMA
    const uint32_t BAR_FULL = bar_base + 64;                // [STAGES] train chunk landed
    const uint32_t BAR_EMPTY = bar_base + 96;               // [STAGES] train chunk consumed
    const uint32_t BAR_TFULL = bar_base + 128;              // [2] accumulator stage complete
    const uint32_t BAR_TEMPTY = bar_base + 144;             // [2] accumulator stage drained
    const uint32_t BAR_UFULL = bar_base + 160;              // [2] unit descriptor published
    const uint32_t BAR_UEMPTY = bar_base + 176;             // [2] unit descriptor released by MMA + epilogue
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_gen + SMEM_BAR + 192);
    TcUnit* unit_ring = reinterpret_cast<TcUnit*>(smem_gen + SMEM_BAR + 256);     // [2]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    pdl_launch_dependents();
    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&map_scratch);
        prefetch_tensormap(&map_store);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < NCHUNK; i++) {
            mbar_init(BAR_QFULL + 8 * i, 1);
            mbar_init(BAR_QEMPTY + 8 * i, 1);
        }
        for (int i = 0; i < STAGES; i++) {
            mbar_init(BAR_FULL + 8 * i, 1);
            mbar_init(BAR_EMPTY + 8 * i, 1);
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(BAR_TFULL + 8 * i, 1);
            mbar_init(BAR_TEMPTY + 8 * i, 8);          // one arrive per epilogue warp
            mbar_init(BAR_UFULL + 8 * i, 1);
            mbar_init(BAR_UEMPTY + 8 * i, 9);          // MMA thread + 8 epilogue warps
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(smem_base + SMEM_BAR + 192, TMEM_COLS);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                 // barriers, TMEM and tensor maps are set up; now the inputs must be complete

    // Persistent CTA: units are handed out by an atomic counter in launch order (range-major,
    // query-tile-minor), so the CTAs running together always work on neighbouring train rows:
    // the database streams from DRAM once and every SM stays busy until the list is empty.
    if (warp == 0) {
        // ===== scheduler + TMA producer =====
        if (lane == 0) {
            int slot = 0;
            uint32_t ph = 0;
            const int nunits = nunits_host;
            // The queue is read one unit ahead: the atomic and the descriptor load of unit i+1 are in flight
            // while unit i streams, so a unit boundary costs no L2 round trips.  When the first-pass list of
            // a compact loop search is exhausted the scheduler turns to the redo queue (take_redo), which it
            // may only wait on AFTER the loads of its current unit are out: that unit's epilogue must be able
            // to finish, or main_done never completes.
            bool phase_a = true;
            int pending_idx = -1;
            TcUnit u;
            bool have = false;
            {
                // the first unit of a CTA is its own index (grid <= units): no atomic round trip before the first
                // load of a short call; the queue hands out gridDim.x, gridDim.x + 1, ...
                const int i = (int)blockIdx.x;
                if (i < nunits) { u = units[i]; have = true; }
                else { phase_a = false; if (fargs) have = take_redo(fargs, u); }
            }
            for (uint32_t ui = 0;; ui++) {
                const int us = ui & 1;
                mbar_wait(BAR_UEMPTY + 8 * us, ((ui >> 1) & 1) ^ 1);
                if (!have) u.t_count = 0;                                    // t_count 0 = no more work
                unit_ring[us] = u;
                mbar_arrive(BAR_UFULL + 8 * us);                             // release: publishes the slot
                if (!have) break;
                // a claimed first-pass unit that has not been scheduled yet (result first used after tile 0 is issued)
                if (pending_idx < 0 && phase_a) {
                    pending_idx = (int)(gridDim.x + atomicAdd(work_counter, 1u));
                    if (pending_idx >= nunits) { pending_idx = -1; phase_a = false; }
                }
                // the redo queue's counters: loaded here, looked at after tile 0 is issued (latency hidden)
                uint32_t redo_cnt = 0, redo_head = 0;
                if (fargs) {
                    redo_cnt = ld_relaxed(&fargs->ctl->produced);
                    redo_head = ld_relaxed(&fargs->ctl->head);
                }
                TcUnit u_next;
                bool have_next = false;
                const CUtensorMap* mq = (u.maps & 1) ? &map_store : &map_scratch;
                const CUtensorMap* mt = (u.maps & 2) ? &map_store : &map_scratch;
                const int ntiles = (u.t_count + TILE_N - 1) / TILE_N;
                for (int n = 0; n < ntiles; n++) {
                    const int row = u.t_row + n * TILE_N;
                    if (u.prefetch && n + L2_AHEAD < ntiles + u.prefetch - 1) {
                        // pull a tile several iterations ahead into L2 (also past the end of this unit,
                        // for whoever takes the next one): the smem ring covers an L2 hit, not a DRAM miss
                        const int prow = row + L2_AHEAD * TILE_N;
                        for (int c = 0; c < NCHUNK; c++) {
                            tma_prefetch_l2_2d(mt, c * KCHUNK, prow);
                            tma_prefetch_l2_2d(mt, c * KCHUNK, prow + TILE_N / 2);
                        }
                    }
                    for (int c = 0; c < NCHUNK; c++) {
                        if (n == 0) {
                            // this unit's query K-chunk c, as soon as the previous unit's MMAs left it
                            mbar_wait(BAR_QEMPTY + 8 * c, (ui & 1) ^ 1);
                            mbar_expect_tx(BAR_QFULL + 8 * c, Q_SUB_BYTES);
                            tma_load_2d(smem_base + SMEM_Q + c * Q_SUB_BYTES, mq, c * KCHUNK, u.q_row, BAR_QFULL + 8 * c);
                        }
                        mbar_wait(BAR_EMPTY + 8 * slot, ph ^ 1);
                        if (DEBUG && u.dump >= 2 && c == 0 && n < 4096) reinterpret_cast<long long*>(dump)[4096 + n] = clock64();
                        mbar_expect_tx(BAR_FULL + 8 * slot, T_STAGE_BYTES);
                        const uint32_t dst = smem_base + SMEM_T + slot * T_STAGE_BYTES;
                        tma_load_2d(dst, mt, c * KCHUNK, row, BAR_FULL + 8 * slot);
                        tma_load_2d(dst + T_STAGE_BYTES / 2, mt, c * KCHUNK, row + TILE_N / 2, BAR_FULL + 8 * slot);
                        if (++slot == STAGES) { slot = 0; ph ^= 1; }
                    }
                    if (n == 0) {
                        // Redo units go FIRST (a compact loop search): taken as soon as they exist, they overlap
                        // the first pass instead of forming a tail of a few busy SMs after it.
                        int redo = -1;
                        if (fargs) {
                            const uint32_t cnt = min(redo_cnt, fargs->unit2_cap);
                            if (redo_head < cnt && atomicCAS(&fargs->ctl->head, redo_head, redo_head + 1u) == redo_head) redo = (int)redo_head;
                        }
                        if (redo >= 0) {
                            while (ld_relaxed(fargs->ready2 + redo) == 0u) {}                    // written right after the reservation
                            __threadfence();
                            u_next = load_unit_cg(fargs->units2 + redo);
                            have_next = true;                                                    // pending_idx keeps its unit for the next round
                        } else if (pending_idx >= 0) {
                            u_next = units[pending_idx];                                         // lands while the other tiles stream
                            pending_idx = -1;
                            have_next = true;
                        }
                    }
                }
                if (!have_next && !phase_a && pending_idx < 0 && fargs) have_next = take_redo(fargs, u_next);
                u = u_next;
                have = have_next;
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            int slot = 0;
            uint32_t ph = 0, tile_it = 0;
            for (uint32_t ui = 0;; ui++) {
                const int us = ui & 1;
                mbar_wait(BAR_UFULL + 8 * us, (ui >> 1) & 1);
                const int t_count = unit_ring[us].t_count;
                const int dbg = DEBUG ? unit_ring[us].dump : 0;
                mbar_arrive(BAR_UEMPTY + 8 * us);
                if (t_count == 0) break;
                const int ntiles = (t_count + TILE_N - 1) / TILE_N;
                for (int n = 0; n < ntiles; n++, tile_it++) {
                    const int st = tile_it & 1;
                    mbar_wait(BAR_TEMPTY + 8 * st, ((tile_it >> 1) & 1) ^ 1);
                    tcgen05_fence_after();
                    const long long t_issue = DEBUG ? clock64() : 0;
                    long long t_wait = 0;
                    const uint32_t d_tmem = tmem_base + st * TILE_N;
                    for (int c = 0; c < NCHUNK; c++) {
                        if (n == 0) mbar_wait(BAR_QFULL + 8 * c, ui & 1);
                        mbar_wait(BAR_FULL + 8 * slot, ph);
                        tcgen05_fence_after();
                        const uint64_t a0 = umma_smem_desc(smem_base + SMEM_Q + c * Q_SUB_BYTES);
                        const uint64_t b0 = umma_smem_desc(smem_base + SMEM_T + slot * T_STAGE_BYTES);
#pragma unroll
                        for (int k = 0; k < KCHUNK / 16; k++)            // 32 bytes = 2 address units per K=16
                            umma_bf16(d_tmem, a0 + 2 * k, b0 + 2 * k, IDESC, (c | k) != 0);
                        umma_commit(BAR_EMPTY + 8 * slot);               // frees the smem stage
                        if (n == ntiles - 1) umma_commit(BAR_QEMPTY + 8 * c);   // last reader of this query chunk
                        if (++slot == STAGES) { slot = 0; ph ^= 1; }
                    }
                    umma_commit(BAR_TFULL + 8 * st);                     // accumulator stage complete
                    if (DEBUG && dbg >= 2 && n < 4096)     // clock at issue | cycles spent waiting for train chunks
                        reinterpret_cast<long long*>(dump)[8192 + n] = (t_issue & 0xFFFFFFFFFFll) | (t_wait << 40);
                }
            }
        }
    } else if (warp >= EPI_WARP0) {
        // ===== epilogue =====
        const int half = (warp - EPI_WARP0) >> 2;
        const int quarter = warp & 3;                              // TMEM lanes this warp may read
        const int row = quarter * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + half * HALF_N;
        uint32_t tile_it = 0;
        for (uint32_t ui = 0;; ui++) {
            const int us = ui & 1;
            mbar_wait(BAR_UFULL + 8 * us, (ui >> 1) & 1);
            const TcUnit u = unit_ring[us];
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR_UEMPTY + 8 * us);
            if (u.t_count == 0) break;
            const int ntiles = (u.t_count + TILE_N - 1) / TILE_N;
            const bool row_valid = row < u.q_valid;
            if (u.maps & 4) {
                // ----- maxima-only unit (per-keyframe ratio search): the record of a slice is
                // (max, second of the group maxima, -inf, -inf); select_kernel either proves from
                // them that the ratio test fails or has the slice re-scanned exactly
                float g0 = -INFINITY, g1 = -INFINITY;
                int seg = 0, seg_tile = 0;
                const bool fused = (u.maps & 8) != 0;
                for (int n = 0; n < ntiles; n++, tile_it++) {
                    const int st = tile_it & 1;
                    mbar_wait(BAR_TFULL + 8 * st, (tile_it >> 1) & 1);
                    tcgen05_fence_after();
                    const uint32_t taddr = lane_addr + st * TILE_N;
                    const int32_t ucol = n * TILE_N + half * HALF_N;
                    const bool full_tile = (n + 1) * TILE_N <= u.t_count;
                    uint32_t ra[32], rb[32];
                    tmem_ld32(taddr, ra);
                    tmem_ld_wait(ra);
                    tmem_ld32(taddr + 32, rb);
                    if (!full_tile) mask32(ra, ucol, u.t_count);
                    scan32_max2(g0, g1, ra);
                    tmem_ld_wait(rb);
                    tmem_ld32(taddr + 64, ra);
                    if (!full_tile) mask32(rb, ucol + 32, u.t_count);
                    scan32_max2(g0, g1, rb);
                    tmem_ld_wait(ra);
                    tmem_ld32(taddr + 96, rb);
                    if (!full_tile) mask32(ra, ucol + 64, u.t_count);
                    scan32_max2(g0, g1, ra);
                    tmem_ld_wait(rb);
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(BAR_TEMPTY + 8 * st);
                    if (!full_tile) mask32(rb, ucol + 96, u.t_count);
                    scan32_max2(g0, g1, rb);
                    if (!fused && (++seg_tile == u.seg_tiles || n == ntiles - 1)) {
                        if (row_valid) {
                            const float4 rec = make_float4(g0, g1, -INFINITY, -INFINITY);
                            *reinterpret_cast<float4*>(recs + u.rec_base + (int64_t)row * u.rec_stride + seg * 2 + half) = rec;
                        }
                        seg++;
                        seg_tile = 0;
                        g0 = g1 = -INFINITY;
                    }
                }
                if (fused) {
                    // ----- the unit was one whole keyframe: decide here whether the ratio test can pass at all.
                    // The two warps that own a row's column halves meet through shared memory (one named
                    // barrier per quarter, two buffers alternating by unit); g0 is the exact maximum of the
                    // keyframe, g1 a value of ANOTHER row -- the a0 / a1 of select_kernel's ratio-only test.
                    float2* xb = reinterpret_cast<float2*>(smem_gen + SMEM_XCHG) + (ui & 1) * TILE_M;
                    if (half == 1) xb[row] = make_float2(g0, g1);
                    asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
                    if (half == 0) {
                        const float2 o = xb[row];
                        const float a0 = fmaxf(g0, o.x);
                        const float a1 = fmaxf(fminf(g0, o.x), fmaxf(g1, o.y));
                        bool open = row_valid;
                        if (row_valid && a1 > VALID_FLOOR) {
                            const float qn2 = __ldg(u.q_n2 + row);
                            float tmin2, tmax2;
                            stats_read(u.t_stats, tmin2, tmax2);
                            const float margin = dot_margin(qn2, tmin2, tmax2);
                            const float lo0 = qn2 + tmin2 - 2.f * (a0 + margin);
                            const float hi1 = qn2 + tmax2 - 2.f * (a1 - margin);
                            if (hi1 > 0.f && lo0 >= u.skip_ratio2 * hi1) open = false;
                        }
                        const unsigned m = __ballot_sync(0xffffffffu, open);
                        if (lane == 0) u.hint[u.rec_base + quarter] = m;
                        if (fargs) {
                            // Open pairs in this quarter: number them, and push the unit back into this
                            // kernel's queue as a top-4 unit (the second pass runs inside the same launch, on
                            // whichever CTA gets to it; only this quarter's 32 rows matter to it).
                            if (m != 0u) {
                                uint32_t u2 = 0, base = 0;
                                if (lane == 0) {
                                    u2 = atomicAdd(&fargs->ctl->produced, 1u);
                                    base = atomicAdd(fargs->counters + 5, (uint32_t)__popc(m));
                                    fargs->word_base[u.rec_base + quarter] = base;
                                }
                                u2 = __shfl_sync(0xffffffffu, u2, 0);
                                base = __shfl_sync(0xffffffffu, base, 0);
                                const bool ok = u2 < fargs->unit2_cap;
                                if (ok) {
                                    reinterpret_cast<uint4*>(fargs->hints2 + (size_t)u2 * TILE_M)[lane] = make_uint4(0u, 0u, 0u, 0u);
                                    if (lane == 0) {
                                        TcUnit t = u;
                                        t.rec_base = (int64_t)u2 * TILE_M * 2;
                                        t.rec_stride = 2;
                                        t.seg_tiles = 64;                       // one slice segment: the whole keyframe
                                        t.maps = 2;                             // train rows in the store; top-4 records
                                        t.prefetch = 1;
                                        t.hint = fargs->hints2 + (size_t)u2 * TILE_M;
                                        fargs->units2[u2] = t;
                                    }
                                    __threadfence();
                                    __syncwarp();
                                    if (lane == 0) *reinterpret_cast<volatile uint32_t*>(fargs->ready2 + u2) = 1u;
                                } else if (lane == 0) {
                                    fargs->counters[7] = 1u;                     // queue full: the call repeats on the record path
                                }
                                if ((m >> lane) & 1u) {
                                    const uint32_t p = base + __popc(m & ((1u << lane) - 1u));
                                    if (p < fargs->pair_cap) {
                                        PairRef r = {u.q_row + row, u.t_index0, ok ? (int32_t)u2 : -1, 0};
                                        fargs->pair_ref[p] = r;
                                    } else {
                                        fargs->counters[7] = 1u;
                                    }
                                }
                                __threadfence();                 // the pushes above, before this quarter counts as finished
                            }
                            __syncwarp();
                            if (lane == 0) atomicAdd(&fargs->ctl->main_done, 1u);
                        }
                    }
                }
                continue;
            }
            if (u.maps & 16) {
                // ----- append unit: one slice per column half over the whole range of the unit
                AppendState a;
                a.g0 = a.g1 = a.G = a.published = -INFINITY;
                a.cnt = 0;
                {
                    const float qn2 = row_valid ? __ldg(u.q_n2 + row) : 0.f;
                    float tmin2, tmax2;
                    stats_read(u.t_stats, tmin2, tmax2);
                    a.margin2 = 2.f * dot_margin(qn2, tmin2, tmax2);
                }
                // records of this warp's 32 rows (this column half); rows past q_valid store nothing
                const int64_t row_floats = (int64_t)u.rec_stride * 4;
                float* rec0 = reinterpret_cast<float*>(recs + u.rec_base + (int64_t)(quarter * 32) * u.rec_stride + half * APPEND_RECS);
                const int rows_valid = u.q_valid - quarter * 32;                 // lanes below this own a real query row
                float* stage = reinterpret_cast<float*>(smem_gen + SMEM_APPEND + (warp - EPI_WARP0) * APPEND_WARP_BYTES);
                volatile uint32_t* hint = u.hint + (row_valid ? row : 0);
                uint32_t h_next = *hint;
                for (int n = 0; n < ntiles; n++, tile_it++) {
                    const int st = tile_it & 1;
                    const uint32_t h = h_next;
                    mbar_wait(BAR_TFULL + 8 * st, (tile_it >> 1) & 1);
                    tcgen05_fence_after();
                    if (h != 0u) a.G = fmaxf(a.G, dec_ordered(h));
                    const uint32_t taddr = lane_addr + st * TILE_N;
                    const int32_t ucol = n * TILE_N + half * HALF_N;
                    const uint32_t scol = (uint32_t)n * HALF_N;
                    const bool full_tile = (n + 1) * TILE_N <= u.t_count;
                    uint32_t ra[32], rb[32];
                    tmem_ld32(taddr, ra);
                    tmem_ld_wait(ra);
                    tmem_ld32(taddr + 32, rb);
                    if (!full_tile) mask32(ra, ucol, u.t_count);
                    append32(a, ra, scol, stage, rec0, row_floats, rows_valid, lane);
                    tmem_ld_wait(rb);
                    tmem_ld32(taddr + 64, ra);
                    if (!full_tile) mask32(rb, ucol + 32, u.t_count);
                    append32(a, rb, scol + 32, stage, rec0, row_floats, rows_valid, lane);
                    tmem_ld_wait(ra);
                    tmem_ld32(taddr + 96, rb);
                    if (!full_tile) mask32(ra, ucol + 64, u.t_count);
                    append32(a, ra, scol + 64, stage, rec0, row_floats, rows_valid, lane);
                    tmem_ld_wait(rb);
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(BAR_TEMPTY + 8 * st);
                    h_next = *hint;
                    if (!full_tile) mask32(rb, ucol + 96, u.t_count);
                    append32(a, rb, scol + 96, stage, rec0, row_floats, rows_valid, lane);
                    const float L = fmaxf(a.G, a.g1);
                    if (row_valid && L > a.published) {
                        atomicMax(const_cast<uint32_t*>(hint), enc_ordered(L));
                        a.published = L;
                    }
                }
                if (row_valid) rec0[(int64_t)lane * row_floats] = __uint_as_float(a.cnt);
                continue;
            }
            if (u.maps & 32) {
                // ----- tile top-2 unit: one record per (query, tile, column half) = the exact two largest keys
                float sc;
                {
                    const float qn2 = row_valid ? __ldg(u.q_n2 + row) : 1.f;
                    float tmin2, tmax2;
                    stats_read(u.t_stats, tmin2, tmax2);
                    sc = t2_scale(qn2, tmax2);
                }
                // one PartialRec per (query, tile): {H, L of column half 0, H, L of column half 1}
                PartialRec* rec = recs + u.rec_base + (int64_t)row * u.rec_stride;
                for (int n = 0; n < ntiles; n++, tile_it++) {
                    const int st = tile_it & 1;
                    mbar_wait(BAR_TFULL + 8 * st, (tile_it >> 1) & 1);
                    tcgen05_fence_after();
                    const uint32_t taddr = lane_addr + st * TILE_N;
                    const int32_t ucol = n * TILE_N + half * HALF_N;
                    const bool full_tile = (n + 1) * TILE_N <= u.t_count;
                    float H = -INFINITY, L = -INFINITY;
                    uint32_t ra[32], rb[32];
                    tmem_ld32(taddr, ra);
                    // the four chunks as a loop of two double steps (not unrolled: 12 KB of straight-line code made
                    // instruction fetch a stall reason of its own)
#pragma unroll 1
                    for (int c = 0; c < 4; c += 2) {
                        tmem_ld_wait(ra);
                        tmem_ld32(taddr + 32 * c + 32, rb);
                        if (!full_tile && ucol + 32 * c + 32 > u.t_count) mask32(ra, ucol + 32 * c, u.t_count);
                        t2_step(H, L, ra, sc, (float)(32 * c) * 1.1920928955078125e-7f);
                        tmem_ld_wait(rb);
                        if (c == 0) {
                            tmem_ld32(taddr + 64, ra);
                        } else {
                            // all TMEM reads of this stage are complete: hand it back to the MMA warp
                            tcgen05_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(BAR_TEMPTY + 8 * st);
                        }
                        if (!full_tile && ucol + 32 * c + 64 > u.t_count) mask32(rb, ucol + 32 * c + 32, u.t_count);
                        t2_step(H, L, rb, sc, (float)(32 * c + 32) * 1.1920928955078125e-7f);
                    }
                    if (row_valid) reinterpret_cast<float2*>(rec + n)[half] = make_float2(H, L);
                }
                continue;
            }
            Top3 s;
            s.b0 = s.b1 = s.b2 = s.b3 = -INFINITY;
            s.G = s.published = -INFINITY;
            {
                const float qn2 = row_valid ? __ldg(u.q_n2 + row) : 0.f;
                float tmin2, tmax2;
                stats_read(u.t_stats, tmin2, tmax2);
                s.margin2 = 2.f * dot_margin(qn2, tmin2, tmax2);
            }
            top3_update_thr(s);
            volatile uint32_t* hint = u.hint + (row_valid ? row : 0);
            int seg = 0, seg_tile = 0;
            // bound published by the other CTAs / warps working on the same query (a second-best
            // of any subset of the train set is a lower bound on the global second-best); the load
            // for tile n+1 is issued during tile n's last scan, so its latency is never exposed
            uint32_t h_next = *hint;
            for (int n = 0; n < ntiles; n++, tile_it++) {
                const int st = tile_it & 1;
                const uint32_t h = h_next;
                mbar_wait(BAR_TFULL + 8 * st, (tile_it >> 1) & 1);
                tcgen05_fence_after();
                if (DEBUG && u.dump >= 2 && threadIdx.x == EPI_WARP0 * 32 && n < 4096)
                    reinterpret_cast<long long*>(dump)[n] = clock64();
                if (h != 0u) { s.G = fmaxf(s.G, dec_ordered(h)); top3_update_thr(s); }
                const uint32_t taddr = lane_addr + st * TILE_N;
                const int32_t ucol = n * TILE_N + half * HALF_N;           // unit-relative column of chunk 0
                const uint32_t scol = (uint32_t)seg_tile * HALF_N;         // slice-relative
                const bool full_tile = (n + 1) * TILE_N <= u.t_count;
                uint32_t ra[32], rb[32];
                int slow = 0;
                tmem_ld32(taddr, ra);
                tmem_ld_wait(ra);
                tmem_ld32(taddr + 32, rb);
                if (DEBUG && u.dump == 1 && n == 0)
                    for (int e = 0; e < 32; e++) dump[row * TILE_N + half * HALF_N + e] = __uint_as_float(ra[e]);
                if (!full_tile) mask32(ra, ucol, u.t_count);
                if (!(DEBUG && u.dump == 3)) scan32(s, ra, scol, DEBUG ? &slow : nullptr);
                tmem_ld_wait(rb);
                tmem_ld32(taddr + 64, ra);
                if (DEBUG && u.dump == 1 && n == 0)
                    for (int e = 0; e < 32; e++) dump[row * TILE_N + half * HALF_N + 32 + e] = __uint_as_float(rb[e]);
                if (!full_tile) mask32(rb, ucol + 32, u.t_count);
                if (!(DEBUG && u.dump == 3)) scan32(s, rb, scol + 32, DEBUG ? &slow : nullptr);
                tmem_ld_wait(ra);
                tmem_ld32(taddr + 96, rb);
                if (DEBUG && u.dump == 1 && n == 0)
                    for (int e = 0; e < 32; e++) dump[row * TILE_N + half * HALF_N + 64 + e] = __uint_as_float(ra[e]);
                if (!full_tile) mask32(ra, ucol + 64, u.t_count);
                if (!(DEBUG && u.dump == 3)) scan32(s, ra, scol + 64, DEBUG ? &slow : nullptr);
                tmem_ld_wait(rb);
                // all TMEM reads of this stage are complete: hand it back to the MMA warp
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR_TEMPTY + 8 * st);
                h_next = *hint;
                if (DEBUG && u.dump == 1 && n == 0)
                    for (int e = 0; e < 32; e++) dump[row * TILE_N + half * HALF_N + 96 + e] = __uint_as_float(rb[e]);
                if (!full_tile) mask32(rb, ucol + 96, u.t_count);
                if (!(DEBUG && u.dump == 3)) scan32(s, rb, scol + 96, DEBUG ? &slow : nullptr);
                if (DEBUG && u.dump >= 2 && threadIdx.x == EPI_WARP0 * 32 && n < 4096)   // clock | slow groups | hint seen
                    reinterpret_cast<long long*>(dump)[12288 + n] =
                        (clock64() & 0xFFFFFFFFFFll) | ((long long)slow << 40) | ((long long)(h != 0u) << 48);

                const float L = fmaxf(s.G, s.b1);
                if (row_valid && L > s.published) {
                    atomicMax(const_cast<uint32_t*>(hint), enc_ordered(L));
                    s.published = L;
                }
                if (++seg_tile == u.seg_tiles || n == ntiles - 1) {
                    if (row_valid) {
                        const float4 rec = make_float4(s.b0, s.b1, s.b2, s.b3);
                        *reinterpret_cast<float4*>(recs + u.rec_base + (int64_t)row * u.rec_stride + seg * 2 + half) = rec;
                    }
                    seg++;
                    seg_tile = 0;
                    top3_reset_slice(s);
                }
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

}  // namespace tc
}  // namespace vsm
