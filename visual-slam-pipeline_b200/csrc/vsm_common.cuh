// vsm_common.cuh -- shared device types and the canonical exact distance.
//
// Exactness contract: every distance this library REPORTS is computed by
// canon_l2sqr_halfwarp() below, which reproduces, operation for operation, the
// fp32 arithmetic of OpenCV's normL2Sqr_ baseline path (4 accumulators x 4 lanes,
// separate mul and add, ((d0+d1)+d2)+d3, then (x0+x2)+(x1+x3)) followed by a
// correctly rounded sqrt -- what cv::BFMatcher(NORM_L2) computes behind
// src/Slam.cpp:1149 / src/LoopCloser.cpp:51 of the reference.  The tensor-core
// pass only NOMINATES candidates; it never decides an index or a distance.
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <float.h>

#define VSM_DIM 256
#define VSM_TOPK 4                 // approximate entries kept per (query, slice)

namespace vsm {

// Programmatic dependent launch: the kernels of one call (prologue -> tensor-core pass -> select ->
// re-scan -> filter) are launched with cudaLaunchAttributeProgrammaticStreamSerialization.  Each
// one lets its successor be scheduled at once (pdl_launch_dependents: the successor's launch
// latency, block scheduling and set-up overlap this kernel) and touches global memory only after
// pdl_wait(), which returns when the predecessor grid has completed and its writes are visible.
// Without a programmatic predecessor both are no-ops.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Geometry of the tensor-core kernel (vsm_tc.cuh).
constexpr int TILE_M = 128;        // queries per CTA (TMEM lanes)
constexpr int TILE_N = 256;        // train rows per MMA tile (TMEM columns of one stage)
constexpr int HALF_N = 128;        // columns one epilogue warp-group owns in every tile

// Approximate per-(query, slice) record written by the tensor-core kernel: the four
// largest bf16 dot products q.t of the slice, descending, PACKED: the low 13 mantissa bits
// hold the column number inside the slice (column c = row (c/128)*256 + half*128 + c%128
// of the slice's range).  -inf = empty.
struct __align__(16) PartialRec {
    float s[VSM_TOPK];
};
constexpr uint32_t PACK_MASK = 0xFFFFE000u;
// Append records (small train sets, TcUnit maps bit 4): per (query, slice) APPEND_CAP floats -- [0] the
// number of values that passed the running threshold (as an integer), [1..] the first APPEND_CAP - 1 of
// them, packed like a PartialRec entry.  A count above APPEND_CAP - 1 = overflow: the slice is re-scanned.
constexpr int APPEND_CAP = 64;
constexpr int APPEND_RECS = APPEND_CAP * 4 / 16;          // PartialRec slots one append record occupies
constexpr float MASKED_VALUE = -3.0e38f;       // a column past the end of the train range
constexpr float VALID_FLOOR = -1.0e38f;        // record entries above this are real columns

// One CTA of the tensor-core kernel: one 128-query tile x a contiguous train range.
struct __align__(16) TcUnit {
    const float*    q_n2;          // squared norms of the tile's query rows
    const uint32_t* t_stats;       // norm statistics of the train set (see stats_read)
    int64_t rec_base;              // record index of (tile row 0, the unit's first slice)
    int32_t rec_stride;            // records per query (= slices of the problem)
    int32_t q_row;                 // first query row (tensor-map row coordinate)
    int32_t t_row;                 // first train row of the range (tensor-map row coordinate)
    int32_t t_count;               // valid train rows in the range (>= 1)
    int32_t t_index0;              // logical train index of the range's first row
    int32_t q_valid;               // valid query rows in the tile (1..128)
    int32_t seg_tiles;             // tiles per slice segment (records flushed every seg_tiles tiles)
    int32_t maps;                  // bit0: query rows in store map, bit1: train rows in store map,
                                   // bit2: maxima-only records (see scan32_max2),
                                   // bit4: APPEND records (see APPEND_CAP): every value above the running
                                   //       threshold is appended, no top-4 state -- for small train sets, where
                                   //       almost every 8-group holds a candidate for some lane of the warp
                                   // bit5: TILE TOP-2 records (small train sets of pair matching): no running state
                                   //       at all -- per (query, tile, column half) the exact two largest values with
                                   //       their column, found by a fixed comparison tree (t2_chunk); data-independent
                                   // bit3 (with bit2): FUSED ratio dismissal -- the unit covers one whole keyframe;
                                   //       no record is written, only a 128-bit mask of the queries the
                                   //       ratio test could not dismiss (rec_base = first mask word of the unit)
    int32_t dump;                  // debug: 1 = raw accumulators of tile 0, 2 = clock64 timeline
    int32_t prefetch;              // this CTA issues the L2 prefetches for its train range
    float skip_ratio2;             // bit3 units: ratio^2 * 1.001 of the caller's ratio test (Problem::skip_ratio2)
    uint32_t* hint;                // per query row: shared lower bound on the global second-best dot
                                   // (bit3 units: the open-pair mask words, uint32 [units][4])
};

// ---- compact loop search (vsm_loop_detect_compact): the second pass runs INSIDE the first pass's kernel ----
// A fused unit (maps bit 3) whose ratio test could not be dismissed for some query re-enters the same
// persistent kernel as a top-4 unit: the epilogue that found the open pairs numbers them and pushes a redo
// unit per 32-row quarter; the schedulers of all CTAs drain that queue after the first-pass list.
struct PairRef {                   // one open (query, keyframe) pair
    int32_t q, slot, unit2, pad;
};
struct RedoCtl {                   // device memory, zeroed per call
    uint32_t produced;             // redo units reserved so far (above unit2_cap: overflow, the call falls back)
    uint32_t head;                 // redo units taken by the schedulers
    uint32_t main_done;            // finished (first-pass unit, quarter) epilogues: 4 per unit, then nothing is produced any more
    uint32_t finish_ticket;        // loop_finish_kernel's last-block ticket
};
struct FusedArgs {                 // lives in the call's descriptor block
    RedoCtl* ctl;
    TcUnit* units2;                // [unit2_cap]
    uint32_t* ready2;              // [unit2_cap] 1 = units2[i] is complete (zeroed per call)
    uint32_t* hints2;              // [unit2_cap][128]
    uint32_t* word_base;           // [mask words] pair index of a word's first open bit
    PairRef* pair_ref;             // [pair_cap]
    uint32_t* counters;            // aux block as uint32: [5] open pairs, [7] overflow flag
    uint32_t unit2_cap, pair_cap;
    uint32_t n_main, pad;          // first-pass units
};

// A slice = the train rows one epilogue thread scanned for one record:
//   half >= 0: columns [half*128, half*128+128) of every 256-row tile of
//              [t_index0, t_index0 + t_count)
//   half <  0: the contiguous range itself (exact SIMT engine).
struct SliceInfo {
    int32_t t_index0;
    int32_t t_count;
    int32_t half;
    int32_t pad;
};

// {min, max} squared norm of a row set, kept so that an all-zero slot means "no rows yet":
// slot[0] = max over rows of ~bits(n2) (i.e. the inverted minimum), slot[1] = max of bits(n2)
// (non-negative floats order like their bit patterns).
__device__ __forceinline__ void stats_read(const uint32_t* slot, float& tmin2, float& tmax2) {
    tmin2 = __uint_as_float(~__ldg(slot));
    tmax2 = __uint_as_float(__ldg(slot + 1));
    if (!(tmin2 <= tmax2)) { tmin2 = 0.f; tmax2 = 0.f; }      // empty set
}

// One kNN problem (query set vs train set), consumed by the select/re-score kernel.
struct Problem {
    const float*    q_f32;         // first query row (fp32 master)
    const float*    t_f32;         // first train row (fp32 master)
    const float*    q_n2;          // squared norms of the query rows
    const uint32_t* t_stats;       // norm statistics of the train set (see stats_read)
    int64_t partial_off;           // first PartialRec of the problem
    int64_t out_off;               // outputs at out_*[(out_off + q) * 2 ...]
    int32_t nq, nt;
    int32_t nslices;
    int32_t slice_off;             // first SliceInfo of this problem
    int32_t exact;                 // bit0: no tensor-core records, scan every slice exactly;
                                   // bit1: the records hold slice maxima only (TcUnit maps bit2): a query
                                   //       that the ratio-only test cannot dismiss is re-scanned exactly
                                   // bit2: append records (TcUnit maps bit4), APPEND_RECS slots per (query, slice)
                                   // bit3: tile top-2 records (TcUnit maps bit5, see t2_scale): one slice per (tile,
                                   //       column half), each holding the EXACT two largest approximate dots
                                   // bit4 (with bit3): only the nearest neighbour's INDEX is wanted (the reverse
                                   //       problem of a mutual test, read by filter_kernel)
    // > 0: the caller only wants ratio-test survivors of this problem (no raw list; a mutual test, if
    // any, is applied on top by filter_kernel) with this ratio^2 * 1.001: a query whose approximate top-2 already PROVES the ratio test fails is
    // answered "no match" without any exact re-score (select_kernel)
    float skip_ratio2;
    float ratio;                   // tile top-2 problems with skip_ratio2 > 0: the caller's ratio itself (fp32 test d0 < ratio * d1)
    int32_t gshift;                // tile top-2 problems: log2 of the queries one warp of t2_select_kernel takes
};

// One pair for the filter kernel (match_features semantics).
struct FilterJob {
    int64_t fwd_off;               // top-2 of query->train at out_*[(fwd_off + q) * 2]
    int64_t back_off;              // top-2 of train->query (mutual only), else -1
    int64_t good_off;              // output slot offsets (in dmatch units)
    int64_t raw_off;               // -1: no raw output
    int32_t nq, nt;
    int32_t img_idx;
    int32_t back_prob;             // >= 0: the reverse problem (index into the call's Problem list) keeps tile top-2
                                   // records and left its ambiguous rows DEFERRED: filter_kernel resolves those it needs
    float   ratio;
    int32_t pad2;
};

struct DMatch {
    int32_t queryIdx, trainIdx, imgIdx;
    float distance;
};

__device__ __forceinline__ bool better(float d, int32_t i, float bd, int32_t bi) {
    // (distance, index) lexicographic == one sequential pass with strict-< insertion
    return d < bd || (d == bd && (uint32_t)i < (uint32_t)bi);
}

__device__ __forceinline__ void insert2(float d, int32_t i, float& d0, int32_t& i0, float& d1, int32_t& i1) {
    if (better(d, i, d1, i1)) {
        if (better(d, i, d0, i0)) { d1 = d0; i1 = i0; d0 = d; i0 = i; }
        else { d1 = d; i1 = i; }
    }
}

// Half-warp exact squared distance.  lane16 owns elements k = 16*i + lane16,
// i.e. OpenCV's accumulator a = lane16/4, SIMD lane l = lane16%4.
// qreg[i] = q[16*i + lane16].  Result valid in lane16 == 0 of each half-warp.
__device__ __forceinline__ float canon_l2sqr_halfwarp_regs(const float (&qreg)[16], const float (&tv)[16]) {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        float d = __fsub_rn(qreg[i], tv[i]);
        acc = __fadd_rn(__fmul_rn(d, d), acc);          // no FMA: mul, then add
    }
    const unsigned full = 0xffffffffu;
    float v1 = __shfl_down_sync(full, acc, 4, 16);
    float v2 = __shfl_down_sync(full, acc, 8, 16);
    float v3 = __shfl_down_sync(full, acc, 12, 16);
    float s = __fadd_rn(__fadd_rn(__fadd_rn(acc, v1), v2), v3);   // ((d0+d1)+d2)+d3, lanes 0..3
    float h = __fadd_rn(s, __shfl_down_sync(full, s, 2, 16));     // (x0+x2), (x1+x3)
    return __fadd_rn(h, __shfl_down_sync(full, h, 1, 16));        // lane 0
}

__device__ __forceinline__ float canon_l2sqr_halfwarp(const float (&qreg)[16],
                                                      const float* __restrict__ t, int lane16) {
    float tv[16];
#pragma unroll
    for (int i = 0; i < 16; i++) tv[i] = __ldg(t + 16 * i + lane16);
    return canon_l2sqr_halfwarp_regs(qreg, tv);
}

__device__ __forceinline__ void load_qreg(float (&qreg)[16], const float* __restrict__ q, int lane16) {
#pragma unroll
    for (int i = 0; i < 16; i++) qreg[i] = __ldg(q + 16 * i + lane16);
}

// Half-width of the interval that is guaranteed to contain the exact dot product q.t
// (and, through it, the ordering by the canonical fp32 distance) around the value the
// tensor-core pass computed from bf16-rounded operands, in dot-product units:
//   |bf16(q).bf16(t) - q.t| <= ((1+2^-9)^2 - 1) |q||t|  <  (2^-8 + 2^-17) |q||t|
//   fp32 accumulation inside the MMA and the rounding of the canonical distance are
//   covered by a 2^-10 |q||t| and the 2^-14 (|q|^2+|t|^2) term; the epilogue's index
//   packing (low 13 mantissa bits) moves a value by < 2^-10 |q||t| more;
//   ranking by the dot product alone (instead of |t|^2 - 2 q.t) is off by at most
//   (max|t|^2 - min|t|^2) / 4 around the mid norm.
// Used by BOTH the tensor-core epilogue and the select kernel (same bits).
__device__ __forceinline__ float dot_margin(float qn2, float tmin2, float tmax2) {
    float qn = sqrtf(qn2), tn = sqrtf(tmax2);
    return 0.0059f * qn * tn + 0.25f * (tmax2 - tmin2) + 6.2e-5f * (qn2 + tmax2);
}


// ---- tile top-2 records (TcUnit maps bit 5, Problem exact bit 3) -------------------------------------------
// Small train sets (a frame pair, a ragged batch: <= T2_MAX_TILES tiles) are match-heavy: with ~1000 train rows
// some lane of every warp holds a candidate in almost every group of eight values, so a threshold-driven epilogue
// runs its insert path all the time.  Here the epilogue carries no state and takes no data-dependent branch: each
// accumulator value v becomes a KEY, a float in [1, 2) whose upper 16 mantissa bits hold v quantised on a grid
// of 2^-16 / s (s = t2_scale, so that |v * s| < 0.49) and whose low 7 bits hold the column inside the tile half:
//     q   = fma(v, s, 192)            -> 192 + n * 2^-16       (the addition rounds v * s to the grid)
//     key = (q - 190.5) + col * 2^-23 -> 1.5 + n * 2^-16 + col * 2^-23, both additions exact
// Keys order like (quantised value, column); the exact top-2 of a tile half is a fixed tree of FMNMX / FMNMX3.
// The quantisation moves a value by <= 2^-17 / s < 1.7e-5 |q||t| -- far inside what dot_margin() reserves for the
// 13-bit packing of the other record kinds.
constexpr int T2_MAX_TILES = 32;
__device__ __forceinline__ float t2_scale(float qn2, float tmax2) {
    return 0.48f / fmaxf(1.01f * sqrtf(qn2) * sqrtf(tmax2), 1e-20f);
}
__device__ __forceinline__ bool t2_valid(float key) { return key >= 1.0f; }            // masked columns / empty: below 1 (or NaN)
__device__ __forceinline__ int t2_col(float key) { return (int)(__float_as_uint(key) & 127u); }
__device__ __forceinline__ float t2_value(float key, float inv_s) {                    // approximate dot product behind a key
    const int n = (int)((__float_as_uint(key) & 0x7FFFFFu) >> 7) - 32768;
    return (float)n * 1.52587890625e-5f * inv_s;
}

}  // namespace vsm
