// vsm_tc2.cuh -- the tensor-core pass on CTA PAIRS (tcgen05 cta_group::2, thread-block cluster of 2).
//
// Same algorithm and epilogue as vsm_tc.cuh; what changes is who feeds the tensor cores.
// With one CTA per SM an M=128 x N=256 x K=16 MMA reads 12 KB of operands from shared memory
// per 128 cycles (96 B/cycle) while TMA refills the ring at 64 B/cycle: together more than the
// 128 B/cycle a shared-memory port delivers, and ncu shows the tensor pipe stuck near 89 %.
// A pair of SMs working on one M=256 x N=256 tile halves the train-row traffic of each:
//   * the two CTAs of a cluster hold two DIFFERENT 128-query tiles (their own A operand) and
//     the same train range; each loads only its HALF (128 of 256 rows) of every train K-chunk;
//   * the leader CTA's MMA thread issues tcgen05.mma.cta_group::2 (M=256): each SM multiplies
//     its 128 queries by all 256 train rows, fetching the other half of B from its partner's
//     shared memory; each SM's TMEM receives its own 128 x 256 accumulator;
//   * per SM and 128 MMA cycles: 8 KB of operand reads + 4 KB of TMA refill -> 96 B/cycle,
//     and the same 128 KB ring now holds two tiles in flight instead of one.
// Synchronisation across the pair: TMA completions of BOTH CTAs land on the LEADER's `full`
// barriers (cta_group::2 loads), tcgen05.commit multicasts `empty` / `tmem_full` to both CTAs,
// and the partner's epilogue warps arrive remotely on the leader's `tmem_empty`.
// Units are taken round-robin by cluster (static: no cross-CTA hand-off of the queue head).
#pragma once

#include "vsm_tc.cuh"

namespace vsm {

// One work unit of the pair kernel: two query tiles x one train range.
struct TcUnit2 {
    const float*    q_n2[2];       // squared norms of each CTA's query rows
    uint32_t*       hint[2];       // shared second-best hints of each CTA's query rows
    const uint32_t* t_stats;
    int64_t rec_base[2];
    int32_t rec_stride;
    int32_t q_row[2];
    int32_t q_valid[2];            // 0 = dummy tile (odd number of query tiles): nothing is written
    int32_t t_row, t_count, t_index0;
    int32_t seg_tiles, maps, prefetch;
    int32_t pad;
};

namespace tc2 {

using namespace tc;

constexpr int STAGES2 = 8;                            // train K-chunk ring: 8 x (128 rows x 64 cols)
constexpr uint32_t T2_STAGE_BYTES = 128 * 128;        // 16 KB per CTA per stage
constexpr uint32_t SMEM2_T = NCHUNK * Q_SUB_BYTES;    // 65536
constexpr uint32_t SMEM2_BAR = SMEM2_T + STAGES2 * T2_STAGE_BYTES;   // 196608
constexpr uint32_t SMEM2_BYTES = SMEM2_BAR + 1024 + 1024;
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;           // clears the CTA-rank bit of a shared::cluster address
constexpr uint32_t IDESC2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t ncluster_x() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    __syncwarp();                                        // roles run on single lanes: re-converge first
    asm volatile("barrier.cluster.arrive.release;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
// bounded wait with cluster-scope acquire (barriers here are also signalled from the partner CTA)
__device__ __forceinline__ void mbar_wait_cl(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    long long t0 = 0;
    for (uint32_t it = 0;; it++) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
        if ((it & 1023u) == 1023u) {
            long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000LL) __trap();
        }
    }
}
// arrive on the LEADER's copy of a barrier (local for rank 0, remote for rank 1)
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar & PEER_MASK) : "memory");
}
// this CTA's half of a train chunk / its query chunk; the bytes are counted on the LEADER's barrier
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, int32_t c0, int32_t c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar & PEER_MASK)
        : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives on the barrier at this offset in BOTH CTAs once all MMAs issued so far have completed
__device__ __forceinline__ void umma2_commit_both(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
tc_top3_pair_kernel(const __grid_constant__ CUtensorMap map_scratch, const __grid_constant__ CUtensorMap map_store,
                    const TcUnit2* __restrict__ units, int nunits, PartialRec* __restrict__ recs) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t bar_base = smem_base + SMEM2_BAR;
    const uint32_t BAR_QFULL = bar_base;                    // [4]  used in the leader
    const uint32_t BAR_QEMPTY = bar_base + 32;              // [4]  both CTAs (multicast commit)
    const uint32_t BAR_FULL = bar_base + 64;                // [8]  used in the leader
    const uint32_t BAR_EMPTY = bar_base + 128;              // [8]  both CTAs
    const uint32_t BAR_TFULL = bar_base + 192;              // [2]  both CTAs
    const uint32_t BAR_TEMPTY = bar_base + 208;             // [2]  used in the leader: 16 arrivals
    const uint32_t BAR_UFULL = bar_base + 224;              // [2]  CTA-local
    const uint32_t BAR_UEMPTY = bar_base + 240;             // [2]  CTA-local
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_gen + SMEM2_BAR + 256);
    TcUnit2* unit_ring = reinterpret_cast<TcUnit2*>(smem_gen + SMEM2_BAR + 512);     // [2]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const uint32_t cid = cluster_id_x(), ncl = ncluster_x();

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&map_scratch);
        prefetch_tensormap(&map_store);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < NCHUNK; i++) {
            mbar_init(BAR_QFULL + 8 * i, 1);
            mbar_init(BAR_QEMPTY + 8 * i, 1);
        }
        for (int i = 0; i < STAGES2; i++) {
            mbar_init(BAR_FULL + 8 * i, 1);
            mbar_init(BAR_EMPTY + 8 * i, 1);
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(BAR_TFULL + 8 * i, 1);
            mbar_init(BAR_TEMPTY + 8 * i, 16);         // 8 epilogue warps of each CTA
            mbar_init(BAR_UFULL + 8 * i, 1);
            mbar_init(BAR_UEMPTY + 8 * i, leader ? 9 : 8);   // (MMA thread, leader only) + 8 epilogue warps
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc2(smem_base + SMEM2_BAR + 256, TMEM_COLS);
    tcgen05_fence_before();
    cluster_sync_all();                                  // barriers of BOTH CTAs are initialised before any remote use
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== unit reader + TMA producer (every CTA loads its own half) =====
        if (lane == 0) {
            int slot = 0;
            uint32_t ph = 0;
            for (uint32_t ui = 0;; ui++) {
                const int us = ui & 1;
                const int idx = (int)(cid + ui * ncl);                       // static round-robin over the clusters
                mbar_wait(BAR_UEMPTY + 8 * us, ((ui >> 1) & 1) ^ 1);
                TcUnit2 u;
                if (idx < nunits) u = units[idx]; else u.t_count = 0;
                unit_ring[us] = u;
                mbar_arrive(BAR_UFULL + 8 * us);
                if (idx >= nunits) break;
                const CUtensorMap* mq = (u.maps & 1) ? &map_store : &map_scratch;
                const CUtensorMap* mt = (u.maps & 2) ? &map_store : &map_scratch;
                const int ntiles = (u.t_count + TILE_N - 1) / TILE_N;
                for (int n = 0; n < ntiles; n++) {
                    const int row = u.t_row + n * TILE_N;
                    if (leader && u.prefetch && n + L2_AHEAD < ntiles + u.prefetch - 1) {
                        const int prow = row + L2_AHEAD * TILE_N;
                        for (int c = 0; c < NCHUNK; c++) {
                            tma_prefetch_l2_2d(mt, c * KCHUNK, prow);
                            tma_prefetch_l2_2d(mt, c * KCHUNK, prow + TILE_N / 2);
                        }
                    }
                    for (int c = 0; c < NCHUNK; c++) {
                        if (n == 0) {
                            mbar_wait_cl(BAR_QEMPTY + 8 * c, (ui & 1) ^ 1);
                            if (leader) mbar_expect_tx(BAR_QFULL + 8 * c, 2 * Q_SUB_BYTES);
                            tma_load_2d_pair(smem_base + SMEM_Q + c * Q_SUB_BYTES, mq, c * KCHUNK, u.q_row[rank], BAR_QFULL + 8 * c);
                        }
                        mbar_wait_cl(BAR_EMPTY + 8 * slot, ph ^ 1);
                        if (leader) mbar_expect_tx(BAR_FULL + 8 * slot, 2 * T2_STAGE_BYTES);
                        tma_load_2d_pair(smem_base + SMEM2_T + slot * T2_STAGE_BYTES, mt, c * KCHUNK, row + (int)rank * 128,
                                         BAR_FULL + 8 * slot);
                        if (++slot == STAGES2) { slot = 0; ph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the leader's thread drives both SMs =====
        if (lane == 0 && leader) {
            int slot = 0;
            uint32_t ph = 0, tile_it = 0;
            for (uint32_t ui = 0;; ui++) {
                const int us = ui & 1;
                mbar_wait(BAR_UFULL + 8 * us, (ui >> 1) & 1);
                const int t_count = unit_ring[us].t_count;
                mbar_arrive(BAR_UEMPTY + 8 * us);
                if (t_count == 0) break;
                const int ntiles = (t_count + TILE_N - 1) / TILE_N;
                for (int n = 0; n < ntiles; n++, tile_it++) {
                    const int st = tile_it & 1;
                    mbar_wait_cl(BAR_TEMPTY + 8 * st, ((tile_it >> 1) & 1) ^ 1);
                    tcgen05_fence_after();
                    const uint32_t d_tmem = tmem_base + st * TILE_N;
                    for (int c = 0; c < NCHUNK; c++) {
                        if (n == 0) mbar_wait_cl(BAR_QFULL + 8 * c, ui & 1);
                        mbar_wait_cl(BAR_FULL + 8 * slot, ph);
                        tcgen05_fence_after();
                        const uint64_t a0 = umma_smem_desc(smem_base + SMEM_Q + c * Q_SUB_BYTES);
                        const uint64_t b0 = umma_smem_desc(smem_base + SMEM2_T + slot * T2_STAGE_BYTES);
#pragma unroll
                        for (int k = 0; k < KCHUNK / 16; k++)
                            umma2_bf16(d_tmem, a0 + 2 * k, b0 + 2 * k, IDESC2, (c | k) != 0);
                        umma2_commit_both(BAR_EMPTY + 8 * slot);
                        if (n == ntiles - 1) umma2_commit_both(BAR_QEMPTY + 8 * c);
                        if (++slot == STAGES2) { slot = 0; ph ^= 1; }
                    }
                    umma2_commit_both(BAR_TFULL + 8 * st);
                }
            }
        }
    } else if (warp >= EPI_WARP0) {
        // ===== epilogue (each CTA scans its own 128 queries) =====
        const int half = (warp - EPI_WARP0) >> 2;
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + half * HALF_N;
        uint32_t tile_it = 0;
        for (uint32_t ui = 0;; ui++) {
            const int us = ui & 1;
            mbar_wait(BAR_UFULL + 8 * us, (ui >> 1) & 1);
            const TcUnit2 u = unit_ring[us];
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR_UEMPTY + 8 * us);
            if (u.t_count == 0) break;
            const int ntiles = (u.t_count + TILE_N - 1) / TILE_N;
            const bool row_valid = row < u.q_valid[rank];
            Top3 s;
            s.b0 = s.b1 = s.b2 = s.b3 = -INFINITY;
            s.G = s.published = -INFINITY;
            {
                const float qn2 = row_valid ? __ldg(u.q_n2[rank] + row) : 0.f;
                float tmin2, tmax2;
                stats_read(u.t_stats, tmin2, tmax2);
                s.margin2 = 2.f * dot_margin(qn2, tmin2, tmax2);
            }
            top3_update_thr(s);
            volatile uint32_t* hint = u.hint[rank] + (row_valid ? row : 0);
            int seg = 0, seg_tile = 0;
            for (int n = 0; n < ntiles; n++, tile_it++) {
                const int st = tile_it & 1;
                const uint32_t h = *hint;
                mbar_wait_cl(BAR_TFULL + 8 * st, (tile_it >> 1) & 1);
                tcgen05_fence_after();
                if (h != 0u && row_valid) { s.G = fmaxf(s.G, dec_ordered(h)); top3_update_thr(s); }
                const uint32_t taddr = lane_addr + st * TILE_N;
                const int32_t ucol = n * TILE_N + half * HALF_N;
                const uint32_t scol = (uint32_t)seg_tile * HALF_N;
                const bool full_tile = (n + 1) * TILE_N <= u.t_count;
                uint32_t ra[32], rb[32];
                tmem_ld32(taddr, ra);
                tmem_ld_wait(ra);
                tmem_ld32(taddr + 32, rb);
                if (!full_tile) mask32(ra, ucol, u.t_count);
                scan32(s, ra, scol);
                tmem_ld_wait(rb);
                tmem_ld32(taddr + 64, ra);
                if (!full_tile) mask32(rb, ucol + 32, u.t_count);
                scan32(s, rb, scol + 32);
                tmem_ld_wait(ra);
                tmem_ld32(taddr + 96, rb);
                if (!full_tile) mask32(ra, ucol + 64, u.t_count);
                scan32(s, ra, scol + 64);
                tmem_ld_wait(rb);
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_leader(BAR_TEMPTY + 8 * st);       // the leader's MMA thread waits for 16
                if (!full_tile) mask32(rb, ucol + 96, u.t_count);
                scan32(s, rb, scol + 96);

                const float L = fmaxf(s.G, s.b1);
                if (row_valid && L > s.published) {
                    atomicMax(const_cast<uint32_t*>(hint), enc_ordered(L));
                    s.published = L;
                }
                if (++seg_tile == u.seg_tiles || n == ntiles - 1) {
                    if (row_valid) {
                        const float4 rec = make_float4(s.b0, s.b1, s.b2, s.b3);
                        *reinterpret_cast<float4*>(recs + u.rec_base[rank] + (int64_t)row * u.rec_stride + seg * 2 + half) = rec;
                    }
                    seg++;
                    seg_tile = 0;
                    top3_reset_slice(s);
                }
            }
        }
    }
    tcgen05_fence_before();
    cluster_sync_all();                                  // the partner may still be read by / signalled from this CTA
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc2(tmem_base, TMEM_COLS);
    }
}

}  // namespace tc2
}  // namespace vsm
