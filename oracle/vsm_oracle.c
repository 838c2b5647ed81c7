/*
 * vsm_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY, see vsm_oracle.h).
 *
 * Restates the arithmetic of cv::BFMatcher(NORM_L2).knnMatch(q, t, k) as
 * published in OpenCV 4.13.0 (third-party dependency of the reference, not
 * vendored under /root/reference, version unpinned there):
 *
 *   modules/core/src/norm.cpp   normL2Sqr_(const float*, const float*, int)
 *       baseline SIMD128 path: four v_float32x4 accumulators, each step
 *       t = a - b; d = d + t*t (v_muladd without FMA on the SSE3 baseline),
 *       then d0+d1+d2+d3 left to right, then v_reduce_sum = (x0+x2)+(x1+x3).
 *   modules/core/src/batch_distance.cpp   batchDistL2_ / BatchDistInvoker
 *       dist[j] = std::sqrt(normL2Sqr_(q, t_j, len)) for all j, then a
 *       strict-< insertion into the K best (ties keep the lowest index);
 *       K-buffers start at FLT_MAX / -1.
 *   modules/features2d/src/matchers.cpp   BFMatcher::knnMatchImpl
 *       one DMatch(queryIdx, trainIdx, imgIdx=0, dist) per filled slot.
 *
 * and the reference's own host-side filter loops:
 *   src/Slam.cpp:1151-1158, src/LoopCloser.cpp:54-62, src/Slam.cpp:569-574.
 *
 * check_vs_cv2.py verifies this file is bit-identical (indices and fp32
 * distance bits) to cv2 4.13.0 on seeded inputs.
 */
#include "vsm_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <emmintrin.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

#define DIM VSM_ORACLE_DIM

/* norm.cpp normL2Sqr_ baseline order.  No FMA: compiled with -ffp-contract=off
 * and explicit SSE2 mul/add so the bits match OpenCV's SSE3-baseline build. */
float vsm_oracle_l2sqr(const float* a, const float* b) {
    __m128 d0 = _mm_setzero_ps(), d1 = _mm_setzero_ps();
    __m128 d2 = _mm_setzero_ps(), d3 = _mm_setzero_ps();
    for (int j = 0; j < DIM; j += 16) {
        __m128 t0 = _mm_sub_ps(_mm_loadu_ps(a + j), _mm_loadu_ps(b + j));
        __m128 t1 = _mm_sub_ps(_mm_loadu_ps(a + j + 4), _mm_loadu_ps(b + j + 4));
        __m128 t2 = _mm_sub_ps(_mm_loadu_ps(a + j + 8), _mm_loadu_ps(b + j + 8));
        __m128 t3 = _mm_sub_ps(_mm_loadu_ps(a + j + 12), _mm_loadu_ps(b + j + 12));
        d0 = _mm_add_ps(_mm_mul_ps(t0, t0), d0);
        d1 = _mm_add_ps(_mm_mul_ps(t1, t1), d1);
        d2 = _mm_add_ps(_mm_mul_ps(t2, t2), d2);
        d3 = _mm_add_ps(_mm_mul_ps(t3, t3), d3);
    }
    __m128 s = _mm_add_ps(_mm_add_ps(_mm_add_ps(d0, d1), d2), d3);
    /* v_reduce_sum(v_float32x4): s + movehl(s) -> (x0+x2, x1+x3), then lane0+lane1 */
    __m128 h = _mm_add_ps(s, _mm_movehl_ps(s, s));
    __m128 r = _mm_add_ss(h, _mm_shuffle_ps(h, h, 1));
    return _mm_cvtss_f32(r);
}

static inline float l2dist(const float* a, const float* b) {
    return sqrtf(vsm_oracle_l2sqr(a, b));
}

/* BatchDistInvoker's K-best insertion (strict <, stable). */
static inline void knn_insert(float d, int64_t j, int k, int64_t* idx, float* dist) {
    if (d < dist[k - 1]) {
        int p;
        for (p = k - 2; p >= 0 && dist[p] > d; p--) {
            idx[p + 1] = idx[p];
            dist[p + 1] = dist[p];
        }
        idx[p + 1] = j;
        dist[p + 1] = d;
    }
}

static int pick_threads(int threads) {
    if (threads <= 0) {
        long n = sysconf(_SC_NPROCESSORS_ONLN);
        threads = n > 0 ? (int)n : 1;
    }
    return threads > 256 ? 256 : threads;
}

/* cv::parallel_for_ over query rows (BatchDistInvoker), here with pthreads and a
 * shared chunk counter. */
typedef struct {
    const float* q; int nq; int64_t stride_q;
    const float* t; int64_t nt; int64_t stride_t;
    int k; int64_t* idx; float* dist;
    atomic_int next;
} knn_job;

static void* knn_worker(void* arg) {
    knn_job* job = (knn_job*)arg;
    const int chunk = 4;
    for (;;) {
        int i0 = atomic_fetch_add(&job->next, chunk);
        if (i0 >= job->nq) break;
        int i1 = i0 + chunk < job->nq ? i0 + chunk : job->nq;
        for (int i = i0; i < i1; i++) {
            int k = job->k;
            int64_t* ii = job->idx + (int64_t)i * k;
            float* dd = job->dist + (int64_t)i * k;
            for (int p = 0; p < k; p++) { ii[p] = -1; dd[p] = FLT_MAX; }
            const float* qi = job->q + (int64_t)i * job->stride_q;
            for (int64_t j = 0; j < job->nt; j++)
                knn_insert(l2dist(qi, job->t + j * job->stride_t), j, k, ii, dd);
        }
    }
    return NULL;
}

void vsm_oracle_knn(const float* q, int nq, int64_t stride_q,
                    const float* t, int64_t nt, int64_t stride_t,
                    int k, int64_t* idx, float* dist, int threads) {
    threads = pick_threads(threads);
    knn_job job = {q, nq, stride_q, t, nt, stride_t, k, idx, dist, 0};
    if (threads > nq) threads = nq > 0 ? nq : 1;
    if (threads <= 1) { knn_worker(&job); return; }
    pthread_t th[256];
    int started = 0;
    for (int i = 0; i < threads - 1; i++)
        if (pthread_create(&th[started], NULL, knn_worker, &job) == 0) started++;
    knn_worker(&job);
    for (int i = 0; i < started; i++) pthread_join(th[i], NULL);
}

int vsm_oracle_match_features(const float* q, int nq, const float* t, int nt,
                              float ratio, int mutual,
                              vsm_oracle_dmatch* good, int* n_good,
                              vsm_oracle_dmatch* raw, int* n_raw, int threads) {
    *n_good = 0;
    if (n_raw) *n_raw = 0;
    if (nq <= 0 || nt <= 0) return 0;               /* Slam.cpp:1143 */
    int64_t* idx = (int64_t*)malloc(sizeof(int64_t) * 2 * (size_t)nq);
    float* dist = (float*)malloc(sizeof(float) * 2 * (size_t)nq);
    int64_t* back = NULL;
    float* backd = NULL;
    vsm_oracle_knn(q, nq, DIM, t, nt, DIM, 2, idx, dist, threads);
    if (mutual) {
        back = (int64_t*)malloc(sizeof(int64_t) * (size_t)nt);
        backd = (float*)malloc(sizeof(float) * (size_t)nt);
        vsm_oracle_knn(t, nt, DIM, q, nq, DIM, 1, back, backd, threads);
    }
    for (int i = 0; i < nq; i++) {
        if (idx[2 * i + 1] < 0) continue;            /* m.size() >= 2, Slam.cpp:1152 */
        vsm_oracle_dmatch m = {i, (int32_t)idx[2 * i], 0, dist[2 * i]};
        if (raw && n_raw) raw[(*n_raw)++] = m;      /* Slam.cpp:1153 */
        float thr = ratio * dist[2 * i + 1];         /* fp32 product, Slam.cpp:1154 */
        if (m.distance < thr) {
            if (mutual && back[m.trainIdx] != i) continue;
            good[(*n_good)++] = m;
        }
    }
    free(idx); free(dist); free(back); free(backd);
    return 0;
}

void vsm_oracle_segmented(const float* q, int nq, const float* db,
                          const int64_t* seg_off, int nseg, float ratio,
                          int32_t* counts, vsm_oracle_dmatch* matches, int threads) {
    int64_t* idx = (int64_t*)malloc(sizeof(int64_t) * 2 * (size_t)(nq > 0 ? nq : 1));
    float* dist = (float*)malloc(sizeof(float) * 2 * (size_t)(nq > 0 ? nq : 1));
    for (int s = 0; s < nseg; s++) {
        int64_t n = seg_off[s + 1] - seg_off[s];
        counts[s] = 0;
        if (n <= 0 || nq <= 0) continue;             /* LoopCloser.cpp:45 */
        vsm_oracle_knn(q, nq, DIM, db + seg_off[s] * DIM, n, DIM, 2, idx, dist, threads);
        for (int i = 0; i < nq; i++) {
            if (idx[2 * i + 1] < 0) continue;        /* LoopCloser.cpp:57 */
            if (dist[2 * i] < ratio * dist[2 * i + 1]) {
                if (matches) {
                    vsm_oracle_dmatch m = {i, (int32_t)idx[2 * i], s, dist[2 * i]};
                    matches[(int64_t)s * nq + counts[s]] = m;
                }
                counts[s]++;
            }
        }
    }
    free(idx); free(dist);
}

void vsm_oracle_merge_top2(const int64_t* idx_in, const float* dist_in, int nshard,
                           int nq, int64_t* idx_out, float* dist_out) {
    for (int i = 0; i < nq; i++) {
        int64_t bi[2] = {-1, -1};
        float bd[2] = {FLT_MAX, FLT_MAX};
        for (int s = 0; s < nshard; s++) {
            for (int p = 0; p < 2; p++) {
                int64_t j = idx_in[((int64_t)s * nq + i) * 2 + p];
                float d = dist_in[((int64_t)s * nq + i) * 2 + p];
                if (j < 0) continue;
                /* order by (distance, global index): equal to one pass over the
                 * concatenated DB with strict-< insertion. */
                for (int r = 0; r < 2; r++) {
                    if (bi[r] < 0 || d < bd[r] || (d == bd[r] && j < bi[r])) {
                        for (int m = 1; m > r; m--) { bi[m] = bi[m - 1]; bd[m] = bd[m - 1]; }
                        bi[r] = j; bd[r] = d;
                        break;
                    }
                }
            }
        }
        idx_out[2 * i] = bi[0]; idx_out[2 * i + 1] = bi[1];
        dist_out[2 * i] = bd[0]; dist_out[2 * i + 1] = bd[1];
    }
}

/* ---- deterministic generator (integer arithmetic; numpy twin in gen.py) ---- */
static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

static inline int irwin_hall8(uint64_t x) {
    int s = 0;
    for (int b = 0; b < 8; b++) s += (int)((x >> (8 * b)) & 0xFF);
    return s - 1020;
}

void vsm_oracle_gen_rows(uint64_t seed, uint64_t set_id, int64_t row0, int64_t n, float* out) {
    uint64_t key = splitmix64(splitmix64(seed) ^ (set_id * 0xD1342543DE82EF95ULL));
    for (int64_t r = 0; r < n; r++) {
        int v[DIM];
        int64_t n2 = 0;
        for (int c = 0; c < DIM; c++) {
            uint64_t ctr = (uint64_t)(row0 + r) * DIM + (uint64_t)c;
            v[c] = irwin_hall8(splitmix64(key + ctr * 0x9E3779B97F4A7C15ULL));
            n2 += (int64_t)v[c] * v[c];
        }
        double inv = 1.0 / sqrt((double)n2);
        for (int c = 0; c < DIM; c++) out[r * DIM + c] = (float)((double)v[c] * inv);
    }
}
