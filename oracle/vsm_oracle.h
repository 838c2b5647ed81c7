/*
 * vsm_oracle.h -- CPU oracle for the descriptor-matching hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is on the product path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product library
 * (libvsm.so) never links, loads or calls it and has no CPU fallback.
 *
 * What it restates.  The reference (salah-dev-stu/visual-slam-pipeline) does
 * its matching by calling OpenCV (not vendored in /root/reference; "OpenCV
 * 4.x", unpinned, CMakeLists.txt:8) and then filtering on the host:
 *   - src/Slam.cpp:1140-1172      Slam::match_features (kNN k=2, ratio test)
 *   - src/LoopCloser.cpp:43-62    per-keyframe kNN k=2 + ratio + >=30 gate
 *   - src/Slam.cpp:546-574        map-point DB search, ratio 0.70
 *   - src/Slam.cpp:744-774        same, frame-range filtered
 * The exact matcher BASELINE.json names is cv::BFMatcher(NORM_L2).knnMatch.
 * Its arithmetic (OpenCV 4.13.0, modules/core/src/batch_distance.cpp and
 * norm.cpp, as published) is restated in vsm_oracle.c.
 *
 * Pinning.  The reference holds no tests or golden vectors for this path
 * ("parity unpinned" by the reference itself).  The
 * oracle is instead pinned against OpenCV itself: oracle/check_vs_cv2.py
 * compares it with cv2 4.13.0's BFMatcher (same C++ code the reference
 * links) bit for bit (indices AND fp32 distances) on seeded inputs, and
 * oracle/make_golden.py writes cv2's answers to tests/golden/.
 */
#ifndef VSM_ORACLE_H
#define VSM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VSM_ORACLE_DIM 256

/* Mirror of cv::DMatch {int queryIdx; int trainIdx; int imgIdx; float distance;} */
typedef struct {
    int32_t queryIdx;
    int32_t trainIdx;
    int32_t imgIdx;
    float distance;
} vsm_oracle_dmatch;

/* fp32 sum of squared differences over 256 floats in OpenCV's baseline-SSE
 * order (4 accumulators x 4 lanes, separate mul and add). */
float vsm_oracle_l2sqr(const float* a, const float* b);

/* kNN with k in {1,2}.  idx/dist are [nq][k]; missing neighbours (nt < k) get
 * idx = -1, dist = FLT_MAX, like cv::batchDistance's initial fill.
 * stride_q / stride_t are row strides in floats (>= 256). threads <= 0: all cores. */
void vsm_oracle_knn(const float* q, int nq, int64_t stride_q,
                    const float* t, int64_t nt, int64_t stride_t,
                    int k, int64_t* idx, float* dist, int threads);

/* Slam::match_features (src/Slam.cpp:1140-1172) for float descriptors, with
 * the ratio as a runtime argument and an optional mutual-NN filter
 * (north-star addition; composed as knn(q,t,2) AND knn(t,q,1)[train]==query).
 * good/raw must hold nq entries.  Returns 0. */
int vsm_oracle_match_features(const float* q, int nq, const float* t, int nt,
                              float ratio, int mutual,
                              vsm_oracle_dmatch* good, int* n_good,
                              vsm_oracle_dmatch* raw, int* n_raw, int threads);

/* LoopCloser::detect matching block (src/LoopCloser.cpp:43-62): for each
 * keyframe segment s (rows seg_off[s]..seg_off[s+1]) top-2 within the segment,
 * ratio test, survivors counted.  counts[s] = number of survivors.  If
 * matches != NULL, the survivors of segment s are written at
 * matches[s*nq ...] in query order (trainIdx is segment-local, imgIdx = s). */
void vsm_oracle_segmented(const float* q, int nq, const float* db,
                          const int64_t* seg_off, int nseg, float ratio,
                          int32_t* counts, vsm_oracle_dmatch* matches, int threads);

/* Merge per-shard top-2 lists into a global top-2 ordered by (distance, index).
 * idx_in/dist_in are [nshard][nq][2] with GLOBAL indices (-1 = empty). */
void vsm_oracle_merge_top2(const int64_t* idx_in, const float* dist_in, int nshard,
                           int nq, int64_t* idx_out, float* dist_out);

/* Deterministic integer-arithmetic descriptor generator (see oracle/gen.py for
 * the numpy twin).  Writes n rows of 256 fp32, unit L2 norm. */
void vsm_oracle_gen_rows(uint64_t seed, uint64_t set_id, int64_t row0, int64_t n, float* out);

#ifdef __cplusplus
}
#endif
#endif
