/*
 * vsm.h -- C ABI of the B200-native descriptor-matching engine (libvsm.so).
 *
 * Drop-in boundary for the ONE data-parallel hot path of
 * salah-dev-stu/visual-slam-pipeline: brute-force kNN (k=2) over 256-d fp32
 * SuperPoint descriptors + Lowe ratio test (+ optional mutual-NN filter), and
 * top-2 search against a device-resident keyframe descriptor database.
 *
 * The reference has no plugin/FFI layer; it calls OpenCV directly.  Each entry
 * point below cites the reference call it replaces (paths relative to the
 * reference root).  `include/vsm_cv.hpp` is the header-only C++ adaptor that
 * gives these calls the reference's own signatures
 * (cv::Mat in, std::vector<cv::DMatch> out).
 *
 * Conventions
 *   - plain C types only; every function returns a vsm_status (0 = OK) and never
 *     throws.  vsm_last_error(ctx) gives the message of the last failure.
 *   - descriptors are row-major fp32, 256 columns, contiguous rows (row stride =
 *     1024 B), exactly what FeatureExtractor produces
 *     (src/FeatureExtractor.cpp:170-205); the *_strided entry points take a row stride.
 *   - the caller owns all host buffers; the library copies in and never keeps a
 *     host pointer after returning.  Output arrays are caller-allocated.
 *   - results equal cv::BFMatcher(NORM_L2).knnMatch on the same input: identical
 *     indices, identical fp32 distance bits (OpenCV 4.13 baseline build), ties
 *     to the lowest train index.
 *   - a context is single-caller (the reference matches on one thread,
 *     src/main.cpp:1520); calls are synchronous.
 *   - there is NO CPU fallback: without a CUDA device vsm_create fails with
 *     VSM_ERR_NO_DEVICE.
 */
#ifndef VSM_H
#define VSM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VSM_DIM 256

typedef enum {
    VSM_OK = 0,
    VSM_ERR_INVALID = 1,     /* bad argument */
    VSM_ERR_NO_DEVICE = 2,   /* no usable CUDA device / not sm_100 */
    VSM_ERR_CUDA = 3,        /* CUDA runtime/driver failure (message in vsm_last_error) */
    VSM_ERR_CAPACITY = 4,    /* store / scratch capacity exceeded */
    VSM_ERR_NOT_FOUND = 5,   /* unknown keyframe handle */
    VSM_ERR_TIMEOUT = 6      /* fused peer-memory exchange: a peer rank never made the matching call */
} vsm_status;

/* Mirror of cv::DMatch {int queryIdx; int trainIdx; int imgIdx; float distance;}
 * (OpenCV core/types.hpp), 16 bytes; vsm_cv.hpp static_asserts the equality. */
typedef struct {
    int32_t queryIdx;
    int32_t trainIdx;
    int32_t imgIdx;
    float distance;
} vsm_dmatch;

typedef enum {
    VSM_ENGINE_AUTO = 0,     /* tensor-core path (tcgen05) with exact fp32 re-score */
    VSM_ENGINE_TENSOR = 1,   /* same, forced */
    VSM_ENGINE_SIMT = 2,     /* exact fp32 CUDA-core brute force (debug / cross-check) */
    VSM_ENGINE_TENSOR_PAIR = 3   /* tensor-core path on CTA pairs (tcgen05 cta_group::2, clusters of 2) */
} vsm_engine;

typedef struct {
    int32_t device;              /* CUDA device ordinal */
    int32_t engine;              /* vsm_engine */
    int64_t scratch_rows;        /* initial capacity for transient descriptor rows (queries,
                                    pair frames, ragged batches); grows on demand; 0 = 8192 */
    int64_t store_rows;          /* initial capacity of the keyframe store in rows; it grows
                                    on demand; 0 = allocate at the first add */
    int32_t reserved[8];         /* [0]: tiles per slice segment (0 = automatic);
                                    [1]: rescan work-list capacity (0 = default; tests shrink it);
                                    [2]: plain frames vsm_track keeps before recycling (0 = 2);
                                    [3]: open (query, keyframe) pairs vsm_loop_detect_compact can hold (0 = 262144; tests shrink it);
                                    [4]: train sets of up to this many 256-row tiles use append records instead of
                                         top-4 records (0 = never: measured slower; kept as a tested option);
                                    [5]: 1 = pair matching (match_features without a raw list, its mutual reverse
                                         problem) keeps the threshold-driven top-4 records instead of the tile top-2
                                         records it uses by default on train sets of up to 8192 rows (A/B, tests) */
} vsm_opts;

typedef struct vsm_ctx vsm_ctx;

/* Counters of the last call (for parity reports and the bench). */
typedef struct {
    int64_t candidates;          /* (query,row) pairs re-scored in exact fp32 */
    int64_t flagged_slices;      /* (query,slice) pairs that fell back to an exact slice scan */
    int64_t kernel_launches;     /* kernels launched by the last call */
    float   device_ms;           /* device time of the last call (CUDA events, incl. copies) */
    float   tc_ms;               /* device time of the tensor-core kernel of the last call */
    float   select_ms;           /* device time of the select / exact re-score kernel */
    int32_t slice_tiles;         /* 256-row tiles per record slice of the last planned search (a big database
                                    search picks 64, or 16 after a search whose slices overflowed often) */
    int32_t reserved;
} vsm_stats;

void        vsm_default_opts(vsm_opts* opts);
int         vsm_create(const vsm_opts* opts, vsm_ctx** out);
void        vsm_destroy(vsm_ctx* ctx);
const char* vsm_last_error(const vsm_ctx* ctx);       /* ctx may be NULL: create() errors */
const char* vsm_version(void);
int         vsm_get_stats(vsm_ctx* ctx, vsm_stats* out);

/* Pinned host memory for callers that want DMA without a staging copy. */
int  vsm_host_alloc(void** ptr, int64_t bytes);
void vsm_host_free(void* ptr);

/* ---- pair matching ------------------------------------------------------- */

/* Replaces cv::DescriptorMatcher::knnMatch(query, train, knn, 2) at
 * src/Slam.cpp:1149, :567, :764 and src/LoopCloser.cpp:51.
 * idx/dist are [nq][2]; a missing neighbour (nt < 2) is idx = -1, dist = FLT_MAX
 * (the reference drops those lists with its `m.size() >= 2` guard). */
int vsm_knn2(vsm_ctx* ctx, const float* query, int32_t nq, const float* train, int32_t nt,
             int32_t* idx, float* dist);

/* Replaces Slam::match_features(desc1, desc2, raw_out) for float descriptors
 * (src/Slam.cpp:1140-1172; decl include/Slam.h:69-70): kNN k=2, then
 *   raw  = every m[0] whose list has 2 entries                     (:1152-1153)
 *   good = those with m[0].distance < ratio * m[1].distance (fp32)  (:1154)
 * both in query order.  ratio is Config::L2_RATIO_THRESHOLD (0.75f) or
 * FLANN_RATIO_THRESHOLD (0.7f) (include/Config.h:53-55).  mutual != 0 adds the
 * north-star mutual-NN filter: keep m only if query is also train's nearest.
 * good and raw must each hold nq entries; raw/n_raw may be NULL.
 * nq == 0 or nt == 0 -> OK with zero outputs (:1143). */
int vsm_match_pair(vsm_ctx* ctx, const float* query, int32_t nq, const float* train, int32_t nt,
                   float ratio, int32_t mutual,
                   vsm_dmatch* good, int32_t* n_good, vsm_dmatch* raw, int32_t* n_raw);

/* The same two calls for rows that are NOT contiguous -- a cv::Mat ROI or any Mat with step > cols * 4
 * (cv::Mat::step): *_stride = bytes from one row to the next, >= 1024 and a multiple of 4.  The rows are
 * packed by one strided DMA; no host copy. */
int vsm_knn2_strided(vsm_ctx* ctx, const float* query, int32_t nq, int64_t q_stride, const float* train, int32_t nt,
                     int64_t t_stride, int32_t* idx, float* dist);
int vsm_match_pair_strided(vsm_ctx* ctx, const float* query, int32_t nq, int64_t q_stride, const float* train, int32_t nt,
                           int64_t t_stride, float ratio, int32_t mutual,
                           vsm_dmatch* good, int32_t* n_good, vsm_dmatch* raw, int32_t* n_raw);

/* Ragged batch of independent pairs in one launch (BASELINE configs[4]).
 * query/train: concatenated rows; q_off/t_off: n_pairs+1 row offsets.
 * good: concatenated, pair p's survivors start at good[q_off[p]]; n_good[p] entries. */
int vsm_match_batch(vsm_ctx* ctx, int32_t n_pairs,
                    const float* query, const int32_t* q_off,
                    const float* train, const int32_t* t_off,
                    float ratio, int32_t mutual, vsm_dmatch* good, int32_t* n_good);

/* ---- device-resident frame store (Frame::descriptors_, Map::frames_) ----------------------------
 * The store mirrors Map::frames_ (src/Map.cpp:7-10): every stored frame has a handle, a frame id and
 * the reference's is_keyframe flag (include/Frame.h; false until Frame::set_keyframe(true)).
 *   - vsm_store_add* adds a KEYFRAME; vsm_track adds the frame being tracked as a PLAIN frame, which the
 *     caller promotes (vsm_store_promote) when the reference calls set_keyframe(true)
 *     (src/Slam.cpp:590, :660, :828, :852, :919, :1065, :1076).
 *   - "the keyframes" below always means Map::get_keyframes() (src/Map.cpp:40-47): the live frames
 *     with the flag set, in insertion order.  Per-keyframe outputs (counts / status arrays) are
 *     indexed by position in that list; vsm_store_keyframes returns its handles.  While frames are
 *     only added with vsm_store_add and never removed, position == handle.
 *   - plain frames are transient: only last_frame_ and the current frame are ever matched again
 *     (src/Slam.cpp:838, :848), so vsm_track keeps the newest `ring` (default 2) plain frames and
 *     recycles the rows of older ones; a tracking sequence does not grow the store.
 *   - vsm_store_remove frees a frame's rows for reuse; its handle value is reused by a later add.
 *   - the arrays grow in place (CUDA virtual memory management: a reserved address range into which
 *     physical chunks are mapped), so adding a keyframe never copies or moves existing rows.
 *   - VSM_ERR_CAPACITY: the device is out of memory (the store is left as it was). */

/* Mirrors Map::add_frame for a keyframe (src/Map.cpp, include/Frame.h:37,61):
 * uploads the N x 256 descriptor matrix once; fp32 master + bf16 shadow live on
 * the device.  *handle identifies the frame. */
int vsm_store_add(vsm_ctx* ctx, int32_t frame_id, const float* desc, int32_t n, int32_t* handle);
/* vsm_store_add for a descriptor matrix with a row stride (cv::Mat::step), >= 1024 bytes. */
int vsm_store_add_strided(vsm_ctx* ctx, int32_t frame_id, const float* desc, int32_t n, int64_t stride_bytes, int32_t* handle);
/* Frame::set_keyframe(true) for a frame stored by vsm_track: it joins the keyframes at its own
 * position in insertion order (a bridge keyframe promoted late, src/Slam.cpp:851-863, included). */
int vsm_store_promote(vsm_ctx* ctx, int32_t handle);
/* Drops a stored frame (keyframe or plain); its rows and its handle value are reused. */
int vsm_store_remove(vsm_ctx* ctx, int32_t handle);
/* First store row, row count, frame id and keyframe flag of a live handle (any pointer may be NULL):
 * what a caller needs to size the good / raw buffers of vsm_match_to_stored and vsm_track, and to turn
 * a store row returned by vsm_db_top2 into (frame, keypoint index = row - row0). */
int vsm_store_frame_info(const vsm_ctx* ctx, int32_t handle, int64_t* row0, int32_t* n_rows, int32_t* frame_id,
                         int32_t* is_keyframe);
/* Handles of the keyframes in Map::get_keyframes() order; *n = their number (also when cap is smaller). */
int vsm_store_keyframes(const vsm_ctx* ctx, int32_t* handles, int32_t cap, int32_t* n);
/* Same, from a device pointer (bulk loads; no host round trip). */
int vsm_store_add_device(vsm_ctx* ctx, int32_t frame_id, const float* d_desc, int64_t n, int32_t* handle);
/* Adopt an externally owned device fp32 matrix as the whole store (no copy of the
 * fp32 master; builds the bf16 shadow).  seg_off: nseg+1 row offsets or NULL (one segment). */
int vsm_store_adopt_device(vsm_ctx* ctx, const float* d_desc, int64_t n_rows,
                           const int64_t* seg_off, int32_t nseg);
/* Bulk-load the reference's feature cache (FeatureExtractor::save_cache / load_cache,
 * src/FeatureExtractor.cpp:269-381: magic 0x53504346 "SPCF", version 1, per entry frame_idx,
 * keypoints (7 x 4 B each), rows, cols, type, raw descriptor bytes): every entry holding an
 * N x 256 CV_32F matrix becomes one keyframe (frame_id = frame_idx), in file order.
 * n_loaded / n_skipped (entries of another type, e.g. the ORB fallback's CV_8U) may be NULL;
 * first_handle receives the handle of the first keyframe added (-1 if none). */
int vsm_store_load_spcf(vsm_ctx* ctx, const char* path, int32_t* n_loaded, int32_t* n_skipped,
                        int32_t* first_handle);
int vsm_store_clear(vsm_ctx* ctx);
/* n_rows = rows in use up to the high-water mark (removed ranges inside it included: the length a
 * vsm_db_top2_masked mask must have); n_keyframes = live keyframes. */
int vsm_store_info(const vsm_ctx* ctx, int64_t* n_rows, int32_t* n_keyframes);

/* Slam::match_features(ref_kf->descriptors(), cur->descriptors()) with the reference
 * keyframe already resident (src/Slam.cpp:841: query = stored keyframe, train = current). */
int vsm_match_to_stored(vsm_ctx* ctx, int32_t handle, const float* cur, int32_t n_cur,
                        float ratio, int32_t mutual,
                        vsm_dmatch* good, int32_t* n_good, vsm_dmatch* raw, int32_t* n_raw);

/* The tracking step of Slam::process_frame (src/Slam.cpp:838-842: ref = last keyframe or last
 * frame; matches = match_features(ref->descriptors(), frame->descriptors(), &raw)) for a
 * sequence: uploads the current frame ONCE, straight into the store as a PLAIN frame
 * (*cur_handle; see vsm_store_promote), and matches the resident reference `ref_handle` -- a keyframe
 * or the previous plain frame -- against it (query = reference, train = current).
 * ref_handle < 0: only store the frame (first frame of a sequence).
 * good / raw must hold as many entries as the reference frame has rows (vsm_store_frame_info). */
int vsm_track(vsm_ctx* ctx, int32_t ref_handle, int32_t frame_id, const float* cur, int32_t n_cur,
              float ratio, int32_t mutual, vsm_dmatch* good, int32_t* n_good, vsm_dmatch* raw, int32_t* n_raw,
              int32_t* cur_handle);

/* Global top-2 per query over the rows of every KEYFRAME of the store -- the stacked-matrix search of
 * src/Slam.cpp:546-574 and :744-774 (knnMatch(frame, all_descs, 2)).  Rows of plain or removed
 * frames are skipped.  idx: [nq][2] store row (+ row_offset, for sharded stores), dist: [nq][2]. */
int vsm_db_top2(vsm_ctx* ctx, const float* query, int32_t nq, int64_t row_offset,
                int64_t* idx, float* dist);

/* LoopCloser::detect matching block (src/LoopCloser.cpp:43-62): for every stored
 * keyframe s: top-2 WITHIN the keyframe, ratio test, survivors counted.
 * counts: [n_keyframes].  matches (may be NULL): [n_keyframes][nq], survivors of
 * keyframe s first in query order, trainIdx keyframe-local, imgIdx = s. */
int vsm_db_segmented(vsm_ctx* ctx, const float* query, int32_t nq, float ratio,
                     int32_t* counts, vsm_dmatch* matches);

/* The same search restricted to the store rows with mask[row] != 0 -- the map-point searches of
 * src/Slam.cpp:546-574 (only valid map points, :553) and :744-774 (only points observed near the
 * loop keyframe, :748-756), which the reference implements by re-stacking the selected
 * descriptors on every call.  The mask is copied to the device as it is and the selection (numbering in
 * ascending row order, gather) happens there.  n_mask must equal the store's row count.  idx holds ORIGINAL store
 * rows (the reference's mp_ids_vec[trainIdx], :768); order and ties are those of the compacted
 * matrix the reference builds (ascending row).  -1 / FLT_MAX when fewer than k rows are selected. */
int vsm_db_top2_masked(vsm_ctx* ctx, const float* query, int32_t nq, const uint8_t* mask, int64_t n_mask,
                       int64_t* idx, float* dist);

/* ---- resident map-point table (Map::map_points_, include/MapPoint.h) -------------------------------
 * The reference re-stacks the descriptors of the selected map points into a fresh cv::Mat on every
 * search (src/Slam.cpp:552-557, :744-759) -- O(map) host work per call.  Here a point's descriptor, its
 * valid_ flag and its observations_ live on the device from birth; a search selects, compacts (ascending
 * point id = the order of the reference's loop) and gathers on the device, nothing is uploaded but the frame.
 *   vsm_points_add            MapPoint(id, pos, desc) + add_observation(frame_id, .) (src/Slam.cpp:1342-1345,
 *                             :1566-1568); ids are 0, 1, 2 ... like the reference's next_id
 *   vsm_points_add_from_frame the same with desc = row kp_idx[i] of a stored frame (row(i).clone(), :1339, :1563):
 *                             a device-to-device copy; the first observation is that frame's id
 *   vsm_points_observe        MapPoint::add_observation (src/Slam.cpp:463)
 *   vsm_points_set_valid      MapPoint::set_valid (src/Slam.cpp:490, :496, :1119, :1123)
 *   vsm_points_top2           knnMatch(frame, stack of the selected points, 2): selected = valid, and -- when
 *                             near_frame_id >= 0 -- observed in a frame f with |f - near_frame_id| < range
 *                             (Config::LC_NEARBY_FRAME_RANGE, :749-754).  idx = POINT IDS (mp_ids_vec[trainIdx],
 *                             :768), -1 = fewer than k points selected; *n_selected = rows of the stacked matrix
 *                             (the reference's `all_mp_descs.rows >= 50` / `mp_descs.rows >= 20` gates, :561, :760). */
int vsm_points_add(vsm_ctx* ctx, const float* desc, int32_t n, int32_t frame_id, int32_t* first_id);
int vsm_points_add_from_frame(vsm_ctx* ctx, int32_t handle, const int32_t* kp_idx, int32_t n, int32_t* first_id);
int vsm_points_observe(vsm_ctx* ctx, const int32_t* point_ids, int32_t n, int32_t frame_id);
int vsm_points_set_valid(vsm_ctx* ctx, const int32_t* point_ids, int32_t n, int32_t valid);
int vsm_points_info(const vsm_ctx* ctx, int64_t* n_points, int64_t* n_valid, int64_t* n_observations);
int vsm_points_clear(vsm_ctx* ctx);
int vsm_points_top2(vsm_ctx* ctx, const float* query, int32_t nq, int32_t near_frame_id, int32_t range, int64_t* idx,
                    float* dist, int32_t* n_selected);

/* ---- projected-window map-point tracking (Slam::track_local_map, src/Slam.cpp:380-469) ---------
 * For every valid map point: project it with the frame pose (fp64, :417-428), take the keypoints
 * within search_radius pixels (:452-454; the reference walks a 30-px cell grid, the same set),
 * keep the one with the smallest descriptor distance below desc_threshold (:456-460, cv::norm in
 * double), then assign map points to keypoints in map-point order, a later point replacing an
 * earlier one only with a strictly smaller distance (:465-470).
 * Distances are sqrt(sum((double)a-(double)b)^2) in fp64; OpenCV's cv::norm uses a dispatch-dependent
 * summation order, so they agree to ~1e-15 relative, decisions except on such near-ties. */
typedef struct {
    double fx, fy, cx, cy;           /* Config::FX, FY, CX, CY (include/Config.h:14-17) */
    int32_t width, height;           /* Config::IMAGE_WIDTH, IMAGE_HEIGHT (:10-11) */
    int32_t cell_size;               /* Config::TRACK_GRID_CELL_SIZE (:108): fixes the visiting order of ties */
    int32_t reserved;
    double depth_min, depth_max;     /* (double)Config::DEPTH_MIN (:29), Config::TRIANG_MAX_DEPTH (:72) */
    double search_radius;            /* Config::TRACK_SEARCH_RADIUS (:109) */
    double desc_threshold;           /* Config::TRACK_DESC_THRESHOLD (:110) */
} vsm_track_cfg;

/* kp_xy: [nkp][2] keypoint pixel coordinates, desc: [nkp][256] the frame's descriptors;
 * mp_pos: [nmp][3] map-point positions, mp_desc: [nmp][256] their descriptors (NULL: points 0 .. nmp-1 of
 * the resident map-point table, vsm_points_*; if that table is empty, the first nmp rows of the frame
 * store), mp_valid: [nmp] (NULL = the table's own validity flags, or all valid without the table);
 * R_cam (row-major 3x3), t_cam: world -> camera (the reference's R.t(), -R.t()*t, :404-407).
 * indices: [nkp] in/out = frame->map_point_indices(); obs_mp / obs_ki: [nmp] the (map point,
 * keypoint) pairs for MapPoint::add_observation in the order the reference adds them (:468),
 * *tracked = their count (the function's return value, :469).  best_ki / best_dist ([nmp], may be
 * NULL): per map point the chosen keypoint (-1 = none) and its distance. */
int vsm_track_local_map(vsm_ctx* ctx, const vsm_track_cfg* cfg, const float* kp_xy, const float* desc, int32_t nkp,
                        const double* mp_pos, const float* mp_desc, const uint8_t* mp_valid, int32_t nmp,
                        const double* R_cam, const double* t_cam, int32_t* indices, int32_t* obs_mp,
                        int32_t* obs_ki, int32_t* tracked, int32_t* best_ki, double* best_dist);

/* Slam::match_features for n_pairs pairs of STORED keyframes in one launch sequence -- the ragged
 * batch of BASELINE configs[4] with every frame already resident (handles from vsm_store_add /
 * vsm_track), e.g. the keyframe-to-keyframe matches before triangulation (src/Slam.cpp:926) for a
 * window of keyframes.  Nothing is uploaded.  queryIdx indexes keyframe q_handle[p], trainIdx
 * keyframe t_handle[p].  good_off: [n_pairs+1] (out) = prefix sums of the query keyframes' row
 * counts; pair p's survivors are good[good_off[p] .. good_off[p] + n_good[p]); good_cap (entries)
 * must be >= good_off[n_pairs] or the call fails with VSM_ERR_INVALID.  good == NULL with
 * good_cap == 0 is a size query: only good_off is filled. */
int vsm_match_batch_stored(vsm_ctx* ctx, int32_t n_pairs, const int32_t* q_handle, const int32_t* t_handle,
                           float ratio, int32_t mutual, vsm_dmatch* good, int64_t good_cap,
                           int32_t* n_good, int64_t* good_off);

/* LoopCloser::detect's candidate loop WITH its eligibility rules (src/LoopCloser.cpp:43-62):
 * stored keyframes are visited in store order; a keyframe is skipped when
 * cur_frame_id - frame_id < min_gap (:44, Config::LC_MIN_FRAME_GAP = 200) or it is empty (:45);
 * of the remaining ones only every `every`-th is matched (:47-48: checked++; checked % 5 != 0 -> skip).
 * status: [n_keyframes], -1 = skipped by those rules, otherwise the number of ratio-test
 * survivors (the caller applies its >= Config::MIN_MATCHES gate, :62).  matches as in
 * vsm_db_segmented.  Only the eligible keyframes are matched on the device. */
int vsm_loop_detect(vsm_ctx* ctx, int32_t cur_frame_id, int32_t min_gap, int32_t every,
                    const float* query, int32_t nq, float ratio, int32_t* status, vsm_dmatch* matches);

/* The same loop for ONE SHARD of a keyframe list partitioned across GPUs (whole keyframes per
 * rank, SURVEY 8e: no collective, the host concatenates the per-rank status arrays).  The
 * every-`every`-th rule counts over the WHOLE list (src/LoopCloser.cpp:47-48), so the caller
 * passes checked_before = the number of keyframes on earlier shards that pass the gap (:44) and
 * non-empty (:45) tests; *checked_after (may be NULL) = checked_before + this shard's own count.
 * vsm_loop_detect is this call with checked_before = 0. */
int vsm_loop_detect_shard(vsm_ctx* ctx, int32_t cur_frame_id, int32_t min_gap, int32_t every,
                          int32_t checked_before, const float* query, int32_t nq, float ratio,
                          int32_t* status, vsm_dmatch* matches, int32_t* checked_after);

/* LoopCloser::detect's candidate loop in COMPACT form (src/LoopCloser.cpp:43-62, gate at :62 included):
 * the reference only ever uses the good_matches of keyframes with at least Config::MIN_MATCHES (30)
 * survivors, so only those come back.  Same eligibility rules and shard arguments as
 * vsm_loop_detect_shard (checked_before = 0 / checked_after = NULL for a whole list).
 *   status  [n_keyframes]: -1 skipped, else the number of ratio-test survivors
 *   cands   up to cand_cap entries, ascending keyframe position: the keyframes with
 *           count >= max(min_matches, 1); entry k's survivors are matches[offset .. offset + count),
 *           in query order, trainIdx keyframe-local, imgIdx = keyframe position
 *   *n_cands / *n_matches: totals (if one exceeds its capacity the arrays hold the first part: call again
 *           with larger ones)
 * On the device the ratio test is dismissed inside the tensor-core epilogue for almost every (query,
 * keyframe) pair; exact scans, the gate and the packing touch the remaining OPEN pairs only -- no buffer of
 * size keyframes x queries exists anywhere. */
typedef struct {
    int32_t keyframe;            /* position in Map::get_keyframes() order */
    int32_t count;               /* survivors (>= min_matches) */
    int64_t offset;              /* of its list in `matches` */
} vsm_loop_candidate;
int vsm_loop_detect_compact(vsm_ctx* ctx, int32_t cur_frame_id, int32_t min_gap, int32_t every, int32_t checked_before,
                            const float* query, int32_t nq, float ratio, int32_t min_matches, int32_t* status,
                            vsm_loop_candidate* cands, int32_t cand_cap, int32_t* n_cands, vsm_dmatch* matches,
                            int64_t match_cap, int64_t* n_matches, int32_t* checked_after);

/* Frame ids of the stored keyframes, in store order (n must equal the keyframe count).  For a
 * store built by vsm_store_adopt_device, whose keyframes otherwise get ids 0..n-1
 * (Frame::id(), include/Frame.h, read by the gap rule src/LoopCloser.cpp:44). */
int vsm_store_set_frame_ids(vsm_ctx* ctx, const int32_t* frame_ids, int32_t n);

/* ---- device-pointer variants (resident data; multi-GPU plumbing) ---------- */

/* vsm_db_top2 with query and outputs already on this context's device.
 * d_idx: int64 [nq][2], d_dist: float [nq][2].  Asynchronous on the context stream
 * unless sync != 0. */
int vsm_db_top2_device(vsm_ctx* ctx, const float* d_query, int32_t nq, int64_t row_offset,
                       int64_t* d_idx, float* d_dist, int32_t sync);

/* Same search, result as two 64-bit keys per query: ~((distance bits << 32) | GLOBAL index), 0 = empty
 * (global index = store row + row_offset, must fit 32 bits).  One buffer to all-gather instead of two. */
int vsm_db_top2_keys_device(vsm_ctx* ctx, const float* d_query, int32_t nq, int64_t row_offset,
                            uint64_t* d_keys, int32_t sync);
/* Merge gathered keys [nshard][nq][2] -> idx [nq][2] (int64, -1 = none) + dist [nq][2]. */
int vsm_merge_keys_device(vsm_ctx* ctx, const uint64_t* d_keys_in, int32_t nshard, int32_t nq,
                          int64_t* d_idx_out, float* d_dist_out, int32_t sync);

/* ---- fused exchange over peer memory (one process per GPU, NVLink / NVSwitch) -------------------
 * Instead of an NCCL all-gather between the search and the merge, every rank STORES its result keys
 * straight into all peers' gather buffers (P2P over NVLink), raises a per-peer flag, waits for the
 * peers' flags and merges -- one kernel after the local search, no collective launch.
 *   vsm_xchg_create   allocates this rank's buffer (2 parities x world x nq_cap keys + flags) and
 *                     returns its 64-byte CUDA IPC handle; exchange the handles between the ranks
 *                     (any transport) and pass all of them, in rank order, to vsm_xchg_connect.
 *                     All ranks must have connected before the first search (barrier).
 *   vsm_db_top2_xchg_device   = vsm_db_top2_keys_device + publish + wait + merge.  Collective: every
 *                     rank of the group must call it the same number of times with the same nq. */
int vsm_xchg_create(vsm_ctx* ctx, int32_t rank, int32_t world, int32_t nq_cap, uint8_t handle_out[64]);
int vsm_xchg_connect(vsm_ctx* ctx, const uint8_t* handles /* [world][64] */);
int vsm_db_top2_xchg_device(vsm_ctx* ctx, const float* d_query, int32_t nq, int64_t row_offset,
                            int64_t* d_idx_out, float* d_dist_out, int32_t sync);
/* The same collective search with HOST buffers in and out (the reference-facing form of one rank of a
 * partitioned loop-closure search): queries are read from host memory, the merged global top-2 is
 * written straight into pinned host memory by the exchange kernel and copied to idx / dist.
 * Synchronous.  Every rank must make the same sequence of exchange calls with the same nq; a peer
 * that is more than VSM_XCHG_TIMEOUT_S (default 10) seconds late makes the call fail with
 * VSM_ERR_TIMEOUT on the ranks that waited for it -- the context stays usable. */
int vsm_db_top2_xchg(vsm_ctx* ctx, const float* query, int32_t nq, int64_t row_offset, int64_t* idx, float* dist);

/* Merge per-shard top-2 lists (e.g. after an NCCL all-gather) by (distance, index).
 * d_idx_in/d_dist_in: [nshard][nq][2] with GLOBAL indices (-1 = empty). */
int vsm_merge_top2_device(vsm_ctx* ctx, const int64_t* d_idx_in, const float* d_dist_in,
                          int32_t nshard, int32_t nq, int64_t* d_idx_out, float* d_dist_out,
                          int32_t sync);

/* ---- several GPUs behind one caller thread ---------------------------------------------------------
 * The reference matches on ONE thread of ONE process (src/main.cpp:1520; LoopCloser::detect,
 * src/LoopCloser.cpp:16-18, runs inside it), so the drop-in for a box of GPUs is a group: one context
 * per device and one worker thread per context inside the library.  A group call fans out on the
 * workers, joins, and merges on the first device -- every member stores its exact local top-2 (keys with
 * the STACKED row index of the whole database) straight into the first device's gather buffer over
 * NVLink (peer access; pinned host memory if there is no peer path) and the first device's stream
 * waits for the members' events before it merges: no CUDA IPC, no process group, no spinning kernel.
 * Keyframes are dealt to the members whole, each to the member that holds the fewest rows.
 * A device may be listed more than once (several contexts on one GPU; how the 1-GPU tests run it).
 * Group keyframe handles are 0, 1, 2 ... in insertion order and are never reused. */
typedef struct vsm_group vsm_group;
int         vsm_group_create(const int32_t* devices, int32_t n, const vsm_opts* opts, vsm_group** out);
void        vsm_group_destroy(vsm_group* g);
const char* vsm_group_last_error(const vsm_group* g);       /* g may be NULL: create() errors */
int         vsm_group_size(const vsm_group* g);
/* Member context (pair matching, statistics): calls on it follow the single-caller rule. */
vsm_ctx*    vsm_group_ctx(vsm_group* g, int32_t member);
/* Map::add_frame for a keyframe (src/Map.cpp:7-10): uploaded to ONE member. */
int vsm_group_store_add(vsm_group* g, int32_t frame_id, const float* desc, int32_t n, int32_t* handle);
int vsm_group_store_remove(vsm_group* g, int32_t handle);
int vsm_group_store_clear(vsm_group* g);
/* rows_per_member: [vsm_group_size] live keyframe rows on each member (may be NULL). */
int vsm_group_store_info(const vsm_group* g, int64_t* n_rows, int32_t* n_keyframes, int64_t* rows_per_member);
/* Bulk load: member `member` adopts a device matrix that lives on ITS device (vsm_store_adopt_device);
 * members must adopt in ascending order on an empty group, which makes stacked row = position in the
 * concatenation of the members' matrices. */
int vsm_group_adopt_device(vsm_group* g, int32_t member, const float* d_desc, int64_t n_rows, const int64_t* seg_off,
                           int32_t nseg);
/* knnMatch(frame, all_descs, 2) over every keyframe row of every member (src/Slam.cpp:546-574, :744-774):
 * host queries in, host result out.  idx: [nq][2] stacked row (-1 = none), dist: [nq][2];
 * kf_handle / kf_row ([nq][2], may be NULL): the keyframe holding the row and the row inside it. */
int vsm_group_db_top2(vsm_group* g, const float* query, int32_t nq, int64_t* idx, float* dist, int32_t* kf_handle,
                      int32_t* kf_row);
/* vsm_loop_detect over the group's keyframe list: status / matches are indexed by position among the
 * live keyframes in insertion order (= handle while nothing has been removed). */
int vsm_group_loop_detect(vsm_group* g, int32_t cur_frame_id, int32_t min_gap, int32_t every, const float* query,
                          int32_t nq, float ratio, int32_t* status, vsm_dmatch* matches);
/* vsm_loop_detect_compact over the group's list: each member gates and packs on its own device, the
 * host orders the surviving keyframes by list position (cands[k].keyframe). */
int vsm_group_loop_detect_compact(vsm_group* g, int32_t cur_frame_id, int32_t min_gap, int32_t every, const float* query,
                                  int32_t nq, float ratio, int32_t min_matches, int32_t* status, vsm_loop_candidate* cands,
                                  int32_t cand_cap, int32_t* n_cands, vsm_dmatch* matches, int64_t match_cap,
                                  int64_t* n_matches);

/* vsm_match_batch with the pairs dealt to the members in contiguous blocks of near-equal input size
 * (replicas: a pair is never split); every member uploads its block over its own PCIe link. */
int vsm_group_match_batch(vsm_group* g, int32_t n_pairs, const float* query, const int32_t* q_off, const float* train,
                          const int32_t* t_off, float ratio, int32_t mutual, vsm_dmatch* good, int32_t* n_good);

/* Raw CUDA stream of the context (cudaStream_t as void*), for event timing. */
void* vsm_stream(vsm_ctx* ctx);
/* Run on a caller-owned stream (e.g. the stream an NCCL collective is enqueued on). */
int   vsm_set_stream(vsm_ctx* ctx, void* cuda_stream);
int   vsm_sync(vsm_ctx* ctx);
/* CUDA events and counters behind vsm_stats (device_ms, tc_ms, select_ms, candidates, flagged_slices);
 * on by default.  Off: those fields read 0 and a call is a few microseconds shorter. */
int   vsm_set_profiling(vsm_ctx* ctx, int32_t on);
/* Device time of the tensor-core kernel in each of the last calls (oldest first, at most the last
 * 64 calls made with profiling on; 0 for a call that did not run it).  Waits for the stream.  Lets
 * a caller that enqueues a loop of asynchronous searches read every launch's duration afterwards
 * without synchronising inside the loop.  ms: [n], *n_out = entries written. */
int   vsm_tc_history(vsm_ctx* ctx, float* ms, int32_t n, int32_t* n_out);

/* Synthetic SuperPoint-shaped descriptors written straight into device memory (benchmarks in C++):
 * rows row0 .. row0+n of stream `seed`, 256 standard normals scaled to unit length
 * (the shape of src/FeatureExtractor.cpp:170-205's output).  d_dst must live on the context's device. */
int vsm_synth_rows_device(vsm_ctx* ctx, float* d_dst, int64_t row0, int64_t n, uint64_t seed);

/* Debug / bring-up: the raw tensor-core accumulators (bf16 dot products q.t) of the
 * first 128 queries x first 256 train rows, written to out[128*256] (host). */
/* Debug: raw copy of the 128 KB debug buffer (tile scores, or -- with VSM_DEBUG_TIMELINE set --
 * clock64 stamps of unit 0: [0..4095] accumulator ready, [4096..] loads issued, [8192..] MMA
 * issued, [12288..] epilogue done, per tile). */
int vsm_debug_fetch_dump(vsm_ctx* ctx, void* out, int64_t bytes);
int vsm_debug_tile_scores(vsm_ctx* ctx, const float* query, int32_t nq, const float* train,
                          int32_t nt, float* out);

#ifdef __cplusplus
}
#endif
#endif /* VSM_H */
