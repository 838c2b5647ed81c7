/*
 * vsm_cv.hpp -- header-only C++ adaptor over the C ABI of libvsm.so (include/vsm.h) that
 * gives the B200 matcher the reference's own call-site signatures:
 *
 *   Slam::match_features(desc1, desc2, raw_out)          src/Slam.cpp:1140-1172, include/Slam.h:69-70
 *   cv::DescriptorMatcher::knnMatch(query, train, knn, 2) src/Slam.cpp:1149, :567, :764; src/LoopCloser.cpp:51
 *   LoopCloser::detect's per-keyframe block               src/LoopCloser.cpp:43-62
 *
 * cv::Mat (N x 256, CV_32F) in, std::vector<cv::DMatch> out.  With OpenCV's headers on the
 * include path the adaptor uses cv::Mat / cv::DMatch directly; without them (this build
 * image has none) it falls back to two minimal stand-ins with the same member names, so
 * the same call-site code compiles either way.
 *
 * Errors surface as std::runtime_error, like the cv::Exception the reference never catches.
 */
#ifndef VSM_CV_HPP
#define VSM_CV_HPP

#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "vsm.h"

#if defined(__has_include)
#if __has_include(<opencv2/core.hpp>) && !defined(VSM_CV_NO_OPENCV)
#include <opencv2/core.hpp>
#define VSM_CV_HAVE_OPENCV 1
#endif
#endif

namespace vsm_cv {

#ifdef VSM_CV_HAVE_OPENCV
using Mat = cv::Mat;
using DMatch = cv::DMatch;
inline bool is_f32_256(const Mat& m) { return m.type() == CV_32F && m.cols == VSM_DIM; }
#else
/* cv::DMatch (opencv2/core/types.hpp): same members, same order, 16 bytes. */
struct DMatch {
    int queryIdx = -1, trainIdx = -1, imgIdx = -1;
    float distance = 3.402823466e+38f;
    DMatch() {}
    DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
    DMatch(int q, int t, int i, float d) : queryIdx(q), trainIdx(t), imgIdx(i), distance(d) {}
    bool operator<(const DMatch& m) const { return distance < m.distance; }
};
/* The slice of cv::Mat the matcher call sites use: a borrowed row-major fp32 matrix. */
struct Mat {
    int rows = 0, cols = 0;
    const float* data = nullptr;
    size_t step = 0;                                   /* bytes per row */
    Mat() {}
    Mat(int r, int c, const float* p, size_t step_bytes = 0)
        : rows(r), cols(c), data(p), step(step_bytes ? step_bytes : (size_t)c * sizeof(float)) {}
    bool empty() const { return rows == 0 || cols == 0 || !data; }
    bool isContinuous() const { return step == (size_t)cols * sizeof(float); }
    template <class T> const T* ptr(int r = 0) const {
        return reinterpret_cast<const T*>(reinterpret_cast<const char*>(data) + (size_t)r * step);
    }
};
inline bool is_f32_256(const Mat& m) { return m.cols == VSM_DIM; }
#endif

static_assert(sizeof(DMatch) == sizeof(vsm_dmatch), "cv::DMatch must be the 16-byte struct vsm_dmatch mirrors");

class DescriptorMatcher {
public:
    /* One GPU (like the reference's matcher_l2_ member, include/Slam.h:197). */
    explicit DescriptorMatcher(int device = 0, int engine = VSM_ENGINE_AUTO) {
        vsm_opts o;
        vsm_default_opts(&o);
        o.device = device;
        o.engine = engine;
        if (vsm_create(&o, &ctx_) != VSM_OK) throw std::runtime_error(std::string("vsm_create: ") + vsm_last_error(nullptr));
    }
    /* Several GPUs behind this one object: the keyframe database (add_keyframe / search_store /
     * detect_loop) is dealt to the devices by whole keyframes and searched on all of them per call
     * (vsm_group_*); pair matching and tracking run on the first device.  The caller stays a single
     * thread, like the reference's slam_thread (src/main.cpp:1520). */
    explicit DescriptorMatcher(const std::vector<int>& devices, int engine = VSM_ENGINE_AUTO) {
        vsm_opts o;
        vsm_default_opts(&o);
        o.engine = engine;
        std::vector<int32_t> d(devices.begin(), devices.end());
        if (vsm_group_create(d.data(), (int32_t)d.size(), &o, &group_) != VSM_OK)
            throw std::runtime_error(std::string("vsm_group_create: ") + vsm_group_last_error(nullptr));
        ctx_ = vsm_group_ctx(group_, 0);
    }
    ~DescriptorMatcher() {
        if (group_) vsm_group_destroy(group_);
        else vsm_destroy(ctx_);
    }
    DescriptorMatcher(const DescriptorMatcher&) = delete;
    DescriptorMatcher& operator=(const DescriptorMatcher&) = delete;

    /* Slam::match_features for float descriptors (src/Slam.cpp:1140-1172): kNN k=2, raw = every
     * m[0] of a two-entry list, result = those passing m[0].distance < ratio * m[1].distance.
     * The reference's ratio is Config::L2_RATIO_THRESHOLD = 0.75f (include/Config.h:53).
     * A Mat with a row stride (ROI, non-continuous) is passed as it is: the library packs the rows. */
    std::vector<DMatch> match_features(const Mat& desc1, const Mat& desc2, std::vector<DMatch>* raw_out = nullptr,
                                       float ratio = 0.75f, bool mutual = false) {
        std::vector<DMatch> good;
        if (raw_out) raw_out->clear();
        if (desc1.empty() || desc2.empty()) return good;                  /* src/Slam.cpp:1143 */
        need_f32_256(desc1);
        need_f32_256(desc2);
        good.resize(desc1.rows);
        if (raw_out) raw_out->resize(desc1.rows);
        int32_t ng = 0, nr = 0;
        check(vsm_match_pair_strided(ctx_, desc1.template ptr<float>(0), desc1.rows, stride_of(desc1),
                                     desc2.template ptr<float>(0), desc2.rows, stride_of(desc2), ratio, mutual ? 1 : 0,
                                     reinterpret_cast<vsm_dmatch*>(good.data()), &ng,
                                     raw_out ? reinterpret_cast<vsm_dmatch*>(raw_out->data()) : nullptr, raw_out ? &nr : nullptr));
        good.resize(ng);
        if (raw_out) raw_out->resize(nr);
        return good;
    }

    /* cv::DescriptorMatcher::knnMatch(query, train, knn, 2): lists may hold fewer than two
     * entries when train has fewer than two rows (the reference guards with m.size() >= 2). */
    void knnMatch(const Mat& query, const Mat& train, std::vector<std::vector<DMatch>>& knn, int k = 2) {
        if (k != 2) throw std::runtime_error("vsm_cv::knnMatch: only k = 2 (the reference's call sites)");
        knn.assign(query.rows, std::vector<DMatch>());
        if (query.empty()) return;
        need_f32_256(query);
        if (!train.empty()) need_f32_256(train);
        std::vector<int32_t> idx((size_t)query.rows * 2);
        std::vector<float> dist((size_t)query.rows * 2);
        check(vsm_knn2_strided(ctx_, query.template ptr<float>(0), query.rows, stride_of(query),
                               train.empty() ? nullptr : train.template ptr<float>(0), train.empty() ? 0 : train.rows,
                               train.empty() ? VSM_DIM * (int64_t)sizeof(float) : stride_of(train), idx.data(), dist.data()));
        for (int i = 0; i < query.rows; i++)
            for (int p = 0; p < 2; p++)
                if (idx[2 * i + p] >= 0) knn[i].push_back(DMatch(i, idx[2 * i + p], 0, dist[2 * i + p]));
    }

    /* Keyframe descriptors kept on the device (Frame::descriptors_, include/Frame.h:61).  With several
     * devices the keyframe goes to the one holding the fewest rows; the handle is then a group handle. */
    int add_keyframe(int frame_id, const Mat& desc) {
        std::vector<float> b;
        int32_t h = -1;
        if (group_) {
            gcheck(vsm_group_store_add(group_, frame_id, desc.empty() ? nullptr : rows_of(desc, b), desc.empty() ? 0 : desc.rows, &h));
            return h;
        }
        if (!desc.empty()) need_f32_256(desc);
        check(vsm_store_add_strided(ctx_, frame_id, desc.empty() ? nullptr : desc.template ptr<float>(0), desc.empty() ? 0 : desc.rows,
                                    desc.empty() ? VSM_DIM * (int64_t)sizeof(float) : stride_of(desc), &h));
        return h;
    }
    void clear_keyframes() {
        if (group_) gcheck(vsm_group_store_clear(group_));
        else check(vsm_store_clear(ctx_));
    }
    /* Frame::set_keyframe(true) (src/Slam.cpp:1065, :1076, :852) / dropping a frame from the device store. */
    void promote(int handle) { single("promote"); check(vsm_store_promote(ctx_, handle)); }
    void remove_frame(int handle) {
        if (group_) gcheck(vsm_group_store_remove(group_, handle));
        else check(vsm_store_remove(ctx_, handle));
    }
    int frame_rows(int handle) {
        int32_t n = 0;
        if (vsm_store_frame_info(ctx_, handle, nullptr, &n, nullptr, nullptr) != VSM_OK)
            throw std::runtime_error("vsm_cv: unknown frame handle");
        return n;
    }
    int keyframe_count() {
        int64_t rows = 0;
        int32_t nkf = 0;
        if (group_) gcheck(vsm_group_store_info(group_, &rows, &nkf, nullptr));
        else check(vsm_store_info(ctx_, &rows, &nkf));
        return nkf;
    }

    /* match_features(ref_kf->descriptors(), cur->descriptors(), raw) with ref_kf resident (src/Slam.cpp:841). */
    std::vector<DMatch> match_features(int keyframe_handle, const Mat& cur,
                                       std::vector<DMatch>* raw_out = nullptr, float ratio = 0.75f, bool mutual = false) {
        single("match_features(handle, ...)");
        const int keyframe_rows = frame_rows(keyframe_handle);        /* the library's own count sizes the buffers */
        std::vector<DMatch> good(keyframe_rows > 0 ? keyframe_rows : 1);
        if (raw_out) raw_out->assign(good.size(), DMatch());
        std::vector<float> b;
        int32_t ng = 0, nr = 0;
        check(vsm_match_to_stored(ctx_, keyframe_handle, cur.empty() ? nullptr : rows_of(cur, b), cur.empty() ? 0 : cur.rows,
                                  ratio, mutual ? 1 : 0, reinterpret_cast<vsm_dmatch*>(good.data()), &ng,
                                  raw_out ? reinterpret_cast<vsm_dmatch*>(raw_out->data()) : nullptr, raw_out ? &nr : nullptr));
        good.resize(ng);
        if (raw_out) raw_out->resize(nr);
        return good;
    }

    /* match_features for many pairs of stored keyframes in one launch sequence (nothing is uploaded):
     * result[p] = match_features(descriptors of q_handles[p], descriptors of t_handles[p]). */
    std::vector<std::vector<DMatch>> match_features_batch(const std::vector<int>& q_handles, const std::vector<int>& t_handles,
                                                          float ratio = 0.75f, bool mutual = false) {
        single("match_features_batch");
        if (q_handles.size() != t_handles.size()) throw std::runtime_error("vsm_cv: handle lists differ in length");
        const int n = (int)q_handles.size();
        std::vector<std::vector<DMatch>> out(n);
        if (n == 0) return out;
        std::vector<int32_t> qh(q_handles.begin(), q_handles.end()), th(t_handles.begin(), t_handles.end()), ng(n, 0);
        std::vector<int64_t> off(n + 1, 0);
        check(vsm_match_batch_stored(ctx_, n, qh.data(), th.data(), ratio, mutual ? 1 : 0, nullptr, 0, ng.data(), off.data()));
        std::vector<DMatch> flat((size_t)(off[n] > 0 ? off[n] : 1));
        check(vsm_match_batch_stored(ctx_, n, qh.data(), th.data(), ratio, mutual ? 1 : 0,
                                     reinterpret_cast<vsm_dmatch*>(flat.data()), (int64_t)flat.size(), ng.data(), off.data()));
        for (int p = 0; p < n; p++) out[p].assign(flat.begin() + off[p], flat.begin() + off[p] + ng[p]);
        return out;
    }

    /* The tracking match of Slam::process_frame (src/Slam.cpp:838-842) for a sequence: the current
     * frame goes to the device once, as a plain frame *cur_handle (promote() it when the reference
     * calls frame->set_keyframe(true)); ref_handle = last_keyframe_ or last_frame_, < 0 = first frame. */
    std::vector<DMatch> track(int ref_handle, int frame_id, const Mat& cur, int* cur_handle,
                              std::vector<DMatch>* raw_out = nullptr, float ratio = 0.75f, bool mutual = false) {
        single("track");
        const int ref_rows = ref_handle >= 0 ? frame_rows(ref_handle) : 0;
        std::vector<DMatch> good(ref_rows > 0 ? ref_rows : 1);
        if (raw_out) raw_out->assign(good.size(), DMatch());
        std::vector<float> b;
        int32_t ng = 0, nr = 0, h = -1;
        check(vsm_track(ctx_, ref_handle, frame_id, cur.empty() ? nullptr : rows_of(cur, b), cur.empty() ? 0 : cur.rows,
                        ratio, mutual ? 1 : 0, reinterpret_cast<vsm_dmatch*>(good.data()), &ng,
                        raw_out ? reinterpret_cast<vsm_dmatch*>(raw_out->data()) : nullptr, raw_out ? &nr : nullptr, &h));
        if (cur_handle) *cur_handle = h;
        good.resize(ng);
        if (raw_out) raw_out->resize(nr);
        return good;
    }

    /* LoopCloser::detect, lines 43-62: per stored keyframe the ratio-test survivors of a kNN
     * inside that keyframe.  good_matches[s] is what the reference builds at :54-60; the caller
     * keeps its own eligibility rules (:44-48) and the >= 30 gate (:62). */
    void detect_candidates(const Mat& cur, float ratio, std::vector<std::vector<DMatch>>& good_matches) {
        single("detect_candidates");
        const int nkf = keyframe_count();
        good_matches.assign(nkf, std::vector<DMatch>());
        if (cur.empty() || nkf == 0) return;
        std::vector<float> b;
        std::vector<int32_t> counts(nkf);
        std::vector<DMatch> flat((size_t)nkf * cur.rows);
        check(vsm_db_segmented(ctx_, rows_of(cur, b), cur.rows, ratio, counts.data(), reinterpret_cast<vsm_dmatch*>(flat.data())));
        for (int s = 0; s < nkf; s++)
            good_matches[s].assign(flat.begin() + (size_t)s * cur.rows, flat.begin() + (size_t)s * cur.rows + counts[s]);
    }

    /* LoopCloser::detect's whole candidate loop, lines 43-62 INCLUDING the eligibility rules (:44-48:
     * gap >= min_gap = Config::LC_MIN_FRAME_GAP, non-empty, every `every`-th = 5) and the gate at :62
     * (good_matches.size() >= min_matches = Config::MIN_MATCHES).  status[s] = -1 for a skipped keyframe,
     * otherwise its number of survivors; candidates = (keyframe position, good_matches) of the keyframes
     * that pass the gate, ascending -- exactly the lists the reference hands to findEssentialMat (:70).
     * Gate and packing run on the device(s); nothing of size keyframes x queries is allocated. */
    struct LoopCandidate {
        int keyframe;
        std::vector<DMatch> good_matches;
    };
    void detect_loop(int cur_frame_id, const Mat& cur, float ratio, int min_gap, int every, int min_matches,
                     std::vector<int>& status, std::vector<LoopCandidate>& candidates) {
        const int nkf = keyframe_count();
        status.assign(nkf, -1);
        candidates.clear();
        if (nkf == 0) return;
        std::vector<float> b;
        const float* q = cur.empty() ? nullptr : rows_of(cur, b);
        const int nq = cur.empty() ? 0 : cur.rows;
        std::vector<int32_t> st(nkf, -1);
        std::vector<vsm_loop_candidate> cands(16);
        std::vector<DMatch> flat((size_t)16 * (nq > 0 ? nq : 1));
        int32_t nc = 0;
        int64_t nm = 0;
        for (;;) {
            if (group_)
                gcheck(vsm_group_loop_detect_compact(group_, cur_frame_id, min_gap, every, q, nq, ratio, min_matches, st.data(),
                                                     cands.data(), (int32_t)cands.size(), &nc,
                                                     reinterpret_cast<vsm_dmatch*>(flat.data()), (int64_t)flat.size(), &nm));
            else
                check(vsm_loop_detect_compact(ctx_, cur_frame_id, min_gap, every, 0, q, nq, ratio, min_matches, st.data(),
                                              cands.data(), (int32_t)cands.size(), &nc,
                                              reinterpret_cast<vsm_dmatch*>(flat.data()), (int64_t)flat.size(), &nm, nullptr));
            if (nc <= (int32_t)cands.size() && nm <= (int64_t)flat.size()) break;
            if (nc > (int32_t)cands.size()) cands.resize(nc);
            if (nm > (int64_t)flat.size()) flat.resize((size_t)nm);
        }
        for (int s = 0; s < nkf; s++) status[s] = st[s];
        for (int k = 0; k < nc; k++) {
            LoopCandidate c;
            c.keyframe = cands[k].keyframe;
            c.good_matches.assign(flat.begin() + cands[k].offset, flat.begin() + cands[k].offset + cands[k].count);
            candidates.push_back(std::move(c));
        }
    }

    /* The record-based form of the same loop: every eligible keyframe's list comes back (no gate). */
    void detect_loop_candidates(int cur_frame_id, const Mat& cur, float ratio, int min_gap, int every,
                                std::vector<int>& status, std::vector<std::vector<DMatch>>& good_matches) {
        const int nkf = keyframe_count();
        status.assign(nkf, -1);
        good_matches.assign(nkf, std::vector<DMatch>());
        if (nkf == 0) return;
        std::vector<float> b;
        std::vector<int32_t> st(nkf, -1);
        std::vector<DMatch> flat(cur.empty() ? 1 : (size_t)nkf * cur.rows);
        const float* q = cur.empty() ? nullptr : rows_of(cur, b);
        if (group_)
            gcheck(vsm_group_loop_detect(group_, cur_frame_id, min_gap, every, q, cur.empty() ? 0 : cur.rows, ratio, st.data(),
                                         reinterpret_cast<vsm_dmatch*>(flat.data())));
        else
            check(vsm_loop_detect(ctx_, cur_frame_id, min_gap, every, q, cur.empty() ? 0 : cur.rows, ratio, st.data(),
                                  reinterpret_cast<vsm_dmatch*>(flat.data())));
        for (int s = 0; s < nkf; s++) {
            status[s] = st[s];
            if (st[s] > 0)
                good_matches[s].assign(flat.begin() + (size_t)s * cur.rows, flat.begin() + (size_t)s * cur.rows + st[s]);
        }
    }

    /* knnMatch(frame_desc, stack of the stored rows with valid[row] != 0, knn, 2): the map-point searches
     * that re-stack a subset per call (valid points src/Slam.cpp:552-557; points seen near the loop
     * keyframe :748-759).  trainIdx is the ORIGINAL store row (the reference's mp_ids_vec[trainIdx], :768). */
    void search_store(const Mat& frame_desc, const std::vector<unsigned char>& valid,
                      std::vector<std::vector<DMatch>>& knn) {
        single("search_store(valid)");
        knn.assign(frame_desc.rows, std::vector<DMatch>());
        if (frame_desc.empty()) return;
        std::vector<float> b;
        std::vector<int64_t> idx((size_t)frame_desc.rows * 2);
        std::vector<float> dist((size_t)frame_desc.rows * 2);
        check(vsm_db_top2_masked(ctx_, rows_of(frame_desc, b), frame_desc.rows, valid.data(), (int64_t)valid.size(),
                                 idx.data(), dist.data()));
        for (int i = 0; i < frame_desc.rows; i++)
            for (int p = 0; p < 2; p++)
                if (idx[2 * i + p] >= 0) knn[i].push_back(DMatch(i, (int)idx[2 * i + p], 0, dist[2 * i + p]));
    }

    /* knnMatch(frame_desc, all_descs, knn, 2) over every keyframe row (src/Slam.cpp:567, :764).  One
     * device: trainIdx = store row; several devices: trainIdx = row of the stacked database. */
    void search_store(const Mat& frame_desc, std::vector<std::vector<DMatch>>& knn) {
        knn.assign(frame_desc.rows, std::vector<DMatch>());
        if (frame_desc.empty()) return;
        std::vector<float> b;
        std::vector<int64_t> idx((size_t)frame_desc.rows * 2);
        std::vector<float> dist((size_t)frame_desc.rows * 2);
        if (group_)
            gcheck(vsm_group_db_top2(group_, rows_of(frame_desc, b), frame_desc.rows, idx.data(), dist.data(), nullptr, nullptr));
        else
            check(vsm_db_top2(ctx_, rows_of(frame_desc, b), frame_desc.rows, 0, idx.data(), dist.data()));
        for (int i = 0; i < frame_desc.rows; i++)
            for (int p = 0; p < 2; p++)
                if (idx[2 * i + p] >= 0) knn[i].push_back(DMatch(i, (int)idx[2 * i + p], 0, dist[2 * i + p]));
    }

    /* ---- Map::map_points_ on the device (first device of a multi-device matcher) --------------------
     * add_map_points: MapPoint births whose descriptor is frame->descriptors().row(kp).clone() of a stored
     * frame (src/Slam.cpp:1337-1347, :1561-1570); returns the first new id (ids follow the reference's next_id).
     * observe_map_points / set_map_points_valid: MapPoint::add_observation (:463) / set_valid (:496, :1119-1123).
     * search_map_points: the searches of src/Slam.cpp:546-574 (near_frame_id < 0: every valid point) and
     * :744-774 (valid points observed within `range` frames of near_frame_id); trainIdx = POINT ID
     * (mp_ids_vec[m[0].trainIdx], :768); returns the rows of the stacked matrix (the >= 50 / >= 20 gates). */
    int add_map_points(int frame_handle, const std::vector<int>& keypoint_idx) {
        std::vector<int32_t> k(keypoint_idx.begin(), keypoint_idx.end());
        int32_t first = -1;
        check(vsm_points_add_from_frame(ctx_, frame_handle, k.data(), (int32_t)k.size(), &first));
        return first;
    }
    void observe_map_points(const std::vector<int>& point_ids, int frame_id) {
        std::vector<int32_t> p(point_ids.begin(), point_ids.end());
        check(vsm_points_observe(ctx_, p.data(), (int32_t)p.size(), frame_id));
    }
    void set_map_points_valid(const std::vector<int>& point_ids, bool valid) {
        std::vector<int32_t> p(point_ids.begin(), point_ids.end());
        check(vsm_points_set_valid(ctx_, p.data(), (int32_t)p.size(), valid ? 1 : 0));
    }
    int search_map_points(const Mat& frame_desc, int near_frame_id, int range, std::vector<std::vector<DMatch>>& knn) {
        knn.assign(frame_desc.rows, std::vector<DMatch>());
        std::vector<float> b;
        std::vector<int64_t> idx((size_t)(frame_desc.rows > 0 ? frame_desc.rows : 1) * 2);
        std::vector<float> dist(idx.size());
        int32_t rows_stacked = 0;
        check(vsm_points_top2(ctx_, frame_desc.empty() ? nullptr : rows_of(frame_desc, b), frame_desc.empty() ? 0 : frame_desc.rows,
                              near_frame_id, range, idx.data(), dist.data(), &rows_stacked));
        for (int i = 0; i < frame_desc.rows; i++)
            for (int p = 0; p < 2; p++)
                if (idx[2 * i + p] >= 0) knn[i].push_back(DMatch(i, (int)idx[2 * i + p], 0, dist[2 * i + p]));
        return rows_stacked;
    }

    vsm_ctx* handle() { return ctx_; }
    vsm_group* group() { return group_; }

private:
    vsm_ctx* ctx_ = nullptr;
    vsm_group* group_ = nullptr;

    void check(int st) {
        if (st != VSM_OK) throw std::runtime_error(std::string("libvsm: ") + vsm_last_error(ctx_));
    }
    void gcheck(int st) {
        if (st != VSM_OK) throw std::runtime_error(std::string("libvsm group: ") + vsm_group_last_error(group_));
    }
    void single(const char* what) {
        if (group_) throw std::runtime_error(std::string("vsm_cv: ") + what + " addresses one device's store; not available on a multi-device matcher");
    }
    static void need_f32_256(const Mat& m) {
        if (!is_f32_256(m)) throw std::runtime_error("vsm_cv: descriptors must be N x 256 CV_32F");
    }
    /* bytes from one row to the next (cv::Mat::step) */
    static int64_t stride_of(const Mat& m) {
        return m.rows > 1 ? (int64_t)(reinterpret_cast<const char*>(m.template ptr<float>(1)) - reinterpret_cast<const char*>(m.template ptr<float>(0)))
                          : (int64_t)VSM_DIM * (int64_t)sizeof(float);
    }
    /* contiguous N x 256 fp32 rows of m (copied only if m is a strided view) */
    static const float* rows_of(const Mat& m, std::vector<float>& tmp) {
        need_f32_256(m);
        if (m.isContinuous()) return m.template ptr<float>(0);
        tmp.resize((size_t)m.rows * VSM_DIM);
        for (int r = 0; r < m.rows; r++) std::memcpy(&tmp[(size_t)r * VSM_DIM], m.template ptr<float>(r), VSM_DIM * sizeof(float));
        return tmp.data();
    }
};

}  // namespace vsm_cv
#endif /* VSM_CV_HPP */
