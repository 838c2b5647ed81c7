// The reference-shaped C++ call sites on top of libvsm.so, checked against the CPU oracle.
// Mirrors how Slam.cpp / LoopCloser.cpp would call the adaptor.
// Build + run: tests/test_cpp_adaptor.py (needs a B200).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "vsm_cv.hpp"
#include "../../oracle/vsm_oracle.h"

using vsm_cv::DMatch;
using vsm_cv::Mat;

static int fails = 0;
#define EXPECT(c)                                                  \
    do {                                                           \
        if (!(c)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); fails++; } \
    } while (0)

static std::vector<float> rows(uint64_t seed, uint64_t set, int n) {
    std::vector<float> v((size_t)n * 256);
    if (n) vsm_oracle_gen_rows(seed, set, 0, n, v.data());
    return v;
}

static bool same(const std::vector<DMatch>& a, const std::vector<vsm_oracle_dmatch>& b, int nb) {
    return (int)a.size() == nb && (nb == 0 || std::memcmp(a.data(), b.data(), (size_t)nb * 16) == 0);
}

static std::vector<int> g_devices;

int main(int argc, char** argv) {
    for (int i = 1; i < argc; i++) g_devices.push_back(std::atoi(argv[i]));
    vsm_cv::DescriptorMatcher matcher;              // like Slam's matcher_l2_ member (include/Slam.h:197)
    // frame B re-observes part of frame A: B = A + small noise on the first 300 rows
    const int n1 = 500, n2 = 640;
    std::vector<float> a = rows(3, 0, n1), b = rows(3, 1, n2), noise = rows(3, 2, 300);
    for (int r = 0; r < 300; r++) {
        double nn = 0;
        for (int c = 0; c < 256; c++) { float v = a[r * 256 + c] + 0.05f * noise[r * 256 + c]; b[(r + 40) * 256 + c] = v; nn += (double)v * v; }
        for (int c = 0; c < 256; c++) b[(r + 40) * 256 + c] = (float)(b[(r + 40) * 256 + c] / std::sqrt(nn));
    }
    Mat desc1(n1, 256, a.data()), desc2(n2, 256, b.data());

    for (int mutual = 0; mutual < 2; mutual++) {
        std::vector<DMatch> raw;
        std::vector<DMatch> good = matcher.match_features(desc1, desc2, &raw, 0.75f, mutual != 0);   // Slam.cpp:841
        std::vector<vsm_oracle_dmatch> og(n1), orw(n1);
        int ng = 0, nr = 0;
        vsm_oracle_match_features(a.data(), n1, b.data(), n2, 0.75f, mutual, og.data(), &ng, orw.data(), &nr, 0);
        EXPECT(ng > 200);
        EXPECT(same(good, og, ng));
        EXPECT(same(raw, orw, nr));
    }
    // empty inputs return empty vectors (Slam.cpp:1143)
    EXPECT(matcher.match_features(Mat(), desc2).empty());
    EXPECT(matcher.match_features(desc1, Mat()).empty());

    // knnMatch + the reference's own loop (LoopCloser.cpp:54-60)
    std::vector<std::vector<DMatch>> knn;
    matcher.knnMatch(desc1, desc2, knn, 2);
    std::vector<DMatch> loop_good;
    for (auto& m : knn)
        if (m.size() >= 2 && m[0].distance < 0.75f * m[1].distance) loop_good.push_back(m[0]);
    {
        std::vector<vsm_oracle_dmatch> og(n1);
        int ng = 0;
        vsm_oracle_match_features(a.data(), n1, b.data(), n2, 0.75f, 0, og.data(), &ng, nullptr, nullptr, 0);
        EXPECT(same(loop_good, og, ng));
    }
    // a strided (non-continuous) view is accepted like a cv::Mat ROI
    {
        std::vector<float> wide((size_t)n1 * 300, 0.f);
        for (int r = 0; r < n1; r++) std::memcpy(&wide[(size_t)r * 300], &a[(size_t)r * 256], 1024);
        Mat view(n1, 256, wide.data(), 300 * sizeof(float));
        std::vector<DMatch> g1 = matcher.match_features(view, desc2), g2 = matcher.match_features(desc1, desc2);
        EXPECT(g1.size() == g2.size() && std::memcmp(g1.data(), g2.data(), g1.size() * 16) == 0);
    }
    // keyframe store: resident reference keyframe + LoopCloser block
    int h0 = matcher.add_keyframe(10, desc1);
    std::vector<float> c = rows(4, 5, 333);
    int h1 = matcher.add_keyframe(20, Mat(333, 256, c.data()));
    EXPECT(h0 == 0 && h1 == 1);
    {
        std::vector<DMatch> good = matcher.match_features(h0, desc2);
        std::vector<vsm_oracle_dmatch> og(n1);
        int ng = 0;
        vsm_oracle_match_features(a.data(), n1, b.data(), n2, 0.75f, 0, og.data(), &ng, nullptr, nullptr, 0);
        EXPECT(same(good, og, ng));
        std::vector<std::vector<DMatch>> per_kf;
        matcher.detect_candidates(desc2, 0.75f, per_kf);
        EXPECT(per_kf.size() == 2);
        std::vector<float> db(a);
        db.insert(db.end(), c.begin(), c.end());
        int64_t seg[3] = {0, n1, n1 + 333};
        int32_t counts[2];
        std::vector<vsm_oracle_dmatch> om((size_t)2 * n2);
        vsm_oracle_segmented(b.data(), n2, db.data(), seg, 2, 0.75f, counts, om.data(), 0);
        for (int s = 0; s < 2; s++) {
            std::vector<vsm_oracle_dmatch> os(om.begin() + (size_t)s * n2, om.begin() + (size_t)s * n2 + counts[s]);
            EXPECT(same(per_kf[s], os, counts[s]));
        }
        EXPECT(counts[0] >= 30);                      // the loop-closure gate at LoopCloser.cpp:62 would pass
    }
    // LoopCloser::detect with its eligibility rules (LoopCloser.cpp:43-48), restated here as the caller's
    // loop; and the map-point search over a re-stacked subset (Slam.cpp:546-574)
    {
        matcher.clear_keyframes();
        const int nkf = 12, kf_rows = 150;
        std::vector<float> db = rows(7, 1, nkf * kf_rows);
        for (int s = 0; s < nkf; s++) matcher.add_keyframe(30 * s, Mat(kf_rows, 256, &db[(size_t)s * kf_rows * 256]));
        const int nq = 90;
        std::vector<float> q = rows(7, 2, nq);
        std::memcpy(q.data(), &db[(size_t)(5 * kf_rows + 3) * 256], 40 * 1024);        // re-observe 40 rows of keyframe 5
        Mat cur(nq, 256, q.data());
        std::vector<int64_t> seg(nkf + 1);
        for (int s = 0; s <= nkf; s++) seg[s] = (int64_t)s * kf_rows;
        std::vector<int32_t> counts(nkf);
        std::vector<vsm_oracle_dmatch> om((size_t)nkf * nq);
        vsm_oracle_segmented(q.data(), nq, db.data(), seg.data(), nkf, 0.75f, counts.data(), om.data(), 0);
        const int cur_id = 400, min_gap = 200, every = 2;
        std::vector<int> status;
        std::vector<std::vector<DMatch>> per_kf;
        matcher.detect_loop_candidates(cur_id, cur, 0.75f, min_gap, every, status, per_kf);
        int checked = 0, matched = 0;
        for (int s = 0; s < nkf; s++) {
            bool eligible = false;
            if (!(cur_id - 30 * s < min_gap)) { checked++; eligible = checked % every == 0; }
            if (!eligible) { EXPECT(status[s] == -1 && per_kf[s].empty()); continue; }
            matched++;
            std::vector<vsm_oracle_dmatch> os(om.begin() + (size_t)s * nq, om.begin() + (size_t)s * nq + counts[s]);
            EXPECT(status[s] == counts[s]);
            EXPECT(same(per_kf[s], os, counts[s]));
        }
        EXPECT(matched == 3);                                   // keyframes 0..6 pass the gap rule, every 2nd of them
        // the compact form: the gate (LoopCloser.cpp:62) on the device, only surviving lists come back
        for (int min_matches : {30, 1}) {
            std::vector<int> st2;
            std::vector<vsm_cv::DescriptorMatcher::LoopCandidate> cands;
            matcher.detect_loop(cur_id, cur, 0.75f, min_gap, every, min_matches, st2, cands);
            EXPECT(st2 == status);
            size_t k = 0;
            for (int s = 0; s < nkf; s++) {
                if (status[s] < min_matches || status[s] <= 0) continue;
                EXPECT(k < cands.size() && cands[k].keyframe == s);
                if (k < cands.size()) {
                    std::vector<vsm_oracle_dmatch> os(om.begin() + (size_t)s * nq, om.begin() + (size_t)s * nq + counts[s]);
                    EXPECT(same(cands[k].good_matches, os, counts[s]));
                }
                k++;
            }
            EXPECT(k == cands.size());
        }
        // pairs of stored keyframes in one call = the per-pair calls
        {
            std::vector<int> qh = {0, 5, 11, 3}, th = {1, 0, 11, 7};
            for (int mutual = 0; mutual < 2; mutual++) {
                std::vector<std::vector<DMatch>> batch = matcher.match_features_batch(qh, th, 0.75f, mutual != 0);
                EXPECT(batch.size() == 4);
                for (size_t p = 0; p < qh.size(); p++) {
                    std::vector<vsm_oracle_dmatch> og(kf_rows);
                    int ng = 0;
                    vsm_oracle_match_features(&db[(size_t)qh[p] * kf_rows * 256], kf_rows, &db[(size_t)th[p] * kf_rows * 256], kf_rows,
                                              0.75f, mutual, og.data(), &ng, nullptr, nullptr, 0);
                    EXPECT(same(batch[p], og, ng));
                }
            }
        }
        // subset search: every third row is a valid map point
        std::vector<unsigned char> valid((size_t)nkf * kf_rows, 0);
        std::vector<float> sub;
        std::vector<int> ids;
        for (int r = 0; r < nkf * kf_rows; r += 3) { valid[r] = 1; ids.push_back(r); sub.insert(sub.end(), &db[(size_t)r * 256], &db[(size_t)r * 256 + 256]); }
        std::vector<std::vector<DMatch>> knn2;
        matcher.search_store(cur, valid, knn2);
        std::vector<int64_t> oi((size_t)nq * 2);
        std::vector<float> od((size_t)nq * 2);
        vsm_oracle_knn(q.data(), nq, 256, sub.data(), (int64_t)ids.size(), 256, 2, oi.data(), od.data(), 0);
        for (int i = 0; i < nq; i++) {
            EXPECT(knn2[i].size() == 2);
            for (int k = 0; k < 2 && knn2[i].size() == 2; k++) {
                EXPECT(knn2[i][k].trainIdx == ids[(size_t)oi[2 * i + k]]);
                EXPECT(std::memcmp(&knn2[i][k].distance, &od[2 * i + k], 4) == 0);
            }
        }
    }
    // the resident map-point table through the adaptor: points born from rows of two stored keyframes, one
    // observed again later, some invalidated; both searches of Slam.cpp (:546-574 all valid, :744-774 near a frame)
    {
        matcher.clear_keyframes();
        std::vector<float> k0 = rows(9, 1, 200), k1 = rows(9, 2, 200);
        const int h0 = matcher.add_keyframe(100, Mat(200, 256, k0.data()));
        const int h1 = matcher.add_keyframe(300, Mat(200, 256, k1.data()));
        std::vector<int> kp0, kp1;
        for (int i = 0; i < 200; i += 2) kp0.push_back(i);
        for (int i = 1; i < 200; i += 4) kp1.push_back(i);
        EXPECT(matcher.add_map_points(h0, kp0) == 0);
        EXPECT(matcher.add_map_points(h1, kp1) == (int)kp0.size());
        matcher.observe_map_points({3, 4, 5}, 310);                         // points of keyframe 100 seen again near frame 300
        matcher.set_map_points_valid({0, 1, 2, 4}, false);
        // the model: descriptors / valid / observation frames per point id
        std::vector<const float*> pd;
        std::vector<std::vector<int>> obs;
        for (int k : kp0) { pd.push_back(&k0[(size_t)k * 256]); obs.push_back({100}); }
        for (int k : kp1) { pd.push_back(&k1[(size_t)k * 256]); obs.push_back({300}); }
        for (int p : {3, 4, 5}) obs[p].push_back(310);
        std::vector<char> valid(pd.size(), 1);
        for (int p : {0, 1, 2, 4}) valid[p] = 0;
        const int nq = 60;
        std::vector<float> q = rows(9, 3, nq);
        std::memcpy(q.data(), k0.data() + (size_t)6 * 256, 1024);            // query 0 = point 3's descriptor (keypoint 6 of keyframe 100)
        std::memcpy(q.data() + 256, k1.data() + (size_t)5 * 256, 1024);      // query 1 = a point of keyframe 300
        Mat fq(nq, 256, q.data());
        for (int near : {-1, 300}) {
            std::vector<int> ids;
            std::vector<float> sub;
            for (size_t p = 0; p < pd.size(); p++) {
                bool sel = valid[p];
                if (sel && near >= 0) {
                    sel = false;
                    for (int f : obs[p]) sel |= std::abs(f - near) < 30;
                }
                if (sel) { ids.push_back((int)p); sub.insert(sub.end(), pd[p], pd[p] + 256); }
            }
            std::vector<std::vector<DMatch>> knn4;
            const int stacked = matcher.search_map_points(fq, near, 30, knn4);
            EXPECT(stacked == (int)ids.size());
            std::vector<int64_t> oi((size_t)nq * 2);
            std::vector<float> od((size_t)nq * 2);
            vsm_oracle_knn(q.data(), nq, 256, sub.data(), (int64_t)ids.size(), 256, 2, oi.data(), od.data(), 0);
            for (int i = 0; i < nq; i++) {
                EXPECT(knn4[i].size() == 2);
                for (int k = 0; k < 2 && knn4[i].size() == 2; k++) {
                    EXPECT(knn4[i][k].trainIdx == ids[(size_t)oi[2 * i + k]]);
                    EXPECT(std::memcmp(&knn4[i][k].distance, &od[2 * i + k], 4) == 0);
                }
            }
            EXPECT(knn4[0][0].trainIdx == 3 && knn4[0][0].distance == 0.f);   // seen at frames 100 and 310: selected in both searches
        }
        vsm_points_clear(matcher.handle());
        matcher.clear_keyframes();
    }
    // several devices behind one matcher object (vsm_group): device list from the command line,
    // e.g. "0 0 0" (three contexts on one GPU) or "0 1"
    if (g_devices.size() > 1) {
        vsm_cv::DescriptorMatcher multi(g_devices);
        const int nkf = 14, kf_rows = 120, nq = 100;
        std::vector<float> db = rows(8, 1, nkf * kf_rows), q = rows(8, 2, nq);
        std::memcpy(q.data(), &db[(size_t)(9 * kf_rows + 5) * 256], 50 * 1024);          // re-observe 50 rows of keyframe 9
        for (int s = 0; s < nkf; s++) EXPECT(multi.add_keyframe(25 * s, Mat(kf_rows, 256, &db[(size_t)s * kf_rows * 256])) == s);
        Mat cur(nq, 256, q.data());
        std::vector<std::vector<DMatch>> knn3;
        multi.search_store(cur, knn3);
        std::vector<int64_t> oi((size_t)nq * 2);
        std::vector<float> od((size_t)nq * 2);
        vsm_oracle_knn(q.data(), nq, 256, db.data(), (int64_t)nkf * kf_rows, 256, 2, oi.data(), od.data(), 0);
        for (int i = 0; i < nq; i++) {
            EXPECT(knn3[i].size() == 2);
            for (int k = 0; k < 2 && knn3[i].size() == 2; k++) {
                EXPECT(knn3[i][k].trainIdx == (int)oi[2 * i + k]);
                EXPECT(std::memcmp(&knn3[i][k].distance, &od[2 * i + k], 4) == 0);
            }
        }
        std::vector<int64_t> seg(nkf + 1);
        for (int s = 0; s <= nkf; s++) seg[s] = (int64_t)s * kf_rows;
        std::vector<int32_t> counts(nkf);
        std::vector<vsm_oracle_dmatch> om((size_t)nkf * nq);
        vsm_oracle_segmented(q.data(), nq, db.data(), seg.data(), nkf, 0.75f, counts.data(), om.data(), 0);
        std::vector<int> st3;
        std::vector<vsm_cv::DescriptorMatcher::LoopCandidate> cands;
        multi.detect_loop(1000, cur, 0.75f, 200, 1, 30, st3, cands);
        EXPECT(cands.size() == 1 && cands[0].keyframe == 9);
        for (int s = 0; s < nkf; s++) EXPECT(st3[s] == counts[s]);
        if (cands.size() == 1) {
            std::vector<vsm_oracle_dmatch> os(om.begin() + (size_t)9 * nq, om.begin() + (size_t)9 * nq + counts[9]);
            for (auto& m : os) m.imgIdx = 9;
            EXPECT(same(cands[0].good_matches, os, counts[9]));
        }
        std::printf("multi-device matcher on %zu contexts: checked\n", g_devices.size());
    }
    std::printf(fails ? "adaptor test: %d FAILURES\n" : "adaptor test: OK\n", fails);
    return fails ? 1 : 0;
}
