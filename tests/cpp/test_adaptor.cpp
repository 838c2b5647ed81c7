// The reference-shaped C++ call sites on top of libvsm.so, checked against the CPU oracle.
// Mirrors how Slam.cpp / LoopCloser.cpp would call the adaptor.
// Build + run: tests/test_cpp_adaptor.py (needs a B200).
#include <cmath>
#include <cstdio>
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

int main() {
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
        std::vector<DMatch> good = matcher.match_features(h0, n1, desc2);
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
    std::printf(fails ? "adaptor test: %d FAILURES\n" : "adaptor test: OK\n", fails);
    return fails ? 1 : 0;
}
