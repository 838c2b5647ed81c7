// Tracking-step latency measured from C++ through the same C ABI the reference would call
// (BASELINE configs[1]: consecutive 1000x1000 frame pairs, mutual-NN + ratio 0.75).
// Host buffers are pinned (vsm_host_alloc); every call uploads the current frame, matches it
// against the resident previous frame and returns the DMatch list to the host.
//   g++ -O2 -std=c++17 -I include bench_cpp/track_latency.cpp -L <libdir> -lvsm -o track_latency
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "vsm.h"

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static inline double uniform() {
    rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
    return (double)(rng_state >> 11) * (1.0 / 9007199254740992.0);
}
static inline float gauss() {
    double u1 = uniform() + 1e-300, u2 = uniform();
    return (float)(std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2));
}
static void normalize(float* r) {
    double n = 0;
    for (int c = 0; c < 256; c++) n += (double)r[c] * r[c];
    const float inv = (float)(1.0 / std::sqrt(n));
    for (int c = 0; c < 256; c++) r[c] *= inv;
}

int main(int argc, char** argv) {
    const int npairs = argc > 1 ? std::atoi(argv[1]) : 2544;
    const int n = 1000, nframes = 64;                       // 64 distinct frames, cycled
    float* frames = nullptr;
    if (vsm_host_alloc(reinterpret_cast<void**>(&frames), (int64_t)nframes * n * 256 * sizeof(float)) != VSM_OK) return 2;
    // frame f+1 re-observes 60 % of frame f with noise (sigma 0.06), the rest is new
    for (int r = 0; r < n; r++) { float* p = frames + (size_t)r * 256; for (int c = 0; c < 256; c++) p[c] = gauss(); normalize(p); }
    for (int f = 1; f < nframes; f++) {
        const float* prev = frames + (size_t)(f - 1) * n * 256;
        float* cur = frames + (size_t)f * n * 256;
        for (int r = 0; r < n; r++) {
            float* p = cur + (size_t)r * 256;
            if (r < 600) { const float* q = prev + (size_t)((r * 7 + f) % n) * 256; for (int c = 0; c < 256; c++) p[c] = q[c] + 0.06f * gauss(); }
            else for (int c = 0; c < 256; c++) p[c] = gauss();
            normalize(p);
        }
    }
    vsm_opts o;
    vsm_default_opts(&o);
    // NO pre-sized store: tracked frames are plain frames whose rows are recycled; every 5th frame is
    // promoted to a keyframe (Frame::set_keyframe, src/Slam.cpp:1076), so the store grows while we time
    vsm_ctx* ctx = nullptr;
    if (vsm_create(&o, &ctx) != VSM_OK) { std::fprintf(stderr, "%s\n", vsm_last_error(nullptr)); return 1; }
    vsm_set_profiling(ctx, 0);
    std::vector<vsm_dmatch> good(n), raw(n);
    int32_t ng = 0, nr = 0, h = -1, h2 = -1;
    // VSM_TRACK_RAW=1: also ask for the raw list, as the reference's own tracking call does (src/Slam.cpp:839-844)
    const bool want_raw = std::getenv("VSM_TRACK_RAW") && std::atoi(std::getenv("VSM_TRACK_RAW"));
    vsm_dmatch* rawp = want_raw ? raw.data() : nullptr;
    int32_t* nrp = want_raw ? &nr : nullptr;
    auto fail = [&](const char* what) { std::fprintf(stderr, "%s: %s\n", what, vsm_last_error(ctx)); std::exit(1); };
    if (vsm_track(ctx, -1, 0, frames, n, 0.75f, 1, good.data(), &ng, nullptr, nullptr, &h) != VSM_OK) fail("first frame");
    for (int f = 1; f <= 20; f++) {                          // warm-up
        if (vsm_track(ctx, h, f, frames + (size_t)(f % nframes) * n * 256, n, 0.75f, 1, good.data(), &ng, rawp, nrp, &h2) != VSM_OK) fail("warm-up");
        h = h2;
    }
    std::vector<double> us(npairs);
    long long matches = 0;
    const auto t_all = std::chrono::steady_clock::now();
    for (int f = 0; f < npairs; f++) {
        const float* cur = frames + (size_t)((f + 21) % nframes) * n * 256;
        const auto t0 = std::chrono::steady_clock::now();
        if (vsm_track(ctx, h, f + 21, cur, n, 0.75f, 1, good.data(), &ng, rawp, nrp, &h2) != VSM_OK) fail("track");
        if (f % 5 == 4 && vsm_store_promote(ctx, h2) != VSM_OK) fail("promote");      // inside the timed step
        us[f] = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
        h = h2;
        matches += ng;
    }
    const double total_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_all).count();
    std::sort(us.begin(), us.end());
    int64_t rows = 0;
    int32_t nkf = 0;
    vsm_store_info(ctx, &rows, &nkf);
    std::printf("{\"pairs\": %d, \"p50_us\": %.2f, \"p99_us\": %.2f, \"min_us\": %.2f, \"pairs_per_s\": %.1f, \"matches_per_s\": %.1f, "
                "\"matches_per_pair\": %.1f, \"p999_us\": %.2f, \"max_us\": %.2f, \"keyframes\": %d, \"store_rows\": %lld, \"timing\": \"std::chrono around vsm_track (+ vsm_store_promote every 5th frame) in C++: pinned H2D of the current frame, match, D2H of the DMatch list; store NOT pre-sized\"}\n",
                npairs, us[npairs / 2], us[(size_t)(npairs * 0.99)], us[0], npairs / total_s, matches / total_s, (double)matches / npairs, us[(size_t)(npairs * 0.999)], us[npairs - 1], nkf, (long long)rows);
    vsm_destroy(ctx);
    vsm_host_free(frames);
    return 0;
}
