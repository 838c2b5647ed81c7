// loop_search.cpp -- LoopCloser::detect's candidate search (src/LoopCloser.cpp:43-62) from C++ through
// vsm_loop_detect_compact: nkf keyframes x rows descriptors resident on the device, nq query descriptors
// from pinned host memory, eligibility rules, ratio test, >= 30 gate; only surviving lists come back.
//   loop_search [nkf=500] [rows=1000] [nq=1000] [every=1] [steps=200] [n_gpus=1]
// With n_gpus > 1 the keyframes are split over the GPUs of the box (whole keyframes, contiguous blocks) and the
// search goes through vsm_group_loop_detect_compact -- still one C++ thread.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#include "vsm.h"

// queries: random unit rows; the first nq/5 re-observe rows of keyframe kf (device pointer kf_rows = its first row)
static void make_queries(vsm_ctx* ctx, float* q, int nq, const float* kf_rows) {
    float* dq = nullptr;
    cudaMalloc(reinterpret_cast<void**>(&dq), (size_t)nq * 1024);
    vsm_synth_rows_device(ctx, dq, 0, nq, 100);
    cudaMemcpy(q, dq, (size_t)nq * 1024, cudaMemcpyDeviceToHost);
    const int np = nq / 5;
    std::vector<float> src((size_t)np * 256), noise((size_t)np * 256);
    cudaMemcpy(src.data(), kf_rows + (size_t)3 * 256, src.size() * 4, cudaMemcpyDeviceToHost);
    vsm_synth_rows_device(ctx, dq, 0, np, 101);
    cudaMemcpy(noise.data(), dq, noise.size() * 4, cudaMemcpyDeviceToHost);
    for (int i = 0; i < np; i++) {
        double nn = 0;
        float* o = q + (size_t)i * 256;
        for (int c = 0; c < 256; c++) { o[c] = src[(size_t)i * 256 + c] + 0.9f * noise[(size_t)i * 256 + c]; nn += (double)o[c] * o[c]; }
        for (int c = 0; c < 256; c++) o[c] = (float)(o[c] / std::sqrt(nn));
    }
    cudaFree(dq);
}

static int group_main(int nkf, int rows, int nq, int every, int steps, int ngpu) {
    std::vector<int32_t> devs(ngpu);
    for (int i = 0; i < ngpu; i++) devs[i] = i;
    vsm_group* g = nullptr;
    if (vsm_group_create(devs.data(), ngpu, nullptr, &g) != VSM_OK) { std::fprintf(stderr, "%s\n", vsm_group_last_error(nullptr)); return 1; }
    auto die = [&](const char* what) { std::fprintf(stderr, "%s: %s\n", what, vsm_group_last_error(g)); std::exit(1); };
    float* q = nullptr;
    vsm_host_alloc(reinterpret_cast<void**>(&q), (int64_t)nq * 1024);
    const int kf_hit = (nkf / 2 / every) * every + every - 1;
    std::vector<float*> shard(ngpu, nullptr);
    for (int r = 0; r < ngpu; r++) {
        const int k0 = (int)((long long)nkf * r / ngpu), k1 = (int)((long long)nkf * (r + 1) / ngpu);
        const long long n = (long long)(k1 - k0) * rows;
        cudaSetDevice(r);
        if (cudaMalloc(reinterpret_cast<void**>(&shard[r]), (size_t)n * 1024) != cudaSuccess) return 2;
        if (vsm_synth_rows_device(vsm_group_ctx(g, r), shard[r], (long long)k0 * rows, n, 99) != VSM_OK) die("synth");
        if (kf_hit >= k0 && kf_hit < k1) make_queries(vsm_group_ctx(g, r), q, nq, shard[r] + (size_t)(kf_hit - k0) * rows * 256);
        std::vector<int64_t> seg(k1 - k0 + 1);
        for (int s = 0; s <= k1 - k0; s++) seg[s] = (int64_t)s * rows;
        if (vsm_group_adopt_device(g, r, shard[r], n, seg.data(), k1 - k0) != VSM_OK) die("adopt");
        std::vector<int32_t> ids(k1 - k0);
        for (int s = 0; s < k1 - k0; s++) ids[s] = k0 + s;
        vsm_store_set_frame_ids(vsm_group_ctx(g, r), ids.data(), k1 - k0);
    }
    std::vector<int32_t> status(nkf);
    std::vector<vsm_loop_candidate> cands(64);
    std::vector<vsm_dmatch> matches((size_t)64 * nq);
    int32_t nc = 0;
    int64_t nm = 0;
    auto call = [&]() {
        if (vsm_group_loop_detect_compact(g, nkf + 1000, 200, every, q, nq, 0.75f, 30, status.data(), cands.data(), 64, &nc,
                                          matches.data(), (int64_t)matches.size(), &nm) != VSM_OK) die("loop_detect");
    };
    for (int w = 0; w < 10; w++) call();
    for (int r = 0; r < ngpu; r++) vsm_set_profiling(vsm_group_ctx(g, r), 0);
    std::vector<double> ms(steps);
    for (int s = 0; s < steps; s++) {
        const auto t0 = std::chrono::steady_clock::now();
        call();
        ms[s] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    std::sort(ms.begin(), ms.end());
    int matched = 0;
    for (int s = 0; s < nkf; s++) matched += status[s] >= 0;
    const double flop = 2.0 * nq * (double)matched * rows * 256;
    std::printf("{\"bench\": \"loop_search\", \"api\": \"vsm_group_loop_detect_compact from ONE C++ thread\", \"n_gpus\": %d, "
                "\"keyframes\": %d, \"rows_per_keyframe\": %d, \"nq\": %d, \"every\": %d, \"keyframes_matched\": %d, \"p50_ms\": %.4f, "
                "\"p99_ms\": %.4f, \"min_ms\": %.4f, \"tflops_e2e_p50\": %.1f, \"candidates\": %d, \"candidate0\": %d, \"survivors\": %lld}\n",
                ngpu, nkf, rows, nq, every, matched, ms[steps / 2], ms[(size_t)(steps * 0.99)], ms[0],
                flop / (ms[steps / 2] * 1e-3) / 1e12, nc, nc ? cands[0].keyframe : -1, (long long)nm);
    vsm_group_destroy(g);
    for (int r = 0; r < ngpu; r++) { cudaSetDevice(r); cudaFree(shard[r]); }
    vsm_host_free(q);
    return 0;
}

int main(int argc, char** argv) {
    const int nkf = argc > 1 ? std::atoi(argv[1]) : 500;
    const int rows = argc > 2 ? std::atoi(argv[2]) : 1000;
    const int nq = argc > 3 ? std::atoi(argv[3]) : 1000;
    const int every = argc > 4 ? std::atoi(argv[4]) : 1;
    const int steps = argc > 5 ? std::atoi(argv[5]) : 200;
    const int ngpu = argc > 6 ? std::atoi(argv[6]) : 1;
    if (ngpu > 1) return group_main(nkf, rows, nq, every, steps, ngpu);
    vsm_ctx* ctx = nullptr;
    if (vsm_create(nullptr, &ctx) != VSM_OK) { std::fprintf(stderr, "%s\n", vsm_last_error(nullptr)); return 1; }
    auto die = [&](const char* what) { std::fprintf(stderr, "%s: %s\n", what, vsm_last_error(ctx)); std::exit(1); };
    const long long total = (long long)nkf * rows;
    float* db = nullptr;
    if (cudaMalloc(reinterpret_cast<void**>(&db), (size_t)total * 1024) != cudaSuccess) return 2;
    if (vsm_synth_rows_device(ctx, db, 0, total, 99) != VSM_OK) die("synth db");
    float* q = nullptr;
    vsm_host_alloc(reinterpret_cast<void**>(&q), (int64_t)nq * 1024);
    {
        float* dq = nullptr;
        cudaMalloc(reinterpret_cast<void**>(&dq), (size_t)nq * 1024);
        vsm_synth_rows_device(ctx, dq, 0, nq, 100);
        cudaMemcpy(q, dq, (size_t)nq * 1024, cudaMemcpyDeviceToHost);
        // the first nq/5 queries re-observe rows of one eligible keyframe
        const int kf = (nkf / 2 / every) * every + every - 1, np = nq / 5;
        std::vector<float> src((size_t)np * 256), noise((size_t)np * 256);
        cudaMemcpy(src.data(), db + ((size_t)kf * rows + 3) * 256, src.size() * 4, cudaMemcpyDeviceToHost);
        vsm_synth_rows_device(ctx, dq, 0, np, 101);
        cudaMemcpy(noise.data(), dq, noise.size() * 4, cudaMemcpyDeviceToHost);
        for (int i = 0; i < np; i++) {
            double nn = 0;
            float* o = q + (size_t)i * 256;
            for (int c = 0; c < 256; c++) { o[c] = src[(size_t)i * 256 + c] + 0.9f * noise[(size_t)i * 256 + c]; nn += (double)o[c] * o[c]; }
            for (int c = 0; c < 256; c++) o[c] = (float)(o[c] / std::sqrt(nn));
        }
        cudaFree(dq);
    }
    std::vector<int64_t> seg(nkf + 1);
    for (int s = 0; s <= nkf; s++) seg[s] = (int64_t)s * rows;
    if (vsm_store_adopt_device(ctx, db, total, seg.data(), nkf) != VSM_OK) die("adopt");
    std::vector<int32_t> status(nkf);
    std::vector<vsm_loop_candidate> cands(64);
    std::vector<vsm_dmatch> matches((size_t)64 * nq);
    int32_t nc = 0;
    int64_t nm = 0;
    auto call = [&]() {
        if (vsm_loop_detect_compact(ctx, nkf + 1000, 200, every, 0, q, nq, 0.75f, 30, status.data(), cands.data(), 64, &nc,
                                    matches.data(), (int64_t)matches.size(), &nm, nullptr) != VSM_OK) die("loop_detect");
    };
    for (int w = 0; w < 10; w++) call();
    vsm_stats st;
    vsm_get_stats(ctx, &st);
    vsm_set_profiling(ctx, 0);
    std::vector<double> ms(steps);
    for (int s = 0; s < steps; s++) {
        const auto t0 = std::chrono::steady_clock::now();
        call();
        ms[s] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    std::sort(ms.begin(), ms.end());
    int matched = 0;
    for (int s = 0; s < nkf; s++) matched += status[s] >= 0;
    const double flop = 2.0 * nq * (double)matched * rows * 256;
    std::printf("{\"bench\": \"loop_search\", \"api\": \"vsm_loop_detect_compact from C++ (pinned host queries in, surviving lists out)\", "
                "\"keyframes\": %d, \"rows_per_keyframe\": %d, \"nq\": %d, \"every\": %d, \"keyframes_matched\": %d, \"p50_ms\": %.4f, "
                "\"p99_ms\": %.4f, \"min_ms\": %.4f, \"tflops_e2e_p50\": %.1f, \"device_ms\": %.4f, \"tc_ms\": %.4f, \"after_tc_ms\": %.4f, "
                "\"launches\": %lld, \"candidates\": %d, \"survivors\": %lld}\n",
                nkf, rows, nq, every, matched, ms[steps / 2], ms[(size_t)(steps * 0.99)], ms[0], flop / (ms[steps / 2] * 1e-3) / 1e12,
                st.device_ms, st.tc_ms, st.select_ms, (long long)st.kernel_launches, nc, (long long)nm);
    vsm_destroy(ctx);
    cudaFree(db);
    vsm_host_free(q);
    return 0;
}
