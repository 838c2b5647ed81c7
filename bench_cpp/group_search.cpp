// group_search.cpp -- BASELINE configs[3] from ONE C++ thread: 2000 query descriptors against a
// 20M-row keyframe database dealt to the N GPUs of the box, through vsm_group_db_top2 (host queries
// in, host top-2 out; include/vsm.h).  What a single-threaded caller such as the reference's
// slam_thread (src/main.cpp:1520) gets, next to bench.py's one-process-per-GPU number.
//
//   group_search <n_gpus> [rows_total=20000000] [nq=2000] [steps=20]
// Prints one JSON line.  Build: g++ -O2 -std=c++17 -I include -I /usr/local/cuda/include
//   bench_cpp/group_search.cpp -L <libdir> -lvsm -L /usr/local/cuda/lib64 -lcudart
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

#include "vsm.h"

static double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char** argv) {
    const int ngpu = argc > 1 ? std::atoi(argv[1]) : 1;
    const long long rows_total = argc > 2 ? std::atoll(argv[2]) : 20000000LL;
    const int nq = argc > 3 ? std::atoi(argv[3]) : 2000;
    const int steps = argc > 4 ? std::atoi(argv[4]) : 20;
    const int n_planted = nq / 5;
    std::vector<int32_t> devs(ngpu);
    for (int i = 0; i < ngpu; i++) devs[i] = i;
    vsm_group* g = nullptr;
    if (vsm_group_create(devs.data(), ngpu, nullptr, &g) != VSM_OK) { std::fprintf(stderr, "%s\n", vsm_group_last_error(nullptr)); return 1; }
    auto die = [&](const char* what) { std::fprintf(stderr, "%s: %s\n", what, vsm_group_last_error(g)); std::exit(1); };

    // queries (pinned): random unit rows; the first n_planted are noisy re-observations of database rows
    float* q = nullptr;
    if (vsm_host_alloc(reinterpret_cast<void**>(&q), (int64_t)nq * 256 * sizeof(float)) != VSM_OK) return 2;
    std::vector<float*> shard(ngpu, nullptr);
    std::vector<long long> off(ngpu + 1, 0);
    for (int r = 0; r < ngpu; r++) off[r + 1] = rows_total * (r + 1) / ngpu;
    const long long stride = rows_total / n_planted;
    {
        float* dq = nullptr;
        cudaSetDevice(0);
        cudaMalloc(reinterpret_cast<void**>(&dq), (size_t)nq * 1024);
        if (vsm_synth_rows_device(vsm_group_ctx(g, 0), dq, 0, nq, 777) != VSM_OK) die("synth queries");
        cudaMemcpy(q, dq, (size_t)nq * 1024, cudaMemcpyDeviceToHost);
        if (vsm_synth_rows_device(vsm_group_ctx(g, 0), dq, 0, n_planted, 778) != VSM_OK) die("synth noise");
        std::vector<float> noise((size_t)n_planted * 256);
        cudaMemcpy(noise.data(), dq, noise.size() * 4, cudaMemcpyDeviceToHost);
        cudaFree(dq);
        for (int r = 0; r < ngpu; r++) {
            const long long n = off[r + 1] - off[r];
            cudaSetDevice(r);
            if (cudaMalloc(reinterpret_cast<void**>(&shard[r]), (size_t)n * 1024) != cudaSuccess) { std::fprintf(stderr, "cudaMalloc shard\n"); return 3; }
            if (vsm_synth_rows_device(vsm_group_ctx(g, r), shard[r], off[r], n, 4242) != VSM_OK) die("synth shard");
            // plant: database row i*stride+17 re-observes query i (sigma 0.05 x 16 = 0.8 of a unit row's component scale)
            for (int i = 0; i < n_planted; i++) {
                const long long row = (long long)i * stride + 17;
                if (row < off[r] || row >= off[r + 1]) continue;
                float v[256];
                double nn = 0;
                for (int c = 0; c < 256; c++) { v[c] = q[(size_t)i * 256 + c] + 0.8f * noise[(size_t)i * 256 + c]; nn += (double)v[c] * v[c]; }
                for (int c = 0; c < 256; c++) v[c] = (float)(v[c] / std::sqrt(nn));
                cudaMemcpy(shard[r] + (size_t)(row - off[r]) * 256, v, 1024, cudaMemcpyHostToDevice);
            }
            if (vsm_group_adopt_device(g, r, shard[r], n, nullptr, 0) != VSM_OK) die("adopt");
        }
    }
    std::vector<int64_t> idx((size_t)nq * 2);
    std::vector<float> dist((size_t)nq * 2);
    for (int w = 0; w < 5; w++)
        if (vsm_group_db_top2(g, q, nq, idx.data(), dist.data(), nullptr, nullptr) != VSM_OK) die("warm-up");
    std::vector<double> ms(steps);
    const double t_all = now_ms();
    for (int s = 0; s < steps; s++) {
        const double t0 = now_ms();
        if (vsm_group_db_top2(g, q, nq, idx.data(), dist.data(), nullptr, nullptr) != VSM_OK) die("search");
        ms[s] = now_ms() - t0;
    }
    const double total = now_ms() - t_all;
    int recovered = 0;
    for (int i = 0; i < n_planted; i++) recovered += idx[(size_t)2 * i] == (long long)i * stride + 17;
    unsigned long long chk = 0;
    for (size_t i = 0; i < idx.size(); i++) chk = chk * 1099511628211ull + (unsigned long long)idx[i];
    double tc_max = 0;
    for (int r = 0; r < ngpu; r++) {
        vsm_stats st;
        vsm_get_stats(vsm_group_ctx(g, r), &st);
        tc_max = std::max(tc_max, (double)st.tc_ms);
    }
    std::sort(ms.begin(), ms.end());
    const double flop = 2.0 * nq * (double)rows_total * 256;
    std::printf("{\"bench\": \"group_search\", \"api\": \"vsm_group_db_top2 (one C++ thread, host queries in, host top-2 out)\", "
                "\"n_gpus\": %d, \"db_rows\": %lld, \"nq\": %d, \"steps\": %d, \"ms_per_search_mean\": %.4f, \"ms_p50\": %.4f, "
                "\"ms_min\": %.4f, \"tflops_e2e\": %.1f, \"tc_kernel_ms_slowest_member\": %.4f, \"planted_recovered\": \"%d/%d\", "
                "\"checksum\": \"%016llx\"}\n",
                ngpu, rows_total, nq, steps, total / steps, ms[steps / 2], ms[0], flop / (total / steps * 1e-3) / 1e12, tc_max,
                recovered, n_planted, chk);
    vsm_group_destroy(g);
    for (int r = 0; r < ngpu; r++) { cudaSetDevice(r); cudaFree(shard[r]); }
    vsm_host_free(q);
    return 0;
}
