// vsm_group.inl -- several GPUs behind ONE caller thread (included by vsm_api.cu, inside extern "C").
//
// The reference matches on one thread of one process (src/main.cpp:1520, src/LoopCloser.cpp:16-18), so
// a drop-in that uses the 8 GPUs of a box cannot ask for one process per GPU.  A vsm_group owns one
// matching context per device and one worker thread per context; a group call fans the search out on
// the workers (each enqueues its device's kernels on that context's stream), joins, and merges on the
// first device:
//   * every member writes its exact local top-2 -- 64-bit keys carrying the STACKED row index of the
//     whole database -- straight into the first device's gather buffer (peer stores over NVLink,
//     cudaDeviceEnablePeerAccess; a pinned host buffer if a pair of devices has no peer path), then
//     records an event;
//   * the first device's stream waits for the events (stream order, no flags, no spinning kernel) and
//     runs merge_keys_kernel, whose output lands in pinned host memory.
// No CUDA IPC, no process group.  Keyframes are dealt to the members WHOLE (the per-keyframe search
// stays local to a device), always to the member that holds the fewest rows.

struct GroupWorker {
    std::thread th;
    std::mutex m;
    std::condition_variable cv;
    std::function<int()> task;
    std::atomic<int> state{0};               // 0 idle, 1 task posted, 2 done
    bool quit = false;
    int status = VSM_OK;
};

struct GKeyframe {
    int32_t member;
    int32_t local_handle;
    int64_t grow0;                           // first row in the stacked (global) numbering
    int32_t count;
    int32_t frame_id;
    uint8_t live;
};

struct vsm_group {
    int n = 0;
    std::vector<int> dev;
    std::vector<vsm_ctx*> ctx;
    std::vector<GroupWorker*> workers;
    std::vector<GKeyframe> kfs;              // by group handle (never reused), insertion order
    int64_t g_rows = 0;                      // stacked rows handed out so far
    std::vector<std::vector<int64_t>> h_tab; // per member: (local row0, count, global row0) triples
    std::vector<DevBuf<int64_t>> d_tab;
    std::vector<char> tab_dirty;
    std::vector<int32_t> tab_n;
    unsigned long long* gather = nullptr;    // [n][nq][2] keys on the first device (or pinned host)
    bool gather_on_host = false;
    int nq_cap = 0;
    std::vector<cudaEvent_t> ev;
    uint8_t* h_out = nullptr;
    size_t h_out_cap = 0;
    float* h_query = nullptr;
    size_t h_query_cap = 0;
    bool adopted = false;
    std::string err;
};

static thread_local std::string g_group_create_error;

static void group_worker_main(GroupWorker* w) {
    for (;;) {
        // spin briefly (back-to-back searches), then block
        int spins = 0;
        while (w->state.load(std::memory_order_acquire) != 1) {
            if (++spins < 2000) { std::this_thread::yield(); continue; }
            std::unique_lock<std::mutex> lk(w->m);
            w->cv.wait(lk, [&] { return w->state.load(std::memory_order_acquire) == 1 || w->quit; });
            if (w->quit) return;
        }
        if (w->quit) return;
        w->status = w->task();
        w->state.store(2, std::memory_order_release);
    }
}

static void group_post(GroupWorker* w, std::function<int()> fn) {
    w->task = std::move(fn);
    {
        std::lock_guard<std::mutex> lk(w->m);
        w->state.store(1, std::memory_order_release);
    }
    w->cv.notify_one();
}

static int group_wait(GroupWorker* w) {
    while (w->state.load(std::memory_order_acquire) != 2) std::this_thread::yield();
    w->state.store(0, std::memory_order_release);
    return w->status;
}

// Runs fn(r) for every member on its worker thread; returns the first failure (its message in g->err).
static int group_fan_out(vsm_group* g, const std::function<int(int)>& fn) {
    for (int r = 0; r < g->n; r++) group_post(g->workers[r], [&fn, r] { return fn(r); });
    int st = VSM_OK;
    for (int r = 0; r < g->n; r++) {
        const int s = group_wait(g->workers[r]);
        if (s != VSM_OK && st == VSM_OK) {
            st = s;
            g->err = "device " + std::to_string(g->dev[r]) + ": " + g->ctx[r]->err;
        }
    }
    return st;
}

#define GCK(call)                                                                                    \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            char b_[512];                                                                            \
            snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            g->err = b_;                                                                             \
            return VSM_ERR_CUDA;                                                                     \
        }                                                                                            \
    } while (0)

static int group_fail(vsm_group* g, int code, const char* msg) {
    g->err = msg;
    return code;
}

const char* vsm_group_last_error(const vsm_group* g) { return g ? g->err.c_str() : g_group_create_error.c_str(); }
int vsm_group_size(const vsm_group* g) { return g ? g->n : 0; }
vsm_ctx* vsm_group_ctx(vsm_group* g, int32_t member) { return (g && member >= 0 && member < g->n) ? g->ctx[member] : nullptr; }

void vsm_group_destroy(vsm_group* g) {
    if (!g) return;
    for (GroupWorker* w : g->workers) {
        {
            std::lock_guard<std::mutex> lk(w->m);
            w->quit = true;
            w->state.store(1, std::memory_order_release);
        }
        w->cv.notify_one();
        if (w->th.joinable()) w->th.join();
        delete w;
    }
    if (!g->ctx.empty() && g->ctx[0]) {
        cudaSetDevice(g->dev[0]);
        if (g->gather) { if (g->gather_on_host) cudaFreeHost(g->gather); else cudaFree(g->gather); }
    }
    for (int r = 0; r < (int)g->ctx.size(); r++) {
        if (!g->ctx[r]) continue;
        cudaSetDevice(g->dev[r]);
        if (r < (int)g->ev.size() && g->ev[r]) cudaEventDestroy(g->ev[r]);
        if (r < (int)g->d_tab.size() && g->d_tab[r].p) cudaFree(g->d_tab[r].p);
        vsm_destroy(g->ctx[r]);
    }
    if (g->h_out) cudaFreeHost(g->h_out);
    if (g->h_query) cudaFreeHost(g->h_query);
    delete g;
}

int vsm_group_create(const int32_t* devices, int32_t n, const vsm_opts* opts, vsm_group** out) {
    if (!out) return VSM_ERR_INVALID;
    *out = nullptr;
    if (!devices || n < 1 || n > XCHG_MAX_WORLD) { g_group_create_error = "vsm_group_create: 1..16 devices"; return VSM_ERR_INVALID; }
    vsm_group* g = new vsm_group();
    g->n = n;
    g->dev.assign(devices, devices + n);
    g->ctx.assign(n, nullptr);
    g->h_tab.resize(n); g->d_tab.resize(n); g->tab_dirty.assign(n, 0); g->tab_n.assign(n, 0);
    g->ev.assign(n, nullptr);
    auto bail = [&](int code, const std::string& msg) {
        g_group_create_error = msg;
        vsm_group_destroy(g);
        return code;
    };
    for (int r = 0; r < n; r++) {
        vsm_opts o;
        if (opts) o = *opts; else vsm_default_opts(&o);
        o.device = devices[r];
        const int st = vsm_create(&o, &g->ctx[r]);
        if (st != VSM_OK) return bail(st, std::string("vsm_group_create: ") + vsm_last_error(nullptr));
        if (cudaEventCreateWithFlags(&g->ev[r], cudaEventDisableTiming) != cudaSuccess) return bail(VSM_ERR_CUDA, "vsm_group_create: event");
    }
    // peer path from every member to the first device (the merge runs there)
    bool peer_ok = true;
    for (int r = 1; r < n; r++) {
        if (devices[r] == devices[0]) continue;
        int can = 0;
        cudaSetDevice(devices[r]);
        if (cudaDeviceCanAccessPeer(&can, devices[r], devices[0]) != cudaSuccess || !can) { cudaGetLastError(); peer_ok = false; continue; }
        const cudaError_t e = cudaDeviceEnablePeerAccess(devices[0], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) peer_ok = false;
        cudaGetLastError();
    }
    g->gather_on_host = !peer_ok;
    for (int r = 0; r < n; r++) {
        g->workers.push_back(new GroupWorker());
        g->workers[r]->th = std::thread(group_worker_main, g->workers[r]);
    }
    *out = g;
    return VSM_OK;
}

static int group_ensure_buffers(vsm_group* g, int nq) {
    GCK(cudaSetDevice(g->dev[0]));
    if (nq > g->nq_cap) {
        for (int r = 0; r < g->n; r++) { cudaSetDevice(g->dev[r]); cudaStreamSynchronize(g->ctx[r]->stream); }
        GCK(cudaSetDevice(g->dev[0]));
        if (g->gather) { if (g->gather_on_host) cudaFreeHost(g->gather); else cudaFree(g->gather); g->gather = nullptr; }
        const int cap = std::max(nq, 2048);
        const size_t bytes = (size_t)g->n * cap * 2 * sizeof(unsigned long long);
        if (g->gather_on_host) GCK(cudaHostAlloc(reinterpret_cast<void**>(&g->gather), bytes, cudaHostAllocPortable | cudaHostAllocMapped));
        else GCK(cudaMalloc(reinterpret_cast<void**>(&g->gather), bytes));
        g->nq_cap = cap;
    }
    const size_t ob = (size_t)nq * 2 * (sizeof(int64_t) + sizeof(float));
    if (ob > g->h_out_cap) {
        if (g->h_out) cudaFreeHost(g->h_out);
        g->h_out = nullptr; g->h_out_cap = 0;
        GCK(cudaHostAlloc(reinterpret_cast<void**>(&g->h_out), ob, cudaHostAllocPortable | cudaHostAllocMapped));
        g->h_out_cap = ob;
    }
    return VSM_OK;
}

// The caller's query rows where every member can DMA them from: the buffer itself if it is pinned,
// else one copy into the group's pinned staging buffer.
static int group_stage_query(vsm_group* g, const float* query, int nq, const float** staged) {
    if (host_pinned(query)) { *staged = query; return VSM_OK; }
    const size_t bytes = (size_t)nq * VSM_DIM * sizeof(float);
    if (bytes > g->h_query_cap) {
        for (int r = 0; r < g->n; r++) { cudaSetDevice(g->dev[r]); cudaStreamSynchronize(g->ctx[r]->stream); }
        if (g->h_query) cudaFreeHost(g->h_query);
        g->h_query = nullptr; g->h_query_cap = 0;
        GCK(cudaHostAlloc(reinterpret_cast<void**>(&g->h_query), bytes, cudaHostAllocPortable | cudaHostAllocMapped));
        g->h_query_cap = bytes;
    }
    memcpy(g->h_query, query, bytes);
    *staged = g->h_query;
    return VSM_OK;
}

static void group_note_keyframe(vsm_group* g, int r, int32_t local_handle, int32_t count, int32_t frame_id, int32_t* handle) {
    vsm_ctx* ctx = g->ctx[r];
    GKeyframe k = {r, local_handle, g->g_rows, count, frame_id, 1};
    g->kfs.push_back(k);
    if (count > 0) {
        g->h_tab[r].push_back(ctx->segs[local_handle].row0);
        g->h_tab[r].push_back(count);
        g->h_tab[r].push_back(g->g_rows);
        g->tab_dirty[r] = 1;
    }
    g->g_rows += count;
    if (handle) *handle = (int32_t)g->kfs.size() - 1;
}

int vsm_group_store_add(vsm_group* g, int32_t frame_id, const float* desc, int32_t n, int32_t* handle) {
    if (!g || n < 0 || (n > 0 && !desc)) return g ? group_fail(g, VSM_ERR_INVALID, "vsm_group_store_add: bad argument") : VSM_ERR_INVALID;
    if (g->adopted) return group_fail(g, VSM_ERR_INVALID, "vsm_group_store_add: the members adopted device matrices; clear first");
    if (g->g_rows + n > 0xFFFFFFFFll) return group_fail(g, VSM_ERR_CAPACITY, "vsm_group_store_add: more than 2^32 stacked rows");
    int r = 0;                                               // the member holding the fewest rows takes the keyframe, whole
    for (int k = 1; k < g->n; k++) if (g->ctx[k]->store_rows < g->ctx[r]->store_rows) r = k;
    int32_t h = -1;
    const int st = vsm_store_add(g->ctx[r], frame_id, desc, n, &h);
    if (st != VSM_OK) { g->err = g->ctx[r]->err; return st; }
    group_note_keyframe(g, r, h, n, frame_id, handle);
    return VSM_OK;
}

int vsm_group_store_remove(vsm_group* g, int32_t handle) {
    if (!g) return VSM_ERR_INVALID;
    if (handle < 0 || handle >= (int32_t)g->kfs.size() || !g->kfs[handle].live)
        return group_fail(g, VSM_ERR_NOT_FOUND, "vsm_group_store_remove: unknown keyframe handle");
    GKeyframe& k = g->kfs[handle];
    vsm_ctx* ctx = g->ctx[k.member];
    const int64_t row0 = ctx->segs[k.local_handle].row0;
    const int st = vsm_store_remove(ctx, k.local_handle);
    if (st != VSM_OK) { g->err = ctx->err; return st; }
    k.live = 0;
    std::vector<int64_t>& t = g->h_tab[k.member];
    for (size_t i = 0; i + 2 < t.size(); i += 3)
        if (t[i] == row0 && t[i + 2] == k.grow0) { t.erase(t.begin() + (long)i, t.begin() + (long)i + 3); break; }
    g->tab_dirty[k.member] = 1;
    return VSM_OK;
}

int vsm_group_store_clear(vsm_group* g) {
    if (!g) return VSM_ERR_INVALID;
    for (int r = 0; r < g->n; r++) {
        const int st = vsm_store_clear(g->ctx[r]);
        if (st != VSM_OK) { g->err = g->ctx[r]->err; return st; }
        g->h_tab[r].clear();
        g->tab_dirty[r] = 1;
    }
    g->kfs.clear();
    g->g_rows = 0;
    g->adopted = false;
    return VSM_OK;
}

int vsm_group_store_info(const vsm_group* g, int64_t* n_rows, int32_t* n_keyframes, int64_t* rows_per_member) {
    if (!g) return VSM_ERR_INVALID;
    int32_t nk = 0;
    for (auto& k : g->kfs) nk += k.live;
    int64_t rows = 0;
    for (int r = 0; r < g->n; r++) {
        int64_t live = 0;
        for (int32_t h : g->ctx[r]->kf_order) live += g->ctx[r]->segs[h].count;
        if (rows_per_member) rows_per_member[r] = live;
        rows += live;
    }
    if (n_rows) *n_rows = rows;
    if (n_keyframes) *n_keyframes = nk;
    return VSM_OK;
}

int vsm_group_adopt_device(vsm_group* g, int32_t member, const float* d_desc, int64_t n_rows, const int64_t* seg_off,
                           int32_t nseg) {
    if (!g || member < 0 || member >= g->n) return g ? group_fail(g, VSM_ERR_INVALID, "vsm_group_adopt_device: bad member") : VSM_ERR_INVALID;
    vsm_ctx* ctx = g->ctx[member];
    if (!ctx->kf_order.empty()) return group_fail(g, VSM_ERR_INVALID, "vsm_group_adopt_device: this member already holds keyframes; clear the group first");
    const int st = vsm_store_adopt_device(ctx, d_desc, n_rows, seg_off, nseg);
    if (st != VSM_OK) { g->err = ctx->err; return st; }
    g->adopted = true;
    for (int32_t h : ctx->kf_order) group_note_keyframe(g, member, h, ctx->segs[h].count, ctx->segs[h].frame_id, nullptr);
    return VSM_OK;
}

// table of a member: local row ranges -> stacked rows, sorted by local row0, on its device
static int group_upload_table(vsm_group* g, int r) {
    vsm_ctx* ctx = g->ctx[r];
    if (!g->tab_dirty[r]) return VSM_OK;
    std::vector<int64_t>& t = g->h_tab[r];
    const size_t ne = t.size() / 3;
    std::vector<size_t> ord(ne);
    for (size_t i = 0; i < ne; i++) ord[i] = i;
    std::sort(ord.begin(), ord.end(), [&](size_t a, size_t b) { return t[3 * a] < t[3 * b]; });
    std::vector<int64_t> s(std::max<size_t>(ne, 1) * 3, 0);
    for (size_t i = 0; i < ne; i++) for (int c = 0; c < 3; c++) s[3 * i + c] = t[3 * ord[i] + c];
    t.assign(s.begin(), s.begin() + (long)(ne * 3));
    TRY(ensure(ctx, g->d_tab[r], s.size()));
    CK(cudaMemcpyAsync(g->d_tab[r].p, s.data(), s.size() * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));                   // `s` is a local
    g->tab_n[r] = (int32_t)ne;
    g->tab_dirty[r] = 0;
    return VSM_OK;
}

int vsm_group_db_top2(vsm_group* g, const float* query, int32_t nq, int64_t* idx, float* dist, int32_t* kf_handle,
                      int32_t* kf_row) {
    if (!g || nq < 0 || (nq > 0 && (!query || !idx || !dist)))
        return g ? group_fail(g, VSM_ERR_INVALID, "vsm_group_db_top2: bad argument") : VSM_ERR_INVALID;
    if (nq == 0) return VSM_OK;
    g->err.clear();
    TRY(group_ensure_buffers(g, nq));
    const float* staged = nullptr;
    TRY(group_stage_query(g, query, nq, &staged));
    unsigned long long* gather = g->gather;
    TRY(group_fan_out(g, [g, staged, nq, gather](int r) -> int {
        vsm_ctx* ctx = g->ctx[r];
        TRY(begin_call(ctx));
        TRY(group_upload_table(g, r));
        TRY(arena_reserve(ctx, ctx->scratch, nq, 0));
        TRY(upload_scratch(ctx, staged, 0, nq));
        HProblem p;
        db_problem(ctx, ctx->scratch.f32, nq, p);
        TRY(run_problems(ctx, {p}, {}, nq, 0));
        // keys with STACKED row indices, stored straight into the first device's gather slot of this member
        stacked_keys_kernel<<<(nq * 2 + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_out_key, nq * 2, g->d_tab[r].p, g->tab_n[r],
                                                                          gather + (size_t)r * nq * 2);
        ctx->launches++;
        CK(cudaGetLastError());
        CK(cudaEventRecord(g->ev[r], ctx->stream));
        return end_call(ctx, false);
    }));
    // join on the first device: wait (in stream order) for every member's keys, merge, result in pinned host memory
    vsm_ctx* c0 = g->ctx[0];
    GCK(cudaSetDevice(g->dev[0]));
    for (int r = 1; r < g->n; r++) GCK(cudaStreamWaitEvent(c0->stream, g->ev[r], 0));
    int64_t* h_idx = reinterpret_cast<int64_t*>(g->h_out);
    float* h_dist = reinterpret_cast<float*>(g->h_out + (size_t)nq * 2 * sizeof(int64_t));
    merge_keys_kernel<<<(nq + 127) / 128, 128, 0, c0->stream>>>(gather, g->n, nq, h_idx, h_dist);
    GCK(cudaGetLastError());
    GCK(cudaStreamSynchronize(c0->stream));
    // (per-member statistics stay pending: vsm_get_stats on a member context collects them on demand)
    memcpy(idx, h_idx, (size_t)nq * 2 * sizeof(int64_t));
    memcpy(dist, h_dist, (size_t)nq * 2 * sizeof(float));
    if (kf_handle || kf_row) {
        // stacked row -> (group keyframe handle, row inside the keyframe): grow0 ascends with the handle
        for (int i = 0; i < nq * 2; i++) {
            int32_t kh = -1, kr = -1;
            if (idx[i] >= 0) {
                size_t lo = 0, hi = g->kfs.size();
                while (hi - lo > 1) { const size_t mid = (lo + hi) / 2; if (g->kfs[mid].grow0 <= idx[i]) lo = mid; else hi = mid; }
                while (lo + 1 < g->kfs.size() && g->kfs[lo].count == 0) lo++;             // empty keyframes share a grow0
                kh = (int32_t)lo;
                kr = (int32_t)(idx[i] - g->kfs[lo].grow0);
            }
            if (kf_handle) kf_handle[i] = kh;
            if (kf_row) kf_row[i] = kr;
        }
    }
    return VSM_OK;
}

// LoopCloser::detect's candidate loop (src/LoopCloser.cpp:43-62) over the group's keyframe list: the
// eligibility rules run over the WHOLE list in insertion order (the every-5th counter is global), each
// member matches its own eligible keyframes, no data crosses devices.
int vsm_group_loop_detect(vsm_group* g, int32_t cur_frame_id, int32_t min_gap, int32_t every, const float* query,
                          int32_t nq, float ratio, int32_t* status, vsm_dmatch* matches) {
    if (!g || nq < 0 || (nq > 0 && !query) || !status || every <= 0)
        return g ? group_fail(g, VSM_ERR_INVALID, "vsm_group_loop_detect: bad argument") : VSM_ERR_INVALID;
    g->err.clear();
    // live keyframes in insertion order; position in this list indexes status / matches
    std::vector<int32_t> live;
    for (int32_t h = 0; h < (int32_t)g->kfs.size(); h++) if (g->kfs[h].live) live.push_back(h);
    const int nkf = (int)live.size();
    std::vector<std::vector<char>> elig(g->n);
    std::vector<std::vector<int32_t>> cnt(g->n);
    std::vector<std::vector<int32_t>> pos_of(g->n);           // member-local keyframe position -> position in `live`
    std::vector<std::vector<int32_t>> local_pos(g->n);        // member: local handle -> position in its kf_order
    for (int r = 0; r < g->n; r++) {
        vsm_ctx* ctx = g->ctx[r];
        const int ln = (int)ctx->kf_order.size();
        elig[r].assign(ln, 0);
        cnt[r].assign(std::max(ln, 1), 0);
        pos_of[r].assign(ln, -1);
        local_pos[r].assign(ctx->segs.size(), -1);
        for (int k = 0; k < ln; k++) local_pos[r][ctx->kf_order[k]] = k;
    }
    int checked = 0;
    bool any = false;
    for (int s = 0; s < nkf; s++) {                                       // src/LoopCloser.cpp:43-48
        const GKeyframe& k = g->kfs[live[s]];
        status[s] = -1;
        const int lp = local_pos[k.member][k.local_handle];
        pos_of[k.member][lp] = s;
        if (cur_frame_id - g->ctx[k.member]->segs[k.local_handle].frame_id < min_gap) continue;
        if (k.count == 0) continue;
        checked++;
        if (checked % every != 0) continue;
        elig[k.member][lp] = 1;
        status[s] = 0;
        any = true;
    }
    if (nq == 0 || !any) return VSM_OK;
    const float* staged = nullptr;
    TRY(group_stage_query(g, query, nq, &staged));
    std::vector<std::vector<vsm_dmatch>> slab(g->n);
    if (matches) for (int r = 0; r < g->n; r++) slab[r].resize((size_t)std::max<size_t>(g->ctx[r]->kf_order.size(), 1) * nq);
    TRY(group_fan_out(g, [&](int r) -> int {
        vsm_ctx* ctx = g->ctx[r];
        bool mine = false;
        for (char e : elig[r]) mine |= e != 0;
        if (!mine) return VSM_OK;
        return segmented_impl(ctx, staged, nq, ratio, &elig[r], cnt[r].data(), matches ? slab[r].data() : nullptr);
    }));
    for (int r = 0; r < g->n; r++)
        for (size_t k = 0; k < elig[r].size(); k++) {
            if (!elig[r][k]) continue;
            const int s = pos_of[r][k];
            status[s] = cnt[r][k];
            if (matches) {
                memcpy(matches + (size_t)s * nq, slab[r].data() + k * (size_t)nq, (size_t)cnt[r][k] * sizeof(vsm_dmatch));
                for (int i = 0; i < cnt[r][k]; i++) matches[(size_t)s * nq + i].imgIdx = s;
            }
        }
    return VSM_OK;
}

// The same loop in compact form (vsm_loop_detect_compact): every member gates and packs the lists of
// its own keyframes on its device; the host orders the surviving keyframes by list position.
int vsm_group_loop_detect_compact(vsm_group* g, int32_t cur_frame_id, int32_t min_gap, int32_t every, const float* query,
                                  int32_t nq, float ratio, int32_t min_matches, int32_t* status, vsm_loop_candidate* cands,
                                  int32_t cand_cap, int32_t* n_cands, vsm_dmatch* matches, int64_t match_cap,
                                  int64_t* n_matches) {
    if (!g || nq < 0 || (nq > 0 && !query) || !status || every <= 0 || cand_cap < 0 || match_cap < 0 || !n_cands || !n_matches ||
        (cand_cap > 0 && !cands) || (match_cap > 0 && !matches))
        return g ? group_fail(g, VSM_ERR_INVALID, "vsm_group_loop_detect_compact: bad argument") : VSM_ERR_INVALID;
    g->err.clear();
    *n_cands = 0;
    *n_matches = 0;
    std::vector<int32_t> live;
    for (int32_t h = 0; h < (int32_t)g->kfs.size(); h++) if (g->kfs[h].live) live.push_back(h);
    const int nkf = (int)live.size();
    std::vector<std::vector<char>> elig(g->n);
    std::vector<std::vector<int32_t>> st(g->n), pos_of(g->n), local_pos(g->n);
    for (int r = 0; r < g->n; r++) {
        vsm_ctx* ctx = g->ctx[r];
        const int ln = (int)ctx->kf_order.size();
        elig[r].assign(ln, 0);
        st[r].assign(std::max(ln, 1), -1);
        pos_of[r].assign(ln, -1);
        local_pos[r].assign(ctx->segs.size(), -1);
        for (int k = 0; k < ln; k++) local_pos[r][ctx->kf_order[k]] = k;
    }
    int checked = 0;
    bool any = false;
    for (int s = 0; s < nkf; s++) {                                       // src/LoopCloser.cpp:43-48
        const GKeyframe& k = g->kfs[live[s]];
        status[s] = -1;
        const int lp = local_pos[k.member][k.local_handle];
        pos_of[k.member][lp] = s;
        if (cur_frame_id - g->ctx[k.member]->segs[k.local_handle].frame_id < min_gap) continue;
        if (k.count == 0) continue;
        checked++;
        if (checked % every != 0) continue;
        elig[k.member][lp] = 1;
        status[s] = 0;
        any = true;
    }
    if (nq == 0 || !any) return VSM_OK;
    const float* staged = nullptr;
    TRY(group_stage_query(g, query, nq, &staged));
    struct Part { std::vector<vsm_loop_candidate> c; std::vector<vsm_dmatch> m; int32_t nc = 0; int64_t nm = 0; };
    std::vector<Part> part(g->n);
    TRY(group_fan_out(g, [&](int r) -> int {
        vsm_ctx* ctx = g->ctx[r];
        bool mine = false;
        for (char e : elig[r]) mine |= e != 0;
        if (!mine) return VSM_OK;
        Part& p = part[r];
        p.c.resize(64);
        p.m.resize((size_t)64 * std::max(nq, 1));
        for (;;) {                                                        // grow the member's buffers if a search returns more
            TRY(loop_compact_eligible(ctx, elig[r], staged, nq, ratio, min_matches, st[r].data(), p.c.data(), (int32_t)p.c.size(),
                                      &p.nc, p.m.data(), (int64_t)p.m.size(), &p.nm));
            if (p.nc <= (int32_t)p.c.size() && p.nm <= (int64_t)p.m.size()) return VSM_OK;
            p.c.resize(std::max<size_t>(p.c.size(), (size_t)p.nc));
            p.m.resize(std::max<size_t>(p.m.size(), (size_t)p.nm));
        }
    }));
    // surviving keyframes of all members in list order
    struct Ref { int32_t pos, member, k; };
    std::vector<Ref> refs;
    for (int r = 0; r < g->n; r++) {
        for (size_t k = 0; k < elig[r].size(); k++) if (elig[r][k]) status[pos_of[r][k]] = st[r][k];
        for (int k = 0; k < part[r].nc; k++) refs.push_back(Ref{pos_of[r][part[r].c[k].keyframe], r, k});
    }
    std::sort(refs.begin(), refs.end(), [](const Ref& a, const Ref& b) { return a.pos < b.pos; });
    int64_t nm = 0;
    int nc = 0;
    for (const Ref& f : refs) {
        const vsm_loop_candidate& c = part[f.member].c[f.k];
        if (nc < cand_cap) cands[nc] = vsm_loop_candidate{f.pos, c.count, nm};
        for (int i = 0; i < c.count; i++)
            if (nm + i < match_cap) {
                matches[nm + i] = part[f.member].m[(size_t)c.offset + i];
                matches[nm + i].imgIdx = f.pos;
            }
        nc++;
        nm += c.count;
    }
    *n_cands = nc;
    *n_matches = nm;
    return VSM_OK;
}

// Ragged batch of independent pairs (vsm_match_batch, BASELINE configs[4]) over the group: pair matching does not
// shard (one pair is ~1 us of tensor time) but independent pairs are replicas -- the batch is cut into
// contiguous blocks of pairs with near-equal input bytes, one block per member, so the upload (what bounds
// the call: 150 MB for 64 pairs of up to 2048 keypoints) runs over every member's own PCIe link at once.
int vsm_group_match_batch(vsm_group* g, int32_t n_pairs, const float* query, const int32_t* q_off, const float* train,
                          const int32_t* t_off, float ratio, int32_t mutual, vsm_dmatch* good, int32_t* n_good) {
    if (!g || n_pairs < 0 || (n_pairs > 0 && (!q_off || !t_off || !n_good)))
        return g ? group_fail(g, VSM_ERR_INVALID, "vsm_group_match_batch: bad argument") : VSM_ERR_INVALID;
    g->err.clear();
    if (n_pairs == 0) return VSM_OK;
    for (int p = 0; p < n_pairs; p++) {
        if (q_off[p + 1] < q_off[p] || t_off[p + 1] < t_off[p]) return group_fail(g, VSM_ERR_INVALID, "vsm_group_match_batch: offsets must be non-decreasing");
        n_good[p] = 0;
    }
    const int64_t total = (int64_t)q_off[n_pairs] + t_off[n_pairs];
    if (q_off[n_pairs] == 0) return VSM_OK;
    if (!query || !good || (t_off[n_pairs] > 0 && !train)) return group_fail(g, VSM_ERR_INVALID, "vsm_group_match_batch: null buffer");
    // cut points: member r takes pairs [cut[r], cut[r+1]) -- the first pair whose cumulative rows reach r/n of the total
    std::vector<int> cut(g->n + 1, n_pairs);
    cut[0] = 0;
    for (int r = 1, p = 0; r < g->n; r++) {
        const int64_t target = total * r / g->n;
        while (p < n_pairs && (int64_t)q_off[p] + t_off[p] < target) p++;
        cut[r] = p;
    }
    return group_fan_out(g, [&](int r) -> int {
        const int p0 = cut[r], p1 = cut[r + 1];
        if (p1 <= p0) return VSM_OK;
        std::vector<int32_t> qo(p1 - p0 + 1), to(p1 - p0 + 1);
        for (int p = p0; p <= p1; p++) { qo[p - p0] = q_off[p] - q_off[p0]; to[p - p0] = t_off[p] - t_off[p0]; }
        return vsm_match_batch(g->ctx[r], p1 - p0, query + (size_t)q_off[p0] * VSM_DIM, qo.data(),
                               train ? train + (size_t)t_off[p0] * VSM_DIM : nullptr, to.data(), ratio, mutual,
                               good + q_off[p0], n_good + p0);
    });
}
