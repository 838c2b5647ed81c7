// vsm_api.cu -- host side of libvsm.so: context, device-resident descriptor arenas,
// launch planning and the C ABI declared in include/vsm.h.
//
// There is no CPU fallback anywhere in this file: every entry point either runs the
// CUDA kernels or returns an error code.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <functional>
#include <mutex>
#include <thread>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/vsm.h"
#include "vsm_common.cuh"
#include "vsm_kernels.cuh"
#include "vsm_tc.cuh"
#include "vsm_tc2.cuh"

using namespace vsm;

namespace {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// virtual memory management (driver API, resolved through cudaGetDriverEntryPoint like the tensor-map encoder)
typedef CUresult (*PFN_cuMemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long);
typedef CUresult (*PFN_cuMemAddressFree)(CUdeviceptr, size_t);
typedef CUresult (*PFN_cuMemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long);
typedef CUresult (*PFN_cuMemRelease)(CUmemGenericAllocationHandle);
typedef CUresult (*PFN_cuMemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
typedef CUresult (*PFN_cuMemUnmap)(CUdeviceptr, size_t);
typedef CUresult (*PFN_cuMemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t);
typedef CUresult (*PFN_cuMemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags);

thread_local std::string g_create_error;

// One array of an arena that grows IN PLACE: a virtual address range reserved once
// (cuMemAddressReserve) into which physical chunks are mapped as the row count grows
// (cuMemCreate + cuMemMap + cuMemSetAccess).  Growth never moves a row and never copies.
struct VmmRange {
    CUdeviceptr base = 0;
    size_t reserved = 0, mapped = 0;
    std::vector<std::pair<CUmemGenericAllocationHandle, size_t>> chunks;    // (handle, bytes), in address order
};

struct Arena {
    float* f32 = nullptr;            // fp32 master rows (exact re-score reads these)
    __nv_bfloat16* b16 = nullptr;    // bf16 shadow (tensor-core operand, TMA source)
    float* n2 = nullptr;             // squared norms
    int64_t cap = 0;                 // rows
    bool own_f32 = true;
    bool vmm = false;                // arrays live in VmmRanges (growth maps chunks, no copy)
    VmmRange r_f32, r_b16, r_n2;
    CUtensorMap map;
};

// One stored frame = a contiguous row range of the store (Frame::descriptors_, include/Frame.h:61).
// Handles index `segs`; the slot of a removed frame is reused.
struct Seg {
    int64_t row0;
    int32_t count;
    int32_t frame_id;
    int64_t seq;                     // insertion order (Map::frames_ order, src/Map.cpp:7-10)
    uint8_t live;                    // 0 = removed (slot waits in free_handles)
    uint8_t is_kf;                   // Frame::is_keyframe() (Map::get_keyframes filters on it, src/Map.cpp:40-47)
};

// A device array that grows in place (virtual memory management; cudaMalloc + copy without it).
struct GrowArr {
    VmmRange r;
    uint8_t* p = nullptr;
    size_t cap = 0;                  // bytes usable
    bool vmm = false;
};

// A run of consecutive store rows that belong to live keyframes (global search skips the rest).
struct Run {
    int64_t row0, count;
};

// Host description of one kNN problem before planning.
struct HProblem {
    const float* q_f32;  const float* q_n2;  int64_t q_row;  int q_store;  int nq;
    const float* t_f32;  int64_t t_row;  int t_store;  int nt;
    int64_t out_off;
    float skip_ratio2 = 0.f;     // see Problem::skip_ratio2
    int32_t maxima_only = 0;     // with skip_ratio2: records = slice maxima only (matches are rare: LoopCloser's loop)
    // pair matching on small train sets: tile top-2 records (Problem::exact bit 3).  1 = forward problem of a
    // ratio-only caller (needs skip_ratio2 > 0 and 0 < ratio <= 1), 2 = reverse problem of a mutual test
    // (only the nearest neighbour's index is read)
    int32_t t2 = 0;
    float ratio = 0.f;
    // train rows = these runs of STORE rows (logical train index = store row; t_row = 0, nt = the store's
    // high-water mark); nullptr = the one contiguous range [t_row, t_row + nt)
    const std::vector<Run>* runs = nullptr;
};

// ratio^2 with slack for a caller that only keeps ratio-test survivors; 0 = never skip
inline float skip_r2(float ratio) { return (ratio > 0.f && ratio <= 1.5f) ? ratio * ratio * 1.001f : 0.f; }

struct HJob {
    int64_t fwd_off, back_off, good_off, raw_off;
    int nq, nt, img_idx;
    float ratio;
    int64_t back_prob = -1;      // index of the reverse problem in the call's list (mutual test), -1 = none
};

static_assert(sizeof(HProblem) == 88 && sizeof(HJob) == 56, "plan keys compare these byte for byte: no padding allowed");

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
};

}  // namespace

struct vsm_ctx {
    int device = 0;
    int engine = VSM_ENGINE_AUTO;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    Arena scratch, store;
    int64_t store_rows = 0;              // high-water mark: rows [0, store_rows) have been handed out
    std::vector<Seg> segs;               // indexed by handle
    std::vector<int32_t> free_handles;   // slots of removed frames
    std::vector<Run> free_rows;          // row ranges of removed frames, sorted by row0, coalesced
    std::vector<int32_t> kf_order;       // live keyframes in insertion order = Map::get_keyframes() (src/Map.cpp:40-47)
    std::vector<int32_t> plain_ring;     // live NON-keyframe frames, oldest first (last_frame_ and the current one)
    int ring_depth = 2;                  // plain frames kept: a new one evicts the oldest (src/Slam.cpp:838 needs last_frame_)
    int64_t next_seq = 0;
    std::vector<Run> kf_runs;            // merged row runs of the live keyframes (global search), see store_runs()
    bool kf_runs_valid = false;
    bool vmm_ok = false;                 // driver entry points below resolved and the device supports them
    size_t vmm_gran = 0;
    PFN_cuMemAddressReserve p_reserve = nullptr;
    PFN_cuMemAddressFree p_addr_free = nullptr;
    PFN_cuMemCreate p_create = nullptr;
    PFN_cuMemRelease p_release = nullptr;
    PFN_cuMemMap p_map = nullptr;
    PFN_cuMemUnmap p_unmap = nullptr;
    PFN_cuMemSetAccess p_set_access = nullptr;
    PFN_cuMemGetAllocationGranularity p_gran = nullptr;
    int64_t arena_grow_copies = 0;       // growth events that had to copy (non-VMM fall-back path); tests read it
    uint32_t* h_status = nullptr;        // pinned, device-visible: kernels report soft failures here (exchange timeout)
    uint32_t* d_store_stats = nullptr;   // {min, max} squared norm over the store
    DevBuf<uint8_t> d_desc;
    uint8_t* h_desc = nullptr;
    size_t h_desc_cap = 0;
    std::vector<uint8_t> desc_build;                 // descriptor block under construction
    std::vector<uint8_t> plan_key_build;
    cudaEvent_t ev_desc = nullptr;
    bool desc_copy_pending = false;
    DevBuf<PartialRec> d_recs;
    unsigned long long* d_out_key = nullptr;     // [query][2] result keys, inside d_aux (zeroed per call)
    DevBuf<uint8_t> d_result;            // DMatch lists followed by the counts
    bool result_on_host = false;         // small results: filter_kernel writes straight into pinned h_result
    uint8_t* h_result = nullptr;
    size_t h_result_cap = 0;
    // zeroed per call: [0] candidates, [1] flagged slices (u64), [2] rescan work count (u32) and the
    // number of ratio-only queries that were not dismissed (u32), unit queue head, statistics slots, then
    // per output query the shared second-best hint (u32) and the two result keys (u64)
    DevBuf<uint8_t> d_aux;
    unsigned long long* d_counters = nullptr;    // = d_aux.p
    DevBuf<WorkItem> d_work;
    DevBuf<int32_t> d_sel;                       // selected store rows of a masked search
    DevBuf<uint8_t> d_track;                     // track_local_map: keypoints, map-point positions, results
    // resident map-point table (MapPoint::descriptor_ / valid_ / observations_, include/MapPoint.h:38-42)
    GrowArr pt_f32, pt_valid, pt_log;            // descriptors [n][256] fp32, validity [n] u8, observation log [(point, frame)]
    int64_t n_points = 0, n_log = 0, n_valid = 0;
    std::vector<uint8_t> pt_valid_h;             // host mirror of the validity flags (set_valid is idempotent, the count is not)
    DevBuf<uint8_t> d_pt_tmp;                    // per-search: near flags, selection flags, block counts / offsets, total
    DevBuf<uint8_t> d_loop;                      // compact loop search: masks, word bases, pair keys, staged matches, offsets
    std::vector<uint8_t> loop_key;               // plan key of the last compact loop search (descriptor block reuse)
    const void* loop_p_desc = nullptr;
    const void* loop_p_loop = nullptr;
    uint32_t pair_cap = 0;                       // open (query, keyframe) pairs the compact loop search can hold (0 = PAIR_CAP)
    int64_t append_max_tiles = 0;                // train sets of up to this many tiles use append records (0 = never)
    bool t2_off = false;                         // vsm_opts.reserved[5] = 1 / VSM_NO_T2: pair matching keeps the top-4 records
    float* d_dump = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // event pairs around the tensor-core kernel of the last TC_RING calls (ev_tc0/ev_tc1 = the current
    // call's pair): asynchronous callers read a whole timed loop afterwards (vsm_tc_history)
    static constexpr int TC_RING = 64;
    cudaEvent_t tc_ring0[TC_RING] = {}, tc_ring1[TC_RING] = {};
    bool tc_ring_valid[TC_RING] = {};
    uint32_t tc_ring_head = 0;               // slot of the next timed call
    cudaEvent_t ev_tc0 = nullptr, ev_tc1 = nullptr, ev_sel1 = nullptr;
    bool timed_tc = false, timed_sel = false, timed_call = false, pending_stats = false;
    vsm_stats stats{};
    int launches = 0;
    std::string err;
    PFN_encodeTiled encode = nullptr;
    // fused exchange over peer memory
    uint8_t* xchg_buf = nullptr;
    void* xchg_peer_ptr[XCHG_MAX_WORLD] = {};
    XchgPeers xchg_peers{};
    int xchg_rank = 0, xchg_world = 0, xchg_nq_cap = 0;
    uint32_t xchg_step = 0;
    bool xchg_connected = false;
    int seg_tiles = 0;                   // 0 = automatic
    bool profiling = true;               // per-kernel events (tc_ms / select_ms)
    uint32_t work_cap = 0;               // rescan work-list capacity (0 = WORK_CAP)
    // plan of the last run_problems call: a call with identical inputs (the loop-closure search and
    // the tracking step repeat their shapes) reuses the descriptor block already on the device
    struct Plan {
        std::vector<uint8_t> key;
        const void* p_desc = nullptr;
        const void* p_aux = nullptr;
        const void* p_stats = nullptr;
        size_t off_prob = 0, off_unit = 0, off_slice = 0, off_job = 0, total = 0, nunits = 0, nunits2 = 0;
        int64_t nrecs = 0;
        int qb_total = 0;                            // 32-query groups of tile top-2 problems
        bool any_classic = false;                    // problems answered by select_kernel
        int t2_gshift = 5;                           // log2 of the queries per warp in t2_select_kernel
        int32_t big_db_nq = 0, slice_tiles = 0;      // what the planning loop noted for the slice-length feedback
        bool valid = false;
    } plan;
    int64_t plan_hits = 0;
    std::vector<ConvJob> pending_conv;   // conversions queued by this call, run by its prologue kernel
    uint32_t call_seq = 0;               // run_problems calls so far: picks the scratch-statistics slot
    const void* aux_zeroed = nullptr;    // d_aux.p at the time its slots were last known to be zero
    // share of the (query, keyframe) problems of the last per-keyframe search that the ratio-only test
    // could NOT dismiss; starts pessimistic.  Decides the epilogue of the next one (segmented_impl).
    float seg_open_rate = 1.f;
    // Slice length of a big database search (more than 4096 tiles), chosen from the data: 64-tile slices keep
    // the record stream and select_kernel's walk over it small, but a slice whose four recorded entries
    // overflow is re-scanned exactly, 8192 rows per (query, slice).  On clustered databases (keyframes of
    // one place hold near-copies of the same descriptors, stored next to each other) almost every query
    // overflows the slice of its own cluster; 16-tile slices make that re-scan four times cheaper for
    // ~1 % more time on friendly data.  The previous big search's overflow count decides (hysteresis).
    bool db_short_slices = false;
    int32_t big_db_nq = 0;               // queries of the big database search of the current call (0 = none)
    int32_t big_db_feedback_nq = 0;      // ... of the previous call, whose overflow count waits in h_status[6..7]
    int32_t last_slice_tiles = 0;        // slice length the last planned problem used (vsm_stats)
};

namespace {

#define CK(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            char b_[512];                                                                            \
            snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            ctx->err = b_;                                                                           \
            return VSM_ERR_CUDA;                                                                     \
        }                                                                                            \
    } while (0)

#define TRY(call)                    \
    do {                             \
        int s_ = (call);             \
        if (s_ != VSM_OK) return s_; \
    } while (0)

int fail(vsm_ctx* ctx, int code, const char* msg) {
    ctx->err = msg;
    return code;
}

template <class T>
int ensure(vsm_ctx* ctx, DevBuf<T>& b, size_t n) {
    if (n <= b.cap) return VSM_OK;
    size_t want = std::max(n, b.cap + b.cap / 2);
    CK(cudaStreamSynchronize(ctx->stream));
    if (b.p) CK(cudaFree(b.p));
    b.p = nullptr; b.cap = 0;
    CK(cudaMalloc(&b.p, want * sizeof(T)));
    b.cap = want;
    return VSM_OK;
}

int ensure_host(vsm_ctx* ctx, uint8_t*& p, size_t& cap, size_t n) {
    if (n <= cap) return VSM_OK;
    size_t want = std::max(n, cap * 2);
    CK(cudaStreamSynchronize(ctx->stream));
    if (p) CK(cudaFreeHost(p));
    p = nullptr; cap = 0;
    CK(cudaMallocHost(&p, want));
    cap = want;
    return VSM_OK;
}

int encode_map(vsm_ctx* ctx, Arena& a) {
    // bf16 [cap rows][256], box = 64 columns x 128 rows, SWIZZLE_128B (rows of 128 bytes)
    cuuint64_t gdim[2] = {(cuuint64_t)VSM_DIM, (cuuint64_t)a.cap};
    cuuint64_t gstr[1] = {(cuuint64_t)VSM_DIM * sizeof(__nv_bfloat16)};
    cuuint32_t box[2] = {64u, 128u};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = ctx->encode(&a.map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a.b16, gdim, gstr, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char b[128];
        snprintf(b, sizeof b, "cuTensorMapEncodeTiled failed: CUresult %d", (int)r);
        ctx->err = b;
        return VSM_ERR_CUDA;
    }
    return VSM_OK;
}

#define CKU(call)                                                                                    \
    do {                                                                                             \
        CUresult r_ = (call);                                                                        \
        if (r_ != CUDA_SUCCESS) {                                                                    \
            char b_[512];                                                                            \
            snprintf(b_, sizeof b_, "%s failed: CUresult %d (%s:%d)", #call, (int)r_, __FILE__, __LINE__); \
            ctx->err = b_;                                                                           \
            return r_ == CUDA_ERROR_OUT_OF_MEMORY ? VSM_ERR_CAPACITY : VSM_ERR_CUDA;                 \
        }                                                                                            \
    } while (0)

// ---- arenas that grow in place (virtual memory management) ------------------------------------
int vmm_reserve(vsm_ctx* ctx, VmmRange& r, size_t bytes) {
    bytes = (bytes + ctx->vmm_gran - 1) / ctx->vmm_gran * ctx->vmm_gran;
    CKU(ctx->p_reserve(&r.base, bytes, 0, 0, 0));
    r.reserved = bytes;
    r.mapped = 0;
    return VSM_OK;
}

// Map physical memory so that [base, base + need) is backed; existing mappings are untouched.
int vmm_back(vsm_ctx* ctx, VmmRange& r, size_t need) {
    need = (need + ctx->vmm_gran - 1) / ctx->vmm_gran * ctx->vmm_gran;
    if (need <= r.mapped) return VSM_OK;
    if (need > r.reserved) return fail(ctx, VSM_ERR_CAPACITY, "descriptor arena: beyond the reserved address range");
    const size_t bytes = need - r.mapped;
    CUmemAllocationProp prop;
    memset(&prop, 0, sizeof prop);
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = ctx->device;
    CUmemGenericAllocationHandle h;
    CKU(ctx->p_create(&h, bytes, &prop, 0));
    CUresult rc = ctx->p_map(r.base + r.mapped, bytes, 0, h, 0);
    if (rc == CUDA_SUCCESS) {
        CUmemAccessDesc acc;
        memset(&acc, 0, sizeof acc);
        acc.location = prop.location;
        acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        rc = ctx->p_set_access(r.base + r.mapped, bytes, &acc, 1);
        if (rc != CUDA_SUCCESS) ctx->p_unmap(r.base + r.mapped, bytes);
    }
    if (rc != CUDA_SUCCESS) {
        ctx->p_release(h);
        char b[160];
        snprintf(b, sizeof b, "descriptor arena: mapping %zu bytes failed (CUresult %d)", bytes, (int)rc);
        return fail(ctx, rc == CUDA_ERROR_OUT_OF_MEMORY ? VSM_ERR_CAPACITY : VSM_ERR_CUDA, b);
    }
    r.chunks.push_back({h, bytes});
    r.mapped = need;
    return VSM_OK;
}

void vmm_release(vsm_ctx* ctx, VmmRange& r) {
    size_t off = 0;
    for (auto& c : r.chunks) {
        ctx->p_unmap(r.base + off, c.second);
        ctx->p_release(c.first);
        off += c.second;
    }
    if (r.base) ctx->p_addr_free(r.base, r.reserved);
    r = VmmRange();
}

// Grow an arena to hold `rows`; the first `keep` rows survive.  With virtual memory management
// (the normal case) growth maps more chunks behind the same addresses: nothing is copied, nothing
// moves, no stream is synchronised.  The cudaMalloc path below is only taken when the driver has no
// VMM support or for an adopted matrix (whose fp32 master is the caller's).
int arena_reserve(vsm_ctx* ctx, Arena& a, int64_t rows, int64_t keep) {
    if (rows <= a.cap && a.b16) return VSM_OK;
    if (a.vmm || (ctx->vmm_ok && a.own_f32 && !a.b16)) {
        if (!a.vmm) {
            // address space for as many rows as the device could ever hold (fp32 + bf16 + norm = 1540 B per row)
            size_t free_b = 0, total_b = 0;
            CK(cudaMemGetInfo(&free_b, &total_b));
            const size_t max_rows = total_b / 1536 + (1u << 20);
            int st = vmm_reserve(ctx, a.r_f32, max_rows * VSM_DIM * sizeof(float));
            if (st == VSM_OK) st = vmm_reserve(ctx, a.r_b16, max_rows * VSM_DIM * sizeof(__nv_bfloat16));
            if (st == VSM_OK) st = vmm_reserve(ctx, a.r_n2, max_rows * sizeof(float));
            if (st != VSM_OK) {
                // no address space to reserve (a restricted environment): this context allocates and copies instead
                vmm_release(ctx, a.r_f32); vmm_release(ctx, a.r_b16); vmm_release(ctx, a.r_n2);
                ctx->vmm_ok = false;
                ctx->err.clear();
                return arena_reserve(ctx, a, rows, keep);
            }
            a.vmm = true;
            a.f32 = reinterpret_cast<float*>(a.r_f32.base);
            a.b16 = reinterpret_cast<__nv_bfloat16*>(a.r_b16.base);
            a.n2 = reinterpret_cast<float*>(a.r_n2.base);
            a.cap = 0;
        }
        auto back = [&](int64_t want) -> int {
            TRY(vmm_back(ctx, a.r_f32, (size_t)want * VSM_DIM * sizeof(float)));
            TRY(vmm_back(ctx, a.r_b16, (size_t)want * VSM_DIM * sizeof(__nv_bfloat16)));
            TRY(vmm_back(ctx, a.r_n2, (size_t)want * sizeof(float)));
            return VSM_OK;
        };
        // head-room: a quarter of the arena, at least 64K rows (96 MB), so that growth is rare
        int64_t want = std::max<int64_t>(rows, a.cap + std::max<int64_t>(a.cap / 4, 65536));
        want = (want + 255) / 256 * 256;
        int st = back(want);
        if (st != VSM_OK) {                                  // the head-room did not fit: exactly what is needed
            want = (rows + 255) / 256 * 256;
            st = back(want);                                 // chunks mapped by the failed attempt are kept and reused
        }
        if (st != VSM_OK) return st;
        a.cap = want;
        return encode_map(ctx, a);
    }
    int64_t want = std::max<int64_t>(rows, a.cap + a.cap / 2);
    want = (want + 255) / 256 * 256;
    CK(cudaStreamSynchronize(ctx->stream));
    float* f32 = nullptr; __nv_bfloat16* b16 = nullptr; float* n2 = nullptr;
    // allocate and copy everything first: on failure (e.g. out of HBM) the old arena stays intact
    auto grow = [&]() -> int {
        if (a.own_f32) CK(cudaMalloc(&f32, (size_t)want * VSM_DIM * sizeof(float)));
        CK(cudaMalloc(&b16, (size_t)want * VSM_DIM * sizeof(__nv_bfloat16)));
        CK(cudaMalloc(&n2, (size_t)want * sizeof(float)));
        if (keep > 0) {
            if (a.own_f32) CK(cudaMemcpy(f32, a.f32, (size_t)keep * VSM_DIM * sizeof(float), cudaMemcpyDeviceToDevice));
            CK(cudaMemcpy(b16, a.b16, (size_t)keep * VSM_DIM * sizeof(__nv_bfloat16), cudaMemcpyDeviceToDevice));
            CK(cudaMemcpy(n2, a.n2, (size_t)keep * sizeof(float), cudaMemcpyDeviceToDevice));
            ctx->arena_grow_copies++;
        }
        return VSM_OK;
    };
    int st = grow();
    const int64_t tight = (rows + 255) / 256 * 256;
    if (st != VSM_OK && want > tight) {                     // the 1.5x head-room did not fit: take exactly what is needed
        cudaFree(f32); cudaFree(b16); cudaFree(n2);
        cudaGetLastError();
        f32 = nullptr; b16 = nullptr; n2 = nullptr;
        want = tight;
        st = grow();
    }
    if (st != VSM_OK) {
        cudaFree(f32); cudaFree(b16); cudaFree(n2);
        cudaGetLastError();
        ctx->err = "descriptor arena: out of device memory (" + ctx->err + ")";
        return VSM_ERR_CAPACITY;
    }
    if (a.own_f32) { if (a.f32) CK(cudaFree(a.f32)); a.f32 = f32; }
    if (a.b16) CK(cudaFree(a.b16));
    if (a.n2) CK(cudaFree(a.n2));
    a.b16 = b16; a.n2 = n2; a.cap = want;
    return encode_map(ctx, a);
}

// Make [p, p + need) usable; existing content keeps its address (VMM) or is copied (fall-back).
int grow_arr(vsm_ctx* ctx, GrowArr& a, size_t need, size_t reserve_bytes) {
    if (need <= a.cap) return VSM_OK;
    if (a.vmm || (ctx->vmm_ok && !a.p)) {
        if (!a.vmm) {
            if (vmm_reserve(ctx, a.r, reserve_bytes) != VSM_OK) {          // see arena_reserve
                ctx->vmm_ok = false;
                ctx->err.clear();
                return grow_arr(ctx, a, need, reserve_bytes);
            }
            a.vmm = true;
            a.p = reinterpret_cast<uint8_t*>(a.r.base);
        }
        const size_t want = std::max(need, a.cap + a.cap / 2);
        int st = vmm_back(ctx, a.r, std::min(want, a.r.reserved));
        if (st != VSM_OK) st = vmm_back(ctx, a.r, need);
        if (st != VSM_OK) return st;
        a.cap = a.r.mapped;
        return VSM_OK;
    }
    const size_t want = std::max(need, a.cap * 2);
    uint8_t* np = nullptr;
    CK(cudaStreamSynchronize(ctx->stream));
    if (cudaMalloc(&np, want) != cudaSuccess) { cudaGetLastError(); return fail(ctx, VSM_ERR_CAPACITY, "map-point table: out of device memory"); }
    if (a.p) {
        CK(cudaMemcpy(np, a.p, a.cap, cudaMemcpyDeviceToDevice));
        CK(cudaFree(a.p));
    }
    a.p = np;
    a.cap = want;
    return VSM_OK;
}

void grow_arr_free(vsm_ctx* ctx, GrowArr& a) {
    if (a.vmm) vmm_release(ctx, a.r);
    else if (a.p) cudaFree(a.p);
    a = GrowArr();
}

void arena_free(vsm_ctx* ctx, Arena& a) {
    if (a.vmm) {
        vmm_release(ctx, a.r_f32); vmm_release(ctx, a.r_b16); vmm_release(ctx, a.r_n2);
    } else {
        if (a.own_f32 && a.f32) cudaFree(a.f32);
        if (a.b16) cudaFree(a.b16);
        if (a.n2) cudaFree(a.n2);
    }
    a = Arena();
}

int launch_convert(vsm_ctx* ctx, const float* src, __nv_bfloat16* dst, float* n2, int64_t rows, uint32_t* stats) {
    if (rows <= 0) return VSM_OK;
    int64_t blocks = std::min<int64_t>((rows + 7) / 8, (int64_t)ctx->num_sms * 16);
    convert_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(src, dst, n2, rows, stats);
    ctx->launches++;
    CK(cudaGetLastError());
    return VSM_OK;
}

size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

constexpr int UNIT_TILES = 64;                   // longest train range of one work unit, in 256-row tiles
constexpr uint32_t WORK_CAP = 1u << 18;          // rescan work items (4 MB); beyond it select scans inline

int begin_call(vsm_ctx* ctx) {
    ctx->err.clear();
    ctx->launches = 0;
    ctx->timed_tc = ctx->timed_sel = false;
    ctx->big_db_nq = 0;
    if (ctx->big_db_feedback_nq) {
        // more than one overflowing slice per 8 queries in the previous big search: shorter slices from now
        // on; back to long ones when overflow is 8x rarer.  (An asynchronous caller that has not waited for
        // that search yet reads the count of the one before: late by one search, never wrong.)
        unsigned long long flagged;
        memcpy(&flagged, ctx->h_status + 6, sizeof flagged);
        if ((int64_t)flagged * 8 > ctx->big_db_feedback_nq) ctx->db_short_slices = true;
        else if ((int64_t)flagged * 64 < ctx->big_db_feedback_nq) ctx->db_short_slices = false;
        ctx->big_db_feedback_nq = 0;
    }
    ctx->pending_conv.clear();
    CK(cudaSetDevice(ctx->device));
    ctx->timed_call = ctx->profiling;
    if (ctx->timed_call) CK(cudaEventRecord(ctx->ev0, ctx->stream));
    return VSM_OK;
}

// Reads the event timings and counters of the last call (the stream must be idle).
int collect_stats(vsm_ctx* ctx) {
    if (!ctx->pending_stats) return VSM_OK;
    ctx->pending_stats = false;
    float ms = 0.f;
    if (ctx->timed_call) CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->stats.device_ms = ms;
    ctx->stats.tc_ms = ctx->stats.select_ms = 0.f;
    if (ctx->timed_tc) CK(cudaEventElapsedTime(&ctx->stats.tc_ms, ctx->ev_tc0, ctx->ev_tc1));
    if (ctx->timed_sel) CK(cudaEventElapsedTime(&ctx->stats.select_ms, ctx->timed_tc ? ctx->ev_tc1 : ctx->ev_tc0, ctx->ev_sel1));
    unsigned long long c[2] = {0, 0};
    if (ctx->d_counters && ctx->profiling) CK(cudaMemcpy(c, ctx->d_counters, sizeof c, cudaMemcpyDeviceToHost));
    ctx->stats.candidates = (int64_t)c[0];
    ctx->stats.flagged_slices = (int64_t)c[1];
    ctx->stats.slice_tiles = ctx->last_slice_tiles;
    return VSM_OK;
}

int end_call(vsm_ctx* ctx, bool sync) {
    if (ctx->big_db_nq && ctx->d_counters) {
        // the overflow count of a big database search travels to pinned host memory behind the call's own
        // work (no synchronisation here); the next call reads it and picks its slice length (begin_call)
        CK(cudaMemcpyAsync(ctx->h_status + 4, ctx->d_counters, 16, cudaMemcpyDeviceToHost, ctx->stream));
        ctx->big_db_feedback_nq = ctx->big_db_nq;
    }
    if (ctx->timed_call) CK(cudaEventRecord(ctx->ev1, ctx->stream));
    ctx->stats.kernel_launches = ctx->launches;
    ctx->pending_stats = true;
    if (sync) {
        CK(cudaStreamSynchronize(ctx->stream));
        return collect_stats(ctx);
    }
    return VSM_OK;
}

// Launch with programmatic stream serialization (see pdl_wait in vsm_common.cuh): the kernel may be
// scheduled while its predecessor in the stream still runs; it orders itself with pdl_wait().
template <class... KArgs, class... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    static const bool no_pdl = getenv("VSM_NO_PDL") != nullptr;         // A/B switch: plain stream order
    cfg.attrs = at;
    cfg.numAttrs = no_pdl ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// Queue the conversion (bf16 shadow + norms) of scratch rows [row0, row0+n) whose fp32 values are
// already on the device at `src` (in the scratch arena itself, or a caller's device matrix).
void queue_convert(vsm_ctx* ctx, const float* src, int64_t row0, int64_t n) {
    if (n <= 0) return;
    ConvJob j = {src, nullptr, ctx->scratch.b16 + row0 * VSM_DIM, ctx->scratch.n2 + row0, nullptr, n};
    ctx->pending_conv.push_back(j);
}

// Runs queued conversions as stand-alone kernels until at most `keep` are left (their statistics
// target must be set).
int flush_conversions(vsm_ctx* ctx, size_t keep) {
    while (ctx->pending_conv.size() > keep) {
        const ConvJob j = ctx->pending_conv.back();
        ctx->pending_conv.pop_back();
        if (j.dst_f32) CK(cudaMemcpyAsync(j.dst_f32, j.src, (size_t)j.rows * VSM_DIM * sizeof(float), cudaMemcpyDefault, ctx->stream));
        TRY(launch_convert(ctx, j.dst_f32 ? j.dst_f32 : j.src, j.dst_b16, j.n2, j.rows, j.stats));
    }
    return VSM_OK;
}

// First kernel of a matching call: zeroes the per-call aux block (keeping the scratch-statistics slot
// this call accumulates into), copies the descriptor block from pinned host memory and runs the
// conversions queued by the call.  Scratch norm statistics live in two alternating 8-byte slots: a call
// accumulates into slot call_seq & 1, which the PREVIOUS call's prologue zeroed, and zeroes the other one.
int launch_prologue(vsm_ctx* ctx, size_t aux_bytes, size_t desc_bytes, bool upload_desc) {
    const int stats_slot = (int)(ctx->call_seq & 1u);
    uint32_t* d_scratch_stats = reinterpret_cast<uint32_t*>(ctx->d_aux.p + 32 + 8 * stats_slot);
    for (auto& j : ctx->pending_conv) if (!j.stats) j.stats = d_scratch_stats;
    TRY(flush_conversions(ctx, MAX_CONV));                          // more row sets than one prologue takes
    Prologue pr;
    memset(&pr, 0, sizeof pr);
    pr.aux = reinterpret_cast<uint4*>(ctx->d_aux.p);
    pr.aux_vecs = (uint32_t)(aux_bytes / 16);
    pr.keep_slot = stats_slot;
    pr.desc_src = reinterpret_cast<const uint4*>(ctx->h_desc);      // pinned, mapped (unified addressing)
    pr.desc_dst = reinterpret_cast<uint4*>(ctx->d_desc.p);
    pr.desc_vecs = upload_desc ? (uint32_t)(desc_bytes / 16) : 0u;
    int64_t max_rows = 0;
    for (auto& j : ctx->pending_conv) {
        pr.conv[pr.nconv++] = j;
        max_rows = std::max(max_rows, j.rows);
    }
    ctx->pending_conv.clear();
    const int64_t want = std::max<int64_t>((max_rows + 7) / 8, ((int64_t)pr.aux_vecs + pr.desc_vecs + 1023) / 1024);
    const unsigned blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)ctx->num_sms * 16));
    CK(launch_pdl(prologue_kernel, dim3(blocks), dim3(256), 0, ctx->stream, pr));
    ctx->launches++;
    if (upload_desc) {
        CK(cudaEventRecord(ctx->ev_desc, ctx->stream));
        ctx->desc_copy_pending = true;
    }
    ctx->call_seq++;
    return VSM_OK;
}

inline uint32_t __float_as_uint_host(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

// Plans, uploads and launches: tensor-core pass -> select/re-score (+ re-scan) -> filter.
// The conversions queued by the call (ctx->pending_conv), the zeroing of the per-call aux block
// (counters, unit queue head, hints, result keys) and the descriptor upload are one prologue kernel.
// Scratch norm statistics live in two alternating 8-byte slots: a call accumulates into slot
// call_seq & 1, which the PREVIOUS call's prologue zeroed, and zeroes the other one.
int run_problems(vsm_ctx* ctx, const std::vector<HProblem>& probs, const std::vector<HJob>& jobs,
                 int64_t total_out, int64_t total_matches, int dump_first = 0) {
    const bool exact = ctx->engine == VSM_ENGINE_SIMT;
    const int P = (int)probs.size();
    vsm_ctx::Plan& pl = ctx->plan;
    // everything the plan depends on, byte for byte (HProblem and HJob have no padding)
    std::vector<uint8_t>& key = ctx->plan_key_build;
    {
        int64_t uses_slot = 0;
        for (auto& p : probs) if (!p.t_store) uses_slot = 1 + (ctx->call_seq & 1);
        const int64_t head[8] = {ctx->engine + 16 * ctx->append_max_tiles + (ctx->t2_off ? 8 : 0), ctx->seg_tiles * 2 + (ctx->db_short_slices ? 1 : 0), ctx->num_sms, P,
                                 (int64_t)jobs.size(), total_out, total_matches, uses_slot};
        key.resize(sizeof head + sizeof(HProblem) * probs.size() + sizeof(HJob) * jobs.size());
        memcpy(key.data(), head, sizeof head);
        if (P) memcpy(key.data() + sizeof head, probs.data(), sizeof(HProblem) * probs.size());
        if (!jobs.empty()) memcpy(key.data() + sizeof head + sizeof(HProblem) * probs.size(), jobs.data(), sizeof(HJob) * jobs.size());
        for (auto& p : probs)                                // a run list is part of the plan, not only its address
            if (p.runs && !p.runs->empty()) {
                const uint8_t* b = reinterpret_cast<const uint8_t*>(p.runs->data());
                key.insert(key.end(), b, b + p.runs->size() * sizeof(Run));
            }
    }
    const bool hit = !dump_first && pl.valid && pl.p_desc == ctx->d_desc.p && pl.p_aux == ctx->d_aux.p &&
                     pl.p_stats == ctx->d_store_stats && pl.key == key;
    std::vector<Problem> dp(hit ? 0 : P);
    std::vector<int32_t> qb(P + 1, 0);
    const bool pairs = ctx->engine == VSM_ENGINE_TENSOR_PAIR && !dump_first;
    std::vector<TcUnit> units;
    std::vector<TcUnit2> units2;
    std::vector<int> unit_prob;
    std::vector<SliceInfo> slices;
    int64_t nrecs = 0;
    bool any_classic = false;                                // some problem is answered by select_kernel
    // queries per warp of t2_select_kernel: enough warps to fill the device first, then up to 32 per warp
    int t2_gshift = 1;
    {
        int64_t nq_all = 0;
        for (auto& p : probs) if (p.t2) nq_all += p.nq;
        while (t2_gshift < 5 && (nq_all >> (t2_gshift + 1)) >= (int64_t)ctx->num_sms * 16) t2_gshift++;
    }

    // scheduling granularity: query tiles on SMs, or pairs of query tiles on SM pairs
    int64_t total_qtiles = 0;
    for (auto& p : probs)
        if (p.nq > 0 && p.nt > 0) total_qtiles += pairs ? ((p.nq + TILE_M - 1) / TILE_M + 1) / 2 : (p.nq + TILE_M - 1) / TILE_M;
    const int64_t nworkers = pairs ? ctx->num_sms / 2 : ctx->num_sms;

    // device addresses inside the descriptor block are fixed up after the layout is known
    for (int i = 0; i < P && !hit; i++) {
        const HProblem& hp = probs[i];
        Problem& d = dp[i];
        memset(&d, 0, sizeof d);
        d.q_f32 = hp.q_f32; d.t_f32 = hp.t_f32; d.q_n2 = hp.q_n2;
        d.out_off = hp.out_off; d.nq = hp.nq; d.nt = hp.nt;
        d.slice_off = (int32_t)slices.size();
        d.partial_off = nrecs;
        d.exact = (exact || hp.nt == 0) ? 1 : 0;
        d.skip_ratio2 = hp.skip_ratio2;
        d.ratio = hp.ratio;
        const bool maxima = hp.maxima_only && hp.skip_ratio2 > 0.f && !d.exact && !pairs;
        if (maxima) d.exact |= 2;
        qb[i + 1] = qb[i];                                   // 32-query groups of the tile top-2 problems (t2_select_kernel)
        if (hp.nq <= 0 || hp.nt <= 0) { d.nslices = 0; continue; }
        if (d.exact & 1) {
            any_classic = true;
            if (hp.runs) {
                for (const Run& rn : *hp.runs) { SliceInfo si = {(int32_t)rn.row0, (int32_t)rn.count, -1, 0}; slices.push_back(si); }
            } else {
                SliceInfo si = {0, hp.nt, -1, 0};
                slices.push_back(si);
            }
            d.nslices = (int)slices.size() - d.slice_off;
            continue;
        }
        // the train rows as runs of consecutive rows (one run unless the store has holes); idx0 = logical
        // index of a run's first row, its tensor-map row is hp.t_row + idx0
        std::vector<Run> one{Run{0, hp.nt}};
        const std::vector<Run>& runs = hp.runs ? *hp.runs : one;
        int64_t ntiles = 0;
        for (const Run& rn : runs) ntiles += (rn.count + TILE_N - 1) / TILE_N;
        // ranges: units of at most UNIT_TILES tiles whose count is (close to) a whole number of
        // waves over the SMs, so that the dynamic scheduler ends every CTA at about the same time
        int64_t nranges = 1;
        for (int64_t k = 1; k <= 8192; k++) {
            nranges = std::min<int64_t>(ntiles, std::max<int64_t>(1, nworkers * k / std::max<int64_t>(total_qtiles, 1)));
            if ((ntiles + nranges - 1) / nranges <= UNIT_TILES || nranges == ntiles) break;
        }
        const int tpr_all = (int)((ntiles + nranges - 1) / nranges);
        // slice length: short slices keep an overflow re-scan cheap, long ones keep the record
        // stream small next to the database stream
        const int seg_pref = ctx->seg_tiles > 0 ? ctx->seg_tiles
                                                : (ntiles > 4096 ? (ctx->db_short_slices ? 16 : 64)
                                                                 : (int)std::max<int64_t>(1, std::min<int64_t>(16, ntiles / 16)));
        if (ntiles > 4096 && ctx->seg_tiles == 0) ctx->big_db_nq = hp.nq;
        ctx->last_slice_tiles = seg_pref;
        const int nqt = (hp.nq + TILE_M - 1) / TILE_M;
        // small train sets (pairs, tracking, ragged batches): append records -- one slice per (range, column
        // half), every value above the running threshold stored, no top-4 state in the epilogue (vsm_tc.cuh)
        // -- OFF by default (vsm_opts.reserved[4] = largest train set, in tiles, that uses it): measured SLOWER than
        // the top-4 records (64 ragged pairs: 0.30 ms against 0.18 ms) because with ~1000 rows
        // the running threshold matures so slowly that most (lane, chunk) pairs hold a passing value
        const bool append = !maxima && !pairs && !hp.runs && !dump_first && ntiles <= ctx->append_max_tiles && ctx->seg_tiles == 0;
        if (append) d.exact |= 4;
        // small train sets of pair matching: tile top-2 records -- a slice per (tile, column half), no running state
        // in the epilogue, one exact distance per matching query in select_kernel (vsm_common.cuh, t2_scale)
        const bool t2 = hp.t2 && !ctx->t2_off && !maxima && !append && !pairs && !hp.runs && !dump_first && ctx->seg_tiles == 0 &&
                        ntiles <= T2_MAX_TILES && (hp.t2 == 2 || (hp.skip_ratio2 > 0.f && hp.ratio > 0.f && hp.ratio <= 1.f));
        if (t2) {
            d.exact |= 8 | (hp.t2 == 2 ? 16 : 0);
            d.gshift = hp.t2 == 2 ? t2_gshift : std::max(1, t2_gshift - 2);
            qb[i + 1] = qb[i] + ((hp.nq + (1 << d.gshift) - 1) >> d.gshift);
        } else {
            any_classic = true;
        }
        const int rec_per_slice = append ? APPEND_RECS : 1;
        struct Range { int64_t idx0, count; int tiles_after; int slice0; int seg; };
        std::vector<Range> ranges;
        for (const Run& rn : runs) {
            const int rt = (int)((rn.count + TILE_N - 1) / TILE_N);
            if (rt == 0) continue;
            const int nr = (rt + tpr_all - 1) / tpr_all;
            const int tpr = (rt + nr - 1) / nr;
            // 64 tiles x 128 columns = the 13 index bits of a packed entry; an append unit is one segment
            const int seg = append ? tpr : t2 ? 1 : std::min(std::min(tpr, seg_pref), 64);
            for (int tile0 = 0; tile0 < rt; tile0 += tpr) {
                const int tile1 = std::min(rt, tile0 + tpr);
                Range g;
                g.idx0 = rn.row0 + (int64_t)tile0 * TILE_N;
                g.count = std::min<int64_t>((int64_t)(tile1 - tile0) * TILE_N, rn.row0 + rn.count - g.idx0);
                g.tiles_after = rt - tile1;
                g.slice0 = (int)slices.size() - d.slice_off;
                g.seg = seg;
                for (int64_t i0 = 0; i0 < g.count; i0 += (int64_t)seg * TILE_N) {
                    const int32_t cnt = (int32_t)std::min<int64_t>((int64_t)seg * TILE_N, g.count - i0);
                    for (int h = 0; h < 2; h++) { SliceInfo si = {(int32_t)(g.idx0 + i0), cnt, h, 0}; slices.push_back(si); }
                }
                ranges.push_back(g);
            }
        }
        d.nslices = (int)slices.size() - d.slice_off;
        for (const Range& g : ranges) {
            for (int qt = 0; pairs && qt < nqt; qt += 2) {
                TcUnit2 u;
                memset(&u, 0, sizeof u);
                for (int k = 0; k < 2; k++) {
                    const bool real = qt + k < nqt;
                    const int q = real ? qt + k : qt;                   // an odd tail pairs the last tile with a dummy
                    u.q_n2[k] = hp.q_n2 + (int64_t)q * TILE_M;
                    u.rec_base[k] = nrecs + (int64_t)q * TILE_M * d.nslices + g.slice0;
                    u.q_row[k] = (int32_t)(hp.q_row + (int64_t)q * TILE_M);
                    u.q_valid[k] = real ? std::min(TILE_M, hp.nq - q * TILE_M) : 0;
                }
                u.rec_stride = d.nslices;
                u.t_row = (int32_t)(hp.t_row + g.idx0);
                u.t_index0 = (int32_t)g.idx0;
                u.t_count = (int32_t)g.count;
                u.seg_tiles = g.seg;
                u.maps = (hp.q_store ? 1 : 0) | (hp.t_store ? 2 : 0);
                u.prefetch = (qt == 0 || qt == ((nqt / 2) & ~1)) ? 1 + std::min(tc::L2_AHEAD, g.tiles_after) : 0;
                units2.push_back(u);
                unit_prob.push_back(i);
            }
            for (int qt = 0; !pairs && qt < nqt; qt++) {
                TcUnit u;
                memset(&u, 0, sizeof u);
                u.q_n2 = hp.q_n2 + (int64_t)qt * TILE_M;
                u.t_stats = nullptr;                                    // fixed up below
                u.rec_base = nrecs + ((int64_t)qt * TILE_M * d.nslices + g.slice0) * rec_per_slice;
                u.rec_stride = d.nslices * rec_per_slice;
                if (t2) {                                               // one record per (query, tile): both column halves
                    u.rec_base = nrecs + ((int64_t)qt * TILE_M * d.nslices + g.slice0) / 2;
                    u.rec_stride = d.nslices / 2;
                }
                u.q_row = (int32_t)(hp.q_row + (int64_t)qt * TILE_M);
                u.t_row = (int32_t)(hp.t_row + g.idx0);
                u.t_index0 = (int32_t)g.idx0;
                u.t_count = (int32_t)g.count;
                u.q_valid = std::min(TILE_M, hp.nq - qt * TILE_M);
                u.seg_tiles = g.seg;
                u.maps = (hp.q_store ? 1 : 0) | (hp.t_store ? 2 : 0) | (maxima ? 4 : 0) | (append ? 16 : 0) | (t2 ? 32 : 0);
                u.dump = (dump_first && units.empty()) ? dump_first : 0;
                // 1 + the number of tiles past this unit's end that may be prefetched as well
                u.prefetch = (qt == 0 || qt == nqt / 2) ? 1 + std::min(tc::L2_AHEAD, g.tiles_after) : 0;
                units.push_back(u);
                unit_prob.push_back(i);
            }
        }
        nrecs += t2 ? (int64_t)hp.nq * (d.nslices / 2) : (int64_t)hp.nq * d.nslices * rec_per_slice;
    }

    // Several problems of different sizes in one call (ragged batches, per-keyframe searches): the persistent
    // CTAs take units in list order, so the longest units go first -- otherwise a 8-tile unit handed out last
    // leaves 147 SMs idle for its whole duration (ncu on 64 ragged pairs: 19 % of the warp samples sat in EXIT).
    if (!hit && P > 1 && units.size() > 1) {
        std::vector<size_t> ord(units.size());
        for (size_t k = 0; k < ord.size(); k++) ord[k] = k;
        std::stable_sort(ord.begin(), ord.end(), [&](size_t a, size_t b) { return units[a].t_count > units[b].t_count; });
        std::vector<TcUnit> u2(units.size());
        std::vector<int> p2(units.size());
        for (size_t k = 0; k < ord.size(); k++) { u2[k] = units[ord[k]]; p2[k] = unit_prob[ord[k]]; }
        units.swap(u2);
        unit_prob.swap(p2);
    }

    // descriptor block: [Problem][q_block0][TcUnit][SliceInfo][FilterJob]
    if (!hit) {
        pl.valid = false;
        pl.off_prob = 0;
        const size_t off_qb0 = align16(pl.off_prob + sizeof(Problem) * P);
        pl.off_unit = align16(off_qb0 + sizeof(int32_t) * (P + 1));
        pl.off_slice = align16(pl.off_unit + sizeof(TcUnit) * units.size() + sizeof(TcUnit2) * units2.size());
        pl.off_job = align16(pl.off_slice + sizeof(SliceInfo) * slices.size());
        pl.total = align16(pl.off_job + sizeof(FilterJob) * jobs.size());
        pl.nunits = units.size();
        pl.nunits2 = units2.size();
        pl.nrecs = nrecs;
        pl.qb_total = qb[P];
        pl.any_classic = any_classic;
        pl.t2_gshift = t2_gshift;
        pl.big_db_nq = ctx->big_db_nq;
        pl.slice_tiles = ctx->last_slice_tiles;
    } else {
        ctx->plan_hits++;
        ctx->big_db_nq = pl.big_db_nq;               // the planning loop did not run: same notes as when the plan was made
        ctx->last_slice_tiles = pl.slice_tiles;
    }
    const size_t off_prob = pl.off_prob;
    const size_t off_qb = align16(off_prob + sizeof(Problem) * P);
    const size_t off_unit = pl.off_unit, off_slice = pl.off_slice, off_job = pl.off_job, total = pl.total;
    const size_t nunits = pl.nunits, nunits2 = pl.nunits2;
    nrecs = pl.nrecs;
    TRY(ensure(ctx, ctx->d_desc, total));
    TRY(ensure_host(ctx, ctx->h_desc, ctx->h_desc_cap, total));
    TRY(ensure(ctx, ctx->d_recs, (size_t)std::max<int64_t>(nrecs, 1)));
    const size_t result_bytes = (size_t)total_matches * sizeof(DMatch) + jobs.size() * 2 * sizeof(int32_t);
    // small result lists are written by filter_kernel straight into the pinned host buffer
    // (zero-copy over PCIe): one stream operation and ~10 us less per tracking step; up to 4 MB of CAPACITY --
    // only the matches that exist cross the link (64 ragged pairs: 1.1 MB of capacity, 0.5 MB of matches,
    // 0.268 -> 0.242 ms around the C call)
    ctx->result_on_host = !jobs.empty() && result_bytes <= ((size_t)4 << 20);
    if (ctx->result_on_host) TRY(ensure_host(ctx, ctx->h_result, ctx->h_result_cap, std::max<size_t>(result_bytes, 16)));
    else TRY(ensure(ctx, ctx->d_result, std::max<size_t>(result_bytes, 16)));

    const size_t nout = (size_t)std::max<int64_t>(total_out, 1);
    const size_t off_keys = align16(48 + nout * sizeof(uint32_t));
    const size_t aux_bytes = off_keys + nout * 2 * sizeof(unsigned long long);
    TRY(ensure(ctx, ctx->d_aux, aux_bytes));
    TRY(ensure(ctx, ctx->d_work, (size_t)WORK_CAP));
    if (ctx->aux_zeroed != ctx->d_aux.p) {                   // a fresh allocation: both statistics slots start at zero
        CK(cudaMemsetAsync(ctx->d_aux.p, 0, ctx->d_aux.cap, ctx->stream));
        ctx->aux_zeroed = ctx->d_aux.p;
    }
    const int stats_slot = (int)(ctx->call_seq & 1u);
    ctx->d_counters = reinterpret_cast<unsigned long long*>(ctx->d_aux.p);
    uint32_t* d_scratch_stats = reinterpret_cast<uint32_t*>(ctx->d_aux.p + 32 + 8 * stats_slot);
    uint32_t* d_hints = reinterpret_cast<uint32_t*>(ctx->d_aux.p + 48);
    ctx->d_out_key = reinterpret_cast<unsigned long long*>(ctx->d_aux.p + off_keys);
    bool upload_desc = false;

    if (!hit) {
    for (int i = 0; i < P; i++) dp[i].t_stats = probs[i].t_store ? ctx->d_store_stats : d_scratch_stats;
    for (size_t k = 0; k < units.size(); k++) {
        const int i = unit_prob[k];
        units[k].t_stats = dp[i].t_stats;
        units[k].hint = d_hints + probs[i].out_off + (units[k].q_row - probs[i].q_row);
    }
    for (size_t k = 0; k < units2.size(); k++) {
        const int i = unit_prob[k];
        units2[k].t_stats = dp[i].t_stats;
        for (int c = 0; c < 2; c++) units2[k].hint[c] = d_hints + probs[i].out_off + (units2[k].q_row[c] - probs[i].q_row);
    }

    // build the block; upload it only if it differs from what the device already holds
    // (tracking calls repeat the same shapes frame after frame)
    std::vector<uint8_t>& blk = ctx->desc_build;
    blk.assign(total, 0);
    uint8_t* h = blk.data();
    memcpy(h + off_prob, dp.data(), sizeof(Problem) * P);
    memcpy(h + off_qb, qb.data(), sizeof(int32_t) * (P + 1));
    if (!units.empty()) memcpy(h + off_unit, units.data(), sizeof(TcUnit) * units.size());
    if (!units2.empty()) memcpy(h + off_unit, units2.data(), sizeof(TcUnit2) * units2.size());
    if (!slices.empty()) memcpy(h + off_slice, slices.data(), sizeof(SliceInfo) * slices.size());
    for (size_t j = 0; j < jobs.size(); j++) {
        FilterJob fj;
        memset(&fj, 0, sizeof fj);
        fj.fwd_off = jobs[j].fwd_off; fj.back_off = jobs[j].back_off;
        fj.good_off = jobs[j].good_off; fj.raw_off = jobs[j].raw_off;
        fj.nq = jobs[j].nq; fj.nt = jobs[j].nt; fj.img_idx = jobs[j].img_idx; fj.ratio = jobs[j].ratio;
        const int64_t bp = jobs[j].back_prob;
        fj.back_prob = (bp >= 0 && bp < P && (dp[bp].exact & 16)) ? (int32_t)bp : -1;
        memcpy(h + off_job + j * sizeof(FilterJob), &fj, sizeof fj);
    }
    if (ctx->desc_copy_pending) CK(cudaEventSynchronize(ctx->ev_desc));       // h_desc may still be read by the last prologue
    memcpy(ctx->h_desc, h, total);
    upload_desc = true;
    ctx->loop_p_desc = nullptr;                              // the compact loop search's block is gone
    if (!dump_first) {
        pl.key = key;
        pl.p_desc = ctx->d_desc.p;
        pl.p_aux = ctx->d_aux.p;
        pl.p_stats = ctx->d_store_stats;
        pl.valid = true;
    }
    }   // !hit

    TRY(launch_prologue(ctx, aux_bytes, total, upload_desc));

    uint8_t* dd = ctx->d_desc.p;
    if (ctx->profiling) {
        const uint32_t slot = ctx->tc_ring_head++ % vsm_ctx::TC_RING;
        ctx->ev_tc0 = ctx->tc_ring0[slot];
        ctx->ev_tc1 = ctx->tc_ring1[slot];
        ctx->tc_ring_valid[slot] = nunits || nunits2;
        CK(cudaEventRecord(ctx->ev_tc0, ctx->stream));
    }
    if (nunits2) {
        const CUtensorMap& ms = ctx->scratch.map;
        const CUtensorMap& mt = ctx->store.b16 ? ctx->store.map : ctx->scratch.map;
        const unsigned nclusters = (unsigned)std::min<size_t>(nunits2, (size_t)(ctx->num_sms / 2));
        tc2::tc_top3_pair_kernel<<<nclusters * 2, tc::THREADS, tc2::SMEM2_BYTES, ctx->stream>>>(
            ms, mt, reinterpret_cast<const TcUnit2*>(dd + off_unit), (int)nunits2, ctx->d_recs.p);
        ctx->launches++;
        CK(cudaGetLastError());
        if (ctx->profiling) {
            CK(cudaEventRecord(ctx->ev_tc1, ctx->stream));
            ctx->timed_tc = true;
        }
    }
    if (nunits) {
        const CUtensorMap& ms = ctx->scratch.map;
        const CUtensorMap& mt = ctx->store.b16 ? ctx->store.map : ctx->scratch.map;
        const unsigned grid = (unsigned)std::min<size_t>(nunits, (size_t)ctx->num_sms);
        uint32_t* d_unit_counter = reinterpret_cast<uint32_t*>(ctx->d_aux.p + 24);
        CK(launch_pdl(dump_first ? tc::tc_top3_kernel<true> : tc::tc_top3_kernel<false>, dim3(grid), dim3(tc::THREADS),
                      tc::SMEM_BYTES, ctx->stream, ms, mt, reinterpret_cast<const TcUnit*>(dd + off_unit), (int)nunits,
                      (const FusedArgs*)nullptr, d_unit_counter, ctx->d_recs.p, ctx->d_dump));
        ctx->launches++;
        if (ctx->profiling) {
            CK(cudaEventRecord(ctx->ev_tc1, ctx->stream));
            ctx->timed_tc = true;
        }
    }
    if (pl.qb_total > 0 || pl.any_classic) {
        if (pl.qb_total > 0) {
            CK(launch_pdl(t2_select_kernel, dim3((unsigned)((pl.qb_total + T2_SELECT_WARPS - 1) / T2_SELECT_WARPS)),
                          dim3(T2_SELECT_WARPS * 32), 0, ctx->stream, reinterpret_cast<const Problem*>(dd + off_prob), P,
                          reinterpret_cast<const int32_t*>(dd + off_qb), (const PartialRec*)ctx->d_recs.p,
                          reinterpret_cast<const SliceInfo*>(dd + off_slice), ctx->d_out_key, ctx->d_counters,
                          ctx->d_work.p, ctx->work_cap));
            ctx->launches++;
        }
        int max_blocks = 0;
        for (int i = 0; i < P; i++) max_blocks = std::max(max_blocks, (probs[i].nq + SELECT_WARPS - 1) / SELECT_WARPS);
        for (int p0 = 0; p0 < P && pl.any_classic; p0 += 65535) {         // gridDim.y limit
            const int np = std::min(65535, P - p0);
            CK(launch_pdl(select_kernel, dim3((unsigned)max_blocks, (unsigned)np), dim3(SELECT_WARPS * 32), 0, ctx->stream,
                          reinterpret_cast<const Problem*>(dd + off_prob), p0, (const PartialRec*)ctx->d_recs.p,
                          reinterpret_cast<const SliceInfo*>(dd + off_slice), ctx->d_out_key, ctx->d_counters,
                          ctx->d_work.p, ctx->work_cap));
            ctx->launches++;
        }
        if (nunits || nunits2) {
            // exact re-scan of the slices whose top-3 overflowed (usually none: the blocks exit at once)
            CK(launch_pdl(rescan_kernel, dim3((unsigned)ctx->num_sms * 2), dim3(256), 0, ctx->stream,
                          (const WorkItem*)ctx->d_work.p, (const unsigned long long*)ctx->d_counters, ctx->work_cap));
            ctx->launches++;
        }
        if (ctx->profiling) {
            CK(cudaEventRecord(ctx->ev_sel1, ctx->stream));
            ctx->timed_sel = true;
        }
    }
    if (!jobs.empty()) {
        uint8_t* rbase = ctx->result_on_host ? ctx->h_result : ctx->d_result.p;
        DMatch* dm = reinterpret_cast<DMatch*>(rbase);
        int32_t* dc = reinterpret_cast<int32_t*>(rbase + (size_t)total_matches * sizeof(DMatch));
        CK(launch_pdl(filter_kernel, dim3((unsigned)jobs.size()), dim3(FILTER_THREADS), 0, ctx->stream,
                      reinterpret_cast<const FilterJob*>(dd + off_job), ctx->d_out_key, dm, dc,
                      reinterpret_cast<const Problem*>(dd + off_prob), (const PartialRec*)ctx->d_recs.p,
                      reinterpret_cast<const SliceInfo*>(dd + off_slice)));
        ctx->launches++;
    }
    return VSM_OK;
}

// Host-mapped device address of a pinned host buffer, or nullptr if `p` is pageable memory.
const float* host_mapped(const float* p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (attr.type != cudaMemoryTypeHost || !attr.devicePointer) return nullptr;
    return static_cast<const float*>(attr.devicePointer);
}

// up to this many rows (1 KB each) the prologue reads a pinned buffer straight over PCIe
// (VSM_ZERO_COPY_ROWS overrides it, for tuning)
static const int64_t ZERO_COPY_ROWS = getenv("VSM_ZERO_COPY_ROWS") ? atoll(getenv("VSM_ZERO_COPY_ROWS")) : 2048;

// Host rows -> scratch rows [row0, row0+n): fp32 master + queued conversion.  A small pinned buffer
// is not copied here at all: the call's prologue kernel reads it over PCIe (one launch for upload,
// conversion and norms).  Otherwise a DMA now, conversion in the prologue.
int upload_scratch(vsm_ctx* ctx, const float* src, int64_t row0, int64_t n, int64_t stride_bytes = VSM_DIM * sizeof(float)) {
    if (n <= 0) return VSM_OK;
    float* dst = ctx->scratch.f32 + row0 * VSM_DIM;
    if (stride_bytes != (int64_t)(VSM_DIM * sizeof(float))) {
        // rows of a wider matrix (a cv::Mat ROI / a non-continuous Mat): one strided DMA packs them
        CK(cudaMemcpy2DAsync(dst, VSM_DIM * sizeof(float), src, (size_t)stride_bytes, VSM_DIM * sizeof(float), (size_t)n,
                             cudaMemcpyHostToDevice, ctx->stream));
        ConvJob j = {dst, nullptr, ctx->scratch.b16 + row0 * VSM_DIM, ctx->scratch.n2 + row0, nullptr, n};
        ctx->pending_conv.push_back(j);
        return VSM_OK;
    }
    const float* mapped = n <= ZERO_COPY_ROWS ? host_mapped(src) : nullptr;
    ConvJob j = {mapped ? mapped : dst, mapped ? dst : nullptr, ctx->scratch.b16 + row0 * VSM_DIM, ctx->scratch.n2 + row0,
                 nullptr, n};
    if (!mapped) CK(cudaMemcpyAsync(dst, src, (size_t)n * VSM_DIM * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    ctx->pending_conv.push_back(j);
    return VSM_OK;
}

// Host rows -> scratch fp32 rows only (no bf16 shadow: track_local_map scores in fp64 from these).
int upload_plain(vsm_ctx* ctx, const float* src, int64_t row0, int64_t n) {
    if (n <= 0) return VSM_OK;
    CK(cudaMemcpyAsync(ctx->scratch.f32 + row0 * VSM_DIM, src, (size_t)n * VSM_DIM * sizeof(float),
                       cudaMemcpyHostToDevice, ctx->stream));
    return VSM_OK;
}

int fetch_result(vsm_ctx* ctx, size_t bytes) {
    if (ctx->result_on_host) return VSM_OK;                  // already there once the stream is idle
    TRY(ensure_host(ctx, ctx->h_result, ctx->h_result_cap, std::max<size_t>(bytes, 16)));
    if (bytes) CK(cudaMemcpyAsync(ctx->h_result, ctx->d_result.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return VSM_OK;
}

// [query][2] result keys of the first nq queries -> pinned host buffer
int fetch_keys(vsm_ctx* ctx, int64_t nq) {
    const size_t nb = (size_t)nq * 2 * sizeof(unsigned long long);
    TRY(ensure_host(ctx, ctx->h_result, ctx->h_result_cap, std::max<size_t>(nb, 16)));
    CK(cudaMemcpyAsync(ctx->h_result, ctx->d_out_key, nb, cudaMemcpyDeviceToHost, ctx->stream));
    return VSM_OK;
}

// host twin of key_decode (vsm_kernels.cuh)
void decode_key(unsigned long long k, int64_t& idx, float& dist) {
    if (k == 0ull) { idx = -1; dist = FLT_MAX; return; }
    k = ~k;
    idx = (int64_t)(uint32_t)k;
    uint32_t b = (uint32_t)(k >> 32);
    memcpy(&dist, &b, 4);
}

HProblem scratch_vs_scratch(vsm_ctx* ctx, int64_t q_row, int nq, int64_t t_row, int nt, int64_t out_off) {
    HProblem p;
    p.q_f32 = ctx->scratch.f32 + q_row * VSM_DIM; p.q_n2 = ctx->scratch.n2 + q_row; p.q_row = q_row;
    p.q_store = 0; p.nq = nq;
    p.t_f32 = ctx->scratch.f32 + t_row * VSM_DIM; p.t_row = t_row; p.t_store = 0; p.nt = nt;
    p.out_off = out_off;
    return p;
}

}  // namespace

// ---- C ABI ---------------------------------------------------------------------------
extern "C" {

void vsm_default_opts(vsm_opts* o) {
    if (!o) return;
    memset(o, 0, sizeof *o);
    o->engine = VSM_ENGINE_AUTO;
}

const char* vsm_version(void) { return "vsm-b200 0.1 (sm_100a, tcgen05)"; }

const char* vsm_last_error(const vsm_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int vsm_create(const vsm_opts* opts, vsm_ctx** out) {
    if (!out) return VSM_ERR_INVALID;
    *out = nullptr;
    vsm_opts o;
    if (opts) o = *opts; else vsm_default_opts(&o);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || o.device < 0 || o.device >= ndev) {
        cudaGetLastError();
        g_create_error = "no usable CUDA device (this library has no CPU fallback)";
        return VSM_ERR_NO_DEVICE;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, o.device) != cudaSuccess || prop.major != 10) {
        g_create_error = "device is not sm_100 (Blackwell B200); kernels are built for sm_100a only";
        return VSM_ERR_NO_DEVICE;
    }
    vsm_ctx* ctx = new vsm_ctx();
    ctx->device = o.device;
    ctx->engine = o.engine;
    ctx->num_sms = prop.multiProcessorCount;
    if (o.reserved[0] > 0) ctx->seg_tiles = o.reserved[0];
    else if (getenv("VSM_SEG_TILES")) ctx->seg_tiles = std::max(0, atoi(getenv("VSM_SEG_TILES")));      // A/B switch
    ctx->work_cap = o.reserved[1] > 0 ? std::min<uint32_t>((uint32_t)o.reserved[1], WORK_CAP) : WORK_CAP;
    auto bail = [&](int code) {
        g_create_error = ctx->err;
        vsm_destroy(ctx);
        return code;
    };
    auto init = [&]() -> int {
        CK(cudaSetDevice(ctx->device));
        CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        CK(cudaEventCreate(&ctx->ev0));
        CK(cudaEventCreate(&ctx->ev1));
        for (int i = 0; i < vsm_ctx::TC_RING; i++) {
            CK(cudaEventCreate(&ctx->tc_ring0[i]));
            CK(cudaEventCreate(&ctx->tc_ring1[i]));
        }
        ctx->ev_tc0 = ctx->tc_ring0[0];
        ctx->ev_tc1 = ctx->tc_ring1[0];
        CK(cudaEventCreate(&ctx->ev_sel1));
        CK(cudaEventCreateWithFlags(&ctx->ev_desc, cudaEventDisableTiming));
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) return fail(ctx, VSM_ERR_CUDA, "cuTensorMapEncodeTiled not available");
        ctx->encode = reinterpret_cast<PFN_encodeTiled>(fn);
        // virtual memory management: arenas grow by mapping chunks (VSM_NO_VMM=1 forces the copying path, for A/B tests)
        {
            auto sym = [&](const char* name) -> void* {
                void* f = nullptr;
                cudaDriverEntryPointQueryResult q;
                if (cudaGetDriverEntryPoint(name, &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
                    cudaGetLastError();
                    return nullptr;
                }
                return f;
            };
            ctx->p_reserve = reinterpret_cast<PFN_cuMemAddressReserve>(sym("cuMemAddressReserve"));
            ctx->p_addr_free = reinterpret_cast<PFN_cuMemAddressFree>(sym("cuMemAddressFree"));
            ctx->p_create = reinterpret_cast<PFN_cuMemCreate>(sym("cuMemCreate"));
            ctx->p_release = reinterpret_cast<PFN_cuMemRelease>(sym("cuMemRelease"));
            ctx->p_map = reinterpret_cast<PFN_cuMemMap>(sym("cuMemMap"));
            ctx->p_unmap = reinterpret_cast<PFN_cuMemUnmap>(sym("cuMemUnmap"));
            ctx->p_set_access = reinterpret_cast<PFN_cuMemSetAccess>(sym("cuMemSetAccess"));
            ctx->p_gran = reinterpret_cast<PFN_cuMemGetAllocationGranularity>(sym("cuMemGetAllocationGranularity"));
            if (ctx->p_reserve && ctx->p_addr_free && ctx->p_create && ctx->p_release && ctx->p_map && ctx->p_unmap &&
                ctx->p_set_access && ctx->p_gran && !getenv("VSM_NO_VMM")) {
                CUmemAllocationProp prop;
                memset(&prop, 0, sizeof prop);
                prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
                prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
                prop.location.id = ctx->device;
                size_t g = 0;
                if (ctx->p_gran(&g, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS && g > 0) {
                    ctx->vmm_gran = g;
                    ctx->vmm_ok = true;
                }
            }
        }
        CK(cudaHostAlloc(reinterpret_cast<void**>(&ctx->h_status), 64, cudaHostAllocMapped));
        memset(ctx->h_status, 0, 64);
        if (o.reserved[2] > 0) ctx->ring_depth = o.reserved[2];
        if (o.reserved[3] > 0) ctx->pair_cap = (uint32_t)o.reserved[3];
        if (o.reserved[4] > 0) ctx->append_max_tiles = o.reserved[4];
        else if (getenv("VSM_APPEND_TILES")) ctx->append_max_tiles = atoll(getenv("VSM_APPEND_TILES"));
        ctx->t2_off = o.reserved[5] == 1 || (getenv("VSM_NO_T2") && atoi(getenv("VSM_NO_T2")));
        CK(cudaMalloc(&ctx->d_store_stats, 16));
        // on the context's own (non-blocking) stream and waited for: a cudaMemset on the legacy default
        // stream is not ordered with it and could land AFTER the first conversion's atomicMax into the
        // slot -- an "empty" norm range shrinks dot_margin to ~6e-5 and the exactness guarantee is gone
        CK(cudaMemsetAsync(ctx->d_store_stats, 0, 16, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaMalloc(&ctx->d_dump, TILE_M * TILE_N * sizeof(float)));
        CK(cudaFuncSetAttribute(tc::tc_top3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::SMEM_BYTES));
        CK(cudaFuncSetAttribute(tc::tc_top3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::SMEM_BYTES));
        CK(cudaFuncSetAttribute(tc2::tc_top3_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc2::SMEM2_BYTES));
        TRY(arena_reserve(ctx, ctx->scratch, o.scratch_rows > 0 ? o.scratch_rows : 8192, 0));
        if (o.store_rows > 0) TRY(arena_reserve(ctx, ctx->store, o.store_rows, 0));
        return VSM_OK;
    };
    int s = init();
    if (s != VSM_OK) return bail(s);
    *out = ctx;
    return VSM_OK;
}

void vsm_destroy(vsm_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    arena_free(ctx, ctx->scratch);
    arena_free(ctx, ctx->store);
    for (int r = 0; r < ctx->xchg_world; r++)
        if (r != ctx->xchg_rank && ctx->xchg_peer_ptr[r]) cudaIpcCloseMemHandle(ctx->xchg_peer_ptr[r]);
    if (ctx->xchg_buf) cudaFree(ctx->xchg_buf);
    if (ctx->d_store_stats) cudaFree(ctx->d_store_stats);
    if (ctx->h_status) cudaFreeHost(ctx->h_status);
    if (ctx->d_desc.p) cudaFree(ctx->d_desc.p);
    if (ctx->h_desc) cudaFreeHost(ctx->h_desc);
    if (ctx->d_recs.p) cudaFree(ctx->d_recs.p);
    if (ctx->d_result.p) cudaFree(ctx->d_result.p);
    if (ctx->h_result) cudaFreeHost(ctx->h_result);
    if (ctx->d_aux.p) cudaFree(ctx->d_aux.p);
    if (ctx->d_work.p) cudaFree(ctx->d_work.p);
    if (ctx->d_sel.p) cudaFree(ctx->d_sel.p);
    if (ctx->d_track.p) cudaFree(ctx->d_track.p);
    if (ctx->d_loop.p) cudaFree(ctx->d_loop.p);
    if (ctx->d_pt_tmp.p) cudaFree(ctx->d_pt_tmp.p);
    grow_arr_free(ctx, ctx->pt_f32);
    grow_arr_free(ctx, ctx->pt_valid);
    grow_arr_free(ctx, ctx->pt_log);
    if (ctx->d_dump) cudaFree(ctx->d_dump);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    for (int i = 0; i < vsm_ctx::TC_RING; i++) {
        if (ctx->tc_ring0[i]) cudaEventDestroy(ctx->tc_ring0[i]);
        if (ctx->tc_ring1[i]) cudaEventDestroy(ctx->tc_ring1[i]);
    }
    if (ctx->ev_sel1) cudaEventDestroy(ctx->ev_sel1);
    if (ctx->ev_desc) cudaEventDestroy(ctx->ev_desc);
    if (ctx->stream && ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int vsm_get_stats(vsm_ctx* ctx, vsm_stats* out) {
    if (!ctx || !out) return VSM_ERR_INVALID;
    if (ctx->pending_stats) {                  // an asynchronous call: wait for it, then read its events
        CK(cudaStreamSynchronize(ctx->stream));
        TRY(collect_stats(ctx));
    }
    *out = ctx->stats;
    return VSM_OK;
}

int vsm_tc_history(vsm_ctx* ctx, float* ms, int32_t n, int32_t* n_out) {
    if (!ctx || !ms || n < 0 || !n_out) return ctx ? fail(ctx, VSM_ERR_INVALID, "vsm_tc_history: bad argument") : VSM_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    const uint32_t have = std::min<uint32_t>(ctx->tc_ring_head, (uint32_t)vsm_ctx::TC_RING);
    const uint32_t take = std::min<uint32_t>(have, (uint32_t)n);
    int32_t k = 0;
    for (uint32_t i = ctx->tc_ring_head - take; i != ctx->tc_ring_head; i++) {      // oldest first
        const uint32_t slot = i % vsm_ctx::TC_RING;
        float t = 0.f;
        if (ctx->tc_ring_valid[slot]) CK(cudaEventElapsedTime(&t, ctx->tc_ring0[slot], ctx->tc_ring1[slot]));
        ms[k++] = t;
    }
    *n_out = k;
    return VSM_OK;
}

int vsm_host_alloc(void** ptr, int64_t bytes) {
    if (!ptr || bytes <= 0) return VSM_ERR_INVALID;
    return cudaMallocHost(ptr, (size_t)bytes) == cudaSuccess ? VSM_OK : VSM_ERR_CUDA;
}
void vsm_host_free(void* ptr) { if (ptr) cudaFreeHost(ptr); }

void* vsm_stream(vsm_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int vsm_set_stream(vsm_ctx* ctx, void* stream) {
    if (!ctx) return VSM_ERR_INVALID;
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream && ctx->stream) CK(cudaStreamDestroy(ctx->stream));
    ctx->stream = (cudaStream_t)stream;
    ctx->own_stream = false;
    return VSM_OK;
}

int vsm_set_profiling(vsm_ctx* ctx, int32_t on) {
    if (!ctx) return VSM_ERR_INVALID;
    ctx->profiling = on != 0;
    return VSM_OK;
}

int vsm_sync(vsm_ctx* ctx) {
    if (!ctx) return VSM_ERR_INVALID;
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->h_status && ctx->h_status[0] == 0x7100u) {          // XCHG_TIMEOUT of an asynchronous exchange
        ctx->h_status[0] = 0;
        ctx->err = "peer-memory exchange timed out waiting for a peer rank";
        return VSM_ERR_TIMEOUT;
    }
    return VSM_OK;
}

// ---- pair matching ---------------------------------------------------------------------
static bool bad_stride(int64_t s) { return s < (int64_t)(VSM_DIM * sizeof(float)) || (s & 3); }

int vsm_knn2(vsm_ctx* ctx, const float* query, int32_t nq, const float* train, int32_t nt, int32_t* idx, float* dist) {
    return vsm_knn2_strided(ctx, query, nq, VSM_DIM * sizeof(float), train, nt, VSM_DIM * sizeof(float), idx, dist);
}

int vsm_knn2_strided(vsm_ctx* ctx, const float* query, int32_t nq, int64_t q_stride, const float* train, int32_t nt,
                     int64_t t_stride, int32_t* idx, float* dist) {
    if (!ctx || nq < 0 || nt < 0 || (nq > 0 && (!query || !idx || !dist)) || (nt > 0 && !train) || bad_stride(q_stride) ||
        bad_stride(t_stride))
        return ctx ? fail(ctx, VSM_ERR_INVALID, "vsm_knn2: bad argument (row stride must be >= 1024 bytes)") : VSM_ERR_INVALID;
    if (nq == 0) return VSM_OK;
    TRY(begin_call(ctx));
    TRY(arena_reserve(ctx, ctx->scratch, (int64_t)nq + nt, 0));
    TRY(upload_scratch(ctx, query, 0, nq, q_stride));
    TRY(upload_scratch(ctx, train, nq, nt, t_stride));
    std::vector<HProblem> probs{scratch_vs_scratch(ctx, 0, nq, nq, nt, 0)};
    TRY(run_problems(ctx, probs, {}, nq, 0));
    TRY(fetch_keys(ctx, nq));
    TRY(end_call(ctx, true));
    const unsigned long long* k = reinterpret_cast<const unsigned long long*>(ctx->h_result);
    for (int i = 0; i < nq * 2; i++) {
        int64_t j;
        decode_key(k[i], j, dist[i]);
        idx[i] = (int32_t)j;
    }
    return VSM_OK;
}

static int match_common(vsm_ctx* ctx, std::vector<HProblem>& probs, int nq, int nt, float ratio, int mutual,
                        vsm_dmatch* good, int32_t* n_good, vsm_dmatch* raw, int32_t* n_raw) {
    const bool want_raw = raw && n_raw;
    if (!want_raw) {                                             // forward problem: a survivor must pass the ratio test
        probs[0].skip_ratio2 = skip_r2(ratio);
        probs[0].t2 = 1;
        probs[0].ratio = ratio;
    }
    // reverse problem: only its nearest index is read.  Only next to a tile top-2 forward problem: a call that mixes the
    // record kinds runs both select kernels, which costs a tracking step with a raw list 3 us (72.5 against 69.7 us)
    if (mutual && probs.size() > 1 && !want_raw) probs[1].t2 = 2;
    HJob j;
    j.fwd_off = 0; j.back_off = mutual ? nq : -1; j.good_off = 0; j.raw_off = want_raw ? nq : -1;
    j.nq = nq; j.nt = nt; j.img_idx = 0; j.ratio = ratio;
    j.back_prob = mutual && probs.size() > 1 ? 1 : -1;
    const int64_t total_matches = (int64_t)nq * (want_raw ? 2 : 1);
    TRY(run_problems(ctx, probs, {j}, (int64_t)nq + (mutual ? nt : 0), total_matches));
    const size_t bytes = (size_t)total_matches * sizeof(DMatch) + 2 * sizeof(int32_t);
    TRY(fetch_result(ctx, bytes));
    TRY(end_call(ctx, true));
    const int32_t* c = reinterpret_cast<const int32_t*>(ctx->h_result + (size_t)total_matches * sizeof(DMatch));
    *n_good = c[0];
    memcpy(good, ctx->h_result, (size_t)c[0] * sizeof(DMatch));
    if (want_raw) {
        *n_raw = c[1];
        memcpy(raw, ctx->h_result + (size_t)nq * sizeof(DMatch), (size_t)c[1] * sizeof(DMatch));
    }
    return VSM_OK;
}

int vsm_match_pair(vsm_ctx* ctx, const float* query, int32_t nq, const float* train, int32_t nt, float ratio,
                   int32_t mutual, vsm_dmatch* good, int32_t* n_good, vsm_dmatch* raw, int32_t* n_raw) {
    return vsm_match_pair_strided(ctx, query, nq, VSM_DIM * sizeof(float), train, nt, VSM_DIM * sizeof(float), ratio, mutual,
                                  good, n_good, raw, n_raw);
}

int vsm_match_pair_strided(vsm_ctx* ctx, const float* query, int32_t nq, int64_t q_stride, const float* train, int32_t nt,
                           int64_t t_stride, float ratio, int32_t mutual, vsm_dmatch* good, int32_t* n_good, vsm_dmatch* raw,
                           int32_t* n_raw) {
    if (!ctx || nq < 0 || nt < 0 || !n_good || (nq > 0 && (!query || !good)) || (nt > 0 && !train) || bad_stride(q_stride) ||
        bad_stride(t_stride))
        return ctx ? fail(ctx, VSM_ERR_INVALID, "vsm_match_pair: bad argument (row stride must be >= 1024 bytes)") : VSM_ERR_INVALID;
    *n_good = 0;
    if (n_raw) *n_raw = 0;
    if (nq == 0 || nt == 0) return VSM_OK;                              // src/Slam.cpp:1143
    TRY(begin_call(ctx));
    TRY(arena_reserve(ctx, ctx->scratch, (int64_t)nq + nt, 0));
    TRY(upload_scratch(ctx, query, 0, nq, q_stride));
    TRY(upload_scratch(ctx, train, nq, nt, t_stride));
    std::vector<HProblem> probs{scratch_vs_scratch(ctx, 0, nq, nq, nt, 0)};
    if (mutual) probs.push_back(scratch_vs_scratch(ctx, nq, nt, 0, nq, nq));
    return match_common(ctx, probs, nq, nt, ratio, mutual, good, n_good, raw, n_raw);
}

int vsm_match_batch(vsm_ctx* ctx, int32_t n_pairs, const float* query, const int32_t* q_off, const float* train,
                    const int32_t* t_off, float ratio, int32_t mutual, vsm_dmatch* good, int32_t* n_good) {
    if (!ctx || n_pairs < 0 || (n_pairs > 0 && (!q_off || !t_off || !n_good)))
        return ctx ? fail(ctx, VSM_ERR_INVALID, "vsm_match_batch: bad argument") : VSM_ERR_INVALID;
    if (n_pairs == 0) return VSM_OK;
    const int64_t NQ = q_off[n_pairs], NT = t_off[n_pairs];
    for (int p = 0; p < n_pairs; p++) {
        if (q_off[p + 1] < q_off[p] || t_off[p + 1] < t_off[p])
            return fail(ctx, VSM_ERR_INVALID, "vsm_match_batch: offsets must be non-decreasing");
        n_good[p] = 0;
    }
    if (NQ == 0) return VSM_OK;
    if (!query || !good || (NT > 0 && !train)) return fail(ctx, VSM_ERR_INVALID, "vsm_match_batch: null buffer");
    TRY(begin_call(ctx));
    TRY(arena_reserve(ctx, ctx->scratch, NQ + NT, 0));
    TRY(upload_scratch(ctx, query, 0, NQ));
    TRY(upload_scratch(ctx, train, NQ, NT));
    std::vector<HProblem> probs;
    std::vector<HJob> jobs;
    for (int p = 0; p < n_pairs; p++) {
        const int nq = q_off[p + 1] - q_off[p], nt = t_off[p + 1] - t_off[p];
        probs.push_back(scratch_vs_scratch(ctx, q_off[p], nq, NQ + t_off[p], nt, q_off[p]));
        probs.back().skip_ratio2 = skip_r2(ratio);               // forward problem only; the reverse one feeds the mutual test
        probs.back().t2 = 1;
        probs.back().ratio = ratio;
        if (mutual) {
            probs.push_back(scratch_vs_scratch(ctx, NQ + t_off[p], nt, q_off[p], nq, NQ + t_off[p]));
            probs.back().t2 = 2;
        }
        HJob j;
        j.fwd_off = q_off[p]; j.back_off = mutual ? NQ + t_off[p] : -1; j.good_off = q_off[p]; j.raw_off = -1;
        j.nq = nq; j.nt = nt; j.img_idx = 0; j.ratio = ratio;
        j.back_prob = mutual ? (int64_t)probs.size() - 1 : -1;
        jobs.push_back(j);
    }
    TRY(run_problems(ctx, probs, jobs, NQ + (mutual ? NT : 0), NQ));
    const size_t bytes = (size_t)NQ * sizeof(DMatch) + (size_t)n_pairs * 2 * sizeof(int32_t);
    TRY(fetch_result(ctx, bytes));
    TRY(end_call(ctx, true));
    const int32_t* c = reinterpret_cast<const int32_t*>(ctx->h_result + (size_t)NQ * sizeof(DMatch));
    for (int p = 0; p < n_pairs; p++) {
        n_good[p] = c[2 * p];
        memcpy(good + q_off[p], ctx->h_result + (size_t)q_off[p] * sizeof(DMatch), (size_t)c[2 * p] * sizeof(DMatch));
    }
    return VSM_OK;
}

// ---- keyframe store ------------------------------------------------------------------
// Rows for a new frame: the first removed range that is large enough, else the end of the store.
static int store_alloc_rows(vsm_ctx* ctx, int64_t n, int64_t* row0) {
    if (n <= 0) { *row0 = ctx->store_rows; return VSM_OK; }
    for (size_t k = 0; k < ctx->free_rows.size(); k++) {
        Run& f = ctx->free_rows[k];
        if (f.count < n) continue;
        *row0 = f.row0;
        f.row0 += n; f.count -= n;
        if (f.count == 0) ctx->free_rows.erase(ctx->free_rows.begin() + (long)k);
        return VSM_OK;
    }
    TRY(arena_reserve(ctx, ctx->store, ctx->store_rows + n, ctx->store_rows));
    *row0 = ctx->store_rows;
    ctx->store_rows += n;
    return VSM_OK;
}

static void store_free_rows(vsm_ctx* ctx, int64_t row0, int64_t n) {
    if (n <= 0) return;
    if (row0 + n == ctx->store_rows) {                       // the tail: lower the high-water mark instead
        ctx->store_rows = row0;
        while (!ctx->free_rows.empty() && ctx->free_rows.back().row0 + ctx->free_rows.back().count == ctx->store_rows) {
            ctx->store_rows = ctx->free_rows.back().row0;
            ctx->free_rows.pop_back();
        }
        return;
    }
    auto it = std::lower_bound(ctx->free_rows.begin(), ctx->free_rows.end(), row0,
                               [](const Run& r, int64_t v) { return r.row0 < v; });
    it = ctx->free_rows.insert(it, Run{row0, n});
    if (it + 1 != ctx->free_rows.end() && it->row0 + it->count == (it + 1)->row0) {      // merge with the next range
        it->count += (it + 1)->count;
        ctx->free_rows.erase(it + 1);
    }
    if (it != ctx->free_rows.begin() && (it - 1)->row0 + (it - 1)->count == it->row0) {  // and with the previous one
        (it - 1)->count += it->count;
        ctx->free_rows.erase(it);
    }
}

static int32_t store_new_seg(vsm_ctx* ctx, int64_t row0, int32_t n, int32_t frame_id, bool is_kf) {
    Seg s = {row0, n, frame_id, ctx->next_seq++, 1, (uint8_t)(is_kf ? 1 : 0)};
    int32_t h;
    if (!ctx->free_handles.empty()) { h = ctx->free_handles.back(); ctx->free_handles.pop_back(); ctx->segs[h] = s; }
    else { h = (int32_t)ctx->segs.size(); ctx->segs.push_back(s); }
    if (is_kf) ctx->kf_order.push_back(h);                   // seq is the largest so far: stays sorted
    else ctx->plain_ring.push_back(h);
    ctx->kf_runs_valid = false;
    return h;
}

static bool seg_live(const vsm_ctx* ctx, int32_t h) { return h >= 0 && h < (int32_t)ctx->segs.size() && ctx->segs[h].live; }

static void store_remove_seg(vsm_ctx* ctx, int32_t h) {
    Seg& s = ctx->segs[h];
    auto& list = s.is_kf ? ctx->kf_order : ctx->plain_ring;
    list.erase(std::find(list.begin(), list.end(), h));
    store_free_rows(ctx, s.row0, s.count);
    s.live = 0;
    ctx->free_handles.push_back(h);
    ctx->kf_runs_valid = false;
}

// Merged row runs of the live keyframes, ascending: what a global search over the store scans.
static const std::vector<Run>& store_runs(vsm_ctx* ctx) {
    if (!ctx->kf_runs_valid) {
        std::vector<Run> r;
        for (int32_t h : ctx->kf_order) if (ctx->segs[h].count > 0) r.push_back(Run{ctx->segs[h].row0, ctx->segs[h].count});
        std::sort(r.begin(), r.end(), [](const Run& a, const Run& b) { return a.row0 < b.row0; });
        ctx->kf_runs.clear();
        for (const Run& x : r) {
            if (!ctx->kf_runs.empty() && ctx->kf_runs.back().row0 + ctx->kf_runs.back().count == x.row0) ctx->kf_runs.back().count += x.count;
            else ctx->kf_runs.push_back(x);
        }
        ctx->kf_runs_valid = true;
    }
    return ctx->kf_runs;
}

static bool host_pinned(const void* p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return attr.type == cudaMemoryTypeHost;
}

static int store_append(vsm_ctx* ctx, int32_t frame_id, const float* src, int64_t n, cudaMemcpyKind kind,
                        bool is_kf, int32_t* handle, int64_t stride_bytes = VSM_DIM * sizeof(float)) {
    if (!ctx->store.own_f32) return fail(ctx, VSM_ERR_INVALID, "store was adopted from a device matrix; clear it first");
    int64_t row0 = 0;
    TRY(store_alloc_rows(ctx, n, &row0));
    if (n > 0) {
        CK(cudaMemcpy2DAsync(ctx->store.f32 + row0 * VSM_DIM, VSM_DIM * sizeof(float), src, (size_t)stride_bytes,
                             VSM_DIM * sizeof(float), (size_t)n, kind, ctx->stream));
        TRY(launch_convert(ctx, ctx->store.f32 + row0 * VSM_DIM, ctx->store.b16 + row0 * VSM_DIM,
                           ctx->store.n2 + row0, n, ctx->d_store_stats));
        // a pageable source has been staged when cudaMemcpyAsync returns; only a pinned one is still
        // being read by the DMA engine, and the caller may reuse it right after this call
        if (kind == cudaMemcpyHostToDevice && host_pinned(src)) CK(cudaStreamSynchronize(ctx->stream));
    }
    const int32_t h = store_new_seg(ctx, row0, (int32_t)n, frame_id, is_kf);
    if (handle) *handle = h;
    return VSM_OK;
}

int vsm_store_add(vsm_ctx* ctx, int32_t frame_id, const float* desc, int32_t n, int32_t* handle) {
    return vsm_store_add_strided(ctx, frame_id, desc, n, VSM_DIM * sizeof(float), handle);
}

int vsm_store_add_strided(vsm_ctx* ctx, int32_t frame_id, const float* desc, int32_t n, int64_t stride_bytes, int32_t* handle) {
    if (!ctx || n < 0 || (n > 0 && !desc) || stride_bytes < (int64_t)(VSM_DIM * sizeof(float)) || (stride_bytes & 3))
        return ctx ? fail(ctx, VSM_ERR_INVALID, "vsm_store_add: bad argument (row stride must be >= 1024 bytes)") : VSM_ERR_INVALID;
    ctx->err.clear();
    CK(cudaSetDevice(ctx->device));
    return store_append(ctx, frame_id, desc, n, cudaMemcpyHostToDevice, true, handle, stride_bytes);
}

int vsm_store_add_device(vsm_ctx* ctx, int32_t frame_id, const float* d_desc, int64_t n, int32_t* handle) {
    if (!ctx || n < 0 || n > INT32_MAX || (n > 0 && !d_desc))
        return ctx ? fail(ctx, VSM_ERR_INVALID, "vsm_store_add_device: bad argument") : VSM_ERR_INVALID;
    ctx->err.clear();
    CK(cudaSetDevice(ctx->device));
    return store_append(ctx, frame_id, d_desc, n, cudaMemcpyDeviceToDevice, true, handle);
}

int vsm_store_remove(vsm_ctx* ctx, int32_t handle) {
    if (!ctx) return VSM_ERR_INVALID;
    ctx->err.clear();
    if (!seg_live(ctx, handle)) return fail(ctx, VSM_ERR_NOT_FOUND, "vsm_store_remove: unknown frame handle");
    if (!ctx->store.own_f32) return fail(ctx, VSM_ERR_INVALID, "store was adopted from a device matrix; clear it instead");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));                  // an asynchronous search may still read the rows
    store_remove_seg(ctx, handle);
    return VSM_OK;
}

int vsm_store_promote(vsm_ctx* ctx, int32_t handle) {
    if (!ctx) return VSM_ERR_INVALID;
    ctx->err.clear();
    if (!seg_live(ctx, handle)) return fail(ctx, VSM_ERR_NOT_FOUND, "vsm_store_promote: unknown frame handle");
    Seg& s = ctx->segs[handle];
    if (s.is_kf) return VSM_OK;                              // Frame::set_keyframe(true) twice is harmless
    ctx->plain_ring.erase(std::find(ctx->plain_ring.begin(), ctx->plain_ring.end(), handle));
    s.is_kf = 1;
    // Map::get_keyframes walks frames_ in insertion order (src/Map.cpp:40-47): a frame promoted late
    // (the bridge keyframe of src/Slam.cpp:851-863) still sits at its own position in that order
    auto it = std::lower_bound(ctx->kf_order.begin(), ctx->kf_order.end(), s.seq,
                               [&](int32_t h, int64_t seq) { return ctx->segs[h].seq < seq; });
    ctx->kf_order.insert(it, handle);
    ctx->kf_runs_valid = false;
    return VSM_OK;
}

int vsm_store_frame_info(const vsm_ctx* ctx, int32_t handle, int64_t* row0, int32_t* n_rows, int32_t* frame_id,
                         int32_t* is_keyframe) {
    if (!ctx) return VSM_ERR_INVALID;
    if (!seg_live(ctx, handle)) return VSM_ERR_NOT_FOUND;
    const Seg& s = ctx->segs[handle];
    if (row0) *row0 = s.row0;
    if (n_rows) *n_rows = s.count;
    if (frame_id) *frame_id = s.frame_id;
    if (is_keyframe) *is_keyframe = s.is_kf;
    return VSM_OK;
}

int vsm_store_keyframes(const vsm_ctx* ctx, int32_t* handles, int32_t cap, int32_t* n) {
    if (!ctx || !n || cap < 0 || (cap > 0 && !handles)) return VSM_ERR_INVALID;
    *n = (int32_t)ctx->kf_order.size();
    for (int32_t k = 0; k < *n && k < cap; k++) handles[k] = ctx->kf_order[k];
    return VSM_OK;
}

int vsm_store_load_spcf(vsm_ctx* ctx, const char* path, int32_t* n_loaded, int32_t* n_skipped, int32_t* first_handle) {
    if (!ctx || !path) return ctx ? fail(ctx, VSM_ERR_INVALID, "vsm_store_load_spcf: bad argument") : VSM_ERR_INVALID;
    ctx->err.clear();
    if (n_loaded) *n_loaded = 0;
    if (n_skipped) *n_skipped = 0;
    if (first_handle) *first_handle = -1;
    FILE* f = fopen(path, "rb");
    if (!f) return fail(ctx, VSM_ERR_NOT_FOUND, "vsm_store_load_spcf: cannot open file");
    std::vector<uint8_t> buf;
    {
        fseek(f, 0, SEEK_END);
        long sz = ftell(f);
        fseek(f, 0, SEEK_SET);
        buf.resize(sz > 0 ? (size_t)sz : 0);
        size_t got = buf.empty() ? 0 : fread(buf.data(), 1, buf.size(), f);
        fclose(f);
        if (got != buf.size()) return fail(ctx, VSM_ERR_INVALID, "vsm_store_load_spcf: short read");
    }
    size_t pos = 0;
    auto rd32 = [&](uint32_t& v) -> bool {
        if (pos + 4 > buf.size()) return false;
        memcpy(&v, buf.data() + pos, 4);
        pos += 4;
        return true;
    };
    uint32_t magic = 0, version = 0, n_entries = 0;
    if (!rd32(magic) || !rd32(version) || !rd32(n_entries) || magic != 0x53504346u || version != 1u)
        return fail(ctx, VSM_ERR_INVALID, "vsm_store_load_spcf: not an SPCF version-1 file");      // FeatureExtractor.cpp:281
    CK(cudaSetDevice(ctx->device));
    int loaded = 0, skipped = 0;
    for (uint32_t e = 0; e < n_entries; e++) {
        uint32_t frame_idx, num_kp, rows, cols, type;
        if (!rd32(frame_idx) || !rd32(num_kp)) return fail(ctx, VSM_ERR_INVALID, "vsm_store_load_spcf: truncated entry");
        pos += (size_t)num_kp * 28;                                     // x, y, size, angle, response, octave, class_id
        if (!rd32(rows) || !rd32(cols) || !rd32(type)) return fail(ctx, VSM_ERR_INVALID, "vsm_store_load_spcf: truncated entry");
        const int depth = type & 7;                                     // CV_MAT_DEPTH: 0 = 8U ... 5 = 32F, 6 = 64F
        const int channels = ((type >> 3) & 511) + 1;
        static const int depth_bytes[8] = {1, 1, 2, 2, 4, 4, 8, 2};
        size_t nbytes = 0;
        if ((int32_t)rows > 0 && (int32_t)cols > 0) nbytes = (size_t)rows * cols * depth_bytes[depth] * channels;
        if (pos + nbytes > buf.size()) return fail(ctx, VSM_ERR_INVALID, "vsm_store_load_spcf: truncated descriptors");
        if (type == 5 && cols == VSM_DIM && (int32_t)rows > 0) {        // CV_32FC1, N x 256
            // file offsets are only 4-byte aligned by construction, which is all the copy needs
            int32_t h = -1;
            TRY(store_append(ctx, (int32_t)frame_idx, reinterpret_cast<const float*>(buf.data() + pos), rows,
                             cudaMemcpyHostToDevice, true, &h));
            if (loaded == 0 && first_handle) *first_handle = h;
            loaded++;
        } else {
            skipped++;
        }
        pos += nbytes;
    }
    if (n_loaded) *n_loaded = loaded;
    if (n_skipped) *n_skipped = skipped;
    return VSM_OK;
}

static void store_reset_tables(vsm_ctx* ctx) {
    ctx->store_rows = 0;
    ctx->segs.clear();
    ctx->free_handles.clear();
    ctx->free_rows.clear();
    ctx->kf_order.clear();
    ctx->plain_ring.clear();
    ctx->kf_runs.clear();
    ctx->kf_runs_valid = false;
    ctx->next_seq = 0;
}

int vsm_store_clear(vsm_ctx* ctx) {
    if (!ctx) return VSM_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    if (!ctx->store.own_f32) arena_free(ctx, ctx->store);
    store_reset_tables(ctx);
    // on the context's own stream: ordered before the next conversion's atomicMax into the slot
    CK(cudaMemsetAsync(ctx->d_store_stats, 0, 16, ctx->stream));
    return VSM_OK;
}

int vsm_store_adopt_device(vsm_ctx* ctx, const float* d_desc, int64_t n_rows, const int64_t* seg_off, int32_t nseg) {
    if (!ctx || !d_desc || n_rows <= 0 || n_rows > INT32_MAX || (seg_off && nseg <= 0))
        return ctx ? fail(ctx, VSM_ERR_INVALID, "vsm_store_adopt_device: bad argument") : VSM_ERR_INVALID;
    if (seg_off) {                                           // validate before anything is changed
        if (seg_off[0] < 0) return fail(ctx, VSM_ERR_INVALID, "vsm_store_adopt_device: bad segment offsets");
        for (int s = 0; s < nseg; s++)
            if (seg_off[s + 1] < seg_off[s] || seg_off[s + 1] > n_rows)
                return fail(ctx, VSM_ERR_INVALID, "vsm_store_adopt_device: bad segment offsets");
    }
    TRY(vsm_store_clear(ctx));
    arena_free(ctx, ctx->store);
    ctx->store.own_f32 = false;
    ctx->store.f32 = const_cast<float*>(d_desc);
    int st = arena_reserve(ctx, ctx->store, n_rows, 0);
    if (st != VSM_OK) { ctx->store = Arena(); return st; }
    ctx->launches = 0;
    TRY(launch_convert(ctx, d_desc, ctx->store.b16, ctx->store.n2, n_rows, ctx->d_store_stats));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->store_rows = n_rows;
    if (seg_off) {
        for (int s = 0; s < nseg; s++) store_new_seg(ctx, seg_off[s], (int32_t)(seg_off[s + 1] - seg_off[s]), s, true);
    } else {
        store_new_seg(ctx, 0, (int32_t)n_rows, 0, true);
    }
    return VSM_OK;
}

int vsm_store_info(const vsm_ctx* ctx, int64_t* n_rows, int32_t* n_keyframes) {
    if (!ctx) return VSM_ERR_INVALID;
    if (n_rows) *n_rows = ctx->store_rows;
    if (n_keyframes) *n_keyframes = (int32_t)ctx->kf_order.size();
    return VSM_OK;
}

int vsm_match_to_stored(vsm_ctx* ctx, int32_t handle, const float* cur, int32_t n_cur, float ratio, int32_t mutual,
                        vsm_dmatch* good, int32_t* n_good, vsm_dmatch* raw, int32_t* n_raw) {
    if (!ctx || n_cur < 0 || !n_good || (n_cur > 0 && !cur))
        return ctx ? fail(ctx, VSM_ERR_INVALID, "vsm_match_to_stored: bad argument") : VSM_ERR_INVALID;
    if (!seg_live(ctx, handle)) return fail(ctx, VSM_ERR_NOT_FOUND, "unknown keyframe handle");
    const Seg sg = ctx->segs[handle];
    *n_good = 0;
    if (n_raw) *n_raw = 0;
    if (sg.count == 0 || n_cur == 0) return VSM_OK;
    if (!good) return fail(ctx, VSM_ERR_INVALID, "vsm_match_to_stored: null output");
    TRY(begin_call(ctx));
    TRY(arena_reserve(ctx, ctx->scratch, n_cur, 0));
    TRY(upload_scratch(ctx, cur, 0, n_cur));
    HProblem f;
    f.q_f32 = ctx->store.f32 + sg.row0 * VSM_DIM; f.q_n2 = ctx->store.n2 + sg.row0; f.q_row = sg.row0;
    f.q_store = 1; f.nq = sg.count;
    f.t_f32 = ctx->scratch.f32; f.t_row = 0; f.t_store = 0; f.nt = n_cur; f.out_off = 0;
    std::vector<HProblem> probs{f};
    if (mutual) {
        HProblem b;
        b.q_f32 = ctx->scratch.f32; b.q_n2 = ctx->scratch.n2; b.q_row = 0; b.q_store = 0; b.nq = n_cur;
        b.t_f32 = f.q_f32; b.t_row = sg.row0; b.t_store = 1; b.nt = sg.count; b.out_off = sg.count;
        probs.push_back(b);
    }
    return match_common(ctx, probs, sg.count, n_cur, ratio, mutual, good, n_good, raw, n_raw);
}

int vsm_match_batch_stored(vsm_ctx* ctx, int32_t n_pairs, const int32_t* q_handle, const int32_t* t_handle, float ratio,
                           int32_t mutual, vsm_dmatch* good, int64_t good_cap, int32_t* n_good, int64_t* good_off) {
    if (!ctx || n_pairs < 0 || (n_pairs > 0 && (!q_handle || !t_handle || !n_good || !good_off)))
        return ctx ? fail(ctx, VSM_ERR_INVALID, "vsm_match_batch_stored: bad argument") : VSM_ERR_INVALID;
    if (n_pairs == 0) return VSM_OK;
    int64_t NQ = 0, NT = 0;
    std::vector<int64_t> t_off(n_pairs + 1, 0);
    for (int p = 0; p < n_pairs; p++) {
        if (!seg_live(ctx, q_handle[p]) || !seg_live(ctx, t_handle[p]))
            return fail(ctx, VSM_ERR_NOT_FOUND, "vsm_match_batch_stored: unknown keyframe handle");
        good_off[p] = NQ;
        t_off[p] = NT;
        NQ += ctx->segs[q_handle[p]].count;
        NT += ctx->segs[t_handle[p]].count;
        n_good[p] = 0;
    }
    good_off[n_pairs] = NQ;
    t_off[n_pairs] = NT;
    if (NQ == 0 || (!good && good_cap == 0)) return VSM_OK;               // size query: good_off is filled
    if (!good || good_cap < NQ) return fail(ctx, VSM_ERR_INVALID, "vsm_match_batch_stored: good_cap is smaller than the query rows");
    TRY(begin_call(ctx));
    std::vector<HProblem> probs;
    std::vector<HJob> jobs;
    for (int p = 0; p < n_pairs; p++) {
        const Seg& a = ctx->segs[q_handle[p]];
        const Seg& b = ctx->segs[t_handle[p]];
        HProblem f;
        f.q_f32 = ctx->store.f32 + a.row0 * VSM_DIM; f.q_n2 = ctx->store.n2 + a.row0; f.q_row = a.row0; f.q_store = 1; f.nq = a.count;
        f.t_f32 = ctx->store.f32 + b.row0 * VSM_DIM; f.t_row = b.row0; f.t_store = 1; f.nt = b.count; f.out_off = good_off[p];
        f.skip_ratio2 = skip_r2(ratio);
        f.t2 = 1;
        f.ratio = ratio;
        probs.push_back(f);
        if (mutual) {
            HProblem r;
            r.t2 = 2;
            r.q_f32 = f.t_f32; r.q_n2 = ctx->store.n2 + b.row0; r.q_row = b.row0; r.q_store = 1; r.nq = b.count;
            r.t_f32 = f.q_f32; r.t_row = a.row0; r.t_store = 1; r.nt = a.count; r.out_off = NQ + t_off[p];
            probs.push_back(r);
        }
        HJob j;
        j.fwd_off = good_off[p]; j.back_off = mutual ? NQ + t_off[p] : -1; j.good_off = good_off[p]; j.raw_off = -1;
        j.nq = a.count; j.nt = b.count; j.img_idx = 0; j.ratio = ratio;
        j.back_prob = mutual ? (int64_t)probs.size() - 1 : -1;
        jobs.push_back(j);
    }
    TRY(run_problems(ctx, probs, jobs, NQ + (mutual ? NT : 0), NQ));
    const size_t bytes = (size_t)NQ * sizeof(DMatch) + (size_t)n_pairs * 2 * sizeof(int32_t);
    TRY(fetch_result(ctx, bytes));
    TRY(end_call(ctx, true));
    const int32_t* c = reinterpret_cast<const int32_t*>(ctx->h_result + (size_t)NQ * sizeof(DMatch));
    for (int p = 0; p < n_pairs; p++) {
        n_good[p] = c[2 * p];
        memcpy(good + good_off[p], ctx->h_result + (size_t)good_off[p] * sizeof(DMatch), (size_t)c[2 * p] * sizeof(DMatch));
    }
    return VSM_OK;
}

int vsm_track(vsm_ctx* ctx, int32_t ref_handle, int32_t frame_id, const float* cur, int32_t n_cur, float ratio,
              int32_t mutual, vsm_dmatch* good, int32_t* n_good, vsm_dmatch* raw, int32_t* n_raw, int32_t* cur_handle) {
    if (!ctx || n_cur < 0 || !n_good || (n_cur > 0 && !cur) || !cur_handle)
        return ctx ? fail(ctx, VSM_ERR_INVALID, "vsm_track: bad argument") : VSM_ERR_INVALID;
    if (ref_handle >= 0 && !seg_live(ctx, ref_handle)) return fail(ctx, VSM_ERR_NOT_FOUND, "unknown frame handle");
    if (!ctx->store.own_f32) return fail(ctx, VSM_ERR_INVALID, "store was adopted from a device matrix; clear it first");
    if (!good && ref_handle >= 0 && n_cur > 0 && ctx->segs[ref_handle].count > 0)
        return fail(ctx, VSM_ERR_INVALID, "vsm_track: null output");
    *n_good = 0;
    if (n_raw) *n_raw = 0;
    TRY(begin_call(ctx));
    // The frame enters the store as a PLAIN frame (Frame::is_keyframe_ = false, src/Frame.cpp:13); the
    // caller promotes it once the reference decides so (src/Slam.cpp:1065, :1076; vsm_store_promote).
    // Only last_frame_ and the frame being processed are ever matched again (src/Slam.cpp:838, :848),
    // so older plain frames are evicted here and their rows reused: tracking does not grow the store.
    while ((int)ctx->plain_ring.size() >= ctx->ring_depth) {
        int32_t victim = -1;
        for (int32_t h : ctx->plain_ring) if (h != ref_handle) { victim = h; break; }     // oldest first; never the reference
        if (victim < 0) break;
        store_remove_seg(ctx, victim);          // stream order protects the rows: every later use is enqueued after this call's
    }
    int64_t row0 = 0;
    TRY(store_alloc_rows(ctx, n_cur, &row0));
    if (n_cur > 0) {
        // a small pinned frame is read over PCIe by the call's prologue kernel (fp32 master + bf16
        // shadow + norms in one launch); otherwise one DMA, conversion in the prologue
        float* dst = ctx->store.f32 + row0 * VSM_DIM;
        const float* mapped = n_cur <= ZERO_COPY_ROWS ? host_mapped(cur) : nullptr;
        if (!mapped) CK(cudaMemcpyAsync(dst, cur, (size_t)n_cur * VSM_DIM * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        ConvJob j = {mapped ? mapped : dst, mapped ? dst : nullptr, ctx->store.b16 + row0 * VSM_DIM, ctx->store.n2 + row0,
                     ctx->d_store_stats, n_cur};
        ctx->pending_conv.push_back(j);
    }
    *cur_handle = store_new_seg(ctx, row0, n_cur, frame_id, false);
    const Seg ref = ref_handle >= 0 ? ctx->segs[ref_handle] : Seg{0, 0, 0, 0, 0, 0};
    if (ref_handle < 0 || ref.count == 0 || n_cur == 0) {
        TRY(flush_conversions(ctx, 0));                                  // no matching step: convert the frame on its own
        CK(cudaStreamSynchronize(ctx->stream));                          // the caller may reuse `cur`
        return end_call(ctx, true);
    }
    HProblem f;
    f.q_f32 = ctx->store.f32 + ref.row0 * VSM_DIM; f.q_n2 = ctx->store.n2 + ref.row0; f.q_row = ref.row0;
    f.q_store = 1; f.nq = ref.count;
    f.t_f32 = ctx->store.f32 + row0 * VSM_DIM; f.t_row = row0; f.t_store = 1; f.nt = n_cur; f.out_off = 0;
    std::vector<HProblem> probs{f};
    if (mutual) {
        HProblem b;
        b.q_f32 = f.t_f32; b.q_n2 = ctx->store.n2 + row0; b.q_row = row0; b.q_store = 1; b.nq = n_cur;
        b.t_f32 = f.q_f32; b.t_row = ref.row0; b.t_store = 1; b.nt = ref.count; b.out_off = ref.count;
        probs.push_back(b);
    }
    return match_common(ctx, probs, ref.count, n_cur, ratio, mutual, good, n_good, raw, n_raw);
}

// ---- database search -------------------------------------------------------------------
// The stacked-matrix search scans the rows of the live keyframes: one contiguous range while nothing
// has been removed and no plain frame sits between keyframes, otherwise the merged runs (the logical
// train index is the store row either way).
static int db_problem(vsm_ctx* ctx, const float* q_f32, int nq, HProblem& p) {
    p.q_f32 = q_f32; p.q_n2 = ctx->scratch.n2; p.q_row = 0; p.q_store = 0; p.nq = nq;
    p.t_f32 = ctx->store.f32; p.t_row = 0; p.t_store = 1; p.nt = (int)ctx->store_rows; p.out_off = 0;
    const std::vector<Run>& runs = store_runs(ctx);
    if (runs.empty()) p.nt = 0;
    else if (runs.size() == 1 && runs[0].row0 == 0) p.nt = (int)runs[0].count;
    else p.runs = &runs;
    return VSM_OK;
}

int vsm_db_top2(vsm_ctx* ctx, const float* query, int32_t nq, int64_t row_offset, int64_t* idx, float* dist) {
    if (!ctx || nq < 0 || (nq > 0 && (!query || !idx || !dist)))
        return ctx ? fail(ctx, VSM_ERR_INVALID, "vsm_db_top2: bad argument") : VSM_ERR_INVALID;
    if (nq == 0) return VSM_OK;
    TRY(begin_call(ctx));
    TRY(arena_reserve(ctx, ctx->scratch, nq, 0));
    TRY(upload_scratch(ctx, query, 0, nq));
    HProblem p;
    db_problem(ctx, ctx->scratch.f32, nq, p);
    TRY(run_problems(ctx, {p}, {}, nq, 0));
    TRY(fetch_keys(ctx, nq));
    TRY(end_call(ctx, true));
    const unsigned long long* k = reinterpret_cast<const unsigned long long*>(ctx->h_result);
    for (int i = 0; i < nq * 2; i++) {
        decode_key(k[i], idx[i], dist[i]);
        if (idx[i] >= 0) idx[i] += row_offset;
    }
    return VSM_OK;
}

int vsm_db_top2_device(vsm_ctx* ctx, const float* d_query, int32_t nq, int64_t row_offset, int64_t* d_idx,
                       float* d_dist, int32_t sync) {
    if (!ctx || nq < 0 || (nq > 0 && (!d_query || !d_idx || !d_dist)))
        return ctx ? fail(ctx, VSM_ERR_INVALID, "vsm_db_top2_device: bad argument") : VSM_ERR_INVALID;
    if (nq == 0) return VSM_OK;
    TRY(begin_call(ctx));
    TRY(arena_reserve(ctx, ctx->scratch, nq, 0));
    HProblem p;
    db_problem(ctx, d_query, nq, p);
    static const int timeline = getenv("VSM_DEBUG_TIMELINE") ? std::max(2, atoi(getenv("VSM_DEBUG_TIMELINE"))) : 0;
    queue_convert(ctx, d_query, 0, nq);
    TRY(run_problems(ctx, {p}, {}, nq, 0, timeline));
    widen_kernel<<<(nq * 2 + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_out_key, nq * 2, row_offset, d_idx, d_dist);
    ctx->launches++;
    CK(cudaGetLastError());
    return end_call(ctx, sync != 0);
}

// Per-keyframe top-2 + ratio test for the keyframes with eligible[k] != 0 (all if NULL); k = position
// in Map::get_keyframes() order (ctx->kf_order).
static int segmented_impl(vsm_ctx* ctx, const float* query, int32_t nq, float ratio, const std::vector<char>* eligible,
                          int32_t* counts, vsm_dmatch* matches) {
    const int nkf = (int)ctx->kf_order.size();
    TRY(begin_call(ctx));
    TRY(arena_reserve(ctx, ctx->scratch, nq, 0));
    TRY(upload_scratch(ctx, query, 0, nq));
    std::vector<HProblem> probs;
    std::vector<HJob> jobs;
    std::vector<int> job_seg;
    for (int s = 0; s < nkf; s++) {
        if (eligible && !(*eligible)[s]) continue;
        const Seg& sg = ctx->segs[ctx->kf_order[s]];
        const int64_t slot = (int64_t)jobs.size();
        HProblem p;
        p.q_f32 = ctx->scratch.f32; p.q_n2 = ctx->scratch.n2; p.q_row = 0; p.q_store = 0; p.nq = nq;
        p.t_f32 = ctx->store.f32 + sg.row0 * VSM_DIM; p.t_row = sg.row0; p.t_store = 1; p.nt = sg.count;
        p.out_off = slot * nq;
        p.skip_ratio2 = skip_r2(ratio);                 // only ratio-test survivors are returned
        // Maxima-only records pay off while matches are rare: a query the ratio test cannot dismiss costs
        // an exact scan of the whole keyframe instead of ~4 re-scores.  The share of such (query, keyframe)
        // pairs is measured by every per-keyframe search (either epilogue); the next search uses the
        // maxima-only epilogue only if it was below 0.1 % (no loop in sight: the usual case).
        p.maxima_only = ctx->seg_open_rate < 1e-3f ? 1 : 0;
        // Keyframes of up to 8192 rows (every keyframe of the reference: SP_MAX_KEYPOINTS = 400, include/Config.h:42)
        // keep tile top-2 records instead: their cost does not depend on the number of matches either, their
        // select pass is per-lane arithmetic plus one exact distance per match, and that beats both record kinds
        // above (500 keyframes x 1000 rows x 1000 queries: tensor-core pass 0.25 ms + select 0.09 ms, against
        // 0.20 + 0.22 ms maxima-only and 0.44 + 0.18 ms top-4)
        if (!ctx->t2_off && ratio > 0.f && ratio <= 1.f && (sg.count + TILE_N - 1) / TILE_N <= T2_MAX_TILES &&
            (sg.count + TILE_N - 1) / TILE_N > ctx->append_max_tiles) {
            p.maxima_only = 0;
            p.t2 = 1;
            p.ratio = ratio;
        }
        probs.push_back(p);
        HJob j;
        j.fwd_off = p.out_off; j.back_off = -1; j.good_off = slot * nq; j.raw_off = -1;
        j.nq = nq; j.nt = sg.count; j.img_idx = s; j.ratio = ratio;
        jobs.push_back(j);
        job_seg.push_back(s);
    }
    if (jobs.empty()) return end_call(ctx, true);
    const int64_t total_matches = (int64_t)jobs.size() * nq;
    TRY(run_problems(ctx, probs, jobs, total_matches, total_matches));
    const size_t mbytes = (size_t)total_matches * sizeof(DMatch);
    const size_t cbytes = jobs.size() * 2 * sizeof(int32_t);
    // one copy of the whole result block (lists + counts) into the pinned buffer, scattered on the host
    if (!ctx->result_on_host) TRY(fetch_result(ctx, matches ? mbytes + cbytes : 0));
    std::vector<int32_t> cnt(jobs.size() * 2);
    if (!ctx->result_on_host && !matches)
        CK(cudaMemcpyAsync(cnt.data(), ctx->d_result.p + mbytes, cbytes, cudaMemcpyDeviceToHost, ctx->stream));
    TRY(end_call(ctx, true));
    if (ctx->result_on_host || matches) memcpy(cnt.data(), ctx->h_result + mbytes, cbytes);
    if (matches)
        for (size_t k = 0; k < jobs.size(); k++)       // slot k -> keyframe job_seg[k]
            memcpy(matches + (size_t)job_seg[k] * nq, ctx->h_result + k * (size_t)nq * sizeof(DMatch),
                   (size_t)cnt[2 * k] * sizeof(DMatch));
    for (size_t k = 0; k < jobs.size(); k++) counts[job_seg[k]] = cnt[2 * k];
    uint32_t open_pairs = 0;                                      // stream is idle: a 4-byte read
    CK(cudaMemcpy(&open_pairs, reinterpret_cast<const uint8_t*>(ctx->d_counters) + 20, sizeof open_pairs, cudaMemcpyDeviceToHost));
    ctx->seg_open_rate = (float)open_pairs / (float)std::max<int64_t>(total_matches, 1);
    return VSM_OK;
}

// ---- resident map-point table ----------------------------------------------------------------------
static int points_append(vsm_ctx* ctx, int32_t n, int32_t frame_id, int32_t* first_id) {
    // room for n more points and n more log entries; validity = 1 (MapPoint's constructor, src/MapPoint.cpp:8-13)
    TRY(grow_arr(ctx, ctx->pt_f32, (size_t)(ctx->n_points + n) * VSM_DIM * sizeof(float), (size_t)1 << 38));
    TRY(grow_arr(ctx, ctx->pt_valid, (size_t)(ctx->n_points + n), (size_t)1 << 30));
    TRY(grow_arr(ctx, ctx->pt_log, (size_t)(ctx->n_log + n) * sizeof(PointObs), (size_t)1 << 34));
    std::vector<PointObs> obs(n);
    for (int i = 0; i < n; i++) obs[i] = PointObs{(int32_t)(ctx->n_points + i), frame_id};
    CK(cudaMemsetAsync(ctx->pt_valid.p + ctx->n_points, 1, (size_t)n, ctx->stream));
    CK(cudaMemcpyAsync(ctx->pt_log.p + (size_t)ctx->n_log * sizeof(PointObs), obs.data(), (size_t)n * sizeof(PointObs),
                       cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));                   // `obs` is a local; the caller may reuse its buffers
    if (first_id) *first_id = (int32_t)ctx->n_points;
    ctx->pt_valid_h.resize((size_t)(ctx->n_points + n), 1);
    ctx->n_points += n;
    ctx->n_log += n;
    ctx->n_valid += n;
    return VSM_OK;
}

int vsm_points_add(vsm_ctx* ctx, const float* desc, int32_t n, int32_t frame_id, int32_t* first_id) {
    if (!ctx || n < 0 || (n > 0 && !desc)) return ctx ? fail(ctx, VSM_ERR_INVALID, "vsm_points_add: bad argument") : VSM_ERR_INVALID;
    ctx->err.clear();
    if (first_id) *first_id = (int32_t)ctx->n_points;
    if (n == 0) return VSM_OK;
    CK(cudaSetDevice(ctx->device));
    TRY(grow_arr(ctx, ctx->pt_f32, (size_t)(ctx->n_points + n) * VSM_DIM * sizeof(float), (size_t)1 << 38));
    CK(cudaMemcpyAsync(ctx->pt_f32.p + (size_t)ctx->n_points * VSM_DIM * sizeof(float), desc, (size_t)n * VSM_DIM * sizeof(float),
                       cudaMemcpyHostToDevice, ctx->stream));
    return points_append(ctx, n, frame_id, first_id);
}

int vsm_points_add_from_frame(vsm_ctx* ctx, int32_t handle, const int32_t* kp_idx, int32_t n, int32_t* first_id) {
    if (!ctx || n < 0 || (n > 0 && !kp_idx)) return ctx ? fail(ctx, VSM_ERR_INVALID, "vsm_points_add_from_frame: bad argument") : VSM_ERR_INVALID;
    ctx->err.clear();
    if (!seg_live(ctx, handle)) return fail(ctx, VSM_ERR_NOT_FOUND, "vsm_points_add_from_frame: unknown frame handle");
    const Seg sg = ctx->segs[handle];
    for (int i = 0; i < n; i++)
        if (kp_idx[i] < 0 || kp_idx[i] >= sg.count) return fail(ctx, VSM_ERR_INVALID, "vsm_points_add_from_frame: keypoint index outside the frame");
    if (first_id) *first_id = (int32_t)ctx->n_points;
    if (n == 0) return VSM_OK;
    CK(cudaSetDevice(ctx->device));
    TRY(grow_arr(ctx, ctx->pt_f32, (size_t)(ctx->n_points + n) * VSM_DIM * sizeof(float), (size_t)1 << 38));
    // descriptor = frame->descriptors().row(kp).clone() (src/Slam.cpp:1339, :1563): a device-side row gather
    std::vector<int32_t> rows(n);
    for (int i = 0; i < n; i++) rows[i] = (int32_t)(sg.row0 + kp_idx[i]);
    TRY(ensure(ctx, ctx->d_sel, (size_t)n));
    CK(cudaMemcpyAsync(ctx->d_sel.p, rows.data(), (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    const int64_t blocks = std::min<int64_t>((n + 7) / 8, (int64_t)ctx->num_sms * 16);
    gather_rows_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(ctx->store.f32, ctx->d_sel.p, n,
                                                                  reinterpret_cast<float*>(ctx->pt_f32.p) + (size_t)ctx->n_points * VSM_DIM);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));                   // `rows` is a local
    return points_append(ctx, n, sg.frame_id, first_id);
}

int vsm_points_observe(vsm_ctx* ctx, const int32_t* point_ids, int32_t n, int32_t frame_id) {
    if (!ctx || n < 0 || (n > 0 && !point_ids)) return ctx ? fail(ctx, VSM_ERR_INVALID, "vsm_points_observe: bad argument") : VSM_ERR_INVALID;
    ctx->err.clear();
    if (n == 0) return VSM_OK;
    for (int i = 0; i < n; i++)
        if (point_ids[i] < 0 || point_ids[i] >= ctx->n_points) return fail(ctx, VSM_ERR_NOT_FOUND, "vsm_points_observe: unknown map point");
    CK(cudaSetDevice(ctx->device));
    TRY(grow_arr(ctx, ctx->pt_log, (size_t)(ctx->n_log + n) * sizeof(PointObs), (size_t)1 << 34));
    std::vector<PointObs> obs(n);
    for (int i = 0; i < n; i++) obs[i] = PointObs{point_ids[i], frame_id};
    CK(cudaMemcpyAsync(ctx->pt_log.p + (size_t)ctx->n_log * sizeof(PointObs), obs.data(), (size_t)n * sizeof(PointObs),
                       cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->n_log += n;
    return VSM_OK;
}

int vsm_points_set_valid(vsm_ctx* ctx, const int32_t* point_ids, int32_t n, int32_t valid) {
    if (!ctx || n < 0 || (n > 0 && !point_ids)) return ctx ? fail(ctx, VSM_ERR_INVALID, "vsm_points_set_valid: bad argument") : VSM_ERR_INVALID;
    ctx->err.clear();
    if (n == 0) return VSM_OK;
    for (int i = 0; i < n; i++)
        if (point_ids[i] < 0 || point_ids[i] >= ctx->n_points) return fail(ctx, VSM_ERR_NOT_FOUND, "vsm_points_set_valid: unknown map point");
    CK(cudaSetDevice(ctx->device));
    const uint8_t f = valid ? 1 : 0;
    for (int i = 0; i < n; i++) {
        uint8_t& h = ctx->pt_valid_h[(size_t)point_ids[i]];
        ctx->n_valid += (int64_t)f - (int64_t)h;
        h = f;
    }
    TRY(ensure(ctx, ctx->d_sel, (size_t)n));
    CK(cudaMemcpyAsync(ctx->d_sel.p, point_ids, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    points_set_valid_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_sel.p, n, f, ctx->pt_valid.p);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));                   // the caller may reuse point_ids
    return VSM_OK;
}

int vsm_points_info(const vsm_ctx* ctx, int64_t* n_points, int64_t* n_valid, int64_t* n_observations) {
    if (!ctx) return VSM_ERR_INVALID;
    if (n_points) *n_points = ctx->n_points;
    if (n_valid) *n_valid = ctx->n_valid;
    if (n_observations) *n_observations = ctx->n_log;
    return VSM_OK;
}

int vsm_points_clear(vsm_ctx* ctx) {
    if (!ctx) return VSM_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->n_points = ctx->n_log = ctx->n_valid = 0;
    ctx->pt_valid_h.clear();
    return VSM_OK;
}

// knnMatch(query, stack of the rows i of src_f32 with d_valid[i] != 0 (and d_near[i] != 0 if given), 2) with the
// selection made on the device: count / scan / scatter the selected row numbers in ascending order (the order
// of the reference's re-stacking loops, src/Slam.cpp:552-557, :744-759), gather the rows, run the ordinary
// search, map trainIdx back to row numbers.  d_tmp: scratch of at least select_tmp_bytes(n) bytes whose first n
// bytes may hold d_near.  One 4-byte read (the selection size) is the only host round trip.
static size_t select_tmp_bytes(int64_t n) {
    const size_t nblocks = (size_t)((n + POINTS_PER_BLOCK - 1) / POINTS_PER_BLOCK);
    return align16((size_t)n) * 2 + align16(nblocks * 4) * 2 + 32;
}
static int select_and_search(vsm_ctx* ctx, const float* query, int32_t nq, const uint8_t* d_valid, const uint8_t* d_near,
                             int64_t n, const float* src_f32, uint8_t* d_tmp, int64_t* idx, float* dist, int32_t* n_selected) {
    const int nblocks = (int)((n + POINTS_PER_BLOCK - 1) / POINTS_PER_BLOCK);
    const size_t o_flag = align16((size_t)n), o_cnt = o_flag + align16((size_t)n), o_off = o_cnt + align16((size_t)nblocks * 4),
                 o_total = o_off + align16((size_t)nblocks * 4);
    TRY(ensure(ctx, ctx->d_sel, (size_t)n));
    points_count_kernel<<<nblocks, 256, 0, ctx->stream>>>(d_valid, d_near, n, d_tmp + o_flag, reinterpret_cast<int32_t*>(d_tmp + o_cnt));
    points_scan_kernel<<<1, 1024, 0, ctx->stream>>>(reinterpret_cast<const int32_t*>(d_tmp + o_cnt), nblocks,
                                                    reinterpret_cast<int32_t*>(d_tmp + o_off), reinterpret_cast<int32_t*>(d_tmp + o_total));
    points_scatter_kernel<<<nblocks, 256, 0, ctx->stream>>>(d_tmp + o_flag, n, reinterpret_cast<const int32_t*>(d_tmp + o_off), ctx->d_sel.p);
    ctx->launches += 3;
    CK(cudaGetLastError());
    int32_t ns = 0;                                              // the search is planned on the host: one 4-byte read
    CK(cudaMemcpyAsync(&ns, d_tmp + o_total, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (n_selected) *n_selected = ns;
    if (nq == 0 || ns == 0) return end_call(ctx, true);
    // the stacked matrix the reference builds by push_back (:556, :757), gathered on the device; then the ordinary search
    TRY(arena_reserve(ctx, ctx->scratch, (int64_t)nq + ns, 0));
    TRY(upload_scratch(ctx, query, 0, nq));
    const int64_t blocks = std::min<int64_t>(((int64_t)ns + 7) / 8, (int64_t)ctx->num_sms * 16);
    gather_rows_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(src_f32, ctx->d_sel.p, ns, ctx->scratch.f32 + (int64_t)nq * VSM_DIM);
    ctx->launches++;
    CK(cudaGetLastError());
    std::vector<HProblem> probs{scratch_vs_scratch(ctx, 0, nq, nq, ns, 0)};
    queue_convert(ctx, ctx->scratch.f32 + (int64_t)nq * VSM_DIM, nq, ns);
    TRY(run_problems(ctx, probs, {}, nq, 0));
    // trainIdx -> row number (the reference's mp_ids_vec[m[0].trainIdx], :768), straight into pinned host memory
    const size_t nb = (size_t)nq * 2 * (sizeof(int64_t) + sizeof(float));
    TRY(ensure_host(ctx, ctx->h_result, ctx->h_result_cap, std::max<size_t>(nb, 16)));
    int64_t* h_idx = reinterpret_cast<int64_t*>(ctx->h_result);
    float* h_dist = reinterpret_cast<float*>(ctx->h_result + (size_t)nq * 2 * sizeof(int64_t));
    points_result_kernel<<<(nq * 2 + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_out_key, nq * 2, ctx->d_sel.p, h_idx, h_dist);
    ctx->launches++;
    CK(cudaGetLastError());
    TRY(end_call(ctx, true));
    memcpy(idx, h_idx, (size_t)nq * 2 * sizeof(int64_t));
    memcpy(dist, h_dist, (size_t)nq * 2 * sizeof(float));
    return VSM_OK;
}

int vsm_points_top2(vsm_ctx* ctx, const float* query, int32_t nq, int32_t near_frame_id, int32_t range, int64_t* idx, float* dist,
                    int32_t* n_selected) {
    if (!ctx || nq < 0 || (nq > 0 && (!query || !idx || !dist)) || (near_frame_id >= 0 && range <= 0))
        return ctx ? fail(ctx, VSM_ERR_INVALID, "vsm_points_top2: bad argument") : VSM_ERR_INVALID;
    if (n_selected) *n_selected = 0;
    for (int i = 0; i < nq * 2; i++) { idx[i] = -1; dist[i] = FLT_MAX; }
    const int64_t np = ctx->n_points;
    if (np == 0) return VSM_OK;
    TRY(begin_call(ctx));
    // valid (:553 / :747) and, for the loop-verification search, seen near the matched keyframe (:748-756)
    TRY(ensure(ctx, ctx->d_pt_tmp, select_tmp_bytes(np)));
    uint8_t* t = ctx->d_pt_tmp.p;
    const uint8_t* near = nullptr;
    if (near_frame_id >= 0) {
        CK(cudaMemsetAsync(t, 0, (size_t)np, ctx->stream));
        if (ctx->n_log > 0)
            points_mark_near_kernel<<<(unsigned)((ctx->n_log + 255) / 256), 256, 0, ctx->stream>>>(
                reinterpret_cast<const PointObs*>(ctx->pt_log.p), ctx->n_log, near_frame_id, range, t);
        near = t;
        ctx->launches++;
    }
    return select_and_search(ctx, query, nq, ctx->pt_valid.p, near, np, reinterpret_cast<const float*>(ctx->pt_f32.p), t, idx, dist,
                             n_selected);
}

// ---- LoopCloser::detect, compact form ------------------------------------------------------------
constexpr uint32_t PAIR_CAP = 1u << 18;          // open (query, keyframe) pairs per search: 4 MB each of keys, references, staged matches
constexpr uint32_t UNIT2_CAP = 8192;             // (keyframe, query tile) units of the second pass: 32 MB of records

// The eligible keyframes (positions in get_keyframes() order) against the query frame: fused ratio
// dismissal in the tensor-core epilogue, exact scans for the open pairs only, gate and packed lists on
// the device.  *overflow: the open pairs did not fit (a scene full of matches): nothing was returned and
// the caller takes the record-based path.
static int loop_compact_impl(vsm_ctx* ctx, const float* query, int32_t nq, float ratio, int32_t min_matches,
                             const std::vector<int32_t>& elig_pos, int32_t* status, vsm_loop_candidate* cands,
                             int32_t cand_cap, int32_t* n_cands, vsm_dmatch* matches, int64_t match_cap,
                             int64_t* n_matches, bool* overflow) {
    *overflow = false;
    const int nslots = (int)elig_pos.size();
    const int nqt = (nq + TILE_M - 1) / TILE_M;
    const int wps = nqt * 4;
    const uint32_t pair_cap = ctx->pair_cap ? ctx->pair_cap : PAIR_CAP;
    TRY(begin_call(ctx));
    TRY(arena_reserve(ctx, ctx->scratch, nq, 0));
    TRY(upload_scratch(ctx, query, 0, nq));

    // device layout
    const uint32_t unit2_cap = UNIT2_CAP;
    const size_t nwords = (size_t)nslots * wps;
    const size_t o_mask = 0, o_base = align16(o_mask + nwords * 4), o_keys = align16(o_base + nwords * 4),
                 o_ref = align16(o_keys + (size_t)pair_cap * 16), o_stage = align16(o_ref + (size_t)pair_cap * sizeof(PairRef)),
                 o_off = align16(o_stage + (size_t)pair_cap * sizeof(DMatch)), o_unit2 = align16(o_off + (size_t)nslots * 8),
                 o_hint2 = align16(o_unit2 + (size_t)unit2_cap * sizeof(TcUnit)),
                 loop_bytes = align16(o_hint2 + (size_t)unit2_cap * TILE_M * 4);
    TRY(ensure(ctx, ctx->d_loop, loop_bytes));
    TRY(ensure(ctx, ctx->d_recs, (size_t)unit2_cap * TILE_M * 2));
    const size_t off_slot = 0, off_unit = align16(off_slot + sizeof(LoopSlot) * nslots),
                 off_fused = align16(off_unit + sizeof(TcUnit) * (size_t)nslots * nqt), total = align16(off_fused + sizeof(FusedArgs));
    TRY(ensure(ctx, ctx->d_desc, total));
    TRY(ensure_host(ctx, ctx->h_desc, ctx->h_desc_cap, total));
    TRY(ensure(ctx, ctx->d_work, (size_t)WORK_CAP));
    // header (incl. RedoCtl at bytes 48..63) + survivors per slot + ready flags of the redo units, zeroed per call
    const size_t o_ready = align16(64 + (size_t)nslots * 4);
    const size_t aux_bytes = align16(o_ready + (size_t)unit2_cap * 4);
    TRY(ensure(ctx, ctx->d_aux, aux_bytes));
    if (ctx->aux_zeroed != ctx->d_aux.p) {
        CK(cudaMemsetAsync(ctx->d_aux.p, 0, ctx->d_aux.cap, ctx->stream));
        ctx->aux_zeroed = ctx->d_aux.p;
    }
    // outputs in pinned host memory, written by the last kernel: head | survivors per slot | candidates | matches
    const size_t h_head = 0, h_good = 16, h_cand = align16(h_good + (size_t)nslots * 4),
                 h_match = align16(h_cand + (size_t)nslots * sizeof(LoopCand)),
                 h_total = h_match + (size_t)pair_cap * sizeof(DMatch);
    TRY(ensure_host(ctx, ctx->h_result, ctx->h_result_cap, h_total));
    ctx->plan.valid = false;                                               // the descriptor block is overwritten

    // descriptor block: rebuilt only when the eligible list (or a buffer) changed since the last search
    std::vector<uint8_t>& key = ctx->plan_key_build;
    {
        const int64_t head[6] = {nq, nslots, (int64_t)__float_as_uint_host(ratio) | ((int64_t)pair_cap << 32), (int64_t)(uintptr_t)ctx->store.f32,
                                 (int64_t)(uintptr_t)ctx->scratch.n2 ^ (int64_t)(uintptr_t)ctx->d_aux.p,
                                 (int64_t)(uintptr_t)ctx->d_store_stats ^ ((int64_t)ctx->store.cap << 20)};
        key.resize(sizeof head + (size_t)nslots * sizeof(LoopSlot));
        memcpy(key.data(), head, sizeof head);
    }
    std::vector<LoopSlot> slots(nslots);
    for (int s = 0; s < nslots; s++) {
        const Seg& sg = ctx->segs[ctx->kf_order[elig_pos[s]]];
        slots[s] = LoopSlot{sg.row0, sg.count, elig_pos[s]};
    }
    if (nslots) memcpy(key.data() + 48, slots.data(), (size_t)nslots * sizeof(LoopSlot));
    const bool hit = ctx->loop_p_desc == ctx->d_desc.p && ctx->loop_p_loop == ctx->d_loop.p && ctx->loop_key == key;
    uint8_t* dl = ctx->d_loop.p;
    if (!hit) {
        std::vector<uint8_t>& blk = ctx->desc_build;
        blk.assign(total, 0);
        memcpy(blk.data() + off_slot, slots.data(), sizeof(LoopSlot) * nslots);
        TcUnit* units = reinterpret_cast<TcUnit*>(blk.data() + off_unit);
        const float r2 = skip_r2(ratio);
        for (int s = 0; s < nslots; s++)
            for (int qt = 0; qt < nqt; qt++) {
                TcUnit& u = units[(size_t)s * nqt + qt];
                u.q_n2 = ctx->scratch.n2 + (int64_t)qt * TILE_M;
                u.t_stats = ctx->d_store_stats;
                u.rec_base = ((int64_t)s * nqt + qt) * 4;                    // first mask word of the unit
                u.rec_stride = 0;
                u.q_row = qt * TILE_M;
                u.t_row = (int32_t)slots[s].row0;
                u.t_count = slots[s].count;
                u.t_index0 = s;                                             // fused units: the keyframe's slot (PairRef::slot)
                u.q_valid = std::min(TILE_M, nq - qt * TILE_M);
                u.seg_tiles = UNIT_TILES;
                u.maps = 2 | 4 | 8;                                         // train rows in the store; maxima only; fused dismissal
                u.prefetch = (qt == 0 || qt == nqt / 2) ? 1 : 0;
                u.skip_ratio2 = r2;
                u.hint = reinterpret_cast<uint32_t*>(dl + o_mask);
            }
        FusedArgs* fa = reinterpret_cast<FusedArgs*>(blk.data() + off_fused);
        fa->ctl = reinterpret_cast<RedoCtl*>(ctx->d_aux.p + 48);
        fa->units2 = reinterpret_cast<TcUnit*>(dl + o_unit2);
        fa->ready2 = reinterpret_cast<uint32_t*>(ctx->d_aux.p + o_ready);
        fa->hints2 = reinterpret_cast<uint32_t*>(dl + o_hint2);
        fa->word_base = reinterpret_cast<uint32_t*>(dl + o_base);
        fa->pair_ref = reinterpret_cast<PairRef*>(dl + o_ref);
        fa->counters = reinterpret_cast<uint32_t*>(ctx->d_aux.p);
        fa->unit2_cap = unit2_cap;
        fa->pair_cap = pair_cap;
        fa->n_main = (uint32_t)((size_t)nslots * nqt);
        if (ctx->desc_copy_pending) CK(cudaEventSynchronize(ctx->ev_desc));
        memcpy(ctx->h_desc, blk.data(), total);
        ctx->loop_key = key;
        ctx->loop_p_desc = ctx->d_desc.p;
        ctx->loop_p_loop = ctx->d_loop.p;
    }
    TRY(launch_prologue(ctx, aux_bytes, total, !hit));
    uint8_t* dd = ctx->d_desc.p;
    const size_t nunits = (size_t)nslots * nqt;
    if (ctx->profiling) {
        const uint32_t slot = ctx->tc_ring_head++ % vsm_ctx::TC_RING;
        ctx->ev_tc0 = ctx->tc_ring0[slot];
        ctx->ev_tc1 = ctx->tc_ring1[slot];
        ctx->tc_ring_valid[slot] = nunits > 0;
        CK(cudaEventRecord(ctx->ev_tc0, ctx->stream));
    }
    {
        const unsigned grid = (unsigned)std::min<size_t>(nunits, (size_t)ctx->num_sms);
        uint32_t* d_unit_counter = reinterpret_cast<uint32_t*>(ctx->d_aux.p + 24);
        CK(launch_pdl(tc::tc_top3_kernel<false>, dim3(grid), dim3(tc::THREADS), tc::SMEM_BYTES, ctx->stream, ctx->scratch.map,
                      ctx->store.map, reinterpret_cast<const TcUnit*>(dd + off_unit), (int)nunits,
                      reinterpret_cast<const FusedArgs*>(dd + off_fused), d_unit_counter, ctx->d_recs.p, ctx->d_dump));
        ctx->launches++;
        if (ctx->profiling) {
            CK(cudaEventRecord(ctx->ev_tc1, ctx->stream));
            ctx->timed_tc = true;
        }
    }
    LoopParams P;
    memset(&P, 0, sizeof P);
    P.slots = reinterpret_cast<const LoopSlot*>(dd + off_slot);
    P.nslots = nslots; P.nq = nq; P.words_per_slot = wps;
    P.q_f32 = ctx->scratch.f32;
    P.q_n2 = ctx->scratch.n2;
    P.store_f32 = ctx->store.f32;
    P.t_stats = ctx->d_store_stats;
    P.masks = reinterpret_cast<const uint32_t*>(dl + o_mask);
    P.word_base = reinterpret_cast<uint32_t*>(dl + o_base);
    P.pair_keys = reinterpret_cast<unsigned long long*>(dl + o_keys);
    P.pair_ref = reinterpret_cast<PairRef*>(dl + o_ref);
    P.stage = reinterpret_cast<DMatch*>(dl + o_stage);
    P.pair_cap = pair_cap;
    P.units2 = reinterpret_cast<TcUnit*>(dl + o_unit2);
    P.hints2 = reinterpret_cast<uint32_t*>(dl + o_hint2);
    P.recs2 = ctx->d_recs.p;
    P.unit2_cap = unit2_cap;
    P.counters = reinterpret_cast<uint32_t*>(ctx->d_aux.p);
    P.work = ctx->d_work.p;
    P.work_cap = ctx->work_cap;
    P.slot_good = reinterpret_cast<int32_t*>(ctx->d_aux.p + 64);          // (the ready flags of the redo units follow)
    P.slot_off = reinterpret_cast<int64_t*>(dl + o_off);
    P.ratio = ratio;
    P.skip_ratio2 = skip_r2(ratio);
    P.min_matches = min_matches;
    P.out_head = reinterpret_cast<int32_t*>(ctx->h_result + h_head);
    P.out_good = reinterpret_cast<int32_t*>(ctx->h_result + h_good);
    P.out_cands = reinterpret_cast<LoopCand*>(ctx->h_result + h_cand);
    P.cand_cap = nslots;
    P.out_matches = reinterpret_cast<DMatch*>(ctx->h_result + h_match);
    P.match_cap = pair_cap;
    const unsigned wblocks = (unsigned)std::max<size_t>(1, std::min<size_t>((nwords + 7) / 8, (size_t)ctx->num_sms * 8));
    ctx->d_counters = reinterpret_cast<unsigned long long*>(ctx->d_aux.p);
    CK(launch_pdl(loop_select_kernel, dim3((unsigned)ctx->num_sms * 4), dim3(SELECT_WARPS * 32), 0, ctx->stream, P));
    CK(launch_pdl(rescan_kernel, dim3((unsigned)ctx->num_sms * 2), dim3(256), 0, ctx->stream, (const WorkItem*)ctx->d_work.p,
                  (const unsigned long long*)ctx->d_counters, ctx->work_cap));
    CK(launch_pdl(loop_finish_kernel, dim3(wblocks), dim3(256), 0, ctx->stream, P));       // its last block gates and emits
    ctx->launches += 3;
    if (ctx->profiling) {
        CK(cudaEventRecord(ctx->ev_sel1, ctx->stream));
        ctx->timed_sel = true;
    }
    TRY(end_call(ctx, true));                                          // every output was written into pinned host memory
    const int32_t* head = reinterpret_cast<const int32_t*>(ctx->h_result + h_head);
    ctx->seg_open_rate = (float)(uint32_t)head[3] / (float)std::max<int64_t>((int64_t)nslots * nq, 1);
    if (head[2]) { *overflow = true; return VSM_OK; }
    const int32_t* good = reinterpret_cast<const int32_t*>(ctx->h_result + h_good);
    for (int s = 0; s < nslots; s++) status[elig_pos[s]] = good[s];
    const int nc = head[0];
    const int64_t nm = head[1];
    if (n_cands) *n_cands = nc;
    if (n_matches) *n_matches = nm;
    if (cands) memcpy(cands, ctx->h_result + h_cand, (size_t)std::min(nc, cand_cap) * sizeof(LoopCand));
    if (matches) memcpy(matches, ctx->h_result + h_match, (size_t)std::min<int64_t>(nm, match_cap) * sizeof(DMatch));
    return VSM_OK;
}

static int loop_eligible(vsm_ctx* ctx, int32_t cur_frame_id, int32_t min_gap, int32_t every, int32_t checked_before,
                         std::vector<char>& eligible, int32_t* status, int32_t* checked_after) {
    const int nseg = (int)ctx->kf_order.size();                       // keyframes only, in Map::get_keyframes() order
    eligible.assign(nseg, 0);
    int checked = checked_before, any = 0;
    for (int s = 0; s < nseg; s++) {                                  // src/LoopCloser.cpp:43-48
        const Seg& sg = ctx->segs[ctx->kf_order[s]];
        status[s] = -1;
        if (cur_frame_id - sg.frame_id < min_gap) continue;
        if (sg.count == 0) continue;
        checked++;
        if (checked % every != 0) continue;
        eligible[s] = 1;
        status[s] = 0;
        any = 1;
    }
    if (checked_after) *checked_after = checked;
    return any;
}

// The compact search over the keyframes with eligible[s] != 0 (positions in get_keyframes() order);
// status[s] of those keyframes is set; cands / matches as in vsm_loop_detect_compact.
static int loop_compact_eligible(vsm_ctx* ctx, const std::vector<char>& eligible, const float* query, int32_t nq, float ratio,
                                 int32_t min_matches, int32_t* status, vsm_loop_candidate* cands, int32_t cand_cap,
                                 int32_t* n_cands, vsm_dmatch* matches, int64_t match_cap, int64_t* n_matches) {
    *n_cands = 0;
    *n_matches = 0;
    std::vector<int32_t> pos;
    bool fits = ctx->engine != VSM_ENGINE_SIMT && ctx->engine != VSM_ENGINE_TENSOR_PAIR;
    for (int s = 0; s < (int)eligible.size(); s++) {
        if (!eligible[s]) continue;
        const int32_t cnt = ctx->segs[ctx->kf_order[s]].count;
        status[s] = 0;
        if (cnt < 2) continue;                                        // no two-entry list, no survivor (:57): status stays 0
        if (cnt > UNIT_TILES * TILE_N) fits = false;                  // a keyframe must be one work unit
        pos.push_back(s);
    }
    if (pos.empty() || nq == 0) return VSM_OK;
    bool overflow = !fits;
    if (fits)
        TRY(loop_compact_impl(ctx, query, nq, ratio, min_matches, pos, status, cands, cand_cap, n_cands, matches, match_cap,
                              n_matches, &overflow));
    if (!overflow) return VSM_OK;
    // too many open pairs for the compact buffers (or an engine / keyframe size the fused units do not
    // cover): the record-based per-keyframe search, packed on the host
    const int nkf = (int)eligible.size();
    std::vector<int32_t> cnt(nkf, 0);
    std::vector<vsm_dmatch> slab((size_t)nkf * nq);
    TRY(segmented_impl(ctx, query, nq, ratio, &eligible, cnt.data(), slab.data()));
    int nc = 0;
    int64_t nm = 0;
    for (int s = 0; s < nkf; s++) {
        if (!eligible[s]) continue;
        status[s] = cnt[s];
        if (cnt[s] < min_matches || cnt[s] == 0) continue;
        if (nc < cand_cap) cands[nc] = vsm_loop_candidate{s, cnt[s], nm};
        for (int i = 0; i < cnt[s]; i++)
            if (nm + i < match_cap) matches[nm + i] = slab[(size_t)s * nq + i];
        nc++;
        nm += cnt[s];
    }
    *n_cands = nc;
    *n_matches = nm;
    return VSM_OK;
}

int vsm_loop_detect_compact(vsm_ctx* ctx, int32_t cur_frame_id, int32_t min_gap, int32_t every, int32_t checked_before,
                            const float* query, int32_t nq, float ratio, int32_t min_matches, int32_t* status,
                            vsm_loop_candidate* cands, int32_t cand_cap, int32_t* n_cands, vsm_dmatch* matches,
                            int64_t match_cap, int64_t* n_matches, int32_t* checked_after) {
    if (!ctx || nq < 0 || (nq > 0 && !query) || !status || every <= 0 || checked_before < 0 || cand_cap < 0 || match_cap < 0 ||
        !n_cands || !n_matches || (cand_cap > 0 && !cands) || (match_cap > 0 && !matches))
        return ctx ? fail(ctx, VSM_ERR_INVALID, "vsm_loop_detect_compact: bad argument") : VSM_ERR_INVALID;
    static_assert(sizeof(vsm_loop_candidate) == sizeof(LoopCand), "vsm_loop_candidate layout");
    *n_cands = 0;
    *n_matches = 0;
    std::vector<char> eligible;
    const int any = loop_eligible(ctx, cur_frame_id, min_gap, every, checked_before, eligible, status, checked_after);
    if (nq == 0 || !any) return VSM_OK;                               // :22 (empty current frame)
    return loop_compact_eligible(ctx, eligible, query, nq, ratio, min_matches, status, cands, cand_cap, n_cands, matches,
                                 match_cap, n_matches);
}

int vsm_track_local_map(vsm_ctx* ctx, const vsm_track_cfg* cfg, const float* kp_xy, const float* desc, int32_t nkp,
                        const double* mp_pos, const float* mp_desc, const uint8_t* mp_valid, int32_t nmp,
                        const double* R_cam, const double* t_cam, int32_t* indices, int32_t* obs_mp, int32_t* obs_ki,
                        int32_t* tracked, int32_t* best_ki, double* best_dist) {
    if (!ctx || !cfg || nkp < 0 || nmp < 0 || !tracked || !R_cam || !t_cam || cfg->cell_size <= 0 ||
        (nkp > 0 && (!kp_xy || !desc || !indices)) || (nmp > 0 && (!mp_pos || !obs_mp || !obs_ki)))
        return ctx ? fail(ctx, VSM_ERR_INVALID, "vsm_track_local_map: bad argument") : VSM_ERR_INVALID;
    *tracked = 0;
    if (nkp == 0 || nmp == 0) return VSM_OK;                               // src/Slam.cpp:384
    // mp_desc == NULL: the resident map-point table (ids 0 .. nmp-1; its validity flags too if mp_valid is
    // NULL), or -- when the table is empty -- the first nmp rows of the frame store
    const bool from_table = !mp_desc && ctx->n_points >= nmp;
    if (!mp_desc && !from_table && ctx->store_rows < nmp)
        return fail(ctx, VSM_ERR_INVALID, "vsm_track_local_map: neither the map-point table nor the store holds nmp rows");
    // the reference's visiting order: its cell grid (:391-401), cells row-major, keypoints in push order
    const int GW = (cfg->width + cfg->cell_size - 1) / cfg->cell_size, GH = (cfg->height + cfg->cell_size - 1) / cfg->cell_size;
    std::vector<std::pair<int, int>> order;                                // (cell, ki)
    order.reserve(nkp);
    for (int ki = 0; ki < nkp; ki++) {
        const int gx = std::min((int)(kp_xy[2 * ki] / cfg->cell_size), GW - 1);
        const int gy = std::min((int)(kp_xy[2 * ki + 1] / cfg->cell_size), GH - 1);
        if (gx >= 0 && gy >= 0) order.push_back({gy * GW + gx, ki});
    }
    std::stable_sort(order.begin(), order.end(), [](const std::pair<int, int>& a, const std::pair<int, int>& b) { return a.first < b.first; });
    const int nord = (int)order.size();
    std::vector<float> sxy((size_t)std::max(nord, 1) * 2);
    std::vector<int32_t> sid(std::max(nord, 1));
    for (int k = 0; k < nord; k++) { sxy[2 * k] = kp_xy[2 * order[k].second]; sxy[2 * k + 1] = kp_xy[2 * order[k].second + 1]; sid[k] = order[k].second; }

    TRY(begin_call(ctx));
    TRY(arena_reserve(ctx, ctx->scratch, (int64_t)nkp + (mp_desc ? nmp : 0), 0));
    TRY(upload_plain(ctx, desc, 0, nkp));
    if (mp_desc) TRY(upload_plain(ctx, mp_desc, nkp, nmp));
    const size_t o_xy = 0, o_id = align16(o_xy + sxy.size() * 4), o_pos = align16(o_id + sid.size() * 4),
                 o_valid = align16(o_pos + (size_t)nmp * 24), o_bk = align16(o_valid + (size_t)nmp),
                 o_bd = align16(o_bk + (size_t)nmp * 4), total = align16(o_bd + (size_t)nmp * 8);
    TRY(ensure(ctx, ctx->d_track, total));
    uint8_t* d = ctx->d_track.p;
    CK(cudaMemcpyAsync(d + o_xy, sxy.data(), sxy.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d + o_id, sid.data(), sid.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d + o_pos, mp_pos, (size_t)nmp * 24, cudaMemcpyHostToDevice, ctx->stream));
    if (mp_valid) CK(cudaMemcpyAsync(d + o_valid, mp_valid, (size_t)nmp, cudaMemcpyHostToDevice, ctx->stream));
    TrackCfg tc;
    tc.fx = cfg->fx; tc.fy = cfg->fy; tc.cx = cfg->cx; tc.cy = cfg->cy;
    tc.depth_min = cfg->depth_min; tc.depth_max = cfg->depth_max;
    tc.radius_sq = cfg->search_radius * cfg->search_radius;                // :409
    tc.desc_threshold = cfg->desc_threshold;
    for (int i = 0; i < 9; i++) tc.R[i] = R_cam[i];
    for (int i = 0; i < 3; i++) tc.t[i] = t_cam[i];
    tc.width = cfg->width; tc.height = cfg->height;
    const float* d_mp_desc = mp_desc ? ctx->scratch.f32 + (int64_t)nkp * VSM_DIM
                                     : (from_table ? reinterpret_cast<const float*>(ctx->pt_f32.p) : ctx->store.f32);
    const uint8_t* d_valid = mp_valid ? d + o_valid : (from_table ? ctx->pt_valid.p : nullptr);
    track_local_map_kernel<<<(unsigned)(((int64_t)nmp * 32 + 255) / 256), 256, 0, ctx->stream>>>(
        tc, reinterpret_cast<const float2*>(d + o_xy), reinterpret_cast<const int32_t*>(d + o_id), nord, ctx->scratch.f32,
        reinterpret_cast<const double*>(d + o_pos), d_mp_desc, d_valid, nmp,
        reinterpret_cast<int32_t*>(d + o_bk), reinterpret_cast<double*>(d + o_bd));
    ctx->launches++;
    CK(cudaGetLastError());
    std::vector<int32_t> bk(nmp);
    std::vector<double> bd(nmp);
    CK(cudaMemcpyAsync(bk.data(), d + o_bk, (size_t)nmp * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(bd.data(), d + o_bd, (size_t)nmp * 8, cudaMemcpyDeviceToHost, ctx->stream));
    TRY(end_call(ctx, true));
    // the sequential part (:465-470): map points in id order, strictly smaller distance replaces
    std::vector<double> best_desc_dist(nkp, 1e9);                          // :389
    int n = 0;
    for (int mp = 0; mp < nmp; mp++) {
        if (best_ki) best_ki[mp] = bk[mp];
        if (best_dist) best_dist[mp] = bd[mp];
        const int ki = bk[mp];
        if (ki >= 0 && bd[mp] < best_desc_dist[ki]) {
            indices[ki] = mp;
            best_desc_dist[ki] = bd[mp];
            obs_mp[n] = mp; obs_ki[n] = ki;
            n++;
        }
    }
    *tracked = n;
    return VSM_OK;
}

int vsm_db_top2_masked(vsm_ctx* ctx, const float* query, int32_t nq, const uint8_t* mask, int64_t n_mask, int64_t* idx,
                       float* dist) {
    if (!ctx || nq < 0 || (nq > 0 && (!query || !idx || !dist)) || n_mask != ctx->store_rows || (n_mask > 0 && !mask))
        return ctx ? fail(ctx, VSM_ERR_INVALID, "vsm_db_top2_masked: bad argument (mask length must equal the store rows)")
                   : VSM_ERR_INVALID;
    if (nq == 0) return VSM_OK;
    for (int i = 0; i < nq * 2; i++) { idx[i] = -1; dist[i] = FLT_MAX; }
    if (n_mask == 0) return VSM_OK;
    TRY(begin_call(ctx));
    // the mask goes to the device as it is (one byte per store row); the selected rows (the reference's
    // mp_ids_vec, src/Slam.cpp:757) are found, numbered and gathered there -- no O(rows) loop on the host
    TRY(ensure(ctx, ctx->d_pt_tmp, select_tmp_bytes(n_mask) + align16((size_t)n_mask)));
    uint8_t* d_mask = ctx->d_pt_tmp.p + select_tmp_bytes(n_mask);
    CK(cudaMemcpyAsync(d_mask, mask, (size_t)n_mask, cudaMemcpyHostToDevice, ctx->stream));
    return select_and_search(ctx, query, nq, d_mask, nullptr, n_mask, ctx->store.f32, ctx->d_pt_tmp.p, idx, dist, nullptr);
}

int vsm_db_top2_keys_device(vsm_ctx* ctx, const float* d_query, int32_t nq, int64_t row_offset, uint64_t* d_keys,
                            int32_t sync) {
    if (!ctx || nq < 0 || (nq > 0 && (!d_query || !d_keys)) || row_offset < 0 ||
        row_offset + ctx->store_rows > 0xFFFFFFFFll)
        return ctx ? fail(ctx, VSM_ERR_INVALID, "vsm_db_top2_keys_device: bad argument") : VSM_ERR_INVALID;
    if (nq == 0) return VSM_OK;
    TRY(begin_call(ctx));
    TRY(arena_reserve(ctx, ctx->scratch, nq, 0));
    HProblem p;
    db_problem(ctx, d_query, nq, p);
    static const int timeline = getenv("VSM_DEBUG_TIMELINE") ? std::max(2, atoi(getenv("VSM_DEBUG_TIMELINE"))) : 0;
    queue_convert(ctx, d_query, 0, nq);
    TRY(run_problems(ctx, {p}, {}, nq, 0, timeline));
    globalize_keys_kernel<<<(nq * 2 + 255) / 256, 256, 0, ctx->stream>>>(
        ctx->d_out_key, nq * 2, (uint32_t)row_offset, reinterpret_cast<unsigned long long*>(d_keys));
    ctx->launches++;
    CK(cudaGetLastError());
    return end_call(ctx, sync != 0);
}

int vsm_xchg_create(vsm_ctx* ctx, int32_t rank, int32_t world, int32_t nq_cap, uint8_t handle_out[64]) {
    if (!ctx || !handle_out || world < 1 || world > XCHG_MAX_WORLD || rank < 0 || rank >= world || nq_cap <= 0)
        return ctx ? fail(ctx, VSM_ERR_INVALID, "vsm_xchg_create: bad argument") : VSM_ERR_INVALID;
    if (ctx->xchg_buf) return fail(ctx, VSM_ERR_INVALID, "vsm_xchg_create: already created");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    CK(cudaSetDevice(ctx->device));
    const size_t bytes = XCHG_FLAG_BYTES + (size_t)2 * world * nq_cap * 2 * sizeof(unsigned long long);
    CK(cudaMalloc(&ctx->xchg_buf, bytes));
    CK(cudaMemset(ctx->xchg_buf, 0, bytes));
    CK(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, ctx->xchg_buf));
    memcpy(handle_out, &h, 64);
    ctx->xchg_rank = rank; ctx->xchg_world = world; ctx->xchg_nq_cap = nq_cap; ctx->xchg_step = 0;
    return VSM_OK;
}

int vsm_xchg_connect(vsm_ctx* ctx, const uint8_t* handles) {
    if (!ctx || !handles || !ctx->xchg_buf) return ctx ? fail(ctx, VSM_ERR_INVALID, "vsm_xchg_connect: create first") : VSM_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    for (int r = 0; r < ctx->xchg_world; r++) {
        if (r == ctx->xchg_rank) { ctx->xchg_peer_ptr[r] = ctx->xchg_buf; }
        else {
            cudaIpcMemHandle_t h;
            memcpy(&h, handles + (size_t)r * 64, 64);
            CK(cudaIpcOpenMemHandle(&ctx->xchg_peer_ptr[r], h, cudaIpcMemLazyEnablePeerAccess));
        }
        ctx->xchg_peers.base[r] = (unsigned long long)(uintptr_t)ctx->xchg_peer_ptr[r];
    }
    ctx->xchg_connected = true;
    return VSM_OK;
}

// Soft failures reported by kernels through the pinned status word (stream must be idle).
static int check_status(vsm_ctx* ctx) {
    if (ctx->h_status && ctx->h_status[0] == XCHG_TIMEOUT) {
        char b[160];
        snprintf(b, sizeof b, "peer-memory exchange timed out waiting for rank %u (it did not make the matching call)",
                 ctx->h_status[1]);
        ctx->h_status[0] = 0;
        ctx->err = b;
        return VSM_ERR_TIMEOUT;
    }
    return VSM_OK;
}

static int xchg_search(vsm_ctx* ctx, const float* d_query, int32_t nq, int64_t row_offset, int64_t* d_idx_out,
                       float* d_dist_out) {
    HProblem p;
    db_problem(ctx, d_query, nq, p);
    TRY(run_problems(ctx, {p}, {}, nq, 0));
    static const double timeout_s = getenv("VSM_XCHG_TIMEOUT_S") ? atof(getenv("VSM_XCHG_TIMEOUT_S")) : 10.0;
    const uint32_t step = ctx->xchg_step + 1;
    xchg_publish_merge_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->d_out_key, nq, (uint32_t)row_offset, ctx->xchg_peers,
                                                          ctx->xchg_rank, ctx->xchg_world, ctx->xchg_nq_cap, step,
                                                          d_idx_out, d_dist_out, (long long)(timeout_s * 2.0e9),
                                                          ctx->h_status);
    CK(cudaGetLastError());
    ctx->xchg_step = step;                       // only once the kernel is enqueued: a failed call does not desynchronise the ranks
    ctx->launches++;
    return VSM_OK;
}

int vsm_db_top2_xchg_device(vsm_ctx* ctx, const float* d_query, int32_t nq, int64_t row_offset, int64_t* d_idx_out,
                            float* d_dist_out, int32_t sync) {
    if (!ctx || nq <= 0 || !d_query || !d_idx_out || !d_dist_out || row_offset < 0 ||
        row_offset + ctx->store_rows > 0xFFFFFFFFll)
        return ctx ? fail(ctx, VSM_ERR_INVALID, "vsm_db_top2_xchg_device: bad argument") : VSM_ERR_INVALID;
    if (!ctx->xchg_connected || nq > ctx->xchg_nq_cap)
        return fail(ctx, VSM_ERR_INVALID, "vsm_db_top2_xchg_device: exchange not connected or nq above its capacity");
    TRY(begin_call(ctx));
    TRY(arena_reserve(ctx, ctx->scratch, nq, 0));
    queue_convert(ctx, d_query, 0, nq);
    TRY(xchg_search(ctx, d_query, nq, row_offset, d_idx_out, d_dist_out));
    TRY(end_call(ctx, sync != 0));
    return sync ? check_status(ctx) : VSM_OK;
}

int vsm_db_top2_xchg(vsm_ctx* ctx, const float* query, int32_t nq, int64_t row_offset, int64_t* idx, float* dist) {
    if (!ctx || nq <= 0 || !query || !idx || !dist || row_offset < 0 || row_offset + ctx->store_rows > 0xFFFFFFFFll)
        return ctx ? fail(ctx, VSM_ERR_INVALID, "vsm_db_top2_xchg: bad argument") : VSM_ERR_INVALID;
    if (!ctx->xchg_connected || nq > ctx->xchg_nq_cap)
        return fail(ctx, VSM_ERR_INVALID, "vsm_db_top2_xchg: exchange not connected or nq above its capacity");
    TRY(begin_call(ctx));
    TRY(arena_reserve(ctx, ctx->scratch, nq, 0));
    TRY(upload_scratch(ctx, query, 0, nq));
    // the merged result goes straight into pinned host memory (zero-copy stores of 24 B per query)
    const size_t nb = (size_t)nq * 2 * (sizeof(int64_t) + sizeof(float));
    TRY(ensure_host(ctx, ctx->h_result, ctx->h_result_cap, std::max<size_t>(nb, 16)));
    int64_t* h_idx = reinterpret_cast<int64_t*>(ctx->h_result);
    float* h_dist = reinterpret_cast<float*>(ctx->h_result + (size_t)nq * 2 * sizeof(int64_t));
    TRY(xchg_search(ctx, ctx->scratch.f32, nq, row_offset, h_idx, h_dist));
    TRY(end_call(ctx, true));
    TRY(check_status(ctx));
    memcpy(idx, h_idx, (size_t)nq * 2 * sizeof(int64_t));
    memcpy(dist, h_dist, (size_t)nq * 2 * sizeof(float));
    return VSM_OK;
}

int vsm_merge_keys_device(vsm_ctx* ctx, const uint64_t* d_keys_in, int32_t nshard, int32_t nq, int64_t* d_idx_out,
                          float* d_dist_out, int32_t sync) {
    if (!ctx || nshard <= 0 || nq < 0 || (nq > 0 && (!d_keys_in || !d_idx_out || !d_dist_out)))
        return ctx ? fail(ctx, VSM_ERR_INVALID, "vsm_merge_keys_device: bad argument") : VSM_ERR_INVALID;
    if (nq == 0) return VSM_OK;
    CK(cudaSetDevice(ctx->device));
    merge_keys_kernel<<<(nq + 127) / 128, 128, 0, ctx->stream>>>(
        reinterpret_cast<const unsigned long long*>(d_keys_in), nshard, nq, d_idx_out, d_dist_out);
    CK(cudaGetLastError());
    if (sync) CK(cudaStreamSynchronize(ctx->stream));
    return VSM_OK;
}

int vsm_db_segmented(vsm_ctx* ctx, const float* query, int32_t nq, float ratio, int32_t* counts, vsm_dmatch* matches) {
    if (!ctx || nq < 0 || (nq > 0 && !query) || !counts)
        return ctx ? fail(ctx, VSM_ERR_INVALID, "vsm_db_segmented: bad argument") : VSM_ERR_INVALID;
    const int nseg = (int)ctx->kf_order.size();
    for (int s = 0; s < nseg; s++) counts[s] = 0;
    if (nq == 0 || nseg == 0) return VSM_OK;
    return segmented_impl(ctx, query, nq, ratio, nullptr, counts, matches);
}

int vsm_loop_detect_shard(vsm_ctx* ctx, int32_t cur_frame_id, int32_t min_gap, int32_t every, int32_t checked_before,
                          const float* query, int32_t nq, float ratio, int32_t* status, vsm_dmatch* matches,
                          int32_t* checked_after) {
    if (!ctx || nq < 0 || (nq > 0 && !query) || !status || every <= 0 || checked_before < 0)
        return ctx ? fail(ctx, VSM_ERR_INVALID, "vsm_loop_detect: bad argument") : VSM_ERR_INVALID;
    std::vector<char> eligible;
    const int any = loop_eligible(ctx, cur_frame_id, min_gap, every, checked_before, eligible, status, checked_after);
    if (nq == 0 || !any) return VSM_OK;                               // :22 (empty current frame)
    return segmented_impl(ctx, query, nq, ratio, &eligible, status, matches);
}

int vsm_loop_detect(vsm_ctx* ctx, int32_t cur_frame_id, int32_t min_gap, int32_t every, const float* query,
                    int32_t nq, float ratio, int32_t* status, vsm_dmatch* matches) {
    return vsm_loop_detect_shard(ctx, cur_frame_id, min_gap, every, 0, query, nq, ratio, status, matches, nullptr);
}

int vsm_store_set_frame_ids(vsm_ctx* ctx, const int32_t* frame_ids, int32_t n) {
    if (!ctx || !frame_ids || n != (int32_t)ctx->kf_order.size())
        return ctx ? fail(ctx, VSM_ERR_INVALID, "vsm_store_set_frame_ids: n must equal the keyframe count") : VSM_ERR_INVALID;
    for (int s = 0; s < n; s++) ctx->segs[ctx->kf_order[s]].frame_id = frame_ids[s];
    return VSM_OK;
}

int vsm_merge_top2_device(vsm_ctx* ctx, const int64_t* d_idx_in, const float* d_dist_in, int32_t nshard, int32_t nq,
                          int64_t* d_idx_out, float* d_dist_out, int32_t sync) {
    if (!ctx || nshard <= 0 || nq < 0 || (nq > 0 && (!d_idx_in || !d_dist_in || !d_idx_out || !d_dist_out)))
        return ctx ? fail(ctx, VSM_ERR_INVALID, "vsm_merge_top2_device: bad argument") : VSM_ERR_INVALID;
    if (nq == 0) return VSM_OK;
    CK(cudaSetDevice(ctx->device));
    merge_kernel<<<(nq + 127) / 128, 128, 0, ctx->stream>>>(d_idx_in, d_dist_in, nshard, nq, d_idx_out, d_dist_out);
    CK(cudaGetLastError());
    if (sync) CK(cudaStreamSynchronize(ctx->stream));
    return VSM_OK;
}

int vsm_synth_rows_device(vsm_ctx* ctx, float* d_dst, int64_t row0, int64_t n, uint64_t seed) {
    if (!ctx || n < 0 || (n > 0 && !d_dst)) return ctx ? fail(ctx, VSM_ERR_INVALID, "vsm_synth_rows_device: bad argument") : VSM_ERR_INVALID;
    if (n == 0) return VSM_OK;
    CK(cudaSetDevice(ctx->device));
    const int64_t blocks = std::min<int64_t>((n + 7) / 8, (int64_t)ctx->num_sms * 16);
    synth_rows_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(d_dst, row0, n, seed);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    return VSM_OK;
}

int vsm_debug_fetch_dump(vsm_ctx* ctx, void* out, int64_t bytes) {
    if (!ctx || !out || bytes <= 0 || bytes > (int64_t)(TILE_M * TILE_N * sizeof(float))) return VSM_ERR_INVALID;
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpy(out, ctx->d_dump, (size_t)bytes, cudaMemcpyDeviceToHost));
    return VSM_OK;
}

int vsm_debug_tile_scores(vsm_ctx* ctx, const float* query, int32_t nq, const float* train, int32_t nt, float* out) {
    if (!ctx || !query || !train || !out || nq <= 0 || nt <= 0)
        return ctx ? fail(ctx, VSM_ERR_INVALID, "vsm_debug_tile_scores: bad argument") : VSM_ERR_INVALID;
    if (ctx->engine == VSM_ENGINE_SIMT) return fail(ctx, VSM_ERR_INVALID, "no tensor-core pass in the SIMT engine");
    TRY(begin_call(ctx));
    TRY(arena_reserve(ctx, ctx->scratch, (int64_t)nq + nt, 0));
    CK(cudaMemsetAsync(ctx->d_dump, 0, TILE_M * TILE_N * sizeof(float), ctx->stream));
    TRY(upload_scratch(ctx, query, 0, nq));
    TRY(upload_scratch(ctx, train, nq, nt));
    std::vector<HProblem> probs{scratch_vs_scratch(ctx, 0, nq, nq, nt, 0)};
    TRY(run_problems(ctx, probs, {}, nq, 0, 1));
    CK(cudaMemcpyAsync(out, ctx->d_dump, TILE_M * TILE_N * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    return end_call(ctx, true);
}

#include "vsm_group.inl"

}  // extern "C"
