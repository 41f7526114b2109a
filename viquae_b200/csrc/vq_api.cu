// C ABI of the engine (include/viquae_b200.h): index object, kernel-variant dispatch, workspace carving.
#include "../../include/viquae_b200.h"

#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <thread>
#include <vector>

#include "vq_common.cuh"
#include "vq_kernels.h"

static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess) return fail(VQ_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
    } while (0)

// Every entry point that works on an index makes the index's device current for the duration of the call and
// restores the caller's device afterwards (an index on cuda:1 must not silently switch a cuda:0 caller).
struct DeviceGuard {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int dev) {
        int cur = -1;
        err = cudaGetDevice(&cur);
        if (err == cudaSuccess && cur != dev) {
            err = cudaSetDevice(dev);
            if (err == cudaSuccess) prev = cur;
        }
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};
#define VQ_ON_DEVICE(dev)                                                                                      \
    DeviceGuard guard__(dev);                                                                                  \
    if (guard__.err != cudaSuccess) return fail(VQ_ERR_CUDA, "cudaSetDevice(%d): %s", (dev), cudaGetErrorString(guard__.err))

#define VQ_TIMING_SLOTS 256

struct vq_index {
    int d, dtype, device, ld, n_sms, variant;
    int kd, kdtype, seg;  // what the kernels contract over: kd columns of kdtype (VQ_F32X3: 6 segments of seg bf16 columns)
    int kpad;             // VQ_F32: rows are stored as k-major tiles of 128 rows, [tile][kpad][128] floats (K1's TMA boxes)
    uint64_t capacity, ntotal;
    int64_t id_offset;
    void* data;
    // engine-owned buffers of the host-buffer entry point
    void* hws;
    size_t hws_bytes;
    // host ingest pipeline: two pinned buffers + two device staging buffers, filled alternately (allocated at the
    // first host-side add, kept for the index's lifetime)
    void* pin[2];
    void* stage[2];
    cudaEvent_t ingest_ev[2];
    int ingest_ev_ok[2];
    size_t ingest_bytes;
    // measurement support: CUDA events around the dominant kernel of each search
    int timing;
    int n_timed;
    uint64_t launches;  // kernels launched by vq_index_search so far
    cudaEvent_t ev0[VQ_TIMING_SLOTS], ev1[VQ_TIMING_SLOTS];
};

static size_t esize(int dtype) { return dtype == VQ_F32 ? 4 : 2; }  // of one stored element (VQ_F32X3 stores bf16)
static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
// bytes of the first `rows` rows of the corpus matrix (VQ_F32: whole 128-row tiles)
static size_t storage_bytes(const vq_index* idx, uint64_t rows);
static bool dtype_ok(int t) { return t == VQ_F32 || t == VQ_BF16 || t == VQ_F16; }  // of vectors crossing the ABI
static bool storage_ok(int t) { return dtype_ok(t) || t == VQ_F32X3; }

constexpr size_t VQ_INGEST_CHUNK = 32ull << 20;  // per pipeline stage

static void ingest_release(vq_index* idx) {
    for (int b = 0; b < 2; b++) {
        if (idx->pin[b]) cudaFreeHost(idx->pin[b]);
        if (idx->stage[b]) cudaFree(idx->stage[b]);
        if (idx->ingest_ev_ok[b]) cudaEventDestroy(idx->ingest_ev[b]);
        idx->pin[b] = idx->stage[b] = nullptr;
        idx->ingest_ev_ok[b] = 0;
    }
    idx->ingest_bytes = 0;
}

static cudaError_t ingest_reserve(vq_index* idx, size_t bytes) {
    if (idx->ingest_bytes >= bytes) return cudaSuccess;
    ingest_release(idx);
    for (int b = 0; b < 2; b++) {
        cudaError_t e = cudaHostAlloc(&idx->pin[b], bytes, cudaHostAllocDefault);
        if (e == cudaSuccess) e = cudaMalloc(&idx->stage[b], bytes);
        if (e == cudaSuccess) {
            e = cudaEventCreateWithFlags(&idx->ingest_ev[b], cudaEventDisableTiming);
            idx->ingest_ev_ok[b] = e == cudaSuccess;
        }
        if (e != cudaSuccess) {
            ingest_release(idx);
            return e;
        }
    }
    idx->ingest_bytes = bytes;
    return cudaSuccess;
}

// pageable -> pinned copy on several host threads (one memcpy stream per thread saturates ~1/4 of a socket)
static void host_copy(void* dst, const void* src, size_t bytes) {
    static const unsigned n_threads = [] {
        unsigned h = std::thread::hardware_concurrency();
        unsigned n = h / 2;
        return n < 1 ? 1u : (n > 8 ? 8u : n);
    }();
    if (bytes < (4u << 20) || n_threads == 1) {
        memcpy(dst, src, bytes);
        return;
    }
    const size_t per = ((bytes / n_threads) + 4095) & ~(size_t)4095;
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < n_threads; t++) {
        const size_t a = (size_t)t * per;
        if (a >= bytes) break;
        const size_t len = bytes - a < per ? bytes - a : per;
        pool.emplace_back([=] { memcpy((uint8_t*)dst + a, (const uint8_t*)src + a, len); });
    }
    memcpy(dst, src, per < bytes ? per : bytes);
    for (auto& th : pool) th.join();
}

static size_t storage_bytes(const vq_index* idx, uint64_t rows) {
    if (idx->dtype == VQ_F32) return (size_t)((rows + 127) / 128) * idx->kpad * 128 * sizeof(float);
    return (size_t)rows * idx->ld * esize(idx->dtype);
}

// [n, d] rows of src_dtype -> rows row0 .. row0+n-1 of a matrix in the index's storage layout at `base` (the corpus, or a
// query operand: the query-side arrangement of VQ_F32X3; VQ_F32: k-major tiles either way)
static int to_storage(const vq_index* idx, const void* src, int src_dtype, void* base, uint64_t row0, long long n, int l2norm, int query_side,
                      cudaStream_t s, long long n_pad = 0) {
    if (idx->dtype == VQ_F32) return vq::tile_rows_launch(src, src_dtype, idx->d, (float*)base, (long long)row0, idx->kpad, n, idx->d, l2norm, s, n_pad);
    void* dst = (uint8_t*)base + row0 * (size_t)idx->ld * esize(idx->dtype);
    if (idx->dtype == VQ_F32X3) return vq::split_rows_launch(src, src_dtype, idx->d, dst, idx->seg, n, idx->d, l2norm, query_side, s, n_pad);
    return vq::prep_rows_launch(src, src_dtype, idx->d, dst, idx->dtype, idx->ld, n, idx->d, l2norm, s, n_pad);
}

// stored corpus rows row0 .. row0+n-1 -> [n, d] rows of dst_dtype
static int from_storage(const vq_index* idx, uint64_t row0, void* dst, int dst_dtype, long long n, cudaStream_t s) {
    if (idx->dtype == VQ_F32) return vq::untile_rows_launch((const float*)idx->data, (long long)row0, idx->kpad, dst, dst_dtype, idx->d, n, idx->d, s);
    const void* src = (const uint8_t*)idx->data + row0 * (size_t)idx->ld * esize(idx->dtype);
    if (idx->dtype == VQ_F32X3) return vq::join_rows_launch(src, idx->seg, dst, dst_dtype, idx->d, n, idx->d, s);
    return vq::prep_rows_launch(src, idx->dtype, idx->ld, dst, dst_dtype, idx->d, n, idx->d, 0, s);
}

extern "C" {

const char* vq_last_error(void) { return g_err; }
const char* vq_version(void) { return "viquae_b200 0.1 (sm_100a)"; }

int vq_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int vq_index_create(vq_index** out, int d, int dtype, int device, uint64_t capacity) {
    if (!out || d <= 0 || !storage_ok(dtype)) return fail(VQ_ERR_INVALID, "vq_index_create: bad arguments (d=%d dtype=%d)", d, dtype);
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(VQ_ERR_INVALID, "vq_index_create: device %d not in [0, %d)", device, ndev);
    VQ_ON_DEVICE(device);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(VQ_ERR_UNSUPPORTED, "viquae_b200 needs an sm_100 GPU, found sm_%d%d", prop.major, prop.minor);
    vq_index* idx = new vq_index();
    idx->d = d;
    idx->dtype = dtype;
    idx->device = device;
    idx->ld = dtype == VQ_F32 ? (d + 3) / 4 * 4 : (d + 7) / 8 * 8;  // 16-byte row stride (float4 / TMA)
    idx->kd = d;
    idx->kdtype = dtype;
    idx->seg = 0;
    idx->kpad = dtype == VQ_F32 ? vq::f32_kpad(d) : 0;
    if (dtype == VQ_F32X3) {
        idx->seg = (d + 7) / 8 * 8;
        idx->ld = idx->kd = 6 * idx->seg;
        idx->kdtype = VQ_BF16;
    }
    idx->n_sms = prop.multiProcessorCount;
    idx->variant = VQ_VARIANT_AUTO;
    idx->capacity = 0;
    idx->ntotal = 0;
    idx->id_offset = 0;
    idx->data = nullptr;
    idx->hws = nullptr;
    idx->hws_bytes = 0;
    idx->timing = 0;
    idx->n_timed = 0;
    idx->launches = 0;
    if (capacity) {
        cudaError_t e = cudaMalloc(&idx->data, storage_bytes(idx, capacity));
        if (e != cudaSuccess) {
            delete idx;
            return fail(VQ_ERR_NOMEM, "vq_index_create: cudaMalloc of %llu rows failed: %s", (unsigned long long)capacity, cudaGetErrorString(e));
        }
        if (dtype == VQ_F32) cudaMemset(idx->data, 0, storage_bytes(idx, capacity));  // rows past ntotal of the last tile are read (and discarded)
        idx->capacity = capacity;
    }
    *out = idx;
    return VQ_OK;
}

void vq_index_destroy(vq_index* idx) {
    if (!idx) return;
    DeviceGuard guard(idx->device);
    if (idx->data) cudaFree(idx->data);
    if (idx->hws) cudaFree(idx->hws);
    ingest_release(idx);
    if (idx->timing)
        for (int i = 0; i < VQ_TIMING_SLOTS; i++) {
            cudaEventDestroy(idx->ev0[i]);
            cudaEventDestroy(idx->ev1[i]);
        }
    delete idx;
}

uint64_t vq_index_ntotal(const vq_index* idx) { return idx ? idx->ntotal : 0; }
uint64_t vq_index_launch_count(const vq_index* idx) { return idx ? idx->launches : 0; }
int vq_index_dim(const vq_index* idx) { return idx ? idx->d : 0; }
int vq_index_dtype(const vq_index* idx) { return idx ? idx->dtype : -1; }

int vq_index_set_id_offset(vq_index* idx, int64_t off) {
    if (!idx) return fail(VQ_ERR_INVALID, "null index");
    idx->id_offset = off;
    return VQ_OK;
}
int vq_index_set_variant(vq_index* idx, int variant) {
    if (!idx || variant < 0 || variant > 2) return fail(VQ_ERR_INVALID, "vq_index_set_variant: bad variant %d", variant);
    idx->variant = variant;
    return VQ_OK;
}

int vq_index_set_timing(vq_index* idx, int enable) {
    if (!idx) return fail(VQ_ERR_INVALID, "null index");
    VQ_ON_DEVICE(idx->device);
    if (enable && !idx->timing) {
        for (int i = 0; i < VQ_TIMING_SLOTS; i++) {
            CUDA_TRY(cudaEventCreate(&idx->ev0[i]));
            CUDA_TRY(cudaEventCreate(&idx->ev1[i]));
        }
        idx->timing = 1;
    }
    if (!enable && idx->timing) {
        for (int i = 0; i < VQ_TIMING_SLOTS; i++) {
            cudaEventDestroy(idx->ev0[i]);
            cudaEventDestroy(idx->ev1[i]);
        }
        idx->timing = 0;
    }
    idx->n_timed = 0;
    return VQ_OK;
}

int vq_index_read_timing(vq_index* idx, double* total_ms, int* n_launches) {
    if (!idx || !total_ms || !n_launches) return fail(VQ_ERR_INVALID, "vq_index_read_timing: null pointer");
    double tot = 0.0;
    for (int i = 0; i < idx->n_timed; i++) {
        CUDA_TRY(cudaEventSynchronize(idx->ev1[i]));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, idx->ev0[i], idx->ev1[i]));
        tot += ms;
    }
    *total_ms = tot;
    *n_launches = idx->n_timed;
    idx->n_timed = 0;
    return VQ_OK;
}

static int is_device_ptr(const void* p) {
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

static int grow(vq_index* idx, uint64_t need, cudaStream_t s) {
    if (need <= idx->capacity) return VQ_OK;
    uint64_t cap = idx->capacity ? idx->capacity : 1024;
    while (cap < need) cap += cap / 2 + 1;
    void* nd = nullptr;
    cudaError_t e = cudaMalloc(&nd, storage_bytes(idx, cap));
    if (e != cudaSuccess && cap > need) {  // retry with the exact size
        cudaGetLastError();
        cap = need;
        e = cudaMalloc(&nd, storage_bytes(idx, cap));
    }
    if (e != cudaSuccess) return fail(VQ_ERR_NOMEM, "vq_index_add: cudaMalloc of %llu rows failed: %s", (unsigned long long)cap, cudaGetErrorString(e));
    const size_t used = storage_bytes(idx, idx->ntotal);
    if (used) CUDA_TRY(cudaMemcpyAsync(nd, idx->data, used, cudaMemcpyDeviceToDevice, s));
    if (idx->dtype == VQ_F32) CUDA_TRY(cudaMemsetAsync((uint8_t*)nd + used, 0, storage_bytes(idx, cap) - used, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    if (idx->data) cudaFree(idx->data);
    idx->data = nd;
    idx->capacity = cap;
    return VQ_OK;
}

int vq_index_add(vq_index* idx, const void* x, uint64_t n, int src_dtype, int l2norm, void* stream) {
    if (!idx || (!x && n) || !dtype_ok(src_dtype)) return fail(VQ_ERR_INVALID, "vq_index_add: bad arguments");
    if (n == 0) return VQ_OK;
    // rows are uint32 in the candidate keys and int32 TMA coordinates in the tensor-core kernel
    if (idx->ntotal + n > 0x7FFFFF00ull) return fail(VQ_ERR_UNSUPPORTED, "vq_index_add: more than 2^31-256 rows per shard");
    // fp32 exact mode: the k-major tiled matrix [tiles * kpad][128] is addressed with int32 TMA coordinates
    if (idx->dtype == VQ_F32 && ((idx->ntotal + n + 127) / 128) * (uint64_t)idx->kpad > 0x7FFFFF00ull)
        return fail(VQ_ERR_UNSUPPORTED, "vq_index_add: fp32 exact mode holds at most 2^31 / %d tiles of 128 rows per shard", idx->kpad);
    cudaStream_t s = (cudaStream_t)stream;
    VQ_ON_DEVICE(idx->device);
    int rc = grow(idx, idx->ntotal + n, s);
    if (rc) return rc;
    if (is_device_ptr(x)) {
        rc = to_storage(idx, x, src_dtype, idx->data, idx->ntotal, (long long)n, l2norm ? 2 : 0, 0, s);
        if (rc) return fail(VQ_ERR_CUDA, "vq_index_add: prep kernel launch failed (%d)", rc);
    } else {
        // host source: pageable rows -> pinned buffer (host threads) -> device staging (DMA) -> conversion kernel,
        // double-buffered so the host copy of chunk c+1 overlaps the DMA and kernel of chunk c
        const size_t src_row = (size_t)idx->d * esize(src_dtype);
        uint64_t chunk = VQ_INGEST_CHUNK / src_row;
        if (chunk < 1) chunk = 1;
        if (chunk > n) chunk = n;  // small adds keep small buffers (they grow with the first large one)
        if (idx->ingest_bytes >= src_row && chunk * src_row > idx->ingest_bytes && idx->ingest_bytes * 8 >= chunk * src_row)
            chunk = idx->ingest_bytes / src_row;  // buffers within 8x of the wish: reuse them rather than reallocate pinned memory
        cudaError_t e = ingest_reserve(idx, chunk * src_row);
        if (e != cudaSuccess) return fail(VQ_ERR_NOMEM, "vq_index_add: ingest buffers: %s", cudaGetErrorString(e));
        uint64_t c = 0;
        for (uint64_t r0 = 0; r0 < n; r0 += chunk, c++) {
            const uint64_t m = (n - r0 < chunk) ? n - r0 : chunk;
            const int b = (int)(c & 1);
            if (c >= 2) e = cudaEventSynchronize(idx->ingest_ev[b]);  // chunk c-2 has left this stage
            if (e == cudaSuccess) {
                host_copy(idx->pin[b], (const uint8_t*)x + r0 * src_row, m * src_row);
                e = cudaMemcpyAsync(idx->stage[b], idx->pin[b], m * src_row, cudaMemcpyHostToDevice, s);
            }
            if (e == cudaSuccess) {
                rc = to_storage(idx, idx->stage[b], src_dtype, idx->data, idx->ntotal + r0, (long long)m, l2norm ? 2 : 0, 0, s);
                if (rc == 0) e = cudaEventRecord(idx->ingest_ev[b], s);
            }
            if (e != cudaSuccess || rc) {
                cudaStreamSynchronize(s);
                return fail(VQ_ERR_CUDA, "vq_index_add: host staging failed: %s (%d)", cudaGetErrorString(e), rc);
            }
        }
        CUDA_TRY(cudaStreamSynchronize(s));  // the caller may free x; the staging buffers are reused by the next call
    }
    idx->ntotal += n;
    return VQ_OK;
}

int vq_index_copy_rows(const vq_index* idx, uint64_t start, uint64_t n, void* dst, int dst_dtype, void* stream) {
    if (!idx || (!dst && n) || !dtype_ok(dst_dtype) || start + n > idx->ntotal) return fail(VQ_ERR_INVALID, "vq_index_copy_rows: bad arguments");
    if (n == 0) return VQ_OK;
    cudaStream_t s = (cudaStream_t)stream;
    VQ_ON_DEVICE(idx->device);
    if (is_device_ptr(dst)) {
        int rc = from_storage(idx, start, dst, dst_dtype, (long long)n, s);
        if (rc) return fail(VQ_ERR_CUDA, "vq_index_copy_rows: launch failed (%d)", rc);
        return VQ_OK;
    }
    const size_t dst_row = (size_t)idx->d * esize(dst_dtype);
    uint64_t chunk = (64ull << 20) / dst_row;
    if (chunk < 1) chunk = 1;
    if (chunk > n) chunk = n;
    void* stage = nullptr;
    CUDA_TRY(cudaMalloc(&stage, chunk * dst_row));
    for (uint64_t r0 = 0; r0 < n; r0 += chunk) {
        const uint64_t m = (n - r0 < chunk) ? n - r0 : chunk;
        int rc = from_storage(idx, start + r0, stage, dst_dtype, (long long)m, s);
        cudaError_t e = rc ? cudaErrorUnknown : cudaMemcpyAsync((uint8_t*)dst + r0 * dst_row, stage, m * dst_row, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) {
            cudaFree(stage);
            return fail(VQ_ERR_CUDA, "vq_index_copy_rows: %s (%d)", cudaGetErrorString(e), rc);
        }
    }
    cudaFree(stage);
    return VQ_OK;
}

// ---- variant choice + workspace carving shared by workspace_bytes() and search()
#define VQ_MAX_PASSES 6
struct SearchPlan {
    int kind;  // 0 = fp32 SIMT, 1 = tensor core
    // tensor-core path: corpus tiles [tile_b[p], tile_b[p+1]) are searched by pass p with thresholds primed
    // from all earlier passes (see make_plan)
    int n_passes;
    int tile_b[VQ_MAX_PASSES + 1];
    int stride[VQ_MAX_PASSES];
    int ratio;
    vq::TcPlan tc[VQ_MAX_PASSES];
    vq::F32Plan f32[VQ_MAX_PASSES];
    int pass_sets;  // largest number of candidate-list sets any pass writes (set 0 is the running merge)
    int nq_pad;
    size_t off_q, off_lists, off_counts, off_thr, off_prog, total;
    int max_items;
};

// Threshold-priming pass schedule (in tiles).  Pass 0 takes every S0-th tile of the corpus (S0 = ratio^m, at most
// `first_cap` tiles: every row of the unprimed pass becomes a candidate), pass p the multiples of S0 / ratio^p that
// the sparser passes did not take, the last pass (stride 1) everything that is left: ~(1 - 1/ratio) of the corpus.
// Every pass therefore samples the WHOLE corpus, so the primed threshold does not depend on the row order (a
// prefix sample made a corpus sorted by relevance 2.9x slower), and the dominant last pass always starts from a
// threshold computed over 1/ratio of the rows and appends only ~ratio * k candidates per query.
// tile_b[] holds cumulative virtual tile counts; stride[] the S_p.
static int pass_schedule(int n_tiles, int first_cap, int ratio, int* tile_b, int* stride) {
    int m = 0;
    long long s0 = 1;
    while (m < VQ_MAX_PASSES - 1 && (n_tiles + s0 - 1) / s0 > first_cap) {
        s0 *= ratio;
        m++;
    }
    tile_b[0] = 0;
    long long prev = 0;
    long long sp = s0;
    for (int p = 0; p <= m; p++) {
        const long long c = n_tiles > 0 ? (n_tiles + sp - 1) / sp : 0;  // multiples of sp in [0, n_tiles)
        stride[p] = (int)sp;
        tile_b[p + 1] = tile_b[p] + (int)(c - prev);
        prev = c;
        sp /= ratio;
    }
    return m + 1;
}

static void make_plan(const vq_index* idx, int nq, SearchPlan* sp) {
    const long long n = (long long)idx->ntotal;
    sp->n_passes = 1;
    sp->max_items = 0;
    if (idx->dtype == VQ_F32) {
        sp->kind = 0;
        // same threshold-priming schedule as the tensor-core path (see below), in tiles of 128 rows
        const int n_tiles = (int)((n + 127) / 128);
        const int m_tiles = (nq + 127) / 128;
        int first = (idx->n_sms * 2) / m_tiles;
        if (first < 1) first = 1;
        if (first > 56) first = 56;
        sp->ratio = 8;
        const int np = pass_schedule(n_tiles, first, sp->ratio, sp->tile_b, sp->stride);
        sp->n_passes = np;
        sp->pass_sets = 0;
        for (int p = 0; p < np; p++) {
            vq::f32_plan(nq, (long long)(sp->tile_b[p + 1] - sp->tile_b[p]) * 128, idx->n_sms, &sp->f32[p]);
            if (sp->f32[p].n_chunks > sp->pass_sets) sp->pass_sets = sp->f32[p].n_chunks;
        }
        sp->nq_pad = sp->f32[0].nq_pad;
    } else {
        int variant = idx->variant;
        if (variant == VQ_VARIANT_AUTO) variant = (nq <= 128) ? VQ_VARIANT_TC_CG1 : VQ_VARIANT_TC_CG2;
        {
            sp->kind = 1;
            const int cg = variant == VQ_VARIANT_TC_CG2 ? 2 : 1;
            // Threshold priming.  A query's k-th best score over ANY subset of the corpus is a lower bound of its
            // final k-th best score, so the corpus is searched in passes of geometrically growing size (x ratio), each
            // a strided sample of the whole corpus (pass_schedule above): the first pass has about one tile per work
            // item (nothing to prune yet), every later pass starts just below the k-th score of everything before it
            // and rejects almost every 32-score block with a single compare; it appends only ~ratio * k candidates
            // per query.  The handed-down threshold is the next float below the k-th score (merge_lists_kernel), so
            // an equal score with a lower id still gets in and the result stays exact.
            const int n_tiles = (int)((n + 255) / 256);
            const int m_tiles = (nq + 128 * cg - 1) / (128 * cg);
            static int env_first = -1, env_ratio = -1;  // experiment knobs: VQ_FIRST_TILES, VQ_PASS_RATIO
            if (env_first < 0) {
                const char* a = getenv("VQ_FIRST_TILES");
                const char* b2 = getenv("VQ_PASS_RATIO");
                env_first = a ? atoi(a) : 0;
                env_ratio = b2 ? atoi(b2) : 0;
            }
            int ratio = env_ratio > 1 ? env_ratio : 8;  // measured: 6..8 beats 16 by 2-4 % (lighter last pass, one more small pass)
            int first = (idx->n_sms / cg) / m_tiles;
            if (first < 1) first = 1;
            if (first > 28) first = 28;  // every row of the first pass becomes a candidate: keep its merge one 8192-key sort
            if (env_first > 0) first = env_first;
            // (Fewer, larger passes for HBM-bound batches were measured and lose: a pass boundary costs ~30 us of fixed
            // latency, but the one-CTA-per-query merge of a 7,000-row unprimed pass costs 51 us and a CTA-wide shared-memory
            // sort is slower still — profiles/README.md, "latency regime".  Ratio 8 stays the schedule for every batch size.)
            sp->ratio = ratio;
            const int np = pass_schedule(n_tiles, first, ratio, sp->tile_b, sp->stride);
            sp->n_passes = np;
            sp->pass_sets = 0;
            for (int p = 0; p < np; p++) {
                vq::tc_plan(nq, (long long)(sp->tile_b[p + 1] - sp->tile_b[p]) * 256, idx->n_sms, cg, &sp->tc[p]);
                if (sp->tc[p].list_sets > sp->pass_sets) sp->pass_sets = sp->tc[p].list_sets;
                if (sp->tc[p].m_tiles * sp->tc[p].n_chunks > sp->max_items) sp->max_items = sp->tc[p].m_tiles * sp->tc[p].n_chunks;
            }
            sp->nq_pad = sp->tc[0].nq_pad;
        }
    }
    size_t off = 0;
    sp->off_q = off;
    off += align256(storage_bytes(idx, (uint64_t)(nq > sp->nq_pad ? nq : sp->nq_pad)));
    sp->off_lists = off;
    off += align256((size_t)(1 + sp->pass_sets) * sp->nq_pad * VQ_CAP * sizeof(uint64_t));
    sp->off_counts = off;
    off += align256((size_t)(1 + sp->pass_sets) * sp->nq_pad * sizeof(int));
    sp->off_thr = off;
    off += align256((size_t)sp->nq_pad * sizeof(float));
    sp->off_prog = off;
    off += align256((size_t)sp->max_items * sizeof(int));
    sp->total = off;
}

// extra workspace of banded searches: one round's [nq][128] results + the per-query bound
static size_t band_bytes(int nq, int nq_pad) {
    return align256((size_t)nq * VQ_KMAX * sizeof(float)) + align256((size_t)nq * VQ_KMAX * sizeof(long long)) +
           align256((size_t)nq_pad * sizeof(float)) + align256((size_t)nq_pad * sizeof(uint32_t));
}

size_t vq_index_workspace_bytes(const vq_index* idx, int nq, int k) {
    if (!idx || nq <= 0) return 0;
    SearchPlan sp;
    make_plan(idx, nq, &sp);
    return sp.total + (k > VQ_KMAX ? band_bytes(nq, sp.nq_pad) : 0);
}

// One full search of at most VQ_KMAX results per query (all priming passes + merges).  With ub_score/ub_row set,
// only candidates ranking strictly after that per-query bound are taken (banded rounds for k > VQ_KMAX).
static int search_round(vq_index* idx, const SearchPlan& sp, uint8_t* ws, const void* qprep, int nq, int k, const float* ub_score,
                        const uint32_t* ub_row, long long id_offset, float* out_scores, long long* out_ids, cudaStream_t s) {
    uint64_t* lists = (uint64_t*)(ws + sp.off_lists);
    int* counts = (int*)(ws + sp.off_counts);
    float* thr = (float*)(ws + sp.off_thr);
    int* progress = sp.max_items ? (int*)(ws + sp.off_prog) : nullptr;
    const long long ntot = (long long)idx->ntotal;
    int rc = 0;
    if (ntot == 0) {
        cudaError_t e = cudaMemsetAsync(counts, 0, (size_t)(1 + sp.pass_sets) * sp.nq_pad * sizeof(int), s);
        if (e != cudaSuccess) return -(int)e;
        rc = vq::merge_lists_launch(lists, counts, 0, 1, sp.pass_sets, sp.nq_pad, nq, k, id_offset, nullptr, out_scores, out_ids, 512, s);
        idx->launches += 1;
    } else if (sp.kind == 1) {
        // Latency regime (one query tile, one CTA per SM, a handful of merge CTAs): chain the kernels of a search with
        // programmatic dependent launches — the merge of pass p becomes resident while the contraction runs, and the
        // contraction of pass p+1 starts streaming tiles while that merge runs (only its epilogue waits for the
        // thresholds).  Needs every merge CTA resident at once (they wait inside the kernel): few queries only.
        static int no_pdl = -1;
        if (no_pdl < 0) no_pdl = getenv("VQ_NO_PDL") ? 1 : 0;
        const bool pdl = !no_pdl && sp.tc[0].cg == 1 && nq <= 128;
        for (int p = 0; p < sp.n_passes && !rc; p++) {
            const bool last = p == sp.n_passes - 1;
            vq::TcPass pass = {sp.tile_b[p], sp.tile_b[p + 1], 1, sp.stride[p], p ? sp.ratio : 0, p ? thr : nullptr, 1, ub_score, ub_row,
                               pdl ? (p ? 2 : 1) : 0};
            rc = vq::tc_topk_launch(qprep, nq, idx->ld, idx->data, ntot, idx->ld, idx->kd, idx->kdtype, k, sp.tc[p], pass, lists, counts, progress, s);
            // first pass: every row is a candidate; primed passes add ~ratio * k candidates per query
            const int expected = p == 0 ? (sp.tile_b[1] - sp.tile_b[0]) * 256 + k : ((sp.ratio + 2) * k > 2048 ? (sp.ratio + 2) * k : 2048);  // ~ (ratio + 1) * k; more takes gather rounds
            if (!rc)
                rc = vq::merge_lists_launch(lists, counts, p > 0, 1, sp.tc[p].list_sets, sp.nq_pad, nq, k, id_offset, last ? nullptr : thr,
                                            last ? out_scores : nullptr, last ? out_ids : nullptr, expected, s, pdl ? 1 : 0);
            idx->launches += 2;
        }
    } else if (sp.kind == 0) {
        for (int p = 0; p < sp.n_passes && !rc; p++) {
            const bool last = p == sp.n_passes - 1;
            rc = vq::f32_topk_launch((const float*)qprep, nq, (const float*)idx->data, ntot, idx->d, k, sp.f32[p],
                                     sp.tile_b[p], sp.tile_b[p + 1], sp.stride[p], p ? sp.ratio : 0, p ? thr : nullptr, ub_score, ub_row,
                                     lists + (size_t)sp.nq_pad * VQ_CAP, counts + sp.nq_pad, s);
            const int expected = p == 0 ? (sp.tile_b[1] - sp.tile_b[0]) * 128 + k : 2048;
            if (!rc)
                rc = vq::merge_lists_launch(lists, counts, p > 0, 1, sp.f32[p].n_chunks, sp.nq_pad, nq, k, id_offset, last ? nullptr : thr,
                                            last ? out_scores : nullptr, last ? out_ids : nullptr, expected, s);
            idx->launches += 2;
        }
    }
    return rc;
}

#define VQ_KMAX_BANDED 2048

int vq_index_search(vq_index* idx, const void* q, int nq, int q_dtype, int l2norm_q, int k, float* out_scores, int64_t* out_ids,
                    void* workspace, size_t workspace_bytes, void* stream) {
    if (!idx || nq < 0 || k < 1 || !dtype_ok(q_dtype)) return fail(VQ_ERR_INVALID, "vq_index_search: bad arguments (nq=%d k=%d)", nq, k);
    if (nq == 0) return VQ_OK;
    if (!q || !out_scores || !out_ids || !workspace) return fail(VQ_ERR_INVALID, "vq_index_search: null pointer");
    if (k > VQ_KMAX_BANDED) return fail(VQ_ERR_UNSUPPORTED, "vq_index_search: k=%d > %d is not served", k, VQ_KMAX_BANDED);
    cudaStream_t s = (cudaStream_t)stream;
    VQ_ON_DEVICE(idx->device);
    SearchPlan sp;
    make_plan(idx, nq, &sp);
    const size_t need = sp.total + (k > VQ_KMAX ? band_bytes(nq, sp.nq_pad) : 0);
    if (workspace_bytes < need) return fail(VQ_ERR_NOMEM, "vq_index_search: workspace %zu < %zu bytes", workspace_bytes, need);
    uint8_t* ws = (uint8_t*)workspace;
    void* qprep = ws + sp.off_q;
    // rows [nq, nq_pad) of the prepared query matrix are real zero rows (written by the same kernel), so the tensor-core
    // path never depends on TMA out-of-bounds fill for the query operand
    int rc = to_storage(idx, q, q_dtype, qprep, 0, nq, l2norm_q ? 1 : 0, 1, s, sp.nq_pad);
    if (rc) return fail(VQ_ERR_CUDA, "vq_index_search: query prep launch failed (%d)", rc);
    idx->launches += 1;
    const bool timed = idx->timing && idx->n_timed < VQ_TIMING_SLOTS && idx->ntotal;
    if (timed) CUDA_TRY(cudaEventRecord(idx->ev0[idx->n_timed], s));
    if (k <= VQ_KMAX) {
        rc = search_round(idx, sp, ws, qprep, nq, k, nullptr, nullptr, idx->id_offset, out_scores, (long long*)out_ids, s);
    } else {
        // k > 128: rounds of up to 128 results; round r takes what ranks strictly after the last entry of round r-1
        uint8_t* band = ws + sp.total;
        float* tmp_s = (float*)band;
        long long* tmp_i = (long long*)(band + align256((size_t)nq * VQ_KMAX * sizeof(float)));
        float* ub_score = (float*)((uint8_t*)tmp_i + align256((size_t)nq * VQ_KMAX * sizeof(long long)));
        uint32_t* ub_row = (uint32_t*)((uint8_t*)ub_score + align256((size_t)sp.nq_pad * sizeof(float)));
        for (int col0 = 0; col0 < k && !rc; col0 += VQ_KMAX) {
            const int kr = k - col0 < VQ_KMAX ? k - col0 : VQ_KMAX;
            rc = search_round(idx, sp, ws, qprep, nq, kr, col0 ? ub_score : nullptr, col0 ? ub_row : nullptr, 0, tmp_s, tmp_i, s);
            if (!rc)
                rc = vq::band_scatter_launch(tmp_s, tmp_i, nq, kr, k, col0, idx->id_offset, out_scores, (long long*)out_ids, ub_score, ub_row, s);
            idx->launches += 1;
        }
    }
    if (rc == -3) return fail(VQ_ERR_UNSUPPORTED, "vq_index_search: variant does not support k > %d", VQ_KMAX);
    if (rc) return fail(VQ_ERR_CUDA, "vq_index_search: kernel launch failed (%d: %s)", rc, rc < 0 && rc > -1000 ? cudaGetErrorString((cudaError_t)(-rc)) : "driver");
    if (timed) {
        CUDA_TRY(cudaEventRecord(idx->ev1[idx->n_timed], s));
        idx->n_timed++;
    }
    return VQ_OK;
}

int vq_index_search_host(vq_index* idx, const float* q_host, int nq, int l2norm_q, int k, float* out_scores_host,
                         int64_t* out_ids_host) {
    if (!idx || nq < 0 || k < 1) return fail(VQ_ERR_INVALID, "vq_index_search_host: bad arguments");
    if (nq == 0) return VQ_OK;
    if (!q_host || !out_scores_host || !out_ids_host) return fail(VQ_ERR_INVALID, "vq_index_search_host: null pointer");
    VQ_ON_DEVICE(idx->device);
    SearchPlan sp;
    make_plan(idx, nq, &sp);
    const size_t q_bytes = align256((size_t)nq * idx->d * sizeof(float));
    const size_t s_bytes = align256((size_t)nq * k * sizeof(float));
    const size_t i_bytes = align256((size_t)nq * k * sizeof(int64_t));
    const size_t ws_need = vq_index_workspace_bytes(idx, nq, k);
    const size_t need = q_bytes + s_bytes + i_bytes + ws_need;
    if (idx->hws_bytes < need) {
        if (idx->hws) cudaFree(idx->hws);
        idx->hws = nullptr;
        idx->hws_bytes = 0;
        cudaError_t e = cudaMalloc(&idx->hws, need);
        if (e != cudaSuccess) return fail(VQ_ERR_NOMEM, "vq_index_search_host: cudaMalloc(%zu) failed: %s", need, cudaGetErrorString(e));
        idx->hws_bytes = need;
    }
    uint8_t* base = (uint8_t*)idx->hws;
    float* dq = (float*)base;
    float* ds = (float*)(base + q_bytes);
    int64_t* di = (int64_t*)(base + q_bytes + s_bytes);
    void* ws = base + q_bytes + s_bytes + i_bytes;
    cudaStream_t s = 0;
    CUDA_TRY(cudaMemcpyAsync(dq, q_host, (size_t)nq * idx->d * sizeof(float), cudaMemcpyHostToDevice, s));
    int rc = vq_index_search(idx, dq, nq, VQ_F32, l2norm_q, k, ds, di, ws, ws_need, (void*)s);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(out_scores_host, ds, (size_t)nq * k * sizeof(float), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(out_ids_host, di, (size_t)nq * k * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return VQ_OK;
}

int vq_prep_rows(const void* src, int src_dtype, void* dst, int dst_dtype, uint64_t n, int d, int l2norm, void* stream) {
    if ((!src || !dst) && n) return fail(VQ_ERR_INVALID, "vq_prep_rows: null pointer");
    if (!dtype_ok(src_dtype) || !dtype_ok(dst_dtype) || d <= 0) return fail(VQ_ERR_INVALID, "vq_prep_rows: bad arguments");
    int rc = vq::prep_rows_launch(src, src_dtype, d, dst, dst_dtype, d, (long long)n, d, l2norm, (cudaStream_t)stream);
    if (rc) return fail(VQ_ERR_CUDA, "vq_prep_rows: launch failed (%d)", rc);
    return VQ_OK;
}

int vq_merge_topk(const float* scores, const int64_t* ids, int n_lists, int nq, int k_in, int k_out, float* out_scores,
                  int64_t* out_ids, void* stream) {
    if (!scores || !ids || !out_scores || !out_ids || n_lists < 1 || nq < 0) return fail(VQ_ERR_INVALID, "vq_merge_topk: bad arguments");
    int rc = vq::merge_topk_launch(scores, (const long long*)ids, (long long)nq * k_in, (long long)nq * k_in, n_lists, nq, k_in, k_out, out_scores,
                                   (long long*)out_ids, (cudaStream_t)stream);
    if (rc == -2) return fail(VQ_ERR_UNSUPPORTED, "vq_merge_topk: k_in=%d k_out=%d not supported", k_in, k_out);
    if (rc) return fail(VQ_ERR_CUDA, "vq_merge_topk: launch failed (%d)", rc);
    return VQ_OK;
}

int vq_merge_topk_strided(const float* scores, const int64_t* ids, int64_t list_stride_scores, int64_t list_stride_ids, int n_lists, int nq,
                          int k_in, int k_out, float* out_scores, int64_t* out_ids, void* stream) {
    if (!scores || !ids || !out_scores || !out_ids || n_lists < 1 || nq < 0) return fail(VQ_ERR_INVALID, "vq_merge_topk_strided: bad arguments");
    int rc = vq::merge_topk_launch(scores, (const long long*)ids, list_stride_scores, list_stride_ids, n_lists, nq, k_in, k_out, out_scores,
                                   (long long*)out_ids, (cudaStream_t)stream);
    if (rc == -2) return fail(VQ_ERR_UNSUPPORTED, "vq_merge_topk_strided: k_in=%d k_out=%d not supported", k_in, k_out);
    if (rc) return fail(VQ_ERR_CUDA, "vq_merge_topk_strided: launch failed (%d)", rc);
    return VQ_OK;
}

int vq_expand_hits(const float* scores, const int64_t* ids, int nq, int k_in, const int64_t* map_offsets, const int64_t* map_targets,
                   int many2one_max, int k, int out_cap, double* out_scores, int64_t* out_ids, int32_t* out_counts, int32_t* overflow_flag,
                   void* stream) {
    if (!scores || !ids || !out_scores || !out_ids || !out_counts || !overflow_flag || nq < 0 || k_in < 1 || k < 1 || out_cap < 1)
        return fail(VQ_ERR_INVALID, "vq_expand_hits: bad arguments");
    if ((map_offsets == nullptr) != (map_targets == nullptr)) return fail(VQ_ERR_INVALID, "vq_expand_hits: offsets and targets go together");
    int rc = vq::expand_hits_launch(scores, (const long long*)ids, nq, k_in, (const long long*)map_offsets, (const long long*)map_targets,
                                    many2one_max, k, out_cap, out_scores, (long long*)out_ids, out_counts, overflow_flag, (cudaStream_t)stream);
    if (rc == -2) return fail(VQ_ERR_UNSUPPORTED, "vq_expand_hits: out_cap=%d too large for shared memory", out_cap);
    if (rc) return fail(VQ_ERR_CUDA, "vq_expand_hits: launch failed (%d)", rc);
    return VQ_OK;
}

// ---------------------------------------------------------------- cross-GPU exchange (vq_xchg.cu)
struct vq_xchg {
    int device, rank, world;
    size_t slot_bytes;
    uint8_t* window;        // [XchgControl | slot 0 | slot 1], cudaMalloc'ed here
    void* mapped[VQ_XCHG_MAX_RANKS];  // cudaIpcOpenMemHandle results (null for the own rank)
    int attached;
    vq::XchgPeers peers;
};

static size_t xchg_ctl_bytes() { return align256(sizeof(vq::XchgControl)); }

int vq_xchg_create(vq_xchg** out, int device, int rank, int world_size, size_t slot_bytes) {
    if (!out || rank < 0 || world_size < 1 || rank >= world_size || world_size > VQ_XCHG_MAX_RANKS || slot_bytes == 0)
        return fail(VQ_ERR_INVALID, "vq_xchg_create: bad arguments (rank=%d world=%d, at most %d ranks)", rank, world_size, VQ_XCHG_MAX_RANKS);
    VQ_ON_DEVICE(device);
    vq_xchg* x = new vq_xchg();
    x->device = device;
    x->rank = rank;
    x->world = world_size;
    x->slot_bytes = align256(slot_bytes);
    x->attached = 1;
    memset(x->mapped, 0, sizeof(x->mapped));
    memset(&x->peers, 0, sizeof(x->peers));
    const size_t total = xchg_ctl_bytes() + 2 * x->slot_bytes;
    cudaError_t e = cudaMalloc((void**)&x->window, total);
    if (e == cudaSuccess) e = cudaMemset(x->window, 0, total);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        if (x->window) cudaFree(x->window);
        delete x;
        return fail(VQ_ERR_NOMEM, "vq_xchg_create: %s", cudaGetErrorString(e));
    }
    x->peers.ctl[rank] = (vq::XchgControl*)x->window;
    x->peers.slots[rank] = x->window + xchg_ctl_bytes();
    *out = x;
    return VQ_OK;
}

int vq_xchg_export(vq_xchg* x, void* handle_out) {
    if (!x || !handle_out) return fail(VQ_ERR_INVALID, "vq_xchg_export: null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == VQ_XCHG_HANDLE_BYTES, "IPC handle size");
    VQ_ON_DEVICE(x->device);
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, x->window));
    memcpy(handle_out, &h, sizeof(h));
    return VQ_OK;
}

int vq_xchg_attach(vq_xchg* x, int peer_rank, const void* handle) {
    if (!x || !handle || peer_rank < 0 || peer_rank >= x->world) return fail(VQ_ERR_INVALID, "vq_xchg_attach: bad arguments");
    if (peer_rank == x->rank || x->mapped[peer_rank]) return VQ_OK;
    VQ_ON_DEVICE(x->device);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    x->mapped[peer_rank] = p;
    x->peers.ctl[peer_rank] = (vq::XchgControl*)p;
    x->peers.slots[peer_rank] = (uint8_t*)p + xchg_ctl_bytes();
    x->attached++;
    return VQ_OK;
}

size_t vq_xchg_slot_bytes(const vq_xchg* x) { return x ? x->slot_bytes : 0; }

static int xchg_ready(const vq_xchg* x, const char* who) {
    if (!x) return fail(VQ_ERR_INVALID, "%s: null exchange", who);
    if (x->attached != x->world) return fail(VQ_ERR_INVALID, "%s: %d of %d ranks attached", who, x->attached, x->world);
    return VQ_OK;
}

int vq_xchg_publish(vq_xchg* x, const void* src, size_t bytes, void* stream) {
    int rc = xchg_ready(x, "vq_xchg_publish");
    if (rc) return rc;
    if (!src || bytes == 0 || bytes > x->slot_bytes) return fail(VQ_ERR_INVALID, "vq_xchg_publish: %zu bytes do not fit the %zu-byte slot", bytes, x->slot_bytes);
    VQ_ON_DEVICE(x->device);
    rc = vq::xchg_publish_launch(x->peers.ctl[x->rank], x->peers.slots[x->rank], x->slot_bytes, src, bytes, x->peers, x->rank, x->world, (cudaStream_t)stream);
    if (rc) return fail(VQ_ERR_CUDA, "vq_xchg_publish: launch failed (%d)", rc);
    return VQ_OK;
}

int vq_xchg_merge_topk(vq_xchg* x, int nq, int k_in, int k_out, float* out_scores, int64_t* out_ids, void* stream) {
    int rc = xchg_ready(x, "vq_xchg_merge_topk");
    if (rc) return rc;
    if (!out_scores || !out_ids || nq < 0) return fail(VQ_ERR_INVALID, "vq_xchg_merge_topk: bad arguments");
    VQ_ON_DEVICE(x->device);
    rc = vq::xchg_merge_launch(x->peers.ctl[x->rank], x->peers, x->slot_bytes, x->world, nq, k_in, k_out, out_scores, (long long*)out_ids, (cudaStream_t)stream);
    if (rc == -2) return fail(VQ_ERR_UNSUPPORTED, "vq_xchg_merge_topk: k_in=%d k_out=%d world=%d not supported", k_in, k_out, x->world);
    if (rc) return fail(VQ_ERR_CUDA, "vq_xchg_merge_topk: launch failed (%d)", rc);
    return VQ_OK;
}

int vq_xchg_gather(vq_xchg* x, size_t bytes, void* dst, void* stream) {
    int rc = xchg_ready(x, "vq_xchg_gather");
    if (rc) return rc;
    if (!dst || bytes == 0) return fail(VQ_ERR_INVALID, "vq_xchg_gather: bad arguments");
    VQ_ON_DEVICE(x->device);
    rc = vq::xchg_gather_launch(x->peers.ctl[x->rank], x->peers, x->slot_bytes, x->world, bytes, dst, (cudaStream_t)stream);
    if (rc == -2) return fail(VQ_ERR_INVALID, "vq_xchg_gather: %zu bytes: must be a multiple of 16 and fit the slot", bytes);
    if (rc) return fail(VQ_ERR_CUDA, "vq_xchg_gather: launch failed (%d)", rc);
    return VQ_OK;
}

void vq_xchg_destroy(vq_xchg* x) {
    if (!x) return;
    DeviceGuard guard(x->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < x->world; r++)
        if (x->mapped[r]) cudaIpcCloseMemHandle(x->mapped[r]);
    if (x->window) cudaFree(x->window);
    delete x;
}

size_t vq_fuse_workspace_bytes(int n_runs, int nq, int k_in) {
    if (n_runs < 1 || nq < 1 || k_in < 1) return 0;
    return vq::fuse_workspace_bytes(n_runs, nq, k_in);
}

int vq_fuse(const double* scores, const int64_t* ids, const double* weights, int n_runs, int nq, int k_in, int defmin, int norm,
            double* out_scores, int64_t* out_ids, int32_t* out_counts, void* workspace, size_t workspace_bytes, void* stream) {
    if (!scores || !ids || !weights || !out_scores || !out_ids || !out_counts || !workspace)
        return fail(VQ_ERR_INVALID, "vq_fuse: null pointer");
    vq::FuseParams fp;
    fp.n_runs = n_runs;
    fp.nq = nq;
    fp.k_in = k_in;
    fp.u_max = n_runs * k_in;
    fp.defmin = defmin;
    fp.norm = norm;
    int rc = vq::fuse_launch(scores, (const long long*)ids, weights, fp, out_scores, (long long*)out_ids, out_counts, workspace,
                             workspace_bytes, (cudaStream_t)stream);
    if (rc == -3) return fail(VQ_ERR_UNSUPPORTED, "vq_fuse: n_runs=%d k_in=%d norm=%d outside the supported range (n_runs*k_in <= 16384, n_runs <= 16)", n_runs, k_in, norm);
    if (rc == -4) return fail(VQ_ERR_NOMEM, "vq_fuse: workspace too small");
    if (rc) return fail(VQ_ERR_CUDA, "vq_fuse: launch failed (%d)", rc);
    return VQ_OK;
}

int vq_inbatch_logprobs(const void* q, const void* c, int dtype, int n_q, int n_c, int d, const int64_t* labels, int64_t ignore_index,
                        float* log_probs, float* nll, int32_t* rank, void* stream) {
    if (!q || !c || !labels || !nll || !rank || n_q < 0 || n_c < 1 || d < 1 || !dtype_ok(dtype))
        return fail(VQ_ERR_INVALID, "vq_inbatch_logprobs: bad arguments");
    int rc = vq::inbatch_launch(q, c, dtype, n_q, n_c, d, (const long long*)labels, (long long)ignore_index, log_probs, nll, rank, (cudaStream_t)stream);
    if (rc == -2) return fail(VQ_ERR_UNSUPPORTED, "vq_inbatch_logprobs: d + n_c = %d floats exceed the shared memory of one CTA", d + n_c);
    if (rc) return fail(VQ_ERR_CUDA, "vq_inbatch_logprobs: launch failed (%d)", rc);
    return VQ_OK;
}

int vq_mark_relevant(const int64_t* ids, int nq, int k, const int64_t* rel_offsets, const int64_t* rel_ids, uint8_t* is_relevant, void* stream) {
    if (!ids || !rel_offsets || !is_relevant || nq < 0 || k < 1) return fail(VQ_ERR_INVALID, "vq_mark_relevant: bad arguments");
    int rc = vq::mark_relevant_launch((const long long*)ids, nq, k, (const long long*)rel_offsets, (const long long*)rel_ids, is_relevant, (cudaStream_t)stream);
    if (rc) return fail(VQ_ERR_CUDA, "vq_mark_relevant: launch failed (%d)", rc);
    return VQ_OK;
}

size_t vq_fuse_grid_workspace_bytes(int n_runs, int nq, int k_in, int n_weights) {
    if (n_runs < 1 || nq < 1 || k_in < 1 || n_weights < 1) return 0;
    return vq::fuse_grid_workspace_bytes(n_runs, nq, k_in, n_weights);
}

int vq_fuse_grid(const double* scores, const int64_t* ids, int n_runs, int nq, int k_in, int defmin, int norm, const double* weights,
                 int n_weights, const int64_t* rel, int max_rel, const int32_t* n_rel_total, int n_judged, int metric, int cut,
                 double* out_metric, int32_t* out_best, void* workspace, size_t workspace_bytes, void* stream) {
    if (!scores || !ids || !weights || !rel || !n_rel_total || !out_metric || !out_best || !workspace)
        return fail(VQ_ERR_INVALID, "vq_fuse_grid: null pointer");
    vq::FuseParams fp;
    fp.n_runs = n_runs;
    fp.nq = nq;
    fp.k_in = k_in;
    fp.u_max = n_runs * k_in;
    fp.defmin = defmin;
    fp.norm = norm;
    int rc = vq::fuse_grid_launch(scores, (const long long*)ids, fp, weights, n_weights, (const long long*)rel, max_rel, n_rel_total, n_judged, metric,
                                  cut, out_metric, out_best, workspace, workspace_bytes, (cudaStream_t)stream);
    if (rc == -3) return fail(VQ_ERR_UNSUPPORTED, "vq_fuse_grid: n_runs=%d k_in=%d norm=%d metric=%d outside the supported range", n_runs, k_in, norm, metric);
    if (rc == -4) return fail(VQ_ERR_NOMEM, "vq_fuse_grid: workspace too small");
    if (rc) return fail(VQ_ERR_CUDA, "vq_fuse_grid: launch failed (%d)", rc);
    return VQ_OK;
}

}  // extern "C"
