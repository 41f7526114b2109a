// A/B harness for K1 (fp32 exact contraction): runs template variants of f32_topk_kernel on one synthetic shape with the
// thresholds set so high that nothing is appended — times the contraction loop alone.  Test infrastructure.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I include -I viquae_b200/csrc \
//        -o scripts/ubench/k1_variants scripts/ubench/k1_variants.cu -lcuda
#include <cstdio>
#include <vector>
#include <algorithm>
#include "k1_experiments.cuh"
#include "k1_wide.cuh"

using namespace vq;

static int g_reps = 4, g_idx = 0;
static unsigned g_mask = 0xffffffffu;
template <bool PACKED, bool ALIGNED, bool LANES48, bool SWZ, bool STAGE = true, bool DUMMY = false>
static void run(const char* name, const float* q, int nq, const float* x, long long n, int d, const float* thr, uint64_t* lists, int* counts, int sms) {
    if (!((g_mask >> g_idx++) & 1)) return;
    F32Plan plan;
    f32_plan(nq, n, sms, &plan);
    const int n_tiles = (int)((n + 127) / 128);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < g_reps; rep++) {
        cudaEventRecord(e0);
        f32::f32_topk_kernel<PACKED, ALIGNED, LANES48, SWZ, STAGE, DUMMY><<<plan.n_ctas, f32::THREADS>>>(q, nq, d, x, n, d, d, 100, plan.m_tiles, 0, n_tiles, 1, 0, plan.n_chunks,
                                                                                        plan.tiles_per_chunk, plan.nq_pad, thr, nullptr, nullptr, lists, counts);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if ((rep || g_reps == 1) && ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    const double flops = 2.0 * plan.nq_pad * (double)n * d;
    printf("%-40s %8.3f ms  %6.2f TFLOP/s (padded queries)  ctas=%d chunks=%d  %s\n", name, best, flops / best / 1e9, plan.n_ctas, plan.n_chunks,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
}

static void run_tma(const char* name, const float* q, int nq, const float* x, long long n, int d, const float* thr, uint64_t* lists, int* counts, int sms) {
    if (!((g_mask >> g_idx++) & 1)) return;
    F32Plan plan;
    f32_plan(nq, n, sms, &plan);
    const int n_tiles = (int)((n + 127) / 128);
    float* qt;
    cudaMalloc(&qt, f32_kmajor_bytes(plan.nq_pad, d));
    f32_queries_kmajor_launch(q, plan.nq_pad, d, d, qt, 0);
    const char* ev = getenv("VQ_F32_TMA_VARIANT");
    float* xt = nullptr;
    if (ev && atoi(ev) >= 5) {  // k-major tiled copy of the corpus (n must be a multiple of 128 here)
        cudaMalloc(&xt, f32_kmajor_bytes((int)n, d));
        f32_queries_kmajor_launch(x, (int)n, d, d, xt, 0);
        x = xt;
    }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    int rc = 0;
    for (int rep = 0; rep < g_reps; rep++) {
        cudaEventRecord(e0);
        rc |= f32_topk_tma_launch(qt, nq, x, n, d, d, 100, plan, 0, n_tiles, 1, 0, thr, nullptr, nullptr, lists, counts, 0);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if ((rep || g_reps == 1) && ms < best) best = ms;
    }
    cudaError_t e = cudaDeviceSynchronize();
    const double flops = 2.0 * plan.nq_pad * (double)n * d;
    printf("%-40s %8.3f ms  %6.2f TFLOP/s (padded queries)  ctas=%d chunks=%d  rc=%d %s\n", name, best, flops / best / 1e9, plan.n_ctas, plan.n_chunks, rc,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(qt);
    if (xt) cudaFree(xt);
}

static int encode_kmajor(CUtensorMap* m, const void* base, uint64_t tiles, int kpad) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    auto enc = reinterpret_cast<PFN_encodeTiled>(ptr);
    cuuint64_t gdim[2] = {128, tiles * (uint64_t)kpad};
    cuuint64_t gstride[1] = {512};
    cuuint32_t box[2] = {128, 32};
    cuuint32_t estr[2] = {1, 1};
    return (int)enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

static void run_wide(const char* name, const float* q, int nq, const float* x, long long n, int d, const float* thr, int sms) {
    if (!((g_mask >> g_idx++) & 1)) return;
    F32Plan plan;
    f32_plan_wide(nq, n, sms, &plan);
    const int n_tiles = (int)((n + 127) / 128), kpad = (d + 31) / 32 * 32;
    float *qt, *xt; uint64_t* lists; int* counts;
    cudaMalloc(&qt, f32_kmajor_bytes(plan.nq_pad, d));
    cudaMalloc(&xt, f32_kmajor_bytes((int)n, d));
    cudaMalloc(&lists, (size_t)plan.n_chunks * plan.nq_pad * VQ_CAP * 8);
    cudaMalloc(&counts, (size_t)plan.n_chunks * plan.nq_pad * 4);
    f32_queries_kmajor_launch(q, plan.nq_pad, d, d, qt, 0);
    f32_queries_kmajor_launch(x, (int)n, d, d, xt, 0);
    CUtensorMap tq, tx;
    int rc = encode_kmajor(&tq, qt, plan.nq_pad / 128, kpad) | encode_kmajor(&tx, xt, n_tiles, kpad);
    cudaFuncSetAttribute(f32w::f32_topk_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f32w::SMEM_BYTES);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < g_reps; rep++) {
        cudaEventRecord(e0);
        f32w::f32_topk_wide_kernel<<<plan.n_ctas, f32w::THREADS, f32w::SMEM_BYTES>>>(tq, tx, nq, n, d, 100, plan.m_tiles, 0, n_tiles, 1, 0, plan.n_chunks,
                                                                                    plan.tiles_per_chunk, plan.nq_pad, thr, nullptr, nullptr, lists, counts);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if ((rep || g_reps == 1) && ms < best) best = ms;
    }
    cudaError_t e = cudaDeviceSynchronize();
    const double flops = 2.0 * plan.nq_pad * (double)n * d;
    printf("%-40s %8.3f ms  %6.2f TFLOP/s (padded queries)  ctas=%d chunks=%d  rc=%d %s\n", name, best, flops / best / 1e9, plan.n_ctas, plan.n_chunks, rc,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(qt); cudaFree(xt); cudaFree(lists); cudaFree(counts);
}

int main(int argc, char** argv) {
    const int nq = argc > 1 ? atoi(argv[1]) : 1280, d = 768;
    const long long n = argc > 2 ? atoll(argv[2]) : 148LL * 2 * 128 * 16;
    if (argc > 3) g_reps = atoi(argv[3]);
    if (argc > 4) g_mask = strtoul(argv[4], nullptr, 0);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *q, *x, *thr; uint64_t* lists; int* counts;
    cudaMalloc(&q, (size_t)(nq + 128) * d * 4);
    cudaMemset(q, 0, (size_t)(nq + 128) * d * 4);
    cudaMalloc(&x, (size_t)n * d * 4);
    cudaMalloc(&thr, (size_t)(nq + 128) * 4);
    std::vector<float> h((size_t)(1 << 22));
    uint32_t sd = 12345;
    for (auto& v : h) { sd = sd * 1664525u + 1013904223u; v = ((sd >> 8) & 0xffff) / 65536.0f - 0.5f; }
    for (size_t off = 0; off < (size_t)n * d; off += h.size())
        cudaMemcpy(x + off, h.data(), std::min(h.size(), (size_t)n * d - off) * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(q, h.data(), (size_t)nq * d * 4, cudaMemcpyHostToDevice);
    std::vector<float> t(nq + 128, 1e30f);
    cudaMemcpy(thr, t.data(), t.size() * 4, cudaMemcpyHostToDevice);
    F32Plan plan;
    f32_plan(nq, n, sms, &plan);
    cudaMalloc(&lists, (size_t)plan.n_chunks * plan.nq_pad * VQ_CAP * 8);
    cudaMalloc(&counts, (size_t)plan.n_chunks * plan.nq_pad * 4);
    printf("nq=%d n=%lld d=%d sms=%d\n", nq, n, d, sms);
    run<false, false, false, false>("FFMA  generic-loads 2x16 padded (r1)", q, nq, x, n, d, thr, lists, counts, sms);
    run<true, false, false, false>("FFMA2 generic-loads 2x16 padded", q, nq, x, n, d, thr, lists, counts, sms);
    run<false, true, false, false>("FFMA  lean-loads    2x16 padded", q, nq, x, n, d, thr, lists, counts, sms);
    run<true, true, false, false>("FFMA2 lean-loads    2x16 padded", q, nq, x, n, d, thr, lists, counts, sms);
    run<false, true, true, false>("FFMA  lean-loads    4x8  padded", q, nq, x, n, d, thr, lists, counts, sms);
    run<true, true, true, false>("FFMA2 lean-loads    4x8  padded", q, nq, x, n, d, thr, lists, counts, sms);
    run<false, true, false, true>("FFMA  lean-loads    2x16 swizzled", q, nq, x, n, d, thr, lists, counts, sms);
    run<true, true, false, true>("FFMA2 lean-loads    2x16 swizzled", q, nq, x, n, d, thr, lists, counts, sms);
    run<false, true, true, true>("FFMA  lean-loads    4x8  swizzled", q, nq, x, n, d, thr, lists, counts, sms);
    run<true, true, true, true>("FFMA2 lean-loads    4x8  swizzled", q, nq, x, n, d, thr, lists, counts, sms);
    setenv("VQ_F32_TMA_VARIANT", argc > 5 ? argv[5] : "0", 1);
    run_tma("FFMA2 TMA-fed, 3 x 32-column stages", q, nq, x, n, d, thr, lists, counts, sms);
    run_wide("FFMA2 TMA k-major 256x128, 16x8/thread, 1 CTA/SM", q, nq, x, n, d, thr, sms);
    run<false, true, false, false, false>("FFMA  NO staging    2x16 padded", q, nq, x, n, d, thr, lists, counts, sms);
    run<true, true, false, false, false>("FFMA2 NO staging    2x16 padded", q, nq, x, n, d, thr, lists, counts, sms);
    run<false, true, true, true, false>("FFMA  NO staging    4x8  swizzled", q, nq, x, n, d, thr, lists, counts, sms);
    run<true, true, true, true, false>("FFMA2 NO staging    4x8  swizzled", q, nq, x, n, d, thr, lists, counts, sms);
    run<true, true, false, false, false, true>("FFMA2 NO loads, dummy STS 2x16 padded", q, nq, x, n, d, thr, lists, counts, sms);
    run<true, true, false, true, false, true>("FFMA2 NO loads, dummy STS 2x16 swizzled", q, nq, x, n, d, thr, lists, counts, sms);
    return 0;
}
