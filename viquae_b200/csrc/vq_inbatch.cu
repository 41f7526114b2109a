// Adjacent rows of the hot path: what consumes / produces the embeddings next to the search.
//
//  * inbatch_kernel — the bi-encoder's in-batch similarity as the reference evaluates it
//    (meerqat/train/trainee.py:454-458: similarities = Q @ C^T; log_probs = log_softmax(similarities, dim=1);
//    NLLLoss; meerqat/train/metrics.py:22-37 batch_retrieval: rank of the labelled context, MRR, hits@1),
//    fused: the similarity row of a question never leaves shared memory between the contraction, the
//    log-sum-exp, the loss term and the rank count.  Forward only (the evaluation step).
//  * mark_relevant_kernel — hard-negative mining after a (64k-query) search (EXPERIMENTS.rst:78-103: the
//    retrieved ids are split into relevant ones and `search_irrelevant_indices`): per (query, hit) membership in
//    the query's sorted list of relevant ids (CSR), by binary search.
#include "vq_common.cuh"
#include "vq_kernels.h"

namespace vq {
namespace inbatch {

constexpr int THREADS = 256;
constexpr int WARPS = THREADS / 32;

template <typename T> __device__ __forceinline__ float ld_f32(const T* p, int j);
template <> __device__ __forceinline__ float ld_f32<float>(const float* p, int j) { return p[j]; }
template <> __device__ __forceinline__ float ld_f32<__nv_bfloat16>(const __nv_bfloat16* p, int j) { return __bfloat162float(p[j]); }
template <> __device__ __forceinline__ float ld_f32<__half>(const __half* p, int j) { return __half2float(p[j]); }

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float t = __shfl_xor_sync(VQ_FULL_MASK, v, o);
        v = is_max ? fmaxf(v, t) : v + t;
    }
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float r = red[0];
#pragma unroll
    for (int w = 1; w < WARPS; w++) r = is_max ? fmaxf(r, red[w]) : r + red[w];  // fixed order: deterministic
    __syncthreads();
    return r;
}

// one CTA per question i: s[j] = <q_i, c_j> (warp per context, lanes stride over d, fp32 accumulation), then
// lse = max + log(sum exp(s - max)); log_probs[i][j] = s[j] - lse; nll[i] = -log_probs[i][label];
// rank[i] = 1 + #{j : s[j] > s[label] or (s[j] == s[label] and j < label)}; ignored labels -> nll 0, rank 0
template <typename T>
__global__ void __launch_bounds__(THREADS)
inbatch_kernel(const T* __restrict__ Q, const T* __restrict__ C, int n_q, int n_c, int d, const long long* __restrict__ labels,
               long long ignore_index, float* __restrict__ log_probs, float* __restrict__ nll, int* __restrict__ rank) {
    extern __shared__ float sm[];   // [d] question, [n_c] similarities
    __shared__ float red[WARPS];
    float* qv = sm;
    float* s = sm + d;
    const int i = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int j = threadIdx.x; j < d; j += THREADS) qv[j] = ld_f32<T>(Q + (size_t)i * d, j);
    __syncthreads();
    for (int c = warp; c < n_c; c += WARPS) {
        const T* row = C + (size_t)c * d;
        float acc = 0.0f;
        for (int j = lane; j < d; j += 32) acc = fmaf(qv[j], ld_f32<T>(row, j), acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(VQ_FULL_MASK, acc, o);
        if (lane == 0) s[c] = acc;
    }
    __syncthreads();
    float mx = -INFINITY;
    for (int c = threadIdx.x; c < n_c; c += THREADS) mx = fmaxf(mx, s[c]);
    mx = block_reduce(mx, red, true);
    float sum = 0.0f;
    for (int c = threadIdx.x; c < n_c; c += THREADS) sum += expf(s[c] - mx);
    sum = block_reduce(sum, red, false);
    const float lse = mx + logf(sum);
    if (log_probs)
        for (int c = threadIdx.x; c < n_c; c += THREADS) log_probs[(size_t)i * n_c + c] = s[c] - lse;
    const long long label = labels[i];
    const bool ignored = label == ignore_index || label < 0 || label >= n_c;
    float before = 0.0f;
    if (!ignored) {
        const float sl = s[label];
        for (int c = threadIdx.x; c < n_c; c += THREADS) before += (s[c] > sl || (s[c] == sl && c < label)) ? 1.0f : 0.0f;
    }
    before = block_reduce(before, red, false);
    if (threadIdx.x == 0) {
        nll[i] = ignored ? 0.0f : -(s[label] - lse);
        rank[i] = ignored ? 0 : (int)before + 1;
    }
}

// hits: ids [nq, k] (id < 0 = padding); relevant ids of query q: rel_ids[rel_off[q] .. rel_off[q+1]) ascending
__global__ void __launch_bounds__(THREADS)
mark_relevant_kernel(const long long* __restrict__ ids, int nq, int k, const long long* __restrict__ rel_off, const long long* __restrict__ rel_ids,
                     unsigned char* __restrict__ is_rel) {
    const long long n = (long long)nq * k;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int q = (int)(e / k);
        const long long id = ids[e];
        long long lo = rel_off[q], hi = rel_off[q + 1];
        bool found = false;
        while (id >= 0 && lo < hi) {
            const long long mid = (lo + hi) >> 1;
            const long long v = rel_ids[mid];
            if (v == id) { found = true; break; }
            if (v < id) lo = mid + 1; else hi = mid;
        }
        is_rel[e] = found ? 1 : 0;
    }
}

}  // namespace inbatch

int inbatch_launch(const void* q, const void* c, int dtype, int n_q, int n_c, int d, const long long* labels, long long ignore_index,
                   float* log_probs, float* nll, int* rank, cudaStream_t stream) {
    if (n_q <= 0) return 0;
    const size_t smem = (size_t)(d + n_c) * sizeof(float);
    if (smem > 200 * 1024) return -2;
#define VQ_INBATCH_LAUNCH(T)                                                                                                          \
    do {                                                                                                                              \
        static size_t attr_dev[64] = {};                                                                                              \
        int dev_ = 0;                                                                                                                 \
        cudaGetDevice(&dev_);                                                                                                         \
        if (smem > 40 * 1024 && smem > attr_dev[dev_ & 63]) {                                                                         \
            cudaError_t e = cudaFuncSetAttribute(inbatch::inbatch_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            if (e != cudaSuccess) return -(int)e;                                                                                     \
            attr_dev[dev_ & 63] = smem;                                                                                               \
        }                                                                                                                             \
        inbatch::inbatch_kernel<T><<<n_q, inbatch::THREADS, smem, stream>>>((const T*)q, (const T*)c, n_q, n_c, d, labels, ignore_index, \
                                                                        log_probs, nll, rank);                                       \
    } while (0)
    switch (dtype) {
        case VQ_DTYPE_F32: VQ_INBATCH_LAUNCH(float); break;
        case VQ_DTYPE_BF16: VQ_INBATCH_LAUNCH(__nv_bfloat16); break;
        case VQ_DTYPE_F16: VQ_INBATCH_LAUNCH(__half); break;
        default: return -3;
    }
#undef VQ_INBATCH_LAUNCH
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -(int)e;
}

int mark_relevant_launch(const long long* ids, int nq, int k, const long long* rel_off, const long long* rel_ids, unsigned char* is_rel,
                         cudaStream_t stream) {
    const long long n = (long long)nq * k;
    if (n <= 0) return 0;
    long long blocks = (n + inbatch::THREADS - 1) / inbatch::THREADS;
    if (blocks > 148 * 8) blocks = 148 * 8;
    inbatch::mark_relevant_kernel<<<(unsigned)blocks, inbatch::THREADS, 0, stream>>>(ids, nq, k, rel_off, rel_ids, is_rel);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -(int)e;
}

}  // namespace vq
