// K0 — ingest / query preparation: dtype conversion with optional row-wise L2 normalisation.
//
// Replaces the reference's numpy L2norm (meerqat/ir/search.py:43-46: q / ||q||_2, no epsilon, so a
// zero row becomes NaN exactly like the reference) and FAISS's NormalizationTransform for
// "L2norm,Flat" (search.py:230-245), fused with the fp32 -> bf16/fp16 cast of the storage dtype.
// One warp per row, lane-strided (fully coalesced) element loads/stores; HBM-bound and off the search path's
// critical section (runs once per ingest batch and once per query batch).
#include "vq_common.cuh"
#include "vq_kernels.h"

namespace vq {

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }

template <typename S, typename D>
__global__ void __launch_bounds__(256)
prep_rows_kernel(const S* __restrict__ src, long long ld_src, D* __restrict__ dst, long long ld_dst, long long n, int d, int l2norm) {
    const int lane = threadIdx.x & 31;
    const long long warps_total = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < n; row += warps_total) {
        const S* s = src + row * ld_src;
        D* o = dst + row * ld_dst;
        float inv_is_div = 1.0f;  // divisor
        if (l2norm) {
            float acc = 0.0f;
            for (int j = lane; j < d; j += 32) {
                const float v = to_f32<S>(s[j]);
                acc = fmaf(v, v, acc);
            }
#pragma unroll
            for (int o2 = 16; o2 > 0; o2 >>= 1) acc += __shfl_xor_sync(VQ_FULL_MASK, acc, o2);
            inv_is_div = sqrtf(acc);
        }
        for (int j = lane; j < (int)ld_dst; j += 32) {
            float v = 0.0f;
            if (j < d) {
                v = to_f32<S>(s[j]);
                if (l2norm) v = v / inv_is_div;  // 0/0 -> NaN like numpy
            }
            o[j] = from_f32<D>(v);
        }
    }
}

template <typename S>
static int prep_dispatch_dst(const void* src, long long ld_src, void* dst, int dst_dtype, long long ld_dst, long long n, int d, int l2norm,
                             cudaStream_t stream) {
    long long blocks = (n + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) return 0;
    switch (dst_dtype) {
        case VQ_DTYPE_F32:
            prep_rows_kernel<S, float><<<(unsigned)blocks, 256, 0, stream>>>((const S*)src, ld_src, (float*)dst, ld_dst, n, d, l2norm);
            break;
        case VQ_DTYPE_BF16:
            prep_rows_kernel<S, __nv_bfloat16><<<(unsigned)blocks, 256, 0, stream>>>((const S*)src, ld_src, (__nv_bfloat16*)dst, ld_dst, n, d, l2norm);
            break;
        case VQ_DTYPE_F16:
            prep_rows_kernel<S, __half><<<(unsigned)blocks, 256, 0, stream>>>((const S*)src, ld_src, (__half*)dst, ld_dst, n, d, l2norm);
            break;
        default:
            return -2;
    }
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -(int)e;
}

int prep_rows_launch(const void* src, int src_dtype, long long ld_src, void* dst, int dst_dtype, long long ld_dst,
                     long long n, int d, int l2norm, cudaStream_t stream) {
    switch (src_dtype) {
        case VQ_DTYPE_F32: return prep_dispatch_dst<float>(src, ld_src, dst, dst_dtype, ld_dst, n, d, l2norm, stream);
        case VQ_DTYPE_BF16: return prep_dispatch_dst<__nv_bfloat16>(src, ld_src, dst, dst_dtype, ld_dst, n, d, l2norm, stream);
        case VQ_DTYPE_F16: return prep_dispatch_dst<__half>(src, ld_src, dst, dst_dtype, ld_dst, n, d, l2norm, stream);
        default: return -2;
    }
}

}  // namespace vq
