// K0 — ingest / query preparation: dtype conversion with optional row-wise L2 normalisation.
//
// Replaces the reference's numpy L2norm (meerqat/ir/search.py:43-46: q / ||q||_2, no epsilon, so a
// zero row becomes NaN exactly like the reference; l2norm == 1, the query side) and FAISS's
// NormalizationTransform for "L2norm,Flat" (search.py:230-245; fvec_renorm_L2 leaves a zero vector
// untouched, so such a corpus row scores 0 instead of becoming unretrievable; l2norm == 2, the ingest
// side), fused with the fp32 -> bf16/fp16 cast of the storage dtype.
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
prep_rows_kernel(const S* __restrict__ src, long long ld_src, D* __restrict__ dst, long long ld_dst, long long n, int d, int l2norm,
                 long long n_pad) {
    const int lane = threadIdx.x & 31;
    const long long warps_total = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < n_pad; row += warps_total) {
        const S* s = src + row * ld_src;
        D* o = dst + row * ld_dst;
        if (row >= n) {  // padding rows of a query operand: real zero rows
            for (int j = lane; j < (int)ld_dst; j += 32) o[j] = from_f32<D>(0.0f);
            continue;
        }
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
            if (l2norm == 2 && inv_is_div == 0.0f) inv_is_div = 1.0f;  // FAISS: zero vectors stay zero
        }
        for (int j = lane; j < (int)ld_dst; j += 32) {
            float v = 0.0f;
            if (j < d) {
                v = to_f32<S>(s[j]);
                if (l2norm) v = v / inv_is_div;  // l2norm == 1: 0/0 -> NaN like numpy
            }
            o[j] = from_f32<D>(v);
        }
    }
}

template <typename S>
static int prep_dispatch_dst(const void* src, long long ld_src, void* dst, int dst_dtype, long long ld_dst, long long n, int d, int l2norm,
                             long long n_pad, cudaStream_t stream) {
    if (n_pad < n) n_pad = n;
    long long blocks = (n_pad + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) return 0;
    switch (dst_dtype) {
        case VQ_DTYPE_F32:
            prep_rows_kernel<S, float><<<(unsigned)blocks, 256, 0, stream>>>((const S*)src, ld_src, (float*)dst, ld_dst, n, d, l2norm, n_pad);
            break;
        case VQ_DTYPE_BF16:
            prep_rows_kernel<S, __nv_bfloat16><<<(unsigned)blocks, 256, 0, stream>>>((const S*)src, ld_src, (__nv_bfloat16*)dst, ld_dst, n, d, l2norm, n_pad);
            break;
        case VQ_DTYPE_F16:
            prep_rows_kernel<S, __half><<<(unsigned)blocks, 256, 0, stream>>>((const S*)src, ld_src, (__half*)dst, ld_dst, n, d, l2norm, n_pad);
            break;
        default:
            return -2;
    }
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -(int)e;
}

// fp32 value -> three bf16 terms with hi + mid + lo == v exactly (each remainder is exact in fp32)
__device__ __forceinline__ void split3(float v, __nv_bfloat16& hi, __nv_bfloat16& mid, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(hi);
    mid = __float2bfloat16_rn(r1);
    const float r2 = r1 - __bfloat162float(mid);
    lo = __float2bfloat16_rn(r2);
}

template <typename S>
__global__ void __launch_bounds__(256)
split_rows_kernel(const S* __restrict__ src, long long ld_src, __nv_bfloat16* __restrict__ dst, int seg, long long n, int d, int l2norm,
                  int query_side, long long n_pad) {
    const int lane = threadIdx.x & 31;
    const long long warps_total = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < n_pad; row += warps_total) {
        const S* s = src + row * ld_src;
        __nv_bfloat16* o = dst + row * 6 * (long long)seg;
        if (row >= n) {
            for (int j = lane; j < 6 * seg; j += 32) o[j] = __float2bfloat16_rn(0.0f);
            continue;
        }
        float div = 1.0f;
        if (l2norm) {  // same arithmetic as prep_rows_kernel: the normalised fp32 values are those of an fp32 index
            float acc = 0.0f;
            for (int j = lane; j < d; j += 32) {
                const float v = to_f32<S>(s[j]);
                acc = fmaf(v, v, acc);
            }
#pragma unroll
            for (int o2 = 16; o2 > 0; o2 >>= 1) acc += __shfl_xor_sync(VQ_FULL_MASK, acc, o2);
            div = sqrtf(acc);
            if (l2norm == 2 && div == 0.0f) div = 1.0f;
        }
        for (int j = lane; j < seg; j += 32) {
            float v = 0.0f;
            if (j < d) {
                v = to_f32<S>(s[j]);
                if (l2norm) v = v / div;
            }
            __nv_bfloat16 hi, mid, lo;
            split3(v, hi, mid, lo);
            if (query_side) {
                o[j] = hi; o[seg + j] = lo; o[2 * seg + j] = mid; o[3 * seg + j] = hi; o[4 * seg + j] = mid; o[5 * seg + j] = hi;
            } else {
                o[j] = lo; o[seg + j] = hi; o[2 * seg + j] = mid; o[3 * seg + j] = mid; o[4 * seg + j] = hi; o[5 * seg + j] = hi;
            }
        }
    }
}

template <typename D>
__global__ void __launch_bounds__(256)
join_rows_kernel(const __nv_bfloat16* __restrict__ src, int seg, D* __restrict__ dst, long long ld_dst, long long n, int d) {
    const int lane = threadIdx.x & 31;
    const long long warps_total = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < n; row += warps_total) {
        const __nv_bfloat16* s = src + row * 6 * (long long)seg;
        D* o = dst + row * ld_dst;
        for (int j = lane; j < d; j += 32) {
            const float lo = __bfloat162float(s[j]), mid = __bfloat162float(s[2 * seg + j]), hi = __bfloat162float(s[5 * seg + j]);
            const float v = hi + (mid + lo);
            o[j] = from_f32<D>(v == 0.0f ? hi : v);  // a zero keeps its sign (hi is then +-0)
        }
    }
}

static unsigned row_blocks(long long n) {
    long long blocks = (n + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    return (unsigned)blocks;
}

int split_rows_launch(const void* src, int src_dtype, long long ld_src, void* dst, int seg, long long n, int d, int l2norm,
                      int query_side, cudaStream_t stream, long long n_pad) {
    if (n_pad < n) n_pad = n;
    if (n_pad <= 0) return 0;
    __nv_bfloat16* o = (__nv_bfloat16*)dst;
    switch (src_dtype) {
        case VQ_DTYPE_F32: split_rows_kernel<float><<<row_blocks(n_pad), 256, 0, stream>>>((const float*)src, ld_src, o, seg, n, d, l2norm, query_side, n_pad); break;
        case VQ_DTYPE_BF16: split_rows_kernel<__nv_bfloat16><<<row_blocks(n_pad), 256, 0, stream>>>((const __nv_bfloat16*)src, ld_src, o, seg, n, d, l2norm, query_side, n_pad); break;
        case VQ_DTYPE_F16: split_rows_kernel<__half><<<row_blocks(n_pad), 256, 0, stream>>>((const __half*)src, ld_src, o, seg, n, d, l2norm, query_side, n_pad); break;
        default: return -2;
    }
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -(int)e;
}

int join_rows_launch(const void* src, int seg, void* dst, int dst_dtype, long long ld_dst, long long n, int d, cudaStream_t stream) {
    if (n <= 0) return 0;
    const __nv_bfloat16* s = (const __nv_bfloat16*)src;
    switch (dst_dtype) {
        case VQ_DTYPE_F32: join_rows_kernel<float><<<row_blocks(n), 256, 0, stream>>>(s, seg, (float*)dst, ld_dst, n, d); break;
        case VQ_DTYPE_BF16: join_rows_kernel<__nv_bfloat16><<<row_blocks(n), 256, 0, stream>>>(s, seg, (__nv_bfloat16*)dst, ld_dst, n, d); break;
        case VQ_DTYPE_F16: join_rows_kernel<__half><<<row_blocks(n), 256, 0, stream>>>(s, seg, (__half*)dst, ld_dst, n, d); break;
        default: return -2;
    }
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -(int)e;
}

// ---- fp32 exact mode: K-MAJOR tiles of 128 rows, [tile][kpad][128] floats (what K1's TMA boxes read, vq_fp32.cu).
// Element (row R, column c) lives at ((R >> 7) * kpad + c) * 128 + (R & 127); columns d..kpad-1 are zero.
// tile_rows: rows [0, n) of src become rows row0 .. row0+n-1 of the tiled matrix (rows [n, n_pad): zero rows); same cast
// and normalisation ARITHMETIC as prep_rows_kernel (lane-strided fmaf sum, xor-shuffle tree), so an fp32 index holds the
// same bits in either layout.  32 rows per CTA pass through a shared-memory transpose: loads coalesced along a row, stores
// 128 bytes of 32 neighbouring rows at one column.
template <typename S>
__global__ void __launch_bounds__(256)
tile_rows_kernel(const S* __restrict__ src, long long ld_src, float* __restrict__ dst, long long row0, int kpad, long long n, int d, int l2norm,
                 long long n_pad) {
    __shared__ float tile[32][33];
    __shared__ float div_s[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long rb = blockIdx.x; rb * 32 < n_pad; rb += gridDim.x) {
        for (int rr = warp; rr < 32; rr += 8) {
            const long long r = rb * 32 + rr;
            float div = 1.0f;
            if (l2norm && r < n) {
                const S* s = src + r * ld_src;
                float acc = 0.0f;
                for (int j = lane; j < d; j += 32) {
                    const float v = to_f32<S>(s[j]);
                    acc = fmaf(v, v, acc);
                }
#pragma unroll
                for (int o2 = 16; o2 > 0; o2 >>= 1) acc += __shfl_xor_sync(VQ_FULL_MASK, acc, o2);
                div = sqrtf(acc);
                if (l2norm == 2 && div == 0.0f) div = 1.0f;  // FAISS: zero vectors stay zero
            }
            if (lane == 0) div_s[rr] = div;
        }
        __syncthreads();
        for (int k0 = 0; k0 < kpad; k0 += 32) {
            for (int rr = warp; rr < 32; rr += 8) {
                const long long r = rb * 32 + rr;
                const int c = k0 + lane;
                float v = 0.0f;
                if (r < n && c < d) {
                    v = to_f32<S>(src[r * ld_src + c]);
                    if (l2norm) v = v / div_s[rr];  // l2norm == 1: 0/0 -> NaN like numpy
                }
                tile[rr][lane] = v;
            }
            __syncthreads();
            const long long r = rb * 32 + lane;
            if (r < n_pad) {
                const long long R = row0 + r;
                float* o = dst + ((R >> 7) * kpad + k0) * 128 + (R & 127);
                for (int kk = warp; kk < 32; kk += 8) o[(long long)kk * 128] = tile[lane][kk];
            }
            __syncthreads();
        }
    }
}

// rows row0 .. row0+n-1 of the tiled matrix -> [n][ld_dst] rows of D (columns < d)
template <typename D>
__global__ void __launch_bounds__(256)
untile_rows_kernel(const float* __restrict__ src, long long row0, int kpad, D* __restrict__ dst, long long ld_dst, long long n, int d) {
    __shared__ float tile[32][33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long rb = blockIdx.x; rb * 32 < n; rb += gridDim.x) {
        for (int k0 = 0; k0 < d; k0 += 32) {
            const long long r = rb * 32 + lane;
            if (r < n) {
                const long long R = row0 + r;
                const float* s = src + ((R >> 7) * kpad + k0) * 128 + (R & 127);
                for (int kk = warp; kk < 32; kk += 8) tile[lane][kk] = s[(long long)kk * 128];
            }
            __syncthreads();
            for (int rr = warp; rr < 32; rr += 8) {
                const long long r2 = rb * 32 + rr;
                const int c = k0 + lane;
                if (r2 < n && c < d) dst[r2 * ld_dst + c] = from_f32<D>(tile[rr][lane]);
            }
            __syncthreads();
        }
    }
}

int tile_rows_launch(const void* src, int src_dtype, long long ld_src, float* dst, long long row0, int kpad, long long n, int d, int l2norm,
                     cudaStream_t stream, long long n_pad) {
    if (n_pad < n) n_pad = n;
    long long blocks = (n_pad + 31) / 32;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) return 0;
    switch (src_dtype) {
        case VQ_DTYPE_F32: tile_rows_kernel<float><<<(unsigned)blocks, 256, 0, stream>>>((const float*)src, ld_src, dst, row0, kpad, n, d, l2norm, n_pad); break;
        case VQ_DTYPE_BF16:
            tile_rows_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, stream>>>((const __nv_bfloat16*)src, ld_src, dst, row0, kpad, n, d, l2norm, n_pad);
            break;
        case VQ_DTYPE_F16: tile_rows_kernel<__half><<<(unsigned)blocks, 256, 0, stream>>>((const __half*)src, ld_src, dst, row0, kpad, n, d, l2norm, n_pad); break;
        default: return -2;
    }
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -(int)e;
}

int untile_rows_launch(const float* src, long long row0, int kpad, void* dst, int dst_dtype, long long ld_dst, long long n, int d, cudaStream_t stream) {
    long long blocks = (n + 31) / 32;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) return 0;
    switch (dst_dtype) {
        case VQ_DTYPE_F32: untile_rows_kernel<float><<<(unsigned)blocks, 256, 0, stream>>>(src, row0, kpad, (float*)dst, ld_dst, n, d); break;
        case VQ_DTYPE_BF16: untile_rows_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, stream>>>(src, row0, kpad, (__nv_bfloat16*)dst, ld_dst, n, d); break;
        case VQ_DTYPE_F16: untile_rows_kernel<__half><<<(unsigned)blocks, 256, 0, stream>>>(src, row0, kpad, (__half*)dst, ld_dst, n, d); break;
        default: return -2;
    }
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -(int)e;
}

int prep_rows_launch(const void* src, int src_dtype, long long ld_src, void* dst, int dst_dtype, long long ld_dst,
                     long long n, int d, int l2norm, cudaStream_t stream, long long n_pad) {
    switch (src_dtype) {
        case VQ_DTYPE_F32: return prep_dispatch_dst<float>(src, ld_src, dst, dst_dtype, ld_dst, n, d, l2norm, n_pad, stream);
        case VQ_DTYPE_BF16: return prep_dispatch_dst<__nv_bfloat16>(src, ld_src, dst, dst_dtype, ld_dst, n, d, l2norm, n_pad, stream);
        case VQ_DTYPE_F16: return prep_dispatch_dst<__half>(src, ld_src, dst, dst_dtype, ld_dst, n, d, l2norm, n_pad, stream);
        default: return -2;
    }
}

}  // namespace vq
