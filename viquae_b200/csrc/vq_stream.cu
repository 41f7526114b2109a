// K2 — small query batches (1..32): HBM-streaming contraction + fused top-k.
//
// For a handful of queries the contraction is bandwidth-bound (arithmetic intensity = nq flop/B for a
// 16-bit corpus, ridge ~211 queries), so the kernel is built around the corpus stream: every warp
// owns 16 corpus rows at a time and reads them straight from HBM into registers with 128-bit
// coalesced loads (two adjacent 16 B chunks per lane: four lanes cover one full 128 B line of a
// row), bypassing L1.  The loaded registers ARE the A fragments of mma.sync m16n8k16 — the K index
// is permuted consistently on the query side, which the dot product does not see — so no shared
// memory staging of the corpus is needed.  Queries are converted once per CTA into
// fragment-ordered shared memory (conflict-free LDS.128).  fp32 accumulation.
//
// Each CTA walks a contiguous slab of the corpus in rounds of 128 rows (8 warps x 16 rows) and
// keeps one candidate list per query (see vq_common.cuh); lists are pruned between rounds.
// Stands in for faiss IndexFlat::search in the latency regime (interact/system.py:24-40 calls the
// Searcher with a batch of one question).
#include "vq_common.cuh"
#include "vq_kernels.h"

namespace vq {
namespace stream {

constexpr int THREADS = 256, WARPS = 8, ROUND_ROWS = WARPS * 16, KB = 64;

__device__ __forceinline__ uint4 ldg_stream(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

template <bool BF16>
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    if constexpr (BF16) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    } else {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
}

// NT = query tiles of 8 (nq <= 8 * NT)
template <int NT, bool BF16>
__global__ void __launch_bounds__(THREADS, 2)
stream_topk_kernel(const uint16_t* __restrict__ Q, int nq, int ld_q, const uint16_t* __restrict__ X, long long n_rows, int ld_x, int k,
                   long long rows_per_cta, int nq_pad, uint64_t* __restrict__ lists, int* __restrict__ counts) {
    extern __shared__ uint4 qfrag[];  // [num_kb][NT][2][32] fragment-ordered queries
    __shared__ float thr_s[8 * NT];
    __shared__ int cnt_s[8 * NT];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int num_kb = (ld_x + KB - 1) / KB;

    // ---- queries -> fragment order.  chunk (kb, nt, h, lane') = Q[nt*8 + g'][kb*64 + t'*16 + h*8 .. +7]
    for (int i = tid; i < num_kb * NT * 64; i += THREADS) {
        const int l2 = i & 31, h = (i >> 5) & 1, nt = (i >> 6) % NT, kb = (i >> 6) / NT;
        const int qr = nt * 8 + (l2 >> 2), k0 = kb * KB + (l2 & 3) * 16 + h * 8;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (qr < nq && k0 < ld_q) v = *reinterpret_cast<const uint4*>(Q + (size_t)qr * ld_q + k0);
        qfrag[i] = v;
    }
    if (tid < 8 * NT) {
        cnt_s[tid] = 0;
        thr_s[tid] = tid < nq ? -FLT_MAX : INFINITY;
    }
    __syncthreads();

    uint64_t* cta_lists = lists + (size_t)blockIdx.x * nq_pad * VQ_CAP;
    const long long row_begin = (long long)blockIdx.x * rows_per_cta;
    const long long row_end = min(n_rows, row_begin + rows_per_cta);
    float thr[NT][2];
#pragma unroll
    for (int nt = 0; nt < NT; nt++) { thr[nt][0] = thr_s[nt * 8 + 2 * t]; thr[nt][1] = thr_s[nt * 8 + 2 * t + 1]; }

    for (long long r0 = row_begin; r0 < row_end; r0 += ROUND_ROWS) {
        const long long ra = r0 + warp * 16 + g, rb = ra + 8;
        const bool va = ra < row_end, vb = rb < row_end;
        const uint16_t* pa = X + (size_t)(va ? ra : row_begin) * ld_x + t * 16;
        const uint16_t* pb = X + (size_t)(vb ? rb : row_begin) * ld_x + t * 16;
        float acc[NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; nt++)
#pragma unroll
            for (int i = 0; i < 4; i++) acc[nt][i] = 0.0f;

        // software pipeline over 64-wide K blocks: next block's 4 x 16 B are in flight while this one is multiplied
        uint4 a_lo[2], b_lo[2], a_nx[2], b_nx[2];
        auto load_kb = [&](int kb, uint4 (&xa)[2], uint4 (&xb)[2]) {
            const int k0 = kb * KB + t * 16;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const bool in = (k0 + h * 8) < ld_x;
                xa[h] = in ? ldg_stream(pa + kb * KB + h * 8) : make_uint4(0u, 0u, 0u, 0u);
                xb[h] = in ? ldg_stream(pb + kb * KB + h * 8) : make_uint4(0u, 0u, 0u, 0u);
            }
        };
        load_kb(0, a_lo, b_lo);
        for (int kb = 0; kb < num_kb; kb++) {
            if (kb + 1 < num_kb) load_kb(kb + 1, a_nx, b_nx);
#pragma unroll
            for (int nt = 0; nt < NT; nt++) {
                const uint4 q0 = qfrag[((kb * NT + nt) * 2 + 0) * 32 + lane];
                const uint4 q1 = qfrag[((kb * NT + nt) * 2 + 1) * 32 + lane];
                // chunk h holds two k-steps: regs (x, y) then (z, w); rows g (a_lo) and g+8 (b_lo)
                mma16816<BF16>(acc[nt], a_lo[0].x, b_lo[0].x, a_lo[0].y, b_lo[0].y, q0.x, q0.y);
                mma16816<BF16>(acc[nt], a_lo[0].z, b_lo[0].z, a_lo[0].w, b_lo[0].w, q0.z, q0.w);
                mma16816<BF16>(acc[nt], a_lo[1].x, b_lo[1].x, a_lo[1].y, b_lo[1].y, q1.x, q1.y);
                mma16816<BF16>(acc[nt], a_lo[1].z, b_lo[1].z, a_lo[1].w, b_lo[1].w, q1.z, q1.w);
            }
            a_lo[0] = a_nx[0]; a_lo[1] = a_nx[1]; b_lo[0] = b_nx[0]; b_lo[1] = b_nx[1];
        }
        // ---- fused top-k: acc[nt] = {(row g, q 2t), (row g, q 2t+1), (row g+8, q 2t), (row g+8, q 2t+1)}
#pragma unroll
        for (int nt = 0; nt < NT; nt++) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int qi = nt * 8 + 2 * t + (i & 1);
                const bool valid = (i < 2) ? va : vb;
                const float s = acc[nt][i];
                if (valid && s > thr[nt][i & 1]) {
                    const int pos = atomicAdd(&cnt_s[qi], 1);
                    __stcg(reinterpret_cast<unsigned long long*>(cta_lists + (size_t)qi * VQ_CAP) + pos,
                           (unsigned long long)pack_raw(s, (uint32_t)((i < 2) ? ra : rb)));
                }
            }
        }
        __syncthreads();
        // a round adds at most ROUND_ROWS = 128 keys to a list
        for (int qi = warp; qi < 8 * NT; qi += WARPS) {
            const int n = cnt_s[qi];
            if (n > VQ_CAP - ROUND_ROWS) {
                float thr_new;
                const int m = warp_prune_list(cta_lists + (size_t)qi * VQ_CAP, n, k, lane, thr_new);
                if (lane == 0) { cnt_s[qi] = m; thr_s[qi] = thr_new; }
            }
        }
        __syncthreads();
#pragma unroll
        for (int nt = 0; nt < NT; nt++) { thr[nt][0] = thr_s[nt * 8 + 2 * t]; thr[nt][1] = thr_s[nt * 8 + 2 * t + 1]; }
    }
    for (int qi = warp; qi < 8 * NT; qi += WARPS) {
        float thr_new;
        const int m = warp_prune_list(cta_lists + (size_t)qi * VQ_CAP, cnt_s[qi], k, lane, thr_new);
        if (lane == 0) counts[(size_t)blockIdx.x * nq_pad + qi] = qi < nq ? m : 0;
    }
}

template <int NT, bool BF16>
static int launch(const void* q16, int nq, int ld_q, const void* x16, long long n_rows, int ld_x, int k, const StreamPlan& plan,
                  uint64_t* lists, int* counts, cudaStream_t stream) {
    const int num_kb = (ld_x + KB - 1) / KB;
    const size_t smem = (size_t)num_kb * NT * 64 * sizeof(uint4);
    if (smem > 200 * 1024) return -2;
    static size_t attr_set_dev[64] = {};  // cudaFuncSetAttribute is per device
    int dev_ = 0;
    cudaGetDevice(&dev_);
    size_t& attr_set = attr_set_dev[dev_ & 63];
    if (smem > attr_set && smem > 0) {
        cudaError_t e = cudaFuncSetAttribute(stream_topk_kernel<NT, BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return -(int)e;
        attr_set = smem;
    }
    stream_topk_kernel<NT, BF16><<<plan.n_ctas, THREADS, smem, stream>>>((const uint16_t*)q16, nq, ld_q, (const uint16_t*)x16, n_rows, ld_x, k,
                                                                        plan.rows_per_cta, plan.nq_pad, lists, counts);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -(int)e;
}

}  // namespace stream

void stream_plan(int nq, long long n_rows, int n_sms, StreamPlan* plan) {
    const int nt = nq <= 8 ? 1 : (nq <= 16 ? 2 : 4);
    plan->nq_pad = 8 * nt;
    long long ctas = (long long)n_sms * 2;
    const long long rounds = (n_rows + stream::ROUND_ROWS - 1) / stream::ROUND_ROWS;
    if (ctas > rounds) ctas = rounds > 0 ? rounds : 1;
    long long rpc = (n_rows + ctas - 1) / ctas;
    rpc = (rpc + stream::ROUND_ROWS - 1) / stream::ROUND_ROWS * stream::ROUND_ROWS;
    if (rpc < stream::ROUND_ROWS) rpc = stream::ROUND_ROWS;
    plan->rows_per_cta = rpc;
    plan->n_ctas = (int)((n_rows + rpc - 1) / rpc);
    if (plan->n_ctas < 1) plan->n_ctas = 1;
    plan->n_chunks = plan->n_ctas;
}

int stream_topk_launch(const void* q16, int nq, int ld_q, const void* x16, long long n_rows, int ld_x, int d, int dtype, int k,
                       const StreamPlan& plan, uint64_t* lists, int* counts, cudaStream_t stream) {
    (void)d;
    if (k < 1 || k > VQ_KMAX || nq > 32) return -2;
    if (nq <= 0) return 0;
    const bool bf = dtype == VQ_DTYPE_BF16;
    if (plan.nq_pad == 8) return bf ? stream::launch<1, true>(q16, nq, ld_q, x16, n_rows, ld_x, k, plan, lists, counts, stream)
                                    : stream::launch<1, false>(q16, nq, ld_q, x16, n_rows, ld_x, k, plan, lists, counts, stream);
    if (plan.nq_pad == 16) return bf ? stream::launch<2, true>(q16, nq, ld_q, x16, n_rows, ld_x, k, plan, lists, counts, stream)
                                     : stream::launch<2, false>(q16, nq, ld_q, x16, n_rows, ld_x, k, plan, lists, counts, stream);
    return bf ? stream::launch<4, true>(q16, nq, ld_q, x16, n_rows, ld_x, k, plan, lists, counts, stream)
              : stream::launch<4, false>(q16, nq, ld_q, x16, n_rows, ld_x, k, plan, lists, counts, stream);
}

}  // namespace vq
