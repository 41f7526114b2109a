// K1 — fp32 "exact" mode: SIMT FFMA contraction with a FIXED reduction order + fused top-k.
//
// Stands in for faiss IndexFlat(d, METRIC_INNER_PRODUCT)::search on an fp32 corpus
// (meerqat/ir/search.py:143-146 casts queries to float32; datasets/search.py:384).  Every score is
//     acc = fmaf(q[j], x[j], acc),  j = 0 .. d-1 ascending, acc0 = +0
// i.e. bit-for-bit what oracle/flat_ip.c computes, so ids match except at exactly tied scores
// (and those follow the same lower-id-first rule).
//
// 128 (queries) x 128 (corpus rows) tile per CTA, 256 threads, 8x8 outputs per thread, K staged
// through double-buffered shared memory 16 columns at a time (one barrier per K block).  FFMA-bound.
#include "vq_common.cuh"
#include "vq_kernels.h"

namespace vq {
namespace f32 {

constexpr int BM = 128, BN = 128, BK = 16, THREADS = 256;
constexpr int LDS = BM + 4;  // padded row of the k-major smem tiles

__global__ void __launch_bounds__(THREADS, 2)
f32_topk_kernel(const float* __restrict__ Q, int nq, int ld_q, const float* __restrict__ X, long long n_rows, int ld_x, int d, int k,
                int m_tiles, int tile_begin, int n_tiles, int map_stride, int map_ratio, int n_chunks, int tiles_per_chunk, int nq_pad,
                const float* __restrict__ init_thr,
                const float* __restrict__ ub_score, const uint32_t* __restrict__ ub_row, uint64_t* __restrict__ lists,
                int* __restrict__ counts) {
    __shared__ __align__(16) float As[2][BK][LDS];  // double-buffered k-major tiles
    __shared__ __align__(16) float Bs[2][BK][LDS];
    __shared__ float thr_s[BM];
    __shared__ int cnt_s[BM];
    __shared__ float ubs_s[BM];      // exclusive upper bound of banded rounds (k > 128); +inf when unbanded
    __shared__ uint32_t ubr_s[BM];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // 16 x 16 threads, 8 x 8 outputs each as 2 x 2 blocks of 4 x 4: rows {4ty..4ty+3, 64+4ty..}, cols {4tx.., 64+4tx..}
    // (16-byte-strided float4 reads of the B tile are bank-conflict free; the A reads are warp broadcasts)
    const int ty = tid >> 4, tx = tid & 15;
    // global -> smem staging: thread loads float4 (4 consecutive k) of row lr for A and B, twice (rows lr, lr+64)
    const int lr = tid >> 2, lk = (tid & 3) * 4;
    const int n_items = m_tiles * n_chunks;
    const int num_kb = (d + BK - 1) / BK;
    const bool vec_ok = ((ld_q & 3) == 0) && ((ld_x & 3) == 0) && ((d & 3) == 0);

    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int m_tile = item % m_tiles, chunk = item / m_tiles;
        const int q0 = m_tile * BM;
        uint64_t* item_lists = lists + ((size_t)chunk * nq_pad + q0) * VQ_CAP;
        if (tid < BM) {
            cnt_s[tid] = 0;
            // primed threshold (next float below the k-th score of the strided samples the earlier passes searched: those
            // rows may have HIGHER ids, so an equal score must still pass the strict compare) or -FLT_MAX
            thr_s[tid] = (q0 + tid < nq) ? (init_thr ? init_thr[q0 + tid] : -FLT_MAX) : INFINITY;
            ubs_s[tid] = (ub_score && q0 + tid < nq) ? ub_score[q0 + tid] : INFINITY;
            ubr_s[tid] = (ub_row && q0 + tid < nq) ? ub_row[q0 + tid] : 0u;
        }
        __syncthreads();
        const int t0 = tile_begin + chunk * tiles_per_chunk;
        const int t1 = min(n_tiles, t0 + tiles_per_chunk);
        for (int t = t0; t < t1; t++) {
            // virtual tile of the pass -> physical corpus tile (every priming pass samples the whole corpus)
            int jt = t - tile_begin;
            if (map_ratio) jt = (jt / (map_ratio - 1)) * map_ratio + (jt % (map_ratio - 1)) + 1;
            const long long x0 = (long long)jt * map_stride * BN;
            float acc[8][8];
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 8; j++) acc[i][j] = 0.0f;

            float4 ra[2], rb[2];
            auto gload = [&](int kb) {
                const int kk = kb * BK + lk;
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int qr = q0 + lr + h * 64;
                    const long long xr = x0 + lr + h * 64;
                    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
                    if (qr < nq) {
                        const float* src = Q + (size_t)qr * ld_q + kk;
                        if (vec_ok && kk + 3 < d) a = *reinterpret_cast<const float4*>(src);
                        else {
                            if (kk + 0 < d) a.x = src[0];
                            if (kk + 1 < d) a.y = src[1];
                            if (kk + 2 < d) a.z = src[2];
                            if (kk + 3 < d) a.w = src[3];
                        }
                    }
                    if (xr < n_rows) {
                        const float* src = X + (size_t)xr * ld_x + kk;
                        if (vec_ok && kk + 3 < d) b = __ldg(reinterpret_cast<const float4*>(src));
                        else {
                            if (kk + 0 < d) b.x = src[0];
                            if (kk + 1 < d) b.y = src[1];
                            if (kk + 2 < d) b.z = src[2];
                            if (kk + 3 < d) b.w = src[3];
                        }
                    }
                    ra[h] = a;
                    rb[h] = b;
                }
            };
            auto sstore = [&](int buf) {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int r = lr + h * 64;
                    As[buf][lk + 0][r] = ra[h].x; As[buf][lk + 1][r] = ra[h].y; As[buf][lk + 2][r] = ra[h].z; As[buf][lk + 3][r] = ra[h].w;
                    Bs[buf][lk + 0][r] = rb[h].x; Bs[buf][lk + 1][r] = rb[h].y; Bs[buf][lk + 2][r] = rb[h].z; Bs[buf][lk + 3][r] = rb[h].w;
                }
            };
            gload(0);
            sstore(0);
            __syncthreads();
            for (int kb = 0; kb < num_kb; kb++) {
                const int buf = kb & 1;
                if (kb + 1 < num_kb) gload(kb + 1);  // global loads fly over the FFMA block below
#pragma unroll
                for (int kk = 0; kk < BK; kk++) {
                    float a[8], b[8];
                    const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
                    const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
                    const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
                    const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
                    a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
                    b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
                    for (int i = 0; i < 8; i++)
#pragma unroll
                        for (int j = 0; j < 8; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
                }
                // the other buffer was last read one iteration ago, before the barrier below: safe to overwrite
                if (kb + 1 < num_kb) sstore(buf ^ 1);
                __syncthreads();
            }
            // ---- fused top-k: threshold filter -> per-query candidate lists
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int r = (i < 4) ? ty * 4 + i : 64 + ty * 4 + (i - 4);
                const float thr = thr_s[r];
                const float ubs = ubs_s[r];
                const uint32_t ubr = ubr_s[r];
                unsigned long long* list = reinterpret_cast<unsigned long long*>(item_lists + (size_t)r * VQ_CAP);
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const long long xr = x0 + ((j < 4) ? tx * 4 + j : 64 + tx * 4 + (j - 4));
                    const float s = acc[i][j];
                    if (xr < n_rows && s > thr && (s < ubs || (s == ubs && (uint32_t)xr > ubr))) {
                        const int pos = atomicAdd(&cnt_s[r], 1);
                        __stcg(list + pos, (unsigned long long)pack_raw(s, (uint32_t)xr));
                    }
                }
            }
            __syncthreads();
            // a tile adds at most BN = 128 keys to a list: prune whenever fewer than 128 slots remain
            for (int r = warp; r < BM; r += THREADS / 32) {
                const int n = cnt_s[r];
                if (n > VQ_CAP - BN) {
                    float thr_new;
                    const int m = warp_prune_list(item_lists + (size_t)r * VQ_CAP, n, k, lane, thr_new);
                    if (lane == 0) {
                        cnt_s[r] = m;
                        thr_s[r] = fmaxf(thr_s[r], thr_new);  // never loosen a primed threshold
                    }
                }
            }
            // (the __syncthreads at the top of the next K loop / below orders these writes)
        }
        __syncthreads();
        // publish the (unsorted) list lengths; the pass's merge kernel sorts
        if (tid < BM) counts[(size_t)chunk * nq_pad + q0 + tid] = (q0 + tid < nq) ? cnt_s[tid] : 0;
        __syncthreads();
    }
}

}  // namespace f32

void f32_plan(int nq, long long n_rows, int n_sms, F32Plan* plan) {
    const int max_ctas = n_sms * 2;
    const int m_tiles = (nq + f32::BM - 1) / f32::BM;
    const int n_tiles = (int)((n_rows + f32::BN - 1) / f32::BN);
    int best_c = 1;
    double best_eff = 0.0;
    int c_max = n_tiles < 4 * max_ctas ? (n_tiles > 0 ? n_tiles : 1) : 4 * max_ctas;
    if (c_max > 1000) c_max = 1000;  // the merge kernel takes at most 1023 list sets
    for (int c = 1; c <= c_max; c++) {
        const int tpc = (n_tiles + c - 1) / c;
        const int c_eff = (n_tiles + tpc - 1) / tpc;
        const long long items = (long long)m_tiles * c_eff;
        const long long rounds = (items + max_ctas - 1) / max_ctas;
        const double eff = (double)m_tiles * n_tiles / ((double)rounds * tpc * max_ctas);
        if (eff > best_eff + 0.02) { best_eff = eff; best_c = c; }
        if (best_eff > 0.97) break;
    }
    const int tpc = n_tiles > 0 ? (n_tiles + best_c - 1) / best_c : 1;
    plan->m_tiles = m_tiles;
    plan->n_tiles = n_tiles;
    plan->tiles_per_chunk = tpc;
    plan->n_chunks = n_tiles > 0 ? (n_tiles + tpc - 1) / tpc : 1;
    plan->nq_pad = m_tiles * f32::BM;
    const long long items = (long long)m_tiles * plan->n_chunks;
    plan->n_ctas = (int)(items < max_ctas ? items : max_ctas);
    if (plan->n_ctas < 1) plan->n_ctas = 1;
}

int f32_topk_launch(const float* q, int nq, int ld_q, const float* x, long long n_rows, int ld_x, int d, int k,
                    const F32Plan& plan, int tile_begin, int tile_end, int map_stride, int map_ratio, const float* init_thr,
                    const float* ub_score, const uint32_t* ub_row, uint64_t* lists, int* counts, cudaStream_t stream) {
    if (k < 1 || k > VQ_KMAX) return -2;
    if (nq <= 0) return 0;
    f32::f32_topk_kernel<<<plan.n_ctas, f32::THREADS, 0, stream>>>(q, nq, ld_q, x, n_rows, ld_x, d, k, plan.m_tiles, tile_begin, tile_end, map_stride, map_ratio,
                                                                  plan.n_chunks, plan.tiles_per_chunk, plan.nq_pad, init_thr, ub_score, ub_row, lists, counts);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -(int)e;
}

}  // namespace vq
