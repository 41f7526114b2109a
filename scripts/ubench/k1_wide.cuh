// EXPERIMENT (round 2): K1 with 256 x 128 CTA tiles, 16 x 8 outputs per thread, one CTA per SM, 4-stage TMA ring.
// Same k-major operand layout and packed FFMA2 as the product kernel (csrc/vq_fp32.cu); per FMA a quarter fewer LDS
// wavefronts, register writes and refill bytes.  Included by k1_variants.cu after k1_experiments.cuh.
#pragma once

namespace vq {
namespace f32w {

constexpr int BM = 256, BN = 128, BK = 32, STAGES = 4, THREADS = 256, WARPS = THREADS / 32;
constexpr uint32_t A_BYTES = BK * BM * 4, B_BYTES = BK * BN * 4, STAGE_BYTES = A_BYTES + B_BYTES;  // 32 + 16 KB
constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 128;

using f32::tma::smem_u32;
using f32::tma::mbar_init;
using f32::tma::mbar_wait;
using f32::tma::mbar_arrive_expect_tx;
using f32::tma::mbar_arrive;
using f32::tma::tma_load_2d;
using f32::tma::lds128;
using f32::tma::lds128f;
using f32::ffma2;

struct Cursor {
    int item, t, t1, kb;
    int qrow, xrow;
};

__global__ void __launch_bounds__(THREADS, 1)
f32_topk_wide_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_x, int nq, long long n_rows, int d, int k,
                     int m_tiles, int tile_begin, int n_tiles, int map_stride, int map_ratio, int n_chunks, int tiles_per_chunk, int nq_pad,
                     const float* __restrict__ init_thr, const float* __restrict__ ub_score, const uint32_t* __restrict__ ub_row,
                     uint64_t* __restrict__ lists, int* __restrict__ counts) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES];
    __shared__ float thr_s[BM];
    __shared__ int cnt_s[BM];
    __shared__ float ubs_s[BM];
    __shared__ uint32_t ubr_s[BM];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ty = tid >> 4, tx = tid & 15;  // queries {64 g + 4 ty + 0..3, g = 0..3}, corpus rows {4tx.., 64+4tx..}
    const uint32_t ring = (smem_u32(smem_raw) + 127u) & ~127u;
    const int n_items = m_tiles * n_chunks;
    const int num_kb = (d + BK - 1) / BK;
    const int kpad = num_kb * BK;

    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto phys_tile = [&](int t) {
        int jt = t - tile_begin;
        if (map_ratio) jt = (jt / (map_ratio - 1)) * map_ratio + (jt % (map_ratio - 1)) + 1;
        return jt * map_stride;
    };
    Cursor pc = {};
    auto pc_set_item = [&]() {
        for (; pc.item < n_items; pc.item += gridDim.x) {
            const int chunk = pc.item / m_tiles;
            pc.t = tile_begin + chunk * tiles_per_chunk;
            pc.t1 = min(n_tiles, pc.t + tiles_per_chunk);
            if (pc.t < pc.t1) {
                pc.qrow = (pc.item - chunk * m_tiles) * 2 * kpad;  // two 128-query tiles per item
                pc.xrow = phys_tile(pc.t) * kpad;
                return;
            }
        }
    };
    pc.item = blockIdx.x;
    pc_set_item();
    auto pc_advance = [&]() {
        if (++pc.kb < num_kb) return;
        pc.kb = 0;
        if (++pc.t < pc.t1) {
            pc.xrow = phys_tile(pc.t) * kpad;
            return;
        }
        pc.item += gridDim.x;
        pc_set_item();
    };
    uint32_t p_stage = 0, p_phase = 1;
    auto issue = [&]() {
        mbar_wait(&empty_bar[p_stage], p_phase);
        mbar_arrive_expect_tx(&full_bar[p_stage], STAGE_BYTES);
        const uint32_t dst = ring + p_stage * STAGE_BYTES;
        tma_load_2d(&tmap_q, &full_bar[p_stage], dst, 0, pc.qrow + pc.kb * BK);
        tma_load_2d(&tmap_q, &full_bar[p_stage], dst + A_BYTES / 2, 0, pc.qrow + kpad + pc.kb * BK);
        tma_load_2d(&tmap_x, &full_bar[p_stage], dst + A_BYTES, 0, pc.xrow + pc.kb * BK);
    };
    auto p_next = [&]() {
        if (++p_stage == STAGES) {
            p_stage = 0;
            p_phase ^= 1;
        }
        pc_advance();
    };
    for (int i = 0; i < STAGES - 1; i++) {
        if (pc.item < n_items) {
            if (tid == 0) issue();
            p_next();
        }
    }

    uint32_t it = 0;
    uint32_t c_stage = 0, c_phase = 0;
    const uint32_t a_off = (uint32_t)ty * 16u, b_off = A_BYTES + (uint32_t)tx * 16u;

    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int m_tile = item % m_tiles, chunk = item / m_tiles;
        const int q0 = m_tile * BM;
        uint64_t* item_lists = lists + ((size_t)chunk * nq_pad + q0) * VQ_CAP;
        const int t0 = tile_begin + chunk * tiles_per_chunk;
        const int t1 = min(n_tiles, t0 + tiles_per_chunk);
        {
            cnt_s[tid] = 0;
            thr_s[tid] = (q0 + tid < nq) ? (init_thr ? init_thr[q0 + tid] : -FLT_MAX) : INFINITY;
            ubs_s[tid] = (ub_score && q0 + tid < nq) ? ub_score[q0 + tid] : INFINITY;
            ubr_s[tid] = (ub_row && q0 + tid < nq) ? ub_row[q0 + tid] : 0u;
        }
        __syncthreads();
        for (int t = t0; t < t1; t++) {
            const long long x0 = (long long)phys_tile(t) * BN;
            unsigned long long acc2[16][4];
#pragma unroll
            for (int i = 0; i < 16; i++)
#pragma unroll
                for (int p = 0; p < 4; p++) acc2[i][p] = 0ull;

            for (int kb = 0; kb < num_kb; kb++, it++) {
                if (pc.item < n_items) {
                    if (warp == (int)(it & (WARPS - 1)) && lane == 0) issue();
                    p_next();
                }
                const uint32_t s = c_stage;
                mbar_wait(&full_bar[s], c_phase);
                if (++c_stage == STAGES) {
                    c_stage = 0;
                    c_phase ^= 1;
                }
                const uint32_t As = ring + s * STAGE_BYTES + a_off, Bs = ring + s * STAGE_BYTES + b_off;
#pragma unroll 8
                for (int kk = 0; kk < BK; kk++) {
                    const float4 a0 = lds128f(As + (uint32_t)kk * 512u);
                    const float4 a1 = lds128f(As + (uint32_t)kk * 512u + 256u);
                    const float4 a2 = lds128f(As + 16384u + (uint32_t)kk * 512u);
                    const float4 a3 = lds128f(As + 16384u + (uint32_t)kk * 512u + 256u);
                    const ulonglong2 b0 = lds128(Bs + (uint32_t)kk * 512u);
                    const ulonglong2 b1 = lds128(Bs + (uint32_t)kk * 512u + 256u);
                    const float a[16] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x, a2.y, a2.z, a2.w, a3.x, a3.y, a3.z, a3.w};
                    const unsigned long long b[4] = {b0.x, b0.y, b1.x, b1.y};
#pragma unroll
                    for (int p = 0; p < 4; p++)
#pragma unroll
                        for (int i = 0; i < 16; i++) ffma2(acc2[i][p], a[i], b[p]);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[s]);
            }
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const int r = (i >> 2) * 64 + ty * 4 + (i & 3);
                const float thr = thr_s[r];
                const float ubs = ubs_s[r];
                const uint32_t ubr = ubr_s[r];
                unsigned long long* list = reinterpret_cast<unsigned long long*>(item_lists + (size_t)r * VQ_CAP);
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const long long xr = x0 + ((j < 4) ? tx * 4 + j : 64 + tx * 4 + (j - 4));
                    const float sc = __uint_as_float((j & 1) ? (uint32_t)(acc2[i][j >> 1] >> 32) : (uint32_t)acc2[i][j >> 1]);
                    if (xr < n_rows && sc > thr && (sc < ubs || (sc == ubs && (uint32_t)xr > ubr))) {
                        const int pos = atomicAdd(&cnt_s[r], 1);
                        __stcg(list + pos, (unsigned long long)pack_raw(sc, (uint32_t)xr));
                    }
                }
            }
            __syncthreads();
            for (int r = warp; r < BM; r += WARPS) {
                const int n = cnt_s[r];
                if (n > VQ_CAP - BN) {
                    float thr_new;
                    const int m = warp_prune_list(item_lists + (size_t)r * VQ_CAP, n, k, lane, thr_new);
                    if (lane == 0) {
                        cnt_s[r] = m;
                        thr_s[r] = fmaxf(thr_s[r], thr_new);
                    }
                }
            }
            __syncthreads();
        }
        counts[(size_t)chunk * nq_pad + q0 + tid] = (q0 + tid < nq) ? cnt_s[tid] : 0;
        __syncthreads();
    }
}

}  // namespace f32w

// plan with 256-query tiles and one CTA per SM
inline void f32_plan_wide(int nq, long long n_rows, int n_sms, F32Plan* plan) {
    const int max_ctas = n_sms;
    const int m_tiles = (nq + f32w::BM - 1) / f32w::BM;
    const int n_tiles = (int)((n_rows + f32w::BN - 1) / f32w::BN);
    int best_c = 1;
    double best_eff = 0.0;
    int c_max = n_tiles < 4 * max_ctas ? (n_tiles > 0 ? n_tiles : 1) : 4 * max_ctas;
    if (c_max > 1000) c_max = 1000;
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
    plan->nq_pad = m_tiles * f32w::BM;
    const long long items = (long long)m_tiles * plan->n_chunks;
    plan->n_ctas = (int)(items < max_ctas ? items : max_ctas);
    if (plan->n_ctas < 1) plan->n_ctas = 1;
}

}  // namespace vq
