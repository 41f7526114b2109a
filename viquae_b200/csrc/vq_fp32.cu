// K1 — fp32 "exact" mode: SIMT contraction with a FIXED reduction order + fused top-k.
//
// Stands in for faiss IndexFlat(d, METRIC_INNER_PRODUCT)::search on an fp32 corpus
// (meerqat/ir/search.py:143-146 casts queries to float32; datasets/search.py:384).  Every score is
//     acc = fmaf(q[j], x[j], acc),  j = 0 .. d-1 ascending, acc0 = +0
// i.e. bit-for-bit what oracle/flat_ip.c computes, so ids match except at exactly tied scores
// (and those follow the same lower-id-first rule).
//
// Round-2 form (measured against six other arrangements, profiles/r2_k1_experiments.txt):
//   * Both operands live K-MAJOR in HBM, in tiles of 128 rows: [tile][kpad][128] floats (kpad = d rounded up to 32; the
//     corpus is stored that way by vq_index_add, the queries are written that way by the query-prep kernel).  A k-block of a
//     tile is one dense 16 KB box, so TMA bulk-tensor copies drop it into shared memory exactly as the inner loop reads it:
//     no staging loads, no transposing stores, no per-k-block __syncthreads — a 3-stage ring (32 k per stage) guarded by
//     mbarriers, the refill issued by one elected lane of a warp whose turn rotates with the k block.
//   * Packed FFMA2 (fma.rn.f32x2, new on sm_100): acc.lo = fma(a, b.lo, acc.lo), acc.hi = fma(a, b.hi, acc.hi) in ONE
//     issue slot.  Each half rounds once, exactly like fmaf, so the sum stays bit-identical to the oracle.  The packed
//     operand is a PAIR of neighbouring corpus rows (one LDS.128 = two pairs), the scalar a query value (ptxas folds the
//     duplicated scalar into the instruction's .F32 broadcast operand and keeps the pair in the operand-reuse cache: 2.03
//     pipe cycles per FFMA2; the mirrored arrangement — pairs of queries, scalar corpus values — costs 2.25).
// 128 (queries) x 128 (corpus rows) tile per CTA, 256 threads, 8 x 8 outputs per thread, 2 CTAs per SM.
#include <cstdlib>
#include <cuda.h>
#include "vq_common.cuh"
#include "vq_kernels.h"

namespace vq {
namespace f32 {

constexpr int BM = 128, BN = 128, BK = 32, STAGES = 3, THREADS = 256, WARPS = THREADS / 32;
constexpr uint32_t A_BYTES = BK * BM * 4, B_BYTES = BK * BN * 4, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 128;  // + alignment slack (TMA destinations: 128-byte aligned)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const void* tmap, uint64_t* bar, uint32_t smem_dst, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_dst),
                 "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ ulonglong2 lds_pairs(uint32_t addr) {  // four floats as two packed pairs
    ulonglong2 v;
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
// acc.lo = fma(a, b.lo, acc.lo), acc.hi = fma(a, b.hi, acc.hi)
__device__ __forceinline__ void ffma2(unsigned long long& acc, float a, unsigned long long b) {
    unsigned long long aa;
    asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(aa), "l"(b));
}

// position of the load stream: (item, tile, k block) and the TMA coordinates that go with it; advanced in step by every
// thread (warp-uniform).  The per-k-block path is an increment and a compare; divisions only when a tile or item ends.
struct Cursor {
    int item, t, t1, kb;
    int qrow;  // row of the item's query tile in the k-major query matrix (m_tile * kpad)
    int xrow;  // row of corpus tile t in the k-major corpus matrix (physical tile * kpad)
};

__global__ void __launch_bounds__(THREADS, 2)
f32_topk_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_x, int nq, long long n_rows, int d, int k,
                int m_tiles, int tile_begin, int n_tiles, int map_stride, int map_ratio, int n_chunks, int tiles_per_chunk, int nq_pad,
                const float* __restrict__ init_thr, const float* __restrict__ ub_score, const uint32_t* __restrict__ ub_row,
                uint64_t* __restrict__ lists, int* __restrict__ counts) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES];
    __shared__ float thr_s[BM];
    __shared__ int cnt_s[BM];
    __shared__ float ubs_s[BM];      // exclusive upper bound of banded rounds (k > 128); +inf when unbanded
    __shared__ uint32_t ubr_s[BM];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // 16 x 16 threads, 8 x 8 outputs each as 2 x 2 blocks of 4 x 4: queries {4ty..4ty+3, 64+4ty..}, corpus rows {4tx.., 64+4tx..}
    // (the float4 reads of the corpus tile are 16 consecutive 16-byte chunks per half-warp, those of the query tile broadcasts)
    const int ty = tid >> 4, tx = tid & 15;
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

    auto phys_tile = [&](int t) {  // virtual tile of the pass -> physical corpus tile (every priming pass samples the whole corpus)
        int jt = t - tile_begin;
        if (map_ratio) jt = (jt / (map_ratio - 1)) * map_ratio + (jt % (map_ratio - 1)) + 1;
        return jt * map_stride;
    };
    // the producer's cursor runs STAGES-1 loads ahead of the consumers'
    Cursor pc = {};
    auto pc_set_item = [&]() {  // first non-empty item at or after pc.item (a chunk past the last tile is empty)
        for (; pc.item < n_items; pc.item += gridDim.x) {
            const int chunk = pc.item / m_tiles;
            pc.t = tile_begin + chunk * tiles_per_chunk;
            pc.t1 = min(n_tiles, pc.t + tiles_per_chunk);
            if (pc.t < pc.t1) {
                pc.qrow = (pc.item - chunk * m_tiles) * kpad;
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
    uint32_t p_stage = 0, p_phase = 1;  // the producer waits for the PREVIOUS phase of a stage's empty barrier
    auto issue = [&]() {  // one elected lane of one warp
        mbar_wait(&empty_bar[p_stage], p_phase);
        mbar_arrive_expect_tx(&full_bar[p_stage], STAGE_BYTES);
        const uint32_t dst = ring + p_stage * STAGE_BYTES;
        tma_load_2d(&tmap_q, &full_bar[p_stage], dst, 0, pc.qrow + pc.kb * BK);
        tma_load_2d(&tmap_x, &full_bar[p_stage], dst + A_BYTES, 0, pc.xrow + pc.kb * BK);
    };
    auto p_next = [&]() {
        if (++p_stage == STAGES) {
            p_stage = 0;
            p_phase ^= 1;
        }
        pc_advance();
    };
    // prologue: STAGES-1 loads in flight
    for (int i = 0; i < STAGES - 1; i++) {
        if (pc.item < n_items) {
            if (tid == 0) issue();
            p_next();
        }
    }

    uint32_t it = 0;  // consumed k blocks so far (producer duty = it mod 8)
    uint32_t c_stage = 0, c_phase = 0;
    const uint32_t a_off = (uint32_t)ty * 16u, b_off = A_BYTES + (uint32_t)tx * 16u;

    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int m_tile = item % m_tiles, chunk = item / m_tiles;
        const int q0 = m_tile * BM;
        uint64_t* item_lists = lists + ((size_t)chunk * nq_pad + q0) * VQ_CAP;
        const int t0 = tile_begin + chunk * tiles_per_chunk;
        const int t1 = min(n_tiles, t0 + tiles_per_chunk);
        if (tid < BM) {
            cnt_s[tid] = 0;
            // primed threshold (next float below the k-th score of the strided samples the earlier passes searched: those
            // rows may have HIGHER ids, so an equal score must still pass the strict compare) or -FLT_MAX
            thr_s[tid] = (q0 + tid < nq) ? (init_thr ? init_thr[q0 + tid] : -FLT_MAX) : INFINITY;
            ubs_s[tid] = (ub_score && q0 + tid < nq) ? ub_score[q0 + tid] : INFINITY;
            ubr_s[tid] = (ub_row && q0 + tid < nq) ? ub_row[q0 + tid] : 0u;
        }
        __syncthreads();
        for (int t = t0; t < t1; t++) {
            const long long x0 = (long long)phys_tile(t) * BN;
            // accumulators as packed pairs of neighbouring corpus rows: acc2[i][p] = (out[i][2p], out[i][2p+1])
            unsigned long long acc2[8][4];
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int p = 0; p < 4; p++) acc2[i][p] = 0ull;

            for (int kb = 0; kb < num_kb; kb++, it++) {
                // producer duty rotates over the warps: refill the stage every warp released one k block ago
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
#pragma unroll
                for (int kk = 0; kk < BK; kk++) {
                    const float4 a0 = lds_f4(As + (uint32_t)kk * 512u);
                    const float4 a1 = lds_f4(As + (uint32_t)kk * 512u + 256u);
                    const ulonglong2 b0 = lds_pairs(Bs + (uint32_t)kk * 512u);
                    const ulonglong2 b1 = lds_pairs(Bs + (uint32_t)kk * 512u + 256u);
                    const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                    const unsigned long long b[4] = {b0.x, b0.y, b1.x, b1.y};
#pragma unroll
                    for (int i = 0; i < 8; i++)
#pragma unroll
                        for (int p = 0; p < 4; p++) ffma2(acc2[i][p], a[i], b[p]);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[s]);
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
                    const float s = __uint_as_float((j & 1) ? (uint32_t)(acc2[i][j >> 1] >> 32) : (uint32_t)acc2[i][j >> 1]);
                    if (xr < n_rows && s > thr && (s < ubs || (s == ubs && (uint32_t)xr > ubr))) {
                        const int pos = atomicAdd(&cnt_s[r], 1);
                        __stcg(list + pos, (unsigned long long)pack_raw(s, (uint32_t)xr));
                    }
                }
            }
            __syncthreads();
            // a tile adds at most BN = 128 keys to a list: prune whenever fewer than 128 slots remain
            for (int r = warp; r < BM; r += WARPS) {
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
            __syncthreads();
        }
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

// ----------------------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled f32_get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess || !ptr) return nullptr;
        fn = reinterpret_cast<PFN_encodeTiled>(ptr);
    }
    return fn;
}
// k-major tiled fp32 matrix [tiles][kpad][128] seen as [tiles * kpad][128]; box = one k block of one tile (32 x 128, dense)
static int f32_encode_kmajor(CUtensorMap* m, const void* base, uint64_t tiles, int kpad) {
    PFN_encodeTiled enc = f32_get_encode();
    if (!enc) return -1;
    cuuint64_t gdim[2] = {(cuuint64_t)f32::BN, tiles * (uint64_t)kpad};
    cuuint64_t gstride[1] = {(cuuint64_t)f32::BN * 4};
    cuuint32_t box[2] = {(cuuint32_t)f32::BN, (cuuint32_t)f32::BK};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -(int)r - 1000;
}

int f32_kpad(int d) { return (d + f32::BK - 1) / f32::BK * f32::BK; }

int f32_topk_launch(const float* qt, int nq, const float* xt, long long n_rows, int d, int k,
                    const F32Plan& plan, int tile_begin, int tile_end, int map_stride, int map_ratio, const float* init_thr,
                    const float* ub_score, const uint32_t* ub_row, uint64_t* lists, int* counts, cudaStream_t stream) {
    if (k < 1 || k > VQ_KMAX) return -2;
    if (nq <= 0) return 0;
    if (n_rows <= 0 || ((uintptr_t)xt & 15) || ((uintptr_t)qt & 15)) return -2;
    const int kpad = f32_kpad(d);
    CUtensorMap tq, tx;
    int rc = f32_encode_kmajor(&tq, qt, (uint64_t)plan.m_tiles, kpad);
    if (rc) return rc;
    rc = f32_encode_kmajor(&tx, xt, (uint64_t)((n_rows + f32::BN - 1) / f32::BN), kpad);
    if (rc) return rc;
    static bool attr_set_dev[64] = {};  // cudaFuncSetAttribute is per device
    int dev_ = 0;
    cudaGetDevice(&dev_);
    if (!attr_set_dev[dev_ & 63]) {
        cudaError_t e = cudaFuncSetAttribute(f32::f32_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f32::SMEM_BYTES);
        if (e != cudaSuccess) return -(int)e;
        attr_set_dev[dev_ & 63] = true;
    }
    f32::f32_topk_kernel<<<plan.n_ctas, f32::THREADS, f32::SMEM_BYTES, stream>>>(tq, tx, nq, n_rows, d, k, plan.m_tiles, tile_begin, tile_end, map_stride, map_ratio,
                                                                               plan.n_chunks, plan.tiles_per_chunk, plan.nq_pad, init_thr, ub_score, ub_row, lists,
                                                                               counts);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -(int)e;
}

}  // namespace vq
