// EXPERIMENT FILE (round 2) — not part of the product build.  A copy of csrc/vq_fp32.cu grown into template variants of K1
// (the fp32 exact contraction) for the A/B harness k1_variants.cu: packed FFMA2 vs scalar FFMA, lane arrangements,
// XOR-swizzled vs padded tiles, lean staging loads, no staging at all (upper bound), and a TMA-fed build with mbarriers.
// Every variant with real operand traffic lands at 0.63-0.67 of the FFMA peak (profiles/r2_k1_experiments.txt), so the
// product keeps the round-1 kernel.  Self-contained: declares its own host entry points.
// K1 — fp32 "exact" mode: SIMT FFMA contraction with a FIXED reduction order + fused top-k.
//
// Stands in for faiss IndexFlat(d, METRIC_INNER_PRODUCT)::search on an fp32 corpus
// (meerqat/ir/search.py:143-146 casts queries to float32; datasets/search.py:384).  Every score is
//     acc = fmaf(q[j], x[j], acc),  j = 0 .. d-1 ascending, acc0 = +0
// i.e. bit-for-bit what oracle/flat_ip.c computes, so ids match except at exactly tied scores
// (and those follow the same lower-id-first rule).
//
// 128 (queries) x 128 (corpus rows) tile per CTA, 256 threads, 8x8 outputs per thread, K staged
// through double-buffered shared memory 16 columns at a time (one barrier per K block).  Bound by the FMA pipe
// (packed FFMA2: two outputs per issue slot) with the shared-memory pipe close behind.
#include <cstdlib>
#include <cuda.h>
#include "vq_common.cuh"
#include "vq_kernels.h"
namespace vq {
size_t f32_kmajor_bytes(int nq_pad, int d);
int f32_queries_kmajor_launch(const float* q, int nq_pad, int ld_q, int d, float* qt, cudaStream_t stream);
int f32_topk_tma_launch(const float* qt, int nq, const float* x, long long n_rows, int ld_x, int d, int k,
                        const F32Plan& plan, int tile_begin, int tile_end, int map_stride, int map_ratio, const float* init_thr,
                        const float* ub_score, const uint32_t* ub_row, uint64_t* lists, int* counts, cudaStream_t stream);
}


namespace vq {
namespace f32 {

constexpr int BM = 128, BN = 128, BK = 16, THREADS = 256;

// Two IEEE fused multiply-adds in one issue slot (Blackwell FFMA2): acc.lo = fma(a, b.lo, acc.lo), acc.hi = fma(a, b.hi, acc.hi).
// Each half rounds once, exactly like fmaf, so the fixed-order sum stays bit-identical to oracle/flat_ip.c; ptxas folds the
// duplicated scalar into the instruction's .F32 broadcast operand (no extra moves).
__device__ __forceinline__ void ffma2(unsigned long long& acc, float a, unsigned long long b) {
    unsigned long long aa;
    asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(aa), "l"(b));
}
__device__ __forceinline__ void ffma1(unsigned long long& acc, float a, unsigned long long b) {
    const float lo = fmaf(a, __uint_as_float((uint32_t)b), __uint_as_float((uint32_t)acc));
    const float hi = fmaf(a, __uint_as_float((uint32_t)(b >> 32)), __uint_as_float((uint32_t)(acc >> 32)));
    acc = ((unsigned long long)__float_as_uint(hi) << 32) | __float_as_uint(lo);
}

// k-major shared-memory tiles [BK][128] without padding.  Column r of k-row kk lives at r ^ swz<SWZ>(kk): the XOR moves whole
// 8-float groups, so the float4 reads of the FFMA loop stay aligned and contiguous, and the transposing stores of the
// staging step (a warp writes 8 rows x 4 k-chunks at once) land in 32 different banks.
template <bool SWZ> __device__ __forceinline__ constexpr int swz(int kk) { return SWZ ? ((kk >> 2) & 3) << 3 : 0; }

// PACKED: FFMA2 (two outputs per instruction) or scalar FFMA.  ALIGNED: d % 16 == 0 and 16-byte row strides — the staging
// loads are unpredicated float4 loads from clamped row pointers; otherwise the generic, fully predicated loads.
template <bool PACKED, bool ALIGNED, bool LANES48 = true, bool SWZ = true, bool STAGE = true, bool DUMMY_STS = false>
__global__ void __launch_bounds__(THREADS, 2)
f32_topk_kernel(const float* __restrict__ Q, int nq, int ld_q, const float* __restrict__ X, long long n_rows, int ld_x, int d, int k,
                int m_tiles, int tile_begin, int n_tiles, int map_stride, int map_ratio, int n_chunks, int tiles_per_chunk, int nq_pad,
                const float* __restrict__ init_thr,
                const float* __restrict__ ub_score, const uint32_t* __restrict__ ub_row, uint64_t* __restrict__ lists,
                int* __restrict__ counts) {
    constexpr int PITCH = SWZ ? BM : BM + 4;
    __shared__ __align__(16) float As[2][BK][PITCH];  // double-buffered k-major tiles (columns XOR-swizzled, see swz)
    __shared__ __align__(16) float Bs[2][BK][PITCH];
    __shared__ float thr_s[BM];
    __shared__ int cnt_s[BM];
    __shared__ float ubs_s[BM];      // exclusive upper bound of banded rounds (k > 128); +inf when unbanded
    __shared__ uint32_t ubr_s[BM];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // 16 x 16 threads, 8 x 8 outputs each as 2 x 2 blocks of 4 x 4: rows {4ty..4ty+3, 64+4ty..}, cols {4tx.., 64+4tx..}.
    // A warp covers 4 (ty) x 8 (tx) threads: per half-warp a float4 read of the A tile touches 2 addresses, one of the B
    // tile 8 consecutive 16-byte chunks (128 bytes) — 2 shared-memory wavefronts per LDS.128 (a 2 x 16 arrangement: 2 and 4).
    const int ty = LANES48 ? (warp >> 1) * 4 + (lane >> 3) : tid >> 4, tx = LANES48 ? (warp & 1) * 8 + (lane & 7) : tid & 15;
    // global -> smem staging: thread loads float4 (4 consecutive k) of row lr for A and B, twice (rows lr, lr+64)
    const int lr = tid >> 2, lk = (tid & 3) * 4;
    const int lsw = SWZ ? (tid & 3) << 3 : 0;  // swz(lk + c), c = 0..3
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
        // ALIGNED: rows past the end are clamped to the last one (their outputs are discarded below: thr = +inf / xr >= n_rows)
        const float* pa[2];
#pragma unroll
        for (int h = 0; h < 2; h++) pa[h] = Q + (size_t)min(q0 + lr + h * 64, nq - 1) * ld_q + lk;
        const int t0 = tile_begin + chunk * tiles_per_chunk;
        const int t1 = min(n_tiles, t0 + tiles_per_chunk);
        for (int t = t0; t < t1; t++) {
            // virtual tile of the pass -> physical corpus tile (every priming pass samples the whole corpus)
            int jt = t - tile_begin;
            if (map_ratio) jt = (jt / (map_ratio - 1)) * map_ratio + (jt % (map_ratio - 1)) + 1;
            const long long x0 = (long long)jt * map_stride * BN;
            const float* pb[2];
#pragma unroll
            for (int h = 0; h < 2; h++) pb[h] = X + (size_t)min(x0 + lr + h * 64, n_rows - 1) * ld_x + lk;
            // accumulators as packed pairs of neighbouring corpus rows: acc2[i][p] = (out[i][2p], out[i][2p+1])
            unsigned long long acc2[8][4];
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int p = 0; p < 4; p++) acc2[i][p] = 0ull;

            float4 ra[2], rb[2];
            auto gload = [&](int kb) {
                if (ALIGNED) {
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        ra[h] = *reinterpret_cast<const float4*>(pa[h] + kb * BK);
                        rb[h] = __ldg(reinterpret_cast<const float4*>(pb[h] + kb * BK));
                    }
                    return;
                }
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
                    const int r = (lr + h * 64) ^ lsw;
                    As[buf][lk + 0][r] = ra[h].x; As[buf][lk + 1][r] = ra[h].y; As[buf][lk + 2][r] = ra[h].z; As[buf][lk + 3][r] = ra[h].w;
                    Bs[buf][lk + 0][r] = rb[h].x; Bs[buf][lk + 1][r] = rb[h].y; Bs[buf][lk + 2][r] = rb[h].z; Bs[buf][lk + 3][r] = rb[h].w;
                }
            };
            gload(0);
            sstore(0);
            if (!STAGE) {  // measurement build: both buffers hold real data, nothing is staged inside the loop
                gload(1);
                sstore(1);
            }
            __syncthreads();
            for (int kb = 0; kb < num_kb; kb++) {
                const int buf = kb & 1;
                if (STAGE && kb + 1 < num_kb) gload(kb + 1);  // global loads fly over the FFMA block below
#pragma unroll
                for (int kk = 0; kk < BK; kk++) {
                    float a[8];
                    const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][(ty * 4) ^ swz<SWZ>(kk)]);
                    const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ((ty * 4) ^ swz<SWZ>(kk))]);
                    const ulonglong2 b0 = *reinterpret_cast<const ulonglong2*>(&Bs[buf][kk][(tx * 4) ^ swz<SWZ>(kk)]);
                    const ulonglong2 b1 = *reinterpret_cast<const ulonglong2*>(&Bs[buf][kk][64 + ((tx * 4) ^ swz<SWZ>(kk))]);
                    a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
                    const unsigned long long b[4] = {b0.x, b0.y, b1.x, b1.y};
#pragma unroll
                    for (int i = 0; i < 8; i++)
#pragma unroll
                        for (int p = 0; p < 4; p++) {
                            if (PACKED) ffma2(acc2[i][p], a[i], b[p]);
                            else ffma1(acc2[i][p], a[i], b[p]);
                        }
                }
                // the other buffer was last read one iteration ago, before the barrier below: safe to overwrite
                if ((STAGE || DUMMY_STS) && kb + 1 < num_kb) sstore(buf ^ 1);
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
                    const float s = __uint_as_float((j & 1) ? (uint32_t)(acc2[i][j >> 1] >> 32) : (uint32_t)acc2[i][j >> 1]);
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


// ============================================================================ TMA-fed kernel
// The same contraction with NO staging instructions: both operand tiles arrive by TMA bulk-tensor copies into a
// 3-stage shared-memory ring (32 k-columns per stage), guarded by mbarriers instead of per-k-block __syncthreads — the
// warps of a CTA drift by up to a stage instead of meeting every 16 columns.
//   * corpus tile: 128 rows x 32 columns straight from the row-major fp32 matrix, 128-byte swizzle (16-byte chunk c of
//     row r sits at chunk c ^ (r & 7)): a thread reads float4 = four k of ONE row, eight lanes read eight neighbouring
//     rows -> 32 distinct banks.  A corpus value is the SCALAR operand of FFMA2.
//   * query tile: 32 k-rows x 128 queries from a k-major copy of the prepared queries ([m_tile][k][128], written once per
//     search by queries_kmajor_kernel): a float4 = two PAIRS of neighbouring queries at one k — the packed operand.
//   acc2[p][j] = (out[query 2p][row j], out[query 2p+1][row j]); each half still sees k ascending, one rounding per step.
// Columns past d are zero in both operands (TMA out-of-bounds fill / zero rows of the k-major copy): fma(0, 0, acc) = acc.
namespace tma {

constexpr int BM = 128, BN = 128, BK = 32, STAGES = 3, THREADS = 256, WARPS = THREADS / 32;
constexpr uint32_t A_BYTES = BK * BM * 4, B_BYTES = BN * BK * 4, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024;  // + alignment slack (the swizzle pattern is address based)

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
__device__ __forceinline__ ulonglong2 lds128(uint32_t addr) {
    ulonglong2 v;
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
// acc.lo = fma(a.lo, b, acc.lo), acc.hi = fma(a.hi, b, acc.hi): the scalar is the broadcast operand
__device__ __forceinline__ void ffma2s(unsigned long long& acc, unsigned long long a, float b) {
    unsigned long long bb;
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(bb));
}
__device__ __forceinline__ ulonglong2 lds64x2(uint32_t addr) {  // two floats as two registers
    ulonglong2 v;
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ float2 lds64f(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}

// position of the load stream: (item, tile, k block) and the TMA coordinates that go with it; advanced in step by every
// thread (warp-uniform).  The per-k-block path is an increment and a compare; divisions only when a tile or item ends.
struct Cursor {
    int item, t, t1, kb;
    int qrow;  // first row of the item's query tile in the k-major query matrix (m_tile * kpad)
    int xrow;  // first corpus row of tile t
};

template <int VAR, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
f32_topk_tma_kernel(const __grid_constant__ CUtensorMap tmap_qt, const __grid_constant__ CUtensorMap tmap_x, int hot, int nq, long long n_rows, int d, int k,
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
    const int ty = tid >> 4, tx = tid & 15;  // queries {4ty..4ty+3, 64+4ty..}, corpus rows {tx + 16 j}
    const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
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

    auto phys_row = [&](int t) {  // virtual tile of the pass -> first row of the physical corpus tile
        int jt = t - tile_begin;
        if (map_ratio) jt = (jt / (map_ratio - 1)) * map_ratio + (jt % (map_ratio - 1)) + 1;
        return jt * map_stride * BN;
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
                pc.xrow = phys_row(pc.t);
                return;
            }
        }
    };
    pc.item = blockIdx.x;
    pc.kb = 0;
    pc.t = pc.t1 = pc.qrow = pc.xrow = 0;
    pc_set_item();
    auto pc_advance = [&]() {
        if (++pc.kb < num_kb) return;
        pc.kb = 0;
        if (++pc.t < pc.t1) {
            pc.xrow = phys_row(pc.t);
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
        tma_load_2d(&tmap_qt, &full_bar[p_stage], dst, 0, hot ? 0 : pc.qrow + pc.kb * BK);
        if (VAR >= 5) tma_load_2d(&tmap_x, &full_bar[p_stage], dst + A_BYTES, 0, hot ? 0 : (pc.xrow / BN) * kpad + pc.kb * BK);
        else tma_load_2d(&tmap_x, &full_bar[p_stage], dst + A_BYTES, hot ? 0 : pc.kb * BK, hot ? 0 : pc.xrow);
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
    // swizzled 16-byte chunk of this thread's corpus rows: row & 7 == tx & 7 for every row tx + 16 j
    const uint32_t b_row = (uint32_t)tx * 128u, b_x = (uint32_t)(tx & 7) << 4;
    const uint32_t a_off = (uint32_t)ty * 16u;

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
            const long long x0 = phys_row(t);
            unsigned long long acc2[4][8];
            float facc[4][8][2];
#pragma unroll
            for (int p = 0; p < 4; p++)
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    acc2[p][j] = 0ull;
                    facc[p][j][0] = facc[p][j][1] = 0.0f;
                }

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
                const uint32_t As = ring + s * STAGE_BYTES + a_off, Bs = ring + s * STAGE_BYTES + A_BYTES + b_row;
                if (VAR >= 5) {
                    // both operands k-major: queries are the scalars, PAIRS of neighbouring corpus rows the packed operand
                    // (the inner loop of the register-staged kernel, fed by TMA)
                    const uint32_t Ak = ring + s * STAGE_BYTES + (uint32_t)ty * 16u, Bk = ring + s * STAGE_BYTES + A_BYTES + (uint32_t)tx * 16u;
#pragma unroll
                    for (int kk = 0; kk < BK; kk++) {
                        const float4 a0 = lds128f(Ak + (uint32_t)kk * 512u);
                        const float4 a1 = lds128f(Ak + (uint32_t)kk * 512u + 256u);
                        const ulonglong2 b0 = lds128(Bk + (uint32_t)kk * 512u);
                        const ulonglong2 b1 = lds128(Bk + (uint32_t)kk * 512u + 256u);
                        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                        const unsigned long long b[4] = {b0.x, b0.y, b1.x, b1.y};
                        // acc2[p][j] doubles as [pair of rows p][query j]: 4 x 8
                        // VAR 7 / 8 / 9: only the first 1 / 2 / 0 pairs of rows go through FFMA2, the rest through scalar FFMA
                        // (mixing the issue-cheap packed form with the pipe-cheap scalar form)
                        constexpr int NP2 = VAR == 7 ? 1 : VAR == 8 ? 2 : VAR == 9 ? 0 : 4;
#pragma unroll
                        for (int i = 0; i < 8; i++)
#pragma unroll
                            for (int p = 0; p < 4; p++) {
                                if (p < NP2) ffma2(acc2[p][i], a[i], b[p]);
                                else {
                                    facc[p][i][0] = fmaf(a[i], __uint_as_float((uint32_t)b[p]), facc[p][i][0]);
                                    facc[p][i][1] = fmaf(a[i], __uint_as_float((uint32_t)(b[p] >> 32)), facc[p][i][1]);
                                }
                            }
                    }
                } else if (VAR == 2) {
                    // corpus operand two k at a time (LDS.64): 16 instead of 32 operand registers
#pragma unroll
                    for (int c = 0; c < BK / 2; c++) {
                        float2 b[8];
#pragma unroll
                        for (int j = 0; j < 8; j++) b[j] = lds64f(Bs + (((uint32_t)(c >> 1) << 4) ^ b_x) + (uint32_t)(c & 1) * 8u + (uint32_t)j * 2048u);
#pragma unroll
                        for (int kk = 0; kk < 2; kk++) {
                            const ulonglong2 a0 = lds128(As + (uint32_t)(c * 2 + kk) * 512u);
                            const ulonglong2 a1 = lds128(As + (uint32_t)(c * 2 + kk) * 512u + 256u);
                            const unsigned long long a[4] = {a0.x, a0.y, a1.x, a1.y};
#pragma unroll
                            for (int j = 0; j < 8; j++) {
                                const float bv = kk == 0 ? b[j].x : b[j].y;
#pragma unroll
                                for (int p = 0; p < 4; p++) ffma2s(acc2[p][j], a[p], bv);
                            }
                        }
                    }
                } else {
#pragma unroll
                for (int c = 0; c < BK / 4; c++) {
                    float4 b[8];  // four k of each of this thread's eight corpus rows
#pragma unroll
                    for (int j = 0; j < 8; j++) b[j] = lds128f(Bs + (((uint32_t)c << 4) ^ b_x) + (uint32_t)j * 2048u);
#pragma unroll
                    for (int kk = 0; kk < 4; kk++) {
                        const ulonglong2 a0 = lds128(As + (uint32_t)(c * 4 + kk) * 512u);
                        const ulonglong2 a1 = lds128(As + (uint32_t)(c * 4 + kk) * 512u + 256u);
                        const unsigned long long a[4] = {a0.x, a0.y, a1.x, a1.y};
                        if (VAR == 0) {
#pragma unroll
                            for (int j = 0; j < 8; j++) {
                                const float bv = kk == 0 ? b[j].x : kk == 1 ? b[j].y : kk == 2 ? b[j].z : b[j].w;
#pragma unroll
                                for (int p = 0; p < 4; p++) ffma2s(acc2[p][j], a[p], bv);
                            }
                        } else if (VAR == 3) {
#pragma unroll
                            for (int j = 0; j < 8; j++) {
                                const float bv = kk == 0 ? b[j].x : kk == 1 ? b[j].y : kk == 2 ? b[j].z : b[j].w;
#pragma unroll
                                for (int p = 0; p < 4; p++) {
                                    const float lo = fmaf(__uint_as_float((uint32_t)a[p]), bv, __uint_as_float((uint32_t)acc2[p][j]));
                                    const float hi = fmaf(__uint_as_float((uint32_t)(a[p] >> 32)), bv, __uint_as_float((uint32_t)(acc2[p][j] >> 32)));
                                    acc2[p][j] = ((unsigned long long)__float_as_uint(hi) << 32) | __float_as_uint(lo);
                                }
                            }
                        } else {
#pragma unroll
                            for (int p = 0; p < 4; p++)
#pragma unroll
                                for (int j = 0; j < 8; j++) {
                                    const float bv = kk == 0 ? b[j].x : kk == 1 ? b[j].y : kk == 2 ? b[j].z : b[j].w;
                                    ffma2s(acc2[p][j], a[p], bv);
                                }
                        }
                    }
                }
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
                    constexpr int NP2E = VAR == 7 ? 1 : VAR == 8 ? 2 : VAR == 9 ? 0 : 4;
                    const long long xr = VAR >= 5 ? x0 + ((j < 4) ? tx * 4 + j : 64 + tx * 4 + (j - 4)) : x0 + tx + 16 * j;
                    const float s = VAR >= 5 ? ((j >> 1) < NP2E ? __uint_as_float((j & 1) ? (uint32_t)(acc2[j >> 1][i] >> 32) : (uint32_t)acc2[j >> 1][i])
                                                                : facc[j >> 1][i][j & 1])
                                             : __uint_as_float((i & 1) ? (uint32_t)(acc2[i >> 1][j] >> 32) : (uint32_t)acc2[i >> 1][j]);
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

// prepared queries [nq_pad][ld_q] (rows past nq already zero) -> k-major tiles qt[m_tile][kpad][128]; columns >= d zero
__global__ void __launch_bounds__(256) queries_kmajor_kernel(const float* __restrict__ q, int ld_q, int d, int kpad, float* __restrict__ qt) {
    __shared__ float tile[32][33];
    const int m_tile = blockIdx.y, k0 = blockIdx.x * 32;
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;  // 32 x 8
    for (int qb = 0; qb < BM; qb += 32) {
#pragma unroll
        for (int r = ly; r < 32; r += 8) {
            const int kcol = k0 + lx;
            tile[r][lx] = kcol < d ? q[(size_t)(m_tile * BM + qb + r) * ld_q + kcol] : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int r = ly; r < 32; r += 8) qt[((size_t)m_tile * kpad + k0 + r) * BM + qb + lx] = tile[lx][r];
        __syncthreads();
    }
}

}  // namespace tma

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

static int f32_topk_staged_launch_unused(const float* q, int nq, int ld_q, const float* x, long long n_rows, int ld_x, int d, int k,
                           const F32Plan& plan, int tile_begin, int tile_end, int map_stride, int map_ratio, const float* init_thr,
                           const float* ub_score, const uint32_t* ub_row, uint64_t* lists, int* counts, cudaStream_t stream) {
    if (k < 1 || k > VQ_KMAX) return -2;
    if (nq <= 0) return 0;
    static int packed = -1;  // experiment knob: VQ_F32_FFMA2=0 selects the scalar-FFMA build of the same kernel
    if (packed < 0) {
        const char* e = getenv("VQ_F32_FFMA2");
        packed = e ? atoi(e) != 0 : 1;
    }
    const bool aligned = (d % f32::BK == 0) && (ld_q % 4 == 0) && (ld_x % 4 == 0) && (((uintptr_t)q | (uintptr_t)x) % 16 == 0) && n_rows > 0;
    auto kern = aligned ? (packed ? f32::f32_topk_kernel<true, true> : f32::f32_topk_kernel<false, true>)
                        : (packed ? f32::f32_topk_kernel<true, false> : f32::f32_topk_kernel<false, false>);
    kern<<<plan.n_ctas, f32::THREADS, 0, stream>>>(q, nq, ld_q, x, n_rows, ld_x, d, k, plan.m_tiles, tile_begin, tile_end, map_stride, map_ratio,
                                                  plan.n_chunks, plan.tiles_per_chunk, plan.nq_pad, init_thr, ub_score, ub_row, lists, counts);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -(int)e;
}


// ----------------------------------------------------------------------------- TMA-fed kernel, host side
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
// fp32 matrix [rows][cols], row stride ld floats; box = box_cols x box_rows; out-of-bounds elements read as zero
static int f32_encode_2d(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_cols, uint32_t box_rows,
                         CUtensorMapSwizzle swizzle) {
    PFN_encodeTiled enc = f32_get_encode();
    if (!enc) return -1;
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {ld * 4};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -(int)r - 1000;
}

size_t f32_kmajor_bytes(int nq_pad, int d) {
    const int kpad = (d + f32::tma::BK - 1) / f32::tma::BK * f32::tma::BK;
    return (size_t)nq_pad * kpad * sizeof(float);
}

int f32_queries_kmajor_launch(const float* q, int nq_pad, int ld_q, int d, float* qt, cudaStream_t stream) {
    if (nq_pad <= 0) return 0;
    const int kpad = (d + f32::tma::BK - 1) / f32::tma::BK * f32::tma::BK;
    f32::tma::queries_kmajor_kernel<<<dim3(kpad / 32, nq_pad / f32::tma::BM), 256, 0, stream>>>(q, ld_q, d, kpad, qt);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -(int)e;
}

int f32_topk_tma_launch(const float* qt, int nq, const float* x, long long n_rows, int ld_x, int d, int k,
                    const F32Plan& plan, int tile_begin, int tile_end, int map_stride, int map_ratio, const float* init_thr,
                    const float* ub_score, const uint32_t* ub_row, uint64_t* lists, int* counts, cudaStream_t stream) {
    namespace T = f32::tma;
    if (k < 1 || k > VQ_KMAX) return -2;
    if (nq <= 0) return 0;
    if (n_rows <= 0 || (ld_x & 3) || ((uintptr_t)x & 15) || ((uintptr_t)qt & 15)) return -2;
    const int kpad = (d + T::BK - 1) / T::BK * T::BK;
    CUtensorMap tq, tx;
    int rc = f32_encode_2d(&tq, qt, (uint64_t)plan.m_tiles * kpad, T::BM, T::BM, T::BM, T::BK, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
    const char* ev = getenv("VQ_F32_TMA_VARIANT");
    if (ev && atoi(ev) >= 5)  // x is the k-major tiled copy [n_tiles_total][kpad][128] (corpus_kmajor_kernel)
        rc = f32_encode_2d(&tx, x, (uint64_t)((n_rows + T::BN - 1) / T::BN) * kpad, T::BN, T::BN, T::BN, T::BK, CU_TENSOR_MAP_SWIZZLE_NONE);
    else
        rc = f32_encode_2d(&tx, x, (uint64_t)n_rows, (uint64_t)d, (uint64_t)ld_x, T::BK, T::BN, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    static int var = -1, hot = 0;  // experiment knobs: VQ_F32_TMA_VARIANT picks the instruction arrangement of the inner loop,
                                   // VQ_F32_TMA_HOT=1 makes every load fetch the same tile (wrong results: measurement only)
    if (var < 0) {
        const char* e = getenv("VQ_F32_TMA_VARIANT");
        var = e ? atoi(e) : 0;
        if (var < 0 || var > 9) var = 0;
        hot = getenv("VQ_F32_TMA_HOT") ? 1 : 0;
    }
    typedef void (*kern_t)(const CUtensorMap, const CUtensorMap, int, int, long long, int, int, int, int, int, int, int, int, int, int, const float*, const float*,
                           const uint32_t*, uint64_t*, int*);
    kern_t kern = var == 0 ? T::f32_topk_tma_kernel<0, 2> : var == 1 ? T::f32_topk_tma_kernel<1, 2> : var == 2 ? T::f32_topk_tma_kernel<2, 2>
                : var == 3 ? T::f32_topk_tma_kernel<3, 2> : var == 4 ? T::f32_topk_tma_kernel<0, 1> : var == 5 ? T::f32_topk_tma_kernel<5, 2> : var == 6 ? T::f32_topk_tma_kernel<5, 1>
                : var == 7 ? T::f32_topk_tma_kernel<7, 2> : var == 8 ? T::f32_topk_tma_kernel<8, 2> : T::f32_topk_tma_kernel<9, 2>;
    static bool attr_set_dev[64][10] = {};  // cudaFuncSetAttribute is per device
    int dev_ = 0;
    cudaGetDevice(&dev_);
    if (!attr_set_dev[dev_ & 63][var]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::SMEM_BYTES);
        if (e != cudaSuccess) return -(int)e;
        attr_set_dev[dev_ & 63][var] = true;
    }
    kern<<<plan.n_ctas, T::THREADS, T::SMEM_BYTES, stream>>>(tq, tx, hot, nq, n_rows, d, k, plan.m_tiles, tile_begin, tile_end, map_stride, map_ratio,
                                                           plan.n_chunks, plan.tiles_per_chunk, plan.nq_pad, init_thr, ub_score, ub_row, lists, counts);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -(int)e;
}

}  // namespace vq
