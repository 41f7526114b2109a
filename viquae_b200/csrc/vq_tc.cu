// K3 — bf16/fp16 query x corpus contraction on the 5th-gen tensor cores with a fused top-k.
//
// Replaces faiss IndexFlat(d, METRIC_INNER_PRODUCT)::search as reached from
// meerqat/ir/search.py:146 -> datasets/search.py:384 for 16-bit corpora.  The score matrix never
// reaches HBM: every 128(query) x 256(corpus row) tile is accumulated in TMEM by tcgen05.mma
// (operands staged in shared memory by TMA, 128-byte swizzle), read back with tcgen05.ld by eight
// epilogue warps, and filtered against a per-query running threshold held in a register.
//
// Layout of one tile:  D[M = queries, N = corpus rows] = Q[M, K] . X[N, K]^T, both operands K-major.
// Queries sit on the TMEM *lanes*, so one epilogue thread owns one query: its threshold and its
// candidate count live in registers and the 256 scores of a tile stream past it column by column.
//
// CG = 1: one CTA per SM, UMMA 128x256x16.   CG = 2: a CTA pair (cluster of 2) issues
// tcgen05.mma.cta_group::2 with UMMA 256x256x16: each CTA holds 128 queries and HALF of the corpus
// tile, halving the L2->SM corpus traffic per query row.
//
// Warp roles (384 threads): warp 0 TMA producer, warp 1 MMA issuer (leader CTA only), warp 2 TMEM
// allocator, warp 3 idle, warps 4..11 epilogue (TMEM lane quarter = warp % 4, column half =
// (warp - 4) / 4).  The epilogue is the critical path next to the MMA: its inner loop is branch-free
// (predicated 8-byte stores of raw (score, row) pairs through inline PTX).
//
// Work decomposition: item = (query tile, residue class g): the item's tiles are g, g + G, g + 2G, ... of
// the pass (see class_tiles / phys_tile), so that at any moment all clusters work inside a window of ~G
// tiles; the clusters of different query tiles that share a class are kept within 8 tiles of each other
// (L2 sharing window in the producer warp) so each corpus tile is pulled from HBM about once.
// One launch = one pass of the threshold-priming schedule (vq_api.cu: pass_schedule).
#include "vq_common.cuh"
#include "vq_kernels.h"

#include <cuda.h>
#include <stdlib.h>

namespace vq {
namespace tc {

constexpr int BLOCK_Q = 128;   // queries per CTA (TMEM lanes)
constexpr int BLOCK_N = 256;   // corpus rows per tile (UMMA_N, TMEM columns)
constexpr int BLOCK_K = 64;    // 64 x 2 B = one 128-byte swizzle atom
constexpr int UMMA_K = 16;
constexpr int ACC_STAGES = 2;  // 2 x 256 columns = all 512 TMEM columns
constexpr int EPI_WARP0 = 4;
constexpr int EPI_WARPS = 8;   // two warps per TMEM lane quarter: each takes half of the tile's columns
constexpr int EPI_HALVES = EPI_WARPS / 4;
constexpr int EPI_COLS = BLOCK_N / EPI_HALVES;  // 128 columns per epilogue warp
constexpr int THREADS = (EPI_WARP0 + EPI_WARPS) * 32;  // 384
constexpr uint32_t A_STAGE_BYTES = BLOCK_Q * BLOCK_K * 2;  // 16 KB

template <int CG> struct Cfg {
    static constexpr int LOAD_N = BLOCK_N / CG;                       // corpus rows each CTA loads
    static constexpr uint32_t B_STAGE_BYTES = LOAD_N * BLOCK_K * 2;   // 32 KB / 16 KB
    static constexpr uint32_t STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int STAGES = (CG == 1) ? 4 : 6;                  // 192 KB of operand ring
    static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

// ----------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
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
// arrive on the barrier at the same smem offset in CTA `cta` of the cluster (works for cta == self)
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
    asm volatile(
        "{\n"
        ".reg .b32 remAddr32;\n"
        "mapa.shared::cluster.u32 remAddr32, %0, %1;\n"
        "mbarrier.arrive.shared::cluster.b64 _, [remAddr32];\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(cta) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// one lane of a converged warp (the instruction needs warp-uniform control flow around it)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void prefetch_tmap(const void* p) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(p) : "memory");
}
template <int CG>
__device__ __forceinline__ void tma_load_2d(const void* tmap, uint64_t* bar, void* smem, int c0, int c1, uint64_t hint) {
    if constexpr (CG == 1) {
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
            " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(smem)),
            "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(hint)
            : "memory");
    } else {
        // both CTAs of the pair execute this; clearing the peer bit routes complete_tx to the leader's barrier
        asm volatile(
            "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
            " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(smem)),
            "l"(tmap), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "l"(hint)
            : "memory");
    }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int CG> __device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    if constexpr (CG == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
}
template <int CG> __device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
    if constexpr (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
template <int CG>
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if constexpr (CG == 1) {
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}
// tcgen05.commit: the mbarrier gets one arrival once all MMAs issued so far by this thread are done
template <int CG> __device__ __forceinline__ void umma_commit(uint64_t* bar) {
    if constexpr (CG == 1) {
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    } else {
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
    }
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// if (s > thr) { list[cnt++] = ((rbase + J) << 32 | bits(s)); }  — predicated, no branch
template <int J>
__device__ __forceinline__ void push_if_gt_imm(unsigned long long* list, uint32_t& cnt, uint32_t s_bits, float thr, uint32_t rbase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 a;\n"
        ".reg .b32 r;\n"
        "setp.gt.f32 p, %2, %3;\n"
        "mad.wide.u32 a, %0, 8, %1;\n"   // temporaries are written unconditionally: a predicated write would
        "add.u32 r, %5, %6;\n"           // make ptxas keep their old values alive across the whole loop
        "@p st.global.cg.v2.b32 [a], {%4, r};\n"
        "@p add.u32 %0, %0, 1;\n"
        "}\n"
        : "+r"(cnt)
        : "l"(list), "f"(__uint_as_float(s_bits)), "f"(thr), "r"(s_bits), "r"(rbase), "n"(J)
        : "memory");
}
// banded variant: additionally requires (s, row) to rank strictly AFTER the bound (ubs, ubr):
// s < ubs, or s == ubs and row > ubr
template <int J>
__device__ __forceinline__ void push_if_banded_imm(unsigned long long* list, uint32_t& cnt, uint32_t s_bits, float thr, uint32_t rbase,
                                                   float ubs, uint32_t ubr) {
    asm volatile(
        "{\n"
        ".reg .pred p, lt, eq, gt;\n"
        ".reg .b64 a;\n"
        ".reg .b32 r;\n"
        "add.u32 r, %5, %6;\n"
        "setp.gt.f32 p, %2, %3;\n"
        "setp.lt.f32 lt, %2, %7;\n"
        "setp.eq.f32 eq, %2, %7;\n"
        "setp.gt.u32 gt, r, %8;\n"
        "and.pred eq, eq, gt;\n"
        "or.pred lt, lt, eq;\n"
        "and.pred p, p, lt;\n"
        "mad.wide.u32 a, %0, 8, %1;\n"
        "@p st.global.cg.v2.b32 [a], {%4, r};\n"
        "@p add.u32 %0, %0, 1;\n"
        "}\n"
        : "+r"(cnt)
        : "l"(list), "f"(__uint_as_float(s_bits)), "f"(thr), "r"(s_bits), "r"(rbase), "n"(J), "f"(ubs), "r"(ubr)
        : "memory");
}
template <int J0, int J1>
__device__ __forceinline__ void push_block_banded(unsigned long long* list, uint32_t& cnt, const uint32_t (&v)[32], float thr, uint32_t rbase,
                                                  int lim, float ubs, uint32_t ubr) {
    if constexpr (J0 < J1) {
        push_if_banded_imm<J0>(list, cnt, v[J0], (J0 < lim) ? thr : INFINITY, rbase, ubs, ubr);
        push_block_banded<J0 + 1, J1>(list, cnt, v, thr, rbase, lim, ubs, ubr);
    }
}
template <int J0, int J1>
__device__ __forceinline__ void push_block(unsigned long long* list, uint32_t& cnt, const uint32_t (&v)[32], float thr, uint32_t rbase, int lim) {
    if constexpr (J0 < J1) {
        push_if_gt_imm<J0>(list, cnt, v[J0], (J0 < lim) ? thr : INFINITY, rbase);
        push_block<J0 + 1, J1>(list, cnt, v, thr, rbase, lim);
    }
}

// K-major, 128-byte-swizzled operand tile: 8-row groups are 1024 B apart (SBO), version 1 (sm_100)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;                    // LBO (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;          // SBO
    d |= (uint64_t)1 << 46;                    // descriptor version
    d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
    return d;
}

// Corpus tiles of residue class `g` in this pass.  The G classes are grouped into slabs of `cps` consecutive
// classes; slab s owns a contiguous range of tiles (proportional to its number of classes) and its classes
// interleave inside it.  With one round of work items cps == G (all clusters sweep the whole pass together);
// with several rounds each round sweeps its own slab, so every corpus byte is pulled from HBM once.
struct TileRange { int first, end, stride; };
// virtual tile index of the pass -> physical corpus tile (see TcParams::map_stride)
__device__ __forceinline__ int phys_tile(const TcParams& p, int t) {
    int j = t - p.tile_begin;
    if (p.map_ratio) j = (j / (p.map_ratio - 1)) * p.map_ratio + (j % (p.map_ratio - 1)) + 1;
    return j * p.map_stride;
}
__device__ __forceinline__ TileRange class_tiles(const TcParams& p, int g) {
    const int s = g / p.classes_per_slab;
    const int c0 = s * p.classes_per_slab;
    const int c1 = min(p.n_chunks, c0 + p.classes_per_slab);
    const long long span = (long long)(p.n_tiles - p.tile_begin);
    TileRange r;
    r.stride = c1 - c0;
    r.first = p.tile_begin + (int)(span * c0 / p.n_chunks) + (g - c0);
    r.end = p.tile_begin + (int)(span * c1 / p.n_chunks);
    return r;
}

// ----------------------------------------------------------------------------- the kernel
template <int CG>
__global__ void __launch_bounds__(THREADS, 1)
tc_topk_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_x, TcParams p) {
    using C = Cfg<CG>;
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment (128B swizzle) computed in the shared window so both CTAs of a pair get the same offsets
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + C::STAGES * A_STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + C::STAGES;
    uint64_t* tmem_full_bar = bars + 2 * C::STAGES;
    uint64_t* tmem_empty_bar = bars + 2 * C::STAGES + ACC_STAGES;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 2 * ACC_STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = (CG == 1) ? 0u : cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x / CG;
    const int n_clusters = gridDim.x / CG;

    if (CG > 1) cluster_sync_all();
    if (p.pdl && threadIdx.x == 0) pdl_launch_dependents();  // the pass's merge kernel may become resident now (it waits for us)
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_q);
        prefetch_tmap(&tmap_x);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < C::STAGES; i++) {
            mbar_init(&full_bar[i], CG);   // one arrival per CTA producer (all land on the leader)
            mbar_init(&empty_bar[i], 1);   // one tcgen05.commit
        }
        for (int i = 0; i < ACC_STAGES; i++) {
            mbar_init(&tmem_full_bar[i], 1);        // one tcgen05.commit
            mbar_init(&tmem_empty_bar[i], CG * EPI_WARPS);  // one arrival per epilogue warp of each CTA
        }
        fence_barrier_init();
    } else if (warp == 2) {
        tmem_alloc<CG>(tmem_ptr_smem, 512);
    }
    tc_fence_before();
    if (CG > 1) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const int num_kb = p.num_kb;
    const int n_items = p.m_tiles * p.n_chunks;

    if (warp == 0) {
        // ===================== TMA producer (lane 0 issues; both CTAs of a pair) =====================
        // L2 sharing window.  The clusters that work on the SAME corpus tiles (same residue class, different query
        // tiles, same round) pull every tile through L2; if they drift apart by more than what L2 retains, each
        // of them fetches the tile from HBM again (measured 2.5x..15x the algorithmic DRAM traffic).  The leader
        // CTA's producer warp therefore publishes its tile counter every SYNC_EVERY tiles and does not run more
        // than SYNC_WINDOW (6) tiles ahead of the slowest member of its group.  The slowest member never waits, so the
        // protocol cannot deadlock; a bounded spin makes a lost update harmless (it only costs locality).
        constexpr int SYNC_MAX_SPINS = 50000;
        const int SYNC_EVERY = p.sync_every, SYNC_WINDOW = p.sync_window;
        uint32_t stage = 0, phase = 0;
        for (int item = cluster_id; item < n_items; item += n_clusters) {
            const int m_tile = item % p.m_tiles, chunk = item / p.m_tiles;
            const int q_row0 = (m_tile * CG + (int)rank) * BLOCK_Q;
            const TileRange tr = class_tiles(p, chunk);
            const int round = item / n_clusters;
            const int g0 = max(chunk * p.m_tiles, round * n_clusters);
            const int g1 = min(min((chunk + 1) * p.m_tiles, (round + 1) * n_clusters), n_items);
            const bool synced = leader && p.progress != nullptr && (g1 - g0) > 1;
            int j = 0;
            for (int t = tr.first; t < tr.end; t += tr.stride, j++) {
                if (synced && (j % SYNC_EVERY) == 0) {
                    if (lane == 0) asm volatile("st.volatile.global.s32 [%0], %1;" ::"l"(p.progress + item), "r"(j) : "memory");
                    const int need = j - SYNC_WINDOW;
                    if (need > 0) {
                        for (int spin = 0; spin < SYNC_MAX_SPINS; spin++) {
                            int mn = 0x7fffffff;
                            for (int i = g0 + lane; i < g1; i += 32) {
                                int v;
                                asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p.progress + i) : "memory");
                                mn = min(mn, v);
                            }
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) mn = min(mn, __shfl_xor_sync(VQ_FULL_MASK, mn, o));
                            if (mn >= need) break;
                            __nanosleep(100);
                        }
                    }
                }
                // the whole warp runs the k loop with warp-uniform state; one elected lane issues the copies, so the
                // descriptors, coordinates and barrier addresses stay in uniform registers
                const int x_row0 = phys_tile(p, t) * BLOCK_N + (int)rank * C::LOAD_N;
                for (int kb = 0; kb < num_kb; kb++) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (elect_one()) {
                        tma_load_2d<CG>(&tmap_q, &full_bar[stage], smem_a + stage * A_STAGE_BYTES, kb * BLOCK_K, q_row0,
                                        0x14F0000000000000ull /*evict last: queries are re-read by every tile*/);
                        tma_load_2d<CG>(&tmap_x, &full_bar[stage], smem_b + stage * C::B_STAGE_BYTES, kb * BLOCK_K, x_row0,
                                        0x1000000000000000ull /*evict normal*/);
                        if (leader) mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES * CG);
                        else mbar_arrive_cluster(&full_bar[stage], 0);
                    }
                    __syncwarp();
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
            }
            // done with this item: never hold the others back
            if (synced && lane == 0) asm volatile("st.volatile.global.s32 [%0], %1;" ::"l"(p.progress + item), "r"(0x7fffffff) : "memory");
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer (warp 1 of the leader CTA) =====================
        // All 32 lanes run the loop with warp-uniform state and ONE elected lane issues the MMAs of a k block: shared
        // memory descriptors, the TMEM address and the barrier addresses then live in uniform registers.  (Issued from
        // inside an `if (lane == 0)` region the same loop costs ~115 SASS instructions per k block — a waterfall of
        // ELECT / R2UR.BROADCAST around every UTCHMMA — which is as long as the 4 MMAs themselves take: the tensor
        // pipe sat at 84 % active, paced by its own issuer.)
        if (leader) {
            const uint32_t idesc = p.idesc_base | ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)((BLOCK_Q * CG) >> 4) << 24);
            const uint64_t adesc0 = make_smem_desc(smem_u32(smem_a));
            const uint64_t bdesc0 = make_smem_desc(smem_u32(smem_b));
            uint32_t stage = 0, phase = 0, iter = 0;
            for (int item = cluster_id; item < n_items; item += n_clusters) {
                const int chunk = item / p.m_tiles;
                const TileRange tr = class_tiles(p, chunk);
                for (int t = tr.first; t < tr.end; t += tr.stride, iter++) {
                    const uint32_t as = iter & 1, aphase = (iter >> 1) & 1;
                    mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + as * BLOCK_N;
                    for (int kb = 0; kb < num_kb; kb++) {
                        mbar_wait(&full_bar[stage], phase);
                        tc_fence_after();
                        if (elect_one()) {
                            // stage s of an operand ring starts s * STAGE_BYTES further: + (bytes >> 4) in the address field
                            const uint64_t adesc = adesc0 + (uint64_t)(stage * (A_STAGE_BYTES >> 4));
                            const uint64_t bdesc = bdesc0 + (uint64_t)(stage * (C::B_STAGE_BYTES >> 4));
                            // +32 B per UMMA_K inside the swizzle atom = +2 in the (addr >> 4) field
                            umma_f16<CG>(tmem_d, adesc, bdesc, idesc, (uint32_t)(kb != 0));
#pragma unroll
                            for (int k4 = 1; k4 < BLOCK_K / UMMA_K; k4++)
                                umma_f16<CG>(tmem_d, adesc + (uint64_t)(k4 * 2), bdesc + (uint64_t)(k4 * 2), idesc, 1u);
                            umma_commit<CG>(&empty_bar[stage]);                       // frees the smem slot (both CTAs)
                            if (kb == num_kb - 1) umma_commit<CG>(&tmem_full_bar[as]);  // accumulator ready (both CTAs)
                        }
                        __syncwarp();
                        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp >= EPI_WARP0) {
        // ===================== epilogue: TMEM -> registers -> threshold filter -> candidate lists =====================
        const int ew = (warp - EPI_WARP0) & 3;    // == warp % 4: the TMEM lane quarter this warp may read
        const int half = (warp - EPI_WARP0) >> 2;  // which EPI_COLS-wide slice of the tile's columns
        const uint32_t lane_field = (uint32_t)(ew * 32) << 16;
        const int k = p.k;
        uint32_t iter = 0;
        // Programmatic dependent launch: this kernel may have started while the previous pass's merge was still running.
        // The operands (queries, corpus) were complete long before; what the merge produces — the primed thresholds — and
        // what it still reads — the candidate lists this epilogue is about to overwrite — are touched only after this wait.
        // Producer and MMA warps never wait: they stream the first tiles into TMEM meanwhile.
        if (p.pdl) pdl_wait();
        for (int item = cluster_id; item < n_items; item += n_clusters) {
            const int m_tile = item % p.m_tiles, chunk = item / p.m_tiles;
            const int q = (m_tile * CG + (int)rank) * BLOCK_Q + ew * 32 + lane;  // this thread's query
            const int q0_warp = q - lane;
            const size_t list_set = (size_t)p.list_set_base + (size_t)chunk * EPI_HALVES + half;  // candidate lists are per (residue class, column half)
            uint64_t* warp_lists = p.lists + (list_set * p.nq_pad + q0_warp) * VQ_CAP;
            unsigned long long* const my_list = reinterpret_cast<unsigned long long*>(warp_lists + (size_t)lane * VQ_CAP);
            uint32_t cnt = 0;
            const bool active = q < p.nq;
            // primed threshold: the k-th best score of an already-searched sample of the corpus (rows below
            // tile_begin * 256), a valid lower bound of the final k-th score; ties lose (higher row id)
            float thr = active ? (p.init_thr ? p.init_thr[(size_t)q * p.init_thr_stride] : -FLT_MAX) : INFINITY;
            const bool banded = p.ub_score != nullptr;
            const float ubs = (banded && active) ? p.ub_score[q] : 0.0f;
            const uint32_t ubr = (banded && active) ? p.ub_row[q] : 0u;
            const TileRange tr = class_tiles(p, chunk);
            for (int t = tr.first; t < tr.end; t += tr.stride, iter++) {
                const uint32_t as = iter & 1, aphase = (iter >> 1) & 1;
                mbar_wait(&tmem_full_bar[as], aphase);
                tc_fence_after();
                const uint32_t row0 = (uint32_t)phys_tile(p, t) * BLOCK_N + half * EPI_COLS;
                const int n_valid = (int)min((long long)EPI_COLS, p.n_rows - (long long)row0);  // may be <= 0
                const uint32_t tcol0 = tmem_base + lane_field + as * BLOCK_N + half * EPI_COLS;
#pragma unroll 1
                for (int cb = 0; cb < EPI_COLS / 32; cb++) {
                    // make room first (nothing else is live across the call): a block appends at most 32 pairs
                    unsigned need = __ballot_sync(VQ_FULL_MASK, cnt > VQ_CAP - 32);
                    while (need) {
                        const int l = __ffs(need) - 1;
                        need &= need - 1;
                        const int n_l = __shfl_sync(VQ_FULL_MASK, (int)cnt, l);
                        float thr_l;
                        const int m_l = warp_prune_list(warp_lists + (size_t)l * VQ_CAP, n_l, k, lane, thr_l);
                        // never loosen a primed threshold
                        if (lane == l) { cnt = (uint32_t)m_l; thr = fmaxf(thr, thr_l); }
                    }
                    uint32_t v[32];
                    tmem_ld32(tcol0 + cb * 32, v);
                    tmem_ld_wait();
                    if (cb == EPI_COLS / 32 - 1) {
                        // all of this warp's reads of the accumulator are done: hand it back to the MMA warp
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if (CG == 1) mbar_arrive(&tmem_empty_bar[as]);
                            else mbar_arrive_cluster(&tmem_empty_bar[as], 0);
                        }
                    }
                    // fast reject: the block's maximum against the threshold (NaN never wins an fmaxf)
                    float mx = fmaxf(__uint_as_float(v[0]), __uint_as_float(v[1]));
#pragma unroll
                    for (int j = 2; j < 32; j++) mx = fmaxf(mx, __uint_as_float(v[j]));
                    if (__any_sync(VQ_FULL_MASK, mx > thr)) {
                        const uint32_t rbase = row0 + cb * 32;
                        // rows past the end of the corpus are TMA zero fill: lim masks them (32 = all valid)
                        const int lim = min(32, n_valid - cb * 32);
                        if (banded) push_block_banded<0, 32>(my_list, cnt, v, thr, rbase, lim, ubs, ubr);
                        else if (lim == 32) push_block<0, 32>(my_list, cnt, v, thr, rbase, 32);
                        else push_block<0, 32>(my_list, cnt, v, thr, rbase, lim);
                    }
                }
            }
            // end of the work item: publish the (unsorted) list length; the pass's merge kernel sorts
            p.counts[list_set * p.nq_pad + q] = active ? (int)cnt : 0;
        }
    }

    // ----- teardown: every async op has been consumed by now (the epilogue waited on the last commit)
    tc_fence_before();
    if (CG > 1) cluster_sync_all(); else __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<CG>(tmem_base, 512);
    }
}

// ----------------------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess || !ptr) return nullptr;
        fn = reinterpret_cast<PFN_encodeTiled>(ptr);
    }
    return fn;
}

// rows x cols 16-bit matrix, row stride ld elements; box = 64 cols x box_rows, 128B swizzle, OOB -> 0
static int encode_2d(CUtensorMap* m, const void* base, int dtype, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return -1;
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMapDataType dt = (dtype == VQ_DTYPE_BF16) ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    CUresult r = enc(m, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -(int)r - 1000;
}

template <int CG>
static int launch_cg(const CUtensorMap& tq, const CUtensorMap& tx, const TcParams& p, int n_clusters, cudaStream_t stream, bool pdl_launch) {
    using C = Cfg<CG>;
    static bool attr_set_dev[64] = {};  // cudaFuncSetAttribute is per device
    int dev_ = 0;
    cudaGetDevice(&dev_);
    bool& attr_set = attr_set_dev[dev_ & 63];
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tc_topk_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES);
        if (e != cudaSuccess) return -(int)e;
        attr_set = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(n_clusters * CG));
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (pdl_launch) {  // may start before the preceding kernel (the previous pass's merge) has finished
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.numAttrs = 2;
    }
    cudaError_t e = cudaLaunchKernelEx(&cfg, tc_topk_kernel<CG>, tq, tx, p);
    return e == cudaSuccess ? 0 : -(int)e;
}

}  // namespace tc

// Plans the (query tile, residue class) decomposition for nq queries over n_rows corpus rows on n_sms SMs.
void tc_plan(int nq, long long n_rows, int n_sms, int force_cg, TcPlan* plan) {
    int cg = force_cg;
    if (cg != 1 && cg != 2) cg = (nq > tc::BLOCK_Q) ? 2 : 1;
    const int n_clusters_max = n_sms / cg;
    const int m_tiles = (nq + tc::BLOCK_Q * cg - 1) / (tc::BLOCK_Q * cg);
    const int n_tiles = (int)((n_rows + tc::BLOCK_N - 1) / tc::BLOCK_N);
    // number of residue classes G: the m_tiles * G items should spread evenly over the clusters;
    // cost model: the slowest cluster runs rounds * ceil(n_tiles / G) tiles.  Best efficiency wins; a smaller G only
    // when it is within 0.5 % (fewer candidate lists for the merge to scan).  HBM-bound batches stream the corpus at
    // ~7 TB/s, where every percent of imbalance is a percent of latency.
    int best_c = 1;
    double best_eff = 0.0;
    int c_max = n_tiles < 4 * n_clusters_max ? (n_tiles > 0 ? n_tiles : 1) : 4 * n_clusters_max;
    if (c_max > 448) c_max = 448;  // 2 list sets per class, the merge kernel takes at most 1023 sets
    for (int c = 1; c <= c_max; c++) {
        const int tpc = (n_tiles + c - 1) / c;
        const long long items = (long long)m_tiles * c;
        const long long rounds = (items + n_clusters_max - 1) / n_clusters_max;
        const double eff = (double)m_tiles * n_tiles / ((double)rounds * tpc * n_clusters_max);
        if (eff > best_eff + 0.005) { best_eff = eff; best_c = c; }
    }
    plan->cg = cg;
    plan->m_tiles = m_tiles;
    plan->n_tiles = n_tiles;
    plan->n_chunks = best_c;
    plan->tiles_per_chunk = n_tiles > 0 ? (n_tiles + best_c - 1) / best_c : 1;
    {   // classes per slab: one round -> a single slab; several rounds -> the classes a round can hold
        const long long items_all = (long long)m_tiles * best_c;
        plan->classes_per_slab = items_all <= n_clusters_max ? best_c : (n_clusters_max + m_tiles - 1) / m_tiles;
        if (plan->classes_per_slab < 1) plan->classes_per_slab = 1;
        if (plan->classes_per_slab > best_c) plan->classes_per_slab = best_c;
    }
    plan->list_sets = best_c * tc::EPI_HALVES;
    plan->nq_pad = m_tiles * tc::BLOCK_Q * cg;
    const long long items = (long long)plan->m_tiles * plan->n_chunks;
    plan->n_clusters = (int)(items < n_clusters_max ? items : n_clusters_max);
    if (plan->n_clusters < 1) plan->n_clusters = 1;
}

// q16: [nq, ld_q] 16-bit queries (device); x16: [n_rows, ld_x] 16-bit corpus (device)
int tc_topk_launch(const void* q16, int nq, int ld_q, const void* x16, long long n_rows, int ld_x, int d, int dtype, int k,
                   const TcPlan& plan, const TcPass& pass, uint64_t* lists, int* counts, int* progress, cudaStream_t stream) {
    if (k < 1 || k > VQ_KMAX) return -2;
    if (n_rows <= 0 || nq <= 0) return 0;
    CUtensorMap tq, tx;
    int rc = tc::encode_2d(&tq, q16, dtype, (uint64_t)plan.nq_pad, (uint64_t)d, (uint64_t)ld_q, tc::BLOCK_Q);  // q16 holds nq_pad rows
    if (rc) return rc;
    rc = tc::encode_2d(&tx, x16, dtype, (uint64_t)n_rows, (uint64_t)d, (uint64_t)ld_x, (uint32_t)(tc::BLOCK_N / plan.cg));
    if (rc) return rc;
    if (plan.m_tiles < 2) progress = nullptr;  // one query tile: nobody shares a corpus tile, no window (and no memset node)
    if (progress) {  // per-item tile counters of the L2 sharing window, zeroed for every pass
        cudaError_t e = cudaMemsetAsync(progress, 0, (size_t)plan.m_tiles * plan.n_chunks * sizeof(int), stream);
        if (e != cudaSuccess) return -(int)e;
    }
    TcParams p;
    p.lists = lists;
    p.counts = counts;
    p.n_rows = n_rows;
    p.nq = nq;
    p.nq_pad = plan.nq_pad;
    p.k = k;
    p.num_kb = (d + tc::BLOCK_K - 1) / tc::BLOCK_K;
    p.m_tiles = plan.m_tiles;
    p.n_tiles = pass.tile_end;
    p.tile_begin = pass.tile_begin;
    p.map_stride = pass.map_stride;
    p.map_ratio = pass.map_ratio;
    p.list_set_base = pass.list_set_base;
    p.init_thr = pass.init_thr;
    p.init_thr_stride = pass.init_thr_stride;
    p.ub_score = pass.ub_score;
    p.ub_row = pass.ub_row;
    p.n_chunks = plan.n_chunks;
    p.tiles_per_chunk = plan.tiles_per_chunk;
    p.classes_per_slab = plan.classes_per_slab;
    p.progress = progress;
    {   // window of the L2 sharing protocol (tiles); tunable for experiments through the environment
        static int every = 0, window = 0;
        if (!every) {
            const char* e = getenv("VQ_SYNC_EVERY");
            const char* w = getenv("VQ_SYNC_WINDOW");
            every = e ? atoi(e) : 4;
            window = w ? atoi(w) : 6;  // measured r2: 6 beats 8 by 3.5 % at nq = 1250 (fewer DRAM re-reads under the power cap), flat at 4096
            if (every < 1) every = 1;
            if (window < 1) window = 1;
        }
        p.sync_every = every;
        p.sync_window = window;
    }
    // instruction descriptor: D = f32 (bits 4-5 = 1), A/B format bf16 = 1 / f16 = 0 (bits 7-9, 10-12), both K-major
    const uint32_t fmt = (dtype == VQ_DTYPE_BF16) ? 1u : 0u;
    p.idesc_base = (1u << 4) | (fmt << 7) | (fmt << 10);
    p.pdl = pass.pdl != 0;
    const bool pdl_launch = pass.pdl == 2;
    return plan.cg == 2 ? tc::launch_cg<2>(tq, tx, p, plan.n_clusters, stream, pdl_launch) : tc::launch_cg<1>(tq, tx, p, plan.n_clusters, stream, pdl_launch);
}

}  // namespace vq
