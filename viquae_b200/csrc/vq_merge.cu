// K4 — k-way merges of candidate lists, one warp per query: selection of the top 128 (bitwise search for the 128-th key,
// ballot compaction, one 128-key register sort) instead of sorting everything.
//
//  * merge_lists: folds the per-(corpus slab, query) candidate-key lists written by K1 / K3 into
//    the final [nq, k] (scores, ids) the reference returns from faiss_index.search
//    (datasets/search.py:384-385: float32 scores, int64 ids, descending, -1 padding).
//  * merge_topk: the cross-GPU step — every rank's local top-k [G, nq, k] (after the NCCL
//    all-gather) -> global top-k.  Top-k of a union == top-k of the per-shard top-k's.
// Both use the engine-wide tie rule: equal scores -> lower id first.  ids must be < 2^32.
#include "vq_common.cuh"
#include "vq_kernels.h"

#include <stdlib.h>

namespace vq {

// largest float strictly below x (x finite; -FLT_MAX stays)
__device__ __forceinline__ float next_below(float x) {
    if (x == 0.0f) return -1.401298464e-45f;
    const uint32_t u = __float_as_uint(x);
    if (x > 0.0f) return __uint_as_float(u - 1);
    if (x <= -FLT_MAX) return x;
    return __uint_as_float(u + 1);
}

constexpr int MERGE_THREADS = 256;
constexpr int MERGE_WARPS = MERGE_THREADS / 32;
constexpr int MERGE_CAP_MAX = 8192;  // 64 KB of keys
constexpr int MERGE_MAX_SETS = 1024;

// a, b: 128 keys each, sorted descending, 4 per lane (element e = r*32 + lane).  a <- the best 128 of the
// union, sorted descending.  max(a[e], b[127-e]) is a bitonic sequence holding exactly those 128 keys;
// one bitonic merge (strides 64, 32 in registers, 16..1 by shuffle) sorts it.
__device__ __forceinline__ void merge_top128(uint64_t (&a)[4], const uint64_t (&b)[4], int lane) {
#pragma unroll
    for (int r = 0; r < 4; r++) a[r] = umax64(a[r], __shfl_sync(VQ_FULL_MASK, b[3 - r], 31 - lane));
#pragma unroll
    for (int r = 0; r < 2; r++) {  // stride 64
        const uint64_t x = a[r], y = a[r + 2];
        a[r] = umax64(x, y);
        a[r + 2] = umin64(x, y);
    }
#pragma unroll
    for (int r = 0; r < 4; r += 2) {  // stride 32
        const uint64_t x = a[r], y = a[r + 1];
        a[r] = umax64(x, y);
        a[r + 1] = umin64(x, y);
    }
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) {
        const bool lower = (lane & j) == 0;
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const uint64_t other = __shfl_xor_sync(VQ_FULL_MASK, a[r], j);
            a[r] = lower ? umax64(a[r], other) : umin64(a[r], other);
        }
    }
}

// ---- Merge step of one pass, one WARP per query, by SELECTION.  Candidate lists are [set][nq_pad][VQ_CAP] raw pairs +
// counts[set][nq_pad]; set 0 is the RUNNING merged top-k of all passes so far.  Takes set 0 (if use_set0) + sets
// [set_begin, set_begin + n_sets) in batches of whole lists (<= 1024 keys, 32 per lane), reduces every batch to its top 128
// (warp_select_top128: bitwise search for the 128-th key + one 128-key sort) and folds it into the running top-128; writes
// the best k back to set 0 (sorted raw pairs + count), the k-th score to thr_out (the primed threshold of the next pass)
// and, for the last pass, the final (scores, ids).  No block-level barrier: the warps of a CTA are independent.
// (Round 1/early round 2 sorted every 256-key chunk in registers: 20.6 K instructions per query, ALU-pipe bound at 73 %;
// selection needs ~1/5 of that: 50 -> 27 us per merge at 1,250 queries, 123 -> 74 us for the unprimed pass.)
constexpr int SEL_R = 32;
constexpr int SEL_BATCH = 32 * SEL_R;
constexpr int SEL_WARPS = 4;

__global__ void __launch_bounds__(SEL_WARPS * 32)
merge_select_kernel(uint64_t* __restrict__ lists, int* __restrict__ counts, int use_set0, int set_begin, int n_sets, int nq_pad, int nq, int k,
                    long long id_offset, float* __restrict__ thr_out, float* __restrict__ out_scores, long long* __restrict__ out_ids) {
    extern __shared__ uint64_t sel_smem[];                       // per warp: owner ids (u16 [SEL_BATCH] = 256 u64) + [128] slots
    __shared__ int off_all[SEL_WARPS][MERGE_MAX_SETS + 2];        // exclusive prefix sums of the list lengths
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // programmatic dependent launch (no-ops on a plain launch): the next pass's contraction may become resident and stream
    // its first tiles while this merge runs (its epilogue waits for us); we wait for the contraction that filled the lists
    if (threadIdx.x == 0) pdl_launch_dependents();
    pdl_wait();
    const int q = blockIdx.x * SEL_WARPS + warp;
    if (q >= nq) return;
    uint64_t* slots = sel_smem + (size_t)warp * (256 + 128);
    int* off = off_all[warp];
    const int L = n_sets + 1;  // logical list l: 0 = running merged list (set 0), l >= 1 = set_begin + l - 1
    // list lengths: every load is issued before the first one is used (the kernel is latency-bound on these round trips)
    constexpr int CPL = (MERGE_MAX_SETS + 1 + 31) / 32;  // lists per lane (covers l = 0 .. MERGE_MAX_SETS), lane-major: lane owns l = lane * CPL + j
    int c[CPL];
#pragma unroll
    for (int j = 0; j < CPL; j++) {
        const int l = lane * CPL + j;
        const int set = l == 0 ? 0 : set_begin + l - 1;
        c[j] = (l >= L || (l == 0 && !use_set0)) ? 0 : __ldcg(counts + (size_t)set * nq_pad + q);
    }
    int mine = 0;
#pragma unroll
    for (int j = 0; j < CPL; j++) mine += c[j];
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(VQ_FULL_MASK, incl, o);
        if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(VQ_FULL_MASK, incl, 31);
    {
        int before = incl - mine;
#pragma unroll
        for (int j = 0; j < CPL; j++) {
            const int l = lane * CPL + j;
            if (l <= L) off[l] = before;
            before += c[j];
        }
    }
    __syncwarp();
    uint16_t* owner = reinterpret_cast<uint16_t*>(slots);  // [SEL_BATCH] list of every staged element (reuses nothing live: slots are written later)
    uint64_t run[4] = {0ull, 0ull, 0ull, 0ull};
    int la = 0;
    while (la < L && off[la] < total) {
        const int base = off[la];
        int lo = la + 1, hi = L;  // the most whole lists that fit in one batch (a list holds at most VQ_CAP <= SEL_BATCH keys)
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (off[mid] - base <= SEL_BATCH) lo = mid; else hi = mid - 1;
        }
        const int lb = lo, n = off[lb] - base;
        // which list does element e of the batch come from?  one lane per list writes its id over the list's span
        // (shared-memory stores only), so that the global loads below are independent and all in flight at once
        for (int l = la + lane; l < lb; l += 32) {
            const int e0 = off[l] - base, e1 = off[l + 1] - base;
            for (int e = e0; e < e1; e++) owner[e] = (uint16_t)l;
        }
        __syncwarp();
        uint64_t key[SEL_R];
#pragma unroll
        for (int r = 0; r < SEL_R; r++) {
            const int e = r * 32 + lane;
            key[r] = 0ull;
            if (e < n) {
                const int l = owner[e];
                const int set = l == 0 ? 0 : set_begin + l - 1;
                key[r] = __ldcg(reinterpret_cast<const unsigned long long*>(lists + ((size_t)set * nq_pad + q) * VQ_CAP) + (base + e - off[l]));
            }
        }
#pragma unroll
        for (int r = 0; r < SEL_R; r++) key[r] = (r * 32 + lane) < n ? raw_to_key(key[r]) : 0ull;
        __syncwarp();
        uint64_t top[4];
        warp_select_top128<SEL_R>(key, top, slots + 256, lane);
        merge_top128(run, top, lane);
        la = lb;
    }
    const int have = total < k ? total : k;
    uint64_t* dst = lists + (size_t)q * VQ_CAP;  // set 0
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int e = r * 32 + lane;
        if (e < have) dst[e] = key_to_raw(run[r]);
        // the next pass samples rows with LOWER ids too, so a score equal to the k-th must still pass its strict
        // compare: hand out the next float below the k-th score
        if (e == k - 1 && thr_out) thr_out[q] = total >= k ? next_below(key_score(run[r])) : -FLT_MAX;
        if (out_scores && e < k) {
            out_scores[(size_t)q * k + e] = e < have ? key_score(run[r]) : -FLT_MAX;
            out_ids[(size_t)q * k + e] = e < have ? (long long)key_row(run[r]) + id_offset : -1;
        }
    }
    if (lane == 0) counts[q] = have;
}

// Cross-GPU merge (NCCL transport; the peer-memory transport fuses it with the exchange, vq_xchg.cu): per query, the G
// per-shard top-k lists [n_lists][nq][k_in] -> global top-k_out, one warp per query, whole lists in batches of <= 1024
// keys reduced by selection (warp_select_top128).
__global__ void __launch_bounds__(SEL_WARPS * 32)
merge_topk_kernel(const float* __restrict__ scores, const long long* __restrict__ ids, long long stride_s, long long stride_i, int n_lists,
                  int nq, int k_in, int k_out, float* __restrict__ out_scores, long long* __restrict__ out_ids) {
    extern __shared__ uint64_t mt_smem[];  // per warp: [128] slots of the selection
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.x * SEL_WARPS + warp;
    if (q >= nq) return;
    uint64_t* slots = mt_smem + (size_t)warp * 128;
    const int per = k_in <= SEL_BATCH ? SEL_BATCH / k_in : 1;  // whole lists per batch
    uint64_t run[4] = {0ull, 0ull, 0ull, 0ull};
    for (int c0 = 0; c0 < n_lists; c0 += per) {
        const int c1 = min(n_lists, c0 + per);
        for (int j0 = 0; j0 < k_in; j0 += SEL_BATCH) {  // (a single list longer than a batch is taken in pieces)
            const int kk = c1 - c0 > 1 ? k_in : min(SEL_BATCH, k_in - j0);
            const int n = (c1 - c0) * kk;
            // all loads of the batch in flight before the first use
            uint64_t key[SEL_R];
            uint32_t sc[SEL_R];
#pragma unroll
            for (int r = 0; r < SEL_R; r++) {
                const int e = r * 32 + lane;
                key[r] = ~0ull;  // id -1
                sc[r] = 0u;
                if (e < n) {
                    const int c = c0 + e / kk, j = j0 + e - (e / kk) * kk;
                    const size_t at = (size_t)q * k_in + j;  // list c starts stride_s / stride_i elements after list c-1
                    key[r] = (uint64_t)ids[(size_t)c * stride_i + at];
                    sc[r] = __float_as_uint(scores[(size_t)c * stride_s + at]);
                }
            }
#pragma unroll
            for (int r = 0; r < SEL_R; r++) {
                const long long id = (long long)key[r];
                key[r] = id >= 0 ? make_key(__uint_as_float(sc[r]), (uint32_t)id) : 0ull;  // padding sorts last
            }
            uint64_t top[4];
            warp_select_top128<SEL_R>(key, top, slots, lane);
            merge_top128(run, top, lane);
            if (c1 - c0 > 1) break;
        }
    }
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int e = r * 32 + lane;
        if (e < k_out) {
            const bool valid = run[r] != 0ull;
            out_scores[(size_t)q * k_out + e] = valid ? key_score(run[r]) : -FLT_MAX;
            out_ids[(size_t)q * k_out + e] = valid ? (long long)key_row(run[r]) : -1;
        }
    }
}

__global__ void band_scatter_kernel(const float* __restrict__ tmp_s, const long long* __restrict__ tmp_i, int nq, int kr, int k, int col0,
                                    long long id_offset, float* __restrict__ out_scores, long long* __restrict__ out_ids,
                                    float* __restrict__ ub_score, uint32_t* __restrict__ ub_row) {
    const int q = blockIdx.x;
    for (int e = threadIdx.x; e < kr; e += blockDim.x) {
        const float s = tmp_s[(size_t)q * kr + e];
        const long long id = tmp_i[(size_t)q * kr + e];
        out_scores[(size_t)q * k + col0 + e] = s;
        out_ids[(size_t)q * k + col0 + e] = id >= 0 ? id + id_offset : -1;
        if (e == kr - 1) {
            // next round takes what ranks strictly after this entry; a short round means the corpus is exhausted
            ub_score[q] = id >= 0 ? s : -INFINITY;
            ub_row[q] = id >= 0 ? (uint32_t)id : 0u;
        }
    }
}

int band_scatter_launch(const float* tmp_s, const long long* tmp_i, int nq, int kr, int k, int col0, long long id_offset,
                        float* out_scores, long long* out_ids, float* ub_score, uint32_t* ub_row, cudaStream_t stream) {
    if (nq <= 0) return 0;
    band_scatter_kernel<<<nq, 128, 0, stream>>>(tmp_s, tmp_i, nq, kr, k, col0, id_offset, out_scores, out_ids, ub_score, ub_row);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -(int)e;
}

// k_out > 128 (banded searches): gather every list of the query and sort them all in shared memory.
__global__ void __launch_bounds__(MERGE_THREADS)
merge_topk_big_kernel(const float* __restrict__ scores, const long long* __restrict__ ids, long long stride_s, long long stride_i, int n_lists,
                      int nq, int k_in, int k_out, float* __restrict__ out_scores, long long* __restrict__ out_ids, int S) {
    extern __shared__ uint64_t keys[];
    const int q = blockIdx.x;
    const int n = n_lists * k_in;
    for (int i = threadIdx.x; i < S; i += MERGE_THREADS) {
        uint64_t key = 0ull;
        if (i < n) {
            const int c = i / k_in, j = i - c * k_in;
            const size_t at = (size_t)q * k_in + j;
            const long long id = ids[(size_t)c * stride_i + at];
            if (id >= 0) key = make_key(scores[(size_t)c * stride_s + at], (uint32_t)id);
        }
        keys[i] = key;
    }
    __syncthreads();
    for (int s2 = 2; s2 <= S; s2 <<= 1)
        for (int j = s2 >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < (S >> 1); i += MERGE_THREADS) {
                const int lo = ((i & ~(j - 1)) << 1) | (i & (j - 1)), hi = lo | j;
                const bool desc = (lo & s2) == 0;
                const uint64_t a = keys[lo], b = keys[hi];
                if (desc ? (a < b) : (a > b)) { keys[lo] = b; keys[hi] = a; }
            }
            __syncthreads();
        }
    for (int e = threadIdx.x; e < k_out; e += MERGE_THREADS) {
        const uint64_t key = e < S ? keys[e] : 0ull;
        out_scores[(size_t)q * k_out + e] = key ? key_score(key) : -FLT_MAX;
        out_ids[(size_t)q * k_out + e] = key ? (long long)key_row(key) : -1;
    }
}

int merge_lists_launch(uint64_t* lists, int* counts, int use_set0, int set_begin, int n_sets, int nq_pad, int nq, int k, long long id_offset,
                       float* thr_out, float* out_scores, long long* out_ids, int expected_total, cudaStream_t stream, int pdl) {
    if (nq <= 0) return 0;
    if (n_sets + 1 > MERGE_MAX_SETS) return -2;
    {
        const size_t smem = (size_t)SEL_WARPS * (256 + 128) * sizeof(uint64_t);  // 12 KB
        static bool attr_sel[64] = {};
        int dev_s = 0;
        cudaGetDevice(&dev_s);
        if (!attr_sel[dev_s & 63]) {
            cudaError_t e = cudaFuncSetAttribute(merge_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return -(int)e;
            attr_sel[dev_s & 63] = true;
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)((nq + SEL_WARPS - 1) / SEL_WARPS));
        cfg.blockDim = dim3(SEL_WARPS * 32);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = pdl ? 1 : 0;
        cudaError_t e = cudaLaunchKernelEx(&cfg, merge_select_kernel, lists, counts, use_set0, set_begin, n_sets, nq_pad, nq, k, id_offset, thr_out,
                                           out_scores, out_ids);
        return e == cudaSuccess ? 0 : -(int)e;
    }
}

int merge_topk_launch(const float* scores, const long long* ids, long long stride_s, long long stride_i, int n_lists, int nq, int k_in,
                      int k_out, float* out_scores, long long* out_ids, cudaStream_t stream) {
    if (nq <= 0) return 0;
    if (k_in < 1 || k_out < 1 || k_in > MERGE_CAP_MAX) return -2;
    if (k_out > VQ_KMAX) {
        long long n = (long long)n_lists * k_in;
        if (n > MERGE_CAP_MAX) return -2;
        int S = 2;
        while (S < n) S <<= 1;
        static bool big_attr[64] = {};
        int dev_b = 0;
        cudaGetDevice(&dev_b);
        if (!big_attr[dev_b & 63]) {
            cudaError_t e = cudaFuncSetAttribute(merge_topk_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MERGE_CAP_MAX * 8);
            if (e != cudaSuccess) return -(int)e;
            big_attr[dev_b & 63] = true;
        }
        merge_topk_big_kernel<<<nq, MERGE_THREADS, (size_t)S * 8, stream>>>(scores, ids, stride_s, stride_i, n_lists, nq, k_in, k_out, out_scores,
                                                                           out_ids, S);
        cudaError_t e = cudaGetLastError();
        return e == cudaSuccess ? 0 : -(int)e;
    }
    const size_t smem = (size_t)SEL_WARPS * 128 * sizeof(uint64_t);  // 4 KB
    merge_topk_kernel<<<(nq + SEL_WARPS - 1) / SEL_WARPS, SEL_WARPS * 32, smem, stream>>>(scores, ids, stride_s, stride_i, n_lists, nq, k_in, k_out,
                                                                                        out_scores, out_ids);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -(int)e;
}

}  // namespace vq
