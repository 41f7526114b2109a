// K4 — k-way merges of sorted candidate lists, one CTA per query, bitonic sort in shared memory.
//
//  * merge_lists: folds the per-(corpus slab, query) candidate-key lists written by K1/K2/K3 into
//    the final [nq, k] (scores, ids) the reference returns from faiss_index.search
//    (datasets/search.py:384-385: float32 scores, int64 ids, descending, -1 padding).
//  * merge_topk: the cross-GPU step — every rank's local top-k [G, nq, k] (after the NCCL
//    all-gather) -> global top-k.  Top-k of a union == top-k of the per-shard top-k's.
// Both use the engine-wide tie rule: equal scores -> lower id first.  ids must be < 2^32.
#include "vq_common.cuh"
#include "vq_kernels.h"

namespace vq {

constexpr int MERGE_THREADS = 256;
constexpr int MERGE_CAP_MAX = 8192;  // 64 KB of keys

__device__ __forceinline__ void block_bitonic_sort_desc(uint64_t* keys, int S) {
    for (int s = 2; s <= S; s <<= 1) {
        for (int j = s >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < (S >> 1); i += MERGE_THREADS) {
                const int lo = ((i & ~(j - 1)) << 1) | (i & (j - 1));
                const int hi = lo | j;
                const bool desc = (lo & s) == 0;
                const uint64_t a = keys[lo], b = keys[hi];
                if (desc ? (a < b) : (a > b)) {
                    keys[lo] = b;
                    keys[hi] = a;
                }
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ int next_pow2(int x) {
    int p = 2;
    while (p < x) p <<= 1;
    return p;
}

// Gathers the (unsorted) candidate lists of one query — the running merged list in set 0 (if
// use_set0) plus sets [set_begin, set_begin + n_sets) — sorts them in shared memory, and writes the
// best k back to set 0 (sorted raw pairs + count), the k-th score to thr_out (the primed threshold
// of the next pass), and, for the last pass, the final (scores, ids).  One CTA per query.
constexpr int MERGE_MAX_SETS = 1024;

__global__ void __launch_bounds__(MERGE_THREADS)
merge_lists_kernel(uint64_t* __restrict__ lists, int* __restrict__ counts, int use_set0, int set_begin, int n_sets, int nq_pad, int k,
                   long long id_offset, float* __restrict__ thr_out, float* __restrict__ out_scores, long long* __restrict__ out_ids,
                   int cap) {
    extern __shared__ uint64_t keys[];
    __shared__ int off_s[MERGE_MAX_SETS + 2];  // exclusive prefix sums of the list lengths (index 0 = set 0)
    __shared__ int round_end;
    const int q = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int L = n_sets + 1;  // logical list l: 0 = running merged list (set 0), l >= 1 = set_begin + l - 1
    // all list lengths at once (the loads are independent), then a serial scan in shared memory
    for (int l = threadIdx.x; l < L; l += MERGE_THREADS) {
        const int set = l == 0 ? 0 : set_begin + l - 1;
        off_s[l + 1] = (l == 0 && !use_set0) ? 0 : counts[(size_t)set * nq_pad + q];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        off_s[0] = 0;
        for (int l = 1; l <= L; l++) { run += off_s[l]; off_s[l] = run; }
    }
    __syncthreads();
    int have = 0, l0 = 0;
    do {
        // this round takes lists [l0, l1): as many as fit beside the `have` keys kept from earlier rounds
        if (threadIdx.x == 0) {
            int l1 = l0;
            while (l1 < L && have + (off_s[l1 + 1] - off_s[l0]) <= cap) l1++;
            round_end = l1;
        }
        __syncthreads();
        const int l1 = round_end;
        const int base = off_s[l0];
        for (int l = l0 + warp; l < l1; l += MERGE_THREADS / 32) {
            const int set = l == 0 ? 0 : set_begin + l - 1;
            const int n = off_s[l + 1] - off_s[l];
            const unsigned long long* src = reinterpret_cast<const unsigned long long*>(lists + ((size_t)set * nq_pad + q) * VQ_CAP);
            for (int i = lane; i < n; i += 32) keys[have + off_s[l] - base + i] = raw_to_key(__ldcg(src + i));
        }
        const int fill = have + off_s[l1] - base;
        const int S = next_pow2(fill);
        for (int i = fill + threadIdx.x; i < S; i += MERGE_THREADS) keys[i] = 0ull;
        __syncthreads();
        block_bitonic_sort_desc(keys, S);
        have = fill < k ? fill : k;
        l0 = l1;
    } while (l0 < L);
    uint64_t* dst = lists + (size_t)q * VQ_CAP;  // set 0
    for (int i = threadIdx.x; i < have; i += MERGE_THREADS) dst[i] = key_to_raw(keys[i]);
    if (threadIdx.x == 0) {
        counts[q] = have;
        if (thr_out) thr_out[q] = have >= k ? key_score(keys[k - 1]) : -FLT_MAX;
    }
    if (out_scores) {
        for (int i = threadIdx.x; i < k; i += MERGE_THREADS) {
            if (i < have) {
                out_scores[(size_t)q * k + i] = key_score(keys[i]);
                out_ids[(size_t)q * k + i] = (long long)key_row(keys[i]) + id_offset;
            } else {
                out_scores[(size_t)q * k + i] = -FLT_MAX;
                out_ids[(size_t)q * k + i] = -1;
            }
        }
    }
}

__global__ void __launch_bounds__(MERGE_THREADS)
merge_topk_kernel(const float* __restrict__ scores, const long long* __restrict__ ids, int n_lists, int nq, int k_in, int k_out,
                  float* __restrict__ out_scores, long long* __restrict__ out_ids, int cap) {
    extern __shared__ uint64_t keys[];
    const int q = blockIdx.x;
    int have = 0, c = 0;
    do {
        int fill = have;
        while (c < n_lists && fill + k_in <= cap) {
            const size_t base = ((size_t)c * nq + q) * k_in;
            for (int i = threadIdx.x; i < k_in; i += MERGE_THREADS) {
                const long long id = ids[base + i];
                keys[fill + i] = id >= 0 ? make_key(scores[base + i], (uint32_t)id) : 0ull;  // padding sorts last
            }
            fill += k_in;
            c++;
        }
        const int S = next_pow2(fill);
        for (int i = fill + threadIdx.x; i < S; i += MERGE_THREADS) keys[i] = 0ull;
        __syncthreads();
        block_bitonic_sort_desc(keys, S);
        have = fill < k_out ? fill : k_out;
    } while (c < n_lists);
    for (int i = threadIdx.x; i < k_out; i += MERGE_THREADS) {
        const uint64_t key = i < have ? keys[i] : 0ull;
        if (key != 0ull) {
            out_scores[(size_t)q * k_out + i] = key_score(key);
            out_ids[(size_t)q * k_out + i] = (long long)key_row(key);
        } else {
            out_scores[(size_t)q * k_out + i] = -FLT_MAX;
            out_ids[(size_t)q * k_out + i] = -1;
        }
    }
}

static int pick_cap(long long total, int k) {
    long long need = total > 2LL * k ? total : 2LL * k;
    int cap = 256;
    while (cap < need && cap < MERGE_CAP_MAX) cap <<= 1;
    return cap;
}

int merge_lists_launch(uint64_t* lists, int* counts, int use_set0, int set_begin, int n_sets, int nq_pad, int nq, int k, long long id_offset,
                       float* thr_out, float* out_scores, long long* out_ids, int expected_total, cudaStream_t stream) {
    if (nq <= 0) return 0;
    if (n_sets + 1 > MERGE_MAX_SETS) return -2;
    // capacity: at least one full list + the running top-k must fit
    const int cap = pick_cap(expected_total, VQ_CAP);
    const size_t smem = (size_t)cap * sizeof(uint64_t);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(merge_lists_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MERGE_CAP_MAX * 8);
        if (e != cudaSuccess) return -(int)e;
        attr_set = true;
    }
    merge_lists_kernel<<<nq, MERGE_THREADS, smem, stream>>>(lists, counts, use_set0, set_begin, n_sets, nq_pad, k, id_offset, thr_out, out_scores,
                                                           out_ids, cap);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -(int)e;
}

int merge_topk_launch(const float* scores, const long long* ids, int n_lists, int nq, int k_in, int k_out,
                      float* out_scores, long long* out_ids, cudaStream_t stream) {
    if (nq <= 0) return 0;
    if (k_in < 1 || k_out < 1 || k_in + k_out > MERGE_CAP_MAX) return -2;
    const int cap = pick_cap((long long)n_lists * k_in, k_in > k_out ? k_in : k_out);
    const size_t smem = (size_t)cap * sizeof(uint64_t);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(merge_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MERGE_CAP_MAX * 8);
        if (e != cudaSuccess) return -(int)e;
        attr_set = true;
    }
    merge_topk_kernel<<<nq, MERGE_THREADS, smem, stream>>>(scores, ids, n_lists, nq, k_in, k_out, out_scores, out_ids, cap);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -(int)e;
}

}  // namespace vq
