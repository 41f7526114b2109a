// K6 — run post-processing on the device: what Searcher.__call__ does per (query, hit) in a
// triple Python loop (meerqat/ir/search.py:413-440):
//
//     for score, i in zip(scores, indices):            # hits of one query, best first
//         for j in index_mapping[i]:                   # article -> passages (one2many) / row -> doc (many2one)
//             many2one is None : run[j] = score - penalty; penalty += 1e-8     # overwrite, keeps dict position
//             many2one == max  : run[j] = max(run.get(j, -inf), score)
//         if len(run) >= k: break
//
// The run is a Python dict: documents keep INSERTION order and duplicates update in place; that order is part
// of the contract (ties are later sorted stably).  The penalty is carried exactly like the reference does under
// NumPy >= 2 promotion rules: it accumulates in double (a Python float), restarts at 0 for every hit, is
// converted to float32 and subtracted in float32 (`np.float32 - float` stays float32).  At DPR score magnitudes
// (~70, ulp 7.6e-6) the subtraction is a no-op; for cosine scores (`L2norm,Flat` indexes, |s| <= 1, ulp <= 6e-8)
// it separates the passages of one article, as in the reference.
//
// One warp per query; the growing run lives in shared memory; membership tests are warp-parallel.
// The mapping is CSR on the device: offsets int64 [n_rows + 1], targets int64 [nnz].
#include "vq_common.cuh"
#include "vq_kernels.h"

namespace vq {

constexpr int POST_WARPS = 4;

__global__ void __launch_bounds__(POST_WARPS * 32)
expand_hits_kernel(const float* __restrict__ scores, const long long* __restrict__ ids, int nq, int k_in,
                   const long long* __restrict__ map_offsets, const long long* __restrict__ map_targets, int many2one_max, int k,
                   int out_cap, double* __restrict__ out_scores, long long* __restrict__ out_ids, int* __restrict__ out_counts,
                   int* __restrict__ overflow) {
    extern __shared__ unsigned char post_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.x * POST_WARPS + warp;
    if (q >= nq) return;
    long long* docs = reinterpret_cast<long long*>(post_smem) + (size_t)warp * out_cap;
    float* vals = reinterpret_cast<float*>(reinterpret_cast<long long*>(post_smem) + (size_t)POST_WARPS * out_cap) + (size_t)warp * out_cap;
    int n = 0;
    bool lost = false;
    for (int h = 0; h < k_in; h++) {
        const long long row = ids[(size_t)q * k_in + h];
        if (row < 0) break;  // padding: no more hits
        const float s = scores[(size_t)q * k_in + h];
        const long long t0 = map_offsets ? map_offsets[row] : row;
        const long long t1 = map_offsets ? map_offsets[row + 1] : row + 1;
        double penalty = 0.0;  // search.py:420: reset per hit; only the one2many mapping path uses it
        for (long long t = t0; t < t1; t++) {
            const long long doc = map_offsets ? map_targets[t] : row;
            // is doc already in the run?  (warp-parallel scan, lowest matching position wins)
            int found = -1;
            for (int base = 0; base < n; base += 32) {
                const int i = base + lane;
                const unsigned m = __ballot_sync(VQ_FULL_MASK, i < n && docs[i] == doc);
                if (m) { found = base + __ffs(m) - 1; break; }
            }
            // one2many through a mapping: score - penalty (float32 arithmetic); many2one / no mapping: the score
            const float v = (map_offsets && !many2one_max) ? __fsub_rn(s, (float)penalty) : s;
            penalty += 1e-8;
            if (lane == 0) {
                if (found >= 0) {
                    if (!many2one_max || vals[found] < v) vals[found] = v;
                } else if (n < out_cap) {
                    docs[n] = doc;
                    vals[n] = v;
                }
            }
            if (found < 0) {
                if (n < out_cap) n++;
                else lost = true;
            }
            __syncwarp();
        }
        if (n >= k) break;  // checked after each HIT, like the reference
    }
    for (int i = lane; i < out_cap; i += 32) {
        out_ids[(size_t)q * out_cap + i] = i < n ? docs[i] : -1;
        out_scores[(size_t)q * out_cap + i] = i < n ? (double)vals[i] : 0.0;
    }
    if (lane == 0) {
        out_counts[q] = n;
        if (lost) atomicExch(overflow, 1);
    }
}

int expand_hits_launch(const float* scores, const long long* ids, int nq, int k_in, const long long* map_offsets,
                       const long long* map_targets, int many2one_max, int k, int out_cap, double* out_scores, long long* out_ids,
                       int* out_counts, int* overflow, cudaStream_t stream) {
    if (nq <= 0) return 0;
    const size_t smem = (size_t)POST_WARPS * out_cap * (sizeof(long long) + sizeof(float));
    if (smem > 200 * 1024) return -2;
    static size_t attr_set_dev[64] = {};  // cudaFuncSetAttribute is per device
    int dev_ = 0;
    cudaGetDevice(&dev_);
    size_t& attr_set = attr_set_dev[dev_ & 63];
    if (smem > attr_set && smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(expand_hits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return -(int)e;
        attr_set = smem;
    }
    expand_hits_kernel<<<(nq + POST_WARPS - 1) / POST_WARPS, POST_WARPS * 32, smem, stream>>>(
        scores, ids, nq, k_in, map_offsets, map_targets, many2one_max, k, out_cap, out_scores, out_ids, out_counts, overflow);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -(int)e;
}

}  // namespace vq
