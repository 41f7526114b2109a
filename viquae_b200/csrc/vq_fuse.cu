// K5 — late fusion of the per-index runs of one query batch (meerqat.ir.fuse), in float64 like the
// reference (numba typed dicts unicode -> float64, meerqat/ir/fuse.py:54-57).
//
//   default_minimum  fuse.py:132-149  per query, U = union of doc ids over runs; every run that has at
//                                     least one result for the query gets its missing docs filled with
//                                     that run's per-query minimum.  Happens BEFORE normalisation.
//   gzmuv            fuse.py:86-129   s' = (s - mu) / max(sigma, 1e-9) with mu / sigma (population)
//                                     over ALL scores of the run (all queries, filled entries included)
//   zmuv, min-max    ranx norms       same formula per (run, query) / (s - min) / max(max - min, 1e-9)
//   wsum             ranx.fuse        fused(d) = sum_r w_r * s'_r(d) over the runs that hold d
//   result                            union per query sorted by descending fused score (ties: lower id)
//
// Small kernels: A builds the union and the filled [U x R] matrix per query (sort by doc id in shared memory),
// B1/B2 reduce the global statistics per run (gzmuv only; deterministic: fixed slices, fixed-order trees, Chan's
// pairwise combination of (n, mean, M2)), C normalises, combines and sorts.  Latency/HBM-bound and tiny next to K3.
#include "vq_common.cuh"
#include "vq_kernels.h"

#include <math.h>

namespace vq {
namespace fuse {

constexpr int THREADS = 256;
constexpr int E_MAX = 16384;  // n_runs * k_in entries per query: keys (8 B) + ids (4 B) per entry in dynamic shared memory
constexpr int R_MAX = 16;

__device__ __forceinline__ void sort_desc_u64(uint64_t* keys, int S) {
    for (int s = 2; s <= S; s <<= 1)
        for (int j = s >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < (S >> 1); i += THREADS) {
                const int lo = ((i & ~(j - 1)) << 1) | (i & (j - 1)), hi = lo | j;
                const bool desc = (lo & s) == 0;
                const uint64_t a = keys[lo], b = keys[hi];
                if (desc ? (a < b) : (a > b)) { keys[lo] = b; keys[hi] = a; }
            }
            __syncthreads();
        }
}

// "a ranks before b": larger key, then smaller id
__device__ __forceinline__ bool before(uint64_t ka, uint32_t ia, uint64_t kb, uint32_t ib) {
    return ka > kb || (ka == kb && ia < ib);
}
__device__ __forceinline__ void sort_pairs_desc(uint64_t* keys, uint32_t* vals, int S) {
    for (int s = 2; s <= S; s <<= 1)
        for (int j = s >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < (S >> 1); i += THREADS) {
                const int lo = ((i & ~(j - 1)) << 1) | (i & (j - 1)), hi = lo | j;
                const bool desc = (lo & s) == 0;
                const uint64_t a = keys[lo], b = keys[hi];
                const uint32_t va = vals[lo], vb = vals[hi];
                const bool sw = desc ? before(b, vb, a, va) : before(a, va, b, vb);
                if (sw) { keys[lo] = b; keys[hi] = a; vals[lo] = vb; vals[hi] = va; }
            }
            __syncthreads();
        }
}

__device__ __forceinline__ int pow2_ge(int x) { int p = 2; while (p < x) p <<= 1; return p; }

__device__ __forceinline__ uint64_t d2ord(double v) {
    unsigned long long u = (unsigned long long)__double_as_longlong(v + 0.0);
    return u ^ ((u >> 63) ? ~0ull : (1ull << 63));
}

// fixed-order block reductions (result broadcast to all threads)
__device__ __forceinline__ double block_sum(double v, double* red) {
    red[threadIdx.x] = v;
    __syncthreads();
    for (int o = THREADS / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    const double r = red[0];
    __syncthreads();
    return r;
}
__device__ __forceinline__ double block_min(double v, double* red) {
    red[threadIdx.x] = v;
    __syncthreads();
    for (int o = THREADS / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] = fmin(red[threadIdx.x], red[threadIdx.x + o]);
        __syncthreads();
    }
    const double r = red[0];
    __syncthreads();
    return r;
}
__device__ __forceinline__ double block_max(double v, double* red) {
    red[threadIdx.x] = v;
    __syncthreads();
    for (int o = THREADS / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + o]);
        __syncthreads();
    }
    const double r = red[0];
    __syncthreads();
    return r;
}

// ---- A: union of doc ids + filled matrix.  mat[q][u][r] (NaN = run r does not hold doc u), doc[q][u], ucount[q]
__global__ void __launch_bounds__(THREADS)
union_kernel(const double* __restrict__ scores, const long long* __restrict__ ids, int R, int nq, int k_in, int defmin,
             double* __restrict__ mat, long long* __restrict__ doc, int* __restrict__ ucount) {
    extern __shared__ uint64_t keys[];  // [pow2_ge(R * k_in)]
    __shared__ double red[THREADS];
    __shared__ int scan[THREADS];
    __shared__ double run_min[R_MAX];
    __shared__ int run_n[R_MAX];
    const int q = blockIdx.x, tid = threadIdx.x;
    const int E = R * k_in, S = pow2_ge(E);
    // per-run minimum and count over the entries present for this query
    for (int r = 0; r < R; r++) {
        double mn = INFINITY;
        double n = 0.0;
        for (int i = tid; i < k_in; i += THREADS) {
            const size_t at = ((size_t)r * nq + q) * k_in + i;
            if (ids[at] >= 0) { mn = fmin(mn, scores[at]); n += 1.0; }
        }
        mn = block_min(mn, red);
        n = block_sum(n, red);
        if (tid == 0) { run_min[r] = mn; run_n[r] = (int)n; }
    }
    // sort present entries by doc id: key = ~(id << 32 | run << 16 | slot); absent -> 0 (sorts last)
    for (int e = tid; e < S; e += THREADS) {
        uint64_t key = 0ull;
        if (e < E) {
            const int r = e / k_in, slot = e - r * k_in;
            const long long id = ids[((size_t)r * nq + q) * k_in + slot];
            if (id >= 0) key = ~(((uint64_t)(uint32_t)id << 32) | ((uint64_t)r << 16) | (uint64_t)slot);
        }
        keys[e] = key;
    }
    __syncthreads();
    sort_desc_u64(keys, S);
    // head flags -> exclusive scan -> union index of every entry (each thread owns a contiguous span)
    const int span = (S + THREADS - 1) / THREADS;
    const int b0 = tid * span, b1 = min(S, b0 + span);
    int local = 0;
    for (int i = b0; i < b1; i++) {
        const uint64_t key = keys[i];
        if (key == 0ull) break;
        const uint32_t id = (uint32_t)((~key) >> 32);
        const bool head = (i == 0) || ((uint32_t)((~keys[i - 1]) >> 32) != id);
        local += head ? 1 : 0;
    }
    scan[tid] = local;
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int i = 0; i < THREADS; i++) { const int v = scan[i]; scan[i] = run; run += v; }
        ucount[q] = run;
    }
    __syncthreads();
    const int U = ucount[q];
    const size_t mbase = (size_t)q * E * R;
    for (int i = tid; i < U * R; i += THREADS) {
        const int r = i % R;
        mat[mbase + i] = (defmin && run_n[r] > 0) ? run_min[r] : nan("");
    }
    __syncthreads();
    int u = scan[tid] - 1;
    for (int i = b0; i < b1; i++) {
        const uint64_t key = keys[i];
        if (key == 0ull) break;
        const uint64_t k2 = ~key;
        const uint32_t id = (uint32_t)(k2 >> 32);
        const int r = (int)((k2 >> 16) & 0xFFFF), slot = (int)(k2 & 0xFFFF);
        const bool head = (i == 0) || ((uint32_t)((~keys[i - 1]) >> 32) != id);
        if (head) { u++; doc[(size_t)q * E + u] = (long long)id; }
        mat[mbase + (size_t)u * R + r] = scores[((size_t)r * nq + q) * k_in + slot];
    }
}

// ---- B: global mean / population std per run over all queries (gzmuv), two deterministic stages:
// B1: block (b, r) reduces its fixed slice of queries to (n, mean, M2 = sum (v - mean)^2) with two passes over its
//     data (fixed-order tree); B2: one thread per run folds the slices in index order with Chan's pairwise update.
constexpr int GSTATS_BLOCKS = 96;

__global__ void __launch_bounds__(THREADS)
gstats_partial_kernel(const double* __restrict__ mat, const int* __restrict__ ucount, int R, int nq, int E, double* __restrict__ partial) {
    __shared__ double red[THREADS];
    const int b = blockIdx.x, r = blockIdx.y;
    const int per = (nq + gridDim.x - 1) / gridDim.x;
    const int qa = b * per, qb = min(nq, qa + per);
    double s = 0.0, n = 0.0;
    for (int q = qa; q < qb; q++) {
        const int U = ucount[q];
        const double* m = mat + (size_t)q * E * R;
        for (int u = threadIdx.x; u < U; u += THREADS) {
            const double v = m[(size_t)u * R + r];
            if (!isnan(v)) { s += v; n += 1.0; }
        }
    }
    s = block_sum(s, red);
    n = block_sum(n, red);
    const double mean = n > 0.0 ? s / n : 0.0;
    double ss = 0.0;
    for (int q = qa; q < qb; q++) {
        const int U = ucount[q];
        const double* m = mat + (size_t)q * E * R;
        for (int u = threadIdx.x; u < U; u += THREADS) {
            const double v = m[(size_t)u * R + r];
            if (!isnan(v)) ss += (v - mean) * (v - mean);
        }
    }
    ss = block_sum(ss, red);
    if (threadIdx.x == 0) {
        double* out = partial + ((size_t)r * gridDim.x + b) * 3;
        out[0] = n;
        out[1] = mean;
        out[2] = ss;
    }
}

__global__ void gstats_final_kernel(const double* __restrict__ partial, int n_blocks, double* __restrict__ gstats) {
    const int r = blockIdx.x;
    if (threadIdx.x != 0) return;
    double n = 0.0, mean = 0.0, m2 = 0.0;
    for (int b = 0; b < n_blocks; b++) {
        const double* in = partial + ((size_t)r * n_blocks + b) * 3;
        const double nb = in[0];
        if (nb <= 0.0) continue;
        const double delta = in[1] - mean;
        const double nt = n + nb;
        mean += delta * nb / nt;
        m2 += in[2] + delta * delta * n * nb / nt;
        n = nt;
    }
    gstats[2 * r] = mean;
    gstats[2 * r + 1] = n > 0.0 ? sqrt(m2 / n) : 0.0;
}

// per-run normalisation constants of one query: s' = (s - sub[r]) / div[r].  Called by all threads of the CTA.
__device__ __forceinline__ void norm_stats(const double* m, int U, int R, int norm, const double* gstats, double* red, double* sub, double* div) {
    const int tid = threadIdx.x;
    for (int r = 0; r < R; r++) {
        double a = 0.0, b = 1.0;
        if (norm == 1) {  // zmuv over this query's entries of run r
            double s = 0.0, n = 0.0;
            for (int u = tid; u < U; u += THREADS) { const double v = m[(size_t)u * R + r]; if (!isnan(v)) { s += v; n += 1.0; } }
            s = block_sum(s, red);
            n = block_sum(n, red);
            const double mean = n > 0.0 ? s / n : 0.0;
            double ss = 0.0;
            for (int u = tid; u < U; u += THREADS) { const double v = m[(size_t)u * R + r]; if (!isnan(v)) ss += (v - mean) * (v - mean); }
            ss = block_sum(ss, red);
            a = mean;
            b = fmax(n > 0.0 ? sqrt(ss / n) : 0.0, 1e-9);
        } else if (norm == 2) {
            a = gstats[2 * r];
            b = fmax(gstats[2 * r + 1], 1e-9);
        } else if (norm == 3) {
            double mn = INFINITY, mx = -INFINITY;
            for (int u = tid; u < U; u += THREADS) { const double v = m[(size_t)u * R + r]; if (!isnan(v)) { mn = fmin(mn, v); mx = fmax(mx, v); } }
            mn = block_min(mn, red);
            mx = block_max(mx, red);
            a = mn;
            b = fmax(mx - mn, 1e-9);
        }
        if (tid == 0) { sub[r] = a; div[r] = b; }
    }
    __syncthreads();
}

// ---- C: normalise + weighted sum + sort
__global__ void __launch_bounds__(THREADS)
combine_kernel(const double* __restrict__ mat, const long long* __restrict__ doc, const int* __restrict__ ucount,
               const double* __restrict__ weights, const double* __restrict__ gstats, int R, int nq, int E, int norm,
               double* __restrict__ out_scores, long long* __restrict__ out_ids, int* __restrict__ out_counts) {
    extern __shared__ uint64_t keys[];  // [S_max] fused-score keys, then [S_max] doc ids (uint32)
    __shared__ double red[THREADS];
    __shared__ double sub[R_MAX], div[R_MAX];
    const int q = blockIdx.x, tid = threadIdx.x;
    const int U = ucount[q];
    uint32_t* vals = reinterpret_cast<uint32_t*>(keys + pow2_ge(E));
    const double* m = mat + (size_t)q * E * R;
    norm_stats(m, U, R, norm, gstats, red, sub, div);
    const int S = pow2_ge(U);
    for (int u = tid; u < S; u += THREADS) {
        uint64_t key = 0ull;
        uint32_t val = 0xFFFFFFFFu;
        if (u < U) {
            double f = 0.0;
            for (int r = 0; r < R; r++) {
                const double v = m[(size_t)u * R + r];
                if (!isnan(v)) f += weights[r] * ((v - sub[r]) / div[r]);
            }
            key = d2ord(f);
            val = (uint32_t)doc[(size_t)q * E + u];
        }
        keys[u] = key;
        vals[u] = val;
    }
    __syncthreads();
    sort_pairs_desc(keys, vals, S);
    for (int i = tid; i < E; i += THREADS) {
        if (i < U) {
            const uint64_t o = keys[i];
            const unsigned long long bits = (o >> 63) ? (o ^ (1ull << 63)) : ~o;
            out_scores[(size_t)q * E + i] = __longlong_as_double((long long)bits);
            out_ids[(size_t)q * E + i] = (long long)vals[i];
        } else {
            out_scores[(size_t)q * E + i] = 0.0;
            out_ids[(size_t)q * E + i] = -1;
        }
    }
    if (tid == 0) out_counts[q] = U;
}


// ---- D: grid search over weight vectors (Fusion.fit, meerqat/ir/fuse.py:193-217 -> ranx.optimize_fusion).
// D1 normalises the filled matrix once per query (the statistics do not depend on the weights) and locates the
// query's relevant documents in its union.  D2 evaluates every weight vector: the metrics the reference optimises
// (mrr@k, precision@k, hit_rate@k, recall@k) only need the RANKS of the relevant documents, and
//     rank(d) = 1 + #{u : fused(u) > fused(d)  or  (fused(u) == fused(d) and id(u) < id(d))}
// is a counting pass over the union — no sort per weight vector.  D3 averages over the judged queries in a fixed
// order and picks the first best weight vector.
constexpr int REL_MAX = 256;  // relevant documents of one query that can appear in its union

__global__ void __launch_bounds__(THREADS)
grid_prepare_kernel(double* __restrict__ mat, const long long* __restrict__ doc, const int* __restrict__ ucount, const double* __restrict__ gstats,
                    int R, int nq, int E, int norm, const long long* __restrict__ rel, int max_rel, int* __restrict__ rel_pos,
                    int* __restrict__ rel_cnt) {
    __shared__ double red[THREADS];
    __shared__ double sub[R_MAX], div[R_MAX];
    __shared__ int n_found;
    const int q = blockIdx.x, tid = threadIdx.x;
    const int U = ucount[q];
    double* m = mat + (size_t)q * E * R;
    norm_stats(m, U, R, norm, gstats, red, sub, div);
    for (int i = tid; i < U * R; i += THREADS) {
        const double v = m[i];
        if (!isnan(v)) m[i] = (v - sub[i % R]) / div[i % R];
    }
    // relevant documents present in the union: doc[q][0..U) ascends by id -> binary search
    if (tid == 0) n_found = 0;
    __syncthreads();
    const long long* dq = doc + (size_t)q * E;
    for (int j = tid; j < max_rel; j += THREADS) {
        const long long want = rel[(size_t)q * max_rel + j];
        if (want < 0) continue;
        int lo = 0, hi = U;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (dq[mid] < want) lo = mid + 1; else hi = mid;
        }
        if (lo < U && dq[lo] == want) {
            const int at = atomicAdd(&n_found, 1);
            if (at < REL_MAX) rel_pos[(size_t)q * REL_MAX + at] = lo;
        }
    }
    __syncthreads();
    if (tid == 0) rel_cnt[q] = n_found < REL_MAX ? n_found : REL_MAX;
}

// grid (nq, ceil(W / warps)): one warp per weight vector
__global__ void __launch_bounds__(THREADS)
grid_eval_kernel(const double* __restrict__ matn, const long long* __restrict__ doc, const int* __restrict__ ucount,
                 const double* __restrict__ weights, int W, int R, int nq, int E, const int* __restrict__ rel_pos, const int* __restrict__ rel_cnt,
                 const int* __restrict__ n_rel_total, int metric, int cut, double* __restrict__ per_query) {
    const int q = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int w = blockIdx.y * (THREADS / 32) + warp;
    if (w >= W) return;
    const int U = ucount[q], nrel = rel_cnt[q];
    const double* m = matn + (size_t)q * E * R;
    const long long* dq = doc + (size_t)q * E;
    double wt[R_MAX];
#pragma unroll
    for (int r = 0; r < R_MAX; r++) wt[r] = r < R ? weights[(size_t)w * R + r] : 0.0;
    const int kk = cut > 0 ? cut : 0x7fffffff;
    int hits = 0, best_rank = 0x7fffffff;
    for (int j0 = 0; j0 < nrel; j0 += 8) {  // eight relevant documents per sweep over the union
        double fd[8];
        long long idd[8];
        int cnt[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            cnt[j] = 0;
            fd[j] = 0.0;
            idd[j] = -1;
            if (j0 + j < nrel) {
                const int u = rel_pos[(size_t)q * REL_MAX + j0 + j];
                double f = 0.0;
#pragma unroll
                for (int r = 0; r < R_MAX; r++)
                    if (r < R) { const double v = m[(size_t)u * R + r]; if (!isnan(v)) f += wt[r] * v; }
                fd[j] = f;
                idd[j] = dq[u];
            }
        }
        for (int u = lane; u < U; u += 32) {
            double f = 0.0;
#pragma unroll
            for (int r = 0; r < R_MAX; r++)
                if (r < R) { const double v = m[(size_t)u * R + r]; if (!isnan(v)) f += wt[r] * v; }
            const long long id = dq[u];
#pragma unroll
            for (int j = 0; j < 8; j++) cnt[j] += (f > fd[j] || (f == fd[j] && id < idd[j])) ? 1 : 0;
        }
#pragma unroll
        for (int j = 0; j < 8; j++) {
            int c = cnt[j];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(VQ_FULL_MASK, c, o);
            if (j0 + j < nrel) {
                const int rank = c + 1;
                if (rank <= kk) hits++;
                best_rank = rank < best_rank ? rank : best_rank;
            }
        }
    }
    if (lane == 0) {
        double v = 0.0;
        if (metric == 0) v = best_rank <= kk ? 1.0 / (double)best_rank : 0.0;                       // mrr@k
        else if (metric == 1) v = (double)hits / (double)(cut > 0 ? cut : (U > 0 ? U : 1));        // precision@k
        else if (metric == 2) v = hits > 0 ? 1.0 : 0.0;                                             // hit_rate@k
        else v = n_rel_total[q] > 0 ? (double)hits / (double)n_rel_total[q] : 0.0;                  // recall@k
        per_query[(size_t)w * nq + q] = v;
    }
}

// one CTA per weight vector: fixed-order sum over the queries / n_judged; the last CTA to finish picks the first best
__global__ void __launch_bounds__(THREADS)
grid_reduce_kernel(const double* __restrict__ per_query, int W, int nq, int n_judged, double* __restrict__ out_metric, int* __restrict__ out_best,
                   unsigned* __restrict__ done) {
    __shared__ double red[THREADS];
    __shared__ bool last;
    const int w = blockIdx.x;
    double s = 0.0;
    for (int q = threadIdx.x; q < nq; q += THREADS) s += per_query[(size_t)w * nq + q];
    s = block_sum(s, red);
    if (threadIdx.x == 0) {
        out_metric[w] = s / (double)(n_judged > 0 ? n_judged : 1);
        __threadfence();
        last = atomicAdd(done, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        int best = 0;
        double best_v = -1.0;
        for (int i = 0; i < W; i++) {
            const double v = reinterpret_cast<volatile double*>(out_metric)[i];
            if (v > best_v + 1e-15) { best_v = v; best = i; }  // first best wins, like ranx.optimize_fusion
        }
        *out_best = best;
        *done = 0;
    }
}

}  // namespace fuse

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static int pow2_ge_host(int x) { int p = 2; while (p < x) p <<= 1; return p; }

size_t fuse_workspace_bytes(int n_runs, int nq, int k_in) {
    const size_t E = (size_t)n_runs * k_in;
    return align256((size_t)nq * E * n_runs * sizeof(double)) + align256((size_t)nq * E * sizeof(long long)) +
           align256((size_t)nq * sizeof(int)) + align256((size_t)2 * n_runs * sizeof(double)) +
           align256((size_t)3 * n_runs * fuse::GSTATS_BLOCKS * sizeof(double));
}

struct FuseWs {
    double* mat;
    long long* doc;
    int* ucount;
    double* gstats;
    double* partial;
    uint8_t* end;
};

static FuseWs carve(void* workspace, int R, int nq, int E) {
    uint8_t* w = reinterpret_cast<uint8_t*>(workspace);
    FuseWs f;
    f.mat = reinterpret_cast<double*>(w);
    w += align256((size_t)nq * E * R * sizeof(double));
    f.doc = reinterpret_cast<long long*>(w);
    w += align256((size_t)nq * E * sizeof(long long));
    f.ucount = reinterpret_cast<int*>(w);
    w += align256((size_t)nq * sizeof(int));
    f.gstats = reinterpret_cast<double*>(w);
    w += align256((size_t)2 * R * sizeof(double));
    f.partial = reinterpret_cast<double*>(w);
    w += align256((size_t)3 * R * fuse::GSTATS_BLOCKS * sizeof(double));
    f.end = w;
    return f;
}

// union + default-minimum fill (+ global statistics for gzmuv): the part of a fusion that does not depend on the weights
static int fuse_front(const double* scores, const long long* ids, const FuseParams& fp, const FuseWs& f, cudaStream_t stream) {
    const int R = fp.n_runs, nq = fp.nq, k_in = fp.k_in, E = R * k_in;
    const size_t smem_u = (size_t)pow2_ge_host(E) * sizeof(uint64_t);
    static size_t attr_dev[64] = {};
    int dev_ = 0;
    cudaGetDevice(&dev_);
    if (smem_u > 40 * 1024 && smem_u > attr_dev[dev_ & 63]) {  // static shared memory counts towards the 48 KB default limit
        cudaError_t e = cudaFuncSetAttribute(fuse::union_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_u);
        if (e != cudaSuccess) return -(int)e;
        attr_dev[dev_ & 63] = smem_u;
    }
    fuse::union_kernel<<<nq, fuse::THREADS, smem_u, stream>>>(scores, ids, R, nq, k_in, fp.defmin, f.mat, f.doc, f.ucount);
    if (fp.norm == 2) {
        const int nb = nq < fuse::GSTATS_BLOCKS ? nq : fuse::GSTATS_BLOCKS;
        fuse::gstats_partial_kernel<<<dim3(nb, R), fuse::THREADS, 0, stream>>>(f.mat, f.ucount, R, nq, E, f.partial);
        fuse::gstats_final_kernel<<<R, 32, 0, stream>>>(f.partial, nb, f.gstats);
    }
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -(int)e;
}

static bool fuse_shape_ok(const FuseParams& fp) {
    const long long E = (long long)fp.n_runs * fp.k_in;
    return fp.n_runs >= 1 && fp.n_runs <= fuse::R_MAX && fp.k_in >= 1 && fp.k_in <= 65535 && E <= fuse::E_MAX && fp.norm >= 0 && fp.norm <= 3;
}

int fuse_launch(const double* scores, const long long* ids, const double* weights, const FuseParams& fp,
                double* out_scores, long long* out_ids, int* out_counts, void* workspace, size_t ws_bytes, cudaStream_t stream) {
    const int R = fp.n_runs, nq = fp.nq, k_in = fp.k_in;
    const int E = R * k_in;
    if (!fuse_shape_ok(fp)) return -3;
    if (nq <= 0) return 0;
    if (ws_bytes < fuse_workspace_bytes(R, nq, k_in)) return -4;
    const FuseWs f = carve(workspace, R, nq, E);
    int rc = fuse_front(scores, ids, fp, f, stream);
    if (rc) return rc;
    const size_t smem_c = (size_t)pow2_ge_host(E) * (sizeof(uint64_t) + sizeof(uint32_t));
    static size_t attr_dev[64] = {};
    int dev_ = 0;
    cudaGetDevice(&dev_);
    if (smem_c > 40 * 1024 && smem_c > attr_dev[dev_ & 63]) {
        cudaError_t e = cudaFuncSetAttribute(fuse::combine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c);
        if (e != cudaSuccess) return -(int)e;
        attr_dev[dev_ & 63] = smem_c;
    }
    fuse::combine_kernel<<<nq, fuse::THREADS, smem_c, stream>>>(f.mat, f.doc, f.ucount, weights, f.gstats, R, nq, E, fp.norm, out_scores, out_ids, out_counts);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -(int)e;
}

size_t fuse_grid_workspace_bytes(int n_runs, int nq, int k_in, int n_weights) {
    return fuse_workspace_bytes(n_runs, nq, k_in) + align256((size_t)nq * fuse::REL_MAX * sizeof(int)) + align256((size_t)nq * sizeof(int)) +
           align256((size_t)n_weights * nq * sizeof(double)) + 256;
}

int fuse_grid_launch(const double* scores, const long long* ids, const FuseParams& fp, const double* weights, int n_weights, const long long* rel,
                     int max_rel, const int* n_rel_total, int n_judged, int metric, int cut, double* out_metric, int* out_best, void* workspace,
                     size_t ws_bytes, cudaStream_t stream) {
    const int R = fp.n_runs, nq = fp.nq, k_in = fp.k_in;
    const int E = R * k_in;
    if (!fuse_shape_ok(fp) || n_weights < 1 || max_rel < 1 || metric < 0 || metric > 3) return -3;
    if (nq <= 0) return 0;
    if (ws_bytes < fuse_grid_workspace_bytes(R, nq, k_in, n_weights)) return -4;
    const FuseWs f = carve(workspace, R, nq, E);
    uint8_t* w = f.end;
    int* rel_pos = reinterpret_cast<int*>(w);
    w += align256((size_t)nq * fuse::REL_MAX * sizeof(int));
    int* rel_cnt = reinterpret_cast<int*>(w);
    w += align256((size_t)nq * sizeof(int));
    double* per_query = reinterpret_cast<double*>(w);
    w += align256((size_t)n_weights * nq * sizeof(double));
    unsigned* done = reinterpret_cast<unsigned*>(w);
    cudaError_t e = cudaMemsetAsync(done, 0, sizeof(unsigned), stream);
    if (e != cudaSuccess) return -(int)e;
    int rc = fuse_front(scores, ids, fp, f, stream);
    if (rc) return rc;
    fuse::grid_prepare_kernel<<<nq, fuse::THREADS, 0, stream>>>(f.mat, f.doc, f.ucount, f.gstats, R, nq, E, fp.norm, rel, max_rel, rel_pos, rel_cnt);
    const int warps = fuse::THREADS / 32;
    fuse::grid_eval_kernel<<<dim3(nq, (n_weights + warps - 1) / warps), fuse::THREADS, 0, stream>>>(f.mat, f.doc, f.ucount, weights, n_weights, R, nq, E,
                                                                                                 rel_pos, rel_cnt, n_rel_total, metric, cut, per_query);
    fuse::grid_reduce_kernel<<<n_weights, fuse::THREADS, 0, stream>>>(per_query, n_weights, nq, n_judged, out_metric, out_best, done);
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -(int)e;
}

}  // namespace vq
