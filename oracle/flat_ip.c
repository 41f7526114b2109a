/* CPU oracle, C restatement of the exact inner-product top-k path.  TEST INFRASTRUCTURE ONLY:
 * built into oracle/_build/liboracle.so by oracle/Makefile; loaded only from tests/,
 * __graft_entry__.smoke() and bench.py's CPU-baseline legs.  Never linked into the product.
 *
 * PARITY UNPINNED for the contraction (see oracle/flat_ip.py header): FAISS is an un-vendored,
 * un-installable dependency of the reference (requirements.txt:14) and the reference holds no
 * golden vectors for this path.  Follows meerqat/ir/search.py:135-146 (float32 queries ->
 * dataset.search_batch) -> datasets/search.py:369-385 -> faiss IndexFlat(d, METRIC_INNER_PRODUCT)
 * .search: per query the k largest inner products, descending, int64 ids, (-FLT_MAX, -1) padding,
 * a candidate enters only if strictly greater than the current k-th (so NaN/-inf never enter).
 *
 * Reduction order is FIXED: acc = fmaf(q[j], x[j], acc) for j = 0..d-1 ascending, acc0 = 0.
 * The engine's fp32 "exact" kernel uses the same order, so scores compare bit-for-bit.
 * Tie rule (FAISS leaves it open): lower id first.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline float bf16_to_f32(uint16_t h) {
    uint32_t u = ((uint32_t)h) << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

static inline float f16_to_f32(uint16_t h) {
    uint32_t sign = ((uint32_t)h & 0x8000u) << 16, exp = (h >> 10) & 0x1f, man = h & 0x3ffu, u;
    if (exp == 0) {
        if (man == 0) u = sign;
        else { /* subnormal */
            int e = -1;
            do { man <<= 1; e++; } while (!(man & 0x400u));
            u = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ffu) << 13);
        }
    } else if (exp == 31) u = sign | 0x7f800000u | (man << 13);
    else u = sign | ((exp + 112) << 23) | (man << 13);
    float f;
    memcpy(&f, &u, 4);
    return f;
}

/* better(a,b): does candidate a rank before b?  higher score, then lower id */
static inline int better(float sa, int64_t ia, float sb, int64_t ib) {
    return sa > sb || (sa == sb && ia < ib);
}

/* sorted insertion into a descending list of current length *n (capacity k) */
static inline void insert(float* D, int64_t* I, int* n, int k, float s, int64_t id) {
    if (!(s > -FLT_MAX)) return;                       /* NaN, -inf, -FLT_MAX never selected */
    if (*n == k && !better(s, id, D[k - 1], I[k - 1])) return;
    int pos = (*n < k) ? *n : k - 1;
    while (pos > 0 && better(s, id, D[pos - 1], I[pos - 1])) {
        D[pos] = D[pos - 1];
        I[pos] = I[pos - 1];
        pos--;
    }
    D[pos] = s;
    I[pos] = id;
    if (*n < k) (*n)++;
}

/* dtype: 0 = f32, 1 = bf16 (uint16 storage), 2 = f16 (uint16 storage); queries use the same dtype.
 * Accumulation is always float32 fmaf in ascending j. */
int vq_oracle_flat_ip(const void* Q, const void* X, int dtype, int64_t nq, int64_t N, int d, int k,
                      int64_t id_offset, float* D, int64_t* I) {
    if (nq < 0 || N < 0 || d <= 0 || k <= 0) return -1;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t q = 0; q < nq; q++) {
        float* qf = (float*)malloc(sizeof(float) * (size_t)d);
        for (int j = 0; j < d; j++) {
            if (dtype == 0) qf[j] = ((const float*)Q)[q * d + j];
            else if (dtype == 1) qf[j] = bf16_to_f32(((const uint16_t*)Q)[q * d + j]);
            else qf[j] = f16_to_f32(((const uint16_t*)Q)[q * d + j]);
        }
        float* Dq = D + q * k;
        int64_t* Iq = I + q * k;
        int n = 0;
        for (int64_t r = 0; r < N; r++) {
            float acc = 0.0f;
            if (dtype == 0) {
                const float* x = (const float*)X + r * d;
                for (int j = 0; j < d; j++) acc = fmaf(qf[j], x[j], acc);
            } else if (dtype == 1) {
                const uint16_t* x = (const uint16_t*)X + r * d;
                for (int j = 0; j < d; j++) acc = fmaf(qf[j], bf16_to_f32(x[j]), acc);
            } else {
                const uint16_t* x = (const uint16_t*)X + r * d;
                for (int j = 0; j < d; j++) acc = fmaf(qf[j], f16_to_f32(x[j]), acc);
            }
            insert(Dq, Iq, &n, k, acc, r + id_offset);
        }
        for (int j = n; j < k; j++) { Dq[j] = -FLT_MAX; Iq[j] = -1; }
        free(qf);
    }
    return 0;
}

/* Row-wise L2 normalisation, meerqat/ir/search.py:43-46 (no epsilon: zero rows -> NaN).
 * norm = sqrtf(sum in ascending j of x*x accumulated with fmaf) — the engine's prep kernel uses
 * a different (tree) order, so compare with a tolerance, not bitwise. */
int vq_oracle_l2norm(const float* x, int64_t n, int d, float* out) {
    for (int64_t r = 0; r < n; r++) {
        float acc = 0.0f;
        for (int j = 0; j < d; j++) acc = fmaf(x[r * d + j], x[r * d + j], acc);
        float nrm = sqrtf(acc);
        for (int j = 0; j < d; j++) out[r * d + j] = x[r * d + j] / nrm;
    }
    return 0;
}
