// Shared device-side pieces of the exact inner-product top-k engine (sm_100a only).
//
// Candidate keys.  Every (score, row) pair travels as ONE 64-bit key
//     key = orderable(score) << 32 | ~row
// so that "larger key" == "ranks earlier": higher score first, then LOWER row id (the tie rule the
// oracle fixes; FAISS IndexFlat leaves it open).  Rows are shard-local uint32 (< 2^32 per GPU).
//
// Candidate lists.  Each (work item, query) owns a list of VQ_CAP (score, row) pairs in the caller-provided
// workspace (HBM, L2-resident in practice).  Kernels append every score that beats the running
// threshold; when a list could overflow, one warp sorts it (bitonic, registers + shuffles), keeps
// the best k and raises the threshold to the k-th score.  Nothing above the true k-th score of
// the slab is ever dropped, so the result is exact.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <float.h>
#include <stdint.h>

#define VQ_CAP 256   // keys per candidate list
#define VQ_KMAX 128  // largest k served by the fused list path (VQ_CAP - VQ_KMAX >= 128 pushes per tile)
#define VQ_FULL_MASK 0xffffffffu

namespace vq {

// Programmatic dependent launch (PTX griddepcontrol).  launch_dependents: kernels queued behind this one with the
// programmatic-stream-serialization attribute may start as soon as every CTA of this grid has executed it (or exited).
// wait: blocks until the kernels this grid depends on have completed and their writes are visible; a no-op for a
// kernel that was launched without the attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ uint32_t f2ord(float f) {
    uint32_t u = __float_as_uint(f);
    return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o) {
    uint32_t u = (o & 0x80000000u) ? (o ^ 0x80000000u) : ~o;
    return __uint_as_float(u);
}
// s + 0.0f folds -0.0 into +0.0 so equal scores have equal keys
__device__ __forceinline__ uint64_t make_key(float s, uint32_t row) {
    return ((uint64_t)f2ord(s + 0.0f) << 32) | (uint32_t)(~row);
}
__device__ __forceinline__ float key_score(uint64_t k) { return ord2f((uint32_t)(k >> 32)); }
__device__ __forceinline__ uint32_t key_row(uint64_t k) { return ~(uint32_t)k; }

// Candidate lists hold RAW pairs (row << 32 | float bits of the score): pushing costs one 8-byte store
// on the hot path; the order-preserving key is built when a list is pruned or merged.
__device__ __forceinline__ uint64_t pack_raw(float s, uint32_t row) { return ((uint64_t)row << 32) | (uint64_t)__float_as_uint(s); }
__device__ __forceinline__ uint64_t raw_to_key(uint64_t raw) { return make_key(__uint_as_float((uint32_t)raw), (uint32_t)(raw >> 32)); }
__device__ __forceinline__ uint64_t key_to_raw(uint64_t key) { return pack_raw(key_score(key), key_row(key)); }

__device__ __forceinline__ uint64_t umax64(uint64_t a, uint64_t b) { return a > b ? a : b; }
__device__ __forceinline__ uint64_t umin64(uint64_t a, uint64_t b) { return a < b ? a : b; }

// Bitonic sort, descending, of 32*E keys held E per lane; element e = r*32 + lane.
// The network is kept ROLLED (runtime stage size s and cross-lane stride j) so that the code stays a
// few hundred instructions and can be inlined wherever a list is pruned; only the in-register
// exchanges (stride >= 32, compile-time register distance) are unrolled.
template <int E, int RJ>
__device__ __forceinline__ void bitonic_inreg_stage(uint64_t (&key)[E], int s) {
#pragma unroll
    for (int r = 0; r < E; r++) {
        if ((r & RJ) == 0) {
            const bool desc = (((r * 32) & s) == 0);
            const uint64_t a = key[r], b = key[r | RJ];
            const bool sw = desc ? (a < b) : (a > b);
            key[r] = sw ? b : a;
            key[r | RJ] = sw ? a : b;
        }
    }
}

template <int E>
__device__ __forceinline__ void warp_bitonic_sort_desc(uint64_t (&key)[E], int lane) {
    static_assert(E == 8, "warp_bitonic_sort_desc is written for 256 keys per warp");
#pragma unroll 1
    for (int s = 2; s <= 32 * E; s <<= 1) {
        if (s >= 256) bitonic_inreg_stage<E, 4>(key, s);
        if (s >= 128) bitonic_inreg_stage<E, 2>(key, s);
        if (s >= 64) bitonic_inreg_stage<E, 1>(key, s);
#pragma unroll 1
        for (int j = (s >> 1) < 16 ? (s >> 1) : 16; j > 0; j >>= 1) {
            const bool lower = (lane & j) == 0;
#pragma unroll
            for (int r = 0; r < E; r++) {
                const int e = r * 32 + lane;
                const uint64_t other = __shfl_xor_sync(VQ_FULL_MASK, key[r], j);
                const bool desc = ((e & s) == 0);
                key[r] = (lower == desc) ? umax64(key[r], other) : umin64(key[r], other);
            }
        }
    }
}

// 128 keys per warp, 4 per lane (element e = r*32 + lane), descending
__device__ __forceinline__ void warp_bitonic_sort_desc4(uint64_t (&key)[4], int lane) {
#pragma unroll 1
    for (int s = 2; s <= 128; s <<= 1) {
        if (s >= 128) bitonic_inreg_stage<4, 2>(key, s);
        if (s >= 64) bitonic_inreg_stage<4, 1>(key, s);
#pragma unroll 1
        for (int j = (s >> 1) < 16 ? (s >> 1) : 16; j > 0; j >>= 1) {
            const bool lower = (lane & j) == 0;
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const int e = r * 32 + lane;
                const uint64_t other = __shfl_xor_sync(VQ_FULL_MASK, key[r], j);
                const bool desc = ((e & s) == 0);
                key[r] = (lower == desc) ? umax64(key[r], other) : umin64(key[r], other);
            }
        }
    }
}

// Top-128 of up to 32*R keys held R per lane (0 = empty slot; keys are distinct), by SELECTION instead of sorting:
// a bitwise search for the 128-th largest score word (one compare per key and bit, starting below the bits all keys
// share), a second search over the row words only if exactly tied scores straddle the cut, ballot compaction of the
// survivors into 128 shared-memory slots and ONE 128-key bitonic sort.  ~1/8 of the instructions of sorting every
// 256-key chunk (the merge kernels are ALU-pipe bound: 64-bit compare + select per exchange).
// out: sorted descending, 4 per lane (element e = r*32 + lane); slots: 128 uint64 of shared memory owned by the warp.
template <int R>
__device__ __forceinline__ void warp_select_top128(const uint64_t (&key)[R], uint64_t (&out)[4], uint64_t* slots, int lane) {
    int nv = 0;
    uint32_t mx = 0u, mn = 0xFFFFFFFFu;
#pragma unroll
    for (int r = 0; r < R; r++) {
        const uint32_t hi = (uint32_t)(key[r] >> 32);
        const bool valid = key[r] != 0ull;
        nv += valid ? 1 : 0;
        mx = max(mx, hi);
        mn = valid ? min(mn, hi) : mn;
    }
    nv = __reduce_add_sync(VQ_FULL_MASK, nv);
    uint32_t T = 0u, Lw = 0u;  // keep: hi > T, or hi == T and lo >= Lw
    if (nv > 128) {
        mx = __reduce_max_sync(VQ_FULL_MASK, mx);
        mn = __reduce_min_sync(VQ_FULL_MASK, mn);
        const uint32_t diff = mx ^ mn;
        uint32_t prefix = mx;
        if (diff) {
            const int top = 31 - __clz(diff);
            prefix = top == 31 ? 0u : (mx & ~((2u << top) - 1u));   // the bits every key shares
#pragma unroll 1
            for (int b = top; b >= 0; b--) {
                const uint32_t cand = prefix | (1u << b);
                int c = 0;
#pragma unroll
                for (int r = 0; r < R; r++) c += ((uint32_t)(key[r] >> 32) >= cand) ? 1 : 0;
                if (__reduce_add_sync(VQ_FULL_MASK, c) >= 128) prefix = cand;
            }
        }
        T = prefix;  // count(hi >= T) >= 128 > count(hi > T)
        int gt = 0, eq = 0;
#pragma unroll
        for (int r = 0; r < R; r++) {
            const uint32_t hi = (uint32_t)(key[r] >> 32);
            gt += hi > T ? 1 : 0;
            eq += hi == T ? 1 : 0;
        }
        gt = __reduce_add_sync(VQ_FULL_MASK, gt);
        eq = __reduce_add_sync(VQ_FULL_MASK, eq);
        const int need = 128 - gt;  // >= 1
        if (eq > need) {
            // exactly tied scores straddle the cut: the `need` largest row words (= lowest ids) among them
            uint32_t pl = 0u;
#pragma unroll 1
            for (int b = 31; b >= 0; b--) {
                const uint32_t cand = pl | (1u << b);
                int c = 0;
#pragma unroll
                for (int r = 0; r < R; r++) c += ((uint32_t)(key[r] >> 32) == T && (uint32_t)key[r] >= cand) ? 1 : 0;
                if (__reduce_add_sync(VQ_FULL_MASK, c) >= need) pl = cand;
            }
            Lw = pl;
        }
    }
    // compaction of the survivors (at most 128) into the warp's slots
    const unsigned lt = (1u << lane) - 1u;
    int base = 0;
#pragma unroll
    for (int r = 0; r < R; r++) {
        const uint32_t hi = (uint32_t)(key[r] >> 32), lo = (uint32_t)key[r];
        const bool keep = key[r] != 0ull && (hi > T || (hi == T && lo >= Lw));
        const unsigned m = __ballot_sync(VQ_FULL_MASK, keep);
        if (keep) slots[base + __popc(m & lt)] = key[r];
        base += __popc(m);
    }
    for (int e = base + lane; e < 128; e += 32) slots[e] = 0ull;
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 4; r++) out[r] = slots[r * 32 + lane];
    __syncwarp();
    warp_bitonic_sort_desc4(out, lane);
}

// Warp-cooperative prune of one candidate list: sort its n (<= VQ_CAP) keys, keep the best
// min(n, k) in place (sorted descending), return the new length.  thr_out = k-th score if the
// list held at least k keys, else -FLT_MAX.  Called by all 32 lanes with identical arguments.
__device__ __forceinline__ int warp_prune_list(uint64_t* list, int n, int k, int lane, float& thr_out) {
    uint64_t key[VQ_CAP / 32];
#pragma unroll
    for (int r = 0; r < VQ_CAP / 32; r++) {
        const int e = r * 32 + lane;
        key[r] = (e < n) ? raw_to_key(__ldcg(reinterpret_cast<const unsigned long long*>(list) + e)) : 0ull;
    }
    warp_bitonic_sort_desc<VQ_CAP / 32>(key, lane);
    const int m = n < k ? n : k;
#pragma unroll
    for (int r = 0; r < VQ_KMAX / 32; r++) {
        const int e = r * 32 + lane;
        if (e < m) __stcg(reinterpret_cast<unsigned long long*>(list) + e, (unsigned long long)key_to_raw(key[r]));
    }
    uint64_t kth = 0;
    const int rk = (k - 1) >> 5, lk = (k - 1) & 31;
#pragma unroll
    for (int r = 0; r < VQ_KMAX / 32; r++) {
        const uint64_t t = __shfl_sync(VQ_FULL_MASK, key[r], lk);
        if (r == rk) kth = t;
    }
    thr_out = (n >= k) ? key_score(kth) : -FLT_MAX;
    __syncwarp();
    return m;
}

}  // namespace vq
