// Cross-GPU exchange over NVLink peer memory (one process per GPU, one node) + the k-way merge FUSED with it.
//
// The path's only exchange step: every rank holds the local top-k of its row shard and every rank
// needs the merged global top-k.  The reference has no counterpart (FAISS replicates the index over the GPUs,
// datasets/search.py:329-340); round 1 used ncclAllGather (1.5 MB per rank at 1,250 queries — latency-bound) and
// a separate merge kernel.  Here every rank owns one cudaMalloc'ed window [control | slot 0 | slot 1], exported with
// cudaIpcGetMemHandle and mapped by every peer:
//
//   publish  copies the rank's packed (scores, ids) into its OWN slot (parity of the next sequence number), then
//            the last CTA fences at system scope, bumps the rank's device-resident sequence number and stores it
//            into flag[rank] of EVERY peer's control block (8 remote 4-byte stores over NVLink);
//   merge    one CTA per query spins on its LOCAL flags until every rank has published this sequence number, then
//            pulls the query's k entries of every rank's list STRAIGHT FROM THE PEERS' SLOTS (ld.volatile: peer data
//            may only be cached in L1, which these loads bypass) and runs the same register/shared-memory k-way
//            merge as the single-GPU path.  The all-gather and the merge are one kernel; the gathered [G, nq, k]
//            buffer never exists.
//
// Sequence numbers live in device memory (not in kernel arguments), so a captured CUDA graph of local search +
// publish + merge replays correctly.  Two slots suffice: a rank can only publish sequence n+2 after its merge n+1
// finished, which needed every peer's flag n+1, which a peer raises only after its own merge n completed — so nobody
// still reads the slot of parity n when it is overwritten.
#include "vq_common.cuh"
#include "vq_kernels.h"

namespace vq {
namespace xchg {

constexpr int PUBLISH_THREADS = 256;

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_volatile_u32(const void* p) {
    unsigned v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_volatile_u64(const void* p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint4 ld_volatile_v4(const void* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

// src (local, 16-byte aligned, bytes % 16 == 0 after padding by the caller) -> own slot of the next sequence number
__global__ void __launch_bounds__(PUBLISH_THREADS)
publish_kernel(XchgControl* ctl, uint8_t* slots, size_t slot_bytes, const uint8_t* __restrict__ src, size_t bytes, XchgPeers peers, int rank,
               int world) {
    const unsigned next = ctl->seq + 1;  // stable until the last CTA of this launch bumps it
    uint8_t* dst = slots + (size_t)(next & 1u) * slot_bytes;
    const size_t n16 = bytes >> 4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x)
        reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
    for (size_t i = (n16 << 4) + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < bytes; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
    __threadfence();
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) last = atomicAdd(&ctl->done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    // every CTA's copy is visible device-wide (fence + atomic); make it visible to the peers, then raise the flags
    __threadfence_system();
    if (threadIdx.x == 0) {
        ctl->done = 0;
        ctl->seq = next;
        __threadfence_system();
    }
    __syncthreads();
    if ((int)threadIdx.x < world) st_release_sys(&peers.ctl[threadIdx.x]->flag[rank], next);
}

// waits until every rank published the local sequence number (thread r watches flag[r]).  Like any collective,
// every rank has to make the call; the spin is bounded (~30 s) so that a rank that died cannot hang this GPU
// forever: the kernel then gives up and poisons its output (ids -2), which the host entry points check
__device__ __forceinline__ bool wait_all(const XchgControl* ctl, int world, unsigned want) {
    __shared__ int ok_s;
    if (threadIdx.x == 0) ok_s = 1;
    __syncthreads();
    if ((int)threadIdx.x < world) {
        bool ok = false;
        for (long long spin = 0; spin < 30000000ll; spin++) {
            if ((int)(ld_acquire_sys(&ctl->flag[threadIdx.x]) - want) >= 0) { ok = true; break; }
            __nanosleep(40);
        }
        if (!ok) ok_s = 0;
    }
    __syncthreads();
    return ok_s != 0;
}

// a, b: 128 keys each, sorted descending, 4 per lane -> a = best 128 of the union (see vq_merge.cu)
__device__ __forceinline__ void merge_top128(uint64_t (&a)[4], const uint64_t (&b)[4], int lane) {
#pragma unroll
    for (int r = 0; r < 4; r++) a[r] = umax64(a[r], __shfl_sync(VQ_FULL_MASK, b[3 - r], 31 - lane));
#pragma unroll
    for (int r = 0; r < 2; r++) { const uint64_t x = a[r], y = a[r + 2]; a[r] = umax64(x, y); a[r + 2] = umin64(x, y); }
#pragma unroll
    for (int r = 0; r < 4; r += 2) { const uint64_t x = a[r], y = a[r + 1]; a[r] = umax64(x, y); a[r + 1] = umin64(x, y); }
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

// One CTA per query: wait for the flags, pull the query's k_in entries of every rank's published list from the
// peers' slots (scores f32 [nq, k_in] then ids i64 [nq, k_in] at ids_off), merge to the global top k_out.
// T threads per CTA: 256, or 128 when that keeps all queries in one wave (see merge_lists_launch).
template <int T>
__global__ void __launch_bounds__(T)
pull_merge_kernel(const XchgControl* ctl, XchgPeers peers, size_t slot_bytes, size_t ids_off, int world, int nq, int k_in, int k_out,
                  float* __restrict__ out_scores, long long* __restrict__ out_ids) {
    constexpr int MERGE_THREADS = T, MERGE_WARPS = T / 32;
    extern __shared__ uint64_t keys[];  // [cap] gathered keys, reused as [WARPS][128] by the tree
    const int q = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned want = ctl->seq;  // this rank published `want` earlier on this stream
    if (!wait_all(ctl, world, want)) {
        for (int e = threadIdx.x; e < k_out; e += MERGE_THREADS) {
            out_scores[(size_t)q * k_out + e] = -FLT_MAX;
            out_ids[(size_t)q * k_out + e] = -2;  // poisoned: a peer never arrived
        }
        return;
    }
    const size_t par = (size_t)(want & 1u) * slot_bytes;
    const int n = world * k_in;
    for (int i = threadIdx.x; i < n; i += MERGE_THREADS) {
        const int c = i / k_in, j = i - c * k_in;
        const uint8_t* base = peers.slots[c] + par;
        const size_t at = (size_t)q * k_in + j;
        const long long id = (long long)ld_volatile_u64(base + ids_off + at * 8);
        const float s = __uint_as_float(ld_volatile_u32(base + at * 4));
        keys[i] = id >= 0 ? make_key(s, (uint32_t)id) : 0ull;  // padding sorts last
    }
    __syncthreads();
    uint64_t run[4] = {0ull, 0ull, 0ull, 0ull};
    for (int c0 = warp * 256; c0 < n; c0 += MERGE_WARPS * 256) {
        uint64_t ch[8];
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const int e = c0 + r * 32 + lane;
            ch[r] = e < n ? keys[e] : 0ull;
        }
        warp_bitonic_sort_desc<8>(ch, lane);
        const uint64_t top[4] = {ch[0], ch[1], ch[2], ch[3]};
        merge_top128(run, top, lane);
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; r++) keys[warp * 128 + r * 32 + lane] = run[r];
    __syncthreads();
    for (int step = 1; step < MERGE_WARPS; step <<= 1) {
        if ((warp & (2 * step - 1)) == 0) {
            uint64_t other[4];
#pragma unroll
            for (int r = 0; r < 4; r++) other[r] = keys[(warp + step) * 128 + r * 32 + lane];
            merge_top128(run, other, lane);
#pragma unroll
            for (int r = 0; r < 4; r++) keys[warp * 128 + r * 32 + lane] = run[r];
        }
        __syncthreads();
    }
    if (warp == 0) {
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
}

// generic all-gather consumer: dst[r * bytes .. ] = rank r's published bytes (queries travelling between ranks)
__global__ void __launch_bounds__(PUBLISH_THREADS)
pull_gather_kernel(const XchgControl* ctl, XchgPeers peers, size_t slot_bytes, int world, size_t bytes, uint8_t* __restrict__ dst) {
    const unsigned want = ctl->seq;
    if (!wait_all(ctl, world, want)) return;
    const size_t par = (size_t)(want & 1u) * slot_bytes;
    const size_t n16 = bytes >> 4;
    for (int c = blockIdx.y; c < world; c += gridDim.y) {
        const uint8_t* src = peers.slots[c] + par;
        uint8_t* d = dst + (size_t)c * bytes;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x)
            reinterpret_cast<uint4*>(d)[i] = ld_volatile_v4(src + (i << 4));
    }
}

}  // namespace xchg

int xchg_publish_launch(XchgControl* ctl, uint8_t* slots, size_t slot_bytes, const void* src, size_t bytes, const XchgPeers& peers, int rank,
                        int world, cudaStream_t stream) {
    if (bytes > slot_bytes) return -2;
    int blocks = (int)((bytes / 16 + xchg::PUBLISH_THREADS - 1) / xchg::PUBLISH_THREADS);
    if (blocks < 1) blocks = 1;
    if (blocks > 64) blocks = 64;
    xchg::publish_kernel<<<blocks, xchg::PUBLISH_THREADS, 0, stream>>>(ctl, slots, slot_bytes, (const uint8_t*)src, bytes, peers, rank, world);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -(int)e;
}

template <int T>
static int xchg_merge_launch_t(const XchgControl* ctl, const XchgPeers& peers, size_t slot_bytes, size_t ids_off, int cap, int world, int nq, int k_in,
                               int k_out, float* out_scores, long long* out_ids, cudaStream_t stream) {
    static bool attr_set_dev[64] = {};
    int dev_ = 0;
    cudaGetDevice(&dev_);
    if (!attr_set_dev[dev_ & 63]) {
        cudaError_t e = cudaFuncSetAttribute(xchg::pull_merge_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 8);
        if (e != cudaSuccess) return -(int)e;
        attr_set_dev[dev_ & 63] = true;
    }
    xchg::pull_merge_kernel<T><<<nq, T, (size_t)cap * 8, stream>>>(ctl, peers, slot_bytes, ids_off, world, nq, k_in, k_out, out_scores, out_ids);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -(int)e;
}

int xchg_merge_launch(const XchgControl* ctl, const XchgPeers& peers, size_t slot_bytes, int world, int nq, int k_in, int k_out,
                      float* out_scores, long long* out_ids, cudaStream_t stream) {
    if (nq <= 0) return 0;
    if (k_in < 1 || k_out < 1 || k_out > VQ_KMAX || (long long)world * k_in > 8192) return -2;
    int cap = 1024;
    while (cap < world * k_in) cap <<= 1;
    const size_t ids_off = ((size_t)nq * k_in * 4 + 7) / 8 * 8;
    if (nq > 1184 && nq <= 2368 && cap <= 2048)
        return xchg_merge_launch_t<128>(ctl, peers, slot_bytes, ids_off, cap, world, nq, k_in, k_out, out_scores, out_ids, stream);
    return xchg_merge_launch_t<256>(ctl, peers, slot_bytes, ids_off, cap, world, nq, k_in, k_out, out_scores, out_ids, stream);
}

int xchg_gather_launch(const XchgControl* ctl, const XchgPeers& peers, size_t slot_bytes, int world, size_t bytes, void* dst, cudaStream_t stream) {
    if (bytes > slot_bytes || (bytes & 15)) return -2;
    int bx = (int)((bytes / 16 + xchg::PUBLISH_THREADS - 1) / xchg::PUBLISH_THREADS);
    if (bx < 1) bx = 1;
    if (bx > 16) bx = 16;
    xchg::pull_gather_kernel<<<dim3(bx, world), xchg::PUBLISH_THREADS, 0, stream>>>(ctl, peers, slot_bytes, world, bytes, (uint8_t*)dst);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -(int)e;
}

}  // namespace vq
