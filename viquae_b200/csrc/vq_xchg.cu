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

constexpr int PM_WARPS = 4;             // queries per CTA (one warp each)
constexpr int PM_R = 32;
constexpr int PM_BATCH = 32 * PM_R;     // keys per selection batch

// One WARP per query (4 per CTA): the CTA waits for the flags, then every warp pulls its query's k_in entries of every
// rank's published list straight from the peers' slots (scores f32 [nq, k_in] then ids i64 [nq, k_in] at ids_off),
// whole lists at a time in batches of <= 1024 keys, selects each batch's top 128 (warp_select_top128) and folds it
// into the running top-128 -> global top k_out.
__global__ void __launch_bounds__(PM_WARPS * 32)
pull_merge_kernel(const XchgControl* ctl, XchgPeers peers, size_t slot_bytes, size_t ids_off, int world, int nq, int k_in, int k_out,
                  float* __restrict__ out_scores, long long* __restrict__ out_ids) {
    extern __shared__ uint64_t pm_smem[];  // per warp: [128] slots of the selection
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.x * PM_WARPS + warp;
    const unsigned want = ctl->seq;  // this rank published `want` earlier on this stream
    const bool arrived = wait_all(ctl, world, want);
    if (q >= nq) return;
    if (!arrived) {
        for (int e = lane; e < k_out; e += 32) {
            out_scores[(size_t)q * k_out + e] = -FLT_MAX;
            out_ids[(size_t)q * k_out + e] = -2;  // poisoned: a peer never arrived
        }
        return;
    }
    uint64_t* slots = pm_smem + (size_t)warp * 128;
    const size_t par = (size_t)(want & 1u) * slot_bytes;
    const int per = PM_BATCH / k_in;  // whole lists per batch (k_in <= 128 -> >= 8)
    uint64_t run[4] = {0ull, 0ull, 0ull, 0ull};
    for (int c0 = 0; c0 < world; c0 += per) {
        const int c1 = min(world, c0 + per);
        const int n = (c1 - c0) * k_in;
        // every remote load of the batch is issued before the first one is used: a peer round trip costs ~2-3 us over
        // NVLink, and a loop that consumes each value before asking for the next would pay it 25 times in a row
        uint64_t key[PM_R];
        uint32_t sc[PM_R];
#pragma unroll
        for (int r = 0; r < PM_R; r++) {
            const int e = r * 32 + lane;
            key[r] = ~0ull;  // id -1
            sc[r] = 0u;
            if (e < n) {
                const int c = c0 + e / k_in, j = e - (e / k_in) * k_in;
                const uint8_t* base = peers.slots[c] + par;
                const size_t at = (size_t)q * k_in + j;
                key[r] = ld_volatile_u64(base + ids_off + at * 8);
                sc[r] = ld_volatile_u32(base + at * 4);
            }
        }
#pragma unroll
        for (int r = 0; r < PM_R; r++) {
            const long long id = (long long)key[r];
            key[r] = id >= 0 ? make_key(__uint_as_float(sc[r]), (uint32_t)id) : 0ull;  // padding sorts last
        }
        uint64_t top[4];
        warp_select_top128<PM_R>(key, top, slots, lane);
        merge_top128(run, top, lane);
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

int xchg_merge_launch(const XchgControl* ctl, const XchgPeers& peers, size_t slot_bytes, int world, int nq, int k_in, int k_out,
                      float* out_scores, long long* out_ids, cudaStream_t stream) {
    if (nq <= 0) return 0;
    if (k_in < 1 || k_in > VQ_KMAX || k_out < 1 || k_out > VQ_KMAX) return -2;
    const size_t ids_off = ((size_t)nq * k_in * 4 + 7) / 8 * 8;
    const size_t smem = (size_t)xchg::PM_WARPS * 128 * sizeof(uint64_t);  // 4 KB
    static bool attr_set_dev[64] = {};
    int dev_ = 0;
    cudaGetDevice(&dev_);
    if (!attr_set_dev[dev_ & 63]) {
        cudaError_t e = cudaFuncSetAttribute(xchg::pull_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return -(int)e;
        attr_set_dev[dev_ & 63] = true;
    }
    xchg::pull_merge_kernel<<<(nq + xchg::PM_WARPS - 1) / xchg::PM_WARPS, xchg::PM_WARPS * 32, smem, stream>>>(ctl, peers, slot_bytes, ids_off, world, nq, k_in,
                                                                                                             k_out, out_scores, out_ids);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -(int)e;
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
