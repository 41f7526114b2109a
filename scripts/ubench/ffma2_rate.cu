// Microbenchmark: issue/pipe rate of FFMA vs FFMA2 (fma.rn.f32x2) on sm_100a.  Prints FMA lanes per clock per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate ffma2_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>  // 0: scalar FFMA, 1: FFMA2 with scalar-broadcast a, 2: FFMA2 with pair a
__global__ void __launch_bounds__(256) rate_kernel(float* out, int iters, float a0, float b0) {
    unsigned long long acc[32];
    float facc[64];
#pragma unroll
    for (int i = 0; i < 32; i++) acc[i] = 0ull;
#pragma unroll
    for (int i = 0; i < 64; i++) facc[i] = 0.f;
    // operands: a0 == 0 -> all-zero data; otherwise per-thread pseudo-random values in (-1, 1) (the accumulators random-walk)
    uint32_t h = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    float a = a0 == 0.f ? 0.f : ((h & 0xffff) / 32768.0f - 1.0f), b = a0 == 0.f ? 0.f : (((h >> 16) & 0xffff) / 32768.0f - 1.0f) * b0;
    unsigned long long bp, ap;
    asm("mov.b64 %0, {%1, %2};" : "=l"(bp) : "f"(b), "f"(b + 1.f));
    asm("mov.b64 %0, {%1, %2};" : "=l"(ap) : "f"(a), "f"(a + 1.f));
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 64; i++) facc[i] = fmaf(a, b, facc[i]);
        } else if (MODE == 1) {
            unsigned long long aa;
            asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
#pragma unroll
            for (int i = 0; i < 32; i++) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[i]) : "l"(aa), "l"(bp));
        } else {
#pragma unroll
            for (int i = 0; i < 32; i++) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[i]) : "l"(ap), "l"(bp));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; i++) s += __uint_as_float((uint32_t)acc[i]) + __uint_as_float((uint32_t)(acc[i] >> 32));
#pragma unroll
    for (int i = 0; i < 64; i++) s += facc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static void run(const char* name, int ctas_per_sm, float a0 = 1.f) {
    int dev = 0, sms = 0, khz = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    float* out;
    const int grid = sms * ctas_per_sm, iters = 20000;
    cudaMalloc(&out, (size_t)grid * 256 * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    rate_kernel<MODE><<<grid, 256>>>(out, 1000, a0, 1.f);
    cudaEventRecord(e0);
    rate_kernel<MODE><<<grid, 256>>>(out, iters, a0, 1.f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fma = (double)grid * 256 * iters * 64;
    printf("%-34s ctas/sm=%d  %.3f ms  %.2f TFLOP/s  %.1f FMA/clk/SM at the max clock %d MHz\n", name, ctas_per_sm, ms, 2 * fma / ms / 1e9,
           fma / (ms * 1e-3) / sms / (khz * 1e3), khz / 1000);
    cudaFree(out);
}

int main() {
    for (int c = 1; c <= 2; c++) {
        run<0>("FFMA (scalar), random data", c);
        run<1>("FFMA2 (.F32 broadcast a), random", c);
        run<2>("FFMA2 (pair a), random", c);
        run<0>("FFMA (scalar), zeros", c, 0.f);
        run<1>("FFMA2 (.F32 broadcast a), zeros", c, 0.f);
    }
    return 0;
}
