// FP32 FMA issue-rate microbenchmark for sm_100a: scalar FFMA vs packed FFMA2 vs a mix.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_rate fma_rate.cu && ./fma_rate
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

template <int MODE>   // 0: scalar FFMA x16 chains, 1: FFMA2 x16 chains, 2: 8 FFMA2 + 8 FFMA, 3: 8 FFMA2 + 16 FFMA, 4: FFMA2 with broadcast-scalar operand
__global__ void k(float *out, int iters, float x)
{
    float s[16];
    u64 p[16];
    for (int i = 0; i < 16; ++i) { s[i] = threadIdx.x + i; p[i] = ((u64)__float_as_uint(s[i]) << 32) | __float_as_uint(s[i] + 1.f); }
    const u64 xx = ((u64)__float_as_uint(x) << 32) | __float_as_uint(x);
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) s[i] = fma1(s[i], x, s[i]);
        } else if (MODE == 1 || MODE == 4) {
#pragma unroll
            for (int i = 0; i < 16; ++i) p[i] = fma2(p[i], xx, p[i]);
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { p[i] = fma2(p[i], xx, p[i]); s[i] = fma1(s[i], x, s[i]); }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) { p[i] = fma2(p[i], xx, p[i]); s[2 * i] = fma1(s[2 * i], x, s[2 * i]); s[2 * i + 1] = fma1(s[2 * i + 1], x, s[2 * i + 1]); }
        }
    }
    float acc = 0;
    for (int i = 0; i < 16; ++i) acc += s[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char *name, double fma_per_thread_iter, int warps_per_sm)
{
    int dev = 0, sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    float *out;
    const int threads = 32 * warps_per_sm, iters = 20000;
    cudaMalloc(&out, sizeof(float) * sms * threads);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<sms, threads>>>(out, 100, 1.0001f);
    cudaEventRecord(a);
    k<MODE><<<sms, threads>>>(out, iters, 1.0001f);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double fma = fma_per_thread_iter * iters * (double)threads * sms;
    printf("%-34s warps/SM %2d: %7.2f TFMA/s = %6.1f FMA/clk/SM at max clock %d MHz (%.3f ms)\n", name, warps_per_sm,
           fma / ms / 1e9, fma / (ms * 1e-3) / sms / (khz * 1e3), khz / 1000, ms);
    cudaFree(out);
}

int main()
{
    for (int w : {4, 8, 16, 32}) {
        run<0>("scalar FFMA (3-reg)", 16, w);
        run<1>("packed FFMA2", 32, w);
        run<2>("8 FFMA2 + 8 FFMA", 24, w);
        run<3>("8 FFMA2 + 16 FFMA", 32, w);
    }
    return 0;
}
