// How do FFMA2 streams share issue slots with LDS / ALU / MUFU work on sm_100a?
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 fma2b(u64 a, float b, u64 c) { u64 d; asm volatile("{.reg .b64 t; mov.b64 t, {%2, %2}; fma.rn.f32x2 %0, %1, t, %3;}" : "=l"(d) : "l"(a), "f"(b), "l"(c)); return d; }
__device__ __forceinline__ float lds(unsigned a) { float v; asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }

// per iteration: 16 FFMA2 (+ NL shared loads feeding them as broadcast operands) (+ NA integer adds) (+ NM mufu)
template <int NL, int NA, int NM>
__global__ void k(float *out, int iters, float x)
{
    __shared__ float sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = 1.0f + 1e-6f * i;
    __syncthreads();
    u64 p[16];
    for (int i = 0; i < 16; ++i) p[i] = ((u64)__float_as_uint(threadIdx.x + i) << 32) | __float_as_uint(threadIdx.x + 1.f);
    unsigned base = (unsigned)__cvta_generic_to_shared(sm) + 4 * (threadIdx.x & 31);
    unsigned ia[8] = {1, 2, 3, 4, 5, 6, 7, 8};
    float mu = x;
    for (int it = 0; it < iters; ++it) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (j < NL) ? lds(base + 128 * j + ((it & 7) << 10)) : x;
#pragma unroll
        for (int j = 0; j < NA; ++j) asm volatile("add.u32 %0, %0, %1;" : "+r"(ia[j & 7]) : "r"(it));
#pragma unroll
        for (int j = 0; j < NM; ++j) mu = __sinf(mu);
#pragma unroll
        for (int i = 0; i < 16; ++i) p[i] = fma2b(p[i], v[i & 7], p[i]);
    }
    float acc = mu;
    for (int i = 0; i < 16; ++i) acc += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
    for (int j = 0; j < 8; ++j) acc += ia[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int NL, int NA, int NM>
void run(int warps)
{
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float *out; cudaMalloc(&out, sizeof(float) * sms * 32 * warps);
    const int iters = 20000;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<NL, NA, NM><<<sms, 32 * warps>>>(out, 100, 1.0001f);
    cudaEventRecord(a);
    k<NL, NA, NM><<<sms, 32 * warps>>>(out, iters, 1.0001f);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double cyc_per_iter_per_smsp = ms * 1e-3 * khz * 1e3 / iters / (warps / 4.0);
    printf("16 FFMA2 + %d LDS + %2d IADD + %d MUFU, %2d warps/SM: %6.1f cycles per warp-iteration per SMSP (FMA floor 32) -> FMA pipe %.0f%%\n",
           NL, NA, NM, warps, cyc_per_iter_per_smsp, 3200.0 / cyc_per_iter_per_smsp);
    cudaFree(out);
}

int main()
{
    for (int w : {4, 12}) {
        run<0, 0, 0>(w);
        run<4, 0, 0>(w);
        run<8, 0, 0>(w);
        run<0, 8, 0>(w);
        run<0, 16, 0>(w);
        run<8, 8, 0>(w);
        run<8, 8, 1>(w);
    }
    return 0;
}
