// Feasibility experiment for DESIGN.md section 10 item 1: tcgen05.mma kind::tf32, M = 128 x N = 32 x K = 8, operands in
// shared memory in the no-swizzle K-major canonical layout, accumulator in TMEM.
//   (1) correctness of hand-built shared-memory / instruction descriptors against a host reference;
//   (2) MMA issue rate at N = 32 (the [S_re | S_im] x 16 antennas operand of the correlator).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 scripts/microbench/umma_tf32.cu -o umma_tf32 && ./umma_tf32
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int M = 128, N = 32, KSTEP = 8;           // one MMA: 128 x 32 x 8 (tf32: 32 bytes of K per row)
constexpr int KSTEPS = 32;                          // 256 samples per correlator tile
constexpr int A_STEP_BYTES = M * KSTEP * 4;         // 4096
constexpr int B_STEP_BYTES = N * KSTEP * 4;         // 1024

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// no-swizzle K-major canonical layout (cute mma_traits_sm100.hpp: ((8,n),2):((1,SBO),LBO) in 16-byte units):
// 8 rows x 16 bytes form one contiguous 128-byte core matrix; SBO steps to the next 8 rows, LBO to the next 16 bytes of K
__host__ __device__ inline int canon_off(int row, int k, int rows)   // byte offset inside one K-step block
{
    return (k / 4) * (rows * 16) + (row / 8) * 128 + (row % 8) * 16 + (k % 4) * 4;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;          // descriptor version (Blackwell)
    return d;                        // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}
constexpr uint32_t kIdesc = (1u << 4)      // D = F32
                            | (2u << 7)    // A = TF32
                            | (2u << 10)   // B = TF32
                            | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);   // K-major A and B

__global__ void __launch_bounds__(128, 1) umma_kernel(const float *A, const float *B, float *D, int repeats, long long *cycles)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *sA = smem;                                   // [KSTEPS][4096]
    unsigned char *sB = smem + KSTEPS * A_STEP_BYTES;           // [KSTEPS][1024]
    __shared__ uint32_t tmem_base;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;

    for (int i = tid; i < M * KSTEP * KSTEPS; i += blockDim.x) {
        const int row = i / (KSTEP * KSTEPS), kk = i % (KSTEP * KSTEPS);
        *reinterpret_cast<float *>(sA + (kk / KSTEP) * A_STEP_BYTES + canon_off(row, kk % KSTEP, M)) = A[i];
    }
    for (int i = tid; i < N * KSTEP * KSTEPS; i += blockDim.x) {
        const int row = i / (KSTEP * KSTEPS), kk = i % (KSTEP * KSTEPS);
        *reinterpret_cast<float *>(sB + (kk / KSTEP) * B_STEP_BYTES + canon_off(row, kk % KSTEP, N)) = B[i];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy smem writes -> visible to the MMA (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;

    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        t0 = clock64();
        for (int r = 0; r < repeats; ++r) {
            for (int j = 0; j < KSTEPS; ++j) {
                const uint64_t da = make_desc(smem_u32(sA + j * A_STEP_BYTES), M * 16, 128);
                const uint64_t db = make_desc(smem_u32(sB + j * B_STEP_BYTES), N * 16, 128);
                const uint32_t acc = (r > 0 || j > 0) ? 1u : 0u;
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                    ::"r"(tmem), "l"(da), "l"(db), "r"(kIdesc), "r"(acc)
                    : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // everybody waits for the MMAs (phase 0)
    {
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    }
    if (tid == 0) {
        t1 = clock64();
        if (cycles) cycles[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // epilogue: warp w owns TMEM lanes 32w .. 32w+31 (= rows), 32 fp32 columns each
    uint32_t v[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
        "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
          "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (blockIdx.x == 0)
        for (int j = 0; j < 32; ++j) D[tid * N + j] = __uint_as_float(v[j]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem) : "memory");
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

int main()
{
    const int K = KSTEP * KSTEPS;
    std::vector<float> hA(M * K), hB(N * K), hD(M * N), ref(M * N, 0.f);
    srand(1);
    for (auto &x : hA) x = (float)(rand() % 7 - 3);          // small integers: exact in TF32, exact sums
    for (auto &x : hB) x = (float)(rand() % 5 - 2);
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < N; ++j) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)hA[i * K + k] * hB[j * K + k];
            ref[i * N + j] = (float)s;
        }
    float *dA, *dB, *dD;
    long long *dC;
    CK(cudaMalloc(&dA, hA.size() * 4)); CK(cudaMalloc(&dB, hB.size() * 4)); CK(cudaMalloc(&dD, hD.size() * 4));
    CK(cudaMalloc(&dC, 148 * sizeof(long long)));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice));
    const int smem = KSTEPS * (A_STEP_BYTES + B_STEP_BYTES);
    CK(cudaFuncSetAttribute(umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    umma_kernel<<<1, 128, smem>>>(dA, dB, dD, 1, dC);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    double worst = 0;
    for (int i = 0; i < M * N; ++i) {
        const double e = std::abs((double)hD[i] - ref[i]);
        if (e > worst) worst = e;
        if (e > 1e-3 && bad++ < 5) std::printf("  mismatch at (%d,%d): got %g want %g\n", i / N, i % N, hD[i], ref[i]);
    }
    std::printf("{\"check\": \"tcgen05.mma kind::tf32 128x32x8 x %d K-steps vs host\", \"max_abs_err\": %g, \"mismatches\": %d}\n", KSTEPS, worst, bad);
    for (int grid : {1, 148}) {
        const int repeats = 200;
        umma_kernel<<<grid, 128, smem>>>(dA, dB, dD, repeats, dC);
        CK(cudaDeviceSynchronize());
        std::vector<long long> hc(grid);
        CK(cudaMemcpy(hc.data(), dC, grid * sizeof(long long), cudaMemcpyDeviceToHost));
        long long mx = 0;
        for (auto c : hc) mx = c > mx ? c : mx;
        const double per = (double)mx / (repeats * KSTEPS);
        std::printf("{\"grid\": %d, \"mmas_per_cta\": %d, \"cycles_per_mma_128x32x8_tf32\": %.2f, \"mac_per_clk_per_sm\": %.0f}\n", grid,
                    repeats * KSTEPS, per, M * N * KSTEP / per);
    }
    return 0;
}
