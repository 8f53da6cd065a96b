// libgat tensor-core path (opt-in, GAT_TENSOR_TF32): the correlator for signal blocks shared by MANY satellite channels,
// as a GEMM on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM).
//
//   C[(k, l), (plane, m)] = sum_n W[(k, l), n] * S[n, (plane, m)],      W[(k, l), n] = chip_{k,l}[n] * conj(carrier_k[n])
//
// One channel's contraction is 3..11 taps x 1..16 antennas -- too skinny for an MMA, which is why the default kernel
// (gat_correlate.cu) runs on the FP32 pipes.  With K >= 32 channels over ONE block the rows (channel, tap) fill the
// 128-row MMA, and the FP32 loop is issue-bound, so the arithmetic moves to the tensor pipe
// and the CUDA cores only GENERATE W:
//   * B operand  = the signal tile [256 samples x (2 planes x 16 antennas)] exactly as ONE 4-D TMA load delivers it in
//                  the no-swizzle K-major canonical layout [sample/4][plane][antenna][sample%4]; a pass over the tile
//                  rounds it to TF32 (truncation would bias the magnitude) and zeroes samples outside the range.
//   * A operand  = W, 32 channels x 4 tap rows = 128 rows, generated in 32-sample chunks into double-buffered
//                  shared memory, directly in the canonical layout: lane = (row % 8, sample % 4), so every store
//                  instruction of a warp writes one contiguous 128-byte core matrix.  Per channel one replica row (sign
//                  bits, ballot-packed) per tile and one carrier row (TF32 cos, -sin) per chunk are shared by the tap
//                  rows, whose entries are then load - sign flip - store.
//   * D          = two 128 x 32 FP32 accumulators in TMEM: C_r = W_re x [S_re | S_im], C_i = W_im x [S_re | S_im];
//                  acc_re = C_r.re - C_i.im, acc_im = C_i.re + C_r.im in the epilogue.
// Work split: the flattened (period, channel group, tile) space is divided evenly over the CTAs (like the default
// kernel); every CTA writes one partial per job it touches and a second, tiny kernel sums the partials in a fixed order
// (deterministic, no float atomics).
// Numerics: W and S are rounded to TF32 (10-bit mantissa, round-to-nearest), products and sums are FP32 in the tensor
// core: integer samples of <= 12 bits are exact, and the error of the accumulators is ~3e-4 * sqrt(N) * rms(s) -- a few
// 1e-6 of the prompt of a full-strength signal over 50 000 samples (tests/test_gpu_tensor.py).  The chip indices are
// the same bit-exact Int64 NCO as in the default kernel.
#include "gat_internal.h"

namespace gat {

namespace {

constexpr int kTcRows = 128, kTcRowsPerSat = 4, kTcSats = kTcRows / kTcRowsPerSat, kTcAnts = 16, kTcCols = 2 * kTcAnts;
constexpr int kTcTile = 256, kTcChunk = 64, kTcChunks = kTcTile / kTcChunk, kTcSteps = kTcChunk / 8;
constexpr int kTcAStep = kTcRows * 8 * 4;                  // one K-step (8 samples) of A: 4096 B
constexpr int kTcAChunk = kTcSteps * kTcAStep;             // 32 KB per part (re / im) and buffer
constexpr int kTcBGroup = kTcCols * 16;                    // 4 samples of all 32 columns: 512 B
constexpr int kTcBTile = (kTcTile / 4) * kTcBGroup;        // 32 KB
constexpr int kTcGenWarps = 16;
constexpr int kTcThreads = 32 * (kTcGenWarps + 1);
constexpr int kTcTabWords = 32;                            // chip table of a channel as sign bits: 1024 chips in 128 B
constexpr int kTcRepWords = 20;                            // replica sign bits per channel and tile: <= 512 entries (+ one spare word)
constexpr int kTcSmemBytes = 2 * kTcBTile + 4 * kTcAChunk + kTcSats * kTcTabWords * 4 + kTcSats * kTcRepWords * 4 + kTcSats * kTcChunk * 8;

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint32_t a, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(n)); }
__device__ __forceinline__ void bar_arrive(uint32_t a) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory"); }
__device__ __forceinline__ void bar_expect(uint32_t a, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait(uint32_t a, uint32_t parity)
{
    // tight polling: the wake-up latency of these waits is on the critical path of the A-buffer hand-over, so the
    // watchdog clock is looked at only once in 256 failed polls (reading %globaltimer in every poll cost ~1 us per wait)
    uint32_t done = 0, polls = 0;
    long long t0 = 0;
    while (true) {
        // (with a suspend-time hint the thread sleeps in hardware until the phase flips -- the 16 generator warps
        // polling here took 28 % of all issued instructions)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity), "r"(20000u) : "memory");
        if (done) break;
        if ((++polls & 255u) == 0) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 8000000000ll) __trap();       // ~4 s: a broken launch becomes a CUDA error, not a hang
        }
    }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *map, int c3, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(dst), "l"(map), "r"(0), "r"(0), "r"(0), "r"(c3), "r"(bar) : "memory");
}
// shared-memory matrix descriptor, no swizzle, K-major: core matrix = 8 rows x 16 B; SBO = next 8 rows, LBO = next 16 B of K
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((addr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | (1ull << 46);
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, N = 32, M = 128
constexpr uint32_t kTcIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTcCols >> 3) << 17) | ((uint32_t)(kTcRows >> 4) << 24);
__device__ __forceinline__ void umma_tf32(uint32_t tmem, uint64_t da, uint64_t db, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem), "l"(da), "l"(db), "r"(kTcIdesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t tf32_rna(float x)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
        "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
          "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ int tc_owner(int64_t x, int grid, int64_t total) { return (int)(((x + 1) * grid - 1) / total); }

}  // namespace

__global__ void __launch_bounds__(kTcThreads, 1) correlate_tc_kernel(const __grid_constant__ TcArgs args)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *sB = smem;                                            // [2][kTcBTile]
    unsigned char *sA = sB + 2 * kTcBTile;                               // [2 buffers][re, im][kTcAChunk]
    uint32_t *sTab = reinterpret_cast<uint32_t *>(sA + 4 * kTcAChunk);   // [32][32] chip tables as sign bits
    uint32_t *sRep = sTab + kTcSats * kTcTabWords;                       // [32][20] sign bits of the tile's replica
    float2 *sCar = reinterpret_cast<float2 *>(sRep + kTcSats * kTcRepWords);        // [32][64] (cos, -sin) as TF32, private to the owning warp
    __shared__ uint32_t tmem_base;
    __shared__ __align__(8) uint64_t bars[7];   // 0,1 B full; 2,3 A full; 4,5 A free; 6 accumulators ready
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar0 = s32(bars);
    const uint32_t B_FULL = bar0, A_FULL = bar0 + 16, A_FREE = bar0 + 32, ACC = bar0 + 48;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            bar_init(B_FULL + 8 * i, 1);
            bar_init(A_FULL + 8 * i, kTcGenWarps);
            bar_init(A_FREE + 8 * i, 1);
        }
        bar_init(ACC, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kTcGenWarps) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(s32(&tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // rows of taps that do not exist (n_taps < 4) are never written: both A buffers start as zeros
    for (int i = tid; i < 4 * kTcAChunk / 16; i += kTcThreads) reinterpret_cast<uint4 *>(sA)[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;

    const int K = args.n_sats, G = args.G, TJ = args.tiles_per_job, L = args.n_taps;
    const int64_t TT = args.total_units;
    const int grid = gridDim.x;
    const int64_t r0 = (int64_t)blockIdx.x * TT / grid, r1 = (int64_t)(blockIdx.x + 1) * TT / grid;

    // lane roles of the tap-row phase: this warp owns row group `warp` = channels 2 warp, 2 warp + 1
    const int r8 = lane >> 2, k4 = lane & 3, tap = r8 & 3;
    const int my_sat = 2 * warp + (r8 >> 2);    // the channel of this lane's tap row
    uint32_t qa = 0;      // running chunk counter (A buffer + parity)
    uint32_t qb = 0;      // running tile counter of this CTA (B stage + parity)
    uint32_t seg = 0;

    for (int64_t u = r0; u < r1; ++seg) {
        const int job = (int)(u / TJ);
        const int t_first = (int)(u - (int64_t)job * TJ);
        const int t_last = (int)min((int64_t)TJ, (int64_t)t_first + (r1 - u));
        const int p = job / G, grp = job % G;
        const TcPeriod *per = &args.periods[p];

        // ---- segment set-up: every generator warp owns channels 2 w and 2 w + 1 of the group -- their chip tables, replica
        // bits, carrier rows and tap rows -- so the generator warps never wait for each other, only for the MMA thread ----
        uint64_t frac[2] = {0, 0}, cphl[2] = {0, 0}, cdel32[2] = {0, 0}, ndel[2] = {0, 0};
        uint32_t bmod[2] = {0, 0}, lc[2] = {1, 1};
        int fp[2] = {32, 32};
        bool live[2] = {false, false};
        if (warp < kTcGenWarps) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int sl = 2 * warp + h, k = grp * kTcSats + sl;
                live[h] = k < K;
                if (live[h]) {
                    const SatDev *sd = &args.sats[(size_t)p * K + k];
                    const int8_t *code = sd->code;
                    const int clen = sd->code_len;
                    // 16 chips per lane and step (one 16-byte load; the columns are zero-padded to 16 B): their sign bits
                    // are squeezed into 16 bits, two lanes make one word -- 2 steps for a 1023-chip code instead of 32
                    // dependent byte loads + ballots
                    for (int i0 = 0; i0 < kTcTabWords * 32; i0 += 512) {
                        const int ci = i0 + lane * 16;
                        uint4 w = make_uint4(0, 0, 0, 0);
                        if (ci < clen) w = *reinterpret_cast<const uint4 *>(code + ci);
                        auto squeeze = [](uint32_t x) { return ((x >> 7) & 1u) | ((x >> 14) & 2u) | ((x >> 21) & 4u) | ((x >> 28) & 8u); };
                        const uint32_t half = squeeze(w.x) | (squeeze(w.y) << 4) | (squeeze(w.z) << 8) | (squeeze(w.w) << 12);
                        const uint32_t other = __shfl_xor_sync(0xffffffffu, half, 1);
                        if (!(lane & 1)) sTab[sl * kTcTabWords + (ci >> 5)] = half | (other << 16);
                    }
                    ndel[h] = (uint64_t)sd->nco_delta;
                    fp[h] = sd->nco_fp;
                    lc[h] = (uint32_t)clen;
                    // state at the first sample of tile t_first (relative index n0 may be < 0 for the alignment head)
                    const int64_t n0 = (int64_t)args.aligned_start + (int64_t)t_first * kTcTile - args.start_sample;
                    const __int128 tot = (__int128)(n0 + args.shift0) * (__int128)sd->nco_delta + (__int128)sd->nco_start;
                    int64_t b = (int64_t)(tot >> sd->nco_fp) % clen;
                    if (b < 0) b += clen;
                    bmod[h] = (uint32_t)b;
                    frac[h] = (uint64_t)tot & ((1ull << sd->nco_fp) - 1ull);
                    // carrier phase (Q0.64) of THIS LANE's sample in the next chunk, advanced by 32 samples per chunk
                    cphl[h] = sd->car_phase + (uint64_t)(n0 + lane) * sd->car_delta;
                    cdel32[h] = 32ull * sd->car_delta;
                }
            }
        }
        __syncthreads();   // tables in place; previous segment's epilogue done (TMEM free)

        for (int t = t_first; t < t_last; ++t, ++qb) {
            const uint32_t st = qb & 1u;
            const int n0 = args.aligned_start + t * kTcTile - args.start_sample;      // relative index of tile sample 0
            const uint32_t bt = s32(sB + st * kTcBTile);
            if (warp == kTcGenWarps) {
                // =========================== control warp: TMA + MMA issue, one thread ===========================
                if (lane == 0) {
                    const int c3 = (args.aligned_start + t * kTcTile) / 4;
                    if (qb == 0 && !(args.debug & 16)) {                             // the CTA's very first tile
                        bar_expect(B_FULL + 8 * st, kTcBTile);
                        tma_load_4d(bt, &per->map, c3, B_FULL + 8 * st);
                    }
                    // prefetch the next tile of this CTA (same job or the next one) into the other stage
                    const int64_t un = u + (t - t_first) + 1;
                    if (un < r1 && !(args.debug & 16)) {
                        const int jn = (int)(un / TJ), tn = (int)(un - (int64_t)jn * TJ);
                        if (qb >= 1) {                                               // its previous reader, tile qb - 1, is done:
                            const uint32_t ql = qb * kTcChunks - 1;                  // ... that tile's last chunk has been committed
                            bar_wait(A_FREE + 8 * (ql & 1u), (ql >> 1) & 1u);
                        }
                        bar_expect(B_FULL + 8 * (st ^ 1u), kTcBTile);
                        tma_load_4d(s32(sB + (st ^ 1u) * kTcBTile), &args.periods[jn / G].map, (args.aligned_start + tn * kTcTile) / 4,
                                    B_FULL + 8 * (st ^ 1u));
                    }
                    // descriptors differ only in their 14-bit start-address field: one base per operand, then 64-bit adds of
                    // small constants -- the issuing thread shares its scheduler with four generator warps, so every
                    // instruction it needs per MMA is time the tensor pipe idles
                    const uint64_t da0 = umma_desc(s32(sA), kTcRows * 16, 128);
                    const uint64_t db0 = umma_desc(bt, kTcBGroup, 128);
                    uint32_t q = qa;
                    uint32_t acc = (t > t_first) ? 1u : 0u;
                    for (int c = 0; c < kTcChunks; ++c, ++q) {
                        const uint32_t buf = q & 1u;
                        bar_wait(A_FULL + 8 * buf, (q >> 1) & 1u);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint64_t da_re = da0 + (uint64_t)((buf * 2 * kTcAChunk) >> 4);
                        const uint64_t da_im = da_re + (uint64_t)(kTcAChunk >> 4);
                        const uint64_t db = db0 + (uint64_t)(((uint32_t)c * (kTcChunk / 4) * kTcBGroup) >> 4);
#pragma unroll
                        for (int j = 0; j < kTcSteps; ++j) {
                            if (args.debug & 2) break;
                            umma_tf32(tmem, da_re + (uint64_t)((j * kTcAStep) >> 4), db + (uint64_t)((j * 2 * kTcBGroup) >> 4), acc);
                            umma_tf32(tmem + 32, da_im + (uint64_t)((j * kTcAStep) >> 4), db + (uint64_t)((j * 2 * kTcBGroup) >> 4), acc);
                            acc = 1u;
                        }
                        umma_commit(A_FREE + 8 * buf);
                    }
                    if (t == t_last - 1) umma_commit(ACC);
                }
                qa += kTcChunks;
                continue;
            }

            // ================================= generator warps =================================
            // ---- replica sign bits of this tile for the warp's two channels: entry e <-> sample n0 + e + shift0 ----
            __syncwarp();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (!live[h] || (args.debug & 64)) continue;
                const int sl = 2 * warp + h;
                const uint32_t tab_s = s32(sTab + sl * kTcTabWords);
                const int sh = fp[h] - 32;
                uint64_t v = frac[h] + (uint64_t)lane * ndel[h];
                const uint64_t v32 = 32ull * ndel[h];
                const int rows = (kTcTile + args.span + 31) >> 5;
                for (int r = 0; r < rows; r += 2) {                 // two independent table lookups in flight (the row count is
                    uint32_t word[2], sft[2];                       // rounded up to even: the buffer has 20 words per channel)
#pragma unroll
                    for (int j = 0; j < 2; ++j, v += v32) {
                        uint32_t idx = bmod[h] + ((uint32_t)(v >> 32) >> sh);
                        idx = min(idx, idx - lc[h]);                // single wrap (host-checked)
                        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(word[j]) : "r"(tab_s + 4u * (idx >> 5)));
                        sft[j] = idx & 31u;
                    }
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const uint32_t bits = __ballot_sync(0xffffffffu, (word[j] >> sft[j]) & 1u);
                        if (lane == j) sRep[sl * kTcRepWords + r + j] = bits;
                    }
                }
                // advance the NCO base to the next tile
                const unsigned __int128 nf = (unsigned __int128)frac[h] + (unsigned __int128)kTcTile * (unsigned __int128)ndel[h];
                uint64_t x = (uint64_t)bmod[h] + (uint64_t)(nf >> fp[h]);
                if (x >= lc[h]) x %= lc[h];
                bmod[h] = (uint32_t)x;
                frac[h] = (uint64_t)nf & ((1ull << fp[h]) - 1ull);
            }
            // ---- the signal tile: wait for the TMA, round to TF32, zero what lies outside [0, n_samples) ----
            if (!(args.debug & 16)) bar_wait(B_FULL + 8 * st, (qb >> 1) & 1u);
            {
                unsigned char *bp = sB + st * kTcBTile;
                const bool edge = n0 < 0 || n0 + kTcTile > args.n_samples;            // only a job's first / last tile
                for (int i = tid; i < kTcBTile / 16 && !(args.debug & 8); i += 32 * kTcGenWarps) {        // 16 B = 4 samples of one column
                    uint4 w = reinterpret_cast<uint4 *>(bp)[i];
                    w.x = tf32_rna(__uint_as_float(w.x));
                    w.y = tf32_rna(__uint_as_float(w.y));
                    w.z = tf32_rna(__uint_as_float(w.z));
                    w.w = tf32_rna(__uint_as_float(w.w));
                    if (edge) {
                        const int n = n0 + 4 * (i >> 5);                              // i >> 5 = sample group in the tile
                        if (n + 0 < 0 || n + 0 >= args.n_samples) w.x = 0u;
                        if (n + 1 < 0 || n + 1 >= args.n_samples) w.y = 0u;
                        if (n + 2 < 0 || n + 2 >= args.n_samples) w.z = 0u;
                        if (n + 3 < 0 || n + 3 >= args.n_samples) w.w = 0u;
                    }
                    reinterpret_cast<uint4 *>(bp)[i] = w;
                }
            }
            const uint32_t car_w = s32(sCar + (2 * warp) * kTcChunk);               // this warp's two carrier rows
            // carrier phase of this lane's sample inside the tile: 32 bits (2^-32 cycle), restarted exactly from the
            // 64-bit accumulator at every tile; a step is 32 samples.  Only a job's first / last tile has samples
            // outside [0, n_samples): there the signal tile is zeroed, so the carrier needs no range check at all.
            uint32_t ph32[2], st32[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                ph32[h] = (uint32_t)(cphl[h] >> 32);
                st32[h] = (uint32_t)((cdel32[h] + 0x80000000ull) >> 32);
                cphl[h] += 8ull * cdel32[h];                                          // next tile: 256 samples on
            }
            for (int c = 0; c < kTcChunks; ++c, ++qa) {
                const uint32_t buf = qa & 1u, use = qa >> 1;
                // ---- carrier rows of this chunk for the warp's two channels: lane = sample ----
                __syncwarp();                                                       // previous chunk's readers are done
#pragma unroll
                for (int h = 0; h < 2; ++h) {
#pragma unroll
                    for (int half = 0; half < kTcChunk / 32; ++half) {
                        float cr, ci;
                        __sincosf((float)(int32_t)ph32[h] * 1.4629180792671596e-9f, &ci, &cr);        // 2 pi / 2^32
                        ph32[h] += st32[h];
                        const uint32_t keep = (live[h] && !(args.debug & 4)) ? 0xffffffffu : 0u;
                        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(car_w + (uint32_t)(h * kTcChunk + half * 32 + lane) * 8u),
                                     "r"(tf32_rna(cr) & keep), "r"((tf32_rna(ci) ^ 0x80000000u) & keep) : "memory");
                    }
                }
                __syncwarp();
                if (use > 0) bar_wait(A_FREE + 8 * buf, (use - 1) & 1u);              // the MMAs that read this buffer are done
                // ---- tap rows: rows 8 w .. 8 w + 7, four samples per step; one 128-byte core matrix per store ----
                if (tap < L && !(args.debug & 1)) {
                    const uint32_t a_re = s32(sA + (buf * 2 + 0) * kTcAChunk) + (uint32_t)warp * 128u + (uint32_t)lane * 4u;
                    const uint32_t a_im = a_re + kTcAChunk;
                    const uint32_t car_s = car_w + (uint32_t)((r8 >> 2) * kTcChunk + k4) * 8u;
                    // the lane's sixteen replica entries e, e + 4, ..., e + 60 sit in three consecutive words
                    const int e = c * kTcChunk + k4 + args.koff[tap];
                    const uint32_t rep_s = s32(sRep + my_sat * kTcRepWords + (e >> 5));
                    uint32_t w0, w1, w2;
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(rep_s));
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w1) : "r"(rep_s + 4u));
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w2) : "r"(rep_s + 8u));
                    const uint32_t bits_lo = __funnelshift_r(w0, w1, e & 31);         // bit 4 i = sign of entry e + 4 i, i < 8
                    const uint32_t bits_hi = __funnelshift_r(w1, w2, e & 31);         // ... of entry e + 32 + 4 (i - 8)
                    // all sixteen carrier loads first, then the sign flips and stores: volatile asm statements keep their
                    // order, so interleaving them made every store wait for its own load (ncu: 29 % of all stall samples)
                    uint32_t cx[kTcChunk / 4], cy[kTcChunk / 4];
#pragma unroll
                    for (int i = 0; i < kTcChunk / 4; ++i)
                        asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(cx[i]), "=r"(cy[i]) : "r"(car_s + 32u * i));
#pragma unroll
                    for (int i = 0; i < kTcChunk / 4; ++i) {
                        const uint32_t sign = ((i < 8 ? bits_lo : bits_hi) << (31 - 4 * (i & 7))) & 0x80000000u;
                        const uint32_t off = (uint32_t)(i >> 1) * kTcAStep + (uint32_t)(i & 1) * 2048u;
                        asm volatile("st.shared.b32 [%0], %1;" ::"r"(a_re + off), "r"(cx[i] ^ sign) : "memory");
                        asm volatile("st.shared.b32 [%0], %1;" ::"r"(a_im + off), "r"(cy[i] ^ sign) : "memory");
                    }
                }
                if (!(args.debug & 32)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) bar_arrive(A_FULL + 8 * buf);
            }
        }
        u += t_last - t_first;

        // ---- epilogue of the segment: warps 0..3 read the accumulators and publish this CTA's partial of the job ----
        if (warp < 4) {
            bar_wait(ACC, seg & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t vr[32], vi[32];
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
            tmem_ld32(taddr, vr);
            tmem_ld32(taddr + 32, vi);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int row = warp * 32 + lane;
            float *pr = args.partials + ((size_t)(job + (int)blockIdx.x) * 2 * kTcRows + row) * kTcAnts;
            float *pi = pr + (size_t)kTcRows * kTcAnts;
            // columns 0..15 of both accumulators belong to plane 0 of the view, 16..31 to plane 1; plane 0 is the re plane
            // unless the slot's im plane lies lower in memory
            const bool sw = per->swapped != 0;
#pragma unroll
            for (int m = 0; m < kTcAnts; m += 4) {
                float o_re[4], o_im[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float cr_re = __uint_as_float(sw ? vr[16 + m + j] : vr[m + j]);      // W_re x S_re
                    const float cr_im = __uint_as_float(sw ? vr[m + j] : vr[16 + m + j]);      // W_re x S_im
                    const float ci_re = __uint_as_float(sw ? vi[16 + m + j] : vi[m + j]);      // W_im x S_re
                    const float ci_im = __uint_as_float(sw ? vi[m + j] : vi[16 + m + j]);      // W_im x S_im
                    o_re[j] = cr_re - ci_im;
                    o_im[j] = ci_re + cr_im;
                }
                *reinterpret_cast<float4 *>(pr + m) = make_float4(o_re[0], o_re[1], o_re[2], o_re[3]);
                *reinterpret_cast<float4 *>(pi + m) = make_float4(o_im[0], o_im[1], o_im[2], o_im[3]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == kTcGenWarps) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

// sum the partials of every job in CTA order (fixed -> deterministic) and write the caller's [M x L x K x P] planes.
// 8 lanes per output element: lane j adds contributors b_first + j, + 8, ...; a fixed xor tree combines the eight sums.
__global__ void __launch_bounds__(256) tc_finalize_kernel(const TcArgs args, int grid)
{
    const int job = blockIdx.x / 128;
    const int G = args.G, TJ = args.tiles_per_job, K = args.n_sats, L = args.n_taps, M = args.n_ants;
    const int p = job / G, grp = job % G;
    const int b_first = tc_owner((int64_t)job * TJ, grid, args.total_units);
    const int b_last = tc_owner((int64_t)(job + 1) * TJ - 1, grid, args.total_units);
    const int x = (blockIdx.x % 128) * 32 + (threadIdx.x >> 3), j = threadIdx.x & 7;      // element of [re, im][128 rows][16 antennas]
    const size_t stride = (size_t)2 * kTcRows * kTcAnts;
    const float *src = args.partials + (size_t)(job + b_first + j) * stride + x;
    float a0 = 0.f, a1 = 0.f;
    int b = b_first + j;
    for (; b + 8 <= b_last; b += 16, src += 16 * stride) {
        a0 += __ldcg(src);
        a1 += __ldcg(src + 8 * stride);
    }
    if (b <= b_last) a0 += __ldcg(src);
    float acc = a0 + a1;
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    const int c = x / (kTcRows * kTcAnts), row = (x / kTcAnts) % kTcRows, m = x % kTcAnts;
    const int k = grp * kTcSats + (row >> 2), tap = row & 3;
    if (j == 0 && k < K && tap < L && m < M) {
        float *out = c ? args.out_im : args.out_re;
        out[(((size_t)p * K + k) * L + tap) * M + m] = acc;
    }
}

cudaError_t configure_tc_kernel()
{
    return cudaFuncSetAttribute((const void *)correlate_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes);
}

cudaError_t launch_correlate_tc(const TcArgs &args, int grid, int jobs, cudaStream_t stream)
{
    correlate_tc_kernel<<<grid, kTcThreads, kTcSmemBytes, stream>>>(args);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    tc_finalize_kernel<<<jobs * 128, 256, 0, stream>>>(args, grid);
    return cudaGetLastError();
}

}  // namespace gat
