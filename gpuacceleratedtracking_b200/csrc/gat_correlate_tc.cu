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
//   * A operand  = W, 32 channels x 4 tap rows = 128 rows, generated in 64-sample chunks straight into TENSOR MEMORY
//                  (tcgen05.st into a ring of three buffers: lane = row, column = sample): with A in shared memory every
//                  MMA re-read 4 KB of it and the generator's stores and the MMA's reads shared one 128 B/clk pipe.
//                  A thread owns one (channel, tap) row and 16
//                  consecutive samples of the chunk; the four warps of a TMEM lane quarter split the chunk's samples.
//                  Per channel one replica row (sign bits, ballot-packed, double-buffered per tile and shared by the
//                  quarter's four warps through a 128-thread named barrier) and one warp-private carrier row (TF32
//                  cos, -sin) per chunk feed the tap rows, whose entries are load - sign flip - tcgen05.st.
//   * D          = two 128 x 32 FP32 accumulators in TMEM: C_r = W_re x [S_re | S_im], C_i = W_im x [S_re | S_im];
//                  acc_re = C_r.re - C_i.im, acc_im = C_i.re + C_r.im in the epilogue.
// Warp roles: 16 generator warps + 1 control warp (TMA loads two tiles ahead into four signal stages, tcgen05.mma issue from
// uniform registers, tcgen05.commit hand-backs).  Reference semantics followed: the multi-correlator sum of
// /root/reference/src/algorithms.jl:482-516 (replica-buffer form) with Tracking.jl's Int64 code NCO.
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
constexpr int kTcDCols = 2 * kTcCols;                      // TMEM columns of the two accumulators C_r | C_i
constexpr int kTcABufCols = 2 * kTcChunk;                  // TMEM columns of one A buffer: W_re | W_im, one column per sample
constexpr int kTcABufs = 3;                                // A ring: a hand-over round trip (arrive -> MMA -> commit -> wake) costs
                                                           // ~0.85 us, more than a chunk's generation: two buffers serialised them
constexpr int kTcBStages = 4;                              // signal tiles in flight (TMA runs two tiles ahead of the MMA)
constexpr int kTcTmemCols = 512;                           // 64 + 3 x 128 = 448, rounded up to the next power of two
constexpr int kTcLaneSamples = kTcChunk / 4;               // samples per thread and chunk (four warps share a lane quarter)
constexpr int kTcCarStride = kTcLaneSamples * 8 + 16;      // bytes per channel in a warp's carrier rows (+16: bank spread)
constexpr int kTcBGroup = kTcCols * 16;                    // 4 samples of all 32 columns: 512 B
constexpr int kTcBTile = (kTcTile / 4) * kTcBGroup;        // 32 KB
constexpr int kTcGenWarps = 16;
constexpr int kTcThreads = 32 * (kTcGenWarps + 1);
constexpr int kTcTabWords = 320;                           // chip table of a channel as sign bits: up to 10 240 chips (GPS L5) in 1 280 B
constexpr int kTcRepWords = 20;                            // replica sign bits per channel and tile: <= 512 entries (+ one spare word)
constexpr int kTcSmemBytes = kTcBStages * kTcBTile + kTcSats * kTcTabWords * 4 + 2 * kTcSats * kTcRepWords * 4 + kTcGenWarps * 8 * kTcCarStride;

static_assert(kTcDCols + kTcABufs * kTcABufCols <= kTcTmemCols, "accumulators + A ring must fit the TMEM allocation");
static_assert(kTcSmemBytes <= 227 * 1024, "dynamic shared memory beyond the sm_100 per-CTA limit");
static_assert((kTcBStages & (kTcBStages - 1)) == 0 && kTcChunk == 4 * kTcLaneSamples && kTcCarStride % 16 == 0, "layout assumptions");
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
// the MMA thread's wait: one thread, on the critical path of every hand-over -> plain polling, no hardware suspend
__device__ __forceinline__ void bar_wait_spin(uint32_t a, uint32_t parity)
{
    uint32_t done = 0, polls = 0;
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (done) break;
        if (++polls > (1u << 26)) __trap();
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
// D[tmem_d] (+)= A[tmem_a] x B[smem]: A = 128 lanes x 8 columns of TF32 in tensor memory (lane = row, column = k)
__device__ __forceinline__ bool elect_one()
{
    uint32_t e;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(e));
    return e != 0;
}
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(kTcIdesc), "r"(accumulate) : "memory");
}
// this thread's TMEM lane, 16 consecutive columns
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
                   "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// round to nearest (ties away from zero) to TF32 = add half an ulp of the 10-bit mantissa to the magnitude, drop 13 bits.
// cvt.rna.tf32.f32 is not a native instruction on sm_100a: ptxas emits exactly this plus an |x| < inf test, which finite
// samples and sines do not need (an infinite or NaN sample comes out as NaN either way).
__device__ __forceinline__ uint32_t tf32_rna(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }
// a ^ (b & c) in one LOP3
__device__ __forceinline__ uint32_t xor_and(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0x78;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
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
// experiment (GAT_TC_DEBUG bit 4096): SM-clock timestamps of the first 64 chunk hand-overs of CTA 0, read back with gat_debug_tc_trace
constexpr int kTcTraceChunks = 64, kTcTraceKinds = 7;      // kind 6: one-off events of CTA 0 (entry, set-up, first tile, epilogue, exit)
__device__ unsigned long long g_tc_trace[kTcTraceKinds * kTcTraceChunks];
__device__ __forceinline__ void tc_trace(bool on, int kind, uint32_t g)
{
    if (on && g < (uint32_t)kTcTraceChunks) g_tc_trace[kind * kTcTraceChunks + g] = clock64();
}
// The segment barrier of all 17 warps (a counted named barrier; both roles reach it at one call site of the segment loop).
__device__ __forceinline__ void seg_barrier() { asm volatile("bar.sync 5, %0;" ::"n"(kTcThreads) : "memory"); }
__device__ __forceinline__ int tc_owner(int64_t x, int grid, int64_t total) { return (int)(((x + 1) * grid - 1) / total); }

}  // namespace

__global__ void __launch_bounds__(kTcThreads, 1) correlate_tc_kernel(const __grid_constant__ TcArgs args)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *sB = smem;                                            // [kTcBStages][kTcBTile]
    uint32_t *sTab = reinterpret_cast<uint32_t *>(sB + kTcBStages * kTcBTile);    // [32][32] chip tables as sign bits
    uint32_t *sRep = sTab + kTcSats * kTcTabWords;                       // [2][32][20] sign bits of a tile's replicas (tile parity)
    unsigned char *sCar = reinterpret_cast<unsigned char *>(sRep + 2 * kTcSats * kTcRepWords);   // [16 warps][8 channels][kTcCarStride]
    __shared__ uint32_t tmem_base;
    __shared__ __align__(8) uint64_t bars[2 * kTcBStages + 2 * kTcABufs + 1];   // B full x 4; B free x 4; A full x 3; A free x 3; accumulators ready
    // (the warp index through a shuffle from lane 0: ptxas then KNOWS it is warp-uniform, and everything the MMA warp derives
    // from it stays in uniform registers -- with operands it could not prove uniform, every tcgen05.mma was wrapped in an
    // ELECT / R2UR.BROADCAST / BRA.U.ANY loop of ~10 instructions, ~45 cycles per MMA: the path's real bottleneck)
    // the finalize kernel is launched with programmatic stream serialisation: its blocks may take their places now and wait
    // (griddepcontrol.wait) until this grid has completed -- its launch latency leaves the call's critical path
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const bool tr_ev = (args.debug & 4096) && blockIdx.x == 0 && threadIdx.x == 0;
    tc_trace(tr_ev, 6, 0);
    const int tid = threadIdx.x, warp = (int)__reduce_max_sync(0xffffffffu, (unsigned)tid >> 5), lane = tid & 31;
    const uint32_t bar0 = s32(bars);
    const uint32_t B_FULL = bar0, B_FREE = B_FULL + 8 * kTcBStages, A_FULL = B_FREE + 8 * kTcBStages, A_FREE = A_FULL + 8 * kTcABufs, ACC = A_FREE + 8 * kTcABufs;

    if (tid == 0) {
        for (int i = 0; i < kTcBStages; ++i) {
            bar_init(B_FULL + 8 * i, 1);
            bar_init(B_FREE + 8 * i, 1);
        }
        for (int i = 0; i < kTcABufs; ++i) {
            bar_init(A_FULL + 8 * i, kTcGenWarps);
            bar_init(A_FREE + 8 * i, 1);
        }
        bar_init(ACC, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kTcGenWarps) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_base)), "n"(kTcTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    tc_trace(tr_ev, 6, 1);
    const uint32_t tmem = __reduce_max_sync(0xffffffffu, tmem_base);            // columns 0..63: accumulators; then the three A buffers of 128 columns
    const uint32_t tmem_a = tmem + kTcDCols;

    const int K = args.n_sats, G = args.G, TJ = args.tiles_per_job;
    const int64_t TT = args.total_units;
    const int grid = gridDim.x;
    const int64_t r0 = (int64_t)blockIdx.x * TT / grid, r1 = (int64_t)(blockIdx.x + 1) * TT / grid;

    // Roles of a generator warp.  TMEM lane quarter q = warp % 4 (hardware: a warp reaches lanes 32 q .. 32 q + 31 only) =
    // channels 8 q .. 8 q + 7 of the group; sub = warp / 4 = which 16 samples of every 64-sample chunk.
    //   replica bits : the warp generates the tile's sign bits of channels 8 q + 2 sub, + 1 (read by the quarter's four warps)
    //   carrier rows : lane = (channel 8 q + lane / 4, samples 4 (lane % 4) .. + 3 of the warp's 16), warp-private
    //   tap rows     : lane = row 32 q + lane = (channel 8 q + lane / 4, tap lane % 4), the warp's 16 samples
    const int q4 = warp & 3, sub = warp >> 2, kq = lane >> 2, tap = lane & 3;
    const int my_sat = 8 * q4 + kq;
    // ================= control warp: its own tile loop and its own counters =================
    // Everything it touches derives from kernel parameters, blockIdx and two REDUX results, so ptxas keeps it in UNIFORM
    // registers and a tcgen05.mma costs the issuing warp ~3 instructions.  (When this code shared the generator warps'
    // tile loop and counters its operands lived in vector registers and every MMA was wrapped in ELECT / R2UR.BROADCAST /
    // VOTEU sequences of 10+ instructions on a scheduler shared with four generator warps: ~45 cycles per MMA.)
    uint32_t c_qb = 0, c_abuf = 0, c_ause = 0;       // tile counter; A ring position and the use count of its buffers
    uint32_t c_gch = 0;
    const bool c_tr = (args.debug & 4096) && blockIdx.x == 0 && lane == 0 && warp == kTcGenWarps;
    auto load_tile = [&](int64_t un, uint32_t qn) {       // unit un of this CTA's range = its tile number qn
            const int jn = (int)(un / TJ), tn = (int)(un - (int64_t)jn * TJ);
            const uint32_t sn = qn & (kTcBStages - 1);
            bar_expect(B_FULL + 8 * sn, kTcBTile);
            tma_load_4d(s32(sB + sn * kTcBTile), &args.periods[jn / G].map, (args.aligned_start + tn * kTcTile) / 4, B_FULL + 8 * sn);
        };
    if (warp == kTcGenWarps) {
        if (elect_one() && !(args.debug & 16)) {              // the first two tiles are on their way during the segment set-up
            if (r0 < r1) load_tile(r0, 0);
            if (r0 + 1 < r1) load_tile(r0 + 1, 1);
        }
        __syncwarp();
    }
    auto control_segment = [&](int64_t u, int t_first, int t_last) {
        {
            uint32_t qb = c_qb, abuf = c_abuf, ause = c_ause, gch = c_gch;
            const bool tr = c_tr;
            for (int t = t_first; t < t_last; ++t, ++qb) {
                const uint32_t bt = s32(sB + (qb & (kTcBStages - 1)) * kTcBTile);
                if (elect_one() && !(args.debug & 16)) {
                    // the TMA runs two tiles ahead: tile qb + 2 goes into the stage tile qb - 2 was read from
                    const int64_t uc = u + (t - t_first);
                    if (uc + 2 < r1) {
                        // (its own barrier per stage, waited for exactly once per use: a parity wait on an A barrier of two
                        // tiles ago could alias with that buffer's later phases)
                        if (qb >= 2) bar_wait_spin(B_FREE + 8 * ((qb + 2) & (kTcBStages - 1)), ((qb - 2) >> 2) & 1u);
                        load_tile(uc + 2, qb + 2);
                    }
                }
                __syncwarp();
                // B descriptors differ only in their 14-bit start-address field: one base, then 64-bit adds of constants
                const uint64_t db0 = umma_desc(bt, kTcBGroup, 128);
                for (int c = 0; c < kTcChunks; ++c) {
                    bar_wait_spin(A_FULL + 8 * abuf, ause & 1u);
                    tc_trace(tr, 0, gch);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t ta_re = tmem_a + abuf * kTcABufCols, ta_im = ta_re + kTcChunk;
                    const uint64_t db = db0 + (uint64_t)(((uint32_t)c * (kTcChunk / 4) * kTcBGroup) >> 4);
                    const uint32_t acc0 = (t > t_first || c > 0) ? 1u : 0u;
                    if (elect_one()) {
                        if (!(args.debug & 2)) {
#pragma unroll
                            for (int j = 0; j < kTcSteps; ++j) {
                                umma_tf32_ts(tmem, ta_re + 8 * j, db + (uint64_t)((j * 2 * kTcBGroup) >> 4), j ? 1u : acc0);
                                umma_tf32_ts(tmem + kTcCols, ta_im + 8 * j, db + (uint64_t)((j * 2 * kTcBGroup) >> 4), j ? 1u : acc0);
                            }
                        }
                        umma_commit(A_FREE + 8 * abuf);
                        if (c == kTcChunks - 1) {
                            umma_commit(B_FREE + 8 * (qb & (kTcBStages - 1)));
                            if (t == t_last - 1) umma_commit(ACC);
                        }
                    }
                    __syncwarp();
                    tc_trace(tr, 1, gch++);
                    if (++abuf == kTcABufs) { abuf = 0; ++ause; }
                }
            }
            c_qb = qb; c_abuf = abuf; c_ause = ause; c_gch = gch;
        }
    };
    const int koff_tap = args.koff[tap];
    const bool tr = (args.debug & 4096) && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 15);
    const int trk = warp == 0 ? 2 : 4;
    uint32_t gch = 0;
    const bool skeleton = (args.debug & 128) != 0;
    uint32_t abuf = 0, ause = 0;      // A ring position and the use count of its buffers
    uint32_t qb = 0;      // running tile counter of this CTA (B stage + parity, replica buffer)
    uint32_t seg = 0;

    for (int64_t u = r0; u < r1; ++seg) {
        const int job = (int)(u / TJ);
        const int t_first = (int)(u - (int64_t)job * TJ);
        const int t_last = (int)min((int64_t)TJ, (int64_t)t_first + (r1 - u));
        const int p = job / G, grp = job % G;
        const TcPeriod *per = &args.periods[p];

        // ---- segment set-up ----
        uint64_t frac[2] = {0, 0}, ndel[2] = {0, 0};
        uint32_t bmod[2] = {0, 0}, lc[2] = {1, 1};
        int fp[2] = {32, 32};
        uint64_t cph = 0, cd1 = 0;       // carrier phase (Q0.64 cycles) of this lane's first sample in the next chunk; step per sample
        if (warp < kTcGenWarps) {
            const int64_t n0 = (int64_t)args.aligned_start + (int64_t)t_first * kTcTile - args.start_sample;   // may be < 0 (alignment head)
            // Slots past the last channel of a ragged group alias the last channel: their rows are generated like any
            // other (no branches in the generators) and never read by the finalize kernel.  All global loads of the
            // set-up are issued before anything consumes them (it was a chain of six dependent round trips, 3.2 us).
            const SatDev *sdp[2];
            const int8_t *code[2];
            int clen[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k = min(grp * kTcSats + 8 * q4 + 2 * sub + h, K - 1);
                sdp[h] = &args.sats[(size_t)p * K + k];
            }
            const SatDev *sdc = &args.sats[(size_t)p * K + min(grp * kTcSats + my_sat, K - 1)];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                code[h] = sdp[h]->code;
                clen[h] = sdp[h]->code_len;
                ndel[h] = (uint64_t)sdp[h]->nco_delta;
                fp[h] = sdp[h]->nco_fp;
            }
            cd1 = sdc->car_delta;
            cph = sdc->car_phase + (uint64_t)(n0 + sub * kTcLaneSamples + 4 * tap) * cd1;      // 64-bit wrap = whole cycles
            // 16 chips per lane and step (one 16-byte load; the columns are zero-padded to 16 B): their sign bits are
            // squeezed into 16 bits, two lanes make one word -- 2 steps for a 1023-chip code
            const int tab_steps = (max(clen[0], clen[1]) + 511) >> 9;       // 2 for a 1 023-chip code, 20 for GPS L5
            for (int i0 = 0; i0 < tab_steps; i0 += 2) {                      // four loads in flight per round
                uint4 cw[2][2];
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int ci = (i0 + i) * 512 + lane * 16;
                        cw[h][i] = make_uint4(0, 0, 0, 0);
                        if (ci < clen[h]) cw[h][i] = *reinterpret_cast<const uint4 *>(code[h] + ci);
                    }
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const uint4 w = cw[h][i];
                        auto squeeze = [](uint32_t x) { return ((x >> 7) & 1u) | ((x >> 14) & 2u) | ((x >> 21) & 4u) | ((x >> 28) & 8u); };
                        const uint32_t half = squeeze(w.x) | (squeeze(w.y) << 4) | (squeeze(w.z) << 8) | (squeeze(w.w) << 12);
                        const uint32_t other = __shfl_xor_sync(0xffffffffu, half, 1);
                        const int ci = (i0 + i) * 512 + lane * 16;
                        if (!(lane & 1) && ci < kTcTabWords * 32)
                            sTab[(8 * q4 + 2 * sub + h) * kTcTabWords + (ci >> 5)] = half | (other << 16);
                    }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                lc[h] = (uint32_t)clen[h];
                // state at the first sample of tile t_first
                const __int128 tot = (__int128)(n0 + args.shift0) * (__int128)(int64_t)ndel[h] + (__int128)sdp[h]->nco_start;
                int64_t b = (int64_t)(tot >> fp[h]) % clen[h];
                if (b < 0) b += clen[h];
                bmod[h] = (uint32_t)b;
                frac[h] = (uint64_t)tot & ((1ull << fp[h]) - 1ull);
            }
        }
        seg_barrier();     // all 17 warps, one call site: tables in place; previous segment's epilogue done (TMEM free)
        if (warp == kTcGenWarps) {
            control_segment(u, t_first, t_last);
            u += t_last - t_first;
            continue;
        }
        tc_trace(tr_ev, 6, 2 + 8 * seg);

        for (int t = t_first; t < t_last; ++t, ++qb) {
            const uint32_t st = qb & (kTcBStages - 1);
            const int n0 = args.aligned_start + t * kTcTile - args.start_sample;      // relative index of tile sample 0
            // ================================= generator warps =================================
            // ---- replica sign bits of this tile for the warp's two channels: entry e <-> sample n0 + e + shift0.  Both
            // channels' table lookups are in flight together (four independent load -> ballot chains per round) ----
            uint32_t *rep_t = sRep + (qb & 1u) * (kTcSats * kTcRepWords);
            if (!(args.debug & 64)) {
                const int rows = (kTcTile + args.span + 31) >> 5;
                uint64_t v[2], v32[2];
                uint32_t tab_s[2];
                int sh[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    v[h] = frac[h] + (uint64_t)lane * ndel[h];
                    v32[h] = 32ull * ndel[h];
                    tab_s[h] = s32(sTab + (8 * q4 + 2 * sub + h) * kTcTabWords);
                    sh[h] = fp[h] - 32;
                }
                if (args.win_ok) {
                    // Oversampled signals (the whole replica of a tile spans < 32 chips -- host-checked): one 32-chip window
                    // per channel and tile, bit j = chip (first chip + j) mod length, fetched by one table lookup; an entry's
                    // sign is then bit (chips advanced since the tile's first entry) of the window -- no wrap, no table
                    // address, no load per row.  Same Int64 NCO, same bits as the general path below.
                    uint32_t win[2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        uint32_t idx = bmod[h] + (uint32_t)lane, word;
                        idx = min(idx, idx - lc[h]);
                        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(word) : "r"(tab_s[h] + 4u * (idx >> 5)));
                        win[h] = __ballot_sync(0xffffffffu, ((word >> (idx & 31u)) & 1u) != 0u);
                    }
                    for (int r = 0; r < rows; r += 2) {
#pragma unroll
                        for (int j = 0; j < 2; ++j)
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const uint32_t rel = (uint32_t)(v[h] >> 32) >> sh[h];
                                const uint32_t bits = __ballot_sync(0xffffffffu, ((win[h] >> rel) & 1u) != 0u);
                                if (lane == 2 * j + h) rep_t[(8 * q4 + 2 * sub + h) * kTcRepWords + r + j] = bits;
                                v[h] += v32[h];
                            }
                    }
                } else
                for (int r = 0; r < rows; r += 2) {                 // (the row count is rounded up to even: 20 words per channel)
                    uint32_t word[2][2], sft[2][2];
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            uint32_t idx = bmod[h] + ((uint32_t)(v[h] >> 32) >> sh[h]);
                            idx = min(idx, idx - lc[h]);            // single wrap (host-checked)
                            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(word[j][h]) : "r"(tab_s[h] + 4u * (idx >> 5)));
                            sft[j][h] = idx & 31u;
                            v[h] += v32[h];
                        }
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const uint32_t bits = __ballot_sync(0xffffffffu, ((word[j][h] >> sft[j][h]) & 1u) != 0u);
                            if (lane == 2 * j + h) rep_t[(8 * q4 + 2 * sub + h) * kTcRepWords + r + j] = bits;
                        }
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {                       // advance the NCO base to the next tile
                    const unsigned __int128 nf = (unsigned __int128)frac[h] + (unsigned __int128)kTcTile * (unsigned __int128)ndel[h];
                    uint64_t x = (uint64_t)bmod[h] + (uint64_t)(nf >> fp[h]);
                    if (x >= lc[h]) x %= lc[h];
                    bmod[h] = (uint32_t)x;
                    frac[h] = (uint64_t)nf & ((1ull << fp[h]) - 1ull);
                }
            }
            // the quarter's four warps exchange their replica rows.  The buffer alternates per tile: a warp that runs ahead
            // into tile t + 1 writes the other buffer, and it cannot reach tile t + 2 before everyone has left tile t.
            asm volatile("bar.sync %0, 128;" ::"r"(1 + q4) : "memory");
            if (args.dump && lane < kTcRepWords) {
                // debug (gat_debug_tc_replica_bits): this tile's sign-bit rows of the warp's two channels, as the tap rows will read them
                const int64_t ug = (int64_t)job * TJ + t;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int sat = 8 * q4 + 2 * sub + h;
                    args.dump[(ug * kTcSats + sat) * kTcRepWords + lane] = rep_t[sat * kTcRepWords + lane];
                }
            }
            if (t == t_first) tc_trace(tr_ev, 6, 3 + 8 * seg);

            // ---- the signal tile: wait for the TMA, round to TF32, zero what lies outside [0, n_samples) ----
            if (!(args.debug & 16)) bar_wait(B_FULL + 8 * st, (qb >> 2) & 1u);
            {
                unsigned char *bp = sB + st * kTcBTile;
                const bool edge = n0 < 0 || n0 + kTcTile > args.n_samples;            // only a job's first / last tile
                for (int i = tid; i < kTcBTile / 16; i += 32 * kTcGenWarps) {        // 16 B = 4 samples of one column
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
            // Only a job's first / last tile has samples outside [0, n_samples): there the signal tile is zeroed, so neither
            // the carrier nor the replica needs a range check.  Rows of taps or channels that do not exist hold finite
            // junk: a row of A only reaches its own row of D, which the finalize kernel never reads.
            if (t == t_first) tc_trace(tr_ev, 6, 4 + 8 * seg);
            const uint32_t car_w = s32(sCar + warp * (8 * kTcCarStride));             // this warp's carrier rows
            const uint32_t car_st = car_w + (uint32_t)(kq * kTcCarStride + tap * 32);  // writer: 4 samples = 32 B
            const uint32_t car_ld = car_w + (uint32_t)(kq * kTcCarStride);             // reader: the channel's 16 samples
            const uint32_t rep_s0 = s32(rep_t + my_sat * kTcRepWords);
            const int e_lane = sub * kTcLaneSamples + koff_tap;
            const uint32_t t_row = tmem_a + ((uint32_t)(32 * q4) << 16) + (uint32_t)(sub * kTcLaneSamples);
            for (int c = 0; c < kTcChunks; ++c) {
                const uint32_t buf = abuf, use = ause;
                if (++abuf == kTcABufs) { abuf = 0; ++ause; }
                if (skeleton) {                                                     // experiment: the hand-over skeleton alone
                    if (use > 0) bar_wait(A_FREE + 8 * buf, (use - 1) & 1u);
                    tc_trace(tr, trk, gch);
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    if (c == 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    tc_trace(tr, trk + 1, gch++);
                    if (lane == 0) bar_arrive(A_FULL + 8 * buf);
                    continue;
                }
                // ---- carrier rows of this chunk: four consecutive samples of one channel per lane ----
                __syncwarp();                                                       // previous chunk's readers are done
                {
                    uint32_t cw[8];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint32_t ph = (uint32_t)((cph + (uint64_t)i * cd1) >> 32);
                        float cr, ci;
                        __sincosf((float)(int32_t)ph * 1.4629180792671596e-9f, &ci, &cr);        // 2 pi / 2^32
                        cw[2 * i] = tf32_rna(cr);
                        cw[2 * i + 1] = tf32_rna(ci) ^ 0x80000000u;
                    }
                    cph += (uint64_t)kTcChunk * cd1;
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(car_st), "r"(cw[0]), "r"(cw[1]), "r"(cw[2]), "r"(cw[3]) : "memory");
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(car_st + 16u), "r"(cw[4]), "r"(cw[5]), "r"(cw[6]), "r"(cw[7]) : "memory");
                }
                __syncwarp();
                // ---- tap rows: this lane's row, the warp's 16 samples of the chunk ----
                // the row's sixteen replica entries e .. e + 15 sit in two consecutive words
                const int e = c * kTcChunk + e_lane;
                uint32_t w0, w1;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(rep_s0 + 4u * (uint32_t)(e >> 5)));
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w1) : "r"(rep_s0 + 4u * (uint32_t)(e >> 5) + 4u));
                uint32_t cx[kTcLaneSamples], cy[kTcLaneSamples];
#pragma unroll
                for (int i = 0; i < kTcLaneSamples / 2; ++i)
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(cx[2 * i]), "=r"(cy[2 * i]), "=r"(cx[2 * i + 1]), "=r"(cy[2 * i + 1])
                                 : "r"(car_ld + 16u * i));
                const uint32_t bits = __funnelshift_r(w0, w1, e & 31);              // bit i = sign of entry e + i
#pragma unroll
                for (int i = 0; i < kTcLaneSamples; ++i) {
                    const uint32_t sh = bits << (31 - i);                           // bit 31 = sign of sample i
                    cx[i] = xor_and(cx[i], sh, 0x80000000u);
                    cy[i] = xor_and(cy[i], sh, 0x80000000u);
                }
                if (use > 0) bar_wait(A_FREE + 8 * buf, (use - 1) & 1u);              // the MMAs that read this buffer are done
                tc_trace(tr, trk, gch);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                tmem_st16(t_row + buf * kTcABufCols, cx);
                tmem_st16(t_row + buf * kTcABufCols + kTcChunk, cy);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                if (c == 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the rounded signal tile -> the MMA's proxy
                __syncwarp();
                tc_trace(tr, trk + 1, gch++);
                if (lane == 0) bar_arrive(A_FULL + 8 * buf);
            }
        }
        u += t_last - t_first;
        tc_trace(tr_ev, 6, 5 + 8 * seg);

        // ---- epilogue of the segment: warps 0..3 read the accumulators and publish this CTA's partial of the job ----
        if (warp < 4) {
            bar_wait(ACC, seg & 1u);
            tc_trace(tr_ev, 6, 6 + 8 * seg);
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
            tc_trace(tr_ev, 6, 7 + 8 * seg);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    tc_trace(tr_ev, 6, 40);
    if (warp == kTcGenWarps) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTcTmemCols) : "memory");
}

// sum the partials of every job in CTA order (fixed -> deterministic) and write the caller's [M x L x K x P] planes.
// 8 lanes per output element: lane j adds contributors b_first + j, + 8, ...; a fixed xor tree combines the eight sums.
__global__ void __launch_bounds__(256) tc_finalize_kernel(const TcArgs args, int grid)
{
    asm volatile("griddepcontrol.wait;" ::: "memory");          // the correlate grid has completed, its partials are visible
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

}  // namespace gat
extern "C" int gat_debug_tc_trace(unsigned long long *out, int n)       // experiment read-back, not part of include/gat.h
{
    if (n > gat::kTcTraceKinds * gat::kTcTraceChunks) n = gat::kTcTraceKinds * gat::kTcTraceChunks;
    return (int)cudaMemcpyFromSymbol(out, gat::g_tc_trace, sizeof(unsigned long long) * n);
}
namespace gat {

cudaError_t configure_tc_kernel()
{
    return cudaFuncSetAttribute((const void *)correlate_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes);
}

cudaError_t launch_correlate_tc(const TcArgs &args, int grid, int jobs, cudaStream_t stream)
{
    correlate_tc_kernel<<<grid, kTcThreads, kTcSmemBytes, stream>>>(args);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(jobs * 128);
    cfg.blockDim = dim3(256);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, tc_finalize_kernel, args, grid);
}

}  // namespace gat
