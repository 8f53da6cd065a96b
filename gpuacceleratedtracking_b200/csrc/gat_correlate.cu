// gat_correlate.cu -- the fused downconvert-and-correlate kernel family for sm_100a.
//
// Replaces, in ONE launch, what the reference spreads over 1-4 launches
// (/root/reference/src/algorithms.jl:869-1545):
//   code replica generation      (src/algorithms.jl:13-140, inline :179-186)
//   carrier replica generation   (CUDA.sincos sites, e.g. :172, :484, :573)
//   carrier wipe-off             (:175-176, :488-489)
//   code wipe-off + accumulate   (:185-186, :494-495)
//   in-block + cross-block reduce (:196-208, :521-533, :625-632, src/reduction.jl)
//
// Design (DESIGN.md has the long form):
//   * persistent CTAs, one per SM; the flattened (period, sat-group, tile) space is split
//     evenly over the grid ("stream-K"), so every SM streams the same number of bytes.
//   * warp W is a producer: it moves [tile x antennas] signal tiles HBM -> smem with 2-D tiled
//     TMA loads (UTMALDG) through a full/empty mbarrier ring and keeps the chip table of every
//     satellite batched on the CTA in shared memory.
//   * warps 0..W-1 are consumers.  A consumer warp owns (satellite s, antenna group ag,
//     sample slice sl); all of them read the SAME staged signal tile, so one HBM/L2 read
//     feeds every satellite on the SM.
//   * carrier: 64-bit integer phase accumulator (exact wrap), top 32 bits -> FP32 -> MUFU
//     sin/cos.  code: each warp generates its tile's code replica (integer NCO or IEEE-double
//     formula -> smem chip table) into a private smem buffer while the signal tile is in flight;
//     every tap is then one shared-memory load.
//   * wipe-off and taps are packed FP32x2 FMAs (FFMA2), two antennas per instruction,
//     accumulators in registers.
//   * reduction: in-warp halving butterfly (reduce-scatter by shuffles) -> smem across
//     slices -> per-CTA partial -> last-arriving CTA of a job sums partials in fixed order
//     (single pass, deterministic, no float atomics, self-cleaning counters).
#include "gat_internal.h"
#include <cstdlib>

#include <cstdio>

namespace gat {

// --------------------------------------------------------------------------------------
// PTX helpers
// --------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_inval(uint64_t *bar)
{
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_s(uint32_t bar_s)       // 32-bit shared address
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_s) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ uint64_t global_timer_ns()
{
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Watchdog: a wait that lasts seconds means a broken launch (a descriptor whose box does not match the
// expected byte count, a lost arrival).  Trap -> the host sees a CUDA error instead of a hung device.
// Only the retry path pays for it.
constexpr uint64_t kWaitLimitNs = 4000000000ull;
__device__ __forceinline__ void mbar_wait_s(uint32_t addr, uint32_t parity)
{
    uint32_t done = 0;
    uint64_t t0 = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) break;
        const uint64_t now = global_timer_ns();
        if (t0 == 0) t0 = now;
        else if (now - t0 > kWaitLimitNs) __trap();
    }
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) { mbar_wait_s(smem_u32(bar), parity); }
// The producer warp waits for whole tiles to be consumed (microseconds): back off between tries so that its
// retry loop does not take issue slots from the consumer warps of its scheduler (measured: 7 % of all executed
// instructions were this loop spinning).
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t *bar, uint32_t parity)
{
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    uint64_t t0 = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity), "r"(2000u)
            : "memory");
        if (done) break;
        __nanosleep(200);
        const uint64_t now = global_timer_ns();
        if (t0 == 0) t0 = now;
        else if (now - t0 > kWaitLimitNs) __trap();
    }
}
// 1-D bulk async copy global -> shared, completion counted in bytes on an mbarrier (TMA unit).
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// 2-D tiled TMA load: one instruction moves a [box_rows x box_cols] tile (UTMALDG)
__device__ __forceinline__ void tma_load_2d(void *dst_smem, const CUtensorMap *map, int c0, int c1, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst_smem)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
// named barrier among the warps that share one code replica (ids 2..): id and count are warp-uniform
__device__ __forceinline__ void group_bar_sync(int id, int threads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ float chip_to_float(int c)   // +-1 int8 -> +-1.0f without a conversion instruction
{
    return __uint_as_float(0x3f800000u | ((uint32_t)c & 0x80000000u));
}
// CTA-wide barrier for code that exists in TWO copies (the reallocation class instantiates the kernel body once per warpgroup kind):
// __syncthreads() is `barrier.sync.aligned`, which promises that every thread of the CTA executes the SAME instruction; the two
// copies meet at different ones, so they use the non-aligned form on the same barrier resource (synccheck flags the aligned one).
template <int ROLE>
__device__ __forceinline__ void cta_sync()
{
    if constexpr (ROLE == 0) __syncthreads();
    else asm volatile("barrier.sync 0;" ::: "memory");
}
__device__ __forceinline__ void consumer_bar_sync(int threads)
{
    asm volatile("bar.sync 1, %0;" ::"r"(threads) : "memory");
}

__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// debug timeline: one thread per role stamps a slot (see gat_get_timeline in include/gat.h)
#define GAT_STAMP(slot)                                                                        \
    do {                                                                                       \
        if (args.timeline && lane == 0 && (warp == 0 || warp == PW))                            \
            args.timeline[(size_t)blockIdx.x * 16 + (slot)] = globaltimer_ns();                \
    } while (0)

// Shared-memory load that ptxas may not move across loop iterations: without it the assembler
// software-pipelines the sample loop and pays ~20 register-rotation moves per iteration (IMAD.MOV on
// the FMA pipe) -- measured in the SASS of the 11-tap kernel.
__device__ __forceinline__ float lds_f32(const float *p)
{
    float v;
    asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(v) : "r"(smem_u32(p)));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(const float *p)
{
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)));
    return v;
}
// int16 -> float.  GAT_S16_MODE 0 (default): sign-extending byte permute / arithmetic shift, then I2FP (ALU pipe).
// GAT_S16_MODE 1: offset-binary magic-number conversion (LOP3 + PRMT under the exponent of 2^23, then one packed
// FADD removes the bias for two antennas).  Measured on B200, 256 periods x 16 antennas x 3 taps: mode 0 265 us,
// mode 1 277 us, (float)(short) -> I2F.S16 265 us: the loop is issue-bound, so the variant with the fewest
// FMA-pipe instructions wins; kept for the record.
#ifndef GAT_S16_MODE
#define GAT_S16_MODE 0
#endif
[[maybe_unused]] constexpr float kS16Bias = 8388608.f + 32768.f;
__device__ __forceinline__ uint32_t s16_flip(uint32_t w) { return w ^ 0x80008000u; }
__device__ __forceinline__ float s16_lo_biased(uint32_t wf) { return __uint_as_float(__byte_perm(wf, 0x4B000000u, 0x7410)); }   // I
__device__ __forceinline__ float s16_hi_biased(uint32_t wf) { return __uint_as_float(__byte_perm(wf, 0x4B000000u, 0x7432)); }   // Q
__device__ __forceinline__ float s16_lo(uint32_t w)   // sign-extending permute (selector msb = replicate the byte's sign) + I2FP
{
    int v;
    asm("prmt.b32 %0, %1, 0, 0x9910;" : "=r"(v) : "r"(w));
    return (float)v;
}
__device__ __forceinline__ float s16_hi(uint32_t w) { return (float)((int)w >> 16); }
__device__ __forceinline__ uint32_t lds_u32_at(uint32_t smem_addr)
{
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_addr));
    return v;
}
__device__ __forceinline__ int lds_s8_at(uint32_t smem_addr)      // chip table entry, sign-extended
{
    int v;
    asm volatile("ld.shared.s8 %0, [%1];" : "=r"(v) : "r"(smem_addr));
    return v;
}
__device__ __forceinline__ void sts_b32_at(uint32_t smem_addr, uint32_t v)
{
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(smem_addr), "r"(v) : "memory");
}
__device__ __forceinline__ float lds_f32_at(uint32_t smem_addr)
{
    float v;
    asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(v) : "r"(smem_addr));
    return v;
}

#ifndef GAT_FULL_UNROLL
#define GAT_FULL_UNROLL 1          // 1: straight-line code for full tiles in the reallocation class (7 / 9 / 11 taps); 2: in every class
#endif
#ifndef GAT_LOOP_UNROLL
#define GAT_LOOP_UNROLL 1          // A/B builds: unroll factor of the sample loop
#endif
constexpr int kLoopUnroll = GAT_LOOP_UNROLL;
typedef unsigned long long f32x2;  // two packed floats: lo = even antenna, hi = odd antenna
__device__ __forceinline__ f32x2 pack2(float lo, float hi)
{
    f32x2 d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c)
{
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b)
{
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b)
{
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// --------------------------------------------------------------------------------------
// replica phase arithmetic
// --------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t floormod64(int64_t a, int64_t m)
{
    int64_t r = a % m;
    return r < 0 ? r + m : r;
}

// Integer NCO, Tracking.jl gen_code_replica! [upstream]:
//   idx(k) = (k * delta + start) >> fp         k = sample offset + tap shift (may be < 0)
// Evaluated per tile as   base_idx + ((frac0 + kk * delta) >> fp),   kk >= 0 relative to the
// tile's first (sample, latest tap): exact because the split is at a multiple of 2^fp.
__device__ __forceinline__ void nco_tile_base(const SatDev &sd, int64_t u0, uint64_t &frac, uint32_t &bmod)
{
    const __int128 tot = (__int128)u0 * (__int128)sd.nco_delta + (__int128)sd.nco_start;
    const int64_t base = (int64_t)(tot >> sd.nco_fp);
    frac = (uint64_t)tot & ((1ull << sd.nco_fp) - 1ull);
    bmod = (uint32_t)floormod64(base, sd.code_len);
}
// chip-table index of replica entry u (u = sample offset in the tile + tap offset from the latest tap).
// host guarantees (tile_len + span + 1) * delta + 2^fp < 2^64, so v = frac + u * delta is exact in 64 bits;
// fp >= 32 always, so only the high word is shifted (sh = fp - 32).
__device__ __forceinline__ uint32_t rep_index_nco(uint64_t v, int sh, uint32_t bmod, uint32_t lc)
{
    uint32_t idx = bmod + ((uint32_t)(v >> 32) >> sh);
    if (idx >= lc) {
        idx -= lc;
        if (idx >= lc) idx %= lc;
    }
    return idx;
}
// IEEE-double form of the reference GPU kernels (src/algorithms.jl:179-182):
//   floor(code_frequency / sampling_frequency * (n + shift) + start_code_phase)
// separate multiply and add (Julia does not contract), then floor.
__device__ __forceinline__ int32_t f64_chip_floor(double ratio, double phase, int32_t u)
{
    const double cp = __dadd_rn(__dmul_rn(ratio, (double)u), phase);
    return __double2int_rd(cp);
}
__device__ __forceinline__ uint32_t rep_index_f64(double ratio, double phase, int32_t u_abs, int32_t b, uint32_t bmod, uint32_t lc)
{
    uint32_t idx = bmod + (uint32_t)(f64_chip_floor(ratio, phase, u_abs) - b);   // floor() is monotone: >= 0
    if (idx >= lc) {
        idx -= lc;
        if (idx >= lc) idx %= lc;
    }
    return idx;
}

// --------------------------------------------------------------------------------------
// in-warp halving butterfly: N values per lane in, N/32 fully reduced values per lane out;
// lane l ends up owning elements [l*N/32, (l+1)*N/32).
// --------------------------------------------------------------------------------------
template <int N, int MASK>
__device__ __forceinline__ void reduce_scatter(float *v, int lane)
{
    const bool up = (lane & MASK) != 0;
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
        const float send = up ? v[i] : v[i + N / 2];
        const float keep = up ? v[i + N / 2] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, MASK);
    }
    if constexpr (MASK > 1) reduce_scatter<N / 2, MASK / 2>(v, lane);
}

// --------------------------------------------------------------------------------------
// rows [row0, row1) of one tile's code replica: rep[u] = +-1.0f chip under (tile sample 0 + latest tap + u), 32 entries
// per row.  Called by the consumer warps that share a replica (each its share of the rows) or by the replica warp
// (all rows).  NCO mode: (frac, bmod) = phase state under entry 0; F64 mode: u0 = absolute index of entry 0.
// --------------------------------------------------------------------------------------
// DUMP layout: [tile][dump_stride] -- entry u of a tile's replica is the chip under (tile sample 0 + latest tap + u).  A visit
// of two tiles generates ONE replica of 2 * tile_len + span entries: entries below dump_stride belong to the first tile's slot,
// entries from tile_len on to the second tile's (shifted by tile_len); the overlap is written to both.
struct DumpDst {
    uint32_t *p;        // first tile's slot
    int32_t stride;     // entries per tile slot
    int32_t second_at;  // first entry that (also) belongs to the second tile's slot; INT_MAX for a one-tile visit
    __device__ __forceinline__ void put(int u, uint32_t idx) const
    {
        if (u < stride) p[u] = idx;
        if (u >= second_at && u - second_at < stride) p[stride + (u - second_at)] = idx;
    }
};
template <bool F64, bool DUMP>
__device__ __forceinline__ void gen_replica_rows(const CorrArgs &args, int lane, int row0, int row1, float *rep, uint32_t rep_s,
                                                 const int8_t *tab, uint32_t tab_s, uint64_t frac, uint32_t bmod, uint64_t delta, int sh,
                                                 uint32_t lc, double ratio, double cphase, int32_t u0, [[maybe_unused]] const DumpDst &dmp)
{
    if constexpr (F64) {
        const int32_t b = f64_chip_floor(ratio, cphase, u0);
        bmod = (uint32_t)floormod64(b, lc);
        int r = row0;
        for (; r + 1 < row1; r += 2) {
            const uint32_t i0 = rep_index_f64(ratio, cphase, u0 + r * 32 + lane, b, bmod, lc);
            const uint32_t i1 = rep_index_f64(ratio, cphase, u0 + r * 32 + 32 + lane, b, bmod, lc);
            const int c0 = tab[i0];
            const int c1 = tab[i1];
            rep[r * 32 + lane] = chip_to_float(c0);
            rep[r * 32 + 32 + lane] = chip_to_float(c1);
            if constexpr (DUMP) {
                dmp.put(r * 32 + lane, i0);
                dmp.put(r * 32 + 32 + lane, i1);
            }
        }
        if (r < row1) {
            const uint32_t i0 = rep_index_f64(ratio, cphase, u0 + r * 32 + lane, b, bmod, lc);
            rep[r * 32 + lane] = chip_to_float(tab[i0]);
            if constexpr (DUMP) dmp.put(r * 32 + lane, i0);
        }
    } else {
        uint64_t v = frac + (uint64_t)(uint32_t)(row0 * 32 + lane) * delta;
        const uint64_t v32 = 32ull * delta;
        int r = row0;
        if (args.rep_single_wrap) {
            // the tile advances the code by less than one period (host-checked): one branch-free wrap per
            // entry, table and replica addressed through 32-bit shared-memory addresses
            uint32_t wa = rep_s + 4u * (uint32_t)(row0 * 32 + lane);
            for (; r + 3 < row1; r += 4, wa += 512u) {   // 4 independent table lookups in flight per lane
                int c[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t idx = bmod + ((uint32_t)(v >> 32) >> sh);
                    c[j] = lds_s8_at(tab_s + min(idx, idx - lc));   // unsigned: idx - lc wraps high when idx < lc
                    if constexpr (DUMP) dmp.put((r + j) * 32 + lane, min(idx, idx - lc));
                    v += v32;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) sts_b32_at(wa + 128u * j, 0x3f800000u | ((uint32_t)c[j] & 0x80000000u));
            }
            for (; r < row1; ++r, v += v32, wa += 128u) {
                const uint32_t idx = bmod + ((uint32_t)(v >> 32) >> sh);
                sts_b32_at(wa, 0x3f800000u | ((uint32_t)lds_s8_at(tab_s + min(idx, idx - lc)) & 0x80000000u));
                if constexpr (DUMP) dmp.put(r * 32 + lane, min(idx, idx - lc));
            }
        } else {
            for (; r + 3 < row1; r += 4) {
                int c[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t idx = rep_index_nco(v, sh, bmod, lc);
                    c[j] = tab[idx];
                    if constexpr (DUMP) dmp.put((r + j) * 32 + lane, idx);
                    v += v32;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) rep[(r + j) * 32 + lane] = chip_to_float(c[j]);
            }
            for (; r < row1; ++r, v += v32) {
                const uint32_t idx = rep_index_nco(v, sh, bmod, lc);
                rep[r * 32 + lane] = chip_to_float(tab[idx]);
                if constexpr (DUMP) dmp.put(r * 32 + lane, idx);
            }
        }
    }
}

// stream-K ownership: CTA b owns global tiles [b*TT/grid, (b+1)*TT/grid)
__device__ __forceinline__ int tile_owner(int64_t x, int grid, int64_t total)
{
    return (int)(((x + 1) * grid - 1) / total);
}

// accumulator slot x of a job -> output element.  Slot layout per role r = ag*S + s:
//   A >= 2: e = ((antenna_pair * L + tap) * 2 + {re,im}) * 2 + {even,odd antenna};  A == 1: e = tap*2 + {re,im}
// RES (resident kernel): the element goes to HOST memory as one 8-byte word {value bits, command sequence number} -- the host
// knows an element has arrived when its sequence number matches, so a command needs no completion fence and no flag
// (a system-scope fence after sysmem stores cost 2 - 11 us per CTA in the first version of the resident kernel).
template <int A, int L, bool RES = false>
__device__ __forceinline__ void emit_output(const CorrArgs &args, int job, int x, float val, [[maybe_unused]] uint32_t res_seq = 0u)
{
    constexpr int R = 2 * A * L;
    constexpr int RP = (R + 31) / 32 * 32;
    const int r = x / RP, e = x % RP;
    if (e >= R) return;
    const int S = args.S, K = args.n_sats, M = args.n_ants;
    const int p = job / args.G, grp = job % args.G;
    const int s2 = r % S, ag2 = (r / S) % args.AG, tg2 = r / (S * args.AG);
    int ml, l, c;
    if constexpr (A >= 2) {
        ml = 2 * ((e >> 2) / L) + (e & 1);
        l = (e >> 2) % L;
        c = (e >> 1) & 1;
    } else {
        ml = 0;
        l = e >> 1;
        c = e & 1;
    }
    const int kk = grp * S + s2, m = ag2 * A + ml;
    l += tg2 * L;                      // tap groups: this role holds taps tg2 * L .. of the call's n_taps
    if (kk >= K || m >= M || l >= args.n_taps) return;
    const size_t idx = (((size_t)p * K + kk) * args.n_taps + l) * M + m;
    if constexpr (RES) {
        uint2 *dst2 = reinterpret_cast<uint2 *>(c ? args.out_im : args.out_re) + idx;
        asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(dst2), "r"(__float_as_uint(val * args.out_scale)), "r"(res_seq) : "memory");
        return;
    }
    float *dst = (c ? args.out_im : args.out_re) + idx;
    val *= args.out_scale;             // raw integer tiles: the (power-of-two) sample scale is applied once, here
    if (args.flags & 1u) val += *dst;  // GAT_ACCUMULATE
    *dst = val;
    if (args.n_peers > 1) {
        // fused gather: the same value goes into slice `my_rank` of every other rank's buffer
        // (posted stores over NVLink to CUDA-IPC peer mappings; no collective launch)
        const size_t off = (size_t)args.my_rank * args.gather_elems + args.gather_off + idx;
        for (int d = 0; d < args.n_peers; ++d)
            if (d != args.my_rank) (c ? args.peer_im[d] : args.peer_re[d])[off] = val;
    }
}

// --------------------------------------------------------------------------------------
// the kernel
// --------------------------------------------------------------------------------------
// SC16: the staged tile holds raw interleaved complex int16 samples (one 32-bit word = I | Q << 16 per
// sample and antenna, a single plane) instead of two FP32 planes; they are converted in registers.
// DUMP: debug instantiation that also writes the chip-table index of every replica entry it generates to args.dump --
// the bit-exactness tests read the HOT kernel's own index arithmetic (tile-to-tile NCO advance, both wrap branches, the
// Float64 mode), not a look-alike.  Used with one period and one channel.
// HELP: one more warp (index W + 1, the "replica warp") generates the code replica of every tile into a double-buffered ring, one
// or two tiles ahead of the consumers, which then neither carry the code-NCO state nor meet at named barriers twice per tile.
// For shapes with few satellites per CTA (one channel per block: 4 consumer warps share a tile, so the per-tile work of a
// warp -- replica rows, two group barriers, bookkeeping -- was as long as its 8 FMA iterations: ncu source view of C4, only
// 50 % of the warp samples inside the FMA loop).
// Visits (reallocation class): the CTA's tiles q = 0, 1, .. go to the sample slices in runs of V consecutive tiles (V = 1 or 2),
// run i to slice i % SL; a consumer warp works through a whole run per visit.  A segment (the CTA's share of one job) may
// cut a run: then each piece is a visit of its own.  first_visit() returns the offset (tiles from the segment's first) of the
// RUN that holds slice sl's first visit in a segment starting at CTA tile q0 (-1 ... : a run may have begun in the previous
// segment); the visit itself covers [max(o_run, 0), min(o_run + V, n_seg)), following runs start V * SL tiles apart.
__device__ __forceinline__ int first_run_offset(uint32_t q0, int sl, int V, int SL)
{
    const int PV = V * SL;
    int o_run = V * sl - (int)(q0 % (uint32_t)PV);
    if (o_run <= -V) o_run += PV;
    return o_run;
}
// ring stage / parity code (2 * stage + parity) advanced by n <= stages tiles
__device__ __forceinline__ int sp_advance(int sp, int n, int stages)
{
    sp += 2 * n;
    if (sp >= 2 * stages) sp = (sp - 2 * stages) ^ 1;
    return sp;
}

// ROLE (register-reallocation class only): the kernel body is instantiated once per warpgroup kind, each copy behind its own
// setmaxnreg, so that ptxas allocates the consumer code against 160 registers and the producer / replica-warp code against
// 32 (one copy with a join after the setmaxnreg made every value that lives across it spill).  0 = all roles in one copy.
enum { kRoleAll = 0, kRoleAux = 1, kRoleConsumer = 2 };
// RES (resident kernel, gat_resident.cu): the body runs once per command inside a kernel that stays on the device; what changes
// from call to call -- the block's descriptors, the channel records, the grid barrier's target and the completion sequence
// number -- comes from `ro` instead of the kernel arguments.
struct ResOverride {
    const PeriodDev *periods;     // descriptors of the selected slot (device global memory)
    const SatDev *sats;           // this command's channel records (shared memory)
    unsigned int barrier_target;
    uint32_t seq;
    uint32_t reinit;              // not the launch's first command: the barriers hold the previous command's objects
};
template <int A, int L, bool F64, bool SC16, bool DUMP, bool HELP, int ROLE, bool RES = false>
__device__ __forceinline__ void correlate_body(const CorrArgs &args, [[maybe_unused]] const ResOverride &ro = ResOverride{})
{
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr int AP = (A >= 2) ? A / 2 : 1;
    constexpr int R = 2 * A * L;
    constexpr int RP = (R + 31) / 32 * 32;
    constexpr int Q = RP / 32;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int W = args.W, S = args.S, AG = args.AG, SL = args.SL, G = args.G;
    const int TG = args.TG;                    // tap groups: warps that split the taps of one (satellite, antenna group)
    const int NR = S * AG * TG;
    const int MP = AG * A;
    const int M = args.n_ants, K = args.n_sats;
    const int stages = args.stages;
    const int tile_len = args.tile_len;
    const int TJ = args.tiles_per_job;
    // register-reallocation class (gat_internal.h): fixed warp positions, producer and replica warp in the fourth warpgroup
    constexpr bool REALLOC = HELP && help_realloc(A, L);
    constexpr int HS = REALLOC ? 1 : kHelperMaxSats;    // satellites per CTA the replica warp serves
    const int PW = REALLOC ? kReallocConsumerWarps : W;   // producer warp; the replica warps are PW + 1 ..
    // replica warps: the reallocation class fills its fourth warpgroup with three of them -- (slice, satellite) group g is
    // served by replica warp g % NREP.  (One warp for all slices was 87 % busy on the 11-tap shape and the consumers waited
    // for it: ncu source view, profiles/r03_ncu_c4_one_replica_warp.txt.)
    constexpr int NREP = REALLOC ? 3 : 1;

    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem);
    uint64_t *empty_bar = full_bar + kMaxStages;
    uint64_t *code_bar = reinterpret_cast<uint64_t *>(smem + 256);   // chip-table bulk copies, one phase per segment
    uint64_t *code_free = code_bar + 1;                              // back-pressure: every consumer warp has SEEN that phase
    const bool split = args.split_tiles != 0;
    float *tiles = reinterpret_cast<float *>(smem + kSmemHeaderBytes);
    const int tile_floats = (SC16 ? 1 : 2) * MP * kTileCap;   // 32-bit words per stage
    float *part = tiles + (size_t)stages * tile_floats;                               // [W][RP]
    float *rep_all = part + (size_t)W * RP;                                           // [rep_bufs][rep_stride]
    int8_t *code_cache = reinterpret_cast<int8_t *>(rep_all + (size_t)args.rep_bufs * args.rep_stride);  // [S][cache_stride]

    GAT_STAMP(warp == PW ? 8 : 0);
    if (tid == 0) {
        // (one thread, one barrier after the other: a lane-parallel version saved 0.3 us per small call and made the resident
        // kernel hang under back-to-back commands -- not understood, reverted)
        if constexpr (RES) {
            // mbarrier.init on a live mbarrier object is undefined (and does hang): invalidate the previous command's first
            if (ro.reinit) {
                for (int s = 0; s < stages; ++s) {
                    mbar_inval(&full_bar[s]);
                    mbar_inval(&empty_bar[s]);
                }
                mbar_inval(code_bar);
                mbar_inval(code_free);
                if constexpr (HELP) {
                    const int groups = (split ? 1 : SL) * S;
                    for (int i = 0; i < 2 * groups; ++i) {
                        mbar_inval(reinterpret_cast<uint64_t *>(smem + kRepBarOff) + i);
                        mbar_inval(reinterpret_cast<uint64_t *>(smem + 2 * kRepBarOff) + i);
                    }
                }
            }
        }
        for (int s = 0; s < stages; ++s) {
            mbar_init(&full_bar[s], 1);                            // producer's expect_tx arrival (+ TMA bytes)
            mbar_init(&empty_bar[s], (uint32_t)(split ? W : NR));  // one arrival per consumer warp that reads the stage
        }
        mbar_init(code_bar, 1);
        mbar_init(code_free, (uint32_t)(W + (HELP ? NREP : 0)));
        if constexpr (HELP) {
            // replica ring: two buffers per (slice, satellite) group; full = the replica warp's arrival, empty = one arrival
            // per consumer warp of the group
            const int groups = (split ? 1 : SL) * S;
            const uint32_t readers = (uint32_t)((split ? AG * SL : AG) * TG);
            for (int i = 0; i < 2 * groups; ++i) {
                mbar_init(reinterpret_cast<uint64_t *>(smem + kRepBarOff) + i, 1);
                mbar_init(reinterpret_cast<uint64_t *>(smem + 2 * kRepBarOff) + i, readers);
            }
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // antenna rows that pad M up to AG*A are never written by the copies: keep them zero
    if (MP > M) {
        const int pad_rows = MP - M;
        constexpr int kPlanes = SC16 ? 1 : 2;
        for (int i = tid; i < stages * kPlanes * pad_rows * kTileCap; i += blockDim.x) {
            const int col = i % kTileCap;
            const int row = (i / kTileCap) % pad_rows;
            const int plane = (i / (kTileCap * pad_rows)) % kPlanes;
            const int st = i / (kTileCap * pad_rows * kPlanes);
            tiles[(size_t)st * tile_floats + (size_t)(plane * MP + M + row) * kTileCap + col] = 0.f;
        }
    }
    cta_sync<ROLE>();
    if constexpr (ROLE == kRoleAux) {
        if (warp > PW + NREP) return;                 // (none with three replica warps)
    }
    if constexpr (ROLE == kRoleConsumer) {
        if (warp >= W) return;                        // plans with fewer than 12 consumer warps
    }

    // small calls carry their TMA descriptors and channel records in the kernel arguments
    const PeriodDev *periods = RES ? ro.periods : (args.use_inline ? reinterpret_cast<const PeriodDev *>(args.inline_blk) : args.periods);
    const SatDev *sats = RES ? ro.sats : (args.use_inline ? reinterpret_cast<const SatDev *>(args.inline_blk + args.inline_sat_off) : args.sats);
    const unsigned int barrier_target = RES ? ro.barrier_target : args.barrier_target;
    const int64_t TT = args.total_tiles;
    const int grid = gridDim.x;
    const int64_t r0 = (int64_t)blockIdx.x * TT / grid;
    const int64_t r1 = (int64_t)(blockIdx.x + 1) * TT / grid;
    uint32_t q = 0;    // running tile counter of this CTA -> ring stage and parity
    uint32_t seg = 0;  // running segment counter -> code_bar parity
    GAT_STAMP(warp == PW ? 9 : 1);

    if (ROLE != kRoleConsumer && warp == PW) {
        // ============================ producer warp ============================
        // Moves signal tiles (two 2-D TMA loads per tile) and keeps the chip table of every
        // satellite batched on this CTA in shared memory (one bulk copy per table change).
        const int8_t *cached_code = nullptr;  // lane s: which table sits in code_cache[s]
        for (int64_t g = r0; g < r1; ++seg) {
            const int job = (int)(g / TJ);
            const int t_first = (int)(g - (int64_t)job * TJ);
            const int t_last = (int)min((int64_t)TJ, (int64_t)t_first + (r1 - g));
            const int p = job / G, grp = job % G;
            const PeriodDev *per = &periods[(size_t)p * args.n_parts];
            const bool sat_ok = (lane < S) && (grp * S + lane < K);
            const int8_t *code = nullptr;
            int code_len = 0;
            if (sat_ok) {
                const SatDev *sd = &sats[(size_t)p * K + grp * S + lane];
                code = sd->code;
                code_len = sd->code_len;
            }
            const bool reload = sat_ok && code != cached_code;
            if (__any_sync(0xffffffffu, reload) && q > 0) {
                // the consumers still read the cached tables while they work on the previous
                // segment: wait until every tile issued so far has been released
                const int live = (int)min(q, (uint32_t)stages);
                for (int st = 0; st < live; ++st) mbar_wait_relaxed(&empty_bar[st], ((q - 1u - (uint32_t)st) / (uint32_t)stages) & 1u);
            }
            for (int t = t_first; t < t_last; ++t, ++q) {
                const int stage = q % stages;
                const uint32_t par = (q / stages) & 1u;
                mbar_wait_relaxed(&empty_bar[stage], par ^ 1u);
                const int ts_rel = t * tile_len;  // relative to aligned_start
                if (lane == 0) {
                    // full boxes always: samples past the block end arrive as zeros and still count
                    mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)((SC16 ? 1 : 2) * M * kTileCap * 4));
                    float *stage_tile = tiles + (size_t)stage * tile_floats;
                    int c0 = args.aligned_start + ts_rel;
                    const PeriodDev *src = per;
                    if (args.n_parts > 1) {
                        // sharded slot: this tile lives in ONE owner's HBM (local or a peer over NVLink); the all-gather of
                        // the block is this kernel's own tile pipeline -- no receive buffer, no second pass over HBM
                        const int part = min(c0 / (args.part_tiles * kTileCap), args.n_parts - 1);
                        src = per + part;
                        c0 -= part * args.part_tiles * kTileCap;
                    }
                    tma_load_2d(stage_tile, &src->re, c0, 0, &full_bar[stage]);   // SC16: the I/Q words
                    if constexpr (!SC16)
                        tma_load_2d(stage_tile + (size_t)MP * kTileCap, &src->im, c0, 0, &full_bar[stage]);
                }
                if (t == t_first) {
                    // chip tables AFTER the first tile is in flight; code_bar completes one phase per
                    // segment (with zero bytes when nothing changed) and releases the consumers.
                    // A parity wait only tells the current phase from the previous one, so the producer must not
                    // complete phase seg before every consumer warp has observed phase seg - 1: with a first
                    // segment shorter than the ring it could otherwise finish two phases before a consumer looks
                    // (the consumer then waits for a parity that is the CURRENT phase's -> dead CTA -> watchdog trap),
                    // or arrive while the previous phase's bulk-copy bytes are still pending.
                    if (seg > 0) mbar_wait_relaxed(code_free, (seg - 1u) & 1u);
                    const uint32_t my_bytes = reload ? (uint32_t)((code_len + 15) & ~15) : 0u;
                    const uint32_t all_bytes = __reduce_add_sync(0xffffffffu, my_bytes);
                    if (lane == 0) mbar_arrive_expect_tx(code_bar, all_bytes);
                    __syncwarp();
                    if (reload) bulk_g2s(code_cache + (size_t)lane * args.cache_stride, code, my_bytes, code_bar);
                    if (sat_ok) cached_code = code;
                    if (q == 0) GAT_STAMP(10);
                }
                if (q == 0) GAT_STAMP(11);
            }
            g += t_last - t_first;
        }
        GAT_STAMP(12);
        return;
    }

    if constexpr (HELP && ROLE != kRoleConsumer) {
        if (warp > PW && warp <= PW + NREP) {
            [[maybe_unused]] const int rj = warp - PW - 1;
            // ============================ replica warp ============================
            // Walks the CTA's tiles in order and writes each tile's code replica for every satellite of the group into the
            // ring buffer (slice, satellite, tile parity) its consumers will read: (frac, bmod) advance tile by tile, exactly
            // the arithmetic the consumer warps use when they generate their own rows.
            const int span = args.span;
            uint32_t qh = 0, segh = 0;
            if constexpr (REALLOC) {
                // One satellite per CTA; replica warp rj serves slice rj (split tiles: warp 0 serves the one group) and walks
                // exactly the visits its consumers walk: one replica of (tiles of the visit) * tile_len + span entries each.
                const int V = args.visit_tiles;
                const bool serve = split ? (rj == 0) : (rj < SL);
                const int grp_id = split ? 0 : rj;
                uint32_t use = 0;                    // visits served so far -> ring buffer and phase
                for (int64_t g = r0; g < r1; ++segh) {
                    const int job = (int)(g / TJ);
                    const int t_first = (int)(g - (int64_t)job * TJ);
                    const int t_last = (int)min((int64_t)TJ, (int64_t)t_first + (r1 - g));
                    const int p = job / G, grp = job % G;
                    const int n_seg = t_last - t_first;
                    const bool act = grp * S < K;
                    SatDev sd{};
                    if (act) sd = sats[(size_t)p * K + grp * S];
                    mbar_wait(code_bar, segh & 1u);      // this segment's chip tables are in shared memory
                    __syncwarp();
                    if (lane == 0) mbar_arrive(code_free);
                    if (serve) {
                        const int stride_o = split ? 1 : V * SL;
                        int o_run = split ? 0 : first_run_offset(qh, rj, V, SL);
                        for (; o_run < n_seg; o_run += stride_o, ++use) {
                            const int o = max(o_run, 0);
                            const int nt = min(o_run + (split ? 1 : V), n_seg) - o;
                            if (!act) continue;
                            const int t = t_first + o;
                            const int ts_rel = t * tile_len;
                            const int len = min(nt * tile_len, args.aligned_len - ts_rel);
                            const int n0 = args.aligned_start + ts_rel - args.start_sample;
                            const int rows = (((len + span + 31) >> 5) + 3) & ~3;
                            [[maybe_unused]] DumpDst dmp{nullptr, args.dump_stride, nt == 2 ? tile_len : 0x7fffffff};
                            if constexpr (DUMP) dmp.p = args.dump + (size_t)t * args.dump_stride;
                            uint64_t frac = 0;
                            uint32_t bmod = 0;
                            if constexpr (!F64) nco_tile_base(sd, (int64_t)n0 + args.phase_off + args.shifts[0], frac, bmod);
                            const int buf = 2 * grp_id + (int)(use & 1u);
                            float *rep = rep_all + (size_t)buf * args.rep_stride;
                            mbar_wait(reinterpret_cast<uint64_t *>(smem + 2 * kRepBarOff) + buf, ((use >> 1) & 1u) ^ 1u);   // its previous readers are done
                            gen_replica_rows<F64, DUMP>(args, lane, 0, rows, rep, smem_u32(rep), code_cache, smem_u32(code_cache), frac, bmod,
                                                        (uint64_t)sd.nco_delta, sd.nco_fp - 32, (uint32_t)sd.code_len, sd.code_ratio, sd.code_phase,
                                                        n0 + args.phase_off + args.shifts[0], dmp);
                            __syncwarp();
                            if (lane == 0) mbar_arrive(reinterpret_cast<uint64_t *>(smem + kRepBarOff) + buf);
                        }
                    }
                    qh += (uint32_t)n_seg;
                    g += n_seg;
                }
            } else
            for (int64_t g = r0; g < r1; ++segh) {
                const int job = (int)(g / TJ);
                const int t_first = (int)(g - (int64_t)job * TJ);
                const int t_last = (int)min((int64_t)TJ, (int64_t)t_first + (r1 - g));
                const int p = job / G, grp = job % G;
                uint64_t delta[HS], frac[HS], adv_frac[HS];
                int64_t nco_start[HS];
                uint32_t bmod[HS], adv_chips[HS], lc[HS];
                int fp[HS];
                double ratio[HS], cphase[HS];
                bool act[HS];
#pragma unroll
                for (int s = 0; s < HS; ++s) {
                    act[s] = s < S && grp * S + s < K;
                    delta[s] = frac[s] = adv_frac[s] = 0;
                    nco_start[s] = 0;
                    bmod[s] = adv_chips[s] = 0;
                    lc[s] = 1;
                    fp[s] = 32;
                    ratio[s] = cphase[s] = 0.0;
                    if (act[s]) {
                        const SatDev *sd = &sats[(size_t)p * K + grp * S + s];
                        delta[s] = (uint64_t)sd->nco_delta;
                        nco_start[s] = sd->nco_start;
                        fp[s] = sd->nco_fp;
                        lc[s] = (uint32_t)sd->code_len;
                        ratio[s] = sd->code_ratio;
                        cphase[s] = sd->code_phase;
                        if constexpr (!F64) {
                            const unsigned __int128 adv = (unsigned __int128)(uint32_t)tile_len * (unsigned __int128)delta[s];
                            adv_frac[s] = (uint64_t)adv & ((1ull << fp[s]) - 1ull);
                            adv_chips[s] = (uint32_t)((uint64_t)(adv >> fp[s]) % lc[s]);
                        }
                    }
                }
                mbar_wait(code_bar, segh & 1u);      // this segment's chip tables are in shared memory
                __syncwarp();
                if (lane == 0) mbar_arrive(code_free);
                for (int t = t_first; t < t_last; ++t, ++qh) {
                    const int ts_rel = t * tile_len;
                    const int len = min(tile_len, args.aligned_len - ts_rel);
                    const int n0 = args.aligned_start + ts_rel - args.start_sample;
                    const int rows = (((len + span + 31) >> 5) + 3) & ~3;          // whole groups of four rows (the buffer is padded)
                    const uint32_t use = split ? qh : qh / (uint32_t)SL;           // how often this tile's group has been served before
                    const int slice = split ? 0 : (int)(qh % (uint32_t)SL);
                    [[maybe_unused]] DumpDst dmp{nullptr, args.dump_stride, 0x7fffffff};
                    if constexpr (DUMP) dmp.p = args.dump + (size_t)t * args.dump_stride;
#pragma unroll
                    for (int s = 0; s < HS; ++s) {
                        if (!act[s]) continue;
                        if constexpr (!F64) {
                            if (t == t_first) {
                                SatDev tmp;
                                tmp.nco_delta = (int64_t)delta[s]; tmp.nco_start = nco_start[s]; tmp.nco_fp = fp[s]; tmp.code_len = (int32_t)lc[s];
                                nco_tile_base(tmp, (int64_t)n0 + args.phase_off + args.shifts[0], frac[s], bmod[s]);
                            } else {
                                frac[s] += adv_frac[s];
                                bmod[s] += adv_chips[s] + (uint32_t)(frac[s] >> fp[s]);
                                frac[s] &= (1ull << fp[s]) - 1ull;
                                if (bmod[s] >= lc[s]) bmod[s] -= lc[s];
                                if (bmod[s] >= lc[s]) bmod[s] -= lc[s];
                            }
                        }
                        if constexpr (NREP > 1) {
                            if ((slice * S + s) % NREP != rj) continue;    // another replica warp's group (the NCO state above still advanced)
                        }
                        const int buf = 2 * (slice * S + s) + (int)(use & 1u);
                        float *rep = rep_all + (size_t)buf * args.rep_stride;
                        const int8_t *tab = code_cache + (size_t)s * args.cache_stride;
                        mbar_wait(reinterpret_cast<uint64_t *>(smem + 2 * kRepBarOff) + buf, ((use >> 1) & 1u) ^ 1u);   // its previous readers are done
                        gen_replica_rows<F64, DUMP>(args, lane, 0, rows, rep, smem_u32(rep), tab, smem_u32(tab), frac[s], bmod[s], delta[s],
                                                    fp[s] - 32, lc[s], ratio[s], cphase[s], n0 + args.phase_off + args.shifts[0], dmp);
                        __syncwarp();
                        if (lane == 0) mbar_arrive(reinterpret_cast<uint64_t *>(smem + kRepBarOff) + buf);
                    }
                }
                g += t_last - t_first;
            }
            return;
        }
    }

    // ============================== consumer warps ==============================
    if constexpr (ROLE == kRoleAux) return;
    const int role = warp % NR;
    const int s_idx = role % S, ag = (role / S) % AG, tg = role / (S * AG);
    const int sl = warp / NR;
    const int consumer_threads = 32 * W;
    const int roles_rp = NR * RP;
    const int span = args.span;
    // warps that work on the same satellite and the same tile share one code replica
    const int gw = (split ? AG * SL : AG) * TG;             // warps per group
    const int gid = split ? s_idx : sl * S + s_idx;         // group id (< W)
    const int gr = (split ? sl * AG + ag : ag) * TG + tg;   // this warp's rank in its group
    // tap group tg holds taps tg * L ..: the host guarantees that their offsets RELATIVE to the group's first tap equal
    // koff4[0 .. L-1] (equally spaced taps), so the group offset goes into the lane's replica address once per tile and
    // the loop keeps addressing its taps through uniform registers
    const uint32_t tg_off = (uint32_t)args.koff4[tg * L];
    float *rep = rep_all + (size_t)(HELP ? 2 * gid : gid) * args.rep_stride;     // HELP: two ring buffers per group
    const int8_t *tab = code_cache + (size_t)s_idx * args.cache_stride;
    // 32-bit shared-memory addresses of everything the per-tile path touches, derived once (generic pointers make
    // ptxas re-derive the shared window base -- S2R CgaCtaId, LDC ... -- at every use) and pinned in registers
    uint32_t rep_s = smem_u32(rep), tab_s = smem_u32(tab);
    uint32_t bars_s = smem_u32(full_bar);                                         // full[i] at +8i, empty[i] at +8(kMaxStages + i)
    uint32_t tile0_s = smem_u32(tiles) + 4u * (uint32_t)(ag * A) * kTileCap;      // this warp's antenna rows in stage 0
    // (pinning pays where 16 / 8 antennas per thread amortise the four registers: same-box A/B, 264-channel block
    // 128.6 -> 124.3 us, int16 batch 169.5 -> 161.2 us; with 4 antennas per thread -- 11 taps, or the 96-register
    // class -- it costs 2-4 %, so those instantiations leave the choice to ptxas)
    if constexpr (A >= 8) asm volatile("" : "+r"(rep_s), "+r"(tab_s), "+r"(bars_s), "+r"(tile0_s));
    const uint32_t tile_bytes = 4u * (uint32_t)tile_floats;
    const uint32_t im_off = 4u * (uint32_t)MP * kTileCap;

    // replica rows (32 entries each) of a full tile and this warp's share of them: fixed for the whole launch;
    // only a job's last, shorter tile recomputes them.  Likewise the tile step of this warp, and the ring stage /
    // parity, advance by additions -- no integer division in the per-tile path.
    // (few values stay live across the FMA loop -- it runs within 7 registers of the 168 cap)
    const int rows_per_full = gw == 1 ? ((((tile_len + span + 31) >> 5) + 3) & ~3) : (((tile_len + span + 31) >> 5) + gw - 1) / gw;
    const int step = split ? 1 : SL;               // this warp works on every step-th tile of the CTA's sequence

    [[maybe_unused]] uint32_t use_run = 0;   // reallocation class: visits of this warp so far -> replica ring buffer and phase
    for (int64_t g = r0; g < r1; ++seg) {
        const int job = (int)(g / TJ);
        const int t_first = (int)(g - (int64_t)job * TJ);
        const int t_last = (int)min((int64_t)TJ, (int64_t)t_first + (r1 - g));
        const int p = job / G, grp = job % G;
        const int k = grp * S + s_idx;
        const bool active = k < K;

        // per-satellite constants
        uint64_t delta = 0, car_phase = 0, car_delta = 0;
        int64_t nco_start = 0;
        int fp = 32;
        uint32_t lc = 1;
        double ratio = 0.0, cphase = 0.0;
        if (active) {
            const SatDev *sd = &sats[(size_t)p * K + k];
            delta = (uint64_t)sd->nco_delta;
            nco_start = sd->nco_start;
            fp = sd->nco_fp;
            lc = (uint32_t)sd->code_len;
            car_phase = sd->car_phase;
            car_delta = sd->car_delta;
            ratio = sd->code_ratio;
            cphase = sd->code_phase;
        }
        const int sh = fp - 32;
        const int tt_stride = args.tt_stride;
        // inside a tile the carrier phase advances in 32 bits (2^-32 cycle per step, <= 8 steps, restarted
        // exactly from the 64-bit accumulator at every tile): error < 2e-9 cycle
        const uint32_t ph_step32 = (uint32_t)(((uint64_t)tt_stride * car_delta + 0x80000000ull) >> 32);
        uint64_t frac = 0;   // NCO state of the tile this warp works on (phase under its first sample, latest tap)
        uint32_t bmod = 0;
        bool have_base = false;
        // consecutive tiles of this warp are a fixed number of samples apart: split that advance into
        // whole chips and a 2^fp fraction ONCE per segment (128-bit), so the per-tile update is a few adds
        uint64_t adv_frac = 0;
        uint32_t adv_chips = 0;
        if (!F64 && active) {
            const unsigned __int128 adv = (unsigned __int128)(uint32_t)((split ? 1 : SL) * tile_len) * (unsigned __int128)delta;
            adv_frac = (uint64_t)adv & ((1ull << fp) - 1ull);
            adv_chips = (uint32_t)((uint64_t)(adv >> fp) % lc);
        }

        f32x2 accRe[AP][L], accIm[AP][L];
        float sRe[L], sIm[L];  // A == 1 path
#pragma unroll
        for (int l = 0; l < L; ++l) {
            sRe[l] = sIm[l] = 0.f;
#pragma unroll
            for (int a = 0; a < AP; ++a) accRe[a][l] = accIm[a][l] = 0ull;
        }

        if (args.flags & kFlagStallConsumers) __nanosleep(20000);   // test hook: let the producer run ahead
        mbar_wait(code_bar, seg & 1u);   // this segment's chip tables are in shared memory
        __syncwarp();                    // every lane is past the wait before the producer may flip the phase again
        if (lane == 0) mbar_arrive(code_free);

        if constexpr (REALLOC) {
            // ---- visits: runs of V consecutive tiles per slice (first_run_offset above); one replica, one prologue per visit ----
            static_assert(!REALLOC || (A >= 2 && !SC16), "reallocation class: packed FP32 tiles");
            const int V = split ? 1 : args.visit_tiles;
            const int n_seg = t_last - t_first;
            const int stride_o = split ? 1 : V * SL;
            int o_run = split ? 0 : first_run_offset(q, sl, V, SL);
            int o_prev = max(o_run, 0);
            int sp = 0;                                        // 2 * ring stage + phase parity of the visit's first tile
            if (o_run < n_seg) {
                const uint32_t qq = q + (uint32_t)o_prev;
                sp = 2 * (int)(qq % (uint32_t)stages) + (int)((qq / (uint32_t)stages) & 1u);
            }
            const bool stamp_first = (q == 0);
            q += (uint32_t)n_seg;
            const int lane_base = split ? sl * 32 + lane : lane;
            for (; o_run < n_seg; o_run += stride_o, ++use_run) {
                const int o = max(o_run, 0);
                const int nt = min(o_run + V, n_seg) - o;
                sp = sp_advance(sp, o - o_prev, stages);
                o_prev = o;
                const int ts_rel = (t_first + o) * tile_len;
                const int n0 = args.aligned_start + ts_rel - args.start_sample;  // relative index of the visit's sample 0
                const uint32_t rbuf = (uint32_t)(2 * gid) + (use_run & 1u);
                if (active) mbar_wait_s(smem_u32(smem + kRepBarOff) + 8u * rbuf, (use_run >> 1) & 1u);   // the replica warp wrote this visit's replica
                int tt0 = lane_base;
                if (n0 + tt0 < 0) tt0 += tt_stride;           // samples staged before start_sample (first tile of a job)
                uint32_t ph = (uint32_t)((car_phase + (uint64_t)(int64_t)(n0 + args.phase_off + tt0) * car_delta) >> 32);
                uint32_t ra = rep_s + (use_run & 1u) * 4u * (uint32_t)args.rep_stride + 4u * (uint32_t)tt0 + tg_off;
                int spj = sp;
#pragma unroll 1
                for (int j = 0; j < nt; ++j) {
                    const int stage = spj >> 1;
                    const int len = min(tile_len, args.aligned_len - ts_rel - j * tile_len);
                    mbar_wait_s(bars_s + 8u * (uint32_t)stage, (uint32_t)spj & 1u);
                    if (stamp_first && o == 0 && j == 0) GAT_STAMP(2);
                    if (active) {
                        // the lane's phase and replica address run on from the first tile of the visit (every lane has done
                        // its 8 x 32 samples there), only the staged tile changes
                        const uint32_t tre_s = tile0_s + (uint32_t)stage * tile_bytes;
                        uint32_t ta_re = tre_s + 4u * (uint32_t)(j == 0 ? tt0 : lane_base);
                        uint32_t ta_im = ta_re + im_off;
                        const uint32_t ta_end = tre_s + 4u * (uint32_t)len;
                        uint32_t astep = 4u * (uint32_t)tt_stride, pstep = ph_step32;
                        asm volatile("" : "+r"(astep), "+r"(pstep), "+r"(ra), "+r"(ph));        // opaque: keep them in registers
                        // one sample of this lane: carrier, the L chips, A antennas' wipe-off and taps
                        auto sample = [&](uint32_t a_re, uint32_t a_im, uint32_t a_rep) {
                            float cr, ci;
                            const float x = (float)(int32_t)ph * 1.4629180792671596e-9f;  // 2 pi / 2^32
                            __sincosf(x, &ci, &cr);
                            ph += pstep;
                            float chip[L];
#pragma unroll
                            for (int l = 0; l < L; ++l) chip[l] = lds_f32_at(a_rep + (uint32_t)args.koff4[l]);
                            const f32x2 CR = pack2(cr, cr), CI = pack2(ci, ci), NCI = pack2(-ci, -ci);
                            f32x2 X[AP], Y[AP];
#pragma unroll
                            for (int a = 0; a < AP; ++a) {
                                X[a] = pack2(lds_f32_at(a_re + 4u * (2 * a) * kTileCap), lds_f32_at(a_re + 4u * (2 * a + 1) * kTileCap));
                                Y[a] = pack2(lds_f32_at(a_im + 4u * (2 * a) * kTileCap), lds_f32_at(a_im + 4u * (2 * a + 1) * kTileCap));
                            }
#pragma unroll
                            for (int a = 0; a < AP; ++a) {
                                const f32x2 Dre = fma2(Y[a], CI, mul2(X[a], CR));
                                const f32x2 Dim = fma2(X[a], NCI, mul2(Y[a], CR));
#pragma unroll
                                for (int l = 0; l < L; ++l) {
                                    const f32x2 CH = pack2(chip[l], chip[l]);
                                    accRe[a][l] = fma2(Dre, CH, accRe[a][l]);
                                    accIm[a][l] = fma2(Dim, CH, accIm[a][l]);
                                }
                            }
                        };
#if GAT_FULL_UNROLL
                        // a full tile of a non-split plan is exactly 8 samples per lane, 128 bytes apart: straight-line code
                        // with immediate offsets, no loop branch (the branch's resolve stall was 7 % of the warp samples)
                        if (len == kTileCap && !split && ta_re == tre_s + 4u * (uint32_t)lane) {
#pragma unroll
                            for (int it = 0; it < kTileCap / 32; ++it) sample(ta_re + 128u * it, ta_im + 128u * it, ra + 128u * it);
                            ra += 128u * (kTileCap / 32);
                        } else
#endif
                        {
#pragma unroll 1
                            for (; ta_re < ta_end; ta_re += astep, ta_im += astep, ra += astep) sample(ta_re, ta_im, ra);
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive_s(bars_s + 8u * (uint32_t)(kMaxStages + stage));
                    spj = sp_advance(spj, 1, stages);
                }
                if (lane == 0 && active) mbar_arrive_s(smem_u32(smem + 2 * kRepBarOff) + 8u * rbuf);
            }
            g += n_seg;
        } else {
        // whole tiles go round-robin over the sample slices (tile q of the CTA belongs to slice q % SL); the first
        // one of this segment that is ours, its ring stage and parity: once per segment
        int t = t_first;
        if (step > 1) t += (sl + step - (int)(q % (uint32_t)step)) % step;
        int sp;                                            // 2 * ring stage + phase parity of tile t
        {
            const uint32_t qq = q + (uint32_t)(t - t_first);
            sp = 2 * (int)(qq % (uint32_t)stages) + (int)((qq / (uint32_t)stages) & 1u);
        }
        [[maybe_unused]] uint32_t use = 0;                 // HELP: how often this warp's group has been served before tile t
        if constexpr (HELP) {
            const uint32_t qq = q + (uint32_t)(t - t_first);
            use = split ? qq : qq / (uint32_t)SL;
        }
        const bool stamp_first = (q == 0);
        q += (uint32_t)(t_last - t_first);                 // every tile of the segment counts, ours or not
        for (; t < t_last; t += step) {
            const int stage = sp >> 1;
            const uint32_t par = (uint32_t)sp & 1u;
            const int ts_rel = t * tile_len;
            const int len = min(tile_len, args.aligned_len - ts_rel);
            const int n0 = args.aligned_start + ts_rel - args.start_sample;  // relative index of tile sample 0
            [[maybe_unused]] uint32_t rep_tile_s = rep_s;
            if constexpr (HELP) {
                // the replica warp wrote this tile's replica into ring buffer (use & 1) of the group
                rep_tile_s = rep_s + (use & 1u) * 4u * (uint32_t)args.rep_stride;
                if (active) mbar_wait_s(smem_u32(smem + kRepBarOff) + 8u * (uint32_t)(2 * gid + (int)(use & 1u)), (use >> 1) & 1u);
            }
            if (!HELP && active) {
                // ---- code replica of this tile, generated while the signal tile is still in flight ----
                // rep[u] = chip under (tile sample 0 + latest tap + u); tap l of sample tt reads rep[tt + koff[l]]
                // (the reference writes the same array to global memory, src/algorithms.jl:100-119, :1513-1525)
                int rows = (len + span + 31) >> 5;                    // 32 entries per row
                if (gw == 1) rows = (rows + 3) & ~3;                  // a warp on its own writes whole groups of four rows
                int rows_per = rows_per_full;                         // (the buffer is padded to 128 entries)
                if (len != tile_len) rows_per = (rows + gw - 1) / gw;
                const int row0 = gr * rows_per, row1 = min(rows, row0 + rows_per);
                [[maybe_unused]] DumpDst dmp{nullptr, args.dump_stride, 0x7fffffff};
                if constexpr (DUMP) dmp.p = args.dump + (size_t)t * args.dump_stride;
                if (gw > 1) group_bar_sync(2 + gid, 32 * gw); else __syncwarp();   // previous tile's readers are done
                if constexpr (!F64) {
                    if (!have_base) {
                        SatDev tmp;
                        tmp.nco_delta = (int64_t)delta; tmp.nco_start = nco_start; tmp.nco_fp = fp; tmp.code_len = (int32_t)lc;
                        nco_tile_base(tmp, (int64_t)n0 + args.phase_off + args.shifts[0], frac, bmod);
                        have_base = true;
                    } else {
                        frac += adv_frac;
                        bmod += adv_chips + (uint32_t)(frac >> fp);
                        frac &= (1ull << fp) - 1ull;
                        if (bmod >= lc) bmod -= lc;
                        if (bmod >= lc) bmod -= lc;
                    }
                }
                gen_replica_rows<F64, DUMP>(args, lane, row0, row1, rep, rep_s, tab, tab_s, frac, bmod, delta, sh, lc, ratio, cphase,
                                            n0 + args.phase_off + args.shifts[0], dmp);
                if (gw > 1) group_bar_sync(2 + gid, 32 * gw); else __syncwarp();
            }
            mbar_wait_s(bars_s + 8u * (uint32_t)stage, par);
            if (stamp_first && t == t_first) GAT_STAMP(2);
            if (active) {
                const uint32_t tre_s = tile0_s + (uint32_t)stage * tile_bytes;
                // tiles start on a 16-byte boundary, so the first tile of a job may stage <= 3 samples that lie
                // before start_sample (n0 < 0): the lanes that own them skip their first iteration.  Samples
                // past the end never reach the loop (tt < len).  The loop itself stays branch-free.
                int tt0 = split ? sl * 32 + lane : lane;
                if (n0 + tt0 < 0) tt0 += tt_stride;
                uint32_t ph = (uint32_t)((car_phase + (uint64_t)(int64_t)(n0 + args.phase_off + tt0) * car_delta) >> 32);
                // running shared-memory addresses of this lane's sample in the re / im planes and in the replica;
                // the loop carries nothing else (no sample counter): ptxas otherwise re-derives the tile base and
                // the phase step from the kernel arguments in every iteration
                uint32_t ta_re = tre_s + 4u * (uint32_t)tt0;
                uint32_t ta_im = ta_re + im_off;
                uint32_t ra = rep_tile_s + 4u * (uint32_t)tt0 + tg_off;
                const uint32_t ta_end = tre_s + 4u * (uint32_t)len;
                uint32_t astep = 4u * (uint32_t)tt_stride, pstep = ph_step32;
                asm volatile("" : "+r"(astep), "+r"(pstep));        // opaque: keep them in registers
                // (unrolling by two was measured slower on the 11-tap shape: occupancy, not per-warp ILP, is
                // what hides the MUFU / shared-memory latencies here)
                // one sample of this lane at the given shared-memory addresses (re plane, im plane, replica)
                auto sample = [&](uint32_t ta_re, uint32_t ta_im, uint32_t ra) {
                    // ---- carrier replica: exp(j 2 pi phase) ----
                    float cr, ci;
                    const float x = (float)(int32_t)ph * 1.4629180792671596e-9f;  // 2 pi / 2^32
                    __sincosf(x, &ci, &cr);
                    ph += pstep;
                    // ---- code replica chips for every tap: one shared-memory load each ----
                    float chip[L];
#pragma unroll
                    for (int l = 0; l < L; ++l) chip[l] = lds_f32_at(ra + (uint32_t)args.koff4[l]);
                    if constexpr (A >= 2) {
                        const f32x2 CR = pack2(cr, cr), CI = pack2(ci, ci), NCI = pack2(-ci, -ci);
                        // all of this sample's loads first (volatile: kept in this order, ahead of the math),
                        // so the packed-FMA stream below never waits on shared-memory latency
                        f32x2 X[AP], Y[AP];
#pragma unroll
                        for (int a = 0; a < AP; ++a) {
                            if constexpr (SC16) {
                                const uint32_t w0 = lds_u32_at(ta_re + 4u * (2 * a) * kTileCap), w1 = lds_u32_at(ta_re + 4u * (2 * a + 1) * kTileCap);
#if GAT_S16_MODE == 1
                                const uint32_t f0 = s16_flip(w0), f1 = s16_flip(w1);
                                const f32x2 NB = pack2(-kS16Bias, -kS16Bias);
                                X[a] = add2(pack2(s16_lo_biased(f0), s16_lo_biased(f1)), NB);
                                Y[a] = add2(pack2(s16_hi_biased(f0), s16_hi_biased(f1)), NB);
#else
                                X[a] = pack2(s16_lo(w0), s16_lo(w1));
                                Y[a] = pack2(s16_hi(w0), s16_hi(w1));
#endif
                            } else {
                                X[a] = pack2(lds_f32_at(ta_re + 4u * (2 * a) * kTileCap), lds_f32_at(ta_re + 4u * (2 * a + 1) * kTileCap));
                                Y[a] = pack2(lds_f32_at(ta_im + 4u * (2 * a) * kTileCap), lds_f32_at(ta_im + 4u * (2 * a + 1) * kTileCap));
                            }
                        }
#pragma unroll
                        for (int a = 0; a < AP; ++a) {
                            // d = s * conj(c):  d_re = s_re c_re + s_im c_im ; d_im = s_im c_re - s_re c_im
                            const f32x2 Dre = fma2(Y[a], CI, mul2(X[a], CR));
                            const f32x2 Dim = fma2(X[a], NCI, mul2(Y[a], CR));
#pragma unroll
                            for (int l = 0; l < L; ++l) {
                                const f32x2 CH = pack2(chip[l], chip[l]);
                                accRe[a][l] = fma2(Dre, CH, accRe[a][l]);
                                accIm[a][l] = fma2(Dim, CH, accIm[a][l]);
                            }
                        }
                    } else {
                        float xr, xi;
                        if constexpr (SC16) {
                            const uint32_t w = lds_u32_at(ta_re);
                            xr = s16_lo(w);
                            xi = s16_hi(w);
                        } else {
                            xr = lds_f32_at(ta_re);
                            xi = lds_f32_at(ta_im);
                        }
                        const float dre = fmaf(xi, ci, xr * cr);
                        const float dim = fmaf(-xr, ci, xi * cr);
#pragma unroll
                        for (int l = 0; l < L; ++l) {
                            sRe[l] = fmaf(dre, chip[l], sRe[l]);
                            sIm[l] = fmaf(dim, chip[l], sIm[l]);
                        }
                    }
                };
                // a full tile walked with one warp per tile is exactly 8 samples per lane, 128 bytes apart: straight-line code
                // for the many-tap shapes outside the reallocation class (8 satellites x 11 taps: 110 -> 97 us; slower at
                // 5 taps, 74 -> 81 us, and on raw int16 tiles, 169 -> 182 us, neutral at 3 taps: eight copies of a 16-antenna
                // body no longer fit the instruction cache next to the prologue -- so only from 7 taps on)
                constexpr bool kStraight = GAT_FULL_UNROLL >= 2 || (GAT_FULL_UNROLL == 1 && L >= 7);
                bool straight = false;
                if constexpr (kStraight) straight = (len == kTileCap && astep == 128u && tt0 == lane);
                if (straight) {
#pragma unroll
                    for (int it = 0; it < kTileCap / 32; ++it) sample(ta_re + 128u * it, ta_im + 128u * it, ra + 128u * it);
                } else {
#pragma unroll kLoopUnroll
                    for (; ta_re < ta_end; ta_re += astep, ta_im += astep, ra += astep) sample(ta_re, ta_im, ra);
                }
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive_s(bars_s + 8u * (uint32_t)(kMaxStages + stage));
                if constexpr (HELP)
                    if (active) mbar_arrive_s(smem_u32(smem + 2 * kRepBarOff) + 8u * (uint32_t)(2 * gid + (int)(use & 1u)));
            }
            if constexpr (HELP) ++use;
            sp += 2 * step;                                  // step <= stages: at most one wrap, which flips the parity
            if (sp >= 2 * stages) sp = (sp - 2 * stages) ^ 1;
        }
        g += t_last - t_first;
        }   // !REALLOC

        if (g >= r1) GAT_STAMP(3);
        // ------------------------------ flush this segment ------------------------------
        {
            float vr[RP];
#pragma unroll
            for (int i = 0; i < RP; ++i) vr[i] = 0.f;
            if constexpr (A >= 2) {
#pragma unroll
                for (int a = 0; a < AP; ++a)
#pragma unroll
                    for (int l = 0; l < L; ++l) {
                        unpack2(accRe[a][l], vr[(a * L + l) * 4 + 0], vr[(a * L + l) * 4 + 1]);
                        unpack2(accIm[a][l], vr[(a * L + l) * 4 + 2], vr[(a * L + l) * 4 + 3]);
                    }
            } else {
#pragma unroll
                for (int l = 0; l < L; ++l) {
                    vr[l * 2 + 0] = sRe[l];
                    vr[l * 2 + 1] = sIm[l];
                }
            }
            reduce_scatter<RP, 16>(vr, lane);
#pragma unroll
            for (int i = 0; i < Q; ++i) part[warp * RP + lane * Q + i] = vr[i];
        }
        consumer_bar_sync(consumer_threads);

        const int b_first = tile_owner((int64_t)job * TJ, grid, TT);
        const int b_last = tile_owner((int64_t)(job + 1) * TJ - 1, grid, TT);
        {
            // a job that lives entirely on this CTA is written straight out; otherwise this
            // CTA's share goes to its partial slot (job + b is unique along the stream-K staircase)
            const bool sole = (b_first == b_last);
            float *my_partial = args.partials + (size_t)(job + (int)blockIdx.x) * roles_rp;
            for (int x = tid; x < roles_rp; x += consumer_threads) {
                const int r = x / RP, e = x % RP;
                float acc = 0.f;
                for (int i = 0; i < SL; ++i) acc += part[(i * NR + r) * RP + e];
                if (sole)
                    emit_output<A, L, RES>(args, job, x, acc, RES ? ro.seq : 0u);
                else
                    __stcg(my_partial + x, acc);
            }
        }
        consumer_bar_sync(consumer_threads);  // `part` is free again
    }

    // ---------------- single-pass grid reduction: barrier, then every CTA finalises a share ----------------
    // All CTAs are co-resident (grid <= number of SMs, one CTA per SM), so a counting barrier is safe.
    __threadfence();
    consumer_bar_sync(consumer_threads);
    GAT_STAMP(4);
    if (tid == 0) {
        atomicAdd(args.grid_barrier, 1u);
        unsigned int seen;
        const uint64_t t0 = global_timer_ns();
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(args.grid_barrier) : "memory");
            if ((int)(seen - barrier_target) < 0 && global_timer_ns() - t0 > 5 * kWaitLimitNs) __trap();
        } while ((int)(seen - barrier_target) < 0);
    }
    consumer_bar_sync(consumer_threads);
    GAT_STAMP(5);
    {
        // GS lanes cooperate on one output element: lane j of the group sums contributors
        // b_first + j, + GS, ... in order, then a fixed xor tree combines them -> bit-reproducible.
        const int jobs = args.n_periods * G;
        const int64_t E = (int64_t)jobs * roles_rp;
        const int GS = args.fin_group;
        const int groups_per_cta = consumer_threads / GS;
        const int64_t n_groups = (int64_t)grid * groups_per_cta;
        const int gl = tid & (GS - 1);
        int64_t e = (int64_t)blockIdx.x * groups_per_cta + tid / GS;
        int64_t warp_e = (int64_t)blockIdx.x * groups_per_cta + (tid & ~31) / GS;   // warp-uniform loop bound
        for (; warp_e < E; warp_e += n_groups, e += n_groups) {
            float acc = 0.f;
            bool emit = false;
            int job = 0, x = 0;
            if (e < E) {
                job = (int)(e / roles_rp);
                x = (int)(e % roles_rp);
                const int b_first = tile_owner((int64_t)job * TJ, grid, TT);
                const int b_last = tile_owner((int64_t)(job + 1) * TJ - 1, grid, TT);
                if (b_first != b_last) {  // otherwise already written by its only owner
                    emit = true;
                    const float *src = args.partials + (size_t)(job + b_first + gl) * roles_rp + x;
                    const size_t step = (size_t)GS * roles_rp;
                    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                    int b = b_first + gl;
                    for (; b + 3 * GS <= b_last; b += 4 * GS, src += 4 * step) {
                        a0 += __ldcg(src);
                        a1 += __ldcg(src + step);
                        a2 += __ldcg(src + 2 * step);
                        a3 += __ldcg(src + 3 * step);
                    }
                    for (; b <= b_last; b += GS, src += step) a0 += __ldcg(src);
                    acc = (a0 + a1) + (a2 + a3);
                }
            }
            for (int o = GS >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (emit && gl == 0) emit_output<A, L, RES>(args, job, x, acc, RES ? ro.seq : 0u);
        }
    }
    if (args.n_peers >= 1) {
        // every CTA fences its (peer) stores system-wide and checks in; the last one releases the flags
        __threadfence_system();
        consumer_bar_sync(consumer_threads);
        if (tid == 0) {
            const unsigned int prev = atomicAdd(args.done_counter, 1u);
            if (prev == (unsigned int)grid - 1u) {
                *args.done_counter = 0u;   // self-cleaning for the next launch
                __threadfence_system();
                for (int d = 0; d < args.n_peers; ++d)
                    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(args.peer_flag[d] + args.my_rank), "r"(RES ? ro.seq : args.gather_seq) : "memory");
            }
        }
    }
    GAT_STAMP(6);
}

#ifndef GAT_RESIDENT_TU     // gat_resident.cu includes this file for the helpers and correlate_body above
template <int A, int L, bool F64, bool SC16, bool DUMP = false, bool HELP = false>
__global__ void __launch_bounds__(HELP ? block_threads_help(A, L) : block_threads_max(A, L), 1) correlate_kernel(const __grid_constant__ CorrArgs args)
{
    if constexpr (HELP && help_realloc(A, L)) {
        // every warp of a warpgroup executes the same setmaxnreg; the consumers' increase waits for the fourth group's decrease
        if ((threadIdx.x >> 5) >= kReallocConsumerWarps) {
            asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kReallocAuxRegs));
            correlate_body<A, L, F64, SC16, DUMP, HELP, kRoleAux>(args);
        } else {
            asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kReallocConsumerRegs));
            correlate_body<A, L, F64, SC16, DUMP, HELP, kRoleConsumer>(args);
        }
    } else {
        correlate_body<A, L, F64, SC16, DUMP, HELP, kRoleAll>(args);
    }
}

// --------------------------------------------------------------------------------------
// instantiation table + launcher
// --------------------------------------------------------------------------------------
typedef void (*KernelFn)(const CorrArgs);

template <int A, int L>
static KernelFn pick_mode(bool f64, bool sc16)
{
    if (sc16) return f64 ? nullptr : (KernelFn)correlate_kernel<A, L, false, true>;   // raw tiles: NCO convention only
    return f64 ? (KernelFn)correlate_kernel<A, L, true, false> : (KernelFn)correlate_kernel<A, L, false, false>;
}

// the replica-index dump exists for one shape of every (accumulator class, generation pattern): one antenna per thread,
// 16 antennas x 3 taps (the headline loop), 8 x 5, 4 x 7 and the tap-group shape 4 x 6 (two warps share 11 taps)
static KernelFn pick_dump_kernel(int A, int L, bool f64)
{
#define GAT_DUMP_CASE(a, l) \
    if (A == a && L == l) return f64 ? (KernelFn)correlate_kernel<a, l, true, false, true> : (KernelFn)correlate_kernel<a, l, false, false, true>;
#ifdef GAT_DEV_ONLY_4_11   // development builds: only the 11-tap instantiations (seconds instead of minutes of ptxas)
    GAT_DUMP_CASE(4, 11)
#else
    GAT_DUMP_CASE(1, 3) GAT_DUMP_CASE(16, 3) GAT_DUMP_CASE(8, 5) GAT_DUMP_CASE(4, 7) GAT_DUMP_CASE(4, 11) GAT_DUMP_CASE(4, 6)
#endif
#undef GAT_DUMP_CASE
    return nullptr;
}
bool dump_kernel_available(int A, int L) { return pick_dump_kernel(A, L, false) != nullptr; }

// replica-warp instantiations: the shapes where several antenna (or tap) groups share a tile with few satellites per CTA
static KernelFn pick_help_kernel(int A, int L, bool f64, bool dump)
{
#define GAT_HELP_CASE(a, l)                                                                                                                  \
    if (A == a && L == l) {                                                                                                                  \
        if (dump) return f64 ? (KernelFn)correlate_kernel<a, l, true, false, true, true> : (KernelFn)correlate_kernel<a, l, false, false, true, true>; \
        return f64 ? (KernelFn)correlate_kernel<a, l, true, false, false, true> : (KernelFn)correlate_kernel<a, l, false, false, false, true>;         \
    }
#ifdef GAT_DEV_ONLY_4_11
    GAT_HELP_CASE(4, 11)
#else
    GAT_HELP_CASE(4, 7) GAT_HELP_CASE(4, 9) GAT_HELP_CASE(8, 5) GAT_HELP_CASE(4, 11)
#endif
#undef GAT_HELP_CASE
    return nullptr;
}
bool help_kernel_available(int A, int L, bool f64, bool dump) { return pick_help_kernel(A, L, f64, dump) != nullptr; }

static KernelFn pick_kernel(int A, int L, bool f64, bool sc16)
{
#define GAT_CASE(a, l) \
    if (A == a && L == l) return pick_mode<a, l>(f64, sc16);
#ifdef GAT_DEV_ONLY_4_11
    GAT_CASE(4, 11)
#else
    GAT_CASE(1, 1) GAT_CASE(2, 1) GAT_CASE(4, 1) GAT_CASE(8, 1) GAT_CASE(16, 1)
    GAT_CASE(1, 3) GAT_CASE(2, 3) GAT_CASE(4, 3) GAT_CASE(8, 3) GAT_CASE(16, 3)
    GAT_CASE(4, 4) GAT_CASE(4, 6)                      // tap-group shapes (many taps split over two warps)
    GAT_CASE(1, 5) GAT_CASE(2, 5) GAT_CASE(4, 5) GAT_CASE(8, 5)
    GAT_CASE(1, 7) GAT_CASE(2, 7) GAT_CASE(4, 7)
    GAT_CASE(1, 9) GAT_CASE(2, 9) GAT_CASE(4, 9)
    GAT_CASE(1, 11) GAT_CASE(2, 11) GAT_CASE(4, 11)
#endif
#undef GAT_CASE
    return nullptr;
}

bool kernel_available(int A, int L) { return pick_kernel(A, L, false, false) != nullptr; }

cudaError_t configure_kernels()
{
    static const int As[] = {1, 2, 4, 8, 16};
    static const int Ls[] = {1, 3, 4, 5, 6, 7, 9, 11};
    for (int A : As)
        for (int L : Ls)
            for (int f = 0; f < 3; ++f) {
                KernelFn fn = pick_kernel(A, L, f == 1, f == 2);
                if (!fn) continue;
                cudaError_t e = cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
                if (e != cudaSuccess) return e;
                KernelFn more[3] = {f < 2 ? pick_dump_kernel(A, L, f == 1) : nullptr, f < 2 ? pick_help_kernel(A, L, f == 1, false) : nullptr,
                                    f < 2 ? pick_help_kernel(A, L, f == 1, true) : nullptr};
                for (KernelFn mf : more)
                    if (mf) {
                        e = cudaFuncSetAttribute((const void *)mf, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
                        if (e != cudaSuccess) return e;
                    }
            }
    return cudaSuccess;
}

cudaError_t launch_correlate(const LaunchPlan &plan, const CorrArgs &args, cudaStream_t stream)
{
    KernelFn fn = plan.help ? pick_help_kernel(plan.A, plan.L, plan.f64, plan.dump)
                            : (plan.dump ? pick_dump_kernel(plan.A, plan.L, plan.f64) : pick_kernel(plan.A, plan.L, plan.f64, plan.sc16));
    if (!fn) return cudaErrorInvalidValue;
    // Cooperative launch: the kernel ends with a grid-wide counting barrier, so all CTAs must be
    // co-resident.  grid <= #SMs with one CTA per SM satisfies that on an idle device; the cooperative
    // attribute makes the driver GUARANTEE it (two such kernels from different streams are then
    // serialised instead of dead-locking each other half-resident).
    void *kargs[] = {const_cast<CorrArgs *>(&args)};
    // (measured: a plain cudaLaunchKernel is not faster -- 27.9 vs 28.6 us per synchronous small call)
    return cudaLaunchCooperativeKernel(reinterpret_cast<const void *>(fn), dim3(plan.grid), dim3(plan.block), kargs, plan.smem, stream);
}

// --------------------------------------------------------------------------------------
// stream-ordered wait for the fused gather: lane r spins until rank r's flag reached `seq`
// --------------------------------------------------------------------------------------
// A peer that died or diverged never raises its flag: after 20 s the wait traps (-> a CUDA error on this rank's
// stream) instead of hanging the device for good.
__global__ void gather_wait_kernel(unsigned int *flags, int world, unsigned int seq)
{
    if ((int)threadIdx.x < world) {
        unsigned int v;
        uint64_t t0 = 0;
        unsigned int polls = 0;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + threadIdx.x) : "memory");
            if ((int)(v - seq) < 0 && (++polls & 1023u) == 0) {
                const uint64_t now = global_timer_ns();
                if (t0 == 0) t0 = now;
                else if (now - t0 > 5 * kWaitLimitNs) __trap();
            }
        } while ((int)(v - seq) < 0);
    }
}

__global__ void flag_signal_kernel(const FlagPtrs dst, int world, int my_rank, unsigned int seq)
{
    // stream order put everything this flag announces (H2D copies, the kernel that read the blocks) before this
    // kernel; the fence + release store make it visible system-wide before the flag
    if ((int)threadIdx.x < world) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(dst.p[threadIdx.x] + my_rank), "r"(seq) : "memory");
    }
}

cudaError_t launch_flag_signal(const FlagPtrs &dst, int world, int my_rank, unsigned int seq, cudaStream_t stream)
{
    flag_signal_kernel<<<1, 32, 0, stream>>>(dst, world, my_rank, seq);
    return cudaGetLastError();
}

cudaError_t launch_flag_wait(unsigned int *local_flags, int world, unsigned int seq, cudaStream_t stream)
{
    gather_wait_kernel<<<1, 32, 0, stream>>>(local_flags, world, seq);
    return cudaGetLastError();
}

cudaError_t launch_gather_wait(unsigned int *const *, unsigned int *local_flags, int world, unsigned int seq, cudaStream_t stream)
{
    gather_wait_kernel<<<1, 32, 0, stream>>>(local_flags, world, seq);
    return cudaGetLastError();
}

// --------------------------------------------------------------------------------------
// integer ingest: interleaved complex int16 / int8 -> FP32 planes (memory-bound streaming kernel)
// --------------------------------------------------------------------------------------
template <typename T>
__global__ void expand_sc_kernel(const T *__restrict__ iq, int64_t ld_in, float *__restrict__ re, float *__restrict__ im,
                                 int64_t ld_out, int n_samples, float scale)
{
    const int m = blockIdx.y;
    const T *src = iq + (int64_t)m * ld_in * 2;
    float *dre = re + (int64_t)m * ld_out, *dim = im + (int64_t)m * ld_out;
    // 4 complex samples per thread and trip when the row is suitably aligned, scalar otherwise
    const bool vec = ((reinterpret_cast<uintptr_t>(src) & (8 * sizeof(T) - 1)) == 0) && ((ld_out & 3) == 0) &&
                     ((reinterpret_cast<uintptr_t>(dre) & 15) == 0) && ((reinterpret_cast<uintptr_t>(dim) & 15) == 0);
    const int stride = gridDim.x * blockDim.x;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (vec) {
        const int quads = n_samples >> 2;
        for (int q = tid; q < quads; q += stride) {
            T v[8];
            if constexpr (sizeof(T) == 2) *reinterpret_cast<int4 *>(v) = __ldg(reinterpret_cast<const int4 *>(src) + q);
            else *reinterpret_cast<int2 *>(v) = __ldg(reinterpret_cast<const int2 *>(src) + q);
            *reinterpret_cast<float4 *>(dre + 4 * q) = make_float4(v[0] * scale, v[2] * scale, v[4] * scale, v[6] * scale);
            *reinterpret_cast<float4 *>(dim + 4 * q) = make_float4(v[1] * scale, v[3] * scale, v[5] * scale, v[7] * scale);
        }
        for (int n = (quads << 2) + tid; n < n_samples; n += stride) {
            dre[n] = (float)src[2 * n] * scale;
            dim[n] = (float)src[2 * n + 1] * scale;
        }
    } else {
        for (int n = tid; n < n_samples; n += stride) {
            dre[n] = (float)src[2 * n] * scale;
            dim[n] = (float)src[2 * n + 1] * scale;
        }
    }
}

cudaError_t launch_expand_sc(const void *iq, int bytes_per_component, int64_t ld_in, float *re, float *im, int64_t ld_out,
                             int n_samples, int n_ants, float scale, cudaStream_t stream)
{
    const int threads = 256;
    const int bx = max(1, min(64, (n_samples / 4 + threads - 1) / threads));
    dim3 grid(bx, n_ants);
    if (bytes_per_component == 2)
        expand_sc_kernel<int16_t><<<grid, threads, 0, stream>>>(static_cast<const int16_t *>(iq), ld_in, re, im, ld_out, n_samples, scale);
    else
        expand_sc_kernel<int8_t><<<grid, threads, 0, stream>>>(static_cast<const int8_t *>(iq), ld_in, re, im, ld_out, n_samples, scale);
    return cudaGetLastError();
}

// --------------------------------------------------------------------------------------
// synthetic signal generator, gen_signal semantics (src/gen_signal.jl:135-152)
// --------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__global__ void gen_signal_kernel(float *re, float *im, int64_t ld, const int8_t *code, int code_len,
                                  double code_ratio, double carrier_freq, double fs, double code_phase,
                                  double carrier_phase_rad, int n_samples, int n_ants, double ant_phase_step,
                                  double noise_sigma, uint64_t seed, int superpose)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_samples) return;
    // code_phases = fc/fs .* (0:N-1) .+ start_code_phase ; chips = codes[1 + mod(floor(Int, .), Lc)]
    const double cp = __dadd_rn(__dmul_rn(code_ratio, (double)i), code_phase);
    const int64_t idx = floormod64((int64_t)floor(cp), code_len);
    const float chip = (float)code[idx];
    // carrier_phases = Float32(2pi * i * f / fs + phase0)  (Float64, left to right, then cast)
    const double two_pi = 6.283185307179586;
    const double phd = __dadd_rn(__ddiv_rn(__dmul_rn(__dmul_rn(two_pi, (double)i), carrier_freq), fs), carrier_phase_rad);
    for (int m = 0; m < n_ants; ++m) {
        const float ph = (float)(phd + (double)m * ant_phase_step);
        float vr = cosf(ph) * chip;
        float vi = sinf(ph) * chip;
        if (noise_sigma > 0.0) {
            const uint64_t h = splitmix64(seed ^ splitmix64(((uint64_t)m << 32) | (uint32_t)i));
            const float u1 = ((float)(uint32_t)(h >> 40) + 0.5f) * (1.0f / 16777216.0f);
            const float u2 = ((float)(uint32_t)((h >> 16) & 0xFFFFFFu) + 0.5f) * (1.0f / 16777216.0f);
            const float rad = sqrtf(-2.0f * logf(u1)) * (float)noise_sigma;
            float sn, cs;
            sincospif(2.0f * u2, &sn, &cs);
            vr += rad * cs;
            vi += rad * sn;
        }
        const int64_t o = (int64_t)m * ld + i;
        if (superpose) {
            re[o] += vr;
            im[o] += vi;
        } else {
            re[o] = vr;
            im[o] = vi;
        }
    }
}

cudaError_t launch_gen_signal(float *re, float *im, int64_t ld, const int8_t *code, int code_len,
                              double code_ratio, double carrier_freq, double fs, double code_phase,
                              double carrier_phase_rad, int n_samples, int n_ants, double ant_phase_step_rad,
                              double noise_sigma, uint64_t seed, int superpose, cudaStream_t stream)
{
    const int threads = 256;
    const int blocks = (n_samples + threads - 1) / threads;
    gen_signal_kernel<<<blocks, threads, 0, stream>>>(re, im, ld, code, code_len, code_ratio, carrier_freq, fs,
                                                      code_phase, carrier_phase_rad, n_samples, n_ants,
                                                      ant_phase_step_rad, noise_sigma, seed, superpose);
    return cudaGetLastError();
}

#endif  // GAT_RESIDENT_TU

}  // namespace gat
