// gat_api.cu -- the C ABI of include/gat.h: contexts, chip tables, signal slots, launch
// planning and the correlate entry points.  Host logic only; the kernels live in
// gat_correlate.cu.  No CPU fallback exists: every entry point needs a live sm_100 device.
#include <new>

#include <atomic>
#include <cstdio>
#include <chrono>
#include "gat_ctx.h"

using namespace gat;

namespace {

// Host staging -> device for the per-call parameter block.  The upload runs on a side stream so
// that, with calls queued back to back, block i+1 is copied while kernel i runs; a ring of pinned
// buffers with two events each keeps host and device copies from being overwritten while in use.
int stage_params(gat_ctx *ctx, const void *src, size_t bytes, unsigned char **d_out, Staging **used)
{
    Staging &s = ctx->stg[ctx->stg_next];
    ctx->stg_next = (ctx->stg_next + 1) % kStagingRing;
    if (s.pending) {
        GAT_CUDA(ctx, cudaEventSynchronize(s.done));   // pinned buffer free again
        s.pending = false;
    }
    if (bytes > s.cap) {
        if (s.h) GAT_CUDA(ctx, cudaFreeHost(s.h));
        if (s.d) {
            GAT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            GAT_CUDA(ctx, cudaFree(s.d));
        }
        s.h = s.d = nullptr;
        s.cap = 0;
        const size_t grow = std::max<size_t>(bytes * 2, 4096);
        GAT_CUDA(ctx, cudaMallocHost(reinterpret_cast<void **>(&s.h), grow));
        GAT_CUDA(ctx, cudaMalloc(reinterpret_cast<void **>(&s.d), grow));
        s.cap = grow;
    }
    if (!s.done) {
        GAT_CUDA(ctx, cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
        GAT_CUDA(ctx, cudaEventCreateWithFlags(&s.consumed, cudaEventDisableTiming));
        GAT_CUDA(ctx, cudaEventRecord(s.consumed, ctx->stream));
    }
    std::memcpy(s.h, src, bytes);
    GAT_CUDA(ctx, cudaStreamWaitEvent(ctx->param_stream, s.consumed, 0));  // last reader of s.d is done
    GAT_CUDA(ctx, cudaMemcpyAsync(s.d, s.h, bytes, cudaMemcpyHostToDevice, ctx->param_stream));
    GAT_CUDA(ctx, cudaEventRecord(s.done, ctx->param_stream));
    GAT_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, s.done, 0));
    s.pending = true;
    *d_out = s.d;
    *used = &s;
    return GAT_OK;
}

// Q0.64 fraction of a real number of cycles (two's complement wrap == mod 1)
uint64_t cycles_to_q64(double x)
{
    double f = x - std::floor(x);  // [0, 1)
    if (!(f >= 0.0) || f >= 1.0) f = 0.0;
    return static_cast<uint64_t>(std::ldexp(f, 64));
}

int nco_fixed_point(int code_len)
{
    int bits = 0;
    while ((1LL << bits) < static_cast<long long>(code_len)) ++bits;
    return 63 - bits;
}

// Pre-digest one channel.  NCO constants follow Tracking.jl gen_code_replica! [upstream]:
//   fixed_point = 63 - ceil(log2(code_length)); delta = floor(fc * 2^fp / fs);
//   start = floor(mod(code_phase, code_length) * 2^fp)
int fill_sat(gat_ctx *ctx, const gat_channel &ch, double fs, SatDev &sd)
{
    if (ch.system_id < 0 || ch.system_id >= GAT_MAX_SYSTEMS) return fail(ctx, GAT_ERR_INVALID, "system_id out of range");
    const CodeTable &tab = ctx->codes[ch.system_id];
    if (!tab.d_chips) return fail(ctx, GAT_ERR_NO_CODES, "no chip table set for system " + std::to_string(ch.system_id));
    if (ch.prn < 1 || ch.prn > tab.n_prn) return fail(ctx, GAT_ERR_NO_CODES, "prn outside the chip table");
    if (!(ch.code_freq_hz > 0.0) || !std::isfinite(ch.code_freq_hz) || !std::isfinite(ch.code_phase_chips) ||
        !std::isfinite(ch.carrier_freq_hz) || !std::isfinite(ch.carrier_phase_cycles))
        return fail(ctx, GAT_ERR_INVALID, "non-finite or non-positive channel parameter");
    sd.code = tab.d_chips + static_cast<size_t>(ch.prn - 1) * tab.col_stride;
    sd.code_len = tab.code_len;
    sd.nco_fp = nco_fixed_point(tab.code_len);
    sd.nco_delta = static_cast<int64_t>(std::floor(ch.code_freq_hz * std::ldexp(1.0, sd.nco_fp) / fs));
    double modded = std::fmod(ch.code_phase_chips, static_cast<double>(tab.code_len));
    if (modded < 0) modded += static_cast<double>(tab.code_len);
    sd.nco_start = static_cast<int64_t>(std::floor(modded * std::ldexp(1.0, sd.nco_fp)));
    sd.car_phase = cycles_to_q64(ch.carrier_phase_cycles);
    sd.car_delta = cycles_to_q64(ch.carrier_freq_hz / fs);
    sd.code_ratio = ch.code_freq_hz / fs;
    sd.code_phase = ch.code_phase_chips;
    return GAT_OK;
}

struct Shape {
    int P, K, M, L, start, n;
    const int32_t *shifts;
    double max_ratio;
    int min_fp;
    int64_t max_delta;
    bool f64;
    int max_code_len;
    bool sc16;
    int min_code_len = 1 << 30;
    int n_parts = 1, part_tiles = 0;   // sharded slots (signal ring)
    int A = 1;                         // antennas per thread (template)
    int n_taps = 0;                    // the caller's tap count (output layout); L = taps per warp (template), TG * L >= n_taps
    int TG = 1;
    bool dump = false;                 // replica-index dump requested (debug instantiations)
    int reserve_smem = 0;              // bytes of dynamic shared memory kept free behind the plan (resident sessions: the command area)
};

// Antennas per thread A, taps per warp L (template arguments) and tap groups TG for a call's tap count -- see the comment
// where correlate_impl calls it.
static void choose_instantiation(int n_taps, int M, bool use_raw, const int32_t *shifts, int &A, int &L, int &TG)
{
    TG = 1;
    if (n_taps <= 3) {
        A = 16;
        L = n_taps == 2 ? 3 : n_taps;
    } else if (n_taps <= 5) {
        A = 8;
        L = 5;
    } else {
        A = 4;
        L = n_taps | 1;
    }
    A = std::min(A, pow2_ceil(M));
    A = std::max(1, std::min(A, env_int("GAT_TUNE_A", A)));
    if (A == 4 && n_taps >= env_int("GAT_TUNE_TG_MIN_TAPS", 99) && !use_raw) {
        // the second group's taps must sit at the first group's offsets shifted by one constant (the kernel addresses a
        // group's taps relative to its first tap): true for the equally spaced sets get_correlator_sample_shifts makes
        const int Lh = (n_taps + 1) / 2;   // 4, 5, 6
        bool same = true;
        for (int l = 0; l < Lh && Lh + l < n_taps; ++l) same = same && (shifts[Lh + l] - shifts[Lh] == shifts[l] - shifts[0]);
        if (same) {
            TG = 2;
            L = Lh;
        }
    }
}

// Choose the kernel instantiation and the CTA decomposition (DESIGN.md "Launch planning").
int make_plan(gat_ctx *ctx, const Shape &sh, LaunchPlan &plan, CorrArgs &a)
{
    const int L = sh.L, M = sh.M, K = sh.K, TG = sh.TG;   // L = taps per warp (template), TG warps split the call's taps
    const int A = sh.A;
    if (!kernel_available(A, L)) return fail(ctx, GAT_ERR_UNSUPPORTED, "no kernel for this (antennas, taps) shape");
    const int AG = (M + A - 1) / A;
    // Replica warp (HELP instantiation): with few satellites per CTA and several antenna / tap groups sharing each tile, one
    // extra warp generates all code replicas and the consumers only read them (gat_correlate.cu).  Needs two ring buffers per
    // (slice, satellite) group inside the W per-warp buffers, i.e. >= 2 warps per group.
    bool help = env_int("GAT_TUNE_REPHELPER", 1) != 0 && !sh.sc16 && AG * TG >= 2 && help_kernel_available(A, L, sh.f64, sh.dump);
    // (the register-reallocation class -- 11 taps -- has fixed warp positions: 12 consumer warps, one satellite per CTA)
    bool realloc_class = help && help_realloc(A, L);
    int w_cap = help ? (realloc_class ? kReallocConsumerWarps : block_threads_help(A, L) / 32 - 2) : max_consumer_warps(A, L);
    // Several satellites with >= 7 taps stay in the reallocation class, ONE satellite per CTA pass (G = K groups): consecutive
    // jobs re-read the period's block out of L2, and 12 consumer warps at 152 registers beat the plain instantiation's 8-10 at
    // 168 by more than the re-reads cost (8 satellites x 11 taps x 8 periods, same-box A/B: 16 antennas 97.4 -> 76.9 us,
    // 12 antennas 89.9 -> 76.9, 8 antennas 66.9 -> 62.9; 5 satellites x 9 taps x 8 antennas x 16 periods 71.1 -> 67.1)
    const bool realloc_multi = realloc_class && K > 1 && env_int("GAT_TUNE_REALLOC_MULTI", 1) != 0;
    if (help && !realloc_multi && std::min(K, std::max(1, w_cap / (AG * TG))) > (realloc_class ? 1 : kHelperMaxSats)) {
        // several satellites per CTA: the plain instantiation, and none of the reallocation class's sizing below (two-tile
        // replicas, two buffers per group) -- with it still on, 5 satellites x 9 / 11 taps x 8 antennas asked for 233 216 B of
        // shared memory and the cooperative launch was refused
        help = false;
        realloc_class = false;
        w_cap = max_consumer_warps(A, L);
    }
    if (AG * TG > w_cap) return fail(ctx, GAT_ERR_UNSUPPORTED, "too many antennas for this tap count");

    const int w_target_multi = std::min(w_cap, env_int("GAT_TUNE_WMAX", w_cap));
    // (two channels per block of 16 antennas -- the per-GPU step of the 2-GPU sample-sharded run: 5 slices x 2 channels = 10 warps over
    // a 5-stage ring, 0.185 -> 0.156 ms per 256 half-blocks against 3 slices over 6 stages)
    const bool two_channels = K == 2 && AG * TG == 1 && w_cap == 11;
    const int w_target_single = std::min(w_cap, env_int("GAT_TUNE_W", w_cap == 11 ? (two_channels ? 10 : 8) : 16));
    const int cache_stride = (sh.max_code_len + kCodeColAlign - 1) / kCodeColAlign * kCodeColAlign;
    const size_t smem_budget = 227 * 1024 - static_cast<size_t>(sh.reserve_smem);
    const int RW = AG * TG;            // warps per satellite and sample slice
    int S = std::max(1, std::min(K, w_target_multi / RW));
    S = std::max(1, std::min(S, env_int("GAT_TUNE_S", S)));
    if (realloc_class) S = 1;
    // replica entries a tile needs beyond its own samples (with two tap groups the pad tap may reach one spacing further)
    int span = sh.shifts[sh.n_taps - 1] - sh.shifts[0];
    if (TG == 2) span = std::max(span, (sh.shifts[L] - sh.shifts[0]) + (sh.shifts[L - 1] - sh.shifts[0]));
    // every satellite batched on a CTA keeps its chip table in smem: leave room for >= 2 stages,
    // the per-warp code replicas and the flush buffer
    {
        const size_t two_stages = kSmemHeaderBytes + 2 * smem_tile_floats(AG, A, sh.sc16) * sizeof(float) +
                                  static_cast<size_t>(w_cap) * (padded_acc(A, L) + kTileCap + span + 128) * sizeof(float);
        if (two_stages + cache_stride > smem_budget)
            return fail(ctx, GAT_ERR_UNSUPPORTED, "chip table too long for the shared-memory cache");
        S = std::max(1, std::min<int>(S, static_cast<int>((smem_budget - two_stages) / cache_stride)));
    }
    const int G = (K + S - 1) / S;
    S = (K + G - 1) / G;  // balance the groups
    const int RP = padded_acc(A, L);
    const size_t tile_bytes = smem_tile_floats(AG, A, sh.sc16) * sizeof(float);
    // tile coordinates stay multiples of 4 samples (16 B); the TMA unit zero-fills past the block end,
    // and the kernel masks the <= 3 samples staged before start_sample
    const int aligned_start = env_int("GAT_TUNE_NOALIGN", 0) ? sh.start : (sh.start & ~3);
    const int aligned_len = sh.start + sh.n - aligned_start;

    const int jobs = sh.P * G;
    // Tiles are always the full 256-sample TMA box.  Shrinking them to spread a small problem over more SMs was
    // measured slower on every single-block shape (same box, back-to-back / synchronous call: 2048 x 1 antenna
    // 17.7 -> 11.3 / 27.6 -> 22.4 us, 50000 x 16 14.4 -> 12.3 / 28.8 -> 26.6 us, 32 satellites 32.8 -> 28.8 us):
    // the per-tile work (replica generation, barriers) and the number of partials to finalise outweigh the idle SMs.
    int tile_len = env_int("GAT_TUNE_TILE", kTileCap);
    if (tile_len < 32 || tile_len > kTileCap || tile_len % 32) return fail(ctx, GAT_ERR_INVALID, "bad GAT_TUNE_TILE");
    if (sh.n_parts > 1 && (tile_len != kTileCap || aligned_start % kTileCap))
        return fail(ctx, GAT_ERR_UNSUPPORTED, "sharded (ring) slots need start_sample to be a multiple of 256");
    // Reallocation class: a consumer warp works through TWO consecutive tiles per visit (one code replica of 2 * tile + span
    // entries, one prologue, 16 loop iterations) unless the tiles are split among the slices.  Everything that is sized per
    // replica uses the longer window.
    int visit_max = realloc_class ? std::max(1, std::min(2, env_int("GAT_TUNE_VISIT", 2))) : 1;
    // per-replica relative NCO phase must fit 64 bits: (window + span + 1) * delta + 2^fp < 2^64 -- at very low sampling rates
    // (several chips per sample) only a one-tile window does
    auto window_need = [&](int len) {
        return static_cast<long double>(len + span + 160) * static_cast<long double>(sh.max_delta) + std::ldexp(1.0L, sh.min_fp);
    };
    if (!sh.f64 && visit_max == 2 && window_need(2 * tile_len) >= std::ldexp(1.0L, 64)) visit_max = 1;
    const int rep_len = visit_max * tile_len;
    if (!sh.f64) {
        const long double need = window_need(rep_len);
        if (need >= std::ldexp(1.0L, 64))
            return fail(ctx, GAT_ERR_UNSUPPORTED, "code rate too high for the fixed-point window (code_freq/fs * tile too large)");
    } else {
        const double worst = sh.max_ratio * (static_cast<double>(sh.n) + std::abs(sh.shifts[0]) + std::abs(sh.shifts[L - 1]));
        if (!(worst < 1.0e9)) return fail(ctx, GAT_ERR_UNSUPPORTED, "code phase range exceeds the f64 window arithmetic");
    }
    const int rep_stride = (rep_len + span + 127) & ~127;   // generated in rows of 32 entries, four rows at a time
    // (per-warp buffers sized for the most consumer warps this instantiation's CTA can hold; the reallocation class keeps two
    // replica buffers per (slice, satellite) group, at most three groups)
    const int rep_bufs_max = realloc_class ? 2 * std::min(3, std::max(1, w_cap / RW)) : w_cap;
    const size_t fixed_bytes = kSmemHeaderBytes + static_cast<size_t>(w_cap) * RP * sizeof(float) +
                               static_cast<size_t>(rep_bufs_max) * rep_stride * sizeof(float) + static_cast<size_t>(S) * cache_stride;

    const int tiles_per_job = (aligned_len + tile_len - 1) / tile_len;
    const int64_t total_tiles = static_cast<int64_t>(jobs) * tiles_per_job;
    if (fixed_bytes + tile_bytes > smem_budget)
        return fail(ctx, GAT_ERR_UNSUPPORTED, "shape does not fit shared memory (antennas x tap span)");
    int stages = static_cast<int>((smem_budget - fixed_bytes) / tile_bytes);
    // (up to 12 stages where the tiles are small -- few antennas, raw int16 words -- so that one channel
    // reading a block alone can still spread over 8..12 sample slices; 32 KB FP32 tiles fit 6)
    stages = std::min(stages, std::min(kMaxStages, env_int("GAT_TUNE_STAGES", 12)));
    stages = std::max(1, static_cast<int>(std::min<int64_t>(stages, std::max<int64_t>(1, total_tiles))));
    // sample slices take whole tiles round-robin, so more slices than stages cannot all be fed
    int SL = std::max(1, std::min(stages, w_target_single / (S * RW)));
    // few tiles per CTA: let every slice work on every tile instead of taking turns ("split").  Then the slices
    // stride through the tile together, 32 * SL samples per step, so SL is a power of two <= 8 (it divides the
    // 256-sample tile: every warp gets the same number of samples) and is no longer bounded by the stage count.
    const int grid_est = static_cast<int>(std::min<int64_t>(total_tiles, ctx->max_ctas > 0 ? std::min(ctx->n_sm, ctx->max_ctas) : ctx->n_sm));
    int split_tiles = (w_target_single / (S * RW) > 1 && total_tiles < static_cast<int64_t>(2) * SL * grid_est) ? 1 : 0;
    split_tiles = env_int("GAT_TUNE_SPLIT", split_tiles);
    // warps sharing a code replica meet at named barriers 2..15: at most 14 such groups
    if (split_tiles && S > 14) split_tiles = 0;
    if (split_tiles) {
        int cap = std::min(8, std::min(w_cap, w_target_single) / (S * RW));
        SL = 1;
        while (2 * SL <= cap && (tile_len % (64 * SL)) == 0) SL *= 2;
        if (SL == 1) split_tiles = 0;
    }
    if (!split_tiles) SL = std::max(1, std::min(stages, w_target_single / (S * RW)));
    SL = std::max(1, std::min(SL, env_int("GAT_TUNE_SL", SL)));
    int visit_tiles = 1;
    if (realloc_class && !split_tiles) {
        // one replica warp per slice; tile PAIRS go round-robin over the slices, so a stage stays with its slice when
        // 2 * SL divides the ring (see below)
        SL = std::min(SL, 3);
        if (visit_max == 2 && stages >= 2 * SL) {
            visit_tiles = 2;
            stages = stages / (2 * SL) * (2 * SL);
        }
    }
    if (!split_tiles && visit_tiles == 1) {
        // Whole tiles go round-robin over the slices AND over the ring stages.  The slice count must DIVIDE the stage count:
        // then a stage is always read by the same slice, and a consumer's parity wait on its `full` barrier can only be one
        // phase ahead.  Otherwise the stage's previous tile belongs to another slice; if that tile's TMA load is still in
        // flight when a faster slice comes back to the stage, the parity wait sees the phase before it as "complete", the
        // slice reads a tile that is not there and releases a stage it never owned -- wrong sums and, once the arrival
        // counts are off, a dead CTA.  (Found in round 2: 4 or 5 slices over 6 stages hung reliably under back-to-back
        // launches, and round 1's int16 plan had 8 slices over 12 stages; a protocol simulation reproduces the stale read.)
        // Either fewer slices (the largest divisor) or a shorter ring (the largest multiple of the slice count): the ring is
        // shortened only if it stays >= 8 stages deep -- small int16 tiles: 8 slices over 8 stages 170 us against 6 over 12
        // 195 us; 32 KB FP32 tiles: 3 slices over 6 stages 76 us against 4 over 4 84 us (5 taps x 16 antennas, 64 periods).
        if (SL > 1 && stages % SL != 0) {
            if (stages / SL * SL >= 8 || (two_channels && SL == 5 && stages > SL)) stages = stages / SL * SL;
            else
                while (SL > 1 && stages % SL != 0) --SL;
        }
    }
    const int W = S * RW * SL;
    if (W > w_cap || S > 32) return fail(ctx, GAT_ERR_UNSUPPORTED, "internal: role count exceeds CTA size");


    const int ctas_per_sm = 1;
    int grid = static_cast<int>(std::min<int64_t>(total_tiles, static_cast<int64_t>(ctx->n_sm) * ctas_per_sm));
    if (ctx->max_ctas > 0) grid = std::min(grid, ctx->max_ctas);
    grid = std::max(1, std::min(grid, env_int("GAT_TUNE_GRID", grid)));

    plan.A = A;
    plan.L = L;
    plan.f64 = sh.f64;
    plan.sc16 = sh.sc16;
    plan.grid = grid;
    plan.block = (help && help_realloc(A, L)) ? 512 : 32 * (W + 1 + (help ? 1 : 0));
    plan.help = help;
    plan.dump = sh.dump;
    a.rep_helper = help ? 1 : 0;
    const int rep_bufs = realloc_class ? 2 * (split_tiles ? S : SL * S) : W;
    plan.smem = kSmemHeaderBytes + stages * tile_bytes + static_cast<size_t>(W) * RP * sizeof(float) +
                static_cast<size_t>(rep_bufs) * rep_stride * sizeof(float) + static_cast<size_t>(S) * cache_stride;
    if (plan.smem > smem_budget) return fail(ctx, GAT_ERR_UNSUPPORTED, "internal: launch plan exceeds the shared-memory budget");
    a.rep_bufs = rep_bufs;
    a.visit_tiles = visit_tiles;
    a.dump_stride = (tile_len + span + 127) & ~127;
    plan.RP = RP;
    plan.jobs = jobs;

    // taps beyond the caller's count (TG * L > n_taps) are computed and dropped by emit_output: one group repeats the last
    // shift; with two groups the pad tap continues the second group's spacing (the kernel reads group offset + koff4[l])
    for (int l = 0; l < kMaxTaps; ++l) a.shifts[l] = sh.shifts[std::min(l, sh.n_taps - 1)];
    for (int l = 0; l <= kMaxTaps; ++l) a.koff4[l] = 4 * (sh.shifts[std::min(l, sh.n_taps - 1)] - sh.shifts[0]);
    a.span = span;
    a.TG = TG;
    a.n_periods = sh.P;
    a.n_sats = K;
    a.n_ants = M;
    a.n_taps = sh.n_taps;
    a.start_sample = sh.start;
    a.n_samples = sh.n;
    a.aligned_start = aligned_start;
    a.aligned_len = aligned_len;
    a.tile_len = tile_len;
    a.tiles_per_job = tiles_per_job;
    a.S = S;
    a.AG = AG;
    a.SL = SL;
    a.W = W;
    a.G = G;
    a.stages = stages;
    a.n_parts = sh.n_parts;
    a.part_tiles = sh.part_tiles;
    a.rep_stride = rep_stride;
    // chips advanced across one replica (tile + tap span, + the 32-entry row granularity) < shortest code
    a.rep_single_wrap = (static_cast<double>(rep_len + span + 160) * sh.max_ratio + 2.0 < static_cast<double>(sh.min_code_len)) ? 1 : 0;
    a.rep_single_wrap = env_int("GAT_TUNE_REPWRAP", a.rep_single_wrap) ? a.rep_single_wrap : 0;
    a.cache_stride = cache_stride;
    a.total_tiles = static_cast<int32_t>(total_tiles);
    {
        // finalize: a job's tiles are spread over at most ceil(TJ / tiles-per-CTA) + 1 CTAs
        const int64_t per_cta = std::max<int64_t>(1, total_tiles / grid);
        const int64_t max_contrib = std::min<int64_t>(grid, (tiles_per_job + per_cta - 1) / per_cta + 1);
        a.fin_group = max_contrib <= 8 ? 1 : 32;
        a.split_tiles = (split_tiles && SL > 1) ? 1 : 0;
        a.tt_stride = a.split_tiles ? 32 * SL : 32;
    }

    gat_launch_info &li = ctx->info;
    li.grid = grid;
    li.block = plan.block;
    li.smem_bytes = static_cast<int32_t>(plan.smem);
    li.ants_per_thread = A;
    li.ant_groups = AG;
    li.sats_per_cta = S;
    li.sample_slices = SL;
    li.consumer_warps = W;
    li.sat_groups = G;
    li.chunks_per_job = static_cast<int32_t>((grid + jobs - 1) / jobs);
    li.chunk_len = static_cast<int32_t>(total_tiles / grid * tile_len);
    li.tile_len = tile_len;
    li.stages = stages;
    li.items = static_cast<int32_t>(total_tiles);
    li.sc16 = sh.sc16 ? 1 : 0;
    return GAT_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// (re)build the two plane descriptors of a slot: dims {n_samples, n_ants}, box {kTileCap, n_ants}
int encode_slot_maps(gat_ctx *ctx, SignalSlot &s)
{
    s.tc_state = 0;
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return fail(ctx, GAT_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(s.n_samples), static_cast<cuuint64_t>(s.n_ants)};
    cuuint64_t row_bytes = static_cast<cuuint64_t>(s.ld) * sizeof(float);
    if (s.n_ants == 1) row_bytes = (static_cast<cuuint64_t>(s.n_samples) * sizeof(float) + 15) / 16 * 16;
    const cuuint64_t strides[1] = {row_bytes};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(kTileCap), static_cast<cuuint32_t>(s.n_ants)};
    const cuuint32_t estr[2] = {1, 1};
    float *planes[2] = {s.re, s.im};
    CUtensorMap *maps[2] = {&s.maps.re, &s.maps.im};
    for (int i = 0; i < 2; ++i) {
        CUresult r = enc(maps[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, planes[i], dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS)
            return fail(ctx, GAT_ERR_ALIGNMENT, "cuTensorMapEncodeTiled rejected the signal layout (CUresult " + std::to_string(r) + ")");
    }
    s.maps_valid = true;
    return GAT_OK;
}

SignalSlot *slot_for(gat_ctx *ctx, int slot)
{
    if (slot < 0 || slot >= 65536) return nullptr;
    return &ctx->slots[slot];
}

// drop the FP32 planes of a slot (owned storage is freed, a zero-copy binding is forgotten); the raw copy stays
int free_planes(gat_ctx *ctx, SignalSlot &s)
{
    if (s.owned && s.re) {
        GAT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        GAT_CUDA(ctx, cudaFree(s.re));
    }
    if (s.peer_base) {
        GAT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        GAT_CUDA(ctx, cudaIpcCloseMemHandle(s.peer_base));
        s.peer_base = nullptr;
    }
    s.re = s.im = nullptr;
    s.ld = 0;
    s.owned = false;
    s.cap_floats = 0;
    s.maps_valid = false;
    s.planes_valid = false;
    return GAT_OK;
}

int release_slot(gat_ctx *ctx, SignalSlot &s)
{
    int rc = free_planes(ctx, s);
    if (rc) return rc;
    if (s.raw) {
        GAT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        GAT_CUDA(ctx, cudaFree(s.raw));
    }
    s = SignalSlot{};
    return GAT_OK;
}

// make `s` hold owned FP32 planes of n_ants x ld floats each (one allocation: re then im).  The caller fills
// them in stream order; unless keep_raw, a raw integer copy of an older block is invalidated.
int own_slot(gat_ctx *ctx, SignalSlot &s, int n_samples, int n_ants, int64_t ld, bool keep_raw = false)
{
    const size_t need = static_cast<size_t>(ld) * n_ants;
    if (!s.owned || s.cap_floats < need) {
        int rc = free_planes(ctx, s);
        if (rc) return rc;
        GAT_CUDA(ctx, cudaMalloc(reinterpret_cast<void **>(&s.re), 2 * need * sizeof(float)));
        s.cap_floats = need;
        s.owned = true;
    }
    s.im = s.re + s.cap_floats;
    const bool same = s.maps_valid && s.ld == ld && s.n_samples == n_samples && s.n_ants == n_ants;
    s.ld = ld;
    s.n_samples = n_samples;
    s.n_ants = n_ants;
    s.planes_valid = true;
    if (!keep_raw) s.raw_valid = false;
    return same ? GAT_OK : encode_slot_maps(ctx, s);
}

// descriptor over the raw I/Q words: dims {n_samples, n_ants} of 32-bit elements, box {kTileCap, n_ants}
int encode_raw_map(gat_ctx *ctx, SignalSlot &s)
{
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return fail(ctx, GAT_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(s.n_samples), static_cast<cuuint64_t>(s.n_ants)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(s.raw_ld) * 4};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(kTileCap), static_cast<cuuint32_t>(s.n_ants)};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&s.raw_map.re, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, s.raw, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, GAT_ERR_ALIGNMENT, "cuTensorMapEncodeTiled rejected the raw layout (CUresult " + std::to_string(r) + ")");
    s.raw_map.im = s.raw_map.re;
    return GAT_OK;
}

// expand the raw integer copy into FP32 planes if that has not happened yet
int ensure_planes(gat_ctx *ctx, SignalSlot &s)
{
    if (s.planes_valid) return GAT_OK;
    if (!s.raw_valid) return fail(ctx, GAT_ERR_NO_SIGNAL, "slot has no signal");
    int rc = own_slot(ctx, s, s.n_samples, s.n_ants, (static_cast<int64_t>(s.n_samples) + 3) & ~3LL, true);
    if (rc) return rc;
    cudaError_t e = launch_expand_sc(s.raw, 2, s.raw_ld, s.re, s.im, s.ld, s.n_samples, s.n_ants, s.raw_scale, ctx->stream);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "expand_sc launch");
    ctx->launches += 1;
    return GAT_OK;
}

// 4-D view {4 samples, antennas, planes, sample groups} of a slot's two FP32 planes: one TMA box {4, 16, 2, 64} is the
// [sample/4][plane][antenna][sample%4] tile the tensor-core kernel uses as its K-major B operand
bool encode_tc_map(SignalSlot &s)
{
    if (s.tc_state) return s.tc_state > 0;
    s.tc_state = -1;
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc || !s.re || !s.im || s.im == s.re || s.n_ants > 16) return false;
    // the view starts at the lower plane; `swapped` tells the kernel's epilogue which one that is
    const bool swapped = s.im < s.re;
    float *lo = swapped ? s.im : s.re, *hi = swapped ? s.re : s.im;
    const uint64_t plane_bytes = static_cast<uint64_t>(reinterpret_cast<uintptr_t>(hi) - reinterpret_cast<uintptr_t>(lo));
    uint64_t row_bytes = static_cast<uint64_t>(s.ld) * sizeof(float);
    if (s.n_ants == 1) row_bytes = (static_cast<uint64_t>(s.n_samples) * sizeof(float) + 15) / 16 * 16;
    if (plane_bytes % 16 || plane_bytes >= (1ull << 40) || row_bytes % 16 || (reinterpret_cast<uintptr_t>(lo) & 15u)) return false;
    const cuuint64_t dims[4] = {4, static_cast<cuuint64_t>(s.n_ants), 2, static_cast<cuuint64_t>((s.n_samples + 3) / 4)};
    const cuuint64_t strides[3] = {row_bytes, plane_bytes, 16};
    const cuuint32_t box[4] = {4, 16, 2, 64};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    s.tc_map.swapped = swapped ? 1 : 0;
    CUresult r = enc(&s.tc_map.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, lo, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return false;
    s.tc_state = 1;
    return true;
}

int check_ctx(gat_ctx *ctx)
{
    if (!ctx) return GAT_ERR_INVALID;
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaSetDevice");
    return GAT_OK;
}

int correlate_impl(gat_ctx *ctx, int n_periods, const int32_t *slots, int n_sats, const gat_channel *channels,
                   double fs_hz, const int32_t *shifts, int n_taps, int start_sample, int n_samples,
                   float *out_re, float *out_im, int out_is_device, unsigned flags)
{
    NvtxRange nvtx_call("gat_correlate");
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (!slots || !channels || !shifts || ((!out_re || !out_im) && !(flags & GAT_GATHER)))
        return fail(ctx, GAT_ERR_INVALID, "null pointer argument");
    if (n_periods < 1 || n_sats < 1) return fail(ctx, GAT_ERR_INVALID, "n_periods and n_sats must be >= 1");
    if (ctx->res.active)
        return fail(ctx, GAT_ERR_INVALID, "a resident session owns the device (gat_resident_correlate / gat_resident_end)");
    if (n_taps < 1 || n_taps > GAT_MAX_TAPS) return fail(ctx, GAT_ERR_UNSUPPORTED, "n_taps must be 1..11");
    if (!(fs_hz > 0.0) || !std::isfinite(fs_hz)) return fail(ctx, GAT_ERR_INVALID, "sampling frequency must be positive");
    if (start_sample < 0 || n_samples < 1) return fail(ctx, GAT_ERR_INVALID, "empty or negative sample range");
    if ((flags & GAT_ACCUMULATE) && !out_is_device)
        return fail(ctx, GAT_ERR_INVALID, "GAT_ACCUMULATE needs device outputs");
    for (int l = 1; l < n_taps; ++l)
        if (shifts[l] < shifts[l - 1]) return fail(ctx, GAT_ERR_INVALID, "sample shifts must be ascending");
    if (static_cast<int64_t>(shifts[n_taps - 1]) - shifts[0] > 4096) return fail(ctx, GAT_ERR_UNSUPPORTED, "tap span > 4096 samples");

    const int32_t *sh_pad = shifts;      // (the tensor-core path below reads the caller's taps as they are)

    // signal slots
    int M = -1;
    bool all_raw = true, all_planes = true;
    float raw_scale = 1.f;
    int n_parts = -1, part_tiles = 0;
    for (int p = 0; p < n_periods; ++p) {
        const int s = slots[p];
        const SignalSlot *slp = find_slot(ctx, s);
        if (!slot_has_signal(slp)) return fail(ctx, GAT_ERR_NO_SIGNAL, "slot " + std::to_string(s) + " has no signal");
        const SignalSlot &sl = *slp;
        const int np = sl.parts.empty() ? 1 : static_cast<int>(sl.parts.size());
        if (n_parts < 0) {
            n_parts = np;
            part_tiles = sl.part_tiles;
        }
        if (np != n_parts || sl.part_tiles != part_tiles)
            return fail(ctx, GAT_ERR_INVALID, "a batch must not mix ring slots and plain slots (or rings of different geometry)");
        if (M < 0) M = sl.n_ants;
        if (sl.n_ants != M) return fail(ctx, GAT_ERR_INVALID, "all periods of a batch must have the same antenna count");
        if (static_cast<int64_t>(start_sample) + n_samples > sl.n_samples)
            return fail(ctx, GAT_ERR_INVALID, "sample range exceeds the signal in slot " + std::to_string(s));
        if (p == 0) raw_scale = sl.raw_scale;
        all_raw = all_raw && sl.raw_valid && sl.raw_scale == raw_scale;
        all_planes = all_planes && sl.planes_valid;
    }
    if (M < 1 || M > kMaxAnts) return fail(ctx, GAT_ERR_UNSUPPORTED, "antenna count must be 1..32");
    // Raw int16 tiles are read directly by the kernel: half the HBM bytes, and no expansion pass (4 B read +
    // 8 B written per sample and antenna, ~1.5 us per 50000 x 16 block).  Every channel converts the words it
    // reads on the issue-bound loop, so with many channels per block the FP32 planes win again.  Measured per
    // 256 channel-blocks of 50000 x 16: one channel per block 171 us raw vs 250 us FP32 (HBM-bound), two 163 vs
    // 174, four 160 vs 162.  So: up to two channels per block always read the raw words; more do so only while
    // nobody has paid for the FP32 planes yet, up to 16 channels per block.  The scale must be a power of two so
    // that applying it to the accumulators is bit-identical to scaling every sample.
    bool use_raw = false;
    {
        int mant_exp = 0;
        const bool pow2 = raw_scale > 0.f && std::frexp(raw_scale, &mant_exp) == 0.5f;
        const int pref = env_int("GAT_TUNE_RAW", -1);
        const bool worth = n_sats <= 2 || (!all_planes && n_sats <= 16);
        use_raw = all_raw && pow2 && !(flags & GAT_CODE_PHASE_F64) && (pref < 0 ? worth : pref != 0);
        if (flags & GAT_TENSOR_TF32) use_raw = false;       // the tensor-core path works on the FP32 planes
    }
    const bool sharded = part_tiles > 0;
    if (sharded) {
        use_raw = false;
        flags &= ~static_cast<unsigned>(GAT_TENSOR_TF32);     // ring slots run on the FP32 kernel
    }
    std::vector<PeriodDev> periods(static_cast<size_t>(n_periods) * n_parts);
    for (int p = 0; p < n_periods; ++p) {
        SignalSlot &sl = ctx->slots[slots[p]];
        if (sharded) {
            for (int j = 0; j < n_parts; ++j) periods[static_cast<size_t>(p) * n_parts + j] = sl.parts[j].maps;
        } else if (use_raw) {
            periods[p] = sl.raw_map;
        } else {
            rc = ensure_planes(ctx, sl);
            if (rc) return rc;
            if (!sl.maps_valid) return fail(ctx, GAT_ERR_NO_SIGNAL, "slot " + std::to_string(slots[p]) + " has no TMA descriptor");
            periods[p] = sl.maps;
        }
    }

    // channels
    const size_t n_ch = static_cast<size_t>(n_periods) * n_sats;
    std::vector<SatDev> sats(n_ch);
    // Kernel shape.  Antennas per thread: as many as ~96 accumulator registers allow -- the carrier and the tap loads are per
    // THREAD, so fewer antennas per thread means more redundant work (measured, 64 periods x 16 antennas: 3 taps x 32
    // satellites A=16 143 us, A=8 168 us, A=4 247 us).  Taps per warp L: the instantiated counts are 1, 3, 5, 7, 9, 11 (and
    // 4, 6 with 4 antennas); a call with fewer taps than its instantiation repeats the last shift and the extra taps are
    // dropped when the accumulators are written.  From 8 taps on (4 antennas per thread) the taps are split over TWO warps
    // ("tap groups"): 2 x 4 x 6 accumulators instead of 88 put the shape into the 19-warp class; the wipe-off is repeated by
    // both, which costs 23 % more FMAs and wins through occupancy (C4: see DESIGN.md).
    int A, L, TG = 1;
    choose_instantiation(n_taps, M, use_raw, shifts, A, L, TG);
    Shape shape{n_periods, n_sats, M, L, start_sample, n_samples, shifts, 0.0, 63, 0, (flags & GAT_CODE_PHASE_F64) != 0, 1, use_raw};
    shape.n_parts = n_parts;
    shape.part_tiles = part_tiles;
    shape.A = A;
    shape.TG = TG;
    shape.n_taps = n_taps;
    shape.dump = (flags & kFlagDumpReplica) != 0;
    shape.reserve_smem = (flags & kFlagResidentPlan) ? kResCmdSmemBytes + 128 : 0;
    for (size_t i = 0; i < n_ch; ++i) {
        rc = fill_sat(ctx, channels[i], fs_hz, sats[i]);
        if (rc) return rc;
        shape.max_ratio = std::max(shape.max_ratio, sats[i].code_ratio);
        shape.min_fp = std::min(shape.min_fp, sats[i].nco_fp);
        shape.max_delta = std::max(shape.max_delta, sats[i].nco_delta);
        shape.max_code_len = std::max(shape.max_code_len, sats[i].code_len);
        shape.min_code_len = std::min(shape.min_code_len, sats[i].code_len);
    }

    // ---- tensor-core path (opt-in): many channels over the same block(s), see gat_correlate_tc.cu ----
    ctx->info.tensor = 0;
    if (ctx->sample_origin_on) flags &= ~static_cast<unsigned>(GAT_TENSOR_TF32);   // sample ranges run on the FP32 kernel
    if (flags & GAT_TENSOR_TF32) {
        const int span = sh_pad[n_taps - 1] - sh_pad[0];
        bool ok = !(flags & (GAT_CODE_PHASE_F64 | GAT_ACCUMULATE | GAT_GATHER)) && n_taps <= 4 && M <= 16 && span <= 224 &&
                  shape.max_code_len <= 10240 &&      // kTcTabWords * 32 chips of sign bits per channel in shared memory (GPS L5: 10 230)
                  static_cast<double>(kTileCap + span + 192) * shape.max_ratio + 2.0 < static_cast<double>(shape.min_code_len);
        {
            const long double need = static_cast<long double>(kTileCap + span + 192) * static_cast<long double>(shape.max_delta) +
                                     std::ldexp(1.0L, shape.min_fp);
            ok = ok && need < std::ldexp(1.0L, 64);
        }
        for (int p = 0; ok && p < n_periods; ++p) ok = encode_tc_map(ctx->slots[slots[p]]);
        if (ok) {
            const int aligned_start = start_sample & ~3;
            const int aligned_len = start_sample + n_samples - aligned_start;
            const int G = (n_sats + 31) / 32;
            const int jobs = n_periods * G;
            const int tiles_per_job = (aligned_len + kTileCap - 1) / kTileCap;
            const int64_t total_units = static_cast<int64_t>(jobs) * tiles_per_job;
            int grid = static_cast<int>(std::min<int64_t>(total_units, ctx->n_sm));
            if (ctx->max_ctas > 0) grid = std::min(grid, ctx->max_ctas);
            // parameter block: [TcPeriod x P][SatDev x P*K]
            const size_t per_bytes = sizeof(TcPeriod) * n_periods;
            std::vector<unsigned char> blk(per_bytes + sizeof(SatDev) * n_ch);
            for (int p = 0; p < n_periods; ++p) std::memcpy(blk.data() + sizeof(TcPeriod) * p, &ctx->slots[slots[p]].tc_map, sizeof(TcPeriod));
            std::memcpy(blk.data() + per_bytes, sats.data(), sizeof(SatDev) * n_ch);
            unsigned char *d_blk = nullptr;
            Staging *stg = nullptr;
            rc = stage_params(ctx, blk.data(), blk.size(), &d_blk, &stg);
            if (rc) return rc;
            rc = ensure_device(ctx, ctx->d_partials, ctx->partials_cap, (static_cast<size_t>(jobs) + grid) * 2 * 128 * 16, false);
            if (rc) return rc;
            const size_t out_elems = n_ch * static_cast<size_t>(n_taps) * M;
            if (!out_is_device && 2 * out_elems > ctx->h_out_cap) {
                if (ctx->h_out) {
                    GAT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                    GAT_CUDA(ctx, cudaFreeHost(ctx->h_out));
                }
                ctx->h_out = nullptr;
                ctx->h_out_cap = 0;
                GAT_CUDA(ctx, cudaHostAlloc(reinterpret_cast<void **>(&ctx->h_out), 4 * out_elems * sizeof(float), cudaHostAllocMapped));
                ctx->h_out_cap = 4 * out_elems;
            }
            TcArgs ta{};
            ta.periods = reinterpret_cast<const TcPeriod *>(d_blk);
            ta.sats = reinterpret_cast<const SatDev *>(d_blk + per_bytes);
            ta.partials = ctx->d_partials;
            ta.out_re = out_is_device ? out_re : ctx->h_out;
            ta.out_im = out_is_device ? out_im : ctx->h_out + out_elems;
            ta.n_periods = n_periods;
            ta.n_sats = n_sats;
            ta.n_ants = M;
            ta.n_taps = n_taps;
            ta.shift0 = sh_pad[0];
            ta.span = span;
            for (int l = 0; l < 4; ++l) ta.koff[l] = l < n_taps ? sh_pad[l] - sh_pad[0] : 0;
            ta.start_sample = start_sample;
            ta.n_samples = n_samples;
            ta.aligned_start = aligned_start;
            ta.tiles_per_job = tiles_per_job;
            ta.G = G;
            ta.total_units = total_units;
            ta.win_ok = (static_cast<double>(kTileCap + span + 64) * shape.max_ratio + 2.0 < 32.0 && !env_int("GAT_TC_NO_WINDOW", 0)) ? 1 : 0;
            ta.debug = env_int("GAT_TC_DEBUG", 0);
            ta.dump = nullptr;
            if (flags & kFlagDumpReplica) {
                const size_t n_dump = static_cast<size_t>(total_units) * 32 * 20;
                rc = ensure_device(ctx, ctx->d_dbg, ctx->d_dbg_cap, n_dump, false);
                if (rc) return rc;
                GAT_CUDA(ctx, cudaMemsetAsync(ctx->d_dbg, 0, n_dump * sizeof(int32_t), ctx->stream));
                ta.dump = reinterpret_cast<uint32_t *>(ctx->d_dbg);
                ctx->dump_tiles = tiles_per_job;
                ctx->dump_aligned_start = aligned_start;
            }
            if (ctx->timing) GAT_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
            NvtxRange nvtx_launch("correlate_tc_kernel + tc_finalize_kernel");
            cudaError_t e = launch_correlate_tc(ta, grid, jobs, ctx->stream);
            if (e != cudaSuccess) return cuda_fail(ctx, e, "tensor-core correlate launch");
            GAT_CUDA(ctx, cudaEventRecord(stg->consumed, ctx->stream));
            if (ctx->timing) GAT_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
            ctx->launches += 2;
            gat_launch_info &li = ctx->info;
            li = gat_launch_info{};
            li.grid = grid;
            li.block = 32 * 17;
            li.sats_per_cta = 32;
            li.sat_groups = G;
            li.tile_len = kTileCap;
            li.stages = 2;
            li.items = static_cast<int32_t>(total_units);
            li.kernels_launched = 2;
            li.tensor = 1;
            if (!out_is_device) {
                GAT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                std::memcpy(out_re, ctx->h_out, out_elems * sizeof(float));
                std::memcpy(out_im, ctx->h_out + out_elems, out_elems * sizeof(float));
            }
            if (ctx->timing) {
                GAT_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
                float ms = 0.f;
                GAT_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
                ctx->info.last_kernel_ms = ms;
            }
            return GAT_OK;
        }
        // otherwise: the shape is outside the tensor-core path's envelope -> the FP32 kernel below
        if (flags & kFlagDumpReplica) return fail(ctx, GAT_ERR_UNSUPPORTED, "shape outside the tensor-core path's envelope");
    }

    LaunchPlan plan{};
    CorrArgs args{};
    rc = make_plan(ctx, shape, plan, args);
    if (rc) return rc;
    args.flags = flags;
    args.out_scale = use_raw ? raw_scale : 1.f;
    args.phase_off = ctx->sample_origin_on ? ctx->sample_origin + start_sample : 0;

    // parameter block: [PeriodDev x P][SatDev x P*K]
    const size_t per_bytes = sizeof(PeriodDev) * periods.size();
    const size_t sat_off = (per_bytes + 63) & ~static_cast<size_t>(63);
    const size_t blk_bytes = sat_off + sizeof(SatDev) * n_ch;
    std::vector<unsigned char> blk(blk_bytes);
    std::memcpy(blk.data(), periods.data(), per_bytes);
    std::memcpy(blk.data() + sat_off, sats.data(), sizeof(SatDev) * n_ch);
    Staging *stg = nullptr;
    if (blk_bytes <= static_cast<size_t>(kInlineBytes)) {
        // small call (one period, up to ~40 channels): the block rides in the kernel arguments
        std::memcpy(args.inline_blk, blk.data(), blk_bytes);
        args.use_inline = 1;
        args.inline_sat_off = static_cast<int32_t>(sat_off);
        args.periods = nullptr;
        args.sats = nullptr;
    } else {
        unsigned char *d_blk = nullptr;
        rc = stage_params(ctx, blk.data(), blk_bytes, &d_blk, &stg);
        if (rc) return rc;
        args.use_inline = 0;
        args.periods = reinterpret_cast<const PeriodDev *>(d_blk);
        args.sats = reinterpret_cast<const SatDev *>(d_blk + sat_off);
    }

    // scratch
    const size_t roles_rp = static_cast<size_t>(args.S) * args.AG * args.TG * plan.RP;
    rc = ensure_device(ctx, ctx->d_partials, ctx->partials_cap, (static_cast<size_t>(plan.jobs) + plan.grid) * roles_rp, false);
    if (rc) return rc;
    if (!ctx->d_barrier) {
        GAT_CUDA(ctx, cudaMalloc(reinterpret_cast<void **>(&ctx->d_barrier), sizeof(unsigned int)));
        GAT_CUDA(ctx, cudaMemsetAsync(ctx->d_barrier, 0, sizeof(unsigned int), ctx->stream));
        ctx->barrier_count = 0;
    }
    // the counter only moves once the launch has succeeded: a call that fails below must not leave the
    // host ahead of the device (the next kernel would wait for arrivals that never come)
    const unsigned int barrier_target = ctx->barrier_count + static_cast<unsigned int>(plan.grid);
    args.partials = ctx->d_partials;
    args.grid_barrier = ctx->d_barrier;
    args.barrier_target = barrier_target;

    const bool gather = (flags & GAT_GATHER) != 0;
    if (gather) {
        if (!ctx->g_connected) return fail(ctx, GAT_ERR_INVALID, "GAT_GATHER needs gat_gather_create + gat_gather_connect");
        if (flags & GAT_ACCUMULATE) return fail(ctx, GAT_ERR_UNSUPPORTED, "GAT_GATHER cannot be combined with GAT_ACCUMULATE");
        if (ctx->g_off + n_ch * static_cast<size_t>(n_taps) * M > ctx->g_elems)
            return fail(ctx, GAT_ERR_INVALID, "gather buffer too small for this call (elements + gat_gather_set_offset)");
    }
    const size_t out_elems = n_ch * n_taps * M;
    // the kernel writes the caller's [n_ants x n_taps x n_sats x n_periods] layout itself (padded taps are dropped at the
    // store), to the device pointers given, or -- host results -- straight into pinned, device-mapped host memory (posted
    // writes over PCIe), so the call needs no D2H copy, just the stream synchronisation
    const bool direct = gather || out_is_device;
    const bool host_direct = !direct;
    if (host_direct && 2 * out_elems > ctx->h_out_cap) {
        if (ctx->h_out) {
            GAT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            GAT_CUDA(ctx, cudaFreeHost(ctx->h_out));
        }
        ctx->h_out = nullptr;
        ctx->h_out_cap = 0;
        GAT_CUDA(ctx, cudaHostAlloc(reinterpret_cast<void **>(&ctx->h_out), 4 * out_elems * sizeof(float), cudaHostAllocMapped));
        ctx->h_out_cap = 4 * out_elems;
    }
    args.n_peers = 0;
    if (gather) {
        args.out_re = ctx->g_re[ctx->g_rank] + static_cast<size_t>(ctx->g_rank) * ctx->g_elems + ctx->g_off;
        args.out_im = ctx->g_im[ctx->g_rank] + static_cast<size_t>(ctx->g_rank) * ctx->g_elems + ctx->g_off;
        args.gather_off = ctx->g_off;
        for (int d = 0; d < ctx->g_world; ++d) {
            args.peer_re[d] = ctx->g_re[d];
            args.peer_im[d] = ctx->g_im[d];
            args.peer_flag[d] = ctx->g_flag[d];
        }
        args.n_peers = ctx->g_world;
        args.my_rank = ctx->g_rank;
        args.gather_seq = ctx->g_seq + 1;   // committed after the launch
        args.gather_elems = ctx->g_elems;
        args.done_counter = ctx->d_done;
    } else if (direct) {
        args.out_re = out_re;
        args.out_im = out_im;
    } else {
        args.out_re = ctx->h_out;
        args.out_im = ctx->h_out + out_elems;
    }

    args.dump = nullptr;
    if (flags & kFlagDumpReplica) {
        // debug: the DUMP instantiation of the hot kernel records the chip-table index of every replica entry
        if (n_periods != 1 || n_sats != 1 || use_raw || !(plan.help || dump_kernel_available(plan.A, plan.L)))
            return fail(ctx, GAT_ERR_UNSUPPORTED, "replica dump: one period, one channel, FP32 planes, antenna x tap class (1,3) (16,3) (8,5) (4,11)");
        const size_t n_dump = static_cast<size_t>(args.tiles_per_job + 1) * args.dump_stride;
        rc = ensure_device(ctx, ctx->d_dbg, ctx->d_dbg_cap, n_dump, false);
        if (rc) return rc;
        GAT_CUDA(ctx, cudaMemsetAsync(ctx->d_dbg, 0xFF, n_dump * sizeof(int32_t), ctx->stream));
        args.dump = reinterpret_cast<uint32_t *>(ctx->d_dbg);
        ctx->dump_tiles = args.tiles_per_job;
        ctx->dump_stride = args.dump_stride;
        ctx->dump_tile_len = args.tile_len;
        ctx->dump_aligned_start = args.aligned_start;
    }
    args.timeline = nullptr;
    if (ctx->timeline_on) {
        rc = ensure_device(ctx, ctx->d_timeline, ctx->timeline_cap, static_cast<size_t>(plan.grid) * 16, false);
        if (rc) return rc;
        GAT_CUDA(ctx, cudaMemsetAsync(ctx->d_timeline, 0, static_cast<size_t>(plan.grid) * 16 * sizeof(unsigned long long), ctx->stream));
        args.timeline = ctx->d_timeline;
        ctx->timeline_ctas = plan.grid;
    }
    if (flags & kFlagResidentPlan) {
        // gat_resident_begin: everything is validated, planned and marshalled -- the launch is the resident kernel's
        if (use_raw || sharded || plan.dump) return fail(ctx, GAT_ERR_UNSUPPORTED, "resident sessions run on FP32 planes of plain slots");
        ctx->res.plan = plan;
        ctx->res.args = args;
        ctx->res.n_ants = M;
        return GAT_OK;
    }
    if (ctx->timing) GAT_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    cudaError_t e;
    {
        NvtxRange nvtx_launch("correlate_kernel");
        e = launch_correlate(plan, args, ctx->stream);
    }
    if (e != cudaSuccess) return cuda_fail(ctx, e, "correlate kernel launch");
    ctx->barrier_count = barrier_target;
    if (gather) ctx->g_seq = args.gather_seq;
    if (stg) GAT_CUDA(ctx, cudaEventRecord(stg->consumed, ctx->stream));
    if (ctx->timing) GAT_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    ctx->launches += 1;
    ctx->info.kernels_launched = 1;

    if (host_direct) {
        GAT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        std::memcpy(out_re, ctx->h_out, out_elems * sizeof(float));
        std::memcpy(out_im, ctx->h_out + out_elems, out_elems * sizeof(float));
    }
    if (ctx->timing) {
        GAT_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
        float ms = 0.f;
        GAT_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        ctx->info.last_kernel_ms = ms;
    }
    return GAT_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// exported functions
// ------------------------------------------------------------------------------------------
extern "C" {

int gat_version(void) { return GAT_VERSION; }

const char *gat_status_string(int status)
{
    switch (status) {
    case GAT_OK: return "ok";
    case GAT_ERR_INVALID: return "invalid argument";
    case GAT_ERR_CUDA: return "CUDA error";
    case GAT_ERR_UNSUPPORTED: return "unsupported shape";
    case GAT_ERR_ALIGNMENT: return "misaligned device signal";
    case GAT_ERR_NO_CODES: return "no chip table";
    case GAT_ERR_NO_SIGNAL: return "no signal bound";
    case GAT_ERR_NO_DEVICE: return "no usable CUDA device";
    default: return "unknown status";
    }
}

int gat_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return GAT_ERR_NO_DEVICE;
    }
    return n;
}

int gat_create(gat_ctx **out, int device_id)
{
    if (!out) return GAT_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n < 1) {
        cudaGetLastError();
        return GAT_ERR_NO_DEVICE;
    }
    if (device_id < 0 || device_id >= n) return GAT_ERR_INVALID;
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, device_id) != cudaSuccess) return GAT_ERR_CUDA;
    if (prop.major != 10) return GAT_ERR_NO_DEVICE;  // sm_100a cubin only; no fallback path exists
    gat_ctx *ctx = new (std::nothrow) gat_ctx();
    if (!ctx) return GAT_ERR_INVALID;
    ctx->device = device_id;
    ctx->n_sm = prop.multiProcessorCount;
    if (cudaSetDevice(device_id) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->param_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess ||
        configure_kernels() != cudaSuccess || configure_tc_kernel() != cudaSuccess) {
        cudaGetLastError();
        delete ctx;
        return GAT_ERR_CUDA;
    }
    ctx->stream = ctx->own_stream;
    *out = ctx;
    return GAT_OK;
}

int gat_set_stream(gat_ctx *ctx, void *cuda_stream)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    GAT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stream = static_cast<cudaStream_t>(cuda_stream);   // NULL == legacy default stream
    return GAT_OK;
}

int gat_use_own_stream(gat_ctx *ctx)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    GAT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stream = ctx->own_stream;
    return GAT_OK;
}

int gat_destroy(gat_ctx *ctx)
{
    if (!ctx) return GAT_ERR_INVALID;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto &kv : ctx->slots) {
        SignalSlot &s = kv.second;
        if (s.owned && s.re) cudaFree(s.re);
        if (s.raw) cudaFree(s.raw);
        if (s.peer_base) cudaIpcCloseMemHandle(s.peer_base);
    }
    for (auto &c : ctx->codes)
        if (c.d_chips) cudaFree(c.d_chips);
    for (auto &s : ctx->stg) {
        if (s.h) cudaFreeHost(s.h);
        if (s.d) cudaFree(s.d);
        if (s.done) cudaEventDestroy(s.done);
        if (s.consumed) cudaEventDestroy(s.consumed);
    }
    if (ctx->d_partials) cudaFree(ctx->d_partials);
    if (ctx->d_barrier) cudaFree(ctx->d_barrier);
    if (ctx->d_out) cudaFree(ctx->d_out);
    if (ctx->h_out) cudaFreeHost(ctx->h_out);
    if (ctx->d_dbg) cudaFree(ctx->d_dbg);
    if (ctx->d_raw) cudaFree(ctx->d_raw);
    if (ctx->d_timeline) cudaFree(ctx->d_timeline);
    if (ctx->d_ing_out) cudaFree(ctx->d_ing_out);
    for (int b = 0; b < kIngestDepth; ++b)
        if (ctx->d_ing_stage[b]) cudaFree(ctx->d_ing_stage[b]);
    for (int b = 0; b < kIngestDepth; ++b) {
        if (ctx->ing_ready[b]) cudaEventDestroy(ctx->ing_ready[b]);
        if (ctx->ing_free[b]) cudaEventDestroy(ctx->ing_free[b]);
    }
    gat_resident_end(ctx);
    gat_gather_destroy(ctx);
    gat_ring_destroy(ctx);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    if (ctx->param_stream) cudaStreamDestroy(ctx->param_stream);
    delete ctx;
    return GAT_OK;
}

const char *gat_last_error(gat_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int gat_sync(gat_ctx *ctx)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    GAT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->ring.local) GAT_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));   // ingest queued by gat_ring_upload*
    return GAT_OK;
}

void *gat_stream(gat_ctx *ctx) { return ctx ? static_cast<void *>(ctx->stream) : nullptr; }

int gat_set_codes(gat_ctx *ctx, int system_id, const int8_t *chips, int code_len, int n_prn)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (system_id < 0 || system_id >= GAT_MAX_SYSTEMS || !chips || code_len < 1 || n_prn < 1)
        return fail(ctx, GAT_ERR_INVALID, "bad chip table arguments");
    const size_t n = static_cast<size_t>(code_len) * n_prn;
    for (size_t i = 0; i < n; ++i)
        if (chips[i] != 1 && chips[i] != -1) return fail(ctx, GAT_ERR_INVALID, "chips must be +1 or -1");
    CodeTable &t = ctx->codes[system_id];
    GAT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (t.d_chips) GAT_CUDA(ctx, cudaFree(t.d_chips));
    t = CodeTable{};
    // device layout: one zero-padded, 16-byte aligned column per PRN (vector loads into smem)
    const int stride = (code_len + kCodeColAlign - 1) / kCodeColAlign * kCodeColAlign;
    std::vector<int8_t> padded(static_cast<size_t>(stride) * n_prn, 0);
    for (int p = 0; p < n_prn; ++p)
        std::memcpy(padded.data() + static_cast<size_t>(p) * stride, chips + static_cast<size_t>(p) * code_len, code_len);
    GAT_CUDA(ctx, cudaMalloc(reinterpret_cast<void **>(&t.d_chips), padded.size()));
    GAT_CUDA(ctx, cudaMemcpy(t.d_chips, padded.data(), padded.size(), cudaMemcpyHostToDevice));
    t.code_len = code_len;
    t.n_prn = n_prn;
    t.col_stride = stride;
    // the built-in ids carry their ICD chip rates; any other table needs gat_set_code_frequency before gat_gen_signal
    t.code_freq_hz = (system_id == GAT_GPSL1 && code_len == 1023) ? 1.023e6 : (system_id == GAT_GPSL5 && code_len == 10230) ? 10.23e6 : 0.0;
    return GAT_OK;
}

int gat_set_code_frequency(gat_ctx *ctx, int system_id, double code_freq_hz)
{
    if (!ctx) return GAT_ERR_INVALID;
    if (system_id < 0 || system_id >= GAT_MAX_SYSTEMS || !ctx->codes[system_id].d_chips)
        return fail(ctx, GAT_ERR_NO_CODES, "no chip table set for this system");
    if (!(code_freq_hz > 0.0) || !std::isfinite(code_freq_hz)) return fail(ctx, GAT_ERR_INVALID, "code frequency must be positive");
    ctx->codes[system_id].code_freq_hz = code_freq_hz;
    return GAT_OK;
}

namespace {
// copy [n_ants x ld] planes into ctx-owned, padded storage of `slot` on `stream` (the ctx stream, or the ingest stream)
int upload_planes(gat_ctx *ctx, int slot, const float *re, const float *im, int n_samples, int n_ants, int ld, int src_is_device,
                  cudaStream_t stream)
{
    if (!re || !im || n_samples < 1 || n_ants < 1 || n_ants > kMaxAnts || ld < n_samples)
        return fail(ctx, GAT_ERR_INVALID, "bad signal arguments");
    SignalSlot *s = slot_for(ctx, slot);
    if (!s) return fail(ctx, GAT_ERR_INVALID, "slot out of range");
    if (!s->parts.empty()) return fail(ctx, GAT_ERR_INVALID, "slot belongs to the signal ring (use gat_ring_upload)");
    int rc = GAT_OK;
    if (!s->owned && s->re) {   // drop a zero-copy binding
        rc = free_planes(ctx, *s);
        if (rc) return rc;
    }
    const cudaMemcpyKind kind = src_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (ld % 4 == 0) {
        // same layout on both sides: one flat copy per plane
        rc = own_slot(ctx, *s, n_samples, n_ants, ld);
        if (rc) return rc;
        const size_t bytes = (static_cast<size_t>(ld) * (n_ants - 1) + n_samples) * sizeof(float);
        GAT_CUDA(ctx, cudaMemcpyAsync(s->re, re, bytes, kind, stream));
        GAT_CUDA(ctx, cudaMemcpyAsync(s->im, im, bytes, kind, stream));
    } else {
        const int64_t dld = (static_cast<int64_t>(n_samples) + 3) & ~3LL;
        rc = own_slot(ctx, *s, n_samples, n_ants, dld);
        if (rc) return rc;
        GAT_CUDA(ctx, cudaMemcpy2DAsync(s->re, dld * sizeof(float), re, static_cast<size_t>(ld) * sizeof(float),
                                        static_cast<size_t>(n_samples) * sizeof(float), n_ants, kind, stream));
        GAT_CUDA(ctx, cudaMemcpy2DAsync(s->im, dld * sizeof(float), im, static_cast<size_t>(ld) * sizeof(float),
                                        static_cast<size_t>(n_samples) * sizeof(float), n_ants, kind, stream));
    }
    return GAT_OK;
}
}  // namespace

int gat_upload_signal(gat_ctx *ctx, int slot, const float *re, const float *im, int n_samples, int n_ants, int ld,
                      int src_is_device)
{
    NvtxRange nvtx_call("gat_upload_signal");
    int rc = check_ctx(ctx);
    if (rc) return rc;
    return upload_planes(ctx, slot, re, im, n_samples, n_ants, ld, src_is_device, ctx->stream);
}

namespace {
int upload_sc(gat_ctx *ctx, int slot, const void *iq, int bytes_per_component, int n_samples, int n_ants, int ld, float scale,
              int src_is_device)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (!iq || n_samples < 1 || n_ants < 1 || n_ants > kMaxAnts || ld < n_samples || !std::isfinite(scale))
        return fail(ctx, GAT_ERR_INVALID, "bad signal arguments");
    SignalSlot *s = slot_for(ctx, slot);
    if (!s) return fail(ctx, GAT_ERR_INVALID, "slot out of range");
    if (!s->owned && s->re) {   // drop a zero-copy binding
        rc = free_planes(ctx, *s);
        if (rc) return rc;
    }
    const cudaMemcpyKind kind = src_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (bytes_per_component == 2) {
        // keep the raw words in the slot (rows padded to 4 samples); FP32 planes are produced only if a later
        // call needs them (gat_correlate with many channels per block, gat_download_signal, ...)
        const int64_t rld = (static_cast<int64_t>(n_samples) + 3) & ~3LL;
        const size_t need = static_cast<size_t>(rld) * n_ants;
        if (s->raw_cap < need) {
            GAT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            if (s->raw) GAT_CUDA(ctx, cudaFree(s->raw));
            s->raw = nullptr;
            s->raw_cap = 0;
            GAT_CUDA(ctx, cudaMalloc(reinterpret_cast<void **>(&s->raw), need * 4));
            s->raw_cap = need;
        }
        GAT_CUDA(ctx, cudaMemcpy2DAsync(s->raw, static_cast<size_t>(rld) * 4, iq, static_cast<size_t>(ld) * 4,
                                        static_cast<size_t>(n_samples) * 4, n_ants, kind, ctx->stream));
        const bool same = s->raw_valid && s->raw_ld == rld && s->n_samples == n_samples && s->n_ants == n_ants;
        // the FP32-plane descriptors describe the OLD shape: a later ensure_planes must not reuse them
        if (s->n_samples != n_samples || s->n_ants != n_ants) s->maps_valid = false;
        s->raw_ld = rld;
        s->n_samples = n_samples;
        s->n_ants = n_ants;
        s->raw_scale = scale;
        s->raw_valid = true;
        s->planes_valid = false;
        return same ? GAT_OK : encode_raw_map(ctx, *s);
    }
    rc = own_slot(ctx, *s, n_samples, n_ants, (static_cast<int64_t>(n_samples) + 3) & ~3LL);
    if (rc) return rc;
    const void *d_src = iq;
    if (!src_is_device) {
        const size_t bytes = (static_cast<size_t>(ld) * (n_ants - 1) + n_samples) * 2 * bytes_per_component;
        rc = ensure_device(ctx, ctx->d_raw, ctx->d_raw_cap, bytes, false);
        if (rc) return rc;
        GAT_CUDA(ctx, cudaMemcpyAsync(ctx->d_raw, iq, bytes, kind, ctx->stream));
        d_src = ctx->d_raw;
    }
    cudaError_t e = launch_expand_sc(d_src, bytes_per_component, ld, s->re, s->im, s->ld, n_samples, n_ants, scale, ctx->stream);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "expand_sc launch");
    ctx->launches += 1;
    return GAT_OK;
}
}  // namespace

int gat_upload_signal_sc16(gat_ctx *ctx, int slot, const int16_t *iq, int n_samples, int n_ants, int ld, float scale, int src_is_device)
{
    return upload_sc(ctx, slot, iq, 2, n_samples, n_ants, ld, scale, src_is_device);
}

int gat_upload_signal_sc8(gat_ctx *ctx, int slot, const int8_t *iq, int n_samples, int n_ants, int ld, float scale, int src_is_device)
{
    return upload_sc(ctx, slot, iq, 1, n_samples, n_ants, ld, scale, src_is_device);
}

int gat_bind_signal(gat_ctx *ctx, int slot, const float *d_re, const float *d_im, int n_samples, int n_ants, int ld)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (!d_re || !d_im || n_samples < 1 || n_ants < 1 || n_ants > kMaxAnts || ld < n_samples)
        return fail(ctx, GAT_ERR_INVALID, "bad signal arguments");
    if ((reinterpret_cast<uintptr_t>(d_re) & 15u) || (reinterpret_cast<uintptr_t>(d_im) & 15u) || (n_ants > 1 && ld % 4 != 0))
        return fail(ctx, GAT_ERR_ALIGNMENT, "zero-copy planes must be 16-byte aligned with ld % 4 == 0 (use gat_upload_signal)");
    // the bulk copies read whole 16-byte groups: the last group of the last row must exist
    if (n_ants == 1 && ld % 4 != 0 && ((n_samples + 3) & ~3) > ld)
        return fail(ctx, GAT_ERR_ALIGNMENT, "single-antenna zero-copy needs ld >= roundup4(n_samples)");
    SignalSlot *s = slot_for(ctx, slot);
    if (!s) return fail(ctx, GAT_ERR_INVALID, "slot out of range");
    rc = release_slot(ctx, *s);
    if (rc) return rc;
    s->re = const_cast<float *>(d_re);
    s->im = const_cast<float *>(d_im);
    s->ld = ld;
    s->n_samples = n_samples;
    s->n_ants = n_ants;
    s->owned = false;
    s->planes_valid = true;
    return encode_slot_maps(ctx, *s);
}

namespace {
struct SlotDesc {                    // the opaque GAT_SLOT_DESC_BYTES blob of gat_slot_export / gat_slot_import
    cudaIpcMemHandle_t handle;       // of the allocation holding both planes
    uint64_t im_offset_bytes;
    int64_t ld;
    int32_t n_samples, n_ants;
    uint32_t magic, reserved;
};
static_assert(sizeof(SlotDesc) == GAT_SLOT_DESC_BYTES, "slot descriptor layout");
constexpr uint32_t kSlotMagic = 0x47415453u;
}  // namespace

int gat_slot_export(gat_ctx *ctx, int slot, unsigned char *desc_out)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    SignalSlot *sp = find_slot(ctx, slot);
    if (!desc_out || !sp || !(sp->planes_valid || sp->raw_valid)) return fail(ctx, GAT_ERR_NO_SIGNAL, "slot has no signal");
    SignalSlot &s = *sp;
    rc = ensure_planes(ctx, s);
    if (rc) return rc;
    if (!s.owned) return fail(ctx, GAT_ERR_INVALID, "only ctx-owned slots (gat_upload_signal*, gat_gen_signal) can be exported");
    SlotDesc d{};
    GAT_CUDA(ctx, cudaIpcGetMemHandle(&d.handle, s.re));
    d.im_offset_bytes = static_cast<uint64_t>(s.cap_floats) * sizeof(float);
    d.ld = s.ld;
    d.n_samples = s.n_samples;
    d.n_ants = s.n_ants;
    d.magic = kSlotMagic;
    std::memcpy(desc_out, &d, sizeof(d));
    return GAT_OK;
}

int gat_slot_import(gat_ctx *ctx, int slot, const unsigned char *desc)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (!desc) return fail(ctx, GAT_ERR_INVALID, "null descriptor");
    SlotDesc d;
    std::memcpy(&d, desc, sizeof(d));
    if (d.magic != kSlotMagic || d.n_samples < 1 || d.n_ants < 1 || d.n_ants > kMaxAnts || d.ld < d.n_samples || d.ld % 4 != 0)
        return fail(ctx, GAT_ERR_INVALID, "not a slot descriptor");
    SignalSlot *s = slot_for(ctx, slot);
    if (!s) return fail(ctx, GAT_ERR_INVALID, "slot out of range");
    rc = release_slot(ctx, *s);
    if (rc) return rc;
    void *base = nullptr;
    GAT_CUDA(ctx, cudaIpcOpenMemHandle(&base, d.handle, cudaIpcMemLazyEnablePeerAccess));
    s->peer_base = base;
    s->re = static_cast<float *>(base);
    s->im = reinterpret_cast<float *>(static_cast<unsigned char *>(base) + d.im_offset_bytes);
    s->ld = d.ld;
    s->n_samples = d.n_samples;
    s->n_ants = d.n_ants;
    s->owned = false;
    s->planes_valid = true;
    return encode_slot_maps(ctx, *s);
}

int gat_gen_signal(gat_ctx *ctx, int slot, int system_id, int prn, double carrier_freq_hz, double fs_hz,
                   double start_code_phase, double start_carrier_phase_rad, int n_samples, int n_ants,
                   double ant_phase_step_rad, double noise_sigma, uint64_t seed, int superpose)
{
    NvtxRange nvtx_call("gat_gen_signal");
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (system_id < 0 || system_id >= GAT_MAX_SYSTEMS || !ctx->codes[system_id].d_chips)
        return fail(ctx, GAT_ERR_NO_CODES, "no chip table for gen_signal");
    const CodeTable &t = ctx->codes[system_id];
    if (prn < 1 || prn > t.n_prn || n_samples < 1 || n_ants < 1 || n_ants > kMaxAnts || !(fs_hz > 0.0))
        return fail(ctx, GAT_ERR_INVALID, "bad gen_signal arguments");
    SignalSlot *s = slot_for(ctx, slot);
    if (!s) return fail(ctx, GAT_ERR_INVALID, "slot out of range");
    if (superpose) {
        if (!(s->planes_valid || s->raw_valid) || s->n_samples != n_samples || s->n_ants != n_ants)
            return fail(ctx, GAT_ERR_NO_SIGNAL, "superpose needs an existing slot of the same shape");
        rc = ensure_planes(ctx, *s);
        if (rc) return rc;
    } else if (!(s->planes_valid && s->re && s->n_samples == n_samples && s->n_ants == n_ants)) {
        // (an owned slot, or caller planes bound with gat_bind_signal, of the right shape is generated into in place)
        if (!s->owned && s->re) {
            rc = free_planes(ctx, *s);
            if (rc) return rc;
        }
        rc = own_slot(ctx, *s, n_samples, n_ants, (static_cast<int64_t>(n_samples) + 3) & ~3LL);
        if (rc) return rc;
    }
    // another process's memory (gat_slot_import) and the ring's shares are inputs only: never written from here
    if (s->peer_base || !s->parts.empty())
        return fail(ctx, GAT_ERR_INVALID, "gat_gen_signal cannot write into an imported or ring slot");
    s->raw_valid = false;   // the planes are about to change
    const double code_freq = t.code_freq_hz;
    if (!(code_freq > 0.0))
        return fail(ctx, GAT_ERR_INVALID, "no chip rate known for this table: call gat_set_code_frequency first");
    cudaError_t e = launch_gen_signal(s->re, s->im, s->ld, t.d_chips + static_cast<size_t>(prn - 1) * t.col_stride, t.code_len,
                                      code_freq / fs_hz, carrier_freq_hz, fs_hz, start_code_phase, start_carrier_phase_rad,
                                      n_samples, n_ants, ant_phase_step_rad, noise_sigma, seed, superpose, ctx->stream);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "gen_signal launch");
    ctx->launches += 1;
    return GAT_OK;
}

int gat_download_signal(gat_ctx *ctx, int slot, float *re, float *im)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    SignalSlot *sp = find_slot(ctx, slot);
    if (!re || !im || !sp || !(sp->planes_valid || sp->raw_valid)) return fail(ctx, GAT_ERR_NO_SIGNAL, "slot has no signal");
    rc = ensure_planes(ctx, *sp);
    if (rc) return rc;
    const SignalSlot &s = *sp;
    const size_t w = static_cast<size_t>(s.n_samples) * sizeof(float);
    GAT_CUDA(ctx, cudaMemcpy2DAsync(re, w, s.re, s.ld * sizeof(float), w, s.n_ants, cudaMemcpyDeviceToHost, ctx->stream));
    GAT_CUDA(ctx, cudaMemcpy2DAsync(im, w, s.im, s.ld * sizeof(float), w, s.n_ants, cudaMemcpyDeviceToHost, ctx->stream));
    GAT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GAT_OK;
}

int gat_correlate(gat_ctx *ctx, int slot, int n_sats, const gat_channel *channels, double fs_hz,
                  const int32_t *sample_shifts, int n_taps, int start_sample, int n_samples, float *out_re,
                  float *out_im, int out_is_device, unsigned flags)
{
    const int32_t s = slot;
    return correlate_impl(ctx, 1, &s, n_sats, channels, fs_hz, sample_shifts, n_taps, start_sample, n_samples, out_re,
                          out_im, out_is_device, flags);
}

int gat_correlate_batch(gat_ctx *ctx, int n_periods, const int32_t *slots, int n_sats, const gat_channel *channels,
                        double fs_hz, const int32_t *sample_shifts, int n_taps, int start_sample, int n_samples,
                        float *out_re, float *out_im, int out_is_device, unsigned flags)
{
    return correlate_impl(ctx, n_periods, slots, n_sats, channels, fs_hz, sample_shifts, n_taps, start_sample, n_samples,
                          out_re, out_im, out_is_device, flags);
}

// ------------------------------------------------------------------------------------------
// resident sessions (gat_resident.cu): one call + synchronisation per block without a kernel launch
// ------------------------------------------------------------------------------------------
namespace {

int resident_launch(gat_ctx *ctx, uint32_t first_seq)
{
    Resident &r = ctx->res;
    // the relay cells may hold the exit notice of an earlier launch under this very sequence number
    GAT_CUDA(ctx, cudaMemsetAsync(r.d_relay, 0, kResMaxCells * 16, r.stream));
    r.ctl.first_seq = first_seq;
    r.args.barrier_target = ctx->barrier_count + static_cast<unsigned int>(r.plan.grid);
    cudaError_t e = launch_resident(r.plan, r.args, r.ctl, r.smem, r.stream);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "resident kernel launch");
    r.launched = true;
    ctx->launches += 1;
    return GAT_OK;
}

void resident_write_command(Resident &r, uint32_t seq, const uint32_t *words, int n_words)
{
    volatile uint32_t *cmd = r.h_cmd;
    const int n_cells = r.ctl.n_cells;
    for (int c = 0; c < n_cells; ++c)
        for (int j = 0; j < 3; ++j) cmd[4 * c + j] = (3 * c + j < n_words) ? words[3 * c + j] : 0u;
    // data before sequence words (x86 stores are observed in program order; the fence keeps the compiler honest)
    std::atomic_thread_fence(std::memory_order_seq_cst);
    for (int c = 0; c < n_cells; ++c) cmd[4 * c + 3] = seq;
    std::atomic_thread_fence(std::memory_order_seq_cst);
}

}  // namespace

int gat_resident_begin(gat_ctx *ctx, const int32_t *slots, int n_slots, int n_sats, const gat_channel *channels, double fs_hz,
                       const int32_t *sample_shifts, int n_taps, int start_sample, int n_samples)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    Resident &r = ctx->res;
    if (r.active) return fail(ctx, GAT_ERR_INVALID, "a resident session is already open");
    if (!slots || n_slots < 1 || n_slots > 4096 || !channels) return fail(ctx, GAT_ERR_INVALID, "bad arguments");
    if (n_sats < 1 || n_sats > kResMaxSats)
        return fail(ctx, GAT_ERR_UNSUPPORTED, "a resident command carries 1.." + std::to_string(kResMaxSats) + " channels");
    // plan + marshal through the ordinary call path (validation, kernel class, shared-memory carve-up), no launch
    std::vector<float> dummy(2 * static_cast<size_t>(n_sats) * n_taps * kMaxAnts);
    rc = correlate_impl(ctx, 1, &slots[0], n_sats, channels, fs_hz, sample_shifts, n_taps, start_sample, n_samples, dummy.data(),
                        dummy.data() + dummy.size() / 2, 0, kFlagResidentPlan);
    if (rc) return rc;
    if (!resident_kernel_available(r.plan.A, r.plan.L, r.plan.help))
        return fail(ctx, GAT_ERR_UNSUPPORTED, "no resident instantiation for this (antennas, taps) class: 1 / 4 / 16 antennas with <= 3 or 7 taps, 16 with 11");
    const int M = r.n_ants;
    std::vector<PeriodDev> maps(n_slots);
    for (int i = 0; i < n_slots; ++i) {
        SignalSlot *sl = find_slot(ctx, slots[i]);
        if (!slot_has_signal(sl) || !sl->parts.empty()) return fail(ctx, GAT_ERR_NO_SIGNAL, "slot " + std::to_string(slots[i]) + " has no (plain) signal");
        if (sl->n_ants != M || static_cast<int64_t>(start_sample) + n_samples > sl->n_samples)
            return fail(ctx, GAT_ERR_INVALID, "all slots of a resident session must hold the range and the same antenna count");
        rc = ensure_planes(ctx, *sl);
        if (rc) return rc;
        if (!sl->maps_valid) return fail(ctx, GAT_ERR_NO_SIGNAL, "slot has no TMA descriptor");
        maps[i] = sl->maps;
    }
    GAT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));     // uploads and expansions queued so far are done
    if (!r.stream) GAT_CUDA(ctx, cudaStreamCreateWithFlags(&r.stream, cudaStreamNonBlocking));
    if (!r.h_cmd) {
        GAT_CUDA(ctx, cudaHostAlloc(reinterpret_cast<void **>(&r.h_cmd), kResMaxCells * 16 + 64, cudaHostAllocMapped));
        r.h_flag = r.h_cmd + kResMaxCells * 4;
        GAT_CUDA(ctx, cudaMalloc(reinterpret_cast<void **>(&r.d_relay), kResMaxCells * 16 + 64));
    }
    std::memset(r.h_cmd, 0, kResMaxCells * 16 + 64);
    GAT_CUDA(ctx, cudaMemset(r.d_relay, 0, kResMaxCells * 16 + 64));
    if (r.d_maps) GAT_CUDA(ctx, cudaFree(r.d_maps));
    r.d_maps = nullptr;
    GAT_CUDA(ctx, cudaMalloc(reinterpret_cast<void **>(&r.d_maps), sizeof(PeriodDev) * n_slots));
    GAT_CUDA(ctx, cudaMemcpy(r.d_maps, maps.data(), sizeof(PeriodDev) * n_slots, cudaMemcpyHostToDevice));
    // results: one 8-byte word {value bits, sequence number} per accumulator, straight into mapped host memory
    r.out_elems = static_cast<size_t>(n_sats) * n_taps * M;
    if (r.h_res) GAT_CUDA(ctx, cudaFreeHost(r.h_res));
    r.h_res = nullptr;
    GAT_CUDA(ctx, cudaHostAlloc(reinterpret_cast<void **>(&r.h_res), 2 * r.out_elems * sizeof(unsigned long long), cudaHostAllocMapped));
    std::memset(r.h_res, 0, 2 * r.out_elems * sizeof(unsigned long long));
    r.args.out_re = reinterpret_cast<float *>(r.h_res);
    r.args.out_im = reinterpret_cast<float *>(r.h_res + r.out_elems);
    r.args.n_peers = 0;
    r.args.use_inline = 0;
    r.args.periods = nullptr;
    r.args.sats = nullptr;
    r.args.flags = 0;
    r.ctl.cmd_host = reinterpret_cast<const uint4 *>(r.h_cmd);
    r.ctl.relay = r.d_relay;
    r.ctl.slot_maps = r.d_maps;
    r.ctl.n_cells = (4 + 16 * n_sats + 2) / 3;
    r.ctl.cmd_off = static_cast<int32_t>((r.plan.smem + 127) & ~static_cast<size_t>(127));
    r.ctl.idle_limit_ms = static_cast<uint32_t>(std::max(1, env_int("GAT_RESIDENT_IDLE_MS", 2000)));
    r.debug = env_int("GAT_RESIDENT_DEBUG", 0) != 0;
    r.ctl.stamps = r.debug ? reinterpret_cast<unsigned long long *>(r.h_flag + 4) : nullptr;
    r.dbg_host_ns = r.dbg_seen_to_body_ns = r.dbg_body_ns = 0;
    r.dbg_calls = 0;
    r.args.timeline = nullptr;
    if (r.debug) {
        rc = ensure_device(ctx, ctx->d_timeline, ctx->timeline_cap, static_cast<size_t>(r.plan.grid) * 16, false);
        if (rc) return rc;
        GAT_CUDA(ctx, cudaMemset(ctx->d_timeline, 0, static_cast<size_t>(r.plan.grid) * 16 * sizeof(unsigned long long)));
        r.args.timeline = ctx->d_timeline;
    }
    r.smem = static_cast<size_t>(r.ctl.cmd_off) + kResCmdSmemBytes;
    if (r.smem > 227 * 1024) return fail(ctx, GAT_ERR_UNSUPPORTED, "shape leaves no shared memory for the command area");
    static_assert(sizeof(SatDev) == 64, "command layout: 16 words per channel");
    static_assert(4 + 16 * kResMaxSats <= 3 * kResMaxCells && (4 + 16 * kResMaxSats) * 4 <= kResCmdSmemBytes, "command size");
    r.n_slots = n_slots;
    r.n_sats = n_sats;
    r.n_taps = n_taps;
    r.fs_hz = fs_hz;
    r.rep_len = std::max(1, r.args.visit_tiles) * r.args.tile_len;
    r.seq = 0;
    r.launched = false;
    r.active = true;
    rc = resident_launch(ctx, 1u);
    if (rc) r.active = false;
    return rc;
}

int gat_resident_correlate(gat_ctx *ctx, int slot_index, const gat_channel *channels, float *out_re, float *out_im)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    Resident &r = ctx->res;
    if (!r.active) return fail(ctx, GAT_ERR_INVALID, "no resident session (gat_resident_begin)");
    if (!channels || !out_re || !out_im || slot_index < 0 || slot_index >= r.n_slots) return fail(ctx, GAT_ERR_INVALID, "bad arguments");
    uint32_t words[4 + 16 * kResMaxSats];
    words[0] = kResOpCorrelate;
    words[1] = static_cast<uint32_t>(slot_index);
    words[2] = static_cast<uint32_t>(r.n_sats);
    words[3] = 0u;
    for (int k = 0; k < r.n_sats; ++k) {
        SatDev sd{};
        rc = fill_sat(ctx, channels[k], r.fs_hz, sd);
        if (rc) return rc;
        // the session's plan fixed the chip-table cache, the replica window and its wrap branch: a channel must fit them
        const long double need = static_cast<long double>(r.rep_len + r.args.span + 160) * static_cast<long double>(sd.nco_delta) +
                                 std::ldexp(1.0L, sd.nco_fp);
        const bool wrap_ok = !r.args.rep_single_wrap ||
                             static_cast<double>(r.rep_len + r.args.span + 160) * sd.code_ratio + 2.0 < static_cast<double>(sd.code_len);
        if (((sd.code_len + kCodeColAlign - 1) / kCodeColAlign * kCodeColAlign) > r.args.cache_stride || need >= std::ldexp(1.0L, 64) || !wrap_ok)
            return fail(ctx, GAT_ERR_UNSUPPORTED, "channel outside the envelope the resident session was planned for (code length / code rate)");
        std::memcpy(&words[4 + 16 * k], &sd, sizeof(sd));
    }
    const uint32_t seq = ++r.seq;
    if (r.launched && cudaStreamQuery(r.stream) == cudaSuccess) r.launched = false;    // ended on its idle limit
    if (!r.launched) {
        rc = resident_launch(ctx, seq);
        if (rc) return rc;
        r.relaunches += 1;
    }
    resident_write_command(r, seq, words, 4 + 16 * r.n_sats);
    // every accumulator arrives as {value, seq}: the command is complete when all of them carry this sequence number
    const volatile unsigned long long *res = r.h_res;
    const size_t n_res = 2 * r.out_elems;
    size_t next = 0;                 // elements [0, next) have arrived
    const auto t0 = std::chrono::steady_clock::now();
    for (uint32_t spins = 1;; ++spins) {
        while (next < n_res && static_cast<uint32_t>(res[next] >> 32) == seq) ++next;
        if (next == n_res) break;
        if ((spins & 0xFFFu) == 0) {
            const cudaError_t q = cudaStreamQuery(r.stream);
            if (q == cudaSuccess && static_cast<uint32_t>(res[next] >> 32) != seq) {
                // the kernel left on its idle limit just as this command was written: start it again for this command
                r.launched = false;
                rc = resident_launch(ctx, seq);
                if (rc) return rc;
                r.relaunches += 1;
            } else if (q != cudaErrorNotReady && q != cudaSuccess) {
                r.launched = false;
                r.active = false;
                return cuda_fail(ctx, q, "resident kernel");
            }
            if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(20)) {
                r.active = false;
                return fail(ctx, GAT_ERR_CUDA, "resident kernel did not answer within 20 s");
            }
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    if (r.debug) {
        const volatile unsigned long long *st = reinterpret_cast<const volatile unsigned long long *>(r.h_flag + 4);
        r.dbg_host_ns += std::chrono::duration<double, std::nano>(std::chrono::steady_clock::now() - t0).count();
        r.dbg_seen_to_body_ns += static_cast<double>(st[1] - st[0]);
        r.dbg_body_ns += static_cast<double>(st[2] - st[1]);
        r.dbg_calls += 1;
    }
    ctx->barrier_count += static_cast<unsigned int>(r.plan.grid);
    for (size_t i = 0; i < r.out_elems; ++i) {
        const uint32_t vr = static_cast<uint32_t>(res[i]), vi = static_cast<uint32_t>(res[r.out_elems + i]);
        std::memcpy(&out_re[i], &vr, sizeof(float));
        std::memcpy(&out_im[i], &vi, sizeof(float));
    }
    ctx->info.kernels_launched = 0;
    return GAT_OK;
}

int gat_resident_end(gat_ctx *ctx)
{
    if (!ctx) return GAT_ERR_INVALID;
    Resident &r = ctx->res;
    int rc = GAT_OK;
    if (r.debug && r.dbg_calls)
        std::fprintf(stderr, "[gat resident] %llu calls: host write -> flag seen %.2f us; device: command seen -> body %.2f us, body (CTA 0) %.2f us\n",
                     static_cast<unsigned long long>(r.dbg_calls), r.dbg_host_ns / r.dbg_calls * 1e-3, r.dbg_seen_to_body_ns / r.dbg_calls * 1e-3,
                     r.dbg_body_ns / r.dbg_calls * 1e-3);
    const bool dbg_timeline = r.debug && r.dbg_calls && r.args.timeline;
    const unsigned long long dbg_t0 = (r.debug && r.h_flag) ? reinterpret_cast<const volatile unsigned long long *>(r.h_flag + 4)[0] : 0ull;
    r.debug = false;
    if (r.active && r.launched && cudaStreamQuery(r.stream) == cudaErrorNotReady) {
        const uint32_t words[4] = {kResOpExit, 0u, 0u, 0u};
        resident_write_command(r, ++r.seq, words, 4);
    }
    if (r.stream) {
        const cudaError_t e = cudaStreamSynchronize(r.stream);
        if (e != cudaSuccess) rc = cuda_fail(ctx, e, "resident kernel");
    }
    if (dbg_timeline && rc == GAT_OK) {
        // per-CTA stamps of the LAST command (GAT_STAMP slots of correlate_body), relative to CTA 0 seeing the command
        static const char *names[16] = {"c.entry", "c.setup", "c.first_tile", "c.last_tile", "c.published", "c.barrier", "c.exit", "",
                                        "p.entry", "p.setup", "p.cached", "p.first_issued", "p.all_issued", "", "", ""};
        const int G = r.plan.grid;
        std::vector<unsigned long long> tl(static_cast<size_t>(G) * 16);
        if (cudaMemcpy(tl.data(), ctx->d_timeline, tl.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost) == cudaSuccess)
            for (int sl = 0; sl < 16; ++sl) {
                std::vector<double> v;
                for (int b = 0; b < G; ++b)
                    if (tl[static_cast<size_t>(b) * 16 + sl]) v.push_back((static_cast<double>(tl[static_cast<size_t>(b) * 16 + sl]) - static_cast<double>(dbg_t0)) * 1e-3);
                if (v.empty()) continue;
                std::sort(v.begin(), v.end());
                std::fprintf(stderr, "[gat resident]   %-14s min %6.2f  median %6.2f  max %6.2f us after the command was seen\n", names[sl], v.front(),
                             v[v.size() / 2], v.back());
            }
    }
    r.active = false;
    r.launched = false;
    if (r.d_maps) cudaFree(r.d_maps);
    r.d_maps = nullptr;
    if (r.d_relay) cudaFree(r.d_relay);
    r.d_relay = nullptr;
    if (r.h_res) cudaFreeHost(r.h_res);
    r.h_res = nullptr;
    if (r.h_cmd) cudaFreeHost(r.h_cmd);
    r.h_cmd = nullptr;
    r.h_flag = nullptr;
    if (r.stream) cudaStreamDestroy(r.stream);
    r.stream = nullptr;
    return rc;
}

int gat_downconvert_and_correlate(gat_ctx *ctx, const float *h_re, const float *h_im, int ld, int n_ants, int n_sats,
                                  const gat_channel *channels, double fs_hz, const int32_t *sample_shifts, int n_taps,
                                  int start_sample, int n_samples, float *h_out_re, float *h_out_im, unsigned flags)
{
    if (!ctx) return GAT_ERR_INVALID;
    const int scratch_slot = 65535;   // (slots are a sparse map: a high id costs nothing)
    if (start_sample < 0 || n_samples < 1 || static_cast<int64_t>(start_sample) + n_samples > INT32_MAX)
        return fail(ctx, GAT_ERR_INVALID, "empty, negative or overflowing sample range");
    int rc = gat_upload_signal(ctx, scratch_slot, h_re, h_im, start_sample + n_samples, n_ants, ld, 0);
    if (rc) return rc;
    return gat_correlate(ctx, scratch_slot, n_sats, channels, fs_hz, sample_shifts, n_taps, start_sample, n_samples,
                         h_out_re, h_out_im, 0, flags);
}

// The CPU-style call over MANY periods with the host<->device traffic pipelined inside the library: the drop-in for a
// host loop of Tracking.downconvert_and_correlate! calls (src/benchmarks.jl:63-79) over consecutive 1 ms blocks whose
// channel parameters are known up front.  Chunks of kIngestChunk periods go H2D on the ingest stream into a ring of
// kIngestDepth staging buffers while the kernel of the previous chunk runs; two events per buffer order the streams.
int gat_ingest_correlate(gat_ctx *ctx, int n_periods, const float *const *h_re, const float *const *h_im, int ld, int n_ants,
                         int n_sats, const gat_channel *channels, double fs_hz, const int32_t *sample_shifts, int n_taps,
                         int start_sample, int n_samples, float *h_out_re, float *h_out_im, unsigned flags)
{
    NvtxRange nvtx_call("gat_ingest_correlate");
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (!h_re || !h_im || !channels || !sample_shifts || !h_out_re || !h_out_im || n_periods < 1 || n_sats < 1)
        return fail(ctx, GAT_ERR_INVALID, "null pointer or empty batch");
    if (flags & (GAT_ACCUMULATE | GAT_GATHER)) return fail(ctx, GAT_ERR_INVALID, "gat_ingest_correlate returns host results: no ACCUMULATE / GATHER");
    if (n_taps < 1 || n_taps > GAT_MAX_TAPS) return fail(ctx, GAT_ERR_UNSUPPORTED, "n_taps must be 1..11");
    if (start_sample < 0 || n_samples < 1 || static_cast<int64_t>(start_sample) + n_samples > INT32_MAX)
        return fail(ctx, GAT_ERR_INVALID, "empty, negative or overflowing sample range");
    for (int b = 0; b < kIngestDepth; ++b)
        if (!ctx->ing_ready[b]) {
            GAT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ing_ready[b], cudaEventDisableTiming));
            GAT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ing_free[b], cudaEventDisableTiming));
        }
    const size_t per_period = static_cast<size_t>(n_sats) * n_taps * n_ants;
    const size_t out_elems = per_period * n_periods;
    rc = ensure_device(ctx, ctx->d_ing_out, ctx->ing_out_cap, 2 * out_elems, false);
    if (rc) return rc;
    float *d_re = ctx->d_ing_out, *d_im = ctx->d_ing_out + out_elems;
    const int n_up = start_sample + n_samples;
    if (n_ants < 1 || n_ants > kMaxAnts || ld < n_up) return fail(ctx, GAT_ERR_INVALID, "bad signal arguments");
    // Staging: every ring buffer is ONE allocation holding a chunk's re planes followed by its im planes, and the chunk's
    // slots are zero-copy views into it.  When the caller's blocks are contiguous too (an array [P][n_ants][ld], the usual
    // case) a chunk crosses PCIe as two large copies instead of 32 small ones (54 instead of 51.7 GB/s from pinned memory).
    const int64_t dld = (ld % 4 == 0) ? ld : ((static_cast<int64_t>(n_up) + 3) & ~3LL);
    const size_t plane = static_cast<size_t>(dld) * n_ants;
    const size_t need = 2 * static_cast<size_t>(kIngestChunk) * plane;
    if (need > ctx->ing_stage_cap || ctx->ing_n != n_up || ctx->ing_m != n_ants || ctx->ing_ld != dld) {
        GAT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        GAT_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
        if (need > ctx->ing_stage_cap) {
            for (int b = 0; b < kIngestDepth; ++b) {
                if (ctx->d_ing_stage[b]) GAT_CUDA(ctx, cudaFree(ctx->d_ing_stage[b]));
                ctx->d_ing_stage[b] = nullptr;
            }
            ctx->ing_stage_cap = 0;
            for (int b = 0; b < kIngestDepth; ++b) GAT_CUDA(ctx, cudaMalloc(reinterpret_cast<void **>(&ctx->d_ing_stage[b]), need * sizeof(float)));
            ctx->ing_stage_cap = need;
        }
        for (int b = 0; b < kIngestDepth; ++b)
            for (int i = 0; i < kIngestChunk; ++i) {
                SignalSlot *s = slot_for(ctx, kIngestSlotBase + b * kIngestChunk + i);
                rc = release_slot(ctx, *s);
                if (rc) return rc;
                s->re = ctx->d_ing_stage[b] + static_cast<size_t>(i) * plane;
                s->im = ctx->d_ing_stage[b] + (static_cast<size_t>(kIngestChunk) + i) * plane;
                s->ld = dld;
                s->n_samples = n_up;
                s->n_ants = n_ants;
                s->owned = false;
                s->planes_valid = true;
                rc = encode_slot_maps(ctx, *s);
                if (rc) return rc;
            }
        ctx->ing_n = n_up;
        ctx->ing_m = n_ants;
        ctx->ing_ld = dld;
    }
    int32_t slot_ids[kIngestChunk];
    // nothing queued earlier on the ctx stream may still be reading the staging slots
    for (int b = 0; b < kIngestDepth; ++b) GAT_CUDA(ctx, cudaEventRecord(ctx->ing_free[b], ctx->stream));
    const size_t host_plane = static_cast<size_t>(ld) * n_ants;
    int chunk = 0;
    for (int p0 = 0; p0 < n_periods; p0 += kIngestChunk, ++chunk) {
        const int b = chunk % kIngestDepth;
        const int cnt = std::min(kIngestChunk, n_periods - p0);
        GAT_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ing_free[b], 0));      // the kernel that read buffer b is done
        bool contiguous = (dld == ld);
        for (int i = 0; i < cnt; ++i) {
            slot_ids[i] = kIngestSlotBase + b * kIngestChunk + i;
            if (!h_re[p0 + i] || !h_im[p0 + i]) return fail(ctx, GAT_ERR_INVALID, "null signal pointer");
            contiguous = contiguous && h_re[p0 + i] == h_re[p0] + i * host_plane && h_im[p0 + i] == h_im[p0] + i * host_plane;
        }
        float *s_re = ctx->d_ing_stage[b], *s_im = ctx->d_ing_stage[b] + static_cast<size_t>(kIngestChunk) * plane;
        if (contiguous) {
            GAT_CUDA(ctx, cudaMemcpyAsync(s_re, h_re[p0], cnt * plane * sizeof(float), cudaMemcpyHostToDevice, ctx->copy_stream));
            GAT_CUDA(ctx, cudaMemcpyAsync(s_im, h_im[p0], cnt * plane * sizeof(float), cudaMemcpyHostToDevice, ctx->copy_stream));
        } else {
            for (int i = 0; i < cnt; ++i) {
                GAT_CUDA(ctx, cudaMemcpy2DAsync(s_re + i * plane, dld * sizeof(float), h_re[p0 + i], static_cast<size_t>(ld) * sizeof(float),
                                                static_cast<size_t>(n_up) * sizeof(float), n_ants, cudaMemcpyHostToDevice, ctx->copy_stream));
                GAT_CUDA(ctx, cudaMemcpy2DAsync(s_im + i * plane, dld * sizeof(float), h_im[p0 + i], static_cast<size_t>(ld) * sizeof(float),
                                                static_cast<size_t>(n_up) * sizeof(float), n_ants, cudaMemcpyHostToDevice, ctx->copy_stream));
            }
        }
        GAT_CUDA(ctx, cudaEventRecord(ctx->ing_ready[b], ctx->copy_stream));
        GAT_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ing_ready[b], 0));
        rc = correlate_impl(ctx, cnt, slot_ids, n_sats, channels + static_cast<size_t>(p0) * n_sats, fs_hz, sample_shifts, n_taps,
                            start_sample, n_samples, d_re + per_period * p0, d_im + per_period * p0, 1, flags);
        if (rc) return rc;
        GAT_CUDA(ctx, cudaEventRecord(ctx->ing_free[b], ctx->stream));
    }
    GAT_CUDA(ctx, cudaMemcpyAsync(h_out_re, d_re, out_elems * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    GAT_CUDA(ctx, cudaMemcpyAsync(h_out_im, d_im, out_elems * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    GAT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GAT_OK;
}

// Page-lock caller memory (a Julia Array, a numpy buffer) so that the H2D copies of gat_upload_signal* /
// gat_ingest_correlate / gat_ring_upload* run asynchronously at full PCIe rate instead of through the driver's bounce buffers.
int gat_host_register(void *ptr, uint64_t bytes)
{
    if (!ptr || !bytes) return GAT_ERR_INVALID;
    cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
    if (e == cudaErrorHostMemoryAlreadyRegistered) {
        cudaGetLastError();
        return GAT_OK;
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        return GAT_ERR_CUDA;
    }
    return GAT_OK;
}

int gat_host_unregister(void *ptr)
{
    if (!ptr) return GAT_ERR_INVALID;
    if (cudaHostUnregister(ptr) != cudaSuccess) {
        cudaGetLastError();
        return GAT_ERR_CUDA;
    }
    return GAT_OK;
}

int gat_slot_shape(gat_ctx *ctx, int slot, int *n_samples_out, int *n_ants_out)
{
    if (!ctx) return GAT_ERR_INVALID;
    const SignalSlot *s = find_slot(ctx, slot);
    if (!slot_has_signal(s)) return fail(ctx, GAT_ERR_NO_SIGNAL, "slot " + std::to_string(slot) + " has no signal");
    if (n_samples_out) *n_samples_out = s->n_samples;
    if (n_ants_out) *n_ants_out = s->n_ants;
    return GAT_OK;
}

int gat_last_launch_info(gat_ctx *ctx, gat_launch_info *out)
{
    if (!ctx || !out) return GAT_ERR_INVALID;
    *out = ctx->info;
    return GAT_OK;
}

int gat_plan_probe(int n_sm, int max_ctas, int n_periods, int n_sats, int n_ants, int n_taps, const int32_t *sample_shifts,
                   int start_sample, int n_samples, double fs_hz, double code_freq_hz, int code_len, unsigned flags,
                   gat_launch_info *out, char *err, int err_cap)
{
    // the planner on a context that never touches CUDA: the same checks, instantiation choice and make_plan as correlate_impl
    gat_ctx probe;
    probe.n_sm = n_sm;
    probe.max_ctas = max_ctas;
    auto finish = [&](int rc) {
        if (err && err_cap > 0) {
            std::strncpy(err, probe.err.c_str(), static_cast<size_t>(err_cap) - 1);
            err[err_cap - 1] = 0;
        }
        return rc;
    };
    if (!sample_shifts || !out || n_sm < 1) return finish(fail(&probe, GAT_ERR_INVALID, "null pointer argument or n_sm < 1"));
    if (n_periods < 1 || n_sats < 1) return finish(fail(&probe, GAT_ERR_INVALID, "n_periods and n_sats must be >= 1"));
    if (n_taps < 1 || n_taps > GAT_MAX_TAPS) return finish(fail(&probe, GAT_ERR_UNSUPPORTED, "n_taps must be 1..11"));
    if (n_ants < 1 || n_ants > kMaxAnts) return finish(fail(&probe, GAT_ERR_UNSUPPORTED, "antenna count must be 1..32"));
    if (!(fs_hz > 0.0) || !(code_freq_hz > 0.0) || code_len < 1) return finish(fail(&probe, GAT_ERR_INVALID, "frequencies and code length must be positive"));
    if (start_sample < 0 || n_samples < 1) return finish(fail(&probe, GAT_ERR_INVALID, "empty or negative sample range"));
    for (int l = 1; l < n_taps; ++l)
        if (sample_shifts[l] < sample_shifts[l - 1]) return finish(fail(&probe, GAT_ERR_INVALID, "sample shifts must be ascending"));
    if (static_cast<int64_t>(sample_shifts[n_taps - 1]) - sample_shifts[0] > 4096)
        return finish(fail(&probe, GAT_ERR_UNSUPPORTED, "tap span > 4096 samples"));
    const bool f64 = (flags & GAT_CODE_PHASE_F64) != 0;
    const bool raw = (flags & GAT_PROBE_INT16) != 0 && !f64;
    int A, L, TG;
    choose_instantiation(n_taps, n_ants, raw, sample_shifts, A, L, TG);
    Shape shape{n_periods, n_sats, n_ants, L, start_sample, n_samples, sample_shifts, 0.0, 63, 0, f64, 1, raw};
    shape.A = A;
    shape.TG = TG;
    shape.n_taps = n_taps;
    shape.reserve_smem = (flags & GAT_PROBE_RESIDENT) ? kResCmdSmemBytes + 128 : 0;
    shape.max_ratio = code_freq_hz / fs_hz;
    shape.min_fp = nco_fixed_point(code_len);
    shape.max_delta = static_cast<int64_t>(std::floor(code_freq_hz * std::ldexp(1.0, shape.min_fp) / fs_hz));
    shape.max_code_len = shape.min_code_len = code_len;
    LaunchPlan plan{};
    CorrArgs args{};
    const int rc = make_plan(&probe, shape, plan, args);
    if (rc == GAT_OK) *out = probe.info;
    return finish(rc);
}

int gat_set_timing(gat_ctx *ctx, int enable)
{
    if (!ctx) return GAT_ERR_INVALID;
    ctx->timing = enable != 0;
    return GAT_OK;
}

int gat_set_max_ctas(gat_ctx *ctx, int max_ctas)
{
    if (!ctx || max_ctas < 0) return GAT_ERR_INVALID;
    ctx->max_ctas = max_ctas;
    return GAT_OK;
}

uint64_t gat_kernel_launch_count(gat_ctx *ctx) { return ctx ? ctx->launches : 0; }

int gat_gather_destroy(gat_ctx *ctx)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    cudaStreamSynchronize(ctx->stream);
    for (int d = 0; d < kMaxPeers; ++d) {
        if (ctx->g_opened[d]) cudaIpcCloseMemHandle(ctx->g_opened[d]);
        ctx->g_opened[d] = nullptr;
        ctx->g_re[d] = ctx->g_im[d] = nullptr;
        ctx->g_flag[d] = nullptr;
    }
    if (ctx->g_local) cudaFree(ctx->g_local);
    if (ctx->d_done) cudaFree(ctx->d_done);
    ctx->g_local = nullptr;
    ctx->d_done = nullptr;
    ctx->g_connected = false;
    ctx->g_world = 0;
    ctx->g_seq = 0;
    ctx->g_off = 0;
    return GAT_OK;
}

namespace {
void gather_views(unsigned char *base, int world, uint64_t elems, float **re, float **im, unsigned int **flag)
{
    const size_t plane = static_cast<size_t>(world) * elems * sizeof(float);
    *re = reinterpret_cast<float *>(base);
    *im = reinterpret_cast<float *>(base + plane);
    *flag = reinterpret_cast<unsigned int *>(base + 2 * plane);
}
}  // namespace

int gat_gather_create(gat_ctx *ctx, int world, int rank, uint64_t elems_per_rank, unsigned char *handle_out)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world || elems_per_rank < 1 || !handle_out)
        return fail(ctx, GAT_ERR_INVALID, "bad gather arguments (1 <= world <= 8)");
    rc = gat_gather_destroy(ctx);
    if (rc) return rc;
    const uint64_t elems = (elems_per_rank + 63) & ~static_cast<uint64_t>(63);   // 256-byte slices
    const size_t bytes = 2 * static_cast<size_t>(world) * elems * sizeof(float) + 256;
    GAT_CUDA(ctx, cudaMalloc(reinterpret_cast<void **>(&ctx->g_local), bytes));
    GAT_CUDA(ctx, cudaMemset(ctx->g_local, 0, bytes));
    GAT_CUDA(ctx, cudaMalloc(reinterpret_cast<void **>(&ctx->d_done), sizeof(unsigned int)));
    GAT_CUDA(ctx, cudaMemset(ctx->d_done, 0, sizeof(unsigned int)));
    cudaIpcMemHandle_t h;
    GAT_CUDA(ctx, cudaIpcGetMemHandle(&h, ctx->g_local));
    static_assert(sizeof(h) == GAT_IPC_HANDLE_BYTES, "IPC handle size");
    std::memcpy(handle_out, &h, sizeof(h));
    ctx->g_world = world;
    ctx->g_rank = rank;
    ctx->g_elems = elems;
    return GAT_OK;
}

int gat_gather_connect(gat_ctx *ctx, const unsigned char *handles)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (!handles || !ctx->g_local) return fail(ctx, GAT_ERR_INVALID, "gat_gather_create first");
    for (int d = 0; d < ctx->g_world; ++d) {
        unsigned char *base = ctx->g_local;
        if (d != ctx->g_rank) {
            cudaIpcMemHandle_t h;
            std::memcpy(&h, handles + static_cast<size_t>(d) * GAT_IPC_HANDLE_BYTES, sizeof(h));
            void *p = nullptr;
            GAT_CUDA(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
            ctx->g_opened[d] = p;
            base = static_cast<unsigned char *>(p);
        }
        gather_views(base, ctx->g_world, ctx->g_elems, &ctx->g_re[d], &ctx->g_im[d], &ctx->g_flag[d]);
    }
    ctx->g_connected = true;
    return GAT_OK;
}

int gat_gather_set_offset(gat_ctx *ctx, uint64_t elem_offset)
{
    if (!ctx) return GAT_ERR_INVALID;
    if (elem_offset >= ctx->g_elems && ctx->g_elems) return fail(ctx, GAT_ERR_INVALID, "gather offset beyond the slice");
    ctx->g_off = elem_offset;
    return GAT_OK;
}

int gat_set_sample_origin(gat_ctx *ctx, int origin)
{
    if (!ctx) return GAT_ERR_INVALID;
    if (ctx->res.active) return fail(ctx, GAT_ERR_INVALID, "a resident session is open");
    ctx->sample_origin_on = origin >= 0;
    ctx->sample_origin = origin >= 0 ? origin : 0;
    return GAT_OK;
}

int gat_gather_sum(gat_ctx *ctx, uint64_t n_elems, float *d_out_re, float *d_out_im)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (!ctx->g_connected || !d_out_re || !d_out_im) return fail(ctx, GAT_ERR_INVALID, "gather not connected, or null outputs");
    if (n_elems < 1 || n_elems > ctx->g_elems) return fail(ctx, GAT_ERR_INVALID, "more elements than a gather slice holds");
    const cudaError_t e = launch_sum_slices(ctx->g_re[ctx->g_rank], ctx->g_im[ctx->g_rank], ctx->g_elems, ctx->g_world, n_elems, d_out_re,
                                            d_out_im, ctx->stream);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "sum_slices_kernel");
    ctx->launches += 1;
    return GAT_OK;
}

int gat_gather_wait(gat_ctx *ctx)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (!ctx->g_connected) return fail(ctx, GAT_ERR_INVALID, "gather not connected");
    cudaError_t e = launch_gather_wait(nullptr, ctx->g_flag[ctx->g_rank], ctx->g_world, ctx->g_seq, ctx->stream);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "gather wait launch");
    ctx->launches += 1;
    return GAT_OK;
}

int gat_gather_read(gat_ctx *ctx, float *h_re, float *h_im)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (!ctx->g_connected || !h_re || !h_im) return fail(ctx, GAT_ERR_INVALID, "gather not connected");
    const size_t bytes = static_cast<size_t>(ctx->g_world) * ctx->g_elems * sizeof(float);
    GAT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    GAT_CUDA(ctx, cudaMemcpy(h_re, ctx->g_re[ctx->g_rank], bytes, cudaMemcpyDeviceToHost));
    GAT_CUDA(ctx, cudaMemcpy(h_im, ctx->g_im[ctx->g_rank], bytes, cudaMemcpyDeviceToHost));
    return GAT_OK;
}

int gat_set_timeline(gat_ctx *ctx, int enable)
{
    if (!ctx) return GAT_ERR_INVALID;
    ctx->timeline_on = enable != 0;
    return GAT_OK;
}

int gat_get_timeline(gat_ctx *ctx, uint64_t *out, int cap_ctas)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (!out || !ctx->d_timeline || ctx->timeline_ctas < 1) return fail(ctx, GAT_ERR_INVALID, "no timeline recorded");
    const int n = std::min(cap_ctas, ctx->timeline_ctas);
    GAT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    GAT_CUDA(ctx, cudaMemcpy(out, ctx->d_timeline, static_cast<size_t>(n) * 16 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return n;
}

int gat_beamform(gat_ctx *ctx, int n_ch, int n_taps, int n_ants, const float *d_acc_re, const float *d_acc_im, const float *d_w_re,
                 const float *d_w_im, float *d_y_re, float *d_y_im)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (!d_acc_re || !d_acc_im || !d_w_re || !d_w_im || !d_y_re || !d_y_im) return fail(ctx, GAT_ERR_INVALID, "null pointer argument");
    if (n_ch < 1 || n_taps < 1 || n_taps > GAT_MAX_TAPS || n_ants < 1 || n_ants > kMaxAnts)
        return fail(ctx, GAT_ERR_INVALID, "bad post-correlation shape");
    cudaError_t e = launch_beamform(d_acc_re, d_acc_im, d_w_re, d_w_im, d_y_re, d_y_im, n_ch, n_taps, n_ants, ctx->stream);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "beamform launch");
    ctx->launches += 1;
    return GAT_OK;
}

int gat_eigen_weights(gat_ctx *ctx, int n_ch, int n_taps, int n_ants, const float *d_acc_re, const float *d_acc_im, int tap, float forget,
                      int iters, float *d_cov_re, float *d_cov_im, float *d_w_re, float *d_w_im)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (!d_acc_re || !d_acc_im || !d_cov_re || !d_cov_im || !d_w_re || !d_w_im) return fail(ctx, GAT_ERR_INVALID, "null pointer argument");
    if (n_ch < 1 || n_taps < 1 || n_taps > GAT_MAX_TAPS || n_ants < 1 || n_ants > kMaxAnts || tap < 0 || tap >= n_taps || iters < 1 ||
        iters > 1024 || !(forget >= 0.f) || !(forget <= 1.f))
        return fail(ctx, GAT_ERR_INVALID, "bad eigen-filter arguments");
    cudaError_t e = launch_eigen_weights(d_acc_re, d_acc_im, n_ch, n_taps, n_ants, tap, forget, iters, d_cov_re, d_cov_im, d_w_re, d_w_im,
                                         ctx->stream);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "eigen-filter launch");
    ctx->launches += 1;
    return GAT_OK;
}

// The replica chip-table indices of a correlate call, as the HOT kernel itself computed them: the call is really made
// (DUMP instantiation of correlate_kernel over an all-zero block of n_ants antennas) and the index of every replica entry
// it generated -- first tile from scratch, the following ones by the per-tile NCO advance, through whichever wrap branch
// the launch plan selected -- is read back.  out[l * n_samples + i] = index used for tap l at sample start_sample + i.
int gat_debug_replica_indices(gat_ctx *ctx, const gat_channel *ch, double fs_hz, const int32_t *sample_shifts, int n_taps, int n_ants,
                              int start_sample, int n_samples, unsigned flags, int32_t *out)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (!ch || !out || !sample_shifts || n_samples < 1 || start_sample < 0 || !(fs_hz > 0.0) || n_taps < 1 || n_taps > GAT_MAX_TAPS ||
        n_ants < 1 || n_ants > kMaxAnts || static_cast<int64_t>(start_sample) + n_samples > (1 << 26))
        return fail(ctx, GAT_ERR_INVALID, "bad arguments");
    const int dbg_slot = 65533;
    SignalSlot *s = slot_for(ctx, dbg_slot);
    const int n_total = start_sample + n_samples;
    const int64_t ld = (static_cast<int64_t>(n_total) + 3) & ~3LL;
    if (!s->owned && s->re) {
        rc = free_planes(ctx, *s);
        if (rc) return rc;
    }
    rc = own_slot(ctx, *s, n_total, n_ants, ld);
    if (rc) return rc;
    GAT_CUDA(ctx, cudaMemsetAsync(s->re, 0, 2 * s->cap_floats * sizeof(float), ctx->stream));
    std::vector<float> acc(2 * static_cast<size_t>(n_taps) * n_ants);
    const int32_t slot_id = dbg_slot;
    rc = correlate_impl(ctx, 1, &slot_id, 1, ch, fs_hz, sample_shifts, n_taps, start_sample, n_samples, acc.data(),
                        acc.data() + acc.size() / 2, 0, (flags & GAT_CODE_PHASE_F64) | kFlagDumpReplica | (flags & GAT_DEBUG_STALL_CONSUMERS));
    if (rc) return rc;
    std::vector<int32_t> dump(static_cast<size_t>(ctx->dump_tiles) * ctx->dump_stride);
    GAT_CUDA(ctx, cudaMemcpyAsync(dump.data(), ctx->d_dbg, dump.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    GAT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int l = 0; l < n_taps; ++l) {
        const int koff = sample_shifts[l] - sample_shifts[0];
        for (int i = 0; i < n_samples; ++i) {
            const int rel = start_sample + i - ctx->dump_aligned_start;
            const int t = rel / ctx->dump_tile_len, tt = rel % ctx->dump_tile_len;
            out[static_cast<size_t>(l) * n_samples + i] = dump[static_cast<size_t>(t) * ctx->dump_stride + tt + koff];
        }
    }
    return GAT_OK;
}

int gat_debug_tc_replica_bits(gat_ctx *ctx, int slot, int n_sats, const gat_channel *channels, double fs_hz, const int32_t *sample_shifts,
                              int n_taps, int start_sample, int n_samples, uint8_t *out)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (!channels || !out || !sample_shifts || n_sats < 1 || n_taps < 1 || n_taps > 4 || n_samples < 1 || start_sample < 0)
        return fail(ctx, GAT_ERR_INVALID, "bad arguments");
    const SignalSlot *sl = find_slot(ctx, slot);
    if (!slot_has_signal(sl)) return fail(ctx, GAT_ERR_NO_SIGNAL, "slot has no signal");
    std::vector<float> acc(2 * static_cast<size_t>(n_sats) * n_taps * sl->n_ants);
    const int32_t slot_id = slot;
    rc = correlate_impl(ctx, 1, &slot_id, n_sats, channels, fs_hz, sample_shifts, n_taps, start_sample, n_samples, acc.data(),
                        acc.data() + acc.size() / 2, 0, GAT_TENSOR_TF32 | kFlagDumpReplica);
    if (rc) return rc;
    if (ctx->info.tensor != 1) return fail(ctx, GAT_ERR_UNSUPPORTED, "the call did not run on the tensor-core path");
    const int G = (n_sats + 31) / 32, TJ = ctx->dump_tiles;
    std::vector<uint32_t> dump(static_cast<size_t>(G) * TJ * 32 * 20);
    GAT_CUDA(ctx, cudaMemcpyAsync(dump.data(), ctx->d_dbg, dump.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    GAT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < n_sats; ++k)
        for (int l = 0; l < n_taps; ++l) {
            const int koff = sample_shifts[l] - sample_shifts[0];
            for (int i = 0; i < n_samples; ++i) {
                const int rel = start_sample + i - ctx->dump_aligned_start;
                const int t = rel / kTileCap, e = rel % kTileCap + koff;
                const size_t row = ((static_cast<size_t>(k / 32) * TJ + t) * 32 + k % 32) * 20;
                out[(static_cast<size_t>(k) * n_taps + l) * n_samples + i] = static_cast<uint8_t>((dump[row + (e >> 5)] >> (e & 31)) & 1u);
            }
        }
    return GAT_OK;
}

// single-tap convenience kept from round 1 (then a look-alike kernel, now the hot kernel's own dump)
int gat_debug_chip_indices(gat_ctx *ctx, const gat_channel *ch, double fs_hz, int shift, int n_samples, unsigned flags,
                           int32_t *out)
{
    if (!ctx || !out || n_samples < 1) return GAT_ERR_INVALID;
    const int32_t shifts[3] = {shift - 1, shift, shift + 1};
    std::vector<int32_t> all(3 * static_cast<size_t>(n_samples));
    int rc = gat_debug_replica_indices(ctx, ch, fs_hz, shifts, 3, 1, 0, n_samples, flags, all.data());
    if (rc) return rc;
    std::memcpy(out, all.data() + n_samples, sizeof(int32_t) * n_samples);
    return GAT_OK;
}

}  // extern "C"
