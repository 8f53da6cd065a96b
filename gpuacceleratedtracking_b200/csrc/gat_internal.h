// gat_internal.h -- structures shared by the kernels (gat_correlate.cu) and the C-ABI host
// layer (gat_api.cu).  Not part of the public interface (include/gat.h is).
#pragma once

#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

namespace gat {

constexpr int kTileCap = 256;      // samples per staged smem row (row pitch, floats)
constexpr int kMaxTaps = 11;       // GAT_MAX_TAPS
constexpr int kMaxAnts = 32;       // rows per plane that fit the staging ring
// Two CTA classes: shapes whose accumulators need <= 32 registers run up to 19 consumer warps
// (+ 1 producer = 640 threads, <= 102 registers/thread); the others 11 (+ 1 = 384 threads, 168 registers).
constexpr int kMaxConsumerWarps = 19;
// Two CTA classes: shapes whose accumulators need <= 48 registers run up to 19 consumer warps (+ 1 producer = 640 threads,
// <= 102 registers/thread); the others 11 (+ 1 = 384 threads, 168 registers: 12 warps are 3 per scheduler, and a 13th
// warp would put 4 on one scheduler's 16 K registers = 128 per thread).
// Many taps (>= 7 with 4 antennas per thread) are split over TWO warps ("tap groups", gat_api.cu make_plan): each holds half
// the taps of the same 4 antennas (2 x 4 x 6 = 48 accumulators), repeats the cheap wipe-off, and lands in the 19-warp class.
// Measured before that (64 periods x 16 antennas, one channel, same box): 11 taps 116 us = 0.54 of the HBM roof with 8 warps
// of 168 registers; a 12-warp / 128-register class took 7 taps 105 -> 82 us but spilled at 9 / 11 taps (11 taps: 187 us;
// with the chips loaded in two halves 148 us); a software-pipelined sample loop (next sample's loads between this sample's
// FMAs) was slower as well (128 us at 168 registers): the loop is bound by warps per scheduler, not by a warp's load phase.
#ifndef GAT_WIDE_MAX_TAPS
#define GAT_WIDE_MAX_TAPS 0      // experiments: 4-antenna shapes with 7 .. this many taps run 12 consumer warps + 1 at 128 registers
#endif
#ifndef GAT_HELP11_THREADS
#define GAT_HELP11_THREADS 512   // CTA class of the replica-warp instantiation of the 11-tap shape (512 = register reallocation)
#endif
// Register reallocation (setmaxnreg): a 512-thread CTA launches at 128 registers per thread; the three consumer warpgroups
// (warps 0..11) then grow to kReallocConsumerRegs while the fourth (producer = warp 12, replica warp = warp 13, two idle
// warps) shrinks to kReallocAuxRegs: 384 * 152 + 128 * 56 = 65 536 (160 + 32 makes the producer and the replica warp spill).  The 11-tap shape (88 accumulators) thereby runs
// 3 consumer warps per scheduler without spilling; at a flat 168 registers only 8 warps (2 per scheduler) fit.
#ifndef GAT_REALLOC_MIN_TAPS
#define GAT_REALLOC_MIN_TAPS 7   // 4-antenna shapes with at least this many taps use the reallocation class
#endif
#ifndef GAT_REALLOC_CONSUMER_REGS
#define GAT_REALLOC_CONSUMER_REGS 152
#endif
#ifndef GAT_REALLOC_AUX_REGS
#define GAT_REALLOC_AUX_REGS 56
#endif
constexpr int kReallocConsumerRegs = GAT_REALLOC_CONSUMER_REGS;
constexpr int kReallocAuxRegs = GAT_REALLOC_AUX_REGS;
constexpr int kReallocConsumerWarps = 12;
__host__ __device__ constexpr int block_threads_max(int A, int L)
{
    return 2 * A * L <= 48 ? 640 : ((A == 4 && L >= 7 && L <= GAT_WIDE_MAX_TAPS) ? 416 : 384);
}
// HELP instantiations (replica warp, see correlate_kernel): the consumer warps carry no code-NCO state and generate nothing,
// which lets the many-tap shapes (4 antennas per thread, >= 7 taps) fit 128 registers: 12 consumer warps + producer +
// replica warp = 448 threads, 3 consumer warps per scheduler.
__host__ __device__ constexpr int block_threads_help(int A, int L)
{
    return (A == 4 && L >= 7) ? (L >= GAT_REALLOC_MIN_TAPS ? GAT_HELP11_THREADS : 448) : block_threads_max(A, L);
}
__host__ __device__ constexpr bool help_realloc(int A, int L) { return block_threads_help(A, L) == 512; }
constexpr int kHelperMaxSats = 4;     // satellites per CTA the one replica warp keeps up with (1 in the reallocation class)
constexpr int kRepBarOff = 512;       // replica ring barriers in the shared-memory header: full[i] at 512 + 8 i, empty[i] at 1024 + 8 i
__host__ __device__ constexpr int max_consumer_warps(int A, int L) { return block_threads_max(A, L) / 32 - 1; }
constexpr int kMaxStages = 16;
constexpr int kMaxPeers = 8;
constexpr int kCodeColAlign = 16;  // device chip-table columns are padded to this many bytes
constexpr int kSmemHeaderBytes = 2048;  // full/empty barriers (0..255), chip-table barrier (256) + its back-pressure barrier (264)
constexpr uint32_t kFlagStallConsumers = 0x100u;   // == GAT_DEBUG_STALL_CONSUMERS (include/gat.h)
constexpr uint32_t kFlagDumpReplica = 0x200u;      // internal: run the DUMP instantiation (gat_debug_replica_indices)
constexpr uint32_t kFlagResidentPlan = 0x400u;     // internal: plan and marshal only -- gat_resident_begin keeps (plan, args)

// One satellite channel of one period, pre-digested on the host (gat_api.cu: fill_sat).
struct SatDev {
    const int8_t *code;   // +-1 chips of (system, prn), length code_len
    int32_t code_len;     // modulo length Lc
    int32_t nco_fp;       // Q-format fractional bits: 63 - ceil(log2(Lc))   [Tracking.jl gen_code_replica!]
    int64_t nco_delta;    // floor(fc * 2^fp / fs)
    int64_t nco_start;    // floor(mod(phase, Lc) * 2^fp)
    uint64_t car_phase;   // carrier phase at relative sample 0, Q0.64 cycles
    uint64_t car_delta;   // carrier phase step per sample, Q0.64 cycles
    double code_ratio;    // fc / fs                (GAT_CODE_PHASE_F64)
    double code_phase;    // start code phase, chips (GAT_CODE_PHASE_F64)
};

// One signal block: 2-D TMA descriptors of the re / im planes, dims {n_samples, n_ants},
// row stride ld*4 bytes, box {kTileCap, n_ants}.  Out-of-range samples are zero-filled by the
// TMA unit, so tiles may start at any sample and run past the end of the block.
struct alignas(64) PeriodDev {
    CUtensorMap re;
    CUtensorMap im;
};

constexpr int kInlineBytes = 3328;   // small parameter blocks ride in the kernel arguments (no H2D copy, no events)

struct alignas(64) CorrArgs {
    // [PeriodDev x P][pad to 64][SatDev x P*K] when the whole block fits; else `periods` / `sats` point to it in global memory
    unsigned char inline_blk[kInlineBytes];
    int32_t use_inline, inline_sat_off;
    const PeriodDev *periods;   // [n_periods]
    const SatDev *sats;         // [n_periods * n_sats]
    float *out_re, *out_im;     // [n_periods][n_sats][n_taps][n_ants]
    float *partials;            // [(jobs + grid)][roles * RP]
    unsigned int *grid_barrier; // monotonically increasing arrival counter (never reset)
    unsigned int barrier_target;// value the counter reaches when every CTA of THIS launch arrived
    int32_t shifts[kMaxTaps];
    int32_t koff4[kMaxTaps + 1];       // 4 * (shifts[l] - shifts[0]): byte offset of tap l in the code replica (+ 1: tap-group padding)
    int32_t span;                      // shifts[last] - shifts[0]
    int32_t TG;                        // tap groups: warps splitting the taps of one (satellite, antenna group); template L = taps per warp
    int32_t tt_stride;                 // samples between two iterations of a lane: 32, or 32 * SL when tiles are split
    int32_t n_periods, n_sats, n_ants, n_taps;
    int32_t start_sample, n_samples;   // integrated range [start, start + n)
    int32_t aligned_start;             // first staged sample (== start_sample with tensor-map TMA)
    int32_t aligned_len;               // staged length (== n_samples)
    int32_t tile_len;                  // <= kTileCap, multiple of 32
    int32_t tiles_per_job;
    int32_t S, AG, SL, W, G;           // sats/CTA, antenna groups, sample slices, consumer warps, sat groups
    int32_t stages;
    int32_t n_parts;                   // sharded slots: `periods` holds n_parts descriptors pairs per period, part j covers tiles
    int32_t part_tiles;                //   [j * part_tiles, (j + 1) * part_tiles) of the block (absolute, kTileCap samples each)
    int32_t rep_stride;                // floats per consumer warp for its code replica (>= tile_len + span)
    int32_t cache_stride;              // bytes per satellite in the smem chip-table cache (multiple of 16)
    int32_t total_tiles;               // jobs * tiles_per_job
    float out_scale;
    int32_t rep_helper;                // 1: warp W + 1 generates every tile's replicas (HELP instantiation), consumers only read them
    int32_t rep_single_wrap;   // a tile (+ tap span) advances every code by less than one period: branch-free index wrap                   // multiplies every accumulator at emit (1 unless raw integer tiles carry a scale)
    int32_t fin_group;                 // lanes cooperating on one output element in the finalize (pow2 <= 32)
    int32_t split_tiles;               // 1: every slice works on every tile (small problems); 0: whole tiles round-robin
    uint32_t flags;
    // fused multi-GPU gather (GAT_GATHER): peer mappings of every rank's buffer, this rank's slice offset
    float *peer_re[kMaxPeers];
    float *peer_im[kMaxPeers];
    unsigned int *peer_flag[kMaxPeers];
    int32_t n_peers, my_rank;
    uint32_t gather_seq;               // value released into flags[my_rank] when this launch is complete
    unsigned long long gather_elems;   // elements per rank slice
    unsigned long long gather_off;     // this call's first element inside the slice (gat_gather_set_offset)
    unsigned int *done_counter;        // CTAs that finished their stores (self-cleaning)
    unsigned long long *timeline;      // debug: [grid][16] globaltimer stamps, or nullptr
    uint32_t *dump;                    // debug (DUMP instantiations): [tiles_per_job][dump_stride] chip-table index of every replica entry
    int32_t dump_stride;               // entries per tile in `dump` (the one-tile replica stride)
    int32_t rep_bufs;                  // code-replica buffers of rep_stride floats in shared memory (W, or 2 per group with replica warps)
    int32_t visit_tiles;               // reallocation class: consecutive tiles a consumer warp works through per visit (1 or 2)
    int32_t phase_off;                 // sample index, in the channel phases' frame, of start_sample (0: the phases refer to start_sample;
                                       // gat_set_sample_origin: the phases refer to sample 0 of the integration period this slot is a range of)
};

static_assert(sizeof(CorrArgs) <= 4096, "kernel parameter space");

struct LaunchPlan {
    int A;            // antennas per thread (template)
    int L;            // taps (template)
    bool f64;
    bool sc16;        // raw int16 I/Q tiles
    bool dump;        // replica-index dump instantiation (debug)
    bool help;        // replica-warp instantiation
    int grid, block;
    size_t smem;
    int RP;           // padded accumulators per role
    int jobs;
};

// smem carve-up, shared by host sizing and device addressing
__host__ __device__ inline size_t smem_tile_floats(int AG, int A, bool sc16 = false) { return (size_t)(sc16 ? 1 : 2) * AG * A * kTileCap; }
__host__ __device__ inline int padded_acc(int A, int L) { return ((2 * A * L) + 31) / 32 * 32; }

// ---- tensor-core path (gat_correlate_tc.cu) ----
struct alignas(64) TcPeriod {
    CUtensorMap map;   // 4-D view of the two planes: {4 samples, antennas, planes, sample groups}, box {4, 16, 2, 64}
    int32_t swapped;   // the im plane lies BELOW the re plane in memory: plane 0 of the view is im
};
struct alignas(64) TcArgs {
    const TcPeriod *periods;
    const SatDev *sats;          // [n_periods][n_sats]
    float *partials;             // [(jobs + grid)][re, im][128 rows][16 antennas]
    float *out_re, *out_im;      // [n_ants x n_taps x n_sats x n_periods]
    int32_t n_periods, n_sats, n_ants, n_taps;
    int32_t shift0, span;        // first tap's sample shift, last - first
    int32_t koff[4];             // tap offsets relative to the first tap
    int32_t start_sample, n_samples, aligned_start;
    int32_t tiles_per_job, G;
    int64_t total_units;
    int32_t win_ok;              // a tile's replica (tile + tap span + row padding) advances < 32 chips on every channel
    uint32_t *dump;              // debug: [total_units][32 channels][kTcRepWords] replica sign-bit words of every tile, or nullptr
    int32_t debug;               // GAT_TC_DEBUG bit mask (experiments; results are wrong with 2..128 set): 2 skip MMAs, 16 skip TMA,
                                 // 64 skip replica rows, 128 hand-over skeleton only, 4096 record the hand-over timeline of CTA 0
};
cudaError_t configure_tc_kernel();
cudaError_t launch_correlate_tc(const TcArgs &args, int grid, int jobs, cudaStream_t stream);

cudaError_t launch_correlate(const LaunchPlan &plan, const CorrArgs &args, cudaStream_t stream);
cudaError_t configure_kernels();   // opt-in to > 48 KB dynamic smem for every instantiation
bool kernel_available(int A, int L);
bool dump_kernel_available(int A, int L);
bool help_kernel_available(int A, int L, bool f64, bool dump);

// ---- resident kernel (gat_resident.cu, gat_resident_* in include/gat.h) ----
constexpr int kResChunks = 6;                 // a command is up to kResChunks warp-wide loads of 32 cells of 16 bytes {d0, d1, d2, seq}
constexpr int kResMaxCells = 32 * kResChunks;
constexpr int kResMaxSats = 32;               // data words: [op, slot index, n_sats, 0][SatDev x n_sats] <= 3 * kResMaxCells
constexpr uint32_t kResOpCorrelate = 1u, kResOpExit = 2u;
constexpr int kResCmdSmemBytes = 2304;        // the command's data words in shared memory, behind the plan's carve-up
struct ResCtl {
    const uint4 *cmd_host;        // the host's command cells (pinned, mapped: device pointer)
    uint4 *relay;                 // device memory: the cells as CTA 0 received them (the other CTAs poll these)
    const PeriodDev *slot_maps;   // device memory: descriptors of the session's slots
    int32_t n_cells;
    int32_t cmd_off;              // byte offset of the command area in dynamic shared memory
    uint32_t idle_limit_ms;       // CTA 0 ends the kernel after this long without a command
    uint32_t first_seq;           // sequence number of the first command this launch serves
    unsigned long long *stamps;   // debug (GAT_RESIDENT_DEBUG): host-mapped [4] globaltimer stamps of the last command, or nullptr
};
bool resident_kernel_available(int A, int L, bool help);
cudaError_t launch_resident(const LaunchPlan &plan, const CorrArgs &args, const ResCtl &ctl, size_t smem_bytes, cudaStream_t stream);

cudaError_t launch_gather_wait(unsigned int *const *flags_unused, unsigned int *local_flags, int world, unsigned int seq,
                               cudaStream_t stream);
// cross-rank flags of the signal ring: lane d release-stores `seq` into dst[d][my_rank]; the wait spins on local flags
struct FlagPtrs { unsigned int *p[kMaxPeers]; };
cudaError_t launch_flag_signal(const FlagPtrs &dst, int world, int my_rank, unsigned int seq, cudaStream_t stream);
cudaError_t launch_flag_wait(unsigned int *local_flags, int world, unsigned int seq, cudaStream_t stream);


// expand interleaved complex integer samples [n_ants][ld_in][2] into FP32 planes [n_ants][ld_out]
cudaError_t launch_sum_slices(const float *in_re, const float *in_im, size_t slice_stride, int n_slices, size_t n, float *out_re,
                              float *out_im, cudaStream_t stream);
cudaError_t launch_beamform(const float *acc_re, const float *acc_im, const float *w_re, const float *w_im, float *y_re, float *y_im,
                            int n_ch, int n_taps, int n_ants, cudaStream_t stream);
cudaError_t launch_eigen_weights(const float *acc_re, const float *acc_im, int n_ch, int n_taps, int n_ants, int tap, float forget, int iters,
                                 float *cov_re, float *cov_im, float *w_re, float *w_im, cudaStream_t stream);
cudaError_t launch_expand_sc(const void *iq, int bytes_per_component, int64_t ld_in, float *re, float *im, int64_t ld_out,
                             int n_samples, int n_ants, float scale, cudaStream_t stream);

cudaError_t launch_gen_signal(float *re, float *im, int64_t ld, const int8_t *code, int code_len,
                              double code_ratio, double carrier_freq_hz, double fs_hz, double code_phase,
                              double carrier_phase_rad, int n_samples, int n_ants,
                              double ant_phase_step_rad, double noise_sigma, uint64_t seed,
                              int superpose, cudaStream_t stream);

}  // namespace gat
