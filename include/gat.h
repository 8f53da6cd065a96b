/*
 * gat.h -- C ABI of libgat, the B200-native GNSS correlator engine.
 *
 * This is the drop-in boundary for the reference's correlate hot path
 * (coezmaden/GPUAcceleratedTracking; citations relative to /root/reference):
 *
 *   GPU-style entry   kernel_algorithm(threads, blocks, shmem, code_replica, codes, code_frequency,
 *                     sampling_frequency, start_code_phase, prn, num_samples, num_of_shifts,
 *                     code_length, accum_re, accum_im, ..., signal_re, signal_im,
 *                     correlator_sample_shifts, carrier_frequency, carrier_phase, num_ants,
 *                     num_corrs, algorithm)                       src/algorithms.jl:1485-1545
 *   CPU-style entry   Tracking.downconvert_and_correlate!(system, signal, correlator, code_replica,
 *                     code_phase, carrier_replica, carrier_phase, downconverted_signal,
 *                     code_frequency, correlator_sample_shifts, carrier_frequency,
 *                     sampling_frequency, signal_start_sample, num_samples_left, prn)
 *                                                                 src/benchmarks.jl:63-79
 *
 * A Julia method of either function forwards to gat_correlate() through `ccall`
 * (julia/GATB200.jl).  Plain pointers and sizes only; no CUDA, torch or
 * C++ types cross this boundary.  Every function returns GAT_OK (0) or a negative
 * gat_status and never throws or aborts; gat_last_error() gives the text.
 *
 * Conventions (identical to the reference's):
 *   - signal is SoA complex Float32: separate `re` and `im` planes, each column-major
 *     [n_samples x n_ants] with leading dimension `ld` (sample index fastest)
 *     (src/gen_signal.jl:181-184).
 *   - code phase in chips, code/carrier/sampling frequency in Hz, carrier phase in CYCLES
 *     (src/algorithms.jl:172), prn 1-based, sample shifts in samples, ascending
 *     (late -> early), start_sample 0-BASED here (the Julia wrapper subtracts 1).
 *   - outputs are [n_ants x n_taps x n_sats (x n_periods)] column-major (antenna fastest):
 *     the accum[M, L, K] layout of src/algorithms.jl:712.
 *   - chips are +-1 int8, table column-major [code_len x n_prn]  (GNSSSignals `system.codes`).
 */
#ifndef GAT_H
#define GAT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GAT_VERSION 100 /* 0.1.0 */

typedef enum gat_status {
    GAT_OK = 0,
    GAT_ERR_INVALID = -1,     /* bad argument (null pointer, size out of range, unsorted shifts ...) */
    GAT_ERR_CUDA = -2,        /* CUDA runtime error; text in gat_last_error */
    GAT_ERR_UNSUPPORTED = -3, /* shape outside the compiled kernel family (n_taps > 11, window too large) */
    GAT_ERR_ALIGNMENT = -4,   /* zero-copy device signal not 16-byte aligned / ld % 4 != 0 */
    GAT_ERR_NO_CODES = -5,    /* channel names a (system_id, prn) with no chip table set */
    GAT_ERR_NO_SIGNAL = -6,   /* slot has no signal bound */
    GAT_ERR_NO_DEVICE = -7    /* no CUDA device / wrong architecture (sm_100 required) */
} gat_status;

/* flags for gat_correlate* */
#define GAT_ACCUMULATE 1u     /* out += result (device outputs only): the `+=` of src/algorithms.jl:207
                                 and the atomic accumulate of :625-632.  Default overwrites.          */
#define GAT_GATHER 4u         /* multi-GPU: the kernel epilogue stores this rank's accumulators into the gather
                                 buffer of EVERY rank (peer memory over NVLink) instead of out_re/out_im, which
                                 are ignored.  Needs gat_gather_create/connect.                             */
#define GAT_TENSOR_TF32 8u    /* allow the tensor-core path (tcgen05.mma kind::tf32) for blocks shared by many channels:
                                 <= 4 taps, <= 16 antennas, NCO convention, chip tables <= 10240 chips (GPS L1 C/A, L5).  W and the
                                 samples are rounded to TF32 (10-bit mantissa), sums are FP32: <= 12-bit integer
                                 samples stay exact, the accumulators carry ~3e-4 * sqrt(N) * rms(sample) of rounding
                                 noise instead of the FP32 kernel's ~1e-7 relative.  Shapes outside the envelope run
                                 on the FP32 kernel as usual (gat_launch_info.tensor tells which one ran).  Measured
                                 on B200 (16 antennas, 3 taps, 50000 samples): 2.0-2.2x the FP32 kernel from 64
                                 channel-blocks per call (264 channels: 126 -> 66 us, 1024 channels: 446 -> 202 us),
                                 on par at 32; the caller decides.                                                  */
#define GAT_CODE_PHASE_F64 2u /* chip index = mod(floor(fc/fs*(n+shift)+phase), Lc) in IEEE double,
                                 bit-exact with the reference's GPU kernels (src/algorithms.jl:179-182).
                                 Default is the Int64 Q-format NCO of Tracking.jl's CPU path, bit-exact
                                 with gen_code_replica! [upstream].                                   */

#define GAT_DEBUG_STALL_CONSUMERS 0x100u /* test hook: the consumer warps of the FP32 kernel sleep ~20 us at every
                                 segment start, so the producer warp runs as far ahead as the ring lets it
                                 (regression test of the chip-table hand-shake; results are unchanged).   */

/* built-in GNSS ids for gat_gen_code / the `system_id` of a channel.  Any other non-negative
 * id may be used for caller-supplied tables (gat_set_codes). */
#define GAT_GPSL1 0
#define GAT_GPSL5 1
#define GAT_MAX_SYSTEMS 8
#define GAT_MAX_TAPS 11
#define GAT_MAX_ANTS 32

typedef struct gat_ctx gat_ctx;

/* One satellite channel for one integration period.  Mirrors the per-call scalars of
 * downconvert_and_correlate! (code_phase, code_frequency, carrier_phase, carrier_frequency, prn). */
typedef struct gat_channel {
    int32_t system_id;           /* which chip table (GAT_GPSL1, GAT_GPSL5, ...) */
    int32_t prn;                 /* 1-based column of that table */
    double code_phase_chips;     /* at the first integrated sample */
    double code_freq_hz;         /* code frequency incl. code Doppler */
    double carrier_phase_cycles; /* at the first integrated sample, cycles */
    double carrier_freq_hz;      /* IF + carrier Doppler */
} gat_channel;

/* ---- life cycle ------------------------------------------------------------------------ */
int gat_version(void);
const char *gat_status_string(int status);
int gat_device_count(void);                       /* visible CUDA devices, <0 on error */
int gat_create(gat_ctx **out, int device_id);     /* one ctx = one device + one non-blocking stream */
int gat_destroy(gat_ctx *ctx);
const char *gat_last_error(gat_ctx *ctx);         /* valid until the next call on ctx */
int gat_sync(gat_ctx *ctx);                       /* cudaStreamSynchronize; CUDA.@sync of src/benchmarks.jl:872 */
void *gat_stream(gat_ctx *ctx);                   /* the ctx's cudaStream_t, for interop */
/* Run on a caller-owned cudaStream_t (e.g. CUDA.jl's task stream, torch's current stream) so
 * work is ordered with the caller's own kernels.  The handle is used as given: NULL means the
 * legacy default stream (stream 0).  gat_use_own_stream goes back to the ctx's private stream. */
int gat_set_stream(gat_ctx *ctx, void *cuda_stream);
int gat_use_own_stream(gat_ctx *ctx);

/* ---- chip tables (replaces the CuTexture / gmem `codes` argument, src/benchmarks.jl:829-837) */
/* Host helper: generate the +-1 chips of one PRN of a built-in system.  Returns code length
 * (1023 / 10230) or <0.  (GNSSSignals.jl GPSL1()/GPSL5() `codes[:, prn]`.) */
int gat_gen_code(int system_id, int prn, int8_t *out, int cap);
/* Upload a table; chips is HOST memory, column-major [code_len x n_prn]. */
int gat_set_codes(gat_ctx *ctx, int system_id, const int8_t *chips, int code_len, int n_prn);
/* Nominal chip rate of a table, used ONLY by gat_gen_signal (a channel carries its own code_freq_hz).  The built-in
 * ids get 1.023 MHz / 10.23 MHz when gat_set_codes installs a table of the ICD length; every other table (other ids,
 * BOC / tiered codes, a custom table under id 0 or 1) has no rate until this is called -- gat_gen_signal then
 * returns GAT_ERR_INVALID instead of silently assuming 1.023 MHz. */
int gat_set_code_frequency(gat_ctx *ctx, int system_id, double code_freq_hz);

/* ---- signal blocks --------------------------------------------------------------------- */
/* A ctx holds a ring of signal slots (one 1 ms block each).  Upload copies (H2D or D2D,
 * asynchronously on the ctx stream, src may be pageable or pinned) into ctx-owned padded
 * storage; bind registers caller-owned DEVICE planes zero-copy (CuArray / torch data_ptr). */
int gat_upload_signal(gat_ctx *ctx, int slot, const float *re, const float *im,
                      int n_samples, int n_ants, int ld, int src_is_device);
int gat_bind_signal(gat_ctx *ctx, int slot, const float *d_re, const float *d_im,
                    int n_samples, int n_ants, int ld);
/* Integer front-end samples: interleaved complex (I, Q) int16 / int8 per antenna, antenna-major
 * [n_ants][ld][2] with `ld` in complex samples -- the usual SDR wire format (2-4x fewer bytes than FP32).
 * Results are those of FP32 planes holding (float)x * scale; scale = 1 keeps the integers exact.
 *  - int16: the slot keeps the raw words in HBM.  gat_correlate* reads them directly (conversion in
 *    registers, `scale` applied to the accumulators) when scale is a power of two, the NCO code-phase
 *    convention is used and at most 16 channels share the block; otherwise, and for gat_download_signal /
 *    gat_gen_signal(superpose), the FP32 planes are expanded once on first need.
 *  - int8: expanded to FP32 planes by a streaming kernel at upload. */
int gat_upload_signal_sc16(gat_ctx *ctx, int slot, const int16_t *iq, int n_samples, int n_ants, int ld,
                           float scale, int src_is_device);
int gat_upload_signal_sc8(gat_ctx *ctx, int slot, const int8_t *iq, int n_samples, int n_ants, int ld,
                          float scale, int src_is_device);
/* Device-side synthetic generator with gen_signal semantics (src/gen_signal.jl:135-152):
 * code phase Float64 -> floor/mod, carrier phase Float64 -> Float32 -> cos/sin, every antenna
 * identical.  Extensions (off when zero): per-antenna phase step [rad], AWGN sigma (seeded),
 * additive superposition onto the existing slot contents.  This call WRITES the slot: a ctx-owned slot, or caller
 * planes registered with gat_bind_signal (then the caller's memory is generated into in place -- that is the point of
 * binding writable planes); imported (gat_slot_import) and ring slots are refused. */
int gat_gen_signal(gat_ctx *ctx, int slot, int system_id, int prn, double carrier_freq_hz,
                   double fs_hz, double start_code_phase, double start_carrier_phase_rad,
                   int n_samples, int n_ants, double ant_phase_step_rad, double noise_sigma,
                   uint64_t seed, int superpose);
/* Shape of the block a slot holds (what the output size of a correlate call over it depends on). */
int gat_slot_shape(gat_ctx *ctx, int slot, int *n_samples_out, int *n_ants_out);
/* Copy a slot's planes back (tests): host column-major [n_samples x n_ants], ld = n_samples. */
int gat_download_signal(gat_ctx *ctx, int slot, float *re, float *im);

/* Multi-GPU ingest without a broadcast (one process per GPU, same node): the ingest process exports a
 * ctx-owned slot as an opaque descriptor (CUDA IPC handle + layout) that the host passes to the other
 * processes by any means; gat_slot_import maps it and binds it like gat_bind_signal, so a correlate call
 * on that rank TMA-loads its signal tiles straight out of the ingest GPU's HBM over NVLink -- the transfer
 * is fused into the kernel's own tile pipeline, no receive buffer, no extra HBM write + read.  The exporter
 * keeps ownership and must not rewrite the slot while importers still read it (host-level ordering, e.g. a
 * barrier or an event); the importer releases the mapping by re-using or destroying the slot. */
#define GAT_SLOT_DESC_BYTES 96
int gat_slot_export(gat_ctx *ctx, int slot, unsigned char *desc_out /* [GAT_SLOT_DESC_BYTES] */);
int gat_slot_import(gat_ctx *ctx, int slot, const unsigned char *desc /* [GAT_SLOT_DESC_BYTES] */);

/* ---- the hot path ---------------------------------------------------------------------- */
/* One integration period: n_sats channels over the same signal block.
 *   out_re/out_im: [n_ants x n_taps x n_sats].  out_is_device = 0: host pointers, the call
 *   returns after the D2H copy (synchronous, like CUDA.@sync + Array()).  out_is_device = 1:
 *   device pointers, the call is asynchronous on the ctx stream (call gat_sync). */
int gat_correlate(gat_ctx *ctx, int slot, int n_sats, const gat_channel *channels, double fs_hz,
                  const int32_t *sample_shifts, int n_taps, int start_sample, int n_samples,
                  float *out_re, float *out_im, int out_is_device, unsigned flags);

/* A batch of integration periods in ONE launch: period p reads slots[p] with channels
 * channels[p * n_sats + k].  All periods share fs, shifts, start_sample, n_samples and the
 * antenna count.  out: [n_ants x n_taps x n_sats x n_periods]. */
int gat_correlate_batch(gat_ctx *ctx, int n_periods, const int32_t *slots, int n_sats,
                        const gat_channel *channels, double fs_hz, const int32_t *sample_shifts,
                        int n_taps, int start_sample, int n_samples, float *out_re, float *out_im,
                        int out_is_device, unsigned flags);

/* Host-buffer convenience = the CPU-style signature: upload + correlate + download. */
int gat_downconvert_and_correlate(gat_ctx *ctx, const float *h_re, const float *h_im, int ld,
                                  int n_ants, int n_sats, const gat_channel *channels, double fs_hz,
                                  const int32_t *sample_shifts, int n_taps, int start_sample,
                                  int n_samples, float *h_out_re, float *h_out_im, unsigned flags);

/* The same over a BATCH of periods with the host<->device traffic pipelined inside the library (a host loop of
 * downconvert_and_correlate! calls over consecutive 1 ms blocks, src/benchmarks.jl:63-79): h_re[p] / h_im[p] are the
 * host planes of period p ([n_ants x ld]); chunks of 16 periods are copied on an internal ingest stream into a ring of
 * staging buffers while the kernel of the previous chunk runs; channels[p * n_sats + k]; results
 * [n_ants x n_taps x n_sats x n_periods] in host memory, one synchronisation at the end.  Pinned host memory
 * (cudaMallocHost, or gat_host_register on caller memory) makes the copies asynchronous at full PCIe rate. */
int gat_ingest_correlate(gat_ctx *ctx, int n_periods, const float *const *h_re, const float *const *h_im, int ld,
                         int n_ants, int n_sats, const gat_channel *channels, double fs_hz,
                         const int32_t *sample_shifts, int n_taps, int start_sample, int n_samples,
                         float *h_out_re, float *h_out_im, unsigned flags);
int gat_host_register(void *ptr, uint64_t bytes);   /* cudaHostRegister (portable); already registered = GAT_OK */
int gat_host_unregister(void *ptr);

/* ---- post-correlation array processing (Tracking.jl `track(...; post_corr_filter)` [upstream]) ----
 * For accumulators that stay on the device: all pointers are DEVICE pointers, both calls are asynchronous on the
 * ctx stream (ordered behind the correlate call that produced `acc`).  n_ch = n_sats x n_periods of that call.
 *   gat_beamform:      y[l, k] = sum_m conj(w[m, k]) * acc[m, l, k];  acc [n_ants x n_taps x n_ch],
 *                      w [n_ants x n_ch], y [n_taps x n_ch].
 *   gat_eigen_weights: eigen-beamformer.  R_k <- forget * R_k + p p^H with p = acc[:, tap, k] (cov: caller-owned
 *                      state [n_ants x n_ants x n_ch], row-major R[i][j] per channel, zero it to start), then
 *                      `iters` power iterations from the previous w (all-zero w = cold start from p); w comes
 *                      back with unit norm and antenna 0 real, non-negative. */
int gat_beamform(gat_ctx *ctx, int n_ch, int n_taps, int n_ants, const float *d_acc_re, const float *d_acc_im,
                 const float *d_w_re, const float *d_w_im, float *d_y_re, float *d_y_im);
int gat_eigen_weights(gat_ctx *ctx, int n_ch, int n_taps, int n_ants, const float *d_acc_re, const float *d_acc_im,
                      int tap, float forget, int iters, float *d_cov_re, float *d_cov_im, float *d_w_re, float *d_w_im);

/* ---- multi-GPU gather fused into the kernel epilogue (one process per GPU, same node) ------ */
/* Every rank allocates [world x elems_per_rank] FP32 re + im planes plus one arrival flag per rank and
 * exports a CUDA IPC handle (64 bytes) that the host exchanges by any means (torch.distributed,
 * MPI, a file).  After gat_gather_connect a correlate call with GAT_GATHER writes its
 * [n_ants x n_taps x n_sats x n_periods] block into slice `rank` of all `world` buffers with plain
 * stores to the peer mappings, then its last CTA release-stores the call's sequence number into
 * flags[rank] of every buffer.  gat_gather_wait queues (on the ctx stream) a wait until every rank's
 * flag reached this rank's own sequence number; all ranks must issue the same sequence of calls. */
#define GAT_IPC_HANDLE_BYTES 64
int gat_gather_create(gat_ctx *ctx, int world, int rank, uint64_t elems_per_rank, unsigned char *handle_out);
int gat_gather_connect(gat_ctx *ctx, const unsigned char *handles /* [world][GAT_IPC_HANDLE_BYTES] */);
/* Where inside its slice the NEXT GAT_GATHER calls put their first element (default 0; sticky): lets a step made of
 * several launches (chunks of periods) fill one gather buffer.  gat_gather_wait after the last launch covers them all. */
int gat_gather_set_offset(gat_ctx *ctx, uint64_t elem_offset);
int gat_gather_wait(gat_ctx *ctx);
int gat_gather_read(gat_ctx *ctx, float *h_re, float *h_im); /* sync + D2H of [world x elems_per_rank] */
/* ---- sample sharding: the decomposition for FEW channels per GPU (SURVEY 8e "alternative worth measuring") ----
 * Sharding the satellites means every GPU needs every block (the signal exchange above); with one or two channels per GPU
 * that exchange (NVLink, <= 0.9 TB/s) is all the step does.  The accumulators are SUMS over samples, so the samples can be
 * sharded instead: every GPU correlates ALL channels over the sample range its own PCIe link delivered -- no signal crosses
 * NVLink -- and the partial sums are added across the GPUs:
 *   gat_set_sample_origin(ctx, o)  (sticky; -1 = off)  the ctx's slots hold samples [o, o + n) of the integration period the
 *       channel phases refer to: the code / carrier phase of slot sample s is taken at period sample o + s, with the same
 *       integer NCO and Q0.64 carrier arithmetic as a whole-block call (bit-exact chip indices), so the partial sums of the
 *       ranges add up to the whole block's accumulators up to FP32 summation order;
 *   gat_correlate*(... GAT_GATHER) puts this rank's partial sums into slice `rank` of every rank's gather buffer;
 *   gat_gather_wait; gat_gather_sum(n, out_re, out_im) adds the `world` slices in rank order into device outputs
 *       (stream-ordered; identical on every rank). */
int gat_set_sample_origin(gat_ctx *ctx, int origin);
int gat_gather_sum(gat_ctx *ctx, uint64_t n_elems, float *d_out_re, float *d_out_im);
int gat_gather_destroy(gat_ctx *ctx);

/* ---- signal ring: the all-gather of the signal blocks fused into the correlate kernel ----------
 * Satellite channels shard across the GPUs of a box, so every GPU needs every signal block (north_star: "each
 * integration period's signal block is NCCL-broadcast over NVLink").  The ring replaces the broadcast + receive
 * buffer by peer reads inside the kernel: a block is cut into `world` contiguous sample ranges of whole 256-sample
 * tiles, rank r keeps range r of each of the ring's `n_slots` blocks in its own HBM (fed through its own PCIe link),
 * every rank maps all owners' memory once, and the correlate kernel's producer warp TMA-loads each tile from the
 * owner it lives on -- the exchange overlaps the math tile by tile, each byte crosses NVLink once per reader and
 * never lands in the reader's HBM.  After gat_ring_connect* the ctx's slots 0 .. n_slots-1 ARE the ring's blocks
 * (use them in gat_correlate* like any slot; start_sample must be a multiple of 256; FP32 kernel only).
 *   one process per GPU : gat_ring_create on every rank, exchange the 64-byte handles, gat_ring_connect.
 *   one process, n GPUs : gat_ring_create on every ctx, gat_ring_connect_local with the array of all ctxs
 *                         (several ctxs may share one device: "logical ranks", used by the single-GPU tests).
 * Cross-rank ordering is by sequence flags in the ring memory, all stream-ordered (no host blocking):
 *   ingest stream (internal) : gat_ring_acquire(releases) -> gat_ring_upload*(slot ...) -> g = gat_ring_publish()
 *   ctx stream               : gat_ring_wait(g) -> gat_correlate*(...) -> r = gat_ring_release()
 * gat_ring_wait(g) holds the ctx stream until EVERY rank has published its g-th generation; gat_ring_acquire(r)
 * holds the ingest stream until every rank has released r generations (pass the count that frees the slots about
 * to be overwritten; <= 0 waits for nothing).  All ranks must issue the same sequence.  A peer that never
 * arrives trips a 20 s watchdog (CUDA error on the waiting rank) instead of hanging the device. */
int gat_ring_create(gat_ctx *ctx, int world, int rank, int n_slots, int n_samples, int n_ants,
                    unsigned char *handle_out /* [GAT_IPC_HANDLE_BYTES] */);
int gat_ring_connect(gat_ctx *ctx, const unsigned char *handles /* [world][GAT_IPC_HANDLE_BYTES] */);
int gat_ring_connect_local(gat_ctx *ctx, gat_ctx *const *peers /* [world], peers[rank] may be NULL */);
int gat_ring_part(gat_ctx *ctx, int rank, int *start_out, int *len_out);   /* sample range owned by `rank` */
/* Copy THIS rank's sample range of a block into ring slot `slot` (asynchronous, ingest stream).  gat_ring_upload takes
 * the planes of the whole block ([n_ants x ld], like gat_upload_signal) and reads only its own range of each row;
 * gat_ring_upload_part takes planes that START at the rank's first sample (a host that holds only its share). */
int gat_ring_upload(gat_ctx *ctx, int slot, const float *re, const float *im, int ld, int src_is_device);
int gat_ring_upload_part(gat_ctx *ctx, int slot, const float *re_part, const float *im_part, int ld, int src_is_device);
int gat_ring_publish(gat_ctx *ctx);                 /* returns this rank's generation count (>= 1), or < 0 */
int gat_ring_wait(gat_ctx *ctx, int generation);
int gat_ring_release(gat_ctx *ctx);                 /* returns this rank's release count (>= 1), or < 0 */
int gat_ring_acquire(gat_ctx *ctx, int releases);
/* Mirror view, for kernels that are compute-bound (many satellites per block and GPU): the copy engines bring the peers'
 * shares into local HBM one generation ahead, under the previous generation's kernel and without using an SM.  After
 * gat_ring_enable_mirror the ctx's slots n_slots .. 2 n_slots-1 are the SAME blocks read from those local copies.
 *   ingest stream : t = gat_ring_prefetch(first_slot, n, generation, releases)  -- waits until every rank published
 *                   `generation` and until this rank's own release number `releases` has run (<= 0: nothing), then copies
 *                   the peers' shares of ring slots first_slot .. first_slot + n-1 (one contiguous copy per peer)
 *   ctx stream    : gat_ring_mirror_wait(t) -> gat_correlate*(slots n_slots + ...) -> gat_ring_release() */
int gat_ring_enable_mirror(gat_ctx *ctx);
int gat_ring_prefetch(gat_ctx *ctx, int first_slot, int n_slots, int generation, int releases);   /* returns a ticket >= 1 */
int gat_ring_mirror_wait(gat_ctx *ctx, int ticket);
int gat_ring_destroy(gat_ctx *ctx);

/* ---- resident sessions: one call + synchronisation per block WITHOUT a kernel launch --------------------------------
 * The reference times one `CUDA.@sync kernel_algorithm(...)` per 1 ms block (src/benchmarks.jl:872, paper/paper.tex:150);
 * a tracking loop makes exactly that call every millisecond.  Through a kernel launch it costs ~18 us on a B200, most of
 * it launch, argument marshalling and completion latency.  A resident session launches the correlate kernel ONCE: it
 * stays on the device, polls a command the host writes into pinned mapped memory, runs the same fused
 * downconvert-and-correlate (same plan, bit-identical sums to gat_correlate) and stores every accumulator straight into
 * host memory together with the command's sequence number; the caller spins until all of them carry it.
 *   gat_resident_begin     fixes the shape (slots that hold blocks of the same geometry, channel count <= 32, sampling
 *                          rate, taps, sample range; `channels` = representative channels: systems / code rates) and
 *                          launches the kernel.  Classes: 1 / 4 / 16 antennas with <= 3 or 7 taps, 16 antennas x 11 taps
 *                          (GAT_ERR_UNSUPPORTED otherwise), integer-NCO code phase, FP32 planes.
 *   gat_resident_correlate one command: the block in slots[slot_index], n_sats channels (any PRN / phases / Doppler of the
 *                          planned systems) -> out [n_ants x n_taps x n_sats] host arrays.  Synchronous.
 *   gat_resident_end       ends the kernel and frees the session (also done by gat_destroy).
 * While a session is open the kernel owns every SM: gat_correlate* on this ctx return GAT_ERR_INVALID, kernels of OTHER
 * contexts or libraries on the device wait; copies (gat_upload_signal of FP32 host data into the session's slots, on
 * the ctx stream) run on the copy engines and are fine.  To bound that wait the kernel leaves by itself after
 * GAT_RESIDENT_IDLE_MS (environment, default 2000) without a command; the next gat_resident_correlate starts it again. */
int gat_resident_begin(gat_ctx *ctx, const int32_t *slots, int n_slots, int n_sats, const gat_channel *channels, double fs_hz,
                       const int32_t *sample_shifts, int n_taps, int start_sample, int n_samples);
int gat_resident_correlate(gat_ctx *ctx, int slot_index, const gat_channel *channels, float *out_re, float *out_im);
int gat_resident_end(gat_ctx *ctx);

/* ---- one host process, all GPUs of the box -----------------------------------------------------
 * The call a single tracking-loop process makes (Julia: one `ccall` per integration period or batch); the accumulators
 * come back in the caller's channel order.  gat_mg_upload_signal sends each device ITS sample range of the block through
 * that device's own PCIe link (asynchronous, n_dev links in parallel).  By default (gat_mg_set_sharding) every device then
 * correlates ALL channels over that range and the host adds the devices' partial sums -- no signal crosses NVLink; in
 * satellite mode the channels are partitioned and the correlate kernels gather the other ranges over NVLink (signal ring).  `devices` may name one device several times (logical shards; used by the
 * single-GPU tests).  Typical loop: upload block t+1 into slot (t+1) % n_slots, THEN gat_mg_correlate on slot t % n_slots:
 * the upload of the next block overlaps the kernels of the current one; a slot is not overwritten before every
 * device has finished reading it (tracked per slot).
 *   out: [n_ants x n_taps x n_sats x n_periods] host arrays; gat_mg_correlate is synchronous. */
typedef struct gat_mg gat_mg;
int gat_mg_create(gat_mg **out, int n_dev, const int *devices);
/* How gat_mg_correlate splits a call over the devices.  0 (default) = SAMPLES: every device correlates all channels over the
 * sample range of the block it already holds and the host adds the devices' partial sums -- no signal crosses NVLink
 * (bit-exact chip indices; FP32 summation order differs from a one-device call).  1 = SATELLITES: the channels are
 * partitioned, every device reads the whole block, the other devices' ranges over NVLink inside the kernel. */
int gat_mg_set_sharding(gat_mg *mg, int mode);
int gat_mg_destroy(gat_mg *mg);
const char *gat_mg_last_error(gat_mg *mg);
int gat_mg_device_count(gat_mg *mg);
gat_ctx *gat_mg_ctx(gat_mg *mg, int i);            /* device i's context (borrowed) for per-device calls, e.g. gat_last_launch_info */
int gat_mg_set_codes(gat_mg *mg, int system_id, const int8_t *chips, int code_len, int n_prn);   /* on every device */
int gat_mg_configure(gat_mg *mg, int n_slots, int n_samples, int n_ants);                          /* (re)builds the ring */
int gat_mg_upload_signal(gat_mg *mg, int slot, const float *h_re, const float *h_im, int ld);     /* host planes [n_ants x ld] */
int gat_mg_correlate(gat_mg *mg, int n_periods, const int32_t *slots, int n_sats, const gat_channel *channels,
                     double fs_hz, const int32_t *sample_shifts, int n_taps, int start_sample, int n_samples,
                     float *h_out_re, float *h_out_im, unsigned flags);
int gat_mg_sync(gat_mg *mg);

/* ---- introspection (bench / tests) ----------------------------------------------------- */
typedef struct gat_launch_info {
    int32_t grid, block, smem_bytes;
    int32_t ants_per_thread, ant_groups, sats_per_cta, sample_slices, consumer_warps;
    int32_t sat_groups, chunks_per_job, chunk_len, tile_len, stages, items;
    int32_t kernels_launched;   /* kernels of OURS enqueued by the last correlate call */
    int32_t sc16;               /* 1 if the kernel read raw int16 I/Q words (gat_upload_signal_sc16 slots) */
    int32_t tensor;             /* 1 if the call ran on the tensor-core path (GAT_TENSOR_TF32) */
    float last_kernel_ms;       /* device time of the last correlate kernel if timing enabled */
} gat_launch_info;
int gat_last_launch_info(gat_ctx *ctx, gat_launch_info *out);
/* Host-only planner probe: NO CUDA call, works without a device.  The launch plan gat_correlate_batch would choose for
 * n_periods blocks x n_sats channels (all of one code length and chip rate) on a device with n_sm SMs and the given
 * gat_set_max_ctas value: grid, CTA size, dynamic shared memory, antennas per thread, satellites per CTA, sample slices,
 * ring stages.  flags: GAT_CODE_PHASE_F64, GAT_PROBE_INT16 (the slots hold raw int16 words and the kernel reads them),
 * GAT_PROBE_RESIDENT (plan of a resident session).  Returns GAT_OK, or the status gat_correlate_batch would return for
 * a shape it cannot plan, with the message copied into err (may be NULL).  Used by the CPU test tier to sweep the
 * planner's invariants (shared-memory budget, warps per CTA, slice / stage divisibility) over the supported shapes, and
 * by hosts that size their channel batches before a device is attached. */
#define GAT_PROBE_INT16 0x10000u
#define GAT_PROBE_RESIDENT 0x20000u
int gat_plan_probe(int n_sm, int max_ctas, int n_periods, int n_sats, int n_ants, int n_taps, const int32_t *sample_shifts,
                   int start_sample, int n_samples, double fs_hz, double code_freq_hz, int code_len, unsigned flags,
                   gat_launch_info *out, char *err, int err_cap);
/* The correlate kernel is persistent: one CTA per SM, all SMs.  A communication kernel that should run
 * CONCURRENTLY (an NCCL broadcast of the next signal blocks, ...) then finds no free SM and the two serialise.
 * max_ctas > 0 caps the grid so that the remaining SMs stay free (measured, 2 GPUs, NCCL broadcast of the next
 * blocks under the kernel: 53 -> 34 us per period with 116 of 148 SMs); 0 restores the default. */
int gat_set_max_ctas(gat_ctx *ctx, int max_ctas);
int gat_set_timing(gat_ctx *ctx, int enable);  /* record cudaEvents around the correlate kernel */
uint64_t gat_kernel_launch_count(gat_ctx *ctx); /* total kernels of ours launched on this ctx */
/* Debug timeline: when enabled, every CTA of the next correlate launches stamps %globaltimer (ns) into
 * 16 slots: consumer warp 0 -> [0] kernel entry, [1] setup done, [2] first tile landed, [3] last tile
 * consumed, [4] partials published, [5] grid barrier passed, [6] exit; producer warp -> [8] entry,
 * [9] setup done, [10] chip tables cached, [11] first tile issued and windows built, [12] all tiles issued.
 * gat_get_timeline syncs and copies [n_ctas x 16] stamps of the LAST launch; returns n_ctas or <0. */
int gat_set_timeline(gat_ctx *ctx, int enable);
int gat_get_timeline(gat_ctx *ctx, uint64_t *out, int cap_ctas);
/* Replica chip-table indices exactly as the HOT kernel computes them, for the bit-exactness tests: the correlate call
 * is really made (a debug instantiation of the same kernel over an all-zero block of n_ants antennas) and the index of
 * every replica entry it generated is read back -- the first tile's base from scratch, the following tiles through the
 * per-tile NCO advance, through whichever wrap branch the launch plan selected, or the Float64 formula with
 * GAT_CODE_PHASE_F64.  out[l * n_samples + i] = index used for tap l at sample start_sample + i (host memory).
 * Shapes: (n_ants, n_taps) classes (1, 3) (16, 3) (8..16, 5) (4..16, 11) -- one per accumulator class of the kernel family. */
int gat_debug_replica_indices(gat_ctx *ctx, const gat_channel *ch, double fs_hz, const int32_t *sample_shifts,
                              int n_taps, int n_ants, int start_sample, int n_samples, unsigned flags, int32_t *out);
/* one tap, one antenna: out[n_samples] */
int gat_debug_chip_indices(gat_ctx *ctx, const gat_channel *ch, double fs_hz, int shift,
                           int n_samples, unsigned flags, int32_t *out);
/* The tensor-core kernel's replica SIGN BITS (it keeps no indices): 1 = chip -1.  out[(k * n_taps + l) * n_samples + i]
 * for channel k, tap l, sample start_sample + i of a gat_correlate call with GAT_TENSOR_TF32 over slot `slot`; both
 * generators (32-chip window / table lookup per entry; force the latter with the environment variable GAT_TC_NO_WINDOW=1). */
int gat_debug_tc_replica_bits(gat_ctx *ctx, int slot, int n_sats, const gat_channel *channels, double fs_hz,
                              const int32_t *sample_shifts, int n_taps, int start_sample, int n_samples, uint8_t *out);

#ifdef __cplusplus
}
#endif
#endif /* GAT_H */
