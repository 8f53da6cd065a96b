/*
 * oracle.h -- CPU restatement of the reference's correlate hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker (or as the timed CPU
 * baseline), never as the thing shipped.
 *
 * The reference (coezmaden/GPUAcceleratedTracking) is Julia; `julia` is not in
 * this image and the arithmetic of its CPU path lives in un-vendored packages
 * (Tracking.jl fork v0.14.8, GNSSSignals.jl v0.15.4 -- Manifest.toml:1392-1398,
 * :440-444), so the reference cannot be built into oracle/_ref.  Every function
 * below cites the reference file:line (relative to /root/reference) or the
 * upstream package behaviour it restates.
 *
 * Pinning status (see DESIGN.md "Oracle"):
 *   - L1 C/A correlate path: PINNED by the reference's own known answer
 *     [1476, 2500, 1476] (test/algorithms.jl:85-86 and 10 more sites).
 *   - C/A code content: PINNED by IS-GPS-200 first-10-chip octal words.
 *   - GPS L5 code content, track()/loop filters: PARITY UNPINNED (no reference
 *     test or fixture touches them; restated from the published ICDs / upstream).
 */
#ifndef GAT_ORACLE_H
#define GAT_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- PRN code tables (GNSSSignals.jl `system.codes`, +-1 chips) ---------------- */
/* chip mapping: logic 0 -> +1, logic 1 -> -1.  Returns code length or <0. */
int orc_gps_l1_ca(int prn, int8_t *out, int cap);   /* 1023 chips, PRN 1..37 (IS-GPS-200) */
int orc_gps_l5_i5(int prn, int8_t *out, int cap);   /* 10230 chips, PRN 1..37 (IS-GPS-705) */

/* ---- get_correlator_sample_shifts [upstream Tracking.jl]; call sites
 *      src/benchmarks.jl:52,106, test/algorithms.jl:16 ---------------------------- */
int orc_sample_shifts(double code_freq_hz, double fs_hz, double preferred_shift_chips,
                      int n_taps, int32_t *out);

/* ---- gen_signal (src/gen_signal.jl:64-70, :86-90) ------------------------------ */
/* re/im are column-major [ld x n_ants] (n fastest), every antenna identical.      */
void orc_gen_signal(const int8_t *code, int code_len, double code_freq_hz,
                    double carrier_freq_hz, double fs_hz, double start_code_phase,
                    double start_carrier_phase_rad, int n_samples, int n_ants, int ld,
                    float *re, float *im);

/* ---- chip index sequences ------------------------------------------------------ */
/* GPU-kernel form: mod(floor(fc/fs*(i+shift)+phase), Lc), Float64
 * (src/algorithms.jl:179-182, src/gen_signal.jl:64-65).  i = 0..n-1.             */
void orc_chip_index_f64(double code_freq_hz, double fs_hz, double code_phase,
                        int code_len, int shift, int n, int32_t *out);
/* Tracking.jl CPU form: Int64 Q-format NCO [upstream gen_code_replica!]:
 *   fp = 63 - ceil(log2(Lc)); delta = floor(fc*2^fp/fs);
 *   start = floor(mod(phase, Lc)*2^fp); idx = ((i+shift)*delta + start) >> fp, wrapped. */
void orc_chip_index_nco(double code_freq_hz, double fs_hz, double code_phase,
                        int code_len, int shift, int n, int32_t *out);

/* ---- correlate, semantic oracle: kernel-1330 formula (src/algorithms.jl:170-187)
 *      evaluated in double precision.  code_mode 0 = f64 chip index, 1 = NCO.
 *      out_re/out_im are [n_ants x n_taps] column-major (antenna fastest), overwritten. */
void orc_correlate_direct(const float *re, const float *im, int ld, int n_ants,
                          int start_sample, int n_samples,
                          const int8_t *code, int code_len,
                          double code_freq_hz, double code_phase,
                          double carrier_freq_hz, double carrier_phase_cycles,
                          double fs_hz, const int32_t *shifts, int n_taps,
                          int code_mode, double *out_re, double *out_im);

/* ---- correlate, Tracking.jl CPU structure (call site src/benchmarks.jl:63-79;
 *      correlate loop paper/paper.tex:286-293): four serial passes in Float32 over
 *      caller-owned scratch.  This is the timed CPU baseline.
 *      scratch sizes: code_rep n+span floats, car_re/car_im n floats, dw_re/dw_im n*n_ants. */
void orc_correlate_tracking(const float *re, const float *im, int ld, int n_ants,
                            int start_sample, int n_samples,
                            const int8_t *code, int code_len,
                            double code_freq_hz, double code_phase,
                            double carrier_freq_hz, double carrier_phase_cycles,
                            double fs_hz, const int32_t *shifts, int n_taps,
                            float *code_rep, float *car_re, float *car_im,
                            float *dw_re, float *dw_im,
                            float *out_re, float *out_im);

/* batch of independent (period, satellite) jobs, OpenMP over jobs.
 * Signals: period p at re + p*period_stride.  Channel arrays are [n_periods*n_sats].
 * out: [n_ants x n_taps x n_sats x n_periods].  Returns threads used. */
int orc_correlate_tracking_batch(const float *re, const float *im, int64_t period_stride,
                                 int ld, int n_ants, int n_samples, int n_periods, int n_sats,
                                 const int8_t *const *codes, const int32_t *code_lens,
                                 const double *code_freq_hz, const double *code_phase,
                                 const double *carrier_freq_hz, const double *carrier_phase_cycles,
                                 double fs_hz, const int32_t *shifts, int n_taps,
                                 int n_threads, float *out_re, float *out_im);

/* minimum wall time [ns] over `reps` single-thread orc_correlate_tracking calls (timed CPU baseline of the
 * reference's per-call sweep, src/benchmarks.jl:63-79 + BenchmarkTools minimum, paper/paper.tex:150) */
double orc_time_tracking(const float *re, const float *im, int ld, int n_ants, int n_samples,
                         const int8_t *code, int code_len, double code_freq_hz, double code_phase,
                         double carrier_freq_hz, double carrier_phase_cycles, double fs_hz,
                         const int32_t *shifts, int n_taps, int reps, float *out_re, float *out_im);

/* ---- tracking loop pieces [upstream Tracking.jl / TrackingLoopFilters.jl, SURVEY A.3]
 *      PARITY UNPINNED.  State layout documented in oracle.c. ---------------------- */
typedef struct {
    double carrier_doppler, code_doppler;      /* Hz */
    double carrier_phase, code_phase;          /* cycles, chips */
    double pll_x1, pll_x2, dll_x1;             /* loop filter states */
    double init_carrier_doppler, init_code_doppler;
} orc_track_state;

void orc_loop_update(orc_track_state *st, const double *prompt_re_im, const double *early_re_im,
                     const double *late_re_im, double early_late_spacing_chips, double dt_s,
                     double code_freq_hz, double center_freq_hz, double pll_bw_hz, double dll_bw_hz);

#ifdef __cplusplus
}
#endif
#endif
