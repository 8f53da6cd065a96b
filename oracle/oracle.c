/*
 * oracle.c -- CPU restatement of the reference's correlate hot path (see oracle.h).
 * TEST INFRASTRUCTURE ONLY: checker + timed CPU baseline, never the product path.
 *
 * Build: make -C oracle   (gcc -O3 -march=native -fopenmp -shared -fPIC)
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* =========================================================================
 * PRN codes.  The reference reads them from GNSSSignals.jl (`system.codes`,
 * src/benchmarks.jl:93, src/gen_signal.jl:65); that package is not vendored,
 * so the tables are regenerated from the public ICDs.
 * ========================================================================= */

/* IS-GPS-200 Table 3-Ia: G2 phase-selector taps for PRN 1..37. */
static const unsigned char CA_TAPS[37][2] = {
    {2, 6},  {3, 7},  {4, 8},  {5, 9},  {1, 9},  {2, 10}, {1, 8},  {2, 9},  {3, 10}, {2, 3},
    {3, 4},  {5, 6},  {6, 7},  {7, 8},  {8, 9},  {9, 10}, {1, 4},  {2, 5},  {3, 6},  {4, 7},
    {5, 8},  {6, 9},  {1, 3},  {4, 6},  {5, 7},  {6, 8},  {7, 9},  {8, 10}, {1, 6},  {2, 7},
    {3, 8},  {4, 9},  {5, 10}, {4, 10}, {1, 7},  {2, 8},  {4, 10}};

int orc_gps_l1_ca(int prn, int8_t *out, int cap)
{
    if (prn < 1 || prn > 37 || cap < 1023) return -1;
    int g1[11], g2[11]; /* stages 1..10 */
    for (int i = 1; i <= 10; ++i) g1[i] = g2[i] = 1;
    const int s1 = CA_TAPS[prn - 1][0], s2 = CA_TAPS[prn - 1][1];
    for (int k = 0; k < 1023; ++k) {
        const int bit = g1[10] ^ g2[s1] ^ g2[s2];
        out[k] = (int8_t)(bit ? -1 : 1);
        const int f1 = g1[3] ^ g1[10];
        const int f2 = g2[2] ^ g2[3] ^ g2[6] ^ g2[8] ^ g2[9] ^ g2[10];
        for (int i = 10; i > 1; --i) { g1[i] = g1[i - 1]; g2[i] = g2[i - 1]; }
        g1[1] = f1;
        g2[1] = f2;
    }
    return 1023;
}

/* IS-GPS-705 Table 3-Ia: XB code advance (chips) of the I5 code, PRN 1..37. */
static const short L5I_ADVANCE[37] = {
    266,  365,  804,  1138, 1509, 1559, 1756, 2084, 2170, 2303, 2527, 2687, 2930,
    3471, 3940, 4132, 4332, 4924, 5343, 5443, 5641, 5816, 5898, 5918, 5955, 6243,
    6345, 6477, 6518, 6875, 7168, 7187, 7329, 7577, 7720, 7777, 8057};

int orc_gps_l5_i5(int prn, int8_t *out, int cap)
{
    if (prn < 1 || prn > 37 || cap < 10230) return -1;
    /* XB natural sequence (period 8191), 1+x+x^3+x^4+x^6+x^7+x^8+x^12+x^13, all-ones start */
    static unsigned char xb_seq[8191];
    int xb[14], xa[14];
    for (int i = 1; i <= 13; ++i) xb[i] = 1;
    for (int k = 0; k < 8191; ++k) {
        xb_seq[k] = (unsigned char)xb[13];
        const int f = xb[1] ^ xb[3] ^ xb[4] ^ xb[6] ^ xb[7] ^ xb[8] ^ xb[12] ^ xb[13];
        for (int i = 13; i > 1; --i) xb[i] = xb[i - 1];
        xb[1] = f;
    }
    /* XA: 1+x^9+x^10+x^12+x^13, short-cycled to 8190 chips, restarted every 10230 */
    for (int i = 1; i <= 13; ++i) xa[i] = 1;
    const int adv = L5I_ADVANCE[prn - 1];
    for (int k = 0; k < 10230; ++k) {
        if (k == 8190)
            for (int i = 1; i <= 13; ++i) xa[i] = 1;
        const int a = xa[13];
        const int b = xb_seq[(k + adv) % 8191];
        out[k] = (int8_t)((a ^ b) ? -1 : 1);
        const int f = xa[9] ^ xa[10] ^ xa[12] ^ xa[13];
        for (int i = 13; i > 1; --i) xa[i] = xa[i - 1];
        xa[1] = f;
    }
    return 10230;
}

/* =========================================================================
 * get_correlator_sample_shifts  [upstream Tracking.jl v0.14.x; SURVEY App. A.1]
 *   s = max(1, round(Int, preferred_code_shift * fs / code_frequency))
 *   shifts = (-(L-1)/2 : (L-1)/2) .* s          (index 1 = latest / most negative)
 * Julia's round() is round-half-to-even == nearbyint() in the default mode.
 * ========================================================================= */
int orc_sample_shifts(double code_freq_hz, double fs_hz, double preferred_shift_chips,
                      int n_taps, int32_t *out)
{
    if (n_taps < 1 || (n_taps & 1) == 0) return -1;
    long s = (long)nearbyint(preferred_shift_chips * fs_hz / code_freq_hz);
    if (s < 1) s = 1;
    const int half = (n_taps - 1) / 2;
    for (int l = 0; l < n_taps; ++l) out[l] = (int32_t)((l - half) * s);
    return 0;
}

static inline int64_t floormod_i64(int64_t a, int64_t m)
{
    int64_t r = a % m;
    return r < 0 ? r + m : r;
}

/* =========================================================================
 * gen_signal  (src/gen_signal.jl:64-70 vector, :86-90 CPU matrix)
 *   code_phases   = fc/fs .* (0:N-1) .+ start_code_phase            (Float64)
 *   chips         = codes[1 .+ mod.(floor.(Int, code_phases), Lc), prn]
 *   carrier_phases= Float32(2pi*(0:N-1)*f/fs .+ start_carrier_phase) (Float64 -> Float32)
 *   re = cos.(carrier_phases) .* chips ; im = sin.(...) .* chips ; same for every antenna
 * ========================================================================= */
void orc_gen_signal(const int8_t *code, int code_len, double code_freq_hz,
                    double carrier_freq_hz, double fs_hz, double start_code_phase,
                    double start_carrier_phase_rad, int n_samples, int n_ants, int ld,
                    float *re, float *im)
{
    const double ratio = code_freq_hz / fs_hz;
    const double two_pi = 2.0 * M_PI;
    for (int i = 0; i < n_samples; ++i) {
        const double cp = ratio * (double)i + start_code_phase;
        const int64_t idx = floormod_i64((int64_t)floor(cp), code_len);
        const float chip = (float)code[idx];
        /* Julia: 2pi * i * f / fs  evaluated left to right */
        const float ph = (float)(two_pi * (double)i * carrier_freq_hz / fs_hz + start_carrier_phase_rad);
        re[i] = cosf(ph) * chip;
        im[i] = sinf(ph) * chip;
    }
    for (int m = 1; m < n_ants; ++m) {
        memcpy(re + (size_t)m * ld, re, sizeof(float) * (size_t)n_samples);
        memcpy(im + (size_t)m * ld, im, sizeof(float) * (size_t)n_samples);
    }
}

/* =========================================================================
 * chip index, two forms
 * ========================================================================= */
void orc_chip_index_f64(double code_freq_hz, double fs_hz, double code_phase,
                        int code_len, int shift, int n, int32_t *out)
{
    /* src/algorithms.jl:179-182: code_frequency / sampling_frequency * ((sample_idx-1)+shift) + phase */
    const double ratio = code_freq_hz / fs_hz;
    for (int i = 0; i < n; ++i) {
        const double cp = ratio * (double)((int64_t)i + shift) + code_phase;
        out[i] = (int32_t)floormod_i64((int64_t)floor(cp), code_len);
    }
}

static int nco_fixed_point(int code_len)
{
    /* fixed_point = 64 - 1 - ceil(log2(code_length * secondary_length)) */
    int bits = 0;
    while ((1LL << bits) < (long long)code_len) ++bits;
    return 63 - bits;
}

typedef struct {
    int fp;
    int64_t delta, start;
} nco_t;

static nco_t nco_make(double code_freq_hz, double fs_hz, double code_phase, int code_len)
{
    nco_t n;
    n.fp = nco_fixed_point(code_len);
    /* delta = floor(Int, code_frequency * 1 << fixed_point / sampling_frequency) */
    n.delta = (int64_t)floor(code_freq_hz * ldexp(1.0, n.fp) / fs_hz);
    double modded = fmod(code_phase, (double)code_len);
    if (modded < 0) modded += (double)code_len;
    n.start = (int64_t)floor(modded * ldexp(1.0, n.fp));
    return n;
}

static inline int32_t nco_index(const nco_t *n, int64_t k, int code_len)
{
    /* 128-bit so the value equals Julia's Int64 result wherever that does not
     * overflow and stays mathematically right where it would. */
    const __int128 v = (__int128)k * n->delta + n->start;
    const int64_t idx = (int64_t)(v >> n->fp);
    return (int32_t)floormod_i64(idx, code_len);
}

void orc_chip_index_nco(double code_freq_hz, double fs_hz, double code_phase,
                        int code_len, int shift, int n, int32_t *out)
{
    const nco_t nco = nco_make(code_freq_hz, fs_hz, code_phase, code_len);
    for (int i = 0; i < n; ++i) out[i] = nco_index(&nco, (int64_t)i + shift, code_len);
}

/* =========================================================================
 * Semantic oracle: the fully fused formula of kernel 1330
 * (src/algorithms.jl:170-187), every operation in double:
 *   carrier = sincos(2pi*((n)*f/fs + phi))
 *   dw      = s * conj(carrier)
 *   acc[m,l] += codes[1+mod(floor(fc/fs*(n+shift_l)+phi_c), Lc)] * dw
 * ========================================================================= */
void orc_correlate_direct(const float *re, const float *im, int ld, int n_ants,
                          int start_sample, int n_samples,
                          const int8_t *code, int code_len,
                          double code_freq_hz, double code_phase,
                          double carrier_freq_hz, double carrier_phase_cycles,
                          double fs_hz, const int32_t *shifts, int n_taps,
                          int code_mode, double *out_re, double *out_im)
{
    const double two_pi = 2.0 * M_PI;
    const double ratio = code_freq_hz / fs_hz;
    const nco_t nco = nco_make(code_freq_hz, fs_hz, code_phase, code_len);
    for (int i = 0; i < n_ants * n_taps; ++i) out_re[i] = out_im[i] = 0.0;
    double chips[64];
    for (int n = 0; n < n_samples; ++n) {
        const double ph = two_pi * ((double)n * carrier_freq_hz / fs_hz + carrier_phase_cycles);
        const double cr = cos(ph), ci = sin(ph);
        for (int l = 0; l < n_taps; ++l) {
            int32_t idx;
            if (code_mode == 0) {
                const double cp = ratio * (double)((int64_t)n + shifts[l]) + code_phase;
                idx = (int32_t)floormod_i64((int64_t)floor(cp), code_len);
            } else {
                idx = nco_index(&nco, (int64_t)n + shifts[l], code_len);
            }
            chips[l] = (double)code[idx];
        }
        for (int m = 0; m < n_ants; ++m) {
            const double sr = re[(size_t)m * ld + start_sample + n];
            const double si = im[(size_t)m * ld + start_sample + n];
            const double dr = sr * cr + si * ci;
            const double di = si * cr - sr * ci;
            for (int l = 0; l < n_taps; ++l) {
                out_re[l * n_ants + m] += chips[l] * dr;
                out_im[l * n_ants + m] += chips[l] * di;
            }
        }
    }
}

/* =========================================================================
 * Tracking.jl CPU structure: four serial Float32 passes.
 * ========================================================================= */

/* Float32 sincos good to ~1 ulp on |x| <= a few thousand rad: Cody-Waite reduction by
 * pi/2 then minimax polynomials.  Written branch-free so gcc vectorises the carrier
 * pass the way LoopVectorization's @avx + SLEEF does for the reference. */
static inline void sincos_f32(float x, float *s, float *c)
{
    const float two_over_pi = 0.636619772367581343f;
    const float q = nearbyintf(x * two_over_pi);
    const int qi = (int)q;
    /* pi/2 split in three parts */
    float r = fmaf(q, -1.5707962512969971f, x);
    r = fmaf(q, -7.5497894158615964e-08f, r);
    r = fmaf(q, -5.3903029534742384e-15f, r);
    const float r2 = r * r;
    float sp = fmaf(r2, 2.6083159809786593541503e-06f, -1.981069071916863322258e-04f);
    sp = fmaf(sp, r2, 8.33307858556509017944336e-03f);
    sp = fmaf(sp, r2, -1.66666597127914428710938e-01f);
    sp = fmaf(sp * r2, r, r);
    float cp = fmaf(r2, -2.6051615e-07f, 2.4760495e-05f);
    cp = fmaf(cp, r2, -1.3888378e-03f);
    cp = fmaf(cp, r2, 4.1666638e-02f);
    cp = fmaf(cp, r2, -0.5f);
    cp = fmaf(cp, r2, 1.0f);
    const int swap = qi & 1;
    float ss = swap ? cp : sp;
    float cc = swap ? sp : cp;
    ss = (qi & 2) ? -ss : ss;
    cc = ((qi + 1) & 2) ? -cc : cc;
    *s = ss;
    *c = cc;
}

void orc_correlate_tracking(const float *re, const float *im, int ld, int n_ants,
                            int start_sample, int n_samples,
                            const int8_t *code, int code_len,
                            double code_freq_hz, double code_phase,
                            double carrier_freq_hz, double carrier_phase_cycles,
                            double fs_hz, const int32_t *shifts, int n_taps,
                            float *code_rep, float *car_re, float *car_im,
                            float *dw_re, float *dw_im,
                            float *out_re, float *out_im)
{
    const int span = shifts[n_taps - 1] - shifts[0];
    /* pass 1: gen_code_replica!  (Int64 fixed-point NCO, replica covers n + span samples) */
    {
        const nco_t nco = nco_make(code_freq_hz, fs_hz, code_phase, code_len);
        /* incremental form of ((i + shifts[0]) * delta + start) >> fp, wrapped */
        __int128 v = (__int128)shifts[0] * nco.delta + nco.start;
        const __int128 wrap = (__int128)code_len << nco.fp;
        v %= wrap;
        if (v < 0) v += wrap;
        uint64_t acc = (uint64_t)v;           /* < code_len * 2^fp < 2^63 */
        const uint64_t uwrap = (uint64_t)wrap;
        const uint64_t dmod = (uint64_t)nco.delta % uwrap;
        for (int i = 0; i < n_samples + span; ++i) {
            code_rep[i] = (float)code[acc >> nco.fp];
            acc += dmod;
            if (acc >= uwrap) acc -= uwrap;
        }
    }
    /* pass 2: gen_carrier_replica!  sincos(T(2pi) * (i*T(f)/T(fs) + T(phase))), T = Float32 */
    {
        const float two_pi = (float)(2.0 * M_PI);
        const float f = (float)carrier_freq_hz, fs = (float)fs_hz, ph0 = (float)carrier_phase_cycles;
#pragma omp simd
        for (int i = 0; i < n_samples; ++i) {
            float s, c;
            sincos_f32(two_pi * ((float)i * f / fs + ph0), &s, &c);
            car_re[i] = c;
            car_im[i] = s;
        }
    }
    /* pass 3: downconvert!  d = s * conj(c) */
    for (int m = 0; m < n_ants; ++m) {
        const float *sr = re + (size_t)m * ld + start_sample;
        const float *si = im + (size_t)m * ld + start_sample;
        float *dr = dw_re + (size_t)m * n_samples;
        float *di = dw_im + (size_t)m * n_samples;
#pragma omp simd
        for (int i = 0; i < n_samples; ++i) {
            dr[i] = sr[i] * car_re[i] + si[i] * car_im[i];
            di[i] = si[i] * car_re[i] - sr[i] * car_im[i];
        }
    }
    /* pass 4: correlate (paper/paper.tex:286-293): a[l] += d[i] * code[i + shift_l - shift_1] */
    for (int m = 0; m < n_ants; ++m) {
        const float *dr = dw_re + (size_t)m * n_samples;
        const float *di = dw_im + (size_t)m * n_samples;
        for (int l = 0; l < n_taps; ++l) {
            const float *cr = code_rep + (shifts[l] - shifts[0]);
            float ar = 0.f, ai = 0.f;
#pragma omp simd reduction(+ : ar, ai)
            for (int i = 0; i < n_samples; ++i) {
                ar += dr[i] * cr[i];
                ai += di[i] * cr[i];
            }
            out_re[l * n_ants + m] = ar;
            out_im[l * n_ants + m] = ai;
        }
    }
}

int orc_correlate_tracking_batch(const float *re, const float *im, int64_t period_stride,
                                 int ld, int n_ants, int n_samples, int n_periods, int n_sats,
                                 const int8_t *const *codes, const int32_t *code_lens,
                                 const double *code_freq_hz, const double *code_phase,
                                 const double *carrier_freq_hz, const double *carrier_phase_cycles,
                                 double fs_hz, const int32_t *shifts, int n_taps,
                                 int n_threads, float *out_re, float *out_im)
{
    const int span = shifts[n_taps - 1] - shifts[0];
    const int jobs = n_periods * n_sats;
    int used = 1;
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
    if (n_threads > jobs) n_threads = jobs;
    used = n_threads;
#pragma omp parallel num_threads(n_threads)
#endif
    {
        float *code_rep = (float *)malloc(sizeof(float) * (size_t)(n_samples + span + 8));
        float *car_re = (float *)malloc(sizeof(float) * (size_t)n_samples);
        float *car_im = (float *)malloc(sizeof(float) * (size_t)n_samples);
        float *dw_re = (float *)malloc(sizeof(float) * (size_t)n_samples * n_ants);
        float *dw_im = (float *)malloc(sizeof(float) * (size_t)n_samples * n_ants);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
        for (int j = 0; j < jobs; ++j) {
            const int p = j / n_sats;
            orc_correlate_tracking(re + (size_t)p * period_stride, im + (size_t)p * period_stride,
                                   ld, n_ants, 0, n_samples, codes[j], code_lens[j],
                                   code_freq_hz[j], code_phase[j], carrier_freq_hz[j],
                                   carrier_phase_cycles[j], fs_hz, shifts, n_taps,
                                   code_rep, car_re, car_im, dw_re, dw_im,
                                   out_re + (size_t)j * n_ants * n_taps,
                                   out_im + (size_t)j * n_ants * n_taps);
        }
        free(code_rep); free(car_re); free(car_im); free(dw_re); free(dw_im);
    }
    return used;
}

/* Minimum wall time [ns] of `reps` single-thread orc_correlate_tracking calls on caller-owned scratch: the
 * estimator of the reference's harness (BenchmarkTools minimum, paper/paper.tex:150; call site
 * src/benchmarks.jl:63-79), measured in C so that no interpreter overhead enters. */
#include <time.h>
double orc_time_tracking(const float *re, const float *im, int ld, int n_ants, int n_samples,
                         const int8_t *code, int code_len, double code_freq_hz, double code_phase,
                         double carrier_freq_hz, double carrier_phase_cycles, double fs_hz,
                         const int32_t *shifts, int n_taps, int reps, float *out_re, float *out_im)
{
    const int span = shifts[n_taps - 1] - shifts[0];
    float *code_rep = (float *)malloc(sizeof(float) * (size_t)(n_samples + span + 8));
    float *car_re = (float *)malloc(sizeof(float) * (size_t)n_samples);
    float *car_im = (float *)malloc(sizeof(float) * (size_t)n_samples);
    float *dw_re = (float *)malloc(sizeof(float) * (size_t)n_samples * n_ants);
    float *dw_im = (float *)malloc(sizeof(float) * (size_t)n_samples * n_ants);
    double best = 1e300;
    for (int r = 0; r < reps + 2; ++r) {   /* two untimed warm-up calls */
        struct timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        orc_correlate_tracking(re, im, ld, n_ants, 0, n_samples, code, code_len, code_freq_hz, code_phase,
                               carrier_freq_hz, carrier_phase_cycles, fs_hz, shifts, n_taps,
                               code_rep, car_re, car_im, dw_re, dw_im, out_re, out_im);
        clock_gettime(CLOCK_MONOTONIC, &t1);
        const double ns = (double)(t1.tv_sec - t0.tv_sec) * 1e9 + (double)(t1.tv_nsec - t0.tv_nsec);
        if (r >= 2 && ns < best) best = ns;
    }
    free(code_rep); free(car_re); free(car_im); free(dw_re); free(dw_im);
    return best;
}

/* =========================================================================
 * Loop closure [upstream Tracking.jl track() + TrackingLoopFilters.jl; SURVEY App. A.3]
 * PARITY UNPINNED: no reference test covers it and the source is not in tree.
 *   pll_disc = atan(Q_P / I_P)                         (Costas, cycles = /2pi)
 *   dll_disc = (|E|-|L|)/(|E|+|L|) / (2*(2-d))         d = early-late spacing in chips
 *   3rd-order bilinear PLL, 2nd-order bilinear DLL (Kaplan & Hegarty Table 5.6)
 * ========================================================================= */
void orc_loop_update(orc_track_state *st, const double *p, const double *e, const double *l,
                     double d, double dt, double code_freq_hz, double center_freq_hz,
                     double pll_bw_hz, double dll_bw_hz)
{
    const double pll_disc = atan(p[1] / p[0]) / (2.0 * M_PI);
    const double ea = hypot(e[0], e[1]), la = hypot(l[0], l[1]);
    const double dll_disc = (ea - la) / (ea + la) / (2.0 * (2.0 - d));
    /* third order bilinear */
    {
        const double w0 = pll_bw_hz * 1.2;
        const double w02 = w0 * w0, w03 = w02 * w0;
        const double x1 = st->pll_x1, x2 = st->pll_x2;
        const double out = x2 + 0.5 * dt * x1 + (0.25 * dt * dt * w03 + 0.55 * dt * w02 + 2.4 * w0) * pll_disc;
        st->pll_x1 = x1 + dt * w03 * pll_disc;
        st->pll_x2 = x2 + dt * (x1 + 0.5 * dt * w03 * pll_disc + 1.1 * w02 * pll_disc);
        st->carrier_doppler = out + st->init_carrier_doppler;
    }
    /* second order bilinear */
    {
        const double w0 = dll_bw_hz * 1.89;
        const double x1 = st->dll_x1;
        const double out = x1 + (0.5 * dt * w0 * w0 + sqrt(2.0) * w0) * dll_disc;
        st->dll_x1 = x1 + dt * w0 * w0 * dll_disc;
        st->code_doppler = out + st->carrier_doppler * code_freq_hz / center_freq_hz + st->init_code_doppler;
    }
}
