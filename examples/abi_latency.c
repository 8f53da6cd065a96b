/*
 * abi_latency.c -- plain-C caller of libgat (include/gat.h): the call sequence a host tracking loop makes,
 * timed per call.  No CUDA headers, no Python: what a `ccall` from Julia costs.
 *
 *   gcc -O2 -std=c11 -Iinclude examples/abi_latency.c -Lgpuacceleratedtracking_b200 -lgat \
 *       -Wl,-rpath,'$ORIGIN/../gpuacceleratedtracking_b200' -lm -o examples/abi_latency
 *   examples/abi_latency            (needs a B200)
 *
 * For every shape of the reference's sweep (scripts/run_benchmarks_gpsl1.jl:5-18: N = 2^11..2^18, M in {1, 4},
 * L in {3, 7}; plus M = 16) it prints one JSON line with the minimum wall time [ns] of
 *   resident : gat_correlate on a signal block already in HBM, accumulators returned to HOST memory
 *              (synchronous, = `CUDA.@sync kernel_algorithm(...)` + `Array(accum)`, src/benchmarks.jl:872)
 *   host     : gat_downconvert_and_correlate -- signal in HOST memory, upload + correlate + download (the CPU-style
 *              call of src/benchmarks.jl:63-79 pointed at the GPU)
 *   session  : gat_resident_correlate inside a resident session (the kernel stays on the device and is fed through mapped
 *              memory): the same synchronous call without a kernel launch; 0 where the class has no resident kernel
 */
#define _POSIX_C_SOURCE 199309L
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include "gat.h"

static double now_ns(void)
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec * 1e9 + (double)t.tv_nsec;
}

#define CHECK(call)                                                                                   \
    do {                                                                                              \
        int rc_ = (call);                                                                             \
        if (rc_ != GAT_OK) {                                                                          \
            fprintf(stderr, "%s -> %d (%s): %s\n", #call, rc_, gat_status_string(rc_), ctx ? gat_last_error(ctx) : ""); \
            return 1;                                                                                 \
        }                                                                                             \
    } while (0)

int main(int argc, char **argv)
{
    const int reps = argc > 1 ? atoi(argv[1]) : 300;
    gat_ctx *ctx = NULL;
    CHECK(gat_create(&ctx, 0));

    /* chip table of GPS L1 C/A, PRN 1..32 (GNSSSignals `GPSL1().codes`) */
    static int8_t table[1023 * 32];
    for (int prn = 1; prn <= 32; ++prn)
        if (gat_gen_code(GAT_GPSL1, prn, table + 1023 * (prn - 1), 1023) != 1023) return 2;
    CHECK(gat_set_codes(ctx, GAT_GPSL1, table, 1023, 32));

    static const int ants[] = {1, 4, 16}, taps[] = {3, 7};
    for (int im = 0; im < 3; ++im)
        for (int il = 0; il < 2; ++il)
            for (int e = 11; e <= 18; ++e) {
                const int M = ants[im], L = taps[il], N = 1 << e;
                const double fs = N / 1e-3;
                /* get_correlator_sample_shifts: half-chip spacing in samples */
                int step = (int)floor(0.5 * fs / 1.023e6 + 0.5);
                if (step < 1) step = 1;
                int32_t shifts[GAT_MAX_TAPS];
                for (int l = 0; l < L; ++l) shifts[l] = (l - (L - 1) / 2) * step;
                CHECK(gat_gen_signal(ctx, 0, GAT_GPSL1, 1, 1500.0, fs, 0.0, 0.0, N, M, 0.0, 0.0, 0, 0));
                const gat_channel ch = {GAT_GPSL1, 1, 0.0, 1.023e6, 0.0, 1500.0};
                float out_re[GAT_MAX_TAPS * 16], out_im[GAT_MAX_TAPS * 16];

                double best_res = 1e300;
                for (int r = 0; r < reps + 10; ++r) {
                    const double t0 = now_ns();
                    CHECK(gat_correlate(ctx, 0, 1, &ch, fs, shifts, L, 0, N, out_re, out_im, 0, 0));
                    const double dt = now_ns() - t0;
                    if (r >= 10 && dt < best_res) best_res = dt;
                }
                if (fabs(out_re[(L / 2) * M] - N) > 1e-3 * N) {
                    fprintf(stderr, "wrong prompt %f for N=%d\n", out_re[(L / 2) * M], N);
                    return 3;
                }

                /* a resident session: the kernel stays on the device, one command per call (gat_resident_*) */
                double best_session = 0.0, med_session = 0.0;
                {
                    const int32_t slot0 = 0;
                    int rc_b = gat_resident_begin(ctx, &slot0, 1, 1, &ch, fs, shifts, L, 0, N);
                    if (rc_b == GAT_OK) {
                        float s_re[GAT_MAX_TAPS * 16], s_im[GAT_MAX_TAPS * 16];
                        double *ts = malloc(sizeof(double) * (size_t)reps);
                        for (int r = 0; r < reps + 10; ++r) {
                            const double t0 = now_ns();
                            CHECK(gat_resident_correlate(ctx, 0, &ch, s_re, s_im));
                            if (r >= 10) ts[r - 10] = now_ns() - t0;
                        }
                        CHECK(gat_resident_end(ctx));
                        for (int i = 0; i < L * M; ++i)
                            if (s_re[i] != out_re[i] || s_im[i] != out_im[i]) {
                                fprintf(stderr, "resident session differs from the launched call at %d\n", i);
                                return 5;
                            }
                        /* min and median */
                        for (int i = 1; i < reps; ++i) {
                            double v = ts[i];
                            int j = i - 1;
                            while (j >= 0 && ts[j] > v) { ts[j + 1] = ts[j]; --j; }
                            ts[j + 1] = v;
                        }
                        best_session = ts[0];
                        med_session = ts[reps / 2];
                        free(ts);
                    } else if (rc_b != GAT_ERR_UNSUPPORTED) {
                        fprintf(stderr, "gat_resident_begin -> %d: %s\n", rc_b, gat_last_error(ctx));
                        return 6;
                    }
                }

                /* the same block from host memory through the CPU-style entry point */
                float *h_re = malloc(sizeof(float) * (size_t)N * M), *h_im = malloc(sizeof(float) * (size_t)N * M);
                CHECK(gat_download_signal(ctx, 0, h_re, h_im));
                double best_host = 1e300;
                const int hreps = reps / 4 + 5;
                for (int r = 0; r < hreps + 3; ++r) {
                    const double t0 = now_ns();
                    CHECK(gat_downconvert_and_correlate(ctx, h_re, h_im, N, M, 1, &ch, fs, shifts, L, 0, N, out_re, out_im, 0));
                    const double dt = now_ns() - t0;
                    if (r >= 3 && dt < best_host) best_host = dt;
                }
                free(h_re);
                free(h_im);
                if (fabs(out_re[(L / 2) * M] - N) > 1e-3 * N) return 4;
                printf("{\"system\": \"GPSL1\", \"num_samples\": %d, \"num_ants\": %d, \"num_correlators\": %d, "
                       "\"resident_call_ns\": %.0f, \"host_call_ns\": %.0f, \"session_call_ns\": %.0f, \"session_call_median_ns\": %.0f}\n",
                       N, M, L, best_res, best_host, best_session, med_session);
                fflush(stdout);
            }
    CHECK(gat_destroy(ctx));
    return 0;
}
