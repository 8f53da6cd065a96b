// Post-correlation array processing: what a multi-antenna receiver does with the [antennas x taps]
// accumulators of a channel right after the correlator -- Tracking.jl's `track(...; post_corr_filter)` hook
// [upstream; a user closure applied to every tap's antenna vector], here for accumulators that stay on the device.
//   beamform_kernel       y[l, k] = sum_m conj(w[m, k]) * acc[m, l, k]
//   eigen_weights_kernel  w[:, k] = dominant eigenvector of R_k, R_k <- forget * R_k + p p^H with p = acc[:, tap, k]
//                         (eigen-beamformer: the steering vector is estimated from the prompt covariance)
// Tiny and latency-bound: one thread per output (beamformer), one warp per channel (eigen filter); both run on
// the ctx stream right behind the correlate kernel.  PARITY UNPINNED: the reference ships no post-correlation
// filter of its own; the tests check these kernels against numpy (float64) restatements of the formulas above.
#include "gat_internal.h"
#include <algorithm>

namespace gat {

__global__ void beamform_kernel(const float *__restrict__ acc_re, const float *__restrict__ acc_im, const float *__restrict__ w_re,
                                const float *__restrict__ w_im, float *__restrict__ y_re, float *__restrict__ y_im, int n_ch, int n_taps,
                                int n_ants)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;     // (l, k), l fastest
    if (t >= n_ch * n_taps) return;
    const int k = t / n_taps;
    const float *ar = acc_re + (size_t)t * n_ants, *ai = acc_im + (size_t)t * n_ants;
    const float *wr = w_re + (size_t)k * n_ants, *wi = w_im + (size_t)k * n_ants;
    float sr = 0.f, si = 0.f;
    for (int m = 0; m < n_ants; ++m) {
        // conj(w) * a = (wr - j wi)(ar + j ai)
        sr = fmaf(wr[m], ar[m], fmaf(wi[m], ai[m], sr));
        si = fmaf(wr[m], ai[m], fmaf(-wi[m], ar[m], si));
    }
    y_re[t] = sr;
    y_im[t] = si;
}

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// one warp per channel; lane i owns row i of R (n_ants <= 32) and element i of w
__global__ void eigen_weights_kernel(const float *__restrict__ acc_re, const float *__restrict__ acc_im, int n_ch, int n_taps, int n_ants,
                                     int tap, float forget, int iters, float *__restrict__ cov_re, float *__restrict__ cov_im,
                                     float *__restrict__ w_re, float *__restrict__ w_im)
{
    const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (k >= n_ch) return;                                   // whole warps leave together
    const bool on = lane < n_ants;
    const size_t a_off = ((size_t)k * n_taps + tap) * n_ants;
    const float pr = on ? acc_re[a_off + lane] : 0.f, pi = on ? acc_im[a_off + lane] : 0.f;
    float *Rr = cov_re + ((size_t)k * n_ants + lane) * n_ants, *Ri = cov_im + ((size_t)k * n_ants + lane) * n_ants;
    // R[i][j] <- forget * R[i][j] + p_i conj(p_j)
    for (int j = 0; j < n_ants; ++j) {
        const float qr = __shfl_sync(0xffffffffu, pr, j), qi = __shfl_sync(0xffffffffu, pi, j);
        if (on) {
            Rr[j] = fmaf(forget, Rr[j], fmaf(pr, qr, pi * qi));
            Ri[j] = fmaf(forget, Ri[j], fmaf(pi, qr, -pr * qi));
        }
    }
    __syncwarp();
    float wr = on ? w_re[(size_t)k * n_ants + lane] : 0.f, wi = on ? w_im[(size_t)k * n_ants + lane] : 0.f;
    if (warp_sum(wr * wr + wi * wi) == 0.f) {                // cold start: the prompt vector itself
        wr = pr;
        wi = pi;
    }
    for (int it = 0; it < iters; ++it) {
        float vr = 0.f, vi = 0.f;
        for (int j = 0; j < n_ants; ++j) {
            const float xr = __shfl_sync(0xffffffffu, wr, j), xi = __shfl_sync(0xffffffffu, wi, j);
            if (on) {
                vr = fmaf(Rr[j], xr, fmaf(-Ri[j], xi, vr));
                vi = fmaf(Rr[j], xi, fmaf(Ri[j], xr, vi));
            }
        }
        const float n2 = warp_sum(vr * vr + vi * vi);
        const float inv = n2 > 0.f ? rsqrtf(n2) : 0.f;
        wr = vr * inv;
        wi = vi * inv;
    }
    // fix the free phase: antenna 0 real and non-negative
    const float r0 = __shfl_sync(0xffffffffu, wr, 0), i0 = __shfl_sync(0xffffffffu, wi, 0);
    const float m0 = sqrtf(r0 * r0 + i0 * i0);
    if (m0 > 0.f) {
        const float cr = r0 / m0, ci = -i0 / m0;             // multiply by conj(w0) / |w0|
        const float tr = wr * cr - wi * ci, ti = wr * ci + wi * cr;
        wr = tr;
        wi = ti;
    }
    if (on) {
        w_re[(size_t)k * n_ants + lane] = wr;
        w_im[(size_t)k * n_ants + lane] = wi;
    }
}

// out[i] = sum over the `n_slices` slices of a gather buffer, in slice order (fixed order: bit-reproducible) -- the
// cross-GPU reduction of sample-sharded correlation (every rank holds partial sums over its own sample range)
__global__ void sum_slices_kernel(const float *__restrict__ in_re, const float *__restrict__ in_im, size_t slice_stride, int n_slices,
                                  size_t n, float *__restrict__ out_re, float *__restrict__ out_im)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float ar = 0.f, ai = 0.f;
        for (int r = 0; r < n_slices; ++r) {
            ar += __ldcg(in_re + (size_t)r * slice_stride + i);     // written by peers' stores: bypass L1
            ai += __ldcg(in_im + (size_t)r * slice_stride + i);
        }
        out_re[i] = ar;
        out_im[i] = ai;
    }
}

cudaError_t launch_sum_slices(const float *in_re, const float *in_im, size_t slice_stride, int n_slices, size_t n, float *out_re,
                              float *out_im, cudaStream_t stream)
{
    const int threads = 256;
    const int blocks = (int)std::min<size_t>((n + threads - 1) / threads, 148 * 8);
    sum_slices_kernel<<<std::max(1, blocks), threads, 0, stream>>>(in_re, in_im, slice_stride, n_slices, n, out_re, out_im);
    return cudaGetLastError();
}

cudaError_t launch_beamform(const float *acc_re, const float *acc_im, const float *w_re, const float *w_im, float *y_re, float *y_im,
                            int n_ch, int n_taps, int n_ants, cudaStream_t stream)
{
    const int total = n_ch * n_taps;
    beamform_kernel<<<(total + 127) / 128, 128, 0, stream>>>(acc_re, acc_im, w_re, w_im, y_re, y_im, n_ch, n_taps, n_ants);
    return cudaGetLastError();
}

cudaError_t launch_eigen_weights(const float *acc_re, const float *acc_im, int n_ch, int n_taps, int n_ants, int tap, float forget, int iters,
                                 float *cov_re, float *cov_im, float *w_re, float *w_im, cudaStream_t stream)
{
    const int warps_per_block = 4;
    eigen_weights_kernel<<<(n_ch + warps_per_block - 1) / warps_per_block, 32 * warps_per_block, 0, stream>>>(
        acc_re, acc_im, n_ch, n_taps, n_ants, tap, forget, iters, cov_re, cov_im, w_re, w_im);
    return cudaGetLastError();
}

}  // namespace gat
