// dmel_dft_kernels.cuh — direct-DFT correctness path for transform lengths without an FFT plan
// Part of libdmel_b200.so (single translation unit, see dmel_kernels.cu for the overview).
#pragma once
#include "dmel_common.cuh"

namespace dmel {

// ----------------------------------------------------------------------------
// direct-DFT path (any n_fft): one CTA per (clip, frame)
// ----------------------------------------------------------------------------
constexpr int DFT_THREADS = 256;

// smem: frw[n_fft] windowed frame, spec[F] (float2: power / Z), out rows staging
template <bool BWD, bool NEED_DX>
__global__ void __launch_bounds__(DFT_THREADS)
dmel_dft_kernel(const float* __restrict__ x, const float* __restrict__ y_in, const float* __restrict__ gy,
                float* __restrict__ y_out, float* __restrict__ dx, Geom g, MelDev mel, int spec_mode,
                const float* __restrict__ win_g, const float2* __restrict__ tw, const float* __restrict__ means,
                WsHeader* hdr, double* __restrict__ clipsum) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int L = g.n_fft, F = L / 2 + 1;
    float* fr = reinterpret_cast<float*>(smem_raw);           // [L] centred frame samples
    float2* X = reinterpret_cast<float2*>(fr + round_up4(L)); // [F]
    float* gp = reinterpret_cast<float*>(X + F);              // [F] power (fwd only)
    double* red = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(gp + F) + 15) & ~uintptr_t(15));
    const int tid = threadIdx.x;
    const int b = blockIdx.x / g.T, t = blockIdx.x - b * g.T;
    const float mu = __ldg(means + b);
    const float* xb = x + (long long)b * g.x_stride;
    const int p0 = t * g.hop - L / 2;
    for (int n = tid; n < L; n += DFT_THREADS) {
        const int p = p0 + n;
        fr[n] = (p >= 0 && p < g.N) ? __ldg(xb + p) - mu : 0.f;
    }
    __syncthreads();
    for (int k = tid; k < F; k += DFT_THREADS) {
        float re = 0.f, im = 0.f;
        int idx = 0;
        for (int n = 0; n < L; ++n) {
            const float a = fr[n] * __ldg(win_g + n);
            const float2 w = __ldg(tw + idx);
            re = fmaf(a, w.x, re);
            im = fmaf(a, w.y, im);
            idx += k;
            if (idx >= L) idx -= L;
        }
        X[k] = make_float2(re, im);
        if (!BWD) gp[k] = fmaf(re, re, im * im);
    }
    __syncthreads();
    const bool do_log = g.log_eps >= 0.f;
    if (!BWD) {
        float ysum_dft = 0.f;
        float* yb = y_out + (long long)b * g.M * g.T + t;
        for (int m = tid; m < g.M; m += DFT_THREADS) {
            float s;
            if (spec_mode) {
                s = gp[m];
            } else {
                const int k0 = __ldg(mel.mel_k0 + m);
                const int o0 = __ldg(mel.mel_off + m), o1 = __ldg(mel.mel_off + m + 1);
                s = 0.f;
                // runs are zero-padded to 4 taps (dmel.h): the padded taps may point past the last bin, where gp holds
                // no spectrum (0 * leftover NaN bits would poison the sum), so the walk stops at bin F-1
                const int cnt = min(o1 - o0, F - k0);
                for (int o = 0; o < cnt; ++o) s = fmaf(__ldg(mel.mel_w + o0 + o), gp[k0 + o], s);
            }
            if (do_log) s = logf(s + g.log_eps);
            if (!spec_mode && mel.out_scale) s = affine_out(mel, m, s);
            yb[(long long)m * g.T] = s;
            ysum_dft += s;
        }
        const double ytot = block_sum(double(ysum_dft), red);
        if (tid == 0) atomicAdd(&hdr->y_sum, ytot);
        return;
    }
    // backward: GP[k], Z[k] = c_k * 2 GP X  (c_k = 1/2 off the DC/Nyquist bins, whose X is real)
    const long long ybase = (long long)b * g.M * g.T + t;
    for (int k = tid; k < F; k += DFT_THREADS) {
        float ga = 0.f;
        if (spec_mode) {
            ga = __ldg(gy + ybase + (long long)k * g.T);
            if (do_log) ga *= expf(-__ldg(y_in + ybase + (long long)k * g.T));
        } else {
            const int m0 = __ldg(mel.bin_m0 + k);
            const int o0 = __ldg(mel.bin_off + k), o1 = __ldg(mel.bin_off + k + 1);
            for (int o = o0; o < o1; ++o) {
                const long long off = ybase + (long long)(m0 + (o - o0)) * g.T;
                float gv = __ldg(gy + off);
                if (mel.out_scale) gv *= affine_grad_factor(mel, m0 + (o - o0), do_log ? __ldg(y_in + off) : 0.f, do_log);
                else if (do_log) gv *= expf(-__ldg(y_in + off));
                ga = fmaf(__ldg(mel.bin_w + o), gv, ga);
            }
        }
        const float2 xk = X[k];
        const bool edge = (k == 0) || (2 * k == L);
        // u[n] = Re sum_k Zs[k] e^{+i theta k n}: store Zs = 2*GP*X (edge: real part only)
        X[k] = edge ? make_float2(2.f * ga * xk.x, 0.f) : make_float2(2.f * ga * xk.x, 2.f * ga * xk.y);
    }
    __syncthreads();
    const float inv_lam3 = hdr->inv_lam3, cprime = hdr->cprime, centre = hdr->centre;
    float dlam = 0.f, part = 0.f;
    float* dxb = NEED_DX ? dx + (long long)b * g.N : nullptr;
    for (int n = tid; n < L; n += DFT_THREADS) {
        float u = 0.f;
        int idx = 0;
        for (int k = 0; k < F; ++k) {
            const float2 z = X[k];
            const float2 w = __ldg(tw + idx);  // e^{-i theta}; e^{+i theta} = conj
            u += z.x * w.x + z.y * w.y;        // Re(z * conj(w))
            idx += n;
            if (idx >= L) idx -= L;
        }
        const float w = __ldg(win_g + n);
        const float d = float(n) - centre;
        const float q = fmaf(d * d, inv_lam3, -cprime);
        dlam = fmaf(w * q, u * fr[n], dlam);
        if (NEED_DX) {
            const int p = p0 + n;
            if (p >= 0 && p < g.N) {
                const float val = u * w;
                part += val;
                atomicAdd(dxb + p, val);
            }
        }
    }
    const double tot = block_sum(double(dlam), red);
    if (tid == 0) atomicAdd(&hdr->dlam_acc, tot);
    if (NEED_DX) {
        const double cs = block_sum(double(part), red);
        if (tid == 0) atomicAdd(clipsum + b, cs);
    }
}

}  // namespace dmel
