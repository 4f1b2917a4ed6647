// dmel_big_kernels.cuh — one-CTA-per-transform kernels for n_fft 8192 / 16384
// Part of libdmel_b200.so (single translation unit, see dmel_kernels.cu for the overview).
#pragma once
#include "dmel_common.cuh"

namespace dmel {

// ----------------------------------------------------------------------------
// big transforms (n_fft 8192 / 16384, e.g. the 300 ms window at 44.1 kHz): one CTA per tile, the whole
// CTA is one FFT group; the L-point spectrum buffer takes most of shared memory, so the window, the
// signal and the sparse bank are read from global/L2 and only the tile, the two 128-entry twiddle
// tables and (backward) the dx span live next to it
// ----------------------------------------------------------------------------
struct BigLay {
    int thi, tile, priv, red, total;   // 4-byte words; buf (2L words) is at 0
};
__host__ __device__ inline BigLay make_big_lay(int L, int span, int tile_words, bool priv_dx) {
    BigLay l;
    int o = 2 * L;
    l.thi = o;  o += 512;                     // t_hi[128] | t_lo[128] as float2
    l.tile = o; o += round_up4(tile_words);
    l.priv = o; o += priv_dx ? round_up4(span) : 0;
    l.red = o;  o += 64;
    l.total = o;
    return l;
}

template <int L>
__device__ __forceinline__ void big_load_pair(float2 (&v)[Plan<L>::E], const float* __restrict__ xb, const float* __restrict__ win_g,
                                              int pa0, int hop, int N, float mu, int tid) {
    using P0 = Pass<L, 0>;
#pragma unroll
    for (int i = 0; i < P0::NB; ++i) {
        const int j = tid + i * P0::GT;
#pragma unroll
        for (int r = 0; r < P0::R; ++r) {
            const int n = j + r * P0::STR;
            const float w = __ldg(win_g + n);
            const int pa = pa0 + n, pb = pa + hop;
            const float xa = (pa >= 0 && pa < N) ? __ldg(xb + pa) - mu : 0.f;
            const float xc = (pb >= 0 && pb < N) ? __ldg(xb + pb) - mu : 0.f;
            v[i * P0::R + r] = make_float2(xa * w, xc * w);
        }
    }
}

template <int L, bool SPEC>
__global__ void __launch_bounds__(L / Plan<L>::E)
dmel_big_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, Geom g, MelDev mel,
                    const float* __restrict__ win_g, const float2* __restrict__ tw_g, const float* __restrict__ means,
                    WsHeader* hdr) {
    constexpr int E = Plan<L>::E, NT = L / E;
    extern __shared__ __align__(16) float smem_f[];
    const int TS = g.TT + 1;
    const BigLay lay = make_big_lay(L, g.span, SPEC ? 0 : g.M * TS, false);
    float2* buf = reinterpret_cast<float2*>(smem_f);
    float2* t_hi = reinterpret_cast<float2*>(smem_f + lay.thi);
    float2* t_lo = t_hi + 128;
    float* tile = smem_f + lay.tile;
    const int tid = threadIdx.x;
    const int b = blockIdx.x / g.tiles_per_clip;
    const int tile_i = blockIdx.x - b * g.tiles_per_clip;
    const int t0 = tile_i * g.TT;
    const int nfr = min(g.TT, g.T - t0);
    const float* xb = x + (long long)b * g.x_stride;
    const float mu = __ldg(means + b);
    const bool do_log = g.log_eps >= 0.f;
    float ysum = 0.f;
    for (int i = tid; i < 256; i += NT) t_hi[i] = __ldg(tw_g + i);
    __syncthreads();

    for (int f = 0; f < nfr; f += 2) {
        float2 v[E];
        big_load_pair<L>(v, xb, win_g, (t0 + f) * g.hop - L / 2, g.hop, g.N, mu, tid);
        big_passes<L, 0, true>(v, buf, t_hi, t_lo, tid);
        for (int k = tid; k <= L / 2; k += NT) {
            const float2 z1 = buf[k];
            const float2 z2 = buf[(L - k) & (L - 1)];
            const float ar = 0.5f * (z1.x + z2.x), ai = 0.5f * (z1.y - z2.y);
            const float br = 0.5f * (z1.y + z2.y), bi = 0.5f * (z2.x - z1.x);
            float pa = fmaf(ar, ar, ai * ai), pb = fmaf(br, br, bi * bi);
            if (SPEC) {
                // (B,1,F,T): the two frames of the pair are neighbours in a row
                float* yr = y + ((long long)b * g.M + k) * g.T + t0 + f;
                if (do_log) { pa = __logf(pa + g.log_eps); pb = __logf(pb + g.log_eps); }
                yr[0] = pa;
                ysum += pa;
                if (f + 1 < nfr) { yr[1] = pb; ysum += pb; }
            } else {
                buf[k] = make_float2(pa, pb);
            }
        }
        if (!SPEC) {
            __syncthreads();
            for (int m = tid; m < g.M; m += NT) {
                const int k0 = __ldg(mel.mel_k0 + m);
                const int o0 = __ldg(mel.mel_off + m), o1 = __ldg(mel.mel_off + m + 1);
                const float2* pp = buf + k0;
                const float4* wp = reinterpret_cast<const float4*>(mel.mel_w + o0);
                float sa = 0.f, sb = 0.f;
                for (int q = 0; q < ((o1 - o0) >> 2); ++q) {
                    const float4 w = __ldg(wp + q);
                    const float2 a0 = pp[4 * q], a1 = pp[4 * q + 1], a2 = pp[4 * q + 2], a3 = pp[4 * q + 3];
                    sa = fmaf(w.x, a0.x, sa); sb = fmaf(w.x, a0.y, sb);
                    sa = fmaf(w.y, a1.x, sa); sb = fmaf(w.y, a1.y, sb);
                    sa = fmaf(w.z, a2.x, sa); sb = fmaf(w.z, a2.y, sb);
                    sa = fmaf(w.w, a3.x, sa); sb = fmaf(w.w, a3.y, sb);
                }
                tile[m * TS + f] = sa;
                tile[m * TS + f + 1] = sb;
            }
            // the barrier before the next scatter (inside big_passes) orders these reads of buf
        }
    }
    if (!SPEC) {
        __syncthreads();
        auto emit = [&](int m, int tt, long long off) {
            float val = tile[m * TS + tt];
            if (do_log) val = __logf(val + g.log_eps);
            if (mel.out_scale) val = affine_out(mel, m, val);
            y[off] = val;
            ysum += val;
        };
        if (!g.time_major) {
            const long long base = (long long)b * g.M * g.T + t0;
            const int tt_shift = ilog2(g.TT);
            for (int idx = tid; idx < (g.M << tt_shift); idx += NT) {
                const int m = idx >> tt_shift, tt = idx & (g.TT - 1);
                if (tt < nfr) emit(m, tt, base + (long long)m * g.T + tt);
            }
        } else {
            // (B,1,T,M): what MelPANNsNet transposes to (models.py:162); consecutive lanes write consecutive mel bins
            const long long base = ((long long)b * g.T + t0) * g.M;
            for (int tt = tid >> 5; tt < nfr; tt += NT / 32)
                for (int m = tid & 31; m < g.M; m += 32) emit(m, tt, base + (long long)tt * g.M + m);
        }
    }
    const double ytot = block_sum(double(ysum), reinterpret_cast<double*>(smem_f + lay.red));
    if (tid == 0) atomicAdd(&hdr->y_sum, ytot);
}

template <int L, bool SPEC, bool NEED_DX>
__global__ void __launch_bounds__(L / Plan<L>::E)
dmel_big_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ gy,
                    float* __restrict__ dx, Geom g, MelDev mel, const float* __restrict__ win_g,
                    const float2* __restrict__ tw_g, const float* __restrict__ means, WsHeader* hdr,
                    double* __restrict__ clipsum) {
    constexpr int E = Plan<L>::E, NT = L / E, NP = Plan<L>::NP;
    using P0 = Pass<L, 0>;
    using PL = Pass<L, NP - 1>;
    extern __shared__ __align__(16) float smem_f[];
    const int TS = g.TT + 1;
    // spectrogram mode reads dL/dy straight from global (its tile would be n_fft/2+1 rows)
    const BigLay lay = make_big_lay(L, g.span, SPEC ? 0 : g.M * TS, NEED_DX);
    float2* buf = reinterpret_cast<float2*>(smem_f);
    float2* t_hi = reinterpret_cast<float2*>(smem_f + lay.thi);
    float2* t_lo = t_hi + 128;
    float* gt = smem_f + lay.tile;
    float* priv = smem_f + lay.priv;
    double* red = reinterpret_cast<double*>(smem_f + lay.red);
    const int tid = threadIdx.x;
    const int b = blockIdx.x / g.tiles_per_clip;
    const int tile_i = blockIdx.x - b * g.tiles_per_clip;
    const int t0 = tile_i * g.TT;
    const int nfr = min(g.TT, g.T - t0);
    const int p0 = t0 * g.hop - L / 2;
    const float* xb = x + (long long)b * g.x_stride;
    const float mu = __ldg(means + b);
    const bool do_log = g.log_eps >= 0.f;
    const long long ybase = (long long)b * g.M * g.T + t0;
    const float inv_lam3 = hdr->inv_lam3, cprime = hdr->cprime, centre = hdr->centre;

    for (int i = tid; i < 256; i += NT) t_hi[i] = __ldg(tw_g + i);
    if (NEED_DX)
        for (int i = tid; i < g.span; i += NT) priv[i] = 0.f;
    if (!SPEC) {
        auto stage = [&](int m, int tt, long long off) {
            float val = 0.f;
            if (tt < nfr) {
                val = __ldg(gy + off);
                if (mel.out_scale) val *= affine_grad_factor(mel, m, do_log ? __ldg(y + off) : 0.f, do_log);
                else if (do_log) val *= __expf(-__ldg(y + off));
            }
            gt[m * TS + tt] = val;
        };
        if (!g.time_major) {
            const int tt_shift = ilog2(g.TT);
            for (int idx = tid; idx < (g.M << tt_shift); idx += NT) {
                const int m = idx >> tt_shift, tt = idx & (g.TT - 1);
                stage(m, tt, ybase + (long long)m * g.T + tt);
            }
        } else {
            const long long base = ((long long)b * g.T + t0) * g.M;
            for (int tt = tid >> 5; tt < g.TT; tt += NT / 32)
                for (int m = tid & 31; m < g.M; m += 32) stage(m, tt, base + (long long)tt * g.M + m);
        }
    }
    __syncthreads();

    float dlam = 0.f;
    for (int f = 0; f < nfr; f += 2) {
        const bool has_b = (f + 1) < nfr;
        const int pa0 = (t0 + f) * g.hop - L / 2;
        float2 v[E];
        big_load_pair<L>(v, xb, win_g, pa0, g.hop, g.N, mu, tid);
        big_passes<L, 0, true>(v, buf, t_hi, t_lo, tid);
        for (int k = tid; k <= L / 2; k += NT) {
            const float2 z1 = buf[k];
            const float2 z2 = buf[(L - k) & (L - 1)];
            const float ar = 0.5f * (z1.x + z2.x), ai = 0.5f * (z1.y - z2.y);
            const float br = 0.5f * (z1.y + z2.y), bi = 0.5f * (z2.x - z1.x);
            float ga = 0.f, gb = 0.f;
            if (SPEC) {
                const long long off = ybase + (long long)k * g.T + f;
                ga = __ldg(gy + off);
                if (do_log) ga *= __expf(-__ldg(y + off));
                if (has_b) {
                    gb = __ldg(gy + off + 1);
                    if (do_log) gb *= __expf(-__ldg(y + off + 1));
                }
            } else {
                const int m0 = __ldg(mel.bin_m0 + k);
                const int o0 = __ldg(mel.bin_off + k), o1 = __ldg(mel.bin_off + k + 1);
                for (int o = o0; o < o1; ++o) {
                    const float w = __ldg(mel.bin_w + o);
                    const float* gp = gt + (m0 + (o - o0)) * TS + f;
                    ga = fmaf(w, gp[0], ga);
                    gb = fmaf(w, gp[1], gb);
                }
            }
            if (k == 0 || k == L / 2) {
                buf[k] = make_float2(2.f * ga * ar, -2.f * gb * br);
            } else {
                buf[k] = make_float2(ga * ar - gb * bi, -(ga * ai + gb * br));
                buf[L - k] = make_float2(ga * ar + gb * bi, ga * ai - gb * br);
            }
        }
        __syncthreads();
        P0::gather(v, buf, tid);
        big_passes<L, 0, false>(v, buf, t_hi, t_lo, tid);
#pragma unroll
        for (int i = 0; i < PL::NB; ++i) {
#pragma unroll
            for (int r = 0; r < PL::R; ++r) {
                const int n = PL::out_index(tid, i, r);
                const float2 u = v[i * PL::R + r];
                const float w = __ldg(win_g + n);
                const float d = float(n) - centre;
                const float q = fmaf(d * d, inv_lam3, -cprime);
                const int pa = pa0 + n, pb = pa + g.hop;
                const float xa = (pa >= 0 && pa < g.N) ? __ldg(xb + pa) - mu : 0.f;
                const float xc = (has_b && pb >= 0 && pb < g.N) ? __ldg(xb + pb) - mu : 0.f;
                dlam = fmaf(w * q, u.x * xa - u.y * xc, dlam);
                v[i * PL::R + r] = make_float2(u.x * w, -u.y * w);
            }
        }
        if (NEED_DX) {
            float* da = priv + f * g.hop;
#pragma unroll
            for (int i = 0; i < PL::NB; ++i)
#pragma unroll
                for (int r = 0; r < PL::R; ++r) da[PL::out_index(tid, i, r)] += v[i * PL::R + r].x;
            __syncthreads();
            if (has_b) {
                float* db = da + g.hop;
#pragma unroll
                for (int i = 0; i < PL::NB; ++i)
#pragma unroll
                    for (int r = 0; r < PL::R; ++r) db[PL::out_index(tid, i, r)] += v[i * PL::R + r].y;
            }
            __syncthreads();
        }
    }
    const double tot = block_sum(double(dlam), red);
    if (tid == 0) atomicAdd(&hdr->dlam_acc, tot);
    if (NEED_DX) {
        const int span_used = (nfr - 1) * g.hop + L;
        const int excl_lo = (tile_i == 0) ? 0 : (t0 - 1) * g.hop + L / 2;
        const int excl_hi = (tile_i == g.tiles_per_clip - 1) ? g.N : (t0 + nfr) * g.hop - L / 2;
        float* dxb = dx + (long long)b * g.N;
        float part = 0.f;
        for (int i = tid; i < span_used; i += NT) {
            const int p = p0 + i;
            if (p >= 0 && p < g.N) {
                const float val = priv[i];
                part += val;
                if (p >= excl_lo && p < excl_hi) dxb[p] = val;
                else atomicAdd(dxb + p, val);
            }
        }
        const double cs = block_sum(double(part), red);
        if (tid == 0) atomicAdd(clipsum + b, cs);
    }
}

}  // namespace dmel
