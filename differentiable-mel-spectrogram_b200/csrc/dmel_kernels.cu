// dmel_kernels.cu — fused DMEL forward / backward kernels for sm_100a and their C ABI.
//
// Replaces, for one batch, the whole per-sample Python loop of the reference
//   models.py:33-56 (MelSpectrogramLayer.forward) / models.py:182-200 (SpectrogramLayer.forward)
//   -> time_frequency.py:21-58 (Gaussian window, torch.stft, |.|^2) -> melscale_fbanks matmul
//   -> torch.log(s + 1e-10) (models.py:72-73)  and autograd's backward of that chain.
//
// Kernels
//   dmel_prep_kernel      per-clip means (models.py:38), window w[n] (time_frequency.py:21-30,
//                         centre-padded to n_fft like torch.stft does), twiddle table, scalars
//   dmel_fwd_kernel<L>    one CTA = one tile of TT frames of one clip: the clip span is staged
//                         once in shared memory (mean removed, zero padded = center=True,
//                         pad_mode='constant'); frames are transformed two at a time as the
//                         real/imag parts of one L-point complex FFT; |X|^2, the sparse mel
//                         gather and the log happen on chip; only y is written
//   dmel_bwd_kernel<L>    recomputes the frames' spectra, forms Z = 2*GP*X, inverse-transforms
//                         two frames at a time, overlap-adds u*w into a shared-memory dx span and
//                         accumulates dL/dlambda = sum u*fr*dw/dlambda in registers
//   dmel_finalize_kernel  dx -= mean(dx) per clip (adjoint of the mean removal), fp64 -> fp32
//                         dL/dlambda with sign(lambda) (adjoint of torch.abs, models.py:38)
//   dmel_dft_*            direct-DFT versions of fwd/bwd for transform lengths without an FFT
//                         plan (non powers of two, < 128, > 4096); correctness path, O(L^2)
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <vector>

#include "../../include/dmel.h"
#include "fft_device.cuh"

namespace dmel {

// ----------------------------------------------------------------------------
// workspace layout (all offsets multiples of 16 bytes)
// ----------------------------------------------------------------------------
struct WsHeader {
    double dlam_acc;   // sum over all CTAs of dL/d|lambda|
    float inv_lam3;    // 1 / (|lambda| + 1e-15)^3
    float cprime;      // sum(w^2 q)/sum(w^2) when normalize_window, else 0
    float centre;      // window centre in n_fft coordinates: left + win_length/2
    float inv_norm;    // 1/||w|| when normalize_window, else 1
    int32_t left;      // (n_fft - win_length)/2
    int32_t pad_[9];
};
static_assert(sizeof(WsHeader) == 64, "header is 64 bytes");

struct WsView {
    WsHeader* hdr;
    float* win;       // [n_fft]
    float2* tw;       // [n_fft]  exp(-2*pi*i*k/n_fft)
    float* means;     // [B]
    double* clipsum;  // [B]
};

static inline size_t align16(size_t v) { return (v + 15) & ~size_t(15); }

static size_t ws_bytes(int64_t B, int32_t n_fft) {
    return sizeof(WsHeader) + align16(size_t(n_fft) * 4) + align16(size_t(n_fft) * 8) + align16(size_t(B) * 4) +
           align16(size_t(B) * 8);
}

static WsView ws_view(void* base, int64_t B, int32_t n_fft) {
    unsigned char* p = static_cast<unsigned char*>(base);
    WsView v;
    v.hdr = reinterpret_cast<WsHeader*>(p);
    p += sizeof(WsHeader);
    v.win = reinterpret_cast<float*>(p);
    p += align16(size_t(n_fft) * 4);
    v.tw = reinterpret_cast<float2*>(p);
    p += align16(size_t(n_fft) * 8);
    v.means = reinterpret_cast<float*>(p);
    p += align16(size_t(B) * 4);
    v.clipsum = reinterpret_cast<double*>(p);
    return v;
}

struct Geom {
    int B, N, T, hop, n_fft, M;  // M = output rows (n_mels, or n_fft/2+1 in spectrogram mode)
    int TT, tiles_per_clip, span;
    long long x_stride;
    float log_eps;
};

struct MelDev {
    const int* mel_k0;
    const int* mel_off;
    const float* mel_w;
    const int* bin_m0;
    const int* bin_off;
    const float* bin_w;
};

// ----------------------------------------------------------------------------
// block reductions
// ----------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// result valid in thread 0; `red` is >= 32 doubles of shared memory
__device__ __forceinline__ double block_sum(double v, double* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    double r = 0.0;
    if (wid == 0) {
        r = lane < nw ? red[lane] : 0.0;
        r = warp_sum(r);
    }
    return r;
}

// ----------------------------------------------------------------------------
// prep: means + window + twiddles
// ----------------------------------------------------------------------------
constexpr int PREP_THREADS = 256;

__global__ void __launch_bounds__(PREP_THREADS)
dmel_prep_kernel(const float* __restrict__ x, long long x_stride, int B, int N, float lam_abs, int win_length,
                 int n_fft, int flags, WsView ws) {
    __shared__ double red[32];
    const int tid = threadIdx.x;
    const int normalize = flags & DMEL_FLAG_NORMALIZE_WINDOW;
    if (int(blockIdx.x) < B) {
        if (flags & DMEL_FLAG_KEEP_MEAN) {
            if (tid == 0) ws.means[blockIdx.x] = 0.f;
            return;
        }
        // models.py:38  x[idx] - torch.mean(x[idx])
        const float* xb = x + (long long)blockIdx.x * x_stride;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        int i = tid;
        for (; i + 3 * PREP_THREADS < N; i += 4 * PREP_THREADS) {
            s0 += __ldg(xb + i);
            s1 += __ldg(xb + i + PREP_THREADS);
            s2 += __ldg(xb + i + 2 * PREP_THREADS);
            s3 += __ldg(xb + i + 3 * PREP_THREADS);
        }
        for (; i < N; i += PREP_THREADS) s0 += __ldg(xb + i);
        double tot = block_sum(double(s0) + double(s1) + double(s2) + double(s3), red);
        if (tid == 0) ws.means[blockIdx.x] = float(tot / double(N));
        return;
    }
    // time_frequency.py:21-30; torch.stft centre-pads the window to n_fft
    const int left = (n_fft - win_length) / 2;
    const float lam = lam_abs + 1e-15f;
    const float c0 = 0.5f * float(win_length);
    const float inv_lam3 = 1.0f / (lam * lam * lam);
    double s2 = 0.0, sq = 0.0;
    for (int n = tid; n < n_fft; n += PREP_THREADS) {
        const int m = n - left;
        float w = 0.f;
        if (m >= 0 && m < win_length) {
            const float d = float(m) - c0;
            const float t = d / lam;
            w = expf(-0.5f * (t * t));
            s2 += double(w) * double(w);
            sq += double(w) * double(w) * double(d * d * inv_lam3);
        }
        ws.win[n] = w;
        double sn, cs;
        sincospi(-2.0 * double(n) / double(n_fft), &sn, &cs);
        ws.tw[n] = make_float2(float(cs), float(sn));
    }
    const double S2 = block_sum(s2, red);
    __shared__ float sh_scale;
    __syncthreads();
    const double SQ = block_sum(sq, red);
    if (tid == 0) {
        // S2 is valid in thread 0 only (block_sum contract)
        float inv_norm = 1.f, cprime = 0.f;
        if (normalize) {
            inv_norm = 1.0f / sqrtf(float(S2));
            cprime = float(SQ / S2);
        }
        sh_scale = inv_norm;
        ws.hdr->inv_lam3 = inv_lam3;
        ws.hdr->cprime = cprime;
        ws.hdr->centre = float(left) + c0;
        ws.hdr->inv_norm = inv_norm;
        ws.hdr->left = left;
    }
    __syncthreads();
    if (normalize) {
        const float sc = sh_scale;
        for (int n = tid; n < n_fft; n += PREP_THREADS) ws.win[n] *= sc;
    }
}

// ----------------------------------------------------------------------------
// shared-memory carve-up shared by host and device
// ----------------------------------------------------------------------------
__host__ __device__ inline int round_up4(int v) { return (v + 3) & ~3; }

// bytes: [buf0 L*8][buf1 L*8][win L*4][xs span4*4][dxs span4*4 (bwd+dx)][tile M*(TT+1)*4][red 256]
static size_t smem_bytes_fft(int L, int span, int M, int TT, bool need_dxs) {
    size_t s = size_t(L) * 16 + size_t(L) * 4;
    s += size_t(round_up4(span)) * 4 * (need_dxs ? 2 : 1);
    s += size_t(round_up4(M * (TT + 1))) * 4;
    s += 256;
    return s;
}

// ----------------------------------------------------------------------------
// forward
// ----------------------------------------------------------------------------
template <int L, bool SPEC>
__global__ void __launch_bounds__(L / Plan<L>::E)
dmel_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, Geom g, MelDev mel,
                const float* __restrict__ win_g, const float2* __restrict__ tw, const float* __restrict__ means) {
    constexpr int E = Plan<L>::E, NT = L / E;
    using P0 = Pass<L, 0>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* buf0 = reinterpret_cast<float2*>(smem_raw);
    float2* buf1 = buf0 + L;
    float* win = reinterpret_cast<float*>(buf1 + L);
    float* xs = win + L;
    float* tile = xs + round_up4(g.span);

    const int tid = threadIdx.x;
    const int b = blockIdx.x / g.tiles_per_clip;
    const int tile_i = blockIdx.x - b * g.tiles_per_clip;
    const int t0 = tile_i * g.TT;
    const int nfr = min(g.TT, g.T - t0);
    const int TS = g.TT + 1;

    for (int i = tid; i < L; i += NT) win[i] = __ldg(win_g + i);
    {
        const float mu = __ldg(means + b);
        const float* xb = x + (long long)b * g.x_stride;
        const int p0 = t0 * g.hop - L / 2;
        for (int i = tid; i < g.span; i += NT) {
            const int p = p0 + i;
            xs[i] = (p >= 0 && p < g.N) ? __ldg(xb + p) - mu : 0.f;
        }
    }
    __syncthreads();

    int sel = 0;
    for (int f = 0; f < nfr; f += 2) {
        const float* xa = xs + f * g.hop;
        const float* xb2 = xa + g.hop;
        float2 v[E];
#pragma unroll
        for (int i = 0; i < P0::NB; ++i) {
            const int j = tid + i * NT;
#pragma unroll
            for (int r = 0; r < P0::R; ++r) {
                const int n = j + r * P0::STR;
                const float w = win[n];
                v[i * P0::R + r] = make_float2(xa[n] * w, xb2[n] * w);
            }
        }
        run_passes<L, 0, true>(v, buf0, buf1, sel, tw, tid);
        float2* Z = sel ? buf0 : buf1;  // buffer written last
        // packed pair -> two one-sided power spectra (time_frequency.py:53)
        for (int k = tid; k <= L / 2; k += NT) {
            const float2 z1 = Z[k];
            const float2 z2 = Z[(L - k) & (L - 1)];
            const float ar = 0.5f * (z1.x + z2.x), ai = 0.5f * (z1.y - z2.y);
            const float br = 0.5f * (z1.y + z2.y), bi = 0.5f * (z2.x - z1.x);
            const float pa = fmaf(ar, ar, ai * ai), pb = fmaf(br, br, bi * bi);
            if (SPEC) {
                tile[k * TS + f] = pa;
                tile[k * TS + f + 1] = pb;
            } else {
                Z[k] = make_float2(pa, pb);
            }
        }
        if (!SPEC) {
            __syncthreads();
            // models.py:53 matmul(spec^T, mel_fb): each mel reads its contiguous run of bins
            for (int m = tid; m < g.M; m += NT) {
                const int k0 = __ldg(mel.mel_k0 + m);
                const int o0 = __ldg(mel.mel_off + m), o1 = __ldg(mel.mel_off + m + 1);
                float sa = 0.f, sb = 0.f;
                for (int o = o0; o < o1; ++o) {
                    const float w = __ldg(mel.mel_w + o);
                    const float2 p = Z[k0 + (o - o0)];
                    sa = fmaf(w, p.x, sa);
                    sb = fmaf(w, p.y, sb);
                }
                tile[m * TS + f] = sa;
                tile[m * TS + f + 1] = sb;
            }
        }
    }
    __syncthreads();
    // models.py:54 pack into (B,1,M,T), T fastest; optional fused log (models.py:72-73)
    float* yb = y + (long long)b * g.M * g.T + t0;
    const bool do_log = g.log_eps >= 0.f;
    for (int idx = tid; idx < g.M * nfr; idx += NT) {
        const int m = idx / nfr, tt = idx - m * nfr;
        float val = tile[m * TS + tt];
        if (do_log) val = logf(val + g.log_eps);
        yb[(long long)m * g.T + tt] = val;
    }
}

// ----------------------------------------------------------------------------
// backward
// ----------------------------------------------------------------------------
template <int L, bool SPEC, bool NEED_DX>
__global__ void __launch_bounds__(L / Plan<L>::E)
dmel_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ gy,
                float* __restrict__ dx, Geom g, MelDev mel, const float* __restrict__ win_g,
                const float2* __restrict__ tw, const float* __restrict__ means, WsHeader* hdr,
                double* __restrict__ clipsum) {
    constexpr int E = Plan<L>::E, NT = L / E, NP = Plan<L>::NP;
    using P0 = Pass<L, 0>;
    using PL = Pass<L, NP - 1>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* buf0 = reinterpret_cast<float2*>(smem_raw);
    float2* buf1 = buf0 + L;
    float* win = reinterpret_cast<float*>(buf1 + L);
    float* xs = win + L;
    float* dxs = xs + round_up4(g.span);
    float* gt = NEED_DX ? dxs + round_up4(g.span) : dxs;
    double* red = reinterpret_cast<double*>(gt + round_up4(g.M * (g.TT + 1)));

    const int tid = threadIdx.x;
    const int b = blockIdx.x / g.tiles_per_clip;
    const int tile_i = blockIdx.x - b * g.tiles_per_clip;
    const int t0 = tile_i * g.TT;
    const int nfr = min(g.TT, g.T - t0);
    const int TS = g.TT + 1;
    const int p0 = t0 * g.hop - L / 2;

    for (int i = tid; i < L; i += NT) win[i] = __ldg(win_g + i);
    {
        const float mu = __ldg(means + b);
        const float* xb = x + (long long)b * g.x_stride;
        for (int i = tid; i < g.span; i += NT) {
            const int p = p0 + i;
            xs[i] = (p >= 0 && p < g.N) ? __ldg(xb + p) - mu : 0.f;
            if (NEED_DX) dxs[i] = 0.f;
        }
    }
    {
        // upstream gradient tile; with the fused log dL/dmel = G / (mel + eps) = G * exp(-y)
        const bool do_log = g.log_eps >= 0.f;
        const long long base = (long long)b * g.M * g.T + t0;
        for (int idx = tid; idx < g.M * g.TT; idx += NT) {
            const int m = idx / g.TT, tt = idx - m * g.TT;
            float val = 0.f;
            if (tt < nfr) {
                const long long off = base + (long long)m * g.T + tt;
                val = __ldg(gy + off);
                if (do_log) val *= expf(-__ldg(y + off));
            }
            gt[m * TS + tt] = val;
        }
    }
    __syncthreads();

    const float inv_lam3 = hdr->inv_lam3, cprime = hdr->cprime, centre = hdr->centre;
    float dlam = 0.f;
    int sel = 0;
    for (int f = 0; f < nfr; f += 2) {
        const float* xa = xs + f * g.hop;
        const float* xb2 = xa + g.hop;
        const bool has_b = (f + 1) < nfr;
        float2 v[E];
#pragma unroll
        for (int i = 0; i < P0::NB; ++i) {
            const int j = tid + i * NT;
#pragma unroll
            for (int r = 0; r < P0::R; ++r) {
                const int n = j + r * P0::STR;
                const float w = win[n];
                v[i * P0::R + r] = make_float2(xa[n] * w, xb2[n] * w);
            }
        }
        run_passes<L, 0, true>(v, buf0, buf1, sel, tw, tid);
        float2* Z = sel ? buf0 : buf1;
        // Z <- conj(V),  V = Zhat_a + i*Zhat_b the Hermitian-extended 2*GP*X of both frames
        for (int k = tid; k <= L / 2; k += NT) {
            const float2 z1 = Z[k];
            const float2 z2 = Z[(L - k) & (L - 1)];
            const float ar = 0.5f * (z1.x + z2.x), ai = 0.5f * (z1.y - z2.y);
            const float br = 0.5f * (z1.y + z2.y), bi = 0.5f * (z2.x - z1.x);
            float ga = 0.f, gb = 0.f;
            if (SPEC) {
                ga = gt[k * TS + f];
                gb = gt[k * TS + f + 1];
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
                Z[k] = make_float2(2.f * ga * ar, -2.f * gb * br);
            } else {
                Z[k] = make_float2(ga * ar - gb * bi, -(ga * ai + gb * br));
                Z[L - k] = make_float2(ga * ar + gb * bi, ga * ai - gb * br);
            }
        }
        __syncthreads();
        // inverse transform as conj(FFT(conj(V))): u_a = Re, u_b = -Im of the FFT output
        P0::gather(v, Z, tid);
        run_passes<L, 0, false>(v, buf0, buf1, sel, tw, tid);
        // dL/dlambda, time-domain form: sum_n u[n] * fr[n] * dw/dlambda[n],
        //   dw/dlambda[n] = w[n] * ((n-centre)^2/lambda^3 - cprime)
#pragma unroll
        for (int i = 0; i < PL::NB; ++i) {
#pragma unroll
            for (int r = 0; r < PL::R; ++r) {
                const int n = PL::out_index(tid, i, r);
                const float2 u = v[i * PL::R + r];
                const float w = win[n];
                const float d = float(n) - centre;
                const float q = fmaf(d * d, inv_lam3, -cprime);
                const float s = u.x * xa[n] - (has_b ? u.y * xb2[n] : 0.f);
                dlam = fmaf(w * q, s, dlam);
                v[i * PL::R + r] = make_float2(u.x * w, -u.y * w);
            }
        }
        if (NEED_DX) {
            // overlap-add (adjoint of the framing): frame a, barrier, frame b, barrier
            float* da = dxs + f * g.hop;
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
        // crop the centre padding; positions no other tile touches are stored, halos are reduced
        const int span_used = (nfr - 1) * g.hop + L;
        const int excl_lo = (tile_i == 0) ? 0 : (t0 - 1) * g.hop + L / 2;
        const int excl_hi = (tile_i == g.tiles_per_clip - 1) ? g.N : (t0 + nfr) * g.hop - L / 2;
        float* dxb = dx + (long long)b * g.N;
        float part = 0.f;
        for (int i = tid; i < span_used; i += NT) {
            const int p = p0 + i;
            if (p >= 0 && p < g.N) {
                const float val = dxs[i];
                part += val;
                if (p >= excl_lo && p < excl_hi) dxb[p] = val;
                else atomicAdd(dxb + p, val);
            }
        }
        const double cs = block_sum(double(part), red);
        if (tid == 0) atomicAdd(clipsum + b, cs);
    }
}

// ----------------------------------------------------------------------------
// finalize: dx -= mean(dx) per clip; dL/dlambda -> fp32 with sign(lambda)
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dmel_finalize_kernel(float* __restrict__ dx, int B, int N, int chunks_per_clip, const double* __restrict__ clipsum,
                     const WsHeader* hdr, float sign, float* __restrict__ grad_lambda) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *grad_lambda = float(double(sign) * hdr->dlam_acc);
    if (dx == nullptr) return;
    const int b = blockIdx.x / chunks_per_clip;
    const int c = blockIdx.x - b * chunks_per_clip;
    if (b >= B) return;
    if (clipsum == nullptr) return;  // DMEL_FLAG_KEEP_MEAN: no mean-removal adjoint
    const float mu = float(clipsum[b] / double(N));
    float* dxb = dx + (long long)b * N;
    const int per = (N + chunks_per_clip - 1) / chunks_per_clip;
    const int lo = c * per, hi = min(N, lo + per);
    for (int i = lo + threadIdx.x; i < hi; i += 256) dxb[i] -= mu;
}

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
        float* yb = y_out + (long long)b * g.M * g.T + t;
        for (int m = tid; m < g.M; m += DFT_THREADS) {
            float s;
            if (spec_mode) {
                s = gp[m];
            } else {
                const int k0 = __ldg(mel.mel_k0 + m);
                const int o0 = __ldg(mel.mel_off + m), o1 = __ldg(mel.mel_off + m + 1);
                s = 0.f;
                for (int o = o0; o < o1; ++o) s = fmaf(__ldg(mel.mel_w + o), gp[k0 + (o - o0)], s);
            }
            if (do_log) s = logf(s + g.log_eps);
            yb[(long long)m * g.T] = s;
        }
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
                if (do_log) gv *= expf(-__ldg(y_in + off));
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

// ----------------------------------------------------------------------------
// host side: planning + dispatch
// ----------------------------------------------------------------------------
constexpr size_t SMEM_MAX = 227 * 1024;

static bool has_fft_plan(int n_fft) {
    switch (n_fft) {
        case 128: case 256: case 512: case 1024: case 2048: case 4096: return true;
        default: return false;
    }
}

static int fft_threads(int n_fft) {
    switch (n_fft) {
        case 128: return 128 / Plan<128>::E;
        case 256: return 256 / Plan<256>::E;
        case 512: return 512 / Plan<512>::E;
        case 1024: return 1024 / Plan<1024>::E;
        case 2048: return 2048 / Plan<2048>::E;
        case 4096: return 4096 / Plan<4096>::E;
        default: return 0;
    }
}

struct LaunchPlan {
    int path;  // 0 FFT, 1 direct DFT
    int TT, span, tiles_per_clip, threads;
    size_t smem;
};

// Tile size: the largest even TT in {32,16,8,4,2} whose shared memory leaves room for
// several CTAs per SM; relax the occupancy target until something fits.
static int plan_launch(int64_t N, int n_fft, int hop, int rows, bool bwd, bool need_dx, LaunchPlan* lp) {
    const int64_t T64 = N / hop + 1;
    if (T64 > INT32_MAX) return DMEL_ERR_BAD_GEOMETRY;
    const int T = int(T64);
    if (!has_fft_plan(n_fft)) {
        if (n_fft < 2 || n_fft > DMEL_MAX_DFT_NFFT || (n_fft & 1)) return DMEL_ERR_UNSUPPORTED_NFFT;
        const int F = n_fft / 2 + 1;
        lp->path = 1;
        lp->TT = 1;
        lp->span = n_fft;
        lp->tiles_per_clip = T;
        lp->threads = DFT_THREADS;
        lp->smem = size_t(round_up4(n_fft)) * 4 + size_t(F) * 8 + size_t(round_up4(F) + 8) * 4 + 512;
        if (lp->smem > SMEM_MAX) return DMEL_ERR_SMEM_EXCEEDED;
        return DMEL_OK;
    }
    const size_t targets[4] = {SMEM_MAX / 4, SMEM_MAX / 3, SMEM_MAX / 2, SMEM_MAX};
    const int tts[5] = {32, 16, 8, 4, 2};
    const int t_even = (T + 1) & ~1;
    for (int ti = 0; ti < 4; ++ti) {
        for (int c = 0; c < 5; ++c) {
            int TT = tts[c];
            if (TT > t_even) TT = t_even;
            const int64_t span64 = int64_t(TT - 1) * hop + n_fft;
            if (span64 > (1 << 20)) continue;
            const size_t s = smem_bytes_fft(n_fft, int(span64), rows, TT, bwd && need_dx);
            if (s <= targets[ti]) {
                lp->path = 0;
                lp->TT = TT;
                lp->span = int(span64);
                lp->tiles_per_clip = (T + TT - 1) / TT;
                lp->threads = fft_threads(n_fft);
                lp->smem = s;
                return DMEL_OK;
            }
        }
    }
    return DMEL_ERR_SMEM_EXCEEDED;
}

// ---- optional instrumentation for bench.py: launch counter + per-kernel CUDA-event timing.
// Off by default; when on, every main-kernel launch is bracketed by a pair of events recorded on
// the launching stream (no synchronisation until dmel_profile_read).
enum { PROF_PREP = 0, PROF_FWD = 1, PROF_BWD = 2, PROF_FINALIZE = 3, PROF_KINDS = 4 };
static std::atomic<long long> g_launches[PROF_KINDS];
struct ProfState {
    bool on = false;
    std::vector<cudaEvent_t> ev[PROF_KINDS];  // start/stop pairs
};
static thread_local ProfState t_prof;

struct ProfScope {
    int kind;
    cudaStream_t st;
    ProfScope(int k, cudaStream_t s) : kind(k), st(s) {
        g_launches[kind].fetch_add(1, std::memory_order_relaxed);
        if (t_prof.on) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            cudaEventRecord(e, st);
            t_prof.ev[kind].push_back(e);
        }
    }
    ~ProfScope() {
        if (t_prof.on) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            cudaEventRecord(e, st);
            t_prof.ev[kind].push_back(e);
        }
    }
};

static int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        fprintf(stderr, "[dmel] %s: %s\n", what, cudaGetErrorString(e));
        return DMEL_ERR_LAUNCH_FAILED;
    }
    return DMEL_OK;
}

template <typename K>
static int set_smem(K kernel, size_t smem) {
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (e != cudaSuccess) {
            fprintf(stderr, "[dmel] cudaFuncSetAttribute(%zu): %s\n", smem, cudaGetErrorString(e));
            (void)cudaGetLastError();
            return DMEL_ERR_LAUNCH_FAILED;
        }
    }
    return DMEL_OK;
}

template <int L>
static int launch_fwd(const float* x, float* y, const Geom& g, const MelDev& mel, bool spec, const WsView& ws,
                      const LaunchPlan& lp, cudaStream_t st) {
    const unsigned grid = unsigned(int64_t(g.B) * g.tiles_per_clip);
    int rc;
    ProfScope prof(PROF_FWD, st);
    if (spec) {
        if ((rc = set_smem(dmel_fwd_kernel<L, true>, lp.smem))) return rc;
        dmel_fwd_kernel<L, true><<<grid, lp.threads, lp.smem, st>>>(x, y, g, mel, ws.win, ws.tw, ws.means);
    } else {
        if ((rc = set_smem(dmel_fwd_kernel<L, false>, lp.smem))) return rc;
        dmel_fwd_kernel<L, false><<<grid, lp.threads, lp.smem, st>>>(x, y, g, mel, ws.win, ws.tw, ws.means);
    }
    return check_launch("dmel_fwd_kernel");
}

template <int L, bool SPEC, bool NEED_DX>
static int launch_bwd_t(const float* x, const float* y, const float* gy, float* dx, const Geom& g, const MelDev& mel,
                        const WsView& ws, const LaunchPlan& lp, cudaStream_t st) {
    const unsigned grid = unsigned(int64_t(g.B) * g.tiles_per_clip);
    int rc;
    if ((rc = set_smem(dmel_bwd_kernel<L, SPEC, NEED_DX>, lp.smem))) return rc;
    ProfScope prof(PROF_BWD, st);
    dmel_bwd_kernel<L, SPEC, NEED_DX>
        <<<grid, lp.threads, lp.smem, st>>>(x, y, gy, dx, g, mel, ws.win, ws.tw, ws.means, ws.hdr, ws.clipsum);
    return check_launch("dmel_bwd_kernel");
}

template <int L>
static int launch_bwd(const float* x, const float* y, const float* gy, float* dx, const Geom& g, const MelDev& mel,
                      bool spec, const WsView& ws, const LaunchPlan& lp, cudaStream_t st) {
    if (spec) {
        return dx ? launch_bwd_t<L, true, true>(x, y, gy, dx, g, mel, ws, lp, st)
                  : launch_bwd_t<L, true, false>(x, y, gy, dx, g, mel, ws, lp, st);
    }
    return dx ? launch_bwd_t<L, false, true>(x, y, gy, dx, g, mel, ws, lp, st)
              : launch_bwd_t<L, false, false>(x, y, gy, dx, g, mel, ws, lp, st);
}

#define DMEL_DISPATCH_L(n_fft, CALL)              \
    switch (n_fft) {                              \
        case 128: rc = CALL(128); break;          \
        case 256: rc = CALL(256); break;          \
        case 512: rc = CALL(512); break;          \
        case 1024: rc = CALL(1024); break;        \
        case 2048: rc = CALL(2048); break;        \
        case 4096: rc = CALL(4096); break;        \
        default: rc = DMEL_ERR_UNSUPPORTED_NFFT;  \
    }

static int validate(const float* x, int64_t B, int64_t N, int64_t stride, int32_t win, int32_t n_fft, int32_t hop,
                    int32_t n_mels, const dmel_melbank* bank, void* ws, size_t ws_bytes_in) {
    if (!x || !ws) return DMEL_ERR_NULL_POINTER;
    if (B <= 0 || N <= 0 || hop <= 0 || win <= 0 || n_fft < win || n_mels < 0 || stride < N) return DMEL_ERR_BAD_GEOMETRY;
    if (B > INT32_MAX / 2 || N > INT32_MAX / 2) return DMEL_ERR_BAD_GEOMETRY;
    if (n_mels > 0) {
        if (!bank) return DMEL_ERR_NULL_POINTER;
        if (bank->n_mels != n_mels || bank->n_freqs != n_fft / 2 + 1) return DMEL_ERR_BAD_MELBANK;
        if (!bank->mel_k0 || !bank->mel_off || !bank->mel_w || !bank->bin_m0 || !bank->bin_off || !bank->bin_w)
            return DMEL_ERR_NULL_POINTER;
    }
    if (ws_bytes_in < ws_bytes(B, n_fft)) return DMEL_ERR_WORKSPACE_TOO_SMALL;
    return DMEL_OK;
}

static MelDev mel_dev(const dmel_melbank* bank) {
    MelDev m{};
    if (bank) {
        m.mel_k0 = bank->mel_k0; m.mel_off = bank->mel_off; m.mel_w = bank->mel_w;
        m.bin_m0 = bank->bin_m0; m.bin_off = bank->bin_off; m.bin_w = bank->bin_w;
    }
    return m;
}

static int run_prep(const float* x, int64_t B, int64_t N, int64_t stride, float lam_abs, int32_t win, int32_t n_fft,
                    int32_t flags, const WsView& ws, cudaStream_t st) {
    ProfScope prof(PROF_PREP, st);
    dmel_prep_kernel<<<unsigned(B + 1), PREP_THREADS, 0, st>>>(x, stride, int(B), int(N), lam_abs, win, n_fft,
                                                                flags, ws);
    return check_launch("dmel_prep_kernel");
}

}  // namespace dmel

// ----------------------------------------------------------------------------
// C ABI
// ----------------------------------------------------------------------------
using namespace dmel;

extern "C" DMEL_API int dmel_abi_version(void) { return DMEL_ABI_VERSION; }

extern "C" DMEL_API const char* dmel_status_string(int status) {
    switch (status) {
        case DMEL_OK: return "ok";
        case DMEL_ERR_NULL_POINTER: return "null pointer argument";
        case DMEL_ERR_BAD_GEOMETRY: return "inconsistent geometry (batch, n_points, hop, win_length, n_fft, n_mels)";
        case DMEL_ERR_UNSUPPORTED_NFFT: return "no kernel for this n_fft";
        case DMEL_ERR_WORKSPACE_TOO_SMALL: return "workspace smaller than dmel_workspace_bytes()";
        case DMEL_ERR_LAUNCH_FAILED: return "CUDA launch failed (see stderr)";
        case DMEL_ERR_SMEM_EXCEEDED: return "no tile size fits in shared memory";
        case DMEL_ERR_BAD_MELBANK: return "melbank does not match n_fft / n_mels";
        default: return "unknown dmel status";
    }
}

extern "C" DMEL_API int dmel_nfft_supported(int32_t n_fft) {
    if (has_fft_plan(n_fft)) return 1;
    return (n_fft >= 2 && n_fft <= DMEL_MAX_DFT_NFFT && !(n_fft & 1)) ? 1 : 0;
}

extern "C" DMEL_API int32_t dmel_window_length_from_lambda(float lambda) {
    // time_frequency.py:39,60-65: int((|lambd| * 6) as fp32), then 1 << (x-1).bit_length()
    const float prod = fabsf(lambda) * 6.0f;
    if (!(prod < 1073741824.0f)) return 0;
    const int32_t xi = int32_t(prod);  // truncation toward zero like int()
    const int64_t v = int64_t(xi) - 1;
    // Python's (-1).bit_length() == 1, (0).bit_length() == 0
    int bits = 0;
    uint64_t a = uint64_t(v < 0 ? -v : v);
    while (a) { ++bits; a >>= 1; }
    return int32_t(1) << bits;
}

extern "C" DMEL_API size_t dmel_workspace_bytes(int64_t batch, int32_t n_fft) {
    if (batch <= 0 || n_fft <= 0) return 0;
    return ws_bytes(batch, n_fft);
}

extern "C" DMEL_API int dmel_query_launch(int64_t n_points, int32_t n_fft, int32_t hop, int32_t rows, int32_t which,
                                 int32_t* tile_frames, int32_t* smem_bytes, int32_t* threads, int32_t* path) {
    if (n_points <= 0 || hop <= 0 || n_fft <= 0 || rows <= 0) return DMEL_ERR_BAD_GEOMETRY;
    LaunchPlan lp{};
    int rc = plan_launch(n_points, n_fft, hop, rows, which != 0, which == 1, &lp);
    if (rc) return rc;
    if (tile_frames) *tile_frames = lp.TT;
    if (smem_bytes) *smem_bytes = int32_t(lp.smem);
    if (threads) *threads = lp.threads;
    if (path) *path = lp.path;
    return DMEL_OK;
}

static Geom make_geom(int64_t B, int64_t N, int64_t stride, int32_t n_fft, int32_t hop, int32_t rows, float log_eps,
                      const LaunchPlan& lp) {
    Geom g{};
    g.B = int(B);
    g.N = int(N);
    g.T = int(N / hop + 1);
    g.hop = hop;
    g.n_fft = n_fft;
    g.M = rows;
    g.TT = lp.TT;
    g.tiles_per_clip = lp.tiles_per_clip;
    g.span = lp.span;
    g.x_stride = stride;
    g.log_eps = log_eps;
    return g;
}

extern "C" DMEL_API int dmel_forward_f32(const float* x, int64_t batch, int64_t n_points, int64_t x_row_stride, float lambda,
                                int32_t win_length, int32_t n_fft, int32_t hop, int32_t n_mels,
                                const dmel_melbank* bank, float log_eps, int32_t flags, float* y,
                                void* workspace, size_t workspace_bytes, void* cuda_stream) {
    if (!y) return DMEL_ERR_NULL_POINTER;
    int rc = validate(x, batch, n_points, x_row_stride, win_length, n_fft, hop, n_mels, bank, workspace, workspace_bytes);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    const bool spec = (n_mels == 0);
    const int rows = spec ? n_fft / 2 + 1 : n_mels;
    LaunchPlan lp{};
    if ((rc = plan_launch(n_points, n_fft, hop, rows, false, false, &lp))) return rc;
    if (int64_t(batch) * lp.tiles_per_clip > INT32_MAX) return DMEL_ERR_BAD_GEOMETRY;
    WsView ws = ws_view(workspace, batch, n_fft);
    if ((rc = run_prep(x, batch, n_points, x_row_stride, fabsf(lambda), win_length, n_fft, flags, ws, st)))
        return rc;
    const Geom g = make_geom(batch, n_points, x_row_stride, n_fft, hop, rows, log_eps, lp);
    const MelDev mel = mel_dev(bank);
    if (lp.path == 1) {
        if ((rc = set_smem(dmel_dft_kernel<false, false>, lp.smem))) return rc;
        ProfScope prof(PROF_FWD, st);
        dmel_dft_kernel<false, false><<<unsigned(int64_t(g.B) * g.T), DFT_THREADS, lp.smem, st>>>(
            x, nullptr, nullptr, y, nullptr, g, mel, spec ? 1 : 0, ws.win, ws.tw, ws.means, ws.hdr, ws.clipsum);
        return check_launch("dmel_dft_kernel<fwd>");
    }
#define CALL_FWD(LL) launch_fwd<LL>(x, y, g, mel, spec, ws, lp, st)
    DMEL_DISPATCH_L(n_fft, CALL_FWD)
#undef CALL_FWD
    return rc;
}

extern "C" DMEL_API int dmel_backward_f32(const float* x, int64_t batch, int64_t n_points, int64_t x_row_stride, float lambda,
                                 int32_t win_length, int32_t n_fft, int32_t hop, int32_t n_mels,
                                 const dmel_melbank* bank, float log_eps, int32_t flags, const float* y,
                                 const float* grad_y, float* grad_x, float* grad_lambda, void* workspace,
                                 size_t workspace_bytes, int32_t prepared, void* cuda_stream) {
    if (!grad_y || !grad_lambda) return DMEL_ERR_NULL_POINTER;
    if (log_eps >= 0.f && !y) return DMEL_ERR_NULL_POINTER;
    int rc = validate(x, batch, n_points, x_row_stride, win_length, n_fft, hop, n_mels, bank, workspace, workspace_bytes);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    const bool spec = (n_mels == 0);
    const int rows = spec ? n_fft / 2 + 1 : n_mels;
    const bool need_dx = grad_x != nullptr;
    LaunchPlan lp{};
    if ((rc = plan_launch(n_points, n_fft, hop, rows, true, need_dx, &lp))) return rc;
    if (int64_t(batch) * lp.tiles_per_clip > INT32_MAX) return DMEL_ERR_BAD_GEOMETRY;
    WsView ws = ws_view(workspace, batch, n_fft);
    if (!prepared) {
        if ((rc = run_prep(x, batch, n_points, x_row_stride, fabsf(lambda), win_length, n_fft, flags, ws, st)))
            return rc;
    }
    // zero the accumulators (and dx: halo positions are reduced with atomics)
    if (cudaMemsetAsync(&ws.hdr->dlam_acc, 0, sizeof(double), st) != cudaSuccess) return check_launch("memset dlam");
    if (need_dx) {
        if (cudaMemsetAsync(ws.clipsum, 0, size_t(batch) * sizeof(double), st) != cudaSuccess)
            return check_launch("memset clipsum");
        if (cudaMemsetAsync(grad_x, 0, size_t(batch) * size_t(n_points) * sizeof(float), st) != cudaSuccess)
            return check_launch("memset grad_x");
    }
    const Geom g = make_geom(batch, n_points, x_row_stride, n_fft, hop, rows, log_eps, lp);
    const MelDev mel = mel_dev(bank);
    if (lp.path == 1) {
        const unsigned grid = unsigned(int64_t(g.B) * g.T);
        ProfScope prof(PROF_BWD, st);
        if (need_dx) {
            if ((rc = set_smem(dmel_dft_kernel<true, true>, lp.smem))) return rc;
            dmel_dft_kernel<true, true><<<grid, DFT_THREADS, lp.smem, st>>>(
                x, y, grad_y, nullptr, grad_x, g, mel, spec ? 1 : 0, ws.win, ws.tw, ws.means, ws.hdr, ws.clipsum);
        } else {
            if ((rc = set_smem(dmel_dft_kernel<true, false>, lp.smem))) return rc;
            dmel_dft_kernel<true, false><<<grid, DFT_THREADS, lp.smem, st>>>(
                x, y, grad_y, nullptr, nullptr, g, mel, spec ? 1 : 0, ws.win, ws.tw, ws.means, ws.hdr, ws.clipsum);
        }
        if ((rc = check_launch("dmel_dft_kernel<bwd>"))) return rc;
    } else {
#define CALL_BWD(LL) launch_bwd<LL>(x, y, grad_y, grad_x, g, mel, spec, ws, lp, st)
        DMEL_DISPATCH_L(n_fft, CALL_BWD)
#undef CALL_BWD
        if (rc) return rc;
    }
    const float sign = lambda > 0.f ? 1.f : (lambda < 0.f ? -1.f : 0.f);
    int chunks = 1, grid = 1;
    if (need_dx) {
        chunks = int((n_points + 8191) / 8192);
        grid = int(batch) * chunks;
    }
    const double* csum = (flags & DMEL_FLAG_KEEP_MEAN) ? nullptr : ws.clipsum;
    if (!csum) grid = 1;
    ProfScope prof(PROF_FINALIZE, st);
    dmel_finalize_kernel<<<unsigned(grid), 256, 0, st>>>(grad_x, int(batch), int(n_points), chunks, csum, ws.hdr,
                                                         sign, grad_lambda);
    return check_launch("dmel_finalize_kernel");
}

extern "C" DMEL_API long long dmel_launch_count(int32_t kind) {
    if (kind < 0) {
        long long t = 0;
        for (int k = 0; k < PROF_KINDS; ++k) t += g_launches[k].load(std::memory_order_relaxed);
        return t;
    }
    return kind < PROF_KINDS ? g_launches[kind].load(std::memory_order_relaxed) : 0;
}

extern "C" DMEL_API void dmel_profile_enable(int32_t on) {
    for (int k = 0; k < PROF_KINDS; ++k) {
        for (cudaEvent_t e : t_prof.ev[k]) cudaEventDestroy(e);
        t_prof.ev[k].clear();
    }
    t_prof.on = on != 0;
}

extern "C" DMEL_API int dmel_profile_read(int32_t kind, double* total_ms, long long* launches) {
    if (kind < 0 || kind >= PROF_KINDS || !total_ms || !launches) return DMEL_ERR_BAD_GEOMETRY;
    double tot = 0.0;
    const std::vector<cudaEvent_t>& v = t_prof.ev[kind];
    for (size_t i = 0; i + 1 < v.size(); i += 2) {
        if (cudaEventSynchronize(v[i + 1]) != cudaSuccess) return DMEL_ERR_LAUNCH_FAILED;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, v[i], v[i + 1]) != cudaSuccess) return DMEL_ERR_LAUNCH_FAILED;
        tot += ms;
    }
    *total_ms = tot;
    *launches = (long long)(v.size() / 2);
    return DMEL_OK;
}
