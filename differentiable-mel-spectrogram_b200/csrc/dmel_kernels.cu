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
//                         pad_mode='constant') next to the window, the twiddle tables and the
//                         sparse filterbank.  NG independent FFT groups (one warp each for
//                         L <= 512) split the tile into contiguous sub-tiles and transform their
//                         frames two at a time as the real/imag parts of one L-point complex FFT;
//                         |X|^2, the sparse mel gather and the log happen on chip; only y is written
//   dmel_bwd_kernel<L>    same decomposition: recomputes the spectra, forms Z = 2*GP*X, inverse-
//                         transforms two frames at a time, overlap-adds u*w into the group's private
//                         shared-memory dx sub-span (merged per tile at the end) and accumulates
//                         dL/dlambda = sum u*fr*dw/dlambda in registers
//   dmel_finalize_kernel  dx -= mean(dx) per clip (adjoint of the mean removal), fp64 -> fp32
//                         dL/dlambda with sign(lambda) (adjoint of torch.abs, models.py:38)
//   dmel_dft_*            direct-DFT versions of fwd/bwd for transform lengths without an FFT
//                         plan (non powers of two, < 128, > 4096); correctness path, O(L^2)
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include <stddef.h>
#include <stdlib.h>

#include <algorithm>
#include <atomic>
#include <cmath>
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
    int32_t pad0_;
    double y_sum;      // sum of every output element of the last forward (train.py:58's running energy)
    int32_t pad_[6];
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
    int TT;                      // frames per CTA tile (= NG * SF)
    int SF;                      // frames per group sub-tile (even)
    int tiles_per_clip, span;    // span = (TT-1)*hop + n_fft samples staged per tile
    int sub_span;                // (SF-1)*hop + n_fft: private dx span of one group
    int csr_ints, csr_floats;    // size of the sparse-bank form the kernel stages (0 in spectrogram mode)
    long long x_stride;
    float log_eps;
    int row_aligned;             // x and x_stride are 16-byte aligned: spans may move as 16-byte quads
    int lookahead;               // 1: span (and dL/dy, y) buffers are doubled and tile i+1 is prefetched during tile i
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

// host mirror of fft_device.cuh's Plan<L> (np == 0: no FFT plan, direct-DFT path)
struct PlanDesc {
    int np;
    int r[4];
};

__global__ void __launch_bounds__(PREP_THREADS)
dmel_prep_kernel(const float* __restrict__ x, long long x_stride, int B, int N, float lam_abs, int win_length,
                 int n_fft, int flags, WsView ws, PlanDesc plan) {
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
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        int i0 = 0;
        if ((reinterpret_cast<uintptr_t>(xb) & 15) == 0) {
            // 16-byte loads, 8 independent ones in flight per thread
            const float4* x4 = reinterpret_cast<const float4*>(xb);
            const int n4 = N >> 2;
            int q = tid;
            for (; q + 7 * PREP_THREADS < n4; q += 8 * PREP_THREADS) {
                float4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = __ldg(x4 + q + u * PREP_THREADS);
#pragma unroll
                for (int u = 0; u < 8; ++u) acc[u] += (v[u].x + v[u].y) + (v[u].z + v[u].w);
            }
            for (; q < n4; q += PREP_THREADS) {
                const float4 v = __ldg(x4 + q);
                acc[0] += (v.x + v.y) + (v.z + v.w);
            }
            i0 = n4 << 2;
        }
        for (int i = i0 + tid; i < N; i += PREP_THREADS) acc[1] += __ldg(xb + i);
        double part = 0.0;
#pragma unroll
        for (int u = 0; u < 8; ++u) part += double(acc[u]);
        double tot = block_sum(part, red);
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
        if (plan.np < 0 && n < 256) {  // big-FFT path: t_hi[h] = exp(-2*pi*i*128h/n_fft), t_lo[l] = exp(-2*pi*i*l/n_fft)
            double sn, cs;
            const double idx = n < 128 ? 128.0 * double(n) : double(n - 128);
            sincospi(-2.0 * idx / double(n_fft), &sn, &cs);
            ws.tw[n] = make_float2(float(cs), float(sn));
        }
        if (plan.np == 0) {  // direct-DFT path: plain table exp(-2*pi*i*n/n_fft)
            double sn, cs;
            sincospi(-2.0 * double(n) / double(n_fft), &sn, &cs);
            ws.tw[n] = make_float2(float(cs), float(sn));
        }
    }
    // FFT path: per-pass tables T_p[r-1][k] = exp(-2*pi*i*k*r/(NS_p*R_p)) (fft_device.cuh, TwOffset)
    {
        int off = 0, ns = plan.r[0];
        for (int p = 1; p < plan.np; ++p) {
            const int R = plan.r[p], cnt = (R - 1) * ns;
            for (int e = tid; e < cnt; e += PREP_THREADS) {
                const int r = e / ns + 1, k = e - (r - 1) * ns;
                double sn, cs;
                sincospi(-2.0 * double(k) * double(r) / (double(ns) * double(R)), &sn, &cs);
                ws.tw[off + e] = make_float2(float(cs), float(sn));
            }
            off += cnt;
            ns *= R;
        }
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
        ws.hdr->y_sum = 0.0;
    }
    __syncthreads();
    if (normalize) {
        const float sc = sh_scale;
        for (int n = tid; n < n_fft; n += PREP_THREADS) ws.win[n] *= sc;
    }
}

// ----------------------------------------------------------------------------
// shared-memory carve-up shared by host and device (offsets in 4-byte words, all multiples of 4)
// ----------------------------------------------------------------------------
__host__ __device__ inline int round_up4(int v) { return (v + 3) & ~3; }

struct SmemLay {
    int twt, win, xs, tile, csr_i, csr_w, red, grp, grp_stride, total;
};

__host__ __device__ inline SmemLay make_lay(int L, int NG, int tw_total, int span, int sub_span, int tile_words,
                                            int csr_ints, int csr_floats, bool priv_dx, int span_bufs) {
    SmemLay l;
    int o = 0;
    l.twt = o;   o += round_up4(2 * tw_total);
    l.win = o;   o += L;
    l.xs = o;    o += span_bufs * round_up4(span + 4);   // (+ up to 3 words of 16-byte phase shift each)
    l.tile = o;  o += round_up4(tile_words);
    l.csr_i = o; o += round_up4(csr_ints);
    l.csr_w = o; o += round_up4(csr_floats);
    l.red = o;   o += 64;
    l.grp = o;
    l.grp_stride = 2 * L + (priv_dx ? round_up4(sub_span) : 0);
    o += NG * l.grp_stride;
    l.total = o;
    return l;
}

// ----------------------------------------------------------------------------
// asynchronous staging helpers (LDGSTS: global -> shared without a register round trip)
// ----------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
// 4-byte copy; valid == false writes a zero instead (src-size 0 => zero fill, nothing is read)
__device__ __forceinline__ void cp_async4_zfill(void* smem_dst, const void* gmem_src, bool valid) {
    const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    const int n = valid ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(gmem_src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// copy `words` 4-byte words (multiple of 4; both pointers 16-byte aligned) with the whole CTA
template <int NT>
__device__ __forceinline__ void stage_table_async(float* dst, const void* src, int words, int tid) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int i = tid; i < (words >> 2); i += NT) cp_async16(d4 + i, s4 + i);
}

// Stage the signal span [p0, p0+span) of one clip: out-of-range samples are zero (center=True,
// pad_mode='constant').  Issue phase only; call span_centre() after cp_async_wait_all().
// The span is stored at xs + shift (shift = p0 mod 4 in [0,3]) so that shared and global
// addresses share their 16-byte phase and whole in-range quads move as one 16-byte copy.
template <int NT>
__device__ __forceinline__ void span_issue(float* xs_sh, const float* __restrict__ xb, int p0, int span, int N,
                                           bool row_aligned, int tid) {
    // quad q covers span indices [4q - shift, 4q - shift + 3]  <->  xs_sh storage words [4q, 4q+3] (before +shift)
    const int shift = p0 & 3;                       // p0 may be negative: & 3 is still the mod-4 phase
    float* base = xs_sh;                            // storage word i + shift holds span index i
    const int nq = (span + shift + 3) >> 2;
    for (int q = tid; q < nq; q += NT) {
        const int i0 = 4 * q - shift;               // span index of the quad's first word
        const int p = p0 + i0;                      // global sample index, multiple of 4 by construction
        float* d = base + 4 * q;
        if (row_aligned && i0 >= 0 && i0 + 3 < span && p >= 0 && p + 3 < N) {
            cp_async16(d, xb + p);
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int pe = p + e, ie = i0 + e;
                const bool ok = (ie >= 0) && (ie < span) && (pe >= 0) && (pe < N);
                cp_async4_zfill(d + e, xb + (ok ? pe : 0), ok);
            }
        }
    }
}

// remove the clip mean from the in-range samples this thread staged (same q loop as span_issue,
// so every thread only touches words whose copies it has itself waited for)
template <int NT>
__device__ __forceinline__ void span_centre(float* xs_sh, int p0, int span, int N, float mu, int tid) {
    const int shift = p0 & 3;
    const int nq = (span + shift + 3) >> 2;
    for (int q = tid; q < nq; q += NT) {
        const int i0 = 4 * q - shift;
        const int p = p0 + i0;
        float4* d = reinterpret_cast<float4*>(xs_sh + 4 * q);
        float4 v = *d;
        v.x = (p + 0 >= 0 && p + 0 < N) ? v.x - mu : 0.f;
        v.y = (p + 1 >= 0 && p + 1 < N) ? v.y - mu : 0.f;
        v.z = (p + 2 >= 0 && p + 2 < N) ? v.z - mu : 0.f;
        v.w = (p + 3 >= 0 && p + 3 < N) ? v.w - mu : 0.f;
        *d = v;
    }
}

// Register budget: the one-warp-per-transform kernels (L <= 512, 128-thread CTAs) are latency bound, so
// they are compiled for DMEL_MINB_SMALL resident CTAs per SM (5 -> 96 registers, 20 warps/SM; measured
// best of 1/5/6 on B200, gpurun_out/exp/minb.log).
#ifndef DMEL_MINB_SMALL
#define DMEL_MINB_SMALL 5
#endif
// Everything else keeps the 128-register budget (512 threads' worth of CTAs per SM).
// The dx-producing backward is limited to 4 CTAs per SM by shared memory anyway and runs faster with the full
// 128 registers (r01e 66 us vs 71 us at 96), so only the forward and the lambda-only backward take the cap.
__host__ __device__ constexpr int min_ctas(int L, int NG, int NT, bool full_bwd = false) {
    return (L <= 512 && NG == 4 && !full_bwd) ? DMEL_MINB_SMALL : (512 / NT > 0 ? 512 / NT : 1);
}

// integer log2 of a power of two
__device__ __forceinline__ int ilog2(int v) { return 31 - __clz(v); }

// ----------------------------------------------------------------------------
// forward (persistent: each CTA stages the constant tables once and walks tiles)
// ----------------------------------------------------------------------------
template <int L, int NG, bool SPEC>
__global__ void __launch_bounds__(NG * (L / Plan<L>::E), min_ctas(L, NG, NG * (L / Plan<L>::E)))
dmel_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, Geom g, MelDev mel,
                const float* __restrict__ win_g, const float2* __restrict__ tw_g, const float* __restrict__ means,
                WsHeader* hdr) {
    constexpr int E = Plan<L>::E, GT = L / E, NT = NG * GT, TWN = TwTotal<L>::value;
    using P0 = Pass<L, 0>;
    extern __shared__ __align__(16) float smem_f[];
    const int TS = g.TT + 1;
    const SmemLay lay = make_lay(L, NG, TWN, g.span, g.sub_span, g.M * TS, g.csr_ints, g.csr_floats, false, 1 + g.lookahead);
    float2* twt = reinterpret_cast<float2*>(smem_f + lay.twt);
    float* win = smem_f + lay.win;
    float* xs_sh = smem_f + lay.xs;
    float* tile = smem_f + lay.tile;
    int* csr_i = reinterpret_cast<int*>(smem_f + lay.csr_i);
    float* csr_w = smem_f + lay.csr_w;

    const int tid = threadIdx.x;
    const int grp = tid / GT, gtid = tid - grp * GT, bar_id = 1 + grp;
    float2* buf = reinterpret_cast<float2*>(smem_f + lay.grp + grp * lay.grp_stride);
    const bool row_aligned = g.row_aligned != 0;
    const int tt_shift = ilog2(g.TT);
    const bool do_log = g.log_eps >= 0.f;

    // constant tables: once per CTA
    stage_table_async<NT>(smem_f + lay.twt, tw_g, 2 * TWN, tid);
    stage_table_async<NT>(win, win_g, L, tid);
    if (!SPEC) {
        // mel-major sparse bank: ints = k0[M4] | off[..], floats = weights (sub-arrays padded to 4 words)
        stage_table_async<NT>(reinterpret_cast<float*>(csr_i), mel.mel_k0, round_up4(g.M), tid);
        stage_table_async<NT>(reinterpret_cast<float*>(csr_i + round_up4(g.M)), mel.mel_off, round_up4(g.M + 1), tid);
        stage_table_async<NT>(csr_w, mel.mel_w, round_up4(g.csr_floats), tid);
    }
    const int* s_k0 = csr_i;
    const int* s_off = csr_i + round_up4(g.M);

    // Persistent tile loop with a one-tile look-ahead: while tile i is transformed, the span of tile i+1 is
    // already landing in the other span buffer (cp.async), so the staging latency is off the critical path.
    const int total_tiles = g.B * g.tiles_per_clip;
    const int xs_words = round_up4(g.span + 4);
    const bool ahead = g.lookahead != 0;
    float ysum = 0.f;
    if (ahead && int(blockIdx.x) < total_tiles) {
        const int b0 = blockIdx.x / g.tiles_per_clip;
        const int p00 = (blockIdx.x - b0 * g.tiles_per_clip) * g.TT * g.hop - L / 2;
        span_issue<NT>(xs_sh, x + (long long)b0 * g.x_stride, p00, g.span, g.N, row_aligned, tid);
    }
    int it = 0;
    for (int tix = blockIdx.x; tix < total_tiles; tix += gridDim.x, ++it) {
        const int b = tix / g.tiles_per_clip;
        const int tile_i = tix - b * g.tiles_per_clip;
        const int t0 = tile_i * g.TT;
        const int nfr = min(g.TT, g.T - t0);
        const int p0 = t0 * g.hop - L / 2;
        float* xs_cur = xs_sh + (ahead ? (it & 1) * xs_words : 0);
        const float* xs = xs_cur + (p0 & 3);

        if (!ahead) span_issue<NT>(xs_cur, x + (long long)b * g.x_stride, p0, g.span, g.N, row_aligned, tid);
        const float mu = __ldg(means + b);
        cp_async_wait_all();
        span_centre<NT>(xs_cur, p0, g.span, g.N, mu, tid);
        __syncthreads();   // also: every warp is done with the previous tile (its span buffer, the output tile)
        if (ahead) {
            const int nx = tix + gridDim.x;
            if (nx < total_tiles) {
                const int bn = nx / g.tiles_per_clip;
                const int pn = (nx - bn * g.tiles_per_clip) * g.TT * g.hop - L / 2;
                span_issue<NT>(xs_sh + ((it + 1) & 1) * xs_words, x + (long long)bn * g.x_stride, pn, g.span, g.N,
                               row_aligned, tid);
            }
        }

        const int f_hi = min(grp * g.SF + g.SF, nfr);
        for (int f = grp * g.SF; f < f_hi; f += 2) {
            const float* xa = xs + f * g.hop;
            const float* xb2 = xa + g.hop;
            float2 v[E];
#pragma unroll
            for (int i = 0; i < P0::NB; ++i) {
                const int j = gtid + i * GT;
#pragma unroll
                for (int r = 0; r < P0::R; ++r) {
                    const int n = j + r * P0::STR;
                    const float w = win[n];
                    v[i * P0::R + r] = make_float2(xa[n] * w, xb2[n] * w);
                }
            }
            run_passes<L, 0, true>(v, buf, twt, gtid, bar_id);
            // packed pair -> two one-sided power spectra (time_frequency.py:53), in place
#pragma unroll 4
            for (int k = gtid; k <= L / 2; k += GT) {
                const float2 z1 = buf[k];
                const float2 z2 = buf[(L - k) & (L - 1)];
                const float ar = 0.5f * (z1.x + z2.x), ai = 0.5f * (z1.y - z2.y);
                const float br = 0.5f * (z1.y + z2.y), bi = 0.5f * (z2.x - z1.x);
                const float pa = fmaf(ar, ar, ai * ai), pb = fmaf(br, br, bi * bi);
                if (SPEC) {
                    tile[k * TS + f] = pa;
                    tile[k * TS + f + 1] = pb;
                } else {
                    buf[k] = make_float2(pa, pb);
                }
            }
            if (!SPEC) {
                group_sync<GT>(bar_id);
                // models.py:53 matmul(spec^T, mel_fb): each mel reads its contiguous run of bins.  Runs are
                // zero-padded to a multiple of 4 taps by the host (dmel.h), so 4 taps go per step with one
                // 16-byte weight load; the bins behind a padded tap are valid shared memory (k0+run <= L/2+3 < L)
                for (int m = gtid; m < g.M; m += GT) {
                    const int k0 = s_k0[m];
                    const int o0 = s_off[m], o1 = s_off[m + 1];
                    const float2* pp = buf + k0;
                    const float4* wp = reinterpret_cast<const float4*>(csr_w + o0);
                    float sa = 0.f, sb = 0.f;
                    for (int q = 0; q < ((o1 - o0) >> 2); ++q) {
                        const float4 w = wp[q];
                        const float2 a0 = pp[4 * q], a1 = pp[4 * q + 1], a2 = pp[4 * q + 2], a3 = pp[4 * q + 3];
                        sa = fmaf(w.x, a0.x, sa); sb = fmaf(w.x, a0.y, sb);
                        sa = fmaf(w.y, a1.x, sa); sb = fmaf(w.y, a1.y, sb);
                        sa = fmaf(w.z, a2.x, sa); sb = fmaf(w.z, a2.y, sb);
                        sa = fmaf(w.w, a3.x, sa); sb = fmaf(w.w, a3.y, sb);
                    }
                    tile[m * TS + f] = sa;
                    tile[m * TS + f + 1] = sb;
                }
            }
        }
        __syncthreads();
        // models.py:54 pack into (B,1,M,T), T fastest; optional fused log (models.py:72-73)
        float* yb = y + (long long)b * g.M * g.T + t0;
        for (int idx = tid; idx < (g.M << tt_shift); idx += NT) {
            const int m = idx >> tt_shift, tt = idx & (g.TT - 1);
            if (tt < nfr) {
                float val = tile[m * TS + tt];
                // lg2.approx * ln2: <= 3 ulp, i.e. ~1e-6 absolute on log-mel values (budget 1e-5 + 1e-4 |y|)
                if (do_log) val = __logf(val + g.log_eps);
                yb[(long long)m * g.T + tt] = val;
                ysum += val;
            }
        }
        // the next iteration's staging only writes xs; its __syncthreads orders it after this epilogue
    }
    cp_async_wait_all();
    // fused by-product: sum of all outputs (what train.py:58 copies the whole batch to the host for)
    const double ytot = block_sum(double(ysum), reinterpret_cast<double*>(smem_f + lay.red));
    if (tid == 0) atomicAdd(&hdr->y_sum, ytot);
}

// ----------------------------------------------------------------------------
// backward (persistent)
// ----------------------------------------------------------------------------
template <int L, int NG, bool SPEC, bool NEED_DX>
__global__ void __launch_bounds__(NG * (L / Plan<L>::E), min_ctas(L, NG, NG * (L / Plan<L>::E), NEED_DX))
dmel_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ gy,
                float* __restrict__ dx, Geom g, MelDev mel, const float* __restrict__ win_g,
                const float2* __restrict__ tw_g, const float* __restrict__ means, WsHeader* hdr,
                double* __restrict__ clipsum) {
    constexpr int E = Plan<L>::E, GT = L / E, NT = NG * GT, NP = Plan<L>::NP, TWN = TwTotal<L>::value;
    constexpr int F = L / 2 + 1;
    using P0 = Pass<L, 0>;
    using PL = Pass<L, NP - 1>;
    extern __shared__ __align__(16) float smem_f[];
    const int TS = g.TT + 1;
    const SmemLay lay = make_lay(L, NG, TWN, g.span, g.sub_span, (2 + 2 * g.lookahead) * round_up4(g.M * TS), g.csr_ints,
                                 g.csr_floats, NEED_DX, 1 + g.lookahead);
    float2* twt = reinterpret_cast<float2*>(smem_f + lay.twt);
    float* win = smem_f + lay.win;
    float* xs_sh = smem_f + lay.xs;
    float* gt = smem_f + lay.tile;
    int* csr_i = reinterpret_cast<int*>(smem_f + lay.csr_i);
    float* csr_w = smem_f + lay.csr_w;
    double* red = reinterpret_cast<double*>(smem_f + lay.red);

    const int tid = threadIdx.x;
    const int grp = tid / GT, gtid = tid - grp * GT, bar_id = 1 + grp;
    float2* buf = reinterpret_cast<float2*>(smem_f + lay.grp + grp * lay.grp_stride);
    float* priv = smem_f + lay.grp + grp * lay.grp_stride + 2 * L;   // this group's dx sub-span
    const bool row_aligned = g.row_aligned != 0;
    const int tt_shift = ilog2(g.TT);
    const bool do_log = g.log_eps >= 0.f;

    stage_table_async<NT>(smem_f + lay.twt, tw_g, 2 * TWN, tid);
    stage_table_async<NT>(win, win_g, L, tid);
    if (!SPEC) {
        // bin-major sparse bank: ints = m0[F4] | off[..], floats = weights
        stage_table_async<NT>(reinterpret_cast<float*>(csr_i), mel.bin_m0, round_up4(F), tid);
        stage_table_async<NT>(reinterpret_cast<float*>(csr_i + round_up4(F)), mel.bin_off, round_up4(F + 1), tid);
        stage_table_async<NT>(csr_w, mel.bin_w, round_up4(g.csr_floats), tid);
    }
    const int* s_m0 = csr_i;
    const int* s_off = csr_i + round_up4(F);
    const float inv_lam3 = hdr->inv_lam3, cprime = hdr->cprime, centre = hdr->centre;
    float dlam = 0.f;

    // Persistent tile loop with a one-tile look-ahead (see the forward kernel): the span and the dL/dy / y
    // tiles of tile i+1 land in the alternate buffers while tile i is processed.
    const int total_tiles = g.B * g.tiles_per_clip;
    const int xs_words = round_up4(g.span + 4);
    const int tile_words = round_up4(g.M * TS);
    auto issue_tile = [&](int tix_, int slot) {
        const int b_ = tix_ / g.tiles_per_clip;
        const int t0_ = (tix_ - b_ * g.tiles_per_clip) * g.TT;
        const int nfr_ = min(g.TT, g.T - t0_);
        span_issue<NT>(xs_sh + slot * xs_words, x + (long long)b_ * g.x_stride, t0_ * g.hop - L / 2, g.span, g.N,
                       row_aligned, tid);
        // upstream gradient tile (and y when the log is fused): rows are T floats apart in HBM, so the copies
        // are 4-byte async ones, all in flight together; columns past the last frame are zero
        float* gt_ = gt + slot * 2 * tile_words;
        float* yt_ = gt_ + tile_words;
        const long long base = (long long)b_ * g.M * g.T + t0_;
        for (int idx = tid; idx < (g.M << tt_shift); idx += NT) {
            const int m = idx >> tt_shift, tt = idx & (g.TT - 1);   // consecutive lanes: consecutive frames
            const bool ok = tt < nfr_;
            const long long off = base + (long long)m * g.T + (ok ? tt : 0);
            cp_async4_zfill(gt_ + m * TS + tt, gy + off, ok);
            if (do_log) cp_async4_zfill(yt_ + m * TS + tt, y + off, ok);
        }
    };
    const bool ahead = g.lookahead != 0;
    if (ahead && int(blockIdx.x) < total_tiles) issue_tile(blockIdx.x, 0);
    int it = 0;
    for (int tix = blockIdx.x; tix < total_tiles; tix += gridDim.x, ++it) {
        const int b = tix / g.tiles_per_clip;
        const int tile_i = tix - b * g.tiles_per_clip;
        const int t0 = tile_i * g.TT;
        const int nfr = min(g.TT, g.T - t0);
        const int p0 = t0 * g.hop - L / 2;
        const int slot = ahead ? (it & 1) : 0;
        if (!ahead) issue_tile(tix, 0);
        float* xs_cur = xs_sh + slot * xs_words;
        const float* xs = xs_cur + (p0 & 3);
        float* gtc = gt + slot * 2 * tile_words;       // this tile's dL/dmel
        const float* ytc = gtc + tile_words;

        const float mu = __ldg(means + b);
        if (NEED_DX)
            for (int i = gtid; i < g.sub_span; i += GT) priv[i] = 0.f;
        cp_async_wait_all();
        span_centre<NT>(xs_cur, p0, g.span, g.N, mu, tid);
        if (do_log) {
            // dL/dmel = G / (mel + eps) = G * exp(-y); each thread touches the words it copied itself
            for (int idx = tid; idx < (g.M << tt_shift); idx += NT) {
                const int m = idx >> tt_shift, tt = idx & (g.TT - 1);
                gtc[m * TS + tt] *= __expf(-ytc[m * TS + tt]);
            }
        }
        __syncthreads();   // also: every warp is done with the previous tile
        if (ahead && tix + int(gridDim.x) < total_tiles) issue_tile(tix + gridDim.x, slot ^ 1);

        const int f_lo = grp * g.SF;
        const int f_hi = min(f_lo + g.SF, nfr);
        for (int f = f_lo; f < f_hi; f += 2) {
            const float* xa = xs + f * g.hop;
            const float* xb2 = xa + g.hop;
            const bool has_b = (f + 1) < nfr;
            float2 v[E];
#pragma unroll
            for (int i = 0; i < P0::NB; ++i) {
                const int j = gtid + i * GT;
#pragma unroll
                for (int r = 0; r < P0::R; ++r) {
                    const int n = j + r * P0::STR;
                    const float w = win[n];
                    v[i * P0::R + r] = make_float2(xa[n] * w, xb2[n] * w);
                }
            }
            run_passes<L, 0, true>(v, buf, twt, gtid, bar_id);
            // buf <- conj(V),  V = Zhat_a + i*Zhat_b the Hermitian-extended 2*GP*X of both frames
#pragma unroll 2
            for (int k = gtid; k <= L / 2; k += GT) {
                const float2 z1 = buf[k];
                const float2 z2 = buf[(L - k) & (L - 1)];
                const float ar = 0.5f * (z1.x + z2.x), ai = 0.5f * (z1.y - z2.y);
                const float br = 0.5f * (z1.y + z2.y), bi = 0.5f * (z2.x - z1.x);
                float ga, gb;
                if (SPEC) {
                    ga = gtc[k * TS + f];
                    gb = gtc[k * TS + f + 1];
                } else {
                    // <= 2 taps per bin for a triangular bank: both issued together; longer runs loop
                    const int m0 = s_m0[k];
                    const int o0 = s_off[k], o1 = s_off[k + 1];
                    const float w0 = (o0 < o1) ? csr_w[o0] : 0.f;
                    const float w1 = (o0 + 1 < o1) ? csr_w[o0 + 1] : 0.f;
                    const float* gp0 = gtc + m0 * TS + f;
                    const float* gp1 = gtc + min(m0 + 1, g.M - 1) * TS + f;
                    ga = fmaf(w1, gp1[0], w0 * gp0[0]);
                    gb = fmaf(w1, gp1[1], w0 * gp0[1]);
                    for (int o = o0 + 2; o < o1; ++o) {
                        const float w = csr_w[o];
                        const float* gp = gtc + (m0 + (o - o0)) * TS + f;
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
            group_sync<GT>(bar_id);
            // inverse transform as conj(FFT(conj(V))): u_a = Re, u_b = -Im of the FFT output
            P0::gather(v, buf, gtid);
            run_passes<L, 0, false>(v, buf, twt, gtid, bar_id);
            // dL/dlambda, time-domain form: sum_n u[n] * fr[n] * dw/dlambda[n],
            //   dw/dlambda[n] = w[n] * ((n-centre)^2/lambda^3 - cprime)
#pragma unroll
            for (int i = 0; i < PL::NB; ++i) {
#pragma unroll
                for (int r = 0; r < PL::R; ++r) {
                    const int n = PL::out_index(gtid, i, r);
                    const float2 u = v[i * PL::R + r];
                    const float w = win[n];
                    const float d = float(n) - centre;
                    const float q = fmaf(d * d, inv_lam3, -cprime);
                    const float sdot = u.x * xa[n] - (has_b ? u.y * xb2[n] : 0.f);
                    dlam = fmaf(w * q, sdot, dlam);
                    v[i * PL::R + r] = make_float2(u.x * w, -u.y * w);
                }
            }
            if (NEED_DX) {
                // overlap-add (adjoint of the framing) into the group's private span:
                // frame a, barrier, frame b, barrier — no shared-memory atomics
                float* da = priv + (f - f_lo) * g.hop;
#pragma unroll
                for (int i = 0; i < PL::NB; ++i)
#pragma unroll
                    for (int r = 0; r < PL::R; ++r) da[PL::out_index(gtid, i, r)] += v[i * PL::R + r].x;
                group_sync<GT>(bar_id);
                if (has_b) {
                    float* db = da + g.hop;
#pragma unroll
                    for (int i = 0; i < PL::NB; ++i)
#pragma unroll
                        for (int r = 0; r < PL::R; ++r) db[PL::out_index(gtid, i, r)] += v[i * PL::R + r].y;
                }
                group_sync<GT>(bar_id);
            }
        }

        if (NEED_DX) {
            __syncthreads();   // all private spans of this tile are complete
            // merge the groups' sub-spans; crop the centre padding; positions no other tile touches are
            // stored, the halos shared with the neighbouring tiles are reduced into the zeroed dx
            const int span_used = (nfr - 1) * g.hop + L;
            const int excl_lo = (tile_i == 0) ? 0 : (t0 - 1) * g.hop + L / 2;
            const int excl_hi = (tile_i == g.tiles_per_clip - 1) ? g.N : (t0 + nfr) * g.hop - L / 2;
            float* dxb = dx + (long long)b * g.N;
            const float* priv0 = smem_f + lay.grp + 2 * L;
            const int gstep = g.SF * g.hop;
            float part = 0.f;
            for (int i = tid; i < span_used; i += NT) {
                const int p = p0 + i;
                if (p >= 0 && p < g.N) {
                    // groups gi whose sub-span covers i: 0 <= i - gi*gstep < sub_span (one unsigned compare each)
                    float val = 0.f;
#pragma unroll
                    for (int gi = 0; gi < NG; ++gi) {
                        const int o = i - gi * gstep;
                        if (unsigned(o) < unsigned(g.sub_span)) val += priv0[gi * lay.grp_stride + o];
                    }
                    part += val;
                    if (p >= excl_lo && p < excl_hi) dxb[p] = val;
                    else atomicAdd(dxb + p, val);
                }
            }
            // per-clip sum of dx for the mean-removal adjoint (one fp64 atomic per warp)
            double ps = warp_sum(double(part));
            if ((tid & 31) == 0) atomicAdd(clipsum + b, ps);
        }
        __syncthreads();   // next tile restages xs / gt / priv
    }
    cp_async_wait_all();
    const double tot = block_sum(double(dlam), red);
    if (tid == 0) atomicAdd(&hdr->dlam_acc, tot);
}

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
        float* yb = y + (long long)b * g.M * g.T + t0;
        const int tt_shift = ilog2(g.TT);
        for (int idx = tid; idx < (g.M << tt_shift); idx += NT) {
            const int m = idx >> tt_shift, tt = idx & (g.TT - 1);
            if (tt < nfr) {
                float val = tile[m * TS + tt];
                if (do_log) val = __logf(val + g.log_eps);
                yb[(long long)m * g.T + tt] = val;
                ysum += val;
            }
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
        const int tt_shift = ilog2(g.TT);
        for (int idx = tid; idx < (g.M << tt_shift); idx += NT) {
            const int m = idx >> tt_shift, tt = idx & (g.TT - 1);
            float val = 0.f;
            if (tt < nfr) {
                const long long off = ybase + (long long)m * g.T + tt;
                val = __ldg(gy + off);
                if (do_log) val *= __expf(-__ldg(y + off));
            }
            gt[m * TS + tt] = val;
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
    const int per = (((N + chunks_per_clip - 1) / chunks_per_clip) + 3) & ~3;
    const int lo = c * per, hi = min(N, lo + per);
    if ((reinterpret_cast<uintptr_t>(dxb) & 15) == 0) {
        float4* d4 = reinterpret_cast<float4*>(dxb);
        const int q_hi = hi >> 2;
        for (int q = (lo >> 2) + threadIdx.x; q < q_hi; q += 256) {
            float4 v = d4[q];
            v.x -= mu; v.y -= mu; v.z -= mu; v.w -= mu;
            d4[q] = v;
        }
        for (int i = (q_hi << 2) + threadIdx.x; i < hi; i += 256) dxb[i] -= mu;
    } else {
        for (int i = lo + threadIdx.x; i < hi; i += 256) dxb[i] -= mu;
    }
}

// one launch instead of three memsets: dx (halos are reduced into it), per-clip sums, dL/dlambda accumulator
__global__ void __launch_bounds__(256)
dmel_zero_kernel(float* __restrict__ dx, long long n_dx, double* __restrict__ clipsum, int B, WsHeader* hdr) {
    const long long gid = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long nthreads = (long long)gridDim.x * 256;
    if (gid == 0) hdr->dlam_acc = 0.0;
    if (clipsum)
        for (long long i = gid; i < B; i += nthreads) clipsum[i] = 0.0;
    if (dx) {
        if ((reinterpret_cast<uintptr_t>(dx) & 15) == 0) {
            float4* d4 = reinterpret_cast<float4*>(dx);
            const long long n4 = n_dx >> 2;
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            for (long long i = gid; i < n4; i += nthreads) d4[i] = z;
            for (long long i = (n4 << 2) + gid; i < n_dx; i += nthreads) dx[i] = 0.f;
        } else {
            for (long long i = gid; i < n_dx; i += nthreads) dx[i] = 0.f;
        }
    }
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
                for (int o = o0; o < o1; ++o) s = fmaf(__ldg(mel.mel_w + o), gp[k0 + (o - o0)], s);
            }
            if (do_log) s = logf(s + g.log_eps);
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
constexpr int NUM_SMS = 148;

template <int L> static PlanDesc plan_desc_of() {
    PlanDesc d;
    d.np = Plan<L>::NP;
    d.r[0] = Plan<L>::R0; d.r[1] = Plan<L>::R1; d.r[2] = Plan<L>::R2; d.r[3] = Plan<L>::R3;
    return d;
}

struct FftInfo {
    bool ok;
    bool big;      // 8192 / 16384: one CTA per transform, tables in global memory
    int gt;        // threads per group
    int tw_total;  // twiddle table entries
    int ng[2];     // group counts per CTA that are instantiated (largest first)
    PlanDesc desc;
};

static FftInfo fft_info(int n_fft) {
    FftInfo f{};
    switch (n_fft) {
#define DMEL_FFT_CASE(LL, NGA, NGB)                                                                   \
    case LL: f.ok = true; f.gt = LL / Plan<LL>::E; f.tw_total = TwTotal<LL>::value; f.ng[0] = NGA;     \
             f.ng[1] = NGB; f.desc = plan_desc_of<LL>(); break;
        DMEL_FFT_CASE(128, 8, 4)
        DMEL_FFT_CASE(256, 8, 4)
        DMEL_FFT_CASE(512, 8, 4)
        DMEL_FFT_CASE(1024, 4, 2)
        DMEL_FFT_CASE(2048, 2, 1)
        DMEL_FFT_CASE(4096, 1, 1)
#undef DMEL_FFT_CASE
        case 8192: f.ok = true; f.big = true; f.gt = 8192 / Plan<8192>::E; f.ng[0] = f.ng[1] = 1; f.desc.np = -1; break;
        case 16384: f.ok = true; f.big = true; f.gt = 16384 / Plan<16384>::E; f.ng[0] = f.ng[1] = 1; f.desc.np = -1; break;
        default: f.ok = false; f.desc.np = 0; break;
    }
    return f;
}

static bool has_fft_plan(int n_fft) { return fft_info(n_fft).ok; }

struct LaunchPlan {
    int path;  // 0 FFT, 1 direct DFT, 2 big FFT
    int lookahead;
    int NG, SF, TT, span, sub_span, tiles_per_clip, threads;
    int csr_ints, csr_floats;
    size_t smem;
};

static size_t smem_bytes_fft(int L, int NG, int tw_total, int span, int sub_span, int M, int TT, int csr_ints,
                             int csr_floats, bool bwd, bool priv_dx, int lookahead) {
    // backward keeps two tiles (dL/dy and, for the fused log, y); the look-ahead doubles them and the span
    const int tile_words = bwd ? (2 + 2 * lookahead) * round_up4(M * (TT + 1)) : M * (TT + 1);
    return size_t(make_lay(L, NG, tw_total, span, sub_span, tile_words, csr_ints, csr_floats, priv_dx, 1 + lookahead).total) * 4;
}

// Tile shape: NG groups x SF frames per CTA tile.  scripts/tune.py sweeps on B200 (profiles/) show a
// flat optimum around 8-16 frames per tile: smaller tiles pay the per-tile fixed cost (span staging,
// barriers, tile IO) too often, larger ones cost occupancy and leave a coarse tail over the 148 SMs.
// The preferred (groups, frames per group) per transform length is tabulated below and shrunk
// until the tile fits shared memory or the clip.
// DMEL_NG / DMEL_SF in the environment pin the choice (tuning experiments).
static int plan_launch(int64_t B, int64_t N, int n_fft, int hop, int rows, int csr_ints, int csr_floats, bool bwd,
                       bool need_dx, LaunchPlan* lp) {
    (void)B;
    const int64_t T64 = N / hop + 1;
    if (T64 > INT32_MAX) return DMEL_ERR_BAD_GEOMETRY;
    const int T = int(T64);
    lp->csr_ints = csr_ints;
    lp->csr_floats = csr_floats;
    const FftInfo fi = fft_info(n_fft);
    if (!fi.ok) {
        if (n_fft < 2 || n_fft > DMEL_MAX_DFT_NFFT || (n_fft & 1)) return DMEL_ERR_UNSUPPORTED_NFFT;
        const int F = n_fft / 2 + 1;
        lp->path = 1;
        lp->NG = 1; lp->SF = 1; lp->TT = 1;
        lp->span = n_fft; lp->sub_span = n_fft;
        lp->tiles_per_clip = T;
        lp->threads = DFT_THREADS;
        lp->smem = size_t(round_up4(n_fft)) * 4 + size_t(F) * 8 + size_t(round_up4(F) + 8) * 4 + 512;
        if (lp->smem > SMEM_MAX) return DMEL_ERR_SMEM_EXCEEDED;
        return DMEL_OK;
    }
    if (fi.big) {
        // largest power-of-two tile that fits next to the L-point buffer (spectrogram mode keeps no tile)
        const bool spec = csr_floats == 0;
        int t_pow2 = 2;
        while (t_pow2 < T) t_pow2 <<= 1;
        for (int TT = std::min(16, t_pow2); TT >= 2; TT >>= 1) {
            const int64_t span64 = int64_t(TT - 1) * hop + n_fft;
            if (span64 > (1 << 20)) continue;
            const size_t sm = size_t(make_big_lay(n_fft, int(span64), spec ? 0 : rows * (TT + 1), bwd && need_dx).total) * 4;
            if (sm > SMEM_MAX) continue;
            lp->path = 2;
            lp->NG = 1; lp->SF = TT; lp->TT = TT;
            lp->span = int(span64); lp->sub_span = int(span64);
            lp->tiles_per_clip = (T + TT - 1) / TT;
            lp->threads = fi.gt;
            lp->smem = sm;
            return DMEL_OK;
        }
        return DMEL_ERR_SMEM_EXCEEDED;
    }
    int pin_ng = 0, pin_sf = 0;
    if (const char* e = getenv("DMEL_NG")) pin_ng = atoi(e);
    if (const char* e = getenv("DMEL_SF")) pin_sf = atoi(e);
    // measured optima (scripts/tune.py on B200, profiles/*tune*.log): (groups, frames per group)
    int ng_pref, sf_pref;
    switch (n_fft) {
        case 1024: ng_pref = (bwd && need_dx) ? 4 : 2; sf_pref = (bwd && need_dx) ? 4 : 8; break;
        case 2048: ng_pref = 1; sf_pref = 8; break;
        case 4096: ng_pref = 1; sf_pref = 32; break;
        default: ng_pref = 4; sf_pref = 4; break;    // one warp per transform
    }
    for (int attempt = 0; attempt < 2; ++attempt) {
        for (int ni = 0; ni < 2; ++ni) {
            // first the preferred NG, then the other instantiated one
            const int NG = (ni == 0) ? ng_pref : (fi.ng[0] == ng_pref ? fi.ng[1] : fi.ng[0]);
            if (ni == 1 && NG == ng_pref) continue;
            if (pin_ng && NG != pin_ng) continue;
            int sf0 = pin_sf ? pin_sf : sf_pref;
            // no point in a tile much longer than the clip
            while (!pin_sf && sf0 > 2 && NG * (sf0 / 2) >= T) sf0 >>= 1;
            for (int SF = sf0; SF >= 2; SF >>= 1) {
                const int tt = NG * SF;
                const int64_t span64 = int64_t(tt - 1) * hop + n_fft, sub64 = int64_t(SF - 1) * hop + n_fft;
                if (span64 <= (1 << 20)) {
                    size_t sm = smem_bytes_fft(n_fft, NG, fi.tw_total, int(span64), int(sub64), rows, tt, csr_ints, csr_floats,
                                               bwd, bwd && need_dx, 0);
                    if (sm <= SMEM_MAX) {
                        // one-tile look-ahead (double buffers) only if it does not cost a resident CTA
                        const size_t sm2 = smem_bytes_fft(n_fft, NG, fi.tw_total, int(span64), int(sub64), rows, tt, csr_ints,
                                                          csr_floats, bwd, bwd && need_dx, 1);
                        const int warps_cta = NG * fi.gt / 32;
                        const int reg_cap = std::max(1, (n_fft <= 512 && NG == 4 && !(bwd && need_dx) ? 20 : 16) / warps_cta);
                        auto ctas = [&](size_t b) { return std::min<int>(reg_cap, int(SMEM_MAX / (b + 1024))); };
                        lp->lookahead = 0;
                        if (!getenv("DMEL_NO_LOOKAHEAD") && sm2 <= SMEM_MAX && ctas(sm2) >= ctas(sm)) {
                            lp->lookahead = 1;
                            sm = sm2;
                        }
                        if (getenv("DMEL_FORCE_LOOKAHEAD") && sm2 <= SMEM_MAX) { lp->lookahead = 1; sm = sm2; }
                        lp->path = 0;
                        lp->NG = NG; lp->SF = SF; lp->TT = tt;
                        lp->span = int(span64); lp->sub_span = int(sub64);
                        lp->tiles_per_clip = (T + tt - 1) / tt;
                        lp->threads = NG * fi.gt;
                        lp->smem = sm;
                        return DMEL_OK;
                    }
                }
                if (pin_sf) break;
            }
        }
        pin_ng = 0;   // a pinned shape that does not fit falls back to the automatic choice
        pin_sf = 0;
    }
    return DMEL_ERR_SMEM_EXCEEDED;
}

// ---- optional instrumentation for bench.py: launch counter + per-kernel CUDA-event timing.
// Off by default; when on, every main-kernel launch is bracketed by a pair of events recorded on
// the launching stream (no synchronisation until dmel_profile_read).
enum { PROF_PREP = 0, PROF_FWD = 1, PROF_BWD = 2, PROF_FINALIZE = 3, PROF_ZERO = 4, PROF_KINDS = 5 };
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

// persistent launch: one resident wave of CTAs (occupancy x SM count), each walking tiles with stride gridDim
template <typename K>
static unsigned persistent_grid(K kernel, const LaunchPlan& lp, int64_t total_tiles) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, lp.threads, lp.smem) != cudaSuccess || per_sm < 1) {
        (void)cudaGetLastError();
        per_sm = 1;
    }
    int sms = NUM_SMS, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) sms = v;
    }
    const int64_t cap = int64_t(per_sm) * sms;
    return unsigned(std::min<int64_t>(total_tiles, cap));
}

template <int L, int NG, bool SPEC>
static int launch_fwd_t(const float* x, float* y, const Geom& g, const MelDev& mel, const WsView& ws,
                        const LaunchPlan& lp, cudaStream_t st) {
    int rc;
    if ((rc = set_smem(dmel_fwd_kernel<L, NG, SPEC>, lp.smem))) return rc;
    const unsigned grid = persistent_grid(dmel_fwd_kernel<L, NG, SPEC>, lp, int64_t(g.B) * g.tiles_per_clip);
    ProfScope prof(PROF_FWD, st);
    dmel_fwd_kernel<L, NG, SPEC><<<grid, lp.threads, lp.smem, st>>>(x, y, g, mel, ws.win, ws.tw, ws.means, ws.hdr);
    return check_launch("dmel_fwd_kernel");
}

template <int L, int NG, bool SPEC, bool NEED_DX>
static int launch_bwd_t(const float* x, const float* y, const float* gy, float* dx, const Geom& g, const MelDev& mel,
                        const WsView& ws, const LaunchPlan& lp, cudaStream_t st) {
    int rc;
    if ((rc = set_smem(dmel_bwd_kernel<L, NG, SPEC, NEED_DX>, lp.smem))) return rc;
    const unsigned grid = persistent_grid(dmel_bwd_kernel<L, NG, SPEC, NEED_DX>, lp, int64_t(g.B) * g.tiles_per_clip);
    ProfScope prof(PROF_BWD, st);
    dmel_bwd_kernel<L, NG, SPEC, NEED_DX>
        <<<grid, lp.threads, lp.smem, st>>>(x, y, gy, dx, g, mel, ws.win, ws.tw, ws.means, ws.hdr, ws.clipsum);
    return check_launch("dmel_bwd_kernel");
}

template <int L, int NG>
static int launch_fwd(const float* x, float* y, const Geom& g, const MelDev& mel, bool spec, const WsView& ws,
                      const LaunchPlan& lp, cudaStream_t st) {
    return spec ? launch_fwd_t<L, NG, true>(x, y, g, mel, ws, lp, st) : launch_fwd_t<L, NG, false>(x, y, g, mel, ws, lp, st);
}

template <int L, int NG>
static int launch_bwd(const float* x, const float* y, const float* gy, float* dx, const Geom& g, const MelDev& mel,
                      bool spec, const WsView& ws, const LaunchPlan& lp, cudaStream_t st) {
    if (spec) {
        return dx ? launch_bwd_t<L, NG, true, true>(x, y, gy, dx, g, mel, ws, lp, st)
                  : launch_bwd_t<L, NG, true, false>(x, y, gy, dx, g, mel, ws, lp, st);
    }
    return dx ? launch_bwd_t<L, NG, false, true>(x, y, gy, dx, g, mel, ws, lp, st)
              : launch_bwd_t<L, NG, false, false>(x, y, gy, dx, g, mel, ws, lp, st);
}

template <int L>
static int launch_big_fwd(const float* x, float* y, const Geom& g, const MelDev& mel, bool spec, const WsView& ws,
                          const LaunchPlan& lp, cudaStream_t st) {
    const unsigned grid = unsigned(int64_t(g.B) * g.tiles_per_clip);
    int rc;
    ProfScope prof(PROF_FWD, st);
    if (spec) {
        if ((rc = set_smem(dmel_big_fwd_kernel<L, true>, lp.smem))) return rc;
        dmel_big_fwd_kernel<L, true><<<grid, lp.threads, lp.smem, st>>>(x, y, g, mel, ws.win, ws.tw, ws.means, ws.hdr);
    } else {
        if ((rc = set_smem(dmel_big_fwd_kernel<L, false>, lp.smem))) return rc;
        dmel_big_fwd_kernel<L, false><<<grid, lp.threads, lp.smem, st>>>(x, y, g, mel, ws.win, ws.tw, ws.means, ws.hdr);
    }
    return check_launch("dmel_big_fwd_kernel");
}

template <int L, bool SPEC, bool NEED_DX>
static int launch_big_bwd_t(const float* x, const float* y, const float* gy, float* dx, const Geom& g, const MelDev& mel,
                            const WsView& ws, const LaunchPlan& lp, cudaStream_t st) {
    const unsigned grid = unsigned(int64_t(g.B) * g.tiles_per_clip);
    int rc;
    if ((rc = set_smem(dmel_big_bwd_kernel<L, SPEC, NEED_DX>, lp.smem))) return rc;
    ProfScope prof(PROF_BWD, st);
    dmel_big_bwd_kernel<L, SPEC, NEED_DX>
        <<<grid, lp.threads, lp.smem, st>>>(x, y, gy, dx, g, mel, ws.win, ws.tw, ws.means, ws.hdr, ws.clipsum);
    return check_launch("dmel_big_bwd_kernel");
}

template <int L>
static int launch_big_bwd(const float* x, const float* y, const float* gy, float* dx, const Geom& g, const MelDev& mel,
                          bool spec, const WsView& ws, const LaunchPlan& lp, cudaStream_t st) {
    if (spec) {
        return dx ? launch_big_bwd_t<L, true, true>(x, y, gy, dx, g, mel, ws, lp, st)
                  : launch_big_bwd_t<L, true, false>(x, y, gy, dx, g, mel, ws, lp, st);
    }
    return dx ? launch_big_bwd_t<L, false, true>(x, y, gy, dx, g, mel, ws, lp, st)
              : launch_big_bwd_t<L, false, false>(x, y, gy, dx, g, mel, ws, lp, st);
}

// (n_fft, NG) -> instantiation; the NG values must match fft_info()
#define DMEL_DISPATCH(n_fft, ng, CALL)                                   \
    switch ((n_fft) * 16 + (ng)) {                                       \
        case 128 * 16 + 8: rc = CALL(128, 8); break;                     \
        case 128 * 16 + 4: rc = CALL(128, 4); break;                     \
        case 256 * 16 + 8: rc = CALL(256, 8); break;                     \
        case 256 * 16 + 4: rc = CALL(256, 4); break;                     \
        case 512 * 16 + 8: rc = CALL(512, 8); break;                     \
        case 512 * 16 + 4: rc = CALL(512, 4); break;                     \
        case 1024 * 16 + 4: rc = CALL(1024, 4); break;                   \
        case 1024 * 16 + 2: rc = CALL(1024, 2); break;                   \
        case 2048 * 16 + 2: rc = CALL(2048, 2); break;                   \
        case 2048 * 16 + 1: rc = CALL(2048, 1); break;                   \
        case 4096 * 16 + 1: rc = CALL(4096, 1); break;                   \
        default: rc = DMEL_ERR_UNSUPPORTED_NFFT;                         \
    }

static int validate(const float* x, int64_t B, int64_t N, int64_t stride, int32_t win, int32_t n_fft, int32_t hop,
                    int32_t n_mels, const dmel_melbank* bank, void* ws, size_t ws_bytes_in) {
    if (!x || !ws) return DMEL_ERR_NULL_POINTER;
    if (B <= 0 || N <= 0 || hop <= 0 || win <= 0 || n_fft < win || n_mels < 0 || stride < N) return DMEL_ERR_BAD_GEOMETRY;
    if (B > INT32_MAX / 2 || N > INT32_MAX / 2) return DMEL_ERR_BAD_GEOMETRY;
    if (n_mels > 0) {
        if (!bank) return DMEL_ERR_NULL_POINTER;
        if (bank->n_mels != n_mels || bank->n_freqs != n_fft / 2 + 1) return DMEL_ERR_BAD_MELBANK;
        if (bank->mel_nnz < 0 || bank->bin_nnz < 0 || bank->mel_nnz > (1 << 22) || bank->bin_nnz > (1 << 22))
            return DMEL_ERR_BAD_MELBANK;
        if (!bank->mel_k0 || !bank->mel_off || !bank->mel_w || !bank->bin_m0 || !bank->bin_off || !bank->bin_w)
            return DMEL_ERR_NULL_POINTER;
        const uintptr_t al = reinterpret_cast<uintptr_t>(bank->mel_k0) | reinterpret_cast<uintptr_t>(bank->mel_off) |
                             reinterpret_cast<uintptr_t>(bank->mel_w) | reinterpret_cast<uintptr_t>(bank->bin_m0) |
                             reinterpret_cast<uintptr_t>(bank->bin_off) | reinterpret_cast<uintptr_t>(bank->bin_w);
        if (al & 15) return DMEL_ERR_BAD_MELBANK;  // the arrays are staged with 16-byte async copies
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
                                                                flags, ws, fft_info(n_fft).desc);
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

extern "C" DMEL_API int dmel_query_launch(int64_t batch, int64_t n_points, int32_t n_fft, int32_t hop, int32_t rows,
                                          int32_t bank_nnz, int32_t which, int32_t* tile_frames,
                                          int32_t* groups, int32_t* smem_bytes, int32_t* threads, int32_t* path) {
    if (batch <= 0 || n_points <= 0 || hop <= 0 || n_fft <= 0 || rows <= 0 || bank_nnz < 0) return DMEL_ERR_BAD_GEOMETRY;
    LaunchPlan lp{};
    const int F = n_fft / 2 + 1;
    const bool spec = bank_nnz == 0;
    const int ci = spec ? 0 : (which == 0 ? round_up4(rows) + round_up4(rows + 1) : round_up4(F) + round_up4(F + 1));
    int rc = plan_launch(batch, n_points, n_fft, hop, rows, ci, bank_nnz, which != 0, which == 1, &lp);
    if (rc) return rc;
    if (tile_frames) *tile_frames = lp.TT;
    if (groups) *groups = lp.NG;
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
    g.SF = lp.SF;
    g.tiles_per_clip = lp.tiles_per_clip;
    g.span = lp.span;
    g.sub_span = lp.sub_span;
    g.csr_ints = lp.csr_ints;
    g.csr_floats = lp.csr_floats;
    g.x_stride = stride;
    g.log_eps = log_eps;
    g.row_aligned = 0;
    g.lookahead = lp.lookahead;
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
    if ((rc = plan_launch(batch, n_points, n_fft, hop, rows, spec ? 0 : round_up4(n_mels) + round_up4(n_mels + 1), spec ? 0 : bank->mel_nnz, false,
                          false, &lp)))
        return rc;
    if (int64_t(batch) * lp.tiles_per_clip > INT32_MAX) return DMEL_ERR_BAD_GEOMETRY;
    WsView ws = ws_view(workspace, batch, n_fft);
    if ((rc = run_prep(x, batch, n_points, x_row_stride, fabsf(lambda), win_length, n_fft, flags, ws, st)))
        return rc;
    Geom g = make_geom(batch, n_points, x_row_stride, n_fft, hop, rows, log_eps, lp);
    g.row_aligned = ((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (x_row_stride & 3) == 0) ? 1 : 0;
    const MelDev mel = mel_dev(bank);
    if (lp.path == 1) {
        if ((rc = set_smem(dmel_dft_kernel<false, false>, lp.smem))) return rc;
        ProfScope prof(PROF_FWD, st);
        dmel_dft_kernel<false, false><<<unsigned(int64_t(g.B) * g.T), DFT_THREADS, lp.smem, st>>>(
            x, nullptr, nullptr, y, nullptr, g, mel, spec ? 1 : 0, ws.win, ws.tw, ws.means, ws.hdr, ws.clipsum);
        return check_launch("dmel_dft_kernel<fwd>");
    }
    if (lp.path == 2) {
        return n_fft == 8192 ? launch_big_fwd<8192>(x, y, g, mel, spec, ws, lp, st)
                             : launch_big_fwd<16384>(x, y, g, mel, spec, ws, lp, st);
    }
#define CALL_FWD(LL, NGG) launch_fwd<LL, NGG>(x, y, g, mel, spec, ws, lp, st)
    DMEL_DISPATCH(n_fft, lp.NG, CALL_FWD)
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
    if ((rc = plan_launch(batch, n_points, n_fft, hop, rows, spec ? 0 : round_up4(n_fft / 2 + 1) + round_up4(n_fft / 2 + 2),
                          spec ? 0 : bank->bin_nnz, true, need_dx, &lp)))
        return rc;
    if (int64_t(batch) * lp.tiles_per_clip > INT32_MAX) return DMEL_ERR_BAD_GEOMETRY;
    WsView ws = ws_view(workspace, batch, n_fft);
    if (!prepared) {
        if ((rc = run_prep(x, batch, n_points, x_row_stride, fabsf(lambda), win_length, n_fft, flags, ws, st)))
            return rc;
    }
    // zero the accumulators (and dx: halo positions are reduced with atomics) in one launch
    {
        const long long n_dx = need_dx ? (long long)batch * n_points : 0;
        long long blocks = need_dx ? (n_dx / 4 + 255) / 256 : 1;
        blocks = std::max<long long>(1, std::min<long long>(blocks, 8LL * NUM_SMS));
        ProfScope prof(PROF_ZERO, st);
        dmel_zero_kernel<<<unsigned(blocks), 256, 0, st>>>(grad_x, n_dx, need_dx ? ws.clipsum : nullptr, int(batch), ws.hdr);
        if ((rc = check_launch("dmel_zero_kernel"))) return rc;
    }
    Geom g = make_geom(batch, n_points, x_row_stride, n_fft, hop, rows, log_eps, lp);
    g.row_aligned = ((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (x_row_stride & 3) == 0) ? 1 : 0;
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
    } else if (lp.path == 2) {
        rc = n_fft == 8192 ? launch_big_bwd<8192>(x, y, grad_y, grad_x, g, mel, spec, ws, lp, st)
                           : launch_big_bwd<16384>(x, y, grad_y, grad_x, g, mel, spec, ws, lp, st);
        if (rc) return rc;
    } else {
#define CALL_BWD(LL, NGG) launch_bwd<LL, NGG>(x, y, grad_y, grad_x, g, mel, spec, ws, lp, st)
        DMEL_DISPATCH(n_fft, lp.NG, CALL_BWD)
#undef CALL_BWD
        if (rc) return rc;
    }
    const float sign = lambda > 0.f ? 1.f : (lambda < 0.f ? -1.f : 0.f);
    int chunks = 1, grid = 1;
    if (need_dx) {
        chunks = int((n_points + 4095) / 4096);
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

extern "C" DMEL_API size_t dmel_workspace_y_sum_offset(void) { return offsetof(WsHeader, y_sum); }
