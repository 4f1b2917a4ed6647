// dmel_tiled2_kernels.cuh — second-generation persistent tiled forward / backward kernels (n_fft 256..4096),
// built on fft_core.cuh.  Same decomposition as the first generation (dmel_tiled_kernels.cuh: one CTA = one tile of
// TT frames of one clip, NG independent FFT groups, two real frames per complex transform, persistent grid), with the
// per-transform instruction count roughly halved:
//   * a thread owns the points / bins  n = gtid + GT*q  (q = 0..15) before AND after every transform, so the window
//     (pre-scaled by 1/2, which removes the 1/2 of the packed-pair split), the derivative window (JVP) and the
//     per-point dL/dlambda accumulators are registers, and all shared-memory addresses are base + immediate;
//   * the packed-pair split reads only the partner half: the thread already holds Z[k]; Z[L-k] comes from the
//     upper half of the buffer that the partner thread stored (8 stores + 8 loads instead of 16 + 18);
//   * backward: conj(V[k]) = s*conj(Z[k]) + d*Z[L-k],  s = ga+gb, d = ga-gb  (ga, gb the two frames' dL/dP at bin
//     k) — 4 FMAs per bin straight from Z, no separate split; the thread that owns k also forms V[L-k] and hands it
//     to the partner through the slot it has just read; the inverse transform starts from registers;
//   * dL/dlambda = sum_n dw[n] * (sum_frames u[n]*fr[n]): the frame sum is accumulated per owned n, the derivative
//     window is applied once at the end of the kernel;
//   * the bin-major bank is packed into one 16-byte record per bin (row offset, two weights) in shared memory.
// Inputs may be fp32, int16 or fp16 (Geom.x_dtype): 16-bit samples are widened while the span is staged, there is
// no intermediate fp32 tensor (the PCM input path, reference train.py:34 / datasets.py:169).
//
// Reference chain replaced: models.py:33-56 -> time_frequency.py:21-58 (Gaussian window, torch.stft, |.|^2)
// -> melscale_fbanks matmul -> torch.log(s + 1e-10) (models.py:72-73) and autograd's backward of it.
#pragma once
#include <cuda_fp16.h>

#include "dmel_common.cuh"
#include "fft_core.cuh"

namespace dmel {
namespace v2 {

// ----------------------------------------------------------------------------
// shared-memory carve-up (offsets in 4-byte words, multiples of 4)
// ----------------------------------------------------------------------------
struct Lay2 {
    int tw, xs, tile, meta, w, red, grp, grp_stride, total;
};

// meta_words: forward 2*M (int2 per mel row), backward 4*F (16-byte record per bin); w_words: forward mel_nnz
__host__ __device__ inline Lay2 make_lay2(int tw_smem, int buf_f2, int NG, int span, int sub_span, int tile_words,
                                          int meta_words, int w_words, bool priv_dx, int span_bufs) {
    Lay2 l;
    int o = 0;
    l.tw = o;    o += round_up4(2 * tw_smem);
    l.xs = o;    o += span_bufs * round_up4(span + 4);
    l.tile = o;  o += round_up4(tile_words);
    l.meta = o;  o += round_up4(meta_words);
    l.w = o;     o += round_up4(w_words);
    l.red = o;   o += 64;
    l.grp = o;
    l.grp_stride = round_up4(2 * buf_f2) + (priv_dx ? round_up4(sub_span) : 0);
    o += NG * l.grp_stride;
    l.total = o;
    return l;
}

// forward, mel mode: the weight region holds either the segment table (seg_capacity float2) or the row-gather
// weights (mel_nnz floats) — the kernel picks at run time (SegHeader.use_seg), so it is sized for the larger
__host__ __device__ inline int fwd_weight_words(int n_fft, int mel_nnz) {
    const int seg = 2 * seg_capacity(n_fft);
    return seg > mel_nnz ? seg : mel_nnz;
}

// ----------------------------------------------------------------------------
// 16-bit inputs: load + widen + centre in one step (no cp.async: the converted value is what is stored)
// storage convention as span_issue: word s = p - (p0 & ~3)
// ----------------------------------------------------------------------------
__device__ __forceinline__ float widen16(unsigned short raw, int dtype) {
    return dtype == DMEL_X_I16 ? float(short(raw)) : __half2float(__ushort_as_half(raw));
}

template <int NT>
__device__ __forceinline__ void span_load16(float* xs_sh, const unsigned short* __restrict__ xb, int p0, int span, int N,
                                            bool row_aligned8, float mu, int dtype, int tid) {
    const int shift = p0 & 3;
    const int nq = (span + shift + 3) >> 2;
    for (int q = tid; q < nq; q += NT) {
        const int i0 = 4 * q - shift;
        const int p = p0 + i0;                     // multiple of 4
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row_aligned8 && p >= 0 && p + 3 < N) {
            const uint2 r = __ldg(reinterpret_cast<const uint2*>(xb + p));
            v.x = widen16((unsigned short)(r.x & 0xffffu), dtype) - mu;
            v.y = widen16((unsigned short)(r.x >> 16), dtype) - mu;
            v.z = widen16((unsigned short)(r.y & 0xffffu), dtype) - mu;
            v.w = widen16((unsigned short)(r.y >> 16), dtype) - mu;
        } else {
            if (p + 0 >= 0 && p + 0 < N) v.x = widen16(__ldg(xb + p + 0), dtype) - mu;
            if (p + 1 >= 0 && p + 1 < N) v.y = widen16(__ldg(xb + p + 1), dtype) - mu;
            if (p + 2 >= 0 && p + 2 < N) v.z = widen16(__ldg(xb + p + 2), dtype) - mu;
            if (p + 3 >= 0 && p + 3 < N) v.w = widen16(__ldg(xb + p + 3), dtype) - mu;
        }
        *reinterpret_cast<float4*>(xs_sh + 4 * q) = v;
    }
}

// one call stages the span of a tile whatever the input type.  fp32: asynchronous copies (span_issue); the caller
// waits and calls span_finish.  16-bit: nothing is issued here, span_finish does the whole job.
template <int NT>
__device__ __forceinline__ void span_begin(float* xs_sh, const void* x, long long row, const Geom& g, int p0, int tid) {
    if (g.x_dtype == DMEL_X_F32)
        span_issue<NT>(xs_sh, static_cast<const float*>(x) + row * g.x_stride, p0, g.span, g.N, g.row_aligned != 0, tid);
}
template <int NT>
__device__ __forceinline__ void span_finish(float* xs_sh, const void* x, long long row, const Geom& g, int p0, float mu,
                                            int tid) {
    if (g.x_dtype == DMEL_X_F32)
        span_centre<NT>(xs_sh, p0, g.span, g.N, mu, tid);
    else
        span_load16<NT>(xs_sh, static_cast<const unsigned short*>(x) + row * g.x_stride, p0, g.span, g.N,
                        g.row_aligned != 0, mu, g.x_dtype, tid);
}

// register budget: the one-warp-or-less-per-transform forward kernels are compiled for 5 resident 128-thread CTAs
__host__ __device__ constexpr int min_ctas2(int NT, bool cap96) {
    return cap96 ? (640 / NT > 0 ? 640 / NT : 1) : (512 / NT > 0 ? 512 / NT : 1);
}

#ifdef DMEL2_FWD_THREADS   // experiment: resident threads per SM the forward is compiled for (register budget)
#define DMEL2_FWD_MINB(NT, CAP96) ((DMEL2_FWD_THREADS) / (NT) > 0 ? (DMEL2_FWD_THREADS) / (NT) : 1)
#else
#define DMEL2_FWD_MINB(NT, CAP96) min_ctas2(NT, CAP96)
#endif

// ----------------------------------------------------------------------------
// forward
// ----------------------------------------------------------------------------
template <int L, int NG, bool SPEC, bool JVP>
__global__ void __launch_bounds__(NG * (L / 16), L == 4096 ? 1 : DMEL2_FWD_MINB(NG * (L / 16), L <= 512 && !JVP))
dmel2_fwd_kernel(const void* __restrict__ x, float* __restrict__ y, float* __restrict__ dy, Geom g, MelDev mel,
                 const float* __restrict__ win_g, const float2* __restrict__ tw_g, const float* __restrict__ means,
                 WsHeader* hdr) {
    using F = Fft<L>;
    constexpr int GT = F::GT, NT = NG * GT;
    extern __shared__ __align__(16) float smem_f[];
    const int TS = g.TT + 1;
    const int plane = round_up4(g.M * TS);   // JVP keeps two planes: mel and d(mel)/d|lambda|
    const Lay2 lay = make_lay2(F::TW_SMEM, F::BUF, NG, g.span, g.sub_span, JVP ? 2 * plane : g.M * TS, SPEC ? 0 : 2 * g.M,
                               SPEC ? 0 : (mel.seg.hdr != nullptr ? fwd_weight_words(L, g.csr_floats) : g.csr_floats), false,
                               1 + g.lookahead);
    float2* tws = reinterpret_cast<float2*>(smem_f + lay.tw);
    float* xs_sh = smem_f + lay.xs;
    float* tile = smem_f + lay.tile;
    float* tile2 = tile + plane;
    int2* meta = reinterpret_cast<int2*>(smem_f + lay.meta);
    float* csr_w = smem_f + lay.w;

    const int tid = threadIdx.x;
    const int grp = tid / GT, gtid = tid - grp * GT, bar_id = group_sync_arg<GT>(grp);
    float2* buf = reinterpret_cast<float2*>(smem_f + lay.grp + grp * lay.grp_stride);
    const int tt_shift = ilog2(g.TT);
    const bool do_log = g.log_eps >= 0.f;
    const bool affine = !SPEC && mel.out_scale != nullptr;

    // constant tables: once per CTA
    {
        const float2* src = tw_g + (F::TW1_REGS ? F::TW1_N : 0);
        stage_table_async<NT>(smem_f + lay.tw, src, 2 * F::TW_SMEM, tid);
    }
    // mel projection: segment-owner form when the prep kernel could build its tables (triangular banks), else the
    // row gather over the mel-major bank
    const bool use_seg = !SPEC && mel.seg.hdr != nullptr && mel.seg.hdr->use_seg != 0;
    const float2* segw = reinterpret_cast<const float2*>(csr_w);
    int seg_np = 0, seg_dummy = 0;
    int pdest[9];
    __shared__ int s_seg[2 * SEG_MAX_PASSES];   // base | maxlen of every pass
    if (use_seg) {
        if (tid < SEG_MAX_PASSES) {
            s_seg[tid] = mel.seg.hdr->base[tid];
            s_seg[SEG_MAX_PASSES + tid] = mel.seg.hdr->maxlen[tid];
        }
        seg_np = mel.seg.hdr->npass;
        seg_dummy = mel.seg.hdr->nslot * (GT + 1);
        stage_table_async<NT>(csr_w, mel.seg.w, round_up4(2 * (seg_dummy + 1)), tid);
#pragma unroll
        for (int q = 0; q < 8; ++q) pdest[q] = __ldg(mel.seg.dest + gtid + GT * q);
        pdest[8] = gtid == 0 ? __ldg(mel.seg.dest + L / 2) : seg_dummy;
        // slots that no bin is written to are read with zero weights: they must hold finite values
        for (int i = gtid; i < F::BUF; i += GT) buf[i] = make_float2(0.f, 0.f);
    } else if (!SPEC) {
        stage_table_async<NT>(csr_w, mel.mel_w, round_up4(g.csr_floats), tid);
        // one record per mel row: first bin | 4-tap chunks << 16, weight offset (runs are zero-padded to 4 taps, dmel.h)
        for (int m = tid; m < g.M; m += NT) {
            const int o0 = __ldg(mel.mel_off + m), o1 = __ldg(mel.mel_off + m + 1);
            meta[m] = make_int2(__ldg(mel.mel_k0 + m) | (((o1 - o0) >> 2) << 16), o0);
        }
#pragma unroll
        for (int q = 0; q < 9; ++q) pdest[q] = 0;
    }
    typename F::Ctx ctx;
    F::init(ctx, buf, tws, tw_g, gtid, bar_id);
    // window of the owned points, pre-scaled by 1/2: Z' = Z/2, so X_a = Z'[k] + conj(Z'[L-k]) needs no halving
    float wh[16];
    float whd[JVP ? 16 : 1];
    {
        float inv_lam3 = 0.f, cprime = 0.f, centre = 0.f;
        if (JVP) { inv_lam3 = hdr->inv_lam3; cprime = hdr->cprime; centre = hdr->centre; }
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int n = gtid + GT * q;
            wh[q] = 0.5f * __ldg(win_g + n);
            if (JVP) {
                // dw/d|lambda|[n] = w[n] * ((n-centre)^2/lambda^3 - cprime)
                const float d = float(n) - centre;
                whd[q] = wh[q] * fmaf(d * d, inv_lam3, -cprime);
            }
        }
    }

    // Persistent tile loop with a one-tile look-ahead: while tile i is transformed, the span of tile i+1 is
    // already landing in the other span buffer (cp.async), so the staging latency is off the critical path.
    const int total_tiles = g.B * g.tiles_per_clip;
    const int xs_words = round_up4(g.span + 4);
    const bool ahead = g.lookahead != 0;
    float ysum = 0.f;
    if (ahead && int(blockIdx.x) < total_tiles) {
        const int b0 = blockIdx.x / g.tiles_per_clip;
        const int p00 = (blockIdx.x - b0 * g.tiles_per_clip) * g.TT * g.hop - L / 2;
        span_begin<NT>(xs_sh, x, b0, g, p00, tid);
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

        if (!ahead) span_begin<NT>(xs_cur, x, b, g, p0, tid);
        const float mu = __ldg(means + b);
        cp_async_wait_all();
        span_finish<NT>(xs_cur, x, b, g, p0, mu, tid);
        __syncthreads();   // also: every warp is done with the previous tile (its span buffer, the output tile)
        if (ahead) {
            const int nx = tix + gridDim.x;
            if (nx < total_tiles) {
                const int bn = nx / g.tiles_per_clip;
                const int pn = (nx - bn * g.tiles_per_clip) * g.TT * g.hop - L / 2;
                span_begin<NT>(xs_sh + ((it + 1) & 1) * xs_words, x, bn, g, pn, tid);
            }
        }

        // frames per group: a short last tile of a clip is re-split evenly over the groups (whole pairs)
        const int sfe = tile_frames_per_group<NG, !JVP>(nfr, g.TT, g.SF);
        const int f_hi = min(grp * sfe + sfe, nfr);
        for (int f = grp * sfe; f < f_hi; f += (JVP ? 1 : 2)) {
            const float* xa = xs + f * g.hop + gtid;
            const float* xb = xa + g.hop;
            float2 z[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const float a = xa[GT * q];
                if (JVP) {
                    z[q] = px::mul2(px::bc(a), make_float2(wh[q], whd[q]));
                } else {
                    const float c = xb[GT * q];
                    z[q] = px::mul2(make_float2(a, c), px::bc(wh[q]));
                }
            }
#ifndef DMEL2_EXP_NOFFT
            F::run(z, ctx);
#endif
            // packed pair -> two one-sided power spectra (time_frequency.py:53).  The thread holds Z[k] for its
            // bins; the partner values Z[L-k], k < L/2, are the upper half of the spectrum: only that half goes
            // through shared memory (linear layout; Z[0] mirrored at index L).
            gsync<GT>(bar_id);
#pragma unroll
            for (int q = 8; q < 16; ++q) buf[gtid + GT * q] = z[q];
            if (gtid == 0) buf[L] = z[0];
            gsync<GT>(bar_id);
            const float2* lp = buf + L - gtid;
            float* tcol = tile + gtid * TS + f;
            auto split = [&](const float2 z1, const float2 z2, float& pa, float& pb) {
                // X_a = (ar, ai), X_b = (br, bi):  (ar, br) = z1 + z2,  (-bi, ai) = z1 - z2
                const float2 sm = px::add2(z1, z2), df = px::sub2(z1, z2);
                // pair: both powers.  JVP: (ar,ai) = X, (br,bi) = X' -> P = |X|^2, dP/d|lambda| = 2 Re(conj(X) X')
                if (JVP) {
                    pa = fmaf(sm.x, sm.x, df.y * df.y);
                    pb = 2.f * fmaf(sm.x, sm.y, -df.y * df.x);
                } else {
                    const float2 p = px::fma2(px::swap(df), px::swap(df), px::mul2(sm, sm));
                    pa = p.x;
                    pb = p.y;
                }
            };
            if (SPEC) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    float pa, pb;
                    split(z[q], lp[-GT * q], pa, pb);
                    tcol[GT * q * TS] = pa;
                    if (JVP) tcol[GT * q * TS + plane] = pb;
                    else tcol[GT * q * TS + 1] = pb;
                }
                if (gtid == 0) {   // bin L/2 pairs with itself
                    float pa, pb;
                    split(z[8], z[8], pa, pb);
                    tile[(L / 2) * TS + f] = pa;
                    if (JVP) tile2[(L / 2) * TS + f] = pb;
                    else tile[(L / 2) * TS + f + 1] = pb;
                }
            } else if (use_seg) {
                // ---- segment-owner projection (SegHeader, dmel_common.cuh).  The powers of the owned bins are formed in
                // registers first: their slots overlap the partner half that other threads are still reading.
                float2 pw[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) split(z[q], lp[-GT * q], pw[q].x, pw[q].y);
                gsync<GT>(bar_id);
#pragma unroll
                for (int q = 0; q < 8; ++q) buf[pdest[q]] = pw[q];
                if (gtid == 0) {   // bin L/2 pairs with itself
                    float2 ph;
                    split(z[8], z[8], ph.x, ph.y);
                    buf[pdest[8]] = ph;
                }
                gsync<GT>(bar_id);
                // models.py:53 matmul(spec^T, mel_fb), segment by segment: mel[s] takes w.x * P over the bins of segment s,
                // mel[s+1] takes w.y * P.  First the upper parts are stored to their rows, then (one barrier later) every
                // owner adds the lower part to its own row.
                for (int i = 0; i < seg_np; ++i) {
                    const int sg = i * GT + gtid;
                    const int o = s_seg[i] * (GT + 1) + gtid;
                    const float2* pp = buf + o;
                    const float2* wp = segw + o;
                    const int len = s_seg[SEG_MAX_PASSES + i];
                    float da = 0.f, db = 0.f, ua = 0.f, ub = 0.f;
#pragma unroll 2
                    for (int j = 0; j < len; ++j) {
                        const float2 pv = pp[j * (GT + 1)], w = wp[j * (GT + 1)];
                        da = fmaf(w.x, pv.x, da); db = fmaf(w.x, pv.y, db);
                        ua = fmaf(w.y, pv.x, ua); ub = fmaf(w.y, pv.y, ub);
                    }
                    float* ra = tile + sg * TS + f;
                    float* rb = JVP ? ra + plane : ra + 1;
                    if (sg == 0) { *ra = 0.f; *rb = 0.f; }
                    if (sg + 1 < g.M) { ra[TS] = ua; rb[TS] = ub; }
                    gsync<GT>(bar_id);
                    if (sg < g.M) { *ra += da; *rb += db; }
                }
            } else {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    float pa, pb;
                    split(z[q], lp[-GT * q], pa, pb);
                    // P overwrites the lower half of the buffer, which nobody reads in this phase
                    buf[gtid + GT * q] = make_float2(pa, pb);
                }
                if (gtid == 0) {   // bin L/2 pairs with itself
                    float pa, pb;
                    split(z[8], z[8], pa, pb);
                    buf[L / 2] = make_float2(pa, pb);
                }
#ifdef DMEL2_EXP_NOGATHER
                if (false) {
#else
                {
#endif
                gsync<GT>(bar_id);
                // models.py:53 matmul(spec^T, mel_fb): each mel reads its contiguous run of bins, 4 taps per step with
                // one 16-byte weight load; the bins behind a padded tap are valid shared memory (k0+run <= L/2+3 < L)
                for (int m = gtid; m < g.M; m += GT) {
                    const int2 mt = meta[m];
                    const float2* pp = buf + (mt.x & 0xffff);
                    const float4* wp = reinterpret_cast<const float4*>(csr_w + mt.y);
                    const int nch = mt.x >> 16;
                    float sa = 0.f, sb = 0.f;
                    for (int c = 0; c < nch; ++c) {
                        const float4 w = wp[c];
                        const float2 a0 = pp[4 * c], a1 = pp[4 * c + 1], a2 = pp[4 * c + 2], a3 = pp[4 * c + 3];
                        sa = fmaf(w.x, a0.x, sa); sb = fmaf(w.x, a0.y, sb);
                        sa = fmaf(w.y, a1.x, sa); sb = fmaf(w.y, a1.y, sb);
                        sa = fmaf(w.z, a2.x, sa); sb = fmaf(w.z, a2.y, sb);
                        sa = fmaf(w.w, a3.x, sa); sb = fmaf(w.w, a3.y, sb);
                    }
                    tile[m * TS + f] = sa;
                    if (JVP) tile2[m * TS + f] = sb;
                    else tile[m * TS + f + 1] = sb;
                }
                }
            }
        }
        __syncthreads();
        // models.py:54 pack into (B,1,M,T), T fastest; optional fused log (models.py:72-73)
        // (with g.time_major the result is (B,1,T,M), M fastest: what MelPANNsNet transposes to, models.py:162)
        // the per-mel output map (dmel.h: out_scale/out_shift) is compiled as a second copy of the loops so the
        // plain path pays nothing for it
        auto emit = [&](auto aff, int m, int tt, long long off) {
            constexpr bool AFF = decltype(aff)::value;
            float val = tile[m * TS + tt];
            if (JVP) {
                // d log(mel + eps) / d|lambda| = dmel / (mel + eps)
                float dv = tile2[m * TS + tt];
                if (do_log) dv = __fdividef(dv, val + g.log_eps);
                dy[off] = AFF ? dv * __ldg(mel.out_scale + m) : dv;
            }
            // lg2.approx * ln2: <= 3 ulp, i.e. ~1e-6 absolute on log-mel values (budget 1e-5 + 1e-4 |y|)
            if (do_log) val = __logf(val + g.log_eps);
            if (AFF) val = affine_out(mel, m, val);
            y[off] = val;
            ysum += val;
        };
        auto emit_tile = [&](auto aff) {
            if (!g.time_major) {
                // a thread keeps its column tt and walks rows m0, m0+rpp, ...: one pointer bump per output
                constexpr bool AFF = decltype(aff)::value;
                const int tt = tid & (g.TT - 1), m0 = tid >> tt_shift, rpp = NT >> tt_shift;
                if (rpp > 0 && tt < nfr) {
                    const long long off0 = (long long)b * g.M * g.T + t0 + (long long)m0 * g.T + tt;
                    float* yp = y + off0;
                    float* dyp = JVP ? dy + off0 : nullptr;
                    const float* tp = tile + m0 * TS + tt;
                    const size_t ystep = size_t(rpp) * g.T;
                    const int tstep = rpp * TS;
                    const float eps = g.log_eps;
                    for (int m = m0; m < g.M; m += rpp, yp += ystep, tp += tstep) {
                        float val = *tp;
                        if (JVP) {
                            float dv = tp[plane];
                            if (do_log) dv = __fdividef(dv, val + eps);
                            *dyp = AFF ? dv * __ldg(mel.out_scale + m) : dv;
                            dyp += ystep;
                        }
                        if (do_log) val = __logf(val + eps);
                        if (AFF) val = affine_out(mel, m, val);
                        *yp = val;
                        ysum += val;
                    }
                } else if (rpp == 0) {   // tiles wider than the CTA (TT > NT): generic walk
                    const long long base = (long long)b * g.M * g.T + t0;
                    for (int idx = tid; idx < (g.M << tt_shift); idx += NT) {
                        const int m = idx >> tt_shift, t2 = idx & (g.TT - 1);
                        if (t2 < nfr) emit(aff, m, t2, base + (long long)m * g.T + t2);
                    }
                }
            } else {
                const long long base = ((long long)b * g.T + t0) * g.M;
                for (int tt = tid >> 5; tt < nfr; tt += NT / 32)
                    for (int m = tid & 31; m < g.M; m += 32) emit(aff, m, tt, base + (long long)tt * g.M + m);
            }
        };
#ifndef DMEL2_EXP_NOEPI
        if (affine) emit_tile(BoolTag<true>{});
        else emit_tile(BoolTag<false>{});
#endif
        // the next iteration's staging only writes xs; its __syncthreads orders it after this epilogue
    }
    cp_async_wait_all();
    // fused by-product: sum of all outputs (what train.py:58 copies the whole batch to the host for)
    const double ytot = block_sum(double(ysum), reinterpret_cast<double*>(smem_f + lay.red));
    if (tid == 0) atomicAdd(&hdr->y_sum, ytot);
}

// ----------------------------------------------------------------------------
// backward
// ----------------------------------------------------------------------------
#ifdef DMEL2_BWD_THREADS   // experiment: resident threads per SM the backward is compiled for (register budget)
#define DMEL2_BWD_MINB(NT) ((DMEL2_BWD_THREADS) / (NT) > 0 ? (DMEL2_BWD_THREADS) / (NT) : 1)
#else
#define DMEL2_BWD_MINB(NT) min_ctas2(NT, false)
#endif
template <int L, int NG, bool SPEC, bool NEED_DX>
// n_fft 1024 with dL/dx in mel mode: 57-75 KB of shared memory per CTA (two groups, their dx sub-spans, the dL/dy and y tiles)
// limit the SM to 3 CTAs whatever the registers, so the kernel is compiled for 3 (up to 168 registers, it takes 160) instead
// of 4 (128): backward -4 % on large_batch (profiles/r02/r03w_bwd1024_registers.log)
// n_fft 4096 (256 threads, 1-2 CTAs per SM by shared memory): compiled without a register cap (-2.6 %, r03y)
__global__ void __launch_bounds__(NG * (L / 16), (L == 1024 && NG == 2 && NEED_DX && !SPEC) ? 3 : (L == 4096 ? 1 : DMEL2_BWD_MINB(NG * (L / 16))))
dmel2_bwd_kernel(const void* __restrict__ x, const float* __restrict__ y, const float* __restrict__ gy,
                 float* __restrict__ dx, Geom g, MelDev mel, const float* __restrict__ win_g,
                 const float2* __restrict__ tw_g, const float* __restrict__ means, WsHeader* hdr,
                 double* __restrict__ clipsum) {
    using F = Fft<L>;
    constexpr int GT = F::GT, NT = NG * GT;
    constexpr int NF = L / 2 + 1;
    extern __shared__ __align__(16) float smem_f[];
    // dL/dmel tile: M+1 rows (the extra row stays zero: a bin's second tap may point one past the last mel) of
    // TS = TT+2 words, so that the two frames of a pair are one aligned 8-byte load
    const int TS = g.TT + 2;
    const int tile_words = round_up4((g.M + 1) * TS);
    const Lay2 lay = make_lay2(F::TW_SMEM, F::BUF, NG, g.span, g.sub_span, (2 + 2 * g.lookahead) * tile_words,
                               SPEC ? 0 : 4 * NF, 0, NEED_DX, 1 + g.lookahead);
    float2* tws = reinterpret_cast<float2*>(smem_f + lay.tw);
    float* xs_sh = smem_f + lay.xs;
    float* gt = smem_f + lay.tile;
    int4* meta = reinterpret_cast<int4*>(smem_f + lay.meta);
    double* red = reinterpret_cast<double*>(smem_f + lay.red);

    const int tid = threadIdx.x;
    const int grp = tid / GT, gtid = tid - grp * GT, bar_id = group_sync_arg<GT>(grp);
    float2* buf = reinterpret_cast<float2*>(smem_f + lay.grp + grp * lay.grp_stride);
    float* priv = smem_f + lay.grp + grp * lay.grp_stride + round_up4(2 * F::BUF);   // this group's dx sub-span
    const int tt_shift = ilog2(g.TT);
    const bool do_log = g.log_eps >= 0.f;
    const bool affine = !SPEC && mel.out_scale != nullptr;

    {
        const float2* src = tw_g + (F::TW1_REGS ? F::TW1_N : 0);
        stage_table_async<NT>(smem_f + lay.tw, src, 2 * F::TW_SMEM, tid);
    }
    if (!SPEC) {
        // bin-major bank, one 16-byte record per bin: {row offset of the first mel in the tile, w0, w1, extra taps}.
        // The Hermitian extension counts bins 0 and L/2 twice (adjoint of the one-sided transform): folded in here.
        for (int k = tid; k < NF; k += NT) {
            const int o0 = __ldg(mel.bin_off + k), cnt = __ldg(mel.bin_off + k + 1) - o0;
            const float e = (k == 0 || k == L / 2) ? 2.f : 1.f;
            const float w0 = cnt > 0 ? e * __ldg(mel.bin_w + o0) : 0.f;
            const float w1 = cnt > 1 ? e * __ldg(mel.bin_w + o0 + 1) : 0.f;
            const int m0 = cnt > 0 ? __ldg(mel.bin_m0 + k) : 0;
            meta[k] = make_int4(m0 * TS, __float_as_int(w0), __float_as_int(w1), cnt > 2 ? cnt - 2 : 0);
        }
    }
    // zero rows (and pad columns) of the gradient tiles once: the staging below only writes [0,M) x [0,TT)
    for (int i = tid; i < (2 + 2 * g.lookahead) * tile_words; i += NT) gt[i] = 0.f;
    typename F::Ctx ctx;
    F::init(ctx, buf, tws, tw_g, gtid, bar_id);
    float wh[16], acc[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        wh[q] = 0.5f * __ldg(win_g + gtid + GT * q);
        acc[q] = 0.f;
    }
    __syncthreads();   // tile zeroing above precedes the first asynchronous copies into it

    // Persistent tile loop with a one-tile look-ahead (see the forward kernel): the span and the dL/dy / y
    // tiles of tile i+1 land in the alternate buffers while tile i is processed.
    const int total_tiles = g.B * g.tiles_per_clip;
    const int xs_words = round_up4(g.span + 4);
    auto issue_tile = [&](int tix_, int slot) {
        const int b_ = tix_ / g.tiles_per_clip;
        const int t0_ = (tix_ - b_ * g.tiles_per_clip) * g.TT;
        const int nfr_ = min(g.TT, g.T - t0_);
        span_begin<NT>(xs_sh + slot * xs_words, x, b_, g, t0_ * g.hop - L / 2, tid);
        // upstream gradient tile (and y when the log is fused): rows are T floats apart in HBM, so the copies
        // are 4-byte async ones, all in flight together; columns past the last frame are zero
        float* gt_ = gt + slot * 2 * tile_words;
        float* yt_ = gt_ + tile_words;
        if (!g.time_major) {
            const long long base = (long long)b_ * g.M * g.T + t0_;
            for (int idx = tid; idx < (g.M << tt_shift); idx += NT) {
                const int m = idx >> tt_shift, tt = idx & (g.TT - 1);   // consecutive lanes: consecutive frames
                const bool ok = tt < nfr_;
                const long long off = base + (long long)m * g.T + (ok ? tt : 0);
                cp_async4_zfill(gt_ + m * TS + tt, gy + off, ok);
                if (do_log) cp_async4_zfill(yt_ + m * TS + tt, y + off, ok);
            }
        } else {
            // (B,1,T,M) layout: consecutive lanes read consecutive mel bins of one frame
            const long long base = ((long long)b_ * g.T + t0_) * g.M;
            for (int tt = tid >> 5; tt < g.TT; tt += NT / 32) {
                const bool ok = tt < nfr_;
                for (int m = tid & 31; m < g.M; m += 32) {
                    const long long off = base + (long long)(ok ? tt : 0) * g.M + m;
                    cp_async4_zfill(gt_ + m * TS + tt, gy + off, ok);
                    if (do_log) cp_async4_zfill(yt_ + m * TS + tt, y + off, ok);
                }
            }
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
        if (NEED_DX) {
            float4* p4 = reinterpret_cast<float4*>(priv);
            const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int i = gtid; i < (round_up4(g.sub_span) >> 2); i += GT) p4[i] = zero4;
        }
        cp_async_wait_all();
        span_finish<NT>(xs_cur, x, b, g, p0, mu, tid);
        if (do_log || affine) {
            // dL/dmel = G / (mel + eps) = G * exp(-y); each thread touches the words it copied itself
            auto to_dmel = [&](auto aff) {
                constexpr bool AFF = decltype(aff)::value;
                auto one = [&](int m, int tt) {
                    const int i = m * TS + tt;
                    gtc[i] *= AFF ? affine_grad_factor(mel, m, do_log ? ytc[i] : 0.f, do_log) : __expf(-ytc[i]);
                };
                if (!g.time_major) {
                    for (int idx = tid; idx < (g.M << tt_shift); idx += NT) one(idx >> tt_shift, idx & (g.TT - 1));
                } else {
                    for (int tt = tid >> 5; tt < g.TT; tt += NT / 32)
                        for (int m = tid & 31; m < g.M; m += 32) one(m, tt);
                }
            };
            if (affine) to_dmel(BoolTag<true>{});
            else to_dmel(BoolTag<false>{});
        }
        __syncthreads();   // also: every warp is done with the previous tile
        if (ahead && tix + int(gridDim.x) < total_tiles) issue_tile(tix + gridDim.x, slot ^ 1);

        const int sfe = tile_frames_per_group<NG, true>(nfr, g.TT, g.SF);   // short last tile: even re-split
        const int f_lo = grp * sfe;
        const int f_hi = min(f_lo + sfe, nfr);
        for (int f = f_lo; f < f_hi; f += 2) {
            const float* xa = xs + f * g.hop + gtid;
            const float* xb = xa + g.hop;
            const bool has_b = (f + 1) < nfr;
            float2 z[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const float a = xa[GT * q], c = xb[GT * q];
                z[q] = px::mul2(make_float2(a, c), px::bc(wh[q]));
            }
            // the forward and the inverse transform share ONE copy of the FFT code (rolled two-trip loop): the
            // unrolled butterflies are most of the kernel's instruction footprint
#ifndef DMEL2_BWD_UNROLLED
            int phase = 0;
#pragma unroll 1
            do {
            F::run(z, ctx);
            asm volatile("" : "+r"(phase));   // opaque to the optimiser: keeps the two-trip loop rolled
            if (phase == 0) {
#else
            {
            F::run(z, ctx);
#endif
            // conj(V) for the inverse transform, V = Zhat_a + i*Zhat_b the Hermitian-extended GP*X of both frames:
            //   conj(V[k]) = s*conj(Z[k]) + d*Z[L-k],  s = ga+gb, d = ga-gb  (window pre-scaled by 1/2)
            // The owner of bin k < L/2 reads Z[L-k] from the partner's slot, forms conj(V[k]) (kept in registers) and
            // conj(V[L-k]) (written back to the same slot, which the partner then reads as its own bin).
            gsync<GT>(bar_id);
#pragma unroll
            for (int q = 8; q < 16; ++q) buf[gtid + GT * q] = z[q];
            if (gtid == 0) buf[L] = z[0];
            gsync<GT>(bar_id);
            float2* lp = buf + L - gtid;
            const float* gcol = gtc + f;
            auto bin_grad = [&](int k, float& ga, float& gb) {
                if (SPEC) {
                    const float2 gg = *reinterpret_cast<const float2*>(gcol + k * TS);
                    const float e = (k == 0 || k == L / 2) ? 2.f : 1.f;
                    ga = e * gg.x;
                    gb = e * gg.y;
                } else {
                    const int4 mt = meta[k];
                    const float* r0 = gcol + mt.x;
                    const float2 g0 = *reinterpret_cast<const float2*>(r0);
                    const float2 g1 = *reinterpret_cast<const float2*>(r0 + TS);
                    const float w0 = __int_as_float(mt.y), w1 = __int_as_float(mt.z);
                    const float2 gab = px::fma2(px::bc(w1), g1, px::mul2(px::bc(w0), g0));
                    ga = gab.x;
                    gb = gab.y;
                    if (mt.w) {   // banks with more than two taps per bin (not the triangular HTK bank): slow path
                        const int o0 = __ldg(mel.bin_off + k);
                        const float e = (k == 0 || k == L / 2) ? 2.f : 1.f;
                        for (int t = 0; t < mt.w; ++t) {
                            const float w = e * __ldg(mel.bin_w + o0 + 2 + t);
                            const float2 gg = *reinterpret_cast<const float2*>(r0 + (2 + t) * TS);
                            ga = fmaf(w, gg.x, ga);
                            gb = fmaf(w, gg.y, gb);
                        }
                    }
                }
            };
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                float ga, gb;
                bin_grad(gtid + GT * q, ga, gb);
                const float s = ga + gb, d = ga - gb;
                const float2 z1 = z[q], z2 = lp[-GT * q];
                const float2 sc = make_float2(s, -s);
                z[q] = px::fma2(sc, z1, px::mul2(px::bc(d), z2));
                lp[-GT * q] = px::fma2(sc, z2, px::mul2(px::bc(d), z1));
            }
            if (gtid == 0) {   // bin L/2 pairs with itself
                float ga, gb;
                bin_grad(L / 2, ga, gb);
                const float s = ga + gb, d = ga - gb;
                buf[L / 2] = make_float2((s + d) * z[8].x, (d - s) * z[8].y);
            }
            gsync<GT>(bar_id);
#pragma unroll
            for (int q = 8; q < 16; ++q) z[q] = buf[gtid + GT * q];
            // inverse transform as conj(FFT(conj(V))): u_a = Re, u_b = -Im of the FFT output
#ifdef DMEL2_BWD_UNROLLED
            F::run(z, ctx);
            }
#else
            }
            } while (++phase < 2);
#endif
            if (!has_b) {
#pragma unroll
                for (int q = 0; q < 16; ++q) z[q].y = 0.f;
            }
            // dL/dlambda, time-domain form: sum_n dw/dlambda[n] * sum_frames u[n]*fr[n]; the frame sum per owned n
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                acc[q] = fmaf(z[q].x, xa[GT * q], acc[q]);
                acc[q] = fmaf(-z[q].y, xb[GT * q], acc[q]);
            }
            if (NEED_DX) {
                // overlap-add (adjoint of the framing) of u*w/2 into the group's private span: frame a, barrier,
                // frame b — no shared-memory atomics.  (The barrier that orders frame b before the next pair's
                // frame a is the first one of the next transform.)
                float* da = priv + (f - f_lo) * g.hop + gtid;
#pragma unroll
                for (int q = 0; q < 16; ++q) da[GT * q] = fmaf(z[q].x, wh[q], da[GT * q]);
                gsync<GT>(bar_id);
                if (has_b) {
                    float* db = da + g.hop;
#pragma unroll
                    for (int q = 0; q < 16; ++q) db[GT * q] = fmaf(-z[q].y, wh[q], db[GT * q]);
                }
            }
        }

        if (NEED_DX) {
            __syncthreads();   // all private spans of this tile are complete
            // merge the groups' sub-spans (x2: the window was pre-scaled by 1/2); crop the centre padding; positions no
            // other tile touches are stored, the halos shared with the neighbouring tiles are reduced into the zeroed dx
            const int span_used = (nfr - 1) * g.hop + L;
            const int excl_lo = (tile_i == 0) ? 0 : (t0 - 1) * g.hop + L / 2;
            const int excl_hi = (tile_i == g.tiles_per_clip - 1) ? g.N : (t0 + nfr) * g.hop - L / 2;
            float* dxb = dx + (long long)b * g.N;
            const float* priv0 = smem_f + lay.grp + round_up4(2 * F::BUF);
            const int gstep = sfe * g.hop;
            float part = 0.f;
            if (g.dx_vec4 && (((p0 | gstep) & 3) == 0)) {
                // dx rows are 16-byte aligned and this tile's first sample and the groups' offsets are multiples of 4: the merge
                // moves 16 bytes per step (LDS.128 per group, one STG.128 or one vector reduction) instead of four scalar round
                // trips.  A sub-span is zero-padded to a multiple of 4, so a quad may run past its end; only the quads that
                // straddle the store / reduce boundaries take the per-sample path.
                const unsigned ss4 = unsigned(round_up4(g.sub_span));
                for (int i = 4 * tid; i < span_used; i += 4 * NT) {
                    const int p = p0 + i;
                    if (p >= 0 && p < g.N) {
                        float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int gi = 0; gi < NG; ++gi) {
                            const int o = i - gi * gstep;
                            if (unsigned(o) < ss4) {
                                const float4 v = *reinterpret_cast<const float4*>(priv0 + gi * lay.grp_stride + o);
                                val.x += v.x; val.y += v.y; val.z += v.z; val.w += v.w;
                            }
                        }
                        val.x *= 2.f; val.y *= 2.f; val.z *= 2.f; val.w *= 2.f;
                        part += (val.x + val.y) + (val.z + val.w);
                        if (p >= excl_lo && p + 3 < excl_hi) {
                            *reinterpret_cast<float4*>(dxb + p) = val;
                        } else if (p + 3 < excl_lo || p >= excl_hi) {
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dxb + p), "f"(val.x), "f"(val.y),
                                         "f"(val.z), "f"(val.w)
                                         : "memory");
                        } else {
                            const float v4[4] = {val.x, val.y, val.z, val.w};
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                if (p + j >= excl_lo && p + j < excl_hi) dxb[p + j] = v4[j];
                                else atomicAdd(dxb + p + j, v4[j]);
                            }
                        }
                    }
                }
            } else
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
                    val *= 2.f;
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
    // dw/d|lambda|[n] = w[n] * ((n-centre)^2/lambda^3 - cprime),  w = 2*wh
    float dlam = 0.f;
    {
        const float inv_lam3 = hdr->inv_lam3, cprime = hdr->cprime, centre = hdr->centre;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const float d = float(gtid + GT * q) - centre;
            dlam = fmaf(acc[q] * (2.f * wh[q]), fmaf(d * d, inv_lam3, -cprime), dlam);
        }
    }
    const double tot = block_sum(double(dlam), red);
    if (tid == 0) atomicAdd(&hdr->dlam_acc, tot);
}

}  // namespace v2
}  // namespace dmel
