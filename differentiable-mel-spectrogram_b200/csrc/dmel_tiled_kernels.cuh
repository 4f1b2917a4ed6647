// dmel_tiled_kernels.cuh — persistent tiled forward / backward kernels (n_fft 128..4096)
// Part of libdmel_b200.so (single translation unit, see dmel_kernels.cu for the overview).
#pragma once
#include "dmel_common.cuh"

namespace dmel {

// ----------------------------------------------------------------------------
// forward (persistent: each CTA stages the constant tables once and walks tiles)
//
// JVP variant (the reference's training case: x is data, only lambd trains — models.py:19, main.py:27): instead of
// two frames, one complex transform carries ONE frame windowed by w (real part) and by dw/d|lambda| (imaginary
// part), i.e. X and X' = STFT(x * dw/dlambda); besides y it writes dy/d|lambda| = fb . 2Re(conj(X) X') (chain-ruled
// through the fused log), so the backward collapses to the dot product sum(dL/dy * dy/dlambda): 1.0 instead of 1.5
// complex transforms per frame for forward + backward, and nothing is recomputed.
// ----------------------------------------------------------------------------
template <int L, int NG, bool SPEC, bool JVP>
__global__ void __launch_bounds__(NG * (L / Plan<L>::E), min_ctas(L, NG, NG * (L / Plan<L>::E)))
dmel_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ dy, Geom g, MelDev mel,
                const float* __restrict__ win_g, const float2* __restrict__ tw_g, const float* __restrict__ means,
                WsHeader* hdr) {
    constexpr int E = Plan<L>::E, GT = L / E, NT = NG * GT, TWN = TwTotal<L>::value;
    using P0 = Pass<L, 0>;
    extern __shared__ __align__(16) float smem_f[];
    const int TS = g.TT + 1;
    const int plane = round_up4(g.M * TS);   // JVP keeps two planes: mel and d(mel)/d|lambda|
    const SmemLay lay = make_lay(L, NG, TWN, g.span, g.sub_span, JVP ? 2 * plane : g.M * TS, g.csr_ints, g.csr_floats, false,
                                 1 + g.lookahead);
    float2* twt = reinterpret_cast<float2*>(smem_f + lay.twt);
    float* win = smem_f + lay.win;
    float* xs_sh = smem_f + lay.xs;
    float* tile = smem_f + lay.tile;
    float* tile2 = tile + plane;
    int* csr_i = reinterpret_cast<int*>(smem_f + lay.csr_i);
    float* csr_w = smem_f + lay.csr_w;

    const int tid = threadIdx.x;
    const int grp = tid / GT, gtid = tid - grp * GT, bar_id = 1 + grp;
    float2* buf = reinterpret_cast<float2*>(smem_f + lay.grp + grp * lay.grp_stride);
    const bool row_aligned = g.row_aligned != 0;
    const int tt_shift = ilog2(g.TT);
    const bool do_log = g.log_eps >= 0.f;
    const bool affine = !SPEC && mel.out_scale != nullptr;
    float inv_lam3 = 0.f, cprime = 0.f, centre = 0.f;
    if (JVP) { inv_lam3 = hdr->inv_lam3; cprime = hdr->cprime; centre = hdr->centre; }

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

        // frames per group: a short last tile of a clip is re-split evenly over the groups (whole pairs)
        const int sfe = tile_frames_per_group<NG, !JVP>(nfr, g.TT, g.SF);
        const int f_hi = min(grp * sfe + sfe, nfr);
        for (int f = grp * sfe; f < f_hi; f += (JVP ? 1 : 2)) {
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
                    if (JVP) {
                        // dw/d|lambda|[n] = w[n] * ((n-centre)^2/lambda^3 - cprime)
                        const float d = float(n) - centre;
                        const float xw = xa[n] * w;
                        v[i * P0::R + r] = make_float2(xw, xw * fmaf(d * d, inv_lam3, -cprime));
                    } else {
                        v[i * P0::R + r] = make_float2(xa[n] * w, xb2[n] * w);
                    }
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
                // pair: both powers.  JVP: (ar,ai) = X, (br,bi) = X' -> P = |X|^2, dP/d|lambda| = 2 Re(conj(X) X')
                const float pa = fmaf(ar, ar, ai * ai);
                const float pb = JVP ? 2.f * fmaf(ar, br, ai * bi) : fmaf(br, br, bi * bi);
                if (SPEC) {
                    tile[k * TS + f] = pa;
                    if (JVP) tile2[k * TS + f] = pb;
                    else tile[k * TS + f + 1] = pb;
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
                    if (JVP) tile2[m * TS + f] = sb;
                    else tile[m * TS + f + 1] = sb;
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
                const long long base = (long long)b * g.M * g.T + t0;
                for (int idx = tid; idx < (g.M << tt_shift); idx += NT) {
                    const int m = idx >> tt_shift, tt = idx & (g.TT - 1);
                    if (tt < nfr) emit(aff, m, tt, base + (long long)m * g.T + tt);
                }
            } else {
                const long long base = ((long long)b * g.T + t0) * g.M;
                for (int tt = tid >> 5; tt < nfr; tt += NT / 32)
                    for (int m = tid & 31; m < g.M; m += 32) emit(aff, m, tt, base + (long long)tt * g.M + m);
            }
        };
        if (affine) emit_tile(BoolTag<true>{});
        else emit_tile(BoolTag<false>{});
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
    const bool affine = !SPEC && mel.out_scale != nullptr;

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
        if (NEED_DX)
            for (int i = gtid; i < g.sub_span; i += GT) priv[i] = 0.f;
        cp_async_wait_all();
        span_centre<NT>(xs_cur, p0, g.span, g.N, mu, tid);
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
            const int gstep = sfe * g.hop;
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

}  // namespace dmel
