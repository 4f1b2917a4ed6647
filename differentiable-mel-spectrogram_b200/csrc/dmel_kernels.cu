// dmel_kernels.cu — fused DMEL forward / backward kernels for sm_100a and their C ABI.  This file holds the host-side
// planner, the dispatch and the extern "C" entry points; the tiled kernels are instantiated in inst_tiled.cu (one
// object per transform length and direction, built in parallel); the device code is in
//   fft_device.cuh          in-register / shared-memory FFT building blocks
//   dmel_common.cuh         workspace + geometry structs, reductions, shared-memory layout, cp.async staging
//   dmel_small_kernels.cuh  prep, zero-init, finalize
//   dmel_tiled_kernels.cuh  persistent tiled forward / backward (n_fft 128..4096)
//   dmel_big_kernels.cuh    n_fft 8192 / 16384
//   dmel_dft_kernels.cuh    direct-DFT correctness path).
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
#include <string.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <vector>

#include "../../include/dmel.h"
#include "fft_device.cuh"

#include "dmel_common.cuh"
#include "dmel_host.cuh"
#include "dmel_small_kernels.cuh"
#include "dmel_tiled_kernels.cuh"   // shared-memory layouts and plan constants for the planner; the tiled kernels
#include "dmel_tiled2_kernels.cuh"  // themselves are instantiated in inst_tiled.cu
#include "dmel_big_kernels.cuh"
#include "dmel_dft_kernels.cuh"
#include "dmel_dftmm_kernels.cuh"

namespace dmel {

// ----------------------------------------------------------------------------
// host side: planning + dispatch
// ----------------------------------------------------------------------------
constexpr size_t SMEM_MAX = 227 * 1024;
// Peer all-reduce spin bound.  A collective has to wait for its slowest rank however late it is (checkpointing or
// evaluation on one rank, a data-loader stall, first-iteration set-up): a short bound would hand NaN to the ranks
// that gave up while the late rank still completes, and the replicated lambda would diverge.  So the default is the
// order of NCCL's own watchdog (30 minutes at ~2 GHz); DMEL_PEER_TIMEOUT_S overrides it (tests use a small value).
static long long peer_timeout_cycles() {
    static const long long cycles = [] {
        double s = 1800.0;
        if (const char* e = getenv("DMEL_PEER_TIMEOUT_S")) {
            const double v = atof(e);
            if (v > 0.0) s = v;
        }
        return (long long)(s * 2.0e9);
    }();
    return cycles;
}
#define PEER_TIMEOUT_CYCLES peer_timeout_cycles()

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

// second-generation tiled kernels (fft_core.cuh / dmel_tiled2_kernels.cuh): n_fft 256..4096, 16 points per thread.
// (the first generation, dmel_tiled_kernels.cuh, still serves n_fft 128)
template <int L> static PlanDesc plan2_desc_of() {
    PlanDesc d;
    d.np = v2::Fft<L>::NP;
    d.r[0] = 16; d.r[1] = v2::Fft<L>::R1; d.r[2] = v2::Fft<L>::R2; d.r[3] = 1;
    return d;
}
struct Fft2Info {
    bool ok;
    int gt, tw_smem, buf;
    int ng[2];
    PlanDesc desc;
};
static Fft2Info fft2_info(int n_fft) {
    Fft2Info f{};
    switch (n_fft) {
#define DMEL_FFT2_CASE(LL, NGA, NGB)                                                                             \
    case LL: f.ok = true; f.gt = v2::Fft<LL>::GT; f.tw_smem = v2::Fft<LL>::TW_SMEM; f.buf = v2::Fft<LL>::BUF;      \
             f.ng[0] = NGA; f.ng[1] = NGB; f.desc = plan2_desc_of<LL>(); break;
        DMEL_FFT2_CASE(256, 8, 4)
        DMEL_FFT2_CASE(512, 8, 4)
        DMEL_FFT2_CASE(1024, 4, 2)
        DMEL_FFT2_CASE(2048, 4, 1)
        DMEL_FFT2_CASE(4096, 1, 1)
#undef DMEL_FFT2_CASE
        default: break;
    }
    return f;
}


// second generation: forward keeps an [M][TT+1] tile (two planes for the JVP forward) and int2-per-mel records + the
// mel-major weights; backward keeps (M+1) x (TT+2) tiles of dL/dy and y (doubled by the look-ahead) and one 16-byte
// record per bin
static size_t smem_bytes_fft2(const Fft2Info& fi, int NG, int span, int sub_span, int M, int F, int TT, int mel_nnz,
                              bool spec, bool bwd, bool priv_dx, int lookahead, bool jvp, bool seg) {
    const int n_fft = 2 * (F - 1);
    int tile_words, meta_words, w_words;
    if (bwd) {
        tile_words = (2 + 2 * lookahead) * round_up4((M + 1) * (TT + 2));
        meta_words = spec ? 0 : 4 * F;
        w_words = 0;
    } else {
        tile_words = jvp ? 2 * round_up4(M * (TT + 1)) : M * (TT + 1);
        meta_words = spec ? 0 : 2 * M;
        w_words = spec ? 0 : (seg ? v2::fwd_weight_words(n_fft, mel_nnz) : mel_nnz);
    }
    return size_t(v2::make_lay2(fi.tw_smem, fi.buf, NG, span, sub_span, tile_words, meta_words, w_words, priv_dx,
                                1 + lookahead).total) * 4;
}

static size_t smem_bytes_fft(int L, int NG, int tw_total, int span, int sub_span, int M, int TT, int csr_ints,
                             int csr_floats, bool bwd, bool priv_dx, int lookahead, bool jvp = false) {
    // backward keeps two tiles (dL/dy and, for the fused log, y); the look-ahead doubles them and the span.
    // the JVP forward keeps two planes (mel, dmel/dlambda)
    const int tile_words = bwd ? (2 + 2 * lookahead) * round_up4(M * (TT + 1))
                               : (jvp ? 2 * round_up4(M * (TT + 1)) : M * (TT + 1));
    return size_t(make_lay(L, NG, tw_total, span, sub_span, tile_words, csr_ints, csr_floats, priv_dx, 1 + lookahead).total) * 4;
}

// Tile shape: NG groups x SF frames per CTA tile.  scripts/tune.py sweeps on B200 (profiles/) show a
// flat optimum around 8-16 frames per tile: smaller tiles pay the per-tile fixed cost (span staging,
// barriers, tile IO) too often, larger ones cost occupancy and leave a coarse tail over the 148 SMs.
// The preferred (groups, frames per group) per transform length is tabulated below and shrunk
// until the tile fits shared memory or the clip.
// DMEL_NG / DMEL_SF in the environment pin the choice (tuning experiments).
static int plan_launch(int64_t B, int64_t N, int n_fft, int hop, int rows, int csr_ints, int csr_floats, bool bwd,
                       bool need_dx, LaunchPlan* lp, bool jvp = false, bool seg = false) {
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
    const Fft2Info f2 = fft2_info(n_fft);
    const bool spec_mode = csr_floats == 0;
    lp->gen = f2.ok ? 2 : 1;
    int pin_ng = 0, pin_sf = 0;
    if (const char* e = getenv("DMEL_NG")) pin_ng = atoi(e);
    if (const char* e = getenv("DMEL_SF")) pin_sf = atoi(e);
    // measured optima (scripts/tune.py on B200, profiles/*tune*.log): (groups, frames per group)
    int ng_pref, sf_pref;
    switch (n_fft) {
        case 1024:
            if (f2.ok) { ng_pref = 2; sf_pref = (bwd && !need_dx) ? 4 : 8; }   // r02h sweep (profiles/)
            else { ng_pref = (bwd && need_dx) ? 4 : 2; sf_pref = (bwd && need_dx) ? 4 : 8; }
            break;
        case 2048:
            // 128 threads per transform: one group per CTA leaves 8-12 warps per SM (shared memory bounds the CTAs), too few
            // to cover the exchange latency; four groups share one copy of the tables (r02o, profiles/r02: backward
            // -9 %, lambda-only backward -10 %, forward -3 %; the forward-mode forward keeps one group)
            if (jvp) { ng_pref = 1; sf_pref = 8; }
            else { ng_pref = 4; sf_pref = (bwd && need_dx) ? 4 : 8; }
            break;
        case 4096: ng_pref = 1; sf_pref = 32; break;
        case 256: ng_pref = f2.ok ? 8 : 4; sf_pref = f2.ok ? 2 : 4; break;   // gen 2: half a warp per transform
        default: ng_pref = 4; sf_pref = 4; break;    // one warp per transform
    }
    const int gt = f2.ok ? f2.gt : fi.gt;
    const int* ngs = f2.ok ? f2.ng : fi.ng;
    for (int attempt = 0; attempt < 2; ++attempt) {
        for (int ni = 0; ni < 2; ++ni) {
            // first the preferred NG, then the other instantiated one
            const int NG = (ni == 0) ? ng_pref : (ngs[0] == ng_pref ? ngs[1] : ngs[0]);
            if (ni == 1 && NG == ng_pref) continue;
            if (pin_ng && NG != pin_ng) continue;
            int sf0 = pin_sf ? pin_sf : sf_pref;
            // no point in a tile much longer than the clip
            while (!pin_sf && sf0 > 2 && NG * (sf0 / 2) >= T) sf0 >>= 1;
            for (int SF = sf0; SF >= 2; SF >>= 1) {
                const int tt = NG * SF;
                const int64_t span64 = int64_t(tt - 1) * hop + n_fft, sub64 = int64_t(SF - 1) * hop + n_fft;
                if (span64 <= (1 << 20)) {
                    auto smem_of = [&](int la) {
                        return f2.ok ? smem_bytes_fft2(f2, NG, int(span64), int(sub64), rows, n_fft / 2 + 1, tt, csr_floats,
                                                       spec_mode, bwd, bwd && need_dx, la, jvp, seg)
                                     : smem_bytes_fft(n_fft, NG, fi.tw_total, int(span64), int(sub64), rows, tt, csr_ints,
                                                      csr_floats, bwd, bwd && need_dx, la, jvp);
                    };
                    size_t sm = smem_of(0);
                    if (sm <= SMEM_MAX) {
                        // one-tile look-ahead (double buffers) only if it does not cost a resident CTA
                        const size_t sm2 = smem_of(1);
                        const int warps_cta = std::max(1, NG * gt / 32);
                        const bool cap96 = f2.ok ? (n_fft <= 512 && !bwd && !jvp) : (n_fft <= 512 && NG == 4 && !(bwd && need_dx));
                        int reg_cap = std::max(1, (cap96 ? 20 : 16) / warps_cta);
                        // dmel2_bwd_kernel<1024,2,mel,dx> is compiled for 3 resident CTAs (160 registers, see its launch bounds)
                        if (f2.ok && n_fft == 1024 && NG == 2 && bwd && need_dx && !spec_mode) reg_cap = 3;
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
                        lp->threads = NG * gt;
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

// launch instrumentation state (declared in dmel_host.cuh, shared with the inst_tiled.cu translation units)
std::atomic<long long> g_launches[PROF_KINDS];
thread_local ProfState t_prof;

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

static int validate(const void* x, int64_t B, int64_t N, int64_t stride, int32_t win, int32_t n_fft, int32_t hop,
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
        if ((bank->out_scale == nullptr) != (bank->out_shift == nullptr)) return DMEL_ERR_BAD_MELBANK;
    }
    if (ws_bytes_in < ws_bytes(B, n_fft)) return DMEL_ERR_WORKSPACE_TOO_SMALL;
    return DMEL_OK;
}

static MelDev mel_dev(const dmel_melbank* bank) {
    MelDev m{};
    if (bank) {
        m.mel_k0 = bank->mel_k0; m.mel_off = bank->mel_off; m.mel_w = bank->mel_w;
        m.bin_m0 = bank->bin_m0; m.bin_off = bank->bin_off; m.bin_w = bank->bin_w;
        m.out_scale = bank->out_scale; m.out_shift = bank->out_shift;
        m.seg = seg_view(const_cast<void*>(bank->seg_tables), 2 * (bank->n_freqs - 1));
    }
    return m;
}

static int run_prep(const void* x, int64_t B, int64_t N, int64_t stride, float lam_abs, int32_t win, int32_t n_fft,
                    int32_t flags, const WsView& ws, cudaStream_t st, bool gen2, const float* means_src = nullptr) {
    ProfScope prof(PROF_PREP, st);
    const int nseg = prep_segments(int(N));
    if ((B + 1) * int64_t(nseg) > INT32_MAX) return DMEL_ERR_BAD_GEOMETRY;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(unsigned((B + 1) * nseg));
    cfg.blockDim = dim3(PREP_THREADS);
    cfg.stream = st;
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = unsigned(nseg);
    attr.val.clusterDim.y = attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, dmel_prep_kernel, x, (long long)stride, int(B), int(N), lam_abs, int(win), int(n_fft),
                       int(flags), ws, gen2 ? fft2_info(n_fft).desc : fft_info(n_fft).desc, means_src);
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
    g.time_major = 0;
    g.x_dtype = DMEL_X_F32;
    return g;
}

// sample type of x from the call flags; 16-bit inputs need the second-generation tiled kernels
static int x_dtype_of(int32_t flags, const LaunchPlan& lp, int* dtype) {
    const int both = DMEL_FLAG_X_I16 | DMEL_FLAG_X_F16;
    if ((flags & both) == both) return DMEL_ERR_BAD_GEOMETRY;
    *dtype = (flags & DMEL_FLAG_X_I16) ? DMEL_X_I16 : ((flags & DMEL_FLAG_X_F16) ? DMEL_X_F16 : DMEL_X_F32);
    if (*dtype != DMEL_X_F32 && !(lp.path == 0 && lp.gen == 2)) return DMEL_ERR_UNSUPPORTED_NFFT;
    return DMEL_OK;
}

// spans move as 16-byte quads (fp32) / 8-byte quads (16-bit samples) when x and its row stride allow
static int row_alignment_ok(const void* x, int64_t stride, int dtype) {
    const uintptr_t mask = dtype == DMEL_X_F32 ? 15 : 7;
    return ((reinterpret_cast<uintptr_t>(x) & mask) == 0 && (stride & 3) == 0) ? 1 : 0;
}

// Short windows (even n_fft < 128), plain forward: DFT by matrix multiply on the tensor cores (dmel_dftmm_kernels.cuh)
// instead of the O(L^2) CUDA-core DFT: 5.6x (n_fft 64, batch 512) to 6.7x (batch 4096) faster.  n_fft 128 itself stays
// on the shared-memory FFT, which is 1.5x faster than the tensor-core kernel at every batch size measured
// (profiles/r02/r02z_dftmm_vs_cuda_cores.log; the subtractive builds next to it show why: the operand builders' loads and
// the single-buffered build -> MMA latency chain, not the MMAs, bound it).  DMEL_NO_DFTMM=1 keeps the CUDA-core kernels, DMEL_DFTMM_128=1 also sends n_fft 128 to the tensor
// cores (A/B measurements, tests).
static bool dftmm_enabled() {
    static const bool off = getenv("DMEL_NO_DFTMM") != nullptr && atoi(getenv("DMEL_NO_DFTMM")) != 0;
    return !off;
}
static bool dftmm_takes_128() {
    static const bool on = getenv("DMEL_DFTMM_128") != nullptr && atoi(getenv("DMEL_DFTMM_128")) != 0;
    return on;
}
static bool dftmm_serves(int32_t n_fft, int32_t n_mels, const dmel_melbank* bank, int32_t flags, const float* dy) {
    if (!dftmm_enabled() || dy != nullptr) return false;
    if (n_fft < 4 || n_fft > mm::MAX_NFFT || (n_fft & 1)) return false;
    if (n_fft == 128 && !dftmm_takes_128()) return false;
    if (flags & (DMEL_FLAG_TIME_MAJOR | DMEL_FLAG_X_I16 | DMEL_FLAG_X_F16)) return false;
    if (n_mels > 0 && bank->out_scale != nullptr) return false;
    // the epilogue's chunk table (shared memory) is sized for this many mels / 4-tap chunks
    if (n_mels > mm::MAX_MELS || (n_mels > 0 && bank->mel_nnz / 4 + n_mels > mm::MAX_CHUNKS)) return false;
    return true;
}
static int launch_dftmm_fwd(const float* x, float* y, int64_t B, int64_t N, int64_t stride, int32_t n_fft, int32_t hop,
                            int32_t n_mels, float log_eps, const MelDev& mel, const WsView& ws, cudaStream_t st) {
    static KernelCache kc;
    int rc;
    if ((rc = set_smem(mm::dmel_dftmm_fwd_kernel, size_t(mm::SMEM_BYTES), &kc))) return rc;
    mm::MmGeom g{};
    g.B = int(B); g.N = int(N); g.T = int(N / hop + 1); g.hop = hop; g.L = n_fft; g.M = n_mels;
    g.x_stride = stride; g.log_eps = log_eps;
    const int64_t tiles = (int64_t(g.B) * g.T + mm::ROWS - 1) / mm::ROWS;
    int sms = NUM_SMS, dev = current_device(), v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) sms = v;
    const unsigned grid = unsigned(std::min<int64_t>(tiles, sms));
    ProfScope prof(PROF_FWD, st);
    mm::dmel_dftmm_fwd_kernel<<<grid, mm::THREADS, mm::SMEM_BYTES, st>>>(x, y, g, mel, ws.win, ws.means, ws.hdr);
    return check_launch("dmel_dftmm_fwd_kernel");
}

static int forward_impl(const void* x, int64_t batch, int64_t n_points, int64_t x_row_stride, float lambda,
                        int32_t win_length, int32_t n_fft, int32_t hop, int32_t n_mels, const dmel_melbank* bank,
                        float log_eps, int32_t flags, float* y, float* dy, void* workspace, size_t workspace_bytes,
                        void* cuda_stream, const float* means_src = nullptr) {
    if (!y) return DMEL_ERR_NULL_POINTER;
    int rc = validate(x, batch, n_points, x_row_stride, win_length, n_fft, hop, n_mels, bank, workspace, workspace_bytes);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    const bool spec = (n_mels == 0);
    const int rows = spec ? n_fft / 2 + 1 : n_mels;
    LaunchPlan lp{};
    if ((rc = plan_launch(batch, n_points, n_fft, hop, rows, spec ? 0 : round_up4(n_mels) + round_up4(n_mels + 1), spec ? 0 : bank->mel_nnz, false,
                          false, &lp, dy != nullptr, !spec && bank->seg_tables != nullptr)))
        return rc;
    if (dy && lp.path != 0) return DMEL_ERR_UNSUPPORTED_NFFT;   // the JVP forward exists for the tiled FFT path only
    // the (B,1,T,M) layout: tiled FFT kernels, and the big-FFT path in mel mode
    if ((flags & DMEL_FLAG_TIME_MAJOR) && !(lp.path == 0 || (lp.path == 2 && !spec))) return DMEL_ERR_UNSUPPORTED_NFFT;
    if (int64_t(batch) * lp.tiles_per_clip > INT32_MAX) return DMEL_ERR_BAD_GEOMETRY;
    int x_dtype = DMEL_X_F32;
    if ((rc = x_dtype_of(flags, lp, &x_dtype))) return rc;
    const bool gen2 = lp.path == 0 && lp.gen == 2;
    WsView ws = ws_view(workspace, batch, n_fft);
    const MelDev mel = mel_dev(bank);
    if ((rc = run_prep(x, batch, n_points, x_row_stride, fabsf(lambda), win_length, n_fft, flags, ws, st, gen2, means_src)))
        return rc;
    Geom g = make_geom(batch, n_points, x_row_stride, n_fft, hop, rows, log_eps, lp);
    g.x_dtype = x_dtype;
    g.row_aligned = row_alignment_ok(x, x_row_stride, x_dtype);
    g.time_major = (flags & DMEL_FLAG_TIME_MAJOR) ? 1 : 0;
    const float* xf = static_cast<const float*>(x);
    if (dftmm_serves(n_fft, n_mels, bank, flags, dy))
        return launch_dftmm_fwd(xf, y, batch, n_points, x_row_stride, n_fft, hop, n_mels, log_eps, mel, ws, st);
    if (lp.path == 1) {
        if ((rc = set_smem(dmel_dft_kernel<false, false>, lp.smem))) return rc;
        ProfScope prof(PROF_FWD, st);
        dmel_dft_kernel<false, false><<<unsigned(int64_t(g.B) * g.T), DFT_THREADS, lp.smem, st>>>(
            xf, nullptr, nullptr, y, nullptr, g, mel, spec ? 1 : 0, ws.win, ws.tw, ws.means, ws.hdr, ws.clipsum);
        return check_launch("dmel_dft_kernel<fwd>");
    }
    if (lp.path == 2) {
        return n_fft == 8192 ? launch_big_fwd<8192>(xf, y, g, mel, spec, ws, lp, st)
                             : launch_big_fwd<16384>(xf, y, g, mel, spec, ws, lp, st);
    }
    if (gen2) {
        switch (n_fft) {
#define DMEL_CASE(LL) case LL: rc = launch_tiled2_fwd_##LL(lp.NG, x, y, dy, g, mel, spec, ws, lp, st); break;
            DMEL_CASE(256) DMEL_CASE(512) DMEL_CASE(1024) DMEL_CASE(2048) DMEL_CASE(4096)
#undef DMEL_CASE
            default: rc = DMEL_ERR_UNSUPPORTED_NFFT;
        }
        return rc;
    }
    rc = launch_tiled1_fwd(n_fft, lp.NG, xf, y, dy, g, mel, spec, ws, lp, st);
    return rc;
}

extern "C" DMEL_API int dmel_forward_f32(const void* x, int64_t batch, int64_t n_points, int64_t x_row_stride, float lambda,
                                int32_t win_length, int32_t n_fft, int32_t hop, int32_t n_mels,
                                const dmel_melbank* bank, float log_eps, int32_t flags, float* y,
                                void* workspace, size_t workspace_bytes, void* cuda_stream) {
    return forward_impl(x, batch, n_points, x_row_stride, lambda, win_length, n_fft, hop, n_mels, bank, log_eps, flags, y,
                        nullptr, workspace, workspace_bytes, cuda_stream);
}

extern "C" DMEL_API int dmel_forward_jvp_f32(const void* x, int64_t batch, int64_t n_points, int64_t x_row_stride,
                                    float lambda, int32_t win_length, int32_t n_fft, int32_t hop, int32_t n_mels,
                                    const dmel_melbank* bank, float log_eps, int32_t flags, float* y,
                                    float* dy_dlambda, void* workspace, size_t workspace_bytes, void* cuda_stream) {
    if (!dy_dlambda) return DMEL_ERR_NULL_POINTER;
    return forward_impl(x, batch, n_points, x_row_stride, lambda, win_length, n_fft, hop, n_mels, bank, log_eps, flags, y,
                        dy_dlambda, workspace, workspace_bytes, cuda_stream);
}

extern "C" DMEL_API int dmel_forward_multi_f32(const void* x, int64_t batch, int64_t n_points, int64_t x_row_stride,
                                               int32_t n_lambda, const float* lambdas, const int32_t* win_lengths,
                                               const int32_t* n_ffts, int32_t hop, int32_t n_mels,
                                               const dmel_melbank* const* banks, float log_eps, int32_t flags,
                                               float* const* ys, void* const* workspaces, const size_t* workspace_bytes,
                                               void* cuda_stream) {
    if (!lambdas || !win_lengths || !n_ffts || !ys || !workspaces || !workspace_bytes) return DMEL_ERR_NULL_POINTER;
    if (n_lambda <= 0 || n_lambda > DMEL_MAX_MULTI_LAMBDA) return DMEL_ERR_BAD_GEOMETRY;
    if (n_mels > 0 && !banks) return DMEL_ERR_NULL_POINTER;
    const float* means0 = nullptr;
    for (int k = 0; k < n_lambda; ++k) {
        // the per-clip means (models.py:38) do not depend on lambda: the first pass sums them, the others copy them
        const int rc = forward_impl(x, batch, n_points, x_row_stride, lambdas[k], win_lengths[k], n_ffts[k], hop, n_mels,
                                    n_mels > 0 ? banks[k] : nullptr, log_eps, flags, ys[k], nullptr, workspaces[k],
                                    workspace_bytes[k], cuda_stream, means0);
        if (rc) return rc;
        if (k == 0 && !(flags & DMEL_FLAG_KEEP_MEAN)) means0 = ws_view(workspaces[0], batch, n_ffts[0]).means;
    }
    return DMEL_OK;
}

extern "C" DMEL_API int dmel_jvp_supported(int32_t n_fft) { const FftInfo fi = fft_info(n_fft); return (fi.ok && !fi.big) ? 1 : 0; }

extern "C" DMEL_API int dmel_time_major_supported(int32_t n_fft, int32_t n_mels) {
    const FftInfo fi = fft_info(n_fft);
    return (fi.ok && (!fi.big || n_mels > 0)) ? 1 : 0;
}

extern "C" DMEL_API int dmel_x16_supported(int32_t n_fft) { return fft2_info(n_fft).ok ? 1 : 0; }

extern "C" DMEL_API size_t dmel_segment_table_bytes(int32_t n_fft) {
    return fft2_info(n_fft).ok ? seg_blob_bytes(n_fft) : 0;
}

extern "C" DMEL_API int dmel_build_segment_tables(const dmel_melbank* bank, int32_t n_fft, void* tables, size_t bytes,
                                                  void* cuda_stream, int32_t* usable) {
    if (usable) *usable = 0;
    if (!bank || !tables) return DMEL_ERR_NULL_POINTER;
    if (!fft2_info(n_fft).ok) return DMEL_ERR_UNSUPPORTED_NFFT;
    if (bank->n_freqs != n_fft / 2 + 1 || bank->n_mels <= 0) return DMEL_ERR_BAD_MELBANK;
    if (!bank->bin_m0 || !bank->bin_off || !bank->bin_w) return DMEL_ERR_NULL_POINTER;
    if (bytes < seg_blob_bytes(n_fft)) return DMEL_ERR_WORKSPACE_TOO_SMALL;
    if (reinterpret_cast<uintptr_t>(tables) & 15) return DMEL_ERR_BAD_MELBANK;
    MelDev m{};
    m.bin_m0 = bank->bin_m0; m.bin_off = bank->bin_off; m.bin_w = bank->bin_w;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    dmel_segment_kernel<<<1, PREP_THREADS, 0, st>>>(m, bank->n_mels, n_fft, n_fft / 16, seg_view(tables, n_fft));
    int rc = check_launch("dmel_segment_kernel");
    if (rc) return rc;
    // one-time set-up call: read the verdict back (this is the only entry point that synchronises its stream)
    SegHeader h{};
    if (cudaMemcpyAsync(&h, tables, sizeof(h), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess)
        return check_launch("dmel_build_segment_tables readback") ? DMEL_ERR_LAUNCH_FAILED : DMEL_ERR_LAUNCH_FAILED;
    if (usable) *usable = h.use_seg ? 1 : 0;
    return DMEL_OK;
}

extern "C" DMEL_API int dmel_lambda_grad_f32(const float* grad_y, const float* dy_dlambda, int64_t count, float lambda,
                                    float* grad_lambda, void* workspace, size_t workspace_bytes, void* cuda_stream) {
    if (!grad_y || !dy_dlambda || !grad_lambda || !workspace) return DMEL_ERR_NULL_POINTER;
    if (count <= 0) return DMEL_ERR_BAD_GEOMETRY;
    if (workspace_bytes < sizeof(WsHeader)) return DMEL_ERR_WORKSPACE_TOO_SMALL;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    const float sign = lambda > 0.f ? 1.f : (lambda < 0.f ? -1.f : 0.f);
    long long blocks = (count / 4 + 255) / 256;
    blocks = std::max<long long>(1, std::min<long long>(blocks, 4LL * NUM_SMS));
    ProfScope prof(PROF_BWD, st);
    dmel_lambda_grad_kernel<<<unsigned(blocks), 256, 0, st>>>(grad_y, dy_dlambda, count, sign,
                                                              static_cast<WsHeader*>(workspace), grad_lambda);
    return check_launch("dmel_lambda_grad_kernel");
}

static int backward_impl(const void* x, int64_t batch, int64_t n_points, int64_t x_row_stride, float lambda,
                         int32_t win_length, int32_t n_fft, int32_t hop, int32_t n_mels, const dmel_melbank* bank,
                         float log_eps, int32_t flags, const float* y, const float* grad_y, float* grad_x,
                         float* grad_lambda, void* workspace, size_t workspace_bytes, int32_t prepared,
                         void* cuda_stream, const void* peer_bases_dev, int32_t rank, int32_t world) {
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
    int x_dtype = DMEL_X_F32;
    if ((rc = x_dtype_of(flags, lp, &x_dtype))) return rc;
    const bool gen2 = lp.path == 0 && lp.gen == 2;
    const float* xf = static_cast<const float*>(x);
    WsView ws = ws_view(workspace, batch, n_fft);
    if (!prepared) {
        if ((rc = run_prep(x, batch, n_points, x_row_stride, fabsf(lambda), win_length, n_fft, flags, ws, st, gen2)))
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
    g.x_dtype = x_dtype;
    g.row_aligned = row_alignment_ok(x, x_row_stride, x_dtype);
    g.time_major = (flags & DMEL_FLAG_TIME_MAJOR) ? 1 : 0;
    g.dx_vec4 = (grad_x != nullptr && (reinterpret_cast<uintptr_t>(grad_x) & 15) == 0 && (n_points & 3) == 0) ? 1 : 0;
    if (g.time_major && !(lp.path == 0 || (lp.path == 2 && !spec))) return DMEL_ERR_UNSUPPORTED_NFFT;
    const MelDev mel = mel_dev(bank);
    if (lp.path == 1) {
        const unsigned grid = unsigned(int64_t(g.B) * g.T);
        ProfScope prof(PROF_BWD, st);
        if (need_dx) {
            if ((rc = set_smem(dmel_dft_kernel<true, true>, lp.smem))) return rc;
            dmel_dft_kernel<true, true><<<grid, DFT_THREADS, lp.smem, st>>>(
                xf, y, grad_y, nullptr, grad_x, g, mel, spec ? 1 : 0, ws.win, ws.tw, ws.means, ws.hdr, ws.clipsum);
        } else {
            if ((rc = set_smem(dmel_dft_kernel<true, false>, lp.smem))) return rc;
            dmel_dft_kernel<true, false><<<grid, DFT_THREADS, lp.smem, st>>>(
                xf, y, grad_y, nullptr, nullptr, g, mel, spec ? 1 : 0, ws.win, ws.tw, ws.means, ws.hdr, ws.clipsum);
        }
        if ((rc = check_launch("dmel_dft_kernel<bwd>"))) return rc;
    } else if (lp.path == 2) {
        rc = n_fft == 8192 ? launch_big_bwd<8192>(xf, y, grad_y, grad_x, g, mel, spec, ws, lp, st)
                           : launch_big_bwd<16384>(xf, y, grad_y, grad_x, g, mel, spec, ws, lp, st);
        if (rc) return rc;
    } else if (gen2) {
        switch (n_fft) {
#define DMEL_CASE(LL) case LL: rc = launch_tiled2_bwd_##LL(lp.NG, x, y, grad_y, grad_x, g, mel, spec, ws, lp, st); break;
            DMEL_CASE(256) DMEL_CASE(512) DMEL_CASE(1024) DMEL_CASE(2048) DMEL_CASE(4096)
#undef DMEL_CASE
            default: rc = DMEL_ERR_UNSUPPORTED_NFFT;
        }
        if (rc) return rc;
    } else {
        rc = launch_tiled1_bwd(n_fft, lp.NG, xf, y, grad_y, grad_x, g, mel, spec, ws, lp, st);
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
                                                         sign, grad_lambda, static_cast<const long long*>(peer_bases_dev),
                                                         rank, world, PEER_TIMEOUT_CYCLES);
    return check_launch("dmel_finalize_kernel");
}

extern "C" DMEL_API int dmel_backward_f32(const void* x, int64_t batch, int64_t n_points, int64_t x_row_stride, float lambda,
                                 int32_t win_length, int32_t n_fft, int32_t hop, int32_t n_mels,
                                 const dmel_melbank* bank, float log_eps, int32_t flags, const float* y,
                                 const float* grad_y, float* grad_x, float* grad_lambda, void* workspace,
                                 size_t workspace_bytes, int32_t prepared, void* cuda_stream) {
    return backward_impl(x, batch, n_points, x_row_stride, lambda, win_length, n_fft, hop, n_mels, bank, log_eps, flags, y,
                         grad_y, grad_x, grad_lambda, workspace, workspace_bytes, prepared, cuda_stream, nullptr, 0, 1);
}

extern "C" DMEL_API int dmel_backward_allreduce_f32(const void* x, int64_t batch, int64_t n_points, int64_t x_row_stride,
                                           float lambda, int32_t win_length, int32_t n_fft, int32_t hop,
                                           int32_t n_mels, const dmel_melbank* bank, float log_eps, int32_t flags,
                                           const float* y, const float* grad_y, float* grad_x, float* grad_lambda,
                                           void* workspace, size_t workspace_bytes, int32_t prepared,
                                           void* cuda_stream, const void* peer_bases_dev, int32_t rank,
                                           int32_t world) {
    if (!peer_bases_dev) return DMEL_ERR_NULL_POINTER;
    if (world < 1 || world > PEER_MAX_WORLD || rank < 0 || rank >= world) return DMEL_ERR_BAD_GEOMETRY;
    return backward_impl(x, batch, n_points, x_row_stride, lambda, win_length, n_fft, hop, n_mels, bank, log_eps, flags, y,
                         grad_y, grad_x, grad_lambda, workspace, workspace_bytes, prepared, cuda_stream, peer_bases_dev,
                         rank, world);
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

extern "C" DMEL_API size_t dmel_peer_buffer_bytes(int32_t world) {
    return (world < 1 || world > PEER_MAX_WORLD) ? 0 : size_t(4 * world + 4) * 4;
}

extern "C" DMEL_API int dmel_peer_allreduce_f32(float* value, const void* peer_bases_dev, int32_t rank, int32_t world,
                                       void* cuda_stream) {
    if (!value || !peer_bases_dev) return DMEL_ERR_NULL_POINTER;
    if (world < 1 || world > PEER_MAX_WORLD || rank < 0 || rank >= world) return DMEL_ERR_BAD_GEOMETRY;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
        dmel_peer_allreduce_kernel<<<1, 32, 0, st>>>(value, static_cast<const long long*>(peer_bases_dev), rank, world,
                                                  PEER_TIMEOUT_CYCLES);
    return check_launch("dmel_peer_allreduce_kernel");
}

// ---- symmetric buffers for the peer all-reduce: raw cudaMalloc (IPC handles refer to whole allocations) owned by
// the caller through these explicit calls; peers are mapped into the CURRENT device's context with peer access.
extern "C" DMEL_API int dmel_peer_alloc(int32_t world, void** buffer, unsigned char* ipc_handle_64) {
    if (!buffer || !ipc_handle_64) return DMEL_ERR_NULL_POINTER;
    const size_t bytes = dmel_peer_buffer_bytes(world);
    if (bytes == 0) return DMEL_ERR_BAD_GEOMETRY;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) return check_launch("cudaMalloc(peer buffer)") ? DMEL_ERR_LAUNCH_FAILED : DMEL_ERR_LAUNCH_FAILED;
    if (cudaMemset(p, 0, bytes) != cudaSuccess) { cudaFree(p); return check_launch("cudaMemset(peer buffer)") ? DMEL_ERR_LAUNCH_FAILED : DMEL_ERR_LAUNCH_FAILED; }
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, p) != cudaSuccess) { cudaFree(p); (void)check_launch("cudaIpcGetMemHandle"); return DMEL_ERR_LAUNCH_FAILED; }
    memcpy(ipc_handle_64, &h, 64);
    *buffer = p;
    return DMEL_OK;
}

extern "C" DMEL_API int dmel_peer_open(const unsigned char* ipc_handle_64, void** mapped) {
    if (!ipc_handle_64 || !mapped) return DMEL_ERR_NULL_POINTER;
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle_64, 64);
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        (void)check_launch("cudaIpcOpenMemHandle");
        return DMEL_ERR_LAUNCH_FAILED;
    }
    *mapped = p;
    return DMEL_OK;
}

extern "C" DMEL_API int dmel_peer_close(void* mapped) {
    if (!mapped) return DMEL_OK;
    return cudaIpcCloseMemHandle(mapped) == cudaSuccess ? DMEL_OK : DMEL_ERR_LAUNCH_FAILED;
}

extern "C" DMEL_API int dmel_peer_free(void* buffer) {
    if (!buffer) return DMEL_OK;
    return cudaFree(buffer) == cudaSuccess ? DMEL_OK : DMEL_ERR_LAUNCH_FAILED;
}

// synchronous read of the rank's own error word (1 after a peer timeout); for self-tests, not for the hot path
extern "C" DMEL_API int dmel_peer_timed_out(const void* buffer, int32_t world, int32_t* timed_out) {
    if (!buffer || !timed_out) return DMEL_ERR_NULL_POINTER;
    if (dmel_peer_buffer_bytes(world) == 0) return DMEL_ERR_BAD_GEOMETRY;
    int v = 0;
    if (cudaMemcpy(&v, static_cast<const int*>(buffer) + 4 * world + 1, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) {
        (void)check_launch("cudaMemcpy(peer status)");
        return DMEL_ERR_LAUNCH_FAILED;
    }
    *timed_out = v;
    return DMEL_OK;
}
