// inst_tiled.cu — launchers + template instantiations of the tiled kernels, compiled once per
//   -DDMEL_INST_GEN=1|2  -DDMEL_INST_L=<n_fft>  -DDMEL_INST_DIR=0 (forward) | 1 (backward)
// so that the heavy instantiations build in parallel (see Makefile).  The group counts per transform length must
// match fft_info() / fft2_info() in dmel_kernels.cu.
#include <stdlib.h>

#include "dmel_host.cuh"
#include "fft_device.cuh"
#if DMEL_INST_GEN == 1
#include "dmel_tiled_kernels.cuh"
#else
#include "dmel_tiled2_kernels.cuh"
#endif

namespace dmel {

#if DMEL_INST_GEN == 1
template <int L, int NG, bool SPEC, bool JVP>
static int launch_fwd_t(const float* x, float* y, float* dy, const Geom& g, const MelDev& mel, const WsView& ws,
                        const LaunchPlan& lp, cudaStream_t st) {
    static KernelCache kc;     // one per instantiation
    int rc;
    if ((rc = set_smem(dmel_fwd_kernel<L, NG, SPEC, JVP>, lp.smem, &kc))) return rc;
    const unsigned grid = persistent_grid(dmel_fwd_kernel<L, NG, SPEC, JVP>, lp, int64_t(g.B) * g.tiles_per_clip, &kc);
    ProfScope prof(PROF_FWD, st);
    dmel_fwd_kernel<L, NG, SPEC, JVP><<<grid, lp.threads, lp.smem, st>>>(x, y, dy, g, mel, ws.win, ws.tw, ws.means, ws.hdr);
    return check_launch("dmel_fwd_kernel");
}

template <int L, int NG, bool SPEC, bool NEED_DX>
static int launch_bwd_t(const float* x, const float* y, const float* gy, float* dx, const Geom& g, const MelDev& mel,
                        const WsView& ws, const LaunchPlan& lp, cudaStream_t st) {
    static KernelCache kc;     // one per instantiation
    int rc;
    if ((rc = set_smem(dmel_bwd_kernel<L, NG, SPEC, NEED_DX>, lp.smem, &kc))) return rc;
    const unsigned grid = persistent_grid(dmel_bwd_kernel<L, NG, SPEC, NEED_DX>, lp, int64_t(g.B) * g.tiles_per_clip, &kc);
    ProfScope prof(PROF_BWD, st);
    dmel_bwd_kernel<L, NG, SPEC, NEED_DX>
        <<<grid, lp.threads, lp.smem, st>>>(x, y, gy, dx, g, mel, ws.win, ws.tw, ws.means, ws.hdr, ws.clipsum);
    return check_launch("dmel_bwd_kernel");
}

template <int L, int NG>
static int launch_fwd(const float* x, float* y, float* dy, const Geom& g, const MelDev& mel, bool spec, const WsView& ws,
                      const LaunchPlan& lp, cudaStream_t st) {
    if (dy) {
        return spec ? launch_fwd_t<L, NG, true, true>(x, y, dy, g, mel, ws, lp, st)
                    : launch_fwd_t<L, NG, false, true>(x, y, dy, g, mel, ws, lp, st);
    }
    return spec ? launch_fwd_t<L, NG, true, false>(x, y, nullptr, g, mel, ws, lp, st)
                : launch_fwd_t<L, NG, false, false>(x, y, nullptr, g, mel, ws, lp, st);
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

#else
// ---- second generation
template <int L, int NG, bool SPEC, bool JVP>
static int launch_fwd2_t(const void* x, float* y, float* dy, const Geom& g, const MelDev& mel, const WsView& ws,
                         const LaunchPlan& lp, cudaStream_t st) {
    static KernelCache kc;     // one per instantiation
    int rc;
    if ((rc = set_smem(v2::dmel2_fwd_kernel<L, NG, SPEC, JVP>, lp.smem, &kc))) return rc;
    const unsigned grid = persistent_grid(v2::dmel2_fwd_kernel<L, NG, SPEC, JVP>, lp, int64_t(g.B) * g.tiles_per_clip, &kc);
    ProfScope prof(PROF_FWD, st);
    v2::dmel2_fwd_kernel<L, NG, SPEC, JVP><<<grid, lp.threads, lp.smem, st>>>(x, y, dy, g, mel, ws.win, ws.tw, ws.means, ws.hdr);
    rc = check_launch("dmel2_fwd_kernel");
    if (rc) {
        cudaFuncAttributes fa{};
        cudaFuncGetAttributes(&fa, v2::dmel2_fwd_kernel<L, NG, SPEC, JVP>);
        fprintf(stderr, "[dmel] dmel2_fwd_kernel<%d,%d,%d,%d>: grid %u threads %d dynamic smem %zu (max dynamic %d, static %zu, regs %d)\n", L, NG,
                int(SPEC), int(JVP), grid, lp.threads, lp.smem, fa.maxDynamicSharedSizeBytes, fa.sharedSizeBytes, fa.numRegs);
    }
    return rc;
}

template <int L, int NG, bool SPEC, bool NEED_DX>
static int launch_bwd2_t(const void* x, const float* y, const float* gy, float* dx, const Geom& g, const MelDev& mel,
                         const WsView& ws, const LaunchPlan& lp, cudaStream_t st) {
    static KernelCache kc;     // one per instantiation
    int rc;
    // DMEL_EXP_PAD_SMEM: occupancy experiments — pad the request so that fewer CTAs fit per SM
    static const size_t pad = getenv("DMEL_EXP_PAD_SMEM") ? size_t(atol(getenv("DMEL_EXP_PAD_SMEM"))) : 0;
    LaunchPlan lpp = lp;
    lpp.smem += pad;
    if ((rc = set_smem(v2::dmel2_bwd_kernel<L, NG, SPEC, NEED_DX>, lpp.smem, &kc))) return rc;
    const unsigned grid = persistent_grid(v2::dmel2_bwd_kernel<L, NG, SPEC, NEED_DX>, lpp, int64_t(g.B) * g.tiles_per_clip, &kc);
    ProfScope prof(PROF_BWD, st);
    v2::dmel2_bwd_kernel<L, NG, SPEC, NEED_DX>
        <<<grid, lp.threads, lpp.smem, st>>>(x, y, gy, dx, g, mel, ws.win, ws.tw, ws.means, ws.hdr, ws.clipsum);
    return check_launch("dmel2_bwd_kernel");
}

template <int L, int NG>
static int launch_fwd2(const void* x, float* y, float* dy, const Geom& g, const MelDev& mel, bool spec, const WsView& ws,
                       const LaunchPlan& lp, cudaStream_t st) {
    if (dy) {
        return spec ? launch_fwd2_t<L, NG, true, true>(x, y, dy, g, mel, ws, lp, st)
                    : launch_fwd2_t<L, NG, false, true>(x, y, dy, g, mel, ws, lp, st);
    }
    return spec ? launch_fwd2_t<L, NG, true, false>(x, y, nullptr, g, mel, ws, lp, st)
                : launch_fwd2_t<L, NG, false, false>(x, y, nullptr, g, mel, ws, lp, st);
}

template <int L, int NG>
static int launch_bwd2(const void* x, const float* y, const float* gy, float* dx, const Geom& g, const MelDev& mel,
                       bool spec, const WsView& ws, const LaunchPlan& lp, cudaStream_t st) {
    if (spec) {
        return dx ? launch_bwd2_t<L, NG, true, true>(x, y, gy, dx, g, mel, ws, lp, st)
                  : launch_bwd2_t<L, NG, true, false>(x, y, gy, dx, g, mel, ws, lp, st);
    }
    return dx ? launch_bwd2_t<L, NG, false, true>(x, y, gy, dx, g, mel, ws, lp, st)
              : launch_bwd2_t<L, NG, false, false>(x, y, gy, dx, g, mel, ws, lp, st);
}


#endif

// group counts instantiated per transform length (largest first)
template <int L> struct InstNG;
template <> struct InstNG<128>  { static constexpr int A = 8, B = 4; };
template <> struct InstNG<256>  { static constexpr int A = 8, B = 4; };
template <> struct InstNG<512>  { static constexpr int A = 8, B = 4; };
template <> struct InstNG<1024> { static constexpr int A = 4, B = 2; };
template <> struct InstNG<2048> { static constexpr int A = 4, B = 1; };
template <> struct InstNG<4096> { static constexpr int A = 1, B = 1; };

#define DMEL_CAT_(a, b) a##b
#define DMEL_CAT(a, b) DMEL_CAT_(a, b)
constexpr int IL = DMEL_INST_L;

#if DMEL_INST_GEN == 1 && DMEL_INST_DIR == 0
int launch_tiled1_fwd(int n_fft, int ng, const float* x, float* y, float* dy, const Geom& g, const MelDev& mel, bool spec,
                      const WsView& ws, const LaunchPlan& lp, cudaStream_t st) {
    if (n_fft != IL) return DMEL_ERR_UNSUPPORTED_NFFT;
    if (ng == InstNG<IL>::A) return launch_fwd<IL, InstNG<IL>::A>(x, y, dy, g, mel, spec, ws, lp, st);
    if (ng == InstNG<IL>::B) return launch_fwd<IL, InstNG<IL>::B>(x, y, dy, g, mel, spec, ws, lp, st);
    return DMEL_ERR_UNSUPPORTED_NFFT;
}
#elif DMEL_INST_GEN == 1
int launch_tiled1_bwd(int n_fft, int ng, const float* x, const float* y, const float* gy, float* dx, const Geom& g,
                      const MelDev& mel, bool spec, const WsView& ws, const LaunchPlan& lp, cudaStream_t st) {
    if (n_fft != IL) return DMEL_ERR_UNSUPPORTED_NFFT;
    if (ng == InstNG<IL>::A) return launch_bwd<IL, InstNG<IL>::A>(x, y, gy, dx, g, mel, spec, ws, lp, st);
    if (ng == InstNG<IL>::B) return launch_bwd<IL, InstNG<IL>::B>(x, y, gy, dx, g, mel, spec, ws, lp, st);
    return DMEL_ERR_UNSUPPORTED_NFFT;
}
#elif DMEL_INST_DIR == 0
int DMEL_CAT(launch_tiled2_fwd_, DMEL_INST_L)(int ng, const void* x, float* y, float* dy, const Geom& g, const MelDev& mel,
                                              bool spec, const WsView& ws, const LaunchPlan& lp, cudaStream_t st) {
    if (ng == InstNG<IL>::A) return launch_fwd2<IL, InstNG<IL>::A>(x, y, dy, g, mel, spec, ws, lp, st);
    if (ng == InstNG<IL>::B) return launch_fwd2<IL, InstNG<IL>::B>(x, y, dy, g, mel, spec, ws, lp, st);
    return DMEL_ERR_UNSUPPORTED_NFFT;
}
#else
int DMEL_CAT(launch_tiled2_bwd_, DMEL_INST_L)(int ng, const void* x, const float* y, const float* gy, float* dx,
                                              const Geom& g, const MelDev& mel, bool spec, const WsView& ws,
                                              const LaunchPlan& lp, cudaStream_t st) {
    if (ng == InstNG<IL>::A) return launch_bwd2<IL, InstNG<IL>::A>(x, y, gy, dx, g, mel, spec, ws, lp, st);
    if (ng == InstNG<IL>::B) return launch_bwd2<IL, InstNG<IL>::B>(x, y, gy, dx, g, mel, spec, ws, lp, st);
    return DMEL_ERR_UNSUPPORTED_NFFT;
}
#endif

}  // namespace dmel
