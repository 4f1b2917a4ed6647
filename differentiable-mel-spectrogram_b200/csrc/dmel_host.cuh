// dmel_host.cuh — host-side state shared by the translation units of libdmel_b200.so: the launch plan, the optional
// launch instrumentation, the per-kernel launch cache and the entry points of the tiled-kernel launchers.
//
// The library is built from several translation units so that the template instantiations of the tiled kernels
// compile in parallel (`make -j`): dmel_kernels.cu holds the planner, the small / big-FFT / direct-DFT kernels and the
// C ABI; inst_tiled.cu is compiled once per (generation, transform length, direction) and holds the launchers of
// those instantiations.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <algorithm>
#include <atomic>
#include <vector>

#include "../../include/dmel.h"
#include "dmel_common.cuh"

namespace dmel {

constexpr int NUM_SMS = 148;

struct LaunchPlan {
    int path;  // 0 FFT, 1 direct DFT, 2 big FFT
    int gen;   // tiled FFT path: 1 first-generation kernels, 2 second generation
    int lookahead;
    int NG, SF, TT, span, sub_span, tiles_per_clip, threads;
    int csr_ints, csr_floats;
    size_t smem;
};

// ---- optional instrumentation for bench.py: launch counter + per-kernel CUDA-event timing.
// Off by default; when on, every main-kernel launch is bracketed by a pair of events recorded on
// the launching stream (no synchronisation until dmel_profile_read).
enum { PROF_PREP = 0, PROF_FWD = 1, PROF_BWD = 2, PROF_FINALIZE = 3, PROF_ZERO = 4, PROF_KINDS = 5 };
extern std::atomic<long long> g_launches[PROF_KINDS];   // defined in dmel_kernels.cu
struct ProfState {
    bool on = false;
    std::vector<cudaEvent_t> ev[PROF_KINDS];  // start/stop pairs
};
extern thread_local ProfState t_prof;                    // defined in dmel_kernels.cu

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

static inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        fprintf(stderr, "[dmel] %s: %s\n", what, cudaGetErrorString(e));
        return DMEL_ERR_LAUNCH_FAILED;
    }
    return DMEL_OK;
}

// Per-kernel launch state that does not change between calls: the opt-in to > 48 KB of dynamic shared memory and the
// occupancy behind the persistent grid.  Both are properties of (kernel, device); they are looked up once per kernel
// instantiation and device instead of on every launch (the host cost of a call is two launches' worth of driver work).
constexpr int MAX_DEVICES = 64;
struct KernelCache {
    std::atomic<size_t> smem_ok[MAX_DEVICES];     // largest dynamic shared memory size already opted in
    std::atomic<uint64_t> occ[MAX_DEVICES];       // (threads << 40 | smem << 8 | per_sm) of the last occupancy query
    std::atomic<int> sms[MAX_DEVICES];
};

static inline int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        (void)cudaGetLastError();
        dev = 0;
    }
    return dev;
}

template <typename K>
static int set_smem(K kernel, size_t smem, KernelCache* kc = nullptr) {
    // the 48 KB a kernel gets without opting in covers dynamic AND static shared memory (the tiled kernels keep a few
    // hundred bytes of static tables): opt in from a margin below it
    if (smem > 47 * 1024) {
        const int dev = current_device();
        if (kc && dev < MAX_DEVICES && kc->smem_ok[dev].load(std::memory_order_relaxed) >= smem) return DMEL_OK;
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (e != cudaSuccess) {
            fprintf(stderr, "[dmel] cudaFuncSetAttribute(%zu): %s\n", smem, cudaGetErrorString(e));
            (void)cudaGetLastError();
            return DMEL_ERR_LAUNCH_FAILED;
        }
        if (kc && dev < MAX_DEVICES) kc->smem_ok[dev].store(smem, std::memory_order_relaxed);
    }
    return DMEL_OK;
}

// persistent launch: one resident wave of CTAs (occupancy x SM count), each walking tiles with stride gridDim
template <typename K>
static unsigned persistent_grid(K kernel, const LaunchPlan& lp, int64_t total_tiles, KernelCache* kc = nullptr) {
    const int dev = current_device();
    const uint64_t key = (uint64_t(lp.threads) << 40) | (uint64_t(lp.smem) << 8);
    int per_sm = 0, sms = 0;
    if (kc && dev < MAX_DEVICES) {
        const uint64_t c = kc->occ[dev].load(std::memory_order_relaxed);
        if ((c & ~uint64_t(0xff)) == key && (c & 0xff) != 0) {
            per_sm = int(c & 0xff);
            sms = kc->sms[dev].load(std::memory_order_relaxed);
        }
    }
    if (per_sm < 1 || sms < 1) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, lp.threads, lp.smem) != cudaSuccess || per_sm < 1) {
            (void)cudaGetLastError();
            per_sm = 1;
        }
        sms = NUM_SMS;
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) sms = v;
        if (kc && dev < MAX_DEVICES) {
            kc->sms[dev].store(sms, std::memory_order_relaxed);
            kc->occ[dev].store(key | uint64_t(per_sm & 0xff), std::memory_order_relaxed);
        }
    }
    const int64_t cap = int64_t(per_sm) * sms;
    return unsigned(std::min<int64_t>(total_tiles, cap));
}

// ---- launchers of the tiled kernels (inst_tiled.cu).  `ng` must be one of the group counts instantiated for the
// transform length (fft_info / fft2_info in dmel_kernels.cu); anything else returns DMEL_ERR_UNSUPPORTED_NFFT.
// first generation (dmel_tiled_kernels.cuh): n_fft 128
int launch_tiled1_fwd(int n_fft, int ng, const float* x, float* y, float* dy, const Geom& g, const MelDev& mel, bool spec,
                      const WsView& ws, const LaunchPlan& lp, cudaStream_t st);
int launch_tiled1_bwd(int n_fft, int ng, const float* x, const float* y, const float* gy, float* dx, const Geom& g,
                      const MelDev& mel, bool spec, const WsView& ws, const LaunchPlan& lp, cudaStream_t st);
// second generation (dmel_tiled2_kernels.cuh): n_fft 256..4096
#define DMEL_TILED2_DECL(LL)                                                                                          \
    int launch_tiled2_fwd_##LL(int ng, const void* x, float* y, float* dy, const Geom& g, const MelDev& mel, bool spec, \
                               const WsView& ws, const LaunchPlan& lp, cudaStream_t st);                              \
    int launch_tiled2_bwd_##LL(int ng, const void* x, const float* y, const float* gy, float* dx, const Geom& g,      \
                               const MelDev& mel, bool spec, const WsView& ws, const LaunchPlan& lp, cudaStream_t st);
DMEL_TILED2_DECL(256)
DMEL_TILED2_DECL(512)
DMEL_TILED2_DECL(1024)
DMEL_TILED2_DECL(2048)
DMEL_TILED2_DECL(4096)
#undef DMEL_TILED2_DECL

}  // namespace dmel
