// dmel_small_kernels.cuh — prep (means, window, twiddles), zero-init and finalize kernels
// Part of libdmel_b200.so (single translation unit, see dmel_kernels.cu for the overview).
#pragma once
#include <cooperative_groups.h>
#include <cuda_fp16.h>

#include "dmel_common.cuh"

namespace dmel {
namespace cg = cooperative_groups;

// ----------------------------------------------------------------------------
// prep: means + window + twiddles
// ----------------------------------------------------------------------------
constexpr int PREP_THREADS = 256;

// host mirror of fft_device.cuh's Plan<L> (np == 0: no FFT plan, direct-DFT path)
struct PlanDesc {
    int np;
    int r[4];
};

// A clip's mean is summed by a thread-block cluster of `prep_segments(N)` CTAs, one segment of the clip each; the
// segment totals meet in CTA 0 of the cluster over distributed shared memory and are added in segment order.  The
// split depends on N alone, so a clip's mean (hence its spectrogram) is bit-identical whatever batch it is part of.
constexpr int PREP_SEG_POINTS = 32768;
constexpr int PREP_MAX_SEGMENTS = 8;   // portable cluster size
__host__ __device__ inline int prep_segments(int N) {
    const int s = (N + PREP_SEG_POINTS - 1) / PREP_SEG_POINTS;
    return s < 1 ? 1 : (s > PREP_MAX_SEGMENTS ? PREP_MAX_SEGMENTS : s);
}

// Segment tables of the second-generation forward (SegHeader in dmel_common.cuh), built by one CTA from the bin-major
// bank.  gt = threads per FFT group (n_fft/16); gt == 0: not wanted (spectrogram mode, other kernel families).
constexpr int SEG_MAX_MELS = 1024;
__global__ void __launch_bounds__(PREP_THREADS)
dmel_segment_kernel(MelDev mel, int n_mels, int n_fft, int gt, SegView sv) {
    __shared__ int seg_start[SEG_MAX_MELS], seg_len[SEG_MAX_MELS];
    __shared__ int pass_max[SEG_MAX_PASSES], pass_base[SEG_MAX_PASSES + 1];
    __shared__ int bad;
    const int tid = threadIdx.x, F = n_fft / 2 + 1;
    const int npass = gt > 0 ? (n_mels + gt - 1) / gt : 0;
    const int cap = seg_capacity(n_fft), pitch = gt + 1;
    if (gt <= 0 || n_mels <= 0 || n_mels > SEG_MAX_MELS || npass > SEG_MAX_PASSES || mel.bin_m0 == nullptr) {
        if (tid == 0) sv.hdr->use_seg = 0;
        return;
    }
    if (tid == 0) bad = 0;
    for (int m = tid; m < n_mels; m += PREP_THREADS) { seg_start[m] = -1; seg_len[m] = 0; }
    if (tid < SEG_MAX_PASSES) pass_max[tid] = 0;
    __syncthreads();
    auto seg_of = [&](int k) {   // segment of bin k, -1 if the bin carries no weight
        const int o0 = __ldg(mel.bin_off + k), cnt = __ldg(mel.bin_off + k + 1) - o0;
        if (cnt > 2) bad = 1;
        return cnt > 0 ? __ldg(mel.bin_m0 + k) : -1;
    };
    // segment starts; a segment that starts twice (weights interrupted by an empty bin) or a bank whose first mels do
    // not increase with frequency is not of the expected shape
    for (int k = tid; k < F; k += PREP_THREADS) {
        const int sk = seg_of(k);
        if (sk >= n_mels) bad = 1;
        if (sk >= 0 && sk < n_mels) {
            const int sp = k > 0 ? seg_of(k - 1) : -1;
            if (sp != sk) {
                if (atomicCAS(&seg_start[sk], -1, k) != -1) bad = 1;
            }
            if (sp > sk) bad = 1;
        }
    }
    __syncthreads();
    for (int k = tid; k < F; k += PREP_THREADS) {
        const int sk = seg_of(k);
        if (sk >= 0 && sk < n_mels && seg_start[sk] >= 0) {
            const int j = k - seg_start[sk];
            if (j < 0) bad = 1;
            else atomicMax(&seg_len[sk], j + 1);
        }
    }
    __syncthreads();
    for (int m = tid; m < n_mels; m += PREP_THREADS) atomicMax(&pass_max[m / gt], seg_len[m]);
    __syncthreads();
    if (tid == 0) {
        int o = 0;
        for (int i = 0; i < npass; ++i) { pass_base[i] = o; o += pass_max[i]; }
        pass_base[npass] = o;
        // slots [0, nslot*pitch) hold the segments, slot nslot*pitch is the dummy bins without weights are written to
        const int use = (!bad && o * pitch + 1 <= cap) ? 1 : 0;
        sv.hdr->use_seg = use;
        sv.hdr->npass = npass;
        sv.hdr->nslot = o;
        sv.hdr->pad_ = 0;
        for (int i = 0; i < SEG_MAX_PASSES; ++i) {
            sv.hdr->base[i] = i < npass ? pass_base[i] : 0;
            sv.hdr->maxlen[i] = i < npass ? pass_max[i] : 0;
        }
        if (!use) bad = 1;
    }
    __syncthreads();
    if (bad) return;
    const int nslot = pass_base[npass];
    for (int i = tid; i < cap; i += PREP_THREADS) sv.w[i] = make_float2(0.f, 0.f);
    __syncthreads();
    for (int k = tid; k < F; k += PREP_THREADS) {
        const int o0 = __ldg(mel.bin_off + k), cnt = __ldg(mel.bin_off + k + 1) - o0;
        int dest = nslot * pitch;
        if (cnt > 0) {
            const int sk = __ldg(mel.bin_m0 + k);
            const int slot = (pass_base[sk / gt] + (k - seg_start[sk])) * pitch + (sk % gt);
            sv.w[slot] = make_float2(__ldg(mel.bin_w + o0), cnt > 1 ? __ldg(mel.bin_w + o0 + 1) : 0.f);
            dest = slot;
        }
        sv.dest[k] = dest;
    }
}

// grid = (B + 1) clusters of prep_segments(N) CTAs: cluster b < B -> mean of clip b; cluster B -> window + twiddles
__global__ void __launch_bounds__(PREP_THREADS)
dmel_prep_kernel(const void* __restrict__ x_any, long long x_stride, int B, int N, float lam_abs, int win_length,
                 int n_fft, int flags, WsView ws, PlanDesc plan, const float* __restrict__ means_src) {
    const float* x = static_cast<const float*>(x_any);
    __shared__ double red[32];
    __shared__ double seg_total;
    const int tid = threadIdx.x;
    const int normalize = flags & DMEL_FLAG_NORMALIZE_WINDOW;
    const int nseg = prep_segments(N);
    const int clip = int(blockIdx.x) / nseg, seg = int(blockIdx.x) - clip * nseg;
    if (clip < B) {
        if (flags & DMEL_FLAG_KEEP_MEAN) {
            if (tid == 0 && seg == 0) ws.means[clip] = 0.f;
            return;
        }
        if (means_src) {   // dmel_forward_multi_f32: the means of this batch already exist (they do not depend on lambda)
            if (tid == 0 && seg == 0) ws.means[clip] = __ldg(means_src + clip);
            return;
        }
        // models.py:38  x[idx] - torch.mean(x[idx])
        const bool x16 = (flags & (DMEL_FLAG_X_I16 | DMEL_FLAG_X_F16)) != 0;
        const int seg_len = x16 ? ((((N + nseg - 1) / nseg) + 7) & ~7)
                                : ((((N + nseg - 1) / nseg) + 3) & ~3);   // segments keep x's 16-byte alignment
        const int lo = min(seg * seg_len, N), hi = min(lo + seg_len, N);
        const float* xb = x + (long long)clip * x_stride + lo;
        const int n = hi - lo;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        int i0 = 0;
        if (x16) {
            // 16-bit samples (int16 PCM or half), widened exactly like the staging of the main kernels
            const bool is_i16 = (flags & DMEL_FLAG_X_I16) != 0;
            const unsigned short* xh = static_cast<const unsigned short*>(x_any) + (long long)clip * x_stride + lo;
            auto wide = [&](unsigned v) {
                return is_i16 ? float(short(v & 0xffffu)) : __half2float(__ushort_as_half((unsigned short)(v & 0xffffu)));
            };
            if ((reinterpret_cast<uintptr_t>(xh) & 15) == 0) {
                const uint4* x8 = reinterpret_cast<const uint4*>(xh);
                const int n8 = n >> 3;
                for (int q = tid; q < n8; q += PREP_THREADS) {
                    const uint4 v = __ldg(x8 + q);
                    acc[0] += ((wide(v.x) + wide(v.x >> 16)) + (wide(v.y) + wide(v.y >> 16))) +
                                  ((wide(v.z) + wide(v.z >> 16)) + (wide(v.w) + wide(v.w >> 16)));
                }
                i0 = n8 << 3;
            }
            for (int i = i0 + tid; i < n; i += PREP_THREADS) acc[1] += wide(__ldg(xh + i));
        } else if ((reinterpret_cast<uintptr_t>(xb) & 15) == 0) {
            // 16-byte loads, 8 independent ones in flight per thread
            const float4* x4 = reinterpret_cast<const float4*>(xb);
            const int n4 = n >> 2;
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
        if (!x16)
            for (int i = i0 + tid; i < n; i += PREP_THREADS) acc[1] += __ldg(xb + i);
        double part = 0.0;
#pragma unroll
        for (int u = 0; u < 8; ++u) part += double(acc[u]);
        double tot = block_sum(part, red);
        if (nseg > 1) {
            cg::cluster_group cluster = cg::this_cluster();
            if (tid == 0) seg_total = tot;
            cluster.sync();
            if (seg == 0 && tid == 0) {
                tot = 0.0;
                for (int r = 0; r < nseg; ++r) tot += *cluster.map_shared_rank(&seg_total, r);
            }
            cluster.sync();     // a CTA's shared memory must outlive the remote reads of it
        }
        if (tid == 0 && seg == 0) ws.means[clip] = float(tot / double(N));
        return;
    }
    if (seg != 0) return;
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
        ws.hdr->jvp_acc = 0.0;
        ws.hdr->jvp_ticket = 0u;
    }
    __syncthreads();
    if (normalize) {
        const float sc = sh_scale;
        for (int n = tid; n < n_fft; n += PREP_THREADS) ws.win[n] *= sc;
    }
}

// ----------------------------------------------------------------------------
// One-shot all-reduce (sum) of ONE fp32 over NVLink peer memory — the path's only collective (the scalar
// dL/dlambda).  Every rank owns a symmetric buffer, IPC-mapped into all peers:
//     float slots[2][world] | int flags[2][world] | int counter | int error
// Rank r stores its value into slots[parity][r] of EVERY peer, fences, then publishes flags[parity][r] = epoch
// with a system-scope release; it then acquires its own flags[parity][*] and sums its own slots in rank order,
// so all ranks obtain the bit-identical total.  `epoch` is a device-side call counter (identical on all ranks, and
// it keeps advancing under CUDA-graph replay); the parity double-buffering is enough because a peer cannot finish
// call e+1 before this rank has published call e+1, i.e. after it has read call e.  Spins are bounded: on a
// timeout the result is NaN and the error word is set (the host falls back to NCCL).
// ----------------------------------------------------------------------------
constexpr int PEER_MAX_WORLD = 32;

__device__ __forceinline__ void st_release_sys(int* p, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// executed by one full warp; returns the total in every lane (NaN + error word on a peer timeout)
__device__ __forceinline__ float peer_exchange_sum(float v, const long long* __restrict__ peer_bases, int rank, int world,
                                                   long long timeout_cycles) {
    const int lane = threadIdx.x & 31;
    int* mine = reinterpret_cast<int*>(peer_bases[rank]);
    int epoch = 0;
    if (lane == 0) epoch = ++mine[4 * world];          // local call counter (only this code touches it)
    epoch = __shfl_sync(0xffffffffu, epoch, 0);
    const int parity = epoch & 1;
    if (lane < world) {
        int* peer = reinterpret_cast<int*>(peer_bases[lane]);
        reinterpret_cast<volatile float*>(peer)[parity * world + rank] = v;
        __threadfence_system();
        st_release_sys(peer + 2 * world + parity * world + rank, epoch);
    }
    float got = 0.f;
    bool ok = true;
    if (lane < world) {
        const int* flag = mine + 2 * world + parity * world + lane;
        const long long t0 = clock64();
        while (ld_acquire_sys(flag) != epoch) {
            if (clock64() - t0 > timeout_cycles) { ok = false; break; }
        }
        got = reinterpret_cast<volatile float*>(mine)[parity * world + lane];
    }
    ok = __all_sync(0xffffffffu, ok);
    float total = 0.f;
    for (int r = 0; r < world; ++r) total += __shfl_sync(0xffffffffu, got, r);   // same order on every rank
    if (!ok) {
        if (lane == 0) mine[4 * world + 1] = 1;
        total = __int_as_float(0x7fc00000);
    }
    return total;
}

// ----------------------------------------------------------------------------
// finalize: dx -= mean(dx) per clip; dL/dlambda -> fp32 with sign(lambda)
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dmel_finalize_kernel(float* __restrict__ dx, int B, int N, int chunks_per_clip, const double* __restrict__ clipsum,
                     const WsHeader* hdr, float sign, float* __restrict__ grad_lambda,
                     const long long* __restrict__ peer_bases, int rank, int world, long long timeout_cycles) {
    if (blockIdx.x == 0 && threadIdx.x < 32) {
        // local dL/dlambda; with a peer ring the sum over the GPUs happens right here (compute + collective
        // in one kernel: the other blocks meanwhile apply the mean-removal adjoint to dx)
        float v = float(double(sign) * hdr->dlam_acc);
        if (peer_bases != nullptr) v = peer_exchange_sum(v, peer_bases, rank, world, timeout_cycles);
        if (threadIdx.x == 0) *grad_lambda = v;
    }
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

// Backward of the JVP forward: dL/dlambda = sign(lambda) * sum(dL/dy * dy/d|lambda|).  One launch: every block adds
// its fp64 partial, the last block to finish (ticket) converts the total and resets the accumulator for the next call.
__global__ void __launch_bounds__(256)
dmel_lambda_grad_kernel(const float* __restrict__ gy, const float* __restrict__ dy, long long n, float sign,
                        WsHeader* hdr, float* __restrict__ grad_lambda) {
    __shared__ double red[32];
    __shared__ bool last;
    const long long gid = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long nthreads = (long long)gridDim.x * 256;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    const long long n4 = ((reinterpret_cast<uintptr_t>(gy) | reinterpret_cast<uintptr_t>(dy)) & 15) == 0 ? (n >> 2) : 0;
    const float4* g4 = reinterpret_cast<const float4*>(gy);
    const float4* d4 = reinterpret_cast<const float4*>(dy);
    for (long long i = gid; i < n4; i += nthreads) {
        const float4 a = __ldg(g4 + i), b = __ldg(d4 + i);
        a0 = fmaf(a.x, b.x, a0); a1 = fmaf(a.y, b.y, a1); a2 = fmaf(a.z, b.z, a2); a3 = fmaf(a.w, b.w, a3);
    }
    for (long long i = (n4 << 2) + gid; i < n; i += nthreads) a0 = fmaf(__ldg(gy + i), __ldg(dy + i), a0);
    const double tot = block_sum((double(a0) + double(a1)) + (double(a2) + double(a3)), red);
    if (threadIdx.x == 0) {
        atomicAdd(&hdr->jvp_acc, tot);
        __threadfence();
        last = atomicAdd(&hdr->jvp_ticket, 1u) == gridDim.x - 1;
        if (last) {
            __threadfence();
            const double total = atomicAdd(&hdr->jvp_acc, 0.0);   // read through L2
            *grad_lambda = float(double(sign) * total);
            hdr->jvp_acc = 0.0;
            hdr->jvp_ticket = 0u;
        }
    }
}

__global__ void __launch_bounds__(32)
dmel_peer_allreduce_kernel(float* __restrict__ value, const long long* __restrict__ peer_bases, int rank, int world,
                           long long timeout_cycles) {
    const float total = peer_exchange_sum(*value, peer_bases, rank, world, timeout_cycles);
    if (threadIdx.x == 0) *value = total;
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

}  // namespace dmel
