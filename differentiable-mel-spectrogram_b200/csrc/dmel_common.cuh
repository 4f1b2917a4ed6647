// dmel_common.cuh — workspace / geometry structs, block reductions, shared-memory layout, cp.async staging helpers
// Part of libdmel_b200.so (single translation unit, see dmel_kernels.cu for the overview).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

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
    double jvp_acc;    // sum(dL/dy * dy/d|lambda|) of dmel_lambda_grad_kernel (self-resetting)
    uint32_t jvp_ticket;
    int32_t pad_[3];
};
static_assert(sizeof(WsHeader) == 64, "header is 64 bytes");

// Segment-owner form of the mel projection (second-generation forward).  Every bin k with a non-zero weight feeds the
// mels m0(k) and m0(k)+1 (triangular bank: the bin-major form of dmel.h); the bins with the same m0 are one contiguous
// *segment* s = m0.  Thread `s mod GT` of an FFT group owns segment s (pass s / GT): the split phase writes P[k] into
// slot (base[pass] + j) * (GT+1) + (s mod GT), j the bin's position inside its segment, so that the owner afterwards
// reads its bins — and their two weights from `w`, same indexing — at conflict-free consecutive-lane addresses:
//   mel[s] = sum_j w[.].x * P[.]  (+)  mel[s+1] += sum_j w[.].y * P[.]
// Built ONCE per (bank, n_fft) by dmel_build_segment_tables (dmel.h) into caller-owned device memory that the bank
// struct points to (dmel_melbank.seg_tables); without it the forward falls back to the row gather.
constexpr int SEG_MAX_PASSES = 8;
struct SegHeader {
    int32_t use_seg;                 // 0: bank is not of that shape / tables do not fit -> row-gather fallback
    int32_t npass, nslot, pad_;
    int32_t base[SEG_MAX_PASSES];    // first slot of each pass
    int32_t maxlen[SEG_MAX_PASSES];  // longest segment of each pass
};
static_assert(sizeof(SegHeader) == 80, "SegHeader is 80 bytes");

struct WsView {
    WsHeader* hdr;
    float* win;       // [n_fft]
    float2* tw;       // [n_fft]  exp(-2*pi*i*k/n_fft)
    float* means;     // [B]
    double* clipsum;  // [B]
};

// view of a segment-table blob: SegHeader | (w0, w1) per slot [seg_capacity] | slot of every bin [n_fft/2+1 (+pad)]
// (bins without weights point at the dummy slot nslot*(GT+1))
struct SegView {
    SegHeader* hdr;   // nullptr: no tables
    float2* w;
    int32_t* dest;
};

// slots available for the segment tables: the P values share the group's FFT buffer (fft_core.cuh: padded(L) + 2)
__host__ __device__ inline int seg_capacity(int n_fft) { return n_fft + (n_fft >> 4) + 2; }
static inline size_t seg_blob_bytes(int n_fft) {
    return 96 + ((size_t(seg_capacity(n_fft)) * 8 + 15) & ~size_t(15)) + ((size_t(n_fft / 2 + 4) * 4 + 15) & ~size_t(15));
}
static inline SegView seg_view(void* blob, int n_fft) {
    SegView v{nullptr, nullptr, nullptr};
    if (!blob) return v;
    unsigned char* p = static_cast<unsigned char*>(blob);
    v.hdr = reinterpret_cast<SegHeader*>(p);
    p += 96;
    v.w = reinterpret_cast<float2*>(p);
    p += (size_t(seg_capacity(n_fft)) * 8 + 15) & ~size_t(15);
    v.dest = reinterpret_cast<int32_t*>(p);
    return v;
}

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
    int time_major;              // y / dL/dy / dy are (B,1,T,rows) instead of (B,1,rows,T)
    int x_dtype;                 // DMEL_X_F32 / DMEL_X_I16 / DMEL_X_F16 (16-bit inputs: second-generation tiled kernels only;
                                 // row_aligned then means 8-byte alignment of x and of the row stride)
    int dx_vec4;                 // backward: dx is 16-byte aligned and N a multiple of 4: tiles whose first sample is 4-aligned merge dx four samples at a time
};

struct MelDev {
    const int* mel_k0;
    const int* mel_off;
    const float* mel_w;
    const int* bin_m0;
    const int* bin_off;
    const float* bin_w;
    const float* out_scale;   // per-mel affine folded into the epilogue (dmel.h); both null = identity
    const float* out_shift;
    SegView seg;              // segment tables of the bank (hdr == nullptr: none)
};

// Frames handled by each of the NG groups of a tile holding nfr (<= TT = NG*SF) frames.  Full tiles: SF.  The
// short last tile of a clip is split evenly instead (a multiple of 2 when frames travel as packed pairs), so its
// transforms run side by side on all groups rather than one after the other on the first ones.
template <int NG, bool PAIRS>
__device__ __forceinline__ int tile_frames_per_group(int nfr, int TT, int SF) {
    if (NG == 1 || nfr >= TT) return SF;
    int s = (nfr + NG - 1) / NG;
    if (PAIRS) s = (s + 1) & ~1;
    return s;
}

template <bool V> struct BoolTag { static constexpr bool value = V; };

// forward epilogue: y <- scale[m]*y + shift[m]
__device__ __forceinline__ float affine_out(const MelDev& mel, int m, float v) {
    return fmaf(v, __ldg(mel.out_scale + m), __ldg(mel.out_shift + m));
}
// backward: factor that turns dL/dy' (y' the mapped output) into dL/dmel.  With the log fused,
// log-mel = (y' - shift)/scale and dL/dmel = G * scale * exp(-log-mel); scale == 0 gives 0.
__device__ __forceinline__ float affine_grad_factor(const MelDev& mel, int m, float y_out, bool do_log) {
    const float a = __ldg(mel.out_scale + m);
    if (!do_log) return a;
    const float ia = a != 0.f ? 1.f / a : 0.f;
    return a * __expf(-(y_out - __ldg(mel.out_shift + m)) * ia);
}

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
    if (row_aligned && p0 - shift >= 0 && p0 - shift + 4 * nq <= N) {
        // interior tile: every quad is in range (the quads at both ends may carry up to 3 samples outside the span;
        // they are inside the clip and nobody reads them)
        const float4* src = reinterpret_cast<const float4*>(xb + (p0 - shift));
        float4* dst = reinterpret_cast<float4*>(base);
        for (int q = tid; q < nq; q += NT) cp_async16(dst + q, src + q);
        return;
    }
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
    if (p0 - shift >= 0 && p0 - shift + 4 * nq <= N) {   // interior tile (same test as span_issue: same words per thread)
        float4* d4 = reinterpret_cast<float4*>(xs_sh);
        for (int q = tid; q < nq; q += NT) {
            float4 v = d4[q];
            v.x -= mu; v.y -= mu; v.z -= mu; v.w -= mu;
            d4[q] = v;
        }
        return;
    }
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
#ifdef DMEL_FULL_BWD_CAP
    full_bwd = false;   // experiment: cap the dx-producing backward too
#endif
    return (L <= 512 && NG == 4 && !full_bwd) ? DMEL_MINB_SMALL : (512 / NT > 0 ? 512 / NT : 1);
}

// integer log2 of a power of two
__device__ __forceinline__ int ilog2(int v) { return 31 - __clz(v); }

}  // namespace dmel
