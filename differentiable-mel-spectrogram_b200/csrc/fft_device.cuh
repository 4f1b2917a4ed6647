// fft_device.cuh — in-register / shared-memory complex FFT building blocks (sm_100a).
//
// A group of GT = L/E threads (one warp for L <= 512) computes one L-point complex FFT at
// a time ("packed pair": two real STFT frames ride as the real and imaginary part of one
// complex transform); a CTA runs several such groups independently.  Each thread keeps
// E complex points in registers; a Stockham pass is
//   load (stride L/R)  ->  twiddle  ->  radix-R DFT in registers  ->  scatter (stride NS)
// with the scatter/gather going through ONE shared-memory buffer per group (a group barrier on each side of a
// scatter; for one-warp groups that is __syncwarp).
// The shared-memory index is XOR-swizzled per pass so that both the strided scatter
// of the early passes and the contiguous gather of the next pass are conflict-free
// for 64-bit accesses (16 lanes x 8 B per wavefront).
//
// This replaces the cuFFT/MKL R2C call under torch.stft that the reference reaches
// from time_frequency.py:48/51 (reference: /root/reference/time_frequency.py).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dmel {

// ----------------------------------------------------------------------------
// complex helpers
// ----------------------------------------------------------------------------
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
// multiply by -i  (forward-transform quarter turn)
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }

// ----------------------------------------------------------------------------
// radix-R DFT on registers, natural order in / natural order out, sign = -1
// ----------------------------------------------------------------------------
template <int R> struct RegDft;

template <> struct RegDft<2> {
    static __device__ __forceinline__ void run(float2 (&v)[2]) {
        float2 a = v[0], b = v[1];
        v[0] = cadd(a, b);
        v[1] = csub(a, b);
    }
};

template <> struct RegDft<4> {
    static __device__ __forceinline__ void run(float2 (&v)[4]) {
        float2 s02 = cadd(v[0], v[2]), d02 = csub(v[0], v[2]);
        float2 s13 = cadd(v[1], v[3]), d13 = mul_mi(csub(v[1], v[3]));
        v[0] = cadd(s02, s13);
        v[1] = cadd(d02, d13);
        v[2] = csub(s02, s13);
        v[3] = csub(d02, d13);
    }
};

template <> struct RegDft<8> {
    static __device__ __forceinline__ void run(float2 (&v)[8]) {
        constexpr float C = 0.70710678118654752440f;
        float2 e[4] = {v[0], v[2], v[4], v[6]};
        float2 o[4] = {v[1], v[3], v[5], v[7]};
        RegDft<4>::run(e);
        RegDft<4>::run(o);
        // o[k] *= exp(-2*pi*i*k/8)
        float2 o1 = make_float2(C * (o[1].x + o[1].y), C * (o[1].y - o[1].x));
        float2 o2 = mul_mi(o[2]);
        float2 o3 = make_float2(C * (o[3].y - o[3].x), -C * (o[3].x + o[3].y));
        v[0] = cadd(e[0], o[0]);  v[4] = csub(e[0], o[0]);
        v[1] = cadd(e[1], o1);    v[5] = csub(e[1], o1);
        v[2] = cadd(e[2], o2);    v[6] = csub(e[2], o2);
        v[3] = cadd(e[3], o3);    v[7] = csub(e[3], o3);
    }
};

template <> struct RegDft<16> {
    static __device__ __forceinline__ void run(float2 (&v)[16]) {
        constexpr float C1 = 0.92387953251128675613f;  // cos(pi/8)
        constexpr float S1 = 0.38268343236508977173f;  // sin(pi/8)
        constexpr float C2 = 0.70710678118654752440f;  // cos(pi/4)
        float2 e[8], o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { e[i] = v[2 * i]; o[i] = v[2 * i + 1]; }
        RegDft<8>::run(e);
        RegDft<8>::run(o);
        // w16^k = (cos(k*pi/8), -sin(k*pi/8))
        const float2 w[8] = {
            make_float2(1.f, 0.f),  make_float2(C1, -S1), make_float2(C2, -C2), make_float2(S1, -C1),
            make_float2(0.f, -1.f), make_float2(-S1, -C1), make_float2(-C2, -C2), make_float2(-C1, -S1)};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float2 t;
            if (k == 0) t = o[0];
            else if (k == 4) t = mul_mi(o[4]);
            else t = cmul(o[k], w[k]);
            v[k] = cadd(e[k], t);
            v[k + 8] = csub(e[k], t);
        }
    }
};

// ----------------------------------------------------------------------------
// FFT plans.  A *group* of GT = L/E threads owns one L-point transform; a CTA hosts
// several independent groups.  L <= 512 use a single warp per transform (E up to 16
// points per thread), so their exchanges need only __syncwarp; larger L use 2..8 warps
// and a named barrier private to the group.  prod R = L.
// ----------------------------------------------------------------------------
template <int L> struct Plan;
template <> struct Plan<128>  { static constexpr int E = 4,  NP = 4, R0 = 4,  R1 = 4,  R2 = 4,  R3 = 2; };
template <> struct Plan<256>  { static constexpr int E = 8,  NP = 3, R0 = 8,  R1 = 8,  R2 = 4,  R3 = 1; };
template <> struct Plan<512>  { static constexpr int E = 16, NP = 3, R0 = 8,  R1 = 8,  R2 = 8,  R3 = 1; };
template <> struct Plan<1024> { static constexpr int E = 16, NP = 3, R0 = 16, R1 = 8,  R2 = 8,  R3 = 1; };
template <> struct Plan<2048> { static constexpr int E = 16, NP = 3, R0 = 16, R1 = 16, R2 = 8,  R3 = 1; };
template <> struct Plan<4096> { static constexpr int E = 16, NP = 3, R0 = 16, R1 = 16, R2 = 16, R3 = 1; };
// "big" plans: one CTA per transform, 32 points per thread, twiddles from two 128-entry tables
template <> struct Plan<8192>  { static constexpr int E = 32, NP = 4, R0 = 16, R1 = 16, R2 = 8,  R3 = 4; };
template <> struct Plan<16384> { static constexpr int E = 32, NP = 4, R0 = 16, R1 = 16, R2 = 8,  R3 = 8; };

template <int L, int P> struct PlanRadix;
template <int L> struct PlanRadix<L, 0> { static constexpr int R = Plan<L>::R0, NS = 1; };
template <int L> struct PlanRadix<L, 1> { static constexpr int R = Plan<L>::R1, NS = Plan<L>::R0; };
template <int L> struct PlanRadix<L, 2> { static constexpr int R = Plan<L>::R2, NS = Plan<L>::R0 * Plan<L>::R1; };
template <int L> struct PlanRadix<L, 3> { static constexpr int R = Plan<L>::R3, NS = Plan<L>::R0 * Plan<L>::R1 * Plan<L>::R2; };

// Per-pass twiddle tables, concatenated: pass p >= 1 owns (R_p - 1) rows of NS_p entries,
//   T_p[r-1][k] = exp(-2*pi*i * k*r / (NS_p*R_p)),  k < NS_p
// so a warp reads consecutive k (conflict-free) at an immediate row offset.
template <int L, int P> struct TwOffset {
    static constexpr int value = TwOffset<L, P - 1>::value + (PlanRadix<L, P - 1>::R - 1) * PlanRadix<L, P - 1>::NS;
};
template <int L> struct TwOffset<L, 1> { static constexpr int value = 0; };
template <int L> struct TwOffset<L, 0> { static constexpr int value = 0; };
template <int L> struct TwTotal {
    static constexpr int value = TwOffset<L, Plan<L>::NP - 1>::value +
                                 (PlanRadix<L, Plan<L>::NP - 1>::R - 1) * PlanRadix<L, Plan<L>::NP - 1>::NS;
};

// XOR swizzle of a float2 index inside a 16-entry (128 B) row. `ROT` is chosen per
// pass (see swz_rot) so that the scatter with stride NS lands on 16 distinct 8-byte
// bank pairs per half-warp; rows stay rows, so contiguous gathers stay conflict-free.
template <int ROT> __device__ __forceinline__ int swz(int i) {
    if (ROT == 0) return i;
    return i ^ ((ROT * (i >> 4)) & 15);
}
__host__ __device__ constexpr int swz_rot(int NS, int R) {
    return NS >= 16 ? 0 : ((NS * R >= 16) ? (16 / R) : NS);
}
// compile-time swizzle of a constant offset.  For every plan below, swz(base + r*stride) == swz(base) ^ swz(r*stride)
// (the bit fields of base and r*stride never overlap; checked exhaustively by tests/test_host_logic.py), so a
// butterfly computes the swizzle of its base index once and reaches its R elements with one XOR (ROT != 0) or a
// plain immediate offset (ROT == 0) each.
__host__ __device__ constexpr int swz_const(int rot, int i) { return rot ? (i ^ ((rot * (i >> 4)) & 15)) : i; }

// barrier over the GT threads of one group (bar_id in 1..15 is private to the group)
template <int GT> __device__ __forceinline__ void group_sync(int bar_id) {
    if (GT == 32) {
        __syncwarp();
    } else {
        asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(GT) : "memory");
    }
}

// ----------------------------------------------------------------------------
// One Stockham pass, split in its three phases so callers can place barriers.
//   v[i*R + r]  holds input  j + r*(L/R),  j = gtid + i*GT   (after gather)
//   after compute it holds output index  (j-k)*R + k + r*NS, k = j mod NS
// ----------------------------------------------------------------------------
template <int L, int P> struct Pass {
    static constexpr int E = Plan<L>::E;
    static constexpr int GT = L / E;
    static constexpr int R = PlanRadix<L, P>::R;
    static constexpr int NS = PlanRadix<L, P>::NS;
    static constexpr int NB = E / R;       // butterflies per thread
    static constexpr int STR = L / R;      // gather stride
    static constexpr int ROT_OUT = swz_rot(NS, R);
    // the layout this pass reads is the one the previous pass wrote
    static constexpr int ROT_IN = (P == 0) ? 0 : swz_rot(PlanRadix<L, (P > 0 ? P - 1 : 0)>::NS,
                                                          PlanRadix<L, (P > 0 ? P - 1 : 0)>::R);
    static constexpr int TW_OFF = TwOffset<L, P>::value;

    static __device__ __forceinline__ void gather(float2 (&v)[E], const float2* __restrict__ in, int gtid) {
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const int base = swz<ROT_IN>(gtid + i * GT);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (ROT_IN == 0) v[i * R + r] = in[base + r * STR];
                else v[i * R + r] = in[base ^ swz_const(ROT_IN, r * STR)];
            }
        }
    }

    // twiddle (shared-memory tables, see TwOffset) + radix-R DFT in registers
    static __device__ __forceinline__ void compute(float2 (&v)[E], const float2* __restrict__ twt, int gtid) {
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const int j = gtid + i * GT;
            float2 t[R];
#pragma unroll
            for (int r = 0; r < R; ++r) t[r] = v[i * R + r];
            if (NS > 1) {
                const float2* row = twt + TW_OFF + (j & (NS - 1));
#pragma unroll
                for (int r = 1; r < R; ++r) t[r] = cmul(t[r], row[(r - 1) * NS]);
            }
            RegDft<R>::run(t);
#pragma unroll
            for (int r = 0; r < R; ++r) v[i * R + r] = t[r];
        }
    }

    static __device__ __forceinline__ void scatter(const float2 (&v)[E], float2* __restrict__ out, int gtid) {
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const int j = gtid + i * GT;
            const int k = j & (NS - 1);
            const int base = swz<ROT_OUT>((j - k) * R + k);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (ROT_OUT == 0) out[base + r * NS] = v[i * R + r];
                else out[base ^ swz_const(ROT_OUT, r * NS)] = v[i * R + r];
            }
        }
    }

    // output index held in v[i*R + r] after compute()
    static __device__ __forceinline__ int out_index(int gtid, int i, int r) {
        const int j = gtid + i * GT;
        const int k = j & (NS - 1);
        return (j - k) * R + k + r * NS;
    }
};

// Runs passes P..NP-1 of the plan for one group, exchanging through ONE shared-memory buffer.
// On entry v holds the gathered inputs of pass P.  A group barrier precedes every scatter (all
// earlier readers of `buf` — the previous pass's gathers, or the previous iteration's consumers —
// are done) and follows it.  If STORE_LAST the natural-order spectrum ends in `buf` (linear
// layout: the last pass always has NS >= 16) behind a barrier; otherwise the last pass leaves
// its outputs in v (index map Pass<L,NP-1>::out_index).
template <int L, int P, bool STORE_LAST>
__device__ __forceinline__ void run_passes(float2 (&v)[Plan<L>::E], float2* buf, const float2* __restrict__ twt,
                                           int gtid, int bar_id) {
    constexpr int GT = L / Plan<L>::E;
    Pass<L, P>::compute(v, twt, gtid);
    if constexpr (P + 1 < Plan<L>::NP) {
        group_sync<GT>(bar_id);
        Pass<L, P>::scatter(v, buf, gtid);
        group_sync<GT>(bar_id);
        Pass<L, P + 1>::gather(v, buf, gtid);
        run_passes<L, P + 1, STORE_LAST>(v, buf, twt, gtid, bar_id);
    } else if constexpr (STORE_LAST) {
        group_sync<GT>(bar_id);
        Pass<L, P>::scatter(v, buf, gtid);
        group_sync<GT>(bar_id);
    }
}

// ----------------------------------------------------------------------------
// Big transforms (8192, 16384 points): the per-pass twiddle tables would not fit in shared
// memory, so exp(-2*pi*i*idx/L) is formed from two 128-entry tables, idx = hi*128 + lo
// (one extra complex multiply, two roundings).  The whole CTA is the group.
// ----------------------------------------------------------------------------
__device__ __forceinline__ float2 tw_split(const float2* __restrict__ t_hi, const float2* __restrict__ t_lo, int idx) {
    return cmul(t_hi[idx >> 7], t_lo[idx & 127]);
}

template <int L, int P>
__device__ __forceinline__ void big_compute(float2 (&v)[Plan<L>::E], const float2* __restrict__ t_hi,
                                            const float2* __restrict__ t_lo, int tid) {
    using PS = Pass<L, P>;
#pragma unroll
    for (int i = 0; i < PS::NB; ++i) {
        const int j = tid + i * PS::GT;
        float2 t[PS::R];
#pragma unroll
        for (int r = 0; r < PS::R; ++r) t[r] = v[i * PS::R + r];
        if (PS::NS > 1) {
            const int base = (j & (PS::NS - 1)) * (L / (PS::NS * PS::R));
#pragma unroll
            for (int r = 1; r < PS::R; ++r) t[r] = cmul(t[r], tw_split(t_hi, t_lo, base * r));
        }
        RegDft<PS::R>::run(t);
#pragma unroll
        for (int r = 0; r < PS::R; ++r) v[i * PS::R + r] = t[r];
    }
}

// same contract as run_passes, with __syncthreads as the group barrier
template <int L, int P, bool STORE_LAST>
__device__ __forceinline__ void big_passes(float2 (&v)[Plan<L>::E], float2* buf, const float2* __restrict__ t_hi,
                                           const float2* __restrict__ t_lo, int tid) {
    big_compute<L, P>(v, t_hi, t_lo, tid);
    if constexpr (P + 1 < Plan<L>::NP) {
        __syncthreads();
        Pass<L, P>::scatter(v, buf, tid);
        __syncthreads();
        Pass<L, P + 1>::gather(v, buf, tid);
        big_passes<L, P + 1, STORE_LAST>(v, buf, t_hi, t_lo, tid);
    } else if constexpr (STORE_LAST) {
        __syncthreads();
        Pass<L, P>::scatter(v, buf, tid);
        __syncthreads();
    }
}

}  // namespace dmel
