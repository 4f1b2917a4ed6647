// fft_core.cuh — second-generation in-register / shared-memory complex FFT (sm_100a), n_fft 256..4096.
//
// What changed against fft_device.cuh (which the n_fft 128 and 8192/16384 kernels keep using), and why — the
// round-1 ncu captures showed the tiled kernels bound by instruction issue and the shared-memory pipe, with more
// integer than floating-point instructions (profiles/r01w_ncu_*.txt):
//   * every thread owns 16 points of an L-point transform (GT = L/16 threads per transform), always the SAME
//     index set on the way in and on the way out:  n = gtid + GT*q, q = 0..15.  So window values, derivative-window
//     values and per-point accumulators live in registers for the whole kernel instead of being re-read per frame;
//   * the first radix is 16, so that every exchange is either "16 consecutive points per thread" (pass 0) or a
//     run of 16 consecutive lanes (later passes).  With ONE padding word per 16 float2 entries,
//     phys(i) = i + (i >> 4), both patterns are bank-conflict-free for 64-bit accesses, and — unlike an XOR
//     swizzle — every address is `per-thread base + compile-time immediate`: no integer instruction per access;
//   * pass-1 twiddles depend on (gtid mod 16) only: 7 complex registers when the radix is 8; the last pass reads
//     its table rows at immediate offsets from one base;
//   * radix-8/16 butterflies fold the sqrt(1/2) rotations into FFMAs.
//
// This replaces the cuFFT/MKL R2C call under torch.stft that the reference reaches from
// /root/reference/time_frequency.py:48/51.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "f32x2.cuh"
#include "fft_device.cuh"   // float2 helpers, RegDft<2/4>, group_sync

namespace dmel {
namespace v2 {

// barrier over the GT threads of one transform group: a (partial-)warp barrier for up to one warp, a named barrier
// private to the group otherwise
// `sync_arg` (see group_sync_arg): the lane mask of the group for GT <= 32, the named-barrier id otherwise
template <int GT> __device__ __forceinline__ void gsync(int sync_arg) {
    if constexpr (GT <= 32) __syncwarp(unsigned(sync_arg));
    else asm volatile("bar.sync %0, %1;" ::"r"(sync_arg), "n"(GT) : "memory");
}
// GT = 16: two groups share a warp and may run different numbers of transforms (short last tiles), so each
// synchronises its own half only
template <int GT> __device__ __forceinline__ int group_sync_arg(int grp) {
    if constexpr (GT > 32) return 1 + grp;
    else if constexpr (GT == 32) return -1;
    else return int(((1u << GT) - 1u) << ((grp * GT) & 31));
}

// ----------------------------------------------------------------------------
// radix butterflies, forward sign, natural order in/out
// ----------------------------------------------------------------------------
__device__ __forceinline__ void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
    const float2 s02 = px::add2(a0, a2), d02 = px::sub2(a0, a2);
    const float2 s13 = px::add2(a1, a3), d13 = px::sub2(a1, a3);
    a0 = px::add2(s02, s13);
    a2 = px::sub2(s02, s13);
    a1 = px::add_mi(d02, d13);   // d02 + d13*(-i)
    a3 = px::sub_mi(d02, d13);
}

// e +- o*W8^1 = e +- C*(o.x+o.y, o.y-o.x);   e +- o*W8^3 = e +- C*(o.y-o.x, -(o.x+o.y))
__device__ __forceinline__ void bf_w8_1(const float2 e, const float2 o, float2& lo, float2& hi) {
    constexpr float C = 0.70710678118654752440f;
    const float2 u = px::fma2(px::swap(o), px::pm(), o);
    lo = px::fma2(px::bc(C), u, e);
    hi = px::fma2(px::bc(-C), u, e);
}
__device__ __forceinline__ void bf_w8_3(const float2 e, const float2 o, float2& lo, float2& hi) {
    constexpr float C = 0.70710678118654752440f;
    const float2 u = px::fma2(px::swap(o), px::pm(), px::neg(o));
    lo = px::fma2(px::bc(C), u, e);
    hi = px::fma2(px::bc(-C), u, e);
}
// e +- o*(-i)
__device__ __forceinline__ void bf_mi(const float2 e, const float2 o, float2& lo, float2& hi) {
    lo = px::add_mi(e, o);
    hi = px::sub_mi(e, o);
}
// e +- o*(wr + i*wi), compile-time constant
__device__ __forceinline__ void bf_w(const float2 e, const float2 o, const float wr, const float wi, float2& lo,
                                     float2& hi) {
    px::bf_const(e, o, wr, wi, lo, hi);
}

__device__ __forceinline__ void dft8(float2 (&v)[8]) {
    float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
    float2 o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
    dft4(e0, e1, e2, e3);
    dft4(o0, o1, o2, o3);
    v[0] = px::add2(e0, o0);
    v[4] = px::sub2(e0, o0);
    bf_w8_1(e1, o1, v[1], v[5]);
    bf_mi(e2, o2, v[2], v[6]);
    bf_w8_3(e3, o3, v[3], v[7]);
}

__device__ __forceinline__ void dft16(float2 (&v)[16]) {
    constexpr float C1 = 0.92387953251128675613f;  // cos(pi/8)
    constexpr float S1 = 0.38268343236508977173f;  // sin(pi/8)
    float2 e[8], o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { e[i] = v[2 * i]; o[i] = v[2 * i + 1]; }
    dft8(e);
    dft8(o);
    v[0] = px::add2(e[0], o[0]);
    v[8] = px::sub2(e[0], o[0]);
    bf_w(e[1], o[1], C1, -S1, v[1], v[9]);
    bf_w8_1(e[2], o[2], v[2], v[10]);
    bf_w(e[3], o[3], S1, -C1, v[3], v[11]);
    bf_mi(e[4], o[4], v[4], v[12]);
    bf_w(e[5], o[5], -S1, -C1, v[5], v[13]);
    bf_w8_3(e[6], o[6], v[6], v[14]);
    bf_w(e[7], o[7], -C1, -S1, v[7], v[15]);
}

// v * exp(-2*pi*i*M/16), M a compile-time constant: the rotations between the butterflies of one thread in the last pass
template <int M> __device__ __forceinline__ float2 mul_w16(const float2 v) {
    constexpr float C1 = 0.92387953251128675613f, S1 = 0.38268343236508977173f, C2 = 0.70710678118654752440f;
    constexpr int m = M & 15;
    if constexpr (m == 0) return v;
    else if constexpr (m == 4) return px::mul_mi(v);
    else if constexpr (m == 8) return px::neg(v);
    else if constexpr (m == 12) return px::neg(px::mul_mi(v));
    else {
        // (wr, wi) = (cos, -sin)(2*pi*m/16)
        constexpr float wr = (m == 1 || m == 15) ? C1 : (m == 2 || m == 14) ? C2 : (m == 3 || m == 13) ? S1
                           : (m == 5 || m == 11) ? -S1 : (m == 6 || m == 10) ? -C2 : -C1;
        constexpr float wi = (m == 1 || m == 7) ? -S1 : (m == 2 || m == 6) ? -C2 : (m == 3 || m == 5) ? -C1
                           : (m == 9 || m == 15) ? S1 : (m == 10 || m == 14) ? C2 : C1;
        return px::cmul_const(v, wr, wi);
    }
}

template <int I, int N, typename Fn> __device__ __forceinline__ void static_for(Fn&& fn) {
    if constexpr (I < N) {
        fn(std::integral_constant<int, I>{});
        static_for<I + 1, N>(fn);
    }
}

template <int R> __device__ __forceinline__ void dftR(float2 (&v)[R]) {
    if constexpr (R == 16) dft16(v);
    else if constexpr (R == 8) dft8(v);
    else if constexpr (R == 4) dft4(v[0], v[1], v[2], v[3]);
    else if constexpr (R == 2) { const float2 a = v[0], b = v[1]; v[0] = px::add2(a, b); v[1] = px::sub2(a, b); }
}

// ----------------------------------------------------------------------------
// plans: L = 16 * R1 * R2 (R2 == 1: two passes)
// ----------------------------------------------------------------------------
template <int L> struct Plan2;
template <> struct Plan2<256>  { static constexpr int R1 = 16, R2 = 1; };
template <> struct Plan2<512>  { static constexpr int R1 = 8,  R2 = 4; };
template <> struct Plan2<1024> { static constexpr int R1 = 8,  R2 = 8; };
template <> struct Plan2<2048> { static constexpr int R1 = 8,  R2 = 16; };
template <> struct Plan2<4096> { static constexpr int R1 = 16, R2 = 16; };

__host__ __device__ constexpr int padded(int i) { return i + (i >> 4); }

template <int L> struct Fft {
    static constexpr int E = 16;
    static constexpr int GT = L / 16;                 // threads per transform
    static constexpr int R1 = Plan2<L>::R1, R2 = Plan2<L>::R2;
    static constexpr int NP = (R2 == 1) ? 2 : 3;
    static constexpr int NS2 = 16 * R1;               // = L / R2
    static constexpr int BUF = padded(L) + 2;         // float2 entries per group buffer (even: keeps 16 B alignment)
    static constexpr bool TW1_REGS = (R1 == 8);       // pass-1 twiddles in registers (7 complex)
    // twiddle tables as the prep kernel writes them (fft_device.cuh layout): pass 1: (R1-1) rows of 16; pass 2:
    // (R2-1) rows of NS2
    static constexpr int TW1_N = (R1 - 1) * 16;
    static constexpr int TW2_N = (R2 == 1) ? 0 : (R2 - 1) * NS2;
    static constexpr int TW_TOTAL = TW1_N + TW2_N;
    // what the kernels keep in shared memory
    // Last-pass twiddles W_L^{(gtid + i*GT) r} = W_L^{gtid*r} * W_16^{i*r}: the first factor is a per-thread constant
    // (registers), the second a compile-time constant.  Shared-memory bandwidth, not arithmetic, bounds these kernels,
    // so the 2..14 table loads per transform are traded for a few multiplies.  R2 = 4: three base factors; R2 = 8:
    // r = 1, 2, 4 are kept and the other four are formed per transform (4 more complex multiplies).
    static constexpr bool TW2_REGS = (R2 == 4 || R2 == 8);
    static constexpr int TW2_NB = TW2_REGS ? (R2 == 4 ? 3 : 3) : 1;
    static constexpr int TW_SMEM = (TW1_REGS ? 0 : TW1_N) + (TW2_REGS ? 0 : TW2_N);

    struct Ctx {
        float2* g;            // buf + padded(gtid): gather base of passes 1, 2
        float2* s0;           // buf + 17*gtid: scatter base of pass 0
        float2* s1;           // buf + 17*R1*(gtid>>4) + (gtid&15): scatter base of pass 1
        const float2* t1;     // shared-memory pass-1 table + (gtid&15)   (unused when TW1_REGS)
        const float2* t2;     // shared-memory pass-2 table + gtid
        float2 tw1[TW1_REGS ? 7 : 1];
        float2 tw2b[TW2_NB];  // TW2_REGS: W_L^{gtid*r} for r = 1,2,3 (R2 = 4) or r = 1,2,4 (R2 = 8)
        int bar_id;
    };

    // tw_smem: the staged tables (TW_SMEM entries); tw_global: the full table in global memory (prep kernel)
    static __device__ __forceinline__ void init(Ctx& c, float2* buf, const float2* tw_smem, const float2* __restrict__ tw_global,
                                                int gtid, int bar_id) {
        c.g = buf + padded(gtid);
        c.s0 = buf + 17 * gtid;
        c.s1 = buf + 17 * R1 * (gtid >> 4) + (gtid & 15);
        c.t1 = tw_smem + (gtid & 15);
        c.t2 = tw_smem + (TW1_REGS ? 0 : TW1_N) + gtid;
        c.bar_id = bar_id;
        if constexpr (TW2_REGS) {
            const float2* t2g = tw_global + TW1_N + gtid;      // pass-2 table rows: r-1
            c.tw2b[0] = __ldg(t2g);
            c.tw2b[1] = __ldg(t2g + NS2);
            c.tw2b[2] = __ldg(t2g + (R2 == 4 ? 2 : 3) * NS2);
        }
        if constexpr (TW1_REGS) {
#pragma unroll
            for (int r = 1; r < 8; ++r) c.tw1[r - 1] = __ldg(tw_global + (r - 1) * 16 + (gtid & 15));
        }
    }

    // In:  z[q] = point gtid + GT*q.  Out: z[q] = bin gtid + GT*q.  The group buffer may still be read by the
    // group's previous consumer on entry (a barrier precedes the first scatter); on exit all gathers of this
    // thread are done but other threads may still be gathering: callers barrier before writing the buffer again.
    static __device__ __forceinline__ void run(float2 (&z)[16], const Ctx& c) {
        // ---- pass 0: radix 16 over stride GT (no twiddles), outputs 16*gtid + r
        dft16(z);
        gsync<GT>(c.bar_id);
#pragma unroll
        for (int r = 0; r < 16; ++r) c.s0[r] = z[r];
        gsync<GT>(c.bar_id);
        // ---- pass 1: radix R1, NS = 16
        constexpr int NB1 = 16 / R1, S1 = L / R1;
        float2 y[16];
#pragma unroll
        for (int i = 0; i < NB1; ++i) {
            float2 t[R1];
#pragma unroll
            for (int r = 0; r < R1; ++r) t[r] = c.g[padded(i * GT + r * S1)];
#pragma unroll
            for (int r = 1; r < R1; ++r) {
                if constexpr (TW1_REGS) t[r] = px::cmul(t[r], c.tw1[r - 1]);
                else t[r] = px::cmul(t[r], c.t1[(r - 1) * 16]);
            }
            dftR<R1>(t);
            if constexpr (NP == 2) {
                // last pass: natural order, bin gtid + 16*r  (GT == 16)
#pragma unroll
                for (int r = 0; r < R1; ++r) z[r] = t[r];
            } else {
#pragma unroll
                for (int r = 0; r < R1; ++r) y[i * R1 + r] = t[r];
            }
        }
        if constexpr (NP == 3) {
            gsync<GT>(c.bar_id);
#pragma unroll
            for (int i = 0; i < NB1; ++i)
#pragma unroll
                for (int r = 0; r < R1; ++r) c.s1[17 * R1 * (i * GT / 16) + 17 * r] = y[i * R1 + r];
            gsync<GT>(c.bar_id);
            // ---- pass 2: radix R2, NS = 16*R1 = L/R2; outputs gtid + GT*(i + r*NB2)
            constexpr int NB2 = 16 / R2, S2 = L / R2;
            float2 bs[R2];   // TW2_REGS: bs[r] = W_L^{gtid*r}
            if constexpr (TW2_REGS && R2 == 4) {
                bs[1] = c.tw2b[0]; bs[2] = c.tw2b[1]; bs[3] = c.tw2b[2];
            } else if constexpr (TW2_REGS && R2 == 8) {
                bs[1] = c.tw2b[0]; bs[2] = c.tw2b[1]; bs[4] = c.tw2b[2];
                bs[3] = px::cmul(bs[1], bs[2]); bs[5] = px::cmul(bs[1], bs[4]);
                bs[6] = px::cmul(bs[2], bs[4]); bs[7] = px::cmul(bs[3], bs[4]);
            }
            static_for<0, NB2>([&](auto ic) {
                constexpr int i = decltype(ic)::value;
                float2 t[R2];
#pragma unroll
                for (int r = 0; r < R2; ++r) t[r] = c.g[padded(i * GT + r * S2)];
                if constexpr (TW2_REGS) {
                    static_for<1, R2>([&](auto rc) {
                        constexpr int r = decltype(rc)::value;
                        t[r] = px::cmul(mul_w16<i * r>(t[r]), bs[r]);
                    });
                } else {
#pragma unroll
                    for (int r = 1; r < R2; ++r) t[r] = px::cmul(t[r], c.t2[(r - 1) * NS2 + i * GT]);
                }
                dftR<R2>(t);
#pragma unroll
                for (int r = 0; r < R2; ++r) z[i + r * NB2] = t[r];
            });
        }
    }
};

}  // namespace v2
}  // namespace dmel
