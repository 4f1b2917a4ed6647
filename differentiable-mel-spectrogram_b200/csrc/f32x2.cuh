// f32x2.cuh — packed single-precision pairs: the sm_100a FADD2 / FMUL2 / FFMA2 instructions (PTX add/mul/fma.rn.f32x2).
//
// One warp-instruction does the arithmetic of two: a complex add is one FADD2, a complex multiply by a twiddle is
// three packed operations instead of four scalar ones, a butterfly with a constant rotation is two FFMA2 per output.
// A float2 lives in an aligned register pair; ptxas folds `mov.b64 {lo,hi}` into operand selectors
// (R.F32x2.HI_LO, .LO_HI = swapped, R.F32 = one register broadcast to both halves), so swap() and bc() below cost no
// instruction.
//
// MEASURED AND NOT ADOPTED (profiles/r02/r02m_f32x2_packed_vs_scalar.log, B200): with the whole FFT core, the window
// multiply, the packed-pair split and the conj(V) formation on packed instructions the kernels' floating-point
// instruction count halves (cuobjdump: FADD 201+FFMA 183+FMUL 144 scalar -> 436 packed in fwd<512>), yet the step time
// moves by -2 % (n_fft 512, 2048) to +5 % (n_fft 1024): a packed instruction occupies the FP32 pipe for both halves,
// so the throughput per issue slot is what the scalar instructions already had, and a rotation needs a sign pattern
// the scalar FFMA gets for free.  The scalar forms are therefore the default; -DDMEL_F32X2 compiles the packed ones.
#pragma once
#include <cuda_runtime.h>

namespace dmel {
namespace px {

typedef unsigned long long u64;

__device__ __forceinline__ u64 pk(float2 v) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(v.x), "f"(v.y));
    return r;
}
__device__ __forceinline__ float2 up(u64 v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ float2 bc(float s) { return make_float2(s, s); }
__device__ __forceinline__ float2 swap(float2 a) { return make_float2(a.y, a.x); }
__device__ __forceinline__ float2 neg(float2 a) { return make_float2(-a.x, -a.y); }

#ifdef DMEL_F32X2
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk(a)), "l"(pk(b)));
    return up(r);
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
    u64 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk(a)), "l"(pk(b)));
    return up(r);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    u64 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk(a)), "l"(pk(b)));
    return up(r);
}
// a*b + c
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pk(a)), "l"(pk(b)), "l"(pk(c)));
    return up(r);
}
#else
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
}
#endif

// (1,-1): the one sign pattern a complex rotation needs.  swap(a)*PM = (a.y, -a.x) = -i*a
__device__ __forceinline__ float2 pm() { return make_float2(1.f, -1.f); }

// a * (-i)
__device__ __forceinline__ float2 mul_mi(float2 a) { return mul2(swap(a), pm()); }
// a + b*(-i),  a - b*(-i)
__device__ __forceinline__ float2 add_mi(float2 a, float2 b) { return fma2(swap(b), pm(), a); }
__device__ __forceinline__ float2 sub_mi(float2 a, float2 b) { return fma2(neg(swap(b)), pm(), a); }

// a * w, w a run-time twiddle: w enters as two broadcast scalars, so it needs no rotated copy in registers
__device__ __forceinline__ float2 cmul(float2 a, float2 w) {
    const float2 t = mul2(swap(a), neg(pm()));          // (-a.y, a.x) = i*a
    return fma2(bc(w.y), t, mul2(bc(w.x), a));
}
// e + o*(wr + i*wi), e - o*(wr + i*wi), w a compile-time constant: two FFMA2 per output
__device__ __forceinline__ void bf_const(float2 e, float2 o, float wr, float wi, float2& lo, float2& hi) {
    const float2 w = make_float2(wr, wi), wq = make_float2(-wi, wr);
    lo = fma2(bc(o.y), wq, fma2(bc(o.x), w, e));
    hi = fma2(bc(o.y), neg(wq), fma2(bc(o.x), neg(w), e));
}
// o * (wr + i*wi), compile-time constant
__device__ __forceinline__ float2 cmul_const(float2 o, float wr, float wi) {
    return fma2(bc(o.y), make_float2(-wi, wr), mul2(bc(o.x), make_float2(wr, wi)));
}

}  // namespace px
}  // namespace dmel
