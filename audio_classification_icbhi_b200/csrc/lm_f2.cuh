// lm_f2.cuh -- a pair of fp32 values in one 64-bit register, and the sm_100a packed ops on it.
//
// Blackwell has 2-wide fp32 instructions (PTX add/sub/mul/fma.rn.f32x2 -> SASS FADD2/FMUL2/FFMA2).
// ptxas folds half swaps, per-half negations and 32-bit broadcasts of the operands into the
// instruction's operand modifiers (R.F32x2.LO_HI, .NP/.PN, R.F32), so `lm_swap`, `lm_conj`,
// `lm_bcast` and literal (t, -t) pairs below normally cost nothing: a complex radix-2 butterfly
// with a constant twiddle is three FFMA2 instead of six FFMA.
//
// Under a host compiler the same names are plain scalar code (fmaf per half), so the generated
// FFTs are checked on the CPU with gcc (tests/test_fft_codegen.py).
#pragma once
#include <math.h>

#ifndef LM_HD
#  ifdef __CUDACC__
#    define LM_HD __host__ __device__
#    define LM_INLINE __forceinline__
#  else
#    define LM_HD
#    define LM_INLINE inline
#  endif
#endif

#if defined(__CUDA_ARCH__)

typedef unsigned long long lm_f2;

__device__ __forceinline__ lm_f2 lm_pack(float lo, float hi) {
    lm_f2 r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float lm_lo(lm_f2 v) {
    float a, b;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    (void)b;
    return a;
}
__device__ __forceinline__ float lm_hi(lm_f2 v) {
    float a, b;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    (void)a;
    return b;
}
__device__ __forceinline__ lm_f2 lm_add2(lm_f2 a, lm_f2 b) {
    lm_f2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ lm_f2 lm_sub2(lm_f2 a, lm_f2 b) {
    lm_f2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ lm_f2 lm_mul2(lm_f2 a, lm_f2 b) {
    lm_f2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ lm_f2 lm_fma2(lm_f2 a, lm_f2 b, lm_f2 c) {
    lm_f2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

#else  // host

struct lm_f2 {
    float lo, hi;
};
LM_INLINE lm_f2 lm_pack(float lo, float hi) { lm_f2 r; r.lo = lo; r.hi = hi; return r; }
LM_INLINE float lm_lo(lm_f2 v) { return v.lo; }
LM_INLINE float lm_hi(lm_f2 v) { return v.hi; }
LM_INLINE lm_f2 lm_add2(lm_f2 a, lm_f2 b) { return lm_pack(a.lo + b.lo, a.hi + b.hi); }
LM_INLINE lm_f2 lm_sub2(lm_f2 a, lm_f2 b) { return lm_pack(a.lo - b.lo, a.hi - b.hi); }
LM_INLINE lm_f2 lm_mul2(lm_f2 a, lm_f2 b) { return lm_pack(a.lo * b.lo, a.hi * b.hi); }
LM_INLINE lm_f2 lm_fma2(lm_f2 a, lm_f2 b, lm_f2 c) { return lm_pack(fmaf(a.lo, b.lo, c.lo), fmaf(a.hi, b.hi, c.hi)); }

#endif

// operand shapes that ptxas turns into modifiers
LM_HD LM_INLINE lm_f2 lm_swap(lm_f2 v) { return lm_pack(lm_hi(v), lm_lo(v)); }          // (hi, lo)
LM_HD LM_INLINE lm_f2 lm_conj(lm_f2 v) { return lm_pack(lm_lo(v), -lm_hi(v)); }         // (lo, -hi)
LM_HD LM_INLINE lm_f2 lm_mul_mi(lm_f2 v) { return lm_pack(lm_hi(v), -lm_lo(v)); }       // -i * (lo + i hi)
LM_HD LM_INLINE lm_f2 lm_bcast(float s) { return lm_pack(s, s); }
