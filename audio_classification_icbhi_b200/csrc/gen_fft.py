#!/usr/bin/env python3
"""Emits straight-line, register-resident radix-2 DIT FFT kernels (fft_gen.cuh).

Every index and twiddle is a compile-time literal, so nvcc keeps whole arrays in registers.
Three flavours of the same forward transform (sign e^{-2 pi i nk/N}), natural order in and out:

  lm_fft{16,32}(float xr[N], float xi[N])
      scalar reference flavour (used by the CPU tests and as documentation of the op count).

  lm_fft32_aos(lm_f2 z[32])
      one complex point per 64-bit register (lo = re, hi = im), packed fp32x2 arithmetic:
      every butterfly is 2 FADD2 or 3 FFMA2 -- half the issue slots of the scalar flavour.
      Used for the first (lane-local) FFT, whose input arrives from shared memory as float2.

  lm_fft32_aos_from2(lm_f2 z[32])
      the same network minus its first stage: the kernel computes z[r] +- z[r+16] itself, fused
      with the Hann window (hann[n + n_fft/2] = 1 - hann[n], so one window load serves both inputs).

  lm_fft32_soa(const lm_f2 pr[16], const lm_f2 pi[16], float xr[32], float xi[32])
      input pr[m] = (re x[2m], re x[2m+1]), pi[m] likewise -- what the LDS.128 reads of the
      32x32 transpose deliver.  In DIT order those two points share every twiddle of stages 1-4
      (they are the even- and odd-sample 16-point sub-transforms), so stages 1-4 run packed with
      broadcast constants; stage 5 combines the halves of each register in scalar code.

Non-trivial butterflies use the 6-FMA (3 packed) Linzer-Feig form with the tangent
(|c| >= |s|) or cotangent (|c| < |s|) ratio so that |ratio| <= 1:

    w = c - i s,  t = s/c:   b~ = (br + t bi, bi - t br)        a +- c b~
    w = c - i s,  t = c/s:   b^ = (br - t bi, bi + t br)        a +- s (b^i, -b^r)

The text only uses the helpers of lm_f2.cuh, which are PTX f32x2 ops under nvcc and plain C
under gcc: tests/test_fft_codegen.py compiles this very file with gcc and checks it against numpy.

Usage:  python gen_fft.py > fft_gen.cuh
"""
import math


def bitrev(i: int, bits: int) -> int:
    r = 0
    for _ in range(bits):
        r = (r << 1) | (i & 1)
        i >>= 1
    return r


def lit(x: float) -> str:
    return f"{x:.9e}f"


def twiddle_kind(j: int, m: int):
    """('one'|'mi'|'tan'|'cot', c, s, t) for w = exp(-2 pi i j / m)."""
    c = math.cos(2 * math.pi * j / m)
    s = math.sin(2 * math.pi * j / m)
    if j == 0:
        return "one", c, s, 0.0
    if 4 * j == m:
        return "mi", c, s, 0.0
    if abs(c) >= abs(s):
        return "tan", c, s, s / c
    return "cot", c, s, c / s


# --------------------------------------------------------------------------------------
def gen_scalar(n: int) -> str:
    bits = n.bit_length() - 1
    out, ops, tmp = [], 0, 0
    emit = out.append
    emit(f"LM_HD LM_INLINE void lm_fft{n}(float (&xr)[{n}], float (&xi)[{n}]) {{")
    cur = [(f"xr[{bitrev(i, bits)}]", f"xi[{bitrev(i, bits)}]") for i in range(n)]
    for s in range(1, bits + 1):
        m, half = 1 << s, 1 << (s - 1)
        for k in range(0, n, m):
            for j in range(half):
                ia, ib = k + j, k + j + half
                (ar, ai), (br, bi) = cur[ia], cur[ib]
                p = f"t{tmp}"
                tmp += 1
                kind, c, sn, t = twiddle_kind(j, m)
                if kind == "one":
                    emit(f"  const float {p}ar = {ar} + {br}, {p}ai = {ai} + {bi};")
                    emit(f"  const float {p}br = {ar} - {br}, {p}bi = {ai} - {bi};")
                    ops += 4
                elif kind == "mi":
                    emit(f"  const float {p}ar = {ar} + {bi}, {p}ai = {ai} - {br};")
                    emit(f"  const float {p}br = {ar} - {bi}, {p}bi = {ai} + {br};")
                    ops += 4
                elif kind == "tan":
                    emit(f"  const float {p}ur = fmaf({lit(t)}, {bi}, {br}), {p}ui = fmaf({lit(-t)}, {br}, {bi});")
                    emit(f"  const float {p}ar = fmaf({lit(c)}, {p}ur, {ar}), {p}ai = fmaf({lit(c)}, {p}ui, {ai});")
                    emit(f"  const float {p}br = fmaf({lit(-c)}, {p}ur, {ar}), {p}bi = fmaf({lit(-c)}, {p}ui, {ai});")
                    ops += 6
                else:
                    emit(f"  const float {p}ur = fmaf({lit(-t)}, {bi}, {br}), {p}ui = fmaf({lit(t)}, {br}, {bi});")
                    emit(f"  const float {p}ar = fmaf({lit(sn)}, {p}ui, {ar}), {p}ai = fmaf({lit(-sn)}, {p}ur, {ai});")
                    emit(f"  const float {p}br = fmaf({lit(-sn)}, {p}ui, {ar}), {p}bi = fmaf({lit(sn)}, {p}ur, {ai});")
                    ops += 6
                cur[ia] = (f"{p}ar", f"{p}ai")
                cur[ib] = (f"{p}br", f"{p}bi")
    for i in range(n):
        emit(f"  xr[{i}] = {cur[i][0]}; xi[{i}] = {cur[i][1]};")
    emit("}")
    emit(f"// lm_fft{n}: {ops} scalar float ops")
    return "\n".join(out)


# --------------------------------------------------------------------------------------
def gen_aos(n: int, first_stage: int = 1) -> str:
    """Packed complex (re, im) per register.  first_stage = 2: z[] already holds the results of the
    first butterfly stage (z[r] +- z[r + n/2] in place) -- the caller fused it with the window."""
    bits = n.bit_length() - 1
    out, ops, tmp = [], 0, 0
    emit = out.append
    name = f"lm_fft{n}_aos" if first_stage == 1 else f"lm_fft{n}_aos_from{first_stage}"
    emit(f"LM_HD LM_INLINE void {name}(lm_f2 (&z)[{n}]) {{")
    cur = [f"z[{bitrev(i, bits)}]" for i in range(n)]
    for s in range(first_stage, bits + 1):
        m, half = 1 << s, 1 << (s - 1)
        for k in range(0, n, m):
            for j in range(half):
                ia, ib = k + j, k + j + half
                a, b = cur[ia], cur[ib]
                p = f"c{tmp}"
                tmp += 1
                kind, c, sn, t = twiddle_kind(j, m)
                if kind == "one":
                    emit(f"  const lm_f2 {p}a = lm_add2({a}, {b}), {p}b = lm_sub2({a}, {b});")
                    ops += 2
                elif kind == "mi":   # w b = -i b = (bi, -br)
                    emit(f"  const lm_f2 {p}a = lm_add2({a}, lm_mul_mi({b})), {p}b = lm_sub2({a}, lm_mul_mi({b}));")
                    ops += 2
                elif kind == "tan":  # b~ = b + t (bi, -br);  a +- c b~
                    emit(f"  const lm_f2 {p}u = lm_fma2(lm_swap({b}), lm_pack({lit(t)}, {lit(-t)}), {b});")
                    emit(f"  const lm_f2 {p}a = lm_fma2({p}u, lm_bcast({lit(c)}), {a}), "
                         f"{p}b = lm_fma2({p}u, lm_bcast({lit(-c)}), {a});")
                    ops += 3
                else:                # b^ = b + t (-bi, br);  a +- s (b^i, -b^r)
                    emit(f"  const lm_f2 {p}u = lm_fma2(lm_swap({b}), lm_pack({lit(-t)}, {lit(t)}), {b});")
                    emit(f"  const lm_f2 {p}a = lm_fma2(lm_swap({p}u), lm_pack({lit(sn)}, {lit(-sn)}), {a}), "
                         f"{p}b = lm_fma2(lm_swap({p}u), lm_pack({lit(-sn)}, {lit(sn)}), {a});")
                    ops += 3
                cur[ia], cur[ib] = f"{p}a", f"{p}b"
    for i in range(n):
        emit(f"  z[{i}] = {cur[i]};")
    emit("}")
    emit(f"// {name}: {ops} packed f32x2 ops")
    return "\n".join(out)


# --------------------------------------------------------------------------------------
def gen_soa32() -> str:
    """Stages 1-4 on (even-sample, odd-sample) register pairs, stage 5 scalar."""
    out, ops2, ops1, tmp = [], 0, 0, 0
    emit = out.append
    emit("LM_HD LM_INLINE void lm_fft32_soa(const lm_f2 (&pr)[16], const lm_f2 (&pi)[16], "
         "float (&xr)[32], float (&xi)[32]) {")
    # DIT-16 array A[i] = P[bitrev4(i)], each entry a (re pair, im pair)
    cur = [(f"pr[{bitrev(i, 4)}]", f"pi[{bitrev(i, 4)}]") for i in range(16)]
    for s in range(1, 5):
        m, half = 1 << s, 1 << (s - 1)
        for k in range(0, 16, m):
            for j in range(half):
                ia, ib = k + j, k + j + half
                (ar, ai), (br, bi) = cur[ia], cur[ib]
                p = f"s{tmp}"
                tmp += 1
                kind, c, sn, t = twiddle_kind(j, m)
                if kind == "one":
                    emit(f"  const lm_f2 {p}ar = lm_add2({ar}, {br}), {p}ai = lm_add2({ai}, {bi});")
                    emit(f"  const lm_f2 {p}br = lm_sub2({ar}, {br}), {p}bi = lm_sub2({ai}, {bi});")
                    ops2 += 4
                elif kind == "mi":
                    emit(f"  const lm_f2 {p}ar = lm_add2({ar}, {bi}), {p}ai = lm_sub2({ai}, {br});")
                    emit(f"  const lm_f2 {p}br = lm_sub2({ar}, {bi}), {p}bi = lm_add2({ai}, {br});")
                    ops2 += 4
                elif kind == "tan":
                    emit(f"  const lm_f2 {p}ur = lm_fma2(lm_bcast({lit(t)}), {bi}, {br}), "
                         f"{p}ui = lm_fma2(lm_bcast({lit(-t)}), {br}, {bi});")
                    emit(f"  const lm_f2 {p}ar = lm_fma2(lm_bcast({lit(c)}), {p}ur, {ar}), "
                         f"{p}ai = lm_fma2(lm_bcast({lit(c)}), {p}ui, {ai});")
                    emit(f"  const lm_f2 {p}br = lm_fma2(lm_bcast({lit(-c)}), {p}ur, {ar}), "
                         f"{p}bi = lm_fma2(lm_bcast({lit(-c)}), {p}ui, {ai});")
                    ops2 += 6
                else:
                    emit(f"  const lm_f2 {p}ur = lm_fma2(lm_bcast({lit(-t)}), {bi}, {br}), "
                         f"{p}ui = lm_fma2(lm_bcast({lit(t)}), {br}, {bi});")
                    emit(f"  const lm_f2 {p}ar = lm_fma2(lm_bcast({lit(sn)}), {p}ui, {ar}), "
                         f"{p}ai = lm_fma2(lm_bcast({lit(-sn)}), {p}ur, {ai});")
                    emit(f"  const lm_f2 {p}br = lm_fma2(lm_bcast({lit(-sn)}), {p}ui, {ar}), "
                         f"{p}bi = lm_fma2(lm_bcast({lit(sn)}), {p}ur, {ai});")
                    ops2 += 6
                cur[ia] = (f"{p}ar", f"{p}ai")
                cur[ib] = (f"{p}br", f"{p}bi")
    # stage 5: X[i] = E[i] + w32^i O[i], X[i+16] = E[i] - w32^i O[i]; E = lo halves, O = hi halves
    for i in range(16):
        R, I = cur[i]
        ar, ai, br, bi = f"lm_lo({R})", f"lm_lo({I})", f"lm_hi({R})", f"lm_hi({I})"
        kind, c, sn, t = twiddle_kind(i, 32)
        p = f"f{i}"
        if kind == "one":
            emit(f"  xr[{i}] = {ar} + {br}; xi[{i}] = {ai} + {bi}; xr[{i + 16}] = {ar} - {br}; xi[{i + 16}] = {ai} - {bi};")
            ops1 += 4
        elif kind == "mi":
            emit(f"  xr[{i}] = {ar} + {bi}; xi[{i}] = {ai} - {br}; xr[{i + 16}] = {ar} - {bi}; xi[{i + 16}] = {ai} + {br};")
            ops1 += 4
        elif kind == "tan":
            emit(f"  const float {p}ur = fmaf({lit(t)}, {bi}, {br}), {p}ui = fmaf({lit(-t)}, {br}, {bi});")
            emit(f"  xr[{i}] = fmaf({lit(c)}, {p}ur, {ar}); xi[{i}] = fmaf({lit(c)}, {p}ui, {ai}); "
                 f"xr[{i + 16}] = fmaf({lit(-c)}, {p}ur, {ar}); xi[{i + 16}] = fmaf({lit(-c)}, {p}ui, {ai});")
            ops1 += 6
        else:
            emit(f"  const float {p}ur = fmaf({lit(-t)}, {bi}, {br}), {p}ui = fmaf({lit(t)}, {br}, {bi});")
            emit(f"  xr[{i}] = fmaf({lit(sn)}, {p}ui, {ar}); xi[{i}] = fmaf({lit(-sn)}, {p}ur, {ai}); "
                 f"xr[{i + 16}] = fmaf({lit(-sn)}, {p}ui, {ar}); xi[{i + 16}] = fmaf({lit(sn)}, {p}ur, {ai});")
            ops1 += 6
    emit("}")
    emit(f"// lm_fft32_soa: {ops2} packed f32x2 ops + {ops1} scalar ops")
    return "\n".join(out)


def main() -> None:
    print("// GENERATED by gen_fft.py -- do not edit.  Register-resident forward FFTs.")
    print("#pragma once")
    print('#include "lm_f2.cuh"')
    for n in (16, 32):
        print()
        print(gen_scalar(n))
    print()
    print(gen_aos(32))
    print()
    print(gen_aos(32, first_stage=2))
    print()
    print(gen_soa32())


if __name__ == "__main__":
    main()
