// logmel_tc_tables.h -- host-side constant tables of the tensor-core FFT (logmel_tc_core.cuh).
// Plain C++ (no CUDA types): used by logmel_capi.cu and by tools/tc_fft_proto.cu.
//
// The 2048-point real FFT of a frame is a 1024-point complex FFT of z[m] = x[2m] + i x[2m+1] done as two
// radix-32 stages, m = 32 m1 + m2, k = q1 + 32 q2 (the decomposition round 1 ran on the CUDA cores), each stage
// a real [rows x 64] . [64 x 64] GEMM on tcgen05 with the same matrix G:
//     K index  kappa = 32 c + idx   (c = 0 real part, 1 imaginary part of the input point idx)
//     N index  nu    = 32 c' + q    (c' = 0 real, 1 imaginary part of output q; "planar", so that neighbouring
//                                    TMEM columns hold the same part of neighbouring outputs)
//     G[kappa][nu]: out_re[q] = sum zr cos(t) + zi sin(t),  out_im[q] = sum -zr sin(t) + zi cos(t),  t = 2 pi idx q / 32
// in fp16 head + fp16 residual (three MMA passes: head.head + residual.head + head.residual).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace lmtc_host {

inline uint16_t f32_to_f16_rn(float f) {   // IEEE round-to-nearest-even, subnormals kept
    uint32_t x;
    std::memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    x &= 0x7fffffffu;
    if (x >= 0x7f800000u) return static_cast<uint16_t>(sign | 0x7c00u | ((x > 0x7f800000u) ? 0x200u : 0u));
    if (x >= 0x477ff000u) return static_cast<uint16_t>(sign | 0x7c00u);             // rounds to inf
    if (x < 0x33000001u) return static_cast<uint16_t>(sign);                        // rounds to zero
    int e = static_cast<int>(x >> 23) - 127;
    uint32_t m = (x & 0x7fffffu) | 0x800000u;
    int shift;
    uint32_t he;
    if (e < -14) { shift = 13 + (-14 - e); he = 0; }                                // subnormal half
    else { shift = 13; he = static_cast<uint32_t>(e + 15); }
    const uint32_t keep = m >> shift, rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1);
    uint32_t h = (he << 10) + (he ? (keep & 0x3ffu) : keep);
    if (rem > half || (rem == half && (keep & 1u))) ++h;                            // carries into the exponent correctly
    return static_cast<uint16_t>(sign | h);
}
inline float f16_to_f32(uint16_t h) {
    const uint32_t sign = (h & 0x8000u) << 16, e = (h >> 10) & 0x1fu, m = h & 0x3ffu;
    uint32_t x;
    if (e == 0) {
        if (m == 0) x = sign;
        else {
            float v = std::ldexp(static_cast<float>(m), -24);
            std::memcpy(&x, &v, 4);
            x |= sign;
        }
    } else if (e == 31) x = sign | 0x7f800000u | (m << 13);
    else x = sign | ((e + 112u) << 23) | (m << 13);
    float f;
    std::memcpy(&f, &x, 4);
    return f;
}

constexpr int kGBytes = 64 * 64 * 2;   // one 64 x 64 fp16 matrix image

// UMMA K-major, no swizzle: element (nu, kappa) at (nu/8)*1024 + (nu%8)*16 + (kappa/8)*128 + (kappa%8)*2
// (core matrix = 8 rows x 16 bytes; LBO = 128 between k-groups, SBO = 1024 between 8-row groups)
inline size_t g_offset(int nu, int kappa) { return static_cast<size_t>(nu / 8) * 1024 + (nu % 8) * 16 + (kappa / 8) * 128 + (kappa % 8) * 2; }

// images of the head and residual matrices, kGBytes each
inline void build_dft32(std::vector<uint8_t>& g_hi, std::vector<uint8_t>& g_lo) {
    g_hi.assign(kGBytes, 0);
    g_lo.assign(kGBytes, 0);
    const double two_pi = 6.283185307179586476925286766559;
    for (int c = 0; c < 2; ++c)
        for (int idx = 0; idx < 32; ++idx)
            for (int cp = 0; cp < 2; ++cp)
                for (int q = 0; q < 32; ++q) {
                    const double t = two_pi * static_cast<double>((idx * q) % 32) / 32.0;
                    double g;
                    if (c == 0 && cp == 0) g = std::cos(t);
                    else if (c == 1 && cp == 0) g = std::sin(t);
                    else if (c == 0 && cp == 1) g = -std::sin(t);
                    else g = std::cos(t);
                    if (std::fabs(g) < 1e-15) g = 0.0;
                    const float gf = static_cast<float>(g);
                    const uint16_t h = f32_to_f16_rn(gf);
                    const uint16_t l = f32_to_f16_rn(static_cast<float>(g - static_cast<double>(f16_to_f32(h))));
                    const size_t o = g_offset(32 * cp + q, 32 * c + idx);
                    std::memcpy(&g_hi[o], &h, 2);
                    std::memcpy(&g_lo[o], &l, 2);
                }
}

constexpr int kTw1Rows = 17;      // j = 0 .. 16
constexpr int kUtwRows = 17;      // j = 0 .. 16 (row 16 is used by the fix-up of the bins k = 16 mod 32)
constexpr int kUtwPitch = 34;     // floats per row: lanes j read 8-byte pairs, 34 keeps them on distinct banks

// tw1[j][m2] = (cos, -sin)(2 pi j m2 / 1024): the inter-stage twiddle W_1024^(j m2)
inline void build_tw1(std::vector<float>& tw1) {
    tw1.assign(static_cast<size_t>(kTw1Rows) * 32 * 2, 0.f);
    const double two_pi = 6.283185307179586476925286766559;
    for (int j = 0; j < kTw1Rows; ++j)
        for (int m2 = 0; m2 < 32; ++m2) {
            const double a = two_pi * static_cast<double>(j * m2) / 1024.0;
            tw1[(static_cast<size_t>(j) * 32 + m2) * 2] = static_cast<float>(std::cos(a));
            tw1[(static_cast<size_t>(j) * 32 + m2) * 2 + 1] = static_cast<float>(-std::sin(a));
        }
}
// utw_c[j][p], utw_s[j][p] = cos / sin (2 pi k / 2048), k = j + 32 p: the real-FFT untangle twiddle, planar
inline void build_utw(std::vector<float>& utw_c, std::vector<float>& utw_s) {
    utw_c.assign(static_cast<size_t>(kUtwRows) * kUtwPitch, 0.f);
    utw_s.assign(static_cast<size_t>(kUtwRows) * kUtwPitch, 0.f);
    const double two_pi = 6.283185307179586476925286766559;
    for (int j = 0; j < kUtwRows; ++j)
        for (int p = 0; p < 32; ++p) {
            const double a = two_pi * static_cast<double>(j + 32 * p) / 2048.0;
            utw_c[static_cast<size_t>(j) * kUtwPitch + p] = static_cast<float>(std::cos(a));
            utw_s[static_cast<size_t>(j) * kUtwPitch + p] = static_cast<float>(std::sin(a));
        }
}

}  // namespace lmtc_host
