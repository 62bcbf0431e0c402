// logmel_aux.cuh -- small companion kernels of the log-mel path.
//
//   resize_finish_kernel   FlexibleAudioPreprocessor.resize_spectrogram + augment_spectrogram + normalize
//                          (R/data/preprocessing_flexible.py:118-154, :106-116, order of :182-190):
//                          bilinear F.interpolate(size=(n_mels, target), align_corners=False).  The mel
//                          axis keeps its size, so the kernel is a 1-D linear interpolation along time
//                          (aten/src/ATen/native/UpSample.h: src = scale*(dst+0.5)-0.5 clamped at 0,
//                          scale = in/out), then the SpecAugment intervals, then (x-mean)/(std+eps).
//   pcm16_roundtrip_kernel the analyzers write every window to a temp .wav with soundfile's default
//                          PCM_16 subtype and read it back (R/realtime_analyzer_parallel.py:181-184):
//                          y = rint(clamp(x) * 32767) / 32768.  "Parity unpinned" (libsndfile is not in
//                          the reference tree); offered so window features can follow that accident.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/logmel_b200.h"

namespace lm {

constexpr int kAuxThreads = 512;

__device__ __forceinline__ float resize_sample(const float* __restrict__ row, int fin, int t, float scale) {
    float src = fmaf(scale, static_cast<float>(t) + 0.5f, -0.5f);
    if (src < 0.0f) src = 0.0f;
    int i0 = static_cast<int>(src);
    if (i0 > fin - 1) i0 = fin - 1;
    float lam = src - static_cast<float>(i0);
    lam = fminf(fmaxf(lam, 0.0f), 1.0f);
    const int i1 = (i0 + 1 < fin) ? i0 + 1 : i0;
    return (1.0f - lam) * row[i0] + lam * row[i1];
}

// one CTA per clip; in [B, n_mels, fin] -> out [B, n_mels, fout]
__global__ void __launch_bounds__(kAuxThreads) resize_finish_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                                     const lm_aug* __restrict__ aug, int n_mels, int fin,
                                                                     int fout, int normalize, float eps) {
    __shared__ double red[2 * (kAuxThreads / 32)];
    __shared__ float bc[2];
    const int clip = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* __restrict__ src = in + static_cast<size_t>(clip) * n_mels * fin;
    float* __restrict__ dst = out + static_cast<size_t>(clip) * n_mels * fout;
    int f0 = 0, f1 = 0, t0 = 0, t1 = 0;
    if (aug != nullptr) { f0 = aug[clip].f0; f1 = aug[clip].f1; t0 = aug[clip].t0; t1 = aug[clip].t1; }
    const float scale = static_cast<float>(fin) / static_cast<float>(fout);
    const int n = n_mels * fout;
    double s = 0.0, q = 0.0;
    for (int i = tid; i < n; i += kAuxThreads) {
        const int m = i / fout, t = i - m * fout;
        float v = (fin == fout) ? src[m * fin + t] : resize_sample(src + m * fin, fin, t, scale);
        if ((m >= f0 && m < f1) || (t >= t0 && t < t1)) v = 0.0f;
        dst[i] = v;
        const double d = static_cast<double>(v);
        s += d;
        q = fma(d, d, q);
    }
    if (!normalize) return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if (lane == 0) { red[warp] = s; red[kAuxThreads / 32 + warp] = q; }
    __syncthreads();
    if (tid == 0) {
        double ss = 0.0, qq = 0.0;
        for (int w = 0; w < kAuxThreads / 32; ++w) { ss += red[w]; qq += red[kAuxThreads / 32 + w]; }
        const double mean = ss / n;
        double var = (qq - ss * mean) / (static_cast<double>(n) - 1.0);
        if (!(var > 0.0)) var = 0.0;
        bc[0] = static_cast<float>(mean);
        bc[1] = static_cast<float>(sqrt(var)) + eps;
    }
    __syncthreads();
    const float mean = bc[0], inv = 1.0f / bc[1];
    for (int i = tid; i < n; i += kAuxThreads) dst[i] = (dst[i] - mean) * inv;   // own writes: same thread
}

__global__ void pcm16_roundtrip_kernel(const float* __restrict__ in, float* __restrict__ out, long long n) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        float x = in[i];
        x = fminf(fmaxf(x, -1.0f), 1.0f);
        out[i] = rintf(x * 32767.0f) * (1.0f / 32768.0f);
    }
}

// T.AmplitudeToDB() on its own (R/src/data/preprocessing.py:46, TA/functional/functional.py:390-391):
// y = multiplier * log10(max(x, amin)) - offset, element-wise; the drop-in preprocessor's `amplitude_to_db` attribute.
__global__ void amplitude_to_db_kernel(const float* __restrict__ in, float* __restrict__ out, long long n, float db_scale, float amin,
                                       float db_offset) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = fmaf(db_scale, log2f(fmaxf(in[i], amin)), -db_offset);
}

// 16-bit PCM -> fp32: x / 32768 (exact), what torchaudio.load(normalize=True) hands the reference for a
// PCM_16 wav (R/src/data/preprocessing.py:57).  Eight samples per thread per pass: one 128-bit load, two
// 128-bit stores; `in` and `out` must be 16-byte aligned, the tail is done element-wise.
__global__ void pcm16_decode_kernel(const int16_t* __restrict__ in, float* __restrict__ out, long long n) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long n8 = n >> 3;
    const int4* __restrict__ in8 = reinterpret_cast<const int4*>(in);
    float4* __restrict__ out4 = reinterpret_cast<float4*>(out);
    constexpr float k = 1.0f / 32768.0f;
    for (long long i = tid; i < n8; i += stride) {
        const int4 v = __ldcs(in8 + i);
        float4 a, b;
        a.x = static_cast<float>(static_cast<short>(v.x & 0xffff)) * k;
        a.y = static_cast<float>(v.x >> 16) * k;
        a.z = static_cast<float>(static_cast<short>(v.y & 0xffff)) * k;
        a.w = static_cast<float>(v.y >> 16) * k;
        b.x = static_cast<float>(static_cast<short>(v.z & 0xffff)) * k;
        b.y = static_cast<float>(v.z >> 16) * k;
        b.z = static_cast<float>(static_cast<short>(v.w & 0xffff)) * k;
        b.w = static_cast<float>(v.w >> 16) * k;
        out4[2 * i] = a;
        out4[2 * i + 1] = b;
    }
    for (long long i = (n8 << 3) + tid; i < n; i += stride) out[i] = static_cast<float>(in[i]) * k;
}

// Polyphase sinc resampler = torchaudio.transforms.Resample(orig, new) with its defaults, the resampler of
// AudioPreprocessor.load_audio (R/src/data/preprocessing.py:63-65; TA/functional/functional.py
// _get_sinc_resample_kernel / _apply_sinc_resample_kernel).  With o = orig/gcd, q = new/gcd and
// xpad = pad(x, (width, width + o)):   y[m q + p] = sum_k kernel[p][k] xpad[m o + k],  k < 2 width + o.
// torchaudio runs that as a dense conv1d; all but ~2 width + 1 taps of a phase are (numerically) zero, so
// the host keeps, per phase, the window of `ntaps` taps around the phase's centre and the kernel walks only
// those.  One output sample per thread; a block's outputs share their input span through L1.
// grid.y = row (channel / clip): rows of equal length `n_in`, `in_stride` / `out_stride` floats apart
__global__ void resample_kernel(const float* __restrict__ x_all, long long n_in, long long in_stride,
                                float* __restrict__ y_all, long long n_out, long long out_stride,
                                const float* __restrict__ taps, const int* __restrict__ k0, int o, int q, int ntaps,
                                int width) {
    const long long j = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (j >= n_out) return;
    const float* __restrict__ x = x_all + static_cast<long long>(blockIdx.y) * in_stride;
    float* __restrict__ y = y_all + static_cast<long long>(blockIdx.y) * out_stride;
    const long long m = j / q;
    const int p = static_cast<int>(j - m * q);
    const float* __restrict__ w = taps + static_cast<size_t>(p) * ntaps;
    const long long base = m * o + k0[p] - width;   // index into x of tap 0
    float acc = 0.0f;
    for (int i = 0; i < ntaps; ++i) {
        const long long xi = base + i;
        const float v = (xi >= 0 && xi < n_in) ? __ldg(x + xi) : 0.0f;
        acc = fmaf(w[i], v, acc);
    }
    y[j] = acc;
}


// The same resampler, tiled: a block produces kRsChunk consecutive outputs of one row from ONE copy of their input
// span in shared memory (loaded coalesced, zero outside the row: each input sample feeds ~ntaps * q / o outputs), taps
// padded to a multiple of four per phase and fetched as 128-bit read-only loads.  The taps are walked in the same
// order with the same fmaf chain as resample_kernel (the padding adds fmaf(0, x, acc) = acc), so the result is
// bit-identical to it.
constexpr int kRsChunk = 1024, kRsThreads = 256;
__global__ void __launch_bounds__(kRsThreads) resample_tiled_kernel(const float* __restrict__ x_all, long long n_in, long long in_stride,
                                                                     float* __restrict__ y_all, long long n_out, long long out_stride,
                                                                     const float4* __restrict__ taps4, const int* __restrict__ k0, int o,
                                                                     int q, int nt4, int width, int k0max) {
    extern __shared__ float sx[];
    const float* __restrict__ x = x_all + static_cast<long long>(blockIdx.y) * in_stride;
    float* __restrict__ y = y_all + static_cast<long long>(blockIdx.y) * out_stride;
    const long long j0 = static_cast<long long>(blockIdx.x) * kRsChunk;
    const int n_here = static_cast<int>((n_out - j0 < kRsChunk) ? (n_out - j0) : kRsChunk);
    const long long m0 = j0 / q;
    const unsigned p0 = static_cast<unsigned>(j0 - m0 * q);
    const long long x_lo = m0 * o - width;                                             // tap 0 of the first output at the earliest
    const int m_span = static_cast<int>((p0 + static_cast<unsigned>(n_here) - 1u) / static_cast<unsigned>(q));
    const int span = m_span * o + k0max + 4 * nt4;                                     // covers the last output's last (padded) tap
    for (int i = threadIdx.x; i < span; i += kRsThreads) {
        const long long xi = x_lo + i;
        sx[i] = (xi >= 0 && xi < n_in) ? __ldg(x + xi) : 0.0f;
    }
    __syncthreads();
    for (int d = threadIdx.x; d < n_here; d += kRsThreads) {
        const unsigned t = p0 + static_cast<unsigned>(d);
        const unsigned dm = t / static_cast<unsigned>(q), p = t - dm * static_cast<unsigned>(q);
        const float4* __restrict__ w = taps4 + static_cast<size_t>(p) * nt4;
        const float* __restrict__ xs = sx + dm * o + k0[p];
        float acc = 0.0f;
        for (int i = 0; i < nt4; ++i) {
            const float4 wv = __ldg(w + i);
            acc = fmaf(wv.x, xs[4 * i], acc);
            acc = fmaf(wv.y, xs[4 * i + 1], acc);
            acc = fmaf(wv.z, xs[4 * i + 2], acc);
            acc = fmaf(wv.w, xs[4 * i + 3], acc);
        }
        y[j0 + d] = acc;
    }
}

}  // namespace lm
