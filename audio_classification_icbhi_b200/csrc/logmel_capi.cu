// logmel_capi.cu -- extern "C" boundary of liblogmel_b200.so (see include/logmel_b200.h).
//
// Host side only: plan construction (twiddles, banded filterbank rows), launch, and the
// host-buffer pipeline.  No torch types, no ATen: plain pointers and sizes.
#include <math.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <new>
#include <vector>

#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: ranges are no-ops unless a profiler is attached

#include "logmel_kernel.cuh"
#include "logmel_aux.cuh"

namespace {

thread_local char g_cuda_err[256] = "";

int cuda_fail(cudaError_t e, const char* what) {
    snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", what, cudaGetErrorString(e));
    return LM_ERR_CUDA;
}
#define LM_CUDA(call)                                        \
    do {                                                     \
        cudaError_t _e = (call);                             \
        if (_e != cudaSuccess) return cuda_fail(_e, #call);  \
    } while (0)

// The companion entry points take raw device pointers and a stream but no plan: make the device that owns the
// pointer current for the launch (a caller on cuda:0 may hold tensors of cuda:1) and restore it afterwards.
struct DeviceOf {
    int prev = -1, dev = -1;
    explicit DeviceOf(const void* ptr) {
        cudaPointerAttributes a{};
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (ptr && cudaPointerGetAttributes(&a, ptr) == cudaSuccess && a.type == cudaMemoryTypeDevice) dev = a.device;
        cudaGetLastError();
        if (dev >= 0 && dev != prev) cudaSetDevice(dev); else dev = -1;
    }
    ~DeviceOf() { if (dev >= 0 && prev >= 0) cudaSetDevice(prev); }
};

// NVTX range around a C-ABI entry point (ranges around plan / execute), visible in Nsight Systems / Compute
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

constexpr int kSlots = 3;   // host pipeline depth
constexpr int kCounters = 1024;      // launches of one plan that may be in flight at once (on any streams); a slot is
                                     // re-zeroed on the launching stream right before its kernel, so launches on ONE
                                     // stream never interfere, and 1024 concurrent streams per plan are out of reach
constexpr int kMaxSplitClips = 512;  // small-batch mode is used below 2 * SM count clips

struct HostSlot {
    cudaStream_t stream = nullptr;
    float* d_wave = nullptr;      size_t cap_wave = 0;     // floats
    int16_t* d_pcm = nullptr;     size_t cap_pcm = 0;      // 16-bit samples (lm_forward_host_pcm16)
    float* d_noise = nullptr;     size_t cap_noise = 0;
    float* d_out = nullptr;       size_t cap_out = 0;
    // per-chunk metadata, one pinned staging block and one device block: [offset int64 x n][aug 40 B x n][length int32 x n]
    // (a copy from the caller's pageable arrays would make cudaMemcpyAsync wait for the stream, i.e. for the
    //  chunk's waveform copy, and stall the pipeline once per chunk)
    unsigned char* d_meta = nullptr;
    unsigned char* h_meta = nullptr;   // pinned
    int cap_clips = 0;
};

}  // namespace

struct lm_plan {
    int device = 0;
    int n_fft = 0, hop = 0, n_mels = 0, T = 0, frames = 0, n_freqs = 0;
    int tile_f = 0, n_tiles = 0, ns = 0, n_dk = 0, fb_nnz = 0;
    int sm_count = 0, max_ctas = 0, use_tma = 1, stagger_ns = 0;
    int host_chunk_clips = 0;   // lm_forward_host chunk size; 0 = automatic
    size_t smem_bytes = 0;
    float db_mult = 10.f, amin = 1e-10f, db_offset = 0.f, floor_db = -100.f, norm_eps = 1e-8f;
    // device constants
    float* d_window = nullptr;
    float2* d_tw = nullptr;
    float2* d_utw = nullptr;
    float4* d_melw = nullptr;
    lm::MelTable* d_tab = nullptr;
    // host pipeline
    HostSlot slots[kSlots];
    bool slots_ready = false;
    std::atomic<long long> launches{0};
    int* d_counters = nullptr;   // kCounters x {work counter, finished groups} for the kernel's dynamic clip scheduling, one pair per launch in flight
    // small-batch mode scratch, one block per launch in flight: [kMaxSplitClips][2] 64-bit sums, then [kMaxSplitClips] arrival counters
    unsigned char* d_split = nullptr;
    int split_override = 0;      // 0 = automatic, 1 = never split, k > 1 = at most k chunks per clip
    std::mutex host_mu;          // lm_forward_host*: one caller at a time per plan (plan-owned staging buffers)
};

namespace {

int free_plan(lm_plan* p) {
    if (!p) return LM_OK;
    cudaSetDevice(p->device);
    cudaFree(p->d_window); cudaFree(p->d_tw); cudaFree(p->d_utw); cudaFree(p->d_melw); cudaFree(p->d_tab); cudaFree(p->d_counters); cudaFree(p->d_split);
    for (auto& s : p->slots) {
        if (s.stream) cudaStreamDestroy(s.stream);
        cudaFree(s.d_wave); cudaFree(s.d_pcm); cudaFree(s.d_noise); cudaFree(s.d_out);
        cudaFree(s.d_meta);
        if (s.h_meta) cudaFreeHost(s.h_meta);
    }
    delete p;
    return LM_OK;
}

lm::KParams make_params(const lm_plan* p) {
    lm::KParams k{};
    k.T = p->T; k.hop = p->hop; k.frames = p->frames; k.n_mels = p->n_mels; k.n_tiles = p->n_tiles;
    k.ns = p->ns; k.n_dk = p->n_dk; k.use_tma = p->use_tma; k.stagger_ns = p->stagger_ns;
    k.db_scale = static_cast<float>(static_cast<double>(p->db_mult) * 0.30102999566398119521);
    k.amin = p->amin; k.db_offset = p->db_offset; k.floor_db = p->floor_db;
    k.norm_eps = p->norm_eps;
    k.window = p->d_window; k.tw = p->d_tw; k.utw = p->d_utw; k.melw = p->d_melw; k.mel_table = p->d_tab;
    return k;
}

int launch(lm_plan* p, const float* wave, const int64_t* offset, const int32_t* length, int32_t B,
           const lm_aug* aug, const float* noise, float* out_norm, float* out_db, float* out_melpow,
           int32_t normalize, cudaStream_t stream, float* const* peers = nullptr, int n_peers = 0,
           float* mc_out = nullptr, bool pcm16 = false) {
    if (B == 0) return LM_OK;
    lm::KParams k = make_params(p);
    for (int r = 0; r < n_peers; ++r) k.peer[r] = peers[r];
    k.n_peer = n_peers; k.mc_out = mc_out;
    k.wave = wave; k.offset = reinterpret_cast<const long long*>(offset); k.length = length;
    k.aug = aug; k.noise = noise; k.out_norm = out_norm; k.out_db = out_db; k.out_melpow = out_melpow;
    k.B = B; k.normalize = normalize;
    const long long seq = p->launches.fetch_add(1);
    k.work_counter = p->d_counters + 2 * (seq % kCounters);   // zero now, left at zero by the kernel (no memset per launch)
    const int cap = p->max_ctas > 0 ? p->max_ctas : p->sm_count;
    // Small batches: fewer clips than 8-warp groups on the GPU.  Cut every clip into chunks of whole tiles so that
    // (almost) every group gets one chunk; statistics are combined with integer atomics (logmel_kernel.cuh).
    k.split = 1; k.tiles_per_chunk = p->n_tiles;
    const int groups = cap * lm::kGroups;
    if (normalize && p->split_override != 1 && B < groups && B <= kMaxSplitClips && p->n_tiles > 1) {
        int want = groups / B;
        if (p->split_override > 1) want = std::min(want, p->split_override);
        want = std::max(1, std::min(want, p->n_tiles));
        const int tpc = (p->n_tiles + want - 1) / want;
        k.tiles_per_chunk = tpc;
        k.split = (p->n_tiles + tpc - 1) / tpc;
    }
    if (!normalize && p->split_override != 1 && B < groups && p->n_tiles > 1) {   // nothing to combine: split freely
        const int want = std::max(1, std::min(groups / B, p->n_tiles));
        k.tiles_per_chunk = (p->n_tiles + want - 1) / want;
        k.split = (p->n_tiles + k.tiles_per_chunk - 1) / k.tiles_per_chunk;
    }
    if (k.split > 1 && normalize) {
        unsigned char* blk = p->d_split + static_cast<size_t>(seq % kCounters) * (kMaxSplitClips * 20);
        k.clip_stats = reinterpret_cast<unsigned long long*>(blk);
        k.clip_cnt = reinterpret_cast<int*>(blk + kMaxSplitClips * 16);
        // zero at plan creation; the group that completes a clip resets its entries
    }
    const int grid = std::min<int>(B * k.split, cap);
    const bool extra = (out_db != nullptr) || (out_melpow != nullptr);
    const bool device_noise = (aug != nullptr) && (noise == nullptr);   // clips may ask for Philox noise drawn in the kernel
    if (pcm16) {   // `wave` holds 16-bit samples (lm_forward_pcm16): expanded inside the staging, no decode kernel
        if (extra) return LM_ERR_UNSUPPORTED;
        if (p->n_fft == 2048) lm::logmel_kernel<2048, false, false, true><<<grid, lm::kThreads, p->smem_bytes, stream>>>(k);
        else lm::logmel_kernel<1024, false, false, true><<<grid, lm::kThreads, p->smem_bytes, stream>>>(k);
    } else if (p->n_fft == 2048) {
        if (extra) lm::logmel_kernel<2048, true><<<grid, lm::kThreads, p->smem_bytes, stream>>>(k);
        else if (device_noise) lm::logmel_kernel<2048, false, true><<<grid, lm::kThreads, p->smem_bytes, stream>>>(k);
        else lm::logmel_kernel<2048, false><<<grid, lm::kThreads, p->smem_bytes, stream>>>(k);
    } else {
        if (extra) lm::logmel_kernel<1024, true><<<grid, lm::kThreads, p->smem_bytes, stream>>>(k);
        else lm::logmel_kernel<1024, false><<<grid, lm::kThreads, p->smem_bytes, stream>>>(k);
    }
    LM_CUDA(cudaGetLastError());
    return LM_OK;
}

int ensure_slots(lm_plan* p) {
    if (p->slots_ready) return LM_OK;
    for (auto& s : p->slots) LM_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    p->slots_ready = true;
    return LM_OK;
}

template <class Tp>
int grow(Tp** ptr, size_t* cap, size_t want) {
    if (want <= *cap) return LM_OK;
    if (*ptr) LM_CUDA(cudaFree(*ptr));
    *ptr = nullptr; *cap = 0;
    LM_CUDA(cudaMalloc(reinterpret_cast<void**>(ptr), want * sizeof(Tp)));
    *cap = want;
    return LM_OK;
}

}  // namespace

extern "C" {

int lm_abi_version(void) { return LM_ABI_VERSION; }

const char* lm_strerror(int status) {
    switch (status) {
        case LM_OK: return "ok";
        case LM_ERR_INVALID_ARG: return "invalid argument";
        case LM_ERR_UNSUPPORTED: return "unsupported configuration (n_fft must be 1024 or 2048; hop even and <= n_fft/4; n_mels <= 256; window must be the periodic Hann window: w[n] + w[n + n_fft/2] = 1)";
        case LM_ERR_FILTERBANK: return "filterbank support too wide for the on-chip table";
        case LM_ERR_CUDA: return "CUDA runtime error (see lm_last_cuda_error)";
        case LM_ERR_NO_DEVICE: return "no usable CUDA device (an sm_100 GPU is required; there is no CPU fallback)";
        case LM_ERR_TOO_SHORT: return "target_len must exceed n_fft/2 (reflect padding)";
        default: return "unknown status";
    }
}

const char* lm_last_cuda_error(void) { return g_cuda_err; }

int lm_plan_create(const lm_config* cfg, int device, lm_plan** out_plan) {
    NvtxRange nvtx("lm_plan_create");
    if (!cfg || !out_plan || !cfg->window || !cfg->fb) return LM_ERR_INVALID_ARG;
    *out_plan = nullptr;
    if (cfg->n_fft != 2048 && cfg->n_fft != 1024) return LM_ERR_UNSUPPORTED;
    if (cfg->hop < 2 || (cfg->hop & 1) || cfg->hop > cfg->n_fft / 4) return LM_ERR_UNSUPPORTED;
    if (cfg->n_mels < 1 || cfg->n_mels > 256) return LM_ERR_UNSUPPORTED;
    if (cfg->target_len <= cfg->n_fft / 2) return LM_ERR_TOO_SHORT;
    if (!(cfg->amin > 0.f)) return LM_ERR_INVALID_ARG;
    // The kernel keeps half of the window and fuses it with the first butterfly through w[n + N/2] = 1 - w[n], which
    // holds for the periodic Hann window of the reference (torch.hann_window, TA/transforms/_transforms.py:86-87) and
    // for nothing else in common use: refuse other windows instead of computing wrong spectra.
    for (int n = 0; n < cfg->n_fft / 2; ++n)
        if (!(fabsf(cfg->window[n] + cfg->window[n + cfg->n_fft / 2] - 1.0f) <= 2e-6f)) return LM_ERR_UNSUPPORTED;

    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device < 0 || device >= n_dev) {
        cudaGetLastError();
        return LM_ERR_NO_DEVICE;
    }
    cudaDeviceProp prop{};
    LM_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return LM_ERR_NO_DEVICE;   // the fatbin holds sm_100a code only
    LM_CUDA(cudaSetDevice(device));

    lm_plan* p = new (std::nothrow) lm_plan();
    if (!p) return LM_ERR_INVALID_ARG;
    p->device = device;
    p->n_fft = cfg->n_fft; p->hop = cfg->hop; p->n_mels = cfg->n_mels; p->T = cfg->target_len;
    p->n_freqs = cfg->n_fft / 2 + 1;
    p->frames = 1 + cfg->target_len / cfg->hop;
    p->tile_f = (cfg->n_fft == 2048) ? lm::Geo<2048>::TILE_F : lm::Geo<1024>::TILE_F;
    p->n_tiles = (p->frames + p->tile_f - 1) / p->tile_f;
    p->ns = (((p->tile_f - 1) * p->hop + p->n_fft) + 3) & ~3;
    p->db_mult = cfg->db_multiplier; p->amin = cfg->amin; p->db_offset = cfg->db_offset;
    p->norm_eps = cfg->norm_eps;
    p->floor_db = cfg->db_multiplier * log10f(cfg->amin) - cfg->db_offset;
    p->sm_count = prop.multiProcessorCount;

    // ---- banded filterbank as mma.sync A fragments (scaled by 1/4: the kernel produces 4 |X|^2) ----
    // Mel tile mt = filters [8mt, 8mt+8).  Its band starts at kb (first non-zero bin, rounded down to
    // 4) and is walked in steps of 16 bins = two MMA k-steps.  For step d and k-step s, lane
    // (g = lane/4, tg = lane%4) holds one float4 = the A fragment (a0, a1, a2, a3) =
    // (head w0, residual w0, head w1, residual w1) with w_j = fb[kb + 16d + 4tg + 2s + j][8mt + g]:
    // MMA rows 0-7 carry the TF32 head (low 13 mantissa bits cleared), rows 8-15 the residual w - head
    // (13 significant bits, of which the tensor core keeps 11), and the k-permutation matches the
    // kernel's LDS.128 loads of the power rows.
    const int n_mt = (p->n_mels + 7) / 8;
    lm::MelTable tab{};
    std::vector<float4> melw;
    int nnz = 0;
    for (int k = 0; k < p->n_freqs; ++k)
        for (int m = 0; m < p->n_mels; ++m) nnz += cfg->fb[static_cast<size_t>(k) * p->n_mels + m] != 0.0f;
    p->fb_nnz = nnz;
    auto fbv = [&](int k, int m) -> float {
        return (k < p->n_freqs && m < p->n_mels) ? 0.25f * cfg->fb[static_cast<size_t>(k) * p->n_mels + m] : 0.0f;
    };
    auto head = [](float w) -> float {
        uint32_t u;
        memcpy(&u, &w, 4);
        u &= 0xffffe000u;
        float h;
        memcpy(&h, &u, 4);
        return h;
    };
    const int row_cap = (p->n_fft == 2048) ? lm::kRowFloats : (lm::kRowFloats - lm::kPbOff);
    for (int mt = 0; mt < n_mt; ++mt) {
        int lo = -1, hi = -1;
        for (int k = 0; k < p->n_freqs; ++k)
            for (int g = 0; g < 8; ++g)
                if (fbv(k, 8 * mt + g) != 0.0f) { if (lo < 0) lo = k; hi = k; }
        int kb = 0, ndk = 0;
        if (lo >= 0) { kb = lo & ~3; ndk = (hi + 1 - kb + 15) / 16; }
        while (ndk > 0 && kb > 0 && kb + 16 * ndk > row_cap) kb -= 4;   // keep the padded band inside the row
        if (kb + 16 * ndk > row_cap) { free_plan(p); return LM_ERR_FILTERBANK; }
        tab.kb[mt] = kb; tab.ndk[mt] = ndk; tab.off[mt] = static_cast<int>(melw.size() / 64);
        for (int d = 0; d < ndk; ++d)
            for (int s = 0; s < 2; ++s)
                for (int lane = 0; lane < 32; ++lane) {
                    const int g = lane >> 2, tg = lane & 3, k0 = kb + 16 * d + 4 * tg + 2 * s, m = 8 * mt + g;
                    const float w0 = fbv(k0, m), w1 = fbv(k0 + 1, m);
                    melw.push_back(make_float4(head(w0), w0 - head(w0), head(w1), w1 - head(w1)));
                }
    }
    if (melw.empty()) melw.assign(64, make_float4(0.f, 0.f, 0.f, 0.f));
    p->n_dk = static_cast<int>(melw.size() / 64);
    if (p->n_dk > lm::kMaxDk) { free_plan(p); return LM_ERR_FILTERBANK; }
    // longest-processing-time assignment of mel tiles to the 8 warps of a group (both groups use the
    // same table); warps 2s and 2s+1... of one scheduler are warp % 4, balanced too
    {
        for (auto& w : tab.warp_tile)
            for (int& t : w) t = -1;
        std::vector<int> order(n_mt);
        for (int i = 0; i < n_mt; ++i) order[i] = i;
        std::sort(order.begin(), order.end(), [&](int a, int b) { return tab.ndk[a] > tab.ndk[b]; });
        int load_w[lm::kGroupWarps] = {0}, load_s[4] = {0, 0, 0, 0}, cnt_w[lm::kGroupWarps] = {0};
        for (int mt : order) {
            int best = -1;
            for (int w = 0; w < lm::kGroupWarps; ++w) {
                if (cnt_w[w] >= lm::kTileSlots) continue;
                if (best < 0) { best = w; continue; }
                const int a = load_w[w] * 64 + load_s[w & 3];
                const int b = load_w[best] * 64 + load_s[best & 3];
                // lightest warp first, then lightest scheduler
                if (a < b) best = w;
            }
            tab.warp_tile[best][cnt_w[best]++] = mt;
            load_w[best] += tab.ndk[mt] + 4;      // a tile's fixed cost (setup + epilogue) is worth ~4 steps
            load_s[best & 3] += tab.ndk[mt] + 4;
        }
    }

    // ---- twiddles ---------------------------------------------------------------------------
    std::vector<float2> tw(32 * lm::kTwRows), utw(512);
    const double two_pi = 6.283185307179586476925286766559;
    for (int r = 0; r < lm::kTwRows; ++r) {
        const int k1 = (lm::kTwRows == 31) ? r + 1 : (r < 3 ? r + 1 : 4 * (r - 2));   // 1..31, or 1, 2, 3, 4, 8, ..., 28
        for (int n2 = 0; n2 < 32; ++n2) {
            const double a = two_pi * static_cast<double>(k1 * n2) / 1024.0;
            tw[r * 32 + n2] = make_float2(static_cast<float>(cos(a)), static_cast<float>(-sin(a)));
        }
    }
    for (int k = 0; k < 512; ++k) {
        const double a = two_pi * static_cast<double>(k) / 2048.0;
        utw[k] = make_float2(static_cast<float>(cos(a)), static_cast<float>(sin(a)));
    }

    auto up = [&](void** dst, const void* src, size_t bytes) -> int {
        LM_CUDA(cudaMalloc(dst, bytes));
        LM_CUDA(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
        return LM_OK;
    };
    int rc = LM_OK;
    if ((rc = up(reinterpret_cast<void**>(&p->d_window), cfg->window, sizeof(float) * p->n_fft)) ||
        (rc = up(reinterpret_cast<void**>(&p->d_tw), tw.data(), sizeof(float2) * tw.size())) ||
        (rc = up(reinterpret_cast<void**>(&p->d_utw), utw.data(), sizeof(float2) * utw.size())) ||
        (rc = up(reinterpret_cast<void**>(&p->d_melw), melw.data(), sizeof(float4) * melw.size())) ||
        (rc = up(reinterpret_cast<void**>(&p->d_tab), &tab, sizeof(tab))) ||
        (rc = (cudaMalloc(&p->d_counters, sizeof(int) * 2 * kCounters) == cudaSuccess &&
               cudaMemset(p->d_counters, 0, sizeof(int) * 2 * kCounters) == cudaSuccess &&
               cudaMalloc(&p->d_split, static_cast<size_t>(kCounters) * kMaxSplitClips * 20) == cudaSuccess &&
               cudaMemset(p->d_split, 0, static_cast<size_t>(kCounters) * kMaxSplitClips * 20) == cudaSuccess)
                  ? LM_OK : cuda_fail(cudaGetLastError(), "work counters"))) {
        free_plan(p);
        return rc;
    }

    // ---- shared memory --------------------------------------------------------------------
    cudaError_t e;
    {
        const size_t want = (p->n_fft == 2048) ? lm::Smem<2048>::total(p->ns, p->n_dk) : lm::Smem<1024>::total(p->ns, p->n_dk);
        if (want > prop.sharedMemPerBlockOptin) { free_plan(p); return LM_ERR_FILTERBANK; }
    }
    if (p->n_fft == 2048) {
        p->smem_bytes = lm::Smem<2048>::total(p->ns, p->n_dk);
        e = cudaFuncSetAttribute(lm::logmel_kernel<2048, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(p->smem_bytes));
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(lm::logmel_kernel<2048, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(p->smem_bytes));
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(lm::logmel_kernel<2048, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(p->smem_bytes));
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(lm::logmel_kernel<2048, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(p->smem_bytes));
    } else {
        p->smem_bytes = lm::Smem<1024>::total(p->ns, p->n_dk);
        e = cudaFuncSetAttribute(lm::logmel_kernel<1024, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(p->smem_bytes));
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(lm::logmel_kernel<1024, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(p->smem_bytes));
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(lm::logmel_kernel<1024, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(p->smem_bytes));
    }
    if (e != cudaSuccess) { free_plan(p); return cuda_fail(e, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)"); }
    *out_plan = p;
    return LM_OK;
}

int lm_plan_destroy(lm_plan* plan) { return free_plan(plan); }

#if LM_TIMING
// debug builds only (tools/build_variants.py "timing"): per-warp phase cycle counters of the last launch
int lm_debug_timing(long long* host_out, int n) {
    cudaDeviceSynchronize();
    return cudaMemcpyFromSymbol(host_out, lm::g_timing, sizeof(long long) * n) == cudaSuccess ? LM_OK : LM_ERR_CUDA;
}
#endif

int lm_plan_frames(const lm_plan* plan) { return plan ? plan->frames : LM_ERR_INVALID_ARG; }

int lm_plan_info(const lm_plan* plan, lm_info* info) {
    if (!plan || !info) return LM_ERR_INVALID_ARG;
    info->abi_version = LM_ABI_VERSION;
    info->frames = plan->frames;
    info->n_freqs = plan->n_freqs;
    info->sm_count = plan->max_ctas > 0 ? plan->max_ctas : plan->sm_count;
    info->threads_per_cta = lm::kThreads;
    info->smem_bytes = static_cast<int32_t>(plan->smem_bytes);
    info->fb_nnz = plan->fb_nnz;
    info->tma_staging = plan->use_tma;
    info->bytes_per_clip = 4LL * plan->T + 4LL * plan->n_mels * plan->frames;
    return LM_OK;
}

int lm_plan_set(lm_plan* plan, const char* key, int value) {
    if (!plan || !key) return LM_ERR_INVALID_ARG;
    if (!strcmp(key, "tma")) { plan->use_tma = value ? 1 : 0; return LM_OK; }
    if (!strcmp(key, "host_chunk_clips")) { if (value < 0) return LM_ERR_INVALID_ARG; plan->host_chunk_clips = value; return LM_OK; }
    if (!strcmp(key, "stagger_ns")) { if (value < 0) return LM_ERR_INVALID_ARG; plan->stagger_ns = value; return LM_OK; }
    if (!strcmp(key, "max_ctas")) { if (value < 0) return LM_ERR_INVALID_ARG; plan->max_ctas = value; return LM_OK; }
    if (!strcmp(key, "split")) { if (value < 0) return LM_ERR_INVALID_ARG; plan->split_override = value; return LM_OK; }
    return LM_ERR_INVALID_ARG;
}

int64_t lm_plan_launch_count(const lm_plan* plan) { return plan ? plan->launches.load() : 0; }

int lm_forward(lm_plan* plan, const float* wave, const int64_t* offset, const int32_t* length, int32_t B,
               const lm_aug* aug, const float* noise, float* out_norm, float* out_db, float* out_melpow,
               int32_t normalize, void* cuda_stream) {
    if (!plan || B < 0) return LM_ERR_INVALID_ARG;
    if (B == 0) return LM_OK;
    if (!wave || !offset || !length || !out_norm) return LM_ERR_INVALID_ARG;
    NvtxRange nvtx("lm_forward");
    int dev = -1;
    LM_CUDA(cudaGetDevice(&dev));
    if (dev != plan->device) LM_CUDA(cudaSetDevice(plan->device));
    const int rc = launch(plan, wave, offset, length, B, aug, noise, out_norm, out_db, out_melpow, normalize,
                          static_cast<cudaStream_t>(cuda_stream));
    if (dev != plan->device) cudaSetDevice(dev);
    return rc;
}

int lm_forward_pcm16(lm_plan* plan, const int16_t* pcm, const int64_t* offset, const int32_t* length, int32_t B,
                     const lm_aug* aug, const float* noise, float* out_norm, int32_t normalize, void* cuda_stream) {
    if (!plan || B < 0) return LM_ERR_INVALID_ARG;
    if (B == 0) return LM_OK;
    if (!pcm || !offset || !length || !out_norm) return LM_ERR_INVALID_ARG;
    NvtxRange nvtx("lm_forward_pcm16");
    int dev = -1;
    LM_CUDA(cudaGetDevice(&dev));
    if (dev != plan->device) LM_CUDA(cudaSetDevice(plan->device));
    const int rc = launch(plan, reinterpret_cast<const float*>(pcm), offset, length, B, aug, noise, out_norm, nullptr, nullptr, normalize,
                          static_cast<cudaStream_t>(cuda_stream), nullptr, 0, nullptr, true);
    if (dev != plan->device) cudaSetDevice(dev);
    return rc;
}

int lm_forward_gather(lm_plan* plan, const float* wave, const int64_t* offset, const int32_t* length, int32_t B,
                      const lm_aug* aug, const float* noise, float* out_slice, float* const* peer_slices,
                      int32_t n_peers, float* mc_slice, void* cuda_stream) {
    if (!plan || B < 0 || n_peers < 0 || n_peers > lm::kMaxPeers) return LM_ERR_INVALID_ARG;
    if (B == 0) return LM_OK;
    if (!wave || !offset || !length || !out_slice || (n_peers > 0 && !peer_slices)) return LM_ERR_INVALID_ARG;
    // the gather rides on the 16-byte stores of the normalisation pass
    if ((static_cast<size_t>(plan->n_mels) * plan->frames) % 4 != 0) return LM_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(out_slice) & 15u) || (reinterpret_cast<uintptr_t>(mc_slice) & 15u)) return LM_ERR_INVALID_ARG;
    for (int r = 0; r < n_peers; ++r)
        if (!peer_slices[r] || (reinterpret_cast<uintptr_t>(peer_slices[r]) & 15u)) return LM_ERR_INVALID_ARG;
    int dev = -1;
    LM_CUDA(cudaGetDevice(&dev));
    if (dev != plan->device) LM_CUDA(cudaSetDevice(plan->device));
    const int rc = launch(plan, wave, offset, length, B, aug, noise, out_slice, nullptr, nullptr, 1,
                          static_cast<cudaStream_t>(cuda_stream), peer_slices, n_peers, mc_slice);
    if (dev != plan->device) cudaSetDevice(dev);
    return rc;
}

int lm_resize_finish(const float* in, int32_t B, int32_t n_mels, int32_t frames_in, int32_t frames_out,
                     const lm_aug* aug, float* out, int32_t normalize, float norm_eps, void* cuda_stream) {
    if (B < 0 || n_mels < 1 || frames_in < 1 || frames_out < 1) return LM_ERR_INVALID_ARG;
    if (B == 0) return LM_OK;
    if (!in || !out || in == out) return LM_ERR_INVALID_ARG;
    DeviceOf guard(out);
    lm::resize_finish_kernel<<<B, lm::kAuxThreads, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
        in, out, aug, n_mels, frames_in, frames_out, normalize, norm_eps);
    LM_CUDA(cudaGetLastError());
    return LM_OK;
}

int lm_pcm16_roundtrip(const float* in, float* out, int64_t n, void* cuda_stream) {
    if (n < 0) return LM_ERR_INVALID_ARG;
    if (n == 0) return LM_OK;
    if (!in || !out) return LM_ERR_INVALID_ARG;
    DeviceOf guard(out);
    const int blocks = static_cast<int>(std::min<int64_t>((n + 1023) / 1024, 148 * 8));
    lm::pcm16_roundtrip_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(in, out, n);
    LM_CUDA(cudaGetLastError());
    return LM_OK;
}

// ---- polyphase sinc resampler (torchaudio.transforms.Resample defaults) ---------------------------------------
struct lm_resampler {
    int device = 0, orig = 1, neu = 1;   // after division by the gcd
    int width = 0, ntaps = 0;
    float* d_taps = nullptr;
    int* d_k0 = nullptr;
    // tiled kernel: taps padded to nt4 float4 per phase; shared memory for the widest input span of a 1024-output chunk
    float4* d_taps4 = nullptr;
    int nt4 = 0, k0max = 0;
    size_t tile_smem = 0;     // 0 = the span does not fit: the untiled kernel is used
};

int lm_resampler_create(int32_t orig_freq, int32_t new_freq, int device, lm_resampler** out) {
    if (!out || orig_freq < 1 || new_freq < 1) return LM_ERR_INVALID_ARG;
    *out = nullptr;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device < 0 || device >= n_dev) {
        cudaGetLastError();
        return LM_ERR_NO_DEVICE;
    }
    LM_CUDA(cudaSetDevice(device));
    int a = orig_freq, b = new_freq;
    while (b) { const int t = a % b; a = b; b = t; }
    const int o = orig_freq / a, q = new_freq / a;
    // TA/functional/functional.py _get_sinc_resample_kernel with lowpass_filter_width = 6, rolloff = 0.99,
    // sinc_interp_hann, float64 index arithmetic, kernel rounded to float32 at the end
    const int lpw = 6;
    const double base_freq = std::min(o, q) * 0.99;
    const int width = static_cast<int>(ceil(lpw * o / base_freq));
    const int K = 2 * width + o;
    const double pi = 3.14159265358979323846;
    std::vector<float> dense(static_cast<size_t>(q) * K);
    for (int p = 0; p < q; ++p)
        for (int k = 0; k < K; ++k) {
            double t = (static_cast<double>(-p) / q + static_cast<double>(k - width) / o) * base_freq;
            t = std::min(std::max(t, -static_cast<double>(lpw)), static_cast<double>(lpw));
            const double c = cos(t * pi / lpw / 2.0);
            const double window = c * c;
            const double tp = t * pi;
            const double sinc = (tp == 0.0) ? 1.0 : sin(tp) / tp;
            dense[static_cast<size_t>(p) * K + k] = static_cast<float>(sinc * window * (base_freq / o));
        }
    // per phase keep the taps that are not negligible (|w| > 1e-12: the clamped tails are ~1e-33)
    int ntaps = 1;
    std::vector<int> k0(q), k1(q);
    for (int p = 0; p < q; ++p) {
        int lo = K, hi = -1;
        for (int k = 0; k < K; ++k)
            if (fabsf(dense[static_cast<size_t>(p) * K + k]) > 1e-12f) { lo = std::min(lo, k); hi = k; }
        if (hi < 0) { lo = 0; hi = 0; }
        k0[p] = lo; k1[p] = hi;
        ntaps = std::max(ntaps, hi - lo + 1);
    }
    std::vector<float> taps(static_cast<size_t>(q) * ntaps, 0.0f);
    for (int p = 0; p < q; ++p) {
        if (k0[p] + ntaps > K) k0[p] = std::max(0, K - ntaps);
        for (int i = 0; i < ntaps && k0[p] + i < K; ++i) taps[static_cast<size_t>(p) * ntaps + i] = dense[static_cast<size_t>(p) * K + k0[p] + i];
    }
    lm_resampler* r = new (std::nothrow) lm_resampler();
    if (!r) return LM_ERR_INVALID_ARG;
    r->device = device; r->orig = o; r->neu = q; r->width = width; r->ntaps = ntaps;
    r->nt4 = (ntaps + 3) / 4;
    r->k0max = *std::max_element(k0.begin(), k0.end());
    std::vector<float> taps4(static_cast<size_t>(q) * r->nt4 * 4, 0.0f);
    for (int p = 0; p < q; ++p)
        for (int i = 0; i < ntaps; ++i) taps4[(static_cast<size_t>(p) * r->nt4) * 4 + i] = taps[static_cast<size_t>(p) * ntaps + i];
    {
        const size_t span = static_cast<size_t>((lm::kRsChunk + q - 1) / q + 1) * o + r->k0max + 4 * r->nt4 + 8;
        r->tile_smem = span * sizeof(float) <= 96 * 1024 ? span * sizeof(float) : 0;
        if (const char* e = getenv("LM_RESAMPLE_UNTILED")) { if (e[0] == '1') r->tile_smem = 0; }   // tests compare the two kernels
        if (r->tile_smem > 48 * 1024 &&
            cudaFuncSetAttribute(lm::resample_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(r->tile_smem)) != cudaSuccess) {
            cudaGetLastError();
            r->tile_smem = 0;
        }
    }
    if (cudaMalloc(&r->d_taps4, sizeof(float) * taps4.size()) != cudaSuccess ||
        cudaMemcpy(r->d_taps4, taps4.data(), sizeof(float) * taps4.size(), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMalloc(&r->d_taps, sizeof(float) * taps.size()) != cudaSuccess ||
        cudaMalloc(&r->d_k0, sizeof(int) * q) != cudaSuccess ||
        cudaMemcpy(r->d_taps, taps.data(), sizeof(float) * taps.size(), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(r->d_k0, k0.data(), sizeof(int) * q, cudaMemcpyHostToDevice) != cudaSuccess) {
        const int rc = cuda_fail(cudaGetLastError(), "resampler tables");
        cudaFree(r->d_taps); cudaFree(r->d_k0); cudaFree(r->d_taps4);
        delete r;
        return rc;
    }
    *out = r;
    return LM_OK;
}

int lm_resampler_destroy(lm_resampler* r) {
    if (!r) return LM_OK;
    cudaSetDevice(r->device);
    cudaFree(r->d_taps); cudaFree(r->d_k0); cudaFree(r->d_taps4);
    delete r;
    return LM_OK;
}

int64_t lm_resampler_out_len(const lm_resampler* r, int64_t in_len) {
    if (!r || in_len < 0) return LM_ERR_INVALID_ARG;
    return (static_cast<int64_t>(r->neu) * in_len + r->orig - 1) / r->orig;   // ceil(new * len / orig)
}

int lm_resample_rows(const lm_resampler* r, const float* in, int64_t in_len, int64_t in_stride, int32_t n_rows,
                     float* out, int64_t out_stride, void* cuda_stream) {
    if (!r || in_len < 0 || n_rows < 0) return LM_ERR_INVALID_ARG;
    const int64_t n_out = lm_resampler_out_len(r, in_len);
    if (n_out == 0 || n_rows == 0) return LM_OK;
    if (!in || !out || in_stride < in_len || out_stride < n_out) return LM_ERR_INVALID_ARG;
    DeviceOf guard(out);
    const int threads = 256;
    const long long blocks = (n_out + threads - 1) / threads;
    if (blocks > 0x7fffffffLL) return LM_ERR_INVALID_ARG;
    for (int32_t r0 = 0; r0 < n_rows && r->tile_smem; r0 += 65535) {   // tiled: one input span per 1024 outputs
        const int32_t nr = std::min<int32_t>(65535, n_rows - r0);
        const long long tblocks = (n_out + lm::kRsChunk - 1) / lm::kRsChunk;
        lm::resample_tiled_kernel<<<dim3(static_cast<unsigned>(tblocks), static_cast<unsigned>(nr)), lm::kRsThreads, r->tile_smem,
                                    static_cast<cudaStream_t>(cuda_stream)>>>(
            in + static_cast<int64_t>(r0) * in_stride, in_len, in_stride, out + static_cast<int64_t>(r0) * out_stride, n_out,
            out_stride, r->d_taps4, r->d_k0, r->orig, r->neu, r->nt4, r->width, r->k0max);
    }
    for (int32_t r0 = 0; r0 < n_rows && !r->tile_smem; r0 += 65535) {   // grid.y limit
        const int32_t nr = std::min<int32_t>(65535, n_rows - r0);
        lm::resample_kernel<<<dim3(static_cast<unsigned>(blocks), static_cast<unsigned>(nr)), threads, 0,
                              static_cast<cudaStream_t>(cuda_stream)>>>(
            in + static_cast<int64_t>(r0) * in_stride, in_len, in_stride, out + static_cast<int64_t>(r0) * out_stride, n_out,
            out_stride, r->d_taps, r->d_k0, r->orig, r->neu, r->ntaps, r->width);
    }
    LM_CUDA(cudaGetLastError());
    return LM_OK;
}

int lm_resample(const lm_resampler* r, const float* in, int64_t in_len, float* out, void* cuda_stream) {
    return lm_resample_rows(r, in, in_len, in_len, 1, out, lm_resampler_out_len(r, in_len), cuda_stream);
}

int lm_amplitude_to_db(const float* in, float* out, int64_t n, float multiplier, float amin, float db_offset, void* cuda_stream) {
    if (n < 0 || !(amin > 0.f)) return LM_ERR_INVALID_ARG;
    if (n == 0) return LM_OK;
    if (!in || !out) return LM_ERR_INVALID_ARG;
    DeviceOf guard(out);
    const int blocks = static_cast<int>(std::min<int64_t>((n + 255) / 256, 148 * 8));
    lm::amplitude_to_db_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
        in, out, n, static_cast<float>(static_cast<double>(multiplier) * 0.30102999566398119521), amin, db_offset);
    LM_CUDA(cudaGetLastError());
    return LM_OK;
}

int lm_pcm16_decode(const int16_t* in, float* out, int64_t n, void* cuda_stream) {
    if (n < 0) return LM_ERR_INVALID_ARG;
    if (n == 0) return LM_OK;
    if (!in || !out || (reinterpret_cast<uintptr_t>(in) & 15u) || (reinterpret_cast<uintptr_t>(out) & 15u)) return LM_ERR_INVALID_ARG;
    DeviceOf guard(out);
    const int blocks = static_cast<int>(std::min<int64_t>((n / 8 + 255) / 256 + 1, 148 * 8));
    lm::pcm16_decode_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(in, out, n);
    LM_CUDA(cudaGetLastError());
    return LM_OK;
}

static int forward_host_impl(lm_plan* plan, const void* wave_any, bool pcm16, int64_t total_samples, const int64_t* offset,
                             const int32_t* length, int32_t B, const lm_aug* aug, const float* noise, float* out,
                             int32_t normalize);

int lm_forward_host(lm_plan* plan, const float* wave, int64_t total_samples, const int64_t* offset,
                    const int32_t* length, int32_t B, const lm_aug* aug, const float* noise, float* out,
                    int32_t normalize) {
    return forward_host_impl(plan, wave, false, total_samples, offset, length, B, aug, noise, out, normalize);
}

int lm_forward_host_pcm16(lm_plan* plan, const int16_t* pcm, int64_t total_samples, const int64_t* offset,
                          const int32_t* length, int32_t B, const lm_aug* aug, const float* noise, float* out,
                          int32_t normalize) {
    return forward_host_impl(plan, pcm, true, total_samples, offset, length, B, aug, noise, out, normalize);
}

static int forward_host_impl(lm_plan* plan, const void* wave_any, bool pcm16, int64_t total_samples, const int64_t* offset,
                             const int32_t* length, int32_t B, const lm_aug* aug, const float* noise, float* out,
                             int32_t normalize) {
    const float* wave = static_cast<const float*>(wave_any);
    const int16_t* pcm = static_cast<const int16_t*>(wave_any);
    if (!plan || B < 0 || total_samples < 0) return LM_ERR_INVALID_ARG;
    if (B == 0) return LM_OK;
    if (!wave || !offset || !length || !out) return LM_ERR_INVALID_ARG;
    for (int i = 0; i < B; ++i)
        if (length[i] < 0 || offset[i] < 0 || offset[i] + length[i] > total_samples) return LM_ERR_INVALID_ARG;
    NvtxRange nvtx(pcm16 ? "lm_forward_host_pcm16" : "lm_forward_host");
    std::lock_guard<std::mutex> one_caller(plan->host_mu);   // the staging slots belong to the plan
    int dev = -1;
    LM_CUDA(cudaGetDevice(&dev));
    LM_CUDA(cudaSetDevice(plan->device));
    int rc = ensure_slots(plan);
    if (rc) return rc;

    const size_t clip_elems = static_cast<size_t>(plan->n_mels) * plan->frames;
    // chunk so that a few chunks are in flight: ~64 MB of waveform, at least 2 SM-waves of clips
    const int auto_chunk = std::max(2 * plan->sm_count, static_cast<int>((64u << 20) / (4u * static_cast<size_t>(plan->T))));
    const int chunk = std::max(1, std::min<int>(B, plan->host_chunk_clips > 0 ? plan->host_chunk_clips : auto_chunk));
    int slot_i = 0;
    for (int c0 = 0; c0 < B && rc == LM_OK; c0 += chunk, slot_i = (slot_i + 1) % kSlots) {
        const int n = std::min(chunk, B - c0);
        HostSlot& s = plan->slots[slot_i];
        if ((rc = cudaStreamSynchronize(s.stream) == cudaSuccess ? LM_OK : cuda_fail(cudaGetLastError(), "slot sync"))) break;
        int64_t lo = INT64_MAX, hi = 0;
        for (int i = c0; i < c0 + n; ++i) {
            if (length[i] == 0) continue;
            lo = std::min(lo, offset[i]);
            hi = std::max(hi, offset[i] + length[i]);
        }
        if (lo == INT64_MAX) { lo = 0; hi = 0; }
        lo &= ~int64_t(7);   // keep 16-byte alignment of clip starts (TMA staging path; 8 PCM samples per int4)
        const size_t n_wave = static_cast<size_t>(hi - lo);
        constexpr size_t kMetaPerClip = sizeof(long long) + sizeof(lm_aug) + sizeof(int);
        if (n > s.cap_clips) {
            cudaFree(s.d_meta);
            if (s.h_meta) cudaFreeHost(s.h_meta);
            s.d_meta = nullptr; s.h_meta = nullptr; s.cap_clips = 0;
            if (cudaMalloc(&s.d_meta, kMetaPerClip * n) != cudaSuccess || cudaMallocHost(&s.h_meta, kMetaPerClip * n) != cudaSuccess) {
                rc = cuda_fail(cudaGetLastError(), "slot metadata alloc");
                break;
            }
            s.cap_clips = n;
        }
        // the blocks are laid out for the slot's capacity so that the three arrays keep their alignment
        const size_t cap = static_cast<size_t>(s.cap_clips);
        long long* h_off = reinterpret_cast<long long*>(s.h_meta);
        lm_aug* h_aug = reinterpret_cast<lm_aug*>(s.h_meta + sizeof(long long) * cap);
        int* h_len = reinterpret_cast<int*>(s.h_meta + (sizeof(long long) + sizeof(lm_aug)) * cap);
        const long long* d_off = reinterpret_cast<const long long*>(s.d_meta);
        const lm_aug* d_aug = reinterpret_cast<const lm_aug*>(s.d_meta + sizeof(long long) * cap);
        const int* d_len = reinterpret_cast<const int*>(s.d_meta + (sizeof(long long) + sizeof(lm_aug)) * cap);
        if (!pcm16 && (rc = grow(&s.d_wave, &s.cap_wave, std::max<size_t>(n_wave + 8, 16)))) break;
        if (pcm16 && (rc = grow(&s.d_pcm, &s.cap_pcm, std::max<size_t>(n_wave + 8, 16)))) break;
        if ((rc = grow(&s.d_out, &s.cap_out, clip_elems * n))) break;
        if (noise && (rc = grow(&s.d_noise, &s.cap_noise, static_cast<size_t>(plan->T) * n))) break;
        for (int i = 0; i < n; ++i) h_off[i] = length[c0 + i] ? offset[c0 + i] - lo : 0;
        memcpy(h_len, length + c0, sizeof(int) * n);
        if (aug) memcpy(h_aug, aug + c0, sizeof(lm_aug) * n);

        cudaError_t e = cudaMemcpyAsync(s.d_meta, s.h_meta, kMetaPerClip * cap, cudaMemcpyHostToDevice, s.stream);
        if (e != cudaSuccess) { rc = cuda_fail(e, "H2D copy (metadata)"); break; }
        if (n_wave && !pcm16) e = cudaMemcpyAsync(s.d_wave, wave + lo, sizeof(float) * n_wave, cudaMemcpyHostToDevice, s.stream);
        if (n_wave && pcm16) e = cudaMemcpyAsync(s.d_pcm, pcm + lo, sizeof(int16_t) * n_wave, cudaMemcpyHostToDevice, s.stream);
        if (e == cudaSuccess && noise)
            e = cudaMemcpyAsync(s.d_noise, noise + static_cast<size_t>(c0) * plan->T, sizeof(float) * plan->T * n,
                                cudaMemcpyHostToDevice, s.stream);
        if (e != cudaSuccess) { rc = cuda_fail(e, "H2D copy"); break; }
        rc = launch(plan, pcm16 ? reinterpret_cast<const float*>(s.d_pcm) : s.d_wave, reinterpret_cast<const int64_t*>(d_off), d_len, n,
                    aug ? d_aug : nullptr, noise ? s.d_noise : nullptr, s.d_out, nullptr, nullptr, normalize, s.stream, nullptr, 0, nullptr,
                    pcm16);
        if (rc) break;
        e = cudaMemcpyAsync(out + clip_elems * c0, s.d_out, sizeof(float) * clip_elems * n, cudaMemcpyDeviceToHost, s.stream);
        if (e != cudaSuccess) { rc = cuda_fail(e, "D2H copy"); break; }
    }
    for (auto& s : plan->slots) {
        const cudaError_t e = cudaStreamSynchronize(s.stream);
        if (e != cudaSuccess && rc == LM_OK) rc = cuda_fail(e, "pipeline drain");
    }
    cudaSetDevice(dev);
    return rc;
}

}  // extern "C"
