// logmel_pipe_kernel.cuh -- the large-batch kernel: SMs specialised by role, fed through L2 (v8).
//
// Why.  The frame transform (window, 2048-point real FFT, 4|X|^2) runs at 432 cycles per frame per SM when 16 warps do
// nothing else (FMA pipe 85 % busy); the mel phase's LDS.128 / HMMA / integer stream costs it half its speed whenever
// the two share an SM, whether they run one after the other (logmel_kernel.cuh) or side by side in specialised warps
// (DESIGN.md section 4.3).  So here they do not share an SM.  One persistent, cooperatively launched grid; of every
// kPipeFan + 1 consecutive CTAs
//
//   kPipeFan "transform" CTAs   two teams of 8 warps; a team owns a clip at a time (dealt dynamically from a global
//               counter) and walks its tiles: the warp that is last to pull its frame out of a staging buffer restages
//               that buffer two tiles ahead (bulk copy + the reflect / zero padding by its own 32 lanes), so there is no
//               barrier and no idle warp; every warp writes its frame's 4|X|^2 row (1025 floats) straight from registers to
//               the team's ring of tile slots in global memory -- which stays in L2 (30 MB for the whole grid) -- and
//               signals the slot one frame later, when the stores have long landed;
//   1 "mel" CTA   15 warps that each own one mel tile (8 filters; the two smallest share a warp) and keep its banded
//               A fragments IN REGISTERS for the whole kernel -- no filterbank in shared memory, one LDS.128 per 16 bins
//               instead of three -- plus a loader warp that polls the six producer rings, pulls ready tiles into a
//               three-deep shared-memory ring with one bulk copy each and hands the slots back.  Epilogue, statistics
//               and the clip-end normalisation are those of logmel_kernel.cuh: outputs are bit-identical to it.
//
// Scope: n_fft = 2048, plain clips (no augmentation records), no extra outputs, no fused gather, at most 16 mel tiles of at
// most kPipeMaxSteps 16-bin steps.  Everything else -- and every batch too small to fill the grid -- takes logmel_kernel.cuh.
#pragma once
#include "logmel_kernel.cuh"

namespace lm {

constexpr int kPipeFan = 3;                    // transform CTAs per mel CTA
constexpr int kPipeSlots = 4;                  // tile slots per producer team in global memory
constexpr int kPipeRowG = 1040;                // floats per power row in the rings (1025 + the padded tail of the last band; == 16 mod 32)
constexpr int kPipeSlotFloats = 8 * kPipeRowG; // one tile = 8 rows
constexpr int kPipeBufs = 4;                   // tiles in the mel CTA's shared memory
constexpr int kPipeMelWarps = 15;
constexpr int kPipeMelThreads = kPipeMelWarps * 32;
constexpr int kPipeMaxSteps = 12;              // 16-bin steps a mel warp can keep in registers (8 registers each)
constexpr int kPipeRings = 2 * kPipeFan;       // producer teams per mel CTA

struct PipeHdr { int clip, tile, flags, pad; };   // flags: 1 a tile, 4 the last real tile of its clip (pad = silent_from: the mel CTA fills the rest), 8 the team has no more work

// per-launch control block in global memory (zeroed by a memset node in front of the kernel)
struct PipeCtl {
    int work_counter;          // next clip
    int pad[3];
    // then, per producer team: PipeHdr hdr[kPipeSlots]; int ready[kPipeSlots]; int consumed; int pad[3]
};
constexpr int kPipeTeamCtlInts = 4 * kPipeSlots + kPipeSlots + 4;
__host__ __device__ inline size_t pipe_ctl_bytes(int n_teams) { return sizeof(PipeCtl) + sizeof(int) * kPipeTeamCtlInts * static_cast<size_t>(n_teams); }

struct PipeMelTable {
    int warp_tile[kPipeMelWarps][2];   // mel tiles of each mel warp (-1 = none); their steps add up to <= kPipeMaxSteps
};

struct PipeParams {
    KParams k;
    float* ring;            // [n_teams][kPipeSlots][8][kPipeRowG]; columns >= 1025 are zero and stay zero
    int* ctl;               // PipeCtl
    const PipeMelTable* mel_warps;
    int n_groups;           // grid = n_groups * (kPipeFan + 1)
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {   // one probe, no loop
    uint32_t ok;
    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
#ifndef LM_PIPE_DEBUG
#define LM_PIPE_DEBUG 0
#endif
#if LM_PIPE_DEBUG
__device__ unsigned long long g_pipe_time[160 * 8];   // per CTA: cycles summed over its warps, 8 categories
#define PT_BEGIN long long pt_last_ = clock64(); long long pt_acc_[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define PT(slot) { const long long n_ = clock64(); pt_acc_[slot] += n_ - pt_last_; pt_last_ = n_; }
#define PT_END if ((threadIdx.x & 31) == 0) { for (int i_ = 0; i_ < 8; ++i_) atomicAdd(&g_pipe_time[blockIdx.x * 8 + i_], (unsigned long long)pt_acc_[i_]); }
__device__ int g_pipe_dbg[4096];   // [0] = entries; then {block, warp, site, a, b, c} per stuck wait (first 600)
__device__ __noinline__ void pipe_dbg(int site, int a, int b, int c) {
    if ((threadIdx.x & 31) != 0) return;
    const int i = atomicAdd(&g_pipe_dbg[0], 1);
    if (i < 600) { int* e = g_pipe_dbg + 8 + 6 * i; e[0] = blockIdx.x; e[1] = threadIdx.x >> 5; e[2] = site; e[3] = a; e[4] = b; e[5] = c; }
}
#define PIPE_WATCH_BEGIN const long long t_watch_ = clock64();
#define PIPE_WATCH(site, a, b, c) if (clock64() - t_watch_ > 400000000LL) { pipe_dbg(site, a, b, c); break; }
#define PIPE_MBAR_WAIT(bar, parity, site, a, b) { PIPE_WATCH_BEGIN while (!mbar_try(bar, parity)) { PIPE_WATCH(site, a, b, 0) } }
#else
#define PT_BEGIN
#define PT(slot)
#define PT_END
#define PIPE_WATCH_BEGIN
#define PIPE_WATCH(site, a, b, c)
#define PIPE_MBAR_WAIT(bar, parity, site, a, b) mbar_wait(bar, parity);
#endif
__device__ __forceinline__ int* pipe_team_ctl(int* ctl, int team) { return ctl + sizeof(PipeCtl) / 4 + team * kPipeTeamCtlInts; }
__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ld_relaxed(const int* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_relaxed_add(int* p, int v) {
    asm volatile("red.relaxed.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// shared -> global bulk copy (TMA) in the thread's bulk group; bulk_wait_all: every committed group of this thread has been
// written to global memory (so a flag set afterwards cannot be seen before the data)
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void st_release(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_add(int* p, int v) {
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// shared memory of a transform CTA
struct PipeFftSmem {
    static constexpr size_t kBar = 0;                      // sb_full[2 teams][2]
    static constexpr size_t kCnt = kBar + 32;              // consumption counters [2][2]
    static constexpr size_t kSeq = kCnt + 16;              // scheduling tickets [2]
    static constexpr size_t kDesc = kSeq + 16;             // PipeHdr (+ team sequence number in pad) [2][2]
    static constexpr size_t kCtx = kDesc + 64;             // ClipCtx [2 teams] + team schedule {tile, t_end, done, n_staged}
    static constexpr size_t kSched = kCtx + 2 * kCtxSlot;
    static constexpr size_t kWin = kSched + 32;
    static constexpr size_t kTw = kWin + sizeof(float) * 1024;
    static constexpr size_t kUtw = kTw + sizeof(float2) * 32 * kTwRows;
    static constexpr size_t kScr = (kUtw + sizeof(float2) * 512 + 15) & ~size_t(15);
    static constexpr size_t kSbuf = kScr + sizeof(float) * 16 * kRowFloats;
    static __host__ __device__ size_t total(int ns) { return kSbuf + sizeof(float) * 4 * static_cast<size_t>(ns); }
};
// shared memory of a mel CTA
struct PipeMelSmem {
    static constexpr size_t kBar = 0;                      // rows_full[kPipeBufs], rows_empty[kPipeBufs], filterbank copy
    static constexpr size_t kDesc = 128;                   // PipeHdr (+ ring in pad) [kPipeBufs]
    static constexpr size_t kRed = kDesc + 16 * kPipeBufs; // reduction scratch 2 x 16 long long + bcast[4]
    static constexpr size_t kStat = kRed + 256 + 16;       // [kPipeRings][16 warps] longlong2: fixed-point (sum, sum of squares) per ring and warp
    static constexpr size_t kRows = (kStat + sizeof(longlong2) * kPipeRings * 16 + 127) & ~size_t(127);
    static constexpr size_t kMelw = kRows + sizeof(float) * kPipeBufs * kPipeSlotFloats;
    static __host__ __device__ size_t total(int n_dk) { return kMelw + sizeof(float4) * 64 * static_cast<size_t>(n_dk); }
};

// Everything the staging warp of a producer team needs (pointers into the CTA's shared memory and the team's control block)
struct PipeTeam {
    uint64_t* sb_full; volatile int* s_seq; volatile PipeHdr* s_desc; ClipCtx* s_ctx; volatile int* s_sched; float* sb;
    PipeHdr* g_hdr; int* g_ready; int* g_consumed;
};
// ---- staging of the team's buffer use number bufq (ONE warp: the one that emptied the buffer, or the prologue warp) ----
// A ticket (s_seq) keeps the tiles in order: the stager of use bufq waits until use bufq - 1 has been scheduled.
// Silent tiles (the zero-padded tail of a short plain clip) never enter the pipeline: the clip's last REAL tile
// carries silent_from and the mel CTA writes the floor for the rest; a clip without any real tile is finished here.
// Out of line on purpose: one warp in eight runs it once per tile, and inlined it would spill the FFT's registers for all.
template <int NFFT>
__device__ __noinline__ void pipe_stage_buffer(const PipeParams& pp, const PipeTeam& tm, int bufq) {
    constexpr int TILE_F = 8;
    constexpr int HALF = NFFT / 2;
    const KParams& p = pp.k;
    const int lane = threadIdx.x & 31;
    const int T = p.T, hop = p.hop, frames = p.frames;
    const size_t clip_elems = static_cast<size_t>(p.n_mels) * frames;
    uint64_t* const sb_full = tm.sb_full; volatile int* const s_seq = tm.s_seq; volatile PipeHdr* const s_desc = tm.s_desc;
    ClipCtx* const s_ctx = tm.s_ctx; volatile int* const s_sched = tm.s_sched; float* const sb = tm.sb;
    PipeHdr* const g_hdr = tm.g_hdr; int* const g_ready = tm.g_ready; int* const g_consumed = tm.g_consumed;

            { PIPE_WATCH_BEGIN while (*s_seq != bufq) { PIPE_WATCH(1, bufq, *s_seq, 0) } }
            const int buf = bufq & 1;
            volatile PipeHdr* const d = &s_desc[buf];
            for (;;) {
                int flags = 0, tile_ = 0, first_stop = 0;
                if (lane == 0) {
                    int done = s_sched[2];
                    if (!done && s_sched[0] == s_sched[1]) {            // the team needs a clip
                        const int nxt = atomicAdd(pp.ctl, 1);
                        if (nxt < p.B) {
                            load_clip(p, nxt, s_ctx, NFFT);
                            const int te = s_ctx->t_end < s_ctx->silent_from ? s_ctx->t_end : s_ctx->silent_from;
                            s_sched[0] = s_ctx->t_begin; s_sched[1] = te;
                            if (te <= s_ctx->t_begin) flags = 2;         // nothing but padding
                        } else {
                            done = 1; s_sched[2] = 1; first_stop = 1;
                        }
                    }
                    if (done) {
                        flags = 8;
                    } else if (flags == 0) {
                        tile_ = s_sched[0];
                        flags = 1 | (tile_ + 1 == s_sched[1] ? 4 : 0);
                        s_sched[0] = tile_ + 1;
                    }
                }
                flags = __shfl_sync(0xffffffffu, flags, 0);
                tile_ = __shfl_sync(0xffffffffu, tile_, 0);
                __syncwarp();
                const ClipCtx c = *s_ctx;                                 // private copy: the next stager may replace the context
                if (flags & 2) {
                    // an all-padding clip: every feature is the floor, the normalised clip is exactly 0 (as logmel_kernel.cuh)
                    float* __restrict__ o = p.out_norm + static_cast<size_t>(c.clip) * clip_elems;
                    const float v = p.normalize ? 0.0f : p.floor_db;
                    for (int i = lane; i < static_cast<int>(clip_elems); i += 32) o[i] = v;
                    __syncwarp();
                    continue;
                }
                int oseq = 0;
                if (lane == 0 && (!(flags & 8) || first_stop)) { oseq = s_sched[3]; s_sched[3] = oseq + 1; }   // ring sequence number of this item
                oseq = __shfl_sync(0xffffffffu, oseq, 0);
                if (flags & 8) {
                    // the team is out of work: ONE marker ends its ring, and its warps are told to leave (both staging buffers get the word)
                    if (lane == 0) {
                        if (first_stop) {
                            { PIPE_WATCH_BEGIN while (ld_acquire(g_consumed) + kPipeSlots <= oseq) { PIPE_WATCH(2, oseq, ld_acquire(g_consumed), 0) } }
                            st_release(&g_ready[oseq % kPipeSlots], 16);   // 16 instead of 8 arrivals: the ring ends here
                        }
                        d->flags = 8;
                        __threadfence_block();
                        *s_seq = bufq + 1;
                        mbar_arrive(&sb_full[buf]);
                    }
                    __syncwarp();
                    return;
                }
                if (lane == 0) { __threadfence_block(); *s_seq = bufq + 1; }   // scheduling done: the next buffer's stager may go on
                const int tf = tile_ * TILE_F;
                const int nf = (frames - tf) < TILE_F ? (frames - tf) : TILE_F;
                const int need = (nf - 1) * hop + NFFT;
                const int j0 = tf * hop - HALF;
                float* const sbuf = sb + buf * p.ns;
                int e_lo = 0, cnt = 0;
                {
                    const int lo = j0 < 0 ? -j0 : 0;
                    int hi = c.lc - j0;
                    if (hi > need) hi = need;
                    if (p.use_tma && hi > lo && (reinterpret_cast<uintptr_t>(c.src + j0 + lo) & 15u) == 0 && (lo & 3) == 0) { e_lo = lo; cnt = (hi - lo) & ~3; }
                }
                const int rest = p.ns - cnt;
                for (int i0 = 0; i0 < rest; i0 += 8 * 32) {
                    float v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int idx = i0 + u * 32 + lane;
                        const int e = idx < e_lo ? idx : idx + cnt;
                        int j = j0 + e;
                        if (j < 0) j = -j;
                        else if (j >= T) j = 2 * (T - 1) - j;
                        v[u] = (idx < rest && e < need && j >= 0 && j < c.lc) ? __ldg(c.src + j) : 0.0f;
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int idx = i0 + u * 32 + lane;
                        if (idx < rest) sbuf[idx < e_lo ? idx : idx + cnt] = v[u];
                    }
                }
                __syncwarp();
                if (lane == 0) {
                    d->clip = c.clip; d->tile = tile_ | (c.silent_from << 16); d->flags = flags; d->pad = oseq;
                    if (cnt != 0) {
                        fence_proxy_async();
                        mbar_expect_tx(&sb_full[buf], static_cast<uint32_t>(cnt) * 4u);
                        bulk_g2s(sbuf + e_lo, c.src + j0 + e_lo, static_cast<uint32_t>(cnt) * 4u, &sb_full[buf]);
                    } else {
                        mbar_arrive(&sb_full[buf]);
                    }
                }
                __syncwarp();
                return;
            }
}

template <int NFFT>
__global__ void __launch_bounds__(kThreads, 1) logmel_pipe_kernel(const PipeParams pp) {
    static_assert(NFFT == 2048, "the pipeline kernel is the n_fft = 2048 path");
    constexpr int TILE_F = 8;
    constexpr int HALF = NFFT / 2;
    const KParams& p = pp.k;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = static_cast<int>(blockIdx.x) / (kPipeFan + 1), role = static_cast<int>(blockIdx.x) % (kPipeFan + 1);
    const int T = p.T, hop = p.hop, frames = p.frames, n_mels = p.n_mels;
    const size_t clip_elems = static_cast<size_t>(n_mels) * frames;

    if (role != 0) {
        // =====================================================================================================
        // transform CTA: two teams of 8 warps
        // =====================================================================================================
        using L = PipeFftSmem;
        const int team = warp >> 3, fw = warp & 7;
        const int gteam = (grp * kPipeFan + (role - 1)) * 2 + team;          // producer team index in the grid
        uint64_t* const sb_full = reinterpret_cast<uint64_t*>(smem_raw + L::kBar) + 2 * team;
        int* const s_cnt = reinterpret_cast<int*>(smem_raw + L::kCnt) + 2 * team;
        volatile int* const s_seq = reinterpret_cast<volatile int*>(smem_raw + L::kSeq) + team;
        volatile PipeHdr* const s_desc = reinterpret_cast<volatile PipeHdr*>(smem_raw + L::kDesc) + 2 * team;
        ClipCtx* const s_ctx = reinterpret_cast<ClipCtx*>(smem_raw + L::kCtx + team * kCtxSlot);
        volatile int* const s_sched = reinterpret_cast<volatile int*>(smem_raw + L::kSched) + 4 * team;   // tile, t_end, done, staged
        float* const s_win = reinterpret_cast<float*>(smem_raw + L::kWin);
        float2* const s_tw = reinterpret_cast<float2*>(smem_raw + L::kTw);
        float2* const s_utw = reinterpret_cast<float2*>(smem_raw + L::kUtw);
        float* const scr = reinterpret_cast<float*>(smem_raw + L::kScr) + warp * kRowFloats;
        float* const sb = reinterpret_cast<float*>(smem_raw + L::kSbuf) + static_cast<size_t>(team) * 2 * p.ns;
        int* const tctl = pipe_team_ctl(pp.ctl, gteam);
        PipeHdr* const g_hdr = reinterpret_cast<PipeHdr*>(tctl);
        int* const g_ready = tctl + 4 * kPipeSlots;
        int* const g_consumed = g_ready + kPipeSlots;
        float* const g_ring = pp.ring + static_cast<size_t>(gteam) * kPipeSlots * kPipeSlotFloats;

        for (int i = tid; i < HALF; i += kThreads) s_win[i] = p.window[i];
        for (int i = tid; i < 32 * kTwRows; i += kThreads) s_tw[i] = p.tw[i];
        for (int i = tid; i < 512; i += kThreads) s_utw[i] = p.utw[i];
        for (int i = tid; i < 16 * kRowFloats; i += kThreads) reinterpret_cast<float*>(smem_raw + L::kScr)[i] = 0.0f;
        if (tid == 0) {
            for (int i = 0; i < 4; ++i) mbar_init(reinterpret_cast<uint64_t*>(smem_raw + L::kBar) + i, 1);
            fence_mbar_init();
        }
        if (tid < 4) reinterpret_cast<int*>(smem_raw + L::kCnt)[tid] = 0;
        if (tid < 2) reinterpret_cast<int*>(smem_raw + L::kSeq)[tid] = 0;
        if (tid < 8) reinterpret_cast<int*>(smem_raw + L::kSched)[tid] = 0;
        __syncthreads();

        PipeTeam tm;
        tm.sb_full = sb_full; tm.s_seq = s_seq; tm.s_desc = s_desc; tm.s_ctx = s_ctx; tm.s_sched = s_sched; tm.sb = sb;
        tm.g_hdr = g_hdr; tm.g_ready = g_ready; tm.g_consumed = g_consumed;
        if (fw == 0) { pipe_stage_buffer<NFFT>(pp, tm, 0); pipe_stage_buffer<NFFT>(pp, tm, 1); }   // prologue: the team's first two tiles

        PT_BEGIN
        float* pend_row = nullptr;      // the row this warp wrote last (its stores may still be in flight) ...
        int* pend_ready = nullptr;      // ... and the slot counter to bump once they have landed
#pragma unroll 1
        for (int q = 0;; ++q) {
            const int buf = q & 1;
            PIPE_MBAR_WAIT(&sb_full[buf], (q >> 1) & 1, 3, q, *s_seq)
            PT(0)
            const int flags = s_desc[buf].flags;
            if (flags & 8) break;
            const int oseq = s_desc[buf].pad, h_clip = s_desc[buf].clip, h_tile = s_desc[buf].tile;   // read before the buffer (and its descriptor) can be restaged
            const float* __restrict__ sbuf = sb + buf * p.ns;
            int seen_consumed = 0;
            if (lane == 0) seen_consumed = ld_relaxed(g_consumed);   // asked for now, needed before the row leaves: the round trip hides behind the FFT
            lm_f2 z[32];
            {
                const float2* __restrict__ s2 = reinterpret_cast<const float2*>(sbuf + fw * hop);
                const float2* __restrict__ w2 = reinterpret_cast<const float2*>(s_win);
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const float2 v1 = s2[32 * r + lane];
                    const float2 v2 = s2[32 * (r + 16) + lane];
                    const float2 w = w2[32 * r + lane];
                    const lm_f2 V1 = lm_pack(v1.x, v1.y), V2 = lm_pack(v2.x, v2.y), W = lm_pack(w.x, w.y);
                    z[r] = lm_fma2(lm_sub2(V1, V2), W, V2);
                    z[r + 16] = lm_fma2(lm_add2(V1, V2), W, lm_pack(-v2.x, -v2.y));
                }
            }
            // the frame is in registers.  The warp that is last to say so restages the buffer two tiles ahead.
            __syncwarp();
            int last = 0;
            if (lane == 0) {
                last = (atomicAdd(&s_cnt[buf], 1) == 7);
                if (last) s_cnt[buf] = 0;
            }
            last = __shfl_sync(0xffffffffu, last, 0);
            lm_fft32_aos_from2(z);
            // the previous frame's row has landed by now: hand it to the mel CTA
            if (pend_ready != nullptr) {
                if (lane == 0) { bulk_wait_all(); red_relaxed_add(pend_ready, 1); }   // the row's bulk copy has been written: count it
                __syncwarp();                                                        // ... and has left the scratch row, which part 2 reuses
                pend_ready = nullptr;
            }
            PT(1)
            if (last) pipe_stage_buffer<NFFT>(pp, tm, q + 2);
            PT(2)

            float xr[32], xi[32];
            warp_cfft1024_part2(z, xr, xi, scr, s_tw, lane);
            // the slot must have been handed back by the mel CTA (kPipeSlots tiles ago)
            if (lane == 0) { PIPE_WATCH_BEGIN while (seen_consumed + kPipeSlots <= oseq) { seen_consumed = ld_relaxed(g_consumed); PIPE_WATCH(4, oseq, seen_consumed, q) } }
            __syncwarp();
            PT(3)
            float* __restrict__ row = g_ring + static_cast<size_t>(oseq % kPipeSlots) * kPipeSlotFloats + fw * kPipeRowG;
            if (fw == 0 && lane == 0) {   // the tile's header rides behind the spectrum of its first row
                int* const hd = reinterpret_cast<int*>(scr + 1028);
                hd[0] = h_clip; hd[1] = h_tile & 0xffff; hd[2] = flags; hd[3] = h_tile >> 16;   // [3]: silent_from
            }
            const int srcl = (32 - lane) & 31;
            const bool l0 = (lane == 0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float s_lr = l0 ? xr[(32 - i) & 31] : xr[31 - i];
                const float s_li = l0 ? xi[(32 - i) & 31] : xi[31 - i];
                const float s_hr = l0 ? xr[(16 - i) & 31] : xr[15 - i];
                const float s_hi = l0 ? xi[(16 - i) & 31] : xi[15 - i];
                const float b_lr = __shfl_sync(0xffffffffu, s_lr, srcl);
                const float b_li = __shfl_sync(0xffffffffu, s_li, srcl);
                const float b_hr = __shfl_sync(0xffffffffu, s_hr, srcl);
                const float b_hi = __shfl_sync(0xffffffffu, s_hi, srcl);
                const lm_f2 Ar = lm_pack(xr[i], xr[i + 16]), Ai = lm_pack(xi[i], xi[i + 16]);
                const lm_f2 Br = lm_pack(b_lr, b_hr), Bi = lm_pack(b_li, b_hi);
                const lm_f2 Er = lm_add2(Ar, Br), Ei = lm_sub2(Ai, Bi), Or = lm_add2(Ai, Bi), Oi = lm_sub2(Br, Ar);
                const float2 cs = s_utw[lane + 32 * i];
                const lm_f2 C = lm_pack(cs.x, -cs.y), S = lm_pack(cs.y, cs.x), nS = lm_pack(-cs.y, -cs.x);
                const lm_f2 Tr = lm_fma2(C, Or, lm_mul2(S, Oi));
                const lm_f2 Ti = lm_fma2(C, Oi, lm_mul2(nS, Or));
                const lm_f2 Ur = lm_add2(Er, Tr), Ui = lm_add2(Ei, Ti), Vr = lm_sub2(Er, Tr), Vi = lm_sub2(Ei, Ti);
                const lm_f2 PU = lm_fma2(Ur, Ur, lm_mul2(Ui, Ui)), PV = lm_fma2(Vr, Vr, lm_mul2(Vi, Vi));
                scr[lane + 32 * i] = lm_lo(PU);
                scr[lane + 32 * i + 512] = lm_hi(PU);
                scr[1024 - lane - 32 * i] = lm_lo(PV);
                scr[512 - lane - 32 * i] = lm_hi(PV);
            }
            {   // lane 0 only: bins 256 and 768 pair with each other (slots 8 and 24), twiddle pi/4
                const float ar = xr[8], ai = xi[8], br = xr[24], bi = xi[24];
                const float er = ar + br, ei = ai - bi, orr = ai + bi, oi = br - ar;
                const float c = 0.70710678118654752440f;
                const float tr = c * (orr + oi), ti = c * (oi - orr);
                const float ur = er + tr, ui = ei + ti, vr = er - tr, vi = ei - ti;
                if (l0) {
                    scr[256] = fmaf(ur, ur, ui * ui);
                    scr[768] = fmaf(vr, vr, vi * vi);
                }
            }
            PT(5)
            __syncwarp();
            PT(6)
            if (lane == 0) {   // the row (1025 floats + header words: 1032 copied; the mel bands read at most kPipeRowG columns, finite values) goes to the team's slot
                fence_proxy_async();
                bulk_s2g(row, scr, 1032u * 4u);
            }
            pend_row = row;
            pend_ready = &g_ready[oseq % kPipeSlots];
            PT(4)
        }
        PT_END
        (void)pend_row;
        if (pend_ready != nullptr && lane == 0) { bulk_wait_all(); red_relaxed_add(pend_ready, 1); }
        return;
    }

    // =========================================================================================================
    // mel CTA: warps 0-14 own the mel tiles (A fragments in registers), warp 15 loads ready tiles from the producers' rings
    // =========================================================================================================
    using M = PipeMelSmem;
    uint64_t* const rows_full = reinterpret_cast<uint64_t*>(smem_raw + M::kBar);
    uint64_t* const rows_empty = rows_full + kPipeBufs;
    volatile PipeHdr* const s_desc = reinterpret_cast<volatile PipeHdr*>(smem_raw + M::kDesc);
    long long* const red = reinterpret_cast<long long*>(smem_raw + M::kRed);
    float* const bcast = reinterpret_cast<float*>(smem_raw + M::kRed + 256);
    longlong2* const s_stat = reinterpret_cast<longlong2*>(smem_raw + M::kStat);
    float* const s_rows = reinterpret_cast<float*>(smem_raw + M::kRows);
    float4* const s_melw = reinterpret_cast<float4*>(smem_raw + M::kMelw);
    uint64_t* const mbar_fb = rows_full + 2 * kPipeBufs;
    if (tid == 0) {
        for (int b = 0; b < kPipeBufs; ++b) { mbar_init(&rows_full[b], 1); mbar_init(&rows_empty[b], kPipeMelWarps); }
        mbar_init(mbar_fb, 1);
        fence_mbar_init();
        mbar_expect_tx(mbar_fb, static_cast<uint32_t>(p.n_dk) * 1024u);   // the banded filterbank fragments, one bulk copy
        bulk_g2s(s_melw, p.melw, static_cast<uint32_t>(p.n_dk) * 1024u, mbar_fb);
    }
    for (int i = tid; i < kPipeRings * 16; i += kThreads) s_stat[i] = make_longlong2(0, 0);
    __syncthreads();

    if (warp == kPipeMelWarps) {
        // ---- loader warp: lane r < kPipeRings watches producer ring r ------------------------------------------
        int* const my_ctl = pipe_team_ctl(pp.ctl, grp * kPipeRings + (lane < kPipeRings ? lane : 0));
        int my_seq = 0;                               // next tile of my ring
        bool my_open = lane < kPipeRings;
        int held_ring[kPipeBufs], held_use[kPipeBufs];   // the ring whose slot is being copied into buffer b and has not been handed back yet (-1: none)
#pragma unroll
        for (int b = 0; b < kPipeBufs; ++b) { held_ring[b] = -1; held_use[b] = 0; }
        // a ring slot goes back to its producers as soon as its bulk copy has landed (the buffer's "full" phase completes)
        auto hand_back = [&](int block_b) {
#pragma unroll
            for (int b = 0; b < kPipeBufs; ++b) {
                if (held_ring[b] < 0) continue;
                bool landed = mbar_try(&rows_full[b], held_use[b] & 1);
                while (b == block_b && !landed) landed = mbar_try(&rows_full[b], held_use[b] & 1);
                if (landed) {
                    if (lane == held_ring[b]) red_release_add(my_ctl + 4 * kPipeSlots + kPipeSlots, 1);   // the lane that reset the slot's counter
                    held_ring[b] = -1;
                }
            }
        };
        int k = 0, rr = 0;
        PT_BEGIN
        PIPE_WATCH_BEGIN
        for (;;) {
            PIPE_WATCH(5, k, my_seq, 0)
            if (__ballot_sync(0xffffffffu, my_open) == 0u) break;
            hand_back(-1);
            // one round trip for all rings: is my ring's next slot complete?  (its header comes with it)
            bool ready = false;
            if (my_open) {
                const int v = ld_relaxed(my_ctl + 4 * kPipeSlots + my_seq % kPipeSlots);   // L2, no L1 invalidation: the rows come by bulk copy
                ready = (v == 8);
                if (v == 16) my_open = false;   // the team is done: its marker needs no buffer
            }
            unsigned mask = __ballot_sync(0xffffffffu, ready);
            if (mask == 0u) { __nanosleep(100); PT(5) continue; }
            PT(5)
            bool second_of_pair = false;
            while (mask != 0u) {
                // round robin over the ready rings
                const unsigned rot = (mask >> rr) | (mask << (32 - rr));
                const int r = (rr + __ffs(rr ? rot : mask) - 1) % 32;
                mask &= ~(1u << r);
                rr = (r + 1) % kPipeRings;
                const bool pair_first = !second_of_pair && mask != 0u;   // another tile follows at once: the mel warps take both together
                second_of_pair = pair_first;
                const int b = k % kPipeBufs;
                PIPE_MBAR_WAIT(&rows_empty[b], ((k / kPipeBufs) & 1) ^ 1, 6, k, b)    // the mel warps are done with the tile that was in this buffer ...
                hand_back(b);                                               // ... so its ring slot has been read: make sure it went back
                PT(6)
                if (lane == r) {
                    const int slot = my_seq % kPipeSlots;
                    s_desc[b].flags = pair_first ? 16 : 0; s_desc[b].pad = r;
                    my_ctl[4 * kPipeSlots + slot] = 0;                      // the counter is ready for the slot's next use
                    fence_proxy_async();
                    mbar_expect_tx(&rows_full[b], static_cast<uint32_t>(kPipeSlotFloats) * 4u);
                    bulk_g2s(s_rows + static_cast<size_t>(b) * kPipeSlotFloats,
                             pp.ring + (static_cast<size_t>(grp * kPipeRings + r) * kPipeSlots + slot) * kPipeSlotFloats,
                             static_cast<uint32_t>(kPipeSlotFloats) * 4u, &rows_full[b]);
                    ++my_seq;
                }
                held_ring[b] = r; held_use[b] = k / kPipeBufs;
                __syncwarp();
                ++k;
                PT(7)
            }
        }
        PT_END
        // every ring has ended: one last marker tells the mel warps
        {
            const int b = k % kPipeBufs;
            PIPE_MBAR_WAIT(&rows_empty[b], ((k / kPipeBufs) & 1) ^ 1, 7, k, b)
            if (lane == 0) { s_desc[b].flags = 8; mbar_arrive(&rows_full[b]); }
        }
        return;
    }

    // ---- mel warps --------------------------------------------------------------------------------------
    const int g = lane >> 2, tg = lane & 3;
    const int mt0 = pp.mel_warps->warp_tile[warp][0], mt1 = pp.mel_warps->warp_tile[warp][1];
    const int nd0 = mt0 >= 0 ? p.mel_table->ndk[mt0] : 0, nd1 = mt1 >= 0 ? p.mel_table->ndk[mt1] : 0;
    const int kb0 = mt0 >= 0 ? p.mel_table->kb[mt0] : 0, kb1 = mt1 >= 0 ? p.mel_table->kb[mt1] : 0;
    const int of0 = mt0 >= 0 ? p.mel_table->off[mt0] : 0, of1 = mt1 >= 0 ? p.mel_table->off[mt1] : 0;
    mbar_wait(mbar_fb, 0);
    const int mtid = tid;   // 0 .. kPipeMelThreads - 1
    PT_BEGIN
    // Items come one or two at a time: when the loader found two tiles ready it marks the first one (flag 16) and the
    // warps take both through the MMA loop together -- every A fragment serves 16 frames, two independent accumulator
    // chains per tile -- which is what a latency-bound loop of 48 dependent HMMAs needs.
#pragma unroll 1
    for (int k = 0;;) {
        const int b0 = k % kPipeBufs;
        PIPE_MBAR_WAIT(&rows_full[b0], (k / kPipeBufs) & 1, 8, k, b0)
        PT(0)
        const int flags0 = s_desc[b0].flags;
        if (flags0 & 8) break;
        const int n_it = (flags0 & 16) ? 2 : 1;
        const int b1 = (k + 1) % kPipeBufs;
        if (n_it == 2) PIPE_MBAR_WAIT(&rows_full[b1], ((k + 1) / kPipeBufs) & 1, 9, k, b1)
        int clip_[2], tile_[2], sfrom_[2], ring_[2], flg_[2];
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int b = it ? b1 : b0;
            const int4 hd = *reinterpret_cast<const int4*>(s_rows + static_cast<size_t>(b) * kPipeSlotFloats + 1028);   // behind row 0
            clip_[it] = hd.x; tile_[it] = hd.y; flg_[it] = hd.z; sfrom_[it] = hd.w;
            ring_[it] = s_desc[b].pad;
        }
        long long si_[2] = {0, 0}, qi_[2] = {0, 0};   // this thread's fixed-point sums of the item(s)
#pragma unroll
        for (int t2 = 0; t2 < 2; ++t2) {
            const int mt = t2 ? mt1 : mt0;
            if (mt < 0) continue;
            const int ndk = t2 ? nd1 : nd0;
            const float4* __restrict__ wp = s_melw + static_cast<size_t>(t2 ? of1 : of0) * 64 + lane;
            const float* const rp0 = s_rows + static_cast<size_t>(b0) * kPipeSlotFloats + g * kPipeRowG + 4 * tg + (t2 ? kb1 : kb0);
            const float* const rp1 = s_rows + static_cast<size_t>(b1) * kPipeSlotFloats + g * kPipeRowG + 4 * tg + (t2 ? kb1 : kb0);
            // four independent accumulator chains per item (head / residual of the power x the two k-steps of a 16-bin step):
            // a dependent HMMA.1688 comes back after ~100 cycles, the tensor pipe takes one every 8
            float acc_h[2][4], acc_l[2][4], acc_h2[2][4], acc_l2[2][4];
#pragma unroll
            for (int it = 0; it < 2; ++it)
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) { acc_h[it][q4] = 0.f; acc_l[it][q4] = 0.f; acc_h2[it][q4] = 0.f; acc_l2[it][q4] = 0.f; }
#pragma unroll 2
            for (int d = 0; d < ndk; ++d) {
                const float4 w1 = wp[64 * d], w2 = wp[64 * d + 32];
#pragma unroll
                for (int it = 0; it < 2; ++it) {
                    if (it < n_it) {
                        const float4 pv = *reinterpret_cast<const float4*>((it ? rp1 : rp0) + 16 * d);
                        const uint32_t p0 = tf32_hi(pv.x), p1 = tf32_hi(pv.y), p2 = tf32_hi(pv.z), p3 = tf32_hi(pv.w);
                        const lm_f2 r01 = lm_sub2(lm_pack(pv.x, pv.y), lm_pack(__uint_as_float(p0), __uint_as_float(p1)));
                        const lm_f2 r23 = lm_sub2(lm_pack(pv.z, pv.w), lm_pack(__uint_as_float(p2), __uint_as_float(p3)));
#ifdef LM_PIPE_X1
                        acc_h[it][0] += w1.x * pv.x + w2.y * lm_lo(r01) + w1.z * lm_hi(r23) + __uint_as_float(p0 ^ p1 ^ p2 ^ p3) + w1.y + w1.w + w2.x + w2.z + w2.w + lm_hi(r01) + lm_lo(r23);
                        continue;
#endif
                        mma_tf32(acc_h[it], __float_as_uint(w1.x), __float_as_uint(w1.y), __float_as_uint(w1.z), __float_as_uint(w1.w), p0, p1);
                        mma_tf32(acc_l[it], __float_as_uint(w1.x), __float_as_uint(w1.y), __float_as_uint(w1.z), __float_as_uint(w1.w),
                                 __float_as_uint(lm_lo(r01)), __float_as_uint(lm_hi(r01)));
                        mma_tf32(acc_h2[it], __float_as_uint(w2.x), __float_as_uint(w2.y), __float_as_uint(w2.z), __float_as_uint(w2.w), p2, p3);
                        mma_tf32(acc_l2[it], __float_as_uint(w2.x), __float_as_uint(w2.y), __float_as_uint(w2.z), __float_as_uint(w2.w),
                                 __float_as_uint(lm_lo(r23)), __float_as_uint(lm_hi(r23)));
                    }
                }
            }
#pragma unroll
            for (int it = 0; it < 2; ++it)
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) { acc_h[it][q4] += acc_h2[it][q4]; acc_l[it][q4] += acc_l2[it][q4]; }
            // epilogue (as logmel_kernel.cuh; plain clips: no masks)
            const int m = mt * 8 + g;
            const bool ok_m = m < n_mels;
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                if (it < n_it) {
                    const int tf = tile_[it] * TILE_F;
                    const int nf = (frames - tf) < TILE_F ? (frames - tf) : TILE_F;
                    float* __restrict__ out = p.out_norm + static_cast<size_t>(clip_[it]) * clip_elems;
                    const lm_f2 mp2 = lm_add2(lm_add2(lm_pack(acc_h[it][0], acc_h[it][1]), lm_pack(acc_h[it][2], acc_h[it][3])),
                                              lm_add2(lm_pack(acc_l[it][0], acc_l[it][1]), lm_pack(acc_l[it][2], acc_l[it][3])));
                    const int fl = 2 * tg, tt = tf + fl;
                    const int o = m * frames + tt;
                    const float mp0 = lm_lo(mp2), mp1 = lm_hi(mp2);
                    const float v0 = (mp0 <= p.amin) ? p.floor_db : fmaf(p.db_scale, lg2_ftz(mp0), -p.db_offset);
                    const float v1 = (mp1 <= p.amin) ? p.floor_db : fmaf(p.db_scale, lg2_ftz(mp1), -p.db_offset);
                    const bool ok0 = ok_m && (fl < nf), ok1 = ok_m && (fl + 1 < nf);
#ifdef LM_PIPE_X2
                    if (ok0 && v0 == 12345.f) out[o] = v0;
#else
                    if (ok0) out[o] = v0;
                    if (ok1) out[o + 1] = v1;
#endif
                    const float u0 = ok0 ? v0 : 0.0f, u1 = ok1 ? v1 : 0.0f;
                    // one fixed-point conversion per (thread, mel tile, item), as logmel_kernel.cuh: the same integers, bit-identical statistics
                    si_[it] += __double2ll_rn(static_cast<double>(u0 + u1) * static_cast<double>(kStatScaleS));
                    qi_[it] += __double2ll_rn(static_cast<double>(fmaf(u0, u0, fmaf(u1, u1, 0.0f))) * static_cast<double>(kStatScaleQ));
                }
            }
        }
        __syncwarp();
        if (lane == 0) { mbar_arrive(&rows_empty[b0]); if (n_it == 2) mbar_arrive(&rows_empty[b1]); }
        // integer sums: any order gives the same total
#pragma unroll
        for (int it = 0; it < 2; ++it) {
#ifdef LM_PIPE_X3
            if (it < n_it && si_[it] == 12345) {
#else
            if (it < n_it) {
#endif
                long long a = si_[it], c = qi_[it];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    a += __shfl_xor_sync(0xffffffffu, a, o);
                    c += __shfl_xor_sync(0xffffffffu, c, o);
                }
                if (lane == 0) { longlong2 st = s_stat[ring_[it] * 16 + warp]; st.x += a; st.y += c; s_stat[ring_[it] * 16 + warp] = st; }
            }
        }
        PT(1)

#pragma unroll 1
        for (int it = 0; it < n_it; ++it) {
            if (!(flg_[it] & 4)) continue;
            const int clip = clip_[it], ring = ring_[it], silent_from = sfrom_[it];
            float* __restrict__ out = p.out_norm + static_cast<size_t>(clip) * clip_elems;
            // ---- the clip's silent tail: tiles [silent_from, n_tiles) are the floor; threads 0-255 walk them exactly as the
            //      256 threads of a group of logmel_kernel.cuh do, so the fixed-point statistics are the same integers ----------
            if (mtid < kGroupThreads) {
                for (int st_tile = silent_from; st_tile < p.n_tiles; ++st_tile) {
                    const int stf = st_tile * TILE_F;
                    const int snf = (frames - stf) < TILE_F ? (frames - stf) : TILE_F;
                    float ss = 0.0f, qq = 0.0f;
                    for (int idx = mtid; idx < n_mels * TILE_F; idx += kGroupThreads) {
                        const int m = idx / TILE_F, f = idx - m * TILE_F;
                        if (f < snf) {
                            const float v = p.floor_db;
                            out[m * frames + stf + f] = v;
                            ss += v;
                            qq = fmaf(v, v, qq);
                        }
                    }
                    long long a = __double2ll_rn(static_cast<double>(ss) * static_cast<double>(kStatScaleS));
                    long long c = __double2ll_rn(static_cast<double>(qq) * static_cast<double>(kStatScaleQ));
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        a += __shfl_xor_sync(0xffffffffu, a, o);
                        c += __shfl_xor_sync(0xffffffffu, c, o);
                    }
                    if (lane == 0) { longlong2 st = s_stat[ring * 16 + warp]; st.x += a; st.y += c; s_stat[ring * 16 + warp] = st; }
                }
            }
            // ---- per-clip normalisation (the clip's tiles came in order through one ring) ------------------------------
            if (p.normalize) {
                asm volatile("bar.sync 1, %0;" ::"n"(kPipeMelThreads) : "memory");   // every warp's sums are in; also orders every mel thread's dB stores before the re-read below
                if (mtid == 0) {
                    long long si = 0, qi = 0;
                    for (int w = 0; w < kPipeMelWarps; ++w) { si += s_stat[ring * 16 + w].x; qi += s_stat[ring * 16 + w].y; s_stat[ring * 16 + w] = make_longlong2(0, 0); }
                    const double sd = static_cast<double>(si) * (1.0 / kStatScaleS), qd = static_cast<double>(qi) * (1.0 / kStatScaleQ);
                    const double n = static_cast<double>(clip_elems);
                    const double mean = sd / n;
                    double var = (qd - sd * mean) / (n - 1.0);   // unbiased, as torch.std
                    if (!(var > 0.0)) var = 0.0;
                    bcast[0] = static_cast<float>(mean);
                    bcast[1] = static_cast<float>(sqrt(var)) + p.norm_eps;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kPipeMelThreads) : "memory");
                const float mean = bcast[0], inv = 1.0f / bcast[1];
                float4* __restrict__ o4 = reinterpret_cast<float4*>(out);
                const int n4 = (reinterpret_cast<uintptr_t>(out) & 15u) == 0 ? static_cast<int>(clip_elems >> 2) : 0;
                int i = mtid;
                for (; i + 3 * kPipeMelThreads < n4; i += 4 * kPipeMelThreads) {   // four loads in flight per thread
                    float4 v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) v[u] = __ldcg(o4 + i + u * kPipeMelThreads);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        v[u].x = (v[u].x - mean) * inv; v[u].y = (v[u].y - mean) * inv;
                        v[u].z = (v[u].z - mean) * inv; v[u].w = (v[u].w - mean) * inv;
                        o4[i + u * kPipeMelThreads] = v[u];
                    }
                }
                for (; i < n4; i += kPipeMelThreads) {
                    float4 v = __ldcg(o4 + i);
                    v.x = (v.x - mean) * inv; v.y = (v.y - mean) * inv; v.z = (v.z - mean) * inv; v.w = (v.w - mean) * inv;
                    o4[i] = v;
                }
                for (int j = (n4 << 2) + mtid; j < static_cast<int>(clip_elems); j += kPipeMelThreads)
                    out[j] = (__ldcg(out + j) - mean) * inv;
                asm volatile("bar.sync 1, %0;" ::"n"(kPipeMelThreads) : "memory");   // bcast is rewritten at the next clip end
            } else if (mtid < kPipeMelWarps) {
                s_stat[ring * 16 + mtid] = make_longlong2(0, 0);
            }
        }
        k += n_it;
        PT(2)
    }
    PT_END
}

}  // namespace lm
