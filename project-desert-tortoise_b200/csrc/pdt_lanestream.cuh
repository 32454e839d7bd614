// pdt_lanestream.cuh — "lane streams": 32 independent serial recurrences per warp, each walking its own contiguous
// stretch of a stream in HBM, fed and drained by the TMA (sm_100a, device only).
//
// Why: the time-tiled loops (PLL track core, AGC) give every lane a private tile.  Read directly, a warp-wide load
// touches 32 different 128-byte lines for 16 bytes each; the L1/LSU serialises those 32 tag look-ups, and with a
// handful of such warps per SM the kernels became LSU-bound at 2-4x the latency of the arithmetic chain (ncu r01d:
// k_pll_core / k_agc_core 50 % long-scoreboard, 240-280 cycles per sample against a 63-cycle chain).  Here every lane
// issues ONE bulk copy (cp.async.bulk, SASS UBLKCP) per LS_R-sample piece of its tile into its own shared-memory row,
// completion is counted on an mbarrier, results go back with one bulk store per row, and the only LSU traffic left is
// the conflict-free LDS.128/STS.128 of the rows (row stride ≡ 4 words mod 32 banks).
#pragma once

#include "pdt_tiled.cuh"

namespace pdt {
namespace tiled {

#ifdef __CUDACC__

constexpr int LS_R  = 64;            // samples per lane per stage (256-byte rows)
constexpr int LS_RS = LS_R + 4;      // row stride in floats: 272 B, 16-byte aligned, bank-conflict-free for 128-bit accesses
constexpr int LS_NIN = 2;            // input stages (the next row is requested while the current one is consumed)
constexpr int LS_NOUT = 2;

struct __align__(128) LaneStreamSmem {
    float in[LS_NIN][32][LS_RS];
    float out[LS_NOUT][32][LS_RS];
    unsigned long long full[LS_NIN];
};

// per-warp handle: the shared-memory rows + how many pieces this warp has streamed so far in this kernel (an mbarrier may
// not be re-initialised without mbarrier.inval, so the barriers are set up once per kernel and the running piece count
// selects stage and phase parity across calls)
struct LaneStream { LaneStreamSmem *sm; unsigned pieces; };

__device__ __forceinline__ uint32_t ls_smem(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ls_mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ls_smem(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void ls_mbar_arrive(unsigned long long *bar)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(ls_smem(bar)) : "memory");
}
__device__ __forceinline__ void ls_mbar_arrive_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(ls_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ls_mbar_wait(unsigned long long *bar, unsigned parity, unsigned dbg_c = 0, unsigned dbg_n = 0, unsigned dbg_m = 0)
{
    // try_wait suspends the thread for a hardware time slice when the phase is not complete; a copy that never lands
    // (it cannot, short of a programming error) traps after ~2 s instead of hanging the device
    unsigned done;
    unsigned long long t0 = 0;
    for (unsigned spins = 0;; spins++) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(ls_smem(bar)), "r"(parity) : "memory");
        if (done) break;
        if ((spins & 63u) == 63u) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t0 == 0) t0 = t;
            else if (t - t0 > 2000000000ull) {
#ifdef LS_DEBUG
                if ((threadIdx.x & 31) < 3)
                printf("ls_mbar_wait stuck: block %d thread %d bar@%u parity %u word %016llx c %u nch %u maxch %u other %016llx\n", (int)blockIdx.x, (int)threadIdx.x,
                       ls_smem(bar), parity, *(volatile unsigned long long *)bar, dbg_c, dbg_n, dbg_m, *(volatile unsigned long long *)((ls_smem(bar) & 8) ? bar - 1 : bar + 1));
#endif
                __trap();
            }
        }
    }
}
__device__ __forceinline__ void ls_bulk_load(void *dst_smem, const void *src_gmem, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(ls_smem(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(ls_smem(bar)) : "memory");
}
__device__ __forceinline__ void ls_bulk_store(void *dst_gmem, const void *src_smem, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(ls_smem(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ls_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void ls_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ls_bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void ls_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Run `step` over in[s0, s1) for this lane; outputs are written to out[sb, s1) (s0 <= sb <= s1: [s0, sb) is a warm-up
// whose outputs are discarded).  s0 and sb must be multiples of 4 (16-byte aligned pieces); s1 may be anything — the
// last piece is then loaded/stored rounded up to 4 samples, inside the padding every workspace row has.
// A lane with nothing to do passes s0 == s1.  All 32 lanes of the warp must call this together.
//   Step:  void quad(const float4 &v, float4 &o);   float one(float v);
// STORE = false: a pure warm-up (sb == s1), no output row is written at all.
// once per kernel, by all 32 lanes of the warp
__device__ __forceinline__ LaneStream lane_stream_init(LaneStreamSmem *sm, const int lane)
{
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < LS_NIN; s++) ls_mbar_init(&sm->full[s], 32);
    }
    ls_fence_async();
    __syncwarp();
    LaneStream h; h.sm = sm; h.pieces = 0;
    return h;
}

template <bool STORE, class Step>
__device__ __forceinline__ void lane_stream(LaneStream &h, const int lane, const float *__restrict__ in, float *__restrict__ out,
                                            const u64 s0, const u64 sb, const u64 s1, Step &step)
{
    LaneStreamSmem &sm = *h.sm;
    const unsigned nch = (unsigned)((s1 - s0 + LS_R - 1) / LS_R);
    unsigned maxch = nch;
#pragma unroll
    for (int o = 16; o; o >>= 1) { const unsigned t = __shfl_xor_sync(0xffffffffu, maxch, o); maxch = t > maxch ? t : maxch; }
    if (maxch == 0) return;
    const unsigned p0 = h.pieces;          // running piece index of this call's piece 0
    h.pieces += maxch;
    __syncwarp();

    auto issue = [&](unsigned c) {
        const int st = (int)((p0 + c) % LS_NIN);
        if (c < nch) {
            const u64 g = s0 + (u64)c * LS_R;
            const unsigned cnt = (s1 - g < (u64)LS_R) ? (unsigned)(s1 - g) : (unsigned)LS_R;
            const unsigned bytes = ((cnt + 3u) & ~3u) * 4u;
            ls_mbar_arrive_tx(&sm.full[st], bytes);
            ls_bulk_load(&sm.in[st][lane][0], in + g, bytes, &sm.full[st]);
        } else ls_mbar_arrive(&sm.full[st]);
    };
#pragma unroll
    for (int c = 0; c < LS_NIN - 1; c++) if ((unsigned)c < maxch) issue((unsigned)c);

    for (unsigned c = 0; c < maxch; c++) {
        const int st = (int)((p0 + c) % LS_NIN), so = (int)(c % LS_NOUT);
        if (c + LS_NIN - 1 < maxch) issue(c + LS_NIN - 1);        // its stage was consumed by every lane in round c-1
        ls_mbar_wait(&sm.full[st], ((p0 + c) / LS_NIN) & 1u, c, nch, maxch);
        if (c < nch) {
            ls_bulk_wait_read<LS_NOUT - 1>();                     // the store that last used out[so] has read its row
            const u64 g = s0 + (u64)c * LS_R;
            const unsigned cnt = (s1 - g < (u64)LS_R) ? (unsigned)(s1 - g) : (unsigned)LS_R;
            const float *__restrict__ ir = sm.in[st][lane];
            float *__restrict__ orow = sm.out[so][lane];
            unsigned j = 0;
            if (cnt == (unsigned)LS_R) {
#pragma unroll 4
                for (; j < (unsigned)LS_R; j += 4) { const float4 v = ld4(ir + j); float4 o; step.quad(v, o); if (STORE) st4(orow + j, o); }
            } else {
                for (; j + 4 <= cnt; j += 4) { const float4 v = ld4(ir + j); float4 o; step.quad(v, o); if (STORE) st4(orow + j, o); }
                for (; j < cnt; j++) { const float o = step.one(ir[j]); if (STORE) orow[j] = o; }
                if (STORE) for (; j & 3u; j++) orow[j] = 0.0f;
            }
            const unsigned sf = (sb > g) ? ((sb - g < (u64)cnt) ? (unsigned)(sb - g) : cnt) : 0u;
            if (STORE && cnt > sf) {
                ls_fence_async();                                 // generic-proxy row writes -> visible to the bulk store
                ls_bulk_store(out + g + sf, orow + sf, ((cnt - sf + 3u) & ~3u) * 4u);
            }
            ls_bulk_commit();
        }
        __syncwarp();
    }
    ls_bulk_wait_all();
    __syncwarp();
}

#endif // __CUDACC__

} // namespace tiled
} // namespace pdt
