// pdt_lanestream.cuh — "lane streams": 32 independent serial recurrences per warp, each walking its own contiguous
// stretch of a stream in HBM, staged through shared memory with warp-cooperative, fully coalesced copies (sm_100a).
//
// Why: the time-tiled loops (PLL track core, AGC) give every lane a private tile.  Read directly, a warp-wide load
// touches 32 different 128-byte lines for 16 bytes each; the L1/LSU serialises those 32 tag look-ups, and with a
// handful of such warps per SM the kernels became LSU-bound at 2-4x the latency of the arithmetic chain (ncu r01d:
// k_pll_core / k_agc_core 50 % long-scoreboard, 240-280 cycles per sample against a 63-cycle chain).
// Here the warp moves LS_R-sample pieces of all 32 tiles together: 16 lanes cover one 256-byte row with one
// cp.async (LDGSTS, 16 B per lane, L2 -> shared memory without a register round trip), so every request is two full
// rows; each lane then runs its recurrence over its own row with conflict-free LDS.128/STS.128 (row stride ≡ 4 words
// mod 32 banks), and the output rows go back the same cooperative way.  [A first version issued one cp.async.bulk (TMA)
// per lane and row; the bulk copy is a warp-uniform instruction, so 32 lanes meant a 32-trip issue loop of ~100 cycles
// per copy — 2.5x slower than the arithmetic (profiles/README.md, r01e).]
#pragma once

#include "pdt_tiled.cuh"

namespace pdt {
namespace tiled {

#ifdef __CUDACC__

constexpr int LS_R  = 64;            // samples per lane per stage (256-byte rows)
constexpr int LS_RS = LS_R + 4;      // row stride in floats: 272 B, 16-byte aligned, bank-conflict-free for 128-bit accesses
constexpr int LS_NIN = 2;            // input stages (the next piece is requested while the current one is consumed)

struct __align__(128) LaneStreamSmem {
    float in[LS_NIN][32][LS_RS];
    float out[32][LS_RS];
};

struct LaneStream { LaneStreamSmem *sm; };

__device__ __forceinline__ uint32_t ls_smem(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ls_cp16(uint32_t dst_smem, const void *src_gmem)
{
    // no "memory" clobber: the copy is ordered against its consumers by ls_wait + __syncwarp (both compiler barriers)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src_gmem));
}
__device__ __forceinline__ void ls_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void ls_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ LaneStream lane_stream_init(LaneStreamSmem *sm, const int lane)
{
    (void)lane;
    LaneStream h; h.sm = sm;
    return h;
}

__device__ __forceinline__ u64 ls_shfl64(u64 v, int src)
{
    const unsigned lo = __shfl_sync(0xffffffffu, (unsigned)v, src), hi = __shfl_sync(0xffffffffu, (unsigned)(v >> 32), src);
    return ((u64)hi << 32) | lo;
}

// Run `step` over in[s0, s1) for this lane; outputs are written to out[sb, s1) (s0 <= sb <= s1: [s0, sb) is a warm-up
// whose outputs are discarded).  `in` / `out` are the SAME for all 32 lanes (the workspace arrays); s0, sb, s1 index them.
// s0 and sb must be multiples of 4 (16-byte aligned pieces); s1 may be anything — the last piece is then loaded/stored
// rounded up to 4 samples, inside the padding every workspace row has.  A lane with nothing to do passes s0 == s1.
// All 32 lanes of the warp must call this together; s1 - s0 < 2^32.
//   Step:  void quad(const float4 &v, float4 &o);   float one(float v);
// STORE = false: a pure warm-up (sb == s1), no output row is written at all.
// HOOK = true: after every COMPLETE 64-sample round c of this lane, step.round_done(c) is called (checkpoints of the loop
// state at fixed positions of the tile, see k_pll_core) — no extra pass, no pipeline drain.
template <bool STORE, class Step, bool HOOK = false>
__device__ __forceinline__ void lane_stream(LaneStream &h, const int lane, const float *__restrict__ in, float *__restrict__ out,
                                            const u64 s0, const u64 sb, const u64 s1, Step &step)
{
    LaneStreamSmem &sm = *h.sm;
    const unsigned len = (unsigned)(s1 - s0);
    const unsigned nch = (len + LS_R - 1) / LS_R;
    unsigned maxch = nch;
#pragma unroll
    for (int o = 16; o; o >>= 1) { const unsigned t = __shfl_xor_sync(0xffffffffu, maxch, o); maxch = t > maxch ? t : maxch; }
    if (maxch == 0) return;
    // this lane's part in the cooperative row copies: 16 lanes cover one row, request i of a round covers rows 2i and 2i+1.
    // The geometry of "its" 16 rows is fetched once per call (registers), so a round's 16 requests are pure arithmetic.
    const int half = lane >> 4, piece = (lane & 15) * 4;
    u64 r_off[16]; unsigned r_n[16], r_sf[16];
    const unsigned my_sf = (unsigned)(sb - s0);
#pragma unroll
    for (int i = 0; i < 16; i++) {
        r_off[i] = ls_shfl64(s0, 2 * i + half) + (u64)piece;
        r_n[i] = __shfl_sync(0xffffffffu, len, 2 * i + half);
        r_sf[i] = __shfl_sync(0xffffffffu, my_sf, 2 * i + half);
    }
    const uint32_t in_u32 = ls_smem(&sm.in[0][0][0]) + 4u * (unsigned)(half * LS_RS + piece);
    __syncwarp();

    auto issue = [&](unsigned c) {
        if (c < maxch) {
            const unsigned off = c * LS_R + (unsigned)piece;
            const uint32_t dst = in_u32 + 4u * (unsigned)((c % LS_NIN) * 32 * LS_RS);
#pragma unroll
            for (int i = 0; i < 16; i++)
                if (off < r_n[i]) ls_cp16(dst + 4u * (unsigned)(2 * i * LS_RS), in + r_off[i] + (u64)c * LS_R);
        }
        ls_commit();
    };
#pragma unroll
    for (int c = 0; c < LS_NIN - 1; c++) issue((unsigned)c);

    for (unsigned c = 0; c < maxch; c++) {
        const int st = (int)(c % LS_NIN);
        ls_wait<LS_NIN - 2>();                                    // this lane's pieces of round c have landed ...
        __syncwarp();                                             // ... and so have everybody else's; round c-1 is fully consumed
        issue(c + LS_NIN - 1);
        if (c < nch) {
            const unsigned g = c * LS_R;
            const unsigned cnt = (len - g < (unsigned)LS_R) ? (len - g) : (unsigned)LS_R;
            const float *__restrict__ ir = sm.in[st][lane];
            float *__restrict__ orow = sm.out[lane];
            unsigned j = 0;
            if (cnt == (unsigned)LS_R) {
                float4 v = ld4(ir);                               // the next quad is fetched before the current one is consumed:
#pragma unroll
                for (; j < (unsigned)LS_R; j += 4) {              // the recurrence never waits for shared memory
                    float4 vn = v;
                    if (j + 4 < (unsigned)LS_R) vn = ld4(ir + j + 4);
                    float4 o; step.quad(v, o);
                    if (STORE) st4(orow + j, o);
                    v = vn;
                }
                if constexpr (HOOK) step.round_done(c);
            } else {
                for (; j + 4 <= cnt; j += 4) { const float4 v = ld4(ir + j); float4 o; step.quad(v, o); if (STORE) st4(orow + j, o); }
                for (; j < cnt; j++) { const float o = step.one(ir[j]); if (STORE) orow[j] = o; }
                if (STORE) for (; j & 3u; j++) orow[j] = 0.0f;
            }
        }
        if (STORE) {
            __syncwarp();
            const unsigned off = c * LS_R + (unsigned)piece;
#pragma unroll
            for (int i = 0; i < 16; i++)
                if (off < r_n[i] && off >= r_sf[i])
                    *reinterpret_cast<float4 *>(out + r_off[i] + (u64)c * LS_R) = ld4(&sm.out[2 * i + half][piece]);
        }
    }
    ls_wait<0>();
    __syncwarp();
}

#endif // __CUDACC__

} // namespace tiled
} // namespace pdt
