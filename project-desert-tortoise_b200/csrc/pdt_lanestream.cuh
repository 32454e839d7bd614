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
    unsigned long long src[32], dst[32];     // per lane: global byte address of its stream's first sample (input / output)
    unsigned n[32], sf[32];                  // per lane: samples in the stream, first sample whose output is kept
};

struct LaneStream { LaneStreamSmem *sm; };

__device__ __forceinline__ uint32_t ls_smem(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ls_cp16(void *dst_smem, unsigned long long src_gmem)
{
    // no "memory" clobber: the copy is ordered against its consumers by ls_wait + __syncwarp (both compiler barriers), and
    // without it the row metadata loads of the 16 requests of a round can be scheduled ahead of the requests
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ls_smem(dst_smem)), "l"(src_gmem));
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

// Run `step` over in[s0, s1) for this lane; outputs are written to out[sb, s1) (s0 <= sb <= s1: [s0, sb) is a warm-up
// whose outputs are discarded).  s0 and sb must be multiples of 4 (16-byte aligned pieces); s1 may be anything — the
// last piece is then loaded/stored rounded up to 4 samples, inside the padding every workspace row has.
// A lane with nothing to do passes s0 == s1.  All 32 lanes of the warp must call this together; s1 - s0 < 2^32.
//   Step:  void quad(const float4 &v, float4 &o);   float one(float v);
// STORE = false: a pure warm-up (sb == s1), no output row is written at all.
template <bool STORE, class Step>
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
    __syncwarp();
    sm.src[lane] = (unsigned long long)(in + s0); sm.dst[lane] = (unsigned long long)(out + s0);
    sm.n[lane] = len; sm.sf[lane] = (unsigned)(sb - s0);
    __syncwarp();
    const int half = lane >> 4, piece = (lane & 15) * 4;       // this lane's part in the cooperative row copies

    auto issue = [&](unsigned c) {
        if (c < maxch) {
            const int st = (int)(c % LS_NIN);
            const unsigned off = c * LS_R + (unsigned)piece;
            unsigned nr[16]; unsigned long long sr[16];
#pragma unroll
            for (int i = 0; i < 16; i++) { nr[i] = sm.n[2 * i + half]; sr[i] = sm.src[2 * i + half]; }
#pragma unroll
            for (int i = 0; i < 16; i++)
                if (off < nr[i]) ls_cp16(&sm.in[st][2 * i + half][piece], sr[i] + 4ull * off);
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
            } else {
                for (; j + 4 <= cnt; j += 4) { const float4 v = ld4(ir + j); float4 o; step.quad(v, o); if (STORE) st4(orow + j, o); }
                for (; j < cnt; j++) { const float o = step.one(ir[j]); if (STORE) orow[j] = o; }
                if (STORE) for (; j & 3u; j++) orow[j] = 0.0f;
            }
        }
        if (STORE) {
            __syncwarp();
            const unsigned off = c * LS_R + (unsigned)piece;
            unsigned nr[16], fr[16];
#pragma unroll
            for (int i = 0; i < 16; i++) { nr[i] = sm.n[2 * i + half]; fr[i] = sm.sf[2 * i + half]; }
#pragma unroll
            for (int i = 0; i < 16; i++)
                if (off < nr[i] && off >= fr[i])
                    *reinterpret_cast<float4 *>(sm.dst[2 * i + half] + 4ull * off) = ld4(&sm.out[2 * i + half][piece]);
        }
    }
    ls_wait<0>();
    __syncwarp();
}

#endif // __CUDACC__

} // namespace tiled
} // namespace pdt
