// pdt_tiled_kernels.cuh — sm_100a kernels of the tiled engine (see pdt_tiled.cuh for the arithmetic and the design).
//
// Launch order per batch (one stream, no host synchronisation in between):
//   k_norm      StaticGain of the first chunk                                  warp per capture
//   k_sp        sample_phase = approx-atan2(Q, I) for every sample             data-parallel, float4
//   k_acquire   PLL acquisition (sweep + lock detector) up to the lock latch   CTA per capture, block-speculative
//   k_estimate  carrier frequency/phase guess per tile (decimate + 1024-FFT)   warp per tile
//   k_pll_core  track-mode phase/frequency recurrence, warm-up + main           lane per tile
//   k_pll_fix   accept tiles whose warm-up state is bit-identical, else re-run  thread per capture
//   k_front<L>  NCO sincos + derotation + ×L interpolating FIR (exact order)    CTA per 3328-sample span
//   k_agc_core  AGC gain recurrence, warm-up + main                             lane per tile
//   k_agc_fix   same verification for the AGC                                   thread per capture
//   k_gardner   Gardner timing recovery -> symbol stream                        warp per capture
//   k_bits      Manchester -> ByteSync -> frame table                           lane per capture
#pragma once

#include "pdt_tiled.cuh"
#include "pdt_lanestream.cuh"

namespace pdt {
namespace tiled {

struct GarRecord { u64 n_sym; float final_next; float pad; };
struct LaneTask { uint32_t cap, k0; };      // one warp's work in a lane-stream kernel: tiles k0 … k0+31 of capture cap

struct TiledArgs {
    ChainConst  cc;
    const void *iq; int pcm16; u64 stride; const u64 *n_samples; u64 n_uniform; uint32_t n_captures;
    u64         ws_stride;   // samples between captures in the workspaces below (multiple of 4: float4 alignment)
    // workspaces (device)
    float      *sp;          // [captures][stride]      sample_phase
    float      *ph;          // [captures][stride]      PLL phase used to derotate each sample
    float      *y;           // [captures][stride·L]    FIR output
    float      *z;           // [captures][stride·L]    AGC output
    AcqResult  *acq;         // [captures]
    LoopState2 *guess, *pll_start, *pll_end;     // [captures][pll.max_tiles]
    LoopState2 *pll_ckpt;    // [captures][pll.max_tiles][pll_nck]  PLL state at tile begin + j·PLL_CK (j >= 1) of the stored run
    unsigned    pll_nck;
    LoopState2 *agc_start, *agc_end;             // [captures][agc_max_tiles]
    uint32_t   *counters;    // [0] PLL tiles re-run, [1] AGC tiles re-run, [2] acquisition flag restarts
    TilePlan    pll;         // in input samples
    u64         agc_min_tile; unsigned agc_max_tiles;   // in interpolated samples
    int         est_decim;   // D
    float       est_fmax;    // peak search limit (Hz)
    u64         acq_first;   // samples covered by the first acquisition pass (0 = whole capture in one pass)
    int         slow_pass;   // 0: kernels process the captures that latched in the first pass, 1: the slow ones
    uint32_t    prelock_from;// captures >= this index skip the acquisition sweep and start in track mode from a carrier
                             // estimate (k_prelock): segments of one stream behind the first (pdt_demod_segments_device)
    uint32_t   *slow_list, *slow_count;  // captures the second acquisition pass continues (k_slow_list -> k_acquire_packed)
    const uint32_t *cap_list, *cap_list_count;   // k_front1<., true>: compact launch over this capture list (the slow captures) or nullptr
    LaneTask   *pll_tasks, *agc_tasks;   // compact work lists of the persistent lane-stream kernels (built on the device)
    uint32_t   *task_counts; // [0] PLL tasks, [1] AGC tasks
    unsigned    pll_tasks_per_cap, agc_tasks_per_cap;
    unsigned    agc_tile_halves;         // AGC tile length in halves of the warm-up length (1: T = W/2)
    unsigned    front_tiles;             // k_front1 tiles per capture of this launch (persistent mode)
    float      *sym;         // [captures][sym_cap]     Gardner symbol stream
    u64        *gidx;        // [captures][sym_cap]     absolute interpolated-sample index of every pick
    GarRecord  *gar;         // [captures]
    u64         sym_cap;
    pdt_capture_stats *stats; pdt_frame *frames; const pdt_traces *traces;
};

PDT_DEV u64 cap_len(const TiledArgs &a, uint32_t c) { return a.n_samples ? a.n_samples[c] : a.n_uniform; }
// does this launch (fast pipeline / slow-capture pipeline) own capture c?
PDT_DEV bool cap_selected(const TiledArgs &a, uint32_t c) { return (a.acq[c].slow != 0) == (a.slow_pass != 0); }

PDT_DEV TrackConst track_const(const TiledArgs &a, const AcqResult &acq)
{
    TrackConst k;
    k.alpha = acq.alpha; k.beta = acq.beta;
    k.max_freq = 2.0 * PDT_PI * a.cc.pll.freq_range / a.cc.pll.Fs;       // CarrierTrackingPLL.c:93-94
    k.min_freq = -2.0 * PDT_PI * a.cc.pll.freq_range / a.cc.pll.Fs;
    return k;
}

// AGC tile plan of one capture, in interpolated samples.  The AGC loop contracts with time constant gain/decay
// samples (AGC.c:120: gain -= (|x·gain| - 1)·rate); ~16 of them bring a 1/mean|y| guess down to a bit-identical
// gain (measured), 22 are used.  The gain the loop will settle at is 1/mean|y|, measured by k_agc_plan
// on the FIR output; an estimate that is off only costs a re-run in k_agc_fix.
PDT_DEV TilePlan agc_plan(const TiledArgs &a, const AcqResult &acq)
{
    TilePlan p;
    float g = acq.agc_gain_est; if (!(g > 0.25f)) g = 0.25f;
    double w = 22.0 * (double)g / (double)a.cc.agc_decay;
    if (w < (double)a.agc_min_tile) w = (double)a.agc_min_tile;
    if (w > 1e15) w = 1e15;
    p.W = ((u64)w + 3) & ~3ull;
    p.T = (p.W * a.agc_tile_halves / 2 + 3) & ~3ull; p.T0 = p.W + p.T; p.max_tiles = a.agc_max_tiles;
    return p;
}

// tiles a capture has under a plan (tile_range() true for k = 0 … count-1)
PDT_DEV unsigned tile_count(u64 first, u64 n, const TilePlan &p)
{
    if (first >= n) return 0;
    const u64 a0 = (first + 3) & ~3ull;
    u64 cnt = 1;
    if (n > a0 + p.T0) cnt += 1 + (n - 1 - a0 - p.T0) / p.T;
    return (unsigned)(cnt < p.max_tiles ? cnt : p.max_tiles);
}

// append ceil(tiles/32) warp tasks for capture `cap` to a work list (order is irrelevant)
__device__ __forceinline__ void push_tasks(LaneTask *list, uint32_t *count, uint32_t cap, unsigned tiles)
{
    const unsigned nt = (tiles + 31) / 32;
    if (nt == 0) return;
    const uint32_t at = atomicAdd(count, nt);
    for (unsigned i = 0; i < nt; i++) list[at + i] = LaneTask{cap, 32u * i};
}

__global__ void __launch_bounds__(128) k_agc_plan(const TiledArgs a)
{
    const uint32_t cap = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (cap >= a.n_captures || !cap_selected(a, cap)) return;
    const u64 nL = cap_len(a, cap) * a.cc.L;
    const float *y = a.y + (u64)cap * a.ws_stride * a.cc.L;
    // 32 evenly spread segments of 256 samples (lane = segment); the largest segment mean bounds the smallest gain,
    // the smallest mean the largest gain: use the median-ish robust choice = overall mean
    float sum = 0.0f; unsigned cnt = 0;
    if (nL > 0) {
        const u64 seg = nL / 32;
        const u64 b0 = (u64)lane * seg;
        const u64 len = seg < 256 ? seg : 256;
        for (u64 i = 0; i < len; i++) { sum += fabsf(y[b0 + i]); cnt++; }
    }
    for (int o = 16; o; o >>= 1) { sum += __shfl_xor_sync(0xffffffffu, sum, o); cnt += __shfl_xor_sync(0xffffffffu, cnt, o); }
    if (lane == 0) {
        AcqResult &acq = a.acq[cap];
        acq.agc_gain_est = (sum > 0.0f) ? (float)cnt / sum : acq.norm;
        push_tasks(a.agc_tasks, &a.task_counts[1], cap, tile_count(0, nL, agc_plan(a, acq)));
    }
}

// work list of the PLL track kernels: one thread per capture
__global__ void __launch_bounds__(128) k_pll_tasks(const TiledArgs a)
{
    const uint32_t cap = blockIdx.x * blockDim.x + threadIdx.x;
    if (cap >= a.n_captures || !cap_selected(a, cap)) return;
    const AcqResult &acq = a.acq[cap];
    if (!acq.locked) return;
    push_tasks(a.pll_tasks, &a.task_counts[0], cap, tile_count(acq.track_begin, cap_len(a, cap), a.pll));
}

// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_norm(const TiledArgs a)
{
    const uint32_t cap = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (cap >= a.n_captures) return;
    AcqResult *r = &a.acq[cap];
    if (lane == 0) { r->prelocked = 0; r->prelock_snr = 0.0f; }
    if (a.cc.norm_override != 0) { if (lane == 0) r->norm = a.cc.norm_override; return; }
    const u64 n = cap_len(a, cap), first = (u64)cap * a.stride;
    const u64 m = n < a.cc.chunk ? n : a.cc.chunk;
    if (m == 0) { if (lane == 0) r->norm = 1.0f; return; }
    float x0, y0;
    load_iq1(a.iq, a.pcm16, first, x0, y0);
    float level = hypot_exact(x0, y0);                                   // AGC.c:58
    for (u64 base = 0; base < m; base += 32) {
        const u64 i = base + lane;
        float h = 0.0f;
        if (i < m) { float p, q; load_iq1(a.iq, a.pcm16, first + i, p, q); h = hypot_exact(p, q); }
        const int cnt = (int)((m - base < 32) ? (m - base) : 32);
        for (int j = 0; j < cnt; j++) {                                  // AGC.c:62-71, strictly serial
            const float hj = __shfl_sync(0xffffffffu, h, j);
            level += hj;
            level /= 2.0;
        }
    }
    if (lane == 0) r->norm = (float)1.0 / level;
}

// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_sp(const TiledArgs a)
{
    const uint32_t cap = blockIdx.y;
    const u64 n = cap_len(a, cap), first = (u64)cap * a.stride;
    float *sp = a.sp + (u64)cap * a.ws_stride;
    const bool vec = !a.pcm16 && !(first & 1) && !(reinterpret_cast<uintptr_t>(a.iq) & 15);
    for (u64 i = ((u64)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += (u64)gridDim.x * blockDim.x * 4) {
        if (i + 4 <= n && vec) {
            const float4 v0 = ld4(reinterpret_cast<const float *>(a.iq) + 2 * (first + i));
            const float4 v1 = ld4(reinterpret_cast<const float *>(a.iq) + 2 * (first + i) + 4);
            float4 o;
            o.x = arctan2_approx(v0.y, v0.x); o.y = arctan2_approx(v0.w, v0.z);     // CarrierTrackingPLL.c:128
            o.z = arctan2_approx(v1.y, v1.x); o.w = arctan2_approx(v1.w, v1.z);
            st4(sp + i, o);
        } else {
            for (u64 j = i; j < n && j < i + 4; j++) {
                float p, q;
                load_iq1(a.iq, a.pcm16, first + j, p, q);
                sp[j] = arctan2_approx(q, p);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Acquisition.  Per sample the reference does (CarrierTrackingPLL.c:102-275)
//   [A] derotate with the current phase, EMA of |output phase|  -> avg_phase            (feeds [C])
//   [B] phase/frequency update from the raw sample phase        (11 dependent flops)
//   [C] while avg_phase looks like noise: frequency += sweep     (feeds the next [B])
//   [D] EMA lock detector; first crossing of the threshold latches the loop into track mode.
// The feedback [A]->[C] is ~300 cycles long but its outcome (a boolean) changes a handful of times per
// capture, so a block of ACQ_B samples is run with the boolean SPECULATED constant: [B]+[C] serially on one
// lane, [A]'s transcendental part on all threads, the two EMAs serially on two lanes of different warps, then
// the booleans are compared; on the first difference the block restarts from that sample with the corrected
// flag.  The committed trajectory is exactly the serial one.
// ---------------------------------------------------------------------------------------------------
// Geometry (r02i/r02l measurements, profiles/README.md): 128-sample blocks and 128-thread CTAs in both passes.  The step's
// fixed cost weighs a little more than with 256-sample blocks, but a CTA holds 21 KB of shared memory and 7 K registers
// instead of 44 KB / 12.5 K — acquisition CTAs sit on their SM for milliseconds (the slow ones for 60 ms), and what they
// hold is what the other batches in flight cannot use.
#ifndef PDT_ACQ_B
#define PDT_ACQ_B 128
#endif
#ifndef PDT_ACQ_THREADS
#define PDT_ACQ_THREADS 128
#endif
constexpr int ACQ_B = PDT_ACQ_B;      // samples per pipeline block
constexpr int ACQ_RING = 4;           // blocks in flight: core | terms | EMAs | decisions
constexpr int ACQ_SERIAL = 64;        // warp 0: the core lane; warp 1: the two EMA lanes (same code, own data)
constexpr int ACQ_THREADS = PDT_ACQ_THREADS;   // 2 serial warps + the helper warps (first pass)
constexpr int ACQ_THREADS_SLOW = 128; // second pass: the few slow captures hold their CTA for tens of ms — a thinner CTA (two helper warps; four warps so
                                      // that the rotated roles reach all four sub-cores) pins fewer registers while several batches are in flight

struct __align__(16) AcqSmem {
    float sp[ACQ_RING][ACQ_B], a[ACQ_RING][ACQ_B], b[ACQ_RING][ACQ_B];               // inputs
    float ph[ACQ_RING][ACQ_B + 4], fr[ACQ_RING][ACQ_B + 4], sw[ACQ_RING][ACQ_B + 4]; // state BEFORE each sample ([cnt] = after the block)
    float aterm[ACQ_RING][ACQ_B], lterm[ACQ_RING][ACQ_B];
    float avg[ACQ_RING][ACQ_B + 4], lks[ACQ_RING][ACQ_B + 4];                        // [0] = before the block, [i+1] = after sample i
    int mism[2], latch[2];                                                           // by step parity
    unsigned char spec[ACQ_B];                                                       // per-sample flag prediction for block 0 of an epoch
};

// [B]+[C] for one sample with the sweep switched on (CarrierTrackingPLL.c:232-246 taken) or off
template <bool ON>
PDT_DEV void acq_step(float &phase, float &freq, float &sweep, float sp, const TrackConst &k)
{
    pll_track_step(phase, freq, sp, k);
    if (ON) {
        const float f2 = freq + sweep;
        float s2 = (f2 >= 0) ? fabsf(sweep) : -fabsf(sweep);
        s2 = (f2 <= k.min_freq) ? -sweep : s2;
        s2 = (f2 >= k.max_freq) ? -sweep : s2;
        freq = f2; sweep = s2;
    }
}

// serial core over one block (cnt samples) with ONE flag value for the whole block (a flag flip ends the pipeline epoch).
// Inputs are fetched a quad ahead and the three state streams leave as float4: one dependent chain that never waits
// on shared memory.  ph/fr/sw[i] = state before sample i, [cnt] = state after the block.
template <bool ON>
PDT_DEV void acq_core(float *__restrict__ ph, float *__restrict__ fr, float *__restrict__ sw, const float *__restrict__ sp, int cnt,
                      float phase, float freq, float sweep, const TrackConst &k)
{
    int i = 0;
    if (cnt >= 4) {
        float4 c = ld4(sp);
        for (; i + 4 <= cnt; i += 4) {
            float4 cn = c;
            if (i + 8 <= cnt) cn = ld4(sp + i + 4);
            float4 p, q, w;
            p.x = phase; q.x = freq; w.x = sweep; acq_step<ON>(phase, freq, sweep, c.x, k);
            p.y = phase; q.y = freq; w.y = sweep; acq_step<ON>(phase, freq, sweep, c.y, k);
            p.z = phase; q.z = freq; w.z = sweep; acq_step<ON>(phase, freq, sweep, c.z, k);
            p.w = phase; q.w = freq; w.w = sweep; acq_step<ON>(phase, freq, sweep, c.w, k);
            st4(ph + i, p); st4(fr + i, q); st4(sw + i, w);
            c = cn;
        }
    }
    for (; i < cnt; i++) { ph[i] = phase; fr[i] = freq; sw[i] = sweep; acq_step<ON>(phase, freq, sweep, sp[i], k); }
    ph[cnt] = phase; fr[cnt] = freq; sw[cnt] = sweep;
}

// block 0 of an epoch that starts at a flag flip: per-sample flags (the prediction in spec[0, spec_n), `tail` behind it)
PDT_DEV void acq_core_spec(float *__restrict__ ph, float *__restrict__ fr, float *__restrict__ sw, const float *__restrict__ sp,
                           const unsigned char *__restrict__ spec, int spec_n, bool tail, int cnt, float phase, float freq, float sweep,
                           const TrackConst &k)
{
    for (int i = 0; i < cnt; i++) {
        ph[i] = phase; fr[i] = freq; sw[i] = sweep;
        const bool on = (i < spec_n) ? (spec[i] != 0) : tail;
        if (on) acq_step<true>(phase, freq, sweep, sp[i], k); else acq_step<false>(phase, freq, sweep, sp[i], k);
    }
    ph[cnt] = phase; fr[cnt] = freq; sw[cnt] = sweep;
}

// x <- (float)((double)x·c + (double)t[i]) over one block: the EMA of CarrierTrackingPLL.c:124 / :220 as one dependent chain
PDT_DEV void acq_ema(float *__restrict__ out, const float *__restrict__ term, int cnt, float x, double c)
{
    out[0] = x;
    int i = 0;
    if (cnt >= 4) {
        float4 t = ld4(term);
        for (; i + 4 <= cnt; i += 4) {
            float4 tn = t;
            if (i + 8 <= cnt) tn = ld4(term + i + 4);
            const double t0 = (double)t.x, t1 = (double)t.y, t2 = (double)t.z, t3 = (double)t.w;
            x = (float)((double)x * c + t0); out[i + 1] = x;
            x = (float)((double)x * c + t1); out[i + 2] = x;
            x = (float)((double)x * c + t2); out[i + 3] = x;
            x = (float)((double)x * c + t3); out[i + 4] = x;
            t = tn;
        }
    }
    for (; i < cnt; i++) { x = (float)((double)x * c + (double)term[i]); out[i + 1] = x; }
}

// CarrierTrackingPLL.c:232 — |π/2 - averagePhase| < 0.05 (float fabs of a double difference, compared in double)
PDT_DEV unsigned char acq_noise_like(float avg) { return (double)fabsf((float)(PDT_PI / 2.0 - (double)avg)) < 0.05; }

// Acquisition as a four-stage software pipeline over ACQ_B-sample blocks, one CTA per capture.  In step s
//     thread 0            runs the serial core of block s        (sweep flag SPECULATED constant = the epoch's flag)
//     warps 3..7          compute the feed-forward terms of block s-1 ([A] derotate + |phase|, [D] lock-detector input),
//                         take the decisions of block s-3, stage the inputs of block s+1
//     threads 32 and 64   run the two EMA chains of block s-2
// with two barriers per step, so a step costs what the core chain costs (63-100 cycles per sample, tools/microbench.cu)
// and nothing else.  The decisions of a block are the per-sample sweep flag (compared with the speculation) and the lock
// latch.  A wrong flag at sample m ends the epoch: samples before m are committed, and the pipeline restarts at m from
// the exact state it had there with the flag flipped — the committed trajectory is exactly the serial one.
//
// pass 0: samples [0, min(n, acq_first)) of every capture.  A capture whose loop has not latched by then is marked
//         `slow` and its loop state is saved;
// pass 1: the slow captures only, from where pass 0 stopped to the end of the capture.
// (The two passes let the host run the rest of the chain for the quickly-locking majority while the few captures
// that lock late, or never, are still in their serial acquisition.)
// cycle accounting of the acquisition pipeline (pass 1 only; read with pdt_debug_acq_prof): [0] steps, [1] step cycles,
// [2] core busy, [3] EMA busy, [4] helper busy in the work phase, [5] decision phase (helper 0), [6] epochs
__device__ unsigned long long g_acq_prof[8];

__global__ void __launch_bounds__(ACQ_THREADS) k_acquire(const TiledArgs a, const int pass)
{
    unsigned long long pf_steps = 0, pf_cyc = 0, pf_busy = 0, pf_dec = 0, pf_epochs = 0;
    __shared__ AcqSmem s;
    const int n_threads = (int)blockDim.x, n_helpers = n_threads - ACQ_SERIAL;     // 2 serial warps + the helper warps
    const uint32_t cap = blockIdx.x;
    // Roles are assigned by a VIRTUAL thread index that rotates the warps from CTA to CTA.  A warp's scheduler (SM sub-core) is
    // its warp index modulo 4: with fixed roles the core lane of every CTA on an SM — the one warp that decides how long a
    // step takes, issuing one instruction every ~4 cycles — sits on sub-core 0, and four or five co-resident CTAs (several
    // batches in flight) ask that one scheduler for more than one instruction per cycle while the other three idle.
    const int n_warps_cta = (int)blockDim.x >> 5;
    const int tid = ((((int)threadIdx.x >> 5) + n_warps_cta - (int)(blockIdx.x % (unsigned)n_warps_cta)) % n_warps_cta) * 32 + ((int)threadIdx.x & 31);
    const u64 n = cap_len(a, cap), first = (u64)cap * a.stride;
    const PllParams &pp = a.cc.pll;
    AcqResult *res = &a.acq[cap];
    const u64 wfirst = (u64)cap * a.ws_stride;
    float *ph_out = a.ph + wfirst;
    if (cap >= a.prelock_from && res->prelocked == 1) return;       // started in track mode by k_prelock (launched before this kernel)

    PllState ps;
    pll_reset(ps);
    pll_begin(ps, pp);
    TrackConst kacq; kacq.alpha = ps.alpha; kacq.beta = ps.beta; kacq.max_freq = ps.max_freq; kacq.min_freq = ps.min_freq;
    const float avg_alpha = 0.00005f;
    const double c_avg = 1.0 - avg_alpha, c_lks = 1.0 - pp.lock_alpha;
    uint32_t restarts = 0;

    u64 i_begin = 0, i_stop = n;
    if (pass == 0) {
        if (a.acq_first && a.acq_first < n) i_stop = a.acq_first;
    } else {
        if (!res->slow || res->resume_at >= n) return;
        i_begin = res->resume_at;
        ps.phase = res->phase; ps.freq = res->freq; ps.sweep = res->sweep; ps.avg_phase = res->avg_phase; ps.locksig = res->locksig;
    }
    if (n == 0) {
        if (tid == 0) {
            res->locked = 0; res->slow = 0; res->resume_at = 0; res->lock_sample = 0; res->track_begin = 0;
            res->phase = ps.phase; res->freq = ps.freq;
            res->sweep = ps.sweep; res->avg_phase = ps.avg_phase; res->locksig = ps.locksig; res->lock_freq_hz = 0;
            res->alpha = kacq.alpha; res->beta = kacq.beta;
        }
        return;
    }

    // epoch state (uniform across the CTA): origin, loop state at the origin, speculated flag
    u64 x0 = i_begin;
    float e_phase = ps.phase, e_freq = ps.freq, e_sweep = ps.sweep, e_avg = ps.avg_phase, e_lks = ps.locksig;
    bool flag = acq_noise_like(e_avg) != 0;                  // true for the initial avg_phase = π/2
    int spec_n = 0;                                          // samples of the epoch's block 0 that carry a per-sample prediction
    const int hid = tid - ACQ_SERIAL;                        // helper index (warps 2..), < 0 for the two serial warps

    // (A ramp of small blocks after a flag flip was tried and lost: a step has ~4k cycles of fixed cost, profiles/README.md r01f.)
    auto block_off = [&](long long j) -> u64 { return (u64)j * (u64)ACQ_B; };
    auto block_cnt = [&](long long j) -> int {               // samples in block j of the current epoch (0 = does not exist)
        if (j < 0) return 0;
        const u64 b0 = x0 + block_off(j);
        if (b0 >= i_stop) return 0;
        return (int)((i_stop - b0 < (u64)ACQ_B) ? (i_stop - b0) : (u64)ACQ_B);
    };
    auto load_inputs = [&](long long j) {                    // helpers only
        const int cnt = block_cnt(j);
        const int slot = (int)(j % ACQ_RING);
        const u64 b0 = x0 + block_off(j);
        for (int i = hid; i < cnt; i += n_helpers) {
            float p, q;
            load_iq1(a.iq, a.pcm16, first + b0 + i, p, q);
            s.a[slot][i] = p; s.b[slot][i] = q; s.sp[slot][i] = a.sp[wfirst + b0 + i];
        }
    };

    for (;;) {                                               // one iteration = one epoch
        pf_epochs++;
        if (hid >= 0) load_inputs(0);
        __syncthreads();
        bool epoch_done = false;
        for (long long st = 0; !epoch_done; st++) {
            const int c_core = block_cnt(st), c_term = block_cnt(st - 1), c_ema = block_cnt(st - 2), c_dec = block_cnt(st - 3);
            if (c_core == 0 && c_term == 0 && c_ema == 0 && c_dec == 0) break;      // pipeline drained: nothing left below i_stop
            const long long t_step = clock64();
            if (tid == 0) {
                if (c_core) {
                    const int slot = (int)(st % ACQ_RING), prev = (int)((st + ACQ_RING - 1) % ACQ_RING);
                    float p = e_phase, f = e_freq, w = e_sweep;
                    if (st > 0) { const int pc = block_cnt(st - 1); p = s.ph[prev][pc]; f = s.fr[prev][pc]; w = s.sw[prev][pc]; }
                    const int so = (int)block_off(st);              // the per-sample prediction covers the first spec_n samples of the epoch
                    if (so < spec_n) acq_core_spec(s.ph[slot], s.fr[slot], s.sw[slot], s.sp[slot], s.spec + so, spec_n - so, flag, c_core, p, f, w, kacq);
                    else if (flag) acq_core<true>(s.ph[slot], s.fr[slot], s.sw[slot], s.sp[slot], c_core, p, f, w, kacq);
                    else           acq_core<false>(s.ph[slot], s.fr[slot], s.sw[slot], s.sp[slot], c_core, p, f, w, kacq);
                }
                s.mism[st & 1] = ACQ_B + 1; s.latch[st & 1] = ACQ_B + 1;
            } else if (tid == 32 || tid == 33) {
                // the two EMA chains run on two lanes of ONE warp: same code, own data — one warp instruction serves both
                if (c_ema) {
                    const int slot = (int)((st - 2) % ACQ_RING), prev = (int)((st - 3 + ACQ_RING) % ACQ_RING);
                    const bool is_avg = tid == 32;
                    float *out = is_avg ? s.avg[slot] : s.lks[slot];
                    const float *term = is_avg ? s.aterm[slot] : s.lterm[slot];
                    const float x0e = (st == 2) ? (is_avg ? e_avg : e_lks) : (is_avg ? s.avg[prev][ACQ_B] : s.lks[prev][ACQ_B]);
                    acq_ema(out, term, c_ema, x0e, is_avg ? c_avg : c_lks);                  // :124 / :220
                }
            } else if (hid >= 0) {
                if (c_term) {                                                        // [A],[D] feed-forward parts of block st-1
                    const int slot = (int)((st - 1) % ACQ_RING);
                    for (int i = hid; i < c_term; i += n_helpers) {
                        float ti, tr;
                        sincos_exact(s.ph[slot][i], ti, tr);                                    // :106-107
                        const float p = s.a[slot][i], q = s.b[slot][i], nti = -ti;
                        const float mre = p * tr - q * nti, mim = p * nti + q * tr;            // :110
                        s.aterm[slot][i] = avg_alpha * fabsf(arctan2_approx(mim, mre));         // :117,:124
                        const float mag2 = p * p + q * q;                                       // :193-220
                        const float inv = q_rsqrt(mag2);
                        const float nre = p * inv, nim = q * inv;
                        s.lterm[slot][i] = pp.lock_alpha * (nre * tr + nim * ti);
                    }
                }
                if (block_cnt(st + 1)) load_inputs(st + 1);
            }
            pf_busy += (unsigned long long)(clock64() - t_step);
            __syncthreads();
            const long long t_dec = clock64();
            // decisions of block st-3 (its EMAs were finished in the previous step)
            if (c_dec && hid >= 0) {
                const int slot = (int)((st - 3) % ACQ_RING);
                for (int i = hid; i < c_dec; i += n_helpers) {
                    const int so = (int)block_off(st - 3) + i;
                    const bool want = (so < spec_n) ? (s.spec[so] != 0) : flag;
                    if ((acq_noise_like(s.avg[slot][i + 1]) != 0) != want) atomicMin(&s.mism[st & 1], i);      // :232
                    if (s.lks[slot][i + 1] > pp.lock_thresh) atomicMin(&s.latch[st & 1], i);                   // :266
                }
            }
            __syncthreads();
            pf_dec += (unsigned long long)(clock64() - t_dec);
            pf_steps++;
            auto pf_flush = [&]() {
                pf_cyc += (unsigned long long)(clock64() - t_step);
                if (pass == 1) {
                    if (tid == 0)  { atomicAdd(&g_acq_prof[0], pf_steps); atomicAdd(&g_acq_prof[1], pf_cyc); atomicAdd(&g_acq_prof[2], pf_busy); atomicAdd(&g_acq_prof[6], pf_epochs); }
                    if (tid == 32) atomicAdd(&g_acq_prof[3], pf_busy);
                    if (tid == ACQ_SERIAL) { atomicAdd(&g_acq_prof[4], pf_busy); atomicAdd(&g_acq_prof[5], pf_dec); }
                }
            };
            if (c_dec) {
                const int slot = (int)((st - 3) % ACQ_RING);
                const u64 b0 = x0 + block_off(st - 3);
                const int mism = s.mism[st & 1] < c_dec ? s.mism[st & 1] : c_dec, latch = s.latch[st & 1] < c_dec ? s.latch[st & 1] : c_dec;
                if (latch < c_dec && latch < mism) {
                    // every flag up to and including the latch sample was right: commit and leave acquisition
                    const int cnt = latch + 1;
                    for (int i = tid; i < cnt; i += n_threads) ph_out[b0 + i] = s.ph[slot][i];
                    if (tid == 0) {
                        const float freq = s.fr[slot][cnt];
                        res->locked = 1; res->lock_sample = b0 + latch; res->track_begin = b0 + cnt; res->resume_at = n;
                        if (pass == 0) res->slow = 0;
                        res->phase = s.ph[slot][cnt]; res->freq = freq; res->sweep = s.sw[slot][cnt];
                        res->avg_phase = s.avg[slot][cnt]; res->locksig = s.lks[slot][cnt];
                        res->lock_freq_hz = freq * pp.Fs / (2.0 * PDT_PI);                                  // :269
                        const float bw = pp.bw_track, damp = ps.damp;                                       // :272-273
                        res->alpha = (4.0 * damp * bw) / (1.0 + 2.0 * damp * bw + bw * bw);
                        res->beta  = (4.0 * bw * bw) / (1.0 + 2.0 * damp * bw + bw * bw);
                        if (restarts) atomicAdd(&a.counters[2], restarts);
                    }
                    pf_flush();
                    return;
                }
                const int good = mism;                                   // samples of the block whose speculated flag was right
                for (int i = tid; i < good; i += n_threads) ph_out[b0 + i] = s.ph[slot][i];
                if (mism < c_dec) {
                    // the flag flips at sample `mism`: new epoch there, from the exact state in front of that sample
                    // The flags the EMA just produced for the rest of this block become the per-sample prediction of the new
                    // epoch's block 0: they come from phases that are about to change, but avg_phase moves by < 2e-4 per
                    // sample whatever the phase is, so a burst of flips around a threshold crossing is predicted in one go.
                    const float np_ = s.ph[slot][mism], nf = s.fr[slot][mism], nw = s.sw[slot][mism];
                    const float na = s.avg[slot][mism], nl_ = s.lks[slot][mism];
                    const int nspec = c_dec - mism;
                    unsigned char mine[4] = {0, 0, 0, 0};              // ceil(ACQ_B / smallest CTA) entries
                    for (int i = tid, q = 0; i < nspec; i += n_threads, q++) mine[q] = acq_noise_like(s.avg[slot][mism + i + 1]);
                    const bool tail = acq_noise_like(s.avg[slot][c_dec]) != 0;
                    __syncthreads();
                    for (int i = tid, q = 0; i < nspec; i += n_threads, q++) s.spec[i] = mine[q];
                    x0 = b0 + (u64)mism; e_phase = np_; e_freq = nf; e_sweep = nw; e_avg = na; e_lks = nl_;
                    spec_n = nspec; flag = tail; restarts++;
                    epoch_done = true;
                } else if (b0 + (u64)c_dec >= i_stop) {
                    // reached the end of this pass without a latch
                    if (tid == 0) {
                        const bool done = (i_stop == n);
                        res->locked = 0; res->lock_sample = 0; res->track_begin = n;
                        res->resume_at = i_stop;
                        if (pass == 0) res->slow = done ? 0 : 1;
                        res->phase = s.ph[slot][c_dec]; res->freq = s.fr[slot][c_dec]; res->sweep = s.sw[slot][c_dec];
                        res->avg_phase = s.avg[slot][c_dec]; res->locksig = s.lks[slot][c_dec];
                        res->lock_freq_hz = 0; res->alpha = kacq.alpha; res->beta = kacq.beta;
                        if (restarts) atomicAdd(&a.counters[2], restarts);
                    }
                    pf_flush();
                    return;
                }
            }
            // the commit above read ring slot (st-3) % 4, which is the slot the core lane fills in the next step: close the
            // step before anybody runs ahead (compute-sanitizer racecheck; ~30 of a step's ~20 000 cycles)
            if (c_dec && !epoch_done) __syncthreads();
            pf_cyc += (unsigned long long)(clock64() - t_step);
        }
        if (!epoch_done) return;        // (not reached: the last block always returns above)
    }
}

// ---------------------------------------------------------------------------------------------------
// k_acquire_packed — the same acquisition (same arithmetic, same speculation, same committed trajectory) with the serial
// chains of 32 captures packed into the lanes of ONE warp each.
//
// k_acquire gives every capture a CTA in which three lanes do serial work (the core recurrence and the two EMAs): a warp
// instruction is issued for one useful lane, 5.4 G warp instructions per 1024 x 1 M batch, and 1024 (first pass) or ~90
// long-lived (second pass, 55 ms) CTAs hold registers and shared memory that the other kernels in flight cannot use —
// together ~15 ms of every 31 ms step (profiles/README.md, r02e-r02p).  Here a CTA serves AQ = 32 captures:
//     warp 0   lane q = core recurrence [B]+[C] of capture q          (sweep flag by select: the lanes stay in lock-step)
//     warp 1   lane q = EMA of |output phase| of capture q   (:124)
//     warp 2   lane q = EMA of the lock detector of capture q (:220)
//     warps 3+ helpers: feed-forward terms (sincos, derotation, approx-atan2, Q_rsqrt), decisions, loads, commits — a warp
//              takes one capture's block at a time (lane = sample), so shared and global accesses are contiguous
// over AB = 32-sample blocks in a three-stage pipeline (core | terms | EMAs + decisions: the EMA lanes leave every sample's
// sweep / latch decision as a bit, the verdict on a block is integer work for one control warp in the same step), every
// capture with its own epoch origin and step counter: a wrong sweep flag restarts THAT capture at the offending sample
// while the others go on.
// Rows of different captures are 129 / 133 floats apart, so the 32 serial lanes hit 32 different banks.
// The first pass needs ceil(captures / 32) CTAs (32 SMs for a 1024-capture batch), the second pass three.
// ---------------------------------------------------------------------------------------------------
constexpr int AQ = 32;                  // captures per CTA
constexpr int AB = 32;                  // samples per block
constexpr int AR = 4;                   // ring slots (blocks in flight per capture)
constexpr int AP_SERIAL = 96;           // three serial warps
constexpr int AP_THREADS = 512;         // + thirteen helper warps (a helper warp serves at most AP_HCAPS captures per step)
constexpr int AP_HCAPS = (AQ + (AP_THREADS - AP_SERIAL) / 32 - 1) / ((AP_THREADS - AP_SERIAL) / 32);
constexpr int AP_RS = AR * AB + 1;      // row stride of the [slot][sample] arrays (129: one bank further per capture)
constexpr int AP_RS1 = AR * (AB + 1) + 1;   // row stride of the [slot][sample + 1] arrays (133 = 5 mod 32)

struct AcqCap {                         // control block of one capture slot (shared memory)
    u64   x0, i_stop, n, first, wfirst; // epoch origin; end of this pass; capture length; sample offsets of the capture's rows
    int   cap;                          // capture index (-1: empty slot)
    int   st;                           // steps since the epoch started
    int   active;
    int   flag, spec_n;                 // speculated flag behind the per-sample predicted prefix (spec_n <= 2·AB samples of the epoch)
    uint32_t specmask[2];               // predicted flags of the epoch's first 64 samples (bit i of word b = sample 32·b + i)
    uint32_t amask[AR], lmask[AR];      // per ring slot, written by the EMA lanes: bit i = sweep flag the reference takes at sample i
                                        // (:232), bit i = lock detector above its threshold after sample i (:266)
    float e_phase, e_freq, e_sweep, e_avg, e_lks;   // loop state at the epoch origin
    uint32_t restarts;
};

struct AcqPackSmem {
    float sp[AQ][AP_RS], a[AQ][AP_RS], b[AQ][AP_RS], aterm[AQ][AP_RS], lterm[AQ][AP_RS];
    float ph[AQ][AP_RS1], fr[AQ][AP_RS1], sw[AQ][AP_RS1];      // [slot][i] = state BEFORE sample i, [slot][cnt] = after the block
    float avg[AQ][AP_RS1], lks[AQ][AP_RS1];                    // [slot][0] = before the block, [slot][i + 1] = after sample i
    AcqCap c[AQ];
    int   n_active;
    float noise_lo, noise_hi;           // acq_noise_like(avg) <=> noise_lo <= avg <= noise_hi (found by bisection at kernel start)
};

// compact list of the captures the second acquisition pass has to continue (one thread per capture, whole batch)
__global__ void __launch_bounds__(128) k_slow_list(const TiledArgs a, uint32_t *__restrict__ list, uint32_t *__restrict__ count)
{
    const uint32_t cap = blockIdx.x * blockDim.x + threadIdx.x;
    if (cap >= a.n_captures) return;
    const AcqResult &r = a.acq[cap];
    if (r.slow && r.resume_at < cap_len(a, cap)) list[atomicAdd(count, 1u)] = cap;
    if (r.slow) list[a.n_captures + atomicAdd(count + 1, 1u)] = cap;       // second list: every capture of the slow pipeline
}

__device__ __forceinline__ void acq_step_sel(float &phase, float &freq, float &sweep, float sp, const TrackConst &k, bool on)
{
    pll_track_step(phase, freq, sp, k);
    const float f2 = freq + sweep;                                              // CarrierTrackingPLL.c:232-246, taken or not by select
    float s2 = (f2 >= 0) ? fabsf(sweep) : -fabsf(sweep);
    s2 = (f2 <= k.min_freq) ? -sweep : s2;
    s2 = (f2 >= k.max_freq) ? -sweep : s2;
    freq = on ? f2 : freq; sweep = on ? s2 : sweep;
}

__global__ void __launch_bounds__(AP_THREADS, 1) k_acquire_packed(const TiledArgs a, const int pass, const uint32_t *__restrict__ slow_list,
                                                                  const uint32_t *__restrict__ slow_count)
{
    extern __shared__ __align__(16) unsigned char ap_raw[];
    AcqPackSmem &s = *reinterpret_cast<AcqPackSmem *>(ap_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int hid = tid - AP_SERIAL, n_helpers = AP_THREADS - AP_SERIAL, n_hwarps = n_helpers / 32, hwarp = warp - 3;
    const PllParams &pp = a.cc.pll;
    PllState ps0; pll_reset(ps0); pll_begin(ps0, pp);
    TrackConst kacq; kacq.alpha = ps0.alpha; kacq.beta = ps0.beta; kacq.max_freq = ps0.max_freq; kacq.min_freq = ps0.min_freq;
    const float avg_alpha = 0.00005f;
    const double c_avg = 1.0 - avg_alpha, c_lks = 1.0 - pp.lock_alpha;

    // ---- slot set-up: capture q of this CTA ---------------------------------------------------------------------
    if (tid < AQ) {
        AcqCap &c = s.c[tid];
        const u64 idx = (u64)blockIdx.x * AQ + tid;
        long long cap = -1;
        if (pass == 0) { if (idx < a.n_captures) cap = (long long)idx; }
        else if (idx < *slow_count) cap = (long long)slow_list[idx];
        c.cap = (int)cap; c.active = 0; c.st = -1; c.restarts = 0; c.spec_n = 0;   // st = -1: the first step only loads block 0
        if (cap >= 0) {
            AcqResult *res = &a.acq[cap];
            const u64 n = cap_len(a, (uint32_t)cap);
            c.n = n; c.first = (u64)cap * a.stride; c.wfirst = (u64)cap * a.ws_stride;
            bool run = true;
            if (pass == 0) {
                if ((uint32_t)cap >= a.prelock_from && res->prelocked == 1) run = false;      // started in track mode by k_prelock
                c.x0 = 0; c.i_stop = (a.acq_first && a.acq_first < n) ? a.acq_first : n;
                c.e_phase = ps0.phase; c.e_freq = ps0.freq; c.e_sweep = ps0.sweep; c.e_avg = ps0.avg_phase; c.e_lks = ps0.locksig;
                if (run && n == 0) {
                    res->locked = 0; res->slow = 0; res->resume_at = 0; res->lock_sample = 0; res->track_begin = 0;
                    res->phase = ps0.phase; res->freq = ps0.freq; res->sweep = ps0.sweep; res->avg_phase = ps0.avg_phase;
                    res->locksig = ps0.locksig; res->lock_freq_hz = 0; res->alpha = kacq.alpha; res->beta = kacq.beta;
                    run = false;
                }
            } else {
                c.x0 = res->resume_at; c.i_stop = n;
                c.e_phase = res->phase; c.e_freq = res->freq; c.e_sweep = res->sweep; c.e_avg = res->avg_phase; c.e_lks = res->locksig;
                if (!res->slow || res->resume_at >= n) run = false;
            }
            c.flag = acq_noise_like(c.e_avg) != 0;
            c.active = run ? 1 : 0;
        }
    }
    __syncthreads();

    auto blk_cnt = [&](const AcqCap &c, int j) -> int {                 // samples in block j of the capture's current epoch
        if (!c.active || j < 0) return 0;
        const u64 b0 = c.x0 + (u64)j * AB;
        if (b0 >= c.i_stop) return 0;
        return (int)((c.i_stop - b0 < (u64)AB) ? (c.i_stop - b0) : (u64)AB);
    };
    // speculated sweep flags of block j, one bit per sample: the predicted prefix of the epoch, the epoch's flag behind it
    auto blk_flags = [&](const AcqCap &c, int j) -> uint32_t {
        const uint32_t tail = c.flag ? 0xffffffffu : 0u;
        if (j < 0 || j > 1) return tail;
        const int left = c.spec_n - j * AB;                       // predicted samples inside this block
        if (left <= 0) return tail;
        const uint32_t valid = left >= 32 ? 0xffffffffu : ((1u << left) - 1u);
        return (c.specmask[j] & valid) | (tail & ~valid);
    };
    if (tid == 0) { int na = 0; for (int q = 0; q < AQ; q++) na += s.c[q].active; s.n_active = na; }
    if (tid == AP_SERIAL) {
        // CarrierTrackingPLL.c:232 compares |π/2 - averagePhase| with 0.05 through a double difference narrowed to float; the
        // narrowing is monotonic, so the samples that satisfy it are exactly the floats of one interval around π/2.  Its two
        // ends are found once, by bisection over the (ordered) bit patterns of positive floats with the reference's own
        // expression — the 32 EMA lanes then decide with two float compares per sample instead of five double/convert ops.
        const uint32_t mid = pdt_f2u((float)(PDT_PI / 2.0));                 // inside the interval
        uint32_t a0 = 0u, a1 = mid;                                         // lo: smallest pattern that satisfies it
        while (a0 < a1) { const uint32_t m = a0 + (a1 - a0) / 2; if (acq_noise_like(pdt_u2f(m))) a1 = m; else a0 = m + 1; }
        uint32_t b0_ = mid, b1 = 0x7f7fffffu;                               // hi: largest pattern that satisfies it
        while (b0_ < b1) { const uint32_t m = b0_ + (b1 - b0_ + 1) / 2; if (acq_noise_like(pdt_u2f(m))) b0_ = m; else b1 = m - 1; }
        s.noise_lo = pdt_u2f(a0); s.noise_hi = pdt_u2f(b0_);
    }
    __syncthreads();
    int any_active = s.n_active;
    const float noise_lo = s.noise_lo, noise_hi = s.noise_hi;

    unsigned long long pf_steps = 0, pf_a = 0, pf_wait = 0, pf_c = 0;     // cycle accounting (pdt_debug_acq_prof): phase A busy / barrier wait / phase C
    while (any_active) {
        const long long t0 = clock64();
        // ================= phase A: the four pipeline stages, each on its own threads =================
        if (warp == 0) {
            // core recurrence of block st (CarrierTrackingPLL.c:128-188 + sweep :232-246), lane = capture
            const AcqCap &c = s.c[lane];
            const int j = c.st, cnt = blk_cnt(c, j);
            const int slot = j & (AR - 1), prev = (j + AR - 1) & (AR - 1);
            float phase = c.e_phase, freq = c.e_freq, sweep = c.e_sweep;
            if (cnt && j > 0) { const int pc = blk_cnt(c, j - 1); phase = s.ph[lane][prev * (AB + 1) + pc]; freq = s.fr[lane][prev * (AB + 1) + pc]; sweep = s.sw[lane][prev * (AB + 1) + pc]; }
            const uint32_t fm = blk_flags(c, j);
            const float *spr = &s.sp[lane][slot * AB];
            float *php = &s.ph[lane][slot * (AB + 1)], *frp = &s.fr[lane][slot * (AB + 1)], *swp = &s.sw[lane][slot * (AB + 1)];
            const unsigned FULL = 0xffffffffu;
            const bool all_on = __all_sync(FULL, cnt == 0 || fm == 0xffffffffu);
            const bool all_off = __all_sync(FULL, cnt == 0 || fm == 0u);
            // All 32 samples of the slot are run unconditionally, four per trip with their inputs fetched ahead of the
            // dependent chain: entry [i] is the state BEFORE sample i, so for a block of cnt < 32 samples (end of a pass)
            // entry [cnt] is its end state and whatever lies behind it is never read.  Idle lanes chew on stale rows.
            auto run = [&](auto step) {
#pragma unroll 2
                for (int i = 0; i < AB; i += 4) {
                    float x[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) x[u] = spr[i + u];
#pragma unroll
                    for (int u = 0; u < 4; u++) { php[i + u] = phase; frp[i + u] = freq; swp[i + u] = sweep; step(x[u], i + u); }
                }
                php[AB] = phase; frp[AB] = freq; swp[AB] = sweep;
            };
            if (__any_sync(FULL, cnt != 0)) {
                if (all_on)       run([&](float x, int) { acq_step<true>(phase, freq, sweep, x, kacq); });
                else if (all_off) run([&](float x, int) { acq_step<false>(phase, freq, sweep, x, kacq); });
                else              run([&](float x, int i) { acq_step_sel(phase, freq, sweep, x, kacq, ((fm >> i) & 1u) != 0); });
            }
        } else if (warp == 1 || warp == 2) {
            // EMA chains of block st-2: x <- (float)((double)x·c + (double)term) (:124 / :220), lane = capture
            const AcqCap &c = s.c[lane];
            const int j = c.st - 2, cnt = blk_cnt(c, j);
            const int slot = j & (AR - 1), prev = (j + AR - 1) & (AR - 1);
            const bool is_avg = warp == 1;
            float (*row)[AP_RS1] = is_avg ? s.avg : s.lks;
            const float *term = is_avg ? &s.aterm[lane][slot * AB] : &s.lterm[lane][slot * AB];
            const double cc_ = is_avg ? c_avg : c_lks;
            float x = is_avg ? c.e_avg : c.e_lks;
            if (cnt && j > 0) x = row[lane][prev * (AB + 1) + blk_cnt(c, j - 1)];
            float *out = &row[lane][slot * (AB + 1)];
            if (__any_sync(0xffffffffu, cnt != 0)) {
                out[0] = x;
                uint32_t m = 0;                                       // the decision each sample implies, one bit per sample
#pragma unroll 2
                for (int i = 0; i < AB; i += 4) {                     // all 32 entries, like the core: [cnt] is the block's end value
                    double t4[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) t4[u] = (double)term[i + u];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        x = (float)((double)x * cc_ + t4[u]); out[i + u + 1] = x;
                        const bool bit = is_avg ? (x >= noise_lo && x <= noise_hi) : (x > pp.lock_thresh);   // :232 / :266 (off the chain)
                        m |= (bit ? 1u : 0u) << (i + u);
                    }
                }
                if (is_avg) s.c[lane].amask[slot] = m; else s.c[lane].lmask[slot] = m;
            }
        } else {
            // helpers: feed-forward terms of block st-1 ([A] derotate + |phase|, [D] lock-detector input), inputs of block st+1
            // the loads of block st+1 are issued first and land while the terms are computed; they go to shared memory last
            float lp[AP_HCAPS], lq[AP_HCAPS], ls[AP_HCAPS];
#pragma unroll
            for (int u = 0; u < AP_HCAPS; u++) {
                const int q = hwarp + u * n_hwarps;
                lp[u] = lq[u] = ls[u] = 0.0f;
                if (q < AQ) {
                    const AcqCap &c = s.c[q];
                    const int jn = c.st + 1;
                    if (lane < blk_cnt(c, jn)) {
                        const u64 i = c.x0 + (u64)jn * AB + lane;
                        load_iq1(a.iq, a.pcm16, c.first + i, lp[u], lq[u]);
                        ls[u] = a.sp[c.wfirst + i];
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < AP_HCAPS; u++) {
                const int q = hwarp + u * n_hwarps;
                if (q >= AQ) break;
                const AcqCap &c = s.c[q];
                const int j = c.st - 1, cnt = blk_cnt(c, j);
                if (lane < cnt) {
                    const int slot = j & (AR - 1), o = slot * AB + lane;
                    const float phase_i = s.ph[q][slot * (AB + 1) + lane];
                    // the phase stream leaves here, speculatively: a restart at an earlier sample recomputes and overwrites it,
                    // and behind a lock latch the track tiles write their own phases from track_begin on
                    a.ph[c.wfirst + c.x0 + (u64)j * AB + lane] = phase_i;
                    float ti, tr;
                    sincos_exact(phase_i, ti, tr);                                          // :106-107
                    const float p = s.a[q][o], qq = s.b[q][o], nti = -ti;
                    const float mre = p * tr - qq * nti, mim = p * nti + qq * tr;             // :110
                    s.aterm[q][o] = avg_alpha * fabsf(arctan2_approx(mim, mre));              // :117,:124
                    const float mag2 = p * p + qq * qq;                                       // :193-220
                    const float inv = q_rsqrt(mag2);
                    const float nre = p * inv, nim = qq * inv;
                    s.lterm[q][o] = pp.lock_alpha * (nre * tr + nim * ti);
                }
            }
#pragma unroll
            for (int u = 0; u < AP_HCAPS; u++) {
                const int q = hwarp + u * n_hwarps;
                if (q < AQ) {
                    const AcqCap &c = s.c[q];
                    const int jn = c.st + 1;
                    if (lane < blk_cnt(c, jn)) { const int o = (jn & (AR - 1)) * AB + lane; s.a[q][o] = lp[u]; s.b[q][o] = lq[u]; s.sp[q][o] = ls[u]; }
                }
            }
        }
        const long long t1 = clock64();
        __syncthreads();
        const long long t2 = clock64();
        // ================= phase C: decisions of block st-2 and control, ONE warp, lane = capture — the EMA lanes left the
        // per-sample decisions as bit masks, so a capture's verdict is a few integer operations =========
        int my_active = 0;
        if (warp == 3) {
            AcqCap &c = s.c[lane];
            if (c.active) {
                my_active = 1;
                const int j = c.st - 2, cnt = blk_cnt(c, j);
                if (cnt == 0) c.st++;                                                // pipeline still filling for this capture
                else {
                    const int slot = j & (AR - 1);
                    const u64 b0 = c.x0 + (u64)j * AB;
                    const float *php = &s.ph[lane][slot * (AB + 1)], *frp = &s.fr[lane][slot * (AB + 1)], *swp = &s.sw[lane][slot * (AB + 1)];
                    const float *avp = &s.avg[lane][slot * (AB + 1)], *lkp = &s.lks[lane][slot * (AB + 1)];
                    const uint32_t valid = cnt >= 32 ? 0xffffffffu : ((1u << cnt) - 1u);
                    const uint32_t actual = c.amask[slot];
                    const uint32_t diff = (actual ^ blk_flags(c, j)) & valid, lat = c.lmask[slot] & valid;
                    const int mism = diff ? (__ffs((int)diff) - 1) : cnt, latch = lat ? (__ffs((int)lat) - 1) : cnt;
                    AcqResult *res = &a.acq[c.cap];
                    if (latch < cnt && latch < mism) {
                        // every flag up to and including the latch sample was right: leave acquisition
                        const int k = latch + 1;
                        const float freq = frp[k];
                        res->locked = 1; res->lock_sample = b0 + latch; res->track_begin = b0 + k; res->resume_at = c.n;
                        if (pass == 0) res->slow = 0;
                        res->phase = php[k]; res->freq = freq; res->sweep = swp[k];
                        res->avg_phase = avp[k]; res->locksig = lkp[k];
                        res->lock_freq_hz = freq * pp.Fs / (2.0 * PDT_PI);                              // :269
                        const float bw = pp.bw_track, damp = ps0.damp;                                  // :272-273
                        res->alpha = (4.0 * damp * bw) / (1.0 + 2.0 * damp * bw + bw * bw);
                        res->beta  = (4.0 * bw * bw) / (1.0 + 2.0 * damp * bw + bw * bw);
                        if (c.restarts) atomicAdd(&a.counters[2], c.restarts);
                        c.active = 0; my_active = 0;
                    } else if (mism < cnt) {
                        // the flag flips at sample `mism`: new epoch there, from the exact state in front of that sample.  The
                        // decisions the EMA just produced for the rest of this block become the per-sample prediction of the
                        // new epoch (avg_phase moves by < 2e-4 per sample whatever the phase is), the last one its flag.
                        const int n0 = cnt - mism;
                        const float np_ = php[mism], nf = frp[mism], nw = swp[mism], na = avp[mism], nl = lkp[mism];
                        c.x0 = b0 + (u64)mism; c.e_phase = np_; c.e_freq = nf; c.e_sweep = nw; c.e_avg = na; c.e_lks = nl;
                        c.specmask[0] = actual >> mism; c.specmask[1] = 0u;
                        c.spec_n = n0; c.flag = (int)((actual >> (cnt - 1)) & 1u); c.restarts++;
                        c.st = -1;                                                   // next step only loads the new epoch's first block
                    } else if (b0 + (u64)cnt >= c.i_stop) {
                        // reached the end of this pass without a latch
                        const bool done = (c.i_stop == c.n);
                        res->locked = 0; res->lock_sample = 0; res->track_begin = c.n;
                        res->resume_at = c.i_stop;
                        if (pass == 0) res->slow = done ? 0 : 1;
                        res->phase = php[cnt]; res->freq = frp[cnt]; res->sweep = swp[cnt];
                        res->avg_phase = avp[cnt]; res->locksig = lkp[cnt];
                        res->lock_freq_hz = 0; res->alpha = kacq.alpha; res->beta = kacq.beta;
                        if (c.restarts) atomicAdd(&a.counters[2], c.restarts);
                        c.active = 0; my_active = 0;
                    } else c.st++;
                }
            }
        }
        const long long t3 = clock64();
        any_active = __syncthreads_or(my_active);
        pf_steps++; pf_a += (unsigned long long)(t1 - t0); pf_wait += (unsigned long long)(t2 - t1); pf_c += (unsigned long long)(t3 - t2);
    }
    if (pass == 1 && lane == 0 && (warp == 0 || warp == 1 || warp == 3)) {
        // [0] steps, [1] core-warp phase A cycles, [2] its barrier wait, [3] EMA-warp phase A, [4] helper-warp phase A, [5] helper barrier wait, [6] phase C (warp 15), [7] phase C (warp 0)
        if (warp == 0)  { atomicAdd(&g_acq_prof[0], pf_steps); atomicAdd(&g_acq_prof[1], pf_a); atomicAdd(&g_acq_prof[2], pf_wait); atomicAdd(&g_acq_prof[7], pf_c); }
        if (warp == 1)  atomicAdd(&g_acq_prof[3], pf_a);
        if (warp == 3)  { atomicAdd(&g_acq_prof[4], pf_a); atomicAdd(&g_acq_prof[5], pf_wait); }
        if (warp == 3)  atomicAdd(&g_acq_prof[6], pf_c);
    }
}

// ---------------------------------------------------------------------------------------------------
// carrier guess per tile: warp per (capture, tile >= 1)
// ---------------------------------------------------------------------------------------------------
constexpr int EST_WARPS = 2;
constexpr int EST_MAX_D = 24;          // staged (coalesced) decimation up to this factor (6 KB per warp), direct strided reads beyond

// carrier frequency (rad/sample) and phase at sample `warm` of a capture, from the EST_FFT·D samples in front of it.
// One warp; z / stg are its shared-memory scratch.  The result is valid on lane 0.
__device__ __forceinline__ LoopState2 est_carrier(const TiledArgs &a, const u64 first, const u64 warm, float2 *z, float2 *stg, const int lane,
                                                  float *snr_out = nullptr)
{
    const int D = a.est_decim;
    const long long w0 = (long long)warm - (long long)EST_FFT * D;
    // decimate by block sums, store bit-reversed.  The warp reads 32·D consecutive samples at a time (coalesced) into a
    // staging tile and lane j sums its own D of them: read directly, lane j's samples are D·8 bytes apart from lane j+1's
    // and every 32-byte sector fetched from HBM would be used for one 8-byte sample.
    for (int blk = 0; blk < EST_FFT / 32; blk++) {
        const long long b0 = w0 + (long long)blk * 32 * D;
        if (D <= EST_MAX_D) {
            for (int t = lane; t < 32 * D; t += 32) {
                const long long i = b0 + t;
                float p = 0.f, q = 0.f;
                if (i >= 0) load_iq1(a.iq, a.pcm16, first + (u64)i, p, q);
                stg[t] = make_float2(p, q);
            }
            __syncwarp();
            float sr = 0.f, si = 0.f;
            for (int d = 0; d < D; d++) { const float2 v = stg[lane * D + d]; sr += v.x; si += v.y; }
            z[__brev((unsigned)(blk * 32 + lane)) >> 22] = make_float2(sr, si);
            __syncwarp();
        } else {
            float sr = 0.f, si = 0.f;
            for (int d = 0; d < D; d++) {
                const long long i = b0 + (long long)lane * D + d;
                if (i >= 0) { float p, q; load_iq1(a.iq, a.pcm16, first + (u64)i, p, q); sr += p; si += q; }
            }
            z[__brev((unsigned)(blk * 32 + lane)) >> 22] = make_float2(sr, si);
        }
    }
    __syncwarp();
    for (int len = 2; len <= EST_FFT; len <<= 1) {
        const int half = len >> 1;
        for (int b = lane; b < EST_FFT / 2; b += 32) {
            const int grp = b / half, pos = b - grp * half;
            const int i0 = grp * len + pos, i1 = i0 + half;
            float sn, cs;
            est_sincos_turns(-(float)pos / (float)len, sn, cs);
            const float2 u = z[i0], v = z[i1];
            const float tr = v.x * cs - v.y * sn, ti = v.x * sn + v.y * cs;
            z[i0] = make_float2(u.x + tr, u.y + ti);
            z[i1] = make_float2(u.x - tr, u.y - ti);
        }
        __syncwarp();
    }
    const float fs_d = a.cc.pll.Fs / (float)D, bin_hz = fs_d / (float)EST_FFT;
    int kmax = (int)(a.est_fmax / bin_hz) + 1;
    if (kmax > EST_FFT / 2 - 2) kmax = EST_FFT / 2 - 2;
    float best = -1.f, total = 0.f; int best_k = 0;
    for (int kk = -kmax + lane; kk <= kmax; kk += 32) {
        const float2 v = z[kk & (EST_FFT - 1)];
        const float m2 = v.x * v.x + v.y * v.y;
        total += m2;
        if (m2 > best) { best = m2; best_k = kk; }
    }
    for (int o = 16; o; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int ok = __shfl_xor_sync(0xffffffffu, best_k, o);
        total += __shfl_xor_sync(0xffffffffu, total, o);
        if (ob > best || (ob == best && ok < best_k)) { best = ob; best_k = ok; }
    }
    if (snr_out) {       // peak against the mean of the OTHER searched bins (noise only: max of n exponentials ~ ln n ~ 7)
        const float rest = total - best;
        *snr_out = (rest > 0.f) ? best * (float)(2 * kmax) / rest : ((best > 0.f) ? 1e30f : 0.f);
    }
    const float delta = est_peak_offset(z[(best_k - 1) & (EST_FFT - 1)], z[best_k & (EST_FFT - 1)], z[(best_k + 1) & (EST_FFT - 1)]);
    const float f_hz = ((float)best_k + delta) * bin_hz;
    const float cyc = f_hz / a.cc.pll.Fs;                  // cycles per sample
    // phase of the carrier at `warm`: coherent sum over the last 1024 samples, derotated by the estimate
    float ar = 0.f, ai = 0.f;
    for (int j = lane; j < 1024; j += 32) {
        const long long i = (long long)warm - 1024 + j;
        if (i >= 0) {
            float p, q, sn, cs;
            load_iq1(a.iq, a.pcm16, first + (u64)i, p, q);
            est_sincos_turns(-cyc * (float)(j - 1024), sn, cs);
            ar += p * cs - q * sn; ai += p * sn + q * cs;
        }
    }
    for (int o = 16; o; o >>= 1) { ar += __shfl_xor_sync(0xffffffffu, ar, o); ai += __shfl_xor_sync(0xffffffffu, ai, o); }
    LoopState2 g;
    g.a = atan2f(ai, ar);
    g.b = 6.283185307179586f * cyc;
    return g;
}

__global__ void __launch_bounds__(EST_WARPS * 32) k_estimate(const TiledArgs a)
{
    __shared__ float2 zs[EST_WARPS][EST_FFT];
    __shared__ float2 stage[EST_WARPS][32 * EST_MAX_D];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned per_cap = a.pll.max_tiles - 1;
    const u64 wid = (u64)blockIdx.x * EST_WARPS + wib;
    if (per_cap == 0 || wid >= (u64)a.n_captures * per_cap) return;
    const uint32_t cap = (uint32_t)(wid / per_cap);
    const unsigned k = 1 + (unsigned)(wid % per_cap);
    const AcqResult &acq = a.acq[cap];
    const u64 n = cap_len(a, cap), first = (u64)cap * a.stride;
    u64 warm, begin, end;
    if (!cap_selected(a, cap) || !acq.locked || !tile_range(acq.track_begin, n, a.pll, k, warm, begin, end)) return;
    const LoopState2 g = est_carrier(a, first, warm, zs[wib], stage[wib], lane);
    if (lane == 0) a.guess[(size_t)cap * a.pll.max_tiles + k] = g;
}

// Pre-locked start of the captures >= prelock_from (segments of one stream behind the first, pdt.h): instead of the
// reference's acquisition sweep from zero (CarrierTrackingPLL.c:115-262), the loop starts in TRACK mode at sample
// pre = EST_FFT·D from the carrier estimated over [0, pre) — the same guess + warm-up the PLL tiles of a capture use.
// The phase stream of [0, pre) is the estimate extrapolated backwards, so the FIR/AGC see a derotated signal from sample 0.
// Writes the AcqResult k_acquire would have written at a lock latch on sample pre-1.
__host__ __device__ inline u64 prelock_span(int est_decim) { return (u64)EST_FFT * (u64)est_decim; }

__global__ void __launch_bounds__(EST_WARPS * 32) k_prelock(const TiledArgs a)
{
    __shared__ float2 zs[EST_WARPS][EST_FFT];
    __shared__ float2 stage[EST_WARPS][32 * EST_MAX_D];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const u64 wid = (u64)blockIdx.x * EST_WARPS + wib;
    const u64 cap64 = (u64)a.prelock_from + wid;
    if (cap64 >= a.n_captures) return;
    const uint32_t cap = (uint32_t)cap64;
    const u64 n = cap_len(a, cap), first = (u64)cap * a.stride;
    AcqResult *res = &a.acq[cap];
    u64 pre = prelock_span(a.est_decim);
    if (pre > n) pre = n;                                   // (the host refuses segments this short; keep the kernel safe)
    float snr = 0.f;
    LoopState2 g = est_carrier(a, first, pre, zs[wib], stage[wib], lane, &snr);
    g.a = __shfl_sync(0xffffffffu, g.a, 0); g.b = __shfl_sync(0xffffffffu, g.b, 0);
    if (!(snr >= PDT_PRELOCK_MIN_SNR)) {
        // no carrier stands out of the searched band: do not pretend a lock — the segment runs the reference's acquisition
        // sweep from zero like a recording that starts here (k_acquire picks it up), and says so in its stats
        if (lane == 0) { res->prelocked = 2; res->prelock_snr = snr; }
        return;
    }
    PllState ps; pll_reset(ps); pll_begin(ps, a.cc.pll);
    if (g.b > ps.max_freq) g.b = ps.max_freq; else if (g.b < ps.min_freq) g.b = ps.min_freq;
    float *ph = a.ph + (u64)cap * a.ws_stride;
    for (u64 i = lane; i < pre; i += 32) {
        double p = (double)g.a - (double)g.b * (double)(pre - i);
        p -= 6.283185307179586 * rint(p / 6.283185307179586);
        ph[i] = (float)p;
    }
    if (lane == 0) {
        res->locked = 1; res->slow = 0; res->resume_at = n;
        res->lock_sample = pre ? pre - 1 : 0; res->track_begin = pre;
        res->phase = g.a; res->freq = g.b; res->sweep = 0.0f;
        res->avg_phase = 0.0f; res->locksig = 1.0f;
        res->prelocked = 1; res->prelock_snr = snr;
        res->lock_freq_hz = g.b * a.cc.pll.Fs / (2.0 * PDT_PI);
        const float bw = a.cc.pll.bw_track, damp = ps.damp;                                         // CarrierTrackingPLL.c:272-273
        res->alpha = (4.0 * damp * bw) / (1.0 + 2.0 * damp * bw + bw * bw);
        res->beta  = (4.0 * bw * bw) / (1.0 + 2.0 * damp * bw + bw * bw);
    }
}

// ---------------------------------------------------------------------------------------------------
// PLL track core: lane per (capture, tile), a warp = 32 consecutive tiles of one capture, streams moved by the TMA
// ---------------------------------------------------------------------------------------------------
#ifndef PDT_LS_WARPS
#define PDT_LS_WARPS 4                                    // (2 -> 4: k_pll_core 13.8 -> 9.9 ms on the whole batch, r02h)
#endif
constexpr int LS_WARPS = PDT_LS_WARPS;                    // warps per CTA of the lane-stream kernels
constexpr size_t LS_SMEM = LS_WARPS * sizeof(LaneStreamSmem);

// Checkpoints.  While a tile's samples are stored, the loop state (phase, frequency) is also recorded every PLL_CK samples.
// A tile whose speculated start state turns out wrong is re-run from the true state — but only until its state at a
// checkpoint is bit-identical to the recorded one: from there on the two runs are the same run, sample for sample, and the
// phases already stored are the true ones (k_pll_fix_par).  A failed tile is typically off by a few ulps at its first
// sample and merges within the first pieces; without the checkpoints each repair cost a whole tile of serial recurrence
// (65 k samples = 2.3 ms at 250 ksps, 523 k samples = 18 ms at 2 Msps) on the critical path of every batch.
constexpr unsigned PLL_CK = 1024;                         // samples between checkpoints (16 lane-stream rounds)
static_assert(PLL_CK % LS_R == 0, "checkpoints fall on lane-stream round boundaries");

struct PllLaneStep {
    float phase, freq; TrackConst k;
    LoopState2 *ck = nullptr;                             // this tile's checkpoint row (nullptr: none recorded)
    unsigned n_ck = 0;
    __device__ __forceinline__ void round_done(unsigned c)
    {
        const unsigned done = (c + 1) * (unsigned)LS_R;   // samples of this pass behind us
        if (ck && (done % PLL_CK) == 0 && done / PLL_CK < n_ck) ck[done / PLL_CK] = LoopState2{phase, freq};
    }
    __device__ __forceinline__ void quad(const float4 &v, float4 &o)
    {
        o.x = phase; pll_track_step(phase, freq, v.x, k); o.y = phase; pll_track_step(phase, freq, v.y, k);
        o.z = phase; pll_track_step(phase, freq, v.z, k); o.w = phase; pll_track_step(phase, freq, v.w, k);
    }
    __device__ __forceinline__ float one(float v) { const float p = phase; pll_track_step(phase, freq, v, k); return p; }
};

__device__ __forceinline__ void k_pll_core_task(const TiledArgs &a, LaneStream &sm, const int lane, const uint32_t cap, const unsigned k)
{
    bool active = cap < a.n_captures && cap_selected(a, cap);
    u64 warm = 0, begin = 0, end = 0, mid = 0, keep = 0;
    PllLaneStep st; st.phase = 0.f; st.freq = 0.f; st.k = TrackConst{0.f, 0.f, 0.f, 0.f};
    const float *sp = a.sp; float *ph = a.ph;
    size_t slot = 0;
    if (active) {
        const AcqResult &acq = a.acq[cap];
        active = acq.locked && tile_range(acq.track_begin, cap_len(a, cap), a.pll, k, warm, begin, end);
        if (active) {
            st.k = track_const(a, acq);
            sp = a.sp + (u64)cap * a.ws_stride; ph = a.ph + (u64)cap * a.ws_stride;
            slot = (size_t)cap * a.pll.max_tiles + k;
            if (k == 0) {
                st.phase = acq.phase; st.freq = acq.freq;
                // tile 0 starts at the (arbitrary) sample behind the lock latch: walk to the first 16-byte boundary
                u64 a0 = (begin + 3) & ~3ull; if (a0 > end) a0 = end;
                for (u64 i = begin; i < a0; i++) { ph[i] = st.phase; pll_track_step(st.phase, st.freq, sp[i], st.k); }
                a.pll_start[slot] = LoopState2{acq.phase, acq.freq};
                // tile 0 needs no warm-up; it is T0 = W + T long and spends its first W samples (stored) while the other
                // lanes of the warp warm up, so that both passes below are equally long for all 32 lanes
                warm = a0; mid = a0 + a.pll.W; if (mid > end) mid = end;
                keep = warm;
            } else {
                const LoopState2 g = a.guess[slot];
                st.phase = g.a; st.freq = g.b;
                if (st.freq > st.k.max_freq) st.freq = st.k.max_freq; else if (st.freq < st.k.min_freq) st.freq = st.k.min_freq;
                mid = begin; keep = begin;
            }
        }
    }
    if (!active) warm = mid = keep = end = 0;
    const u64 cb = active ? (u64)cap * a.ws_stride : 0;              // lane streams index the whole workspace arrays
    lane_stream<true>(sm, lane, a.sp, a.ph, cb + warm, cb + keep, cb + mid, st);   // warm-up (nothing kept) / first W samples of tile 0
    if (active && k != 0) {
        a.pll_start[slot] = LoopState2{st.phase, st.freq};
        st.ck = a.pll_ckpt + slot * a.pll_nck; st.n_ck = a.pll_nck;               // checkpoints of the stored pass (mid == begin)
    }
    lane_stream<true, PllLaneStep, true>(sm, lane, a.sp, a.ph, cb + mid, cb + mid, cb + end, st);
    if (active) a.pll_end[slot] = LoopState2{st.phase, st.freq};
}

// persistent: a fixed, resident grid of warps walks the compact work list (no CTA is launched for tiles that do not exist)
__global__ void __launch_bounds__(LS_WARPS * 32) k_pll_core(const TiledArgs a)
{
    extern __shared__ __align__(128) unsigned char ls_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    LaneStream sm = lane_stream_init(&reinterpret_cast<LaneStreamSmem *>(ls_raw)[wib], lane);
    const uint32_t n_tasks = a.task_counts[0];
    for (uint32_t t = blockIdx.x * LS_WARPS + wib; t < n_tasks; t += gridDim.x * LS_WARPS) {
        const LaneTask tk = a.pll_tasks[t];
        k_pll_core_task(a, sm, lane, tk.cap, tk.k0 + (unsigned)lane);
    }
}

PDT_DEV bool same_bits(const LoopState2 &x, const LoopState2 &y) { return pdt_f2u(x.a) == pdt_f2u(y.a) && pdt_f2u(x.b) == pdt_f2u(y.b); }

PDT_DEV LoopState2 ld_state(const LoopState2 *p)
{
    // one 8-byte access: a concurrent writer (parallel repair pass) is seen either entirely or not at all
    const unsigned long long v = *reinterpret_cast<const volatile unsigned long long *>(p);
    LoopState2 r; r.a = pdt_u2f((uint32_t)v); r.b = pdt_u2f((uint32_t)(v >> 32));
    return r;
}
PDT_DEV void st_state(LoopState2 *p, const LoopState2 &v)
{
    *reinterpret_cast<volatile unsigned long long *>(p) = (unsigned long long)pdt_f2u(v.a) | ((unsigned long long)pdt_f2u(v.b) << 32);
}

// Parallel repair pass: every tile whose warm-up state differs from the end state of its predecessor is re-run
// from that end state, all such tiles at once.  A predecessor that is itself being repaired in the same pass may
// still change; the state actually used is recorded in pll_start, so the next pass (or the final serial sweep in
// k_pll_fix) sees the difference.  Tiles converge long before their end, so one pass almost always suffices.
__device__ __forceinline__ void k_pll_fix_par_task(const TiledArgs &a, LaneStream &sm, const int lane, const uint32_t cap, const unsigned k)
{
    bool active = cap < a.n_captures && k != 0 && cap_selected(a, cap);
    u64 warm = 0, begin = 0, end = 0;
    PllLaneStep st; st.phase = 0.f; st.freq = 0.f; st.k = TrackConst{0.f, 0.f, 0.f, 0.f};
    size_t slot = 0;
    LoopState2 truth{0.f, 0.f};
    if (active) {
        const AcqResult &acq = a.acq[cap];
        active = acq.locked && tile_range(acq.track_begin, cap_len(a, cap), a.pll, k, warm, begin, end);
        if (active) {
            slot = (size_t)cap * a.pll.max_tiles + k;
            truth = ld_state(&a.pll_end[slot - 1]);
            active = !same_bits(ld_state(&a.pll_start[slot]), truth);
            st.k = track_const(a, acq); st.phase = truth.a; st.freq = truth.b;
        }
    }
    if (!__any_sync(0xffffffffu, active)) return;
    if (!active) begin = end = 0;
    const u64 off = (u64)(active ? cap : 0) * a.ws_stride;
    // re-run piece by piece (PLL_CK samples) and stop at the first checkpoint where the state is bit-identical to the one
    // the stored run recorded there: everything behind it, including the tile's end state, is already the true trajectory
    LoopState2 *ck = a.pll_ckpt + slot * a.pll_nck;
    bool running = active, merged = false;
    u64 at = begin;
    for (unsigned j = 1; __any_sync(0xffffffffu, running); j++) {
        u64 stop = at + PLL_CK; if (stop > end) stop = end;
        const u64 p0 = running ? off + at : 0, p1 = running ? off + stop : 0;
        lane_stream<true>(sm, lane, a.sp, a.ph, p0, p0, p1, st);
        if (running) {
            at = stop;
            if (at >= end) running = false;                               // ran to the end of the tile: new end state below
            else if (j < a.pll_nck) {
                const LoopState2 now{st.phase, st.freq};
                if (same_bits(ld_state(&ck[j]), now)) { running = false; merged = true; }
                else st_state(&ck[j], now);
            }
        }
    }
    if (active) {
        st_state(&a.pll_start[slot], truth);
        if (!merged) st_state(&a.pll_end[slot], LoopState2{st.phase, st.freq});
        atomicAdd(&a.counters[0], 1u);
    }
}

// persistent: a fixed, resident grid of warps walks the compact work list (no CTA is launched for tiles that do not exist)
__global__ void __launch_bounds__(LS_WARPS * 32) k_pll_fix_par(const TiledArgs a)
{
    extern __shared__ __align__(128) unsigned char ls_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    LaneStream sm = lane_stream_init(&reinterpret_cast<LaneStreamSmem *>(ls_raw)[wib], lane);
    const uint32_t n_tasks = a.task_counts[0];
    for (uint32_t t = blockIdx.x * LS_WARPS + wib; t < n_tasks; t += gridDim.x * LS_WARPS) {
        const LaneTask tk = a.pll_tasks[t];
        k_pll_fix_par_task(a, sm, lane, tk.cap, tk.k0 + (unsigned)lane);
    }
}

__global__ void __launch_bounds__(128) k_pll_fix(const TiledArgs a)
{
    const uint32_t cap = blockIdx.x * blockDim.x + threadIdx.x;
    if (cap >= a.n_captures) return;
    const AcqResult &acq = a.acq[cap];
    if (!cap_selected(a, cap) || !acq.locked) return;
    const u64 n = cap_len(a, cap), first = (u64)cap * a.stride;
    const TrackConst kc = track_const(a, acq);
    uint32_t fixed = 0;
    for (unsigned k = 1; k < a.pll.max_tiles; k++) {
        u64 warm, begin, end;
        if (!tile_range(acq.track_begin, n, a.pll, k, warm, begin, end)) break;
        const size_t slot = (size_t)cap * a.pll.max_tiles + k;
        const LoopState2 truth = a.pll_end[slot - 1];
        if (same_bits(a.pll_start[slot], truth)) continue;
        float phase = truth.a, freq = truth.b;                        // speculation failed: re-run from the true state
        pll_track_run<true>(a.sp + (u64)cap * a.ws_stride, a.ph + (u64)cap * a.ws_stride, begin, end, phase, freq, kc);
        a.pll_start[slot] = truth;
        a.pll_end[slot] = LoopState2{phase, freq};
        fixed++;
    }
    if (fixed) atomicAdd(&a.counters[0], fixed);
}

// ---------------------------------------------------------------------------------------------------
// front: out = Im(x·e^{-jφ}) (CarrierTrackingPLL.c:106-113) then the ×L zero-stuff FIR (LowPassFilter.c:43-70)
// ---------------------------------------------------------------------------------------------------
// geometry per interpolation factor: 26 input samples per thread; the CTA shrinks for L >= 5 so that the L-times larger
// output staging still leaves several CTAs per SM.  For L >= 2 every thread's 26·L output row is padded to an odd stride
// (26·L is ≡ 8 / 16 mod 32 for L = 4 / 8: 8- and 16-way bank conflicts on the row-major stores otherwise).
__host__ __device__ constexpr int front_threads(int L) { return L >= 5 ? 64 : 128; }
__host__ __device__ constexpr int front_span(int L) { return front_threads(L) * FIR_K; }          // 3328 / 1664 input samples per CTA
__host__ __device__ constexpr int front_row(int L) { return L == 1 ? FIR_K : FIR_K * L + 1; }
__host__ __device__ constexpr size_t front_smem_bytes(int L)
{
    return sizeof(float) * ((size_t)front_span(L) + FIR_K + 2 + (size_t)front_threads(L) * front_row(L));
}

template <int L>
__global__ void __launch_bounds__(front_threads(L)) k_front(const TiledArgs a, const __grid_constant__ TapsRev taps)
{
    constexpr int FRONT_THREADS = front_threads(L), FRONT_SPAN = front_span(L), ROW = front_row(L);
    extern __shared__ __align__(16) float fsm[];
    float *outs = fsm;                                  // [FRONT_SPAN + FIR_K]  (one block of history in front)
    float *ys = fsm + FRONT_SPAN + FIR_K + 2;           // [FRONT_THREADS][ROW]
    const uint32_t cap = blockIdx.y;
    const int tid = threadIdx.x;
    const u64 n = cap_len(a, cap), first = (u64)cap * a.stride;
    const u64 base = (u64)blockIdx.x * FRONT_SPAN;
    if (base >= n || !cap_selected(a, cap)) return;
    const float *ph = a.ph + (u64)cap * a.ws_stride;
    const pdt_traces *tr = a.traces ? &a.traces[cap] : nullptr;
    {
        // staging pass: sample i = base - FIR_K + idx, valid for idx in [lo, hi); everything is addressed from per-CTA bases
        // with 32-bit offsets, the (rare) trace pointer is fetched once
        const int lo = (base >= (u64)FIR_K) ? 0 : (FIR_K - (int)base);
        const u64 left = n - base;                                              // >= 1
        const int hi = (left + FIR_K < (u64)(FRONT_SPAN + FIR_K)) ? (int)(left + FIR_K) : (FRONT_SPAN + FIR_K);
        const float *ph_b = ph + base - FIR_K;                                  // only dereferenced at idx >= lo
        float *trace_out = (tr && tr->pll_out) ? reinterpret_cast<float *>(tr->pll_out) + base - FIR_K : nullptr;
        if (a.pcm16) {
            const short2 *iq_b = reinterpret_cast<const short2 *>(a.iq) + first + base - FIR_K;
            for (int idx = tid; idx < FRONT_SPAN + FIR_K; idx += FRONT_THREADS) {
                float o = 0.0f;
                if (idx >= lo && idx < hi) {
                    const short2 v = iq_b[idx];
                    const float p = v.x / 32768.0f, q = v.y / 32768.0f;
                    float ti, tr_;
                    sincos_exact(ph_b[idx], ti, tr_);
                    const float nti = -ti;
                    o = p * nti + q * tr_;                                          // :110,:113
                    if (trace_out && idx >= FIR_K) trace_out[idx] = o;
                }
                outs[idx] = o;
            }
        } else {
            const float2 *iq_b = reinterpret_cast<const float2 *>(a.iq) + first + base - FIR_K;
            for (int idx = tid; idx < FRONT_SPAN + FIR_K; idx += FRONT_THREADS) {
                float o = 0.0f;
                if (idx >= lo && idx < hi) {
                    const float2 v = iq_b[idx];
                    float ti, tr_;
                    sincos_exact(ph_b[idx], ti, tr_);
                    const float nti = -ti;
                    o = v.x * nti + v.y * tr_;                                      // :110,:113
                    if (trace_out && idx >= FIR_K) trace_out[idx] = o;
                }
                outs[idx] = o;
            }
        }
    }
    __syncthreads();
    {
        const u64 j0 = base + (u64)tid * FIR_K;
        if (j0 < n) {
            float prev[FIR_K], cur[FIR_K];
#pragma unroll
            for (int s = 0; s < FIR_K; s++) { prev[s] = outs[tid * FIR_K + s]; cur[s] = outs[(tid + 1) * FIR_K + s]; }
            float *dst = ys + (size_t)tid * ROW;
            if constexpr (L == 1) fir_block26<1>(prev, cur, taps, [&](int o, float v) { dst[o] = v; });
            else                  fir_block26_branches<L>(prev, cur, taps, [&](int o, float v) { dst[o] = v; });
        }
    }
    __syncthreads();
    const u64 span = (n - base < FRONT_SPAN) ? (n - base) : FRONT_SPAN;
    float *y = a.y + (u64)cap * a.ws_stride * L + base * L;
    if constexpr (L == 1) {
        for (u64 o = tid; o < span; o += FRONT_THREADS) y[o] = ys[o];
    } else {
        const unsigned total = (unsigned)span * L;
        for (unsigned o = tid; o < total; o += FRONT_THREADS) y[o] = ys[o + o / (unsigned)(FIR_K * L)];      // skip one pad word per row
    }
}

// ---------------------------------------------------------------------------------------------------
// k_front1: the L = 1 front kernel (250 ksps and up: no interpolation, N = 26), rebuilt around Blackwell's PACKED fp32
// pipe.  Same arithmetic as k_front<1> — out = Im(x·e^{-jφ}) with the glibc-exact sincos, then the 26-tap FIR in the
// reference's rotating summation order, every product and every sum rounded separately (LowPassFilter.c:58-64) — but
//   * a thread owns TWO aligned 26-sample blocks, 128 blocks apart (blocks t+1 and t+129 of the CTA's 257 staged blocks),
//     and runs them as the two halves of f32x2 operands: per slot s one `mul.rn.f32x2` and one `fma.rn.f32x2(p, 1, acc)` for
//     both blocks (the 1.0 comes from the parameter bank at run time — with a literal, or with add.rn.f32x2, ptxas folds
//     the pair into a single-rounding FFMA2, tools/f32x2_test.cu).  Both blocks have the same position c in the ring, so
//     they use the same tap for the same slot: the tap pairs (h,h) sit in UNIFORM registers, loaded once per thread from
//     the parameter bank (SASS: FMUL2 R, R, UR / FFMA2 R, R, UR, R).  26 + 26 packed instructions per 2 outputs, against
//     52 + 52 scalar ones;
//   * the loop is slot-major: all 26 accumulator pairs stay in registers and each staged operand pair is loaded once
//     (LDS.128 = two slots), ~70 registers instead of the 116 an output-major packed variant needed (profiles/README.md);
//   * staging writes the derotated samples as (block j, block j+128) float2 pairs — exactly the operand layout — with one
//     conflict-free STS.64 per two samples, no index division: thread q handles samples q and q + 3328 of the span;
//   * the finished 6656-sample tile leaves through ONE bulk copy (cp.async.bulk shared -> global, SASS UBLKCP) issued by
//     one thread, instead of a per-thread LDS/STG loop; the tile aliases the staging buffer (26.8 KB per CTA in total).
// ---------------------------------------------------------------------------------------------------
constexpr int F1_THREADS = 128;
constexpr int F1_BLOCKS = 2 * F1_THREADS;                 // 26-sample blocks produced per CTA
constexpr int F1_SPAN = F1_BLOCKS * FIR_K;                // 6656 input (= output) samples per CTA
constexpr int F1_HALF = F1_THREADS * FIR_K;               // 3328: distance between the two samples of a staged pair
constexpr int F1_PAIRS = (F1_THREADS + 1) * FIR_K;        // 3354 staged float2 pairs (one history block in front)
struct TapsPair { float2 hh[FIR_K]; float one, pad; };    // hh[u] = (hr[u], hr[u]); one = 1.0f (run-time operand, see above)

__device__ __forceinline__ u64 f1_pk(float lo, float hi) { return ((u64)__float_as_uint(hi) << 32) | (u64)__float_as_uint(lo); }
__device__ __forceinline__ u64 f1_mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 f1_fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

// Staging.  Pair q of the operand buffer is (sample i0+q, sample i0+q+3328), i0 = base-26.  A thread takes pairs 2t, 2t+1 of
// every 256-pair round: its two "low" samples and its two "high" samples are adjacent in memory — one 16-byte IQ load and
// one 8-byte phase load each — and the two finished pairs leave as one STS.128.  The loads of round r+1 are issued before
// round r is computed (explicit register double-buffering): ncu on the first version showed half of the staging time
// waiting for the global loads at their first use (profiles/README.md r02b).
struct F1Raw { float4 iq_lo, iq_hi; float2 ph_lo, ph_hi; };      // two IQ samples + two phases, low and high half

// checked loads (edge CTAs: capture start / end inside the span, and the last block's pairs): indices are clamped for the
// loads, f1_derot4<true> selects 0 outside [0, n) — no branch.  PCM: int16 IQ (wave.c:141-166: /32768).
template <bool PCM, bool CHECK>
__device__ __forceinline__ void f1_load(F1Raw &r, const void *__restrict__ iq_cap, const float *__restrict__ ph, long long ilo, long long n)
{
    static_assert(CHECK, "interior CTAs stage through cp.async (f1_stage)");
    auto two = [&](long long i, float4 &iq, float2 &p2) {
        long long i0 = i, i1 = i + 1;
        i0 = i0 < 0 ? 0 : (i0 >= n ? n - 1 : i0); i1 = i1 < 0 ? 0 : (i1 >= n ? n - 1 : i1);
        float a0, b0, a1, b1;
        load_iq1(iq_cap, PCM ? 1 : 0, (u64)i0, a0, b0); load_iq1(iq_cap, PCM ? 1 : 0, (u64)i1, a1, b1);
        iq = make_float4(a0, b0, a1, b1);
        p2 = make_float2(ph[i0], ph[i1]);
    };
    two(ilo, r.iq_lo, r.ph_lo);
    two(ilo + F1_HALF, r.iq_hi, r.ph_hi);
}

template <bool CHECK>
__device__ __forceinline__ float4 f1_derot4(const F1Raw &r, long long ilo, long long n, bool &big)
{
    auto one = [&](float p, float q, float phase, long long i) -> float {
        big |= !sincos_in_core_range(phase);
        float ti, tr;
        sincos_core(phase, ti, tr);
        const float nti = -ti;
        const float o = p * nti + q * tr;                                           // CarrierTrackingPLL.c:110,:113
        return (CHECK && (i < 0 || i >= n)) ? 0.0f : o;
    };
    float4 o;                                                                       // (lo q, hi q, lo q+1, hi q+1) = pairs q, q+1
    o.x = one(r.iq_lo.x, r.iq_lo.y, r.ph_lo.x, ilo);
    o.y = one(r.iq_hi.x, r.iq_hi.y, r.ph_hi.x, ilo + F1_HALF);
    o.z = one(r.iq_lo.z, r.iq_lo.w, r.ph_lo.y, ilo + 1);
    o.w = one(r.iq_hi.z, r.iq_hi.w, r.ph_hi.y, ilo + 1 + F1_HALF);
    return o;
}

// 8-byte cp.async (LDGSTS): the NCO phases of the next round travel global -> shared memory without holding registers, so
// nothing tempts ptxas into sinking the loads next to their first use (it did, under the 72-register budget: r02b)
__device__ __forceinline__ void f1_cp8(uint32_t dst_smem, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ float2 f1_lds8(uint32_t src_smem)
{
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(src_smem) : "memory");
    return v;
}

template <bool PCM, bool CHECK>
__device__ __forceinline__ void f1_stage(float2 *__restrict__ P, float2 (*__restrict__ RP)[F1_THREADS], const void *__restrict__ iq_cap,
                                         const float *__restrict__ ph, const long long base, const long long n, const int tid,
                                         float *__restrict__ trace_out)
{
    constexpr int ROUNDS = FIR_K / 2;                                               // 13 rounds of 256 pairs, then 26 pairs
    const long long i0 = base - FIR_K;
    bool big = false;
    float4 *P4 = reinterpret_cast<float4 *>(P);
    auto at = [&](int r) { return 2 * tid + r * 2 * F1_THREADS; };
    F1Raw A;
    if (!CHECK) {
        // interior CTA.  Round r: the phases of round r arrive through the thread's own two shared-memory slots (requested a
        // whole round earlier), the IQ samples through plain loads issued at the top of the round — they are only needed at
        // the derotation, behind ~40 dependent double-precision operations.
        // Two slot sets: round r first requests round r+1 into the other set and THEN waits for its own data
        // (wait_group 1) — the request is ordered in front of the wait, the wait in front of the phase reads, the
        // arithmetic depends on those: the copy has a whole round of arithmetic to land, whatever the scheduler does.
        // (pointers and shared-memory offsets advance by constants: the round costs no address arithmetic to speak of)
        const float *php = ph + (i0 + 2 * tid);                                   // phases of this thread's low pair, round r
        const char *iqp = PCM ? (const char *)(reinterpret_cast<const short2 *>(iq_cap) + (i0 + 2 * tid))
                              : (const char *)(reinterpret_cast<const float2 *>(iq_cap) + (i0 + 2 * tid));
        constexpr int IQ_B = PCM ? 4 : 8;                                         // bytes per IQ sample
        constexpr uint32_t HALF_B = F1_THREADS * sizeof(float2);                  // RP[k+1][tid] - RP[k][tid] in bytes
        uint32_t slot = (uint32_t)__cvta_generic_to_shared(&RP[0][tid]);          // set 0: RP[0], RP[1]; set 1: RP[2], RP[3]
        uint32_t other = slot + 2 * HALF_B;
        float4 *dst = P4 + tid;
        f1_cp8(slot, php); f1_cp8(slot + HALF_B, php + F1_HALF);
        asm volatile("cp.async.commit_group;" ::: "memory");
#pragma unroll 1
        for (int r = 0; r < ROUNDS; r++) {
            if (r + 1 < ROUNDS) { f1_cp8(other, php + 2 * F1_THREADS); f1_cp8(other + HALF_B, php + 2 * F1_THREADS + F1_HALF); }
            asm volatile("cp.async.commit_group;" ::: "memory");
            if (PCM) {
                const short4 lo = *reinterpret_cast<const short4 *>(iqp);
                const short4 hi = *reinterpret_cast<const short4 *>(iqp + (size_t)F1_HALF * IQ_B);
                A.iq_lo = make_float4(lo.x / 32768.0f, lo.y / 32768.0f, lo.z / 32768.0f, lo.w / 32768.0f);
                A.iq_hi = make_float4(hi.x / 32768.0f, hi.y / 32768.0f, hi.z / 32768.0f, hi.w / 32768.0f);
            } else {
                A.iq_lo = *reinterpret_cast<const float4 *>(iqp);
                A.iq_hi = *reinterpret_cast<const float4 *>(iqp + (size_t)F1_HALF * IQ_B);
            }
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            A.ph_lo = f1_lds8(slot); A.ph_hi = f1_lds8(slot + HALF_B);
            *dst = f1_derot4<false>(A, 0, n, big);
            php += 2 * F1_THREADS; iqp += (size_t)2 * F1_THREADS * IQ_B; dst += F1_THREADS;
            const uint32_t t_ = slot; slot = other; other = t_;
        }
    } else {
#pragma unroll 1
        for (int r = 0; r < ROUNDS; r++) {
            f1_load<PCM, true>(A, iq_cap, ph, i0 + at(r), n);
            P4[at(r) >> 1] = f1_derot4<true>(A, i0 + at(r), n, big);
        }
    }
    // the last block's 26 pairs go to threads 0…12 (checked loads: their high half may pass the end of the capture)
    if (tid < FIR_K / 2) {
        f1_load<PCM, true>(A, iq_cap, ph, i0 + F1_HALF + 2 * tid, n);
        P4[(F1_HALF >> 1) + tid] = f1_derot4<true>(A, i0 + F1_HALF + 2 * tid, n, big);
    }
    if (big || trace_out) {
        // never taken by PLL phases (wrapped to ±2π): phases beyond the branch-free sincos's range (a corrupted phase stream)
        // go through the general routine; the rare trace tap is written here too, off the hot loop
        for (int q = tid; q < F1_PAIRS; q += F1_THREADS) {
            float2 v = P[q];
            const long long idx[2] = {i0 + q, i0 + q + F1_HALF};
            for (int h = 0; h < 2; h++) {
                const long long i = idx[h];
                if (i < 0 || i >= n) continue;
                float p_, q_;
                load_iq1(iq_cap, PCM ? 1 : 0, (u64)i, p_, q_);
                float ti, tr_;
                sincos_exact(ph[i], ti, tr_);
                const float nti = -ti;
                const float o = p_ * nti + q_ * tr_;
                if (h) v.y = o; else v.x = o;
                if (trace_out && i >= base && (h || q < F1_HALF + FIR_K)) trace_out[i] = o;
            }
            P[q] = v;
        }
    }
}

// MODE 0: one CTA per (tile, capture) cell of the grid.
// MODE 1: the launch covers a.cap_list (the slow captures, ~9 % of a batch) with blockIdx.y striding over it, instead of one
//         grid row per capture of which 91 % would exit at once (0.24 ms of empty CTAs per batch, profiles/r02z3_bench.json).
// MODE 2: persistent — a 1-D grid of resident CTAs strides over all (capture, tile) items of this pass (experiment: PDT_FRONT_PERSIST).
template <bool PCM, int MODE>
__global__ void __launch_bounds__(F1_THREADS, 7) k_front1(const TiledArgs a, const __grid_constant__ TapsPair taps)
{
    __shared__ __align__(128) float2 P[F1_PAIRS];          // staged operand pairs; re-used as the output tile (6656 floats)
    __shared__ __align__(16) float2 RP[4][F1_THREADS];     // per-thread landing slots of the phase prefetch: two sets of (low, high half)
    const int tid = threadIdx.x;
    const uint32_t n_list = MODE == 1 ? *a.cap_list_count : 0u;
    const unsigned long long n_items = MODE == 2 ? (unsigned long long)a.n_captures * a.front_tiles : 0ull;
  for (unsigned long long it = (MODE == 2) ? blockIdx.x : blockIdx.y; MODE == 0 || (MODE == 1 ? it < n_list : it < n_items);
       it += (MODE == 2) ? gridDim.x : gridDim.y) {
    uint32_t cap; long long base;
    if (MODE == 2) { cap = (uint32_t)(it / a.front_tiles); base = (long long)(it - (unsigned long long)cap * a.front_tiles) * F1_SPAN; }
    else           { cap = MODE == 1 ? a.cap_list[it] : (uint32_t)it; base = (long long)blockIdx.x * F1_SPAN; }
    const long long n = (long long)cap_len(a, cap);
    if (MODE == 0 && (base >= n || !cap_selected(a, cap))) return;
    if (MODE == 1 && base >= n) continue;                  // (uniform over the CTA; nothing of this capture touched shared memory)
    if (MODE == 2 && (base >= n || !cap_selected(a, cap))) continue;
    const float *__restrict__ ph = a.ph + (u64)cap * a.ws_stride;
    const void *__restrict__ iq_cap = PCM ? (const void *)(reinterpret_cast<const short2 *>(a.iq) + (u64)cap * a.stride)
                                          : (const void *)(reinterpret_cast<const float2 *>(a.iq) + (u64)cap * a.stride);
    const pdt_traces *tr = a.traces ? &a.traces[cap] : nullptr;
    float *trace_out = (tr && tr->pll_out) ? reinterpret_cast<float *>(tr->pll_out) : nullptr;

    // ---- staging: pair q = (sample base-26+q, sample base-26+q+3328) ------------------------------------------------
    // interior CTA whose sample pairs are naturally aligned (the vector loads take two IQ samples at once: a capture that
    // starts at an odd sample offset — odd stride, odd capture index — stages through the scalar, checked path instead)
    const uintptr_t pair_mask = PCM ? 7u : 15u;
    const bool aligned = ((reinterpret_cast<uintptr_t>(iq_cap) + (uintptr_t)(base - FIR_K) * (PCM ? 4u : 8u)) & pair_mask) == 0;
    if (aligned && base >= FIR_K && base + F1_SPAN <= n) f1_stage<PCM, false>(P, RP, iq_cap, ph, base, n, tid, trace_out);
    else                                                 f1_stage<PCM, true>(P, RP, iq_cap, ph, base, n, tid, trace_out);
    __syncthreads();

    // ---- FIR: blocks tid+1 (low halves) and tid+129 (high halves), slot-major --------------------------------------
    u64 acc[FIR_K];
    {
        const u64 one2 = f1_pk(taps.one, taps.one);
#pragma unroll
        for (int c = 0; c < FIR_K; c++) acc[c] = 0ull;                              // +0.0f, +0.0f
        const float4 *prev4 = reinterpret_cast<const float4 *>(P + tid * FIR_K);    // 208 B rows: 16-byte aligned, and the
        const float4 *cur4 = reinterpret_cast<const float4 *>(P + (tid + 1) * FIR_K);   // 8 lanes of an LDS.128 phase hit 32 distinct banks
#pragma unroll
        for (int s2 = 0; s2 < FIR_K / 2; s2++) {
            const float4 xc4 = cur4[s2], xp4 = prev4[s2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int s = 2 * s2 + h;
                const u64 xc = h ? f1_pk(xc4.z, xc4.w) : f1_pk(xc4.x, xc4.y);
                const u64 xp = h ? f1_pk(xp4.z, xp4.w) : f1_pk(xp4.x, xp4.y);
#pragma unroll
                for (int c = 0; c < FIR_K; c++) {
                    // output c of a block sums ring slots s = 0…25 in this order; slot s holds the current block's sample
                    // for s <= c and the previous block's otherwise, times hr[(c - s) mod 26]   (pdt_tiled.cuh::fir_block26)
                    const u64 x = (s <= c) ? xc : xp;
                    const float2 hh = taps.hh[(c - s + FIR_K) % FIR_K];
                    acc[c] = f1_fma2(f1_mul2(x, f1_pk(hh.x, hh.y)), one2, acc[c]);
                }
            }
        }
    }
    __syncthreads();                                       // every operand has been read: the buffer becomes the output tile
    float *ys = reinterpret_cast<float *>(P);
#pragma unroll
    for (int c = 0; c < FIR_K; c++) {
        ys[tid * FIR_K + c] = __uint_as_float((unsigned)acc[c]);
        ys[tid * FIR_K + c + F1_HALF] = __uint_as_float((unsigned)(acc[c] >> 32));
    }
    const long long left = n - base;
    const unsigned span = (unsigned)(left < F1_SPAN ? left : F1_SPAN);
    float *y = a.y + (u64)cap * a.ws_stride + (u64)base;
    const unsigned bytes = ((span + 3u) & ~3u) * 4u;       // rounded up into the row padding (ws_stride is a multiple of 4)
    if ((reinterpret_cast<uintptr_t>(y) & 15) == 0) {
        // generic-proxy writes of the tile -> visible to the async proxy, then one bulk copy shared -> global
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         :: "l"(y), "r"((uint32_t)__cvta_generic_to_shared(ys)), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");       // the tile has been read: the CTA may retire
        }
    } else {
        __syncthreads();
        for (unsigned o = tid; o < span; o += F1_THREADS) y[o] = ys[o];
    }
    if (MODE == 0) return;
    __syncthreads();                                       // the tile has left (thread 0 waited for the bulk read): next item
  }
}

// ---------------------------------------------------------------------------------------------------
// AGC core: lane per (capture, tile)
// ---------------------------------------------------------------------------------------------------
PDT_DEV float agc_guess(const float *y, u64 at)
{
    // equilibrium of AGC.c:98-131 is E|y·g| = 1  ->  g ≈ 1 / mean|y| over the samples in front of the warm-up
    const u64 cnt = at < 1024 ? at : 1024;
    float sum = 0.0f;
    for (u64 i = at - cnt; i < at; i++) sum += fabsf(y[i]);
    if (!(sum > 0.0f)) return 1.0f;
    return (float)cnt / sum;
}

struct AgcFastLaneStep {
    float gain, decay; AgcProof pr;
    __device__ __forceinline__ void quad(const float4 &v, float4 &o)
    {
        o.x = agc_step_fast(gain, v.x, decay, pr); o.y = agc_step_fast(gain, v.y, decay, pr);
        o.z = agc_step_fast(gain, v.z, decay, pr); o.w = agc_step_fast(gain, v.w, decay, pr);
    }
    __device__ __forceinline__ float one(float v) { return agc_step_fast(gain, v, decay, pr); }
};
struct AgcLaneStep {
    AgcState st; float attack, decay;
    __device__ __forceinline__ void quad(const float4 &v, float4 &o)
    {
        o.x = agc_step(st, v.x, attack, decay); o.y = agc_step(st, v.y, attack, decay);
        o.z = agc_step(st, v.z, attack, decay); o.w = agc_step(st, v.w, attack, decay);
    }
    __device__ __forceinline__ float one(float v) { return agc_step(st, v, attack, decay); }
};

// Two lock-step passes for every lane of the warp: [s0, mid) of which [keep, mid) is stored (a tile's warm-up: keep == mid;
// the first W samples of tile 0, which needs no warm-up: keep == s0), then [mid, end) stored.  The proven fast regime
// first; lanes whose proof fails (and only they) repeat their range with the general recurrence.
// Returns the gain at `mid` in mid_gain and the final gain in gain.
__device__ __forceinline__ void agc_tile_warp(LaneStream &sm, const int lane, const float *x, float *z, u64 s0, u64 keep, u64 mid, u64 end,
                                              float &gain, float &mid_gain, float attack, float decay)
{
    const float g0 = gain;
    AgcFastLaneStep fs; fs.gain = g0; fs.decay = decay; fs.pr.lo = 0.0f; fs.pr.hi = 0.0f;
    lane_stream<true>(sm, lane, x, z, s0, keep, mid, fs);
    float gm = fs.gain;
    lane_stream<true>(sm, lane, x, z, mid, mid, end, fs);
    float g = fs.gain;
    const bool redo = (end > s0) && !agc_proof_ok(fs.pr);
    if (__any_sync(0xffffffffu, redo)) {
        AgcLaneStep gsx; gsx.st.init = 1; gsx.st.gain = g0; gsx.attack = attack; gsx.decay = decay;
        const u64 a2 = redo ? s0 : 0, k2 = redo ? keep : 0, m2 = redo ? mid : 0, e2 = redo ? end : 0;
        lane_stream<true>(sm, lane, x, z, a2, k2, m2, gsx);
        if (redo) gm = gsx.st.gain;
        lane_stream<true>(sm, lane, x, z, m2, m2, e2, gsx);
        if (redo) g = gsx.st.gain;
    }
    mid_gain = gm; gain = g;
}

__device__ __forceinline__ void k_agc_core_task(const TiledArgs &a, LaneStream &sm, const int lane, const uint32_t cap, const unsigned k)
{
    const int L = a.cc.L;
    bool active = cap < a.n_captures && cap_selected(a, cap);
    u64 warm = 0, begin = 0, end = 0, first = 0;
    float gain = 1.0f, start_gain = 1.0f;
    size_t slot = 0;
    if (active) {
        const AcqResult &acq = a.acq[cap];
        const u64 nL = cap_len(a, cap) * L;
        first = (u64)cap * a.ws_stride * L;
        const TilePlan plan = agc_plan(a, acq);
        active = tile_range(0, nL, plan, k, warm, begin, end);
        if (active) {
            slot = (size_t)cap * a.agc_max_tiles + k;
            if (k == 0) gain = acq.norm;                                 // AGC.c:92-96: first call seeds gain with `initial`
            else        gain = agc_guess(a.y + first, warm);
        }
    }
    if (!__any_sync(0xffffffffu, active)) return;
    if (!active) { warm = begin = end = 0; first = 0; }
    u64 keep = begin, mid = begin;
    if (active && k == 0) {                // tile 0: no warm-up, its first W samples run (stored) beside the others' warm-up
        const TilePlan plan = agc_plan(a, a.acq[cap]);
        keep = warm; mid = warm + plan.W; if (mid > end) mid = end;
    }
    agc_tile_warp(sm, lane, a.y, a.z, first + warm, first + keep, first + mid, first + end, gain, start_gain, a.cc.agc_attack, a.cc.agc_decay);
    if (active) {
        a.agc_start[slot] = LoopState2{start_gain, 0.0f};
        a.agc_end[slot] = LoopState2{gain, 0.0f};
    }
}

// persistent: a fixed, resident grid of warps walks the compact work list (no CTA is launched for tiles that do not exist)
__global__ void __launch_bounds__(LS_WARPS * 32) k_agc_core(const TiledArgs a)
{
    extern __shared__ __align__(128) unsigned char ls_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    LaneStream sm = lane_stream_init(&reinterpret_cast<LaneStreamSmem *>(ls_raw)[wib], lane);
    const uint32_t n_tasks = a.task_counts[1];
    for (uint32_t t = blockIdx.x * LS_WARPS + wib; t < n_tasks; t += gridDim.x * LS_WARPS) {
        const LaneTask tk = a.agc_tasks[t];
        k_agc_core_task(a, sm, lane, tk.cap, tk.k0 + (unsigned)lane);
    }
}

__device__ __forceinline__ void k_agc_fix_par_task(const TiledArgs &a, LaneStream &sm, const int lane, const uint32_t cap, const unsigned k)
{
    const int L = a.cc.L;
    bool active = cap < a.n_captures && k != 0 && cap_selected(a, cap);
    u64 warm = 0, begin = 0, end = 0, first = 0;
    size_t slot = 0;
    LoopState2 truth{0.f, 0.f};
    if (active) {
        const AcqResult &acq = a.acq[cap];
        const u64 nL = cap_len(a, cap) * L;
        first = (u64)cap * a.ws_stride * L;
        const TilePlan plan = agc_plan(a, acq);
        active = tile_range(0, nL, plan, k, warm, begin, end);
        if (active) {
            slot = (size_t)cap * a.agc_max_tiles + k;
            truth = ld_state(&a.agc_end[slot - 1]);
            active = !same_bits(ld_state(&a.agc_start[slot]), truth);
        }
    }
    if (!__any_sync(0xffffffffu, active)) return;
    if (!active) { begin = end = 0; first = 0; }
    float gain = truth.a, sg;
    agc_tile_warp(sm, lane, a.y, a.z, first + begin, first + begin, first + begin, first + end, gain, sg, a.cc.agc_attack, a.cc.agc_decay);
    if (active) {
        st_state(&a.agc_start[slot], truth);
        st_state(&a.agc_end[slot], LoopState2{gain, 0.0f});
        atomicAdd(&a.counters[1], 1u);
    }
}

// persistent: a fixed, resident grid of warps walks the compact work list (no CTA is launched for tiles that do not exist)
__global__ void __launch_bounds__(LS_WARPS * 32) k_agc_fix_par(const TiledArgs a)
{
    extern __shared__ __align__(128) unsigned char ls_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    LaneStream sm = lane_stream_init(&reinterpret_cast<LaneStreamSmem *>(ls_raw)[wib], lane);
    const uint32_t n_tasks = a.task_counts[1];
    for (uint32_t t = blockIdx.x * LS_WARPS + wib; t < n_tasks; t += gridDim.x * LS_WARPS) {
        const LaneTask tk = a.agc_tasks[t];
        k_agc_fix_par_task(a, sm, lane, tk.cap, tk.k0 + (unsigned)lane);
    }
}

__global__ void __launch_bounds__(128) k_agc_fix(const TiledArgs a)
{
    const uint32_t cap = blockIdx.x * blockDim.x + threadIdx.x;
    if (cap >= a.n_captures || !cap_selected(a, cap)) return;
    const AcqResult &acq = a.acq[cap];
    const int L = a.cc.L;
    const u64 nL = cap_len(a, cap) * L, first = (u64)cap * a.ws_stride * L;
    const TilePlan plan = agc_plan(a, acq);
    uint32_t fixed = 0;
    for (unsigned k = 1; k < a.agc_max_tiles; k++) {
        u64 warm, begin, end;
        if (!tile_range(0, nL, plan, k, warm, begin, end)) break;
        const size_t slot = (size_t)cap * a.agc_max_tiles + k;
        const LoopState2 truth = a.agc_end[slot - 1];
        if (same_bits(a.agc_start[slot], truth)) continue;
        float gain = truth.a, sg;
        agc_tile(a.y + first, a.z + first, begin, begin, end, gain, sg, a.cc.agc_attack, a.cc.agc_decay);
        a.agc_start[slot] = truth;
        a.agc_end[slot] = LoopState2{gain, 0.0f};
        fixed++;
    }
    if (fixed) atomicAdd(&a.counters[1], fixed);
}

// ---------------------------------------------------------------------------------------------------
// back end, two kernels:
//   k_gardner  GardenerClockRecovery.c:24-111 — warp per capture, chunk by chunk (the chunk length is part of the
//              reference's numerics: the sampling position is a float in chunk-relative samples).  The AGC
//              output streams through a per-warp shared-memory window; lane 0 runs the timing recurrence and nothing
//              else (its dependent chain — round, window read, error, clamp, advance — is what bounds this kernel);
//              symbols and pick indices are staged in shared memory and flushed coalesced by the whole warp.
//   k_bits     ManchesterDecode.c:27-97 + ByteSync.c:42-148 over the symbol stream, lane per capture.
// ---------------------------------------------------------------------------------------------------
constexpr int GAR_WIN = 4096;
#ifndef PDT_GAR_WARPS
#define PDT_GAR_WARPS 2
#endif
constexpr int GAR_WARPS = PDT_GAR_WARPS;
constexpr int GAR_STAGE = 512;          // symbols staged per flush (a window yields window/step of them; a full stage just flushes early)


// rintf(x) as an integer.  FAST: x + 1.5·2^23 rounds to nearest-even at unit granularity for |x| < 2^22, and the integer is
// the low mantissa — two dependent ALU ops instead of FRND + F2I (25 cycles, tools/microbench.cu).
template <bool FAST>
PDT_DEV int rint_index(float x)
{
    if (FAST) return (int)pdt_f2u(x + 12582912.0f) - 0x4B400000;
    const float r = rintf(x);
    return (r < 0.0f) ? -1 : (int)(unsigned)r;
}

// one symbol with every check of the reference loop (used for the first symbol of a chunk and near the window / chunk end)
template <bool FAST>
__device__ __forceinline__ void gardner_careful(float &next, float &prev, float &half, bool &first_symbol, int at, const float *__restrict__ win,
                                                int w_lo, int w_hi, const float *__restrict__ z, u64 ibase, unsigned n_out, unsigned full_out,
                                                bool has_prev, float kp, float range, float step, float &sym, float &err)
{
    const float cur = win[min((unsigned)(at - w_lo), (unsigned)(GAR_WIN - 1))];
    const int hi = rint_index<FAST>(half);                                             // :28 index-then-value reuse
    float hv;
    if (!first_symbol && hi >= w_lo && hi < w_hi) hv = win[hi - w_lo];
    else hv = stale_lookup(z, ibase, (unsigned)hi, n_out, full_out, has_prev);
    first_symbol = false;
    const float e = clamp_like_ifs(kp * (cur - prev) * hv, -range, range);             // :43-57
    next = next - e;
    half = next + step / 2.0;                                                          // :59
    next = next + step;
    prev = cur;
    sym = cur; err = e;
}

template <bool FAST>
__device__ __forceinline__ void gardner_capture(const TiledArgs &a, const uint32_t cap, float *__restrict__ win, float *__restrict__ ssym,
                                                float *__restrict__ serr, unsigned *__restrict__ sidx, const int lane)
{
    const ChainConst &cc = a.cc;
    const int L = cc.L;
    const u64 n = cap_len(a, cap);
    const float *__restrict__ z = a.z + (u64)cap * a.ws_stride * L;
    float *__restrict__ sym_out = a.sym + (u64)cap * a.sym_cap;
    u64 *__restrict__ gidx_out = a.gidx + (u64)cap * a.sym_cap;
    const pdt_traces *tr = a.traces ? &a.traces[cap] : nullptr;
    const unsigned full_out = cc.chunk * (unsigned)L;
    const float kp = cc.g_kp, range = cc.g_range;

    GardnerState gs = GardnerState();
    gardner_begin(gs, cc.gardner_fs, cc.baud);
    const float step = gs.step;
    // half = next + step/2.0 is a double sum narrowed to float (GardenerClockRecovery.c:59).  The double sum of two floats
    // is exact when their exponents are < 29 apart, and then the float sum is the same correctly rounded value: the batch
    // loop below uses the float sum and re-runs the batch with the double expression if `next` ever was tiny but non-zero.
    const float half_step = (float)((double)step / 2.0);
    const bool half_float_ok = ((double)half_step == (double)step / 2.0) && step > 1.0f && step < 1.0e6f;
    const float inv_adv = 1.0f / (step + fabsf(range) + 0.01f);
    u64 n_sym = 0;

    for (u64 base = 0; base < n; base += cc.chunk) {
        const unsigned m = (unsigned)((n - base < cc.chunk) ? (n - base) : cc.chunk);
        const unsigned n_out = m * (unsigned)L;
        const u64 ibase = base * (u64)L;
        const bool has_prev = base > 0;
        bool first_symbol = true, reload = true;
        unsigned w0 = 0;
        for (;;) {
            // resident window: chunk-relative [w0, w0 + GAR_WIN)
            if (!reload) {
            } else if (((ibase + w0) & 3) == 0) {
#pragma unroll 8
                for (unsigned i = lane * 4; i < GAR_WIN; i += 128) {
                    const unsigned idx = w0 + i;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (idx + 4 <= n_out) v = ld4(z + ibase + idx);
                    else {
                        if (idx < n_out) v.x = z[ibase + idx];
                        if (idx + 1 < n_out) v.y = z[ibase + idx + 1];
                        if (idx + 2 < n_out) v.z = z[ibase + idx + 2];
                    }
                    st4(win + i, v);
                }
            } else {
#pragma unroll 8
                for (unsigned i = lane; i < GAR_WIN; i += 32) {
                    const unsigned idx = w0 + i;
                    win[i] = (idx < n_out) ? z[ibase + idx] : 0.0f;
                }
            }
            __syncwarp();
            int finished = 0, cnt = 0;
            if (lane == 0) {
                float next = gs.next, prev = gs.prev, half = gs.half;
                const int w_lo = (int)w0, w_hi = (int)(w0 + GAR_WIN);
                const int limit = ((int)n_out < w_hi) ? (int)n_out : w_hi;
                for (;;) {
                    const int at = rint_index<FAST>(next);
                    if (at >= (int)n_out) { finished = 1; break; }                     // GardenerClockRecovery.c:24
                    if (at >= w_hi || cnt >= GAR_STAGE) break;                         // refill / flush
                    // symbols that certainly stay below `limit`: the position advances by at most step + |range| per symbol
                    int k = 0;
                    if (!first_symbol && half_float_ok) {
                        const float room = (float)limit - 4.0f - next;
                        k = (room > 0.0f) ? 1 + (int)(room * inv_adv) : 0;
                        if (k > GAR_STAGE - cnt) k = GAR_STAGE - cnt;
                    }
                    if (k < 2) {
                        float sv, ev;
                        gardner_careful<FAST>(next, prev, half, first_symbol, at, win, w_lo, w_hi, z, ibase, n_out, full_out, has_prev,
                                              kp, range, step, sv, ev);
                        ssym[cnt] = sv; sidx[cnt] = (unsigned)at; if (serr) serr[cnt] = ev;
                        cnt++;
                        continue;
                    }
                    // ---- the hot loop: k symbols, no exit test, no window test (both proven above), float mid-point sum ----
                    const float next0 = next, prev0 = prev, half0 = half;
                    bool tiny = false;
                    for (int j = 0; j < k; j++) {
                        const int aj = rint_index<FAST>(next);
                        const float cur = win[min((unsigned)(aj - w_lo), (unsigned)(GAR_WIN - 1))];
                        const int hj = rint_index<FAST>(half);
                        const float hv = win[min((unsigned)(hj - w_lo), (unsigned)(GAR_WIN - 1))];
                        const float e = clamp_like_ifs(kp * (cur - prev) * hv, -range, range);
                        next = next - e;
                        tiny |= !(fabsf(next) >= 9.5367431640625e-07f || next == 0.0f);
                        half = next + half_step;
                        next = next + step;
                        prev = cur;
                        ssym[cnt + j] = cur; sidx[cnt + j] = (unsigned)aj; if (serr) serr[cnt + j] = e;
                    }
                    if (tiny) {                 // never observed; keeps the result exact by construction
                        next = next0; prev = prev0; half = half0;
                        for (int j = 0; j < k; j++) {
                            const int aj = rint_index<FAST>(next);
                            float sv, ev;
                            gardner_careful<FAST>(next, prev, half, first_symbol, aj, win, w_lo, w_hi, z, ibase, n_out, full_out, has_prev,
                                                  kp, range, step, sv, ev);
                            ssym[cnt + j] = sv; sidx[cnt + j] = (unsigned)aj; if (serr) serr[cnt + j] = ev;
                        }
                    }
                    cnt += k;
                }
                gs.next = next; gs.prev = prev; gs.half = half;
            }
            finished = __shfl_sync(0xffffffffu, finished, 0);
            cnt = __shfl_sync(0xffffffffu, cnt, 0);
            __syncwarp();
            for (int j = lane; j < cnt; j += 32) {
                const u64 k = n_sym + (u64)j;
                if (k < a.sym_cap) { sym_out[k] = ssym[j]; gidx_out[k] = ibase + sidx[j]; }
                if (tr && k < tr->cap) {
                    if (tr->sym)         reinterpret_cast<float *>(tr->sym)[k] = ssym[j];
                    if (tr->gardner_err) reinterpret_cast<float *>(tr->gardner_err)[k] = serr[j];
                    if (tr->gardner_idx) tr->gardner_idx[k] = ibase + sidx[j];
                }
            }
            n_sym += (u64)cnt;
            if (finished) break;
            // next window starts a little before the pick so that the mid-sample (behind the pick) stays resident
            int at = 0;
            if (lane == 0) at = rint_index<FAST>(gs.next);
            at = __shfl_sync(0xffffffffu, at, 0);
            const unsigned back = (unsigned)step + 8;
            const unsigned nw0 = (unsigned)at > back ? (unsigned)at - back : 0;
            reload = (unsigned)at >= w0 + GAR_WIN;                             // a pure stage flush keeps the window
            if (reload) w0 = nw0 & ~3u;
            __syncwarp();
        }
        if (lane == 0) gs.next = gs.next - n_out;                                      // :111
        __syncwarp();
    }
    if (lane == 0) { GarRecord r; r.n_sym = n_sym; r.final_next = gs.next; r.pad = 0.f; a.gar[cap] = r; }
}

// GAR_SPREAD: the CTA is launched with 4 warps of which 2 work — warps {0,1} in even CTAs, {2,3} in odd ones; the other two
// retire at once.  A warp's scheduler is its index modulo 4: with plain 2-warp CTAs every capture's serial lane would sit
// on sub-cores 0 and 1 of its SM and the other two schedulers would never see this kernel.
#ifndef PDT_GAR_SPREAD
#define PDT_GAR_SPREAD 0
#endif
constexpr int GAR_CTA_THREADS = PDT_GAR_SPREAD ? 128 : GAR_WARPS * 32;
__global__ void __launch_bounds__(GAR_CTA_THREADS) k_gardner(const TiledArgs a)
{
    __shared__ __align__(16) float wins[GAR_WARPS][GAR_WIN];
    __shared__ float ssym[GAR_WARPS][GAR_STAGE], serr[GAR_WARPS][GAR_STAGE];
    __shared__ unsigned sidx[GAR_WARPS][GAR_STAGE];
    const int lane = threadIdx.x & 31;
    int wib = threadIdx.x >> 5;
    if (PDT_GAR_SPREAD) {
        wib -= 2 * (int)(blockIdx.x & 1u);
        if (wib < 0 || wib >= GAR_WARPS) return;
    }
    const uint32_t cap = blockIdx.x * GAR_WARPS + wib;
    if (cap >= a.n_captures || !cap_selected(a, cap)) return;
    const bool fast = (double)a.cc.chunk * a.cc.L + 64.0 < 4.0e6;
    float *e = (a.traces && a.traces[cap].gardner_err) ? serr[wib] : nullptr;
    if (fast) gardner_capture<true>(a, cap, wins[wib], ssym[wib], e, sidx[wib], lane);
    else      gardner_capture<false>(a, cap, wins[wib], ssym[wib], e, sidx[wib], lane);
}

// ---------------------------------------------------------------------------------------------------
// k_bits: ManchesterDecode.c:27-97 + ByteSync.c:42-148 over the symbol stream, warp per capture, 32 symbols per step.
//  * Manchester: the only state carried from symbol to symbol is the pair phase `clockmod`, and a symbol either sets it
//    to its own parity (two strong equal-sign symbols in front of it, :42-50) or leaves it alone, so the phase seen
//    by lane l is the parity of the last such symbol at or before l (one ballot), else the carried one.
//  * ByteSync: every new bit gets the 32-bit history ending at it (a funnel shift of the carried history and this step's
//    bits), all sync comparisons of the step happen at once, and the frame shifter only walks the few events.
// ---------------------------------------------------------------------------------------------------
constexpr int BITS_WARPS = 4;
constexpr int BITS_FRAME_WORDS = (PDT_FRAME_MAX_BYTES + 3) / 4;

__global__ void __launch_bounds__(BITS_WARPS * 32) k_bits(const TiledArgs a)
{
    __shared__ unsigned fbuf_all[BITS_WARPS][BITS_FRAME_WORDS];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t cap = blockIdx.x * BITS_WARPS + wib;
    if (cap >= a.n_captures || !cap_selected(a, cap)) return;
    const ChainConst &cc = a.cc;
    const SyncParams sp = cc.sync;
    const u64 n = cap_len(a, cap);
    const int L = cc.L;
    const GarRecord gr = a.gar[cap];
    const float *__restrict__ sym = a.sym + (u64)cap * a.sym_cap;
    const u64 *__restrict__ gidx = a.gidx + (u64)cap * a.sym_cap;
    pdt_frame *frames = a.frames + (size_t)cap * cc.max_frames;
    const pdt_traces *tr = a.traces ? &a.traces[cap] : nullptr;
    unsigned *fbuf = fbuf_all[wib];
    const float thr = cc.man_thresh;
    const u64 ns = gr.n_sym < a.sym_cap ? gr.n_sym : a.sym_cap;
    const unsigned lt_mask = (1u << lane) - 1u, le_mask = lt_mask | (1u << lane);
    // frame geometry: bits a frame takes after the sync bit, and where its first bit lands in the stored byte string
    const int frame_bits = 8 * (sp.last_idx - 1) - sp.carry_bits;              // 813 (POES) / 56 (ARGOS)
    const int q_base = 8 * cc.prefix_bytes + sp.carry_bits;
    const int full_bytes = cc.prefix_bytes + (sp.last_idx - 1);

    unsigned cm = 0, hist = 0;                 // Manchester pair phase; ByteSync history (newest bit = LSB)
    float carry1 = 0.0f, carry2 = 0.0f;        // sym[base-1], sym[base-2]
    int in_frame = 0, fb = 0, inv = 0, cur_frame = -1;
    u64 n_bits = 0; uint32_t n_frames = 0;

    auto flush_frame = [&](int n_bytes, int complete) {        // whole warp; cur_frame >= 0
        __syncwarp();
        pdt_frame &f = frames[cur_frame];
        for (int t = lane; t < PDT_FRAME_MAX_BYTES; t += 32) {
            const unsigned w = fbuf[t >> 2];
            f.bytes[t] = (t < n_bytes) ? (uint8_t)(w >> (24 - 8 * (t & 3))) : (uint8_t)0;
        }
        if (lane == 0) { f.n_bytes = (uint8_t)n_bytes; f.complete = (uint8_t)complete; }
        __syncwarp();
    };

    for (u64 base = 0; base < ns; base += 32) {
        const u64 i = base + (u64)lane;
        const bool valid = i < ns;
        const float c = valid ? sym[i] : 0.0f;
        float p = __shfl_up_sync(FULL, c, 1), pp = __shfl_up_sync(FULL, c, 2);
        if (lane == 0) { p = carry1; pp = carry2; }
        if (lane == 1) pp = carry1;
        const bool cond = valid && sign_of(pp) == sign_of(p) && fabsf(pp) > thr && fabsf(p) > thr;      // ManchesterDecode.c:42-50
        const unsigned below = __ballot_sync(FULL, cond) & le_mask;
        const unsigned par = (unsigned)(i & 1);
        const unsigned cm_l = below ? (((unsigned)base + (31u - (unsigned)__clz((int)below))) & 1u) : cm;
        const bool emit = valid && par == cm_l;                                                        // :52
        const unsigned bit = (fabsf(p) > fabsf(c)) ? (p > 0 ? 1u : 0u) : (c > 0 ? 0u : 1u);           // :60-82
        const unsigned eb = __ballot_sync(FULL, emit);
        const int k = __popc(eb), pos = __popc(eb & lt_mask);
        cm = __shfl_sync(FULL, cm_l, 31);
        carry2 = __shfl_sync(FULL, c, 30); carry1 = __shfl_sync(FULL, c, 31);
        if (k == 0) continue;
        const unsigned v = __reduce_or_sync(FULL, emit ? (bit << (k - 1 - pos)) : 0u);                 // this step's bits, oldest = MSB
        if (tr && tr->bits && emit && n_bits + (u64)pos < tr->cap) tr->bits[n_bits + (u64)pos] = (uint8_t)('0' + bit);
        const unsigned long long H = ((unsigned long long)hist << k) | (unsigned long long)v;
        const unsigned hj = (lane < k) ? (unsigned)(H >> (k - 1 - lane)) : 0u;                         // history ending at new bit `lane`
        const unsigned bn = __ballot_sync(FULL, lane < k && (hj & sp.mask) == sp.word);                // ByteSync.c:104-118
        const unsigned bi = sp.inverse_enabled ? __ballot_sync(FULL, lane < k && (~hj & sp.mask) == sp.word) : 0u;   // :120-133
        int j0 = 0, sf = 0;
        for (;;) {
            if (in_frame) {
                int take = k - j0; if (take > frame_bits - fb) take = frame_bits - fb;
                if (lane >= j0 && lane < j0 + take && cur_frame >= 0) {
                    const unsigned b = ((v >> (k - 1 - lane)) & 1u) ^ (unsigned)inv;                   // :46-52 zero/one swap for an inverted stream
                    const int q = q_base + fb + (lane - j0);
                    if (b) atomicOr(&fbuf[q >> 5], 1u << (31 - (q & 31)));
                }
                fb += take; j0 += take;
                if (fb < frame_bits) break;                       // all new bits went into the open frame
                if (cur_frame >= 0) flush_frame(full_bytes, 1);   // :58-72 frame complete
                in_frame = 0; cur_frame = -1;
                sf = j0 - 1;                                      // the closing bit itself is examined for a sync word (emission precedes detection)
            }
            const unsigned m = (bn | bi) & ~((sf >= 32) ? FULL : ((1u << sf) - 1u));
            if (m == 0) break;
            const int jm = __ffs((int)m) - 1;
            inv = ((bn >> jm) & 1u) ? 0 : 1;
            cur_frame = (n_frames < cc.max_frames) ? (int)n_frames : -1;
            n_frames++;
            if (cur_frame >= 0) {
                __syncwarp();
                for (int t = lane; t < BITS_FRAME_WORDS; t += 32) fbuf[t] = (t == 0 && cc.prefix_bytes) ? 0xEDE20000u : 0u;
                if (emit && pos == jm) {
                    pdt_frame &f = frames[cur_frame];
                    f.sample_index = gidx[i]; f.bit_index = (uint32_t)(n_bits + (u64)jm);
                    f.inverse = (uint8_t)inv; f.complete = 0; f.pad = 0; f.n_bytes = (uint8_t)cc.prefix_bytes;
                    if (cc.prefix_bytes) { f.bytes[0] = 0xED; f.bytes[1] = 0xE2; }
                }
                __syncwarp();
            }
            in_frame = 1; fb = 0; j0 = jm + 1;
            if (j0 >= k) break;
        }
        n_bits += (u64)k;
        hist = (unsigned)H;
    }
    // a capture may end inside a frame: the bytes completed so far are kept
    if (in_frame && cur_frame >= 0) flush_frame(cc.prefix_bytes + (sp.carry_bits + fb) / 8, 0);
    if (lane == 0) {
        const AcqResult &acq = a.acq[cap];
        pdt_capture_stats s;
        s.n_samples = n; s.n_symbols = gr.n_sym; s.n_bits = n_bits; s.n_frames = n_frames;
        s.locked = acq.locked; s.lock_sample = acq.lock_sample; s.lock_freq_hz = acq.lock_freq_hz;
        s.norm_factor = acq.norm; s.avg_phase = acq.avg_phase;
        s.final_phase = acq.phase; s.final_freq = acq.freq;
        if (acq.locked) {
            u64 warm, begin, end;
            for (unsigned k = 0; k < a.pll.max_tiles; k++) {
                if (!tile_range(acq.track_begin, n, a.pll, k, warm, begin, end)) break;
                const LoopState2 e = a.pll_end[(size_t)cap * a.pll.max_tiles + k];
                s.final_phase = e.a; s.final_freq = e.b;
            }
        }
        s.final_gain = acq.norm;
        {
            const TilePlan plan = agc_plan(a, acq);
            u64 warm, begin, end;
            for (unsigned k = 0; k < a.agc_max_tiles; k++) {
                if (!tile_range(0, n * L, plan, k, warm, begin, end)) break;
                s.final_gain = a.agc_end[(size_t)cap * a.agc_max_tiles + k].a;
            }
        }
        s.final_next = gr.final_next;
        s.prelocked = acq.prelocked; s.prelock_snr = acq.prelock_snr;
        a.stats[cap] = s;
    }
}

} // namespace tiled
} // namespace pdt
