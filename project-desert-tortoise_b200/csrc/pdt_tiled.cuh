// pdt_tiled.cuh — the TILED engine (float / POES chain): the serial recurrences of the reference are split into
//   * a thin exact serial core per loop (PLL phase/frequency, AGC gain, Gardner timing), run lane-per-tile, and
//   * fat feed-forward work (approx-atan2 of the raw sample, NCO sincos + derotation, FIR/interpolator) run
//     data-parallel,
// and the loops are parallelised in TIME: every tile re-derives its start state from a guess over a warm-up
// window and the result is accepted only if that state is BIT-IDENTICAL to the end state of the previous tile
// (otherwise the tile is re-run from the true state).  Results are therefore exactly those of the serial
// reference order; the speculation only decides how fast they are obtained.
//
// Reference being restated (file:line in the reference repo):
//   CarrierTrackingPLL.c:102-275 (per-sample PLL), LowPassFilter.c:43-70 (zero-stuff interpolating FIR),
//   AGC.c:98-131 (NormalizingAGC), AGC.c:48-75 (StaticGain), GardenerClockRecovery.c:24-111,
//   ManchesterDecode.c:27-97, POESTIPdemod/ByteSync.c:42-148; driver loop POESTIPdemod/main.c:373-482.
//
// Everything that decides a result is __host__ __device__ so that host-side study tools (tools/acq_study.cu) run the very same
// arithmetic on the CPU against the oracle; the kernels below only map work items to threads.
#pragma once

#include "pdt_common.cuh"

namespace pdt {
namespace tiled {

typedef unsigned long long u64;

// ---------------------------------------------------------------------------------------------------
// exact float-only forms of the reference's double-precision 2π wraps.
//   (float)((double)d - 2*M_PI) == (d - HI) - MID   for every float d in [3, 10.5]   (and mirrored)
// checked exhaustively over all 15.2 M floats of that range in tests/test_tiled_math.py.  d - HI is exact
// (Sterbenz), MID = float(2π - HI), and the residual 2π - HI - MID = -7.1e-15 never reaches a rounding boundary.
//   (double)d >  M_PI   <=>  d >=  PI_UP      (PI_UP = float(π) is the first float above π)
//   (double)p > 2*M_PI  <=>  p >=  TWO_PI_HI  (float(2π) is the first float above 2π)
// ---------------------------------------------------------------------------------------------------
#define PDT_PI_UP      0x1.921fb6p+1f
#define PDT_TWO_PI_HI  0x1.921fb6p+2f
#define PDT_TWO_PI_MID (-0x1.777a5cp-23f)

PDT_DEV float wrap_down(float d) { return (d - PDT_TWO_PI_HI) - PDT_TWO_PI_MID; }
PDT_DEV float wrap_up(float d)   { return (d + PDT_TWO_PI_HI) + PDT_TWO_PI_MID; }

struct TrackConst { float alpha, beta, max_freq, min_freq; };

// The state-carrying part of CarrierTrackingPLL.c:128-188 once the lock latch has fired (no sweep, constant
// gains): sample_phase `sp` is state-independent and precomputed.  Written branch-free (both wrap candidates are
// computed, then selected): ~13 dependent float ops ≈ 55 cycles per sample on sm_100a, against ~190 for the
// branchy form (tools/microbench.cu).  IEEE add/sub is sign-symmetric, so the selects reproduce the
// reference's if / else-if chains bit for bit.
PDT_DEV float wrap_pm(float v, float thresh)
{
    const float dn = wrap_down(v), up = wrap_up(v);
    float r = v;
    r = (v <= -thresh) ? up : r;
    r = (v >= thresh) ? dn : r;
    return r;
}

// if (f > hi) f = hi; else if (f < lo) f = lo;   with the comparisons' NaN behaviour (a NaN stays a NaN)
PDT_DEV float clamp_like_ifs(float f, float lo, float hi)
{
#ifdef __CUDA_ARCH__
    float r;
    asm("{\n\t.reg .f32 t;\n\tmax.NaN.f32 t, %1, %2;\n\tmin.NaN.f32 %0, t, %3;\n\t}" : "=f"(r) : "f"(f), "f"(lo), "f"(hi));
    return r;
#else
    return (f > hi) ? hi : ((f < lo) ? lo : f);
#endif
}

PDT_DEV void pll_track_step(float &phase, float &freq, float sp, const TrackConst &k)
{
    const float err = wrap_pm(sp - phase, PDT_PI_UP);             // :165-170
    const float f = freq + k.beta * err;                          // :174
    const float p = phase + f + k.alpha * err;                    // :175
    phase = wrap_pm(p, PDT_TWO_PI_HI);                            // :178-182 (one step suffices: |Δphase| < 2π, checked at create)
    freq = clamp_like_ifs(f, k.min_freq, k.max_freq);             // :185-188
}

// ---------------------------------------------------------------------------------------------------
// per-capture records
// ---------------------------------------------------------------------------------------------------
struct AcqResult {
    int   locked;               // latch fired inside the capture
    int   slow;                 // not latched within the first acquisition pass: handled by the slow-capture pipeline
    u64   resume_at;            // where the second acquisition pass continues (loop state below is the state before it)
    u64   lock_sample;          // ℓ : absolute index of the latch sample
    u64   track_begin;          // first sample handled by the tiled track core (ℓ+1, or n when never locked)
    float phase, freq;          // PLL state BEFORE sample track_begin
    float alpha, beta;          // track gains (CarrierTrackingPLL.c:272-273)
    float avg_phase, locksig;   // EMA states after sample track_begin-1
    float norm;                 // StaticGain (main.c:384-389) or the override
    float sweep;
    float agc_gain_est;         // 1 / mean|FIR output| (k_agc_plan): where the AGC gain will settle
    float prelock_snr;          // peak / mean of the carrier estimate's spectrum (k_prelock)
    double lock_freq_hz;
    int   prelocked, pad3;      // pdt_capture_stats.prelocked: 0 reference acquisition, 1 started in track mode from the estimate,
                                // 2 estimate rejected -> reference acquisition from zero
};

struct LoopState2 { float a, b; };   // (phase, freq) for the PLL, (gain, init) for the AGC

// tile geometry: tile 0 = [first, a0+T0) starts from the exact state; tile k>=1 = [a0+T0+(k-1)T, +T) with a
// warm-up window of W samples in front of it.  a0 = first rounded up to 4 so that every later boundary is
// float4-aligned.  T0 >= W guarantees that no warm-up reaches in front of `first`.
struct TilePlan { u64 T, W, T0; unsigned max_tiles; };

PDT_DEV bool tile_range(u64 first, u64 n, const TilePlan &p, unsigned k, u64 &warm, u64 &begin, u64 &end)
{
    if (first >= n) return false;
    const u64 a0 = (first + 3) & ~3ull;
    if (k == 0) { warm = begin = first; end = a0 + p.T0; }
    else { begin = a0 + p.T0 + (u64)(k - 1) * p.T; end = begin + p.T; warm = begin - p.W; }
    if (begin >= n) return false;
    if (end > n) end = n;
    return true;
}

PDT_DEV float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
PDT_DEV void   st4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }

// Lane-serial streaming loops.  Each lane walks its own tile, so its loads are 16-byte pieces of a private stream:
// nothing hides their DRAM latency (~1000-1500 cycles under load, ncu: 50 % long-scoreboard stalls with one quad
// group of look-ahead) except distance.  PF_Q quads are kept in flight: a quad is re-loaded for the NEXT 4·PF_Q-sample
// round as soon as it has been consumed, i.e. 4·PF_Q samples (>= 2000 cycles of recurrence) ahead of its use.
constexpr int PF_Q = 8;

// run the track core over sp[i0, i1); STORE: write the phase used to derotate sample i to ph[i]
template <bool STORE>
PDT_DEV void pll_track_run(const float *__restrict__ sp, float *__restrict__ ph, u64 i0, u64 i1, float &phase, float &freq,
                           const TrackConst &k)
{
    u64 i = i0;
    for (; i < i1 && (i & 3); i++) { if (STORE) ph[i] = phase; pll_track_step(phase, freq, sp[i], k); }
    if (i + 4 * PF_Q <= i1) {
        float4 c[PF_Q];
#pragma unroll
        for (int q = 0; q < PF_Q; q++) c[q] = ld4(sp + i + 4 * q);
        for (; i + 4 * PF_Q <= i1; i += 4 * PF_Q) {
            const bool more = i + 8 * PF_Q <= i1;
#pragma unroll
            for (int q = 0; q < PF_Q; q++) {
                const float4 v = c[q];
                if (more) c[q] = ld4(sp + i + 4 * PF_Q + 4 * q);
                float4 o;
                o.x = phase; pll_track_step(phase, freq, v.x, k); o.y = phase; pll_track_step(phase, freq, v.y, k);
                o.z = phase; pll_track_step(phase, freq, v.z, k); o.w = phase; pll_track_step(phase, freq, v.w, k);
                if (STORE) st4(ph + i + 4 * q, o);
            }
        }
    }
    for (; i < i1; i++) { if (STORE) ph[i] = phase; pll_track_step(phase, freq, sp[i], k); }
}

// same shape for the AGC (AGC.c:98-131); x -> z = x*gain
template <bool STORE>
PDT_DEV void agc_run(const float *__restrict__ x, float *__restrict__ z, u64 i0, u64 i1, AgcState &st, float attack, float decay)
{
    u64 i = i0;
    for (; i < i1 && (i & 3); i++) { const float v = agc_step(st, x[i], attack, decay); if (STORE) z[i] = v; }
    if (i + 4 * PF_Q <= i1) {
        float4 c[PF_Q];
#pragma unroll
        for (int q = 0; q < PF_Q; q++) c[q] = ld4(x + i + 4 * q);
        for (; i + 4 * PF_Q <= i1; i += 4 * PF_Q) {
            const bool more = i + 8 * PF_Q <= i1;
#pragma unroll
            for (int q = 0; q < PF_Q; q++) {
                const float4 v = c[q];
                if (more) c[q] = ld4(x + i + 4 * PF_Q + 4 * q);
                float4 o;
                o.x = agc_step(st, v.x, attack, decay); o.y = agc_step(st, v.y, attack, decay);
                o.z = agc_step(st, v.z, attack, decay); o.w = agc_step(st, v.w, attack, decay);
                if (STORE) st4(z + i + 4 * q, o);
            }
        }
    }
    for (; i < i1; i++) { const float v = agc_step(st, x[i], attack, decay); if (STORE) z[i] = v; }
}

// The AGC loop almost always sits in one regime: |error| <= gain, so the DECAY rate applies (AGC.c:108-118 picks the
// attack rate only while the gain is below the error, i.e. for gains < ~1), and neither clamp (:124-131) fires.  There
// the recurrence is four dependent float ops:  z = x·g;  e = |z| - 1;  g = g - e·decay   (16 cycles against 45 for the
// general step with its predicated selects, tools/microbench.cu).  agc_run_fast assumes that regime and PROVES it on the
// side: `lo` = min over the run of (gain - |error|) and of the new gain, `hi` = max of the new gain, NaN-propagating.  The
// caller accepts the result only if lo >= 0 and hi <= 5000 and otherwise re-runs the range with the general step, so
// the output is the reference's in every case.
struct AgcProof { float lo, hi; };

PDT_DEV float min_nan(float a, float b)
{
#ifdef __CUDA_ARCH__
    float r; asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r;
#else
    return (a != a || b != b) ? NAN : (a < b ? a : b);
#endif
}
PDT_DEV float max_nan(float a, float b)
{
#ifdef __CUDA_ARCH__
    float r; asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r;
#else
    return (a != a || b != b) ? NAN : (a > b ? a : b);
#endif
}

PDT_DEV float agc_step_fast(float &gain, float x, float decay, AgcProof &pr)
{
    const float zv = x * gain;
    const float err = fabsf(zv) - 1.0f;
    const float g = gain - err * decay;
    pr.lo = min_nan(pr.lo, min_nan(gain - fabsf(err), g));
    pr.hi = max_nan(pr.hi, g);
    gain = g;
    return zv;
}

PDT_DEV bool agc_proof_ok(const AgcProof &pr) { return pr.lo >= 0.0f && pr.hi <= 5000.0f; }

template <bool STORE>
PDT_DEV void agc_run_fast(const float *__restrict__ x, float *__restrict__ z, u64 i0, u64 i1, float &gain, float decay, AgcProof &pr)
{
    u64 i = i0;
    for (; i < i1 && (i & 3); i++) { const float v = agc_step_fast(gain, x[i], decay, pr); if (STORE) z[i] = v; }
    if (i + 4 * PF_Q <= i1) {
        float4 c[PF_Q];
#pragma unroll
        for (int q = 0; q < PF_Q; q++) c[q] = ld4(x + i + 4 * q);
        for (; i + 4 * PF_Q <= i1; i += 4 * PF_Q) {
            const bool more = i + 8 * PF_Q <= i1;
#pragma unroll
            for (int q = 0; q < PF_Q; q++) {
                const float4 v = c[q];
                if (more) c[q] = ld4(x + i + 4 * PF_Q + 4 * q);
                float4 o;
                o.x = agc_step_fast(gain, v.x, decay, pr); o.y = agc_step_fast(gain, v.y, decay, pr);
                o.z = agc_step_fast(gain, v.z, decay, pr); o.w = agc_step_fast(gain, v.w, decay, pr);
                if (STORE) st4(z + i + 4 * q, o);
            }
        }
    }
    for (; i < i1; i++) { const float v = agc_step_fast(gain, x[i], decay, pr); if (STORE) z[i] = v; }
}

// a whole [warm-up +] tile: the proven fast regime first, the general recurrence if the proof fails
PDT_DEV void agc_tile(const float *__restrict__ x, float *__restrict__ z, u64 warm, u64 begin, u64 end, float &gain, float &start_gain,
                      float attack, float decay)
{
    const float g0 = gain;
    AgcProof pr; pr.lo = 0.0f; pr.hi = 0.0f;
    float g = g0;
    if (warm < begin) agc_run_fast<false>(x, z, warm, begin, g, decay, pr);
    float gs = g;
    agc_run_fast<true>(x, z, begin, end, g, decay, pr);
    if (!agc_proof_ok(pr)) {
        AgcState st; st.init = 1; st.gain = g0;
        if (warm < begin) agc_run<false>(x, z, warm, begin, st, attack, decay);
        gs = st.gain;
        agc_run<true>(x, z, begin, end, st, attack, decay);
        g = st.gain;
    }
    start_gain = gs; gain = g;
}

// ---------------------------------------------------------------------------------------------------
// IQ access (cf32 or int16 PCM, wave.c:141-166: int16 / 32768)
// ---------------------------------------------------------------------------------------------------
PDT_DEV void load_iq1(const void *base, int pcm16, u64 idx, float &a, float &b)
{
    if (pcm16) {
        const short2 v = reinterpret_cast<const short2 *>(base)[idx];
        a = v.x / 32768.0f; b = v.y / 32768.0f;
    } else {
        const float2 v = reinterpret_cast<const float2 *>(base)[idx];
        a = v.x; b = v.y;
    }
}

// ---------------------------------------------------------------------------------------------------
// FIR: one aligned block of 26 input samples -> 26·L outputs, in the reference's rotating summation order
// (LowPassFilter.c:58-64; derivation in pdt_device.cuh::fir_interp_exact).  With K = N/L = 26 taps per
// branch and j = 26m + c the ring holds the current block in slots 0..c and the previous block in slots
// c+1..25, and the reference sums slot 0 first:
//     y[L·j + p] = Σ_{s=0..c} hr[p + L(c-s)]·cur[s]  +  Σ_{s=c+1..25} hr[p + L(c-s+26)]·prev[s]
// with hr[u] = h[N-1-u].  Fully unrolled: every tap index is a compile-time constant, so the taps are read
// straight from the kernel-parameter constant bank.
// ---------------------------------------------------------------------------------------------------
constexpr int FIR_K = 26;
constexpr int FIR_MAX_L = 8;         // 8 = the historical 8x interpolator (18.75 ksps recordings)
struct TapsRev { float hr[FIR_K * FIR_MAX_L]; };

template <int L, typename OUT>
PDT_DEV void fir_block26(const float (&prev)[FIR_K], const float (&cur)[FIR_K], const TapsRev &t, OUT &&put)
{
#pragma unroll
    for (int c = 0; c < FIR_K; c++) {
#pragma unroll
        for (int p = 0; p < L; p++) {
            float acc = 0.0f;
#pragma unroll
            for (int s = 0; s <= c; s++) acc += t.hr[p + L * (c - s)] * cur[s];
#pragma unroll
            for (int s = c + 1; s < FIR_K; s++) acc += t.hr[p + L * (c - s + FIR_K)] * prev[s];
            put(c * L + p, acc);
        }
    }
}

// The same sums for L >= 2, one polyphase branch at a time: for a fixed p the 26 outputs y[L·j + p] are the L = 1 block
// with the branch taps hp[k] = hr[p + L·k].  The branch loop is NOT unrolled — the body is the 676-MAC L = 1 block with
// its taps in registers (26 indexed constant-bank loads per branch), so the code stays ~22 KB for every L instead of
// L x 22 KB (k_front<8> fully unrolled is 170 KB of straight-line code and runs out of the instruction cache).
template <int L, typename OUT>
PDT_DEV void fir_block26_branches(const float (&prev)[FIR_K], const float (&cur)[FIR_K], const TapsRev &t, OUT &&put)
{
#pragma unroll 1
    for (int p = 0; p < L; p++) {
        float hp[FIR_K];
#pragma unroll
        for (int k = 0; k < FIR_K; k++) hp[k] = t.hr[p + L * k];
#pragma unroll
        for (int c = 0; c < FIR_K; c++) {
            float acc = 0.0f;
#pragma unroll
            for (int s = 0; s <= c; s++) acc += hp[c - s] * cur[s];
#pragma unroll
            for (int s = c + 1; s < FIR_K; s++) acc += hp[c - s + FIR_K] * prev[s];
            put(c * L + p, acc);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// carrier guess for a tile: frequency and phase of the residual carrier at sample `at`, from the Wd = 1024·D
// samples in front of it.  Only a GUESS (it seeds the warm-up); nothing here needs to be bit-reproducible
// against the reference, and the accept/re-run verification makes the final result independent of it.
// ---------------------------------------------------------------------------------------------------
constexpr int EST_FFT = 1024;

PDT_DEV void est_sincos_turns(float turns, float &s, float &c)
{
#ifdef __CUDA_ARCH__
    sincospif(2.0f * turns, &s, &c);
#else
    s = sinf(6.2831853071795865f * turns); c = cosf(6.2831853071795865f * turns);
#endif
}

// Jacobsen 3-bin peak interpolation on a rectangular-window DFT; returns the fractional bin offset
PDT_DEV float est_peak_offset(float2 xm, float2 x0, float2 xp)
{
    const float nr = xm.x - xp.x, ni = xm.y - xp.y;
    const float dr = 2.0f * x0.x - xm.x - xp.x, di = 2.0f * x0.y - xm.y - xp.y;
    const float den = dr * dr + di * di;
    if (den <= 0.0f) return 0.0f;
    float d = (nr * dr + ni * di) / den;
    if (d > 0.5f) d = 0.5f; else if (d < -0.5f) d = -0.5f;
    return d;
}

// ---------------------------------------------------------------------------------------------------
// Gardner + Manchester + ByteSync over one capture, streaming through a resident window of the AGC output.
// ---------------------------------------------------------------------------------------------------
struct BackState {
    GardnerState    gar;
    ManchesterState man;
    SyncState       sync;
    u64             n_sym, n_bits;
    uint32_t        n_frames;
    int             cur_frame;      // slot being filled (-1 none / table full)
    int             cur_n;          // bytes written to it so far
};

PDT_DEV void back_reset(BackState &st)
{
    st = BackState();
    st.sync.one = 1;
    st.cur_frame = -1;
}

// symbol -> Manchester -> ByteSync -> frame table (same bookkeeping as the exact engine)
PDT_DEV void back_consume(BackState &st, const ChainConst &cc, float sym, u64 abs_interp_idx, pdt_frame *frames,
                          const pdt_traces *tr)
{
    unsigned char bit;
    if (!manchester_step(st.man, sym, cc.man_thresh, bit)) return;
    if (tr && tr->bits && st.n_bits < tr->cap) tr->bits[st.n_bits] = bit;
    int emit, eol; unsigned char byte;
    const int ev = sync_step(st.sync, cc.sync, bit, emit, byte, eol);
    if (emit && st.cur_frame >= 0) {
        pdt_frame &f = frames[st.cur_frame];
        if (st.cur_n < PDT_FRAME_MAX_BYTES) f.bytes[st.cur_n++] = byte;
        if (eol) { f.n_bytes = (uint8_t)st.cur_n; f.complete = 1; st.cur_frame = -1; }
    } else if (eol) st.cur_frame = -1;
    if (ev != EV_NONE) {
        if (st.n_frames < cc.max_frames) {
            st.cur_frame = (int)st.n_frames;
            pdt_frame &f = frames[st.cur_frame];
            f.sample_index = abs_interp_idx; f.bit_index = (uint32_t)st.n_bits;
            f.inverse = (ev == EV_SYNC_INV); f.complete = 0; f.pad = 0;
            f.n_bytes = (uint8_t)cc.prefix_bytes; st.cur_n = cc.prefix_bytes;
            if (cc.prefix_bytes) { f.bytes[0] = 0xED; f.bytes[1] = 0xE2; }
        } else st.cur_frame = -1;
        st.n_frames++;
    }
    st.n_bits++;
}

// same, for the two-kernel back end: the pick index of a symbol is only needed when its bit completes a sync word
PDT_DEV void back_consume_lazy(BackState &st, const ChainConst &cc, float sym, const u64 *gidx, u64 sym_index, pdt_frame *frames,
                               const pdt_traces *tr)
{
    unsigned char bit;
    if (!manchester_step(st.man, sym, cc.man_thresh, bit)) return;
    if (tr && tr->bits && st.n_bits < tr->cap) tr->bits[st.n_bits] = bit;
    int emit, eol; unsigned char byte;
    const int ev = sync_step(st.sync, cc.sync, bit, emit, byte, eol);
    if (emit && st.cur_frame >= 0) {
        pdt_frame &f = frames[st.cur_frame];
        if (st.cur_n < PDT_FRAME_MAX_BYTES) f.bytes[st.cur_n++] = byte;
        if (eol) { f.n_bytes = (uint8_t)st.cur_n; f.complete = 1; st.cur_frame = -1; }
    } else if (eol) st.cur_frame = -1;
    if (ev != EV_NONE) {
        if (st.n_frames < cc.max_frames) {
            st.cur_frame = (int)st.n_frames;
            pdt_frame &f = frames[st.cur_frame];
            f.sample_index = gidx[sym_index]; f.bit_index = (uint32_t)st.n_bits;
            f.inverse = (ev == EV_SYNC_INV); f.complete = 0; f.pad = 0;
            f.n_bytes = (uint8_t)cc.prefix_bytes; st.cur_n = cc.prefix_bytes;
            if (cc.prefix_bytes) { f.bytes[0] = 0xED; f.bytes[1] = 0xE2; }
        } else st.cur_frame = -1;
        st.n_frames++;
    }
    st.n_bits++;
}

// a capture may end inside a frame: record how many bytes of it were produced
PDT_DEV void back_finish(BackState &st, pdt_frame *frames)
{
    if (st.cur_frame >= 0) frames[st.cur_frame].n_bytes = (uint8_t)st.cur_n;
}

// Value the reference would read at chunk-relative index `idx` of the interpolated/AGC'd chunk buffer when idx
// lies at or beyond the n_out samples just written (GardenerClockRecovery.c:28 after a roll-over):
// the buffer is chunk·L long and is refilled from index 0 every chunk, so [n_out, chunk·L) still holds the
// PREVIOUS chunk's samples (only possible for a short last chunk) and everything beyond is never written (0).
PDT_DEV float stale_lookup(const float *z_capture, u64 chunk_base_interp, unsigned idx, unsigned n_out, unsigned full_out, bool has_prev)
{
    if (idx < n_out) return z_capture[chunk_base_interp + idx];
    if (idx < full_out && has_prev) return z_capture[chunk_base_interp - full_out + idx];
    return 0.0f;
}

} // namespace tiled
} // namespace pdt
