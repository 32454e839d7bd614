// pdt_pll_pipe.cuh — CarrierTrackPLL over one call's samples by a whole CTA, bit-identical to the one-thread loop
// (pdt_device.cuh::pll_step, CarrierTrackingPLL.c:102-275), float and double builds, acquisition and track mode alike.
//
// One thread walking pll_step spends ~1 000 (float) to ~3 000 (double) cycles per sample on a dependent instruction
// stream of which only the loop filter (:165-188, ~100 cycles) is a true recurrence.  Everything else either does not
// depend on the loop state (first-order atan2 of the sample, its Q_rsqrt normalisation) or depends on it only through
// the NCO phase of the same sample (sincos, derotation, the two averaged terms).  The loop state is fed back from those
// only through two booleans per sample: "averaged phase looks like noise" (→ sweep the frequency, :232-246) and
// "lock signal above threshold" (→ latch, :266-274).  So a block of up to PP_B samples runs in four phases:
//
//   P  all threads   sample k: load, sample phase, normalised sample            (state-independent)
//   C  thread 0      loop filter over the block, keeping the state BEFORE every sample, ASSUMING the sweep flag keeps
//                    the value it had after the previous sample and the latch does not fire
//   H  all threads   sample k: sincos of its NCO phase, derotation, output, the two EMA terms
//   E  threads 0/32  the two EMAs in order; each stops at the first sample that contradicts the assumption
//
// If sample j contradicts it, samples ≤ j are exact as computed (their phases depend on the flags of samples < j only):
// the state before j is restored, sample j's filter step repeated with its actual flag, the latch applied, and the next
// block starts at j + 1 — short at first, doubling while the assumption holds.  In track mode nothing is assumed.
// Nothing here is approximate: every value is produced by the same statement as in pll_step.
#pragma once

#include "pdt_device.cuh"
#if PDT_USE_FLOATS
#include "pdt_tiled.cuh"             // pll_track_step: the float-only exact forms of the 2π wraps
#endif

namespace pdt {

// Branch-free forms of pll_loop_core / pll_sweep_core (selects instead of if-chains and while-loops: ~65 (float) / ~100
// (double) cycles per sample against ~320 for the branchy form, tools/chain_prof.py).  Bit-identical WHILE ONE 2π WRAP PER
// SAMPLE SUFFICES, i.e. |Δphase| <= max_freq + |sweep| + (alpha + beta)·π < 4 (and, float build, the wrapped values stay
// inside [3, 10.5], the range tests/test_tiled_math.py checks the float-only wrap forms over): pll_fast_ok() — the caller
// falls back to the reference-shaped functions otherwise.
PDT_DEV bool pll_fast_ok(const PllState &s)
{
    return (double)s.max_freq + 10.0 * ((double)s.alpha + (double)s.beta) < 4.0 && (double)s.min_freq == -(double)s.max_freq;
}

#if PDT_USE_FLOATS
PDT_DEV void pll_loop_fast(float &phase, float &freq, float sp, const tiled::TrackConst &k) { tiled::pll_track_step(phase, freq, sp, k); }
#else
namespace tiled { struct TrackConst { double alpha, beta, max_freq, min_freq; }; }
PDT_DEV void pll_loop_fast(double &phase, double &freq, double sp, const tiled::TrackConst &k)
{
    const double d = sp - phase;                                      // :165-170
    double err = d;
    err = (d < -PDT_PI) ? d + 2 * PDT_PI : err;
    err = (d > PDT_PI) ? d - 2 * PDT_PI : err;
    double f = freq + k.beta * err;                                   // :174
    double ph = phase + f + k.alpha * err;                            // :175
    ph = (ph > 2 * PDT_PI) ? ph - 2.0 * PDT_PI : ph;                  // :178-182, one trip each
    ph = (ph < -2 * PDT_PI) ? ph + 2.0 * PDT_PI : ph;
    f = (f > k.max_freq) ? k.max_freq : ((f < k.min_freq) ? k.min_freq : f);     // :185-188
    phase = ph; freq = f;
}
#endif

PDT_DEV void pll_sweep_fast(real_t &freq, real_t &sweep, real_t max_freq, real_t min_freq)      // :233-246 by selects
{
    const real_t f2 = freq + sweep;
    real_t s2 = (f2 >= 0) ? r_fabs(sweep) : -r_fabs(sweep);
    s2 = (f2 <= min_freq) ? -sweep : s2;
    s2 = (f2 >= max_freq) ? -sweep : s2;
    freq = f2; sweep = s2;
}

constexpr int PP_B = 256;            // samples per block = threads of the CTA that runs it
constexpr int PP_B_MIN = 8;          // block length right after a contradicted assumption

struct PllPipeSmem {
    real_t pa[PP_B], pb[PP_B], sp[PP_B], nre[PP_B], nim[PP_B];   // P: inputs that do not depend on the loop
    real_t ph[PP_B], fq[PP_B], sw[PP_B];                         // C: loop state BEFORE each sample
    real_t out[PP_B], at[PP_B], lt[PP_B];                        // H: derotated output, EMA terms
    real_t av[PP_B], lk[PP_B];                                   // E: EMA values AFTER each sample
    real_t end_phase, end_freq, end_sweep;
    int    ja, jl;                                               // first contradicted sweep flag / first latch (block length if none)
};

// `load(i, a, b)`: sample i of this call;  `emit(i, out, lock, phase_before, freq_before)`: results of sample i (called by
// one thread per sample, any order inside a block, each sample exactly once with its final values).
// `s` lives in shared memory; all CTA threads call this with identical arguments (blockDim.x == PP_B).
// `pf` (optional, thread 0 only): cycle accounting [0] P, [1] C, [2] H, [3] E, [4] emit, [5] blocks, [6] contradicted blocks, [7] control.
template <class Load, class Emit>
__device__ __forceinline__ void pll_run_blocks(PllState &s, const PllParams &p, unsigned long long n, unsigned long long abs0, PllPipeSmem &S,
                                               Load load, Emit emit, unsigned long long *pf = nullptr)
{
    const int tid = threadIdx.x;
    const real_t avg_alpha = 0.00005;
    int cur = PP_B;
    unsigned long long i = 0;
    while (i < n) {
        const int cnt = (int)((n - i < (unsigned long long)cur) ? (n - i) : (unsigned long long)cur);
        const bool acq = s.stage == 1;
        const bool guess = acq && pll_noise_like(s.avg_phase);
        const long long t0 = clock64();
        // ---- P ----------------------------------------------------------------------------------------------------
        if (tid < cnt) {
            real_t a, b;
            load(i + tid, a, b);
            S.pa[tid] = a; S.pb[tid] = b;
            S.sp[tid] = arctan2_approx(b, a);                    // :128
            real_t nre = a, nim = b;                             // :193-218
            const real_t mag2 = nre * nre + nim * nim;
            const real_t inv  = q_rsqrt((float)mag2);
            nre *= inv; nim *= inv;
            S.nre[tid] = nre; S.nim[tid] = nim;
        }
        __syncthreads();
        const long long t1 = clock64();
        // ---- C ----------------------------------------------------------------------------------------------------
        if (tid == 0) {
            real_t phase = s.phase, freq = s.freq, sweep = s.sweep;
            const real_t alpha = s.alpha, beta = s.beta, maxf = s.max_freq, minf = s.min_freq;
            if (pll_fast_ok(s)) {
                tiled::TrackConst kc; kc.alpha = alpha; kc.beta = beta; kc.max_freq = maxf; kc.min_freq = minf;
                int k = 0;
                for (; k + 4 <= cnt; k += 4) {                   // the four sample phases are fetched ahead of the dependent chain
                    const real_t s0 = S.sp[k], s1 = S.sp[k + 1], s2 = S.sp[k + 2], s3 = S.sp[k + 3];
                    S.ph[k] = phase; S.fq[k] = freq; S.sw[k] = sweep;
                    pll_loop_fast(phase, freq, s0, kc); if (guess) pll_sweep_fast(freq, sweep, maxf, minf);
                    S.ph[k + 1] = phase; S.fq[k + 1] = freq; S.sw[k + 1] = sweep;
                    pll_loop_fast(phase, freq, s1, kc); if (guess) pll_sweep_fast(freq, sweep, maxf, minf);
                    S.ph[k + 2] = phase; S.fq[k + 2] = freq; S.sw[k + 2] = sweep;
                    pll_loop_fast(phase, freq, s2, kc); if (guess) pll_sweep_fast(freq, sweep, maxf, minf);
                    S.ph[k + 3] = phase; S.fq[k + 3] = freq; S.sw[k + 3] = sweep;
                    pll_loop_fast(phase, freq, s3, kc); if (guess) pll_sweep_fast(freq, sweep, maxf, minf);
                }
                for (; k < cnt; k++) {
                    S.ph[k] = phase; S.fq[k] = freq; S.sw[k] = sweep;
                    pll_loop_fast(phase, freq, S.sp[k], kc); if (guess) pll_sweep_fast(freq, sweep, maxf, minf);
                }
            } else {
                for (int k = 0; k < cnt; k++) {
                    S.ph[k] = phase; S.fq[k] = freq; S.sw[k] = sweep;
                    pll_loop_core(phase, freq, S.sp[k], alpha, beta, maxf, minf);
                    if (guess) pll_sweep_core(freq, sweep, maxf, minf);
                }
            }
            S.end_phase = phase; S.end_freq = freq; S.end_sweep = sweep;
        }
        __syncthreads();
        const long long t2 = clock64();
        // ---- H ----------------------------------------------------------------------------------------------------
        if (tid < cnt) {
            real_t ti, tr;
            sincos_exact(S.ph[tid], ti, tr);                     // :106-107
            const real_t a = S.pa[tid], b = S.pb[tid], nti = -ti;
            const real_t mre = a * tr - b * nti;                 // :110
            const real_t mim = a * nti + b * tr;
            S.out[tid] = mim;                                    // :113
            S.at[tid] = avg_alpha * r_fabs(arctan2_approx(mim, mre));            // :117, :124
            S.lt[tid] = p.lock_alpha * (S.nre[tid] * tr + S.nim[tid] * ti);      // :220
        }
        __syncthreads();
        const long long t3 = clock64();
        // ---- E ----------------------------------------------------------------------------------------------------
        if (tid == 0) {
            real_t avg = s.avg_phase;
            int ja = cnt;
            for (int k0 = 0; k0 < cnt && ja == cnt; k0 += 4) {
                real_t v[4];
#pragma unroll
                for (int q = 0; q < 4; q++) v[q] = S.at[(k0 + q < PP_B) ? k0 + q : PP_B - 1];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int k = k0 + q;
                    if (k < cnt && ja == cnt) {
                        avg = avg * (1.0 - avg_alpha) + v[q];    // :124
                        S.av[k] = avg;
                        if (acq && pll_noise_like(avg) != guess) ja = k;
                    }
                }
            }
            S.ja = ja;
        } else if (tid == 32) {
            real_t lks = s.locksig;
            int jl = cnt;
            for (int k0 = 0; k0 < cnt && jl == cnt; k0 += 4) {
                real_t v[4];
#pragma unroll
                for (int q = 0; q < 4; q++) v[q] = S.lt[(k0 + q < PP_B) ? k0 + q : PP_B - 1];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int k = k0 + q;
                    if (k < cnt && jl == cnt) {
                        lks = lks * (1.0 - p.lock_alpha) + v[q]; // :220
                        S.lk[k] = lks;
                        if (acq && lks > p.lock_thresh) jl = k;
                    }
                }
            }
            S.jl = jl;
        }
        __syncthreads();
        const long long t4 = clock64();
        const int j = (S.ja < S.jl) ? S.ja : S.jl;
        const int valid = (j < cnt) ? j + 1 : cnt;
        if (tid < valid) emit(i + tid, S.out[tid], S.lk[tid], S.ph[tid], S.fq[tid]);
        __syncthreads();                                         // every reader of `s` and of this block's arrays is done
        const long long t4b = clock64();
        if (tid == 0) {
            if (j < cnt) {                                       // sample j again, with what it really saw
                real_t phase = S.ph[j], freq = S.fq[j], sweep = S.sw[j];
                pll_loop_core(phase, freq, S.sp[j], s.alpha, s.beta, s.max_freq, s.min_freq);
                if (pll_noise_like(S.av[j])) pll_sweep_core(freq, sweep, s.max_freq, s.min_freq);
                s.phase = phase; s.freq = freq; s.sweep = sweep;
                s.avg_phase = S.av[j]; s.locksig = S.lk[j];
                if (S.lk[j] > p.lock_thresh) pll_latch(s, p, abs0 + i + j);
            } else {
                s.phase = S.end_phase; s.freq = S.end_freq; s.sweep = S.end_sweep;
                s.avg_phase = S.av[cnt - 1]; s.locksig = S.lk[cnt - 1];
            }
        }
        cur = (j < cnt) ? PP_B_MIN : ((cur * 2 < PP_B) ? cur * 2 : PP_B);
        i += valid;
        __syncthreads();
        if (pf && tid == 0) {
            const long long t5 = clock64();
            pf[0] += t1 - t0; pf[1] += t2 - t1; pf[2] += t3 - t2; pf[3] += t4 - t3; pf[4] += t4b - t4; pf[5] += 1; pf[6] += (j < cnt); pf[7] += t5 - t4b;
        }
    }
}

} // namespace pdt
