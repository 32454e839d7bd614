// pdt_pll_pipe.cuh — CarrierTrackPLL over one call's samples by a whole CTA, bit-identical to the one-thread loop
// (pdt_device.cuh::pll_step, CarrierTrackingPLL.c:102-275), float and double builds, acquisition and track mode alike.
//
// One thread walking pll_step spends ~1 000 (float) to ~3 000 (double) cycles per sample on a dependent instruction
// stream of which only the loop filter (:165-188) is a true recurrence.  Everything else either does not depend on the
// loop state (first-order atan2 of the sample, its Q_rsqrt normalisation) or depends on it only through the NCO phase
// of the same sample (sincos, derotation, the two averaged terms).  The loop state is fed back from those only through
// two booleans per sample: "averaged phase looks like noise" (→ sweep the frequency, :232-246) and "lock signal above
// threshold" (→ latch, :266-274).  The call is therefore cut into blocks of up to PP_B samples and each block goes
// through
//
//   P  one thread per sample   load, sample phase, normalised sample                     (state-independent)
//   C  thread 0                loop filter over the block, keeping the state BEFORE every sample, ASSUMING the sweep flag
//                              keeps the value it had and the latch does not fire
//   H  one thread per sample   sincos of its NCO phase, derotation, output, the two EMA terms
//   E  threads 32 / 64         the two EMAs in order; each stops at the first sample that contradicts the assumption
//
// software-pipelined over two buffers:  step k runs  C(block k) ‖ E(block k-1),  then  H(block k) and P(block k+1).
// If sample j of block k-1 contradicts the assumption, its samples ≤ j are exact as computed (their phases depend on the
// flags of samples < j only): the loop state before j is restored, sample j's filter step repeated with its actual flag,
// the latch applied, block k dropped, and the pipeline restarts at j + 1 — with short blocks at first, doubling while
// the assumption holds.  In track mode nothing is assumed.  Nothing here is approximate: every value is produced by the
// same statement as in pll_step (or by a select form of it that is bit-identical under a stated condition, below).
#pragma once

#include "pdt_device.cuh"
#if PDT_USE_FLOATS
#include "pdt_tiled.cuh"             // pll_track_step: the float-only exact forms of the 2π wraps
#endif

namespace pdt {

// Branch-free forms of pll_loop_core / pll_sweep_core (selects instead of if-chains and while-loops: ~320 cycles per
// sample for the branchy form, tools/chain_prof.py).  Bit-identical WHILE ONE 2π WRAP PER SAMPLE SUFFICES, i.e.
// |Δphase| <= max_freq + |sweep| + (alpha + beta)·π < 4 (and, float build, the wrapped values stay inside [3, 10.5], the
// range tests/test_tiled_math.py checks the float-only wrap forms over): pll_fast_ok() — otherwise the runner uses the
// reference-shaped functions.
PDT_DEV bool pll_fast_ok(const PllState &s)
{
    return (double)s.max_freq + 10.0 * ((double)s.alpha + (double)s.beta) < 4.0 && (double)s.min_freq == -(double)s.max_freq;
}

#if PDT_USE_FLOATS
typedef tiled::TrackConst PllLoopConst;
PDT_DEV void pll_loop_fast(float &phase, float &freq, float sp, const PllLoopConst &k) { tiled::pll_track_step(phase, freq, sp, k); }
#else
struct PllLoopConst { double alpha, beta, max_freq, min_freq; };
PDT_DEV void pll_loop_fast(double &phase, double &freq, double sp, const PllLoopConst &k)
{
    // all comparisons of a stage are taken on the same value, so they issue side by side; a value that was wrapped down
    // (it was > π resp. > 2π) cannot satisfy the opposite condition afterwards, so one select chain equals the if-chains
    const double d = sp - phase;                                      // :165-170
    const double d_dn = d - 2 * PDT_PI, d_up = d + 2 * PDT_PI;
    double err = (d < -PDT_PI) ? d_up : d;
    err = (d > PDT_PI) ? d_dn : err;
    const double f = freq + k.beta * err;                             // :174
    const double p0 = phase + f + k.alpha * err;                      // :175
    const double p_dn = p0 - 2.0 * PDT_PI, p_up = p0 + 2.0 * PDT_PI;  // :178-182, one trip each
    double ph = (p0 < -2 * PDT_PI) ? p_up : p0;
    ph = (p0 > 2 * PDT_PI) ? p_dn : ph;
    double fc = (f < k.min_freq) ? k.min_freq : f;                    // :185-188
    fc = (f > k.max_freq) ? k.max_freq : fc;
    phase = ph; freq = fc;
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

constexpr int PP_B = 128;            // samples per block
constexpr int PP_THREADS = 128;      // threads of the CTA that runs it: one per sample in P and H; C on thread 0, the EMAs on 32 and 64
                                     // (a narrow CTA on purpose: the runner is latency-bound, residency is what buys throughput)
constexpr int PP_B_MIN = 8;          // block length right after a contradicted assumption

struct PllBlockBuf {
    real_t pa[PP_B], pb[PP_B], sp[PP_B], nre[PP_B], nim[PP_B];   // P: inputs that do not depend on the loop
    real_t ph[PP_B], fq[PP_B], sw[PP_B];                         // C: loop state BEFORE each sample
    real_t at[PP_B], lt[PP_B];                                   // H: the two EMA terms
    real_t av[PP_B], lk[PP_B];                                   // E: EMA values AFTER each sample
};
struct PllPipeSmem {
    PllBlockBuf b[2];
    int ja, jl;                      // first contradicted sweep flag / first latch of the block E just walked (its length if none)
    int j;                           // control's verdict on that block: its first contradicted sample, or its length
    int guess;                       // the sweep flag C assumes
    unsigned long long prof_c, prof_e, prof_l;
};

// `load(i, a, b)`: sample i of this call.  `emit(i, out, phase_before, freq_before)` and `emit_lock(i, lock)` (one thread per
// sample each) store the results of sample i; `emit` may see a sample more than once, the last time with its final values,
// `emit_lock` sees verified values only.  `s` lives in shared memory; all PP_THREADS threads call this with identical arguments.
// `pf` (optional, thread 0 only): cycle accounting [0] P alone (prologue, restarts) [1] C ‖ E [2] H ‖ P [3] C busy [4] E busy
// [5] blocks [6] contradicted blocks [7] control (with its barrier) [8] control, thread 0's work alone [9] lock-EMA thread busy.
template <class Load, class Emit, class EmitLock>
__device__ __forceinline__ void pll_run_blocks(PllState &s, const PllParams &p, unsigned long long n, unsigned long long abs0, PllPipeSmem &S,
                                               Load load, Emit emit, EmitLock emit_lock, unsigned long long *pf = nullptr)
{
    const int tid = threadIdx.x;
    const real_t avg_alpha = 0.00005;

    auto phase_P = [&](PllBlockBuf &B, unsigned long long i0, int cnt) {
        const int k = tid;
        if (k < cnt) {
            real_t a, b;
            load(i0 + k, a, b);
            B.pa[k] = a; B.pb[k] = b;
            B.sp[k] = arctan2_approx(b, a);                      // :128
            real_t nre = a, nim = b;                             // :193-218
            const real_t mag2 = nre * nre + nim * nim;
            const real_t inv  = q_rsqrt((float)mag2);
            nre *= inv; nim *= inv;
            B.nre[k] = nre; B.nim[k] = nim;
        }
    };

    int a = 0, cur = PP_B;
    unsigned long long iT = 0, iP = 0;
    int cntT = (int)((n < (unsigned long long)cur) ? n : (unsigned long long)cur), cntP = 0;
    bool have_prev = false;
    long long tq = clock64();
    if (tid == 0) { S.guess = (s.stage == 1) && pll_noise_like(s.avg_phase); S.prof_c = 0; S.prof_e = 0; S.prof_l = 0; }
    phase_P(S.b[a], iT, cntT);
    __syncthreads();
    if (pf && tid == 0) { const long long now = clock64(); pf[0] += now - tq; tq = now; }

    while (cntT > 0 || have_prev) {
        const bool acq = s.stage == 1;
        const bool guess = S.guess != 0;
        PllBlockBuf &T = S.b[a], &Pv = S.b[a ^ 1];
        // ---- C(T) ‖ E(prev) -----------------------------------------------------------------------------------------
        if (tid == 0) {
            if (cntT > 0) {
                const long long c0 = clock64();
                real_t phase = s.phase, freq = s.freq, sweep = s.sweep;
                const real_t alpha = s.alpha, beta = s.beta, maxf = s.max_freq, minf = s.min_freq;
                if (pll_fast_ok(s)) {
                    PllLoopConst kc; kc.alpha = alpha; kc.beta = beta; kc.max_freq = maxf; kc.min_freq = minf;
                    int k = 0;
                    for (; k + 4 <= cntT; k += 4) {              // the four sample phases are fetched ahead of the dependent chain
                        const real_t s0 = T.sp[k], s1 = T.sp[k + 1], s2 = T.sp[k + 2], s3 = T.sp[k + 3];
                        T.ph[k] = phase; T.fq[k] = freq; T.sw[k] = sweep;
                        pll_loop_fast(phase, freq, s0, kc); if (guess) pll_sweep_fast(freq, sweep, maxf, minf);
                        T.ph[k + 1] = phase; T.fq[k + 1] = freq; T.sw[k + 1] = sweep;
                        pll_loop_fast(phase, freq, s1, kc); if (guess) pll_sweep_fast(freq, sweep, maxf, minf);
                        T.ph[k + 2] = phase; T.fq[k + 2] = freq; T.sw[k + 2] = sweep;
                        pll_loop_fast(phase, freq, s2, kc); if (guess) pll_sweep_fast(freq, sweep, maxf, minf);
                        T.ph[k + 3] = phase; T.fq[k + 3] = freq; T.sw[k + 3] = sweep;
                        pll_loop_fast(phase, freq, s3, kc); if (guess) pll_sweep_fast(freq, sweep, maxf, minf);
                    }
                    for (; k < cntT; k++) {
                        T.ph[k] = phase; T.fq[k] = freq; T.sw[k] = sweep;
                        pll_loop_fast(phase, freq, T.sp[k], kc); if (guess) pll_sweep_fast(freq, sweep, maxf, minf);
                    }
                } else {
                    for (int k = 0; k < cntT; k++) {
                        T.ph[k] = phase; T.fq[k] = freq; T.sw[k] = sweep;
                        pll_loop_core(phase, freq, T.sp[k], alpha, beta, maxf, minf);
                        if (guess) pll_sweep_core(freq, sweep, maxf, minf);
                    }
                }
                s.phase = phase; s.freq = freq; s.sweep = sweep;         // the loop runs ahead of the EMAs; a contradiction restores it
                S.prof_c += (unsigned long long)(clock64() - c0);
            }
        } else if (tid == 32) {
            if (have_prev) {
                const long long e0 = clock64();
                real_t avg = s.avg_phase;
                int ja = cntP;
                if (acq) {
                    for (int k0 = 0; k0 < cntP && ja == cntP; k0 += 4) {
                        real_t v[4];
#pragma unroll
                        for (int q = 0; q < 4; q++) v[q] = Pv.at[(k0 + q < PP_B) ? k0 + q : PP_B - 1];
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const int k = k0 + q;
                            if (k < cntP && ja == cntP) {
                                avg = avg * (1.0 - avg_alpha) + v[q];    // :124
                                Pv.av[k] = avg;
                                if (pll_noise_like(avg) != guess) ja = k;
                            }
                        }
                    }
                } else {                                         // locked: nothing to verify, only the last value is needed
                    int k = 0;
                    for (; k + 4 <= cntP; k += 4) {
                        const real_t v0 = Pv.at[k], v1 = Pv.at[k + 1], v2 = Pv.at[k + 2], v3 = Pv.at[k + 3];
                        avg = avg * (1.0 - avg_alpha) + v0; avg = avg * (1.0 - avg_alpha) + v1;
                        avg = avg * (1.0 - avg_alpha) + v2; avg = avg * (1.0 - avg_alpha) + v3;
                    }
                    for (; k < cntP; k++) avg = avg * (1.0 - avg_alpha) + Pv.at[k];
                    Pv.av[cntP - 1] = avg;
                }
                S.ja = ja;
                S.prof_e += (unsigned long long)(clock64() - e0);
            }
        } else if (tid == 64) {
            if (have_prev) {
                const long long l0 = clock64();
                real_t lks = s.locksig;
                int jl = cntP;
                const auto keep = 1.0 - p.lock_alpha;            // the expression's own type (double in both builds), hoisted
                const real_t thresh = p.lock_thresh;
                if (acq) {
                    for (int k0 = 0; k0 < cntP && jl == cntP; k0 += 4) {
                        real_t v[4];
#pragma unroll
                        for (int q = 0; q < 4; q++) v[q] = Pv.lt[(k0 + q < PP_B) ? k0 + q : PP_B - 1];
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const int k = k0 + q;
                            if (k < cntP && jl == cntP) {
                                lks = lks * keep + v[q];         // :220
                                Pv.lk[k] = lks;
                                if (lks > thresh) jl = k;
                            }
                        }
                    }
                } else {
                    int k = 0;
                    for (; k + 4 <= cntP; k += 4) {
                        const real_t v0 = Pv.lt[k], v1 = Pv.lt[k + 1], v2 = Pv.lt[k + 2], v3 = Pv.lt[k + 3];
                        lks = lks * keep + v0; Pv.lk[k] = lks;
                        lks = lks * keep + v1; Pv.lk[k + 1] = lks;
                        lks = lks * keep + v2; Pv.lk[k + 2] = lks;
                        lks = lks * keep + v3; Pv.lk[k + 3] = lks;
                    }
                    for (; k < cntP; k++) { lks = lks * keep + Pv.lt[k]; Pv.lk[k] = lks; }
                }
                S.jl = jl;
                S.prof_l += (unsigned long long)(clock64() - l0);
            }
        }
        __syncthreads();
        if (pf && tid == 0) { const long long now = clock64(); pf[1] += now - tq; tq = now; }
        // ---- control: commit the previous block, or roll back to its first contradicted sample ----------------------------
        if (tid == 0) {
            int jj = cntP;
            if (have_prev) {
                jj = (S.ja < S.jl) ? S.ja : S.jl;
                if (jj < cntP) {                                 // sample jj again, with what it really saw
                    real_t phase = Pv.ph[jj], freq = Pv.fq[jj], sweep = Pv.sw[jj];
                    pll_loop_core(phase, freq, Pv.sp[jj], s.alpha, s.beta, s.max_freq, s.min_freq);
                    const bool flag = pll_noise_like(Pv.av[jj]);
                    if (flag) pll_sweep_core(freq, sweep, s.max_freq, s.min_freq);
                    s.phase = phase; s.freq = freq; s.sweep = sweep;
                    s.avg_phase = Pv.av[jj]; s.locksig = Pv.lk[jj];
                    if (Pv.lk[jj] > p.lock_thresh) pll_latch(s, p, abs0 + iP + jj);
                    S.guess = (s.stage == 1) && flag;
                } else {
                    s.avg_phase = Pv.av[cntP - 1]; s.locksig = Pv.lk[cntP - 1];
                }
            }
            S.j = jj;
            if (pf) pf[8] += clock64() - tq;
        }
        __syncthreads();
        const int j = S.j;                                       // (rewritten only behind the next C ‖ E barrier)
        const bool rolled = have_prev && j < cntP;
        if (pf && tid == 0) { pf[5] += have_prev; pf[6] += rolled; }
        if (have_prev && tid < (rolled ? j + 1 : cntP)) emit_lock(iP + tid, Pv.lk[tid]);    // the verified part of the previous block
        if (rolled) {                                            // drop block T, restart behind sample j of the previous block
            iT = iP + j + 1;
            cur = PP_B_MIN;
            cntT = (iT < n) ? (int)((n - iT < (unsigned long long)cur) ? (n - iT) : (unsigned long long)cur) : 0;
            have_prev = false;
            if (pf && tid == 0) { const long long now = clock64(); pf[7] += now - tq; tq = now; }
            phase_P(S.b[a], iT, cntT);
            __syncthreads();
            if (pf && tid == 0) { const long long now = clock64(); pf[0] += now - tq; tq = now; }
            continue;
        }
        if (pf && tid == 0) { const long long now = clock64(); pf[7] += now - tq; tq = now; }
        // ---- H(T), P(next) ------------------------------------------------------------------------------------------
        if (have_prev) cur = (cur * 2 < PP_B) ? cur * 2 : PP_B;
        const unsigned long long iU = iT + (unsigned long long)cntT;
        const int cntU = (cntT > 0 && iU < n) ? (int)((n - iU < (unsigned long long)cur) ? (n - iU) : (unsigned long long)cur) : 0;
        if (tid < cntT) {
            real_t ti, tr;
            const real_t ph = T.ph[tid];
            sincos_exact(ph, ti, tr);                            // :106-107
            const real_t x = T.pa[tid], y = T.pb[tid], nti = -ti;
            const real_t mre = x * tr - y * nti;                 // :110
            const real_t mim = x * nti + y * tr;
            T.at[tid] = avg_alpha * r_fabs(arctan2_approx(mim, mre));            // :117, :124
            T.lt[tid] = p.lock_alpha * (T.nre[tid] * tr + T.nim[tid] * ti);      // :220
            emit(iT + tid, mim, ph, T.fq[tid]);                  // :113
        }
        phase_P(Pv, iU, cntU);
        __syncthreads();
        if (pf && tid == 0) { const long long now = clock64(); pf[2] += now - tq; tq = now; }
        have_prev = cntT > 0; iP = iT; cntP = cntT;
        iT = iU; cntT = cntU; a ^= 1;
    }
    if (pf && tid == 0) { pf[3] += S.prof_c; pf[4] += S.prof_e; pf[9] += S.prof_l; }
}

} // namespace pdt
