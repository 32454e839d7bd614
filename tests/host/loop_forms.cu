// tests/host/loop_forms.cu — HOST-side check (compiled by nvcc, run on the CPU) that the branch-free forms the kernels use
// (pdt_pll_pipe.cuh: pll_loop_fast, pll_sweep_fast; pdt_device.cuh: agc_step4) give bit-identical results to the reference-shaped
// functions (pdt_device.cuh: pll_loop_core, pll_sweep_core = CarrierTrackingPLL.c:165-188, :233-246 wherever pll_fast_ok() holds;
// agc_step = AGC.c:98-131 always).
// Built twice: -DPDT_USE_FLOATS=1 (float-only 2π wraps of pdt_tiled.cuh) and =0 (double selects).  Prints "OK <cases>".
#include <cstdio>
#include <cstring>
#include <cmath>
#include <random>
#include "pdt_pll_pipe.cuh"

using namespace pdt;

static bool same(real_t a, real_t b) { return std::memcmp(&a, &b, sizeof a) == 0; }

int main()
{
    std::mt19937_64 rng(20261017);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    unsigned long long cases = 0;
    const double fss[] = {5000.0, 18750.0, 50000.0, 250000.0, 2000000.0};
    const double bws[] = {0.0005, 0.002, 0.02, 0.1, 0.3};            // per-sample loop bandwidths (acquisition and track)
    for (double fs : fss) for (double bw : bws) {
        PllState s; pll_reset(s);
        PllParams p; p.Fs = (real_t)fs; p.freq_range = 4500; p.lock_thresh = (real_t)0.1; p.lock_alpha = (real_t)0.002;
        p.bw_acq = (real_t)bw; p.bw_track = (real_t)(bw / 10);
        pll_begin(s, p);
        if (!pll_fast_ok(s)) continue;                                  // the kernels take the reference-shaped path there
        PllLoopConst k; k.alpha = s.alpha; k.beta = s.beta; k.max_freq = s.max_freq; k.min_freq = s.min_freq;
        // (1) independent random states, with the edges over-represented
        for (int i = 0; i < 400000; i++) {
            double ph = (U(rng) * 2 - 1) * 2 * PDT_PI, sp = (U(rng) * 2 - 1) * PDT_PI, fq = (U(rng) * 2 - 1) * 1.2 * (double)s.max_freq;
            const int e = (int)(U(rng) * 16);
            if (e == 0) ph = 2 * PDT_PI; else if (e == 1) ph = -2 * PDT_PI; else if (e == 2) sp = PDT_PI; else if (e == 3) sp = -PDT_PI;
            else if (e == 4) ph = sp - PDT_PI; else if (e == 5) ph = sp + PDT_PI; else if (e == 6) fq = s.max_freq; else if (e == 7) fq = s.min_freq;
            else if (e == 8) { ph = 2 * PDT_PI - U(rng) * 1e-6; sp = -PDT_PI + U(rng) * 1e-6; }
            real_t p1 = (real_t)ph, f1 = (real_t)fq, p2 = p1, f2 = f1;
            pll_loop_core(p1, f1, (real_t)sp, s.alpha, s.beta, s.max_freq, s.min_freq);
            pll_loop_fast(p2, f2, (real_t)sp, k);
            if (!same(p1, p2) || !same(f1, f2)) {
                std::printf("MISMATCH loop fs=%g bw=%g ph=%.17g fq=%.17g sp=%.17g -> core (%.17g, %.17g) fast (%.17g, %.17g)\n", fs, bw, ph, fq, sp,
                            (double)p1, (double)f1, (double)p2, (double)f2);
                return 1;
            }
            real_t s1 = (real_t)((U(rng) < 0.5 ? -1 : 1) * 0.2 * (2.0 * PDT_PI / fs)), s2 = s1;
            real_t g1 = f1, g2 = f1;
            if (e == 9) g1 = g2 = s.max_freq - s1; else if (e == 10) g1 = g2 = s.min_freq - s1; else if (e == 11) g1 = g2 = -s1;
            pll_sweep_core(g1, s1, s.max_freq, s.min_freq);
            pll_sweep_fast(g2, s2, s.max_freq, s.min_freq);
            if (!same(g1, g2) || !same(s1, s2)) {
                std::printf("MISMATCH sweep fs=%g freq=%.17g -> core (%.17g, %.17g) fast (%.17g, %.17g)\n", fs, (double)f1, (double)g1, (double)s1,
                            (double)g2, (double)s2);
                return 1;
            }
            cases += 2;
        }
        // (2) a trajectory: the loop run on noise + a tone with the sweep on, state fed back
        real_t p1 = (real_t)0.1, f1 = 0, w1 = s.sweep, p2 = p1, f2 = f1, w2 = w1;
        double tone = 0.0;
        for (int i = 0; i < 300000; i++) {
            tone += 2 * PDT_PI * 1234.5 / fs;
            const double x = std::cos(tone) + 0.7 * (U(rng) - 0.5), y = std::sin(tone) + 0.7 * (U(rng) - 0.5);
            const real_t sp = arctan2_approx((real_t)y, (real_t)x);
            pll_loop_core(p1, f1, sp, s.alpha, s.beta, s.max_freq, s.min_freq);
            pll_loop_fast(p2, f2, sp, k);
            if ((i / 5000) % 2 == 0) { pll_sweep_core(f1, w1, s.max_freq, s.min_freq); pll_sweep_fast(f2, w2, s.max_freq, s.min_freq); }
            if (!same(p1, p2) || !same(f1, f2) || !same(w1, w2)) { std::printf("MISMATCH trajectory fs=%g bw=%g at %d\n", fs, bw, i); return 1; }
            cases++;
        }
    }
    // (3) AGC: groups of four through the common-regime chain (agc_step4) against four general steps (agc_step = AGC.c:98-131),
    //     on signals that sit in the decay regime, that hit the attack branch (bursts) and that hit both clamps
    bool hit_max = false, hit_min = false, hit_attack = false;
    for (int sc = 0; sc < 6; sc++) {
        AgcState a1, a2; a1.init = a2.init = 1; a1.gain = a2.gain = (real_t)(sc == 4 ? 4999.0 : (sc == 5 ? 1e-4 : 1.0));
        const real_t attack = (real_t)(sc == 3 ? 0.5 : 0.01), decay = (real_t)((sc == 2 || sc == 4) ? 0.3 : 0.0001);
        for (int i = 0; i < 500000; i++) {
            real_t x[4], w[4], v[4];
            for (int q = 0; q < 4; q++) {
                double amp = (sc == 1 && (i / 3000) % 2) ? 40.0 : 1.0;                       // bursts: |x·g| - 1 > g -> attack branch
                if (sc == 4) amp = 1e-7;                                                        // gain climbs into the 5000 clamp
                if (sc == 5) amp = 1e5 * U(rng);                                                // gain pushed below zero -> 10e-5
                x[q] = (real_t)(amp * (U(rng) * 2 - 1));
            }
            for (int q = 0; q < 4; q++) {
                hit_attack |= std::fabs(std::fabs((double)x[q] * (double)a1.gain) - 1.0) > (double)a1.gain;
                w[q] = agc_step(a1, x[q], attack, decay);
                hit_max |= a1.gain == (real_t)5000; hit_min |= a1.gain == (real_t)10e-5;
            }
            agc_step4(a2, x[0], x[1], x[2], x[3], attack, decay, v[0], v[1], v[2], v[3]);
            for (int q = 0; q < 4; q++)
                if (!same(w[q], v[q]) || !same(a1.gain, a2.gain)) { std::printf("MISMATCH agc scenario %d group %d\n", sc, i); return 1; }
            cases += 4;
        }
    }
    if (!hit_max || !hit_min || !hit_attack) { std::printf("AGC scenarios did not reach every branch: max %d min %d attack %d\n", hit_max, hit_min, hit_attack); return 1; }
    std::printf("OK %llu\n", cases);
    return 0;
}
