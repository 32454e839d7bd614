// tools/acq_study.cu — HOST-ONLY study (runs on the CPU): how long does the acquisition-mode PLL recurrence take to
// forget its start state?  For a capture that has not latched after R samples, the (phase, freq, sweep) recurrence is
// restarted at later points from the state it had at R and run forward on the true inputs; the merge time is the
// number of samples until the restarted trajectory is bit-identical to the true one.
// build: nvcc -O2 -fmad=false -Xcompiler -ffp-contract=off -I../project-desert-tortoise_b200/csrc -I../include -o acq_study acq_study.cu
// usage: acq_study capture.cf32 Fs R stride maxwarm
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include "pdt_tiled.cuh"
using namespace pdt; using namespace pdt::tiled;

static void acq_step_h(float &phase, float &freq, float &sweep, float sp, bool on, const TrackConst &k)
{
    pll_track_step(phase, freq, sp, k);
    const float f2 = freq + sweep;
    float s2 = (f2 >= 0) ? fabsf(sweep) : -fabsf(sweep);
    s2 = (f2 <= k.min_freq) ? -sweep : s2;
    s2 = (f2 >= k.max_freq) ? -sweep : s2;
    freq = on ? f2 : freq;
    sweep = on ? s2 : sweep;
}

int main(int argc, char **argv)
{
    if (argc < 6) { fprintf(stderr, "usage\n"); return 2; }
    const double Fs = atof(argv[2]);
    const long R = atol(argv[3]), stride = atol(argv[4]), maxwarm = atol(argv[5]);
    FILE *f = fopen(argv[1], "rb"); if (!f) { perror("open"); return 2; }
    fseek(f, 0, SEEK_END); const long n = ftell(f) / 8; fseek(f, 0, SEEK_SET);
    std::vector<float> iq(2 * n); if (fread(iq.data(), 8, n, f) != (size_t)n) return 2; fclose(f);
    PllParams pp; const double w = 2.0 * M_PI / (float)Fs;
    pp.Fs = (float)Fs; pp.freq_range = 4500.0; pp.lock_thresh = 0.08; pp.lock_alpha = 0.3979 * w; pp.bw_acq = 127.3240 * w; pp.bw_track = 10.3451 * w;
    PllState ps; pll_reset(ps); pll_begin(ps, pp);
    TrackConst k; k.alpha = ps.alpha; k.beta = ps.beta; k.max_freq = ps.max_freq; k.min_freq = ps.min_freq;
    std::vector<float> ph(n + 1), fr(n + 1), sw(n + 1), sp(n); std::vector<unsigned char> nl(n);
    long lock = -1;
    for (long i = 0; i < n; i++) {
        ph[i] = ps.phase; fr[i] = ps.freq; sw[i] = ps.sweep;
        sp[i] = arctan2_approx(iq[2 * i + 1], iq[2 * i]);
        const float f_before = ps.freq;
        float out, lk;
        pll_step(ps, pp, iq[2 * i], iq[2 * i + 1], out, lk, i);
        nl[i] = (double)fabsf((float)(PDT_PI / 2.0 - (double)ps.avg_phase)) < 0.05;
        if (ps.stage == 2) { lock = i; break; }
        (void)f_before;
    }
    const long m = lock >= 0 ? lock : n;
    ph[m] = ps.phase; fr[m] = ps.freq; sw[m] = ps.sweep;
    // self-check of the restated step against the full step
    { float p = ph[0], q = fr[0], s = sw[0]; long bad = 0;
      for (long i = 0; i < m; i++) { if (p != ph[i] || q != fr[i] || s != sw[i]) { bad++; if (bad < 3) printf("selfcheck mismatch at %ld\n", i); p = ph[i]; q = fr[i]; s = sw[i]; } acq_step_h(p, q, s, sp[i], nl[i], k); }
      printf("n=%ld lock=%ld selfcheck_bad=%ld  flags1=%.4f  freq@R=%.1f Hz\n", n, lock, bad, (double)std::count(nl.begin(), nl.begin() + m, 1) / m, R < m ? fr[R] * Fs / (2 * M_PI) : 0.0); }
    if (R >= m) { printf("latched before R\n"); return 0; }
    std::vector<long> merge;
    long fails = 0, total = 0;
    for (long b = R + maxwarm; b + 1 < m; b += stride) {
        float p = ph[R], q = fr[R], s = sw[R];
        long t = -1;
        for (long i = b - maxwarm; i < b; i++) {
            if (p == ph[i] && q == fr[i] && s == sw[i]) { t = i - (b - maxwarm); break; }
            acq_step_h(p, q, s, sp[i], nl[i], k);
        }
        total++;
        if (t < 0) fails++; else merge.push_back(t);
    }
    std::sort(merge.begin(), merge.end());
    auto pct = [&](double x) { return merge.empty() ? -1L : merge[(size_t)(x * (merge.size() - 1))]; };
    printf("restarts=%ld not merged within %ld: %ld   merge time p50=%ld p90=%ld p99=%ld max=%ld\n", total, maxwarm, fails, pct(0.5), pct(0.9), pct(0.99), pct(1.0));
    return 0;
}
