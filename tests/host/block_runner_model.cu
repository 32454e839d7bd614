// tests/host/block_runner_model.cu — HOST MODEL of the PLL block runner (csrc/pdt_pll_pipe.cuh::pll_run_blocks): the same
// schedule — C(block k) speculating the sweep flag and the open latch while E(block k-1) is still unverified, commit or roll
// back to the first contradicted sample, short blocks doubling after a roll-back — executed phase by phase on the CPU with the
// same per-sample functions, and compared with the one-thread loop (pll_step) over a capture: every output, lock value,
// phase / frequency trace and the final state must be bit-identical, however the calls are cut.  It pins the ALGORITHM
// (samples up to the first contradiction are exact; what is restored and repeated) under the CPU suite; the device code is
// held against the oracle by the GPU suite.
//   block_runner_model <fs> <iq.bin: interleaved real_t> <bw_acq> <bw_track> <lock_thresh> [cut lengths…]
//   -> "OK samples=<n> blocks=<b> rolled=<r> locked=<0|1>"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "pdt_pll_pipe.cuh"

using namespace pdt;

static bool same(real_t a, real_t b) { return std::memcmp(&a, &b, sizeof a) == 0; }

struct Out { std::vector<real_t> out, lock, ph, fq; };

// one call of the runner over iq[0 … n): the phases of pll_run_blocks in their order, all "threads" of a phase in a loop
static void run_blocks_model(PllState &s, const PllParams &p, const real_t *iq, unsigned long long n, unsigned long long abs0, Out &o,
                             unsigned long long &blocks, unsigned long long &rolled_n)
{
    const real_t avg_alpha = 0.00005;
    static PllBlockBuf B[2];
    auto phase_P = [&](PllBlockBuf &b, unsigned long long i0, int cnt) {
        for (int k = 0; k < cnt; k++) {
            const real_t a = iq[2 * (i0 + k)], q = iq[2 * (i0 + k) + 1];
            b.pa[k] = a; b.pb[k] = q; b.sp[k] = arctan2_approx(q, a);
            real_t nre = a, nim = q;
            const real_t mag2 = nre * nre + nim * nim;
            const real_t inv = q_rsqrt((float)mag2);
            nre *= inv; nim *= inv;
            b.nre[k] = nre; b.nim[k] = nim;
        }
    };
    int a = 0, cur = PP_B;
    unsigned long long iT = 0, iP = 0;
    int cntT = (int)((n < (unsigned long long)cur) ? n : (unsigned long long)cur), cntP = 0;
    bool have_prev = false;
    bool guess = (s.stage == 1) && pll_noise_like(s.avg_phase);
    phase_P(B[a], iT, cntT);
    while (cntT > 0 || have_prev) {
        const bool acq = s.stage == 1;
        PllBlockBuf &T = B[a], &Pv = B[a ^ 1];
        int ja = cntP, jl = cntP;
        // ---- C(T) ‖ E(prev): on the device three threads at once; none reads what another writes in this phase ----
        real_t avg0 = s.avg_phase, lks0 = s.locksig;                 // E starts from the committed EMA state
        if (cntT > 0) {
            real_t phase = s.phase, freq = s.freq, sweep = s.sweep;
            const bool fast = pll_fast_ok(s);
            PllLoopConst kc; kc.alpha = s.alpha; kc.beta = s.beta; kc.max_freq = s.max_freq; kc.min_freq = s.min_freq;
            for (int k = 0; k < cntT; k++) {
                T.ph[k] = phase; T.fq[k] = freq; T.sw[k] = sweep;
                if (fast) { pll_loop_fast(phase, freq, T.sp[k], kc); if (guess) pll_sweep_fast(freq, sweep, s.max_freq, s.min_freq); }
                else      { pll_loop_core(phase, freq, T.sp[k], s.alpha, s.beta, s.max_freq, s.min_freq); if (guess) pll_sweep_core(freq, sweep, s.max_freq, s.min_freq); }
            }
            s.phase = phase; s.freq = freq; s.sweep = sweep;
        }
        if (have_prev) {
            real_t avg = avg0;
            for (int k = 0; k < cntP; k++) {
                avg = avg * (1.0 - avg_alpha) + Pv.at[k];
                Pv.av[k] = avg;
                if (acq && pll_noise_like(avg) != guess) { ja = k; break; }
            }
            real_t lks = lks0;
            const auto keep = 1.0 - p.lock_alpha;
            for (int k = 0; k < cntP; k++) {
                lks = lks * keep + Pv.lt[k];
                Pv.lk[k] = lks;
                if (acq && lks > p.lock_thresh) { jl = k; break; }
            }
        }
        // ---- control ----
        int j = cntP;
        if (have_prev) {
            j = ja < jl ? ja : jl;
            if (j < cntP) {
                real_t phase = Pv.ph[j], freq = Pv.fq[j], sweep = Pv.sw[j];
                pll_loop_core(phase, freq, Pv.sp[j], s.alpha, s.beta, s.max_freq, s.min_freq);
                const bool flag = pll_noise_like(Pv.av[j]);
                if (flag) pll_sweep_core(freq, sweep, s.max_freq, s.min_freq);
                s.phase = phase; s.freq = freq; s.sweep = sweep;
                s.avg_phase = Pv.av[j]; s.locksig = Pv.lk[j];
                if (Pv.lk[j] > p.lock_thresh) pll_latch(s, p, abs0 + iP + j);
                guess = (s.stage == 1) && flag;
            } else { s.avg_phase = Pv.av[cntP - 1]; s.locksig = Pv.lk[cntP - 1]; }
            blocks++;
        }
        const bool rolled = have_prev && j < cntP;
        if (have_prev) for (int k = 0; k < (rolled ? j + 1 : cntP); k++) o.lock[abs0 + iP + k] = Pv.lk[k];      // emit_lock: verified values only
        if (rolled) {
            rolled_n++;
            iT = iP + j + 1; cur = PP_B_MIN;
            cntT = (iT < n) ? (int)((n - iT < (unsigned long long)cur) ? (n - iT) : (unsigned long long)cur) : 0;
            have_prev = false;
            phase_P(B[a], iT, cntT);
            continue;
        }
        // ---- H(T), P(next) ----
        if (have_prev) cur = (cur * 2 < PP_B) ? cur * 2 : PP_B;
        const unsigned long long iU = iT + (unsigned long long)cntT;
        const int cntU = (cntT > 0 && iU < n) ? (int)((n - iU < (unsigned long long)cur) ? (n - iU) : (unsigned long long)cur) : 0;
        for (int k = 0; k < cntT; k++) {
            real_t ti, tr;
            sincos_exact(T.ph[k], ti, tr);
            const real_t x = T.pa[k], y = T.pb[k], nti = -ti;
            const real_t mre = x * tr - y * nti, mim = x * nti + y * tr;
            T.at[k] = avg_alpha * r_fabs(arctan2_approx(mim, mre));
            T.lt[k] = p.lock_alpha * (T.nre[k] * tr + T.nim[k] * ti);
            o.out[abs0 + iT + k] = mim; o.ph[abs0 + iT + k] = T.ph[k]; o.fq[abs0 + iT + k] = T.fq[k];       // emit: may be overwritten later
        }
        phase_P(Pv, iU, cntU);
        have_prev = cntT > 0; iP = iT; cntP = cntT;
        iT = iU; cntT = cntU; a ^= 1;
    }
}

int main(int argc, char **argv)
{
    if (argc < 6) { std::fprintf(stderr, "usage: block_runner_model <fs> <iq.bin> <bw_acq> <bw_track> <lock_thresh> [cuts…]\n"); return 2; }
    std::vector<real_t> iq;
    {
        FILE *f = std::fopen(argv[2], "rb");
        if (!f) { std::perror(argv[2]); return 2; }
        std::fseek(f, 0, SEEK_END); const long bytes = std::ftell(f); std::fseek(f, 0, SEEK_SET);
        iq.resize((size_t)bytes / sizeof(real_t));
        if (std::fread(iq.data(), sizeof(real_t), iq.size(), f) != iq.size()) return 2;
        std::fclose(f);
    }
    const unsigned long long n = iq.size() / 2;
    PllParams p; p.Fs = (real_t)std::atof(argv[1]); p.freq_range = 4500; p.lock_thresh = (real_t)std::atof(argv[5]); p.lock_alpha = (real_t)0.002;
    p.bw_acq = (real_t)std::atof(argv[3]); p.bw_track = (real_t)std::atof(argv[4]);
    std::vector<unsigned long long> cuts;
    for (int i = 6; i < argc; i++) cuts.push_back((unsigned long long)std::atoll(argv[i]));

    // the one-thread loop
    Out ref; ref.out.resize(n); ref.lock.resize(n); ref.ph.resize(n); ref.fq.resize(n);
    PllState s1; pll_reset(s1); pll_begin(s1, p);
    for (unsigned long long i = 0; i < n; i++) { ref.ph[i] = s1.phase; ref.fq[i] = s1.freq; pll_step(s1, p, iq[2 * i], iq[2 * i + 1], ref.out[i], ref.lock[i], i); }

    // the block runner, call after call
    Out got; got.out.assign(n, 0); got.lock.assign(n, 0); got.ph.assign(n, 0); got.fq.assign(n, 0);
    PllState s2; pll_reset(s2);
    unsigned long long blocks = 0, rolled = 0, at = 0; size_t ci = 0;
    while (at < n) {
        unsigned long long len = ci < cuts.size() ? cuts[ci++] : n - at;
        if (len > n - at) len = n - at;
        pll_begin(s2, p);
        run_blocks_model(s2, p, iq.data() + 2 * at, len, at, got, blocks, rolled);
        at += len;
    }
    for (unsigned long long i = 0; i < n; i++)
        if (!same(ref.out[i], got.out[i]) || !same(ref.lock[i], got.lock[i]) || !same(ref.ph[i], got.ph[i]) || !same(ref.fq[i], got.fq[i])) {
            std::printf("MISMATCH at sample %llu: out %.9g/%.9g lock %.9g/%.9g phase %.9g/%.9g freq %.9g/%.9g\n", i, (double)ref.out[i], (double)got.out[i],
                        (double)ref.lock[i], (double)got.lock[i], (double)ref.ph[i], (double)got.ph[i], (double)ref.fq[i], (double)got.fq[i]);
            return 1;
        }
    if (!same(s1.phase, s2.phase) || !same(s1.freq, s2.freq) || !same(s1.sweep, s2.sweep) || !same(s1.avg_phase, s2.avg_phase) ||
        !same(s1.locksig, s2.locksig) || s1.stage != s2.stage || s1.lock_sample != s2.lock_sample || !same(s1.alpha, s2.alpha)) {
        std::printf("MISMATCH in the final state\n");
        return 1;
    }
    std::printf("OK samples=%llu blocks=%llu rolled=%llu locked=%d\n", n, blocks, rolled, s1.stage == 2);
    return 0;
}
