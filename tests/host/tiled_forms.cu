// tests/host/tiled_forms.cu — HOST-side check (nvcc, CPU) of the tiled engine's building blocks (csrc/pdt_tiled.cuh, float build)
// against the reference-shaped functions of pdt_device.cuh, bit for bit:
//   fir_block26<L>, fir_block26_branches<L>  vs  fir_interp_exact (LowPassFilter.c:43-70, rotating summation order), L = 1…8
//   pll_track_run                            vs  pll_loop_core per sample (CarrierTrackingPLL.c:165-188)
//   agc_tile (common-regime chain + side proof + fallback)  vs  agc_step per sample (AGC.c:98-131), incl. signals that break the proof
// Prints "OK <cases>".
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>
#include "pdt_tiled.cuh"

using namespace pdt;
using namespace pdt::tiled;

static bool same(float a, float b) { return std::memcmp(&a, &b, sizeof a) == 0; }
static std::mt19937_64 rng(777);
static std::uniform_real_distribution<double> U(-1.0, 1.0);

template <int L>
static int fir_case(unsigned long long &cases)
{
    const int N = FIR_K * L, blocks = 24, n = FIR_K * blocks;
    std::vector<float> h(N), x(FIR_K + n, 0.0f);                         // FIR_K zeros of history in front
    for (auto &v : h) v = (float)(U(rng) * 0.1);
    for (int i = 0; i < n; i++) x[FIR_K + i] = (float)U(rng);
    TapsRev t; std::memset(&t, 0, sizeof t);
    for (int u = 0; u < N; u++) t.hr[u] = h[N - 1 - u];                  // as pdt_batch.cu::tiled_setup fills it
    for (int b = 0; b < blocks; b++) {
        float prev[FIR_K], cur[FIR_K], y1[FIR_K * L], y2[FIR_K * L];
        for (int s = 0; s < FIR_K; s++) { cur[s] = x[FIR_K + b * FIR_K + s]; prev[s] = x[b * FIR_K + s]; }
        fir_block26<L>(prev, cur, t, [&](int o, float v) { y1[o] = v; });
        fir_block26_branches<L>(prev, cur, t, [&](int o, float v) { y2[o] = v; });
        for (int c = 0; c < FIR_K; c++) for (int p = 0; p < L; p++) {
            const int j = b * FIR_K + c;                                  // absolute input index, blocks aligned to multiples of 26
            const float want = fir_interp_exact(h.data(), x.data() + FIR_K + j, N, L, FIR_K, p, j % FIR_K);
            if (!same(want, y1[c * L + p]) || !same(want, y2[c * L + p])) { std::printf("MISMATCH fir L=%d block %d c=%d p=%d\n", L, b, c, p); return 1; }
            cases++;
        }
    }
    return 0;
}

int main()
{
    unsigned long long cases = 0;
    if (fir_case<1>(cases) || fir_case<2>(cases) || fir_case<3>(cases) || fir_case<4>(cases) || fir_case<5>(cases) || fir_case<6>(cases) ||
        fir_case<7>(cases) || fir_case<8>(cases)) return 1;

    // ---- track-mode PLL over a stream, from misaligned starts and with ragged ends ----
    {
        const int n = 40000;
        alignas(16) static float sp[40000], ph[40000];
        for (int i = 0; i < n; i++) sp[i] = (float)(U(rng) * 3.14159);
        TrackConst k; k.alpha = 0.0026f; k.beta = 3.4e-6f; k.max_freq = 0.565f; k.min_freq = -0.565f;
        for (int trial = 0; trial < 40; trial++) {
            const unsigned long long i0 = (unsigned long long)((U(rng) + 1) * 500), i1 = n - (unsigned long long)((U(rng) + 1) * 500);
            float p1 = (float)(U(rng) * 6.2), f1 = (float)(U(rng) * 0.5), p2 = p1, f2 = f1;
            pll_track_run<true>(sp, ph, i0, i1, p2, f2, k);
            for (unsigned long long i = i0; i < i1; i++) {
                if (!same(ph[i], p1)) { std::printf("MISMATCH track phase at %llu (trial %d)\n", i, trial); return 1; }
                pll_loop_core(p1, f1, sp[i], k.alpha, k.beta, k.max_freq, k.min_freq);
                cases++;
            }
            if (!same(p1, p2) || !same(f1, f2)) { std::printf("MISMATCH track end state (trial %d)\n", trial); return 1; }
        }
    }
    // ---- AGC tiles: quiet signal (proof holds), bursts (attack branch), tiny and huge signals (clamps) ----
    {
        const int n = 60000;
        alignas(16) static float x[60000], z[60000], w[60000];
        bool proof_failed_somewhere = false, proof_held_somewhere = false;
        for (int sc = 0; sc < 5; sc++) {
            for (int i = 0; i < n; i++) {
                double amp = 1.0;
                if (sc == 1 && (i / 7000) % 2) amp = 30.0;
                if (sc == 2) amp = 1e-7;
                if (sc == 3) amp = 3e4;
                x[i] = (float)(amp * U(rng));
            }
            const float attack = 0.01f, decay = (sc == 2) ? 0.3f : 0.0001f;
            for (int trial = 0; trial < 12; trial++) {
                const unsigned long long warm = (unsigned long long)((U(rng) + 1) * 3000), begin = warm + (trial % 3 ? (unsigned long long)((U(rng) + 1) * 4000) : 0),
                                         end = begin + 4 + (unsigned long long)((U(rng) + 1) * 9000);
                float g = (sc == 2) ? 4990.0f : (float)(1.0 + U(rng) * 0.5), gs = 0;
                AgcState st; st.init = 1; st.gain = g;
                agc_tile(x, z, warm, begin, end, g, gs, attack, decay);
                bool clamped_or_attacked = false;
                for (unsigned long long i = warm; i < end; i++) {
                    if (i == begin && !same(st.gain, gs)) { std::printf("MISMATCH agc start gain (scenario %d trial %d)\n", sc, trial); return 1; }
                    const float gin = st.gain;
                    w[i] = agc_step(st, x[i], attack, decay);
                    clamped_or_attacked |= (fabsf(fabsf(x[i] * gin) - 1.0f) > gin) || st.gain == 5000.0f || st.gain == 10e-5f;
                    if (i >= begin && !same(w[i], z[i])) { std::printf("MISMATCH agc output at %llu (scenario %d trial %d)\n", i, sc, trial); return 1; }
                    cases++;
                }
                if (!same(st.gain, g)) { std::printf("MISMATCH agc end gain (scenario %d trial %d)\n", sc, trial); return 1; }
                proof_failed_somewhere |= clamped_or_attacked; proof_held_somewhere |= !clamped_or_attacked;
            }
        }
        if (!proof_failed_somewhere || !proof_held_somewhere) { std::printf("AGC scenarios did not cover both regimes\n"); return 1; }
    }
    std::printf("OK %llu\n", cases);
    return 0;
}
