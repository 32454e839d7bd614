// pdt_device.cuh — device-side stage state and EXACT-ORDER stage arithmetic (sm_100a).
//
// Every function here reproduces the floating-point operation order of the reference stage it cites
// (file:line relative to the reference repo), so that with -fmad=false the float build is bit-identical
// to the reference compiled with -O2 -ffp-contract=off on x86-64.  The time-parallel reformulations of
// the recurrences live in pdt_tiled.cuh / pdt_tiled_kernels.cuh and are validated against these.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#ifndef PDT_USE_FLOATS
#define PDT_USE_FLOATS 1
#endif
#if PDT_USE_FLOATS
typedef float real_t;
#else
typedef double real_t;
#endif

#define PDT_PI 3.14159265358979323846
#define PDT_MAX_TAPS 1024
#define PDT_MAX_PHASE_TAPS 64   // taps per polyphase branch (N / L); the reference uses 26 (main.c:104) or 50

// Every stage function is __host__ __device__: the device build is the product; the host build of the very same
// code is used only by host-side study tools (tools/acq_study.cu: CPU emulation of the recurrences).
#define PDT_DEV __host__ __device__ __forceinline__

#include <cmath>
#include <cstring>
PDT_DEV uint32_t pdt_f2u(float x)
{
#ifdef __CUDA_ARCH__
    return __float_as_uint(x);
#else
    uint32_t u; memcpy(&u, &x, 4); return u;
#endif
}
PDT_DEV float pdt_u2f(uint32_t u)
{
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float x; memcpy(&x, &u, 4); return x;
#endif
}
PDT_DEV int pdt_d2i_rz(double x)
{
#ifdef __CUDA_ARCH__
    return __double2int_rz(x);
#else
    return (int)x;
#endif
}

// ---------------------------------------------------------------------------------------------------
// small overload helpers
// ---------------------------------------------------------------------------------------------------
PDT_DEV float  r_fabs(float x)  { return fabsf(x); }
PDT_DEV double r_fabs(double x) { return fabs(x); }
PDT_DEV float  r_rint(float x)  { return rintf(x); }
PDT_DEV double r_rint(double x) { return rint(x); }

// ---------------------------------------------------------------------------------------------------
// sinf/cosf that match glibc >= 2.28 bit for bit (float build).
// glibc's sinf/cosf (sysdeps/ieee754/flt-32/s_sincosf.h, from ARM Optimized Routines) evaluate a short
// double-precision polynomial after a fast quadrant reduction and round once to float.  The algorithm
// and its published constants are restated here; double arithmetic on the GPU is IEEE, so the result
// equals the host libm's for |x| < 120 (the PLL phase lives in [-2π, 2π]).  tests/test_gpu_parity.py
// checks this against the oracle over the whole phase range.
// ---------------------------------------------------------------------------------------------------
// sine / cosine kernels on the reduced argument (s_sincosf.h: sincosf_poly).  Both are evaluated for every sample and
// swapped / negated afterwards: IEEE +,× are sign-symmetric, so sin_poly(-x) == -sin_poly(x) and a cosine evaluated with
// the negated coefficient set (glibc's second table entry) is the negated cosine, bit for bit — no per-lane branches.
#ifdef __CUDACC__
// the polynomial / reduction constants live in the constant bank: a DMUL/DADD takes them as c[bank][offset] operands,
// where immediates would be rebuilt in uniform registers (two UMOVs per constant) for every sample
__constant__ double pdt_sc_tab[10] = {
    -0x1.555545995a603p-3, 0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13,                                  // s1 s2 s3
    -0x1.ffffffd0c621cp-2, 0x1.55553e1068f19p-5, -0x1.6c087e89a359dp-10, 0x1.99343027bf8c3p-16,           // c1 c2 c3 c4
    0x1.45F306DC9C883p+23, 0x1.921FB54442D18p0, 0x1p0 };                                                  // 2^24·2/π, π/2, 1
#endif
#ifdef __CUDA_ARCH__
#define PDT_SC(i, lit) pdt_sc_tab[i]
#else
#define PDT_SC(i, lit) (lit)
#endif

PDT_DEV double sc_sin_poly(double x, double x2)
{
    const double s1 = PDT_SC(0, -0x1.555545995a603p-3), s2 = PDT_SC(1, 0x1.1107605230bc4p-7), s3 = PDT_SC(2, -0x1.994eb3774cf24p-13);
    const double x3 = x * x2;
    const double t1 = s2 + x2 * s3;
    const double x7 = x3 * x2;
    const double s  = x + x3 * s1;
    return s + x7 * t1;
}
PDT_DEV double sc_cos_poly(double x2)
{
    const double c0 = PDT_SC(9, 0x1p0), c1 = PDT_SC(3, -0x1.ffffffd0c621cp-2), c2 = PDT_SC(4, 0x1.55553e1068f19p-5),
                 c3 = PDT_SC(5, -0x1.6c087e89a359dp-10), c4 = PDT_SC(6, 0x1.99343027bf8c3p-16);
    const double x4 = x2 * x2;
    const double t2 = c3 + x2 * c4;
    const double t1 = c0 + x2 * c1;
    const double x6 = x4 * x2;
    const double c  = t1 + x4 * c2;
    return c + x6 * t2;
}

PDT_DEV uint32_t abstop12(float x) { return (pdt_f2u(x) >> 20) & 0x7ffu; }

// the branch-free body, valid for |y| < 120 and y != -0 (sincos_in_core_range): what the hot loops call directly, so that
// several evaluations interleave in one straight-line block (a per-sample range branch fences the scheduler off).
//   * glibc's three ranges collapse into its general path: for |y| < π/4 the quadrant is n = 0 and the reduction x - 0·(π/2)
//     is exact, which is its small-argument path; and for |y| < 2^-12, where glibc returns (y, 1) without evaluating
//     anything, the polynomials round to exactly that: x·(1 - x²/6 …) is within 2^-26.5 of x, below half a float ulp, and
//     1 - x²/2 … is less than 2^-25 below 1, half the float spacing there.  The one exception is y = -0 (the sine kernel
//     gives +0), which the range predicate therefore excludes.  tests/test_sincos_exact.py checks a numpy restatement of
//     this function against glibc over every float of [2^-13, 8] and samples below.
//   * signs are applied AFTER the conversion to float (rounding to nearest is symmetric: (float)(-d) == -(float)d) as one
//     XOR each with a mask taken straight from the quadrant bits — no double negations, no 64-bit selects.
PDT_DEV bool sincos_in_core_range(float y) { return (r_fabs(y) < 120.0f) && (pdt_f2u(y) != 0x80000000u); }

PDT_DEV void sincos_core(float y, float &s, float &c)
{
    const double x = (double)y;
    const double hpi_inv = PDT_SC(7, 0x1.45F306DC9C883p+23), hpi = PDT_SC(8, 0x1.921FB54442D18p0);
    const double r = x * hpi_inv;
    const int t1 = pdt_d2i_rz(r) + 0x800000;                             // quadrant n = t1 >> 24 (arithmetic): bits 24, 25 = n & 3
    const int n = t1 >> 24;
    const double xr = x - (double)n * hpi;
    const double x2 = xr * xr;
    const double sp = sc_sin_poly(xr, x2), cp = sc_cos_poly(x2);
    // sine kernel negated for n & 3 in {1, 2}  <=>  bit 1 of n + 1;  cosine kernel negated (glibc's second table) for n & 2
    const uint32_t ms = ((uint32_t)(t1 + 0x1000000) << 6) & 0x80000000u;
    const uint32_t mc = ((uint32_t)t1 << 6) & 0x80000000u;
    const float sv = pdt_u2f(pdt_f2u((float)sp) ^ ms), cv = pdt_u2f(pdt_f2u((float)cp) ^ mc);
    const bool odd = (t1 & 0x1000000) != 0;
    s = odd ? cv : sv; c = odd ? sv : cv;
}

PDT_DEV void sincos_exact(float y, float &s, float &c)
{
    if (!sincos_in_core_range(y)) { s = sinf(y); c = cosf(y); return; }   // |y| >= 120: never reached by the PLL (phase is wrapped to ±2π)
    sincos_core(y, s, c);
}
PDT_DEV void sincos_exact(double y, double &s, double &c)
{
#ifdef __CUDA_ARCH__
    sincos(y, &s, &c);
#else
    s = sin(y); c = cos(y);
#endif
}

// hypot as glibc's cabsf()/cabs() compute it (AGC.c:58,66)
PDT_DEV float  hypot_exact(float re, float im)  { return (float)sqrt((double)re * (double)re + (double)im * (double)im); }
PDT_DEV double hypot_exact(double re, double im)
{
    // sqrt(x²+y²) with one exact-residual correction step (error < 0.51 ulp, like glibc's hypot)
    double ax = fabs(re), ay = fabs(im);
    if (ax < ay) { double t = ax; ax = ay; ay = t; }
    if (ay == 0.0) return ax;
    double h  = sqrt(fma(ax, ax, ay * ay));
    double h2 = h * h, hx = fma(h, h, -h2);
    double x2 = ax * ax, xx = fma(ax, ax, -x2);
    double y2 = ay * ay, yy = fma(ay, ay, -y2);
    double res = ((x2 - h2) + y2) + ((xx + yy) - hx);
    return h + res / (2.0 * h);
}

// ---------------------------------------------------------------------------------------------------
// CarrierTrackingPLL.c:15-40 (first-order atan2) and :43-52 (Q_rsqrt)
// ---------------------------------------------------------------------------------------------------
PDT_DEV real_t arctan2_approx(real_t y, real_t x)
{
    real_t abs_y = r_fabs(y) + 1e-10;                 // double add, narrowed (:21)
    real_t r, angle;
    if (x >= 0) { r = (x - abs_y) / (x + abs_y); angle = 0.78539816339744825 - 0.78539816339744825 * r; }
    else        { r = (x + abs_y) / (abs_y - x); angle = 2.35619449019234475 - 0.78539816339744825 * r; }
    return (y < 0) ? -angle : angle;
}

PDT_DEV float q_rsqrt(float x)
{
    float half = 0.5f * x;
    int bits = (int)pdt_f2u(x);
    bits = 0x5f3759df - (bits >> 1);
    x = pdt_u2f((uint32_t)bits);
    x = x * (1.5f - half * x * x);
    x = x * (1.5f - half * x * x);
    return x;
}

PDT_DEV int sign_of(real_t x) { return (x > 0) - (x < 0); }

// ---------------------------------------------------------------------------------------------------
// PLL — CarrierTrackingPLL.c:54-278
// ---------------------------------------------------------------------------------------------------
struct PllParams { real_t Fs, freq_range, lock_thresh, lock_alpha, bw_acq, bw_track; };

struct PllState {
    int      stage;          // 0 = uninitialised (firstLock==-2), 1 = searching (-1), 2 = locked
    real_t   damp, alpha, beta, phase, freq, max_freq, min_freq, avg_phase, locksig, sweep;
    double   lock_freq_hz;
    unsigned long long samples_seen, lock_sample;
    int      lock_event;     // set when the latch fires during the current call (host prints the message)
};

PDT_DEV void pll_reset(PllState &s) { s = PllState(); s.damp = 0.999; }

PDT_DEV void pll_begin(PllState &s, const PllParams &p)
{
    if (s.stage != 0) return;                            // :88-100
    const real_t bw = p.bw_acq;
    s.alpha = (4 * s.damp * bw) / (1 + 2 * s.damp * bw + bw * bw);
    s.beta  = (4 * bw * bw) / (1 + 2 * s.damp * bw + bw * bw);
    s.phase = 0.1;
    s.freq  = 2.0 * PDT_PI * 0 / p.Fs;
    s.max_freq = 2.0 * PDT_PI * p.freq_range / p.Fs;
    s.min_freq = -2.0 * PDT_PI * p.freq_range / p.Fs;
    s.stage = 1;
    s.avg_phase = PDT_PI / 2.0;
    s.sweep = 0.2 * (2.0 * PDT_PI / p.Fs);
}

// The pieces of one PLL sample, shared by the one-thread loop (pll_step) and the block-parallel runner (pdt_pll_pipe.cuh)
// so that both execute the same statements.
// loop filter + NCO advance, CarrierTrackingPLL.c:165-188
PDT_DEV void pll_loop_core(real_t &phase, real_t &freq, real_t sample_phase, real_t alpha, real_t beta, real_t max_freq, real_t min_freq)
{
    real_t err;                                          // :165-170
    if ((sample_phase - phase) > PDT_PI)        err = (sample_phase - phase) - 2 * PDT_PI;
    else if ((sample_phase - phase) < -PDT_PI)  err = (sample_phase - phase) + 2 * PDT_PI;
    else                                        err = sample_phase - phase;

    freq  = freq + beta * err;                           // :174
    phase = phase + freq + alpha * err;                  // :175
    while (phase > 2 * PDT_PI)  phase = phase - 2.0 * PDT_PI;    // :178-182
    while (phase < -2 * PDT_PI) phase = phase + 2.0 * PDT_PI;
    if (freq > max_freq)      freq = max_freq;           // :185-188
    else if (freq < min_freq) freq = min_freq;
}

PDT_DEV bool pll_noise_like(real_t avg_phase)
{
#if PDT_USE_FLOATS
    return fabsf((float)(PDT_PI / 2.0 - avg_phase)) < 0.05;      // :232
#else
    return fabs(PDT_PI / 2.0 - avg_phase) < 0.05;                // :248
#endif
}

// the acquisition sweep of a sample whose averaged phase looks like noise, :233-246
PDT_DEV void pll_sweep_core(real_t &freq, real_t &sweep, real_t max_freq, real_t min_freq)
{
    freq = freq + sweep;
    if (freq >= max_freq)       sweep = sweep * -1.0;
    else if (freq <= min_freq)  sweep = sweep * -1.0;
    else if (freq >= 0)         sweep = r_fabs(sweep);
    else                        sweep = r_fabs(sweep) * -1.0;
}

// the lock latch, :266-274
PDT_DEV void pll_latch(PllState &s, const PllParams &p, unsigned long long abs_index)
{
    s.lock_freq_hz = s.freq * p.Fs / (2.0 * PDT_PI);
    s.stage = 2;
    s.lock_sample = abs_index;
    s.lock_event = 1;
    const real_t bw = p.bw_track;
    s.alpha = (4.0 * s.damp * bw) / (1.0 + 2.0 * s.damp * bw + bw * bw);
    s.beta  = (4.0 * bw * bw) / (1.0 + 2.0 * s.damp * bw + bw * bw);
}

// one sample; `abs_index` = absolute sample index (for the lock record)
PDT_DEV void pll_step(PllState &s, const PllParams &p, real_t a, real_t b, real_t &out, real_t &lock,
                      unsigned long long abs_index)
{
    const real_t avg_alpha = 0.00005;
    real_t ti, tr;
    sincos_exact(s.phase, ti, tr);                       // :106-107
    const real_t nti = -ti;
    const real_t mre = a * tr - b * nti;                 // :110, four separately rounded products
    const real_t mim = a * nti + b * tr;
    out = mim;                                           // :113

    const real_t out_phase = arctan2_approx(mim, mre);   // :117
    s.avg_phase = s.avg_phase * (1.0 - avg_alpha) + avg_alpha * r_fabs(out_phase);   // :124

    const real_t sample_phase = arctan2_approx(b, a);    // :128
    pll_loop_core(s.phase, s.freq, sample_phase, s.alpha, s.beta, s.max_freq, s.min_freq);

    real_t nre = a, nim = b;                             // :193-220
    const real_t mag2 = nre * nre + nim * nim;
    const real_t inv  = q_rsqrt((float)mag2);
    nre *= inv; nim *= inv;
    s.locksig = s.locksig * (1.0 - p.lock_alpha) + p.lock_alpha * (nre * tr + nim * ti);
    lock = s.locksig;

    if (pll_noise_like(s.avg_phase) && s.stage == 1) pll_sweep_core(s.freq, s.sweep, s.max_freq, s.min_freq);
    if (s.locksig > p.lock_thresh && s.stage == 1) pll_latch(s, p, abs_index);
}

// ---------------------------------------------------------------------------------------------------
// AGC — AGC.c:78-132 (NormalizingAGC), :48-75 (StaticGain)
// ---------------------------------------------------------------------------------------------------
struct AgcState { int init; real_t gain; };

PDT_DEV real_t agc_step(AgcState &s, real_t x, real_t attack, real_t decay)
{
    // AGC.c:98-131, written so that the compiler emits selects instead of branches (both rate products are formed,
    // the one the reference would use is selected): the gain recurrence is the serial bottleneck of this stage.
    const real_t reference = 1.0, max_gain = 5000;
    x *= s.gain;
    const real_t err = r_fabs(x) - reference;
    const real_t da = err * attack, dd = err * decay;
    real_t g = s.gain - ((r_fabs(err) > s.gain) ? da : dd);
    g = (g < 0.0) ? (real_t)10e-5 : g;
    g = (g > max_gain) ? max_gain : g;
    s.gain = g;
    return x;
}

// Four samples of AGC.c:98-131 in its common regime — decay branch, no clamp: gain' = gain - (|x·gain| - 1)·decay, four dependent
// operations per sample — with the conditions of the other branches evaluated beside the chain; a group in which any of them
// holds is redone by the general step from the saved gain.  Same results as four agc_step calls, always
// (tests/host/loop_forms.cu compares the two on the host).
PDT_DEV void agc_step4(AgcState &s, real_t x0, real_t x1, real_t x2, real_t x3, real_t attack, real_t decay,
                       real_t &v0, real_t &v1, real_t &v2, real_t &v3)
{
    const real_t g_in = s.gain;
    real_t g = g_in;
    bool other = false;
    v0 = x0 * g; { const real_t e = r_fabs(v0) - (real_t)1.0; other |= r_fabs(e) > g; g = g - e * decay; other |= (g < 0.0) | (g > (real_t)5000); }
    v1 = x1 * g; { const real_t e = r_fabs(v1) - (real_t)1.0; other |= r_fabs(e) > g; g = g - e * decay; other |= (g < 0.0) | (g > (real_t)5000); }
    v2 = x2 * g; { const real_t e = r_fabs(v2) - (real_t)1.0; other |= r_fabs(e) > g; g = g - e * decay; other |= (g < 0.0) | (g > (real_t)5000); }
    v3 = x3 * g; { const real_t e = r_fabs(v3) - (real_t)1.0; other |= r_fabs(e) > g; g = g - e * decay; other |= (g < 0.0) | (g > (real_t)5000); }
    if (!other) { s.gain = g; return; }
    s.gain = g_in;
    v0 = agc_step(s, x0, attack, decay); v1 = agc_step(s, x1, attack, decay);
    v2 = agc_step(s, x2, attack, decay); v3 = agc_step(s, x3, attack, decay);
}

PDT_DEV real_t static_gain_serial(const real_t *iq, unsigned long long n, real_t desired)
{
    real_t level = hypot_exact(iq[0], iq[1]);
    for (unsigned long long i = 0; i < n; i++) {
        level += hypot_exact(iq[2 * i], iq[2 * i + 1]);
        level /= 2.0;
    }
    return desired / level;
}

// ---------------------------------------------------------------------------------------------------
// FIR — LowPassFilter.c:13-71 (zero-stuffing interpolator) and :76-125 (plain)
//
// The reference keeps an N-slot ring and, for every output t (global counter), sums only the slots
// that hold real samples, IN RING-SLOT ORDER.  Restated without the ring: with p = t mod L,
// j = t div L (index of the newest real sample) and K = N/L taps per branch,
//     y[t] = Σ_k h[N-1-p-kL] · x[j-k]   summed in the order k = k0, k0-1, …, 0, K-1, …, k0+1,  k0 = j mod K
// starting from +0 (x[<0] = 0).  `xh` points at x[j] inside a buffer that has K-1 history samples
// in front of it.
// ---------------------------------------------------------------------------------------------------
PDT_DEV real_t fir_interp_exact(const real_t *__restrict__ h, const real_t *__restrict__ xh, int N, int L, int K,
                                int p, int k0)
{
    real_t acc = 0;
    const real_t *hp = h + (N - 1 - p);
    for (int k = k0; k >= 0; k--)      acc += hp[-k * L] * xh[-k];
    for (int k = K - 1; k > k0; k--)   acc += hp[-k * L] * xh[-k];
    return acc;
}

// plain FIR: y = Σ_{k=0}^{N-1} h[k]·x[o-(N-1)+k], oldest first (:113-116); xh points at x[o]
PDT_DEV real_t fir_plain_exact(const real_t *__restrict__ h, const real_t *__restrict__ xh, int N)
{
    real_t acc = 0;
    for (int k = 0; k < N; k++) acc += h[k] * xh[k - (N - 1)];
    return acc;
}

// ---------------------------------------------------------------------------------------------------
// Gardner — GardenerClockRecovery.c:5-114 ; Manchester — ManchesterDecode.c:10-100
// ---------------------------------------------------------------------------------------------------
struct GardnerState { int init; real_t next, prev, half, step; };
struct ManchesterState { unsigned clockmod; real_t cur, prev, prevprev; unsigned char even_odd; };

PDT_DEV void gardner_begin(GardnerState &s, int Fs, real_t baud) { if (!s.init) { s.step = Fs / baud; s.init = 1; } }

// one symbol; returns the picked index, symbol value in `sym`, clamped error in `err`
// `x`: anything indexable by an unsigned sample index (a pointer, or a window accessor)
template <class X>
PDT_DEV unsigned gardner_step(GardnerState &s, const X &x, real_t range, real_t kp, real_t &sym, real_t &err)
{
    const unsigned at = (unsigned)(r_rint(s.next));
    const real_t cur = x[at];
    s.half = x[(unsigned)(r_rint(s.half))];              // index-then-value reuse (:28)
    real_t e = kp * (cur - s.prev) * (s.half);           // :43
    if (e > range) e = range; else if (e < -range) e = -range;
    s.next = (s.next - e);
    s.half = s.next + s.step / 2.0;                      // :59
    s.next = s.next + s.step;
    s.prev = cur;
    sym = cur; err = e;
    return at;
}

// Mueller & Müller — MMClockRecovery.c:5-84 (exported by the reference; its call is commented out in both drivers)
struct MMState { int init; real_t step, next, last; };

PDT_DEV real_t mm_rint(real_t v)
{
#if PDT_USE_FLOATS
    return rintf(v);                       // <tgmath.h> rint() of a float (:23,:26)
#else
    return (real_t)rintf((float)v);        // the double build rounds through rintf (:53,:56) — kept
#endif
}
PDT_DEV void mm_begin(MMState &s, int Fs, real_t baud) { if (!s.init) { s.step = Fs / (baud); s.init = 1; } }
// one symbol at the current position (the caller checks mm_rint(next) < n first); returns the picked index
template <class X>
PDT_DEV unsigned mm_step(MMState &s, const X &x, real_t step_min, real_t step_max, real_t kp, real_t &sym)
{
    const unsigned at = (unsigned)(mm_rint(s.next));
    const real_t cur = x[at];
    const real_t err = sign_of(s.last) * cur - sign_of(cur) * s.last;          // :35
    s.step = s.step + kp * err;
    if (s.step > step_max) s.step = step_max;
    if (s.step < step_min) s.step = step_min;
    s.next = s.next + s.step;
    s.last = cur;
    sym = cur;
    return at;
}

// one symbol; returns 1 and sets `bit` ('0'/'1') when a bit is produced
PDT_DEV int manchester_step(ManchesterState &s, real_t v, real_t thresh, unsigned char &bit)
{
    int produced = 0;
    s.prevprev = s.prev; s.prev = s.cur; s.cur = v;
    if ((unsigned)(s.even_odd % 2) != s.clockmod) {
        if (sign_of(s.prevprev) == sign_of(s.prev))
            if (r_fabs(s.prevprev) > thresh && r_fabs(s.prev) > thresh) s.clockmod = (s.even_odd % 2);
    }
    if ((unsigned)(s.even_odd % 2) == s.clockmod) {
        if (r_fabs(s.prev) > r_fabs(s.cur)) bit = (s.prev > 0) ? '1' : '0';
        else                                 bit = (s.cur > 0) ? '0' : '1';
        produced = 1;
    }
    s.even_odd++;
    return produced;
}

// ---------------------------------------------------------------------------------------------------
// ByteSync — POESTIPdemod/ByteSync.c:16-150, ARGOSdemod/ByteSync.c:17-150
// The circular ASCII history is restated as a 32-bit shift register (newest bit = LSB); a sync hit is
// (hist & mask) == word, the inverse hit (~hist & mask) == word.
// ---------------------------------------------------------------------------------------------------
struct SyncParams {
    uint32_t word, mask; int len;
    int last_idx;        // 103 (POES) / 8 (ARGOS): frame ends when frame_byte_idx > last_idx
    int carry_bits;      // bitIdx after accept: 3 (POES: 19 = 16 + 3) / 0 (ARGOS)
    int inverse_enabled; // POES 1 / ARGOS 0 (:115)
};
struct SyncState { uint32_t hist; int in_frame, frame_byte_idx, bit_idx; unsigned char zero, one, byte; };

enum { EV_NONE = 0, EV_SYNC = 1, EV_SYNC_INV = 2 };

// one bit; returns EV_* for an accepted sync; `emit` = 1 when a byte was completed (value in out_byte),
// `eol` = 1 when that byte closed the frame.  Byte emission precedes sync detection, as in the reference.
PDT_DEV int sync_step(SyncState &s, const SyncParams &p, unsigned char bit, int &emit, unsigned char &out_byte, int &eol)
{
    emit = 0; eol = 0;
    if (s.in_frame == 1) {
        s.byte = (unsigned char)(s.byte << 1);
        s.byte |= (bit == '0') ? s.zero : s.one;
        s.bit_idx++;
        if (s.bit_idx > 7) {
            emit = 1; out_byte = s.byte;
            s.byte = 0; s.bit_idx = 0; s.frame_byte_idx++;
            if (s.frame_byte_idx > p.last_idx) { s.in_frame = 0; eol = 1; }
        }
    }
    s.hist = (s.hist << 1) | (uint32_t)(bit != '0');     // anything that is not '0' compares unequal to '0'
    int ev = EV_NONE;
    // NB: a non-'0'/'1' byte can never equal a sync character; '0'/'1' streams are all the chain produces.
    if (((s.hist & p.mask) == p.word) && s.in_frame == 0) {
        s.frame_byte_idx = 2; s.in_frame = 1; s.bit_idx = p.carry_bits; s.byte = 0; s.zero = 0; s.one = 1;
        ev = EV_SYNC;
    }
    if (p.inverse_enabled && ((~s.hist & p.mask) == p.word) && s.in_frame == 0) {
        s.frame_byte_idx = 2; s.in_frame = 1; s.bit_idx = p.carry_bits; s.byte = 0; s.zero = 1; s.one = 0;
        ev = EV_SYNC_INV;
    }
    return ev;
}
