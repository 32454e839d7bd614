/*
 * pdt_oracle.c — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).  See pdt_oracle.h.
 *
 * Every function cites the reference file:line it restates.  Expressions deliberately keep the
 * reference's literal types (1.0 vs 1, 2.0*M_PI …) because with DECIMAL_TYPE=float the C usual
 * arithmetic conversions decide which sub-expressions are evaluated in double.
 * Compile with -ffp-contract=off (see oracle/Makefile) so no FMA contraction changes rounding.
 */
#define _GNU_SOURCE
#include "pdt_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#if PDT_USE_FLOATS
#define R_SIN  sinf
#define R_COS  cosf
#define R_FABS fabsf
#define R_RINT rintf
#define R_HYPOT(re, im) cabsf_like(re, im)
#else
#define R_SIN  sin
#define R_COS  cos
#define R_FABS fabs
#define R_RINT rint
#define R_HYPOT(re, im) cabs_like(re, im)
#endif

/* cabsf()/cabs() of glibc are hypotf()/hypot(); call those directly (AGC.c:58,66). */
static inline float  cabsf_like(float re, float im)   { return hypotf(re, im); }
static inline double cabs_like(double re, double im)  { return hypot(re, im); }

/* ------------------------------------------------------------------------------------------------
 * helpers: CarrierTrackingPLL.c:15-52, MMClockRecovery.c:86-89
 * ---------------------------------------------------------------------------------------------- */
#define ATAN_C1 (0.78539816339744825)
#define ATAN_C2 (2.35619449019234475)

pdto_real pdto_arctan2(pdto_real y, pdto_real x)
{
    /* first-order approximation; |y| gets the DOUBLE literal 1e-10 added before narrowing (:21/:23) */
    pdto_real abs_y = R_FABS(y) + 1e-10;
    pdto_real r, angle;
    if (x >= 0) {
        r = (x - abs_y) / (x + abs_y);
        angle = ATAN_C1 - ATAN_C1 * r;          /* double arithmetic, narrowed on store (:29) */
    } else {
        r = (x + abs_y) / (abs_y - x);
        angle = ATAN_C2 - ATAN_C1 * r;          /* (:34) */
    }
    return (y < 0) ? -angle : angle;
}

float pdto_q_rsqrt(float x)
{
    /* 0x5f3759df estimate + two Newton steps, float even in the double build (:43-52) */
    float half = 0.5f * x;
    int32_t bits;
    memcpy(&bits, &x, sizeof bits);
    bits = 0x5f3759df - (bits >> 1);
    memcpy(&x, &bits, sizeof x);
    x = x * (1.5f - half * x * x);
    x = x * (1.5f - half * x * x);
    return x;
}

int pdto_sign(pdto_real x) { return (x > 0) - (x < 0); }

/* ------------------------------------------------------------------------------------------------
 * AGC.c:48-75 StaticGain, AGC.c:24-46 Squelch
 * ---------------------------------------------------------------------------------------------- */
pdto_real pdto_static_gain(const pdto_real *iq, unsigned int n, pdto_real desired)
{
    /* NOT a mean: level = (level + |x|)/2 per sample, seeded with |x[0]| (:58-71) */
    pdto_real level = R_HYPOT(iq[0], iq[1]);
    for (unsigned long i = 0; i < n; i++) {
        level += R_HYPOT(iq[2 * i], iq[2 * i + 1]);
        level /= 2.0;
    }
    return desired / level;
}

void pdto_squelch(pdto_real *x, const pdto_real *lock, unsigned long n, pdto_real thresh)
{
    for (unsigned long i = 0; i < n; i++)
        if (lock[i] < thresh) x[i] = 0;
}

/* ------------------------------------------------------------------------------------------------
 * CarrierTrackingPLL.c:54-278
 * ---------------------------------------------------------------------------------------------- */
void pdto_pll_reset(pdto_pll *s)
{
    memset(s, 0, sizeof *s);
    s->first_lock = -2;
    s->damp = 0.999;
}

static void pll_gains_init(pdto_pll *s, pdto_real bw)
{
    /* integer literals: evaluated in DECIMAL_TYPE (:90-91) */
    s->alpha = (4 * s->damp * bw) / (1 + 2 * s->damp * bw + bw * bw);
    s->beta  = (4 * bw * bw) / (1 + 2 * s->damp * bw + bw * bw);
}
static void pll_gains_track(pdto_pll *s, pdto_real bw)
{
    /* double literals: evaluated in double, narrowed on store (:272-273) */
    s->alpha = (4.0 * s->damp * bw) / (1.0 + 2.0 * s->damp * bw + bw * bw);
    s->beta  = (4.0 * bw * bw) / (1.0 + 2.0 * s->damp * bw + bw * bw);
}

pdto_real pdto_pll_run(pdto_pll *s, const pdto_real *iq, pdto_real *real_out, pdto_real *lock_out,
                       unsigned int n, pdto_real Fs, pdto_real freq_range, pdto_real lock_thresh,
                       pdto_real lock_alpha, pdto_real bw_acq, pdto_real bw_track,
                       pdto_real *trace_phase, pdto_real *trace_freq)
{
    const pdto_real avg_alpha = 0.00005;                 /* :80 */
    if (s->first_lock == -2) {                           /* :88-100 */
        pll_gains_init(s, bw_acq);
        s->phase     = 0.1;
        s->freq      = 2.0 * M_PI * 0 / Fs;
        s->max_freq  = 2.0 * M_PI * freq_range / Fs;
        s->min_freq  = -2.0 * M_PI * freq_range / Fs;
        s->first_lock = -1;
        s->avg_phase = M_PI / 2.0;
        s->sweep     = 0.2 * (2.0 * M_PI / Fs);
    }

    for (unsigned int i = 0; i < n; i++) {
        const pdto_real a = iq[2 * i], b = iq[2 * i + 1];
        if (trace_phase) trace_phase[i] = s->phase;
        if (trace_freq)  trace_freq[i]  = s->freq;

        const pdto_real ti = R_SIN(s->phase);            /* :106-107 / :134-135 */
        const pdto_real tr = R_COS(s->phase);
        /* x * (tr - j*ti): the four products are rounded separately (:110) */
        const pdto_real nti = -ti;
        const pdto_real mre = a * tr - b * nti;
        const pdto_real mim = a * nti + b * tr;
        real_out[i] = mim;                               /* data is on the imaginary axis (:113) */

        const pdto_real out_phase = pdto_arctan2(mim, mre);                     /* :117 */
        s->avg_phase = s->avg_phase * (1.0 - avg_alpha) + avg_alpha * R_FABS(out_phase);   /* :124 */

        const pdto_real sample_phase = pdto_arctan2(b, a);                      /* :128 */
        pdto_real err;                                                          /* :165-170 */
        if ((sample_phase - s->phase) > M_PI)        err = (sample_phase - s->phase) - 2 * M_PI;
        else if ((sample_phase - s->phase) < -M_PI)  err = (sample_phase - s->phase) + 2 * M_PI;
        else                                         err = sample_phase - s->phase;

        s->freq  = s->freq + s->beta * err;                                     /* :174 */
        s->phase = s->phase + s->freq + s->alpha * err;                         /* :175 */
        while (s->phase > 2 * M_PI)  s->phase = s->phase - 2.0 * M_PI;          /* :178-182 */
        while (s->phase < -2 * M_PI) s->phase = s->phase + 2.0 * M_PI;
        if (s->freq > s->max_freq)      s->freq = s->max_freq;                  /* :185-188 */
        else if (s->freq < s->min_freq) s->freq = s->min_freq;

        /* lock detector on the Q_rsqrt-normalised input (:193-220) */
        pdto_real nre = a, nim = b;
        const pdto_real mag2 = nre * nre + nim * nim;
        const pdto_real inv  = pdto_q_rsqrt((float)mag2);
        nre *= inv; nim *= inv;
        s->locksig = s->locksig * (1.0 - lock_alpha) + lock_alpha * (nre * tr + nim * ti);
        if (lock_out) lock_out[i] = s->locksig;                                 /* :222-223 */

        /* sweep while the averaged |phase| still looks like noise (:231-263) */
#if PDT_USE_FLOATS
        if (fabsf(M_PI / 2.0 - s->avg_phase) < 0.05 && s->first_lock == -1) {
#else
        if (fabs(M_PI / 2.0 - s->avg_phase) < 0.05 && s->first_lock == -1) {
#endif
            s->freq = s->freq + s->sweep;
            if (s->freq >= s->max_freq)       s->sweep = s->sweep * -1.0;
            else if (s->freq <= s->min_freq)  s->sweep = s->sweep * -1.0;
            else if (s->freq >= 0)            s->sweep = R_FABS(s->sweep);
            else                              s->sweep = R_FABS(s->sweep) * -1.0;
        }

        /* one-way lock latch + gain switch (:266-274) */
        if (s->locksig > lock_thresh && s->first_lock == -1) {
            s->lock_freq_hz = s->freq * Fs / (2.0 * M_PI);
            s->first_lock   = i;
            s->lock_sample  = s->samples_seen + i;
            pll_gains_track(s, bw_track);
        }
    }
    s->samples_seen += n;
    return s->avg_phase;
}

/* ------------------------------------------------------------------------------------------------
 * LowPassFilter.c:127-175 MakeLPFIR
 * ---------------------------------------------------------------------------------------------- */
int pdto_make_lpfir(pdto_real *h, int N, pdto_real Fc, pdto_real Fs, int L)
{
    pdto_real T   = 1.0 / Fs;
    pdto_real wc  = 2.0 * M_PI * Fc * T;
    pdto_real tou = (N - 1.0) / 2.0;
    int n;
    for (n = 0; n < N; n++) {
        /* DECIMAL_TYPE sine over a double denominator, narrowed (:148) */
        pdto_real hd = (R_SIN(wc * (n - tou))) / (M_PI * (n - tou));
        if ((n == tou) && ((int)((N / 2) * 2) != N)) hd = wc / M_PI;            /* odd N centre (:151-154) */
        pdto_real wn = 0.42 - 0.5 * cos((2 * M_PI * n) / (N - 1)) + 0.08 * cos((4 * M_PI * n) / (N - 1));
        h[n] = hd * wn * (pdto_real)(L);                                         /* :166 */
    }
    return n;
}

/* ------------------------------------------------------------------------------------------------
 * LowPassFilter.c:13-71 LowPassFilterInterp, :76-125 LowPassFilter
 * ---------------------------------------------------------------------------------------------- */
void pdto_fir_reset(pdto_fir *s) { memset(s, 0, sizeof *s); }

void pdto_fir_interp_run(pdto_fir *s, const pdto_real *in_time, const pdto_real *in, pdto_real *out,
                         pdto_real *out_time, unsigned long n, const pdto_real *h, int N, int L)
{
    if (L <= 0) return;              /* L=0: reference loop bound n*L is 0 -> emits nothing */
    s->init = 1;
    unsigned long src = 0;
    const unsigned long n_out = n * (unsigned long)L;
    for (unsigned long o = 0; o < n_out; o++, s->interp_counter++) {
        if ((s->interp_counter % L) == 0) {               /* a real sample enters the ring (:45-49) */
            s->ring[s->oldest] = in[src++];
            s->interp_counter = 0;
        } else {
            s->ring[s->oldest] = 0;                       /* zero stuffing (:51) */
        }
        /* only every L-th ring slot can be non-zero; summation runs in RING-SLOT order, i.e. it is
         * rotated by `oldest`, and the taps are time-reversed: newest sample meets h[N-1] (:58-64) */
        pdto_real acc = 0;
        for (int slot = 0; slot < N; slot += L) {
            int tap = (N - 1 - s->oldest + slot) % N;     /* == (N-(oldest-slot+1)) % N */
            acc += h[tap] * s->ring[slot];
        }
        out[o] = acc;
        if (out_time) out_time[o] = in_time[src];         /* post-incremented index: NEXT input's time,
                                                             one-past-the-end on the last L-1 outputs (:68) */
        s->oldest = (s->oldest + 1) % N;
    }
}

void pdto_fir_run(pdto_fir *s, pdto_real *x, unsigned long n, const pdto_real *h, int N)
{
    s->init = 1;
    for (unsigned long o = 0; o < n; o++) {
        s->ring[s->oldest] = x[o];
        pdto_real acc = 0;
        for (int k = 0; k < N; k++)                       /* h[0] on the oldest sample (:113-116) */
            acc += h[k] * s->ring[(s->oldest + k + 1) % N];
        x[o] = acc;
        s->oldest = (s->oldest + 1) % N;
    }
}

/* ------------------------------------------------------------------------------------------------
 * AGC.c:78-132 NormalizingAGC
 * ---------------------------------------------------------------------------------------------- */
void pdto_agc_reset(pdto_agc *s) { s->init = 0; s->gain = 1; }

void pdto_agc_run(pdto_agc *s, pdto_real *x, unsigned long n, pdto_real initial, pdto_real attack,
                  pdto_real decay, pdto_real *trace_gain)
{
    const pdto_real reference = 1.0, max_gain = 5000;
    if (!s->init) { s->init = 1; s->gain = initial; }      /* first call latches the seed (:92-96) */
    for (unsigned long i = 0; i < n; i++) {
        if (trace_gain) trace_gain[i] = s->gain;
        x[i] *= s->gain;
        pdto_real err  = R_FABS(x[i]) - reference;
        pdto_real rate = decay;
        if (R_FABS(err) > s->gain) rate = attack;          /* :112-119 */
        s->gain -= err * rate;
        if (s->gain < 0.0) s->gain = 10e-5;                /* :124-125 */
        if (max_gain > 0.0 && s->gain > max_gain) s->gain = max_gain;
    }
}

/* ------------------------------------------------------------------------------------------------
 * AGC.c:164-200 NormalizingAGCC, AGC.c:6-20 FindSignalAmplitude (exported by the reference, not called by its drivers)
 * ---------------------------------------------------------------------------------------------- */
void pdto_agcc_run(pdto_agc *s, pdto_real *iq, unsigned long n, pdto_real initial, pdto_real loop_gain)
{
    const pdto_real desired = 5;
    if (!s->init) { s->init = 1; s->gain = initial; }      /* :177-181 */
    for (unsigned long i = 0; i < n; i++) {
        iq[2 * i] *= s->gain; iq[2 * i + 1] *= s->gain;    /* complex *= real (:185) */
#if PDT_USE_FLOATS
        /* :188 calls fabsf() on a float complex: fabsf is NOT one of <tgmath.h>'s macros, so the argument converts to
         * float by dropping the imaginary part - the float build measures |Re| */
        pdto_real mag = fabsf(iq[2 * i]);
#else
        /* :190 calls fabs(), which <tgmath.h> dispatches to cabs() for a complex argument */
        pdto_real mag = R_HYPOT(iq[2 * i], iq[2 * i + 1]);
#endif
        pdto_real err = desired - (s->gain * mag);
        s->gain = s->gain + loop_gain * err;               /* :198 */
    }
}

pdto_real pdto_amp_run(pdto_real *average, const pdto_real *x, unsigned long n, pdto_real alpha)
{
    pdto_real avg = *average;
    for (unsigned long i = 0; i < n; i++)
        avg = avg * (1.0 - alpha) + alpha * R_FABS(x[i]);  /* :14/:16, the 1.0 makes the first product double */
    *average = avg;
    return avg;
}

/* ------------------------------------------------------------------------------------------------
 * GardenerClockRecovery.c:5-114
 * ---------------------------------------------------------------------------------------------- */
void pdto_gardner_reset(pdto_gardner *s) { memset(s, 0, sizeof *s); }

unsigned long pdto_gardner_run(pdto_gardner *s, const pdto_real *x, pdto_real *time, unsigned long n,
                               pdto_real *out, int Fs, pdto_real baud, pdto_real step_range, pdto_real kp,
                               uint32_t *trace_idx, pdto_real *trace_err)
{
    unsigned long count = 0;
    if (!s->init) { s->step = Fs / baud; s->init = 1; }    /* int / DECIMAL (:19) */
    while (R_RINT(s->next) < n) {
        const unsigned int at = (unsigned int)(R_RINT(s->next));
        const pdto_real cur = x[at];
        /* `half` holds an INDEX here and a VALUE afterwards; after a chunk rollover the index is
         * still expressed in the previous chunk's coordinates (:28) */
        s->half = x[(unsigned int)(R_RINT(s->half))];
        out[count] = cur;
        if (time) time[count] = time[at];
        pdto_real err = kp * (cur - s->prev) * (s->half);   /* :43 */
        if (err > step_range) err = step_range;
        else if (err < -step_range) err = -step_range;
        if (trace_idx) trace_idx[count] = at;
        if (trace_err) trace_err[count] = err;
        s->next = (s->next - err);
        s->half = s->next + s->step / 2.0;                   /* double intermediate (:59) */
        s->next = s->next + s->step;
        s->prev = cur;
        count++;
    }
    if (time) time[count] = time[(unsigned int)(R_RINT(s->next))];   /* :65 (reads past n) */
    s->next = s->next - n;                                   /* chunk-relative coordinates (:111) */
    return count;
}

/* MMClockRecovery.c:5-84 (exported, never called by either main.c) */
void pdto_mm_reset(pdto_mm *s) { memset(s, 0, sizeof *s); s->step = 3.0; }

unsigned long pdto_mm_run(pdto_mm *s, const pdto_real *x, pdto_real *time, unsigned long n, pdto_real *out,
                          int Fs, pdto_real baud, pdto_real step_range, pdto_real kp)
{
    const pdto_real step_max = Fs / (baud - step_range);
    const pdto_real step_min = Fs / (baud + step_range);
    unsigned long count = 0;
    if (!s->init) { s->step = Fs / (baud); s->init = 1; }
    /* quirk kept: the float build rounds with rint(), the double build with rintf() (:27 vs :55) */
#if PDT_USE_FLOATS
#define MM_RINT(v) rintf(v)
#else
#define MM_RINT(v) rintf((float)(v))
#endif
    while (MM_RINT(s->next) < n) {
        const unsigned int at = (unsigned int)(MM_RINT(s->next));
        const pdto_real cur = x[at];
        out[count] = cur;
        if (time) time[count] = time[at];
        count = count + 1;
        pdto_real err = pdto_sign(s->last) * cur - pdto_sign(cur) * s->last;
        s->step = s->step + kp * err;
        if (s->step > step_max) s->step = step_max;
        if (s->step < step_min) s->step = step_min;
        s->next = s->next + s->step;
        s->last = cur;
    }
#undef MM_RINT
    s->next = s->next - n;
    return count;
}

/* ------------------------------------------------------------------------------------------------
 * ManchesterDecode.c:10-100
 * ---------------------------------------------------------------------------------------------- */
void pdto_manchester_reset(pdto_manchester *s) { memset(s, 0, sizeof *s); }

unsigned long pdto_manchester_run(pdto_manchester *s, const pdto_real *sym, pdto_real *time, unsigned long n,
                                  unsigned char *bits, pdto_real resync_thresh)
{
    unsigned long o = 0;
    for (unsigned long i = 0; i < n; i++, s->even_odd++) {
        s->prevprev = s->prev;
        s->prev     = s->cur;
        s->cur      = sym[i];
        /* off-boundary: two equal-signed strong symbols in a row mean the pair boundary is wrong (:35-53) */
        if ((s->even_odd % 2) != (int)s->clockmod) {
            if (pdto_sign(s->prevprev) == pdto_sign(s->prev))
                if (R_FABS(s->prevprev) > resync_thresh && R_FABS(s->prev) > resync_thresh)
                    s->clockmod = (s->even_odd % 2);
        }
        /* boundary: decide on the stronger half of the pair (:57-93) */
        if ((s->even_odd % 2) == (int)s->clockmod) {
            unsigned char bit;
            if (R_FABS(s->prev) > R_FABS(s->cur)) bit = (s->prev > 0) ? '1' : '0';
            else                                   bit = (s->cur > 0) ? '0' : '1';
            bits[o] = bit;
            if (time) time[o] = time[i];
            o++;
        }
    }
    return o;
}

/* ------------------------------------------------------------------------------------------------
 * ByteSync: POESTIPdemod/ByteSync.c:16-150, ARGOSdemod/ByteSync.c:17-150
 * ---------------------------------------------------------------------------------------------- */
static void sink_printf(pdto_bytesync *s, const char *fmt, double v, int as_int)
{
    char tmp[64];
    int len = as_int ? snprintf(tmp, sizeof tmp, fmt, (unsigned)v) : snprintf(tmp, sizeof tmp, fmt, v);
    if (s->text_len + (size_t)len + 1 > s->text_cap) {
        s->text_cap = (s->text_cap ? s->text_cap * 2 : 4096) + (size_t)len;
        s->text = (char *)realloc(s->text, s->text_cap);
        if (!s->text) { printf("Error in malloc\n"); exit(1); }
    }
    memcpy(s->text + s->text_len, tmp, (size_t)len + 1);
    s->text_len += (size_t)len;
}
static void sink_str(pdto_bytesync *s, const char *str) { sink_printf(s, str, 0, 0); }

void pdto_bytesync_reset(pdto_bytesync *s)
{
    char *t = s->text; size_t cap = s->text_cap;
    memset(s, 0, sizeof *s);
    s->text = t; s->text_cap = cap; s->one = 1;
    if (s->text) s->text[0] = 0;
}
void pdto_bytesync_free(pdto_bytesync *s) { free(s->text); s->text = NULL; s->text_cap = s->text_len = 0; }

/* frame_last_idx: 103 for the 104-byte TIP minor frame, 8 for ARGOS; poes selects the literal ED E2
 * prefix, the 3-bit carry-in (19-bit sync = 2 bytes + 3 bits) and the enabled inverse search. */
/* `poes`: 1 = POESTIPdemod/ByteSync.c, 0 = ARGOSdemod/ByteSync.c, 2 = the parameterised common/ByteSync.c:16-144
 * (frame ends when frameByteIdx > frameLength, bitIdx = startBit after a sync word, inverse search on, literal ED E2). */
static int bytesync_core_ex(pdto_bytesync *s, const unsigned char *bits, const pdto_real *time, unsigned long n,
                            const char *sync, unsigned int len, int poes, int frame_len, int start_bit)
{
    int found = 0;
    const int last_idx = poes == 2 ? frame_len : (poes ? 103 : 8);
    const int carry = poes == 2 ? start_bit : (poes ? 3 : 0);
    if (!s->init) { s->init = 1; memset(s->hist, 48, len); }
    for (unsigned long i = 0; i < n; i++) {
        if (s->in_frame == 1) {                               /* shift payload bits into bytes (:45-72) */
            s->byte = (unsigned char)(s->byte << 1);
            s->byte |= (bits[i] == '0') ? s->zero : s->one;
            s->bit_idx++;
            if (s->bit_idx > 7) {
                sink_printf(s, "%.2X ", s->byte, 1);
                s->byte = 0; s->bit_idx = 0; s->frame_byte_idx++;
                if (s->frame_byte_idx > last_idx) { s->in_frame = 0; sink_str(s, "\n"); }
            }
        }
        s->hist[s->oldest] = (char)bits[i];                   /* :75 */
        int hit = 1, hit_inv = poes ? 1 : 0;                  /* ARGOS: inverse disabled (:115) */
        for (unsigned int k = 0; k < len; k++) {
            char hb = s->hist[(s->oldest + k + 1) % len];
            if (sync[k] != hb) hit = 0;
            if (sync[k] == hb) hit_inv = 0;
        }
        const double t = time ? (double)time[i] : 0.0;
        if (hit && s->in_frame == 0) {                        /* accepted only outside a frame (:93) */
            sink_printf(s, "%.5f ", t, 0);
            if (poes) { sink_printf(s, "%.2X ", 0xED, 1); sink_printf(s, "%.2X ", 0xE2, 1); }
            s->frame_byte_idx = 2; s->in_frame = 1; found++;
            s->bit_idx = carry; s->byte = 0; s->zero = 0; s->one = 1;
        }
        if (hit_inv && s->in_frame == 0) {                    /* :127-144 */
            sink_printf(s, "%.5fi ", t, 0);
            if (poes) { sink_printf(s, "%.2X ", 0xED, 1); sink_printf(s, "%.2X ", 0xE2, 1); }
            s->frame_byte_idx = 2; s->in_frame = 1; found++;
            s->bit_idx = carry; s->byte = 0; s->zero = 1; s->one = 0;
        }
        s->oldest = (s->oldest + 1) % (int)len;
    }
    return found;
}

static int bytesync_core(pdto_bytesync *s, const unsigned char *bits, const pdto_real *time, unsigned long n,
                         const char *sync, unsigned int len, int poes)
{ return bytesync_core_ex(s, bits, time, n, sync, len, poes, 0, 0); }

int pdto_bytesync_generic_run(pdto_bytesync *s, const unsigned char *bits, const pdto_real *time, unsigned long n,
                              const char *sync, unsigned int len, int frame_len, int start_bit)
{ return bytesync_core_ex(s, bits, time, n, sync, len, 2, frame_len, start_bit); }

int pdto_bytesync_poes_run(pdto_bytesync *s, const unsigned char *bits, const pdto_real *time, unsigned long n,
                           const char *sync, unsigned int len)
{ return bytesync_core(s, bits, time, n, sync, len, 1); }

int pdto_bytesync_argos_run(pdto_bytesync *s, const unsigned char *bits, const pdto_real *time, unsigned long n,
                            const char *sync, unsigned int len)
{ return bytesync_core(s, bits, time, n, sync, len, 0); }

/* ------------------------------------------------------------------------------------------------
 * chain drivers: POESTIPdemod/main.c:346-482, ARGOSdemod/main.c:244-300
 * ---------------------------------------------------------------------------------------------- */
#define POES_MAX_DEV     (4500.0)
#define POES_ACQ_GAIN    127.3240
#define POES_TRCK_GAIN   10.3451
#define POES_LOCK_ALPHA  0.3979
#define POES_LOCK_THRESH (0.08)
#define POES_BAUD        (8320*2+0.3)
#define AGC_ATCK         (79.5775)
#define AGC_DCY          (159.1549)
#define POES_LPF_FC      (11000.0)
#define POES_LPF_ORDER   (26)
#define ARGOS_MAX_DEV    (550.0)
#define ARGOS_LOCK_ALPHA (3.1831)
#define ARGOS_ACQ_GAIN   (16)
#define ARGOS_LOCK_THRESH (0.1)
#define ARGOS_SQLCH      (0.15)
#define ARGOS_LPF_FC     (700)
#define ARGOS_LPF_ORDER  (50)
#define ARGOS_BAUD       (400*2.0)

pdto_chain *pdto_chain_new(int argos, double Fs_hz, unsigned long chunk, int force_min_L1)
{
    pdto_chain *c = (pdto_chain *)calloc(1, sizeof *c);
    if (!c) return NULL;
    c->argos = argos; c->Fs = Fs_hz; c->chunk = chunk; c->force_min_L1 = force_min_L1;
    const pdto_real Fs = (pdto_real)(unsigned int)Fs_hz;               /* main.c:346 */
    if (argos) { c->L = 1; c->N = ARGOS_LPF_ORDER; }
    else {
        c->L = (int)rint(150000.0 / Fs);                                /* main.c:347 */
        if (force_min_L1 && c->L < 1) c->L = 1;
        c->N = POES_LPF_ORDER * c->L;
    }
    const int Lb = c->L > 0 ? c->L : 1, Nb = c->N > 0 ? c->N : 1;
    /* Gardner reads its first mid-sample of every chunk up to step/2 past the end of the chunk
     * (GardenerClockRecovery.c:28); the reference's buffers are over-allocated (chunk*N,
     * main.c:356) and never written there, i.e. zero.  Keep a zero pad of at least one symbol. */
    const size_t pad = 16 + (size_t)((double)Fs * Lb / (argos ? ARGOS_BAUD : POES_BAUD));
    pdto_pll_reset(&c->pll); pdto_fir_reset(&c->fir); pdto_agc_reset(&c->agc);
    pdto_gardner_reset(&c->gardner); pdto_manchester_reset(&c->man); pdto_bytesync_reset(&c->sync);
    c->wave_time = 0; c->wave_ts = 1.0 / (pdto_real)(unsigned int)Fs_hz;   /* wave.c:96-97 */
    c->h        = (pdto_real *)calloc((size_t)Nb, sizeof(pdto_real));
    c->time_in  = (pdto_real *)calloc(chunk + pad, sizeof(pdto_real));
    c->real_s   = (pdto_real *)calloc(chunk + pad, sizeof(pdto_real));
    c->lock     = (pdto_real *)calloc(chunk + 16, sizeof(pdto_real));
    c->lpf      = (pdto_real *)calloc(chunk * (size_t)Lb + pad, sizeof(pdto_real));
    c->lpf_time = (pdto_real *)calloc(chunk * (size_t)Lb + pad, sizeof(pdto_real));
    c->sym      = (pdto_real *)calloc(chunk * (size_t)Lb + pad, sizeof(pdto_real));
    c->bits     = (unsigned char *)calloc(chunk * (size_t)Lb + pad, 1);
    if (argos) pdto_make_lpfir(c->h, c->N, ARGOS_LPF_FC, Fs, 1);                     /* ARGOS main.c:248 */
    else if (c->L > 0) pdto_make_lpfir(c->h, c->N, POES_LPF_FC, Fs * c->L, c->L);   /* POES main.c:369 */
    return c;
}

void pdto_chain_free(pdto_chain *c)
{
    if (!c) return;
    free(c->h); free(c->time_in); free(c->real_s); free(c->lock); free(c->lpf); free(c->lpf_time);
    free(c->sym); free(c->bits); pdto_bytesync_free(&c->sync); free(c);
}

static void chain_chunk(pdto_chain *c, const pdto_real *iq, unsigned long n)
{
    const pdto_real Fs = (pdto_real)(unsigned int)c->Fs;
    const uint64_t base_in = c->total_samples;
    for (unsigned long i = 0; i < n; i++) {                  /* wave.c:166-167 */
        c->wave_time += c->wave_ts;
        c->time_in[i] = c->wave_time;
    }
    if (c->chunks == 0 && c->norm_factor == 0)               /* main.c:384-389 */
        c->norm_factor = pdto_static_gain(iq, (unsigned int)n, 1.0);

    pdto_real *trp = c->tr_phase ? c->tr_phase + base_in : NULL;
    pdto_real *trf = c->tr_freq ? c->tr_freq + base_in : NULL;
    unsigned long n_sym, n_bits, n_interp;
    uint32_t *gidx = NULL; pdto_real *gerr = NULL;

    if (!c->argos) {
        const int L = c->L;
        c->avg_phase = pdto_pll_run(&c->pll, iq, c->real_s, NULL, (unsigned int)n, Fs, POES_MAX_DEV,
                                    POES_LOCK_THRESH, POES_LOCK_ALPHA * (2.0 * M_PI / Fs),
                                    POES_ACQ_GAIN * (2.0 * M_PI / Fs), POES_TRCK_GAIN * (2.0 * M_PI / Fs),
                                    trp, trf);                                   /* main.c:413 */
        if (c->tr_pll_out) memcpy(c->tr_pll_out + base_in, c->real_s, n * sizeof(pdto_real));
        n_interp = n * (unsigned long)L;
        const uint64_t base_i = base_in * (uint64_t)L;
        pdto_fir_interp_run(&c->fir, c->time_in, c->real_s, c->lpf, c->lpf_time, n, c->h, c->N, L);   /* :419 */
        if (c->tr_lpf) memcpy(c->tr_lpf + base_i, c->lpf, n_interp * sizeof(pdto_real));
        pdto_agc_run(&c->agc, c->lpf, n_interp, c->norm_factor, AGC_ATCK * (2.0 * M_PI / (Fs * L)),
                     AGC_DCY * (2.0 * M_PI / (Fs * L)), NULL);                   /* :429 */
        if (c->tr_agc) memcpy(c->tr_agc + base_i, c->lpf, n_interp * sizeof(pdto_real));
        if (c->tr_gidx) { gidx = (uint32_t *)malloc((n_interp + 16) * sizeof *gidx); gerr = (pdto_real *)malloc((n_interp + 16) * sizeof *gerr); }
        n_sym = pdto_gardner_run(&c->gardner, c->lpf, c->lpf_time, n_interp, c->sym, (int)(Fs * L), POES_BAUD,
                                 0.1, 3.0, gidx, gerr);                          /* :438 */
        for (unsigned long k = 0; k < n_sym && c->tr_sym && c->total_symbols + k < c->tr_sym_cap; k++) {
            c->tr_sym[c->total_symbols + k] = c->sym[k];
            if (gidx) { c->tr_gidx[c->total_symbols + k] = base_i + gidx[k]; c->tr_gerr[c->total_symbols + k] = gerr[k]; }
        }
        n_bits = pdto_manchester_run(&c->man, c->sym, c->lpf_time, n_sym, c->bits, 1.0);            /* :445 */
        for (unsigned long k = 0; k < n_bits && c->tr_bits && c->total_bits + k < c->tr_bits_cap; k++)
            c->tr_bits[c->total_bits + k] = c->bits[k];
        c->total_frames += (uint64_t)pdto_bytesync_poes_run(&c->sync, c->bits, c->lpf_time, n_bits,
                                                            "1110110111100010000", 19);             /* :454 */
    } else {
        c->avg_phase = pdto_pll_run(&c->pll, iq, c->real_s, c->lock, (unsigned int)n, Fs, ARGOS_MAX_DEV,
                                    ARGOS_LOCK_THRESH, ARGOS_LOCK_ALPHA * (2.0 * M_PI / Fs),
                                    ARGOS_ACQ_GAIN * (2.0 * M_PI / Fs), ARGOS_ACQ_GAIN * (2.0 * M_PI / Fs),
                                    trp, trf);                                   /* ARGOS main.c:265 */
        if (c->tr_pll_out) memcpy(c->tr_pll_out + base_in, c->real_s, n * sizeof(pdto_real));
        n_interp = n;
        pdto_fir_run(&c->fir, c->real_s, n, c->h, c->N);                          /* :268 */
        if (c->tr_lpf) memcpy(c->tr_lpf + base_in, c->real_s, n * sizeof(pdto_real));
        pdto_agc_run(&c->agc, c->real_s, n, c->norm_factor, AGC_ATCK * (2.0 * M_PI / Fs),
                     AGC_DCY * (2.0 * M_PI / Fs), NULL);                          /* :270 */
        pdto_squelch(c->real_s, c->lock, n, ARGOS_SQLCH);                         /* :276 */
        if (c->tr_agc) memcpy(c->tr_agc + base_in, c->real_s, n * sizeof(pdto_real));
        if (c->tr_gidx) { gidx = (uint32_t *)malloc((n + 16) * sizeof *gidx); gerr = (pdto_real *)malloc((n + 16) * sizeof *gerr); }
        n_sym = pdto_gardner_run(&c->gardner, c->real_s, c->time_in, n, c->sym, (int)Fs, ARGOS_BAUD, 0.1, 3.0,
                                 gidx, gerr);                                     /* :278 */
        for (unsigned long k = 0; k < n_sym && c->tr_sym && c->total_symbols + k < c->tr_sym_cap; k++) {
            c->tr_sym[c->total_symbols + k] = c->sym[k];
            if (gidx) { c->tr_gidx[c->total_symbols + k] = base_in + gidx[k]; c->tr_gerr[c->total_symbols + k] = gerr[k]; }
        }
        n_bits = pdto_manchester_run(&c->man, c->sym, c->time_in, n_sym, c->bits, 0.5);             /* :282 */
        for (unsigned long k = 0; k < n_bits && c->tr_bits && c->total_bits + k < c->tr_bits_cap; k++)
            c->tr_bits[c->total_bits + k] = c->bits[k];
        c->total_frames += (uint64_t)pdto_bytesync_argos_run(&c->sync, c->bits, c->time_in, n_bits,
                                                             "0001011110000", 13);                  /* :284 */
    }
    free(gidx); free(gerr);
    c->chunks++; c->total_samples += n; c->total_symbols += n_sym; c->total_bits += n_bits;
}

void pdto_chain_feed(pdto_chain *c, const pdto_real *iq, uint64_t n)
{
    uint64_t done = 0;
    while (done < n) {
        unsigned long take = (unsigned long)((n - done < c->chunk) ? (n - done) : c->chunk);
        chain_chunk(c, iq + 2 * done, take);
        done += take;
    }
}

const char *pdto_chain_text(const pdto_chain *c, size_t *len)
{
    if (len) *len = c->sync.text_len;
    return c->sync.text ? c->sync.text : "";
}

void pdto_pcm16_to_complex(const int16_t *pcm, uint64_t n, pdto_real *iq)
{
    const pdto_real maxsize = 32768;                          /* wave.c:116 */
    for (uint64_t i = 0; i < 2 * n; i++) iq[i] = pcm[i] / maxsize;
}

/* libm's sinf/cosf over an array: the reference calls them once per sample (CarrierTrackingPLL.c:106-107); tests compare the
 * kernels' restatement of glibc's algorithm against the real thing, float by float. */
void pdto_sincosf_array(const float *y, uint64_t n, float *s, float *c)
{
    for (uint64_t i = 0; i < n; i++) { s[i] = sinf(y[i]); c[i] = cosf(y[i]); }
}

size_t pdto_sizeof(const char *what)
{
    if (!strcmp(what, "pll")) return sizeof(pdto_pll);
    if (!strcmp(what, "fir")) return sizeof(pdto_fir);
    if (!strcmp(what, "agc")) return sizeof(pdto_agc);
    if (!strcmp(what, "gardner")) return sizeof(pdto_gardner);
    if (!strcmp(what, "mm")) return sizeof(pdto_mm);
    if (!strcmp(what, "manchester")) return sizeof(pdto_manchester);
    if (!strcmp(what, "bytesync")) return sizeof(pdto_bytesync);
    if (!strcmp(what, "chain")) return sizeof(pdto_chain);
    if (!strcmp(what, "real")) return sizeof(pdto_real);
    return 0;
}
