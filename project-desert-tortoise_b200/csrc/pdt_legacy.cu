// pdt_legacy.cu — the reference's own stage-function ABI (include/pdt_legacy.h) on top of sm_100a kernels.
//
// POESTIPdemod/main.c and ARGOSdemod/main.c link unmodified against these symbols.  Like the
// reference, the state of every stage is a process-wide singleton that is latched on first use; here it
// lives in device memory.  Each call stages the caller's host buffers to the GPU, runs the stage kernel
// and copies the result back before returning.  CarrierTrackPLL runs on one CTA through the block runner of the exact
// engine (pdt_pll_pipe.cuh: only the loop filter is serial), the float AGC through the proved common-regime chain of the
// batch engine; Gardner / M&M, Manchester and ByteSync run in the reference's exact operation order on one thread; the
// FIRs are data-parallel with the reference's exact (rotating) summation order.  The time-axis arrays are pure index bookkeeping and are
// compacted on the host from the pick indices the kernels return.
#include <cmath>
#include <vector>

#include "pdt_common.cuh"
#include "pdt_pll_pipe.cuh"
#if PDT_USE_FLOATS
#include "pdt_tiled.cuh"             // pll_track_step, agc_run_fast: the float-exact loop bodies the tiled engine is verified with
#endif
#include "../../include/pdt_legacy.h"

using namespace pdt;

namespace {

struct LegacyState {
    PllState pll; AgcState agc; GardnerState gar; MMState mm; ManchesterState man; SyncState sync;
    int agcc_init; real_t agcc_gain; real_t amp_avg;
    unsigned long long fir_j;          // inputs consumed by LowPassFilterInterp so far
    // per-call results
    unsigned long long count; real_t ret;
};

struct DevBuf {
    void *p = nullptr; size_t cap = 0;
    void *ensure(size_t bytes)
    {
        if (bytes > cap) {
            if (p) cudaFree(p);
            size_t want = bytes + bytes / 2 + 256;
            if (cudaMalloc(&p, want) != cudaSuccess) { printf("Error in malloc\n"); exit(1); }
            cudaMemset(p, 0, want);
            cap = want;
        }
        return p;
    }
};

struct Legacy {
    LegacyState *d_state = nullptr;
    DevBuf in, out, aux, aux2, taps, hist;
    bool ready = false;
} G;

void die(const char *what, cudaError_t e);
// [hist | n new] buffer that survives growth with its first `hist` elements intact (zero history on first use)
real_t *hist_buffer(DevBuf &b, int hist, size_t n)
{
    const size_t need = sizeof(real_t) * ((size_t)hist + n + 8);
    if (b.cap < need) {
        std::vector<real_t> keep((size_t)hist, 0);
        if (b.p && hist) cudaMemcpy(keep.data(), b.p, sizeof(real_t) * hist, cudaMemcpyDeviceToHost);
        b.ensure(need);
        if (hist) cudaMemcpy(b.p, keep.data(), sizeof(real_t) * hist, cudaMemcpyHostToDevice);
    }
    return (real_t *)b.p;
}

void die(const char *what, cudaError_t e)
{
    fprintf(stderr, "pdt: %s failed: %s (no CPU fallback; a CUDA device is required)\n", what, cudaGetErrorString(e));
    exit(1);
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) die(#call, e_); } while (0)

__global__ void k_leg_reset(LegacyState *s)
{
    *s = LegacyState();
    pll_reset(s->pll);
    s->agc.gain = 1; s->sync.one = 1; s->mm.step = 3.0;
}

void ensure_ready()
{
    if (G.ready) return;
    if (!device_ok()) { fprintf(stderr, "pdt: %s\n", pdt_last_error()); exit(1); }
    CK(cudaMalloc(&G.d_state, sizeof(LegacyState)));
    k_leg_reset<<<1, 1>>>(G.d_state); count_launch();
    CK(cudaDeviceSynchronize());
    G.ready = true;
}

LegacyState fetch_state()
{
    LegacyState h;
    CK(cudaMemcpy(&h, G.d_state, sizeof h, cudaMemcpyDeviceToHost));
    return h;
}

// ---- kernels -----------------------------------------------------------------------------------------
__global__ void k_leg_static_gain(LegacyState *s, const real_t *iq, unsigned n, real_t desired)
{
    s->ret = static_gain_serial(iq, n, desired);
}

// CarrierTrackPLL by one CTA, block by block (pdt_pll_pipe.cuh): the statements of the one-thread loop, only the loop filter
// serial; acquisition and track mode, float and double.
__global__ void __launch_bounds__(PP_THREADS) k_leg_pll_blocks(LegacyState *s, PllParams p, const real_t *__restrict__ iq, real_t *__restrict__ out,
                                                        real_t *__restrict__ lock, unsigned n)
{
    __shared__ PllState st;
    __shared__ PllPipeSmem blk;
    if (threadIdx.x == 0) {
        st = s->pll;
        st.lock_event = 0;
        pll_begin(st, p);
    }
    __syncthreads();
    pll_run_blocks(st, p, n, st.samples_seen, blk,
        [&](unsigned long long i, real_t &a, real_t &b) { a = iq[2 * i]; b = iq[2 * i + 1]; },
        [&](unsigned long long i, real_t o, real_t, real_t) { out[i] = o; },
        [&](unsigned long long i, real_t l) { if (lock) lock[i] = l; });
    if (threadIdx.x == 0) {
        st.samples_seen += n;
        s->pll = st;
        s->ret = st.avg_phase;
    }
}

#if PDT_USE_FLOATS
// NormalizingAGC of the float build: the proven fast regime of the batch engine (four dependent operations per sample and a
// side proof that no clamp / attack branch would have fired, pdt_tiled.cuh), the general recurrence if the proof fails.
__global__ void k_leg_agc_fast(LegacyState *s, const float *__restrict__ x, float *__restrict__ z, unsigned long long n, float initial,
                               float attack, float decay)
{
    using namespace pdt::tiled;
    AgcState st = s->agc;
    if (!st.init) { st.init = 1; st.gain = initial; }
    float gain = st.gain, start_gain;
    agc_tile(x, z, 0, 0, n, gain, start_gain, attack, decay);
    st.gain = gain;
    s->agc = st;
}
#endif

// xext = [K-1 history | n new inputs]; out[n*L]
__global__ void k_leg_fir_interp(const real_t *__restrict__ taps, const real_t *__restrict__ xext, real_t *__restrict__ out,
                                 unsigned long long n, int N, int L, int K, unsigned long long j0)
{
    const unsigned long long n_out = n * (unsigned long long)L;
    for (unsigned long long o = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; o < n_out;
         o += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long jl = o / L; const int p = (int)(o - jl * L);
        const int k0 = (int)((j0 + jl) % (unsigned)K);
        out[o] = fir_interp_exact(taps, xext + (K - 1) + jl, N, L, K, p, k0);
    }
}

__global__ void k_leg_fir_plain(const real_t *__restrict__ taps, const real_t *__restrict__ xext, real_t *__restrict__ out,
                                unsigned long long n, int N)
{
    for (unsigned long long o = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; o < n;
         o += (unsigned long long)gridDim.x * blockDim.x)
        out[o] = fir_plain_exact(taps, xext + (N - 1) + o, N);
}

__global__ void k_leg_slide(real_t *xext, unsigned long long n, int hist)
{
    // keep the last `hist` inputs at the front (single block; read-then-write)
    extern __shared__ unsigned char sm_raw[];
    real_t *tmp = reinterpret_cast<real_t *>(sm_raw);
    for (int i = threadIdx.x; i < hist; i += blockDim.x) tmp[i] = xext[n + i];
    __syncthreads();
    for (int i = threadIdx.x; i < hist; i += blockDim.x) xext[i] = tmp[i];
}

__global__ void k_leg_agc(LegacyState *s, real_t *x, unsigned long long n, real_t initial, real_t attack, real_t decay)
{
    AgcState st = s->agc;
    if (!st.init) { st.init = 1; st.gain = initial; }
    for (unsigned long long i = 0; i < n; i++) x[i] = agc_step(st, x[i], attack, decay);
    s->agc = st;
}

__global__ void k_leg_agcc(LegacyState *s, real_t *iq, unsigned long long n, real_t initial, real_t loop_gain)
{
    // AGC.c:164-200 NormalizingAGCC (exported; its call is commented out in both drivers)
    real_t gain = s->agcc_init ? s->agcc_gain : initial;
    const real_t desired = 5;
    for (unsigned long long i = 0; i < n; i++) {
        iq[2 * i] *= gain; iq[2 * i + 1] *= gain;                     // complex *= real
#if PDT_USE_FLOATS
        // AGC.c:188 calls fabsf() on a float complex: not a <tgmath.h> macro, the argument converts to float by dropping the
        // imaginary part — the float build measures |Re|
        const real_t mag = fabsf(iq[2 * i]);
#else
        const real_t mag = hypot_exact(iq[2 * i], iq[2 * i + 1]);     // AGC.c:190: <tgmath.h> fabs() of a complex = cabs()
#endif
        real_t err = desired - (gain * mag);
        gain = gain + loop_gain * err;
    }
    s->agcc_init = 1; s->agcc_gain = gain;
}

__global__ void k_leg_amp(LegacyState *s, const real_t *x, unsigned long long n, real_t alpha)
{
    real_t avg = s->amp_avg;                                           // AGC.c:6-20 FindSignalAmplitude
    for (unsigned long long i = 0; i < n; i++) avg = avg * (1.0 - alpha) + alpha * r_fabs(x[i]);
    s->amp_avg = avg; s->ret = avg;
}

__global__ void k_leg_squelch(real_t *x, const real_t *lock, unsigned long long n, real_t thr)
{
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x)
        if (lock[i] < thr) x[i] = 0;
}

__global__ void k_leg_gardner(LegacyState *s, const real_t *x, unsigned long long n, real_t *out, unsigned *idx,
                              int Fs, real_t baud, real_t range, real_t kp)
{
    GardnerState st = s->gar;
    gardner_begin(st, Fs, baud);
    unsigned long long count = 0;
    while (r_rint(st.next) < n) {
        real_t sym, err;
        const unsigned at = gardner_step(st, x, range, kp, sym, err);
        out[count] = sym; idx[count] = at; count++;
    }
    idx[count] = (unsigned)(r_rint(st.next));            // GardenerClockRecovery.c:65
    st.next = st.next - n;
    s->gar = st; s->count = count;
}

__global__ void k_leg_mm(LegacyState *s, const real_t *x, unsigned long long n, real_t *out, unsigned *idx,
                         int Fs, real_t baud, real_t range, real_t kp)
{
    // MMClockRecovery.c:5-84 (exported, not called by the drivers)
    MMState st = s->mm;
    const real_t step_max = Fs / (baud - range), step_min = Fs / (baud + range);
    mm_begin(st, Fs, baud);
    unsigned long long count = 0;
    while (mm_rint(st.next) < n) {
        real_t sym;
        const unsigned at = mm_step(st, x, step_min, step_max, kp, sym);
        out[count] = sym; idx[count] = at; count++;
    }
    st.next = st.next - n;
    s->mm = st; s->count = count;
}

__global__ void k_leg_manchester(LegacyState *s, const real_t *sym, unsigned long long n, unsigned char *bits,
                                 unsigned *src, real_t thresh)
{
    ManchesterState st = s->man;
    unsigned long long o = 0;
    for (unsigned long long i = 0; i < n; i++) {
        unsigned char b;
        if (manchester_step(st, sym[i], thresh, b)) { bits[o] = b; src[o] = (unsigned)i; o++; }
    }
    s->man = st; s->count = o;
}

struct SyncEvent { unsigned bit; unsigned char type, value; };   // type: 1 sync, 2 inverse sync, 3 byte, 4 byte+EOL

__global__ void k_leg_bytesync(LegacyState *s, SyncParams p, const unsigned char *bits, unsigned long long n, SyncEvent *ev)
{
    SyncState st = s->sync;
    unsigned long long ne = 0;
    for (unsigned long long i = 0; i < n; i++) {
        int emit, eol; unsigned char byte;
        const int e = sync_step(st, p, bits[i], emit, byte, eol);
        if (emit) { ev[ne].bit = (unsigned)i; ev[ne].type = eol ? 4 : 3; ev[ne].value = byte; ne++; }
        if (e)    { ev[ne].bit = (unsigned)i; ev[ne].type = (unsigned char)e; ev[ne].value = 0; ne++; }
    }
    s->sync = st; s->count = ne;
}

inline int grid_for(unsigned long long n) { return (int)std::min<unsigned long long>((n + 255) / 256 + 1, 148ull * 8); }

uint32_t word_of(const char *w, unsigned len)
{
    uint32_t v = 0;
    for (unsigned i = 0; i < len; i++) v = (v << 1) | (uint32_t)(w[i] == '1');
    return v;
}

int bytesync_common(unsigned char *bits, DECIMAL_TYPE *time, unsigned long n, char *syncWord, unsigned len, FILE *fp, int poes)
{
    ensure_ready();
    if (len < 1 || len > 31) { fprintf(stderr, "pdt: sync word length %u unsupported\n", len); exit(1); }
    SyncParams p;
    p.len = (int)len; p.word = word_of(syncWord, len); p.mask = (1u << len) - 1u;
    p.last_idx = poes ? 103 : 8; p.carry_bits = poes ? 3 : 0; p.inverse_enabled = poes ? 1 : 0;
    if (n == 0) return 0;
    unsigned char *d_bits = (unsigned char *)G.in.ensure(n);
    SyncEvent *d_ev = (SyncEvent *)G.out.ensure(sizeof(SyncEvent) * (n + n / 4 + 8));
    CK(cudaMemcpy(d_bits, bits, n, cudaMemcpyHostToDevice));
    k_leg_bytesync<<<1, 1>>>(G.d_state, p, d_bits, n, d_ev); count_launch();
    LegacyState h = fetch_state();
    std::vector<SyncEvent> ev(h.count);
    if (h.count) CK(cudaMemcpy(ev.data(), d_ev, sizeof(SyncEvent) * h.count, cudaMemcpyDeviceToHost));
    int found = 0;
    for (const SyncEvent &e : ev) {
        if (e.type == 3 || e.type == 4) {
            fprintf(fp, "%.2X ", e.value);
            if (!poes) printf("%.2X ", e.value);                       // ARGOS echoes to stdout (ByteSync.c:65)
            if (e.type == 4) { fprintf(fp, "\n"); if (!poes) printf("\n"); }
        } else {
            const double t = (double)time[e.bit];
            if (e.type == 1) { fprintf(fp, "%.5f ", t); if (!poes) printf("%.5f ", t); }
            else             { fprintf(fp, "%.5fi ", t); if (!poes) printf("\t%.5fi ", t); }
            if (poes) { fprintf(fp, "%.2X ", 0xED); fprintf(fp, "%.2X ", 0xE2); }
            found++;
        }
    }
    return found;
}

} // namespace

extern "C" {

void pdt_legacy_reset(void)
{
    ensure_ready();
    k_leg_reset<<<1, 1>>>(G.d_state); count_launch();
    CK(cudaDeviceSynchronize());
    if (G.hist.p) cudaMemset(G.hist.p, 0, G.hist.cap);
    if (G.aux2.p) cudaMemset(G.aux2.p, 0, G.aux2.cap);
}

DECIMAL_TYPE StaticGain(DECIMAL_TYPE *complexData, unsigned int nSamples, DECIMAL_TYPE desiredLevel)
{
    ensure_ready();
    const size_t bytes = sizeof(real_t) * 2 * (size_t)(nSamples ? nSamples : 1);
    real_t *d = (real_t *)G.in.ensure(bytes);
    CK(cudaMemcpy(d, complexData, bytes, cudaMemcpyHostToDevice));
    k_leg_static_gain<<<1, 1>>>(G.d_state, d, nSamples, desiredLevel); count_launch();
    return fetch_state().ret;
}

DECIMAL_TYPE CarrierTrackPLL(DECIMAL_TYPE *complexDataIn, DECIMAL_TYPE *realDataOut, DECIMAL_TYPE *lockSignalStreamOut,
                             unsigned int nSamples, DECIMAL_TYPE Fs, DECIMAL_TYPE freqRange, DECIMAL_TYPE d_lock_threshold,
                             DECIMAL_TYPE lockSigAlpha, DECIMAL_TYPE loopbw_acq, DECIMAL_TYPE loopbw_track)
{
    ensure_ready();
    PllParams p; p.Fs = Fs; p.freq_range = freqRange; p.lock_thresh = d_lock_threshold; p.lock_alpha = lockSigAlpha;
    p.bw_acq = loopbw_acq; p.bw_track = loopbw_track;
    const size_t n = nSamples;
    real_t *d_in = (real_t *)G.in.ensure(sizeof(real_t) * 2 * (n + 1));
    real_t *d_out = (real_t *)G.out.ensure(sizeof(real_t) * (n + 1));
    real_t *d_lock = lockSignalStreamOut ? (real_t *)G.aux.ensure(sizeof(real_t) * (n + 1)) : nullptr;
    if (n) CK(cudaMemcpy(d_in, complexDataIn, sizeof(real_t) * 2 * n, cudaMemcpyHostToDevice));
    k_leg_pll_blocks<<<1, PP_THREADS>>>(G.d_state, p, d_in, d_out, d_lock, nSamples);
    count_launch();
    if (n) CK(cudaMemcpy(realDataOut, d_out, sizeof(real_t) * n, cudaMemcpyDeviceToHost));
    if (n && d_lock) CK(cudaMemcpy(lockSignalStreamOut, d_lock, sizeof(real_t) * n, cudaMemcpyDeviceToHost));
    LegacyState h = fetch_state();
    if (h.pll.lock_event) printf(" : PLL locked at %0.2fHz\n", h.pll.lock_freq_hz);      // CarrierTrackingPLL.c:269
    return h.ret;
}

int MakeLPFIR(DECIMAL_TYPE *h, int N, DECIMAL_TYPE Fc, DECIMAL_TYPE Fs, int interpFactor)
{
    make_lpfir_host(h, N, Fc, Fs, interpFactor);      // one-off filter design (LowPassFilter.c:127-175), host libm like the reference
    return N;
}

void LowPassFilterInterp(DECIMAL_TYPE *inTime, DECIMAL_TYPE *in, DECIMAL_TYPE *out, DECIMAL_TYPE *outTime,
                         unsigned long nSamples, DECIMAL_TYPE *filterCoeffs, int N, int L)
{
    ensure_ready();
    if (L <= 0 || nSamples == 0) return;              // L=0: the reference's loop bound is 0 -> no output
    if (N % L || N > PDT_MAX_TAPS) { fprintf(stderr, "pdt: LowPassFilterInterp N=%d L=%d unsupported\n", N, L); exit(1); }
    const int K = N / L;
    const size_t n = nSamples, n_out = n * (size_t)L;
    real_t *xext = hist_buffer(G.hist, K - 1, n);
    real_t *d_taps = (real_t *)G.taps.ensure(sizeof(real_t) * N);
    real_t *d_out = (real_t *)G.out.ensure(sizeof(real_t) * n_out);
    CK(cudaMemcpy(d_taps, filterCoeffs, sizeof(real_t) * N, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(xext + (K - 1), in, sizeof(real_t) * n, cudaMemcpyHostToDevice));
    LegacyState h = fetch_state();
    k_leg_fir_interp<<<grid_for(n_out), 256>>>(d_taps, xext, d_out, n, N, L, K, h.fir_j); count_launch();
    k_leg_slide<<<1, 256, sizeof(real_t) * (K - 1)>>>(xext, n, K - 1); count_launch();
    h.fir_j += n;
    CK(cudaMemcpy(&G.d_state->fir_j, &h.fir_j, sizeof h.fir_j, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(out, d_out, sizeof(real_t) * n_out, cudaMemcpyDeviceToHost));
    // time axis: out_time[o] = in_time[idxI] with idxI already post-incremented (LowPassFilter.c:47,68): the NEXT
    // input's time, i.e. one-past-the-end for the last L outputs — the same out-of-bounds read the reference performs.
    if (outTime && inTime)
        for (size_t o = 0; o < n_out; o++) outTime[o] = inTime[o / L + 1];
}

void LowPassFilter(DECIMAL_TYPE *dataStream, unsigned long nSamples, DECIMAL_TYPE *filterCoeffs, int N)
{
    ensure_ready();
    if (nSamples == 0) return;
    if (N < 1 || N > PDT_MAX_TAPS) { fprintf(stderr, "pdt: LowPassFilter N=%d unsupported\n", N); exit(1); }
    const size_t n = nSamples;
    real_t *xext = hist_buffer(G.aux2, N - 1, n);
    real_t *d_taps = (real_t *)G.taps.ensure(sizeof(real_t) * N);
    real_t *d_out = (real_t *)G.out.ensure(sizeof(real_t) * n);
    CK(cudaMemcpy(d_taps, filterCoeffs, sizeof(real_t) * N, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(xext + (N - 1), dataStream, sizeof(real_t) * n, cudaMemcpyHostToDevice));
    k_leg_fir_plain<<<grid_for(n), 256>>>(d_taps, xext, d_out, n, N); count_launch();
    k_leg_slide<<<1, 256, sizeof(real_t) * (N - 1)>>>(xext, n, N - 1); count_launch();
    CK(cudaMemcpy(dataStream, d_out, sizeof(real_t) * n, cudaMemcpyDeviceToHost));
}

void NormalizingAGC(DECIMAL_TYPE *x, unsigned long nSamples, DECIMAL_TYPE initial, DECIMAL_TYPE attack, DECIMAL_TYPE decay)
{
    ensure_ready();
    const size_t n = nSamples;
    real_t *d = (real_t *)G.in.ensure(sizeof(real_t) * (n + 1));
    if (n) CK(cudaMemcpy(d, x, sizeof(real_t) * n, cudaMemcpyHostToDevice));
#if PDT_USE_FLOATS
    real_t *dz = (real_t *)G.out.ensure(sizeof(real_t) * (n + 4));
    k_leg_agc_fast<<<1, 1>>>(G.d_state, d, dz, n, initial, attack, decay); count_launch();
    if (n) CK(cudaMemcpy(x, dz, sizeof(real_t) * n, cudaMemcpyDeviceToHost));
#else
    k_leg_agc<<<1, 1>>>(G.d_state, d, n, initial, attack, decay); count_launch();
    if (n) CK(cudaMemcpy(x, d, sizeof(real_t) * n, cudaMemcpyDeviceToHost));
#endif
}

void NormalizingAGCC(DECIMAL_TYPE *iq, unsigned long nSamples, DECIMAL_TYPE initial, DECIMAL_TYPE loop_gain)
{
    ensure_ready();
    const size_t n = nSamples;
    if (!n) return;
    real_t *d = (real_t *)G.in.ensure(sizeof(real_t) * 2 * n);
    CK(cudaMemcpy(d, iq, sizeof(real_t) * 2 * n, cudaMemcpyHostToDevice));
    k_leg_agcc<<<1, 1>>>(G.d_state, d, n, initial, loop_gain); count_launch();
    CK(cudaMemcpy(iq, d, sizeof(real_t) * 2 * n, cudaMemcpyDeviceToHost));
}

DECIMAL_TYPE FindSignalAmplitude(DECIMAL_TYPE *x, unsigned long nSamples, DECIMAL_TYPE alpha)
{
    ensure_ready();
    const size_t n = nSamples;
    real_t *d = (real_t *)G.in.ensure(sizeof(real_t) * (n + 1));
    if (n) CK(cudaMemcpy(d, x, sizeof(real_t) * n, cudaMemcpyHostToDevice));
    k_leg_amp<<<1, 1>>>(G.d_state, d, n, alpha); count_launch();
    return fetch_state().ret;
}

void Squelch(DECIMAL_TYPE *x, DECIMAL_TYPE *lock, unsigned long nSamples, DECIMAL_TYPE thr)
{
    ensure_ready();
    const size_t n = nSamples;
    if (!n) return;
    real_t *d = (real_t *)G.in.ensure(sizeof(real_t) * n);
    real_t *dl = (real_t *)G.aux.ensure(sizeof(real_t) * n);
    CK(cudaMemcpy(d, x, sizeof(real_t) * n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dl, lock, sizeof(real_t) * n, cudaMemcpyHostToDevice));
    k_leg_squelch<<<grid_for(n), 256>>>(d, dl, n, thr); count_launch();
    CK(cudaMemcpy(x, d, sizeof(real_t) * n, cudaMemcpyDeviceToHost));
}

static unsigned long clock_recovery(int mm, DECIMAL_TYPE *x, DECIMAL_TYPE *time, unsigned long numSamples, DECIMAL_TYPE *out,
                                    int Fs, DECIMAL_TYPE baud, DECIMAL_TYPE range, DECIMAL_TYPE kp)
{
    ensure_ready();
    const size_t n = numSamples;
    // The reference reads dataStreamIn[rint(halfSample)] with a stale index of up to n + step/2 (GardenerClockRecovery.c:28,
    // SURVEY §5.9): stage exactly that many of the caller's elements so the same stale values are seen.
    const size_t halo = mm ? 0 : (size_t)ceil((double)Fs / (double)baud / 2.0) + 1;
    real_t *d_x = (real_t *)G.in.ensure(sizeof(real_t) * (n + halo + 1));
    real_t *d_out = (real_t *)G.out.ensure(sizeof(real_t) * (n + 8));
    unsigned *d_idx = (unsigned *)G.aux.ensure(sizeof(unsigned) * (n + 8));
    CK(cudaMemcpy(d_x, x, sizeof(real_t) * (n + halo), cudaMemcpyHostToDevice));
    if (mm) k_leg_mm<<<1, 1>>>(G.d_state, d_x, n, d_out, d_idx, Fs, baud, range, kp);
    else    k_leg_gardner<<<1, 1>>>(G.d_state, d_x, n, d_out, d_idx, Fs, baud, range, kp);
    count_launch();
    LegacyState h = fetch_state();
    const size_t cnt = h.count;
    if (cnt) CK(cudaMemcpy(out, d_out, sizeof(real_t) * cnt, cudaMemcpyDeviceToHost));
    std::vector<unsigned> idx(cnt + 1);
    CK(cudaMemcpy(idx.data(), d_idx, sizeof(unsigned) * (cnt + 1), cudaMemcpyDeviceToHost));
    if (time) {
        for (size_t k = 0; k < cnt; k++) time[k] = time[idx[k]];        // in-place compaction (:31)
        if (!mm) time[cnt] = time[idx[cnt]];                            // :65
    }
    return cnt;
}

unsigned long GardenerClockRecovery(DECIMAL_TYPE *x, DECIMAL_TYPE *time, unsigned long numSamples, DECIMAL_TYPE *out, int Fs,
                                    DECIMAL_TYPE baud, DECIMAL_TYPE stepRange, DECIMAL_TYPE kp)
{ return clock_recovery(0, x, time, numSamples, out, Fs, baud, stepRange, kp); }

unsigned long MMClockRecovery(DECIMAL_TYPE *x, DECIMAL_TYPE *time, unsigned long numSamples, DECIMAL_TYPE *out, int Fs,
                              DECIMAL_TYPE baud, DECIMAL_TYPE stepRange, DECIMAL_TYPE kp)
{ return clock_recovery(1, x, time, numSamples, out, Fs, baud, stepRange, kp); }

unsigned long ManchesterDecode(DECIMAL_TYPE *sym, DECIMAL_TYPE *time, unsigned long nSymbols, unsigned char *bitStream,
                               DECIMAL_TYPE resyncThreshold)
{
    ensure_ready();
    const size_t n = nSymbols;
    if (!n) return 0;
    real_t *d_sym = (real_t *)G.in.ensure(sizeof(real_t) * n);
    unsigned char *d_bits = (unsigned char *)G.out.ensure(n + 8);
    unsigned *d_src = (unsigned *)G.aux.ensure(sizeof(unsigned) * (n + 8));
    CK(cudaMemcpy(d_sym, sym, sizeof(real_t) * n, cudaMemcpyHostToDevice));
    k_leg_manchester<<<1, 1>>>(G.d_state, d_sym, n, d_bits, d_src, resyncThreshold); count_launch();
    LegacyState h = fetch_state();
    const size_t cnt = h.count;
    if (cnt) CK(cudaMemcpy(bitStream, d_bits, cnt, cudaMemcpyDeviceToHost));
    if (time && cnt) {
        std::vector<unsigned> src(cnt);
        CK(cudaMemcpy(src.data(), d_src, sizeof(unsigned) * cnt, cudaMemcpyDeviceToHost));
        for (size_t k = 0; k < cnt; k++) time[k] = time[src[k]];        // ManchesterDecode.c:86
    }
    return cnt;
}

int ByteSyncOnSyncword(unsigned char *bits, DECIMAL_TYPE *time, unsigned long n, char *syncWord, unsigned int len, FILE *fp)
{ return bytesync_common(bits, time, n, syncWord, len, fp, 1); }

int FindSyncWords(unsigned char *bits, DECIMAL_TYPE *time, unsigned long n, char *syncWord, unsigned int len, FILE *fp)
{ return bytesync_common(bits, time, n, syncWord, len, fp, 0); }

// Scalar helpers the reference library also exports (CarrierTrackingPLL.c:15-52, MMClockRecovery.c:86-89).
// They are not on the per-sample path of the shim (the kernels carry their own device versions); kept for link compatibility.
DECIMAL_TYPE arctan2(DECIMAL_TYPE y, DECIMAL_TYPE x)
{
#if PDT_USE_FLOATS
    DECIMAL_TYPE abs_y = fabsf(y) + 1e-10;
#else
    DECIMAL_TYPE abs_y = fabs(y) + 1e-10;
#endif
    DECIMAL_TYPE r, angle;
    if (x >= 0) { r = (x - abs_y) / (x + abs_y); angle = 0.78539816339744825 - 0.78539816339744825 * r; }
    else        { r = (x + abs_y) / (abs_y - x); angle = 2.35619449019234475 - 0.78539816339744825 * r; }
    return (y < 0) ? -angle : angle;
}

float Q_rsqrt(float x)
{
    float half = 0.5f * x; int bits;
    memcpy(&bits, &x, 4); bits = 0x5f3759df - (bits >> 1); memcpy(&x, &bits, 4);
    x = x * (1.5f - half * x * x);
    x = x * (1.5f - half * x * x);
    return x;
}

int sign(DECIMAL_TYPE x) { return (x > 0) - (x < 0); }

} // extern "C"
