// pdt_chain_kernel.cuh — the fused whole-chain kernel: one CTA per IQ capture, the capture is walked in
// reference-sized chunks (the chunk length is part of the reference's numerical behaviour: Gardner works
// in chunk-relative float coordinates), all intermediate streams live in shared memory
// (or a per-CTA global workspace when the chunk does not fit).
//
// Data flow per chunk (reference: POESTIPdemod/main.c:379-454, ARGOSdemod/main.c:252-284):
//   IQ (HBM, cf32/cf64 or int16 PCM) --PLL--> R[chunk] --FIR(+×L interp)--> Y[chunk·L] --AGC(+squelch)--> Y
//   --Gardner--> symbols --Manchester--> bits --ByteSync--> frame table (HBM)
// Only the IQ read and the (tiny) frame table touch HBM: 8 B (float) / 16 B (double) / 4 B (pcm16) per sample.
#pragma once

#include "pdt_common.cuh"

namespace pdt {

struct ChainState {
    PllState        pll;
    AgcState        agc;
    GardnerState    gar;
    MMState         mm;
    ManchesterState man;
    SyncState       sync;
    real_t          norm;
    unsigned long long n_sym, n_bits, fir_j;   // fir_j: absolute index of the next input sample entering the FIR
    uint32_t        n_frames;
    int             cur_frame;                 // slot being filled (-1 none / overflow)
    real_t          avg_phase;
    unsigned long long in0;                    // live mode: input samples consumed by earlier pushes (absolute index of this push's sample 0)
    int             live_init;                 // live mode: this state has been initialised (the stream is running)
};

struct ChainArgs {
    ChainConst      cc;
    const real_t   *taps;           // [N] device
    const void     *iq;             // device base
    int             pcm16;
    unsigned long long stride;      // samples between captures
    const unsigned long long *n_samples;   // device [n_captures] or nullptr
    unsigned long long n_uniform;
    uint32_t        n_captures;
    real_t         *workspace;      // per-CTA global workspace (used when !use_smem)
    unsigned long long ws_stride;   // reals per CTA
    int             use_smem;
    pdt_capture_stats *stats;       // [n_captures]
    pdt_frame      *frames;         // [n_captures][max_frames]
    const pdt_traces *traces;       // device [n_captures] or nullptr
    ChainState     *persist;        // live mode (pdt_live_*): per-stream state carried from push to push, else nullptr.  The
                                    // workspace is then per STREAM (FIR history and the chunk buffer Gardner looks back into
                                    // survive between pushes like the reference's static buffers do) and the frame table a ring.
};

constexpr int CHAIN_THREADS = 256;
constexpr int IQ_TILE = 1024;       // samples staged per PLL tile

// reals needed: Rext[K-1+chunk] + LOCK[chunk] (ARGOS only) + Y[chunk*L+ypad] + IQ tile[2*IQ_TILE]
__host__ __device__ inline size_t chain_ws_reals(const ChainConst &cc)
{
    return (size_t)(cc.K - 1 + cc.chunk) + (cc.argos ? cc.chunk : 0) + (size_t)cc.chunk * cc.L + cc.ypad + 2 * IQ_TILE;
}

PDT_DEV void load_iq(const void *base, int pcm16, unsigned long long idx, real_t &a, real_t &b)
{
    if (pcm16) {
        const short2 v = reinterpret_cast<const short2 *>(base)[idx];
        const real_t maxsize = 32768;                 // wave.c:116,151,156: int16 / DECIMAL maxsize
        a = v.x / maxsize; b = v.y / maxsize;
    } else {
        const real_t *p = reinterpret_cast<const real_t *>(base) + 2 * idx;
        a = p[0]; b = p[1];
    }
}

// symbol -> Manchester -> ByteSync -> frame table; executed by the serial lane
PDT_DEV void consume_symbol(ChainState &st, const ChainConst &cc, real_t sym, unsigned long long abs_interp_idx,
                            pdt_frame *frames, const pdt_traces *tr)
{
    unsigned char bit;
    if (!manchester_step(st.man, sym, cc.man_thresh, bit)) return;
    if (tr && tr->bits && st.n_bits < tr->cap) tr->bits[st.n_bits] = bit;
    int emit, eol; unsigned char byte;
    const int ev = sync_step(st.sync, cc.sync, bit, emit, byte, eol);
    if (emit && st.cur_frame >= 0) {
        pdt_frame &f = frames[st.cur_frame];
        if (f.n_bytes < PDT_FRAME_MAX_BYTES) f.bytes[f.n_bytes++] = byte;
        if (eol) { f.complete = 1; st.cur_frame = -1; }
    } else if (eol) st.cur_frame = -1;
    if (ev != EV_NONE) {
        if (st.n_frames < cc.max_frames || cc.ring_frames) {
            st.cur_frame = (int)(st.n_frames % cc.max_frames);          // live mode: frame k of a stream sits in slot k mod max_frames
            pdt_frame &f = frames[st.cur_frame];
            f.sample_index = abs_interp_idx; f.bit_index = (uint32_t)st.n_bits;
            f.inverse = (ev == EV_SYNC_INV); f.complete = 0; f.pad = 0;
            f.n_bytes = (uint8_t)cc.prefix_bytes;
            if (cc.prefix_bytes) { f.bytes[0] = 0xED; f.bytes[1] = 0xE2; }
        } else st.cur_frame = -1;
        st.n_frames++;
    }
    st.n_bits++;
}

// ---------------------------------------------------------------------------------------------------
// v1 chain kernel: exact-order serial loops on lane 0, FIR on all threads.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CHAIN_THREADS) k_chain_exact(const ChainArgs args)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ ChainState st;
    __shared__ real_t taps_s[PDT_MAX_TAPS];

    const ChainConst &cc = args.cc;
    const int tid = threadIdx.x;
    const size_t y_cap = (size_t)cc.chunk * cc.L + cc.ypad;

    for (int i = tid; i < cc.N; i += CHAIN_THREADS) taps_s[i] = args.taps[i];

    for (uint32_t cap = blockIdx.x; cap < args.n_captures; cap += gridDim.x) {
        real_t *ws = args.use_smem ? reinterpret_cast<real_t *>(smem_raw)
                                   : args.workspace + (size_t)(args.persist ? cap : blockIdx.x) * args.ws_stride;
        real_t *Rext = ws;                                   // [K-1 + chunk]
        real_t *LOCK = Rext + (cc.K - 1 + cc.chunk);         // [chunk] (ARGOS)
        real_t *Y    = LOCK + (cc.argos ? cc.chunk : 0);     // [chunk*L + ypad]
        real_t *IQT  = Y + (size_t)cc.chunk * cc.L + cc.ypad;     // [2*IQ_TILE]
        const unsigned long long n = args.n_samples ? args.n_samples[cap] : args.n_uniform;
        const unsigned long long first = (unsigned long long)cap * args.stride;
        pdt_frame *frames = args.frames + (size_t)cap * cc.max_frames;
        const pdt_traces *tr = args.traces ? &args.traces[cap] : nullptr;

        __syncthreads();
        const bool resume = args.persist && args.persist[cap].live_init;    // live mode, second push onwards: carry on
        if (tid == 0) {
            if (resume) st = args.persist[cap];
            else {
                st = ChainState();
                pll_reset(st.pll);
                st.agc.gain = 1;
                st.sync.one = 1;
                st.cur_frame = -1;
                st.norm = cc.norm_override;
                st.live_init = 1;
            }
        }
        if (!resume) {
            for (size_t i = tid; i < (size_t)cc.K - 1; i += CHAIN_THREADS) Rext[i] = 0;
            for (size_t i = tid; i < y_cap; i += CHAIN_THREADS) Y[i] = 0;    // fresh zeroed buffer, like the malloc'd one
        }
        __syncthreads();
        const unsigned long long in0 = st.in0;                              // absolute index of this call's first sample (0 outside live mode)

        for (unsigned long long base = 0; base < n; base += cc.chunk) {
            const uint32_t m = (uint32_t)((n - base < cc.chunk) ? (n - base) : cc.chunk);

            // ---- StaticGain on the first chunk (main.c:384-389) --------------------------------------
            if (base == 0 && in0 == 0 && tid == 0 && st.norm == 0) {
                real_t level; real_t a, b;
                load_iq(args.iq, args.pcm16, first, a, b);
                level = hypot_exact(a, b);
                for (uint32_t i = 0; i < m; i++) {
                    load_iq(args.iq, args.pcm16, first + i, a, b);
                    level += hypot_exact(a, b);
                    level /= 2.0;
                }
                st.norm = (real_t)1.0 / level;
            }

            // ---- PLL: tiles of IQ staged cooperatively, recurrence on lane 0 -------------------------
            if (tid == 0) { pll_begin(st.pll, cc.pll); }
            for (uint32_t t0 = 0; t0 < m; t0 += IQ_TILE) {
                const uint32_t tn = (m - t0 < IQ_TILE) ? (m - t0) : IQ_TILE;
                __syncthreads();
                for (uint32_t i = tid; i < tn; i += CHAIN_THREADS) {
                    real_t a, b;
                    load_iq(args.iq, args.pcm16, first + base + t0 + i, a, b);
                    IQT[2 * i] = a; IQT[2 * i + 1] = b;
                }
                __syncthreads();
                if (tid == 0) {
                    for (uint32_t i = 0; i < tn; i++) {
                        const unsigned long long g = in0 + base + t0 + i;
                        if (tr) {
                            if (tr->pll_phase) reinterpret_cast<real_t *>(tr->pll_phase)[g] = st.pll.phase;
                            if (tr->pll_freq)  reinterpret_cast<real_t *>(tr->pll_freq)[g]  = st.pll.freq;
                        }
                        real_t out, lock;
                        pll_step(st.pll, cc.pll, IQT[2 * i], IQT[2 * i + 1], out, lock, g);
                        Rext[cc.K - 1 + t0 + i] = out;
                        if (cc.argos) LOCK[t0 + i] = lock;
                        if (tr) {
                            if (tr->pll_out) reinterpret_cast<real_t *>(tr->pll_out)[g] = out;
                            if (tr->lock)    reinterpret_cast<real_t *>(tr->lock)[g] = lock;
                        }
                    }
                    st.avg_phase = st.pll.avg_phase;
                }
            }
            __syncthreads();

            // ---- FIR (all threads), exact summation order -------------------------------------------
            const unsigned long long j0 = st.fir_j;                     // absolute index of R[0] of this chunk
            const uint32_t n_out = m * (uint32_t)cc.L;
            if (!cc.argos) {
                for (uint32_t o = tid; o < n_out; o += CHAIN_THREADS) {
                    const uint32_t jl = o / cc.L; const int p = (int)(o - jl * cc.L);
                    const int k0 = (int)((j0 + jl) % (unsigned)cc.K);
                    const real_t y = fir_interp_exact(taps_s, Rext + (cc.K - 1) + jl, cc.N, cc.L, cc.K, p, k0);
                    Y[o] = y;
                    if (tr && tr->lpf) reinterpret_cast<real_t *>(tr->lpf)[base * cc.L + o] = y;
                }
            } else {
                for (uint32_t o = tid; o < n_out; o += CHAIN_THREADS) {
                    const real_t y = fir_plain_exact(taps_s, Rext + (cc.K - 1) + o, cc.N);
                    Y[o] = y;
                    if (tr && tr->lpf) reinterpret_cast<real_t *>(tr->lpf)[base + o] = y;
                }
            }
            __syncthreads();
            // slide the FIR history: keep the last K-1 inputs
            {
                // K - 1 <= 1023 (build_chain_const) may exceed the CTA: every thread carries up to 4 elements through registers,
                // all reads before all writes (source and destination overlap when the chunk is shorter than the history)
                constexpr int SLIDE = (1024 + CHAIN_THREADS - 1) / CHAIN_THREADS;
                real_t keep[SLIDE];
#pragma unroll
                for (int q = 0; q < SLIDE; q++) { const int i = (int)tid + q * CHAIN_THREADS; keep[q] = (i < cc.K - 1) ? Rext[m + i] : (real_t)0; }
                __syncthreads();
#pragma unroll
                for (int q = 0; q < SLIDE; q++) { const int i = (int)tid + q * CHAIN_THREADS; if (i < cc.K - 1) Rext[i] = keep[q]; }
            }

            // ---- AGC (+ squelch) then Gardner -> Manchester -> ByteSync on lane 0 ---------------------
            if (tid == 0) {
                st.fir_j = j0 + m;
                if (!st.agc.init) { st.agc.init = 1; st.agc.gain = st.norm; }
                for (uint32_t o = 0; o < n_out; o++) {
                    real_t v = agc_step(st.agc, Y[o], cc.agc_attack, cc.agc_decay);
                    if (cc.argos && LOCK[o] < cc.squelch) v = 0;                // AGC.c:24-46, ARGOS main.c:276
                    Y[o] = v;
                    if (tr && tr->agc) reinterpret_cast<real_t *>(tr->agc)[base * cc.L + o] = v;
                }
                gardner_begin(st.gar, cc.gardner_fs, cc.baud);
                const unsigned long long ibase = (in0 + base) * (unsigned long long)cc.L;
                if (cc.use_mm) {
                    // MMClockRecovery.c:5-84 in the Gardner loop's place (the call the drivers keep commented out)
                    mm_begin(st.mm, cc.gardner_fs, cc.baud);
                    const real_t step_max = cc.gardner_fs / (cc.baud - cc.mm_range), step_min = cc.gardner_fs / (cc.baud + cc.mm_range);
                    while (mm_rint(st.mm.next) < n_out) {
                        real_t sym;
                        const unsigned at = mm_step(st.mm, Y, step_min, step_max, cc.mm_kp, sym);
                        if (tr && st.n_sym < tr->cap) {
                            if (tr->sym)         reinterpret_cast<real_t *>(tr->sym)[st.n_sym] = sym;
                            if (tr->gardner_idx) tr->gardner_idx[st.n_sym] = ibase + at;
                        }
                        st.n_sym++;
                        consume_symbol(st, cc, sym, ibase + at, frames, tr);
                    }
                    st.mm.next = st.mm.next - n_out;                             // :80
                } else
                while (r_rint(st.gar.next) < n_out) {
                    real_t sym, err;
                    const unsigned at = gardner_step(st.gar, Y, cc.g_range, cc.g_kp, sym, err);
                    if (tr && st.n_sym < tr->cap) {
                        if (tr->sym)         reinterpret_cast<real_t *>(tr->sym)[st.n_sym] = sym;
                        if (tr->gardner_err) reinterpret_cast<real_t *>(tr->gardner_err)[st.n_sym] = err;
                        if (tr->gardner_idx) tr->gardner_idx[st.n_sym] = ibase + at;
                    }
                    st.n_sym++;
                    consume_symbol(st, cc, sym, ibase + at, frames, tr);
                }
                if (!cc.use_mm) st.gar.next = st.gar.next - n_out;               // :111
            }
            __syncthreads();
        }

        if (tid == 0) {
            if (args.persist) { st.in0 = in0 + n; args.persist[cap] = st; }
            pdt_capture_stats s;
            s.n_samples = in0 + n; s.n_symbols = st.n_sym; s.n_bits = st.n_bits; s.n_frames = st.n_frames;
            s.locked = (st.pll.stage == 2); s.lock_sample = st.pll.lock_sample; s.lock_freq_hz = st.pll.lock_freq_hz;
            s.norm_factor = st.norm; s.avg_phase = st.avg_phase;
            s.final_phase = st.pll.phase; s.final_freq = st.pll.freq; s.final_gain = st.agc.gain; s.final_next = st.gar.next;
            s.prelocked = 0; s.prelock_snr = 0.0f;
            args.stats[cap] = s;
        }
    }
}

} // namespace pdt
