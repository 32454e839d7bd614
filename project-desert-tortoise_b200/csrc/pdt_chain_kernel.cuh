// pdt_chain_kernel.cuh — the fused whole-chain kernel: one CTA per IQ capture, the capture is walked in
// reference-sized chunks (the chunk length is part of the reference's numerical behaviour: Gardner works
// in chunk-relative float coordinates).  The chunk-sized intermediate streams live in a per-CTA
// workspace — global memory (L1/L2-resident, touched by the parallel phases only) or shared memory, whichever keeps
// more CTAs resident (pdt_create decides) — and every serial phase works on a window of it staged in shared memory.
//
// Data flow per chunk (reference: POESTIPdemod/main.c:379-454, ARGOSdemod/main.c:252-284):
//   IQ (HBM, cf32/cf64 or int16 PCM) --PLL--> R[chunk] --FIR(+×L interp)--> Y[chunk·L] --AGC(+squelch)--> Y
//   --Gardner--> symbols --Manchester--> bits --ByteSync--> frame table (HBM)
//   PLL: the whole CTA, in speculated and verified blocks (pdt_pll_pipe.cuh); FIR: all threads; StaticGain, AGC: one thread
//   over staged windows (magnitudes / squelch by all threads); clock recovery -> bits: one thread, state in registers.
// HBM sees the IQ read and the (tiny) frame table: 8 B (float) / 16 B (double) / 4 B (pcm16) per sample.  The kernel is
// latency-bound by design (a handful of dependent chains per capture): CTAs are narrow and many.
#pragma once

#include "pdt_common.cuh"
#include "pdt_pll_pipe.cuh"

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

constexpr int CHAIN_THREADS = 128;      // narrow CTAs: the chain is latency-bound on a few serial threads, throughput comes from resident CTAs
constexpr int CHAIN_MIN_CTAS = 6;       // register budget asked of the compiler (65536 / (128 · 6) = 85 per thread)
static_assert(CHAIN_THREADS == PP_THREADS, "the PLL block runner assigns its roles by thread index");

// reals needed: Rext[K-1+chunk] + LOCK[chunk] (ARGOS only) + Y[chunk*L+ypad]
__host__ __device__ inline size_t chain_ws_reals(const ChainConst &cc)
{
    return (size_t)(cc.K - 1 + cc.chunk) + (cc.argos ? cc.chunk : 0) + (size_t)cc.chunk * cc.L + cc.ypad;
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

// What clock recovery -> Manchester -> ByteSync carry from symbol to symbol: kept in REGISTERS of the serial lane while a
// chunk is walked (every field of the shared ChainState costs a 29-cycle LDS on the dependent path), stored back per chunk.
struct BackState {
    GardnerState gar; MMState mm; ManchesterState man; SyncState sync;
    unsigned long long n_sym, n_bits;
    uint32_t n_frames; int cur_frame;
    uint32_t cur_bytes;                        // bytes already in the open frame (mirror of frames[cur_frame].n_bytes)
};
PDT_DEV void back_load(BackState &b, const ChainState &st, const pdt_frame *frames)
{
    b.gar = st.gar; b.mm = st.mm; b.man = st.man; b.sync = st.sync;
    b.n_sym = st.n_sym; b.n_bits = st.n_bits; b.n_frames = st.n_frames; b.cur_frame = st.cur_frame;
    b.cur_bytes = (st.cur_frame >= 0) ? frames[st.cur_frame].n_bytes : 0u;
}
PDT_DEV void back_store(ChainState &st, const BackState &b)
{
    st.gar = b.gar; st.mm = b.mm; st.man = b.man; st.sync = b.sync;
    st.n_sym = b.n_sym; st.n_bits = b.n_bits; st.n_frames = b.n_frames; st.cur_frame = b.cur_frame;
}

// symbol -> Manchester -> ByteSync -> frame table; executed by the serial lane
PDT_DEV void consume_symbol(BackState &st, const ChainConst &cc, real_t sym, unsigned long long abs_interp_idx,
                            pdt_frame *frames, const pdt_traces *tr)
{
    unsigned char bit;
    if (!manchester_step(st.man, sym, cc.man_thresh, bit)) return;
    if (tr && tr->bits && st.n_bits < tr->cap) tr->bits[st.n_bits] = bit;
    int emit, eol; unsigned char byte;
    const int ev = sync_step(st.sync, cc.sync, bit, emit, byte, eol);
    if (emit && st.cur_frame >= 0) {
        pdt_frame &f = frames[st.cur_frame];
        if (st.cur_bytes < PDT_FRAME_MAX_BYTES) { f.bytes[st.cur_bytes++] = byte; f.n_bytes = (uint8_t)st.cur_bytes; }
        if (eol) { f.complete = 1; st.cur_frame = -1; }
    } else if (eol) st.cur_frame = -1;
    if (ev != EV_NONE) {
        if (st.n_frames < cc.max_frames || cc.ring_frames) {
            st.cur_frame = (int)(st.n_frames % cc.max_frames);          // live mode: frame k of a stream sits in slot k mod max_frames
            pdt_frame &f = frames[st.cur_frame];
            f.sample_index = abs_interp_idx; f.bit_index = (uint32_t)st.n_bits;
            f.inverse = (ev == EV_SYNC_INV); f.complete = 0; f.pad = 0;
            f.n_bytes = (uint8_t)cc.prefix_bytes; st.cur_bytes = (uint32_t)cc.prefix_bytes;
            if (cc.prefix_bytes) { f.bytes[0] = 0xED; f.bytes[1] = 0xE2; }
        } else st.cur_frame = -1;
        st.n_frames++;
    }
    st.n_bits++;
}

// The PLL block arrays double as the staging window of the other serial stages (they never run at the same time).
constexpr int WS_REALS = (int)(sizeof(PllPipeSmem) / sizeof(real_t)) & ~3;
constexpr int CLK_BACK = 512, CLK_WIN = 1536;            // clock recovery: look-back kept in front of every window
constexpr int CLK_Q = 256;                               // symbols per queue (two queues: values as reals, pick indices as uint32)
static_assert(CLK_BACK + CLK_WIN + 2 * CLK_Q + (2 * CLK_Q * 4 + sizeof(real_t) - 1) / sizeof(real_t) <= WS_REALS,
              "clock-recovery window and symbol queues must fit the shared staging area");
struct ClkQueueCtl { int cnt[2], more[2]; };

// chunk samples [lo, hi) from shared memory, everything else from the chunk buffer
struct WindowView {
    const real_t *win, *chunk; uint32_t lo, hi;
    PDT_DEV real_t operator[](unsigned i) const { return (i >= lo && i < hi) ? win[i - lo] : chunk[i]; }
};

// cycle accounting of the exact engine (thread 0 of every CTA, summed over captures; read with pdt_debug_chain_prof):
// [0] StaticGain  [1] PLL total  [2] its P alone (prologue, restarts)  [3] C ‖ E  [4] H ‖ P  [5] C busy  [6] E busy  [7] PLL blocks  [8] contradicted blocks
// [9] FIR + history slide  [10] AGC (+squelch)  [11] clock recovery + Manchester + ByteSync  [12] whole capture  [13] samples
// [14] PLL control (commit / roll back between blocks)  [15] its work alone  [16] lock-EMA thread busy
__device__ unsigned long long g_chain_prof[20];

// ---------------------------------------------------------------------------------------------------
// the exact engine's kernel: one CTA walks one capture (or live stream) chunk by chunk
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CHAIN_THREADS, CHAIN_MIN_CTAS) k_chain_exact(const ChainArgs args)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ ChainState st;
    __shared__ real_t taps_s[PDT_MAX_TAPS];
    __shared__ __align__(16) PllPipeSmem pll_blk;
    __shared__ ClkQueueCtl clk_q;
    real_t *const WS = reinterpret_cast<real_t *>(&pll_blk);

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
        const bool need_norm = st.norm == 0;                                // no StaticGain override: measure it on the first chunk
        unsigned long long pf[15] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // thread 0: [0..9] PLL phases/blocks, [10] gain [11] pll [12] fir [13] agc [14] back
        const long long pf_begin = clock64();

        for (unsigned long long base = 0; base < n; base += cc.chunk) {
            const uint32_t m = (uint32_t)((n - base < cc.chunk) ? (n - base) : cc.chunk);

            long long pf_t = clock64();
            // ---- StaticGain on the first chunk (main.c:384-389): magnitudes by all threads, the halving average on one -------
            if (base == 0 && in0 == 0 && need_norm) {
                real_t level = 0;
                for (uint32_t t0 = 0; t0 < m; t0 += WS_REALS) {
                    const uint32_t tn = (m - t0 < (uint32_t)WS_REALS) ? (m - t0) : (uint32_t)WS_REALS;
                    for (uint32_t i = tid; i < tn; i += CHAIN_THREADS) {
                        real_t a, b;
                        load_iq(args.iq, args.pcm16, first + t0 + i, a, b);
                        WS[i] = hypot_exact(a, b);
                    }
                    __syncthreads();
                    if (tid == 0) {
                        if (t0 == 0) level = WS[0];                      // AGC.c:58: the first sample primes the average (and is added again)
                        for (uint32_t i = 0; i < tn; i++) {
                            level += WS[i];
                            level /= 2.0;
                        }
                    }
                    __syncthreads();
                }
                if (tid == 0) st.norm = (real_t)1.0 / level;
            }

            // ---- PLL: the whole CTA, block by block (pdt_pll_pipe.cuh); only the loop filter runs on one thread ----------
            { const long long now = clock64(); pf[10] += now - pf_t; pf_t = now; }
            if (tid == 0) { pll_begin(st.pll, cc.pll); }
            __syncthreads();
            {
                const unsigned long long g0 = in0 + base, src0 = first + base;
                pll_run_blocks(st.pll, cc.pll, m, g0, pll_blk,
                    [&](unsigned long long i, real_t &a, real_t &b) { load_iq(args.iq, args.pcm16, src0 + i, a, b); },
                    [&](unsigned long long i, real_t out, real_t ph_before, real_t fq_before) {
                        Rext[cc.K - 1 + i] = out;
                        if (tr) {
                            const unsigned long long g = g0 + i;
                            if (tr->pll_phase) reinterpret_cast<real_t *>(tr->pll_phase)[g] = ph_before;
                            if (tr->pll_freq)  reinterpret_cast<real_t *>(tr->pll_freq)[g]  = fq_before;
                            if (tr->pll_out)   reinterpret_cast<real_t *>(tr->pll_out)[g] = out;
                        }
                    },
                    [&](unsigned long long i, real_t lock) {
                        if (cc.argos) LOCK[i] = lock;
                        if (tr && tr->lock) reinterpret_cast<real_t *>(tr->lock)[g0 + i] = lock;
                    }, pf);
                if (tid == 0) st.avg_phase = st.pll.avg_phase;
            }
            __syncthreads();
            { const long long now = clock64(); pf[11] += now - pf_t; pf_t = now; }

            // ---- FIR (all threads), exact summation order -------------------------------------------
            const unsigned long long j0 = st.fir_j;                     // absolute index of R[0] of this chunk
            const uint32_t n_out = m * (uint32_t)cc.L;
            // tiles of inputs (with their K-1 samples of history in front) pass through the shared staging area: the chunk buffer
            // itself may live in global memory, and every output reads K inputs
            const uint32_t fir_tile = (uint32_t)(WS_REALS - (cc.K - 1)) & ~127u;          // K - 1 <= 1023 < WS_REALS
            for (uint32_t t0 = 0; t0 < m; t0 += fir_tile) {
                const uint32_t tn = (m - t0 < fir_tile) ? (m - t0) : fir_tile;
                for (uint32_t i = tid; i < tn + (uint32_t)cc.K - 1; i += CHAIN_THREADS) WS[i] = Rext[t0 + i];
                __syncthreads();
                const uint32_t o_lo = t0 * (uint32_t)cc.L, o_hi = (t0 + tn) * (uint32_t)cc.L;
                if (!cc.argos) {
                    for (uint32_t o = o_lo + tid; o < o_hi; o += CHAIN_THREADS) {
                        const uint32_t jl = o / cc.L; const int p = (int)(o - jl * cc.L);
                        const int k0 = (int)((j0 + jl) % (unsigned)cc.K);
                        const real_t y = fir_interp_exact(taps_s, WS + (cc.K - 1) + (jl - t0), cc.N, cc.L, cc.K, p, k0);
                        Y[o] = y;
                        if (tr && tr->lpf) reinterpret_cast<real_t *>(tr->lpf)[base * cc.L + o] = y;
                    }
                } else {
                    for (uint32_t o = o_lo + tid; o < o_hi; o += CHAIN_THREADS) {
                        const real_t y = fir_plain_exact(taps_s, WS + (cc.K - 1) + (o - t0), cc.N);
                        Y[o] = y;
                        if (tr && tr->lpf) reinterpret_cast<real_t *>(tr->lpf)[base + o] = y;
                    }
                }
                __syncthreads();
            }
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

            // ---- AGC (+ squelch): windows of the chunk staged in shared memory, the gain recurrence on one thread ------------
            { const long long now = clock64(); pf[12] += now - pf_t; pf_t = now; }
            if (tid == 0) {
                st.fir_j = j0 + m;
                if (!st.agc.init) { st.agc.init = 1; st.agc.gain = st.norm; }
            }
            for (uint32_t w0 = 0; w0 < n_out; w0 += WS_REALS) {
                const uint32_t wn = (n_out - w0 < (uint32_t)WS_REALS) ? (n_out - w0) : (uint32_t)WS_REALS;
                for (uint32_t k = tid; k < wn; k += CHAIN_THREADS) WS[k] = Y[w0 + k];
                __syncthreads();
                if (tid == 0) {
                    AgcState ag = st.agc;
                    uint32_t k = 0;
                    for (; k + 4 <= wn; k += 4) {                        // inputs fetched ahead of the dependent gain chain
                        const real_t x0 = WS[k], x1 = WS[k + 1], x2 = WS[k + 2], x3 = WS[k + 3];
                        real_t v0, v1, v2, v3;
                        agc_step4(ag, x0, x1, x2, x3, cc.agc_attack, cc.agc_decay, v0, v1, v2, v3);   // common-regime chain, verified
                        WS[k] = v0; WS[k + 1] = v1; WS[k + 2] = v2; WS[k + 3] = v3;
                    }
                    for (; k < wn; k++) WS[k] = agc_step(ag, WS[k], cc.agc_attack, cc.agc_decay);
                    st.agc = ag;
                }
                __syncthreads();
                for (uint32_t k = tid; k < wn; k += CHAIN_THREADS) {
                    real_t v = WS[k];
                    if (cc.argos && LOCK[w0 + k] < cc.squelch) v = 0;            // AGC.c:24-46, ARGOS main.c:276
                    Y[w0 + k] = v;
                    if (tr && tr->agc) reinterpret_cast<real_t *>(tr->agc)[base * cc.L + w0 + k] = v;
                }
                __syncthreads();
            }

            // ---- clock recovery (thread 0) -> symbol queue -> Manchester + ByteSync (thread 32).  The chunk passes through shared
            //      memory in windows with CLK_BACK samples of look-back (anything outside the window, e.g. the stale half-way index
            //      the reference carries across a chunk boundary, is read from the chunk buffer itself); thread 0 fills one
            //      symbol queue while thread 32 drains the other ----------------------------------------------------------------
            { const long long now = clock64(); pf[13] += now - pf_t; pf_t = now; }
            BackState bk;                                   // thread 0 owns gar / mm / n_sym, thread 32 the rest
            if (tid == 0 || tid == 32) back_load(bk, st, frames);
            if (tid == 0) {
                gardner_begin(bk.gar, cc.gardner_fs, cc.baud);
                if (cc.use_mm) mm_begin(bk.mm, cc.gardner_fs, cc.baud);              // MMClockRecovery.c:5-84 (the call the drivers keep commented out)
            }
            {
                const unsigned long long ibase = (in0 + base) * (unsigned long long)cc.L;
                const real_t step_max = cc.gardner_fs / (cc.baud - cc.mm_range), step_min = cc.gardner_fs / (cc.baud + cc.mm_range);
                real_t   *const q_sym = WS + CLK_BACK + CLK_WIN;                                       // [2][CLK_Q]
                uint32_t *const q_at  = reinterpret_cast<uint32_t *>(q_sym + 2 * CLK_Q);               // [2][CLK_Q]
                int fill = 0;                               // queue being filled; the other one (if `pending`) is being drained
                bool pending = false;
                auto drain = [&](int qi) {                  // thread 32
                    const int cnt = clk_q.cnt[qi];
                    const real_t *qs = q_sym + qi * CLK_Q; const uint32_t *qa = q_at + qi * CLK_Q;
                    for (int k = 0; k < cnt; k++) consume_symbol(bk, cc, qs[k], ibase + qa[k], frames, tr);
                };
                for (uint32_t w0 = 0; w0 < n_out; w0 += CLK_WIN) {
                    const uint32_t w_hi = (n_out - w0 < (uint32_t)CLK_WIN) ? n_out : w0 + CLK_WIN;
                    const uint32_t w_lo = (w0 > (uint32_t)CLK_BACK) ? w0 - CLK_BACK : 0u;
                    __syncthreads();
                    for (uint32_t k = w_lo + tid; k < w_hi; k += CHAIN_THREADS) WS[k - w_lo] = Y[k];
                    __syncthreads();
                    bool more = true;
                    while (more) {                          // one trip per window unless a queue fills up (very short symbols)
                        if (tid == 0) {
                            const WindowView yv{WS, Y, w_lo, w_hi};
                            real_t *qs = q_sym + fill * CLK_Q; uint32_t *qa = q_at + fill * CLK_Q;
                            int cnt = 0;
                            if (cc.use_mm) {
                                while (cnt < CLK_Q && mm_rint(bk.mm.next) < w_hi) {
                                    real_t sym;
                                    const unsigned at = mm_step(bk.mm, yv, step_min, step_max, cc.mm_kp, sym);
                                    if (tr && bk.n_sym < tr->cap) {
                                        if (tr->sym)         reinterpret_cast<real_t *>(tr->sym)[bk.n_sym] = sym;
                                        if (tr->gardner_idx) tr->gardner_idx[bk.n_sym] = ibase + at;
                                    }
                                    bk.n_sym++;
                                    qs[cnt] = sym; qa[cnt] = at; cnt++;
                                }
                                clk_q.more[fill] = mm_rint(bk.mm.next) < w_hi;
                            } else {
                                while (cnt < CLK_Q && r_rint(bk.gar.next) < w_hi) {
                                    real_t sym, err;
                                    const unsigned at = gardner_step(bk.gar, yv, cc.g_range, cc.g_kp, sym, err);
                                    if (tr && bk.n_sym < tr->cap) {
                                        if (tr->sym)         reinterpret_cast<real_t *>(tr->sym)[bk.n_sym] = sym;
                                        if (tr->gardner_err) reinterpret_cast<real_t *>(tr->gardner_err)[bk.n_sym] = err;
                                        if (tr->gardner_idx) tr->gardner_idx[bk.n_sym] = ibase + at;
                                    }
                                    bk.n_sym++;
                                    qs[cnt] = sym; qa[cnt] = at; cnt++;
                                }
                                clk_q.more[fill] = r_rint(bk.gar.next) < w_hi;
                            }
                            clk_q.cnt[fill] = cnt;
                        } else if (tid == 32 && pending) drain(fill ^ 1);
                        __syncthreads();
                        more = clk_q.more[fill] != 0;       // (rewritten two barriers from here at the earliest)
                        pending = true; fill ^= 1;
                    }
                }
                if (tid == 32 && pending) drain(fill ^ 1);
                if (tid == 0) {
                    if (cc.use_mm) bk.mm.next = bk.mm.next - n_out;                  // :80
                    else           bk.gar.next = bk.gar.next - n_out;                // :111
                    st.gar = bk.gar; st.mm = bk.mm; st.n_sym = bk.n_sym;
                } else if (tid == 32) {
                    st.man = bk.man; st.sync = bk.sync; st.n_bits = bk.n_bits; st.n_frames = bk.n_frames; st.cur_frame = bk.cur_frame;
                }
            }
            __syncthreads();
            if (tid == 0) pf[14] += clock64() - pf_t;
        }
        if (tid == 0) {
            atomicAdd(&g_chain_prof[0], pf[10]); atomicAdd(&g_chain_prof[1], pf[11]);
            for (int q = 0; q < 7; q++) atomicAdd(&g_chain_prof[2 + q], pf[q]);
            atomicAdd(&g_chain_prof[9], pf[12]); atomicAdd(&g_chain_prof[10], pf[13]); atomicAdd(&g_chain_prof[11], pf[14]);
            atomicAdd(&g_chain_prof[14], pf[7]); atomicAdd(&g_chain_prof[15], pf[8]); atomicAdd(&g_chain_prof[16], pf[9]);
            atomicAdd(&g_chain_prof[12], (unsigned long long)(clock64() - pf_begin)); atomicAdd(&g_chain_prof[13], n);
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
