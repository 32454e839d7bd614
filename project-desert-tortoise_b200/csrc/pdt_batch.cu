// pdt_batch.cu — host side of the batch / context C-ABI (include/pdt.h).
//
// Mirrors, for whole captures, the driver loops of POESTIPdemod/main.c:346-482 and ARGOSdemod/main.c:244-300:
// parameters are derived exactly as the reference call sites derive them (double expression narrowed to
// DECIMAL_TYPE), the filter is designed once on the host (LowPassFilter.c:127-175), and the per-sample work
// runs in ONE fused sm_100a kernel per batch (pdt_chain_kernel.cuh).  No CPU fallback exists.
#include <cmath>
#include <vector>
#include <algorithm>
#include <new>

#include "pdt_chain_kernel.cuh"
#include "pdt_synth.cuh"
#if PDT_USE_FLOATS
#include "pdt_tiled_kernels.cuh"
#endif

namespace pdt {

thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

bool device_ok()
{
    static int state = -1;
    if (state < 0) {
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        state = (e == cudaSuccess && n > 0) ? 1 : 0;
        if (!state) { fail(PDT_ENODEV, "no usable CUDA device (%s)", cudaGetErrorString(e)); cudaGetLastError(); }
    }
    return state == 1;
}

// LowPassFilter.c:127-175 — run once per context on the host with the host libm, exactly as the reference does.
void make_lpfir_host(real_t *h, int N, real_t Fc, real_t Fs, int L)
{
    real_t T = 1.0 / Fs;
    real_t wc = 2.0 * M_PI * Fc * T;
    real_t tou = (N - 1.0) / 2.0;
    for (int n = 0; n < N; n++) {
#if PDT_USE_FLOATS
        real_t hd = (sinf(wc * (n - tou))) / (M_PI * (n - tou));
#else
        real_t hd = (sin(wc * (n - tou))) / (M_PI * (n - tou));
#endif
        if ((n == tou) && ((int)((N / 2) * 2) != N)) hd = wc / M_PI;
        real_t wn = 0.42 - 0.5 * cos((2 * M_PI * n) / (N - 1)) + 0.08 * cos((4 * M_PI * n) / (N - 1));
        h[n] = hd * wn * (real_t)(L);
    }
}

static uint32_t sync_bits(const char *w, int len)
{
    uint32_t v = 0;
    for (int i = 0; i < len; i++) v = (v << 1) | (uint32_t)(w[i] == '1');
    return v;
}

int build_chain_const(const pdt_params &p, ChainConst &cc)
{
    memset(&cc, 0, sizeof cc);
    if (p.sync_len < 1 || p.sync_len > 31) return fail(PDT_EINVAL, "sync_len %d out of range", p.sync_len);
    if (p.chunk < 64) return fail(PDT_EINVAL, "chunk %u too small", p.chunk);
    const real_t Fs = (real_t)(unsigned int)p.sample_rate;          // main.c:346 (header.sample_rate is unsigned)
    cc.argos = (p.mode == PDT_MODE_ARGOS);
    cc.L = p.interp; cc.N = p.taps;
    if (cc.L < 0 || cc.N > PDT_MAX_TAPS) return fail(PDT_EINVAL, "interp %d / taps %d unsupported", cc.L, cc.N);
    if (cc.L > 0) {
        if (cc.N % cc.L) return fail(PDT_EINVAL, "taps must be a multiple of interp");
        cc.K = cc.argos ? cc.N : cc.N / cc.L;
        if (cc.K < 1 || cc.K > 1024) return fail(PDT_EINVAL, "taps per branch %d unsupported", cc.K);
    }
    cc.chunk = p.chunk;
    const double w = 2.0 * M_PI / Fs;                                // (2.0*M_PI/Fs), Fs is DECIMAL_TYPE
    cc.pll.Fs = Fs;
    cc.pll.freq_range = p.max_carrier_dev;
    cc.pll.lock_thresh = p.pll_lock_thresh;
    cc.pll.lock_alpha = p.pll_lock_alpha * w;                        // main.c:413
    cc.pll.bw_acq = p.pll_acq_gain * w;
    cc.pll.bw_track = p.pll_track_gain * w;
    if (cc.argos) {
        cc.agc_attack = p.agc_attack * (2.0 * M_PI / Fs);            // ARGOS main.c:270
        cc.agc_decay  = p.agc_decay * (2.0 * M_PI / Fs);
        cc.gardner_fs = (int)Fs;                                     // ARGOS main.c:278
    } else {
        const real_t FsL = Fs * cc.L;                                // Fs*dspLPFInterp in DECIMAL_TYPE
        cc.agc_attack = p.agc_attack * (2.0 * M_PI / FsL);           // main.c:429
        cc.agc_decay  = p.agc_decay * (2.0 * M_PI / FsL);
        cc.gardner_fs = (int)FsL;                                    // main.c:438
    }
    cc.norm_override = p.norm_factor;
    cc.baud = p.baud; cc.g_range = p.gardner_err_lim; cc.g_kp = p.gardner_gain;
    cc.man_thresh = p.manchester_resync; cc.squelch = p.squelch_thresh;
    cc.sync.len = p.sync_len;
    cc.sync.word = sync_bits(p.sync_word, p.sync_len);
    cc.sync.mask = (p.sync_len == 32) ? 0xffffffffu : ((1u << p.sync_len) - 1u);
    cc.sync.last_idx = cc.argos ? 8 : 103;
    cc.sync.carry_bits = cc.argos ? 0 : 3;
    cc.sync.inverse_enabled = cc.argos ? 0 : 1;
    cc.prefix_bytes = cc.argos ? 0 : 2;
    if (p.sync_generic) {                                            // common/ByteSync.c:16-144
        if (p.sync_frame_len < 2 || p.sync_frame_len > PDT_FRAME_MAX_BYTES - 1 || p.sync_start_bit < 0 || p.sync_start_bit > 7)
            return fail(PDT_EINVAL, "generic sync: frame length %d / start bit %d out of range", p.sync_frame_len, p.sync_start_bit);
        cc.sync.last_idx = p.sync_frame_len; cc.sync.carry_bits = p.sync_start_bit;
        cc.sync.inverse_enabled = 1; cc.prefix_bytes = 2;
    }
    cc.use_mm = (p.clock_recovery == PDT_CLOCK_MM);
    cc.mm_range = p.mm_step_range; cc.mm_kp = p.mm_gain;
    if (cc.use_mm && !(p.baud - p.mm_step_range > 0)) return fail(PDT_EINVAL, "M&M step range %g >= baud", p.mm_step_range);
    cc.ypad = 16 + 4 * (int)((double)cc.gardner_fs / p.baud / 4.0 + 1.0);
    return PDT_OK;
}

} // namespace pdt

using namespace pdt;

struct pdt_ctx {
    pdt_params  params;
    ChainConst  cc;
    uint32_t    max_captures, max_frames;
    uint64_t    max_samples;
    real_t      taps_h[PDT_MAX_TAPS];
    real_t     *d_taps = nullptr;
    real_t     *d_ws = nullptr;          // global workspace (only when the chunk does not fit in shared memory)
    size_t      ws_stride = 0;
    int         use_smem = 1;
    size_t      smem_bytes = 0;
    int         grid = 0;
    pdt_capture_stats *d_stats = nullptr;
    pdt_frame  *d_frames = nullptr;
    unsigned long long *d_nsamp = nullptr;
    pdt_traces *d_traces = nullptr;
    pdt_frame_quality *d_quality = nullptr;
    ChainState *d_live = nullptr;        // live mode (pdt_live_*): per-stream chain state carried from push to push
    real_t     *d_live_ws = nullptr;     //   and per-stream workspaces (FIR history, chunk buffer)
    size_t      live_ws_stride = 0;
    void       *d_stage = nullptr;       // staging for pdt_demod_host
    size_t      stage_bytes = 0;
    int         device = 0, sm_count = 0;
    int         engine = PDT_ENGINE_EXACT;
    int         max_groups = 0;          // pdt_set_groups: 0 = default (PDT_MAX_GROUPS or 3)
    static constexpr int MAX_GROUPS = 5;      // upper bound; the default in use is 3 (group_plan): with 4 batches in flight, 4·(3 + slow) + callers stay within the 32 hardware queues
    cudaStream_t h2d_stream = nullptr;       // chunked staging of pdt_demod_host
    cudaEvent_t  ev_h2d[MAX_GROUPS] = {}, ev_done = nullptr;
#if PDT_USE_FLOATS
    uint32_t prelock_from = 0xFFFFFFFFu; // this call: captures >= this index start pre-locked (pdt_demod_segments_device)
    tiled::TiledArgs ta;                 // workspace pointers + plans of the tiled engine (engine == PDT_ENGINE_TILED)
    tiled::TapsRev   taps_rev;
    tiled::AcqResult *acq_backup = nullptr;   // PDT_DEBUG_ONLY_ACQ1 (timing experiments): acquisition results of the first batch
    int         acq_packed = 1;          // k_acquire_packed (32 captures per CTA) instead of k_acquire (one CTA per capture)
    int         ws_aliased = 0;          // y/z share the rows of sp/ph (L = 1)
    float      *y_sep = nullptr, *z_sep = nullptr;   // separate y/z of a traced call on an aliased context (allocated on demand)
    tiled::TapsPair  taps_pair;          // L = 1: duplicated tap pairs of the packed front kernel (k_front1)
    int         front_packed = 0;        // k_front1 in use (L = 1; PDT_FRONT_SCALAR=1 keeps the scalar k_front<1> for A/B runs)
    size_t      front_smem = 0;
    static constexpr int MAX_MARKS = 512;
    int         mark_group[MAX_MARKS] = {};
    cudaStream_t gstream[MAX_GROUPS] = {}, sstream = nullptr;
    cudaEvent_t  ev_fork = nullptr, ev_join[MAX_GROUPS] = {}, ev_acq[MAX_GROUPS] = {}, ev_sjoin = nullptr;
    int         profiling = 0, n_marks = 0;
    cudaEvent_t marks[MAX_MARKS] = {};
    const char *mark_names[MAX_MARKS] = {};
#endif
};

// how a batch is cut into capture groups (shared by the kernel launches and by the chunked host->device staging)
static int group_plan(const pdt_ctx *c, uint32_t n_captures, uint32_t &per)
{
    static int env_groups = 0;                        // PDT_MAX_GROUPS: experiment knob (default 3)
    if (env_groups == 0) {
        const char *e = getenv("PDT_MAX_GROUPS");
        env_groups = e ? std::max(1, std::min(atoi(e), (int)pdt_ctx::MAX_GROUPS)) : 3;
    }
    const int max_groups = c->max_groups > 0 ? std::min(c->max_groups, (int)pdt_ctx::MAX_GROUPS) : env_groups;
    const int groups = (int)std::min<uint32_t>((uint32_t)max_groups, (n_captures + 63) / 64);
    per = groups > 0 ? (n_captures + groups - 1) / groups : n_captures;
    return groups;
}

#if PDT_USE_FLOATS
// ---- tiled engine: eligibility, workspace, launch sequence -----------------------------------------------
typedef void (*FrontKernel)(const tiled::TiledArgs, const tiled::TapsRev);
static FrontKernel front_kernel(int L)          // one instantiation per interpolation factor the reference can choose (main.c:354)
{
    static const FrontKernel k[tiled::FIR_MAX_L] = {tiled::k_front<1>, tiled::k_front<2>, tiled::k_front<3>, tiled::k_front<4>,
                                                    tiled::k_front<5>, tiled::k_front<6>, tiled::k_front<7>, tiled::k_front<8>};
    return k[L - 1];
}

static bool tiled_applicable(const pdt_params &p, const ChainConst &cc, uint32_t max_captures)
{
    if (cc.argos || cc.L < 1 || cc.L > tiled::FIR_MAX_L || cc.N != tiled::FIR_K * cc.L) return false;
    if (cc.use_mm) return false;                                         // the M&M loop lives in the exact engine
    if (max_captures > 65535u) return false;
    if ((double)cc.chunk * cc.L > 1e9) return false;
    // one 2π wrap per sample must suffice (pdt_tiled.cuh::pll_track_step): |Δphase| <= max_freq + (alpha+beta)·3π < 2π
    PllState ps; pll_reset(ps); pll_begin(ps, cc.pll);
    if (!((double)ps.max_freq + 10.0 * ((double)ps.alpha + (double)ps.beta) < 4.0)) return false;     // 2π + 4 < 10.5: inside the checked range
    if (!(cc.pll.bw_track > 0) || !(cc.agc_decay > 0)) return false;
    const double S = (double)cc.gardner_fs / (double)cc.baud;            // k_gardner: the window must hold several symbols
    if (!(S > 2.0) || S > tiled::GAR_WIN / 8) return false;
    return true;
}

static int tiled_setup(pdt_ctx *c)
{
    using namespace tiled;
    TiledArgs &t = c->ta;
    memset(&t, 0, sizeof t);
    t.cc = c->cc;
    const ChainConst &cc = c->cc;
    const u64 stride = c->max_samples;
    // PLL plan: the track loop is critically damped with time constant 2/alpha ≈ 1/(2·bw) samples; 34 time
    // constants take a (2 Hz, 0.2 rad) carrier guess down to a bit-identical state (measured).
    u64 W = c->params.pll_warm ? c->params.pll_warm : (u64)(17.0 / (double)cc.pll.bw_track);
    W = (W + 3) & ~3ull; if (W < 1024) W = 1024;
    u64 T = c->params.pll_tile ? c->params.pll_tile : W;
    T = (T + 3) & ~3ull; if (T < 1024) T = 1024;
    t.pll.W = W; t.pll.T = T; t.pll.T0 = W + T;
    t.pll.max_tiles = 1 + (unsigned)((stride + T - 1) / T);
    t.agc_min_tile = c->params.agc_min_tile ? ((c->params.agc_min_tile + 3) & ~3u) : 4096u * (unsigned)cc.L;
    { const char *e = getenv("PDT_AGC_TILE_HALVES"); t.agc_tile_halves = e ? (unsigned)std::max(1, std::min(atoi(e), 16)) : 1u; }   // experiment knob
    {
        const u64 t_min = ((t.agc_min_tile / 2) + 3) & ~3ull;       // agc_plan: T = W/2 >= agc_min_tile/2
        t.agc_max_tiles = 2 + (unsigned)((stride * cc.L + t_min - 1) / t_min);
    }
    t.est_fmax = (float)c->params.max_carrier_dev + 600.0f;
    int D = (int)((double)cc.pll.Fs / (2.5 * (double)t.est_fmax));
    t.est_decim = D < 1 ? 1 : D;
    t.ws_stride = (stride + 3) & ~3ull;
    const size_t nin = (size_t)c->max_captures * t.ws_stride, nout = nin * cc.L;
    cudaError_t e;
#define TA(ptr, bytes) if ((e = cudaMalloc((void **)&(ptr), (bytes))) != cudaSuccess) return fail(PDT_ENOMEM, "tiled workspace (%zu bytes): %s", (size_t)(bytes), cudaGetErrorString(e))
    TA(t.sp, nin * sizeof(float)); TA(t.ph, nin * sizeof(float));
    // L = 1: the FIR output y re-uses the rows of sp and the AGC output z those of ph.  Row-wise this is safe by stream
    // order: capture c's y is written by k_front after the last reader of sp[c] (its PLL kernels), and z[c] by the AGC
    // after k_front has consumed ph[c]; kernels of the other pass only touch the rows of their own captures.  It halves
    // the workspaces of a context (1024 x 1 M: 17 -> 9 GB), which is what lets more batches be in flight — the serial
    // acquisition of a never-locking capture is 60 ms of latency that only other batches can hide.
    // A call with trace taps needs ph and y beside z at the end: it gets separate buffers on demand (tiled_run).
    c->ws_aliased = (cc.L == 1);
    if (c->ws_aliased) { t.y = t.sp; t.z = t.ph; }
    else { TA(t.y, nout * sizeof(float)); TA(t.z, nout * sizeof(float)); }
    TA(t.acq, sizeof(AcqResult) * c->max_captures);
    const size_t np = (size_t)c->max_captures * t.pll.max_tiles, na = (size_t)c->max_captures * t.agc_max_tiles;
    TA(t.guess, np * sizeof(LoopState2)); TA(t.pll_start, np * sizeof(LoopState2)); TA(t.pll_end, np * sizeof(LoopState2));
    t.pll_nck = (unsigned)(t.pll.T / PLL_CK) + 2;
    TA(t.pll_ckpt, np * t.pll_nck * sizeof(LoopState2));
    TA(t.agc_start, na * sizeof(LoopState2)); TA(t.agc_end, na * sizeof(LoopState2));
    TA(t.counters, 4 * sizeof(uint32_t));
    TA(t.slow_list, 2 * (size_t)c->max_captures * sizeof(uint32_t)); TA(t.slow_count, 2 * sizeof(uint32_t));
    // work lists of the persistent lane-stream kernels: one region per pass (fast groups / slow captures run concurrently)
    t.pll_tasks_per_cap = (t.pll.max_tiles + 31) / 32;
    t.agc_tasks_per_cap = (t.agc_max_tiles + 31) / 32;
    TA(t.pll_tasks, 2 * (size_t)c->max_captures * t.pll_tasks_per_cap * sizeof(LaneTask));
    TA(t.agc_tasks, 2 * (size_t)c->max_captures * t.agc_tasks_per_cap * sizeof(LaneTask));
    TA(t.task_counts, 2 * (pdt_ctx::MAX_GROUPS + 1) * 2 * sizeof(uint32_t));
    {
        const double S = (double)cc.gardner_fs / (double)cc.baud;        // samples per symbol
        t.sym_cap = (u64)((double)stride * cc.L / (S - 0.2)) + stride / cc.chunk + 64;
        t.sym_cap = (t.sym_cap + 3) & ~3ull;
        TA(t.sym, (size_t)c->max_captures * t.sym_cap * sizeof(float));
        TA(t.gidx, (size_t)c->max_captures * t.sym_cap * sizeof(u64));
        TA(t.gar, sizeof(GarRecord) * c->max_captures);
    }
#undef TA
    for (int u = 0; u < cc.N; u++) c->taps_rev.hr[u] = c->taps_h[cc.N - 1 - u];
    if (cc.L == 1) {
        for (int u = 0; u < FIR_K; u++) c->taps_pair.hh[u] = make_float2(c->taps_rev.hr[u], c->taps_rev.hr[u]);
        c->taps_pair.one = 1.0f; c->taps_pair.pad = 0.0f;
        const char *e = getenv("PDT_FRONT_SCALAR");
        c->front_packed = !(e && atoi(e) != 0);
    }
    c->front_smem = front_smem_bytes(cc.L);
    if ((e = cudaFuncSetAttribute(front_kernel(cc.L), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->front_smem)) != cudaSuccess)
        return fail(PDT_ECUDA, "cudaFuncSetAttribute(k_front): %s", cudaGetErrorString(e));
    if ((e = cudaFuncSetAttribute(k_acquire_packed, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AcqPackSmem))) != cudaSuccess)
        return fail(PDT_ECUDA, "cudaFuncSetAttribute(k_acquire_packed): %s", cudaGetErrorString(e));
    {
        const char *ev = getenv("PDT_ACQ_PACKED");           // 0: the one-CTA-per-capture acquisition (k_acquire), kept for A/B runs
        c->acq_packed = !(ev && atoi(ev) == 0);
    }
    for (void (*kl)(const TiledArgs) : {k_pll_core, k_pll_fix_par, k_agc_core, k_agc_fix_par})
        if ((e = cudaFuncSetAttribute(kl, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LS_SMEM)) != cudaSuccess)
            return fail(PDT_ECUDA, "cudaFuncSetAttribute(lane-stream kernel): %s", cudaGetErrorString(e));
    return PDT_OK;
}

static void tiled_free(pdt_ctx *c)
{
    tiled::TiledArgs &t = c->ta;
    cudaFree(t.sp); cudaFree(t.ph); cudaFree(t.acq); cudaFree(t.guess);
    if (!c->ws_aliased) { cudaFree(t.y); cudaFree(t.z); }
    cudaFree(c->y_sep); cudaFree(c->z_sep);
    cudaFree(t.pll_start); cudaFree(t.pll_end); cudaFree(t.pll_ckpt); cudaFree(t.slow_list); cudaFree(t.slow_count); cudaFree(t.agc_start); cudaFree(t.agc_end); cudaFree(t.counters);
    cudaFree(t.sym); cudaFree(t.gidx); cudaFree(t.gar); cudaFree(t.pll_tasks); cudaFree(t.agc_tasks); cudaFree(t.task_counts);
}

// Kernel sequences for captures [c0, c0+cnt) of the batch.  `head` = StaticGain, sample phases and the first
// acquisition pass; `pipeline` = everything behind the lock latch, for the captures the pass owns (slow_pass 0: those
// that latched within the first acquisition pass, 1: the rest, after their second acquisition pass).
struct GroupLaunch {
    pdt_ctx *c; tiled::TiledArgs t; uint32_t cnt; tiled::u64 n_max; bool marks; int gid = 0;
    bool slow_listed = false;            // acquire_rest has built the list of this launch group's slow captures
    static unsigned blocks(tiled::u64 items, unsigned per) { return (unsigned)((items + per - 1) / per); }
    void mark(cudaStream_t s, const char *name)
    {
        if (!marks || c->n_marks >= pdt_ctx::MAX_MARKS) return;
        if (!c->marks[c->n_marks]) cudaEventCreate(&c->marks[c->n_marks]);
        cudaEventRecord(c->marks[c->n_marks], s);
        c->mark_group[c->n_marks] = gid;
        c->mark_names[c->n_marks++] = name;
    }
    void head(cudaStream_t s)
    {
        using namespace tiled;
        mark(s, "begin");
        k_norm<<<blocks((u64)cnt * 32, 128), 128, 0, s>>>(t);
        mark(s, "k_norm");
        dim3 g(std::min<unsigned>(blocks(n_max, 1024), 4096), cnt);
        k_sp<<<g, 256, 0, s>>>(t);
        mark(s, "k_sp");
        // captures >= prelock_from start in track mode from a carrier estimate (k_prelock); one whose estimate is rejected falls
        // back to the reference's acquisition sweep, so k_acquire covers every capture and skips the pre-locked ones
        const uint32_t serial = std::min(cnt, t.prelock_from);
        if (serial < cnt) { k_prelock<<<blocks(cnt - serial, EST_WARPS), EST_WARPS * 32, 0, s>>>(t); mark(s, "k_prelock"); count_launch(1); }
        static int acq0_threads = 0;                      // PDT_ACQ0_THREADS: experiment knob (CTA width of the first acquisition pass)
        if (acq0_threads == 0) {
            const char *e = getenv("PDT_ACQ0_THREADS");
            acq0_threads = e ? std::max(96, std::min((atoi(e) / 32) * 32, (int)ACQ_THREADS)) : ACQ_THREADS;
        }
        if (dbg_skip("k_acquire0")) {
        } else if (c->acq_packed) k_acquire_packed<<<blocks(cnt, AQ), AP_THREADS, sizeof(AcqPackSmem), s>>>(t, 0, nullptr, nullptr);
        else k_acquire<<<cnt, acq0_threads, 0, s>>>(t, 0);
        mark(s, "k_acquire");
        count_launch(3);
    }
    void acquire_rest(cudaStream_t s)
    {
        using namespace tiled;
        cudaMemsetAsync(t.slow_count, 0, 2 * sizeof(uint32_t), s);
        k_slow_list<<<blocks(cnt, 128), 128, 0, s>>>(t, t.slow_list, t.slow_count);
        slow_listed = true;
        count_launch(1);
        if (c->acq_packed) k_acquire_packed<<<blocks(cnt, AQ), AP_THREADS, sizeof(AcqPackSmem), s>>>(t, 1, t.slow_list, t.slow_count);
        else k_acquire<<<std::max(1u, cnt), ACQ_THREADS_SLOW, 0, s>>>(t, 1);
        mark(s, "k_acquire");
        count_launch(1);
    }
    static bool dbg_skip(const char *name)                // PDT_DEBUG_SKIP=k_gardner,k_bits,…: TIMING EXPERIMENTS ONLY (results are garbage)
    {
        static const char *list = getenv("PDT_DEBUG_SKIP");
        return list && strstr(list, name) != nullptr;
    }
    void pipeline(cudaStream_t s, int slow_pass)
    {
        using namespace tiled;
        TiledArgs q = t;
        q.slow_pass = slow_pass;
        const int L = c->cc.L;
        // work lists: fast pass of group `slot` -> first region, counters 2·slot; slow pass -> second region
        const int slot = (gid >= 0 && gid < pdt_ctx::MAX_GROUPS) ? gid : pdt_ctx::MAX_GROUPS;
        if (slow_pass) {
            q.pll_tasks += (size_t)c->max_captures * q.pll_tasks_per_cap;
            q.agc_tasks += (size_t)c->max_captures * q.agc_tasks_per_cap;
        }
        q.task_counts = c->ta.task_counts + 2 * (size_t)(slow_pass * (pdt_ctx::MAX_GROUPS + 1) + slot);
        static unsigned ls_per_sm = 0;                    // PDT_LS_CAP: experiment knob (CTAs per SM of a persistent lane-stream launch)
        if (ls_per_sm == 0) { const char *e = getenv("PDT_LS_CAP"); ls_per_sm = e ? (unsigned)std::max(1, std::min(atoi(e), 8)) : 2u; }
        const unsigned ls_cap = ls_per_sm * (unsigned)std::max(c->sm_count, 1);    // resident CTAs of a persistent lane-stream launch
        const unsigned pll_grid = std::min(blocks((u64)cnt * q.pll_tasks_per_cap, LS_WARPS), ls_cap);
        const unsigned agc_grid = std::min(blocks((u64)cnt * q.agc_tasks_per_cap, LS_WARPS), ls_cap);
        k_pll_tasks<<<blocks(cnt, 128), 128, 0, s>>>(q);
        if (q.pll.max_tiles > 1)
            k_estimate<<<blocks((u64)cnt * (q.pll.max_tiles - 1), EST_WARPS), EST_WARPS * 32, 0, s>>>(q);
        mark(s, "k_estimate");
        if (!dbg_skip("k_pll_core")) k_pll_core<<<pll_grid, LS_WARPS * 32, LS_SMEM, s>>>(q);
        mark(s, "k_pll_core");
        if (!dbg_skip("k_pll_fix_par")) {
            k_pll_fix_par<<<pll_grid, LS_WARPS * 32, LS_SMEM, s>>>(q);
            k_pll_fix_par<<<pll_grid, LS_WARPS * 32, LS_SMEM, s>>>(q);
        }
        mark(s, "k_pll_fix_par");
        k_pll_fix<<<blocks(cnt, 128), 128, 0, s>>>(q);
        mark(s, "k_pll_fix");
        if (dbg_skip("k_front")) {
        } else if (c->front_packed && slow_pass && slow_listed) {
            q.cap_list = t.slow_list + t.n_captures; q.cap_list_count = t.slow_count + 1;
            dim3 g(blocks(n_max, F1_SPAN), std::max(1u, blocks(cnt, 8)));         // rows stride over the list: any count is covered
            if (q.pcm16) k_front1<true, 1><<<g, F1_THREADS, 0, s>>>(q, c->taps_pair);
            else         k_front1<false, 1><<<g, F1_THREADS, 0, s>>>(q, c->taps_pair);
        } else if (c->front_packed) {
            static int persist = -1;                      // PDT_FRONT_PERSIST=k: experiment — k resident CTAs per SM stride over the items
            if (persist < 0) { const char *e = getenv("PDT_FRONT_PERSIST"); persist = e ? std::max(0, std::min(atoi(e), 7)) : 0; }
            q.front_tiles = blocks(n_max, F1_SPAN);
            if (persist) {
                const unsigned gp = (unsigned)std::min<u64>((u64)cnt * q.front_tiles, (u64)persist * (unsigned)std::max(c->sm_count, 1));
                if (q.pcm16) k_front1<true, 2><<<gp, F1_THREADS, 0, s>>>(q, c->taps_pair);
                else         k_front1<false, 2><<<gp, F1_THREADS, 0, s>>>(q, c->taps_pair);
            } else {
                dim3 g(q.front_tiles, cnt);
                if (q.pcm16) k_front1<true, 0><<<g, F1_THREADS, 0, s>>>(q, c->taps_pair);
                else         k_front1<false, 0><<<g, F1_THREADS, 0, s>>>(q, c->taps_pair);
            }
        } else {
            dim3 g(blocks(n_max, front_span(L)), cnt);
            front_kernel(L)<<<g, front_threads(L), c->front_smem, s>>>(q, c->taps_rev);
        }
        mark(s, "k_front");
        k_agc_plan<<<blocks((u64)cnt * 32, 128), 128, 0, s>>>(q);
        mark(s, "k_agc_plan");
        if (!dbg_skip("k_agc_core")) k_agc_core<<<agc_grid, LS_WARPS * 32, LS_SMEM, s>>>(q);
        mark(s, "k_agc_core");
        k_agc_fix_par<<<agc_grid, LS_WARPS * 32, LS_SMEM, s>>>(q);
        mark(s, "k_agc_fix_par");
        k_agc_fix<<<blocks(cnt, 128), 128, 0, s>>>(q);
        mark(s, "k_agc_fix");
        if (!dbg_skip("k_gardner")) k_gardner<<<blocks(cnt, GAR_WARPS), GAR_CTA_THREADS, 0, s>>>(q);
        mark(s, "k_gardner");
        if (!dbg_skip("k_bits")) k_bits<<<blocks(cnt, BITS_WARPS), BITS_WARPS * 32, 0, s>>>(q);
        mark(s, "k_bits");
        count_launch(q.pll.max_tiles > 1 ? 13 : 12);
    }
};

static GroupLaunch make_group(pdt_ctx *c, const tiled::TiledArgs &base, uint32_t c0, uint32_t cnt, tiled::u64 n_max, bool marks)
{
    using namespace tiled;
    GroupLaunch g{c, base, cnt, n_max, marks};
    TiledArgs &t = g.t;
    const int L = c->cc.L;
    const size_t elem = t.pcm16 ? 2 * sizeof(int16_t) : 2 * sizeof(float);
    t.iq = (const char *)base.iq + (size_t)c0 * t.stride * elem;
    if (t.n_samples) t.n_samples += c0;
    t.n_captures = cnt;
    t.prelock_from = base.prelock_from > c0 ? std::min(base.prelock_from - c0, cnt) : 0;
    t.sp += (size_t)c0 * t.ws_stride; t.ph += (size_t)c0 * t.ws_stride;
    t.y += (size_t)c0 * t.ws_stride * L; t.z += (size_t)c0 * t.ws_stride * L;
    t.acq += c0;
    t.guess += (size_t)c0 * t.pll.max_tiles; t.pll_start += (size_t)c0 * t.pll.max_tiles; t.pll_end += (size_t)c0 * t.pll.max_tiles;
    t.pll_ckpt += (size_t)c0 * t.pll.max_tiles * t.pll_nck;
    t.agc_start += (size_t)c0 * t.agc_max_tiles; t.agc_end += (size_t)c0 * t.agc_max_tiles;
    t.sym += (size_t)c0 * t.sym_cap; t.gidx += (size_t)c0 * t.sym_cap; t.gar += c0;
    t.pll_tasks += (size_t)c0 * t.pll_tasks_per_cap; t.agc_tasks += (size_t)c0 * t.agc_tasks_per_cap;
    t.stats += c0; t.frames += (size_t)c0 * c->max_frames;
    if (t.traces) t.traces += c0;
    return g;
}

// Captures are independent, and the kernels that carry the serial recurrences (k_acquire, k_pll_core, k_agc_core,
// k_back) are latency-bound with few warps, while k_sp / k_front are throughput-bound.  The batch is therefore cut
// into groups that run the same kernel sequence on internal streams, forked from and joined to the caller's
// stream: one group's serial kernels overlap another group's bulk kernels.  Inside a group the captures that have
// not latched after the first acquisition pass (`acq_first` samples) move to a second stream, where their serial
// acquisition continues while the majority is already running the rest of the chain.
// `ready` (optional): one event per capture group, recorded when that group's samples have arrived on the device
static int tiled_run(pdt_ctx *c, const void *d_iq, int pcm16, uint32_t n_captures, uint64_t stride, const uint64_t *n_samples,
                     const pdt_traces *traces, cudaStream_t s, const cudaEvent_t *ready = nullptr)
{
    using namespace tiled;
    TiledArgs t = c->ta;
    if (traces && c->ws_aliased) {       // trace taps read ph / y / z after the run: give this call its own y and z
        const size_t nout = (size_t)c->max_captures * t.ws_stride * c->cc.L;
        if (!c->y_sep) PDT_CUDA(cudaMalloc((void **)&c->y_sep, nout * sizeof(float)));
        if (!c->z_sep) PDT_CUDA(cudaMalloc((void **)&c->z_sep, nout * sizeof(float)));
        t.y = c->y_sep; t.z = c->z_sep;
    }
    t.iq = d_iq; t.pcm16 = pcm16; t.stride = stride; t.n_captures = n_captures;
    t.n_samples = n_samples ? c->d_nsamp : nullptr; t.n_uniform = stride;
    t.stats = c->d_stats; t.frames = c->d_frames; t.traces = traces ? c->d_traces : nullptr;
    u64 n_max = stride;
    if (n_samples) { n_max = 0; for (uint32_t i = 0; i < n_captures; i++) n_max = std::max<u64>(n_max, n_samples[i]); }
    if (n_max == 0) return PDT_OK;
    const int L = c->cc.L;
    static u64 acq_env = 0;                           // PDT_ACQ_FIRST: experiment knob
    if (acq_env == 0) { const char *e = getenv("PDT_ACQ_FIRST"); acq_env = e ? (u64)atoll(e) : 131072; if (acq_env < 1024) acq_env = 131072; }
    const u64 acq_first = c->params.acq_first ? c->params.acq_first : acq_env;
    t.acq_first = (n_max > 2 * acq_first) ? acq_first : 0;
    t.prelock_from = std::min(c->prelock_from, n_captures);
    if (t.prelock_from == 0) t.acq_first = 0;          // nobody runs the acquisition sweep: no slow-capture pass
    const bool two_pass = t.acq_first != 0;
    static int dbg_only_acq1 = -1;       // PDT_DEBUG_ONLY_ACQ1=1: TIMING EXPERIMENTS ONLY — after the first batch of a context, a batch
    if (dbg_only_acq1 < 0) { const char *e = getenv("PDT_DEBUG_ONLY_ACQ1"); dbg_only_acq1 = e ? atoi(e) : 0; }   // is just the second acquisition pass
    if (dbg_only_acq1 && two_pass) {
        if (!c->acq_backup) {
            // first call: run normally below, then keep what the first pass left behind (slow flags + resume states)
            c->prelock_from = c->prelock_from;    // (no-op; the backup is taken at the end of this function)
        } else {
            PDT_CUDA(cudaMemcpyAsync(t.acq, c->acq_backup, sizeof(AcqResult) * n_captures, cudaMemcpyDeviceToDevice, s));
            if (!c->sstream) {
                int lo = 0, hi = 0;
                cudaDeviceGetStreamPriorityRange(&lo, &hi);
                PDT_CUDA(cudaStreamCreateWithPriority(&c->sstream, cudaStreamNonBlocking, hi));
            }
            if (!c->ev_fork) PDT_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
            if (!c->ev_sjoin) PDT_CUDA(cudaEventCreateWithFlags(&c->ev_sjoin, cudaEventDisableTiming));
            PDT_CUDA(cudaEventRecord(c->ev_fork, s));
            PDT_CUDA(cudaStreamWaitEvent(c->sstream, c->ev_fork, 0));
            GroupLaunch whole = make_group(c, t, 0, n_captures, n_max, false);
            whole.acquire_rest(c->sstream);
            PDT_CUDA(cudaEventRecord(c->ev_sjoin, c->sstream));
            PDT_CUDA(cudaStreamWaitEvent(s, c->ev_sjoin, 0));
            return PDT_OK;
        }
    }
    PDT_CUDA(cudaMemsetAsync(t.counters, 0, 4 * sizeof(uint32_t), s));
    PDT_CUDA(cudaMemsetAsync(t.task_counts, 0, 2 * (pdt_ctx::MAX_GROUPS + 1) * 2 * sizeof(uint32_t), s));
    c->n_marks = 0;
    uint32_t per = 0;
    const int groups = group_plan(c, n_captures, per);
    if (c->profiling == 1 || traces || (groups < 2 && !two_pass)) {
        if (ready) for (int gi = 0; gi < groups; gi++) PDT_CUDA(cudaStreamWaitEvent(s, ready[gi], 0));
        GroupLaunch g = make_group(c, t, 0, n_captures, n_max, c->profiling != 0);
        g.head(s);
        g.pipeline(s, 0);
        if (two_pass) { g.acquire_rest(s); g.pipeline(s, 1); }
    } else {
        const bool tl = c->profiling == 2;        // timeline mode: timing events on every group stream
        if (!c->ev_fork) PDT_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
        if (tl) { GroupLaunch g0 = make_group(c, t, 0, n_captures, n_max, true); g0.gid = -1; g0.mark(s, "fork"); }
        PDT_CUDA(cudaEventRecord(c->ev_fork, s));
        int used = 0;
        for (int gi = 0; gi < groups; gi++) {
            const uint32_t c0 = (uint32_t)gi * per;
            if (c0 >= n_captures) break;
            const uint32_t cnt = std::min<uint32_t>(per, n_captures - c0);
            if (!c->gstream[gi]) PDT_CUDA(cudaStreamCreateWithFlags(&c->gstream[gi], cudaStreamNonBlocking));
            if (!c->ev_join[gi]) PDT_CUDA(cudaEventCreateWithFlags(&c->ev_join[gi], cudaEventDisableTiming));
            PDT_CUDA(cudaStreamWaitEvent(c->gstream[gi], c->ev_fork, 0));
            if (ready) PDT_CUDA(cudaStreamWaitEvent(c->gstream[gi], ready[gi], 0));
            GroupLaunch g = make_group(c, t, c0, cnt, n_max, tl);
            g.gid = gi;
            g.head(c->gstream[gi]);
            if (two_pass) {
                if (!c->ev_acq[gi]) PDT_CUDA(cudaEventCreateWithFlags(&c->ev_acq[gi], cudaEventDisableTiming));
                PDT_CUDA(cudaEventRecord(c->ev_acq[gi], c->gstream[gi]));
            }
            g.pipeline(c->gstream[gi], 0);
            PDT_CUDA(cudaEventRecord(c->ev_join[gi], c->gstream[gi]));
            PDT_CUDA(cudaStreamWaitEvent(s, c->ev_join[gi], 0));
            used = gi + 1;
        }
        static int dbg_skip_slow = -1;                    // PDT_DEBUG_SKIP_SLOW=1: TIMING EXPERIMENTS ONLY — the slow captures are left undone
        if (dbg_skip_slow < 0) { const char *e = getenv("PDT_DEBUG_SKIP_SLOW"); dbg_skip_slow = e ? atoi(e) : 0; }
        if (two_pass && dbg_skip_slow != 1) {
            // ONE slow-capture stream for the whole batch (streams beyond the device's 8 hardware queues alias, and a
            // launch waiting behind the long second acquisition pass would block every stream sharing its queue)
            if (!c->sstream) {
                int lo = 0, hi = 0;
                cudaDeviceGetStreamPriorityRange(&lo, &hi);
                PDT_CUDA(cudaStreamCreateWithPriority(&c->sstream, cudaStreamNonBlocking, hi));
            }
            if (!c->ev_sjoin) PDT_CUDA(cudaEventCreateWithFlags(&c->ev_sjoin, cudaEventDisableTiming));
            for (int gi = 0; gi < used; gi++) PDT_CUDA(cudaStreamWaitEvent(c->sstream, c->ev_acq[gi], 0));
            GroupLaunch whole = make_group(c, t, 0, n_captures, n_max, tl);
            whole.gid = 99;
            whole.mark(c->sstream, "begin");
            if (dbg_skip_slow != 3) whole.acquire_rest(c->sstream);       // (2: second acquisition pass only, 3: slow pipeline only)
            if (dbg_skip_slow != 2) whole.pipeline(c->sstream, 1);
            PDT_CUDA(cudaEventRecord(c->ev_sjoin, c->sstream));
            PDT_CUDA(cudaStreamWaitEvent(s, c->ev_sjoin, 0));
        }
    }
    PDT_CUDA(cudaGetLastError());
    if (dbg_only_acq1 && two_pass && !c->acq_backup) {
        // the first batch ran with PDT_DEBUG_SKIP_SLOW=1 semantics expected (set both): acq[] still holds the first pass's results
        PDT_CUDA(cudaMalloc((void **)&c->acq_backup, sizeof(AcqResult) * c->max_captures));
        PDT_CUDA(cudaMemcpyAsync(c->acq_backup, t.acq, sizeof(AcqResult) * n_captures, cudaMemcpyDeviceToDevice, s));
    }
    if (traces) {      // trace taps that are whole workspaces: copy them out (test/debug path)
        for (uint32_t i = 0; i < n_captures; i++) {
            const u64 n = n_samples ? n_samples[i] : stride;
            if (traces[i].pll_phase) PDT_CUDA(cudaMemcpyAsync(traces[i].pll_phase, t.ph + (size_t)i * t.ws_stride, n * sizeof(float), cudaMemcpyDeviceToDevice, s));
            if (traces[i].lpf) PDT_CUDA(cudaMemcpyAsync(traces[i].lpf, t.y + (size_t)i * t.ws_stride * L, n * L * sizeof(float), cudaMemcpyDeviceToDevice, s));
            if (traces[i].agc) PDT_CUDA(cudaMemcpyAsync(traces[i].agc, t.z + (size_t)i * t.ws_stride * L, n * L * sizeof(float), cudaMemcpyDeviceToDevice, s));
        }
    }
    return PDT_OK;
}
#endif

// checkParity.m:20-90 + daytimeDecode.m:4 on the device frame table: one thread per frame slot, then a serial pass per
// capture for the counter continuity (a capture has tens of frames)
__global__ void k_frame_checks(const pdt_frame *frames, const pdt_capture_stats *stats, pdt_frame_quality *q, uint32_t n_captures,
                               uint32_t max_frames)
{
    const uint32_t cap = blockIdx.x;
    if (cap >= n_captures) return;
    const uint32_t nf = stats[cap].n_frames < max_frames ? stats[cap].n_frames : max_frames;
    const pdt_frame *fr = frames + (size_t)cap * max_frames;
    pdt_frame_quality *qq = q + (size_t)cap * max_frames;
    for (uint32_t f = threadIdx.x; f < max_frames; f += blockDim.x) {
        pdt_frame_quality r; r.counter = 0; r.spacecraft = 0; r.parity_ok = 0; r.parity_bits = 0; r.continuous = 0; r.valid = 0; r.pad = 0;
        if (f < nf && fr[f].complete && fr[f].n_bytes == PDT_FRAME_MAX_BYTES) {
            const uint8_t *b = fr[f].bytes;
            r.valid = 1;
            r.counter = (uint16_t)(((b[4] & 1u) << 8) | b[5]);
            r.spacecraft = b[2];
            const int lo[5] = {2, 19, 36, 53, 70}, hi[5] = {18, 35, 52, 69, 86};
            unsigned bad = 0;
            for (int g = 0; g < 5; g++) {
                unsigned ones = 0;
                for (int k = lo[g]; k <= hi[g]; k++) ones += __popc((unsigned)b[k]);
                if ((ones & 1u) != ((b[103] >> (5 - g)) & 1u)) bad |= 1u << (4 - g);
            }
            r.parity_bits = (uint8_t)bad; r.parity_ok = bad == 0;
        }
        qq[f] = r;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int prev = -1;
        for (uint32_t f = 0; f < nf; f++) {
            if (!qq[f].valid) continue;
            qq[f].continuous = (prev < 0) || (qq[f].counter == (unsigned)((prev + 1) % 320));
            prev = qq[f].counter;
        }
    }
}

extern "C" {

const char *pdt_version(void) { return "pdt-b200 0.1 (sm_100a, "
#if PDT_USE_FLOATS
    "f32"
#else
    "f64"
#endif
    ")"; }
const char *pdt_last_error(void) { return g_err; }
int pdt_real_size(void) { return (int)sizeof(real_t); }
uint64_t pdt_launch_count(void) { return g_launches.load(); }

int pdt_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int pdt_set_device(int ordinal)
{
    if (!device_ok()) return PDT_ENODEV;
    PDT_CUDA(cudaSetDevice(ordinal));
    return PDT_OK;
}

int pdt_params_default(pdt_params *p, int mode, double sample_rate)
{
    if (!p || sample_rate < 1) return fail(PDT_EINVAL, "bad arguments");
    memset(p, 0, sizeof *p);
    p->mode = mode; p->sample_rate = sample_rate;
    const real_t Fs = (real_t)(unsigned int)sample_rate;
    if (mode == PDT_MODE_POES) {                      // POESTIPdemod/main.c:30-104
        p->chunk = 10000;
        p->interp = (int)rint(150000.0 / Fs);         // main.c:347
        p->taps = 26 * p->interp;                     // main.c:348
        p->max_carrier_dev = 4500.0;
        p->pll_acq_gain = 127.3240; p->pll_track_gain = 10.3451; p->pll_lock_alpha = 0.3979;
        p->pll_lock_thresh = 0.08;
        p->agc_attack = 79.5775; p->agc_decay = 159.1549;
        p->lpf_fc = 11000.0;
        p->baud = 8320 * 2 + 0.3;
        p->gardner_err_lim = 0.1; p->gardner_gain = 3.0;
        p->manchester_resync = 1.0;                   // main.c:445 passes the literal, not DSP_MCHSTR_RESYNC_LVL
        strcpy(p->sync_word, "1110110111100010000"); p->sync_len = 19;
        p->sync_frame_len = 103; p->sync_start_bit = 3;   // what sync_generic = 1 would need to equal the application copy
        p->mm_step_range = 3; p->mm_gain = 0.15;          // main.c:435 (commented call)
    } else if (mode == PDT_MODE_ARGOS) {              // ARGOSdemod/main.c:27-65
        p->chunk = 2400;
        p->interp = 1; p->taps = 50;
        p->max_carrier_dev = 550.0;
        p->pll_acq_gain = 16; p->pll_track_gain = 16; p->pll_lock_alpha = 3.1831;
        p->pll_lock_thresh = 0.1;
        p->agc_attack = 79.5775; p->agc_decay = 159.1549;
        p->lpf_fc = 700;
        p->baud = 400 * 2.0;
        p->gardner_err_lim = 0.1; p->gardner_gain = 3.0;
        p->manchester_resync = 0.5;
        p->squelch_thresh = 0.15;
        strcpy(p->sync_word, "0001011110000"); p->sync_len = 13;
        p->sync_frame_len = 8; p->sync_start_bit = 0;
        p->mm_step_range = 3; p->mm_gain = 0.15;          // ARGOSdemod/main.c:277 (commented call)
    } else return fail(PDT_EINVAL, "unknown mode %d", mode);
    return PDT_OK;
}

pdt_ctx *pdt_create(const pdt_params *p, uint32_t max_captures, uint64_t max_samples, uint32_t max_frames)
{
    if (!p || !max_captures || !max_samples || !max_frames) { fail(PDT_EINVAL, "bad arguments"); return nullptr; }
    if (!device_ok()) return nullptr;
    pdt_ctx *c = new (std::nothrow) pdt_ctx();
    if (!c) { fail(PDT_ENOMEM, "out of host memory"); return nullptr; }
    c->params = *p;
    if (c->params.force_min_interp1 && c->params.mode == PDT_MODE_POES && c->params.interp < 1) {
        c->params.interp = 1; c->params.taps = 26;
    }
    c->max_captures = max_captures; c->max_samples = max_samples; c->max_frames = max_frames;
    if (build_chain_const(c->params, c->cc) != PDT_OK) { delete c; return nullptr; }
    c->cc.max_frames = max_frames;
    auto bail = [&](cudaError_t e, const char *what) -> pdt_ctx * {
        fail(PDT_ECUDA, "%s: %s", what, cudaGetErrorString(e));
        pdt_destroy(c);
        return nullptr;
    };
    cudaError_t e;
    if ((e = cudaGetDevice(&c->device)) != cudaSuccess) return bail(e, "cudaGetDevice");
    cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, c->device);
    c->engine = PDT_ENGINE_EXACT;
#if PDT_USE_FLOATS
    if (c->params.engine != PDT_ENGINE_EXACT && tiled_applicable(c->params, c->cc, max_captures)) c->engine = PDT_ENGINE_TILED;
#endif
    if (c->params.engine == PDT_ENGINE_TILED && c->engine != PDT_ENGINE_TILED) {
        fail(PDT_EINVAL, "the tiled engine needs the float POES chain with 1 <= interp <= 8 and taps = 26*interp");
        delete c;
        return nullptr;
    }
    if (c->cc.L > 0) {
        const real_t Fs = (real_t)(unsigned int)c->params.sample_rate;
        if (c->cc.argos) make_lpfir_host(c->taps_h, c->cc.N, (real_t)c->params.lpf_fc, Fs, 1);            // ARGOS main.c:248
        else             make_lpfir_host(c->taps_h, c->cc.N, (real_t)c->params.lpf_fc, Fs * c->cc.L, c->cc.L);   // main.c:369
        if ((e = cudaMalloc(&c->d_taps, sizeof(real_t) * c->cc.N)) != cudaSuccess) return bail(e, "cudaMalloc taps");
        if ((e = cudaMemcpy(c->d_taps, c->taps_h, sizeof(real_t) * c->cc.N, cudaMemcpyHostToDevice)) != cudaSuccess) return bail(e, "taps H2D");
        // shared-memory plan
        int max_optin = 0;
        cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device);
        c->smem_bytes = chain_ws_reals(c->cc) * sizeof(real_t);
        cudaFuncAttributes fa{};
        if ((e = cudaFuncGetAttributes(&fa, k_chain_exact)) != cudaSuccess) return bail(e, "cudaFuncGetAttributes");
        const size_t static_smem = fa.sharedSizeBytes + 1024;
        // the chain is latency-bound on a few serial threads per CTA: what buys throughput is the number of resident CTAs.  The
        // chunk buffers go to shared memory only when that does not cost residency (ARGOS: 58 KB of chunk per CTA would leave
        // 2 CTAs per SM where the per-CTA global workspace, L1/L2-resident and touched by the parallel phases only, allows 6).
        const bool fits = (c->smem_bytes + static_smem <= (size_t)max_optin);
        int per_smem = 0, per_glob = 0;
        if (fits) {
            if ((e = cudaFuncSetAttribute(k_chain_exact, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_bytes)) != cudaSuccess)
                return bail(e, "cudaFuncSetAttribute");
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_smem, k_chain_exact, CHAIN_THREADS, c->smem_bytes);
        }
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_glob, k_chain_exact, CHAIN_THREADS, 0);
        per_glob = std::min(per_glob, 8);
        c->use_smem = fits && per_smem >= per_glob;
        int per_sm = c->use_smem ? per_smem : per_glob;
        if (!c->use_smem) c->smem_bytes = 0;
        if (per_sm < 1) per_sm = 1;
        c->grid = (int)std::min<uint64_t>(max_captures, (uint64_t)c->sm_count * per_sm);
        if (!c->use_smem) {
            c->ws_stride = (chain_ws_reals(c->cc) + 31) & ~(size_t)31;
            if ((e = cudaMalloc(&c->d_ws, c->ws_stride * sizeof(real_t) * c->grid)) != cudaSuccess) return bail(e, "cudaMalloc workspace");
        }
    }
#if PDT_USE_FLOATS
    if (c->engine == PDT_ENGINE_TILED && tiled_setup(c) != PDT_OK) { pdt_destroy(c); return nullptr; }
#endif
    if ((e = cudaMalloc(&c->d_stats, sizeof(pdt_capture_stats) * max_captures)) != cudaSuccess) return bail(e, "cudaMalloc stats");
    if ((e = cudaMalloc(&c->d_frames, sizeof(pdt_frame) * (size_t)max_captures * max_frames)) != cudaSuccess) return bail(e, "cudaMalloc frames");
    if ((e = cudaMalloc(&c->d_nsamp, sizeof(unsigned long long) * max_captures)) != cudaSuccess) return bail(e, "cudaMalloc nsamp");
    if ((e = cudaMalloc(&c->d_traces, sizeof(pdt_traces) * max_captures)) != cudaSuccess) return bail(e, "cudaMalloc traces");
    return c;
}

void pdt_destroy(pdt_ctx *c)
{
    if (!c) return;
    cudaFree(c->d_taps); cudaFree(c->d_ws); cudaFree(c->d_stats); cudaFree(c->d_frames);
    cudaFree(c->d_nsamp); cudaFree(c->d_traces); cudaFree(c->d_stage); cudaFree(c->d_quality);
    cudaFree(c->d_live); cudaFree(c->d_live_ws);
#if PDT_USE_FLOATS
    if (c->engine == PDT_ENGINE_TILED) tiled_free(c);
    for (cudaEvent_t e : c->marks) if (e) cudaEventDestroy(e);
    for (cudaStream_t gs : c->gstream) if (gs) cudaStreamDestroy(gs);
    if (c->sstream) cudaStreamDestroy(c->sstream);
    for (cudaEvent_t e : c->ev_join) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : c->ev_acq) if (e) cudaEventDestroy(e);
    if (c->ev_sjoin) cudaEventDestroy(c->ev_sjoin);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
#endif
    if (c->h2d_stream) cudaStreamDestroy(c->h2d_stream);
    for (cudaEvent_t e : c->ev_h2d) if (e) cudaEventDestroy(e);
    if (c->ev_done) cudaEventDestroy(c->ev_done);
    delete c;
}

int pdt_get_params(const pdt_ctx *c, pdt_params *out)
{
    if (!c || !out) return fail(PDT_EINVAL, "bad arguments");
    *out = c->params;
    return PDT_OK;
}

int pdt_get_taps(const pdt_ctx *c, void *h_out)
{
    if (!c || !h_out) return fail(PDT_EINVAL, "bad arguments");
    memcpy(h_out, c->taps_h, sizeof(real_t) * (size_t)std::max(c->cc.N, 0));
    return PDT_OK;
}

static int demod_device_impl(pdt_ctx *c, const void *d_iq, int pcm16, uint32_t n_captures, uint64_t stride_samples,
                             const uint64_t *n_samples, const pdt_traces *traces, void *stream, const cudaEvent_t *ready);

int pdt_demod_device(pdt_ctx *c, const void *d_iq, int pcm16, uint32_t n_captures, uint64_t stride_samples,
                     const uint64_t *n_samples, const pdt_traces *traces, void *stream)
{
    return demod_device_impl(c, d_iq, pcm16, n_captures, stride_samples, n_samples, traces, stream, nullptr);
}

static int demod_device_impl(pdt_ctx *c, const void *d_iq, int pcm16, uint32_t n_captures, uint64_t stride_samples,
                             const uint64_t *n_samples, const pdt_traces *traces, void *stream, const cudaEvent_t *ready)
{
    if (!c || !d_iq || !n_captures || n_captures > c->max_captures) return fail(PDT_EINVAL, "bad arguments");
    if (!device_ok()) return PDT_ENODEV;
    cudaStream_t s = (cudaStream_t)stream;
    PDT_CUDA(cudaMemsetAsync(c->d_stats, 0, sizeof(pdt_capture_stats) * n_captures, s));
    PDT_CUDA(cudaMemsetAsync(c->d_frames, 0, sizeof(pdt_frame) * (size_t)n_captures * c->max_frames, s));
    if (n_samples) {
        for (uint32_t i = 0; i < n_captures; i++)
            if (n_samples[i] > c->max_samples) return fail(PDT_EINVAL, "capture %u too long", i);      // > stride: overlapping segments
        PDT_CUDA(cudaMemcpyAsync(c->d_nsamp, n_samples, sizeof(uint64_t) * n_captures, cudaMemcpyHostToDevice, s));
    } else if (stride_samples > c->max_samples) return fail(PDT_EINVAL, "captures longer than the context allows");
    if (traces) PDT_CUDA(cudaMemcpyAsync(c->d_traces, traces, sizeof(pdt_traces) * n_captures, cudaMemcpyHostToDevice, s));
    if (c->cc.L <= 0) {
        // Fs >= 300 ksps: L = rint(150000/Fs) = 0 and the reference silently emits nothing. Same here.
        return PDT_OK;
    }
#if PDT_USE_FLOATS
    if (c->engine == PDT_ENGINE_TILED) {
        bool needs_exact = false;      // per-sample frequency / lock-detector taps only exist in the exact engine
        if (traces) for (uint32_t i = 0; i < n_captures; i++) needs_exact |= (traces[i].pll_freq || traces[i].lock);
        if (!needs_exact) return tiled_run(c, d_iq, pcm16, n_captures, stride_samples, n_samples, traces, s, ready);
    }
#endif
    if (ready) { uint32_t per = 0; const int groups = group_plan(c, n_captures, per); for (int gi = 0; gi < groups; gi++) PDT_CUDA(cudaStreamWaitEvent(s, ready[gi], 0)); }
    ChainArgs a;
    a.cc = c->cc; a.taps = c->d_taps; a.iq = d_iq; a.pcm16 = pcm16; a.stride = stride_samples;
    a.n_samples = n_samples ? c->d_nsamp : nullptr; a.n_uniform = stride_samples; a.n_captures = n_captures;
    a.workspace = c->d_ws; a.ws_stride = c->ws_stride; a.use_smem = c->use_smem;
    a.stats = c->d_stats; a.frames = c->d_frames; a.traces = traces ? c->d_traces : nullptr; a.persist = nullptr;
    const int grid = (int)std::min<uint32_t>(n_captures, (uint32_t)c->grid);
    k_chain_exact<<<grid, CHAIN_THREADS, c->smem_bytes, s>>>(a);
    count_launch();
    PDT_CUDA(cudaGetLastError());
    return PDT_OK;
}

// ---- live mode: bounded-latency streaming with carried state (POESTIPdemodPortAudio/main.c:324-401) ----------------
int pdt_live_begin(pdt_ctx *c)
{
    if (!c) return fail(PDT_EINVAL, "bad arguments");
    if (!device_ok()) return PDT_ENODEV;
    if (c->cc.L <= 0) return fail(PDT_EINVAL, "live mode needs interp >= 1 (set force_min_interp1 above 300 ksps)");
    if (!c->d_live) {
        c->live_ws_stride = (chain_ws_reals(c->cc) + 31) & ~(size_t)31;
        PDT_CUDA(cudaMalloc((void **)&c->d_live, sizeof(ChainState) * c->max_captures));
        PDT_CUDA(cudaMalloc((void **)&c->d_live_ws, sizeof(real_t) * c->live_ws_stride * c->max_captures));
    }
    PDT_CUDA(cudaMemset(c->d_live, 0, sizeof(ChainState) * c->max_captures));
    PDT_CUDA(cudaMemset(c->d_live_ws, 0, sizeof(real_t) * c->live_ws_stride * c->max_captures));
    PDT_CUDA(cudaMemset(c->d_stats, 0, sizeof(pdt_capture_stats) * c->max_captures));
    PDT_CUDA(cudaMemset(c->d_frames, 0, sizeof(pdt_frame) * (size_t)c->max_captures * c->max_frames));
    return PDT_OK;
}

int pdt_live_push_device(pdt_ctx *c, const void *d_iq, int pcm16, uint32_t n_streams, uint64_t stride_samples, uint64_t n, void *stream)
{
    if (!c || !d_iq || !n_streams || n_streams > c->max_captures || n > stride_samples) return fail(PDT_EINVAL, "bad arguments");
    if (!c->d_live) return fail(PDT_EINVAL, "pdt_live_begin first");
    if (n == 0) return PDT_OK;
    ChainArgs a;
    a.cc = c->cc; a.cc.ring_frames = 1;
    a.taps = c->d_taps; a.iq = d_iq; a.pcm16 = pcm16; a.stride = stride_samples;
    a.n_samples = nullptr; a.n_uniform = n; a.n_captures = n_streams;
    a.workspace = c->d_live_ws; a.ws_stride = c->live_ws_stride; a.use_smem = 0;
    a.stats = c->d_stats; a.frames = c->d_frames; a.traces = nullptr; a.persist = c->d_live;
    const int grid = (int)std::min<uint32_t>(n_streams, (uint32_t)std::max(c->sm_count, 1) * 8u);
    k_chain_exact<<<grid, CHAIN_THREADS, 0, (cudaStream_t)stream>>>(a);
    count_launch();
    PDT_CUDA(cudaGetLastError());
    return PDT_OK;
}

int pdt_live_push_host(pdt_ctx *c, const void *h_iq, int pcm16, uint32_t n_streams, uint64_t stride_samples, uint64_t n,
                       pdt_capture_stats *stats_out, pdt_frame *frames_out)
{
    if (!c || !h_iq) return fail(PDT_EINVAL, "bad arguments");
    if (!device_ok()) return PDT_ENODEV;
    const size_t elem = pcm16 ? 2 * sizeof(int16_t) : 2 * sizeof(real_t);
    const size_t bytes = elem * stride_samples * n_streams;
    if (bytes > c->stage_bytes) {
        PDT_CUDA(cudaDeviceSynchronize());
        cudaFree(c->d_stage); c->d_stage = nullptr; c->stage_bytes = 0;
        PDT_CUDA(cudaMalloc(&c->d_stage, bytes));
        c->stage_bytes = bytes;
    }
    PDT_CUDA(cudaMemcpyAsync(c->d_stage, h_iq, bytes, cudaMemcpyHostToDevice, nullptr));
    const int rc = pdt_live_push_device(c, c->d_stage, pcm16, n_streams, stride_samples, n, nullptr);
    if (rc != PDT_OK) return rc;
    return pdt_fetch(c, n_streams, stats_out, frames_out, nullptr);
}

int pdt_fetch(pdt_ctx *c, uint32_t n_captures, pdt_capture_stats *stats_out, pdt_frame *frames_out, void *stream)
{
    if (!c || n_captures > c->max_captures) return fail(PDT_EINVAL, "bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    if (stats_out) PDT_CUDA(cudaMemcpyAsync(stats_out, c->d_stats, sizeof(pdt_capture_stats) * n_captures, cudaMemcpyDeviceToHost, s));
    if (frames_out) PDT_CUDA(cudaMemcpyAsync(frames_out, c->d_frames, sizeof(pdt_frame) * (size_t)n_captures * c->max_frames, cudaMemcpyDeviceToHost, s));
    PDT_CUDA(cudaStreamSynchronize(s));
    return PDT_OK;
}

int pdt_engine(const pdt_ctx *c) { return c ? c->engine : PDT_EINVAL; }

int pdt_tiled_counters(pdt_ctx *c, uint32_t out[4], void *stream)
{
    if (!c || !out) return fail(PDT_EINVAL, "bad arguments");
    out[0] = out[1] = out[2] = out[3] = 0;
#if PDT_USE_FLOATS
    if (c->engine == PDT_ENGINE_TILED) {
        PDT_CUDA(cudaMemcpyAsync(out, c->ta.counters, 3 * sizeof(uint32_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
        PDT_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
        out[3] = c->ta.pll.max_tiles;
    }
#endif
    return PDT_OK;
}

int pdt_set_groups(pdt_ctx *c, int max_groups)
{
    if (!c || max_groups < 0) return fail(PDT_EINVAL, "bad arguments");
    c->max_groups = max_groups;
    return PDT_OK;
}

int pdt_set_profiling(pdt_ctx *c, int enable)
{
    if (!c) return fail(PDT_EINVAL, "bad arguments");
#if PDT_USE_FLOATS
    c->profiling = enable;
#endif
    return PDT_OK;
}

int pdt_kernel_times(pdt_ctx *c, const char **names, float *ms, int cap)
{
    if (!c || !names || !ms) return fail(PDT_EINVAL, "bad arguments");
#if PDT_USE_FLOATS
    if (c->engine != PDT_ENGINE_TILED || c->n_marks < 2) return 0;
    if (c->profiling == 2) return 0;
    PDT_CUDA(cudaEventSynchronize(c->marks[c->n_marks - 1]));
    int k = 0;
    for (int i = 1; i < c->n_marks && k < cap; i++, k++) {
        names[k] = c->mark_names[i];
        PDT_CUDA(cudaEventElapsedTime(&ms[k], c->marks[i - 1], c->marks[i]));
    }
    return k;
#else
    return 0;
#endif
}

int pdt_timeline(pdt_ctx *c, const char **names, int *groups, float *end_ms, int cap)
{
    if (!c || !names || !groups || !end_ms) return fail(PDT_EINVAL, "bad arguments");
#if PDT_USE_FLOATS
    if (c->engine != PDT_ENGINE_TILED || c->profiling != 2 || c->n_marks < 2) return 0;
    PDT_CUDA(cudaDeviceSynchronize());
    int k = 0;
    for (int i = 1; i < c->n_marks && k < cap; i++, k++) {
        names[k] = c->mark_names[i]; groups[k] = c->mark_group[i];
        PDT_CUDA(cudaEventElapsedTime(&end_ms[k], c->marks[0], c->marks[i]));
    }
    return k;
#else
    return 0;
#endif
}

int pdt_debug_chain_prof(uint64_t out[20], int reset)
{
    if (!out) return fail(PDT_EINVAL, "bad arguments");
    unsigned long long h[20];
    PDT_CUDA(cudaDeviceSynchronize());
    PDT_CUDA(cudaMemcpyFromSymbol(h, g_chain_prof, sizeof h));
    for (int i = 0; i < 20; i++) out[i] = h[i];
    if (reset) { unsigned long long z[20] = {}; PDT_CUDA(cudaMemcpyToSymbol(g_chain_prof, z, sizeof z)); }
    return PDT_OK;
}

int pdt_debug_acq_prof(uint64_t out[8], int reset)
{
#if PDT_USE_FLOATS
    unsigned long long h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    PDT_CUDA(cudaDeviceSynchronize());
    PDT_CUDA(cudaMemcpyFromSymbol(h, tiled::g_acq_prof, sizeof h));
    for (int i = 0; i < 8; i++) out[i] = h[i];
    if (reset) { unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0}; PDT_CUDA(cudaMemcpyToSymbol(tiled::g_acq_prof, z, sizeof z)); }
#else
    for (int i = 0; i < 8; i++) out[i] = 0;
#endif
    return PDT_OK;
}

int pdt_frame_checks(pdt_ctx *c, uint32_t n_captures, pdt_frame_quality *quality_out, void *stream)
{
    if (!c || !quality_out || n_captures > c->max_captures) return fail(PDT_EINVAL, "bad arguments");
    if (!device_ok()) return PDT_ENODEV;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t bytes = sizeof(pdt_frame_quality) * (size_t)c->max_captures * c->max_frames;
    if (!c->d_quality) PDT_CUDA(cudaMalloc(&c->d_quality, bytes));
    k_frame_checks<<<n_captures, 64, 0, s>>>(c->d_frames, c->d_stats, c->d_quality, n_captures, c->max_frames);
    count_launch();
    PDT_CUDA(cudaGetLastError());
    PDT_CUDA(cudaMemcpyAsync(quality_out, c->d_quality, sizeof(pdt_frame_quality) * (size_t)n_captures * c->max_frames, cudaMemcpyDeviceToHost, s));
    PDT_CUDA(cudaStreamSynchronize(s));
    return PDT_OK;
}

int pdt_result_tables(pdt_ctx *c, void **d_stats, void **d_frames, uint32_t *max_frames)
{
    if (!c) return fail(PDT_EINVAL, "bad arguments");
    if (d_stats) *d_stats = c->d_stats;
    if (d_frames) *d_frames = c->d_frames;
    if (max_frames) *max_frames = c->max_frames;
    return PDT_OK;
}

int pdt_demod_host_async(pdt_ctx *c, const void *h_iq, int pcm16, uint32_t n_captures, uint64_t stride_samples,
                         const uint64_t *n_samples, void *stream)
{
    if (!c || !h_iq) return fail(PDT_EINVAL, "bad arguments");
    if (!device_ok()) return PDT_ENODEV;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t elem = pcm16 ? 2 * sizeof(int16_t) : 2 * sizeof(real_t);
    const size_t bytes = elem * stride_samples * n_captures;
    if (bytes > c->stage_bytes) {
        PDT_CUDA(cudaDeviceSynchronize());
        cudaFree(c->d_stage); c->d_stage = nullptr; c->stage_bytes = 0;
        PDT_CUDA(cudaMalloc(&c->d_stage, bytes));
        c->stage_bytes = bytes;
    }
    if (n_samples)
        for (uint32_t i = 0; i < n_captures; i++)
            if (n_samples[i] > stride_samples) return fail(PDT_EINVAL, "overlapping captures (segments of a stream) need device-resident input");
    // The samples go up one capture group at a time on a copy stream; every group starts its kernels as soon as its own
    // samples have landed, so all but the first group's transfer is hidden behind the kernels of the groups before it.
    uint32_t per = 0;
    const int groups = group_plan(c, n_captures, per);
    if (!c->h2d_stream) PDT_CUDA(cudaStreamCreateWithFlags(&c->h2d_stream, cudaStreamNonBlocking));
    if (!c->ev_done) PDT_CUDA(cudaEventCreateWithFlags(&c->ev_done, cudaEventDisableTiming));
    else PDT_CUDA(cudaStreamWaitEvent(c->h2d_stream, c->ev_done, 0));      // the previous batch has finished reading the staging buffer
    for (int gi = 0; gi < groups; gi++) {
        const uint32_t c0 = (uint32_t)gi * per;
        const uint32_t cnt = c0 < n_captures ? std::min<uint32_t>(per, n_captures - c0) : 0;
        if (!c->ev_h2d[gi]) PDT_CUDA(cudaEventCreateWithFlags(&c->ev_h2d[gi], cudaEventDisableTiming));
        if (cnt) {
            const size_t off = elem * stride_samples * c0, len = elem * stride_samples * cnt;
            PDT_CUDA(cudaMemcpyAsync((char *)c->d_stage + off, (const char *)h_iq + off, len, cudaMemcpyHostToDevice, c->h2d_stream));
        }
        PDT_CUDA(cudaEventRecord(c->ev_h2d[gi], c->h2d_stream));
    }
    const int rc = demod_device_impl(c, c->d_stage, pcm16, n_captures, stride_samples, n_samples, nullptr, s, c->ev_h2d);
    if (rc != PDT_OK) return rc;
    PDT_CUDA(cudaEventRecord(c->ev_done, s));
    return PDT_OK;
}

int pdt_demod_host(pdt_ctx *c, const void *h_iq, int pcm16, uint32_t n_captures, uint64_t stride_samples,
                   const uint64_t *n_samples, pdt_capture_stats *stats_out, pdt_frame *frames_out)
{
    const int rc = pdt_demod_host_async(c, h_iq, pcm16, n_captures, stride_samples, n_samples, nullptr);
    if (rc != PDT_OK) return rc;
    return pdt_fetch(c, n_captures, stats_out, frames_out, nullptr);
}

long pdt_format_frames(const pdt_ctx *c, const pdt_frame *frames, uint32_t n_frames, char *buf, size_t cap)
{
    if (!c || !frames || !buf) return fail(PDT_EINVAL, "bad arguments");
    // time column: wave.c:91-167 accumulates `time += Ts` in DECIMAL_TYPE per sample; emulate up to the last frame.
    const ChainConst &cc = c->cc;
    const int L = cc.L > 0 ? cc.L : 1;
    std::vector<uint64_t> want(n_frames);
    uint64_t last = 0;
    for (uint32_t i = 0; i < n_frames; i++) {
        uint64_t in_idx = frames[i].sample_index / L;
        // POES: LowPassFilterInterp hands out the NEXT input's time (LowPassFilter.c:68); ARGOS passes waveDataTime through.
        want[i] = cc.argos ? in_idx : in_idx + 1;
        last = std::max(last, want[i]);
    }
    std::vector<real_t> tval(n_frames, 0);
    {
        std::vector<uint32_t> order(n_frames);
        for (uint32_t i = 0; i < n_frames; i++) order[i] = i;
        std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return want[a] < want[b]; });
        real_t t = 0; const real_t Ts = 1.0 / (real_t)(unsigned int)c->params.sample_rate;
        uint64_t g = 0; size_t oi = 0;
        while (oi < order.size()) {
            const uint64_t target = want[order[oi]];
            while (g <= target) { t += Ts; g++; }          // sample g (0-based) carries the time after g+1 additions
            tval[order[oi]] = t; oi++;
        }
    }
    size_t pos = 0;
    auto put = [&](const char *fmt, auto v) {
        if (pos < cap) { int k = snprintf(buf + pos, cap - pos, fmt, v); if (k > 0) pos += (size_t)k; }
    };
    for (uint32_t i = 0; i < n_frames; i++) {
        const pdt_frame &f = frames[i];
        put(f.inverse ? "%.5fi " : "%.5f ", (double)tval[i]);
        for (int b = 0; b < f.n_bytes; b++) put("%.2X ", (unsigned)f.bytes[b]);
        if (f.complete) put("%s", "\n");
    }
    if (pos >= cap) return fail(PDT_EINVAL, "buffer too small");
    buf[pos] = 0;
    return (long)pos;
}

int pdt_demod_segments_device(pdt_ctx *c, const void *d_iq, int pcm16, uint32_t n_segments, uint64_t stride_samples,
                              const uint64_t *n_samples, uint32_t n_serial, void *stream)
{
#if PDT_USE_FLOATS
    if (!c || !n_samples) return fail(PDT_EINVAL, "bad arguments");
    if (c->engine != PDT_ENGINE_TILED) return fail(PDT_EINVAL, "stream segments need the tiled engine");
    const uint64_t pre = tiled::prelock_span(c->ta.est_decim);
    for (uint32_t i = n_serial; i < n_segments; i++)
        if (n_samples[i] < 2 * pre) return fail(PDT_EINVAL, "segment %u is shorter than twice the carrier-estimate span (%llu samples)", i, (unsigned long long)pre);
    c->prelock_from = n_serial;
    const int rc = demod_device_impl(c, d_iq, pcm16, n_segments, stride_samples, n_samples, nullptr, stream, nullptr);
    c->prelock_from = 0xFFFFFFFFu;
    return rc;
#else
    (void)c; (void)d_iq; (void)pcm16; (void)n_segments; (void)stride_samples; (void)n_samples; (void)n_serial; (void)stream;
    return fail(PDT_EINVAL, "stream segments exist for the float POES chain only");
#endif
}

// ---- one long stream as overlapping segments (pdt.h) ---------------------------------------------------------
int pdt_stream_plan_make(pdt_stream_plan *plan, const pdt_params *p, uint64_t total_samples, uint64_t segment, uint64_t lead,
                         uint64_t tail)
{
    if (!plan || !p || !total_samples || !segment || !(p->sample_rate > 0)) return fail(PDT_EINVAL, "bad arguments");
    plan->total_samples = total_samples;
    plan->segment = segment;
    plan->lead = lead ? lead : (uint64_t)std::ceil(0.3 * p->sample_rate);
    plan->tail = tail ? tail : (uint64_t)std::ceil(0.13 * p->sample_rate) + 4096;
    plan->interp = (uint32_t)std::max(p->interp, 1);
    {   // two symbols, in interpolated samples (GardenerClockRecovery.c:20: step = Fs·L / baud)
        const double sym = (double)plan->interp * p->sample_rate / (p->baud > 0 ? p->baud : 16640.3);
        plan->seam_tol = (uint32_t)std::ceil(2.0 * sym);
        plan->pad = 0;
    }
    // the last segment owns everything from its window start to the end of the stream
    const uint64_t k = total_samples > plan->lead ? (total_samples - plan->lead + segment - 1) / segment : 1;
    if (k > 0xFFFFFFFFull) return fail(PDT_EINVAL, "too many segments");
    plan->n_segments = (uint32_t)std::max<uint64_t>(k, 1);
    return PDT_OK;
}

uint64_t pdt_stream_segment_length(const pdt_stream_plan *plan, uint32_t s)
{
    if (!plan || s >= plan->n_segments) return 0;
    const uint64_t start = (uint64_t)s * plan->segment, full = plan->lead + plan->segment + plan->tail;
    return std::min(full, plan->total_samples - start);
}

long pdt_stream_stitch(const pdt_stream_plan *plan, uint32_t first, uint32_t n, const pdt_capture_stats *stats,
                       const pdt_frame *frames, uint32_t max_frames, pdt_frame *out, uint32_t out_cap)
{
    if (!plan || !stats || !frames || !out || (uint64_t)first + n > plan->n_segments) return fail(PDT_EINVAL, "bad arguments");
    const uint64_t L = plan->interp, tol = plan->seam_tol;
    uint32_t k = 0;
    bool have_prev = false; uint64_t prev_g = 0;
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t s = first + i;
        const uint64_t start = (uint64_t)s * plan->segment;
        const bool last = s + 1 == plan->n_segments;
        // ownership window, widened by the seam tolerance on both sides: neighbouring segments see the same sync word up to
        // a symbol apart, so a frame near an edge is claimed by both and de-duplicated by position below (never by neither)
        // (The outer edges of the range being stitched stay exact, so that ranges stitched separately — one per rank — still
        // concatenate to a partition of the stream; only a stitch over ALL segments is jitter-proof at every seam.)
        const uint64_t lo_exact = s == 0 ? 0 : (start + plan->lead) * L;
        const uint64_t lo = (i == 0) ? lo_exact : (lo_exact > tol ? lo_exact - tol : 0);
        const uint64_t hi = last ? ~0ull : (start + plan->segment + plan->lead) * L + (i + 1 == n ? 0 : tol);
        const uint32_t nf = std::min(stats[i].n_frames, max_frames);
        for (uint32_t f = 0; f < nf; f++) {
            const pdt_frame &fr = frames[(size_t)i * max_frames + f];
            const uint64_t g = start * L + fr.sample_index;
            if (g < lo || g >= hi) continue;
            if (!fr.complete && !last) continue;            // cannot happen with tail > one frame; never emit a seam fragment
            if (have_prev && g <= prev_g + 2 * tol) continue;               // the frame just emitted, seen again by this segment
            if (k >= out_cap) return fail(PDT_EINVAL, "stitched frame table too small");
            out[k] = fr;
            out[k].sample_index = g;
            prev_g = g; have_prev = true;
            k++;
        }
    }
    return (long)k;
}

// Host twin of k_frame_checks for a STITCHED frame list (the device table is per segment; continuity across the seams only
// exists after the stitch).  A few hundred bytes per frame, ten frames per second of signal: host work by nature.
int pdt_stream_frame_checks(const pdt_frame *frames, uint32_t n_frames, pdt_frame_quality *quality_out)
{
    if ((!frames || !quality_out) && n_frames) return fail(PDT_EINVAL, "bad arguments");
    int prev = -1;
    for (uint32_t f = 0; f < n_frames; f++) {
        pdt_frame_quality r; memset(&r, 0, sizeof r);
        if (frames[f].complete && frames[f].n_bytes == PDT_FRAME_MAX_BYTES) {
            const uint8_t *b = frames[f].bytes;
            r.valid = 1;
            r.counter = (uint16_t)(((b[4] & 1u) << 8) | b[5]);                          // daytimeDecode.m:4
            r.spacecraft = b[2];
            const int lo[5] = {2, 19, 36, 53, 70}, hi[5] = {18, 35, 52, 69, 86};        // checkParity.m:20-90
            unsigned bad = 0;
            for (int g = 0; g < 5; g++) {
                unsigned ones = 0;
                for (int k = lo[g]; k <= hi[g]; k++) ones += (unsigned)__builtin_popcount((unsigned)b[k]);
                if ((ones & 1u) != ((b[103] >> (5 - g)) & 1u)) bad |= 1u << (4 - g);
            }
            r.parity_bits = (uint8_t)bad; r.parity_ok = bad == 0;
            r.continuous = (prev < 0) || (r.counter == (unsigned)((prev + 1) % 320));
            prev = r.counter;
        }
        quality_out[f] = r;
    }
    return PDT_OK;
}

int pdt_synth_poes_stream_device(void *d_iq, int pcm16, uint64_t start_sample, uint64_t n_samples, uint64_t total_samples,
                                 double sample_rate, uint64_t seed, void *stream)
{
    if (!device_ok()) return PDT_ENODEV;
    if (!d_iq || !n_samples || start_sample + n_samples > total_samples) return fail(PDT_EINVAL, "bad arguments");
    return synth_poes_launch(d_iq, pcm16, 1, n_samples, n_samples, sample_rate, seed, (cudaStream_t)stream, start_sample, total_samples);
}

int pdt_synth_poes_device(void *d_iq, int pcm16, uint32_t n_captures, uint64_t stride_samples, uint64_t n_samples,
                          double sample_rate, uint64_t seed, void *stream)
{
    if (!d_iq || !n_captures || n_samples > stride_samples) return fail(PDT_EINVAL, "bad arguments");
    if (!device_ok()) return PDT_ENODEV;
    return synth_poes_launch(d_iq, pcm16, n_captures, stride_samples, n_samples, sample_rate, seed, (cudaStream_t)stream);
}

} // extern "C"
