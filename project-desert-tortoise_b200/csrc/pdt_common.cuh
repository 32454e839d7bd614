// pdt_common.cuh — host-side plumbing shared by the batch API (pdt_batch.cu) and the legacy shim (pdt_legacy.cu).
#pragma once

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>

#include "pdt_device.cuh"
#include "../../include/pdt.h"

namespace pdt {

extern thread_local char g_err[512];
extern std::atomic<unsigned long long> g_launches;

int  fail(int code, const char *fmt, ...);
bool device_ok();                      // true when a CUDA device is usable (cached)

#define PDT_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return pdt::fail(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver ? PDT_ENODEV : PDT_ECUDA, \
                             "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));   \
    } while (0)

inline void count_launch(unsigned n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---- derived chain constants (everything the kernels need, computed exactly like the reference call sites) ----
struct ChainConst {
    int      argos, L, N, K;            // K = taps per polyphase branch (N/L), ARGOS: K = N
    uint32_t chunk;
    PllParams pll;
    real_t   agc_attack, agc_decay, norm_override;
    int      gardner_fs;
    real_t   baud, g_range, g_kp, man_thresh, squelch;
    SyncParams sync;
    uint32_t max_frames;
    int      prefix_bytes;              // 2 for POES (literal ED E2), 0 for ARGOS
    int      ring_frames;               // live mode: the frame table is a ring (slot = frame number mod max_frames)
    int      use_mm;                    // PDT_CLOCK_MM: MMClockRecovery.c:5-84 instead of the Gardner loop (exact engine)
    real_t   mm_range, mm_kp;
    int      ypad;                      // zero pad behind the interpolated chunk: Gardner's first mid-sample of a chunk reads up
                                        // to step/2 past its end (GardenerClockRecovery.c:28; the reference's buffer is chunk*N long)
};

int  build_chain_const(const pdt_params &p, ChainConst &cc);
void make_lpfir_host(real_t *h, int N, real_t Fc, real_t Fs, int L);   // LowPassFilter.c:127-175

} // namespace pdt
