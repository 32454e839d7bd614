// tools/microbench.cu — single-warp latency probes for the serial cores of the tiled engine (B200, sm_100a).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -I../project-desert-tortoise_b200/csrc -o microbench microbench.cu
#include <cstdio>
#include "pdt_tiled.cuh"
using namespace pdt;
using namespace pdt::tiled;

#define N_IT 4096
__global__ void probe(float *out, long long *cyc, const float *in, int active_lanes)
{
    __shared__ float sm[N_IT];
    for (int i = threadIdx.x; i < N_IT; i += blockDim.x) sm[i] = in[i];
    __syncthreads();
    if ((int)threadIdx.x >= active_lanes) return;
    long long t0, t1; int k = 0;
    float x = in[0] + threadIdx.x * 1e-3f, y = in[1];
    double d = in[2];
    // 0: dependent FADD chain
    t0 = clock64(); for (int i = 0; i < N_IT; i++) x = x + y; t1 = clock64(); cyc[k++] = t1 - t0;
    // 1: dependent FMUL+FADD (no fma)
    t0 = clock64(); for (int i = 0; i < N_IT; i++) x = x * 0.999f + y; t1 = clock64(); cyc[k++] = t1 - t0;
    // 2: dependent DADD
    t0 = clock64(); for (int i = 0; i < N_IT; i++) d = d + (double)y; t1 = clock64(); cyc[k++] = t1 - t0;
    // 3: dependent DMUL+DADD
    t0 = clock64(); for (int i = 0; i < N_IT; i++) d = d * 0.99999 + 1e-5; t1 = clock64(); cyc[k++] = t1 - t0;
    // 4: float EMA through double like the reference: f = (float)((double)f*(1-a) + (double)t)
    t0 = clock64(); for (int i = 0; i < N_IT; i++) x = (float)((double)x * (1.0 - 0.00005f) + (double)y); t1 = clock64(); cyc[k++] = t1 - t0;
    // 5: same EMA reading its input from shared memory
    t0 = clock64(); for (int i = 0; i < N_IT; i++) x = (float)((double)x * (1.0 - 0.00005f) + (double)sm[i]); t1 = clock64(); cyc[k++] = t1 - t0;
    // 6: pll_track_step chain, sp from shared memory
    { TrackConst kc; kc.alpha = 1.04e-3f; kc.beta = 2.7e-7f; kc.max_freq = 0.113f; kc.min_freq = -0.113f;
      float ph = x, fr = 0.01f;
      t0 = clock64(); for (int i = 0; i < N_IT; i++) pll_track_step(ph, fr, sm[i], kc); t1 = clock64(); cyc[k++] = t1 - t0; x += ph + fr; }
    // 7: agc_step chain from shared memory
    { AgcState st; st.init = 1; st.gain = 5.0f; float acc = 0;
      t0 = clock64(); for (int i = 0; i < N_IT; i++) acc += agc_step(st, sm[i], 2e-3f, 4e-3f); t1 = clock64(); cyc[k++] = t1 - t0; x += acc + st.gain; }
    // 8: sincos_exact throughput-ish (dependent through the argument)
    { float s, c, a = x;
      t0 = clock64(); for (int i = 0; i < N_IT / 8; i++) { sincos_exact(a, s, c); a = s + c; } t1 = clock64(); cyc[k++] = (t1 - t0) * 8; x += a; }
    // 9: arctan2_approx dependent
    { float a = x;
      t0 = clock64(); for (int i = 0; i < N_IT / 8; i++) { a = arctan2_approx(a, sm[i]); } t1 = clock64(); cyc[k++] = (t1 - t0) * 8; x += a; }
    // 10: dependent LDS chain (pointer chasing)
    { int idx = ((int)x) & 1;
      t0 = clock64(); for (int i = 0; i < N_IT; i++) idx = ((int)sm[idx]) & (N_IT - 1); t1 = clock64(); cyc[k++] = t1 - t0; x += idx; }
    // 11: F2D + D2F round trip chain
    t0 = clock64(); for (int i = 0; i < N_IT; i++) x = (float)((double)x + 1e-9); t1 = clock64(); cyc[k++] = t1 - t0;
    // 12: rintf + F2I + LDS (gardner pick)
    { float nx = 3.0f; float acc = 0;
      t0 = clock64(); for (int i = 0; i < N_IT / 4; i++) { unsigned at = (unsigned)rintf(nx); float v = sm[at & (N_IT - 1)]; nx = nx - v * 1e-3f + 3.7f; if (nx > 4000.f) nx -= 4000.f; acc += v; } t1 = clock64(); cyc[k++] = (t1 - t0) * 4; x += acc; }
    out[threadIdx.x] = x + (float)d;
}

int main()
{
    float *in, *out; long long *cyc;
    cudaMallocManaged(&in, N_IT * 4); cudaMallocManaged(&out, 1024 * 4); cudaMallocManaged(&cyc, 64 * 8);
    for (int i = 0; i < N_IT; i++) in[i] = 0.3f * sinf(0.37f * i) + 0.1f;
    const char *names[] = {"FADD chain", "FMUL+FADD chain", "DADD chain", "DMUL+DADD chain", "float EMA via double (reg)", "float EMA via double (LDS in)",
                           "pll_track_step (LDS in)", "agc_step (LDS in)", "sincos_exact dependent", "arctan2_approx dependent", "LDS pointer chase", "F2D+DADD+D2F chain",
                           "gardner pick (rint+F2I+LDS+3 flop)"};
    for (int lanes : {1, 32}) {
        for (int rep = 0; rep < 2; rep++) { probe<<<1, 32>>>(out, cyc, in, lanes); cudaDeviceSynchronize(); }
        printf("active lanes = %d\n", lanes);
        for (int k = 0; k < 13; k++) printf("  %-38s %7.1f cycles/iter\n", names[k], (double)cyc[k] / N_IT);
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
