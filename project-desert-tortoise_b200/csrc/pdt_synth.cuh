// pdt_synth.cuh — seeded synthetic POES-TIP IQ captures generated ON THE DEVICE (bench / scale-test workloads).
//
// Signal model of SURVEY.md §8(d) (same as tests/synth_ref.py, which is the numpy twin used for small cases):
//   104-byte minor frames (ED E2, spacecraft id, 9-bit counter in bytes 4-5, parity word 103 per
//   standalone_matlab/Functionized/checkParity.m:20-90) -> split-phase chips at 16640.3 chips/s
//   -> x[n] = A·exp(j(2π(f0·t + ½·ḟ·t²) + θ0 + 1.169·chip)) + σ·CN(0,1)
// Per-capture amplitude / Doppler / drift / SNR / frame phase derive from hash(seed, capture).
#pragma once

#include "pdt_common.cuh"

namespace pdt {

PDT_DEV unsigned long long mix64(unsigned long long z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
PDT_DEV double u01(unsigned long long h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); }

struct SynthCap { double f0, drift, amp, sigma, theta0; unsigned bit_start, n_frames, counter0; };

constexpr int SYNTH_FRAME_BYTES = 104;

// stream_seconds > 0: ONE long stream (capture 0) — the linear drift is chosen so that the carrier crosses from f0 to -f0
// over the whole stream (a pass-like Doppler S-curve stays inside the PLL's +-4500 Hz range however long the stream is).
__global__ void k_synth_frames(unsigned char *tab, SynthCap *caps, uint32_t n_captures, unsigned frames_per_cap,
                               double sps, unsigned long long seed, double stream_seconds)
{
    const uint32_t c = blockIdx.x;
    if (c >= n_captures) return;
    const unsigned long long hc = mix64(seed * 0x100000001B3ull + c);
    if (threadIdx.x == 0) {
        SynthCap sc;
        sc.f0 = (u01(mix64(hc + 1)) * 2.0 - 1.0) * 3000.0;
        sc.drift = (u01(mix64(hc + 2)) * 2.0 - 1.0) * 50.0;
        sc.amp = 0.05 + 0.45 * u01(mix64(hc + 3));
        const double esn0_db[4] = {20.0, 18.0, 16.0, 14.0};
        const double esn0 = pow(10.0, esn0_db[c & 3] / 10.0);
        sc.sigma = sc.amp * sin(1.169) * sqrt(sps / esn0);
        sc.theta0 = u01(mix64(hc + 4));                       // in turns
        sc.n_frames = frames_per_cap;
        sc.bit_start = (unsigned)(mix64(hc + 5) % (unsigned long long)(2 * SYNTH_FRAME_BYTES * 8u));   // < 2 frames: counter stays continuous
        sc.counter0 = (unsigned)(mix64(hc + 6) % 320u);
        if (stream_seconds > 0.0) sc.drift = -2.0 * sc.f0 / stream_seconds;
        caps[c] = sc;
    }
    const unsigned counter0 = (unsigned)(mix64(hc + 6) % 320u);
    for (unsigned f = threadIdx.x; f < frames_per_cap; f += blockDim.x) {
        unsigned char *fr = tab + ((size_t)c * frames_per_cap + f) * SYNTH_FRAME_BYTES;
        for (int k = 0; k < SYNTH_FRAME_BYTES; k += 8) {
            unsigned long long r = mix64(hc ^ (0xABCDull + (unsigned long long)f * 131 + k));
            for (int b = 0; b < 8 && k + b < SYNTH_FRAME_BYTES; b++) fr[k + b] = (unsigned char)(r >> (8 * b));
        }
        fr[0] = 0xED; fr[1] = 0xE2; fr[2] = 0x08;
        const unsigned cnt = (counter0 + f) % 320u;
        fr[4] = (unsigned char)((fr[4] & 0xFE) | (cnt >> 8));
        fr[5] = (unsigned char)(cnt & 0xFF);
        unsigned w = fr[103] & 0xC1u;
        const int lo[5] = {2, 19, 36, 53, 70}, hi[5] = {18, 35, 52, 69, 86};
        for (int g = 0; g < 5; g++) {
            unsigned ones = 0;
            for (int k = lo[g]; k <= hi[g]; k++) ones += __popc((unsigned)fr[k]);
            w |= (ones & 1u) << (5 - g);
        }
        fr[103] = (unsigned char)w;
    }
}

template <typename OUT>
__global__ void k_synth_iq(OUT *iq, const unsigned char *tab, const SynthCap *caps, unsigned long long stride,
                           unsigned long long n, double fs, double sps, unsigned long long seed, int pcm16,
                           unsigned long long i0)           // i0: stream index of the first sample written (slices of one stream)
{
    const uint32_t c = blockIdx.y;
    const SynthCap sc = caps[c];
    const unsigned long long hc = mix64(seed * 0x100000001B3ull + c) ^ 0x5EEDull;
    const unsigned total_bits = sc.n_frames * SYNTH_FRAME_BYTES * 8u;
    const unsigned char *ft = tab + (size_t)c * sc.n_frames * SYNTH_FRAME_BYTES;
    for (unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; k < n;
         k += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long i = i0 + k;
        const unsigned long long chip = (unsigned long long)floor((double)i / sps);
        const unsigned bitpos = (unsigned)((sc.bit_start + (chip >> 1)) % total_bits);
        const int bit = (ft[bitpos >> 3] >> (7 - (bitpos & 7))) & 1;
        const double d = ((bit != 0) == ((chip & 1) == 0)) ? 1.0 : -1.0;       // '1' = (+,-), '0' = (-,+)
        const double t = (double)i / fs;
        double cyc = sc.f0 * t + 0.5 * sc.drift * t * t + sc.theta0 + d * (1.169 / (2.0 * PDT_PI));
        cyc -= floor(cyc);
        double sn, cs;
        sincospi(2.0 * cyc, &sn, &cs);
        const unsigned long long r = mix64(hc + i * 2), r2 = mix64(hc + i * 2 + 1);
        const double u1 = u01(r) + 1e-300, u2 = u01(r2);
        const double mag = sc.sigma * sqrt(-log(u1));                        // CN(0,σ²): each axis σ²/2
        double n1, n2;
        sincospi(2.0 * u2, &n1, &n2);
        const double re = sc.amp * cs + mag * n2, im = sc.amp * sn + mag * n1;
        OUT *o = iq + ((size_t)c * stride + k) * 2;
        if (pcm16) {
            o[0] = (OUT)fmin(fmax(rint(re * 32768.0), -32768.0), 32767.0);
            o[1] = (OUT)fmin(fmax(rint(im * 32768.0), -32768.0), 32767.0);
        } else { o[0] = (OUT)re; o[1] = (OUT)im; }
    }
}

inline int synth_poes_launch(void *d_iq, int pcm16, uint32_t n_captures, uint64_t stride, uint64_t n, double fs,
                             uint64_t seed, cudaStream_t s, uint64_t stream_start = 0, uint64_t stream_total = 0)
{
    const double sps = fs / (8320 * 2 + 0.3);
    const uint64_t span = stream_total ? stream_total : n;
    const unsigned frames_per_cap = (unsigned)(span / sps / (SYNTH_FRAME_BYTES * 16.0)) + 5;
    unsigned char *tab = nullptr; SynthCap *caps = nullptr;
    PDT_CUDA(cudaMalloc(&tab, (size_t)n_captures * frames_per_cap * SYNTH_FRAME_BYTES));
    PDT_CUDA(cudaMalloc(&caps, sizeof(SynthCap) * n_captures));
    k_synth_frames<<<n_captures, 64, 0, s>>>(tab, caps, n_captures, frames_per_cap, sps, seed, stream_total ? (double)stream_total / fs : 0.0);
    const unsigned bx = (unsigned)std::min<uint64_t>((n + 255) / 256, 2048);
    dim3 grid(bx, n_captures);
    if (pcm16) k_synth_iq<short><<<grid, 256, 0, s>>>((short *)d_iq, tab, caps, stride, n, fs, sps, seed, 1, stream_start);
    else       k_synth_iq<real_t><<<grid, 256, 0, s>>>((real_t *)d_iq, tab, caps, stride, n, fs, sps, seed, 0, stream_start);
    count_launch(2);
    cudaError_t e = cudaGetLastError();
    cudaStreamSynchronize(s);
    cudaFree(tab); cudaFree(caps);
    if (e != cudaSuccess) return fail(PDT_ECUDA, "synth launch: %s", cudaGetErrorString(e));
    return PDT_OK;
}

} // namespace pdt
