// tests/host/sincos_host.cu — the branch-free exact sincos of the kernels (pdt_device.cuh::sincos_core, `__host__ __device__`) run on
// the HOST for EVERY float of ±[2^-13, 8] (the PLL phase lives in [-2π, 2π]) and for samples of the rest of its declared range
// (|y| < 120, tiny values), against this machine's glibc sinf / cosf — the functions the reference calls
// (CarrierTrackingPLL.c:106-107).  The double arithmetic inside is IEEE on both sides, so equality here is equality on the device
// (the GPU parity tests compare whole phase streams besides).  Prints "OK <values>".
#include <cstdio>
#include <cstring>
#include <cmath>
#include <cstdint>
#include "pdt_device.cuh"


static int check(float y, unsigned long long &n)
{
    if (!sincos_in_core_range(y)) return 0;
    float s, c;
    sincos_core(y, s, c);
    const float ws = sinf(y), wc = cosf(y);
    n++;
    if (std::memcmp(&s, &ws, 4) || std::memcmp(&c, &wc, 4)) {
        std::printf("MISMATCH y=%a: sin %a / glibc %a, cos %a / glibc %a\n", y, s, ws, c, wc);
        return 1;
    }
    return 0;
}

int main()
{
    unsigned long long n = 0;
    const uint32_t lo = pdt_f2u(0x1p-13f), hi = pdt_f2u(8.0f);
    for (uint32_t u = lo; u <= hi; u++) {                       // every float of [2^-13, 8], both signs
        const float y = pdt_u2f(u);
        if (check(y, n) || check(-y, n)) return 1;
    }
    for (uint32_t u = pdt_f2u(8.0f); u < pdt_f2u(120.0f); u += 37) {      // the rest of the declared range, sampled
        const float y = pdt_u2f(u);
        if (check(y, n) || check(-y, n)) return 1;
    }
    for (uint32_t u = 1; u < lo; u += 4099) {                   // tiny values down to the subnormals (glibc returns (y, 1) there)
        const float y = pdt_u2f(u);
        if (check(y, n) || check(-y, n)) return 1;
    }
    if (check(0.0f, n)) return 1;
    std::printf("OK %llu\n", n);
    return 0;
}
