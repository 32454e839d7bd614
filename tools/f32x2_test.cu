// tools/f32x2_test.cu — are Blackwell's packed fp32 ops (mul.rn.f32x2 / add.rn.f32x2) kept UNFUSED by ptxas?
// The FIR must round every product and every sum separately (the reference is built without FMA contraction).
#include <cstdio>
#include <cstdlib>
__device__ __forceinline__ unsigned long long pk(float a, float b) { return ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(a); }
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) { unsigned long long r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) { unsigned long long r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__global__ void k(const float *h, const float *x, float *out_pk, float *out_sc, float *out_fma, int n)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long acc = pk(0.f, 0.f);
    float s0 = 0.f, s1 = 0.f, f0 = 0.f, f1 = 0.f;
    for (int i = 0; i < n; i++) {
        const float a0 = h[i], a1 = h[i + 1], b0 = x[t + i], b1 = x[t + i + 7];
        acc = add2(acc, mul2(pk(a0, a1), pk(b0, b1)));
        s0 = __fadd_rn(s0, __fmul_rn(a0, b0)); s1 = __fadd_rn(s1, __fmul_rn(a1, b1));
        f0 = __fmaf_rn(a0, b0, f0); f1 = __fmaf_rn(a1, b1, f1);
    }
    out_pk[2 * t] = __uint_as_float((unsigned)acc); out_pk[2 * t + 1] = __uint_as_float((unsigned)(acc >> 32));
    out_sc[2 * t] = s0; out_sc[2 * t + 1] = s1; out_fma[2 * t] = f0; out_fma[2 * t + 1] = f1;
}
int main()
{
    const int n = 26, T = 1 << 16;
    float *h, *x, *a, *b, *c;
    cudaMallocManaged(&h, (n + 1) * 4); cudaMallocManaged(&x, (T + n + 8) * 4);
    cudaMallocManaged(&a, 2 * T * 4); cudaMallocManaged(&b, 2 * T * 4); cudaMallocManaged(&c, 2 * T * 4);
    srand(1);
    for (int i = 0; i <= n; i++) h[i] = (rand() / (float)RAND_MAX - 0.5f) * 0.3f;
    for (int i = 0; i < T + n + 8; i++) x[i] = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
    k<<<T / 256, 256>>>(h, x, a, b, c, n);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed\n"); return 1; }
    long d_sc = 0, d_fma = 0;
    for (int i = 0; i < 2 * T; i++) { d_sc += a[i] != b[i]; d_fma += a[i] != c[i]; }
    printf("packed vs separately rounded: %ld differences; packed vs fused: %ld differences (of %d)\n", d_sc, d_fma, 2 * T);
    return 0;
}
