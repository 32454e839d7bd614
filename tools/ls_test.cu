// tools/ls_test.cu — standalone check of the TMA lane-stream helper (pdt_lanestream.cuh).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -I../project-desert-tortoise_b200/csrc -I../include -o ls_test ls_test.cu
#include <cstdio>
#include <vector>
#include "pdt_lanestream.cuh"
using namespace pdt; using namespace pdt::tiled;

struct AddStep {
    float acc;
    __device__ __forceinline__ void quad(const float4 &v, float4 &o) { o.x = v.x + acc; acc += 1.f; o.y = v.y + acc; acc += 1.f; o.z = v.z + acc; acc += 1.f; o.w = v.w + acc; acc += 1.f; }
    __device__ __forceinline__ float one(float v) { float o = v + acc; acc += 1.f; return o; }
};

__global__ void k(const float *in, float *out, const u64 *s0, const u64 *sb, const u64 *s1, float *accs, int rounds)
{
    extern __shared__ __align__(128) unsigned char raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    LaneStream sm = lane_stream_init(&reinterpret_cast<LaneStreamSmem *>(raw)[wib], lane);
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    AddStep st; st.acc = 0.f;
    for (int r = 0; r < rounds; r++) {
        lane_stream<false>(sm, lane, in, out, s0[gid], sb[gid], sb[gid], st);
        lane_stream<true>(sm, lane, in, out, sb[gid], sb[gid], s1[gid], st);
    }
    accs[gid] = st.acc;
}

int main()
{
    const int warps = 2, blocks = 3, nl = warps * 32 * blocks;
    const u64 N = 1 << 20;
    std::vector<float> hin(N); for (u64 i = 0; i < N; i++) hin[i] = (float)(i % 1000) * 0.5f;
    std::vector<u64> s0(nl), sb(nl), s1(nl);
    u64 pos = 0;
    for (int l = 0; l < nl; l++) {
        u64 warm = (l % 5 == 0) ? 0 : 4 * (u64)((l * 37) % 300);
        u64 len = (l % 7 == 0) ? 0 : (u64)((l * 131) % 3000) + (l % 3);     // ragged, some empty, some not multiple of 4
        if (l % 11 == 0) { warm = 0; len = 0; }
        s0[l] = pos; sb[l] = pos + warm; s1[l] = sb[l] + len;
        pos = (s1[l] + 3 + 64) & ~3ull;
    }
    printf("total span %llu of %llu\n", (unsigned long long)pos, (unsigned long long)N);
    float *din, *dout, *dacc; u64 *d0, *db, *d1;
    cudaMalloc(&din, N * 4); cudaMalloc(&dout, N * 4); cudaMalloc(&dacc, nl * 4);
    cudaMalloc(&d0, nl * 8); cudaMalloc(&db, nl * 8); cudaMalloc(&d1, nl * 8);
    cudaMemcpy(din, hin.data(), N * 4, cudaMemcpyHostToDevice); cudaMemset(dout, 0, N * 4);
    cudaMemcpy(d0, s0.data(), nl * 8, cudaMemcpyHostToDevice); cudaMemcpy(db, sb.data(), nl * 8, cudaMemcpyHostToDevice); cudaMemcpy(d1, s1.data(), nl * 8, cudaMemcpyHostToDevice);
    const size_t smem = warps * sizeof(LaneStreamSmem);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<blocks, warps * 32, smem>>>(din, dout, d0, db, d1, dacc, 1);   // (rounds > 1 would need per-round expectations)
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> hout(N), hacc(nl);
    cudaMemcpy(hout.data(), dout, N * 4, cudaMemcpyDeviceToHost); cudaMemcpy(hacc.data(), dacc, nl * 4, cudaMemcpyDeviceToHost);
    long bad = 0;
    for (int l = 0; l < nl; l++) {
        float acc = 0.f;
        for (u64 i = s0[l]; i < sb[l]; i++) acc += 1.f;
        for (u64 i = sb[l]; i < s1[l]; i++) { float want = hin[i] + acc; acc += 1.f; if (hout[i] != want) { if (bad < 5) printf("lane %d i %llu got %f want %f\n", l, (unsigned long long)i, hout[i], want); bad++; } }
        if (hacc[l] != acc) { if (bad < 5) printf("lane %d acc %f want %f\n", l, hacc[l], acc); bad++; }
        for (u64 i = s0[l]; i < sb[l]; i++) if (hout[i] != 0.f) { if (bad < 5) printf("lane %d warm region written at %llu\n", l, (unsigned long long)i); bad++; }
    }
    printf("bad = %ld\n", bad);
    return bad != 0;
}
