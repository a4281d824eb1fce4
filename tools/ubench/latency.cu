// Dependent-chain latency of FFMA vs FFMA2 vs MUFU.EX2 vs SHFL vs LDS on B200 (one warp, clock64 around a long chain).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long f2u(float2 v) { return *reinterpret_cast<unsigned long long*>(&v); }
__device__ __forceinline__ float2 u2f(unsigned long long v) { return *reinterpret_cast<float2*>(&v); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(f2u(a)), "l"(f2u(b)), "l"(f2u(c))); return u2f(d);
}
__device__ __forceinline__ float ex2a(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int MODE> __global__ void k(float* out, long long* cyc) {
    __shared__ float sm[1024];
    for (int i = threadIdx.x; i < 1024; i += 32) sm[i] = (float)((i * 7 + 3) & 1023);
    __syncwarp();
    float v = threadIdx.x * 1e-3f; float2 w = make_float2(v, v + 1.f);
    const float2 a = make_float2(0.999f, 1.001f), b = make_float2(1e-3f, 2e-3f);
    int idx = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < 1024; ++i) {
        if (MODE == 0) v = fmaf(v, 0.999f, 1e-3f);
        if (MODE == 1) w = fma2(w, a, b);
        if (MODE == 2) v = ex2a(v);
        if (MODE == 3) v = __shfl_up_sync(0xffffffffu, v, 1, 16);
        if (MODE == 4) { idx = (int)sm[idx & 1023]; }
        if (MODE == 5) { float4 t = *reinterpret_cast<const float4*>(&sm[(idx & 255) * 4]); idx = (int)(t.x + t.y + t.z + t.w) & 1023; }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = v + w.x + w.y + idx;
}
template <int MODE> void run(const char* n) {
    float* d; long long* c; cudaMalloc(&d, 256); cudaMalloc(&c, 8);
    k<MODE><<<1, 32>>>(d, c); cudaDeviceSynchronize(); k<MODE><<<1, 32>>>(d, c); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("%-12s %.2f cycles per dependent op\n", n, h / 1024.0);
}
int main() { run<0>("FFMA"); run<1>("FFMA2"); run<2>("MUFU.EX2"); run<3>("SHFL.UP"); run<4>("LDS.32+cvt"); run<5>("LDS.128+add"); return 0; }
