// Micro-benchmarks of the SM pipes the scan kernels lean on (B200): MUFU.EX2, FFMA, FFMA2, LDS, SHFL.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipes pipes.cu ; prints warp-instructions / clk / SM.
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2048
__device__ __forceinline__ float ex2a(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ unsigned long long f2u(float2 v) { return *reinterpret_cast<unsigned long long*>(&v); }
__device__ __forceinline__ float2 u2f(unsigned long long v) { return *reinterpret_cast<float2*>(&v); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(f2u(a)), "l"(f2u(b)), "l"(f2u(c))); return u2f(d);
}

template <int MODE>
__global__ void k(float* out, int n) {
    __shared__ float sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * 1e-3f;
    __syncthreads();
    float v[8];
    float2 w[8];
    for (int i = 0; i < 8; ++i) { v[i] = threadIdx.x * 1e-3f + i; w[i] = make_float2(v[i], v[i] + 1.f); }
    const float2 a = make_float2(0.999f, 1.001f), b = make_float2(1e-3f, 2e-3f);
    int lane = threadIdx.x & 31;
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 0) {            // MUFU.EX2 x8
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = ex2a(v[i]);
        } else if (MODE == 1) {     // FFMA x8
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i], 0.999f, 1e-3f + n);
        } else if (MODE == 2) {     // FFMA2 x8
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = fma2(w[i], a, b);
        } else if (MODE == 3) {     // LDS.128 broadcast (all lanes same address)
#pragma unroll
            for (int i = 0; i < 8; ++i) { float4 t = *reinterpret_cast<const float4*>(&sm[((it + i) & 255) * 4]); v[i] += t.x + t.w; }
        } else if (MODE == 4) {     // LDS.128 distinct (32 lanes x 16B contiguous)
#pragma unroll
            for (int i = 0; i < 8; ++i) { float4 t = *reinterpret_cast<const float4*>(&sm[(((it + i) & 7) * 32 + lane) * 4]); v[i] += t.x + t.w; }
        } else if (MODE == 5) {     // LDS.32 broadcast
#pragma unroll
            for (int i = 0; i < 8; ++i) { v[i] += sm[(it + i) & 1023]; }
        } else if (MODE == 6) {     // SHFL
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __shfl_up_sync(0xffffffffu, v[i], 1, 16) + 1.f;
        } else if (MODE == 7) {     // mixed: 4 MUFU + 8 FFMA2
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = ex2a(v[i]);
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = fma2(w[i], a, b);
        } else if (MODE == 8) {     // LDS.64 broadcast pairs (16 distinct 8B addresses, 2 lanes each)
#pragma unroll
            for (int i = 0; i < 8; ++i) { float2 t = *reinterpret_cast<const float2*>(&sm[(((it + i) & 63) * 16 + (lane >> 1)) * 2]); v[i] += t.x + t.y; }
        } else if (MODE == 9) {     // LDS.128: 8 distinct addresses (4 lanes each share) -- the fwd16 B/C pattern
#pragma unroll
            for (int i = 0; i < 8; ++i) { float4 t = *reinterpret_cast<const float4*>(&sm[(((it + i) & 31) * 8 + (lane & 7)) * 4]); v[i] += t.x + t.w; }
        }
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) s += v[i] + w[i].x + w[i].y;
    if (s == 123.456f) out[0] = s;
}

template <int MODE>
void run(const char* name, int inst_per_iter) {
    float* d; cudaMalloc(&d, 4);
    int dev_sms = 148, ctas = dev_sms * 4, thr = 256;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<ctas, thr>>>(d, 0); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k<MODE><<<ctas, thr>>>(d, 0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double warp_inst = (double)ctas * (thr / 32) * ITERS * inst_per_iter;
    double cycles = ms * 1e-3 * clk_khz * 1e3;
    printf("%-28s %8.3f ms  %.3f warp-inst/clk/SM (at %d MHz nominal)  -> %.1f lanes/clk/SM\n", name, ms, warp_inst / cycles / dev_sms, clk_khz / 1000,
           warp_inst / cycles / dev_sms * 32);
    cudaFree(d);
}

int main() {
    run<0>("MUFU.EX2", 8);
    run<1>("FFMA", 8);
    run<2>("FFMA2", 8);
    run<3>("LDS.128 broadcast", 8);
    run<4>("LDS.128 distinct", 8);
    run<5>("LDS.32 broadcast", 8);
    run<6>("SHFL.UP", 8);
    run<7>("4 MUFU + 8 FFMA2", 12);
    run<8>("LDS.64 16 addr", 8);
    run<9>("LDS.128 8 addr", 8);
    return 0;
}
