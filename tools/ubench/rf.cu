// Register-operand throughput of fp32 ops on B200: 3-register FFMA vs immediate forms, packed FFMA2/FMUL2 with pair / scalar operands.
// Independent accumulators (8 per thread), operands are live registers loaded from memory (no immediates / no CSE).
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
__device__ __forceinline__ unsigned long long f2u(float2 v) { return *reinterpret_cast<unsigned long long*>(&v); }
__device__ __forceinline__ float2 u2f(unsigned long long v) { return *reinterpret_cast<float2*>(&v); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(f2u(a)), "l"(f2u(b)), "l"(f2u(c))); return u2f(d); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { unsigned long long d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2u(a)), "l"(f2u(b))); return u2f(d); }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float mul1(float a, float b) { float d; asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }

template <int MODE> __global__ void k(const float* in, float* out) {
    float v[8], a[8], b[8]; float2 w[8], p[8], q[8];
    for (int i = 0; i < 8; ++i) { v[i] = in[threadIdx.x + 32 * i]; a[i] = in[threadIdx.x + 256 + 32 * i]; b[i] = in[threadIdx.x + 512 + 32 * i];
        w[i] = make_float2(v[i], a[i]); p[i] = make_float2(a[i], b[i]); q[i] = make_float2(b[i], v[i]); }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) v[i] = fma1(v[i], a[i], b[i]);                 // FFMA  d = d*a + b   (3 distinct regs)
            if (MODE == 1) v[i] = fma1(a[i], b[(i + 1) & 7], v[i]);       // FFMA  d = a*b' + d  (3 distinct regs, acc)
            if (MODE == 2) v[i] = mul1(v[i], a[i]);                       // FMUL reg*reg
            if (MODE == 3) w[i] = fma2(w[i], p[i], q[i]);                 // FFMA2 3 distinct pairs
            if (MODE == 4) w[i] = mul2(w[i], p[i]);                       // FMUL2 pair*pair
            if (MODE == 5) w[i] = mul2(w[i], make_float2(a[i], a[i]));    // FMUL2 pair*scalar-broadcast
            if (MODE == 6) w[i] = fma2(p[i], make_float2(a[i], a[i]), w[i]);   // FFMA2 pair*scalar + acc
            if (MODE == 7) v[i] = fmaf(v[i], 0.999f, 0.001f);            // FFMA immediates
        }
    }
    float s = 0; for (int i = 0; i < 8; ++i) s += v[i] + w[i].x + w[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name) {
    float *in, *out; cudaMalloc(&in, 4096 * 4); cudaMemset(in, 0, 4096 * 4); cudaMalloc(&out, 148 * 4 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 4, 256>>>(in, out); cudaDeviceSynchronize();
    cudaEventRecord(e0); for (int r = 0; r < 5; ++r) k<MODE><<<148 * 4, 256>>>(in, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    double wi = 148.0 * 4 * 8 * ITERS * 8, cyc = ms * 1e-3 * 1.965e9;
    printf("%-34s %.3f warp-inst/clk/SM  (%.2f cycles per warp-inst per SMSP)\n", name, wi / cyc / 148, 4.0 / (wi / cyc / 148));
}
int main() {
    run<0>("FFMA d=d*a+b (3 regs)"); run<1>("FFMA d=a*b'+d (3 regs)"); run<2>("FMUL reg*reg"); run<3>("FFMA2 3 pairs");
    run<4>("FMUL2 pair*pair"); run<5>("FMUL2 pair*scalar"); run<6>("FFMA2 pair*scalar+acc"); run<7>("FFMA immediates");
    return 0;
}
