// fm_common.cuh -- device helpers shared by the sm_100a scan kernels.
//
// Math contract (restated from the reference, not copied):
//   softplus threshold 20, log1p(exp(x))          selective_scan_fwd_kernel.cuh:153-156
//   a = exp2(delta * A * log2(e))                  selective_scan_fwd_kernel.cuh:169-171, 216
//   scan monoid (a0,b0)o(a1,b1) = (a1*a0, a1*b0+b1) selective_scan_common.h:110-115
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fm_scan.h"

namespace fm {

constexpr float kLog2e = 1.4426950408889634f;
// A lane owns S consecutive timesteps of a chunk (S = 8 or 16).  In shared memory a lane segment is padded to
// S + 4 floats so that the G float4 reads of a quarter warp fall in distinct bank groups (conflict-free LDS.128).
__host__ __device__ constexpr int seg_pad(int S) { return S == 4 ? 4 : S + 4; }

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// exp(-|x|) with the exponent scaled in two pieces (log2(e) = hi + lo) so that large |x| keep full relative accuracy
__device__ __forceinline__ float exp_neg_abs(float x) {
    const float ax = fabsf(x);
    return ex2_approx(fmaf(-ax, 1.4426950216293335f, -ax * 1.9259629911266175e-8f));
}

// Branch-free softplus with the reference's semantics (F.softplus / log1pf(expf(x)), threshold 20):
//   softplus(x) = max(x, 0) + log1p(exp(-|x|)),   log1p(t) = 2 atanh(t / (2 + t)),  t in (0, 1] -> s <= 1/3,
// odd series in s truncated at s^13 (remainder < 2e-8 relative).  Max relative error vs float64 ~3e-7; for x > 20 the
// correction is below half an ulp of x, so the result equals x exactly like the reference's threshold branch.
__device__ __forceinline__ float softplus_fast(float x) {
    const float t = exp_neg_abs(x);
    const float s = t * rcp_approx(2.f + t);
    const float s2 = s * s;
    float p = fmaf(s2, 1.f / 13.f, 1.f / 11.f);
    p = fmaf(s2, p, 1.f / 9.f);
    p = fmaf(s2, p, 1.f / 7.f);
    p = fmaf(s2, p, 1.f / 5.f);
    p = fmaf(s2, p, 1.f / 3.f);
    p = fmaf(s2, p, 1.f);
    return fmaf(2.f * s, p, fmaxf(x, 0.f));
}

// 1 - exp(-d) for d >= 0 with full relative accuracy: alternating series up to d^6/720 below 0.125 (remainder < 3e-9 relative),
// 1 - ex2(-d log2 e) above (cancellation there costs < 2^-21 relative).  Equals sigmoid(x) when d = softplus(x).
__device__ __forceinline__ float one_minus_exp_neg(float d) {
    float p = fmaf(d, -1.f / 720.f, 1.f / 120.f);
    p = fmaf(d, p, -1.f / 24.f);
    p = fmaf(d, p, 1.f / 6.f);
    p = fmaf(d, p, -0.5f);
    p = fmaf(d, p, 1.f);
    const float small = d * p;
    const float big = 1.f - ex2_approx(fmaf(-d, 1.4426950216293335f, -d * 1.9259629911266175e-8f));
    return d < 0.125f ? small : big;
}

// sigmoid(x) = 1/(1+e^-x), evaluated through e^-|x| so that both tails keep full relative accuracy
__device__ __forceinline__ float sigmoid_f(float x) {
    const float t = exp_neg_abs(x);
    const float r = rcp_approx(1.f + t);
    return x >= 0.f ? r : t * r;
}

// ---------------------------------------------------------------------------------------------
// Blackwell packed fp32 math (sm_100 FFMA2 / FMUL2 / FADD2): two IEEE fp32 operations per issue slot, bit-identical
// to the scalar fma.rn / mul.rn / add.rn.  The kernels are issue-bound, so the time-adjacent element pairs of a lane
// segment are processed with these wherever the recurrence does not serialise them.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long& as_u64(float2& v) { return *reinterpret_cast<unsigned long long*>(&v); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(as_u64(d)) : "l"(as_u64(a)), "l"(as_u64(b)), "l"(as_u64(c)));
    return d;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    float2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(as_u64(d)) : "l"(as_u64(a)), "l"(as_u64(b)));
    return d;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    float2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(as_u64(d)) : "l"(as_u64(a)), "l"(as_u64(b)));
    return d;
}
__device__ __forceinline__ float2 bcast2(float s) { return make_float2(s, s); }   // SASS: scalar-broadcast operand (R.F32)
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ float4 lds128(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void sts128(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// ---------------------------------------------------------------------------------------------
// dtype traits
// ---------------------------------------------------------------------------------------------
template <typename T> struct Cvt;
template <> struct Cvt<float> {
    static __device__ __forceinline__ float to_f(float v) { return v; }
    static __device__ __forceinline__ float from_f(float v) { return v; }
};
template <> struct Cvt<__half> {
    static __device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
    static __device__ __forceinline__ __half from_f(float v) { return __float2half_rn(v); }
};
template <> struct Cvt<__nv_bfloat16> {
    static __device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
    static __device__ __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};

// Load kSeg consecutive elements starting at p into f[] (fp32). `nvalid` = elements that exist
// (<= 0: none); `vec` = pointer and row pitch are 16-byte aligned. Missing elements read as 0.
template <typename T, int kSeg>
__device__ __forceinline__ void load_seg(const T* __restrict__ p, int nvalid, bool vec, float (&f)[kSeg]) {
    if (vec && nvalid >= kSeg) {
        if constexpr (sizeof(T) == 4) {
            const float4* q = reinterpret_cast<const float4*>(p);
#pragma unroll
            for (int i = 0; i < kSeg / 4; ++i) {
                float4 v = __ldg(q + i);
                f[4 * i + 0] = v.x; f[4 * i + 1] = v.y; f[4 * i + 2] = v.z; f[4 * i + 3] = v.w;
            }
        } else if constexpr (kSeg % 8 == 0) {
            const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
            for (int i = 0; i < kSeg / 8; ++i) {
                uint4 v = __ldg(q + i);
                const T* e = reinterpret_cast<const T*>(&v);
#pragma unroll
                for (int j = 0; j < 8; ++j) f[8 * i + j] = Cvt<T>::to_f(e[j]);
            }
        } else {
            const uint2* q = reinterpret_cast<const uint2*>(p);
#pragma unroll
            for (int i = 0; i < kSeg / 4; ++i) {
                uint2 v = __ldg(q + i);
                const T* e = reinterpret_cast<const T*>(&v);
#pragma unroll
                for (int j = 0; j < 4; ++j) f[4 * i + j] = Cvt<T>::to_f(e[j]);
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < kSeg; ++i) f[i] = (i < nvalid) ? Cvt<T>::to_f(p[i]) : 0.f;
    }
}

// Raw (unconverted) segment: the bits of kSeg elements as loaded.  Kernels that software-pipeline their global loads
// keep these in loop-carried registers and widen at the point of use -- converting right behind the load would make
// the warp wait for the data there.
template <typename T, int kSeg>
struct SegRaw { uint32_t w[kSeg * sizeof(T) / 4]; };

template <typename T, int kSeg>
__device__ __forceinline__ void load_seg_raw(const T* __restrict__ p, int nvalid, bool vec, SegRaw<T, kSeg>& r) {
    constexpr int NWD = kSeg * sizeof(T) / 4;
    if (vec && nvalid >= kSeg) {
        if constexpr (NWD % 4 == 0) {
            const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
            for (int i = 0; i < NWD / 4; ++i) {
                const uint4 v = __ldg(q + i);
                r.w[4 * i] = v.x; r.w[4 * i + 1] = v.y; r.w[4 * i + 2] = v.z; r.w[4 * i + 3] = v.w;
            }
        } else {
            const uint2* q = reinterpret_cast<const uint2*>(p);
#pragma unroll
            for (int i = 0; i < NWD / 2; ++i) {
                const uint2 v = __ldg(q + i);
                r.w[2 * i] = v.x; r.w[2 * i + 1] = v.y;
            }
        }
    } else {
        T e[kSeg];
#pragma unroll
        for (int i = 0; i < kSeg; ++i) e[i] = (i < nvalid) ? p[i] : Cvt<T>::from_f(0.f);
#pragma unroll
        for (int i = 0; i < NWD; ++i) r.w[i] = reinterpret_cast<const uint32_t*>(e)[i];
    }
}

template <typename T, int kSeg>
__device__ __forceinline__ void widen_seg(const SegRaw<T, kSeg>& r, float (&f)[kSeg]) {
    if constexpr (sizeof(T) == 4) {
#pragma unroll
        for (int i = 0; i < kSeg; ++i) f[i] = __uint_as_float(r.w[i]);
    } else {
#pragma unroll
        for (int i = 0; i < kSeg / 2; ++i) {
            const uint32_t w = r.w[i];
            const T* e = reinterpret_cast<const T*>(&w);
            f[2 * i] = Cvt<T>::to_f(e[0]);
            f[2 * i + 1] = Cvt<T>::to_f(e[1]);
        }
    }
}

template <typename T, int kSeg>
__device__ __forceinline__ void store_seg(T* __restrict__ p, int nvalid, bool vec, const float (&f)[kSeg]) {
    if (vec && nvalid >= kSeg) {
        if constexpr (sizeof(T) == 4) {
            float4* q = reinterpret_cast<float4*>(p);
#pragma unroll
            for (int i = 0; i < kSeg / 4; ++i) q[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
        } else if constexpr (kSeg % 8 == 0) {
            uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
            for (int i = 0; i < kSeg / 8; ++i) {
                uint4 v;
                T* e = reinterpret_cast<T*>(&v);
#pragma unroll
                for (int j = 0; j < 8; ++j) e[j] = Cvt<T>::from_f(f[8 * i + j]);
                q[i] = v;
            }
        } else {
            uint2* q = reinterpret_cast<uint2*>(p);
#pragma unroll
            for (int i = 0; i < kSeg / 4; ++i) {
                uint2 v;
                T* e = reinterpret_cast<T*>(&v);
#pragma unroll
                for (int j = 0; j < 4; ++j) e[j] = Cvt<T>::from_f(f[4 * i + j]);
                q[i] = v;
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < kSeg; ++i)
            if (i < nvalid) p[i] = Cvt<T>::from_f(f[i]);
    }
}

// ---------------------------------------------------------------------------------------------
// cp.async (LDGSTS) helpers for the fp32 B/C tile
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
    unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// Stage one [dstate][TC] tile of a (batch, group, dstate, L) tensor into smem as fp32 with the
// lane-segment-padded layout: element (n, tt) lives at n*rowp + (tt/S)*(S+4) + tt%S.
// Out-of-range timesteps are zero-filled (=> b = 0 and C*h contributes nothing).
// Swizzled alternative to the padded layout for 8-element segments (kSwz): no padding, the two 16-byte halves of segment
// s swap places when (s >> 2) & 1.  A quarter-warp of lanes reading "its" half i of 8 consecutive segments then covers
// all 32 banks once -- conflict-free 128-bit accesses at 2/3 of the padded footprint.
template <int kSeg, bool kSwz>
__host__ __device__ constexpr int tile_seg_pitch() { return kSwz ? kSeg : seg_pad(kSeg); }
template <int kSeg, bool kSwz>
__device__ __forceinline__ int tile_off(int tt) {           // offset of timestep tt (within the chunk) inside a state row
    const int sg = tt / kSeg, r = tt % kSeg;
    if constexpr (kSwz) {
        static_assert(kSeg == 8, "swizzled tiles are for 8-element segments");
        return sg * 8 + ((((r >> 2) ^ (sg >> 2)) & 1) << 2) + (r & 3);
    } else {
        return sg * seg_pad(kSeg) + r;
    }
}

template <typename T, int TC, int kSeg, bool kSwz = false>
__device__ __forceinline__ void stage_tile(float* __restrict__ dst, const T* __restrict__ src, int64_t dstate_stride,
                                           int dstate, int t0, int L, bool vec, int tid, int nthreads) {
    constexpr int ROWP = (TC / kSeg) * tile_seg_pitch<kSeg, kSwz>();
    if (vec) {
        if constexpr (sizeof(T) == 4) {
            constexpr int QPR = TC / 4;  // float4 per state row
            for (int s = tid; s < dstate * QPR; s += nthreads) {
                int n = s / QPR, q = s % QPR;
                int t = t0 + 4 * q;
                int rem = (L - t) * 4;
                int bytes = rem >= 16 ? 16 : (rem > 0 ? rem : 0);
                const T* g = src + n * dstate_stride + (bytes > 0 ? t : 0);
                cp_async16(dst + n * ROWP + tile_off<kSeg, kSwz>(4 * q), g, bytes);
            }
        } else {
            constexpr int QPR = TC / 8;  // 8-element (16 B) packets per state row
            for (int s = tid; s < dstate * QPR; s += nthreads) {
                int n = s / QPR, q = s % QPR;
                int t = t0 + 8 * q;
                float f[8];
                if (t + 8 <= L) {
                    uint4 v = __ldg(reinterpret_cast<const uint4*>(src + n * dstate_stride + t));
                    const T* e = reinterpret_cast<const T*>(&v);
#pragma unroll
                    for (int j = 0; j < 8; ++j) f[j] = Cvt<T>::to_f(e[j]);
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) f[j] = (t + j < L) ? Cvt<T>::to_f(src[n * dstate_stride + t + j]) : 0.f;
                }
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int tt = 8 * q + 4 * hh;     // timestep of this half within the chunk
                    *reinterpret_cast<float4*>(dst + n * ROWP + tile_off<kSeg, kSwz>(tt)) =
                        make_float4(f[4 * hh], f[4 * hh + 1], f[4 * hh + 2], f[4 * hh + 3]);
                }
            }
        }
    } else {
        for (int s = tid; s < dstate * TC; s += nthreads) {
            int n = s / TC, tt = s % TC;
            int t = t0 + tt;
            dst[n * ROWP + tile_off<kSeg, kSwz>(tt)] = (t < L) ? Cvt<T>::to_f(src[n * dstate_stride + t]) : 0.f;
        }
    }
}

__host__ __device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace fm
