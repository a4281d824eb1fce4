// fm_norm.cu -- SS2D epilogue for sm_100a: transpose + LayerNorm(D) + cast in ONE pass.
//
// Replaces, on the inference path, the three full-tensor passes that follow the merge in the reference's SS2D core
// (models/cross.py:334-335 and :337):   y.transpose(1, 2).contiguous()  ->  out_norm(y)  ->  y.to(x.dtype)
//   src  y   (batch, D, P) fp32, P = H*W positions contiguous   (what the scan's fused merge store writes)
//   dst  out (batch, P, D) in the output dtype, normalised over D with nn.LayerNorm semantics (biased variance, eps inside the
//        sqrt, affine weight / bias in fp32), optionally multiplied by SiLU(gate) read from the channels-last in_proj output
//        (the y * z of SS2D.forward, models/cross.py:728-729, 740).
// One CTA owns 32 consecutive positions of one batch item and all D channels:
//   pass 1  per-position sum and sum of squares, shifted by the position's first channel (no cancellation), read with
//           lanes along P (128-byte coalesced rows of y), reduced over the CTA's warps through shared memory;
//   pass 2  32x32 (channel x position) tiles re-read (L2-resident: the CTA touched them microseconds ago), normalised,
//           transposed through a padded shared tile and stored with lanes along D (coalesced rows of out).
// HBM roofline: 4*D*P*batch bytes read + s*D*P*batch written; no tensor cores (no GEMM shape).
#include "fm_common.cuh"
#include "fm_launch.h"

namespace fm {

template <typename TO, int NW>
__global__ void __launch_bounds__(NW * 32)
merge_norm_kernel(const float* __restrict__ y, const float* __restrict__ w, const float* __restrict__ bsh, TO* __restrict__ out,
                  int D, int P, float eps, const TO* __restrict__ gate, int64_t gcs, int goff) {
    constexpr int TP = 32;                                  // positions per CTA
    __shared__ float s_sum[NW][TP], s_sq[NW][TP];
    __shared__ float s_mean[TP], s_rstd[TP];
    __shared__ float tile[NW][32][33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const int p0 = blockIdx.x * TP;
    const int p = p0 + lane;
    const bool pok = p < P;
    const float* yb = y + static_cast<int64_t>(b) * D * P;
    const float shift = pok ? __ldg(yb + p) : 0.f;          // channel 0 of this position

    float s = 0.f, q = 0.f;
    for (int d = warp; d < D; d += NW) {
        const float v = pok ? __ldg(yb + static_cast<int64_t>(d) * P + p) - shift : 0.f;
        s += v;
        q = fmaf(v, v, q);
    }
    s_sum[warp][lane] = s;
    s_sq[warp][lane] = q;
    __syncthreads();
    if (warp == 0) {
        float ts = 0.f, tq = 0.f;
#pragma unroll
        for (int k = 0; k < NW; ++k) { ts += s_sum[k][lane]; tq += s_sq[k][lane]; }
        const float m = ts / D;
        const float var = fmaxf(tq / D - m * m, 0.f);
        s_mean[lane] = m + shift;
        s_rstd[lane] = rsqrtf(var + eps);
    }
    __syncthreads();

    TO* ob = out + (static_cast<int64_t>(b) * P + p0) * D;
    for (int d0 = warp * 32; d0 < D; d0 += NW * 32) {
        // load a 32(d) x 32(p) tile with lanes along p, normalise with the position's statistics
        const float mean = s_mean[lane], rstd = s_rstd[lane];
#pragma unroll 8
        for (int r = 0; r < 32; ++r) {
            const int d = d0 + r;
            float v = 0.f;
            if (d < D && pok) v = (__ldg(yb + static_cast<int64_t>(d) * P + p) - mean) * rstd;
            tile[warp][r][lane] = v;
        }
        __syncwarp();
        // store with lanes along d: out[b, p0 + r, d0 + lane]
        const int d = d0 + lane;
        if (d < D) {
            const float wd = w ? __ldg(w + d) : 1.f, bd = bsh ? __ldg(bsh + d) : 0.f;
#pragma unroll 8
            for (int r = 0; r < 32; ++r)
                if (p0 + r < P) {
                    float v = fmaf(tile[warp][lane][r], wd, bd);
                    if (gate != nullptr) {
                        const float g = Cvt<TO>::to_f(gate[(static_cast<int64_t>(b) * P + p0 + r) * gcs + goff + d]);
                        // round both factors to the output dtype first: same values as the reference's separate LayerNorm / SiLU ops
                        v = Cvt<TO>::to_f(Cvt<TO>::from_f(v)) * Cvt<TO>::to_f(Cvt<TO>::from_f(g * sigmoid_f(g)));
                    }
                    ob[static_cast<int64_t>(r) * D + d] = Cvt<TO>::from_f(v);
                }
        }
        __syncwarp();
    }
}

// Channels-last source (batch, positions, dim) -- what the scan's FM_MAP_EFFICIENT_V2_CL store writes: plain row LayerNorm,
// one warp per position, the row held in registers (dim <= 2048), lanes along channels for loads and stores.
template <typename TO, int NW, int MAXI>
__global__ void __launch_bounds__(NW * 32)
row_norm_kernel(const float* __restrict__ y, const float* __restrict__ w, const float* __restrict__ bsh, TO* __restrict__ out,
                int D, int64_t rows, float eps, const TO* __restrict__ gate, int64_t gcs, int goff) {
    const int lane = threadIdx.x & 31;   // MAXI >= ceil(D / 32): register slots per lane (2 ... 64)
    const int64_t row = static_cast<int64_t>(blockIdx.x) * NW + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* src = y + row * D;
    const int ni = (D + 31) >> 5;
    float v[MAXI];
    const float shift = __ldg(src);
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int i = 0; i < MAXI; ++i) {
        if (i < ni) {
            const int d = lane + 32 * i;
            const float x = d < D ? __ldg(src + d) - shift : 0.f;
            v[i] = x;
            s += x;
            q = fmaf(x, x, q);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    const float m = s / D;
    const float rstd = rsqrtf(fmaxf(q / D - m * m, 0.f) + eps);
    TO* dst = out + row * D;
    const TO* g = gate ? gate + row * gcs + goff : nullptr;
#pragma unroll
    for (int i = 0; i < MAXI; ++i) {
        if (i < ni) {
            const int d = lane + 32 * i;
            if (d < D) {
                float r = (v[i] - m) * rstd;
                r = fmaf(r, w ? __ldg(w + d) : 1.f, bsh ? __ldg(bsh + d) : 0.f);
                if (g != nullptr) {
                    const float gv = Cvt<TO>::to_f(g[d]);
                    r = Cvt<TO>::to_f(Cvt<TO>::from_f(r)) * Cvt<TO>::to_f(Cvt<TO>::from_f(gv * sigmoid_f(gv)));
                }
                dst[d] = Cvt<TO>::from_f(r);
            }
        }
    }
}

// Vectorised row LayerNorm for dim % 4 == 0 (every SS2D in the model): LP lanes share one position, each lane owns NV
// 4-channel vectors (128-bit loads of y, 64/128-bit loads of the gate and stores of out), 32 / LP positions per warp pass,
// warps walk the positions grid-stride with the affine weights parked in registers (NV <= 4).  Same arithmetic as
// row_norm_kernel (shifted one-pass moments, biased variance, double rounding of the gated product) at roughly a quarter
// of its instructions per element: the scalar kernel is issue-bound at 2 TB/s (profiles/r01_ss2d_block_ncu_stage0.txt).
template <typename TO> struct Vec4 { using type = uint2; };          // 4 x 16-bit
template <> struct Vec4<float> { using type = float4; };

template <typename TO>
__device__ __forceinline__ void unpack4(const typename Vec4<TO>::type& r, float (&f)[4]) {
    if constexpr (sizeof(TO) == 4) {
        f[0] = r.x; f[1] = r.y; f[2] = r.z; f[3] = r.w;
    } else {
        const TO* e = reinterpret_cast<const TO*>(&r);
#pragma unroll
        for (int i = 0; i < 4; ++i) f[i] = Cvt<TO>::to_f(e[i]);
    }
}

template <typename TO, int NW, int LP, int NV, bool kRegW>
__global__ void __launch_bounds__(NW * 32)
row_norm_vec_kernel(const float* __restrict__ y, const float* __restrict__ w, const float* __restrict__ bsh, TO* __restrict__ out,
                    int D, int64_t rows, float eps, const TO* __restrict__ gate, int64_t gcs, int goff) {
    using V4 = typename Vec4<TO>::type;
    constexpr int PW = 32 / LP;                              // positions per warp pass
    const int lane = threadIdx.x & 31;
    const int sub = lane % LP, pw = lane / LP;
    const int V = D >> 2;                                    // vectors per position
    const float inv_d = 1.f / D;
    float4 wr[kRegW ? NV : 1], br[kRegW ? NV : 1];
    if constexpr (kRegW) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int j = sub + LP * i;
            wr[i] = (w != nullptr && j < V) ? __ldg(reinterpret_cast<const float4*>(w) + j) : make_float4(1.f, 1.f, 1.f, 1.f);
            br[i] = (bsh != nullptr && j < V) ? __ldg(reinterpret_cast<const float4*>(bsh) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    const int64_t stride = static_cast<int64_t>(gridDim.x) * NW * PW;
    for (int64_t row0 = (static_cast<int64_t>(blockIdx.x) * NW + (threadIdx.x >> 5)) * PW; row0 < rows; row0 += stride) {
        const int64_t row = row0 + pw;
        const bool rok = row < rows;
        const float4* src = reinterpret_cast<const float4*>(y + (rok ? row : 0) * D);
        float4 v[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int j = sub + LP * i;
            v[i] = (rok && j < V) ? __ldg(src + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        V4 gr[NV];
        if (gate != nullptr) {
            const V4* g = reinterpret_cast<const V4*>(gate + (rok ? row : 0) * gcs + goff);
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int j = sub + LP * i;
                if (rok && j < V) gr[i] = __ldg(g + j);
            }
        }
        const float shift = __shfl_sync(0xffffffffu, v[0].x, pw * LP);          // channel 0 of this position
        float s = 0.f, q = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int j = sub + LP * i;
            if (j < V) {
                v[i].x -= shift; v[i].y -= shift; v[i].z -= shift; v[i].w -= shift;
                s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
                q = fmaf(v[i].x, v[i].x, q); q = fmaf(v[i].y, v[i].y, q); q = fmaf(v[i].z, v[i].z, q); q = fmaf(v[i].w, v[i].w, q);
            }
        }
#pragma unroll
        for (int o = LP / 2; o > 0; o >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            q += __shfl_xor_sync(0xffffffffu, q, o);
        }
        const float m = s * inv_d;
        const float rstd = rsqrtf(fmaxf(q * inv_d - m * m, 0.f) + eps);
        const float nm = -m * rstd;
        V4* dst = reinterpret_cast<V4*>(out + (rok ? row : 0) * D);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int j = sub + LP * i;
            if (rok && j < V) {
                float4 wv, bv;
                if constexpr (kRegW) { wv = wr[i]; bv = br[i]; }
                else {
                    wv = w != nullptr ? __ldg(reinterpret_cast<const float4*>(w) + j) : make_float4(1.f, 1.f, 1.f, 1.f);
                    bv = bsh != nullptr ? __ldg(reinterpret_cast<const float4*>(bsh) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                float r[4] = {fmaf(fmaf(v[i].x, rstd, nm), wv.x, bv.x), fmaf(fmaf(v[i].y, rstd, nm), wv.y, bv.y),
                              fmaf(fmaf(v[i].z, rstd, nm), wv.z, bv.z), fmaf(fmaf(v[i].w, rstd, nm), wv.w, bv.w)};
                if (gate != nullptr) {
                    float gv[4];
                    unpack4<TO>(gr[i], gv);
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        r[e] = Cvt<TO>::to_f(Cvt<TO>::from_f(r[e])) * Cvt<TO>::to_f(Cvt<TO>::from_f(gv[e] * sigmoid_f(gv[e])));
                }
                V4 o;
                if constexpr (sizeof(TO) == 4) {
                    o = make_float4(r[0], r[1], r[2], r[3]);
                } else {
                    TO* e = reinterpret_cast<TO*>(&o);
#pragma unroll
                    for (int k = 0; k < 4; ++k) e[k] = Cvt<TO>::from_f(r[k]);
                }
                dst[j] = o;
            }
        }
    }
}

template <typename TO, int NW, int LP, int NV>
static cudaError_t launch_row_norm_vec(const FmNormParams& p, cudaStream_t st, int64_t rows) {
    constexpr int PW = 32 / LP;
    constexpr bool kRegW = NV <= 4;
    const int64_t passes = (rows + NW * PW - 1) / (NW * PW);
    const int64_t cap = 148 * 8 * 4;                         // a few resident waves; the warps loop over the rest
    const unsigned grid = (unsigned)(passes < cap ? passes : cap);
    row_norm_vec_kernel<TO, NW, LP, NV, kRegW><<<grid, NW * 32, 0, st>>>(
        static_cast<const float*>(p.src), static_cast<const float*>(p.weight), static_cast<const float*>(p.bias),
        static_cast<TO*>(p.dst), p.dim, rows, p.eps, static_cast<const TO*>(p.gate), p.gate_channel_stride, p.gate_channel_offset);
    return cudaGetLastError();
}

template <typename TO>
static bool row_norm_vec_ok(const FmNormParams& p) {
    if (p.dim % 4 != 0 || p.dim > 2048) return false;
    auto al = [](const void* q, size_t a) { return (reinterpret_cast<uintptr_t>(q) % a) == 0; };
    if (!al(p.src, 16) || !al(p.dst, 4 * sizeof(TO))) return false;
    if ((p.weight && !al(p.weight, 16)) || (p.bias && !al(p.bias, 16))) return false;
    if (p.gate && (!al(p.gate, 4 * sizeof(TO)) || p.gate_channel_stride % 4 != 0 || p.gate_channel_offset % 4 != 0)) return false;
    return true;
}

template <typename TO>
static cudaError_t launch_norm_T(const FmNormParams& p, cudaStream_t st) {
    constexpr int NW = 8;
    if (p.src_channels_last) {
        if (p.dim > 2048) return cudaErrorInvalidConfiguration;
        const int64_t rows = static_cast<int64_t>(p.batch) * p.positions;
        if (row_norm_vec_ok<TO>(p) && env_int("FM_NORM_VEC", 1)) {
            const int V = p.dim / 4;
            cudaError_t e;
            if (V <= 8) e = launch_row_norm_vec<TO, NW, 8, 1>(p, st, rows);
            else if (V <= 16) e = launch_row_norm_vec<TO, NW, 8, 2>(p, st, rows);
            else if (V <= 32) e = launch_row_norm_vec<TO, NW, 16, 2>(p, st, rows);
            else if (V <= 48) e = launch_row_norm_vec<TO, NW, 16, 3>(p, st, rows);
            else if (V <= 64) e = launch_row_norm_vec<TO, NW, 16, 4>(p, st, rows);
            else if (V <= 96) e = launch_row_norm_vec<TO, NW, 32, 3>(p, st, rows);
            else if (V <= 128) e = launch_row_norm_vec<TO, NW, 32, 4>(p, st, rows);
            else if (V <= 192) e = launch_row_norm_vec<TO, NW, 32, 6>(p, st, rows);
            else if (V <= 256) e = launch_row_norm_vec<TO, NW, 32, 8>(p, st, rows);
            else if (V <= 384) e = launch_row_norm_vec<TO, NW, 32, 12>(p, st, rows);
            else e = launch_row_norm_vec<TO, NW, 32, 16>(p, st, rows);
            count_launch();
            return e;
        }
        const int ni = (p.dim + 31) / 32;
#define FM_ROWNORM(mi)                                                                                                        \
    row_norm_kernel<TO, NW, mi><<<(unsigned)((rows + NW - 1) / NW), NW * 32, 0, st>>>(                                        \
        static_cast<const float*>(p.src), static_cast<const float*>(p.weight), static_cast<const float*>(p.bias),            \
        static_cast<TO*>(p.dst), p.dim, rows, p.eps, static_cast<const TO*>(p.gate), p.gate_channel_stride, p.gate_channel_offset)
        if (ni <= 2) FM_ROWNORM(2);
        else if (ni <= 4) FM_ROWNORM(4);
        else if (ni <= 8) FM_ROWNORM(8);
        else if (ni <= 16) FM_ROWNORM(16);
        else if (ni <= 32) FM_ROWNORM(32);
        else FM_ROWNORM(64);
#undef FM_ROWNORM
        count_launch();
        return cudaGetLastError();
    }
    dim3 grid((p.positions + 31) / 32, p.batch);
    merge_norm_kernel<TO, NW><<<grid, NW * 32, 0, st>>>(static_cast<const float*>(p.src), static_cast<const float*>(p.weight),
                                                         static_cast<const float*>(p.bias), static_cast<TO*>(p.dst), p.dim,
                                                         p.positions, p.eps, static_cast<const TO*>(p.gate), p.gate_channel_stride,
                                                         p.gate_channel_offset);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_merge_norm(const FmNormParams& p, cudaStream_t st) {
    switch (p.out_dtype) {
        case FM_F32: return launch_norm_T<float>(p, st);
        case FM_F16: return launch_norm_T<__half>(p, st);
        default: return launch_norm_T<__nv_bfloat16>(p, st);
    }
}

}  // namespace fm
