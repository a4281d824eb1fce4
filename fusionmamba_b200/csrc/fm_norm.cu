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

template <typename TO>
static cudaError_t launch_norm_T(const FmNormParams& p, cudaStream_t st) {
    constexpr int NW = 8;
    if (p.src_channels_last) {
        if (p.dim > 2048) return cudaErrorInvalidConfiguration;
        const int64_t rows = static_cast<int64_t>(p.batch) * p.positions;
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
