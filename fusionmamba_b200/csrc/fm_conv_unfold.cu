// fm_conv_unfold.cu -- SS2D prologue for sm_100a: depthwise 3x3 conv + bias + SiLU + EfficientScan unfold in ONE pass.
//
// Replaces, on the inference path, the four full-tensor passes between in_proj and the scan in the reference's SS2D.forward /
// cross_selective_scan (models/cross.py:727-731 and :297):
//     x = xz[..., :D].permute(0, 3, 1, 2).contiguous()  ->  conv2d(x) (depthwise 3x3, padding 1)  ->  SiLU  ->  EfficientScan
//   src  xz  (batch, H, W, Cs) channels-last, the x half = channels [c_off, c_off + D)      (what in_proj writes)
//   dst  xs  (batch, 4, D, L), L = ceil(H/2)*ceil(W/2): the four stride-2 sub-grids, k in {0,2} row-major, {1,3} column-major
//        (index map of fm_permute.cu / models/cross.py:139-169); padded positions of odd sizes are written as 0.
// One CTA owns a TxT pixel tile (even origin, T = 8/16/32 by image size) of 16 channels: the halo tile is loaded with 16-byte
// vectors along channels (one 32-byte sector per pixel for 16-bit inputs) and kept in the input dtype, the conv runs from
// shared memory with fp32 accumulation and a 3-row register window,
// results are transposed through a second shared tile and stored with lanes along l (16 contiguous elements per sub-grid line).
// HBM roofline: s*D*H*W*batch bytes read + the same written; no tensor cores (9 MACs per element).
#include "fm_common.cuh"
#include "fm_launch.h"

namespace fm {

template <typename TI, typename TO, int T>
__global__ void __launch_bounds__(256)
conv_silu_unfold_kernel(const TI* __restrict__ xz, const float* __restrict__ wgt, const float* __restrict__ bias,
                        TO* __restrict__ xs, int D, int H, int W, int64_t Cs, int c_off) {
    constexpr int CH = 16, TI_ = T + 2;                   // T = tile side (8, 16 or 32: small images use small tiles), channels per CTA, halo side
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TI* s_in = reinterpret_cast<TI*>(smem_raw);                              // [TI_][TI_][CH]  (input dtype: 2-3 CTAs per SM)
    TO* s_out = reinterpret_cast<TO*>(s_in + TI_ * TI_ * CH);                // [CH][T][T + 2]   (padded rows)
    constexpr int OP = T + 2;
    constexpr int CHS = T * OP + (sizeof(TO) == 4 ? 1 : 2);                  // channel pitch: distinct banks for the 16 channel lanes

    const int tiles_w = (W + T - 1) / T;
    const int h0 = (blockIdx.x / tiles_w) * T, w0 = (blockIdx.x % tiles_w) * T;
    const int c0 = blockIdx.y * CH;
    const int b = blockIdx.z;
    const int tid = threadIdx.x;
    const int Hp = (H + 1) >> 1, Wp = (W + 1) >> 1;
    const int64_t L = static_cast<int64_t>(Hp) * Wp;

    // ---- load the halo tile: 16-byte vectors along channels (8 x 16-bit or 4 x fp32 per thread), zero outside the image --------
    {
        constexpr int VE = 16 / sizeof(TI);                // elements per 16-byte vector
        constexpr int VPP = CH / VE;                       // vectors per pixel
        const bool vec_ok = (c0 + CH <= D) && ((Cs * sizeof(TI)) % 16 == 0) && (((c_off + c0) * sizeof(TI)) % 16 == 0) &&
                            ((reinterpret_cast<uintptr_t>(xz) & 15u) == 0);
        if (vec_ok) {
            for (int e = tid; e < TI_ * TI_ * VPP; e += 256) {
                const int v = e % VPP, pix = e / VPP;
                const int hh = h0 - 1 + pix / TI_, ww = w0 - 1 + pix % TI_;
                uint4 val = make_uint4(0u, 0u, 0u, 0u);
                if (hh >= 0 && hh < H && ww >= 0 && ww < W)
                    val = __ldg(reinterpret_cast<const uint4*>(xz + ((static_cast<int64_t>(b) * H + hh) * W + ww) * Cs + c_off + c0 + v * VE));
                *reinterpret_cast<uint4*>(s_in + pix * CH + v * VE) = val;
            }
        } else {
            for (int e = tid; e < TI_ * TI_ * CH; e += 256) {
                const int c = e % CH, pix = e / CH;
                const int hh = h0 - 1 + pix / TI_, ww = w0 - 1 + pix % TI_;
                TI v = Cvt<TI>::from_f(0.f);
                if (hh >= 0 && hh < H && ww >= 0 && ww < W && c0 + c < D)
                    v = xz[((static_cast<int64_t>(b) * H + hh) * W + ww) * Cs + c_off + c0 + c];
                s_in[e] = v;
            }
        }
    }
    __syncthreads();

    // ---- depthwise 3x3 + bias + SiLU: thread = (channel, column), walks the 32 rows with a 3-row window --------
    {
        const int c = tid % CH;
        float w9[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) w9[i] = (c0 + c < D) ? __ldg(wgt + static_cast<int64_t>(c0 + c) * 9 + i) : 0.f;
        const float bv = (bias != nullptr && c0 + c < D) ? __ldg(bias + c0 + c) : 0.f;
        for (int col = tid / CH; col < T; col += 256 / CH) {
            float r0[3], r1[3], r2[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                r0[j] = Cvt<TI>::to_f(s_in[((0) * TI_ + col + j) * CH + c]);
                r1[j] = Cvt<TI>::to_f(s_in[((1) * TI_ + col + j) * CH + c]);
            }
            for (int row = 0; row < T; ++row) {
#pragma unroll
                for (int j = 0; j < 3; ++j) r2[j] = Cvt<TI>::to_f(s_in[((row + 2) * TI_ + col + j) * CH + c]);
                float acc = bv;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    acc = fmaf(w9[j], r0[j], acc);
                    acc = fmaf(w9[3 + j], r1[j], acc);
                    acc = fmaf(w9[6 + j], r2[j], acc);
                }
                const float y = acc * sigmoid_f(acc);
                const bool inside = (h0 + row < H) && (w0 + col < W);
                s_out[c * CHS + row * OP + col] = Cvt<TO>::from_f(inside ? y : 0.f);    // EfficientScan zero-pads odd sizes
#pragma unroll
                for (int j = 0; j < 3; ++j) { r0[j] = r1[j]; r1[j] = r2[j]; }
            }
        }
    }
    __syncthreads();

    // ---- unfold store: 4 sub-grids x CH channels x T/2 lines of T/2 contiguous l; consecutive threads walk a line ----------
    constexpr int HT = T / 2;
    TO* xsb = xs + static_cast<int64_t>(b) * 4 * D * L;
    for (int e = tid; e < 4 * CH * HT * HT; e += 256) {
        const int ln = e % HT, line = (e / HT) % HT, c = (e / (HT * HT)) % CH, k = e / (HT * HT * CH);
        if (c0 + c >= D) continue;
        // k & 1: row parity (h = 2i + (k&1)); k >> 1: column parity (w = 2j + (k>>1)); odd k is stored column-major
        int row, col;
        int64_t l;
        if (k & 1) {   // column-major: line = jj (fixed j), threads along ii
            row = 2 * ln + 1; col = 2 * line + (k >> 1);
            const int i = (h0 >> 1) + ln, j = (w0 >> 1) + line;
            if (i >= Hp || j >= Wp) continue;
            l = static_cast<int64_t>(j) * Hp + i;
        } else {       // row-major: line = ii (fixed i), threads along jj
            row = 2 * line; col = 2 * ln + (k >> 1);
            const int i = (h0 >> 1) + line, j = (w0 >> 1) + ln;
            if (i >= Hp || j >= Wp) continue;
            l = static_cast<int64_t>(i) * Wp + j;
        }
        xsb[(static_cast<int64_t>(k) * D + c0 + c) * L + l] = s_out[c * CHS + row * OP + col];
    }
}

template <typename TI, typename TO, int T>
static cudaError_t launch_cu_TT(const FmConvUnfoldParams& p, cudaStream_t st) {
    constexpr int CH = 16;
    const size_t smem = sizeof(TI) * (T + 2) * (T + 2) * CH + sizeof(TO) * CH * (T * (T + 2) + 2);
    auto kern = conv_silu_unfold_kernel<TI, TO, T>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid(((p.h + T - 1) / T) * ((p.w + T - 1) / T), (p.dim + CH - 1) / CH, p.batch);
    kern<<<grid, 256, smem, st>>>(static_cast<const TI*>(p.src), static_cast<const float*>(p.weight),
                                  static_cast<const float*>(p.bias), static_cast<TO*>(p.dst), p.dim, p.h, p.w,
                                  p.src_channel_stride, p.src_channel_offset);
    count_launch();
    return cudaGetLastError();
}

template <typename TI, typename TO>
static cudaError_t launch_cu_T(const FmConvUnfoldParams& p, cudaStream_t st) {
    const int m = p.h > p.w ? p.h : p.w;          // tile side: the smallest of 8 / 16 / 32 that covers the image, else 32
    if (m <= 8) return launch_cu_TT<TI, TO, 8>(p, st);
    if (m <= 16) return launch_cu_TT<TI, TO, 16>(p, st);
    return launch_cu_TT<TI, TO, 32>(p, st);
}

cudaError_t launch_conv_unfold(const FmConvUnfoldParams& p, cudaStream_t st) {
    switch (p.dtype) {
        case FM_F32: return launch_cu_T<float, float>(p, st);
        case FM_F16: return launch_cu_T<__half, __half>(p, st);
        default: return launch_cu_T<__nv_bfloat16, __nv_bfloat16>(p, st);
    }
}

}  // namespace fm
