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
    TO* s_out = reinterpret_cast<TO*>(s_in + TI_ * TI_ * CH);                // [CH][4 sub-grids][T/2 lines][T/2]: already unfolded
    constexpr int HT = T / 2;
    constexpr int PP = HT * HT;                                              // one sub-grid plane of the tile
    constexpr int CHS = 4 * PP + (sizeof(TO) == 4 ? 1 : 2);                  // channel pitch: an odd number of 32-bit words

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
            // asynchronous 16-byte copies (zero-filled outside the image): every thread has all its ~10 requests in flight
            // at once instead of one load -> store round trip per vector
            const TI* xzb = xz + static_cast<int64_t>(b) * H * W * Cs + c_off + c0;
            for (int e = tid; e < TI_ * TI_ * VPP; e += 256) {
                const int v = e % VPP, pix = e / VPP;
                const int hh = h0 - 1 + pix / TI_, ww = w0 - 1 + pix % TI_;
                const bool in = hh >= 0 && hh < H && ww >= 0 && ww < W;
                cp_async16(s_in + pix * CH + v * VE, in ? xzb + (static_cast<int64_t>(hh) * W + ww) * Cs + v * VE : xzb, in ? 16 : 0);
            }
            cp_async_commit();
            cp_async_wait<0>();
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

    // ---- depthwise 3x3 + bias + SiLU: thread = (channel, column), walks the rows with a 3-row window; the result goes to
    //      its unfolded place: sub-grid k = (row & 1) | (col & 1) << 1, row-major (line = row/2, ln = col/2) for even rows,
    //      column-major (line = col/2, ln = row/2) for odd rows.  The two columns a warp holds are col and col + 2 (same
    //      sub-grid, neighbouring elements) --------
    {
        const int c = tid % CH;
        float w9[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) w9[i] = (c0 + c < D) ? __ldg(wgt + static_cast<int64_t>(c0 + c) * 9 + i) : 0.f;
        const float bv = (bias != nullptr && c0 + c < D) ? __ldg(bias + c0 + c) : 0.f;
        constexpr int CPP = 256 / CH;                      // columns per pass
        const int g = tid / 32, hsel = (tid / CH) & 1;
        const int colp = 4 * (g >> 1) + (g & 1) + 2 * hsel;   // 0 .. CPP-1, a permutation
        for (int col = colp; col < T; col += CPP) {
            TO* oe = s_out + c * CHS + ((col & 1) << 1) * PP + (col >> 1);             // even rows: + (row/2) * HT
            TO* oo = s_out + c * CHS + (1 | ((col & 1) << 1)) * PP + (col >> 1) * HT;  // odd rows:  + (row/2)
            const bool col_in = w0 + col < W;
            float r0[3], r1[3], r2[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                r0[j] = Cvt<TI>::to_f(s_in[((0) * TI_ + col + j) * CH + c]);
                r1[j] = Cvt<TI>::to_f(s_in[((1) * TI_ + col + j) * CH + c]);
            }
#pragma unroll 6
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
                const bool inside = col_in && (h0 + row < H);
                const TO o = Cvt<TO>::from_f(inside ? y : 0.f);                         // EfficientScan zero-pads odd sizes
                if (row & 1) oo[row >> 1] = o; else oe[(row >> 1) * HT] = o;
#pragma unroll
                for (int j = 0; j < 3; ++j) { r0[j] = r1[j]; r1[j] = r2[j]; }
            }
        }
    }
    __syncthreads();

    // ---- unfold store: 4 sub-grids x CH channels x T/2 lines of T/2 contiguous l; one thread stores VW consecutive l with
    //      one 16-byte (fp32: two) vector when the destination line is whole and aligned, element-wise at ragged edges --------
    constexpr int VW = HT < 8 ? HT : 8;
    constexpr int VPL = HT / VW;                           // vectors per line
    TO* xsb = xs + static_cast<int64_t>(b) * 4 * D * L;
    const bool vec_st = (Hp % VW == 0) && (Wp % VW == 0) && ((reinterpret_cast<uintptr_t>(xs) & 15u) == 0);
    for (int e = tid; e < 4 * CH * HT * VPL; e += 256) {
        const int lv = e % VPL, line = (e / VPL) % HT, c = (e / (VPL * HT)) % CH, k = e / (VPL * HT * CH);
        if (c0 + c >= D) continue;
        const TO* sp = s_out + c * CHS + k * PP + line * HT + lv * VW;
        // k & 1: row parity (h = 2i + (k&1)); k >> 1: column parity (w = 2j + (k>>1)); odd k is stored column-major
        int64_t l;
        int lim;                                           // valid elements from l on this line
        if (k & 1) {   // column-major: line = jj (fixed j), elements along ii
            const int i = (h0 >> 1) + lv * VW, j = (w0 >> 1) + line;
            if (i >= Hp || j >= Wp) continue;
            l = static_cast<int64_t>(j) * Hp + i;
            lim = Hp - i;
        } else {       // row-major: line = ii (fixed i), elements along jj
            const int i = (h0 >> 1) + line, j = (w0 >> 1) + lv * VW;
            if (i >= Hp || j >= Wp) continue;
            l = static_cast<int64_t>(i) * Wp + j;
            lim = Wp - j;
        }
        TO* dp = xsb + (static_cast<int64_t>(k) * D + c0 + c) * L + l;
        if (vec_st && lim >= VW) {
            constexpr int NWD = VW * sizeof(TO) / 4;       // 32-bit words per thread (shared side is only 4-byte aligned)
            uint32_t wv[NWD];
#pragma unroll
            for (int q = 0; q < NWD; ++q) wv[q] = reinterpret_cast<const uint32_t*>(sp)[q];
            if constexpr (NWD % 4 == 0) {
#pragma unroll
                for (int q = 0; q < NWD / 4; ++q)
                    reinterpret_cast<uint4*>(dp)[q] = make_uint4(wv[4 * q], wv[4 * q + 1], wv[4 * q + 2], wv[4 * q + 3]);
            } else {
#pragma unroll
                for (int q = 0; q < NWD / 2; ++q) reinterpret_cast<uint2*>(dp)[q] = make_uint2(wv[2 * q], wv[2 * q + 1]);
            }
        } else {
            for (int q = 0; q < VW && q < lim; ++q) dp[q] = sp[q];
        }
    }
}

template <typename TI, typename TO, int T>
static cudaError_t launch_cu_TT(const FmConvUnfoldParams& p, cudaStream_t st) {
    constexpr int CH = 16;
    const size_t smem = sizeof(TI) * (T + 2) * (T + 2) * CH + sizeof(TO) * CH * (T * T + 2);
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
