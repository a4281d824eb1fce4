// fm_conv_unfold_bwd.cu -- backward of the SS2D prologue (fm_conv_unfold.cu) for sm_100a, one pass.
//
// Replaces, under autograd, the backward of the reference's   x.permute(0,3,1,2).contiguous() -> conv2d (depthwise 3x3, padding 1)
// -> SiLU -> EfficientScan   chain (models/cross.py:727-731, :297, :171-190): torch runs EfficientScan.backward, the SiLU backward,
// two cuDNN depthwise-conv backward kernels (data + weight) and the permute backward as separate full-tensor passes.
//   xz   (batch, H, W, Cs) channels-last, x half = channels [c_off, c_off + D)        the saved input of the forward
//   dxs  (batch, 4, D, L), L = ceil(H/2)*ceil(W/2)                                      gradient of the unfolded output
//   dx   (batch, H, W, Cd) channels-last, written at channels [d_off, d_off + D)       gradient w.r.t. the x half
//   dW (D, 9), dbias (D)  fp32, ACCUMULATED with atomics (the caller zeroes them)
// One CTA owns a 16x16 pixel tile (even origin) of 16 channels.  z = conv(x) + bias is recomputed on the tile plus a one-pixel
// halo (x is loaded with a two-pixel halo), dconv = dy * SiLU'(z) replaces dy in shared memory (kept in the unfolded order it
// was loaded in: lanes along l on the global side, channel pitch odd on the shared side), the interior pixels feed the nine
// weight-gradient sums and the bias sum from the conv's own register window, and dx is the correlation of dconv with the
// flipped taps, transposed through shared memory into 16-byte channels-last stores.
#include "fm_common.cuh"
#include "fm_launch.h"

namespace fm {

namespace cub_ {   // (conv-unfold backward; no relation to the CUB library)
constexpr int T = 16, CH = 16, XI = T + 4, DI = T + 2, HP = T / 2 + 1;   // tile, channels, x halo side, dconv halo side, sub-grid patch side
constexpr int PP = HP * HP;                    // one sub-grid patch of the dconv halo
constexpr int CHS = 4 * PP + 1;                // channel pitch of the unfolded tile: odd
}

// A halo pixel (hh, ww) -- coordinates inside the (T+2)^2 dconv region, origin (h0-1, w0-1) with h0, w0 even -- is image pixel
// (h0-1+hh, w0-1+ww): row parity r = (hh+1)&1, column parity s = (ww+1)&1, sub-grid k = r | s << 1, patch-local sub-grid
// coordinates (pi, pj) = (hh >> 1, ww >> 1) (image sub-grid coordinates (h0/2 - r + pi, w0/2 - s + pj)).  Inside a patch the
// elements keep the global order: row-major (line pi, element pj) for even k, column-major (line pj, element pi) for odd k.
template <typename TI>
__global__ void __launch_bounds__(256)
conv_silu_unfold_bwd_kernel(const TI* __restrict__ xz, const float* __restrict__ wgt, const float* __restrict__ bias,
                            const TI* __restrict__ dxs, TI* __restrict__ dx, float* __restrict__ dW, float* __restrict__ dbias,
                            int D, int H, int W, int64_t Cs, int c_off, int64_t Cd, int d_off) {
    using namespace cub_;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TI* s_x = reinterpret_cast<TI*>(smem_raw);                                   // [XI][XI][CH]; later dx [T][T][CH]
    float* s_d = reinterpret_cast<float*>(s_x + XI * XI * CH);                   // [CH][CHS] dy -> dconv; later the reduction tile

    const int tiles_w = (W + T - 1) / T;
    const int h0 = (blockIdx.x / tiles_w) * T, w0 = (blockIdx.x % tiles_w) * T;
    const int c0 = blockIdx.y * CH;
    const int b = blockIdx.z;
    const int tid = threadIdx.x;
    const int Hp = (H + 1) >> 1, Wp = (W + 1) >> 1;
    const int64_t L = static_cast<int64_t>(Hp) * Wp;

    // ---- x with a two-pixel halo: 16-byte vectors along channels, zero outside the image (= the conv's zero padding) ------------
    {
        constexpr int VE = 16 / sizeof(TI), VPP = CH / VE;
        const bool vec_ok = (c0 + CH <= D) && ((Cs * sizeof(TI)) % 16 == 0) && (((c_off + c0) * sizeof(TI)) % 16 == 0) &&
                            ((reinterpret_cast<uintptr_t>(xz) & 15u) == 0);
        if (vec_ok) {
            const TI* xzb = xz + static_cast<int64_t>(b) * H * W * Cs + c_off + c0;
            for (int e = tid; e < XI * XI * VPP; e += 256) {
                const int v = e % VPP, pix = e / VPP;
                const int hh = h0 - 2 + pix / XI, ww = w0 - 2 + pix % XI;
                const bool in = hh >= 0 && hh < H && ww >= 0 && ww < W;
                cp_async16(s_x + pix * CH + v * VE, in ? xzb + (static_cast<int64_t>(hh) * W + ww) * Cs + v * VE : xzb, in ? 16 : 0);
            }
            cp_async_commit();
        } else {
            for (int e = tid; e < XI * XI * CH; e += 256) {
                const int c = e % CH, pix = e / CH;
                const int hh = h0 - 2 + pix / XI, ww = w0 - 2 + pix % XI;
                TI v = Cvt<TI>::from_f(0.f);
                if (hh >= 0 && hh < H && ww >= 0 && ww < W && c0 + c < D)
                    v = xz[((static_cast<int64_t>(b) * H + hh) * W + ww) * Cs + c_off + c0 + c];
                s_x[e] = v;
            }
        }
    }
    // ---- dy with a one-pixel halo, read in its unfolded order (lanes along l), 0 outside the image ---------------------------------
    {
        const TI* dxb = dxs + static_cast<int64_t>(b) * 4 * D * L;
        const int i0 = (h0 >> 1), j0 = (w0 >> 1);
        for (int e = tid; e < 4 * CH * PP; e += 256) {
            const int el = e % HP, line = (e / HP) % HP, c = (e / PP) % CH, k = e / (PP * CH);
            const int r = k & 1, s = k >> 1;
            // odd parity starts one sub-grid step before the tile (image row h0 - 1), even parity at the tile origin
            const int pi = (r ? el : line), pj = (r ? line : el);              // patch-local (row, column) sub-grid indices
            const int i = i0 - r + pi, j = j0 - s + pj;                          // sub-grid coordinates in the image
            const int h = 2 * i + r, w = 2 * j + s;
            // the halo region is rows h0-1 .. h0+T: parity-1 rows h0-1 .. h0+T-1 (pi = 0 .. T/2), parity-0 rows h0 .. h0+T
            float v = 0.f;
            if (i >= 0 && j >= 0 && h < H && w < W && c0 + c < D) {
                const int64_t l = r ? static_cast<int64_t>(j) * Hp + i : static_cast<int64_t>(i) * Wp + j;
                v = Cvt<TI>::to_f(dxb[(static_cast<int64_t>(k) * D + c0 + c) * L + l]);
            }
            s_d[c * CHS + k * PP + line * HP + el] = v;
        }
    }
    cp_async_wait<0>();
    __syncthreads();

    const int c = tid % CH, colp = tid / CH;               // thread = (channel, column)
    float w9[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) w9[i] = (c0 + c < D) ? __ldg(wgt + static_cast<int64_t>(c0 + c) * 9 + i) : 0.f;
    const float bv = (bias != nullptr && c0 + c < D) ? __ldg(bias + c0 + c) : 0.f;

    // ---- z = conv(x) + bias on the halo region, dconv = dy * SiLU'(z) in place; interior pixels feed dW / dbias -------------------
    float gw[9], gb = 0.f;
#pragma unroll
    for (int i = 0; i < 9; ++i) gw[i] = 0.f;
    float* sdc = s_d + c * CHS;
    for (int col = colp; col < DI; col += 256 / CH) {       // halo column: image column w0 - 1 + col
        const bool col_int = col >= 1 && col <= T;
        const int s = (col + 1) & 1, pj = col >> 1;
        float r0[3], r1[3], r2[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            r0[j] = Cvt<TI>::to_f(s_x[((0) * XI + col + j) * CH + c]);
            r1[j] = Cvt<TI>::to_f(s_x[((1) * XI + col + j) * CH + c]);
        }
#pragma unroll 2
        for (int row = 0; row < DI; ++row) {                // halo row: image row h0 - 1 + row
#pragma unroll
            for (int j = 0; j < 3; ++j) r2[j] = Cvt<TI>::to_f(s_x[((row + 2) * XI + col + j) * CH + c]);
            float z = bv;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                z = fmaf(w9[j], r0[j], z);
                z = fmaf(w9[3 + j], r1[j], z);
                z = fmaf(w9[6 + j], r2[j], z);
            }
            const float sg = sigmoid_f(z);
            const float dsilu = sg * fmaf(z, 1.f - sg, 1.f);
            const int r = (row + 1) & 1, pi = row >> 1;
            float* slot = sdc + (r | (s << 1)) * PP + (r ? pj * HP + pi : pi * HP + pj);
            const float dc = *slot * dsilu;                  // dy is 0 outside the image
            *slot = dc;
            if (col_int && row >= 1 && row <= T) {
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    gw[j] = fmaf(dc, r0[j], gw[j]);
                    gw[3 + j] = fmaf(dc, r1[j], gw[3 + j]);
                    gw[6 + j] = fmaf(dc, r2[j], gw[6 + j]);
                }
                gb += dc;
            }
#pragma unroll
            for (int j = 0; j < 3; ++j) { r0[j] = r1[j]; r1[j] = r2[j]; }
        }
    }
    __syncthreads();

    // ---- dx = correlation of dconv with the flipped taps: dx[q] = sum_{dr,dc} w[dr][dc] dconv[q - (dr-1, dc-1)] ------------------------
    {
        TI* s_o = s_x;                                       // [T][T][CH], the x tile is dead
        const int col = colp;                                // tile column 0 .. 15: halo columns col .. col + 2
        // halo column col + 2 - dc holds dconv for tap column dc
        int sl[3], pjv[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) { sl[j] = (col + j + 1) & 1; pjv[j] = (col + j) >> 1; }
        auto ldd = [&](int row, int j) {                     // dconv at halo (row, col + j)
            const int r = (row + 1) & 1, pi = row >> 1;
            return sdc[(r | (sl[j] << 1)) * PP + (r ? pjv[j] * HP + pi : pi * HP + pjv[j])];
        };
        float d0[3], d1[3], d2[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) { d0[j] = ldd(0, j); d1[j] = ldd(1, j); }
#pragma unroll 2
        for (int row = 0; row < T; ++row) {                  // tile row: halo rows row .. row + 2
#pragma unroll
            for (int j = 0; j < 3; ++j) d2[j] = ldd(row + 2, j);
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < 3; ++j) {                    // halo (row + a, col + j) = q - (dr-1, dc-1) with dr = 2 - a, dc = 2 - j
                acc = fmaf(w9[6 + (2 - j)], d0[j], acc);
                acc = fmaf(w9[3 + (2 - j)], d1[j], acc);
                acc = fmaf(w9[0 + (2 - j)], d2[j], acc);
            }
            s_o[(row * T + col) * CH + c] = Cvt<TI>::from_f(acc);
#pragma unroll
            for (int j = 0; j < 3; ++j) { d0[j] = d1[j]; d1[j] = d2[j]; }
        }
    }
    __syncthreads();

    // ---- dx store: 16-byte vectors along channels; weight / bias gradients: per-channel sums over the CTA's 16 columns --------------
    {
        const TI* s_o = s_x;
        constexpr int VE = 16 / sizeof(TI), VPP = CH / VE;
        const bool vec_ok = (c0 + CH <= D) && ((Cd * sizeof(TI)) % 16 == 0) && (((d_off + c0) * sizeof(TI)) % 16 == 0) &&
                            ((reinterpret_cast<uintptr_t>(dx) & 15u) == 0);
        TI* dxb = dx + static_cast<int64_t>(b) * H * W * Cd + d_off + c0;
        if (vec_ok) {
            for (int e = tid; e < T * T * VPP; e += 256) {
                const int v = e % VPP, pix = e / VPP;
                const int hh = h0 + pix / T, ww = w0 + pix % T;
                if (hh < H && ww < W)
                    *reinterpret_cast<uint4*>(dxb + (static_cast<int64_t>(hh) * W + ww) * Cd + v * VE) =
                        *reinterpret_cast<const uint4*>(s_o + pix * CH + v * VE);
            }
        } else {
            for (int e = tid; e < T * T * CH; e += 256) {
                const int cc = e % CH, pix = e / CH;
                const int hh = h0 + pix / T, ww = w0 + pix % T;
                if (hh < H && ww < W && c0 + cc < D) dxb[(static_cast<int64_t>(hh) * W + ww) * Cd + cc] = s_o[e];
            }
        }
        float* s_r = s_d;                                    // [10][16 columns][CH]
#pragma unroll
        for (int i = 0; i < 9; ++i) s_r[(i * 16 + colp) * CH + c] = gw[i];
        s_r[(9 * 16 + colp) * CH + c] = gb;
    }
    __syncthreads();
    if (tid < 10 * CH) {
        const int i = tid / CH, cc = tid % CH;
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) sum += s_d[(i * 16 + k) * CH + cc];
        if (c0 + cc < D) {
            if (i < 9) atomicAdd(dW + static_cast<int64_t>(c0 + cc) * 9 + i, sum);
            else if (dbias != nullptr) atomicAdd(dbias + c0 + cc, sum);
        }
    }
}

template <typename TI>
static cudaError_t launch_cub_T(const FmConvUnfoldBwdParams& p, cudaStream_t st) {
    using namespace cub_;
    const size_t smem = sizeof(TI) * XI * XI * CH + sizeof(float) * CH * CHS;
    auto kern = conv_silu_unfold_bwd_kernel<TI>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid(((p.h + T - 1) / T) * ((p.w + T - 1) / T), (p.dim + CH - 1) / CH, p.batch);
    kern<<<grid, 256, smem, st>>>(static_cast<const TI*>(p.src), static_cast<const float*>(p.weight), static_cast<const float*>(p.bias),
                                  static_cast<const TI*>(p.dxs), static_cast<TI*>(p.dsrc), static_cast<float*>(p.dweight),
                                  static_cast<float*>(p.dbias), p.dim, p.h, p.w, p.src_channel_stride, p.src_channel_offset,
                                  p.dsrc_channel_stride, p.dsrc_channel_offset);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_conv_unfold_bwd(const FmConvUnfoldBwdParams& p, cudaStream_t st) {
    switch (p.dtype) {
        case FM_F32: return launch_cub_T<float>(p, st);
        case FM_F16: return launch_cub_T<__half>(p, st);
        default: return launch_cub_T<__nv_bfloat16>(p, st);
    }
}

}  // namespace fm
