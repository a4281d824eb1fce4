// fm_dt_proj.cu -- the rank-R dt projection of the SS2D core for sm_100a.
//
// Replaces  dts = einsum("b k r l, k d r -> b k d l", dts, dt_projs_weight)   (models/cross.py:309-310)
// on the inference path.  The contraction length is dt_rank = ceil(d_model / 16) (6 / 12 at the two long-sequence stages):
// far too short for a tensor-core tile -- cuBLAS answers with an sm_80 fallback kernel that takes 30 us for 50 MB of
// output at stage 0.  This is a bandwidth-bound outer-product kernel instead: a lane keeps VE consecutive l of all R rank
// rows in registers (fp32), a warp walks a tile of channels with the weight rows broadcast from shared memory, and every
// result leaves as one 16-byte store.  fp32 accumulation, output rounded to the I/O dtype (as the GEMM does).
//   src    dts    (batch, K, R, L) strided view of x_dbl, last dim contiguous
//   weight W      (K, D, R) contiguous, I/O dtype (the autocast copy) or fp32
//   dst    delta  (batch, K, D, L) contiguous
// HBM roofline: s*B*K*D*L bytes written (+ the small dts read).
#include "fm_common.cuh"
#include "fm_launch.h"

namespace fm {

template <typename T, typename TW, int R>
__global__ void __launch_bounds__(128)
dt_proj_kernel(const T* __restrict__ dts, const TW* __restrict__ w, T* __restrict__ out, int K, int D, int L,
               int64_t s_b, int64_t s_k, int64_t s_r, int vec) {
    constexpr int VE = 16 / sizeof(T);
    constexpr int DT = 64;                                  // channels per CTA (16 per warp)
    __shared__ float sW[DT][R];
    const int bk = blockIdx.z, b = bk / K, k = bk % K;
    const int d0 = blockIdx.y * DT;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int l = (blockIdx.x * 32 + lane) * VE;
    for (int i = threadIdx.x; i < DT * R; i += 128) {
        const int d = d0 + i / R;
        sW[i / R][i % R] = d < D ? Cvt<TW>::to_f(w[(static_cast<int64_t>(k) * D + d) * R + i % R]) : 0.f;
    }
    float x[R][VE];
    const T* src = dts + b * s_b + k * s_k + l;
    if (l < L) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (vec && l + VE <= L) {
                const uint4 q = __ldg(reinterpret_cast<const uint4*>(src + r * s_r));
                const T* e = reinterpret_cast<const T*>(&q);
#pragma unroll
                for (int j = 0; j < VE; ++j) x[r][j] = Cvt<T>::to_f(e[j]);
            } else {
#pragma unroll
                for (int j = 0; j < VE; ++j) x[r][j] = (l + j < L) ? Cvt<T>::to_f(src[r * s_r + j]) : 0.f;
            }
        }
    }
    __syncthreads();
    if (l >= L) return;
    T* dst = out + (static_cast<int64_t>(bk) * D + d0) * L + l;
#pragma unroll 2
    for (int dd = warp; dd < DT; dd += 4) {
        if (d0 + dd >= D) break;
        float acc[VE];
#pragma unroll
        for (int j = 0; j < VE; ++j) acc[j] = 0.f;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float wv = sW[dd][r];
#pragma unroll
            for (int j = 0; j < VE; ++j) acc[j] = fmaf(wv, x[r][j], acc[j]);
        }
        T* o = dst + static_cast<int64_t>(dd) * L;
        if (vec && l + VE <= L) {
            uint4 q;
            T* e = reinterpret_cast<T*>(&q);
#pragma unroll
            for (int j = 0; j < VE; ++j) e[j] = Cvt<T>::from_f(acc[j]);
            *reinterpret_cast<uint4*>(o) = q;
        } else {
#pragma unroll
            for (int j = 0; j < VE; ++j)
                if (l + j < L) o[j] = Cvt<T>::from_f(acc[j]);
        }
    }
}

template <typename T, typename TW>
static cudaError_t launch_dt_T(const FmDtProjParams& p, cudaStream_t st) {
    constexpr int VE = 16 / (int)sizeof(T);
    const int vec = (p.seqlen % VE == 0) && aligned16(p.src) && aligned16(p.dst) && p.src_batch_stride % VE == 0 &&
                    p.src_group_stride % VE == 0 && p.src_rank_stride % VE == 0;
    dim3 grid((p.seqlen + 32 * VE - 1) / (32 * VE), (p.dim + 63) / 64, p.batch * p.n_groups);
#define FM_DT(r)                                                                                                              \
    case r:                                                                                                                   \
        dt_proj_kernel<T, TW, r><<<grid, 128, 0, st>>>(static_cast<const T*>(p.src), static_cast<const TW*>(p.weight),        \
                                                       static_cast<T*>(p.dst), p.n_groups, p.dim, p.seqlen, p.src_batch_stride, \
                                                       p.src_group_stride, p.src_rank_stride, vec);                           \
        break;
    switch (p.rank) {
        FM_DT(1) FM_DT(2) FM_DT(3) FM_DT(4) FM_DT(5) FM_DT(6) FM_DT(7) FM_DT(8) FM_DT(9) FM_DT(10) FM_DT(11) FM_DT(12)
        default: return cudaErrorInvalidConfiguration;
    }
#undef FM_DT
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_dt_proj(const FmDtProjParams& p, cudaStream_t st) {
    const bool w32 = p.weight_dtype == FM_F32;
    switch (p.dtype) {
        case FM_F32: return w32 ? launch_dt_T<float, float>(p, st) : cudaErrorInvalidConfiguration;
        case FM_F16: return w32 ? launch_dt_T<__half, float>(p, st) : launch_dt_T<__half, __half>(p, st);
        default: return w32 ? launch_dt_T<__nv_bfloat16, float>(p, st) : launch_dt_T<__nv_bfloat16, __nv_bfloat16>(p, st);
    }
}

}  // namespace fm
