// fm_norm_bwd.cu -- LayerNorm(D) backward over channels-last rows for sm_100a (training side of the SS2D epilogue).
//
// The reference normalises the merged scan output with nn.LayerNorm (out_norm, models/cross.py:334-335) and every VSSBlock_new
// wraps the SS2D path in three more (models/cross.py:1334, 1352, 748).  In the training step (BASELINE configs[3]) their
// backward is two ATen kernels per call -- the input gradient and a gamma/beta column reduction that alone takes as long as
// the forward (profiles/r02_train_breakdown_patched.json: 19.6 + 9.6 + 18.8 ms of a 384 ms step).  This is the one-pass
// replacement, same row layout as the forward kernel (fm_norm.cu: LP lanes share a row, NV 128-bit vectors per lane):
//   per row   mean / rstd are RECOMPUTED from x (x is read anyway: no saved statistics, nothing extra kept from the forward),
//             gw = dy * w,  c1 = mean(gw),  c2 = mean(gw * xhat),  dx = rstd * (gw - c1 - xhat * c2)     (one pass over x, dy)
//   columns   dgamma += dy * xhat, dbeta += dy accumulate in registers over the rows a lane walks, are folded over the CTA's
//             row groups through shared memory and leave as ONE partial row per CTA; a second tiny kernel sums the partials
//             (deterministic: no atomics).
// HBM: 8*D bytes read + 4*D written per row -- the roofline of the op; no tensor cores (no GEMM shape).
#include "fm_common.cuh"
#include "fm_launch.h"

namespace fm {

template <int NW, int LP, int NV>
__global__ void __launch_bounds__(NW * 32)
layer_norm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ w,
                      float* __restrict__ dx, float* __restrict__ partial, int D, int64_t rows, float eps) {
    constexpr int PW = 32 / LP;                              // rows per warp pass
    extern __shared__ __align__(16) float s_red[];           // [NW * PW][2][D]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane % LP, pw = lane / LP;
    const int V = D >> 2;
    const float inv_d = 1.f / D;
    float4 wr[NV], dg[NV], db[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int j = sub + LP * i;
        wr[i] = (w != nullptr && j < V) ? __ldg(reinterpret_cast<const float4*>(w) + j) : make_float4(1.f, 1.f, 1.f, 1.f);
        dg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const int64_t stride = static_cast<int64_t>(gridDim.x) * NW * PW;
    for (int64_t row0 = (static_cast<int64_t>(blockIdx.x) * NW + warp) * PW; row0 < rows; row0 += stride) {
        const int64_t row = row0 + pw;
        const bool rok = row < rows;
        const float4* xs = reinterpret_cast<const float4*>(x + (rok ? row : 0) * D);
        const float4* gs = reinterpret_cast<const float4*>(dy + (rok ? row : 0) * D);
        float4 v[NV], g[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int j = sub + LP * i;
            const bool ok = rok && j < V;
            v[i] = ok ? __ldg(xs + j) : make_float4(0.f, 0.f, 0.f, 0.f);
            g[i] = ok ? __ldg(gs + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // statistics of the row, shifted by its first channel like the forward (no cancellation)
        const float shift = __shfl_sync(0xffffffffu, v[0].x, pw * LP);
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
        // xhat in place of v; the two row means of the gradient
        float c1 = 0.f, c2 = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int j = sub + LP * i;
            if (j < V) {
                v[i].x = fmaf(v[i].x, rstd, nm); v[i].y = fmaf(v[i].y, rstd, nm);
                v[i].z = fmaf(v[i].z, rstd, nm); v[i].w = fmaf(v[i].w, rstd, nm);
                const float gx = g[i].x * wr[i].x, gy = g[i].y * wr[i].y, gz = g[i].z * wr[i].z, gw = g[i].w * wr[i].w;
                c1 += (gx + gy) + (gz + gw);
                c2 = fmaf(gx, v[i].x, c2); c2 = fmaf(gy, v[i].y, c2); c2 = fmaf(gz, v[i].z, c2); c2 = fmaf(gw, v[i].w, c2);
                if (rok) {
                    dg[i].x = fmaf(g[i].x, v[i].x, dg[i].x); dg[i].y = fmaf(g[i].y, v[i].y, dg[i].y);
                    dg[i].z = fmaf(g[i].z, v[i].z, dg[i].z); dg[i].w = fmaf(g[i].w, v[i].w, dg[i].w);
                    db[i].x += g[i].x; db[i].y += g[i].y; db[i].z += g[i].z; db[i].w += g[i].w;
                }
            }
        }
#pragma unroll
        for (int o = LP / 2; o > 0; o >>= 1) {
            c1 += __shfl_xor_sync(0xffffffffu, c1, o);
            c2 += __shfl_xor_sync(0xffffffffu, c2, o);
        }
        c1 *= inv_d; c2 *= inv_d;
        float4* dst = reinterpret_cast<float4*>(dx + (rok ? row : 0) * D);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int j = sub + LP * i;
            if (rok && j < V) {
                dst[j] = make_float4(rstd * (g[i].x * wr[i].x - c1 - v[i].x * c2), rstd * (g[i].y * wr[i].y - c1 - v[i].y * c2),
                                     rstd * (g[i].z * wr[i].z - c1 - v[i].z * c2), rstd * (g[i].w * wr[i].w - c1 - v[i].w * c2));
            }
        }
    }
    // fold the per-lane column sums over the CTA's NW * PW row groups, one partial row per CTA
    const int grp = warp * PW + pw;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int j = sub + LP * i;
        if (j < V) {
            reinterpret_cast<float4*>(s_red + (grp * 2 + 0) * D)[j] = dg[i];
            reinterpret_cast<float4*>(s_red + (grp * 2 + 1) * D)[j] = db[i];
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < 2 * D; c += NW * 32) {
        const int which = c / D, ch = c % D;
        float acc = 0.f;
#pragma unroll 4
        for (int r = 0; r < NW * PW; ++r) acc += s_red[(r * 2 + which) * D + ch];
        partial[(static_cast<int64_t>(blockIdx.x) * 2 + which) * D + ch] = acc;
    }
}

// 32 columns x 8 row groups per CTA: every thread sums the partial rows r = rg, rg + 8, ... with eight loads in flight per round
// (one thread per column walking all <= 296 rows four at a time was a chain of 74 L2 round trips: 13 us per launch, 237 launches
// per training step), then the 8 row groups are added in a fixed order -- the result stays deterministic.
__global__ void __launch_bounds__(256)
layer_norm_bwd_finish_kernel(const float* __restrict__ partial, float* __restrict__ dgamma, float* __restrict__ dbeta, int D, int n_part) {
    __shared__ float s_acc[8][33];
    const int cl = threadIdx.x & 31, rg = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    const bool ok = c < 2 * D;
    const int which = ok ? c / D : 0, ch = ok ? c % D : 0;
    float acc = 0.f;
    if (ok) {
        const float* base = partial + static_cast<int64_t>(which) * D + ch;        // partial row r at + r * 2 * D
        float a[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = 0.f;
        for (int r0 = rg; r0 < n_part; r0 += 64) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = r0 + 8 * i;
                v[i] = r < n_part ? __ldg(base + static_cast<int64_t>(r) * 2 * D) : 0.f;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] += v[i];
        }
        acc = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
    }
    s_acc[rg][cl] = acc;
    __syncthreads();
    if (rg == 0 && ok) {
        float t = s_acc[0][cl];
#pragma unroll
        for (int k = 1; k < 8; ++k) t += s_acc[k][cl];
        if (which == 0) { if (dgamma) dgamma[ch] = t; } else { if (dbeta) dbeta[ch] = t; }
    }
}

constexpr int kLnBwdNW = 8;
constexpr int kLnBwdMaxCtas = 148 * 2;

// number of CTAs (= partial rows) the backward uses for this shape
int layer_norm_bwd_ctas(int dim, int64_t rows) {
    const int V = dim / 4;
    const int LP = V <= 16 ? 8 : (V <= 64 ? 16 : 32);
    const int64_t passes = (rows + kLnBwdNW * (32 / LP) - 1) / (kLnBwdNW * (32 / LP));
    return static_cast<int>(passes < kLnBwdMaxCtas ? passes : kLnBwdMaxCtas);
}

template <int LP, int NV>
static cudaError_t launch_ln_bwd(const FmNormBwdParams& p, cudaStream_t st) {
    constexpr int NW = kLnBwdNW;
    const int grid = layer_norm_bwd_ctas(p.dim, p.rows);
    const size_t smem = sizeof(float) * NW * (32 / LP) * 2 * static_cast<size_t>(p.dim);
    auto kern = layer_norm_bwd_kernel<NW, LP, NV>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    kern<<<grid, NW * 32, smem, st>>>(static_cast<const float*>(p.x), static_cast<const float*>(p.dy),
                                      static_cast<const float*>(p.weight), static_cast<float*>(p.dx),
                                      static_cast<float*>(p.workspace), p.dim, p.rows, p.eps);
    count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    layer_norm_bwd_finish_kernel<<<(2 * p.dim + 31) / 32, 256, 0, st>>>(static_cast<const float*>(p.workspace),
                                                                          static_cast<float*>(p.dweight),
                                                                          static_cast<float*>(p.dbias), p.dim, grid);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_layer_norm_bwd(const FmNormBwdParams& p, cudaStream_t st) {
    const int V = p.dim / 4;
    if (V <= 8) return launch_ln_bwd<8, 1>(p, st);
    if (V <= 16) return launch_ln_bwd<8, 2>(p, st);
    if (V <= 32) return launch_ln_bwd<16, 2>(p, st);
    if (V <= 48) return launch_ln_bwd<16, 3>(p, st);
    if (V <= 64) return launch_ln_bwd<16, 4>(p, st);
    if (V <= 96) return launch_ln_bwd<32, 3>(p, st);
    if (V <= 128) return launch_ln_bwd<32, 4>(p, st);
    if (V <= 192) return launch_ln_bwd<32, 6>(p, st);
    if (V <= 256) return launch_ln_bwd<32, 8>(p, st);
    return cudaErrorInvalidConfiguration;
}

}  // namespace fm
