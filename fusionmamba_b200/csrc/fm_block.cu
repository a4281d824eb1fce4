// fm_block.cu -- the inference tail of the reference's VSSBlock_new around the SS2D path, for sm_100a.
//
// After the SS2D op the reference block runs (models/cross.py:1362-1377)
//     x_  = ECA(x_ssm)            x_ssm * sigmoid(conv1d_k3(mean_hw(x_ssm)))                      :1236-1259 (eca_layer)
//     x   = x_ssm + x_            -> LDC conv                                                       :1368-1369
//     x   = se(x_ssm) + se(x_conv)      se(v) = v * sigmoid(W2 gelu(W1 mean_hw(LayerNorm(v)) + b1) + b2)   :744-768 (BiAttn)
//     x   = input + x ;  x = x + mlp(norm2(x))                                                        :1373-1376
// as ~45 ATen kernels per block -- per-(batch, channel) means, tiny linears, broadcasts, residual adds, casts -- each a few
// microseconds on tensors that fit L2 at three of the four stages, i.e. launch bound even inside a CUDA graph (2169 of the 3235
// kernels of one forward, half of its GPU time: profiles/r02_breakdown_swapped_ln.json).  Three kernels replace them:
//   fm_block_gates        one pass over v (B, P, C): per-(b, c) mean of v AND of LayerNorm(v) (row statistics per position, column
//                         sums in registers, one partial row per CTA), then per batch item the ECA gate 1 + sigmoid(conv1d(mean))
//                         and the BiAttn gate sigmoid(W2 gelu(W1 m + b1) + b2)  -- (B, C) fp32 each
//   fm_block_scale        y = v + v * (s - 1)   with the reference's roundings (ECA apply + the add feeding the LDC conv)
//   fm_block_combine_norm x' = input + (x_ssm * a1 + x_conv * a2)  (fp32 residual stream)  and  LayerNorm(x') in the activation
//                         dtype for mlp.fc1 -- the BiAttn applies, both adds and norm2 in one row pass
// Activations are fp32, bf16 or fp16 (the autocast dtype); gates, statistics and the residual stream are fp32.  Products and
// sums of 16-bit activations are rounded where the reference's separate ATen ops round, so the tail is not "more accurate than"
// but equal to the reference within one rounding of the gate values (which the reference computes in 16 bits, here fp32).
#include "fm_common.cuh"
#include "fm_launch.h"

namespace fm {

template <typename T> struct BV4 { using type = uint2; };
template <> struct BV4<float> { using type = float4; };

template <typename T>
__device__ __forceinline__ float4 bload4(const T* p) {
    if constexpr (sizeof(T) == 4) {
        return __ldg(reinterpret_cast<const float4*>(p));
    } else {
        const uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
        const T* e = reinterpret_cast<const T*>(&r);
        return make_float4(Cvt<T>::to_f(e[0]), Cvt<T>::to_f(e[1]), Cvt<T>::to_f(e[2]), Cvt<T>::to_f(e[3]));
    }
}
template <typename T>
__device__ __forceinline__ void bstore4(T* p, float4 v) {
    if constexpr (sizeof(T) == 4) {
        *reinterpret_cast<float4*>(p) = v;
    } else {
        uint2 o;
        T* e = reinterpret_cast<T*>(&o);
        e[0] = Cvt<T>::from_f(v.x); e[1] = Cvt<T>::from_f(v.y); e[2] = Cvt<T>::from_f(v.z); e[3] = Cvt<T>::from_f(v.w);
        *reinterpret_cast<uint2*>(p) = o;
    }
}
template <typename T> __device__ __forceinline__ float rnd(float v) { return Cvt<T>::to_f(Cvt<T>::from_f(v)); }

// ---- pass 1: per-(b, c) sums of v and of (v - mean_row) * rstd_row -------------------------------------------------------------
template <typename T, int NW, int LP, int NV>
__global__ void __launch_bounds__(NW * 32)
block_stats_kernel(const T* __restrict__ v, float* __restrict__ partial, int C, int P, int rows_per_slab, float eps) {
    constexpr int PW = 32 / LP;
    extern __shared__ __align__(16) float s_red[];           // [NW * PW][2][C]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane % LP, pw = lane / LP;
    const int V = C >> 2;
    const float inv_c = 1.f / C;
    const int b = blockIdx.y, slab = blockIdx.x;
    const int p0 = slab * rows_per_slab, p1 = min(P, p0 + rows_per_slab);
    const T* vb = v + static_cast<int64_t>(b) * P * C;
    float4 sx[NV], sn[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) { sx[i] = make_float4(0.f, 0.f, 0.f, 0.f); sn[i] = make_float4(0.f, 0.f, 0.f, 0.f); }
    for (int r0 = p0 + warp * PW; r0 < p1; r0 += NW * PW) {
        const int row = r0 + pw;
        const bool rok = row < p1;
        float4 x[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int j = sub + LP * i;
            x[i] = (rok && j < V) ? bload4<T>(vb + static_cast<int64_t>(row) * C + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const float shift = __shfl_sync(0xffffffffu, x[0].x, pw * LP);
        float s = 0.f, q = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int j = sub + LP * i;
            if (j < V) {
                const float a0 = x[i].x - shift, a1 = x[i].y - shift, a2 = x[i].z - shift, a3 = x[i].w - shift;
                s += (a0 + a1) + (a2 + a3);
                q = fmaf(a0, a0, q); q = fmaf(a1, a1, q); q = fmaf(a2, a2, q); q = fmaf(a3, a3, q);
            }
        }
#pragma unroll
        for (int o = LP / 2; o > 0; o >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            q += __shfl_xor_sync(0xffffffffu, q, o);
        }
        const float m = s * inv_c;
        const float rstd = rsqrtf(fmaxf(q * inv_c - m * m, 0.f) + eps);
        const float nm = -(m + shift) * rstd;
        if (rok) {
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                sx[i].x += x[i].x; sx[i].y += x[i].y; sx[i].z += x[i].z; sx[i].w += x[i].w;
                sn[i].x += fmaf(x[i].x, rstd, nm); sn[i].y += fmaf(x[i].y, rstd, nm);
                sn[i].z += fmaf(x[i].z, rstd, nm); sn[i].w += fmaf(x[i].w, rstd, nm);
            }
        }
    }
    const int grp = warp * PW + pw;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int j = sub + LP * i;
        if (j < V) {
            reinterpret_cast<float4*>(s_red + (grp * 2 + 0) * C)[j] = sx[i];
            reinterpret_cast<float4*>(s_red + (grp * 2 + 1) * C)[j] = sn[i];
        }
    }
    __syncthreads();
    float* dst = partial + (static_cast<int64_t>(b) * gridDim.x + slab) * 2 * C;
    for (int c = threadIdx.x; c < 2 * C; c += NW * 32) {
        const int which = c / C, ch = c % C;
        float acc = 0.f;
#pragma unroll 4
        for (int r = 0; r < NW * PW; ++r) acc += s_red[(r * 2 + which) * C + ch];
        dst[which * C + ch] = acc;
    }
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.7071067811865476f)); }

// ---- pass 2 (one CTA per batch item): the two gates ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
block_gates_finish_kernel(const float* __restrict__ partial, int n_slab, int C, int P, int R,
                          const float* __restrict__ ln_w, const float* __restrict__ ln_b, const float* __restrict__ eca_w,
                          const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
                          const float* __restrict__ b2, float* __restrict__ eca_scale, float* __restrict__ se_gate) {
    extern __shared__ __align__(16) float sm[];              // mean_x[C] | m[C] | g[R]
    float* mean_x = sm;
    float* mvec = sm + C;
    float* gvec = sm + 2 * C;
    const int b = blockIdx.x;
    const float inv_p = 1.f / P;
    const float* pb = partial + static_cast<int64_t>(b) * n_slab * 2 * C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        // eight slabs per round with independent accumulators: the kernel is a chain of L2 round trips, not of adds
        float ax8[8], an8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { ax8[i] = 0.f; an8[i] = 0.f; }
        for (int s = 0; s < n_slab; s += 8) {
            float vx[8], vn[8];                                  // loads first, adds after: one round trip per eight slabs
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const bool in = s + i < n_slab;
                vx[i] = in ? __ldg(pb + (s + i) * 2 * C + c) : 0.f;
                vn[i] = in ? __ldg(pb + (s + i) * 2 * C + C + c) : 0.f;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) { ax8[i] += vx[i]; an8[i] += vn[i]; }
        }
        const float ax = ((ax8[0] + ax8[1]) + (ax8[2] + ax8[3])) + ((ax8[4] + ax8[5]) + (ax8[6] + ax8[7]));
        const float an = ((an8[0] + an8[1]) + (an8[2] + an8[3])) + ((an8[4] + an8[5]) + (an8[6] + an8[7]));
        mean_x[c] = ax * inv_p;
        mvec[c] = fmaf(ln_w ? ln_w[c] : 1.f, an * inv_p, ln_b ? ln_b[c] : 0.f);     // mean_hw(LayerNorm(v))[c]
    }
    __syncthreads();
    if (eca_scale != nullptr) {
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            const float l = c > 0 ? mean_x[c - 1] : 0.f, r = c + 1 < C ? mean_x[c + 1] : 0.f;
            const float y = eca_w[0] * l + eca_w[1] * mean_x[c] + eca_w[2] * r;          // Conv1d(1, 1, 3, padding 1, no bias) over channels
            eca_scale[static_cast<int64_t>(b) * C + c] = 1.f / (1.f + __expf(-y));
        }
    }
    if (se_gate != nullptr) {
        // global_reduce (R x C) : one warp per output row, lanes along the contraction with 16-byte loads (a thread per row walks
        // C strided elements serially: 30 us of pure load latency per launch at C = 768, profiles/r02_breakdown_fused_blocks.json)
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
        const bool v4 = (C & 3) == 0 && (reinterpret_cast<uintptr_t>(w1) & 15u) == 0;
        // four output rows per round (their loads in flight together), rows warp, warp + nw, ...
        for (int j0 = warp; j0 < R; j0 += 4 * nw) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            if (v4 && C <= 768) {
                // C / 128 <= 6 column steps x 4 rows: every load of the round is issued before the first use
                float4 wv[6][4];
#pragma unroll
                for (int it = 0; it < 6; ++it) {
                    const int c4 = lane + 32 * it;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int j = j0 + i * nw;
                        wv[it][i] = (c4 < (C >> 2) && j < R) ? __ldg(reinterpret_cast<const float4*>(w1 + static_cast<int64_t>(j) * C) + c4)
                                                             : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
#pragma unroll
                for (int it = 0; it < 6; ++it) {
                    const int c4 = lane + 32 * it;
                    if (c4 < (C >> 2)) {
                        const float4 m = *reinterpret_cast<const float4*>(mvec + 4 * c4);
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            acc[i] = fmaf(wv[it][i].x, m.x, fmaf(wv[it][i].y, m.y, fmaf(wv[it][i].z, m.z, fmaf(wv[it][i].w, m.w, acc[i]))));
                    }
                }
            } else {
                for (int c = lane; c < C; c += 32) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int j = j0 + i * nw;
                        if (j < R) acc[i] = fmaf(w1[static_cast<int64_t>(j) * C + c], mvec[c], acc[i]);
                    }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
            }
            if (lane == 0) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int j = j0 + i * nw;
                    if (j < R) gvec[j] = gelu_erf(acc[i] + (b1 ? b1[j] : 0.f));
                }
            }
        }
        __syncthreads();
        // channel_select (C x R): a thread per output channel, its R-element row read as whole 16-byte pieces
        const bool r4 = (R & 3) == 0 && (reinterpret_cast<uintptr_t>(w2) & 15u) == 0 && ((2 * C) & 3) == 0;
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            float acc = b2 ? b2[c] : 0.f;
            const float* wr = w2 + static_cast<int64_t>(c) * R;
            if (r4 && R <= 96) {
                float4 wv[24];                           // R / 4 <= 24 pieces: one round trip for the whole row
#pragma unroll
                for (int j4 = 0; j4 < 24; ++j4)
                    wv[j4] = j4 < (R >> 2) ? __ldg(reinterpret_cast<const float4*>(wr) + j4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int j4 = 0; j4 < 24; ++j4) {
                    if (j4 < (R >> 2)) {
                        const float4 g = *reinterpret_cast<const float4*>(gvec + 4 * j4);
                        acc = fmaf(wv[j4].x, g.x, fmaf(wv[j4].y, g.y, fmaf(wv[j4].z, g.z, fmaf(wv[j4].w, g.w, acc))));
                    }
                }
            } else {
                for (int j = 0; j < R; ++j) acc = fmaf(wr[j], gvec[j], acc);
            }
            se_gate[static_cast<int64_t>(b) * C + c] = 1.f / (1.f + __expf(-acc));
        }
    }
}

// ---- ECA apply + add: y = v + v * g  (each step rounded to T like the reference's separate ops) ----------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
block_scale_kernel(const T* __restrict__ v, const float* __restrict__ gate, T* __restrict__ y, int C, int P, int64_t n_vec) {
    const int V = C >> 2;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_vec; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int j = static_cast<int>(i % V);
        const int64_t row = i / V;
        const int b = static_cast<int>(row / P);
        const float4 x = bload4<T>(v + 4 * i);
        const float4 g = __ldg(reinterpret_cast<const float4*>(gate + static_cast<int64_t>(b) * C) + j);
        float4 o;
        o.x = rnd<T>(x.x + rnd<T>(x.x * rnd<T>(g.x))); o.y = rnd<T>(x.y + rnd<T>(x.y * rnd<T>(g.y)));
        o.z = rnd<T>(x.z + rnd<T>(x.z * rnd<T>(g.z))); o.w = rnd<T>(x.w + rnd<T>(x.w * rnd<T>(g.w)));
        bstore4<T>(y + 4 * i, o);
    }
}

// ---- combine + norm2: x' = input + (x_ssm * a1 + x_conv * a2);  y = LayerNorm(x') in T ------------------------------------------------
// TI: dtype of the residual stream (fp32 at stage 0; the autocast dtype wherever a Linear produced it, e.g. behind PatchMerging2D)
template <typename T, typename TI, int NW, int LP, int NV>
__global__ void __launch_bounds__(NW * 32)
block_combine_norm_kernel(const TI* __restrict__ input, const T* __restrict__ xs, const T* __restrict__ xc,
                          const float* __restrict__ a1, const float* __restrict__ a2, const float* __restrict__ w,
                          const float* __restrict__ bsh, TI* __restrict__ xout, T* __restrict__ yout, int C, int P,
                          int64_t rows, float eps) {
    constexpr int PW = 32 / LP;
    const int lane = threadIdx.x & 31;
    const int sub = lane % LP, pw = lane / LP;
    const int V = C >> 2;
    const float inv_c = 1.f / C;
    float4 wr[NV], br[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int j = sub + LP * i;
        wr[i] = (w != nullptr && j < V) ? __ldg(reinterpret_cast<const float4*>(w) + j) : make_float4(1.f, 1.f, 1.f, 1.f);
        br[i] = (bsh != nullptr && j < V) ? __ldg(reinterpret_cast<const float4*>(bsh) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const int64_t stride = static_cast<int64_t>(gridDim.x) * NW * PW;
    for (int64_t row0 = (static_cast<int64_t>(blockIdx.x) * NW + (threadIdx.x >> 5)) * PW; row0 < rows; row0 += stride) {
        const int64_t row = row0 + pw;
        const bool rok = row < rows;
        const int64_t rr = rok ? row : 0;
        const int b = static_cast<int>(rr / P);
        float4 x[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int j = sub + LP * i;
            if (rok && j < V && xs == nullptr) {
                x[i] = bload4<TI>(input + rr * C + 4 * j);           // norm-only form: y = LayerNorm(input) in T, nothing else written
            } else if (rok && j < V) {
                const float4 in = bload4<TI>(input + rr * C + 4 * j);
                const float4 s = bload4<T>(xs + rr * C + 4 * j), c = bload4<T>(xc + rr * C + 4 * j);
                const float4 g1 = __ldg(reinterpret_cast<const float4*>(a1 + static_cast<int64_t>(b) * C) + j);
                const float4 g2 = __ldg(reinterpret_cast<const float4*>(a2 + static_cast<int64_t>(b) * C) + j);
                x[i].x = rnd<TI>(in.x + rnd<T>(rnd<T>(s.x * rnd<T>(g1.x)) + rnd<T>(c.x * rnd<T>(g2.x))));
                x[i].y = rnd<TI>(in.y + rnd<T>(rnd<T>(s.y * rnd<T>(g1.y)) + rnd<T>(c.y * rnd<T>(g2.y))));
                x[i].z = rnd<TI>(in.z + rnd<T>(rnd<T>(s.z * rnd<T>(g1.z)) + rnd<T>(c.z * rnd<T>(g2.z))));
                x[i].w = rnd<TI>(in.w + rnd<T>(rnd<T>(s.w * rnd<T>(g1.w)) + rnd<T>(c.w * rnd<T>(g2.w))));
                bstore4<TI>(xout + rr * C + 4 * j, x[i]);
            } else {
                x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        const float shift = __shfl_sync(0xffffffffu, x[0].x, pw * LP);
        float s = 0.f, q = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int j = sub + LP * i;
            if (j < V) {
                x[i].x -= shift; x[i].y -= shift; x[i].z -= shift; x[i].w -= shift;
                s += (x[i].x + x[i].y) + (x[i].z + x[i].w);
                q = fmaf(x[i].x, x[i].x, q); q = fmaf(x[i].y, x[i].y, q); q = fmaf(x[i].z, x[i].z, q); q = fmaf(x[i].w, x[i].w, q);
            }
        }
#pragma unroll
        for (int o = LP / 2; o > 0; o >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            q += __shfl_xor_sync(0xffffffffu, q, o);
        }
        const float m = s * inv_c;
        const float rstd = rsqrtf(fmaxf(q * inv_c - m * m, 0.f) + eps);
        const float nm = -m * rstd;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int j = sub + LP * i;
            if (rok && j < V) {
                bstore4<T>(yout + rr * C + 4 * j,
                           make_float4(fmaf(fmaf(x[i].x, rstd, nm), wr[i].x, br[i].x), fmaf(fmaf(x[i].y, rstd, nm), wr[i].y, br[i].y),
                                       fmaf(fmaf(x[i].z, rstd, nm), wr[i].z, br[i].z), fmaf(fmaf(x[i].w, rstd, nm), wr[i].w, br[i].w)));
            }
        }
    }
}

// ---- launchers --------------------------------------------------------------------------------------------------------------------------
constexpr int kBlkNW = 8;

int block_gates_slabs(int batch, int positions) {
    int n = (2 * 592 + batch - 1) / batch;                    // ~2 CTAs per SM sub-partition over the whole batch
    const int max_slabs = (positions + 31) / 32;              // at least 32 rows per CTA
    if (n > max_slabs) n = max_slabs;
    if (n > 16) n = 16;                                       // (the finish kernel walks the slabs: 16 = two load rounds)
    return n < 1 ? 1 : n;
}

template <typename T, int LP, int NV>
static cudaError_t launch_stats(const FmBlockGatesParams& p, cudaStream_t st, int n_slab) {
    constexpr int NW = kBlkNW;
    const int rows_per_slab = (p.positions + n_slab - 1) / n_slab;
    const size_t smem = sizeof(float) * NW * (32 / LP) * 2 * static_cast<size_t>(p.dim);
    auto kern = block_stats_kernel<T, NW, LP, NV>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    kern<<<dim3(n_slab, p.batch), NW * 32, smem, st>>>(static_cast<const T*>(p.x), static_cast<float*>(p.workspace), p.dim,
                                                       p.positions, rows_per_slab, p.eps);
    count_launch();
    return cudaGetLastError();
}

template <typename T>
static cudaError_t launch_block_gates_T(const FmBlockGatesParams& p, cudaStream_t st) {
    const int n_slab = block_gates_slabs(p.batch, p.positions);
    const int V = p.dim / 4;
    cudaError_t e;
    if (V <= 8) e = launch_stats<T, 8, 1>(p, st, n_slab);
    else if (V <= 16) e = launch_stats<T, 8, 2>(p, st, n_slab);
    else if (V <= 32) e = launch_stats<T, 16, 2>(p, st, n_slab);
    else if (V <= 48) e = launch_stats<T, 16, 3>(p, st, n_slab);
    else if (V <= 64) e = launch_stats<T, 16, 4>(p, st, n_slab);
    else if (V <= 96) e = launch_stats<T, 32, 3>(p, st, n_slab);
    else if (V <= 128) e = launch_stats<T, 32, 4>(p, st, n_slab);
    else if (V <= 192) e = launch_stats<T, 32, 6>(p, st, n_slab);
    else if (V <= 256) e = launch_stats<T, 32, 8>(p, st, n_slab);
    else return cudaErrorInvalidConfiguration;
    if (e != cudaSuccess) return e;
    const size_t smem = sizeof(float) * (2 * static_cast<size_t>(p.dim) + p.reduce_dim);
    block_gates_finish_kernel<<<p.batch, 256, smem, st>>>(
        static_cast<const float*>(p.workspace), n_slab, p.dim, p.positions, p.reduce_dim, static_cast<const float*>(p.ln_weight),
        static_cast<const float*>(p.ln_bias), static_cast<const float*>(p.eca_weight), static_cast<const float*>(p.w1),
        static_cast<const float*>(p.b1), static_cast<const float*>(p.w2), static_cast<const float*>(p.b2),
        static_cast<float*>(p.eca_scale), static_cast<float*>(p.se_gate));
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_block_gates(const FmBlockGatesParams& p, cudaStream_t st) {
    switch (p.dtype) {
        case FM_F32: return launch_block_gates_T<float>(p, st);
        case FM_F16: return launch_block_gates_T<__half>(p, st);
        default: return launch_block_gates_T<__nv_bfloat16>(p, st);
    }
}

template <typename T>
static cudaError_t launch_block_scale_T(const FmBlockScaleParams& p, cudaStream_t st) {
    const int64_t n_vec = static_cast<int64_t>(p.batch) * p.positions * (p.dim / 4);
    const int64_t blocks = (n_vec + 255) / 256;
    const unsigned grid = static_cast<unsigned>(blocks < 148 * 16 ? blocks : 148 * 16);
    block_scale_kernel<T><<<grid, 256, 0, st>>>(static_cast<const T*>(p.x), static_cast<const float*>(p.gate), static_cast<T*>(p.y),
                                                p.dim, p.positions, n_vec);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_block_scale(const FmBlockScaleParams& p, cudaStream_t st) {
    switch (p.dtype) {
        case FM_F32: return launch_block_scale_T<float>(p, st);
        case FM_F16: return launch_block_scale_T<__half>(p, st);
        default: return launch_block_scale_T<__nv_bfloat16>(p, st);
    }
}

template <typename T, typename TI, int LP, int NV>
static cudaError_t launch_combine(const FmBlockCombineParams& p, cudaStream_t st) {
    constexpr int NW = kBlkNW, PW = 32 / LP;
    const int64_t rows = static_cast<int64_t>(p.batch) * p.positions;
    const int64_t passes = (rows + NW * PW - 1) / (NW * PW);
    const int64_t cap = 148 * 8 * 4;
    const unsigned grid = static_cast<unsigned>(passes < cap ? passes : cap);
    block_combine_norm_kernel<T, TI, NW, LP, NV><<<grid, NW * 32, 0, st>>>(
        static_cast<const TI*>(p.input), static_cast<const T*>(p.x_ssm), static_cast<const T*>(p.x_conv),
        static_cast<const float*>(p.gate_ssm), static_cast<const float*>(p.gate_conv), static_cast<const float*>(p.ln_weight),
        static_cast<const float*>(p.ln_bias), static_cast<TI*>(p.x_out), static_cast<T*>(p.y_out), p.dim, p.positions, rows, p.eps);
    count_launch();
    return cudaGetLastError();
}

template <typename T, typename TI>
static cudaError_t launch_block_combine_TT(const FmBlockCombineParams& p, cudaStream_t st) {
    const int V = p.dim / 4;
    if (V <= 8) return launch_combine<T, TI, 8, 1>(p, st);
    if (V <= 16) return launch_combine<T, TI, 8, 2>(p, st);
    if (V <= 32) return launch_combine<T, TI, 16, 2>(p, st);
    if (V <= 48) return launch_combine<T, TI, 16, 3>(p, st);
    if (V <= 64) return launch_combine<T, TI, 16, 4>(p, st);
    if (V <= 96) return launch_combine<T, TI, 32, 3>(p, st);
    if (V <= 128) return launch_combine<T, TI, 32, 4>(p, st);
    if (V <= 192) return launch_combine<T, TI, 32, 6>(p, st);
    if (V <= 256) return launch_combine<T, TI, 32, 8>(p, st);
    return cudaErrorInvalidConfiguration;
}

// the residual stream is fp32 or has the activation dtype
cudaError_t launch_block_combine(const FmBlockCombineParams& p, cudaStream_t st) {
    const bool in32 = p.input_dtype == FM_F32;
    switch (p.dtype) {
        case FM_F32: return launch_block_combine_TT<float, float>(p, st);
        case FM_F16: return in32 ? launch_block_combine_TT<__half, float>(p, st) : launch_block_combine_TT<__half, __half>(p, st);
        default: return in32 ? launch_block_combine_TT<__nv_bfloat16, float>(p, st) : launch_block_combine_TT<__nv_bfloat16, __nv_bfloat16>(p, st);
    }
}

}  // namespace fm
