// fm_scan_fwd.cuh -- selective-scan forward for sm_100a.
//
// Replaces selective_scan_fwd_kernel (selective_scan/selective_scan_fwd_kernel.cuh:67-303) with a different
// decomposition (this is not a port):
//   * one CTA owns R = NW*(32/G) channel rows of ONE (batch, group) and walks the sequence in chunks of
//     TC = S*G timesteps; the [dstate x TC] B and C tiles are staged ONCE per chunk in shared memory
//     (cp.async double buffer, lane-segment-padded layout => conflict-free LDS.128) and shared by all R rows --
//     the reference re-reads them from L2 for every row.
//   * a row is scanned by G lanes of one warp; each lane owns S consecutive timesteps: thread-serial recurrence
//     from zero (up-sweep), G-lane warp-shuffle combine of (decay, state) aggregates with the monoid
//     (a0,b0)o(a1,b1) = (a1*a0, a1*b0+b1), then a second serial pass seeded with the lane's incoming state that
//     also accumulates y += C*h.  a_t is computed once (one ex2 per (t, state)) and kept in registers between
//     the two passes; a segment's aggregate decay is exp2(A * sum(delta)) (one ex2 per lane, not S multiplies).
//   * NS states are processed per loop trip (independent dependency chains interleaved for ILP).
//   * the running state is carried across chunks per (row, state) in shared memory, owned by one lane; warps are
//     independent inside the state loop -- only two block barriers per chunk guard the B/C tile ring.
//   * u / delta / z / out move as 128-bit vector accesses; every per-chunk index (checkpoint slots, tails) is
//     computed once outside the state loop.
#pragma once
#include "fm_common.cuh"
#include "fm_launch.h"
#include "fm_scan_fwd16.cuh"
#include "fm_scan_fwd_rp.cuh"

namespace fm {

template <typename T, int S, int G, int NW, int NS, bool kHasZ, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB)
scan_fwd_kernel(const FmScanFwdParams p, const int vec_io, const int vec_bc) {
    constexpr int TC = G * S;               // timesteps per chunk
    constexpr int RW = 32 / G;              // rows per warp
    constexpr int R = NW * RW;              // rows per CTA
    constexpr int SP = seg_pad(S);
    constexpr int ROWP = G * SP;            // smem pitch of one state row of the B/C tile (floats)
    constexpr int NT = NW * 32;

    const int N = p.dstate;
    const int L = p.seqlen;
    const int dg = p.dim / p.n_groups;                  // channels per group
    const int tiles_per_group = (dg + R - 1) / R;
    const int group = blockIdx.x / tiles_per_group;
    const int tile = blockIdx.x % tiles_per_group;
    const int b = blockIdx.y;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int seg = lane % G;                           // which S-step segment of the chunk
    const int rl = warp * RW + lane / G;                // row within CTA
    const int dloc = tile * R + rl;                     // channel within group
    const bool row_ok = dloc < dg;
    const int d = group * dg + (row_ok ? dloc : 0);     // invalid rows shadow row 0 of the group and never store

    extern __shared__ __align__(16) float smem[];
    float* sBC = smem;                                  // [2 stages][B|C][N][ROWP]
    float* sA2 = sBC + 4 * N * ROWP;                    // [R][N]  A * log2(e)
    float* sH = sA2 + R * N;                            // [R][N]  running state (touched only by the seg==0 lane)

    const T* __restrict__ Bg = reinterpret_cast<const T*>(p.B) + b * p.B_batch_stride + group * p.B_group_stride;
    const T* __restrict__ Cg = reinterpret_cast<const T*>(p.C) + b * p.C_batch_stride + group * p.C_group_stride;
    const T* __restrict__ urow = reinterpret_cast<const T*>(p.u) + b * p.u_batch_stride + d * p.u_d_stride;
    const T* __restrict__ drow = reinterpret_cast<const T*>(p.delta) + b * p.delta_batch_stride + d * p.delta_d_stride;
    T* __restrict__ orow = reinterpret_cast<T*>(p.out) + b * p.out_batch_stride + d * p.out_d_stride;
    const T* __restrict__ zrow = nullptr;
    T* __restrict__ ozrow = nullptr;
    if constexpr (kHasZ) {
        zrow = reinterpret_cast<const T*>(p.z) + b * p.z_batch_stride + d * p.z_d_stride;
        ozrow = reinterpret_cast<T*>(p.out_z) + b * p.out_z_batch_stride + d * p.out_z_d_stride;
    }
    const int64_t rowid = static_cast<int64_t>(b) * p.dim + d;
    float* __restrict__ xrow = reinterpret_cast<float*>(p.x) + rowid * p.n_chunks * 2 * N;
    float* __restrict__ hckrow = p.hck ? reinterpret_cast<float*>(p.hck) + rowid * p.n_hck * N : nullptr;

    const float Dval = p.D ? reinterpret_cast<const float*>(p.D)[d] : 0.f;
    const float bias = p.delta_bias ? reinterpret_cast<const float*>(p.delta_bias)[d] : 0.f;

    for (int i = tid; i < R * N; i += NT) {
        int r = i / N, n = i % N;
        int dl_ = tile * R + r;
        int dd = group * dg + (dl_ < dg ? dl_ : 0);
        sA2[i] = reinterpret_cast<const float*>(p.A)[dd * p.A_d_stride + n * p.A_dstate_stride] * kLog2e;
        sH[i] = 0.f;
    }

    const int n_chunks = (L + TC - 1) / TC;
    stage_tile<T, TC, S>(sBC, Bg, p.B_dstate_stride, N, 0, L, vec_bc, tid, NT);
    stage_tile<T, TC, S>(sBC + N * ROWP, Cg, p.C_dstate_stride, N, 0, L, vec_bc, tid, NT);
    cp_async_commit();

    float sum_total = 0.f;  // running sum of delta over the row (only for the decay product stored in x)
    const float* myA2 = sA2 + rl * N;
    float* myH = sH + rl * N;

    for (int c = 0; c < n_chunks; ++c) {
        const int stage = c & 1;
        if (c + 1 < n_chunks) {
            float* nxt = sBC + (stage ^ 1) * 2 * N * ROWP;
            stage_tile<T, TC, S>(nxt, Bg, p.B_dstate_stride, N, (c + 1) * TC, L, vec_bc, tid, NT);
            stage_tile<T, TC, S>(nxt + N * ROWP, Cg, p.C_dstate_stride, N, (c + 1) * TC, L, vec_bc, tid, NT);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();

        const int t0 = c * TC + seg * S;
        const int nvalid = L - t0;                      // may be <= 0 or > S
        constexpr int H = S / 2;                        // time-adjacent element pairs of the lane segment
        float2 dl2[H], du2[H], y2[H];
        {
            float uv[S], dl[S];
            load_seg<T, S>(urow + t0, nvalid, vec_io, uv);
            load_seg<T, S>(drow + t0, nvalid, vec_io, dl);
#pragma unroll
            for (int i = 0; i < S; ++i) {
                float xv = dl[i] + bias;
                float sp = p.delta_softplus ? softplus_fast(xv) : xv;
                dl[i] = (i < nvalid) ? sp : 0.f;        // masked steps: a = 1, b = 0
            }
#pragma unroll
            for (int j = 0; j < H; ++j) {
                const float2 u2 = make_float2(uv[2 * j], uv[2 * j + 1]);
                dl2[j] = make_float2(dl[2 * j], dl[2 * j + 1]);
                du2[j] = mul2(dl2[j], u2);
                y2[j] = mul2(u2, bcast2(Dval));
            }
        }
        float sumd = 0.f;
#pragma unroll
        for (int j = 0; j < H; ++j) sumd += dl2[j].x + dl2[j].y;
        float rowsum = sumd;
#pragma unroll
        for (int o = 1; o < G; o <<= 1) rowsum += __shfl_xor_sync(0xffffffffu, rowsum, o, G);
        sum_total += rowsum;

        // per-chunk bookkeeping, hoisted out of the state loop (no integer division per state)
        const int t_end = min((c + 1) * TC, L);         // exclusive end of this chunk
        float* xslot = nullptr;                         // seg-0 lane writes the x checkpoint at slot ends / at L
        if (seg == 0 && row_ok && ((t_end % p.chunk_len == 0) || t_end == L))
            xslot = xrow + ((t_end - 1) / p.chunk_len) * 2 * N;
        float* hslot = nullptr;                         // dense checkpoint: state at the end of this lane's segment
        {
            const int te = t0 + S;
            if (hckrow != nullptr && row_ok && te < L && te % p.hck_len == 0) hslot = hckrow + (te / p.hck_len - 1) * N;
        }

        const float4* Bv = reinterpret_cast<const float4*>(sBC + stage * 2 * N * ROWP + seg * SP);
        const float4* Cv = Bv + (N * ROWP) / 4;

#pragma unroll 1
        for (int n = 0; n < N; n += NS) {
            float2 a2[NS][H], b2[NS][H];
            float h[NS], P[NS], hrun[NS];
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                const float A2 = myA2[n + s];
                hrun[s] = myH[n + s];
#pragma unroll
                for (int k = 0; k < S / 4; ++k) {
                    const float4 v = Bv[(s * ROWP) / 4 + k];
                    b2[s][2 * k] = mul2(du2[2 * k], make_float2(v.x, v.y));
                    b2[s][2 * k + 1] = mul2(du2[2 * k + 1], make_float2(v.z, v.w));
                }
#pragma unroll
                for (int j = 0; j < H; ++j) {
                    const float2 x2 = mul2(dl2[j], bcast2(A2));
                    a2[s][j].x = ex2_approx(x2.x);
                    a2[s][j].y = ex2_approx(x2.y);
                }
                P[s] = ex2_approx(A2 * sumd);
            }
            // up-sweep: segment state from zero (NS independent chains)
#pragma unroll
            for (int s = 0; s < NS; ++s) h[s] = fmaf(a2[s][0].y, b2[s][0].x, b2[s][0].y);
#pragma unroll
            for (int j = 1; j < H; ++j)
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    h[s] = fmaf(a2[s][j].x, h[s], b2[s][j].x);
                    h[s] = fmaf(a2[s][j].y, h[s], b2[s][j].y);
                }
            if (seg == 0)
#pragma unroll
                for (int s = 0; s < NS; ++s) h[s] = fmaf(P[s], hrun[s], h[s]);
            // inclusive combine over the G lanes of the row
#pragma unroll
            for (int o = 1; o < G; o <<= 1) {
                float Pp[NS], hp[NS];
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    hp[s] = __shfl_up_sync(0xffffffffu, h[s], o, G);
                    if (2 * o < G) Pp[s] = __shfl_up_sync(0xffffffffu, P[s], o, G);
                }
                if (seg >= o) {
#pragma unroll
                    for (int s = 0; s < NS; ++s) {
                        h[s] = fmaf(P[s], hp[s], h[s]);
                        if (2 * o < G) P[s] *= Pp[s];
                    }
                }
            }
            float hin[NS];
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                if (hslot != nullptr) hslot[n + s] = h[s];
                hin[s] = __shfl_up_sync(0xffffffffu, h[s], 1, G);
                const float hlast = __shfl_sync(0xffffffffu, h[s], G - 1, G);
                if (seg == 0) {
                    hin[s] = hrun[s];
                    myH[n + s] = hlast;
                    if (xslot != nullptr) {
                        xslot[2 * (n + s)] = ex2_approx(myA2[n + s] * sum_total);
                        xslot[2 * (n + s) + 1] = hlast;
                    }
                }
            }
            // down-sweep with the true incoming state; y += C * h (packed over time-adjacent pairs)
#pragma unroll
            for (int k = 0; k < S / 4; ++k) {
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    const float4 cv = Cv[(s * ROWP) / 4 + k];
                    const float h0 = fmaf(a2[s][2 * k].x, hin[s], b2[s][2 * k].x);
                    const float h1 = fmaf(a2[s][2 * k].y, h0, b2[s][2 * k].y);
                    const float h2_ = fmaf(a2[s][2 * k + 1].x, h1, b2[s][2 * k + 1].x);
                    const float h3 = fmaf(a2[s][2 * k + 1].y, h2_, b2[s][2 * k + 1].y);
                    hin[s] = h3;
                    y2[2 * k] = fma2(make_float2(cv.x, cv.y), make_float2(h0, h1), y2[2 * k]);
                    y2[2 * k + 1] = fma2(make_float2(cv.z, cv.w), make_float2(h2_, h3), y2[2 * k + 1]);
                }
            }
            Bv += (NS * ROWP) / 4;
            Cv += (NS * ROWP) / 4;
        }

        if (row_ok && nvalid > 0) {
            float y[S];
#pragma unroll
            for (int j = 0; j < H; ++j) { y[2 * j] = y2[j].x; y[2 * j + 1] = y2[j].y; }
            store_seg<T, S>(orow + t0, nvalid, vec_io, y);
            if constexpr (kHasZ) {
                float zv[S];
                load_seg<T, S>(zrow + t0, nvalid, vec_io, zv);
#pragma unroll
                for (int i = 0; i < S; ++i) y[i] = y[i] * zv[i] * sigmoid_f(zv[i]);
                store_seg<T, S>(ozrow + t0, nvalid, vec_io, y);
            }
        }
        __syncthreads();  // all warps done with this stage before it is refilled (chunk c+2)
    }
}

// ---------------------------------------------------------------------------------------------
// host launcher
// ---------------------------------------------------------------------------------------------
template <int S, int G, int NW>
constexpr size_t fwd_smem_bytes(int dstate) {
    return sizeof(float) * (4 * (size_t)dstate * G * seg_pad(S) + 2 * (size_t)NW * (32 / G) * dstate);
}

template <typename T, int S, int G, int NW, int NS, int MINB>
static cudaError_t launch_cfg(const FmScanFwdParams& p, cudaStream_t st, int vec_io, int vec_bc) {
    constexpr int RW = 32 / G, R = NW * RW;
    const int dg = p.dim / p.n_groups;
    const int tiles = (dg + R - 1) / R;
    dim3 grid(tiles * p.n_groups, p.batch);
    const size_t smem = fwd_smem_bytes<S, G, NW>(p.dstate);
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    auto kern = p.z ? scan_fwd_kernel<T, S, G, NW, NS, true, MINB> : scan_fwd_kernel<T, S, G, NW, NS, false, MINB>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, NW * 32, smem, st>>>(p, vec_io, vec_bc);
    count_launch();
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_scan_fwd_T(const FmScanFwdParams& p, cudaStream_t st) {
    const int es = sizeof(T);
    const int64_t al = 16 / es;  // elements per 16 bytes
    auto ok = [&](const void* ptr, int64_t s0, int64_t s1) { return aligned16(ptr) && s0 % al == 0 && s1 % al == 0; };
    int vec_io = ok(p.u, p.u_batch_stride, p.u_d_stride) && ok(p.delta, p.delta_batch_stride, p.delta_d_stride) &&
                 (p.out_map != FM_MAP_LINEAR || ok(p.out, p.out_batch_stride, p.out_d_stride));   // fused merge stores are scalar
    if (p.z) vec_io = vec_io && ok(p.z, p.z_batch_stride, p.z_d_stride) && ok(p.out_z, p.out_z_batch_stride, p.out_z_d_stride);
    int vec_bc = ok(p.B, p.B_batch_stride, p.B_group_stride) && p.B_dstate_stride % al == 0 &&
                 ok(p.C, p.C_batch_stride, p.C_group_stride) && p.C_dstate_stride % al == 0;

    // dstate == 16 (the only state size FusionMamba uses): lane-serial single-pass kernel (fm_scan_fwd16.cuh)
    if (p.dstate == 16 && env_int("FM_SCAN_FWD16", 1) != 0) {
        const cudaError_t e16 = launch_scan_fwd16_T<T>(p, st, vec_io, vec_bc);
        if (e16 != cudaErrorInvalidConfiguration) return e16;   // no instance for this shape: use the generic kernel
    }
    if (p.out_map != FM_MAP_LINEAR) return cudaErrorInvalidConfiguration;   // only the dstate-16 kernel fuses the merge
    // default: row-pair kernel (fm_scan_fwd_rp.cuh)
    if (env_int("FM_SCAN_FWD_RP", 1) != 0) return launch_scan_fwd_rp_T<T>(p, st, vec_io, vec_bc);

    // Launch shape.  S = 8 steps per lane keeps the kernel at <= 64 registers (8 warps per scheduler);
    // G lanes per row: enough lanes to fill 148 SMs, never more than the sequence can use.
    const int64_t rows = (int64_t)p.batch * p.dim;
    int S = env_int("FM_SCAN_FWD_S", 8);
    int G = scan_lanes_per_row(rows, p.seqlen, S, "FM_SCAN_FWD_G");
    int NW = env_int("FM_SCAN_FWD_NW", 8);
    const int NS = (p.dstate % 2 == 0) ? 2 : 1;
    // shared-memory budget: the double-buffered B/C tile is 4*dstate*G*(S+4) floats
    while (G > 1 && sizeof(float) * 4 * (size_t)p.dstate * G * (S + 4) > 96 * 1024) G >>= 1;

    const int MB = env_int("FM_SCAN_FWD_MINB", 0);   // tuning: alternative register budgets
#define FM_CASE(s, g, nw, minb, sel)                                                                \
    if (S == s && G == g && NW == nw && (sel)) {                                                    \
        return NS == 2 ? launch_cfg<T, s, g, nw, 2, minb>(p, st, vec_io, vec_bc)                    \
                       : launch_cfg<T, s, g, nw, 1, minb>(p, st, vec_io, vec_bc);                   \
    }
    FM_CASE(8, 16, 8, 2, MB == 2) FM_CASE(8, 16, 4, 4, MB == 4) FM_CASE(8, 16, 4, 5, MB == 5)
    FM_CASE(8, 1, 8, 3, true) FM_CASE(8, 2, 8, 3, true) FM_CASE(8, 4, 8, 3, true) FM_CASE(8, 8, 8, 3, true)
    FM_CASE(8, 16, 8, 3, true) FM_CASE(8, 32, 8, 2, true)
    FM_CASE(8, 16, 4, 6, true) FM_CASE(8, 32, 4, 4, true)
    FM_CASE(16, 16, 8, 2, true)
#undef FM_CASE
    return cudaErrorInvalidConfiguration;
}

}  // namespace fm
