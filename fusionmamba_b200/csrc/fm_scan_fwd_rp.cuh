// fm_scan_fwd_rp.cuh -- selective-scan forward for sm_100a, "row-pair" kernel (any dstate).
//
// Replaces selective_scan_fwd_kernel (selective_scan/selective_scan_fwd_kernel.cuh:67-303).  Not a port:
//   * a lane owns TWO channel rows of one (batch, group) and S = 8 consecutive timesteps of a chunk; the two rows
//     ride the two halves of the Blackwell packed-fp32 pipe (FMUL2 / FFMA2): every B_t / C_t value read from
//     shared memory is a scalar-broadcast operand that serves both rows, so the shared-memory -> register traffic
//     per (t, row, state) is half of a one-row-per-lane layout and every recurrence step is one packed issue slot.
//   * a row pair is scanned by G lanes of a warp: thread-serial up-sweep from zero over the lane's S steps, G-lane
//     warp-shuffle combine of the (decay, state) aggregates with the monoid (a0,b0)o(a1,b1) = (a1*a0, a1*b0+b1),
//     then a down-sweep seeded with the lane's true incoming state that also accumulates y += C*h.  a_t is
//     computed once (one MUFU.EX2 per (t, row, state)) and stays in registers between the sweeps; a segment's
//     aggregate decay is exp2(A * sum(delta)).
//   * one CTA owns R = NW*(32/G)*2 rows and walks the sequence in chunks of TC = S*G steps; the [dstate x TC]
//     B and C tiles are staged once per chunk (cp.async double buffer) and shared by all rows.  The running
//     state is carried across chunks per (row pair, state) in shared memory by the seg-0 lane.
//   * u / delta / z / out move as 128-bit vector accesses.
#pragma once
#include "fm_common.cuh"
#include "fm_launch.h"

namespace fm {

__device__ __forceinline__ float2 shfl_up2(float2 v, int o, int w) {
    return make_float2(__shfl_up_sync(0xffffffffu, v.x, o, w), __shfl_up_sync(0xffffffffu, v.y, o, w));
}
__device__ __forceinline__ float2 shfl_idx2(float2 v, int i, int w) {
    return make_float2(__shfl_sync(0xffffffffu, v.x, i, w), __shfl_sync(0xffffffffu, v.y, i, w));
}

template <typename T, int G, int NW, bool kHasZ>
__global__ void __launch_bounds__(NW * 32)
scan_fwd_rp_kernel(const FmScanFwdParams p, const int vec_io, const int vec_bc) {
    constexpr int S = 8;
    constexpr int TC = G * S;               // timesteps per chunk
    constexpr int PW = 32 / G;              // row pairs per warp
    constexpr int RP = NW * PW;             // row pairs per CTA
    constexpr int R = 2 * RP;               // rows per CTA
    constexpr int SP = seg_pad(S);
    constexpr int ROWP = G * SP;            // smem pitch of one state row of the B/C tile (floats)
    constexpr int NT = NW * 32;

    const int N = p.dstate;
    const int L = p.seqlen;
    const int dg = p.dim / p.n_groups;
    const int tiles_per_group = (dg + R - 1) / R;
    const int group = blockIdx.x / tiles_per_group;
    const int tile = blockIdx.x % tiles_per_group;
    const int b = blockIdx.y;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int seg = lane % G;                           // which S-step segment of the chunk
    const int rp = warp * PW + lane / G;                // row pair within CTA
    const int dloc0 = tile * R + 2 * rp, dloc1 = dloc0 + 1;
    const bool ok0 = dloc0 < dg, ok1 = dloc1 < dg;      // invalid rows shadow row 0 of the group and never store
    const int d0 = group * dg + (ok0 ? dloc0 : 0), d1 = group * dg + (ok1 ? dloc1 : 0);

    extern __shared__ __align__(16) float smem[];
    float* sBC = smem;                                  // [2 stages][B|C][N][ROWP]
    float2* sA2 = reinterpret_cast<float2*>(sBC + 4 * N * ROWP);   // [RP][N]  A * log2(e) of both rows
    float2* sH = sA2 + RP * N;                          // [RP][N]  running state (touched only by the seg==0 lane)

    const T* __restrict__ Bg = reinterpret_cast<const T*>(p.B) + b * p.B_batch_stride + group * p.B_group_stride;
    const T* __restrict__ Cg = reinterpret_cast<const T*>(p.C) + b * p.C_batch_stride + group * p.C_group_stride;
    const T* __restrict__ ub = reinterpret_cast<const T*>(p.u) + b * p.u_batch_stride;
    const T* __restrict__ db = reinterpret_cast<const T*>(p.delta) + b * p.delta_batch_stride;
    T* __restrict__ ob = reinterpret_cast<T*>(p.out) + b * p.out_batch_stride;
    const int64_t rowid0 = static_cast<int64_t>(b) * p.dim + d0, rowid1 = static_cast<int64_t>(b) * p.dim + d1;
    float* __restrict__ xb = reinterpret_cast<float*>(p.x);
    float* __restrict__ hb = reinterpret_cast<float*>(p.hck);

    const float2 Dv = p.D ? make_float2(reinterpret_cast<const float*>(p.D)[d0], reinterpret_cast<const float*>(p.D)[d1])
                          : make_float2(0.f, 0.f);
    const float2 bias = p.delta_bias ? make_float2(reinterpret_cast<const float*>(p.delta_bias)[d0],
                                                   reinterpret_cast<const float*>(p.delta_bias)[d1])
                                     : make_float2(0.f, 0.f);

    for (int i = tid; i < RP * N; i += NT) {
        const int r = i / N, n = i % N;
        const int l0 = tile * R + 2 * r, l1 = l0 + 1;
        const int e0 = group * dg + (l0 < dg ? l0 : 0), e1 = group * dg + (l1 < dg ? l1 : 0);
        const float* Ap = reinterpret_cast<const float*>(p.A);
        sA2[i] = make_float2(Ap[e0 * p.A_d_stride + n * p.A_dstate_stride] * kLog2e,
                             Ap[e1 * p.A_d_stride + n * p.A_dstate_stride] * kLog2e);
        sH[i] = make_float2(0.f, 0.f);
    }

    const int n_chunks = (L + TC - 1) / TC;
    stage_tile<T, TC, S>(sBC, Bg, p.B_dstate_stride, N, 0, L, vec_bc, tid, NT);
    stage_tile<T, TC, S>(sBC + N * ROWP, Cg, p.C_dstate_stride, N, 0, L, vec_bc, tid, NT);
    cp_async_commit();

    float2 sum_lane = make_float2(0.f, 0.f);   // sum of this lane's delta over all chunks so far (x's decay product)
    const float2* myA2 = sA2 + rp * N;
    float2* myH = sH + rp * N;

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
        float2 dl2[S], du2[S], y2[S];                   // (row0, row1) per timestep
        {
            float u0[S], u1[S], e0[S], e1[S];
            load_seg<T, S>(ub + d0 * p.u_d_stride + t0, nvalid, vec_io, u0);
            load_seg<T, S>(ub + d1 * p.u_d_stride + t0, nvalid, vec_io, u1);
            load_seg<T, S>(db + d0 * p.delta_d_stride + t0, nvalid, vec_io, e0);
            load_seg<T, S>(db + d1 * p.delta_d_stride + t0, nvalid, vec_io, e1);
#pragma unroll
            for (int i = 0; i < S; ++i) {
                const float x0 = e0[i] + bias.x, x1 = e1[i] + bias.y;
                const float s0 = p.delta_softplus ? softplus_fast(x0) : x0;
                const float s1 = p.delta_softplus ? softplus_fast(x1) : x1;
                dl2[i] = (i < nvalid) ? make_float2(s0, s1) : make_float2(0.f, 0.f);   // masked steps: a = 1, b = 0
                const float2 uu = make_float2(u0[i], u1[i]);
                du2[i] = mul2(dl2[i], uu);
                y2[i] = mul2(uu, Dv);
            }
        }
        float2 sumd2 = dl2[0];
#pragma unroll
        for (int i = 1; i < S; ++i) sumd2 = add2(sumd2, dl2[i]);
        sum_lane = add2(sum_lane, sumd2);

        // per-chunk bookkeeping, hoisted out of the state loop
        const int t_end = min((c + 1) * TC, L);         // exclusive end of this chunk
        const bool xwrite = (t_end % p.chunk_len == 0) || t_end == L;    // CTA-uniform
        float2 sum_row = make_float2(0.f, 0.f);
        if (xwrite) {
            sum_row = sum_lane;
#pragma unroll
            for (int o = 1; o < G; o <<= 1) {
                sum_row.x += __shfl_xor_sync(0xffffffffu, sum_row.x, o, G);
                sum_row.y += __shfl_xor_sync(0xffffffffu, sum_row.y, o, G);
            }
        }
        const int xoff = ((t_end - 1) / p.chunk_len) * 2 * N;
        int hoff = -1;                                  // dense checkpoint: state at the end of this lane's segment
        {
            const int te = t0 + S;
            if (hb != nullptr && te < L && te % p.hck_len == 0) hoff = (te / p.hck_len - 1) * N;
        }

        const float* tB = sBC + stage * 2 * N * ROWP + seg * SP;
        const float* tC = tB + N * ROWP;

#pragma unroll 1
        for (int n = 0; n < N; ++n) {
            const float2 A2 = myA2[n];
            const float2 hrun = myH[n];
            float2 a2[S], b2[S];
            {
                const float4 v0 = lds128(tB + n * ROWP), v1 = lds128(tB + n * ROWP + 4);
                const float bv[S] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
                for (int j = 0; j < S; ++j) {
                    b2[j] = mul2(du2[j], bcast2(bv[j]));
                    const float2 x2 = mul2(dl2[j], A2);
                    a2[j] = make_float2(ex2_approx(x2.x), ex2_approx(x2.y));
                }
            }
            const float2 ps = mul2(A2, sumd2);
            float2 P = make_float2(ex2_approx(ps.x), ex2_approx(ps.y));
            // up-sweep: segment state from zero
            float2 h = b2[0];
#pragma unroll
            for (int j = 1; j < S; ++j) h = fma2(a2[j], h, b2[j]);
            if (seg == 0) h = fma2(P, hrun, h);
            // inclusive combine over the G lanes of the row pair
#pragma unroll
            for (int o = 1; o < G; o <<= 1) {
                const float2 hp = shfl_up2(h, o, G);
                float2 Pp = make_float2(1.f, 1.f);
                if (2 * o < G) Pp = shfl_up2(P, o, G);
                if (seg >= o) {
                    h = fma2(P, hp, h);
                    if (2 * o < G) P = mul2(P, Pp);
                }
            }
            float2 hin = shfl_up2(h, 1, G);
            const float2 hlast = shfl_idx2(h, G - 1, G);
            if (hoff >= 0) {
                if (ok0) hb[rowid0 * p.n_hck * N + hoff + n] = h.x;
                if (ok1) hb[rowid1 * p.n_hck * N + hoff + n] = h.y;
            }
            if (seg == 0) {
                hin = hrun;
                myH[n] = hlast;
                if (xwrite) {
                    const float2 q = mul2(A2, sum_row);
                    if (ok0) *reinterpret_cast<float2*>(xb + rowid0 * p.n_chunks * 2 * N + xoff + 2 * n) = make_float2(ex2_approx(q.x), hlast.x);
                    if (ok1) *reinterpret_cast<float2*>(xb + rowid1 * p.n_chunks * 2 * N + xoff + 2 * n) = make_float2(ex2_approx(q.y), hlast.y);
                }
            }
            // down-sweep with the true incoming state; y += C * h
            {
                const float4 v0 = lds128(tC + n * ROWP), v1 = lds128(tC + n * ROWP + 4);
                const float cv[S] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
                for (int j = 0; j < S; ++j) {
                    hin = fma2(a2[j], hin, b2[j]);
                    y2[j] = fma2(bcast2(cv[j]), hin, y2[j]);
                }
            }
        }

        if (nvalid > 0) {
            float y0[S], y1[S];
#pragma unroll
            for (int j = 0; j < S; ++j) { y0[j] = y2[j].x; y1[j] = y2[j].y; }
            if (ok0) store_seg<T, S>(ob + d0 * p.out_d_stride + t0, nvalid, vec_io, y0);
            if (ok1) store_seg<T, S>(ob + d1 * p.out_d_stride + t0, nvalid, vec_io, y1);
            if constexpr (kHasZ) {
                const T* zb = reinterpret_cast<const T*>(p.z) + b * p.z_batch_stride;
                T* ozb = reinterpret_cast<T*>(p.out_z) + b * p.out_z_batch_stride;
                float zv[S];
                load_seg<T, S>(zb + d0 * p.z_d_stride + t0, nvalid, vec_io, zv);
#pragma unroll
                for (int i = 0; i < S; ++i) y0[i] = y0[i] * zv[i] * sigmoid_f(zv[i]);
                if (ok0) store_seg<T, S>(ozb + d0 * p.out_z_d_stride + t0, nvalid, vec_io, y0);
                load_seg<T, S>(zb + d1 * p.z_d_stride + t0, nvalid, vec_io, zv);
#pragma unroll
                for (int i = 0; i < S; ++i) y1[i] = y1[i] * zv[i] * sigmoid_f(zv[i]);
                if (ok1) store_seg<T, S>(ozb + d1 * p.out_z_d_stride + t0, nvalid, vec_io, y1);
            }
        }
        __syncthreads();  // all warps done with this stage before it is refilled (chunk c+2)
    }
}

template <int G, int NW>
constexpr size_t fwd_rp_smem_bytes(int dstate) {
    return sizeof(float) * (4 * (size_t)dstate * G * seg_pad(8) + 4 * (size_t)NW * (32 / G) * dstate);
}

template <typename T, int G, int NW>
static cudaError_t launch_fwd_rp_cfg(const FmScanFwdParams& p, cudaStream_t st, int vec_io, int vec_bc) {
    constexpr int R = 2 * NW * (32 / G);
    const int dg = p.dim / p.n_groups;
    const int tiles = (dg + R - 1) / R;
    dim3 grid(tiles * p.n_groups, p.batch);
    const size_t smem = fwd_rp_smem_bytes<G, NW>(p.dstate);
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    auto kern = p.z ? scan_fwd_rp_kernel<T, G, NW, true> : scan_fwd_rp_kernel<T, G, NW, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, NW * 32, smem, st>>>(p, vec_io, vec_bc);
    count_launch();
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_scan_fwd_rp_T(const FmScanFwdParams& p, cudaStream_t st, int vec_io, int vec_bc) {
    constexpr int S = 8;
    const int64_t pairs = ((int64_t)p.batch * p.dim + 1) / 2;
    int G = scan_lanes_per_row(pairs, p.seqlen, S, "FM_SCAN_FWD_G");
    int NW = env_int("FM_SCAN_FWD_NW", 0);
    // shared-memory budget: the double-buffered B/C tile is 4*dstate*G*(S+4) floats
    while (G > 1 && sizeof(float) * 4 * (size_t)p.dstate * G * (S + 4) > 96 * 1024) G >>= 1;
    if (NW != 1 && NW != 2 && NW != 4) {
        // rows per CTA = 2*NW*32/G: prefer a divisor of the channels per group, and enough CTAs for 148 SMs
        const int dg = p.dim / p.n_groups;
        NW = 4;
        while (NW > 1 && (dg % (2 * NW * (32 / G)) != 0)) NW >>= 1;
        while (NW > 1 && (int64_t)p.batch * p.n_groups * ((dg + 2 * NW * (32 / G) - 1) / (2 * NW * (32 / G))) < 2 * 148) NW >>= 1;
    }
    // the per-row-pair state / decay arrays grow with NW * (32 / G) * dstate: wide states with few lanes per row (dstate >= 128,
    // large batch) would not fit 227 KB at NW = 4 -- take fewer warps per CTA before giving up
    auto smem_need = [&](int g, int nw) { return sizeof(float) * (4 * (size_t)p.dstate * g * seg_pad(8) + 4 * (size_t)nw * (32 / g) * p.dstate); };
    while (NW > 1 && smem_need(G, NW) > 227 * 1024) NW >>= 1;
    while (G < 32 && smem_need(G, NW) > 227 * 1024) G <<= 1;     // more lanes per row = fewer rows per CTA
#define FM_CASE_RP(g, nw) if (G == g && NW == nw) return launch_fwd_rp_cfg<T, g, nw>(p, st, vec_io, vec_bc);
    FM_CASE_RP(1, 1) FM_CASE_RP(1, 2) FM_CASE_RP(1, 4)
    FM_CASE_RP(2, 1) FM_CASE_RP(2, 2) FM_CASE_RP(2, 4)
    FM_CASE_RP(4, 1) FM_CASE_RP(4, 2) FM_CASE_RP(4, 4)
    FM_CASE_RP(8, 1) FM_CASE_RP(8, 2) FM_CASE_RP(8, 4)
    FM_CASE_RP(16, 1) FM_CASE_RP(16, 2) FM_CASE_RP(16, 4)
    FM_CASE_RP(32, 1) FM_CASE_RP(32, 2) FM_CASE_RP(32, 4)
#undef FM_CASE_RP
    return cudaErrorInvalidConfiguration;
}

}  // namespace fm
