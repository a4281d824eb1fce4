// fm_scan_bwd_ls.cuh -- selective-scan backward for sm_100a, "lane-serial" kernel (dstate == 16, no z).
//
// Replaces selective_scan_bwd_kernel (selective_scan/selective_scan_bwd_kernel.cuh:75-489); same math (SURVEY.md section
// 3.5), a third decomposition next to fm_scan_bwd_rp.cuh (time-parallel row pairs) and fm_scan_bwd.cuh (generic):
//   * one WARP is an independent work unit: 8 channel rows of one (batch, group), walked in 8-step sub-chunks from the end of
//     the sequence to its start.  No block-level barrier, no cross-lane scan combine: the adjoint recurrence
//     dh_t = C_t dy_t + a_{t+1} dh_{t+1} is carried in registers along the walk, exactly like the forward state in
//     fm_scan_fwd16.cuh.
//   * a lane owns TWO rows (the halves of a packed fp32 pair) and TWO of the 16 states: every B_t / C_t value fetched from
//     shared memory is a scalar-broadcast operand of FMUL2 / FFMA2 serving both rows.  Lane (g = lane/8, s = lane%8) holds rows
//     2g, 2g+1 and states 2s, 2s+1.
//   * the forward states of a sub-chunk are rebuilt from the DENSE checkpoints the forward kernel leaves every 8 timesteps
//     (FmScanFwdParams.hck, hck_len == 8): 8 steps of h = a h + b into registers (a_t and h_t kept: 64 registers), then the
//     same 8 steps are walked backwards.  One MUFU.EX2 per (t, row, state), 12 packed fp32 operations per (t, row pair, state).
//     The checkpoints cost E*N/8*4 bytes of HBM traffic in each direction (201 MB at BASELINE configs[1]); HBM is the idle
//     resource of this kernel, the shared-memory -> register path and the issue slots are not.
//   * per-(row, t) operands (softplus(delta + bias), delta*u, dout) are evaluated once by a "staging" role -- lane i owns row
//     i/4, timesteps 2(i%4), 2(i%4)+1 of the sub-chunk, loads them 64 bits at a time, keeps them for the epilogue -- and
//     broadcast through a 1 KB shared tile; the next sub-chunk's global loads are issued before the current one is computed.
//   * du_t / ddelta_t need the sum over the 16 states = over the 8 lanes of a row pair: per-lane partials go through a
//     padded shared tile and come back to the staging lane of that (row, t), which already holds delta, u, dout and
//     stores du / ddelta 64 bits at a time.
//   * dB_t / dC_t need the sum over the rows of the group: summed over the lane's two rows in registers, then over the 4 row
//     pairs of the warp by a two-round reduce-scatter butterfly (24 shuffles per 8 steps), and leave the warp as two
//     red.global.add.v4.f32 per lane.  dA / dD / ddelta_bias: registers over the whole row, one atomic per (row, state).
#pragma once
#include "fm_common.cuh"
#include "fm_launch.h"

#ifndef FM_LS_DIAG
#define FM_LS_DIAG 0      // diagnostic builds only (tools/ubench/ls_diag.cu): bit 0 no dB/dC shuffles, 1 no partial tile, 2 no MUFU
#endif

namespace fm {

namespace ls {
constexpr int S = 8;                       // timesteps per sub-chunk (== hck_len)
constexpr int RT_RP = 20;                  // floats per row pair in a per-(row,t) tile: 8 t x 2 rows + 4 pad (bank spread)
constexpr int RT_ARR = 4 * RT_RP;          // one array (4 row pairs)
constexpr int BC_ARR = 128;                // [tq 2][parity 2][sg 8][4 t]
constexpr int PT_J = 20;                   // floats per timestep row of the partial tile: 8 lanes x (sB, sA) + 4 pad
constexpr int PT_ROW = 8 * PT_J + 4;       // one channel row (8 timesteps) + 4 pad
constexpr int PT_RP = 2 * PT_ROW + 8;      // one row pair; == 16 (mod 32) floats
constexpr int WARP_FLOATS = 3 * RT_ARR + 2 * BC_ARR + 4 * PT_RP;

template <typename T> struct Raw2 { using type = uint32_t; };
template <> struct Raw2<float> { using type = uint2; };
template <typename T> struct Raw8 { uint4 v[sizeof(T) == 4 ? 2 : 1]; };

template <typename T>
__device__ __forceinline__ void widen2(typename Raw2<T>::type r, float& a, float& b) {
    if constexpr (sizeof(T) == 4) {
        a = __uint_as_float(r.x); b = __uint_as_float(r.y);
    } else {
        const T* e = reinterpret_cast<const T*>(&r);
        a = Cvt<T>::to_f(e[0]); b = Cvt<T>::to_f(e[1]);
    }
}
template <typename T>
__device__ __forceinline__ typename Raw2<T>::type load2(const T* __restrict__ p, int nvalid, bool vec) {
    typename Raw2<T>::type r;
    if (vec && nvalid >= 2) {
        r = __ldg(reinterpret_cast<const typename Raw2<T>::type*>(p));
    } else {
        T e[2];
        e[0] = nvalid > 0 ? p[0] : Cvt<T>::from_f(0.f);
        e[1] = nvalid > 1 ? p[1] : Cvt<T>::from_f(0.f);
        r = *reinterpret_cast<const typename Raw2<T>::type*>(e);
    }
    return r;
}
template <typename T>
__device__ __forceinline__ void store2(T* __restrict__ p, int nvalid, bool vec, float a, float b) {
    if (nvalid <= 0) return;
    if (vec && nvalid >= 2) {
        T e[2] = {Cvt<T>::from_f(a), Cvt<T>::from_f(b)};
        *reinterpret_cast<typename Raw2<T>::type*>(p) = *reinterpret_cast<const typename Raw2<T>::type*>(e);
    } else {
        p[0] = Cvt<T>::from_f(a);
        if (nvalid > 1) p[1] = Cvt<T>::from_f(b);
    }
}
template <typename T>
__device__ __forceinline__ Raw8<T> load8(const T* __restrict__ p, int nvalid, bool vec) {
    Raw8<T> r;
    if (vec && nvalid >= 8) {
        const uint4* q = reinterpret_cast<const uint4*>(p);
        r.v[0] = __ldg(q);
        if constexpr (sizeof(T) == 4) r.v[1] = __ldg(q + 1);
    } else {
        T e[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) e[i] = (i < nvalid) ? p[i] : Cvt<T>::from_f(0.f);
        r.v[0] = reinterpret_cast<const uint4*>(e)[0];
        if constexpr (sizeof(T) == 4) r.v[1] = reinterpret_cast<const uint4*>(e)[1];
    }
    return r;
}
template <typename T>
__device__ __forceinline__ void widen8(const Raw8<T>& r, float (&f)[8]) {
    if constexpr (sizeof(T) == 4) {
        const float* e = reinterpret_cast<const float*>(r.v);
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = e[i];
    } else {
        const T* e = reinterpret_cast<const T*>(r.v);
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = Cvt<T>::to_f(e[i]);
    }
}
__device__ __forceinline__ void red_add_v4_ls(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
}  // namespace ls

template <typename T, int NW>
__global__ void __launch_bounds__(NW * 32, 1)
scan_bwd_ls_kernel(const FmScanBwdParams q, const int vec2_io, const int vec_bc, const int vec_dbc) {
    using namespace ls;
    const FmScanFwdParams& p = q.f;
    constexpr int N = 16;
    const int L = p.seqlen;
    const int dg = p.dim / p.n_groups;
    const int tiles = dg >> 3;                               // launcher guarantees dg % 8 == 0
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t unit = static_cast<int64_t>(blockIdx.x) * NW + warp;
    const int64_t n_units = static_cast<int64_t>(p.batch) * p.n_groups * tiles;
    if (unit >= n_units) return;                             // whole warp leaves; there is no block-level barrier in this kernel
    const int tile = static_cast<int>(unit % tiles);
    const int group = static_cast<int>((unit / tiles) % p.n_groups);
    const int b = static_cast<int>(unit / (static_cast<int64_t>(tiles) * p.n_groups));
    const int row0 = group * dg + tile * 8;

    extern __shared__ __align__(16) float smem_ls[];
    float* sRT = smem_ls + warp * WARP_FLOATS;               // [delta | delta*u | dy][row pair][t][row]
    float* sBC = sRT + 3 * RT_ARR;                           // [B | C][tq][state parity][sg][4 t]
    float* sPT = sBC + 2 * BC_ARR;                           // [row pair][row][t][sg] x (sB, sA)

    // ---- compute role -------------------------------------------------------------------------------------------------
    const int rp = lane >> 3, sg = lane & 7;
    const int d0 = row0 + 2 * rp, d1 = d0 + 1;
    float2 A2[2];                                            // A * log2(e) of (row d0, row d1), states 2sg and 2sg+1
    {
        const float* Ap = reinterpret_cast<const float*>(p.A);
#pragma unroll
        for (int s = 0; s < 2; ++s)
            A2[s] = make_float2(Ap[d0 * p.A_d_stride + (2 * sg + s) * p.A_dstate_stride] * kLog2e,
                                Ap[d1 * p.A_d_stride + (2 * sg + s) * p.A_dstate_stride] * kLog2e);
    }
    const float* __restrict__ hck0 =
        p.hck ? reinterpret_cast<const float*>(p.hck) + (static_cast<int64_t>(b) * p.dim + d0) * p.n_hck * N + 2 * sg : nullptr;
    const float* __restrict__ hck1 = p.hck ? hck0 + static_cast<int64_t>(p.n_hck) * N : nullptr;
    // final owner of the warp-reduced dB / dC values: quantity rp&1 (0: dB, 1: dC), timesteps 4*(rp>>1) .. +3, states 2sg, 2sg+1
    const int fin_q = rp & 1, fin_h = rp >> 1;
    float* __restrict__ dbc0 = (fin_q ? q.dC + b * q.dC_batch_stride + group * q.dC_group_stride + (2 * sg) * q.dC_dstate_stride
                                      : q.dB + b * q.dB_batch_stride + group * q.dB_group_stride + (2 * sg) * q.dB_dstate_stride) +
                               4 * fin_h;
    const int64_t dbc_ns = fin_q ? q.dC_dstate_stride : q.dB_dstate_stride;

    // ---- staging / epilogue role: lane -> (row, two timesteps) ---------------------------------------------------------------
    const int srow = lane >> 2, tp = lane & 3;
    const int ds = row0 + srow;
    const T* __restrict__ us = reinterpret_cast<const T*>(p.u) + b * p.u_batch_stride + ds * p.u_d_stride + 2 * tp;
    const T* __restrict__ es = reinterpret_cast<const T*>(p.delta) + b * p.delta_batch_stride + ds * p.delta_d_stride + 2 * tp;
    const T* __restrict__ gs = reinterpret_cast<const T*>(q.dout) + b * q.dout_batch_stride + ds * q.dout_d_stride + 2 * tp;
    T* __restrict__ dus = reinterpret_cast<T*>(q.du) + b * q.du_batch_stride + ds * q.du_d_stride + 2 * tp;
    T* __restrict__ dds = reinterpret_cast<T*>(q.ddelta) + b * q.ddelta_batch_stride + ds * q.ddelta_d_stride + 2 * tp;
    const float Dv = p.D ? reinterpret_cast<const float*>(p.D)[ds] : 0.f;
    const float bias = p.delta_bias ? reinterpret_cast<const float*>(p.delta_bias)[ds] : 0.f;
    const int rt_w = (srow >> 1) * RT_RP + 4 * tp + (srow & 1);              // + 2*i for timestep i of the lane's pair
    const int pt_r = (srow >> 1) * PT_RP + (srow & 1) * PT_ROW + 2 * tp * PT_J;   // + i*PT_J
    // B / C loader: lane -> (B or C, state n), 8 timesteps
    const int bc_which = lane >> 4, bc_n = lane & 15;
    const T* __restrict__ bcs = bc_which
        ? reinterpret_cast<const T*>(p.C) + b * p.C_batch_stride + group * p.C_group_stride + bc_n * p.C_dstate_stride
        : reinterpret_cast<const T*>(p.B) + b * p.B_batch_stride + group * p.B_group_stride + bc_n * p.B_dstate_stride;
    const int bc_w = bc_which * BC_ARR + (bc_n & 1) * 32 + (bc_n >> 1) * 4;   // + tq*64

    // ---- loop-carried state ------------------------------------------------------------------------------------------------
    float2 dh[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};           // adjoint state entering from the later sub-chunk
    float2 an[2] = {make_float2(1.f, 1.f), make_float2(1.f, 1.f)};           // a of the first timestep of the later sub-chunk
    float2 dA2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
    float dD_acc = 0.f, dbias_acc = 0.f;

    typename Raw2<T>::type pe, pu, pg;                                       // prefetched raw delta / u / dout of the staging pair
    Raw8<T> pbc;
    float2 ph0, ph1;
    auto prefetch = [&](int c) {
        const int t0 = c * S;
        const int nv = L - (t0 + 2 * tp);
        pe = load2<T>(es + t0, nv, vec2_io);
        pu = load2<T>(us + t0, nv, vec2_io);
        pg = load2<T>(gs + t0, nv, vec2_io);
        pbc = load8<T>(bcs + t0, L - t0, vec_bc);
        if (c > 0) {
            ph0 = __ldg(reinterpret_cast<const float2*>(hck0 + static_cast<int64_t>(c - 1) * N));
            ph1 = __ldg(reinterpret_cast<const float2*>(hck1 + static_cast<int64_t>(c - 1) * N));
        } else {
            ph0 = make_float2(0.f, 0.f); ph1 = make_float2(0.f, 0.f);
        }
    };

    const int n_sub = (L + S - 1) / S;
    prefetch(n_sub - 1);

    for (int c = n_sub - 1; c >= 0; --c) {
        const int t0 = c * S;
        // ---- stage: per-(row, t) operands of the lane's pair, B / C packets, checkpointed state --------------------------------
        float sdl[2], su[2], sdy[2];
        {
            float e0, e1;
            widen2<T>(pe, e0, e1); widen2<T>(pu, su[0], su[1]); widen2<T>(pg, sdy[0], sdy[1]);
            const float ee[2] = {e0, e1};
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const float x = ee[i] + bias;
                const float sp = p.delta_softplus ? softplus_fast(x) : x;
                const bool in = t0 + 2 * tp + i < L;
                sdl[i] = in ? sp : 0.f;                       // masked steps: a = 1, b = 0, dy = 0 -> they change nothing
                if (!in) { su[i] = 0.f; sdy[i] = 0.f; }
                dD_acc = fmaf(sdy[i], su[i], dD_acc);
                sRT[rt_w + 2 * i] = sdl[i];
                sRT[RT_ARR + rt_w + 2 * i] = sdl[i] * su[i];
                sRT[2 * RT_ARR + rt_w + 2 * i] = sdy[i];
            }
            float f[8];
            widen8<T>(pbc, f);
            sts128(sBC + bc_w, make_float4(f[0], f[1], f[2], f[3]));
            sts128(sBC + bc_w + 64, make_float4(f[4], f[5], f[6], f[7]));
        }
        float2 h[2] = {make_float2(ph0.x, ph1.x), make_float2(ph0.y, ph1.y)};
        __syncwarp();
        if (c > 0) prefetch(c - 1);                           // in flight behind the whole sub-chunk

        // ---- operands of the sub-chunk into registers --------------------------------------------------------------------------
        float2 dl2[S], du2[S];
        float Bv[2][S];
#pragma unroll
        for (int k = 0; k < S / 2; ++k) {
            const float4 v = lds128(sRT + rp * RT_RP + 4 * k);
            dl2[2 * k] = make_float2(v.x, v.y); dl2[2 * k + 1] = make_float2(v.z, v.w);
            const float4 w = lds128(sRT + RT_ARR + rp * RT_RP + 4 * k);
            du2[2 * k] = make_float2(w.x, w.y); du2[2 * k + 1] = make_float2(w.z, w.w);
        }
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int tq = 0; tq < 2; ++tq) {
                const float4 v = lds128(sBC + tq * 64 + s * 32 + sg * 4);
                Bv[s][4 * tq] = v.x; Bv[s][4 * tq + 1] = v.y; Bv[s][4 * tq + 2] = v.z; Bv[s][4 * tq + 3] = v.w;
            }

        // ---- forward: rebuild a_t and h_t of the 8 steps ------------------------------------------------------------------------
        float2 a[S][2], hs[S][2];
#pragma unroll
        for (int j = 0; j < S; ++j)
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const float2 x2 = mul2(dl2[j], A2[s]);
                a[j][s] = (FM_LS_DIAG & 4) ? mul2(x2, x2) : make_float2(ex2_approx(x2.x), ex2_approx(x2.y));
                h[s] = fma2(a[j][s], h[s], mul2(du2[j], bcast2(Bv[s][j])));
                hs[j][s] = h[s];
            }

        // ---- adjoint: walk the 8 steps backwards ------------------------------------------------------------------------------------
        float kept[2][S];                                      // dB (rp even) or dC (rp odd) summed over row pairs {rp, rp^1}
        float4 dyv, cq[2];                                     // dout of two steps (both rows); C of four steps per state
#pragma unroll
        for (int j = S - 1; j >= 0; --j) {
            if (j & 1) dyv = lds128(sRT + 2 * RT_ARR + rp * RT_RP + 4 * (j >> 1));
            if ((j & 3) == 3) {
                cq[0] = lds128(sBC + BC_ARR + (j >> 2) * 64 + sg * 4);
                cq[1] = lds128(sBC + BC_ARR + (j >> 2) * 64 + 32 + sg * 4);
            }
            const float2 dy2 = (j & 1) ? make_float2(dyv.z, dyv.w) : make_float2(dyv.x, dyv.y);
            const float2 ndu = make_float2(-du2[j].x, -du2[j].y);
            float2 sB2, sA2;
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const float cv = (j & 3) == 0 ? cq[s].x : ((j & 3) == 1 ? cq[s].y : ((j & 3) == 2 ? cq[s].z : cq[s].w));
                dh[s] = fma2(an[s], dh[s], mul2(dy2, bcast2(cv)));                 // dh_t
                an[s] = a[j][s];
                const float2 tc = mul2(dy2, hs[j][s]);
                const float dCv = tc.x + tc.y;
                sB2 = (s == 0) ? mul2(dh[s], bcast2(Bv[s][j])) : fma2(dh[s], bcast2(Bv[s][j]), sB2);
                const float2 g = fma2(ndu, bcast2(Bv[s][j]), hs[j][s]);            // h_t - b_t = a_t h_{t-1}
                const float2 w = mul2(dh[s], g);
                sA2 = (s == 0) ? mul2(w, A2[s]) : fma2(w, A2[s], sA2);            // x log2(e); folded back in the epilogue
                dA2[s] = fma2(dl2[j], w, dA2[s]);
                const float2 tb = mul2(dh[s], du2[j]);
                const float dBv = tb.x + tb.y;
                // reduce-scatter round 1 (lanes 8 apart): row pairs {rp, rp^1}; even rp keeps dB, odd rp keeps dC
                const float send = fin_q ? dBv : dCv;
                const float keep = fin_q ? dCv : dBv;
                kept[s][j] = (FM_LS_DIAG & 1) ? keep + send : keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
            // state-sum partials of this timestep, one (sB, sA) pair per row
            if (!(FM_LS_DIAG & 2) || j == 0) {
                *reinterpret_cast<float2*>(sPT + rp * PT_RP + j * PT_J + 2 * sg) = make_float2(sB2.x, sA2.x);
                *reinterpret_cast<float2*>(sPT + rp * PT_RP + PT_ROW + j * PT_J + 2 * sg) = make_float2(sB2.y, sA2.y);
            } else {
                dA2[0] = add2(dA2[0], add2(sB2, sA2));
            }
        }

        // ---- dB / dC: reduce-scatter round 2 (lanes 16 apart) splits the timestep halves; two vector reds per lane ---------------
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            float fin[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const float send = fin_h ? kept[s][jj] : kept[s][jj + 4];
                const float keep = fin_h ? kept[s][jj + 4] : kept[s][jj];
                fin[jj] = (FM_LS_DIAG & 1) ? keep + send : keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
            float* dst = dbc0 + s * dbc_ns + t0;
            const int tb = t0 + 4 * fin_h;
            if (vec_dbc && tb + 4 <= L) {
                red_add_v4_ls(dst, fin[0], fin[1], fin[2], fin[3]);
            } else {
#pragma unroll
                for (int jj = 0; jj < 4; ++jj)
                    if (tb + jj < L) atomicAdd(dst + jj, fin[jj]);
            }
        }
        __syncwarp();

        // ---- epilogue of the staging pair: sum the 8 lanes' partials, du and ddelta leave 64 bits at a time ---------------------
        {
            float o_du[2], o_dd[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const float* src = sPT + pt_r + i * PT_J;
                float4 v = lds128(src);
                float sB = v.x + v.z, sA = v.y + v.w;
#pragma unroll
                for (int k = 1; k < ((FM_LS_DIAG & 2) ? 1 : 4); ++k) {
                    v = lds128(src + 4 * k);
                    sB += v.x + v.z; sA += v.y + v.w;
                }
                o_du[i] = fmaf(sdl[i], sB, Dv * sdy[i]);
                float g = fmaf(su[i], sB, sA * 0.6931471805599453f);                 // d(loss)/d(Delta_t)
                if (p.delta_softplus) g *= one_minus_exp_neg(sdl[i]);                // sigmoid(x) = 1 - exp(-softplus(x))
                if (!(t0 + 2 * tp + i < L)) g = 0.f;
                o_dd[i] = g;
                dbias_acc += g;
            }
            const int nv = L - (t0 + 2 * tp);
            store2<T>(dus + t0, nv, vec2_io, o_du[0], o_du[1]);
            store2<T>(dds + t0, nv, vec2_io, o_dd[0], o_dd[1]);
        }
        // the next iteration's shared stores follow its own __syncwarp-separated reads; sPT is rewritten only after that barrier
    }

    // ---- whole-row results ------------------------------------------------------------------------------------------------------
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        atomicAdd(q.dA + static_cast<int64_t>(d0) * N + 2 * sg + s, dA2[s].x);
        atomicAdd(q.dA + static_cast<int64_t>(d1) * N + 2 * sg + s, dA2[s].y);
    }
#pragma unroll
    for (int o = 1; o < 4; o <<= 1) {
        dD_acc += __shfl_xor_sync(0xffffffffu, dD_acc, o);
        dbias_acc += __shfl_xor_sync(0xffffffffu, dbias_acc, o);
    }
    if (tp == 0) {
        if (q.dD) atomicAdd(q.dD + ds, dD_acc);
        if (q.ddelta_bias) atomicAdd(q.ddelta_bias + ds, dbias_acc);
    }
}

// Preconditions (cudaErrorInvalidConfiguration otherwise -> the caller falls back to the row-pair / generic kernels):
// dstate == 16, no z, channels per group a multiple of 8, dense checkpoints every 8 steps (or the sequence fits one sub-chunk).
template <typename T>
cudaError_t launch_scan_bwd_ls_T(const FmScanBwdParams& q, cudaStream_t st, int vec_bc, int vec_dbc) {
    const FmScanFwdParams& p = q.f;
    const int dg = p.dim / p.n_groups;
    if (p.dstate != 16 || p.z != nullptr || dg % 8 != 0) return cudaErrorInvalidConfiguration;
    if (p.seqlen > ls::S && !(p.hck != nullptr && p.hck_len == ls::S)) return cudaErrorInvalidConfiguration;
    const int64_t a2 = 2;   // elements per 64-bit (fp32) / 32-bit (16-bit types) access
    auto ok2 = [&](const void* ptr, int64_t s0, int64_t s1) {
        return (reinterpret_cast<uintptr_t>(ptr) % (2 * sizeof(T)) == 0) && s0 % a2 == 0 && s1 % a2 == 0;
    };
    const int vec2_io = ok2(p.u, p.u_batch_stride, p.u_d_stride) && ok2(p.delta, p.delta_batch_stride, p.delta_d_stride) &&
                        ok2(q.dout, q.dout_batch_stride, q.dout_d_stride) && ok2(q.du, q.du_batch_stride, q.du_d_stride) &&
                        ok2(q.ddelta, q.ddelta_batch_stride, q.ddelta_d_stride);
    const int64_t units = static_cast<int64_t>(p.batch) * p.n_groups * (dg / 8);
    int NW = env_int("FM_SCAN_BWD_LS_NW", 1);
    if (NW != 1 && NW != 2 && NW != 4) NW = 1;
    const int64_t blocks = (units + NW - 1) / NW;
    if (blocks > 0x7fffffff) return cudaErrorInvalidConfiguration;
    const size_t smem = sizeof(float) * ls::WARP_FLOATS * NW;
    void (*kern)(const FmScanBwdParams, int, int, int) =
        NW == 1 ? scan_bwd_ls_kernel<T, 1> : (NW == 2 ? scan_bwd_ls_kernel<T, 2> : scan_bwd_ls_kernel<T, 4>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<static_cast<unsigned>(blocks), NW * 32, smem, st>>>(q, vec2_io, vec_bc, vec_dbc);
    count_launch();
    return cudaGetLastError();
}

}  // namespace fm
