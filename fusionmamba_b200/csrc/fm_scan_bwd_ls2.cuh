// fm_scan_bwd_ls2.cuh -- selective-scan backward for sm_100a, software-pipelined lane-serial kernel (dstate == 16, no z).
//
// Replaces selective_scan_bwd_kernel (selective_scan/selective_scan_bwd_kernel.cuh:75-489); same math.
// Second generation of fm_scan_bwd_ls.cuh, rebuilt around the two things ncu showed that kernel waiting on (issue slots spent on
// bookkeeping, and a lone warp stalled on its own dependent chains -- DESIGN.md section 4):
//   * work split as before: a WARP owns 8 channel rows of one (batch, group) and walks the sequence backwards in 8-step
//     sub-chunks from the dense 8-step checkpoints of the forward; lane (rp = lane/8, sg = lane%8) owns the row PAIR 2rp, 2rp+1
//     (the halves of a packed fp32 pair) and states 2sg, 2sg+1.  No block barrier, no scan combine.
//   * THREE independent instruction streams per loop trip, in one basic block, so that a lone warp has something to issue while
//     a chain waits:   epilogue of sub-chunk k+1  |  compute of sub-chunk k  |  staging of sub-chunk k-1
//     (+ the global loads of k-2).  The streams talk through per-warp shared-memory rings (4 / 2 deep); one __syncwarp per trip.
//   * adjoint carried as e_t = a_t dh_t:  dh_t = fma(dy_t, C_t, e_{t+1}),  e_t = a_t dh_t,  w_t = e_t h_{t-1}  -- w (the
//     gradient of the exponent) needs neither b_t nor a second product with a_t: 13 packed fp32 + 2 MUFU per (t, row pair, state).
//   * both cross-lane reductions go through shared memory as transposes (plain STS.128 / LDS.128 + packed adds), no shuffles
//     and no lane-dependent selects:  state sums (du, ddelta: over the 8 lanes of a row pair) are picked up by the lane that
//     staged that (row pair, t) and stores du / ddelta;  row sums (dB, dC: over the warp's 4 row pairs) by the lane that owns
//     that (dB | dC, state) row of the output.
//   * per-(row, t) work (softplus, sigmoid, delta*u, masks) is evaluated on row pairs with packed fp32, once, by lane (rp, t).
//   * the binding resource is the SM's L1 / shared-memory data pipe (one wavefront per clock), so global
//     accesses are shaped for it: B / C arrive as 32-step tiles fetched with 8 lanes per 128-byte row, one quarter of a tile per
//     trip while the previous tile is consumed (a lane per 32-byte row piece costs one wavefront per LANE); the warp-reduced
//     dB / dC of four sub-chunks wait in a 32-step shared tile and leave as coalesced red.global.add.v4.f32.
// Preconditions beyond the first kernel's: B, C, dB, dC rows 16-byte aligned, seqlen a whole number of 16-byte chunks.
#pragma once
#include "fm_common.cuh"
#include "fm_launch.h"
#include "fm_scan_bwd_ls.cuh"

namespace fm {

namespace ls2 {
constexpr int S = 8;                          // timesteps per sub-chunk (== hck_len)
constexpr int RT_RP = 36, RT_BUF = 4 * RT_RP; // [rp][t] (delta.x, delta.y, delta*u.x, delta*u.y); row-pair pitch 9 x 16 B (odd)
constexpr int DY_RP = 20, DY_BUF = 4 * DY_RP; // [rp][t] (dy.x, dy.y); pitch 5 x 16 B
constexpr int ET_RP = 36, ET_BUF = 4 * ET_RP; // [rp][t] (u.x, u.y, sigmoid.x, sigmoid.y): staging -> epilogue of the same lane
constexpr int GT = 32;                        // timesteps per B / C tile and per dB / dC flush (4 sub-chunks)
constexpr int BC_PAIR = 2 * GT + 4;           // [state pair][even state: 32 t | odd state: 32 t | pad 4]: pitch 17 x 16 B (odd)
constexpr int BC_ARR = 8 * BC_PAIR, BC_BUF = 2 * BC_ARR;     // (B | C) x 8 pairs
constexpr int DBW_ROW = GT + 4;               // [dB | dC][state][32 t | pad 4]: warp-reduced dB / dC waiting for the coalesced flush
constexpr int DBW_BUF = 32 * DBW_ROW;
constexpr int PT_BUF = 1024;                  // [rp][t][sg] (sB.x, sB.y, sA.x, sA.y)
constexpr int RR_BUF = 1024;                  // [rp][dB | dC][s][t half][sg] x 4 t
constexpr int OFF_RT = 0;
constexpr int OFF_DY = OFF_RT + 4 * RT_BUF;
constexpr int OFF_ET = OFF_DY + 4 * DY_BUF;
constexpr int OFF_BC = OFF_ET + 4 * ET_BUF;
constexpr int OFF_PT = OFF_BC + 2 * BC_BUF;
constexpr int OFF_RR = OFF_PT + 2 * PT_BUF;
constexpr int OFF_DBW = OFF_RR + 2 * RR_BUF;
constexpr int WARP_FLOATS = OFF_DBW + DBW_BUF;        // 8896 floats = 34.75 KB per warp (6 warps per SM fit 227 KB)

template <typename T> struct Raw1 { using type = unsigned short; };
template <> struct Raw1<float> { using type = float; };
template <typename T>
__device__ __forceinline__ float widen1(typename Raw1<T>::type r) {
    if constexpr (sizeof(T) == 4) return r;
    else return Cvt<T>::to_f(*reinterpret_cast<const T*>(&r));
}
template <typename T>
__device__ __forceinline__ typename Raw1<T>::type ld_raw(const T* __restrict__ p, bool in) {
    typename Raw1<T>::type r = 0;
    if (in) r = __ldg(reinterpret_cast<const typename Raw1<T>::type*>(p));
    return r;
}
__device__ __forceinline__ float2 lo2(float4 v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi2(float4 v) { return make_float2(v.z, v.w); }
__device__ __forceinline__ float comp(const float4& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }
}  // namespace ls2

template <typename T, int NW>
__global__ void __launch_bounds__(NW * 32, 1)
scan_bwd_ls2_kernel(const FmScanBwdParams q, const int vec_bc, const int vec_dbc) {
    using namespace ls2;
    using R1 = typename Raw1<T>::type;
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

    const int n_sub = (L + S - 1) / S;
    extern __shared__ __align__(16) float smem_ls2[];
    float* const sw = smem_ls2 + warp * WARP_FLOATS;

    // ---- compute role: lane -> (row pair rp, states 2sg, 2sg+1) -----------------------------------------------------------------
    const int rp = lane >> 3, sg = lane & 7;
    const int d0 = row0 + 2 * rp, d1 = d0 + 1;
    float2 A2[2];                                            // A * log2(e) of (row d0, row d1)
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
    const int c_rt = OFF_RT + rp * RT_RP, c_dy = OFF_DY + rp * DY_RP, c_bc = OFF_BC + sg * BC_PAIR;
    const int c_pt = OFF_PT + 4 * (rp * 64 + sg), c_rr = OFF_RR + 4 * (rp * 64 + sg);

    // ---- staging / epilogue role: lane -> (row pair rp, timestep te of the sub-chunk) ---------------------------------------------
    const int te = lane & 7;
    const T* __restrict__ us = reinterpret_cast<const T*>(p.u) + b * p.u_batch_stride + d0 * p.u_d_stride + te;
    const T* __restrict__ es = reinterpret_cast<const T*>(p.delta) + b * p.delta_batch_stride + d0 * p.delta_d_stride + te;
    const T* __restrict__ gs = reinterpret_cast<const T*>(q.dout) + b * q.dout_batch_stride + d0 * q.dout_d_stride + te;
    T* __restrict__ dus = reinterpret_cast<T*>(q.du) + b * q.du_batch_stride + d0 * q.du_d_stride + te;
    T* __restrict__ dds = reinterpret_cast<T*>(q.ddelta) + b * q.ddelta_batch_stride + d0 * q.ddelta_d_stride + te;
    const int64_t u_ds = p.u_d_stride, e_ds = p.delta_d_stride, g_ds = q.dout_d_stride, du_ds = q.du_d_stride, dd_ds = q.ddelta_d_stride;
    const float2 Dv2 = p.D ? make_float2(reinterpret_cast<const float*>(p.D)[d0], reinterpret_cast<const float*>(p.D)[d1])
                           : make_float2(0.f, 0.f);
    const float2 bias2 = p.delta_bias ? make_float2(reinterpret_cast<const float*>(p.delta_bias)[d0],
                                                    reinterpret_cast<const float*>(p.delta_bias)[d1])
                                      : make_float2(0.f, 0.f);
    const bool do_sp = p.delta_softplus != 0;
    const int s_rt = OFF_RT + rp * RT_RP + 4 * te, s_dy = OFF_DY + rp * DY_RP + 2 * te, s_et = OFF_ET + rp * ET_RP + 4 * te;
    const int e_pt = OFF_PT + 4 * (rp * 64 + te * 8);        // + 4 * (kk ^ te)
    // B / C loader: 32-step tiles [B | C][16 states][32 t] = 32 rows of CPR 16-byte chunks, fetched with CPR lanes per row (one
    // 128-byte line per row for fp32) -- a lane per row piece would cost one L1 wavefront per lane and instruction.  A tile is
    // fetched in 4 parts of 8 rows, one part per loop trip, while the previous tile is being consumed.
    constexpr int EPC = 16 / (int)sizeof(T), CPR = GT / EPC, CPL = CPR / 4;      // elements per chunk, chunks per row, chunks per lane and part
    const T* __restrict__ Bg = reinterpret_cast<const T*>(p.B) + b * p.B_batch_stride + group * p.B_group_stride;
    const T* __restrict__ Cg = reinterpret_cast<const T*>(p.C) + b * p.C_batch_stride + group * p.C_group_stride;
    const int64_t B_ns = p.B_dstate_stride, C_ns = p.C_dstate_stride;
    const int bc_c = lane % CPR, bc_r = lane / CPR;         // chunk column; row within the 32 / CPR rows one instruction covers
    const T* __restrict__ Bl = Bg + bc_r * B_ns + bc_c * EPC;       // the lane's chunk of row bc_r at t = 0
    const T* __restrict__ Cl = Cg + bc_r * C_ns + bc_c * EPC;
    float* const dBl = q.dB + b * q.dB_batch_stride + group * q.dB_group_stride + (lane >> 3) * q.dB_dstate_stride + 4 * (lane & 7);
    float* const dCl = q.dC + b * q.dC_batch_stride + group * q.dC_group_stride + (lane >> 3) * q.dC_dstate_stride + 4 * (lane & 7);
    const int64_t dB_4ns = 4 * q.dB_dstate_stride, dC_4ns = 4 * q.dC_dstate_stride;
    // owner of one (dB | dC, state) output row: fq = lane/16, state 2*(lane%8) + (lane/8)%2
    const int fq = lane >> 4, fs = (lane >> 3) & 1, fsg = lane & 7;
    const int f_rr = OFF_RR + 4 * ((fq * 2 + fs) * 16 + fsg);
    const int f_dbw = OFF_DBW + (fq * 16 + 2 * fsg + fs) * DBW_ROW;

    // ---- loop-carried state ------------------------------------------------------------------------------------------------
    float2 e2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};           // a_t dh_t of the first step of the later sub-chunk
    float2 dA2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
    float2 dD2 = make_float2(0.f, 0.f), dbias2 = make_float2(0.f, 0.f);
    R1 pe0, pe1, pu0, pu1, pg0, pg1;                                         // raw delta / u / dout of (rows d0, d1; timestep te)
    uint4 pbc[CPL];                                                          // raw chunks of the B / C tile part in flight
    float2 ph0 = make_float2(0.f, 0.f), ph1 = make_float2(0.f, 0.f);         // checkpointed state entering the next computed sub-chunk

    auto prefetch_raw = [&](int m) {                                         // m may be negative: everything masked
        const int t0 = m * S;
        const bool in = static_cast<unsigned>(t0 + te) < static_cast<unsigned>(L);
        pe0 = ld_raw<T>(es + t0, in); pe1 = ld_raw<T>(es + e_ds + t0, in);
        pu0 = ld_raw<T>(us + t0, in); pu1 = ld_raw<T>(us + u_ds + t0, in);
        pg0 = ld_raw<T>(gs + t0, in); pg1 = ld_raw<T>(gs + g_ds + t0, in);
    };
    // part `part` of the B / C tile of 32-step group G: rows 8 part .. 8 part + 7 (0-15: B states, 16-31: C states); chunks beyond
    // L (L is a multiple of the chunk length, checked by the launcher) and groups before the sequence read as zero
    auto bc_load = [&](int G, int part, uint4 (&r)[CPL]) {
        const int t = G * GT + bc_c * EPC;
        const bool in = G >= 0 && t < L;
        const T* src = ((part >> 1) ? Cl + (part & 1) * 8 * C_ns : Bl + (part & 1) * 8 * B_ns) + G * GT;
        const int64_t step = (32 / CPR) * ((part >> 1) ? C_ns : B_ns);
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
            r[i] = make_uint4(0u, 0u, 0u, 0u);
            if (in) r[i] = __ldg(reinterpret_cast<const uint4*>(src + i * step));
        }
    };
    auto bc_store = [&](int G, int part, const uint4 (&r)[CPL]) {
        float* const tile = sw + OFF_BC + (G & 1) * BC_BUF + (part >> 1) * BC_ARR;
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
            const int n = (part & 1) * 8 + i * (32 / CPR) + bc_r;
            float* dst = tile + (n >> 1) * BC_PAIR + (n & 1) * GT + bc_c * EPC;
            if constexpr (sizeof(T) == 4) {
                *reinterpret_cast<uint4*>(dst) = r[i];
            } else {
                const T* e = reinterpret_cast<const T*>(&r[i]);
                sts128(dst, make_float4(Cvt<T>::to_f(e[0]), Cvt<T>::to_f(e[1]), Cvt<T>::to_f(e[2]), Cvt<T>::to_f(e[3])));
                sts128(dst + 4, make_float4(Cvt<T>::to_f(e[4]), Cvt<T>::to_f(e[5]), Cvt<T>::to_f(e[6]), Cvt<T>::to_f(e[7])));
            }
        }
    };
    // coalesced flush of the warp-reduced dB / dC of group G: 8 lanes per (dB | dC, state) row, one 128-byte line per row
    auto dbc_flush = [&](int G) {
        const bool in = G * GT + 4 * (lane & 7) < L;
        const float* src = sw + OFF_DBW + (lane >> 3) * DBW_ROW + 4 * (lane & 7);
        float* pb = dBl + G * GT;
        float* pc = dCl + G * GT;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 vb = lds128(src + 4 * i * DBW_ROW), vc = lds128(src + (16 + 4 * i) * DBW_ROW);
            if (in) {
                ls::red_add_v4_ls(pb, vb.x, vb.y, vb.z, vb.w);
                ls::red_add_v4_ls(pc, vc.x, vc.y, vc.z, vc.w);
            }
            pb += dB_4ns; pc += dC_4ns;
        }
    };
    auto prefetch_hck = [&](int m) {                                         // state entering sub-chunk m (m <= 0: zero)
        ph0 = make_float2(0.f, 0.f); ph1 = make_float2(0.f, 0.f);
        if (m > 0) {
            ph0 = __ldg(reinterpret_cast<const float2*>(hck0 + static_cast<int64_t>(m - 1) * N));
            ph1 = __ldg(reinterpret_cast<const float2*>(hck1 + static_cast<int64_t>(m - 1) * N));
        }
    };

    // ---- staging of sub-chunk m from the raw registers: per-(row pair, t) operands, B / C piece -----------------------------------------
    auto stage = [&](int m, float2 x, float2 u2, float2 g2) {              // x = delta + bias; masked loads read 0
        const bool in = static_cast<unsigned>(m * S + te) < static_cast<unsigned>(L);
        float2 dl, sig;
        {   // softplus and its derivative on the row pair, branch-free (selected against the identity when delta_softplus is off)
            const float2 t = make_float2(exp_neg_abs(x.x), exp_neg_abs(x.y));
            const float2 tp2 = add2(t, bcast2(2.f)), tp1 = add2(t, bcast2(1.f));
            const float2 s = mul2(t, make_float2(rcp_approx(tp2.x), rcp_approx(tp2.y)));
            const float2 r1 = make_float2(rcp_approx(tp1.x), rcp_approx(tp1.y));
            const float2 s2 = mul2(s, s);
            float2 pl = fma2(s2, bcast2(1.f / 13.f), bcast2(1.f / 11.f));
            pl = fma2(s2, pl, bcast2(1.f / 9.f));
            pl = fma2(s2, pl, bcast2(1.f / 7.f));
            pl = fma2(s2, pl, bcast2(1.f / 5.f));
            pl = fma2(s2, pl, bcast2(1.f / 3.f));
            pl = fma2(s2, pl, bcast2(1.f));
            const float2 sp = fma2(add2(s, s), pl, make_float2(fmaxf(x.x, 0.f), fmaxf(x.y, 0.f)));
            const float2 tr = mul2(t, r1);
            dl = make_float2(do_sp ? sp.x : x.x, do_sp ? sp.y : x.y);
            sig = make_float2(do_sp ? (x.x >= 0.f ? r1.x : tr.x) : 1.f, do_sp ? (x.y >= 0.f ? r1.y : tr.y) : 1.f);
        }
        if (!in) { dl = make_float2(0.f, 0.f); sig = make_float2(0.f, 0.f); }   // masked steps: a = 1, b = 0, dy = 0 -> they change nothing
        const float2 dlu = mul2(dl, u2);
        dD2 = fma2(g2, u2, dD2);
        float* const rt = sw + (m & 3) * RT_BUF;
        sts128(rt + s_rt, make_float4(dl.x, dl.y, dlu.x, dlu.y));
        *reinterpret_cast<float2*>(sw + (m & 3) * DY_BUF + s_dy) = g2;
        sts128(sw + (m & 3) * ET_BUF + s_et, make_float4(u2.x, u2.y, sig.x, sig.y));
    };

    // ---- compute sub-chunk k: rebuild a_t, h_t forwards, walk the adjoint backwards ------------------------------------------------------
    auto compute = [&](int k, float2 hin0, float2 hin1) {
        const float* const rt = sw + (k & 3) * RT_BUF + c_rt;
        const float* const dyp = sw + (k & 3) * DY_BUF + c_dy;
        const float* const bc = sw + ((k >> 2) & 1) * BC_BUF + c_bc + (k & 3) * S;
        float* const pt = sw + (k & 1) * PT_BUF + c_pt;
        float* const rr = sw + (k & 1) * RR_BUF + c_rr;
        float4 Bq[2][2];
#pragma unroll
        for (int s = 0; s < 2; ++s) { Bq[s][0] = lds128(bc + s * GT); Bq[s][1] = lds128(bc + s * GT + 4); }
        float2 a[S][2], hs[S][2];
        float2 h[2] = {hin0, hin1};
#pragma unroll
        for (int j = 0; j < S; ++j) {
            const float4 v = lds128(rt + 4 * j);
            const float2 dl2 = lo2(v), du2 = hi2(v);
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const float2 x2 = mul2(dl2, A2[s]);
                a[j][s] = make_float2(ex2_approx(x2.x), ex2_approx(x2.y));
                h[s] = fma2(a[j][s], h[s], mul2(du2, bcast2(comp(Bq[s][j >> 2], j & 3))));
                hs[j][s] = h[s];
            }
        }
        float4 Cq[2], dyv, kB[2], kC[2];
#pragma unroll
        for (int j = S - 1; j >= 0; --j) {
            if ((j & 3) == 3) {
                Cq[0] = lds128(bc + BC_ARR + (j >> 2) * 4);
                Cq[1] = lds128(bc + BC_ARR + GT + (j >> 2) * 4);
            }
            if (j & 1) dyv = lds128(dyp + 2 * (j - 1));
            const float4 v = lds128(rt + 4 * j);
            const float2 dl2 = lo2(v), du2 = hi2(v);
            const float2 dy2 = (j & 1) ? hi2(dyv) : lo2(dyv);
            float2 sB2, sA2;
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const float2 dh = fma2(dy2, bcast2(comp(Cq[s], j & 3)), e2[s]);          // dh_t = C_t dy_t + a_{t+1} dh_{t+1}
                e2[s] = mul2(a[j][s], dh);
                const float2 hp = j > 0 ? hs[j - 1][s] : (s == 0 ? hin0 : hin1);
                const float2 w = mul2(e2[s], hp);                                           // d(loss) / d(delta_t A), per row
                dA2[s] = fma2(w, dl2, dA2[s]);
                sA2 = (s == 0) ? mul2(w, A2[0]) : fma2(w, A2[1], sA2);                      // x log2(e); folded back in the epilogue
                sB2 = (s == 0) ? mul2(dh, bcast2(comp(Bq[0][j >> 2], j & 3))) : fma2(dh, bcast2(comp(Bq[1][j >> 2], j & 3)), sB2);
                const float2 tc = mul2(dy2, hs[j][s]);
                const float2 tb = mul2(dh, du2);
                const float dCv = tc.x + tc.y, dBv = tb.x + tb.y;
                if ((j & 3) == 0) { kB[s].x = dBv; kC[s].x = dCv; }
                else if ((j & 3) == 1) { kB[s].y = dBv; kC[s].y = dCv; }
                else if ((j & 3) == 2) { kB[s].z = dBv; kC[s].z = dCv; }
                else { kB[s].w = dBv; kC[s].w = dCv; }
            }
            sts128(pt + 4 * (j * 8), make_float4(sB2.x, sB2.y, sA2.x, sA2.y));
            if ((j & 3) == 0) {
                const int half = j >> 2;
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    sts128(rr + 4 * ((0 * 2 + s) * 16 + half * 8), kB[s]);
                    sts128(rr + 4 * ((1 * 2 + s) * 16 + half * 8), kC[s]);
                }
            }
        }
    };

    // ---- epilogue of sub-chunk m: state sums -> du, ddelta of (row pair, te); row sums -> dB / dC of the lane's output row -------------
    auto epilogue = [&](int m) {
        const int t0 = m * S;
        const bool in = t0 + te < L;
        const float* const pt = sw + (m & 1) * PT_BUF + e_pt;
        float2 sB = make_float2(0.f, 0.f), sA = make_float2(0.f, 0.f);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
            const float4 v = lds128(pt + 4 * (kk ^ te));
            sB = add2(sB, lo2(v)); sA = add2(sA, hi2(v));
        }
        const float* const rr = sw + (m & 1) * RR_BUF + f_rr;
        float4 fin[2];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const float4 v0 = lds128(rr + 4 * (0 * 64 + half * 8)), v1 = lds128(rr + 4 * (1 * 64 + half * 8));
            const float4 v2 = lds128(rr + 4 * (2 * 64 + half * 8)), v3 = lds128(rr + 4 * (3 * 64 + half * 8));
            const float2 lo = add2(add2(lo2(v0), lo2(v1)), add2(lo2(v2), lo2(v3)));
            const float2 hi = add2(add2(hi2(v0), hi2(v1)), add2(hi2(v2), hi2(v3)));
            fin[half] = make_float4(lo.x, lo.y, hi.x, hi.y);
        }
        const float4 rtv = lds128(sw + (m & 3) * RT_BUF + s_rt);
        const float2 dy2 = *reinterpret_cast<const float2*>(sw + (m & 3) * DY_BUF + s_dy);
        const float4 etv = lds128(sw + (m & 3) * ET_BUF + s_et);
        const float2 du2 = fma2(lo2(rtv), sB, mul2(Dv2, dy2));
        float2 g = fma2(lo2(etv), sB, mul2(sA, bcast2(0.6931471805599453f)));             // d(loss)/d(Delta_t)
        g = mul2(g, hi2(etv));                                                             // x sigmoid(delta + bias) (1 without softplus)
        if (!in) g = make_float2(0.f, 0.f);
        dbias2 = add2(dbias2, g);
        if (in) {
            dus[t0] = Cvt<T>::from_f(du2.x); dus[du_ds + t0] = Cvt<T>::from_f(du2.y);
            dds[t0] = Cvt<T>::from_f(g.x); dds[dd_ds + t0] = Cvt<T>::from_f(g.y);
        }
        // warp-reduced dB / dC of this sub-chunk wait in the 32-step tile; a completed group (the 4 sub-chunks after this one) is
        // flushed first -- its loads precede this sub-chunk's stores into the same tile
        if ((m & 3) == 3 && m + 1 < n_sub) {
            dbc_flush((m + 1) >> 2);
            __syncwarp();
        }
        sts128(sw + f_dbw + (m & 3) * S, fin[0]);
        sts128(sw + f_dbw + (m & 3) * S + 4, fin[1]);
    };

    // ---- pipeline fill: the B / C tile of the last group, then stage the last sub-chunk ------------------------------------------------------
    {
        const int Gl = (n_sub - 1) >> 2;
#pragma unroll
        for (int part = 0; part < 4; ++part) {
            uint4 r[CPL];
            bc_load(Gl, part, r);
            bc_store(Gl, part, r);
        }
        // a partial last group has fewer than 4 trips: the parts of the next tile that those trips would have brought in
        for (int part = ((n_sub - 1) & 3) + 1; part < 4; ++part) {
            uint4 r[CPL];
            bc_load(Gl - 1, part, r);
            bc_store(Gl - 1, part, r);
        }
    }
    prefetch_raw(n_sub - 1);
    {
        const float2 x = add2(make_float2(widen1<T>(pe0), widen1<T>(pe1)), bias2);
        const float2 u2 = make_float2(widen1<T>(pu0), widen1<T>(pu1)), g2 = make_float2(widen1<T>(pg0), widen1<T>(pg1));
        prefetch_raw(n_sub - 2);
        prefetch_hck(n_sub - 1);
        bc_load(((n_sub - 1) >> 2) - 1, (n_sub - 1) & 3, pbc);
        stage(n_sub - 1, x, u2, g2);
        __syncwarp();
    }
    // ---- steady state: trip k = epilogue(k+1) | compute(k) | stage(k-1), loads of k-2 in flight; B / C tile part (k & 3) of the next
    //      group (k / 4 - 1) is stored, the part of trip k-1 is fetched ------------------------------------------------------------------------
    // (trip 0 stages sub-chunk -1 from masked loads into ring slots nobody reads.)
#pragma unroll 1
    for (int k = n_sub - 1; k >= 0; --k) {
        // everything fetched during the previous trip is consumed first (it landed long ago), so the loads of this trip go
        // straight into the loop-carried registers: no copy that would wait for them in mid-trip
        const float2 x = add2(make_float2(widen1<T>(pe0), widen1<T>(pe1)), bias2);
        const float2 u2 = make_float2(widen1<T>(pu0), widen1<T>(pu1)), g2 = make_float2(widen1<T>(pg0), widen1<T>(pg1));
        const float2 hin0 = make_float2(ph0.x, ph1.x), hin1 = make_float2(ph0.y, ph1.y);
        bc_store((k >> 2) - 1, k & 3, pbc);
        prefetch_raw(k - 2);
        prefetch_hck(k - 1);
        bc_load(((k - 1) >> 2) - 1, (k - 1) & 3, pbc);
        if (k + 1 < n_sub) epilogue(k + 1);
        compute(k, hin0, hin1);
        stage(k - 1, x, u2, g2);
        __syncwarp();
    }
    epilogue(0);
    __syncwarp();
    dbc_flush(0);

    // ---- whole-row results ------------------------------------------------------------------------------------------------------
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        atomicAdd(q.dA + static_cast<int64_t>(d0) * N + 2 * sg + s, dA2[s].x);
        atomicAdd(q.dA + static_cast<int64_t>(d1) * N + 2 * sg + s, dA2[s].y);
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        dD2.x += __shfl_xor_sync(0xffffffffu, dD2.x, o); dD2.y += __shfl_xor_sync(0xffffffffu, dD2.y, o);
        dbias2.x += __shfl_xor_sync(0xffffffffu, dbias2.x, o); dbias2.y += __shfl_xor_sync(0xffffffffu, dbias2.y, o);
    }
    if (te == 0) {
        if (q.dD) { atomicAdd(q.dD + d0, dD2.x); atomicAdd(q.dD + d1, dD2.y); }
        if (q.ddelta_bias) { atomicAdd(q.ddelta_bias + d0, dbias2.x); atomicAdd(q.ddelta_bias + d1, dbias2.y); }
    }
}

// Preconditions (cudaErrorInvalidConfiguration otherwise -> the caller falls back to the other backward kernels):
// dstate == 16, no z, channels per group a multiple of 8, dense checkpoints every 8 steps (or the sequence fits one sub-chunk).
template <typename T>
cudaError_t launch_scan_bwd_ls2_T(const FmScanBwdParams& q, cudaStream_t st, int vec_bc, int vec_dbc) {
    const FmScanFwdParams& p = q.f;
    const int dg = p.dim / p.n_groups;
    if (p.dstate != 16 || p.z != nullptr || dg % 8 != 0) return cudaErrorInvalidConfiguration;
    if (p.seqlen > ls2::S && !(p.hck != nullptr && p.hck_len == ls2::S)) return cudaErrorInvalidConfiguration;
    // 16-byte chunks of B / C and dB / dC: aligned rows, whole chunks (else: the first lane-serial kernel)
    if (!vec_bc || !vec_dbc || p.seqlen % (16 / (int)sizeof(T)) != 0) return cudaErrorInvalidConfiguration;
    const int64_t units = static_cast<int64_t>(p.batch) * p.n_groups * (dg / 8);
    int NW = env_int("FM_SCAN_BWD_LS2_NW", 1);
    if (NW != 1 && NW != 2 && NW != 4) NW = 1;
    const int64_t blocks = (units + NW - 1) / NW;
    if (blocks > 0x7fffffff) return cudaErrorInvalidConfiguration;
    const size_t smem = sizeof(float) * ls2::WARP_FLOATS * NW;
    void (*kern)(const FmScanBwdParams, int, int) =
        NW == 1 ? scan_bwd_ls2_kernel<T, 1> : (NW == 2 ? scan_bwd_ls2_kernel<T, 2> : scan_bwd_ls2_kernel<T, 4>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<static_cast<unsigned>(blocks), NW * 32, smem, st>>>(q, vec_bc, vec_dbc);
    count_launch();
    return cudaGetLastError();
}

}  // namespace fm
