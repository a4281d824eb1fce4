// fm_scan_bwd_rp.cuh -- selective-scan backward for sm_100a, "row-pair" kernel.
//
// Replaces selective_scan_bwd_kernel (selective_scan/selective_scan_bwd_kernel.cuh:75-489).  Same math
// (SURVEY.md section 3.5), different decomposition -- and a different one from fm_scan_bwd.cuh (kept as the generic
// fallback): the kernel is bound by the shared-memory / shuffle data path (128 B/clk/SM of register fill), so
//   * a lane owns TWO channel rows (float2 = (row0, row1)) and S = 8 timesteps; every B_t / C_t value fetched from
//     shared memory is a scalar-broadcast operand of a packed FMUL2/FFMA2 and serves both rows (half the fills per
//     (t, row, state)); B stays in registers between the forward recompute and the adjoint sweep;
//   * dB_t / dC_t are summed over the lane's two rows in registers before they touch the CTA's shared reduction tile
//     (half the read-modify-write traffic); the CTA's row pairs walk the states in a rotated order so that at any
//     step they add into different state rows with plain vector read-modify-writes (no atomics), and the tile is
//     flushed once per chunk with red.global.add.v4.f32;
//   * per (row pair, state): a_t is computed once (one MUFU.EX2 per (t, row, state)) and kept in registers; the
//     forward states of the chunk are rebuilt by an up-sweep + G-lane shuffle combine seeded from the dense
//     checkpoint `hck` written by the forward; the adjoint recurrence dh_t = C_t dy_t + a_{t+1} dh_{t+1} uses the
//     mirrored combine (shfl_down) seeded by the carried dh of the later chunk.  Chunks are walked in reverse.
//   * du, ddelta, dz leave as 128-bit stores; dA, dD, ddelta_bias are reduced in registers / shared memory over
//     the whole row and hit global memory with ONE atomic per (row, state) per CTA.
#pragma once
#include "fm_common.cuh"
#include "fm_launch.h"
#include "fm_scan_bwd.cuh"
#include "fm_scan_fwd_rp.cuh"

namespace fm {

__device__ __forceinline__ float2 shfl_down2(float2 v, int o, int w) {
    return make_float2(__shfl_down_sync(0xffffffffu, v.x, o, w), __shfl_down_sync(0xffffffffu, v.y, o, w));
}

// resident CTAs per SM the register allocation is capped for (occupancy vs spills, tuned on B200)
constexpr int bwd_rp_minb(int NW) { return NW == 4 ? 3 : (NW == 2 ? 6 : (NW == 1 ? 12 : 1)); }

#ifndef FM_BWD_PIPE_F32
#define FM_BWD_PIPE_F32 0   // fp32 I/O: the same prefetch costs 5-12 % (profiles/r01_bwd_prefetch_ab.jsonl), kept off
#endif
template <typename T, int S, int G, int NW, bool kHasZ, int MINB, int kN>
__global__ void __launch_bounds__(NW * 32, MINB)
scan_bwd_rp_kernel(const FmScanBwdParams q, const int vec_io, const int vec_bc, const int vec_dbc) {
    const FmScanFwdParams& p = q.f;
    constexpr int TC = G * S;
    constexpr int PW = 32 / G;              // row pairs per warp
    constexpr int RP = NW * PW;             // row pairs per CTA
    constexpr int R = 2 * RP;
    constexpr bool SWZ = (S == 8);          // 8-element segments: swizzled, unpadded tile rows (fm_common.cuh)
    constexpr int SP = tile_seg_pitch<S, SWZ>();
    constexpr int ROWP = G * SP;
    constexpr int NT = NW * 32;

    const int N = kN > 0 ? kN : p.dstate;   // kN: compile-time dstate (index math of the tile loops folds to shifts)
    const int L = p.seqlen;
    const int dg = p.dim / p.n_groups;
    const int tiles_per_group = dg / R;     // launcher guarantees dg % R == 0 (no shadow rows) and RP <= N
    const int group = blockIdx.x / tiles_per_group;
    const int tile = blockIdx.x % tiles_per_group;
    const int b = blockIdx.y;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int seg = lane % G;
    const int rp = warp * PW + lane / G;
    const int d0 = group * dg + tile * R + 2 * rp, d1 = d0 + 1;

    extern __shared__ __align__(16) float smem[];
    float* sBC = smem;                                               // [B|C][N][ROWP]  (single stage: one chunk = dstate long
                                                                     //  state iterations, the tile load is amortised)
    float* sdBC = sBC + 2 * N * ROWP;                                // [dB|dC][N][ROWP] on-chip reduction tile
    float2* sA = reinterpret_cast<float2*>(sdBC + 2 * N * ROWP);     // [RP][N]  A (natural units) of both rows
    float2* sHs = sA + RP * N;                                       // [RP][N]  forward state at chunk start
    float4* sCar = reinterpret_cast<float4*>(sHs + RP * N);          // [RP][N]  (dh.x, dh.y, a.x, a.y) of the first step of the later chunk
    float2* sdA = reinterpret_cast<float2*>(sCar + RP * N);          // [N][NT]  per-thread dA partials

    const T* __restrict__ Bg = reinterpret_cast<const T*>(p.B) + b * p.B_batch_stride + group * p.B_group_stride;
    const T* __restrict__ Cg = reinterpret_cast<const T*>(p.C) + b * p.C_batch_stride + group * p.C_group_stride;
    float* __restrict__ dBg = q.dB + b * q.dB_batch_stride + group * q.dB_group_stride;
    float* __restrict__ dCg = q.dC + b * q.dC_batch_stride + group * q.dC_group_stride;
    const T* __restrict__ ub = reinterpret_cast<const T*>(p.u) + b * p.u_batch_stride;
    const T* __restrict__ db = reinterpret_cast<const T*>(p.delta) + b * p.delta_batch_stride;
    const T* __restrict__ gb = reinterpret_cast<const T*>(q.dout) + b * q.dout_batch_stride;
    T* __restrict__ dub = reinterpret_cast<T*>(q.du) + b * q.du_batch_stride;
    T* __restrict__ ddb = reinterpret_cast<T*>(q.ddelta) + b * q.ddelta_batch_stride;
    const float* __restrict__ hck0 =
        p.hck ? reinterpret_cast<const float*>(p.hck) + (static_cast<int64_t>(b) * p.dim + d0) * p.n_hck * N : nullptr;
    const float* __restrict__ hck1 = p.hck ? hck0 + static_cast<int64_t>(p.n_hck) * N : nullptr;

    const float2 Dv = p.D ? make_float2(reinterpret_cast<const float*>(p.D)[d0], reinterpret_cast<const float*>(p.D)[d1])
                          : make_float2(0.f, 0.f);
    const float2 bias = p.delta_bias ? make_float2(reinterpret_cast<const float*>(p.delta_bias)[d0],
                                                   reinterpret_cast<const float*>(p.delta_bias)[d1])
                                     : make_float2(0.f, 0.f);

    for (int i = tid; i < RP * N; i += NT) {
        const int r = i / N, n = i % N;
        const int e0 = group * dg + tile * R + 2 * r;
        const float* Ap = reinterpret_cast<const float*>(p.A);
        sA[i] = make_float2(Ap[e0 * p.A_d_stride + n * p.A_dstate_stride], Ap[(e0 + 1) * p.A_d_stride + n * p.A_dstate_stride]);
        sCar[i] = make_float4(0.f, 0.f, 1.f, 1.f);
    }
    for (int i = tid; i < N * NT; i += NT) sdA[i] = make_float2(0.f, 0.f);
    for (int i = tid; i < 2 * N * ROWP; i += NT) sdBC[i] = 0.f;

    const int n_chunks = (L + TC - 1) / TC;

    float2 dD_acc = make_float2(0.f, 0.f), dbias_acc = make_float2(0.f, 0.f);
    float2 dfirst_next = make_float2(0.f, 0.f);   // softplus'd delta of the first step of the later chunk
    // Rotated state order (see header): row pair rp starts at state rp*RSTEP and walks upwards, so at any step the RP row
    // pairs of the CTA update RP different state rows of the dB/dC tile.  With a compile-time dstate that is a multiple
    // of RP the starts are RSTEP > 1 rows apart: two warps that are up to RSTEP-1 steps out of phase still touch
    // different rows, and one CTA barrier per RSTEP steps keeps them that close.
    constexpr int RSTEP = (kN > 0 && kN % RP == 0) ? kN / RP : 1;
    const int n_first = (rp * RSTEP) % N;

    // 16-bit I/O, one CTA per SM: the u / delta / dout segments of the chunk about to be processed are loaded one chunk
    // ahead (behind the previous chunk's outputs, in front of its dB/dC flush) into loop-carried raw registers, so the
    // prologue does not wait on them.  Measured on B200 (profiles/r01_bwd_prefetch_ab.jsonl): -5 % for bf16; for fp32
    // I/O (twice the registers, or a cp.async staging area in smem) the same change costs 5-10 %, and the
    // register-capped variants would spill, so those keep the plain loads.
    constexpr bool kPipeR = (MINB <= 2) && (sizeof(T) == 2 || FM_BWD_PIPE_F32) && S == 8;
    SegRaw<T, S> ru0, ru1, re0, re1, rg0, rg1;
    auto load_chunk = [&](int cc) {
        const int tt = cc * TC + seg * S;
        const int nv = L - tt;
        load_seg_raw<T, S>(ub + d0 * p.u_d_stride + tt, nv, vec_io, ru0);
        load_seg_raw<T, S>(ub + d1 * p.u_d_stride + tt, nv, vec_io, ru1);
        load_seg_raw<T, S>(db + d0 * p.delta_d_stride + tt, nv, vec_io, re0);
        load_seg_raw<T, S>(db + d1 * p.delta_d_stride + tt, nv, vec_io, re1);
        load_seg_raw<T, S>(gb + d0 * q.dout_d_stride + tt, nv, vec_io, rg0);
        load_seg_raw<T, S>(gb + d1 * q.dout_d_stride + tt, nv, vec_io, rg1);
    };
    if constexpr (kPipeR) load_chunk(n_chunks - 1);

    for (int it = 0; it < n_chunks; ++it) {
        const int c = n_chunks - 1 - it;
        // every warp is past the previous chunk's state loop (its last step ends with a barrier); the flush of the
        // reduction tile touches a different region, so the B/C tile can be refilled now
        // fp32 tiles arrive by cp.async.  16-bit tiles need a widening pass: with a compile-time dstate their 16-byte
        // packets are only LOADED here (raw bits, parked in registers) and widened into the tile at the end of the
        // prologue, so the loads fly behind the rest of the prologue instead of stalling it at the first convert.
        constexpr bool kRawBC = sizeof(T) == 2 && kN > 0 && S == 8;
        constexpr int QPB = TC / 8;                                   // 8-element packets per state row
        constexpr int NQ = kRawBC ? (2 * (kN > 0 ? kN : 1) * QPB + NT - 1) / NT : 1;
        uint4 rawbc[NQ];
        if constexpr (kRawBC) {
#pragma unroll
            for (int i = 0; i < NQ; ++i) {
                const int sidx = tid + i * NT;
                const int which = sidx / (kN * QPB), rem = sidx % (kN * QPB);
                const int n = rem / QPB, t = c * TC + 8 * (rem % QPB);
                const T* g = (which ? Cg + n * p.C_dstate_stride : Bg + n * p.B_dstate_stride) + t;
                rawbc[i] = make_uint4(0u, 0u, 0u, 0u);
                if (sidx < 2 * kN * QPB && vec_bc && t + 8 <= L) rawbc[i] = __ldg(reinterpret_cast<const uint4*>(g));
            }
        } else {
            stage_tile<T, TC, S, SWZ>(sBC, Bg, p.B_dstate_stride, N, c * TC, L, vec_bc, tid, NT);
            stage_tile<T, TC, S, SWZ>(sBC + N * ROWP, Cg, p.C_dstate_stride, N, c * TC, L, vec_bc, tid, NT);
        }
        cp_async_commit();
        // forward state at the start of this chunk -> sHs (lane seg loads states seg, seg+G, ...).  With a compile-time
        // dstate the loads are issued here and parked in registers until the segment loads below are in flight too
        // (otherwise the shared store right behind them serialises two memory latencies per chunk).
        const int hoff = (c > 0) ? (c * TC / p.hck_len - 1) * N : -1;
        constexpr int NH = kN > 0 ? (kN + G - 1) / G : 1;
        float2 hreg[NH];
        if constexpr (kN > 0) {
#pragma unroll
            for (int i = 0; i < NH; ++i) {
                const int n = seg + i * G;
                hreg[i] = (hoff >= 0 && n < kN) ? make_float2(__ldg(hck0 + hoff + n), __ldg(hck1 + hoff + n)) : make_float2(0.f, 0.f);
            }
        } else {
            for (int n = seg; n < N; n += G)
                sHs[rp * N + n] = hoff >= 0 ? make_float2(hck0[hoff + n], hck1[hoff + n]) : make_float2(0.f, 0.f);
        }
        const int t0 = c * TC + seg * S;
        const int nvalid = L - t0;
        float2 dl2[S], du2[S], dy2[S], s2[S], dd2[S];
        {
            float u0[S], u1[S], e0[S], e1[S], g0[S], g1[S];
            if constexpr (kPipeR) {
                widen_seg<T, S>(ru0, u0); widen_seg<T, S>(ru1, u1);
                widen_seg<T, S>(re0, e0); widen_seg<T, S>(re1, e1);
                widen_seg<T, S>(rg0, g0); widen_seg<T, S>(rg1, g1);
            } else {
                load_seg<T, S>(ub + d0 * p.u_d_stride + t0, nvalid, vec_io, u0);
                load_seg<T, S>(ub + d1 * p.u_d_stride + t0, nvalid, vec_io, u1);
                load_seg<T, S>(db + d0 * p.delta_d_stride + t0, nvalid, vec_io, e0);
                load_seg<T, S>(db + d1 * p.delta_d_stride + t0, nvalid, vec_io, e1);
                load_seg<T, S>(gb + d0 * q.dout_d_stride + t0, nvalid, vec_io, g0);
                load_seg<T, S>(gb + d1 * q.dout_d_stride + t0, nvalid, vec_io, g1);
            }
            if constexpr (kHasZ) {
                const T* zb = reinterpret_cast<const T*>(p.z) + b * p.z_batch_stride;
                const T* yb = reinterpret_cast<const T*>(p.out) + b * p.out_batch_stride;
                T* dzb = reinterpret_cast<T*>(q.dz) + b * q.dz_batch_stride;
                T* ozb = p.out_z ? reinterpret_cast<T*>(p.out_z) + b * p.out_z_batch_stride : nullptr;
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const int d = r ? d1 : d0;
                    float* gg = r ? g1 : g0;
                    float zv[S], yv[S], dzv[S];
                    load_seg<T, S>(zb + d * p.z_d_stride + t0, nvalid, vec_io, zv);
                    load_seg<T, S>(yb + d * p.out_d_stride + t0, nvalid, vec_io, yv);
#pragma unroll
                    for (int i = 0; i < S; ++i) {
                        const float sg_ = sigmoid_f(zv[i]);
                        const float g = gg[i];
                        dzv[i] = g * yv[i] * sg_ * (1.f + zv[i] * (1.f - sg_));
                        gg[i] = g * zv[i] * sg_;
                        yv[i] = yv[i] * zv[i] * sg_;          // recomputed out_z
                    }
                    if (nvalid > 0) {
                        store_seg<T, S>(dzb + d * q.dz_d_stride + t0, nvalid, vec_io, dzv);
                        if (ozb) store_seg<T, S>(ozb + d * p.out_z_d_stride + t0, nvalid, vec_io, yv);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < S; ++i) {
                const float x0 = e0[i] + bias.x, x1 = e1[i] + bias.y;
                const float sp0 = p.delta_softplus ? softplus_fast(x0) : x0;
                const float sp1 = p.delta_softplus ? softplus_fast(x1) : x1;
                const bool in = i < nvalid;
                dl2[i] = in ? make_float2(sp0, sp1) : make_float2(0.f, 0.f);   // masked steps: a = 1, b = 0
                const float2 uu = make_float2(u0[i], u1[i]);
                dy2[i] = in ? make_float2(g0[i], g1[i]) : make_float2(0.f, 0.f);
                du2[i] = mul2(dl2[i], uu);
                dD_acc = fma2(dy2[i], uu, dD_acc);
                s2[i] = make_float2(0.f, 0.f);
                dd2[i] = make_float2(0.f, 0.f);
            }
        }
        float2 sumd2 = dl2[0];
#pragma unroll
        for (int i = 1; i < S; ++i) sumd2 = add2(sumd2, dl2[i]);
        // shifted sum: sum over the segment of delta_{t+1}
        float2 dnext0 = shfl_down2(dl2[0], 1, G);
        if (seg == G - 1) dnext0 = dfirst_next;
        const float2 sumd_sh = add2(add2(sumd2, make_float2(-dl2[0].x, -dl2[0].y)), dnext0);
        dfirst_next = shfl_idx2(dl2[0], 0, G);

        if constexpr (kN > 0) {
#pragma unroll
            for (int i = 0; i < NH; ++i)
                if (seg + i * G < kN) sHs[rp * N + seg + i * G] = hreg[i];
        }
        if constexpr (kRawBC) {
#pragma unroll
            for (int i = 0; i < NQ; ++i) {
                const int sidx = tid + i * NT;
                if (sidx >= 2 * kN * QPB) continue;
                const int which = sidx / (kN * QPB), rem = sidx % (kN * QPB);
                const int n = rem / QPB, q8 = rem % QPB, t = c * TC + 8 * q8;
                float f[8];
                if (vec_bc && t + 8 <= L) {
                    const T* e = reinterpret_cast<const T*>(&rawbc[i]);
#pragma unroll
                    for (int j = 0; j < 8; ++j) f[j] = Cvt<T>::to_f(e[j]);
                } else {
                    const T* g = (which ? Cg + n * p.C_dstate_stride : Bg + n * p.B_dstate_stride) + t;
#pragma unroll
                    for (int j = 0; j < 8; ++j) f[j] = (t + j < L) ? Cvt<T>::to_f(g[j]) : 0.f;
                }
                float* row = sBC + (which * N + n) * ROWP;
                sts128(row + tile_off<S, SWZ>(8 * q8), make_float4(f[0], f[1], f[2], f[3]));
                sts128(row + tile_off<S, SWZ>(8 * q8 + 4), make_float4(f[4], f[5], f[6], f[7]));
            }
        }
        cp_async_wait<0>();
        __syncthreads();      // B/C tile, sHs and the cleared reduction tile are visible to every warp

        // pull the earlier chunk's u / delta / dout segments of this lane into L2 while this chunk's state loop runs
        if (c > 0) {
            const int tn = t0 - TC;
            prefetch_l2(ub + d0 * p.u_d_stride + tn); prefetch_l2(ub + d1 * p.u_d_stride + tn);
            prefetch_l2(db + d0 * p.delta_d_stride + tn); prefetch_l2(db + d1 * p.delta_d_stride + tn);
            prefetch_l2(gb + d0 * q.dout_d_stride + tn); prefetch_l2(gb + d1 * q.dout_d_stride + tn);
            if (c > 1 && seg * 32 < N) {    // its checkpoint rows too (N floats per row: one 128-byte line per 32 states)
                const int hn = ((c - 1) * TC / p.hck_len - 1) * N + seg * 32;
                prefetch_l2(hck0 + hn); prefetch_l2(hck1 + hn);
            }
        }

        const float* tB = sBC + seg * SP;
        const float* tC = tB + N * ROWP;
        float* tdB = sdBC + seg * SP;
        float* tdC = tdB + N * ROWP;
        // offsets of this lane's 4-timestep quarters inside its segment (halves swapped on swizzled rows)
        int qoff[S / 4];
#pragma unroll
        for (int i = 0; i < S / 4; ++i) qoff[i] = SWZ ? ((i ^ (seg >> 2)) & 1) << 2 : 4 * i;
        const int rbase = rp * N;

#pragma unroll 1
        for (int k = 0; k < N; ++k) {
            int n = n_first + k;
            if (n >= N) n -= N;
            const int nro = n * ROWP;
            const float2 An = sA[rbase + n];
            const float2 A2 = mul2(An, bcast2(kLog2e));
            const float2 hstart = sHs[rbase + n];
            const float4 car = sCar[rbase + n];
            const float2 dhrun = make_float2(car.x, car.y);

            float bv[S];
            float2 a2[S], g2[S];                         // g2 holds b_t first, then g_t = a_t * h_{t-1}
            {
#pragma unroll
                for (int i = 0; i < S / 4; ++i) {
                    const float4 v = lds128(tB + nro + qoff[i]);
                    bv[4 * i] = v.x; bv[4 * i + 1] = v.y; bv[4 * i + 2] = v.z; bv[4 * i + 3] = v.w;
                }
#pragma unroll
                for (int j = 0; j < S; ++j) {
                    g2[j] = mul2(du2[j], bcast2(bv[j]));
                    const float2 x2 = mul2(dl2[j], A2);
                    a2[j] = make_float2(ex2_approx(x2.x), ex2_approx(x2.y));
                }
            }
            // C_t dy_t for the adjoint recurrence dh_t = C_t dy_t + a_{t+1} dh_{t+1}
            float2 cd2[S];
            {
                float cv[S];
#pragma unroll
                for (int i = 0; i < S / 4; ++i) {
                    const float4 v = lds128(tC + nro + qoff[i]);
                    cv[4 * i] = v.x; cv[4 * i + 1] = v.y; cv[4 * i + 2] = v.z; cv[4 * i + 3] = v.w;
                }
#pragma unroll
                for (int j = 0; j < S; ++j) cd2[j] = mul2(dy2[j], bcast2(cv[j]));
            }
            float2 anext = shfl_down2(a2[0], 1, G);
            if (seg == G - 1) anext = make_float2(car.z, car.w);
            // ---- the forward-state scan and the adjoint scan are independent until the gradient products: their
            //      up-sweeps and G-lane combines are interleaved (two dependency chains in flight per lane) ----------
            float2 h = g2[0];                            // forward: segment state from zero
            float2 r = cd2[S - 1];                       // adjoint (right to left): dh at the first step given dh_in = 0
#pragma unroll
            for (int j = 1; j < S; ++j) {
                h = fma2(a2[j], h, g2[j]);
                r = fma2(a2[S - j], r, cd2[S - 1 - j]);
            }
            const float2 ps = mul2(A2, sumd2), prs = mul2(A2, sumd_sh);
            float2 P = make_float2(ex2_approx(ps.x), ex2_approx(ps.y));
            float2 Pr = make_float2(ex2_approx(prs.x), ex2_approx(prs.y));
            if (seg == 0) h = fma2(P, hstart, h);
            if (seg == G - 1) r = fma2(Pr, dhrun, r);
#pragma unroll
            for (int o = 1; o < G; o <<= 1) {
                const float2 hp = shfl_up2(h, o, G);
                const float2 rp_ = shfl_down2(r, o, G);
                float2 Pp = make_float2(1.f, 1.f), Prp = make_float2(1.f, 1.f);
                if (2 * o < G) { Pp = shfl_up2(P, o, G); Prp = shfl_down2(Pr, o, G); }
                if (seg >= o) {
                    h = fma2(P, hp, h);
                    if (2 * o < G) P = mul2(P, Pp);
                }
                if (seg + o < G) {
                    r = fma2(Pr, rp_, r);
                    if (2 * o < G) Pr = mul2(Pr, Prp);
                }
            }
            float2 hin = shfl_up2(h, 1, G);
            if (seg == 0) hin = hstart;
            float2 dh = shfl_down2(r, 1, G);             // dh at the first step of the next lane
            if (seg == G - 1) dh = dhrun;
            __syncwarp();
            if (seg == 0) sCar[rbase + n] = make_float4(r.x, r.y, a2[0].x, a2[0].y);   // carry to the earlier chunk
            // forward down-sweep: g_t = a_t h_{t-1}, h_t = g_t + b_t; dC_t = sum over the two rows of dy_t h_t
            {
                float dc[S];
#pragma unroll
                for (int j = 0; j < S; ++j) {
                    const float2 gj = mul2(a2[j], hin);
                    hin = add2(gj, g2[j]);
                    g2[j] = gj;
                    const float2 t2 = mul2(dy2[j], hin);
                    dc[j] = t2.x + t2.y;
                }
#pragma unroll
                for (int i = 0; i < S / 4; ++i) {
                    const float4 o = lds128(tdC + nro + qoff[i]);
                    sts128(tdC + nro + qoff[i], make_float4(o.x + dc[4 * i], o.y + dc[4 * i + 1], o.z + dc[4 * i + 2], o.w + dc[4 * i + 3]));
                }
            }
            // down-sweep (right to left) with the packed gradient products
            float2 dA2 = make_float2(0.f, 0.f);
            {
                float dbs[S];
#pragma unroll
                for (int j = S - 1; j >= 0; --j) {
                    const float2 an = (j == S - 1) ? anext : a2[j + 1];
                    dh = fma2(an, dh, cd2[j]);                           // dh_j
                    s2[j] = fma2(dh, bcast2(bv[j]), s2[j]);
                    const float2 w2 = mul2(dh, g2[j]);                   // dh_t * (h_t - b_t)
                    dd2[j] = fma2(w2, An, dd2[j]);
                    dA2 = fma2(dl2[j], w2, dA2);
                    const float2 t2 = mul2(dh, du2[j]);
                    dbs[j] = t2.x + t2.y;
                }
#pragma unroll
                for (int i = 0; i < S / 4; ++i) {
                    const float4 o = lds128(tdB + nro + qoff[i]);
                    sts128(tdB + nro + qoff[i], make_float4(o.x + dbs[4 * i], o.y + dbs[4 * i + 1], o.z + dbs[4 * i + 2], o.w + dbs[4 * i + 3]));
                }
            }
            sdA[n * NT + tid] = add2(sdA[n * NT + tid], dA2);
            if (RSTEP == 1 || (k + 1) % RSTEP == 0)
                __syncthreads();   // keep the row rotation aligned (one step = one state row per CTA row pair)
        }

        // per-element outputs.  With softplus nothing is re-read: u*s = (delta*u)*s / delta, and the softplus derivative is
        // sigmoid(x) = 1 - exp(-softplus(x)), evaluated from the delta already in registers (series for small delta: no
        // cancellation).  delta == 0 (masked step, or exp underflow where sigmoid(x) is 0 anyway) contributes exactly 0.
        {
            float o0[S], o1[S];
#pragma unroll
            for (int j = 0; j < S; ++j) {
                const float2 o2 = fma2(dl2[j], s2[j], mul2(dy2[j], Dv));
                o0[j] = o2.x; o1[j] = o2.y;
            }
            if (nvalid > 0) {
                store_seg<T, S>(dub + d0 * q.du_d_stride + t0, nvalid, vec_io, o0);
                store_seg<T, S>(dub + d1 * q.du_d_stride + t0, nvalid, vec_io, o1);
            }
            if (p.delta_softplus) {
#pragma unroll
                for (int j = 0; j < S; ++j) {
                    const float2 d = dl2[j];
                    const float2 us = mul2(mul2(du2[j], s2[j]), make_float2(rcp_approx(d.x), rcp_approx(d.y)));
                    float2 g = add2(us, dd2[j]);
                    g = mul2(g, make_float2(one_minus_exp_neg(d.x), one_minus_exp_neg(d.y)));
                    if (!(d.x > 0.f)) g.x = 0.f;
                    if (!(d.y > 0.f)) g.y = 0.f;
                    o0[j] = g.x; o1[j] = g.y;
                    dbias_acc = add2(dbias_acc, g);
                }
            } else {
                float u0[S], u1[S];
                load_seg<T, S>(ub + d0 * p.u_d_stride + t0, nvalid, vec_io, u0);   // re-read (L2 hit) instead of holding registers
                load_seg<T, S>(ub + d1 * p.u_d_stride + t0, nvalid, vec_io, u1);
#pragma unroll
                for (int j = 0; j < S; ++j) {
                    float2 g = fma2(make_float2(u0[j], u1[j]), s2[j], dd2[j]);
                    if (j >= nvalid) g = make_float2(0.f, 0.f);
                    o0[j] = g.x; o1[j] = g.y;
                    dbias_acc = add2(dbias_acc, g);
                }
            }
            if (nvalid > 0) {
                store_seg<T, S>(ddb + d0 * q.ddelta_d_stride + t0, nvalid, vec_io, o0);
                store_seg<T, S>(ddb + d1 * q.ddelta_d_stride + t0, nvalid, vec_io, o1);
            }
        }

        if constexpr (kPipeR) if (c > 0) load_chunk(c - 1);   // in flight across the flush and the next B/C staging

        // flush the CTA's dB/dC tile: one vector red per 4 timesteps per state, then clear it for the next chunk
        {
            constexpr int QPR = TC / 4;
            const int tc0 = c * TC;
            for (int s = tid; s < 2 * N * QPR; s += NT) {
                const int which = s / (N * QPR);
                const int rem = s % (N * QPR);
                const int n = rem / QPR, qq = rem % QPR;
                float4* src = reinterpret_cast<float4*>(sdBC + which * N * ROWP + n * ROWP + tile_off<S, SWZ>(4 * qq));
                const float4 v = *src;
                *src = make_float4(0.f, 0.f, 0.f, 0.f);
                const int t = tc0 + 4 * qq;
                float* dst = (which ? dCg + n * q.dC_dstate_stride : dBg + n * q.dB_dstate_stride) + t;
                if (vec_dbc && t + 4 <= L) {
                    red_add_v4(dst, v.x, v.y, v.z, v.w);
                } else {
                    if (t < L) atomicAdd(dst, v.x);
                    if (t + 1 < L) atomicAdd(dst + 1, v.y);
                    if (t + 2 < L) atomicAdd(dst + 2, v.z);
                    if (t + 3 < L) atomicAdd(dst + 3, v.w);
                }
            }
        }
        // the next iteration's __syncthreads (after staging) orders the tile clear before the next adds
    }

    // row reductions -> one atomic per (row, state) / row
#pragma unroll
    for (int o = 1; o < G; o <<= 1) {
        dD_acc.x += __shfl_xor_sync(0xffffffffu, dD_acc.x, o, G);
        dD_acc.y += __shfl_xor_sync(0xffffffffu, dD_acc.y, o, G);
        dbias_acc.x += __shfl_xor_sync(0xffffffffu, dbias_acc.x, o, G);
        dbias_acc.y += __shfl_xor_sync(0xffffffffu, dbias_acc.y, o, G);
    }
    if (seg == 0) {
        if (q.dD) { atomicAdd(q.dD + d0, dD_acc.x); atomicAdd(q.dD + d1, dD_acc.y); }
        if (q.ddelta_bias) { atomicAdd(q.ddelta_bias + d0, dbias_acc.x); atomicAdd(q.ddelta_bias + d1, dbias_acc.y); }
    }
    for (int n = 0; n < N; ++n) {
        float2 v = sdA[n * NT + tid];
#pragma unroll
        for (int o = 1; o < G; o <<= 1) {
            v.x += __shfl_xor_sync(0xffffffffu, v.x, o, G);
            v.y += __shfl_xor_sync(0xffffffffu, v.y, o, G);
        }
        if (seg == 0) {
            atomicAdd(q.dA + static_cast<int64_t>(d0) * N + n, v.x);
            atomicAdd(q.dA + static_cast<int64_t>(d1) * N + n, v.y);
        }
    }
}

template <int S, int G, int NW>
constexpr size_t bwd_rp_smem_bytes(int dstate) {
    // B/C tile (2) + reduction tile (2) planes of dstate*G*seg_pad(S) floats; sA, sHs (float2), sCar (float4) per (pair, state);
    // per-thread dA partials (float2)
    return sizeof(float) * (4 * (size_t)dstate * G * (S == 8 ? 8 : seg_pad(S)) + 8 * (size_t)NW * (32 / G) * dstate + 2 * (size_t)dstate * NW * 32);
}

template <typename T, int S, int G, int NW>
static cudaError_t launch_bwd_rp_cfg(const FmScanBwdParams& q, cudaStream_t st, int vec_io, int vec_bc, int vec_dbc) {
    const FmScanFwdParams& p = q.f;
    constexpr int R = 2 * NW * (32 / G), NT = NW * 32;
    const int dg = p.dim / p.n_groups;
    dim3 grid((dg / R) * p.n_groups, p.batch);
    const size_t smem = bwd_rp_smem_bytes<S, G, NW>(p.dstate);
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    constexpr int MB = (S == 4) ? (NW == 8 ? 2 : (NW == 4 ? 4 : 8)) : bwd_rp_minb(NW);
    // dstate == 16 (every SS2D in the reference) gets a compile-time-dstate instance for the production tile shapes
    constexpr bool kFixed16 = (S == 8 && NW >= 4);
    void (*kern)(const FmScanBwdParams, const int, const int, const int) =
        p.z ? scan_bwd_rp_kernel<T, S, G, NW, true, MB, 0> : scan_bwd_rp_kernel<T, S, G, NW, false, MB, 0>;
    if constexpr (kFixed16) {
        if (p.dstate == 16 && env_int("FM_SCAN_BWD_FIXN", 1))
            kern = p.z ? scan_bwd_rp_kernel<T, S, G, NW, true, MB, 16> : scan_bwd_rp_kernel<T, S, G, NW, false, MB, 16>;
    }
    if constexpr (kFixed16 && NW == 4) {
        // two register-uncapped CTAs per SM instead of three capped (spilling) ones; fits since the swizzled tiles dropped
        // the padding.  Default whenever the sequence spans several chunks (single-chunk shapes prefer the occupancy);
        // measured in profiles/r01_bwd_two_cta_ab.jsonl
        if (p.dstate == 16 && env_int("FM_SCAN_BWD_MINB", (G == 32 || p.seqlen > G * S) ? 2 : 3) == 2)
            kern = p.z ? scan_bwd_rp_kernel<T, S, G, NW, true, 2, 16> : scan_bwd_rp_kernel<T, S, G, NW, false, 2, 16>;
    }
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, NT, smem, st>>>(q, vec_io, vec_bc, vec_dbc);
    count_launch();
    return cudaGetLastError();
}

// Returns cudaErrorInvalidConfiguration when the shape does not fit this kernel's preconditions (the caller then
// uses the generic kernel of fm_scan_bwd.cuh): whole tiles only (dg % R == 0), R/2 <= dstate (row-pair rotation),
// multi-chunk runs need hck with TC % hck_len == 0.
template <typename T>
cudaError_t launch_scan_bwd_rp_T(const FmScanBwdParams& q, cudaStream_t st, int vec_io, int vec_bc, int vec_dbc) {
    const FmScanFwdParams& p = q.f;
    const int dg = p.dim / p.n_groups;
    int S = env_int("FM_SCAN_BWD_S", 8);
    if (S != 4 && S != 8) S = 8;
    int G = env_int("FM_SCAN_BWD_G", 0);
    if (G <= 0) G = scan_lanes_per_row(((int64_t)p.batch * p.dim + 1) / 2, p.seqlen, S, "FM_SCAN_BWD_G");
    if (!p.hck) {
        G = 1;
        while (G < 32 && G * S < p.seqlen) G <<= 1;
        if (G * S < p.seqlen) return cudaErrorInvalidConfiguration;
    } else if (G * S < p.seqlen) {
        while (G < 32 && (G * S) % p.hck_len != 0) G <<= 1;
        if ((G * S) % p.hck_len != 0) return cudaErrorInvalidConfiguration;
    }
    int NW = env_int("FM_SCAN_BWD_NW", 0);
    // tuned on B200 (profiles/r01_bwd_rp_tune.jsonl, r01_bwd_two_cta_ab.jsonl): G = 32 runs two 4-warp CTAs per SM when the
    // compile-time-dstate instance applies, else one 8-warp CTA
    if (NW != 1 && NW != 2 && NW != 4 && NW != 8) NW = (G == 32 && p.dstate != 16) ? 8 : 4;
    while (NW > 1 && (NW * (32 / G) > p.dstate || dg % (2 * NW * (32 / G)) != 0)) NW >>= 1;
    if (NW * (32 / G) > p.dstate || dg % (2 * NW * (32 / G)) != 0) return cudaErrorInvalidConfiguration;
    const size_t need = sizeof(float) * (4 * (size_t)p.dstate * G * (S == 8 ? 8 : seg_pad(S)) + 8 * (size_t)NW * (32 / G) * p.dstate + 2 * (size_t)p.dstate * NW * 32);
    if (need > 200 * 1024) return cudaErrorInvalidConfiguration;
#define FM_CASE_BRP(s_, g, nw) if (S == s_ && G == g && NW == nw) return launch_bwd_rp_cfg<T, s_, g, nw>(q, st, vec_io, vec_bc, vec_dbc);
    FM_CASE_BRP(8, 2, 1)
    FM_CASE_BRP(8, 4, 1) FM_CASE_BRP(8, 4, 2)
    FM_CASE_BRP(8, 8, 1) FM_CASE_BRP(8, 8, 2) FM_CASE_BRP(8, 8, 4)
    FM_CASE_BRP(8, 16, 1) FM_CASE_BRP(8, 16, 2) FM_CASE_BRP(8, 16, 4) FM_CASE_BRP(8, 16, 8)
    FM_CASE_BRP(8, 32, 1) FM_CASE_BRP(8, 32, 2) FM_CASE_BRP(8, 32, 4) FM_CASE_BRP(8, 32, 8)
    FM_CASE_BRP(4, 16, 4) FM_CASE_BRP(4, 16, 8) FM_CASE_BRP(4, 32, 4) FM_CASE_BRP(4, 32, 8) FM_CASE_BRP(4, 16, 2) FM_CASE_BRP(4, 8, 4)
#undef FM_CASE_BRP
    return cudaErrorInvalidConfiguration;
}

}  // namespace fm
