// fm_scan_bwd.cuh -- selective-scan backward for sm_100a.
//
// Replaces selective_scan_bwd_kernel (selective_scan/selective_scan_bwd_kernel.cuh:75-489).  Same math
// (SURVEY.md section 3.5), different decomposition:
//   * same CTA tiling as the forward (R rows of one (batch, group); B and C tiles staged once per chunk in
//     shared memory and shared by all rows); chunks are walked in REVERSE order.
//   * per (row, state): a_t is computed once (one ex2) and kept in registers; the forward states of the
//     chunk are rebuilt by an up-sweep + G-lane shuffle combine seeded from the dense checkpoint `hck`
//     written by the forward (no block-wide scan, no re-run of earlier chunks); the adjoint recurrence
//     dh_t = C_t dy_t + a_{t+1} dh_{t+1} uses the mirrored combine (shfl_down) seeded by the carried dh of
//     the later chunk.  No cub BlockScan / BlockReverseScan / BlockExchange.
//   * dB / dC are reduced over the CTA's R rows ON CHIP: the rows walk the states in a rotated order
//     (row r handles state (r + k) mod N at step k), so at any step the R rows add into R different state rows
//     of one shared fp32 tile with plain vector read-modify-writes (no atomics); one block barrier per step keeps
//     the rotation aligned.  The tile is flushed once per chunk with red.global.add.v4.f32 -- R times fewer
//     L2 atomics than one atomic per (row, state, t) (the reference issues 2*B*D*N*L scalar atomics).
//   * du, ddelta, dz are written once with 128-bit stores; dA, dD, ddelta_bias are reduced in registers /
//     shared memory over the whole row and hit global memory with ONE atomic per (row, state) per CTA.
#pragma once
#include "fm_common.cuh"
#include "fm_launch.h"

namespace fm {

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <typename T, int S, int G, int NW, bool kHasZ, bool kSmemRed, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB)
scan_bwd_kernel(const FmScanBwdParams q, const int vec_io, const int vec_bc, const int vec_dbc) {
    const FmScanFwdParams& p = q.f;
    constexpr int TC = G * S;
    constexpr int RW = 32 / G;
    constexpr int R = NW * RW;
    constexpr int SP = seg_pad(S);
    constexpr int ROWP = G * SP;
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
    const int seg = lane % G;
    const int rl = warp * RW + lane / G;
    const int dloc = tile * R + rl;
    const bool row_ok = dloc < dg;
    const int d = group * dg + (row_ok ? dloc : 0);

    extern __shared__ __align__(16) float smem[];
    float* sBC = smem;                      // [2 stages][B|C][N][ROWP]
    float* sA = sBC + 4 * N * ROWP;         // [R][N]  A (natural units)
    float* sHs = sA + R * N;                // [R][N]  forward state at chunk start
    float* sDh = sHs + R * N;               // [R][N]  dh at the first step of the later chunk (carried)
    float* sAf = sDh + R * N;               // [R][N]  a of the first step of the later chunk (carried)
    float* sdA = sAf + R * N;               // [N][NT] per-thread dA partials
    float* sdBC = sdA + N * NT;             // [dB|dC][N][ROWP] on-chip reduction tile (kSmemRed only)

    const T* __restrict__ Bg = reinterpret_cast<const T*>(p.B) + b * p.B_batch_stride + group * p.B_group_stride;
    const T* __restrict__ Cg = reinterpret_cast<const T*>(p.C) + b * p.C_batch_stride + group * p.C_group_stride;
    float* __restrict__ dBg = q.dB + b * q.dB_batch_stride + group * q.dB_group_stride;
    float* __restrict__ dCg = q.dC + b * q.dC_batch_stride + group * q.dC_group_stride;
    const T* __restrict__ urow = reinterpret_cast<const T*>(p.u) + b * p.u_batch_stride + d * p.u_d_stride;
    const T* __restrict__ drow = reinterpret_cast<const T*>(p.delta) + b * p.delta_batch_stride + d * p.delta_d_stride;
    const T* __restrict__ gorow = reinterpret_cast<const T*>(q.dout) + b * q.dout_batch_stride + d * q.dout_d_stride;
    T* __restrict__ durow = reinterpret_cast<T*>(q.du) + b * q.du_batch_stride + d * q.du_d_stride;
    T* __restrict__ ddrow = reinterpret_cast<T*>(q.ddelta) + b * q.ddelta_batch_stride + d * q.ddelta_d_stride;
    const T* __restrict__ zrow = nullptr;
    const T* __restrict__ yrow = nullptr;
    T* __restrict__ dzrow = nullptr;
    T* __restrict__ ozrow = nullptr;
    if constexpr (kHasZ) {
        zrow = reinterpret_cast<const T*>(p.z) + b * p.z_batch_stride + d * p.z_d_stride;
        yrow = reinterpret_cast<const T*>(p.out) + b * p.out_batch_stride + d * p.out_d_stride;
        dzrow = reinterpret_cast<T*>(q.dz) + b * q.dz_batch_stride + d * q.dz_d_stride;
        if (p.out_z) ozrow = reinterpret_cast<T*>(p.out_z) + b * p.out_z_batch_stride + d * p.out_z_d_stride;
    }
    const float* __restrict__ hck =
        p.hck ? reinterpret_cast<const float*>(p.hck) + (static_cast<int64_t>(b) * p.dim + d) * p.n_hck * N : nullptr;

    const float Dval = p.D ? reinterpret_cast<const float*>(p.D)[d] : 0.f;
    const float bias = p.delta_bias ? reinterpret_cast<const float*>(p.delta_bias)[d] : 0.f;

    for (int i = tid; i < R * N; i += NT) {
        int r = i / N, n = i % N;
        int dl_ = tile * R + r;
        int dd = group * dg + (dl_ < dg ? dl_ : 0);
        sA[i] = reinterpret_cast<const float*>(p.A)[dd * p.A_d_stride + n * p.A_dstate_stride];
        sDh[i] = 0.f;
        sAf[i] = 1.f;
    }
    for (int i = tid; i < N * NT; i += NT) sdA[i] = 0.f;
    if constexpr (kSmemRed)
        for (int i = tid; i < 2 * N * ROWP; i += NT) sdBC[i] = 0.f;

    const int n_chunks = (L + TC - 1) / TC;
    stage_tile<T, TC, S>(sBC, Bg, p.B_dstate_stride, N, (n_chunks - 1) * TC, L, vec_bc, tid, NT);
    stage_tile<T, TC, S>(sBC + N * ROWP, Cg, p.C_dstate_stride, N, (n_chunks - 1) * TC, L, vec_bc, tid, NT);
    cp_async_commit();

    float dD_acc = 0.f, dbias_acc = 0.f;
    float dfirst_next = 0.f;   // softplus'd delta of the first step of the later chunk
    const float* myA = sA + rl * N;
    const int n_first = kSmemRed ? (rl % N) : 0;   // rotated state order (see header)

    for (int it = 0; it < n_chunks; ++it) {
        const int c = n_chunks - 1 - it;
        const int stage = it & 1;
        if (c > 0) {
            float* nxt = sBC + (stage ^ 1) * 2 * N * ROWP;
            stage_tile<T, TC, S>(nxt, Bg, p.B_dstate_stride, N, (c - 1) * TC, L, vec_bc, tid, NT);
            stage_tile<T, TC, S>(nxt + N * ROWP, Cg, p.C_dstate_stride, N, (c - 1) * TC, L, vec_bc, tid, NT);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        // forward state at the start of this chunk -> sHs (lane seg loads states seg, seg+G, ...)
        {
            const float* hk = (c > 0) ? hck + (c * TC / p.hck_len - 1) * N : nullptr;
            for (int n = seg; n < N; n += G) sHs[rl * N + n] = hk ? hk[n] : 0.f;
        }
        __syncthreads();

        const int t0 = c * TC + seg * S;
        const int nvalid = L - t0;
        constexpr int H = S / 2;                         // time-adjacent element pairs of the lane segment
        float2 dl2[H], du2[H], dy2[H], s12[H], dd2[H];
        {
            float uu[S], dl[S], dy[S];
            load_seg<T, S>(urow + t0, nvalid, vec_io, uu);
            load_seg<T, S>(drow + t0, nvalid, vec_io, dl);
            load_seg<T, S>(gorow + t0, nvalid, vec_io, dy);
            if constexpr (kHasZ) {
                float zv[S], yv[S], dzv[S];
                load_seg<T, S>(zrow + t0, nvalid, vec_io, zv);
                load_seg<T, S>(yrow + t0, nvalid, vec_io, yv);
#pragma unroll
                for (int i = 0; i < S; ++i) {
                    float sg = sigmoid_f(zv[i]);
                    float g = dy[i];
                    dzv[i] = g * yv[i] * sg * (1.f + zv[i] * (1.f - sg));
                    dy[i] = g * zv[i] * sg;
                    yv[i] = yv[i] * zv[i] * sg;          // recomputed out_z
                }
                if (row_ok && nvalid > 0) {
                    store_seg<T, S>(dzrow + t0, nvalid, vec_io, dzv);
                    if (ozrow) store_seg<T, S>(ozrow + t0, nvalid, vec_io, yv);
                }
            }
#pragma unroll
            for (int i = 0; i < S; ++i) {
                float xv = dl[i] + bias;
                float sp = p.delta_softplus ? softplus_fast(xv) : xv;
                sp = (i < nvalid) ? sp : 0.f;            // masked steps: a = 1, b = 0
                dl[i] = sp;
                dD_acc = fmaf(dy[i], uu[i], dD_acc);
            }
#pragma unroll
            for (int j = 0; j < H; ++j) {
                dl2[j] = make_float2(dl[2 * j], dl[2 * j + 1]);
                du2[j] = mul2(dl2[j], make_float2(uu[2 * j], uu[2 * j + 1]));
                dy2[j] = make_float2(dy[2 * j], dy[2 * j + 1]);
                s12[j] = make_float2(0.f, 0.f);
                dd2[j] = make_float2(0.f, 0.f);
            }
        }
        float sumd = 0.f;
#pragma unroll
        for (int j = 0; j < H; ++j) sumd += dl2[j].x + dl2[j].y;
        // shifted sum: sum over the segment of delta_{t+1}
        float dnext0 = __shfl_down_sync(0xffffffffu, dl2[0].x, 1, G);
        if (seg == G - 1) dnext0 = dfirst_next;
        const float sumd_sh = sumd - dl2[0].x + dnext0;
        dfirst_next = __shfl_sync(0xffffffffu, dl2[0].x, 0, G);

        const float* tB = sBC + stage * 2 * N * ROWP + seg * SP;
        const int tCoff = N * ROWP;
        const bool red_vec = vec_dbc && row_ok && nvalid >= S;
        const int rbase = rl * N;

#pragma unroll 1
        for (int k = 0; k < N; ++k) {
            int n = n_first + k;
            if (n >= N) n -= N;
            const int nro = n * ROWP;
            const int rn = rbase + n;
            const float An = sA[rn];
            const float A2 = An * kLog2e;
            const float4* Bv = reinterpret_cast<const float4*>(tB + nro);
            const float4* Cv = reinterpret_cast<const float4*>(tB + nro + tCoff);
            float4* tdB = reinterpret_cast<float4*>(sdBC + nro + seg * SP);
            float4* tdC = reinterpret_cast<float4*>(sdBC + nro + seg * SP + tCoff);
            float* dBp = dBg + n * q.dB_dstate_stride + t0;
            float* dCp = dCg + n * q.dC_dstate_stride + t0;

            float2 a2[H], g2[H];                         // g2 holds b_t first, then g_t = a_t * h_{t-1}
#pragma unroll
            for (int j = 0; j < S / 4; ++j) {
                const float4 v = Bv[j];
                g2[2 * j] = mul2(du2[2 * j], make_float2(v.x, v.y));
                g2[2 * j + 1] = mul2(du2[2 * j + 1], make_float2(v.z, v.w));
            }
#pragma unroll
            for (int j = 0; j < H; ++j) {
                const float2 x2 = mul2(dl2[j], bcast2(A2));
                a2[j].x = ex2_approx(x2.x);
                a2[j].y = ex2_approx(x2.y);
            }
            // ---- forward states of the segment: up-sweep from zero + G-lane combine ------------------------------
            float h = g2[0].x;
            h = fmaf(a2[0].y, h, g2[0].y);
#pragma unroll
            for (int j = 1; j < H; ++j) {
                h = fmaf(a2[j].x, h, g2[j].x);
                h = fmaf(a2[j].y, h, g2[j].y);
            }
            float P = ex2_approx(A2 * sumd);
            const float hstart = sHs[rn];
            if (seg == 0) h = fmaf(P, hstart, h);
#pragma unroll
            for (int o = 1; o < G; o <<= 1) {
                float hp = __shfl_up_sync(0xffffffffu, h, o, G);
                float Pp = 1.f;
                if (2 * o < G) Pp = __shfl_up_sync(0xffffffffu, P, o, G);
                if (seg >= o) {
                    h = fmaf(P, hp, h);
                    if (2 * o < G) P *= Pp;
                }
            }
            float hin = __shfl_up_sync(0xffffffffu, h, 1, G);
            if (seg == 0) hin = hstart;
            // forward down-sweep: g_t = a_t h_{t-1}, h_t = g_t + b_t; dC_t = dy_t h_t goes straight to the tile
            h = hin;
#pragma unroll
            for (int j = 0; j < S / 4; ++j) {
                float2 dc[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int jj = 2 * j + e;
                    const float gx = a2[jj].x * h;
                    const float hx = gx + g2[jj].x;
                    const float gy = a2[jj].y * hx;
                    h = gy + g2[jj].y;
                    g2[jj] = make_float2(gx, gy);
                    dc[e] = mul2(dy2[jj], make_float2(hx, h));
                }
                if constexpr (kSmemRed) {
                    float4 oc = tdC[j];
                    const float2 lo = add2(make_float2(oc.x, oc.y), dc[0]), hi = add2(make_float2(oc.z, oc.w), dc[1]);
                    tdC[j] = make_float4(lo.x, lo.y, hi.x, hi.y);
                } else {
                    if (red_vec) {
                        red_add_v4(dCp + 4 * j, dc[0].x, dc[0].y, dc[1].x, dc[1].y);
                    } else if (row_ok) {
                        const float v4[4] = {dc[0].x, dc[0].y, dc[1].x, dc[1].y};
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (4 * j + e < nvalid) atomicAdd(dCp + 4 * j + e, v4[e]);
                    }
                }
            }
            // ---- adjoint recurrence dh_t = C_t dy_t + a_{t+1} dh_{t+1} -------------------------------------------
            float anext = __shfl_down_sync(0xffffffffu, a2[0].x, 1, G);
            const float dhrun = sDh[rn];
            if (seg == G - 1) anext = sAf[rn];
            float2 cd2[H];
#pragma unroll
            for (int j = 0; j < S / 4; ++j) {
                const float4 v = Cv[j];
                cd2[2 * j] = mul2(make_float2(v.x, v.y), dy2[2 * j]);
                cd2[2 * j + 1] = mul2(make_float2(v.z, v.w), dy2[2 * j + 1]);
            }
            // up-sweep (right to left) from zero: r = dh at the first step given dh_in = 0
            float r = cd2[H - 1].y;
            r = fmaf(a2[H - 1].y, r, cd2[H - 1].x);
#pragma unroll
            for (int j = H - 2; j >= 0; --j) {
                r = fmaf(a2[j + 1].x, r, cd2[j].y);
                r = fmaf(a2[j].y, r, cd2[j].x);
            }
            float Pr = ex2_approx(A2 * sumd_sh);
            if (seg == G - 1) r = fmaf(Pr, dhrun, r);
#pragma unroll
            for (int o = 1; o < G; o <<= 1) {
                float rp = __shfl_down_sync(0xffffffffu, r, o, G);
                float Pp = 1.f;
                if (2 * o < G) Pp = __shfl_down_sync(0xffffffffu, Pr, o, G);
                if (seg + o < G) {
                    r = fmaf(Pr, rp, r);
                    if (2 * o < G) Pr *= Pp;
                }
            }
            float dh = __shfl_down_sync(0xffffffffu, r, 1, G);   // dh at the first step of the next lane
            if (seg == G - 1) dh = dhrun;
            __syncwarp();
            if (seg == 0) {                       // carry to the earlier chunk
                sDh[rn] = r;
                sAf[rn] = a2[0].x;
            }
            // down-sweep (right to left) with the packed gradient products
            float2 dA2 = make_float2(0.f, 0.f);
#pragma unroll
            for (int j = S / 4 - 1; j >= 0; --j) {
                const float4 bv = Bv[j];
                float2 db[2];
#pragma unroll
                for (int e = 1; e >= 0; --e) {
                    const int jj = 2 * j + e;
                    const float an = (jj == H - 1) ? anext : a2[jj + 1].x;
                    const float dhy = fmaf(an, dh, cd2[jj].y);             // dh_{2jj+1}
                    dh = fmaf(a2[jj].y, dhy, cd2[jj].x);                   // dh_{2jj}
                    const float2 dh2 = make_float2(dh, dhy);
                    s12[jj] = fma2(dh2, e ? make_float2(bv.z, bv.w) : make_float2(bv.x, bv.y), s12[jj]);
                    const float2 w2 = mul2(dh2, g2[jj]);                   // dh_t * (h_t - b_t)
                    dd2[jj] = fma2(w2, bcast2(An), dd2[jj]);
                    dA2 = fma2(dl2[jj], w2, dA2);
                    db[e] = mul2(dh2, du2[jj]);
                }
                if constexpr (kSmemRed) {
                    float4 ob = tdB[j];
                    const float2 lo = add2(make_float2(ob.x, ob.y), db[0]), hi = add2(make_float2(ob.z, ob.w), db[1]);
                    tdB[j] = make_float4(lo.x, lo.y, hi.x, hi.y);
                } else {
                    if (red_vec) {
                        red_add_v4(dBp + 4 * j, db[0].x, db[0].y, db[1].x, db[1].y);
                    } else if (row_ok) {
                        const float v4[4] = {db[0].x, db[0].y, db[1].x, db[1].y};
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (4 * j + e < nvalid) atomicAdd(dBp + 4 * j + e, v4[e]);
                    }
                }
            }
            sdA[n * NT + tid] += dA2.x + dA2.y;
            if constexpr (kSmemRed) __syncthreads();   // keep the row rotation aligned (one step = one state row per CTA row)
        }

        // per-element outputs
        {
            float ov[S], uu[S], dr[S];
            load_seg<T, S>(urow + t0, nvalid, vec_io, uu);       // re-read (L1/L2 hit) instead of holding 2*S registers
            load_seg<T, S>(drow + t0, nvalid, vec_io, dr);
#pragma unroll
            for (int j = 0; j < H; ++j) {
                const float2 o2 = fma2(dl2[j], s12[j], mul2(dy2[j], bcast2(Dval)));
                ov[2 * j] = o2.x;
                ov[2 * j + 1] = o2.y;
            }
            if (row_ok && nvalid > 0) store_seg<T, S>(durow + t0, nvalid, vec_io, ov);
#pragma unroll
            for (int j = 0; j < H; ++j) {
                const float2 g2_ = fma2(make_float2(uu[2 * j], uu[2 * j + 1]), s12[j], dd2[j]);
                ov[2 * j] = g2_.x;
                ov[2 * j + 1] = g2_.y;
            }
#pragma unroll
            for (int i = 0; i < S; ++i) {
                float g = ov[i];
                if (p.delta_softplus) g *= sigmoid_f(dr[i] + bias);   // d softplus(x) / dx
                g = (i < nvalid) ? g : 0.f;
                ov[i] = g;
                dbias_acc += g;
            }
            if (row_ok && nvalid > 0) store_seg<T, S>(ddrow + t0, nvalid, vec_io, ov);
        }

        if constexpr (kSmemRed) {
            // flush the CTA's dB/dC tile: one vector red per 4 timesteps per state, then clear it for the next chunk
            constexpr int QPR = TC / 4;
            const int tc0 = c * TC;
            for (int s = tid; s < 2 * N * QPR; s += NT) {
                const int which = s / (N * QPR);
                const int rem = s % (N * QPR);
                const int n = rem / QPR, qq = rem % QPR;
                float4* src = reinterpret_cast<float4*>(sdBC + which * N * ROWP + n * ROWP + (qq / (S / 4)) * SP + (qq % (S / 4)) * 4);
                float4 v = *src;
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
        } else {
            __syncthreads();
        }
    }

    // row reductions -> one atomic per (row, state) / row
#pragma unroll
    for (int o = 1; o < G; o <<= 1) {
        dD_acc += __shfl_xor_sync(0xffffffffu, dD_acc, o, G);
        dbias_acc += __shfl_xor_sync(0xffffffffu, dbias_acc, o, G);
    }
    if (seg == 0 && row_ok) {
        if (q.dD) atomicAdd(q.dD + d, dD_acc);
        if (q.ddelta_bias) atomicAdd(q.ddelta_bias + d, dbias_acc);
    }
    for (int n = 0; n < N; ++n) {
        float v = sdA[n * NT + tid];
#pragma unroll
        for (int o = 1; o < G; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o, G);
        if (seg == 0 && row_ok) atomicAdd(q.dA + static_cast<int64_t>(d) * N + n, v);
    }
}

template <int S, int G, int NW>
constexpr size_t bwd_smem_bytes(int dstate, bool smem_red) {
    return sizeof(float) * (4 * (size_t)dstate * G * seg_pad(S) + 4 * (size_t)NW * (32 / G) * dstate +
                            (size_t)dstate * NW * 32 + (smem_red ? 2 * (size_t)dstate * G * seg_pad(S) : 0));
}

template <typename T, int S, int G, int NW, int MINB>
static cudaError_t launch_bwd_cfg(const FmScanBwdParams& q, cudaStream_t st, int vec_io, int vec_bc, int vec_dbc) {
    const FmScanFwdParams& p = q.f;
    constexpr int RW = 32 / G, R = NW * RW, NT = NW * 32;
    const int dg = p.dim / p.n_groups;
    const int tiles = (dg + R - 1) / R;
    dim3 grid(tiles * p.n_groups, p.batch);
    // on-chip dB/dC reduction needs distinct state rows for the CTA's rows at every step (R <= dstate) and no
    // shadow rows (the shadow rows of a partial tile would double count row 0 of the group)
    const bool smem_red = (R <= p.dstate) && (dg % R == 0) && env_int("FM_SCAN_BWD_SMEMRED", 1) != 0;
    const size_t smem = bwd_smem_bytes<S, G, NW>(p.dstate, smem_red);
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    void (*kern)(const FmScanBwdParams, int, int, int);
    if (p.z) kern = smem_red ? scan_bwd_kernel<T, S, G, NW, true, true, MINB> : scan_bwd_kernel<T, S, G, NW, true, false, MINB>;
    else kern = smem_red ? scan_bwd_kernel<T, S, G, NW, false, true, MINB> : scan_bwd_kernel<T, S, G, NW, false, false, MINB>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, NT, smem, st>>>(q, vec_io, vec_bc, vec_dbc);
    count_launch();
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_scan_bwd_rp_T(const FmScanBwdParams& q, cudaStream_t st, int vec_io, int vec_bc, int vec_dbc);   // fm_scan_bwd_rp.cuh
template <typename T>
cudaError_t launch_scan_bwd_ls_T(const FmScanBwdParams& q, cudaStream_t st, int vec_bc, int vec_dbc);               // fm_scan_bwd_ls.cuh
template <typename T>
cudaError_t launch_scan_bwd_ls2_T(const FmScanBwdParams& q, cudaStream_t st, int vec_bc, int vec_dbc);              // fm_scan_bwd_ls2.cuh

template <typename T>
cudaError_t launch_scan_bwd_T(const FmScanBwdParams& q, cudaStream_t st) {
    const FmScanFwdParams& p = q.f;
    const int64_t al = 16 / (int)sizeof(T);
    auto ok = [&](const void* ptr, int64_t s0, int64_t s1) { return aligned16(ptr) && s0 % al == 0 && s1 % al == 0; };
    int vec_io = ok(p.u, p.u_batch_stride, p.u_d_stride) && ok(p.delta, p.delta_batch_stride, p.delta_d_stride) &&
                 ok(q.dout, q.dout_batch_stride, q.dout_d_stride) && ok(q.du, q.du_batch_stride, q.du_d_stride) &&
                 ok(q.ddelta, q.ddelta_batch_stride, q.ddelta_d_stride);
    if (p.z) {
        vec_io = vec_io && ok(p.z, p.z_batch_stride, p.z_d_stride) && ok(p.out, p.out_batch_stride, p.out_d_stride) &&
                 ok(q.dz, q.dz_batch_stride, q.dz_d_stride);
        if (p.out_z) vec_io = vec_io && ok(p.out_z, p.out_z_batch_stride, p.out_z_d_stride);
    }
    int vec_bc = ok(p.B, p.B_batch_stride, p.B_group_stride) && p.B_dstate_stride % al == 0 &&
                 ok(p.C, p.C_batch_stride, p.C_group_stride) && p.C_dstate_stride % al == 0;
    auto ok4 = [&](const void* ptr, int64_t s0, int64_t s1, int64_t s2) {
        return aligned16(ptr) && s0 % 4 == 0 && s1 % 4 == 0 && s2 % 4 == 0;
    };
    int vec_dbc = ok4(q.dB, q.dB_batch_stride, q.dB_group_stride, q.dB_dstate_stride) &&
                  ok4(q.dC, q.dC_batch_stride, q.dC_group_stride, q.dC_dstate_stride);

    // dstate == 16 without z (every SS2D scan of the model) on the dense 8-step checkpoints: the software-pipelined lane-serial kernel
    // (fm_scan_bwd_ls2.cuh: 16-byte-aligned B / C / dB / dC rows, whole 16-byte chunks), else the first one (fm_scan_bwd_ls.cuh)
    // (sequences under 128 steps: the pipeline fill / drain of the second kernel costs more than its trips save,
    //  profiles/r02_ls2_ab.jsonl)
    if (env_int("FM_SCAN_BWD_LS2", 1) != 0 && p.seqlen >= env_int("FM_SCAN_BWD_LS2_MINL", 128)) {
        const cudaError_t e = launch_scan_bwd_ls2_T<T>(q, st, vec_bc, vec_dbc);
        if (e != cudaErrorInvalidConfiguration) return e;
    }
    if (env_int("FM_SCAN_BWD_LS", 1) != 0) {
        const cudaError_t e = launch_scan_bwd_ls_T<T>(q, st, vec_bc, vec_dbc);
        if (e != cudaErrorInvalidConfiguration) return e;
    }
    // otherwise: row-pair kernel (fm_scan_bwd_rp.cuh); shapes outside its preconditions use the generic kernel below
    if (env_int("FM_SCAN_BWD_RP", 1) != 0) {
        const cudaError_t e = launch_scan_bwd_rp_T<T>(q, st, vec_io, vec_bc, vec_dbc);
        if (e != cudaErrorInvalidConfiguration) return e;
    }

    constexpr int S = 8;
    int G = scan_lanes_per_row((int64_t)p.batch * p.dim, p.seqlen, S, "FM_SCAN_BWD_G");
    int NW = env_int("FM_SCAN_BWD_NW", 8);
    int Gmin = 1;
    if (!p.hck) {
        // no dense checkpoints: the whole sequence must fit one chunk (checked by the C ABI: seqlen <= 256)
        G = 1;
        while (G < 32 && G * S < p.seqlen) G <<= 1;
        Gmin = G;
    } else if (G * S < p.seqlen) {
        // a multi-chunk backward needs the chunk length to be a multiple of the checkpoint spacing
        while (G < 32 && (G * S) % p.hck_len != 0) G <<= 1;
        Gmin = p.hck_len / S > 0 ? p.hck_len / S : 1;
    }
    // prefer a CTA row count that allows the on-chip dB/dC reduction (R = NW*32/G <= dstate)
    while (NW > 4 && NW * (32 / G) > p.dstate) NW >>= 1;
    // shared-memory budget (B/C ring + reduction tile grow with dstate * G): shrink G, then NW
    auto smem_need = [&](int g, int nw) {
        const bool red = (nw * (32 / g) <= p.dstate) && ((p.dim / p.n_groups) % (nw * (32 / g)) == 0);
        return sizeof(float) * ((4 + (red ? 2 : 0)) * (size_t)p.dstate * g * seg_pad(S) + 4 * (size_t)nw * (32 / g) * p.dstate +
                                (size_t)p.dstate * nw * 32);
    };
    while (smem_need(G, NW) > 200 * 1024 && G > Gmin) G >>= 1;
    if (smem_need(G, NW) > 200 * 1024) NW = 4;
    if (G < 8 && NW > 4) NW = 4;   // only (G >= 8, NW = 8) instances exist; short sequences / wide states take 4-warp CTAs
    // wide states (dstate > 64): the per-row state arrays (NW * 32/G rows x dstate) and the B/C tile (dstate x G) pull in opposite
    // directions; the few-lane small-CTA instances below are the ones that fit 227 KB up to dstate 256
    if (smem_need(G, NW) > 220 * 1024) {
        bool found = false;
        for (int nw = 2; nw >= 1 && !found; nw >>= 1)
            for (int g = 2; g >= 1 && !found; g >>= 1) {
                const bool g_ok = g * S >= p.seqlen || (p.hck && (g * S) % p.hck_len == 0);
                if (g_ok && !(g == 1 && nw == 2) && smem_need(g, nw) <= 220 * 1024) { G = g; NW = nw; found = true; }
            }
        if (!found) return cudaErrorInvalidConfiguration;
    }
#define FM_CASE(g, nw, minb) if (G == g && NW == nw) return launch_bwd_cfg<T, S, g, nw, minb>(q, st, vec_io, vec_bc, vec_dbc);
    FM_CASE(1, 4, 4) FM_CASE(2, 4, 4) FM_CASE(4, 4, 4) FM_CASE(8, 4, 4) FM_CASE(16, 4, 4) FM_CASE(32, 4, 4)
    FM_CASE(8, 8, 2) FM_CASE(16, 8, 2) FM_CASE(32, 8, 2)
    FM_CASE(2, 2, 4) FM_CASE(2, 1, 4) FM_CASE(1, 1, 4)
#undef FM_CASE
    return cudaErrorInvalidConfiguration;
}

}  // namespace fm

#include "fm_scan_bwd_rp.cuh"
#include "fm_scan_bwd_ls.cuh"
#include "fm_scan_bwd_ls2.cuh"
