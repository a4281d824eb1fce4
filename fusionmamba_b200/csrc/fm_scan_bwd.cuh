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
//   * du, ddelta, dz are written once with 128-bit stores; dA, dD, ddelta_bias are reduced in registers /
//     shared memory over the whole row and hit global memory with ONE atomic per (row, state) per CTA.
//   * dB / dC: per-lane 16-step segments are added to the fp32 accumulators with vector red.global.add.v4.f32
//     (4 per lane per state instead of 16 scalar atomics).
#pragma once
#include "fm_common.cuh"
#include "fm_launch.h"

namespace fm {

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <typename T, int G, int NW, bool kHasZ>
__global__ void __launch_bounds__(NW * 32)
scan_bwd_kernel(const FmScanBwdParams q, const int vec_io, const int vec_bc, const int vec_dbc) {
    const FmScanFwdParams& p = q.f;
    constexpr int TC = G * kSeg;
    constexpr int RW = 32 / G;
    constexpr int R = NW * RW;
    constexpr int ROWP = G * kSegPad;
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

    const int n_chunks = (L + TC - 1) / TC;
    stage_tile<T, TC>(sBC, Bg, p.B_dstate_stride, N, (n_chunks - 1) * TC, L, vec_bc, tid, NT);
    stage_tile<T, TC>(sBC + N * ROWP, Cg, p.C_dstate_stride, N, (n_chunks - 1) * TC, L, vec_bc, tid, NT);
    cp_async_commit();

    float dD_acc = 0.f, dbias_acc = 0.f;
    float dfirst_next = 0.f;   // softplus'd delta of the first step of the later chunk (valid on every lane of the row)

    for (int it = 0; it < n_chunks; ++it) {
        const int c = n_chunks - 1 - it;
        const int stage = it & 1;
        if (c > 0) {
            float* nxt = sBC + (stage ^ 1) * 2 * N * ROWP;
            stage_tile<T, TC>(nxt, Bg, p.B_dstate_stride, N, (c - 1) * TC, L, vec_bc, tid, NT);
            stage_tile<T, TC>(nxt + N * ROWP, Cg, p.C_dstate_stride, N, (c - 1) * TC, L, vec_bc, tid, NT);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        // forward state at the start of this chunk -> sHs (lane seg loads states seg, seg+G, ...)
        for (int n = seg; n < N; n += G)
            sHs[rl * N + n] = (c > 0) ? hck[(c * TC / p.hck_len - 1) * N + n] : 0.f;
        __syncthreads();

        const int t0 = c * TC + seg * kSeg;
        const int nvalid = L - t0;
        float dl[kSeg], uu[kSeg], du_[kSeg], dy[kSeg], s1[kSeg], dd[kSeg];
        load_seg<T>(urow + t0, nvalid, vec_io, uu);
        load_seg<T>(drow + t0, nvalid, vec_io, dl);
        load_seg<T>(gorow + t0, nvalid, vec_io, dy);
        if constexpr (kHasZ) {
            float zv[kSeg], yv[kSeg];
            load_seg<T>(zrow + t0, nvalid, vec_io, zv);
            load_seg<T>(yrow + t0, nvalid, vec_io, yv);
            float dzv[kSeg], ozv[kSeg];
#pragma unroll
            for (int i = 0; i < kSeg; ++i) {
                float sg = sigmoid_f(zv[i]);
                float g = dy[i];
                dzv[i] = g * yv[i] * sg * (1.f + zv[i] * (1.f - sg));
                ozv[i] = yv[i] * zv[i] * sg;
                dy[i] = g * zv[i] * sg;
            }
            if (row_ok && nvalid > 0) {
                store_seg<T>(dzrow + t0, nvalid, vec_io, dzv);
                if (ozrow) store_seg<T>(ozrow + t0, nvalid, vec_io, ozv);
            }
        }
        float sumd = 0.f;
#pragma unroll
        for (int i = 0; i < kSeg; ++i) {
            float xv = dl[i] + bias;
            float sp = p.delta_softplus ? softplus_ref(xv) : xv;
            sp = (i < nvalid) ? sp : 0.f;
            dl[i] = sp;
            du_[i] = sp * uu[i];
            s1[i] = 0.f;
            dd[i] = 0.f;
            dD_acc = fmaf(dy[i], uu[i], dD_acc);
            sumd += sp;
        }
        // shifted sum: sum over the segment of delta_{t+1}
        float dnext0 = __shfl_down_sync(0xffffffffu, dl[0], 1, G);
        if (seg == G - 1) dnext0 = dfirst_next;
        const float sumd_sh = sumd - dl[0] + dnext0;
        dfirst_next = __shfl_sync(0xffffffffu, dl[0], 0, G);

        const float* tB = sBC + stage * 2 * N * ROWP + seg * kSegPad;
        const float* tC = tB + N * ROWP;
        const bool red_vec = vec_dbc && row_ok && nvalid >= kSeg;

#pragma unroll 1
        for (int n = 0; n < N; ++n) {
            const float An = sA[rl * N + n];
            const float A2 = An * kLog2e;
            float a[kSeg], hs[kSeg];
            const float4* Bv = reinterpret_cast<const float4*>(tB + n * ROWP);
            const float4* Cv = reinterpret_cast<const float4*>(tC + n * ROWP);
#pragma unroll
            for (int k = 0; k < kSeg / 4; ++k) {
                float4 v = Bv[k];
                hs[4 * k + 0] = du_[4 * k + 0] * v.x;
                hs[4 * k + 1] = du_[4 * k + 1] * v.y;
                hs[4 * k + 2] = du_[4 * k + 2] * v.z;
                hs[4 * k + 3] = du_[4 * k + 3] * v.w;
            }
#pragma unroll
            for (int i = 0; i < kSeg; ++i) a[i] = ex2_approx(dl[i] * A2);
            // ---- forward states of the segment -------------------------------------------------
            float h = 0.f;
#pragma unroll
            for (int i = 0; i < kSeg; ++i) h = fmaf(a[i], h, hs[i]);
            float P = ex2_approx(A2 * sumd);
            const float hstart = sHs[rl * N + n];
            if (seg == 0) h = fmaf(P, hstart, h);
#pragma unroll
            for (int o = 1; o < G; o <<= 1) {
                float Pp = __shfl_up_sync(0xffffffffu, P, o, G);
                float hp = __shfl_up_sync(0xffffffffu, h, o, G);
                if (seg >= o) {
                    h = fmaf(P, hp, h);
                    P *= Pp;
                }
            }
            float hin = __shfl_up_sync(0xffffffffu, h, 1, G);
            if (seg == 0) hin = hstart;
            h = hin;
#pragma unroll
            for (int i = 0; i < kSeg; ++i) {
                h = fmaf(a[i], h, hs[i]);
                hs[i] = h;                       // hs[i] = h_t
            }
            // ---- adjoint recurrence ------------------------------------------------------------
            float anext = __shfl_down_sync(0xffffffffu, a[0], 1, G);
            if (seg == G - 1) anext = sAf[rl * N + n];
            // up-sweep (right to left) from zero: r = dh at the first step given dh_in = 0
            float r = 0.f;
            {
                float4 cv3 = Cv[3], cv2 = Cv[2], cv1 = Cv[1], cv0 = Cv[0];
                const float cd[kSeg] = {cv0.x, cv0.y, cv0.z, cv0.w, cv1.x, cv1.y, cv1.z, cv1.w,
                                        cv2.x, cv2.y, cv2.z, cv2.w, cv3.x, cv3.y, cv3.z, cv3.w};
                r = cd[kSeg - 1] * dy[kSeg - 1];
#pragma unroll
                for (int i = kSeg - 2; i >= 0; --i) r = fmaf(a[i + 1], r, cd[i] * dy[i]);
            }
            float Pr = ex2_approx(A2 * sumd_sh);
            const float dhrun = sDh[rl * N + n];
            if (seg == G - 1) r = fmaf(Pr, dhrun, r);
#pragma unroll
            for (int o = 1; o < G; o <<= 1) {
                float Pp = __shfl_down_sync(0xffffffffu, Pr, o, G);
                float rp = __shfl_down_sync(0xffffffffu, r, o, G);
                if (seg + o < G) {
                    r = fmaf(Pr, rp, r);
                    Pr *= Pp;
                }
            }
            float dh = __shfl_down_sync(0xffffffffu, r, 1, G);   // dh at the first step of the next lane
            if (seg == G - 1) dh = dhrun;
            __syncwarp();
            if (seg == 0) {                       // carry to the earlier chunk
                sDh[rl * N + n] = r;
                sAf[rl * N + n] = a[0];
            }
            // down-sweep (right to left) with gradient products
            float dA_part = 0.f;
            float* dBp = dBg + n * q.dB_dstate_stride + t0;
            float* dCp = dCg + n * q.dC_dstate_stride + t0;
#pragma unroll
            for (int k = kSeg / 4 - 1; k >= 0; --k) {
                float4 bv = Bv[k], cv = Cv[k];
                const float bq[4] = {bv.x, bv.y, bv.z, bv.w};
                const float cq[4] = {cv.x, cv.y, cv.z, cv.w};
                float dbq[4], dcq[4];
#pragma unroll
                for (int j = 3; j >= 0; --j) {
                    const int i = 4 * k + j;
                    const float an = (i == kSeg - 1) ? anext : a[i + 1];
                    dh = fmaf(an, dh, cq[j] * dy[i]);            // dh_t
                    s1[i] = fmaf(dh, bq[j], s1[i]);
                    const float hp = (i == 0) ? hin : hs[i - 1];
                    const float w = dh * (a[i] * hp);            // dh_t * (h_t - b_t)
                    dd[i] = fmaf(An, w, dd[i]);
                    dA_part = fmaf(dl[i], w, dA_part);
                    dbq[j] = dh * du_[i];
                    dcq[j] = dy[i] * hs[i];
                }
                if (red_vec) {
                    red_add_v4(dBp + 4 * k, dbq[0], dbq[1], dbq[2], dbq[3]);
                    red_add_v4(dCp + 4 * k, dcq[0], dcq[1], dcq[2], dcq[3]);
                } else if (row_ok) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (4 * k + j < nvalid) {
                            atomicAdd(dBp + 4 * k + j, dbq[j]);
                            atomicAdd(dCp + 4 * k + j, dcq[j]);
                        }
                }
            }
            sdA[n * NT + tid] += dA_part;
        }

        // per-element outputs
        {
            float ov[kSeg];
#pragma unroll
            for (int i = 0; i < kSeg; ++i) ov[i] = fmaf(dl[i], s1[i], Dval * dy[i]);
            if (row_ok && nvalid > 0) store_seg<T>(durow + t0, nvalid, vec_io, ov);
#pragma unroll
            for (int i = 0; i < kSeg; ++i) {
                float g = fmaf(uu[i], s1[i], dd[i]);
                if (p.delta_softplus) g *= -expm1f(-dl[i]);      // sigmoid(x) = 1 - exp(-softplus(x))
                g = (i < nvalid) ? g : 0.f;
                ov[i] = g;
                dbias_acc += g;
            }
            if (row_ok && nvalid > 0) store_seg<T>(ddrow + t0, nvalid, vec_io, ov);
        }
        __syncthreads();
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

template <typename T, int G, int NW>
static cudaError_t launch_bwd_cfg(const FmScanBwdParams& q, cudaStream_t st, int vec_io, int vec_bc, int vec_dbc) {
    const FmScanFwdParams& p = q.f;
    constexpr int RW = 32 / G, R = NW * RW, ROWP = G * kSegPad, NT = NW * 32;
    const int dg = p.dim / p.n_groups;
    const int tiles = (dg + R - 1) / R;
    dim3 grid(tiles * p.n_groups, p.batch);
    size_t smem = sizeof(float) * (4 * (size_t)p.dstate * ROWP + 4 * (size_t)R * p.dstate + (size_t)p.dstate * NT);
    auto kern = p.z ? scan_bwd_kernel<T, G, NW, true> : scan_bwd_kernel<T, G, NW, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, NT, smem, st>>>(q, vec_io, vec_bc, vec_dbc);
    count_launch();
    return cudaGetLastError();
}

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

    int G = scan_lanes_per_row((int64_t)p.batch * p.dim, p.seqlen, p.dstate, "FM_SCAN_BWD_G");
    if (!p.hck) {
        // no dense checkpoints: the whole sequence must fit one chunk (checked by the C ABI: seqlen <= 512)
        G = 1;
        while (G < 32 && G * kSeg < p.seqlen) G <<= 1;
    } else if (G * kSeg < p.seqlen) {
        // a multi-chunk backward needs the chunk length to be a multiple of the checkpoint spacing
        while (G < 32 && (G * kSeg) % p.hck_len != 0) G <<= 1;
    }
    int NW = env_int("FM_SCAN_BWD_NW", 4);
#define FM_CASE(g, nw) if (G == g && NW == nw) return launch_bwd_cfg<T, g, nw>(q, st, vec_io, vec_bc, vec_dbc);
    FM_CASE(1, 4) FM_CASE(2, 4) FM_CASE(4, 4) FM_CASE(8, 4) FM_CASE(16, 4) FM_CASE(32, 4)
    FM_CASE(8, 8) FM_CASE(16, 8) FM_CASE(8, 2) FM_CASE(16, 2)
#undef FM_CASE
    return launch_bwd_cfg<T, 8, 4>(q, st, vec_io, vec_bc, vec_dbc);
}

}  // namespace fm
