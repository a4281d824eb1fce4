// fm_scan_fwd.cuh -- selective-scan forward for sm_100a.
//
// Replaces selective_scan_fwd_kernel (selective_scan/selective_scan_fwd_kernel.cuh:67-303) with a different
// decomposition (this is not a port):
//   * one CTA owns R = NW*(32/G) channel rows of ONE (batch, group) and walks the sequence in chunks of
//     TC = 16*G timesteps; the [dstate x TC] B and C tiles are staged ONCE per chunk in shared memory
//     (cp.async double buffer, lane-segment-padded layout) and shared by all R rows -- the reference
//     re-reads them from L2 for every row.
//   * a row is scanned by G lanes of one warp; each lane owns 16 consecutive timesteps: thread-serial
//     recurrence from zero (up-sweep), G-lane warp-shuffle combine of (decay, state) aggregates with the
//     monoid (a0,b0)o(a1,b1) = (a1*a0, a1*b0+b1), then a second serial pass seeded with the lane's incoming
//     state that also accumulates y += C*h.  a_t is computed once (one ex2 per (t, state)) and kept in
//     registers between the two passes; a segment's aggregate decay is exp2(A * sum(delta)) (one ex2 per lane).
//   * the running state is carried across chunks per (row, state) in shared memory; no block-wide barrier
//     is needed inside the state loop (warps are independent), only two per chunk for the B/C tile ring.
//   * u / delta / z / out move as 128-bit vector accesses, 64 contiguous bytes per lane.
#include "fm_common.cuh"
#include "fm_launch.h"

#pragma once
namespace fm {

template <typename T, int G, int NW, bool kHasZ>
__global__ void __launch_bounds__(NW * 32)
scan_fwd_kernel(const FmScanFwdParams p, const int vec_io, const int vec_bc) {
    constexpr int TC = G * kSeg;            // timesteps per chunk
    constexpr int RW = 32 / G;              // rows per warp
    constexpr int R = NW * RW;              // rows per CTA
    constexpr int ROWP = G * kSegPad;       // smem pitch of one state row of the B/C tile (floats)
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
    const int seg = lane % G;                           // which 16-step segment of the chunk
    const int rl = warp * RW + lane / G;                // row within CTA
    const int dloc = tile * R + rl;                     // channel within group
    const bool row_ok = dloc < dg;
    const int d = group * dg + (row_ok ? dloc : 0);     // clamp: invalid rows compute on row 0 of the group, never store

    extern __shared__ __align__(16) float smem[];
    float* sBC = smem;                                  // [2 stages][B|C][N][ROWP]
    float* sA2 = sBC + 4 * N * ROWP;                    // [R][N]  A * log2(e)
    float* sH = sA2 + R * N;                            // [R][N]  running state (owned by the seg==0 lane of the row)

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
    float* __restrict__ xrow = reinterpret_cast<float*>(p.x) + (static_cast<int64_t>(b) * p.dim + d) * p.n_chunks * 2 * N;

    const float Dval = p.D ? reinterpret_cast<const float*>(p.D)[d] : 0.f;
    const float bias = p.delta_bias ? reinterpret_cast<const float*>(p.delta_bias)[d] : 0.f;

    // per-row constants -> smem
    for (int i = tid; i < R * N; i += NT) {
        int r = i / N, n = i % N;
        int dl = tile * R + r;
        int dd = group * dg + (dl < dg ? dl : 0);
        sA2[i] = reinterpret_cast<const float*>(p.A)[dd * p.A_d_stride + n * p.A_dstate_stride] * kLog2e;
        sH[i] = 0.f;
    }

    const int n_chunks = (L + TC - 1) / TC;
    // prologue: stage chunk 0
    stage_tile<T, TC>(sBC, Bg, p.B_dstate_stride, N, 0, L, vec_bc, tid, NT);
    stage_tile<T, TC>(sBC + N * ROWP, Cg, p.C_dstate_stride, N, 0, L, vec_bc, tid, NT);
    cp_async_commit();

    float sum_total = 0.f;  // running sum of delta over the row (for the decay stored in x)

    for (int c = 0; c < n_chunks; ++c) {
        const int stage = c & 1;
        if (c + 1 < n_chunks) {
            float* nxt = sBC + (stage ^ 1) * 2 * N * ROWP;
            stage_tile<T, TC>(nxt, Bg, p.B_dstate_stride, N, (c + 1) * TC, L, vec_bc, tid, NT);
            stage_tile<T, TC>(nxt + N * ROWP, Cg, p.C_dstate_stride, N, (c + 1) * TC, L, vec_bc, tid, NT);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();

        const int t0 = c * TC + seg * kSeg;
        const int nvalid = L - t0;                      // may be <= 0 or > 16
        float dl[kSeg], du[kSeg], y[kSeg];
        {
            float uv[kSeg];
            load_seg<T>(urow + t0, nvalid, vec_io, uv);
            load_seg<T>(drow + t0, nvalid, vec_io, dl);
#pragma unroll
            for (int i = 0; i < kSeg; ++i) {
                float xv = dl[i] + bias;
                float sp = p.delta_softplus ? softplus_ref(xv) : xv;
                sp = (i < nvalid) ? sp : 0.f;           // masked steps: a = 1, b = 0
                dl[i] = sp;
                du[i] = sp * uv[i];
                y[i] = Dval * uv[i];
            }
        }
        float sumd = 0.f;
#pragma unroll
        for (int i = 0; i < kSeg; ++i) sumd += dl[i];
        // row total of this chunk (for the stored decay product only)
        float rowsum = sumd;
#pragma unroll
        for (int o = 1; o < G; o <<= 1) rowsum += __shfl_xor_sync(0xffffffffu, rowsum, o, G);
        sum_total += rowsum;

        const int t_end = min((c + 1) * TC, L);         // exclusive end of this chunk
        const bool ckpt = (t_end % p.chunk_len == 0) || (t_end == L);
        const int slot = (t_end - 1) / p.chunk_len;

        const float* tB = sBC + stage * 2 * N * ROWP + seg * kSegPad;
        const float* tC = tB + N * ROWP;

#pragma unroll 1
        for (int n = 0; n < N; ++n) {
            const float A2 = sA2[rl * N + n];
            float a[kSeg], bb[kSeg];
            const float4* Bv = reinterpret_cast<const float4*>(tB + n * ROWP);
#pragma unroll
            for (int q = 0; q < kSeg / 4; ++q) {
                float4 v = Bv[q];
                bb[4 * q + 0] = du[4 * q + 0] * v.x;
                bb[4 * q + 1] = du[4 * q + 1] * v.y;
                bb[4 * q + 2] = du[4 * q + 2] * v.z;
                bb[4 * q + 3] = du[4 * q + 3] * v.w;
            }
#pragma unroll
            for (int i = 0; i < kSeg; ++i) a[i] = ex2_approx(dl[i] * A2);
            // up-sweep: segment state from zero
            float h = 0.f;
#pragma unroll
            for (int i = 0; i < kSeg; ++i) h = fmaf(a[i], h, bb[i]);
            float P = ex2_approx(A2 * sumd);
            const float hrun = sH[rl * N + n];
            if (seg == 0) h = fmaf(P, hrun, h);
            // inclusive combine over the G lanes of the row
#pragma unroll
            for (int o = 1; o < G; o <<= 1) {
                float Pp = __shfl_up_sync(0xffffffffu, P, o, G);
                float hp = __shfl_up_sync(0xffffffffu, h, o, G);
                if (seg >= o) {
                    h = fmaf(P, hp, h);
                    P *= Pp;
                }
            }
            if (p.hck != nullptr && row_ok) {   // dense checkpoint: state at the end of this lane's segment
                const int te = t0 + kSeg;
                if (te < L && te % p.hck_len == 0)
                    reinterpret_cast<float*>(p.hck)[((static_cast<int64_t>(b) * p.dim + d) * p.n_hck + (te / p.hck_len - 1)) * N + n] = h;
            }
            float hin = __shfl_up_sync(0xffffffffu, h, 1, G);
            if (seg == 0) hin = hrun;
            const float hlast = __shfl_sync(0xffffffffu, h, G - 1, G);
            if (seg == 0) {
                sH[rl * N + n] = hlast;
                if (ckpt && row_ok) {
                    xrow[slot * 2 * N + 2 * n] = ex2_approx(A2 * sum_total);
                    xrow[slot * 2 * N + 2 * n + 1] = hlast;
                }
            }
            // down-sweep with the true incoming state; y += C * h
            const float4* Cv = reinterpret_cast<const float4*>(tC + n * ROWP);
            h = hin;
#pragma unroll
            for (int q = 0; q < kSeg / 4; ++q) {
                float4 v = Cv[q];
                h = fmaf(a[4 * q + 0], h, bb[4 * q + 0]); y[4 * q + 0] = fmaf(v.x, h, y[4 * q + 0]);
                h = fmaf(a[4 * q + 1], h, bb[4 * q + 1]); y[4 * q + 1] = fmaf(v.y, h, y[4 * q + 1]);
                h = fmaf(a[4 * q + 2], h, bb[4 * q + 2]); y[4 * q + 2] = fmaf(v.z, h, y[4 * q + 2]);
                h = fmaf(a[4 * q + 3], h, bb[4 * q + 3]); y[4 * q + 3] = fmaf(v.w, h, y[4 * q + 3]);
            }
        }

        if (row_ok && nvalid > 0) {
            store_seg<T>(orow + t0, nvalid, vec_io, y);
            if constexpr (kHasZ) {
                float zv[kSeg];
                load_seg<T>(zrow + t0, nvalid, vec_io, zv);
#pragma unroll
                for (int i = 0; i < kSeg; ++i) y[i] = y[i] * zv[i] * sigmoid_f(zv[i]);
                store_seg<T>(ozrow + t0, nvalid, vec_io, y);
            }
        }
        __syncthreads();  // all warps done with this stage before it is refilled (chunk c+2)
    }
}

// ---------------------------------------------------------------------------------------------
// host launcher
// ---------------------------------------------------------------------------------------------
template <typename T, int G, int NW>
static cudaError_t launch_cfg(const FmScanFwdParams& p, cudaStream_t st, int vec_io, int vec_bc) {
    constexpr int RW = 32 / G, R = NW * RW, ROWP = G * kSegPad;
    const int dg = p.dim / p.n_groups;
    const int tiles = (dg + R - 1) / R;
    dim3 grid(tiles * p.n_groups, p.batch);
    size_t smem = sizeof(float) * (4 * (size_t)p.dstate * ROWP + 2 * (size_t)R * p.dstate);
    auto kern = p.z ? scan_fwd_kernel<T, G, NW, true> : scan_fwd_kernel<T, G, NW, false>;
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
                 ok(p.out, p.out_batch_stride, p.out_d_stride);
    if (p.z) vec_io = vec_io && ok(p.z, p.z_batch_stride, p.z_d_stride) && ok(p.out_z, p.out_z_batch_stride, p.out_z_d_stride);
    int vec_bc = ok(p.B, p.B_batch_stride, p.B_group_stride) && p.B_dstate_stride % al == 0 &&
                 ok(p.C, p.C_batch_stride, p.C_group_stride) && p.C_dstate_stride % al == 0;

    // lanes per row: enough lanes to fill the machine, never more than the sequence can use
    int G = scan_lanes_per_row((int64_t)p.batch * p.dim, p.seqlen, p.dstate, "FM_SCAN_FWD_G");
    int NW = env_int("FM_SCAN_FWD_NW", 4);
#define FM_CASE(g, nw) if (G == g && NW == nw) return launch_cfg<T, g, nw>(p, st, vec_io, vec_bc);
    FM_CASE(1, 4) FM_CASE(2, 4) FM_CASE(4, 4) FM_CASE(8, 4) FM_CASE(16, 4) FM_CASE(32, 4)
    FM_CASE(8, 8) FM_CASE(16, 8) FM_CASE(8, 2) FM_CASE(16, 2)
#undef FM_CASE
    return launch_cfg<T, 8, 4>(p, st, vec_io, vec_bc);
}


}  // namespace fm
