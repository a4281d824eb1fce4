// fm_scan_fwd16.cuh -- selective-scan forward, dstate == 16 fast path for sm_100a ("lane-serial" kernel).
//
// Replaces selective_scan_fwd_kernel (selective_scan/selective_scan_fwd_kernel.cuh:67-303) for the state size
// FusionMamba uses (d_state = 16, models/cross.py:423).  Not a port -- the decomposition is different:
//   * a lane owns ONE channel row and SPL (2 or 4) of its 16 states and walks the sequence serially: the
//     recurrence h_t = a_t h_{t-1} + b_t is evaluated exactly once per (t, state) -- no block scan, no
//     up-sweep/down-sweep, no cross-lane prefix combine.  The carried state lives in registers for the whole row.
//     State pairs ride the Blackwell packed-fp32 pipe (FMUL2/FFMA2 with a scalar-broadcast operand), so one
//     (t, state) costs 2 packed issue slots + 1 MUFU.EX2.
//   * a CTA owns R = NW*32*SPL/16 rows of one (batch, group).  Per chunk of TC timesteps it stages, with
//     coalesced 128-bit loads,  the [16 x TC] B and C tiles (once, shared by all R rows -- the reference re-reads
//     them per row), and the [R x TC] u / delta tiles, evaluating softplus(delta + bias) and delta*u ONCE per
//     element on the way in.  The next chunk is prefetched into registers while the current one is scanned.
//   * B / C sit in shared memory time-major in 16-byte packets (TW = 4/SPL steps x SPL states) so that the
//     16/SPL lanes of a row read one contiguous line and the rows of a warp broadcast.
//   * y_t = sum_n C h is reduced over the 16/SPL lanes of a row through a small shared tile and leaves the SM
//     as 128-bit stores together with the D*u skip and the SiLU(z) gate.
// Masked (t >= L) steps are staged as delta = 0, B = C = 0, i.e. a = 1, b = 0: they leave h untouched.
#pragma once
#include "fm_common.cuh"
#include "fm_launch.h"

namespace fm {

template <int SPL> struct Fwd16Cfg {
    static constexpr int N = 16;
    static constexpr int LPR = N / SPL;                 // lanes per row
    static constexpr int RW = 32 / LPR;                 // rows per warp
    static constexpr int TW = 4 / SPL;                  // timesteps per 16-byte B/C packet
    static constexpr int PB = LPR * 4 + (SPL == 2 ? 8 : 4);   // floats per time block of packets (padded)
};

// load 4 consecutive elements (fp32) with tail / alignment fallback
template <typename T>
__device__ __forceinline__ float4 load4(const T* __restrict__ p, int nvalid, bool vec) {
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (nvalid <= 0) return r;
    if (vec && nvalid >= 4) {
        if constexpr (sizeof(T) == 4) {
            r = __ldg(reinterpret_cast<const float4*>(p));
        } else {
            const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
            const T* e = reinterpret_cast<const T*>(&v);
            r = make_float4(Cvt<T>::to_f(e[0]), Cvt<T>::to_f(e[1]), Cvt<T>::to_f(e[2]), Cvt<T>::to_f(e[3]));
        }
    } else {
        r.x = Cvt<T>::to_f(p[0]);
        if (nvalid > 1) r.y = Cvt<T>::to_f(p[1]);
        if (nvalid > 2) r.z = Cvt<T>::to_f(p[2]);
        if (nvalid > 3) r.w = Cvt<T>::to_f(p[3]);
    }
    return r;
}

template <typename T>
__device__ __forceinline__ void store4(T* __restrict__ p, int nvalid, bool vec, float4 v) {
    if (nvalid <= 0) return;
    if (vec && nvalid >= 4) {
        if constexpr (sizeof(T) == 4) {
            *reinterpret_cast<float4*>(p) = v;
        } else {
            uint2 o;
            T* e = reinterpret_cast<T*>(&o);
            e[0] = Cvt<T>::from_f(v.x); e[1] = Cvt<T>::from_f(v.y); e[2] = Cvt<T>::from_f(v.z); e[3] = Cvt<T>::from_f(v.w);
            *reinterpret_cast<uint2*>(p) = o;
        }
    } else {
        p[0] = Cvt<T>::from_f(v.x);
        if (nvalid > 1) p[1] = Cvt<T>::from_f(v.y);
        if (nvalid > 2) p[2] = Cvt<T>::from_f(v.z);
        if (nvalid > 3) p[3] = Cvt<T>::from_f(v.w);
    }
}

// B / C packets (4 elements) in shared memory keep the I/O dtype: 16-bit inputs are stored as they arrive (8-byte packets) and
// widened in registers after the load, which halves the shared-memory -> register fill of the scan loop (its bottleneck).
template <typename T>
__device__ __forceinline__ void sts_packet(T* p, float4 v) {
    if constexpr (sizeof(T) == 4) {
        *reinterpret_cast<float4*>(p) = v;
    } else {
        uint2 o;
        T* e = reinterpret_cast<T*>(&o);
        e[0] = Cvt<T>::from_f(v.x); e[1] = Cvt<T>::from_f(v.y); e[2] = Cvt<T>::from_f(v.z); e[3] = Cvt<T>::from_f(v.w);   // exact: v came from T
        *reinterpret_cast<uint2*>(p) = o;
    }
}
template <typename T>
__device__ __forceinline__ float4 lds_packet(const T* p) {
    if constexpr (sizeof(T) == 4) {
        return *reinterpret_cast<const float4*>(p);
    } else {
        const uint2 v = *reinterpret_cast<const uint2*>(p);
        const T* e = reinterpret_cast<const T*>(&v);
        return make_float4(Cvt<T>::to_f(e[0]), Cvt<T>::to_f(e[1]), Cvt<T>::to_f(e[2]), Cvt<T>::to_f(e[3]));
    }
}

// Raw 4-element vectors: the register prefetch holds the loaded bits untouched and widens them only when the chunk is staged --
// a conversion right after the load would make the warp wait for the global load before it can start scanning.
template <typename T> struct Raw4 { using type = uint2; };
template <> struct Raw4<float> { using type = float4; };
template <typename T>
__device__ __forceinline__ typename Raw4<T>::type ldg4_raw(const T* __restrict__ p) {
    return __ldg(reinterpret_cast<const typename Raw4<T>::type*>(p));
}
template <typename T>
__device__ __forceinline__ typename Raw4<T>::type pack4_raw(float4 v) {
    if constexpr (sizeof(T) == 4) {
        return v;
    } else {
        uint2 o;
        T* e = reinterpret_cast<T*>(&o);
        e[0] = Cvt<T>::from_f(v.x); e[1] = Cvt<T>::from_f(v.y); e[2] = Cvt<T>::from_f(v.z); e[3] = Cvt<T>::from_f(v.w);
        return o;
    }
}
template <typename T>
__device__ __forceinline__ float4 widen4(typename Raw4<T>::type v) {
    if constexpr (sizeof(T) == 4) {
        return v;
    } else {
        const T* e = reinterpret_cast<const T*>(&v);
        return make_float4(Cvt<T>::to_f(e[0]), Cvt<T>::to_f(e[1]), Cvt<T>::to_f(e[2]), Cvt<T>::to_f(e[3]));
    }
}

template <typename T>
__device__ __forceinline__ float4 ldg4_fast(const T* __restrict__ p) {   // 4 elements, 16-byte (fp32) / 8-byte aligned
    if constexpr (sizeof(T) == 4) {
        return __ldg(reinterpret_cast<const float4*>(p));
    } else {
        const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
        const T* e = reinterpret_cast<const T*>(&v);
        return make_float4(Cvt<T>::to_f(e[0]), Cvt<T>::to_f(e[1]), Cvt<T>::to_f(e[2]), Cvt<T>::to_f(e[3]));
    }
}

// ---- TMA bulk copies (cp.async.bulk, SASS: UBLKCP) + mbarrier completion for the B / C tiles -------------------------------------
// One 1-D bulk copy per (B | C, state) row of a chunk lands the raw TC-element row in a double-buffered shared staging area; the
// copy engine fetches the next chunk while the SM scans the current one, with no registers parked for it (the register prefetch
// it replaces held 16-32 registers per thread) and no LDG / wait in the instruction stream.  The transpose into the kernel's
// time-major packets happens at staging time from shared memory.
#ifndef FM_FWD16_TMA
#define FM_FWD16_TMA 1
#endif
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra.uni WAIT_DONE;\n"
        "bra.uni WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// (at least 3 resident CTAs of 128 threads: 384 CTAs at BASELINE configs[1] must fit 148 SMs in one wave)
// TS: storage type of the B / C packets in shared memory (T, or float when the launch is latency-bound and the widening
// instructions in the scan loop cost more than the halved register fill saves)
#ifndef FM_FWD16_UNROLL4
#define FM_FWD16_UNROLL4 1
#endif
// Time-split ("segmented") forward for shapes with too few rows to fill the GPU (one 1024x1024 pair: 768 rows, BASELINE
// configs[4]).  The sequence is cut into n_seg segments of seg_chunks chunks, each handled by its own CTA (grid.y = batch*n_seg):
//   kMode 1  aggregate pass: every segment runs the recurrence from h = 0 and leaves, per (row, state), the decay product of the
//            segment and its local end state (plus the row's sum of delta) in the workspace -- no C, no y, no output;
//   (carry)  scan_fwd16_carry_kernel folds the aggregates serially over the segments into the state ENTERING each segment;
//   kMode 2  final pass: the normal kernel, started from that state, restricted to the segment's chunks.
// 1.6x the arithmetic of the single pass for n_seg x the parallelism.  kMode 0 is the ordinary single pass.
struct FwdSeg {
    int n_seg;          // segments per row (1: no split)
    int seg_chunks;     // chunks (of this instance's TC timesteps) per segment
    float* ws;          // workspace: (batch*dim, n_seg, kWsRec) floats
    int tma_min_chunks; // B / C tiles arrive by TMA bulk copies when the sequence spans at least this many chunks (0: never)
    int lane_map;       // 0: lane = (row lane / LPR, state group lane % LPR); 1: lane = (row lane % RW, state group lane / RW) -- a
                        // quarter warp then shares its B / C packets (one 16-byte address per state group: two quarter warps per
                        // shared-memory wavefront instead of one) and reads RW different rows of delta / delta*u
};

template <typename T, typename TO, typename TS, int SPL, int NW, int KT, bool kHasZ, int kMode = 0>
__global__ void __launch_bounds__(NW * 32, (NW == 4 && KT == 2 ? 3 : 0))
scan_fwd16_kernel(const FmScanFwdParams p, const int vec_io, const int vec_bc, const FwdSeg sgm) {
    using Cf = Fwd16Cfg<SPL>;
    constexpr int N = 16, LPR = Cf::LPR, RW = Cf::RW, TW = Cf::TW, PB = Cf::PB;
    constexpr int NP = SPL / 2;                          // state pairs per lane
    constexpr int R = NW * RW;                           // rows per CTA
    constexpr int NT = NW * 32;
    constexpr int TC = 4 * LPR * KT;                     // timesteps per chunk: every thread stages KT float4 of u and delta
    constexpr int TQ = TC / 4;                           // float4 columns per row (= groups of 4 timesteps)
    constexpr int TCP = TC + 4;                          // padded row pitch of the u/delta/y tiles
    constexpr int NBLK = TC / TW;                        // packet time blocks per chunk
    constexpr int NBC = 2 * LPR * TQ;                    // B + C staging tasks per chunk (SPL states x 4 steps each)
    constexpr int KBC = (NBC + NT - 1) / NT;
    static_assert(TQ >= 4 && TQ % 2 == 0, "pipeline needs an even number (>= 4) of 4-step groups per chunk");
    constexpr bool kUnroll4 = FM_FWD16_UNROLL4;

    const int L = p.seqlen;
    const int dg = p.dim / p.n_groups;
    const int tiles_per_group = (dg + R - 1) / R;
    const int group = blockIdx.x / tiles_per_group;
    const int tile = blockIdx.x % tiles_per_group;
    const int b = kMode ? blockIdx.y / sgm.n_seg : blockIdx.y;
    const int seg = kMode ? blockIdx.y % sgm.n_seg : 0;
    if (kMode == 1 && seg == sgm.n_seg - 1) return;     // nobody consumes the last segment's aggregate
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    // 128-bit (64-bit for 16-bit TO) stores of `out` need its own alignment when TO != T
    const bool vec_out = sizeof(TO) == sizeof(T) ||
                         (((reinterpret_cast<uintptr_t>(p.out) & 15u) == 0) && p.out_batch_stride % 4 == 0 && p.out_d_stride % 4 == 0);

    extern __shared__ __align__(16) float smem[];
    TS* sB = reinterpret_cast<TS*>(smem);                // [NBLK][PB] packets of 4 elements, in the I/O dtype (or fp32)
    TS* sC = sB + NBLK * PB;
    float* sDl = smem + 2 * NBLK * PB;                   // [R][TCP]   softplus(delta + bias), 0 beyond L
    float* sDu = sDl + R * TCP;                          // [R][TCP]   delta * u
    float* sY = sDu + R * TCP;                           // [R*LPR][TCP] per-lane partial sums of C h
    float* sT = sY + R * LPR * TCP;                      // [TC][R + 1]  y transposed (channels-last fused-merge store only)
    // raw B / C rows landed by the copy engine: [B | C][16 states][TC] of T (one stage: it is refilled right after the barrier that
    // follows its consumption and lands while the chunk is scanned), then its mbarrier
    // row pitch TC + 16 bytes: the 8 lanes of a quarter warp (4 state groups x 2 column quads) then hit 8 different bank groups
    constexpr int RAWP = TC + 16 / (int)sizeof(T);
    T* sRaw = reinterpret_cast<T*>(sT + TC * (R + 1) + TC + ((4 - ((TC * (R + 1) + TC) & 3)) & 3));
    uint64_t* sBar = reinterpret_cast<uint64_t*>(reinterpret_cast<float*>(sRaw) + 2 * N * (TC + 4));   // (sized for fp32 rows)

    // ---- scan role: lane -> (row, state group) ------------------------------------------------------------------
    const int sg = sgm.lane_map ? lane / RW : lane % LPR;
    const int rc = warp * RW + (sgm.lane_map ? lane % RW : lane / LPR);   // row within CTA
    const int dloc_c = tile * R + rc;
    const bool rowc_ok = dloc_c < dg;
    const int dc = group * dg + (rowc_ok ? dloc_c : 0);  // invalid rows shadow row 0 of the group and never store
    float2 A2[NP], h2[NP];
#pragma unroll
    for (int q = 0; q < NP; ++q) {
        const float* Ap = reinterpret_cast<const float*>(p.A) + dc * p.A_d_stride;
        A2[q].x = Ap[(sg * SPL + 2 * q) * p.A_dstate_stride] * kLog2e;
        A2[q].y = Ap[(sg * SPL + 2 * q + 1) * p.A_dstate_stride] * kLog2e;
        h2[q] = make_float2(0.f, 0.f);
    }
    const int64_t rowid_c = static_cast<int64_t>(b) * p.dim + dc;
    float* __restrict__ xrow = reinterpret_cast<float*>(p.x) + rowid_c * p.n_chunks * 2 * N + sg * SPL * 2;
    float* __restrict__ hckrow = p.hck ? reinterpret_cast<float*>(p.hck) + rowid_c * p.n_hck * N + sg * SPL : nullptr;
    const int hck_mask = p.hck_len - 1, hck_shift = 31 - __clz(p.hck_len > 0 ? p.hck_len : 1);
    // x-slot boundaries: chunk_len is 2048 in every caller (a power of two) -> mask / shift instead of a division per chunk
    const int cl_mask = p.chunk_len - 1, cl_shift = 31 - __clz(p.chunk_len);
    const bool cl_pow2 = (p.chunk_len & cl_mask) == 0;
    float2 sumd2 = make_float2(0.f, 0.f);                // running sum of delta over the row (decay product stored in x)
    float* __restrict__ wsrow = kMode ? sgm.ws + (rowid_c * sgm.n_seg + seg) * kWsRec : nullptr;
    if constexpr (kMode == 2) {                          // state and delta sum entering this segment (written by the carry kernel)
#pragma unroll
        for (int q = 0; q < NP; ++q)
            h2[q] = make_float2(wsrow[2 * (sg * SPL + 2 * q) + 1], wsrow[2 * (sg * SPL + 2 * q + 1) + 1]);
        sumd2.x = wsrow[2 * N];
    }

    // ---- staging / output role: thread -> KT x (row, float4 column) ---------------------------------------------
    const T* uptr[KT];                                   // element 4*tq of the row; chunk c adds c*TC
    const T* dptr[KT];
    TO* optr[KT];                                        // TO = T, or float for a 16-bit forward with fp32 output
    const T* zptr[KT];
    T* ozptr[KT];
    float Dv[KT], bias[KT];
    int rs[KT], tq[KT];
    bool rows_ok[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) {
        const int e = tid + k * NT;
        rs[k] = e / TQ;
        tq[k] = e % TQ;
        const int dl_ = tile * R + rs[k];
        rows_ok[k] = dl_ < dg;
        const int d = group * dg + (rows_ok[k] ? dl_ : 0);
        uptr[k] = reinterpret_cast<const T*>(p.u) + b * p.u_batch_stride + d * p.u_d_stride + 4 * tq[k];
        dptr[k] = reinterpret_cast<const T*>(p.delta) + b * p.delta_batch_stride + d * p.delta_d_stride + 4 * tq[k];
        optr[k] = reinterpret_cast<TO*>(p.out) + b * p.out_batch_stride +
                  (p.out_map == FM_MAP_LINEAR ? d * p.out_d_stride + 4 * tq[k] : (d - group * dg) * p.out_d_stride);
        if constexpr (kHasZ) {
            zptr[k] = reinterpret_cast<const T*>(p.z) + b * p.z_batch_stride + d * p.z_d_stride + 4 * tq[k];
            ozptr[k] = reinterpret_cast<T*>(p.out_z) + b * p.out_z_batch_stride + d * p.out_z_d_stride + 4 * tq[k];
        }
        Dv[k] = p.D ? reinterpret_cast<const float*>(p.D)[d] : 0.f;
        bias[k] = p.delta_bias ? reinterpret_cast<const float*>(p.delta_bias)[d] : 0.f;
    }
    // B / C staging tasks: SPL states x 4 timesteps each; (tq_hi, sg, tq_lo) order: two lanes fill one 32-byte sector
    const T* bcptr[KBC];
    int64_t bcst[KBC];
    int bcdst[KBC], bct[KBC], bcraw[KBC];
#pragma unroll
    for (int k = 0; k < KBC; ++k) {
        const int task = (tid + k * NT) % NBC;
        const int which = task / (LPR * TQ);             // 0: B, 1: C
        const int rem = task % (LPR * TQ);
        const int tql = rem & 1, sgs = (rem >> 1) % LPR, tqh = rem / (2 * LPR);
        const int tqq = 2 * tqh + tql;
        bcst[k] = which ? p.C_dstate_stride : p.B_dstate_stride;
        bcptr[k] = (which ? reinterpret_cast<const T*>(p.C) + b * p.C_batch_stride + group * p.C_group_stride
                          : reinterpret_cast<const T*>(p.B) + b * p.B_batch_stride + group * p.B_group_stride) +
                   (sgs * SPL) * bcst[k] + 4 * tqq;
        bcdst[k] = (which ? NBLK * PB : 0) + sgs * 4 + tqq * SPL * PB;
        bct[k] = 4 * tqq;
        bcraw[k] = (which * N + sgs * SPL) * (TC + 16 / (int)sizeof(T)) + 4 * tqq;      // same elements inside the raw [B | C][16][TC + pad] stage
    }

    const int c_first = kMode ? seg * sgm.seg_chunks : 0;       // first chunk of this CTA
    // register prefetch buffers for one chunk
    typename Raw4<T>::type pu[KT], pd[KT], pbc[KBC][SPL];   // raw bits; widened at staging time
    // long rows only: a one- or two-chunk sequence has nothing to overlap the copy with, and pays the barrier set-up
    const bool use_tma = FM_FWD16_TMA && vec_bc && sgm.tma_min_chunks > 0 && (L + TC - 1) / TC >= sgm.tma_min_chunks;
    auto tma_chunk = [&](int c) { return use_tma && (c + 1) * TC <= L; };
    auto prefetch = [&](int c) {
        const int t0 = c * TC;
        if (tma_chunk(c) && lane < 32 / NW) {
            // the 32 (B | C, state) rows of the chunk are shared out over the warps (32 / NW copies each, one per lane): TC contiguous
            // elements, 16-byte aligned (vec_bc).  Every issuing warp announces its own share of the bytes.
            const int row = warp * (32 / NW) + lane;
            const int which = row >> 4, n = row & 15;
            if (lane == 0) mbar_expect_tx(sBar, (32u / NW) * TC * sizeof(T));
            __syncwarp((32 / NW) == 32 ? 0xffffffffu : ((1u << (32 / NW)) - 1u));
            const T* src = (which ? reinterpret_cast<const T*>(p.C) + b * p.C_batch_stride + group * p.C_group_stride + n * p.C_dstate_stride
                                  : reinterpret_cast<const T*>(p.B) + b * p.B_batch_stride + group * p.B_group_stride + n * p.B_dstate_stride) + t0;
            tma_bulk_g2s(sRaw + (which * N + n) * RAWP, src, TC * sizeof(T), sBar);
        }
        if (vec_io && vec_bc && t0 + TC <= L) {          // CTA-uniform fast path: whole chunk in range, 128-bit loads
#pragma unroll
            for (int k = 0; k < KT; ++k) {
                pu[k] = ldg4_raw<T>(uptr[k] + t0);
                pd[k] = ldg4_raw<T>(dptr[k] + t0);
            }
            if (!tma_chunk(c)) {
#pragma unroll
                for (int k = 0; k < KBC; ++k)
                    if (NBC % NT == 0 || tid + k * NT < NBC) {
#pragma unroll
                        for (int j = 0; j < SPL; ++j) pbc[k][j] = ldg4_raw<T>(bcptr[k] + j * bcst[k] + t0);
                    }
            }
        } else {
#pragma unroll
            for (int k = 0; k < KT; ++k) {
                const int t = t0 + 4 * tq[k];
                pu[k] = pack4_raw<T>(load4<T>(uptr[k] + t0, L - t, vec_io));
                pd[k] = pack4_raw<T>(load4<T>(dptr[k] + t0, L - t, vec_io));
            }
#pragma unroll
            for (int k = 0; k < KBC; ++k)
                if (NBC % NT == 0 || tid + k * NT < NBC) {
                    const int t = t0 + bct[k];
#pragma unroll
                    for (int j = 0; j < SPL; ++j) pbc[k][j] = pack4_raw<T>(load4<T>(bcptr[k] + j * bcst[k] + t0, L - t, vec_bc));
                }
        }
    };

    // ---- scan-phase pipeline stages (see the loop below) -----------------------------------------------------------
    struct Raw { float4 d4, u4, bp[SPL], cp[SPL]; };                 // shared-memory reads of one 4-step group
    struct Cmp { float2 a[4][NP], b[4][NP], c[4][NP]; };             // decay, input term, C of one 4-step group
    const float* pDl = sDl + rc * TCP;
    const float* pDu = sDu + rc * TCP;
    const TS* pB = sB + sg * 4;
    const TS* pC = sC + sg * 4;
    // partial sums of a row: LPR lines, row-interleaved under lane_map 1 (the lanes of a quarter warp then store to distinct banks)
    float* pY = sY + (sgm.lane_map ? sg * R + rc : rc * LPR + sg) * TCP;
    auto load_raw = [&](int t4, Raw& r) {
        r.d4 = lds128(pDl + 4 * t4);
        r.u4 = lds128(pDu + 4 * t4);
#pragma unroll
        for (int blk = 0; blk < SPL; ++blk) {
            r.bp[blk] = lds_packet<TS>(pB + (t4 * SPL + blk) * PB);
            if constexpr (kMode != 1) r.cp[blk] = lds_packet<TS>(pC + (t4 * SPL + blk) * PB);
        }
    };
    auto compute = [&](const Raw& r, Cmp& g) {
        const float dls[4] = {r.d4.x, r.d4.y, r.d4.z, r.d4.w};
        const float dus[4] = {r.u4.x, r.u4.y, r.u4.z, r.u4.w};
        sumd2 = add2(sumd2, add2(make_float2(r.d4.x, r.d4.y), make_float2(r.d4.z, r.d4.w)));
#pragma unroll
        for (int blk = 0; blk < SPL; ++blk) {
            const float bv[4] = {r.bp[blk].x, r.bp[blk].y, r.bp[blk].z, r.bp[blk].w};
            float cv[4] = {0.f, 0.f, 0.f, 0.f};
            if constexpr (kMode != 1) { cv[0] = r.cp[blk].x; cv[1] = r.cp[blk].y; cv[2] = r.cp[blk].z; cv[3] = r.cp[blk].w; }
#pragma unroll
            for (int tt = 0; tt < TW; ++tt) {
                const int i = blk * TW + tt;
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    const float2 x2 = mul2(bcast2(dls[i]), A2[q]);
                    g.a[i][q] = make_float2(ex2_approx(x2.x), ex2_approx(x2.y));
                    g.b[i][q] = mul2(bcast2(dus[i]), make_float2(bv[tt * SPL + 2 * q], bv[tt * SPL + 2 * q + 1]));
                    if constexpr (kMode != 1) g.c[i][q] = make_float2(cv[tt * SPL + 2 * q], cv[tt * SPL + 2 * q + 1]);
                }
            }
        }
    };
    auto chain = [&](const Cmp& g, int t4) {
        float ys[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 acc;
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                h2[q] = fma2(g.a[i][q], h2[q], g.b[i][q]);
                if constexpr (kMode != 1) acc = (q == 0) ? mul2(g.c[i][q], h2[q]) : fma2(g.c[i][q], h2[q], acc);
            }
            if constexpr (kMode != 1) ys[i] = acc.x + acc.y;
        }
        if constexpr (kMode != 1) sts128(pY + 4 * t4, make_float4(ys[0], ys[1], ys[2], ys[3]));
    };

    // dense 8-step checkpoints (hck_len == 8, read by fm_scan_bwd_ls.cuh): the state after every 8th timestep, stored from
    // inside the scan loop (a 4-step group pair ends on a multiple of 8); coarser spacings are stored at chunk ends below
    const bool hck8 = kMode != 1 && hckrow != nullptr && p.hck_len == 8;
    auto store_hck8 = [&](int te) {
        if (te < L && rowc_ok) {
            float* dst = hckrow + static_cast<int64_t>((te >> 3) - 1) * N;
#pragma unroll
            for (int q = 0; q < NP; ++q) *reinterpret_cast<float2*>(dst + 2 * q) = h2[q];
        }
    };
    const int n_chunks_all = (L + TC - 1) / TC;
    const int c_begin = c_first;
    const int n_chunks = kMode ? min(n_chunks_all, c_begin + sgm.seg_chunks) : n_chunks_all;
    if (use_tma) {
        if (tid == 0) {
            mbar_init(sBar, NW);                       // one arrival (with its byte count) per issuing warp and chunk
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
    prefetch(c_begin);

    for (int c = c_begin; c < n_chunks; ++c) {
        const int t0 = c * TC;
        // ---- stage chunk c from the prefetch registers ----------------------------------------------------------
        float4 du4[KT];                                   // D * u, kept for the output phase
#pragma unroll
        for (int k = 0; k < KT; ++k) {
            const int t = t0 + 4 * tq[k];
            const float4 pdk = widen4<T>(pd[k]), puk = widen4<T>(pu[k]);
            float dl[4] = {pdk.x, pdk.y, pdk.z, pdk.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float xv = dl[i] + bias[k];
                const float sp = p.delta_softplus ? softplus_fast(xv) : xv;
                dl[i] = (t + i < L) ? sp : 0.f;
            }
            sts128(sDl + rs[k] * TCP + 4 * tq[k], make_float4(dl[0], dl[1], dl[2], dl[3]));
            sts128(sDu + rs[k] * TCP + 4 * tq[k], make_float4(dl[0] * puk.x, dl[1] * puk.y, dl[2] * puk.z, dl[3] * puk.w));
            du4[k] = make_float4(Dv[k] * puk.x, Dv[k] * puk.y, Dv[k] * puk.z, Dv[k] * puk.w);
        }
        const bool by_tma = tma_chunk(c);
        if (by_tma) mbar_wait(sBar, (c - c_first) & 1);
#pragma unroll
        for (int k = 0; k < KBC; ++k) {
            if (NBC % NT == 0 || tid + k * NT < NBC) {
                float g[SPL][4];
                if (by_tma) {
                    const T* rawt = sRaw + bcraw[k];
#pragma unroll
                    for (int j = 0; j < SPL; ++j) pbc[k][j] = *reinterpret_cast<const typename Raw4<T>::type*>(rawt + j * RAWP);
                }
#pragma unroll
                for (int j = 0; j < SPL; ++j) { const float4 w4 = widen4<T>(pbc[k][j]); g[j][0] = w4.x; g[j][1] = w4.y; g[j][2] = w4.z; g[j][3] = w4.w; }
#pragma unroll
                for (int i = 0; i < SPL; ++i) {          // SPL = 4/TW time blocks per task
                    float e[4];
#pragma unroll
                    for (int tt = 0; tt < TW; ++tt)
#pragma unroll
                        for (int j = 0; j < SPL; ++j) e[tt * SPL + j] = g[j][i * TW + tt];
                    sts_packet<TS>(sB + bcdst[k] + i * PB, make_float4(e[0], e[1], e[2], e[3]));
                }
            }
        }
        __syncthreads();
        if (c + 1 < n_chunks) prefetch(c + 1);

        // ---- scan chunk c: serial recurrence, SPL states per lane -----------------------------------------------
        // Three-stage software pipeline over groups of 4 timesteps, carried through the registers of a ROLLED loop:
        //   iteration i:  shared-memory reads of group i+2 | delta*A, ex2, delta*u*B of group i+1 | h-chain of group i
        // Every consumer's operands were produced one iteration (~50 instructions) earlier, so the LDS and MUFU
        // latencies are covered inside one warp; the loop body holds two iterations so the buffers alternate by name.
        {
            Raw r0, r1;
            Cmp c0, c1;
            load_raw(0, r1);
            compute(r1, c0);
            load_raw(1, r1);
            // invariant at the top of an (even) iteration i: r1 = raw of group i+1, c0 = computed group i
            auto body = [&](int i) {
                load_raw(i + 2, r0);
                compute(r1, c1);
                chain(c0, i);
                load_raw(i + 3, r1);
                compute(r0, c0);
                chain(c1, i + 1);
                if (hck8) store_hck8(t0 + 4 * (i + 2));
            };
            // two bodies per trip: the running shared-memory address registers are bumped half as often (each bump waits
            // for the queued LDS/STS that still read them)
            int i = 0;
            if (kUnroll4) {
#pragma unroll 1
                for (; i + 4 <= TQ - 2; i += 4) { body(i); body(i + 2); }
            }
#pragma unroll 1
            for (; i < TQ - 2; i += 2) body(i);
            compute(r1, c1);
            chain(c0, TQ - 2);
            chain(c1, TQ - 1);
            if (hck8) store_hck8(t0 + TC);

            // dense checkpoint (state after timestep te-1, te % hck_len == 0, interior boundaries only); the launcher
            // guarantees hck_len % TC == 0, so boundaries fall on chunk ends
            if (kMode != 1 && hckrow != nullptr && !hck8) {
                const int te = t0 + TC;
                if ((te & hck_mask) == 0 && te < L && rowc_ok) {
                    float* dst = hckrow + ((te >> hck_shift) - 1) * N;
#pragma unroll
                    for (int q = 0; q < NP; ++q) *reinterpret_cast<float2*>(dst + 2 * q) = h2[q];
                }
            }
            // x checkpoint: (running decay product, state) at slot ends / at L   (selective_scan_fwd_kernel.cuh:253)
            const int t_end = min(t0 + TC, L);
            const bool slot_end = cl_pow2 ? ((t_end & cl_mask) == 0) : (t_end % p.chunk_len == 0);
            if (kMode != 1 && rowc_ok && (slot_end || t_end == L)) {
                const float sumd = sumd2.x + sumd2.y;
                float* xs = xrow + (cl_pow2 ? ((t_end - 1) >> cl_shift) : ((t_end - 1) / p.chunk_len)) * 2 * N;
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    *reinterpret_cast<float4*>(xs + 4 * q) =
                        make_float4(ex2_approx(A2[q].x * sumd), h2[q].x, ex2_approx(A2[q].y * sumd), h2[q].y);
                }
            }
        }
        __syncthreads();

        // ---- output chunk c: reduce the LPR partial sums, add D*u, gate, store ----------------------------------
        if constexpr (kMode != 1) {
#pragma unroll
        for (int k = 0; k < KT; ++k) {
            const int t = t0 + 4 * tq[k];
            float2 ya = make_float2(du4[k].x, du4[k].y), yb = make_float2(du4[k].z, du4[k].w);
            const float* src = sY + (sgm.lane_map ? rs[k] : rs[k] * LPR) * TCP + 4 * tq[k];
            const int jst = sgm.lane_map ? R * TCP : TCP;
#pragma unroll
            for (int j = 0; j < LPR; ++j) {
                const float4 v = lds128(src + j * jst);
                ya = add2(ya, make_float2(v.x, v.y));
                yb = add2(yb, make_float2(v.z, v.w));
            }
            float4 y = make_float4(ya.x, ya.y, yb.x, yb.y);
            if (rows_ok[k]) {
                if (p.out_map == FM_MAP_LINEAR) {
                    store4<TO>(optr[k] + t0, L - t, vec_io && vec_out, y);
                } else if (p.out_map == FM_MAP_EFFICIENT_V2_CL) {
                    float* dstT = sT + (4 * tq[k]) * (R + 1) + rs[k];         // transposed: the store below runs with lanes along channels
                    dstT[0] = y.x; dstT[R + 1] = y.y; dstT[2 * (R + 1)] = y.z; dstT[3 * (R + 1)] = y.w;
                } else {
                    // fused EfficientMerge (models/cross.py:34-58): direction k = group, element l -> pixel of sub-grid k;
                    // every pixel of y (B, D, H*W) is written exactly once, padded positions of odd sizes are dropped
                    const int H = p.map_h, W = p.map_w, Hp = (H + 1) >> 1, Wp = (W + 1) >> 1;
                    const float yv[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int l = t + i;
                        if (l < L) {
                            int ii, jj;
                            if (group & 1) { jj = l / Hp; ii = l - jj * Hp; } else { ii = l / Wp; jj = l - ii * Wp; }
                            const int h = 2 * ii + (group & 1), w = 2 * jj + (group >> 1);
                            if (h < H && w < W) optr[k][h * W + w] = Cvt<TO>::from_f(yv[i]);
                        }
                    }
                }
                if constexpr (kHasZ) {
                    const float4 z = load4<T>(zptr[k] + t0, L - t, vec_io);
                    y.x *= z.x * sigmoid_f(z.x); y.y *= z.y * sigmoid_f(z.y);
                    y.z *= z.z * sigmoid_f(z.z); y.w *= z.w * sigmoid_f(z.w);
                    store4<T>(ozptr[k] + t0, L - t, vec_io, y);
                }
            }
        }
        if (p.out_map == FM_MAP_EFFICIENT_V2_CL) {
            // fused EfficientMerge, channels-last: out[b, pixel(group, l), channel] -- the R rows of the CTA at one l are R
            // consecutive channels, i.e. one contiguous run per pixel
            int* sPix = reinterpret_cast<int*>(sT + TC * (R + 1));          // [TC] pixel of each timestep of the chunk, or -1
            if (tid < TC) {                                                  // one index computation per timestep, not per element
                const int H = p.map_h, W = p.map_w, Hp = (H + 1) >> 1, Wp = (W + 1) >> 1;
                const int l = t0 + tid;
                int pix = -1;
                if (l < L) {
                    int ii, jj;
                    if (group & 1) { jj = l / Hp; ii = l - jj * Hp; } else { ii = l / Wp; jj = l - ii * Wp; }
                    const int h = 2 * ii + (group & 1), w = 2 * jj + (group >> 1);
                    if (h < H && w < W) pix = h * W + w;
                }
                sPix[tid] = pix;
            }
            __syncthreads();
            TO* ocl = reinterpret_cast<TO*>(p.out) + b * p.out_batch_stride + tile * R;
            constexpr int RSH = (R == 4 ? 2 : R == 8 ? 3 : R == 16 ? 4 : R == 32 ? 5 : R == 64 ? 6 : 0);
            static_assert((1 << RSH) == R, "rows per CTA must be a power of two");
            const int rmax = dg - tile * R;                                  // valid rows of this tile
            // one thread stores 4 consecutive channels of one pixel (16 / 8 bytes) when the pixel rows are aligned for it
            const bool vec_cl = (p.out_d_stride % 4 == 0) && (p.out_batch_stride % 4 == 0) &&
                                ((reinterpret_cast<uintptr_t>(p.out) & (4 * sizeof(TO) - 1)) == 0);
            constexpr int R4 = R / 4;
            for (int e = tid; e < TC * R4; e += NT) {
                const int ll = e / R4, r4 = (e % R4) * 4;
                const int pix = sPix[ll];
                if (pix < 0 || r4 >= rmax) continue;
                const float* sp = sT + ll * (R + 1) + r4;
                store4<TO>(ocl + static_cast<int64_t>(pix) * p.out_d_stride + r4, rmax - r4, vec_cl, make_float4(sp[0], sp[1], sp[2], sp[3]));
            }
        }
        }   // kMode != 1
        // no barrier: the next staging writes sDl/sDu/sB/sC (scan reads finished at the barrier above); sY / sT are next
        // written after the post-staging barrier of the next chunk.
    }
    if constexpr (kMode == 1) {
        // segment aggregate: decay product over the segment and the local end state, interleaved like x; the row's delta sum once
        if (rowc_ok) {
            const float sumd = sumd2.x + sumd2.y;
#pragma unroll
            for (int q = 0; q < NP; ++q)
                *reinterpret_cast<float4*>(wsrow + 2 * (sg * SPL + 2 * q)) =
                    make_float4(ex2_approx(A2[q].x * sumd), h2[q].x, ex2_approx(A2[q].y * sumd), h2[q].y);
            if (sg == 0) wsrow[2 * N] = sumd;
        }
    }
}

// Carry between the two passes of the time-split forward: one thread per (row, state) walks the segments in order and replaces
// (decay product, local end state) by (unused, state entering the segment); the thread of state 0 also turns the per-segment
// delta sums into the sum entering each segment.  rows x 16 threads, n_seg serial steps each.
static __global__ void scan_fwd16_carry_kernel(float* __restrict__ ws, const int64_t rows, const int n_seg) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= rows * 16) return;
    const int64_t row = i >> 4;
    const int n = static_cast<int>(i & 15);
    float* rec = ws + row * n_seg * kWsRec;
    float h = 0.f, sd = 0.f;
    for (int j = 0; j < n_seg; ++j, rec += kWsRec) {
        const float P = rec[2 * n], hl = rec[2 * n + 1];
        rec[2 * n + 1] = h;
        h = fmaf(P, h, hl);
        if (n == 0) {
            const float d = rec[32];
            rec[32] = sd;
            sd += d;
        }
    }
}

template <int SPL, int NW, int KT>
constexpr size_t fwd16_smem_bytes() {
    using Cf = Fwd16Cfg<SPL>;
    constexpr int TC = 4 * Cf::LPR * KT, R = NW * Cf::RW;
    return sizeof(float) * (2 * (size_t)(TC / Cf::TW) * Cf::PB + 2 * (size_t)R * (TC + 4) + (size_t)R * Cf::LPR * (TC + 4) +
                            (size_t)TC * (R + 1) + (size_t)TC + 4 /* alignment */ + 2 * 16 * (size_t)(TC + 4) /* raw B/C stage */ + 4 /* mbarrier */);
}

template <typename T, int SPL, int NW, int KT, int kMode = 0>
static cudaError_t launch_fwd16_cfg(const FmScanFwdParams& p, cudaStream_t st, int vec_io, int vec_bc, FwdSeg sgm = FwdSeg{1, 0, nullptr, 0, 0}) {
    // B / C tiles by TMA bulk copies: measured on B200 (profiles/r02_tma_ab.jsonl) -7 % on fp32 rows of 16+ chunks (256-byte row
    // copies overlap the scan), +2 ... +10 % on 16-bit rows (128-byte copies: the issue cost per byte doubles) and on sequences
    // of a few chunks (nothing to overlap with) -- so it is on for the former only, and only then is the staging area allocated
    // (the extra shared memory would otherwise just shrink the L1 of the non-TMA launches).
    constexpr int TCc = 4 * Fwd16Cfg<SPL>::LPR * KT;
    const int min_chunks = env_int("FM_SCAN_FWD16_TMA_MINCHUNKS", sizeof(T) == 4 ? 16 : 0);
    const bool tma_on = FM_FWD16_TMA && vec_bc && min_chunks > 0 && (p.seqlen + TCc - 1) / TCc >= min_chunks;
    sgm.tma_min_chunks = tma_on ? min_chunks : 0;
    // quarter warps that share their B / C packets: measured -1 ... -7 % on fp32 rows, +2 % on 16-bit packets (profiles/r02_fwd_lm_ab.jsonl)
    sgm.lane_map = env_int("FM_SCAN_FWD16_LM", sizeof(T) == 4 ? 1 : 0) != 0;
    constexpr int R = NW * Fwd16Cfg<SPL>::RW;
    const int dg = p.dim / p.n_groups;
    const int tiles = (dg + R - 1) / R;
    dim3 grid(tiles * p.n_groups, p.batch * (kMode ? sgm.n_seg : 1));
    const size_t smem = fwd16_smem_bytes<SPL, NW, KT>() - (tma_on ? 0 : sizeof(float) * (2 * 16 * (size_t)(TCc + 4) + 4));
    using KernT = void (*)(const FmScanFwdParams, int, int, const FwdSeg);
    KernT kern;
    if constexpr (kMode != 0) {
        // time-split passes: no z; fp32 output from 16-bit inputs is the only mixed form (out_dtype)
        kern = scan_fwd16_kernel<T, T, T, SPL, NW, KT, false, kMode>;
        if constexpr (sizeof(T) == 2) {
            if (p.out_dtype == FM_F32) kern = scan_fwd16_kernel<T, float, T, SPL, NW, KT, false, kMode>;
        }
    } else {
        kern = p.z ? scan_fwd16_kernel<T, T, T, SPL, NW, KT, true> : scan_fwd16_kernel<T, T, T, SPL, NW, KT, false>;
        if constexpr (sizeof(T) == 2) {
            // fewer than ~2 warps per SM sub-partition: the kernel is bound by one warp's instruction latency, not by shared-memory fill
            const bool few_warps = (int64_t)grid.x * grid.y * NW < 2 * 592;
            if (p.out_dtype == FM_F32)                     // z == NULL checked by the C ABI
                kern = few_warps ? scan_fwd16_kernel<T, float, float, SPL, NW, KT, false> : scan_fwd16_kernel<T, float, T, SPL, NW, KT, false>;
            else if (few_warps)
                kern = p.z ? scan_fwd16_kernel<T, T, float, SPL, NW, KT, true> : scan_fwd16_kernel<T, T, float, SPL, NW, KT, false>;
        }
    }
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, NW * 32, smem, st>>>(p, vec_io, vec_bc, sgm);
    count_launch();
    return cudaGetLastError();
}

// Launch-shape heuristic (tuned on B200, profiles/r01_fwd16_tune.jsonl): SPL = 2 (8 lanes per row) until the grid has
// enough rows that SPL = 4 still fills 148 SMs; long rows use 64-step chunks; the CTA row count divides the
// channels of a group where possible (no shadow rows).
template <typename T>
cudaError_t launch_scan_fwd16_T(const FmScanFwdParams& p, cudaStream_t st, int vec_io, int vec_bc) {
    const int64_t rows = (int64_t)p.batch * p.dim;
    const int dg = p.dim / p.n_groups;
    // few rows, long sequence (one 1024x1024 pair): aggregate pass + carry + final pass over n_seg segments, if the caller
    // provided the workspace (fm_scan_fwd_workspace_bytes)
    {
        const Fwd16Split sp = fwd16_split_plan(p);
        if (sp.n_seg > 1 && p.workspace != nullptr && p.workspace_bytes >= sp.ws_bytes &&
            (!p.hck || p.hck_len == 8 || p.hck_len % 64 == 0)) {
            FwdSeg sgm{sp.n_seg, sp.seg_chunks, reinterpret_cast<float*>(p.workspace), 0, 0};
            cudaError_t e = launch_fwd16_cfg<T, 2, 4, 2, 1>(p, st, vec_io, vec_bc, sgm);
            if (e != cudaSuccess) return e;
            const int64_t n = rows * 16;
            scan_fwd16_carry_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(sgm.ws, rows, sp.n_seg);
            count_launch();
            e = cudaGetLastError();
            if (e != cudaSuccess) return e;
            return launch_fwd16_cfg<T, 2, 4, 2, 2>(p, st, vec_io, vec_bc, sgm);
        }
    }
    int SPL = env_int("FM_SCAN_FWD16_SPL", 0);
    if (SPL != 2 && SPL != 4) SPL = (rows >= 24576) ? 4 : 2;      // (4 states per lane once its half as many warps still give 5+ per SM sub-partition)
    int NW = env_int("FM_SCAN_FWD16_NW", 0);
    if (NW != 1 && NW != 2 && NW != 4 && NW != 8) {
        NW = 4;
        const int rw = SPL == 2 ? 4 : 8;
        while (NW > 1 && dg % (NW * rw) != 0) NW >>= 1;
        // few rows (e.g. one 1024x1024 pair, BASELINE configs[4]): smaller CTAs so that more SMs get one
        while (NW > 2 && (int64_t)p.batch * p.n_groups * ((dg + NW * rw - 1) / (NW * rw)) < 148) NW >>= 1;
    }
    int KT = env_int("FM_SCAN_FWD16_KT", 0);
    // two staged items per thread once the sequence spans several chunks (profiles/r01_fwd16_tune.jsonl, r01_stage_shapes_tune.log)
    if (KT != 1 && KT != 2) KT = ((SPL == 2 && p.seqlen >= 1024) || (SPL == 4 && p.seqlen >= 256)) ? 2 : 1;
    // dense checkpoints must fall on chunk ends (TC = 4 * (16 / SPL) * KT timesteps)
    // (hck_len == 8 is stored from inside the scan loop and has no such constraint)
    if (p.hck && p.hck_len != 8 && p.hck_len % (4 * (16 / SPL) * KT) != 0) KT = 1;
    if (p.hck && p.hck_len != 8 && p.hck_len % (4 * (16 / SPL) * KT) != 0) return cudaErrorInvalidConfiguration;
#define FM_CASE16(spl, nw, kt) \
    if (SPL == spl && NW == nw && KT == kt) return launch_fwd16_cfg<T, spl, nw, kt>(p, st, vec_io, vec_bc);
    FM_CASE16(2, 1, 1) FM_CASE16(2, 2, 1) FM_CASE16(2, 4, 1) FM_CASE16(2, 8, 1)
    FM_CASE16(2, 1, 2) FM_CASE16(2, 2, 2) FM_CASE16(2, 4, 2)
    FM_CASE16(4, 1, 2) FM_CASE16(4, 2, 2) FM_CASE16(4, 4, 2)
    FM_CASE16(4, 1, 1) FM_CASE16(4, 2, 1) FM_CASE16(4, 4, 1)
#undef FM_CASE16
    return cudaErrorInvalidConfiguration;
}

}  // namespace fm
