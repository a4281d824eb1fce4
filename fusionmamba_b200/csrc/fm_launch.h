// fm_launch.h -- host-side helpers shared by the launchers and the C-ABI translation unit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fm_scan.h"

namespace fm {

void count_launch();
int env_int(const char* name, int dflt);
// Lanes of a warp that cooperate on one channel row (power of two, 1..32). Enough lanes to fill 148 SMs,
// never more than ceil(seqlen / seg_len); overridable through the named environment variable (tuning only).
int scan_lanes_per_row(int64_t rows, int seqlen, int seg_len, const char* env_name);

// ---- time-split forward (fm_scan_fwd16.cuh, FwdSeg) --------------------------------------------------------------------------
constexpr int kWsRec = 36;   // workspace floats per (row, segment): 16 x (decay product | state) interleaved like x, delta sum, pad
// Time-split plan (see FwdSeg): segments per row for a shape whose single-pass grid leaves most of the GPU idle, else 1.
// The split instance is (SPL 2, NW 4, KT 2): 16 rows per CTA, 64-step chunks.
struct Fwd16Split { int n_seg; int seg_chunks; int64_t ws_bytes; };
inline Fwd16Split fwd16_split_plan(const FmScanFwdParams& p) {
    Fwd16Split r{1, 0, 0};
    const int dg = p.dim / p.n_groups;
    if (p.dstate != 16 || p.z != nullptr || dg % 16 != 0 || p.seqlen < 2048 || env_int("FM_SCAN_FWD16_SPLIT", 1) == 0) return r;   // (L = 1024 measured slower split: profiles/r02_split_p1024.jsonl)
    const int64_t warps = (int64_t)p.batch * p.n_groups * (dg / 16) * 4;    // single-pass grid of the (2, 4, 2) instance
    if (warps >= 2 * 592) return r;                                         // two warps per SM sub-partition already
    int J = env_int("FM_SCAN_FWD16_NSEG", 0);
    if (J <= 0) {
        J = (int)((6 * 592 + warps - 1) / warps);     // ~6 warps per SM sub-partition over both passes (sweep: profiles/r02_split_bench.jsonl)
        if (J > 16) J = 16;
        if (J > p.seqlen / 512) J = p.seqlen / 512;      // segments of at least 8 chunks
    }
    if (J < 2 || (int64_t)p.batch * J > 65535) return r;
    constexpr int TC = 64;
    const int n_chunks = (p.seqlen + TC - 1) / TC;
    r.seg_chunks = (n_chunks + J - 1) / J;     // (checkpoints sit on global chunk ends whatever the segment boundaries are)
    r.n_seg = (n_chunks + r.seg_chunks - 1) / r.seg_chunks;
    if (r.n_seg < 2) return Fwd16Split{1, 0, 0};
    r.ws_bytes = (int64_t)p.batch * p.dim * r.n_seg * kWsRec * (int64_t)sizeof(float);
    return r;
}


cudaError_t launch_scan_fwd(const FmScanFwdParams& p, cudaStream_t st);
cudaError_t launch_scan_bwd(const FmScanBwdParams& p, cudaStream_t st);
cudaError_t launch_unfold(const FmPermuteParams& p, cudaStream_t st);
cudaError_t launch_merge(const FmPermuteParams& p, cudaStream_t st);
cudaError_t launch_merge_norm(const FmNormParams& p, cudaStream_t st);
cudaError_t launch_layer_norm_bwd(const FmNormBwdParams& p, cudaStream_t st);
cudaError_t launch_block_gates(const FmBlockGatesParams& p, cudaStream_t st);
cudaError_t launch_block_scale(const FmBlockScaleParams& p, cudaStream_t st);
cudaError_t launch_block_combine(const FmBlockCombineParams& p, cudaStream_t st);
int block_gates_slabs(int batch, int positions);
int layer_norm_bwd_ctas(int dim, int64_t rows);
cudaError_t launch_conv_unfold(const FmConvUnfoldParams& p, cudaStream_t st);
cudaError_t launch_conv_unfold_bwd(const FmConvUnfoldBwdParams& p, cudaStream_t st);
cudaError_t launch_dt_proj(const FmDtProjParams& p, cudaStream_t st);

}  // namespace fm
