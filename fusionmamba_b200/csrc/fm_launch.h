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

cudaError_t launch_scan_fwd(const FmScanFwdParams& p, cudaStream_t st);
cudaError_t launch_scan_bwd(const FmScanBwdParams& p, cudaStream_t st);
cudaError_t launch_unfold(const FmPermuteParams& p, cudaStream_t st);
cudaError_t launch_merge(const FmPermuteParams& p, cudaStream_t st);
cudaError_t launch_merge_norm(const FmNormParams& p, cudaStream_t st);
cudaError_t launch_conv_unfold(const FmConvUnfoldParams& p, cudaStream_t st);
cudaError_t launch_dt_proj(const FmDtProjParams& p, cudaStream_t st);

}  // namespace fm
