// fm_scan_fwd.cuh -- forward launcher: alignment analysis + dispatch to the two forward kernels.
//   dstate == 16  -> fm_scan_fwd16.cuh (lane-serial, single pass; optional fused EfficientMerge store)
//   otherwise     -> fm_scan_fwd_rp.cuh (row-pair, time-parallel)
// Both replace selective_scan_fwd_kernel (selective_scan/selective_scan_fwd_kernel.cuh:67-303) and its host
// launcher selective_scan_fwd_launch (:305-345).
#pragma once
#include "fm_common.cuh"
#include "fm_launch.h"
#include "fm_scan_fwd16.cuh"
#include "fm_scan_fwd_rp.cuh"

namespace fm {

template <typename T>
cudaError_t launch_scan_fwd_T(const FmScanFwdParams& p, cudaStream_t st) {
    const int es = sizeof(T);
    const int64_t al = 16 / es;  // elements per 16 bytes
    auto ok = [&](const void* ptr, int64_t s0, int64_t s1) { return aligned16(ptr) && s0 % al == 0 && s1 % al == 0; };
    int vec_io = ok(p.u, p.u_batch_stride, p.u_d_stride) && ok(p.delta, p.delta_batch_stride, p.delta_d_stride) &&
                 (p.out_map != FM_MAP_LINEAR || ok(p.out, p.out_batch_stride, p.out_d_stride));   // fused merge stores are scalar
    if (p.z) vec_io = vec_io && ok(p.z, p.z_batch_stride, p.z_d_stride) && ok(p.out_z, p.out_z_batch_stride, p.out_z_d_stride);
    int vec_bc = ok(p.B, p.B_batch_stride, p.B_group_stride) && p.B_dstate_stride % al == 0 &&
                 ok(p.C, p.C_batch_stride, p.C_group_stride) && p.C_dstate_stride % al == 0;

    // dstate == 16 (the only state size FusionMamba uses): lane-serial single-pass kernel (fm_scan_fwd16.cuh)
    if (p.dstate == 16 && env_int("FM_SCAN_FWD16", 1) != 0) {
        const cudaError_t e16 = launch_scan_fwd16_T<T>(p, st, vec_io, vec_bc);
        if (e16 != cudaErrorInvalidConfiguration) return e16;   // no instance for this shape: use the generic kernel
    }
    if (p.out_map != FM_MAP_LINEAR) return cudaErrorInvalidConfiguration;   // only the dstate-16 kernel fuses the merge
    if (p.hck && p.hck_len == 8) return cudaErrorInvalidConfiguration;      // ... and writes the dense 8-step checkpoints
    // any other state size: row-pair kernel (fm_scan_fwd_rp.cuh)

    return launch_scan_fwd_rp_T<T>(p, st, vec_io, vec_bc);
}

}  // namespace fm
