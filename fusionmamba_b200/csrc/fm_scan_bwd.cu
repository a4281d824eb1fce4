// fm_scan_bwd.cu -- dtype dispatch for the backward scan (kernels: fm_scan_bwd.cuh).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "fm_launch.h"
namespace fm {
template <typename T> cudaError_t launch_scan_bwd_T(const FmScanBwdParams&, cudaStream_t);
extern template cudaError_t launch_scan_bwd_T<float>(const FmScanBwdParams&, cudaStream_t);
extern template cudaError_t launch_scan_bwd_T<__half>(const FmScanBwdParams&, cudaStream_t);
extern template cudaError_t launch_scan_bwd_T<__nv_bfloat16>(const FmScanBwdParams&, cudaStream_t);

cudaError_t launch_scan_bwd(const FmScanBwdParams& q, cudaStream_t st) {
    switch (q.f.dtype) {
        case FM_F32: return launch_scan_bwd_T<float>(q, st);
        case FM_F16: return launch_scan_bwd_T<__half>(q, st);
        default: return launch_scan_bwd_T<__nv_bfloat16>(q, st);
    }
}
}  // namespace fm
