// fm_scan_fwd.cu -- dtype dispatch for the forward scan (kernels: fm_scan_fwd.cuh).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "fm_launch.h"
namespace fm {
template <typename T> cudaError_t launch_scan_fwd_T(const FmScanFwdParams&, cudaStream_t);
extern template cudaError_t launch_scan_fwd_T<float>(const FmScanFwdParams&, cudaStream_t);
extern template cudaError_t launch_scan_fwd_T<__half>(const FmScanFwdParams&, cudaStream_t);
extern template cudaError_t launch_scan_fwd_T<__nv_bfloat16>(const FmScanFwdParams&, cudaStream_t);

cudaError_t launch_scan_fwd(const FmScanFwdParams& p, cudaStream_t st) {
    switch (p.dtype) {
        case FM_F32: return launch_scan_fwd_T<float>(p, st);
        case FM_F16: return launch_scan_fwd_T<__half>(p, st);
        default: return launch_scan_fwd_T<__nv_bfloat16>(p, st);
    }
}
}  // namespace fm
