// explicit instantiation of the backward scan for __nv_bfloat16 I/O (one TU per dtype: parallel compilation)
#include "fm_scan_bwd.cuh"
namespace fm {
template cudaError_t launch_scan_bwd_T<__nv_bfloat16>(const FmScanBwdParams&, cudaStream_t);
}
