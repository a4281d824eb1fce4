// explicit instantiation of the backward scan for float I/O (one TU per dtype: parallel compilation)
#include "fm_scan_bwd.cuh"
namespace fm {
template cudaError_t launch_scan_bwd_T<float>(const FmScanBwdParams&, cudaStream_t);
}
