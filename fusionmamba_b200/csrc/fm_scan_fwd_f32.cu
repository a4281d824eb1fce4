// explicit instantiation of the forward scan for float I/O (one TU per dtype: parallel compilation)
#include "fm_scan_fwd.cuh"
namespace fm {
template cudaError_t launch_scan_fwd_T<float>(const FmScanFwdParams&, cudaStream_t);
}
