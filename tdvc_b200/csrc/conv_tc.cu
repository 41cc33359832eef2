// placeholder until the tcgen05 kernels land
#include "common.cuh"
namespace tdvc {
int conv2d_tc_supported(const TdvcConvParams&) { return 0; }
int conv2d_tc(const TdvcConvParams&, cudaStream_t) { set_error("tcgen05 conv not built"); return TDVC_EINVAL; }
int dcn_tc_supported(const TdvcDcnParams&) { return 0; }
int dcn_tc(const TdvcDcnParams&, cudaStream_t) { set_error("tcgen05 dcn not built"); return TDVC_EINVAL; }
}
