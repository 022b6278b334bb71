// conv_tc.cu -- placeholder until the tcgen05 kernel lands: nothing is routed here yet.
#include "conv_tc.h"
bool conv_tc_supported(int, int, int, int, int, int) { return false; }
void launch_conv_tc(int, const ConvArgs&, cudaStream_t) { xu_set_kernel_error("conv_tc: not built"); }
