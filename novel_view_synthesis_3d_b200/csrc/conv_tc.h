// conv_tc.h -- tcgen05/TMEM/TMA implicit-GEMM convolution (bf16 in, fp32 accumulate in TMEM).
#pragma once
#include "kernels.h"
// true if (dtype, shape) is handled by the tcgen05 kernel; everything else goes to the SIMT kernel.
bool conv_tc_supported(int dtype, int Ci, int Co, int ks, int stride, int nseg);
void launch_conv_tc(int dtype, const ConvArgs& a, cudaStream_t s);
