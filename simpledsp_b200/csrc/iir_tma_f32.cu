// iir_tma_f32.cu -- fp32 instantiations of the TMA-fed IIR kernels (split by precision to build in parallel)
#define SDSP_TMA_TYPE float
#define SDSP_TMA_SUFFIX f32
#include "iir_tma_impl.cuh"
