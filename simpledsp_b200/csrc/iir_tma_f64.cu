// iir_tma_f64.cu -- fp64 instantiations of the TMA-fed IIR kernels (split by precision to build in parallel)
#define SDSP_TMA_TYPE double
#define SDSP_TMA_SUFFIX f64
#include "iir_tma_impl.cuh"
