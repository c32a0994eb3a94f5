// iir_scan_f32.cu -- fp32 instantiations of the look-back scan kernels (split by precision to build in parallel)
#define SDSP_SCAN_TYPE float
#define SDSP_SCAN_SUFFIX f32
#include "iir_scan_impl.cuh"
