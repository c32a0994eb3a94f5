// iir_scan_f64.cu -- fp64 instantiations of the look-back scan kernels (split by precision to build in parallel)
#define SDSP_SCAN_TYPE double
#define SDSP_SCAN_SUFFIX f64
#include "iir_scan_impl.cuh"
