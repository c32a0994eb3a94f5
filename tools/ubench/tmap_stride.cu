// Does cuTensorMapEncodeTiled accept a dimension whose stride is SMALLER than the previous one's?
// (d0 = 32 samples, d1 = channel [stride = pitch], d2 = 32-sample block [stride = 128 B])
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
int main()
{
    cudaFree(0);
    void *h = dlopen("libcuda.so.1", RTLD_NOW);
    typedef CUresult (*fn_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                             const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    fn_t enc = (fn_t)dlsym(h, "cuTensorMapEncodeTiled");
    float *d;
    cudaMalloc(&d, 64 * 4096);
    CUtensorMap m;
    cuuint64_t gdim[3] = { 32, 64, 32 };
    cuuint64_t gstr[2] = { 4096, 128 };
    cuuint32_t box[3] = { 32, 8, 2 }, es[3] = { 1, 1, 1 };
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("non-monotonic strides (pitch, 128): result %d\n", (int)r);
    cuuint64_t gdim2[3] = { 32, 32, 64 };
    cuuint64_t gstr2[2] = { 128, 4096 };
    r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, gdim2, gstr2, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("monotonic strides (128, pitch): result %d\n", (int)r);
    return 0;
}
