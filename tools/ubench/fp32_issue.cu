// Micro-benchmark (B200): how fast can ONE warp on an SMSP issue dependent / independent fp32 operations?
// The lane-per-channel IIR kernel of a 16384-channel bank runs 512 warps on 592 SMSPs, so its speed is the
// single-warp issue rate of FFMA / FADD / FFMA2 chains.  K independent chains of dependent ops per thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_issue fp32_issue.cu && ./fp32_issue
#include <cstdio>
#include <cuda_runtime.h>

template <int K, int MODE>
__global__ void __launch_bounds__(128) chains(float *out, const float *in, int iters, long long *cycles)
{
    float a[K], b[K], c[K];
    float2 a2[K], b2[K], c2[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
        a[k] = in[threadIdx.x + 32 * k];
        b[k] = in[threadIdx.x + 32 * k + 1024];
        c[k] = in[threadIdx.x + 32 * k + 2048];
        a2[k] = make_float2(a[k], b[k]);
        b2[k] = make_float2(b[k], c[k]);
        c2[k] = make_float2(c[k], a[k]);
    }
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int k = 0; k < K; k++) {
                if (MODE == 0)
                    a[k] = __fmaf_rn(a[k], b[k], c[k]);
                else if (MODE == 1)
                    a2[k] = __ffma2_rn(a2[k], b2[k], c2[k]);
                else if (MODE == 2)
                    a[k] = __fadd_rn(a[k], b[k]);
                else if (MODE == 4) // all chains share the multiplier register (operand-reuse cache candidates)
                    a[k] = __fmaf_rn(b[0], a[k], c[k]);
                else if (MODE == 5) // share multiplier and addend
                    a[k] = __fmaf_rn(b[0], a[k], c[0]);
                else if (MODE == 6) // pairs share the multiplier
                    a[k] = __fmaf_rn(b[k / 2], a[k], c[k]);
                else if (MODE == 3) { // the delta-form section: 3 fma + fma + add, as a dependent chain
                    float t = __fmaf_rn(b[k], c[k], a[k]);
                    t = __fmaf_rn(c[k], a[k], t);
                    t = __fmaf_rn(b[k], t, c[k]);
                    b[k] = __fmaf_rn(a[k], b[k], t);
                    a[k] = __fadd_rn(a[k], b[k]);
                }
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int k = 0; k < K; k++)
        s += a[k] + a2[k].x + a2[k].y + b[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0)
        *cycles = t1 - t0;
}

template <int K, int MODE>
void run(const char *name, int warps_per_block, float *out, float *in, long long *cyc)
{
    const int iters = 4096;
    chains<K, MODE><<<148, 32 * warps_per_block>>>(out, in, iters, cyc);
    chains<K, MODE><<<148, 32 * warps_per_block>>>(out, in, iters, cyc);
    cudaDeviceSynchronize();
    long long h;
    cudaMemcpy(&h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    const double ops = (double)iters * 8 * K * (MODE == 3 ? 5 : 1);
    printf("%-10s chains=%d warps/SM=%d : %.2f cycles per warp-instruction (%.2f per chain step)\n", name, K, warps_per_block, h / ops,
           h / ((double)iters * 8));
}

int main()
{
    float *in, *out;
    long long *cyc;
    cudaMalloc(&in, 4096 * 4);
    cudaMemset(in, 0, 4096 * 4);
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&cyc, 8);
    for (int w = 1; w <= 4; w *= 4) {
        run<1, 0>("ffma", w, out, in, cyc);
        run<2, 0>("ffma", w, out, in, cyc);
        run<4, 0>("ffma", w, out, in, cyc);
        run<8, 0>("ffma", w, out, in, cyc);
        run<1, 1>("ffma2", w, out, in, cyc);
        run<2, 1>("ffma2", w, out, in, cyc);
        run<4, 1>("ffma2", w, out, in, cyc);
        run<8, 1>("ffma2", w, out, in, cyc);
        run<1, 2>("fadd", w, out, in, cyc);
        run<4, 2>("fadd", w, out, in, cyc);
        run<8, 2>("fadd", w, out, in, cyc);
        run<4, 4>("ffma_sh1", w, out, in, cyc);
        run<8, 4>("ffma_sh1", w, out, in, cyc);
        run<8, 5>("ffma_sh2", w, out, in, cyc);
        run<8, 6>("ffma_pair", w, out, in, cyc);
        run<1, 3>("section", w, out, in, cyc);
        run<2, 3>("section", w, out, in, cyc);
        run<4, 3>("section", w, out, in, cyc);
    }
    return 0;
}
