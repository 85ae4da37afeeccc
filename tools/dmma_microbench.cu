// Register-only ceilings of the FP64 pipes on this GPU: DMMA.8x8x4 (mma.sync m8n8k4 f64) and DFMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/dmma_microbench tools/dmma_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int NACC>
__global__ void __launch_bounds__(256) dmma_loop(double *out, int iters, double a0, double b0) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i][0] = c[i][1] = 0.0;
    double a = a0 + threadIdx.x, b = b0 - threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void __launch_bounds__(256) dfma_loop(double *out, int iters, double a0, double b0) {
    double c[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = i;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    double *out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(double) * 4);
    const int iters = 20000;
    for (int ctas = 1; ctas <= 4; ctas *= 2) {
        int grid = 148 * ctas;
        float ms = timeit([&] { dmma_loop<16><<<grid, 256>>>(out, iters, 1.0, 2.0); });
        double flops = 2.0 * 256 * 16.0 * iters * (double)grid * 8;   // 256 FMA per DMMA, 8 warps per CTA
        printf("DMMA 16 acc/warp, %d CTA/SM x 8 warps: %.2f TFLOP/s\n", ctas, flops / ms / 1e9);
        ms = timeit([&] { dmma_loop<4><<<grid, 256>>>(out, iters, 1.0, 2.0); });
        flops = 2.0 * 256 * 4.0 * iters * (double)grid * 8;
        printf("DMMA  4 acc/warp, %d CTA/SM x 8 warps: %.2f TFLOP/s\n", ctas, flops / ms / 1e9);
        ms = timeit([&] { dfma_loop<16><<<grid, 256>>>(out, iters, 1.000000001, 1e-9); });
        flops = 2.0 * 16.0 * iters * (double)grid * 256;
        printf("DFMA 16 chains/thread, %d CTA/SM x 256 thr: %.2f TFLOP/s\n", ctas, flops / ms / 1e9);
    }
    return 0;
}
