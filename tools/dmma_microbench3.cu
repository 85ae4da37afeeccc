// Per-warp DMMA issue rate: pure register DMMA loops with 1, 2, 3, 4 warps per SM sub-partition.
#include <cstdio>
#include <cuda_runtime.h>
template <int NACC>
__global__ void dmma_loop(double *out, int iters, double a0, double b0) {
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
int main() {
    double *out; cudaMalloc(&out, 148 * 1024 * sizeof(double));
    const int iters = 20000;
    for (int threads = 32; threads <= 512; threads += (threads < 128 ? 32 : 128)) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        dmma_loop<16><<<148, threads>>>(out, iters, 1.0, 2.0); cudaDeviceSynchronize();
        cudaEventRecord(e0); dmma_loop<16><<<148, threads>>>(out, iters, 1.0, 2.0); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double flops = 2.0 * 256 * 16.0 * iters * 148.0 * (threads / 32);
        printf("%3d threads/SM (%d warps): %.2f TFLOP/s  (%.1f cycles per DMMA per warp @1965MHz)\n", threads, threads / 32, flops / ms / 1e9,
               ms * 1e-3 * 1.965e9 / (16.0 * iters));
    }
    return 0;
}
