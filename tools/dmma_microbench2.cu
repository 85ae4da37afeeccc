// Where does the ComplexF64 main loop lose DMMA issue slots? Same register pattern as CoreZ::compute
// (gett.cu): V1 registers only, V2 + LDS.128 fragment loads, V3 + one __syncthreads per k-block (BK=8).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double neg_bits(double x) { return __hiloint2double(__double2hiint(x) ^ 0x80000000, __double2loint(x)); }
constexpr int MT = 4, NT = 4, LDA = 130, LDB = 66, BK = 8;
template <int MODE>
__global__ void __launch_bounds__(256, 1) zloop(double *out, int kblocks) {
    extern __shared__ double2 sm[];
    double2 *sa = sm, *sb = sm + 4 * BK * LDA;
    for (int i = threadIdx.x; i < 4 * BK * (LDA + LDB); i += 256) sm[i] = make_double2(1e-3 * i, -1e-3 * i);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wm = (warp % 4) * 32, wn = (warp / 4) * 32, fr = lane >> 2, fk = lane & 3;
    double re[MT][NT][2] = {}, im[MT][NT][2] = {};
    double2 af[MT], bf[NT];
    for (int i = 0; i < MT; i++) af[i] = make_double2(1.0 + i + lane, 0.5 - i);
    for (int j = 0; j < NT; j++) bf[j] = make_double2(2.0 - j, 0.25 + j + lane);
    for (int kb = 0; kb < kblocks; kb++) {
        if (MODE >= 3) __syncthreads();
        const double2 *pa = sa + (kb & 3) * BK * LDA, *pb = sb + (kb & 3) * BK * LDB;
#pragma unroll
        for (int kk = 0; kk < BK / 4; kk++) {
            double nai[MT];
            if (MODE >= 2) {
#pragma unroll
                for (int i = 0; i < MT; i++) af[i] = pa[(kk * 4 + fk) * LDA + wm + i * 8 + fr];
#pragma unroll
                for (int j = 0; j < NT; j++) bf[j] = pb[(kk * 4 + fk) * LDB + wn + j * 8 + fr];
            }
#pragma unroll
            for (int i = 0; i < MT; i++) nai[i] = neg_bits(af[i].y);
#pragma unroll
            for (int i = 0; i < MT; i++)
#pragma unroll
                for (int j = 0; j < NT; j++) dmma(re[i][j][0], re[i][j][1], af[i].x, bf[j].x);
#pragma unroll
            for (int i = 0; i < MT; i++)
#pragma unroll
                for (int j = 0; j < NT; j++) dmma(im[i][j][0], im[i][j][1], af[i].x, bf[j].y);
#pragma unroll
            for (int i = 0; i < MT; i++)
#pragma unroll
                for (int j = 0; j < NT; j++) dmma(re[i][j][0], re[i][j][1], nai[i], bf[j].y);
#pragma unroll
            for (int i = 0; i < MT; i++)
#pragma unroll
                for (int j = 0; j < NT; j++) dmma(im[i][j][0], im[i][j][1], af[i].y, bf[j].x);
        }
    }
    double s = 0;
    for (int i = 0; i < MT; i++) for (int j = 0; j < NT; j++) s += re[i][j][0] + re[i][j][1] + im[i][j][0] + im[i][j][1];
    out[blockIdx.x * 256 + threadIdx.x] = s;
}
template <class F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    double *out; cudaMalloc(&out, 148 * 256 * sizeof(double));
    const int kblocks = 20000;
    const size_t smem = 4 * BK * (LDA + LDB) * sizeof(double2);
    cudaFuncSetAttribute(zloop<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(zloop<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(zloop<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    double flops = 2.0 * 256 * 128.0 * kblocks * 148.0 * 8;   // 128 DMMA per warp per k-block
    float ms = timeit([&] { zloop<1><<<148, 256, smem>>>(out, kblocks); });
    printf("V1 registers only          : %.2f TFLOP/s\n", flops / ms / 1e9);
    ms = timeit([&] { zloop<2><<<148, 256, smem>>>(out, kblocks); });
    printf("V2 + LDS.128 fragment loads: %.2f TFLOP/s\n", flops / ms / 1e9);
    ms = timeit([&] { zloop<3><<<148, 256, smem>>>(out, kblocks); });
    printf("V3 + __syncthreads / kblock: %.2f TFLOP/s\n", flops / ms / 1e9);
    return 0;
}
