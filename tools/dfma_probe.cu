// measures FP64 FMA throughput per SM per clock (independent chains, register-resident)
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP> __global__ void k(double* out, double a, double b, int iters) {
    double x[ILP];
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
    }
    double s = 0; for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    int dev = 0; cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    double* out; cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 1024 * 4);
    for (int threads : {256, 512, 1024}) {
        const int iters = 20000; constexpr int ILP = 8;
        k<ILP><<<p.multiProcessorCount, threads>>>(out, 1.0000001, 1e-9, 10);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k<ILP><<<p.multiProcessorCount, threads>>>(out, 1.0000001, 1e-9, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fmas = (double)p.multiProcessorCount * threads * ILP * iters;
        int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
        printf("threads/SM=%4d: %.2f TFMA/s = %.2f TFLOP/s FP64 ; per SM per clk (at %d MHz nominal): %.1f FMA\n", threads,
               fmas / ms / 1e9, 2 * fmas / ms / 1e9, clk / 1000, fmas / ms / 1e-3 / p.multiProcessorCount / (clk * 1e3));
    }
    return 0;
}
