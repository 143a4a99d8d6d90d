// FP64 throughput probes for the gate bodies of the tile engine (B200, sm_100a):
//   A. independent FMA chains with constant-bank operands      -> the FP64 pipe's peak
//   B. the same with all three operands in registers           -> register-operand cost
//   C. the 1-qubit dense butterfly on 16 register-resident complex amplitudes (the body of reg_dense1)
//   D. the 2-qubit dense butterfly (the body of reg_dense2)
// each at 8 / 16 / 32 warps per SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dfma_probe tools/dfma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP> __global__ void kA(double* out, double a, double b, int iters) {
    double x[ILP];
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
    }
    double s = 0; for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP> __global__ void kB(double* out, const double* in, int iters) {
    double x[ILP];
    const double a = in[threadIdx.x & 1], b = in[2 + (threadIdx.x & 1)];
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
    }
    double s = 0; for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

struct c2 { double x, y; };
__device__ __forceinline__ c2 cmul(c2 a, c2 b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__device__ __forceinline__ c2 cfma(c2 a, c2 b, c2 c) { return {fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y))}; }

__global__ void kC(double* out, const double* in, int iters) {
    c2 v[16], m[4];
    for (int i = 0; i < 4; i++) m[i] = {in[2 * i], in[2 * i + 1]};
    for (int i = 0; i < 16; i++) v[i] = {threadIdx.x * 1e-3 + i, 0.5 * i};
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 4; k++) {              // one gate per register bit
#pragma unroll
            for (int u = 0; u < 16; u++) {
                if (u & (1 << k)) continue;
                const int u1 = u | (1 << k);
                c2 a0 = v[u], a1 = v[u1];
                v[u] = cfma(m[1], a1, cmul(m[0], a0)); v[u1] = cfma(m[3], a1, cmul(m[2], a0));
            }
        }
    }
    double s = 0; for (int i = 0; i < 16; i++) s += v[i].x + v[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void kD(double* out, const double* in, int iters) {
    c2 v[16], m[16];
    for (int i = 0; i < 16; i++) m[i] = {in[2 * i], in[2 * i + 1]};
    for (int i = 0; i < 16; i++) v[i] = {threadIdx.x * 1e-3 + i, 0.5 * i};
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int pair = 0; pair < 2; pair++) {     // targets (0,1) then (2,3)
            const int K0 = 2 * pair, K1 = 2 * pair + 1;
#pragma unroll
            for (int u = 0; u < 16; u++) {
                if (u & ((1 << K0) | (1 << K1))) continue;
                const int i1 = u | (1 << K0), i2 = u | (1 << K1), i3 = i1 | i2;
                c2 a0 = v[u], a1 = v[i1], a2 = v[i2], a3 = v[i3];
                v[u]  = cfma(m[3], a3, cfma(m[2], a2, cfma(m[1], a1, cmul(m[0], a0))));
                v[i1] = cfma(m[7], a3, cfma(m[6], a2, cfma(m[5], a1, cmul(m[4], a0))));
                v[i2] = cfma(m[11], a3, cfma(m[10], a2, cfma(m[9], a1, cmul(m[8], a0))));
                v[i3] = cfma(m[15], a3, cfma(m[14], a2, cfma(m[13], a1, cmul(m[12], a0))));
            }
        }
    }
    double s = 0; for (int i = 0; i < 16; i++) s += v[i].x + v[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    int dev = 0; cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    double* out; cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 1024 * 4);
    double hin[32]; for (int i = 0; i < 32; i++) hin[i] = (i % 3 == 0) ? 0.7071 : ((i % 3 == 1) ? -0.5 : 0.25);
    double* in; cudaMalloc(&in, sizeof hin); cudaMemcpy(in, hin, sizeof hin, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto report = [&](const char* name, int threads, double fmas, float ms) {
        printf("%-46s threads/SM=%4d: %6.2f TFLOP/s FP64, %5.1f FP64-instr lanes/clk/SM (at %d MHz)\n", name, threads,
               2 * fmas / ms / 1e9, fmas / (ms * 1e-3) / p.multiProcessorCount / (clk * 1e3), clk / 1000);
    };
    const int sms = p.multiProcessorCount;
    for (int threads : {256, 512, 1024}) {
        float ms; const int iters = 20000;
        kA<8><<<sms, threads>>>(out, 1.0000001, 1e-9, 10);
        cudaEventRecord(e0); kA<8><<<sms, threads>>>(out, 1.0000001, 1e-9, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); report("A: 8 chains, constant operands", threads, (double)sms * threads * 8 * iters, ms);
        kB<8><<<sms, threads>>>(out, in, 10);
        cudaEventRecord(e0); kB<8><<<sms, threads>>>(out, in, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); report("B: 8 chains, register operands", threads, (double)sms * threads * 8 * iters, ms);
    }
    for (int threads : {128, 256, 384, 512}) {
        float ms; const int iters = 2000;
        kC<<<sms, threads>>>(out, in, 10);
        cudaEventRecord(e0); kC<<<sms, threads>>>(out, in, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); report("C: dense1 butterfly x4 on 16 reg amps", threads, (double)sms * threads * 4 * 128 * iters, ms);
        kD<<<sms, threads>>>(out, in, 10);
        cudaEventRecord(e0); kD<<<sms, threads>>>(out, in, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); report("D: dense2 butterfly x2 on 16 reg amps", threads, (double)sms * threads * 2 * 256 * iters, ms);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
