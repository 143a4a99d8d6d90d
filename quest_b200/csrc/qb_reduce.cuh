// qb_reduce.cuh -- deterministic grid reduction: warp shuffles -> block -> fixed-order final pass by the
// last block to finish (ticket counter).  Replaces the Thrust transform_reduce / inner_product calls of
// quest/src/gpu/gpu_thrust.cuh:745-1000.  The result is independent of block scheduling, so repeated
// runs give bit-identical probabilities (which keeps measurement outcomes reproducible).
#pragma once
#include "qb_common.cuh"

#define QB_RED_MAX_BLOCKS 4096
#define QB_RED_MAX_OUT 1024                       // doubles a single reduction may return
#define QB_RED_SCRATCH_DOUBLES (QB_RED_MAX_BLOCKS * 2 + QB_RED_MAX_BLOCKS * 64)

__device__ __forceinline__ double warp_sum(double v) {
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

// sums (re, im) over the block; result valid in thread 0
__device__ __forceinline__ void block_sum2(double& re, double& im, double* sm /* [2*8] */) {
    re = warp_sum(re);
    im = warp_sum(im);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) { sm[2 * w] = re; sm[2 * w + 1] = im; }
    __syncthreads();
    if (w == 0) {
        re = (l < QB_BLOCK / 32) ? sm[2 * l] : 0.0;
        im = (l < QB_BLOCK / 32) ? sm[2 * l + 1] : 0.0;
        re = warp_sum(re);
        im = warp_sum(im);
    }
}

// F: __device__ void operator()(qindex n, double& re, double& im) const  -- accumulates item n
template <typename F>
__global__ void __launch_bounds__(QB_BLOCK) k_reduce2(qindex numItems, F f, double* partials,
                                                     unsigned int* ticket, double* out) {
    __shared__ double sm[2 * (QB_BLOCK / 32)];
    __shared__ bool isLast;
    double re0 = 0, im0 = 0, re1 = 0, im1 = 0, re2 = 0, im2 = 0, re3 = 0, im3 = 0;
    const qindex stride = (qindex)gridDim.x * QB_BLOCK;
    qindex n = (qindex)blockIdx.x * QB_BLOCK + threadIdx.x;
    for (; n + 3 * stride < numItems; n += 4 * stride) {
        f(n, re0, im0);
        f(n + stride, re1, im1);
        f(n + 2 * stride, re2, im2);
        f(n + 3 * stride, re3, im3);
    }
    for (; n < numItems; n += stride) f(n, re0, im0);
    double re = (re0 + re1) + (re2 + re3);
    double im = (im0 + im1) + (im2 + im3);
    block_sum2(re, im, sm);
    if (threadIdx.x == 0) {
        partials[2 * blockIdx.x] = re;
        partials[2 * blockIdx.x + 1] = im;
        __threadfence();
        unsigned int t = atomicInc(ticket, gridDim.x - 1);   // wraps to 0 after the last block
        isLast = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!isLast) return;
    __threadfence();
    re = 0; im = 0;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += QB_BLOCK) {
        re += __ldcg(&partials[2 * b]);
        im += __ldcg(&partials[2 * b + 1]);
    }
    block_sum2(re, im, sm);
    if (threadIdx.x == 0) { out[0] = re; out[1] = im; }
}

// host: run the reduction and bring (re, im) back synchronously
template <typename F>
static int qb_reduce2(qindex numItems, F f, double* outRe, double* outIm) {
    QB_READY();
    qindex blocks = (numItems + QB_BLOCK - 1) / QB_BLOCK;
    qindex maxBlocks = (qindex)g_qb.numSMs * 8;
    if (maxBlocks > QB_RED_MAX_BLOCKS) maxBlocks = QB_RED_MAX_BLOCKS;
    if (blocks > maxBlocks) blocks = maxBlocks;
    if (blocks < 1) blocks = 1;
    k_reduce2<F><<<(unsigned int)blocks, QB_BLOCK, 0, g_qb.stream>>>(numItems, f, g_qb.redPartials,
                                                                   g_qb.redTicket, g_qb.redOutDev);
    QB_LAUNCH_CHECK();
    QB_CUDA(cudaMemcpyAsync(g_qb.redOutHost, g_qb.redOutDev, 2 * sizeof(double), cudaMemcpyDeviceToHost, g_qb.stream));
    QB_CUDA(cudaStreamSynchronize(g_qb.stream));
    if (outRe) *outRe = g_qb.redOutHost[0];
    if (outIm) *outIm = g_qb.redOutHost[1];
    return 0;
}
