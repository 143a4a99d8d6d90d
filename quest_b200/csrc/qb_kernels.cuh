// qb_kernels.cuh -- the three generic "direct" kernel shapes every amplitude update reduces to.
//
//   k_tuple : an item owns M amplitudes (pair / quadruple / ...): gather -> mix in registers -> scatter.
//   k_map   : an item owns one amplitude, optionally combined with one amplitude of a second array
//             (the communication buffer or another Qureg).
//   k_fill  : an item writes one amplitude without reading.
//
// All loads of a thread's ITEMS items are issued before any store, so each thread keeps ITEMS*M
// independent 128-bit loads in flight (HBM latency hiding), and consecutive threads touch consecutive
// item numbers so that every warp access is made of whole 32-byte sectors.
// The reference equivalents are the 39 one-work-item-per-thread kernels of quest/src/gpu/gpu_kernels.cuh.
#pragma once
#include "qb_common.cuh"

template <typename Op, int ITEMS>
__global__ void __launch_bounds__(QB_BLOCK) k_tuple(cplx* __restrict__ amps, qindex numItems, const Op op) {
    constexpr int M = Op::M;
    qindex idx[ITEMS][M];
    cplx v[ITEMS][M];
    const qindex base = (qindex)blockIdx.x * (QB_BLOCK * ITEMS) + threadIdx.x;
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
        const qindex n = base + (qindex)j * QB_BLOCK;
        if (ITEMS == 1 && n >= numItems) return;
        op.indices(n, idx[j]);
#pragma unroll
        for (int m = 0; m < M; m++) v[j][m] = amps[idx[j][m]];
    }
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
        op.apply(idx[j], v[j]);
#pragma unroll
        for (int m = 0; m < M; m++)
            if (op.writes(m)) amps[idx[j][m]] = v[j][m];
    }
}

template <typename Op>
static int qb_launch_tuple(cplx* amps, qindex numItems, const Op& op) {
    if (numItems <= 0) return 0;
    constexpr int ITEMS = (Op::M >= 4) ? 2 : 4;
    if (numItems >= (qindex)QB_BLOCK * ITEMS)
        k_tuple<Op, ITEMS><<<qb_grid(numItems, ITEMS), QB_BLOCK, 0, g_qb.stream>>>(amps, numItems, op);
    else
        k_tuple<Op, 1><<<qb_grid(numItems, 1), QB_BLOCK, 0, g_qb.stream>>>(amps, numItems, op);
    QB_LAUNCH_CHECK();
    return 0;
}

// Op: qindex index(n); cplx second(n, i); cplx apply(n, i, a, b)
template <typename Op, int ITEMS>
__global__ void __launch_bounds__(QB_BLOCK) k_map(cplx* __restrict__ amps, qindex numItems, const Op op) {
    qindex idx[ITEMS];
    cplx a[ITEMS], b[ITEMS];
    const qindex base = (qindex)blockIdx.x * (QB_BLOCK * ITEMS) + threadIdx.x;
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
        const qindex n = base + (qindex)j * QB_BLOCK;
        if (ITEMS == 1 && n >= numItems) return;
        idx[j] = op.index(n);
        a[j] = Op::READS ? amps[idx[j]] : mk(0, 0);
        b[j] = op.second(n, idx[j]);
    }
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
        const qindex n = base + (qindex)j * QB_BLOCK;
        amps[idx[j]] = op.apply(n, idx[j], a[j], b[j]);
    }
}

template <typename Op>
static int qb_launch_map(cplx* amps, qindex numItems, const Op& op) {
    if (numItems <= 0) return 0;
    constexpr int ITEMS = 4;
    if (numItems >= (qindex)QB_BLOCK * ITEMS)
        k_map<Op, ITEMS><<<qb_grid(numItems, ITEMS), QB_BLOCK, 0, g_qb.stream>>>(amps, numItems, op);
    else
        k_map<Op, 1><<<qb_grid(numItems, 1), QB_BLOCK, 0, g_qb.stream>>>(amps, numItems, op);
    QB_LAUNCH_CHECK();
    return 0;
}

// common argument validation for entry points taking (state, ctrls)
#define QB_CHECK_STATE(q) QB_REQUIRE((q) && (q)->amps && (q)->numAmpsPerNode > 0 && \
    ((q)->numAmpsPerNode == pow2((q)->logNumAmpsPerNode)), "bad qb_state")
#define QB_CHECK_SUFFIX(arr, n, q) QB_REQUIRE(qb_check_qubits(arr, n, (q)->logNumAmpsPerNode), "qubit list out of local range")
#define QB_CHECK_GLOBAL(arr, n) QB_REQUIRE(qb_check_qubits(arr, n, 63), "qubit list out of range")
