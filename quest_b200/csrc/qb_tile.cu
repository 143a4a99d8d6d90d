// qb_tile.cu -- the tile engine: deferred gate queue + planner + persistent TMA-pipelined fused-gate kernel.
//
// WHY.  A single gate kernel already moves the state at ~85% of the HBM peak (profiles/r1_*), so gates/s can only
// grow by applying SEVERAL gates per pass over HBM.  QuEST's API issues one gate per call, therefore the backend
// defers: the per-gate entry points (qb_gates.cu) append the gate to a host-side queue and return; any entry point
// that observes or otherwise touches device memory (reductions, copies, exchanges, qb_sync, non-fusable gates ...)
// first flushes the queue (QB_READY does it), so the deferral is invisible through the C ABI.
//
// PLAN.  Consecutive queued gates are greedily grouped into PASSES.  A pass owns a set S of T = 12 index bits:
// bits 0..5 (so the state is always touched in >= 1 KiB contiguous runs) plus up to six arbitrary higher bits.  All
// NON-diagonal targets of the pass's gates must lie in S; controls and diagonal targets may lie anywhere.  The 2^n
// amplitudes then split into 2^(n-12) independent TILES of 4096 amplitudes (64 KiB): fix the bits outside S.
// Runs of controlled phase shifts sharing one qubit (the QFT's ladders, api/operations.cpp:1934-1953, which the
// reference marks as a todo to merge) are folded into ONE "phase star" op evaluated from small lookup tables.
//
// KERNEL.  One persistent CTA per SM (grid = #SMs), 512 threads, 3 stages x 64 KiB of shared memory.  Thread 0
// drives the copy engine: cp.async.bulk (TMA, 1-D) loads the chunks of tile k+2 into a free stage, signalling an
// mbarrier with complete_tx; all threads wait on the stage's mbarrier, apply the pass's gates to the tile in shared
// memory (__syncthreads between gates), fence.proxy.async, and thread 0 writes the stage back with
// cp.async.bulk.global.shared::cta (bulk async-group).  Loads of two tiles and the store of one are in flight while a
// tile is being computed, so HBM stays busy; per tile the gates cost shared-memory bandwidth only.
// Algorithmic bytes per pass: every gate of the pass counts its own 2*16*N/2^c bytes, physical bytes
// are 2*16*N once per pass (or less: controls shared by every gate of a pass prune whole tiles before they are loaded).
#include "qb_common.cuh"
#include "qb_kernels.cuh"
#include "qb_tile.cuh"
#include <math.h>
#include <string.h>
#include <vector>
#include <algorithm>

#define TILE_BITS 12
#define TILE_LOW 6
#define TILE_AMPS (1 << TILE_BITS)
#define TILE_STAGES 3
#define TILE_THREADS 512
#define TILE_MAX_CHUNKS (1 << (TILE_BITS - TILE_LOW))
#define QUEUE_MAX 512
#define FUSE_MIN_LOG_AMPS 13          // smaller states use the direct kernels immediately
#define STAR_SEGS 6                   // external-control tables: 6 segments x 6 bits of the global index

enum { OP_DENSE1 = 0, OP_DENSE2, OP_PAULI, OP_SWAP, OP_DIAG, OP_PARITY, OP_STAR };

// ------------------------------------------------------------------------------------------
// queued gate, in global (local-shard) qubit coordinates
// ------------------------------------------------------------------------------------------
struct QOp {
    int kind;
    unsigned long long ctrlMask, ctrlVals;     // suffix controls
    int t0, t1, numT;                          // targets; diag targets may be >= logN (prefix): resolved at enqueue
    unsigned long long maskA, maskB;           // pauli: XY, YZ; parity gadget: target mask
    cplx m[16];                                // matrix / diagonal / factors
    std::vector<std::pair<int,double>> star;   // OP_STAR: (other qubit, theta); centre qubit = t0
    qindex algBytes;
};

static std::vector<QOp> s_queue;
static qb_state s_qstate;                      // identity of the state the queue refers to
static bool s_qvalid = false;
static int s_status = 0;
static bool s_inFlush = false;

// device-side op, in tile coordinates
struct TileOp {
    int kind, p0, p1, numT;                    // in-tile positions (p < 0: target outside the tile, diag only)
    int e0, e1;                                // global bit index of an external diag target
    unsigned int inCtrlMask, inCtrlVals;
    unsigned long long extCtrlMask, extCtrlVals;
    unsigned int inMaskA, inMaskB;
    unsigned long long extMaskB;               // pauli / parity: sign bits outside the tile
    int tab;                                   // OP_STAR: index of its table block
    int pad;
    cplx m[16];
};

struct StarTab { cplx in[2][64]; cplx ext[STAR_SEGS][64]; };

struct PassHdr {
    int numOps, numChunks, chunkAmps, pad;
    qindex numTiles;
    BitIns tileIns;                            // tile number -> global base index (S bits zero, pruning controls set)
    qindex chunkOff[TILE_MAX_CHUNKS];          // global offset of chunk c inside a tile
};

// ------------------------------------------------------------------------------------------
// PTX helpers (sm_100a): mbarrier + 1-D bulk tensor-memory-accelerator copies
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}"
        :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load(void* smemDst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(smemDst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store(void* gdst, const void* smemSrc, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(gdst), "r"(smem_u32(smemSrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ unsigned ins0(unsigned v, int p) { return ((v >> p) << (p + 1)) | (v & ((1u << p) - 1u)); }

// ------------------------------------------------------------------------------------------
// applying one op to one tile held in shared memory
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void apply_op(cplx* __restrict__ t, const TileOp& op, qindex base, const StarTab* __restrict__ tabs, cplx* scratch) {
    const int tid = threadIdx.x;
    const unsigned cm = op.inCtrlMask, cv = op.inCtrlVals;
    switch (op.kind) {
    case OP_DENSE1: {
        const int p = op.p0; const unsigned bit = 1u << p;
        const cplx m00 = op.m[0], m01 = op.m[1], m10 = op.m[2], m11 = op.m[3];
        for (unsigned n = tid; n < TILE_AMPS / 2; n += TILE_THREADS) {
            unsigned j0 = ins0(n, p);
            if ((j0 & cm) != cv) continue;
            cplx a0 = t[j0], a1 = t[j0 | bit];
            t[j0] = cfma(m01, a1, cmul(m00, a0));
            t[j0 | bit] = cfma(m11, a1, cmul(m10, a0));
        }
    } break;
    case OP_DENSE2: {
        const int lo = min(op.p0, op.p1), hi = max(op.p0, op.p1);
        const unsigned b0 = 1u << op.p0, b1 = 1u << op.p1;
        for (unsigned n = tid; n < TILE_AMPS / 4; n += TILE_THREADS) {
            unsigned j = ins0(ins0(n, lo), hi);
            if ((j & cm) != cv) continue;
            cplx a0 = t[j], a1 = t[j | b0], a2 = t[j | b1], a3 = t[j | b0 | b1];
#pragma unroll
            for (int r = 0; r < 4; r++) {
                cplx v = cfma(op.m[4 * r + 3], a3, cfma(op.m[4 * r + 2], a2, cfma(op.m[4 * r + 1], a1, cmul(op.m[4 * r], a0))));
                t[j | ((r & 1) ? b0 : 0u) | ((r & 2) ? b1 : 0u)] = v;
            }
        }
    } break;
    case OP_PAULI: {
        const unsigned xy = op.inMaskA, yz = op.inMaskB;
        const int h = 31 - __clz(xy);
        const int extPar = parity64((unsigned long long)base & op.extMaskB);
        const cplx af = op.m[0], pf = op.m[1];
        for (unsigned n = tid; n < TILE_AMPS / 2; n += TILE_THREADS) {
            unsigned jA = ins0(n, h);
            if ((jA & cm) != cv) continue;
            unsigned jB = jA ^ xy;
            double sA = 1.0 - 2.0 * ((__popc(jA & yz) + extPar) & 1);
            double sB = 1.0 - 2.0 * ((__popc(jB & yz) + extPar) & 1);
            cplx a = t[jA], b = t[jB];
            t[jA] = cfma(pf, cscale(sB, b), cmul(af, a));
            t[jB] = cfma(pf, cscale(sA, a), cmul(af, b));
        }
    } break;
    case OP_SWAP: {
        const int lo = min(op.p0, op.p1), hi = max(op.p0, op.p1);
        const unsigned b0 = 1u << op.p0, b1 = 1u << op.p1;
        for (unsigned n = tid; n < TILE_AMPS / 4; n += TILE_THREADS) {
            unsigned j = ins0(ins0(n, lo), hi);
            if ((j & cm) != cv) continue;
            cplx a = t[j | b0], b = t[j | b1];
            t[j | b0] = b; t[j | b1] = a;
        }
    } break;
    case OP_DIAG: {
        // element index bit k <- target k; external targets read their (tile-constant) bit from the base index
        const int x0 = (op.p0 < 0) ? getBit(base, op.e0) : 0;
        const int x1 = (op.numT > 1 && op.p1 < 0) ? getBit(base, op.e1) : 0;
        for (unsigned j = tid; j < TILE_AMPS; j += TILE_THREADS) {
            if ((j & cm) != cv) continue;
            int k = (op.p0 < 0) ? x0 : ((j >> op.p0) & 1);
            if (op.numT > 1) k |= ((op.p1 < 0) ? x1 : ((j >> op.p1) & 1)) << 1;
            t[j] = cmul(t[j], op.m[k]);
        }
    } break;
    case OP_PARITY: {
        const int extPar = parity64((unsigned long long)base & op.extMaskB);
        for (unsigned j = tid; j < TILE_AMPS; j += TILE_THREADS) {
            if ((j & cm) != cv) continue;
            int par = (__popc(j & op.inMaskA) + extPar) & 1;
            t[j] = cmul(t[j], op.m[par]);
        }
    } break;
    case OP_STAR: {
        // all amplitudes with centre bit = 1 gain exp(i * sum_c theta_c * bit_c): product of per-segment table entries
        const StarTab& tb = tabs[op.tab];
        if (op.p0 < 0 && !getBit(base, op.e0)) break;
        if (tid == 0) {
            cplx f = mk(1, 0);
#pragma unroll
            for (int s = 0; s < STAR_SEGS; s++) f = cmul(f, tb.ext[s][(base >> (6 * s)) & 63]);
            scratch[0] = f;
        }
        __syncthreads();
        const cplx f = scratch[0];
        if (op.p0 >= 0) {
            const int p = op.p0;
            for (unsigned n = tid; n < TILE_AMPS / 2; n += TILE_THREADS) {
                unsigned j = ins0(n, p) | (1u << p);
                cplx e = cmul(cmul(tb.in[0][j & 63], tb.in[1][j >> 6]), f);
                t[j] = cmul(t[j], e);
            }
        } else {
            for (unsigned j = tid; j < TILE_AMPS; j += TILE_THREADS) {
                cplx e = cmul(cmul(tb.in[0][j & 63], tb.in[1][j >> 6]), f);
                t[j] = cmul(t[j], e);
            }
        }
        __syncthreads();       // scratch[0] is reused by the next star
    } break;
    }
}

__global__ void __launch_bounds__(TILE_THREADS, 1) k_tile_pass(cplx* __restrict__ amps, const PassHdr* __restrict__ hdrp,
                                                               const TileOp* __restrict__ gops, const StarTab* __restrict__ tabs) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cplx* stageBuf = reinterpret_cast<cplx*>(smem_raw);                                  // [STAGES][TILE_AMPS]
    TileOp* ops = reinterpret_cast<TileOp*>(smem_raw + (size_t)TILE_STAGES * TILE_AMPS * sizeof(cplx));
    __shared__ unsigned long long full[TILE_STAGES];
    __shared__ PassHdr hdr;
    __shared__ cplx scratch[2];

    const int tid = threadIdx.x;
    for (int i = tid; i < (int)(sizeof(PassHdr) / 4); i += TILE_THREADS) ((int*)&hdr)[i] = ((const int*)hdrp)[i];
    __syncthreads();
    const int numOps = hdr.numOps;
    for (int i = tid; i < numOps * (int)(sizeof(TileOp) / 4); i += TILE_THREADS) ((int*)ops)[i] = ((const int*)gops)[i];
    if (tid == 0) {
        for (int s = 0; s < TILE_STAGES; s++) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const qindex numTiles = hdr.numTiles;
    const int numChunks = hdr.numChunks;
    const unsigned chunkBytes = (unsigned)hdr.chunkAmps * (unsigned)sizeof(cplx);
    const qindex myCount = (numTiles > (qindex)blockIdx.x) ? (numTiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    auto issue_load = [&](qindex k) {          // thread 0 only
        const int s = (int)(k % TILE_STAGES);
        const qindex base = hdr.tileIns((qindex)blockIdx.x + k * gridDim.x);
        mbar_expect_tx(&full[s], TILE_AMPS * (unsigned)sizeof(cplx));
        cplx* dst = stageBuf + (size_t)s * TILE_AMPS;
        for (int c = 0; c < numChunks; c++)
            tma_load(dst + (size_t)c * hdr.chunkAmps, amps + base + hdr.chunkOff[c], chunkBytes, &full[s]);
    };

    if (tid == 0) {
        if (myCount > 0) issue_load(0);
        if (myCount > 1) issue_load(1);
    }

    for (qindex k = 0; k < myCount; k++) {
        const int s = (int)(k % TILE_STAGES);
        const unsigned parity = (unsigned)((k / TILE_STAGES) & 1);
        const qindex base = hdr.tileIns((qindex)blockIdx.x + k * gridDim.x);
        cplx* t = stageBuf + (size_t)s * TILE_AMPS;
        mbar_wait(&full[s], parity);

        for (int o = 0; o < numOps; o++) {
            const TileOp& op = ops[o];
            if (((unsigned long long)base & op.extCtrlMask) == op.extCtrlVals)     // tile-uniform: no divergence
                apply_op(t, op, base, tabs, scratch);
            __syncthreads();
        }
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            for (int c = 0; c < numChunks; c++)
                tma_store(amps + base + hdr.chunkOff[c], t + (size_t)c * hdr.chunkAmps, chunkBytes);
            tma_commit();
            // the stage that tile k+2 will use held tile k-1, whose store was committed one iteration ago
            tma_wait_read<1>();
            if (k + 2 < myCount) issue_load(k + 2);
        }
    }
    if (tid == 0) tma_wait_all<0>();
}

// ------------------------------------------------------------------------------------------
// host: planner
// ------------------------------------------------------------------------------------------
struct Pass { std::vector<int> opIdx; unsigned long long high = 0; };

static inline unsigned long long nonDiagTargets(const QOp& o) {
    switch (o.kind) {
    case OP_DENSE1: return 1ULL << o.t0;
    case OP_DENSE2: case OP_SWAP: return (1ULL << o.t0) | (1ULL << o.t1);
    case OP_PAULI: return o.maskA;
    default: return 0;
    }
}

// run the direct (unfused) kernel for one queued op
static int run_direct(const qb_state* q, const QOp& o);

static int s_maxOpsPerPass = 48;

static int emit_pass(const qb_state* q, const std::vector<QOp>& ops, const Pass& pass,
                     std::vector<PassHdr>& hdrs, std::vector<TileOp>& tops, std::vector<StarTab>& tabs, std::vector<int>& opCount) {
    const int n = q->logNumAmpsPerNode;
    // tile bit set S: the low bits, the required high bits, then filler (lowest unused bits) up to TILE_BITS
    unsigned long long S = ((1ULL << TILE_LOW) - 1) | pass.high;
    for (int b = TILE_LOW; b < n && __builtin_popcountll(S) < TILE_BITS; b++) S |= 1ULL << b;
    int pos[64]; int sbits[TILE_BITS]; int T = 0;
    for (int b = 0; b < 64; b++) pos[b] = -1;
    for (int b = 0; b < n; b++) if ((S >> b) & 1) { pos[b] = T; sbits[T++] = b; }
    // controls shared (same qubit, same value) by EVERY op of the pass and lying outside S prune whole tiles
    unsigned long long common = ~0ULL, commonVals = 0;
    for (size_t i = 0; i < pass.opIdx.size(); i++) {
        const QOp& o = ops[pass.opIdx[i]];
        unsigned long long m = o.ctrlMask & ~S;
        if (i == 0) { common = m; commonVals = o.ctrlVals & m; }
        else { common &= m; common &= ~((o.ctrlVals ^ commonVals) & common); commonVals &= common; }
    }
    if (pass.opIdx.empty()) common = 0;

    PassHdr h; memset(&h, 0, sizeof h);
    int contiguous = 0; while (contiguous < T && sbits[contiguous] == contiguous) contiguous++;
    h.chunkAmps = 1 << contiguous;
    h.numChunks = 1 << (T - contiguous);
    for (int c = 0; c < h.numChunks; c++) {
        qindex off = 0;
        for (int b = contiguous; b < T; b++) if ((c >> (b - contiguous)) & 1) off |= (qindex)1 << sbits[b];
        h.chunkOff[c] = off;
    }
    // tile enumeration inserts zeros at S bits and the pruning-control bits (with their required values)
    int fixedQ[64], fixedV[64], nf = 0;
    for (int b = 0; b < n; b++) {
        if ((S >> b) & 1) { fixedQ[nf] = b; fixedV[nf++] = 0; }
        else if ((common >> b) & 1) { fixedQ[nf] = b; fixedV[nf++] = (int)((commonVals >> b) & 1); }
    }
    h.tileIns = qb_make_ins(fixedQ, fixedV, nf, nullptr, nullptr, 0);
    h.numTiles = (qindex)1 << (n - nf);
    h.numOps = (int)pass.opIdx.size();

    const unsigned long long inMaskAll = S;
    auto toIn = [&](unsigned long long gm) { unsigned v = 0; for (int p = 0; p < T; p++) if ((gm >> sbits[p]) & 1) v |= 1u << p; return v; };
    for (int idx : pass.opIdx) {
        const QOp& o = ops[idx];
        TileOp t; memset(&t, 0, sizeof t);
        t.kind = o.kind; t.numT = o.numT;
        t.inCtrlMask = toIn(o.ctrlMask & inMaskAll); t.inCtrlVals = toIn(o.ctrlVals & o.ctrlMask & inMaskAll);
        t.extCtrlMask = o.ctrlMask & ~inMaskAll; t.extCtrlVals = o.ctrlVals & t.extCtrlMask;
        t.p0 = t.p1 = -1; t.e0 = t.e1 = 0;
        for (int i = 0; i < 16; i++) t.m[i] = o.m[i];
        switch (o.kind) {
        case OP_DENSE1: t.p0 = pos[o.t0]; break;
        case OP_DENSE2: case OP_SWAP: t.p0 = pos[o.t0]; t.p1 = pos[o.t1]; break;
        case OP_PAULI: t.inMaskA = toIn(o.maskA); t.inMaskB = toIn(o.maskB & inMaskAll); t.extMaskB = o.maskB & ~inMaskAll; break;
        case OP_PARITY: t.inMaskA = toIn(o.maskA & inMaskAll); t.extMaskB = o.maskA & ~inMaskAll; break;
        case OP_DIAG:
            t.p0 = (o.t0 < n) ? pos[o.t0] : -1; t.e0 = o.t0;
            if (o.numT > 1) { t.p1 = (o.t1 < n) ? pos[o.t1] : -1; t.e1 = o.t1; }
            break;
        case OP_STAR: {
            t.p0 = pos[o.t0]; t.e0 = o.t0;
            StarTab tb;
            long double angIn[2][64] = {{0}}, angExt[STAR_SEGS][64] = {{0}};
            for (auto& ce : o.star) {
                int c = ce.first; long double th = ce.second;
                if (pos[c] >= 0) { int p = pos[c]; for (int v = 0; v < 64; v++) if ((v >> (p % 6)) & 1) angIn[p / 6][v] += th; }
                else { for (int v = 0; v < 64; v++) if ((v >> (c % 6)) & 1) angExt[c / 6][v] += th; }
            }
            for (int s = 0; s < 2; s++) for (int v = 0; v < 64; v++) tb.in[s][v] = mk((double)cosl(angIn[s][v]), (double)sinl(angIn[s][v]));
            for (int s = 0; s < STAR_SEGS; s++) for (int v = 0; v < 64; v++) tb.ext[s][v] = mk((double)cosl(angExt[s][v]), (double)sinl(angExt[s][v]));
            t.tab = (int)tabs.size();
            tabs.push_back(tb);
        } break;
        }
        tops.push_back(t);
    }
    hdrs.push_back(h);
    opCount.push_back(h.numOps);
    return 0;
}

static bool is_cphase(const QOp& o) {
    if (o.kind != OP_DIAG || o.numT != 1) return false;
    if (__builtin_popcountll(o.ctrlMask) != 1 || o.ctrlVals != o.ctrlMask) return false;
    if (o.m[0].x != 1.0 || o.m[0].y != 0.0) return false;
    return fabs(hypot(o.m[1].x, o.m[1].y) - 1.0) < 1e-14;
}

// device scratch for pass descriptors (grown on demand)
static char* s_devDesc = nullptr; static size_t s_devDescBytes = 0;

static int flush_queue() {
    if (s_queue.empty() || s_inFlush) return 0;
    s_inFlush = true;
    std::vector<QOp> ops; ops.swap(s_queue);
    qb_state q = s_qstate; s_qvalid = false;
    const int n = q.logNumAmpsPerNode;
    int rc = 0;

    // 1. merge ladders of controlled phases that share a qubit into phase stars
    std::vector<QOp> merged;
    for (size_t i = 0; i < ops.size(); ) {
        if (is_cphase(ops[i]) && ops[i].t0 < n) {
            size_t j = i + 1;
            int a = ops[i].t0, b = __builtin_ctzll(ops[i].ctrlMask), centre = -1;
            while (j < ops.size() && is_cphase(ops[j]) && ops[j].t0 < n) {
                int c = ops[j].t0, d = __builtin_ctzll(ops[j].ctrlMask);
                if (centre < 0) { if (c == a || d == a) centre = a; else if (c == b || d == b) centre = b; else break; }
                else if (c != centre && d != centre) break;
                j++;
            }
            if (j - i >= 2) {
                QOp s; s.kind = OP_STAR; s.ctrlMask = s.ctrlVals = 0; s.t0 = centre; s.t1 = 0; s.numT = 1; s.maskA = s.maskB = 0; s.algBytes = 0;
                for (int z = 0; z < 16; z++) s.m[z] = mk(0, 0);
                for (size_t k = i; k < j; k++) {
                    int c = ops[k].t0, d = __builtin_ctzll(ops[k].ctrlMask);
                    s.star.push_back({c == centre ? d : c, atan2(ops[k].m[1].y, ops[k].m[1].x)});
                    s.algBytes += ops[k].algBytes;
                }
                merged.push_back(s);
                i = j;
                continue;
            }
        }
        merged.push_back(ops[i++]);
    }

    // 2. greedy grouping into passes
    std::vector<Pass> passes;
    Pass cur;
    const int maxHigh = TILE_BITS - TILE_LOW;
    const unsigned long long lowMask = (1ULL << TILE_LOW) - 1;
    for (size_t i = 0; i < merged.size(); i++) {
        unsigned long long need = nonDiagTargets(merged[i]) & ~lowMask;
        if (merged[i].kind == OP_STAR && merged[i].t0 >= TILE_LOW) need |= 1ULL << merged[i].t0;   // keep the centre in-tile when cheap
        unsigned long long nh = cur.high | need;
        if (!cur.opIdx.empty() && (__builtin_popcountll(nh) > maxHigh || (int)cur.opIdx.size() >= s_maxOpsPerPass)) {
            passes.push_back(cur); cur = Pass(); nh = need;
        }
        if (__builtin_popcountll(nh) > maxHigh) {          // a single op that cannot fit (e.g. Pauli string on > 6 high qubits)
            if (!cur.opIdx.empty()) { passes.push_back(cur); cur = Pass(); }
            Pass solo; solo.opIdx.push_back((int)i); solo.high = ~0ULL;   // marker: run direct
            passes.push_back(solo);
            continue;
        }
        cur.high = nh; cur.opIdx.push_back((int)i);
    }
    if (!cur.opIdx.empty()) passes.push_back(cur);

    // 3. emit: single-op passes use the direct kernels (already at the HBM roofline), multi-op passes the tile kernel
    std::vector<PassHdr> hdrs; std::vector<TileOp> tops; std::vector<StarTab> tabs; std::vector<int> opCount;
    std::vector<int> passKind;     // -1: direct op index, else index into hdrs
    std::vector<int> passArg;
    for (auto& p : passes) {
        bool direct = (p.high == ~0ULL) || (p.opIdx.size() == 1 && merged[p.opIdx[0]].kind != OP_STAR) || n < TILE_BITS;
        if (direct) { for (int idx : p.opIdx) { passKind.push_back(-1); passArg.push_back(idx); } }
        else { passKind.push_back((int)hdrs.size()); passArg.push_back(0); emit_pass(&q, merged, p, hdrs, tops, tabs, opCount); }
    }
    if (!hdrs.empty()) {
        size_t bh = hdrs.size() * sizeof(PassHdr), bo = tops.size() * sizeof(TileOp), bt = tabs.size() * sizeof(StarTab);
        size_t need = bh + bo + bt + 256;
        if (need > s_devDescBytes) {
            cudaStreamSynchronize(g_qb.stream);
            if (s_devDesc) cudaFree(s_devDesc);
            s_devDescBytes = need * 2;
            if (cudaMalloc(&s_devDesc, s_devDescBytes) != cudaSuccess) { s_devDesc = nullptr; s_devDescBytes = 0; rc = qb_set_error((int)cudaErrorMemoryAllocation, "tile descriptors", __FILE__, __LINE__); }
        }
        if (!rc) {
            // pageable source: the runtime stages the bytes before returning, so the vectors may die right after
            cudaMemcpyAsync(s_devDesc, hdrs.data(), bh, cudaMemcpyHostToDevice, g_qb.stream);
            cudaMemcpyAsync(s_devDesc + bh, tops.data(), bo, cudaMemcpyHostToDevice, g_qb.stream);
            if (bt) cudaMemcpyAsync(s_devDesc + bh + bo, tabs.data(), bt, cudaMemcpyHostToDevice, g_qb.stream);
        }
    }
    static bool attrSet = false;
    const size_t smemBytes = (size_t)TILE_STAGES * TILE_AMPS * sizeof(cplx) + (size_t)s_maxOpsPerPass * sizeof(TileOp);
    if (!attrSet && !rc) {
        cudaError_t e = cudaFuncSetAttribute(k_tile_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes);
        if (e != cudaSuccess) rc = qb_set_error((int)e, "cudaFuncSetAttribute(k_tile_pass)", __FILE__, __LINE__);
        attrSet = true;
    }
    size_t opBase = 0;
    for (size_t i = 0; i < passKind.size() && !rc; i++) {
        if (passKind[i] < 0) { rc = run_direct(&q, merged[passArg[i]]); continue; }
        int hi = passKind[i];
        const PassHdr* dh = (const PassHdr*)s_devDesc + hi;
        const TileOp* dops = (const TileOp*)(s_devDesc + hdrs.size() * sizeof(PassHdr)) + opBase;
        const StarTab* dt = (const StarTab*)(s_devDesc + hdrs.size() * sizeof(PassHdr) + tops.size() * sizeof(TileOp));
        qindex tiles = hdrs[hi].numTiles;
        unsigned grid = (unsigned)std::min<qindex>(tiles, g_qb.numSMs);
        k_tile_pass<<<grid, TILE_THREADS, smemBytes, g_qb.stream>>>((cplx*)q.amps, dh, dops, dt);
        g_qb.launches++;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) rc = qb_set_error((int)e, "k_tile_pass launch", __FILE__, __LINE__);
        opBase += opCount[hi];
    }
    s_inFlush = false;
    if (rc) s_status = rc;
    return rc;
}

int qb_flush_internal() { return flush_queue(); }

// ------------------------------------------------------------------------------------------
// enqueue API used by the per-gate entry points
// ------------------------------------------------------------------------------------------
static bool can_fuse(const qb_state* q) {
    return g_qb.tileEngine && !s_inFlush && q->logNumAmpsPerNode >= FUSE_MIN_LOG_AMPS && q->logNumAmpsPerNode <= 36;
}

static bool same_state(const qb_state* q) {
    return s_qvalid && s_qstate.amps == q->amps && s_qstate.numAmpsPerNode == q->numAmpsPerNode && s_qstate.rank == q->rank;
}

static int enqueue(const qb_state* q, QOp& o) {
    if (!same_state(q)) { int r = flush_queue(); if (r) { s_status = r; return 1; } s_qstate = *q; s_qvalid = true; }
    s_queue.push_back(o);
    s_status = 0;
    if (s_queue.size() >= QUEUE_MAX) s_status = flush_queue();
    return 1;
}

static void set_ctrls(QOp& o, const int* ctrls, const int* cs, int nc) {
    o.ctrlMask = o.ctrlVals = 0;
    for (int i = 0; i < nc; i++) { o.ctrlMask |= 1ULL << ctrls[i]; if (!cs || cs[i]) o.ctrlVals |= 1ULL << ctrls[i]; }
}

static QOp blank(int kind, const qb_state* q, int nc) {
    QOp o; o.kind = kind; o.ctrlMask = o.ctrlVals = 0; o.t0 = o.t1 = 0; o.numT = 1; o.maskA = o.maskB = 0;
    for (int i = 0; i < 16; i++) o.m[i] = mk(0, 0);
    o.algBytes = (2 * (qindex)sizeof(cplx) * q->numAmpsPerNode) >> nc;
    return o;
}

int qb_tile_status() { return s_status; }

int qb_tile_try_dense(const qb_state* q, const int* ctrls, const int* cs, int nc, const int* targs, int nt, const qb_cplx* m) {
    if (!can_fuse(q) || nt > 2) return 0;
    QOp o = blank(nt == 1 ? OP_DENSE1 : OP_DENSE2, q, nc);
    set_ctrls(o, ctrls, cs, nc);
    o.t0 = targs[0]; o.t1 = nt > 1 ? targs[1] : 0; o.numT = nt;
    for (int i = 0; i < (nt == 1 ? 4 : 16); i++) o.m[i] = mk(m[i]);
    return enqueue(q, o);
}

int qb_tile_try_denseK(const qb_state*, const int*, const int*, int, const int*, int, const qb_cplx*, int) { return 0; }

int qb_tile_try_diag(const qb_state* q, const int* ctrls, const int* cs, int nc, const int* targs, int nt, const qb_cplx* e) {
    if (!can_fuse(q) || nt > 2) return 0;
    QOp o = blank(OP_DIAG, q, nc);
    set_ctrls(o, ctrls, cs, nc);
    o.numT = nt;
    // prefix targets (>= logN) are bits of the rank: fold them into the element table now
    const int n = q->logNumAmpsPerNode;
    int loc[2], nl = 0; cplx tab[4];
    for (int i = 0; i < (1 << nt); i++) tab[i] = mk(e[i]);
    if (nt == 1) {
        if (targs[0] >= n) { int b = (q->rank >> (targs[0] - n)) & 1; o.m[0] = o.m[1] = tab[b]; o.t0 = 0; o.numT = 1; }
        else { o.m[0] = tab[0]; o.m[1] = tab[1]; o.t0 = targs[0]; }
        return enqueue(q, o);
    }
    // two targets: element index = bit(t1) << 1 | bit(t0)  (getTwoBits(i, targ2, targ1), cpu_subroutines.cpp:643)
    int fixedBit[2] = {-1, -1};
    for (int i = 0; i < 2; i++) { if (targs[i] >= n) fixedBit[i] = (q->rank >> (targs[i] - n)) & 1; else loc[nl++] = i; }
    if (nl == 2) { o.t0 = targs[0]; o.t1 = targs[1]; for (int i = 0; i < 4; i++) o.m[i] = tab[i]; }
    else if (nl == 1) {
        int i = loc[0], other = 1 - i;
        o.numT = 1; o.t0 = targs[i];
        for (int b = 0; b < 2; b++) { int k = (i == 0) ? (b | (fixedBit[other] << 1)) : ((b << 1) | fixedBit[other]); o.m[b] = tab[k]; }
    } else { o.numT = 1; o.t0 = 0; o.m[0] = o.m[1] = tab[fixedBit[0] | (fixedBit[1] << 1)]; }
    return enqueue(q, o);
}

int qb_tile_try_pauli(const qb_state* q, const int* ctrls, const int* cs, int nc, unsigned long long maskXY, unsigned long long maskYZ, cplx ampFac, cplx pairFac) {
    if (!can_fuse(q)) return 0;
    QOp o = blank(OP_PAULI, q, nc);
    set_ctrls(o, ctrls, cs, nc);
    o.maskA = maskXY; o.maskB = maskYZ; o.m[0] = ampFac; o.m[1] = pairFac;
    return enqueue(q, o);
}

int qb_tile_try_phase(const qb_state* q, const int* ctrls, const int* cs, int nc, unsigned long long targMask, cplx f0, cplx f1) {
    if (!can_fuse(q)) return 0;
    QOp o = blank(OP_PARITY, q, nc);
    set_ctrls(o, ctrls, cs, nc);
    o.maskA = targMask; o.m[0] = f0; o.m[1] = f1;
    return enqueue(q, o);
}

int qb_tile_try_swap(const qb_state* q, const int* ctrls, const int* cs, int nc, int t1, int t2) {
    if (!can_fuse(q)) return 0;
    QOp o = blank(OP_SWAP, q, nc + 1);
    set_ctrls(o, ctrls, cs, nc);
    o.t0 = t1; o.t1 = t2; o.numT = 2;
    return enqueue(q, o);
}

// ------------------------------------------------------------------------------------------
// direct execution of a queued op (bypasses the queue: s_inFlush is set while this runs)
// ------------------------------------------------------------------------------------------
static int run_direct(const qb_state* q, const QOp& o) {
    int ctrls[64], cs[64], nc = 0;
    for (int b = 0; b < 64; b++) if ((o.ctrlMask >> b) & 1) { ctrls[nc] = b; cs[nc++] = (int)((o.ctrlVals >> b) & 1); }
    qb_cplx m[16];
    for (int i = 0; i < 16; i++) { m[i].re = o.m[i].x; m[i].im = o.m[i].y; }
    switch (o.kind) {
    case OP_DENSE1: return qb_statevec_anyCtrlOneTargDenseMatr_subA(q, ctrls, cs, nc, o.t0, m);
    case OP_DENSE2: return qb_statevec_anyCtrlTwoTargDenseMatr_sub(q, ctrls, cs, nc, o.t0, o.t1, m);
    case OP_SWAP: return qb_statevec_anyCtrlSwap_subA(q, ctrls, cs, nc, o.t0, o.t1);
    case OP_DIAG:
        if (o.numT == 1) return qb_statevec_anyCtrlOneTargDiagMatr_sub(q, ctrls, cs, nc, o.t0, m);
        return qb_statevec_anyCtrlTwoTargDiagMatr_sub(q, ctrls, cs, nc, o.t0, o.t1, m);
    case OP_PARITY: {
        int t[64], nt = 0;
        for (int b = 0; b < 64; b++) if ((o.maskA >> b) & 1) t[nt++] = b;
        return qb_statevector_anyCtrlAnyTargZOrPhaseGadget_sub(q, ctrls, cs, nc, t, nt, m[0], m[1]);
    }
    case OP_PAULI:
        return qb_pauli_raw(q, ctrls, cs, nc, o.maskA, o.maskB, m[0], m[1]);
    case OP_STAR: {
        // a star that did not end up in a tile pass: apply its controlled phases one by one
        for (auto& ce : o.star) {
            int c = ce.first; int one = 1;
            qb_cplx e[2] = {{1, 0}, {cos(ce.second), sin(ce.second)}};
            int r = qb_statevec_anyCtrlOneTargDiagMatr_sub(q, &c, &one, 1, o.t0, e);
            if (r) return r;
        }
        return 0;
    }
    }
    return qb_set_error(-1, "tile engine: unknown op", __FILE__, __LINE__);
}
