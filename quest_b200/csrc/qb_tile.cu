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
// Inside a pass, gates are further grouped into ROUNDS: a round owns 4 of the 12 tile bits; every thread pulls the
// 16 amplitudes spanned by those bits into registers, applies every gate of the round whose non-diagonal targets are
// among the 4 bits (diagonal gates always qualify) and writes the 16 amplitudes back: one shared-memory round trip
// and one barrier per ROUND instead of per gate.
//
// KERNEL.  One persistent CTA per SM (grid = #SMs): two compute warpgroups (128 threads each, 232 registers per thread
// after setmaxnreg; each owns one tile at a time and walks it in two 16-amplitude register blocks per thread and round)
// + 1 copy warp (its warpgroup gives its registers away),
// 3 stages x 64 KiB of shared memory.  The copy warp drives the copy engine: cp.async.bulk (TMA, 1-D) loads the chunks
// of a tile into a free stage, signalling the stage's `full` mbarrier with complete_tx; the compute warps wait on it,
// run the pass's rounds on the tile, fence.proxy.async and arrive on the stage's `done` mbarrier; the copy warp then
// writes the stage back with cp.async.bulk.global.shared::cta (bulk async-groups) and refills it with the tile three
// ahead.  So while one tile is computed, another is landing and a third is draining: HBM never waits for the SM.
// Algorithmic bytes per pass: every gate of the pass counts its own 2*16*N/2^c bytes, physical bytes
// are 2*16*N once per pass (or less: controls shared by every gate of a pass prune whole tiles before they are loaded).
#include "qb_common.cuh"
#include "qb_kernels.cuh"
#include "qb_tile.cuh"
#include "qb_pauli_group.cuh"
#ifdef QB_SELFTEST
#include "../../include/quest_b200_selftest.h"
#endif
#include <math.h>
#include <stddef.h>
#include <string.h>
#include <vector>
#include <algorithm>

#define TILE_BITS 12
#define TILE_LOW 6
#define TILE_AMPS (1 << TILE_BITS)
#define TILE_STAGES 3
#define TILE_THREADS 256              // compute threads: two independent warpgroups of WG_THREADS, each working on its own tile
#define WG_THREADS 128
#define WG_ITERS (TILE_AMPS / (WG_THREADS * RAMPS))   // register blocks per thread and round (2)
#define TILE_BLOCK (TILE_THREADS + 128)  // + one warpgroup whose first warp is the copy warp (register re-allocation is per warpgroup)
#define COMPUTE_REGS 240              // setmaxnreg: 384 x 168 at launch -> 256 x 240 (compute) + 128 x 24 (copy warpgroup)
#define COPY_REGS 24
#define RB 4                          // tile bits held in registers per round
#define RAMPS (1 << RB)               // amplitudes per register block
#define TILE_MAX_CHUNKS (1 << (TILE_BITS - TILE_LOW))
#define MAX_OPS_PER_PASS 48
#define QUEUE_MAX 2048                 // deferred gates before a forced flush (cfg 2 issues 680 per step)
#define PAULI_STREAM_FLUSH 48           // see enqueue(): early flush of pure wide-Pauli streams (a multiple of PG_K_GADGET)
#define FUSE_MIN_LOG_AMPS 13          // smaller states use the direct kernels immediately
#define STAR_SEGS 6                   // external-control tables: 6 segments x 6 bits of the global index

enum { OP_DENSE1 = 0, OP_DENSE2, OP_PAULI, OP_SWAP, OP_DIAG, OP_PARITY, OP_STAR, OP_HSTAR };   // OP_HSTAR: Hadamard on t0 fused with the phase star centred on t0 (one QFT stage)
enum { ROUND_REG = 0, ROUND_SMEM = 1 };

// ------------------------------------------------------------------------------------------
// queued gate, in global (local-shard) qubit coordinates
// ------------------------------------------------------------------------------------------
struct QOp {
    int kind;
    unsigned long long ctrlMask, ctrlVals;     // suffix controls
    int t0, t1, numT;                          // targets (prefix diagonal targets are resolved at enqueue)
    unsigned long long maskA, maskB;           // pauli: XY, YZ; parity gadget: target mask
    cplx m[16];                                // matrix / diagonal / factors
    std::vector<std::pair<int,double>> star;   // OP_STAR: (other qubit, theta); centre qubit = t0
    qindex algBytes;
};

// One deferred queue PER STATE (keyed by the amplitude pointer): programs that alternate gates between several Quregs
// -- the reference's own psi / rho co-evolution test, tests/integration/densitymatrix.cpp:68-163 -- keep fusing on each
// of them; with a single global queue every switch would flush and the engine would degrade to one gate per pass.
// Queues of different states are independent (disjoint memory; a queue whose range overlaps a newcomer's is flushed
// first), so "flush" may run them in any order: it runs them in creation order.
struct StateQueue { qb_state st; std::vector<QOp> ops; };
#define MAX_STATE_QUEUES 8
static std::vector<StateQueue> s_queues;
static int s_status = 0;
static bool s_inFlush = false;
static unsigned long long s_flushEpoch = 0;       // bumped whenever a non-empty queue is executed
// cumulative execution statistics (qb_tile_stats): what the planner made of the gates it was given
static unsigned long long s_statPasses = 0, s_statRounds = 0, s_statTileOps = 0, s_statDirectOps = 0, s_statQueuedGates = 0;
static double s_statFmaAmps = 0;                   // FP64 fused multiply-adds issued per pass, summed: ops x amplitudes x FMA per amplitude
static double s_statPassBytes = 0;                 // bytes the launched passes stream through HBM (read + write of what they touch)
// Restricted flush (qb_tile_flush_restricted): the queue is applied only to the amplitudes whose index bits `mask`
// hold `vals` -- as if every queued gate carried those extra controls, but without disturbing gate absorption or the
// phase-star merge (the restriction is a property of the passes: whole tiles are pruned).  Used to overlap a
// half-shard exchange with the gates that commute with it (qb_p2p.cu).
static unsigned long long s_restrictMask = 0, s_restrictVals = 0;

// device-side op, in tile coordinates
struct TileOp {
    // quad 0: everything the round's dispatcher needs, fetched with one 128-bit shared-memory load (and prefetched
    // one op ahead): `code` selects the fully specialised gate body (see op_code)
    int code, kind;
    unsigned int inCtrlMask, inCtrlVals;
    int p0, p1, numT, tab;                     // in-tile positions (p < 0: target outside the tile, diag/star only); OP_STAR: table block
    int e0, e1;                                // global bit index of an external diag target / star centre
    int l0, l1;                                // register rounds: index (0..3) of the target bits among the round's bits
    unsigned int lmaskA, lmaskB;               // register rounds: pauli XY mask / YZ-or-parity mask over the round's bits
    unsigned int inMaskA, inMaskB;             // pauli XY / YZ (or parity) masks over the 12 tile bits
    unsigned long long extCtrlMask, extCtrlVals;
    unsigned long long extMaskB, pad;          // pauli / parity: sign bits outside the tile
    cplx m[16];
};
static_assert(sizeof(TileOp) % 16 == 0 && offsetof(TileOp, m) % 16 == 0, "TileOp is read with 128-bit loads");

// dispatch codes: one per specialised body
enum { CODE_DENSE1 = 0,      // + 2 * l0 + hasInTileCtrl                (8)
       CODE_DENSE2 = 8,      // + 2 * pairIndex(l0, l1) + hasInTileCtrl (12)
       CODE_SWAP = 20,       // + pairIndex                             (6)
       CODE_PAULI = 26,      // + lmaskA - 1                            (15)
       CODE_DIAG = 41, CODE_PARITY = 42, CODE_STAR = 43,     // + l0 (centre among the round's bits), + 4 (centre elsewhere)
       CODE_HSTAR = 48 };                                     // + l0 (the centre is a non-diagonal target: always a round bit)
static inline int pair_index(int a, int b) { return a == 0 ? b - 1 : (a == 1 ? b + 1 : 5); }   // (0,1)(0,2)(0,3)(1,2)(1,3)(2,3) -> 0..5

struct RoundHdr { int kind, opBase, numOps, pad; int b[RB]; };
// phase-star tables: `in` = three 16-entry tables over the 12 tile bits (4 bits each), looked up per thread and block;
// `ext` = six 64-entry tables over the bits outside the tile, folded into one factor per tile.  The `in` tables of the
// first STAR_SMEM_SLOTS stars of a pass are copied into shared memory by the kernel (768 bytes each): fetched from
// global they made the QFT passes wait on L1/L2 (ncu: long_scoreboard 3.5 per issued instruction, profiles/r2_ncu_summary.md)
#define STAR_SMEM_SLOTS 8
#define STAR_IN_ENTRIES 48
struct StarTab { cplx in[3][16]; cplx ext[STAR_SEGS][64]; };

struct PassHdr {
    int numOps, numRounds, numChunks, chunkAmps;
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

// optional cycle accounting of the compute warpgroups (probe builds only: -DQB_TILE_TIMING, tools/tile_timing_probe.py)
#ifdef QB_TILE_TIMING
#define TT_SLOTS 12
__device__ unsigned long long g_tileTiming[TT_SLOTS];
#endif
#if defined(QB_TILE_TIMING) && defined(__CUDA_ARCH__)
__device__ __forceinline__ unsigned long long tt_now() { unsigned long long t; asm volatile("mov.u64 %0, %clock64;" : "=l"(t) :: "memory"); return t; }
#define TT_DECL unsigned long long tt_prev = tt_now(), tt_acc[TT_SLOTS] = {0}
#define TT_MARK(slot) do { unsigned long long n_ = tt_now(); tt_acc[slot] += n_ - tt_prev; tt_prev = n_; } while (0)
#define TT_ARGS , unsigned long long& tt_prev, unsigned long long (&tt_acc)[TT_SLOTS]
#define TT_PASS , tt_prev, tt_acc
#define TT_FLUSH do { if ((threadIdx.x & 127) == 0) for (int i_ = 0; i_ < TT_SLOTS; i_++) atomicAdd(&g_tileTiming[i_], tt_acc[i_]); } while (0)
#else
#define TT_DECL
#define TT_MARK(slot)
#define TT_ARGS
#define TT_PASS
#define TT_FLUSH
#endif

// the gate bodies and the round driver below are compiled for the device (k_tile_pass) AND for the host, where the
// CPU self-test qb_selftest_tile_emulation runs the very same code on the descriptors emit_pass produced
#ifdef __CUDA_ARCH__
#define QB_LDG(p) __ldg(p)
#define QB_POPC(x) __popc(x)
#define QB_CLZ(x) __clz(x)
#else
#define QB_LDG(p) (*(p))
#define QB_POPC(x) __builtin_popcount(x)
#define QB_CLZ(x) __builtin_clz(x)
#endif
#define QB_HD __host__ __device__ __forceinline__

QB_HD unsigned ins0(unsigned v, int p) { return ((v >> p) << (p + 1)) | (v & ((1u << p) - 1u)); }
QB_HD cplx csel(bool c, cplx a, cplx b) { return mk(c ? a.x : b.x, c ? a.y : b.y); }

// ------------------------------------------------------------------------------------------
// register-round gate bodies: v[u] is the amplitude whose round-bit pattern is u (bit k of u <-> round bit k);
// `ok` has bit u set when amplitude u satisfies the gate's in-tile controls
// ------------------------------------------------------------------------------------------
// CTRL = false: the gate has no in-tile controls, every amplitude is updated and no selects are emitted
template <int K, bool CTRL>
QB_HD void reg_dense1(cplx (&v)[RAMPS], cplx m00, cplx m01, cplx m10, cplx m11, unsigned ok) {
#pragma unroll
    for (int u = 0; u < RAMPS; u++) {
        if (u & (1 << K)) continue;
        const int u1 = u | (1 << K);
        cplx a0 = v[u], a1 = v[u1];
        cplx n0 = cfma(m01, a1, cmul(m00, a0)), n1 = cfma(m11, a1, cmul(m10, a0));
        const bool c = CTRL ? (bool)((ok >> u) & 1) : true;
        v[u] = csel(c, n0, a0); v[u1] = csel(c, n1, a1);
    }
}

// K0 < K1; matrix index bit 0 <-> K0, bit 1 <-> K1 (the host re-orders the matrix to make it so); the first matrix row
// arrives in registers (prefetched while the previous gate ran), the other three are loaded while it is being used.
// QB_D2Q quadruples are updated together, column by column of the matrix: 8 x QB_D2Q independent FMA chains in flight
// (the FP64 pipe's dependent-issue latency is ~30 cycles at 2 cycles per warp instruction: 8 chains leave it half idle)
#ifndef QB_D2Q
#define QB_D2Q 2
#endif
template <int K0, int K1, bool CTRL>
QB_HD void reg_dense2(cplx (&v)[RAMPS], cplx r0, cplx r1, cplx r2, cplx r3, const cplx* __restrict__ mp, unsigned ok) {
    constexpr int B0 = 1 << K0, B1 = 1 << K1;
    // the two register bits that are not targets enumerate the four quadruples
    constexpr int O0 = (K0 != 0 && K1 != 0) ? 0 : ((K0 != 1 && K1 != 1) ? 1 : 2);
    constexpr int O1 = (K0 != 3 && K1 != 3) ? 3 : ((K0 != 2 && K1 != 2) ? 2 : 1);
    cplx m[16];
    m[0] = r0; m[1] = r1; m[2] = r2; m[3] = r3;
#pragma unroll
    for (int i = 4; i < 16; i++) m[i] = mp[i];
#pragma unroll
    for (int g = 0; g < 4; g += QB_D2Q) {
        cplx a[QB_D2Q][4], n[QB_D2Q][4];
#pragma unroll
        for (int q = 0; q < QB_D2Q; q++) {
            const int u = (((g + q) & 1) << O0) | (((g + q) >> 1) << O1);
            a[q][0] = v[u]; a[q][1] = v[u | B0]; a[q][2] = v[u | B1]; a[q][3] = v[u | B0 | B1];
        }
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int q = 0; q < QB_D2Q; q++) n[q][r] = cmul(m[4 * r], a[q][0]);
#pragma unroll
        for (int c = 1; c < 4; c++)
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int q = 0; q < QB_D2Q; q++) n[q][r] = cfma(m[4 * r + c], a[q][c], n[q][r]);
#pragma unroll
        for (int q = 0; q < QB_D2Q; q++) {
            const int u = (((g + q) & 1) << O0) | (((g + q) >> 1) << O1);
            const bool cnd = CTRL ? (bool)((ok >> u) & 1) : true;
            v[u] = csel(cnd, n[q][0], a[q][0]); v[u | B0] = csel(cnd, n[q][1], a[q][1]);
            v[u | B1] = csel(cnd, n[q][2], a[q][2]); v[u | B0 | B1] = csel(cnd, n[q][3], a[q][3]);
        }
    }
}

template <int K0, int K1>
QB_HD void reg_swap(cplx (&v)[RAMPS], unsigned ok) {
#pragma unroll
    for (int u = 0; u < RAMPS; u++) {
        if (!(u & (1 << K0)) || (u & (1 << K1))) continue;       // u has K0 = 1, K1 = 0
        const int w = u ^ ((1 << K0) | (1 << K1));
        const bool c = (ok >> u) & 1;
        cplx a = v[u], b = v[w];
        v[u] = csel(c, b, a); v[w] = csel(c, a, b);
    }
}

// pairs (u, u ^ LXY); sign of amplitude u = (-1)^(parity of its Y/Z bits) = basePar ^ popc(u & lyz)
template <int LXY>
QB_HD void reg_pauli(cplx (&v)[RAMPS], unsigned lyz, int basePar, cplx af, cplx pf, unsigned ok) {
    constexpr int H = (LXY & 8) ? 3 : (LXY & 4) ? 2 : (LXY & 2) ? 1 : 0;
#pragma unroll
    for (int u = 0; u < RAMPS; u++) {
        if (u & (1 << H)) continue;
        const int w = u ^ LXY;
        const double sA = 1.0 - 2.0 * ((QB_POPC(u & lyz) + basePar) & 1);
        const double sB = 1.0 - 2.0 * ((QB_POPC(w & lyz) + basePar) & 1);
        cplx a = v[u], b = v[w];
        cplx nA = cfma(pf, cscale(sB, b), cmul(af, a)), nB = cfma(pf, cscale(sA, a), cmul(af, b));
        const bool c = (ok >> u) & 1;
        v[u] = csel(c, nA, a); v[w] = csel(c, nB, b);
    }
}

// phase star on the amplitudes selected by `on` (bit u): v[u] *= eb * m[u]
QB_HD void reg_star(cplx (&v)[RAMPS], cplx eb, const cplx* __restrict__ mu, unsigned on) {
#pragma unroll
    for (int u = 0; u < RAMPS; u++) {
        const cplx f = cmul(eb, mu[u]);
        v[u] = csel((on >> u) & 1, cmul(v[u], f), v[u]);
    }
}
template <int L>      // centre = round bit L: only the 8 amplitudes with that bit set are touched (no selects)
QB_HD void reg_star_bit(cplx (&v)[RAMPS], cplx eb, const cplx* __restrict__ mu) {
#pragma unroll
    for (int u = 0; u < RAMPS; u++) {
        if (!(u & (1 << L))) continue;
        v[u] = cmul(v[u], cmul(eb, mu[u]));
    }
}

// one QFT stage on round bit L: Hadamard, then the phase star on the amplitudes whose bit L is set:
//   v0' = (v0 + v1)/sqrt2,  v1' = (v0 - v1) * (eb * m[u1]) with eb already carrying the 1/sqrt2
// 7 FP64 instructions per amplitude instead of 12 for the generic dense gate followed by the star
template <int L>
QB_HD void reg_hstar_bit(cplx (&v)[RAMPS], cplx ebs, const cplx* __restrict__ mu) {
    const double s = 0.70710678118654752440;
#pragma unroll
    for (int u = 0; u < RAMPS; u++) {
        if (u & (1 << L)) continue;
        const int u1 = u | (1 << L);
        const cplx a0 = v[u], a1 = v[u1];
        const cplx f = cmul(ebs, mu[u1]);
        v[u] = mk(s * (a0.x + a1.x), s * (a0.y + a1.y));
        v[u1] = cmul(mk(a0.x - a1.x, a0.y - a1.y), f);
    }
}

// ------------------------------------------------------------------------------------------
// one register round on one tile.  The op loop is software-pipelined: while gate o runs on the FP64 pipe, the
// dispatch quad and the first four matrix entries (or the phase-star table entries) of gate o+1 are already in flight,
// so the shared-memory latency of the descriptors never sits between two gates.
// `active` (bit i <-> op i of the pass) holds the tile-uniform tests: external controls, external star centres.
// ------------------------------------------------------------------------------------------
QB_HD void reg_round(cplx* __restrict__ t, const RoundHdr& rd, const TileOp* __restrict__ ops, qindex base,
                                          unsigned long long active0, const StarTab* __restrict__ tabs, const cplx* __restrict__ starF, int wtid TT_ARGS) {
  for (int it = 0; it < WG_ITERS; it++) {
    const int tid = wtid + it * WG_THREADS;
    unsigned long long active = active0;
    const int b0 = rd.b[0], b1 = rd.b[1], b2 = rd.b[2], b3 = rd.b[3];          // ascending tile-bit positions
    const unsigned jb = ins0(ins0(ins0(ins0((unsigned)tid, b0), b1), b2), b3);
    const unsigned o0 = 1u << b0, o1 = 1u << b1, o2 = 1u << b2, o3 = 1u << b3;
#define OFF(u) ((((u) & 1) ? o0 : 0u) | (((u) & 2) ? o1 : 0u) | (((u) & 4) ? o2 : 0u) | (((u) & 8) ? o3 : 0u))
    const int first = rd.opBase, num = rd.numOps;
    const TileOp* op = ops + first;
    active >>= first;

    // Two-deep software pipeline of the descriptors: while gate o runs, the four complex operands of gate o+1 are
    // fetched -- which ones depends on its dispatch quad, loaded one gate EARLIER and long since arrived, so no
    // load sits between two gates -- together with the dispatch quad of gate o+2.
    int4 d, d1; cplx pa, pb, pc, pd;
#define LOAD_OPERANDS(q, D, A, B, C_, D_) do { \
        if (D.x >= CODE_STAR) { const StarTab& tb_ = tabs[(q)->tab]; A = QB_LDG(&tb_.in[0][jb & 15]); B = cmul(QB_LDG(&tb_.in[1][(jb >> 4) & 15]), QB_LDG(&tb_.in[2][jb >> 8])); C_ = starF[(q) - ops]; D_ = C_; } \
        else { A = (q)->m[0]; B = (q)->m[1]; C_ = (q)->m[2]; D_ = (q)->m[3]; } } while (0)
    d = *reinterpret_cast<const int4*>(op);
    d1 = d;
    if (num > 1) d1 = *reinterpret_cast<const int4*>(op + 1);
    LOAD_OPERANDS(op, d, pa, pb, pc, pd);

    cplx v[RAMPS];
#pragma unroll
    for (int u = 0; u < RAMPS; u++) v[u] = t[jb | OFF(u)];
    TT_MARK(4);      // address arithmetic + issue of the 16 shared-memory loads

    for (int o = 0; o < num; o++, op++, active >>= 1) {
        int4 d2 = d1; cplx na = pa, nb = pb, nc_ = pc, nd = pd;
        if (o + 1 < num) LOAD_OPERANDS(op + 1, d1, na, nb, nc_, nd);
        if (o + 2 < num) d2 = *reinterpret_cast<const int4*>(op + 2);
        TT_MARK(10);     // per-gate: issue of the prefetches
        if (active & 1) {
            unsigned ok = 0xFFFFu;
            const unsigned cm = (unsigned)d.z, cv = (unsigned)d.w;
            if (cm) {
                ok = 0;
#pragma unroll
                for (int u = 0; u < RAMPS; u++) ok |= (unsigned)(((jb | OFF(u)) & cm) == cv) << u;
            }
            switch (d.x) {
#define D1(L) case CODE_DENSE1 + 2 * L: reg_dense1<L, false>(v, pa, pb, pc, pd, ok); break; \
              case CODE_DENSE1 + 2 * L + 1: reg_dense1<L, true>(v, pa, pb, pc, pd, ok); break;
            D1(0) D1(1) D1(2) D1(3)
#undef D1
#define D2(I, A, B) case CODE_DENSE2 + 2 * I: reg_dense2<A, B, false>(v, pa, pb, pc, pd, op->m, ok); break; \
                    case CODE_DENSE2 + 2 * I + 1: reg_dense2<A, B, true>(v, pa, pb, pc, pd, op->m, ok); break;
            D2(0, 0, 1) D2(1, 0, 2) D2(2, 0, 3) D2(3, 1, 2) D2(4, 1, 3) D2(5, 2, 3)
#undef D2
            case CODE_SWAP + 0: reg_swap<0, 1>(v, ok); break;
            case CODE_SWAP + 1: reg_swap<0, 2>(v, ok); break;
            case CODE_SWAP + 2: reg_swap<0, 3>(v, ok); break;
            case CODE_SWAP + 3: reg_swap<1, 2>(v, ok); break;
            case CODE_SWAP + 4: reg_swap<1, 3>(v, ok); break;
            case CODE_SWAP + 5: reg_swap<2, 3>(v, ok); break;
#define PC(X) case CODE_PAULI + X - 1: reg_pauli<X>(v, op->lmaskB, (QB_POPC(jb & op->inMaskB) + parity64((unsigned long long)base & op->extMaskB)) & 1, pa, pb, ok); break;
            // sign bits of the Y/Z mask: those among the round's bits vary with u (lmaskB), the rest is fixed for this thread
            PC(1) PC(2) PC(3) PC(4) PC(5) PC(6) PC(7) PC(8) PC(9) PC(10) PC(11) PC(12) PC(13) PC(14) PC(15)
#undef PC
            case CODE_DIAG: {
                const int p0 = op->p0, p1 = op->p1, numT = op->numT;
                const int x0 = (p0 < 0) ? getBit(base, op->e0) : 0;
                const int x1 = (numT > 1 && p1 < 0) ? getBit(base, op->e1) : 0;
#pragma unroll
                for (int u = 0; u < RAMPS; u++) {
                    const unsigned j = jb | OFF(u);
                    const int k0 = (p0 < 0) ? x0 : ((j >> p0) & 1);
                    const int k1 = (numT > 1) ? ((p1 < 0) ? x1 : ((j >> p1) & 1)) : 0;
                    const cplx f = k1 ? (k0 ? pd : pc) : (k0 ? pb : pa);
                    v[u] = csel((ok >> u) & 1, cmul(v[u], f), v[u]);
                }
            } break;
            case CODE_PARITY: {
                const unsigned inA = op->inMaskA;
                const int extPar = parity64((unsigned long long)base & op->extMaskB);
#pragma unroll
                for (int u = 0; u < RAMPS; u++) {
                    const int par = (QB_POPC((jb | OFF(u)) & inA) + extPar) & 1;
                    v[u] = csel((ok >> u) & 1, cmul(v[u], par ? pb : pa), v[u]);
                }
            } break;
            // phase star: amplitudes with the centre bit set gain exp(i sum_c theta_c bit_c).  The phase is additive over
            // index bits, so it factors into a per-thread part (tile bits outside the round: two 64-entry tables, times
            // the per-tile factor of the bits outside the tile -- all three prefetched) and a per-register part op->m[u]
            case CODE_STAR + 0: reg_star_bit<0>(v, cmul(cmul(pa, pb), pc), op->m); break;
            case CODE_STAR + 1: reg_star_bit<1>(v, cmul(cmul(pa, pb), pc), op->m); break;
            case CODE_STAR + 2: reg_star_bit<2>(v, cmul(cmul(pa, pb), pc), op->m); break;
            case CODE_STAR + 3: reg_star_bit<3>(v, cmul(cmul(pa, pb), pc), op->m); break;
            case CODE_HSTAR + 0: reg_hstar_bit<0>(v, cscale(0.70710678118654752440, cmul(cmul(pa, pb), pc)), op->m); break;
            case CODE_HSTAR + 1: reg_hstar_bit<1>(v, cscale(0.70710678118654752440, cmul(cmul(pa, pb), pc)), op->m); break;
            case CODE_HSTAR + 2: reg_hstar_bit<2>(v, cscale(0.70710678118654752440, cmul(cmul(pa, pb), pc)), op->m); break;
            case CODE_HSTAR + 3: reg_hstar_bit<3>(v, cscale(0.70710678118654752440, cmul(cmul(pa, pb), pc)), op->m); break;
            default:            // CODE_STAR + 4: centre outside the round (a tile bit tested through `ok`, or external and set)
                if (ok) reg_star(v, cmul(cmul(pa, pb), pc), op->m, ok);
                break;
            }
        }
        TT_MARK(11);     // per-gate: control mask + dispatch + body
        d = d1; d1 = d2; pa = na; pb = nb; pc = nc_; pd = nd;
    }
    TT_MARK(5);      // loop exit
#pragma unroll
    for (int u = 0; u < RAMPS; u++) t[jb | OFF(u)] = v[u];
    TT_MARK(6);      // 16 shared-memory stores
  }
#undef LOAD_OPERANDS
#undef OFF
}

// ------------------------------------------------------------------------------------------
// The same round with BOTH register blocks of a thread resident (2 x 16 amplitudes): every gate is decoded and
// dispatched once per thread and round instead of once per block.  Measured (profiles/r2_round_probe_*.txt): decode +
// indirect branch + operand latency cost ~500 cycles per dispatch against 256 (1-qubit) .. 1024 (2-qubit) cycles of FP64
// issue per block, so halving the dispatches per amplitude is worth more than the extra registers cost.
// ------------------------------------------------------------------------------------------
QB_HD void reg_diag_block(cplx (&v)[RAMPS], unsigned jb, unsigned o0, unsigned o1, unsigned o2, unsigned o3, const TileOp* op, qindex base,
                          cplx pa, cplx pb, cplx pc, cplx pd, unsigned ok) {
#define OFF(u) ((((u) & 1) ? o0 : 0u) | (((u) & 2) ? o1 : 0u) | (((u) & 4) ? o2 : 0u) | (((u) & 8) ? o3 : 0u))
    const int p0 = op->p0, p1 = op->p1, numT = op->numT;
    const int x0 = (p0 < 0) ? getBit(base, op->e0) : 0;
    const int x1 = (numT > 1 && p1 < 0) ? getBit(base, op->e1) : 0;
#pragma unroll
    for (int u = 0; u < RAMPS; u++) {
        const unsigned j = jb | OFF(u);
        const int k0 = (p0 < 0) ? x0 : ((j >> p0) & 1);
        const int k1 = (numT > 1) ? ((p1 < 0) ? x1 : ((j >> p1) & 1)) : 0;
        const cplx f = k1 ? (k0 ? pd : pc) : (k0 ? pb : pa);
        v[u] = csel((ok >> u) & 1, cmul(v[u], f), v[u]);
    }
}
QB_HD void reg_parity_block(cplx (&v)[RAMPS], unsigned jb, unsigned o0, unsigned o1, unsigned o2, unsigned o3, const TileOp* op, qindex base,
                            cplx pa, cplx pb, unsigned ok) {
    const unsigned inA = op->inMaskA;
    const int extPar = parity64((unsigned long long)base & op->extMaskB);
#pragma unroll
    for (int u = 0; u < RAMPS; u++) {
        const int par = (QB_POPC((jb | OFF(u)) & inA) + extPar) & 1;
        v[u] = csel((ok >> u) & 1, cmul(v[u], par ? pb : pa), v[u]);
    }
}
QB_HD unsigned ctrl_mask_block(unsigned jb, unsigned o0, unsigned o1, unsigned o2, unsigned o3, unsigned cm, unsigned cv) {
    unsigned ok = 0;
#pragma unroll
    for (int u = 0; u < RAMPS; u++) ok |= (unsigned)(((jb | OFF(u)) & cm) == cv) << u;
    return ok;
#undef OFF
}

QB_HD void reg_round2(cplx* __restrict__ t, const RoundHdr& rd, const TileOp* __restrict__ ops, qindex base,
                      unsigned long long active, const StarTab* __restrict__ tabs, const cplx* __restrict__ starF, int wtid,
                      const cplx* __restrict__ starIn TT_ARGS) {
    static_assert(WG_ITERS == 2, "the dual-block round driver holds exactly two register blocks per thread");
    const int b0 = rd.b[0], b1 = rd.b[1], b2 = rd.b[2], b3 = rd.b[3];          // ascending tile-bit positions
    const unsigned jbA = ins0(ins0(ins0(ins0((unsigned)wtid, b0), b1), b2), b3);
    const unsigned jbB = ins0(ins0(ins0(ins0((unsigned)(wtid + WG_THREADS), b0), b1), b2), b3);
    const unsigned o0 = 1u << b0, o1 = 1u << b1, o2 = 1u << b2, o3 = 1u << b3;
#define OFF(u) ((((u) & 1) ? o0 : 0u) | (((u) & 2) ? o1 : 0u) | (((u) & 4) ? o2 : 0u) | (((u) & 8) ? o3 : 0u))
    const int first = rd.opBase, num = rd.numOps;
    const TileOp* op = ops + first;
    active >>= first;

    // the dispatch quads are fetched two gates ahead; the operands of a gate are fetched when it is dispatched (their
    // latency is paid once per 32 amplitudes here, and holding a prefetched set for both blocks would cost ~40 registers)
    int4 d, d1;
#define LOAD_OPERANDS(q, D, A, B, C_, D_, QA, QB_) do { \
        if (D.x >= CODE_STAR) { const unsigned sl_ = (unsigned)(q)->pad; \
                                const cplx* ti_ = (starIn && sl_ < STAR_SMEM_SLOTS) ? starIn + sl_ * STAR_IN_ENTRIES : &tabs[(q)->tab].in[0][0]; \
                                A = ti_[jbA & 15]; B = cmul(ti_[16 + ((jbA >> 4) & 15)], ti_[32 + (jbA >> 8)]); \
                                QA = ti_[jbB & 15]; QB_ = cmul(ti_[16 + ((jbB >> 4) & 15)], ti_[32 + (jbB >> 8)]); C_ = starF[(q) - ops]; D_ = C_; } \
        else { A = (q)->m[0]; B = (q)->m[1]; C_ = (q)->m[2]; D_ = (q)->m[3]; QA = A; QB_ = B; } } while (0)
    d = *reinterpret_cast<const int4*>(op);
    d1 = d;
    if (num > 1) d1 = *reinterpret_cast<const int4*>(op + 1);

    cplx vA[RAMPS], vB[RAMPS];
#pragma unroll
    for (int u = 0; u < RAMPS; u++) vA[u] = t[jbA | OFF(u)];
#pragma unroll
    for (int u = 0; u < RAMPS; u++) vB[u] = t[jbB | OFF(u)];
    TT_MARK(4);

    for (int o = 0; o < num; o++, op++, active >>= 1) {
        int4 d2 = d1;
        if (o + 2 < num) d2 = *reinterpret_cast<const int4*>(op + 2);
        TT_MARK(10);
        if (active & 1) {
            cplx pa, pb, pc, pd, qa, qb;
            LOAD_OPERANDS(op, d, pa, pb, pc, pd, qa, qb);
            unsigned okA = 0xFFFFu, okB = 0xFFFFu;
            const unsigned cm = (unsigned)d.z, cv = (unsigned)d.w;
            if (cm) { okA = ctrl_mask_block(jbA, o0, o1, o2, o3, cm, cv); okB = ctrl_mask_block(jbB, o0, o1, o2, o3, cm, cv); }
            switch (d.x) {
#define D1(L) case CODE_DENSE1 + 2 * L: reg_dense1<L, false>(vA, pa, pb, pc, pd, okA); reg_dense1<L, false>(vB, pa, pb, pc, pd, okB); break; \
              case CODE_DENSE1 + 2 * L + 1: reg_dense1<L, true>(vA, pa, pb, pc, pd, okA); reg_dense1<L, true>(vB, pa, pb, pc, pd, okB); break;
            D1(0) D1(1) D1(2) D1(3)
#undef D1
#define D2(I, A, B) case CODE_DENSE2 + 2 * I: reg_dense2<A, B, false>(vA, pa, pb, pc, pd, op->m, okA); reg_dense2<A, B, false>(vB, pa, pb, pc, pd, op->m, okB); break; \
                    case CODE_DENSE2 + 2 * I + 1: reg_dense2<A, B, true>(vA, pa, pb, pc, pd, op->m, okA); reg_dense2<A, B, true>(vB, pa, pb, pc, pd, op->m, okB); break;
            D2(0, 0, 1) D2(1, 0, 2) D2(2, 0, 3) D2(3, 1, 2) D2(4, 1, 3) D2(5, 2, 3)
#undef D2
#define SW(I, A, B) case CODE_SWAP + I: reg_swap<A, B>(vA, okA); reg_swap<A, B>(vB, okB); break;
            SW(0, 0, 1) SW(1, 0, 2) SW(2, 0, 3) SW(3, 1, 2) SW(4, 1, 3) SW(5, 2, 3)
#undef SW
#define PC(X) case CODE_PAULI + X - 1: { const int ep_ = parity64((unsigned long long)base & op->extMaskB); const unsigned lm_ = op->lmaskB, im_ = op->inMaskB; \
                reg_pauli<X>(vA, lm_, (QB_POPC(jbA & im_) + ep_) & 1, pa, pb, okA); reg_pauli<X>(vB, lm_, (QB_POPC(jbB & im_) + ep_) & 1, pa, pb, okB); } break;
            PC(1) PC(2) PC(3) PC(4) PC(5) PC(6) PC(7) PC(8) PC(9) PC(10) PC(11) PC(12) PC(13) PC(14) PC(15)
#undef PC
            case CODE_DIAG:
                reg_diag_block(vA, jbA, o0, o1, o2, o3, op, base, pa, pb, pc, pd, okA);
                reg_diag_block(vB, jbB, o0, o1, o2, o3, op, base, pa, pb, pc, pd, okB);
                break;
            case CODE_PARITY:
                reg_parity_block(vA, jbA, o0, o1, o2, o3, op, base, pa, pb, okA);
                reg_parity_block(vB, jbB, o0, o1, o2, o3, op, base, pa, pb, okB);
                break;
#define ST(L) case CODE_STAR + L: reg_star_bit<L>(vA, cmul(cmul(pa, pb), pc), op->m); reg_star_bit<L>(vB, cmul(cmul(qa, qb), pc), op->m); break; \
              case CODE_HSTAR + L: reg_hstar_bit<L>(vA, cscale(0.70710678118654752440, cmul(cmul(pa, pb), pc)), op->m); \
                                   reg_hstar_bit<L>(vB, cscale(0.70710678118654752440, cmul(cmul(qa, qb), pc)), op->m); break;
            ST(0) ST(1) ST(2) ST(3)
#undef ST
            default:            // CODE_STAR + 4: centre outside the round (a tile bit tested through `ok`, or external and set)
                if (okA) reg_star(vA, cmul(cmul(pa, pb), pc), op->m, okA);
                if (okB) reg_star(vB, cmul(cmul(qa, qb), pc), op->m, okB);
                break;
            }
        }
        TT_MARK(11);
        d = d1; d1 = d2;
    }
    TT_MARK(5);
    // the 32 shared-memory addresses are RE-computed for the stores (an opaque copy of their six ingredients keeps the
    // compiler from holding 32 address registers live across the whole round, which is what pushed it into spilling)
    {
        unsigned sA = jbA, sB = jbB, p0 = o0, p1 = o1, p2 = o2, p3 = o3;
#ifdef __CUDA_ARCH__
        asm volatile("" : "+r"(sA), "+r"(sB), "+r"(p0), "+r"(p1), "+r"(p2), "+r"(p3));
#endif
#define SOFF(u) ((((u) & 1) ? p0 : 0u) | (((u) & 2) ? p1 : 0u) | (((u) & 4) ? p2 : 0u) | (((u) & 8) ? p3 : 0u))
#pragma unroll
        for (int u = 0; u < RAMPS; u++) t[sA | SOFF(u)] = vA[u];
#pragma unroll
        for (int u = 0; u < RAMPS; u++) t[sB | SOFF(u)] = vB[u];
#undef SOFF
    }
    TT_MARK(6);
#undef LOAD_OPERANDS
#undef OFF
}

#ifndef QB_DUAL
#define QB_DUAL 1
#endif
#if QB_DUAL
#define REG_ROUND reg_round2
#define REG_ROUND_EXTRA(p) , p
#else
#define REG_ROUND reg_round
#define REG_ROUND_EXTRA(p)
#endif

// shared-memory fallback for the one gate shape that cannot live in a 4-bit register round:
// Pauli strings with X/Y on more than four tile bits
QB_HD void smem_pauli(cplx* __restrict__ t, const TileOp& op, qindex base, int wtid) {
    const unsigned cm = op.inCtrlMask, cv = op.inCtrlVals;
    const unsigned xy = op.inMaskA, yz = op.inMaskB;
    const int h = 31 - QB_CLZ(xy);
    const int extPar = parity64((unsigned long long)base & op.extMaskB);
    const cplx af = op.m[0], pf = op.m[1];
    for (unsigned n = wtid; n < TILE_AMPS / 2; n += WG_THREADS) {
        unsigned jA = ins0(n, h);
        if ((jA & cm) != cv) continue;
        unsigned jB = jA ^ xy;
        double sA = 1.0 - 2.0 * ((QB_POPC(jA & yz) + extPar) & 1);
        double sB = 1.0 - 2.0 * ((QB_POPC(jB & yz) + extPar) & 1);
        cplx a = t[jA], b = t[jB];
        t[jA] = cfma(pf, cscale(sB, b), cmul(af, a));
        t[jB] = cfma(pf, cscale(sA, a), cmul(af, b));
    }
}

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void wg_sync(int wg) { asm volatile("bar.sync %0, %1;" :: "r"(1 + wg), "n"(WG_THREADS) : "memory"); }

__global__ void __launch_bounds__(TILE_BLOCK, 1) k_tile_pass(cplx* __restrict__ amps, const PassHdr* __restrict__ hdrp,
        const RoundHdr* __restrict__ grounds, const TileOp* __restrict__ gops, const StarTab* __restrict__ tabs) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cplx* stageBuf = reinterpret_cast<cplx*>(smem_raw);                                  // [STAGES][TILE_AMPS]
    TileOp* ops = reinterpret_cast<TileOp*>(smem_raw + (size_t)TILE_STAGES * TILE_AMPS * sizeof(cplx));
    __shared__ unsigned long long full[TILE_STAGES];     // copy engine -> compute warps: the tile has landed
    __shared__ unsigned long long done[TILE_STAGES];     // compute warps -> copy warp: the tile is computed
    __shared__ PassHdr hdr;
    __shared__ RoundHdr rounds[MAX_OPS_PER_PASS];
    __shared__ cplx starF[2][2][MAX_OPS_PER_PASS];  // per warpgroup, double-buffered: a fast warp may start its next tile while a slow one finishes
    __shared__ cplx starIn[STAR_SMEM_SLOTS][STAR_IN_ENTRIES];   // in-tile tables of the pass's first phase stars

    const int tid = threadIdx.x;
    for (int i = tid; i < (int)(sizeof(PassHdr) / 4); i += TILE_BLOCK) ((int*)&hdr)[i] = ((const int*)hdrp)[i];
    __syncthreads();
    const int numOps = hdr.numOps, numRounds = hdr.numRounds;
    for (int i = tid; i < numOps * (int)(sizeof(TileOp) / 4); i += TILE_BLOCK) ((int*)ops)[i] = ((const int*)gops)[i];
    for (int i = tid; i < numRounds * (int)(sizeof(RoundHdr) / 4); i += TILE_BLOCK) ((int*)rounds)[i] = ((const int*)grounds)[i];
    if (tid == 0) {
        for (int s = 0; s < TILE_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&done[s], WG_THREADS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    for (int o = 0; o < numOps; o++) {                 // (ops[] is in shared memory by now)
        const unsigned sl = (unsigned)ops[o].pad;
        if ((ops[o].kind == OP_STAR || ops[o].kind == OP_HSTAR) && sl < STAR_SMEM_SLOTS && tid < STAR_IN_ENTRIES)
            starIn[sl][tid] = __ldg(&tabs[ops[o].tab].in[0][0] + tid);
    }
    __syncthreads();

    const qindex numTiles = hdr.numTiles;
    const int numChunks = hdr.numChunks;
    const unsigned chunkBytes = (unsigned)hdr.chunkAmps * (unsigned)sizeof(cplx);
    const qindex myCount = (numTiles > (qindex)blockIdx.x) ? (numTiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int lane = tid & 31;

    if (tid >= TILE_THREADS) {
        // the copy warpgroup hands most of its registers to the two compute warpgroups; only its first warp works
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(COPY_REGS));
        if (tid >= TILE_THREADS + 32) return;
        // ---------------- copy warp: keeps every stage either loading, being computed, or draining ----------------
        auto issue_load = [&](qindex k) {
            const int s = (int)(k % TILE_STAGES);
            const qindex base = hdr.tileIns((qindex)blockIdx.x + k * gridDim.x);
            if (lane == 0) mbar_expect_tx(&full[s], TILE_AMPS * (unsigned)sizeof(cplx));
            __syncwarp();
            cplx* dst = stageBuf + (size_t)s * TILE_AMPS;
            for (int c = lane; c < numChunks; c += 32)
                tma_load(dst + (size_t)c * hdr.chunkAmps, amps + base + hdr.chunkOff[c], chunkBytes, &full[s]);
        };
        for (qindex k = 0; k < TILE_STAGES && k < myCount; k++) issue_load(k);
        for (qindex k = 0; k < myCount; k++) {
            const int s = (int)(k % TILE_STAGES);
            const unsigned parity = (unsigned)((k / TILE_STAGES) & 1);
            const qindex base = hdr.tileIns((qindex)blockIdx.x + k * gridDim.x);
            cplx* t = stageBuf + (size_t)s * TILE_AMPS;
            mbar_wait(&done[s], parity);                       // the compute warps fenced their writes before arriving
            for (int c = lane; c < numChunks; c += 32)
                tma_store(amps + base + hdr.chunkOff[c], t + (size_t)c * hdr.chunkAmps, chunkBytes);
            tma_commit();
            if (k + TILE_STAGES < myCount) {
                tma_wait_read<0>();                            // the stage may be overwritten once its stores have read it
                __syncwarp();
                issue_load(k + TILE_STAGES);
            }
        }
        tma_wait_all<0>();
        return;
    }

    // ---------------- compute warps ----------------
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(COMPUTE_REGS));
    // two warpgroups, each taking every other tile of this CTA and running it alone (own named barrier): their phases
    // drift apart, so one group's shared-memory traffic, barriers and gate dispatch overlap the other's FP64 work
    const int wg = tid / WG_THREADS, wtid = tid % WG_THREADS;
    TT_DECL;
    for (qindex k = wg; k < myCount; k += 2) {
        const int s = (int)(k % TILE_STAGES);
        const unsigned parity = (unsigned)((k / TILE_STAGES) & 1);
        const qindex base = hdr.tileIns((qindex)blockIdx.x + k * gridDim.x);
        cplx* t = stageBuf + (size_t)s * TILE_AMPS;
        cplx* sf = starF[wg][(k >> 1) & 1];

        // per-tile factor of every phase star: the part of its phase that depends on bits OUTSIDE the tile
        if (wtid < numOps && (ops[wtid].kind == OP_STAR || ops[wtid].kind == OP_HSTAR)) {
            const StarTab& tb = tabs[ops[wtid].tab];
            cplx f = mk(1, 0);
#pragma unroll
            for (int sg = 0; sg < STAR_SEGS; sg++) f = cmul(f, __ldg(&tb.ext[sg][(base >> (6 * sg)) & 63]));
            sf[wtid] = f;
        }
        // tile-uniform tests of every op, once per tile: external controls and external phase-star centres
        unsigned long long active = 0;
        for (int b = 0; b < numOps; b += 32) {
            const int o = b + lane;
            bool a = false;
            if (o < numOps) {
                const TileOp& op = ops[o];
                a = ((unsigned long long)base & op.extCtrlMask) == op.extCtrlVals;
                if (op.kind == OP_STAR && op.p0 < 0) a = a && getBit(base, op.e0);
            }
            active |= (unsigned long long)__ballot_sync(0xffffffffu, a) << b;
        }
        // The stage's previous tile (k - 3) belonged to the OTHER warpgroup.  A parity wait only tells the current phase
        // from the preceding one, so this group must not poll `full` for tile k while the barrier may still be in the
        // phase of tile k - 3: first see that tile consumed (its `done` phase is unambiguous here -- the phase before it
        // was completed by this very group), after which `full` can only be in tile k's phase or past it.
        TT_MARK(0);      // per-tile prologue (star factors, tile-uniform tests)
        if (k >= TILE_STAGES) mbar_wait(&done[s], (unsigned)(((k - TILE_STAGES) / TILE_STAGES) & 1));
        TT_MARK(1);      // waiting for the stage's previous tile to be consumed
        mbar_wait(&full[s], parity);
        TT_MARK(2);      // waiting for the tile to land (TMA load)
        wg_sync(wg);
        TT_MARK(3);      // warpgroup barrier at tile start

        for (int r = 0; r < numRounds; r++) {
            const RoundHdr& rd = rounds[r];
            if (rd.kind == ROUND_REG) {
                REG_ROUND(t, rd, ops, base, active, tabs, sf, wtid REG_ROUND_EXTRA(&starIn[0][0]) TT_PASS);
            } else {
                if ((active >> rd.opBase) & 1) smem_pauli(t, ops[rd.opBase], base, wtid);
                TT_MARK(7);
            }
            if (r + 1 < numRounds) { wg_sync(wg); TT_MARK(8); }     // warpgroup barrier between rounds
        }
        // hand the tile to the copy warp: make this warp's shared-memory writes visible to the async proxy, then arrive
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&done[s]);
        TT_MARK(9);      // fence + arrive
    }
    TT_FLUSH;
}

#ifdef QB_TILE_TIMING
extern "C" int qb_tile_timing_read(unsigned long long* out12, int reset) {
    unsigned long long z[TT_SLOTS] = {0};
    cudaDeviceSynchronize();
    if (out12) cudaMemcpyFromSymbol(out12, g_tileTiming, sizeof z);
    if (reset) cudaMemcpyToSymbol(g_tileTiming, z, sizeof z);
    return 0;
}
#endif

// ------------------------------------------------------------------------------------------
// host: planner
// ------------------------------------------------------------------------------------------
struct Pass { std::vector<int> opIdx; unsigned long long high = 0; };
#define PASS_DIRECT (~0ULL)          // Pass::high markers: run the ops one by one through the direct kernels ...
#define PASS_PGROUP (~1ULL)          // ... or together as one coset-blocked Pauli pass (qb_pauli_group.cu)

static inline unsigned long long nonDiagTargets(const QOp& o) {
    switch (o.kind) {
    case OP_DENSE1: case OP_HSTAR: return 1ULL << o.t0;
    case OP_DENSE2: case OP_SWAP: return (1ULL << o.t0) | (1ULL << o.t1);
    case OP_PAULI: return o.maskA;
    default: return 0;
    }
}

// qubits an op involves only DIAGONALLY (controls, diagonal targets, Z sites, phase-star members): two ops commute
// unless one's non-diagonal targets meet any qubit of the other
static inline unsigned long long diagQubits(const QOp& o) {
    unsigned long long d = o.ctrlMask;
    switch (o.kind) {
    case OP_PAULI: d |= o.maskB & ~o.maskA; break;
    case OP_DIAG: d |= 1ULL << o.t0; if (o.numT > 1) d |= 1ULL << o.t1; break;
    case OP_PARITY: d |= o.maskA; break;
    case OP_STAR: d |= 1ULL << o.t0; for (auto& ce : o.star) d |= 1ULL << ce.first; break;
    case OP_HSTAR: for (auto& ce : o.star) d |= 1ULL << ce.first; break;
    default: break;
    }
    return d;
}

// first-fit list scheduling shared by the pass and the round planners: walk `cand` in program order, take an op when
// `fits(op)` accepts it and it commutes with every op skipped so far; skipped ops block their qubits.  Taken ops keep
// their relative order, and each one only moves ahead of ops it commutes with, so the product is unchanged.
template <class Fits>
static void first_fit(const std::vector<QOp>& ops, const std::vector<int>& cand, size_t maxTake, Fits fits,
                      std::vector<int>& taken, std::vector<int>& rest) {
    unsigned long long blockedND = 0, blockedAny = 0;
    for (size_t i = 0; i < cand.size(); i++) {
        const QOp& o = ops[cand[i]];
        const unsigned long long nd = nonDiagTargets(o), dg = diagQubits(o);
        const bool free = !(nd & blockedAny) && !(dg & blockedND);
        if (free && taken.size() < maxTake && fits(o, taken.empty())) taken.push_back(cand[i]);
        else { blockedND |= nd; blockedAny |= nd | dg; rest.push_back(cand[i]); }
    }
}

static inline cplx hmul(cplx a, cplx b) { return mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// 4x4 product helpers for gate absorption; matrices are row-major, index bit 0 <-> t0, bit 1 <-> t1
static void embed1(const cplx* u, int slot, cplx* e) {       // e = u acting on index bit `slot`, identity on the other
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) {
        const int rs = (r >> slot) & 1, cs = (c >> slot) & 1, ro = (r >> (1 - slot)) & 1, co = (c >> (1 - slot)) & 1;
        e[4 * r + c] = (ro == co) ? u[2 * rs + cs] : mk(0, 0);
    }
}
static void matmul(const cplx* a, const cplx* b, int d, cplx* out) {      // out = a * b (d x d)
    cplx t[16];
    for (int r = 0; r < d; r++) for (int c = 0; c < d; c++) {
        cplx s = mk(0, 0);
        for (int k = 0; k < d; k++) { cplx p = hmul(a[d * r + k], b[d * k + c]); s.x += p.x; s.y += p.y; }
        t[d * r + c] = s;
    }
    for (int i = 0; i < d * d; i++) out[i] = t[i];
}

// Gate absorption: a control-free 1-qubit dense gate is multiplied into the neighbouring control-free dense gate on
// the same qubit (the previous one if that is the last op touching the qubit, else the next 2-qubit gate that finds it
// still "open"), and consecutive 2-qubit gates on the same pair are multiplied together.  Host-side 2x2 / 4x4 products;
// the amplitudes then see one gate instead of two.
static void absorb_gates(std::vector<QOp>& ops) {
    std::vector<char> dead(ops.size(), 0);
    int last[64];
    for (int b = 0; b < 64; b++) last[b] = -1;
    auto plain = [&](int i, int kind) { return i >= 0 && !dead[i] && ops[i].kind == kind && ops[i].ctrlMask == 0; };
    for (size_t i = 0; i < ops.size(); i++) {
        QOp& o = ops[i];
        if (o.kind == OP_DENSE1 && o.ctrlMask == 0) {
            const int p = last[o.t0];
            if (plain(p, OP_DENSE1)) { matmul(o.m, ops[p].m, 2, ops[p].m); ops[p].algBytes += o.algBytes; dead[i] = 1; continue; }
            if (plain(p, OP_DENSE2)) {
                cplx e[16]; embed1(o.m, ops[p].t0 == o.t0 ? 0 : 1, e);
                matmul(e, ops[p].m, 4, ops[p].m); ops[p].algBytes += o.algBytes; dead[i] = 1; continue;
            }
        } else if (o.kind == OP_DENSE2 && o.ctrlMask == 0) {
            const int pa = last[o.t0], pb = last[o.t1];
            if (pa == pb && plain(pa, OP_DENSE2)) {          // same pair (either order): multiply into the earlier gate
                QOp& e = ops[pa];
                cplx m2[16];
                const bool flip = e.t0 != o.t0;
                for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) {
                    const int rs = flip ? (((r & 1) << 1) | (r >> 1)) : r, cs = flip ? (((c & 1) << 1) | (c >> 1)) : c;
                    m2[4 * rs + cs] = o.m[4 * r + c];
                }
                matmul(m2, e.m, 4, e.m); e.algBytes += o.algBytes; dead[i] = 1; continue;
            }
            const int tq[2] = {o.t0, o.t1};
            for (int s = 0; s < 2; s++) {
                const int p = last[tq[s]];
                if (plain(p, OP_DENSE1)) { cplx e[16]; embed1(ops[p].m, s, e); matmul(o.m, e, 4, o.m); o.algBytes += ops[p].algBytes; dead[p] = 1; }
            }
        }
        unsigned long long all = nonDiagTargets(o) | diagQubits(o);
        for (int b = 0; b < 64; b++) if ((all >> b) & 1) last[b] = (int)i;
    }
    size_t w = 0;
    for (size_t i = 0; i < ops.size(); i++) if (!dead[i]) { if (w != i) ops[w] = ops[i]; w++; }
    ops.resize(w);
}

static int run_direct(const qb_state* q, const QOp& o);

struct Emitted {
    std::vector<PassHdr> hdrs; std::vector<RoundHdr> rounds; std::vector<TileOp> ops; std::vector<StarTab> tabs;
    std::vector<int> opBase, roundBase;
};

// tile bit set of a pass: the low bits, the required high bits, then filler (lowest unused bits) up to TILE_BITS
static unsigned long long tile_bit_set(int n, const Pass& pass) {
    unsigned long long S = ((1ULL << TILE_LOW) - 1) | pass.high;
    for (int b = TILE_LOW; b < n && __builtin_popcountll(S) < TILE_BITS; b++) if (!((s_restrictMask >> b) & 1)) S |= 1ULL << b;
    return S;
}

// rounds: first-fit again, now over the pass's ops -- a round takes every op whose non-diagonal targets still fit into
// its RB register bits and that commutes with the ops it overtakes.  `order` lists the pass's ops round by round.
static void order_rounds(const std::vector<QOp>& ops, const Pass& pass, unsigned long long S, bool reorder,
                         std::vector<int>& order, std::vector<int>& roundLen, std::vector<int>& roundKind) {
    std::vector<int> remaining = pass.opIdx;
    while (!remaining.empty()) {
        const QOp& first = ops[remaining[0]];
        if (first.kind == OP_PAULI && __builtin_popcountll(first.maskA & S) > RB) {      // shared-memory round of its own
            order.push_back(remaining[0]); roundLen.push_back(1); roundKind.push_back(ROUND_SMEM);
            remaining.erase(remaining.begin());
            continue;
        }
        std::vector<int> taken, rest;
        unsigned long long bits = 0;
        first_fit(ops, remaining, 1000000, [&](const QOp& o, bool) {
            const unsigned long long nb = bits | (nonDiagTargets(o) & S);
            if (__builtin_popcountll(nb) > RB) return false;
            bits = nb; return true;
        }, taken, rest);
        if (!reorder) {
            size_t keep = 0;
            while (keep < taken.size() && taken[keep] == remaining[keep]) keep++;
            taken.resize(keep);
            rest.assign(remaining.begin() + keep, remaining.end());
        }
        for (int idx : taken) order.push_back(idx);
        roundLen.push_back((int)taken.size()); roundKind.push_back(ROUND_REG);
        remaining.swap(rest);
    }
}

static void emit_pass(const qb_state* q, const std::vector<QOp>& ops, const Pass& pass, Emitted& E, bool reorder) {
    const int n = q->logNumAmpsPerNode;
    const unsigned long long S = tile_bit_set(n, pass);
    int pos[64]; int sbits[TILE_BITS]; int T = 0;
    for (int b = 0; b < 64; b++) pos[b] = -1;
    for (int b = 0; b < n; b++) if ((S >> b) & 1) { pos[b] = T; sbits[T++] = b; }
    // controls shared (same qubit, same value) by EVERY op of the pass and lying outside S prune whole tiles
    unsigned long long common = 0, commonVals = 0;
    for (size_t i = 0; i < pass.opIdx.size(); i++) {
        const QOp& o = ops[pass.opIdx[i]];
        unsigned long long m = o.ctrlMask & ~S;
        if (i == 0) { common = m; commonVals = o.ctrlVals & m; }
        else { common &= m; common &= ~((o.ctrlVals ^ commonVals) & common); commonVals &= common; }
    }

    common |= s_restrictMask & ~S; commonVals = (commonVals & ~s_restrictMask) | (s_restrictVals & s_restrictMask & ~S);
    PassHdr h; memset(&h, 0, sizeof h);
    int contiguous = 0; while (contiguous < T && sbits[contiguous] == contiguous) contiguous++;
    h.chunkAmps = 1 << contiguous;
    h.numChunks = 1 << (T - contiguous);
    for (int c = 0; c < h.numChunks; c++) {
        qindex off = 0;
        for (int b = contiguous; b < T; b++) if ((c >> (b - contiguous)) & 1) off |= (qindex)1 << sbits[b];
        h.chunkOff[c] = off;
    }
    int fixedQ[64], fixedV[64], nf = 0;
    for (int b = 0; b < n; b++) {
        if ((S >> b) & 1) { fixedQ[nf] = b; fixedV[nf++] = 0; }
        else if ((common >> b) & 1) { fixedQ[nf] = b; fixedV[nf++] = (int)((commonVals >> b) & 1); }
    }
    h.tileIns = qb_make_ins(fixedQ, fixedV, nf, nullptr, nullptr, 0);
    h.numTiles = (qindex)1 << (n - nf);
    h.numOps = (int)pass.opIdx.size();

    auto toIn = [&](unsigned long long gm) { unsigned v = 0; for (int p = 0; p < T; p++) if ((gm >> sbits[p]) & 1) v |= 1u << p; return v; };
    const size_t opStart = E.ops.size();

    std::vector<int> order, roundLen, roundKind;
    order_rounds(ops, pass, S, reorder, order, roundLen, roundKind);

    std::vector<unsigned> needIn;            // per op: tile-bit positions its non-diagonal targets occupy
    unsigned long long numStars = 0;
    for (int idx : order) {
        const QOp& o = ops[idx];
        TileOp t; memset(&t, 0, sizeof t);
        t.kind = o.kind; t.numT = o.numT;
        t.inCtrlMask = toIn(o.ctrlMask & S); t.inCtrlVals = toIn(o.ctrlVals & o.ctrlMask & S);
        t.extCtrlMask = o.ctrlMask & ~S; t.extCtrlVals = o.ctrlVals & t.extCtrlMask;
        t.p0 = t.p1 = -1;
        for (int i = 0; i < 16; i++) t.m[i] = o.m[i];
        switch (o.kind) {
        case OP_DENSE1: t.p0 = pos[o.t0]; break;
        case OP_DENSE2: case OP_SWAP: t.p0 = pos[o.t0]; t.p1 = pos[o.t1]; break;
        case OP_PAULI: t.inMaskA = toIn(o.maskA); t.inMaskB = toIn(o.maskB & S); t.extMaskB = o.maskB & ~S; break;
        case OP_PARITY: t.inMaskA = toIn(o.maskA & S); t.extMaskB = o.maskA & ~S; break;
        case OP_DIAG:
            t.p0 = pos[o.t0]; t.e0 = o.t0;
            if (o.numT > 1) { t.p1 = pos[o.t1]; t.e1 = o.t1; }
            break;
        case OP_STAR: case OP_HSTAR: {
            t.p0 = pos[o.t0]; t.e0 = o.t0;
            StarTab tb;
            long double angIn[3][16] = {{0}}, angExt[STAR_SEGS][64] = {{0}};
            for (auto& ce : o.star) {
                int c = ce.first; long double th = ce.second;
                if (pos[c] >= 0) { int p = pos[c]; for (int v = 0; v < 16; v++) if ((v >> (p % 4)) & 1) angIn[p / 4][v] += th; }
                else { for (int v = 0; v < 64; v++) if ((v >> (c % 6)) & 1) angExt[c / 6][v] += th; }
            }
            for (int s = 0; s < 3; s++) for (int v = 0; v < 16; v++) tb.in[s][v] = mk((double)cosl(angIn[s][v]), (double)sinl(angIn[s][v]));
            t.pad = numStars++;              // shared-memory slot of its in-tile tables (the kernel keeps the first STAR_SMEM_SLOTS)
            for (int s = 0; s < STAR_SEGS; s++) for (int v = 0; v < 64; v++) tb.ext[s][v] = mk((double)cosl(angExt[s][v]), (double)sinl(angExt[s][v]));
            t.tab = (int)E.tabs.size();
            E.tabs.push_back(tb);
        } break;
        }
        unsigned need = 0;
        if (o.kind == OP_DENSE1 || o.kind == OP_HSTAR) need = 1u << t.p0;
        else if (o.kind == OP_DENSE2 || o.kind == OP_SWAP) need = (1u << t.p0) | (1u << t.p1);
        else if (o.kind == OP_PAULI) need = t.inMaskA;
        needIn.push_back(need);
        E.ops.push_back(t);
    }

    // per-round register bits and the ops' coordinates relative to them
    const size_t roundStart = E.rounds.size();
    unsigned curBits = 0; int curBase = 0, curCount = 0;
    auto close_round = [&]() {
        if (!curCount) return;
        // fill the unused register bits with the highest free tile positions (keeps shared-memory accesses conflict-free)
        for (int p = T - 1; p >= 0 && __builtin_popcount(curBits) < RB; p--) if (!((curBits >> p) & 1)) curBits |= 1u << p;
        RoundHdr r; memset(&r, 0, sizeof r);
        r.kind = ROUND_REG; r.opBase = curBase; r.numOps = curCount;
        int k = 0, local[TILE_BITS];
        for (int p = 0; p < T; p++) { local[p] = -1; if ((curBits >> p) & 1) { local[p] = k; r.b[k++] = p; } }
        for (int o = curBase; o < curBase + curCount; o++) {
            TileOp& t = E.ops[opStart + o];
            const int hasCtrl = t.inCtrlMask != 0;
            if (t.kind == OP_DENSE1) { t.l0 = local[t.p0]; t.code = CODE_DENSE1 + 2 * t.l0 + hasCtrl; }
            else if (t.kind == OP_SWAP) { t.l0 = std::min(local[t.p0], local[t.p1]); t.l1 = std::max(local[t.p0], local[t.p1]); t.code = CODE_SWAP + pair_index(t.l0, t.l1); }
            else if (t.kind == OP_DENSE2) {
                int a = local[t.p0], b = local[t.p1];
                if (a > b) {        // re-order the matrix so that its index bit 0 belongs to the lower register bit
                    cplx m2[16];
                    for (int r2 = 0; r2 < 4; r2++) for (int c2 = 0; c2 < 4; c2++) {
                        int rs = ((r2 & 1) << 1) | (r2 >> 1), cs = ((c2 & 1) << 1) | (c2 >> 1);
                        m2[4 * rs + cs] = t.m[4 * r2 + c2];
                    }
                    for (int i = 0; i < 16; i++) t.m[i] = m2[i];
                    std::swap(a, b);
                }
                t.l0 = a; t.l1 = b; t.code = CODE_DENSE2 + 2 * pair_index(a, b) + hasCtrl;
            } else if (t.kind == OP_STAR || t.kind == OP_HSTAR) {
                // per-register phase factors over the round's own bits; the tables keep every other bit
                const QOp& qo = ops[order[o]];
                t.l0 = (t.p0 >= 0) ? local[t.p0] : -1;
                long double ang[RAMPS] = {0};
                for (auto& ce : qo.star) {
                    int p = pos[ce.first];
                    if (p >= 0 && local[p] >= 0)
                        for (int u = 0; u < RAMPS; u++) if ((u >> local[p]) & 1) ang[u] += ce.second;
                }
                for (int u = 0; u < RAMPS; u++) t.m[u] = mk((double)cosl(ang[u]), (double)sinl(ang[u]));
                // centre among the round's bits: specialised body; centre elsewhere in the tile: a per-thread test
                // carried by the in-tile control fields; centre outside the tile: the per-tile `active` test
                if (t.kind == OP_HSTAR) t.code = CODE_HSTAR + t.l0;
                else if (t.l0 >= 0) t.code = CODE_STAR + t.l0;
                else { t.code = CODE_STAR + 4; if (t.p0 >= 0) { t.inCtrlMask = 1u << t.p0; t.inCtrlVals = 1u << t.p0; } }
            } else if (t.kind == OP_PAULI) {
                t.lmaskA = t.lmaskB = 0;
                for (int p = 0; p < T; p++) if (local[p] >= 0) {
                    if ((t.inMaskA >> p) & 1) t.lmaskA |= 1u << local[p];
                    if ((t.inMaskB >> p) & 1) t.lmaskB |= 1u << local[p];
                }
                t.inMaskB &= ~curBits;          // the kernel adds the round bits' parity through lmaskB
                t.code = CODE_PAULI + (int)t.lmaskA - 1;
            } else if (t.kind == OP_DIAG) t.code = CODE_DIAG;
            else if (t.kind == OP_PARITY) t.code = CODE_PARITY;
        }
        E.rounds.push_back(r);
        curBits = 0; curCount = 0;
    };
    for (size_t r = 0, o = 0; r < roundLen.size(); o += roundLen[r], r++) {
        if (roundKind[r] == ROUND_SMEM) {
            RoundHdr rh; memset(&rh, 0, sizeof rh);
            rh.kind = ROUND_SMEM; rh.opBase = (int)o; rh.numOps = 1;
            E.rounds.push_back(rh);
            continue;
        }
        curBase = (int)o; curCount = roundLen[r]; curBits = 0;
        for (int k = 0; k < curCount; k++) curBits |= needIn[o + k];
        close_round();
    }
    h.numRounds = (int)(E.rounds.size() - roundStart);
    // opBase inside RoundHdr is relative to the pass's first op
    E.hdrs.push_back(h);
    E.opBase.push_back((int)opStart);
    E.roundBase.push_back((int)roundStart);
}

static bool is_cphase(const QOp& o) {
    if (o.kind != OP_DIAG || o.numT != 1) return false;
    if (__builtin_popcountll(o.ctrlMask) != 1 || o.ctrlVals != o.ctrlMask) return false;
    if (o.m[0].x != 1.0 || o.m[0].y != 0.0) return false;
    return fabs(hypot(o.m[1].x, o.m[1].y) - 1.0) < 1e-14;
}

// device scratch for pass descriptors (grown on demand)
static char* s_devDesc = nullptr; static size_t s_devDescBytes = 0;

// the planner proper (host only, no CUDA): gate absorption, phase-star merging, Hadamard+star fusion, pass grouping
static void plan_passes(std::vector<QOp>& ops, bool reorder, std::vector<QOp>& merged, std::vector<Pass>& passes) {
    if (reorder) absorb_gates(ops);

    // 1. merge ladders of controlled phases that share a qubit into phase stars
    merged.clear();
    for (size_t i = 0; i < ops.size(); ) {
        if (is_cphase(ops[i])) {
            size_t j = i + 1;
            int a = ops[i].t0, b = __builtin_ctzll(ops[i].ctrlMask), centre = -1;
            while (j < ops.size() && is_cphase(ops[j])) {
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

    // 1b. a Hadamard immediately followed by the phase star centred on the same qubit is one QFT stage: fuse them
    if (reorder) {
        std::vector<QOp> fused;
        auto isH = [](const QOp& o) {
            if (o.kind != OP_DENSE1 || o.ctrlMask) return false;
            const double s = 0.70710678118654752440, e = 1e-15;
            return fabs(o.m[0].x - s) < e && fabs(o.m[1].x - s) < e && fabs(o.m[2].x - s) < e && fabs(o.m[3].x + s) < e &&
                   o.m[0].y == 0 && o.m[1].y == 0 && o.m[2].y == 0 && o.m[3].y == 0;
        };
        for (size_t i = 0; i < merged.size(); i++) {
            if (i + 1 < merged.size() && isH(merged[i]) && merged[i + 1].kind == OP_STAR && merged[i + 1].t0 == merged[i].t0) {
                QOp hs = merged[i + 1];
                hs.kind = OP_HSTAR; hs.algBytes += merged[i].algBytes;
                fused.push_back(hs);
                i++;
            } else fused.push_back(merged[i]);
        }
        merged.swap(fused);
    }

    // 2. grouping into passes: first-fit over the whole queue -- a pass takes every op (in program order) whose high
    //    non-diagonal targets still fit into its six free tile bits and that commutes with everything it overtakes
    passes.clear();
    const int maxHigh = TILE_BITS - TILE_LOW;
    const unsigned long long lowMask = (1ULL << TILE_LOW) - 1;
    auto highNeed = [&](const QOp& o) {
        unsigned long long need = nonDiagTargets(o) & ~lowMask;
        if (o.kind == OP_STAR && o.t0 >= TILE_LOW) need |= 1ULL << o.t0;       // keep the centre in-tile when cheap
        return need;
    };
    std::vector<int> remaining(merged.size());
    for (size_t i = 0; i < merged.size(); i++) remaining[i] = (int)i;
    while (!remaining.empty()) {
        Pass cur;
        {
            // A Pauli string with X/Y on more than six high qubits fits no tile.  It and the control-free Pauli / parity
            // gadgets around it in program order (Trotter circuits are nothing else) share ONE pass as long as their
            // X/Y masks stay linearly independent: they only mix amplitudes within cosets of the masks' span.  A group
            // is formed when it starts with a control-free Pauli string and holds at least one string that fits no tile
            // (narrow strings between wide ones ride along instead of costing a pass of their own).
            const QOp& head = merged[remaining[0]];
            const bool headWide = __builtin_popcountll(highNeed(head)) > maxHigh;
            unsigned long long masks[PG_K]; int nm = 0; size_t take = 0; bool anyWide = false;
            if (head.kind == OP_PAULI && !head.ctrlMask)
                for (; take < remaining.size() && take < PG_MAX_OPS; take++) {
                    const QOp& o = merged[remaining[take]];
                    if (o.ctrlMask || (o.kind != OP_PAULI && o.kind != OP_PARITY)) break;
                    if (o.kind == OP_PAULI) {
                        if (nm == PG_K_GADGET) break;
                        masks[nm] = o.maskA;
                        if (pg_rank(masks, nm + 1, nullptr) != nm + 1) break;
                        nm++;
                        anyWide |= __builtin_popcountll(highNeed(o)) > maxHigh;
                    }
                }
            const bool group = take >= 2 && anyWide;
            if (group || headWide) {
                if (!group) take = 1;
                cur.opIdx.assign(remaining.begin(), remaining.begin() + take);
                cur.high = group ? PASS_PGROUP : PASS_DIRECT;
                remaining.erase(remaining.begin(), remaining.begin() + take);
                passes.push_back(cur);
                continue;
            }
        }
        std::vector<int> rest;
        first_fit(merged, remaining, reorder ? MAX_OPS_PER_PASS : 1000000, [&](const QOp& o, bool) {
            const unsigned long long nh = cur.high | highNeed(o);
            if (__builtin_popcountll(nh) > maxHigh) return false;
            cur.high = nh; return true;
        }, cur.opIdx, rest);
        if (!reorder) {
            // program order only: cut the pass at the first op that did not fit
            size_t keep = 0;
            while (keep < cur.opIdx.size() && cur.opIdx[keep] == remaining[keep] && keep < MAX_OPS_PER_PASS) keep++;
            cur.opIdx.resize(keep); cur.high = 0;
            for (int idx : cur.opIdx) cur.high |= highNeed(merged[idx]);
            rest.assign(remaining.begin() + keep, remaining.end());
        }
        remaining.swap(rest);
        passes.push_back(cur);
    }

}

// FP64 fused multiply-adds per touched amplitude of each op kind (complex arithmetic written out: cmul = 4, cfma = 4)
static double fma_per_amp(const QOp& o) {
    double f;
    const double share = 1.0 / (double)(1ULL << __builtin_popcountll(s_restrictMask));
    switch (o.kind) {
    case OP_DENSE1: f = 8; break;
    case OP_DENSE2: f = 16; break;
    case OP_PAULI: f = 8; break;
    case OP_SWAP: f = 0; break;
    case OP_HSTAR: f = 7; break;
    default: f = 4; break;                    // diagonal / parity / star: one complex multiply
    }
    return share * f / (double)(1ULL << __builtin_popcountll(o.ctrlMask));
}

static int flush_one(StateQueue& sq);

static int flush_queue() {
    if (s_queues.empty() || s_inFlush) return 0;
    int rc = 0;
    std::vector<StateQueue> all; all.swap(s_queues);
    for (StateQueue& sq : all) { int r = flush_one(sq); if (r && !rc) rc = r; }
    return rc;
}

static int flush_one(StateQueue& sq) {
    if (sq.ops.empty()) return 0;
    s_inFlush = true;
    s_flushEpoch++;
    std::vector<QOp> ops; ops.swap(sq.ops);
    qb_state q = sq.st;
    s_statQueuedGates += ops.size();
    const int n = q.logNumAmpsPerNode;
    int rc = 0;
    const bool reorder = g_qb.tileEngine != 2;
    std::vector<QOp> merged; std::vector<Pass> passes;
    plan_passes(ops, reorder, merged, passes);

    // 3. emit: single-op passes use the direct kernels (already at the HBM roofline), multi-op passes the tile kernel
    Emitted E;
    std::vector<int> passKind, passArg;     // -1: direct op index, else index into E.hdrs
    for (auto& p : passes) {
        bool direct = (p.high == PASS_DIRECT) || (p.opIdx.size() == 1 && merged[p.opIdx[0]].kind != OP_STAR && merged[p.opIdx[0]].kind != OP_HSTAR) || (n < TILE_BITS && p.high != PASS_PGROUP);
        if (p.high == PASS_PGROUP && p.opIdx.size() > 1) { passKind.push_back(-2); passArg.push_back((int)(&p - &passes[0])); }
        else if (direct) { for (int idx : p.opIdx) { passKind.push_back(-1); passArg.push_back(idx); } }
        else {
            passKind.push_back((int)E.hdrs.size()); passArg.push_back(0); emit_pass(&q, merged, p, E, reorder);
            for (int idx : p.opIdx) s_statFmaAmps += fma_per_amp(merged[idx]) * (double)q.numAmpsPerNode;
        }
    }
    const size_t bh = E.hdrs.size() * sizeof(PassHdr), br = E.rounds.size() * sizeof(RoundHdr),
                 bo = E.ops.size() * sizeof(TileOp), bt = E.tabs.size() * sizeof(StarTab);
    if (!E.hdrs.empty()) {
        size_t need = bh + br + bo + bt + 256;
        if (need > s_devDescBytes) {
            cudaStreamSynchronize(g_qb.stream);
            if (s_devDesc) cudaFree(s_devDesc);
            s_devDescBytes = need * 2;
            if (cudaMalloc(&s_devDesc, s_devDescBytes) != cudaSuccess) { s_devDesc = nullptr; s_devDescBytes = 0; rc = qb_set_error((int)cudaErrorMemoryAllocation, "tile descriptors", __FILE__, __LINE__); }
        }
        if (!rc) {
            // pageable source: the runtime stages the bytes before returning, so the vectors may die right after
            cudaMemcpyAsync(s_devDesc, E.hdrs.data(), bh, cudaMemcpyHostToDevice, g_qb.stream);
            cudaMemcpyAsync(s_devDesc + bh, E.rounds.data(), br, cudaMemcpyHostToDevice, g_qb.stream);
            cudaMemcpyAsync(s_devDesc + bh + br, E.ops.data(), bo, cudaMemcpyHostToDevice, g_qb.stream);
            if (bt) cudaMemcpyAsync(s_devDesc + bh + br + bo, E.tabs.data(), bt, cudaMemcpyHostToDevice, g_qb.stream);
        }
    }
    static bool attrSet = false;
    const size_t smemBytes = (size_t)TILE_STAGES * TILE_AMPS * sizeof(cplx) + (size_t)MAX_OPS_PER_PASS * sizeof(TileOp);
    if (!attrSet && !rc) {
        cudaError_t e = cudaFuncSetAttribute(k_tile_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes);
        if (e != cudaSuccess) rc = qb_set_error((int)e, "cudaFuncSetAttribute(k_tile_pass)", __FILE__, __LINE__);
        attrSet = true;
    }
    for (size_t i = 0; i < passKind.size() && !rc; i++) {
        const double shardBytes = 2.0 * sizeof(cplx) * (double)q.numAmpsPerNode / (double)(1ULL << __builtin_popcountll(s_restrictMask));
        if (passKind[i] == -2) {              // coset-blocked Pauli pass
            s_statPassBytes += shardBytes;
            const Pass& pp = passes[passArg[i]];
            PGOp gops[PG_MAX_OPS]; int ng = 0;
            for (int idx : pp.opIdx) {
                const QOp& o = merged[idx];
                PGOp& g = gops[ng++];
                if (o.kind == OP_PAULI) { g.xy = o.maskA; g.yz = o.maskB; } else { g.xy = 0; g.yz = o.maskA; }
                g.c = o.m[0]; g.f = o.m[1];
                s_statFmaAmps += fma_per_amp(o) * (double)q.numAmpsPerNode;
            }
            rc = qb_pauli_group_apply(&q, gops, ng, s_restrictMask, s_restrictVals);
            s_statPasses++; s_statTileOps += ng;
            continue;
        }
        if (passKind[i] < 0) { s_statPassBytes += shardBytes / (double)(1ULL << __builtin_popcountll(merged[passArg[i]].ctrlMask)) * (merged[passArg[i]].kind == OP_SWAP ? 0.5 : 1.0);
                               rc = run_direct(&q, merged[passArg[i]]); s_statDirectOps++; s_statFmaAmps += fma_per_amp(merged[passArg[i]]) * (double)q.numAmpsPerNode; continue; }
        const int hi = passKind[i];
        s_statPasses++; s_statRounds += E.hdrs[hi].numRounds; s_statTileOps += E.hdrs[hi].numOps;
        s_statPassBytes += 2.0 * sizeof(cplx) * (double)TILE_AMPS * (double)E.hdrs[hi].numTiles;
        const PassHdr* dh = (const PassHdr*)s_devDesc + hi;
        const RoundHdr* dr = (const RoundHdr*)(s_devDesc + bh) + E.roundBase[hi];
        const TileOp* dops = (const TileOp*)(s_devDesc + bh + br) + E.opBase[hi];
        const StarTab* dt = (const StarTab*)(s_devDesc + bh + br + bo);
        unsigned grid = (unsigned)std::min<qindex>(E.hdrs[hi].numTiles, g_qb.numSMs);
        k_tile_pass<<<grid, TILE_BLOCK, smemBytes, g_qb.stream>>>((cplx*)q.amps, dh, dr, dops, dt);
        g_qb.launches++;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) rc = qb_set_error((int)e, "k_tile_pass launch", __FILE__, __LINE__);
    }
    s_inFlush = false;
    if (rc) s_status = rc;
    return rc;
}

int qb_flush_internal() { return flush_queue(); }

// what the deferred queue of this state still holds: number of gates, and every suffix qubit any of them involves
// (targets, controls, diagonal / Z sites).  The sharding layer uses it to run a half-shard swap AHEAD of the queued
// gates when they commute with it (no queued gate touches the swapped suffix qubit), instead of flushing the queue.
static StateQueue* find_queue(const qb_state* q) {
    for (StateQueue& sq : s_queues) if (sq.st.amps == q->amps) return &sq;
    return nullptr;
}

extern "C" int qb_queue_info(const qb_state* q, unsigned long long* touchedSuffixMask, unsigned long long* flushEpoch) {
    unsigned long long mask = 0; int len = 0;
    if (q) if (StateQueue* sq = find_queue(q)) {
        len = (int)sq->ops.size();
        for (const QOp& o : sq->ops) mask |= nonDiagTargets(o) | diagQubits(o);
    }
    if (touchedSuffixMask) *touchedSuffixMask = mask;
    if (flushEpoch) *flushEpoch = s_flushEpoch;
    return len;
}

// cumulative planner / execution statistics since the library was loaded:
// out[0] tile passes launched, [1] rounds in them, [2] ops executed inside tile passes, [3] ops run as direct kernels,
// [4] gates received by the queue, [5] FP64 fused multiply-adds executed for them (thread-level count),
// [6] bytes those passes and kernels streamed through HBM (what they read + wrote), [7] reserved
extern "C" int qb_tile_stats(double out[8]) {
    if (!out) return -1;
    out[6] = s_statPassBytes; out[7] = 0;
    out[0] = (double)s_statPasses; out[1] = (double)s_statRounds; out[2] = (double)s_statTileOps; out[3] = (double)s_statDirectOps;
    out[4] = (double)s_statQueuedGates; out[5] = s_statFmaAmps;
    return 0;
}

// Runs the queue of `q` on the half (quarter ...) of the shard whose index bits `mask` hold `vals`.  keepQueue: the gates
// stay queued (the caller will run them on the remaining amplitudes later).  No queued gate may involve a bit of `mask`.
int qb_tile_flush_restricted(const qb_state* q, unsigned long long mask, unsigned long long vals, bool keepQueue) {
    StateQueue* sq = find_queue(q);
    if (!sq || sq->ops.empty() || s_inFlush) return 0;
    // a restriction is realised by pruning whole tiles, so its bits must lie outside every tile: not among the low
    // TILE_LOW index bits, which belong to all tiles
    if (mask & ((1ULL << TILE_LOW) - 1)) return qb_set_error(-1, "restricted flush: restricted bits must be >= 6", __FILE__, __LINE__);
    for (const QOp& o : sq->ops)
        if ((nonDiagTargets(o) | diagQubits(o)) & mask) return qb_set_error(-1, "restricted flush: a queued gate involves a restricted bit", __FILE__, __LINE__);
    StateQueue work = *sq;
    if (!keepQueue) sq->ops.clear();
    s_restrictMask = mask; s_restrictVals = vals & mask;
    int r = flush_one(work);
    s_restrictMask = s_restrictVals = 0;
    return r;
}

// called by qb_free: whatever is still queued for memory about to be released must run first (QB_READY in qb_free
// has already flushed); nothing may keep referring to the pointer
void qb_tile_forget(const void* amps) {
    for (size_t i = 0; i < s_queues.size(); i++) if (s_queues[i].st.amps == amps) { s_queues.erase(s_queues.begin() + i); return; }
}

// ------------------------------------------------------------------------------------------
// enqueue API used by the per-gate entry points
// ------------------------------------------------------------------------------------------
static bool can_fuse(const qb_state* q) {
    return g_qb.tileEngine && !s_inFlush && q->logNumAmpsPerNode >= FUSE_MIN_LOG_AMPS && q->logNumAmpsPerNode <= 36;
}

static int enqueue(const qb_state* q, QOp& o) {
    s_status = 0;
    StateQueue* sq = find_queue(q);
    if (sq && (sq->st.numAmpsPerNode != q->numAmpsPerNode || sq->st.rank != q->rank || sq->st.logNumAmpsPerNode != q->logNumAmpsPerNode)) {
        int r = flush_one(*sq); if (r) { s_status = r; return 1; }      // same memory seen through a different view
        sq->st = *q;
    }
    if (!sq) {
        // a newcomer whose memory overlaps a queued state's (a view into the middle of another allocation) must not
        // overtake it; and the table is kept small: the oldest queue is run when it is full
        const char* lo = (const char*)q->amps; const char* hi = lo + (size_t)q->numAmpsPerNode * sizeof(cplx);
        for (size_t i = 0; i < s_queues.size(); ) {
            const char* a = (const char*)s_queues[i].st.amps; const char* b = a + (size_t)s_queues[i].st.numAmpsPerNode * sizeof(cplx);
            if (a < hi && lo < b) { int r = flush_one(s_queues[i]); s_queues.erase(s_queues.begin() + i); if (r) { s_status = r; return 1; } }
            else i++;
        }
        if (s_queues.size() >= MAX_STATE_QUEUES) { int r = flush_one(s_queues[0]); s_queues.erase(s_queues.begin()); if (r) { s_status = r; return 1; } }
        s_queues.push_back(StateQueue{*q, {}});
        sq = &s_queues.back();
    }
    sq->ops.push_back(o);
    if (sq->ops.size() >= QUEUE_MAX) s_status = flush_one(*sq);
    // Trotter-like streams (nothing but control-free Pauli / phase gadgets): the planner groups them four per pass in
    // program order, so a longer queue buys no better plan -- start executing every PAULI_STREAM_FLUSH gadgets, so that
    // the device works while the host is still issuing the rest (the API's per-gadget host cost is ~0.15 ms)
    else if (sq->ops.size() >= PAULI_STREAM_FLUSH && sq->ops.size() % PAULI_STREAM_FLUSH == 0) {
        bool stream = true;
        for (size_t i = sq->ops.size() - PAULI_STREAM_FLUSH; i < sq->ops.size() && stream; i++) {
            const QOp& g = sq->ops[i];
            stream = !g.ctrlMask && (g.kind == OP_PARITY || g.kind == OP_PAULI);
        }
        if (stream && sq->ops.size() == PAULI_STREAM_FLUSH) s_status = flush_one(*sq);
    }
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
    const unsigned long long cmask = o.ctrlMask | s_restrictMask, cvals = (o.ctrlVals & ~s_restrictMask) | s_restrictVals;
    for (int b = 0; b < 64; b++) if ((cmask >> b) & 1) { ctrls[nc] = b; cs[nc++] = (int)((cvals >> b) & 1); }
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
    case OP_HSTAR: {
        const double s = 0.70710678118654752440;
        qb_cplx h[4] = {{s, 0}, {s, 0}, {s, 0}, {-s, 0}};
        int r = qb_statevec_anyCtrlOneTargDenseMatr_subA(q, ctrls, cs, nc, o.t0, h);
        if (r) return r;
    }   // fall through: then the star's controlled phases
    case OP_STAR: {
        // a star that did not end up in a tile pass: apply its controlled phases one by one
        for (auto& ce : o.star) {
            ctrls[nc] = ce.first; cs[nc] = 1;
            qb_cplx e[2] = {{1, 0}, {(qb_real)cos(ce.second), (qb_real)sin(ce.second)}};
            int r = qb_statevec_anyCtrlOneTargDiagMatr_sub(q, ctrls, cs, nc + 1, o.t0, e);
            if (r) return r;
        }
        return 0;
    }
    }
    return qb_set_error(-1, "tile engine: unknown op", __FILE__, __LINE__);
}

#if defined(QB_SELFTEST) && !defined(QB_TILE_TIMING)
// ------------------------------------------------------------------------------------------
// host-only self-test of the planner (no CUDA): a random gate list is applied to a small random state vector twice --
// gate by gate in program order, and in the planner's order (after gate absorption, star merging, Hadamard+star
// fusion, pass grouping and round ordering) -- with a plain host simulator.  Exercised by tests/test_abi_cpu.py.
// ------------------------------------------------------------------------------------------
#include <complex>
#include <random>

typedef std::complex<double> hc;
static inline hc tohc(cplx c) { return hc(c.x, c.y); }

static void host_apply(std::vector<hc>& a, const QOp& o) {
    const unsigned long long N = a.size();
    auto ok = [&](unsigned long long i) { return (i & o.ctrlMask) == o.ctrlVals; };
    switch (o.kind) {
    case OP_DENSE1: {
        const unsigned long long t = 1ULL << o.t0;
        for (unsigned long long i = 0; i < N; i++) if (!(i & t) && ok(i)) {
            hc x = a[i], y = a[i | t];
            a[i] = tohc(o.m[0]) * x + tohc(o.m[1]) * y; a[i | t] = tohc(o.m[2]) * x + tohc(o.m[3]) * y;
        }
    } break;
    case OP_DENSE2: {
        const unsigned long long t0 = 1ULL << o.t0, t1 = 1ULL << o.t1;
        for (unsigned long long i = 0; i < N; i++) if (!(i & (t0 | t1)) && ok(i)) {
            hc v[4] = {a[i], a[i | t0], a[i | t1], a[i | t0 | t1]}, w[4];
            for (int r = 0; r < 4; r++) { w[r] = 0; for (int c = 0; c < 4; c++) w[r] += tohc(o.m[4 * r + c]) * v[c]; }
            a[i] = w[0]; a[i | t0] = w[1]; a[i | t1] = w[2]; a[i | t0 | t1] = w[3];
        }
    } break;
    case OP_SWAP: {
        const unsigned long long t0 = 1ULL << o.t0, t1 = 1ULL << o.t1;
        for (unsigned long long i = 0; i < N; i++) if ((i & t0) && !(i & t1) && ok(i)) std::swap(a[i], a[i ^ t0 ^ t1]);
    } break;
    case OP_DIAG:
        for (unsigned long long i = 0; i < N; i++) if (ok(i)) {
            int k = (int)((i >> o.t0) & 1);
            if (o.numT > 1) k |= (int)((i >> o.t1) & 1) << 1;
            a[i] *= tohc(o.m[k]);
        }
        break;
    case OP_PARITY:
        for (unsigned long long i = 0; i < N; i++) if (ok(i)) a[i] *= tohc(o.m[__builtin_parityll(i & o.maskA)]);
        break;
    case OP_PAULI: {
        const int h = 63 - __builtin_clzll(o.maskA);
        for (unsigned long long i = 0; i < N; i++) if (!((i >> h) & 1) && ok(i)) {
            const unsigned long long w = i ^ o.maskA;
            const double si = __builtin_parityll(i & o.maskB) ? -1.0 : 1.0, sw = __builtin_parityll(w & o.maskB) ? -1.0 : 1.0;
            hc x = a[i], y = a[w];
            a[i] = tohc(o.m[0]) * x + tohc(o.m[1]) * sw * y; a[w] = tohc(o.m[0]) * y + tohc(o.m[1]) * si * x;
        }
    } break;
    case OP_HSTAR: {
        const unsigned long long t = 1ULL << o.t0; const double s = 0.70710678118654752440;
        for (unsigned long long i = 0; i < N; i++) if (!(i & t)) { hc x = a[i], y = a[i | t]; a[i] = s * (x + y); a[i | t] = s * (x - y); }
    }   // fall through to the star
    case OP_STAR:
        for (unsigned long long i = 0; i < N; i++) if ((i >> o.t0) & 1) {
            double ang = 0;
            for (auto& ce : o.star) if ((i >> ce.first) & 1) ang += ce.second;
            a[i] *= hc(cos(ang), sin(ang));
        }
        break;
    }
}

// random gate list for the self-tests: every fusable kind, random controls, QFT-like stages
static void selftest_random_ops(int n, int numOps, std::mt19937_64& rng, std::vector<QOp>& ops) {
    auto rnd = [&](int k) { return (int)(rng() % (unsigned long long)k); };
    auto unif = [&]() { return (double)(rng() >> 11) * (1.0 / 9007199254740992.0); };
    auto rc = [&]() { return mk(2 * unif() - 1, 2 * unif() - 1); };
    qb_state q; memset(&q, 0, sizeof q); q.numAmpsPerNode = 1LL << n; q.logNumAmpsPerNode = n; q.numQubits = n;
    auto pick = [&](int k, int* out) { for (int i = 0; i < k; ) { int c = rnd(n); bool dup = false; for (int j = 0; j < i; j++) dup |= out[j] == c; if (!dup) out[i++] = c; } };
    while ((int)ops.size() < numOps) {
        const int kind = rnd(10), nc = (rnd(3) == 0) ? rnd(3) : 0;
        int qs[12]; pick(std::min(n, nc + 8), qs);
        QOp o = blank(OP_DENSE1, &q, nc);
        for (int i = 0; i < nc; i++) { o.ctrlMask |= 1ULL << qs[i]; if (rnd(2)) o.ctrlVals |= 1ULL << qs[i]; }
        const int* t = qs + nc;
        if (kind < 3) { o.kind = OP_DENSE1; o.t0 = t[0]; for (int i = 0; i < 4; i++) o.m[i] = rc(); ops.push_back(o); }
        else if (kind < 6) { o.kind = OP_DENSE2; o.numT = 2; o.t0 = t[0]; o.t1 = t[1]; for (int i = 0; i < 16; i++) o.m[i] = rc(); ops.push_back(o); }
        else if (kind == 6) { o.kind = OP_SWAP; o.numT = 2; o.t0 = t[0]; o.t1 = t[1]; ops.push_back(o); }
        else if (kind == 7) {
            if (rnd(2)) { o.kind = OP_DIAG; o.numT = 1 + rnd(2); o.t0 = t[0]; o.t1 = t[1]; for (int i = 0; i < 4; i++) o.m[i] = rc(); }
            else { o.kind = OP_PARITY; for (int i = 0; i < 1 + rnd(3); i++) o.maskA |= 1ULL << t[i]; o.m[0] = rc(); o.m[1] = rc(); }
            ops.push_back(o);
        } else if (kind == 8) {
            o.kind = OP_PAULI;
            const int k = 1 + rnd(std::min(rnd(3) ? 4 : 8, n - nc));        // long strings: shared-memory rounds / direct passes
            for (int i = 0; i < k; i++) { int p = rnd(3); if (p != 2) o.maskA |= 1ULL << t[i]; if (p != 0) o.maskB |= 1ULL << t[i]; }
            if (!o.maskA) o.maskA = 1ULL << t[0];
            o.m[0] = rc(); o.m[1] = rc();
            ops.push_back(o);
        } else {
            // a QFT stage: Hadamard (sometimes absent) + a ladder of controlled phases on the same qubit
            const int centre = rnd(n); const double s = 0.70710678118654752440;
            if (rnd(4)) { QOp h = blank(OP_DENSE1, &q, 0); h.t0 = centre; h.m[0] = h.m[1] = h.m[2] = mk(s, 0); h.m[3] = mk(-s, 0); ops.push_back(h); }
            const int len = 1 + rnd(std::min(n - 1, 6));
            int others[8]; int cnt = 0;
            while (cnt < len) { int c = rnd(n); bool bad = c == centre; for (int j = 0; j < cnt; j++) bad |= others[j] == c; if (!bad) others[cnt++] = c; }
            for (int i = 0; i < len; i++) {
                QOp d = blank(OP_DIAG, &q, 1); const double th = 6.28 * unif();
                const bool flip = rnd(2);
                d.t0 = flip ? others[i] : centre; d.ctrlMask = d.ctrlVals = 1ULL << (flip ? centre : others[i]);
                d.m[0] = mk(1, 0); d.m[1] = mk(cos(th), sin(th));
                ops.push_back(d);
            }
        }
    }
}

extern "C" int qb_selftest_planner(int numQubits, int numOps, unsigned seed, int reorder, double* maxErr, int* numPasses, int* numRounds, int* numOpsPlanned) {
    if (numQubits < 3 || numQubits > 20 || numOps < 1) return -1;
    std::mt19937_64 rng(seed);
    auto unif = [&]() { return (double)(rng() >> 11) * (1.0 / 9007199254740992.0); };
    const int n = numQubits;
    std::vector<QOp> ops;
    selftest_random_ops(n, numOps, rng, ops);

    std::vector<hc> ref((size_t)1 << n), got;
    for (auto& v : ref) v = hc(2 * unif() - 1, 2 * unif() - 1);
    got = ref;
    for (const QOp& o : ops) host_apply(ref, o);

    std::vector<QOp> queue = ops, merged; std::vector<Pass> passes;
    plan_passes(queue, reorder != 0, merged, passes);
    int rounds = 0, planned = 0;
    for (const Pass& p : passes) {
        if (p.high == PASS_DIRECT || p.high == PASS_PGROUP || n < TILE_BITS) { for (int idx : p.opIdx) { host_apply(got, merged[idx]); planned++; } continue; }
        std::vector<int> order, roundLen, roundKind;
        order_rounds(merged, p, tile_bit_set(n, p), reorder != 0, order, roundLen, roundKind);
        if (order.size() != p.opIdx.size()) return -2;
        for (int idx : order) { host_apply(got, merged[idx]); planned++; }
        rounds += (int)roundLen.size();
    }
    if (planned != (int)merged.size()) return -3;
    double err = 0, norm = 0;
    for (size_t i = 0; i < ref.size(); i++) { err = std::max(err, std::abs(ref[i] - got[i])); norm = std::max(norm, std::abs(ref[i])); }
    if (maxErr) *maxErr = norm > 0 ? err / norm : err;
    if (numPasses) *numPasses = (int)passes.size();
    if (numRounds) *numRounds = rounds;
    if (numOpsPlanned) *numOpsPlanned = planned;
    return 0;
}

// ------------------------------------------------------------------------------------------
// host emulation of k_tile_pass: the SAME round driver and gate bodies (reg_round, smem_pauli: compiled for host and
// device) run tile by tile on the descriptors emit_pass produced -- tile base enumeration, chunk offsets, per-tile
// control tests, per-tile star factors, rounds -- so that the whole host side of the tile engine (planner + emission)
// and the device gate bodies are checked against plain gate-by-gate application without a GPU.
// What stays GPU-only: the TMA / mbarrier pipeline that moves the tiles.
// ------------------------------------------------------------------------------------------
static void emulate_pass(std::vector<cplx>& amps, const Emitted& E, int hi) {
    const PassHdr& hdr = E.hdrs[hi];
    const RoundHdr* rounds = E.rounds.data() + E.roundBase[hi];
    const TileOp* ops = E.ops.data() + E.opBase[hi];
    const StarTab* tabs = E.tabs.data();
    std::vector<cplx> t(TILE_AMPS);
    cplx starF[MAX_OPS_PER_PASS];
    for (qindex k = 0; k < hdr.numTiles; k++) {
        const qindex base = hdr.tileIns(k);
        for (int c = 0; c < hdr.numChunks; c++)                                   // what the copy warp's bulk loads do
            for (int e = 0; e < hdr.chunkAmps; e++) t[(size_t)c * hdr.chunkAmps + e] = amps[base + hdr.chunkOff[c] + e];
        unsigned long long active = 0;
        for (int o = 0; o < hdr.numOps; o++) {
            const TileOp& op = ops[o];
            if (op.kind == OP_STAR || op.kind == OP_HSTAR) {
                const StarTab& tb = tabs[op.tab];
                cplx f = mk(1, 0);
                for (int sg = 0; sg < STAR_SEGS; sg++) f = cmul(f, tb.ext[sg][(base >> (6 * sg)) & 63]);
                starF[o] = f;
            }
            bool a = ((unsigned long long)base & op.extCtrlMask) == op.extCtrlVals;
            if (op.kind == OP_STAR && op.p0 < 0) a = a && getBit(base, op.e0);
            active |= (unsigned long long)a << o;
        }
        for (int r = 0; r < hdr.numRounds; r++) {
            const RoundHdr& rd = rounds[r];
            for (int wtid = 0; wtid < WG_THREADS; wtid++) {                       // the threads of a round touch disjoint amplitudes
                if (rd.kind == ROUND_REG) REG_ROUND(t.data(), rd, ops, base, active, tabs, starF, wtid REG_ROUND_EXTRA(nullptr));
                else if ((active >> rd.opBase) & 1) smem_pauli(t.data(), ops[rd.opBase], base, wtid);
            }
        }
        for (int c = 0; c < hdr.numChunks; c++)
            for (int e = 0; e < hdr.chunkAmps; e++) amps[base + hdr.chunkOff[c] + e] = t[(size_t)c * hdr.chunkAmps + e];
    }
}

extern "C" int qb_selftest_tile_emulation(int numQubits, int numOps, unsigned seed, int reorder, double* maxErr, int* numTilePasses, int* numDirectOps) {
    if (numQubits < TILE_BITS || numQubits > 20 || numOps < 1) return -1;
    std::mt19937_64 rng(seed);
    auto unif = [&]() { return (double)(rng() >> 11) * (1.0 / 9007199254740992.0); };
    const int n = numQubits;
    std::vector<QOp> ops;
    selftest_random_ops(n, numOps, rng, ops);
    qb_state q; memset(&q, 0, sizeof q); q.numAmpsPerNode = 1LL << n; q.logNumAmpsPerNode = n; q.numQubits = n;

    std::vector<hc> ref((size_t)1 << n);
    for (auto& v : ref) v = hc(2 * unif() - 1, 2 * unif() - 1);
    std::vector<cplx> state(ref.size());
    for (size_t i = 0; i < ref.size(); i++) state[i] = mk(ref[i].real(), ref[i].imag());
    for (const QOp& o : ops) host_apply(ref, o);

    std::vector<QOp> queue = ops, merged; std::vector<Pass> passes;
    plan_passes(queue, reorder != 0, merged, passes);
    int tilePasses = 0, directOps = 0;
    std::vector<hc> tmp;
    for (const Pass& p : passes) {
        const bool direct = (p.high == PASS_DIRECT) || (p.high == PASS_PGROUP) || (p.opIdx.size() == 1 && merged[p.opIdx[0]].kind != OP_STAR && merged[p.opIdx[0]].kind != OP_HSTAR);
        if (direct) {                                     // the product runs these through the direct / coset kernels (GPU tests)
            tmp.resize(state.size());
            for (size_t i = 0; i < state.size(); i++) tmp[i] = tohc(state[i]);
            for (int idx : p.opIdx) { host_apply(tmp, merged[idx]); directOps++; }
            for (size_t i = 0; i < state.size(); i++) state[i] = mk(tmp[i].real(), tmp[i].imag());
            continue;
        }
        Emitted E;
        emit_pass(&q, merged, p, E, reorder != 0);
        emulate_pass(state, E, 0);
        tilePasses++;
    }
    double err = 0, norm = 0;
    for (size_t i = 0; i < ref.size(); i++) { err = std::max(err, std::abs(ref[i] - tohc(state[i]))); norm = std::max(norm, std::abs(ref[i])); }
    if (maxErr) *maxErr = norm > 0 ? err / norm : err;
    if (numTilePasses) *numTilePasses = tilePasses;
    if (numDirectOps) *numDirectOps = directOps;
    return 0;
}

// ------------------------------------------------------------------------------------------
// host check of the RESTRICTED flush that the exchange / compute overlap is built on (qb_tile_flush_restricted,
// qb_p2p_swapHalvesOverlapped): a random gate list that never touches index bit `bit` is planned and emulated twice,
// once restricted to bit == 0 and once to bit == 1 -- the two halves of a shard that meet the gates at different
// times -- and the result must equal plain application of the list to the whole state.  Also checks that every
// restricted pass really skips the other half (numTiles halves) and that gate absorption and the phase-star merge,
// which extra per-gate controls would have disabled, still take place (numOpsPlanned < numOps).
// ------------------------------------------------------------------------------------------
extern "C" int qb_selftest_restricted_flush(int numQubits, int numOps, unsigned seed, int bit, double* maxErr, int* numTilePasses, int* numOpsPlanned) {
    if (numQubits < TILE_BITS + 1 || numQubits > 20 || numOps < 1 || bit < 0 || bit >= numQubits) return -1;
    if (bit < TILE_LOW) return -5;                 // same precondition as qb_tile_flush_restricted
    std::mt19937_64 rng(seed);
    auto unif = [&]() { return (double)(rng() >> 11) * (1.0 / 9007199254740992.0); };
    const int n = numQubits;
    // random ops on n-1 qubits, then re-labelled so that index bit `bit` is never involved
    std::vector<QOp> ops;
    selftest_random_ops(n - 1, numOps, rng, ops);
    auto spread = [&](unsigned long long m) { const unsigned long long lo = m & ((1ULL << bit) - 1); return ((m >> bit) << (bit + 1)) | lo; };
    auto qb = [&](int t) { return t >= bit ? t + 1 : t; };
    for (QOp& o : ops) {
        o.ctrlMask = spread(o.ctrlMask); o.ctrlVals = spread(o.ctrlVals); o.maskA = spread(o.maskA); o.maskB = spread(o.maskB);
        o.t0 = qb(o.t0); o.t1 = qb(o.t1);
        for (auto& ce : o.star) ce.first = qb(ce.first);
    }
    qb_state q; memset(&q, 0, sizeof q); q.numAmpsPerNode = 1LL << n; q.logNumAmpsPerNode = n; q.numQubits = n;
    std::vector<hc> ref((size_t)1 << n);
    for (auto& v : ref) v = hc(2 * unif() - 1, 2 * unif() - 1);
    std::vector<cplx> state(ref.size());
    for (size_t i = 0; i < ref.size(); i++) state[i] = mk(ref[i].real(), ref[i].imag());
    for (const QOp& o : ops) host_apply(ref, o);

    int tilePasses = 0, planned = 0, rc = 0;
    std::vector<hc> tmp;
    for (int half = 0; half < 2 && !rc; half++) {
        s_restrictMask = 1ULL << bit; s_restrictVals = half ? s_restrictMask : 0;
        std::vector<QOp> queue = ops, merged; std::vector<Pass> passes;
        plan_passes(queue, true, merged, passes);
        planned = (int)merged.size();
        for (const Pass& p : passes) {
            const bool direct = (p.high == PASS_DIRECT) || (p.high == PASS_PGROUP) || (p.opIdx.size() == 1 && merged[p.opIdx[0]].kind != OP_STAR && merged[p.opIdx[0]].kind != OP_HSTAR);
            if (direct) {          // the product adds the restriction as a control (run_direct) or as a fixed coset bit (coset kernel)
                tmp.resize(state.size());
                for (size_t i = 0; i < state.size(); i++) tmp[i] = tohc(state[i]);
                for (int idx : p.opIdx) {
                    QOp o = merged[idx];
                    if (o.kind == OP_STAR || o.kind == OP_HSTAR) {      // host_apply knows no controlled stars: split the halves by hand
                        std::vector<hc> full = tmp; host_apply(full, o);
                        for (size_t i = 0; i < tmp.size(); i++) if ((((unsigned long long)i >> bit) & 1ULL) == (unsigned long long)half) tmp[i] = full[i];
                    } else { o.ctrlMask |= s_restrictMask; o.ctrlVals |= s_restrictVals; host_apply(tmp, o); }
                }
                for (size_t i = 0; i < state.size(); i++) state[i] = mk(tmp[i].real(), tmp[i].imag());
                continue;
            }
            Emitted E;
            emit_pass(&q, merged, p, E, true);
            if (E.hdrs[0].numTiles != ((qindex)1 << (n - TILE_BITS - 1))) { rc = -4; break; }       // the other half must be pruned
            emulate_pass(state, E, 0);
            tilePasses++;
        }
    }
    s_restrictMask = s_restrictVals = 0;
    if (rc) return rc;
    double err = 0, norm = 0;
    for (size_t i = 0; i < ref.size(); i++) { err = std::max(err, std::abs(ref[i] - tohc(state[i]))); norm = std::max(norm, std::abs(ref[i])); }
    if (maxErr) *maxErr = norm > 0 ? err / norm : err;
    if (numTilePasses) *numTilePasses = tilePasses;
    if (numOpsPlanned) *numOpsPlanned = planned;
    return 0;
}
#endif  // QB_SELFTEST

#ifdef QB_TILE_TIMING
// ------------------------------------------------------------------------------------------
// compute-path probe (timing builds only): the round driver on a shared-memory tile, no TMA, no HBM -- what do the
// gate bodies + dispatch cost per gate with one or two warps per scheduler?  tools/tile_round_probe.py
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TILE_THREADS, 1) k_round_probe(const PassHdr* __restrict__ hdrp, const RoundHdr* __restrict__ grounds,
        const TileOp* __restrict__ gops, const StarTab* __restrict__ tabs, int reps, double* sink) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cplx* stageBuf = reinterpret_cast<cplx*>(smem_raw);
    TileOp* ops = reinterpret_cast<TileOp*>(smem_raw + (size_t)2 * TILE_AMPS * sizeof(cplx));
    __shared__ RoundHdr rounds[MAX_OPS_PER_PASS];
    __shared__ cplx starF[MAX_OPS_PER_PASS];
    const int tid = threadIdx.x, numOps = hdrp->numOps, numRounds = hdrp->numRounds;
    for (int i = tid; i < numOps * (int)(sizeof(TileOp) / 4); i += blockDim.x) ((int*)ops)[i] = ((const int*)gops)[i];
    for (int i = tid; i < numRounds * (int)(sizeof(RoundHdr) / 4); i += blockDim.x) ((int*)rounds)[i] = ((const int*)grounds)[i];
    for (int i = tid; i < 2 * TILE_AMPS; i += blockDim.x) stageBuf[i] = mk(1e-3 * (i & 255), 2e-3 * (i & 127));
    if (tid < MAX_OPS_PER_PASS) starF[tid] = mk(1, 0);
    __syncthreads();
    const int wg = tid / WG_THREADS, wtid = tid % WG_THREADS;
    cplx* t = stageBuf + (size_t)wg * TILE_AMPS;
    TT_DECL;
    for (int rep = 0; rep < reps; rep++)
        for (int r = 0; r < numRounds; r++) {
            REG_ROUND(t, rounds[r], ops, 0, ~0ULL, tabs, starF, wtid REG_ROUND_EXTRA(nullptr) TT_PASS);
            wg_sync(wg);
        }
    if (wtid == 0) sink[blockIdx.x * 2 + wg] = t[blockIdx.x & 1023].x;
    TT_FLUSH;
}

// kind: 1 = dense 1-qubit gates, 2 = dense 2-qubit gates; all on the four highest qubits (one register round)
extern "C" int qb_tile_round_probe(int kind, int numGates, int numWG, int reps, float* msOut, int* roundsOut) {
    QB_READY();
    qb_state q; memset(&q, 0, sizeof q); q.numAmpsPerNode = 1LL << 30; q.logNumAmpsPerNode = 30; q.numQubits = 30;
    std::vector<QOp> ops;
    for (int g = 0; g < numGates; g++) {
        QOp o = blank(kind == 1 ? OP_DENSE1 : OP_DENSE2, &q, 0);
        if (kind == 1) { o.t0 = 26 + (g & 3); }
        else { o.numT = 2; o.t0 = (g & 1) ? 28 : 26; o.t1 = o.t0 + 1; }
        for (int i = 0; i < 16; i++) o.m[i] = mk(0.25 + 0.01 * i, 0.1 - 0.02 * i);
        ops.push_back(o);
    }
    std::vector<QOp> merged; std::vector<Pass> passes;
    plan_passes(ops, false, merged, passes);
    if (passes.size() != 1) return qb_set_error(-1, "round probe: expected one pass", __FILE__, __LINE__);
    Emitted E; emit_pass(&q, merged, passes[0], E, false);
    if (roundsOut) *roundsOut = E.hdrs[0].numRounds;
    char* dev; const size_t bh = sizeof(PassHdr), br = E.rounds.size() * sizeof(RoundHdr), bo = E.ops.size() * sizeof(TileOp);
    QB_CUDA(cudaMalloc(&dev, bh + br + bo + sizeof(StarTab) + 4096));
    QB_CUDA(cudaMemcpy(dev, E.hdrs.data(), bh, cudaMemcpyHostToDevice));
    QB_CUDA(cudaMemcpy(dev + bh, E.rounds.data(), br, cudaMemcpyHostToDevice));
    QB_CUDA(cudaMemcpy(dev + bh + br, E.ops.data(), bo, cudaMemcpyHostToDevice));
    double* sink = (double*)(dev + bh + br + bo + sizeof(StarTab));
    const size_t smemBytes = (size_t)2 * TILE_AMPS * sizeof(cplx) + (size_t)MAX_OPS_PER_PASS * sizeof(TileOp);
    QB_CUDA(cudaFuncSetAttribute(k_round_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 2; w++) {
        cudaEventRecord(e0, g_qb.stream);
        k_round_probe<<<g_qb.numSMs, WG_THREADS * numWG, smemBytes, g_qb.stream>>>((const PassHdr*)dev, (const RoundHdr*)(dev + bh), (const TileOp*)(dev + bh + br),
                                                                                   (const StarTab*)(dev + bh + br + bo), reps, sink);
        cudaEventRecord(e1, g_qb.stream);
        QB_CUDA(cudaEventSynchronize(e1));
    }
    cudaEventElapsedTime(msOut, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(dev);
    return 0;
}
#endif
