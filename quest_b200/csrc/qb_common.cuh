// qb_common.cuh -- shared device/host helpers of the B200 amplitude-update backend.
//
// Index algebra restates quest/src/core/bitwise.hpp (insertBit :99-105, insertBits :164-171,
// insertBitsWithMaskedValues :206-210, setBits :174-183, getValueOfBits :186-195, parity :238-255)
// for device code; all of it is 64-bit integer arithmetic and must be bit-exact.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/quest_b200.h"

// amplitude type: follows the library's build-time precision (include/quest_b200.h, QuEST's FLOAT_PRECISION)
#if QB_PRECISION == 1
typedef float2 cplx;
typedef float real;
#else
typedef double2 cplx;
typedef double real;
#endif
static_assert(sizeof(cplx) == sizeof(qb_cplx), "cplx must be layout-identical to qb_cplx / qcomp");
typedef long long qindex;

#define QB_MAX_QUBITS 63
#define QB_BLOCK 256

// ------------------------------------------------------------------------------------------
// host-side runtime state (qb_runtime.cu)
// ------------------------------------------------------------------------------------------
struct QbRuntime {
    int device = -1;             // bound device, -1 before qb_bind_device
    int numSMs = 148;
    cudaStream_t stream = 0;     // all compute goes here (default: legacy stream 0)
    unsigned long long launches = 0;
    int tileEngine = 1;          // 0: direct kernels only; 1: fused passes with gate absorption + commuting re-order; 2: fused, program order
    // reduction scratch (device) + pinned host landing zone
    double* redPartials = nullptr;   // [QB_RED_MAX_BLOCKS * 2 * QB_RED_MAX_OUT]
    unsigned int* redTicket = nullptr;
    double* redOutDev = nullptr;
    double* redOutHost = nullptr;    // pinned
    // user-visible cache (gpu_getCacheOfSize)
    cplx* cache = nullptr;
    qindex cacheLen = 0;
};
extern QbRuntime g_qb;

int  qb_set_error(int code, const char* what, const char* file, int line);
void qb_p2p_note_alloc(void* base, size_t bytes);   // qb_p2p.cu: allocation registry for IPC export
int  qb_p2p_note_free(void* base);
int  qb_ensure_ready();   // binds device 0 lazily, allocates scratch; returns 0 or error

#define QB_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) \
    return qb_set_error((int)e__, #call, __FILE__, __LINE__); } while (0)
int  qb_flush_internal();  // qb_tile.cu: run every deferred (queued) gate now
void qb_tile_forget(const void* amps);   // qb_tile.cu: drop the (already flushed) queue of a state that is being freed
// every entry point first makes the device ready and drains the deferred-gate queue; the fusable gate entry points
// use QB_READY_NOFLUSH and either append to the queue or flush explicitly before running a direct kernel
#define QB_READY_NOFLUSH() do { int r__ = qb_ensure_ready(); if (r__) return r__; } while (0)
#define QB_FLUSH() do { int r__ = qb_flush_internal(); if (r__) return r__; } while (0)
#define QB_READY() do { QB_READY_NOFLUSH(); QB_FLUSH(); } while (0)
#define QB_LAUNCH_CHECK() do { g_qb.launches++; cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) \
    return qb_set_error((int)e__, "kernel launch", __FILE__, __LINE__); } while (0)
#define QB_REQUIRE(cond, msg) do { if (!(cond)) return qb_set_error(-1, msg, __FILE__, __LINE__); } while (0)

// ------------------------------------------------------------------------------------------
// complex arithmetic on cplx (double2 / float2).  Scalar helpers take `double`: in the fp32 build mixed expressions
// are evaluated in double and narrowed on assignment, which only ever adds accuracy
// ------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ cplx mk(double re, double im) { cplx c; c.x = (real)re; c.y = (real)im; return c; }
__host__ __device__ __forceinline__ cplx mk(qb_cplx c) { return mk(c.re, c.im); }
__host__ __device__ __forceinline__ cplx cadd(cplx a, cplx b) { return mk(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ cplx csub(cplx a, cplx b) { return mk(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ cplx cmul(cplx a, cplx b) { return mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__host__ __device__ __forceinline__ cplx cscale(double s, cplx a) { return mk((real)s * a.x, (real)s * a.y); }
__host__ __device__ __forceinline__ cplx cconj(cplx a) { return mk(a.x, -a.y); }
__host__ __device__ __forceinline__ double cnorm(cplx a) { return (double)a.x * a.x + (double)a.y * a.y; }
// acc + a*b
__host__ __device__ __forceinline__ cplx cfma(cplx a, cplx b, cplx acc) {
    return mk(acc.x + a.x * b.x - a.y * b.y, acc.y + a.x * b.y + a.y * b.x);
}

// complex power a^b = exp(b * log a), the definition std::pow(complex, complex) uses
// (reference device version: quest/src/gpu/gpu_types.cuh:248-270)
__device__ __forceinline__ cplx cpow(cplx a, cplx b) {
    double r = hypot((double)a.x, (double)a.y);
    if (r == 0.0) {
        // 0^0 = 1, 0^b = 0 (matches std::pow for positive real exponents)
        return (b.x == 0.0 && b.y == 0.0) ? mk(1.0, 0.0) : mk(0.0, 0.0);
    }
    double lr = log(r), th = atan2((double)a.y, (double)a.x);
    double ere = b.x * lr - b.y * th;       // Re(b*log a)
    double eim = b.x * th + b.y * lr;       // Im(b*log a)
    double m = exp(ere), s, c;
    sincos(eim, &s, &c);
    return mk(m * c, m * s);
}

// ------------------------------------------------------------------------------------------
// bit algebra
// ------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ qindex pow2(int n) { return (qindex)1 << n; }
__host__ __device__ __forceinline__ int getBit(qindex v, int q) { return (int)((v >> q) & 1); }
__host__ __device__ __forceinline__ qindex insertZeroBit(qindex v, int q) {
    qindex low = v & (pow2(q) - 1);
    return ((v >> q) << (q + 1)) | low;
}
__host__ __device__ __forceinline__ int parity64(unsigned long long v) {
#ifdef __CUDA_ARCH__
    return __popcll(v) & 1;
#else
    return __builtin_popcountll(v) & 1;
#endif
}

// insertBitsWithMaskedValues(n, sortedQubits, numQubits, mask): zero bits are inserted at the given
// (increasing) positions, then the value mask is OR-ed in.  Up to four positions are inserted one by
// one from scalar members (the common case: a target plus a few controls); longer lists use the
// branch-free "expand" (bit deposit) of Hacker's Delight 7-5 with six precomputed move masks, so the
// kernels never index an array held in kernel parameters (which would force a local-memory copy).
struct BitIns {
    int n;
    int p0, p1, p2, p3;                 // first four positions (valid when n <= 4)
    unsigned long long keep;            // 1 = bit position that receives a bit of the item number
    unsigned long long mv[6];           // expand move masks (valid when n > 4)
    unsigned long long mask;            // values OR-ed into the inserted positions
    __host__ __device__ __forceinline__ qindex operator()(qindex item) const {
        unsigned long long v = (unsigned long long)item;
        if (n <= 4) {
            if (n >= 1) v = (unsigned long long)insertZeroBit((qindex)v, p0);
            if (n >= 2) v = (unsigned long long)insertZeroBit((qindex)v, p1);
            if (n >= 3) v = (unsigned long long)insertZeroBit((qindex)v, p2);
            if (n >= 4) v = (unsigned long long)insertZeroBit((qindex)v, p3);
        } else {
#pragma unroll
            for (int i = 5; i >= 0; i--) {
                unsigned long long t = v << (1 << i);
                v = (v & ~mv[i]) | (t & mv[i]);
            }
            v &= keep;
        }
        return (qindex)(v | mask);
    }
};

// qubit list in USER order: bit j of a value <-> position q[j]  (setBits / getValueOfBits)
struct BitList {
    int n;
    unsigned char q[QB_MAX_QUBITS];
    __host__ __device__ __forceinline__ qindex gather(qindex idx) const {   // getValueOfBits
        qindex v = 0;
        for (int j = 0; j < n; j++) v |= (qindex)getBit(idx, q[j]) << j;
        return v;
    }
    __host__ __device__ __forceinline__ qindex scatter(qindex v) const {    // bits of v placed at q[j]
        qindex m = 0;
        for (int j = 0; j < n; j++) m |= (qindex)((v >> j) & 1) << q[j];
        return m;
    }
};

// host helpers (qb_runtime.cu)
BitIns   qb_make_ins(const int* a, const int* aStates, int na, const int* b, const int* bStates, int nb);
BitList  qb_make_list(const int* q, int n);
unsigned long long qb_make_mask(const int* q, int n);
int      qb_check_qubits(const int* q, int n, int limit);

// ------------------------------------------------------------------------------------------
// launch geometry: numItems is always a power of two in this code base.
// ------------------------------------------------------------------------------------------
static inline unsigned int qb_grid(qindex numItems, int itemsPerThread) {
    qindex per = (qindex)QB_BLOCK * itemsPerThread;
    qindex g = (numItems + per - 1) / per;
    return (unsigned int)(g < 1 ? 1 : g);
}

// streaming 128-bit accesses. Amplitudes are touched exactly once per pass, so keep them out of L1.
__device__ __forceinline__ cplx ld_stream(const cplx* p) {
    cplx r;
#if QB_PRECISION == 1
    asm volatile("ld.global.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
#else
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
#endif
    return r;
}
__device__ __forceinline__ void st_stream(cplx* p, cplx v) {
#if QB_PRECISION == 1
    asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1, %2};" :: "l"(p), "f"(v.x), "f"(v.y) : "memory");
#else
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" :: "l"(p), "d"(v.x), "d"(v.y) : "memory");
#endif
}
