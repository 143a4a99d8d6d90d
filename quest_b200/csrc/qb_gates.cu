// qb_gates.cu -- statevector gate kernels (direct path) behind the C ABI.
//
// One entry point per reference routine of quest/src/gpu/gpu_subroutines.cpp; the arithmetic follows
// the reference's CPU implementation (quest/src/cpu/cpu_subroutines.cpp, the parity oracle) line for
// line in meaning, but the enumeration of work items is re-derived for coalescing:
//   * every kernel enumerates item n -> base index via insertBitsWithMaskedValues (bitwise.hpp:206),
//   * Pauli pairs are enumerated by inserting ONE zero bit at the highest X/Y target instead of the
//     reference's (outer n, inner v) split (cpu_subroutines.cpp:832-905): the pairs {i, i^maskXY} are
//     the same set and the update of each pair is identical, but no per-target loop is needed.
#include "qb_common.cuh"
#include "qb_kernels.cuh"
#include "qb_tile.cuh"

// ------------------------------------------------------------------------------------------
// tuple ops
// ------------------------------------------------------------------------------------------
struct OpDense1 {                       // cpu_subroutines.cpp:362-391
    static constexpr int M = 2;
    BitIns ins; qindex tbit; cplx m00, m01, m10, m11;
    __device__ __forceinline__ void indices(qindex n, qindex* idx) const { idx[0] = ins(n); idx[1] = idx[0] | tbit; }
    __device__ __forceinline__ void apply(const qindex*, cplx* v) const {
        cplx a0 = v[0], a1 = v[1];
        v[0] = cfma(m01, a1, cmul(m00, a0));
        v[1] = cfma(m11, a1, cmul(m10, a0));
    }
    __device__ __forceinline__ bool writes(int) const { return true; }
};

struct OpDense2 {                       // cpu_subroutines.cpp:435-472
    static constexpr int M = 4;
    BitIns ins; qindex b1, b2; cplx m[16];
    __device__ __forceinline__ void indices(qindex n, qindex* idx) const {
        idx[0] = ins(n); idx[1] = idx[0] | b1; idx[2] = idx[0] | b2; idx[3] = idx[1] | b2;
    }
    __device__ __forceinline__ void apply(const qindex*, cplx* v) const {
        cplx a0 = v[0], a1 = v[1], a2 = v[2], a3 = v[3];
#pragma unroll
        for (int r = 0; r < 4; r++)
            v[r] = cfma(m[4 * r + 3], a3, cfma(m[4 * r + 2], a2, cfma(m[4 * r + 1], a1, cmul(m[4 * r], a0))));
    }
    __device__ __forceinline__ bool writes(int) const { return true; }
};

struct OpSwapA {                        // cpu_subroutines.cpp:259-283
    static constexpr int M = 2;
    BitIns ins; qindex flip;
    __device__ __forceinline__ void indices(qindex n, qindex* idx) const { idx[0] = ins(n); idx[1] = idx[0] ^ flip; }
    __device__ __forceinline__ void apply(const qindex*, cplx* v) const { cplx t = v[0]; v[0] = v[1]; v[1] = t; }
    __device__ __forceinline__ bool writes(int) const { return true; }
};

struct OpPauliA {                       // cpu_subroutines.cpp:804-905
    static constexpr int M = 2;
    BitIns ins; qindex maskXY, maskYZ; cplx ampFac, pairFac;   // pairFac already includes i^numY
    __device__ __forceinline__ void indices(qindex n, qindex* idx) const { idx[0] = ins(n); idx[1] = idx[0] ^ maskXY; }
    __device__ __forceinline__ void apply(const qindex* idx, cplx* v) const {
        double sA = 1.0 - 2.0 * parity64((unsigned long long)(idx[0] & maskYZ));
        double sB = 1.0 - 2.0 * parity64((unsigned long long)(idx[1] & maskYZ));
        cplx a = v[0], b = v[1];
        v[0] = cfma(pairFac, cscale(sB, b), cmul(ampFac, a));
        v[1] = cfma(pairFac, cscale(sA, a), cmul(ampFac, b));
    }
    __device__ __forceinline__ bool writes(int) const { return true; }
};

// ------------------------------------------------------------------------------------------
// map ops
// ------------------------------------------------------------------------------------------
struct OpDiag1 {                        // cpu_subroutines.cpp:582-610 (targ may be a prefix qubit)
    static constexpr bool READS = true;
    BitIns ins; qindex rankBits; int targ; cplx e0, e1;
    __device__ __forceinline__ qindex index(qindex n) const { return ins(n); }
    __device__ __forceinline__ cplx second(qindex, qindex) const { return mk(0, 0); }
    __device__ __forceinline__ cplx apply(qindex, qindex j, cplx a, cplx) const {
        qindex i = rankBits | j;
        return cmul(a, getBit(i, targ) ? e1 : e0);
    }
};

struct OpDiag2 {                        // cpu_subroutines.cpp:620-648
    static constexpr bool READS = true;
    BitIns ins; qindex rankBits; int t1, t2; cplx e[4];
    __device__ __forceinline__ qindex index(qindex n) const { return ins(n); }
    __device__ __forceinline__ cplx second(qindex, qindex) const { return mk(0, 0); }
    __device__ __forceinline__ cplx apply(qindex, qindex j, cplx a, cplx) const {
        qindex i = rankBits | j;
        int k = (getBit(i, t2) << 1) | getBit(i, t1);
        return cmul(a, e[k]);
    }
};

struct OpDiagK {                        // cpu_subroutines.cpp:658-706
    static constexpr bool READS = true;
    BitIns ins; qindex rankBits; BitList targs; const cplx* elems; int conj, hasPower; cplx expo;
    __device__ __forceinline__ qindex index(qindex n) const { return ins(n); }
    __device__ __forceinline__ cplx second(qindex, qindex j) const {
        qindex t = targs.gather(rankBits | j);
        cplx e = __ldg(&elems[t]);
        if (hasPower) e = cpow(e, expo);
        if (conj) e.y = -e.y;
        return e;
    }
    __device__ __forceinline__ cplx apply(qindex, qindex, cplx a, cplx e) const { return cmul(a, e); }
};

struct OpAllDiagSV {                    // cpu_subroutines.cpp:714-737
    static constexpr bool READS = true;
    const cplx* elems; int hasPower; cplx expo;
    __device__ __forceinline__ qindex index(qindex n) const { return n; }
    __device__ __forceinline__ cplx second(qindex n, qindex) const {
        cplx e = elems[n];
        if (hasPower) e = cpow(e, expo);
        return e;
    }
    __device__ __forceinline__ cplx apply(qindex, qindex, cplx a, cplx e) const { return cmul(a, e); }
};

struct OpAllDiagDM {                    // cpu_subroutines.cpp:741-787
    static constexpr bool READS = true;
    const cplx* elems; qindex dim; qindex rankBits; int hasPower, mulOnly; cplx expo;
    __device__ __forceinline__ qindex index(qindex n) const { return n; }
    __device__ __forceinline__ cplx second(qindex n, qindex) const {
        cplx fac = __ldg(&elems[n % dim]);
        if (hasPower) fac = cpow(fac, expo);
        if (!mulOnly) {
            qindex m = rankBits | n;
            cplx term = __ldg(&elems[m / dim]);
            if (hasPower) term = cpow(term, expo);
            fac = cmul(fac, cconj(term));
        }
        return fac;
    }
    __device__ __forceinline__ cplx apply(qindex, qindex, cplx a, cplx e) const { return cmul(a, e); }
};

struct OpPhaseGadget {                  // cpu_subroutines.cpp:963-993
    static constexpr bool READS = true;
    BitIns ins; qindex targMask; cplx f0, f1;
    __device__ __forceinline__ qindex index(qindex n) const { return ins(n); }
    __device__ __forceinline__ cplx second(qindex, qindex) const { return mk(0, 0); }
    __device__ __forceinline__ cplx apply(qindex, qindex i, cplx a, cplx) const {
        return cmul(a, parity64((unsigned long long)(i & targMask)) ? f1 : f0);
    }
};

struct OpDense1B {                      // cpu_subroutines.cpp:394-423
    static constexpr bool READS = true;
    BitIns ins; const cplx* buf; cplx f0, f1;
    __device__ __forceinline__ qindex index(qindex n) const { return ins(n); }
    __device__ __forceinline__ cplx second(qindex n, qindex) const { return buf[n]; }
    __device__ __forceinline__ cplx apply(qindex, qindex, cplx a, cplx b) const { return cfma(f1, b, cmul(f0, a)); }
};

struct OpPauliB {                       // cpu_subroutines.cpp:910-950
    static constexpr bool READS = true;
    BitIns ins; const cplx* buf; qindex maskXY, maskYZ, bufMaskXY; cplx ampFac, pairFac;
    __device__ __forceinline__ qindex index(qindex n) const { return ins(n); }
    __device__ __forceinline__ cplx second(qindex n, qindex) const { return buf[n ^ bufMaskXY]; }
    __device__ __forceinline__ cplx apply(qindex, qindex i, cplx a, cplx b) const {
        qindex k = i ^ maskXY;
        double s = 1.0 - 2.0 * parity64((unsigned long long)(k & maskYZ));
        return cfma(pairFac, cscale(s, b), cmul(ampFac, a));
    }
};

struct OpUnpack {                       // swap subB / subC: cpu_subroutines.cpp:286-349
    static constexpr bool READS = false;
    BitIns ins; const cplx* buf;
    __device__ __forceinline__ qindex index(qindex n) const { return ins(n); }
    __device__ __forceinline__ cplx second(qindex n, qindex) const { return buf[n]; }
    __device__ __forceinline__ cplx apply(qindex, qindex, cplx, cplx b) const { return b; }
};

// packing writes the buffer, reading amps: run k_map over the *buffer* as the destination
struct OpPack {                         // cpu_subroutines.cpp:192-220
    static constexpr bool READS = false;
    BitIns ins; const cplx* amps;
    __device__ __forceinline__ qindex index(qindex n) const { return n; }
    __device__ __forceinline__ cplx second(qindex n, qindex) const { return amps[ins(n)]; }
    __device__ __forceinline__ cplx apply(qindex, qindex, cplx, cplx b) const { return b; }
};

struct OpPackPairSum {                  // cpu_subroutines.cpp:223-249
    static constexpr bool READS = false;
    BitIns ins; const cplx* amps; qindex flip;
    __device__ __forceinline__ qindex index(qindex n) const { return n; }
    __device__ __forceinline__ cplx second(qindex n, qindex) const {
        qindex i0b0 = ins(n);
        return cadd(amps[i0b0], amps[i0b0 ^ flip]);
    }
    __device__ __forceinline__ cplx apply(qindex, qindex, cplx, cplx b) const { return b; }
};

struct OpSuperpose {                    // cpu_subroutines.cpp:1002-1014
    static constexpr bool READS = true;
    const cplx* in1; const cplx* in2; cplx fOut, f1, f2;
    __device__ __forceinline__ qindex index(qindex n) const { return n; }
    __device__ __forceinline__ cplx second(qindex n, qindex) const { return cfma(f2, in2[n], cmul(f1, in1[n])); }
    __device__ __forceinline__ cplx apply(qindex, qindex, cplx a, cplx b) const { return cadd(cmul(fOut, a), b); }
};

struct OpProjSV {                       // cpu_subroutines.cpp:2282-2307
    static constexpr bool READS = true;
    qindex qubitMask, retainMask; double renorm;
    __device__ __forceinline__ qindex index(qindex n) const { return n; }
    __device__ __forceinline__ cplx second(qindex, qindex) const { return mk(0, 0); }
    __device__ __forceinline__ cplx apply(qindex n, qindex, cplx a, cplx) const {
        double fac = ((n & qubitMask) == retainMask) ? renorm : 0.0;
        return cscale(fac, a);
    }
};

struct OpFill {                         // cpu_subroutines.cpp:2356-2361
    static constexpr bool READS = false;
    cplx amp;
    __device__ __forceinline__ qindex index(qindex n) const { return n; }
    __device__ __forceinline__ cplx second(qindex, qindex) const { return mk(0, 0); }
    __device__ __forceinline__ cplx apply(qindex, qindex, cplx, cplx) const { return amp; }
};

struct OpDebug {                        // cpu_subroutines.cpp:2364-2376
    static constexpr bool READS = false;
    qindex rankBits;
    __device__ __forceinline__ qindex index(qindex n) const { return n; }
    __device__ __forceinline__ cplx second(qindex, qindex) const { return mk(0, 0); }
    __device__ __forceinline__ cplx apply(qindex n, qindex, cplx, cplx) const {
        qindex i = rankBits | n;
        return mk((double)(2 * i) / 10., (double)(2 * i + 1) / 10.);
    }
};

// counter-based RNG (SplitMix64 per variate); stream differs from the CPU's by design, as the
// reference's own GPU path does (gpu_thrust.cuh:581-634); the distribution is the same:
// |amp|^2 ~ chi-squared(2) (sum of two unit normals squared), phase ~ U[0, 2pi).
struct OpRandom {
    static constexpr bool READS = false;
    unsigned long long seed; qindex rankBits;
    __device__ __forceinline__ static unsigned long long mix(unsigned long long z) {
        z += 0x9e3779b97f4a7c15ULL;
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
        return z ^ (z >> 31);
    }
    __device__ __forceinline__ static double u01(unsigned long long r) { return ((r >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
    __device__ __forceinline__ qindex index(qindex n) const { return n; }
    __device__ __forceinline__ cplx second(qindex, qindex) const { return mk(0, 0); }
    __device__ __forceinline__ cplx apply(qindex n, qindex, cplx, cplx) const {
        unsigned long long k = mix(seed ^ mix((unsigned long long)(rankBits | n)));
        double u1 = u01(mix(k)), u2 = u01(mix(k + 1));
        // n1^2 + n2^2 for two unit normals == -2 ln(u1) (Box-Muller radius squared)
        double prob = -2.0 * log(u1);
        double s, c;
        sincospi(2.0 * u2, &s, &c);
        double r = sqrt(prob);
        return mk(r * c, r * s);
    }
};

// ------------------------------------------------------------------------------------------
// k-target dense matrix, 3 <= k <= 5: the 2^k amplitudes of a tuple live in registers; the matrix is
// read through the read-only path (every thread of a warp reads the same element -> broadcast).
// cpu_subroutines.cpp:483-572
// ------------------------------------------------------------------------------------------
template <int K, bool CONJ>
__global__ void __launch_bounds__(128) k_denseK_reg(cplx* __restrict__ amps, qindex numItems, BitIns ins,
                                                    BitList targs, const cplx* __restrict__ matr) {
    constexpr int D = 1 << K;
    // the matrix is staged in shared memory once per block (conjugated there if asked): every thread then reads each
    // element with a warp-uniform (broadcast) shared-memory load instead of a global one -- at K = 4 (2-qubit Kraus
    // superoperators, BASELINE cfg 4) the 256 global loads per thread were what kept the kernel at half its roofline
    __shared__ cplx sm[D * D];
    for (int i = threadIdx.x; i < D * D; i += 128) { cplx e = __ldg(&matr[i]); if (CONJ) e.y = -e.y; sm[i] = e; }
    __syncthreads();
    const qindex n = (qindex)blockIdx.x * 128 + threadIdx.x;
    if (n >= numItems) return;
    const qindex i0 = ins(n);
    cplx v[D];
#pragma unroll
    for (int j = 0; j < D; j++) {
        qindex o = 0;
#pragma unroll
        for (int b = 0; b < K; b++) o |= (qindex)((j >> b) & 1) << targs.q[b];
        v[j] = amps[i0 | o];
    }
    // rows are produced two at a time straight from the register copy, so no second buffer is needed
#pragma unroll 1
    for (int r = 0; r < D; r += (D >= 2 ? 2 : 1)) {
        cplx acc0 = mk(0, 0), acc1 = mk(0, 0);
#pragma unroll
        for (int c = 0; c < D; c++) {
            acc0 = cfma(sm[r * D + c], v[c], acc0);
            if (D >= 2) acc1 = cfma(sm[(r + 1) * D + c], v[c], acc1);
        }
        qindex o0 = 0, o1 = 0;
#pragma unroll
        for (int b = 0; b < K; b++) { o0 |= (qindex)((r >> b) & 1) << targs.q[b]; o1 |= (qindex)(((r + 1) >> b) & 1) << targs.q[b]; }
        amps[i0 | o0] = acc0;
        if (D >= 2) amps[i0 | o1] = acc1;
    }
}

// k-target dense matrix, k >= 6: a block stages TUP tuples of 2^k amplitudes in shared memory
// (layout [amp j][tuple]), every warp then owns rows r = warp, warp+W, ... for 32 tuples at a time;
// the matrix element is warp-uniform (broadcast load), the amplitude load is conflict-free.
template <bool CONJ>
__global__ void __launch_bounds__(256) k_denseK_smem(cplx* __restrict__ amps, qindex numItems, BitIns ins,
                                                     BitList targs, const cplx* __restrict__ matr, int K, int tupLog) {
    extern __shared__ cplx sm[];                       // [D][TUP]
    const int D = 1 << K, TUP = 1 << tupLog;
    const qindex firstItem = (qindex)blockIdx.x * TUP;
    // gather
    for (int e = threadIdx.x; e < D * TUP; e += blockDim.x) {
        int t = e & (TUP - 1), j = e >> tupLog;
        qindex n = firstItem + t;
        if (n < numItems) sm[j * TUP + t] = amps[ins(n) | targs.scatter(j)];
    }
    __syncthreads();
    // multiply: thread handles (tuple t, row r)
    for (int e = threadIdx.x; e < D * TUP; e += blockDim.x) {
        int t = e & (TUP - 1), r = e >> tupLog;
        qindex n = firstItem + t;
        if (n >= numItems) continue;
        cplx acc = mk(0, 0);
        const cplx* row = matr + (qindex)r * D;
        for (int c = 0; c < D; c++) {
            cplx m = __ldg(&row[c]);
            if (CONJ) m.y = -m.y;
            acc = cfma(m, sm[c * TUP + t], acc);
        }
        amps[ins(n) | targs.scatter(r)] = acc;
    }
}

static int qb_denseK(const qb_state* q, const BitIns& ins, const BitList& targs, qindex numItems,
                     const cplx* matr, int conj) {
    cplx* amps = (cplx*)q->amps;
    const int K = targs.n;
#define QB_DK(KK) do { unsigned int g = (unsigned int)((numItems + 127) / 128); \
        if (conj) k_denseK_reg<KK, true><<<g, 128, 0, g_qb.stream>>>(amps, numItems, ins, targs, matr); \
        else      k_denseK_reg<KK, false><<<g, 128, 0, g_qb.stream>>>(amps, numItems, ins, targs, matr); } while (0)
    if (K == 1) QB_DK(1);
    else if (K == 2) QB_DK(2);
    else if (K == 3) QB_DK(3);
    else if (K == 4) QB_DK(4);
    else if (K == 5) QB_DK(5);
    else {
        // shared tile of D*TUP amplitudes <= 64 KiB (4096 amps); at least one tuple per block
        QB_REQUIRE(K <= 12, "dense matrices on more than 12 targets exceed this backend's shared-memory tile");
        int tupLog = 12 - K;
        if (tupLog > 5) tupLog = 5;
        while (tupLog > 0 && pow2(tupLog) > numItems) tupLog--;
        size_t smem = sizeof(cplx) << (K + tupLog);
        unsigned int g = (unsigned int)((numItems + pow2(tupLog) - 1) >> tupLog);
        if (conj) {
            QB_CUDA(cudaFuncSetAttribute(k_denseK_smem<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
            k_denseK_smem<true><<<g, 256, smem, g_qb.stream>>>(amps, numItems, ins, targs, matr, K, tupLog);
        } else {
            QB_CUDA(cudaFuncSetAttribute(k_denseK_smem<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
            k_denseK_smem<false><<<g, 256, smem, g_qb.stream>>>(amps, numItems, ins, targs, matr, K, tupLog);
        }
    }
#undef QB_DK
    QB_LAUNCH_CHECK();
    return 0;
}

// direct Pauli pair kernel from raw masks; pairFac already carries i^numY
int qb_pauli_raw(const qb_state* q, const int* ctrls, const int* cs, int nc, unsigned long long maskXY, unsigned long long maskYZ, qb_cplx ampFac, qb_cplx pairFac) {
    int high = 63 - __builtin_clzll(maskXY), zero = 0;
    OpPauliA op; op.ins = qb_make_ins(ctrls, cs, nc, &high, &zero, 1);
    op.maskXY = (qindex)maskXY; op.maskYZ = (qindex)maskYZ; op.ampFac = mk(ampFac); op.pairFac = mk(pairFac);
    return qb_launch_tuple((cplx*)q->amps, q->numAmpsPerNode >> (1 + nc), op);
}

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" {

int qb_statevec_packAmpsIntoBuffer(const qb_state* q, const int* qubits, const int* states, int nq, qb_index* numPacked) {
    QB_READY(); QB_CHECK_STATE(q); QB_CHECK_SUFFIX(qubits, nq, q);
    QB_REQUIRE(q->buffer, "packAmpsIntoBuffer: state has no communication buffer");
    qindex numIts = q->numAmpsPerNode >> nq;
    OpPack op; op.ins = qb_make_ins(qubits, states, nq, nullptr, nullptr, 0); op.amps = (const cplx*)q->amps;
    cplx* dst = (cplx*)q->buffer + q->numAmpsPerNode / 2;      // comm_indices.hpp:21-28
    if (numPacked) *numPacked = numIts;
    return qb_launch_map(dst, numIts, op);
}

int qb_statevec_packPairSummedAmpsIntoBuffer(const qb_state* q, int q1, int q2, int q3, int bit2, qb_index* numPacked) {
    QB_READY(); QB_CHECK_STATE(q);
    QB_REQUIRE(q->buffer, "packPairSummedAmpsIntoBuffer: state has no communication buffer");
    QB_REQUIRE(q1 < q2 && q2 < q3 && q1 >= 0 && q3 < q->logNumAmpsPerNode, "packPairSummed: qubits must be increasing and local");
    qindex numIts = q->numAmpsPerNode / 8;
    int qs[3] = {q1, q2, q3}, st[3] = {0, bit2, 0};
    OpPackPairSum op; op.ins = qb_make_ins(qs, st, 3, nullptr, nullptr, 0); op.amps = (const cplx*)q->amps;
    op.flip = pow2(q1) | pow2(q3);
    cplx* dst = (cplx*)q->buffer + q->numAmpsPerNode / 2;
    if (numPacked) *numPacked = numIts;
    return qb_launch_map(dst, numIts, op);
}

int qb_statevec_anyCtrlSwap_subA(const qb_state* q, const int* ctrls, const int* cs, int nc, int t1, int t2) {
    QB_READY_NOFLUSH(); QB_CHECK_STATE(q); QB_CHECK_SUFFIX(ctrls, nc, q);
    int ts[2] = {t2, t1}, tst[2] = {0, 1};
    QB_CHECK_SUFFIX(ts, 2, q); QB_REQUIRE(t1 != t2, "swap: identical targets");
    if (qb_tile_try_swap(q, ctrls, cs, nc, t1, t2)) return qb_tile_status();
    QB_FLUSH();
    OpSwapA op; op.ins = qb_make_ins(ctrls, cs, nc, ts, tst, 2); op.flip = pow2(t1) | pow2(t2);
    return qb_launch_tuple((cplx*)q->amps, q->numAmpsPerNode >> (2 + nc), op);
}

int qb_statevec_anyCtrlSwap_subB(const qb_state* q, const int* ctrls, const int* cs, int nc) {
    QB_READY(); QB_CHECK_STATE(q); QB_CHECK_SUFFIX(ctrls, nc, q);
    QB_REQUIRE(q->buffer, "swap subB: state has no communication buffer");
    OpUnpack op; op.ins = qb_make_ins(ctrls, cs, nc, nullptr, nullptr, 0); op.buf = (const cplx*)q->buffer;
    return qb_launch_map((cplx*)q->amps, q->numAmpsPerNode >> nc, op);
}

int qb_statevec_anyCtrlSwap_subC(const qb_state* q, const int* ctrls, const int* cs, int nc, int targ, int targState) {
    QB_READY(); QB_CHECK_STATE(q); QB_CHECK_SUFFIX(ctrls, nc, q); QB_CHECK_SUFFIX(&targ, 1, q);
    QB_REQUIRE(q->buffer, "swap subC: state has no communication buffer");
    OpUnpack op; op.ins = qb_make_ins(ctrls, cs, nc, &targ, &targState, 1); op.buf = (const cplx*)q->buffer;
    return qb_launch_map((cplx*)q->amps, q->numAmpsPerNode >> (1 + nc), op);
}

int qb_statevec_anyCtrlOneTargDenseMatr_subA(const qb_state* q, const int* ctrls, const int* cs, int nc, int targ, const qb_cplx m[4]) {
    QB_READY_NOFLUSH(); QB_CHECK_STATE(q); QB_CHECK_SUFFIX(ctrls, nc, q); QB_CHECK_SUFFIX(&targ, 1, q);
    if (qb_tile_try_dense(q, ctrls, cs, nc, &targ, 1, m)) return qb_tile_status();
    QB_FLUSH();
    int zero = 0;
    OpDense1 op; op.ins = qb_make_ins(ctrls, cs, nc, &targ, &zero, 1); op.tbit = pow2(targ);
    op.m00 = mk(m[0]); op.m01 = mk(m[1]); op.m10 = mk(m[2]); op.m11 = mk(m[3]);
    return qb_launch_tuple((cplx*)q->amps, q->numAmpsPerNode >> (1 + nc), op);
}

int qb_statevec_anyCtrlOneTargDenseMatr_subB(const qb_state* q, const int* ctrls, const int* cs, int nc, qb_cplx f0, qb_cplx f1) {
    QB_READY(); QB_CHECK_STATE(q); QB_CHECK_SUFFIX(ctrls, nc, q);
    QB_REQUIRE(q->buffer, "dense subB: state has no communication buffer");
    OpDense1B op; op.ins = qb_make_ins(ctrls, cs, nc, nullptr, nullptr, 0); op.buf = (const cplx*)q->buffer;
    op.f0 = mk(f0); op.f1 = mk(f1);
    return qb_launch_map((cplx*)q->amps, q->numAmpsPerNode >> nc, op);
}

int qb_statevec_anyCtrlTwoTargDenseMatr_sub(const qb_state* q, const int* ctrls, const int* cs, int nc, int t1, int t2, const qb_cplx m[16]) {
    QB_READY_NOFLUSH(); QB_CHECK_STATE(q); QB_CHECK_SUFFIX(ctrls, nc, q);
    int ts[2] = {t1, t2}, z[2] = {0, 0};
    QB_CHECK_SUFFIX(ts, 2, q); QB_REQUIRE(t1 != t2, "dense2: identical targets");
    if (qb_tile_try_dense(q, ctrls, cs, nc, ts, 2, m)) return qb_tile_status();
    QB_FLUSH();
    OpDense2 op; op.ins = qb_make_ins(ctrls, cs, nc, ts, z, 2); op.b1 = pow2(t1); op.b2 = pow2(t2);
    for (int i = 0; i < 16; i++) op.m[i] = mk(m[i]);
    return qb_launch_tuple((cplx*)q->amps, q->numAmpsPerNode >> (2 + nc), op);
}

int qb_statevec_anyCtrlAnyTargDenseMatr_sub(const qb_state* q, const int* ctrls, const int* cs, int nc,
        const int* targs, int nt, const qb_cplx* devMatr, int conj) {
    QB_READY(); QB_CHECK_STATE(q); QB_CHECK_SUFFIX(ctrls, nc, q); QB_CHECK_SUFFIX(targs, nt, q);
    QB_REQUIRE(nt >= 1 && devMatr, "denseK: need >= 1 target and a device matrix");
    QB_REQUIRE(nc + nt <= q->logNumAmpsPerNode, "denseK: more qubits than the local state holds");
    int z[QB_MAX_QUBITS] = {0};
    BitIns ins = qb_make_ins(ctrls, cs, nc, targs, z, nt);
    BitList tl = qb_make_list(targs, nt);
    return qb_denseK(q, ins, tl, q->numAmpsPerNode >> (nc + nt), (const cplx*)devMatr, conj);
}

int qb_statevec_anyCtrlOneTargDiagMatr_sub(const qb_state* q, const int* ctrls, const int* cs, int nc, int targ, const qb_cplx e[2]) {
    QB_READY_NOFLUSH(); QB_CHECK_STATE(q); QB_CHECK_SUFFIX(ctrls, nc, q); QB_CHECK_GLOBAL(&targ, 1);
    if (qb_tile_try_diag(q, ctrls, cs, nc, &targ, 1, e)) return qb_tile_status();
    QB_FLUSH();
    OpDiag1 op; op.ins = qb_make_ins(ctrls, cs, nc, nullptr, nullptr, 0);
    op.rankBits = (qindex)q->rank << q->logNumAmpsPerNode; op.targ = targ; op.e0 = mk(e[0]); op.e1 = mk(e[1]);
    return qb_launch_map((cplx*)q->amps, q->numAmpsPerNode >> nc, op);
}

int qb_statevec_anyCtrlTwoTargDiagMatr_sub(const qb_state* q, const int* ctrls, const int* cs, int nc, int t1, int t2, const qb_cplx e[4]) {
    QB_READY_NOFLUSH(); QB_CHECK_STATE(q); QB_CHECK_SUFFIX(ctrls, nc, q);
    int ts[2] = {t1, t2};
    QB_CHECK_GLOBAL(ts, 2);
    if (qb_tile_try_diag(q, ctrls, cs, nc, ts, 2, e)) return qb_tile_status();
    QB_FLUSH();
    OpDiag2 op; op.ins = qb_make_ins(ctrls, cs, nc, nullptr, nullptr, 0);
    op.rankBits = (qindex)q->rank << q->logNumAmpsPerNode; op.t1 = t1; op.t2 = t2;
    for (int i = 0; i < 4; i++) op.e[i] = mk(e[i]);
    return qb_launch_map((cplx*)q->amps, q->numAmpsPerNode >> nc, op);
}

int qb_statevec_anyCtrlAnyTargDiagMatr_sub(const qb_state* q, const int* ctrls, const int* cs, int nc,
        const int* targs, int nt, const qb_cplx* devElems, int conj, int hasPower, qb_cplx expo) {
    QB_READY(); QB_CHECK_STATE(q); QB_CHECK_SUFFIX(ctrls, nc, q); QB_CHECK_GLOBAL(targs, nt);
    QB_REQUIRE(devElems, "diagK: null device elements");
    OpDiagK op; op.ins = qb_make_ins(ctrls, cs, nc, nullptr, nullptr, 0);
    op.rankBits = (qindex)q->rank << q->logNumAmpsPerNode; op.targs = qb_make_list(targs, nt);
    op.elems = (const cplx*)devElems; op.conj = conj; op.hasPower = hasPower; op.expo = mk(expo);
    return qb_launch_map((cplx*)q->amps, q->numAmpsPerNode >> nc, op);
}

int qb_statevec_allTargDiagMatr_sub(const qb_state* q, const qb_cplx* devElems, int hasPower, qb_cplx expo) {
    QB_READY(); QB_CHECK_STATE(q); QB_REQUIRE(devElems, "allTargDiag: null device elements");
    OpAllDiagSV op; op.elems = (const cplx*)devElems; op.hasPower = hasPower; op.expo = mk(expo);
    return qb_launch_map((cplx*)q->amps, q->numAmpsPerNode, op);
}

int qb_densmatr_allTargDiagMatr_sub(const qb_state* q, const qb_cplx* devElems, qb_index dim, int hasPower, int mulOnly, qb_cplx expo) {
    QB_READY(); QB_CHECK_STATE(q); QB_REQUIRE(devElems && dim > 0, "allTargDiag(dm): bad matrix");
    OpAllDiagDM op; op.elems = (const cplx*)devElems; op.dim = dim;
    op.rankBits = (qindex)q->rank << q->logNumAmpsPerNode; op.hasPower = hasPower; op.mulOnly = mulOnly; op.expo = mk(expo);
    return qb_launch_map((cplx*)q->amps, q->numAmpsPerNode, op);
}

static cplx powerOfI(int n) {           // util_getPowerOfI, core/utilities.cpp
    switch (n & 3) { case 0: return mk(1, 0); case 1: return mk(0, 1); case 2: return mk(-1, 0); default: return mk(0, -1); }
}

int qb_statevector_anyCtrlPauliTensorOrGadget_subA(const qb_state* q, const int* ctrls, const int* cs, int nc,
        const int* x, int nx, const int* y, int ny, const int* z, int nz, qb_cplx ampFac, qb_cplx pairAmpFac) {
    QB_READY_NOFLUSH(); QB_CHECK_STATE(q); QB_CHECK_SUFFIX(ctrls, nc, q);
    QB_CHECK_SUFFIX(x, nx, q); QB_CHECK_SUFFIX(y, ny, q); QB_CHECK_SUFFIX(z, nz, q);
    QB_REQUIRE(nx + ny >= 1, "pauli subA: needs at least one X or Y target");
    unsigned long long maskXY = qb_make_mask(x, nx) | qb_make_mask(y, ny);
    unsigned long long maskYZ = qb_make_mask(y, ny) | qb_make_mask(z, nz);
    cplx pf = cmul(mk(pairAmpFac), powerOfI(ny));
    if (qb_tile_try_pauli(q, ctrls, cs, nc, maskXY, maskYZ, mk(ampFac), pf)) return qb_tile_status();
    QB_FLUSH();
    qb_cplx pfc = {pf.x, pf.y};
    return qb_pauli_raw(q, ctrls, cs, nc, maskXY, maskYZ, ampFac, pfc);
}

int qb_statevector_anyCtrlPauliTensorOrGadget_subB(const qb_state* q, const int* ctrls, const int* cs, int nc,
        const int* x, int nx, const int* y, int ny, const int* z, int nz, qb_cplx ampFac, qb_cplx pairAmpFac, qb_index bufMaskXY) {
    QB_READY(); QB_CHECK_STATE(q); QB_CHECK_SUFFIX(ctrls, nc, q);
    QB_CHECK_SUFFIX(x, nx, q); QB_CHECK_SUFFIX(y, ny, q); QB_CHECK_SUFFIX(z, nz, q);
    QB_REQUIRE(q->buffer, "pauli subB: state has no communication buffer");
    OpPauliB op; op.ins = qb_make_ins(ctrls, cs, nc, nullptr, nullptr, 0); op.buf = (const cplx*)q->buffer;
    op.maskXY = (qindex)(qb_make_mask(x, nx) | qb_make_mask(y, ny));
    op.maskYZ = (qindex)(qb_make_mask(y, ny) | qb_make_mask(z, nz));
    op.bufMaskXY = bufMaskXY; op.ampFac = mk(ampFac); op.pairFac = cmul(mk(pairAmpFac), powerOfI(ny));
    return qb_launch_map((cplx*)q->amps, q->numAmpsPerNode >> nc, op);
}

int qb_statevector_anyCtrlAnyTargZOrPhaseGadget_sub(const qb_state* q, const int* ctrls, const int* cs, int nc,
        const int* targs, int nt, qb_cplx f0, qb_cplx f1) {
    QB_READY_NOFLUSH(); QB_CHECK_STATE(q); QB_CHECK_SUFFIX(ctrls, nc, q); QB_CHECK_SUFFIX(targs, nt, q);
    unsigned long long targMask = qb_make_mask(targs, nt);
    if (qb_tile_try_phase(q, ctrls, cs, nc, targMask, mk(f0), mk(f1))) return qb_tile_status();
    QB_FLUSH();
    OpPhaseGadget op; op.ins = qb_make_ins(ctrls, cs, nc, nullptr, nullptr, 0);
    op.targMask = (qindex)targMask; op.f0 = mk(f0); op.f1 = mk(f1);
    return qb_launch_map((cplx*)q->amps, q->numAmpsPerNode >> nc, op);
}

int qb_statevec_setQuregToSuperposition_sub(qb_cplx fOut, const qb_state* out, qb_cplx f1, const qb_state* in1, qb_cplx f2, const qb_state* in2) {
    QB_READY(); QB_CHECK_STATE(out); QB_CHECK_STATE(in1); QB_CHECK_STATE(in2);
    QB_REQUIRE(out->numAmpsPerNode == in1->numAmpsPerNode && out->numAmpsPerNode == in2->numAmpsPerNode, "superposition: size mismatch");
    OpSuperpose op; op.in1 = (const cplx*)in1->amps; op.in2 = (const cplx*)in2->amps;
    op.fOut = mk(fOut); op.f1 = mk(f1); op.f2 = mk(f2);
    return qb_launch_map((cplx*)out->amps, out->numAmpsPerNode, op);
}

int qb_statevec_multiQubitProjector_sub(const qb_state* q, const int* qubits, const int* outcomes, int nq, double prob) {
    QB_READY(); QB_CHECK_STATE(q); QB_CHECK_SUFFIX(qubits, nq, q);
    OpProjSV op; op.qubitMask = (qindex)qb_make_mask(qubits, nq); op.retainMask = 0;
    for (int i = 0; i < nq; i++) if (outcomes[i]) op.retainMask |= pow2(qubits[i]);
    op.renorm = 1.0 / sqrt(prob);
    return qb_launch_map((cplx*)q->amps, q->numAmpsPerNode, op);
}

int qb_statevec_initUniformState_sub(const qb_state* q, qb_cplx amp) {
    QB_READY(); QB_CHECK_STATE(q);
    OpFill op; op.amp = mk(amp);
    return qb_launch_map((cplx*)q->amps, q->numAmpsPerNode, op);
}

int qb_statevec_initDebugState_sub(const qb_state* q) {
    QB_READY(); QB_CHECK_STATE(q);
    OpDebug op; op.rankBits = (qindex)q->rank << q->logNumAmpsPerNode;
    return qb_launch_map((cplx*)q->amps, q->numAmpsPerNode, op);
}

int qb_statevec_initUnnormalisedUniformlyRandomPureStateAmps_sub(const qb_state* q, unsigned seed) {
    QB_READY(); QB_CHECK_STATE(q);
    OpRandom op; op.seed = seed; op.rankBits = (qindex)q->rank << q->logNumAmpsPerNode;
    return qb_launch_map((cplx*)q->amps, q->numAmpsPerNode, op);
}

} // extern "C"
