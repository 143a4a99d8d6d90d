// qb_densmatr.cu -- density-matrix kernels: decoherence channels, mixing, projector, partial trace and
// Pauli-sum initialisation.  A density matrix of n qubits is a column-major Choi vector of 2n "qubits":
// ket qubit q <-> bit q, bra qubit q <-> bit q+n of the flat index (core/utilities.cpp:50-55,
// core/fastmath.hpp:87-91).  Semantics follow quest/src/cpu/cpu_subroutines.cpp:1017-1700, 2311-2345;
// channel coefficients follow core/utilities.cpp:922-990.
#include "qb_common.cuh"
#include "qb_kernels.cuh"

// ---- channel factors (core/utilities.cpp:922-990) ---------------------------------------------------
static inline double dephFac1(double p) { return 1 - 2 * p; }                 // util_getOneQubitDephasingFactor
static inline double dephTerm2(double p) { return -4 * p / 3; }               // util_getTwoQubitDephasingTerm
struct Fac3 { double c1, c2, c3, c4; };
static inline Fac3 depolFac1(double p) { return {1 - 2 * p / 3, 2 * p / 3, 1 - 4 * p / 3, 0}; }
static inline Fac3 depolFac2(double p) { return {1 - 4 * p / 5, 4 * p / 15, -16 * p / 15, 0}; }
static inline Fac3 dampFac(double p) { return {sqrt(1 - p), 1 - p, 0, 0}; }
static inline Fac3 pauliFac(double pI, double pX, double pY, double pZ) {
    return {pI + pZ, pX + pY, pI - pZ, pX - pY};
}

static inline int braOf(const qb_state* q, int ket) { return ket + q->numQubits; }   // util_getBraQubit
// util_getRankBitOfBraQubit: the bra qubit lies in the prefix, i.e. is a bit of the rank
static inline int rankBitOfBra(const qb_state* q, int ket) {
    int pref = braOf(q, ket) - q->logNumAmpsPerNode;
    return (q->rank >> pref) & 1;
}

// ---- ops --------------------------------------------------------------------------------------------
struct OpScaleWhere {           // amps[ins(n)] *= fac   (dephasing subB, damping subB/subC)
    static constexpr bool READS = true;
    BitIns ins; double fac;
    __device__ __forceinline__ qindex index(qindex n) const { return ins(n); }
    __device__ __forceinline__ cplx second(qindex, qindex) const { return mk(0, 0); }
    __device__ __forceinline__ cplx apply(qindex, qindex, cplx a, cplx) const { return cscale(fac, a); }
};

struct OpAddBuf {               // amps[ins(n)] += c * buf[n]   (damping subD, 2q-depol subF)
    static constexpr bool READS = true;
    BitIns ins; const cplx* buf; double c;
    __device__ __forceinline__ qindex index(qindex n) const { return ins(n); }
    __device__ __forceinline__ cplx second(qindex n, qindex) const { return buf[n]; }
    __device__ __forceinline__ cplx apply(qindex, qindex, cplx a, cplx b) const { return mk(a.x + c * b.x, a.y + c * b.y); }
};

// amps[n] *= (flag(global i) ? fA : fB) where flag tests XOR / equality of bit pairs
struct OpDeph2 {                // cpu_subroutines.cpp:1148-1175  (also used for subA)
    static constexpr bool READS = true;
    qindex rankBits; int kA, bA, kB, bB; double term;
    __device__ __forceinline__ qindex index(qindex n) const { return n; }
    __device__ __forceinline__ cplx second(qindex, qindex) const { return mk(0, 0); }
    __device__ __forceinline__ cplx apply(qindex n, qindex, cplx a, cplx) const {
        qindex i = rankBits | n;
        int flag = (getBit(i, kA) ^ getBit(i, bA)) | (getBit(i, kB) ^ getBit(i, bB));
        return cscale(1 + flag * term, a);
    }
};

struct OpDepol2A {              // cpu_subroutines.cpp:1264-1287
    static constexpr bool READS = true;
    int k1, b1, k2, b2; double c3;
    __device__ __forceinline__ qindex index(qindex n) const { return n; }
    __device__ __forceinline__ cplx second(qindex, qindex) const { return mk(0, 0); }
    __device__ __forceinline__ cplx apply(qindex n, qindex, cplx a, cplx) const {
        bool f1 = getBit(n, k1) == getBit(n, b1), f2 = getBit(n, k2) == getBit(n, b2);
        int mod = !(f1 & f2);
        return cscale(1 + c3 * mod, a);
    }
};

struct OpDepol2C {              // cpu_subroutines.cpp:1330-1359
    static constexpr bool READS = true;
    int k1, b1, k2, braBit2; double c3;
    __device__ __forceinline__ qindex index(qindex n) const { return n; }
    __device__ __forceinline__ cplx second(qindex, qindex) const { return mk(0, 0); }
    __device__ __forceinline__ cplx apply(qindex n, qindex, cplx a, cplx) const {
        bool f1 = getBit(n, k1) == getBit(n, b1), f2 = getBit(n, k2) == braBit2;
        int mod = !(f1 & f2);
        return cscale(1 + c3 * mod, a);
    }
};

struct OpDepol2E {              // cpu_subroutines.cpp:1399-1424
    static constexpr bool READS = true;
    int k1, k2, braBit1, braBit2; double fac0, fac1;
    __device__ __forceinline__ qindex index(qindex n) const { return n; }
    __device__ __forceinline__ cplx second(qindex, qindex) const { return mk(0, 0); }
    __device__ __forceinline__ cplx apply(qindex n, qindex, cplx a, cplx) const {
        int flag = (getBit(n, k1) == braBit1) & (getBit(n, k2) == braBit2);
        return cscale(fac1 * flag + fac0, a);
    }
};

struct OpDeph1A {               // cpu_subroutines.cpp:1077-1106: scale |.0.><.1.| and |.1.><.0.|
    static constexpr int M = 2;
    BitIns ins; qindex flip; double fac;          // ins: bra=0, ket=1  -> i01; i10 = i01 ^ (bra|ket)
    __device__ __forceinline__ void indices(qindex n, qindex* idx) const { idx[0] = ins(n); idx[1] = idx[0] ^ flip; }
    __device__ __forceinline__ void apply(const qindex*, cplx* v) const { v[0] = cscale(fac, v[0]); v[1] = cscale(fac, v[1]); }
    __device__ __forceinline__ bool writes(int) const { return true; }
};

struct OpQuad {                 // the four amps of a (ket,bra) qubit pair: idx = {i00, i01(ket), i10(bra), i11}
    BitIns ins; qindex ketBit, braBit;
    __device__ __forceinline__ void indices(qindex n, qindex* idx) const {
        idx[0] = ins(n); idx[1] = idx[0] | ketBit; idx[2] = idx[0] | braBit; idx[3] = idx[1] | braBit;
    }
};

struct OpDepol1A : OpQuad {     // cpu_subroutines.cpp:1183-1215
    static constexpr int M = 4;
    double fAA, fBB, fAB;
    __device__ __forceinline__ void apply(const qindex*, cplx* v) const {
        cplx a00 = v[0], a11 = v[3];
        v[0] = mk(fAA * a00.x + fBB * a11.x, fAA * a00.y + fBB * a11.y);
        v[1] = cscale(fAB, v[1]);
        v[2] = cscale(fAB, v[2]);
        v[3] = mk(fAA * a11.x + fBB * a00.x, fAA * a11.y + fBB * a00.y);
    }
    __device__ __forceinline__ bool writes(int) const { return true; }
};

struct OpPauliChA : OpQuad {    // cpu_subroutines.cpp:1455-1497
    static constexpr int M = 4;
    double fAA, fBB, fAB, fBA;
    __device__ __forceinline__ void apply(const qindex*, cplx* v) const {
        cplx a00 = v[0], a01 = v[1], a10 = v[2], a11 = v[3];
        v[0] = mk(fAA * a00.x + fBB * a11.x, fAA * a00.y + fBB * a11.y);
        v[1] = mk(fAB * a01.x + fBA * a10.x, fAB * a01.y + fBA * a10.y);
        v[2] = mk(fAB * a10.x + fBA * a01.x, fAB * a10.y + fBA * a01.y);
        v[3] = mk(fAA * a11.x + fBB * a00.x, fAA * a11.y + fBB * a00.y);
    }
    __device__ __forceinline__ bool writes(int) const { return true; }
};

struct OpDampA : OpQuad {       // cpu_subroutines.cpp:1550-1579
    static constexpr int M = 4;
    double prob, c1, c2;
    __device__ __forceinline__ void apply(const qindex*, cplx* v) const {
        v[0] = mk(v[0].x + prob * v[3].x, v[0].y + prob * v[3].y);
        v[1] = cscale(c1, v[1]);
        v[2] = cscale(c1, v[2]);
        v[3] = cscale(c2, v[3]);
    }
    __device__ __forceinline__ bool writes(int) const { return true; }
};

struct OpDepol2B {              // cpu_subroutines.cpp:1290-1327
    static constexpr int M = 4;
    BitIns ins; qindex f1, f2; double c1, c2;          // f1 = bra1|ket1, f2 = bra2|ket2; c1 already has c2 subtracted
    __device__ __forceinline__ void indices(qindex n, qindex* idx) const {
        idx[0] = ins(n); idx[1] = idx[0] ^ f1; idx[2] = idx[0] ^ f2; idx[3] = idx[1] ^ f2;
    }
    __device__ __forceinline__ void apply(const qindex*, cplx* v) const {
        // the reference sums left to right: ((a0000 + a0101) + a1010) + a1111
        cplx term = cadd(cadd(cadd(v[0], v[1]), v[2]), v[3]);
#pragma unroll
        for (int m = 0; m < 4; m++) v[m] = mk(c1 * v[m].x + c2 * term.x, c1 * v[m].y + c2 * term.y);
    }
    __device__ __forceinline__ bool writes(int) const { return true; }
};

// kernels that mix two local amps with one buffer amp, or two local amps with two buffer amps
struct OpDepol1B {              // cpu_subroutines.cpp:1218-1256: item n -> iAA (mix with buf[n]) and iAB (scale)
    BitIns insAA, insAB; const cplx* buf; double fAA, fBB, fAB;
};
__global__ void __launch_bounds__(QB_BLOCK) k_depol1B(cplx* __restrict__ amps, qindex numItems, OpDepol1B op) {
    qindex n = (qindex)blockIdx.x * QB_BLOCK + threadIdx.x;
    if (n >= numItems) return;
    qindex iAA = op.insAA(n), iAB = op.insAB(n);
    cplx a = amps[iAA], b = op.buf[n], c = amps[iAB];
    amps[iAA] = mk(op.fAA * a.x + op.fBB * b.x, op.fAA * a.y + op.fBB * b.y);
    amps[iAB] = cscale(op.fAB, c);
}

struct OpPauliChB {             // cpu_subroutines.cpp:1500-1541 (buffer holds the partner's full state)
    BitIns insAA; qindex ketBit; const cplx* buf; double fAA, fBB, fAB, fBA;
};
__global__ void __launch_bounds__(QB_BLOCK) k_pauliChB(cplx* __restrict__ amps, qindex numItems, OpPauliChB op) {
    qindex n = (qindex)blockIdx.x * QB_BLOCK + threadIdx.x;
    if (n >= numItems) return;
    qindex iAA = op.insAA(n), iAB = iAA ^ op.ketBit;
    cplx aAA = amps[iAA], aAB = amps[iAB], bBB = op.buf[iAB], bBA = op.buf[iAA];
    amps[iAA] = mk(op.fAA * aAA.x + op.fBB * bBB.x, op.fAA * aAA.y + op.fBB * bBB.y);
    amps[iAB] = mk(op.fAB * aAB.x + op.fBA * bBA.x, op.fAB * aAB.y + op.fBA * bBA.y);
}

struct OpDepol2D {              // cpu_subroutines.cpp:1362-1396
    BitIns ins; qindex flip; const cplx* buf; double c1, c2;
};
__global__ void __launch_bounds__(QB_BLOCK) k_depol2D(cplx* __restrict__ amps, qindex numItems, OpDepol2D op) {
    qindex n = (qindex)blockIdx.x * QB_BLOCK + threadIdx.x;
    if (n >= numItems) return;
    qindex i0 = op.ins(n), i1 = i0 ^ op.flip;
    cplx a0 = amps[i0], a1 = amps[i1], b = op.buf[n];
    amps[i0] = mk(op.c1 * a0.x + op.c2 * (a1.x + b.x), op.c1 * a0.y + op.c2 * (a1.y + b.y));
    amps[i1] = mk(op.c1 * a1.x + op.c2 * (a0.x + b.x), op.c1 * a1.y + op.c2 * (a0.y + b.y));
}

struct OpMixDM {                // cpu_subroutines.cpp:1017-1026
    static constexpr bool READS = true;
    const cplx* in; double pOut, pIn;
    __device__ __forceinline__ qindex index(qindex n) const { return n; }
    __device__ __forceinline__ cplx second(qindex n, qindex) const { return in[n]; }
    __device__ __forceinline__ cplx apply(qindex, qindex, cplx a, cplx b) const {
        return mk(pOut * a.x + pIn * b.x, pOut * a.y + pIn * b.y);
    }
};

struct OpMixSV {                // cpu_subroutines.cpp:1029-1068 (subB: psi local; subC: psi in buffer, global index)
    static constexpr bool READS = true;
    const cplx* psi; qindex dim; qindex rankBits; double pOut, pIn;
    __device__ __forceinline__ qindex index(qindex n) const { return n; }
    __device__ __forceinline__ cplx second(qindex n, qindex) const {
        qindex m = rankBits | n;
        cplx r = __ldg(&psi[m % dim]), c = __ldg(&psi[m / dim]);
        return cmul(cscale(pIn, r), cconj(c));          // (inProb * in[i]) * conj(in[j])
    }
    __device__ __forceinline__ cplx apply(qindex, qindex, cplx a, cplx b) const {
        return mk(pOut * a.x + b.x, pOut * a.y + b.y);
    }
};

struct OpProjDM {               // cpu_subroutines.cpp:2311-2345
    static constexpr bool READS = true;
    qindex rankBits; int numQubits; qindex qubitMask, retainMask; double renorm;
    __device__ __forceinline__ qindex index(qindex n) const { return n; }
    __device__ __forceinline__ cplx second(qindex, qindex) const { return mk(0, 0); }
    __device__ __forceinline__ cplx apply(qindex n, qindex, cplx a, cplx) const {
        qindex i = rankBits | n;
        qindex r = i & (pow2(numQubits) - 1), c = i >> numQubits;
        bool keep = ((r & qubitMask) == retainMask) && ((c & qubitMask) == retainMask);
        return cscale(keep ? renorm : 0.0, a);
    }
};

// partial trace: out[n] = sum_j in[k | targs=j | pairTargs=j]      cpu_subroutines.cpp:1645-1696
__global__ void __launch_bounds__(QB_BLOCK) k_partialTrace(const cplx* __restrict__ in, cplx* __restrict__ out,
        qindex numOut, BitIns ins, BitList targs, BitList pairs) {
    qindex n = (qindex)blockIdx.x * QB_BLOCK + threadIdx.x;
    if (n >= numOut) return;
    qindex k = ins(n);
    qindex numInner = pow2(targs.n);
    cplx acc = mk(0, 0);
    for (qindex j = 0; j < numInner; j++) {
        qindex i = k | targs.scatter(j) | pairs.scatter(j);
        acc = cadd(acc, in[i]);
    }
    out[n] = acc;
}

// <r| sum_t c_t P_t |c>   (core/fastmath.hpp:124-197).  One Pauli per qubit: I/Z are diagonal, X/Y flip.
// matrix element of a single string = prod over qubits; non-zero iff (r^c) == maskXY, then
// value = i^{numY} * (-1)^{popc(c & maskY)} ... evaluated here literally per qubit to keep the reference's
// arithmetic (products of exact 0, +-1, +-i) -- the result is exact in either formulation.
__device__ __forceinline__ cplx pauliStrElem(unsigned long long lo, unsigned long long hi, qindex row, qindex col) {
    // decode base-4 masks into X/Y/Z bit masks (2 bits per qubit: I=0 X=1 Y=2 Z=3, api/paulis.cpp:306-331)
    unsigned long long x = 0, y = 0, z = 0;
#pragma unroll 1
    for (int t = 0; t < 32; t++) {
        int p = (int)((lo >> (2 * t)) & 3);
        x |= (unsigned long long)(p == 1) << t; y |= (unsigned long long)(p == 2) << t; z |= (unsigned long long)(p == 3) << t;
        int ph = (int)((hi >> (2 * t)) & 3);
        x |= (unsigned long long)(ph == 1) << (t + 32); y |= (unsigned long long)(ph == 2) << (t + 32); z |= (unsigned long long)(ph == 3) << (t + 32);
    }
    unsigned long long flip = (unsigned long long)(row ^ col);
    if (flip != (x | y)) return mk(0, 0);
    // Z contributes (-1)^{bit}; Y = [[0,-i],[i,0]]: element (row=1,col=0) = +i, (row=0,col=1) = -i
    int numY = __popcll(y);
    int neg = (__popcll((unsigned long long)col & z) + __popcll((unsigned long long)col & y)) & 1; // Y with col bit 1 -> -i
    // i^{numY} * (-1)^{neg}
    int ph = (numY + 2 * neg) & 3;
    switch (ph) { case 0: return mk(1, 0); case 1: return mk(0, 1); case 2: return mk(-1, 0); default: return mk(0, -1); }
}

struct PauliSumArgs { const cplx* coeffs; const unsigned long long* strings; qindex numTerms; };

__global__ void __launch_bounds__(QB_BLOCK) k_setDMToPauliSum(cplx* __restrict__ amps, qindex numAmps, qindex rankBits,
        qindex dim, PauliSumArgs a) {
    qindex n = (qindex)blockIdx.x * QB_BLOCK + threadIdx.x;
    if (n >= numAmps) return;
    qindex i = rankBits | n, r = i % dim, c = i / dim;
    cplx acc = mk(0, 0);
    for (qindex t = 0; t < a.numTerms; t++)
        acc = cadd(acc, cmul(a.coeffs[t], pauliStrElem(a.strings[2 * t], a.strings[2 * t + 1], r, c)));
    amps[n] = acc;
}

__global__ void __launch_bounds__(QB_BLOCK) k_setDiagToPauliSum(cplx* __restrict__ elems, qindex numElems, qindex rankBits,
        PauliSumArgs a) {
    qindex n = (qindex)blockIdx.x * QB_BLOCK + threadIdx.x;
    if (n >= numElems) return;
    qindex i = rankBits | n;
    cplx acc = mk(0, 0);
    for (qindex t = 0; t < a.numTerms; t++)
        acc = cadd(acc, cmul(a.coeffs[t], pauliStrElem(a.strings[2 * t], a.strings[2 * t + 1], i, i)));
    elems[n] = acc;
}

static int uploadPauliSum(const qb_cplx* coeffs, const unsigned long long* strings, qindex numTerms,
                          cplx** dCoeffs, unsigned long long** dStrings) {
    QB_CUDA(cudaMalloc(dCoeffs, sizeof(cplx) * numTerms));
    QB_CUDA(cudaMalloc(dStrings, sizeof(unsigned long long) * 2 * numTerms));
    QB_CUDA(cudaMemcpyAsync(*dCoeffs, coeffs, sizeof(cplx) * numTerms, cudaMemcpyHostToDevice, g_qb.stream));
    QB_CUDA(cudaMemcpyAsync(*dStrings, strings, sizeof(unsigned long long) * 2 * numTerms, cudaMemcpyHostToDevice, g_qb.stream));
    return 0;
}

#define QB_CHECK_DM(q) do { QB_CHECK_STATE(q); QB_REQUIRE((q)->isDensityMatrix && (q)->numQubits > 0, "state is not a density matrix"); } while (0)
#define QB_CHECK_KET_SUFFIX_BRA(q, k) QB_REQUIRE((k) >= 0 && braOf(q, k) < (q)->logNumAmpsPerNode, "ket/bra qubit not local")
#define QB_CHECK_KET_PREFIX_BRA(q, k) QB_REQUIRE((k) >= 0 && (k) < (q)->logNumAmpsPerNode && braOf(q, k) >= (q)->logNumAmpsPerNode, "bra qubit must be a prefix qubit")
#define QB_GRID1(n) (unsigned int)(((n) + QB_BLOCK - 1) / QB_BLOCK)

extern "C" {

int qb_densmatr_mixQureg_subA(double pOut, const qb_state* out, double pIn, const qb_state* in) {
    QB_READY(); QB_CHECK_STATE(out); QB_CHECK_STATE(in);
    QB_REQUIRE(out->numAmpsPerNode == in->numAmpsPerNode, "mixQureg subA: size mismatch");
    OpMixDM op; op.in = (const cplx*)in->amps; op.pOut = pOut; op.pIn = pIn;
    return qb_launch_map((cplx*)out->amps, out->numAmpsPerNode, op);
}

int qb_densmatr_mixQureg_subB(double pOut, const qb_state* out, double pIn, const qb_state* sv) {
    QB_READY(); QB_CHECK_DM(out); QB_CHECK_STATE(sv);
    OpMixSV op; op.psi = (const cplx*)sv->amps; op.dim = pow2(out->numQubits);
    op.rankBits = 0;   // subB is only used when out is not distributed (localiser.cpp:1383-1411)
    op.pOut = pOut; op.pIn = pIn;
    return qb_launch_map((cplx*)out->amps, out->numAmpsPerNode, op);
}

int qb_densmatr_mixQureg_subC(double pOut, const qb_state* out, double pIn) {
    QB_READY(); QB_CHECK_DM(out); QB_REQUIRE(out->buffer, "mixQureg subC: no communication buffer");
    OpMixSV op; op.psi = (const cplx*)out->buffer; op.dim = pow2(out->numQubits);
    op.rankBits = (qindex)out->rank << out->logNumAmpsPerNode; op.pOut = pOut; op.pIn = pIn;
    return qb_launch_map((cplx*)out->amps, out->numAmpsPerNode, op);
}

int qb_densmatr_oneQubitDephasing_subA(const qb_state* q, int ket, double prob) {
    QB_READY(); QB_CHECK_DM(q); QB_CHECK_KET_SUFFIX_BRA(q, ket);
    int bra = braOf(q, ket), qs[2] = {ket, bra}, st[2] = {1, 0};
    OpDeph1A op; op.ins = qb_make_ins(qs, st, 2, nullptr, nullptr, 0); op.flip = pow2(ket) | pow2(bra); op.fac = dephFac1(prob);
    return qb_launch_tuple((cplx*)q->amps, q->numAmpsPerNode / 4, op);
}

int qb_densmatr_oneQubitDephasing_subB(const qb_state* q, int ket, double prob) {
    QB_READY(); QB_CHECK_DM(q); QB_CHECK_KET_PREFIX_BRA(q, ket);
    int st = !rankBitOfBra(q, ket);
    OpScaleWhere op; op.ins = qb_make_ins(&ket, &st, 1, nullptr, nullptr, 0); op.fac = dephFac1(prob);
    return qb_launch_map((cplx*)q->amps, q->numAmpsPerNode / 2, op);
}

int qb_densmatr_twoQubitDephasing_subB(const qb_state* q, int kA, int kB, double prob) {
    QB_READY(); QB_CHECK_DM(q);
    QB_REQUIRE(kA >= 0 && kB >= 0 && kA < q->numQubits && kB < q->numQubits && kA != kB, "twoQubitDephasing: bad qubits");
    OpDeph2 op; op.rankBits = (qindex)q->rank << q->logNumAmpsPerNode;
    op.kA = kA; op.bA = braOf(q, kA); op.kB = kB; op.bB = braOf(q, kB); op.term = dephTerm2(prob);
    return qb_launch_map((cplx*)q->amps, q->numAmpsPerNode, op);
}

int qb_densmatr_twoQubitDephasing_subA(const qb_state* q, int kA, int kB, double prob) {
    return qb_densmatr_twoQubitDephasing_subB(q, kA, kB, prob);      // identical, cpu_subroutines.cpp:1135-1145
}

static OpQuad makeQuad(const qb_state* q, int ket) {
    int bra = braOf(q, ket), qs[2] = {ket, bra}, st[2] = {0, 0};
    OpQuad o; o.ins = qb_make_ins(qs, st, 2, nullptr, nullptr, 0); o.ketBit = pow2(ket); o.braBit = pow2(bra);
    return o;
}

int qb_densmatr_oneQubitDepolarising_subA(const qb_state* q, int ket, double prob) {
    QB_READY(); QB_CHECK_DM(q); QB_CHECK_KET_SUFFIX_BRA(q, ket);
    Fac3 f = depolFac1(prob);
    OpDepol1A op; (OpQuad&)op = makeQuad(q, ket); op.fAA = f.c1; op.fBB = f.c2; op.fAB = f.c3;
    return qb_launch_tuple((cplx*)q->amps, q->numAmpsPerNode / 4, op);
}

int qb_densmatr_oneQubitDepolarising_subB(const qb_state* q, int ket, double prob) {
    QB_READY(); QB_CHECK_DM(q); QB_CHECK_KET_PREFIX_BRA(q, ket); QB_REQUIRE(q->buffer, "depolarising subB: no buffer");
    int braBit = rankBitOfBra(q, ket), nb = !braBit;
    Fac3 f = depolFac1(prob);
    OpDepol1B op; op.insAA = qb_make_ins(&ket, &braBit, 1, nullptr, nullptr, 0); op.insAB = qb_make_ins(&ket, &nb, 1, nullptr, nullptr, 0);
    op.buf = (const cplx*)q->buffer; op.fAA = f.c1; op.fBB = f.c2; op.fAB = f.c3;
    qindex numIts = q->numAmpsPerNode / 2;
    k_depol1B<<<QB_GRID1(numIts), QB_BLOCK, 0, g_qb.stream>>>((cplx*)q->amps, numIts, op);
    QB_LAUNCH_CHECK();
    return 0;
}

int qb_densmatr_twoQubitDepolarising_subA(const qb_state* q, int k1, int k2, double prob) {
    QB_READY(); QB_CHECK_DM(q); QB_CHECK_KET_SUFFIX_BRA(q, k1); QB_CHECK_KET_SUFFIX_BRA(q, k2);
    OpDepol2A op; op.k1 = k1; op.b1 = braOf(q, k1); op.k2 = k2; op.b2 = braOf(q, k2); op.c3 = depolFac2(prob).c3;
    return qb_launch_map((cplx*)q->amps, q->numAmpsPerNode, op);
}

int qb_densmatr_twoQubitDepolarising_subB(const qb_state* q, int k1, int k2, double prob) {
    QB_READY(); QB_CHECK_DM(q); QB_CHECK_KET_SUFFIX_BRA(q, k1); QB_CHECK_KET_SUFFIX_BRA(q, k2);
    int b1 = braOf(q, k1), b2 = braOf(q, k2), qs[4] = {k1, k2, b1, b2}, st[4] = {0, 0, 0, 0};
    Fac3 f = depolFac2(prob);
    OpDepol2B op; op.ins = qb_make_ins(qs, st, 4, nullptr, nullptr, 0);
    op.f1 = pow2(b1) | pow2(k1); op.f2 = pow2(b2) | pow2(k2); op.c2 = f.c2; op.c1 = f.c1 - f.c2;
    return qb_launch_tuple((cplx*)q->amps, q->numAmpsPerNode / 16, op);
}

int qb_densmatr_twoQubitDepolarising_subC(const qb_state* q, int k1, int k2, double prob) {
    QB_READY(); QB_CHECK_DM(q); QB_CHECK_KET_SUFFIX_BRA(q, k1); QB_CHECK_KET_PREFIX_BRA(q, k2);
    OpDepol2C op; op.k1 = k1; op.b1 = braOf(q, k1); op.k2 = k2; op.braBit2 = rankBitOfBra(q, k2); op.c3 = depolFac2(prob).c3;
    return qb_launch_map((cplx*)q->amps, q->numAmpsPerNode, op);
}

int qb_densmatr_twoQubitDepolarising_subD(const qb_state* q, int k1, int k2, double prob) {
    QB_READY(); QB_CHECK_DM(q); QB_CHECK_KET_SUFFIX_BRA(q, k1); QB_CHECK_KET_PREFIX_BRA(q, k2); QB_REQUIRE(q->buffer, "depolarising subD: no buffer");
    int b1 = braOf(q, k1), braBit2 = rankBitOfBra(q, k2);
    int qs[3] = {k1, k2, b1}, st[3] = {0, braBit2, 0};
    Fac3 f = depolFac2(prob);
    OpDepol2D op; op.ins = qb_make_ins(qs, st, 3, nullptr, nullptr, 0); op.flip = pow2(b1) | pow2(k1);
    op.buf = (const cplx*)q->buffer; op.c1 = f.c1; op.c2 = f.c2;
    qindex numIts = q->numAmpsPerNode / 8;
    k_depol2D<<<QB_GRID1(numIts), QB_BLOCK, 0, g_qb.stream>>>((cplx*)q->amps, numIts, op);
    QB_LAUNCH_CHECK();
    return 0;
}

int qb_densmatr_twoQubitDepolarising_subE(const qb_state* q, int k1, int k2, double prob) {
    QB_READY(); QB_CHECK_DM(q); QB_CHECK_KET_PREFIX_BRA(q, k1); QB_CHECK_KET_PREFIX_BRA(q, k2);
    Fac3 f = depolFac2(prob);
    OpDepol2E op; op.k1 = k1; op.k2 = k2; op.braBit1 = rankBitOfBra(q, k1); op.braBit2 = rankBitOfBra(q, k2);
    op.fac0 = 1 + f.c3; op.fac1 = f.c1 - op.fac0;
    return qb_launch_map((cplx*)q->amps, q->numAmpsPerNode, op);
}

int qb_densmatr_twoQubitDepolarising_subF(const qb_state* q, int k1, int k2, double prob) {
    QB_READY(); QB_CHECK_DM(q); QB_CHECK_KET_PREFIX_BRA(q, k1); QB_CHECK_KET_PREFIX_BRA(q, k2); QB_REQUIRE(q->buffer, "depolarising subF: no buffer");
    int qs[2] = {k1, k2}, st[2] = {rankBitOfBra(q, k1), rankBitOfBra(q, k2)};
    OpAddBuf op; op.ins = qb_make_ins(qs, st, 2, nullptr, nullptr, 0); op.buf = (const cplx*)q->buffer; op.c = depolFac2(prob).c2;
    return qb_launch_map((cplx*)q->amps, q->numAmpsPerNode / 4, op);
}

int qb_densmatr_oneQubitPauliChannel_subA(const qb_state* q, int ket, double pI, double pX, double pY, double pZ) {
    QB_READY(); QB_CHECK_DM(q); QB_CHECK_KET_SUFFIX_BRA(q, ket);
    Fac3 f = pauliFac(pI, pX, pY, pZ);
    OpPauliChA op; (OpQuad&)op = makeQuad(q, ket); op.fAA = f.c1; op.fBB = f.c2; op.fAB = f.c3; op.fBA = f.c4;
    return qb_launch_tuple((cplx*)q->amps, q->numAmpsPerNode / 4, op);
}

int qb_densmatr_oneQubitPauliChannel_subB(const qb_state* q, int ket, double pI, double pX, double pY, double pZ) {
    QB_READY(); QB_CHECK_DM(q); QB_CHECK_KET_PREFIX_BRA(q, ket); QB_REQUIRE(q->buffer, "pauli channel subB: no buffer");
    int braBit = rankBitOfBra(q, ket);
    Fac3 f = pauliFac(pI, pX, pY, pZ);
    OpPauliChB op; op.insAA = qb_make_ins(&ket, &braBit, 1, nullptr, nullptr, 0); op.ketBit = pow2(ket);
    op.buf = (const cplx*)q->buffer; op.fAA = f.c1; op.fBB = f.c2; op.fAB = f.c3; op.fBA = f.c4;
    qindex numIts = q->numAmpsPerNode / 2;
    k_pauliChB<<<QB_GRID1(numIts), QB_BLOCK, 0, g_qb.stream>>>((cplx*)q->amps, numIts, op);
    QB_LAUNCH_CHECK();
    return 0;
}

int qb_densmatr_oneQubitDamping_subA(const qb_state* q, int ket, double prob) {
    QB_READY(); QB_CHECK_DM(q); QB_CHECK_KET_SUFFIX_BRA(q, ket);
    Fac3 f = dampFac(prob);
    OpDampA op; (OpQuad&)op = makeQuad(q, ket); op.prob = prob; op.c1 = f.c1; op.c2 = f.c2;
    return qb_launch_tuple((cplx*)q->amps, q->numAmpsPerNode / 4, op);
}

int qb_densmatr_oneQubitDamping_subB(const qb_state* q, int qubit, double prob) {
    QB_READY(); QB_CHECK_DM(q); QB_CHECK_SUFFIX(&qubit, 1, q);
    int one = 1;
    OpScaleWhere op; op.ins = qb_make_ins(&qubit, &one, 1, nullptr, nullptr, 0); op.fac = dampFac(prob).c2;
    return qb_launch_map((cplx*)q->amps, q->numAmpsPerNode / 2, op);
}

int qb_densmatr_oneQubitDamping_subC(const qb_state* q, int ket, double prob) {
    QB_READY(); QB_CHECK_DM(q); QB_CHECK_KET_PREFIX_BRA(q, ket);
    int st = !rankBitOfBra(q, ket);
    OpScaleWhere op; op.ins = qb_make_ins(&ket, &st, 1, nullptr, nullptr, 0); op.fac = dampFac(prob).c1;
    return qb_launch_map((cplx*)q->amps, q->numAmpsPerNode / 2, op);
}

int qb_densmatr_oneQubitDamping_subD(const qb_state* q, int qubit, double prob) {
    QB_READY(); QB_CHECK_DM(q); QB_CHECK_SUFFIX(&qubit, 1, q); QB_REQUIRE(q->buffer, "damping subD: no buffer");
    int zero = 0;
    OpAddBuf op; op.ins = qb_make_ins(&qubit, &zero, 1, nullptr, nullptr, 0); op.buf = (const cplx*)q->buffer; op.c = prob;
    return qb_launch_map((cplx*)q->amps, q->numAmpsPerNode / 2, op);
}

int qb_densmatr_partialTrace_sub(const qb_state* in, const qb_state* out, const int* targs, const int* pairTargs, int nt) {
    QB_READY(); QB_CHECK_STATE(in); QB_CHECK_STATE(out);
    QB_CHECK_SUFFIX(targs, nt, in); QB_CHECK_SUFFIX(pairTargs, nt, in);
    QB_REQUIRE(out->numAmpsPerNode == (in->numAmpsPerNode >> (2 * nt)), "partialTrace: output size mismatch");
    BitIns ins = qb_make_ins(targs, nullptr, nt, pairTargs, nullptr, nt);
    k_partialTrace<<<QB_GRID1(out->numAmpsPerNode), QB_BLOCK, 0, g_qb.stream>>>((const cplx*)in->amps, (cplx*)out->amps,
        out->numAmpsPerNode, ins, qb_make_list(targs, nt), qb_make_list(pairTargs, nt));
    QB_LAUNCH_CHECK();
    return 0;
}

int qb_densmatr_multiQubitProjector_sub(const qb_state* q, const int* qubits, const int* outcomes, int nq, double prob) {
    QB_READY(); QB_CHECK_DM(q); QB_REQUIRE(qb_check_qubits(qubits, nq, q->numQubits), "projector: qubit out of range");
    OpProjDM op; op.rankBits = (qindex)q->rank << q->logNumAmpsPerNode; op.numQubits = q->numQubits;
    op.qubitMask = (qindex)qb_make_mask(qubits, nq); op.retainMask = 0;
    for (int i = 0; i < nq; i++) if (outcomes[i]) op.retainMask |= pow2(qubits[i]);
    op.renorm = 1.0 / prob;
    return qb_launch_map((cplx*)q->amps, q->numAmpsPerNode, op);
}

int qb_densmatr_setAmpsToPauliStrSum_sub(const qb_state* q, const qb_cplx* coeffs, const unsigned long long* strings, qb_index numTerms) {
    QB_READY(); QB_CHECK_DM(q); QB_REQUIRE(coeffs && strings && numTerms > 0, "setAmpsToPauliStrSum: empty sum");
    cplx* dC = nullptr; unsigned long long* dS = nullptr;
    int r = uploadPauliSum(coeffs, strings, numTerms, &dC, &dS);
    if (r) return r;
    PauliSumArgs a = {dC, dS, numTerms};
    k_setDMToPauliSum<<<QB_GRID1(q->numAmpsPerNode), QB_BLOCK, 0, g_qb.stream>>>((cplx*)q->amps, q->numAmpsPerNode,
        (qindex)q->rank << q->logNumAmpsPerNode, pow2(q->numQubits), a);
    QB_LAUNCH_CHECK();
    QB_CUDA(cudaStreamSynchronize(g_qb.stream));
    cudaFree(dC); cudaFree(dS);
    return 0;
}

int qb_fullstatediagmatr_setElemsToPauliStrSum(qb_cplx* devElems, qb_index numElemsPerNode, int rank,
        const qb_cplx* coeffs, const unsigned long long* strings, qb_index numTerms) {
    QB_READY(); QB_REQUIRE(devElems && numElemsPerNode > 0 && coeffs && strings && numTerms > 0, "setElemsToPauliStrSum: bad arguments");
    int logN = 0; while (pow2(logN) < numElemsPerNode) logN++;
    cplx* dC = nullptr; unsigned long long* dS = nullptr;
    int r = uploadPauliSum(coeffs, strings, numTerms, &dC, &dS);
    if (r) return r;
    PauliSumArgs a = {dC, dS, numTerms};
    k_setDiagToPauliSum<<<QB_GRID1(numElemsPerNode), QB_BLOCK, 0, g_qb.stream>>>((cplx*)devElems, numElemsPerNode,
        (qindex)rank << logN, a);
    QB_LAUNCH_CHECK();
    QB_CUDA(cudaStreamSynchronize(g_qb.stream));
    cudaFree(dC); cudaFree(dS);
    return 0;
}

} // extern "C"
