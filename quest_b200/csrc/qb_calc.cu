// qb_calc.cu -- read-only passes: probabilities, inner products, expectation values.
// Each is one streaming pass over the local amplitudes (or the local diagonal of a density matrix) into
// the deterministic grid reduction of qb_reduce.cuh.  Semantics follow
// quest/src/cpu/cpu_subroutines.cpp:1704-2270 (the oracle); the reference GPU path used Thrust
// (quest/src/gpu/gpu_thrust.cuh:745-1000) and global atomics (gpu_kernels.cuh:1146-1200).
#include "qb_common.cuh"
#include "qb_kernels.cuh"
#include "qb_reduce.cuh"
#include "qb_pauli_group.cuh"
#include <vector>
#include <algorithm>

// ---- functors ---------------------------------------------------------------------------------------
struct RNorm {                  // sum |a_i|^2 over i = ins(n)      (:1704, :1761)
    const cplx* amps; BitIns ins;
    __device__ __forceinline__ void operator()(qindex n, double& re, double&) const { re += cnorm(amps[ins(n)]); }
};
struct RNormAll {
    const cplx* amps;
    __device__ __forceinline__ void operator()(qindex n, double& re, double&) const { re += cnorm(amps[n]); }
};
struct RDiagRe {                // sum Re rho_dd over local diagonal indices i = ins(n)   (:1729, :1791)
    const cplx* amps; BitIns ins; qindex firstDiag, stride;
    __device__ __forceinline__ void operator()(qindex n, double& re, double&) const {
        re += amps[firstDiag + ins(n) * stride].x;
    }
};
struct RInner {                 // sum conj(a) b   (:1919)
    const cplx* a; const cplx* b;
    __device__ __forceinline__ void operator()(qindex n, double& re, double& im) const {
        cplx x = a[n], y = b[n];
        re += x.x * y.x + x.y * y.y;
        im += x.x * y.y - x.y * y.x;
    }
};
struct RHilbert {               // sum |a-b|^2    (:1940)
    const cplx* a; const cplx* b;
    __device__ __forceinline__ void operator()(qindex n, double& re, double&) const { re += cnorm(csub(a[n], b[n])); }
};
struct RFidelity {              // sum rho_rc psi_r* psi_c  (or the conjugated variant)   (:1956-1998)
    const cplx* rho; const cplx* psi; qindex rankBits; int numQubits; int conj;
    __device__ __forceinline__ void operator()(qindex n, double& re, double& im) const {
        qindex i = rankBits | n;
        qindex r = i & (pow2(numQubits) - 1), c = i >> numQubits;
        cplx rhoAmp = rho[n], rowAmp = __ldg(&psi[r]), colAmp = __ldg(&psi[c]);
        if (conj) { rhoAmp.y = -rhoAmp.y; colAmp.y = -colAmp.y; } else rowAmp.y = -rowAmp.y;
        cplx t = cmul(cmul(rhoAmp, rowAmp), colAmp);
        re += t.x; im += t.y;
    }
};
struct RExpecZ {                // sum (+-) |a_n|^2   (:2006)
    const cplx* amps; qindex mask;
    __device__ __forceinline__ void operator()(qindex n, double& re, double&) const {
        double s = 1.0 - 2.0 * parity64((unsigned long long)(n & mask));
        re += s * cnorm(amps[n]);
    }
};
struct RExpecZDM {              // sum (+-) rho_dd     (:2027)
    const cplx* amps; qindex mask, firstDiag, stride;
    __device__ __forceinline__ void operator()(qindex n, double& re, double& im) const {
        qindex r = n + firstDiag;
        double s = 1.0 - 2.0 * parity64((unsigned long long)(r & mask));
        cplx a = amps[firstDiag + n * stride];
        re += s * a.x; im += s * a.y;
    }
};
struct RExpecPauli {            // sum sign(j) conj(a_n) b_j, j = n ^ maskXY; b = amps (subA) or buffer (subB)  (:2059, :2090)
    const cplx* amps; const cplx* other; qindex maskXY, maskYZ;
    __device__ __forceinline__ void operator()(qindex n, double& re, double& im) const {
        qindex j = n ^ maskXY;
        double s = 1.0 - 2.0 * parity64((unsigned long long)(j & maskYZ));
        cplx x = amps[n], y = other[j];
        re += s * (x.x * y.x + x.y * y.y);
        im += s * (x.x * y.y - x.y * y.x);
    }
};
struct RExpecPauliDM {          // (:2129)
    const cplx* amps; qindex maskXY, maskYZ, firstDiag, dim;
    __device__ __forceinline__ void operator()(qindex n, double& re, double& im) const {
        qindex r = n + firstDiag;
        qindex i = r ^ maskXY;
        qindex m = i + n * dim;
        double s = 1.0 - 2.0 * parity64((unsigned long long)(i & maskYZ));
        cplx a = amps[m];
        re += s * a.x; im += s * a.y;
    }
};
struct RExpecDiag {             // sum d_n^p |a_n|^2 (SV) or d_n^p rho_dd (DM)   (:2176, :2216)
    const cplx* amps; const cplx* elems; int isDM; qindex firstDiag, stride; int hasPower, realPow; cplx expo;
    __device__ __forceinline__ void operator()(qindex n, double& re, double& im) const {
        cplx e = elems[n];
        if (hasPower && !realPow) e = cpow(e, expo);
        if (hasPower && realPow) e = mk(pow(e.x, expo.x), 0.0);
        cplx t;
        if (isDM) t = cmul(e, amps[firstDiag + n * stride]);
        else t = cscale(cnorm(amps[n]), e);
        re += t.x; im += t.y;
    }
};

// ---- all-outcome probabilities (histogram)   (:1828-1905) --------------------------------------------
struct HistArgs { const cplx* amps; qindex numItems; qindex rankBits; BitList qubits; int isDM; qindex firstDiag, stride; };

__global__ void __launch_bounds__(QB_BLOCK) k_hist(HistArgs a, double* __restrict__ out, int numBinsLog, int useSmem) {
    extern __shared__ double bins[];
    const int numBins = 1 << numBinsLog;
    if (useSmem) {
        for (int b = threadIdx.x; b < numBins; b += QB_BLOCK) bins[b] = 0.0;
        __syncthreads();
    }
    const qindex stride = (qindex)gridDim.x * QB_BLOCK;
    for (qindex n = (qindex)blockIdx.x * QB_BLOCK + threadIdx.x; n < a.numItems; n += stride) {
        double p; qindex local;
        if (a.isDM) { local = a.firstDiag + n * a.stride; p = a.amps[local].x; }
        else { local = n; p = cnorm(a.amps[n]); }
        qindex j = a.qubits.gather(a.rankBits | local);
        if (useSmem) atomicAdd(&bins[j], p); else atomicAdd(&out[j], p);
    }
    if (useSmem) {
        __syncthreads();
        for (int b = threadIdx.x; b < numBins; b += QB_BLOCK) if (bins[b] != 0.0) atomicAdd(&out[b], bins[b]);
    }
}

static int qb_hist(const HistArgs& a, int k, double* outProbs) {
    QB_REQUIRE(k >= 0 && k <= 40 && outProbs, "calcProbsOfAllMultiQubitOutcomes: bad arguments");
    qindex numBins = pow2(k);
    // grow-only device scratch for the bins (a cudaMalloc/cudaFree pair per call would synchronise the device twice)
    static double* s_bins = nullptr; static qindex s_binsLen = 0;
    if (numBins > s_binsLen) {
        QB_CUDA(cudaStreamSynchronize(g_qb.stream));
        if (s_bins) cudaFree(s_bins);
        s_bins = nullptr; s_binsLen = 0;
        QB_CUDA(cudaMalloc(&s_bins, sizeof(double) * numBins));
        s_binsLen = numBins;
    }
    double* dOut = s_bins;
    QB_CUDA(cudaMemsetAsync(dOut, 0, sizeof(double) * numBins, g_qb.stream));
    int useSmem = k <= 11;                                  // 2^11 doubles = 16 KiB of shared memory
    qindex blocks = (a.numItems + QB_BLOCK - 1) / QB_BLOCK;
    qindex maxBlocks = (qindex)g_qb.numSMs * 8;
    if (blocks > maxBlocks) blocks = maxBlocks;
    k_hist<<<(unsigned int)blocks, QB_BLOCK, useSmem ? sizeof(double) * numBins : 0, g_qb.stream>>>(a, dOut, k, useSmem);
    g_qb.launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(outProbs, dOut, sizeof(double) * numBins, cudaMemcpyDeviceToHost, g_qb.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g_qb.stream);
    if (e != cudaSuccess) return qb_set_error((int)e, "calcProbsOfAllMultiQubitOutcomes", __FILE__, __LINE__);
    return 0;
}

// ---- fused Pauli-string batch: every term sharing ONE pass over the state ------------------------------
// terms are processed in groups of up to QB_PAULI_BATCH per pass; each thread keeps the group's partial
// sums in registers, the state element a_n is loaded once per group and its partners a_{n^maskXY} come
// mostly from L2 (they lie in the same 2^k-aligned neighbourhood when the masks are low, else stream).
#define QB_PAULI_BATCH 8
struct PauliBatch { const cplx* amps; const cplx* other; qindex maskXY[QB_PAULI_BATCH], maskYZ[QB_PAULI_BATCH]; int count; };

__global__ void __launch_bounds__(QB_BLOCK) k_pauliBatch(qindex numItems, PauliBatch pb, double* partials,
                                                        unsigned int* ticket, double* out) {
    __shared__ double sm[2 * (QB_BLOCK / 32)];
    __shared__ bool isLast;
    double re[QB_PAULI_BATCH], im[QB_PAULI_BATCH];
#pragma unroll
    for (int t = 0; t < QB_PAULI_BATCH; t++) { re[t] = 0; im[t] = 0; }
    const qindex stride = (qindex)gridDim.x * QB_BLOCK;
    for (qindex n = (qindex)blockIdx.x * QB_BLOCK + threadIdx.x; n < numItems; n += stride) {
        cplx x = pb.amps[n];
#pragma unroll
        for (int t = 0; t < QB_PAULI_BATCH; t++) {
            if (t < pb.count) {
                qindex j = n ^ pb.maskXY[t];
                double s = 1.0 - 2.0 * parity64((unsigned long long)(j & pb.maskYZ[t]));
                cplx y = pb.other[j];
                re[t] += s * (x.x * y.x + x.y * y.y);
                im[t] += s * (x.x * y.y - x.y * y.x);
            }
        }
    }
#pragma unroll
    for (int t = 0; t < QB_PAULI_BATCH; t++) {
        double r = re[t], i = im[t];
        block_sum2(r, i, sm);
        if (threadIdx.x == 0) {
            partials[(2 * t) * QB_RED_MAX_BLOCKS + blockIdx.x] = r;
            partials[(2 * t + 1) * QB_RED_MAX_BLOCKS + blockIdx.x] = i;
        }
    }
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned int tk = atomicInc(ticket, gridDim.x - 1);
        isLast = (tk == gridDim.x - 1);
    }
    __syncthreads();
    if (!isLast) return;
    __threadfence();
    for (int t = 0; t < QB_PAULI_BATCH; t++) {
        double r = 0, i = 0;
        for (unsigned int b = threadIdx.x; b < gridDim.x; b += QB_BLOCK) {
            r += __ldcg(&partials[(2 * t) * QB_RED_MAX_BLOCKS + b]);
            i += __ldcg(&partials[(2 * t + 1) * QB_RED_MAX_BLOCKS + b]);
        }
        block_sum2(r, i, sm);
        if (threadIdx.x == 0) { out[2 * t] = r; out[2 * t + 1] = i; }
    }
}

// complex result of a reduction (accumulated in double) handed to a qb_cplx of the library's precision
template <typename F>
static int qb_reduce2c(qindex numItems, F f, qb_cplx* out) {
    double re = 0, im = 0;
    int r = qb_reduce2(numItems, f, &re, &im);
    if (r) return r;
    out->re = (qb_real)re; out->im = (qb_real)im;
    return 0;
}

static inline qindex firstDiagOf(const qb_state* q) { return (qindex)q->rank * pow2(q->logNumColsPerNode); }
static inline qindex diagStrideOf(const qb_state* q) { return pow2(q->numQubits) + 1; }
#define QB_CHECK_DM(q) do { QB_CHECK_STATE(q); QB_REQUIRE((q)->isDensityMatrix && (q)->numQubits > 0, "state is not a density matrix"); } while (0)

extern "C" {

int qb_statevec_calcTotalProb_sub(const qb_state* q, double* out) {
    QB_READY(); QB_CHECK_STATE(q); QB_REQUIRE(out, "null output");
    RNormAll f = {(const cplx*)q->amps};
    return qb_reduce2(q->numAmpsPerNode, f, out, nullptr);
}

int qb_densmatr_calcTotalProb_sub(const qb_state* q, double* out) {
    QB_READY(); QB_CHECK_DM(q); QB_REQUIRE(out, "null output");
    RDiagRe f; f.amps = (const cplx*)q->amps; f.ins = qb_make_ins(nullptr, nullptr, 0, nullptr, nullptr, 0);
    f.firstDiag = firstDiagOf(q); f.stride = diagStrideOf(q);
    return qb_reduce2(pow2(q->logNumColsPerNode), f, out, nullptr);
}

int qb_statevec_calcProbOfMultiQubitOutcome_sub(const qb_state* q, const int* qubits, const int* outcomes, int nq, double* out) {
    QB_READY(); QB_CHECK_STATE(q); QB_CHECK_SUFFIX(qubits, nq, q); QB_REQUIRE(out, "null output");
    RNorm f; f.amps = (const cplx*)q->amps; f.ins = qb_make_ins(qubits, outcomes, nq, nullptr, nullptr, 0);
    return qb_reduce2(q->numAmpsPerNode >> nq, f, out, nullptr);
}

int qb_densmatr_calcProbOfMultiQubitOutcome_sub(const qb_state* q, const int* qubits, const int* outcomes, int nq, double* out) {
    QB_READY(); QB_CHECK_DM(q); QB_REQUIRE(qb_check_qubits(qubits, nq, q->logNumColsPerNode), "qubits must have suffix bra qubits");
    QB_REQUIRE(out, "null output");
    RDiagRe f; f.amps = (const cplx*)q->amps; f.ins = qb_make_ins(qubits, outcomes, nq, nullptr, nullptr, 0);
    f.firstDiag = firstDiagOf(q); f.stride = diagStrideOf(q);
    return qb_reduce2(pow2(q->logNumColsPerNode - nq), f, out, nullptr);
}

int qb_statevec_calcProbsOfAllMultiQubitOutcomes_sub(double* outProbs, const qb_state* q, const int* qubits, int nq) {
    QB_READY(); QB_CHECK_STATE(q); QB_CHECK_GLOBAL(qubits, nq);
    HistArgs a; a.amps = (const cplx*)q->amps; a.numItems = q->numAmpsPerNode;
    a.rankBits = (qindex)q->rank << q->logNumAmpsPerNode; a.qubits = qb_make_list(qubits, nq); a.isDM = 0; a.firstDiag = 0; a.stride = 0;
    return qb_hist(a, nq, outProbs);
}

int qb_densmatr_calcProbsOfAllMultiQubitOutcomes_sub(double* outProbs, const qb_state* q, const int* qubits, int nq) {
    QB_READY(); QB_CHECK_DM(q); QB_CHECK_GLOBAL(qubits, nq);
    HistArgs a; a.amps = (const cplx*)q->amps; a.numItems = pow2(q->logNumColsPerNode);
    a.rankBits = (qindex)q->rank << q->logNumAmpsPerNode; a.qubits = qb_make_list(qubits, nq); a.isDM = 1;
    a.firstDiag = firstDiagOf(q); a.stride = diagStrideOf(q);
    return qb_hist(a, nq, outProbs);
}

int qb_statevec_calcInnerProduct_sub(const qb_state* a, const qb_state* b, qb_cplx* out) {
    QB_READY(); QB_CHECK_STATE(a); QB_CHECK_STATE(b); QB_REQUIRE(out && a->numAmpsPerNode == b->numAmpsPerNode, "innerProduct: bad arguments");
    RInner f = {(const cplx*)a->amps, (const cplx*)b->amps};
    return qb_reduce2c(a->numAmpsPerNode, f, out);
}

int qb_densmatr_calcHilbertSchmidtDistance_sub(const qb_state* a, const qb_state* b, double* out) {
    QB_READY(); QB_CHECK_STATE(a); QB_CHECK_STATE(b); QB_REQUIRE(out && a->numAmpsPerNode == b->numAmpsPerNode, "HS distance: bad arguments");
    RHilbert f = {(const cplx*)a->amps, (const cplx*)b->amps};
    return qb_reduce2(a->numAmpsPerNode, f, out, nullptr);
}

int qb_densmatr_calcFidelityWithPureState_sub(const qb_state* rho, const qb_state* psi, int conj, qb_cplx* out) {
    QB_READY(); QB_CHECK_DM(rho); QB_CHECK_STATE(psi); QB_REQUIRE(out, "null output");
    QB_REQUIRE(psi->numAmpsPerNode >= pow2(rho->numQubits), "fidelity: psi must hold the full pure state locally");
    RFidelity f; f.rho = (const cplx*)rho->amps; f.psi = (const cplx*)psi->amps;
    f.rankBits = (qindex)rho->rank << rho->logNumAmpsPerNode; f.numQubits = rho->numQubits; f.conj = conj;
    return qb_reduce2c(rho->numAmpsPerNode, f, out);
}

int qb_statevec_calcExpecAnyTargZ_sub(const qb_state* q, const int* targs, int nt, double* out) {
    QB_READY(); QB_CHECK_STATE(q); QB_CHECK_SUFFIX(targs, nt, q); QB_REQUIRE(out, "null output");
    RExpecZ f = {(const cplx*)q->amps, (qindex)qb_make_mask(targs, nt)};
    return qb_reduce2(q->numAmpsPerNode, f, out, nullptr);
}

int qb_densmatr_calcExpecAnyTargZ_sub(const qb_state* q, const int* targs, int nt, qb_cplx* out) {
    QB_READY(); QB_CHECK_DM(q); QB_REQUIRE(qb_check_qubits(targs, nt, q->numQubits) && out, "expecZ(dm): bad arguments");
    RExpecZDM f = {(const cplx*)q->amps, (qindex)qb_make_mask(targs, nt), firstDiagOf(q), diagStrideOf(q)};
    return qb_reduce2c(pow2(q->logNumColsPerNode), f, out);
}

static cplx powI(int n) {
    switch (n & 3) { case 0: return mk(1, 0); case 1: return mk(0, 1); case 2: return mk(-1, 0); default: return mk(0, -1); }
}

static int expecPauliSV(const qb_state* q, const cplx* other, const int* x, int nx, const int* y, int ny, const int* z, int nz, qb_cplx* out) {
    QB_READY(); QB_CHECK_STATE(q); QB_CHECK_SUFFIX(x, nx, q); QB_CHECK_SUFFIX(y, ny, q); QB_CHECK_SUFFIX(z, nz, q);
    QB_REQUIRE(out && other, "expecPauliStr: bad arguments");
    RExpecPauli f; f.amps = (const cplx*)q->amps; f.other = other;
    f.maskXY = (qindex)(qb_make_mask(x, nx) | qb_make_mask(y, ny)); f.maskYZ = (qindex)(qb_make_mask(y, ny) | qb_make_mask(z, nz));
    double re, im;
    int r = qb_reduce2(q->numAmpsPerNode, f, &re, &im);
    if (r) return r;
    cplx v = cmul(mk(re, im), powI(ny));
    out->re = v.x; out->im = v.y;
    return 0;
}

int qb_statevec_calcExpecPauliStr_subA(const qb_state* q, const int* x, int nx, const int* y, int ny, const int* z, int nz, qb_cplx* out) {
    QB_REQUIRE(q, "null state");
    return expecPauliSV(q, (const cplx*)q->amps, x, nx, y, ny, z, nz, out);
}

int qb_statevec_calcExpecPauliStr_subB(const qb_state* q, const int* x, int nx, const int* y, int ny, const int* z, int nz, qb_cplx* out) {
    QB_REQUIRE(q && q->buffer, "expecPauliStr subB: no communication buffer");
    return expecPauliSV(q, (const cplx*)q->buffer, x, nx, y, ny, z, nz, out);
}

int qb_densmatr_calcExpecPauliStr_sub(const qb_state* q, const int* x, int nx, const int* y, int ny, const int* z, int nz, qb_cplx* out) {
    QB_READY(); QB_CHECK_DM(q);
    QB_REQUIRE(qb_check_qubits(x, nx, q->numQubits) && qb_check_qubits(y, ny, q->numQubits) && qb_check_qubits(z, nz, q->numQubits) && out, "expecPauliStr(dm): bad arguments");
    RExpecPauliDM f; f.amps = (const cplx*)q->amps;
    f.maskXY = (qindex)(qb_make_mask(x, nx) | qb_make_mask(y, ny)); f.maskYZ = (qindex)(qb_make_mask(y, ny) | qb_make_mask(z, nz));
    f.firstDiag = firstDiagOf(q); f.dim = pow2(q->numQubits);
    double re, im;
    int r = qb_reduce2(pow2(q->logNumColsPerNode), f, &re, &im);
    if (r) return r;
    cplx v = cmul(mk(re, im), powI(ny));
    out->re = v.x; out->im = v.y;
    return 0;
}

int qb_statevec_calcExpecFullStateDiagMatr_sub(const qb_state* q, const qb_cplx* devElems, int hasPower, int realPow, qb_cplx expo, qb_cplx* out) {
    QB_READY(); QB_CHECK_STATE(q); QB_REQUIRE(devElems && out, "expecFullStateDiagMatr: bad arguments");
    RExpecDiag f; f.amps = (const cplx*)q->amps; f.elems = (const cplx*)devElems; f.isDM = 0; f.firstDiag = 0; f.stride = 0;
    f.hasPower = hasPower; f.realPow = realPow; f.expo = mk(expo);
    return qb_reduce2c(q->numAmpsPerNode, f, out);
}

int qb_densmatr_calcExpecFullStateDiagMatr_sub(const qb_state* q, const qb_cplx* devElems, int hasPower, int realPow, qb_cplx expo, qb_cplx* out) {
    QB_READY(); QB_CHECK_DM(q); QB_REQUIRE(devElems && out, "expecFullStateDiagMatr(dm): bad arguments");
    RExpecDiag f; f.amps = (const cplx*)q->amps; f.elems = (const cplx*)devElems; f.isDM = 1;
    f.firstDiag = firstDiagOf(q); f.stride = diagStrideOf(q);
    f.hasPower = hasPower; f.realPow = realPow; f.expo = mk(expo);
    return qb_reduce2c(pow2(q->logNumColsPerNode), f, out);
}

} // extern "C"

// Terms whose partner amplitudes live in the state itself are evaluated PG_K at a time from register-resident cosets
// (qb_pauli_group.cu): one read of the state per PG_K terms, where the kernel below re-streams the state once per term
// for the partners.  Terms are taken in order; a group closes when it is full or the next mask is linearly dependent
// on it; Z-only terms (mask 0) and whatever cannot be grouped go through the generic batch kernel.
static int pauliBatch(const qb_state* q, const cplx* other, const unsigned long long* masks, int numTerms, qb_cplx* outTerms);

static int pauliGrouped(const qb_state* q, const unsigned long long* masks, int numTerms, qb_cplx* outTerms) {
    QB_READY(); QB_CHECK_STATE(q); QB_REQUIRE(masks && outTerms && numTerms >= 0, "pauli batch: bad arguments");
    if (q->logNumAmpsPerNode < 12) return pauliBatch(q, (const cplx*)q->amps, masks, numTerms, outTerms);
    std::vector<int> rest;                                   // terms left to the generic kernel
    std::vector<std::vector<int>> groups;
    std::vector<int> cur; unsigned long long curMasks[PG_K];
    for (int t = 0; t < numTerms; t++) {
        const unsigned long long xy = masks[2 * t];
        QB_REQUIRE(xy < (unsigned long long)q->numAmpsPerNode, "pauli batch: X/Y mask reaches prefix qubits");
        if (!xy) { rest.push_back(t); continue; }
        curMasks[cur.size()] = xy;
        if (pg_rank(curMasks, (int)cur.size() + 1, nullptr) == (int)cur.size() + 1) cur.push_back(t);
        else {                                               // dependent on the open group: close it, start a new one
            if (!cur.empty()) groups.push_back(cur);
            cur.assign(1, t); curMasks[0] = xy;
        }
        if ((int)cur.size() == PG_K_EXPEC) { groups.push_back(cur); cur.clear(); }
    }
    if (!cur.empty()) groups.push_back(cur);
    // several groups per host synchronisation: each writes its 2k doubles to its own slot of the result area
    const int perSync = QB_RED_MAX_OUT / (2 * PG_K);
    for (size_t g0 = 0; g0 < groups.size(); g0 += perSync) {
        const size_t g1 = std::min(groups.size(), g0 + perSync);
        for (size_t g = g0; g < g1; g++) {
            unsigned long long m[2 * PG_K];
            for (size_t i = 0; i < groups[g].size(); i++) { m[2 * i] = masks[2 * groups[g][i]]; m[2 * i + 1] = masks[2 * groups[g][i] + 1]; }
            int r = qb_pauli_group_expec(q, m, (int)groups[g].size(), g_qb.redOutDev + (g - g0) * 2 * PG_K); if (r) return r;
        }
        QB_CUDA(cudaMemcpyAsync(g_qb.redOutHost, g_qb.redOutDev, (g1 - g0) * 2 * PG_K * sizeof(double), cudaMemcpyDeviceToHost, g_qb.stream));
        QB_CUDA(cudaStreamSynchronize(g_qb.stream));
        for (size_t g = g0; g < g1; g++)
            for (size_t i = 0; i < groups[g].size(); i++) {
                outTerms[groups[g][i]].re = g_qb.redOutHost[(g - g0) * 2 * PG_K + 2 * i];
                outTerms[groups[g][i]].im = g_qb.redOutHost[(g - g0) * 2 * PG_K + 2 * i + 1];
            }
    }
    if (!rest.empty()) {
        std::vector<unsigned long long> m; std::vector<qb_cplx> o(rest.size());
        for (int t : rest) { m.push_back(masks[2 * t]); m.push_back(masks[2 * t + 1]); }
        int r = pauliBatch(q, (const cplx*)q->amps, m.data(), (int)rest.size(), o.data()); if (r) return r;
        for (size_t i = 0; i < rest.size(); i++) outTerms[rest[i]] = o[i];
    }
    return 0;
}

static int pauliBatch(const qb_state* q, const cplx* other, const unsigned long long* masks, int numTerms, qb_cplx* outTerms) {
    QB_READY(); QB_CHECK_STATE(q); QB_REQUIRE(masks && outTerms && other && numTerms >= 0, "pauli batch: bad arguments");
    qindex blocks = (q->numAmpsPerNode + QB_BLOCK - 1) / QB_BLOCK;
    qindex maxBlocks = (qindex)g_qb.numSMs * 4;
    if (maxBlocks > QB_RED_MAX_BLOCKS) maxBlocks = QB_RED_MAX_BLOCKS;
    if (blocks > maxBlocks) blocks = maxBlocks;
    for (int base = 0; base < numTerms; base += QB_PAULI_BATCH) {
        PauliBatch pb; pb.amps = (const cplx*)q->amps; pb.other = other;
        pb.count = numTerms - base < QB_PAULI_BATCH ? numTerms - base : QB_PAULI_BATCH;
        for (int t = 0; t < QB_PAULI_BATCH; t++) {
            bool live = t < pb.count;
            pb.maskXY[t] = live ? (qindex)masks[2 * (base + t)] : 0;
            pb.maskYZ[t] = live ? (qindex)masks[2 * (base + t) + 1] : 0;
            QB_REQUIRE(!live || pb.maskXY[t] < q->numAmpsPerNode, "pauli batch: X/Y mask reaches prefix qubits");
        }
        k_pauliBatch<<<(unsigned int)blocks, QB_BLOCK, 0, g_qb.stream>>>(q->numAmpsPerNode, pb, g_qb.redPartials,
                                                                      g_qb.redTicket, g_qb.redOutDev);
        QB_LAUNCH_CHECK();
        QB_CUDA(cudaMemcpyAsync(g_qb.redOutHost, g_qb.redOutDev, 2 * QB_PAULI_BATCH * sizeof(double), cudaMemcpyDeviceToHost, g_qb.stream));
        QB_CUDA(cudaStreamSynchronize(g_qb.stream));
        for (int t = 0; t < pb.count; t++) { outTerms[base + t].re = g_qb.redOutHost[2 * t]; outTerms[base + t].im = g_qb.redOutHost[2 * t + 1]; }
    }
    return 0;
}

extern "C" {

int qb_statevec_calcExpecPauliStrBatch_subA(const qb_state* q, const unsigned long long* masks, int numTerms, qb_cplx* outTerms) {
    QB_REQUIRE(q, "null state");
    return pauliGrouped(q, masks, numTerms, outTerms);
}

int qb_statevec_calcExpecPauliStrBatch_subB(const qb_state* q, const unsigned long long* masks, int numTerms, qb_cplx* outTerms) {
    QB_REQUIRE(q && q->buffer, "pauli batch subB: no communication buffer");
    return pauliBatch(q, (const cplx*)q->buffer, masks, numTerms, outTerms);
}

} // extern "C"
