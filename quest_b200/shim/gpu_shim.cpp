/* gpu_shim.cpp -- defines the reference's GPU backend symbols on top of the quest_b200 C ABI.
 *
 * QuEST's dispatch layer (quest/src/core/accelerator.cpp:71-181) takes the ADDRESS of every explicit
 * template instantiation of the gpu_* routines declared in quest/src/gpu/gpu_subroutines.hpp:24-196,
 * and the api/ + core/ layers call the C++ functions of quest/src/gpu/gpu_config.hpp:41-120.  This
 * translation unit provides exactly those symbols (instantiated with the reference's own INSTANTIATE_*
 * macros, core/accelerator.hpp:45-126) and forwards each to one `extern "C"` entry point of
 * include/quest_b200.h, turning template parameters into run-time ints and C++ containers into
 * pointer + length.  It replaces quest/src/gpu/gpu_subroutines.cpp and quest/src/gpu/gpu_config.cpp;
 * it contains no arithmetic on amplitudes.
 *
 * Error convention: the C ABI returns a status; anything non-zero is routed into the reference's
 * internal-error path (core/errors.cpp:41-52 -> print + exit), exactly like CUDA_CHECK in the
 * reference (gpu/gpu_config.hpp:28-31).  Device OOM in gpu_allocArray is the one soft error (nullptr).
 */
#include "quest/include/modes.h"
#include "quest/include/types.h"
#include "quest/include/qureg.h"
#include "quest/include/paulis.h"
#include "quest/include/matrices.h"
#include "quest/include/channels.h"
#include "quest/include/environment.h"

#include "quest/src/core/errors.hpp"
#include "quest/src/core/memory.hpp"
#include "quest/src/core/utilities.hpp"
#include "quest/src/core/randomiser.hpp"
#include "quest/src/core/accelerator.hpp"
#include "quest/src/comm/comm_config.hpp"
#include "quest/src/comm/comm_routines.hpp"
#include "quest/src/gpu/gpu_config.hpp"
#include "quest/src/gpu/gpu_subroutines.hpp"

#include "quest_b200.h"
#include "qubit_map.hpp"

#include <array>
#include <set>
#include <string>
#include <vector>

using std::vector;

static_assert(sizeof(qcomp) == sizeof(qb_cplx), "quest_b200 is an fp64 (FLOAT_PRECISION=2) backend");
static_assert(sizeof(qindex) == sizeof(qb_index), "qindex mismatch");
static_assert(sizeof(PauliStr) == 2 * sizeof(unsigned long long), "PauliStr layout");


/*
 * helpers
 */

#define QB_CHECK(call) qbCheck((call), #call, __func__, __FILE__, __LINE__)

static void qbCheck(int status, const char* call, const char* caller, const char* file, int line) {
    if (status != 0)
        error_cudaCallFailed(qb_error_string(), call, caller, file, line);
}

// declared by gpu_config.hpp when COMPILE_CUDA=1; comm/ and core/ never call it but keep the symbol
void assertCudaCallSucceeded(int code, const char* call, const char* caller, const char* file, int line) {
    if (code != 0)
        error_cudaCallFailed("CUDA call failed", call, caller, file, line);
}

static qb_state st(Qureg q) {
    qb_state s;
    s.amps = reinterpret_cast<qb_cplx*>(q.gpuAmps);
    s.buffer = reinterpret_cast<qb_cplx*>(q.gpuCommBuffer);
    s.numAmpsPerNode = q.numAmpsPerNode;
    s.logNumAmpsPerNode = (int) q.logNumAmpsPerNode;
    s.rank = q.rank;
    s.numQubits = q.numQubits;
    s.logNumColsPerNode = (int) q.logNumColsPerNode;
    s.isDensityMatrix = q.isDensityMatrix;
    return s;
}

static qb_cplx qc(qcomp c) { return qb_cplx{ std::real(c), std::imag(c) }; }
static qcomp cq(qb_cplx c) { return qcomp(c.re, c.im); }
static const qb_cplx* qp(const qcomp* p) { return reinterpret_cast<const qb_cplx*>(p); }
static qb_cplx* qp(qcomp* p) { return reinterpret_cast<qb_cplx*>(p); }


/*
 * HARDWARE AVAILABILITY (gpu_config.cpp:124-283)
 */

static bool hasGpuBeenBound = false;

bool gpu_isGpuCompiled() { return true; }

bool gpu_isCuQuantumCompiled() { return false; }

int gpu_getNumberOfLocalGpus() { return qb_num_devices(); }

bool gpu_isGpuAvailable() { return qb_is_device_available() != 0; }

bool gpu_isDirectGpuCommPossible() {
    // amplitudes always travel GPU-to-GPU over NCCL/NVLink; there is no host-staged path
    return gpu_isGpuAvailable();
}

int gpu_getComputeCapability() {
    assert_gpuHasBeenBound(hasGpuBeenBound);
    return qb_compute_capability();
}

size_t gpu_getCurrentAvailableMemoryInBytes() {
    assert_gpuHasBeenBound(hasGpuBeenBound);
    size_t f = 0, t = 0;
    QB_CHECK( qb_mem_info(&f, &t) );
    return f;
}

size_t gpu_getTotalMemoryInBytes() {
    assert_gpuHasBeenBound(hasGpuBeenBound);
    size_t f = 0, t = 0;
    QB_CHECK( qb_mem_info(&f, &t) );
    return t;
}

bool gpu_doesGpuSupportMemPools() {
    assert_gpuHasBeenBound(hasGpuBeenBound);
    return qb_supports_mem_pools() != 0;
}

qindex gpu_getMaxNumConcurrentThreads() {
    assert_gpuHasBeenBound(hasGpuBeenBound);
    return qb_max_concurrent_threads();
}


/*
 * ENVIRONMENT MANAGEMENT (gpu_config.cpp:332-395)
 */

void gpu_bindLocalGPUsToNodes() {
    int numLocalGpus = gpu_getNumberOfLocalGpus();
    int localGpuInd = (numLocalGpus > 0)? comm_getRank() % numLocalGpus : 0;
    QB_CHECK( qb_bind_device(localGpuInd) );
    hasGpuBeenBound = true;
}

bool gpu_areAnyNodesBoundToSameGpu() {
    assert_gpuHasBeenBound(hasGpuBeenBound);

    if (!comm_isInit())
        return false;

    // the shared-memory / CUDA-IPC transport exists precisely so that ranks may share a device (it is only chosen
    // when there are fewer GPUs than ranks, or on request): the run-time twin of PERMIT_NODES_TO_SHARE_GPU
    // (api/environment.cpp:110, core/validation.cpp:1365)
    if (qb_comm_transport() == 1)
        return false;

    // 16 raw UUID bytes are hex-encoded so that they survive the string gather intact
    char raw[16];
    QB_CHECK( qb_device_uuid(raw) );
    std::array<char,33> hex;
    static const char* digits = "0123456789abcdef";
    for (int i=0; i<16; i++) {
        hex[2*i]   = digits[(raw[i] >> 4) & 0xF];
        hex[2*i+1] = digits[ raw[i]       & 0xF];
    }
    hex[32] = '\0';

    auto allUuids = comm_gatherStringsToRoot(hex.data(), (int) hex.size());
    auto uniqueUuids = std::set<std::string>(allUuids.begin(), allUuids.end());
    bool localGpusAreUnique = allUuids.size() == uniqueUuids.size();
    bool globalGpusAreUnique = comm_isTrueOnRootNode(localGpusAreUnique);
    return ! globalGpusAreUnique;
}

void gpu_sync() { qbmap_canonicaliseAll(); QB_CHECK( qb_sync() ); }     // syncQuESTEnv(): restore canonical qubit order first

void gpu_initCuQuantum()     { error_cuQuantumInitOrFinalizedButNotCompiled(); }
void gpu_finalizeCuQuantum() { error_cuQuantumInitOrFinalizedButNotCompiled(); }


/*
 * MEMORY MANAGEMENT (gpu_config.cpp:399-623)
 */

qcomp* gpu_allocArray(qindex length) {
    int status = 0;
    qb_cplx* ptr = qb_alloc(length, &status);
    QB_CHECK( status );
    return reinterpret_cast<qcomp*>(ptr);   // nullptr on out-of-memory, handled by validation
}

void gpu_deallocArray(qcomp* amps) { qbmap_forget(amps); QB_CHECK( qb_free(qp(amps)) ); }

void gpu_copyArray(qcomp* dest, qcomp* src, qindex dim) { QB_CHECK( qb_copy_d2d(qp(dest), qp(src), dim) ); }

void gpu_copyCpuToGpu(qcomp* cpuArr, qcomp* gpuArr, qindex numElems) { qbmap_canonicaliseHolding(gpuArr); QB_CHECK( qb_copy_h2d(qp(gpuArr), qp(cpuArr), numElems) ); }
void gpu_copyGpuToCpu(qcomp* gpuArr, qcomp* cpuArr, qindex numElems) { qbmap_canonicaliseHolding(gpuArr); QB_CHECK( qb_copy_d2h(qp(cpuArr), qp(gpuArr), numElems) ); }

void gpu_copyCpuToGpu(Qureg qureg, qcomp* cpuArr, qcomp* gpuArr, qindex numElems) {
    assert_quregIsGpuAccelerated(qureg);
    gpu_copyCpuToGpu(cpuArr, gpuArr, numElems);
}
void gpu_copyGpuToCpu(Qureg qureg, qcomp* gpuArr, qcomp* cpuArr, qindex numElems) {
    assert_quregIsGpuAccelerated(qureg);
    gpu_copyGpuToCpu(gpuArr, cpuArr, numElems);
}

void gpu_copyCpuToGpu(Qureg qureg) { gpu_copyCpuToGpu(qureg, qureg.cpuAmps, qureg.gpuAmps, qureg.numAmpsPerNode); }
void gpu_copyGpuToCpu(Qureg qureg) { gpu_copyGpuToCpu(qureg, qureg.gpuAmps, qureg.cpuAmps, qureg.numAmpsPerNode); }

template <typename T>
static void assertHeapObjectGpuMemIsAllocated(T obj) {
    if (! mem_isAllocated(util_getGpuMemPtr(obj)) || ! getQuESTEnv().isGpuAccelerated)
        error_gpuCopyButMatrixNotGpuAccelerated();
}

void gpu_copyCpuToGpu(CompMatr matr) {
    assertHeapObjectGpuMemIsAllocated(matr);
    gpu_copyCpuToGpu(matr.cpuElemsFlat, matr.gpuElemsFlat, matr.numRows * matr.numRows);
}
void gpu_copyGpuToCpu(CompMatr matr) {
    assertHeapObjectGpuMemIsAllocated(matr);
    gpu_copyGpuToCpu(matr.gpuElemsFlat, matr.cpuElemsFlat, matr.numRows * matr.numRows);
}
void gpu_copyCpuToGpu(DiagMatr matr) {
    assertHeapObjectGpuMemIsAllocated(matr);
    gpu_copyCpuToGpu(matr.cpuElems, matr.gpuElems, matr.numElems);
}
void gpu_copyGpuToCpu(DiagMatr matr) {
    assertHeapObjectGpuMemIsAllocated(matr);
    gpu_copyGpuToCpu(matr.gpuElems, matr.cpuElems, matr.numElems);
}
void gpu_copyCpuToGpu(SuperOp op) {
    assertHeapObjectGpuMemIsAllocated(op);
    gpu_copyCpuToGpu(op.cpuElemsFlat, op.gpuElemsFlat, op.numRows * op.numRows);
}
void gpu_copyGpuToCpu(SuperOp op) {
    assertHeapObjectGpuMemIsAllocated(op);
    gpu_copyGpuToCpu(op.gpuElemsFlat, op.cpuElemsFlat, op.numRows * op.numRows);
}
void gpu_copyCpuToGpu(FullStateDiagMatr matr) {
    assertHeapObjectGpuMemIsAllocated(matr);
    gpu_copyCpuToGpu(matr.cpuElems, matr.gpuElems, matr.numElemsPerNode);
}


/*
 * CACHE MANAGEMENT (gpu_config.cpp:635-687)
 */

qcomp* gpu_getCacheOfSize(qindex numElemsPerThread, qindex numThreads) {
    int status = 0;
    qb_cplx* ptr = qb_get_cache(numElemsPerThread * numThreads, &status);
    QB_CHECK( status );
    return reinterpret_cast<qcomp*>(ptr);
}

void gpu_clearCache() { QB_CHECK( qb_clear_cache() ); }

size_t gpu_getCacheMemoryInBytes() { return qb_cache_bytes(); }


/*
 * GETTERS / SETTERS (gpu_subroutines.cpp:73-124)
 */

qcomp gpu_statevec_getAmp_sub(Qureg qureg, qindex ind) {
    auto s = st(qureg);
    qb_cplx out;
    QB_CHECK( qb_statevec_getAmp_sub(&s, ind, &out) );
    return cq(out);
}

void gpu_densmatr_setAmpsToPauliStrSum_sub(Qureg qureg, PauliStrSum sum) {
    assert_highPauliStrSumMaskIsZero(sum);
    auto s = st(qureg);
    QB_CHECK( qb_densmatr_setAmpsToPauliStrSum_sub(&s, qp(sum.coeffs),
        reinterpret_cast<const unsigned long long*>(sum.strings), sum.numTerms) );
}

void gpu_fullstatediagmatr_setElemsToPauliStrSum(FullStateDiagMatr out, PauliStrSum in) {
    int rank = out.isDistributed? comm_getRank() : 0;
    QB_CHECK( qb_fullstatediagmatr_setElemsToPauliStrSum(qp(out.gpuElems), out.numElemsPerNode, rank,
        qp(in.coeffs), reinterpret_cast<const unsigned long long*>(in.strings), in.numTerms) );
}


/*
 * COMMUNICATION BUFFER PACKING (gpu_subroutines.cpp:135-190)
 */

template <int NumQubits>
qindex gpu_statevec_packAmpsIntoBuffer(Qureg qureg, vector<int> qubits, vector<int> qubitStates) {
    assert_numQubitsMatchesQubitStatesAndTemplateParam(qubits.size(), qubitStates.size(), NumQubits);
    auto s = st(qureg);
    qb_index n = 0;
    QB_CHECK( qb_statevec_packAmpsIntoBuffer(&s, qubits.data(), qubitStates.data(), (int) qubits.size(), &n) );
    return n;
}

qindex gpu_statevec_packPairSummedAmpsIntoBuffer(Qureg qureg, int qubit1, int qubit2, int qubit3, int bit2) {
    assert_bufferPackerGivenIncreasingQubits(qubit1, qubit2, qubit3);
    auto s = st(qureg);
    qb_index n = 0;
    QB_CHECK( qb_statevec_packPairSummedAmpsIntoBuffer(&s, qubit1, qubit2, qubit3, bit2, &n) );
    return n;
}

INSTANTIATE_FUNC_OPTIMISED_FOR_NUM_TARGS( qindex, gpu_statevec_packAmpsIntoBuffer, (Qureg, vector<int>, vector<int>) )


/*
 * SWAPS (gpu_subroutines.cpp:198-280)
 */

template <int NumCtrls>
void gpu_statevec_anyCtrlSwap_subA(Qureg qureg, vector<int> ctrls, vector<int> ctrlStates, int targ1, int targ2) {
    assert_numCtrlsMatchesNumCtrlStatesAndTemplateParam(ctrls.size(), ctrlStates.size(), NumCtrls);
    auto s = st(qureg);
    QB_CHECK( qb_statevec_anyCtrlSwap_subA(&s, ctrls.data(), ctrlStates.data(), (int) ctrls.size(), targ1, targ2) );
}

template <int NumCtrls>
void gpu_statevec_anyCtrlSwap_subB(Qureg qureg, vector<int> ctrls, vector<int> ctrlStates) {
    assert_numCtrlsMatchesNumCtrlStatesAndTemplateParam(ctrls.size(), ctrlStates.size(), NumCtrls);
    auto s = st(qureg);
    QB_CHECK( qb_statevec_anyCtrlSwap_subB(&s, ctrls.data(), ctrlStates.data(), (int) ctrls.size()) );
}

template <int NumCtrls>
void gpu_statevec_anyCtrlSwap_subC(Qureg qureg, vector<int> ctrls, vector<int> ctrlStates, int targ, int targState) {
    assert_numCtrlsMatchesNumCtrlStatesAndTemplateParam(ctrls.size(), ctrlStates.size(), NumCtrls);
    auto s = st(qureg);
    QB_CHECK( qb_statevec_anyCtrlSwap_subC(&s, ctrls.data(), ctrlStates.data(), (int) ctrls.size(), targ, targState) );
}

INSTANTIATE_FUNC_OPTIMISED_FOR_NUM_CTRLS( void, gpu_statevec_anyCtrlSwap_subA, (Qureg, vector<int>, vector<int>, int, int) )
INSTANTIATE_FUNC_OPTIMISED_FOR_NUM_CTRLS( void, gpu_statevec_anyCtrlSwap_subB, (Qureg, vector<int>, vector<int>) )
INSTANTIATE_FUNC_OPTIMISED_FOR_NUM_CTRLS( void, gpu_statevec_anyCtrlSwap_subC, (Qureg, vector<int>, vector<int>, int, int) )


/*
 * DENSE MATRICES (gpu_subroutines.cpp:287-510)
 */

template <int NumCtrls>
void gpu_statevec_anyCtrlOneTargDenseMatr_subA(Qureg qureg, vector<int> ctrls, vector<int> ctrlStates, int targ, CompMatr1 matr) {
    assert_numCtrlsMatchesNumCtrlStatesAndTemplateParam(ctrls.size(), ctrlStates.size(), NumCtrls);
    auto s = st(qureg);
    QB_CHECK( qb_statevec_anyCtrlOneTargDenseMatr_subA(&s, ctrls.data(), ctrlStates.data(), (int) ctrls.size(), targ, qp(&matr.elems[0][0])) );
}

template <int NumCtrls>
void gpu_statevec_anyCtrlOneTargDenseMatr_subB(Qureg qureg, vector<int> ctrls, vector<int> ctrlStates, qcomp fac0, qcomp fac1) {
    assert_numCtrlsMatchesNumCtrlStatesAndTemplateParam(ctrls.size(), ctrlStates.size(), NumCtrls);
    auto s = st(qureg);
    QB_CHECK( qb_statevec_anyCtrlOneTargDenseMatr_subB(&s, ctrls.data(), ctrlStates.data(), (int) ctrls.size(), qc(fac0), qc(fac1)) );
}

template <int NumCtrls>
void gpu_statevec_anyCtrlTwoTargDenseMatr_sub(Qureg qureg, vector<int> ctrls, vector<int> ctrlStates, int targ1, int targ2, CompMatr2 matr) {
    assert_numCtrlsMatchesNumCtrlStatesAndTemplateParam(ctrls.size(), ctrlStates.size(), NumCtrls);
    auto s = st(qureg);
    QB_CHECK( qb_statevec_anyCtrlTwoTargDenseMatr_sub(&s, ctrls.data(), ctrlStates.data(), (int) ctrls.size(), targ1, targ2, qp(&matr.elems[0][0])) );
}

template <int NumCtrls, int NumTargs, bool ApplyConj>
void gpu_statevec_anyCtrlAnyTargDenseMatr_sub(Qureg qureg, vector<int> ctrls, vector<int> ctrlStates, vector<int> targs, CompMatr matr) {
    assert_numCtrlsMatchesNumCtrlStatesAndTemplateParam(ctrls.size(), ctrlStates.size(), NumCtrls);
    assert_numTargsMatchesTemplateParam(targs.size(), NumTargs);
    auto s = st(qureg);
    QB_CHECK( qb_statevec_anyCtrlAnyTargDenseMatr_sub(&s, ctrls.data(), ctrlStates.data(), (int) ctrls.size(),
        targs.data(), (int) targs.size(), qp(matr.gpuElemsFlat), ApplyConj) );
}

INSTANTIATE_FUNC_OPTIMISED_FOR_NUM_CTRLS( void, gpu_statevec_anyCtrlOneTargDenseMatr_subA, (Qureg, vector<int>, vector<int>, int, CompMatr1) )
INSTANTIATE_FUNC_OPTIMISED_FOR_NUM_CTRLS( void, gpu_statevec_anyCtrlOneTargDenseMatr_subB, (Qureg, vector<int>, vector<int>, qcomp, qcomp) )
INSTANTIATE_FUNC_OPTIMISED_FOR_NUM_CTRLS( void, gpu_statevec_anyCtrlTwoTargDenseMatr_sub, (Qureg, vector<int>, vector<int>, int, int, CompMatr2) )
INSTANTIATE_CONJUGABLE_FUNC_OPTIMISED_FOR_NUM_CTRLS_AND_TARGS( void, gpu_statevec_anyCtrlAnyTargDenseMatr_sub, (Qureg, vector<int>, vector<int>, vector<int>, CompMatr) )


/*
 * DIAGONAL MATRICES (gpu_subroutines.cpp:517-775)
 */

template <int NumCtrls>
void gpu_statevec_anyCtrlOneTargDiagMatr_sub(Qureg qureg, vector<int> ctrls, vector<int> ctrlStates, int targ, DiagMatr1 matr) {
    assert_numCtrlsMatchesNumCtrlStatesAndTemplateParam(ctrls.size(), ctrlStates.size(), NumCtrls);
    auto s = st(qureg);
    QB_CHECK( qb_statevec_anyCtrlOneTargDiagMatr_sub(&s, ctrls.data(), ctrlStates.data(), (int) ctrls.size(), targ, qp(matr.elems)) );
}

template <int NumCtrls>
void gpu_statevec_anyCtrlTwoTargDiagMatr_sub(Qureg qureg, vector<int> ctrls, vector<int> ctrlStates, int targ1, int targ2, DiagMatr2 matr) {
    assert_numCtrlsMatchesNumCtrlStatesAndTemplateParam(ctrls.size(), ctrlStates.size(), NumCtrls);
    auto s = st(qureg);
    QB_CHECK( qb_statevec_anyCtrlTwoTargDiagMatr_sub(&s, ctrls.data(), ctrlStates.data(), (int) ctrls.size(), targ1, targ2, qp(matr.elems)) );
}

template <int NumCtrls, int NumTargs, bool ApplyConj, bool HasPower>
void gpu_statevec_anyCtrlAnyTargDiagMatr_sub(Qureg qureg, vector<int> ctrls, vector<int> ctrlStates, vector<int> targs, DiagMatr matr, qcomp exponent) {
    assert_numCtrlsMatchesNumCtrlStatesAndTemplateParam(ctrls.size(), ctrlStates.size(), NumCtrls);
    assert_numTargsMatchesTemplateParam(targs.size(), NumTargs);
    assert_exponentMatchesTemplateParam(exponent, HasPower);
    auto s = st(qureg);
    QB_CHECK( qb_statevec_anyCtrlAnyTargDiagMatr_sub(&s, ctrls.data(), ctrlStates.data(), (int) ctrls.size(),
        targs.data(), (int) targs.size(), qp(matr.gpuElems), ApplyConj, HasPower, qc(exponent)) );
}

template <bool HasPower>
void gpu_statevec_allTargDiagMatr_sub(Qureg qureg, FullStateDiagMatr matr, qcomp exponent) {
    assert_exponentMatchesTemplateParam(exponent, HasPower);
    auto s = st(qureg);
    QB_CHECK( qb_statevec_allTargDiagMatr_sub(&s, qp(matr.gpuElems), HasPower, qc(exponent)) );
}

template <bool HasPower, bool MultiplyOnly>
void gpu_densmatr_allTargDiagMatr_sub(Qureg qureg, FullStateDiagMatr matr, qcomp exponent) {
    assert_exponentMatchesTemplateParam(exponent, HasPower);
    auto s = st(qureg);
    QB_CHECK( qb_densmatr_allTargDiagMatr_sub(&s, qp(matr.gpuElems), matr.numElems, HasPower, MultiplyOnly, qc(exponent)) );
}

INSTANTIATE_FUNC_OPTIMISED_FOR_NUM_CTRLS( void, gpu_statevec_anyCtrlOneTargDiagMatr_sub, (Qureg, vector<int>, vector<int>, int, DiagMatr1) )
INSTANTIATE_FUNC_OPTIMISED_FOR_NUM_CTRLS( void, gpu_statevec_anyCtrlTwoTargDiagMatr_sub, (Qureg, vector<int>, vector<int>, int, int, DiagMatr2) )
INSTANTIATE_EXPONENTIABLE_CONJUGABLE_FUNC_OPTIMISED_FOR_NUM_CTRLS_AND_TARGS( void, gpu_statevec_anyCtrlAnyTargDiagMatr_sub, (Qureg, vector<int>, vector<int>, vector<int>, DiagMatr, qcomp) )

template void gpu_statevec_allTargDiagMatr_sub<true >(Qureg, FullStateDiagMatr, qcomp);
template void gpu_statevec_allTargDiagMatr_sub<false>(Qureg, FullStateDiagMatr, qcomp);
template void gpu_densmatr_allTargDiagMatr_sub<true, true>  (Qureg, FullStateDiagMatr, qcomp);
template void gpu_densmatr_allTargDiagMatr_sub<true, false> (Qureg, FullStateDiagMatr, qcomp);
template void gpu_densmatr_allTargDiagMatr_sub<false, true> (Qureg, FullStateDiagMatr, qcomp);
template void gpu_densmatr_allTargDiagMatr_sub<false, false>(Qureg, FullStateDiagMatr, qcomp);


/*
 * PAULI TENSOR AND GADGET (gpu_subroutines.cpp:780-895)
 */

template <int NumCtrls, int NumTargs>
void gpu_statevector_anyCtrlPauliTensorOrGadget_subA(Qureg qureg, vector<int> ctrls, vector<int> ctrlStates, vector<int> x, vector<int> y, vector<int> z, qcomp ampFac, qcomp pairAmpFac) {
    assert_numCtrlsMatchesNumCtrlStatesAndTemplateParam(ctrls.size(), ctrlStates.size(), NumCtrls);
    assert_numTargsMatchesTemplateParam(x.size() + y.size(), NumTargs);
    auto s = st(qureg);
    QB_CHECK( qb_statevector_anyCtrlPauliTensorOrGadget_subA(&s, ctrls.data(), ctrlStates.data(), (int) ctrls.size(),
        x.data(), (int) x.size(), y.data(), (int) y.size(), z.data(), (int) z.size(), qc(ampFac), qc(pairAmpFac)) );
}

template <int NumCtrls>
void gpu_statevector_anyCtrlPauliTensorOrGadget_subB(Qureg qureg, vector<int> ctrls, vector<int> ctrlStates, vector<int> x, vector<int> y, vector<int> z, qcomp ampFac, qcomp pairAmpFac, qindex bufferMaskXY) {
    assert_numCtrlsMatchesNumCtrlStatesAndTemplateParam(ctrls.size(), ctrlStates.size(), NumCtrls);
    auto s = st(qureg);
    QB_CHECK( qb_statevector_anyCtrlPauliTensorOrGadget_subB(&s, ctrls.data(), ctrlStates.data(), (int) ctrls.size(),
        x.data(), (int) x.size(), y.data(), (int) y.size(), z.data(), (int) z.size(), qc(ampFac), qc(pairAmpFac), bufferMaskXY) );
}

template <int NumCtrls>
void gpu_statevector_anyCtrlAnyTargZOrPhaseGadget_sub(Qureg qureg, vector<int> ctrls, vector<int> ctrlStates, vector<int> targs, qcomp fac0, qcomp fac1) {
    assert_numCtrlsMatchesNumCtrlStatesAndTemplateParam(ctrls.size(), ctrlStates.size(), NumCtrls);
    auto s = st(qureg);
    QB_CHECK( qb_statevector_anyCtrlAnyTargZOrPhaseGadget_sub(&s, ctrls.data(), ctrlStates.data(), (int) ctrls.size(),
        targs.data(), (int) targs.size(), qc(fac0), qc(fac1)) );
}

INSTANTIATE_FUNC_OPTIMISED_FOR_NUM_CTRLS_AND_TARGS( void, gpu_statevector_anyCtrlPauliTensorOrGadget_subA, (Qureg, vector<int>, vector<int>, vector<int>, vector<int>, vector<int>, qcomp, qcomp) )
INSTANTIATE_FUNC_OPTIMISED_FOR_NUM_CTRLS( void, gpu_statevector_anyCtrlPauliTensorOrGadget_subB, (Qureg, vector<int>, vector<int>, vector<int>, vector<int>, vector<int>, qcomp, qcomp, qindex) )
INSTANTIATE_FUNC_OPTIMISED_FOR_NUM_CTRLS( void, gpu_statevector_anyCtrlAnyTargZOrPhaseGadget_sub, (Qureg, vector<int>, vector<int>, vector<int>, qcomp, qcomp) )


/*
 * QUREG COMBINATION (gpu_subroutines.cpp:899-960)
 */

void gpu_statevec_setQuregToSuperposition_sub(qcomp facOut, Qureg outQureg, qcomp fac1, Qureg inQureg1, qcomp fac2, Qureg inQureg2) {
    assert_superposedQuregDimsAndDeploysMatch(outQureg, inQureg1, inQureg2);
    auto o = st(outQureg), a = st(inQureg1), b = st(inQureg2);
    QB_CHECK( qb_statevec_setQuregToSuperposition_sub(qc(facOut), &o, qc(fac1), &a, qc(fac2), &b) );
}

void gpu_densmatr_mixQureg_subA(qreal outProb, Qureg outQureg, qreal inProb, Qureg inDensMatr) {
    auto o = st(outQureg), a = st(inDensMatr);
    QB_CHECK( qb_densmatr_mixQureg_subA(outProb, &o, inProb, &a) );
}

void gpu_densmatr_mixQureg_subB(qreal outProb, Qureg outQureg, qreal inProb, Qureg inStateVec) {
    auto o = st(outQureg), a = st(inStateVec);
    QB_CHECK( qb_densmatr_mixQureg_subB(outProb, &o, inProb, &a) );
}

void gpu_densmatr_mixQureg_subC(qreal outProb, Qureg outQureg, qreal inProb) {
    auto o = st(outQureg);
    QB_CHECK( qb_densmatr_mixQureg_subC(outProb, &o, inProb) );
}


/*
 * DECOHERENCE (gpu_subroutines.cpp:965-1390)
 */

#define QB_CHANNEL(name) \
    void gpu_densmatr_##name(Qureg qureg, int qubit, qreal prob) { \
        auto s = st(qureg); \
        QB_CHECK( qb_densmatr_##name(&s, qubit, prob) ); \
    }
#define QB_CHANNEL2(name) \
    void gpu_densmatr_##name(Qureg qureg, int qubit1, int qubit2, qreal prob) { \
        auto s = st(qureg); \
        QB_CHECK( qb_densmatr_##name(&s, qubit1, qubit2, prob) ); \
    }

QB_CHANNEL(oneQubitDephasing_subA)
QB_CHANNEL(oneQubitDephasing_subB)
QB_CHANNEL2(twoQubitDephasing_subA)
QB_CHANNEL2(twoQubitDephasing_subB)
QB_CHANNEL(oneQubitDepolarising_subA)
QB_CHANNEL(oneQubitDepolarising_subB)
QB_CHANNEL2(twoQubitDepolarising_subA)
QB_CHANNEL2(twoQubitDepolarising_subB)
QB_CHANNEL2(twoQubitDepolarising_subC)
QB_CHANNEL2(twoQubitDepolarising_subD)
QB_CHANNEL2(twoQubitDepolarising_subE)
QB_CHANNEL2(twoQubitDepolarising_subF)
QB_CHANNEL(oneQubitDamping_subA)
QB_CHANNEL(oneQubitDamping_subB)
QB_CHANNEL(oneQubitDamping_subC)
QB_CHANNEL(oneQubitDamping_subD)

void gpu_densmatr_oneQubitPauliChannel_subA(Qureg qureg, int ketQubit, qreal pI, qreal pX, qreal pY, qreal pZ) {
    auto s = st(qureg);
    QB_CHECK( qb_densmatr_oneQubitPauliChannel_subA(&s, ketQubit, pI, pX, pY, pZ) );
}

void gpu_densmatr_oneQubitPauliChannel_subB(Qureg qureg, int ketQubit, qreal pI, qreal pX, qreal pY, qreal pZ) {
    auto s = st(qureg);
    QB_CHECK( qb_densmatr_oneQubitPauliChannel_subB(&s, ketQubit, pI, pX, pY, pZ) );
}


/*
 * PARTIAL TRACE (gpu_subroutines.cpp:1398-1425)
 */

template <int NumTargs>
void gpu_densmatr_partialTrace_sub(Qureg inQureg, Qureg outQureg, vector<int> targs, vector<int> pairTargs) {
    assert_numTargsMatchesTemplateParam(targs.size(), NumTargs);
    auto i = st(inQureg), o = st(outQureg);
    QB_CHECK( qb_densmatr_partialTrace_sub(&i, &o, targs.data(), pairTargs.data(), (int) targs.size()) );
}

INSTANTIATE_FUNC_OPTIMISED_FOR_NUM_TARGS( void, gpu_densmatr_partialTrace_sub, (Qureg, Qureg, vector<int>, vector<int>) )


/*
 * PROBABILITIES (gpu_subroutines.cpp:1431-1590)
 */

qreal gpu_statevec_calcTotalProb_sub(Qureg qureg) {
    auto s = st(qureg);
    double out = 0;      // reductions come back in double in both precisions; narrowed to qreal on return
    QB_CHECK( qb_statevec_calcTotalProb_sub(&s, &out) );
    return out;
}

qreal gpu_densmatr_calcTotalProb_sub(Qureg qureg) {
    auto s = st(qureg);
    double out = 0;      // reductions come back in double in both precisions; narrowed to qreal on return
    QB_CHECK( qb_densmatr_calcTotalProb_sub(&s, &out) );
    return out;
}

template <int NumQubits>
qreal gpu_statevec_calcProbOfMultiQubitOutcome_sub(Qureg qureg, vector<int> qubits, vector<int> outcomes) {
    assert_numTargsMatchesTemplateParam(qubits.size(), NumQubits);
    auto s = st(qureg);
    double out = 0;      // reductions come back in double in both precisions; narrowed to qreal on return
    QB_CHECK( qb_statevec_calcProbOfMultiQubitOutcome_sub(&s, qubits.data(), outcomes.data(), (int) qubits.size(), &out) );
    return out;
}

template <int NumQubits>
qreal gpu_densmatr_calcProbOfMultiQubitOutcome_sub(Qureg qureg, vector<int> qubits, vector<int> outcomes) {
    assert_numTargsMatchesTemplateParam(qubits.size(), NumQubits);
    auto s = st(qureg);
    double out = 0;      // reductions come back in double in both precisions; narrowed to qreal on return
    QB_CHECK( qb_densmatr_calcProbOfMultiQubitOutcome_sub(&s, qubits.data(), outcomes.data(), (int) qubits.size(), &out) );
    return out;
}

template <int NumQubits>
void gpu_statevec_calcProbsOfAllMultiQubitOutcomes_sub(qreal* outProbs, Qureg qureg, vector<int> qubits) {
    assert_numTargsMatchesTemplateParam(qubits.size(), NumQubits);
    auto s = st(qureg);
    vector<double> probs((size_t) 1 << qubits.size());
    QB_CHECK( qb_statevec_calcProbsOfAllMultiQubitOutcomes_sub(probs.data(), &s, qubits.data(), (int) qubits.size()) );
    for (size_t i = 0; i < probs.size(); i++)
        outProbs[i] = (qreal) probs[i];
}

template <int NumQubits>
void gpu_densmatr_calcProbsOfAllMultiQubitOutcomes_sub(qreal* outProbs, Qureg qureg, vector<int> qubits) {
    assert_numTargsMatchesTemplateParam(qubits.size(), NumQubits);
    auto s = st(qureg);
    vector<double> probs((size_t) 1 << qubits.size());
    QB_CHECK( qb_densmatr_calcProbsOfAllMultiQubitOutcomes_sub(probs.data(), &s, qubits.data(), (int) qubits.size()) );
    for (size_t i = 0; i < probs.size(); i++)
        outProbs[i] = (qreal) probs[i];
}

INSTANTIATE_FUNC_OPTIMISED_FOR_NUM_TARGS( qreal, gpu_statevec_calcProbOfMultiQubitOutcome_sub, (Qureg, vector<int>, vector<int>) )
INSTANTIATE_FUNC_OPTIMISED_FOR_NUM_TARGS( qreal, gpu_densmatr_calcProbOfMultiQubitOutcome_sub, (Qureg, vector<int>, vector<int>) )
INSTANTIATE_FUNC_OPTIMISED_FOR_NUM_TARGS( void, gpu_statevec_calcProbsOfAllMultiQubitOutcomes_sub, (qreal* outProbs, Qureg, vector<int>) )
INSTANTIATE_FUNC_OPTIMISED_FOR_NUM_TARGS( void, gpu_densmatr_calcProbsOfAllMultiQubitOutcomes_sub, (qreal* outProbs, Qureg, vector<int>) )


/*
 * INNER PRODUCTS (gpu_subroutines.cpp:1597-1645)
 */

qcomp gpu_statevec_calcInnerProduct_sub(Qureg quregA, Qureg quregB) {
    auto a = st(quregA), b = st(quregB);
    qb_cplx out;
    QB_CHECK( qb_statevec_calcInnerProduct_sub(&a, &b, &out) );
    return cq(out);
}

qreal gpu_densmatr_calcHilbertSchmidtDistance_sub(Qureg quregA, Qureg quregB) {
    auto a = st(quregA), b = st(quregB);
    double out = 0;      // reductions come back in double in both precisions; narrowed to qreal on return
    QB_CHECK( qb_densmatr_calcHilbertSchmidtDistance_sub(&a, &b, &out) );
    return out;
}

template <bool Conj>
qcomp gpu_densmatr_calcFidelityWithPureState_sub(Qureg rho, Qureg psi) {
    auto r = st(rho), p = st(psi);
    qb_cplx out;
    QB_CHECK( qb_densmatr_calcFidelityWithPureState_sub(&r, &p, Conj, &out) );
    return cq(out);
}

template qcomp gpu_densmatr_calcFidelityWithPureState_sub<true >(Qureg, Qureg);
template qcomp gpu_densmatr_calcFidelityWithPureState_sub<false>(Qureg, Qureg);


/*
 * EXPECTATION VALUES (gpu_subroutines.cpp:1649-1775)
 */

qreal gpu_statevec_calcExpecAnyTargZ_sub(Qureg qureg, vector<int> targs) {
    auto s = st(qureg);
    double out = 0;      // reductions come back in double in both precisions; narrowed to qreal on return
    QB_CHECK( qb_statevec_calcExpecAnyTargZ_sub(&s, targs.data(), (int) targs.size(), &out) );
    return out;
}

qcomp gpu_densmatr_calcExpecAnyTargZ_sub(Qureg qureg, vector<int> targs) {
    auto s = st(qureg);
    qb_cplx out;
    QB_CHECK( qb_densmatr_calcExpecAnyTargZ_sub(&s, targs.data(), (int) targs.size(), &out) );
    return cq(out);
}

qcomp gpu_statevec_calcExpecPauliStr_subA(Qureg qureg, vector<int> x, vector<int> y, vector<int> z) {
    auto s = st(qureg);
    qb_cplx out;
    QB_CHECK( qb_statevec_calcExpecPauliStr_subA(&s, x.data(), (int) x.size(), y.data(), (int) y.size(), z.data(), (int) z.size(), &out) );
    return cq(out);
}

qcomp gpu_statevec_calcExpecPauliStr_subB(Qureg qureg, vector<int> x, vector<int> y, vector<int> z) {
    auto s = st(qureg);
    qb_cplx out;
    QB_CHECK( qb_statevec_calcExpecPauliStr_subB(&s, x.data(), (int) x.size(), y.data(), (int) y.size(), z.data(), (int) z.size(), &out) );
    return cq(out);
}

qcomp gpu_densmatr_calcExpecPauliStr_sub(Qureg qureg, vector<int> x, vector<int> y, vector<int> z) {
    auto s = st(qureg);
    qb_cplx out;
    QB_CHECK( qb_densmatr_calcExpecPauliStr_sub(&s, x.data(), (int) x.size(), y.data(), (int) y.size(), z.data(), (int) z.size(), &out) );
    return cq(out);
}

template <bool HasPower, bool UseRealPow>
qcomp gpu_statevec_calcExpecFullStateDiagMatr_sub(Qureg qureg, FullStateDiagMatr matr, qcomp exponent) {
    assert_exponentMatchesTemplateParam(exponent, HasPower, UseRealPow);
    auto s = st(qureg);
    qb_cplx out;
    QB_CHECK( qb_statevec_calcExpecFullStateDiagMatr_sub(&s, qp(matr.gpuElems), HasPower, UseRealPow, qc(exponent), &out) );
    return cq(out);
}

template <bool HasPower, bool UseRealPow>
qcomp gpu_densmatr_calcExpecFullStateDiagMatr_sub(Qureg qureg, FullStateDiagMatr matr, qcomp exponent) {
    assert_exponentMatchesTemplateParam(exponent, HasPower, UseRealPow);
    auto s = st(qureg);
    qb_cplx out;
    QB_CHECK( qb_densmatr_calcExpecFullStateDiagMatr_sub(&s, qp(matr.gpuElems), HasPower, UseRealPow, qc(exponent), &out) );
    return cq(out);
}

template qcomp gpu_statevec_calcExpecFullStateDiagMatr_sub<true, true >(Qureg, FullStateDiagMatr, qcomp);
template qcomp gpu_statevec_calcExpecFullStateDiagMatr_sub<true, false>(Qureg, FullStateDiagMatr, qcomp);
template qcomp gpu_statevec_calcExpecFullStateDiagMatr_sub<false,false>(Qureg, FullStateDiagMatr, qcomp);
template qcomp gpu_statevec_calcExpecFullStateDiagMatr_sub<false,true >(Qureg, FullStateDiagMatr, qcomp);

template qcomp gpu_densmatr_calcExpecFullStateDiagMatr_sub<true, true >(Qureg, FullStateDiagMatr, qcomp);
template qcomp gpu_densmatr_calcExpecFullStateDiagMatr_sub<true, false>(Qureg, FullStateDiagMatr, qcomp);
template qcomp gpu_densmatr_calcExpecFullStateDiagMatr_sub<false,false>(Qureg, FullStateDiagMatr, qcomp);
template qcomp gpu_densmatr_calcExpecFullStateDiagMatr_sub<false,true >(Qureg, FullStateDiagMatr, qcomp);


/*
 * PROJECTORS (gpu_subroutines.cpp:1783-1825)
 */

template <int NumQubits>
void gpu_statevec_multiQubitProjector_sub(Qureg qureg, vector<int> qubits, vector<int> outcomes, qreal prob) {
    assert_numTargsMatchesTemplateParam(qubits.size(), NumQubits);
    auto s = st(qureg);
    QB_CHECK( qb_statevec_multiQubitProjector_sub(&s, qubits.data(), outcomes.data(), (int) qubits.size(), prob) );
}

template <int NumQubits>
void gpu_densmatr_multiQubitProjector_sub(Qureg qureg, vector<int> qubits, vector<int> outcomes, qreal prob) {
    assert_numTargsMatchesTemplateParam(qubits.size(), NumQubits);
    auto s = st(qureg);
    QB_CHECK( qb_densmatr_multiQubitProjector_sub(&s, qubits.data(), outcomes.data(), (int) qubits.size(), prob) );
}

INSTANTIATE_FUNC_OPTIMISED_FOR_NUM_TARGS( void, gpu_statevec_multiQubitProjector_sub, (Qureg, vector<int>, vector<int>, qreal) )
INSTANTIATE_FUNC_OPTIMISED_FOR_NUM_TARGS( void, gpu_densmatr_multiQubitProjector_sub, (Qureg, vector<int>, vector<int>, qreal) )


/*
 * STATE INITIALISATION (gpu_subroutines.cpp:1831-1862)
 */

void gpu_statevec_initUniformState_sub(Qureg qureg, qcomp amp) {
    auto s = st(qureg);
    QB_CHECK( qb_statevec_initUniformState_sub(&s, qc(amp)) );
}

void gpu_statevec_initDebugState_sub(Qureg qureg) {
    auto s = st(qureg);
    QB_CHECK( qb_statevec_initDebugState_sub(&s) );
}

void gpu_statevec_initUnnormalisedUniformlyRandomPureStateAmps_sub(Qureg qureg) {
    // the seed is drawn from QuEST's host generator exactly as the reference does
    // (gpu_subroutines.cpp:1853-1860), so user seeding (setSeeds) still determines the state
    unsigned seed = rand_getThreadSharedRandomSeed(qureg.isDistributed);
    auto s = st(qureg);
    QB_CHECK( qb_statevec_initUnnormalisedUniformlyRandomPureStateAmps_sub(&s, seed) );
}
