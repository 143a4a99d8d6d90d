/* quest_b200.h -- C ABI of the B200-native amplitude-update backend for QuEST v4.1.0
 *
 * This is the drop-in boundary.  Every entry point replaces one
 * function of the reference's GPU backend interface
 *     quest/src/gpu/gpu_subroutines.hpp:24-196   (the 63 gpu_*_sub simulation routines)
 *     quest/src/gpu/gpu_config.hpp:41-120        (device queries, memory, copies, cache)
 *     quest/src/comm/comm_routines.hpp:29-77     (amplitude exchange / reductions)
 *     quest/src/comm/comm_config.hpp:15-27       (bootstrap)
 * but with a plain C signature: POD structs, raw pointers, sizes and ints.  No C++ types,
 * no torch types, no templates.  The reference's template parameters (NumCtrls, NumTargs,
 * ApplyConj, HasPower, ...) become ordinary run-time arguments.  The C++ shim in
 * quest_b200/shim/ defines the reference's own gpu_* / comm_* symbols on top of this ABI
 * so that quest/src/core/accelerator.cpp links against it unchanged.
 *
 * Conventions
 *   - qb_cplx is layout-identical to qcomp = std::complex<qreal> (quest/include/types.h:45; qreal = double for
 *     FLOAT_PRECISION=2, float for 1, quest/include/precision.h:80-96) and to CUDA's double2 / float2.
 *   - qb_index is qindex = long long (quest/include/precision.h:36).
 *   - qb_state carries the fields of Qureg (quest/include/qureg.h:49-80) that device code needs.
 *     amps/buffer are DEVICE pointers, 16-byte aligned; they may be offset views into a larger
 *     allocation (the reference "spoofs" Quregs, core/localiser.cpp:264-440).
 *   - qubit lists are HOST int arrays; matrices passed as `const qb_cplx m[...]` are HOST,
 *     row-major (core/fastmath.hpp:93-97); arguments named dev* are DEVICE pointers
 *     (CompMatr.gpuElemsFlat, DiagMatr.gpuElems, FullStateDiagMatr.gpuElems, SuperOp.gpuElemsFlat).
 *   - every function returns 0 on success or a non-zero cudaError_t / ncclResult_t-derived
 *     code; qb_error_string() describes the most recent failure on the calling thread.
 *     There is no CPU fallback: with no usable device every compute call fails.
 *   - gate kernels are asynchronous on the library stream (qb_get_stream); reductions, copies
 *     and qb_sync() are synchronous w.r.t. that stream, matching gpu_config.cpp:449-469.
 */
#ifndef QUEST_B200_H
#define QUEST_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Precision (quest/include/precision.h:80-96): QB_PRECISION follows QuEST's FLOAT_PRECISION -- 2 (default): qreal = double;
 * 1: qreal = float.  It is a build-time property of the library, exactly as in the reference (one libQuEST per precision).
 * Amplitudes, matrices and complex factors travel as qb_cplx = {qb_real re, im} (layout of qcomp and of CUDA's
 * double2 / float2).  Real scalars (probabilities, angles) and reduction results stay `double` in BOTH builds: reductions
 * accumulate in double whatever the amplitude type, and the caller narrows the result to its qreal. */
#ifndef QB_PRECISION
#  ifdef FLOAT_PRECISION
#    define QB_PRECISION FLOAT_PRECISION
#  else
#    define QB_PRECISION 2
#  endif
#endif
#if QB_PRECISION == 1
typedef float qb_real;
#elif QB_PRECISION == 2
typedef double qb_real;
#else
#  error "quest_b200 supports FLOAT_PRECISION 1 (float) and 2 (double); quad precision has no GPU backend (precision.h:117-119)"
#endif
typedef struct { qb_real re, im; } qb_cplx;
int         qb_precision(void);                          /* QB_PRECISION this library was built with */
typedef long long qb_index;

typedef struct qb_state {
    qb_cplx* amps;              /* Qureg.gpuAmps                                         */
    qb_cplx* buffer;            /* Qureg.gpuCommBuffer, NULL if the Qureg is not distributed */
    qb_index numAmpsPerNode;    /* Qureg.numAmpsPerNode (a power of two)                 */
    int      logNumAmpsPerNode; /* Qureg.logNumAmpsPerNode                               */
    int      rank;              /* Qureg.rank  (0 when not distributed)                  */
    int      numQubits;         /* Qureg.numQubits (ket qubits for a density matrix)     */
    int      logNumColsPerNode; /* Qureg.logNumColsPerNode (density matrices only)       */
    int      isDensityMatrix;   /* Qureg.isDensityMatrix                                 */
} qb_state;

/* ------------------------------------------------------------------------------------------
 * runtime: device, memory, stream            (replaces gpu_config.hpp:41-120)
 * ---------------------------------------------------------------------------------------- */
int         qb_abi_version(void);
const char* qb_error_string(void);
int         qb_num_devices(void);                        /* gpu_getNumberOfLocalGpus   gpu_config.cpp:149 */
int         qb_is_device_available(void);                /* gpu_isGpuAvailable         gpu_config.cpp:165 */
int         qb_bind_device(int deviceIndex);             /* gpu_bindLocalGPUsToNodes   gpu_config.cpp:332 */
int         qb_bound_device(void);                       /* -1 until bound */
int         qb_compute_capability(void);                 /* gpu_getComputeCapability   gpu_config.cpp:124 */
int         qb_mem_info(size_t* freeBytes, size_t* totalBytes); /* gpu_config.cpp:220,239 */
int         qb_supports_mem_pools(void);                 /* gpu_doesGpuSupportMemPools gpu_config.cpp:254 */
qb_index    qb_max_concurrent_threads(void);             /* gpu_getMaxNumConcurrentThreads :269 */
int         qb_device_uuid(char out16[16]);              /* getBoundGpuUuid            gpu_config.cpp:298 */
int         qb_sync(void);                               /* gpu_sync                   gpu_config.cpp:382 */
int         qb_flush(void);                              /* launch every deferred gate now (no host wait); implied by every non-gate entry point */
void*       qb_get_stream(void);                         /* cudaStream_t all compute is issued on */
int         qb_set_stream(void* cudaStream);
qb_cplx*    qb_alloc(qb_index numAmps, int* status);     /* gpu_allocArray :399 (NULL + status 0 on OOM)  */
int         qb_free(qb_cplx* devPtr);                    /* gpu_deallocArray           gpu_config.cpp:424 */
int         qb_copy_h2d(qb_cplx* dev, const qb_cplx* host, qb_index numElems); /* gpu_copyCpuToGpu :535 */
int         qb_copy_d2h(qb_cplx* host, const qb_cplx* dev, qb_index numElems); /* gpu_copyGpuToCpu :538 */
int         qb_copy_d2d(qb_cplx* dst, const qb_cplx* src, qb_index numElems);  /* gpu_copyArray    :521 */
qb_cplx*    qb_get_cache(qb_index numElems, int* status);/* gpu_getCacheOfSize         gpu_config.cpp:639 */
int         qb_clear_cache(void);                        /* gpu_clearCache             gpu_config.cpp:665 */
size_t      qb_cache_bytes(void);                        /* gpu_getCacheMemoryInBytes  gpu_config.cpp:683 */
unsigned long long qb_launch_count(void);                /* kernels launched by this library so far */
int         qb_set_tile_engine(int enabled);             /* 1 (default): fused TMA tile passes (gate absorption + commuting re-order); 2: fused, program order; 0: direct kernels only */
/* cumulative statistics of the deferred-gate engine since load: out[0] fused tile passes launched, [1] register rounds
 * in them, [2] gates executed inside tile passes, [3] gates run as direct kernels, [4] gates received by the queues,
 * [5] FP64 fused multiply-adds executed for them (bench.py derives the FP64-pipe fraction of the roofline from it),
 * [6] bytes those passes streamed through HBM (read + write of the amplitudes they touch), [7] reserved */
int         qb_tile_stats(double out[8]);

/* ------------------------------------------------------------------------------------------
 * getters / setters                          (gpu_subroutines.hpp:24-36)
 * ---------------------------------------------------------------------------------------- */
int qb_statevec_getAmp_sub(const qb_state* q, qb_index localInd, qb_cplx* out);
/* strings: numTerms x {lowPaulis, highPaulis} (quest/include/paulis.h:53-61); coeffs: HOST */
int qb_densmatr_setAmpsToPauliStrSum_sub(const qb_state* q, const qb_cplx* coeffs,
        const unsigned long long* strings, qb_index numTerms);
int qb_fullstatediagmatr_setElemsToPauliStrSum(qb_cplx* devElems, qb_index numElemsPerNode, int rank,
        const qb_cplx* coeffs, const unsigned long long* strings, qb_index numTerms);

/* ------------------------------------------------------------------------------------------
 * communication-buffer packing               (gpu_subroutines.hpp:39-45)
 * ---------------------------------------------------------------------------------------- */
int qb_statevec_packAmpsIntoBuffer(const qb_state* q, const int* qubits, const int* qubitStates,
        int numQubits, qb_index* numPacked);
int qb_statevec_packPairSummedAmpsIntoBuffer(const qb_state* q, int qubit1, int qubit2, int qubit3,
        int bit2, qb_index* numPacked);

/* ------------------------------------------------------------------------------------------
 * swaps                                      (gpu_subroutines.hpp:48-54)
 * ---------------------------------------------------------------------------------------- */
int qb_statevec_anyCtrlSwap_subA(const qb_state* q, const int* ctrls, const int* ctrlStates, int numCtrls,
        int targ1, int targ2);
int qb_statevec_anyCtrlSwap_subB(const qb_state* q, const int* ctrls, const int* ctrlStates, int numCtrls);
int qb_statevec_anyCtrlSwap_subC(const qb_state* q, const int* ctrls, const int* ctrlStates, int numCtrls,
        int targ, int targState);

/* ------------------------------------------------------------------------------------------
 * dense matrices                             (gpu_subroutines.hpp:57-67)
 * ---------------------------------------------------------------------------------------- */
int qb_statevec_anyCtrlOneTargDenseMatr_subA(const qb_state* q, const int* ctrls, const int* ctrlStates,
        int numCtrls, int targ, const qb_cplx matr[4]);
int qb_statevec_anyCtrlOneTargDenseMatr_subB(const qb_state* q, const int* ctrls, const int* ctrlStates,
        int numCtrls, qb_cplx fac0, qb_cplx fac1);
int qb_statevec_anyCtrlTwoTargDenseMatr_sub(const qb_state* q, const int* ctrls, const int* ctrlStates,
        int numCtrls, int targ1, int targ2, const qb_cplx matr[16]);
int qb_statevec_anyCtrlAnyTargDenseMatr_sub(const qb_state* q, const int* ctrls, const int* ctrlStates,
        int numCtrls, const int* targs, int numTargs, const qb_cplx* devMatrFlat, int applyConj);

/* ------------------------------------------------------------------------------------------
 * diagonal matrices                          (gpu_subroutines.hpp:70-82)
 * ---------------------------------------------------------------------------------------- */
int qb_statevec_anyCtrlOneTargDiagMatr_sub(const qb_state* q, const int* ctrls, const int* ctrlStates,
        int numCtrls, int targ, const qb_cplx elems[2]);
int qb_statevec_anyCtrlTwoTargDiagMatr_sub(const qb_state* q, const int* ctrls, const int* ctrlStates,
        int numCtrls, int targ1, int targ2, const qb_cplx elems[4]);
int qb_statevec_anyCtrlAnyTargDiagMatr_sub(const qb_state* q, const int* ctrls, const int* ctrlStates,
        int numCtrls, const int* targs, int numTargs, const qb_cplx* devElems,
        int applyConj, int hasPower, qb_cplx exponent);
int qb_statevec_allTargDiagMatr_sub(const qb_state* q, const qb_cplx* devElems, int hasPower, qb_cplx exponent);
int qb_densmatr_allTargDiagMatr_sub(const qb_state* q, const qb_cplx* devElems, qb_index matrNumElems,
        int hasPower, int multiplyOnly, qb_cplx exponent);

/* ------------------------------------------------------------------------------------------
 * Pauli tensors and gadgets                  (gpu_subroutines.hpp:85-93)
 * ---------------------------------------------------------------------------------------- */
int qb_statevector_anyCtrlPauliTensorOrGadget_subA(const qb_state* q, const int* ctrls, const int* ctrlStates,
        int numCtrls, const int* x, int numX, const int* y, int numY, const int* z, int numZ,
        qb_cplx ampFac, qb_cplx pairAmpFac);
int qb_statevector_anyCtrlPauliTensorOrGadget_subB(const qb_state* q, const int* ctrls, const int* ctrlStates,
        int numCtrls, const int* x, int numX, const int* y, int numY, const int* z, int numZ,
        qb_cplx ampFac, qb_cplx pairAmpFac, qb_index bufferMaskXY);
int qb_statevector_anyCtrlAnyTargZOrPhaseGadget_sub(const qb_state* q, const int* ctrls, const int* ctrlStates,
        int numCtrls, const int* targs, int numTargs, qb_cplx fac0, qb_cplx fac1);

/* ------------------------------------------------------------------------------------------
 * Qureg combination                          (gpu_subroutines.hpp:96-104)
 * ---------------------------------------------------------------------------------------- */
int qb_statevec_setQuregToSuperposition_sub(qb_cplx facOut, const qb_state* out, qb_cplx fac1,
        const qb_state* in1, qb_cplx fac2, const qb_state* in2);
int qb_densmatr_mixQureg_subA(double outProb, const qb_state* out, double inProb, const qb_state* inDensMatr);
int qb_densmatr_mixQureg_subB(double outProb, const qb_state* out, double inProb, const qb_state* inStateVec);
int qb_densmatr_mixQureg_subC(double outProb, const qb_state* out, double inProb);

/* ------------------------------------------------------------------------------------------
 * decoherence                                (gpu_subroutines.hpp:107-132)
 * ---------------------------------------------------------------------------------------- */
int qb_densmatr_oneQubitDephasing_subA(const qb_state* q, int qubit, double prob);
int qb_densmatr_oneQubitDephasing_subB(const qb_state* q, int qubit, double prob);
int qb_densmatr_twoQubitDephasing_subA(const qb_state* q, int qubitA, int qubitB, double prob);
int qb_densmatr_twoQubitDephasing_subB(const qb_state* q, int qubitA, int qubitB, double prob);
int qb_densmatr_oneQubitDepolarising_subA(const qb_state* q, int qubit, double prob);
int qb_densmatr_oneQubitDepolarising_subB(const qb_state* q, int qubit, double prob);
int qb_densmatr_twoQubitDepolarising_subA(const qb_state* q, int qubit1, int qubit2, double prob);
int qb_densmatr_twoQubitDepolarising_subB(const qb_state* q, int qubit1, int qubit2, double prob);
int qb_densmatr_twoQubitDepolarising_subC(const qb_state* q, int qubit1, int qubit2, double prob);
int qb_densmatr_twoQubitDepolarising_subD(const qb_state* q, int qubit1, int qubit2, double prob);
int qb_densmatr_twoQubitDepolarising_subE(const qb_state* q, int qubit1, int qubit2, double prob);
int qb_densmatr_twoQubitDepolarising_subF(const qb_state* q, int qubit1, int qubit2, double prob);
int qb_densmatr_oneQubitPauliChannel_subA(const qb_state* q, int qubit, double pI, double pX, double pY, double pZ);
int qb_densmatr_oneQubitPauliChannel_subB(const qb_state* q, int qubit, double pI, double pX, double pY, double pZ);
int qb_densmatr_oneQubitDamping_subA(const qb_state* q, int qubit, double prob);
int qb_densmatr_oneQubitDamping_subB(const qb_state* q, int qubit, double prob);
int qb_densmatr_oneQubitDamping_subC(const qb_state* q, int qubit, double prob);
int qb_densmatr_oneQubitDamping_subD(const qb_state* q, int qubit, double prob);

/* ------------------------------------------------------------------------------------------
 * partial trace                              (gpu_subroutines.hpp:135-139)
 * ---------------------------------------------------------------------------------------- */
int qb_densmatr_partialTrace_sub(const qb_state* in, const qb_state* out, const int* targs,
        const int* pairTargs, int numTargs);

/* ------------------------------------------------------------------------------------------
 * probabilities                              (gpu_subroutines.hpp:142-153)
 * ---------------------------------------------------------------------------------------- */
int qb_statevec_calcTotalProb_sub(const qb_state* q, double* out);
int qb_densmatr_calcTotalProb_sub(const qb_state* q, double* out);
int qb_statevec_calcProbOfMultiQubitOutcome_sub(const qb_state* q, const int* qubits, const int* outcomes,
        int numQubits, double* out);
int qb_densmatr_calcProbOfMultiQubitOutcome_sub(const qb_state* q, const int* qubits, const int* outcomes,
        int numQubits, double* out);
int qb_statevec_calcProbsOfAllMultiQubitOutcomes_sub(double* outProbs, const qb_state* q, const int* qubits,
        int numQubits);
int qb_densmatr_calcProbsOfAllMultiQubitOutcomes_sub(double* outProbs, const qb_state* q, const int* qubits,
        int numQubits);

/* ------------------------------------------------------------------------------------------
 * inner products                             (gpu_subroutines.hpp:156-164)
 * ---------------------------------------------------------------------------------------- */
int qb_statevec_calcInnerProduct_sub(const qb_state* a, const qb_state* b, qb_cplx* out);
int qb_densmatr_calcHilbertSchmidtDistance_sub(const qb_state* a, const qb_state* b, double* out);
int qb_densmatr_calcFidelityWithPureState_sub(const qb_state* rho, const qb_state* psi, int conj, qb_cplx* out);

/* ------------------------------------------------------------------------------------------
 * expectation values                         (gpu_subroutines.hpp:167-180)
 * ---------------------------------------------------------------------------------------- */
int qb_statevec_calcExpecAnyTargZ_sub(const qb_state* q, const int* targs, int numTargs, double* out);
int qb_densmatr_calcExpecAnyTargZ_sub(const qb_state* q, const int* targs, int numTargs, qb_cplx* out);
int qb_statevec_calcExpecPauliStr_subA(const qb_state* q, const int* x, int numX, const int* y, int numY,
        const int* z, int numZ, qb_cplx* out);
int qb_statevec_calcExpecPauliStr_subB(const qb_state* q, const int* x, int numX, const int* y, int numY,
        const int* z, int numZ, qb_cplx* out);
int qb_densmatr_calcExpecPauliStr_sub(const qb_state* q, const int* x, int numX, const int* y, int numY,
        const int* z, int numZ, qb_cplx* out);
int qb_statevec_calcExpecFullStateDiagMatr_sub(const qb_state* q, const qb_cplx* devElems, int hasPower,
        int useRealPow, qb_cplx exponent, qb_cplx* out);
int qb_densmatr_calcExpecFullStateDiagMatr_sub(const qb_state* q, const qb_cplx* devElems, int hasPower,
        int useRealPow, qb_cplx exponent, qb_cplx* out);
/* fused extension (SURVEY.md 8f / BASELINE.md cfg 5): all terms of a Pauli-string sum that have
 * only suffix X/Y in ONE pass over the state. masks: numTerms x {maskXY, maskYZ}; outTerms[t] =
 * i^{numY_t}-free raw sum  sum_n (-1)^{popc(j&maskYZ)} conj(a_n) a_j , j = n ^ maskXY  (HOST out). */
int qb_statevec_calcExpecPauliStrBatch_subA(const qb_state* q, const unsigned long long* masks,
        int numTerms, qb_cplx* outTerms);
/* same with the partner amplitudes a_j read from the communication buffer (after a full exchange) */
int qb_statevec_calcExpecPauliStrBatch_subB(const qb_state* q, const unsigned long long* masks,
        int numTerms, qb_cplx* outTerms);

/* ------------------------------------------------------------------------------------------
 * projectors                                 (gpu_subroutines.hpp:183-188)
 * ---------------------------------------------------------------------------------------- */
int qb_statevec_multiQubitProjector_sub(const qb_state* q, const int* qubits, const int* outcomes,
        int numQubits, double prob);
int qb_densmatr_multiQubitProjector_sub(const qb_state* q, const int* qubits, const int* outcomes,
        int numQubits, double prob);

/* ------------------------------------------------------------------------------------------
 * state initialisation                       (gpu_subroutines.hpp:191-196)
 * ---------------------------------------------------------------------------------------- */
int qb_statevec_initUniformState_sub(const qb_state* q, qb_cplx amp);
int qb_statevec_initDebugState_sub(const qb_state* q);
int qb_statevec_initUnnormalisedUniformlyRandomPureStateAmps_sub(const qb_state* q, unsigned seed);

/* ------------------------------------------------------------------------------------------
 * communication over NCCL / NVLink            (replaces quest/src/comm, MPI -> NCCL)
 * one process per GPU; ranks exchange amplitudes pairwise with rank ^ mask.
 * ---------------------------------------------------------------------------------------- */
#define QB_COMM_ID_BYTES 128
int qb_comm_get_unique_id(char id[QB_COMM_ID_BYTES]);              /* rank 0, then share out-of-band */
int qb_comm_init(int rank, int numRanks, const char id[QB_COMM_ID_BYTES]); /* comm_init  comm_config.cpp:101 */
int qb_comm_end(void);                                              /* comm_end   comm_config.cpp:114 */
int qb_comm_is_init(void);
int qb_comm_rank(void);
int qb_comm_num_ranks(void);
/* transport of the NEXT communicator id made by qb_comm_get_unique_id: 0 = NCCL over NVLink (default; also taken from
 * QUEST_B200_TRANSPORT=nccl|shm), 1 = shared host memory + CUDA IPC, for boxes with fewer GPUs than ranks (ranks share
 * a device, as the reference allows with PERMIT_NODES_TO_SHARE_GPU, CMakeLists.txt:235-240, api/environment.cpp:110).
 * The id carries the choice, so qb_comm_init on every rank follows rank 0. */
int qb_comm_set_transport(int transport);
int qb_comm_transport(void);                                        /* of the live communicator: 0 NCCL, 1 shared memory */
int qb_comm_barrier(void);                                          /* comm_sync  comm_config.cpp:130 */
/* send numAmps from dev `send` to pairRank while receiving numAmps into dev `recv` (comm_routines.cpp:209-232) */
int qb_comm_exchange(const qb_cplx* devSend, qb_cplx* devRecv, qb_index numAmps, int pairRank);
int qb_comm_send(const qb_cplx* devSend, qb_index numAmps, int pairRank);   /* comm_routines.cpp:521 */
int qb_comm_recv(qb_cplx* devRecv, qb_index numAmps, int pairRank);         /* comm_routines.cpp:537 */
/* every rank contributes numAmpsPerRank amps; result (numRanks*numAmpsPerRank) lands in devRecv on all
 * ranks (comm_routines.cpp:553-607, Ibcast-from-every-rank == all-gather) */
int qb_comm_allgather(const qb_cplx* devSend, qb_cplx* devRecv, qb_index numAmpsPerRank);
int qb_comm_allreduce_sum(double* hostValues, qb_index numValues);   /* comm_reduceAmp/Real/Reals :714-744 */
int qb_comm_allreduce_and(int* hostFlag);                            /* comm_isTrueOnAllNodes     :747 */
int qb_comm_broadcast_bytes(void* hostBuf, size_t numBytes, int rootRank); /* comm_broadcast*     :632-695 */
int qb_comm_gather_bytes(const void* hostSend, void* hostRecvOnRoot, size_t numBytesPerRank, int rootRank);
int qb_comm_sendrecv_host(const qb_cplx* hostSend, qb_cplx* hostRecv, qb_index numAmps, int sendRank, int recvRank);

/* ------------------------------------------------------------------------------------------
 * fused compute + exchange over NVLink peer memory (no reference equivalent: replaces the
 * "exchange into buffer, then combine" pairs of core/localiser.cpp:854-869 and :941-953).
 * The partner rank's amplitudes are mapped into this process with CUDA IPC; ONE kernel per GPU reads and
 * writes both ranks' amplitudes, so no communication buffer, no pack/unpack pass and no combine pass exist,
 * and the NVLink transfer overlaps the arithmetic element by element.  Both ranks of a pair must call the
 * same function (SPMD); ordering between the two GPUs uses system-scope flags in IPC-shared memory.
 * ---------------------------------------------------------------------------------------- */
int qb_p2p_is_available(void);          /* 1 when comm is initialised and every peer's memory is mappable */
int qb_p2p_set_enabled(int enabled);    /* 0 forces the NCCL exchange path (used by parity tests)          */
/* gate on a PREFIX target: this rank holds target bit `rankBit`; pairRank holds the other half of every pair.
 * matr = row-major 2x2.  ctrls are suffix controls. (localiser.cpp:941-953 + gpu_subroutines.cpp:319-345) */
int qb_p2p_anyCtrlOneTargDenseMatr(const qb_state* q, const int* ctrls, const int* ctrlStates, int numCtrls,
        int pairRank, int rankBit, const qb_cplx matr[4]);
/* SWAP of a prefix qubit (partner = pairRank) with suffix qubit suffixTarg: the half-shards whose suffix
 * bit differs from the rank bit trade places. (localiser.cpp:854-869 + gpu_subroutines.cpp:251-280) */
int qb_p2p_swapHalves(const qb_state* q, int suffixTarg, int pairRank);
int         qb_p2p_swapHalvesDeferred(const qb_state* q, int suffixTarg, int pairRank); /* same, but may overtake queued gates that do not touch suffixTarg (caller vouches none depends on the rank bit) */
/* same exchange through the partner's communication buffer on a second stream (copy engines), OVERLAPPED with the deferred
 * gates of q: they run on the staying half while the leaving half is in flight, then on the arrived half.  Drains q's queue.
 * Collective over the pair: both ranks must call it (decide from rank-independent state). */
int         qb_p2p_swapHalvesOverlapped(const qb_state* q, int suffixTarg, int pairRank);
unsigned long long qb_p2p_overlapped_count(int withQueuedGatesOnly); /* overlapped swaps issued so far (all / those that had gates to overlap) */
int         qb_queue_info(const qb_state* q, unsigned long long* touchedSuffixMask, unsigned long long* flushEpoch); /* deferred gates of q: count, suffix qubits they involve, flush counter */
int         qb_p2p_set_swap_mode(int mode);              /* half-shard swap: 0 (default) in-place exchange kernel; 1: kernel push into the partner's buffer + local unpack; 2: copy engines (all within 10%: profiles/) */
int         qb_p2p_stats(unsigned long long* numExchanges, unsigned long long* linkBytesPerDir); /* peer-memory exchanges issued by this rank so far, and the bytes each sent one way */

#ifdef __cplusplus
}
#endif
#endif /* QUEST_B200_H */
