"""ctypes binding of QuEST v4.1.0's public C API (quest/include/*.h).

The same binding drives BOTH libraries that export this API:
  * quest_b200/lib/libQuEST.so -- the drop-in: QuEST's unmodified host layers over the B200 backend
  * oracle/_ref/libQuEST.so    -- the unmodified reference CPU/OpenMP build (the parity oracle)
so the parity tests read like the reference's own tests: `q.applyCompMatr1(qureg, t, m)` on each
library, then compare amplitudes.  One process can hold only ONE of them (both export the same symbols
and keep process-global singletons: api/environment.cpp:49, core/randomiser.cpp:48-50); the tests run
the oracle in a worker subprocess (tests/_worker.py).

Struct layouts mirror quest/include/{qureg.h:49-80, environment.h:33-44, matrices.h:68-230,
channels.h:74-118, paulis.h:53-80}.  qcomp (std::complex<qreal> / qreal _Complex) is passed and returned BY VALUE as a
{qreal, qreal} struct, which the SysV x86-64 ABI classifies identically.

Precision is a build-time property of a QuEST library (FLOAT_PRECISION, quest/include/precision.h:80-96), so it is an
import-time property of this binding: QUEST_PRECISION=1 in the environment selects qreal = float (the *_f32 libraries);
the default is 2 (double).  The test workers set it before importing (tests/_worker.py).
"""
import ctypes as C
import os
import numpy as np

PRECISION = int(os.environ.get("QUEST_PRECISION", "2"))
assert PRECISION in (1, 2), "QUEST_PRECISION must be 1 (float) or 2 (double)"
c_qreal = C.c_float if PRECISION == 1 else C.c_double
np_qcomp = np.complex64 if PRECISION == 1 else np.complex128
c_qindex = C.c_longlong


class qcomp(C.Structure):
    _fields_ = [("re", c_qreal), ("im", c_qreal)]

    def __complex__(self):
        return complex(self.re, self.im)


def _qc(z):
    z = complex(z)
    return qcomp(z.real, z.imag)


class Qureg(C.Structure):
    _fields_ = [
        ("isMultithreaded", C.c_int), ("isGpuAccelerated", C.c_int), ("isDistributed", C.c_int),
        ("rank", C.c_int), ("numNodes", C.c_int), ("logNumNodes", C.c_int),
        ("isDensityMatrix", C.c_int), ("numQubits", C.c_int),
        ("numAmps", c_qindex), ("logNumAmps", c_qindex),
        ("numAmpsPerNode", c_qindex), ("logNumAmpsPerNode", c_qindex), ("logNumColsPerNode", c_qindex),
        ("cpuAmps", C.c_void_p), ("gpuAmps", C.c_void_p),
        ("cpuCommBuffer", C.c_void_p), ("gpuCommBuffer", C.c_void_p),
    ]


class QuESTEnv(C.Structure):
    _fields_ = [("isMultithreaded", C.c_int), ("isGpuAccelerated", C.c_int), ("isDistributed", C.c_int),
                ("rank", C.c_int), ("numNodes", C.c_int)]


class CompMatr1(C.Structure):
    _fields_ = [("numQubits", C.c_int), ("numRows", c_qindex), ("elems", qcomp * 2 * 2)]


class CompMatr2(C.Structure):
    _fields_ = [("numQubits", C.c_int), ("numRows", c_qindex), ("elems", qcomp * 4 * 4)]


class CompMatr(C.Structure):
    _fields_ = [("numQubits", C.c_int), ("numRows", c_qindex),
                ("isApproxUnitary", C.POINTER(C.c_int)), ("isApproxHermitian", C.POINTER(C.c_int)),
                ("wasGpuSynced", C.POINTER(C.c_int)),
                ("cpuElems", C.c_void_p), ("cpuElemsFlat", C.c_void_p), ("gpuElemsFlat", C.c_void_p)]


class DiagMatr1(C.Structure):
    _fields_ = [("numQubits", C.c_int), ("numElems", c_qindex), ("elems", qcomp * 2)]


class DiagMatr2(C.Structure):
    _fields_ = [("numQubits", C.c_int), ("numElems", c_qindex), ("elems", qcomp * 4)]


class DiagMatr(C.Structure):
    _fields_ = [("numQubits", C.c_int), ("numElems", c_qindex),
                ("isApproxUnitary", C.POINTER(C.c_int)), ("isApproxHermitian", C.POINTER(C.c_int)),
                ("isApproxNonZero", C.POINTER(C.c_int)), ("isStrictlyNonNegative", C.POINTER(C.c_int)),
                ("wasGpuSynced", C.POINTER(C.c_int)),
                ("cpuElems", C.c_void_p), ("gpuElems", C.c_void_p)]


class FullStateDiagMatr(C.Structure):
    _fields_ = [("numQubits", C.c_int), ("numElems", c_qindex),
                ("isGpuAccelerated", C.c_int), ("isMultithreaded", C.c_int), ("isDistributed", C.c_int),
                ("numElemsPerNode", c_qindex),
                ("isApproxUnitary", C.POINTER(C.c_int)), ("isApproxHermitian", C.POINTER(C.c_int)),
                ("isApproxNonZero", C.POINTER(C.c_int)), ("isStrictlyNonNegative", C.POINTER(C.c_int)),
                ("wasGpuSynced", C.POINTER(C.c_int)),
                ("cpuElems", C.c_void_p), ("gpuElems", C.c_void_p)]


class SuperOp(C.Structure):
    _fields_ = [("numQubits", C.c_int), ("numRows", c_qindex),
                ("cpuElems", C.c_void_p), ("cpuElemsFlat", C.c_void_p), ("gpuElemsFlat", C.c_void_p),
                ("wasGpuSynced", C.POINTER(C.c_int))]


class KrausMap(C.Structure):
    _fields_ = [("numQubits", C.c_int), ("numMatrices", C.c_int), ("numRows", c_qindex),
                ("matrices", C.c_void_p), ("superop", SuperOp), ("isApproxCPTP", C.POINTER(C.c_int))]


class PauliStr(C.Structure):
    _fields_ = [("lowPaulis", C.c_ulonglong), ("highPaulis", C.c_ulonglong)]


class PauliStrSum(C.Structure):
    _fields_ = [("numTerms", c_qindex), ("strings", C.POINTER(PauliStr)), ("coeffs", C.POINTER(qcomp)),
                ("isApproxHermitian", C.POINTER(C.c_int))]


_I, _D, _Q, _P, _IP, _U = C.c_int, c_qreal, c_qindex, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_uint)
_CP = C.POINTER(qcomp)

# name -> (restype, [argtypes])
_SIGS = {
    # environment.h / debug.h
    "initQuESTEnv": (None, []), "initCustomQuESTEnv": (None, [_I, _I, _I]), "finalizeQuESTEnv": (None, []),
    "syncQuESTEnv": (None, []), "reportQuESTEnv": (None, []), "isQuESTEnvInit": (_I, []),
    "getQuESTEnv": (QuESTEnv, []),
    "setSeeds": (None, [_U, _I]), "setSeedsToDefault": (None, []), "getNumSeeds": (_I, []),
    "setValidationOn": (None, []), "setValidationOff": (None, []), "setValidationEpsilon": (None, [_D]),
    "setValidationEpsilonToDefault": (None, []), "getValidationEpsilon": (_D, []),
    "getGpuCacheSize": (_Q, []), "clearGpuCache": (None, []),
    # qureg.h
    "createQureg": (Qureg, [_I]), "createDensityQureg": (Qureg, [_I]), "createForcedQureg": (Qureg, [_I]),
    "createForcedDensityQureg": (Qureg, [_I]), "createCustomQureg": (Qureg, [_I, _I, _I, _I, _I]),
    "createCloneQureg": (Qureg, [Qureg]), "destroyQureg": (None, [Qureg]),
    "reportQuregParams": (None, [Qureg]), "reportQureg": (None, [Qureg]),
    "syncQuregToGpu": (None, [Qureg]), "syncQuregFromGpu": (None, [Qureg]),
    "getQuregAmps": (None, [_P, Qureg, _Q, _Q]),
    # qcomp-returning API functions are exposed to C through out-parameter wrappers (include/wrappers.h:46-101)
    "_wrap_getQuregAmp": (None, [_CP, Qureg, _Q]), "_wrap_getDensityQuregAmp": (None, [_CP, Qureg, _Q, _Q]),
    "_wrap_calcInnerProduct": (None, [Qureg, Qureg, _CP]),
    "_wrap_calcExpecNonHermitianPauliStrSum": (None, [_CP, Qureg, PauliStrSum]),
    "_wrap_calcExpecNonHermitianFullStateDiagMatr": (None, [_CP, Qureg, FullStateDiagMatr]),
    # initialisations.h
    "initBlankState": (None, [Qureg]), "initZeroState": (None, [Qureg]), "initPlusState": (None, [Qureg]),
    "initPureState": (None, [Qureg, Qureg]), "initClassicalState": (None, [Qureg, _Q]),
    "initDebugState": (None, [Qureg]), "initArbitraryPureState": (None, [Qureg, _P]),
    "initRandomPureState": (None, [Qureg]), "initRandomMixedState": (None, [Qureg, _Q]),
    "setQuregAmps": (None, [Qureg, _Q, _P, _Q]), "setDensityQuregFlatAmps": (None, [Qureg, _Q, _P, _Q]),
    "setQuregToClone": (None, [Qureg, Qureg]),
    "setQuregToSuperposition": (None, [qcomp, Qureg, qcomp, Qureg, qcomp, Qureg]),
    "setQuregToRenormalized": (_D, [Qureg]), "setQuregToPauliStrSum": (None, [Qureg, PauliStrSum]),
    "setQuregToPartialTrace": (None, [Qureg, Qureg, _IP, _I]),
    "setQuregToReducedDensityMatrix": (None, [Qureg, Qureg, _IP, _I]),
    # matrices.h / channels.h / paulis.h
    "createCompMatr": (CompMatr, [_I]), "destroyCompMatr": (None, [CompMatr]), "syncCompMatr": (None, [CompMatr]),
    "createDiagMatr": (DiagMatr, [_I]), "destroyDiagMatr": (None, [DiagMatr]), "syncDiagMatr": (None, [DiagMatr]),
    "createFullStateDiagMatr": (FullStateDiagMatr, [_I]),
    "createCustomFullStateDiagMatr": (FullStateDiagMatr, [_I, _I, _I, _I]),
    "destroyFullStateDiagMatr": (None, [FullStateDiagMatr]), "syncFullStateDiagMatr": (None, [FullStateDiagMatr]),
    "setFullStateDiagMatr": (None, [FullStateDiagMatr, _Q, _P, _Q]),
    "setFullStateDiagMatrFromPauliStrSum": (None, [FullStateDiagMatr, PauliStrSum]),
    "createFullStateDiagMatrFromPauliStrSum": (FullStateDiagMatr, [PauliStrSum]),
    "createKrausMap": (KrausMap, [_I, _I]), "destroyKrausMap": (None, [KrausMap]), "syncKrausMap": (None, [KrausMap]),
    "createSuperOp": (SuperOp, [_I]), "destroySuperOp": (None, [SuperOp]), "syncSuperOp": (None, [SuperOp]),
    "_getPauliStrFromInts": (PauliStr, [_IP, _IP, _I]),
    "createPauliStrSum": (PauliStrSum, [C.POINTER(PauliStr), _CP, _Q]), "destroyPauliStrSum": (None, [PauliStrSum]),
    # operations.h
    "applyCompMatr1": (None, [Qureg, _I, CompMatr1]), "applyControlledCompMatr1": (None, [Qureg, _I, _I, CompMatr1]),
    "applyMultiControlledCompMatr1": (None, [Qureg, _IP, _I, _I, CompMatr1]),
    "applyMultiStateControlledCompMatr1": (None, [Qureg, _IP, _IP, _I, _I, CompMatr1]),
    "applyCompMatr2": (None, [Qureg, _I, _I, CompMatr2]), "applyControlledCompMatr2": (None, [Qureg, _I, _I, _I, CompMatr2]),
    "applyMultiControlledCompMatr2": (None, [Qureg, _IP, _I, _I, _I, CompMatr2]),
    "applyMultiStateControlledCompMatr2": (None, [Qureg, _IP, _IP, _I, _I, _I, CompMatr2]),
    "applyCompMatr": (None, [Qureg, _IP, _I, CompMatr]), "applyControlledCompMatr": (None, [Qureg, _I, _IP, _I, CompMatr]),
    "applyMultiControlledCompMatr": (None, [Qureg, _IP, _I, _IP, _I, CompMatr]),
    "applyMultiStateControlledCompMatr": (None, [Qureg, _IP, _IP, _I, _IP, _I, CompMatr]),
    "applyDiagMatr1": (None, [Qureg, _I, DiagMatr1]), "applyControlledDiagMatr1": (None, [Qureg, _I, _I, DiagMatr1]),
    "applyMultiStateControlledDiagMatr1": (None, [Qureg, _IP, _IP, _I, _I, DiagMatr1]),
    "applyDiagMatr2": (None, [Qureg, _I, _I, DiagMatr2]), "applyControlledDiagMatr2": (None, [Qureg, _I, _I, _I, DiagMatr2]),
    "applyMultiStateControlledDiagMatr2": (None, [Qureg, _IP, _IP, _I, _I, _I, DiagMatr2]),
    "applyDiagMatr": (None, [Qureg, _IP, _I, DiagMatr]), "applyControlledDiagMatr": (None, [Qureg, _I, _IP, _I, DiagMatr]),
    "applyMultiStateControlledDiagMatr": (None, [Qureg, _IP, _IP, _I, _IP, _I, DiagMatr]),
    "applyDiagMatrPower": (None, [Qureg, _IP, _I, DiagMatr, qcomp]),
    "applyMultiStateControlledDiagMatrPower": (None, [Qureg, _IP, _IP, _I, _IP, _I, DiagMatr, qcomp]),
    "applyFullStateDiagMatr": (None, [Qureg, FullStateDiagMatr]),
    "applyFullStateDiagMatrPower": (None, [Qureg, FullStateDiagMatr, qcomp]),
    "applyHadamard": (None, [Qureg, _I]), "applyControlledHadamard": (None, [Qureg, _I, _I]),
    "applyMultiStateControlledHadamard": (None, [Qureg, _IP, _IP, _I, _I]),
    "applyS": (None, [Qureg, _I]), "applyT": (None, [Qureg, _I]), "applyControlledS": (None, [Qureg, _I, _I]),
    "applyControlledT": (None, [Qureg, _I, _I]),
    "applyPauliX": (None, [Qureg, _I]), "applyPauliY": (None, [Qureg, _I]), "applyPauliZ": (None, [Qureg, _I]),
    "applyControlledPauliX": (None, [Qureg, _I, _I]), "applyControlledPauliY": (None, [Qureg, _I, _I]),
    "applyControlledPauliZ": (None, [Qureg, _I, _I]),
    "applyMultiControlledPauliX": (None, [Qureg, _IP, _I, _I]),
    "applyMultiStateControlledPauliX": (None, [Qureg, _IP, _IP, _I, _I]),
    "applyMultiStateControlledPauliY": (None, [Qureg, _IP, _IP, _I, _I]),
    "applyMultiStateControlledPauliZ": (None, [Qureg, _IP, _IP, _I, _I]),
    "applyPauliStr": (None, [Qureg, PauliStr]), "applyControlledPauliStr": (None, [Qureg, _I, PauliStr]),
    "applyMultiStateControlledPauliStr": (None, [Qureg, _IP, _IP, _I, PauliStr]),
    "applyPauliGadget": (None, [Qureg, PauliStr, _D]), "applyControlledPauliGadget": (None, [Qureg, _I, PauliStr, _D]),
    "applyMultiStateControlledPauliGadget": (None, [Qureg, _IP, _IP, _I, PauliStr, _D]),
    "applyPhaseGadget": (None, [Qureg, _IP, _I, _D]), "applyControlledPhaseGadget": (None, [Qureg, _I, _IP, _I, _D]),
    "applyMultiStateControlledPhaseGadget": (None, [Qureg, _IP, _IP, _I, _IP, _I, _D]),
    "applyRotateX": (None, [Qureg, _I, _D]), "applyRotateY": (None, [Qureg, _I, _D]), "applyRotateZ": (None, [Qureg, _I, _D]),
    "applyControlledRotateX": (None, [Qureg, _I, _I, _D]), "applyControlledRotateY": (None, [Qureg, _I, _I, _D]),
    "applyControlledRotateZ": (None, [Qureg, _I, _I, _D]),
    "applyMultiStateControlledRotateX": (None, [Qureg, _IP, _IP, _I, _I, _D]),
    "applyMultiStateControlledRotateZ": (None, [Qureg, _IP, _IP, _I, _I, _D]),
    "applyRotateAroundAxis": (None, [Qureg, _I, _D, _D, _D, _D]),
    "applySwap": (None, [Qureg, _I, _I]), "applyControlledSwap": (None, [Qureg, _I, _I, _I]),
    "applyMultiStateControlledSwap": (None, [Qureg, _IP, _IP, _I, _I, _I]),
    "applySqrtSwap": (None, [Qureg, _I, _I]), "applyMultiStateControlledSqrtSwap": (None, [Qureg, _IP, _IP, _I, _I, _I]),
    "applyPhaseFlip": (None, [Qureg, _I]), "applyPhaseShift": (None, [Qureg, _I, _D]),
    "applyTwoQubitPhaseFlip": (None, [Qureg, _I, _I]), "applyTwoQubitPhaseShift": (None, [Qureg, _I, _I, _D]),
    "applyMultiQubitPhaseFlip": (None, [Qureg, _IP, _I]), "applyMultiQubitPhaseShift": (None, [Qureg, _IP, _I, _D]),
    "applyMultiQubitNot": (None, [Qureg, _IP, _I]), "applyControlledMultiQubitNot": (None, [Qureg, _I, _IP, _I]),
    "applyMultiStateControlledMultiQubitNot": (None, [Qureg, _IP, _IP, _I, _IP, _I]),
    "applyQuantumFourierTransform": (None, [Qureg, _IP, _I]), "applyFullQuantumFourierTransform": (None, [Qureg]),
    "applyTrotterizedPauliStrSumGadget": (None, [Qureg, PauliStrSum, _D, _I, _I]),
    "applyQubitProjector": (None, [Qureg, _I, _I]), "applyMultiQubitProjector": (None, [Qureg, _IP, _IP, _I]),
    "applyQubitMeasurement": (_I, [Qureg, _I]), "applyQubitMeasurementAndGetProb": (_I, [Qureg, _I, C.POINTER(_D)]),
    "applyMultiQubitMeasurement": (_Q, [Qureg, _IP, _I]),
    "applyForcedQubitMeasurement": (_D, [Qureg, _I, _I]),
    "applyForcedMultiQubitMeasurement": (_D, [Qureg, _IP, _IP, _I]),
    "multiplyCompMatr1": (None, [Qureg, _I, CompMatr1]), "multiplyCompMatr2": (None, [Qureg, _I, _I, CompMatr2]),
    "multiplyCompMatr": (None, [Qureg, _IP, _I, CompMatr]), "multiplyDiagMatr1": (None, [Qureg, _I, DiagMatr1]),
    "multiplyPauliStr": (None, [Qureg, PauliStr]), "multiplyPauliGadget": (None, [Qureg, PauliStr, _D]),
    "multiplyPauliStrSum": (None, [Qureg, PauliStrSum, Qureg]),
    # decoherence.h
    "mixDephasing": (None, [Qureg, _I, _D]), "mixTwoQubitDephasing": (None, [Qureg, _I, _I, _D]),
    "mixDepolarising": (None, [Qureg, _I, _D]), "mixTwoQubitDepolarising": (None, [Qureg, _I, _I, _D]),
    "mixDamping": (None, [Qureg, _I, _D]), "mixPaulis": (None, [Qureg, _I, _D, _D, _D]),
    "mixQureg": (None, [Qureg, Qureg, _D]), "mixKrausMap": (None, [Qureg, _IP, _I, KrausMap]),
    "mixSuperOp": (None, [Qureg, _IP, _I, SuperOp]),
    # calculations.h
    "calcTotalProb": (_D, [Qureg]), "calcPurity": (_D, [Qureg]), "calcFidelity": (_D, [Qureg, Qureg]),
    "calcDistance": (_D, [Qureg, Qureg]),
    "calcProbOfBasisState": (_D, [Qureg, _Q]), "calcProbOfQubitOutcome": (_D, [Qureg, _I, _I]),
    "calcProbOfMultiQubitOutcome": (_D, [Qureg, _IP, _IP, _I]),
    "calcProbsOfAllMultiQubitOutcomes": (None, [C.POINTER(_D), Qureg, _IP, _I]),
    "calcExpecPauliStr": (_D, [Qureg, PauliStr]), "calcExpecPauliStrSum": (_D, [Qureg, PauliStrSum]),
    "calcExpecFullStateDiagMatr": (_D, [Qureg, FullStateDiagMatr]),
    "calcExpecFullStateDiagMatrPower": (_D, [Qureg, FullStateDiagMatr, _D]),
    "calcPartialTrace": (Qureg, [Qureg, _IP, _I]), "calcReducedDensityMatrix": (Qureg, [Qureg, _IP, _I]),
}

REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B200_LIB = os.path.join(REPO_ROOT, "quest_b200", "lib", "libQuEST_f32.so" if PRECISION == 1 else "libQuEST.so")
REF_LIB = os.path.join(REPO_ROOT, "oracle", "_ref_f32" if PRECISION == 1 else "_ref", "libQuEST.so")   # the parity oracle (tests / bench only)


def _ints(seq):
    seq = [int(v) for v in seq]
    return (C.c_int * max(len(seq), 1))(*seq), len(seq)


class QuEST:
    """One loaded QuEST library. Methods carry the reference's API names and argument order."""

    def __init__(self, lib_path=B200_LIB):
        if not os.path.exists(lib_path):
            raise FileNotFoundError(f"{lib_path} is missing; run `make` (or __graft_entry__.build()) first")
        self.path = lib_path
        self.lib = C.CDLL(lib_path, mode=C.RTLD_GLOBAL)
        for name, (res, args) in _SIGS.items():
            fn = getattr(self.lib, name)
            fn.restype, fn.argtypes = res, args
        self._keep = []
        # python-side equivalents of the C++-only qcomp-returning functions
        L = self.lib

        def _ret(fn, out_first):
            def f(*args):
                out = qcomp()
                fn(*((C.byref(out),) + args if out_first else args + (C.byref(out),)))
                return out
            return f
        L.getQuregAmp = _ret(L._wrap_getQuregAmp, True)
        L.getDensityQuregAmp = _ret(L._wrap_getDensityQuregAmp, True)
        L.calcInnerProduct = _ret(L._wrap_calcInnerProduct, False)
        L.calcExpecNonHermitianPauliStrSum = _ret(L._wrap_calcExpecNonHermitianPauliStrSum, True)
        L.calcExpecNonHermitianFullStateDiagMatr = _ret(L._wrap_calcExpecNonHermitianFullStateDiagMatr, True)

    def __getattr__(self, name):  # raw access for everything without python sugar below
        if name != "lib" and hasattr(self.lib, name):
            return getattr(self.lib, name)
        raise AttributeError(name)

    # ---- matrices -----------------------------------------------------------------------------------
    @staticmethod
    def getCompMatr1(m):
        m = np.asarray(m, dtype=np.complex128).reshape(2, 2)
        out = CompMatr1(); out.numQubits = 1; out.numRows = 2
        for r in range(2):
            for c in range(2):
                out.elems[r][c] = _qc(m[r, c])
        return out

    @staticmethod
    def getCompMatr2(m):
        m = np.asarray(m, dtype=np.complex128).reshape(4, 4)
        out = CompMatr2(); out.numQubits = 2; out.numRows = 4
        for r in range(4):
            for c in range(4):
                out.elems[r][c] = _qc(m[r, c])
        return out

    @staticmethod
    def getDiagMatr1(d):
        out = DiagMatr1(); out.numQubits = 1; out.numElems = 2
        for i in range(2):
            out.elems[i] = _qc(d[i])
        return out

    @staticmethod
    def getDiagMatr2(d):
        out = DiagMatr2(); out.numQubits = 2; out.numElems = 4
        for i in range(4):
            out.elems[i] = _qc(d[i])
        return out

    @staticmethod
    def _view(ptr, n):
        buf = (c_qreal * (2 * n)).from_address(ptr)
        return np.frombuffer(buf, dtype=np_qcomp)

    def newCompMatr(self, m):
        m = np.asarray(m, dtype=np.complex128)
        k = int(round(np.log2(m.shape[0])))
        out = self.lib.createCompMatr(k)
        self._view(out.cpuElemsFlat, m.size)[:] = m.reshape(-1)   # cpuElems is a 2D alias of this memory
        self.lib.syncCompMatr(out)
        return out

    def newDiagMatr(self, d):
        d = np.asarray(d, dtype=np.complex128)
        k = int(round(np.log2(d.size)))
        out = self.lib.createDiagMatr(k)
        self._view(out.cpuElems, d.size)[:] = d
        self.lib.syncDiagMatr(out)
        return out

    def newFullStateDiagMatr(self, d, custom=None):
        d = np.asarray(d, dtype=np.complex128)
        k = int(round(np.log2(d.size)))
        out = self.lib.createFullStateDiagMatr(k) if custom is None else self.lib.createCustomFullStateDiagMatr(k, *custom)
        arr = np.ascontiguousarray(d, dtype=np_qcomp)
        self.lib.setFullStateDiagMatr(out, 0, arr.ctypes.data, d.size)
        return out

    def newKrausMap(self, ops):
        ops = [np.asarray(o, dtype=np.complex128) for o in ops]
        k = int(round(np.log2(ops[0].shape[0])))
        out = self.lib.createKrausMap(k, len(ops))
        dim = 1 << k
        # map.matrices is qcomp*** : numMatrices pointers to arrays of row pointers
        mats = C.cast(out.matrices, C.POINTER(C.POINTER(C.c_void_p)))
        for n, o in enumerate(ops):
            for r in range(dim):
                self._view(mats[n][r], dim)[:] = o[r, :]
        self.lib.syncKrausMap(out)
        return out

    def newSuperOp(self, m):
        m = np.asarray(m, dtype=np.complex128)
        k = int(round(np.log2(m.shape[0]) / 2))
        out = self.lib.createSuperOp(k)
        self._view(out.cpuElemsFlat, m.size)[:] = m.reshape(-1)
        self.lib.syncSuperOp(out)
        return out

    def getPauliStr(self, paulis, indices):
        """paulis: string over IXYZ (or ints 0..3), one per index in `indices`."""
        codes = ["IXYZ".index(p) if isinstance(p, str) else int(p) for p in paulis]
        ca, n = _ints(codes)
        ia, _ = _ints(indices)
        return self.lib._getPauliStrFromInts(ca, ia, n)

    def newPauliStrSum(self, strings, coeffs):
        n = len(strings)
        sa = (PauliStr * n)(*strings)
        ca = (qcomp * n)(*[_qc(c) for c in coeffs])
        return self.lib.createPauliStrSum(sa, ca, n)

    # ---- amplitudes -----------------------------------------------------------------------------------
    def getAmps(self, qureg):
        """All amplitudes as a flat complex128 array (density matrices: column-major flat vector)."""
        n = qureg.numAmps
        out = np.empty(n, dtype=np_qcomp)
        if qureg.isDensityMatrix:
            # getDensityQuregAmps wants qcomp**; go through the flat statevector view instead
            self.lib.syncQuregFromGpu(qureg) if qureg.isGpuAccelerated else None
            if qureg.isDistributed:
                raise NotImplementedError("use getDensityAmpsDistributed")
            out[:] = self._view(qureg.cpuAmps, n)
        else:
            self.lib.getQuregAmps(out.ctypes.data, qureg, 0, n)
        return out

    def getLocalAmps(self, qureg):
        n = qureg.numAmpsPerNode
        if qureg.isGpuAccelerated:
            self.lib.syncQuregFromGpu(qureg)
        return self._view(qureg.cpuAmps, n).copy()

    def setAmps(self, qureg, amps):
        amps = np.ascontiguousarray(amps, dtype=np_qcomp)
        if qureg.isDensityMatrix:
            self.lib.setDensityQuregFlatAmps(qureg, 0, amps.ctypes.data, amps.size)
        else:
            self.lib.setQuregAmps(qureg, 0, amps.ctypes.data, amps.size)

    def setSeeds(self, seeds):
        arr = (C.c_uint * len(seeds))(*[int(s) for s in seeds])
        self.lib.setSeeds(arr, len(seeds))

    # ---- list-taking calls ------------------------------------------------------------------------------
    def call(self, name, *args):
        """Generic call: python lists of ints become (int*, ...) and complex become qcomp.
        A list argument expands to pointer only; pass its length explicitly as the C API does."""
        conv = []
        for a in args:
            if isinstance(a, (list, tuple, np.ndarray)):
                arr, _ = _ints(a)
                self._keep.append(arr)
                conv.append(arr)
            elif isinstance(a, complex):
                conv.append(_qc(a))
            else:
                conv.append(a)
        out = getattr(self.lib, name)(*conv)
        self._keep.clear()
        if isinstance(out, qcomp):
            return complex(out)
        return out
