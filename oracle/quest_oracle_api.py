"""API-level interpreter over the numpy oracle -- TEST INFRASTRUCTURE ONLY (see quest_oracle.py).

Executes the program format of quest_b200/program.py with oracle/quest_oracle.py, restating the thin
host logic that sits between QuEST's public API and the backend subroutines for NON-distributed Quregs:
  * unitaries on density matrices run twice: ket qubits, then bra qubits (= +numQubits) with the
    conjugated operator                                        (api/operations.cpp:38-62)
  * Pauli tensors / gadgets and their density-matrix sign rules  (api/operations.cpp:1061-1089, 1449-1477;
    core/localiser.cpp:1247-1360)
  * gadget angle -> phase = -angle/2                             (core/utilities.cpp:911-914)
  * Kraus map -> superoperator sum conj(K) (x) K on ket+bra targets (core/utilities.cpp:770-800,
    core/localiser.cpp:1670-1686)
  * measurement / projector renormalisation                      (api/operations.cpp:1740-1827)
Its only purpose is to pin the oracle: tests/test_oracle_golden.py replays the committed programs of
tests/golden/ through this interpreter and compares with what the UNMODIFIED reference library produced.
"""
import numpy as np

from . import quest_oracle as qo
from .quest_rng import MT19937_64


def _dec(m):
    a = np.asarray(m, dtype=np.float64)
    return a[..., 0] + 1j * a[..., 1]


def _arg(a):
    if isinstance(a, dict):
        (k, v), = a.items()
        if k in ("m1", "m2", "m", "d1", "d2", "d", "fsd", "superop"):
            return _dec(v)
        if k == "kraus":
            return [_dec(x) for x in v]
        if k == "c":
            return complex(v[0], v[1])
        if k == "pauli":
            return ("pauli", v[0], list(v[1]))
        if k == "paulisum":
            return ("paulisum", [(t[0], list(t[1]), complex(t[2][0], t[2][1])) for t in v])
        if k == "out_reals":
            return ("out_reals", int(v))
    return a


def _xyz(pauli):
    _, chars, inds = pauli
    x = [q for c, q in zip(chars, inds) if c == "X"]
    y = [q for c, q in zip(chars, inds) if c == "Y"]
    z = [q for c, q in zip(chars, inds) if c == "Z"]
    return x, y, z


def _masks(chars, inds):
    lo = hi = 0
    for c, q in zip(chars, inds):
        code = "IXYZ".index(c)
        if q < 32:
            lo |= code << (2 * q)
        else:
            hi |= code << (2 * (q - 32))
    return lo, hi


class OracleMachine:
    def __init__(self, prog):
        self.prog = prog
        self.q = {}
        self.rng = MT19937_64(prog["seeds"]) if "seeds" in prog else None      # setSeeds (core/randomiser.cpp:59-82)

    # ---- helpers --------------------------------------------------------------------------------
    def _bra(self, st, qubits):
        return [q + st.numQubits for q in qubits]

    def _unitary(self, st, fn, ctrls, states, targs, matr, conj_fn):
        """apply on ket; for density matrices again on bra qubits with the conjugated operator"""
        states = list(states) if states else [1] * len(ctrls)
        fn(st, list(ctrls), states, list(targs), matr)
        if st.isDensityMatrix:
            fn(st, self._bra(st, ctrls), states, self._bra(st, targs), conj_fn(matr))

    @staticmethod
    def _dense(st, ctrls, states, targs, m):
        if len(targs) == 1:
            qo.statevec_anyCtrlOneTargDenseMatr_subA(st, ctrls, states, targs[0], m)
        elif len(targs) == 2:
            qo.statevec_anyCtrlTwoTargDenseMatr_sub(st, ctrls, states, targs[0], targs[1], m)
        else:
            qo.statevec_anyCtrlAnyTargDenseMatr_sub(st, ctrls, states, targs, m, False)

    @staticmethod
    def _diag(st, ctrls, states, targs, d):
        if len(targs) == 1:
            qo.statevec_anyCtrlOneTargDiagMatr_sub(st, ctrls, states, targs[0], d)
        elif len(targs) == 2:
            qo.statevec_anyCtrlTwoTargDiagMatr_sub(st, ctrls, states, targs[0], targs[1], d)
        else:
            qo.statevec_anyCtrlAnyTargDiagMatr_sub(st, ctrls, states, targs, d)

    def _pauliTensor(self, st, ctrls, states, x, y, z, factor):
        if x or y:
            qo.statevector_anyCtrlPauliTensorOrGadget_subA(st, ctrls, states, x, y, z, 0 * factor, 1 * factor)
        else:
            qo.statevector_anyCtrlAnyTargZOrPhaseGadget_sub(st, ctrls, states, z, 1, -1)

    def _pauliGadget(self, st, ctrls, states, x, y, z, phase):
        if x or y:
            qo.statevector_anyCtrlPauliTensorOrGadget_subA(st, ctrls, states, x, y, z, np.cos(phase), 1j * np.sin(phase))
        else:
            qo.statevector_anyCtrlAnyTargZOrPhaseGadget_sub(st, ctrls, states, z, np.exp(1j * phase), np.exp(-1j * phase))

    def applyPauliStr(self, st, ctrls, states, pauli):
        x, y, z = _xyz(pauli)
        states = list(states) if states else [1] * len(ctrls)
        n = st.numQubits
        if st.isDensityMatrix and not ctrls:
            factor = -1 if len(y) % 2 else 1
            self._pauliTensor(st, [], [], x + [q + n for q in x], y + [q + n for q in y], z + [q + n for q in z], factor)
            return
        self._pauliTensor(st, list(ctrls), states, x, y, z, 1)
        if st.isDensityMatrix:
            factor = -1 if len(y) % 2 else 1
            self._pauliTensor(st, self._bra(st, ctrls), states, self._bra(st, x), self._bra(st, y), self._bra(st, z), factor)

    def applyPauliGadget(self, st, ctrls, states, pauli, angle):
        x, y, z = _xyz(pauli)
        states = list(states) if states else [1] * len(ctrls)
        if not (x or y or z) and not ctrls and st.isDensityMatrix:
            return
        phase = -angle / 2
        self._pauliGadget(st, list(ctrls), states, x, y, z, phase)
        if st.isDensityMatrix:
            phase *= 1 if len(y) % 2 else -1
            self._pauliGadget(st, self._bra(st, ctrls), states, self._bra(st, x), self._bra(st, y), self._bra(st, z), phase)

    def applyPhaseGadget(self, st, ctrls, states, targs, angle):
        states = list(states) if states else [1] * len(ctrls)
        phase = -angle / 2
        qo.statevector_anyCtrlAnyTargZOrPhaseGadget_sub(st, list(ctrls), states, list(targs), np.exp(1j * phase), np.exp(-1j * phase))
        if st.isDensityMatrix:
            phase = -phase
            qo.statevector_anyCtrlAnyTargZOrPhaseGadget_sub(st, self._bra(st, ctrls), states, self._bra(st, targs),
                                                            np.exp(1j * phase), np.exp(-1j * phase))

    # ---- op table -------------------------------------------------------------------------------
    def op(self, name, a):
        Q = self.q
        conj = np.conj

        if name in ("applyMultiStateControlledCompMatr1", "applyMultiStateControlledCompMatr2"):
            st = Q[a[0]]; nt = 1 if name.endswith("1") else 2
            self._unitary(st, self._dense, a[1], a[2], list(a[4:4 + nt]), a[4 + nt], conj); return None
        if name in ("applyCompMatr1", "applyCompMatr2"):
            st = Q[a[0]]; nt = 1 if name.endswith("1") else 2
            self._unitary(st, self._dense, [], [], list(a[1:1 + nt]), a[1 + nt], conj); return None
        if name == "applyMultiStateControlledCompMatr":
            self._unitary(Q[a[0]], self._dense, a[1], a[2], a[4], a[6], conj); return None
        if name == "applyCompMatr":
            self._unitary(Q[a[0]], self._dense, [], [], a[1], a[3], conj); return None
        if name == "multiplyCompMatr1":
            self._dense(Q[a[0]], [], [], [a[1]], a[2]); return None
        if name in ("applyMultiStateControlledDiagMatr1", "applyMultiStateControlledDiagMatr2"):
            st = Q[a[0]]; nt = 1 if name.endswith("1") else 2
            self._unitary(st, self._diag, a[1], a[2], list(a[4:4 + nt]), a[4 + nt], conj); return None
        if name == "applyMultiStateControlledDiagMatr":
            self._unitary(Q[a[0]], self._diag, a[1], a[2], a[4], a[6], conj); return None
        if name == "applyMultiStateControlledDiagMatrPower":
            st, ctrls, states, targs, d, expo = Q[a[0]], a[1], a[2], a[4], a[6], a[7]
            states = list(states) if states else [1] * len(ctrls)
            qo.statevec_anyCtrlAnyTargDiagMatr_sub(st, list(ctrls), states, list(targs), d, False, True, expo)
            if st.isDensityMatrix:
                qo.statevec_anyCtrlAnyTargDiagMatr_sub(st, self._bra(st, ctrls), states, self._bra(st, targs), d, True, True, expo)
            return None
        if name in ("applyFullStateDiagMatr", "applyFullStateDiagMatrPower", "multiplyFullStateDiagMatr"):
            st, d = Q[a[0]], a[1]
            hasPower = name.endswith("Power")
            expo = a[2] if hasPower else 1
            if st.isDensityMatrix:
                qo.densmatr_allTargDiagMatr_sub(st, d, hasPower, name.startswith("multiply"), expo)
            else:
                qo.statevec_allTargDiagMatr_sub(st, d, hasPower, expo)
            return None
        if name == "applyHadamard":
            h = np.array([[1, 1], [1, -1]]) / np.sqrt(2)
            self._unitary(Q[a[0]], self._dense, [], [], [a[1]], h, conj); return None
        if name == "applyMultiStateControlledPauliStr":
            self.applyPauliStr(Q[a[0]], a[1], a[2], a[4]); return None
        if name == "applyPauliStr":
            self.applyPauliStr(Q[a[0]], [], [], a[1]); return None
        if name == "applyControlledPauliX":
            self.applyPauliStr(Q[a[0]], [a[1]], [1], ("pauli", "X", [a[2]])); return None
        if name == "applyMultiStateControlledPauliGadget":
            self.applyPauliGadget(Q[a[0]], a[1], a[2], a[4], a[5]); return None
        if name == "applyPauliGadget":
            self.applyPauliGadget(Q[a[0]], [], [], a[1], a[2]); return None
        if name in ("applyRotateX", "applyRotateY", "applyRotateZ"):
            self.applyPauliGadget(Q[a[0]], [], [], ("pauli", name[-1], [a[1]]), a[2]); return None
        if name == "applyMultiStateControlledPhaseGadget":
            self.applyPhaseGadget(Q[a[0]], a[1], a[2], a[4], a[6]); return None
        if name == "applyMultiStateControlledSwap":
            st, ctrls, states, t1, t2 = Q[a[0]], a[1], a[2], a[4], a[5]
            states = list(states) if states else [1] * len(ctrls)
            qo.statevec_anyCtrlSwap_subA(st, list(ctrls), states, t1, t2)
            if st.isDensityMatrix:
                qo.statevec_anyCtrlSwap_subA(st, self._bra(st, ctrls), states, t1 + st.numQubits, t2 + st.numQubits)
            return None
        if name == "applySwap":
            return self.op("applyMultiStateControlledSwap", [a[0], [], [], 0, a[1], a[2]])
        if name == "applyDiagMatr1":
            self._unitary(Q[a[0]], self._diag, [], [], [a[1]], a[2], conj); return None
        if name == "applyPhaseGadget":
            self.applyPhaseGadget(Q[a[0]], [], [], a[1], a[3]); return None
        if name == "applyTwoQubitPhaseShift":
            # api/operations.cpp:1596-1614: a (numTargets-1)-controlled diag{1, e^{i angle}} on targets[0], controls targets[1:]
            d = np.array([1, np.exp(1j * a[3])])
            self._unitary(Q[a[0]], self._diag, [a[2]], [1], [a[1]], d, conj); return None
        if name == "applyFullQuantumFourierTransform":
            # api/operations.cpp:1934-1964: per qubit (top down) Hadamard + ladder of controlled phases, then the swaps
            st = Q[a[0]]; n = st.numQubits
            for t in range(n - 1, -1, -1):
                self.op("applyHadamard", [a[0], t])
                for m in range(t):
                    self.op("applyTwoQubitPhaseShift", [a[0], t, t - m - 1, np.pi / (1 << (m + 1))])
            for t in range(n // 2):
                self.op("applySwap", [a[0], t, n - 1 - t])
            return None
        if name == "setQuregToSuperposition":
            qo.statevec_setQuregToSuperposition_sub(a[0], Q[a[1]], a[2], Q[a[3]], a[4], Q[a[5]]); return None
        if name == "applyMultiQubitProjector":
            st, qubits, outcomes = Q[a[0]], list(a[1]), list(a[2])
            if st.isDensityMatrix:
                qo.densmatr_multiQubitProjector_sub(st, qubits, outcomes, 1.0)
            else:
                qo.statevec_multiQubitProjector_sub(st, qubits, outcomes, 1.0)
            return None
        if name == "applyForcedMultiQubitMeasurement":
            st, qubits, outcomes = Q[a[0]], list(a[1]), list(a[2])
            if st.isDensityMatrix:
                prob = qo.densmatr_calcProbOfMultiQubitOutcome_sub(st, qubits, outcomes)
                qo.densmatr_multiQubitProjector_sub(st, qubits, outcomes, prob)
            else:
                prob = qo.statevec_calcProbOfMultiQubitOutcome_sub(st, qubits, outcomes)
                qo.statevec_multiQubitProjector_sub(st, qubits, outcomes, prob)
            return prob
        if name == "applyQubitMeasurement":
            # api/operations.cpp:1796-1827: two reductions, one uniform draw from the host mt19937_64, projector with the
            # probability of the sampled outcome.  Returns the outcome (an integer: compared bit-exactly).
            st, t = Q[a[0]], a[1]
            if self.rng is None:
                raise NotImplementedError("applyQubitMeasurement needs the program's seeds (default seeds come from std::random_device)")
            calc = qo.densmatr_calcProbOfMultiQubitOutcome_sub if st.isDensityMatrix else qo.statevec_calcProbOfMultiQubitOutcome_sub
            probs = [calc(st, [t], [0]), calc(st, [t], [1])]
            outcome = self.rng.single_qubit_outcome(probs[0])
            proj = qo.densmatr_multiQubitProjector_sub if st.isDensityMatrix else qo.statevec_multiQubitProjector_sub
            proj(st, [t], [outcome], probs[outcome])
            return outcome
        if name == "applyTrotterizedPauliStrSumGadget":
            # api/operations.cpp:1133-1197 (orders 1 and 2; higher orders recurse through the same first-order sweep)
            terms, angle, order, reps = a[1][1], a[2], a[3], a[4]
            if angle == 0:
                return None

            def first_order(ang, reverse):
                seq = reversed(terms) if reverse else terms
                for chars, inds, coeff in seq:
                    self.op("applyPauliGadget", [a[0], ("pauli", chars, inds), 2 * ang * complex(coeff).real])

            for _ in range(reps):
                ang = angle / reps
                if order == 1:
                    first_order(ang, False)
                elif order == 2:
                    first_order(ang / 2, False); first_order(ang / 2, True)
                else:
                    raise NotImplementedError("Trotter order > 2 is not restated")
            return None
        # ---- channels ----
        if name == "mixDephasing":
            qo.densmatr_oneQubitDephasing_subA(Q[a[0]], a[1], a[2]); return None
        if name == "mixTwoQubitDephasing":
            qo.densmatr_twoQubitDephasing_subA(Q[a[0]], a[1], a[2], a[3]); return None
        if name == "mixDepolarising":
            qo.densmatr_oneQubitDepolarising_subA(Q[a[0]], a[1], a[2]); return None
        if name == "mixTwoQubitDepolarising":
            q1, q2 = sorted([a[1], a[2]])
            qo.densmatr_twoQubitDepolarising_subA(Q[a[0]], q1, q2, a[3])
            qo.densmatr_twoQubitDepolarising_subB(Q[a[0]], q1, q2, a[3]); return None
        if name == "mixDamping":
            qo.densmatr_oneQubitDamping_subA(Q[a[0]], a[1], a[2]); return None
        if name == "mixPaulis":
            pX, pY, pZ = a[2], a[3], a[4]
            qo.densmatr_oneQubitPauliChannel_subA(Q[a[0]], a[1], 1 - pX - pY - pZ, pX, pY, pZ); return None
        if name == "mixQureg":
            out, other, prob = Q[a[0]], Q[a[1]], a[2]
            if other.isDensityMatrix:
                qo.densmatr_mixQureg_subA(1 - prob, out, prob, other)
            else:
                qo.densmatr_mixQureg_subB(1 - prob, out, prob, other)
            return None
        if name in ("mixKrausMap", "mixSuperOp"):
            st, targs = Q[a[0]], list(a[1])
            if name == "mixKrausMap":
                sup = sum(np.kron(np.conj(k), k) for k in a[3])
            else:
                sup = a[3]
            qo.statevec_anyCtrlAnyTargDenseMatr_sub(st, [], [], targs + self._bra(st, targs), sup, False)
            return None
        # ---- calculations ----
        if name == "calcTotalProb":
            st = Q[a[0]]
            return qo.densmatr_calcTotalProb_sub(st) if st.isDensityMatrix else qo.statevec_calcTotalProb_sub(st)
        if name == "calcPurity":
            st = Q[a[0]]
            p = qo.statevec_calcTotalProb_sub(st)
            return p if st.isDensityMatrix else p * p
        if name in ("calcProbOfMultiQubitOutcome", "calcProbOfQubitOutcome"):
            st = Q[a[0]]
            qubits, outcomes = ([a[1]], [a[2]]) if name.endswith("QubitOutcome") and not isinstance(a[1], list) else (list(a[1]), list(a[2]))
            fn = qo.densmatr_calcProbOfMultiQubitOutcome_sub if st.isDensityMatrix else qo.statevec_calcProbOfMultiQubitOutcome_sub
            return fn(st, qubits, outcomes)
        if name == "calcProbsOfAllMultiQubitOutcomes":
            st = Q[a[1]]
            fn = qo.densmatr_calcProbsOfAllMultiQubitOutcomes_sub if st.isDensityMatrix else qo.statevec_calcProbsOfAllMultiQubitOutcomes_sub
            return list(fn(st, list(a[2])))
        if name == "calcInnerProduct":
            z = qo.statevec_calcInnerProduct_sub(Q[a[0]], Q[a[1]])
            return [z.real, z.imag]
        if name == "calcFidelity":
            rho, psi = Q[a[0]], Q[a[1]]
            if rho.isDensityMatrix:
                return qo.densmatr_calcFidelityWithPureState_sub(rho, psi, False).real
            return abs(qo.statevec_calcInnerProduct_sub(rho, psi)) ** 2
        if name == "calcDistance":
            A, B = Q[a[0]], Q[a[1]]
            if A.isDensityMatrix:
                return float(np.sqrt(qo.densmatr_calcHilbertSchmidtDistance_sub(A, B)))
            # Bures distance between pure states: sqrt(2 - 2|<a|b>|)  (api/calculations.cpp)
            mag = abs(qo.statevec_calcInnerProduct_sub(A, B))
            return float(np.sqrt(max(0.0, 2 - 2 * min(mag, 1.0))))
        if name == "calcExpecPauliStr":
            st = Q[a[0]]; x, y, z = _xyz(a[1])
            if st.isDensityMatrix:
                if not (x or y):
                    return qo.densmatr_calcExpecAnyTargZ_sub(st, z).real if z else qo.densmatr_calcTotalProb_sub(st)
                return qo.densmatr_calcExpecPauliStr_sub(st, x, y, z).real
            if not (x or y or z):
                return qo.statevec_calcTotalProb_sub(st)
            if not (x or y):
                return qo.statevec_calcExpecAnyTargZ_sub(st, z)
            return qo.statevec_calcExpecPauliStr_subA(st, x, y, z).real
        if name == "calcExpecPauliStrSum":
            st = Q[a[0]]; total = 0
            for chars, inds, coeff in a[1][1]:
                total += coeff * self.op("calcExpecPauliStr", [a[0], ("pauli", chars, inds)])
            return complex(total).real
        if name in ("calcExpecFullStateDiagMatr", "calcExpecFullStateDiagMatrPower"):
            st, d = Q[a[0]], a[1]
            hasPower = name.endswith("Power")
            expo = a[2] if hasPower else 1
            fn = qo.densmatr_calcExpecFullStateDiagMatr_sub if st.isDensityMatrix else qo.statevec_calcExpecFullStateDiagMatr_sub
            return fn(st, d, hasPower, hasPower, expo).real
        if name == "calcPartialTrace":
            st, targs = Q[a[0]], list(a[1])
            out = qo.new_state(st.numQubits - len(targs), 1)
            qo.densmatr_partialTrace_sub(st, out, targs, self._bra(st, targs))
            rname = f"_ret{self.opIndex}"
            Q[rname] = out
            return rname
        if name == "setQuregToPauliStrSum":
            st = Q[a[0]]
            terms = a[1][1]
            qo.densmatr_setAmpsToPauliStrSum_sub(st, [t[2] for t in terms], [_masks(t[0], t[1]) for t in terms]); return None
        if name == "initDebugState":
            qo.statevec_initDebugState_sub(Q[a[0]]); return None
        if name == "initPlusState":
            st = Q[a[0]]
            dim = 1 << st.numQubits
            qo.statevec_initUniformState_sub(st, 1 / dim if st.isDensityMatrix else 1 / np.sqrt(dim)); return None
        raise NotImplementedError(f"oracle API interpreter has no rule for {name}")

    def run(self):
        prog = self.prog
        for name, spec in prog["quregs"].items():
            st = qo.new_state(int(spec["n"]), int(spec.get("dm", 0)))
            init = spec.get("init", "zero")
            if init == "debug":
                qo.statevec_initDebugState_sub(st)
            elif init == "zero":
                st.amps[0] = 1
            elif init == "plus":
                dim = 1 << st.numQubits
                st.amps[:] = 1 / dim if st.isDensityMatrix else 1 / np.sqrt(dim)
            elif init == "blank":
                pass
            elif isinstance(init, (list, tuple)) and init[0] == "amps":
                st.amps[:] = _dec(init[1])
            else:
                raise NotImplementedError(init)
            self.q[name] = st
        results = []
        for self.opIndex, op in enumerate(prog["ops"]):
            results.append(self.op(op[0], [_arg(x) for x in op[1:]]))
        dumps = {name: self.q[name].amps.copy() for name in prog.get("dump", [])}
        return {"results": results, "dumps": dumps}


def run_program(prog):
    return OracleMachine(prog).run()
