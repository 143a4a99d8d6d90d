"""Seeded program generators (quest_b200/program.py format) shared by the golden-fixture generator,
the CPU oracle tests and the GPU parity tests.  Everything is derived from numpy's PCG64 with an explicit
seed -- never from QuEST's own RNG -- so the reference library and the backend see identical inputs.

Random unitaries follow the reference tests' recipe (tests/utils/random.cpp:348-376): QR of a complex
Gaussian matrix with R's diagonal phases divided out; Kraus maps are random unitaries scaled by the
square roots of normalised random weights (tests/utils/random.cpp:405-430).
"""
import numpy as np

from quest_b200.program import enc_c, enc_mat


def rand_unitary(rng, dim):
    z = rng.normal(size=(dim, dim)) + 1j * rng.normal(size=(dim, dim))
    q, r = np.linalg.qr(z)
    d = np.diag(r)
    return q * (d / np.abs(d))


def rand_kraus(rng, dim, num_ops):
    w = rng.random(num_ops)
    w /= w.sum()
    return [np.sqrt(wi) * rand_unitary(rng, dim) for wi in w]


def rand_diag_unitary(rng, dim):
    return np.exp(1j * rng.uniform(0, 2 * np.pi, size=dim))


def rand_state(rng, n):
    v = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    return v / np.linalg.norm(v)


def rand_density(rng, n, rank=3):
    dim = 1 << n
    rho = np.zeros((dim, dim), dtype=np.complex128)
    w = rng.random(rank); w /= w.sum()
    for wi in w:
        v = rand_state(rng, n)
        rho += wi * np.outer(v, v.conj())
    return rho.reshape(-1, order="F")      # column-major flat: index = row + col * dim


def _pick(rng, n, k):
    return [int(q) for q in rng.choice(n, size=k, replace=False)]


def _ctrl_targ(rng, n, nc, nt):
    qs = _pick(rng, n, nc + nt)
    ctrls, targs = qs[:nc], qs[nc:]
    states = [int(s) for s in rng.integers(0, 2, size=nc)]
    return ctrls, states, targs


def _pauli(rng, n, k, alphabet="XYZ"):
    qs = _pick(rng, n, k)
    chars = "".join(rng.choice(list(alphabet), size=k))
    return chars, qs


def gates_program(n, seed, dm=0, num_rounds=2, max_ctrls=2, init="debug"):
    """Every unitary-gate family of the hot path with random controls / control-states / targets."""
    rng = np.random.default_rng(seed)
    ops = []
    q = "rho" if dm else "psi"
    for _ in range(num_rounds):
        for nc in range(0, max_ctrls + 1):
            c, s, t = _ctrl_targ(rng, n, nc, 1)
            ops.append(["applyMultiStateControlledCompMatr1", q, c, s, nc, t[0], {"m1": enc_mat(rand_unitary(rng, 2))}])
            c, s, t = _ctrl_targ(rng, n, nc, 2)
            ops.append(["applyMultiStateControlledCompMatr2", q, c, s, nc, t[0], t[1], {"m2": enc_mat(rand_unitary(rng, 4))}])
            for nt in (1, 2, 3):
                if nc + nt > n:
                    continue
                c, s, t = _ctrl_targ(rng, n, nc, nt)
                ops.append(["applyMultiStateControlledCompMatr", q, c, s, nc, t, nt, {"m": enc_mat(rand_unitary(rng, 1 << nt))}])
                c, s, t = _ctrl_targ(rng, n, nc, nt)
                ops.append(["applyMultiStateControlledDiagMatr", q, c, s, nc, t, nt, {"d": enc_mat(rand_diag_unitary(rng, 1 << nt))}])
            c, s, t = _ctrl_targ(rng, n, nc, 1)
            ops.append(["applyMultiStateControlledDiagMatr1", q, c, s, nc, t[0], {"d1": enc_mat(rand_diag_unitary(rng, 2))}])
            c, s, t = _ctrl_targ(rng, n, nc, 2)
            ops.append(["applyMultiStateControlledDiagMatr2", q, c, s, nc, t[0], t[1], {"d2": enc_mat(rand_diag_unitary(rng, 4))}])
            c, s, t = _ctrl_targ(rng, n, nc, 2)
            ops.append(["applyMultiStateControlledDiagMatrPower", q, c, s, nc, t, 2, {"d": enc_mat(rand_diag_unitary(rng, 4))},
                        {"c": enc_c(rng.uniform(-2, 2))}])
            for k in (1, 2, 3):
                if nc + k > n:
                    continue
                c, s, t = _ctrl_targ(rng, n, nc, k)
                chars = "".join(rng.choice(list("XYZ"), size=k))
                ops.append(["applyMultiStateControlledPauliStr", q, c, s, nc, {"pauli": [chars, t]}])
                c, s, t = _ctrl_targ(rng, n, nc, k)
                chars = "".join(rng.choice(list("XYZ"), size=k))
                ops.append(["applyMultiStateControlledPauliGadget", q, c, s, nc, {"pauli": [chars, t]}, float(rng.uniform(-3, 3))])
                c, s, t = _ctrl_targ(rng, n, nc, k)
                ops.append(["applyMultiStateControlledPhaseGadget", q, c, s, nc, t, k, float(rng.uniform(-3, 3))])
            c, s, t = _ctrl_targ(rng, n, nc, 1)
            ops.append(["applyMultiStateControlledPauliStr", q, c, s, nc, {"pauli": ["Z", t]}])
            c, s, t = _ctrl_targ(rng, n, nc, 2)
            ops.append(["applyMultiStateControlledSwap", q, c, s, nc, t[0], t[1]])
        ops.append(["applyFullStateDiagMatr", q, {"fsd": enc_mat(rand_diag_unitary(rng, 1 << n))}])
        ops.append(["applyFullStateDiagMatrPower", q, {"fsd": enc_mat(rand_diag_unitary(rng, 1 << n))}, {"c": enc_c(0.7)}])
        ops.append(["applyHadamard", q, int(rng.integers(n))])
        ops.append(["applyRotateX", q, int(rng.integers(n)), float(rng.uniform(0, 6))])
        ops.append(["applyRotateY", q, int(rng.integers(n)), float(rng.uniform(0, 6))])
        ops.append(["applyRotateZ", q, int(rng.integers(n)), float(rng.uniform(0, 6))])
        a, b = _pick(rng, n, 2)
        ops.append(["applyControlledPauliX", q, a, b])
    ops.append(["calcTotalProb", q])
    return {"quregs": {q: {"n": n, "dm": dm, "init": init}}, "ops": ops, "dump": [q]}


def big_dense_program(n, seed, nt, nc=1):
    """k-target dense matrices beyond the register kernels (k >= 6 uses the shared-memory kernel)."""
    rng = np.random.default_rng(seed)
    ops = []
    for _ in range(2):
        c, s, t = _ctrl_targ(rng, n, nc, nt)
        ops.append(["applyMultiStateControlledCompMatr", "psi", c, s, nc, t, nt, {"m": enc_mat(rand_unitary(rng, 1 << nt))}])
        t = _pick(rng, n, nt)
        ops.append(["applyCompMatr", "psi", t, nt, {"m": enc_mat(rand_unitary(rng, 1 << nt))}])
    ops.append(["calcTotalProb", "psi"])
    return {"quregs": {"psi": {"n": n, "init": ["amps", enc_mat(rand_state(rng, n))]}}, "ops": ops, "dump": ["psi"]}


def calcs_program_sv(n, seed):
    rng = np.random.default_rng(seed)
    psi = rand_state(rng, n)
    phi = rand_state(rng, n)
    ops = [["calcTotalProb", "psi"], ["calcPurity", "psi"], ["calcInnerProduct", "psi", "phi"],
           ["calcFidelity", "psi", "phi"], ["calcDistance", "psi", "phi"]]
    for k in (1, 2, 3):
        qs = _pick(rng, n, k)
        ops.append(["calcProbOfMultiQubitOutcome", "psi", qs, [int(b) for b in rng.integers(0, 2, size=k)], k])
        ops.append(["calcProbsOfAllMultiQubitOutcomes", {"out_reals": 1 << k}, "psi", _pick(rng, n, k), k])
    for k in (1, 2, 4):
        chars, qs = _pauli(rng, n, k)
        ops.append(["calcExpecPauliStr", "psi", {"pauli": [chars, qs]}])
    chars, qs = _pauli(rng, n, 3, "Z")
    ops.append(["calcExpecPauliStr", "psi", {"pauli": [chars, qs]}])
    terms = []
    for _ in range(6):
        chars, qs = _pauli(rng, n, int(rng.integers(1, min(n, 4) + 1)), "IXYZ")
        terms.append([chars, qs, enc_c(rng.uniform(-1, 1))])
    ops.append(["calcExpecPauliStrSum", "psi", {"paulisum": terms}])
    d = rng.uniform(0.1, 2, size=1 << n)
    ops.append(["calcExpecFullStateDiagMatr", "psi", {"fsd": enc_mat(d)}])
    ops.append(["calcExpecFullStateDiagMatrPower", "psi", {"fsd": enc_mat(d)}, 1.5])
    qs = _pick(rng, n, 2)
    ops.append(["applyForcedMultiQubitMeasurement", "psi", qs, [1, 0], 2])
    ops.append(["calcTotalProb", "psi"])
    ops.append(["setQuregToSuperposition", {"c": enc_c(0.3 + 0.1j)}, "psi", {"c": enc_c(-0.2j)}, "phi", {"c": enc_c(0.5)}, "phi"])
    qs = _pick(rng, n, 2)
    ops.append(["applyMultiQubitProjector", "phi", qs, [0, 1], 2])
    return {"quregs": {"psi": {"n": n, "init": ["amps", enc_mat(psi)]}, "phi": {"n": n, "init": ["amps", enc_mat(phi)]}},
            "ops": ops, "dump": ["psi", "phi"]}


def channels_program_dm(n, seed):
    rng = np.random.default_rng(seed)
    rho = rand_density(rng, n)
    sigma = rand_density(rng, n)
    psi = rand_state(rng, n)
    ops = []
    for q in range(n):
        ops.append(["mixDephasing", "rho", q, float(rng.uniform(0, 0.5))])
        ops.append(["mixDepolarising", "rho", q, float(rng.uniform(0, 0.75))])
        ops.append(["mixDamping", "rho", q, float(rng.uniform(0, 1))])
        p = rng.uniform(0, 0.2, size=3)
        ops.append(["mixPaulis", "rho", q, float(p[0]), float(p[1]), float(p[2])])
    for _ in range(3):
        a, b = _pick(rng, n, 2)
        ops.append(["mixTwoQubitDephasing", "rho", a, b, float(rng.uniform(0, 0.75))])
        a, b = _pick(rng, n, 2)
        ops.append(["mixTwoQubitDepolarising", "rho", a, b, float(rng.uniform(0, 15 / 16))])
    ops.append(["mixKrausMap", "rho", _pick(rng, n, 1), 1, {"kraus": [enc_mat(k) for k in rand_kraus(rng, 2, 3)]}])
    ops.append(["mixKrausMap", "rho", _pick(rng, n, 2), 2, {"kraus": [enc_mat(k) for k in rand_kraus(rng, 4, 4)]}])
    ops.append(["mixQureg", "rho", "sigma", 0.3])
    ops.append(["mixQureg", "rho", "psi", 0.2])
    ops.append(["calcTotalProb", "rho"]); ops.append(["calcPurity", "rho"])
    ops.append(["calcFidelity", "rho", "psi"]); ops.append(["calcDistance", "rho", "sigma"])
    for k in (1, 2):
        qs = _pick(rng, n, k)
        ops.append(["calcProbOfMultiQubitOutcome", "rho", qs, [int(b) for b in rng.integers(0, 2, size=k)], k])
        ops.append(["calcProbsOfAllMultiQubitOutcomes", {"out_reals": 1 << k}, "rho", _pick(rng, n, k), k])
    for k in (1, 2, 3):
        chars, qs = _pauli(rng, n, min(k, n))
        ops.append(["calcExpecPauliStr", "rho", {"pauli": [chars, qs]}])
    chars, qs = _pauli(rng, n, 2, "Z")
    ops.append(["calcExpecPauliStr", "rho", {"pauli": [chars, qs]}])
    d = rng.uniform(0.1, 2, size=1 << n)
    ops.append(["calcExpecFullStateDiagMatr", "rho", {"fsd": enc_mat(d)}])
    ops.append(["applyFullStateDiagMatr", "sigma", {"fsd": enc_mat(rand_diag_unitary(rng, 1 << n))}])
    ops.append(["calcPartialTrace", "rho", _pick(rng, n, 1), 1])
    qs = _pick(rng, n, 2)
    ops.append(["applyMultiQubitProjector", "rho", qs, [1, 0], 2])
    terms = []
    for _ in range(4):
        chars, qs = _pauli(rng, n, 2, "IXYZ")
        terms.append([chars, qs, enc_c(rng.uniform(-1, 1) + 1j * rng.uniform(-1, 1))])
    ops.append(["setQuregToPauliStrSum", "sigma", {"paulisum": terms}])
    pt = next(i for i, o in enumerate(ops) if o[0] == "calcPartialTrace")
    return {"quregs": {"rho": {"n": n, "dm": 1, "init": ["amps", enc_mat(rho)]},
                       "sigma": {"n": n, "dm": 1, "init": ["amps", enc_mat(sigma)]},
                       "psi": {"n": n, "init": ["amps", enc_mat(psi)]}},
            "ops": ops, "dump": ["rho", "sigma", "_ret%d" % pt]}


# ---- the BASELINE.json configurations, parametrised by size ------------------------
def cfg1_program(n=20, seed=12345, num_gates=200):
    """cfg 1: H on all qubits, then random {H, CNOT, RotateX, applyCompMatr1}."""
    rng = np.random.default_rng(seed)
    ops = [["applyHadamard", "psi", q] for q in range(n)]
    for _ in range(num_gates):
        kind = int(rng.integers(4))
        if kind == 0:
            ops.append(["applyHadamard", "psi", int(rng.integers(n))])
        elif kind == 1:
            a, b = _pick(rng, n, 2)
            ops.append(["applyControlledPauliX", "psi", a, b])
        elif kind == 2:
            ops.append(["applyRotateX", "psi", int(rng.integers(n)), float(rng.uniform(0, 2 * np.pi))])
        else:
            ops.append(["applyCompMatr1", "psi", int(rng.integers(n)), {"m1": enc_mat(rand_unitary(rng, 2))}])
    ops.append(["calcTotalProb", "psi"])
    for q in range(n):
        ops.append(["calcProbOfQubitOutcome", "psi", q, 0])
    return {"quregs": {"psi": {"n": n, "init": "zero"}}, "ops": ops, "dump": ["psi"]}


def cfg2_program(n=30, seed=20302, num_gates=200, dump=True):
    """cfg 2: full QFT of |0..0>, then random dense 1- and 2-qubit gates on uniformly random targets."""
    rng = np.random.default_rng(seed)
    ops = [["applyFullQuantumFourierTransform", "psi"]]
    for _ in range(num_gates):
        if rng.integers(2):
            ops.append(["applyCompMatr1", "psi", int(rng.integers(n)), {"m1": enc_mat(rand_unitary(rng, 2))}])
        else:
            a, b = _pick(rng, n, 2)
            ops.append(["applyCompMatr2", "psi", a, b, {"m2": enc_mat(rand_unitary(rng, 4))}])
    ops.append(["calcTotalProb", "psi"])
    ops.append(["calcProbOfQubitOutcome", "psi", n - 1, 0])
    return {"quregs": {"psi": {"n": n, "init": "zero"}}, "ops": ops, "dump": ["psi"] if dump else []}


def cfg4_program(n=14, seed=14014, layers=10, dump=True):
    """cfg 4: noisy density-matrix circuit: H layer, CNOT chain, depolarising, 2-qubit Kraus maps, projector."""
    rng = np.random.default_rng(seed)
    kraus = [enc_mat(k) for k in rand_kraus(rng, 4, 4)]
    ops = []
    for _ in range(layers):
        ops += [["applyHadamard", "rho", q] for q in range(n)]
        ops += [["applyControlledPauliX", "rho", q, q + 1] for q in range(n - 1)]
        ops += [["mixDepolarising", "rho", q, 0.01] for q in range(n)]
        ops += [["mixKrausMap", "rho", [q, q + 1], 2, {"kraus": kraus}] for q in range(0, n - 1, 2)]
    ops.append(["applyMultiQubitProjector", "rho", [0, n - 1], [0, 1], 2])
    ops.append(["calcTotalProb", "rho"])
    ops.append(["calcPurity", "rho"])
    return {"quregs": {"rho": {"n": n, "dm": 1, "init": "plus"}}, "ops": ops, "dump": ["rho"] if dump else []}


def cfg5_program(n=28, seed=28200, num_terms=200, dump=True):
    """cfg 5: second-order Trotterised evolution under a random Pauli Hamiltonian, then its expectation value."""
    rng = np.random.default_rng(seed)
    terms = []
    for _ in range(num_terms):
        chars = "".join(rng.choice(list("IXYZ"), size=n, p=[0.5, 1 / 6, 1 / 6, 1 / 6]))
        if set(chars) == {"I"}:
            chars = "Z" + chars[1:]
        terms.append([chars, list(range(n)), enc_c(rng.uniform(-1, 1))])
    ops = [["applyTrotterizedPauliStrSumGadget", "psi", {"paulisum": terms}, 0.1, 2, 1],
           ["calcExpecPauliStrSum", "psi", {"paulisum": terms}],
           ["calcTotalProb", "psi"]]
    return {"quregs": {"psi": {"n": n, "init": "plus"}}, "ops": ops, "dump": ["psi"] if dump else []}


def measurement_program(n, seed):
    """Measurement outcomes must be bit-exact: both libraries are seeded identically (setSeeds) and each
    measurement draws exactly one uniform from the host mt19937_64 (core/randomiser.cpp:125-166)."""
    rng = np.random.default_rng(seed)
    ops = [["applyHadamard", "psi", q] for q in range(n)]
    for _ in range(3 * n):
        ops.append(["applyCompMatr1", "psi", int(rng.integers(n)), {"m1": enc_mat(rand_unitary(rng, 2))}])
        a, b = _pick(rng, n, 2)
        ops.append(["applyControlledPauliX", "psi", a, b])
    for q in range(n):
        ops.append(["applyQubitMeasurement", "psi", q])
    ops.append(["calcTotalProb", "psi"])
    return {"seeds": [int(seed), 7], "quregs": {"psi": {"n": n, "init": "zero"}}, "ops": ops, "dump": ["psi"]}


def relabel_program(n, seed, num_ops=160, reads=True):
    """Stress of the lazy qubit relabelling (quest_b200/shim/localiser_b200.cpp): uncontrolled SWAPs (pure relabelling)
    and dense gates on every qubit (on several GPUs: targets on rank bits, pulled into the shard and left there)
    interleaved with the relabelling-aware gates and reductions, and with operations that must first restore the
    canonical order (expectation values, full-state diagonals, amplitude reads at the end)."""
    rng = np.random.default_rng(seed)
    ops = []
    for _ in range(num_ops):
        r = int(rng.integers(14 if reads else 12))        # reads=False: gates only, so the backend's queue is never flushed by a read
        if r < 3:
            a, b = _pick(rng, n, 2)
            ops.append(["applySwap", "psi", a, b])
        elif r < 5:
            ops.append(["applyCompMatr1", "psi", int(rng.integers(n)), {"m1": enc_mat(rand_unitary(rng, 2))}])
        elif r < 7:
            a, b = _pick(rng, n, 2)
            ops.append(["applyCompMatr2", "psi", a, b, {"m2": enc_mat(rand_unitary(rng, 4))}])
        elif r == 7:
            c, s, t = _ctrl_targ(rng, n, 1, 2)
            ops.append(["applyMultiStateControlledCompMatr2", "psi", c, s, 1, t[0], t[1], {"m2": enc_mat(rand_unitary(rng, 4))}])
        elif r == 8:
            nt = min(3, n)
            ops.append(["applyCompMatr", "psi", _pick(rng, n, nt), nt, {"m": enc_mat(rand_unitary(rng, 1 << nt))}])
        elif r == 9:
            a, b = _pick(rng, n, 2)
            ops.append(["applyTwoQubitPhaseShift", "psi", a, b, float(rng.uniform(0, 6))])
            ops.append(["applyDiagMatr1", "psi", int(rng.integers(n)), {"d1": enc_mat(rand_diag_unitary(rng, 2))}])
        elif r == 10:
            k = int(rng.integers(1, min(n, 4) + 1))
            chars, qs = _pauli(rng, n, k)
            ops.append(["applyPauliGadget", "psi", {"pauli": [chars, qs]}, float(rng.uniform(-3, 3))])
            chars, qs = _pauli(rng, n, k)
            ops.append(["applyPauliStr", "psi", {"pauli": [chars, qs]}])
        elif r == 11:
            k = int(rng.integers(1, min(n, 3) + 1))
            ops.append(["applyPhaseGadget", "psi", _pick(rng, n, k), k, float(rng.uniform(-3, 3))])
            c, s, t = _ctrl_targ(rng, n, 1, 2)
            ops.append(["applyMultiStateControlledSwap", "psi", c, s, 1, t[0], t[1]])
        elif r == 12:
            ops.append(["calcProbOfQubitOutcome", "psi", int(rng.integers(n)), int(rng.integers(2))])
            qs = _pick(rng, n, 2)
            ops.append(["calcProbOfMultiQubitOutcome", "psi", qs, [int(b) for b in rng.integers(0, 2, size=2)], 2])
        else:
            which = int(rng.integers(4))
            if which == 0:
                chars, qs = _pauli(rng, n, min(n, 3))
                ops.append(["calcExpecPauliStr", "psi", {"pauli": [chars, qs]}])
            elif which == 1:
                ops.append(["calcProbsOfAllMultiQubitOutcomes", {"out_reals": 4}, "psi", _pick(rng, n, 2), 2])
            elif which == 2:
                ops.append(["applyFullStateDiagMatr", "psi", {"fsd": enc_mat(rand_diag_unitary(rng, 1 << n))}])
            else:
                ops.append(["applyMultiQubitProjector", "psi", _pick(rng, n, 1), [int(rng.integers(2))], 1])
                ops.append(["calcTotalProb", "psi"])
    ops.append(["calcTotalProb", "psi"])
    return {"quregs": {"psi": {"n": n, "init": ["amps", enc_mat(rand_state(rng, n))]}}, "ops": ops, "dump": ["psi"]}


def rank_divergence_program(n, logp, seed, num_rounds=12):
    """Operations after which the ranks' deferred-gate queues DIFFER -- setQuregAmps of a few amplitudes (runs on the
    owning ranks only), a controlled SWAP of two rank-bit qubits (runs on the ranks whose two bits differ), gates with
    rank-bit controls (dropped on half of the ranks) -- each followed by dense gates on rank-bit qubits, whose swap-in
    victim must still be chosen identically on every rank.  n - logp local qubits; the top logp are rank bits."""
    rng = np.random.default_rng(seed)
    nl = n - logp
    ops = []
    for _ in range(num_rounds):
        # leave gates in the queue that touch the HIGH suffix qubits (the preferred swap-in victims)
        for _ in range(int(rng.integers(1, 4))):
            ops.append(["applyCompMatr1", "psi", int(rng.integers(max(0, nl - 3), nl)), {"m1": enc_mat(rand_unitary(rng, 2))}])
        which = int(rng.integers(4))
        if which == 0:
            k = int(rng.integers(1, 4))
            start = int(rng.integers(0, (1 << n) - k))
            vals = rng.normal(size=k) + 1j * rng.normal(size=k)
            ops.append(["setQuregAmps", "psi", start, {"amps": enc_mat(0.1 * vals)}, k])
        elif which == 1 and logp >= 2:
            a, b = [nl + int(x) for x in rng.choice(logp, size=2, replace=False)]
            c = int(rng.integers(nl))
            ops.append(["applyMultiStateControlledSwap", "psi", [c], [int(rng.integers(2))], 1, a, b])
        elif which == 2:
            c = nl + int(rng.integers(logp))
            ops.append(["applyMultiStateControlledCompMatr1", "psi", [c], [int(rng.integers(2))], 1, int(rng.integers(nl)),
                        {"m1": enc_mat(rand_unitary(rng, 2))}])
        else:
            a, b = _pick(rng, n, 2)
            ops.append(["applySwap", "psi", a, b])
        # now gates whose targets sit on rank bits (or were relabelled there)
        for _ in range(int(rng.integers(1, 3))):
            ops.append(["applyCompMatr1", "psi", nl + int(rng.integers(logp)), {"m1": enc_mat(rand_unitary(rng, 2))}])
        a, b = _pick(rng, n, 2)
        ops.append(["applyCompMatr2", "psi", a, b, {"m2": enc_mat(rand_unitary(rng, 4))}])
    ops.append(["calcTotalProb", "psi"])
    return {"quregs": {"psi": {"n": n, "init": ["amps", enc_mat(rand_state(rng, n))]}}, "ops": ops, "dump": ["psi"]}


def coevolution_program(n, seed, num_ops=60):
    """The reference's psi / rho co-evolution pattern (tests/integration/densitymatrix.cpp:68-163): the same random
    gates applied alternately to a statevector and to a density matrix, which are then compared through
    calcFidelity.  For the backend this interleaves fusable gates on two (here three) different Quregs."""
    rng = np.random.default_rng(seed)
    ops = []
    for _ in range(num_ops):
        r = int(rng.integers(5))
        if r == 0:
            t = int(rng.integers(n))
            for name in ("psi", "rho", "phi"):
                ops.append(["applyHadamard", name, t])
        elif r == 1:
            t = int(rng.integers(n)); m = {"m1": enc_mat(rand_unitary(rng, 2))}
            for name in ("psi", "rho", "phi"):
                ops.append(["applyCompMatr1", name, t, m])
        elif r == 2:
            a, b = _pick(rng, n, 2); m = {"m2": enc_mat(rand_unitary(rng, 4))}
            for name in ("psi", "rho", "phi"):
                ops.append(["applyCompMatr2", name, a, b, m])
        elif r == 3:
            a, b = _pick(rng, n, 2)
            for name in ("psi", "rho", "phi"):
                ops.append(["applyControlledPauliX", name, a, b])
        elif r == 4:
            a, b = _pick(rng, n, 2); th = float(rng.uniform(0, 6))
            for name in ("psi", "rho", "phi"):
                ops.append(["applyTwoQubitPhaseShift", name, a, b, th])
    ops.append(["calcFidelity", "rho", "psi"])
    ops.append(["calcTotalProb", "phi"])
    return {"quregs": {"psi": {"n": n, "init": "zero"}, "rho": {"n": n, "dm": 1, "init": "zero"}, "phi": {"n": n, "init": "plus"}},
            "ops": ops, "dump": ["psi", "rho", "phi"]}


def two_quregs_program(n, seed_a, seed_b, num_ops=40):
    """two statevectors evolved in lock-step with a syncQuESTEnv() in the middle and at the end (with several GPUs: two live
    qubit maps, restored in creation order on every rank)"""
    a, b = relabel_program(n, seed_a, num_ops=num_ops, reads=False), relabel_program(n, seed_b, num_ops=num_ops, reads=False)
    ops = []
    for x, y in zip(a["ops"], b["ops"]):
        ops.append(x)
        ops.append([y[0], "chi"] + list(y[2:]))
    ops.insert(len(ops) // 2, ["syncQuESTEnv"])
    ops.append(["syncQuESTEnv"])
    return {"quregs": {"psi": a["quregs"]["psi"], "chi": b["quregs"]["psi"]}, "ops": ops, "dump": ["psi", "chi"]}


def lookahead_programs(logp):
    """what the look-ahead gate log (quest_b200/shim/lookahead.hpp) has to survive: reads and canonical-order restores in
    the middle of a logged window, heap matrices (never logged), measurements, Pauli gadgets, a gate-only stretch several
    windows long, two Quregs sharing the one log"""
    return [relabel_program(logp + 7, 6302), relabel_program(logp + 13, 6303, num_ops=120),
            relabel_program(logp + 13, 6304, num_ops=600, reads=False), cfg2_program(logp + 13, 6403, 40),
            measurement_program(logp + 5, 6102), cfg5_program(logp + 6, 6103, num_terms=30),
            two_quregs_program(logp + 13, 6601, 6602)]
