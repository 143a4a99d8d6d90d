"""End-to-end parity through the drop-in boundary: the SAME programs (QuEST public API calls) run on
quest_b200/lib/libQuEST.so (reference host layers + our sm_100a backend, GPU-accelerated Quregs only) and are
compared with (a) the committed golden outputs of the unmodified reference CPU library and (b) the live
reference library oracle/_ref/libQuEST.so at larger sizes.  fp64 tolerance 1e-12 relative L2; measurement
outcomes bit-exact.  The BASELINE configurations at their FULL sizes are compared with the reference in
tests/test_fullsize_gpu.py."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from tests import helpers as H       # noqa: E402
from tests import programs as P      # noqa: E402

FIXTURES = ["gates_sv.pkl", "gates_dm.pkl", "calcs_sv.pkl", "channels_dm.pkl", "dense_big.pkl", "configs_small.pkl"]


def _measure_ops(prog):
    return {i for i, op in enumerate(prog["ops"]) if "Measurement" in op[0] and "Forced" not in op[0]}


@pytest.mark.parametrize("fname", FIXTURES)
def test_backend_reproduces_golden(fname):
    fx = H.load_golden(fname)
    outs = H.run_programs("b200", fx["programs"])
    for k, (prog, got, want) in enumerate(zip(fx["programs"], outs, fx["outputs"])):
        H.assert_outputs_match(got, want, label=f"{fname}[{k}]", int_exact_ops=_measure_ops(prog))


def _live(progs):
    if not os.path.exists(H.REF_LIB):
        pytest.skip("oracle/_ref/libQuEST.so not present")
    want = H.run_programs("ref", progs, env={"OMP_NUM_THREADS": str(os.cpu_count() or 8)})
    got = H.run_programs("b200", progs)
    for k, (prog, g, w) in enumerate(zip(progs, got, want)):
        H.assert_outputs_match(g, w, label=f"live[{k}]", int_exact_ops=_measure_ops(prog))


def test_live_gates_medium():
    _live([P.gates_program(12, 4001), P.gates_program(16, 4002, num_rounds=1), P.gates_program(7, 4003, dm=1, num_rounds=1),
           P.big_dense_program(13, 4004, 6), P.big_dense_program(12, 4005, 5)])


def test_live_calcs_and_channels():
    _live([P.calcs_program_sv(14, 4101), P.channels_program_dm(7, 4102), P.measurement_program(12, 4103)])


def test_live_lazy_qubit_relabelling():
    """SWAP relabelling + restoring the canonical order where an operation needs it (single GPU: no rank bits)"""
    _live([P.relabel_program(6, 4201), P.relabel_program(14, 4202), P.relabel_program(17, 4203, num_ops=120),
           P.relabel_program(15, 4204, num_ops=2600, reads=False)])


def test_live_cfg1_20q():
    """BASELINE cfg 1 exactly: 20 qubits, H layer + 200 random {H, CNOT, RotateX, CompMatr1}"""
    _live([P.cfg1_program(20, 12345, 200)])


def test_live_cfg2_qft_dense():
    _live([P.cfg2_program(20, 20302, 100), P.cfg2_program(12, 20302, 200)])


def test_live_cfg4_noisy_dm():
    _live([P.cfg4_program(9, 14014, layers=3), P.cfg4_program(6, 14014, layers=10)])


def test_live_cfg5_trotter_paulisum():
    _live([P.cfg5_program(16, 28200, num_terms=40), P.cfg5_program(10, 28200, num_terms=200)])


def test_qft_known_answer_26q():
    """size-independent property at a size the CPU does not need to replay: QFT|0..0> is uniform 2^-n/2,
    the QFT of a basis state has |amp| = 2^-n/2 everywhere, total probability stays 1."""
    n = 26
    prog = {"quregs": {"psi": {"n": n, "init": ["classical", 12345]}},
            "ops": [["applyFullQuantumFourierTransform", "psi"], ["calcTotalProb", "psi"],
                    ["calcProbOfQubitOutcome", "psi", n - 1, 0]], "dump": ["psi"]}
    out = H.run_programs("b200", [prog])[0]
    amps = out["dumps"]["psi"]
    assert np.max(np.abs(np.abs(amps) - 2.0 ** (-n / 2))) < 1e-13
    # DFT convention of the reference tests (tests/utils/linalg.cpp:160-176): out[x] = 2^-n/2 exp(+2 pi i x y / 2^n)
    x = np.arange(0, 1 << n, 65537)
    want = 2.0 ** (-n / 2) * np.exp(2j * np.pi * ((x * 12345) % (1 << n)) / (1 << n))
    assert np.max(np.abs(amps[x] - want)) < 1e-12
    assert abs(out["results"][1] - 1) < 1e-12 and abs(out["results"][2] - 0.5) < 1e-12


def test_live_interleaved_quregs_coevolution():
    """the reference's psi / rho co-evolution pattern (tests/integration/densitymatrix.cpp:68-163): identical gates applied
    alternately to a statevector, a density matrix and a third register; the backend keeps one deferred queue per state"""
    _live([P.coevolution_program(9, 4301), P.coevolution_program(13, 4302, num_ops=40)])
