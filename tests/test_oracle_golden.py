"""Pins the numpy oracle (oracle/quest_oracle.py) against the UNMODIFIED reference library.

The committed fixtures tests/golden/*.pkl hold seeded programs and what oracle/_ref/libQuEST.so (built
from /root/reference by oracle/Makefile, generator: tests/golden/make_golden.py) returned for them; the
oracle's API-level interpreter must reproduce every dump and every scalar.  Tolerance: 1e-12 relative L2
(the north-star's fp64 bound) -- the two differ only in floating-point summation order."""
import os

import numpy as np
import pytest

from oracle import quest_oracle as qo
from oracle.quest_oracle_api import run_program
from tests import helpers as H

FIXTURES = ["gates_sv.pkl", "gates_dm.pkl", "calcs_sv.pkl", "channels_dm.pkl", "dense_big.pkl", "relabel_sv.pkl"]


@pytest.mark.parametrize("fname", FIXTURES)
def test_oracle_reproduces_reference(fname):
    fx = H.load_golden(fname)
    for k, (prog, want) in enumerate(zip(fx["programs"], fx["outputs"])):
        got = run_program(prog)
        H.assert_outputs_match(got, want, tol=H.TOL, label=f"{fname}[{k}]")


def _measure_ops(prog):
    return {i for i, op in enumerate(prog["ops"]) if op[0] == "applyQubitMeasurement"}


def test_oracle_configs_small():
    """the BASELINE.json configurations at toy size: cfg1 (H / CNOT / RotateX / CompMatr1), cfg2 (QFT + dense gates), cfg4
    (noisy density-matrix circuit), cfg5 (2nd-order Trotter + Pauli-sum expectation) and the measurement program, whose
    OUTCOMES must be bit-exact: the oracle restates the host mt19937_64 stream too (oracle/quest_rng.py)."""
    fx = H.load_golden("configs_small.pkl")
    for k, (prog, want) in enumerate(zip(fx["programs"], fx["outputs"])):
        got = run_program(prog)
        H.assert_outputs_match(got, want, label=f"cfg[{k}]", int_exact_ops=_measure_ops(prog))


def test_measurement_rng_stream_against_live_reference():
    """std::seed_seq + std::mt19937_64 + uniform_real_distribution restated in oracle/quest_rng.py: many measurement outcomes
    under several seed lists, against the live reference (skipped where oracle/_ref is absent)"""
    if not os.path.exists(H.REF_LIB):
        pytest.skip("oracle/_ref/libQuEST.so not built")
    from tests import programs as P
    progs = [P.measurement_program(9, 11), P.measurement_program(10, 123456789), P.measurement_program(8, 4242)]
    progs[1]["seeds"] = [5, 6, 7, 8, 9]            # seed lists of other lengths exercise seed_seq's mixing
    progs[2]["seeds"] = [4294967295]
    wants = H.run_programs("ref", progs)
    for k, (prog, want) in enumerate(zip(progs, wants)):
        H.assert_outputs_match(run_program(prog), want, label=f"meas[{k}]", int_exact_ops=_measure_ops(prog))


def test_oracle_against_live_reference():
    """when oracle/_ref is present (build container and GPU box), re-run a fresh seed live."""
    if not os.path.exists(H.REF_LIB):
        pytest.skip("oracle/_ref/libQuEST.so not built")
    from tests import programs as P
    progs = [P.gates_program(5, 9001), P.gates_program(3, 9002, dm=1, max_ctrls=1), P.calcs_program_sv(5, 9003),
             P.channels_program_dm(3, 9004), P.big_dense_program(7, 9005, 4), P.cfg1_program(8, 9006, 40), P.cfg2_program(8, 9007, 30),
             P.cfg4_program(4, 9008, layers=2), P.cfg5_program(7, 9009, num_terms=8), P.relabel_program(7, 9010)]
    wants = H.run_programs("ref", progs)
    for k, (prog, want) in enumerate(zip(progs, wants)):
        H.assert_outputs_match(run_program(prog), want, label=f"live[{k}]")


def test_debug_state_formula():
    # tests/utils/qvector.cpp:136-141: a_i = 2i/10 + i(2i+1)/10
    st = qo.new_state(4)
    qo.statevec_initDebugState_sub(st)
    i = np.arange(16, dtype=np.float64)
    assert np.array_equal(st.amps.real, (2 * i) / 10.) and np.array_equal(st.amps.imag, (2 * i + 1) / 10.)


def test_insert_bits_matches_definition():
    rng = np.random.default_rng(5)
    for _ in range(200):
        k = int(rng.integers(0, 6))
        qs = sorted(int(q) for q in rng.choice(20, size=k, replace=False))
        n = int(rng.integers(0, 1 << 12))
        got = int(qo.insertBits(np.int64(n), qs, 0))
        bits = [b for b in range(40)]
        want, src = 0, n
        pos = 0
        for b in bits:
            if b in qs:
                continue
            want |= ((src >> pos) & 1) << b
            pos += 1
        assert got == want


def test_debug_state_bit_exact_vs_reference():
    """initDebugState is pure index arithmetic + two real divisions: the oracle must equal the reference BITWISE"""
    fx = H.load_golden("gates_sv.pkl")
    prog = {"quregs": {"psi": {"n": 6, "init": "debug"}}, "ops": [], "dump": ["psi"]}
    got = run_program(prog)["dumps"]["psi"]
    if os.path.exists(H.REF_LIB):
        want = H.run_programs("ref", [prog])[0]["dumps"]["psi"]
        assert np.array_equal(got, want)
    assert fx["programs"][0]["quregs"]["psi"]["init"] == "debug"
