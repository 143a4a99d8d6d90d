"""Single precision: QuEST's FLOAT_PRECISION=1 build of the backend -- the same kernels compiled with
cplx = float2 (quest_b200/lib/libquest_b200_f32.so) behind the reference's host layers compiled with qreal = float
(libQuEST_f32.so) -- against the UNMODIFIED reference compiled at FLOAT_PRECISION=1 (oracle/_ref_f32/libQuEST.so).
Tolerance: the north star's 1e-5 relative L2 for amplitudes and expectation values (reductions accumulate in double on
the GPU, in float on the CPU reference)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from tests import helpers as H       # noqa: E402
from tests import programs as P      # noqa: E402

TOL32 = 1e-5
REF32 = os.path.join(H.ROOT, "oracle", "_ref_f32", "libQuEST.so")
LIB32 = os.path.join(H.ROOT, "quest_b200", "lib", "libQuEST_f32.so")


def _check32(progs, world=0):
    if not (os.path.exists(REF32) and os.path.exists(LIB32)):
        pytest.skip("fp32 libraries not built (make kernels32 oracle32 quest32)")
    want = H.run_programs("ref32", progs, env={"OMP_NUM_THREADS": str(os.cpu_count() or 8)})
    got = H.run_programs_distributed(progs, world, which="b200dist32") if world else H.run_programs("b20032", progs)
    for k, (g, w) in enumerate(zip(got, want)):
        for name in w["dumps"]:
            assert g["dumps"][name].dtype == np.complex64, "the fp32 build must hold complex<float> amplitudes"
        H.assert_outputs_match(g, w, tol=TOL32, label=f"fp32 prog[{k}]")


def test_fp32_gates_statevector_and_density_matrix():
    _check32([P.gates_program(12, 8001), P.gates_program(16, 8002, num_rounds=1), P.gates_program(6, 8003, dm=1, num_rounds=1),
              P.big_dense_program(12, 8004, 5), P.big_dense_program(13, 8005, 6)])


def test_fp32_calculations_and_channels():
    _check32([P.calcs_program_sv(13, 8101), P.channels_program_dm(6, 8102)])


def test_fp32_baseline_configs_fused_paths():
    """the deferred-queue / tile-engine / coset-kernel paths at sizes that enable them (>= 13 local qubits)"""
    _check32([P.cfg1_program(16, 12345, 120), P.cfg2_program(18, 20302, 100), P.cfg4_program(8, 14014, layers=3),
              P.cfg5_program(15, 28200, num_terms=40), P.relabel_program(15, 8201, num_ops=120)])


def test_fp32_sharded_two_ranks():
    _check32([P.gates_program(7, 8301, max_ctrls=2), P.cfg2_program(15, 8302, 60), P.relabel_program(15, 8303, num_ops=100),
              P.channels_program_dm(4, 8304)], world=2)
