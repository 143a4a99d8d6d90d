"""Multi-rank parity: the SAME programs on P ranks (one process per rank, state sharded on the top
log2 P qubits) against the single-process reference CPU library.
With n-qubit states over P ranks only n - log2 P qubits are local, so small n makes nearly every gate hit the
prefix paths -- the same trick the reference's CI uses (16 ranks on 6-qubit states).

On a box with >= P GPUs the ranks talk NCCL / NVLink peer memory.  On a box with FEWER GPUs (the driver's 1-GPU test
box) the ranks share devices and the backend's shared-memory / CUDA-IPC transport carries the very same calls
(quest_b200/csrc/qb_comm_shm.cu) -- the reference tests its distributed GPU code the same way
(PERMIT_NODES_TO_SHARE_GPU, CMakeLists.txt:235-240).  Either way nothing here is skipped when a GPU exists."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from tests import helpers as H       # noqa: E402
from tests import programs as P      # noqa: E402

WORLDS = [2, 4, 8] if H.num_gpus() >= 1 else []


def _measure_ops(prog):
    return {i for i, op in enumerate(prog["ops"]) if "Measurement" in op[0] and "Forced" not in op[0]}


def _check(progs, world, env=None):
    if not os.path.exists(H.REF_LIB):
        pytest.skip("oracle/_ref/libQuEST.so not present")
    want = H.run_programs("ref", progs)
    got = H.run_programs_distributed(progs, world, env=env)
    for k, (prog, g, w) in enumerate(zip(progs, got, want)):
        H.assert_outputs_match(g, w, label=f"P={world} prog[{k}]", int_exact_ops=_measure_ops(prog))
    return got


@pytest.mark.skipif(not WORLDS, reason="needs a GPU")
@pytest.mark.parametrize("world", WORLDS)
@pytest.mark.parametrize("p2p", ["1", "0"], ids=["nvlink-p2p", "nccl-only"])
def test_statevector_gates_sharded(world, p2p):
    logp = world.bit_length() - 1
    got = _check([P.gates_program(logp + 4, 6001, max_ctrls=2), P.gates_program(logp + 7, 6002, num_rounds=1),
                  P.cfg1_program(logp + 8, 6003, 120), P.cfg2_program(logp + 6, 6004, 60)], world, env={"QUEST_B200_P2P": p2p})
    # the fused NVLink peer-memory kernels must really have been the path under test (and really off otherwise)
    assert got[0]["p2p_available"] == int(p2p), "NVLink peer-memory path availability is not what the test asked for"
    assert got[0]["transport"] == (1 if world > H.num_gpus() else 0), "unexpected transport"


@pytest.mark.skipif(not WORLDS, reason="needs a GPU")
@pytest.mark.parametrize("world", WORLDS)
@pytest.mark.parametrize("p2p", ["1", "0"], ids=["nvlink-p2p", "nccl-only"])
def test_lazy_qubit_relabelling_sharded(world, p2p):
    """dense gates on rank bits pull the qubit into the shard and leave it there; SWAPs only relabel; everything else
    restores the canonical order first -- all invisible through the API"""
    logp = world.bit_length() - 1
    # the last program is long enough to overflow the backend's 2048-gate queue while gates with rank-bit controls have
    # been dropped on some ranks only (the ranks' queues then differ in length)
    got = _check([P.relabel_program(logp + 3, 6301), P.relabel_program(logp + 7, 6302), P.relabel_program(logp + 13, 6303, num_ops=120),
                  P.relabel_program(logp + 13, 6304, num_ops=2600, reads=False)],
                 world, env={"QUEST_B200_P2P": p2p, "QUEST_B200_OVERLAP_MIN_GATES": "3"})      # overlap even short queues
    if p2p == "1":
        # the exchange overlapped with the deferred gates (qb_p2p_swapHalvesOverlapped) must have been part of what ran
        assert got[-1]["overlapped_swaps"] > 0, "no swap-in overlapped queued gates"


@pytest.mark.skipif(not WORLDS, reason="needs a GPU")
@pytest.mark.parametrize("world", WORLDS)
def test_swap_in_without_overlap_sharded(world):
    """QUEST_B200_OVERLAP=0: every swap-in through the in-place exchange kernel, the queue overtaken but not split"""
    logp = world.bit_length() - 1
    got = _check([P.relabel_program(logp + 13, 6303, num_ops=120), P.cfg2_program(logp + 13, 6403, 40)], world, env={"QUEST_B200_OVERLAP": "0"})
    assert got[-1]["overlapped_swaps"] == 0


@pytest.mark.skipif(not WORLDS, reason="needs a GPU")
@pytest.mark.parametrize("world", WORLDS)
def test_eager_swap_in_and_back_sharded(world):
    """QUEST_B200_RELABEL=0: the reference's own strategy (swap prefix targets in, apply, swap back; SWAPs move
    amplitudes; 1-target dense gates on a rank bit through the fused exchange kernel) stays correct"""
    logp = world.bit_length() - 1
    _check([P.relabel_program(logp + 6, 6401), P.cfg1_program(logp + 8, 6402, 100), P.cfg2_program(logp + 13, 6403, 40)],
           world, env={"QUEST_B200_RELABEL": "0"})


@pytest.mark.skipif(not WORLDS, reason="needs a GPU")
@pytest.mark.parametrize("world", WORLDS)
def test_calcs_measurement_sharded(world):
    logp = world.bit_length() - 1
    _check([P.calcs_program_sv(logp + 5, 6101), P.measurement_program(logp + 5, 6102), P.cfg5_program(logp + 6, 6103, num_terms=30),
            P.big_dense_program(logp + 7, 6104, 4), P.big_dense_program(logp + 8, 6105, 6, nc=0)], world)


@pytest.mark.skipif(not WORLDS, reason="needs a GPU")
@pytest.mark.parametrize("world", WORLDS)
def test_density_matrix_sharded(world):
    logp = world.bit_length() - 1
    n = max(logp + 1, 4)
    _check([P.gates_program(n, 6201, dm=1, num_rounds=1, max_ctrls=1), P.channels_program_dm(n, 6202), P.cfg4_program(n + 1, 6203, layers=2)], world)


@pytest.mark.skipif(not WORLDS, reason="needs a GPU")
@pytest.mark.parametrize("world", WORLDS)
def test_rank_dependent_flushes_keep_ranks_in_step(world):
    """after operations that drain the deferred-gate queue on SOME ranks only (setQuregAmps on the
    owning ranks, controlled SWAP of two rank bits, rank-bit controls) every rank must still choose the same swap-in
    victim -- the choice is made from the shim's rank-independent record, not from the backend's queue"""
    logp = world.bit_length() - 1
    _check([P.rank_divergence_program(logp + 14, logp, 6501), P.rank_divergence_program(logp + 13, logp, 6502, num_rounds=30),
            P.rank_divergence_program(logp + 5, logp, 6503)], world)


@pytest.mark.skipif(not WORLDS, reason="needs a GPU")
@pytest.mark.parametrize("world", WORLDS[:2])
def test_two_relabelled_quregs_and_sync(world):
    """two distributed statevectors with live qubit maps, restored by syncQuESTEnv() in creation order on every rank
    (ADVICE r1, medium), and interleaved gates on several Quregs"""
    logp = world.bit_length() - 1
    _check([P.two_quregs_program(logp + 13, 6601, 6602)], world)


def _exchanges_of(outs, k):
    """exchanges program k issued (the workers report a cumulative counter)"""
    return outs[k]["p2p_exchanges"] - (outs[k - 1]["p2p_exchanges"] if k else 0)


@pytest.mark.skipif(not WORLDS, reason="needs a GPU")
@pytest.mark.parametrize("world", WORLDS[1:2])
def test_look_ahead_victim_choice_sharded(world):
    """QUEST_B200_LOOKAHEAD=W (opt-in): gate calls are logged and replayed W gates late so that a swap-in can evict the
    qubit needed farthest in the future (quest_b200/shim/lookahead.hpp).  Same results as the reference through reads,
    restores, heap matrices, measurements and two Quregs sharing the log -- and no more exchanges than the default rule
    on the gate-only stretch (4 ranks, window 8: 86 against 107, profiles/r2_lookahead_gpu_check.txt)"""
    logp = world.bit_length() - 1
    progs = P.lookahead_programs(logp)
    got = _check(progs, world, env={"QUEST_B200_LOOKAHEAD": "8"})
    base = H.run_programs_distributed(progs, world)
    assert _exchanges_of(got, 2) <= _exchanges_of(base, 2), "look-ahead cost exchanges on a gate-only program"
