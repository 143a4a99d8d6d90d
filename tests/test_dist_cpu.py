"""CPU coverage of the N>1 path: world_size-2 and -4 `gloo` process groups run the distributed restatement of the
sharding logic (oracle/quest_oracle_dist.py -- the CPU model of quest_b200/shim/localiser_b200.cpp: pair ranks,
packing, buffer masks, prefix sign rules) and the re-assembled global state must equal the single-process oracle.
With n-qubit states over P ranks only n - log2 P qubits are local, so nearly every gate takes a prefix path."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

torch = pytest.importorskip("torch")
import torch.distributed as dist             # noqa: E402
import torch.multiprocessing as mp            # noqa: E402

from oracle import quest_oracle as qo         # noqa: E402
from oracle import quest_oracle_dist as qd    # noqa: E402
from tests.programs import rand_unitary, rand_state, rand_density   # noqa: E402


def _pick(rng, n, k):
    return [int(q) for q in rng.choice(n, size=k, replace=False)]


def _sv_ops(n, seed, count):
    rng = np.random.default_rng(seed)
    ops = []
    for _ in range(count):
        kind = rng.choice(["dense1", "swap", "dense2", "dense3", "diag1", "pauli", "gadget", "phase"])
        nc = int(rng.integers(0, 3))
        if kind == "dense1":
            q = _pick(rng, n, nc + 1)
            ops.append(("dense1", q[:nc], [int(b) for b in rng.integers(0, 2, nc)], q[nc], rand_unitary(rng, 2)))
        elif kind == "swap":
            q = _pick(rng, n, nc + 2)
            ops.append(("swap", q[:nc], [int(b) for b in rng.integers(0, 2, nc)], q[nc], q[nc + 1]))
        elif kind in ("dense2", "dense3"):
            k = 2 if kind == "dense2" else 3
            if nc + k > n:
                continue
            q = _pick(rng, n, nc + k)
            ops.append(("denseK", q[:nc], [int(b) for b in rng.integers(0, 2, nc)], q[nc:], rand_unitary(rng, 1 << k)))
        elif kind == "diag1":
            q = _pick(rng, n, nc + 1)
            ops.append(("diag1", q[:nc], [int(b) for b in rng.integers(0, 2, nc)], q[nc], np.exp(1j * rng.uniform(0, 6, 2))))
        elif kind in ("pauli", "gadget"):
            k = int(rng.integers(1, min(4, n - nc) + 1))
            q = _pick(rng, n, nc + k)
            chars = rng.choice(list("XYZ"), size=k)
            if not any(c in "XY" for c in chars):
                chars[0] = "X"
            t = q[nc:]
            x = [a for c, a in zip(chars, t) if c == "X"]; y = [a for c, a in zip(chars, t) if c == "Y"]; z = [a for c, a in zip(chars, t) if c == "Z"]
            th = float(rng.uniform(0, 6))
            facs = (0j, 1 + 0j) if kind == "pauli" else (complex(np.cos(th)), 1j * np.sin(th))
            ops.append(("pauli", q[:nc], [int(b) for b in rng.integers(0, 2, nc)], x, y, z, facs[0], facs[1]))
        else:
            k = int(rng.integers(1, min(4, n - nc) + 1))
            q = _pick(rng, n, nc + k)
            ops.append(("phase", q[:nc], [int(b) for b in rng.integers(0, 2, nc)], q[nc:], float(rng.uniform(0, 6))))
    return ops


def _apply_global(st, op):
    k = op[0]
    if k == "dense1": qo.statevec_anyCtrlOneTargDenseMatr_subA(st, op[1], op[2], op[3], op[4])
    elif k == "swap": qo.statevec_anyCtrlSwap_subA(st, op[1], op[2], min(op[3], op[4]), max(op[3], op[4]))
    elif k == "denseK": qo.statevec_anyCtrlAnyTargDenseMatr_sub(st, op[1], op[2], op[3], op[4], False)
    elif k == "diag1": qo.statevec_anyCtrlOneTargDiagMatr_sub(st, op[1], op[2], op[3], op[4])
    elif k == "pauli": qo.statevector_anyCtrlPauliTensorOrGadget_subA(st, op[1], op[2], op[3], op[4], op[5], op[6], op[7])
    elif k == "phase": qo.statevector_anyCtrlAnyTargZOrPhaseGadget_sub(st, op[1], op[2], op[3], np.exp(1j * op[4]), np.exp(-1j * op[4]))


def _apply_dist(sh, op):
    k = op[0]
    if k == "dense1": qd.dense1(sh, op[1], op[2], op[3], op[4])
    elif k == "swap": qd.swap(sh, op[1], op[2], op[3], op[4])
    elif k == "denseK": qd.denseK(sh, op[1], op[2], op[3], op[4])
    elif k == "diag1": qd.diag1(sh, op[1], op[2], op[3], op[4])
    elif k == "pauli": qd.pauli(sh, op[1], op[2], op[3], op[4], op[5], op[6], op[7])
    elif k == "phase": qd.phase_gadget(sh, op[1], op[2], op[3], op[4])


def _worker(rank, world, port, n, seed, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        logp = world.bit_length() - 1
        rng = np.random.default_rng(seed)
        psi = rand_state(rng, n)
        N = (1 << n) // world
        st = qo.State(psi[rank * N:(rank + 1) * N].copy(), n, 0, rank, logp, np.zeros(N, dtype=np.complex128))
        sh = qd.Shard(st, rank, world)
        glob = qo.State(psi.copy(), n)
        ops = _sv_ops(n, seed + 1, 60)
        for op in ops:
            _apply_dist(sh, op)
            _apply_global(glob, op)
        # reductions
        qs = _pick(rng, n, 2)
        p_d = qd.prob_of_outcome(sh, qs, [1, 0]); p_g = qo.statevec_calcProbOfMultiQubitOutcome_sub(glob, qs, [1, 0])
        t_d = qd.total_prob(sh); t_g = qo.statevec_calcTotalProb_sub(glob)
        exps = []
        for _ in range(6):
            k = int(rng.integers(1, 4)); t = _pick(rng, n, k); chars = rng.choice(list("XYZ"), size=k)
            x = [a for c, a in zip(chars, t) if c == "X"]; y = [a for c, a in zip(chars, t) if c == "Y"]; z = [a for c, a in zip(chars, t) if c == "Z"]
            want = qo.statevec_calcExpecPauliStr_subA(glob, x, y, z) if (x or y) else qo.statevec_calcExpecAnyTargZ_sub(glob, z)
            exps.append(abs(qd.expec_pauli(sh, x, y, z) - want))
        # inner product with a second sharded register, then projectors on suffix and on rank-bit qubits
        phi = rand_state(np.random.default_rng(seed + 3), n)
        shB = qd.Shard(qo.State(phi[rank * N:(rank + 1) * N].copy(), n, 0, rank, logp, np.zeros(N, dtype=np.complex128)), rank, world)
        ip_err = abs(qd.inner_product(sh, shB) - qo.statevec_calcInnerProduct_sub(glob, qo.State(phi.copy(), n)))
        for qs in ([0], [n - 1], [1, n - 1]):
            outs = [1] * len(qs)
            pr = qo.statevec_calcProbOfMultiQubitOutcome_sub(glob, qs, outs)
            qd.projector(sh, qs, outs, pr); qo.statevec_multiQubitProjector_sub(glob, qs, outs, pr)
        gathered = [torch.empty(2 * N, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(st.amps.view(np.float64).copy()))
        full = np.concatenate([g.numpy().view(np.complex128) for g in gathered])
        err = float(np.linalg.norm(full - glob.amps) / np.linalg.norm(glob.amps)) + ip_err

        # density matrix channels with prefix bra qubits
        m = max(logp + 1, 3)
        rho = rand_density(np.random.default_rng(seed + 7), m)
        Nd = rho.size // world
        dst = qo.State(rho[rank * Nd:(rank + 1) * Nd].copy(), m, 1, rank, logp, np.zeros(Nd, dtype=np.complex128))
        dsh = qd.Shard(dst, rank, world)
        dglob = qo.State(rho.copy(), m, 1)
        for ket in range(m):
            qd.depolarising(dsh, ket, 0.3); qo.densmatr_oneQubitDepolarising_subA(dglob, ket, 0.3)
            qd.damping(dsh, ket, 0.2); qo.densmatr_oneQubitDamping_subA(dglob, ket, 0.2)
            qd.pauli_channel(dsh, ket, 0.1, 0.05, 0.2); qo.densmatr_oneQubitPauliChannel_subA(dglob, ket, 0.65, 0.1, 0.05, 0.2)
            qd.dephasing(dsh, ket, 0.15); qo.densmatr_oneQubitDephasing_subA(dglob, ket, 0.15)
        qd.two_qubit_dephasing(dsh, 0, m - 1, 0.1); qo.densmatr_twoQubitDephasing_subA(dglob, 0, m - 1, 0.1)
        pr = qo.densmatr_calcProbOfMultiQubitOutcome_sub(dglob, [m - 1], [0])
        qd.densmatr_projector(dsh, [m - 1], [0], pr); qo.densmatr_multiQubitProjector_sub(dglob, [m - 1], [0], pr)
        gathered = [torch.empty(2 * Nd, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(dst.amps.view(np.float64).copy()))
        dfull = np.concatenate([g.numpy().view(np.complex128) for g in gathered])
        derr = float(np.linalg.norm(dfull - dglob.amps) / np.linalg.norm(dglob.amps))
        if rank == 0:
            q.put({"err": err, "derr": derr, "prob": abs(p_d - p_g), "tot": abs(t_d - t_g), "exp": max(exps)})
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 5), (2, 8), (4, 6)])
def test_sharded_oracle_matches_single_process(world, n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29800 + world * 10 + n
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, 1000 + n, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0, "a distributed worker failed"
    res = q.get(timeout=10)
    assert res["err"] <= 1e-12, res
    assert res["derr"] <= 1e-12, res
    assert res["prob"] <= 1e-12 and res["tot"] <= 1e-12 and res["exp"] <= 1e-12, res


def _relabel_worker(rank, world, port, n, seed, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        logp = world.bit_length() - 1
        rng = np.random.default_rng(seed)
        psi = rand_state(rng, n)
        N = (1 << n) // world
        st = qo.State(psi[rank * N:(rank + 1) * N].copy(), n, 0, rank, logp, np.zeros(N, dtype=np.complex128))
        rs = qd.RelabelledShard(qd.Shard(st, rank, world), n)
        glob = qo.State(psi.copy(), n)
        probs = []
        for op in _sv_ops(n, seed + 1, 80):
            k = op[0]
            if k == "dense1": rs.dense(op[1], op[2], [op[3]], op[4])
            elif k == "swap": rs.swap(op[1], op[2], op[3], op[4])
            elif k == "denseK": rs.dense(op[1], op[2], list(op[3]), op[4])
            elif k == "diag1": rs.diag1(op[1], op[2], op[3], op[4])
            elif k == "pauli": rs.pauli(op[1], op[2], op[3], op[4], op[5], op[6], op[7])
            elif k == "phase": rs.phase_gadget(op[1], op[2], op[3], op[4])
            _apply_global(glob, op)
            if rng.integers(6) == 0:        # a relabelling-aware read in the middle of the permuted state
                qs = _pick(rng, n, 2)
                probs.append(abs(rs.prob_of_outcome(qs, [1, 0]) - qo.statevec_calcProbOfMultiQubitOutcome_sub(glob, qs, [1, 0])))
        permuted = rs.phys != list(range(n))
        exchanges_before_restore = rs.exchanges
        rs.canonicalise()
        gathered = [torch.empty(2 * N, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(st.amps.view(np.float64).copy()))
        full = np.concatenate([g.numpy().view(np.complex128) for g in gathered])
        err = float(np.linalg.norm(full - glob.amps) / np.linalg.norm(glob.amps))
        if rank == 0:
            q.put({"err": err, "prob": max(probs) if probs else 0.0, "permuted": permuted, "exchanges": exchanges_before_restore})
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 6), (4, 7)])
def test_lazy_relabelling_model_matches_single_process(world, n):
    """the lazy qubit relabelling of the sharding shim (swap = permutation edit, prefix targets pulled into the shard and
    left there, relabelling-aware reads, canonical order restored at the end), restated in oracle/quest_oracle_dist.py and
    run on gloo ranks, against the single-process oracle"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + world * 10 + n
    procs = [ctx.Process(target=_relabel_worker, args=(r, world, port, n, 2000 + n, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0, "a distributed worker failed"
    res = q.get(timeout=10)
    assert res["err"] <= 1e-12 and res["prob"] <= 1e-12, res
    assert res["permuted"] and res["exchanges"] > 0, "the test circuit never exercised the relabelling"
