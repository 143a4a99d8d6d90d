#!/usr/bin/env python
"""bench.py -- the reference's headline workload on the B200 backend (BASELINE.json cfg 2), 1..8 GPUs of one box.

workload: a (30 + log2 N)-qubit fp64 statevector, 2^30 amplitudes per GPU (weak scaling; N=1 is BASELINE's 30-qubit
configuration).  One STEP = applyFullQuantumFourierTransform (30+ H, the controlled-phase ladders, n/2 SWAPs;
api/operations.cpp:1934-1953) + 200 random dense 1- and 2-qubit unitaries on uniformly random targets (seed 20302,
SURVEY.md 8d).  Every arm and every N drives QuEST's PUBLIC API (the call a user makes): the product
arm on the drop-in quest_b200/lib/libQuEST.so (sharding shim + deferred gate queue + sm_100a kernels all on the measured
path), the reference arm on the unmodified CPU/OpenMP build oracle/_ref/libQuEST.so.

  metric  "30q fp64 gates/s" -- the unit is ONE gate applied to ONE 2^30-amplitude shard; at N GPUs every circuit gate
          acts on N shards, so value = N x gates per step / step time.  The same string at every N and in both arms.
  value   state resident in HBM, CUDA events on the backend's stream, max over ranks.  The backend defers work (fusable
          gates are queued; uncontrolled SWAPs -- the QFT's last layer -- only relabel qubits until an operation needs
          the canonical order), so the timed region is K x (step + flush of the queue) and ENDS with syncQuESTEnv(),
          which restores the canonical qubit order and drains the stream: nothing is left undone when the clock stops.
          The restore's cost on its own is reported (roofline.sections.restore_ms), as is the gate count without the
          relabelled SWAPs (config.gates_per_step_without_relabelled_swaps).
  e2e     the same K steps end to end with HOST buffers: every gate operand travels host -> device inside its API
          call and every step ends with calcProbOfQubitOutcome read back device -> host (which forces the step to
          execute); the final syncQuESTEnv() is inside the timed region too; host wall clock, max over ranks.
          (A statevector never crosses PCIe in QuEST -- createQureg/initZeroState build it on the device,
          api/qureg.cpp:143-174 -- so the per-step host inputs are the gate matrices and angles.)
  roofline  dominant kernel k_tile_pass (the fused multi-gate pass): achieved = algorithmic bytes of the step's gates
          (2*16*2^30 / 2^controls each) / step time, against MEASURED_PEAKS.json hbm_gbs.  frac > 1 is the
          gain of applying several gates per pass over HBM; physical_frac is the traffic actually moved (one read + one
          write of the shard per launch; `traffic` = ncu's dram bytes per launch, profiles/); fp64_frac is the share of the
          FP64 pipe's measured peak the executed fused multiply-adds amount to -- the pass is FP64/shared-memory bound,
          not HBM bound, once it fuses more than ~3 gates.  N>1 adds `nvlink` (bytes each GPU sent per direction / time
          spent exchanging, against 900 GB/s).
  cpu_baseline / --impl reference   the UNMODIFIED reference on the box's host cores, all threads, on a bounded sample of
          the same gate stream (every k-th gate of the 680) at 30 qubits = the metric's unit.
  secondary  the other BASELINE configurations measured in the same run (cfg 3: random circuit at 2^31 amplitudes per GPU,
          i.e. 34 qubits on 8 GPUs; cfg 4: 14-qubit noisy density matrix; cfg 5: 28-qubit Trotter + 200-term Pauli sum).
"""
import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20302
NUM_DENSE = 200
AMP_BYTES = 16
METRIC = "30q fp64 gates/s"
FP64_PEAK_TFLOPS = 36.9          # measured on this pool's B200s with tools/dfma_probe.cu (profiles/r1_fp64_probe.txt)
NVLINK_PEAK_GBS = 900.0          # per direction per GPU (BASELINE.md / north star)


# ------------------------------------------------------------------------------------------------
# the gate streams (identical for every arm; tests/test_fullsize_gpu.py replays them against the reference)
# ------------------------------------------------------------------------------------------------
def rand_unitary(rng, dim):
    z = rng.normal(size=(dim, dim)) + 1j * rng.normal(size=(dim, dim))
    q, r = np.linalg.qr(z)
    d = np.diag(r)
    return q * (d / np.abs(d))


def qft_stream(n):
    """applyFullQuantumFourierTransform expanded exactly as api/operations.cpp:1934-1953 issues it"""
    ops = []
    for t in range(n - 1, -1, -1):
        ops.append(("h", t))
        for m in range(t):
            ops.append(("cphase", t, t - m - 1, math.pi / (1 << (m + 1))))   # target t, control t-m-1
    for t in range(n // 2):
        ops.append(("swap", t, n - 1 - t))
    return ops


def dense_stream(n, seed=SEED, num=NUM_DENSE):
    rng = np.random.default_rng(seed)
    ops = []
    for _ in range(num):
        if rng.integers(2):
            ops.append(("m1", int(rng.integers(n)), rand_unitary(rng, 2)))
        else:
            a, b = (int(q) for q in rng.choice(n, size=2, replace=False))
            ops.append(("m2", a, b, rand_unitary(rng, 4)))
    return ops


def cfg3_stream(n, seed=34008, num=100):
    """BASELINE cfg 3: gates drawn uniformly from {H, Rx, CompMatr1, CNOT, CompMatr2}, targets uniform
    over ALL n qubits, so ~log2(P)/n of the targets sit on rank bits"""
    rng = np.random.default_rng(seed)
    ops = []
    for _ in range(num):
        k = int(rng.integers(5))
        if k == 0:
            ops.append(("h", int(rng.integers(n))))
        elif k == 1:
            ops.append(("rx", int(rng.integers(n)), float(rng.uniform(0, 2 * math.pi))))
        elif k == 2:
            ops.append(("m1", int(rng.integers(n)), rand_unitary(rng, 2)))
        elif k == 3:
            c, t = (int(q) for q in rng.choice(n, size=2, replace=False))
            ops.append(("cnot", c, t))
        else:
            a, b = (int(q) for q in rng.choice(n, size=2, replace=False))
            ops.append(("m2", a, b, rand_unitary(rng, 4)))
    return ops


def algorithmic_bytes(op, local_amps):
    """a gate on N local amps with c controls reads+writes 2*B*N/2^c; SWAP counts as c=1."""
    full = 2 * AMP_BYTES * local_amps
    return full // 2 if op[0] in ("cphase", "swap", "cnot") else full


def build_config(n_local, world, workload="cfg2", cpu_sample=24):
    """the `config` object -- a pure function of the workload, so both arms print the same one"""
    n = n_local + int(math.log2(max(1, world)))
    if workload == "cfg3":
        return {"workload": f"cfg3: {n}q fp64 statevector, initPlusState + 100 random gates from {{H, Rx, CompMatr1, CNOT, CompMatr2}}, targets uniform over all qubits",
                "qubits": n, "amps_per_gpu": 1 << n_local, "gates_per_step": 100, "seed": 34008,
                "l2_policy": "state per GPU is far larger than L2; every pass streams it from HBM"}
    gates = len(qft_stream(n)) + NUM_DENSE
    return {"workload": f"cfg2: {n}q fp64 statevector ({1 << int(math.log2(max(1, world)))} x 2^{n_local} amplitudes), applyFullQuantumFourierTransform + {NUM_DENSE} random dense 1/2-qubit gates",
            "qubits": n, "amps_per_gpu": 1 << n_local, "gates_per_step": gates, "seed": SEED,
            "gates_per_step_without_relabelled_swaps": gates - n // 2,
            "unit": f"one gate applied to one 2^{n_local}-amplitude shard (a circuit gate counts once per GPU)",
            "l2_policy": "state (16 GiB per GPU) is far larger than L2; every pass streams it from HBM",
            "reference_arm": f"unmodified CPU/OpenMP reference at {n_local} qubits (the metric's unit) on every {max(1, (len(qft_stream(n_local)) + NUM_DENSE) // cpu_sample)}th gate "
                             f"of the {len(qft_stream(n_local)) + NUM_DENSE}-gate stream ({cpu_sample} gates per step), all host threads"}


# ------------------------------------------------------------------------------------------------
# clocks sampling
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.proc, self.lines, self.index = None, [], index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0])); smax.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the unmodified CPU library on a bounded stratified sample
# ------------------------------------------------------------------------------------------------
_REF_WORKER = r"""
import json, os, sys, time
sys.path.insert(0, {root!r})
import numpy as np
import bench
from quest_b200 import quest_api as qa
n, steps, warmup, sample_size, budget = {n}, {steps}, {warmup}, {sample}, {budget}
Q = qa.QuEST(os.path.join({root!r}, "oracle", "_ref", "libQuEST.so"))      # the unmodified reference CPU build
Q.initCustomQuESTEnv(0, 0, 1)
q = Q.createCustomQureg(n, 0, 0, 0, 1)
Q.initZeroState(q)
stream = bench.qft_stream(n) + bench.dense_stream(n)
stride = max(1, len(stream) // sample_size)
sample = stream[::stride][:sample_size]
def apply(op):
    if op[0] == "h": Q.applyHadamard(q, op[1])
    elif op[0] == "cphase": Q.applyTwoQubitPhaseShift(q, op[1], op[2], op[3])
    elif op[0] == "swap": Q.applySwap(q, op[1], op[2])
    elif op[0] == "m1": Q.applyCompMatr1(q, op[1], Q.getCompMatr1(op[2]))
    else: Q.applyCompMatr2(q, op[1], op[2], Q.getCompMatr2(op[3]))
apply(sample[0])                                    # touches every page once
times, gates = [], []
for s in range(warmup + steps):
    t0 = time.perf_counter(); done = 0
    for op in sample:
        apply(op); done += 1
        if time.perf_counter() - t0 > budget: break
    dt = time.perf_counter() - t0
    if s >= warmup: times.append(dt); gates.append(done)
prob = Q.calcTotalProb(q)
kinds = dict()
for op in sample: kinds[op[0]] = kinds.get(op[0], 0) + 1
print("REFJSON " + json.dumps(dict(times=times, gates=gates, prob=prob, stride=stride, kinds=kinds, threads=int(os.environ.get("OMP_NUM_THREADS", "1")))))
"""


def run_reference(n, steps, warmup, sample_size, budget_s):
    cores = os.cpu_count() or 1
    env = dict(os.environ, OMP_NUM_THREADS=str(cores), OMP_PROC_BIND="spread", OMP_PLACES="cores")
    code = _REF_WORKER.format(root=ROOT, n=n, steps=steps, warmup=warmup, sample=sample_size, budget=budget_s)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=3600)
    for line in r.stdout.splitlines():
        if line.startswith("REFJSON "):
            d = json.loads(line[8:])
            total_t, total_g = sum(d["times"]), sum(d["gates"])
            return {"value": total_g / total_t, "unit": "gates/s", "cores": cores, "kind": "reference",
                    "sample": f"{d['gates'][0]} gates/step: every {d['stride']}th gate of the {len(qft_stream(n)) + NUM_DENSE}-gate cfg-2 stream at {n} qubits "
                              f"({d['kinds']}), {len(d['times'])} timed step(s), OMP_NUM_THREADS={cores}",
                    "ms_per_step": 1e3 * total_t / len(d["times"]), "total_prob": d["prob"]}
    raise RuntimeError(f"reference worker failed rc={r.returncode}\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")


# ------------------------------------------------------------------------------------------------
# the product arm
# ------------------------------------------------------------------------------------------------
class Product:
    """the drop-in library driven through QuEST's public API, on 1..8 ranks"""

    def __init__(self, rank, world, local_rank):
        import torch
        from quest_b200 import capi, quest_api as qa
        self.torch, self.capi, self.rank, self.world = torch, capi, rank, world
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the quest_b200 backend has no CPU fallback")
        torch.cuda.set_device(local_rank)
        self.dev = torch.device("cuda", local_rank)
        self.dist = None
        if world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist
        capi.call("qb_bind_device", local_rank)
        if world > 1:
            # the backend's own NCCL communicator: its unique id travels over torch.distributed (control plane only)
            idbuf = (C.c_char * 128)()
            if rank == 0:
                capi.call("qb_comm_get_unique_id", idbuf)
            t = torch.tensor(list(bytes(idbuf)), dtype=torch.uint8, device=self.dev)
            self.dist.broadcast(t, src=0)
            os.environ["QUEST_B200_NCCL_ID"] = bytes(t.cpu().tolist()).hex()
        self.Q = qa.QuEST(qa.B200_LIB)
        self.Q.initCustomQuESTEnv(1 if world > 1 else 0, 1, 0)
        self.lib = capi.lib()

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist:
            self.dist.barrier()

    def max_over_ranks(self, x):
        if not self.dist:
            return float(x)
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.item()

    def create(self, n, dm=0):
        q = self.Q.createCustomQureg(n, dm, 1 if self.world > 1 else 0, 1, 0)
        assert q.isGpuAccelerated == 1 and q.numNodes == self.world
        return q

    def stats(self):
        s = (C.c_double * 8)()
        self.lib.qb_tile_stats(s)
        a, b = C.c_ulonglong(), C.c_ulonglong()
        self.lib.qb_p2p_stats(C.byref(a), C.byref(b))
        return {"launches": self.lib.qb_launch_count(), "passes": s[0], "rounds": s[1], "tile_ops": s[2], "direct_ops": s[3],
                "fma": s[5], "pass_bytes": s[6], "exchanges": a.value, "link_bytes": b.value}

    def gate(self, q, op, m=None):
        Q = self.Q
        k = op[0]
        if k == "m1": Q.applyCompMatr1(q, op[1], m)
        elif k == "m2": Q.applyCompMatr2(q, op[1], op[2], m)
        elif k == "h": Q.applyHadamard(q, op[1])
        elif k == "rx": Q.applyRotateX(q, op[1], op[2])
        elif k == "cnot": Q.applyControlledPauliX(q, op[1], op[2])
        elif k == "cphase": Q.applyTwoQubitPhaseShift(q, op[1], op[2], op[3])
        elif k == "swap": Q.applySwap(q, op[1], op[2])
        else: raise ValueError(k)

    def mats(self, ops):
        Q = self.Q
        return [(op, Q.getCompMatr1(op[2]) if op[0] == "m1" else (Q.getCompMatr2(op[3]) if op[0] == "m2" else None)) for op in ops]

    def timed(self, fn, steps, warmup):
        """warm up, then time `steps` calls of fn with CUDA events on the backend's stream (the legacy default stream,
        which is torch's current stream), barrier + synchronize on both sides, max over ranks -> ms per call"""
        torch = self.torch
        for _ in range(warmup):
            fn()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1)) / steps


def secondary_configs(P, n_local, full):
    torch = P.torch
    """the other BASELINE configurations, through the public API, device-timed (events) incl. the final syncQuESTEnv"""
    from quest_b200.program import _Interp
    from tests import programs as TP
    Q, world = P.Q, P.world
    logw = int(math.log2(world))
    out = {}

    # cfg 3: random circuit, 2^31 amplitudes per GPU (34 qubits on 8 GPUs); on 4 GPUs additionally 34 qubits (2^32 per GPU)
    sizes = [31 + logw] + ([34] if world == 4 and full else [])
    for n3 in sizes:
        try:
            q = P.create(n3)
        except Exception as exc:      # not enough memory on this box: report, never fake
            out[f"cfg3_{n3}q"] = {"error": str(exc)[:200]}
            continue
        ops = P.mats(cfg3_stream(n3))
        reps = 2
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(reps)]

        def circuit(ev=None):
            if ev: ev[0].record()
            Q.initPlusState(q)
            if ev: ev[1].record()
            for op, m in ops:
                P.gate(q, op, m)
            P.capi.call("qb_flush")
            if ev: ev[2].record()
            Q.syncQuESTEnv()                    # restores the canonical qubit order (exchanges, when rank bits were relabelled)
            if ev: ev[3].record()
        circuit()
        P.barrier()
        s0 = P.stats()
        for r in range(reps):
            circuit(evs[r])
        P.barrier()
        s1 = P.stats()
        ms = P.max_over_ranks(sum(e[0].elapsed_time(e[3]) for e in evs) / reps)
        ms_gates = P.max_over_ranks(sum(e[1].elapsed_time(e[2]) for e in evs) / reps)
        ms_restore = P.max_over_ranks(sum(e[2].elapsed_time(e[3]) for e in evs) / reps)
        prob = Q.calcTotalProb(q)
        Q.destroyQureg(q)
        out[f"cfg3_{n3}q"] = {"workload": f"cfg3: {n3}q random circuit (100 gates from {{H,Rx,CompMatr1,CNOT,CompMatr2}}, seed 34008) incl. initPlusState + syncQuESTEnv",
                              "n_gpus": world, "ms_per_circuit": ms, "gates_ms": ms_gates, "restore_canonical_order_ms": ms_restore,
                              "circuit_gates_per_s": 100 / (ms * 1e-3), "exchanges_per_circuit": (s1["exchanges"] - s0["exchanges"]) / reps,
                              "link_bytes_per_dir_per_circuit": (s1["link_bytes"] - s0["link_bytes"]) / reps, "launches_per_circuit": (s1["launches"] - s0["launches"]) / reps,
                              "total_prob": prob}

    def run_prog(prog, reps):
        """time a tests/programs.py program (creation and destruction of its Quregs excluded, everything else included)"""
        for spec in prog["quregs"].values():
            spec["custom"] = [1 if world > 1 else 0, 1, 0]
        it = _Interp(Q, dict(prog, dump=[]))
        # split: create quregs once, then time the op list
        quregs = {}
        for name, spec in prog["quregs"].items():
            quregs[name] = Q.createCustomQureg(int(spec["n"]), int(spec.get("dm", 0)), *spec["custom"])
        it.quregs = quregs
        results = []

        # operands (heap matrices, Kraus maps, Pauli sums) are created ONCE, outside the timed region: a user builds a
        # channel or a Hamiltonian once and applies it many times; creating and freeing device objects per call would time
        # cudaMalloc/cudaFree (which synchronise the device), not the operators
        calls = []
        for op in prog["ops"]:
            it.keep, it.outarr = [], None
            cargs = [it.conv(a) for a in op[1:]]
            calls.append((getattr(Q.lib, op[0]), cargs, it.keep))
        made = list(it.cleanup); it.cleanup.clear()

        def body():
            for name, spec in prog["quregs"].items():
                Q.initPlusState(quregs[name])
            results.clear()
            for fn, cargs, _keep in calls:
                results.append(fn(*cargs))
            Q.syncQuESTEnv()
        ms = P.timed(body, reps, 1)
        for fn, obj in made:
            getattr(Q.lib, fn)(obj)
        for qq in quregs.values():
            Q.destroyQureg(qq)
        return ms, list(results)

    # cfg 4: 14-qubit density matrix (2^28 amplitudes in total), 10 noisy layers + projector + trace + purity
    prog = TP.cfg4_program(14, 14014, layers=10, dump=False)
    nops = sum(1 for op in prog["ops"] if op[0].startswith(("apply", "mix")))
    ms, res = run_prog(prog, 2)
    out["cfg4_14q_density_matrix"] = {"workload": "cfg4: 14q density matrix, 10 layers of (H x14, CNOT chain, mixDepolarising x14, 2-qubit mixKrausMap x7) + applyMultiQubitProjector + calcTotalProb + calcPurity",
                                      "n_gpus": world, "ms_per_circuit": ms, "ops": nops, "ops_per_s": nops / (ms * 1e-3),
                                      "algorithmic_gbs_per_gpu": nops * 2 * AMP_BYTES * (1 << 28) / world / (ms * 1e-3) / 1e9,
                                      "total_prob": res[-2], "purity": res[-1]}

    # cfg 5: 28 qubits, 400 Pauli gadgets (2nd-order Trotter of a 200-term Hamiltonian) then calcExpecPauliStrSum
    prog = TP.cfg5_program(28, 28200, num_terms=200, dump=False)
    ms_t, _ = run_prog(dict(prog, ops=[prog["ops"][0]]), 2)
    ms_e, res = run_prog(dict(prog, ops=[prog["ops"][1]]), 2)
    out["cfg5_28q_trotter_paulisum"] = {"workload": "cfg5: 28q, applyTrotterizedPauliStrSumGadget (order 2, 400 gadgets) and calcExpecPauliStrSum (200 terms), each incl. initPlusState + syncQuESTEnv",
                                        "n_gpus": world, "trotter_ms": ms_t, "gadgets_per_s": 400 / (ms_t * 1e-3), "expec_ms": ms_e,
                                        "expec_algorithmic_gbs_per_gpu": 200 * AMP_BYTES * (1 << 28) / world / (ms_e * 1e-3) / 1e9, "expec_value": res[0]}
    return out


def run_product(args, rank, world, local_rank):
    P = Product(rank, world, local_rank)
    Q, torch, capi = P.Q, P.torch, P.capi
    n_local = args.qubits or 30
    logw = int(math.log2(world))
    n = n_local + logw
    local_amps = 1 << n_local
    config = build_config(n_local, world, args.workload, args.cpu_sample)
    cfg3 = args.workload == "cfg3"

    qft = [] if cfg3 else qft_stream(n)
    dense_ops = cfg3_stream(n) if cfg3 else dense_stream(n)
    dense = P.mats(dense_ops)
    num_gates = len(qft) + len(dense)
    bytes_step = sum(algorithmic_bytes(op, local_amps) for op in qft) + sum(algorithmic_bytes(op, local_amps) for op in dense_ops)
    qureg = P.create(n)

    def circuit():
        if not cfg3:
            Q.applyFullQuantumFourierTransform(qureg)
        for op, m in dense:
            P.gate(qureg, op, m)

    def step():                       # one step, executed (the queue is flushed), qubit order possibly still relabelled
        circuit()
        capi.call("qb_flush")

    def timed_steps(fn, steps):
        """K steps + the syncQuESTEnv() that leaves nothing undone, CUDA events, barrier both sides, max over ranks"""
        P.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        Q.syncQuESTEnv()
        e1.record()
        P.barrier()
        return P.max_over_ranks(e0.elapsed_time(e1)) / steps

    # ---------------- device-resident timing ----------------
    Q.initPlusState(qureg)
    Q.syncQuESTEnv()
    timed_steps(step, args.warmup)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    s0 = P.stats()
    ms_per_step = timed_steps(step, args.steps)
    s1 = P.stats()
    clocks = sampler.stop() if rank == 0 else None
    launches = s1["launches"] - s0["launches"]
    total_prob = Q.calcTotalProb(qureg)

    # ---------------- e2e: + state initialisation, host operands in, a probability read back ----------------
    e2e = None
    if not args.no_e2e:
        h2d = sum(64 if op[0] == "m1" else (256 if op[0] == "m2" else 16) for op in dense_ops) + 32 * sum(1 for op in qft if op[0] != "swap")

        def e2e_step():
            circuit()
            return Q.calcProbOfQubitOutcome(qureg, n - 1, 0)       # device -> host read of the step's result

        # same starting point as the device-timed loop (fresh state, identity qubit map, W warm-up steps): with lazy
        # relabelling the number of exchanges a step needs depends on the steps before it
        Q.initPlusState(qureg)
        for _ in range(args.warmup):
            e2e_step()
        Q.syncQuESTEnv()
        P.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            prob = e2e_step()
        Q.syncQuESTEnv()
        P.barrier()
        dt = P.max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": world * num_gates * args.steps / dt, "unit": "gates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8,
               "ms_per_step": 1e3 * dt / args.steps, "result_prob_of_top_qubit_0": prob,
               "api": "K x (applyFullQuantumFourierTransform + applyCompMatr1/2 x200 + calcProbOfQubitOutcome) + syncQuESTEnv"}

    # ---------------- section breakdown (untimed extra steps): QFT / dense / restore, and gate-by-gate exchange cost ----------------
    def ev():
        return torch.cuda.Event(enable_timing=True)
    sections = None
    if not cfg3:
        Q.initPlusState(qureg); Q.syncQuESTEnv(); P.barrier()
        a, b, c, d = ev(), ev(), ev(), ev()
        sa = P.stats()
        a.record(); Q.applyFullQuantumFourierTransform(qureg); capi.call("qb_flush"); b.record()
        sb = P.stats()
        for op, m in dense:
            P.gate(qureg, op, m)
        capi.call("qb_flush"); c.record()
        sc = P.stats()
        Q.syncQuESTEnv(); d.record(); P.barrier()
        sd = P.stats()
        sections = {"note": "one extra step with a flush after each section (the timed steps plan the whole step at once)"
                            + ("; with --lookahead the shim replays gate calls late, so the split between sections is not meaningful" if os.environ.get("QUEST_B200_LOOKAHEAD", "0") not in ("", "0") else ""),
                    "qft_ms": P.max_over_ranks(a.elapsed_time(b)), "dense_ms": P.max_over_ranks(b.elapsed_time(c)), "restore_ms": P.max_over_ranks(c.elapsed_time(d)),
                    "qft_launches": sb["launches"] - sa["launches"], "dense_launches": sc["launches"] - sb["launches"], "restore_launches": sd["launches"] - sc["launches"],
                    "dense_tile_passes": sc["passes"] - sb["passes"], "dense_rounds": sc["rounds"] - sb["rounds"], "dense_gates_after_absorption": (sc["tile_ops"] - sb["tile_ops"]) + (sc["direct_ops"] - sb["direct_ops"])}
        sections["dense_algorithmic_gbs"] = sum(algorithmic_bytes(op, local_amps) for op in dense_ops) / (sections["dense_ms"] * 1e-3) / 1e9
        sections["dense_avg_launch_ms"] = sections["dense_ms"] / max(1, sections["dense_launches"])
        sections["dense_fp64_tflops"] = 2 * (sc["fma"] - sb["fma"]) / (sections["dense_ms"] * 1e-3) / 1e12
    nvlink = None
    if world > 1:
        # exchange cost in isolation: the step again, gate by gate with a flush after each, events around the gates that
        # made the backend exchange half-shards (a target on a rank bit is pulled into the shard, or the final restore)
        Q.initPlusState(qureg); Q.syncQuESTEnv(); P.barrier()
        link_ms, link_bytes, nonlocal_gates = 0.0, 0, 0
        stream = [(op, None) for op in qft] + dense
        recs = []
        for op, m in stream:
            x0 = P.stats(); a, b = ev(), ev()
            a.record(); P.gate(qureg, op, m); capi.call("qb_flush"); b.record()
            x1 = P.stats()
            if x1["exchanges"] > x0["exchanges"]:
                recs.append((a, b, x1["link_bytes"] - x0["link_bytes"]))
        x0 = P.stats(); a, b = ev(), ev()
        a.record(); Q.syncQuESTEnv(); b.record()
        x1 = P.stats()
        if x1["exchanges"] > x0["exchanges"]:
            recs.append((a, b, x1["link_bytes"] - x0["link_bytes"]))
        P.barrier()
        for a, b, nb in recs:
            link_ms += a.elapsed_time(b); link_bytes += nb; nonlocal_gates += 1
        link_ms = P.max_over_ranks(link_ms)
        ach = link_bytes / (link_ms * 1e-3) / 1e9 if link_ms > 0 else None
        nvlink = {"achieved_gbs_per_dir_per_gpu": ach, "peak_gbs_per_dir": NVLINK_PEAK_GBS, "frac": (ach / NVLINK_PEAK_GBS) if ach else None,
                  "note": "bytes this GPU sent per direction / time of the gates that exchanged (their local gate work included), unfused gate-by-gate replay",
                  "exchanging_calls_per_step": nonlocal_gates, "exchanges_per_timed_step": (s1["exchanges"] - s0["exchanges"]) / args.steps,
                  "link_bytes_per_dir_per_timed_step": (s1["link_bytes"] - s0["link_bytes"]) / args.steps,
                  "min_exchange_ms_per_timed_step_at_peak": (s1["link_bytes"] - s0["link_bytes"]) / args.steps / (NVLINK_PEAK_GBS * 1e9) * 1e3}
    Q.destroyQureg(qureg)

    # ---------------- roofline of the dominant kernel ----------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
    traffic, traffic_src = None, None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        if n_local == tr.get("local_qubits"):
            traffic, traffic_src = tr["dram_bytes_per_launch"], tr["source"]
    except Exception:
        pass
    passes = (s1["passes"] + s1["direct_ops"] - s0["passes"] - s0["direct_ops"]) / args.steps
    kernel_launches = launches / args.steps
    achieved = bytes_step / (ms_per_step * 1e-3) / 1e9
    # bytes the gate passes streamed (the backend counts them per launch: tiles x 128 KiB, halves for restricted passes)
    physical = (s1["pass_bytes"] - s0["pass_bytes"]) / args.steps / (ms_per_step * 1e-3) / 1e9
    fp64 = 2 * (s1["fma"] - s0["fma"]) / args.steps / (ms_per_step * 1e-3) / 1e12
    roofline = {"kernel": "k_tile_pass (fused multi-gate pass over the shard; >95% of the step's device time) + the few direct single-gate kernels",
                "bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "launches_per_step": kernel_launches, "gates_per_launch": num_gates / max(kernel_launches, 1),
                "algorithmic_bytes_per_launch": bytes_step / max(kernel_launches, 1), "avg_launch_ms": ms_per_step / max(kernel_launches, 1),
                "physical_gbs_estimate": physical, "physical_frac": physical / peak_gbs,
                "fp64_tflops": fp64, "fp64_peak_tflops": FP64_PEAK_TFLOPS, "fp64_frac": fp64 / FP64_PEAK_TFLOPS,
                "reading": "frac > 1 = fusion gain over one gate per HBM pass; the fused pass itself sits under BOTH roofs: physical_frac of HBM, fp64_frac of the FP64 pipe",
                "sections": sections}
    if nvlink:
        roofline["nvlink"] = nvlink

    secondary = None
    if not args.no_secondary and not cfg3:
        try:
            secondary = secondary_configs(P, n_local, full=True)
        except Exception as exc:                       # report, never fake
            secondary = {"error": f"{type(exc).__name__}: {exc}"[:400]}
    # the 34-qubit cfg-3 circuit time (BASELINE metric) where this run holds it: 8 GPUs at 2^31 amplitudes each, 4 GPUs at 2^32.
    # It is NOT put into `config`, which stays a pure function of the workload so that both arms print the same object.
    cfg3_headline = None
    for key, val in (secondary or {}).items():
        if key.startswith("cfg3_") and isinstance(val, dict) and "ms_per_circuit" in val:
            rec = {"qubits": int(key[5:-1]), "n_gpus": world, "ms_per_circuit": val["ms_per_circuit"], "gates_ms": val.get("gates_ms"),
                   "restore_canonical_order_ms": val.get("restore_canonical_order_ms")}
            if cfg3_headline is None or rec["qubits"] > cfg3_headline["qubits"]:
                cfg3_headline = rec

    cpu_baseline = None
    if rank == 0 and not args.no_cpu_baseline:
        try:
            ref = run_reference(n_local, 1, 1, args.cpu_sample, args.cpu_budget)
            cpu_baseline = {k: ref[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as exc:                       # report, never fake
            cpu_baseline = {"value": None, "unit": "gates/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {exc}"[:300]}

    if rank == 0:
        detail = {"total_prob_after_run": total_prob, "circuit_gates_per_s": num_gates / (ms_per_step * 1e-3),
                  "parallelism": f"state sharded over {world} GPUs on the top {logw} qubits" if world > 1 else "single GPU",
                  "p2p_nvlink_kernels": bool(P.lib.qb_p2p_is_available()) if world > 1 else None,
                  "lookahead_window": int(os.environ.get("QUEST_B200_LOOKAHEAD", "0") or 0),
                  "timed_region": f"{args.steps} x (circuit + queue flush) + syncQuESTEnv (canonical qubit order restored), CUDA events, max over ranks"}
        line = {"metric": METRIC, "value": world * num_gates / (ms_per_step * 1e-3), "unit": "gates/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config, "roofline": roofline,
                "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "detail": detail,
                "cfg3": cfg3_headline, "secondary": secondary}
        print(json.dumps(line))
    P.barrier()
    Q.finalizeQuESTEnv()
    if P.dist:
        P.dist.barrier()
        P.dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--qubits", type=int, default=0, help="override the local-qubit count (default 30 per GPU)")
    ap.add_argument("--cpu-sample", type=int, default=24, help="gates per step of the CPU reference sample")
    ap.add_argument("--cpu-budget", type=float, default=25.0, help="seconds of CPU work per reference step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the cfg 3/4/5 records")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3"], help="cfg3: 100 random {H,Rx,CompMatr1,CNOT,CompMatr2} gates on initPlusState")
    ap.add_argument("--lookahead", type=int, default=None,
                    help="N>1: QUEST_B200_LOOKAHEAD window of the sharding shim (opt-in look-ahead swap-in victim choice; default: the library's, off)")
    args = ap.parse_args()
    if args.lookahead is not None:
        os.environ["QUEST_B200_LOOKAHEAD"] = str(args.lookahead)     # read by libQuEST.so at the first gate

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_local = args.qubits or 30

    if args.impl == "reference":
        if rank != 0:
            return 0
        gpus = max(args.gpus, world)
        config = build_config(n_local, gpus, args.workload, args.cpu_sample)
        ref = run_reference(n_local, args.steps, max(args.warmup, 1), args.cpu_sample, args.cpu_budget)
        line = {"impl": "reference", "metric": METRIC, "value": ref["value"], "unit": "gates/s", "n_gpus": gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ref["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {k: ref[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": ref["value"], "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0,
                "detail": {"note": "value = gates of the sample / CPU time: the reference applies one gate per pass over the 16 GiB state, so per-gate cost is "
                                   "what the sample measures; ms_per_step is the time of the SAMPLE, not of the 680-gate step"}}
        print(json.dumps(line))
        return 0
    return run_product(args, rank, world, local_rank)


if __name__ == "__main__":
    sys.exit(main())
