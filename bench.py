#!/usr/bin/env python
"""bench.py -- the reference's headline workload on the B200 backend (BASELINE.json cfg 2).

workload (N=1): 30-qubit fp64 statevector; one STEP = applyFullQuantumFourierTransform (30 H + 435
controlled phase shifts + 15 SWAPs, api/operations.cpp:1934-1953) followed by 200 random dense 1- and
2-qubit unitaries on uniformly random targets (seed 20302) = 680 gates per step.
For N>1 GPUs the state has 30+log2(N) qubits sharded over the ranks (2^30 amplitudes per GPU, weak scaling).

Both timings drive QuEST's public API on the drop-in libQuEST.so (the call a user makes; the reference-facing
boundary), so the sharding shim, the deferred gate queue and the kernels are all on the measured path.

  value   gates/s with the state resident in HBM, CUDA events on the backend's stream.  The timed region is
          K x (QFT + 200 dense gates) and ENDS with syncQuESTEnv(): the backend defers work (fusable gates are queued,
          uncontrolled SWAPs only relabel qubits), and syncQuESTEnv() is what guarantees nothing is left undone
          (it restores the canonical qubit order and drains the stream).  N>1: the unit is one gate applied to one
          2^30-amplitude shard, so value = N x circuit gates / step time (quest_b200/dist_bench.py).
  e2e     the same metric end to end with HOST buffers: initZeroState, the circuit (every matrix travels host ->
          device inside its call), calcProbOfQubitOutcome read back device -> host; host wall clock.  A statevector
          simulator's state is created on the device by the reference's own API (createQureg/initZeroState,
          api/qureg.cpp:143-174) and never crosses PCIe; its per-step host inputs are the gate operands.
  roofline   the dense-gate section: algorithmic bytes of its 200 gates (2*16*2^30 each) / CUDA-event
          time of the section, against MEASURED_PEAKS.json hbm_gbs; frac > 1 is the gain of fusing several gates into
          one HBM pass; physical_frac is the traffic actually moved (one read + one write of the state per launch,
          confirmed by ncu in profiles/) against the same peak.
  cpu_baseline / --impl reference: the UNMODIFIED reference CPU/OpenMP library (oracle/_ref/libQuEST.so) on the
          box's host cores, timed on a bounded stratified sample of the same 680-gate stream at the same size.
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


# ------------------------------------------------------------------------------------------------
# the gate stream (identical for every arm)
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
print("REFJSON " + json.dumps(dict(times=times, gates=gates, prob=prob, stride=stride, threads=int(os.environ.get("OMP_NUM_THREADS", "1")))))
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
                    "sample": f"{d['gates'][0]} gates/step: every {d['stride']}th gate of the 680-gate cfg-2 stream at {n} qubits, "
                              f"{len(d['times'])} timed step(s), OMP_NUM_THREADS={cores}",
                    "ms_per_step": 1e3 * total_t / len(d["times"]), "total_prob": d["prob"]}
    raise RuntimeError(f"reference worker failed rc={r.returncode}\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")


# ------------------------------------------------------------------------------------------------
# the product arm
# ------------------------------------------------------------------------------------------------
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
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3"], help="cfg3 (N>1 only): 100 random {H,Rx,CompMatr1,CNOT,CompMatr2} gates on initPlusState")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_local = args.qubits or 30
    n = n_local + int(math.log2(max(1, world)))
    config = {"workload": f"cfg2: {n}q fp64 statevector, applyFullQuantumFourierTransform + {NUM_DENSE} random dense 1/2-qubit gates",
              "qubits": n, "amps_per_gpu": 1 << n_local, "gates_per_step": len(qft_stream(n)) + NUM_DENSE, "seed": SEED,
              "l2_policy": "state (16 GiB per GPU) is far larger than L2; every gate streams it from HBM"}

    if args.workload == "cfg3":
        config.update({"workload": f"cfg3: {n}q fp64 statevector, initPlusState + 100 random gates from {{H, Rx, CompMatr1, CNOT, CompMatr2}}, targets uniform over all qubits",
                       "gates_per_step": 100, "seed": 34008})

    if args.impl == "reference":
        if rank != 0:
            return 0
        ref = run_reference(n_local, args.steps, max(args.warmup, 1), args.cpu_sample, args.cpu_budget)
        if world > 1 or args.gpus > 1:
            config["reference_note"] = (f"the metric's unit is one gate applied to 2^{n_local} amplitudes; the CPU reference is timed on that unit "
                                        f"({n_local}-qubit state, all host cores): it cannot hold the {n}-qubit sharded state")
        line = {"impl": "reference", "metric": "30q fp64 gates/s", "value": ref["value"], "unit": "gates/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ref["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {k: ref[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": ref["value"], "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    import torch
    from quest_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the quest_b200 backend has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    capi.call("qb_bind_device", local_rank)

    if world > 1:
        from quest_b200 import dist_bench
        return dist_bench.run(args, rank, world, local_rank, n, n_local, config, dist)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")

    qft, dense = qft_stream(n), dense_stream(n)
    local_amps = 1 << n_local
    bytes_qft = sum(algorithmic_bytes(op, local_amps) for op in qft)
    bytes_dense = sum(algorithmic_bytes(op, local_amps) for op in dense)
    num_gates = len(qft) + len(dense)

    # Both timings drive QuEST's public API on the drop-in libQuEST.so (the call a user makes); the backend library
    # underneath is the same shared object that `capi` binds, so qb_flush / qb_launch_count see the same queue.
    from quest_b200 import quest_api as qa
    Q = qa.QuEST(qa.B200_LIB)
    Q.initCustomQuESTEnv(0, 1, 0)
    qureg = Q.createCustomQureg(n, 0, 0, 1, 0)
    assert qureg.isGpuAccelerated == 1
    mats = [(op, Q.getCompMatr1(op[2]) if op[0] == "m1" else Q.getCompMatr2(op[3])) for op in dense]

    def apply_dense():
        for op, m in mats:
            if op[0] == "m1":
                Q.applyCompMatr1(qureg, op[1], m)
            else:
                Q.applyCompMatr2(qureg, op[1], op[2], m)

    # ---------------- device-resident timing: state already in HBM, CUDA events on the backend's stream ----------------
    # The backend defers work: fusable gates are queued until qb_flush, and uncontrolled SWAPs (the QFT's final layer)
    # only relabel qubits until something needs the canonical order.  The timed region therefore ENDS with
    # syncQuESTEnv(), which restores the canonical order and drains the stream, so nothing is left undone.
    Q.initPlusState(qureg)
    for _ in range(args.warmup):
        Q.applyFullQuantumFourierTransform(qureg); capi.call("qb_flush")
        apply_dense(); capi.call("qb_flush")
    Q.syncQuESTEnv()
    launches0 = capi.lib().qb_launch_count()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    ev_end = torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local_rank); sampler.start()
    torch.cuda.synchronize()
    dense_launches = 0
    for k in range(args.steps):
        ev[k][0].record()
        Q.applyFullQuantumFourierTransform(qureg); capi.call("qb_flush")
        ev[k][1].record()
        l0 = capi.lib().qb_launch_count()
        apply_dense(); capi.call("qb_flush")
        dense_launches += capi.lib().qb_launch_count() - l0
        ev[k][2].record()
    Q.syncQuESTEnv()
    ev_end.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    launches = capi.lib().qb_launch_count() - launches0
    t_qft = sum(e[0].elapsed_time(e[1]) for e in ev) / args.steps          # ms
    t_dense = sum(e[1].elapsed_time(e[2]) for e in ev) / args.steps
    t_restore = ev[-1][2].elapsed_time(ev_end)                             # canonical qubit order restored once, after K steps
    ms_per_step = ev[0][0].elapsed_time(ev_end) / args.steps
    total_prob = Q.calcTotalProb(qureg)
    config["timed_region"] = "K x (QFT + 200 dense gates) + syncQuESTEnv (restores canonical qubit order after lazily relabelled SWAPs)"
    config["restore_ms_total"] = t_restore

    # ---------------- e2e through the same API: + state initialisation, host matrices in, probability out ----------------
    e2e = None
    if not args.no_e2e:
        h2d = sum(64 if op[0] == "m1" else 256 for op in dense) + 32 * sum(1 for op in qft if op[0] != "swap")

        def e2e_step():
            Q.initZeroState(qureg)
            Q.applyFullQuantumFourierTransform(qureg)
            apply_dense()
            return Q.calcProbOfQubitOutcome(qureg, n - 1, 0)          # device->host read of the step's result

        for _ in range(args.warmup):
            e2e_step()
        Q.syncQuESTEnv()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            prob = e2e_step()
        Q.syncQuESTEnv()
        dt = time.perf_counter() - t0
        e2e = {"value": num_gates * args.steps / dt, "unit": "gates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8,
               "ms_per_step": 1e3 * dt / args.steps, "result_prob_of_top_qubit_0": prob,
               "api": "initZeroState + applyFullQuantumFourierTransform + applyCompMatr1/2 x200 + calcProbOfQubitOutcome"}
    Q.destroyQureg(qureg)

    # the 200 dense gates run as `passes` kernel launches (tile-engine passes fusing several gates, or direct kernels);
    # each launch streams the 2*16*2^n-byte state once, while its ALGORITHMIC bytes are the sum over the gates it applies
    passes = dense_launches / args.steps
    achieved = bytes_dense / (t_dense * 1e-3) / 1e9
    physical = passes * 2 * AMP_BYTES * local_amps / (t_dense * 1e-3) / 1e9
    roofline = {"kernel": "dense 1/2-qubit gate section: k_tile_pass (fused multi-gate passes) + direct k_tuple kernels",
                "bound": "hbm", "achieved": achieved, "peak": peak_gbs,
                "unit": "GB/s", "frac": achieved / peak_gbs,
                # dram__bytes_read.sum + dram__bytes_write.sum per k_tile_pass launch, ncu --set full of this command
                # (profiles/r1_final_launches_and_ncu.md): one read + one write of the state, whatever the number of fused gates
                "traffic": 34.30e9 if n_local == 30 else None, "traffic_source": "ncu --set full, profiles/r1_final_launches_and_ncu.md",
                "peak_source": peak_src,
                "launches_per_step": passes, "gates_per_launch": len(dense) / max(passes, 1),
                "algorithmic_bytes_per_launch": bytes_dense / max(passes, 1), "avg_launch_ms": t_dense / max(passes, 1),
                "physical_gbs_estimate": physical, "physical_frac": physical / peak_gbs,
                "qft_section": {"achieved_gbs": bytes_qft / (t_qft * 1e-3) / 1e9, "ms": t_qft, "gates": len(qft)},
                "whole_step_gbs": (bytes_qft + bytes_dense) / (ms_per_step * 1e-3) / 1e9}
    config["total_prob_after_run"] = total_prob

    cpu_baseline = None
    if not args.no_cpu_baseline:
        try:
            ref = run_reference(n_local, 1, 1, args.cpu_sample, args.cpu_budget)
            cpu_baseline = {k: ref[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as exc:                       # report, never fake
            cpu_baseline = {"value": None, "unit": "gates/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {exc}"[:300]}

    line = {"metric": "30q fp64 gates/s", "value": num_gates / (ms_per_step * 1e-3), "unit": "gates/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config, "roofline": roofline,
            "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
