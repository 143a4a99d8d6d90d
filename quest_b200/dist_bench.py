"""bench.py's multi-GPU arm: the cfg-2 workload weak-scaled over N GPUs of one box (2^30 amplitudes per GPU,
30 + log2 N qubits), one process per GPU, driven through QuEST's public API on the drop-in library so that
the sharding logic (quest_b200/shim/localiser_b200.cpp), the NCCL control plane and the NVLink peer-memory
kernels are all on the measured path.

Timing: CUDA events on the library's stream (the legacy default stream, which is also torch's current
stream), bracketed by barrier + synchronize, max over ranks.  The metric's unit is one gate applied to 2^30
amplitudes (BASELINE.json: "30q fp64 gates/s"): every rank applies every gate of the circuit to its 2^30-amplitude
shard, so the job processes world x num_gates such units per step and `value` = world * num_gates / step time.
"""
import json
import math
import os
import time

import numpy as np


def _classify(op, n_local):
    """which qubits of a gate are rank bits, and what that costs on the link (bytes per GPU per direction)"""
    full = 16 * (1 << n_local)
    kind = op[0]
    if kind == "h" or kind == "m1":
        return ("dense1_prefix", full) if op[1] >= n_local else ("local", 0)
    if kind == "cphase":
        return ("local", 0)                       # diagonal: never communicates
    if kind == "swap":
        hi = max(op[1], op[2]); lo = min(op[1], op[2])
        if lo >= n_local:
            return ("swap_prefix_prefix", full // 2)     # half of the ranks trade whole shards
        return ("swap_prefix_suffix", full // 2) if hi >= n_local else ("local", 0)
    if kind == "m2":
        npre = sum(1 for t in (op[1], op[2]) if t >= n_local)
        return (f"dense2_{npre}prefix", 2 * npre * (full // 2)) if npre else ("local", 0)
    return ("local", 0)


def run(args, rank, world, local_rank, n, n_local, config, dist):
    import torch
    import bench
    from quest_b200 import capi, quest_api as qa

    dev = torch.device("cuda", local_rank)
    # share NCCL's unique id for the backend's own communicator over torch.distributed (control plane only)
    idbuf = (__import__("ctypes").c_char * 128)()
    if rank == 0:
        capi.call("qb_comm_get_unique_id", idbuf)
    t = torch.tensor(list(bytes(idbuf)), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=0)
    os.environ["QUEST_B200_NCCL_ID"] = bytes(t.cpu().tolist()).hex()

    Q = qa.QuEST(qa.B200_LIB)
    Q.initCustomQuESTEnv(1, 1, 0)
    qureg = Q.createCustomQureg(n, 0, 1, 1, 0)
    assert qureg.isGpuAccelerated == 1 and qureg.isDistributed == 1 and qureg.numNodes == world
    p2p = int(capi.lib().qb_p2p_is_available())

    cfg3 = getattr(args, "workload", "cfg2") == "cfg3"
    qft = [] if cfg3 else bench.qft_stream(n)
    dense = bench.cfg3_stream(n) if cfg3 else bench.dense_stream(n)
    mats = [(op, Q.getCompMatr1(op[2]) if op[0] == "m1" else (Q.getCompMatr2(op[3]) if op[0] == "m2" else None)) for op in dense]
    num_gates = len(qft) + len(dense)

    def apply_dense(op, m):
        if op[0] == "m1":
            Q.applyCompMatr1(qureg, op[1], m)
        elif op[0] == "m2":
            Q.applyCompMatr2(qureg, op[1], op[2], m)
        elif op[0] == "h":
            Q.applyHadamard(qureg, op[1])
        elif op[0] == "rx":
            Q.applyRotateX(qureg, op[1], op[2])
        else:
            Q.applyControlledPauliX(qureg, op[1], op[2])

    def gates():
        if not cfg3:
            Q.applyFullQuantumFourierTransform(qureg)
        for op, m in mats:
            apply_dense(op, m)
        capi.call("qb_flush")          # launch whatever the backend deferred for fusion, so that events bracket it

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()

    Q.initPlusState(qureg) if cfg3 else Q.initZeroState(qureg)
    for _ in range(args.warmup):
        gates()
    barrier()
    launches0 = capi.lib().qb_launch_count()
    import ctypes as _C
    _nx = _C.c_ulonglong(); capi.lib().qb_p2p_stats(_C.byref(_nx), None); exch0 = _nx.value
    sampler = bench.ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        gates()
    Q.syncQuESTEnv()                   # restores the canonical qubit order (lazy relabelling) inside the timed region
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = capi.lib().qb_launch_count() - launches0
    capi.lib().qb_p2p_stats(_C.byref(_nx), None); timed_exchanges = _nx.value - exch0
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = ms.item() / args.steps

    # end-to-end: state init + circuit + a probability read back to the host, host wall clock around the API calls
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        Q.initPlusState(qureg) if cfg3 else Q.initZeroState(qureg)
        gates()
        prob = Q.calcProbOfQubitOutcome(qureg, n - 1, 0)
    barrier()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_s = dt.item() / args.steps

    # per-gate breakdown (one extra untimed step, gates issued and flushed one at a time): with lazy qubit relabelling
    # a gate is "non-local" when it made the backend exchange half-shards (a target sat on a rank bit and was pulled
    # into the shard) -- seen as a step of the backend's exchange counter while the gate was being issued
    import ctypes as C
    def exchanges():
        a, b = C.c_ulonglong(), C.c_ulonglong()
        capi.lib().qb_p2p_stats(C.byref(a), C.byref(b))
        return a.value, b.value
    classes = {}
    stream = [(op, None) for op in qft] + mats
    evs = []
    barrier()
    x0 = exchanges()
    for op, m in stream:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        before = exchanges()
        a.record()
        if op[0] == "cphase":
            Q.applyTwoQubitPhaseShift(qureg, op[1], op[2], op[3])
        elif op[0] == "swap":
            Q.applySwap(qureg, op[1], op[2])
        else:
            apply_dense(op, m)
        capi.call("qb_flush")          # gate-by-gate here: this profile shows the unfused per-gate cost
        b.record()
        after = exchanges()
        evs.append((op, a, b, after[0] - before[0], after[1] - before[1]))
    Q.syncQuESTEnv()
    barrier()
    x1 = exchanges()
    for op, a, b, nx, link in evs:
        cls = "relabelled_swap (no amplitude moves)" if op[0] == "swap" else (f"{'dense' if op[0] in ('h', 'm1', 'm2') else op[0]}_with_{nx}_swap_in" if nx else "local")
        c = classes.setdefault(cls, {"gates": 0, "ms": 0.0, "link_bytes_per_dir": 0, "hbm_algorithmic_bytes": 0})
        c["gates"] += 1
        c["ms"] += a.elapsed_time(b)
        c["link_bytes_per_dir"] += link
        c["hbm_algorithmic_bytes"] += bench.algorithmic_bytes(op, 1 << n_local)
    for c in classes.values():
        c["link_gbs_per_dir"] = c["link_bytes_per_dir"] / (c["ms"] * 1e-3) / 1e9 if c["link_bytes_per_dir"] else None
        c["hbm_algorithmic_gbs"] = c["hbm_algorithmic_bytes"] / (c["ms"] * 1e-3) / 1e9 if c["ms"] > 0 else None
        c["ms_per_gate"] = c["ms"] / c["gates"]
    classes["exchanges_in_breakdown_step_incl_restore"] = x1[0] - x0[0]
    total_prob = Q.calcTotalProb(qureg)
    Q.destroyQureg(qureg)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(bench.ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_gbs = peaks.get("hbm_gbs", 6650.0)
        local = classes.get("local", {"hbm_algorithmic_gbs": 0.0, "hbm_algorithmic_bytes": 0, "gates": 1})
        nonlocal_cls = {k: c for k, c in classes.items() if isinstance(c, dict) and c["link_bytes_per_dir"]}
        link_bytes = sum(c["link_bytes_per_dir"] for c in nonlocal_cls.values())
        link_ms = sum(c["ms"] for c in nonlocal_cls.values())
        config.update({"unit_definition": "one gate applied to one 2^30-amplitude shard; a gate of the sharded circuit counts once per GPU",
                       "circuit_gates_per_s": num_gates / (ms_per_step * 1e-3),
                       "parallelism": f"state sharded over {world} GPUs on the top {int(math.log2(world))} qubits",
                       "p2p_nvlink_kernels": bool(p2p), "total_prob_after_run": total_prob, "gate_classes": classes})
        line = {"metric": f"{n_local}q-per-GPU fp64 gates/s (weak scaling)", "value": world * num_gates / (ms_per_step * 1e-3), "unit": "gates/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "roofline": {"kernel": "local gate kernels (per GPU)", "bound": "hbm", "achieved": local["hbm_algorithmic_gbs"], "peak": peak_gbs,
                             "unit": "GB/s", "frac": local["hbm_algorithmic_gbs"] / peak_gbs, "traffic": None,
                             "peak_source": "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)",
                             "nvlink": {"achieved_gbs_per_dir_per_gpu": link_bytes / (link_ms * 1e-3) / 1e9 if link_ms else None,
                                        "peak_gbs_per_dir": 770.0, "peak_source": "B200_PROFILING.md measured peer copy (900 nominal)",
                                        "non_local_gates": sum(c["gates"] for c in nonlocal_cls.values()),
                                        "exchanges_per_timed_step": timed_exchanges / args.steps}},
                "cpu_baseline": None,
                "e2e": {"value": world * num_gates / e2e_s, "unit": "gates/s", "h2d_bytes_per_step": sum(64 if op[0] == "m1" else (256 if op[0] == "m2" else 16) for op in dense),
                        "d2h_bytes_per_step": 8, "ms_per_step": 1e3 * e2e_s, "result_prob_of_top_qubit_0": prob},
                "gpu_launches": int(launches), "clocks": clocks}
        print(json.dumps(line))
    Q.finalizeQuESTEnv()
    dist.barrier()
    dist.destroy_process_group()
    return 0
