#!/usr/bin/env python
"""Secondary measurements: the BASELINE.json configurations that are not bench.py's headline (cfg 1, 4, 5), through
QuEST's public API on the drop-in library, with the unmodified reference CPU library timed beside them on a bounded
sample.  Prints one JSON object; `python tools/bench_configs.py > profiles/rNN_configs.json` on the GPU box."""
import json
import os
import pickle
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from quest_b200 import quest_api as qa            # noqa: E402
from quest_b200.program import run_program        # noqa: E402
from tests import programs as P                   # noqa: E402

_REF = r"""
import os, pickle, sys, time
sys.path.insert(0, {root!r})
from quest_b200 import quest_api as qa
from quest_b200.program import run_program
prog = pickle.load(open({src!r}, "rb"))
Q = qa.QuEST(os.path.join({root!r}, "oracle", "_ref", "libQuEST.so")); Q.initCustomQuESTEnv(0, 0, 1)
t0 = time.perf_counter(); out = run_program(Q, prog); dt = time.perf_counter() - t0
pickle.dump(dict(seconds=dt, results=out["results"]), open({dst!r}, "wb"))
"""


def time_ref(prog):
    with tempfile.TemporaryDirectory() as d:
        src, dst = os.path.join(d, "p.pkl"), os.path.join(d, "o.pkl")
        pickle.dump(prog, open(src, "wb"))
        env = dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count()), OMP_PROC_BIND="spread")
        subprocess.run([sys.executable, "-c", _REF.format(root=ROOT, src=src, dst=dst)], check=True, env=env, timeout=3000)
        return pickle.load(open(dst, "rb"))


def time_gpu(Q, prog, reps=3):
    for spec in prog["quregs"].values():
        spec["custom"] = [0, 1, 0]
    best, out = None, None
    for _ in range(reps):
        Q.syncQuESTEnv()
        t0 = time.perf_counter()
        out = run_program(Q, prog)
        Q.syncQuESTEnv()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best, out


def count_gates(prog):
    return sum(1 for op in prog["ops"] if op[0].startswith(("apply", "mix")))


def main():
    Q = qa.QuEST(qa.B200_LIB)
    Q.initCustomQuESTEnv(0, 1, 0)
    res = {"host_cores": os.cpu_count()}

    # cfg 1: exactly the BASELINE configuration on both arms
    p = P.cfg1_program(20, 12345, 200); p["dump"] = []
    g, og = time_gpu(Q, p)
    r = time_ref(P.cfg1_program(20, 12345, 200) | {"dump": []})
    n = count_gates(p)
    res["cfg1_20q_random_circuit"] = {"gates": n, "gpu_s": g, "gpu_gates_per_s": n / g, "cpu_s": r["seconds"], "cpu_gates_per_s": n / r["seconds"],
                                      "note": "whole program incl. createQureg/init and 21 probability reductions; CPU = reference OpenMP build",
                                      "total_prob_gpu": og["results"][-21], "total_prob_cpu": r["results"][-21]}

    # cfg 4: 14-qubit density matrix, 10 noisy layers; CPU arm: 1 layer of the same circuit
    p = P.cfg4_program(14, 14014, layers=10, dump=False)
    g, og = time_gpu(Q, p, reps=2)
    p1 = P.cfg4_program(14, 14014, layers=1, dump=False)
    r = time_ref(p1)
    n, n1 = count_gates(p), count_gates(p1)
    res["cfg4_14q_density_matrix"] = {"ops": n, "gpu_s": g, "gpu_ops_per_s": n / g, "cpu_sample_ops": n1, "cpu_s": r["seconds"], "cpu_ops_per_s": n1 / r["seconds"],
                                      "algorithmic_bytes_per_op": 2 * 16 * (1 << 28), "gpu_algorithmic_gbs": n * 2 * 16 * (1 << 28) / g / 1e9,
                                      "total_prob_gpu": og["results"][-2], "purity_gpu": og["results"][-1],
                                      "note": "H and CNOT count once per API call although each is two passes (ket, bra); the 2-qubit Kraus map is a 4-target dense pass"}

    # cfg 5: 28 qubits, 200-term Hamiltonian: 400 Pauli gadgets (2nd-order Trotter) + calcExpecPauliStrSum
    p = P.cfg5_program(28, 28200, num_terms=200, dump=False)
    ptrot = dict(p, ops=[p["ops"][0]]); pexp = dict(p, ops=[p["ops"][1]])
    gt, _ = time_gpu(Q, ptrot, reps=2)
    ge, oe = time_gpu(Q, pexp, reps=2)
    ps = P.cfg5_program(28, 28200, num_terms=8, dump=False)
    r = time_ref(dict(ps, ops=[ps["ops"][0]]))
    r2 = time_ref(dict(ps, ops=[ps["ops"][1]]))
    res["cfg5_28q_trotter_paulisum"] = {"gadgets": 400, "gpu_trotter_s": gt, "gpu_gadgets_per_s": 400 / gt, "gpu_expec_200_terms_s": ge,
                                        "gpu_expec_algorithmic_gbs": 200 * 16 * (1 << 28) / ge / 1e9, "expec_value_gpu": oe["results"][0],
                                        "cpu_sample": "8-term Hamiltonian: 16 gadgets, 8-term expectation (incl. createQureg + initPlusState)",
                                        "cpu_trotter_s": r["seconds"], "cpu_gadgets_per_s": 16 / r["seconds"], "cpu_expec_8_terms_s": r2["seconds"]}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
