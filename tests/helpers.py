"""Shared test plumbing: run programs on a QuEST library in a worker process, compare outputs."""
import os
import pickle
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libQuEST.so")
B200_LIB = os.path.join(ROOT, "quest_b200", "lib", "libQuEST.so")

# the north-star tolerance: 1e-12 relative L2 on amplitudes / expectation values in fp64
TOL = 1e-12


def run_programs(which, progs, timeout=1800, env=None):
    """which: 'b200' (the product, GPU) or 'ref' (unmodified reference CPU build)."""
    with tempfile.TemporaryDirectory() as d:
        src, dst = os.path.join(d, "in.pkl"), os.path.join(d, "out.pkl")
        pickle.dump(progs, open(src, "wb"))
        e = dict(os.environ)
        e.update(env or {})
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_worker.py"), which, src, dst],
                           capture_output=True, text=True, timeout=timeout, env=e)
        if r.returncode != 0 or not os.path.exists(dst):
            raise RuntimeError(f"worker({which}) failed rc={r.returncode}\nstdout:\n{r.stdout[-3000:]}\nstderr:\n{r.stderr[-3000:]}")
        return pickle.load(open(dst, "rb"))


def rel_l2(a, b):
    a, b = np.asarray(a, dtype=np.complex128).ravel(), np.asarray(b, dtype=np.complex128).ravel()
    denom = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / (denom if denom > 0 else 1.0))


def flatten_result(r):
    if r is None or isinstance(r, str):
        return None
    if isinstance(r, np.ndarray) and np.iscomplexobj(r):
        return np.ascontiguousarray(r).view(np.float64).ravel()
    return np.asarray(r, dtype=np.float64).ravel()


def assert_outputs_match(got, want, tol=TOL, label="", int_exact_ops=()):
    """amplitude dumps: relative L2 <= tol; scalar results: |d| <= tol * max(1, |want|)."""
    for name, w in want["dumps"].items():
        g = got["dumps"][name]
        assert g.shape == w.shape, f"{label} dump {name}: shape {g.shape} vs {w.shape}"
        err = rel_l2(g, w)
        assert err <= tol, f"{label} dump {name}: rel-L2 {err:.3e} > {tol:g}"
    assert len(got["results"]) == len(want["results"])
    for i, (g, w) in enumerate(zip(got["results"], want["results"])):
        gf, wf = flatten_result(g), flatten_result(w)
        if wf is None:
            continue
        assert gf is not None and gf.shape == wf.shape, f"{label} result {i}: {g} vs {w}"
        if i in int_exact_ops:
            assert np.array_equal(gf, wf), f"{label} result {i} (bit-exact): {g} vs {w}"
            continue
        scale = max(1.0, float(np.max(np.abs(wf)))) if wf.size else 1.0
        err = float(np.max(np.abs(gf - wf))) if wf.size else 0.0
        assert err <= tol * scale, f"{label} result {i}: {g} vs {w} (abs err {err:.3e})"


def load_golden(name):
    return pickle.load(open(os.path.join(GOLDEN_DIR, name), "rb"))


def num_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def run_programs_distributed(progs, world, timeout=420, env=None, port=29731, which="b200dist"):
    """run the programs on `world` GPUs, one worker process per GPU (the launcher model of the backend:
    RANK / WORLD_SIZE / LOCAL_RANK / MASTER_PORT in the environment, like torchrun).  Returns rank 0's outputs
    with every density-matrix / distributed dump re-assembled from the per-rank shards (rank r holds the
    global indices [r*N, (r+1)*N), api/qureg.cpp:42-74)."""
    with tempfile.TemporaryDirectory() as d:
        src, dst = os.path.join(d, "in.pkl"), os.path.join(d, "out.pkl")
        pickle.dump(progs, open(src, "wb"))
        procs = []
        for r in range(world):
            e = dict(os.environ)
            e.update(env or {})
            # fewer GPUs than ranks: ranks share devices and the backend switches to its shared-memory / CUDA-IPC transport
            e.update(RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r % max(1, num_gpus())), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                     QUEST_B200_ID_FILE=os.path.join(d, "nccl_id"))
            procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_worker.py"), which, src, dst],
                                          stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=e))
        outs, fails = [], []
        for r, p in enumerate(procs):
            try:
                so, se = p.communicate(timeout=timeout)
            except subprocess.TimeoutExpired:
                for q in procs:
                    q.kill()
                raise RuntimeError(f"distributed worker {r} timed out")
            if p.returncode != 0:
                fails.append(f"rank {r} rc={p.returncode}\n{so[-2000:]}\n{se[-3000:]}")
        if fails:
            raise RuntimeError("distributed workers failed:\n" + "\n".join(fails))
        per_rank = [pickle.load(open(f"{dst}.{r}", "rb")) for r in range(world)]
    merged = per_rank[0]
    for k, out in enumerate(merged):
        for name in list(out["dumps"]):
            info = out["info"][name]
            if info["isDistributed"] and out["dumps"][name].size == info["numAmpsPerNode"]:
                out["dumps"][name] = np.concatenate([per_rank[r][k]["dumps"][name] for r in range(world)])
        # every rank must have computed identical scalars (they are all-reduced)
        for r in range(1, world):
            for a, b in zip(out["results"], per_rank[r][k]["results"]):
                fa, fb = flatten_result(a), flatten_result(b)
                assert (fa is None and fb is None) or np.array_equal(fa, fb), f"rank {r} disagrees with rank 0: {a} vs {b}"
    return merged
