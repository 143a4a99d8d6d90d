"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/quest_b200.h declares, its host-side index algebra is exact, and compute entry points FAIL LOUDLY
when there is no CUDA device (there is no CPU fallback)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from quest_b200 import capi


def test_library_exports_every_declared_symbol():
    lib = capi.lib()
    names = capi.declared_symbols()
    assert len(names) >= 100
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in quest_b200.h but not exported: {missing}"
    assert lib.qb_abi_version() == 1
    # the host-side self-tests are compiled OUT of the product library (they live in the -DQB_SELFTEST twin)
    assert not any(hasattr(lib, n) for n in ("qb_selftest_planner", "qb_selftest_tile_emulation", "qb_selftest_bitins"))
    st = capi.selftest_lib()
    assert all(hasattr(st, n) for n in capi.prototypes(capi.SELFTEST_HEADER_PATH))


def test_header_prototypes_parse():
    protos = capi.prototypes()
    assert set(protos) == set(capi.declared_symbols())
    res, args = protos["qb_statevec_anyCtrlOneTargDenseMatr_subB"]
    assert res is C.c_int and args[-1] is capi.qb_cplx and args[-2] is capi.qb_cplx


def test_dropin_library_exports_reference_api():
    """the drop-in libQuEST.so must export the reference's public C API (a sample of each header)"""
    path = os.path.join(capi.REPO_ROOT, "quest_b200", "lib", "libQuEST.so")
    if not os.path.exists(path):
        pytest.skip("drop-in library not built (needs /root/reference at build time)")
    out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True).stdout
    for sym in ("createQureg", "applyCompMatr1", "applyCompMatr2", "applyCompMatr", "applyDiagMatr", "applyPauliGadget",
                "mixKrausMap", "mixDepolarising", "calcProbOfQubitOutcome", "calcExpecPauliStr", "calcExpecPauliStrSum",
                "applyFullQuantumFourierTransform", "applyMultiQubitProjector", "initQuESTEnv", "syncQuESTEnv"):
        assert f" T {sym}\n" in out, sym
    # and it must NOT carry the reference's GPU backend: every gpu_* symbol comes from our shim
    assert "N6thrust" not in out and "custatevec" not in out and "kernel_statevec" not in out


def test_bit_insertion_matches_reference_definition():
    """BitIns (<=4 scalar inserts, else the branch-free expand) == insertBitsWithMaskedValues (bitwise.hpp:206)"""
    lib = capi.selftest_lib()
    rng = np.random.default_rng(3)
    out = C.c_longlong()
    for _ in range(3000):
        n = int(rng.integers(0, 16))
        qs = [int(q) for q in rng.choice(48, size=n, replace=False)]
        st = [int(b) for b in rng.integers(0, 2, size=n)]
        item = int(rng.integers(0, 1 << 14))
        assert lib.qb_selftest_bitins(capi.ints(qs), capi.ints(st), n, item, C.byref(out)) == 0


def test_compute_fails_loudly_without_gpu():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    lib = capi.lib()
    assert lib.qb_num_devices() == 0 and lib.qb_is_device_available() == 0
    s = capi.qb_state()
    out = C.c_double()
    rc = lib.qb_statevec_calcTotalProb_sub(C.byref(s), C.byref(out))
    assert rc != 0 and b"no CPU fallback" in lib.qb_error_string()
    status = C.c_int(0)
    assert lib.qb_alloc(1024, C.byref(status)) is None and status.value != 0


@pytest.mark.parametrize("reorder", [1, 0], ids=["absorb+reorder", "program-order"])
@pytest.mark.parametrize("n", [4, 9, 13, 16])
def test_tile_planner_preserves_the_circuit(n, reorder):
    """host logic of the tile engine (quest_b200/csrc/qb_tile.cu: gate absorption, phase-star merging, Hadamard+star
    fusion, commuting first-fit grouping into passes and rounds): a random gate list applied in the planner's order to a
    small host state must equal gate-by-gate application in program order.  No GPU involved."""
    lib = capi.selftest_lib()
    err, passes, rounds, planned = C.c_double(), C.c_int(), C.c_int(), C.c_int()
    for seed in range(8):
        num_ops = 200
        rc = lib.qb_selftest_planner(n, num_ops, 7000 + seed, reorder, C.byref(err), C.byref(passes), C.byref(rounds), C.byref(planned))
        assert rc == 0, f"planner self-test failed structurally (rc={rc})"
        assert err.value <= 1e-12, f"n={n} seed={seed}: planned order changes the state by {err.value:.3e}"
        assert planned.value <= num_ops + 8
        if n >= 13:
            assert passes.value < planned.value       # gates really are grouped
        if n >= 13 and reorder:
            assert planned.value < num_ops            # absorption / merging shortened the list


def test_reference_example_links_against_the_dropin_library(tmp_path):
    """the drop-in boundary at the API level: the reference's own tutorial (examples/tutorials/min_example.c) compiles
    against QuEST's unchanged public headers, links against quest_b200/lib/libQuEST.so and runs.  Needs the reference tree
    (build container only); without a GPU QuEST's auto-deployer places the Qureg on its own, untouched CPU backend."""
    import shutil
    import subprocess
    ref = "/root/reference"
    lib_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "quest_b200", "lib")
    if not os.path.isdir(ref) or not os.path.exists(os.path.join(lib_dir, "libQuEST.so")) or not shutil.which("gcc"):
        pytest.skip("reference tree / drop-in library / gcc not available")
    # what CMake generates from quest/include/quest.h.in for this build's options (Makefile: SHIM_DEFS)
    (tmp_path / "quest.h").write_text(
        "#ifndef QUEST_H\n#define QUEST_H\n#define FLOAT_PRECISION 2\n#define COMPILE_MPI 1\n#define COMPILE_OPENMP 1\n"
        "#define COMPILE_CUDA 1\n#define COMPILE_CUQUANTUM 0\n" +
        "".join(f'#include "quest/include/{h}.h"\n' for h in
                ("version", "modes", "precision", "types", "calculations", "debug", "decoherence", "environment",
                 "initialisations", "channels", "operations", "paulis", "qureg", "matrices", "wrappers")) + "#endif\n")
    exe = tmp_path / "min_example"
    subprocess.run(["gcc", "-std=c11", f"-I{ref}", f"-I{tmp_path}", f"{ref}/examples/tutorials/min_example.c", "-o", str(exe),
                    f"-L{lib_dir}", "-lQuEST", f"-Wl,-rpath,{lib_dir}", "-lm"], check=True, capture_output=True, timeout=120)
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "Total probability" in out.stdout and "isGpuCompiled...........1" in out.stdout


@pytest.mark.parametrize("reorder", [1, 0], ids=["absorb+reorder", "program-order"])
@pytest.mark.parametrize("n", [12, 13, 15, 17])
def test_tile_kernel_emulation_on_host(n, reorder):
    """the tile engine end to end WITHOUT a GPU: planner -> emit_pass descriptors (tile enumeration, chunk offsets, round
    bits, dispatch codes, phase-star tables) -> the kernel's own round driver and gate bodies, which are compiled for the
    host as well as for the device (same source, quest_b200/csrc/qb_tile.cu) -> compared with plain gate-by-gate
    application.  Only the TMA / mbarrier tile pipeline itself is left to the GPU tests."""
    lib = capi.selftest_lib()
    err, tile_passes, direct_ops = C.c_double(), C.c_int(), C.c_int()
    total_tile_passes = 0
    for seed in range(6):
        rc = lib.qb_selftest_tile_emulation(n, 160, 8100 + seed, reorder, C.byref(err), C.byref(tile_passes), C.byref(direct_ops))
        assert rc == 0
        assert err.value <= 1e-12, f"n={n} seed={seed}: emulated tile passes differ from gate-by-gate application by {err.value:.3e}"
        total_tile_passes += tile_passes.value
    assert total_tile_passes > 0


def test_single_precision_twin_library():
    """FLOAT_PRECISION=1: the fp32 build of the kernel library exports the same ABI and reports its
    precision; one library per precision, as in the reference"""
    path = os.path.join(capi.REPO_ROOT, "quest_b200", "lib", "libquest_b200_f32.so")
    if not os.path.exists(path):
        pytest.skip("fp32 library not built (make kernels32)")
    lib32 = C.CDLL(path, mode=C.RTLD_LOCAL)
    assert lib32.qb_precision() == 1 and capi.lib().qb_precision() == 2
    missing = [n for n in capi.declared_symbols() if not hasattr(lib32, n)]
    assert not missing, missing


def test_single_precision_reference_pins_to_double_reference():
    """the fp32 oracle (the reference compiled at FLOAT_PRECISION=1) agrees with the committed fp64 golden outputs to
    single-precision accuracy -- so it really is the same algorithm at another precision"""
    from tests import helpers as H
    if not os.path.exists(os.path.join(capi.REPO_ROOT, "oracle", "_ref_f32", "libQuEST.so")):
        pytest.skip("oracle/_ref_f32 not built")
    fx = H.load_golden("configs_small.pkl")
    outs = H.run_programs("ref32", fx["programs"][:4])
    for got, want in zip(outs, fx["outputs"][:4]):
        H.assert_outputs_match(got, want, tol=2e-5, label="fp32 reference vs fp64 golden")


@pytest.mark.parametrize("n,bit", [(13, 12), (13, 6), (14, 7), (16, 15), (17, 9)])
def test_restricted_flush_on_host(n, bit):
    """the half-shard flush that the exchange / compute overlap relies on (qb_tile_flush_restricted): gates that never touch
    index bit `bit`, run first on the half with bit == 0 and then on the half with bit == 1 -- through the planner, emit_pass
    and the kernel's own round driver on the host -- must equal plain application; the other half's tiles must be pruned,
    and absorption / star merging must survive the restriction"""
    lib = capi.selftest_lib()
    err, passes, planned = C.c_double(), C.c_int(), C.c_int()
    # bits 0..5 belong to every tile and cannot be restricted by tile pruning: refused (the shim only overlaps on bits >= 10)
    assert lib.qb_selftest_restricted_flush(n, 10, 1, 3, C.byref(err), C.byref(passes), C.byref(planned)) == -5
    total_passes = 0
    for seed in range(5):
        rc = lib.qb_selftest_restricted_flush(n, 120, 9300 + seed, bit, C.byref(err), C.byref(passes), C.byref(planned))
        assert rc == 0, f"restricted flush self-test failed structurally (rc={rc})"
        assert err.value <= 1e-12, f"n={n} bit={bit} seed={seed}: halves differ from plain application by {err.value:.3e}"
        assert planned.value < 120 + 60          # (QFT-like stages expand to several queued gates; merging must shrink them again)
        total_passes += passes.value
    assert total_passes > 0


@pytest.mark.parametrize("n,restrict_bit", [(5, -1), (9, -1), (14, -1), (12, 11), (13, 6), (14, 0)])
def test_coset_pauli_kernel_on_host(n, restrict_bit):
    """the coset-blocked Pauli kernel without a GPU (quest_b200/csrc/qb_pauli_group.cu): the planner's grouping rule, the
    product's own descriptor builder (pivots of the masks' echelon form, coset offsets, sign bits) and the kernel's own
    per-thread body, compiled for the host, against pair-by-pair application of the definition -- on the whole state and
    restricted to each half of it (the exchange / compute overlap)"""
    lib = capi.selftest_lib()
    err, passes = C.c_double(), C.c_int()
    for seed in range(6):
        rc = lib.qb_selftest_pauli_group(n, 60, 9700 + seed, restrict_bit, C.byref(err), C.byref(passes))
        assert rc == 0, f"coset self-test failed structurally (rc={rc})"
        assert err.value <= 1e-12, f"n={n} restrict={restrict_bit} seed={seed}: coset passes differ from the definition by {err.value:.3e}"
        assert 0 < passes.value < 60 * (2 if restrict_bit >= 0 else 1)
