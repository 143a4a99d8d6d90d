"""GPU probe: where do the compute warpgroups of k_tile_pass spend their cycles?  Needs the -DQB_TILE_TIMING build
(`make timing` -> build/timing/libquest_b200_timing.so).  Prints, per scenario, the
time per launch and the share of warpgroup cycles per phase."""
import ctypes as C, sys, os, math
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quest_b200 import capi
capi.LIB_PATH = os.environ.get("QB_TIMING_LIB", os.path.join(ROOT, "build", "timing", "libquest_b200_timing.so"))
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
amps = torch.empty(1 << n, dtype=torch.complex128, device="cuda")
s = capi.state(amps, n); ref = C.byref(s)
capi.call("qb_statevec_initUniformState_sub", ref, capi.cplx(2.0 ** (-n / 2)))
lib = capi.lib()
has_timing = hasattr(lib, "qb_tile_timing_read")
E = capi.ints([])
rng = np.random.default_rng(0)
u4c = capi.cplx_array(np.linalg.qr(rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4)))[0])
h = capi.cplx_array(np.array([[1, 1], [1, -1]]) / math.sqrt(2))
SLOTS = ["prologue", "wait_done", "wait_full(TMA)", "wg_sync@start", "lds_issue", "gates(+lds wait)", "sts", "smem_pauli", "wg_sync@round", "fence+arrive", "-", "-"]

def issue(gates):
    for g in gates:
        if g[0] == "h": lib.qb_statevec_anyCtrlOneTargDenseMatr_subA(ref, E, E, 0, g[1], h)
        elif g[0] == "m2": lib.qb_statevec_anyCtrlTwoTargDenseMatr_sub(ref, E, E, 0, g[1], g[2], g[3] if len(g) > 3 else u4c)
        elif g[0] == "m1": lib.qb_statevec_anyCtrlOneTargDenseMatr_subA(ref, E, E, 0, g[1], g[2])
    capi.call("qb_flush")

def run(label, gates, mode, reps=3):
    capi.call("qb_set_tile_engine", mode)
    issue(gates); capi.sync()
    buf = (C.c_ulonglong * 12)()
    if has_timing: lib.qb_tile_timing_read(buf, 1)
    l0 = lib.qb_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): issue(gates)
    e1.record(); torch.cuda.synchronize()
    launches = (lib.qb_launch_count() - l0) / reps
    ms = e0.elapsed_time(e1) / reps
    line = f"{label:44s} gates={len(gates):3d} launches={launches:5.1f} ms={ms:8.3f} ms/launch={ms/launches:7.3f}"
    if has_timing:
        lib.qb_tile_timing_read(buf, 1)
        tot = sum(buf) or 1
        line += "\n      " + "  ".join(f"{nm}={100*v/tot:.1f}%" for nm, v in zip(SLOTS, buf) if v)
        line += f"\n      cycles per tile per warpgroup = {tot / reps / (launches * (1 << (n - 12)))  :.0f}"
    print(line, flush=True)

hb = [26, 27, 28, 29]
run("2x m2 (26,27),(28,29) (1 round)", [("m2", 26, 27), ("m2", 28, 29)], 2)
run("8x m2 same x4 (1 round)", [("m2", 26, 27), ("m2", 28, 29)] * 4, 2)
run("16xH on q26..29 x4 (1 round)", [("h", q) for q in hb * 4], 2)
run("12xH: q26..29, q6..9, q26..29 (3 rounds)", [("h", q) for q in hb + [6, 7, 8, 9] + hb], 2)
run("8xH: q26..29 then q0..3 (8-way conflicts)", [("h", q) for q in hb + [0, 1, 2, 3]], 2)
dense = []
for op in bench.dense_stream(n):
    if op[0] == "m1": dense.append(("m1", op[1], capi.cplx_array(op[2])))
    else: dense.append(("m2", op[1], op[2], capi.cplx_array(op[3])))
run("cfg2 dense section (200 gates, planner on)", dense, 1)
