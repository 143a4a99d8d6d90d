"""GPU micro-probe of the tile engine: time fused passes for chosen target sets (chunk sizes / op counts)."""
import ctypes as C, sys, os, math
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quest_b200 import capi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
amps = torch.empty(1 << n, dtype=torch.complex128, device="cuda")
s = capi.state(amps, n); ref = C.byref(s)
capi.call("qb_statevec_initUniformState_sub", ref, capi.cplx(2.0 ** (-n / 2)))
h = capi.cplx_array(np.array([[1, 1], [1, -1]]) / math.sqrt(2))
rng = np.random.default_rng(0)
u4 = np.linalg.qr(rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4)))[0]
u4c = capi.cplx_array(u4)
E = capi.ints([])
lib = capi.lib()

ONLY = os.environ.get("PROBE_ONLY", "")

def run(label, gates, reps=5):
    if ONLY and not any(k in label for k in ONLY.split("|")):
        return
    def issue():
        for g in gates:
            if g[0] == "h": lib.qb_statevec_anyCtrlOneTargDenseMatr_subA(ref, E, E, 0, g[1], h)
            elif g[0] == "m2": lib.qb_statevec_anyCtrlTwoTargDenseMatr_sub(ref, E, E, 0, g[1], g[2], u4c)
            elif g[0] == "cp": lib.qb_statevec_anyCtrlOneTargDiagMatr_sub(ref, capi.ints([g[2]]), capi.ints([1]), 1, g[1], capi.cplx_array([1, np.exp(0.3j)]))
        capi.call("qb_flush")
    issue(); capi.sync()
    l0 = lib.qb_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): issue()
    e1.record(); torch.cuda.synchronize()
    launches = (lib.qb_launch_count() - l0) / reps
    ms = e0.elapsed_time(e1) / reps
    print(f"{label:58s} gates={len(gates):3d} launches={launches:5.1f} ms={ms:8.3f} ms/launch={ms/launches:7.3f}", flush=True)

capi.call("qb_set_tile_engine", 2)      # program order, no gate absorption: the probe repeats gates on purpose
hb = [26, 27, 28, 29]
run("4xH on q26..29 (1 round, 16 chunks)", [("h", q) for q in hb])
run("8xH on q26..29 x2 (1 round)", [("h", q) for q in hb * 2])
run("16xH on q26..29 x4 (1 round)", [("h", q) for q in hb * 4])
run("32xH on q26..29 x8 (1 round)", [("h", q) for q in hb * 8])
run("8xH: q26..29 then q6..9 (2 rounds, no conflicts)", [("h", q) for q in hb + [6, 7, 8, 9]])
run("12xH: q26..29, q6..9, q26..29 (3 rounds)", [("h", q) for q in hb + [6, 7, 8, 9] + hb])
run("8xH: q26..29 then q2..5 (2 rounds, 2-way conflicts)", [("h", q) for q in hb + [2, 3, 4, 5]])
run("8xH: q26..29 then q0..3 (2 rounds, 8-way conflicts)", [("h", q) for q in hb + [0, 1, 2, 3]])
run("2x m2 (26,27),(28,29) (1 round)", [("m2", 26, 27), ("m2", 28, 29)])
run("4x m2 same x2 (1 round)", [("m2", 26, 27), ("m2", 28, 29)] * 2)
run("8x m2 same x4 (1 round)", [("m2", 26, 27), ("m2", 28, 29)] * 4)
run("6xH: q26..29 then q6,q7 (2 rounds, no conflicts)", [("h", q) for q in hb + [6, 7]])
run("10xH: q26..29, q6,q7, q26..29 (3 rounds, no conflicts)", [("h", q) for q in hb + [6, 7] + hb])
run("6xH: q26..29 then q4,q5 (2 rounds)", [("h", q) for q in hb + [4, 5]])
run("6xH on q24..q29 (2 rounds, 64 chunks of 1 KiB)", [("h", q) for q in range(24, 30)])
run("QFT block: H29 + 29 cphase + H28 + 28 cphase", [("h", 29)] + [("cp", 29, c) for c in range(29)] + [("h", 28)] + [("cp", 28, c) for c in range(28)])
capi.call("qb_set_tile_engine", 0)
run("direct: 1xH q29", [("h", 29)])
run("direct: 1xH q0", [("h", 0)])
run("direct: 1x m2 (24,25)", [("m2", 24, 25)])
capi.call("qb_set_tile_engine", 2)
run("dbg: H8,H9 (contiguous tile)", [("h", 8), ("h", 9)])
run("dbg: H6,H7 (contiguous tile)", [("h", 6), ("h", 7)])
run("dbg: H0,H1 (contiguous tile)", [("h", 0), ("h", 1)])
run("dbg: H12,H13", [("h", 12), ("h", 13)])
