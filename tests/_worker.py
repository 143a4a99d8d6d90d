"""Worker process: runs a list of programs (quest_b200/program.py) on ONE QuEST library and pickles the
outputs.  usage: python tests/_worker.py <which: b200|ref> <programs.pkl> <out.pkl>
A separate process per library is required because both export the same symbols and QuEST keeps
process-global singletons (api/environment.cpp:49)."""
import os
import pickle
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

# `which` ending in "32" selects the single-precision twins (libQuEST_f32.so / oracle/_ref_f32): precision is an
# import-time property of the binding, like FLOAT_PRECISION is a build-time property of a QuEST library
if len(sys.argv) > 1 and sys.argv[1].endswith("32"):
    os.environ["QUEST_PRECISION"] = "1"
    sys.argv[1] = sys.argv[1][:-2]

from quest_b200 import quest_api as qa          # noqa: E402
from quest_b200.program import run_program      # noqa: E402


CORE_LIB = "libquest_b200_f32.so" if qa.PRECISION == 1 else "libquest_b200.so"


def main():
    which, src, dst = sys.argv[1:4]
    progs = pickle.load(open(src, "rb"))
    if which == "ref":
        Q = qa.QuEST(qa.REF_LIB)                 # the checker (oracle/_ref or oracle/_ref_f32): test infrastructure only
        Q.initCustomQuESTEnv(0, 0, 1)            # reference: CPU + OpenMP, the parity oracle
    elif which == "b200dist":
        # one process per GPU (RANK / WORLD_SIZE / LOCAL_RANK from the launcher); every Qureg is sharded
        Q = qa.QuEST(qa.B200_LIB)
        if os.environ.get("QUEST_B200_P2P", "1") == "0":
            import ctypes
            ctypes.CDLL(os.path.join(os.path.dirname(qa.B200_LIB), CORE_LIB), mode=ctypes.RTLD_GLOBAL).qb_p2p_set_enabled(0)
        Q.initCustomQuESTEnv(1, 1, 0)
        dst = dst + "." + os.environ.get("RANK", "0")
    else:
        Q = qa.QuEST(qa.B200_LIB)
        Q.initCustomQuESTEnv(0, 1, 0)            # product: GPU only; fails loudly without a device
    if qa.PRECISION == 1:
        # the programs' operators are generated in double and narrowed to float: unitarity / CPTP hold to ~1e-7 per
        # element, which the reference's own fp32 validation threshold (summed over large matrices) can reject
        Q.lib.setValidationOff()
    outs = []
    for p in progs:
        if which != "ref":
            # small Quregs are auto-deployed to the CPU (core/autodeployer.hpp:17-21); the product under test
            # is the GPU backend, so force GPU acceleration exactly as createCustomQureg allows a user to
            for spec in p["quregs"].values():
                spec.setdefault("custom", [1 if which == "b200dist" else 0, 1, 0])
        out = run_program(Q, p)
        if which != "ref":
            for name, inf in out["info"].items():
                assert inf["isGpuAccelerated"] == 1, f"qureg {name} is not GPU-accelerated: refusing CPU path"
        if which == "b200dist":
            import ctypes
            core = ctypes.CDLL(os.path.join(os.path.dirname(qa.B200_LIB), CORE_LIB), mode=ctypes.RTLD_GLOBAL)
            out["p2p_available"] = int(core.qb_p2p_is_available())
            out["transport"] = int(core.qb_comm_transport())
            core.qb_p2p_overlapped_count.restype = ctypes.c_ulonglong
            out["overlapped_swaps"] = int(core.qb_p2p_overlapped_count(1))
            nex = ctypes.c_ulonglong(0)
            core.qb_p2p_stats(ctypes.byref(nex), None)
            out["p2p_exchanges"] = int(nex.value)          # cumulative over the programs this worker has run
        outs.append(out)
    Q.finalizeQuESTEnv()
    pickle.dump(outs, open(dst, "wb"))


if __name__ == "__main__":
    main()
