#!/usr/bin/env python
"""One-shot GPU check of the look-ahead gate log (QUEST_B200_LOOKAHEAD): the programs of
tests/test_dist_gpu.py::test_look_ahead_victim_choice_sharded on `world` ranks (sharing GPUs if need be), against the
reference CPU library, printing each stage as it completes.   python tools/lookahead_gpu_check.py [world ...]"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tests import helpers as H, programs as P   # noqa: E402


def main():
    worlds = [int(x) for x in sys.argv[1:]] or [4]
    window = os.environ.get("LOOKAHEAD_WINDOW", "8")
    t0 = time.time()
    for world in worlds:
        logp = world.bit_length() - 1
        progs = P.lookahead_programs(logp)
        want = H.run_programs("ref", progs)
        got = H.run_programs_distributed(progs, world, env={"QUEST_B200_LOOKAHEAD": window})
        for k, (g, w) in enumerate(zip(got, want)):
            H.assert_outputs_match(g, w, label=f"P={world} prog[{k}]",
                                   int_exact_ops={i for i, op in enumerate(progs[k]["ops"]) if "Measurement" in op[0] and "Forced" not in op[0]})
        ex = [got[k]["p2p_exchanges"] - (got[k - 1]["p2p_exchanges"] if k else 0) for k in range(len(got))]
        print(f"[{time.time() - t0:5.1f}s] world {world} window {window}: PARITY OK, exchanges per program {ex}", flush=True)
        base = H.run_programs_distributed(progs, world)
        exb = [base[k]["p2p_exchanges"] - (base[k - 1]["p2p_exchanges"] if k else 0) for k in range(len(base))]
        print(f"[{time.time() - t0:5.1f}s] world {world} default rule:            exchanges per program {exb}", flush=True)


if __name__ == "__main__":
    main()
