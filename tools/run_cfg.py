"""Runs one BASELINE configuration once through the public API on the drop-in library (for ncu launch lists / profiles):
   python tools/run_cfg.py cfg4|cfg5|cfg2 [qubits]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quest_b200 import quest_api as qa
from quest_b200.program import run_program
from tests import programs as P

which = sys.argv[1]
Q = qa.QuEST(qa.B200_LIB); Q.initCustomQuESTEnv(0, 1, 0)
if which == "cfg4": prog = P.cfg4_program(int(sys.argv[2]) if len(sys.argv) > 2 else 14, 14014, layers=10, dump=False)
elif which == "cfg5": prog = P.cfg5_program(int(sys.argv[2]) if len(sys.argv) > 2 else 28, 28200, num_terms=200, dump=False)
else: prog = P.cfg2_program(int(sys.argv[2]) if len(sys.argv) > 2 else 30, 20302, 200, dump=False)
for spec in prog["quregs"].values(): spec["custom"] = [0, 1, 0]
for rep in range(2):
    Q.syncQuESTEnv(); t0 = time.perf_counter()
    out = run_program(Q, prog)
    Q.syncQuESTEnv(); print(which, "rep", rep, "seconds", time.perf_counter() - t0, out["results"][-2:], flush=True)
