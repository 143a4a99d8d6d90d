"""Generates tests/golden/*.pkl by running the committed program generators (tests/programs.py) on the
UNMODIFIED reference library oracle/_ref/libQuEST.so (built from /root/reference by oracle/Makefile).
Run in the build container only:   python tests/golden/make_golden.py
Each fixture is {"programs": [...], "outputs": [...]} -- the exact inputs and what the reference returned.
"""
import os
import pickle
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from tests import programs as P          # noqa: E402
from tests.helpers import run_programs   # noqa: E402

FIXTURES = {
    "gates_sv.pkl": [P.gates_program(6, 101), P.gates_program(5, 102, max_ctrls=3), P.gates_program(3, 103, max_ctrls=1)],
    "gates_dm.pkl": [P.gates_program(3, 201, dm=1, max_ctrls=1), P.gates_program(4, 202, dm=1, num_rounds=1)],
    "calcs_sv.pkl": [P.calcs_program_sv(6, 301), P.calcs_program_sv(4, 302)],
    "channels_dm.pkl": [P.channels_program_dm(4, 401), P.channels_program_dm(3, 402)],
    "dense_big.pkl": [P.big_dense_program(8, 501, 6), P.big_dense_program(7, 503, 4), P.big_dense_program(6, 504, 5, nc=0)],
    "configs_small.pkl": [P.cfg1_program(10, num_gates=60), P.cfg2_program(9, num_gates=40), P.cfg4_program(5, layers=2),
                          P.cfg5_program(8, num_terms=12), P.measurement_program(7, 77)],
    "relabel_sv.pkl": [P.relabel_program(6, 601), P.relabel_program(8, 602, num_ops=120), P.relabel_program(7, 603, num_ops=200, reads=False)],
}

if __name__ == "__main__":
    only = set(sys.argv[1:])          # optional: regenerate just the named fixtures
    for fname, progs in FIXTURES.items():
        if only and fname not in only:
            continue
        outs = run_programs("ref", progs)
        path = os.path.join(ROOT, "tests", "golden", fname)
        pickle.dump({"programs": progs, "outputs": outs}, open(path, "wb"), protocol=4)
        print(fname, os.path.getsize(path), "bytes")
