"""Generates tests/golden/fullsize_*.pkl: the BASELINE configurations at FULL size (30-qubit cfg 2, 14-qubit density
matrix cfg 4, 28-qubit cfg 5; programs defined in tests/test_fullsize_gpu.py) run on the UNMODIFIED reference library
oracle/_ref/libQuEST.so.  Only scalars and amplitude windows are stored (a few MiB), never the 16 GiB states.
Run in the build container (needs ~20 GiB of RAM and several minutes of CPU):   python tests/golden/make_fullsize.py
"""
import os
import pickle
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from tests.helpers import run_programs          # noqa: E402
from tests.test_fullsize_gpu import FULLSIZE    # noqa: E402

if __name__ == "__main__":
    only = set(sys.argv[1:])
    for fname, make in FULLSIZE.items():
        if only and fname not in only:
            continue
        prog = make()
        out = run_programs("ref", [prog], timeout=7200, env={"OMP_NUM_THREADS": str(os.cpu_count() or 8)})
        path = os.path.join(ROOT, "tests", "golden", fname)
        pickle.dump({"programs": [prog], "outputs": out}, open(path, "wb"), protocol=4)
        print(fname, os.path.getsize(path), "bytes", flush=True)
