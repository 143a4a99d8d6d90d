"""ctypes binding of the quest_b200 C ABI (include/quest_b200.h -> quest_b200/lib/libquest_b200.so).

This is the kernel-level entry used by the `-m gpu` parity tests and by bench.py's device-resident
timing: amplitudes live in torch CUDA tensors (torch is plumbing for device memory and streams only),
every compute call goes straight into the hand-written CUDA library.  There is no fallback: if the
library is missing, or no GPU is present, calls raise.
"""
import ctypes as C
import os
import re

REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(REPO_ROOT, "quest_b200", "lib", "libquest_b200.so")
HEADER_PATH = os.path.join(REPO_ROOT, "include", "quest_b200.h")
# test-only twin of the library carrying the host-side self-tests (include/quest_b200_selftest.h); never used by the product
SELFTEST_LIB_PATH = os.path.join(REPO_ROOT, "quest_b200", "lib", "libquest_b200_selftest.so")
SELFTEST_HEADER_PATH = os.path.join(REPO_ROOT, "include", "quest_b200_selftest.h")


class qb_cplx(C.Structure):
    _fields_ = [("re", C.c_double), ("im", C.c_double)]


class qb_state(C.Structure):
    _fields_ = [("amps", C.c_void_p), ("buffer", C.c_void_p), ("numAmpsPerNode", C.c_longlong),
                ("logNumAmpsPerNode", C.c_int), ("rank", C.c_int), ("numQubits", C.c_int),
                ("logNumColsPerNode", C.c_int), ("isDensityMatrix", C.c_int)]


class QbError(RuntimeError):
    pass


def declared_symbols(header=HEADER_PATH):
    """Every function name declared in include/quest_b200.h (used by the CPU-side export test)."""
    text = open(header).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qb_[A-Za-z0-9_]+)\s*\(", text)))


_lib = None

_SCALARS = {"int": C.c_int, "double": C.c_double, "qb_index": C.c_longlong, "size_t": C.c_size_t,
            "unsigned": C.c_uint, "unsigned long long": C.c_ulonglong, "qb_cplx": qb_cplx, "void": None}


def _ctype(decl):
    """Map one C parameter / return declaration of quest_b200.h to a ctypes type (pointers -> void*)."""
    decl = decl.strip()
    if "*" in decl or "[" in decl:
        return C.c_char_p if decl.startswith("const char*") and "[" not in decl else C.c_void_p
    decl = re.sub(r"\bconst\b", "", decl).strip()
    for name in sorted(_SCALARS, key=len, reverse=True):     # type is everything but the parameter name
        if decl == name or decl.startswith(name + " "):
            return _SCALARS[name]
    raise QbError(f"quest_b200.h: cannot map C declaration '{decl}'")


def prototypes(header=HEADER_PATH):
    """{name: (restype, [argtypes])} parsed from the header, so python can never drift from the ABI."""
    text = re.sub(r"/\*.*?\*/", "", open(header).read(), flags=re.S)
    text = re.sub(r"^\s*#.*$", "", text, flags=re.M)
    out = {}
    for m in re.finditer(r"([A-Za-z_][A-Za-z0-9_ \*]*?)\b(qb_[A-Za-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        ret, name, params = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
        args = [] if params in ("", "void") else [_ctype(p) for p in params.split(",")]
        out[name] = (_ctype(ret + " ") if "*" not in ret else (C.c_char_p if "char" in ret else C.c_void_p), args)
    return out


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise QbError(f"{LIB_PATH} not built: run `make kernels` / __graft_entry__.build(); there is no fallback path")
        _lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in prototypes().items():
            fn = getattr(_lib, name)
            fn.restype, fn.argtypes = res, args
    return _lib


_selftest_lib = None


def selftest_lib():
    """the -DQB_SELFTEST build of the library (host-only planner / index-algebra self-tests for the CPU test-suite)"""
    global _selftest_lib
    if _selftest_lib is None:
        if not os.path.exists(SELFTEST_LIB_PATH):
            raise QbError(f"{SELFTEST_LIB_PATH} not built: run `make selftest`")
        _selftest_lib = C.CDLL(SELFTEST_LIB_PATH, mode=C.RTLD_LOCAL)
        for name, (res, args) in prototypes(SELFTEST_HEADER_PATH).items():
            fn = getattr(_selftest_lib, name)
            fn.restype, fn.argtypes = res, args
    return _selftest_lib


def check(status, what=""):
    if status != 0:
        raise QbError(f"{what}: status {status}: {lib().qb_error_string().decode()}")


def cplx(z):
    z = complex(z)
    return qb_cplx(z.real, z.imag)


def cplx_array(values):
    import numpy as np
    v = np.ascontiguousarray(np.asarray(values, dtype=np.complex128).reshape(-1))
    arr = (qb_cplx * v.size)()
    C.memmove(arr, v.ctypes.data, v.nbytes)
    return arr


def ints(seq):
    seq = [int(v) for v in seq]
    return (C.c_int * max(1, len(seq)))(*seq)


def state(amps, numQubits, isDensityMatrix=0, rank=0, logNumNodes=0, buffer=None):
    """qb_state over a torch complex128 CUDA tensor (and an optional same-sized buffer tensor)."""
    n = amps.numel()
    log_n = n.bit_length() - 1
    assert (1 << log_n) == n, "amplitude count must be a power of two"
    s = qb_state()
    s.amps = amps.data_ptr()
    s.buffer = buffer.data_ptr() if buffer is not None else None
    s.numAmpsPerNode = n
    s.logNumAmpsPerNode = log_n
    s.rank = rank
    s.numQubits = numQubits
    s.logNumColsPerNode = (numQubits - logNumNodes) if isDensityMatrix else 0
    s.isDensityMatrix = isDensityMatrix
    return s


def call(name, *args):
    """Call a status-returning qb_* entry point and raise on failure."""
    fn = getattr(lib(), name)
    check(fn(*args), name)


def sync():
    call("qb_sync")
