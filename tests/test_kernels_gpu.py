"""GPU parity tests proper: every C-ABI entry point (include/quest_b200.h) against the numpy oracle
(oracle/quest_oracle.py, pinned to the reference by tests/test_oracle_golden.py) on identical seeded inputs.

Amplitudes live in torch CUDA tensors (plumbing only); every compute call goes through ctypes into
quest_b200/lib/libquest_b200.so.  Tolerance: 1e-12 relative L2 (fp64 north-star bound); integer / index
results bit-exact.  Each case runs with the TMA tile engine enabled AND disabled (direct kernels)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import quest_oracle as qo          # noqa: E402
from quest_b200 import capi                    # noqa: E402
from tests.helpers import rel_l2, TOL          # noqa: E402
from tests.programs import rand_unitary, rand_state, rand_density   # noqa: E402

torch = pytest.importorskip("torch")


class Dev:
    """a State mirrored on the device"""

    def __init__(self, st):
        self.st = st
        self.amps = torch.from_numpy(st.amps.copy()).cuda()
        self.buf = torch.from_numpy(st.buffer.copy()).cuda() if st.buffer is not None else None
        torch.cuda.synchronize()
        self.c = capi.state(self.amps, st.numQubits, st.isDensityMatrix, st.rank, st.logNumNodes, self.buf)
        self.ref = C.byref(self.c)

    def host(self):
        capi.sync()
        return self.amps.cpu().numpy()

    def host_buf(self):
        capi.sync()
        return self.buf.cpu().numpy()


def rand_sv(rng, n, rank=0, logNodes=0, buffer=False):
    st = qo.new_state(n, 0, rank, logNodes, with_buffer=buffer or logNodes > 0)
    st.amps[:] = rand_state(rng, st.logNumAmpsPerNode)
    if st.buffer is not None:
        st.buffer[:] = rand_state(rng, st.logNumAmpsPerNode)
    return st


def rand_dm(rng, n, rank=0, logNodes=0):
    st = qo.new_state(n, 1, rank, logNodes, with_buffer=logNodes > 0)
    st.amps[:] = rand_state(rng, st.logNumAmpsPerNode) * 3
    if st.buffer is not None:
        st.buffer[:] = rand_state(rng, st.logNumAmpsPerNode)
    return st


def pick(rng, n, k):
    return [int(q) for q in rng.choice(n, size=k, replace=False)]


def ctrl_targ(rng, n, nc, nt):
    qs = pick(rng, n, nc + nt)
    return qs[:nc], [int(b) for b in rng.integers(0, 2, size=nc)], qs[nc:]


def check_state(dev, st, label):
    err = rel_l2(dev.host(), st.amps)
    assert err <= TOL, f"{label}: rel-L2 {err:.3e}"


def dev_matrix(m):
    t = torch.from_numpy(np.ascontiguousarray(m, dtype=np.complex128).reshape(-1)).cuda()
    torch.cuda.synchronize()
    return t


@pytest.fixture(params=[1, 0], ids=["tile", "direct"])
def engine(request):
    capi.call("qb_set_tile_engine", request.param)
    yield request.param
    capi.call("qb_set_tile_engine", 1)


SIZES = [1, 2, 5, 10, 13, 16, 21]


@pytest.mark.parametrize("n", SIZES)
def test_dense1(n, engine):
    rng = np.random.default_rng(100 + n)
    for nc in range(0, min(4, n)):
        for trial in range(3):
            c, s, t = ctrl_targ(rng, n, nc, 1)
            if trial == 0 and nc == 0:
                t = [0]
            if trial == 1 and nc == 0:
                t = [n - 1]
            m = rand_unitary(rng, 2)
            st = rand_sv(rng, n); dev = Dev(st)
            capi.call("qb_statevec_anyCtrlOneTargDenseMatr_subA", dev.ref, capi.ints(c), capi.ints(s), nc, t[0], capi.cplx_array(m))
            qo.statevec_anyCtrlOneTargDenseMatr_subA(st, c, s, t[0], m)
            check_state(dev, st, f"dense1 n={n} c={c} s={s} t={t}")


@pytest.mark.parametrize("n", [2, 5, 10, 13, 16, 21])
def test_dense2(n, engine):
    rng = np.random.default_rng(200 + n)
    for nc in range(0, min(3, n - 1)):
        for trial in range(4):
            c, s, t = ctrl_targ(rng, n, nc, 2)
            if nc == 0 and trial == 0:
                t = [0, n - 1]
            if nc == 0 and trial == 1:
                t = [1, 0]
            m = rand_unitary(rng, 4)
            st = rand_sv(rng, n); dev = Dev(st)
            capi.call("qb_statevec_anyCtrlTwoTargDenseMatr_sub", dev.ref, capi.ints(c), capi.ints(s), nc, t[0], t[1], capi.cplx_array(m))
            qo.statevec_anyCtrlTwoTargDenseMatr_sub(st, c, s, t[0], t[1], m)
            check_state(dev, st, f"dense2 n={n} c={c} s={s} t={t}")


@pytest.mark.parametrize("n,nt", [(3, 1), (4, 2), (5, 3), (8, 4), (12, 4), (9, 5), (14, 5), (8, 6), (13, 6), (9, 7), (10, 8), (18, 3), (20, 4)])
def test_denseK(n, nt, engine):
    rng = np.random.default_rng(300 + n * 10 + nt)
    for nc in range(0, min(3, n - nt + 1)):
        for conj in (0, 1):
            c, s, t = ctrl_targ(rng, n, nc, nt)
            m = rand_unitary(rng, 1 << nt)
            dm = dev_matrix(m)
            st = rand_sv(rng, n); dev = Dev(st)
            capi.call("qb_statevec_anyCtrlAnyTargDenseMatr_sub", dev.ref, capi.ints(c), capi.ints(s), nc, capi.ints(t), nt, dm.data_ptr(), conj)
            qo.statevec_anyCtrlAnyTargDenseMatr_sub(st, c, s, t, m, bool(conj))
            check_state(dev, st, f"denseK n={n} nt={nt} c={c} s={s} t={t} conj={conj}")


@pytest.mark.parametrize("n", [1, 3, 6, 12, 17, 21])
@pytest.mark.parametrize("rank,logNodes", [(0, 0), (5, 3)])
def test_diag(n, rank, logNodes, engine):
    """diagonal targets may be PREFIX qubits (bits of the rank): qubits range over n + logNodes"""
    rng = np.random.default_rng(400 + n)
    nglob = n + logNodes
    for nc in range(0, min(3, n)):
        c = pick(rng, n, nc); s = [int(b) for b in rng.integers(0, 2, size=nc)]
        free = [q for q in range(nglob) if q not in c]
        # one target
        t = int(rng.choice(free)); e = np.exp(1j * rng.uniform(0, 6, 2))
        st = rand_sv(rng, nglob, rank, logNodes); dev = Dev(st)
        capi.call("qb_statevec_anyCtrlOneTargDiagMatr_sub", dev.ref, capi.ints(c), capi.ints(s), nc, t, capi.cplx_array(e))
        qo.statevec_anyCtrlOneTargDiagMatr_sub(st, c, s, t, e)
        check_state(dev, st, f"diag1 n={n} c={c} t={t}")
        if len(free) >= 2:
            t2 = [int(q) for q in rng.choice(free, size=2, replace=False)]; e = np.exp(1j * rng.uniform(0, 6, 4))
            st = rand_sv(rng, nglob, rank, logNodes); dev = Dev(st)
            capi.call("qb_statevec_anyCtrlTwoTargDiagMatr_sub", dev.ref, capi.ints(c), capi.ints(s), nc, t2[0], t2[1], capi.cplx_array(e))
            qo.statevec_anyCtrlTwoTargDiagMatr_sub(st, c, s, t2[0], t2[1], e)
            check_state(dev, st, f"diag2 n={n} c={c} t={t2}")
        for nt in (1, 2, 3, 5):
            if nt > len(free):
                continue
            tk = [int(q) for q in rng.choice(free, size=nt, replace=False)]
            e = np.exp(1j * rng.uniform(0, 6, 1 << nt)) * rng.uniform(0.5, 1.5, 1 << nt)
            de = dev_matrix(e)
            for conj, hasPow, expo in ((0, 0, 1), (1, 0, 1), (0, 1, 0.5 - 0.3j), (1, 1, 2.0)):
                st = rand_sv(rng, nglob, rank, logNodes); dev = Dev(st)
                capi.call("qb_statevec_anyCtrlAnyTargDiagMatr_sub", dev.ref, capi.ints(c), capi.ints(s), nc, capi.ints(tk), nt,
                          de.data_ptr(), conj, hasPow, capi.cplx(expo))
                qo.statevec_anyCtrlAnyTargDiagMatr_sub(st, c, s, tk, e, bool(conj), bool(hasPow), expo)
                check_state(dev, st, f"diagK n={n} nt={nt} conj={conj} pow={hasPow}")


@pytest.mark.parametrize("n", [1, 2, 6, 11, 16, 21])
def test_pauli_and_phase(n, engine):
    rng = np.random.default_rng(500 + n)
    for nc in range(0, min(3, n)):
        for k in sorted({1, min(2, n - nc), min(5, n - nc), n - nc}):
            if k < 1:
                continue
            c, s, t = ctrl_targ(rng, n, nc, k)
            chars = rng.choice(list("XYZ"), size=k)
            if not any(ch in "XY" for ch in chars):
                chars[0] = "X"
            x = [q for ch, q in zip(chars, t) if ch == "X"]; y = [q for ch, q in zip(chars, t) if ch == "Y"]
            z = [q for ch, q in zip(chars, t) if ch == "Z"]
            af, pf = complex(rng.normal(), rng.normal()), complex(rng.normal(), rng.normal())
            st = rand_sv(rng, n); dev = Dev(st)
            capi.call("qb_statevector_anyCtrlPauliTensorOrGadget_subA", dev.ref, capi.ints(c), capi.ints(s), nc,
                      capi.ints(x), len(x), capi.ints(y), len(y), capi.ints(z), len(z), capi.cplx(af), capi.cplx(pf))
            qo.statevector_anyCtrlPauliTensorOrGadget_subA(st, c, s, x, y, z, af, pf)
            check_state(dev, st, f"pauliA n={n} c={c} x={x} y={y} z={z}")
            # phase gadget on the same targets
            f0, f1 = np.exp(1j * rng.uniform(0, 6, 2))
            st = rand_sv(rng, n); dev = Dev(st)
            capi.call("qb_statevector_anyCtrlAnyTargZOrPhaseGadget_sub", dev.ref, capi.ints(c), capi.ints(s), nc, capi.ints(t), k,
                      capi.cplx(f0), capi.cplx(f1))
            qo.statevector_anyCtrlAnyTargZOrPhaseGadget_sub(st, c, s, t, f0, f1)
            check_state(dev, st, f"phase n={n} c={c} t={t}")


@pytest.mark.parametrize("n", [2, 5, 12, 16, 21])
def test_swap(n, engine):
    rng = np.random.default_rng(600 + n)
    for nc in range(0, min(3, n - 1)):
        for _ in range(3):
            c, s, t = ctrl_targ(rng, n, nc, 2)
            st = rand_sv(rng, n); dev = Dev(st)
            capi.call("qb_statevec_anyCtrlSwap_subA", dev.ref, capi.ints(c), capi.ints(s), nc, t[0], t[1])
            qo.statevec_anyCtrlSwap_subA(st, c, s, t[0], t[1])
            err = np.max(np.abs(dev.host() - st.amps))
            assert err == 0.0, f"swap must be bit-exact, n={n} c={c} t={t}"


@pytest.mark.parametrize("n", [3, 8, 14])
def test_buffer_kernels(n):
    """the post-exchange kernels: pack, swap subB/subC, dense subB, Pauli subB (buffer contents are arbitrary)"""
    rng = np.random.default_rng(700 + n)
    for nc in range(0, min(3, n - 1)):
        c, s, t = ctrl_targ(rng, n, nc, 1)
        # pack
        if nc > 0:
            st = rand_sv(rng, n, buffer=True); dev = Dev(st)
            npacked = C.c_longlong()
            capi.call("qb_statevec_packAmpsIntoBuffer", dev.ref, capi.ints(c), capi.ints(s), nc, C.byref(npacked))
            want = qo.statevec_packAmpsIntoBuffer(st, c, s)
            assert npacked.value == want
            assert np.array_equal(dev.host_buf(), st.buffer), "pack must be bit-exact"
        # swap subB / subC
        st = rand_sv(rng, n, buffer=True); dev = Dev(st)
        capi.call("qb_statevec_anyCtrlSwap_subB", dev.ref, capi.ints(c), capi.ints(s), nc)
        qo.statevec_anyCtrlSwap_subB(st, c, s)
        assert np.array_equal(dev.host(), st.amps)
        st = rand_sv(rng, n, buffer=True); dev = Dev(st)
        capi.call("qb_statevec_anyCtrlSwap_subC", dev.ref, capi.ints(c), capi.ints(s), nc, t[0], 1)
        qo.statevec_anyCtrlSwap_subC(st, c, s, t[0], 1)
        assert np.array_equal(dev.host(), st.amps)
        # dense subB
        f0, f1 = complex(rng.normal(), rng.normal()), complex(rng.normal(), rng.normal())
        st = rand_sv(rng, n, buffer=True); dev = Dev(st)
        capi.call("qb_statevec_anyCtrlOneTargDenseMatr_subB", dev.ref, capi.ints(c), capi.ints(s), nc, capi.cplx(f0), capi.cplx(f1))
        qo.statevec_anyCtrlOneTargDenseMatr_subB(st, c, s, f0, f1)
        check_state(dev, st, f"dense subB n={n}")
        # pauli subB
        free = [q for q in range(n) if q not in c]
        k = min(3, len(free))
        tq = [int(q) for q in rng.choice(free, size=k, replace=False)]
        chars = rng.choice(list("XYZ"), size=k)
        x = [q for ch, q in zip(chars, tq) if ch == "X"]; y = [q for ch, q in zip(chars, tq) if ch == "Y"]
        z = [q for ch, q in zip(chars, tq) if ch == "Z"]
        maskXY = qo.getBitMask(x + y)
        # bufferMaskXY = removeBits(suffixMaskXY, sortedCtrls)  (core/localiser.cpp:1309-1313)
        bufMask, out_bit = 0, 0
        for b in range(n):
            if b in c:
                continue
            if (maskXY >> b) & 1:
                bufMask |= 1 << out_bit
            out_bit += 1
        st = rand_sv(rng, n, buffer=True); dev = Dev(st)
        capi.call("qb_statevector_anyCtrlPauliTensorOrGadget_subB", dev.ref, capi.ints(c), capi.ints(s), nc,
                  capi.ints(x), len(x), capi.ints(y), len(y), capi.ints(z), len(z), capi.cplx(f0), capi.cplx(f1), bufMask)
        qo.statevector_anyCtrlPauliTensorOrGadget_subB(st, c, s, x, y, z, f0, f1, bufMask)
        check_state(dev, st, f"pauli subB n={n}")
    if n >= 3:
        q1, q2, q3 = sorted(pick(rng, n, 3))
        st = rand_sv(rng, n, buffer=True); dev = Dev(st)
        npacked = C.c_longlong()
        capi.call("qb_statevec_packPairSummedAmpsIntoBuffer", dev.ref, q1, q2, q3, 1, C.byref(npacked))
        qo.statevec_packPairSummedAmpsIntoBuffer(st, q1, q2, q3, 1)
        assert rel_l2(dev.host_buf(), st.buffer) <= TOL


@pytest.mark.parametrize("n", [1, 4, 9, 15, 22])
def test_reductions_sv(n):
    rng = np.random.default_rng(800 + n)
    st = rand_sv(rng, n); dev = Dev(st)
    out = C.c_double(); outc = capi.qb_cplx()
    capi.call("qb_statevec_calcTotalProb_sub", dev.ref, C.byref(out))
    assert abs(out.value - qo.statevec_calcTotalProb_sub(st)) <= TOL
    for k in range(0, min(4, n) + 1):
        qs = pick(rng, n, k); oc = [int(b) for b in rng.integers(0, 2, size=k)]
        capi.call("qb_statevec_calcProbOfMultiQubitOutcome_sub", dev.ref, capi.ints(qs), capi.ints(oc), k, C.byref(out))
        assert abs(out.value - qo.statevec_calcProbOfMultiQubitOutcome_sub(st, qs, oc)) <= TOL
        probs = (C.c_double * (1 << k))()
        capi.call("qb_statevec_calcProbsOfAllMultiQubitOutcomes_sub", probs, dev.ref, capi.ints(qs), k)
        assert np.max(np.abs(np.array(probs) - qo.statevec_calcProbsOfAllMultiQubitOutcomes_sub(st, qs))) <= TOL
        capi.call("qb_statevec_calcExpecAnyTargZ_sub", dev.ref, capi.ints(qs), k, C.byref(out))
        assert abs(out.value - qo.statevec_calcExpecAnyTargZ_sub(st, qs)) <= TOL
        chars = rng.choice(list("XYZ"), size=k)
        x = [q for ch, q in zip(chars, qs) if ch == "X"]; y = [q for ch, q in zip(chars, qs) if ch == "Y"]
        z = [q for ch, q in zip(chars, qs) if ch == "Z"]
        capi.call("qb_statevec_calcExpecPauliStr_subA", dev.ref, capi.ints(x), len(x), capi.ints(y), len(y), capi.ints(z), len(z), C.byref(outc))
        want = qo.statevec_calcExpecPauliStr_subA(st, x, y, z)
        assert abs(complex(outc.re, outc.im) - want) <= TOL
    if n >= 12:
        probs = (C.c_double * (1 << 12))()
        qs = pick(rng, n, 12)
        capi.call("qb_statevec_calcProbsOfAllMultiQubitOutcomes_sub", probs, dev.ref, capi.ints(qs), 12)
        assert np.max(np.abs(np.array(probs) - qo.statevec_calcProbsOfAllMultiQubitOutcomes_sub(st, qs))) <= TOL
    other = rand_sv(rng, n); dev2 = Dev(other)
    capi.call("qb_statevec_calcInnerProduct_sub", dev.ref, dev2.ref, C.byref(outc))
    assert abs(complex(outc.re, outc.im) - qo.statevec_calcInnerProduct_sub(st, other)) <= TOL
    d = rng.uniform(0.2, 2, size=1 << n) + 1j * rng.uniform(-1, 1, size=1 << n)
    dd = dev_matrix(d)
    for hasPow, realPow, expo in ((0, 0, 1), (1, 0, 1.5 + 0.2j), (1, 1, 2.5)):
        capi.call("qb_statevec_calcExpecFullStateDiagMatr_sub", dev.ref, dd.data_ptr(), hasPow, realPow, capi.cplx(expo), C.byref(outc))
        want = qo.statevec_calcExpecFullStateDiagMatr_sub(st, d, bool(hasPow), bool(realPow), expo)
        assert abs(complex(outc.re, outc.im) - want) <= TOL * max(1, abs(want)) * 10
    # fused Pauli batch == the per-term reduction
    terms, masks = [], []
    for _ in range(11):
        k = int(rng.integers(1, min(n, 5) + 1)); qs = pick(rng, n, k); chars = rng.choice(list("XYZ"), size=k)
        x = [q for ch, q in zip(chars, qs) if ch == "X"]; y = [q for ch, q in zip(chars, qs) if ch == "Y"]
        z = [q for ch, q in zip(chars, qs) if ch == "Z"]
        terms.append((x, y, z)); masks += [qo.getBitMask(x + y), qo.getBitMask(y + z)]
    arr = (C.c_ulonglong * len(masks))(*masks)
    outs = (capi.qb_cplx * len(terms))()
    capi.call("qb_statevec_calcExpecPauliStrBatch_subA", dev.ref, arr, len(terms), outs)
    for (x, y, z), o in zip(terms, outs):
        want = qo.statevec_calcExpecPauliStr_subA(st, x, y, z) / qo.POWERS_OF_I[len(y) % 4]
        assert abs(complex(o.re, o.im) - want) <= TOL


@pytest.mark.parametrize("n", [1, 4, 9, 16, 21])
def test_elementwise_sv(n):
    rng = np.random.default_rng(900 + n)
    # projector
    k = min(2, n); qs = pick(rng, n, k); oc = [int(b) for b in rng.integers(0, 2, size=k)]
    st = rand_sv(rng, n); dev = Dev(st)
    capi.call("qb_statevec_multiQubitProjector_sub", dev.ref, capi.ints(qs), capi.ints(oc), k, 0.37)
    qo.statevec_multiQubitProjector_sub(st, qs, oc, 0.37)
    check_state(dev, st, "projector")
    # all-target diagonal
    d = np.exp(1j * rng.uniform(0, 6, 1 << n)); dd = dev_matrix(d)
    for hasPow, expo in ((0, 1), (1, 0.3 + 0.1j)):
        st = rand_sv(rng, n); dev = Dev(st)
        capi.call("qb_statevec_allTargDiagMatr_sub", dev.ref, dd.data_ptr(), hasPow, capi.cplx(expo))
        qo.statevec_allTargDiagMatr_sub(st, d, bool(hasPow), expo)
        check_state(dev, st, "allTargDiag")
    # superposition
    a, b, c = rand_sv(rng, n), rand_sv(rng, n), rand_sv(rng, n)
    da, db, dc = Dev(a), Dev(b), Dev(c)
    f = [complex(rng.normal(), rng.normal()) for _ in range(3)]
    capi.call("qb_statevec_setQuregToSuperposition_sub", capi.cplx(f[0]), da.ref, capi.cplx(f[1]), db.ref, capi.cplx(f[2]), dc.ref)
    qo.statevec_setQuregToSuperposition_sub(f[0], a, f[1], b, f[2], c)
    check_state(da, a, "superposition")
    # init
    for rank, logNodes in ((0, 0), (3, 2)):
        st = qo.new_state(n + logNodes, 0, rank, logNodes, with_buffer=False); dev = Dev(st)
        capi.call("qb_statevec_initDebugState_sub", dev.ref)
        qo.statevec_initDebugState_sub(st)
        assert np.array_equal(dev.host(), st.amps), "debug state must be bit-exact"
        capi.call("qb_statevec_initUniformState_sub", dev.ref, capi.cplx(0.25 - 0.5j))
        assert np.all(dev.host() == 0.25 - 0.5j)
    # random state: statistical check only (RNG streams are backend specific by design)
    st = qo.new_state(max(n, 12)); dev = Dev(st)
    capi.call("qb_statevec_initUnnormalisedUniformlyRandomPureStateAmps_sub", dev.ref, 1234)
    a = dev.host()
    assert abs(np.mean(np.abs(a) ** 2) - 2.0) < 0.2 and abs(np.mean(a)) < 0.1
    capi.call("qb_statevec_initUnnormalisedUniformlyRandomPureStateAmps_sub", dev.ref, 1234)
    assert np.array_equal(a, dev.host()), "same seed must reproduce the same state"


@pytest.mark.parametrize("n", [2, 3, 5, 8, 10])
def test_channels_local(n):
    """density-matrix channels, bra qubit in the suffix (no communication)"""
    rng = np.random.default_rng(1000 + n)
    for q in sorted({0, n // 2, n - 1}):
        for fname, args in (("oneQubitDephasing_subA", (0.2,)), ("oneQubitDepolarising_subA", (0.4,)), ("oneQubitDamping_subA", (0.3,)),
                            ("oneQubitPauliChannel_subA", (0.6, 0.1, 0.2, 0.1))):
            st = rand_dm(rng, n); dev = Dev(st)
            capi.call("qb_densmatr_" + fname, dev.ref, q, *args)
            getattr(qo, "densmatr_" + fname)(st, q, *args)
            check_state(dev, st, f"{fname} n={n} q={q}")
    for _ in range(3):
        a, b = pick(rng, n, 2)
        for fname in ("twoQubitDephasing_subA", "twoQubitDephasing_subB", "twoQubitDepolarising_subA", "twoQubitDepolarising_subB"):
            st = rand_dm(rng, n); dev = Dev(st)
            capi.call("qb_densmatr_" + fname, dev.ref, a, b, 0.35)
            getattr(qo, "densmatr_" + fname)(st, a, b, 0.35)
            check_state(dev, st, f"{fname} n={n} a={a} b={b}")


@pytest.mark.parametrize("n,logNodes", [(3, 1), (4, 2), (6, 3), (8, 3)])
def test_channels_prefix(n, logNodes):
    """bra qubit in the prefix: the post-exchange combine kernels, for every rank"""
    rng = np.random.default_rng(1100 + n)
    for rank in range(1 << logNodes):
        lo = n - logNodes            # ket qubits >= lo have prefix bra qubits
        for q in range(lo, n):
            for fname, args in (("oneQubitDephasing_subB", (0.2,)), ("oneQubitDepolarising_subB", (0.4,)), ("oneQubitDamping_subB", (0.3,)),
                                ("oneQubitDamping_subC", (0.3,)), ("oneQubitDamping_subD", (0.3,)), ("oneQubitPauliChannel_subB", (0.6, 0.1, 0.2, 0.1))):
                st = rand_dm(rng, n, rank, logNodes); dev = Dev(st)
                capi.call("qb_densmatr_" + fname, dev.ref, q, *args)
                getattr(qo, "densmatr_" + fname)(st, q, *args)
                check_state(dev, st, f"{fname} n={n} rank={rank} q={q}")
        if lo >= 1:
            k1, k2 = int(rng.integers(0, lo)), int(rng.integers(lo, n))
            for fname in ("twoQubitDepolarising_subC", "twoQubitDepolarising_subD"):
                st = rand_dm(rng, n, rank, logNodes); dev = Dev(st)
                capi.call("qb_densmatr_" + fname, dev.ref, k1, k2, 0.5)
                getattr(qo, "densmatr_" + fname)(st, k1, k2, 0.5)
                check_state(dev, st, f"{fname} n={n} rank={rank}")
        if logNodes >= 2:
            k1, k2 = sorted(pick(rng, logNodes, 2))
            k1, k2 = k1 + lo, k2 + lo
            for fname in ("twoQubitDepolarising_subE", "twoQubitDepolarising_subF"):
                st = rand_dm(rng, n, rank, logNodes); dev = Dev(st)
                capi.call("qb_densmatr_" + fname, dev.ref, k1, k2, 0.5)
                getattr(qo, "densmatr_" + fname)(st, k1, k2, 0.5)
                check_state(dev, st, f"{fname} n={n} rank={rank}")
        # rank-aware elementwise kernels
        st = rand_dm(rng, n, rank, logNodes); dev = Dev(st)
        qs = pick(rng, n, 2); oc = [1, 0]
        capi.call("qb_densmatr_multiQubitProjector_sub", dev.ref, capi.ints(qs), capi.ints(oc), 2, 0.8)
        qo.densmatr_multiQubitProjector_sub(st, qs, oc, 0.8)
        check_state(dev, st, "dm projector")
        st = rand_dm(rng, n, rank, logNodes); dev = Dev(st)
        a, b = pick(rng, n, 2)
        capi.call("qb_densmatr_twoQubitDephasing_subB", dev.ref, a, b, 0.3)
        qo.densmatr_twoQubitDephasing_subB(st, a, b, 0.3)
        check_state(dev, st, "2q dephasing rank-aware")
        st = rand_dm(rng, n, rank, logNodes); dev = Dev(st)
        st.buffer[:1 << n] = rand_state(rng, n); dev = Dev(st)
        capi.call("qb_densmatr_mixQureg_subC", C.c_double(0.7), dev.ref, C.c_double(0.3))
        qo.densmatr_mixQureg_subC(0.7, st, 0.3)
        check_state(dev, st, "mixQureg subC")


@pytest.mark.parametrize("n", [2, 4, 7, 10])
def test_reductions_dm(n):
    rng = np.random.default_rng(1200 + n)
    for rank, logNodes in ((0, 0), (1, 1), (2, 2)):
        if logNodes > n:
            continue
        st = rand_dm(rng, n, rank, logNodes); dev = Dev(st)
        out = C.c_double(); outc = capi.qb_cplx()
        capi.call("qb_densmatr_calcTotalProb_sub", dev.ref, C.byref(out))
        assert abs(out.value - qo.densmatr_calcTotalProb_sub(st)) <= TOL
        loc = n - logNodes
        for k in range(0, min(2, loc) + 1):
            qs = pick(rng, loc, k); oc = [int(b) for b in rng.integers(0, 2, size=k)]
            capi.call("qb_densmatr_calcProbOfMultiQubitOutcome_sub", dev.ref, capi.ints(qs), capi.ints(oc), k, C.byref(out))
            assert abs(out.value - qo.densmatr_calcProbOfMultiQubitOutcome_sub(st, qs, oc)) <= TOL
        for k in range(0, min(3, n) + 1):
            qs = pick(rng, n, k)
            probs = (C.c_double * (1 << k))()
            capi.call("qb_densmatr_calcProbsOfAllMultiQubitOutcomes_sub", probs, dev.ref, capi.ints(qs), k)
            assert np.max(np.abs(np.array(probs) - qo.densmatr_calcProbsOfAllMultiQubitOutcomes_sub(st, qs))) <= TOL
            capi.call("qb_densmatr_calcExpecAnyTargZ_sub", dev.ref, capi.ints(qs), k, C.byref(outc))
            assert abs(complex(outc.re, outc.im) - qo.densmatr_calcExpecAnyTargZ_sub(st, qs)) <= TOL
            chars = rng.choice(list("XYZ"), size=k)
            x = [q for ch, q in zip(chars, qs) if ch == "X"]; y = [q for ch, q in zip(chars, qs) if ch == "Y"]
            z = [q for ch, q in zip(chars, qs) if ch == "Z"]
            if logNodes == 0:
                capi.call("qb_densmatr_calcExpecPauliStr_sub", dev.ref, capi.ints(x), len(x), capi.ints(y), len(y), capi.ints(z), len(z), C.byref(outc))
                assert abs(complex(outc.re, outc.im) - qo.densmatr_calcExpecPauliStr_sub(st, x, y, z)) <= TOL
        other = rand_dm(rng, n, rank, logNodes); dev2 = Dev(other)
        capi.call("qb_densmatr_calcHilbertSchmidtDistance_sub", dev.ref, dev2.ref, C.byref(out))
        assert abs(out.value - qo.densmatr_calcHilbertSchmidtDistance_sub(st, other)) <= TOL * 10
        psi = rand_sv(rng, n); dpsi = Dev(psi)
        for conj in (0, 1):
            capi.call("qb_densmatr_calcFidelityWithPureState_sub", dev.ref, dpsi.ref, conj, C.byref(outc))
            assert abs(complex(outc.re, outc.im) - qo.densmatr_calcFidelityWithPureState_sub(st, psi, bool(conj))) <= TOL * 10
        d = rng.uniform(0.2, 2, size=1 << loc) + 0j; dd = dev_matrix(d)
        capi.call("qb_densmatr_calcExpecFullStateDiagMatr_sub", dev.ref, dd.data_ptr(), 1, 1, capi.cplx(2.0), C.byref(outc))
        assert abs(complex(outc.re, outc.im) - qo.densmatr_calcExpecFullStateDiagMatr_sub(st, d, True, True, 2.0)) <= TOL * 10
        # full-state diagonal on a density matrix + mixing + partial trace + Pauli-sum init
        dfull = np.exp(1j * rng.uniform(0, 6, 1 << n)); ddf = dev_matrix(dfull)
        for hasPow, mulOnly, expo in ((0, 0, 1), (0, 1, 1), (1, 0, 0.5)):
            s2 = rand_dm(rng, n, rank, logNodes); d2 = Dev(s2)
            capi.call("qb_densmatr_allTargDiagMatr_sub", d2.ref, ddf.data_ptr(), 1 << n, hasPow, mulOnly, capi.cplx(expo))
            qo.densmatr_allTargDiagMatr_sub(s2, dfull, bool(hasPow), bool(mulOnly), expo)
            check_state(d2, s2, "dm allTargDiag")
        s2 = rand_dm(rng, n, rank, logNodes); d2 = Dev(s2)
        capi.call("qb_densmatr_mixQureg_subA", C.c_double(0.6), d2.ref, C.c_double(0.4), dev.ref)
        qo.densmatr_mixQureg_subA(0.6, s2, 0.4, st)
        check_state(d2, s2, "mixQureg subA")
        terms = [(complex(rng.normal(), rng.normal()), (int(rng.integers(0, 4 ** n)), 0)) for _ in range(5)]
        coeffs = capi.cplx_array([t[0] for t in terms])
        strs = (C.c_ulonglong * 10)(*[v for t in terms for v in t[1]])
        s2 = rand_dm(rng, n, rank, logNodes); d2 = Dev(s2)
        capi.call("qb_densmatr_setAmpsToPauliStrSum_sub", d2.ref, coeffs, strs, 5)
        qo.densmatr_setAmpsToPauliStrSum_sub(s2, [t[0] for t in terms], [t[1] for t in terms])
        check_state(d2, s2, "setAmpsToPauliStrSum")
    if n >= 3:
        st = rand_dm(rng, n); dev = Dev(st)
        sv = rand_sv(rng, n); dsv = Dev(sv)
        capi.call("qb_densmatr_mixQureg_subB", C.c_double(0.6), dev.ref, C.c_double(0.4), dsv.ref)
        qo.densmatr_mixQureg_subB(0.6, st, 0.4, sv)
        check_state(dev, st, "mixQureg subB")
        for k in (1, 2):
            targs = pick(rng, n, k); pairs = [t + n for t in targs]
            out_st = qo.new_state(n - k, 1); dout = Dev(out_st)
            capi.call("qb_densmatr_partialTrace_sub", dev.ref, dout.ref, capi.ints(targs), capi.ints(pairs), k)
            qo.densmatr_partialTrace_sub(st, out_st, targs, pairs)
            check_state(dout, out_st, "partial trace")
        diag = qo.fullstatediagmatr_setElemsToPauliStrSum(1 << n, 0, [1.5, -0.5j], [(0b11, 0), (0b1100 if n > 1 else 0b11, 0)])
        dd2 = torch.zeros(1 << n, dtype=torch.complex128, device="cuda")
        capi.call("qb_fullstatediagmatr_setElemsToPauliStrSum", dd2.data_ptr(), 1 << n, 0, capi.cplx_array([1.5, -0.5j]),
                  (C.c_ulonglong * 4)(0b11, 0, 0b1100 if n > 1 else 0b11, 0), 2)
        capi.sync()
        assert rel_l2(dd2.cpu().numpy(), diag) <= TOL


def test_errors_are_loud():
    s = capi.qb_state()
    with pytest.raises(capi.QbError):
        capi.call("qb_statevec_calcTotalProb_sub", C.byref(s), None)
    st = qo.new_state(4); dev = Dev(st)
    with pytest.raises(capi.QbError):
        capi.call("qb_statevec_anyCtrlOneTargDenseMatr_subA", dev.ref, capi.ints([]), capi.ints([]), 0, 9, capi.cplx_array(np.eye(2)))
    assert capi.lib().qb_launch_count() > 0


def _random_fusable_op(rng, n, st, dev, prefix_bits=0):
    """apply one random fusable gate to both the oracle state and the device state (queued by the tile engine)"""
    kind = rng.choice(["dense1", "dense2", "pauli", "swap", "diag1", "diag2", "parity", "ladder"])
    nc = int(rng.integers(0, 3))
    if kind == "dense1":
        c, s, t = ctrl_targ(rng, n, nc, 1); m = rand_unitary(rng, 2)
        capi.call("qb_statevec_anyCtrlOneTargDenseMatr_subA", dev.ref, capi.ints(c), capi.ints(s), nc, t[0], capi.cplx_array(m))
        qo.statevec_anyCtrlOneTargDenseMatr_subA(st, c, s, t[0], m)
    elif kind == "dense2":
        c, s, t = ctrl_targ(rng, n, nc, 2); m = rand_unitary(rng, 4)
        capi.call("qb_statevec_anyCtrlTwoTargDenseMatr_sub", dev.ref, capi.ints(c), capi.ints(s), nc, t[0], t[1], capi.cplx_array(m))
        qo.statevec_anyCtrlTwoTargDenseMatr_sub(st, c, s, t[0], t[1], m)
    elif kind == "pauli":
        k = int(rng.integers(1, 5)); c, s, t = ctrl_targ(rng, n, nc, k)
        chars = rng.choice(list("XYZ"), size=k)
        if not any(ch in "XY" for ch in chars):
            chars[0] = "Y"
        x = [q for ch, q in zip(chars, t) if ch == "X"]; y = [q for ch, q in zip(chars, t) if ch == "Y"]; z = [q for ch, q in zip(chars, t) if ch == "Z"]
        th = rng.uniform(0, 6); af, pf = complex(np.cos(th)), 1j * np.sin(th)
        capi.call("qb_statevector_anyCtrlPauliTensorOrGadget_subA", dev.ref, capi.ints(c), capi.ints(s), nc,
                  capi.ints(x), len(x), capi.ints(y), len(y), capi.ints(z), len(z), capi.cplx(af), capi.cplx(pf))
        qo.statevector_anyCtrlPauliTensorOrGadget_subA(st, c, s, x, y, z, af, pf)
    elif kind == "swap":
        c, s, t = ctrl_targ(rng, n, nc, 2)
        capi.call("qb_statevec_anyCtrlSwap_subA", dev.ref, capi.ints(c), capi.ints(s), nc, t[0], t[1])
        qo.statevec_anyCtrlSwap_subA(st, c, s, t[0], t[1])
    elif kind == "diag1":
        c, s, _ = ctrl_targ(rng, n, nc, 0)
        t = int(rng.choice([q for q in range(n + prefix_bits) if q not in c])); e = np.exp(1j * rng.uniform(0, 6, 2))
        capi.call("qb_statevec_anyCtrlOneTargDiagMatr_sub", dev.ref, capi.ints(c), capi.ints(s), nc, t, capi.cplx_array(e))
        qo.statevec_anyCtrlOneTargDiagMatr_sub(st, c, s, t, e)
    elif kind == "diag2":
        c, s, _ = ctrl_targ(rng, n, nc, 0)
        t = [int(q) for q in rng.choice([q for q in range(n + prefix_bits) if q not in c], size=2, replace=False)]
        e = np.exp(1j * rng.uniform(0, 6, 4))
        capi.call("qb_statevec_anyCtrlTwoTargDiagMatr_sub", dev.ref, capi.ints(c), capi.ints(s), nc, t[0], t[1], capi.cplx_array(e))
        qo.statevec_anyCtrlTwoTargDiagMatr_sub(st, c, s, t[0], t[1], e)
    elif kind == "parity":
        k = int(rng.integers(1, 6)); c, s, t = ctrl_targ(rng, n, nc, k); f0, f1 = np.exp(1j * rng.uniform(0, 6, 2))
        capi.call("qb_statevector_anyCtrlAnyTargZOrPhaseGadget_sub", dev.ref, capi.ints(c), capi.ints(s), nc, capi.ints(t), k, capi.cplx(f0), capi.cplx(f1))
        qo.statevector_anyCtrlAnyTargZOrPhaseGadget_sub(st, c, s, t, f0, f1)
    else:
        # a QFT-style ladder: controlled phases sharing one qubit (merged into a phase star by the engine)
        centre = int(rng.integers(n)); others = [q for q in range(n) if q != centre]
        for o in rng.choice(others, size=min(len(others), int(rng.integers(2, 9))), replace=False):
            o = int(o); th = float(rng.uniform(0, 6)); e = [1, np.exp(1j * th)]
            targ, ctrl = (centre, o) if rng.integers(2) else (o, centre)
            capi.call("qb_statevec_anyCtrlOneTargDiagMatr_sub", dev.ref, capi.ints([ctrl]), capi.ints([1]), 1, targ, capi.cplx_array(e))
            qo.statevec_anyCtrlOneTargDiagMatr_sub(st, [ctrl], [1], targ, e)


@pytest.mark.parametrize("n", [13, 14, 17, 20, 22])
@pytest.mark.parametrize("rank,logNodes", [(0, 0), (2, 2)])
def test_fused_gate_sequences(n, rank, logNodes):
    """the tile engine proper: long random runs of fusable gates are queued, planned into multi-gate passes and
    executed by the TMA tile kernel; the result must equal gate-by-gate application by the oracle"""
    capi.call("qb_set_tile_engine", 1)
    rng = np.random.default_rng(1300 + n + rank)
    for trial in range(3):
        st = rand_sv(rng, n + logNodes, rank, logNodes); dev = Dev(st)
        launches0 = capi.lib().qb_launch_count()
        nops = [5, 40, 120][trial]
        for _ in range(nops):
            _random_fusable_op(rng, n, st, dev, prefix_bits=logNodes)
        err = rel_l2(dev.host(), st.amps)
        launches = capi.lib().qb_launch_count() - launches0
        assert err <= TOL * 5, f"fused n={n} ops={nops}: rel-L2 {err:.3e}"
        assert launches < nops, f"no fusion happened: {launches} launches for {nops} gates"


@pytest.mark.parametrize("mode", [1, 2], ids=["absorb+reorder", "program-order"])
@pytest.mark.parametrize("n,span", [(13, 4), (16, 16), (20, 7), (22, 22)])
def test_fused_dense_streams(n, span, mode):
    """bench-like streams of control-free dense 1/2-qubit gates (here crowded onto `span` qubits so that nearly every
    gate has a neighbour to be multiplied into), salted with controlled and diagonal gates that must block absorption
    and re-ordering across them: planner mode 1 (gate absorption + commuting re-order) and mode 2 (program order)"""
    capi.call("qb_set_tile_engine", mode)
    rng = np.random.default_rng(1700 + n + span)
    qubits = [int(q) for q in rng.choice(n, size=span, replace=False)]
    st = rand_sv(rng, n); dev = Dev(st)
    launches0 = capi.lib().qb_launch_count()
    nops = 150
    for _ in range(nops):
        r = rng.integers(10)
        if r < 4:
            t = int(rng.choice(qubits)); m = rand_unitary(rng, 2)
            capi.call("qb_statevec_anyCtrlOneTargDenseMatr_subA", dev.ref, capi.ints([]), capi.ints([]), 0, t, capi.cplx_array(m))
            qo.statevec_anyCtrlOneTargDenseMatr_subA(st, [], [], t, m)
        elif r < 8:
            t = [int(q) for q in rng.choice(qubits, size=2, replace=False)]; m = rand_unitary(rng, 4)
            capi.call("qb_statevec_anyCtrlTwoTargDenseMatr_sub", dev.ref, capi.ints([]), capi.ints([]), 0, t[0], t[1], capi.cplx_array(m))
            qo.statevec_anyCtrlTwoTargDenseMatr_sub(st, [], [], t[0], t[1], m)
        else:
            _random_fusable_op(rng, n, st, dev)
    err = rel_l2(dev.host(), st.amps)
    launches = capi.lib().qb_launch_count() - launches0
    capi.call("qb_set_tile_engine", 1)
    assert err <= TOL * 5, f"dense stream n={n} span={span} mode={mode}: rel-L2 {err:.3e}"
    assert launches < nops // 2


@pytest.mark.parametrize("n", [13, 18, 21])
def test_fused_qft_matches_gate_by_gate(n):
    """the QFT ladder (api/operations.cpp:1934-1953) through the phase-star merge vs the same gates one at a time"""
    rng = np.random.default_rng(1400 + n)
    psi = rand_sv(rng, n)
    results = []
    for engine in (1, 0):
        capi.call("qb_set_tile_engine", engine)
        st = qo.State(psi.amps.copy(), n); dev = Dev(st)
        h = np.array([[1, 1], [1, -1]]) / np.sqrt(2)
        for t in range(n - 1, -1, -1):
            capi.call("qb_statevec_anyCtrlOneTargDenseMatr_subA", dev.ref, capi.ints([]), capi.ints([]), 0, t, capi.cplx_array(h))
            for m in range(t):
                e = [1, np.exp(1j * np.pi / (1 << (m + 1)))]
                capi.call("qb_statevec_anyCtrlOneTargDiagMatr_sub", dev.ref, capi.ints([t - m - 1]), capi.ints([1]), 1, t, capi.cplx_array(e))
        for t in range(n // 2):
            capi.call("qb_statevec_anyCtrlSwap_subA", dev.ref, capi.ints([]), capi.ints([]), 0, t, n - 1 - t)
        results.append(dev.host())
    capi.call("qb_set_tile_engine", 1)
    # reference: DFT with the convention of tests/utils/linalg.cpp:160-176
    want = np.fft.ifft(psi.amps) * np.sqrt(1 << n)
    assert rel_l2(results[1], want) <= 1e-11, "gate-by-gate QFT disagrees with the DFT"
    assert rel_l2(results[0], want) <= 1e-11, "fused QFT disagrees with the DFT"
    assert rel_l2(results[0], results[1]) <= TOL * 5


@pytest.mark.parametrize("n", [3, 9, 16, 20])
def test_pauli_expectation_with_buffer_partner_subB(n):
    """calcExpecPauliStr_subB / Batch_subB (gpu_subroutines.cpp:1698): the partner amplitude a_j comes from the
    communication buffer (what a full exchange with rank ^ prefixXY leaves there), single GPU, rank emulated"""
    rng = np.random.default_rng(2100 + n)
    outc = capi.qb_cplx()
    for rank, logNodes in ((0, 1), (1, 1), (2, 2), (5, 3)):
        st = rand_sv(rng, n + logNodes, rank, logNodes, buffer=True); dev = Dev(st)
        terms, masks = [], []
        for _ in range(9):
            k = int(rng.integers(1, min(n, 5) + 1)); qs = pick(rng, n, k); chars = rng.choice(list("XYZ"), size=k)
            x = [q for ch, q in zip(chars, qs) if ch == "X"]; y = [q for ch, q in zip(chars, qs) if ch == "Y"]
            z = [q for ch, q in zip(chars, qs) if ch == "Z"]
            capi.call("qb_statevec_calcExpecPauliStr_subB", dev.ref, capi.ints(x), len(x), capi.ints(y), len(y), capi.ints(z), len(z), C.byref(outc))
            want = qo.statevec_calcExpecPauliStr_subB(st, x, y, z)
            assert abs(complex(outc.re, outc.im) - want) <= TOL, f"subB n={n} rank={rank}: {complex(outc.re, outc.im)} vs {want}"
            terms.append((x, y, z)); masks += [qo.getBitMask(x + y), qo.getBitMask(y + z)]
        arr = (C.c_ulonglong * len(masks))(*masks)
        outs = (capi.qb_cplx * len(terms))()
        capi.call("qb_statevec_calcExpecPauliStrBatch_subB", dev.ref, arr, len(terms), outs)
        for (x, y, z), o in zip(terms, outs):
            want = qo.statevec_calcExpecPauliStr_subB(st, x, y, z) / qo.POWERS_OF_I[len(y) % 4]
            assert abs(complex(o.re, o.im) - want) <= TOL


def test_scratch_cache_functions():
    """gpu_getCacheOfSize / gpu_clearCache / gpu_getCacheMemoryInBytes (gpu_config.cpp:639-687): grows monotonically,
    is reused when large enough, reports its size, and is released by clear"""
    lib = capi.lib()
    capi.call("qb_clear_cache")
    assert lib.qb_cache_bytes() == 0
    st = C.c_int(0)
    p1 = lib.qb_get_cache(1000, C.byref(st)); assert p1 and st.value == 0
    assert lib.qb_cache_bytes() == 1000 * 16
    p2 = lib.qb_get_cache(10, C.byref(st)); assert p2 == p1 and lib.qb_cache_bytes() == 1000 * 16     # reused, not shrunk
    p3 = lib.qb_get_cache(5000, C.byref(st)); assert p3 and lib.qb_cache_bytes() == 5000 * 16
    # the k >= 6 dense path is the cache's user in the reference (gpu_subroutines.cpp:475-498); here it must not disturb it
    rng = np.random.default_rng(5)
    s = rand_sv(rng, 10); dev = Dev(s)
    m = rand_unitary(rng, 64); dm = dev_matrix(m)
    targs = pick(rng, 10, 6)
    capi.call("qb_statevec_anyCtrlAnyTargDenseMatr_sub", dev.ref, capi.ints([]), capi.ints([]), 0, capi.ints(targs), 6, dm.data_ptr(), 0)
    qo.statevec_anyCtrlAnyTargDenseMatr_sub(s, [], [], targs, m, False)
    check_state(dev, s, "dense6 with a live cache")
    assert lib.qb_cache_bytes() >= 5000 * 16
    capi.call("qb_clear_cache")
    assert lib.qb_cache_bytes() == 0


@pytest.mark.parametrize("n", [13, 17])
def test_interleaved_states_keep_fusing(n):
    """per-state deferred queues: gates issued alternately on three states (the reference's psi / rho co-evolution
    pattern, tests/integration/densitymatrix.cpp:68-163) must not flush on every switch"""
    capi.call("qb_set_tile_engine", 1)
    rng = np.random.default_rng(2300 + n)
    sts = [rand_sv(rng, n) for _ in range(3)]
    devs = [Dev(s) for s in sts]
    stats0 = (C.c_double * 8)(); capi.call("qb_tile_stats", stats0)
    launches0 = capi.lib().qb_launch_count()
    nops = 60
    for _ in range(nops):
        for s, d in zip(sts, devs):
            _random_fusable_op(rng, n, s, d)
    errs = [rel_l2(d.host(), s.amps) for s, d in zip(sts, devs)]
    launches = capi.lib().qb_launch_count() - launches0
    stats1 = (C.c_double * 8)(); capi.call("qb_tile_stats", stats1)
    assert max(errs) <= TOL * 5, f"interleaved states: rel-L2 {errs}"
    assert launches < nops, f"switching states flushed the queues: {launches} launches for 3 x {nops} gates"
    assert stats1[4] - stats0[4] >= 3 * nops * 0.9 and stats1[0] > stats0[0]


def _wide_pauli(rng, n, min_sites=8):
    """a Pauli string with X/Y on many qubits, most of them high: fits no shared-memory tile"""
    k = int(rng.integers(min_sites, n + 1))
    t = pick(rng, n, k)
    chars = rng.choice(list("XYZ"), size=k, p=[0.4, 0.4, 0.2])
    x = [q for ch, q in zip(chars, t) if ch == "X"]; y = [q for ch, q in zip(chars, t) if ch == "Y"]; z = [q for ch, q in zip(chars, t) if ch == "Z"]
    return x, y, z


@pytest.mark.parametrize("n", [13, 16, 19, 21])
def test_wide_pauli_gadgets_share_passes(n):
    """Trotter-like streams: control-free Pauli gadgets on many high qubits, interleaved with Z-only phase gadgets and
    the occasional repeated (linearly dependent) string.  The engine runs them PG_K at a time through the coset-blocked
    kernel (quest_b200/csrc/qb_pauli_group.cu); the result must equal gate-by-gate application by the oracle."""
    capi.call("qb_set_tile_engine", 1)
    rng = np.random.default_rng(2500 + n)
    st = rand_sv(rng, n); dev = Dev(st)
    launches0 = capi.lib().qb_launch_count()
    nops, last = 45, None
    for i in range(nops):
        r = rng.integers(10)
        if r < 7 or last is None:
            x, y, z = _wide_pauli(rng, n)
            if not x and not y:
                x, z = z[:1], z[1:]
            last = (x, y, z)
        elif r < 8:
            x, y, z = last                                      # the same string again: dependent mask, closes the group
        else:
            k = int(rng.integers(1, 6)); t = pick(rng, n, k); f0, f1 = np.exp(1j * rng.uniform(0, 6, 2))
            capi.call("qb_statevector_anyCtrlAnyTargZOrPhaseGadget_sub", dev.ref, capi.ints([]), capi.ints([]), 0, capi.ints(t), k, capi.cplx(f0), capi.cplx(f1))
            qo.statevector_anyCtrlAnyTargZOrPhaseGadget_sub(st, [], [], t, f0, f1)
            continue
        th = rng.uniform(0, 6); af, pf = complex(np.cos(th)), 1j * np.sin(th)
        capi.call("qb_statevector_anyCtrlPauliTensorOrGadget_subA", dev.ref, capi.ints([]), capi.ints([]), 0,
                  capi.ints(x), len(x), capi.ints(y), len(y), capi.ints(z), len(z), capi.cplx(af), capi.cplx(pf))
        qo.statevector_anyCtrlPauliTensorOrGadget_subA(st, [], [], x, y, z, af, pf)
    err = rel_l2(dev.host(), st.amps)
    launches = capi.lib().qb_launch_count() - launches0
    assert err <= TOL * 5, f"wide gadgets n={n}: rel-L2 {err:.3e}"
    assert launches <= nops // 2, f"gadgets were not grouped: {launches} launches for {nops} gadgets"


@pytest.mark.parametrize("n", [12, 15, 20])
def test_pauli_expectation_batch_wide_terms(n):
    """calcExpecPauliStrBatch_subA on a random Hamiltonian-like term list: wide strings (grouped PG_K per pass), Z-only
    terms, repeated strings and short strings mixed; every term against the oracle's per-term reduction"""
    rng = np.random.default_rng(2700 + n)
    st = rand_sv(rng, n); dev = Dev(st)
    terms, masks = [], []
    for i in range(37):
        r = rng.integers(10)
        if r < 6:
            x, y, z = _wide_pauli(rng, n, min_sites=min(n, 6))
        elif r < 7 and terms:
            x, y, z = terms[-1]
        elif r < 8:
            x, y, z = [], [], pick(rng, n, int(rng.integers(1, 5)))
        else:
            qs = pick(rng, n, 2); x, y, z = [qs[0]], [qs[1]], []
        terms.append((x, y, z)); masks += [qo.getBitMask(x + y), qo.getBitMask(y + z)]
    arr = (C.c_ulonglong * len(masks))(*masks)
    outs = (capi.qb_cplx * len(terms))()
    capi.call("qb_statevec_calcExpecPauliStrBatch_subA", dev.ref, arr, len(terms), outs)
    for k, ((x, y, z), o) in enumerate(zip(terms, outs)):
        want = qo.statevec_calcExpecPauliStr_subA(st, x, y, z) / qo.POWERS_OF_I[len(y) % 4]
        assert abs(complex(o.re, o.im) - want) <= TOL, f"term {k} ({x},{y},{z}): {complex(o.re, o.im)} vs {want}"
