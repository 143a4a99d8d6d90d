"""Distributed restatement of the amplitude-update path over torch.distributed -- TEST INFRASTRUCTURE ONLY.

A numpy re-statement of what quest/src/core/localiser.cpp does when a Qureg is sharded over P = 2^p ranks (rank r
holds global indices [r*N, (r+1)*N); the top p qubits are bits of the rank): which rank pairs exchange, what is packed,
and which post-exchange routine of oracle/quest_oracle.py (= quest/src/cpu/cpu_subroutines.cpp) combines the buffer.
It is the CPU model of quest_b200/shim/localiser_b200.cpp: tests/test_dist_cpu.py runs it on world_size-2/4 `gloo`
process groups and compares the re-assembled global state with the single-process oracle, which proves the
decomposition (pair ranks, packing order, buffer masks, sign rules) independently of any GPU.
Each function cites the localiser function it mirrors.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import quest_oracle as qo


class Shard:
    """the local part of a distributed register plus its communication buffer"""

    def __init__(self, st, rank, world):
        self.st, self.rank, self.world = st, rank, world
        self.logN = st.logNumAmpsPerNode

    def is_suffix(self, q):
        return q < self.logN

    def rank_bit(self, q):
        return (self.rank >> (q - self.logN)) & 1

    def pair_rank(self, prefix_qubits):
        r = self.rank
        for q in prefix_qubits:
            r ^= 1 << (q - self.logN)
        return r


def _exchange(send, pair):
    """pairwise exchange of equal-sized complex arrays (comm_routines.cpp:209-232)"""
    s = torch.from_numpy(np.ascontiguousarray(send).view(np.float64).copy())
    r = torch.empty_like(s)
    ops = [dist.P2POp(dist.isend, s, pair), dist.P2POp(dist.irecv, r, pair)]
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    return r.numpy().view(np.complex128)


def _reduce(x):
    t = torch.tensor(np.atleast_1d(np.asarray(x, dtype=np.complex128)).view(np.float64).copy())
    dist.all_reduce(t)
    out = t.numpy().view(np.complex128)
    return out[0] if out.size == 1 else out


def _localise_ctrls(sh, ctrls, states):
    """localiser.cpp:102-143: False if this rank's prefix bits violate a control, else the suffix controls"""
    c, s = [], []
    for q, b in zip(ctrls, states):
        if sh.is_suffix(q):
            c.append(q); s.append(b)
        elif sh.rank_bit(q) != b:
            return None
    return c, s


def _exchange_where(sh, pair, qubits, states):
    """localiser.cpp:448-460: whole shard if unconstrained, else pack then exchange sub-buffers"""
    st = sh.st
    if not qubits:
        st.buffer[:] = _exchange(st.amps, pair)
        return
    n = qo.statevec_packAmpsIntoBuffer(st, qubits, states)
    off = st.numAmpsPerNode // 2
    st.buffer[:n] = _exchange(st.buffer[off:off + n], pair)


def dense1(sh, ctrls, states, targ, m):
    """localiser.cpp:941-974"""
    cs = _localise_ctrls(sh, ctrls, states)
    if cs is None:
        return
    c, s = cs
    m = np.asarray(m, dtype=np.complex128).reshape(2, 2)
    if sh.is_suffix(targ):
        qo.statevec_anyCtrlOneTargDenseMatr_subA(sh.st, c, s, targ, m)
        return
    _exchange_where(sh, sh.pair_rank([targ]), c, s)
    b = sh.rank_bit(targ)
    qo.statevec_anyCtrlOneTargDenseMatr_subB(sh.st, c, s, m[b, b], m[b, 1 - b])


def swap(sh, ctrls, states, t1, t2):
    """localiser.cpp:836-903"""
    if t1 > t2:
        t1, t2 = t2, t1
    cs = _localise_ctrls(sh, ctrls, states)
    if cs is None:
        return
    c, s = cs
    pre1, pre2 = not sh.is_suffix(t1), not sh.is_suffix(t2)
    if pre1 and pre2:
        if sh.rank_bit(t1) == sh.rank_bit(t2):
            return
        _exchange_where(sh, sh.pair_rank([t1, t2]), c, s)
        qo.statevec_anyCtrlSwap_subB(sh.st, c, s)
    elif pre2:
        state = 1 - sh.rank_bit(t2)
        _exchange_where(sh, sh.pair_rank([t2]), c + [t1], s + [state])
        qo.statevec_anyCtrlSwap_subC(sh.st, c, s, t1, state)
    else:
        qo.statevec_anyCtrlSwap_subA(sh.st, c, s, t1, t2)


def denseK(sh, ctrls, states, targs, m):
    """localiser.cpp:997-1040 with quest_b200's choice of swap partner (highest free suffix qubit)"""
    if not all(sh.is_suffix(q) or sh.rank_bit(q) == b for q, b in zip(ctrls, states)):
        return
    if all(sh.is_suffix(t) for t in targs):
        c, s = _localise_ctrls(sh, ctrls, states)
        qo.statevec_anyCtrlAnyTargDenseMatr_sub(sh.st, c, s, list(targs), m, False)
        return
    new_ctrls, new_targs = list(ctrls), list(targs)
    free = sh.logN - 1
    for i, t in enumerate(targs):
        if sh.is_suffix(t):
            continue
        while free in new_targs:
            free -= 1
        if free in new_ctrls:
            new_ctrls[new_ctrls.index(free)] = t
        new_targs[i] = free
    for a, b in zip(targs, new_targs):
        if a != b:
            swap(sh, [], [], min(a, b), max(a, b))
    cs = _localise_ctrls(sh, new_ctrls, states)
    if cs is not None:
        qo.statevec_anyCtrlAnyTargDenseMatr_sub(sh.st, cs[0], cs[1], new_targs, m, False)
    for a, b in zip(targs, new_targs):
        if a != b:
            swap(sh, [], [], min(a, b), max(a, b))


def diag1(sh, ctrls, states, targ, elems):
    """localiser.cpp:1089-1104: never communicates; a prefix target reads the rank inside the routine"""
    cs = _localise_ctrls(sh, ctrls, states)
    if cs is not None:
        qo.statevec_anyCtrlOneTargDiagMatr_sub(sh.st, cs[0], cs[1], targ, elems)


def _prefix_pauli_elem(sh, prefixY, prefixZ):
    """paulis_getPrefixPaulisElem, api/paulis.cpp:185-208: each prefix Z gives (-1)^bit, each prefix Y gives
    +i (rank bit 1) or -i (rank bit 0)"""
    f = 1 + 0j
    for q in prefixY:
        f *= (1j if sh.rank_bit(q) else -1j)
    for q in prefixZ:
        f *= 1 - 2 * sh.rank_bit(q)
    return f


def pauli(sh, ctrls, states, x, y, z, ampFac, pairAmpFac):
    """localiser.cpp:1271-1316"""
    cs = _localise_ctrls(sh, ctrls, states)
    if cs is None:
        return
    c, s = cs
    sx = [q for q in x if sh.is_suffix(q)]; px = [q for q in x if not sh.is_suffix(q)]
    sy = [q for q in y if sh.is_suffix(q)]; py = [q for q in y if not sh.is_suffix(q)]
    sz = [q for q in z if sh.is_suffix(q)]; pz = [q for q in z if not sh.is_suffix(q)]
    pairAmpFac = pairAmpFac * _prefix_pauli_elem(sh, py, pz)
    if not px and not py:
        qo.statevector_anyCtrlPauliTensorOrGadget_subA(sh.st, c, s, sx, sy, sz, ampFac, pairAmpFac)
        return
    _exchange_where(sh, sh.pair_rank(px + py), c, s)
    mask = qo.getBitMask(sx + sy)
    buf_mask, out_bit = 0, 0
    for b in range(sh.logN):                     # removeBits(suffixMaskXY, sortedCtrls), localiser.cpp:1309-1313
        if b in c:
            continue
        if (mask >> b) & 1:
            buf_mask |= 1 << out_bit
        out_bit += 1
    qo.statevector_anyCtrlPauliTensorOrGadget_subB(sh.st, c, s, sx, sy, sz, ampFac, pairAmpFac, buf_mask)


def phase_gadget(sh, ctrls, states, targs, phase):
    """localiser.cpp:1247-1268"""
    cs = _localise_ctrls(sh, ctrls, states)
    if cs is None:
        return
    sign = 1
    for q in targs:
        if not sh.is_suffix(q):
            sign *= 1 - 2 * sh.rank_bit(q)
    suffix = [q for q in targs if sh.is_suffix(q)]
    qo.statevector_anyCtrlAnyTargZOrPhaseGadget_sub(sh.st, cs[0], cs[1], suffix, np.exp(1j * phase * sign), np.exp(-1j * phase * sign))


def prob_of_outcome(sh, qubits, outcomes):
    """localiser.cpp:1848-1866"""
    prob = 0.0
    if all(sh.is_suffix(q) or sh.rank_bit(q) == b for q, b in zip(qubits, outcomes)):
        qs = [(q, b) for q, b in zip(qubits, outcomes) if sh.is_suffix(q)]
        prob = qo.statevec_calcProbOfMultiQubitOutcome_sub(sh.st, [q for q, _ in qs], [b for _, b in qs])
    return float(_reduce(prob).real)


def projector(sh, qubits, outcomes, prob):
    """localiser.cpp:2295-2314: ranks whose prefix bits contradict the outcome zero themselves; the others project on
    their suffix qubits (or only renormalise when every projected qubit is a rank bit)"""
    if not all(sh.is_suffix(q) or sh.rank_bit(q) == b for q, b in zip(qubits, outcomes)):
        qo.statevec_initUniformState_sub(sh.st, 0)
        return
    qs = [(q, b) for q, b in zip(qubits, outcomes) if sh.is_suffix(q)]
    if not qs:
        sh.st.amps *= 1 / np.sqrt(prob)
    else:
        qo.statevec_multiQubitProjector_sub(sh.st, [q for q, _ in qs], [b for _, b in qs], prob)


def densmatr_projector(sh, qubits, outcomes, prob):
    """localiser.cpp:2317-2324: purely local; the kernel derives the global row/column from the rank"""
    qo.densmatr_multiQubitProjector_sub(sh.st, list(qubits), list(outcomes), prob)


def inner_product(shA, shB):
    """localiser.cpp:2193-2215: local sums, then one all-reduce of a complex"""
    return complex(_reduce(qo.statevec_calcInnerProduct_sub(shA.st, shB.st)))


def two_qubit_dephasing(sh, ketA, ketB, prob):
    """localiser.cpp:1439-1455: never communicates; the prefix variant reads the bra bits from the rank"""
    if not sh.is_suffix(max(ketA, ketB) + sh.st.numQubits):
        qo.densmatr_twoQubitDephasing_subB(sh.st, ketA, ketB, prob)
    else:
        qo.densmatr_twoQubitDephasing_subA(sh.st, ketA, ketB, prob)


def total_prob(sh):
    return float(_reduce(qo.statevec_calcTotalProb_sub(sh.st)).real)


def expec_pauli(sh, x, y, z):
    """localiser.cpp:2000-2036"""
    sx = [q for q in x if sh.is_suffix(q)]; px = [q for q in x if not sh.is_suffix(q)]
    sy = [q for q in y if sh.is_suffix(q)]; py = [q for q in y if not sh.is_suffix(q)]
    sz = [q for q in z if sh.is_suffix(q)]; pz = [q for q in z if not sh.is_suffix(q)]
    if not px and not py:
        if not (sx or sy or sz):
            v = qo.statevec_calcTotalProb_sub(sh.st)
        elif not (sx or sy):
            v = qo.statevec_calcExpecAnyTargZ_sub(sh.st, sz)
        else:
            v = qo.statevec_calcExpecPauliStr_subA(sh.st, sx, sy, sz)
    else:
        sh.st.buffer[:] = _exchange(sh.st.amps, sh.pair_rank(px + py))
        v = qo.statevec_calcExpecPauliStr_subB(sh.st, sx, sy, sz)
    return complex(_reduce(v * _prefix_pauli_elem(sh, py, pz)))


# ---- density-matrix channels whose bra qubit is a prefix qubit (localiser.cpp:1458-1629) -----------------

def _bra_pair(sh, ket):
    st = sh.st
    return sh.rank ^ (1 << (ket - st.logNumColsPerNode))


def depolarising(sh, ket, prob):
    st = sh.st
    if ket + st.numQubits < sh.logN:
        qo.densmatr_oneQubitDepolarising_subA(st, ket, prob)
        return
    braBit = (sh.rank >> (ket - st.logNumColsPerNode)) & 1
    _exchange_where(sh, _bra_pair(sh, ket), [ket], [braBit])
    qo.densmatr_oneQubitDepolarising_subB(st, ket, prob)


def damping(sh, ket, prob):
    st = sh.st
    if ket + st.numQubits < sh.logN:
        qo.densmatr_oneQubitDamping_subA(st, ket, prob)
        return
    braBit = (sh.rank >> (ket - st.logNumColsPerNode)) & 1
    pair = _bra_pair(sh, ket)
    half = st.numAmpsPerNode // 2
    if braBit == 1:
        qo.statevec_packAmpsIntoBuffer(st, [ket], [1])
        t = torch.from_numpy(st.buffer[half:2 * half].view(np.float64).copy())
        dist.send(t, pair)
        qo.densmatr_oneQubitDamping_subB(st, ket, prob)
    qo.densmatr_oneQubitDamping_subC(st, ket, prob)
    if braBit == 0:
        t = torch.empty(2 * half, dtype=torch.float64)
        dist.recv(t, pair)
        st.buffer[:half] = t.numpy().view(np.complex128)
        qo.densmatr_oneQubitDamping_subD(st, ket, prob)


def pauli_channel(sh, ket, pX, pY, pZ):
    st = sh.st
    pI = 1 - pX - pY - pZ
    if ket + st.numQubits < sh.logN:
        qo.densmatr_oneQubitPauliChannel_subA(st, ket, pI, pX, pY, pZ)
        return
    st.buffer[:] = _exchange(st.amps, _bra_pair(sh, ket))
    qo.densmatr_oneQubitPauliChannel_subB(st, ket, pI, pX, pY, pZ)


def dephasing(sh, ket, prob):
    st = sh.st
    if ket + st.numQubits < sh.logN:
        qo.densmatr_oneQubitDephasing_subA(st, ket, prob)
    else:
        qo.densmatr_oneQubitDephasing_subB(st, ket, prob)


# ------------------------------------------------------------------------------------------------
# lazy qubit relabelling: the CPU model of the "LAZY QUBIT RELABELLING" layer of quest_b200/shim/localiser_b200.cpp
# ------------------------------------------------------------------------------------------------
class RelabelledShard:
    """A Shard plus the permutation logical qubit -> index bit.  Uncontrolled swaps only edit the permutation; a dense or
    Pauli-X/Y target found on a rank bit is pulled into the shard with ONE prefix<->suffix swap against the least-recently-
    used free suffix qubit and stays there (the reference swaps in, applies, swaps back: localiser.cpp:997-1040);
    `canonicalise` undoes the permutation with physical swaps.  Every rank runs the same calls, so the copies agree."""

    def __init__(self, sh, num_qubits):
        self.sh = sh
        self.phys = list(range(num_qubits))
        self.logi = list(range(num_qubits))
        self.last_use = [0] * num_qubits
        self.clock = 0
        self.exchanges = 0

    def _map(self, qubits):
        out = []
        for q in qubits:
            self.clock += 1
            self.last_use[q] = self.clock
            out.append(self.phys[q])
        return out

    def _physical_swap(self, a, b):
        if a == b:
            return
        if not (self.sh.is_suffix(a) and self.sh.is_suffix(b)):
            self.exchanges += 1
        swap(self.sh, [], [], min(a, b), max(a, b))
        la, lb = self.logi[a], self.logi[b]
        self.logi[a], self.logi[b] = lb, la
        self.phys[la], self.phys[lb] = b, a

    def _pull(self, targs, ctrls):
        nl = self.sh.logN
        for i, t in enumerate(targs):
            if t < nl:
                continue
            used = set(targs) | set(ctrls)
            free = [p for p in range(nl) if p not in used]
            if not free:
                return
            victim = min(free, key=lambda p: (self.last_use[self.logi[p]], -p))      # least recently used; ties: highest bit
            self._physical_swap(victim, t)
            targs[i] = victim

    def relabel_swap(self, a, b):
        pa, pb = self.phys[a], self.phys[b]
        self.phys[a], self.phys[b] = pb, pa
        self.logi[pa], self.logi[pb] = b, a

    def dense(self, ctrls, states, targs, m):
        c, t = self._map(ctrls), self._map(targs)
        self._pull(t, c)
        if len(t) == 1:
            dense1(self.sh, c, states, t[0], m)
        else:
            denseK(self.sh, c, states, t, m)

    def swap(self, ctrls, states, t1, t2):
        if not ctrls:
            self.relabel_swap(t1, t2)
            return
        c = self._map(ctrls)
        a, b = self._map([t1, t2])
        swap(self.sh, c, states, min(a, b), max(a, b))

    def diag1(self, ctrls, states, targ, elems):
        c = self._map(ctrls)
        diag1(self.sh, c, states, self._map([targ])[0], elems)

    def pauli(self, ctrls, states, x, y, z, ampFac, pairAmpFac):
        c = self._map(ctrls)
        xy = self._map(list(x) + list(y))
        self._pull(xy, c)
        pauli(self.sh, c, states, xy[:len(x)], xy[len(x):], self._map(z), ampFac, pairAmpFac)

    def phase_gadget(self, ctrls, states, targs, phase):
        phase_gadget(self.sh, self._map(ctrls), states, self._map(targs), phase)

    def prob_of_outcome(self, qubits, outcomes):
        return prob_of_outcome(self.sh, self._map(qubits), outcomes)

    def canonicalise(self):
        for l in range(len(self.phys)):
            if self.phys[l] != l:
                self._physical_swap(self.phys[l], l)
