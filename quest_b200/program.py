"""A tiny JSON-able "program" format for driving QuEST's public API, and its interpreter.

The same program is executed on the B200 drop-in library, on the reference CPU library (in a worker
process, tests/_worker.py) and -- for the ops the numpy oracle restates -- by oracle/quest_oracle_api.py,
so parity tests are data: a list of API calls with the reference's own names and argument order.

program = {
  "seeds": [1, 2],                       # optional: setSeeds before anything else
  "quregs": {"psi": {"n": 6, "dm": 0, "init": "debug" | "zero" | "plus" | "blank" | ["classical", k],
                     "custom": [useDistrib, useGpuAccel, useMultithread]  (optional createCustomQureg)}},
  "ops": [ ["applyCompMatr1", "psi", 3, {"m1": [[re,im]*4]}], ["calcTotalProb", "psi"], ... ],
  "dump": ["psi"],                       # quregs whose full amplitudes are returned
  "dump_windows": {"rho": [[start, width], ...]}   # (optional) slices of the local flat amplitude array, returned as
}                                                  # dumps["rho@start"]: for states too large to dump whole
Special argument encodings (everything else is passed through as int/float/list-of-int):
  "name"                -> the Qureg of that name (only where the API takes a Qureg)
  {"m1": M} {"m2": M}   -> CompMatr1 / CompMatr2 (by value);   {"m": M} -> heap CompMatr (created, synced, destroyed)
  {"d1": D} {"d2": D}   -> DiagMatr1 / DiagMatr2;              {"d": D} -> heap DiagMatr
  {"fsd": D}            -> FullStateDiagMatr;  {"kraus": [K...]} -> KrausMap;  {"superop": S} -> SuperOp
  {"pauli": ["XZY", [q0,q1,q2]]} -> PauliStr;  {"paulisum": [["XZ", [q..], [re,im]], ...]} -> PauliStrSum
  {"c": [re, im]}       -> qcomp by value
  {"out_reals": k}      -> a qreal[k] output array (returned as the op's result)
  {"amps": [[re,im]..]} -> a host qcomp array passed by pointer (setQuregAmps, setDensityQuregFlatAmps ...)
  {"out_amps": k}       -> a qcomp[k] output array passed by pointer (getQuregAmps); returned as a complex ndarray
Complex matrices are nested lists of [re, im] pairs (or anything np.asarray(...).view understands).
"""
import ctypes as C
import numpy as np

from . import quest_api as qa


def enc_c(z):
    z = complex(z)
    return [z.real, z.imag]


def enc_mat(m):
    m = np.asarray(m, dtype=np.complex128)
    return np.stack([m.real, m.imag], axis=-1)          # float64 [..., 2]; .tolist() gives the JSON form


def dec_mat(m):
    a = np.asarray(m, dtype=np.float64)
    return a[..., 0] + 1j * a[..., 1]


class _Interp:
    def __init__(self, Q, prog):
        self.Q, self.prog = Q, prog
        self.quregs = {}
        self.cleanup = []

    def qureg(self, name):
        return self.quregs[name]

    def conv(self, a):
        Q = self.Q
        if isinstance(a, str):
            return self.quregs[a]
        if isinstance(a, dict):
            (k, v), = a.items()
            if k == "m1": return Q.getCompMatr1(dec_mat(v))
            if k == "m2": return Q.getCompMatr2(dec_mat(v))
            if k == "d1": return Q.getDiagMatr1(dec_mat(v))
            if k == "d2": return Q.getDiagMatr2(dec_mat(v))
            if k == "c": return qa.qcomp(float(v[0]), float(v[1]))
            if k == "m":
                obj = Q.newCompMatr(dec_mat(v)); self.cleanup.append(("destroyCompMatr", obj)); return obj
            if k == "d":
                obj = Q.newDiagMatr(dec_mat(v)); self.cleanup.append(("destroyDiagMatr", obj)); return obj
            if k == "fsd":
                obj = Q.newFullStateDiagMatr(dec_mat(v)); self.cleanup.append(("destroyFullStateDiagMatr", obj)); return obj
            if k == "kraus":
                obj = Q.newKrausMap([dec_mat(x) for x in v]); self.cleanup.append(("destroyKrausMap", obj)); return obj
            if k == "superop":
                obj = Q.newSuperOp(dec_mat(v)); self.cleanup.append(("destroySuperOp", obj)); return obj
            if k == "pauli":
                return Q.getPauliStr(v[0], v[1])
            if k == "paulisum":
                strs = [Q.getPauliStr(t[0], t[1]) for t in v]
                obj = Q.newPauliStrSum(strs, [complex(t[2][0], t[2][1]) for t in v])
                self.cleanup.append(("destroyPauliStrSum", obj)); return obj
            if k == "amps":
                arr = np.ascontiguousarray(dec_mat(v), dtype=qa.np_qcomp); self.keep.append(arr); return arr.ctypes.data
            if k == "out_amps":
                arr = np.zeros(int(v), dtype=qa.np_qcomp); self.outarr = arr; return arr.ctypes.data
            if k == "out_reals":
                arr = (qa.c_qreal * int(v))(); self.outarr = arr; return arr
            raise ValueError(f"unknown argument encoding {k}")
        if isinstance(a, (list, tuple)):
            arr = (C.c_int * max(1, len(a)))(*[int(x) for x in a])
            self.keep.append(arr)
            return arr
        return a

    def run(self):
        Q, prog = self.Q, self.prog
        if "seeds" in prog:
            Q.setSeeds(prog["seeds"])
        for name, spec in prog["quregs"].items():
            n, dm = int(spec["n"]), int(spec.get("dm", 0))
            if "custom" in spec:
                q = Q.createCustomQureg(n, dm, *[int(x) for x in spec["custom"]])
            else:
                q = Q.createDensityQureg(n) if dm else Q.createQureg(n)
            self.quregs[name] = q
            init = spec.get("init", "zero")
            if init == "debug": Q.initDebugState(q)
            elif init == "zero": Q.initZeroState(q)
            elif init == "plus": Q.initPlusState(q)
            elif init == "blank": Q.initBlankState(q)
            elif isinstance(init, (list, tuple)) and init[0] == "classical": Q.initClassicalState(q, int(init[1]))
            elif isinstance(init, (list, tuple)) and init[0] == "amps": Q.setAmps(q, dec_mat(init[1]))
            else: raise ValueError(f"unknown init {init}")
        results = []
        for op in prog["ops"]:
            name, args = op[0], op[1:]
            self.keep, self.outarr = [], None
            cargs = [self.conv(a) for a in args]
            out = getattr(Q.lib, name)(*cargs)
            if isinstance(self.outarr, np.ndarray):
                out = self.outarr
            elif self.outarr is not None:
                out = list(self.outarr)
            elif isinstance(out, qa.qcomp):
                out = [out.re, out.im]
            elif isinstance(out, qa.Qureg):
                # API calls creating a new Qureg (calcPartialTrace, createCloneQureg...): register as "_ret<k>"
                rname = f"_ret{len(results)}"
                self.quregs[rname] = out
                out = rname
            results.append(out)
            for fn, obj in self.cleanup:
                getattr(Q.lib, fn)(obj)
            self.cleanup.clear()
        dumps = {}
        for name in prog.get("dump", []):
            q = self.quregs[name]
            dumps[name] = Q.getLocalAmps(q) if (q.isDensityMatrix or q.isDistributed) else Q.getAmps(q)
        for name, wins in prog.get("dump_windows", {}).items():
            q = self.quregs[name]
            if q.isGpuAccelerated:
                Q.lib.syncQuregFromGpu(q)
            flat = Q._view(q.cpuAmps, q.numAmpsPerNode)
            for start, width in wins:
                dumps[f"{name}@{int(start)}"] = np.array(flat[int(start):int(start) + int(width)])
        info = {name: {f: getattr(q, f) for f in ("isGpuAccelerated", "isDistributed", "isMultithreaded", "rank",
                                                   "numNodes", "numAmpsPerNode")} for name, q in self.quregs.items()}
        for q in self.quregs.values():
            Q.destroyQureg(q)
        return {"results": results, "dumps": dumps, "info": info}


def run_program(Q, prog):
    """Execute `prog` on the loaded library `Q` (a quest_api.QuEST whose env is already initialised)."""
    return _Interp(Q, prog).run()
