#!/usr/bin/env python
"""Host-only model of the swap-in policy of quest_b200/shim/localiser_b200.cpp (pullTargetsIntoShard) on bench.py's
cfg-2 and cfg-3 gate streams: counts half-shard exchanges per circuit for

  lru      the shipped rule -- the least-recently-used shard qubit the gate does not touch (high bits first)
  window W farthest-next-use over the next W gates of the stream (Belady restricted to a look-ahead window; qubits
           not used inside the window count as "never", ties broken by LRU) -- the opt-in QUEST_B200_LOOKAHEAD=W of
           quest_b200/shim/lookahead.hpp, whose log/replay blocking is modelled too.

No GPU and no library is involved; the model only tracks which logical qubits sit on rank bits.  It exists to size
the look-ahead victim choice of DESIGN.md section 5; tests/test_lookahead_cpu.py checks that the C++ log and victim
rule the shim uses give exactly these counts.

    python tools/exchange_policy_study.py [--local 30] [--steps 6]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402  (the gate streams; importing bench does not touch the GPU)


def needs(op):
    """(non-diagonal targets, every qubit the gate touches); swaps are relabelled for free by the shim"""
    k = op[0]
    if k in ("h", "rx", "m1"):
        return [op[1]], [op[1]]
    if k == "cphase":
        return [], [op[1], op[2]]
    if k == "cnot":                      # ("cnot", control, target)
        return [op[2]], [op[1], op[2]]
    if k == "m2":
        return [op[1], op[2]], [op[1], op[2]]
    if k == "swap":
        return None, [op[1], op[2]]
    raise ValueError(k)


def stream_lines(stream):
    """the stream in the text format of tests/native/lookahead_harness.cpp"""
    out = []
    for op in stream:
        nd, touched = needs(op)
        if nd is None:
            out.append(f"s {op[1]} {op[2]}")
        else:
            other = [q for q in touched if q not in nd]
            out.append("g " + " ".join(map(str, [len(nd)] + nd + [len(other)] + other)))
    return out


def simulate(stream, n, n_local, window, steps=1):
    """exchanges per circuit.  window 0 = the shipped LRU rule; W > 0 = farthest next use over the next W gates, with
    the log replayed in blocks exactly like quest_b200/shim/lookahead.hpp (2W gates logged -> the older W run), so a
    gate sees between W and 2W - 1 gates of future, never past the end of the program"""
    phys = list(range(n))                # logical -> index bit
    logi = list(range(n))                # index bit -> logical
    last_use = [0] * n
    clock = 0
    exchanges = 0
    prog = stream * steps
    info = [needs(op) for op in prog]

    # log_end[i] = size of the log (as a program position) while gate i is replayed
    log_end = [0] * len(prog)
    if window:
        start = 0
        for issued in range(1, len(prog) + 1):
            if issued - start >= 2 * window:
                for i in range(start, issued - window):
                    log_end[i] = issued
                start = issued - window
        for i in range(start, len(prog)):
            log_end[i] = len(prog)

    for i, op in enumerate(prog):
        nd, touched = info[i]
        if nd is None:                   # uncontrolled SWAP: relabel
            a, b = op[1], op[2]
            pa, pb = phys[a], phys[b]
            phys[a], phys[b] = pb, pa
            logi[pa], logi[pb] = b, a
            continue
        for q in [q for q in touched if q not in nd] + nd:
            clock += 1
            last_use[q] = clock
        targs = [phys[q] for q in nd]
        others = [phys[q] for q in touched if q not in nd]
        for k in range(len(targs)):
            t = targs[k]
            if t < n_local:
                continue
            used = set(targs) | set(others)
            nxt = None
            if window:
                nxt = [1 << 62] * n
                name_now = list(range(n))
                for j in range(i + 1, min(log_end[i], i + 1 + window)):
                    ndj = info[j][0]
                    if ndj is None:
                        a, b = prog[j][1], prog[j][2]
                        name_now[a], name_now[b] = name_now[b], name_now[a]
                        continue
                    for u in ndj:
                        if nxt[name_now[u]] == 1 << 62:
                            nxt[name_now[u]] = j
            victim = -1
            for lo in ((16 if n_local > 20 else 0), 0):
                for p in range(n_local - 1, lo - 1, -1):
                    if p in used:
                        continue
                    if victim < 0:
                        victim = p
                        continue
                    if nxt is not None and nxt[logi[p]] != nxt[logi[victim]]:
                        if nxt[logi[p]] > nxt[logi[victim]]:
                            victim = p
                    elif last_use[logi[p]] < last_use[logi[victim]]:
                        victim = p
                if victim >= 0:
                    break
            lt, lv = logi[t], logi[victim]
            logi[t], logi[victim] = lv, lt
            phys[lt], phys[lv] = victim, t
            targs[k] = victim
            exchanges += 1
    return exchanges / steps


def bench_stream(name, n):
    if name == "cfg2":
        return bench.qft_stream(n) + [o[:3] if o[0] == "m2" else o[:2] for o in bench.dense_stream(n)]
    return [o[:3] if o[0] in ("cnot", "m2") else o[:2] for o in bench.cfg3_stream(n)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--local", type=int, default=30)
    ap.add_argument("--steps", type=int, default=6)
    a = ap.parse_args()
    windows = (0, 16, 64, 256)
    print(f"{'workload':<10}{'GPUs':>5}{'qubits':>7}  " + "".join(f"{('lru' if w == 0 else 'W=' + str(w)):>10}" for w in windows))
    for name in ("cfg2", "cfg3"):
        for world in (2, 4, 8):
            n = a.local + world.bit_length() - 1
            stream = bench_stream(name, n)
            row = [simulate(stream, n, a.local, w, a.steps) for w in windows]
            print(f"{name:<10}{world:>5}{n:>7}  " + "".join(f"{x:>10.1f}" for x in row))


if __name__ == "__main__":
    main()
