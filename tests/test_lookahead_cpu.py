"""The look-ahead gate log of the sharding shim (quest_b200/shim/lookahead.hpp, opt-in QUEST_B200_LOOKAHEAD=W) on the
host: the header is compiled alone into tests/native/lookahead_harness.cpp, which keeps the shim's qubit-map bookkeeping
with the exchange replaced by a counter, and is driven with bench.py's cfg-2 / cfg-3 gate streams.  Checked: logged gates
run exactly once and in program order, the log never holds 2W gates, a foreign access drains it, and the number of
half-shard exchanges equals tools/exchange_policy_study.py's independent Python model of the same rule -- for the default
LRU rule (W = 0) that model reproduces the exchange counts measured on the GPUs (profiles/r2_bench_{2,4,8}gpu.json)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import exchange_policy_study as S   # noqa: E402


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("lookahead") / "harness")
    subprocess.run(["g++", "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "quest_b200", "shim"),
                    os.path.join(ROOT, "tests", "native", "lookahead_harness.cpp"), "-o", exe], check=True)
    return exe


def _run(exe, n, n_local, window, lines):
    out = subprocess.run([exe], input=f"{n} {n_local} {window}\n" + "\n".join(lines) + "\n", text=True, capture_output=True, check=True)
    exchanges, ran, max_log, ok = (int(x) for x in out.stdout.split())
    return exchanges, ran, max_log, ok


@pytest.mark.parametrize("name", ["cfg2", "cfg3"])
@pytest.mark.parametrize("world", [2, 8])
@pytest.mark.parametrize("window", [0, 16, 64])
def test_exchange_counts_match_the_model(harness, name, world, window):
    n_local, steps = 30, 2
    n = n_local + world.bit_length() - 1
    stream = S.bench_stream(name, n)
    lines = S.stream_lines(stream * steps)
    exchanges, ran, max_log, ok = _run(harness, n, n_local, window, lines)
    assert ok == 1 and ran == len(lines)
    assert max_log <= max(0, 2 * window - 1)
    assert exchanges == round(S.simulate(stream, n, n_local, window, steps) * steps)


def test_look_ahead_cuts_exchanges_on_the_bench_circuit(harness):
    """the point of it: on the weak-scaled headline circuit a 64-gate window at least halves the exchanges"""
    for world in (2, 4, 8):
        n = 30 + world.bit_length() - 1
        lines = S.stream_lines(S.bench_stream("cfg2", n) * 3)
        lru = _run(harness, n, 30, 0, lines)[0]
        win = _run(harness, n, 30, 64, lines)[0]
        assert win * 2 <= lru, (world, lru, win)


def test_foreign_access_drains_and_swaps_rename(harness):
    # 4 qubits, 2 in the shard.  With the whole program in view, the H on qubit 3 evicts qubit 1: qubit 0 is needed next
    # -- under the name 2, after the SWAP renames it
    lines = ["g 1 3 0", "s 0 2", "g 1 2 0", "g 1 0 0"]
    ex_win = _run(harness, 4, 2, 8, lines)[0]
    ex_lru = _run(harness, 4, 2, 0, lines)[0]
    assert ex_win <= ex_lru
    # a drain after every gate leaves no future to look at: same count as LRU, still every gate exactly once
    drained = [x for ln in lines for x in (ln, "r")]
    ex, ran, max_log, ok = _run(harness, 4, 2, 8, drained)
    assert ok == 1 and ran == len(lines) and max_log <= 1 and ex == ex_lru
