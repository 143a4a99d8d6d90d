"""GPU probe of the tile engine's COMPUTE path alone (timing build): cycles per gate and per 16-amplitude register block
for rounds of G dense gates on a shared-memory tile, with one or two warps per scheduler.  No HBM traffic, no TMA."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
lib = C.CDLL(os.environ.get("QB_TIMING_LIB", os.path.join(ROOT, "build", "timing", "libquest_b200_timing.so")))
CLK = 1.965e9
for kind in (2, 1):
    for wgs in (1, 2):
        for gates in (2, 8):
            ms, rounds = C.c_float(), C.c_int()
            reps = 2000
            buf = (C.c_ulonglong * 12)()
            lib.qb_tile_timing_read(buf, 1)
            rc = lib.qb_tile_round_probe(kind, gates, wgs, reps, C.byref(ms), C.byref(rounds))
            assert rc == 0, rc
            lib.qb_tile_timing_read(buf, 1)
            tot = sum(buf) or 1
            per = 148 * wgs * 2 * reps * 2 * gates          # (dual-block builds dispatch once per two blocks: halve the per-gate slot figures)          # SMs x warpgroups x launches(2) x reps x blocks x gates
            slots = f"prefetch {buf[10] / per:5.0f} clk/gate, dispatch+body {buf[11] / per:5.0f} clk/gate, lds {100 * buf[4] / tot:4.1f}% sts {100 * buf[6] / tot:4.1f}% exit {100 * buf[5] / tot:4.1f}%"
            clk_round = ms.value * 1e-3 * CLK / reps / rounds.value            # per round per warpgroup (2 blocks)
            ideal = gates * (256 if kind == 2 else 128) * 2 * 2 * (wgs)          # DFMA issue cycles per SMSP: instrs x 2 clk x 2 blocks x warps sharing the pipe
            print(f"kind={kind}q wgs={wgs} gates/round={gates:2d} rounds={rounds.value}: {clk_round:9.0f} clk/round, "
                  f"{clk_round / gates / 2:7.0f} clk per gate per block (FP64-pipe floor {ideal / gates / 2:5.0f}) -> pipe {100 * ideal / clk_round:5.1f}% | {slots}", flush=True)
