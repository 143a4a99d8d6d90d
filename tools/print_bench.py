"""Prints the essentials of a bench.py JSON line: python tools/print_bench.py file.json"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d.get("roofline") or {}
print(f"N={d['n_gpus']} value {d['value']:.1f} {d['unit']}  ms/step {d['ms_per_step']:.2f}  e2e {d['e2e']['value'] if d.get('e2e') else None}")
print("  roofline frac %.2f physical %.3f fp64 %.3f launches/step %.1f" % (r.get("frac", 0), r.get("physical_frac", 0), r.get("fp64_frac", 0), r.get("launches_per_step", 0)))
if r.get("sections"): print("  sections", {k: round(v, 1) for k, v in r["sections"].items() if isinstance(v, (int, float))})
if r.get("nvlink"): print("  nvlink", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in r["nvlink"].items() if k != "note"})
print("  cpu_baseline", d.get("cpu_baseline"))
print("  clocks", d.get("clocks"))
for k, v in (d.get("secondary") or {}).items():
    print("  ", k, {a: (round(b, 2) if isinstance(b, float) else b) for a, b in v.items() if a != "workload"} if isinstance(v, dict) else v)
