"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: python tools/ncu_launch_summary.py file.csv"""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]; kn, mv = H.index("Kernel Name"), H.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[hdr + 1:]:
    if len(r) <= mv: continue
    try: v = float(r[mv].replace(",", ""))
    except ValueError: continue
    name = re.sub(r"\(.*", "", r[kn])
    agg[name].append(v / 1e6)     # ns -> ms
tot = sum(sum(v) for v in agg.values())
print(f"{sum(len(v) for v in agg.values())} launches, {tot:.1f} ms total")
print("| kernel | launches | total ms | share | avg ms | min | max |\n|---|---|---|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"| `{k[:70]}` | {len(v)} | {sum(v):.2f} | {100 * sum(v) / tot:.1f}% | {sum(v) / len(v):.3f} | {min(v):.3f} | {max(v):.3f} |")
