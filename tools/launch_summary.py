"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
H = rows[hdr]; ki = H.index("Kernel Name"); vi = H.index("Metric Value"); ui = H.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
    agg[r[ki][:56]][0] += 1; agg[r[ki][:56]][1] += v
tot = sum(v[1] for v in agg.values())
print(f"total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 14]:
    print(f"{k:58s} n={v[0]:5d} total_us={v[1]:10.1f} share={v[1]/tot:.3f} avg_us={v[1]/v[0]:.1f}")
