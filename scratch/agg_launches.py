# Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.
import csv, sys, collections
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]; ki, vi, ui = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(',', '')); u = r[ui]
    v = v / 1e3 if u in ('ns', 'nsecond') else (v * 1e3 if u in ('ms', 'msecond') else v)
    a = agg.setdefault(r[ki][:int(sys.argv[2]) if len(sys.argv) > 2 else 70], [0, 0.0, 0.0]); a[0] += 1; a[1] += v; a[2] = max(a[2], v)
tot = sum(a[1] for a in agg.values())
print(f"total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches")
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:72s} n={a[0]:5d} total_us={a[1]:12.1f} avg_us={a[1]/a[0]:10.1f} max_us={a[2]:10.1f} share={a[1]/tot:.3f}")
