"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel name, share of total."""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
cols = rows[hdr]; ki, mi, vi = cols.index("Kernel Name"), cols.index("Metric Name"), cols.index("Metric Value")
idi = cols.index("ID")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg = collections.OrderedDict(); total = 0.0
for r in rows[hdr + 1:]:
    if len(r) <= vi or r[mi] != "gpu__time_duration.sum": continue
    if int(r[idi]) < skip: continue
    name = re.sub(r"\(.*", "", r[ki]); name = re.sub(r"^void |sb::", "", name)
    t = float(r[vi].replace(",", "")) / 1e3  # ns -> us
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += t; total += t
print(f"total {total/1e3:.2f} ms over {sum(a[0] for a in agg.values())} launches")
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t/1e3:9.3f} ms {100*t/total:5.1f}%  n={n:5d}  avg {t/n:9.1f} us  {name}")
