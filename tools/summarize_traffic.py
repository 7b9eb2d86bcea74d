"""Summarise an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv` log per kernel name:
    python tools/summarize_traffic.py launches.csv [kernel-name-filter]"""
import collections, csv, re, sys

rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
flt = sys.argv[2] if len(sys.argv) > 2 else ""
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
cols = rows[hdr]
ki, mi, vi, ui = cols.index("Kernel Name"), cols.index("Metric Name"), cols.index("Metric Value"), cols.index("Metric Unit")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3}
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi or flt not in r[ki]:
        continue
    name = re.sub(r"^void |sb::", "", re.sub(r"\(.*", "", r[ki]))
    a = agg.setdefault(name, {"n": 0, "rd": 0.0, "wr": 0.0, "t": 0.0})
    v = float(r[vi].replace(",", "")) * scale.get(r[ui], 1)
    if r[mi] == "dram__bytes_read.sum":
        a["rd"] += v
    elif r[mi] == "dram__bytes_write.sum":
        a["wr"] += v
    elif r[mi] == "gpu__time_duration.sum":
        a["t"] += v
        a["n"] += 1
tot_b = sum(a["rd"] + a["wr"] for a in agg.values())
tot_t = sum(a["t"] for a in agg.values())
n = sum(a["n"] for a in agg.values())
print(f"DRAM traffic {tot_b / 1e9:.2f} GB (read {sum(a['rd'] for a in agg.values()) / 1e9:.2f}, write {sum(a['wr'] for a in agg.values()) / 1e9:.2f}) "
      f"in {tot_t * 1e3:.1f} ms (serialised under ncu) over {n} launches; {tot_b / max(n, 1) / 1e6:.2f} MB per launch")
for name, a in sorted(agg.items(), key=lambda kv: -(kv[1]["rd"] + kv[1]["wr"]))[:16]:
    b = a["rd"] + a["wr"]
    print(f"  {b / 1e9:7.2f} GB  rd {a['rd'] / 1e9:6.2f} wr {a['wr'] / 1e9:6.2f}  {a['t'] * 1e3:7.2f} ms n={a['n']:4d}  {b / a['t'] / 1e9 if a['t'] else 0:7.0f} GB/s  {name[:70]}")
