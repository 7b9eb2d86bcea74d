"""Hottest SASS instructions (warp-stall samples) of every kernel in an .ncu-rep captured with --import-source on:
    python tools/ncu_hot.py file.ncu-rep [top_n]
Prints, per kernel, the top instructions by stall samples with their dominant stall reason — run here, no GPU."""
import csv, io, subprocess, sys
rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(out)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and row and row[0] == "Address":
        cur["hdr"] = row
    elif cur is not None and cur["hdr"] and row:
        cur["rows"].append(row)
for b in blocks:
    h = b["hdr"]
    si = h.index("# Samples")
    stalls = [i for i, k in enumerate(h) if k.startswith("stall_") and "Not Issued" not in k]
    total = sum(int(r[si] or 0) for r in b["rows"])
    print(f"== {b['name'][:90]}  total samples {total}")
    agg = {}
    for r in b["rows"]:
        for i in stalls:
            agg[h[i]] = agg.get(h[i], 0) + int(r[i] or 0)
    print("   stall mix:", ", ".join(f"{k[6:]} {v / max(total, 1):.0%}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    ranked = sorted(enumerate(b["rows"]), key=lambda ir: -int(ir[1][si] or 0))[:top]
    for idx, r in sorted(ranked):
        n = int(r[si] or 0)
        why = max(stalls, key=lambda i: int(r[i] or 0))
        print(f"   #{idx:5d} {n / max(total, 1):6.1%}  {h[why][6:]:12s} {r[1].strip()[:100]}")
