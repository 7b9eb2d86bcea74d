"""Stall samples per SOURCE line: joins the per-instruction table printed by tools/ncu_hot.py (run on the GPU box) with the line
info of the locally built object:  python tools/ncu_lines.py hot.txt build/file.o kernel_substring [top]"""
import re, subprocess, sys, tempfile, os, collections
hot, obj, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(txt) if l.startswith(".text.") and kern in l][0]
insts, line = [], None
for l in txt[start + 1:]:
    if l.startswith("//-----"):
        break
    m = re.search(r'//## File ".*?/([^/"]+)", line (\d+)', l)
    if m:
        line = (m.group(1), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        insts.append((m.group(2), line))
lines = open(hot).read().split("\n")
sec = [i for i, l in enumerate(lines) if l.startswith("== sb::") and kern in l][0]
agg, why = collections.Counter(), collections.defaultdict(collections.Counter)
for l in lines[sec + 2:]:
    if l.startswith("=="):
        break
    m = re.match(r"\s+#\s*(\d+)\s+([\d.]+)%\s+(\S+)\s+(.*)", l)
    if m and int(m.group(1)) < len(insts):
        ln = insts[int(m.group(1))][1]
        agg[ln] += float(m.group(2))
        why[ln][m.group(3)] += float(m.group(2))
print(lines[sec]); print(lines[sec + 1])
for ln, v in agg.most_common(top):
    print(f"  {v:5.1f}%  {ln[0]}:{ln[1]:<5d} {dict((k, round(x, 1)) for k, x in why[ln].most_common(3))}")
