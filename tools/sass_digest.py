"""Per-kernel SASS digest of libsemabs_b200.so (run anywhere, no GPU):  python tools/sass_digest.py > profiles/rNN_sass_digest.txt
Counts the mnemonics that prove which hardware path a kernel uses (B200_PROFILING.md): UTCHMMA = tcgen05.mma,
LDTM / STTM = tcgen05.ld / st (TMEM), UTMALDG / UTMASTG = TMA load / store, UTCBAR = tcgen05.commit, SYNCS = mbarrier,
HMMA = mma.sync (legacy tensor path), LDGSTS = cp.async."""
import collections, os, re, subprocess, sys

so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                        "semantic-abstraction_b200", "libsemabs_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
MN = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "HMMA", "LDGSTS", "LDSM", "RED", "ATOM"]
kern, counts, regs = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern)
        counts[kern] = collections.Counter()
        continue
    if kern is None:
        continue
    m = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1).split(".")[0]
        counts[kern]["_total"] += 1
        for k in MN:
            if op.startswith(k):
                counts[kern][k] += 1
print(f"# SASS digest of {os.path.basename(so)} (cuobjdump -sass, sm_100a); columns = instruction counts per kernel")
print(f"{'kernel':78s} {'insts':>6s} " + " ".join(f"{k:>7s}" for k in MN))
tot = collections.Counter()
for k, c in counts.items():
    print(f"{k[:78]:78s} {c['_total']:6d} " + " ".join(f"{c[m]:7d}" for m in MN))
    tot.update(c)
print(f"{'TOTAL':78s} {tot['_total']:6d} " + " ".join(f"{tot[m]:7d}" for m in MN))
