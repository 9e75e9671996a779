"""Opcode evidence per kernel from the built library:  python tools/sass_summary.py > profiles/r02_sass_summary.txt
Counts the SASS mnemonics that show the Blackwell-native paths (B200_PROFILING.md): UTC*MMA = tcgen05.mma, LDTM / STTM =
tcgen05.ld / st, UBLKCP = cp.async.bulk (TMA engine, 1-D bulk), UTMALDG / UTMASTG = tensor-map TMA, UTCBAR = tcgen05.commit,
SYNCS = mbarrier ops, plus HMMA (legacy mma.sync: must be absent) and LDL / STL (local-memory spills)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "graphnets.jl_b200", "libgnb200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WATCH = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UBLKCP", "UBLKPF", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "HMMA", "LDGSTS", "LDL", "STL", "FFMA2", "FADD2"]
kern, counts, total = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        total[kern] = 0
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if kern and m:
        op = m.group(1)
        total[kern] += 1
        base = op.split(".")[0]
        if base == "UTCHMMA":
            counts[kern]["UTCHMMA.2CTA" if ".2CTA" in op else "UTCHMMA"] += 1
        elif base in WATCH:
            counts[kern][base] += 1
print("# SASS opcode counts per kernel of graphnets.jl_b200/libgnb200.so (cuobjdump -sass), sm_100a")
print("# %-58s %7s  %s" % ("kernel (demangled name abbreviated)", "instrs", "  ".join(WATCH)))
agg = collections.Counter()
for k, c in counts.items():
    name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"\(.*", "", name)[:58]
    print("%-60s %7d  %s" % (name, total[k], "  ".join("%*d" % (len(w), c[w]) for w in WATCH)))
    agg.update(c)
print("%-60s %7d  %s" % ("TOTAL", sum(total.values()), "  ".join("%*d" % (len(w), agg[w]) for w in WATCH)))
