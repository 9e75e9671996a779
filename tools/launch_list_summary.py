"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv --log-file F): time per kernel name and its share.

    python tools/launch_list_summary.py gpurun_out/r02_launches.csv [first_id last_id]"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0] != "ID"]
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 30
agg = collections.OrderedDict()
for r in rows:
    if not (lo <= int(r[0]) <= hi):
        continue
    name = r[4].replace("void <unnamed>::", "").replace("<unnamed>::", "").split("(")[0]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[-1].replace(",", "")) / 1e3      # ns -> us
tot = sum(v[1] for v in agg.values())
print("launches %d..%d: %d kernels, %.1f us in total (cold-cache, serialised: shares, not absolute times)" % (lo, min(hi, int(rows[-1][0])), sum(v[0] for v in agg.values()), tot))
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%5d x %9.1f us  %5.1f %%  %s" % (c, t, 100 * t / tot, n))
