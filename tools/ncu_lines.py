"""Aggregate ncu source-page CSV (--print-source cuda,sass) into per-CUDA-line stall samples."""
import csv, collections, sys
path, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 35
rows = list(csv.reader(open(path)))
sec = None; hdr = None; data = collections.defaultdict(list)
for r in rows:
    if len(r) >= 2 and r[0] == "Function Name": sec = r[1]; continue
    if len(r) > 4 and r[0] == "Line No": hdr = r; continue
    if sec and hdr and len(r) == len(hdr): data[sec].append(r)
for sec, rs in data.items():
    if pat not in sec: continue
    i_samp = hdr.index("# Samples"); i_inst = hdr.index("Instructions Executed")
    src_rows = [r for r in rs if r[0] != ""]
    tot = sum(int(r[i_samp]) for r in src_rows)
    toti = sum(int(r[i_inst]) for r in src_rows)
    print(sec, "total samples", tot, "inst", toti)
    lines = sorted(((int(r[i_samp]), int(r[i_inst]), r[0], r[1][:100]) for r in src_rows), reverse=True)
    for s, i, l, src in lines[:top]:
        print("%6d %5.1f%% inst=%9d (%4.1f%%) L%s: %s" % (s, 100 * s / max(tot, 1), i, 100 * i / max(toti, 1), l, src))
    stall_cols = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    agg = {k: sum(int(r[hdr.index(k)]) for r in src_rows) for k in stall_cols}
    print(sorted(agg.items(), key=lambda kv: -kv[1])[:8])
