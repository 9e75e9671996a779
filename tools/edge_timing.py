"""clock64 phase stamps of one steady-state pass of the fused edge kernel k_edge5 (csrc/tc_edge.cu, EDBG) + per-launch times.

    GNB_LIB_VARIANT=timing python tools/edge_timing.py [graphs]

Prints, relative to the start of the stamped pass on the MMA warp: when every block was issued, when the DRAIN warps saw /
finished every hidden chunk, how long the LayerNorm warps and the OUT slices took.  The variant library is a test-only build
(build.py VARIANTS); the product library carries no stamps."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np      # noqa: E402
import torch            # noqa: E402
import graphnets_b200 as gn      # noqa: E402
import workloads as W            # noqa: E402
from bench import synth          # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
adj, ef, nf = synth("cfg4", B, 1000)
model = W.to_gn_model(gn, W.model_params("cfg4"))
x = gn.batch_compact(adj, ef, nf)
assert gn.lib.gnb_debug_tc_timing(None, 0) == 0
eng = x.graphs.engine
for _ in range(3):
    y = model(x, precision="auto")
eng.sync()
n = 148 * 18 * 32
buf = (C.c_ulonglong * n)()
assert gn.lib.gnb_debug_tc_timing(buf, n) == 0
t = np.array(buf[:], dtype=np.int64).reshape(148, 18, 32)
NOUT = 8 if os.environ.get("GNB_LIB_VARIANT", "").endswith("18") else 6
MMA, DRAIN, LN, OUT = 8 + NOUT, [8, 9, 10, 11], [0, 1, 2, 3], [4, 5, 6, 7] + list(range(12, 8 + NOUT))


def show(name, warps, slots, labels):
    print("==", name)
    for cta in (0, 76):      # pair leaders (the peer CTA does not issue MMAs)
        base = t[cta, MMA, 0]      # MMA warp, start of the stamped pass
        for w in warps:
            print("  cta %3d warp %2d: " % (cta, w) + " ".join("%s=%d" % (l, t[cta, w, k] - base) for k, l in zip(slots, labels)))
    d = np.stack([t[:, w, :] for w in warps], 1).astype(float)
    dd = np.diff(d[:, :, slots], axis=2).mean(axis=(0, 1))
    print("  mean deltas:", " ".join("%s=%.0f" % (l, v) for l, v in zip(labels[1:], dd)))


show("MMA warp (start of each block)", [MMA], list(range(10)), ["up0", "up1", "up2", "blk", "dn0", "up3", "dn1", "dn2", "dn3", "end"])
show("DRAIN warps", DRAIN, list(range(12)), ["start", "hf0", "hs0", "hf1", "hs1", "hf2", "hs2", "hf3", "hs3", "outdone", "stgempty", "stgfull"])
show("LN warps", LN, list(range(7)), ["start", "aempty", "g0", "g1", "g2", "g3", "arrive"])
show("OUT warps (last slice of the stamped pass per warp)", OUT, list(range(3)), ["start", "stgfull", "done"])
eng.set_profiling(True)
eng.read_profile()
for _ in range(5):
    y = model(x, precision="auto")
prof = eng.read_profile()
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
    print("%-20s %8.3f ms/launch x %d per forward" % (k, v["ms"] / v["launches"], v["launches"] // 5))
