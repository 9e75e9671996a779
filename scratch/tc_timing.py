"""clock64 phase timing of one steady-state tile pair of the fused edge kernel (gnb_debug_tc_timing)."""
import sys, ctypes as C; sys.path.insert(0, '.')
import numpy as np, torch
import graphnets_b200 as gn, workloads as W
from bench import synth
adj, ef, nf = synth("cfg4", 4096, 1000)
model = W.to_gn_model(gn, W.model_params("cfg4"))
x = gn.batch_compact(adj, ef, nf)
assert gn.lib.gnb_debug_tc_timing(None, 0) == 0
for _ in range(3): y = model(x, precision="bf16")
torch.cuda.synchronize()
n = 148 * 18 * 32
buf = (C.c_ulonglong * n)()
assert gn.lib.gnb_debug_tc_timing(buf, n) == 0
t = np.array(buf[:], dtype=np.int64).reshape(148, 18, 32)
def show(name, warps, slots, labels):
    print("==", name)
    for cta in (0, 77):
        base = t[cta, 16, 0]      # MMA warp block 0 start
        for w in warps:
            print("  cta %3d warp %2d: " % (cta, w) + " ".join("%s=%d" % (l, t[cta, w, k] - base) for k, l in zip(slots, labels)))
    d = np.stack([t[:, w, :] for w in warps], 1).astype(float)  # [cta, nw, 32]
    dd = np.diff(d[:, :, slots], axis=2).mean(axis=(0, 1))
    print("  mean deltas:", " ".join("%s=%.0f" % (l, v) for l, v in zip(labels[1:], dd)))
show("MMA warp: block start / after W wait / after issue", [16], list(range(9)) + [], ["b%d" % b for b in range(9)])
show("MMA warp: after weight wait", [16], [20 + b for b in range(9)], ["w%d" % b for b in range(9)])
show("MMA warp: issue done", [16], [9 + b for b in range(9)], ["i%d" % b for b in range(9)])
show("prologue warps", list(range(8, 16)), list(range(12)), ["start", "issued0", "aempty"] + ["g%d" % g for g in range(8)] + ["arrive"])
show("drain warps", list(range(0, 8)), list(range(21)), ["start", "afull"] + sum([["hf%d" % c, "hs%d" % c] for c in range(4)], []) + ["pref", "outdone"] + ["ss%d" % i for i in range(8)] + ["end"])
show("loader", [17], list(range(18)), ["h%d" % i for i in range(18)])
