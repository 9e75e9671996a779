"""clock64 phase timing of one steady-state tile of the fused edge kernel (gnb_debug_tc_timing)."""
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
        base = t[cta, 12, 0]      # MMA warp, start of the tile
        for w in warps:
            print("  cta %3d warp %2d: " % (cta, w) + " ".join("%s=%d" % (l, t[cta, w, k] - base) for k, l in zip(slots, labels)))
    d = np.stack([t[:, w, :] for w in warps], 1).astype(float)
    dd = np.diff(d[:, :, slots], axis=2).mean(axis=(0, 1))
    print("  mean deltas:", " ".join("%s=%.0f" % (l, v) for l, v in zip(labels[1:], dd)))
show("MMA warp (start of each block)", [12], list(range(10)), ["up0", "up1", "blk", "dn0", "dn1", "up2", "up3", "dn2", "dn3", "end"])
show("DRAIN warps", [8, 9, 10, 11], list(range(12)), ["start", "hf0", "hs0", "hf1", "hs1", "hf2", "hs2", "hf3", "hs3", "outdone", "stgempty", "stgfull"])
show("LN warps", [0, 1, 2, 3], list(range(7)), ["start", "aempty", "g0", "g1", "g2", "g3", "arrive"])
show("OUT warps", [4, 5, 6, 7], list(range(3)), ["start", "stgfull", "done"])
