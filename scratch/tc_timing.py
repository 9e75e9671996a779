import sys, ctypes as C; sys.path.insert(0,'.')
import numpy as np, torch
import graphnets_b200 as gn, workloads as W
from bench import synth
adj, ef, nf = synth("cfg4", 4096, 1000)
model = W.to_gn_model(gn, W.model_params("cfg4"))
x = gn.batch_compact(adj, ef, nf)
for _ in range(3): y = model(x, precision="bf16")
torch.cuda.synchronize()
buf = (C.c_ulonglong * (148*64))()
assert gn.lib.gnb_debug_tc_timing(buf, 148*64) == 0
t = np.array(buf[:], dtype=np.int64).reshape(148, 2, 32)
names = ["start","A ready"] + sum([["hid%d wait done"%c, "hid%d epi done"%c] for c in range(4)], []) + ["outdone"] + sum([["q%d A"%q,"q%d C"%q,"q%d E"%q] for q in range(4)], [])
for cta in (0, 77):
    for s in (0,1):
        tt = t[cta, s, :len(names)] - t[cta, s, 0]
        print("cta", cta, "sub", s, " ".join("%s=%d" % (n, v) for n, v in zip(names, tt)))
d = (t[:, :, :len(names)] - t[:, :, :1]).astype(float)
dd = np.diff(d, axis=2).mean(axis=(0,1))
print("mean phase durations (cycles):")
for n, v in zip(names[1:], dd): print("  %-16s %8.0f" % (n, v))
print("total", d[:, :, len(names)-1].mean())
