"""Two host threads x (context + stream) alternating e2e steps; variants: LOCK=1 serialises the steps."""
import sys, os, threading, time, ctypes as C; sys.path.insert(0, '.')
import numpy as np, torch
import graphnets_b200 as gn, workloads as W
from bench import synth
B = int(os.environ.get("GRAPHS", "4096"))
adj, ef, nf = synth("cfg4", B, 1000)
model = W.to_gn_model(gn, W.model_params("cfg4"))
x = gn.batch_compact(adj, ef, nf)
eng = x.graphs.engine
y = model(x, precision="auto"); torch.cuda.synchronize()
mask = torch.from_numpy(np.ascontiguousarray((adj == 1).transpose(0, 2, 1)).astype(np.uint8)).pin_memory()
h_ef, h_nf = torch.from_numpy(ef).pin_memory(), torch.from_numpy(nf).pin_memory()
E, N = x.graphs.E, x.graphs.N
nn = (C.c_int32 * B)(*([64] * B))
mh = model._model(eng)
P = lambda t: C.c_void_p(t.data_ptr())
engs = [eng, gn.pkg.engine.Engine(0)]
streams = [torch.cuda.Stream() for _ in range(2)]
for e_, s_ in zip(engs, streams): gn.pkg._lib.check(gn.lib.gnb_ctx_set_stream(e_.ctx, C.c_void_p(s_.cuda_stream)))
outs = [(torch.empty((E, 3)).pin_memory(), torch.empty((N, 4)).pin_memory(), torch.empty((B, 5)).pin_memory()) for _ in range(2)]
lock = threading.Lock() if os.environ.get("LOCK") else None
def step(k):
    h = C.c_void_p()
    gn.pkg._lib.check(gn.lib.gnb_graph_lower(engs[k].ctx, P(mask), 1, 0, nn, 64, B, B, C.byref(h)))
    gn.pkg._lib.check(gn.lib.gnb_model_forward_host(engs[k].ctx, mh, h, P(h_ef), P(h_nf), None, P(outs[k][0]), P(outs[k][1]), P(outs[k][2]), gn.pkg._lib.PRECISIONS[os.environ.get("PREC", "auto")]))
    gn.lib.gnb_graph_destroy(h)
def worker(k, n):
    torch.cuda.set_device(0)
    for _ in range(n):
        if lock:
            with lock: step(k)
        else: step(k)
def run(n):
    ts = [threading.Thread(target=worker, args=(k, n)) for k in range(2)]
    [t.start() for t in ts]; [t.join() for t in ts]
run(2); torch.cuda.synchronize()
t0 = time.perf_counter(); run(5); torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
print("LOCK" if lock else "free", os.environ.get("PREC", "auto"), "ms/step %.3f" % (dt * 1e3))
