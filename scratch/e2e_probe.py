import sys, time, ctypes as C
sys.path.insert(0, '.')
import numpy as np, torch
import graphnets_b200 as gn, workloads as W
from bench import synth
B=4096
adj, ef, nf = synth("cfg4", B, 1000)
layers = W.model_params("cfg4"); model = W.to_gn_model(gn, layers)
x = gn.batch_compact(adj, ef, nf); g=x.graphs; eng=g.engine
for _ in range(3): y = model(x, precision="auto")
torch.cuda.synchronize()
t0=time.perf_counter(); y = model(x, precision="auto"); torch.cuda.synchronize(); print("device fwd wall", time.perf_counter()-t0)
mask = torch.from_numpy(np.ascontiguousarray((adj == 1).transpose(0, 2, 1)).astype(np.uint8)).pin_memory()
h_ef, h_nf = torch.from_numpy(ef).pin_memory(), torch.from_numpy(nf).pin_memory()
E,N=g.E,g.N
h_oe = torch.empty((E, 3)).pin_memory(); h_on = torch.empty((N, 4)).pin_memory(); h_og = torch.empty((B, 5)).pin_memory()
nn = (C.c_int32 * B)(*([64] * B)); mh = model._model(eng); P = lambda t: C.c_void_p(t.data_ptr())
for i in range(5):
    t0=time.perf_counter()
    h = C.c_void_p()
    rc=gn.lib.gnb_graph_lower(eng.ctx, P(mask), 1, 0, nn, 64, B, B, C.byref(h)); assert rc==0
    t1=time.perf_counter()
    rc=gn.lib.gnb_model_forward_host(eng.ctx, mh, h, P(h_ef), P(h_nf), None, P(h_oe), P(h_on), P(h_og), 3); assert rc==0
    t2=time.perf_counter()
    gn.lib.gnb_graph_destroy(h)
    t3=time.perf_counter()
    print("lower %.1f ms  fwd_host %.1f ms  destroy %.1f ms" % ((t1-t0)*1e3,(t2-t1)*1e3,(t3-t2)*1e3))
