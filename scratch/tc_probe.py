import sys; sys.path.insert(0,'.')
import numpy as np, torch
import graphnets_b200 as gn, workloads as W
from oracle import gn_oracle as O
for B in (1, 6, 64):
    w = W.make_workload("cfg4", B=B)
    layers = W.model_params("cfg4")
    model = W.to_gn_model(gn, layers)
    x = gn.batch(W.as_batch_input(w))
    y32 = model(x, precision="fp32")
    y = model(x, precision="bf16")
    torch.cuda.synchronize()
    g = O.lower(W.adj_list(w)); ef,nf,gf = W.compact_inputs(w)
    ref = O.forward_sparse(layers, g, ef, nf, gf)
    for n, f, f32, r in zip("eng", (y.ef,y.nf,y.gf), (y32.ef,y32.nf,y32.gf), ref):
        print(B, n, "bf16 rel", O.rel_err(f.compact.cpu().numpy(), r), "fp32 rel", O.rel_err(f32.compact.cpu().numpy(), r), "nan", bool(torch.isnan(f.compact).any()))
