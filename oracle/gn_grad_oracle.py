"""TEST INFRASTRUCTURE ONLY (like everything under oracle/): gradients of the reference forward, for the parity tests of the
training step.

The reference differentiates its forward with Zygote inside Flux.withgradient (examples/sort/sort.jl:122-132).  Here the same
forward - the sparse-semantics restatement of oracle/gn_oracle.py (gnblock_sparse / gncore / forward_sparse, which follow
src/gnblock.jl:63-69, src/edgefninput.jl, src/nodefninput.jl, src/graphfninput.jl, src/gncore.jl:56-68, src/gnfeedforward.jl:27-31,
src/gngraphnorm.jl) - is restated op for op on torch float64 CPU tensors and differentiated by torch.autograd.  The forward
values of this restatement are checked against gn_oracle.forward_sparse in tests/test_oracle.py, so the gradients are those of
the pinned forward.  Parity unpinned in the same sense as the forward (no Julia here)."""
import numpy as np
import torch

EPS_SQRT_VAR_EPS2, EPS_STD_PLUS_EPS, EPS_SQRT_VAR_EPS = 0, 1, 2


def _ln(x, l, eps_mode):
    mu = x.mean(dim=-1, keepdim=True)
    xc = x - mu
    var = (xc * xc).mean(dim=-1, keepdim=True)
    eps = float(l.get("eps", 1e-5))
    if eps_mode == EPS_SQRT_VAR_EPS2:
        den = torch.sqrt(var + eps * eps)
    elif eps_mode == EPS_STD_PLUS_EPS:
        den = torch.sqrt(var) + eps
    else:
        den = torch.sqrt(var + eps)
    return (xc / den) * l["gamma"] + l["beta"]


def _dense(x, W, b, relu=False):
    y = x @ W.T + b
    return torch.relu(y) if relu else y


def _cat(parts, rows):
    parts = [p for p in parts if p is not None]
    if not parts:
        return torch.zeros((rows, 0), dtype=torch.float64)
    return torch.cat(parts, dim=-1)


def _block(p, g, ef, nf, gf):
    E, N, B = g["E"], g["N"], g["B"]
    src, dst, eg, ng = (torch.as_tensor(np.asarray(g[k], np.int64)) for k in ("edge_src", "edge_dst", "edge_graph", "node_graph"))
    xe = _cat([ef, None if nf is None else nf[src], None if nf is None else nf[dst], None if gf is None else gf[eg]], E)
    h_e = _dense(xe, p["We"], p["be"])
    agg = torch.zeros((N, h_e.shape[1]), dtype=torch.float64).index_add(0, dst, h_e)
    xv = _cat([agg, nf, None if gf is None else gf[ng]], N)
    h_v = _dense(xv, p["Wn"], p["bn"])
    se = torch.zeros((B, h_e.shape[1]), dtype=torch.float64).index_add(0, eg, h_e)
    sv = torch.zeros((B, h_v.shape[1]), dtype=torch.float64).index_add(0, ng, h_v)
    xu = _cat([se, sv, gf], B)
    h_u = _dense(xu, p["Wg"], p["bg"])
    z = lambda t: None if t.shape[1] == 0 else t
    return z(h_e), z(h_v), z(h_u)


def _core(c, g, ef, nf, gf, eps_mode):
    xs = [ef, nf, gf]
    n1 = [_ln(xs[i], c["ln1"][i], eps_mode) for i in range(3)]
    n2 = [_ln(xs[i], c["ln2"][i], eps_mode) for i in range(3)]
    blk = _block(c["block"], g, n1[0], n1[1], n1[2])
    ff = [_dense(_dense(n2[i], c["ffn"][i]["W1"], c["ffn"][i]["b1"], relu=True), c["ffn"][i]["W2"], c["ffn"][i]["b2"]) for i in range(3)]
    return tuple((xs[i] + blk[i]) + ff[i] for i in range(3))


def _to_torch(layers):
    """numpy parameter tree -> the same tree of float64 leaf tensors that require grad."""
    def conv(o):
        if isinstance(o, np.ndarray) and o.dtype.kind == "f":
            return torch.tensor(o.astype(np.float64), requires_grad=True)
        if isinstance(o, dict):
            return {k: conv(v) for k, v in o.items()}
        if isinstance(o, list):
            return [conv(v) for v in o]
        return o
    return [(kind, conv(p)) for kind, p in layers]


def forward_and_grads(layers, g, ef, nf, gf, cot, eps_mode=EPS_SQRT_VAR_EPS2):
    """Forward in float64, then the gradient of  L = sum_k <y_k, cot_k>  (cot: cotangents dL/dy_k for ef / nf / gf, None where
    the output is `nothing`) w.r.t. every parameter and every input.  Returns (outputs, layers-shaped tree of parameter
    gradients in the (out, in) layout of the parameters, input gradients)."""
    tl = _to_torch(layers)
    tin = [None if a is None else torch.tensor(np.asarray(a, np.float64), requires_grad=True) for a in (ef, nf, gf)]
    x = list(tin)
    for kind, p in tl:
        x = list(_block(p, g, *x)) if kind == "block" else list(_core(p, g, *x, eps_mode))
    L = sum((y * torch.as_tensor(np.asarray(c, np.float64))).sum() for y, c in zip(x, cot) if y is not None and c is not None)
    L.backward()

    def grads(o):
        if isinstance(o, torch.Tensor):
            return None if o.grad is None else o.grad.numpy().copy()
        if isinstance(o, dict):
            return {k: grads(v) for k, v in o.items()}
        if isinstance(o, list):
            return [grads(v) for v in o]
        return o
    outs = tuple(None if y is None else y.detach().numpy() for y in x)
    return outs, [(kind, grads(p)) for kind, p in tl], tuple(None if t is None or t.grad is None else t.grad.numpy().copy() for t in tin)
