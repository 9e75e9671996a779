"""Timed CPU baseline: the reference FORMULATION (dense broadcasters + batched matmul on padded
slots) restated op-for-op on torch CPU tensors, float32, all host threads.

TEST / BENCH INFRASTRUCTURE ONLY (see oracle/gn_oracle.py header).  This is what
`bench.py --impl reference` and bench.py's `cpu_baseline` leg time, because the real reference needs
Julia + Flux 0.14, which are not installed (kind = "port").  Every op mirrors a reference call site:
NNlib.batched_mul -> torch.bmm (threaded over the batch like NNlib's CPU path), Dense -> one GEMM on
the (K, PE*B) reshape, LayerNorm over the feature dim, vcat -> torch.cat (materialised, as in
src/edgefninput.jl:2).  Checked against the float64 sparse oracle in tests/test_oracle.py.
"""
import numpy as np
import torch

from . import gn_oracle as O


class TorchDenseBatch:
    """GNGraphBatch(adj_mats) (src/gngraphbatch.jl:33-54) as torch float32 tensors [B][rows][cols]."""

    def __init__(self, adj_mats):
        d = O.DenseBatch(adj_mats)
        self.adj_mats = adj_mats
        self.B, self.PN, self.PE = d.B, d.PN, d.PE
        t = torch.from_numpy
        self.src, self.dst, self.g2e = t(d.src), t(d.dst), t(d.g2e)
        self.e2n, self.g2n, self.e2g, self.n2g = t(d.e2n), t(d.g2n), t(d.e2g), t(d.n2g)


def _bm(F, M):
    # Julia batched_mul(F (D,T,B), M (T,T',B)) -> (D,T',B); here F [B][T][D], M [B][T][T'] -> [B][T'][D]
    return torch.bmm(M.transpose(1, 2), F)


def _dense(x, W, b, relu=False):
    y = torch.addmm(b, x.reshape(-1, x.shape[-1]), W.t()).reshape(*x.shape[:-1], W.shape[0])
    return torch.relu_(y) if relu else y


def _ln(x, l, eps_mode):
    mu = x.mean(-1, keepdim=True)
    xc = x - mu
    var = (xc * xc).mean(-1, keepdim=True)
    eps = l["eps"]
    den = torch.sqrt(var + eps * eps) if eps_mode == 0 else (torch.sqrt(var) + eps if eps_mode == 1 else torch.sqrt(var + eps))
    return xc / den * l["gamma"] + l["beta"]


def to_torch_params(layers):
    def cv(p):
        if isinstance(p, dict):
            return {k: cv(v) for k, v in p.items()}
        if isinstance(p, list):
            return [cv(v) for v in p]
        if isinstance(p, np.ndarray):
            return torch.from_numpy(np.ascontiguousarray(p, dtype=np.float32))
        return p
    return [(k, cv(p)) for k, p in layers]


def block_dense(p, d, ef, nf, gf):
    """src/gnblock.jl:63-69 with src/edgefninput.jl:1-8, src/nodefninput.jl:1-7, src/graphfninput.jl:1-7"""
    parts = []
    if ef is not None:
        parts.append(ef)
    if nf is not None:
        parts += [_bm(nf, d.src), _bm(nf, d.dst)]
    if gf is not None:
        parts.append(_bm(gf, d.g2e))
    h_e = _dense(torch.cat(parts, -1), p["We"], p["be"])
    parts = [_bm(h_e, d.e2n)]
    if nf is not None:
        parts.append(nf)
    if gf is not None:
        parts.append(_bm(gf, d.g2n))
    h_v = _dense(torch.cat(parts, -1), p["Wn"], p["bn"])
    parts = [_bm(h_e, d.e2g), _bm(h_v, d.n2g)]
    if gf is not None:
        parts.append(gf)
    h_u = _dense(torch.cat(parts, -1), p["Wg"], p["bg"])
    z = lambda a: None if a.shape[-1] == 0 else a
    return z(h_e), z(h_v), z(h_u)


def forward(layers, d, ef, nf, gf, eps_mode=0):
    for kind, p in layers:
        if kind == "block":
            ef, nf, gf = block_dense(p, d, ef, nf, gf)
        else:
            xs = [ef, nf, gf]
            n1 = [_ln(xs[i], p["ln1"][i], eps_mode) for i in range(3)]
            n2 = [_ln(xs[i], p["ln2"][i], eps_mode) for i in range(3)]
            blk = block_dense(p["block"], d, *n1)
            ff = [_dense(_dense(n2[i], p["ffn"][i]["W1"], p["ffn"][i]["b1"], True), p["ffn"][i]["W2"], p["ffn"][i]["b2"])
                  for i in range(3)]
            ef, nf, gf = ((xs[i] + blk[i]) + ff[i] for i in range(3))
    return ef, nf, gf


def pad_inputs(adj_mats, ef_c, nf_c, gf_c, g):
    """batch(x) (src/batch.jl:53-64): compact oracle arrays -> padded torch tensors."""
    ep, npz = g["graph_edge_ptr"], g["graph_node_ptr"]
    B = g["B"]
    t = lambda a: None if a is None else torch.from_numpy(a)
    efp = None if ef_c is None else O.padef(adj_mats, [ef_c[ep[b]:ep[b + 1]] for b in range(B)], ef_c.shape[1])
    nfp = None if nf_c is None else O.padnf(adj_mats, [nf_c[npz[b]:npz[b + 1]] for b in range(B)], nf_c.shape[1])
    gfp = None if gf_c is None else np.ascontiguousarray(gf_c[:, None, :])
    return t(efp), t(nfp), t(gfp)
