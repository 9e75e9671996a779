"""Generates tests/golden/*.npz - small seeded input / parameter / output vectors.

The reference itself cannot be executed in this environment (no Julia toolchain), so
the OUTPUT vectors are produced by the float64 sparse oracle (oracle/gn_oracle.py) after it has been
cross-checked against the float32 dense-broadcaster mirror of the reference formulation; the INDEX
goldens are the matrices written out in the reference's own tests (test/runtests.jl:480-508,
655-682), typed in by hand below.  Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import gn_oracle as O  # noqa: E402
import workloads as W  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def flatten_params(layers):
    out = {}
    for li, (kind, p) in enumerate(layers):
        pre = "L%d_%s_" % (li, kind)
        if kind == "block":
            for k in ("We", "be", "Wn", "bn", "Wg", "bg"):
                out[pre + k] = p[k]
            out[pre + "dims"] = np.array(p["din"] + p["dout"], np.int32)
        else:
            for k in ("We", "be", "Wn", "bn", "Wg", "bg"):
                out[pre + "blk_" + k] = p["block"][k]
            out[pre + "dims"] = np.array(p["dims"], np.int32)
            for i in range(3):
                for k in ("W1", "b1", "W2", "b2"):
                    out[pre + "ffn%d_%s" % (i, k)] = p["ffn"][i][k]
                for ln in ("ln1", "ln2"):
                    out[pre + "%s%d_gamma" % (ln, i)] = p[ln][i]["gamma"]
                    out[pre + "%s%d_beta" % (ln, i)] = p[ln][i]["beta"]
    return out


def unflatten_params(z):
    layers = []
    li = 0
    while True:
        if "L%d_block_dims" % li in z:
            pre = "L%d_block_" % li
            d = [int(v) for v in z[pre + "dims"]]
            p = dict(din=tuple(d[:3]), dout=tuple(d[3:]))
            for k in ("We", "be", "Wn", "bn", "Wg", "bg"):
                p[k] = z[pre + k]
            layers.append(("block", p))
        elif "L%d_core_dims" % li in z:
            pre = "L%d_core_" % li
            dims = tuple(int(v) for v in z[pre + "dims"])
            blk = dict(din=dims, dout=dims)
            for k in ("We", "be", "Wn", "bn", "Wg", "bg"):
                blk[k] = z[pre + "blk_" + k]
            p = dict(dims=dims, block=blk, ffn=[], ln1=[], ln2=[])
            for i in range(3):
                p["ffn"].append({k: z[pre + "ffn%d_%s" % (i, k)] for k in ("W1", "b1", "W2", "b2")})
                for ln in ("ln1", "ln2"):
                    p[ln].append(dict(gamma=z[pre + "%s%d_gamma" % (ln, i)], beta=z[pre + "%s%d_beta" % (ln, i)],
                                      eps=1e-5))
            layers.append(("core", p))
        else:
            break
        li += 1
    return layers


def make_case(name, layers, adjs, ef, nf, gf, eps_mode=0):
    g = O.lower(adjs)
    ye, yn, yg = O.forward_sparse(layers, g, ef, nf, gf, eps_mode=eps_mode)
    # cross-check with the dense-broadcaster mirror of the reference formulation (float32)
    d = O.DenseBatch(adjs)
    ep, npz = g["graph_edge_ptr"], g["graph_node_ptr"]
    efs = None if ef is None else [ef[ep[b]:ep[b + 1]] for b in range(g["B"])]
    nfs = None if nf is None else [nf[npz[b]:npz[b + 1]] for b in range(g["B"])]
    efp = None if ef is None else O.padef(adjs, efs, ef.shape[1])
    nfp = None if nf is None else O.padnf(adjs, nfs, nf.shape[1])
    gfp = None if gf is None else gf[:, None, :]
    de, dn, dg = O.forward_dense(layers, d, efp, nfp, gfp, eps_mode=eps_mode)
    if ye is not None:
        assert O.rel_err(np.concatenate(O.unpadef(adjs, de)), ye) < 2e-5, name
    if yn is not None:
        assert O.rel_err(np.concatenate(O.unpadnf(adjs, dn)), yn) < 2e-5, name
    if yg is not None:
        assert O.rel_err(dg[:, 0, :], yg) < 2e-5, name
    out = flatten_params(layers)
    out["n_graphs"] = np.int32(len(adjs))
    for b, a in enumerate(adjs):
        out["adj_%d" % b] = np.asarray(a, np.uint8)
    for k, v in (("ef", ef), ("nf", nf), ("gf", gf), ("ye", ye), ("yn", yn), ("yg", yg)):
        if v is not None:
            out[k] = v
    out["eps_mode"] = np.int32(eps_mode)
    for k in ("edge_src", "edge_dst", "edge_slot", "graph_edge_ptr", "graph_node_ptr", "node_in_ptr"):
        out["idx_" + k] = g[k].astype(np.int32)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, "E=%d N=%d B=%d" % (g["E"], g["N"], g["B"]))


def main():
    # cfg1: README example 1 (test/runtests.jl:180-216)
    w = W.make_workload("cfg1")
    ef, nf, gf = W.compact_inputs(w)
    make_case("cfg1_readme_block", W.model_params("cfg1"), W.adj_list(w), ef, nf, gf)
    # cfg2 at B=8: enc -> 2 x GNCore(10,5,3) -> dec on one shared 16-node structure
    w = W.make_workload("cfg2", B=8)
    ef, nf, gf = W.compact_inputs(w)
    make_case("cfg2_small", W.model_params("cfg2"), W.adj_list(w), ef, nf, gf)
    # variable-structure batch with graph features in, all three LayerNorm eps conventions
    rng = np.random.default_rng(77)
    adjs = [(rng.random((n, n)) < 0.45).astype(np.uint8) for n in (3, 7, 5, 1, 6)]
    g = O.lower(adjs)
    dims = (6, 5, 4)
    layers = [("block", W.block_params(rng, (3, 2, 4), dims)), ("core", W.core_params(rng, dims)),
              ("core", W.core_params(rng, dims)), ("block", W.block_params(rng, dims, (2, 3, 1)))]
    ef = rng.random((g["E"], 3), dtype=np.float32)
    nf = rng.random((g["N"], 2), dtype=np.float32)
    gf = rng.random((g["B"], 4), dtype=np.float32)
    for mode in (0, 1, 2):
        make_case("varbatch_eps%d" % mode, layers, adjs, ef, nf, gf, eps_mode=mode)


if __name__ == "__main__":
    main()
