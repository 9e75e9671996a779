"""Helpers shared by the -m gpu parity tests: run a model through the product (C ABI) and through
the float64 sparse oracle on the same inputs / weights."""
import numpy as np

from oracle import gn_oracle as O
import workloads as W

FP32_TOL = 1e-5     # north_star: fp32 path features within 1e-5 relative (SURVEY 8c definition)
BF16_TOL = 1e-2     # tensor-core path


def run_product(gn, layers, w, precision="fp32", eps_mode=0):
    model = W.to_gn_model(gn, layers, eps_mode=eps_mode)
    x = gn.batch(W.as_batch_input(w))
    y = model(x, precision=precision)
    c = lambda f: None if f is None else f.compact.cpu().numpy()
    return x, y, (c(y.ef), c(y.nf), c(y.gf))


def run_oracle(layers, w, eps_mode=0):
    g = O.lower(W.adj_list(w))
    ef, nf, gf = W.compact_inputs(w)
    return g, O.forward_sparse(layers, g, ef, nf, gf, eps_mode=eps_mode)


def assert_parity(got, ref, tol, what=""):
    for name, a, b in zip(("ef", "nf", "gf"), got, ref):
        assert (a is None) == (b is None), "%s %s: nothing-ness differs" % (what, name)
        if a is None:
            continue
        assert a.shape == b.shape, "%s %s: shape %s vs %s" % (what, name, a.shape, b.shape)
        assert np.isfinite(a).all(), "%s %s: non-finite output" % (what, name)
        err = O.rel_err(a, b)
        assert err <= tol, "%s %s: rel err %.3e > %.1e" % (what, name, err, tol)
