"""Synthetic workloads of BASELINE.json's configs - shared by tests, bench.py and
smoke().  Pure numpy; seeds fixed; features U[0,1) float32 like the reference's rand(Float32, ...);
weights glorot-uniform, biases U(-0.1, 0.1), LayerNorm scale U(0.5,1.5) / bias U(-0.1,0.1) so every
parameter path is exercised.

A model is a list of ("block" | "core", params) with numpy `(out, in)` weights - the format the
oracle consumes; `to_gn_model` loads the same numbers into the product's layer objects."""
import numpy as np

README_ADJ = np.array([[1, 0, 1], [1, 1, 0], [0, 0, 1]], np.uint8)   # README.md / test/runtests.jl:190-194

CONFIGS = {
    # name: (enc_in, hidden, n_cores, dec_out, default B)
    "cfg1": dict(enc=None, block=((10, 5, 0), (3, 4, 5)), B=2, mode="single"),
    "cfg2": dict(enc=(10, 5, 0), hidden=(10, 5, 3), cores=2, dec=(3, 4, 5), B=1024, mode="single"),
    "cfg3": dict(enc=(0, 100, 0), hidden=(384, 384, 384), cores=2, dec=(2, 2, 0), B=4096, mode="vector"),
    "cfg4": dict(enc=(10, 5, 0), hidden=(128, 128, 128), cores=4, dec=(3, 4, 5), B=4096, mode="vector"),
    "cfg5": dict(enc=(10, 5, 0), hidden=(256, 256, 256), cores=4, dec=(3, 4, 5), B=65536, mode="vector"),
}
SEEDS = {"cfg1": 1, "cfg2": 2, "cfg3": 3, "cfg4": 4, "cfg5": 5}


# ------------------------------------------------------------------------------ parameters
def _glorot(rng, out, inn):
    lim = np.sqrt(6.0 / (inn + out)) if inn + out > 0 else 0.0
    return rng.uniform(-lim, lim, size=(out, inn)).astype(np.float32)


def block_params(rng, din, dout):
    a, b, c = din
    p, q, r = dout
    bias = lambda n: rng.uniform(-0.1, 0.1, n).astype(np.float32)
    return dict(din=tuple(din), dout=tuple(dout),
                We=_glorot(rng, p, a + 2 * b + c), be=bias(p),
                Wn=_glorot(rng, q, p + b + c), bn=bias(q),
                Wg=_glorot(rng, r, p + q + c), bg=bias(r))


def core_params(rng, dims, eps=1e-5):
    def ffn(d):
        return dict(W1=_glorot(rng, 4 * d, d), b1=rng.uniform(-0.1, 0.1, 4 * d).astype(np.float32),
                    W2=_glorot(rng, d, 4 * d), b2=rng.uniform(-0.1, 0.1, d).astype(np.float32))

    def ln(d):
        return dict(gamma=rng.uniform(0.5, 1.5, d).astype(np.float32),
                    beta=rng.uniform(-0.1, 0.1, d).astype(np.float32), eps=eps)
    return dict(dims=tuple(dims), block=block_params(rng, dims, dims),
                ffn=[ffn(d) for d in dims], ln1=[ln(d) for d in dims], ln2=[ln(d) for d in dims])


def model_params(name, seed=None):
    cfg = CONFIGS[name]
    rng = np.random.default_rng(1000 + (SEEDS[name] if seed is None else seed))
    if "block" in cfg:
        return [("block", block_params(rng, *cfg["block"]))]
    layers = [("block", block_params(rng, cfg["enc"], cfg["hidden"]))]
    layers += [("core", core_params(rng, cfg["hidden"])) for _ in range(cfg["cores"])]
    layers.append(("block", block_params(rng, cfg["hidden"], cfg["dec"])))
    return layers


def to_gn_model(gn, layers, eps_mode=0):
    """Load oracle-format parameters into product layer objects; returns gn.GNSequential."""
    objs = []
    for kind, p in layers:
        if kind == "block":
            objs.append(_fill_block(gn.GNBlock(p["din"], p["dout"]), p))
        else:
            c = gn.GNCore(p["dims"], eps_mode=eps_mode)
            _fill_block(c.block, p["block"])
            for ch, f in zip((c.ffwd.eff, c.ffwd.nff, c.ffwd.gff), p["ffn"]):
                ch[0].set(f["W1"], f["b1"])
                ch[1].set(f["W2"], f["b2"])
            for gnorm, key in ((c.gn1, "ln1"), (c.gn2, "ln2")):
                for ln, l in zip((gnorm.edgeln, gnorm.nodeln, gnorm.graphln), p[key]):
                    ln.set(l["gamma"], l["beta"])
                    ln.eps = float(l["eps"])
            objs.append(c)
    return gn.GNSequential(*objs)


def _fill_block(blk, p):
    blk.edgefn[0].set(p["We"], p["be"])
    blk.nodefn[0].set(p["Wn"], p["bn"])
    blk.graphfn[0].set(p["Wg"], p["bg"])
    return blk


# ------------------------------------------------------------------------------ structures
def random_cells_adj(rng, B, n, m, chunk=2048):
    """B adjacency matrices (n x n, uint8) with exactly m distinct active cells each, chosen
    uniformly without replacement (self-loops allowed) - cfg4 / cfg5."""
    out = np.zeros((B, n * n), np.uint8)
    for b0 in range(0, B, chunk):
        b1 = min(B, b0 + chunk)
        keys = rng.random((b1 - b0, n * n), dtype=np.float32)
        sel = np.argpartition(keys, m - 1, axis=1)[:, :m]
        np.put_along_axis(out[b0:b1], sel, 1, axis=1)
    return out.reshape(B, n, n)


def make_workload(name, B=None, seed=None, n_nodes=None, n_edges=None):
    """Returns dict(mode, graphs, ef, nf, gf) in the reference's UNBATCHED input form:
    single mode: graphs (n,n), ef (DE,m,B), nf (DN,n,B), gf (DG,B);
    vector mode: lists of per-graph arrays ef (DE,m_b), nf (DN,n_b), gf (DG,)."""
    cfg = CONFIGS[name]
    B = cfg["B"] if B is None else B
    rng = np.random.default_rng(SEEDS[name] if seed is None else seed)
    f = lambda *s: rng.random(s, dtype=np.float32)
    if name == "cfg1":
        adj = README_ADJ
        m, n = int(adj.sum()), 3
        return dict(mode="single", graphs=adj, ef=np.asfortranarray(f(B, m, 10).transpose(2, 1, 0)),
                    nf=np.asfortranarray(f(B, n, 5).transpose(2, 1, 0)), gf=None)
    if name == "cfg2":
        n = 16
        adj = (rng.random((n, n)) < 0.25).astype(np.uint8)
        m = int(adj.sum())
        return dict(mode="single", graphs=adj, ef=np.asfortranarray(f(B, m, 10).transpose(2, 1, 0)),
                    nf=np.asfortranarray(f(B, n, 5).transpose(2, 1, 0)), gf=None)
    if name == "cfg3":
        lo, hi = (8, 64) if n_nodes is None else n_nodes
        ns = rng.integers(lo, hi + 1, size=B)
        graphs = [np.ones((n, n), np.uint8) for n in ns]                 # examples/sort/sort.jl:14
        nf = []
        for n in ns:
            labels = rng.integers(0, 100, size=n)
            oh = np.zeros((100, n), np.float32)
            oh[labels, np.arange(n)] = 1.0
            nf.append(oh)
        return dict(mode="vector", graphs=graphs, ef=None, nf=nf, gf=None)
    # cfg4 / cfg5
    n = 64 if n_nodes is None else n_nodes
    m = 512 if n_edges is None else n_edges
    adj = random_cells_adj(rng, B, n, m)
    graphs = [adj[b] for b in range(B)]
    ef = [np.asfortranarray(f(m, 10).T) for _ in range(B)]
    nf = [np.asfortranarray(f(n, 5).T) for _ in range(B)]
    return dict(mode="vector", graphs=graphs, ef=ef, nf=nf, gf=None)


def as_batch_input(w):
    return dict(graphs=w["graphs"], ef=w["ef"], nf=w["nf"], gf=w["gf"])


def compact_inputs(w):
    """Compact [rows][D] float32 arrays (oracle / engine layout) of a workload."""
    if w["mode"] == "single":
        t = lambda a: None if a is None else np.ascontiguousarray(a.transpose()).reshape(-1, a.shape[0])
        return t(w["ef"]), t(w["nf"]), t(w["gf"])
    cat = lambda xs: None if xs is None else np.ascontiguousarray(
        np.concatenate([x.T if x.ndim == 2 else x[None, :] for x in xs], 0))
    return cat(w["ef"]), cat(w["nf"]), cat(w["gf"])


def adj_list(w):
    """List of B adjacency matrices (single mode replicates the shared structure)."""
    if w["mode"] == "single":
        B = (w["ef"] if w["ef"] is not None else w["nf"]).shape[2]
        return [w["graphs"]] * B
    return w["graphs"]


# ------------------------------------------------------------------------------ canonical work
def canonical_work(layers, E, N, G):
    """Algorithmic flops / bytes of SURVEY 8d: reference-model work on real (unpadded) entities,
    fp32 I/O, each layer reads its input once and writes its output once (+ int32 index)."""
    flops = 0
    byts = 0
    idx = 8 * E + 4 * (N + 1) + 4 * (G + 1)
    for kind, p in layers:
        if kind == "block":
            (a, b, c), (pp, q, r) = p["din"], p["dout"]
            flops += 2 * E * (a + 2 * b + c) * pp + 2 * N * (b + pp + c) * q + 2 * G * (q + pp + c) * r
            byts += 4 * (E * (a + pp) + N * (b + q) + G * (c + r)) + idx
        else:
            de, dn, dg = p["dims"]
            blk = 2 * E * (de + 2 * dn + dg) * de + 2 * N * (dn + de + dg) * dn + 2 * G * (dn + de + dg) * dg
            ffn = 16 * (E * de * de + N * dn * dn + G * dg * dg)
            flops += blk + ffn
            byts += 8 * (E * de + N * dn + G * dg) + idx
    return flops, byts
