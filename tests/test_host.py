"""CPU-side tests: the C-ABI library loads and exports every symbol include/gnb200.h declares,
the host mirror raises the reference's AssertionErrors (src/checks.jl), sharding logic."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ADJ = np.array([[1, 0, 1], [1, 1, 0], [0, 0, 1]])


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "gnb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gnb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(gn):
    names = _header_symbols()
    assert len(names) >= 24
    lib = ctypes.CDLL(gn.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "libgnb200.so does not export %s" % n
    # and the ctypes binding covers the same set
    assert set(names) == set(gn.pkg._lib.SIGNATURES.keys())
    assert lib.gnb_version() >= 100


def test_struct_layout_matches_header(gn):
    L = gn.pkg._lib
    assert ctypes.sizeof(L.BlockParams) == 6 * 4 + 6 * 8
    assert ctypes.sizeof(L.FfnParams) == 32 and ctypes.sizeof(L.LnParams) == 24
    assert ctypes.sizeof(L.CoreParams) == 72 + 3 * 32 + 6 * 24
    assert ctypes.sizeof(L.Layer) == 8 + 72 + ctypes.sizeof(L.CoreParams)


def test_no_cpu_fallback_without_device(gn):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    x = dict(graphs=ADJ, ef=np.zeros((10, 5, 2), np.float32), nf=np.zeros((5, 3, 2), np.float32), gf=None)
    with pytest.raises(gn.pkg.GnbError):
        gn.batch(x)


def test_checks_assertions(gn):
    f = lambda *s: np.zeros(s, np.float32)
    bad = [
        dict(graphs=ADJ, ef=f(10, 4, 2), nf=f(5, 3, 2), gf=None),          # wrong edge count
        dict(graphs=ADJ, ef=f(10, 5, 2), nf=f(5, 4, 2), gf=None),          # wrong node count
        dict(graphs=ADJ, ef=f(10, 5, 2), nf=f(5, 3, 3), gf=None),          # batch sizes differ
        dict(graphs=ADJ, ef=f(10, 5), nf=f(5, 3, 2), gf=None),             # ef not 3-D
        dict(graphs=ADJ, ef=f(10, 5, 2), nf=f(5, 3, 2), gf=f(4, 2, 1)),    # gf not 2-D
        dict(graphs=ADJ, ef=None, nf=None, gf=None),                       # nothing at all
        dict(graphs=[ADJ, ADJ], ef=[f(10, 5)], nf=[f(5, 3), f(5, 3)], gf=None),   # list lengths
        dict(graphs=[ADJ], ef=[f(10, 5, 1)], nf=[f(5, 3)], gf=None),       # ef[i] not 2-D
        dict(graphs=[ADJ], ef=[f(10, 5)], nf=[f(5, 3)], gf=[f(4, 1)]),     # gf[i] not 1-D
        dict(graphs=[], ef=[], nf=[], gf=None),
    ]
    for x in bad:
        with pytest.raises(AssertionError):
            gn.batch(x)
    with pytest.raises(AssertionError):
        gn.batch(dict(graphs=ADJ, ef=f(10, 5, 2), nf=f(5, 3, 2)))          # missing key (src/batch.jl:54)
    with pytest.raises(AssertionError):
        gn.GNBlock((0, 0, 0), (1, 1, 1))
    with pytest.raises(AssertionError):
        gn.GNBlock((1, 1, 1), (0, 0, 0))
    with pytest.raises(AssertionError):
        gn.GNCore((3, 0, 5))        # GNFeedForward: all(dims .> 0) (src/gnfeedforward.jl:18)


def test_layer_fields_and_shapes(gn):
    b = gn.GNBlock((10, 5, 0), (3, 4, 5))
    assert b.edgefn[0].weight.shape == (3, 20) and b.nodefn[0].weight.shape == (4, 8)
    assert b.graphfn[0].weight.shape == (5, 7) and b.dropout.p == 0
    assert (b.edgefn[0].bias == 0).all()
    lim = np.sqrt(6 / 23)
    assert np.abs(b.edgefn[0].weight).max() <= lim
    c = gn.GNCore((3, 4, 5))
    assert c.block.edgefn[0].weight.shape == (3, 3 + 8 + 5)
    assert c.ffwd.eff[0].weight.shape == (12, 3) and c.ffwd.eff[1].weight.shape == (3, 12)
    assert c.gn1.nodeln.scale.shape == (4,) and c.gn2.graphln.eps == pytest.approx(1e-5)
    cl = gn.GNCoreList([c, gn.GNCore((3, 4, 5))])
    assert len(cl.list) == 2


def test_shard_ranges(gn):
    r = gn.shard_ranges([512] * 4096, 8)
    assert r == [(i * 512, (i + 1) * 512) for i in range(8)]
    rng = np.random.default_rng(0)
    m = rng.integers(64, 4097, size=1000)
    for ws in (1, 2, 3, 4, 8):
        r = gn.shard_ranges(m, ws)
        assert r[0][0] == 0 and r[-1][1] == 1000 and all(r[i][1] == r[i + 1][0] for i in range(ws - 1))
        tot = [m[a:b].sum() for a, b in r]
        assert max(tot) - min(tot) <= 2 * m.max()
    assert gn.shard_ranges([], 2) == [(0, 0), (0, 0)]
    assert gn.shard_ranges([0, 0, 0, 0], 2) == [(0, 2), (2, 4)]


def test_default_precision_is_fp32(gn):
    """The reference computes in Float32: the drop-in answers at fp32 parity unless the caller opts in to the tensor path."""
    assert gn.get_precision() == "fp32"
    gn.set_precision("auto")
    assert gn.get_precision() == "auto"
    gn.set_precision("fp32")
    with pytest.raises(AssertionError):
        gn.set_precision("fp16")


# names the reference exports for this path (src/GraphNets.jl:12-50) and north_star lists
REFERENCE_EXPORTS = ["GNGraphBatch", "batch", "unbatch", "GNBlock", "zerodim2nothing", "GNCore", "GNCoreList", "efview", "nfview",
                     "gfview", "flatunpaddednf", "flatunpaddedef", "collapsef", "unpaddedcollapsedef", "flatunpaddedcollapsedef"]


def test_julia_shim_binds_only_declared_symbols_and_defines_the_reference_exports(gn):
    """julia/GraphNetsB200.jl cannot run here (no Julia toolchain): check statically that every symbol it ccalls is declared in
    include/gnb200.h and exported by the library, that it exports every name the reference exports for this path, and that every
    exported name is defined in the file."""
    src = open(os.path.join(ROOT, "julia", "GraphNetsB200.jl")).read()
    code = "\n".join(l.split("#")[0] if not l.lstrip().startswith("#") else "" for l in src.splitlines())
    called = set(re.findall(r"ccall\(\(:(gnb_[a-z0-9_]+),\s*LIB\)", code))
    declared = set(_header_symbols())
    assert called and called <= declared, "undeclared symbols in the Julia shim: %s" % sorted(called - declared)
    lib = ctypes.CDLL(gn.LIB_PATH)
    for n in called:
        assert hasattr(lib, n), n
    # the model path (the benchmarked one) and the edge-list entry are reachable from Julia
    for n in ("gnb_model_create", "gnb_model_forward", "gnb_corelist_forward", "gnb_core_forward", "gnb_block_forward",
              "gnb_graph_lower", "gnb_graph_from_coo", "gnb_pad_edges", "gnb_pad_nodes", "gnb_collapse_edges", "gnb_graph_export_host"):
        assert n in called, "%s is not bound by the Julia shim" % n
    m = re.search(r"^export\s+(.*?)^end", code, flags=re.S | re.M)
    assert m, "no export list"
    exported = set(re.findall(r"[A-Za-z_][A-Za-z0-9_!]*", m.group(1)))
    missing = [n for n in REFERENCE_EXPORTS if n not in exported]
    assert not missing, "reference exports missing from the Julia shim: %s" % missing
    for n in exported:
        pat = r"(^|\s)(function\s+%s\b|(mutable\s+)?struct\s+%s\b|%s\([^)]*\)(\s+where\s+\{[^}]*\})?\s*=|const\s+%s\b)" % ((re.escape(n),) * 4)
        assert re.search(pat, code, flags=re.M), "%s is exported but not defined" % n
    # reference field names of the layer structs (src/gnblock.jl:39-45, src/gncore.jl:38-44, src/gncorelist.jl:29-31)
    for pat in (r"struct GNBlock\s+edgefn; nodefn; graphfn; dropout", r"struct GNCore\s+block; ffwd; gn1; gn2",
                r"struct GNCoreList\s+list", r"struct GNFeedForward\s+eff; nff; gff", r"struct GNGraphNorm\s+edgeln; nodeln; graphln"):
        assert re.search(pat, code), pat


def test_trainer_layer_conversion_roundtrip(gn):
    """train.layers_of: layer objects -> the oracle-format parameter list the Trainer (and the oracle) consume (no GPU needed)."""
    import workloads as W
    from oracle import gn_oracle as O
    layers = W.model_params("cfg2")
    model = W.to_gn_model(gn, layers)
    back = gn.pkg.train.layers_of(model)
    assert [k for k, _ in back] == [k for k, _ in layers]
    for (_, a), (_, b) in zip(back, layers):
        if "We" in a:
            for k in ("We", "be", "Wn", "bn", "Wg", "bg"):
                assert np.array_equal(a[k], b[k])
        else:
            assert np.array_equal(a["block"]["We"], b["block"]["We"])
            for i in range(3):
                for k in ("W1", "b1", "W2", "b2"):
                    assert np.array_equal(a["ffn"][i][k], b["ffn"][i][k])
                assert np.array_equal(a["ln1"][i]["gamma"], b["ln1"][i]["gamma"]) and np.array_equal(a["ln2"][i]["beta"], b["ln2"][i]["beta"])
    # the converted list drives the oracle to the same result as the original one
    w = W.make_workload("cfg2", B=3)
    g = O.lower(W.adj_list(w))
    ef, nf, gf = W.compact_inputs(w)
    y0, y1 = O.forward_sparse(layers, g, ef, nf, gf), O.forward_sparse(back, g, ef, nf, gf)
    assert all(np.allclose(p, q, rtol=1e-12, atol=1e-12) for p, q in zip(y0, y1))      # (the weights' memory order differs: BLAS blocking)
