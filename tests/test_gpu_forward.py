"""-m gpu: forward parity of the CUDA path (through the C ABI) against the float64 oracle and the
committed golden fixtures.  fp32 path: rel err <= 1e-5; bf16 tensor path: <= 1e-2."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import gn_oracle as O
import workloads as W
from tests.golden.make_golden import unflatten_params
from tests.gpu_util import run_product, run_oracle, assert_parity, FP32_TOL, BF16_TOL
from tests.test_oracle import PATTERNS

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "*.npz"))))
def test_golden_fixtures(gn, path):
    z = np.load(path)
    layers = unflatten_params(z)
    adjs = [z["adj_%d" % b] for b in range(int(z["n_graphs"]))]
    get = lambda k: z[k] if k in z else None
    g = O.lower(adjs)
    ep, npz = g["graph_edge_ptr"], g["graph_node_ptr"]
    B = g["B"]
    ef, nf, gf = get("ef"), get("nf"), get("gf")
    x = gn.batch(dict(graphs=adjs,
                      ef=None if ef is None else [ef[ep[b]:ep[b + 1]].T for b in range(B)],
                      nf=None if nf is None else [nf[npz[b]:npz[b + 1]].T for b in range(B)],
                      gf=None if gf is None else [gf[b] for b in range(B)]))
    idx = x.graphs.index()
    for k in ("edge_src", "edge_dst", "edge_slot", "graph_edge_ptr", "graph_node_ptr", "node_in_ptr"):
        assert np.array_equal(idx[k], z["idx_" + k]), k
    model = W.to_gn_model(gn, layers, eps_mode=int(z["eps_mode"]))
    y = model(x, precision="fp32")
    c = lambda f: None if f is None else f.compact.cpu().numpy()
    assert_parity((c(y.ef), c(y.nf), c(y.gf)), (get("ye"), get("yn"), get("yg")), FP32_TOL, os.path.basename(path))


@pytest.mark.parametrize("din,dout", PATTERNS)
def test_block_nothing_patterns(gn, din, dout):
    """All 7 + 4 + 2 `Nothing` methods (src/edgefninput.jl, nodefninput.jl, graphfninput.jl) and
    zerodim2nothing (src/gnblock.jl:71-78)."""
    rng = np.random.default_rng(11)
    adjs = [(rng.random((n, n)) < 0.5).astype(np.uint8) for n in (3, 5, 4, 9)]
    layers = [("block", W.block_params(rng, din, dout))]
    w = dict(mode="vector", graphs=adjs,
             ef=[rng.random((din[0], int(a.sum())), dtype=np.float32) for a in adjs] if din[0] else None,
             nf=[rng.random((din[1], a.shape[0]), dtype=np.float32) for a in adjs] if din[1] else None,
             gf=[rng.random(din[2], dtype=np.float32) for a in adjs] if din[2] else None)
    _, y, got = run_product(gn, layers, w)
    _, ref = run_oracle(layers, w)
    assert_parity(got, ref, FP32_TOL, "%s=>%s" % (din, dout))
    assert (y.ef is None) == (dout[0] == 0) and (y.nf is None) == (dout[1] == 0) and (y.gf is None) == (dout[2] == 0)


@pytest.mark.parametrize("name,B", [("cfg1", 2), ("cfg2", 64), ("cfg2", 1024)])
def test_single_mode_configs(gn, name, B):
    w = W.make_workload(name, B=B)
    layers = W.model_params(name)
    _, _, got = run_product(gn, layers, w)
    _, ref = run_oracle(layers, w)
    assert_parity(got, ref, FP32_TOL, name)


@pytest.mark.parametrize("dims,sizes", [((16, 12, 8), (5, 9, 2)), ((33, 17, 70), (12, 3)), ((128, 128, 128), (20, 7)),
                                        ((64, 96, 32), (6, 6, 6))])
def test_core_stack_fp32(gn, dims, sizes):
    rng = np.random.default_rng(21)
    adjs = [(rng.random((n, n)) < 0.4).astype(np.uint8) for n in sizes]
    layers = [("block", W.block_params(rng, (3, 2, 0), dims)), ("core", W.core_params(rng, dims)),
              ("core", W.core_params(rng, dims)), ("block", W.block_params(rng, dims, (2, 3, 4)))]
    w = dict(mode="vector", graphs=adjs, ef=[rng.random((3, int(a.sum())), dtype=np.float32) for a in adjs],
             nf=[rng.random((2, a.shape[0]), dtype=np.float32) for a in adjs], gf=None)
    for mode in (0, 1, 2):
        _, _, got = run_product(gn, layers, w, eps_mode=mode)
        _, ref = run_oracle(layers, w, eps_mode=mode)
        assert_parity(got, ref, FP32_TOL, "dims=%s eps_mode=%d" % (dims, mode))


def test_cfg3_shape_small(gn):
    """examples/sort model shape: (0,100,0) => 384 x3, 2 cores, => (2,2,0); fully connected graphs."""
    w = W.make_workload("cfg3", B=6, n_nodes=(3, 12))
    layers = W.model_params("cfg3")
    _, y, got = run_product(gn, layers, w, precision="fp32")
    _, ref = run_oracle(layers, w)
    assert_parity(got, ref, FP32_TOL, "cfg3")
    assert y.gf is None


def test_cfg4_shape_small_fp32(gn):
    w = W.make_workload("cfg4", B=4)
    layers = W.model_params("cfg4")
    _, _, got = run_product(gn, layers, w, precision="fp32")
    _, ref = run_oracle(layers, w)
    assert_parity(got, ref, FP32_TOL, "cfg4 fp32")


def test_empty_and_isolated(gn):
    """Graphs with no edges, isolated nodes, a 1-node graph: aggregates are zero, nothing crashes."""
    rng = np.random.default_rng(5)
    adjs = [np.zeros((3, 3), np.uint8), np.array([[0, 1], [0, 0]], np.uint8), np.ones((1, 1), np.uint8),
            np.zeros((1, 1), np.uint8)]
    dims = (8, 8, 8)
    layers = [("block", W.block_params(rng, (2, 3, 0), dims)), ("core", W.core_params(rng, dims)),
              ("block", W.block_params(rng, dims, (1, 2, 3)))]
    w = dict(mode="vector", graphs=adjs, ef=[rng.random((2, int(a.sum())), dtype=np.float32) for a in adjs],
             nf=[rng.random((3, a.shape[0]), dtype=np.float32) for a in adjs], gf=None)
    _, _, got = run_product(gn, layers, w)
    _, ref = run_oracle(layers, w)
    assert_parity(got, ref, FP32_TOL, "empty/isolated")


def test_batch_invariance_bit_exact(gn):
    """test/runtests.jl:62-116 asserts `approx`; deterministic segmented sums make it BIT-exact here."""
    rng = np.random.default_rng(9)
    A, Bm = np.ones((2, 2), np.uint8), np.ones((3, 3), np.uint8)
    layers = [("block", W.block_params(rng, (0, 2, 0), (2, 2, 2))), ("block", W.block_params(rng, (2, 2, 2), (2, 2, 2)))]
    nfA, nfB = rng.random((2, 2), dtype=np.float32), rng.random((2, 3), dtype=np.float32)
    model = W.to_gn_model(gn, layers)
    y1 = model(gn.batch(dict(graphs=[A], ef=None, nf=[nfA], gf=None)), precision="fp32")
    yn = model(gn.batch(dict(graphs=[A, Bm], ef=None, nf=[nfA, nfB], gf=None)), precision="fp32")
    assert gn.nfview(y1, slice(None), slice(None), 0).shape == gn.nfview(yn, slice(None), slice(None), 0).shape
    assert torch.equal(gn.nfview(y1, slice(None), slice(None), 0), gn.nfview(yn, slice(None), slice(None), 0))
    assert torch.equal(gn.efview(y1, slice(None), slice(None), 0), gn.efview(yn, slice(None), slice(None), 0))
    assert torch.equal(gn.gfview(y1, slice(None), 0), gn.gfview(yn, slice(None), 0))


def test_sharded_equals_unsharded(gn):
    """graphs are independent, so running contiguous shards separately (as each rank of a
    multi-GPU job does) reproduces the single-run result bit-exactly - emulated on one GPU."""
    w = W.make_workload("cfg4", B=8, n_nodes=16, n_edges=40)
    layers = W.model_params("cfg2")   # (10,5,0) -> cores (10,5,3) -> (3,4,5)
    model = W.to_gn_model(gn, layers)
    full = model(gn.batch(W.as_batch_input(w)), precision="fp32")
    parts = []
    for r in range(2):
        shard, _ = gn.shard_batch(W.as_batch_input(w), r, 2)
        parts.append(model(gn.batch(shard), precision="fp32"))
    for f in ("ef", "nf", "gf"):
        cat = torch.cat([getattr(p, f).compact for p in parts])
        assert torch.equal(cat, getattr(full, f).compact), f


def test_forward_host_variant_matches_device(gn):
    import ctypes as C
    w = W.make_workload("cfg2", B=16)
    layers = W.model_params("cfg2")
    model = W.to_gn_model(gn, layers)
    x = gn.batch(W.as_batch_input(w))
    y = model(x, precision="fp32")
    eng = x.graphs.engine
    h = model._model(eng)
    ef, nf, _ = W.compact_inputs(w)
    oe = np.empty((x.graphs.E, 3), np.float32)
    on = np.empty((x.graphs.N, 4), np.float32)
    og = np.empty((x.graphs.B, 5), np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = gn.lib.gnb_model_forward_host(eng.ctx, h, x.graphs.handle, p(ef), p(nf), None, p(oe), p(on), p(og), 0)
    assert rc == 0, gn.lib.gnb_last_error()
    assert np.array_equal(oe, y.ef.compact.cpu().numpy()) and np.array_equal(on, y.nf.compact.cpu().numpy())
    assert np.array_equal(og, y.gf.compact.cpu().numpy())


@pytest.mark.parametrize("precision", ["auto", "fp32"])
def test_forward_host_pipelined_copies(gn, precision):
    """Large enough for the overlapped path of gnb_model_forward_host (edge rows uploaded in chunks behind which the
    encoder's edge kernel starts, edge output downloaded under the decoder's node / graph kernels): identical to the
    device-resident forward, twice in a row (buffer reuse)."""
    import ctypes as C
    w = W.make_workload("cfg4", B=512)
    layers = W.model_params("cfg4")
    model = W.to_gn_model(gn, layers)
    x = gn.batch(W.as_batch_input(w))
    y = model(x, precision=precision)
    eng = x.graphs.engine
    h = model._model(eng)
    ef, nf, _ = W.compact_inputs(w)
    assert ef.nbytes >= 8 << 20
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    for rep in range(2):
        oe = np.full((x.graphs.E, 3), np.nan, np.float32)
        on = np.full((x.graphs.N, 4), np.nan, np.float32)
        og = np.full((x.graphs.B, 5), np.nan, np.float32)
        rc = gn.lib.gnb_model_forward_host(eng.ctx, h, x.graphs.handle, p(ef), p(nf), None, p(oe), p(on), p(og),
                                           gn.pkg._lib.PRECISIONS[precision])
        assert rc == 0, gn.lib.gnb_last_error()
        assert np.array_equal(oe, y.ef.compact.cpu().numpy()), rep
        assert np.array_equal(on, y.nf.compact.cpu().numpy()), rep
        assert np.array_equal(og, y.gf.compact.cpu().numpy()), rep


@pytest.mark.parametrize("precision,side_stream", [("auto", False), ("auto", True), ("fp32", False)])
def test_forward_replayed_as_cuda_graph(gn, precision, side_stream):
    """gnb_model_forward captures the second forward with the same (model, graph, buffers) into a CUDA graph and replays it
    from then on (model.cu::forward_graphed): replays must equal the eager result bit for bit - with the context bound to
    torch's legacy default stream (captured on the library's side stream) and to a side stream - see new inputs written into
    the same buffers, and a call with different buffers in between must not disturb the cached graph."""
    import ctypes as C
    w = W.make_workload("cfg4", B=64, seed=11)
    layers = W.model_params("cfg4")
    model = W.to_gn_model(gn, layers)
    x = gn.batch(W.as_batch_input(w))
    eng = x.graphs.engine
    stream = torch.cuda.Stream() if side_stream else torch.cuda.current_stream()
    with torch.cuda.stream(stream):
        y0 = model(x, precision=precision)      # eager (new output buffers each call -> new keys)
        ref = [t.compact.clone() for t in (y0.ef, y0.nf, y0.gf)]
        h = model._model(eng)
        eng.bind_stream()
        outs = [torch.full_like(r, float("nan")) for r in ref]
        ptr = lambda t: C.c_void_p(t.data_ptr())
        prec = gn.pkg._lib.PRECISIONS[precision]
        call = lambda ef_t, o: gn.lib.gnb_model_forward(eng.ctx, h, x.graphs.handle, ptr(ef_t), ptr(x.nf.compact), None,
                                                        ptr(o[0]), ptr(o[1]), ptr(o[2]), prec)
        l0 = eng.launches
        for rep in range(5):      # 1: eager, 2: capture + replay, 3..: replay
            for o in outs: o.fill_(float("nan"))
            assert call(x.ef.compact, outs) == 0, gn.lib.gnb_last_error()
            eng.sync()
            for o, r in zip(outs, ref): assert torch.equal(o, r), (rep, precision)
            if rep == 2:      # another key in between
                other = [torch.empty_like(r) for r in ref]
                assert call(x.ef.compact, other) == 0
                eng.sync()
                for o, r in zip(other, ref): assert torch.equal(o, r)
        assert eng.launches > l0      # replays are accounted like eager launches
        # new data in the same input buffer: the graph reads the buffer, not a snapshot
        ef2 = x.ef.compact.clone()
        x.ef.compact.mul_(0.5)
        assert call(x.ef.compact, outs) == 0
        eng.sync()
        y_half = [o.clone() for o in outs]
        x.ef.compact.copy_(ef2)
        eager = [torch.empty_like(r) for r in ref]      # fresh buffers -> eager
        x_half = ef2 * 0.5
        assert call(x_half, eager) == 0
        eng.sync()
        for a_, b_ in zip(y_half, eager): assert torch.equal(a_, b_)
    torch.cuda.synchronize()


def test_single_layer_abi_entry_points(gn):
    """gnb_block_forward / gnb_core_forward / gnb_corelist_forward on caller-owned device weights."""
    import ctypes as C
    L = gn.pkg._lib
    rng = np.random.default_rng(2)
    adjs = [(rng.random((n, n)) < 0.5).astype(np.uint8) for n in (4, 6)]
    dims = (6, 5, 4)
    cores = [W.core_params(rng, dims), W.core_params(rng, dims)]
    w = dict(mode="vector", graphs=adjs, ef=[rng.random((6, int(a.sum())), dtype=np.float32) for a in adjs],
             nf=[rng.random((5, a.shape[0]), dtype=np.float32) for a in adjs],
             gf=[rng.random(4, dtype=np.float32) for a in adjs])
    x = gn.batch(W.as_batch_input(w))
    eng = x.graphs.engine
    dev = eng.torch_device
    keep = []

    def d(a):   # (out,in) numpy -> device column-major
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(a, np.float32).T)).to(dev)
        keep.append(t)
        return t.data_ptr()

    def blk(p):
        b = L.BlockParams()
        (b.in_e, b.in_n, b.in_g), (b.out_e, b.out_n, b.out_g) = p["din"], p["dout"]
        b.We, b.be, b.Wn, b.bn, b.Wg, b.bg = d(p["We"]), d(p["be"]), d(p["Wn"]), d(p["bn"]), d(p["Wg"]), d(p["bg"])
        return b
    arr = (L.CoreParams * 2)()
    for i, c in enumerate(cores):
        arr[i].block = blk(c["block"])
        for k in range(3):
            f = c["ffn"][k]
            arr[i].ffn[k].W1, arr[i].ffn[k].b1, arr[i].ffn[k].W2, arr[i].ffn[k].b2 = d(f["W1"]), d(f["b1"]), d(f["W2"]), d(f["b2"])
            for name in ("ln1", "ln2"):
                ln = getattr(arr[i], name)[k]
                ln.gamma, ln.beta, ln.eps, ln.eps_mode = d(c[name][k]["gamma"]), d(c[name][k]["beta"]), 1e-5, 0
    oe, on, og = eng.empty(x.graphs.E, 6), eng.empty(x.graphs.N, 5), eng.empty(x.graphs.B, 4)
    P = lambda t: C.c_void_p(t.data_ptr())
    eng.bind_stream()
    rc = gn.lib.gnb_corelist_forward(eng.ctx, x.graphs.handle, arr, 2, P(x.ef.compact), P(x.nf.compact),
                                     P(x.gf.compact), P(oe), P(on), P(og), 0)
    assert rc == 0, gn.lib.gnb_last_error()
    g = O.lower(adjs)
    ef, nf, gf = W.compact_inputs(w)
    ref = O.forward_sparse([("core", c) for c in cores], g, ef, nf, gf)
    assert_parity((oe.cpu().numpy(), on.cpu().numpy(), og.cpu().numpy()), ref, FP32_TOL, "corelist ABI")
    rc = gn.lib.gnb_core_forward(eng.ctx, x.graphs.handle, C.byref(arr[0]), P(x.ef.compact), P(x.nf.compact),
                                 P(x.gf.compact), P(oe), P(on), P(og), 0)
    assert rc == 0
    ref = O.forward_sparse([("core", cores[0])], g, ef, nf, gf)
    assert_parity((oe.cpu().numpy(), on.cpu().numpy(), og.cpu().numpy()), ref, FP32_TOL, "core ABI")
    b = blk(W.block_params(rng, dims, (2, 0, 3)))
    oe2, og2 = eng.empty(x.graphs.E, 2), eng.empty(x.graphs.B, 3)
    rc = gn.lib.gnb_block_forward(eng.ctx, x.graphs.handle, C.byref(b), P(x.ef.compact), P(x.nf.compact),
                                  P(x.gf.compact), P(oe2), None, P(og2), 0)
    assert rc == 0, gn.lib.gnb_last_error()
    # error behaviour: dimension chain mismatch -> GNB_ERR_INVALID
    arr[1].block.in_e = 7
    rc = gn.lib.gnb_corelist_forward(eng.ctx, x.graphs.handle, arr, 2, P(x.ef.compact), P(x.nf.compact),
                                     P(x.gf.compact), P(oe), P(on), P(og), 0)
    assert rc == -1


def test_linearity_property_full_size(gn):
    """Size-independent property at cfg4's full per-graph size (B=256): a GNBlock is affine, so
    f(x1) + f(x2) - f(0) == f(x1 + x2) up to rounding; checks the whole gather/aggregate path at scale."""
    rng = np.random.default_rng(4)
    adj = W.random_cells_adj(rng, 256, 64, 512)
    blk = gn.GNBlock((10, 5, 3), (16, 12, 8), rng=np.random.default_rng(1))
    E, N, B = 256 * 512, 256 * 64, 256
    mk = lambda r, d: rng.random((r, d), dtype=np.float32)
    xs = [(mk(E, 10), mk(N, 5), mk(B, 3)) for _ in range(2)]
    run = lambda e, n, g: blk(gn.batch_compact(adj, e, n, g), precision="fp32")
    y1, y2 = run(*xs[0]), run(*xs[1])
    y0 = run(np.zeros((E, 10), np.float32), np.zeros((N, 5), np.float32), np.zeros((B, 3), np.float32))
    y12 = run(*(a + b for a, b in zip(*xs)))
    for f in ("ef", "nf", "gf"):
        lhs = getattr(y1, f).compact.double() + getattr(y2, f).compact.double() - getattr(y0, f).compact.double()
        rhs = getattr(y12, f).compact.double()
        assert float((lhs - rhs).abs().max() / rhs.abs().max()) < 1e-5, f


def test_bf16_path_when_supported(gn):
    """Tensor-core path parity (1e-2) on a GNCore stack the tcgen05 path supports."""
    w = W.make_workload("cfg4", B=6)
    layers = W.model_params("cfg4")
    model = W.to_gn_model(gn, layers)
    x = gn.batch(W.as_batch_input(w))
    try:
        y = model(x, precision="bf16")
    except gn.pkg.GnbError as e:
        if "not supported" in str(e):
            pytest.skip("tcgen05 path not available for hidden=128 in this build: %s" % e)
        raise
    c = lambda f: None if f is None else f.compact.cpu().numpy()
    _, ref = run_oracle(layers, w)
    assert_parity((c(y.ef), c(y.nf), c(y.gf)), ref, BF16_TOL, "cfg4 bf16")
