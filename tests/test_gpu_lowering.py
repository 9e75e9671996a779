"""-m gpu: batch lowering (adjacency -> receiver-sorted COO/CSR) is BIT-EXACT against the oracle,
through the C ABI (gnb_graph_lower / gnb_graph_export_host / gnb_pad_* / gnb_unpad_*)."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import gn_oracle as O

pytestmark = pytest.mark.gpu
KEYS = ("edge_src", "edge_dst", "edge_slot", "edge_graph", "graph_edge_ptr", "graph_node_ptr", "node_in_ptr")


def _check_index(gb, adjs):
    g = O.lower(adjs)
    idx = gb.index()
    assert gb.E == g["E"] and gb.N == g["N"] and gb.B == g["B"] and gb.node_block_size == g["PN"]
    for k in KEYS:
        assert np.array_equal(idx[k], g[k].astype(np.int32)), k


def test_readme_graphs(gn):
    a1 = np.array([[1, 0, 1], [1, 1, 0], [0, 0, 1]])
    a2 = np.array([[0, 1, 0], [0, 0, 1], [1, 1, 0]])
    gb = gn.GNGraphBatch([a1, a2])
    _check_index(gb, [a1, a2])
    idx = gb.index()
    # golden (test/runtests.jl:480-508): senders / receivers per active slot
    assert idx["edge_slot"].tolist() == [0, 1, 4, 6, 8, 2, 3, 5, 7]
    assert idx["edge_src"].tolist() == [0, 1, 1, 0, 2, 5, 3, 5, 4]
    assert idx["edge_dst"].tolist() == [0, 0, 1, 2, 2, 3, 4, 4, 5]


@pytest.mark.parametrize("seed,sizes,p", [(0, (1, 2, 3, 4, 5), 0.5), (1, (33, 7, 64, 1, 40), 0.3),
                                          (2, (70, 100, 3), 0.1), (3, (5, 5, 5), 0.0), (4, (6, 2), 1.0)])
def test_random_variable_batches(gn, seed, sizes, p):
    rng = np.random.default_rng(seed)
    adjs = [(rng.random((n, n)) < p).astype(np.uint8) for n in sizes]
    _check_index(gn.GNGraphBatch(adjs), adjs)


def test_non_binary_adjacency_uses_isone(gn):
    # only entries equal to one are edges (isone, src/pad.jl:30)
    a = np.array([[2.0, 1.0, 0.5], [1.0, 0.0, 1.0], [3.0, 1.0, 1.0]], np.float32)
    _check_index(gn.GNGraphBatch([a]), [a])


def test_single_adjacency_mode_replicates_structure(gn):
    rng = np.random.default_rng(7)
    a = (rng.random((16, 16)) < 0.25).astype(np.uint8)
    B = 37
    gb = gn.GNGraphBatch([a], B=B)
    _check_index(gb, [a] * B)
    assert gb.single


def test_full_size_properties(gn):
    """cfg4-shaped structure at full per-graph size, B=512: size-independent properties
    (sortedness, CSR consistency, slot <-> (src,dst) bijection, counts)."""
    import workloads as W
    rng = np.random.default_rng(4)
    adj = W.random_cells_adj(rng, 512, 64, 512)
    gb = gn.GNGraphBatch(None, _stacked=adj)
    idx = gb.index()
    assert gb.E == 512 * 512 and gb.N == 512 * 64
    assert (np.diff(idx["graph_edge_ptr"]) == 512).all() and (np.diff(idx["graph_node_ptr"]) == 64).all()
    assert (np.diff(idx["edge_dst"]) >= 0).all()                       # globally receiver-sorted
    e = np.arange(gb.E)
    b = idx["edge_graph"]
    assert (b == e // 512).all()
    i, j = idx["edge_src"] - 64 * b, idx["edge_dst"] - 64 * b
    assert (idx["edge_slot"] == i + 64 * j).all()
    assert (adj[b, i, j] == 1).all()                                    # every listed edge is active
    cnt = np.bincount(idx["edge_dst"], minlength=gb.N)
    assert (np.diff(idx["node_in_ptr"]) == cnt).all()
    # slots strictly ascending within a graph
    d = np.diff(idx["edge_slot"].astype(np.int64))
    assert (d[np.diff(b) == 0] > 0).all()


def test_pad_unpad_roundtrip_exact(gn):
    rng = np.random.default_rng(3)
    adjs = [(rng.random((n, n)) < 0.5).astype(np.uint8) for n in (3, 6, 4)]
    g = O.lower(adjs)
    efs = [rng.random((10, int(a.sum())), dtype=np.float32) for a in adjs]
    nfs = [rng.random((5, a.shape[0]), dtype=np.float32) for a in adjs]
    x = gn.batch(dict(graphs=adjs, ef=efs, nf=nfs, gf=None))
    assert x.ef.shape == (10, 36, 3) and x.nf.shape == (5, 6, 3)
    ref_e = O.padef(adjs, [e.T for e in efs], 10)      # [B][PE][D]
    ref_n = O.padnf(adjs, [n.T for n in nfs], 5)
    assert np.array_equal(x.ef.numpy(), ref_e.transpose(2, 1, 0))
    assert np.array_equal(x.nf.numpy(), ref_n.transpose(2, 1, 0))
    # unpad through the ABI
    eng = x.graphs.engine
    lib = gn.lib
    pe = x.ef.padded().permute(2, 1, 0).contiguous()
    out = torch.empty_like(x.ef.compact)
    assert lib.gnb_unpad_edges(eng.ctx, x.graphs.handle, C.c_void_p(pe.data_ptr()), 10, C.c_void_p(out.data_ptr())) == 0
    assert torch.equal(out, x.ef.compact)
    pn = x.nf.padded().permute(2, 1, 0).contiguous()
    outn = torch.empty_like(x.nf.compact)
    assert lib.gnb_unpad_nodes(eng.ctx, x.graphs.handle, C.c_void_p(pn.data_ptr()), 5, C.c_void_p(outn.data_ptr())) == 0
    assert torch.equal(outn, x.nf.compact)
    # batch |> unbatch == identity, exact (test/runtests.jl:328-366)
    u = gn.unbatch(x)
    for a, b in zip(u.ef, efs):
        assert np.array_equal(a.cpu().numpy(), b)
    for a, b in zip(u.nf, nfs):
        assert np.array_equal(a.cpu().numpy(), b)
    assert u.gf is None and u.graphs is x.graphs.adj_mats


def test_lower_rejects_bad_arguments(gn):
    eng = gn.get_engine()
    h = C.c_void_p()
    nn = (C.c_int32 * 1)(5)
    buf = np.zeros((4, 4), np.uint8)
    rc = gn.lib.gnb_graph_lower(eng.ctx, buf.ctypes.data_as(C.c_void_p), 1, 0, nn, 4, 1, 1, C.byref(h))
    assert rc == -1 and b"n_nodes" in gn.lib.gnb_last_error()
    rc = gn.lib.gnb_graph_lower(eng.ctx, buf.ctypes.data_as(C.c_void_p), 1, 0, nn, 4, 2, 3, C.byref(h))
    assert rc == -1


def _random_batch(rng, B, nmax):
    adjs = []
    for _ in range(B):
        n = int(rng.integers(1, nmax + 1))
        adjs.append((rng.random((n, n)) < rng.uniform(0.0, 0.9)).astype(np.uint8))
    adjs[0] = np.zeros((3, 3), np.uint8)      # an empty graph in front
    return adjs


def _coo_of(adjs, PN):
    src, dst, ep = [], [], [0]
    for a in adjs:
        j, i = np.nonzero(a.T == 1)          # receiver-major: ascending i + PN*j
        src.append(i)
        dst.append(j)
        ep.append(ep[-1] + i.size)
    return np.concatenate(src).astype(np.int32), np.concatenate(dst).astype(np.int32), np.array(ep, np.int32)


def test_coo_lowering_equals_dense_lowering(gn):
    """the index from COO edge lists == the index from the dense adjacency (== the oracle), bit for bit, and a
    forward on it gives the same features."""
    rng = np.random.default_rng(41)
    adjs = _random_batch(rng, 37, 19)
    PN = max(a.shape[0] for a in adjs)
    src, dst, ep = _coo_of(adjs, PN)
    nn = np.array([a.shape[0] for a in adjs], np.int32)
    gd = gn.GNGraphBatch(adjs)
    gc = gn.GNGraphBatch.from_coo(src, dst, ep, nn)
    ref = O.lower(adjs)
    assert (gc.E, gc.N, gc.B, gc.node_block_size) == (gd.E, gd.N, gd.B, gd.node_block_size) == (ref["E"], ref["N"], ref["B"], PN)
    for k in ("edge_src", "edge_dst", "edge_slot", "edge_graph", "graph_edge_ptr", "graph_node_ptr", "node_in_ptr"):
        assert np.array_equal(gc.index()[k], gd.index()[k]), k
        assert np.array_equal(gc.index()[k], ref[k].astype(np.int32)), k
    # the reference-facing fields are still there (built on demand)
    assert all(np.array_equal(a, b) for a, b in zip(gc.adj_mats, adjs))
    ef = rng.random((gd.E, 4), dtype=np.float32)
    nf = rng.random((gd.N, 3), dtype=np.float32)
    blk = gn.GNBlock((4, 3, 0), (5, 6, 7), rng=np.random.default_rng(3))
    xd = gn.GNData(gd, gn.Padded("e", torch.from_numpy(ef).cuda(), gd), gn.Padded("n", torch.from_numpy(nf).cuda(), gd), None)
    xc = gn.batch_coo(src, dst, ep, nn, ef=ef, nf=nf)
    yd, yc = blk(xd), blk(xc)
    for f in ("ef", "nf", "gf"):
        assert torch.equal(getattr(yd, f).compact, getattr(yc, f).compact), f


def test_coo_lowering_rejects_bad_edge_lists(gn):
    nn = np.array([3, 2], np.int32)
    ep = np.array([0, 2, 3], np.int32)
    ok = (np.array([0, 2, 1], np.int32), np.array([0, 1, 1], np.int32))
    gn.GNGraphBatch.from_coo(ok[0], ok[1], ep, nn)
    for src, dst in [(np.array([2, 0, 1], np.int32), np.array([1, 0, 1], np.int32)),      # not receiver-major
                     (np.array([0, 0, 1], np.int32), np.array([0, 0, 1], np.int32)),      # duplicate edge
                     (np.array([0, 3, 1], np.int32), np.array([0, 1, 1], np.int32)),      # sender out of range
                     (np.array([0, 2, 1], np.int32), np.array([0, 1, 2], np.int32))]:     # receiver out of range (graph of 2 nodes)
        with pytest.raises(AssertionError):
            gn.GNGraphBatch.from_coo(src, dst, ep, nn)


def test_bit_packed_adjacency_equals_byte_mask(gn):
    """GNB_ADJ_BITS (what `batch` uploads) against the uint8 / float32 / int32 entry points of gnb_graph_lower, PN not a multiple
    of 32 bits per graph."""
    import ctypes as C
    rng = np.random.default_rng(43)
    B, PN = 21, 7
    adj = (rng.random((B, PN, PN)) < 0.4).astype(np.uint8)
    nn = np.full(B, PN, np.int32)
    eng = gn.get_engine()
    lib, L = gn.lib, gn.pkg._lib
    mask = np.ascontiguousarray(adj.transpose(0, 2, 1))
    outs = []
    for dtype, arr in ((L.ADJ_U8, mask), (L.ADJ_F32, mask.astype(np.float32)), (L.ADJ_I32, mask.astype(np.int32)),
                       (L.ADJ_BITS, gn.pack_adjacency_bits(mask))):
        h = C.c_void_p()
        eng.bind_stream()
        L.check(lib.gnb_graph_lower(eng.ctx, arr.ctypes.data_as(C.c_void_p), dtype, 0, nn.ctypes.data_as(L.i32p), PN, B, B, C.byref(h)))
        E = C.c_int64()
        L.check(lib.gnb_graph_counts(h, C.byref(E), None, None, None))
        src, dst, slot = (np.empty(E.value, np.int32) for _ in range(3))
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        L.check(lib.gnb_graph_export_host(eng.ctx, h, p(src), p(dst), p(slot), None, None, None, None))
        lib.gnb_graph_destroy(h)
        outs.append((E.value, src, dst, slot))
    assert outs[0][0] == int(adj.sum())
    for o in outs[1:]:
        assert o[0] == outs[0][0] and all(np.array_equal(a, b) for a, b in zip(o[1:], outs[0][1:]))
