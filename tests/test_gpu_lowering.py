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
