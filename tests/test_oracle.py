"""Pins the CPU oracle against every known-answer the reference's own tests hold for the path
(SURVEY.md 8c): golden broadcaster matrices (test/runtests.jl:480-508, 655-682), the collapse
identity (:41-50), batch/unbatch round trip (:362-365, :386-389), batch invariance (:111-115),
plus dense-mirror == sparse-oracle equivalence and the committed golden fixtures."""
import glob
import os

import numpy as np
import pytest

from oracle import gn_oracle as O
import workloads as W
from tests.golden.make_golden import unflatten_params

GOLD = os.path.join(os.path.dirname(__file__), "golden")

ADJ1 = np.array([[1, 0, 1], [1, 1, 0], [0, 0, 1]], np.float32)
ADJ2 = np.array([[0, 1, 0], [0, 0, 1], [1, 1, 0]], np.float32)


def test_golden_node2edge_broadcasters():
    # test/runtests.jl:480-508 (commented-out known-answer test)
    d = O.DenseBatch([ADJ1, ADJ2])
    assert d.src.shape == (2, 3, 9)
    assert (d.src[0] == np.array([[1, 0, 0, 0, 0, 0, 1, 0, 0], [0, 1, 0, 0, 1, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0, 0, 0, 1]])).all()
    assert (d.src[1] == np.array([[0, 0, 0, 1, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0, 0, 1, 0], [0, 0, 1, 0, 0, 1, 0, 0, 0]])).all()
    assert (d.dst[0] == np.array([[1, 1, 0, 0, 0, 0, 0, 0, 0], [0, 0, 0, 0, 1, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0, 1, 0, 1]])).all()
    assert (d.dst[1] == np.array([[0, 0, 1, 0, 0, 0, 0, 0, 0], [0, 0, 0, 1, 0, 1, 0, 0, 0], [0, 0, 0, 0, 0, 0, 0, 1, 0]])).all()


def test_golden_edge2node_broadcaster():
    # test/runtests.jl:655-682
    d = O.DenseBatch([ADJ1, ADJ2])
    e1 = np.array([[1, 0, 0], [1, 0, 0], [0, 0, 0], [0, 0, 0], [0, 1, 0], [0, 0, 0], [0, 0, 1], [0, 0, 0], [0, 0, 1]])
    e2 = np.array([[0, 0, 0], [0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 0], [0, 1, 0], [0, 0, 0], [0, 0, 1], [0, 0, 0]])
    assert (d.e2n[0] == e1).all() and (d.e2n[1] == e2).all()


def test_lowering_matches_broadcasters():
    """The COO index is exactly the one-hot content of the reference broadcasters."""
    rng = np.random.default_rng(5)
    adjs = [(rng.random((n, n)) < 0.4).astype(np.float32) for n in (4, 2, 6, 1, 5)]
    g, d = O.lower(adjs), O.DenseBatch(adjs)
    for e in range(g["E"]):
        b, k = g["edge_graph"][e], g["edge_slot"][e]
        base = g["graph_node_ptr"][b]
        assert d.src[b, g["edge_src"][e] - base, k] == 1 and d.src[b, :, k].sum() == 1
        assert d.dst[b, g["edge_dst"][e] - base, k] == 1 and d.dst[b, :, k].sum() == 1
        assert d.e2n[b, k, g["edge_dst"][e] - base] == 1
        assert d.g2e[b, 0, k] == 1 and d.e2g[b, k, 0] == 1
    assert g["E"] == int(sum(d.g2e[b].sum() for b in range(d.B)))
    # receiver-sorted within each graph, slots ascending
    for b in range(g["B"]):
        s = slice(g["graph_edge_ptr"][b], g["graph_edge_ptr"][b + 1])
        assert (np.diff(g["edge_slot"][s]) > 0).all()
        assert (np.diff(g["edge_dst"][s]) >= 0).all()
    # CSR over receivers
    for v in range(g["N"]):
        assert (g["edge_dst"][g["node_in_ptr"][v]:g["node_in_ptr"][v + 1]] == v).all()


def test_readme_graph_index():
    g = O.lower([ADJ1])
    assert g["edge_slot"].tolist() == [0, 1, 4, 6, 8]
    assert g["edge_src"].tolist() == [0, 1, 1, 0, 2]
    assert g["edge_dst"].tolist() == [0, 0, 1, 2, 2]
    assert g["node_in_ptr"].tolist() == [0, 2, 3, 5]


PATTERNS = [((3, 2, 4), (5, 4, 3)), ((3, 2, 0), (5, 4, 3)), ((3, 0, 4), (5, 4, 3)), ((3, 0, 0), (5, 4, 3)),
            ((0, 2, 4), (5, 4, 3)), ((0, 2, 0), (5, 4, 3)), ((0, 0, 4), (5, 4, 3)),
            ((3, 2, 4), (0, 4, 3)), ((3, 2, 4), (5, 0, 3)), ((3, 2, 4), (5, 4, 0)), ((3, 2, 0), (0, 0, 2))]


@pytest.mark.parametrize("din,dout", PATTERNS)
def test_dense_mirror_equals_sparse_block(din, dout):
    """Every `Nothing` method of getedgefninput / getnodefninput / getgraphfninput
    (src/edgefninput.jl:1-48, src/nodefninput.jl:1-25, src/graphfninput.jl:1-14)."""
    rng = np.random.default_rng(11)
    adjs = [(rng.random((n, n)) < 0.5).astype(np.float32) for n in (3, 5, 4)]
    g, d = O.lower(adjs), O.DenseBatch(adjs)
    p = W.block_params(rng, din, dout)
    ef = rng.random((g["E"], din[0]), dtype=np.float32) if din[0] else None
    nf = rng.random((g["N"], din[1]), dtype=np.float32) if din[1] else None
    gf = rng.random((g["B"], din[2]), dtype=np.float32) if din[2] else None
    ys = O.gnblock_sparse(p, g, ef, nf, gf)
    ep, npz = g["graph_edge_ptr"], g["graph_node_ptr"]
    efp = None if ef is None else O.padef(adjs, [ef[ep[b]:ep[b + 1]] for b in range(3)], din[0])
    nfp = None if nf is None else O.padnf(adjs, [nf[npz[b]:npz[b + 1]] for b in range(3)], din[1])
    # garbage in padded node slots must not leak
    if nfp is not None:
        for b, a in enumerate(adjs):
            nfp[b, a.shape[0]:, :] = 1e3
    gfp = None if gf is None else gf[:, None, :]
    yd = O.gnblock_dense(p, d, efp, nfp, gfp)
    for i, (s, dd) in enumerate(zip(ys, yd)):
        assert (s is None) == (dd is None) == (dout[i] == 0)
        if s is None:
            continue
        got = [lambda: np.concatenate(O.unpadef(adjs, dd)), lambda: np.concatenate(O.unpadnf(adjs, dd)),
               lambda: dd[:, 0, :]][i]()
        assert O.rel_err(got, s) < 1e-5


def test_dense_mirror_equals_sparse_core_stack():
    rng = np.random.default_rng(3)
    adjs = [(rng.random((n, n)) < 0.5).astype(np.float32) for n in (4, 6)]
    g, d = O.lower(adjs), O.DenseBatch(adjs)
    dims = (8, 6, 5)
    layers = [("core", W.core_params(rng, dims)), ("core", W.core_params(rng, dims))]
    ef = rng.random((g["E"], 8), dtype=np.float32)
    nf = rng.random((g["N"], 6), dtype=np.float32)
    gf = rng.random((g["B"], 5), dtype=np.float32)
    for mode in (0, 1, 2):
        ys = O.forward_sparse(layers, g, ef, nf, gf, eps_mode=mode)
        ep, npz = g["graph_edge_ptr"], g["graph_node_ptr"]
        efp = O.padef(adjs, [ef[ep[b]:ep[b + 1]] for b in range(2)], 8)
        nfp = O.padnf(adjs, [nf[npz[b]:npz[b + 1]] for b in range(2)], 6)
        yd = O.forward_dense(layers, d, efp, nfp, gf[:, None, :], eps_mode=mode)
        assert O.rel_err(np.concatenate(O.unpadef(adjs, yd[0])), ys[0]) < 1e-5
        assert O.rel_err(np.concatenate(O.unpadnf(adjs, yd[1])), ys[1]) < 1e-5
        assert O.rel_err(yd[2][:, 0, :], ys[2]) < 1e-5


def test_pad_unpad_roundtrip_exact():
    # batch_inverse_2D (test/runtests.jl:328-366): exact ==
    adj2 = np.array([[1, 0, 1, 0], [1, 1, 0, 1], [0, 0, 1, 0], [1, 1, 0, 1]], np.float32)
    adjs = [ADJ1, adj2]
    rng = np.random.default_rng(0)
    efs = [rng.random((int(a.sum()), 10), dtype=np.float32) for a in adjs]
    nfs = [rng.random((a.shape[0], 5), dtype=np.float32) for a in adjs]
    efp, nfp = O.padef(adjs, efs, 10), O.padnf(adjs, nfs, 5)
    assert efp.shape == (2, 16, 10) and nfp.shape == (2, 4, 5)
    for a, b in zip(O.unpadef(adjs, efp), efs):
        assert (a == b).all()
    for a, b in zip(O.unpadnf(adjs, nfp), nfs):
        assert (a == b).all()


def test_collapse_identity():
    # "Test edge collapsing" (test/runtests.jl:4-59): fully connected 2- and 3-node graphs, PN = 3
    rng = np.random.default_rng(2)
    A, Bm = np.ones((2, 2), np.float32), np.ones((3, 3), np.float32)
    adjs = [A, Bm]
    efp = rng.random((2, 9, 2), dtype=np.float32)      # padded (D=2, PE=9, B=2) as [B][PE][D]
    col = O.collapsef_dense(efp, 3)
    idxs = O.collapsed_edge_idxs(O.padadjmats(adjs))
    flat = np.concatenate([col[b][idxs[b]] for b in range(2)])    # flatunpaddedcollapsedef
    e = lambda k, b: efp[b - 1, k - 1]                               # 1-based like the reference test
    assert np.allclose(flat[0], e(1, 1))
    assert np.allclose(flat[1], (e(2, 1) + e(4, 1)) / 2)
    assert np.allclose(flat[2], e(5, 1))
    assert np.allclose(flat[3], e(1, 2))
    assert np.allclose(flat[4], (e(2, 2) + e(4, 2)) / 2)
    assert np.allclose(flat[5], (e(3, 2) + e(7, 2)) / 2)
    assert np.allclose(flat[6], e(5, 2))
    assert np.allclose(flat[7], (e(6, 2) + e(8, 2)) / 2)
    assert np.allclose(flat[8], e(9, 2))
    assert flat.shape == (9, 2)


def test_batch_invariance_oracle():
    # "GNBlock batch invariance" (test/runtests.jl:62-116)
    rng = np.random.default_rng(9)
    A, Bm = np.ones((2, 2), np.float32), np.ones((3, 3), np.float32)
    layers = [("block", W.block_params(rng, (0, 2, 0), (2, 2, 2))), ("block", W.block_params(rng, (2, 2, 2), (2, 2, 2)))]
    nfA, nfB = rng.random((2, 2), dtype=np.float32), rng.random((3, 2), dtype=np.float32)
    y1 = O.forward_sparse(layers, O.lower([A]), None, nfA, None)
    g2 = O.lower([A, Bm])
    y2 = O.forward_sparse(layers, g2, None, np.concatenate([nfA, nfB]), None)
    assert np.allclose(y1[0], y2[0][:4]) and np.allclose(y1[1], y2[1][:2]) and np.allclose(y1[2], y2[2][:1])


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "*.npz"))))
def test_oracle_reproduces_golden_fixtures(path):
    z = np.load(path)
    layers = unflatten_params(z)
    adjs = [z["adj_%d" % b] for b in range(int(z["n_graphs"]))]
    g = O.lower(adjs)
    for k in ("edge_src", "edge_dst", "edge_slot", "graph_edge_ptr", "graph_node_ptr", "node_in_ptr"):
        assert (g[k] == z["idx_" + k]).all()
    get = lambda k: z[k] if k in z else None
    ys = O.forward_sparse(layers, g, get("ef"), get("nf"), get("gf"), eps_mode=int(z["eps_mode"]))
    for y, k in zip(ys, ("ye", "yn", "yg")):
        if y is None:
            assert k not in z
        else:
            assert O.rel_err(y, z[k]) < 1e-12


def test_layernorm_eps_variants_differ_at_tolerance():
    x = np.random.default_rng(0).random((4, 16))
    g, b = np.ones(16), np.zeros(16)
    y0, y1, y2 = (O.layernorm(x, g, b, 1e-5, m) for m in (0, 1, 2))
    assert 0 < np.abs(y0 - y1).max() < 1e-3 and 0 < np.abs(y0 - y2).max() < 1e-3


def test_canonical_work_matches_survey():
    # cfg4 = 3.709 TFLOP, 12.33 GB; one GNCore = 920.6 GFLOP, 2.420 GB (+17.8 MB idx)
    E, N, G = 2097152, 262144, 4096
    fl, by = W.canonical_work(W.model_params("cfg4"), E, N, G)
    assert abs(fl / 1e12 - 3.709) < 0.01
    assert abs(by / 1e9 - 12.33) < 0.15
    fl1, by1 = W.canonical_work([("core", dict(dims=(128, 128, 128)))], E, N, G)
    assert abs(fl1 / 1e9 - 920.6) < 1.0 and abs(by1 / 1e9 - 2.438) < 0.01
