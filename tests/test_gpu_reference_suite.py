"""-m gpu: the reference's own test suite (test/runtests.jl) restated against the product API.
Same testsets, same assertions (shapes, `nothing`, round trip, collapse identity); indices 0-based."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ADJ = np.array([[1, 0, 1], [1, 1, 0], [0, 0, 1]])
ADJ2 = np.array([[1, 0, 1, 0], [1, 1, 0, 1], [0, 0, 1, 0], [1, 1, 0, 1]])
rand = lambda *s: np.asfortranarray(np.random.default_rng(sum(s)).random(s, dtype=np.float32))
ALL = slice(None)


def test_edge_collapsing(gn):
    # test/runtests.jl:4-59
    enc, dec = gn.GNBlock((0, 2, 0), (2, 2, 2)), gn.GNBlock((2, 2, 2), (2, 2, 2))
    A, B = np.ones((2, 2), int), np.ones((3, 3), int)
    x = gn.batch(dict(graphs=[A, B], ef=None, nf=[rand(2, 2), rand(2, 3)], gf=None))
    y = dec(enc(x))
    flat = gn.flatunpaddedcollapsedef(y)
    ef = y.ef      # padded face (D, PE, B)
    ap = lambda a, b: torch.allclose(a, b, rtol=3.5e-4, atol=0)
    assert ap(flat[:, 0], ef[:, 0, 0])
    assert ap(flat[:, 1], (ef[:, 1, 0] + ef[:, 3, 0]) / 2)
    assert ap(flat[:, 2], ef[:, 4, 0])
    assert ap(flat[:, 3], ef[:, 0, 1])
    assert ap(flat[:, 4], (ef[:, 1, 1] + ef[:, 3, 1]) / 2)
    assert ap(flat[:, 5], (ef[:, 2, 1] + ef[:, 6, 1]) / 2)
    assert ap(flat[:, 6], ef[:, 4, 1])
    assert ap(flat[:, 7], (ef[:, 5, 1] + ef[:, 7, 1]) / 2)
    assert ap(flat[:, 8], ef[:, 8, 1])
    assert gn.collapsef(y).shape == (2, 6, 2)


def test_no_graph_features_output(gn):
    # test/runtests.jl:118-164
    block = gn.GNBlock((10, 5, 0), (3, 4, 0))
    x = gn.batch(dict(graphs=ADJ, ef=rand(10, 5, 2), nf=rand(5, 3, 2), gf=None))
    y = gn.unbatch(block(x))
    assert y.ef.shape == (3, 5, 2) and y.nf.shape == (4, 3, 2) and y.gf is None
    assert y.ef[:, :, 0].shape == (3, 5) and y.nf[:, :, 0].shape == (4, 3)


def test_readme_example_1(gn):
    # test/runtests.jl:180-216
    block = gn.GNBlock((10, 5, 0), (3, 4, 5))
    x = gn.batch(dict(graphs=ADJ, ef=rand(10, 5, 2), nf=rand(5, 3, 2), gf=None))
    assert isinstance(x.graphs, gn.GNGraphBatch)
    assert x.ef.shape == (10, x.graphs.edge_block_size, 2) and x.nf.shape == (5, x.graphs.node_block_size, 2)
    y = gn.unbatch(block(x))
    assert y.ef.shape == (3, 5, 2) and y.nf.shape == (4, 3, 2) and y.gf.shape == (5, 2)
    assert y.gf[:, 1].shape == (5,)


def test_readme_example_2(gn):
    # test/runtests.jl:218-271: graphs of different structure
    block = gn.GNBlock((10, 5, 0), (3, 4, 5))
    x = gn.batch(dict(graphs=[ADJ, ADJ2], ef=[rand(10, 5), rand(10, 9)], nf=[rand(5, 3), rand(5, 4)], gf=None))
    assert x.ef.shape == (10, 16, 2) and x.nf.shape == (5, 4, 2)
    yb = block(x)
    y = gn.unbatch(yb)
    assert gn.efview(yb, ALL, ALL, 0).shape == (3, 5) and gn.nfview(yb, ALL, ALL, 0).shape == (4, 3)
    assert gn.gfview(yb, ALL, 0).shape == (5,)
    assert gn.efview(yb, ALL, ALL, 1).shape == (3, 9) and gn.nfview(yb, ALL, ALL, 1).shape == (4, 4)
    assert gn.gfview(yb, ALL, 1).shape == (5,)
    assert y.ef[0].shape == (3, 5) and y.nf[0].shape == (4, 3) and y.gf[0].shape == (5,)
    assert y.ef[1].shape == (3, 9) and y.nf[1].shape == (4, 4) and y.gf[1].shape == (5,)
    assert yb.gf.shape == (5, 1, 2) and yb.ef.shape == (3, 16, 2)
    # views alias the batched storage (src/views.jl uses @view)
    v = gn.efview(yb, ALL, ALL, 1)
    v[0, 0] = 123.0
    assert float(yb.ef.compact[5, 0]) == 123.0


def test_readme_example_3(gn):
    # test/runtests.jl:273-324, but actually running encoder -> core_list -> decoder (SURVEY 4 gap)
    enc = gn.GNBlock((10, 5, 0), (10, 5, 3))
    cores = gn.GNCoreList([gn.GNCore((10, 5, 3)) for _ in range(2)])
    dec = gn.GNBlock((10, 5, 3), (3, 4, 5))
    x = gn.batch(dict(graphs=ADJ, ef=rand(10, 5, 2), nf=rand(5, 3, 2), gf=None))
    y = gn.unbatch(dec(cores(enc(x))))
    assert y.ef.shape == (3, 5, 2) and y.nf.shape == (4, 3, 2) and y.gf.shape == (5, 2)
    # fused single-call form gives the same numbers
    y2 = gn.unbatch(gn.GNSequential(enc, cores, dec)(x))
    assert torch.equal(y.ef, y2.ef) and torch.equal(y.nf, y2.nf) and torch.equal(y.gf, y2.gf)


def test_batch_inverse_2d(gn):
    # test/runtests.jl:328-366
    efs, nfs = [rand(10, 5), rand(10, 9)], [rand(5, 3), rand(5, 4)]
    x = dict(graphs=[ADJ, ADJ2], ef=efs, nf=nfs, gf=None)
    xh = gn.unbatch(gn.batch(x))
    assert xh.graphs is not None and all(np.array_equal(a, b) for a, b in zip(xh.graphs, x["graphs"]))
    assert all(np.array_equal(a.cpu().numpy(), b) for a, b in zip(xh.ef, efs))
    assert all(np.array_equal(a.cpu().numpy(), b) for a, b in zip(xh.nf, nfs))
    assert xh.gf is None


def test_batch_inverse_3d(gn):
    # test/runtests.jl:368-390
    ef, nf = rand(10, 5, 2), rand(5, 3, 2)
    xh = gn.unbatch(gn.batch(dict(graphs=ADJ, ef=ef, nf=nf, gf=None)))
    assert np.array_equal(xh.graphs, ADJ)
    assert np.array_equal(xh.ef.cpu().numpy(), ef) and np.array_equal(xh.nf.cpu().numpy(), nf) and xh.gf is None


def test_gnblock(gn):
    # test/runtests.jl:627-652
    y = gn.unbatch(gn.GNBlock((10, 5, 0), (3, 4, 5))(gn.batch(dict(graphs=ADJ, ef=rand(10, 5, 2), nf=rand(5, 3, 2), gf=None))))
    assert y.ef.shape == (3, 5, 2) and y.nf.shape == (4, 3, 2) and y.gf.shape == (5, 2)


def test_gncore(gn):
    # test/runtests.jl:685-709
    x = gn.batch(dict(graphs=ADJ, ef=rand(3, 5, 2), nf=rand(4, 3, 2), gf=rand(5, 2)))
    assert x.gf.shape == (5, 1, 2)
    y = gn.unbatch(gn.GNCore((3, 4, 5))(x))
    assert y.ef.shape == (3, 5, 2) and y.nf.shape == (4, 3, 2) and y.gf.shape == (5, 2)


def test_gncorelist(gn):
    # test/runtests.jl:711-735
    x = gn.batch(dict(graphs=ADJ, ef=rand(3, 5, 2), nf=rand(4, 3, 2), gf=rand(5, 2)))
    y = gn.unbatch(gn.GNCoreList([gn.GNCore((3, 4, 5)), gn.GNCore((3, 4, 5))])(x))
    assert y.ef.shape == (3, 5, 2) and y.nf.shape == (4, 3, 2) and y.gf.shape == (5, 2)


def test_single_graph_vector_unbatches_in_single_form(gn):
    # src/unbatch.jl:15-17: length(adj_mats) == 1 -> single-adjacency form
    x = gn.batch(dict(graphs=[ADJ], ef=[rand(10, 5)], nf=[rand(5, 3)], gf=None))
    y = gn.unbatch(x)
    assert y.ef.shape == (10, 5, 1) and y.nf.shape == (5, 3, 1)


def test_flat_unpadded_views(gn):
    x = gn.batch(dict(graphs=[ADJ, ADJ2], ef=[rand(10, 5), rand(10, 9)], nf=[rand(5, 3), rand(5, 4)], gf=None))
    assert gn.flatunpaddedef(x).shape == (10, 14) and gn.flatunpaddednf(x).shape == (5, 7)
    # same content as masking the padded tensor with the unpadders (src/views.jl:80-98)
    pe = x.ef.padded().permute(2, 1, 0).reshape(-1, 10).cpu().numpy()
    assert np.array_equal(pe[x.graphs.flat_edge_unpadder], gn.flatunpaddedef(x).t().cpu().numpy())
    pn = x.nf.padded().permute(2, 1, 0).reshape(-1, 5).cpu().numpy()
    assert np.array_equal(pn[x.graphs.flat_node_unpadder], gn.flatunpaddednf(x).t().cpu().numpy())


def test_logitcrossentropy_over_compact_views(gn):
    """examples/sort/sort.jl:76-78: loss = logitcrossentropy(flatunpaddednf(y), flatunpaddednf(t)) + the same over the edges."""
    import torch
    from oracle import gn_oracle as O
    rng = np.random.default_rng(61)
    adjs = [np.ones((n, n), np.uint8) for n in (2, 5, 3, 9, 4)]
    N, E = sum(a.shape[0] for a in adjs), sum(int(a.sum()) for a in adjs)
    for D, R in ((2, N), (2, E), (7, N), (100, E), (1, 3)):
        logits = (rng.standard_normal((R, D)) * 3).astype(np.float32)
        lab = rng.integers(0, D, size=R)
        onehot = np.zeros((R, D), np.float32)
        onehot[np.arange(R), lab] = 1
        soft = rng.random((R, D), dtype=np.float32)
        for tgt in (onehot, soft):
            got = float(gn.logitcrossentropy(torch.from_numpy(logits).cuda(), torch.from_numpy(tgt).cuda()).cpu())
            ref = O.logit_cross_entropy(logits, tgt)
            assert abs(got - ref) <= 1e-5 * max(abs(ref), 1e-3), (D, R, got, ref)
    # through the batched API: model output and targets as Padded features
    x = gn.batch(dict(graphs=adjs, ef=None, nf=[rng.random((6, a.shape[0]), dtype=np.float32) for a in adjs], gf=None))
    blk = gn.GNBlock((0, 6, 0), (2, 2, 0), rng=np.random.default_rng(5))
    y = blk(x)
    tn = np.eye(2, dtype=np.float32)[rng.integers(0, 2, size=N)]
    te = np.eye(2, dtype=np.float32)[rng.integers(0, 2, size=E)]
    t = gn.batch(dict(graphs=adjs, ef=[te[a:b].T for a, b in zip(x.graphs.index()["graph_edge_ptr"][:-1], x.graphs.index()["graph_edge_ptr"][1:])],
                      nf=[tn[a:b].T for a, b in zip(x.graphs.index()["graph_node_ptr"][:-1], x.graphs.index()["graph_node_ptr"][1:])], gf=None))
    loss = float((gn.logitcrossentropy(y.nf, t.nf) + gn.logitcrossentropy(y.ef, t.ef)).cpu())
    ref = O.logit_cross_entropy(y.nf.compact.cpu().numpy(), tn) + O.logit_cross_entropy(y.ef.compact.cpu().numpy(), te)
    assert abs(loss - ref) <= 1e-5 * abs(ref), (loss, ref)
