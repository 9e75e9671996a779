"""-m gpu: training step - gradients of the fp32 path against torch float64 autograd of the oracle forward
(oracle/gn_grad_oracle.py), the AdamW kernel against its formula, and a few optimisation steps that must reduce the loss."""
import numpy as np
import pytest
import torch

import workloads as W
from oracle import gn_oracle as O, gn_grad_oracle as G

pytestmark = pytest.mark.gpu
GRAD_TOL = 2e-4      # max-norm relative, fp32 accumulation against float64


@pytest.fixture(scope="module")
def gn():
    import graphnets_b200 as g
    return g


def _rel(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def _compare_trees(got, ref, path=""):
    worst = 0.0
    if isinstance(ref, dict):
        for k, v in ref.items():
            if k in ("din", "dout", "dims", "eps"):
                continue
            worst = max(worst, _compare_trees(got[k], v, path + "/" + k))
    elif isinstance(ref, (list, tuple)):
        for i, v in enumerate(ref):
            worst = max(worst, _compare_trees(got[i] if not isinstance(v, tuple) else got[i], v, path + "/%d" % i))
    elif isinstance(ref, np.ndarray):
        if ref.size == 0:
            return 0.0
        err = _rel(np.asarray(got, np.float64), ref)
        assert err <= GRAD_TOL, "%s: gradient rel err %.3e > %.0e (|ref| max %.3e)" % (path, err, GRAD_TOL, np.max(np.abs(ref)))
        worst = err
    return worst


def _case(gn, layers, w, eps_mode=0, seed=0):
    x = gn.batch(W.as_batch_input(w))
    tr = gn.Trainer(layers, eps_mode=eps_mode)
    y = tr.forward(x)
    g = O.lower(W.adj_list(w))
    ef, nf, gf = W.compact_inputs(w)
    rng = np.random.default_rng(seed)
    cot = [None if t is None else rng.standard_normal(tuple(t.shape)).astype(np.float32) for t in y]
    outs, pgrads, igrads = G.forward_and_grads(layers, g, ef, nf, gf, cot, eps_mode=eps_mode)
    for a, b in zip(y, outs):
        assert (a is None) == (b is None)
        if a is not None:
            assert _rel(a.cpu().numpy().astype(np.float64), b) <= 1e-5      # the training forward is the fp32 forward
    dev = tr.eng.torch_device
    din = tr.backward(*[None if c is None else torch.from_numpy(c).to(dev) for c in cot])
    worst = _compare_trees([p for _, p in tr.param_grads()], [p for _, p in pgrads], "params")
    for name, a, b in zip(("ef", "nf", "gf"), din, igrads):
        if b is not None:
            assert a is not None, name
            err = _rel(a.cpu().numpy().astype(np.float64), b)
            assert err <= GRAD_TOL, "input cotangent %s: %.3e" % (name, err)
            worst = max(worst, err)
    return tr, x, worst


def test_block_gradients_readme(gn):
    """README example 1: a single GNBlock (10,5,0)=>(3,4,5) on the 3-node graph, batch of 2 (test/runtests.jl:180-216)."""
    w = W.make_workload("cfg1")
    _case(gn, W.model_params("cfg1"), w)


@pytest.mark.parametrize("eps_mode", [0, 1, 2])
def test_core_stack_gradients(gn, eps_mode):
    """enc -> 2 x GNCore(10,5,3) -> dec on variable-size graphs (ragged, an isolated node), all three LayerNorm conventions."""
    rng = np.random.default_rng(7)
    adjs = []
    for n in (5, 9, 3, 12, 1, 7):
        a = (rng.random((n, n)) < 0.4).astype(np.uint8)
        if n > 2:
            a[:, 1] = 0      # node 1 receives nothing
        adjs.append(a)
    ef = [rng.random((10, int(a.sum())), dtype=np.float32) for a in adjs]      # (DE, m_b)
    nf = [rng.random((5, a.shape[0]), dtype=np.float32) for a in adjs]         # (DN, n_b)
    w = dict(mode="vector", graphs=adjs, ef=ef, nf=nf, gf=None)
    _case(gn, W.model_params("cfg2"), w, eps_mode=eps_mode)


def test_sort_model_shape_gradients(gn):
    """The reference's training example (examples/sort/sort.jl:86-88): no edge / graph inputs, node-only one-hot input, decoder
    without a graph output - zero-width pieces on both ends - at a reduced hidden width."""
    rng = np.random.default_rng(3)
    layers = [("block", W.block_params(rng, (0, 12, 0), (16, 16, 16))), ("core", W.core_params(rng, (16, 16, 16))),
              ("block", W.block_params(rng, (16, 16, 16), (2, 2, 0)))]
    sizes = [4, 7, 3, 8]
    graphs = [np.ones((n, n), np.uint8) for n in sizes]
    nf = [np.eye(12, dtype=np.float32)[rng.integers(0, 12, n)].T.copy() for n in sizes]      # (DN, n_b)
    w = dict(mode="vector", graphs=graphs, ef=None, nf=nf, gf=None)
    _case(gn, layers, w)


def test_adamw_and_descent(gn):
    """AdamW kernel against its formula; a few steps on a fixed batch must reduce a quadratic loss."""
    w = W.make_workload("cfg2", B=8)
    layers = W.model_params("cfg2")
    x = gn.batch(W.as_batch_input(w))
    tr = gn.Trainer(layers)
    p0 = tr.params.cpu().numpy().astype(np.float64)
    losses = []
    m = np.zeros_like(p0); v = np.zeros_like(p0); p = p0.copy()
    for it in range(4):
        y = tr.forward(x)
        losses.append(sum(float((t.double() ** 2).sum()) for t in y) * 0.5)
        tr.backward(*y)                                   # d(0.5 |y|^2)/dy = y
        g = tr.grads.cpu().numpy().astype(np.float64)
        tr.step(lr=1e-2, weight_decay=1e-2)
        if it == 0:
            m = 0.1 * g; v = 0.001 * g * g
            ref = p - 1e-2 * ((m / 0.1) / (np.sqrt(v / 0.001) + 1e-8) + 1e-2 * p)
            assert _rel(tr.params.cpu().numpy().astype(np.float64), ref) <= 1e-5
    assert losses[-1] < losses[0], losses


def test_cross_entropy_gradient(gn):
    """d logitcrossentropy / d logits against torch autograd (examples/sort/sort.jl:76-78)."""
    rng = np.random.default_rng(5)
    R, D = 1000, 7
    x = rng.standard_normal((R, D)).astype(np.float32)
    t = np.eye(D, dtype=np.float32)[rng.integers(0, D, R)]
    tr = gn.Trainer(W.model_params("cfg1"))
    dev = tr.eng.torch_device
    loss, dl = tr.cross_entropy(torch.from_numpy(x).to(dev), torch.from_numpy(t).to(dev))
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    ref = -(torch.tensor(t, dtype=torch.float64) * torch.log_softmax(xt, dim=1)).sum(dim=1).mean()
    ref.backward()
    assert abs(float(loss.cpu()) - float(ref)) <= 1e-5 * abs(float(ref))
    assert _rel(dl.cpu().numpy().astype(np.float64), xt.grad.numpy()) <= 1e-5


def test_bf16_gemm_training_step(gn):
    """Trainer(precision="bf16"): the GEMMs of the step with >= 256 rows and 128-multiple widths run with bf16 operands on the
    tensor cores (k_tc_lin, weights packed per call); gradients stay within bf16-operand distance of float64 autograd."""
    rng = np.random.default_rng(9)
    dims = (128, 128, 128)
    layers = [("block", W.block_params(rng, (10, 5, 0), dims)), ("core", W.core_params(rng, dims)), ("block", W.block_params(rng, dims, (3, 4, 5)))]
    adjs = [(rng.random((16, 16)) < 0.4).astype(np.uint8) for _ in range(64)]      # E ~ 6500 >= 4096: tensor-core wgrad on the edges
    ef = [rng.random((10, int(a.sum())), dtype=np.float32) for a in adjs]
    nf = [rng.random((5, 16), dtype=np.float32) for a in adjs]
    w = dict(mode="vector", graphs=adjs, ef=ef, nf=nf, gf=None)
    x = gn.batch(W.as_batch_input(w))
    tr = gn.Trainer(layers, precision="bf16")
    eng = tr.eng
    eng.set_profiling(True); eng.read_profile()
    y = tr.forward(x)
    g = O.lower(adjs)
    cef, cnf, cgf = W.compact_inputs(w)
    cot = [rng.standard_normal(tuple(t.shape)).astype(np.float32) for t in y]
    outs, pgrads, igrads = G.forward_and_grads(layers, g, cef, cnf, cgf, cot)
    dev = eng.torch_device
    tr.backward(*[torch.from_numpy(c).to(dev) for c in cot])
    prof = eng.read_profile()
    eng.set_profiling(False)
    assert prof.get("tc_linear", {}).get("launches", 0) >= 10, list(prof)      # forward + backward GEMMs on the tensor cores
    assert prof.get("train_wgrad_tc", {}).get("launches", 0) >= 3, list(prof)   # edge-row weight gradients on the tensor cores
    for a, b in zip(y, outs):
        assert _rel(a.cpu().numpy().astype(np.float64), b) <= 1e-2
    worst = 0.0
    got = [p for _, p in tr.param_grads()]
    ref = [p for _, p in pgrads]

    def walk(a, b, path):
        nonlocal worst
        if isinstance(b, dict):
            for k, v in b.items():
                if k not in ("din", "dout", "dims", "eps"):
                    walk(a[k], v, path + "/" + k)
        elif isinstance(b, list):
            for i, v in enumerate(b):
                walk(a[i], v, path + "/%d" % i)
        elif isinstance(b, np.ndarray) and b.size:
            err = _rel(np.asarray(a, np.float64), b)
            assert err <= 3e-2, "%s: bf16 gradient rel err %.3e" % (path, err)
            worst = max(worst, err)
    walk(got, ref, "params")
    assert worst > 1e-5      # (it really ran in reduced precision)


@pytest.mark.parametrize("R,K,N,gather", [(5000, 200, 136, True), (4096, 128, 128, False), (70001, 64, 320, False)])
def test_tensor_core_wgrad_operator(gn, R, K, N, gather):
    """gnb_op_wgrad on the tensor cores (k_tc_wgrad: bf16 operands, fp32 accumulation, split over row chunks) against the exact
    product of the bf16-rounded operands; ragged K / N tiles, a ragged last slab, gathered rows, accumulation into dW."""
    import ctypes as C
    from oracle.gn_oracle import round_bf16
    rng = np.random.default_rng(R)
    nx = 3000 if gather else R
    X = rng.standard_normal((nx, K)).astype(np.float32)
    dY = rng.standard_normal((R, N)).astype(np.float32)
    idx = rng.integers(0, nx, R).astype(np.int32) if gather else None
    dW0 = rng.standard_normal((K, N)).astype(np.float32)
    eng = gn.get_engine()
    eng.bind_stream()
    dev = eng.torch_device
    tX, tY, tW = (torch.from_numpy(a).to(dev) for a in (X, dY, dW0))
    tI = None if idx is None else torch.from_numpy(idx).to(dev)
    P = lambda t: None if t is None else C.c_void_p(t.data_ptr())
    L = gn.pkg._lib
    L.check(gn.lib.gnb_op_wgrad(eng.ctx, P(tX), K, K, P(tI), P(tY), N, N, R, P(tW), N, L.PRECISIONS["bf16"]))
    eng.sync()
    Xg = X if idx is None else X[idx]
    ref = dW0.astype(np.float64) + round_bf16(Xg).T @ round_bf16(dY)
    err = _rel(tW.cpu().numpy().astype(np.float64), ref)
    assert err <= 2e-5, err


def test_trainer_from_model_and_write_back(gn):
    """Trainer.from_model over layer objects; after a step the parameters written back make the product forward equal the
    Trainer's own forward (fp32 path, same kernels)."""
    w = W.make_workload("cfg2", B=8)
    model = W.to_gn_model(gn, W.model_params("cfg2"))
    x = gn.batch(W.as_batch_input(w))
    tr = gn.Trainer.from_model(model)
    y = tr.forward(x)
    tr.backward(*y)
    tr.step(lr=1e-3)
    tr.write_back(model)
    y_tr = tr.forward(x)
    y_model = model(x, precision="fp32")
    for a, b in zip(y_tr, (y_model.ef, y_model.nf, y_model.gf)):
        assert _rel(a.cpu().numpy().astype(np.float64), b.compact.cpu().numpy().astype(np.float64)) <= 1e-5
