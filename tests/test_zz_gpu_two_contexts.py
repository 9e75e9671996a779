"""-m gpu: double-buffered host-buffer forward with two contexts (runs last: file name sorts after the other GPU tests).

Opt-in (GNB_TEST_TWO_CONTEXTS=1): the mode is experimental - it ran clean at 1, 2 and 4 GPUs, but one 8-GPU bench run
trapped inside a forward and the cause is not understood yet."""
import os

import numpy as np
import pytest
import torch

import workloads as W

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gn():
    import graphnets_b200 as g
    return g


@pytest.mark.skipif(os.environ.get("GNB_TEST_TWO_CONTEXTS") != "1", reason="experimental two-context pipelining: opt-in")
def test_two_contexts_pipelined_host_forward(gn):
    """Double-buffered input pipeline: two host threads, each with its own context and stream, alternate batches through the
    synchronous host-buffer calls (bench.py `e2e`).  The library serialises the forwards of different contexts on the device
    (stream-ordered), so every result equals the single-context one."""
    import ctypes as C
    import threading
    w = W.make_workload("cfg4", B=512)
    layers = W.model_params("cfg4")
    model = W.to_gn_model(gn, layers)
    x = gn.batch(W.as_batch_input(w))
    y = model(x, precision="auto")
    torch.cuda.synchronize()
    eng = x.graphs.engine
    mh = model._model(eng)
    ef, nf, _ = W.compact_inputs(w)
    adj = np.stack(W.adj_list(w))
    B, n = adj.shape[0], adj.shape[1]
    mask = np.ascontiguousarray((adj == 1).transpose(0, 2, 1)).astype(np.uint8)
    nn = (C.c_int32 * B)(*([n] * B))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    engs = [eng, gn.pkg.engine.Engine(eng.device)]
    streams = [torch.cuda.Stream() for _ in engs]
    for e_, s_ in zip(engs, streams):
        gn.pkg._lib.check(gn.lib.gnb_ctx_set_stream(e_.ctx, C.c_void_p(s_.cuda_stream)))
    outs = [[(np.full((x.graphs.E, 3), np.nan, np.float32), np.full((x.graphs.N, 4), np.nan, np.float32),
              np.full((B, 5), np.nan, np.float32)) for _ in range(2)] for _ in engs]
    errors = []

    def worker(k):
        try:
            torch.cuda.set_device(eng.device)
            for i in range(2):
                h = C.c_void_p()
                gn.pkg._lib.check(gn.lib.gnb_graph_lower(engs[k].ctx, p(mask), 1, 0, nn, n, B, B, C.byref(h)))
                oe, on, og = outs[k][i]
                gn.pkg._lib.check(gn.lib.gnb_model_forward_host(engs[k].ctx, mh, h, p(ef), p(nf), None, p(oe), p(on), p(og),
                                                               gn.pkg._lib.PRECISIONS["auto"]))
                gn.lib.gnb_graph_destroy(h)
        except Exception as e:      # noqa: BLE001
            errors.append(e)
    ts = [threading.Thread(target=worker, args=(k,)) for k in range(2)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    eng.bind_stream()
    assert not errors, errors
    ref = (y.ef.compact.cpu().numpy(), y.nf.compact.cpu().numpy(), y.gf.compact.cpu().numpy())
    for k in range(2):
        for i in range(2):
            for a, b in zip(outs[k][i], ref):
                assert np.array_equal(a, b), (k, i)
