"""Helper of tests/test_gpu_watchdog.py (run as a subprocess so that GNB_LIB_VARIANT / GNB_WATCHDOG_MS / the drain-delay
hook can be chosen per case): runs the cfg4 tensor-path forward twice - once on a normal context, once on a context whose
k_tc_proj drain warps are stalled per tile - and prints one JSON line {"timeout": bool, "equal": bool, "usable": bool}."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import graphnets_b200 as gn      # noqa: E402
import workloads as W            # noqa: E402


def main():
    B, delay_ns = int(sys.argv[1]), sys.argv[2]
    lib, L = gn.lib, gn.pkg._lib
    w = W.make_workload("cfg4", B=B)
    model = W.to_gn_model(gn, W.model_params("cfg4"))
    x = gn.batch(W.as_batch_input(w))
    eng = x.graphs.engine
    old = os.environ.get("GNB_LIB_VARIANT", "") == "oldproj"
    if old:
        # The round-1 protocol is a race: since round 2's kernels changed the relative speed of its roles it dead-locks even
        # without the artificial stall, so the tensor path cannot provide the reference here.  What this mode shows: the
        # watchdog turns the dead-lock into GNB_ERR_TIMEOUT, and the CUDA context keeps computing (fp32 path, bit for bit).
        y32 = model(x, precision="fp32")
        torch.cuda.synchronize()
        ref32 = [t.compact.clone() for t in (y32.ef, y32.nf, y32.gf)]
        res = {"timeout": False, "equal": False, "usable": False}
        os.environ["GNB_DEBUG_PROJ_DRAIN_DELAY_NS"] = delay_ns
        slow = gn.pkg.engine.Engine(eng.device)
        mh = model._model(eng)
        out = [torch.empty_like(t) for t in ref32]
        P = lambda t: C.c_void_p(t.data_ptr())
        slow.bind_stream()
        L.check(lib.gnb_model_forward(slow.ctx, mh, x.graphs.handle, P(x.ef.compact), P(x.nf.compact), None,
                                      P(out[0]), P(out[1]), P(out[2]), L.PRECISIONS["auto"]))
        try:
            slow.sync()
        except L.GnbTimeout:
            res["timeout"] = True
        y32b = model(x, precision="fp32")
        eng.sync()
        res["usable"] = all(torch.equal(a, b.compact) for a, b in zip(ref32, (y32b.ef, y32b.nf, y32b.gf)))
        print(json.dumps(res))
        return
    y = model(x, precision="auto")
    torch.cuda.synchronize()
    ref = [t.compact.clone() for t in (y.ef, y.nf, y.gf)]
    mh = model._model(eng)

    def run(engine):
        out = [torch.full_like(t, float("nan")) for t in ref]
        P = lambda t: C.c_void_p(t.data_ptr())
        engine.bind_stream()
        L.check(lib.gnb_model_forward(engine.ctx, mh, x.graphs.handle, P(x.ef.compact), P(x.nf.compact), None,
                                      P(out[0]), P(out[1]), P(out[2]), L.PRECISIONS["auto"]))
        engine.sync()      # raises GnbTimeout when a kernel watchdog fired
        return out

    os.environ["GNB_DEBUG_PROJ_DRAIN_DELAY_NS"] = delay_ns
    slow = gn.pkg.engine.Engine(eng.device)      # the hook is read at context creation
    res = {"timeout": False, "equal": False, "usable": False}
    try:
        out = run(slow)
        res["equal"] = all(torch.equal(a, b) for a, b in zip(out, ref))
    except L.GnbTimeout:
        res["timeout"] = True
    # the CUDA context must survive a fired watchdog: the same process keeps computing correct results
    out = run(eng)
    res["usable"] = all(torch.equal(a, b) for a, b in zip(out, ref))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
