"""-m gpu parity tests of the generic bf16 tcgen05 linear layers (csrc/tc_gemm.cu) that carry the GNCore layers whose
hidden width is a multiple of 128 but not 128 (BASELINE configs 3 and 5: 384, 256).  Tolerance: 1e-2 relative
(north_star, tensor-core MLP path), against the float64 oracle on the same inputs and weights."""
import numpy as np
import pytest
import torch

import workloads as W
from tests.gpu_util import BF16_TOL, FP32_TOL, assert_parity, run_oracle

# the wide cores carry more bf16 operand error than the 128-wide ones (tests/gpu_util.py::assert_wide_parity): RMS bound = max-norm bound
WIDE_RMS_TOL = 1.5e-2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gn():
    import graphnets_b200 as g
    return g


def _run(gn, layers, w, precision):
    model = W.to_gn_model(gn, layers)
    x = gn.batch(W.as_batch_input(w))
    eng = x.graphs.engine
    eng.set_profiling(True)
    eng.read_profile()
    y = model(x, precision=precision)
    torch.cuda.synchronize()
    prof = eng.read_profile()
    eng.set_profiling(False)
    c = lambda f: None if f is None else f.compact.cpu().numpy()
    return (c(y.ef), c(y.nf), c(y.gf)), prof


@pytest.mark.parametrize("B", [3, 16])
def test_cfg5_shape_tensor_path(gn, B):
    """hidden 256: edge / node Dense layers and FFNs run on tc_linear (row counts with a ragged last tile)."""
    w = W.make_workload("cfg5", B=B, n_nodes=37, n_edges=301)
    layers = W.model_params("cfg5")
    got, prof = _run(gn, layers, w, "auto")
    _, ref = run_oracle(layers, w)
    assert_parity(got, ref, BF16_TOL, "cfg5 auto", rms_tol=WIDE_RMS_TOL)
    assert "tc_linear" in prof and prof["tc_linear"]["launches"] > 0, "the tcgen05 linear kernel did not run: %s" % list(prof)
    got32, prof32 = _run(gn, layers, w, "fp32")
    assert_parity(got32, ref, FP32_TOL, "cfg5 fp32")
    assert "tc_linear" not in prof32


def test_cfg5_many_small_graphs(gn):
    """300 graphs of 4 nodes / 5 edges: the graph rows (B >= 256) also take the tensor-core kernels (fused FFN, tc_linear)."""
    w = W.make_workload("cfg5", B=300, n_nodes=4, n_edges=5)
    layers = W.model_params("cfg5")
    got, prof = _run(gn, layers, w, "auto")
    _, ref = run_oracle(layers, w)
    assert_parity(got, ref, BF16_TOL, "cfg5 small graphs", rms_tol=WIDE_RMS_TOL)
    assert prof["tc_ffn256"]["launches"] == 3 * 4, prof["tc_ffn256"]      # edge, node and graph FFN of each of the 4 cores


def test_cfg3_shape_tensor_path(gn):
    """hidden 384 (3 output blocks: odd block count), node-only inputs, fully connected graphs of 8-24 nodes."""
    w = W.make_workload("cfg3", B=20, n_nodes=(8, 24))
    layers = W.model_params("cfg3")
    got, prof = _run(gn, layers, w, "auto")
    _, ref = run_oracle(layers, w)
    assert_parity(got, ref, BF16_TOL, "cfg3 auto", rms_tol=WIDE_RMS_TOL)
    assert "tc_linear" in prof


def test_bf16_mode_accepts_wide_cores(gn):
    w = W.make_workload("cfg5", B=4)
    layers = W.model_params("cfg5")
    got, prof = _run(gn, layers, w, "bf16")
    _, ref = run_oracle(layers, w)
    assert_parity(got, ref, BF16_TOL, "cfg5 bf16", rms_tol=WIDE_RMS_TOL)


def test_repeated_forward_is_deterministic(gn):
    """Packed weights are cached per model; a second forward reuses them and reproduces the result bit for bit."""
    w = W.make_workload("cfg5", B=5, n_nodes=40, n_edges=333)
    layers = W.model_params("cfg5")
    model = W.to_gn_model(gn, layers)
    x = gn.batch(W.as_batch_input(w))
    y1 = model(x, precision="auto")
    a = [f.compact.clone() for f in (y1.ef, y1.nf, y1.gf)]
    y2 = model(x, precision="auto")
    for u, f in zip(a, (y2.ef, y2.nf, y2.gf)):
        assert torch.equal(u, f.compact)
