"""Layer objects of the reference API: GNBlock, GNCore, GNCoreList (+ GNFeedForward, GNGraphNorm).

Reference: src/gnblock.jl, src/gncore.jl, src/gncorelist.jl, src/gnfeedforward.jl,
src/gngraphnorm.jl.  Layers are callable parameter holders with the reference's public field
names (`edgefn/nodefn/graphfn/dropout`, `block/ffwd/gn1/gn2`, `eff/nff/gff`,
`edgeln/nodeln/graphln`, `list`).  Parameters live on the host as float32 numpy arrays in Flux's
layout (`Dense.weight` is `(out, in)`, column-major); calling a layer uploads them once into a
`gnb_model` (the analogue of `model |> gpu`) and runs the CUDA forward through the C ABI.
Call `.sync()` after mutating parameters in place.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import lib, check
from .api import GNData, Padded, _fields
from .engine import _ptr

# The reference computes in Float32, so the drop-in default is the fp32 path (1e-5 parity with the float64 oracle).  The
# tcgen05 bf16 path is an explicit opt-in: set_precision("auto" | "bf16") or the per-call `precision=` argument.
_default_precision = "fp32"


def set_precision(p):
    """Default precision of layer calls:
    'fp32' (default) - CUDA-core fp32 path for every layer, parity 1e-5 relative;
    'auto'           - tcgen05 bf16 tensor-core path (fp32 accumulate, parity 1e-2 relative) for the GNCore shapes it
                       supports (hidden width a multiple of 128) and the algebraically re-associated streaming kernels for
                       narrow encoder / decoder blocks; fp32 for everything else;
    'bf16'           - like 'auto', but a GNCore the tensor path cannot run is an error instead of a silent fp32 layer."""
    global _default_precision
    assert p in _lib.PRECISIONS
    _default_precision = p


def get_precision():
    return _default_precision


def _glorot(rng, out, inn):
    # Flux.glorot_uniform: U(-x, x), x = sqrt(6 / (fan_in + fan_out))
    lim = np.sqrt(6.0 / (inn + out)) if inn + out > 0 else 0.0
    return np.asfortranarray(rng.uniform(-lim, lim, size=(out, inn)).astype(np.float32))


class Dense:
    """Flux.Dense(in => out, sigma): weight (out, in), bias (out,)."""

    def __init__(self, inn, out, activation=None, rng=None):
        rng = rng or np.random.default_rng()
        self.weight = _glorot(rng, out, inn)
        self.bias = np.zeros(out, np.float32)
        self.activation = activation

    def set(self, weight=None, bias=None):
        if weight is not None:
            w = np.asarray(weight, np.float32)
            assert w.shape == self.weight.shape, "weight shape %s != %s" % (w.shape, self.weight.shape)
            self.weight = np.asfortranarray(w)
        if bias is not None:
            b = np.asarray(bias, np.float32)
            assert b.shape == self.bias.shape
            self.bias = np.ascontiguousarray(b)
        return self


class Chain:
    """Flux.Chain - a tuple of layers (`m.edgefn[0]` is the Dense)."""

    def __init__(self, *layers):
        self.layers = tuple(layers)

    def __getitem__(self, i):
        return self.layers[i]

    def __len__(self):
        return len(self.layers)


class Dropout:
    """Flux.Dropout(p).  Identity: GNBlock never applies it (src/gnblock.jl:59 vs :63-69) and
    GNFeedForward's is inactive outside training."""

    def __init__(self, p=0):
        self.p = p


class LayerNorm:
    """Flux.LayerNorm(d): diag.scale (gamma) = 1, diag.bias (beta) = 0, eps = 1f-5."""

    def __init__(self, d, eps=1e-5, eps_mode=0):
        self.scale = np.ones(d, np.float32)
        self.bias = np.zeros(d, np.float32)
        self.eps = float(eps)
        self.eps_mode = int(eps_mode)

    def set(self, scale=None, bias=None):
        if scale is not None:
            self.scale = np.ascontiguousarray(np.asarray(scale, np.float32).reshape(self.scale.shape))
        if bias is not None:
            self.bias = np.ascontiguousarray(np.asarray(bias, np.float32).reshape(self.bias.shape))
        return self


def _wp(a):
    """host pointer of a parameter array (Fortran/contiguous float32) or None when empty"""
    if a is None or a.size == 0:
        return None
    assert a.dtype == np.float32
    assert a.flags.f_contiguous or a.flags.c_contiguous and a.ndim == 1
    return a.ctypes.data


class _Runnable:
    """Shared engine plumbing: build/cached gnb_model per device, forward on Padded features."""

    precision = None   # None -> module default

    def _layers_desc(self):   # -> list of _lib.Layer
        raise NotImplementedError

    def sync(self):
        """Drop the device copies of the parameters (re-uploaded on next call)."""
        for m in getattr(self, "_models", {}).values():
            lib.gnb_model_destroy(m)
        self._models = {}

    def __del__(self):
        try:
            self.sync()
        except Exception:
            pass

    def _model(self, engine):
        models = self.__dict__.setdefault("_models", {})
        if engine.device not in models:
            descs, keep = self._layers_desc()
            arr = (_lib.Layer * len(descs))(*descs)
            h = C.c_void_p()
            engine.bind_stream()
            check(lib.gnb_model_create(engine.ctx, arr, len(descs), 0, C.byref(h)))
            del keep
            models[engine.device] = h
        return models[engine.device]

    def _in_dims(self):
        raise NotImplementedError

    def _out_dims(self):
        raise NotImplementedError

    def __call__(self, x, precision=None):
        graphs, ef, nf, gf = _fields(x)
        g = graphs
        din, dout = self._in_dims(), self._out_dims()
        for name, f, d in (("ef", ef, din[0]), ("nf", nf, din[1]), ("gf", gf, din[2])):
            got = 0 if f is None else f.D
            assert got == d, "DimensionMismatch: layer expects %s width %d, got %d" % (name, d, got)
        eng = g.engine
        model = self._model(eng)
        eng.bind_stream()
        oe = eng.empty(g.E, dout[0]) if dout[0] else None
        on = eng.empty(g.N, dout[1]) if dout[1] else None
        og = eng.empty(g.B, dout[2]) if dout[2] else None
        prec = _lib.PRECISIONS[precision or self.precision or _default_precision]
        c = lambda f: None if f is None else _ptr(f.compact)
        check(lib.gnb_model_forward(eng.ctx, model, g.handle, c(ef), c(nf), c(gf), _ptr(oe), _ptr(on), _ptr(og), prec))
        wrap = lambda k, t: None if t is None else Padded(k, t, g)     # zerodim2nothing (src/gnblock.jl:71-78)
        return GNData(g, wrap("e", oe), wrap("n", on), wrap("g", og))


def _block_struct(blk, keep):
    p = _lib.BlockParams()
    (p.in_e, p.in_n, p.in_g), (p.out_e, p.out_n, p.out_g) = blk.in_dims, blk.out_dims
    for name, dense in (("e", blk.edgefn[0]), ("n", blk.nodefn[0]), ("g", blk.graphfn[0])):
        w, b = np.asfortranarray(dense.weight), np.ascontiguousarray(dense.bias)
        keep += [w, b]
        setattr(p, "W" + name, _wp(w))
        setattr(p, "b" + name, _wp(b))
    return p


class GNBlock(_Runnable):
    """GNBlock((X_DE,X_DN,X_DG) => (Y_DE,Y_DN,Y_DG); dropout=0)   (src/gnblock.jl:47-61)"""

    def __init__(self, in_dims, out_dims=None, dropout=0, rng=None):
        if out_dims is None:      # GNBlock((in, out)) pair form
            in_dims, out_dims = in_dims
        in_dims, out_dims = tuple(int(v) for v in in_dims), tuple(int(v) for v in out_dims)
        assert len(in_dims) == 3 and len(out_dims) == 3
        assert any(v > 0 for v in in_dims)
        assert any(v > 0 for v in out_dims)
        rng = rng or np.random.default_rng()
        ei, ni, gi = in_dims
        eo, no, go = out_dims
        self.in_dims, self.out_dims = in_dims, out_dims
        self.edgefn = Chain(Dense(ei + 2 * ni + gi, eo, rng=rng))
        self.nodefn = Chain(Dense(ni + eo + gi, no, rng=rng))
        self.graphfn = Chain(Dense(no + eo + gi, go, rng=rng))
        self.dropout = Dropout(dropout)

    def _layers_desc(self):
        keep = []
        L = _lib.Layer()
        L.kind = _lib.LAYER_BLOCK
        L.block = _block_struct(self, keep)
        return [L], keep

    def _in_dims(self):
        return self.in_dims

    def _out_dims(self):
        return self.out_dims


class GNFeedForward:
    """GNFeedForward(dims; dropout=0): eff/nff/gff = Chain(Dense(d=>4d, relu), Dense(4d=>d), Dropout)
    (src/gnfeedforward.jl:17-31)"""

    def __init__(self, dims, dropout=0, rng=None):
        assert all(d > 0 for d in dims)
        rng = rng or np.random.default_rng()
        mk = lambda d: Chain(Dense(d, 4 * d, "relu", rng=rng), Dense(4 * d, d, rng=rng), Dropout(dropout))
        self.eff, self.nff, self.gff = (mk(d) for d in dims)


class GNGraphNorm:
    """GNGraphNorm(dims): edgeln/nodeln/graphln = LayerNorm(d)   (src/gngraphnorm.jl:9-17)"""

    def __init__(self, dims, eps_mode=0):
        assert all(d > 0 for d in dims)
        self.edgeln, self.nodeln, self.graphln = (LayerNorm(d, eps_mode=eps_mode) for d in dims)


def _core_struct(core, keep):
    p = _lib.CoreParams()
    p.block = _block_struct(core.block, keep)
    for i, ch in enumerate((core.ffwd.eff, core.ffwd.nff, core.ffwd.gff)):
        w1, b1 = np.asfortranarray(ch[0].weight), np.ascontiguousarray(ch[0].bias)
        w2, b2 = np.asfortranarray(ch[1].weight), np.ascontiguousarray(ch[1].bias)
        keep += [w1, b1, w2, b2]
        p.ffn[i].W1, p.ffn[i].b1, p.ffn[i].W2, p.ffn[i].b2 = _wp(w1), _wp(b1), _wp(w2), _wp(b2)
    for arr, gn in ((p.ln1, core.gn1), (p.ln2, core.gn2)):
        for i, ln in enumerate((gn.edgeln, gn.nodeln, gn.graphln)):
            s, b = np.ascontiguousarray(ln.scale), np.ascontiguousarray(ln.bias)
            keep += [s, b]
            arr[i].gamma, arr[i].beta, arr[i].eps, arr[i].eps_mode = _wp(s), _wp(b), ln.eps, ln.eps_mode
    return p


class GNCore(_Runnable):
    """GNCore(dims; dropout=0): x + block(gn1(x)) + ffwd(gn2(x))   (src/gncore.jl:46-59)"""

    def __init__(self, dims, dropout=0, rng=None, eps_mode=0):
        dims = tuple(int(d) for d in dims)
        assert any(d > 0 for d in dims)
        rng = rng or np.random.default_rng()
        self.dims = dims
        self.block = GNBlock(dims, dims, dropout=dropout, rng=rng)
        self.ffwd = GNFeedForward(dims, dropout=dropout, rng=rng)
        self.gn1 = GNGraphNorm(dims, eps_mode=eps_mode)
        self.gn2 = GNGraphNorm(dims, eps_mode=eps_mode)

    def _layer(self, keep):
        L = _lib.Layer()
        L.kind = _lib.LAYER_CORE
        L.core = _core_struct(self, keep)
        return L

    def _layers_desc(self):
        keep = []
        return [self._layer(keep)], keep

    def _in_dims(self):
        return self.dims

    def _out_dims(self):
        return self.dims


class GNCoreList(_Runnable):
    """GNCoreList(list): left fold of cores (src/gncorelist.jl:37-45).  The whole list runs as one
    engine call (one gnb_model)."""

    def __init__(self, cores):
        if isinstance(cores, dict):
            cores = list(cores.values())
        self.list = list(cores)
        assert len(self.list) > 0

    def _layers_desc(self):
        keep = []
        descs = []
        for c in self.list:
            if isinstance(c, GNCore):
                descs.append(c._layer(keep))
            elif isinstance(c, GNBlock):
                d, k = c._layers_desc()
                descs += d
                keep += k
            else:
                raise AssertionError("GNCoreList elements must be GNCore / GNBlock")
        return descs, keep

    def _in_dims(self):
        return self.list[0]._in_dims()

    def _out_dims(self):
        return self.list[-1]._out_dims()


class GNSequential(GNCoreList):
    """`decoder o core_list o encoder` fused into ONE engine call: any sequence of GNBlock / GNCore /
    GNCoreList (README.md:133-149 builds the same model by function composition)."""

    def __init__(self, *layers):
        flat = []
        for l in layers:
            flat += l.list if isinstance(l, GNCoreList) else [l]
        super().__init__(flat)
