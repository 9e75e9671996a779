"""Training step of a GNBlock / GNCore model on the fp32 CUDA-core path.

The reference trains with `Flux.withgradient(model) do m ... end` + `Flux.update!` (examples/sort/sort.jl:122-132): Zygote
differentiates the forward of src/gnblock.jl:63-69 / src/gncore.jl:56-68.  Here the adjoint of every layer is composed on the host
from the primitive operators of csrc/train.cu (C ABI `gnb_op_*`, include/gnb200.h) - the forward's own fused linear kernel for
every `dX = dY W^T`, deterministic segmented sums for the adjoints of the gathers (taken BEFORE the GEMM, by linearity),
two-stage atomic-free reductions for `dW = X^T dY`, LayerNorm / relu adjoints - over one flat parameter / gradient buffer, which
is all-reduced across ranks (NCCL through torch.distributed: plumbing) and applied with the AdamW kernel.

    tr = Trainer(layers)                      # layers: [("block" | "core", params)] with numpy (out, in) weights (workloads.py)
    y = tr.forward(x)                         # x = gn.batch(...): (ef, nf, gf) compact torch tensors, activations kept
    tr.backward(d_ef, d_nf, d_gf)             # cotangents of the outputs -> tr.grads (flat), returns the input cotangents
    tr.step(lr=1e-3)                          # (all-reduce when torch.distributed is initialised) + AdamW

PyTorch provides device memory and the process group; every arithmetic operation is a kernel of libgnb200.so."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import lib, check

_P = lambda t: None if t is None else C.c_void_p(t.data_ptr())


class _Ops:
    """ctypes wrappers of the gnb_op_* entry points on torch device tensors (fp32, contiguous)."""

    def __init__(self, eng, precision="fp32"):
        self.eng, self.ctx, self.dev = eng, eng.ctx, eng.torch_device
        self.prec = _lib.PRECISIONS[precision]

    def empty(self, *shape):
        return torch.empty(*shape, dtype=torch.float32, device=self.dev)

    def zeros(self, *shape):
        return torch.zeros(*shape, dtype=torch.float32, device=self.dev)

    def linear(self, R, nout, srcs, ldw, bias=None, adds=(), relu=False, out=None):
        """out[R][nout] = act(sum_s x_s W_s + bias + sum_j a_j[idx_j]);  srcs: [(x [R][d], W pointer-tensor view [d][ldw])]."""
        srcs = [(x, W) for x, W in srcs if x is not None and x.shape[1] > 0]
        assert 1 <= len(srcs) <= 3 and len(adds) <= 4
        out = self.empty(R, nout) if out is None else out
        a = _lib.LinArgs()
        a.R, a.Nout, a.ldw, a.nsrc = R, nout, ldw, len(srcs)
        for i, (x, W) in enumerate(srcs):
            a.src[i].x, a.src[i].d, a.src[i].ldx, a.src[i].W = x.data_ptr(), x.shape[1], x.stride(0), W.data_ptr()
            a.src[i].gamma = a.src[i].beta = None
        a.bias = None if bias is None else bias.data_ptr()
        a.nadd = len(adds)
        for j, (t, idx) in enumerate(adds):
            a.add[j].a, a.add[j].idx, a.add[j].lda = t.data_ptr(), (None if idx is None else idx.data_ptr()), t.stride(0)
        a.relu, a.out, a.ldo, a.precision = int(relu), out.data_ptr(), out.stride(0), self.prec
        check(lib.gnb_op_linear(self.ctx, C.byref(a)))
        return out

    def segsum(self, x, ptr, S, perm=None):
        out = self.empty(S, x.shape[1])
        check(lib.gnb_op_segsum(self.ctx, _P(x), x.shape[1], _P(ptr), S, _P(perm), _P(out)))
        return out

    def layernorm(self, x, gamma, beta, eps, mode):
        y = torch.empty_like(x)
        check(lib.gnb_op_layernorm(self.ctx, _P(x), x.shape[0], x.shape[1], _P(gamma), _P(beta), eps, mode, _P(y)))
        return y

    def layernorm_bwd(self, x, g, gamma, eps, mode, dx):
        gx = torch.empty_like(x)
        check(lib.gnb_op_layernorm_bwd(self.ctx, _P(x), _P(g), x.shape[0], x.shape[1], _P(gamma), eps, mode, _P(dx), _P(gx)))
        return gx

    def wgrad(self, X, dY, dW, ldw, idx=None):
        """dW[k][n] += sum_r X[idx[r]][k] dY[r][n]   (dW: pointer view into the flat gradient buffer, rows of ldw floats)"""
        if X is None or X.shape[1] == 0 or dY.shape[1] == 0:
            return
        check(lib.gnb_op_wgrad(self.ctx, _P(X), X.stride(0), X.shape[1], _P(idx), _P(dY), dY.stride(0), dY.shape[1], dY.shape[0], _P(dW), ldw, self.prec))

    def colsum(self, X, out):
        check(lib.gnb_op_colsum(self.ctx, _P(X), X.stride(0), X.shape[1], X.shape[0], _P(out)))

    def relu_mask(self, t, h):
        check(lib.gnb_op_relu_mask(self.ctx, _P(t), _P(h), t.numel()))

    def gather_add(self, R, D, a=None, b1=None, idx1=None, b2=None, idx2=None, out=None):
        out = self.empty(R, D) if out is None else out
        check(lib.gnb_op_gather_add(self.ctx, _P(out), _P(a), _P(b1), _P(idx1), _P(b2), _P(idx2), R, D))
        return out

    def transpose(self, W, rows, cols, ld):
        """W: view of a [rows][ld] block (first `cols` columns used) -> [cols][rows]"""
        out = self.empty(cols, rows)
        check(lib.gnb_op_transpose(self.ctx, _P(W), rows, cols, ld, _P(out)))
        return out


def allreduce_flat(grads, average=True):
    """One all-reduce of the flat gradient buffer over the default process group (NCCL for device tensors, gloo on the CPU in the
    tests); a no-op without an initialised group.  Data-parallel training over graph shards: every rank holds the gradient of ITS
    shard's loss, the sum (or mean) is the gradient of the whole batch (weights are replicated)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(grads, op=dist.ReduceOp.SUM)
        if average:
            grads.div_(dist.get_world_size())
    return grads


class _Index:
    """Device index arrays of a lowered batch (owned by the graph handle) + the sender-sorted permutation of the edges."""

    def __init__(self, graphs):
        self.E, self.N, self.B = graphs.E, graphs.N, graphs.B
        dev = graphs.engine.torch_device
        idx = graphs.index()      # host copy (numpy)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.int32)).to(dev)
        self.src, self.dst = t(idx["edge_src"]), t(idx["edge_dst"])
        self.gep, self.gnp, self.nip = t(idx["graph_edge_ptr"]), t(idx["graph_node_ptr"]), t(idx["node_in_ptr"])
        eg = np.repeat(np.arange(self.B, dtype=np.int32), np.diff(idx["graph_edge_ptr"]))
        ng = np.repeat(np.arange(self.B, dtype=np.int32), np.diff(idx["graph_node_ptr"]))
        self.eg, self.ng = t(eg), t(ng)
        # the adjoint of the gather by sender is a segmented sum over SENDERS: stable sort of the (receiver-sorted) edges by
        # sender -> deterministic order inside every segment; host-side preprocessing of the batch, like the lowering's own
        perm = np.argsort(idx["edge_src"], kind="stable").astype(np.int32)
        out_ptr = np.zeros(self.N + 1, np.int64)
        np.cumsum(np.bincount(idx["edge_src"], minlength=self.N), out=out_ptr[1:])
        self.src_perm, self.node_out_ptr32 = t(perm), t(out_ptr)      # t() stores int32


def layers_of(model):
    """The oracle-format parameter list [("block" | "core", params)] of a model built from the layer objects of layers.py
    (GNBlock, GNCore, GNCoreList / GNSequential): weights as numpy (out, in) arrays, like Flux's Dense.weight."""
    from .layers import GNBlock, GNCore, GNCoreList

    def blk(b):
        return dict(din=tuple(b.in_dims), dout=tuple(b.out_dims), We=np.asarray(b.edgefn[0].weight), be=np.asarray(b.edgefn[0].bias),
                    Wn=np.asarray(b.nodefn[0].weight), bn=np.asarray(b.nodefn[0].bias), Wg=np.asarray(b.graphfn[0].weight), bg=np.asarray(b.graphfn[0].bias))

    def one(m):
        if isinstance(m, GNBlock):
            return [("block", blk(m))]
        if isinstance(m, GNCore):
            ffn = [dict(W1=np.asarray(ch[0].weight), b1=np.asarray(ch[0].bias), W2=np.asarray(ch[1].weight), b2=np.asarray(ch[1].bias))
                   for ch in (m.ffwd.eff, m.ffwd.nff, m.ffwd.gff)]
            ln = lambda g: [dict(gamma=np.asarray(l.scale), beta=np.asarray(l.bias), eps=float(l.eps)) for l in (g.edgeln, g.nodeln, g.graphln)]
            return [("core", dict(dims=tuple(m.dims), block=blk(m.block), ffn=ffn, ln1=ln(m.gn1), ln2=ln(m.gn2)))]
        if isinstance(m, GNCoreList):
            return [l for c in m.list for l in one(c)]
        raise TypeError("Trainer: unsupported layer %r" % type(m).__name__)
    return one(model)


class Trainer:
    @classmethod
    def from_model(cls, model, **kw):
        """Trainer over the parameters of a GNBlock / GNCore / GNCoreList / GNSequential built with the layer objects; the
        LayerNorm convention (eps_mode) is taken from the model's first GNCore."""
        from .layers import GNCore, GNCoreList
        cores = [c for c in (model.list if isinstance(model, GNCoreList) else [model]) if isinstance(c, GNCore)]
        kw.setdefault("eps_mode", cores[0].gn1.edgeln.eps_mode if cores else 0)
        return cls(layers_of(model), **kw)

    def write_back(self, model):
        """Copies the (trained) parameters back into the layer objects `from_model` was given (and re-syncs their device copies)."""
        from .layers import GNBlock, GNCore, GNCoreList
        p = self.params.cpu().numpy()
        mods = model.list if isinstance(model, GNCoreList) else [model]

        def W(off, rows_in, cols_out):
            return p[off:off + rows_in * cols_out].reshape(rows_in, cols_out).T.copy()

        def put_block(b, s):
            (a, bb, c), (pp, q, r) = s["din"], s["dout"]
            b.edgefn[0].set(W(s["We"], a + 2 * bb + c, pp), p[s["be"]:s["be"] + pp])
            b.nodefn[0].set(W(s["Wn"], pp + bb + c, q), p[s["bn"]:s["bn"] + q])
            b.graphfn[0].set(W(s["Wg"], pp + q + c, r), p[s["bg"]:s["bg"] + r])
        flat = [c for m in mods for c in (m.list if isinstance(m, GNCoreList) else [m])]
        assert len(flat) == len(self.spec)
        for m, (kind, s) in zip(flat, self.spec):
            if kind == "block":
                assert isinstance(m, GNBlock)
                put_block(m, s)
            else:
                assert isinstance(m, GNCore)
                put_block(m.block, s["block"])
                for ch, f, d in zip((m.ffwd.eff, m.ffwd.nff, m.ffwd.gff), s["ffn"], s["dims"]):
                    ch[0].set(W(f["W1"], d, 4 * d), p[f["b1"]:f["b1"] + 4 * d])
                    ch[1].set(W(f["W2"], 4 * d, d), p[f["b2"]:f["b2"] + d])
                for g, key in ((m.gn1, "ln1"), (m.gn2, "ln2")):
                    for l, ls, d in zip((g.edgeln, g.nodeln, g.graphln), s[key], s["dims"]):
                        l.set(p[ls["gamma"]:ls["gamma"] + d], p[ls["beta"]:ls["beta"] + d])
        if hasattr(model, "sync"):
            model.sync()
        return model

    def __init__(self, layers, eps_mode=0, engine=None, precision="fp32"):
        """precision: "fp32" (default: gradients to 2e-4 of float64 autograd) or "bf16" / "auto": every GEMM of the step with
        enough rows and tensor-core friendly widths (forward, recomputation, dX = dY W^T) runs with bf16 operands and fp32
        accumulation on the generic tcgen05 linear kernel; weight gradients, LayerNorm, reductions and the optimiser stay fp32."""
        from .engine import get_engine
        self.eng = engine or get_engine()
        self.ops = _Ops(self.eng, precision)
        self.eps_mode = int(eps_mode)
        # ---- flat parameter buffer; every weight in the ABI layout [in][out] (== Flux (out, in) column-major)
        self.spec, chunks, off = [], [], 0

        def add(arr):
            nonlocal off
            a = np.ascontiguousarray(arr, np.float32).reshape(-1)
            chunks.append(a)
            o = off
            off += a.size
            return o

        def add_block(p):
            return dict(din=tuple(p["din"]), dout=tuple(p["dout"]), We=add(p["We"].T), be=add(p["be"]), Wn=add(p["Wn"].T), bn=add(p["bn"]),
                        Wg=add(p["Wg"].T), bg=add(p["bg"]))

        for kind, p in layers:
            if kind == "block":
                self.spec.append(("block", add_block(p)))
            else:
                d = dict(dims=tuple(p["dims"]), block=add_block(dict(p["block"], din=p["dims"], dout=p["dims"])), ffn=[], ln1=[], ln2=[])
                for f in p["ffn"]:
                    d["ffn"].append(dict(W1=add(f["W1"].T), b1=add(f["b1"]), W2=add(f["W2"].T), b2=add(f["b2"])))
                for key in ("ln1", "ln2"):
                    for l in p[key]:
                        d[key].append(dict(gamma=add(l["gamma"]), beta=add(l["beta"]), eps=float(l.get("eps", 1e-5))))
                self.spec.append(("core", d))
        flat = np.concatenate(chunks) if chunks else np.zeros(0, np.float32)
        self.params = torch.from_numpy(flat).to(self.eng.torch_device)
        self.grads = torch.zeros_like(self.params)
        self.m, self.v, self.t = torch.zeros_like(self.params), torch.zeros_like(self.params), 0
        self._saved = None

    # views into the flat buffers
    def _p(self, off):
        return self.params[off:]

    def _g(self, off):
        return self.grads[off:]

    def param_grads(self):
        """Gradients as a layers-shaped tree of numpy arrays in the (out, in) layout of the input parameters (for tests)."""
        g = self.grads.cpu().numpy()

        def W(off, rows_in, cols_out):
            return g[off:off + rows_in * cols_out].reshape(rows_in, cols_out).T.copy()

        def vec(off, n):
            return g[off:off + n].copy()

        def blk(b):
            (a, bb, c), (p, q, r) = b["din"], b["dout"]
            return dict(We=W(b["We"], a + 2 * bb + c, p), be=vec(b["be"], p), Wn=W(b["Wn"], p + bb + c, q), bn=vec(b["bn"], q),
                        Wg=W(b["Wg"], p + q + c, r), bg=vec(b["bg"], r))
        out = []
        for kind, s in self.spec:
            if kind == "block":
                out.append(("block", blk(s)))
            else:
                d = dict(block=blk(s["block"]), ffn=[], ln1=[], ln2=[])
                for f, dd in zip(s["ffn"], s["dims"]):
                    d["ffn"].append(dict(W1=W(f["W1"], dd, 4 * dd), b1=vec(f["b1"], 4 * dd), W2=W(f["W2"], 4 * dd, dd), b2=vec(f["b2"], dd)))
                for key in ("ln1", "ln2"):
                    for l, dd in zip(s[key], s["dims"]):
                        d[key].append(dict(gamma=vec(l["gamma"], dd), beta=vec(l["beta"], dd)))
                out.append(("core", d))
        return out

    # ------------------------------------------------------------------ forward (activations kept)
    def _block_fwd(self, b, ix, e, v, u):
        o = self.ops
        (a, bb, c), (p, q, r) = b["din"], b["dout"]
        E, N, B = ix.E, ix.N, ix.B
        We, Wn, Wg = self._p(b["We"]), self._p(b["Wn"]), self._p(b["Wg"])
        row = lambda Wt, k, ld: Wt[k * ld:]
        he = hv = hu = None
        if p > 0:
            adds = []
            if bb > 0:
                adds += [(o.linear(N, p, [(v, row(We, a, p))], p), ix.src), (o.linear(N, p, [(v, row(We, a + bb, p))], p), ix.dst)]
            if c > 0:
                adds.append((o.linear(B, p, [(u, row(We, a + 2 * bb, p))], p), ix.eg))
            if a > 0:
                he = o.linear(E, p, [(e, We)], p, bias=self._p(b["be"]), adds=adds)
            else:      # no edge input (the sort model's encoder): the gathered addends and the bias alone
                bias_row = self._p(b["be"])[:p].reshape(1, p)
                acc = o.gather_add(E, p, None, bias_row, ix.zeros(E), adds[0][0] if adds else None, adds[0][1] if adds else None)
                for t, idx in adds[1:]:
                    acc = o.gather_add(E, p, acc, t, idx)
                he = acc
        agg = o.segsum(he, ix.nip, N) if p > 0 else None
        if q > 0:
            adds = [(o.linear(B, q, [(u, row(Wn, p + bb, q))], q), ix.ng)] if c > 0 else []
            hv = o.linear(N, q, [(agg, Wn), (v, row(Wn, p, q))], q, bias=self._p(b["bn"]), adds=adds)
        se = o.segsum(he, ix.gep, B) if p > 0 else None
        sv = o.segsum(hv, ix.gnp, B) if q > 0 else None
        if r > 0:
            hu = o.linear(B, r, [(se, Wg), (sv, row(Wg, p, r)), (u, row(Wg, p + q, r))], r, bias=self._p(b["bg"]))
        return (he, hv, hu), dict(e=e, v=v, u=u, he=he, hv=hv, agg=agg, se=se, sv=sv)

    def forward(self, x):
        """x: the batched NamedTuple of gn.batch (compact torch tensors inside).  Returns (ef, nf, gf) compact device tensors."""
        self.eng.bind_stream()
        if getattr(self, "_ix", None) is None or self._ix[0] is not x.graphs:      # index tensors + sender permutation: once per batch
            ix = _Index(x.graphs)
            zeros_idx = torch.zeros(max(ix.E, ix.N, ix.B, 1), dtype=torch.int32, device=self.eng.torch_device)
            ix.zeros = lambda n: zeros_idx
            self._ix = (x.graphs, ix)
        ix = self._ix[1]
        o = self.ops
        cur = [None if f is None else f.compact for f in (x.ef, x.nf, x.gf)]
        R = (ix.E, ix.N, ix.B)
        saved = []
        for kind, s in self.spec:
            if kind == "block":
                cur, sv = self._block_fwd(s, ix, *cur)
                cur = list(cur)
                saved.append(sv)
            else:
                xs = cur
                a1 = [o.layernorm(xs[k], self._p(s["ln1"][k]["gamma"]), self._p(s["ln1"][k]["beta"]), s["ln1"][k]["eps"], self.eps_mode) for k in range(3)]
                a2 = [o.layernorm(xs[k], self._p(s["ln2"][k]["gamma"]), self._p(s["ln2"][k]["beta"]), s["ln2"][k]["eps"], self.eps_mode) for k in range(3)]
                blk, bsv = self._block_fwd(s["block"], ix, *a1)
                ys, hs = [], []
                for k in range(3):
                    d, f = s["dims"][k], s["ffn"][k]
                    h = o.linear(R[k], 4 * d, [(a2[k], self._p(f["W1"]))], 4 * d, bias=self._p(f["b1"]), relu=True)
                    ys.append(o.linear(R[k], d, [(h, self._p(f["W2"]))], d, bias=self._p(f["b2"]), adds=[(xs[k], None), (blk[k], None)]))
                    hs.append(h)
                saved.append(dict(x=xs, a1=a1, a2=a2, h=hs, blk=bsv))
                cur = ys
        self._saved = (ix, saved)
        return tuple(cur)

    # ------------------------------------------------------------------ backward
    def _block_bwd(self, b, ix, sv, dhe, dhv, dhu):
        """Adjoint of _block_fwd: cotangents of (he, hv, hu) -> cotangents of (e, v, u); parameter gradients accumulated."""
        o = self.ops
        (a, bb, c), (p, q, r) = b["din"], b["dout"]
        E, N, B = ix.E, ix.N, ix.B
        We, Wn, Wg = self._p(b["We"]), self._p(b["Wn"]), self._p(b["Wg"])
        gWe, gWn, gWg = self._g(b["We"]), self._g(b["Wn"]), self._g(b["Wg"])
        row = lambda Wt, k, ld: Wt[k * ld:]
        e, v, u = sv["e"], sv["v"], sv["u"]
        de = dv = du = None

        def acc(cur, R, D, srcs, ldw):      # cur += sum_s x_s W_s  (cur may be None)
            srcs = [(x_, W_) for x_, W_ in srcs if x_ is not None]
            if not srcs:
                return cur
            return o.linear(R, D, srcs, ldw, adds=[] if cur is None else [(cur, None)], out=cur)

        d_se = d_sv = None
        # ---- graph update  hu = [se | sv | u] Wg + bg
        if r > 0 and dhu is not None:
            o.wgrad(sv["se"], dhu, gWg, r)
            o.wgrad(sv["sv"], dhu, row(gWg, p, r), r)
            o.wgrad(u, dhu, row(gWg, p + q, r), r)
            o.colsum(dhu, self._g(b["bg"]))
            WgT = o.transpose(Wg, p + q + c, r, r)      # [r][p+q+c]
            if p > 0:
                d_se = o.linear(B, p, [(dhu, WgT)], p + q + c)
            if q > 0:
                d_sv = o.linear(B, q, [(dhu, WgT[:, p:])], p + q + c)
            if c > 0:
                du = acc(du, B, c, [(dhu, WgT[:, p + q:])], p + q + c)
        # ---- node update  hv = [agg | v | u[g]] Wn + bn ;  total cotangent of hv = dhv + d_sv[graph]
        d_agg = None
        if q > 0 and (dhv is not None or d_sv is not None):
            dv_t = o.gather_add(N, q, dhv, d_sv, ix.ng if d_sv is not None else None)
            o.wgrad(sv["agg"], dv_t, gWn, q)
            o.wgrad(v, dv_t, row(gWn, p, q), q)
            o.wgrad(u, dv_t, row(gWn, p + bb, q), q, idx=ix.ng)
            o.colsum(dv_t, self._g(b["bn"]))
            WnT = o.transpose(Wn, p + bb + c, q, q)     # [q][p+b+c]
            if p > 0:
                d_agg = o.linear(N, p, [(dv_t, WnT)], p + bb + c)
            if bb > 0:
                dv = acc(dv, N, bb, [(dv_t, WnT[:, p:])], p + bb + c)
            if c > 0:      # adjoint of the gather u[graph]: segmented sum over the graph's nodes, then the GEMM on B rows
                du = acc(du, B, c, [(o.segsum(dv_t, ix.gnp, B), WnT[:, p + bb:])], p + bb + c)
        # ---- edge update  he = [e | v[src] | v[dst] | u[g]] We + be ;  total cotangent = dhe + d_agg[dst] + d_se[graph]
        if p > 0 and (dhe is not None or d_agg is not None or d_se is not None):
            de_t = o.gather_add(E, p, dhe, d_agg, ix.dst if d_agg is not None else None, d_se, ix.eg if d_se is not None else None)
            o.wgrad(e, de_t, gWe, p)
            o.wgrad(v, de_t, row(gWe, a, p), p, idx=ix.src)
            o.wgrad(v, de_t, row(gWe, a + bb, p), p, idx=ix.dst)
            o.wgrad(u, de_t, row(gWe, a + 2 * bb, p), p, idx=ix.eg)
            o.colsum(de_t, self._g(b["be"]))
            WeT = o.transpose(We, a + 2 * bb + c, p, p)     # [p][a+2b+c]
            K = a + 2 * bb + c
            if a > 0:
                de = o.linear(E, a, [(de_t, WeT)], K)
            if bb > 0:      # adjoints of the two gathers: segmented sums over senders (permutation) / receivers (CSR), then N-row GEMMs
                s_src = o.segsum(de_t, ix.node_out_ptr32, N, perm=ix.src_perm)
                s_dst = o.segsum(de_t, ix.nip, N)
                dv = acc(dv, N, bb, [(s_src, WeT[:, a:]), (s_dst, WeT[:, a + bb:])], K)
            if c > 0:
                du = acc(du, B, c, [(o.segsum(de_t, ix.gep, B), WeT[:, a + 2 * bb:])], K)
        return de, dv, du

    def backward(self, d_ef=None, d_nf=None, d_gf=None, zero_grad=True):
        """Cotangents of the outputs of the last forward (compact device tensors or None) -> self.grads; returns the cotangents
        of the inputs."""
        assert self._saved is not None, "backward() needs a forward() first"
        ix, saved = self._saved
        o = self.ops
        if zero_grad:
            self.grads.zero_()
        d = [d_ef, d_nf, d_gf]
        R = (ix.E, ix.N, ix.B)
        for (kind, s), sv in zip(reversed(self.spec), reversed(saved)):
            if kind == "block":
                d = list(self._block_bwd(s, ix, sv, *d))
                continue
            dy = d
            dx = [None if t is None else t.clone() for t in dy]      # residual: dx = dy + ...
            da1 = list(self._block_bwd(s["block"], ix, sv["blk"], *dy))
            for k in range(3):
                if dy[k] is None:
                    continue
                dk, f = s["dims"][k], s["ffn"][k]
                x_k, a2, h = sv["x"][k], sv["a2"][k], sv["h"][k]
                # feed-forward branch  y += W2 relu(W1 LN2(x) + b1) + b2
                o.wgrad(h, dy[k], self._g(f["W2"]), dk)
                o.colsum(dy[k], self._g(f["b2"]))
                W2T = o.transpose(self._p(f["W2"]), 4 * dk, dk, dk)        # [d][4d]
                dh = o.linear(R[k], 4 * dk, [(dy[k], W2T)], 4 * dk)
                o.relu_mask(dh, h)
                o.wgrad(a2, dh, self._g(f["W1"]), 4 * dk)
                o.colsum(dh, self._g(f["b1"]))
                W1T = o.transpose(self._p(f["W1"]), dk, 4 * dk, 4 * dk)    # [4d][d]
                da2 = o.linear(R[k], dk, [(dh, W1T)], dk)
                for key, g_ in (("ln2", da2), ("ln1", da1[k])):
                    if g_ is None:
                        continue
                    l = s[key][k]
                    gx = o.layernorm_bwd(x_k, g_, self._p(l["gamma"]), l["eps"], self.eps_mode, dx[k])
                    o.colsum(gx, self._g(l["gamma"]))
                    o.colsum(g_, self._g(l["beta"]))
            d = dx
        return tuple(d)

    # ------------------------------------------------------------------ loss of the reference's training example
    def cross_entropy(self, logits, targets, scale=1.0):
        """Flux.logitcrossentropy over compact rows (examples/sort/sort.jl:76-78): returns (loss: device scalar tensor, dlogits)."""
        R, D = logits.shape
        loss = torch.empty(1, dtype=torch.float32, device=logits.device)
        dl = torch.empty_like(logits)
        check(lib.gnb_logit_cross_entropy(self.eng.ctx, _P(logits), _P(targets), D, R, _P(loss), None))
        check(lib.gnb_logit_cross_entropy_bwd(self.eng.ctx, _P(logits), _P(targets), D, R, float(scale), _P(dl)))
        return loss, dl

    # ------------------------------------------------------------------ optimiser step
    def step(self, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, average=True):
        """Gradient all-reduce over the process group (when torch.distributed is initialised; NCCL on GPUs) + AdamW."""
        allreduce_flat(self.grads, average)
        self.t += 1
        check(lib.gnb_op_adamw(self.eng.ctx, _P(self.params), _P(self.grads), _P(self.m), _P(self.v), self.params.numel(), lr, betas[0], betas[1],
                               eps, weight_decay, self.t))
