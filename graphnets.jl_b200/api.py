"""Host-side mirror of the reference's data API: batch / unbatch / views / checks.

Reference: src/batch.jl, src/unbatch.jl, src/pad.jl, src/unpad.jl, src/views.jl, src/checks.jl,
src/gngraphbatch.jl.  Shapes follow the Julia convention `(D, T, B)` with COLUMN-MAJOR strides
(feature dim contiguous) so user code reads like the reference's; indices are 0-based.

What differs by design: `batch` lowers the dense adjacency on the GPU to a receiver-sorted
COO/CSR (`GNGraphBatch.handle`), and feature tensors live COMPACT on the device
(`(D, E)` / `(D, N)` / `(D, B)` in `flatunpaddedef` order).  The documented padded shapes
(`size(x.ef) == (DE, PN^2, B)`, src/batch.jl:48-50) are served by `Padded`, a lazy view that
scatters into the padded layout only when somebody indexes it.
"""
import collections
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import lib, check
from .engine import get_engine, _ptr

GNData = collections.namedtuple("GNData", "graphs ef nf gf")
_KEYS = ("graphs", "ef", "nf", "gf")


def _fields(t):
    """(; graphs, ef, nf, gf) = t with the reference's key check (src/batch.jl:54)."""
    if isinstance(t, GNData):
        return t
    if hasattr(t, "_asdict"):
        t = t._asdict()
    assert isinstance(t, collections.abc.Mapping), "expected a mapping / named tuple with keys graphs, ef, nf, gf"
    assert set(t.keys()) == set(_KEYS), "keys must be exactly (graphs, ef, nf, gf), got %s" % (sorted(t.keys()),)
    return GNData(t["graphs"], t["ef"], t["nf"], t["gf"])


def _np(a):
    if isinstance(a, torch.Tensor):
        return a.detach().cpu().numpy()
    return np.asarray(a)


def _ndim(a):
    return a.dim() if isinstance(a, torch.Tensor) else np.ndim(a)


def _shape(a):
    return tuple(a.shape) if hasattr(a, "shape") else np.shape(a)


# ----------------------------------------------------------------------------- checks
def _count_edges(adj):
    return int(np.count_nonzero(_np(adj) == 1))     # length(filter(isone, adj_mat))


def _check_counts(adj, ef, nf):
    """checknodeedgecounts (src/checks.jl:42-57)"""
    if ef is not None:
        m = _count_edges(adj)
        assert _shape(ef)[1] == m, "%d != num_edges (%d)" % (_shape(ef)[1], m)
    if nf is not None:
        n = _shape(adj)[0]
        assert _shape(nf)[1] == n, "%d != %d" % (_shape(nf)[1], n)


def checks(graphs, ef, nf, gf):
    """src/checks.jl:1-21.  Raises AssertionError like the reference's @assert."""
    if _is_single(graphs):
        # checkshapes3d / checksamebatchsize3d
        if ef is not None:
            assert _ndim(ef) == 3, "ef must be (C, T, B)"
        if nf is not None:
            assert _ndim(nf) == 3, "nf must be (C, T, B)"
        if gf is not None:
            assert _ndim(gf) == 2, "%d != 2" % _ndim(gf)
        bs = [s for s in (None if ef is None else _shape(ef)[2], None if nf is None else _shape(nf)[2],
                          None if gf is None else _shape(gf)[1]) if s is not None]
        assert all(b == bs[0] for b in bs), "batch sizes differ: %s" % (bs,)
        assert _ndim(graphs) == 2 and _shape(graphs)[0] == _shape(graphs)[1], "adjacency must be (N, N)"
        _check_counts(graphs, ef, nf)
    else:
        assert len(graphs) > 0
        for x in (ef, nf, gf):
            if x is not None:
                assert len(x) == len(graphs), "batch sizes differ"
        for a in graphs:
            assert _ndim(a) == 2 and _shape(a)[0] == _shape(a)[1], "adjacency must be (N, N)"
        for i in range(len(graphs)):
            if ef is not None:
                assert _ndim(ef[i]) == 2, "ef[i] must be (C, T)"
            if nf is not None:
                assert _ndim(nf[i]) == 2, "%d != 2" % _ndim(nf[i])
            if gf is not None:
                assert _ndim(gf[i]) == 1, "gf[i] must be (C,)"
            _check_counts(graphs[i], None if ef is None else ef[i], None if nf is None else nf[i])


def _is_single(graphs):
    """graphs::AbstractMatrix (one structure shared by the batch) vs ::AbstractVector of matrices"""
    if isinstance(graphs, (np.ndarray, torch.Tensor)):
        return _ndim(graphs) == 2
    if isinstance(graphs, (list, tuple)):
        return len(graphs) > 0 and not hasattr(graphs[0], "shape") and np.ndim(graphs[0]) == 1
    raise AssertionError("graphs must be an adjacency matrix or a vector of adjacency matrices")


# ----------------------------------------------------------------------------- GNGraphBatch
def pack_adjacency_bits(mask):
    """uint8 `isone` mask [B][PN][PN] (element (i,j,b) at i + PN*j + PN^2*b) -> GNB_ADJ_BITS words: bit (i + PN*j) of graph b,
    little-endian, every graph padded to whole 32-bit words."""
    B = mask.shape[0]
    flat = np.ascontiguousarray(mask.reshape(B, -1))
    packed = np.packbits(flat, axis=1, bitorder="little")
    pad = (-packed.shape[1]) % 4
    if pad:
        packed = np.concatenate([packed, np.zeros((B, pad), np.uint8)], axis=1)
    return np.ascontiguousarray(packed).view(np.uint32)


class GNGraphBatch:
    """Lowered structure of a batch (reference struct: src/gngraphbatch.jl:1-17).  Instead of
    seven dense broadcaster tensors it holds a device-resident receiver-sorted COO + CSR."""

    def __init__(self, adj_mats, B=None, device=None, _stacked=None):
        if _stacked is not None:
            # fast path: adjacency already stacked as (B, n, n) with a common n (adj[b, i, j])
            stacked = np.asarray(_stacked)
            Badj = stacked.shape[0]
            self.adj_mats = stacked
            self.n_nodes = np.full(Badj, stacked.shape[1], np.int32)
            mask = np.ascontiguousarray((stacked == 1).transpose(0, 2, 1)).astype(np.uint8)
        else:
            adj_np = [_np(a) for a in adj_mats]
            for a in adj_np:
                assert a.ndim == 2 and a.shape[0] == a.shape[1], "adjacency must be (N, N)"
            assert len(adj_np) > 0
            self.adj_mats = list(adj_mats)
            Badj = len(adj_np)
            self.n_nodes = np.array([a.shape[0] for a in adj_np], np.int32)
            mask = None
        self.single = Badj == 1                 # length(graphs.adj_mats) == 1 (src/unbatch.jl:13)
        self.B = int(B) if B is not None else Badj
        assert Badj == 1 or Badj == self.B
        self.node_block_size = int(self.n_nodes.max())       # PN (src/gngraphbatch.jl:35)
        self.edge_block_size = self.node_block_size ** 2     # PE (src/gngraphbatch.jl:36)
        PN = self.node_block_size
        # padadjmats (src/pad.jl:1-10) as a uint8 `isone` mask, element (i,j,b) at i + PN*j + PN^2*b
        if mask is None:
            mask = np.zeros((Badj, PN, PN), np.uint8)
            for b, a in enumerate(adj_np):
                n = a.shape[0]
                mask[b, :n, :n] = (a == 1).T
        self._mask = mask
        self.engine = get_engine(device)
        self.engine.bind_stream()
        h = C.c_void_p()
        # the `isone` mask goes up bit-packed (GNB_ADJ_BITS): 8x fewer PCIe bytes than one byte per cell
        bits = pack_adjacency_bits(mask)
        check(lib.gnb_graph_lower(self.engine.ctx, bits.ctypes.data_as(C.c_void_p), _lib.ADJ_BITS, 0,
                                  self.n_nodes.ctypes.data_as(_lib.i32p), PN, Badj, self.B, C.byref(h)))
        self._finish(h)

    def _finish(self, h):
        self.handle = h
        E, N = C.c_int64(), C.c_int64()
        check(lib.gnb_graph_counts(self.handle, C.byref(E), C.byref(N), None, None))
        self.E, self.N = E.value, N.value
        self._index = None

    @classmethod
    def from_coo(cls, src, dst, graph_edge_ptr, n_nodes, PN=None, device=None):
        """Lower a batch given as COO edge lists (C ABI: gnb_graph_from_coo): graph b owns edges
        [graph_edge_ptr[b], graph_edge_ptr[b+1]) with local sender ids `src` and receiver ids `dst`; each graph's edges
        strictly ascending in src + PN*dst (receiver-major = the order of `findall(isone, adj[:])`, src/pad.jl:30, in which the
        reference takes edge features).  The index equals the one `batch` builds from the equivalent adjacency matrices."""
        self = cls.__new__(cls)
        src = np.ascontiguousarray(src, np.int32)
        dst = np.ascontiguousarray(dst, np.int32)
        ep = np.ascontiguousarray(graph_edge_ptr, np.int32)
        self.n_nodes = np.ascontiguousarray(n_nodes, np.int32)
        assert ep.ndim == 1 and ep.size == self.n_nodes.size + 1 and self.n_nodes.size > 0
        assert src.shape == dst.shape == (int(ep[-1]),), "edge list length != graph_edge_ptr[-1]"
        self.B = int(self.n_nodes.size)
        self.single = self.B == 1
        self.node_block_size = int(PN) if PN else max(int(self.n_nodes.max()), 1)
        self.edge_block_size = self.node_block_size ** 2
        self._mask = None
        self._coo = (src, dst, ep)
        self.engine = get_engine(device)
        self.engine.bind_stream()
        h = C.c_void_p()
        check(lib.gnb_graph_from_coo(self.engine.ctx, src.ctypes.data_as(C.c_void_p), dst.ctypes.data_as(C.c_void_p), 0,
                                     ep.ctypes.data_as(_lib.i32p), self.n_nodes.ctypes.data_as(_lib.i32p),
                                     self.node_block_size, self.B, C.byref(h)))
        self._finish(h)
        return self

    def __getattr__(self, name):
        # `graphs.adj_mats` of a batch that came in as edge lists: dense matrices are built only when somebody asks for them
        if name == "adj_mats" and self.__dict__.get("_coo") is not None:
            self._dense_mask()
            return self.__dict__["adj_mats"]
        raise AttributeError(name)

    def _dense_mask(self):
        """uint8 `isone` mask [b][j][i] (built on demand for graphs that came in as edge lists)"""
        if self._mask is None:
            src, dst, ep = self._coo
            PN = self.node_block_size
            mask = np.zeros((self.B, PN, PN), np.uint8)
            b = np.repeat(np.arange(self.B), np.diff(ep))
            mask[b, dst, src] = 1
            self._mask = mask
            self.adj_mats = [np.ascontiguousarray(mask[i, :n, :n].T) for i, n in enumerate(self.n_nodes)]
        return self._mask

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                lib.gnb_graph_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    @property
    def padded_adj_mats(self):
        """(PN, PN, B') Float32 like src/pad.jl:1-10 (B' = number of distinct structures)."""
        return np.asfortranarray(self._dense_mask().transpose(2, 1, 0).astype(np.float32))

    def index(self):
        """Host copy of the lowered index (dict of int32 numpy arrays)."""
        if self._index is None:
            E, N, B = self.E, self.N, self.B
            out = dict(edge_src=np.empty(E, np.int32), edge_dst=np.empty(E, np.int32),
                       edge_slot=np.empty(E, np.int32), edge_graph=np.empty(E, np.int32),
                       graph_edge_ptr=np.empty(B + 1, np.int32), graph_node_ptr=np.empty(B + 1, np.int32),
                       node_in_ptr=np.empty(N + 1, np.int32))
            p = lambda k: out[k].ctypes.data_as(C.c_void_p)
            check(lib.gnb_graph_export_host(self.engine.ctx, self.handle, p("edge_src"), p("edge_dst"),
                                            p("edge_slot"), p("edge_graph"), p("graph_edge_ptr"),
                                            p("graph_node_ptr"), p("node_in_ptr")))
            self._index = out
        return self._index

    @property
    def uniform_nodes(self):
        return bool((self.n_nodes == self.node_block_size).all())

    # reference field names kept for drop-in reads
    @property
    def flat_node_unpadder(self):
        mask = np.zeros(self.B * self.node_block_size, bool)
        nn = self.n_nodes if len(self.n_nodes) == self.B else np.repeat(self.n_nodes, self.B)
        for b, n in enumerate(nn):
            mask[b * self.node_block_size:b * self.node_block_size + n] = True
        return mask

    @property
    def flat_edge_unpadder(self):
        """Mask of ACTIVE padded slots, graph-major.  (The reference, src/gngraphbatch.jl:125-134,
        copies the unpadded matrix into the first n^2 slots, which is only the active-slot mask
        when n == PN; the compact order used here is the one `unpadef` defines, src/unpad.jl:6-10.)"""
        idx = self.index()
        mask = np.zeros(self.B * self.edge_block_size, bool)
        mask[idx["edge_graph"].astype(np.int64) * self.edge_block_size + idx["edge_slot"]] = True
        return mask


# ----------------------------------------------------------------------------- Padded
class Padded:
    """A batched feature tensor: compact device storage + the padded `(D, T, B)` face of the
    reference (`ef (DE, PN^2, B)`, `nf (DN, PN, B)`, `gf (DG, 1, B)`, src/batch.jl:44-50)."""

    def __init__(self, kind, compact, graphs):
        assert kind in ("e", "n", "g")
        self.kind, self.compact, self.graphs = kind, compact, graphs
        self._padded = None

    @property
    def D(self):
        return int(self.compact.shape[1])

    @property
    def shape(self):
        g = self.graphs
        T = {"e": g.edge_block_size, "n": g.node_block_size, "g": 1}[self.kind]
        return (self.D, T, g.B)

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    def padded(self):
        """Materialise (cached) the padded tensor, shape (D, T, B) with column-major strides."""
        if self._padded is None:
            g = self.graphs
            D, T, B = self.shape
            if self.kind == "g":
                flat = self.compact.view(B, 1, D)
            elif self.kind == "n" and g.uniform_nodes:
                flat = self.compact.view(B, T, D)
            else:
                flat = torch.empty((B, T, D), dtype=torch.float32, device=self.compact.device)
                g.engine.bind_stream()
                fn = lib.gnb_pad_edges if self.kind == "e" else lib.gnb_pad_nodes
                check(fn(g.engine.ctx, g.handle, _ptr(self.compact), D, _ptr(flat)))
            self._padded = flat.permute(2, 1, 0)
        return self._padded

    def __getitem__(self, idx):
        return self.padded()[idx]

    def cpu(self):
        return self.padded().cpu()

    def numpy(self):
        return self.padded().cpu().numpy()

    def __repr__(self):
        return "Padded(%s, shape=%s, compact=%s)" % (self.kind, self.shape, tuple(self.compact.shape))


def _to_compact(x, device):
    """Julia-shaped (D, T[, B]) array -> contiguous device tensor [B*T][D] (feature contiguous)."""
    if isinstance(x, torch.Tensor):
        t = x.to(device=device, dtype=torch.float32)
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(x, np.float32).transpose())).to(device)
        return t.reshape(-1, t.shape[-1])
    perm = tuple(range(t.dim() - 1, -1, -1))
    t = t.permute(perm).contiguous()
    return t.reshape(-1, t.shape[-1])


# ----------------------------------------------------------------------------- batch
def batch(t, device=None):
    """batch(t::NamedTuple) (src/batch.jl:53-64): validate, lower the graphs, move features to
    the device.  Returns GNData(graphs::GNGraphBatch, ef, nf, gf) with `Padded` features."""
    graphs, ef, nf, gf = _fields(t)
    assert ef is not None or nf is not None or gf is not None
    checks(graphs, ef, nf, gf)
    if _is_single(graphs):
        B = _shape(ef)[2] if ef is not None else (_shape(nf)[2] if nf is not None else _shape(gf)[1])
        gb = GNGraphBatch([graphs], B=B, device=device)
        dev = gb.engine.torch_device
        c_ef = None if ef is None else _to_compact(ef, dev)
        c_nf = None if nf is None else _to_compact(nf, dev)
        c_gf = None if gf is None else _to_compact(gf, dev)
    else:
        gb = GNGraphBatch(list(graphs), device=device)
        dev = gb.engine.torch_device

        def cat(xs):
            if xs is None:
                return None
            if all(isinstance(x, torch.Tensor) for x in xs):
                return torch.cat([x.to(dev, torch.float32).t() if x.dim() == 2 else x.to(dev, torch.float32)[None, :]
                                  for x in xs], 0).contiguous()
            rows = [np.asarray(x, np.float32).T if np.ndim(x) == 2 else np.asarray(x, np.float32)[None, :] for x in xs]
            return torch.from_numpy(np.ascontiguousarray(np.concatenate(rows, 0))).to(dev)
        c_ef, c_nf, c_gf = cat(ef), cat(nf), cat(gf)
    wrap = lambda k, c: None if c is None else Padded(k, c, gb)
    return GNData(gb, wrap("e", c_ef), wrap("n", c_nf), wrap("g", c_gf))


def batch_compact(adj_stacked, ef=None, nf=None, gf=None, device=None):
    """Fast constructor for large same-size batches: `adj_stacked` (B, n, n) and features already
    compact `[rows][D]` (numpy or torch, host or device).  Same result as `batch` on the equivalent
    vector-mode input, without the per-graph Python loop."""
    gb = GNGraphBatch(None, device=device, _stacked=adj_stacked)
    dev = gb.engine.torch_device

    def mv(x, rows):
        if x is None:
            return None
        t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x, np.float32))
        t = t.to(dev, torch.float32).contiguous()
        assert t.dim() == 2 and t.shape[0] == rows, "expected %d rows, got %s" % (rows, tuple(t.shape))
        return t
    wrap = lambda k, c: None if c is None else Padded(k, c, gb)
    return GNData(gb, wrap("e", mv(ef, gb.E)), wrap("n", mv(nf, gb.N)), wrap("g", mv(gf, gb.B)))


def batch_coo(src, dst, graph_edge_ptr, n_nodes, ef=None, nf=None, gf=None, PN=None, device=None):
    """`batch` for graphs given as COO edge lists (`GNGraphBatch.from_coo`) and compact features `[rows][D]` in the same
    (graph, receiver, sender) order - no dense adjacency anywhere, neither on the host nor on the PCIe bus."""
    gb = GNGraphBatch.from_coo(src, dst, graph_edge_ptr, n_nodes, PN=PN, device=device)
    dev = gb.engine.torch_device

    def mv(x, rows):
        if x is None:
            return None
        t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x, np.float32))
        t = t.to(dev, torch.float32).contiguous()
        assert t.dim() == 2 and t.shape[0] == rows, "expected %d rows, got %s" % (rows, tuple(t.shape))
        return t
    wrap = lambda k, c: None if c is None else Padded(k, c, gb)
    return GNData(gb, wrap("e", mv(ef, gb.E)), wrap("n", mv(nf, gb.N)), wrap("g", mv(gf, gb.B)))


# ----------------------------------------------------------------------------- unbatch / views
def _compact(x):
    return None if x is None else x.compact


def unbatch(t):
    """unbatch(t::NamedTuple) (src/unbatch.jl:6-48).  Returned tensors are VIEWS of the batched
    (compact) storage, like the reference's views of the padded arrays (src/unpad.jl)."""
    graphs, ef, nf, gf = _fields(t)
    assert ef is not None or nf is not None or gf is not None
    g = graphs
    B = g.B
    if g.single:
        m = g.E // B if B else 0
        n = g.N // B if B else 0
        return GNData(
            g.adj_mats[0],
            None if ef is None else ef.compact.view(B, m, ef.D).permute(2, 1, 0),   # (DE, m, B)
            None if nf is None else nf.compact.view(B, n, nf.D).permute(2, 1, 0),   # (DN, n, B)
            None if gf is None else gf.compact.t(),                                  # (DG, B)
        )
    idx = g.index()
    ep, npz = idx["graph_edge_ptr"], idx["graph_node_ptr"]
    return GNData(
        g.adj_mats,
        None if ef is None else [ef.compact[ep[b]:ep[b + 1]].t() for b in range(B)],
        None if nf is None else [nf.compact[npz[b]:npz[b + 1]].t() for b in range(B)],
        None if gf is None else [gf.compact[b] for b in range(B)],
    )


def efview(t, d1, d2, d3):
    """efview(t, d1, d2, d3) (src/views.jl:6-31): unpadded edge features; 0-based indices/slices."""
    f = t._asdict() if hasattr(t, "_asdict") else t
    assert {"graphs", "ef"} <= set(f.keys())
    g, ef = f["graphs"], f["ef"]
    if ef is None:
        return None
    if g.single:
        m = g.E // g.B
        return ef.compact.view(g.B, m, ef.D).permute(2, 1, 0)[d1, d2, d3]
    assert isinstance(d3, (int, np.integer)), "d3 must be an integer for batches of different structure"
    ep = g.index()["graph_edge_ptr"]
    return ef.compact[ep[d3]:ep[d3 + 1]].t()[d1, d2]


def nfview(t, d1, d2, d3):
    """nfview (src/views.jl:38-61)"""
    f = t._asdict() if hasattr(t, "_asdict") else t
    assert {"graphs", "nf"} <= set(f.keys())
    g, nf = f["graphs"], f["nf"]
    if nf is None:
        return None
    if g.single:
        n = g.N // g.B
        return nf.compact.view(g.B, n, nf.D).permute(2, 1, 0)[d1, d2, d3]
    assert isinstance(d3, (int, np.integer)), "d3 must be an integer for batches of different structure"
    npz = g.index()["graph_node_ptr"]
    return nf.compact[npz[d3]:npz[d3 + 1]].t()[d1, d2]


def gfview(t, d1, d2):
    """gfview (src/views.jl:68-78): gf[d1, 1, d2]"""
    f = t._asdict() if hasattr(t, "_asdict") else t
    assert {"graphs", "gf"} <= set(f.keys())
    gf = f["gf"]
    if gf is None:
        return None
    return gf.compact.t()[d1, d2]


def flatunpaddednf(t):
    """(DN, N) view over all real nodes (src/views.jl:80-88) - free in the compact layout."""
    return _fields(t).nf.compact.t()


def flatunpaddedef(t):
    """(DE, E) view over all active edges (src/views.jl:90-98) - free in the compact layout."""
    return _fields(t).ef.compact.t()


# ----------------------------------------------------------------------------- loss over the compact views
def logitcrossentropy(yhat, y):
    """Flux.logitcrossentropy(flatunpaddednf(yhat_batch), flatunpaddednf(target_batch)) as in the reference's training example
    (examples/sort/sort.jl:76-78).  `yhat`, `y`: `Padded` features (or compact `[rows][D]` device tensors) of the same shape;
    returns a 0-dim device tensor: mean over the real rows of -sum_d y * logsoftmax(yhat)."""
    a = yhat.compact if isinstance(yhat, Padded) else yhat
    b = y.compact if isinstance(y, Padded) else y
    assert tuple(a.shape) == tuple(b.shape) and a.dim() == 2, "logits / targets must be compact [rows][D] of equal shape"
    a, b = a.contiguous(), b.to(a.device, torch.float32).contiguous()
    eng = yhat.graphs.engine if isinstance(yhat, Padded) else get_engine(a.device)
    eng.bind_stream()
    out = torch.empty((), dtype=torch.float32, device=a.device)
    check(lib.gnb_logit_cross_entropy(eng.ctx, _ptr(a), _ptr(b), int(a.shape[1]), int(a.shape[0]), _ptr(out), None))
    return out


# ----------------------------------------------------------------------------- edge collapsing
def collapsef(t):
    """collapsef (src/gngraphbatch.jl:83-85): (DE, PN(PN+1)/2, B), mean of slots (i,j) and (j,i)."""
    g, ef = _fields(t).graphs, _fields(t).ef
    D, PE, B = ef.shape
    PN = g.node_block_size
    C_ = PN * (PN + 1) // 2
    padded = ef.padded().permute(2, 1, 0).contiguous()       # [B][PE][D]
    out = torch.empty((B, C_, D), dtype=torch.float32, device=padded.device)
    g.engine.bind_stream()
    check(lib.gnb_collapse_edges(g.engine.ctx, g.handle, _ptr(padded), D, _ptr(out)))
    return out.permute(2, 1, 0)


def _collapse_idxs(g):
    """getcollapsededgeidxs (src/gngraphbatch.jl:60-65): per structure, lower-triangular coordinates
    (column-major, i >= j) whose adjacency entry is one."""
    PN = g.node_block_size
    out = []
    for b in range(g._dense_mask().shape[0]):
        a = g._dense_mask()[b].T        # a[i, j]
        c = 0
        sel = []
        for j in range(PN):
            for i in range(j, PN):
                if a[i, j] == 1:
                    sel.append(c)
                c += 1
        out.append(np.array(sel, np.int64))
    return out


def unpaddedcollapsedef(t):
    """src/gngraphbatch.jl:87-109"""
    g = _fields(t).graphs
    col = collapsef(t)
    idxs = _collapse_idxs(g)
    res = []
    for b in range(g.B):
        sel = torch.from_numpy(idxs[0 if len(idxs) == 1 else b]).to(col.device)
        res.append(col[:, :, b][:, sel])
    return res


def flatunpaddedcollapsedef(t):
    """reduce(hcat, unpaddedcollapsedef(graph)) (src/gngraphbatch.jl:111-113)"""
    return torch.cat(unpaddedcollapsedef(t), dim=1)
