"""CPU oracle for the GraphNets.jl GNBlock / GNCore forward path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this file; it is used by
`tests/`, by `__graft_entry__.smoke()` and by `bench.py`'s cpu_baseline / `--impl reference`
legs as the checker / timed CPU baseline.

Two independent restatements of the reference (citations relative to /root/reference):

* "sparse" (float64): the mathematical semantics of the forward on compact COO data
  (SURVEY.md Appendix A).  Ground truth for tolerance checks.
* "dense mirror" (float32): op-for-op restatement of the reference formulation - padded
  (D, PN^2, B) tensors, seven dense 0/1 broadcaster tensors, every gather/aggregate a
  batched matmul (src/gngraphbatch.jl, src/edgefninput.jl, src/nodefninput.jl,
  src/graphfninput.jl).  Used to prove the sparse oracle equals the reference formulation
  and as the timed "reference algorithm on CPU" baseline.

PARITY PINNING.  The reference cannot be executed here (no Julia toolchain) and its test
suite holds no numeric forward vectors.  What IS pinned against the reference's own
known-answer material: the edge-index convention / broadcaster matrices
(test/runtests.jl:480-508, 655-682), the collapse identity (test/runtests.jl:41-50), the
batch/unbatch round trip (:362-365, :386-389) and all output shapes.  The floating-point
forward is "parity unpinned": it rests on this restatement of Flux 0.14 `Dense`,
`LayerNorm`, `relu` (third-party, un-vendored, Project.toml:11-15) - see DESIGN.md.

Array conventions: Julia arrays are column-major (D, T, B).  Here every feature array is the
C-order transpose: padded `[B][T][D]`, compact `[E][D]` / `[N][D]` / `[B][D]`, feature dim
contiguous - the same bytes.  Weights are numpy `(out, in)` matrices like Flux `Dense.weight`.
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------------------
# structure lowering (src/pad.jl:1-10, src/gngraphbatch.jl:113-134,194-211)
# --------------------------------------------------------------------------------------


def padadjmats(adj_mats):
    """src/pad.jl:1-10.  Zero-pad to the largest node count; returns float32 adj[b, i, j]."""
    B = len(adj_mats)
    PN = max(a.shape[0] for a in adj_mats)
    out = np.zeros((B, PN, PN), np.float32)
    for b, a in enumerate(adj_mats):
        n = a.shape[0]
        out[b, :n, :n] = a
    return out


def active_slots(padded_adj_b):
    """findall(isone, view(adj, :)) (src/pad.jl:30) - 0-based column-major flat slots
    k = i + PN*j of the entries equal to one, ascending."""
    flat = np.asarray(padded_adj_b).T.reshape(-1)  # column-major flattening of adj[i, j]
    return np.nonzero(flat == 1)[0].astype(np.int64)


def lower(adj_mats):
    """Receiver-sorted COO + CSR of a batch.  Order of edges = graph-major, then ascending
    padded flat slot (src/gngraphbatch.jl:125-134 `flat_edge_unpadder`), sender = row i,
    receiver = column j (src/gngraphbatch.jl:194-211).  Node ids are GLOBAL compact ids
    (graph_node_ptr[b] + local index)."""
    padded = padadjmats(adj_mats)
    B, PN, _ = padded.shape
    n_nodes = np.array([a.shape[0] for a in adj_mats], np.int64)
    graph_node_ptr = np.zeros(B + 1, np.int64)
    graph_node_ptr[1:] = np.cumsum(n_nodes)
    src, dst, slot, eg = [], [], [], []
    graph_edge_ptr = np.zeros(B + 1, np.int64)
    for b in range(B):
        k = active_slots(padded[b])
        i = k % PN
        j = k // PN
        src.append(graph_node_ptr[b] + i)
        dst.append(graph_node_ptr[b] + j)
        slot.append(k)
        eg.append(np.full(k.shape, b, np.int64))
        graph_edge_ptr[b + 1] = graph_edge_ptr[b] + k.size
    src = np.concatenate(src) if src else np.zeros(0, np.int64)
    dst = np.concatenate(dst) if dst else np.zeros(0, np.int64)
    slot = np.concatenate(slot) if slot else np.zeros(0, np.int64)
    eg = np.concatenate(eg) if eg else np.zeros(0, np.int64)
    N = int(graph_node_ptr[-1])
    node_in_ptr = np.zeros(N + 1, np.int64)
    np.add.at(node_in_ptr, dst + 1, 1)
    node_in_ptr = np.cumsum(node_in_ptr)
    node_graph = np.repeat(np.arange(B, dtype=np.int64), n_nodes)
    return dict(
        B=B, PN=PN, E=int(src.size), N=N, n_nodes=n_nodes,
        edge_src=src, edge_dst=dst, edge_slot=slot, edge_graph=eg,
        graph_edge_ptr=graph_edge_ptr, graph_node_ptr=graph_node_ptr,
        node_in_ptr=node_in_ptr, node_graph=node_graph,
    )


# --------------------------------------------------------------------------------------
# dense broadcasters (src/gngraphbatch.jl:136-211) - op-for-op, float32
# stored as bc[b] = Julia matrix [:, :, b] with numpy shape (rows, cols)
# --------------------------------------------------------------------------------------


def node2edge_broadcaster(padded, dst=False):
    """src/gngraphbatch.jl:194-211.  (PN, PN^2) per graph; column k one-hot at the sender
    row index (src) or receiver column index (dst, via `transpose` of the index matrix)."""
    B, PN, _ = padded.shape
    idx = np.repeat(np.arange(1, PN + 1)[:, None], PN, axis=1)  # repeat(1:PN, 1, PN): idx[i,j]=i
    if dst:
        idx = idx.T
    out = np.zeros((B, PN, PN * PN), np.float32)
    for b in range(B):
        masked = padded[b] * idx
        flat = masked.T.reshape(-1)
        active_idx = np.nonzero(flat != 0)[0]
        active = flat[active_idx]
        for k, v in zip(active_idx, active):
            out[b, int(v) - 1, k] += 1.0  # onehotbatch + scatter!(+)
    return out


def graph2edge_broadcaster(padded):
    B, PN, _ = padded.shape
    out = np.zeros((B, 1, PN * PN), np.float32)
    for b in range(B):
        out[b, 0, active_slots(padded[b])] = 1.0
    return out


def edge2node_broadcaster(padded):
    """src/gngraphbatch.jl:158-170: block j of rows holds adjacency column j."""
    B, PN, _ = padded.shape
    out = np.zeros((B, PN * PN, PN), np.float32)
    for b in range(B):
        for j in range(PN):
            out[b, PN * j:PN * j + PN, j] = padded[b, :, j]
    return out


def graph2node_broadcaster(adj_mats, PN):
    out = np.zeros((len(adj_mats), 1, PN), np.float32)
    for b, a in enumerate(adj_mats):
        out[b, 0, :a.shape[0]] = 1.0
    return out


def edge2graph_broadcaster(padded):
    B, PN, _ = padded.shape
    out = np.zeros((B, PN * PN, 1), np.float32)
    for b in range(B):
        out[b, :, 0] = padded[b].T.reshape(-1)
    return out


def node2graph_broadcaster(adj_mats, PN):
    out = np.zeros((len(adj_mats), PN, 1), np.float32)
    for b, a in enumerate(adj_mats):
        out[b, :a.shape[0], 0] = 1.0
    return out


def flat_node_unpadder(adj_mats, PN):
    """src/gngraphbatch.jl:113-123"""
    mask = np.zeros(len(adj_mats) * PN, bool)
    for b, a in enumerate(adj_mats):
        mask[b * PN:b * PN + a.shape[0]] = True
    return mask


def flat_edge_unpadder(adj_mats, PE):
    """src/gngraphbatch.jl:125-134.  NOTE: the reference copies `view(adj_mat, :)` of the
    UNPADDED matrix into the first n^2 slots, which equals the padded-slot mask only when
    n == PN; for smaller graphs it is a reference quirk.  The padded-slot mask (what
    unpadef uses, src/unpad.jl:6-10) is what the compact order follows."""
    mask = np.zeros(len(adj_mats) * PE, bool)
    for b, a in enumerate(adj_mats):
        v = (np.asarray(a).T.reshape(-1) != 0)
        mask[b * PE:b * PE + v.size] = v
    return mask


class DenseBatch:
    """GNGraphBatch(adj_mats) (src/gngraphbatch.jl:33-54), numpy float32."""

    def __init__(self, adj_mats):
        self.adj_mats = adj_mats
        self.padded = padadjmats(adj_mats)
        self.B, self.PN, _ = self.padded.shape
        self.PE = self.PN * self.PN
        self.src = node2edge_broadcaster(self.padded)
        self.dst = node2edge_broadcaster(self.padded, dst=True)
        self.g2e = graph2edge_broadcaster(self.padded)
        self.e2n = edge2node_broadcaster(self.padded)
        self.g2n = graph2node_broadcaster(adj_mats, self.PN)
        self.e2g = edge2graph_broadcaster(self.padded)
        self.n2g = node2graph_broadcaster(adj_mats, self.PN)


# --------------------------------------------------------------------------------------
# padding (src/pad.jl:12-67) / unpadding (src/unpad.jl) on [B][T][D] arrays
# --------------------------------------------------------------------------------------


def padef(adj_mats, efs, DE):
    """src/pad.jl:48-63: scatter each graph's (m_b, DE) rows into the PN^2 slots."""
    padded = padadjmats(adj_mats)
    B, PN, _ = padded.shape
    out = np.zeros((B, PN * PN, DE), np.float32)
    for b in range(B):
        k = active_slots(padded[b])
        out[b, k, :] += np.asarray(efs[b], np.float32)
    return out


def padnf(adj_mats, nfs, DN):
    B = len(adj_mats)
    PN = max(a.shape[0] for a in adj_mats)
    out = np.zeros((B, PN, DN), np.float32)
    for b, nf in enumerate(nfs):
        out[b, :nf.shape[0], :] = nf
    return out


def unpadef(adj_mats, ef_padded):
    padded = padadjmats(adj_mats)
    return [ef_padded[b, active_slots(padded[b]), :] for b in range(len(adj_mats))]


def unpadnf(adj_mats, nf_padded):
    return [nf_padded[b, :a.shape[0], :] for b, a in enumerate(adj_mats)]


# --------------------------------------------------------------------------------------
# layers
# --------------------------------------------------------------------------------------

EPS_SQRT_VAR_EPS2 = 0   # (x-mu)/sqrt(var + eps^2)   Flux >= 0.14 `normalise` (default here)
EPS_STD_PLUS_EPS = 1    # (x-mu)/(std + eps)         older Flux
EPS_SQRT_VAR_EPS = 2    # (x-mu)/sqrt(var + eps)     PyTorch convention


def layernorm(x, gamma, beta, eps=1e-5, eps_mode=EPS_SQRT_VAR_EPS2):
    """Flux LayerNorm(d) over the feature dim (src/gngraphnorm.jl:13-15,22-24); uncorrected
    variance; affine scale/bias.  x: [..., d]."""
    dt = x.dtype
    mu = x.mean(axis=-1, keepdims=True)
    xc = x - mu
    var = (xc * xc).mean(axis=-1, keepdims=True)
    eps = dt.type(eps)
    if eps_mode == EPS_SQRT_VAR_EPS2:
        den = np.sqrt(var + eps * eps)
    elif eps_mode == EPS_STD_PLUS_EPS:
        den = np.sqrt(var) + eps
    else:
        den = np.sqrt(var + eps)
    return (xc / den) * gamma.astype(dt) + beta.astype(dt)


def round_bf16(a):
    """Round-to-nearest-even to bfloat16 precision (8 significand bits), returned in float64."""
    a = np.ascontiguousarray(a, np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    r = (((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16).astype(np.uint32)
    return r.view(np.float32).astype(np.float64)


# When set (only by `forward_sparse(..., bf16_operands=True)`), every Dense rounds BOTH matmul operands to bfloat16 and
# accumulates in float64: the arithmetic model of a bf16 tensor-core GEMM with exact accumulation.  Used by the -m gpu tests
# of the wide tensor path to tell kernel error from the error bf16 operands carry by construction.
_BF16_OPERANDS = False


def dense(x, W, b, relu=False):
    """Flux Dense: sigma.(W*x .+ b), applied to the trailing (feature) dim of x."""
    if _BF16_OPERANDS:
        y = round_bf16(x) @ round_bf16(W).T + np.asarray(b, np.float64)
    else:
        y = x @ W.astype(x.dtype).T + b.astype(x.dtype)
    return np.maximum(y, 0) if relu else y


def _cat(parts, rows, dt):
    parts = [p for p in parts if p is not None]
    if not parts:
        return np.zeros((rows, 0), dt)
    return np.concatenate(parts, axis=-1)


def gnblock_sparse(p, g, ef, nf, gf, dt=np.float64):
    """GNBlock forward on compact data (src/gnblock.jl:63-69,
    src/edgefninput.jl:1-48, src/nodefninput.jl:1-25, src/graphfninput.jl:1-14).
    p: dict(We,be,Wn,bn,Wg,bg); any of ef/nf/gf may be None (`nothing`).
    Returns (h_e [E][DE'], h_v [N][DN'], h_u [B][DG']); zero-width outputs -> None."""
    E, N, B = g["E"], g["N"], g["B"]
    cast = lambda a: None if a is None else np.asarray(a, dt)
    ef, nf, gf = cast(ef), cast(nf), cast(gf)
    src, dst, eg, ng = g["edge_src"], g["edge_dst"], g["edge_graph"], g["node_graph"]
    xe = _cat([ef,
               None if nf is None else nf[src],
               None if nf is None else nf[dst],
               None if gf is None else gf[eg]], E, dt)
    h_e = dense(xe, p["We"], p["be"])
    agg = np.zeros((N, h_e.shape[1]), dt)
    np.add.at(agg, dst, h_e)
    xv = _cat([agg, nf, None if gf is None else gf[ng]], N, dt)
    h_v = dense(xv, p["Wn"], p["bn"])
    se = np.zeros((B, h_e.shape[1]), dt)
    np.add.at(se, eg, h_e)
    sv = np.zeros((B, h_v.shape[1]), dt)
    np.add.at(sv, ng, h_v)
    xu = _cat([se, sv, gf], B, dt)
    h_u = dense(xu, p["Wg"], p["bg"])
    z = lambda a: None if a.shape[-1] == 0 else a   # zerodim2nothing, src/gnblock.jl:71-78
    return z(h_e), z(h_v), z(h_u)


def gnblock_dense(p, d: DenseBatch, ef, nf, gf):
    """GNBlock forward in the reference formulation, float32, padded [B][T][D] tensors.
    Julia `batched_mul(F (D,T,B), M (T,T',B))` == here `M[b].T @ F[b]`."""
    f32 = np.float32
    bm = lambda F, M: np.einsum("btd,btu->bud", F, M, dtype=f32, optimize=True)
    parts = []
    if ef is not None:
        parts.append(ef)
    if nf is not None:
        parts.append(bm(nf, d.src))      # edgefninput.jl:4
        parts.append(bm(nf, d.dst))      # edgefninput.jl:5
    if gf is not None:
        parts.append(bm(gf, d.g2e))      # edgefninput.jl:6
    xe = np.concatenate(parts, axis=-1)
    h_e = dense(xe.astype(f32), p["We"].astype(f32), p["be"].astype(f32))
    parts = [bm(h_e, d.e2n)]             # nodefninput.jl:3
    if nf is not None:
        parts.append(nf)
    if gf is not None:
        parts.append(bm(gf, d.g2n))      # nodefninput.jl:5
    h_v = dense(np.concatenate(parts, axis=-1), p["Wn"].astype(f32), p["bn"].astype(f32))
    parts = [bm(h_e, d.e2g), bm(h_v, d.n2g)]   # graphfninput.jl:3-4
    if gf is not None:
        parts.append(gf)
    h_u = dense(np.concatenate(parts, axis=-1), p["Wg"].astype(f32), p["bg"].astype(f32))
    z = lambda a: None if a.shape[-1] == 0 else a
    return z(h_e), z(h_v), z(h_u)


def feedforward(x, f):
    """src/gnfeedforward.jl:27-31: Dense(d=>4d, relu) -> Dense(4d=>d) -> Dropout(identity)."""
    return dense(dense(x, f["W1"], f["b1"], relu=True), f["W2"], f["b2"])


def gncore(c, block_fn, ef, nf, gf, eps_mode=EPS_SQRT_VAR_EPS2):
    """src/gncore.jl:56-68: (x + block(gn1(x))) + ffwd(gn2(x)), per entity kind.
    c: dict(block=..., ffn=[e,n,g], ln1=[e,n,g], ln2=[e,n,g]); ln = dict(gamma,beta,eps)."""
    xs = [ef, nf, gf]
    ln = lambda l, x: layernorm(x, l["gamma"], l["beta"], l.get("eps", 1e-5), eps_mode)
    n1 = [ln(c["ln1"][i], xs[i]) for i in range(3)]
    n2 = [ln(c["ln2"][i], xs[i]) for i in range(3)]
    blk = block_fn(c["block"], n1[0], n1[1], n1[2])
    ff = [feedforward(n2[i], c["ffn"][i]) for i in range(3)]
    return tuple((xs[i] + blk[i]) + ff[i] for i in range(3))


def forward_sparse(layers, g, ef, nf, gf, eps_mode=EPS_SQRT_VAR_EPS2, dt=np.float64, bf16_operands=False):
    """Sequential model: list of ("block", params) / ("core", params) (GNCoreList is a left
    fold, src/gncorelist.jl:43-45).  bf16_operands: the Dense layers of the GNCore layers round their matmul operands to
    bfloat16 (see _BF16_OPERANDS) - NOT the reference semantics, only an error model for the tensor-path tests."""
    global _BF16_OPERANDS
    cast = lambda a: None if a is None else np.asarray(a, dt)
    ef, nf, gf = cast(ef), cast(nf), cast(gf)
    for kind, p in layers:
        if kind == "block":
            ef, nf, gf = gnblock_sparse(p, g, ef, nf, gf, dt)
        else:
            _BF16_OPERANDS = bool(bf16_operands)
            try:
                ef, nf, gf = gncore(p, lambda bp, a, b, c: gnblock_sparse(bp, g, a, b, c, dt),
                                    ef, nf, gf, eps_mode)
            finally:
                _BF16_OPERANDS = False
    return ef, nf, gf


def forward_dense(layers, d: DenseBatch, ef, nf, gf, eps_mode=EPS_SQRT_VAR_EPS2):
    for kind, p in layers:
        if kind == "block":
            ef, nf, gf = gnblock_dense(p, d, ef, nf, gf)
        else:
            ef, nf, gf = gncore(p, lambda bp, a, b, c: gnblock_dense(bp, d, a, b, c),
                                ef, nf, gf, eps_mode)
    return ef, nf, gf


# --------------------------------------------------------------------------------------
# edge collapsing (src/gngraphbatch.jl:56-111) - "next" row
# --------------------------------------------------------------------------------------


def lower_tri_coords(PN):
    """getlowertriangularcoords: CartesianIndices in column-major order with i >= j."""
    return [(i, j) for j in range(PN) for i in range(PN) if i >= j]


def collapsef_dense(ef_padded, PN):
    """collapsef: ef (D,PE,B) x edge_collapser (PE, PN(PN+1)/2) / 2  (gngraphbatch.jl:69-85).
    A diagonal coordinate gets weight 2 (copy[i,j]+=1 twice), hence /2 gives the slot itself."""
    coords = lower_tri_coords(PN)
    B, PE, D = ef_padded.shape
    out = np.zeros((B, len(coords), D), ef_padded.dtype)
    for c, (i, j) in enumerate(coords):
        out[:, c, :] = (ef_padded[:, i + PN * j, :] + ef_padded[:, j + PN * i, :]) / 2
    return out


def collapsed_edge_idxs(padded):
    """getcollapsededgeidxs (gngraphbatch.jl:60-65): lower-triangular coords whose adjacency
    entry is one."""
    B, PN, _ = padded.shape
    coords = lower_tri_coords(PN)
    return [np.array([c for c, (i, j) in enumerate(coords) if padded[b, i, j] == 1], np.int64)
            for b in range(B)]


# --------------------------------------------------------------------------------------
# synthetic parameters: glorot-uniform weights, non-zero biases
# --------------------------------------------------------------------------------------


def _glorot(rng, out, inn):
    lim = np.sqrt(6.0 / (inn + out)) if inn + out > 0 else 0.0
    return rng.uniform(-lim, lim, size=(out, inn)).astype(np.float32)


def make_block_params(rng, din, dout, zero_bias=False):
    a, b, c = din
    p, q, r = dout
    bias = (lambda n: np.zeros(n, np.float32)) if zero_bias else \
        (lambda n: rng.uniform(-0.1, 0.1, n).astype(np.float32))
    return dict(din=tuple(din), dout=tuple(dout),
                We=_glorot(rng, p, a + 2 * b + c), be=bias(p),
                Wn=_glorot(rng, q, p + b + c), bn=bias(q),
                Wg=_glorot(rng, r, p + q + c), bg=bias(r))


def make_core_params(rng, dims, eps=1e-5):
    def ffn(d):
        return dict(W1=_glorot(rng, 4 * d, d), b1=rng.uniform(-0.1, 0.1, 4 * d).astype(np.float32),
                    W2=_glorot(rng, d, 4 * d), b2=rng.uniform(-0.1, 0.1, d).astype(np.float32))

    def ln(d):
        return dict(gamma=rng.uniform(0.5, 1.5, d).astype(np.float32),
                    beta=rng.uniform(-0.1, 0.1, d).astype(np.float32), eps=eps)
    return dict(dims=tuple(dims), block=make_block_params(rng, dims, dims),
                ffn=[ffn(d) for d in dims], ln1=[ln(d) for d in dims], ln2=[ln(d) for d in dims])


def rel_err(y, ref):
    """SURVEY 8c tolerance: max|y - ref| / max|ref| over the (active) entries given."""
    ref = np.asarray(ref, np.float64)
    y = np.asarray(y, np.float64)
    if ref.size == 0:
        return 0.0
    return float(np.max(np.abs(y - ref)) / max(np.max(np.abs(ref)), 1e-30))


def elementwise_err(y, ref):
    """SURVEY 8c second metric: worst element-wise |y - ref| / (|ref| + 1e-3 max|ref|) - relative error per element with
    a floor that keeps near-zero reference entries from dividing by nothing."""
    ref = np.asarray(ref, np.float64)
    y = np.asarray(y, np.float64)
    if ref.size == 0:
        return 0.0
    floor = 1e-3 * max(np.max(np.abs(ref)), 1e-30)
    return float(np.max(np.abs(y - ref) / (np.abs(ref) + floor)))


def rms_err(y, ref):
    """||y - ref||_2 / ||ref||_2: the average relative error (insensitive to a single unlucky element)."""
    ref = np.asarray(ref, np.float64)
    y = np.asarray(y, np.float64)
    if ref.size == 0:
        return 0.0
    return float(np.sqrt(np.sum((y - ref) ** 2)) / max(np.sqrt(np.sum(ref ** 2)), 1e-30))


def logit_cross_entropy(logits, targets):
    """Flux.logitcrossentropy(yhat, y) with the defaults dims=1, agg=mean, on compact [rows][D] arrays (the transposes of the
    (D, rows) matrices flatunpaddednf / flatunpaddedef return, examples/sort/sort.jl:76-78):
    mean_r( -sum_d y[r, d] * logsoftmax(yhat[r, :])[d] ), float64."""
    x = np.asarray(logits, np.float64)
    t = np.asarray(targets, np.float64)
    m = x.max(axis=1, keepdims=True)
    lse = m + np.log(np.exp(x - m).sum(axis=1, keepdims=True))
    return float(np.mean(-(t * (x - lse)).sum(axis=1))) if x.shape[0] else 0.0
