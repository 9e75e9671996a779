# GraphNetsB200.jl - drop-in for the GNBlock / GNCore forward path of GraphNets.jl over libgnb200.so.
#
# NOT EXECUTED IN THE BUILD ENVIRONMENT (no Julia toolchain there).  Every `ccall` below binds a symbol of
# include/gnb200.h 1:1 (tests/test_host.py parses this file and checks that every symbol it calls is declared in the
# header and exported by the library, and that every name of the reference's export list is defined and exported here),
# and the same entry points are exercised through the Python ctypes host by the -m gpu tests.
#
# Exports = the reference's (src/GraphNets.jl:12-50) for this path:
#   GNGraphBatch, batch, unbatch, GNBlock, zerodim2nothing, GNCore, GNCoreList, efview, nfview, gfview,
#   flatunpaddednf, flatunpaddedef, collapsef, unpaddedcollapsedef, flatunpaddedcollapsedef
# plus `batch_coo` (edge-list input), `compile` / `GNModel` (one engine call for a whole model; the only way onto the
# bf16 tensor-core path) and `set_precision!`.
#
# `const GraphNets = GraphNetsB200` (or `import GraphNetsB200 as GraphNets`) switches user code over.
#
# Data layout: Julia's column-major (D, T) IS the ABI's compact [T][D] layout, so CuArrays cross the boundary without
# copies.  Batched features are `Padded` arrays: they present the documented padded face `size(x.ef) == (DE, PN^2, B)`
# (src/batch.jl:44-50) but store only the compact (D, E) / (D, N) / (D, B) matrix on the device; the padded tensor is
# materialised (gnb_pad_edges / gnb_pad_nodes) only if somebody indexes it.
module GraphNetsB200

using CUDA
using Flux: Flux, Dense, Chain, Dropout, LayerNorm, relu

const LIB = get(ENV, "GNB200_LIB", joinpath(@__DIR__, "..", "graphnets.jl_b200", "libgnb200.so"))

# precision codes of include/gnb200.h
const PREC = Dict(:fp32 => Cint(0), :bf16 => Cint(2), :auto => Cint(3))
const DEFAULT_PRECISION = Ref(:fp32)          # the reference computes in Float32
set_precision!(p::Symbol) = (@assert haskey(PREC, p); DEFAULT_PRECISION[] = p)

struct GnbError <: Exception
    code::Cint
    msg::String
end

function check(rc::Cint)
    rc == 0 && return
    msg = unsafe_string(ccall((:gnb_last_error, LIB), Cstring, ()))
    rc == -1 && throw(AssertionError(msg))          # the reference uses @assert (src/checks.jl)
    rc == -3 && throw(OutOfMemoryError())
    throw(GnbError(rc, msg))                        # -2 CUDA, -4 unsupported, -5 kernel watchdog
end

# ---------------------------------------------------------------- context (one per device, bound to the task's stream)
mutable struct Ctx
    ptr::Ptr{Cvoid}
end
const CTXS = Dict{Int,Ctx}()
function ctx()
    dev = CUDA.deviceid(CUDA.device())
    c = get!(CTXS, dev) do
        err = Ref{Cint}(0)
        p = ccall((:gnb_ctx_create, LIB), Ptr{Cvoid}, (Cint, Ptr{Cint}), dev, err)
        p == C_NULL && check(err[] == 0 ? Cint(-2) : err[])
        Ctx(p)
    end
    check(ccall((:gnb_ctx_set_stream, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), c.ptr, CUDA.stream().handle))
    c.ptr
end
sync() = check(ccall((:gnb_sync, LIB), Cint, (Ptr{Cvoid},), ctx()))

# ---------------------------------------------------------------- parameter structs (include/gnb200.h, same field order)
struct BlockParams
    in_e::Int32; in_n::Int32; in_g::Int32; out_e::Int32; out_n::Int32; out_g::Int32
    We::CuPtr{Float32}; be::CuPtr{Float32}; Wn::CuPtr{Float32}; bn::CuPtr{Float32}
    Wg::CuPtr{Float32}; bg::CuPtr{Float32}
end
struct FfnParams
    W1::CuPtr{Float32}; b1::CuPtr{Float32}; W2::CuPtr{Float32}; b2::CuPtr{Float32}
end
struct LnParams
    gamma::CuPtr{Float32}; beta::CuPtr{Float32}; eps::Float32; eps_mode::Int32
end
struct CoreParams
    block::BlockParams
    ffn::NTuple{3,FfnParams}
    ln1::NTuple{3,LnParams}
    ln2::NTuple{3,LnParams}
end
struct Layer                      # gnb_layer: { int32 kind; int32 _pad; gnb_block_params block; gnb_core_params core; }
    kind::Int32
    _pad::Int32
    block::BlockParams
    core::CoreParams
end
const LAYER_BLOCK = Int32(0)
const LAYER_CORE = Int32(1)

devptr(::Nothing) = CU_NULL
devptr(x) = length(x) == 0 ? CU_NULL : pointer(x)
const NULL_BLOCK = BlockParams(0, 0, 0, 0, 0, 0, CU_NULL, CU_NULL, CU_NULL, CU_NULL, CU_NULL, CU_NULL)
const NULL_FFN = FfnParams(CU_NULL, CU_NULL, CU_NULL, CU_NULL)
const NULL_LN = LnParams(CU_NULL, CU_NULL, 0f0, 0)
const NULL_CORE = CoreParams(NULL_BLOCK, (NULL_FFN, NULL_FFN, NULL_FFN), (NULL_LN, NULL_LN, NULL_LN), (NULL_LN, NULL_LN, NULL_LN))

# ---------------------------------------------------------------- GNGraphBatch (src/gngraphbatch.jl:1-54)
# Instead of seven dense broadcaster tensors: a device-resident receiver-sorted COO + CSR behind `handle`.
mutable struct GNGraphBatch
    adj_mats                    # Vector of adjacency matrices (length 1: one structure shared by the batch)
    handle::Ptr{Cvoid}
    B::Int; E::Int; N::Int
    node_block_size::Int        # PN
    edge_block_size::Int        # PN^2
    n_nodes::Vector{Int32}
    graph_edge_ptr::Vector{Int32}     # host copies of the per-graph offsets (views / unbatch)
    graph_node_ptr::Vector{Int32}
end

function finish_graph(adj_mats, h::Ptr{Cvoid}, B::Int, PN::Int, nn::Vector{Int32})
    E = Ref{Int64}(0); N = Ref{Int64}(0)
    check(ccall((:gnb_graph_counts, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int32}, Ptr{Int32}),
                h, E, N, C_NULL, C_NULL))
    ep = Vector{Int32}(undef, B + 1); np_ = Vector{Int32}(undef, B + 1)
    check(ccall((:gnb_graph_export_host, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}),
                ctx(), h, C_NULL, C_NULL, C_NULL, C_NULL, ep, np_, C_NULL))
    g = GNGraphBatch(adj_mats, h, B, E[], N[], PN, PN^2, nn, ep, np_)
    finalizer(x -> ccall((:gnb_graph_destroy, LIB), Cint, (Ptr{Cvoid},), x.handle), g)
    g
end

function GNGraphBatch(adj_mats::AbstractVector; B::Int=length(adj_mats))
    @assert length(adj_mats) > 0
    PN = maximum(size.(adj_mats, 1))
    Badj = length(adj_mats)
    @assert Badj == 1 || Badj == B
    mask = zeros(UInt8, PN, PN, Badj)                      # padadjmats (src/pad.jl:1-10) as an `isone` mask
    for (b, a) in enumerate(adj_mats)
        n = size(a, 1)
        @assert size(a, 2) == n
        mask[1:n, 1:n, b] .= isone.(a)
    end
    nn = Int32.(size.(adj_mats, 1))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:gnb_graph_lower, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Ptr{Int32}, Cint, Cint, Cint, Ptr{Ptr{Cvoid}}),
                ctx(), mask, 1 #= GNB_ADJ_U8 =#, 0, nn, PN, Badj, B, h))
    finish_graph(adj_mats, h[], B, PN, Badj == 1 ? fill(nn[1], B) : nn)
end

# Edge-list input (gnb_graph_from_coo): 1-based local ids like everything in Julia; each graph's edges ascending in
# (receiver, sender) = the order of findall(isone, adj[:]) (src/pad.jl:30).
function GNGraphBatch(src::Vector{<:Integer}, dst::Vector{<:Integer}, graph_edge_ptr::Vector{<:Integer}, n_nodes::Vector{<:Integer})
    B = length(n_nodes)
    @assert length(graph_edge_ptr) == B + 1 && graph_edge_ptr[1] == 0
    s0 = Int32.(src .- 1); d0 = Int32.(dst .- 1); ep = Int32.(graph_edge_ptr); nn = Int32.(n_nodes)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:gnb_graph_from_coo, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}, Cint, Ptr{Int32}, Ptr{Int32}, Cint, Cint, Ptr{Ptr{Cvoid}}),
                ctx(), s0, d0, 0, ep, nn, 0, B, h))
    PN = max(Int(maximum(nn)), 1)
    adjs = map(1:B) do b                                    # the reference-facing field, small host matrices
        a = zeros(Int, nn[b], nn[b])
        for e in ep[b]+1:ep[b+1]
            a[src[e], dst[e]] = 1
        end
        a
    end
    finish_graph(adjs, h[], B, PN, nn)
end

# ---------------------------------------------------------------- Padded: the (D, T, B) face over compact storage
mutable struct Padded{K} <: AbstractArray{Float32,3}       # K = :e | :n | :g
    compact::CuMatrix{Float32}                             # (D, E) | (D, N) | (D, B)
    graphs::GNGraphBatch
    padded::Union{Nothing,CuArray{Float32,3}}
end
Padded{K}(c::CuMatrix{Float32}, g::GNGraphBatch) where {K} = Padded{K}(c, g, nothing)
Base.size(x::Padded{:e}) = (size(x.compact, 1), x.graphs.edge_block_size, x.graphs.B)
Base.size(x::Padded{:n}) = (size(x.compact, 1), x.graphs.node_block_size, x.graphs.B)
Base.size(x::Padded{:g}) = (size(x.compact, 1), 1, x.graphs.B)
uniform_nodes(g::GNGraphBatch) = all(==(g.node_block_size), g.n_nodes)

"Materialise (and cache) the padded tensor; inactive slots are zero."
function padded(x::Padded{K}) where {K}
    x.padded === nothing || return x.padded
    D, T, B = size(x)
    g = x.graphs
    if K === :g || (K === :n && uniform_nodes(g)) || (K === :e && g.E == T * B)
        x.padded = reshape(x.compact, D, T, B)             # same bytes
    else
        out = CUDA.zeros(Float32, D, T, B)
        fn = K === :e ? :gnb_pad_edges : :gnb_pad_nodes
        if K === :e
            check(ccall((:gnb_pad_edges, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, CuPtr{Float32}, Cint, CuPtr{Float32}),
                        ctx(), g.handle, pointer(x.compact), D, pointer(out)))
        else
            check(ccall((:gnb_pad_nodes, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, CuPtr{Float32}, Cint, CuPtr{Float32}),
                        ctx(), g.handle, pointer(x.compact), D, pointer(out)))
        end
        x.padded = out
    end
    x.padded
end
Base.getindex(x::Padded, I...) = getindex(padded(x), I...)
Base.view(x::Padded, I...) = view(padded(x), I...)
Base.Array(x::Padded) = Array(padded(x))
compact(::Nothing) = nothing
compact(x::Padded) = x.compact
compact(x::CuMatrix{Float32}) = x
wrap(K, ::Nothing, g) = nothing
wrap(K, c, g) = Padded{K}(c, g)

# ---------------------------------------------------------------- checks (src/checks.jl)
function checks(graphs::AbstractMatrix, ef, nf, gf)
    isnothing(ef) || @assert ndims(ef) == 3
    isnothing(nf) || @assert ndims(nf) == 3
    isnothing(gf) || @assert ndims(gf) == 2
    bs = filter(!isnothing, (isnothing(ef) ? nothing : size(ef, 3), isnothing(nf) ? nothing : size(nf, 3),
                             isnothing(gf) ? nothing : size(gf, 2)))
    @assert all(==(first(bs)), bs)
    isnothing(ef) || @assert size(ef, 2) == count(isone, graphs)
    isnothing(nf) || @assert size(nf, 2) == size(graphs, 1)
end
function checks(graphs::AbstractVector, ef, nf, gf)
    @assert length(graphs) > 0
    for x in (ef, nf, gf)
        isnothing(x) || @assert length(x) == length(graphs)
    end
    for i in eachindex(graphs)
        isnothing(ef) || (@assert ndims(ef[i]) == 2; @assert size(ef[i], 2) == count(isone, graphs[i]))
        isnothing(nf) || (@assert ndims(nf[i]) == 2; @assert size(nf[i], 2) == size(graphs[i], 1))
        isnothing(gf) || @assert ndims(gf[i]) == 1
    end
end

# ---------------------------------------------------------------- batch / unbatch (src/batch.jl:53-76, src/unbatch.jl:6-48)
function batch(t::NamedTuple)
    @assert Set(keys(t)) == Set((:graphs, :ef, :nf, :gf))
    (; graphs, ef, nf, gf) = t
    @assert !isnothing(ef) || !isnothing(nf) || !isnothing(gf)
    checks(graphs, ef, nf, gf)
    if graphs isa AbstractMatrix                         # one structure shared by the batch (src/batch.jl:66-76)
        B = !isnothing(ef) ? size(ef, 3) : (!isnothing(nf) ? size(nf, 3) : size(gf, 2))
        g = GNGraphBatch([graphs]; B=B)
        flat(x) = isnothing(x) ? nothing : CuArray{Float32}(reshape(x, size(x, 1), :))
        return (graphs=g, ef=wrap(:e, flat(ef), g), nf=wrap(:n, flat(nf), g), gf=wrap(:g, flat(gf), g))
    end
    g = GNGraphBatch(collect(graphs))
    cat2(xs) = isnothing(xs) ? nothing : CuArray{Float32}(reduce(hcat, xs))
    (graphs=g, ef=wrap(:e, cat2(ef), g), nf=wrap(:n, cat2(nf), g), gf=wrap(:g, cat2(gf), g))
end

"`batch` for graphs given as edge lists and compact features (D, E) / (D, N) / (D, B) in the same order."
function batch_coo(src, dst, graph_edge_ptr, n_nodes; ef=nothing, nf=nothing, gf=nothing)
    g = GNGraphBatch(collect(src), collect(dst), collect(graph_edge_ptr), collect(n_nodes))
    mv(x) = isnothing(x) ? nothing : CuArray{Float32}(x)
    (graphs=g, ef=wrap(:e, mv(ef), g), nf=wrap(:n, mv(nf), g), gf=wrap(:g, mv(gf), g))
end

function unbatch(t::NamedTuple)
    @assert Set(keys(t)) == Set((:graphs, :ef, :nf, :gf))
    (; graphs, ef, nf, gf) = t
    @assert !isnothing(ef) || !isnothing(nf) || !isnothing(gf)
    g = graphs
    ce, cn, cg = compact(ef), compact(nf), compact(gf)
    if length(g.adj_mats) == 1                             # src/unbatch.jl:13-17: a 1-graph batch unbatches to the single form
        m = g.B == 0 ? 0 : g.E ÷ g.B
        n = g.B == 0 ? 0 : g.N ÷ g.B
        return (graphs=g.adj_mats[1],
                ef=isnothing(ce) ? nothing : reshape(ce, size(ce, 1), m, g.B),     # aliases the batched storage
                nf=isnothing(cn) ? nothing : reshape(cn, size(cn, 1), n, g.B),
                gf=cg)
    end
    ep, np_ = g.graph_edge_ptr, g.graph_node_ptr
    (graphs=g.adj_mats,
     ef=isnothing(ce) ? nothing : [view(ce, :, ep[b]+1:ep[b+1]) for b in 1:g.B],
     nf=isnothing(cn) ? nothing : [view(cn, :, np_[b]+1:np_[b+1]) for b in 1:g.B],
     gf=isnothing(cg) ? nothing : [view(cg, :, b) for b in 1:g.B])
end

# ---------------------------------------------------------------- views (src/views.jl) - all alias the compact storage
function efview(t::NamedTuple, d1, d2, d3)
    @assert issubset(Set((:graphs, :ef)), Set(keys(t)))
    isnothing(t.ef) && return nothing
    g, ce = t.graphs, compact(t.ef)
    if length(g.adj_mats) == 1
        return view(reshape(ce, size(ce, 1), g.E ÷ g.B, g.B), d1, d2, d3)
    end
    d3 isa Integer || throw(MethodError(efview, (t, d1, d2, d3)))      # src/views.jl:26
    view(view(ce, :, g.graph_edge_ptr[d3]+1:g.graph_edge_ptr[d3+1]), d1, d2)
end
function nfview(t::NamedTuple, d1, d2, d3)
    @assert issubset(Set((:graphs, :nf)), Set(keys(t)))
    isnothing(t.nf) && return nothing
    g, cn = t.graphs, compact(t.nf)
    if length(g.adj_mats) == 1
        return view(reshape(cn, size(cn, 1), g.N ÷ g.B, g.B), d1, d2, d3)
    end
    d3 isa Integer || throw(MethodError(nfview, (t, d1, d2, d3)))      # src/views.jl:57
    view(view(cn, :, g.graph_node_ptr[d3]+1:g.graph_node_ptr[d3+1]), d1, d2)
end
function gfview(t::NamedTuple, d1, d2)
    @assert issubset(Set((:graphs, :gf)), Set(keys(t)))
    isnothing(t.gf) && return nothing
    view(compact(t.gf), d1, d2)
end
flatunpaddednf(t::NamedTuple) = compact(t.nf)              # (DN, N): free in the compact layout (src/views.jl:80-88)
flatunpaddedef(t::NamedTuple) = compact(t.ef)              # (DE, E)                              (src/views.jl:90-98)

# ---------------------------------------------------------------- loss over the compact views (examples/sort/sort.jl:76-78)
"`Flux.logitcrossentropy(flatunpaddednf(ŷ), flatunpaddednf(targets))` on the device, straight from the compact matrices."
function logitcrossentropy(yhat, y)
    a, b = compact(yhat), compact(y)
    @assert size(a) == size(b)
    loss = CUDA.zeros(Float32, 1)
    check(ccall((:gnb_logit_cross_entropy, LIB), Cint,
                (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, Cint, Int64, CuPtr{Float32}, CuPtr{Float32}),
                ctx(), pointer(a), pointer(b), size(a, 1), size(a, 2), pointer(loss), CU_NULL))
    loss
end

# gradient of the loss above w.r.t. the logits (the cotangent a backward pass starts from)
function logitcrossentropy_grad(yhat, y; scale::Real=1)
    a, b = compact(yhat), compact(y)
    d = similar(a)
    check(ccall((:gnb_logit_cross_entropy_bwd, LIB), Cint,
                (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, Cint, Int64, Cfloat, CuPtr{Float32}),
                ctx(), pointer(a), pointer(b), size(a, 1), size(a, 2), Float32(scale), pointer(d)))
    d
end

# ---------------------------------------------------------------- training step (examples/sort/sort.jl:118-132), fp32
# Thin wrappers of the backward operators (include/gnb200.h, "training step"): an `rrule` per layer composes them exactly as
# graphnets.jl_b200/train.py::Trainer does (INTEGRATION.md).  All arrays are compact (D, rows) CuMatrices.
struct LinSrc;  x::CuPtr{Float32}; d::Cint; ldx::Cint; W::CuPtr{Float32}; gamma::CuPtr{Float32}; beta::CuPtr{Float32}; eps::Cfloat; eps_mode::Cint; end
struct LinAdd;  a::CuPtr{Float32}; idx::CuPtr{Int32}; lda::Cint; end
struct LinArgs
    R::Int64; Nout::Cint; ldw::Cint; nsrc::Cint
    src::NTuple{3,LinSrc}
    bias::CuPtr{Float32}
    nadd::Cint; add::NTuple{4,LinAdd}
    relu::Cint
    out::CuPtr{Float32}; ldo::Cint
    precision::Cint
end
op_linear(a::LinArgs) = check(ccall((:gnb_op_linear, LIB), Cint, (Ptr{Cvoid}, Ref{LinArgs}), ctx(), Ref(a)))
op_segsum!(out, x, ptr, S; perm=CU_NULL) = check(ccall((:gnb_op_segsum, LIB), Cint,
    (Ptr{Cvoid}, CuPtr{Float32}, Cint, CuPtr{Int32}, Int64, CuPtr{Int32}, CuPtr{Float32}), ctx(), pointer(x), size(x, 1), pointer(ptr), S, perm, pointer(out)))
op_layernorm!(y, x, gamma, beta, eps, mode) = check(ccall((:gnb_op_layernorm, LIB), Cint,
    (Ptr{Cvoid}, CuPtr{Float32}, Int64, Cint, CuPtr{Float32}, CuPtr{Float32}, Cfloat, Cint, CuPtr{Float32}),
    ctx(), pointer(x), size(x, 2), size(x, 1), pointer(gamma), pointer(beta), eps, mode, pointer(y)))
op_layernorm_bwd!(dx, gxhat, x, g, gamma, eps, mode) = check(ccall((:gnb_op_layernorm_bwd, LIB), Cint,
    (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, Int64, Cint, CuPtr{Float32}, Cfloat, Cint, CuPtr{Float32}, CuPtr{Float32}),
    ctx(), pointer(x), pointer(g), size(x, 2), size(x, 1), pointer(gamma), eps, mode, pointer(dx), pointer(gxhat)))
op_wgrad!(dW, ldw, X, dY; idx=CU_NULL, precision=0) = check(ccall((:gnb_op_wgrad, LIB), Cint,
    (Ptr{Cvoid}, CuPtr{Float32}, Cint, Cint, CuPtr{Int32}, CuPtr{Float32}, Cint, Cint, Int64, CuPtr{Float32}, Cint, Cint),
    ctx(), pointer(X), size(X, 1), size(X, 1), idx, pointer(dY), size(dY, 1), size(dY, 1), size(dY, 2), dW, ldw, precision))
op_colsum!(out, X) = check(ccall((:gnb_op_colsum, LIB), Cint, (Ptr{Cvoid}, CuPtr{Float32}, Cint, Cint, Int64, CuPtr{Float32}),
    ctx(), pointer(X), size(X, 1), size(X, 1), size(X, 2), out))
op_relu_mask!(t, h) = check(ccall((:gnb_op_relu_mask, LIB), Cint, (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, Int64), ctx(), pointer(t), pointer(h), length(t)))
op_gather_add!(out, a, b1, idx1, b2, idx2) = check(ccall((:gnb_op_gather_add, LIB), Cint,
    (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Int32}, CuPtr{Float32}, CuPtr{Int32}, Int64, Cint),
    ctx(), pointer(out), a, b1, idx1, b2, idx2, size(out, 2), size(out, 1)))
op_transpose!(out, W, rows, cols, ld) = check(ccall((:gnb_op_transpose, LIB), Cint, (Ptr{Cvoid}, CuPtr{Float32}, Cint, Cint, Cint, CuPtr{Float32}),
    ctx(), W, rows, cols, ld, pointer(out)))
op_adamw!(p, g, m, v, step; lr=1f-3, beta1=0.9f0, beta2=0.999f0, eps=1f-8, weight_decay=0f0) = check(ccall((:gnb_op_adamw, LIB), Cint,
    (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Int64, Cfloat, Cfloat, Cfloat, Cfloat, Cfloat, Cint),
    ctx(), pointer(p), pointer(g), pointer(m), pointer(v), length(p), lr, beta1, beta2, eps, weight_decay, step))

# ---------------------------------------------------------------- edge collapsing (src/gngraphbatch.jl:56-111)
function collapsef(t::NamedTuple)
    g = t.graphs
    ef = t.ef isa Padded ? padded(t.ef) : t.ef
    D, PN = size(ef, 1), g.node_block_size
    out = CUDA.zeros(Float32, D, PN * (PN + 1) ÷ 2, g.B)
    check(ccall((:gnb_collapse_edges, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, CuPtr{Float32}, Cint, CuPtr{Float32}),
                ctx(), g.handle, pointer(ef), D, pointer(out)))
    out
end
function collapsededgeidxs(a::AbstractMatrix, PN::Int)     # getcollapsededgeidxs (src/gngraphbatch.jl:60-65)
    idx = Int[]; c = 0
    for j in 1:PN, i in j:PN
        c += 1
        i <= size(a, 1) && j <= size(a, 1) && isone(a[i, j]) && push!(idx, c)
    end
    idx
end
function unpaddedcollapsedef(t::NamedTuple)
    g = t.graphs
    col = collapsef(t)
    [view(col, :, collapsededgeidxs(g.adj_mats[length(g.adj_mats) == 1 ? 1 : b], g.node_block_size), b) for b in 1:g.B]
end
flatunpaddedcollapsedef(t::NamedTuple) = reduce(hcat, unpaddedcollapsedef(t))

# ---------------------------------------------------------------- layers: the reference's structs and field names
# Parameters live in Flux's own Dense / LayerNorm containers (so `gpu`, `Flux.setup`, `trainmode!` keep working) and must be
# CuArrays when a layer is called.
struct GNBlock                                            # src/gnblock.jl:39-61
    edgefn; nodefn; graphfn; dropout
end
Flux.@functor GNBlock
function GNBlock(io::Pair; dropout=0)
    (ei, ni, gi), (eo, no, go) = io
    @assert any((ei, ni, gi) .> 0)
    @assert any((eo, no, go) .> 0)
    GNBlock(Chain(Dense(ei + 2ni + gi => eo)), Chain(Dense(ni + eo + gi => no)), Chain(Dense(no + eo + gi => go)), Dropout(dropout))
end
zerodim2nothing(x) = (isnothing(x) || size(x, 1) == 0) ? nothing : x      # src/gnblock.jl:71-78
function blockdims(m::GNBlock)
    We, Wn, Wg = m.edgefn[1].weight, m.nodefn[1].weight, m.graphfn[1].weight
    eo, no, go = size(We, 1), size(Wn, 1), size(Wg, 1)
    # in_n + in_g from the node Dense, in_e + 2 in_n + in_g from the edge Dense, in_g from the graph Dense
    gi = size(Wg, 2) - no - eo
    ni = size(Wn, 2) - eo - gi
    ei = size(We, 2) - 2ni - gi
    (ei, ni, gi), (eo, no, go)
end
function params(m::GNBlock)
    (ei, ni, gi), (eo, no, go) = blockdims(m)
    BlockParams(ei, ni, gi, eo, no, go, devptr(m.edgefn[1].weight), devptr(m.edgefn[1].bias), devptr(m.nodefn[1].weight),
                devptr(m.nodefn[1].bias), devptr(m.graphfn[1].weight), devptr(m.graphfn[1].bias))
end

struct GNFeedForward                                      # src/gnfeedforward.jl:17-31
    eff; nff; gff
end
Flux.@functor GNFeedForward
function GNFeedForward(dims; dropout=0)
    @assert all(dims .> 0)
    mk(d) = Chain(Dense(d => 4d, relu), Dense(4d => d), Dropout(dropout))
    GNFeedForward(mk(dims[1]), mk(dims[2]), mk(dims[3]))
end
struct GNGraphNorm                                        # src/gngraphnorm.jl:9-17
    edgeln; nodeln; graphln
end
Flux.@functor GNGraphNorm
function GNGraphNorm(dims)
    @assert all(dims .> 0)
    GNGraphNorm(LayerNorm(dims[1]), LayerNorm(dims[2]), LayerNorm(dims[3]))
end
struct GNCore                                             # src/gncore.jl:38-59
    block; ffwd; gn1; gn2
end
Flux.@functor GNCore
GNCore(dims; dropout=0) = GNCore(GNBlock(dims => dims; dropout), GNFeedForward(dims; dropout), GNGraphNorm(dims), GNGraphNorm(dims))
struct GNCoreList                                         # src/gncorelist.jl:29-45
    list
end
Flux.@functor GNCoreList

ffnparams(c::Chain) = FfnParams(devptr(c[1].weight), devptr(c[1].bias), devptr(c[2].weight), devptr(c[2].bias))
# Flux 0.14 LayerNorm: diag.scale / diag.bias, eps field `ϵ`; denominator sqrt(var + eps^2) = GNB_EPS_SQRT_VAR_EPS2 (0)
lnparams(l::LayerNorm) = LnParams(devptr(l.diag.scale), devptr(l.diag.bias), Float32(l.ϵ), 0)
params(m::GNCore) = CoreParams(params(m.block), (ffnparams(m.ffwd.eff), ffnparams(m.ffwd.nff), ffnparams(m.ffwd.gff)),
                               (lnparams(m.gn1.edgeln), lnparams(m.gn1.nodeln), lnparams(m.gn1.graphln)),
                               (lnparams(m.gn2.edgeln), lnparams(m.gn2.nodeln), lnparams(m.gn2.graphln)))
layerdesc(m::GNBlock) = Layer(LAYER_BLOCK, 0, params(m), NULL_CORE)
layerdesc(m::GNCore) = Layer(LAYER_CORE, 0, NULL_BLOCK, params(m))
layerdescs(m::Union{GNBlock,GNCore}) = [layerdesc(m)]
layerdescs(m::GNCoreList) = reduce(vcat, [layerdescs(l) for l in values(m.list)])
outdims(m::GNBlock) = blockdims(m)[2]
outdims(m::GNCore) = blockdims(m.block)[2]
outdims(m::GNCoreList) = outdims(last(collect(values(m.list))))

function outputs(g::GNGraphBatch, dims)
    eo, no, go = dims
    (eo > 0 ? CUDA.zeros(Float32, eo, g.E) : nothing, no > 0 ? CUDA.zeros(Float32, no, g.N) : nothing,
     go > 0 ? CUDA.zeros(Float32, go, g.B) : nothing)
end
result(g, oe, on, og) = (graphs=g, ef=wrap(:e, oe, g), nf=wrap(:n, on, g), gf=wrap(:g, og, g))   # same `graphs` object

# ---- single layers: weights used in place (fp32 path; (m::GNBlock)(x) src/gnblock.jl:63-69, (m::GNCore)(x) src/gncore.jl:56-59)
function (m::GNBlock)(x::NamedTuple)
    g = x.graphs
    oe, on, og = outputs(g, outdims(m))
    p = Ref(params(m))
    GC.@preserve m x p check(ccall((:gnb_block_forward, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{BlockParams}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32},
         CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Cint),
        ctx(), g.handle, p, devptr(compact(x.ef)), devptr(compact(x.nf)), devptr(compact(x.gf)),
        devptr(oe), devptr(on), devptr(og), PREC[:fp32]))
    result(g, oe, on, og)
end
function (m::GNCore)(x::NamedTuple)
    g = x.graphs
    oe, on, og = outputs(g, outdims(m))
    p = Ref(params(m))
    GC.@preserve m x p check(ccall((:gnb_core_forward, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{CoreParams}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32},
         CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Cint),
        ctx(), g.handle, p, devptr(compact(x.ef)), devptr(compact(x.nf)), devptr(compact(x.gf)),
        devptr(oe), devptr(on), devptr(og), PREC[:fp32]))
    result(g, oe, on, og)
end
function (m::GNCoreList)(x::NamedTuple)                    # left fold (src/gncorelist.jl:43-45), one engine call
    cores = collect(values(m.list))
    if all(c -> c isa GNCore, cores)
        g = x.graphs
        oe, on, og = outputs(g, outdims(m))
        ps = [params(c) for c in cores]
        GC.@preserve m x ps check(ccall((:gnb_corelist_forward, LIB), Cint,
            (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{CoreParams}, Cint, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32},
             CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Cint),
            ctx(), g.handle, ps, length(ps), devptr(compact(x.ef)), devptr(compact(x.nf)), devptr(compact(x.gf)),
            devptr(oe), devptr(on), devptr(og), PREC[:fp32]))
        return result(g, oe, on, og)
    end
    foldl((h, l) -> l(h), cores; init=x)
end

# ---- whole models: one gnb_model (weights copied and, for the tensor path, packed to bf16 - the analogue of `model |> gpu`)
mutable struct GNModel
    handle::Ptr{Cvoid}
    out_dims::NTuple{3,Int}
    precision::Symbol
end
"""
    compile(layers...; precision=:auto)

`decoder ∘ core_list ∘ encoder` (README.md:133-149) as ONE engine call.  `precision`: `:fp32` (1e-5 parity), `:auto` (tcgen05 bf16
tensor-core path for the GNCore shapes it supports, 1e-2 parity) or `:bf16` (like `:auto`, unsupported cores are an error).
Re-`compile` after changing parameters.
"""
function compile(layers...; precision::Symbol=DEFAULT_PRECISION[])
    descs = reduce(vcat, [layerdescs(l) for l in layers])
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve layers descs check(ccall((:gnb_model_create, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Layer}, Cint, Cint, Ptr{Ptr{Cvoid}}), ctx(), descs, length(descs), 1 #= weights on device =#, h))
    oe = Ref{Int32}(0); on = Ref{Int32}(0); og = Ref{Int32}(0)
    check(ccall((:gnb_model_out_dims, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}), h[], oe, on, og))
    m = GNModel(h[], (Int(oe[]), Int(on[]), Int(og[])), precision)
    finalizer(x -> ccall((:gnb_model_destroy, LIB), Cint, (Ptr{Cvoid},), x.handle), m)
    m
end
function (m::GNModel)(x::NamedTuple)
    g = x.graphs
    oe, on, og = outputs(g, m.out_dims)
    GC.@preserve m x check(ccall((:gnb_model_forward, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32},
         CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Cint),
        ctx(), m.handle, g.handle, devptr(compact(x.ef)), devptr(compact(x.nf)), devptr(compact(x.gf)),
        devptr(oe), devptr(on), devptr(og), PREC[m.precision]))
    result(g, oe, on, og)
end

export GNGraphBatch, batch, unbatch, GNBlock, zerodim2nothing, GNCore, GNCoreList, efview, nfview, gfview,
       flatunpaddednf, flatunpaddedef, collapsef, unpaddedcollapsedef, flatunpaddedcollapsedef,
       batch_coo, compile, GNModel, set_precision!, Padded, padded, logitcrossentropy
end
