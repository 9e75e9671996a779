# GraphNetsB200.jl - drop-in for the GNBlock / GNCore forward path of GraphNets.jl over libgnb200.so.
#
# NOT EXECUTED IN THE BUILD ENVIRONMENT (no Julia toolchain there); kept 1:1 with include/gnb200.h so
# every behaviour is testable through the Python binding.  It keeps the reference's exported names
# (src/GraphNets.jl:12-50): batch, unbatch, GNGraphBatch, GNBlock, GNCore, GNCoreList, efview, nfview,
# gfview, flatunpaddedef, flatunpaddednf.  Feature arrays are CuArray{Float32}; Julia's column-major
# (D, T) layout is exactly the ABI's compact [T][D] layout, so no copies are made at the boundary.
module GraphNetsB200

using CUDA

const LIB = get(ENV, "GNB200_LIB", joinpath(@__DIR__, "..", "graphnets.jl_b200", "libgnb200.so"))

struct GnbError <: Exception
    code::Cint
    msg::String
end

function check(rc::Cint)
    rc == 0 && return
    msg = unsafe_string(ccall((:gnb_last_error, LIB), Cstring, ()))
    rc == -1 && throw(AssertionError(msg))          # the reference uses @assert (src/checks.jl)
    rc == -3 && throw(OutOfMemoryError())
    throw(GnbError(rc, msg))
end

# ---------------------------------------------------------------- context
mutable struct Ctx
    ptr::Ptr{Cvoid}
end
const CTX = Ref{Union{Nothing,Ctx}}(nothing)
function ctx()
    if CTX[] === nothing
        err = Ref{Cint}(0)
        p = ccall((:gnb_ctx_create, LIB), Ptr{Cvoid}, (Cint, Ptr{Cint}), CUDA.deviceid(CUDA.device()), err)
        p == C_NULL && check(err[])
        CTX[] = Ctx(p)
    end
    c = CTX[]
    check(ccall((:gnb_ctx_set_stream, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), c.ptr, CUDA.stream().handle))
    c.ptr
end

# ---------------------------------------------------------------- parameter structs (include/gnb200.h)
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

# ---------------------------------------------------------------- GNGraphBatch (src/gngraphbatch.jl:1-54)
mutable struct GNGraphBatch
    adj_mats
    handle::Ptr{Cvoid}
    B::Int; E::Int; N::Int
    node_block_size::Int
    edge_block_size::Int
    n_nodes::Vector{Int32}
end

function GNGraphBatch(adj_mats::AbstractVector; B::Int=length(adj_mats))
    @assert length(adj_mats) > 0
    PN = maximum(size.(adj_mats, 1))
    Badj = length(adj_mats)
    mask = zeros(UInt8, PN, PN, Badj)                      # padadjmats (src/pad.jl:1-10) as an isone mask
    for (b, a) in enumerate(adj_mats)
        n = size(a, 1)
        mask[1:n, 1:n, b] .= isone.(a)
    end
    nn = Int32.(size.(adj_mats, 1))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:gnb_graph_lower, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Ptr{Int32}, Cint, Cint, Cint, Ptr{Ptr{Cvoid}}),
                ctx(), mask, 1, 0, nn, PN, Badj, B, h))
    E = Ref{Int64}(0); N = Ref{Int64}(0)
    check(ccall((:gnb_graph_counts, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int32}, Ptr{Int32}),
                h[], E, N, C_NULL, C_NULL))
    g = GNGraphBatch(adj_mats, h[], B, E[], N[], PN, PN^2, nn)
    finalizer(x -> ccall((:gnb_graph_destroy, LIB), Cint, (Ptr{Cvoid},), x.handle), g)
    g
end

# ---------------------------------------------------------------- batch / unbatch (src/batch.jl, src/unbatch.jl)
# Batched features are COMPACT CuMatrix (D, E) / (D, N) / (D, B): `flatunpaddedef` / `flatunpaddednf` order.
function batch(t::NamedTuple)
    @assert Set(keys(t)) == Set((:graphs, :ef, :nf, :gf))
    (; graphs, ef, nf, gf) = t
    @assert !isnothing(ef) || !isnothing(nf) || !isnothing(gf)
    if graphs isa AbstractMatrix                         # one structure shared by the batch
        B = !isnothing(ef) ? size(ef, 3) : (!isnothing(nf) ? size(nf, 3) : size(gf, 2))
        !isnothing(ef) && @assert size(ef, 2) == count(isone, graphs)
        !isnothing(nf) && @assert size(nf, 2) == size(graphs, 1)
        g = GNGraphBatch([graphs]; B=B)
        flat(x) = isnothing(x) ? nothing : CuArray{Float32}(reshape(x, size(x, 1), :))
        return (graphs=g, ef=flat(ef), nf=flat(nf), gf=flat(gf))
    end
    g = GNGraphBatch(collect(graphs))
    cat2(xs) = isnothing(xs) ? nothing : CuArray{Float32}(reduce(hcat, xs))
    (graphs=g, ef=cat2(ef), nf=cat2(nf), gf=cat2(gf))
end

flatunpaddedef(t) = t.ef
flatunpaddednf(t) = t.nf

# ---------------------------------------------------------------- layers
devptr(x) = x === nothing ? CU_NULL : pointer(x)

mutable struct GNBlock
    dims::Pair
    We; be; Wn; bn; Wg; bg            # CuArrays, Flux layout (out, in)
end
glorot(out, inn) = (rand(Float32, out, inn) .- 0.5f0) .* 2f0 .* sqrt(6f0 / (inn + out))
function GNBlock(p::Pair; dropout=0)
    (ei, ni, gi), (eo, no, go) = p
    @assert any((ei, ni, gi) .> 0) && any((eo, no, go) .> 0)
    GNBlock(p, CuArray(glorot(eo, ei + 2ni + gi)), CUDA.zeros(Float32, eo), CuArray(glorot(no, eo + ni + gi)),
            CUDA.zeros(Float32, no), CuArray(glorot(go, eo + no + gi)), CUDA.zeros(Float32, go))
end
params(m::GNBlock) = BlockParams(m.dims[1]..., m.dims[2]..., devptr(m.We), devptr(m.be), devptr(m.Wn),
                                 devptr(m.bn), devptr(m.Wg), devptr(m.bg))

function (m::GNBlock)(x)
    g = x.graphs
    eo, no, go = m.dims[2]
    oe = eo > 0 ? CUDA.zeros(Float32, eo, g.E) : nothing
    on = no > 0 ? CUDA.zeros(Float32, no, g.N) : nothing
    og = go > 0 ? CUDA.zeros(Float32, go, g.B) : nothing
    p = Ref(params(m))
    GC.@preserve m x p check(ccall((:gnb_block_forward, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{BlockParams}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32},
         CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Cint),
        ctx(), g.handle, p, devptr(x.ef), devptr(x.nf), devptr(x.gf), devptr(oe), devptr(on), devptr(og), 0))
    (graphs=g, ef=oe, nf=on, gf=og)                       # zerodim2nothing (src/gnblock.jl:71-78)
end

# GNCore / GNCoreList follow the same pattern with CoreParams and gnb_corelist_forward; a model that
# wants the bf16 tensor-core path builds one gnb_model (gnb_model_create) and calls gnb_model_forward.

export batch, GNGraphBatch, GNBlock, flatunpaddedef, flatunpaddednf
end
