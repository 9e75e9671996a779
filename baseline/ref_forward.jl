# Runs the UNMODIFIED reference (GraphNets.jl + Flux 0.14) on the committed golden fixtures and prints the relative error of
# the committed oracle outputs against it - one run closes the "parity unpinned" loop of DESIGN.md section 4 for anyone who
# has Julia (the build image does not: no julia binary, no network).
#
#   julia --project=/path/to/GraphNets.jl baseline/ref_forward.jl tests/golden/*.npz
#
# Needs NPZ.jl.  Handles every fixture layout tests/golden/make_golden.py writes: any sequence of
#   L<i>_block_{We,be,Wn,bn,Wg,bg,dims}                                   -> GNBlock
#   L<i>_core_{dims, blk_*, ffn<k>_{W1,b1,W2,b2}, ln1<k>_{gamma,beta}, ln2<k>_{gamma,beta}}   (k = 0 edge, 1 node, 2 graph) -> GNCore
# composed left to right (GNCoreList semantics, src/gncorelist.jl:43-45).  The three varbatch_eps<m> fixtures differ only in the
# LayerNorm denominator the oracle used (eps_mode 0: sqrt(var + eps^2), 1: std + eps, 2: sqrt(var + eps)): the one with the
# smallest error tells which convention the installed Flux implements (SURVEY Appendix D) - the product's default is mode 0.
using GraphNets, Flux, NPZ

relerr(a, b) = maximum(abs.(a .- b)) / max(maximum(abs.(b)), 1f-30)

function setdense!(d, W, b)
    d.weight .= W                       # fixtures hold numpy (out, in) matrices == Flux's Dense.weight
    d.bias .= vec(b)
end

function build_layers(z)
    layers = Any[]
    li = 0
    while true
        if haskey(z, "L$(li)_block_dims")
            pre = "L$(li)_block_"
            d = Int.(z[pre * "dims"])
            m = GNBlock(Tuple(d[1:3]) => Tuple(d[4:6]))
            setdense!(m.edgefn[1], z[pre * "We"], z[pre * "be"])
            setdense!(m.nodefn[1], z[pre * "Wn"], z[pre * "bn"])
            setdense!(m.graphfn[1], z[pre * "Wg"], z[pre * "bg"])
            push!(layers, m)
        elseif haskey(z, "L$(li)_core_dims")
            pre = "L$(li)_core_"
            dims = Tuple(Int.(z[pre * "dims"]))
            c = GNCore(dims)
            setdense!(c.block.edgefn[1], z[pre * "blk_We"], z[pre * "blk_be"])
            setdense!(c.block.nodefn[1], z[pre * "blk_Wn"], z[pre * "blk_bn"])
            setdense!(c.block.graphfn[1], z[pre * "blk_Wg"], z[pre * "blk_bg"])
            for (k, ch) in enumerate((c.ffwd.eff, c.ffwd.nff, c.ffwd.gff))
                setdense!(ch[1], z[pre * "ffn$(k-1)_W1"], z[pre * "ffn$(k-1)_b1"])
                setdense!(ch[2], z[pre * "ffn$(k-1)_W2"], z[pre * "ffn$(k-1)_b2"])
            end
            for (name, gn) in (("ln1", c.gn1), ("ln2", c.gn2))
                for (k, ln) in enumerate((gn.edgeln, gn.nodeln, gn.graphln))
                    ln.diag.scale .= vec(z[pre * "$(name)$(k-1)_gamma"])
                    ln.diag.bias .= vec(z[pre * "$(name)$(k-1)_beta"])
                end
            end
            push!(layers, c)
        else
            break
        end
        li += 1
    end
    layers
end

function run_fixture(path)
    z = npzread(path)
    layers = build_layers(z)
    B = Int(z["n_graphs"])
    adjs = [Int.(z["adj_$(b-1)"]) for b in 1:B]
    ep = Int.(z["idx_graph_edge_ptr"]); np_ = Int.(z["idx_graph_node_ptr"])
    # fixtures store compact [rows][D] arrays (C order); the reference wants (D, rows) per graph
    ef = haskey(z, "ef") ? [permutedims(z["ef"][ep[b]+1:ep[b+1], :]) for b in 1:B] : nothing
    nf = haskey(z, "nf") ? [permutedims(z["nf"][np_[b]+1:np_[b+1], :]) for b in 1:B] : nothing
    gf = haskey(z, "gf") ? [vec(z["gf"][b, :]) for b in 1:B] : nothing
    x = (graphs=adjs, ef=ef, nf=nf, gf=gf) |> batch
    y = foldl((h, l) -> l(h), layers; init=x)
    println(basename(path), "  (", length(layers), " layers, eps_mode of the oracle = ", Int(z["eps_mode"]), ")")
    haskey(z, "ye") && println("   ef rel err, oracle vs reference: ", relerr(z["ye"], permutedims(flatunpaddedef(y))))
    haskey(z, "yn") && println("   nf rel err, oracle vs reference: ", relerr(z["yn"], permutedims(flatunpaddednf(y))))
    haskey(z, "yg") && println("   gf rel err, oracle vs reference: ", relerr(z["yg"], permutedims(reshape(y.gf, size(y.gf, 1), :))))
    # the structure goldens: the reference's own compact order must reproduce the oracle's edge index
    padded = GraphNets.padadjmats(adjs)
    for b in 1:B
        k = findall(isone, view(padded, :, :, b)[:]) .- 1          # 0-based padded slots, ascending
        @assert k == Int.(z["idx_edge_slot"][ep[b]+1:ep[b+1]]) "edge order differs in graph $b"
    end
    println("   edge order (findall(isone, adj[:])) == oracle edge_slot: ok")
end

isempty(ARGS) && error("usage: julia --project=<GraphNets.jl> baseline/ref_forward.jl tests/golden/*.npz")
foreach(run_fixture, ARGS)
