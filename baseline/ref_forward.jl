# Runs the UNMODIFIED reference (GraphNets.jl + Flux 0.14) on a golden fixture and prints the relative
# error of the committed oracle outputs against it - closes the "parity unpinned" loop of DESIGN.md §4
# for anyone who has Julia (the build image does not).
#
#   julia --project=/path/to/GraphNets.jl baseline/ref_forward.jl tests/golden/cfg1_readme_block.npz
#
# Needs NPZ.jl.  Only the single-GNBlock fixture layout (L0_block_*) is handled here.
using GraphNets, Flux, NPZ
z = npzread(ARGS[1])
dims = Int.(z["L0_block_dims"])
block = GNBlock(Tuple(dims[1:3]) => Tuple(dims[4:6]))
block.edgefn[1].weight .= z["L0_block_We"];  block.edgefn[1].bias .= z["L0_block_be"]
block.nodefn[1].weight .= z["L0_block_Wn"];  block.nodefn[1].bias .= z["L0_block_bn"]
block.graphfn[1].weight .= z["L0_block_Wg"]; block.graphfn[1].bias .= z["L0_block_bg"]
B = Int(z["n_graphs"])
adjs = [Int.(z["adj_$(b-1)"]) for b in 1:B]
ep = Int.(z["idx_graph_edge_ptr"]); np_ = Int.(z["idx_graph_node_ptr"])
ef = haskey(z, "ef") ? [permutedims(z["ef"][ep[b]+1:ep[b+1], :]) for b in 1:B] : nothing
nf = haskey(z, "nf") ? [permutedims(z["nf"][np_[b]+1:np_[b+1], :]) for b in 1:B] : nothing
x = (graphs=adjs, ef=ef, nf=nf, gf=nothing) |> batch
y = block(x)
relerr(a, b) = maximum(abs.(a .- b)) / maximum(abs.(b))
haskey(z, "ye") && println("ef rel err vs oracle: ", relerr(permutedims(flatunpaddedef(y)), z["ye"]))
haskey(z, "yn") && println("nf rel err vs oracle: ", relerr(permutedims(flatunpaddednf(y)), z["yn"]))
haskey(z, "yg") && println("gf rel err vs oracle: ", relerr(permutedims(reshape(y.gf, size(y.gf, 1), :)), z["yg"]))
