// Graph-level (B rows) stages of the 128-wide GNCore tensor path (internal interface).
#pragma once
#include "common.cuh"

struct GraphPreArgs {
  const float* xg;       // [B][128] graph features u
  int64_t B;
  const float *gamma, *beta;   // LN1 (graph)
  float eps;
  int eps_mode;
  const float* Weu;      // [128][128] rows [3H,4H) of the edge Dense  (k-major)
  const float* Wnu;      // [128][128] rows [2H,3H) of the node Dense
  const float *ce, *cn;  // [128] biases with the folded LayerNorm shifts of the tensor path
  float *Pue, *Pun;      // out [B][128]
};

struct GraphPostArgs {
  const float* xg;
  int64_t B;
  // edge -> graph sum of the block output (== sum of the node aggregates, src/graphfninput.jl:3), by linearity over the partial
  // rows of the aggregate kernel (tc.cu::k_tc_agg2):  s_e = (g1e . sum SE_part) W_ee + sum SG_part
  const float* SEpart;   // [n_nparts][128] partial sums of the nodes' summed normalised edge rows
  const float* SGpart;   // [n_nparts][128] partial sums of the nodes' summed edge addends
  const float* Wee;      // [128][128] rows [0,H) of the edge Dense (k-major), without the LayerNorm scale
  const float* g1e;      // [128] LN1 (edge) scale
  // node -> graph sum of the block output h_v = W_nv' v^ + addends, by linearity over the partial rows of the node kernel
  const int32_t* graph_npart_ptr;   // [B+1]
  const float* Vpart;    // [n_nparts][128] partial sums of the normalised node rows v^
  const float* Npart;    // [n_nparts][128] partial sums of the node addends
  const float* Wnv;      // [128][128] rows [H,2H) of the node Dense (k-major), without the LayerNorm scale
  const float* g1n;      // [128] LN1 (node) scale
  const float *g1, *b1ln; float eps1; int eps_mode1;   // LN1 (graph)
  const float *g2, *b2ln; float eps2; int eps_mode2;   // LN2 (graph)
  const float *Wg, *bg;  // graph Dense (3H -> H), k-major
  const float *W1, *b1, *W2, *b2;   // graph FFN
  float* yg;             // out [B][128]
  // optional: the per-graph rows of the NEXT core (k_graph_pre fused into this launch); next_Pue == nullptr: none
  const float *next_gamma, *next_beta; float next_eps; int next_eps_mode;
  const float *next_Weu, *next_Wnu, *next_ce, *next_cn;
  float *next_Pue, *next_Pun;
};

int launch_graph_pre(gnb_ctx* ctx, const GraphPreArgs& a);
int launch_graph_post(gnb_ctx* ctx, const GraphPostArgs& a, int64_t N);
