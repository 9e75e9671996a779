// tcgen05 / TMA bf16 tensor-core path for GNCore layers (internal interface).
#pragma once
#include "common.cuh"

struct TcCorePack;  // packed bf16 weights + folded LayerNorm affine of one GNCore

// true when a GNCore with these (edge, node, graph) widths can run on the tcgen05 path
bool tc_core_supported(int de, int dn, int dg);
int tc_core_pack(gnb_ctx* ctx, const gnb_block_params& blk, const gnb_ffn_params* ffn, const gnb_ln_params* ln1,
                 const gnb_ln_params* ln2, TcCorePack** out);
void tc_core_pack_free(TcCorePack* p);
// Per-graph rows (P_ue, P_un, [B][128] each) of a core.  A core's graph-level tail kernel can produce the rows of the
// NEXT core (same launch), so consecutive tensor-path cores chain them through `pre_in` / `pre_out`.
struct TcPreRows { float* Pue = nullptr; float* Pun = nullptr; };
struct TcNextCore {      // what the tail kernel needs to know about the next core (nullptr pack: no next tensor-path core)
  const TcCorePack* pk = nullptr;
  const gnb_block_params* blk = nullptr;
  const gnb_ln_params* ln1 = nullptr;
};
// Fused narrow decoder (model.cu): the edge kernel of the LAST core stores y_e . W4 instead of y_e (tc_edge.cuh, EdgeArgs::decW)
struct TcDecFuse { const float* W4 = nullptr; float* partial = nullptr; };
// pre_in.Pue != nullptr: the rows of this core were already produced by the previous core's tail kernel
int tc_core_forward(gnb_ctx* ctx, const gnb_graph* g, const TcCorePack* pk, const gnb_block_params& blk,
                    const gnb_ffn_params* ffn, const gnb_ln_params* ln1, const gnb_ln_params* ln2,
                    const float* xe, const float* xn, const float* xg, float* ye, float* yn, float* yg,
                    TcPreRows pre_in = TcPreRows(), TcNextCore next = TcNextCore(), TcPreRows pre_out = TcPreRows(),
                    TcDecFuse dec = TcDecFuse());
