// tcgen05 / TMA bf16 tensor-core path for GNCore layers (internal interface).
#pragma once
#include "common.cuh"

struct TcCorePack;  // packed bf16 weights + folded LayerNorm affine of one GNCore

// true when a GNCore with these (edge, node, graph) widths can run on the tcgen05 path
bool tc_core_supported(int de, int dn, int dg);
int tc_core_pack(gnb_ctx* ctx, const gnb_block_params& blk, const gnb_ffn_params* ffn, const gnb_ln_params* ln1,
                 const gnb_ln_params* ln2, TcCorePack** out);
void tc_core_pack_free(TcCorePack* p);
int tc_core_forward(gnb_ctx* ctx, const gnb_graph* g, const TcCorePack* pk, const gnb_block_params& blk,
                    const gnb_ffn_params* ffn, const gnb_ln_params* ln1, const gnb_ln_params* ln2,
                    const float* xe, const float* xn, const float* xg, float* ye, float* yn, float* yg);
