// Fused GNCore edge kernel (tcgen05, generation 5); internal interface.
#pragma once
#include <cuda.h>      // CUtensorMap (types only: the driver entry point is resolved at run time, no -lcuda)
#include <cuda_bf16.h>
#include "common.cuh"

constexpr int TC_BIAS_SLAB_BYTES = 4096;                  // 128 rows x 16 bf16 (one K = 16 step), no swizzle
constexpr int TC_BIAS_PACK_BYTES = 5 * TC_BIAS_SLAB_BYTES;

struct EdgeArgs {
  const float* x;     // [R][128] edge features
  float* y;           // [R][128] core output
  int64_t R;
  int num_tiles;
  const __nv_bfloat16* wpack;   // 9 packed 32 KB weight blocks (tc.cu::tc_core_pack, edge order)
  const __nv_bfloat16* bias_pack;   // 5 bias slabs of 4 KB (tc.cu::k_pack_bias): b1' chunks 0-3 (LN2 shift folded in), b2
  float eps;
  int eps_mode;
  // gathered addend rows: g = add1[idx1[r]] + add2[idx2[r]]   (sender projection, receiver projection + per-graph row);
  // add_bf16: both are bf16 matrices (the edges' P_s | P_r'), else fp32 (the nodes' P_agg, P_un); ld in elements
  const void* add1; const int32_t* idx1; int ld1;   // idx1 == nullptr: the row itself
  const void* add2; const int32_t* idx2; int ld2;
  int add_bf16;
  // optional L2 prefetch of the rows the OUT warps will gather (edges): graph of every row + the graphs' row ranges in add1
  // (graph_node_ptr); add1 / add2 are then the two halves of ONE row-interleaved matrix (P_s | P_r'), one range covers both
  const int32_t* pf_row_graph; const int32_t* pf_graph_ptr;
  const int32_t* part;   // partial-row id per edge (32-row blocks, receiver runs)
  float* Epart;          // out [n_parts][128] partial sums of the normalised edge rows
  float* Gpart;          // out [n_parts][128] partial sums of the gathered addends
  int part_bf16;         // both stored as bf16 rows (256 B) instead of fp32
  unsigned long long* dbg;
  WatchArgs wd;          // kernel watchdog (tc_ptx.cuh)
  // Fused narrow decoder (last core of a model followed by a GNBlock with out_e <= 4, src/gnblock.jl:65): instead of storing y
  // the kernel stores dec_out[r][0..4) = y[r] . decW  (the edge-feature rows of the decoder's edge Dense, fp32, padded to 4
  // outputs) - y itself is consumed by nothing else, so 2 x R x 512 B of HBM traffic and one full streaming pass disappear.
  const float* decW;     // [128][4] k-major, nullptr: store y as usual
  float* dec_out;        // [R][4]
  // y as a 2-D tensor map (launch_edge5 fills it): the OUT warps write the finished rows back into the staging tile and one lane
  // stores a 16-row slice with four cp.async.bulk.tensor (TMA) stores - 8 KB bursts instead of 512 B per-row stores
  int y_tma;
  alignas(64) CUtensorMap ymap;
};

int launch_edge5(gnb_ctx* ctx, const EdgeArgs& a, const char* name, double flops, double bytes);
