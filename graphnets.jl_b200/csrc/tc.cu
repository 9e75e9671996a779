// tcgen05 / TMA bf16 tensor-core path (placeholder until the kernels land; nothing routes here
// while tc_core_supported() is false).
#include "tc.cuh"

struct TcCorePack { int dummy; };

bool tc_core_supported(int, int, int) { return false; }
int tc_core_pack(gnb_ctx*, const gnb_block_params&, const gnb_ffn_params*, const gnb_ln_params*,
                 const gnb_ln_params*, TcCorePack** out) { *out = nullptr; return GNB_OK; }
void tc_core_pack_free(TcCorePack* p) { delete p; }
int tc_core_forward(gnb_ctx*, const gnb_graph*, const TcCorePack*, const gnb_block_params&, const gnb_ffn_params*,
                    const gnb_ln_params*, const gnb_ln_params*, const float*, const float*, const float*, float*,
                    float*, float*) {
  gnb_set_error("tcgen05 path not built");
  return GNB_ERR_UNSUPPORTED;
}
