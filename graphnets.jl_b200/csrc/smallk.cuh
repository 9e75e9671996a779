// fp32 kernels for GNBlocks with narrow inputs (encoder) or narrow outputs (decoder); internal interface.
#pragma once
#include "common.cuh"

// one piece of the concatenated narrow input row:  in_p[idx ? idx[r] : r][0..d)
struct WidePiece {
  const float* x;
  const int32_t* idx;
  int d, ldx;
  const float* W;      // [d][ldw] k-major rows of the Dense weight that multiply this piece
};
struct WideArgs {
  int64_t R;
  int Nout, ldw, np;
  WidePiece pc[5];
  const float* bias;   // [Nout] or nullptr
  float* out;          // [R][ldo]
  int ldo;
  int K4;              // (set by the launcher)
};
int launch_wide(gnb_ctx* ctx, const WideArgs& a);

// Z[v] = [ sum_{e->v} e_e ; sum_{e->v} v_src(e) ; deg v_v ; deg u_g ; deg ]   width de + 2 dn + dg + 1 <= 32
struct ZsumArgs {
  int64_t N;
  const int32_t *node_in_ptr, *edge_src, *node_graph;
  const float *ef, *nf, *gf;
  int de, dn, dg;
  float* Z;
};
int launch_zsum(gnb_ctx* ctx, const ZsumArgs& a);

constexpr int NARROW_KMAX = 544;
// W: [d][ldw]; optional second weight block W2 [d][ldw]: output columns j >= n1 take W2[k][j - n1] (two narrow projections of the
// same rows in one pass, e.g. the decoder's P_s | P_r)
struct NarrowSrc { const float* x; int d, ldx; const float* W; const float* W2 = nullptr; int n1 = 0; };
struct NarrowAdd { const float* a; const int32_t* idx; int lda; };
struct NarrowArgs {
  int64_t R;
  int No, ldw, nsrc;
  NarrowSrc src[3];
  const float* bias;
  int nadd;
  NarrowAdd add[4];
  float* out;
  int ldo;
};
int launch_narrow(gnb_ctx* ctx, const NarrowArgs& a);
