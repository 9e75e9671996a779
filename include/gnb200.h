/* gnb200.h - C ABI of the B200-native GraphNets.jl GNBlock / GNCore forward path.
 *
 * The reference (JuliaMLTools/GraphNets.jl v0.1.7) has no FFI of its own: its boundary is the
 * Julia API (src/GraphNets.jl:12-50).  Every entry point below names the reference function it
 * replaces; the Julia `ccall` shim (julia/GraphNetsB200.jl) and the Python ctypes host mirror
 * (graphnets.jl_b200/) bind exactly these symbols.
 *
 * Conventions
 *  - Every function returns 0 (GNB_OK) or a negative error code and never throws across the
 *    ABI; gnb_last_error() returns a thread-local message.
 *  - Arrays are fp32, column-major in Julia terms with the FEATURE dimension contiguous:
 *    Julia (D, T) == C [T][D].  Feature tensors at the ABI are COMPACT: ef (DE, E),
 *    nf (DN, N), gf (DG, B) in the order of flatunpaddedef / flatunpaddednf
 *    (src/views.jl:80-98): graph-major, then ascending padded slot i + PN*j.
 *  - Weights keep Flux's layout: Dense.weight is (out, in) column-major, i.e. C [in][out].
 *  - Pointers are DEVICE pointers unless the function name ends in _host.
 *  - A width of 0 means `nothing` (src/gnblock.jl:71-78); the matching pointer may be NULL.
 *  - One gnb_ctx per GPU and per host thread; a ctx is not thread-safe.
 */
#ifndef GNB200_H
#define GNB200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define GNB_OK               0
#define GNB_ERR_INVALID     -1   /* AssertionError in the reference (src/checks.jl) */
#define GNB_ERR_CUDA        -2
#define GNB_ERR_OOM         -3
#define GNB_ERR_UNSUPPORTED -4
#define GNB_ERR_TIMEOUT     -5   /* a kernel's watchdog fired (barrier wait > GNB_WATCHDOG_MS, default 10 s): the kernel
                                    drained and exited, the CUDA context stays usable, the results of that forward are invalid */

#define GNB_PREC_FP32 0   /* CUDA-core fp32 path, parity 1e-5 relative                       */
#define GNB_PREC_BF16 2   /* tcgen05 bf16 tensor-core path (fp32 accumulate), parity 1e-2   */
#define GNB_PREC_AUTO 3   /* bf16 tensor path where the layer shape supports it, else fp32   */

#define GNB_ADJ_F32 0
#define GNB_ADJ_U8  1
#define GNB_ADJ_I32 2
#define GNB_ADJ_BITS 3   /* bit-packed `isone` mask: bit (i + PN*j) of graph b in little-endian uint32 words, every graph starting
                            on a word boundary (ceil(PN*PN/32) words per graph) - 8x fewer bytes than the uint8 mask */

/* LayerNorm denominator (Flux `normalise`, third-party) */
#define GNB_EPS_SQRT_VAR_EPS2 0  /* sqrt(var + eps^2)  - default (Flux 0.14)  */
#define GNB_EPS_STD_PLUS_EPS  1  /* std + eps          - older Flux           */
#define GNB_EPS_SQRT_VAR_EPS  2  /* sqrt(var + eps)    - PyTorch              */

typedef struct gnb_ctx   gnb_ctx;
typedef struct gnb_graph gnb_graph;
typedef struct gnb_model gnb_model;

/* GNBlock((in_e,in_n,in_g) => (out_e,out_n,out_g))   src/gnblock.jl:47-61
 *   We (out_e, in_e+2*in_n+in_g)   input rows ordered [e | v_src | v_dst | u]   src/edgefninput.jl:2-7
 *   Wn (out_n, out_e+in_n+in_g)    input rows ordered [sum_in e' | v | u]       src/nodefninput.jl:2-6
 *   Wg (out_g, out_e+out_n+in_g)   input rows ordered [sum e' | sum v' | u]     src/graphfninput.jl:2-6 */
typedef struct {
  int32_t in_e, in_n, in_g, out_e, out_n, out_g;
  const float *We, *be, *Wn, *bn, *Wg, *bg;
} gnb_block_params;

/* FeedForward(d): Dense(d=>4d, relu) -> Dense(4d=>d)   src/gnfeedforward.jl:27-31 */
typedef struct { const float *W1, *b1, *W2, *b2; } gnb_ffn_params;

/* Flux.LayerNorm(d): gamma (scale), beta (bias), eps   src/gngraphnorm.jl:13-15 */
typedef struct { const float *gamma, *beta; float eps; int32_t eps_mode; } gnb_ln_params;

/* GNCore(dims): x + block(gn1(x)) + ffwd(gn2(x))       src/gncore.jl:46-59; index 0/1/2 = e/n/g */
typedef struct {
  gnb_block_params block;
  gnb_ffn_params   ffn[3];
  gnb_ln_params    ln1[3], ln2[3];
} gnb_core_params;

#define GNB_LAYER_BLOCK 0
#define GNB_LAYER_CORE  1
typedef struct { int32_t kind; int32_t _pad; gnb_block_params block; gnb_core_params core; } gnb_layer;

/* ---------------------------------------------------------------- context ----------- */
int         gnb_version(void);
const char* gnb_last_error(void);
gnb_ctx*    gnb_ctx_create(int device, int* err);
int         gnb_ctx_destroy(gnb_ctx*);
/* Launch on this CUDA stream (a cudaStream_t; NULL = legacy default stream). */
int         gnb_ctx_set_stream(gnb_ctx*, void* cuda_stream);
int         gnb_sync(gnb_ctx*);
/* number of kernels this ctx has launched since creation (bench `gpu_launches`) */
int64_t     gnb_ctx_launch_count(const gnb_ctx*);
/* Per-kernel CUDA-event profile (bench.py roofline): when on, every launch is bracketed by events
 * on the ctx stream; gnb_ctx_profile_read synchronises, returns the totals per kernel kind since the
 * last read (algorithmic bytes / flops as stated in DESIGN.md) and clears them. */
typedef struct { char name[48]; int64_t launches; double ms; double alg_bytes; double alg_flops; } gnb_prof_entry;
int         gnb_ctx_set_profiling(gnb_ctx*, int on);
int         gnb_ctx_profile_read(gnb_ctx*, gnb_prof_entry* out, int cap, int* n);
/* Diagnostics of the fused tcgen05 edge kernel: the first call (out == NULL) switches on clock64 stamping
 * of one steady-state tile pair per CTA; later calls copy the stamps [CTA 148][warp 18][slot 32] of the
 * most recent edge launch to `out` (n = capacity in 64-bit words).  Used by tools/edge_timing.py (test-only library variant "timing") only. */
int         gnb_debug_tc_timing(unsigned long long* out, int n);

/* ---------------------------------------------------------------- lowering ---------- */
/* Replaces GNGraphBatch(adj_mats) (src/gngraphbatch.jl:33-54) and padadjmats (src/pad.jl:1-10):
 * dense adjacency -> device-resident receiver-sorted COO + CSR.
 *   adj      (PN, PN, Badj) column-major: element (i,j,b) at i + PN*j + PN*PN*b; an edge is an
 *            entry equal to one (`isone`, src/pad.jl:30); sender = row i, receiver = column j.
 *   n_nodes  HOST int32[Badj]; rows/columns >= n_nodes[b] are ignored (zero padding).
 *   Badj     1 (single-adjacency mode, structure shared by all B graphs, src/batch.jl:66) or B. */
int gnb_graph_lower(gnb_ctx*, const void* adj, int adj_dtype, int adj_on_device,
                    const int32_t* n_nodes, int PN, int Badj, int B, gnb_graph** out);
/* The same lowering from COO edge lists, without the dense detour (batch on edge lists: src/batch.jl:53-64 + the `findall`
 * of src/pad.jl:26-46).  Graph b owns edges [graph_edge_ptr[b], graph_edge_ptr[b+1]) with LOCAL node ids sender src[e] (row i)
 * and receiver dst[e] (column j) in [0, n_nodes[b]).  Each graph's edges must be strictly ascending in the padded slot
 * src + PN*dst (receiver-major) - the order in which the reference lists the active entries of an adjacency matrix and in
 * which edge features are given; anything else is GNB_ERR_INVALID.  src / dst: int32, HOST or device (coo_on_device);
 * graph_edge_ptr [B+1] and n_nodes [B]: HOST.  PN <= 0: the largest n_nodes (padadjmats, src/pad.jl:3).
 * The resulting index is bit-identical to gnb_graph_lower on the equivalent adjacency matrices. */
int gnb_graph_from_coo(gnb_ctx*, const int32_t* src, const int32_t* dst, int coo_on_device,
                       const int32_t* graph_edge_ptr, const int32_t* n_nodes, int PN, int B, gnb_graph** out);
int gnb_graph_destroy(gnb_graph*);
int gnb_graph_counts(const gnb_graph*, int64_t* E, int64_t* N, int32_t* B, int32_t* PN);
/* Copy the index to HOST buffers (any may be NULL).  edge_src/edge_dst are global compact node
 * ids, edge_slot the 0-based padded slot i + PN*j, *_ptr are CSR offsets. */
int gnb_graph_export_host(gnb_ctx*, const gnb_graph*, int32_t* edge_src, int32_t* edge_dst,
                          int32_t* edge_slot, int32_t* edge_graph,
                          int32_t* graph_edge_ptr /*B+1*/, int32_t* graph_node_ptr /*B+1*/,
                          int32_t* node_in_ptr /*N+1*/);

/* compact <-> padded (DE,PE,B)/(DN,PN,B) layouts: padef/padnf (src/pad.jl:14-63),
 * unpadef/unpadnf (src/unpad.jl:1-17).  Padded slots that are inactive are written as zero. */
int gnb_pad_edges  (gnb_ctx*, const gnb_graph*, const float* ef_compact, int D, float* ef_padded);
int gnb_unpad_edges(gnb_ctx*, const gnb_graph*, const float* ef_padded, int D, float* ef_compact);
int gnb_pad_nodes  (gnb_ctx*, const gnb_graph*, const float* nf_compact, int D, float* nf_padded);
int gnb_unpad_nodes(gnb_ctx*, const gnb_graph*, const float* nf_padded, int D, float* nf_compact);
/* collapsef (src/gngraphbatch.jl:83-85): out (D, PN(PN+1)/2, B); padded in, padded out. */
int gnb_collapse_edges(gnb_ctx*, const gnb_graph*, const float* ef_padded, int D, float* out);

/* ---------------------------------------------------------------- layers ------------ */
/* A model is a left fold of layers (GNCoreList, src/gncorelist.jl:43-45; user-level
 * `decoder o core_list o encoder`, README.md:133-149).  Weights are copied (and, for the
 * tensor-core path, packed to bf16) at creation - the analogue of `model |> gpu`. */
int gnb_model_create(gnb_ctx*, const gnb_layer* layers, int n_layers, int weights_on_device,
                     gnb_model** out);
int gnb_model_destroy(gnb_model*);
int gnb_model_out_dims(const gnb_model*, int32_t* out_e, int32_t* out_n, int32_t* out_g);
/* (m::GNBlock)(x) src/gnblock.jl:63-69, (m::GNCore)(x) src/gncore.jl:56-59,
 * (m::GNCoreList)(x) src/gncorelist.jl:43-45.  Inputs/outputs compact, device. */
int gnb_model_forward(gnb_ctx*, const gnb_model*, const gnb_graph*,
                      const float* ef, const float* nf, const float* gf,
                      float* out_ef, float* out_nf, float* out_gf, int precision);
/* Same call with HOST buffers: stages H2D, runs, copies the outputs D2H, synchronises. */
int gnb_model_forward_host(gnb_ctx*, const gnb_model*, const gnb_graph*,
                           const float* ef, const float* nf, const float* gf,
                           float* out_ef, float* out_nf, float* out_gf, int precision);
/* Single-layer conveniences over the same engine (weights used in place, device pointers). */
int gnb_block_forward(gnb_ctx*, const gnb_graph*, const gnb_block_params*,
                      const float* ef, const float* nf, const float* gf,
                      float* out_ef, float* out_nf, float* out_gf, int precision);
int gnb_core_forward(gnb_ctx*, const gnb_graph*, const gnb_core_params*,
                     const float* ef, const float* nf, const float* gf,
                     float* out_ef, float* out_nf, float* out_gf, int precision);
int gnb_corelist_forward(gnb_ctx*, const gnb_graph*, const gnb_core_params* cores, int n_cores,
                         const float* ef, const float* nf, const float* gf,
                         float* out_ef, float* out_nf, float* out_gf, int precision);

/* ---------------------------------------------------------------- loss ---------------- */
/* Flux.logitcrossentropy over the compact views, as the reference's training example computes it
 * (examples/sort/sort.jl:76-78: logitcrossentropy(flatunpaddednf(y), flatunpaddednf(targets)), likewise for the edges;
 * src/views.jl:80-98):  *loss = mean over the R rows of  -sum_d targets[r][d] * logsoftmax(logits[r])[d].
 * logits / targets: compact (D, R) device matrices (feature dim contiguous); loss: DEVICE float; per_row: optional device
 * [R] (NULL: not written).  Deterministic (fixed-order sums). */
int gnb_logit_cross_entropy(gnb_ctx*, const float* logits, const float* targets, int D, int64_t R,
                            float* loss, float* per_row);
/* dlogits = scale * d loss / d logits of the mean cross-entropy above (the cotangent the backward pass starts from) */
int gnb_logit_cross_entropy_bwd(gnb_ctx*, const float* logits, const float* targets, int D, int64_t R, float scale, float* dlogits);

/* ---------------------------------------------------------------- training step ---------------- */
/* Primitive operators of the backward pass of the GNBlock / GNCore forward (the reference differentiates it with Zygote inside
 * Flux.withgradient, examples/sort/sort.jl:122-132; forward: src/gnblock.jl:63-69, src/gncore.jl:56-68).  fp32, device pointers,
 * deterministic (no atomics).  The host side (graphnets.jl_b200/train.py) composes them into the adjoint of each layer; the
 * flat gradient buffer is all-reduced over NCCL by the caller (torch.distributed) and applied with gnb_op_adamw. */
typedef struct { const float* x; int d, ldx; const float* W; const float* gamma; const float* beta; float eps; int eps_mode; } gnb_lin_src;
typedef struct { const float* a; const int32_t* idx; int lda; } gnb_lin_add;
typedef struct {
  int64_t R; int Nout, ldw, nsrc;      /* out = act( sum_s LN_s(x_s) W_s + bias + sum_j add_j[idx_j] ),  W_s: [d_s][ldw] */
  gnb_lin_src src[3];
  const float* bias;
  int nadd; gnb_lin_add add[4];
  int relu;
  float* out; int ldo;
  int precision;                       /* GNB_PREC_FP32 | GNB_PREC_BF16 / GNB_PREC_AUTO: bf16 operands on the tensor cores where the shape allows
                                          (rows >= 256, widths multiples of 64 / 128), fp32 accumulation */
} gnb_lin_args;
int gnb_op_linear(gnb_ctx*, const gnb_lin_args*);      /* the forward's fused linear kernel (csrc/fp32.cu) */
/* out[s] = sum over p in [ptr[s], ptr[s+1]) of x[perm ? perm[p] : p], ascending p */
int gnb_op_segsum(gnb_ctx*, const float* x, int D, const int32_t* ptr, int64_t S, const int32_t* perm, float* out);
int gnb_op_layernorm(gnb_ctx*, const float* x, int64_t R, int D, const float* gamma, const float* beta, float eps, int eps_mode, float* y);
/* dx += adjoint of LayerNorm at x applied to g;  gxhat = g * xhat (column sums: d gamma; column sums of g: d beta) */
int gnb_op_layernorm_bwd(gnb_ctx*, const float* x, const float* g, int64_t R, int D, const float* gamma, float eps, int eps_mode,
                         float* dx, float* gxhat);
/* dW[k][n] += sum_r X[idx ? idx[r] : r][k] * dY[r][n]      (dW: [K][ldw], the layout of a Flux Dense.weight block) */
/* precision GNB_PREC_BF16 / AUTO: bf16 operands on the tensor cores (split over row chunks, fp32 accumulation) when R >= 4096 */
int gnb_op_wgrad(gnb_ctx*, const float* X, int ldx, int K, const int32_t* idx, const float* dY, int ldy, int N, int64_t R, float* dW, int ldw,
                 int precision);
int gnb_op_colsum(gnb_ctx*, const float* X, int ldx, int D, int64_t R, float* out /* += */);
int gnb_op_relu_mask(gnb_ctx*, float* t, const float* h, int64_t n);      /* t[i] = h[i] > 0 ? t[i] : 0 */
/* out[r] = a[r] + b1[idx1 ? idx1[r] : r] + b2[idx2 ? idx2[r] : r]   (a, b1, b2 optional; rows of width D) */
int gnb_op_gather_add(gnb_ctx*, float* out, const float* a, const float* b1, const int32_t* idx1, const float* b2, const int32_t* idx2,
                      int64_t R, int D);
int gnb_op_transpose(gnb_ctx*, const float* in, int rows, int cols, int ld_in, float* out /* [cols][rows] */);
int gnb_op_adamw(gnb_ctx*, float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                 float weight_decay, int step);
/* device pointers of the lowered index (owned by the graph) */
int gnb_graph_device_index(const gnb_graph*, const int32_t** edge_src, const int32_t** edge_dst, const int32_t** edge_graph,
                           const int32_t** node_graph, const int32_t** graph_edge_ptr, const int32_t** graph_node_ptr,
                           const int32_t** node_in_ptr);

#ifdef __cplusplus
}
#endif
#endif
