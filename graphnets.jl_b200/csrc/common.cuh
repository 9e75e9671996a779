// Internal structures shared by the gnb200 translation units (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include <vector>
#include <atomic>
#include <string>
#include "../../include/gnb200.h"

void gnb_set_error(const char* fmt, ...);

#define GNB_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (call);                                                             \
    if (_e != cudaSuccess) {                                                             \
      gnb_set_error("CUDA error %s at %s:%d (%s)", cudaGetErrorString(_e), __FILE__,     \
                    __LINE__, #call);                                                    \
      return (_e == cudaErrorMemoryAllocation) ? GNB_ERR_OOM : GNB_ERR_CUDA;             \
    }                                                                                    \
  } while (0)

#define GNB_CHECK(cond, ...)                                                             \
  do {                                                                                   \
    if (!(cond)) {                                                                       \
      gnb_set_error(__VA_ARGS__);                                                        \
      return GNB_ERR_INVALID;                                                            \
    }                                                                                    \
  } while (0)

#define GNB_TRY(expr)                                                                    \
  do {                                                                                   \
    int _r = (expr);                                                                     \
    if (_r != GNB_OK) return _r;                                                         \
  } while (0)

// Scratch arena: bump allocation out of cudaMalloc'd chunks, reset at the start of every
// forward.  Chunks persist across calls so steady-state forwards do no cudaMalloc.
struct Arena {
  struct Chunk { char* base; size_t cap; size_t used; };
  std::vector<Chunk> chunks;
  uint64_t gen = 0;      // bumped whenever the set of chunks changes (pointers handed out before are then stale)
  size_t min_chunk = (size_t)64 << 20;
  int alloc(size_t bytes, void** out);
  void reset();
  void release();
};

// Kernel watchdog arguments of the tcgen05 kernels (tc_ptx.cuh): device-global abort flag + wait limit.
struct WatchArgs {
  int* flag;                        // gnb_ctx::d_abort
  unsigned long long limit_ns;      // GNB_WATCHDOG_MS (default 10 s)
};

struct ProfRec { int tag; cudaEvent_t a, b; double bytes, flops; };
struct ProfTag { const char* name; int64_t launches; double ms, bytes, flops; };

// A whole forward captured as a CUDA graph (model.cu::gnb_model_forward): keyed by everything the enqueued work depends on.
struct FwdGraph {
  uint64_t model_id, graph_uid, arena_gen;
  const void* ptr[6];
  int precision, env_sig;
  int seen;                  // eager calls with this key so far (the graph is captured on the second one)
  cudaGraphExec_t exec;      // nullptr: not captured (yet), or capture not possible
  bool no_graph;
  int64_t launches;          // kernels in the graph (gnb_ctx::launches accounting)
  uint64_t last_use;
};

struct gnb_ctx {
  int device = 0;
  bool profiling = false;
  void* train_ws = nullptr;      // scratch of the two-stage reductions of the training operators (train.cu), grown on demand
  size_t train_ws_bytes = 0;
  // training: weights change every step, so the tensor-core linear layer packs its bf16 weight slabs per call into this
  // stream-ordered scratch instead of the per-model cache (tc_gemm.cu::get_pack)
  bool tc_lin_nocache = false;
  void* pack_ws = nullptr;
  size_t pack_ws_bytes = 0;
  std::vector<FwdGraph> fwd_graphs;
  uint64_t fwd_tick = 0;
  // the legacy default stream cannot be captured: forwards bound to it are captured / replayed on this side stream, forked from
  // and joined back into the caller's stream with events (same ordering as an eager forward)
  cudaStream_t gstream = nullptr;
  cudaEvent_t g_fork = nullptr, g_join = nullptr;
  std::vector<ProfRec> prof_recs;
  std::vector<ProfTag> prof_tags;
  std::vector<cudaEvent_t> ev_pool;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  Arena arena;
  int64_t launches = 0;
  // host-variant staging (device side), grown on demand
  Arena staging;
  // generic bf16 tensor-core linear layers (tc_gemm.cu): enabled per forward by the precision mode; packed weights are
  // cached per (model id, weight block)
  // one-time per-device setup (cudaFuncSetAttribute, memory-pool attributes) is tracked per context, not per process:
  // a process may own one context per GPU
  uint64_t once_mask = 0;
  cudaEvent_t done_ev = nullptr;      // end of this context's last forward (cross-context serialisation, model.cu)
  // host-buffer forward (gnb_model_forward_host): PCIe copies overlapped with compute on a second stream.  The edge input is
  // uploaded in row chunks and the first consumer (the encoder's edge kernel) starts on a chunk as soon as it has landed; the
  // edge output is downloaded as soon as the decoder's edge kernel has written it, under the node / graph kernels.
  struct HostPipe {
    cudaStream_t copy = nullptr;
    cudaEvent_t ev_chunk[8] = {}, ev_ready = nullptr, ev_in = nullptr;
    const float* d_ef = nullptr; int64_t ef_rows = 0; int nchunks = 0; bool ef_pending = false;      // upload in flight
    const float* d_out_ef = nullptr; float* h_out_ef = nullptr; size_t out_bytes = 0; bool out_pending = false, out_early = false;
  } pipe;
  bool use_tc_lin = false;
  uint64_t cur_model_id = 0;
  void* lin_cache = nullptr;
  // kernel watchdog: raised by a tcgen05 kernel whose mbarrier wait timed out; mirrored to pinned host memory behind every forward
  int* d_abort = nullptr;
  int* h_abort = nullptr;
  unsigned long long wd_limit_ns = 10000000000ull;
  // test hook (GNB_DEBUG_PROJ_DRAIN_DELAY_NS): stalls the accumulator-drain warps of k_tc_proj per tile, to exercise the
  // barrier protocol with one role far behind the others
  unsigned int dbg_proj_drain_delay_ns = 0;
};
static inline WatchArgs ctx_watch(const gnb_ctx* c) { return WatchArgs{c->d_abort, c->wd_limit_ns}; }
// Returns GNB_ERR_TIMEOUT (and clears the flag) if a kernel watchdog fired; call after the stream has been synchronised.
int ctx_check_watchdog(gnb_ctx* c);

enum { ONCE_EDGE5 = 0, ONCE_PROJ_LN2, ONCE_PROJ_AGG1, ONCE_PROJ_OTHER, ONCE_TC_LIN, ONCE_TC_FFN, ONCE_TC_FFN384, ONCE_WIDE, ONCE_NARROW2, ONCE_POOL, ONCE_AGG2, ONCE_TC_WGRAD, ONCE_GRAPH_POST };
// true exactly once per (context, key)
static inline bool ctx_first(gnb_ctx* c, int key) {
  const uint64_t b = 1ull << key;
  if (c->once_mask & b) return false;
  c->once_mask |= b;
  return true;
}

// Cut of the partial-row index of the tensor path (edge_part / node_gpart): a partial row never spans a 16-row boundary,
// so that each of the eight 16-row epilogue warps of the fused kernel (tc_edge.cu) owns whole partial rows.
constexpr int GNB_PART_ROWS = 16;

inline std::atomic<uint64_t> g_graph_uids{1};
struct gnb_graph {
  int device = 0;
  uint64_t uid = g_graph_uids.fetch_add(1);      // never reused (a destroyed graph's address may be)
  int32_t B = 0, PN = 0;
  int64_t E = 0, N = 0;
  // device arrays
  int32_t* edge_src = nullptr;    // [E] global compact node id of the sender (row i)
  int32_t* edge_dst = nullptr;    // [E] global compact node id of the receiver (column j)
  int32_t* edge_slot = nullptr;   // [E] padded slot i + PN*j within the graph
  int32_t* edge_graph = nullptr;  // [E]
  int32_t* node_graph = nullptr;  // [N]
  int32_t* graph_edge_ptr = nullptr;  // [B+1]
  int32_t* graph_node_ptr = nullptr;  // [B+1]
  int32_t* node_in_ptr = nullptr;     // [N+1] CSR over receivers (edges are receiver-sorted)
  // tensor-core path: per (GNB_PART_ROWS-edge block, receiver) run one partial row.
  // node v sums partial rows [node_part_ptr[v], node_part_ptr[v+1])  (deterministic, no atomics)
  int32_t* edge_part = nullptr;       // [E]   partial-row id of each edge
  int32_t* node_part_ptr = nullptr;   // [N+1]
  int32_t* graph_part_ptr = nullptr;  // [B+1] partial rows of graph b = [graph_part_ptr[b], graph_part_ptr[b+1])
  int64_t n_parts = 0;
  // same for the node -> graph sums of the tensor path: partial rows per (GNB_PART_ROWS-node block, graph) run
  int32_t* node_gpart = nullptr;       // [N]   partial-row id of each node
  int32_t* graph_npart_ptr = nullptr;  // [B+1] partial rows of graph b = [graph_npart_ptr[b], graph_npart_ptr[b+1])
  int64_t n_nparts = 0;
  void* all = nullptr;  // single (stream-ordered) allocation backing everything above
  cudaStream_t stream = nullptr;   // stream the allocation is ordered on
};

// RAII bracket of one kernel launch: counts it and, when profiling, times it with CUDA events.
struct Launch {
  gnb_ctx* c;
  int rec = -1;
  Launch(gnb_ctx* ctx, const char* name, double alg_bytes = 0, double alg_flops = 0);
  ~Launch();
};

template <typename T>
static inline T* arena_ptr(Arena& a, size_t n, int* rc) {
  void* p = nullptr;
  int r = a.alloc(n * sizeof(T), &p);
  if (r != GNB_OK) *rc = r;
  return (T*)p;
}

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
