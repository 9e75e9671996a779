// Batch lowering: dense adjacency -> receiver-sorted COO + CSR, device resident.
// Replaces GNGraphBatch(adj_mats) (reference src/gngraphbatch.jl:33-54,113-211) and the
// padded-slot bookkeeping of src/pad.jl / src/unpad.jl.  Integer work, HBM-bound, bit-exact.
//
// Order of the compact edge list = graph-major, then ascending padded slot k = i + PN*j
// (findall(isone, view(adj,:)), src/pad.jl:30) which is (receiver j, sender i) lexicographic,
// i.e. already receiver-sorted.  One warp owns one adjacency column: ballot + popc give a
// deterministic position for every active entry, no atomics anywhere.
#include "common.cuh"

namespace {

constexpr int SCAN_ITEMS = 2048;  // items per block in the 3-phase scan (256 threads x 8)

__global__ void k_scan_block_sums(const int32_t* __restrict__ in, int64_t n, int32_t* __restrict__ sums) {
  __shared__ int32_t ws[8];
  int64_t base = (int64_t)blockIdx.x * SCAN_ITEMS;
  int32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    int64_t idx = base + i * 256 + threadIdx.x;
    if (idx < n) s += in[idx];
  }
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int32_t t = 0;
    for (int i = 0; i < 8; i++) t += ws[i];
    sums[blockIdx.x] = t;
  }
}

// single block: exclusive scan of `sums` in place, total -> sums[nb]
__global__ void k_scan_sums(int32_t* sums, int nb) {
  __shared__ int32_t part[1024];
  __shared__ int32_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += 1024) {
    int i = base + threadIdx.x;
    int32_t v = i < nb ? sums[i] : 0;
    part[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      int32_t t = threadIdx.x >= o ? part[threadIdx.x - o] : 0;
      __syncthreads();
      part[threadIdx.x] += t;
      __syncthreads();
    }
    int32_t incl = part[threadIdx.x];
    int32_t carry = carry_s;
    if (i < nb) sums[i] = carry + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) sums[nb] = carry_s;
}

// out[i] = exclusive prefix; also out[n] = total when write_total
__global__ void k_scan_apply(const int32_t* __restrict__ in, int64_t n, const int32_t* __restrict__ sums,
                             int32_t* __restrict__ out, int nb, int write_total) {
  __shared__ int32_t tsum[256];
  int64_t base = (int64_t)blockIdx.x * SCAN_ITEMS + (int64_t)threadIdx.x * 8;
  int32_t v[8];
  int32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    v[i] = (base + i < n) ? in[base + i] : 0;
    s += v[i];
  }
  tsum[threadIdx.x] = s;
  __syncthreads();
  for (int o = 1; o < 256; o <<= 1) {
    int32_t t = threadIdx.x >= o ? tsum[threadIdx.x - o] : 0;
    __syncthreads();
    tsum[threadIdx.x] += t;
    __syncthreads();
  }
  int32_t run = sums[blockIdx.x] + tsum[threadIdx.x] - s;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    if (base + i < n) out[base + i] = run;
    run += v[i];
  }
  if (write_total && blockIdx.x == 0 && threadIdx.x == 0) out[n] = sums[nb];
}

// Bit-packed adjacency (GNB_ADJ_BITS): bit (i + PN*j) of graph b, little-endian in 32-bit words, every graph starting on a word
// boundary (words_per_graph = ceil(PN*PN / 32)).  8x less PCIe traffic than the uint8 mask: what the host-buffer path uploads.
struct BitAdj { uint32_t w; };      // tag type: a `const BitAdj*` is the word array
template <typename T>
static __device__ __forceinline__ bool adj_is_one(const T* adj, size_t graph, int PN, int i, int j) {
  return adj[(graph * PN + j) * PN + i] == (T)1;
}
static __device__ __forceinline__ bool adj_is_one(const BitAdj* adj, size_t graph, int PN, int i, int j) {
  const size_t wpg = ((size_t)PN * PN + 31) / 32;
  const uint32_t bit = (uint32_t)(i + PN * j);
  return (adj[graph * wpg + (bit >> 5)].w >> (bit & 31)) & 1u;
}

template <typename T>
__global__ void k_count_cols(const T* __restrict__ adj, const int32_t* __restrict__ n_nodes, int PN,
                             int Badj, int64_t total_cols, int32_t* __restrict__ colcount) {
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= total_cols) return;
  int b = (int)(warp / PN), j = (int)(warp % PN);
  int ba = Badj == 1 ? 0 : b;
  int n = n_nodes[ba];
  int cnt = 0;
  if (j < n) {
    for (int i0 = 0; i0 < n; i0 += 32) {
      int i = i0 + lane;
      bool a = (i < n) && adj_is_one(adj, (size_t)ba, PN, i, j);
      cnt += __popc(__ballot_sync(0xffffffffu, a));
    }
  }
  if (lane == 0) colcount[warp] = cnt;
}

template <typename T>
__global__ void k_fill_edges(const T* __restrict__ adj, const int32_t* __restrict__ n_nodes, int PN,
                             int Badj, int64_t total_cols, const int32_t* __restrict__ coloff,
                             const int32_t* __restrict__ gnp, int32_t* __restrict__ edge_src,
                             int32_t* __restrict__ edge_dst, int32_t* __restrict__ edge_slot,
                             int32_t* __restrict__ edge_graph, int32_t* __restrict__ node_in_ptr,
                             int32_t* __restrict__ node_graph, int32_t* __restrict__ graph_edge_ptr) {
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= total_cols) return;
  int b = (int)(warp / PN), j = (int)(warp % PN);
  int ba = Badj == 1 ? 0 : b;
  int n = n_nodes[ba];
  int off = coloff[warp];
  if (j == 0 && lane == 0) graph_edge_ptr[b] = off;
  if (j >= n) return;
  int nbase = gnp[b];
  if (lane == 0) {
    node_in_ptr[nbase + j] = off;
    node_graph[nbase + j] = b;
  }
  for (int i0 = 0; i0 < n; i0 += 32) {
    int i = i0 + lane;
    bool a = (i < n) && adj_is_one(adj, (size_t)ba, PN, i, j);
    unsigned m = __ballot_sync(0xffffffffu, a);
    if (a) {
      int e = off + __popc(m & ((1u << lane) - 1u));
      edge_src[e] = nbase + i;
      edge_dst[e] = nbase + j;
      edge_slot[e] = i + PN * j;
      edge_graph[e] = b;
    }
    off += __popc(m);
  }
}

__global__ void k_fill_const(int32_t* p, int64_t n, int32_t v) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// run-start flags for the 128-edge-tile partial sums of the tensor-core path
__global__ void k_part_flags(const int32_t* __restrict__ edge_dst, int64_t E, int tile, int32_t* __restrict__ flag) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  flag[e] = (e % tile == 0 || edge_dst[e] != edge_dst[e - 1]) ? 1 : 0;
}
// part id of edge e = (inclusive scan of flags)[e] - 1 = excl[e] + flag[e] - 1
__global__ void k_part_finish(const int32_t* __restrict__ flag, const int32_t* __restrict__ excl, int64_t E,
                              int32_t* __restrict__ edge_part) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < E) edge_part[e] = excl[e] + flag[e] - 1;
}
__global__ void k_node_part_ptr(const int32_t* __restrict__ node_in_ptr, const int32_t* __restrict__ excl /*E+1*/,
                                int64_t N, int32_t* __restrict__ node_part_ptr) {
  int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v <= N) node_part_ptr[v] = excl[node_in_ptr[v]];
}

__global__ void k_graph_part_ptr(const int32_t* __restrict__ gnp, const int32_t* __restrict__ npp, int B,
                                 int32_t* __restrict__ out) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b <= B) out[b] = npp[gnp[b]];
}

// ---- pad / unpad / collapse ---------------------------------------------------------
__global__ void k_pad_edges(const float* __restrict__ src, const int32_t* __restrict__ edge_slot,
                            const int32_t* __restrict__ edge_graph, int64_t E, int D, int64_t PE,
                            float* __restrict__ dst, int to_padded) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= E * D) return;
  int64_t e = t / D;
  int d = (int)(t % D);
  int64_t p = ((int64_t)edge_graph[e] * PE + edge_slot[e]) * D + d;
  if (to_padded) dst[p] = src[t];
  else dst[t] = src[p];
}
__global__ void k_pad_nodes(const float* __restrict__ src, const int32_t* __restrict__ node_graph,
                            const int32_t* __restrict__ gnp, int64_t N, int D, int PN,
                            float* __restrict__ dst, int to_padded) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N * D) return;
  int64_t v = t / D;
  int d = (int)(t % D);
  int b = node_graph[v];
  int64_t p = ((int64_t)b * PN + (v - gnp[b])) * D + d;
  if (to_padded) dst[p] = src[t];
  else dst[t] = src[p];
}
// collapsef (src/gngraphbatch.jl:69-85): lower-triangular coordinate c (column-major, i>=j)
// gets (ef[i+PN*j] + ef[j+PN*i]) / 2.
__global__ void k_collapse(const float* __restrict__ ef, int B, int PN, int D, float* __restrict__ out) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = (int64_t)B * PN * PN * D;
  if (t >= total) return;
  int d = (int)(t % D);
  int64_t r = t / D;
  int i = (int)(r % PN);
  r /= PN;
  int j = (int)(r % PN);
  int b = (int)(r / PN);
  if (i < j) return;
  int64_t PE = (int64_t)PN * PN;
  int64_t C = (int64_t)PN * (PN + 1) / 2;
  int64_t c = (int64_t)j * PN - (int64_t)j * (j - 1) / 2 + (i - j);
  float a = ef[((int64_t)b * PE + i + (int64_t)PN * j) * D + d];
  float bb = ef[((int64_t)b * PE + j + (int64_t)PN * i) * D + d];
  out[((int64_t)b * C + c) * D + d] = (a + bb) / 2.0f;
}

int exclusive_scan(gnb_ctx* ctx, const int32_t* in, int64_t n, int32_t* out, int write_total, int32_t* sums_buf) {
  int nb = ceil_div(n, SCAN_ITEMS);
  if (nb == 0) {
    if (write_total) GNB_CUDA(cudaMemsetAsync(out, 0, sizeof(int32_t), ctx->stream));
    return GNB_OK;
  }
  k_scan_block_sums<<<nb, 256, 0, ctx->stream>>>(in, n, sums_buf);
  k_scan_sums<<<1, 1024, 0, ctx->stream>>>(sums_buf, nb);
  k_scan_apply<<<nb, 256, 0, ctx->stream>>>(in, n, sums_buf, out, nb, write_total);
  ctx->launches += 3;
  GNB_CUDA(cudaGetLastError());
  return GNB_OK;
}

template <typename T>
int lower_typed(gnb_ctx* ctx, const T* adj_dev, const int32_t* nn_dev, int PN, int Badj, int B,
                int32_t* colcount, int32_t* coloff, int32_t* sums, gnb_graph* g, int phase) {
  int64_t total_cols = (int64_t)B * PN;
  int blocks = ceil_div(total_cols * 32, 256);
  if (phase == 0) {
    k_count_cols<T><<<blocks, 256, 0, ctx->stream>>>(adj_dev, nn_dev, PN, Badj, total_cols, colcount);
    ctx->launches++;
    GNB_CUDA(cudaGetLastError());
    GNB_TRY(exclusive_scan(ctx, colcount, total_cols, coloff, 1, sums));
  } else {
    k_fill_edges<T><<<blocks, 256, 0, ctx->stream>>>(adj_dev, nn_dev, PN, Badj, total_cols, coloff,
                                                     g->graph_node_ptr, g->edge_src, g->edge_dst,
                                                     g->edge_slot, g->edge_graph, g->node_in_ptr,
                                                     g->node_graph, g->graph_edge_ptr);
    ctx->launches++;
    GNB_CUDA(cudaGetLastError());
  }
  return GNB_OK;
}

size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// one stream-ordered allocation backing every index array of the graph (B, E, N set by the caller)
int graph_alloc(gnb_ctx* ctx, gnb_graph* g) {
  const int64_t E = g->E, N = g->N;
  const int B = g->B;
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off += align_up(n * sizeof(int32_t)); return o; };
  size_t o_src = take(E), o_dst = take(E), o_slot = take(E), o_eg = take(E), o_ng = take(N);
  size_t o_gep = take(B + 1), o_gnp = take(B + 1), o_nip = take(N + 1), o_ep = take(E), o_npp = take(N + 1);
  size_t o_gpp = take(B + 1);
  size_t o_ngp = take(N), o_gnpp = take(B + 1);
  // stream-ordered pool allocation: a lowering per batch (the e2e path) must not pay cudaMalloc / cudaFree device syncs
  if (ctx_first(ctx, ONCE_POOL)) {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess) {
      uint64_t thr = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    cudaGetLastError();
  }
  g->stream = ctx->stream;
  cudaError_t ce = cudaMallocAsync(&g->all, off ? off : 256, ctx->stream);
  if (ce != cudaSuccess) {
    gnb_set_error("graph lowering: cudaMalloc(%zu) failed: %s", off, cudaGetErrorString(ce));
    delete g;
    return GNB_ERR_OOM;
  }
  char* base = (char*)g->all;
  g->edge_src = (int32_t*)(base + o_src); g->edge_dst = (int32_t*)(base + o_dst);
  g->edge_slot = (int32_t*)(base + o_slot); g->edge_graph = (int32_t*)(base + o_eg);
  g->node_graph = (int32_t*)(base + o_ng); g->graph_edge_ptr = (int32_t*)(base + o_gep);
  g->graph_node_ptr = (int32_t*)(base + o_gnp); g->node_in_ptr = (int32_t*)(base + o_nip);
  g->edge_part = (int32_t*)(base + o_ep); g->node_part_ptr = (int32_t*)(base + o_npp);
  g->graph_part_ptr = (int32_t*)(base + o_gpp);
  g->node_gpart = (int32_t*)(base + o_ngp); g->graph_npart_ptr = (int32_t*)(base + o_gnpp);
  return GNB_OK;
}

// sentinels + the partial-row indexes of the tensor path; expects edge_dst, node_in_ptr[0..N), node_graph, graph_node_ptr and
// graph_edge_ptr[0..B) on the device.  Synchronises the stream.
int graph_finish(gnb_ctx* ctx, gnb_graph* g) {
  const int64_t E = g->E, N = g->N;
  const int B = g->B;
  int ret = GNB_OK;
  do {
    // sentinels: graph_edge_ptr[B] = E, node_in_ptr[N] = E
    k_fill_const<<<1, 32, 0, ctx->stream>>>(g->graph_edge_ptr + B, 1, (int32_t)E);
    k_fill_const<<<1, 32, 0, ctx->stream>>>(g->node_in_ptr + N, 1, (int32_t)E);
    ctx->launches += 2;
    // tensor-path partial-row index (cut every GNB_PART_ROWS edges)
    if (E > 0) {
      int rc2 = GNB_OK;
      int32_t* flag = arena_ptr<int32_t>(ctx->arena, E, &rc2);
      int32_t* excl = arena_ptr<int32_t>(ctx->arena, E + 1, &rc2);
      int32_t* sums2 = arena_ptr<int32_t>(ctx->arena, ceil_div(E, SCAN_ITEMS) + 2, &rc2);
      if (rc2 != GNB_OK) { ret = rc2; break; }
      int eb = ceil_div(E, 256);
      k_part_flags<<<eb, 256, 0, ctx->stream>>>(g->edge_dst, E, GNB_PART_ROWS, flag);
      ctx->launches++;
      if ((ret = exclusive_scan(ctx, flag, E, excl, 1, sums2)) != GNB_OK) break;
      k_part_finish<<<eb, 256, 0, ctx->stream>>>(flag, excl, E, g->edge_part);
      k_node_part_ptr<<<ceil_div(N + 1, 256), 256, 0, ctx->stream>>>(g->node_in_ptr, excl, N, g->node_part_ptr);
      ctx->launches += 2;
      int32_t np = 0;
      if (cudaMemcpyAsync(&np, excl + E, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) { ret = GNB_ERR_CUDA; break; }
      if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { ret = GNB_ERR_CUDA; break; }
      g->n_parts = np;
    } else {
      cudaMemsetAsync(g->node_part_ptr, 0, sizeof(int32_t) * (N + 1), ctx->stream);
    }
    k_graph_part_ptr<<<ceil_div(B + 1, 256), 256, 0, ctx->stream>>>(g->graph_node_ptr, g->node_part_ptr, B, g->graph_part_ptr);
    ctx->launches++;
    // node -> graph partial-row index (GNB_PART_ROWS-node blocks, graph runs): same construction keyed by the node's graph
    if (N > 0) {
      int rc2 = GNB_OK;
      int32_t* flag = arena_ptr<int32_t>(ctx->arena, N, &rc2);
      int32_t* excl = arena_ptr<int32_t>(ctx->arena, N + 1, &rc2);
      int32_t* sums2 = arena_ptr<int32_t>(ctx->arena, ceil_div(N, SCAN_ITEMS) + 2, &rc2);
      if (rc2 != GNB_OK) { ret = rc2; break; }
      k_part_flags<<<ceil_div(N, 256), 256, 0, ctx->stream>>>(g->node_graph, N, GNB_PART_ROWS, flag);
      ctx->launches++;
      if ((ret = exclusive_scan(ctx, flag, N, excl, 1, sums2)) != GNB_OK) break;
      k_part_finish<<<ceil_div(N, 256), 256, 0, ctx->stream>>>(flag, excl, N, g->node_gpart);
      k_node_part_ptr<<<ceil_div(B + 1, 256), 256, 0, ctx->stream>>>(g->graph_node_ptr, excl, B, g->graph_npart_ptr);
      ctx->launches += 2;
      int32_t np = 0;
      if (cudaMemcpyAsync(&np, excl + N, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) { ret = GNB_ERR_CUDA; break; }
      if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { ret = GNB_ERR_CUDA; break; }
      g->n_nparts = np;
    } else {
      cudaMemsetAsync(g->graph_npart_ptr, 0, sizeof(int32_t) * (B + 1), ctx->stream);
    }
    cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
    if (e2 == cudaSuccess) e2 = cudaGetLastError();
    if (e2 != cudaSuccess) { gnb_set_error("graph lowering: %s", cudaGetErrorString(e2)); ret = GNB_ERR_CUDA; }
  } while (0);
  return ret;
}

}  // namespace

extern "C" int gnb_graph_lower(gnb_ctx* ctx, const void* adj, int adj_dtype, int adj_on_device,
                               const int32_t* n_nodes, int PN, int Badj, int B, gnb_graph** out) {
  GNB_CHECK(ctx && adj && n_nodes && out, "gnb_graph_lower: null argument");
  GNB_CHECK(B > 0 && PN > 0, "gnb_graph_lower: need B > 0 and PN > 0 (length(adj_mats) > 0, src/checks.jl:8)");
  GNB_CHECK(Badj == 1 || Badj == B, "gnb_graph_lower: Badj must be 1 or B");
  GNB_CHECK(adj_dtype == GNB_ADJ_F32 || adj_dtype == GNB_ADJ_U8 || adj_dtype == GNB_ADJ_I32 || adj_dtype == GNB_ADJ_BITS,
            "gnb_graph_lower: bad adj_dtype");
  GNB_CHECK((int64_t)B * PN * PN < ((int64_t)1 << 31), "gnb_graph_lower: B*PN^2 exceeds int32 slot range");
  int64_t N = 0;
  std::vector<int32_t> gnp(B + 1, 0);
  for (int b = 0; b < B; b++) {
    int n = n_nodes[Badj == 1 ? 0 : b];
    GNB_CHECK(n >= 0 && n <= PN, "gnb_graph_lower: n_nodes[%d]=%d outside [0, PN=%d]", b, n, PN);
    N += n;
    gnp[b + 1] = (int32_t)N;
  }
  GNB_CHECK(N < ((int64_t)1 << 31), "gnb_graph_lower: too many nodes");
  GNB_CUDA(cudaSetDevice(ctx->device));
  size_t esz = adj_dtype == GNB_ADJ_U8 ? 1 : 4;
  size_t adj_bytes = adj_dtype == GNB_ADJ_BITS ? (size_t)Badj * (((size_t)PN * PN + 31) / 32) * 4 : (size_t)Badj * PN * PN * esz;
  // temporaries out of the scratch arena
  ctx->arena.reset();
  int rc = GNB_OK;
  int64_t total_cols = (int64_t)B * PN;
  int32_t* nn_dev = arena_ptr<int32_t>(ctx->arena, Badj, &rc);
  int32_t* colcount = arena_ptr<int32_t>(ctx->arena, total_cols, &rc);
  int32_t* coloff = arena_ptr<int32_t>(ctx->arena, total_cols + 1, &rc);
  int32_t* sums = arena_ptr<int32_t>(ctx->arena, ceil_div(total_cols, SCAN_ITEMS) + 2, &rc);
  const void* adj_dev = adj;
  if (!adj_on_device) {
    void* tmp = nullptr;
    int r2 = ctx->arena.alloc(adj_bytes, &tmp);
    if (r2 != GNB_OK) rc = r2;
    adj_dev = tmp;
  }
  if (rc != GNB_OK) return rc;
  if (!adj_on_device) GNB_CUDA(cudaMemcpyAsync((void*)adj_dev, adj, adj_bytes, cudaMemcpyHostToDevice, ctx->stream));
  GNB_CUDA(cudaMemcpyAsync(nn_dev, n_nodes, sizeof(int32_t) * Badj, cudaMemcpyHostToDevice, ctx->stream));

  gnb_graph tmpg;
#define DISPATCH(phase, gp)                                                                              \
  (adj_dtype == GNB_ADJ_F32 ? lower_typed<float>(ctx, (const float*)adj_dev, nn_dev, PN, Badj, B, colcount, coloff, sums, gp, phase) \
   : adj_dtype == GNB_ADJ_U8 ? lower_typed<uint8_t>(ctx, (const uint8_t*)adj_dev, nn_dev, PN, Badj, B, colcount, coloff, sums, gp, phase) \
   : adj_dtype == GNB_ADJ_BITS ? lower_typed<BitAdj>(ctx, (const BitAdj*)adj_dev, nn_dev, PN, Badj, B, colcount, coloff, sums, gp, phase) \
                             : lower_typed<int32_t>(ctx, (const int32_t*)adj_dev, nn_dev, PN, Badj, B, colcount, coloff, sums, gp, phase))
  GNB_TRY(DISPATCH(0, &tmpg));
  int32_t E32 = 0;
  GNB_CUDA(cudaMemcpyAsync(&E32, coloff + total_cols, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  GNB_CUDA(cudaStreamSynchronize(ctx->stream));
  int64_t E = E32;
  GNB_CHECK(E >= 0, "gnb_graph_lower: edge count overflow");

  gnb_graph* g = new gnb_graph();
  g->device = ctx->device;
  g->B = B; g->PN = PN; g->E = E; g->N = N;
  GNB_TRY(graph_alloc(ctx, g));
  int ret = GNB_OK;
  do {
    if (cudaMemcpyAsync(g->graph_node_ptr, gnp.data(), sizeof(int32_t) * (B + 1), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) { ret = GNB_ERR_CUDA; break; }
    // nodes beyond the last column with edges / isolated columns all get a pointer: fill by kernel
    if ((ret = DISPATCH(1, g)) != GNB_OK) break;
    ret = graph_finish(ctx, g);
  } while (0);
#undef DISPATCH
  if (ret != GNB_OK) {
    cudaFreeAsync(g->all, ctx->stream);
    delete g;
    return ret;
  }
  *out = g;
  return GNB_OK;
}

// ---- lowering from COO edge lists (src/batch.jl:53-64 + src/pad.jl:26-46 without the dense detour) ---------------------
namespace {
// first index in [0, n) with a[idx] > v  (a ascending)
__device__ __forceinline__ int upper_bound_i32(const int32_t* __restrict__ a, int n, int v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] <= v) lo = mid + 1; else hi = mid;
  }
  return lo;
}
__global__ void k_coo_fill(const int32_t* __restrict__ src, const int32_t* __restrict__ dst, const int32_t* __restrict__ gep,
                           const int32_t* __restrict__ gnp, int B, int PN, int64_t E, int32_t* __restrict__ edge_src,
                           int32_t* __restrict__ edge_dst, int32_t* __restrict__ edge_slot, int32_t* __restrict__ edge_graph,
                           int* __restrict__ err) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int b = upper_bound_i32(gep, B + 1, (int)e) - 1;      // gep[b] <= e < gep[b+1]
  const int n = gnp[b + 1] - gnp[b];
  const int i = src[e], j = dst[e];
  bool ok = i >= 0 && i < n && j >= 0 && j < n;
  if (ok && e > gep[b]) {
    // ascending padded slot i + PN*j inside a graph == the reference's findall(isone, adj[:]) order; strict: no duplicates
    const int64_t prev = (int64_t)src[e - 1] + (int64_t)PN * dst[e - 1];
    ok = (int64_t)i + (int64_t)PN * j > prev;
  }
  if (!ok) { atomicOr(err, 1); return; }      // error reporting only: not on the data path
  edge_src[e] = gnp[b] + i;
  edge_dst[e] = gnp[b] + j;
  edge_slot[e] = i + PN * j;
  edge_graph[e] = b;
}
// node_in_ptr[v] = first edge whose receiver is >= v (receivers ascend globally); node_graph[v]
__global__ void k_coo_nodes(const int32_t* __restrict__ edge_dst, int64_t E, const int32_t* __restrict__ gnp, int B, int64_t N,
                            int32_t* __restrict__ node_in_ptr, int32_t* __restrict__ node_graph) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v > N) return;
  node_in_ptr[v] = v == 0 ? 0 : upper_bound_i32(edge_dst, (int)E, (int)v - 1);
  if (v < N) node_graph[v] = upper_bound_i32(gnp, B + 1, (int)v) - 1;
}
}  // namespace

extern "C" int gnb_graph_from_coo(gnb_ctx* ctx, const int32_t* src, const int32_t* dst, int coo_on_device,
                                  const int32_t* graph_edge_ptr, const int32_t* n_nodes, int PN, int B, gnb_graph** out) {
  GNB_CHECK(ctx && graph_edge_ptr && n_nodes && out, "gnb_graph_from_coo: null argument");
  GNB_CHECK(B > 0, "gnb_graph_from_coo: need B > 0 (length(adj_mats) > 0, src/checks.jl:8)");
  GNB_CHECK(graph_edge_ptr[0] == 0, "gnb_graph_from_coo: graph_edge_ptr[0] must be 0");
  int64_t N = 0;
  int maxn = 0;
  std::vector<int32_t> gnp(B + 1, 0);
  for (int b = 0; b < B; b++) {
    GNB_CHECK(n_nodes[b] >= 0, "gnb_graph_from_coo: n_nodes[%d] < 0", b);
    GNB_CHECK(graph_edge_ptr[b + 1] >= graph_edge_ptr[b], "gnb_graph_from_coo: graph_edge_ptr must ascend");
    maxn = n_nodes[b] > maxn ? n_nodes[b] : maxn;
    N += n_nodes[b];
    gnp[b + 1] = (int32_t)N;
  }
  if (PN <= 0) PN = maxn > 0 ? maxn : 1;      // padadjmats: common size = largest graph (src/pad.jl:3)
  GNB_CHECK(PN >= maxn, "gnb_graph_from_coo: PN=%d smaller than the largest graph (%d nodes)", PN, maxn);
  GNB_CHECK(N < ((int64_t)1 << 31) && (int64_t)B * PN * PN < ((int64_t)1 << 31), "gnb_graph_from_coo: index range exceeds int32");
  const int64_t E = graph_edge_ptr[B];
  GNB_CHECK(E == 0 || (src && dst), "gnb_graph_from_coo: null edge list");
  GNB_CUDA(cudaSetDevice(ctx->device));
  ctx->arena.reset();
  int rc = GNB_OK;
  const int32_t *d_src = src, *d_dst = dst;
  if (!coo_on_device && E > 0) {
    int32_t* t0 = arena_ptr<int32_t>(ctx->arena, E, &rc);
    int32_t* t1 = arena_ptr<int32_t>(ctx->arena, E, &rc);
    if (rc != GNB_OK) return rc;
    GNB_CUDA(cudaMemcpyAsync(t0, src, sizeof(int32_t) * E, cudaMemcpyHostToDevice, ctx->stream));
    GNB_CUDA(cudaMemcpyAsync(t1, dst, sizeof(int32_t) * E, cudaMemcpyHostToDevice, ctx->stream));
    d_src = t0; d_dst = t1;
  }
  int* d_err = arena_ptr<int>(ctx->arena, 1, &rc);
  if (rc != GNB_OK) return rc;
  gnb_graph* g = new gnb_graph();
  g->device = ctx->device;
  g->B = B; g->PN = PN; g->E = E; g->N = N;
  GNB_TRY(graph_alloc(ctx, g));
  int ret = GNB_OK;
  int h_err = 0;
  do {
    if (cudaMemsetAsync(d_err, 0, sizeof(int), ctx->stream) != cudaSuccess ||
        cudaMemcpyAsync(g->graph_node_ptr, gnp.data(), sizeof(int32_t) * (B + 1), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
        cudaMemcpyAsync(g->graph_edge_ptr, graph_edge_ptr, sizeof(int32_t) * (B + 1), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) { ret = GNB_ERR_CUDA; break; }
    if (E > 0) {
      k_coo_fill<<<ceil_div(E, 256), 256, 0, ctx->stream>>>(d_src, d_dst, g->graph_edge_ptr, g->graph_node_ptr, B, PN, E, g->edge_src,
                                                            g->edge_dst, g->edge_slot, g->edge_graph, d_err);
      ctx->launches++;
    }
    if (cudaMemcpyAsync(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess) { ret = GNB_ERR_CUDA; break; }
    if (h_err) {
      gnb_set_error("gnb_graph_from_coo: every graph's edges must have 0 <= src, dst < n_nodes and be strictly ascending in the "
                    "padded slot src + PN*dst (receiver-major, the order of findall(isone, adj[:]), src/pad.jl:30)");
      ret = GNB_ERR_INVALID;
      break;
    }
    k_coo_nodes<<<ceil_div(N + 1, 256), 256, 0, ctx->stream>>>(g->edge_dst, E, g->graph_node_ptr, B, N, g->node_in_ptr, g->node_graph);
    ctx->launches++;
    ret = graph_finish(ctx, g);
  } while (0);
  if (ret != GNB_OK) {
    cudaFreeAsync(g->all, ctx->stream);
    delete g;
    return ret;
  }
  *out = g;
  return GNB_OK;
}

extern "C" int gnb_graph_destroy(gnb_graph* g) {
  if (!g) return GNB_OK;
  cudaSetDevice(g->device);
  if (g->all) cudaFreeAsync(g->all, g->stream);
  delete g;
  return GNB_OK;
}

extern "C" int gnb_graph_counts(const gnb_graph* g, int64_t* E, int64_t* N, int32_t* B, int32_t* PN) {
  GNB_CHECK(g, "gnb_graph_counts: null graph");
  if (E) *E = g->E;
  if (N) *N = g->N;
  if (B) *B = g->B;
  if (PN) *PN = g->PN;
  return GNB_OK;
}

extern "C" int gnb_graph_export_host(gnb_ctx* ctx, const gnb_graph* g, int32_t* edge_src, int32_t* edge_dst,
                                     int32_t* edge_slot, int32_t* edge_graph, int32_t* graph_edge_ptr,
                                     int32_t* graph_node_ptr, int32_t* node_in_ptr) {
  GNB_CHECK(ctx && g, "gnb_graph_export_host: null argument");
  GNB_CUDA(cudaSetDevice(ctx->device));
  auto cp = [&](int32_t* dst, const int32_t* src, int64_t n) -> cudaError_t {
    if (!dst || n == 0) return cudaSuccess;
    return cudaMemcpyAsync(dst, src, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx->stream);
  };
  GNB_CUDA(cp(edge_src, g->edge_src, g->E));
  GNB_CUDA(cp(edge_dst, g->edge_dst, g->E));
  GNB_CUDA(cp(edge_slot, g->edge_slot, g->E));
  GNB_CUDA(cp(edge_graph, g->edge_graph, g->E));
  GNB_CUDA(cp(graph_edge_ptr, g->graph_edge_ptr, g->B + 1));
  GNB_CUDA(cp(graph_node_ptr, g->graph_node_ptr, g->B + 1));
  GNB_CUDA(cp(node_in_ptr, g->node_in_ptr, g->N + 1));
  GNB_CUDA(cudaStreamSynchronize(ctx->stream));
  return GNB_OK;
}

static int pad_common(gnb_ctx* ctx, const gnb_graph* g, const float* src, int D, float* dst, int edges, int to_padded) {
  GNB_CHECK(ctx && g && D >= 0, "pad/unpad: bad argument");
  if (D == 0) return GNB_OK;
  GNB_CHECK(src && dst, "pad/unpad: null buffer");
  GNB_CUDA(cudaSetDevice(ctx->device));
  int64_t PE = (int64_t)g->PN * g->PN;
  if (to_padded) {
    size_t bytes = sizeof(float) * (size_t)g->B * (edges ? PE : g->PN) * D;
    GNB_CUDA(cudaMemsetAsync(dst, 0, bytes, ctx->stream));
  }
  int64_t rows = edges ? g->E : g->N;
  if (rows > 0) {
    int blocks = ceil_div(rows * D, 256);
    if (edges) k_pad_edges<<<blocks, 256, 0, ctx->stream>>>(src, g->edge_slot, g->edge_graph, g->E, D, PE, dst, to_padded);
    else k_pad_nodes<<<blocks, 256, 0, ctx->stream>>>(src, g->node_graph, g->graph_node_ptr, g->N, D, g->PN, dst, to_padded);
    ctx->launches++;
    GNB_CUDA(cudaGetLastError());
  }
  return GNB_OK;
}

extern "C" int gnb_pad_edges(gnb_ctx* c, const gnb_graph* g, const float* s, int D, float* d) { return pad_common(c, g, s, D, d, 1, 1); }
extern "C" int gnb_unpad_edges(gnb_ctx* c, const gnb_graph* g, const float* s, int D, float* d) { return pad_common(c, g, s, D, d, 1, 0); }
extern "C" int gnb_pad_nodes(gnb_ctx* c, const gnb_graph* g, const float* s, int D, float* d) { return pad_common(c, g, s, D, d, 0, 1); }
extern "C" int gnb_unpad_nodes(gnb_ctx* c, const gnb_graph* g, const float* s, int D, float* d) { return pad_common(c, g, s, D, d, 0, 0); }

extern "C" int gnb_collapse_edges(gnb_ctx* ctx, const gnb_graph* g, const float* ef_padded, int D, float* out) {
  GNB_CHECK(ctx && g && D >= 0, "gnb_collapse_edges: bad argument");
  if (D == 0) return GNB_OK;
  GNB_CHECK(ef_padded && out, "gnb_collapse_edges: null buffer");
  GNB_CUDA(cudaSetDevice(ctx->device));
  int64_t total = (int64_t)g->B * g->PN * g->PN * D;
  k_collapse<<<ceil_div(total, 256), 256, 0, ctx->stream>>>(ef_padded, g->B, g->PN, D, out);
  ctx->launches++;
  GNB_CUDA(cudaGetLastError());
  return GNB_OK;
}
