// Context, model (weights) and forward orchestration of the GNBlock / GNCore path.
// Reference: src/gnblock.jl:63-69, src/gncore.jl:56-68, src/gncorelist.jl:43-45,
// src/gnfeedforward.jl:27-40, src/gngraphnorm.jl:19-26.
#include "kernels.cuh"
#include "tc.cuh"
#include "smallk.cuh"
#include "tc_gemm.cuh"
#include <atomic>
#include <stdlib.h>
#include <mutex>

// ------------------------------------------------------------------ errors / arena / ctx
static thread_local char g_err[1024] = "";

void gnb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char* gnb_last_error(void) { return g_err; }
extern "C" int gnb_version(void) { return 100; }

int Arena::alloc(size_t bytes, void** out) {
  bytes = (bytes + 255) / 256 * 256;
  if (bytes == 0) bytes = 256;
  for (auto& c : chunks) {
    // only the most recent chunks with space are tried in order; chunks are never interleaved
    if (c.cap - c.used >= bytes) {
      *out = c.base + c.used;
      c.used += bytes;
      return GNB_OK;
    }
  }
  size_t cap = bytes > min_chunk ? bytes : min_chunk;
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, cap);
  if (e != cudaSuccess) {
    cudaGetLastError();
    gnb_set_error("workspace cudaMalloc(%zu bytes) failed: %s", cap, cudaGetErrorString(e));
    return GNB_ERR_OOM;
  }
  chunks.push_back({(char*)p, cap, bytes});
  gen++;
  *out = p;
  return GNB_OK;
}
void Arena::reset() {
  // Coalesce: if the previous use spilled over several chunks, replace them by one big chunk so
  // that steady-state forwards bump-allocate out of a single region.
  if (chunks.size() > 1) {
    size_t total = 0;
    for (auto& c : chunks) total += c.cap;
    void* p = nullptr;
    for (auto& c : chunks) cudaFree(c.base);
    chunks.clear();
    gen++;
    if (cudaMalloc(&p, total) == cudaSuccess) chunks.push_back({(char*)p, total, 0});
    else cudaGetLastError();
  }
  for (auto& c : chunks) c.used = 0;
}
void Arena::release() {
  for (auto& c : chunks) cudaFree(c.base);
  gen++;
  chunks.clear();
}

// Forwards of DIFFERENT contexts on the same device are ordered on the device (stream-ordered event chain, the host does not
// block): every persistent tcgen05 kernel is sized to own all SMs, so two forwards interleaved kernel by kernel only take
// turns at the SMs and evict each other's L2 working set.  This is a THROUGHPUT choice, not a correctness requirement: the
// round-1 dead-lock of interleaved forwards was a barrier-protocol bug of k_tc_proj (fixed, see tc.cu) and
// tests/test_zz_gpu_two_contexts.py runs the interleaved mode (GNB_CHAIN_FORWARDS=0) against the oracle-checked result.
// Copies, the lowering and the host-side work of one context still overlap the forward of the other - which is what a
// double-buffered input pipeline needs.
static std::mutex g_chain_mu;
// live contexts of the process: gnb_model_destroy drops a model's per-context state (packed-weight cache entries, captured
// forward graphs) from every one of them
static std::mutex g_ctx_mu;
static std::vector<gnb_ctx*> g_ctxs;
static cudaEvent_t g_chain_ev[64] = {};
static gnb_ctx* g_chain_ctx[64] = {};

extern "C" gnb_ctx* gnb_ctx_create(int device, int* err) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || device < 0 || device >= n) {
    gnb_set_error("gnb_ctx_create: no CUDA device %d (%s)", device,
                  e != cudaSuccess ? cudaGetErrorString(e) : "index out of range");
    if (err) *err = GNB_ERR_CUDA;
    return nullptr;
  }
  cudaSetDevice(device);
  gnb_ctx* c = new gnb_ctx();
  c->device = device;
  cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
  if (cudaMalloc((void**)&c->d_abort, 256) != cudaSuccess || cudaMemset(c->d_abort, 0, 256) != cudaSuccess ||
      cudaHostAlloc((void**)&c->h_abort, sizeof(int), cudaHostAllocDefault) != cudaSuccess) {
    gnb_set_error("gnb_ctx_create: watchdog flag allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
    if (c->d_abort) cudaFree(c->d_abort);
    delete c;
    if (err) *err = GNB_ERR_CUDA;
    return nullptr;
  }
  *c->h_abort = 0;
  if (const char* e = getenv("GNB_WATCHDOG_MS")) {
    const long long ms = atoll(e);
    if (ms > 0) c->wd_limit_ns = (unsigned long long)ms * 1000000ull;
  }
  if (const char* e = getenv("GNB_DEBUG_PROJ_DRAIN_DELAY_NS")) c->dbg_proj_drain_delay_ns = (unsigned int)atoi(e);
  if (err) *err = GNB_OK;
  { std::lock_guard<std::mutex> lk(g_ctx_mu); g_ctxs.push_back(c); }
  return c;
}
extern "C" int gnb_ctx_destroy(gnb_ctx* c) {
  if (!c) return GNB_OK;
  {
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    for (size_t i = 0; i < g_ctxs.size(); i++) if (g_ctxs[i] == c) { g_ctxs.erase(g_ctxs.begin() + i); break; }
  }
  cudaSetDevice(c->device);
  c->arena.release();
  c->staging.release();
  {
    std::lock_guard<std::mutex> lk(g_chain_mu);
    if (c->device < 64 && g_chain_ctx[c->device] == c) { g_chain_ctx[c->device] = nullptr; g_chain_ev[c->device] = nullptr; }
  }
  if (c->done_ev) cudaEventDestroy(c->done_ev);
  if (c->d_abort) cudaFree(c->d_abort);
  if (c->h_abort) cudaFreeHost(c->h_abort);
  for (auto& fg : c->fwd_graphs) if (fg.exec) cudaGraphExecDestroy(fg.exec);
  if (c->train_ws) cudaFree(c->train_ws);
  if (c->pack_ws) cudaFree(c->pack_ws);
  if (c->gstream) cudaStreamDestroy(c->gstream);
  if (c->g_fork) cudaEventDestroy(c->g_fork);
  if (c->g_join) cudaEventDestroy(c->g_join);
  tc_lin_cache_free(c->lin_cache);
  if (c->pipe.copy) {
    cudaStreamDestroy(c->pipe.copy);
    for (auto e : c->pipe.ev_chunk) if (e) cudaEventDestroy(e);
    if (c->pipe.ev_ready) cudaEventDestroy(c->pipe.ev_ready);
    if (c->pipe.ev_in) cudaEventDestroy(c->pipe.ev_in);
  }
  for (auto& r : c->prof_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  for (auto e : c->ev_pool) cudaEventDestroy(e);
  delete c;
  return GNB_OK;
}
extern "C" int gnb_ctx_set_stream(gnb_ctx* c, void* s) {
  GNB_CHECK(c, "gnb_ctx_set_stream: null ctx");
  c->stream = (cudaStream_t)s;
  return GNB_OK;
}
// Kernel watchdog (tc_ptx.cuh): the flag is mirrored to pinned host memory behind every forward; a raised flag turns the
// next synchronising call of this context into GNB_ERR_TIMEOUT and is then cleared (the context stays usable).
int ctx_check_watchdog(gnb_ctx* c) {
  if (!c->h_abort || *c->h_abort == 0) return GNB_OK;
  *c->h_abort = 0;
  cudaMemsetAsync(c->d_abort, 0, sizeof(int), c->stream);
  cudaStreamSynchronize(c->stream);
  gnb_set_error("kernel watchdog: a barrier wait inside a tcgen05 kernel exceeded %llu ms (protocol dead-lock); the kernel "
                "was drained, the results of that forward are invalid", c->wd_limit_ns / 1000000ull);
  return GNB_ERR_TIMEOUT;
}
extern "C" int gnb_sync(gnb_ctx* c) {
  GNB_CHECK(c, "gnb_sync: null ctx");
  GNB_CUDA(cudaSetDevice(c->device));
  GNB_CUDA(cudaStreamSynchronize(c->stream));
  return ctx_check_watchdog(c);
}
extern "C" int64_t gnb_ctx_launch_count(const gnb_ctx* c) { return c ? c->launches : 0; }

Launch::Launch(gnb_ctx* ctx, const char* name, double alg_bytes, double alg_flops) : c(ctx) {
  c->launches++;
  if (!c->profiling) return;
  int tag = -1;
  for (size_t i = 0; i < c->prof_tags.size(); i++)
    if (c->prof_tags[i].name == name || strcmp(c->prof_tags[i].name, name) == 0) { tag = (int)i; break; }
  if (tag < 0) { c->prof_tags.push_back({name, 0, 0, 0, 0}); tag = (int)c->prof_tags.size() - 1; }
  ProfRec r;
  r.tag = tag; r.bytes = alg_bytes; r.flops = alg_flops;
  auto get = [&]() {
    cudaEvent_t e;
    if (!c->ev_pool.empty()) { e = c->ev_pool.back(); c->ev_pool.pop_back(); }
    else cudaEventCreate(&e);
    return e;
  };
  r.a = get(); r.b = get();
  cudaEventRecord(r.a, c->stream);
  c->prof_recs.push_back(r);
  rec = (int)c->prof_recs.size() - 1;
}
Launch::~Launch() {
  if (rec >= 0) cudaEventRecord(c->prof_recs[rec].b, c->stream);
}

extern "C" int gnb_ctx_set_profiling(gnb_ctx* c, int on) {
  GNB_CHECK(c, "gnb_ctx_set_profiling: null ctx");
  c->profiling = on != 0;
  return GNB_OK;
}
extern "C" int gnb_ctx_profile_read(gnb_ctx* c, gnb_prof_entry* out, int cap, int* n) {
  GNB_CHECK(c && n, "gnb_ctx_profile_read: null argument");
  GNB_CUDA(cudaSetDevice(c->device));
  GNB_CUDA(cudaStreamSynchronize(c->stream));
  for (auto& r : c->prof_recs) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.a, r.b);
    ProfTag& t = c->prof_tags[r.tag];
    t.launches++; t.ms += ms; t.bytes += r.bytes; t.flops += r.flops;
    c->ev_pool.push_back(r.a); c->ev_pool.push_back(r.b);
  }
  c->prof_recs.clear();
  int k = 0;
  for (auto& t : c->prof_tags) {
    if (t.launches == 0) continue;
    if (out && k < cap) {
      memset(&out[k], 0, sizeof(gnb_prof_entry));
      strncpy(out[k].name, t.name, sizeof(out[k].name) - 1);
      out[k].launches = t.launches; out[k].ms = t.ms; out[k].alg_bytes = t.bytes; out[k].alg_flops = t.flops;
    }
    k++;
    t.launches = 0; t.ms = t.bytes = t.flops = 0;
  }
  *n = k;
  return GNB_OK;
}

// ------------------------------------------------------------------ model
struct LayerW {
  int kind;
  gnb_block_params blk;
  gnb_ffn_params ffn[3];
  gnb_ln_params ln1[3], ln2[3];
  TcCorePack* tc = nullptr;  // packed bf16 weights (cores the tensor path supports)
  float* wnx = nullptr;      // narrow-input blocks: [We ; be] . Wn_agg  ((in_e+2in_n+in_g+1) x out_n), see run_block_wide
  float* decW4 = nullptr;    // narrow-output blocks behind a tensor-path core: rows [0, in_e) of the edge Dense padded to 4 outputs
};

static std::atomic<uint64_t> g_model_ids{1};



struct gnb_model {
  int device = 0;
  uint64_t id = g_model_ids.fetch_add(1);      // key of the per-context packed-weight cache (tc_gemm.cu)
  std::vector<LayerW> layers;
  float* wbuf = nullptr;  // owned device copy of every weight (nullptr: weights by reference)
  int in_dims[3] = {0, 0, 0};
  int out_dims[3] = {0, 0, 0};
};

// out[k][n] = sum_j X[k][j] Wn[j][n],  X = [We ; be]  ((ke+1) x p),  Wn_agg = rows [0,p) of the node Dense (p x q)
__global__ void k_fold_wnx(const float* __restrict__ We, const float* __restrict__ be, const float* __restrict__ Wn,
                           int ke, int p, int q, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (ke + 1) * q) return;
  const int k = i / q, n = i % q;
  const float* x = k < ke ? We + (size_t)k * p : be;
  float s = 0.f;
  for (int j = 0; j < p; j++) s = fmaf(x[j], Wn[(size_t)j * q + n], s);
  out[i] = s;
}

// out[k][j] = j < p ? We[k][j] : 0     (k < K; We k-major with leading dimension p)
__global__ void k_pad_w4(const float* __restrict__ We, int K, int p, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K * 4) return;
  const int k = i >> 2, j = i & 3;
  out[i] = j < p ? We[(size_t)k * p + j] : 0.f;
}

static int check_block(const gnb_block_params& b, int li) {
  GNB_CHECK(b.in_e >= 0 && b.in_n >= 0 && b.in_g >= 0 && b.out_e >= 0 && b.out_n >= 0 && b.out_g >= 0,
            "layer %d: negative dimension", li);
  GNB_CHECK(b.in_e > 0 || b.in_n > 0 || b.in_g > 0, "layer %d: GNBlock needs any(in .> 0) (src/gnblock.jl:48)", li);
  GNB_CHECK(b.out_e > 0 || b.out_n > 0 || b.out_g > 0, "layer %d: GNBlock needs any(out .> 0) (src/gnblock.jl:49)", li);
  int ke = b.in_e + 2 * b.in_n + b.in_g, kn = b.out_e + b.in_n + b.in_g, kg = b.out_e + b.out_n + b.in_g;
  GNB_CHECK(b.out_e == 0 || (b.be && (ke == 0 || b.We)), "layer %d: edgefn weights missing", li);
  GNB_CHECK(b.out_n == 0 || (b.bn && (kn == 0 || b.Wn)), "layer %d: nodefn weights missing", li);
  GNB_CHECK(b.out_g == 0 || (b.bg && (kg == 0 || b.Wg)), "layer %d: graphfn weights missing", li);
  return GNB_OK;
}

static int build_model(gnb_ctx* ctx, const gnb_layer* layers, int n_layers, int mode /*0 host copy,1 dev copy,2 by reference*/,
                       gnb_model** out) {
  GNB_CHECK(ctx && layers && out && n_layers > 0, "gnb_model_create: bad argument");
  GNB_CUDA(cudaSetDevice(ctx->device));
  gnb_model* m = new gnb_model();
  m->device = ctx->device;
  int cur[3] = {-1, -1, -1};
  size_t total = 0;
  auto cnt = [&](size_t n) { total += (n + 63) / 64 * 64; };
  int rc = GNB_OK;
  for (int li = 0; li < n_layers && rc == GNB_OK; li++) {
    const gnb_layer& L = layers[li];
    LayerW w;
    w.kind = L.kind;
    if (L.kind == GNB_LAYER_BLOCK) {
      w.blk = L.block;
    } else if (L.kind == GNB_LAYER_CORE) {
      w.blk = L.core.block;
      for (int i = 0; i < 3; i++) { w.ffn[i] = L.core.ffn[i]; w.ln1[i] = L.core.ln1[i]; w.ln2[i] = L.core.ln2[i]; }
    } else {
      gnb_set_error("layer %d: unknown kind %d", li, L.kind);
      rc = GNB_ERR_INVALID;
      break;
    }
    if ((rc = check_block(w.blk, li)) != GNB_OK) break;
    const gnb_block_params& b = w.blk;
    if (L.kind == GNB_LAYER_CORE) {
      // GNFeedForward / GNGraphNorm assert all(dims .> 0) (src/gnfeedforward.jl:18, src/gngraphnorm.jl:10)
      if (!(b.in_e > 0 && b.in_n > 0 && b.in_g > 0 && b.in_e == b.out_e && b.in_n == b.out_n && b.in_g == b.out_g)) {
        gnb_set_error("layer %d: GNCore needs all(dims .> 0) and in == out", li);
        rc = GNB_ERR_INVALID;
        break;
      }
      for (int i = 0; i < 3; i++) {
        if (!(w.ffn[i].W1 && w.ffn[i].b1 && w.ffn[i].W2 && w.ffn[i].b2 && w.ln1[i].gamma && w.ln1[i].beta &&
              w.ln2[i].gamma && w.ln2[i].beta)) {
          gnb_set_error("layer %d: GNCore parameter pointer missing (kind %d)", li, i);
          rc = GNB_ERR_INVALID;
          break;
        }
      }
      if (rc != GNB_OK) break;
    }
    if (cur[0] >= 0 && (cur[0] != b.in_e || cur[1] != b.in_n || cur[2] != b.in_g)) {
      gnb_set_error("layer %d: input dims (%d,%d,%d) do not match previous output (%d,%d,%d)", li, b.in_e, b.in_n,
                    b.in_g, cur[0], cur[1], cur[2]);
      rc = GNB_ERR_INVALID;
      break;
    }
    if (li == 0) { m->in_dims[0] = b.in_e; m->in_dims[1] = b.in_n; m->in_dims[2] = b.in_g; }
    cur[0] = b.out_e; cur[1] = b.out_n; cur[2] = b.out_g;
    m->layers.push_back(w);
  }
  if (rc != GNB_OK) { delete m; return rc; }
  m->out_dims[0] = cur[0]; m->out_dims[1] = cur[1]; m->out_dims[2] = cur[2];

  if (mode != 2) {
    // size pass
    for (auto& w : m->layers) {
      const gnb_block_params& b = w.blk;
      cnt((size_t)b.out_e * (b.in_e + 2 * b.in_n + b.in_g)); cnt(b.out_e);
      cnt((size_t)b.out_n * (b.out_e + b.in_n + b.in_g)); cnt(b.out_n);
      cnt((size_t)b.out_g * (b.out_e + b.out_n + b.in_g)); cnt(b.out_g);
      if (w.kind == GNB_LAYER_CORE) {
        int d[3] = {b.in_e, b.in_n, b.in_g};
        for (int i = 0; i < 3; i++) { cnt((size_t)4 * d[i] * d[i]); cnt(4 * d[i]); cnt((size_t)4 * d[i] * d[i]); cnt(d[i]); cnt(d[i]); cnt(d[i]); cnt(d[i]); cnt(d[i]); }
      }
    }
    cudaError_t e = cudaMalloc((void**)&m->wbuf, (total ? total : 64) * sizeof(float));
    if (e != cudaSuccess) {
      cudaGetLastError();
      delete m;
      gnb_set_error("gnb_model_create: cudaMalloc failed: %s", cudaGetErrorString(e));
      return GNB_ERR_OOM;
    }
    size_t off = 0;
    cudaError_t ce = cudaSuccess;
    auto put = [&](const float*& p, size_t n) {
      if (n == 0 || p == nullptr) { p = nullptr; return; }
      float* dst = m->wbuf + off;
      off += (n + 63) / 64 * 64;
      cudaError_t r = cudaMemcpyAsync(dst, p, n * sizeof(float), mode == 0 ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, ctx->stream);
      if (r != cudaSuccess && ce == cudaSuccess) ce = r;
      p = dst;
    };
    for (auto& w : m->layers) {
      gnb_block_params& b = w.blk;
      put(b.We, (size_t)b.out_e * (b.in_e + 2 * b.in_n + b.in_g)); put(b.be, b.out_e);
      put(b.Wn, (size_t)b.out_n * (b.out_e + b.in_n + b.in_g)); put(b.bn, b.out_n);
      put(b.Wg, (size_t)b.out_g * (b.out_e + b.out_n + b.in_g)); put(b.bg, b.out_g);
      if (w.kind == GNB_LAYER_CORE) {
        int d[3] = {b.in_e, b.in_n, b.in_g};
        for (int i = 0; i < 3; i++) {
          put(w.ffn[i].W1, (size_t)4 * d[i] * d[i]); put(w.ffn[i].b1, 4 * d[i]);
          put(w.ffn[i].W2, (size_t)4 * d[i] * d[i]); put(w.ffn[i].b2, d[i]);
          put(w.ln1[i].gamma, d[i]); put(w.ln1[i].beta, d[i]);
          put(w.ln2[i].gamma, d[i]); put(w.ln2[i].beta, d[i]);
        }
      }
    }
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream);
    if (ce != cudaSuccess) {
      gnb_set_error("gnb_model_create: weight upload failed: %s", cudaGetErrorString(ce));
      cudaFree(m->wbuf);
      delete m;
      return GNB_ERR_CUDA;
    }
    // narrow-input blocks (encoders): the node update of the aggregated edge output is folded onto the
    // aggregated narrow inputs,  Wn_agg (We Z + deg be) = ([We ; be] Wn_agg) [Z ; deg]
    for (auto& w : m->layers) {
      const gnb_block_params& b = w.blk;
      const int ke = b.in_e + 2 * b.in_n + b.in_g;
      if (w.kind == GNB_LAYER_BLOCK && b.out_e > 0 && b.out_n > 0 && ke > 0 && ke + 1 + b.in_n + b.in_g <= 32 && b.out_e >= 32) {
        if (cudaMalloc((void**)&w.wnx, (size_t)(ke + 1) * b.out_n * sizeof(float)) != cudaSuccess) {
          cudaGetLastError();
          gnb_model_destroy(m);
          gnb_set_error("gnb_model_create: cudaMalloc failed");
          return GNB_ERR_OOM;
        }
        k_fold_wnx<<<ceil_div((int64_t)(ke + 1) * b.out_n, 128), 128, 0, ctx->stream>>>(b.We, b.be, b.Wn, ke, b.out_e, b.out_n, w.wnx);
      }
    }
    GNB_CUDA(cudaStreamSynchronize(ctx->stream));
    // tensor-core packs for the cores the tcgen05 path supports
    for (auto& w : m->layers) {
      if (w.kind == GNB_LAYER_CORE && tc_core_supported(w.blk.in_e, w.blk.in_n, w.blk.in_g)) {
        int r = tc_core_pack(ctx, w.blk, w.ffn, w.ln1, w.ln2, &w.tc);
        if (r != GNB_OK) {
          gnb_model_destroy(m);
          return r;
        }
      }
    }
    // a narrow-output block right behind a tensor-path core can be fused into that core's edge kernel (dec_fusable below)
    for (size_t li = 1; li < m->layers.size(); li++) {
      LayerW& w = m->layers[li];
      const LayerW& prev = m->layers[li - 1];
      if (w.kind == GNB_LAYER_BLOCK && prev.kind == GNB_LAYER_CORE && prev.tc && w.blk.in_e == 128 && w.blk.out_e > 0 && w.blk.out_e <= 4) {
        if (cudaMalloc((void**)&w.decW4, 128 * 4 * sizeof(float)) != cudaSuccess) {
          cudaGetLastError();
          gnb_model_destroy(m);
          gnb_set_error("gnb_model_create: cudaMalloc failed");
          return GNB_ERR_OOM;
        }
        k_pad_w4<<<2, 256, 0, ctx->stream>>>(w.blk.We, 128, w.blk.out_e, w.decW4);
      }
    }
    GNB_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  *out = m;
  return GNB_OK;
}

extern "C" int gnb_model_create(gnb_ctx* ctx, const gnb_layer* layers, int n_layers, int weights_on_device, gnb_model** out) {
  return build_model(ctx, layers, n_layers, weights_on_device ? 1 : 0, out);
}
extern "C" int gnb_model_destroy(gnb_model* m) {
  if (!m) return GNB_OK;
  cudaSetDevice(m->device);
  for (auto& w : m->layers) {
    if (w.tc) tc_core_pack_free(w.tc);
    if (w.wnx) cudaFree(w.wnx);
    if (w.decW4) cudaFree(w.decW4);
  }
  if (m->wbuf) cudaFree(m->wbuf);
  {
    // per-context state keyed by this model (the caller must not destroy a model while a forward with it is being enqueued)
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    for (gnb_ctx* c : g_ctxs) {
      if (c->device != m->device) continue;
      tc_lin_cache_evict(c->lin_cache, m->id);
      for (size_t i = 0; i < c->fwd_graphs.size();) {
        if (c->fwd_graphs[i].model_id == m->id) {
          if (c->fwd_graphs[i].exec) cudaGraphExecDestroy(c->fwd_graphs[i].exec);
          c->fwd_graphs.erase(c->fwd_graphs.begin() + i);
        } else i++;
      }
    }
  }
  delete m;
  return GNB_OK;
}
extern "C" int gnb_model_out_dims(const gnb_model* m, int32_t* e, int32_t* n, int32_t* g) {
  GNB_CHECK(m, "gnb_model_out_dims: null model");
  if (e) *e = m->out_dims[0];
  if (n) *n = m->out_dims[1];
  if (g) *g = m->out_dims[2];
  return GNB_OK;
}

// ------------------------------------------------------------------ fp32 forward
// (internal linkage via static where not shared with tc.cu)

struct Feat { const float* e; const float* n; const float* g; };
struct FeatOut { float* e; float* n; float* g; };

int launch_linear(gnb_ctx* ctx, const LinArgs& a) {
  if (ctx->use_tc_lin && tc_lin_supported(a)) return launch_linear_tc(ctx, a);
  return launch_linear_fp32(ctx, a);
}

LinSrc mk_src(const float* x, int d, const float* W, const gnb_ln_params* ln) {
  LinSrc s{};
  s.x = x; s.d = d; s.ldx = d; s.W = W;
  s.gamma = ln ? ln->gamma : nullptr;
  s.beta = ln ? ln->beta : nullptr;
  s.eps = ln ? ln->eps : 0.f;
  s.eps_mode = ln ? ln->eps_mode : 0;
  return s;
}

// GNBlock forward (optionally with LayerNorm `ln[3]` applied to the inputs = block(gn1(x))).
// Writes h_e [E][p], h_v [N][q], h_u [B][r].
static int run_block_fp32(gnb_ctx* ctx, const gnb_graph* g, const gnb_block_params& b, const gnb_ln_params* ln,
                   Feat x, FeatOut h) {
  const int a = b.in_e, bn_ = b.in_n, c = b.in_g, p = b.out_e, q = b.out_n, r = b.out_g;
  const int64_t E = g->E, N = g->N, B = g->B;
  const gnb_ln_params* lne = ln ? &ln[0] : nullptr;
  const gnb_ln_params* lnn = ln ? &ln[1] : nullptr;
  const gnb_ln_params* lng = ln ? &ln[2] : nullptr;
  int rc = GNB_OK;
  float* agg = nullptr;
  if (p > 0) {
    float *Ps = nullptr, *Pr = nullptr, *Pu = nullptr;
    if (bn_ > 0) {
      Ps = arena_ptr<float>(ctx->arena, (size_t)N * p, &rc);
      Pr = arena_ptr<float>(ctx->arena, (size_t)N * p, &rc);
      if (rc != GNB_OK) return rc;
      LinArgs la{};
      la.R = N; la.Nout = p; la.ldw = p; la.nsrc = 1; la.ldo = p;
      la.src[0] = mk_src(x.n, bn_, b.We + (size_t)a * p, lnn);
      la.out = Ps;
      GNB_TRY(launch_linear(ctx, la));
      la.src[0] = mk_src(x.n, bn_, b.We + (size_t)(a + bn_) * p, lnn);
      la.out = Pr;
      if (!(ctx->use_tc_lin && c > 0)) GNB_TRY(launch_linear(ctx, la));      // else: launched below, with the per-graph row folded in
    }
    bool pu_folded = false;
    if (c > 0) {
      Pu = arena_ptr<float>(ctx->arena, (size_t)B * p, &rc);
      if (rc != GNB_OK) return rc;
      LinArgs la{};
      la.R = B; la.Nout = p; la.ldw = p; la.nsrc = 1; la.ldo = p;
      la.src[0] = mk_src(x.g, c, b.We + (size_t)(a + 2 * bn_) * p, lng);
      la.bias = b.be;
      la.out = Pu;
      GNB_TRY(launch_linear(ctx, la));
      if (ctx->use_tc_lin && bn_ > 0) {
        // tensor-core precision modes: P_r'[v] = P_r[v] + P_u[graph(v)] once per node (N rows), so that the per-edge launch
        // gathers two rows instead of three (its epilogue is what bounds it)
        LinArgs lr{};
        lr.R = N; lr.Nout = p; lr.ldw = p; lr.nsrc = 1; lr.ldo = p;
        lr.src[0] = mk_src(x.n, bn_, b.We + (size_t)(a + bn_) * p, lnn);
        lr.add[lr.nadd++] = LinAdd{Pu, g->node_graph, p};
        lr.out = Pr;
        GNB_TRY(launch_linear(ctx, lr));
        pu_folded = true;
      }
    }
    LinArgs la{};
    la.R = E; la.Nout = p; la.ldw = p; la.ldo = p; la.out = h.e;
    if (a > 0) { la.nsrc = 1; la.src[0] = mk_src(x.e, a, b.We, lne); }
    if (Ps) {
      la.add[la.nadd++] = LinAdd{Ps, g->edge_src, p};
      la.add[la.nadd++] = LinAdd{Pr, g->edge_dst, p};
    }
    if (Pu && !pu_folded) la.add[la.nadd++] = LinAdd{Pu, g->edge_graph, p};
    else if (!Pu) la.bias = b.be;
    GNB_TRY(launch_linear(ctx, la));
    // edge -> node aggregation over the receiver CSR (src/nodefninput.jl:3)
    if (q > 0 || r > 0) {
      agg = arena_ptr<float>(ctx->arena, (size_t)N * p, &rc);
      if (rc != GNB_OK) return rc;
      GNB_TRY(launch_segsum(ctx, h.e, p, g->node_in_ptr, N, agg));
    }
  }
  if (q > 0) {
    float* Pu = nullptr;
    if (c > 0) {
      Pu = arena_ptr<float>(ctx->arena, (size_t)B * q, &rc);
      if (rc != GNB_OK) return rc;
      LinArgs la{};
      la.R = B; la.Nout = q; la.ldw = q; la.nsrc = 1; la.ldo = q;
      la.src[0] = mk_src(x.g, c, b.Wn + (size_t)(p + bn_) * q, lng);
      la.bias = b.bn;
      la.out = Pu;
      GNB_TRY(launch_linear(ctx, la));
    }
    LinArgs la{};
    la.R = N; la.Nout = q; la.ldw = q; la.ldo = q; la.out = h.n;
    if (p > 0) la.src[la.nsrc++] = mk_src(agg, p, b.Wn, nullptr);
    if (bn_ > 0) la.src[la.nsrc++] = mk_src(x.n, bn_, b.Wn + (size_t)p * q, lnn);
    if (Pu) la.add[la.nadd++] = LinAdd{Pu, g->node_graph, q};
    else la.bias = b.bn;
    GNB_TRY(launch_linear(ctx, la));
  }
  if (r > 0) {
    float *se = nullptr, *sv = nullptr;
    if (p > 0) {
      // sum of a graph's edges == sum of its nodes' incoming-edge aggregates (src/graphfninput.jl:3)
      se = arena_ptr<float>(ctx->arena, (size_t)B * p, &rc);
      if (rc != GNB_OK) return rc;
      GNB_TRY(launch_segsum(ctx, agg, p, g->graph_node_ptr, B, se));
    }
    if (q > 0) {
      sv = arena_ptr<float>(ctx->arena, (size_t)B * q, &rc);
      if (rc != GNB_OK) return rc;
      GNB_TRY(launch_segsum(ctx, h.n, q, g->graph_node_ptr, B, sv));
    }
    LinArgs la{};
    la.R = B; la.Nout = r; la.ldw = r; la.ldo = r; la.out = h.g; la.bias = b.bg;
    if (p > 0) la.src[la.nsrc++] = mk_src(se, p, b.Wg, nullptr);
    if (q > 0) la.src[la.nsrc++] = mk_src(sv, q, b.Wg + (size_t)p * r, nullptr);
    if (c > 0) la.src[la.nsrc++] = mk_src(x.g, c, b.Wg + (size_t)(p + q) * r, lng);
    GNB_TRY(launch_linear(ctx, la));
  }
  return GNB_OK;
}


// Shared graph-level tail of the block fast paths: h_u = Wg [sum agg ; sum h_v ; u] + bg   (src/graphfninput.jl:2-6)
static int run_graph_update_fp32(gnb_ctx* ctx, const gnb_graph* g, const gnb_block_params& b, const float* agg,
                                 const float* hn, const float* xg, float* hg) {
  const int c = b.in_g, p = b.out_e, q = b.out_n, r = b.out_g;
  const int64_t B = g->B;
  int rc = GNB_OK;
  float *se = nullptr, *sv = nullptr;
  if (p > 0) {
    se = arena_ptr<float>(ctx->arena, (size_t)B * p, &rc);
    if (rc != GNB_OK) return rc;
    GNB_TRY(launch_segsum(ctx, agg, p, g->graph_node_ptr, B, se));
  }
  if (q > 0) {
    sv = arena_ptr<float>(ctx->arena, (size_t)B * q, &rc);
    if (rc != GNB_OK) return rc;
    GNB_TRY(launch_segsum(ctx, hn, q, g->graph_node_ptr, B, sv));
  }
  LinArgs la{};
  la.R = B; la.Nout = r; la.ldw = r; la.ldo = r; la.out = hg; la.bias = b.bg;
  if (p > 0) la.src[la.nsrc++] = mk_src(se, p, b.Wg, nullptr);
  if (q > 0) la.src[la.nsrc++] = mk_src(sv, q, b.Wg + (size_t)p * r, nullptr);
  if (c > 0) la.src[la.nsrc++] = mk_src(xg, c, b.Wg + (size_t)(p + q) * r, nullptr);
  return launch_linear_fp32(ctx, la);
}

// GNBlock with NARROW inputs (encoder): the [e | v_src | v_dst | u] concat is at most 31 wide.
// Edge update straight from the raw inputs (no node projections, no gathered 128-wide rows); the
// edge -> node sum is taken over the narrow inputs and transformed afterwards (linearity of Dense).
static bool block_wide_ok(const LayerW& w) {
  const gnb_block_params& b = w.blk;
  const int ke = b.in_e + 2 * b.in_n + b.in_g;
  return w.kind == GNB_LAYER_BLOCK && b.out_e >= 32 && (b.out_e & 3) == 0 && ke > 0 && ke + 1 <= 32 &&
         (b.out_n == 0 || (w.wnx != nullptr && (b.out_n & 3) == 0));
}
static int run_block_wide(gnb_ctx* ctx, const gnb_graph* g, const LayerW& w, Feat x, FeatOut h) {
  const gnb_block_params& b = w.blk;
  const int a = b.in_e, bn_ = b.in_n, c = b.in_g, p = b.out_e, q = b.out_n, r = b.out_g;
  const int ke = a + 2 * bn_ + c, kz = ke + 1;
  const int64_t E = g->E, N = g->N;
  int rc = GNB_OK;
  {
    WideArgs wa{};
    wa.R = E; wa.Nout = p; wa.ldw = p; wa.bias = b.be; wa.out = h.e; wa.ldo = p;
    if (a > 0) wa.pc[wa.np++] = WidePiece{x.e, nullptr, a, a, b.We};
    if (bn_ > 0) {
      wa.pc[wa.np++] = WidePiece{x.n, g->edge_src, bn_, bn_, b.We + (size_t)a * p};
      wa.pc[wa.np++] = WidePiece{x.n, g->edge_dst, bn_, bn_, b.We + (size_t)(a + bn_) * p};
    }
    if (c > 0) wa.pc[wa.np++] = WidePiece{x.g, g->edge_graph, c, c, b.We + (size_t)(a + 2 * bn_) * p};
    if (ctx->pipe.ef_pending && x.e == ctx->pipe.d_ef && a > 0) {
      // the edge rows are still being uploaded (gnb_model_forward_host): one launch per chunk, each behind its copy
      const int nch = ctx->pipe.nchunks;
      const int64_t per = (E + nch - 1) / nch;
      for (int ci = 0; ci < nch; ci++) {
        const int64_t r0 = ci * per, r1 = r0 + per < E ? r0 + per : E;
        if (r0 >= r1) break;
        GNB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->pipe.ev_chunk[ci], 0));
        WideArgs wc = wa;
        wc.R = r1 - r0; wc.out = h.e + (size_t)r0 * p;
        for (int i = 0; i < wc.np; i++) {
          if (wc.pc[i].idx) wc.pc[i].idx += r0;
          else wc.pc[i].x += (size_t)r0 * wc.pc[i].ldx;
        }
        GNB_TRY(launch_wide(ctx, wc));
      }
      ctx->pipe.ef_pending = false;
    } else {
      GNB_TRY(launch_wide(ctx, wa));
    }
  }
  if (q == 0 && r == 0) return GNB_OK;
  float* Z = arena_ptr<float>(ctx->arena, (size_t)N * kz, &rc);
  if (rc != GNB_OK) return rc;
  {
    ZsumArgs za{};
    za.N = N; za.node_in_ptr = g->node_in_ptr; za.edge_src = g->edge_src; za.node_graph = g->node_graph;
    za.ef = x.e; za.nf = x.n; za.gf = x.g; za.de = a; za.dn = bn_; za.dg = c; za.Z = Z;
    GNB_TRY(launch_zsum(ctx, za));
  }
  float* agg = nullptr;
  if (r > 0) {   // the graph update needs sum_v agg_v
    agg = arena_ptr<float>(ctx->arena, (size_t)N * p, &rc);
    if (rc != GNB_OK) return rc;
    WideArgs wa{};
    wa.R = N; wa.Nout = p; wa.ldw = p; wa.out = agg; wa.ldo = p;
    wa.pc[wa.np++] = WidePiece{Z, nullptr, ke, kz, b.We};
    wa.pc[wa.np++] = WidePiece{Z + ke, nullptr, 1, kz, b.be};
    GNB_TRY(launch_wide(ctx, wa));
  }
  if (q > 0) {
    WideArgs wa{};
    wa.R = N; wa.Nout = q; wa.ldw = q; wa.bias = b.bn; wa.out = h.n; wa.ldo = q;
    wa.pc[wa.np++] = WidePiece{Z, nullptr, kz, kz, w.wnx};
    if (bn_ > 0) wa.pc[wa.np++] = WidePiece{x.n, nullptr, bn_, bn_, b.Wn + (size_t)p * q};
    if (c > 0) wa.pc[wa.np++] = WidePiece{x.g, g->node_graph, c, c, b.Wn + (size_t)(p + bn_) * q};
    GNB_TRY(launch_wide(ctx, wa));
  }
  if (r > 0) GNB_TRY(run_graph_update_fp32(ctx, g, b, agg, h.n, x.g, h.g));
  return GNB_OK;
}

// GNBlock with NARROW outputs (decoder): out_e, out_n <= 8; wide inputs streamed once, coalesced.
static bool block_narrow_ok(const LayerW& w) {
  const gnb_block_params& b = w.blk;
  return w.kind == GNB_LAYER_BLOCK && b.out_e > 0 && b.out_e <= 8 && b.out_n <= 8 && b.in_e >= 32 && b.in_e <= 512 &&
         b.in_n <= 512 && (b.in_e & 3) == 0 && (b.in_n & 3) == 0;
}
// Edge update of a narrow decoder whose edge-feature term was already computed by the previous core's edge kernel
// (partial[e][0..4) = y_e . We[0:128, :]):  h_e = partial + Ps[src] + Pr[dst] + (Pu[graph] | be)      (src/gnblock.jl:65)
__global__ void k_dec_finish(const float* __restrict__ partial, const float* __restrict__ Ps, const float* __restrict__ Pr, int ldP,
                             const float* __restrict__ Pu, const float* __restrict__ be, const int32_t* __restrict__ src,
                             const int32_t* __restrict__ dst, const int32_t* __restrict__ eg, int64_t E, int p, float* __restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const float4 v4 = __ldg(reinterpret_cast<const float4*>(partial) + e);
  float v[4] = {v4.x, v4.y, v4.z, v4.w};
  const int s_ = Ps ? src[e] : 0, d_ = Pr ? dst[e] : 0, g_ = Pu ? eg[e] : 0;
  for (int j = 0; j < p; j++) {
    float t = v[j];
    if (Ps) t += __ldg(Ps + (size_t)s_ * ldP + j) + __ldg(Pr + (size_t)d_ * ldP + j);
    t += Pu ? __ldg(Pu + (size_t)g_ * p + j) : be[j];
    out[(size_t)e * p + j] = t;
  }
}

static bool dec_fusable(const LayerW& core, const LayerW& dec) {
  return core.kind == GNB_LAYER_CORE && core.tc != nullptr && dec.decW4 != nullptr && block_narrow_ok(dec);
}

static int run_block_narrow(gnb_ctx* ctx, const gnb_graph* g, const LayerW& w, Feat x, FeatOut h, const float* edge_partial = nullptr) {
  const gnb_block_params& b = w.blk;
  const int a = b.in_e, bn_ = b.in_n, c = b.in_g, p = b.out_e, q = b.out_n, r = b.out_g;
  const int64_t E = g->E, N = g->N, B = g->B;
  int rc = GNB_OK;
  float *Ps = nullptr, *Pr = nullptr, *Pu = nullptr;
  int ldP = p;      // row stride of P_s / P_r
  if (bn_ > 0) {
    // P_s | P_r interleaved per node ([N][2p]): both projections in ONE pass over the node rows
    Ps = arena_ptr<float>(ctx->arena, (size_t)N * 2 * p, &rc);
    if (rc != GNB_OK) return rc;
    Pr = Ps + p;
    ldP = 2 * p;
    if (2 * p <= 8) {
      NarrowArgs na{};
      na.R = N; na.No = 2 * p; na.ldw = p; na.nsrc = 1; na.ldo = 2 * p;
      na.src[0] = NarrowSrc{x.n, bn_, bn_, b.We + (size_t)a * p, b.We + (size_t)(a + bn_) * p, p};
      na.out = Ps;
      GNB_TRY(launch_narrow(ctx, na));
    } else {
      NarrowArgs na{};
      na.R = N; na.No = p; na.ldw = p; na.nsrc = 1; na.ldo = 2 * p;
      na.src[0] = NarrowSrc{x.n, bn_, bn_, b.We + (size_t)a * p};
      na.out = Ps;
      GNB_TRY(launch_narrow(ctx, na));
      na.src[0].W = b.We + (size_t)(a + bn_) * p;
      na.out = Pr;
      GNB_TRY(launch_narrow(ctx, na));
    }
  }
  if (c > 0) {
    Pu = arena_ptr<float>(ctx->arena, (size_t)B * p, &rc);
    if (rc != GNB_OK) return rc;
    LinArgs la{};
    la.R = B; la.Nout = p; la.ldw = p; la.nsrc = 1; la.ldo = p;
    la.src[0] = mk_src(x.g, c, b.We + (size_t)(a + 2 * bn_) * p, nullptr);
    la.bias = b.be; la.out = Pu;
    GNB_TRY(launch_linear_fp32(ctx, la));
  }
  {
    if (edge_partial) {      // the edge-feature term came out of the previous core's edge kernel
      if (E > 0) {
        Launch L(ctx, "dec_finish", 4.0 * E * (4 + p + 3), 0);
        k_dec_finish<<<ceil_div(E, 256), 256, 0, ctx->stream>>>(edge_partial, Ps, Pr, ldP, Pu, b.be, g->edge_src, g->edge_dst, g->edge_graph, E, p, h.e);
        GNB_CUDA(cudaGetLastError());
      }
    } else {
      NarrowArgs na{};
      na.R = E; na.No = p; na.ldw = p; na.nsrc = 1; na.ldo = p; na.out = h.e;
      na.src[0] = NarrowSrc{x.e, a, a, b.We};
      if (Ps) { na.add[na.nadd++] = NarrowAdd{Ps, g->edge_src, ldP}; na.add[na.nadd++] = NarrowAdd{Pr, g->edge_dst, ldP}; }
      if (Pu) na.add[na.nadd++] = NarrowAdd{Pu, g->edge_graph, p};
      else na.bias = b.be;
      GNB_TRY(launch_narrow(ctx, na));
    }
    if (ctx->pipe.out_pending && h.e == ctx->pipe.d_out_ef) {
      // the edge output is final: download it now, under the node / graph kernels of this block (gnb_model_forward_host)
      GNB_CUDA(cudaEventRecord(ctx->pipe.ev_ready, ctx->stream));
      GNB_CUDA(cudaStreamWaitEvent(ctx->pipe.copy, ctx->pipe.ev_ready, 0));
      GNB_CUDA(cudaMemcpyAsync(ctx->pipe.h_out_ef, h.e, ctx->pipe.out_bytes, cudaMemcpyDeviceToHost, ctx->pipe.copy));
      ctx->pipe.out_pending = false;
      ctx->pipe.out_early = true;
    }
  }
  if (q == 0 && r == 0) return GNB_OK;
  float* agg = arena_ptr<float>(ctx->arena, (size_t)N * p, &rc);
  if (rc != GNB_OK) return rc;
  GNB_TRY(launch_segsum(ctx, h.e, p, g->node_in_ptr, N, agg));
  if (q > 0) {
    float* Pun = nullptr;
    if (c > 0) {
      Pun = arena_ptr<float>(ctx->arena, (size_t)B * q, &rc);
      if (rc != GNB_OK) return rc;
      LinArgs la{};
      la.R = B; la.Nout = q; la.ldw = q; la.nsrc = 1; la.ldo = q;
      la.src[0] = mk_src(x.g, c, b.Wn + (size_t)(p + bn_) * q, nullptr);
      la.bias = b.bn; la.out = Pun;
      GNB_TRY(launch_linear_fp32(ctx, la));
    }
    NarrowArgs na{};
    na.R = N; na.No = q; na.ldw = q; na.ldo = q; na.out = h.n;
    na.src[na.nsrc++] = NarrowSrc{agg, p, p, b.Wn};
    if (bn_ > 0) na.src[na.nsrc++] = NarrowSrc{x.n, bn_, bn_, b.Wn + (size_t)p * q};
    if (Pun) na.add[na.nadd++] = NarrowAdd{Pun, g->node_graph, q};
    else na.bias = b.bn;
    GNB_TRY(launch_narrow(ctx, na));
  }
  if (r > 0) GNB_TRY(run_graph_update_fp32(ctx, g, b, agg, h.n, x.g, h.g));
  return GNB_OK;
}

// y = (x + h) + W2 relu(W1 LN2(x) + b1) + b2       (src/gncore.jl:56-68, src/gnfeedforward.jl:27-31)
int run_ffn_residual_fp32(gnb_ctx* ctx, int64_t R, int d, const gnb_ffn_params& f, const gnb_ln_params& ln2,
                          const float* x, const float* h, float* y) {
  if (R <= 0) return GNB_OK;
  if (ctx->use_tc_lin && tc_ffn256_supported(R, d)) return launch_ffn256_tc(ctx, R, f, ln2, x, h, y, d);      // fused, hidden on chip
  // tensor-core precision modes: both Dense layers on the generic tcgen05 kernel with the hidden activation kept in bf16
  // (exactly what the down-projection's MMA consumes anyway: halves the HBM round trip of the 4d-wide hidden rows)
  bool bf16_hidden = false;
  if (ctx->use_tc_lin) {
    LinArgs t1{}, t2{};
    t1.R = R; t1.Nout = 4 * d; t1.nsrc = 1; t1.ldo = 4 * d; t1.src[0] = mk_src(x, d, f.W1, &ln2);
    t2.R = R; t2.Nout = d; t2.nsrc = 1; t2.ldo = d; t2.src[0] = mk_src(x, 4 * d, f.W2, nullptr); t2.src[0].x_bf16 = 1;
    t2.nadd = 2; t2.add[0] = LinAdd{x, nullptr, d}; t2.add[1] = LinAdd{h, nullptr, d};
    bf16_hidden = tc_lin_supported(t1) && tc_lin_supported(t2);
  }
  const size_t cap_bytes = (size_t)1 << 30;  // hidden scratch per chunk
  const size_t esz = bf16_hidden ? 2 : 4;
  int64_t chunk = (int64_t)(cap_bytes / ((size_t)4 * d * esz));
  if (chunk < 1024) chunk = 1024;
  if (chunk > R) chunk = R;
  int rc = GNB_OK;
  float* hid = reinterpret_cast<float*>(arena_ptr<char>(ctx->arena, (size_t)chunk * 4 * d * esz, &rc));
  if (rc != GNB_OK) return rc;
  for (int64_t r0 = 0; r0 < R; r0 += chunk) {
    int64_t rows = R - r0 < chunk ? R - r0 : chunk;
    // a short last chunk falls back to fp32 hidden rows only if the tensor-core kernel refuses it (R < 256): the buffer is
    // large enough either way when chunk == R, otherwise rows == chunk except for the tail
    LinArgs l1{};
    l1.R = rows; l1.Nout = 4 * d; l1.ldw = 4 * d; l1.nsrc = 1; l1.ldo = 4 * d;
    l1.src[0] = mk_src(x + (size_t)r0 * d, d, f.W1, &ln2);
    l1.bias = f.b1; l1.relu = 1; l1.out = hid;
    LinArgs l2{};
    l2.R = rows; l2.Nout = d; l2.ldw = d; l2.nsrc = 1; l2.ldo = d;
    l2.src[0] = mk_src(hid, 4 * d, f.W2, nullptr);
    l2.bias = f.b2;
    l2.add[l2.nadd++] = LinAdd{x + (size_t)r0 * d, nullptr, d};
    l2.add[l2.nadd++] = LinAdd{h + (size_t)r0 * d, nullptr, d};
    l2.out = y + (size_t)r0 * d;
    bool bf = bf16_hidden;
    if (bf) {
      l1.out_bf16 = 1; l2.src[0].x_bf16 = 1;
      bf = tc_lin_supported(l1) && tc_lin_supported(l2);
      if (!bf) { l1.out_bf16 = 0; l2.src[0].x_bf16 = 0; }      // tail chunk too short for the tensor-core kernel
    }
    if (!bf && bf16_hidden && (size_t)rows * 4 * d * 4 > (size_t)chunk * 4 * d * esz) {
      gnb_set_error("run_ffn_residual: hidden scratch too small for an fp32 tail chunk");
      return GNB_ERR_INVALID;
    }
    GNB_TRY(launch_linear(ctx, l1));
    GNB_TRY(launch_linear(ctx, l2));
  }
  return GNB_OK;
}

static int run_core_fp32(gnb_ctx* ctx, const gnb_graph* g, const LayerW& w, Feat x, FeatOut y) {
  const int d[3] = {w.blk.in_e, w.blk.in_n, w.blk.in_g};
  int rc = GNB_OK;
  FeatOut h;
  h.e = arena_ptr<float>(ctx->arena, (size_t)g->E * d[0], &rc);
  h.n = arena_ptr<float>(ctx->arena, (size_t)g->N * d[1], &rc);
  h.g = arena_ptr<float>(ctx->arena, (size_t)g->B * d[2], &rc);
  if (rc != GNB_OK) return rc;
  GNB_TRY(run_block_fp32(ctx, g, w.blk, w.ln1, x, h));
  GNB_TRY(run_ffn_residual_fp32(ctx, g->E, d[0], w.ffn[0], w.ln2[0], x.e, h.e, y.e));
  GNB_TRY(run_ffn_residual_fp32(ctx, g->N, d[1], w.ffn[1], w.ln2[1], x.n, h.n, y.n));
  GNB_TRY(run_ffn_residual_fp32(ctx, g->B, d[2], w.ffn[2], w.ln2[2], x.g, h.g, y.g));
  return GNB_OK;
}

struct ArenaMark { size_t nchunks; std::vector<size_t> used; };
static ArenaMark arena_mark(Arena& a) {
  ArenaMark m;
  m.nchunks = a.chunks.size();
  for (auto& c : a.chunks) m.used.push_back(c.used);
  return m;
}
static void arena_rewind(Arena& a, const ArenaMark& m) {
  for (size_t i = 0; i < a.chunks.size(); i++) a.chunks[i].used = i < m.nchunks ? m.used[i] : 0;
}

static int forward_device_impl(gnb_ctx* ctx, const gnb_model* m, const gnb_graph* g, const float* ef, const float* nf,
                   const float* gf, float* out_ef, float* out_nf, float* out_gf, int precision, bool reset_arena) {
  GNB_CHECK(ctx && m && g, "gnb_model_forward: null argument");
  GNB_CHECK(ctx->device == m->device && ctx->device == g->device, "gnb_model_forward: ctx/model/graph on different devices");
  GNB_CHECK(precision == GNB_PREC_FP32 || precision == GNB_PREC_BF16 || precision == GNB_PREC_AUTO,
            "gnb_model_forward: unknown precision %d", precision);
  GNB_CHECK(m->in_dims[0] == 0 || g->E == 0 || ef, "gnb_model_forward: ef is NULL but the model expects in_e=%d", m->in_dims[0]);
  GNB_CHECK(m->in_dims[1] == 0 || g->N == 0 || nf, "gnb_model_forward: nf is NULL but the model expects in_n=%d", m->in_dims[1]);
  GNB_CHECK(m->in_dims[2] == 0 || gf, "gnb_model_forward: gf is NULL but the model expects in_g=%d", m->in_dims[2]);
  GNB_CHECK(m->out_dims[0] == 0 || g->E == 0 || out_ef, "gnb_model_forward: out_ef is NULL");
  GNB_CHECK(m->out_dims[1] == 0 || g->N == 0 || out_nf, "gnb_model_forward: out_nf is NULL");
  GNB_CHECK(m->out_dims[2] == 0 || out_gf, "gnb_model_forward: out_gf is NULL");
  GNB_CUDA(cudaSetDevice(ctx->device));
  if (reset_arena) ctx->arena.reset();
  ctx->use_tc_lin = precision != GNB_PREC_FP32 && m->wbuf != nullptr;      // needs model-owned weights (cache key = model id)
  ctx->cur_model_id = m->id;
  const int L = (int)m->layers.size();
  // ping-pong activation buffers sized for the widest intermediate layer output
  size_t me = 0, mn = 0, mg = 0;
  for (int i = 0; i + 1 < L; i++) {
    const gnb_block_params& b = m->layers[i].blk;
    if ((size_t)b.out_e > me) me = b.out_e;
    if ((size_t)b.out_n > mn) mn = b.out_n;
    if ((size_t)b.out_g > mg) mg = b.out_g;
  }
  int rc = GNB_OK;
  FeatOut pp[2] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
  int nbuf = L > 2 ? 2 : (L > 1 ? 1 : 0);
  for (int i = 0; i < nbuf; i++) {
    pp[i].e = arena_ptr<float>(ctx->arena, (size_t)g->E * me, &rc);
    pp[i].n = arena_ptr<float>(ctx->arena, (size_t)g->N * mn, &rc);
    pp[i].g = arena_ptr<float>(ctx->arena, (size_t)g->B * mg, &rc);
  }
  // per-graph rows of consecutive tensor-path cores are chained (produced by the previous core's tail kernel): two
  // buffer pairs that survive the per-layer arena rewind
  TcPreRows pre[2];
  bool chain = false;
  for (int i = 0; i + 1 < L; i++)
    chain |= precision != GNB_PREC_FP32 && m->layers[i].kind == GNB_LAYER_CORE && m->layers[i].tc && m->layers[i + 1].kind == GNB_LAYER_CORE && m->layers[i + 1].tc;
  if (chain) {
    for (int i = 0; i < 2; i++) {
      pre[i].Pue = arena_ptr<float>(ctx->arena, (size_t)g->B * 128, &rc);
      pre[i].Pun = arena_ptr<float>(ctx->arena, (size_t)g->B * 128, &rc);
    }
  }
  // fused narrow decoder: the edge kernel of a tensor-path core hands y_e . W_dec to the narrow block behind it
  const char* fuse_env = getenv("GNB_FUSE_DECODER");
  const bool fuse_dec = precision != GNB_PREC_FP32 && !(fuse_env && atoi(fuse_env) == 0);
  float* dec_partial = nullptr;
  if (fuse_dec) {
    for (int i = 0; i + 1 < L && !dec_partial; i++)
      if (dec_fusable(m->layers[i], m->layers[i + 1])) dec_partial = arena_ptr<float>(ctx->arena, (size_t)g->E * 4, &rc);
  }
  if (rc != GNB_OK) return rc;
  bool have_partial = false;      // dec_partial holds the edge term of layer li
  Feat x{ef, nf, gf};
  if (ctx->pipe.ef_pending && !(precision != GNB_PREC_FP32 && block_wide_ok(m->layers[0]) && m->layers[0].blk.in_e > 0)) {
    // chunked upload in flight but the first layer reads all edge rows at once: wait for the last chunk
    for (int ci = 0; ci < ctx->pipe.nchunks; ci++) GNB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->pipe.ev_chunk[ci], 0));
    ctx->pipe.ef_pending = false;
  }
  bool have_pre = false;      // pre[li & 1] holds the rows of layer li
  ArenaMark mark = arena_mark(ctx->arena);
  for (int li = 0; li < L; li++) {
    const LayerW& w = m->layers[li];
    FeatOut y = (li == L - 1) ? FeatOut{out_ef, out_nf, out_gf} : pp[li & 1];
    if (w.blk.in_e == 0) x.e = nullptr;
    if (w.blk.in_n == 0) x.n = nullptr;
    if (w.blk.in_g == 0) x.g = nullptr;
    arena_rewind(ctx->arena, mark);
    if (w.kind == GNB_LAYER_BLOCK) {
      // GNB_PREC_FP32 keeps the reference's operation order (generic path); the other modes may take the
      // algebraically equivalent streaming paths for narrow-input / narrow-output blocks
      if (precision != GNB_PREC_FP32 && block_wide_ok(w)) GNB_TRY(run_block_wide(ctx, g, w, x, y));
      else if (precision != GNB_PREC_FP32 && block_narrow_ok(w)) GNB_TRY(run_block_narrow(ctx, g, w, x, y, have_partial ? dec_partial : nullptr));
      else GNB_TRY(run_block_fp32(ctx, g, w.blk, nullptr, x, y));
      have_pre = false;
      have_partial = false;
    } else {
      bool use_tc = (precision != GNB_PREC_FP32) && w.tc != nullptr;
      const bool wide_ok = (w.blk.in_e % 128 == 0) && (w.blk.in_n % 128 == 0) && (w.blk.in_g % 128 == 0) && m->wbuf != nullptr;
      if (precision == GNB_PREC_BF16 && !w.tc && !wide_ok) {      // wide_ok: the generic tcgen05 linear layers (tc_gemm.cu) apply
        gnb_set_error("layer %d: GNCore dims (%d,%d,%d) are not supported by the tcgen05 bf16 path "
                      "(use GNB_PREC_AUTO or GNB_PREC_FP32)", li, w.blk.in_e, w.blk.in_n, w.blk.in_g);
        return GNB_ERR_UNSUPPORTED;
      }
      if (use_tc) {
        TcNextCore next;
        if (chain && li + 1 < L && m->layers[li + 1].kind == GNB_LAYER_CORE && m->layers[li + 1].tc) {
          next.pk = m->layers[li + 1].tc; next.blk = &m->layers[li + 1].blk; next.ln1 = m->layers[li + 1].ln1;
        }
        TcDecFuse dec;
        if (dec_partial && li + 1 < L && dec_fusable(w, m->layers[li + 1])) { dec.W4 = m->layers[li + 1].decW4; dec.partial = dec_partial; }
        GNB_TRY(tc_core_forward(ctx, g, w.tc, w.blk, w.ffn, w.ln1, w.ln2, x.e, x.n, x.g, y.e, y.n, y.g,
                                have_pre ? pre[li & 1] : TcPreRows(), next, next.pk ? pre[(li + 1) & 1] : TcPreRows(), dec));
        have_pre = next.pk != nullptr;
        have_partial = dec.W4 != nullptr;
      } else {
        GNB_TRY(run_core_fp32(ctx, g, w, x, y));
        have_pre = false;
        have_partial = false;
      }
    }
    x = Feat{y.e, y.n, y.g};
  }
  return GNB_OK;
}



static int forward_device(gnb_ctx* ctx, const gnb_model* m, const gnb_graph* g, const float* ef, const float* nf,
                          const float* gf, float* out_ef, float* out_nf, float* out_gf, int precision, bool reset_arena) {
  const char* chain_env = getenv("GNB_CHAIN_FORWARDS");
  if (!ctx || ctx->device < 0 || ctx->device >= 64 || (chain_env && atoi(chain_env) == 0)) {
    const int rc0 = forward_device_impl(ctx, m, g, ef, nf, gf, out_ef, out_nf, out_gf, precision, reset_arena);
    if (ctx && ctx->h_abort) cudaMemcpyAsync(ctx->h_abort, ctx->d_abort, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    return rc0;
  }
  std::lock_guard<std::mutex> lk(g_chain_mu);      // held while the kernels are enqueued (host time only)
  const int d = ctx->device;
  cudaSetDevice(d);
  if (g_chain_ctx[d] && g_chain_ctx[d] != ctx && g_chain_ev[d]) cudaStreamWaitEvent(ctx->stream, g_chain_ev[d], 0);
  const int rc = forward_device_impl(ctx, m, g, ef, nf, gf, out_ef, out_nf, out_gf, precision, reset_arena);
  cudaMemcpyAsync(ctx->h_abort, ctx->d_abort, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
  if (!ctx->done_ev) cudaEventCreateWithFlags(&ctx->done_ev, cudaEventDisableTiming);
  if (ctx->done_ev && cudaEventRecord(ctx->done_ev, ctx->stream) == cudaSuccess) { g_chain_ev[d] = ctx->done_ev; g_chain_ctx[d] = ctx; }
  return rc;
}

// ------------------------------------------------------------------ forward as a CUDA graph
// A device-resident forward is 30-50 dependent launches of 5-900 us kernels: the launch gaps are 4 % of config 4's step and
// most of config 2's.  The second forward with the same (model, graph, buffers, precision, workspace) is captured into a CUDA
// graph and replayed from then on (steady-state serving / evaluation loops call with the same buffers).  Nothing in a forward
// depends on host state other than the key: scratch comes from the bump arena (same sequence -> same pointers while the arena
// generation is unchanged).  GNB_CUDA_GRAPH=0 disables; profiling (per-launch events) and callers that are themselves
// capturing the stream always run eagerly.
static int graphs_enabled() {
  static const int on = [] { const char* e = getenv("GNB_CUDA_GRAPH"); return e ? atoi(e) : 1; }();
  return on;
}
static int env_signature() {      // environment toggles that are read per launch (tests flip them between calls)
  const char *p = getenv("GNB_EDGE_CTA_PAIR"), *q = getenv("GNB_FFN_CTA_PAIR"), *r = getenv("GNB_FUSE_DECODER");
  return (p ? 1 + atoi(p) : 0) | ((q ? 1 + atoi(q) : 0) << 4) | ((r ? 1 + atoi(r) : 0) << 8);
}
static int forward_graphed(gnb_ctx* ctx, const gnb_model* m, const gnb_graph* g, const float* ef, const float* nf,
                           const float* gf, float* out_ef, float* out_nf, float* out_gf, int precision) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (!ctx || !m || !g || !graphs_enabled() || ctx->profiling ||
      cudaStreamIsCapturing(ctx->stream, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone)
    return forward_device(ctx, m, g, ef, nf, gf, out_ef, out_nf, out_gf, precision, true);
  const void* ptr[6] = {ef, nf, gf, out_ef, out_nf, out_gf};
  const int sig = env_signature();
  const bool legacy = ctx->stream == nullptr || ctx->stream == cudaStreamLegacy;
  FwdGraph* fg = nullptr;
  for (auto& e : ctx->fwd_graphs)
    if (e.model_id == m->id && e.graph_uid == g->uid && e.arena_gen == ctx->arena.gen && e.precision == precision &&
        e.env_sig == sig && memcmp(e.ptr, ptr, sizeof(ptr)) == 0) { fg = &e; break; }
  if (!fg) {
    // first sight: run eagerly (sizes the arena, packs weights, sets function attributes); remember the key
    const int rc = forward_device(ctx, m, g, ef, nf, gf, out_ef, out_nf, out_gf, precision, true);
    if (rc != GNB_OK) return rc;
    ctx->arena.reset();      // coalesce the workspace now (waits for the forward if it has to re-allocate): its addresses are final
    if (ctx->fwd_graphs.size() >= 8) {      // evict the least recently used entry
      size_t lru = 0;
      for (size_t i = 1; i < ctx->fwd_graphs.size(); i++) if (ctx->fwd_graphs[i].last_use < ctx->fwd_graphs[lru].last_use) lru = i;
      if (ctx->fwd_graphs[lru].exec) cudaGraphExecDestroy(ctx->fwd_graphs[lru].exec);
      ctx->fwd_graphs.erase(ctx->fwd_graphs.begin() + lru);
    }
    FwdGraph e{};
    e.model_id = m->id; e.graph_uid = g->uid; e.arena_gen = ctx->arena.gen; e.precision = precision; e.env_sig = sig;
    memcpy(e.ptr, ptr, sizeof(ptr));
    e.seen = 1; e.last_use = ++ctx->fwd_tick;
    ctx->fwd_graphs.push_back(e);
    return GNB_OK;
  }
  fg->last_use = ++ctx->fwd_tick;
  if (!fg->exec && !fg->no_graph) {
    // capture this forward (the capture itself executes nothing), then fall through to the replay
    cudaSetDevice(ctx->device);
    const int64_t l0 = ctx->launches;
    const uint64_t gen0 = ctx->arena.gen;
    cudaGraph_t graph = nullptr;
    cudaStream_t user = ctx->stream;
    bool ok = true;
    if (legacy) {
      if (!ctx->gstream) ok = cudaStreamCreateWithFlags(&ctx->gstream, cudaStreamNonBlocking) == cudaSuccess &&
                              cudaEventCreateWithFlags(&ctx->g_fork, cudaEventDisableTiming) == cudaSuccess &&
                              cudaEventCreateWithFlags(&ctx->g_join, cudaEventDisableTiming) == cudaSuccess;
      if (ok) ctx->stream = ctx->gstream;
    }
    ok = ok && cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed) == cudaSuccess;
    int rc = GNB_OK;
    if (ok) {
      rc = forward_device_impl(ctx, m, g, ef, nf, gf, out_ef, out_nf, out_gf, precision, true);
      ok = cudaStreamEndCapture(ctx->stream, &graph) == cudaSuccess && rc == GNB_OK && graph != nullptr && gen0 == ctx->arena.gen;
    }
    ctx->stream = user;
    if (ok) ok = cudaGraphInstantiate(&fg->exec, graph, 0) == cudaSuccess;
    if (graph) cudaGraphDestroy(graph);
    fg->launches = ctx->launches - l0;
    ctx->launches = l0;
    if (!ok) {
      cudaGetLastError();
      fg->exec = nullptr;
      fg->no_graph = true;
      if (gen0 != ctx->arena.gen) fg->arena_gen = ~0ull;      // the workspace moved under the capture: this key is dead
    }
  }
  if (!fg->exec) return forward_device(ctx, m, g, ef, nf, gf, out_ef, out_nf, out_gf, precision, true);
  // replay, with the same cross-context ordering and watchdog mirror as an eager forward (forward_device)
  const char* chain_env = getenv("GNB_CHAIN_FORWARDS");
  const bool chain = ctx->device >= 0 && ctx->device < 64 && !(chain_env && atoi(chain_env) == 0);
  std::unique_lock<std::mutex> lk(g_chain_mu, std::defer_lock);
  if (chain) lk.lock();
  const int d = ctx->device;
  cudaSetDevice(d);
  if (chain && g_chain_ctx[d] && g_chain_ctx[d] != ctx && g_chain_ev[d]) cudaStreamWaitEvent(ctx->stream, g_chain_ev[d], 0);
  if (legacy) {
    GNB_CUDA(cudaEventRecord(ctx->g_fork, ctx->stream));
    GNB_CUDA(cudaStreamWaitEvent(ctx->gstream, ctx->g_fork, 0));
    GNB_CUDA(cudaGraphLaunch(fg->exec, ctx->gstream));
    GNB_CUDA(cudaEventRecord(ctx->g_join, ctx->gstream));
    GNB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->g_join, 0));
  } else {
    GNB_CUDA(cudaGraphLaunch(fg->exec, ctx->stream));
  }
  ctx->launches += fg->launches;
  cudaMemcpyAsync(ctx->h_abort, ctx->d_abort, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
  if (chain) {
    if (!ctx->done_ev) cudaEventCreateWithFlags(&ctx->done_ev, cudaEventDisableTiming);
    if (ctx->done_ev && cudaEventRecord(ctx->done_ev, ctx->stream) == cudaSuccess) { g_chain_ev[d] = ctx->done_ev; g_chain_ctx[d] = ctx; }
  }
  return GNB_OK;
}

extern "C" int gnb_model_forward(gnb_ctx* ctx, const gnb_model* m, const gnb_graph* g, const float* ef,
                                 const float* nf, const float* gf, float* out_ef, float* out_nf, float* out_gf,
                                 int precision) {
  return forward_graphed(ctx, m, g, ef, nf, gf, out_ef, out_nf, out_gf, precision);
}

extern "C" int gnb_model_forward_host(gnb_ctx* ctx, const gnb_model* m, const gnb_graph* g, const float* ef,
                                      const float* nf, const float* gf, float* out_ef, float* out_nf,
                                      float* out_gf, int precision) {
  GNB_CHECK(ctx && m && g, "gnb_model_forward_host: null argument");
  GNB_CUDA(cudaSetDevice(ctx->device));
  ctx->staging.reset();
  int rc = GNB_OK;
  const size_t ie = (size_t)g->E * m->in_dims[0], in = (size_t)g->N * m->in_dims[1], ig = (size_t)g->B * m->in_dims[2];
  const size_t oe = (size_t)g->E * m->out_dims[0], on = (size_t)g->N * m->out_dims[1], og = (size_t)g->B * m->out_dims[2];
  float* d_ie = ie && ef ? arena_ptr<float>(ctx->staging, ie, &rc) : nullptr;
  float* d_in = in && nf ? arena_ptr<float>(ctx->staging, in, &rc) : nullptr;
  float* d_ig = ig && gf ? arena_ptr<float>(ctx->staging, ig, &rc) : nullptr;
  float* d_oe = oe ? arena_ptr<float>(ctx->staging, oe, &rc) : nullptr;
  float* d_on = on ? arena_ptr<float>(ctx->staging, on, &rc) : nullptr;
  float* d_og = og ? arena_ptr<float>(ctx->staging, og, &rc) : nullptr;
  if (rc != GNB_OK) return rc;
  auto& P = ctx->pipe;
  if (!P.copy) {
    GNB_CUDA(cudaStreamCreateWithFlags(&P.copy, cudaStreamNonBlocking));
    for (auto& e : P.ev_chunk) GNB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    GNB_CUDA(cudaEventCreateWithFlags(&P.ev_ready, cudaEventDisableTiming));
    GNB_CUDA(cudaEventCreateWithFlags(&P.ev_in, cudaEventDisableTiming));
  }
  // the copy stream starts behind whatever the compute stream has queued on these buffers (previous forward)
  GNB_CUDA(cudaEventRecord(P.ev_in, ctx->stream));
  GNB_CUDA(cudaStreamWaitEvent(P.copy, P.ev_in, 0));
  if (d_in) GNB_CUDA(cudaMemcpyAsync(d_in, nf, in * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  if (d_ig) GNB_CUDA(cudaMemcpyAsync(d_ig, gf, ig * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  P.ef_pending = false; P.out_pending = false; P.out_early = false;
  if (d_ie) {
    const int64_t E = g->E;
    const int de = m->in_dims[0];
    if (ie * sizeof(float) >= ((size_t)8 << 20)) {      // worth pipelining: 4 row chunks on the copy stream
      P.nchunks = 4;
      const int64_t per = (E + P.nchunks - 1) / P.nchunks;
      for (int ci = 0; ci < P.nchunks; ci++) {
        const int64_t r0 = ci * per, r1 = r0 + per < E ? r0 + per : E;
        if (r0 < r1)
          GNB_CUDA(cudaMemcpyAsync(d_ie + (size_t)r0 * de, ef + (size_t)r0 * de, (size_t)(r1 - r0) * de * sizeof(float),
                                   cudaMemcpyHostToDevice, P.copy));
        GNB_CUDA(cudaEventRecord(P.ev_chunk[ci], P.copy));
      }
      P.d_ef = d_ie; P.ef_rows = E; P.ef_pending = true;
    } else {
      GNB_CUDA(cudaMemcpyAsync(d_ie, ef, ie * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    }
  }
  if (d_oe && out_ef && oe * sizeof(float) >= ((size_t)4 << 20)) {
    P.d_out_ef = d_oe; P.h_out_ef = out_ef; P.out_bytes = oe * sizeof(float); P.out_pending = true;
  }
  int frc = forward_device(ctx, m, g, d_ie, d_in, d_ig, d_oe, d_on, d_og, precision, true);
  const bool early = P.out_early;
  P.ef_pending = false; P.out_pending = false; P.out_early = false;
  if (frc != GNB_OK) {
    cudaStreamSynchronize(P.copy);
    return frc;
  }
  if (d_oe && out_ef && !early) GNB_CUDA(cudaMemcpyAsync(out_ef, d_oe, oe * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  if (d_on && out_nf) GNB_CUDA(cudaMemcpyAsync(out_nf, d_on, on * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  if (d_og && out_gf) GNB_CUDA(cudaMemcpyAsync(out_gf, d_og, og * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  GNB_CUDA(cudaStreamSynchronize(ctx->stream));
  GNB_CUDA(cudaStreamSynchronize(P.copy));
  return ctx_check_watchdog(ctx);
}

static int forward_by_reference(gnb_ctx* ctx, const gnb_graph* g, const gnb_layer* layers, int n, const float* ef,
                                const float* nf, const float* gf, float* oe, float* on, float* og, int precision) {
  GNB_CHECK(precision == GNB_PREC_FP32 || precision == GNB_PREC_AUTO,
            "single-layer forward runs the fp32 path on caller-owned weights; create a gnb_model for the bf16 tensor path");
  gnb_model* m = nullptr;
  GNB_TRY(build_model(ctx, layers, n, 2, &m));
  int r = forward_device(ctx, m, g, ef, nf, gf, oe, on, og, GNB_PREC_FP32, true);
  gnb_model_destroy(m);
  return r;
}

extern "C" int gnb_block_forward(gnb_ctx* ctx, const gnb_graph* g, const gnb_block_params* p, const float* ef,
                                 const float* nf, const float* gf, float* oe, float* on, float* og, int precision) {
  GNB_CHECK(p, "gnb_block_forward: null params");
  gnb_layer L{};
  L.kind = GNB_LAYER_BLOCK;
  L.block = *p;
  return forward_by_reference(ctx, g, &L, 1, ef, nf, gf, oe, on, og, precision);
}
extern "C" int gnb_core_forward(gnb_ctx* ctx, const gnb_graph* g, const gnb_core_params* p, const float* ef,
                                const float* nf, const float* gf, float* oe, float* on, float* og, int precision) {
  return gnb_corelist_forward(ctx, g, p, 1, ef, nf, gf, oe, on, og, precision);
}
extern "C" int gnb_corelist_forward(gnb_ctx* ctx, const gnb_graph* g, const gnb_core_params* cores, int n_cores,
                                    const float* ef, const float* nf, const float* gf, float* oe, float* on,
                                    float* og, int precision) {
  GNB_CHECK(cores && n_cores > 0, "gnb_corelist_forward: no cores");
  std::vector<gnb_layer> L(n_cores);
  for (int i = 0; i < n_cores; i++) {
    memset(&L[i], 0, sizeof(gnb_layer));
    L[i].kind = GNB_LAYER_CORE;
    L[i].core = cores[i];
  }
  return forward_by_reference(ctx, g, L.data(), n_cores, ef, nf, gf, oe, on, og, precision);
}
