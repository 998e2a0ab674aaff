// step.cu — the whole mini-batch step of the hot loop behind ONE C-ABI call.
//
// ngnn_sage_step = loop body of PipelineCO.train (reference src/pipeline.py:152-169) for the SAGE network
// of reference src/models/layers/sage.py:30-40 on one sampled block:
//     forward (per layer: K-AGG -> K-GEMM with fused bias/ReLU/dropout) -> softmax-CE + accuracy count on
//     the seed rows -> backward (per layer: K-WGRAD straight into the flat gradient bucket, K-DGRAD,
//     CSR transpose, K-AGG-T with the producing layer's ReLU/dropout gate folded in).
// Every layer is trimmed to the rows its seed outputs depend on (prefixes of the block; exact — SURVEY §8
// trimming note), layer 1 aggregates straight from the resident feature table by global ids.  The launches
// are issued back to back from C++ on the caller's stream (no host synchronisation, no allocation: all
// scratch comes from one caller-owned arena), so the step costs one FFI crossing instead of ~80 and can be
// captured in a CUDA graph.  The optimizer (ngnn_adam_step) is a separate call so a data-parallel caller can
// all-reduce the gradient bucket in between.
#include "common.cuh"

namespace ngnn {

struct LayerPlan {
  int64_t F, O;                 // in / out channels
  int64_t ldo;                  // leading dimension of this layer's out / dy buffers (O rounded up to 4: TMA-addressable)
  int64_t ldf;                  // leading dimension of this layer's mean / root / dmean buffers (F rounded up to 4)
  bool concat;                  // layer 1: mean and root rows side by side in ONE [n_dst, 2F] buffer (one K = 2F contraction)
  int a, b;                     // hop indices of this layer's extents: n_dst = nodes[a], e_lim = edges[b], n_src = nodes[b]
  int64_t n_dst, e_lim, n_src;  // trimmed extents of this step (host mode; = the capacities in device mode)
  int64_t n_dst_max, e_max, n_src_max;
  size_t off_wl, off_b, off_wr; // element offsets into the flat parameter / gradient buckets
  // arena regions (byte offsets)
  size_t mean, root, out, dy, colptr_t, row_t, perm_t;
  size_t prep_fwd, prep_dg;     // split weight planes of this layer (forward pack / data-gradient pack)
  size_t prep_fwd_bytes, prep_dg_bytes;
};

struct StepPlan {
  int L;
  LayerPlan layer[16];
  size_t dmean, droot, gemm_ws, dgrad_ws, wgrad_ws, wgrad_ws_aux, sort_ws, ce_rows, total;
  size_t agg1_alt;              // byte distance from layer 1's [mean | root] buffer 0 to buffer 1 (ngnn_sage_agg1)
  size_t gemm_ws_bytes, dgrad_ws_bytes, wgrad_ws_bytes, sort_ws_bytes;
  int64_t n_params;
};

static bool make_plan(const ngnn_sage_model_t* m, int32_t H, const int64_t* max_hop_nodes, const int64_t* max_hop_edges,
                      const int32_t* hop_nodes, const int32_t* hop_edges, StepPlan& pl) {
  const int L = m->num_layers;
  if (L < 1 || L > 16 || H < 1 || m->in_dim < 1 || m->out_dim < 1 || (L > 1 && m->hidden_dim < 1)) return false;
  pl.L = L;
  size_t off = 0, poff = 0;
  auto take = [&](size_t bytes) { size_t r = off; off += align_up(bytes > 0 ? bytes : 4, 256); return r; };
  size_t max_dx = 0, gws = 0, dws = 0, wws = 0, sws = 0;
  for (int i = 0; i < L; ++i) {
    LayerPlan& lp = pl.layer[i];
    lp.F = i == 0 ? m->in_dim : m->hidden_dim;
    lp.O = i == L - 1 ? m->out_dim : m->hidden_dim;
    lp.ldo = (lp.O + 3) / 4 * 4;
    lp.ldf = (lp.F + 3) / 4 * 4;
    lp.concat = i == 0 && lp.F % 4 == 0 && (2 * lp.F + 31) / 32 < 2 * ((lp.F + 31) / 32);   // only where it saves a K-block
    if (lp.concat) lp.ldf = 2 * lp.F;
    const int d = L - 1 - i;   // hops between this layer's outputs and the seeds
    const int a = d < H ? d : H, b = d + 1 < H ? d + 1 : H;
    lp.a = a; lp.b = b;
    lp.n_dst_max = max_hop_nodes[a]; lp.e_max = max_hop_edges[b]; lp.n_src_max = max_hop_nodes[b];
    if (hop_nodes) { lp.n_dst = hop_nodes[a]; lp.e_lim = hop_edges[b]; lp.n_src = hop_nodes[b]; }
    else { lp.n_dst = lp.n_dst_max; lp.e_lim = lp.e_max; lp.n_src = lp.n_src_max; }
    if (lp.n_dst > lp.n_dst_max || lp.e_lim > lp.e_max || lp.n_src > lp.n_src_max) return false;
    lp.off_wl = poff; poff += (size_t)lp.O * lp.F;
    lp.off_b = poff; poff += (size_t)lp.O;
    lp.off_wr = poff; poff += (size_t)lp.O * lp.F;
    lp.mean = take((size_t)lp.n_dst_max * lp.ldf * 4);
    lp.root = i == 0 ? (lp.concat ? lp.mean + (size_t)lp.F * 4 : take((size_t)lp.n_dst_max * lp.ldf * 4)) : 0;
    if (i == 0) {     // a second copy of layer 1's aggregation buffers: the NEXT block's aggregation runs beside this block's step
      const size_t first = lp.mean;
      const size_t alt = take((size_t)lp.n_dst_max * lp.ldf * 4);
      if (!lp.concat) take((size_t)lp.n_dst_max * lp.ldf * 4);
      pl.agg1_alt = alt - first;
    }
    lp.out = take((size_t)lp.n_dst_max * lp.ldo * 4);
    lp.dy = take((size_t)lp.n_dst_max * lp.ldo * 4);
    if (i > 0) {
      lp.colptr_t = take((size_t)(lp.n_src_max + 1) * 4);
      lp.row_t = take((size_t)lp.e_max * 4);
      lp.perm_t = take((size_t)lp.e_max * 4);
      if ((size_t)lp.n_dst_max * lp.ldf * 4 > max_dx) max_dx = (size_t)lp.n_dst_max * lp.ldf * 4;
      const size_t s = ngnn_csr_transpose_workspace_bytes(lp.e_max, lp.n_src_max);
      if (s > sws) sws = s;
      const size_t dg = ngnn_sage_dgrad_workspace_bytes(lp.F, lp.O);
      if (dg > dws) dws = dg;
    }
    const size_t g = ngnn_sage_gemm_workspace_bytes(lp.F, lp.O);
    if (g > gws) gws = g;
    lp.prep_fwd_bytes = g; lp.prep_fwd = take(g);
    lp.prep_dg_bytes = i > 0 ? ngnn_sage_dgrad_workspace_bytes(lp.F, lp.O) : 0;
    lp.prep_dg = i > 0 ? take(lp.prep_dg_bytes) : 0;
    const size_t w = ngnn_sage_wgrad_workspace_bytes(lp.n_dst_max, lp.F, lp.O);
    if (w > wws) wws = w;
  }
  pl.n_params = (int64_t)poff;
  pl.dmean = take(max_dx); pl.droot = take(max_dx);
  pl.gemm_ws = take(gws); pl.dgrad_ws = take(dws); pl.wgrad_ws = take(wws); pl.wgrad_ws_aux = take(wws); pl.sort_ws = take(sws);
  pl.gemm_ws_bytes = gws; pl.dgrad_ws_bytes = dws; pl.wgrad_ws_bytes = wws; pl.sort_ws_bytes = sws;
  pl.ce_rows = take((size_t)max_hop_nodes[0] * 2 * 4);
  pl.total = off + 256;
  return true;
}

// The weight gradients of layers >= 2 are off the backward's critical path (dY_l -> dgrad -> transpose-sum -> dY_{l-1}):
// they are forked onto an auxiliary stream (event fork / join, graph-capturable) so their small grids run under the
// dgrad / K-AGG-T chain instead of after it.
// One set per device (a process may step models on several GPUs); one stepping thread per device at a time.
struct AuxCtx { cudaStream_t stream; cudaEvent_t fork, join; };
static AuxCtx g_auxs[64] = {};
static bool g_use_aux = true;

static int32_t ensure_aux(AuxCtx** out) {
  int dev = 0;
  NGNN_CUDA(cudaGetDevice(&dev));
  NGNN_REQUIRE(dev >= 0 && dev < 64, NGNN_E_UNSUPPORTED, "sage_step: device ordinal %d out of range", dev);
  AuxCtx& a = g_auxs[dev];
  if (a.stream == nullptr) {
    NGNN_CUDA(cudaStreamCreateWithFlags(&a.stream, cudaStreamNonBlocking));
    NGNN_CUDA(cudaEventCreateWithFlags(&a.fork, cudaEventDisableTiming));
    NGNN_CUDA(cudaEventCreateWithFlags(&a.join, cudaEventDisableTiming));
  }
  *out = &a;
  return NGNN_OK;
}

// Optional in-situ timing of the layer-1 aggregation launch inside ngnn_sage_step (bench.py's roofline):
// CUDA events recorded on the caller's stream around that one launch, a pair per step.
static cudaEvent_t* g_probe_ev = nullptr;
static cudaEvent_t* g_probe_t_ev = nullptr;           // second pair per sample: the K-AGG-T launch into layer 1's output rows
static unsigned char* g_probe_t_done = nullptr;
static int g_probe_cap = 0, g_probe_n = 0;
static unsigned long long* g_probe_clk = nullptr;     // device [2 * cap]: (min start, max end) %globaltimer of each probed launch

}  // namespace ngnn

using namespace ngnn;

extern "C" {

int32_t ngnn_probe_enable(int32_t max_samples) {
  for (int i = 0; i < 2 * g_probe_cap; ++i) { cudaEventDestroy(g_probe_ev[i]); cudaEventDestroy(g_probe_t_ev[i]); }
  delete[] g_probe_ev;
  delete[] g_probe_t_ev;
  delete[] g_probe_t_done;
  g_probe_t_ev = nullptr; g_probe_t_done = nullptr;
  if (g_probe_clk) cudaFree(g_probe_clk);
  g_probe_ev = nullptr; g_probe_clk = nullptr; g_probe_cap = 0; g_probe_n = 0;
  if (max_samples <= 0) return NGNN_OK;
  g_probe_ev = new cudaEvent_t[2 * (size_t)max_samples];
  g_probe_t_ev = new cudaEvent_t[2 * (size_t)max_samples];
  g_probe_t_done = new unsigned char[(size_t)max_samples]();
  for (int i = 0; i < 2 * max_samples; ++i) { NGNN_CUDA(cudaEventCreate(&g_probe_ev[i])); NGNN_CUDA(cudaEventCreate(&g_probe_t_ev[i])); }
  {   // (a measurement facility: the one place the library allocates) start words = all ones, end words = 0
    NGNN_CUDA(cudaMalloc(&g_probe_clk, 2 * (size_t)max_samples * sizeof(unsigned long long)));
    unsigned long long* h = new unsigned long long[2 * (size_t)max_samples];
    for (int i = 0; i < max_samples; ++i) { h[2 * i] = ~0ull; h[2 * i + 1] = 0ull; }
    cudaError_t e = cudaMemcpy(g_probe_clk, h, 2 * (size_t)max_samples * sizeof(unsigned long long), cudaMemcpyHostToDevice);
    delete[] h;
    NGNN_CUDA(e);
  }
  g_probe_cap = max_samples;
  return NGNN_OK;
}

int32_t ngnn_probe_read(float* ms, int32_t cap, int32_t* n) {
  NGNN_REQUIRE(ms && n, NGNN_E_INVALID, "probe_read: null pointer");
  const int m = g_probe_n < cap ? g_probe_n : cap;
  for (int i = 0; i < m; ++i) {
    NGNN_CUDA(cudaEventSynchronize(g_probe_ev[2 * i + 1]));
    NGNN_CUDA(cudaEventElapsedTime(&ms[i], g_probe_ev[2 * i], g_probe_ev[2 * i + 1]));
  }
  *n = m;
  return NGNN_OK;
}

int32_t ngnn_probe_read_agg_t(float* ms, int32_t cap, int32_t* n) {
  NGNN_REQUIRE(ms && n, NGNN_E_INVALID, "probe_read_agg_t: null pointer");
  int m = 0;
  for (int i = 0; i < g_probe_n && m < cap; ++i) {
    if (!g_probe_t_done || !g_probe_t_done[i]) continue;
    NGNN_CUDA(cudaEventSynchronize(g_probe_t_ev[2 * i + 1]));
    NGNN_CUDA(cudaEventElapsedTime(&ms[m], g_probe_t_ev[2 * i], g_probe_t_ev[2 * i + 1]));
    ++m;
  }
  *n = m;
  return NGNN_OK;
}

int32_t ngnn_probe_read_device_clock(float* ms, int32_t cap, int32_t* n) {
  NGNN_REQUIRE(ms && n, NGNN_E_INVALID, "probe_read_device_clock: null pointer");
  const int m = g_probe_n < cap ? g_probe_n : cap;
  *n = 0;
  if (m == 0 || g_probe_clk == nullptr) return NGNN_OK;
  NGNN_CUDA(cudaDeviceSynchronize());
  unsigned long long* h = new unsigned long long[2 * (size_t)m];
  cudaError_t e = cudaMemcpy(h, g_probe_clk, 2 * (size_t)m * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess)
    for (int i = 0; i < m; ++i) ms[i] = h[2 * i + 1] > h[2 * i] ? (float)((double)(h[2 * i + 1] - h[2 * i]) * 1e-6) : 0.f;
  delete[] h;
  NGNN_CUDA(e);
  *n = m;
  return NGNN_OK;
}

int64_t ngnn_sage_num_params(const ngnn_sage_model_t* model) {
  if (!model) return -1;
  int64_t n = 0;
  for (int i = 0; i < model->num_layers; ++i) {
    const int64_t F = i == 0 ? model->in_dim : model->hidden_dim;
    const int64_t O = i == model->num_layers - 1 ? model->out_dim : model->hidden_dim;
    n += 2 * O * F + O;
  }
  return n;
}

size_t ngnn_sage_step_workspace_bytes(const ngnn_sage_model_t* model, int32_t num_hops, const int64_t* max_hop_nodes,
                                      const int64_t* max_hop_edges) {
  StepPlan pl;
  if (!model || !max_hop_nodes || !max_hop_edges || !make_plan(model, num_hops, max_hop_nodes, max_hop_edges, nullptr, nullptr, pl))
    return 0;
  return pl.total;
}

// Collects the weight-pack jobs of every layer (forward packs; data-gradient packs of layers >= 2 when training) into the
// thread's batch; which layers take the tensor-core path comes back in the two flag arrays.  prep_batch_launch runs them.
static int32_t prep_weights_collect(const StepPlan& pl, const float* params, bool train, char* base, bool* prep_fwd_ok,
                                    bool* prep_dg_ok) {
  prep_batch_begin();
  for (int i = 0; i < pl.L; ++i) {
    const LayerPlan& lp = pl.layer[i];
    int32_t rc = prep_batch_add(lp.concat ? 2 : 0, params + lp.off_wl, params + lp.off_wr, lp.F, lp.O, base + lp.prep_fwd, lp.prep_fwd_bytes);
    if (rc != NGNN_OK && rc != NGNN_E_UNSUPPORTED) return rc;
    prep_fwd_ok[i] = rc == NGNN_OK;
    prep_dg_ok[i] = false;
    if (train && i > 0) {
      rc = prep_batch_add(1, params + lp.off_wl, params + lp.off_wr, lp.F, lp.O, base + lp.prep_dg, lp.prep_dg_bytes);
      if (rc != NGNN_OK && rc != NGNN_E_UNSUPPORTED) return rc;
      prep_dg_ok[i] = rc == NGNN_OK;
    }
  }
  return NGNN_OK;
}

// Layer 1's aggregation: mean of the sampled in-neighbours + root gather, straight from the resident feature table by global
// id, into copy `buffer` (0 / 1) of the arena's [mean | root] region.  It depends on the block and the table only — not on the
// parameters — so a caller may run it for the NEXT block beside the current block's step (ngnn_sage_agg1).
static int32_t agg1_launch(const ngnn_block_t* block, const StepPlan& pl, const float* table, int64_t ld_table, char* base,
                           int buffer, cudaStream_t st) {
  const LayerPlan& lp = pl.layer[0];
  const bool dev = block->counts != nullptr;
  const Ext n_dst = dev ? ext_dev(block->counts + lp.a, lp.n_dst_max) : ext_host(lp.n_dst);
  const size_t off = buffer == 1 ? pl.agg1_alt : 0;
  float* mean = reinterpret_cast<float*>(base + lp.mean + off);
  float* root = reinterpret_cast<float*>(base + lp.root + off);
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(st, &cap);
  const bool probe = g_probe_n < g_probe_cap && cap == cudaStreamCaptureStatusNone;
  if (probe) cudaEventRecord(g_probe_ev[2 * g_probe_n], st);
  const bool remapped = block->col_table != nullptr && block->n_table != nullptr;      // table stored hot rows first
  const int32_t rc = agg_fwd_table_impl(block->rowptr, remapped ? block->col_table : block->col_global, table, ld_table, n_dst, lp.F,
                                        mean, lp.ldf, remapped ? block->n_table : block->n_id, root, lp.ldf,
                                        remapped ? block->hot_rows : -1, st, probe && g_probe_clk ? g_probe_clk + 2 * g_probe_n : nullptr);
  if (probe) { cudaEventRecord(g_probe_ev[2 * g_probe_n + 1], st); ++g_probe_n; }
  return rc;
}

// phase 0: forward + loss + backward (ngnn_sage_step); phase 1: training-mode forward only, activations stay in ws
// (ngnn_sage_forward); phase 2: backward only from a caller-supplied top-layer gradient (ngnn_sage_backward).
//
// Extents.  Host mode (block->counts == NULL): the per-hop extents come from block->hop_nodes / hop_edges and every kernel
// is launched for exactly those sizes.  Device mode (block->counts != NULL): the extents stay in the sampler's device-side
// `counts`, every launch is sized for the declared capacities (max_hop_*) and reads its row count on the device, so the
// launch sequence is the same for every block and the whole step can be captured once in a CUDA graph and replayed.
static int32_t sage_step_body(const ngnn_sage_model_t* model, const float* params, float* grads, const ngnn_block_t* block,
                              const StepPlan& pl, const float* table, int64_t ld_table,
                              const int64_t* target_global, const int64_t* label_global, uint64_t drop_seed, uint64_t drop_offset,
                              float* stats, float* logits_out, int64_t ld_logits, char* base, ngnn_stream_t stream,
                              int32_t phase, const float* dlogits_in, int64_t ld_dlogits, AuxCtx* aux, bool& aux_used) {
  auto F32 = [&](size_t off) { return reinterpret_cast<float*>(base + off); };
  auto I32 = [&](size_t off) { return reinterpret_cast<int32_t*>(base + off); };
  const int L = pl.L;
  const bool dev = block->counts != nullptr;
  const int H = block->num_hops;
  const int64_t bs = dev ? block->batch_size : block->hop_nodes[0];
  const bool train = grads != nullptr || phase != 0;
  const float p_drop = (train && model->training) ? model->dropout : 0.f;
  const StepCtl* ctl = reinterpret_cast<const StepCtl*>(block->ctl);
  cudaStream_t st = as_stream(stream);
  // extent of "nodes within hop a" / "edges within hop b" for a kernel: device word + capacity, or the host value
  auto nodes_ext = [&](int a, int64_t cap, int64_t host) { return dev ? ext_dev(block->counts + a, cap) : ext_host(host); };
  const size_t a1off = block->agg1_buffer == 2 ? pl.agg1_alt : 0;     // which copy of layer 1's [mean | root] this step reads
  int32_t rc;

  // ---------------- split weight planes of every layer: ONE launch, before the step's first kernel ----------------
  bool prep_fwd_ok[16], prep_dg_ok[16];
  rc = prep_weights_collect(pl, params, train, base, prep_fwd_ok, prep_dg_ok);
  if (rc != NGNN_OK) return rc;
  if (phase != 2 && !block->weights_prepared) {   // phase 2: the planes of the matching forward call are still valid;
    rc = prep_batch_launch(st);                    // weights_prepared: ngnn_sage_prep_weights already ran for these parameters
    if (rc != NGNN_OK) return rc;
  }

  // ---------------- forward ----------------
  for (int i = 0; i < L && phase != 2; ++i) {
    const LayerPlan& lp = pl.layer[i];
    const Ext n_dst = nodes_ext(lp.a, lp.n_dst_max, lp.n_dst);
    const float* root;
    int64_t ld_root;
    if (i == 0) {   // aggregate from the resident table by global ids; gather the root rows in the same launch
      rc = NGNN_OK;
      if (block->agg1_buffer == 0) rc = agg1_launch(block, pl, table, ld_table, base, 0, st);   // else: done by ngnn_sage_agg1
      root = F32(lp.root + a1off); ld_root = lp.ldf;
    } else {
      const LayerPlan& prev = pl.layer[i - 1];
      rc = agg_fwd_impl(block->rowptr, block->col, F32(prev.out), prev.ldo, n_dst, lp.F, F32(lp.mean), lp.ldf, st);
      root = F32(prev.out); ld_root = prev.ldo;
    }
    if (rc != NGNN_OK) return rc;
    const bool last = i == L - 1;
    rc = gemm_fwd_impl(F32(lp.mean + (i == 0 ? a1off : 0)), lp.ldf, root, ld_root, params + lp.off_wl, params + lp.off_wr, params + lp.off_b,
                       n_dst.cap, lp.F, lp.O, last ? NGNN_ACT_NONE : NGNN_ACT_RELU, last ? 0.f : p_drop, drop_seed,
                       drop_offset + (uint64_t)i, F32(lp.out), lp.ldo, nullptr, base + lp.prep_fwd, lp.prep_fwd_bytes,
                       st, prep_fwd_ok[i], n_dst.dev, ctl, (uint32_t)i, lp.concat && prep_fwd_ok[i]);
    if (rc != NGNN_OK) return rc;
  }
  const LayerPlan& top = pl.layer[L - 1];
  if (logits_out && phase != 2) {
    NGNN_REQUIRE(ld_logits >= top.O, NGNN_E_INVALID, "sage_step: ld_logits < out_dim");
    NGNN_CUDA(cudaMemcpy2DAsync(logits_out, (size_t)ld_logits * 4, F32(top.out), (size_t)top.ldo * 4, (size_t)top.O * 4, (size_t)bs,
                                cudaMemcpyDeviceToDevice, st));
  }
  if (phase == 1) return NGNN_OK;                 // forward of a split step: the caller computes the loss gradient
  if (phase == 2) {                               // top-layer gradient supplied by the caller (e.g. the co-teaching loss)
    NGNN_CUDA(cudaMemcpy2DAsync(F32(top.dy), (size_t)top.ldo * 4, dlogits_in, (size_t)ld_dlogits * 4, (size_t)top.O * 4,
                                (size_t)bs, cudaMemcpyDeviceToDevice, st));
  } else {
    if (target_global == nullptr) return NGNN_OK;   // inference: forward only

    // ---------------- loss on the seed rows (labels gathered by global id) ----------------
    rc = ce_impl(F32(top.out), top.ldo, target_global, label_global, block->n_id, bs, top.O, 1.0f, stats,
                 train ? F32(top.dy) : nullptr, top.ldo, F32(pl.ce_rows), ctl, st);
    if (rc != NGNN_OK) return rc;
    if (!train) return NGNN_OK;
  }

  // ---------------- backward ----------------
  const bool use_aux = aux != nullptr;
  for (int i = L - 1; i >= 0; --i) {
    const LayerPlan& lp = pl.layer[i];
    const float* root = i == 0 ? F32(lp.root + a1off) : F32(pl.layer[i - 1].out);
    const float* mean = F32(lp.mean + (i == 0 ? a1off : 0));
    const int64_t ld_root = i == 0 ? lp.ldf : pl.layer[i - 1].ldo;
    // only the first bs rows of the top layer carry a gradient
    const Ext n_rows = i == L - 1 ? ext_host(bs) : nodes_ext(lp.a, lp.n_dst_max, lp.n_dst);
    if (use_aux && i > 0) {    // fork: dY_i is complete on the main stream here
      NGNN_CUDA(cudaEventRecord(aux->fork, st));
      NGNN_CUDA(cudaStreamWaitEvent(aux->stream, aux->fork, 0));
      aux_used = true;
      rc = wgrad_impl(F32(lp.dy), lp.ldo, mean, lp.ldf, root, ld_root, n_rows.cap, n_rows.dev, lp.F, lp.O,
                      grads + lp.off_wl, grads + lp.off_wr, grads + lp.off_b, 0, base + pl.wgrad_ws_aux, pl.wgrad_ws_bytes,
                      aux->stream);
    } else {
      rc = wgrad_impl(F32(lp.dy), lp.ldo, mean, lp.ldf, root, ld_root, n_rows.cap, n_rows.dev, lp.F, lp.O,
                      grads + lp.off_wl, grads + lp.off_wr, grads + lp.off_b, 0, base + pl.wgrad_ws, pl.wgrad_ws_bytes, st);
    }
    if (rc != NGNN_OK) return rc;
    if (i == 0) break;   // features are leaves: no data gradient for layer 1
    const LayerPlan& prev = pl.layer[i - 1];
    rc = dgrad_impl(F32(lp.dy), lp.ldo, params + lp.off_wl, params + lp.off_wr, block->rowptr, n_rows.cap, lp.F, lp.O,
                    F32(pl.dmean), lp.ldf, F32(pl.droot), lp.ldf, base + lp.prep_dg, lp.prep_dg_bytes, st, prep_dg_ok[i], n_rows.dev);
    if (rc != NGNN_OK) return rc;
    const int32_t* colptr_t = I32(lp.colptr_t);
    const int32_t* row_t = I32(lp.row_t);
    const int hop_b = lp.b;                             // hop prefix this layer's edges span
    if (hop_b < 8 && block->colptr_t[hop_b] != nullptr && block->row_t[hop_b] != nullptr) {
      colptr_t = block->colptr_t[hop_b];           // built by the sampler (or by the loader on its side stream)
      row_t = block->row_t[hop_b];
    } else {
      NGNN_REQUIRE(!dev, NGNN_E_INVALID, "sage_step: device-side extents need the block's transposes (colptr_t[%d] / row_t[%d])",
                   hop_b, hop_b);
      const int64_t e_lim = i == L - 1 ? block->hop_edges[1 < H ? 1 : H] : lp.e_lim;
      rc = ngnn_csr_transpose(block->rowptr, block->col, n_rows.cap, e_lim, lp.n_src, I32(lp.colptr_t), I32(lp.row_t), I32(lp.perm_t),
                              base + pl.sort_ws, pl.sort_ws_bytes, stream);
      if (rc != NGNN_OK) return rc;
    }
    // dY of the previous layer = gate(prev output) * (transpose-sum of dmean + droot on the root rows)
    // (second probe of ngnn_probe_*: the K-AGG-T launch into layer 1's output rows, the widest of the step; the probe index was
    //  advanced by the layer-1 K-AGG of this step)
    bool probe_t = i == 1 && g_probe_t_ev != nullptr && g_probe_n >= 1 && g_probe_n <= g_probe_cap && !g_probe_t_done[g_probe_n - 1];
    if (probe_t) {
      cudaStreamCaptureStatus capturing = cudaStreamCaptureStatusNone;
      if (cudaStreamIsCapturing(st, &capturing) != cudaSuccess || capturing != cudaStreamCaptureStatusNone) { cudaGetLastError(); probe_t = false; }
    }
    if (probe_t) cudaEventRecord(g_probe_t_ev[2 * (g_probe_n - 1)], st);
    rc = agg_bwd_impl(colptr_t, row_t, F32(pl.dmean), lp.ldf, nodes_ext(lp.b, lp.n_src_max, lp.n_src), lp.F, F32(pl.droot), lp.ldf,
                      n_rows, F32(prev.out), prev.ldo, 1.0f / (1.0f - p_drop), F32(prev.dy), prev.ldo, st);
    if (probe_t) { cudaEventRecord(g_probe_t_ev[2 * (g_probe_n - 1) + 1], st); g_probe_t_done[g_probe_n - 1] = 1; }
    if (rc != NGNN_OK) return rc;
  }
  return NGNN_OK;
}

static int32_t sage_step_impl(const ngnn_sage_model_t* model, const float* params, float* grads, const ngnn_block_t* block,
                              const int64_t* max_hop_nodes, const int64_t* max_hop_edges, const float* table, int64_t ld_table,
                              const int64_t* target_global, const int64_t* label_global, uint64_t drop_seed, uint64_t drop_offset,
                              float* stats, float* logits_out, int64_t ld_logits, void* ws, size_t ws_bytes, ngnn_stream_t stream,
                              int32_t phase, const float* dlogits_in, int64_t ld_dlogits) {
  NGNN_REQUIRE(model && params && block && table && ws && (stats || phase != 0), NGNN_E_INVALID, "sage_step: null pointer");
  NGNN_REQUIRE(block->rowptr && block->col && block->col_global && block->n_id, NGNN_E_INVALID, "sage_step: incomplete block");
  const bool dev = block->counts != nullptr;
  NGNN_REQUIRE(dev || (block->hop_nodes && block->hop_edges), NGNN_E_INVALID, "sage_step: the block carries no extents");
  NGNN_REQUIRE(!dev || block->batch_size > 0, NGNN_E_INVALID, "sage_step: device-side extents need block->batch_size");
  NGNN_REQUIRE(model->dropout >= 0.f && model->dropout < 1.f, NGNN_E_INVALID, "sage_step: dropout outside [0,1)");
  NGNN_REQUIRE(block->agg1_buffer >= 0 && block->agg1_buffer <= 2, NGNN_E_INVALID, "sage_step: agg1_buffer must be 0, 1 or 2");
  StepPlan pl;
  NGNN_REQUIRE(make_plan(model, block->num_hops, max_hop_nodes, max_hop_edges, dev ? nullptr : block->hop_nodes,
                         dev ? nullptr : block->hop_edges, pl),
               NGNN_E_INVALID, "sage_step: bad model / block extents (block larger than the declared capacity?)");
  NGNN_REQUIRE(!dev || block->batch_size <= max_hop_nodes[0], NGNN_E_INVALID, "sage_step: batch_size above the declared capacity");
  NGNN_REQUIRE(ws_bytes >= pl.total, NGNN_E_WORKSPACE, "sage_step: workspace too small (%zu < %zu)", ws_bytes, pl.total);
  NGNN_REQUIRE(ld_table >= model->in_dim, NGNN_E_INVALID, "sage_step: ld_table < in_dim");
  const bool train = grads != nullptr || phase != 0;
  NGNN_REQUIRE(!train || phase != 0 || target_global != nullptr, NGNN_E_INVALID, "sage_step: training needs targets");
  NGNN_REQUIRE(phase != 2 || (grads != nullptr && dlogits_in != nullptr && ld_dlogits >= model->out_dim), NGNN_E_INVALID,
               "sage_backward: needs the gradient bucket and dlogits");
  char* base = reinterpret_cast<char*>(align_up(reinterpret_cast<size_t>(ws), 256));
  AuxCtx* aux = nullptr;
  if (g_use_aux && pl.L > 1 && train && phase != 1) {
    const int32_t rc = ensure_aux(&aux);
    if (rc != NGNN_OK) return rc;
  }
  bool aux_used = false;
  const int32_t rc = sage_step_body(model, params, grads, block, pl, table, ld_table, target_global, label_global, drop_seed,
                                    drop_offset, stats, logits_out, ld_logits, base, stream, phase, dlogits_in, ld_dlogits, aux,
                                    aux_used);
  if (aux_used) {   // join on EVERY path (also a failed one): the caller's stream continues only after the auxiliary stream's work —
                    // a forked stream left dangling would also break an enclosing graph capture
    cudaError_t e1 = cudaEventRecord(aux->join, aux->stream);
    cudaError_t e2 = cudaStreamWaitEvent(as_stream(stream), aux->join, 0);
    if (rc == NGNN_OK && (e1 != cudaSuccess || e2 != cudaSuccess))
      return set_error(NGNN_E_CUDA, "sage_step: joining the auxiliary stream failed: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
  }
  return rc;
}

int32_t ngnn_sage_agg1(const ngnn_sage_model_t* model, const ngnn_block_t* block, const int64_t* max_hop_nodes,
                       const int64_t* max_hop_edges, const float* table, int64_t ld_table, int32_t buffer, void* ws, size_t ws_bytes,
                       ngnn_stream_t stream) {
  NGNN_REQUIRE(model && block && table && ws && max_hop_nodes && max_hop_edges, NGNN_E_INVALID, "sage_agg1: null pointer");
  NGNN_REQUIRE(block->rowptr && block->col_global && block->n_id, NGNN_E_INVALID, "sage_agg1: incomplete block");
  NGNN_REQUIRE(buffer == 0 || buffer == 1, NGNN_E_INVALID, "sage_agg1: buffer must be 0 or 1");
  const bool dev = block->counts != nullptr;
  NGNN_REQUIRE(dev || (block->hop_nodes && block->hop_edges), NGNN_E_INVALID, "sage_agg1: the block carries no extents");
  StepPlan pl;
  NGNN_REQUIRE(make_plan(model, block->num_hops, max_hop_nodes, max_hop_edges, dev ? nullptr : block->hop_nodes,
                         dev ? nullptr : block->hop_edges, pl),
               NGNN_E_INVALID, "sage_agg1: bad model / block extents");
  NGNN_REQUIRE(ws_bytes >= pl.total, NGNN_E_WORKSPACE, "sage_agg1: workspace too small (%zu < %zu)", ws_bytes, pl.total);
  NGNN_REQUIRE(ld_table >= model->in_dim, NGNN_E_INVALID, "sage_agg1: ld_table < in_dim");
  char* base = reinterpret_cast<char*>(align_up(reinterpret_cast<size_t>(ws), 256));
  return agg1_launch(block, pl, table, ld_table, base, buffer, as_stream(stream));
}

int32_t ngnn_sage_prep_weights(const ngnn_sage_model_t* model, const float* params, int32_t num_hops, const int64_t* max_hop_nodes,
                               const int64_t* max_hop_edges, void* ws, size_t ws_bytes, ngnn_stream_t stream) {
  NGNN_REQUIRE(model && params && ws && max_hop_nodes && max_hop_edges, NGNN_E_INVALID, "sage_prep_weights: null pointer");
  StepPlan pl;
  NGNN_REQUIRE(make_plan(model, num_hops, max_hop_nodes, max_hop_edges, nullptr, nullptr, pl), NGNN_E_INVALID,
               "sage_prep_weights: bad model / capacities");
  NGNN_REQUIRE(ws_bytes >= pl.total, NGNN_E_WORKSPACE, "sage_prep_weights: workspace too small (%zu < %zu)", ws_bytes, pl.total);
  char* base = reinterpret_cast<char*>(align_up(reinterpret_cast<size_t>(ws), 256));
  bool f_ok[16], d_ok[16];
  const int32_t rc = prep_weights_collect(pl, params, true, base, f_ok, d_ok);
  if (rc != NGNN_OK) return rc;
  return prep_batch_launch(as_stream(stream));
}

int32_t ngnn_sage_step(const ngnn_sage_model_t* model, const float* params, float* grads, const ngnn_block_t* block,
                       const int64_t* max_hop_nodes, const int64_t* max_hop_edges, const float* table, int64_t ld_table,
                       const int64_t* target_global, const int64_t* label_global, uint64_t drop_seed, uint64_t drop_offset,
                       float* stats, float* logits_out, int64_t ld_logits, void* ws, size_t ws_bytes, ngnn_stream_t stream) {
  return sage_step_impl(model, params, grads, block, max_hop_nodes, max_hop_edges, table, ld_table, target_global, label_global,
                        drop_seed, drop_offset, stats, logits_out, ld_logits, ws, ws_bytes, stream, 0, nullptr, 0);
}

int32_t ngnn_sage_forward(const ngnn_sage_model_t* model, const float* params, const ngnn_block_t* block,
                          const int64_t* max_hop_nodes, const int64_t* max_hop_edges, const float* table, int64_t ld_table,
                          uint64_t drop_seed, uint64_t drop_offset, float* logits_out, int64_t ld_logits, void* ws,
                          size_t ws_bytes, ngnn_stream_t stream) {
  NGNN_REQUIRE(logits_out != nullptr, NGNN_E_INVALID, "sage_forward: logits_out is null");
  return sage_step_impl(model, params, nullptr, block, max_hop_nodes, max_hop_edges, table, ld_table, nullptr, nullptr, drop_seed,
                        drop_offset, nullptr, logits_out, ld_logits, ws, ws_bytes, stream, 1, nullptr, 0);
}

int32_t ngnn_sage_backward(const ngnn_sage_model_t* model, const float* params, float* grads, const ngnn_block_t* block,
                           const int64_t* max_hop_nodes, const int64_t* max_hop_edges, const float* table, int64_t ld_table,
                           const float* dlogits, int64_t ld_dlogits, void* ws, size_t ws_bytes, ngnn_stream_t stream) {
  return sage_step_impl(model, params, grads, block, max_hop_nodes, max_hop_edges, table, ld_table, nullptr, nullptr, 0, 0,
                        nullptr, nullptr, 0, ws, ws_bytes, stream, 2, dlogits, ld_dlogits);
}

int32_t ngnn_set_step_overlap(int32_t on) { g_use_aux = on != 0; return NGNN_OK; }

}  // extern "C"
