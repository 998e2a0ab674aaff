// gemm.cu — C-ABI entry points of the dense half of SAGEConv: K-GEMM (forward projection,
// PyG ops K4-K8 of SURVEY §2.3), K-DGRAD and K-WGRAD (K11).  Reference call site of the op
// being replaced: torch_geometric.nn.SAGEConv.forward via src/models/layers/sage.py:34.
//
// Dispatch: the tcgen05/TMA path (gemm_tc.cuh, 3xTF32 split for fp32-grade accuracy) when the
// operands are TMA-addressable (F % 4 == 0, 16-byte aligned bases and leading dimensions),
// the SIMT fp32 path (gemm_simt.cuh) otherwise.  Both are device code in this library; there
// is no host fallback.
#include "common.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"

namespace ngnn {

static int g_force_simt = 0;   // ngnn_set_gemm_path: 1 => always take the SIMT kernels (tests, A/B timing)

static int32_t wgrad_splits(int64_t n, int64_t F, int64_t O) {
  const int64_t tiles = ceil_div(O, SG_BM) * ceil_div(F, SG_BN);
  int64_t s = ceil_div(4 * kNumSMs, tiles);            // aim for ~4 CTAs per SM
  const int64_t max_s = ceil_div(n, 8 * SG_BK);          // at least 8 k-tiles per slice
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  if (s > 1024) s = 1024;
  return (int32_t)s;
}
static int32_t colsum_slices(int64_t n, int64_t O) {
  // ~2 CTAs per SM over all 32-column strips together; the fixed-order reduce over the slices is one dependent chain
  // per output column, so fewer slices is also a shorter tail (296 slices cost 24 us in that reduce)
  const int64_t strips = ceil_div(O, 32);
  int64_t s = ceil_div(2 * kNumSMs, strips);
  const int64_t max_s = ceil_div(n, 64);
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  return (int32_t)s;
}

int32_t prep_weights_impl(int32_t mode, const float* w_l, const float* w_r, int64_t F, int64_t O, void* ws, size_t ws_bytes,
                          cudaStream_t st) {
  if (g_force_simt || get_encode_fn() == nullptr) return NGNN_E_UNSUPPORTED;
  return mode == 0 ? tc_prep_fwd(w_l, w_r, w_l != nullptr, w_r != nullptr, F, O, ws, ws_bytes, st)
                   : tc_prep_dgrad(w_l, w_r, w_l != nullptr, w_r != nullptr, F, O, ws, ws_bytes, st);
}

static thread_local PrepBatch t_prep_batch;

void prep_batch_begin() { t_prep_batch.n_jobs = 0; t_prep_batch.first_block[0] = 0; }

int32_t prep_batch_add(int32_t mode, const float* w_l, const float* w_r, int64_t F, int64_t O, void* ws, size_t ws_bytes) {
  if (g_force_simt || get_encode_fn() == nullptr) return NGNN_E_UNSUPPORTED;
  PrepBatch& b = t_prep_batch;
  if (b.n_jobs >= PREP_MAX_JOBS) return NGNN_E_UNSUPPORTED;
  PrepParams pp{};
  const int32_t rc = mode != 1 ? tc_prep_fwd(w_l, w_r, w_l != nullptr, w_r != nullptr, F, O, ws, ws_bytes, nullptr, &pp, mode == 2)
                               : tc_prep_dgrad(w_l, w_r, w_l != nullptr, w_r != nullptr, F, O, ws, ws_bytes, nullptr, &pp);
  if (rc != NGNN_OK) return rc;
  b.job[b.n_jobs] = pp;
  b.first_block[b.n_jobs + 1] = b.first_block[b.n_jobs] + (int32_t)ceil_div((int64_t)pp.rows_out * pp.Kpack, 256);
  ++b.n_jobs;
  return NGNN_OK;
}

int32_t prep_batch_launch(cudaStream_t st) {
  const PrepBatch& b = t_prep_batch;
  if (b.n_jobs == 0) return NGNN_OK;
  launch_chain(k_prep_weights_batch, dim3((unsigned)b.first_block[b.n_jobs]), dim3(256), 0, st, b);
  NGNN_LAUNCH_CHECK();
  return NGNN_OK;
}

int32_t gemm_fwd_impl(const float* a_l, int64_t ld_al, const float* a_r, int64_t ld_ar, const float* w_l, const float* w_r,
                      const float* bias, int64_t n, int64_t F, int64_t O, int32_t act, float drop_p, uint64_t seed,
                      uint64_t offset, float* out, int64_t ld_out, int32_t* path, void* ws, size_t ws_bytes, cudaStream_t st,
                      bool prepped, const int32_t* n_dev, const StepCtl* ctl, uint32_t ctl_layer, bool concat_k) {
  NGNN_REQUIRE(n >= 0 && F >= 0 && O >= 0, NGNN_E_INVALID, "gemm_fwd: negative size");
  NGNN_REQUIRE(act == NGNN_ACT_NONE || act == NGNN_ACT_RELU, NGNN_E_INVALID, "gemm_fwd: unknown activation %d", act);
  NGNN_REQUIRE(drop_p >= 0.f && drop_p < 1.f, NGNN_E_INVALID, "gemm_fwd: dropout p=%f outside [0,1)", (double)drop_p);
  if (path) *path = 0;
  if (n == 0 || O == 0) return NGNN_OK;
  NGNN_REQUIRE(out && ld_out >= O, NGNN_E_INVALID, "gemm_fwd: bad output");
  NGNN_REQUIRE(a_l == nullptr || w_l != nullptr, NGNN_E_INVALID, "gemm_fwd: a_l without w_l");
  NGNN_REQUIRE(a_r == nullptr || w_r != nullptr, NGNN_E_INVALID, "gemm_fwd: a_r without w_r");
  NGNN_REQUIRE(a_l == nullptr || ld_al >= F, NGNN_E_INVALID, "gemm_fwd: ld_al < F");
  NGNN_REQUIRE(a_r == nullptr || ld_ar >= F, NGNN_E_INVALID, "gemm_fwd: ld_ar < F");

  if (!g_force_simt) {
    int32_t rc = tc_gemm_fwd(a_l, ld_al, a_r, ld_ar, w_l, w_r, bias, n, F, O, act, drop_p, seed, offset, out, ld_out,
                             ws, ws_bytes, st, prepped, n_dev, ctl, ctl_layer, concat_k);
    if (rc == NGNN_OK) { if (path) *path = 1; return NGNN_OK; }
    if (rc != NGNN_E_UNSUPPORTED) return rc;
  }
  NGNN_REQUIRE(!concat_k, NGNN_E_UNSUPPORTED, "gemm_fwd: the one-contraction form needs the tensor-core path");

  SimtGemmParams p{};
  if (a_l) { p.A1 = {a_l, ld_al, 1}; p.B1 = {w_l, F, 1}; p.K1 = F; }
  if (a_r) { p.A2 = {a_r, ld_ar, 1}; p.B2 = {w_r, F, 1}; p.K2 = F; }
  p.M = n; p.N = O; p.C = out; p.ldc = ld_out; p.bias = bias; p.act = act; p.drop_p = drop_p;
  p.M_dev = n_dev; p.ctl = ctl; p.ctl_layer = ctl_layer;
  p.seed_lo = (uint32_t)seed; p.seed_hi = (uint32_t)(seed >> 32);
  p.off_lo = (uint32_t)offset; p.off_hi = (uint32_t)(offset >> 32);
  return launch_simt_gemm(p, 1, st);
}

int32_t dgrad_impl(const float* dy, int64_t ld_dy, const float* w_l, const float* w_r, const int32_t* rowptr, int64_t n,
                   int64_t F, int64_t O, float* dmean_scaled, int64_t ld_dmean, float* dx_root, int64_t ld_root, void* ws,
                   size_t ws_bytes, cudaStream_t st, bool prepped, const int32_t* n_dev) {
  NGNN_REQUIRE(n >= 0 && F >= 0 && O >= 0, NGNN_E_INVALID, "dgrad: negative size");
  if (n == 0 || F == 0) return NGNN_OK;
  NGNN_REQUIRE(dy && ld_dy >= O, NGNN_E_INVALID, "dgrad: bad dy");
  NGNN_REQUIRE(dmean_scaled == nullptr || (w_l && ld_dmean >= F), NGNN_E_INVALID, "dgrad: bad dmean output");
  NGNN_REQUIRE(dx_root == nullptr || (w_r && ld_root >= F), NGNN_E_INVALID, "dgrad: bad dx_root output");
  // out[i,f] = sum_o dy[i,o] * W[o,f]  : A = dy (K = O contiguous), B(n=f,k=o) = W[o*F + f]
  if (!g_force_simt) {
    int32_t rc = tc_gemm_dgrad(dy, ld_dy, w_l, w_r, rowptr, n, F, O, dmean_scaled, ld_dmean, dx_root, ld_root, ws,
                               ws_bytes, st, prepped, n_dev);
    if (rc != NGNN_E_UNSUPPORTED) return rc;
  }
  if (dmean_scaled) {
    SimtGemmParams p{};
    p.A1 = {dy, ld_dy, 1}; p.B1 = {w_l, 1, F}; p.K1 = O;
    p.M = n; p.M_dev = n_dev; p.N = F; p.C = dmean_scaled; p.ldc = ld_dmean; p.rowptr_scale = rowptr;
    int32_t rc = launch_simt_gemm(p, 1, st);
    if (rc != NGNN_OK) return rc;
  }
  if (dx_root) {
    SimtGemmParams p{};
    p.A1 = {dy, ld_dy, 1}; p.B1 = {w_r, 1, F}; p.K1 = O;
    p.M = n; p.M_dev = n_dev; p.N = F; p.C = dx_root; p.ldc = ld_root;
    int32_t rc = launch_simt_gemm(p, 1, st);
    if (rc != NGNN_OK) return rc;
  }
  return NGNN_OK;
}

static size_t colsum_ws_bytes(int64_t n, int64_t O) { return align_up((size_t)colsum_slices(n, O) * (size_t)O * sizeof(float), 256); }

// n_dev (optional): device-side number of rows; n is then the capacity every launch is sized for.
int32_t wgrad_impl(const float* dy, int64_t ld_dy, const float* a_l, int64_t ld_al, const float* a_r, int64_t ld_ar, int64_t n,
                   const int32_t* n_dev, int64_t F, int64_t O, float* dw_l, float* dw_r, float* db, int32_t accumulate,
                   void* ws, size_t ws_bytes, cudaStream_t st) {
  NGNN_REQUIRE(n >= 0 && F >= 0 && O >= 0, NGNN_E_INVALID, "wgrad: negative size");
  if (O == 0) return NGNN_OK;
  if (n == 0) {  // empty block: gradients are zero
    if (!accumulate) {
      if (dw_l && F) NGNN_CUDA(cudaMemsetAsync(dw_l, 0, (size_t)O * F * sizeof(float), st));
      if (dw_r && F) NGNN_CUDA(cudaMemsetAsync(dw_r, 0, (size_t)O * F * sizeof(float), st));
      if (db) NGNN_CUDA(cudaMemsetAsync(db, 0, (size_t)O * sizeof(float), st));
    }
    return NGNN_OK;
  }
  NGNN_REQUIRE(dy && ld_dy >= O, NGNN_E_INVALID, "wgrad: bad dy");
  NGNN_REQUIRE(dw_l == nullptr || (a_l && ld_al >= F), NGNN_E_INVALID, "wgrad: dw_l without a_l");
  NGNN_REQUIRE(dw_r == nullptr || (a_r && ld_ar >= F), NGNN_E_INVALID, "wgrad: dw_r without a_r");
  NGNN_REQUIRE(ws && ws_bytes >= ngnn_sage_wgrad_workspace_bytes(n, F, O), NGNN_E_WORKSPACE,
               "wgrad: workspace too small (%zu < %zu)", ws_bytes, ngnn_sage_wgrad_workspace_bytes(n, F, O));
  float* cpart = reinterpret_cast<float*>(align_up(reinterpret_cast<size_t>(ws), 256));
  float* part = cpart + colsum_ws_bytes(n, O) / sizeof(float);
  const size_t part_bytes = ws_bytes - 256 - colsum_ws_bytes(n, O);

  // dW[o,f] = sum_i dy[i,o] * a[i,f] : A(m=o,k=i) = dy[i*ld+o], B(n=f,k=i) = a[i*ld+f]
  bool done = false, db_done = false;
  if (!g_force_simt && F > 0 && (dw_l || dw_r)) {
    int32_t rc = tc_gemm_wgrad(dy, ld_dy, a_l, ld_al, a_r, ld_ar, n, n_dev, F, O, dw_l, dw_r, db, &db_done, accumulate, part,
                               part_bytes, st);
    if (rc == NGNN_OK) done = true;
    else if (rc != NGNN_E_UNSUPPORTED) return rc;
  }
  if (!done && (dw_l || dw_r)) {
    const int32_t S = wgrad_splits(n, F, O);
    const float* as[2] = {a_l, a_r};
    const int64_t lds[2] = {ld_al, ld_ar};
    float* dws[2] = {dw_l, dw_r};
    for (int w = 0; w < 2; ++w) {
      if (!dws[w] || F == 0) continue;
      SimtGemmParams p{};
      p.A1 = {dy, 1, ld_dy}; p.B1 = {as[w], 1, lds[w]}; p.K1 = n; p.K1_dev = n_dev;
      p.M = O; p.N = F; p.ldc = F;
      if (S == 1 && !accumulate) {
        p.C = dws[w]; p.split_stride = 0;
        int32_t rc = launch_simt_gemm(p, 1, st);
        if (rc != NGNN_OK) return rc;
      } else {
        p.C = part; p.split_stride = O * F;
        int32_t rc = launch_simt_gemm(p, S, st);
        if (rc != NGNN_OK) return rc;
        k_reduce_partials<<<(unsigned)ceil_div(O * F, 256), 256, 0, st>>>(part, O * F, S, O * F, dws[w], accumulate);
        NGNN_LAUNCH_CHECK();
      }
    }
  }
  if (db && !db_done) {
    const int32_t Cs = colsum_slices(n, O);
    dim3 grid((unsigned)ceil_div(O, 32), (unsigned)Cs);
    k_colsum_partial<<<grid, 256, 0, st>>>(dy, ld_dy, Ext{n_dev, n}, O, cpart);
    NGNN_LAUNCH_CHECK();
    k_reduce_partials<<<(unsigned)ceil_div(O, 256), 256, 0, st>>>(cpart, O, Cs, O, db, accumulate);
    NGNN_LAUNCH_CHECK();
  }
  return NGNN_OK;
}

}  // namespace ngnn

using namespace ngnn;

extern "C" {

int32_t ngnn_sage_gemm_fwd(const float* a_l, int64_t ld_al, const float* a_r, int64_t ld_ar, const float* w_l,
                           const float* w_r, const float* bias, int64_t n, int64_t F, int64_t O, int32_t act,
                           float drop_p, uint64_t seed, uint64_t offset, float* out, int64_t ld_out, int32_t* path,
                           void* ws, size_t ws_bytes, ngnn_stream_t stream) {
  return gemm_fwd_impl(a_l, ld_al, a_r, ld_ar, w_l, w_r, bias, n, F, O, act, drop_p, seed, offset, out, ld_out, path, ws,
                       ws_bytes, as_stream(stream), false, nullptr, nullptr, 0, false);
}

int32_t ngnn_sage_dgrad(const float* dy, int64_t ld_dy, const float* w_l, const float* w_r, const int32_t* rowptr,
                        int64_t n, int64_t F, int64_t O, float* dmean_scaled, int64_t ld_dmean, float* dx_root,
                        int64_t ld_root, void* ws, size_t ws_bytes, ngnn_stream_t stream) {
  return dgrad_impl(dy, ld_dy, w_l, w_r, rowptr, n, F, O, dmean_scaled, ld_dmean, dx_root, ld_root, ws, ws_bytes,
                    as_stream(stream), false, nullptr);
}

int32_t ngnn_set_gemm_tile(int32_t bn_max) { g_tc_bn_max = bn_max; return NGNN_OK; }   // reached through ngnn_set_tuning(4, .)
int32_t ngnn_set_gemm_ts(int32_t on) { g_tc_ts = on; return NGNN_OK; }                  // reached through ngnn_set_tuning(6, .)
int32_t ngnn_set_wgrad_splits(int32_t s) { g_tc_wgrad_splits = s; return NGNN_OK; }     // reached through ngnn_set_tuning(8, .)

int32_t ngnn_debug_set_trace(void* device_buffer) {
  g_tc_trace = reinterpret_cast<long long*>(device_buffer);
  return NGNN_OK;
}

int32_t ngnn_set_gemm_path(int32_t mode) {
  NGNN_REQUIRE(mode == 0 || mode == 1, NGNN_E_INVALID, "set_gemm_path: mode must be 0 (auto) or 1 (force SIMT)");
  g_force_simt = mode;
  return NGNN_OK;
}

size_t ngnn_sage_gemm_workspace_bytes(int64_t F, int64_t O) { return (F > 0 && O > 0) ? tc_fwd_ws_bytes(F, O) : 256; }
size_t ngnn_sage_dgrad_workspace_bytes(int64_t F, int64_t O) { return (F > 0 && O > 0) ? tc_dgrad_ws_bytes(F, O) : 256; }

size_t ngnn_sage_wgrad_workspace_bytes(int64_t n, int64_t F, int64_t O) {
  if (n <= 0 || F < 0 || O <= 0) return 256;
  const size_t simt = align_up((size_t)wgrad_splits(n, F, O) * (size_t)O * (size_t)F * sizeof(float), 256);
  const size_t tc = F >= 1 ? tc_wgrad_ws_bytes(n, F, O) : 0;
  return colsum_ws_bytes(n, O) + (simt > tc ? simt : tc) + 512;
}

int32_t ngnn_sage_wgrad(const float* dy, int64_t ld_dy, const float* a_l, int64_t ld_al, const float* a_r,
                        int64_t ld_ar, int64_t n, int64_t F, int64_t O, float* dw_l, float* dw_r, float* db,
                        int32_t accumulate, void* ws, size_t ws_bytes, ngnn_stream_t stream) {
  return wgrad_impl(dy, ld_dy, a_l, ld_al, a_r, ld_ar, n, nullptr, F, O, dw_l, dw_r, db, accumulate, ws, ws_bytes,
                    as_stream(stream));
}

int32_t ngnn_act_bwd(const float* dh, int64_t ld_dh, const float* h, int64_t ld_h, int64_t n, int64_t O, float scale,
                     float* dz, int64_t ld_dz, ngnn_stream_t stream) {
  NGNN_REQUIRE(n >= 0 && O >= 0, NGNN_E_INVALID, "act_bwd: negative size");
  if (n == 0 || O == 0) return NGNN_OK;
  NGNN_REQUIRE(dh && h && dz, NGNN_E_INVALID, "act_bwd: null pointer");
  k_act_bwd<<<(unsigned)ceil_div(n * O, 256), 256, 0, as_stream(stream)>>>(dh, ld_dh, h, ld_h, n, O, scale, dz, ld_dz);
  NGNN_LAUNCH_CHECK();
  return NGNN_OK;
}

}  // extern "C"
