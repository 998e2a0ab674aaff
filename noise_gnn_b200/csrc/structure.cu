// structure.cu — block structure kernels: COO -> CSR (stable by destination), CSR transpose
// (CSC by source, for the atomic-free backward), CSR -> COO export, row gather.
//
// Replaces the COO->CSC conversion and relabelled edge_index handling that PyG performs around
// SAGEConv / NeighborLoader (reference call sites src/models/layers/sage.py:34, src/pipeline.py:75-83).
// The stable key sort itself is cub::DeviceRadixSort (a device-wide primitive, like a scan);
// everything domain-specific (key build, rowptr by binary search, relabel, export) is ours.
#include "common.cuh"
#include <cub/device/device_radix_sort.cuh>

namespace ngnn {

static int key_bits(int64_t n) {
  int b = 1;
  while ((1LL << b) < n && b < 31) ++b;
  return b;
}

__global__ void k_coo_keys(const int64_t* __restrict__ dst, int64_t e, int32_t* __restrict__ keys,
                           int32_t* __restrict__ vals) {
  int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p < e) { keys[p] = (int32_t)dst[p]; vals[p] = (int32_t)p; }
}

__global__ void k_permute_src(const int64_t* __restrict__ src, const int32_t* __restrict__ perm, int64_t e,
                              int32_t* __restrict__ col) {
  int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p < e) col[p] = (int32_t)src[perm[p]];
}

// ptr[i] = number of sorted keys < i  (i = 0..n)
__global__ void k_ptr_from_sorted(const int32_t* __restrict__ keys, int64_t e, int64_t n, int32_t* __restrict__ ptr) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i > n) return;
  int64_t lo = 0, hi = e;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (keys[mid] < (int32_t)i) lo = mid + 1; else hi = mid;
  }
  ptr[i] = (int32_t)lo;
}

// destination row of CSR position p: largest i with rowptr[i] <= p
__device__ __forceinline__ int32_t row_of(const int32_t* __restrict__ rowptr, int64_t n_rows, int32_t p) {
  int64_t lo = 0, hi = n_rows;  // invariant: rowptr[lo] <= p < rowptr[hi]
  while (hi - lo > 1) {
    int64_t mid = (lo + hi) >> 1;
    if (rowptr[mid] <= p) lo = mid; else hi = mid;
  }
  return (int32_t)lo;
}

__global__ void k_transpose_keys(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                 int64_t n_rows, int64_t e, int32_t* __restrict__ keys,
                                 int32_t* __restrict__ vals, int32_t* __restrict__ dst_of) {
  int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= e) return;
  keys[p] = col[p];
  vals[p] = (int32_t)p;
  dst_of[p] = row_of(rowptr, n_rows, (int32_t)p);
}

__global__ void k_permute_i32(const int32_t* __restrict__ src, const int32_t* __restrict__ perm, int64_t e,
                              int32_t* __restrict__ out) {
  int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p < e) out[p] = src[perm[p]];
}

__global__ void k_csr_to_coo(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n_rows,
                             int64_t e, int64_t* __restrict__ ei) {
  int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= e) return;
  ei[p] = col[p];
  ei[e + p] = row_of(rowptr, n_rows, (int32_t)p);
}

template <bool VEC>
__global__ void k_gather_rows(const float* __restrict__ table, int64_t ldt, const int32_t* __restrict__ idx,
                              int64_t n, int64_t F, float* __restrict__ out, int64_t ldo) {
  int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= n) return;
  const float* s = table + (int64_t)idx[warp] * ldt;
  float* d = out + warp * ldo;
  if (VEC) {
    int64_t f4 = F >> 2;
    const float4* s4 = reinterpret_cast<const float4*>(s);
    float4* d4 = reinterpret_cast<float4*>(d);
    for (int64_t c = lane; c < f4; c += 32) d4[c] = ldg_nc_f4(s4 + c);
  } else {
    for (int64_t c = lane; c < F; c += 32) d[c] = __ldg(s + c);
  }
}

struct SortWs {
  int32_t *keys_in, *keys_out, *vals_in, *aux;
  void* cub_tmp;
  size_t cub_bytes;
};

static size_t sort_cub_bytes(int64_t e) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const int32_t*)nullptr, (int32_t*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, (int)e, 0, 32);
  return bytes;
}

static size_t sort_ws_bytes(int64_t e) {
  size_t a = align_up((size_t)(e > 0 ? e : 1) * sizeof(int32_t), 256);
  return 4 * a + align_up(sort_cub_bytes(e > 0 ? e : 1), 256) + 256;
}

static bool carve(SortWs& w, void* ws, size_t ws_bytes, int64_t e) {
  if (ws_bytes < sort_ws_bytes(e)) return false;
  size_t a = align_up((size_t)(e > 0 ? e : 1) * sizeof(int32_t), 256);
  char* p = reinterpret_cast<char*>(align_up(reinterpret_cast<size_t>(ws), 256));
  w.keys_in = (int32_t*)p; p += a;
  w.keys_out = (int32_t*)p; p += a;
  w.vals_in = (int32_t*)p; p += a;
  w.aux = (int32_t*)p; p += a;
  w.cub_tmp = p;
  w.cub_bytes = sort_cub_bytes(e > 0 ? e : 1);
  return true;
}

}  // namespace ngnn

using namespace ngnn;

extern "C" {

size_t ngnn_coo_to_csr_workspace_bytes(int64_t e, int64_t /*n_rows*/) { return sort_ws_bytes(e); }
size_t ngnn_csr_transpose_workspace_bytes(int64_t e, int64_t /*n_cols*/) { return sort_ws_bytes(e); }

int32_t ngnn_coo_to_csr(const int64_t* src, const int64_t* dst, int64_t e, int64_t n_rows, int32_t* rowptr,
                        int32_t* col, int32_t* perm, void* ws, size_t ws_bytes, ngnn_stream_t stream) {
  NGNN_REQUIRE(e >= 0 && n_rows >= 0, NGNN_E_INVALID, "coo_to_csr: negative size");
  NGNN_REQUIRE(e < (1LL << 31) && n_rows < (1LL << 31) - 1, NGNN_E_UNSUPPORTED, "coo_to_csr: int32 index range exceeded");
  NGNN_REQUIRE(rowptr, NGNN_E_INVALID, "coo_to_csr: rowptr is null");
  NGNN_REQUIRE(e == 0 || (src && dst && col && perm), NGNN_E_INVALID, "coo_to_csr: null pointer");
  cudaStream_t st = as_stream(stream);
  const int T = 256;
  if (e > 0) {
    SortWs w;
    NGNN_REQUIRE(ws && carve(w, ws, ws_bytes, e), NGNN_E_WORKSPACE, "coo_to_csr: workspace too small (%zu < %zu)",
                 ws_bytes, sort_ws_bytes(e));
    k_coo_keys<<<(unsigned)ceil_div(e, T), T, 0, st>>>(dst, e, w.keys_in, w.vals_in);
    NGNN_LAUNCH_CHECK();
    size_t cb = w.cub_bytes;
    NGNN_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_tmp, cb, (const int32_t*)w.keys_in, w.keys_out,
                                              (const int32_t*)w.vals_in, perm, (int)e, 0, key_bits(n_rows), st));
    count_launches(2 + (key_bits(n_rows) + 7) / 8);   // cub onesweep: histogram, scan, one pass per 8 key bits
    k_permute_src<<<(unsigned)ceil_div(e, T), T, 0, st>>>(src, perm, e, col);
    NGNN_LAUNCH_CHECK();
    k_ptr_from_sorted<<<(unsigned)ceil_div(n_rows + 1, T), T, 0, st>>>(w.keys_out, e, n_rows, rowptr);
    NGNN_LAUNCH_CHECK();
  } else {
    NGNN_CUDA(cudaMemsetAsync(rowptr, 0, (size_t)(n_rows + 1) * sizeof(int32_t), st));
  }
  return NGNN_OK;
}

int32_t ngnn_csr_transpose(const int32_t* rowptr, const int32_t* col, int64_t n_rows, int64_t e_limit,
                           int64_t n_cols, int32_t* colptr_t, int32_t* row_t, int32_t* perm_t, void* ws,
                           size_t ws_bytes, ngnn_stream_t stream) {
  NGNN_REQUIRE(e_limit >= 0 && n_rows >= 0 && n_cols >= 0, NGNN_E_INVALID, "csr_transpose: negative size");
  NGNN_REQUIRE(colptr_t, NGNN_E_INVALID, "csr_transpose: colptr_t is null");
  NGNN_REQUIRE(e_limit == 0 || (rowptr && col && row_t && perm_t), NGNN_E_INVALID, "csr_transpose: null pointer");
  cudaStream_t st = as_stream(stream);
  const int T = 256;
  if (e_limit > 0) {
    SortWs w;
    NGNN_REQUIRE(ws && carve(w, ws, ws_bytes, e_limit), NGNN_E_WORKSPACE,
                 "csr_transpose: workspace too small (%zu < %zu)", ws_bytes, sort_ws_bytes(e_limit));
    k_transpose_keys<<<(unsigned)ceil_div(e_limit, T), T, 0, st>>>(rowptr, col, n_rows, e_limit, w.keys_in,
                                                                  w.vals_in, w.aux);
    NGNN_LAUNCH_CHECK();
    size_t cb = w.cub_bytes;
    NGNN_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_tmp, cb, (const int32_t*)w.keys_in, w.keys_out,
                                              (const int32_t*)w.vals_in, perm_t, (int)e_limit, 0,
                                              key_bits(n_cols), st));
    count_launches(2 + (key_bits(n_cols) + 7) / 8);
    k_permute_i32<<<(unsigned)ceil_div(e_limit, T), T, 0, st>>>(w.aux, perm_t, e_limit, row_t);
    NGNN_LAUNCH_CHECK();
    k_ptr_from_sorted<<<(unsigned)ceil_div(n_cols + 1, T), T, 0, st>>>(w.keys_out, e_limit, n_cols, colptr_t);
    NGNN_LAUNCH_CHECK();
  } else {
    NGNN_CUDA(cudaMemsetAsync(colptr_t, 0, (size_t)(n_cols + 1) * sizeof(int32_t), st));
  }
  return NGNN_OK;
}

int32_t ngnn_csr_to_coo(const int32_t* rowptr, const int32_t* col, int64_t n_rows, int64_t e, int64_t* edge_index,
                        ngnn_stream_t stream) {
  NGNN_REQUIRE(e >= 0 && n_rows >= 0, NGNN_E_INVALID, "csr_to_coo: negative size");
  if (e == 0) return NGNN_OK;
  NGNN_REQUIRE(rowptr && col && edge_index, NGNN_E_INVALID, "csr_to_coo: null pointer");
  k_csr_to_coo<<<(unsigned)ceil_div(e, 256), 256, 0, as_stream(stream)>>>(rowptr, col, n_rows, e, edge_index);
  NGNN_LAUNCH_CHECK();
  return NGNN_OK;
}

int32_t ngnn_gather_rows(const float* table, int64_t ld_table, const int32_t* idx, int64_t n, int64_t F, float* out,
                         int64_t ld_out, ngnn_stream_t stream) {
  NGNN_REQUIRE(n >= 0 && F >= 0, NGNN_E_INVALID, "gather_rows: negative size");
  if (n == 0 || F == 0) return NGNN_OK;
  NGNN_REQUIRE(table && idx && out, NGNN_E_INVALID, "gather_rows: null pointer");
  NGNN_REQUIRE(ld_table >= F && ld_out >= F, NGNN_E_INVALID, "gather_rows: leading dimension < F");
  bool vec = (F % 4 == 0) && (ld_table % 4 == 0) && (ld_out % 4 == 0) && is_aligned(table, 16) && is_aligned(out, 16);
  const int T = 256;
  unsigned grid = (unsigned)ceil_div(n * 32, T);
  if (vec) k_gather_rows<true><<<grid, T, 0, as_stream(stream)>>>(table, ld_table, idx, n, F, out, ld_out);
  else     k_gather_rows<false><<<grid, T, 0, as_stream(stream)>>>(table, ld_table, idx, n, F, out, ld_out);
  NGNN_LAUNCH_CHECK();
  return NGNN_OK;
}

}  // extern "C"
