// sampler.cu — K-SAMPLE / K-RELABEL / K-CSR: GPU fan-out neighbour sampler that emits CSR
// message-flow blocks.  Replaces torch_geometric.loader.NeighborLoader's per-batch call into
// pyg-lib neighbor_sample (reference call sites src/pipeline.py:75-83 ctor, :152 iteration;
// semantics restated in SURVEY §8 row A1): hop-by-hop expansion in discovery order, take-all when
// deg <= fanout, distinct uniform positions otherwise (Robert Floyd's subset algorithm), first-seen
// relabelling with the seeds first, edges emitted grouped by destination => CSR by construction.
//
// B200-first design: the graph (CSC), the feature table and two N-sized relabel maps stay resident
// in HBM; a block is built by a fixed sequence of launches with worst-case grids whose real extents
// live in device memory (`counts`), so sampling needs no host round trip and can be captured in a
// CUDA graph or run ahead on a side stream.  Draws are counter-based (Philox4x32-10 keyed by
// (seed) with counter (node, hop|draw/4, batch, epoch)): a block is a pure function of
// (seed, epoch, batch_idx, seeds), independent of launch geometry and GPU count.
// First-seen order is made deterministic with an atomicMin over candidate positions + a scan.
#include "common.cuh"
#include <cub/device/device_scan.cuh>
#include <limits.h>

namespace ngnn {

constexpr int kMaxFanout = 64;   // Floyd's set lives in per-thread local storage

struct SampleCaps {
  int64_t fr_max[16];   // worst-case frontier size per hop
  int64_t e_max[16];    // worst-case sampled edges per hop
  int64_t max_nodes, max_edges, fr_cap, e_cap;
};

static bool sample_caps(int32_t bs, const int32_t* fanouts, int32_t H, int64_t N, SampleCaps& c) {
  if (H < 1 || H > 15 || bs < 0) return false;
  int64_t fr = bs, nodes = bs, edges = 0, frc = 0, ec = 0;
  for (int h = 0; h < H; ++h) {
    if (fanouts[h] < 1) return false;
    c.fr_max[h] = fr;
    c.e_max[h] = fr * (int64_t)fanouts[h];
    if (c.e_max[h] >= (1LL << 31)) return false;
    edges += c.e_max[h];
    if (fr > frc) frc = fr;
    if (c.e_max[h] > ec) ec = c.e_max[h];
    fr = c.e_max[h] < N ? c.e_max[h] : N;     // new nodes discovered at hop h
    nodes += fr;
  }
  if (edges >= (1LL << 31)) return false;
  c.max_nodes = nodes < N + bs ? nodes : N + bs;
  c.max_edges = edges;
  c.fr_cap = frc;
  c.e_cap = ec;
  return true;
}

struct SampleWs {
  int32_t *local_of, *first_pos, *cnt, *off, *flag, *rank;
  void* cub_tmp;
  size_t cub_bytes;
};

static size_t scan_cub_bytes(int64_t n) {
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const int32_t*)nullptr, (int32_t*)nullptr, (int)n);
  return bytes;
}

static size_t sample_ws_layout(int64_t N, const SampleCaps& c, void* ws, SampleWs* out) {
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += align_up(bytes, 256); return r; };
  const size_t o_local = take((size_t)N * 4), o_first = take((size_t)N * 4);
  const size_t o_cnt = take((size_t)(c.fr_cap + 1) * 4), o_off = take((size_t)(c.fr_cap + 1) * 4);
  const size_t o_flag = take((size_t)(c.e_cap + 1) * 4), o_rank = take((size_t)(c.e_cap + 1) * 4);
  const size_t cb = scan_cub_bytes((c.e_cap > c.fr_cap ? c.e_cap : c.fr_cap) + 1);
  const size_t o_cub = take(cb);
  if (out) {
    char* b = reinterpret_cast<char*>(align_up(reinterpret_cast<size_t>(ws), 256));
    out->local_of = (int32_t*)(b + o_local); out->first_pos = (int32_t*)(b + o_first);
    out->cnt = (int32_t*)(b + o_cnt); out->off = (int32_t*)(b + o_off);
    out->flag = (int32_t*)(b + o_flag); out->rank = (int32_t*)(b + o_rank);
    out->cub_tmp = b + o_cub; out->cub_bytes = cb;
  }
  return o + 256;
}

__global__ void k_fill_i32(int32_t* p, int64_t n, int32_t v) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

__global__ void k_seed_init(const int64_t* __restrict__ seeds, int32_t bs, int32_t H, int32_t* __restrict__ n_id,
                            int32_t* __restrict__ local_of, int32_t* __restrict__ counts, int32_t* __restrict__ rowptr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < bs) {
    const int32_t g = (int32_t)seeds[i];
    n_id[i] = g;
    local_of[g] = i;
  }
  if (i == 0) { counts[0] = bs; counts[H + 1] = 0; rowptr[0] = 0; }
}

// cnt[i] = number of in-neighbours frontier node i will emit (0 for padding slots)
__global__ void k_count(const int32_t* __restrict__ colptr, const int32_t* __restrict__ n_id,
                        const int32_t* __restrict__ counts, int32_t h, int32_t fanout, int32_t replace,
                        int64_t fr_max, int32_t* __restrict__ cnt) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i > fr_max) return;
  const int32_t lo = h == 0 ? 0 : counts[h - 1], hi = counts[h];
  int32_t c = 0;
  if (i < hi - lo) {
    const int32_t v = n_id[lo + i];
    const int32_t d = __ldg(colptr + v + 1) - __ldg(colptr + v);
    c = replace ? (d > 0 ? fanout : 0) : min(d, fanout);
  }
  cnt[i] = c;
}

// One WARP per frontier node (fan-out <= 32): lane j owns draw j.  The positions are decided in registers (Floyd's
// subset algorithm run cooperatively: one shuffle + one vote per draw, no memory), then every lane reads ITS neighbour
// id and relabel slot at once — one round of dependent DRAM latencies per node instead of one per draw (the
// thread-per-node kernel below spent 35 us on the 512 seeds of a products batch, all of it latency).
// Same positions, same emission order and the same Philox stream as k_draw_serial / the C oracle.
__global__ void k_draw_warp(const int32_t* __restrict__ colptr, const int32_t* __restrict__ row,
                            const int32_t* __restrict__ n_id, int32_t* __restrict__ counts, int32_t h, int32_t H,
                            int32_t fanout, int32_t replace, int64_t fr_max, const int32_t* __restrict__ off,
                            uint32_t seed_lo, uint32_t seed_hi, uint32_t epoch, uint32_t batch_idx,
                            int32_t* __restrict__ rowptr, int32_t* __restrict__ col_global, int32_t* __restrict__ e_pos,
                            const int32_t* __restrict__ local_of, int32_t* __restrict__ first_pos) {
  const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const unsigned full = 0xffffffffu;
  const int32_t lo = h == 0 ? 0 : counts[h - 1], hi = counts[h];
  const int32_t e_base = counts[H + 1 + h];
  if (i == 0 && lane == 0) counts[H + 2 + h] = e_base + off[fr_max];
  if (i >= hi - lo) return;                                  // warp-uniform
  const int32_t v = n_id[lo + i];
  const int32_t beg = __ldg(colptr + v), d = __ldg(colptr + v + 1) - beg;
  const int32_t o = off[i];
  const int32_t k = replace ? (d > 0 ? fanout : 0) : min(d, fanout);
  if (lane == 0) rowptr[lo + i + 1] = e_base + o + k;
  if (k == 0) return;

  int32_t pos = lane;                                        // take all, stored order
  if (replace || d > fanout) {
    const Philox4 r = philox4x32_10((uint32_t)v, ((uint32_t)h << 16) | (uint32_t)(lane >> 2), batch_idx, epoch, seed_lo, seed_hi);
    const uint32_t w = (lane & 3) == 0 ? r.x : (lane & 3) == 1 ? r.y : (lane & 3) == 2 ? r.z : r.w;
    if (replace) {
      pos = (int32_t)mulhi32(w, (uint32_t)d);
    } else {
      // Robert Floyd: for jj = d-k .. d-1: t = U[0,jj]; take t unless already taken, else take jj
      int32_t mine = -1;
      for (int32_t j = 0; j < k; ++j) {
        const uint32_t wj = __shfl_sync(full, w, j);
        const int32_t jj = d - k + j;
        int32_t t = (int32_t)mulhi32(wj, (uint32_t)(jj + 1));
        if (__any_sync(full, lane < j && mine == t)) t = jj;
        if (lane == j) mine = t;
      }
      pos = mine;
    }
  }
  if (lane < k) {
    const int32_t g = __ldg(row + beg + pos);
    col_global[e_base + o + lane] = g;
    if (e_pos) e_pos[e_base + o + lane] = beg + pos;
    if (local_of[g] < 0) atomicMin(first_pos + g, o + lane);
  }
}

// One thread per frontier node (any fan-out): draw positions, emit global neighbour ids + CSC positions, close the
// CSR row, and vote (atomicMin) for the first candidate position of every not-yet-labelled neighbour.
__global__ void k_draw(const int32_t* __restrict__ colptr, const int32_t* __restrict__ row,
                       const int32_t* __restrict__ n_id, int32_t* __restrict__ counts, int32_t h, int32_t H,
                       int32_t fanout, int32_t replace, int64_t fr_max, const int32_t* __restrict__ off,
                       uint32_t seed_lo, uint32_t seed_hi, uint32_t epoch, uint32_t batch_idx,
                       int32_t* __restrict__ rowptr, int32_t* __restrict__ col_global, int32_t* __restrict__ e_pos,
                       const int32_t* __restrict__ local_of, int32_t* __restrict__ first_pos) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int32_t lo = h == 0 ? 0 : counts[h - 1], hi = counts[h];
  const int32_t e_base = counts[H + 1 + h];
  if (i == 0) counts[H + 2 + h] = e_base + off[fr_max];
  if (i >= hi - lo) return;
  const int32_t v = n_id[lo + i];
  const int32_t beg = __ldg(colptr + v), d = __ldg(colptr + v + 1) - beg;
  const int32_t o = off[i];
  const int32_t k = replace ? (d > 0 ? fanout : 0) : min(d, fanout);
  rowptr[lo + i + 1] = e_base + o + k;

  auto emit = [&](int32_t j, int32_t pos) {
    const int32_t g = __ldg(row + beg + pos);
    col_global[e_base + o + j] = g;
    if (e_pos) e_pos[e_base + o + j] = beg + pos;
    if (local_of[g] < 0) atomicMin(first_pos + g, o + j);
  };

  if (!replace && d <= fanout) {
    for (int32_t j = 0; j < d; ++j) emit(j, j);          // take all, stored order
    return;
  }
  Philox4 r{0, 0, 0, 0};
  if (replace) {
    for (int32_t j = 0; j < k; ++j) {
      if ((j & 3) == 0) r = philox4x32_10((uint32_t)v, ((uint32_t)h << 16) | (uint32_t)(j >> 2), batch_idx, epoch, seed_lo, seed_hi);
      const uint32_t w = (j & 3) == 0 ? r.x : (j & 3) == 1 ? r.y : (j & 3) == 2 ? r.z : r.w;
      emit(j, (int32_t)mulhi32(w, (uint32_t)d));
    }
    return;
  }
  // Robert Floyd: for jj = d-k .. d-1: t = U[0,jj]; take t unless already taken, else take jj
  int32_t S[kMaxFanout];
  for (int32_t j = 0; j < k; ++j) {
    if ((j & 3) == 0) r = philox4x32_10((uint32_t)v, ((uint32_t)h << 16) | (uint32_t)(j >> 2), batch_idx, epoch, seed_lo, seed_hi);
    const uint32_t w = (j & 3) == 0 ? r.x : (j & 3) == 1 ? r.y : (j & 3) == 2 ? r.z : r.w;
    const int32_t jj = d - k + j;
    int32_t t = (int32_t)mulhi32(w, (uint32_t)(jj + 1));
    for (int32_t q = 0; q < j; ++q) if (S[q] == t) { t = jj; break; }
    S[j] = t;
    emit(j, t);
  }
}

__global__ void k_flag(const int32_t* __restrict__ col_global, const int32_t* __restrict__ counts, int32_t h, int32_t H,
                       int64_t e_max, const int32_t* __restrict__ local_of, const int32_t* __restrict__ first_pos,
                       int32_t* __restrict__ flag) {
  const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p > e_max) return;
  const int32_t e_base = counts[H + 1 + h], e_h = counts[H + 2 + h] - e_base;
  int32_t f = 0;
  if (p < e_h) {
    const int32_t g = col_global[e_base + p];
    f = (local_of[g] < 0 && first_pos[g] == (int32_t)p) ? 1 : 0;
  }
  flag[p] = f;
}

__global__ void k_assign(const int32_t* __restrict__ col_global, int32_t* __restrict__ counts, int32_t h, int32_t H,
                         int64_t e_max, const int32_t* __restrict__ flag, const int32_t* __restrict__ rank,
                         int32_t* __restrict__ n_id, int32_t* __restrict__ local_of) {
  const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int32_t n_prev = counts[h];
  if (p == 0) counts[h + 1] = n_prev + rank[e_max];
  const int32_t e_base = counts[H + 1 + h], e_h = counts[H + 2 + h] - e_base;
  if (p >= e_h || !flag[p]) return;
  const int32_t g = col_global[e_base + p];
  const int32_t lid = n_prev + rank[p];
  n_id[lid] = g;
  local_of[g] = lid;
}

__global__ void k_relabel(const int32_t* __restrict__ col_global, const int32_t* __restrict__ counts, int32_t h,
                          int32_t H, const int32_t* __restrict__ local_of, int32_t* __restrict__ first_pos,
                          int32_t* __restrict__ col) {
  const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int32_t e_base = counts[H + 1 + h], e_h = counts[H + 2 + h] - e_base;
  if (p >= e_h) return;
  const int32_t g = col_global[e_base + p];
  col[e_base + p] = local_of[g];
  first_pos[g] = INT_MAX;
}

// close the (empty) rows of the nodes discovered in the last hop and restore the relabel map
__global__ void k_finish(const int32_t* __restrict__ counts, int32_t H, const int32_t* __restrict__ n_id,
                         int32_t* __restrict__ rowptr, int32_t* __restrict__ local_of) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int32_t n = counts[H], n_exp = counts[H - 1], e = counts[2 * H + 1];
  if (i >= n) return;
  if (i >= n_exp) rowptr[i + 1] = e;
  local_of[n_id[i]] = -1;
}

// Feature-table addresses of a block: the table may be stored in a different row order than the node ids (hot rows
// first, see NeighborLoader(hot_feature_rows=...)); remap[global id] = table row.
__global__ void k_table_index(const int32_t* __restrict__ remap, const int32_t* __restrict__ col_global,
                              const int32_t* __restrict__ n_id, const int32_t* __restrict__ counts, int32_t H,
                              int32_t* __restrict__ col_table, int32_t* __restrict__ n_table) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int32_t n = counts[H], e = counts[2 * H + 1];
  if (i < e) col_table[i] = __ldg(remap + col_global[i]);
  if (i < n) n_table[i] = __ldg(remap + n_id[i]);
}

}  // namespace ngnn

using namespace ngnn;

extern "C" {

int32_t ngnn_block_table_index(const int32_t* remap, const int32_t* col_global, const int32_t* n_id, const int32_t* counts,
                               int32_t H, int64_t max_nodes, int64_t max_edges, int32_t* col_table, int32_t* n_table,
                               ngnn_stream_t stream) {
  NGNN_REQUIRE(remap && col_global && n_id && counts && col_table && n_table, NGNN_E_INVALID, "block_table_index: null pointer");
  NGNN_REQUIRE(H >= 1 && max_nodes >= 0 && max_edges >= 0, NGNN_E_INVALID, "block_table_index: bad sizes");
  const int64_t m = max_nodes > max_edges ? max_nodes : max_edges;
  if (m == 0) return NGNN_OK;
  k_table_index<<<(unsigned)ceil_div(m, 256), 256, 0, as_stream(stream)>>>(remap, col_global, n_id, counts, H, col_table, n_table);
  NGNN_LAUNCH_CHECK();
  return NGNN_OK;
}

int32_t ngnn_sample_capacity(int32_t bs, const int32_t* fanouts, int32_t H, int64_t N, int64_t* max_nodes,
                             int64_t* max_edges) {
  NGNN_REQUIRE(fanouts, NGNN_E_INVALID, "sample_capacity: fanouts is null");
  SampleCaps c;
  NGNN_REQUIRE(sample_caps(bs, fanouts, H, N, c), NGNN_E_INVALID,
               "sample_capacity: need 1 <= H <= 15, fanouts >= 1 and < 2^31 edges");
  if (max_nodes) *max_nodes = c.max_nodes;
  if (max_edges) *max_edges = c.max_edges;
  return NGNN_OK;
}

size_t ngnn_sample_workspace_bytes(int64_t N, int32_t bs, const int32_t* fanouts, int32_t H) {
  SampleCaps c;
  if (!fanouts || N < 0 || !sample_caps(bs, fanouts, H, N, c)) return 0;
  return sample_ws_layout(N, c, nullptr, nullptr);
}

int32_t ngnn_sample_workspace_init(void* ws, size_t ws_bytes, int64_t N, ngnn_stream_t stream) {
  NGNN_REQUIRE(ws && N >= 0, NGNN_E_INVALID, "sample_workspace_init: bad arguments");
  NGNN_REQUIRE(ws_bytes >= 2 * align_up((size_t)N * 4, 256) + 256, NGNN_E_WORKSPACE, "sample_workspace_init: workspace too small");
  if (N == 0) return NGNN_OK;
  char* b = reinterpret_cast<char*>(align_up(reinterpret_cast<size_t>(ws), 256));
  int32_t* local_of = (int32_t*)b;
  int32_t* first_pos = (int32_t*)(b + align_up((size_t)N * 4, 256));
  cudaStream_t st = as_stream(stream);
  k_fill_i32<<<(unsigned)ceil_div(N, 256), 256, 0, st>>>(local_of, N, -1);
  NGNN_LAUNCH_CHECK();
  k_fill_i32<<<(unsigned)ceil_div(N, 256), 256, 0, st>>>(first_pos, N, INT_MAX);
  NGNN_LAUNCH_CHECK();
  return NGNN_OK;
}

int32_t ngnn_sample_block(const int32_t* colptr, const int32_t* row, int64_t N, const int64_t* seeds, int32_t bs,
                          const int32_t* fanouts, int32_t H, int32_t replace, uint64_t seed, uint32_t epoch,
                          uint32_t batch_idx, int32_t* n_id, int32_t* rowptr, int32_t* col, int32_t* col_global,
                          int32_t* e_pos, int32_t* counts, void* ws, size_t ws_bytes, ngnn_stream_t stream) {
  NGNN_REQUIRE(colptr && row && fanouts && n_id && rowptr && col && col_global && counts && ws, NGNN_E_INVALID,
               "sample_block: null pointer");
  NGNN_REQUIRE(N > 0 && N < (1LL << 31) - 1, NGNN_E_INVALID, "sample_block: N out of int32 range");
  NGNN_REQUIRE(bs > 0 && seeds, NGNN_E_INVALID, "sample_block: need at least one seed");
  SampleCaps c;
  NGNN_REQUIRE(sample_caps(bs, fanouts, H, N, c), NGNN_E_INVALID,
               "sample_block: need 1 <= H <= 15, fanouts >= 1 and < 2^31 edges");
  for (int h = 0; h < H; ++h)
    NGNN_REQUIRE(replace || fanouts[h] <= kMaxFanout, NGNN_E_UNSUPPORTED,
                 "sample_block: fanout %d > %d without replacement", fanouts[h], kMaxFanout);
  SampleWs w;
  NGNN_REQUIRE(ws_bytes >= sample_ws_layout(N, c, ws, &w), NGNN_E_WORKSPACE, "sample_block: workspace too small");
  cudaStream_t st = as_stream(stream);
  const int T = 256;
  const uint32_t seed_lo = (uint32_t)seed, seed_hi = (uint32_t)(seed >> 32);

  k_seed_init<<<(unsigned)ceil_div(bs, T), T, 0, st>>>(seeds, bs, H, n_id, w.local_of, counts, rowptr);
  NGNN_LAUNCH_CHECK();
  for (int h = 0; h < H; ++h) {
    const int64_t fr = c.fr_max[h], em = c.e_max[h];
    k_count<<<(unsigned)ceil_div(fr + 1, T), T, 0, st>>>(colptr, n_id, counts, h, fanouts[h], replace, fr, w.cnt);
    NGNN_LAUNCH_CHECK();
    size_t cb = w.cub_bytes;
    NGNN_CUDA(cub::DeviceScan::ExclusiveSum(w.cub_tmp, cb, (const int32_t*)w.cnt, w.off, (int)(fr + 1), st));
    count_launches(2);   // cub scan: init + scan kernels
    if (fanouts[h] <= 32)
      k_draw_warp<<<(unsigned)ceil_div(fr * 32, 256), 256, 0, st>>>(colptr, row, n_id, counts, h, H, fanouts[h], replace, fr, w.off,
                                                                   seed_lo, seed_hi, epoch, batch_idx, rowptr, col_global, e_pos,
                                                                   w.local_of, w.first_pos);
    else
      k_draw<<<(unsigned)ceil_div(fr, 128), 128, 0, st>>>(colptr, row, n_id, counts, h, H, fanouts[h], replace, fr, w.off,
                                                         seed_lo, seed_hi, epoch, batch_idx, rowptr, col_global, e_pos,
                                                         w.local_of, w.first_pos);
    NGNN_LAUNCH_CHECK();
    k_flag<<<(unsigned)ceil_div(em + 1, T), T, 0, st>>>(col_global, counts, h, H, em, w.local_of, w.first_pos, w.flag);
    NGNN_LAUNCH_CHECK();
    cb = w.cub_bytes;
    NGNN_CUDA(cub::DeviceScan::ExclusiveSum(w.cub_tmp, cb, (const int32_t*)w.flag, w.rank, (int)(em + 1), st));
    count_launches(2);
    k_assign<<<(unsigned)ceil_div(em, T), T, 0, st>>>(col_global, counts, h, H, em, w.flag, w.rank, n_id, w.local_of);
    NGNN_LAUNCH_CHECK();
    k_relabel<<<(unsigned)ceil_div(em, T), T, 0, st>>>(col_global, counts, h, H, w.local_of, w.first_pos, col);
    NGNN_LAUNCH_CHECK();
  }
  k_finish<<<(unsigned)ceil_div(c.max_nodes, T), T, 0, st>>>(counts, H, n_id, rowptr, w.local_of);
  NGNN_LAUNCH_CHECK();
  return NGNN_OK;
}

}  // extern "C"
