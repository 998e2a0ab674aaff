// sampler.cu — K-SAMPLE / K-RELABEL / K-CSR: GPU fan-out neighbour sampler that emits CSR
// message-flow blocks (and, for the backward pass, their CSC transposes).  Replaces
// torch_geometric.loader.NeighborLoader's per-batch call into pyg-lib neighbor_sample (reference call sites
// src/pipeline.py:75-83 ctor, :152 iteration; semantics restated in SURVEY §8 row A1): hop-by-hop expansion in
// discovery order, take-all when deg <= fanout, distinct uniform positions otherwise (Robert Floyd's subset
// algorithm), first-seen relabelling with the seeds first, edges emitted grouped by destination => CSR by
// construction.
//
// B200-first design: the graph (CSC), the feature table and two N-sized relabel maps stay resident in HBM; a block is
// built by a FIXED sequence of 5 + 2H launches with worst-case grids whose real extents live in device memory
// (`counts`), so sampling needs no host round trip and is captured in the step's CUDA graph.  Per hop:
//   k_hop_draw    count -> exclusive scan -> draw, ONE kernel: each CTA counts the picks of its 64 frontier nodes, publishes
//                 the tile total, sums the totals of the tiles before it (they are published before anything is waited
//                 for, so there is no serial chain) and draws — one warp per node, lane = draw (Floyd's algorithm run
//                 cooperatively in registers, then every lane fetches its neighbour at once).
//   k_hop_assign  flag -> scan -> assign: first-seen relabelling made deterministic by an atomicMin over candidate
//                 positions (written by the draw), the same tile-scan, new local ids in position order.
// then k_relabel_all (local source ids for every edge + the transposes' histograms), k_tscan / k_tplace (CSC of the hop
// prefixes the backward needs: counting sort by local source id) and k_finish (closes the empty rows, restores the maps,
// puts every transposed row in canonical ascending order so that the backward's summation order is reproducible).
// Draws are counter-based (Philox4x32-10 keyed by (seed) with counter (node, hop|draw/4, batch, epoch)): a block is a pure
// function of (seed, epoch, batch_idx, seeds), independent of launch geometry and GPU count.  No library primitive.
#include "common.cuh"
#include <limits.h>

namespace ngnn {

constexpr int kMaxFanout = 64;   // Floyd's set lives in per-thread local storage
constexpr int kMaxTranspose = 4; // hop prefixes whose CSC a block can carry
constexpr int kDrawTile = 64;    // frontier nodes per CTA of k_hop_draw: 8 warps x 8 nodes.  The draw is a chain of dependent
                                 // global loads per node (neighbour id -> relabel slot), so what matters is how many nodes are
                                 // in flight: 256 nodes per CTA (32 per warp, one after the other) took 79 us on the 76.8 k
                                 // frontier of a products block, one node per warp (round 1) 14 us.
constexpr int kDrawThreads = 256;
int g_draw_group = 8;            // ngnn_set_tuning(16, g): fewest lanes per frontier node in k_hop_draw (8 / 16 / 32; 32 = one node per warp pass)
constexpr int kScanTile = 1024;  // positions per CTA of k_hop_assign / k_tscan (256 threads x 4)
constexpr int kSortSmem = 4096;  // longest transposed row sorted in shared memory by a CTA

struct SampleCaps {
  int64_t fr_max[16];   // worst-case frontier size per hop
  int64_t e_max[16];    // worst-case sampled edges per hop
  int64_t nodes_cum[17], edges_cum[17];   // worst-case cumulative nodes / edges after hop h-1
  int64_t max_nodes, max_edges;
};

static bool sample_caps(int32_t bs, const int32_t* fanouts, int32_t H, int64_t N, SampleCaps& c) {
  if (H < 1 || H > 15 || bs < 0) return false;
  int64_t fr = bs, nodes = bs, edges = 0;
  c.nodes_cum[0] = bs; c.edges_cum[0] = 0;
  for (int h = 0; h < H; ++h) {
    if (fanouts[h] < 1) return false;
    c.fr_max[h] = fr;
    c.e_max[h] = fr * (int64_t)fanouts[h];
    if (c.e_max[h] >= (1LL << 31)) return false;
    edges += c.e_max[h];
    fr = c.e_max[h] < N ? c.e_max[h] : N;     // new nodes discovered at hop h
    nodes += fr;
    c.nodes_cum[h + 1] = nodes < N + bs ? nodes : N + bs;
    c.edges_cum[h + 1] = edges;
  }
  if (edges >= (1LL << 31)) return false;
  c.max_nodes = c.nodes_cum[H];
  c.max_edges = edges;
  return true;
}

struct SampleWs {
  int32_t *local_of, *first_pos;
  int32_t* tcnt[kMaxTranspose];           // per transposed prefix: histogram over local source ids (zero at rest)
  unsigned long long* st_draw[16];        // per hop: published tile totals of k_hop_draw   (zero at rest)
  unsigned long long* st_assign[16];      //          of k_hop_assign
  unsigned long long* st_tscan[kMaxTranspose];
  uint32_t* tickets;                      // [2*16 + kMaxTranspose] dynamic tile ids (zero at rest)
  void* scratch_begin;                    // everything from here on is zero between calls
  size_t scratch_bytes;
  int32_t n_states;                       // total published-state words (for the reset in k_finish)
  unsigned long long* states_begin;
};

static inline int64_t draw_tiles(int64_t fr) { return ceil_div(fr > 0 ? fr : 1, kDrawTile); }
static inline int64_t scan_tiles(int64_t n) { return ceil_div(n > 0 ? n : 1, kScanTile); }

static size_t sample_ws_layout(int64_t N, int32_t H, const SampleCaps& c, void* ws, SampleWs* out) {
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += align_up(bytes > 0 ? bytes : 4, 256); return r; };
  const size_t o_local = take((size_t)N * 4), o_first = take((size_t)N * 4);
  // Everything behind the two maps is ZERO AT REST (k_finish restores it), so the partition below may change from call to
  // call (a short last batch, fewer hops) without any re-initialisation; nothing dirty may live in this region.
  const size_t o_scratch = o;
  size_t o_tcnt[kMaxTranspose];
  const int T = H < kMaxTranspose ? H : kMaxTranspose;
  for (int b = 1; b <= T; ++b) o_tcnt[b - 1] = take((size_t)(c.nodes_cum[b] + 1) * 4);
  const size_t o_states = o;
  size_t o_sd[16], o_sa[16], o_st[kMaxTranspose];
  int64_t n_states = 0;
  for (int h = 0; h < H; ++h) {
    o_sd[h] = take((size_t)draw_tiles(c.fr_max[h]) * 8);
    o_sa[h] = take((size_t)scan_tiles(c.e_max[h]) * 8);
  }
  for (int b = 1; b <= T; ++b) o_st[b - 1] = take((size_t)scan_tiles(c.max_nodes + 1) * 8);   // k_tscan's grid covers the widest prefix
  const size_t o_tick = take((2 * 16 + kMaxTranspose) * 4);
  n_states = (int64_t)((o - o_states) / 8);
  if (out) {
    char* b0 = reinterpret_cast<char*>(align_up(reinterpret_cast<size_t>(ws), 256));
    out->local_of = (int32_t*)(b0 + o_local); out->first_pos = (int32_t*)(b0 + o_first);
    for (int b = 0; b < kMaxTranspose; ++b) { out->tcnt[b] = b < T ? (int32_t*)(b0 + o_tcnt[b]) : nullptr; out->st_tscan[b] = b < T ? (unsigned long long*)(b0 + o_st[b]) : nullptr; }
    for (int h = 0; h < 16; ++h) { out->st_draw[h] = h < H ? (unsigned long long*)(b0 + o_sd[h]) : nullptr; out->st_assign[h] = h < H ? (unsigned long long*)(b0 + o_sa[h]) : nullptr; }
    out->tickets = (uint32_t*)(b0 + o_tick);
    out->scratch_begin = b0 + o_scratch; out->scratch_bytes = o - o_scratch;
    out->states_begin = (unsigned long long*)(b0 + o_states); out->n_states = (int32_t)n_states;
  }
  return o + 256;
}

__global__ void k_fill_i32(int32_t* p, int64_t n, int32_t v) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ---------------------------------------------------------------------------------------------------
// Tile scan shared by k_hop_draw / k_hop_assign / k_tscan: every CTA takes a ticket (its tile id, in scheduling
// order), scans its own values, PUBLISHES the tile total (valid bit | total, one 64-bit word), then adds up the totals
// of all earlier tiles.  A tile only waits for tiles with smaller tickets, which are running or done and publish before
// they wait for anything: no serial chain and no co-residency requirement.  The words are zero at rest (k_finish).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

struct TileScan {
  int excl;        // exclusive prefix of this thread's value over the whole sequence
  int total;       // inclusive prefix at the end of this tile (= grand total for the last tile)
};

// blockDim.x == 256, `tile` = this CTA's ticket.  `mine` = this thread's value (threads are in sequence order within the tile).
__device__ __forceinline__ TileScan tile_scan_256(int mine, int tile, unsigned long long* state) {
  __shared__ int s_wsum[8], s_psum[8], s_prefix;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int incl = warp_incl_scan(mine, lane);
  if (lane == 31) s_wsum[warp] = incl;
  __syncthreads();
  int woff = 0, total = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) { const int s = s_wsum[w]; if (w < warp) woff += s; total += s; }
  if (tid == 0) {
    __threadfence();
    *reinterpret_cast<volatile unsigned long long*>(state + tile) = (1ull << 32) | (unsigned long long)(uint32_t)total;
  }
  int part = 0;
  for (int j = tid; j < tile; j += 256) {
    unsigned long long w;
    do { w = *reinterpret_cast<volatile unsigned long long*>(state + j); } while ((w >> 32) == 0ull);
    part += (int)(uint32_t)w;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if (lane == 0) s_psum[warp] = part;
  __syncthreads();
  if (tid == 0) { int s = 0; for (int w = 0; w < 8; ++w) s += s_psum[w]; s_prefix = s; }
  __syncthreads();
  TileScan r;
  r.excl = s_prefix + woff + incl - mine;
  r.total = s_prefix + total;
  return r;
}

__device__ __forceinline__ int take_ticket(uint32_t* ticket) {
  __shared__ int s_ticket;
  if (threadIdx.x == 0) s_ticket = (int)atomicAdd(ticket, 1u);
  __syncthreads();
  return s_ticket;
}

__global__ void k_seed_init(const int64_t* __restrict__ seeds, int32_t bs, int32_t H, int64_t N, int32_t* __restrict__ n_id,
                            int32_t* __restrict__ local_of, int32_t* __restrict__ counts, int32_t* __restrict__ rowptr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < bs) {
    const int64_t g64 = seeds[i];
    const int32_t g = (g64 >= 0 && g64 < N) ? (int32_t)g64 : 0;   // ids are validated on the host; never write out of bounds
    n_id[i] = g;
    atomicMax(local_of + g, i);     // duplicated seeds: the last occurrence wins, deterministically (like the oracle)
  }
  if (i == 0) { counts[0] = bs; counts[H + 1] = 0; rowptr[0] = 0; }
}

struct DrawArgs {
  const int32_t* colptr; const int32_t* row; const int32_t* n_id; int32_t* counts;
  int32_t h, H, fanout, replace;
  uint32_t seed_lo, seed_hi, epoch, batch_idx;
  const StepCtl* ctl;
  int32_t *rowptr, *col_global, *e_pos, *edge_dst;
  const int32_t* local_of; int32_t* first_pos;
  unsigned long long* state; uint32_t* ticket;
};

// Fused count -> scan -> draw of one hop.  WARP: one warp per frontier node, lane = draw (fan-out <= 32): the positions are
// decided in registers (Floyd's subset algorithm run cooperatively: one shuffle + one vote per draw, no memory), then
// every lane reads ITS neighbour id and relabel slot at once — one round of dependent DRAM latencies per node instead of
// one per draw.  !WARP: one thread per node, any fan-out.  Same positions, emission order and Philox stream as the C oracle.
// WARP = lanes per frontier node (8, 16 or 32: the smallest that holds the fan-out, so a fan-out of 5 runs four nodes per warp
// pass — the draw is bound by instruction issue, Philox + Floyd per node, not by bytes), or 0 = one thread per node.
template <int WARP>
__global__ void __launch_bounds__(kDrawThreads) k_hop_draw(const DrawArgs a) {
  __shared__ int s_v[kDrawTile], s_beg[kDrawTile], s_d[kDrawTile], s_off[kDrawTile];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int32_t h = a.h, H = a.H, fanout = a.fanout, replace = a.replace;
  const int32_t lo = h == 0 ? 0 : a.counts[h - 1], hi = a.counts[h];
  const int32_t e_base = a.counts[H + 1 + h];
  const int32_t fr = hi - lo;
  uint32_t epoch = a.epoch, batch_idx = a.batch_idx;
  if (a.ctl != nullptr) { epoch = a.ctl->epoch; batch_idx = a.ctl->batch_idx; }

  const int tile = take_ticket(a.ticket);
  const int64_t i = (int64_t)tile * kDrawTile + tid;
  int32_t v = -1, beg = 0, d = 0, k = 0;
  if (tid < kDrawTile && i < fr) {
    v = a.n_id[lo + i];
    beg = __ldg(a.colptr + v);
    d = __ldg(a.colptr + v + 1) - beg;
    k = replace ? (d > 0 ? fanout : 0) : min(d, fanout);
  }
  const TileScan sc = tile_scan_256(k, tile, a.state);
  const int32_t off = sc.excl;
  if (tile == (int)gridDim.x - 1 && tid == kDrawThreads - 1) a.counts[H + 2 + h] = e_base + sc.total;   // edges after this hop
  if (tid < kDrawTile && i < fr) a.rowptr[lo + i + 1] = e_base + off + k;

  if (WARP == 0) {
    // one thread per node
    if (k == 0) return;
    auto emit = [&](int32_t j, int32_t pos) {
      const int32_t g = __ldg(a.row + beg + pos);
      const int32_t p = e_base + off + j;
      a.col_global[p] = g;
      if (a.edge_dst) a.edge_dst[p] = lo + (int32_t)i;
      if (a.e_pos) a.e_pos[p] = beg + pos;
      if (a.local_of[g] < 0) atomicMin(a.first_pos + g, off + j);
    };
    if (!replace && d <= fanout) {
      for (int32_t j = 0; j < d; ++j) emit(j, j);          // take all, stored order
      return;
    }
    Philox4 r{0, 0, 0, 0};
    if (replace) {
      for (int32_t j = 0; j < k; ++j) {
        if ((j & 3) == 0) r = philox4x32_10((uint32_t)v, ((uint32_t)h << 16) | (uint32_t)(j >> 2), batch_idx, epoch, a.seed_lo, a.seed_hi);
        const uint32_t w = (j & 3) == 0 ? r.x : (j & 3) == 1 ? r.y : (j & 3) == 2 ? r.z : r.w;
        emit(j, (int32_t)mulhi32(w, (uint32_t)d));
      }
      return;
    }
    // Robert Floyd: for jj = d-k .. d-1: t = U[0,jj]; take t unless already taken, else take jj
    int32_t S[kMaxFanout];
    for (int32_t j = 0; j < k; ++j) {
      if ((j & 3) == 0) r = philox4x32_10((uint32_t)v, ((uint32_t)h << 16) | (uint32_t)(j >> 2), batch_idx, epoch, a.seed_lo, a.seed_hi);
      const uint32_t w = (j & 3) == 0 ? r.x : (j & 3) == 1 ? r.y : (j & 3) == 2 ? r.z : r.w;
      const int32_t jj = d - k + j;
      int32_t t = (int32_t)mulhi32(w, (uint32_t)(jj + 1));
      for (int32_t q = 0; q < j; ++q) if (S[q] == t) { t = jj; break; }
      S[j] = t;
      emit(j, t);
    }
    return;
  }

  // G lanes per node, 8 nodes per warp in NP passes of 32 / G nodes.  The positions of all the warp's nodes are decided first
  // (registers only), then their neighbour-id loads go out together, then the relabel-map loads: two rounds of DRAM latency
  // per warp instead of two per node.
  if (tid < kDrawTile) { s_v[tid] = v; s_beg[tid] = beg; s_d[tid] = d; s_off[tid] = off; }
  __syncthreads();
  constexpr int G = WARP > 0 ? WARP : 32;
  constexpr int NPW = kDrawTile / (kDrawThreads / 32);       // nodes per warp
  constexpr int NP = NPW * G / 32;                           // passes
  const int gl = lane & (G - 1), sg = lane / G;
  const unsigned gmask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (sg * G));
  int32_t pos[NP], g[NP];
#pragma unroll
  for (int u = 0; u < NP; ++u) {
    const int jn = warp + (u * (32 / G) + sg) * (kDrawThreads / 32);
    const int32_t nv = s_v[jn], nd = s_d[jn];
    const int32_t nk = nv < 0 ? 0 : (replace ? (nd > 0 ? fanout : 0) : min(nd, fanout));
    int32_t p = gl;                                            // take all, stored order
    if (nk > 0 && (replace || nd > fanout)) {                  // uniform within the node's lane group
      const Philox4 r = philox4x32_10((uint32_t)nv, ((uint32_t)h << 16) | (uint32_t)(gl >> 2), batch_idx, epoch, a.seed_lo, a.seed_hi);
      const uint32_t w = (gl & 3) == 0 ? r.x : (gl & 3) == 1 ? r.y : (gl & 3) == 2 ? r.z : r.w;
      if (replace) {
        p = (int32_t)mulhi32(w, (uint32_t)nd);
      } else {
        // Robert Floyd: for jj = d-k .. d-1: t = U[0,jj]; take t unless already taken, else take jj
        int32_t mine = -1;
        for (int32_t j = 0; j < nk; ++j) {
          const uint32_t wj = __shfl_sync(gmask, w, j, G);
          const int32_t jj = nd - nk + j;
          int32_t t = (int32_t)mulhi32(wj, (uint32_t)(jj + 1));
          if (__any_sync(gmask, gl < j && mine == t)) t = jj;
          if (gl == j) mine = t;
        }
        p = mine;
      }
    }
    pos[u] = gl < nk ? p : -1;
  }
#pragma unroll
  for (int u = 0; u < NP; ++u) {
    const int jn = warp + (u * (32 / G) + sg) * (kDrawThreads / 32);
    g[u] = pos[u] >= 0 ? __ldg(a.row + s_beg[jn] + pos[u]) : -1;
  }
  int32_t lof[NP];
#pragma unroll
  for (int u = 0; u < NP; ++u) {
    const int jn = warp + (u * (32 / G) + sg) * (kDrawThreads / 32);
    lof[u] = 0;
    if (g[u] >= 0) {
      const int32_t p = e_base + s_off[jn] + gl;
      a.col_global[p] = g[u];
      if (a.edge_dst) a.edge_dst[p] = lo + tile * kDrawTile + jn;
      if (a.e_pos) a.e_pos[p] = s_beg[jn] + pos[u];
      lof[u] = a.local_of[g[u]];
    }
  }
#pragma unroll
  for (int u = 0; u < NP; ++u) {
    const int jn = warp + (u * (32 / G) + sg) * (kDrawThreads / 32);
    if (g[u] >= 0 && lof[u] < 0) atomicMin(a.first_pos + g[u], s_off[jn] + gl);
  }
}

// Fused flag -> scan -> assign of one hop: position p of this hop's edges introduces a new node iff its neighbour is not
// labelled yet and p is the smallest position that names it; new nodes get consecutive local ids in position order.
__global__ void __launch_bounds__(256) k_hop_assign(const int32_t* __restrict__ col_global, int32_t* __restrict__ counts, int32_t h,
                                                    int32_t H, int32_t* __restrict__ local_of, const int32_t* __restrict__ first_pos,
                                                    int32_t* __restrict__ n_id, unsigned long long* state, uint32_t* ticket) {
  const int tid = threadIdx.x;
  const int tile = take_ticket(ticket);
  const int32_t n_prev = counts[h];
  const int32_t e_base = counts[H + 1 + h], e_h = counts[H + 2 + h] - e_base;
  const int64_t p0 = (int64_t)tile * kScanTile + tid * 4;
  // the two map reads of a position are random DRAM accesses: all eight of a thread are issued before any is looked at
  int32_t g[4], lof[4], fp[4];
  int f = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) g[q] = (p0 + q < e_h) ? __ldg(col_global + e_base + p0 + q) : -1;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    lof[q] = g[q] >= 0 ? local_of[g[q]] : 0;
    fp[q] = g[q] >= 0 ? __ldg(first_pos + g[q]) : -1;
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (g[q] >= 0 && lof[q] < 0 && fp[q] == (int32_t)(p0 + q)) ++f;
    else g[q] = -1;
  }
  const TileScan sc = tile_scan_256(f, tile, state);
  int32_t lid = n_prev + sc.excl;
  if (tile == (int)gridDim.x - 1 && tid == 255) counts[h + 1] = n_prev + sc.total;     // nodes after this hop
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (g[q] >= 0) {
      n_id[lid] = g[q];
      local_of[g[q]] = lid;     // only the one position that owns this neighbour writes; every other reader's flag is 0 either way
      ++lid;
    }
  }
}

struct TransposeArgs {
  int32_t num;                          // prefixes 1..num
  int32_t* colptr_t[kMaxTranspose];
  int32_t* row_t[kMaxTranspose];
  int32_t* tcnt[kMaxTranspose];
  unsigned long long* state[kMaxTranspose];
};

// local source id of every sampled edge; restore first_pos; histogram of the transposed prefixes
__global__ void k_relabel_all(const int32_t* __restrict__ col_global, const int32_t* __restrict__ counts, int32_t H,
                              const int32_t* __restrict__ local_of, int32_t* __restrict__ first_pos, int32_t* __restrict__ col,
                              const TransposeArgs t) {
  const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= counts[2 * H + 1]) return;
  const int32_t g = col_global[p];
  const int32_t c = local_of[g];
  col[p] = c;
  first_pos[g] = INT_MAX;
  for (int b = 1; b <= t.num; ++b)
    if (p < counts[H + 1 + b]) atomicAdd(t.tcnt[b - 1] + c, 1);
}

// colptr_t of prefix b (blockIdx.y = b - 1) = exclusive scan of its histogram over the counts[b] local nodes it spans
__global__ void __launch_bounds__(256) k_tscan(const int32_t* __restrict__ counts, int32_t H, const TransposeArgs t, uint32_t* tickets) {
  const int b = blockIdx.y + 1;
  const int tid = threadIdx.x;
  const int tile = take_ticket(tickets + blockIdx.y);
  const int32_t n_cols = counts[b];
  const int32_t* cnt = t.tcnt[b - 1];
  int32_t* colptr_t = t.colptr_t[b - 1];
  unsigned long long* state = t.state[b - 1];
  const int64_t j0 = (int64_t)tile * kScanTile + tid * 4;
  int v[4], f = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) { v[q] = (j0 + q < n_cols) ? cnt[j0 + q] : 0; f += v[q]; }
  const TileScan sc = tile_scan_256(f, tile, state);
  int run = sc.excl;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (j0 + q <= n_cols) colptr_t[j0 + q] = run;     // entry n_cols closes the last row (= number of edges of the prefix)
    run += v[q];
  }
}

// scatter the destinations into their source's row; counting the histogram back down leaves it zero for the next block
__global__ void k_tplace(const int32_t* __restrict__ col, const int32_t* __restrict__ edge_dst, const int32_t* __restrict__ counts,
                         int32_t H, const TransposeArgs t) {
  const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= counts[H + 1 + t.num]) return;
  const int32_t c = col[p], d = edge_dst[p];
  for (int b = 1; b <= t.num; ++b) {
    if (p < counts[H + 1 + b]) {
      const int32_t slot = t.colptr_t[b - 1][c] + atomicSub(t.tcnt[b - 1] + c, 1) - 1;
      t.row_t[b - 1][slot] = d;
    }
  }
}

__device__ __forceinline__ void heap_sift(int32_t* a, int32_t start, int32_t end) {
  int32_t root = start;
  while (2 * root + 1 <= end) {
    int32_t child = 2 * root + 1;
    if (child + 1 <= end && a[child] < a[child + 1]) ++child;
    if (a[root] < a[child]) { const int32_t tmp = a[root]; a[root] = a[child]; a[child] = tmp; root = child; }
    else return;
  }
}

// Closes the (empty) rows of the nodes discovered in the last hop, restores the relabel map, zeroes the published scan
// states / tickets, and sorts every transposed row ascending (the scatter above fills a row in arrival order; ascending
// destination order = ascending edge position = the stable order, so the backward's summation order is reproducible).
// Rows up to 32 entries: insertion sort by their thread; longer rows: bitonic sort by the whole CTA in shared memory
// (rows beyond kSortSmem entries fall back to an in-place heapsort by one thread — star-shaped blocks only).
__global__ void __launch_bounds__(256) k_finish(const int32_t* __restrict__ counts, int32_t H, const int32_t* __restrict__ n_id,
                                                int32_t* __restrict__ rowptr, int32_t* __restrict__ local_of,
                                                const TransposeArgs t, unsigned long long* states, int32_t n_states,
                                                uint32_t* tickets, int32_t n_tickets) {
  __shared__ int32_t s_buf[kSortSmem];
  __shared__ int s_long[256], s_nlong;
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int32_t n = counts[H], n_exp = counts[H - 1], e = counts[2 * H + 1];
  if (i < n) {
    if (i >= n_exp) rowptr[i + 1] = e;
    local_of[n_id[i]] = -1;
  }
  if (i < n_states) states[i] = 0ull;
  if (i < n_tickets) tickets[i] = 0u;
  for (int b = 1; b <= t.num; ++b) {
    if (threadIdx.x == 0) s_nlong = 0;
    __syncthreads();
    const int32_t n_cols = counts[b];
    int32_t* row_t = t.row_t[b - 1];
    if (i < n_cols) {
      const int32_t beg = t.colptr_t[b - 1][i], end = t.colptr_t[b - 1][i + 1];
      const int32_t len = end - beg;
      if (len > 32) {
        s_long[atomicAdd(&s_nlong, 1)] = threadIdx.x;          // list order does not matter: each row is sorted on its own
      } else if (len > 1) {
        int32_t* a = row_t + beg;
        for (int32_t x = 1; x < len; ++x) {
          const int32_t key = a[x];
          int32_t y = x - 1;
          while (y >= 0 && a[y] > key) { a[y + 1] = a[y]; --y; }
          a[y + 1] = key;
        }
      }
    }
    __syncthreads();
    const int nlong = s_nlong;
    for (int k = 0; k < nlong; ++k) {
      const int64_t r = blockIdx.x * (int64_t)blockDim.x + s_long[k];
      const int32_t beg = t.colptr_t[b - 1][r], len = t.colptr_t[b - 1][r + 1] - beg;
      int32_t* a = row_t + beg;
      if (len > kSortSmem) {
        if (threadIdx.x == 0) {
          for (int32_t s0 = (len - 2) / 2; s0 >= 0; --s0) heap_sift(a, s0, len - 1);
          for (int32_t end2 = len - 1; end2 > 0; --end2) {
            const int32_t tmp = a[end2]; a[end2] = a[0]; a[0] = tmp;
            heap_sift(a, 0, end2 - 1);
          }
        }
        __syncthreads();
        continue;
      }
      int32_t m = 64;
      while (m < len) m <<= 1;
      for (int32_t x = threadIdx.x; x < m; x += 256) s_buf[x] = x < len ? a[x] : INT_MAX;
      __syncthreads();
      for (int32_t kk = 2; kk <= m; kk <<= 1) {
        for (int32_t j = kk >> 1; j > 0; j >>= 1) {
          for (int32_t x = threadIdx.x; x < m; x += 256) {
            const int32_t y = x ^ j;
            if (y > x) {
              const int32_t ax = s_buf[x], ay = s_buf[y];
              const bool up = (x & kk) == 0;
              if ((ax > ay) == up) { s_buf[x] = ay; s_buf[y] = ax; }
            }
          }
          __syncthreads();
        }
      }
      for (int32_t x = threadIdx.x; x < len; x += 256) a[x] = s_buf[x];
      __syncthreads();
    }
  }
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

__global__ void k_step_ctl_set(StepCtl* ctl, uint32_t epoch, uint32_t batch_idx, uint32_t off_lo, uint32_t off_hi, float loss_scale) {
  ctl->epoch = epoch; ctl->batch_idx = batch_idx; ctl->drop_off_lo = off_lo; ctl->drop_off_hi = off_hi;
  ctl->loss_scale_bits = __float_as_uint(loss_scale);
  ctl->ticket = 0u;
}

}  // namespace ngnn

using namespace ngnn;

extern "C" {

int32_t ngnn_step_ctl_set(ngnn_step_ctl_t* ctl, uint32_t epoch, uint32_t batch_idx, uint64_t drop_offset, float loss_scale,
                          ngnn_stream_t stream) {
  NGNN_REQUIRE(ctl != nullptr, NGNN_E_INVALID, "step_ctl_set: null pointer");
  k_step_ctl_set<<<1, 1, 0, as_stream(stream)>>>(reinterpret_cast<StepCtl*>(ctl), epoch, batch_idx, (uint32_t)drop_offset,
                                                (uint32_t)(drop_offset >> 32), loss_scale);
  NGNN_LAUNCH_CHECK();
  return NGNN_OK;
}

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
  return sample_ws_layout(N, H, c, nullptr, nullptr);
}

int32_t ngnn_sample_workspace_init(void* ws, size_t ws_bytes, int64_t N, ngnn_stream_t stream) {
  NGNN_REQUIRE(ws && N >= 0, NGNN_E_INVALID, "sample_workspace_init: bad arguments");
  NGNN_REQUIRE(ws_bytes >= 2 * align_up((size_t)N * 4, 256) + 256, NGNN_E_WORKSPACE, "sample_workspace_init: workspace too small");
  char* b = reinterpret_cast<char*>(align_up(reinterpret_cast<size_t>(ws), 256));
  const size_t maps = 2 * align_up((size_t)(N > 0 ? N : 1) * 4, 256);
  cudaStream_t st = as_stream(stream);
  // everything behind the two maps (histograms, published scan states, tickets) is zero at rest
  char* end = reinterpret_cast<char*>(ws) + ws_bytes;
  if (end > b + maps) NGNN_CUDA(cudaMemsetAsync(b + maps, 0, (size_t)(end - (b + maps)), st));
  if (N == 0) return NGNN_OK;
  int32_t* local_of = (int32_t*)b;
  int32_t* first_pos = (int32_t*)(b + align_up((size_t)N * 4, 256));
  k_fill_i32<<<(unsigned)ceil_div(N, 256), 256, 0, st>>>(local_of, N, -1);
  NGNN_LAUNCH_CHECK();
  k_fill_i32<<<(unsigned)ceil_div(N, 256), 256, 0, st>>>(first_pos, N, INT_MAX);
  NGNN_LAUNCH_CHECK();
  return NGNN_OK;
}

int32_t ngnn_sample_block_ex(const int32_t* colptr, const int32_t* row, int64_t N, const int64_t* seeds, int32_t bs,
                             const int32_t* fanouts, int32_t H, int32_t replace, uint64_t seed, uint32_t epoch,
                             uint32_t batch_idx, const ngnn_step_ctl_t* ctl, int32_t* n_id, int32_t* rowptr, int32_t* col,
                             int32_t* col_global, int32_t* e_pos, int32_t* edge_dst, int32_t* counts, int32_t num_transposes,
                             int32_t* const* colptr_t, int32_t* const* row_t, void* ws, size_t ws_bytes, ngnn_stream_t stream) {
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
  NGNN_REQUIRE(num_transposes >= 0 && num_transposes <= kMaxTranspose && num_transposes <= H, NGNN_E_INVALID,
               "sample_block: 0 <= num_transposes <= min(H, %d)", kMaxTranspose);
  NGNN_REQUIRE(num_transposes == 0 || (colptr_t && row_t && edge_dst), NGNN_E_INVALID,
               "sample_block: transposes requested without buffers (colptr_t, row_t, edge_dst)");
  SampleWs w;
  NGNN_REQUIRE(ws_bytes >= sample_ws_layout(N, H, c, ws, &w), NGNN_E_WORKSPACE, "sample_block: workspace too small");
  cudaStream_t st = as_stream(stream);
  const int T = 256;

  k_seed_init<<<(unsigned)ceil_div(bs, T), T, 0, st>>>(seeds, bs, H, N, n_id, w.local_of, counts, rowptr);
  NGNN_LAUNCH_CHECK();
  for (int h = 0; h < H; ++h) {
    DrawArgs a{};
    a.colptr = colptr; a.row = row; a.n_id = n_id; a.counts = counts; a.h = h; a.H = H; a.fanout = fanouts[h]; a.replace = replace;
    a.seed_lo = (uint32_t)seed; a.seed_hi = (uint32_t)(seed >> 32); a.epoch = epoch; a.batch_idx = batch_idx;
    a.ctl = reinterpret_cast<const StepCtl*>(ctl);
    a.rowptr = rowptr; a.col_global = col_global; a.e_pos = e_pos; a.edge_dst = edge_dst;
    a.local_of = w.local_of; a.first_pos = w.first_pos; a.state = w.st_draw[h]; a.ticket = w.tickets + 2 * h;
    const unsigned gd = (unsigned)draw_tiles(c.fr_max[h]);
    if (fanouts[h] <= 8 && g_draw_group <= 8) k_hop_draw<8><<<gd, kDrawThreads, 0, st>>>(a);
    else if (fanouts[h] <= 16 && g_draw_group <= 16) k_hop_draw<16><<<gd, kDrawThreads, 0, st>>>(a);
    else if (fanouts[h] <= 32) k_hop_draw<32><<<gd, kDrawThreads, 0, st>>>(a);
    else k_hop_draw<0><<<gd, kDrawThreads, 0, st>>>(a);
    NGNN_LAUNCH_CHECK();
    k_hop_assign<<<(unsigned)scan_tiles(c.e_max[h]), 256, 0, st>>>(col_global, counts, h, H, w.local_of, w.first_pos, n_id,
                                                                 w.st_assign[h], w.tickets + 2 * h + 1);
    NGNN_LAUNCH_CHECK();
  }
  TransposeArgs t{};
  t.num = num_transposes;
  for (int b = 0; b < num_transposes; ++b) {
    NGNN_REQUIRE(colptr_t[b] && row_t[b], NGNN_E_INVALID, "sample_block: null transpose buffer %d", b);
    t.colptr_t[b] = colptr_t[b]; t.row_t[b] = row_t[b]; t.tcnt[b] = w.tcnt[b]; t.state[b] = w.st_tscan[b];
  }
  k_relabel_all<<<(unsigned)ceil_div(c.max_edges > 0 ? c.max_edges : 1, T), T, 0, st>>>(col_global, counts, H, w.local_of,
                                                                                      w.first_pos, col, t);
  NGNN_LAUNCH_CHECK();
  if (num_transposes > 0) {
    dim3 gs((unsigned)scan_tiles(c.nodes_cum[num_transposes] + 1), (unsigned)num_transposes);
    k_tscan<<<gs, 256, 0, st>>>(counts, H, t, w.tickets + 32);
    NGNN_LAUNCH_CHECK();
    k_tplace<<<(unsigned)ceil_div(c.edges_cum[num_transposes] > 0 ? c.edges_cum[num_transposes] : 1, T), T, 0, st>>>(
        col, edge_dst, counts, H, t);
    NGNN_LAUNCH_CHECK();
  }
  int64_t fin = c.max_nodes;
  if (w.n_states > fin) fin = w.n_states;
  if (fin < 64) fin = 64;
  k_finish<<<(unsigned)ceil_div(fin, T), T, 0, st>>>(counts, H, n_id, rowptr, w.local_of, t, w.states_begin, w.n_states,
                                                   w.tickets, 2 * 16 + kMaxTranspose);
  NGNN_LAUNCH_CHECK();
  return NGNN_OK;
}

int32_t ngnn_sample_block(const int32_t* colptr, const int32_t* row, int64_t N, const int64_t* seeds, int32_t bs,
                          const int32_t* fanouts, int32_t H, int32_t replace, uint64_t seed, uint32_t epoch,
                          uint32_t batch_idx, int32_t* n_id, int32_t* rowptr, int32_t* col, int32_t* col_global,
                          int32_t* e_pos, int32_t* counts, void* ws, size_t ws_bytes, ngnn_stream_t stream) {
  return ngnn_sample_block_ex(colptr, row, N, seeds, bs, fanouts, H, replace, seed, epoch, batch_idx, nullptr, n_id, rowptr, col,
                              col_global, e_pos, nullptr, counts, 0, nullptr, nullptr, ws, ws_bytes, stream);
}

}  // extern "C"
