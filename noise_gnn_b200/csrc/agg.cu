// agg.cu — K-AGG / K-AGG-T: atomic-free CSR segment reductions over fp32 feature rows.
//
// Forward  (SAGEConv mean aggregation, PyG ops K1-K3 in SURVEY §2.3; reference call site
//           src/models/layers/sage.py:34):    mean[i] = 1/max(deg_i,1) * sum_p x[col[p]]
// Backward (transpose segment sum, K9-K10):   dx[j]   = gate_j * (sum_q dmean[row_t[q]] + dx_root[j])
//
// One group of G lanes (G = 8/16/32, picked from the row width) owns one output row, so the
// reduction needs no atomics and its summation order is fixed (deterministic).  Feature rows are
// read as 128-bit vectors, coalesced across the group; up to 4 neighbour rows x VPL vectors are
// in flight per lane.  The segment's indices are loaded once, coalesced, by the group and
// broadcast with warp shuffles.  HBM-bound: algorithmic bytes = 4F*(distinct sources) + 4e +
// 4(n_dst+1) + 4F*n_dst.
#include "common.cuh"

namespace ngnn {

struct AggParams {
  const int32_t* ptr;      // rowptr (fwd) / colptr_t (bwd), [n_rows+1]
  const int32_t* idx;      // col (fwd) / row_t (bwd), [e]
  const float* x;          // source rows
  int64_t ld_x;
  int64_t n_rows;          // number of output rows (the launch bound / capacity when n_rows_dev != nullptr)
  const int32_t* n_rows_dev;  // optional device-side row count (a sampled block's extent), clamped to n_rows
  int64_t F;
  float* out;
  int64_t ld_out;
  int32_t mean;            // 1: scale by 1/max(deg,1)
  const float* add;        // optional rows added for i < n_add (bwd: dx_root)
  int64_t ld_add;
  int64_t n_add;
  const int32_t* n_add_dev; // optional device-side n_add (clamped to n_add)
  const float* act_ref;    // optional gate: out *= (act_ref > 0 ? act_scale : 0)
  int64_t ld_act;
  float act_scale;
  const float* bias;       // optional per-column bias added after the reduction (GCNConv: out = A_sum z + b)
  int32_t keep_l2;         // table-mode gathers (root_idx != NULL) carry an L2 priority:
  int64_t hot_rows;        //   < 0: every row evict_last; >= 0: rows < hot_rows evict_last, the others evict_first
  int32_t l1_alloc;        // gathers allocate in L1 (wide rows re-gathered by neighbouring output rows)
  int32_t long_row;        // rows with more neighbours than this are reduced by the whole CTA (hub rows)
  unsigned long long* clock; // optional [2]: min %globaltimer at CTA start / max at CTA end (in-kernel duration, ngnn_probe_*)
  const int32_t* root_idx; // optional fused root gather (fwd): root[i] = x[root_idx[i]]
  float* root;
  int64_t ld_root;
};

__device__ __forceinline__ void f4_add(float4& a, const float4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }

__device__ __forceinline__ int64_t agg_n_add(const AggParams& p) {
  if (p.n_add_dev == nullptr) return p.n_add;
  const int64_t v = (int64_t)__ldg(p.n_add_dev);
  return v < p.n_add ? (v < 0 ? 0 : v) : p.n_add;
}
__device__ __forceinline__ int64_t agg_rows(const AggParams& p) {
  if (p.n_rows_dev == nullptr) return p.n_rows;
  const int64_t v = (int64_t)__ldg(p.n_rows_dev);
  return v < p.n_rows ? (v < 0 ? 0 : v) : p.n_rows;
}

// Gather-accumulate of the neighbours [beg, end) of one output row into acc (columns c0 + gl + v*G), U rows in flight.
template <int G, int VPL, int U>
__device__ __forceinline__ void seg_accumulate(const AggParams& p, int beg, int end, int c0, int F4, int gl, unsigned gmask,
                                               float4 (&acc)[VPL]) {
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int base = beg; base < end; base += G) {
    const int cnt = min(G, end - base);
    const int my = (gl < cnt) ? __ldg(p.idx + base + gl) : 0;
    for (int j = 0; j < cnt; j += U) {
      // up to U neighbour rows in flight; out-of-range slots are masked
      float4 v4[U][VPL];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int su = __shfl_sync(gmask, my, min(j + u, cnt - 1), G);
        const float4* src = reinterpret_cast<const float4*>(p.x + (int64_t)su * p.ld_x);
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          const int c = c0 + gl + v * G;
          v4[u][v] = (c < F4 && j + u < cnt) ? (p.l1_alloc ? __ldg(src + c) : ldg_nc_f4(src + c)) : zero4;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int v = 0; v < VPL; ++v) f4_add(acc[v], v4[u][v]);
    }
  }
}

// scale, bias, add row, gate, store of one finished column chunk of `row`.  PRE: the gate / add rows were requested
// before the gather (narrow rows: the registers are there); otherwise they are loaded here.
template <int G, int VPL, bool PRE>
__device__ __forceinline__ void seg_finish(const AggParams& p, int64_t row, int c0, int F4, int gl, float scale, bool has_add,
                                           const float4 (&acc)[VPL], const float4 (&gate)[PRE ? VPL : 1],
                                           const float4 (&addv)[PRE ? VPL : 1]) {
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int c = c0 + gl + v * G;
    if (c >= F4) continue;
    float4 r = acc[v];
    r.x *= scale; r.y *= scale; r.z *= scale; r.w *= scale;
    if (p.bias != nullptr) f4_add(r, __ldg(reinterpret_cast<const float4*>(p.bias) + c));
    if (has_add) f4_add(r, PRE ? addv[PRE ? v : 0] : ldg_nc_f4(reinterpret_cast<const float4*>(p.add + row * p.ld_add) + c));
    if (p.act_ref != nullptr) {
      const float4 h = PRE ? gate[PRE ? v : 0] : ldg_nc_f4(reinterpret_cast<const float4*>(p.act_ref + row * p.ld_act) + c);
      r.x = h.x > 0.f ? r.x * p.act_scale : 0.f;
      r.y = h.y > 0.f ? r.y * p.act_scale : 0.f;
      r.z = h.z > 0.f ? r.z * p.act_scale : 0.f;
      r.w = h.w > 0.f ? r.w * p.act_scale : 0.f;
    }
    reinterpret_cast<float4*>(p.out + row * p.ld_out)[c] = r;
  }
}

// Generic one-row-per-group kernel (all widths, optional add / gate / root gather).  U = neighbour rows in flight.
// G == 32 only: rows longer than kLongRow (hubs of a power-law graph: un-sampled convolutions, the transposed blocks of
// the backward) are not walked by their one warp — the CTA's warps each reduce a contiguous slice of the row and the
// slices are combined through shared memory in warp order (still atomic-free and run-to-run deterministic).
// Measured on the Computers-shaped sweep (transposed rows: mean 10, p99 58, max 139 neighbours; profiles/prof_wide.py): lowering
// the threshold from 1024 to 64 takes the F = 512 transpose-sum from 108 us to 80 us and F = 1024 from 219 us to 166 us — the
// few rows an order of magnitude above the mean were each walked by one warp while the rest of the machine idled.
constexpr int kLongRow = 64;
constexpr int kLongWarps = 8;
template <int G, int VPL>
constexpr int seg_max_threads() { return (G == 16 && VPL == 4) ? 128 : 512; }   // the half-warp-per-row variant runs small CTAs
template <int G, int VPL, int U>
__global__ void __launch_bounds__(seg_max_threads<G, VPL>()) k_seg_reduce_v4(AggParams p) {
  pdl_trigger();
  pdl_wait();
  constexpr int GROUPS_PER_WARP = 32 / G;
  // hub rows are handed to the whole CTA, a full warp per slice: CV vectors per lane cover the same G*VPL columns of a pass
  constexpr bool COOP = (G >= 16) && (G * VPL >= 32);
  constexpr int CV = COOP ? G * VPL / 32 : 1;
  // VPL == 4 (F 260..512): the 2 x VPL float4 of prefetch registers drop the kernel from 3 to 2 CTAs per SM (measured
  // 2.5x slower on the C5 sweep); VPL == 8 is at 2 CTAs per SM either way and measured faster with the prefetch.
  constexpr bool PRE = (VPL != 4);
  constexpr int NPRE = PRE ? VPL : 1;
  __shared__ float4 s_part[COOP ? kLongWarps : 1][COOP ? CV * 32 : 1];
  __shared__ int s_long[64];
  __shared__ int s_nlong;
  const int lane = threadIdx.x & 31;
  const int gl = lane & (G - 1);                 // lane within group
  const int gw = lane / G;                       // group within warp
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (gw * G));
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t row = warp * GROUPS_PER_WARP + gw;
  const int64_t n_rows = agg_rows(p);
  const bool valid = row < n_rows;
  if (!COOP && !valid) return;
  if (COOP) {
    if (threadIdx.x == 0) s_nlong = 0;
    __syncthreads();
  }

  const int F4 = (int)((p.F + 3) >> 2);             // a partial last vector reads / writes the rows' padding (ld >= 4*F4)
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  int beg = 0, end = 0;
  if (valid) {
    // The extents and the indices are two small arrays read strictly in row order: each group asks L2 for the sectors
    // the rows kPrefetchRows further on will need.
    constexpr int64_t kPrefetchRows = 16384;
    const int e_total = __ldg(p.ptr + n_rows);
    if (gl == 0 && row + kPrefetchRows <= n_rows) prefetch_l2(p.ptr + row + kPrefetchRows);
    beg = __ldg(p.ptr + row); end = __ldg(p.ptr + row + 1);
    if (gl == 0 && (int64_t)beg + kPrefetchRows < e_total) prefetch_l2(p.idx + beg + kPrefetchRows);
  }
  const bool is_long = COOP && valid && (end - beg) > p.long_row;
  if (valid && !is_long) {
    const float scale = p.mean ? 1.0f / (float)max(end - beg, 1) : 1.0f;
    const bool has_add = p.add != nullptr && row < agg_n_add(p);
    for (int c0 = blockIdx.y * (G * VPL); c0 < F4; c0 += gridDim.y * (G * VPL)) {
      float4 acc[VPL];
#pragma unroll
      for (int v = 0; v < VPL; ++v) acc[v] = zero4;
      // the gate / add rows of THIS output row do not depend on the extents -> indices -> rows chain: request them
      // first so their DRAM latency runs under it (backward: 2 of the 3 streams of the kernel)
      float4 gate[NPRE], addv[NPRE];
      if (PRE) {
#pragma unroll
        for (int v = 0; v < NPRE; ++v) {
          const int c = c0 + gl + v * G;
          gate[v] = (p.act_ref != nullptr && c < F4) ? ldg_nc_f4(reinterpret_cast<const float4*>(p.act_ref + row * p.ld_act) + c) : zero4;
          addv[v] = (has_add && c < F4) ? ldg_nc_f4(reinterpret_cast<const float4*>(p.add + row * p.ld_add) + c) : zero4;
        }
      }
      seg_accumulate<G, VPL, U>(p, beg, end, c0, F4, gl, gmask, acc);
      seg_finish<G, VPL, PRE>(p, row, c0, F4, gl, scale, has_add, acc, gate, addv);
    }
  }
  if (COOP) {
    const int wib = threadIdx.x >> 5, nw = min((int)(blockDim.x >> 5), kLongWarps);
    if (is_long && gl == 0) s_long[atomicAdd(&s_nlong, 1)] = wib * GROUPS_PER_WARP + gw;   // list order is irrelevant to each row's result
    __syncthreads();
    const int nlong = s_nlong;                                            // CTA-uniform
    for (int k = 0; k < nlong; ++k) {
      const int64_t lrow = (int64_t)blockIdx.x * (blockDim.x >> 5) * GROUPS_PER_WARP + s_long[k];
      const int lbeg = __ldg(p.ptr + lrow), lend = __ldg(p.ptr + lrow + 1);
      const int slice = ((lend - lbeg + nw - 1) / nw + 31) & ~31;          // whole 32-index windows per warp
      const int sb = min(lend, lbeg + wib * slice), se = wib < nw ? min(lend, sb + slice) : sb;
      const float scale = p.mean ? 1.0f / (float)max(lend - lbeg, 1) : 1.0f;
      const bool has_add = p.add != nullptr && lrow < agg_n_add(p);
      for (int c0 = blockIdx.y * (32 * CV); c0 < F4; c0 += gridDim.y * (32 * CV)) {
        float4 acc[CV];
#pragma unroll
        for (int v = 0; v < CV; ++v) acc[v] = zero4;
        seg_accumulate<32, CV, U>(p, sb, se, c0, F4, lane, 0xffffffffu, acc);
        if (wib < nw) {
#pragma unroll
          for (int v = 0; v < CV; ++v) s_part[wib][v * 32 + lane] = acc[v];
        }
        __syncthreads();
        if (wib == 0) {
          const float4 none[1] = {zero4};
#pragma unroll
          for (int v = 0; v < CV; ++v)
            for (int w = 1; w < nw; ++w) f4_add(acc[v], s_part[w][v * 32 + lane]);     // fixed warp order
          seg_finish<32, CV, false>(p, lrow, c0, F4, lane, scale, has_add, acc, none, none);
        }
        __syncthreads();
      }
    }
    if (!valid) return;
  }
  if (p.root_idx != nullptr && blockIdx.y == 0) {
    const float4* src = reinterpret_cast<const float4*>(p.x + (int64_t)__ldg(p.root_idx + row) * p.ld_x);
    float4* dst = reinterpret_cast<float4*>(p.root + row * p.ld_root);
    for (int c = gl; c < F4; c += G) dst[c] = ldg_nc_f4(src + c);
  }
}

// ---------------------------------------------------------------------------------------------------
// K-AGG-T with the gate rows staged through shared memory (training backward of a sampled block, 64 < F <= 256):
//     dX[j] = gate_j * ( sum_{q in CSC row j} dmean[row_t[q]]  +  dX_root[j] )        gate_j = (h_j > 0) * act_scale
// Nearly every transposed row of a sampled block has ONE entry, and dmean / dX_root (a few thousand rows, just written by
// K-DGRAD) sit in L2, so the kernel is a stream: read the gate row (the saved layer output h), write the dX row.  In the
// generic kernel those reads are bounded by registers — 56 rows in flight per SM, ncu: 2.4 TB/s, 57 % of the stall samples
// waiting for the gate loads.  Here a CTA owns 8 x RPG consecutive rows and the copy engine brings their gate rows into shared
// memory (`cp.async.bulk`, one per row, completing on an mbarrier) while the half-warps walk extents -> indices -> rows,
// RPG rows each: the gate bytes in flight cost no registers.  RPG = 2 measured best (products step 0.410 ms with the
// generic kernel, 0.394 with RPG = 4, 0.391 with RPG = 2; the arxiv-shaped block, whose transposed rows are longer, 0.233 /
// 0.243 / 0.229: with 4 rows per half-warp too few gather chains are in flight).  Same summation order as the generic kernel (stored order, then
// the root row), long rows handed to the whole CTA in the same way => the two kernels are bitwise interchangeable.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t agg_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void agg_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; spin < (1u << 26); ++spin) {              // bounded: a protocol bug traps instead of hanging the GPU
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) return;
  }
  __trap();
}

constexpr int kStageThreads = 128;  // 8 half-warps x RPG rows
template <int VPL, int RPG>         // vectors per lane of a 16-lane group: F4 <= 16 * VPL; rows per group
__global__ void __launch_bounds__(kStageThreads, RPG == 4 ? 6 : (RPG == 2 ? 7 : 8)) k_aggT_staged(AggParams p) {
  constexpr int G = 16, kStageTile = RPG * (kStageThreads / G), NW = kStageThreads / 32;
  constexpr int CV = G * VPL / 32;
  static_assert(CV >= 1, "a full warp covers the group's columns");
  extern __shared__ __align__(128) unsigned char stage_raw[];       // gate tile: kStageTile rows x F4 vectors
  __shared__ float4 s_part[NW - 1][CV * 32];                       // warp 0 keeps its own partial sums
  __shared__ int s_long[kStageTile];
  __shared__ int s_nlong;
  __shared__ __align__(8) uint64_t s_bar;
  pdl_trigger();
  pdl_wait();
  float4* s_gate = reinterpret_cast<float4*>(stage_raw);
  const int F4 = (int)((p.F + 3) >> 2);
  const int64_t n_rows = agg_rows(p);
  const int64_t row0 = (int64_t)blockIdx.x * kStageTile;
  if (row0 >= n_rows) return;                                        // CTA-uniform
  const int tile_rows = (int)((n_rows - row0) < kStageTile ? (n_rows - row0) : kStageTile);
  const int tid = threadIdx.x, lane = tid & 31, gl = lane & (G - 1), grp = tid / G, wib = tid >> 5;
  const unsigned gmask = 0xffffu << ((lane >> 4) * 16);
  const uint32_t bar = agg_smem_u32(&s_bar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    s_nlong = 0;
  }
  __syncthreads();
  if (wib == 0) {
    if (lane == 0)
      asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}"
                   ::"r"(bar), "r"((uint32_t)(tile_rows * F4 * 16)) : "memory");
    __syncwarp();
    if (lane < tile_rows)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(agg_smem_u32(s_gate + (size_t)lane * F4)), "l"(p.act_ref + (row0 + lane) * p.ld_act),
                     "r"((uint32_t)(F4 * 16)), "r"(bar) : "memory");
  }

  // rows of this group: tile-local r = i * 8 + grp (consecutive groups take consecutive rows).  The three dependent loads of
  // the four rows are issued as three batches.
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const int64_t n_add = p.add != nullptr ? agg_n_add(p) : 0;
  int beg[RPG], end[RPG], my[RPG];
#pragma unroll
  for (int i = 0; i < RPG; ++i) {
    const int r = i * (kStageThreads / G) + grp;
    beg[i] = end[i] = 0;
    if (r < tile_rows) { beg[i] = __ldg(p.ptr + row0 + r); end[i] = __ldg(p.ptr + row0 + r + 1); }
  }
#pragma unroll
  for (int i = 0; i < RPG; ++i) my[i] = (beg[i] + gl < end[i]) ? __ldg(p.idx + beg[i] + gl) : 0;

  bool waited = false;
#pragma unroll
  for (int i = 0; i < RPG; ++i) {
    const int r = i * (kStageThreads / G) + grp;
    if (r >= tile_rows) continue;
    const int64_t row = row0 + r;
    const int deg = end[i] - beg[i];
    if (deg > p.long_row) {                                          // hub row: the whole CTA reduces it below
      if (gl == 0) s_long[atomicAdd(&s_nlong, 1)] = r;
      continue;
    }
    float4 acc[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) acc[v] = zero4;
    // first window of <= 16 indices is already in registers
    {
      const int cnt = min(G, deg);
      for (int j = 0; j < cnt; j += 2) {
        float4 v4[2][VPL];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int su = __shfl_sync(gmask, my[i], min(j + u, cnt - 1), G);
          const float4* src = reinterpret_cast<const float4*>(p.x + (int64_t)su * p.ld_x);
#pragma unroll
          for (int v = 0; v < VPL; ++v) {
            const int c = gl + v * G;
            v4[u][v] = (c < F4 && j + u < cnt) ? ldg_nc_f4(src + c) : zero4;
          }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
          for (int v = 0; v < VPL; ++v) f4_add(acc[v], v4[u][v]);
      }
    }
    if (deg > G) seg_accumulate<G, VPL, 2>(p, beg[i] + G, end[i], 0, F4, gl, gmask, acc);
    if (!waited) { agg_mbar_wait(bar, 0); waited = true; }
    const bool has_add = row < n_add;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int c = gl + v * G;
      if (c >= F4) continue;
      float4 o = acc[v];
      if (has_add) f4_add(o, ldg_nc_f4(reinterpret_cast<const float4*>(p.add + row * p.ld_add) + c));
      const float4 h = s_gate[(size_t)r * F4 + c];
      o.x = h.x > 0.f ? o.x * p.act_scale : 0.f;
      o.y = h.y > 0.f ? o.y * p.act_scale : 0.f;
      o.z = h.z > 0.f ? o.z * p.act_scale : 0.f;
      o.w = h.w > 0.f ? o.w * p.act_scale : 0.f;
      reinterpret_cast<float4*>(p.out + row * p.ld_out)[c] = o;
    }
  }
  __syncthreads();
  const int nlong = s_nlong;                                         // CTA-uniform
  if (nlong == 0) return;
  if (!waited) agg_mbar_wait(bar, 0);
  for (int k = 0; k < nlong; ++k) {
    const int r = s_long[k];
    const int64_t lrow = row0 + r;
    const int lbeg = __ldg(p.ptr + lrow), lend = __ldg(p.ptr + lrow + 1);
    const int slice = ((lend - lbeg + NW - 1) / NW + 31) & ~31;      // whole 32-index windows per warp
    const int sb = min(lend, lbeg + wib * slice), se = min(lend, sb + slice);
    const bool has_add = lrow < n_add;
    for (int c0 = 0; c0 < F4; c0 += 32 * CV) {
      float4 acc[CV];
#pragma unroll
      for (int v = 0; v < CV; ++v) acc[v] = zero4;
      seg_accumulate<32, CV, 2>(p, sb, se, c0, F4, lane, 0xffffffffu, acc);
      if (wib > 0) {
#pragma unroll
        for (int v = 0; v < CV; ++v) s_part[wib - 1][v * 32 + lane] = acc[v];
      }
      __syncthreads();
      if (wib == 0) {
#pragma unroll
        for (int v = 0; v < CV; ++v) {
          for (int w = 1; w < NW; ++w) f4_add(acc[v], s_part[w - 1][v * 32 + lane]);  // fixed warp order
          const int c = c0 + lane + v * 32;
          if (c >= F4) continue;
          float4 o = acc[v];
          if (has_add) f4_add(o, ldg_nc_f4(reinterpret_cast<const float4*>(p.add + lrow * p.ld_add) + c));
          const float4 h = s_gate[(size_t)r * F4 + c];
          o.x = h.x > 0.f ? o.x * p.act_scale : 0.f;
          o.y = h.y > 0.f ? o.y * p.act_scale : 0.f;
          o.z = h.z > 0.f ? o.z * p.act_scale : 0.f;
          o.w = h.w > 0.f ? o.w * p.act_scale : 0.f;
          reinterpret_cast<float4*>(p.out + lrow * p.ld_out)[c] = o;
        }
      }
      __syncthreads();
    }
  }
}

// Scalar variant for rows that are not 16-byte addressable (F % 4 != 0, e.g. 1433 or 767).
template <int VPL>
__global__ void __launch_bounds__(256) k_seg_reduce_scalar(AggParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (row >= agg_rows(p)) return;
  const int F = (int)p.F;
  const int beg = __ldg(p.ptr + row), end = __ldg(p.ptr + row + 1);
  const float scale = p.mean ? 1.0f / (float)max(end - beg, 1) : 1.0f;

  for (int c0 = 0; c0 < F; c0 += 32 * VPL) {
    float acc[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) acc[v] = 0.f;
    for (int base = beg; base < end; base += 32) {
      const int cnt = min(32, end - base);
      const int my = (lane < cnt) ? __ldg(p.idx + base + lane) : 0;
      for (int j = 0; j < cnt; j += 2) {
        const int s0 = __shfl_sync(0xffffffffu, my, j);
        const int s1 = __shfl_sync(0xffffffffu, my, min(j + 1, cnt - 1));
        const float* r0 = p.x + (int64_t)s0 * p.ld_x;
        const float* r1 = p.x + (int64_t)s1 * p.ld_x;
        float a[VPL], b[VPL];
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          const int c = c0 + lane + v * 32;
          a[v] = (c < F) ? __ldg(r0 + c) : 0.f;
          b[v] = (c < F && j + 1 < cnt) ? __ldg(r1 + c) : 0.f;
        }
#pragma unroll
        for (int v = 0; v < VPL; ++v) { acc[v] += a[v]; acc[v] += b[v]; }
      }
    }
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int c = c0 + lane + v * 32;
      if (c >= F) continue;
      float r = acc[v] * scale;
      if (p.bias != nullptr) r += __ldg(p.bias + c);
      if (p.add != nullptr && row < agg_n_add(p)) r += __ldg(p.add + row * p.ld_add + c);
      if (p.act_ref != nullptr) r = __ldg(p.act_ref + row * p.ld_act + c) > 0.f ? r * p.act_scale : 0.f;
      p.out[row * p.ld_out + c] = r;
    }
  }
  if (p.root_idx != nullptr) {
    const float* src = p.x + (int64_t)__ldg(p.root_idx + row) * p.ld_x;
    float* dst = p.root + row * p.ld_root;
    for (int c = lane; c < F; c += 32) dst[c] = __ldg(src + c);
  }
}

// ---------------------------------------------------------------------------------------------------
// Software-pipelined persistent forward kernel (the layer-1 hot case: short rows, F <= 256).
// A sampled row costs three dependent memory round trips (extents -> indices -> feature rows) and only the
// third moves real bytes, so a warp that walks one row at a time keeps HBM idle two thirds of the time (ncu:
// 32 % DRAM, 43 % warps active, everything latency-bound).  Here every warp owns rows r, r+W, r+2W, ... and
// runs a 3-stage pipeline in registers: while the feature rows of row i are in flight it has already issued
// the index load of row i+1 and the extent / root-id loads of row i+2, so each iteration exposes ONE latency.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

template <int VPL, int U, bool ROOT>
__global__ void __launch_bounds__(256) k_agg_fwd_pipe(AggParams p) {
  pdl_trigger();
  pdl_wait();
  if (p.clock != nullptr && threadIdx.x == 0) atomicMin(p.clock, global_ns());
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t n = agg_rows(p);
  const int F4 = (int)((p.F + 3) >> 2);
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

  // ROOT = gathering from the resident feature table (layer 1): keep its rows in L2 ahead of the streaming activations
  const bool keep = ROOT && p.keep_l2;
  const uint64_t pol_hot = l2_policy_evict_last(), pol_cold = p.hot_rows >= 0 ? l2_policy_evict_first() : pol_hot;
  const int64_t hot = p.hot_rows >= 0 ? p.hot_rows : INT64_MAX;
  int64_t row0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int64_t row1 = row0 + nwarps;
  int beg0 = 0, end0 = 0, rid0 = 0, beg1 = 0, end1 = 0, rid1 = 0, my0 = 0;
  if (row0 < n) { beg0 = __ldg(p.ptr + row0); end0 = __ldg(p.ptr + row0 + 1); if (ROOT) rid0 = __ldg(p.root_idx + row0); }
  if (row1 < n) { beg1 = __ldg(p.ptr + row1); end1 = __ldg(p.ptr + row1 + 1); if (ROOT) rid1 = __ldg(p.root_idx + row1); }
  if (row0 < n && lane < min(32, end0 - beg0)) my0 = __ldg(p.idx + beg0 + lane);

  while (row0 < n) {
    // stage A: extents + root id of the row after next
    const int64_t row2 = row1 + nwarps;
    int beg2 = 0, end2 = 0, rid2 = 0;
    if (row2 < n) { beg2 = __ldg(p.ptr + row2); end2 = __ldg(p.ptr + row2 + 1); if (ROOT) rid2 = __ldg(p.root_idx + row2); }
    // stage B: first index chunk of the next row
    int my1 = 0;
    if (row1 < n && lane < min(32, end1 - beg1)) my1 = __ldg(p.idx + beg1 + lane);

    // stage C: gather the feature rows of row0
    float4 acc[VPL], rootv[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int c = lane + v * 32;
      acc[v] = zero4;
      rootv[v] = (ROOT && c < F4) ? (keep ? ldg_nc_f4_hint(reinterpret_cast<const float4*>(p.x + (int64_t)rid0 * p.ld_x) + c, rid0 < hot ? pol_hot : pol_cold)
                                          : ldg_nc_f4(reinterpret_cast<const float4*>(p.x + (int64_t)rid0 * p.ld_x) + c)) : zero4;
    }
    int my = my0;
    for (int base = beg0; base < end0; base += 32) {
      const int cnt = min(32, end0 - base);
      if (base != beg0) my = (lane < cnt) ? __ldg(p.idx + base + lane) : 0;     // long rows only (deg > 32)
      for (int j = 0; j < cnt; j += U) {
        float4 v4[U][VPL];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int su = __shfl_sync(0xffffffffu, my, min(j + u, cnt - 1));
          const float4* src = reinterpret_cast<const float4*>(p.x + (int64_t)su * p.ld_x);
#pragma unroll
          for (int v = 0; v < VPL; ++v) {
            const int c = lane + v * 32;
            v4[u][v] = (c < F4 && j + u < cnt) ? (keep ? ldg_nc_f4_hint(src + c, su < hot ? pol_hot : pol_cold) : ldg_nc_f4(src + c)) : zero4;
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int v = 0; v < VPL; ++v) f4_add(acc[v], v4[u][v]);
      }
    }
    const float scale = p.mean ? 1.0f / (float)max(end0 - beg0, 1) : 1.0f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int c = lane + v * 32;
      if (c >= F4) continue;
      float4 r = acc[v];
      r.x *= scale; r.y *= scale; r.z *= scale; r.w *= scale;
      reinterpret_cast<float4*>(p.out + row0 * p.ld_out)[c] = r;
      if (ROOT) reinterpret_cast<float4*>(p.root + row0 * p.ld_root)[c] = rootv[v];
    }
    // rotate the pipeline registers
    row0 = row1; beg0 = beg1; end0 = end1; rid0 = rid1; my0 = my1;
    row1 = row2; beg1 = beg2; end1 = end2; rid1 = rid2;
  }
  if (p.clock != nullptr) {
    __syncthreads();
    if (threadIdx.x == 0) atomicMax(p.clock + 1, global_ns());
  }
}

static int g_tune_small = 1;     // ngnn_set_tuning(15, 0|1): few-row form of the pipelined K-AGG (64-thread CTAs, 8 rows in flight)
template <int VPL, int U, bool ROOT>
static void launch_pipe(const AggParams& p, cudaStream_t st) {
  static int blocks_per_sm = 0;
  if (blocks_per_sm == 0) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_agg_fwd_pipe<VPL, U, ROOT>, 256, 0) != cudaSuccess ||
        blocks_per_sm < 1) {
      cudaGetLastError();
      blocks_per_sm = 2;
    }
  }
  int64_t grid = (int64_t)kNumSMs * blocks_per_sm;                 // persistent: one resident wave
  const int64_t need = ceil_div(p.n_rows, 8);                        // 8 warps per CTA
  if (need < kNumSMs && g_tune_small) {
    // fewer rows than one 256-thread CTA per SM (the top layers of a step): 64-thread CTAs spread the rows over every SM
    launch_chain(k_agg_fwd_pipe<VPL, U, ROOT>, dim3((unsigned)ceil_div(p.n_rows, 2)), dim3(64), 0, st, p);
    return;
  }
  if (grid > need) grid = need;
  launch_chain(k_agg_fwd_pipe<VPL, U, ROOT>, dim3((unsigned)grid), dim3(256), 0, st, p);
}

static int g_tune_unroll = 0;    // ngnn_set_tuning(0, u): 0 = default, else force U in {2,4,8} for F <= 128
static int g_tune_threads = 256; // ngnn_set_tuning(1, t): CTA size 128 / 256 / 512
static int g_tune_pipe = 1;      // ngnn_set_tuning(3, 0/1): software-pipelined persistent forward kernel
static int g_tune_group = 32;    // ngnn_set_tuning(2, g): lanes per row for 64 < F <= 128 (32 / 16 / 8)
static int g_tune_keep = 1;      // ngnn_set_tuning(7, 0|1): L2 evict_last priority on the layer-1 table gathers
static int g_tune_long = kLongRow;  // ngnn_set_tuning(11, n): hub-row threshold of the generic kernel
static int g_tune_xwide = 0;     // ngnn_set_tuning(12, v): F > 256: 0 = column chunks on grid.y (default); 1 = the row's warp walks them; 2 = 128-thread CTAs
static int g_tune_stage = 1;     // ngnn_set_tuning(14, v): staged K-AGG-T: 0 off, 1 = 2 rows per half-warp (default), 2 = also for few rows (tests), 3 = 4 rows, 4 = 1 row
static int g_tune_l1 = 0;        // ngnn_set_tuning(13, v): gathers of the generic kernel allocate in L1: 0 = wide rows only, 1 = always, 2 = never
static int g_tune_wide = 0;      // ngnn_set_tuning(9, v): generic kernel for 128 < F <= 256: 0 = half-warp/row x4 vectors, 128-thread
                                 //   CTAs (default); 1 = warp/row x2 vectors unroll 2; 2 = warp/row unroll 4

// split_cols: the column chunks (G*VPL vectors each) of a wide row go to different CTAs (grid.y) instead of being walked
// one after the other by the row's warp — more, smaller units of work and more rows in flight.
template <int G, int VPL, int U>
static void launch_v4(const AggParams& p, cudaStream_t st, int threads = 0, bool split_cols = false) {
  int T = threads > 0 && g_tune_threads == 256 ? threads : g_tune_threads;
  if (T > seg_max_threads<G, VPL>()) T = seg_max_threads<G, VPL>();
  const int64_t rows_per_block = (T / 32) * (32 / G);
  const unsigned chunks = split_cols ? (unsigned)ceil_div((p.F + 3) / 4, (int64_t)G * VPL) : 1u;
  launch_chain(k_seg_reduce_v4<G, VPL, U>, dim3((unsigned)ceil_div(p.n_rows, rows_per_block), chunks), dim3(T), 0, st, p);
}

static int32_t run_agg(const AggParams& p_in, cudaStream_t st) {
  AggParams p = p_in;
  p.long_row = g_tune_long;
  p.l1_alloc = g_tune_l1 == 1 || (g_tune_l1 == 0 && p.F > 256);
  if (p.n_rows == 0 || p.F == 0) return NGNN_OK;
  // 128-bit path: rows addressed as whole float4 vectors.  A width that is not a multiple of 4 (1433, 767) qualifies when
  // every row has the padding behind it (ld >= 4*ceil(F/4), as the loader's table and the step arena guarantee): the last
  // vector then reads / writes pad columns, which no consumer looks at.
  const int64_t F4 = (p.F + 3) / 4, Fv = 4 * F4;
  auto rows_ok = [&](const void* base, int64_t ld) { return (ld % 4 == 0) && ld >= Fv && is_aligned(base, 16); };
  bool vec = rows_ok(p.x, p.ld_x) && rows_ok(p.out, p.ld_out);
  if (p.add) vec = vec && rows_ok(p.add, p.ld_add);
  if (p.act_ref) vec = vec && rows_ok(p.act_ref, p.ld_act);
  if (p.root_idx) vec = vec && rows_ok(p.root, p.ld_root);
  if (p.bias) vec = vec && is_aligned(p.bias, 16) && p.F % 4 == 0;
  if (vec) {
    const bool fwd_plain = p.add == nullptr && p.act_ref == nullptr && p.bias == nullptr;   // forward aggregation (mean [+ root gather])
    // The persistent pipelined kernel deals rows to warps statically: right for the near-uniform row lengths of a sampled
    // block's forward (<= fan-out), wrong for a transposed block / an un-sampled graph (sum form), whose hub rows would each be
    // walked by one warp — Computers-shaped transpose, F = 256, fan-out 25: 90 us there, 61 us in the generic kernel.
    if (fwd_plain && p.mean && g_tune_pipe && F4 > 16 && F4 <= 64) {
      if (F4 <= 32) {
        const int u = g_tune_unroll;
        if (p.root_idx) {
          // measured on B200, products layer 1 (profiles/r01_agg_sweep.txt): U=6 is the sweet spot (61 % of HBM peak)
          if (u == 8) launch_pipe<1, 8, true>(p, st); else if (u == 5) launch_pipe<1, 5, true>(p, st);
          else if (u == 4) launch_pipe<1, 4, true>(p, st); else if (u == 3) launch_pipe<1, 3, true>(p, st);
          else launch_pipe<1, 6, true>(p, st);
        } else {
          if (u == 8) launch_pipe<1, 8, false>(p, st); else if (u == 4) launch_pipe<1, 4, false>(p, st);
          else launch_pipe<1, 6, false>(p, st);
        }
      } else {
        // few rows (the last layer: one row per seed): occupancy is no concern, 8 neighbour rows in flight per lane
        if (p.root_idx) launch_pipe<2, 4, true>(p, st);
        else if (p.n_rows <= 2048 && g_tune_small) launch_pipe<2, 8, false>(p, st);
        else launch_pipe<2, 4, false>(p, st);
      }
      NGNN_LAUNCH_CHECK();
      return NGNN_OK;
    }
    // training backward of a sampled block (gate present, sum form): gate rows staged through shared memory
    // (from ~4 tiles per SM up: the 7.6 k-row transpose of the layer above measured 9.6 us in the generic kernel, 14 us here)
    if (g_tune_stage && !p.mean && p.act_ref != nullptr && p.bias == nullptr && p.root_idx == nullptr && F4 > 16 && F4 <= 64 &&
        (p.n_rows >= (int64_t)kNumSMs * 4 * 32 || g_tune_stage == 2)) {
      const int rpg = g_tune_stage == 3 ? 4 : (g_tune_stage == 4 ? 1 : 2), tile = rpg * (kStageThreads / 16);
      const unsigned grid = (unsigned)ceil_div(p.n_rows, (int64_t)tile);
      const size_t smem = (size_t)tile * F4 * 16;
      if (F4 <= 32) {
        if (rpg == 2) launch_chain(k_aggT_staged<2, 2>, dim3(grid), dim3(kStageThreads), smem, st, p);
        else if (rpg == 1) launch_chain(k_aggT_staged<2, 1>, dim3(grid), dim3(kStageThreads), smem, st, p);
        else launch_chain(k_aggT_staged<2, 4>, dim3(grid), dim3(kStageThreads), smem, st, p);
      } else {
        if (rpg == 2) launch_chain(k_aggT_staged<4, 2>, dim3(grid), dim3(kStageThreads), smem, st, p);
        else if (rpg == 1) launch_chain(k_aggT_staged<4, 1>, dim3(grid), dim3(kStageThreads), smem, st, p);
        else launch_chain(k_aggT_staged<4, 4>, dim3(grid), dim3(kStageThreads), smem, st, p);
      }
      NGNN_LAUNCH_CHECK();
      return NGNN_OK;
    }
    if (F4 <= 8) launch_v4<8, 1, 8>(p, st);
    else if (F4 <= 16) launch_v4<16, 1, 8>(p, st);
    else if (F4 <= 32) {
      const int u = g_tune_unroll;
      if (g_tune_group == 16) { if (u == 2) launch_v4<16, 2, 2>(p, st); else if (u == 8) launch_v4<16, 2, 8>(p, st); else launch_v4<16, 2, 4>(p, st); }
      else if (g_tune_group == 8) { if (u == 2) launch_v4<8, 4, 2>(p, st); else if (u == 8) launch_v4<8, 4, 8>(p, st); else launch_v4<8, 4, 4>(p, st); }
      else { if (u == 2) launch_v4<32, 1, 2>(p, st); else if (u == 8) launch_v4<32, 1, 8>(p, st); else launch_v4<32, 1, 4>(p, st); }
    }
    else if (F4 <= 64) {
      // measured on the layer-2 backward of a products block (77 k rows x 256, mostly one transposed neighbour per row;
      // profiles/prof_aggT.py): warp per row 61 us, half-warp per row with 128-thread CTAs 47 us (twice the rows in flight)
      // ... and (profiles/prof_mid.py) on transposed rows of mean length 10-25 without a gate: warp per row, 4 rows in flight 61 us,
      // against 82 us for the half-warp form
      const bool short_rows = p.act_ref != nullptr || p.add != nullptr;      // training backward of a sampled block
      if (g_tune_wide == 2 || (g_tune_wide == 0 && !short_rows)) launch_v4<32, 2, 4>(p, st);
      else if (g_tune_wide == 1) launch_v4<32, 2, 2>(p, st);
      else launch_v4<16, 4, 2>(p, st, 128);
    }
    else {
      // F > 256 (profiles/prof_wide.py, Computers-shaped sweep): each 128-vector column chunk of a row goes to its own CTA
      // row (grid.y), 4 vectors per lane x 2 neighbour rows in flight, gathers allocating in L1.  Against the row's warp
      // walking its chunks with 8 vectors per lane (128 registers, 2 CTAs per SM): F = 1433, fan-out 25 forward 258 -> 129 us,
      // transpose-sum 516 -> 190 us; F = 512 55 -> 49 / 106 -> 82 us.  Narrower chunks (1-2 vectors per lane) re-read the
      // indices too often and were slower than no split; 1 or 4 rows in flight, 512-thread CTAs and the scalar kernel too.
      if (g_tune_xwide == 1) launch_v4<32, 8, 1>(p, st);                        // round-1 form, kept for A/B
      else if (g_tune_xwide == 2) launch_v4<32, 4, 2>(p, st, 128, true);
      else launch_v4<32, 4, 2>(p, st, 0, true);
    }
  } else {
    const unsigned grid = (unsigned)ceil_div(p.n_rows * 32, 256);
    if (p.F <= 128) k_seg_reduce_scalar<4><<<grid, 256, 0, st>>>(p);
    else k_seg_reduce_scalar<8><<<grid, 256, 0, st>>>(p);
  }
  NGNN_LAUNCH_CHECK();
  return NGNN_OK;
}

}  // namespace ngnn

namespace ngnn {
int32_t agg_fwd_table_impl(const int32_t* rowptr, const int32_t* col_table, const float* table, int64_t ld_table, Ext n_dst,
                           int64_t F, float* mean, int64_t ld_mean, const int32_t* root_table, float* root, int64_t ld_root,
                           int64_t hot_rows, cudaStream_t st, unsigned long long* clock) {
  if (n_dst.cap == 0 || F == 0) return NGNN_OK;
  NGNN_REQUIRE(rowptr && table && mean && ld_table >= F && ld_mean >= F, NGNN_E_INVALID, "agg_fwd_table: bad arguments");
  // (Measured, round 1: reserving an L2 persisting set-aside for the hot rows — cudaLimitPersistingL2CacheSize — made this
  //  kernel slower, 57 us vs 52 us with the plain cache hints, so no device-wide limit is touched.)
  AggParams p{};
  p.ptr = rowptr; p.idx = col_table; p.x = table; p.ld_x = ld_table; p.n_rows = n_dst.cap; p.n_rows_dev = n_dst.dev; p.F = F;
  p.out = mean; p.ld_out = ld_mean; p.mean = 1;
  p.root_idx = root_table; p.root = root; p.ld_root = ld_root;
  p.keep_l2 = g_tune_keep; p.hot_rows = hot_rows; p.clock = clock;
  return run_agg(p, st);
}

int32_t agg_fwd_impl(const int32_t* rowptr, const int32_t* col, const float* x, int64_t ld_x, Ext n_dst, int64_t F, float* mean,
                     int64_t ld_mean, cudaStream_t st) {
  if (n_dst.cap == 0 || F == 0) return NGNN_OK;
  NGNN_REQUIRE(rowptr && x && mean && ld_x >= F && ld_mean >= F, NGNN_E_INVALID, "agg_fwd: bad arguments");
  AggParams p{};
  p.ptr = rowptr; p.idx = col; p.x = x; p.ld_x = ld_x; p.n_rows = n_dst.cap; p.n_rows_dev = n_dst.dev; p.F = F;
  p.out = mean; p.ld_out = ld_mean; p.mean = 1;
  p.keep_l2 = g_tune_keep; p.hot_rows = -1;
  return run_agg(p, st);
}

int32_t agg_bwd_impl(const int32_t* colptr_t, const int32_t* row_t, const float* dmean_scaled, int64_t ld_dmean, Ext n_src,
                     int64_t F, const float* dx_root, int64_t ld_root, Ext n_root, const float* act_ref, int64_t ld_act,
                     float act_scale, float* dx, int64_t ld_dx, cudaStream_t st) {
  if (n_src.cap == 0 || F == 0) return NGNN_OK;
  NGNN_REQUIRE(colptr_t && dmean_scaled && dx, NGNN_E_INVALID, "agg_bwd: null pointer");
  NGNN_REQUIRE(ld_dmean >= F && ld_dx >= F, NGNN_E_INVALID, "agg_bwd: leading dimension < F");
  NGNN_REQUIRE(dx_root == nullptr || ld_root >= F, NGNN_E_INVALID, "agg_bwd: ld_root < F");
  NGNN_REQUIRE(act_ref == nullptr || ld_act >= F, NGNN_E_INVALID, "agg_bwd: ld_act < F");
  AggParams p{};
  p.ptr = colptr_t; p.idx = row_t; p.x = dmean_scaled; p.ld_x = ld_dmean; p.n_rows = n_src.cap; p.n_rows_dev = n_src.dev; p.F = F;
  p.out = dx; p.ld_out = ld_dx; p.mean = 0;
  p.add = dx_root; p.ld_add = ld_root; p.n_add = dx_root ? n_root.cap : 0; p.n_add_dev = dx_root ? n_root.dev : nullptr;
  p.act_ref = act_ref; p.ld_act = ld_act; p.act_scale = act_scale;
  return run_agg(p, st);
}
}  // namespace ngnn

using namespace ngnn;

extern "C" int32_t ngnn_set_gemm_tile(int32_t bn_max);   // gemm.cu
extern "C" int32_t ngnn_set_gemm_ts(int32_t on);         // gemm.cu
extern "C" int32_t ngnn_set_wgrad_splits(int32_t s);     // gemm.cu

extern "C" {

int32_t ngnn_set_tuning(int32_t key, int32_t value) {
  if (key == 16 && (value == 8 || value == 16 || value == 32)) { ngnn::g_draw_group = value; return NGNN_OK; }
  if (key == 0) { g_tune_unroll = value; return NGNN_OK; }
  if (key == 1 && (value == 128 || value == 256 || value == 512)) { g_tune_threads = value; return NGNN_OK; }
  if (key == 2 && (value == 32 || value == 16 || value == 8)) { g_tune_group = value; return NGNN_OK; }
  if (key == 3 && (value == 0 || value == 1)) { g_tune_pipe = value; return NGNN_OK; }
  if (key == 4 && (value == 128 || value == 256)) return ngnn_set_gemm_tile(value);
  if (key == 8 && value >= 0 && value <= 4096) return ngnn_set_wgrad_splits(value);
  if (key == 9 && value >= 0 && value <= 2) { g_tune_wide = value; return NGNN_OK; }
  if (key == 10 && (value == 0 || value == 1)) { g_use_pdl = value; return NGNN_OK; }
  if (key == 12 && value >= 0 && value <= 2) { g_tune_xwide = value; return NGNN_OK; }
  if (key == 13 && value >= 0 && value <= 2) { g_tune_l1 = value; return NGNN_OK; }
  if (key == 15 && (value == 0 || value == 1)) { g_tune_small = value; return NGNN_OK; }
  if (key == 14 && value >= 0 && value <= 4) { g_tune_stage = value; return NGNN_OK; }   // 2 = also for few rows (tests)
  if (key == 11 && value >= 32 && value <= (1 << 20)) { g_tune_long = value; return NGNN_OK; }
  if (key == 6 && (value == 0 || value == 1)) return ngnn_set_gemm_ts(value);
  if (key == 7 && (value == 0 || value == 1)) { g_tune_keep = value; return NGNN_OK; }
  return ngnn::set_error(NGNN_E_INVALID, "set_tuning: unknown key/value %d/%d", key, value);
}

int32_t ngnn_sage_agg_fwd(const int32_t* rowptr, const int32_t* col, const float* x, int64_t ld_x, int64_t n_dst,
                          int64_t F, float* mean, int64_t ld_mean, const int32_t* root_idx, float* root,
                          int64_t ld_root, ngnn_stream_t stream) {
  NGNN_REQUIRE(n_dst >= 0 && F >= 0, NGNN_E_INVALID, "agg_fwd: negative size");
  if (n_dst == 0 || F == 0) return NGNN_OK;
  NGNN_REQUIRE(rowptr && x && mean, NGNN_E_INVALID, "agg_fwd: null pointer");
  NGNN_REQUIRE(ld_x >= F && ld_mean >= F, NGNN_E_INVALID, "agg_fwd: leading dimension < F");
  NGNN_REQUIRE(root_idx == nullptr || (root != nullptr && ld_root >= F), NGNN_E_INVALID, "agg_fwd: root buffer missing");
  AggParams p{};
  p.ptr = rowptr; p.idx = col; p.x = x; p.ld_x = ld_x; p.n_rows = n_dst; p.F = F;
  p.out = mean; p.ld_out = ld_mean; p.mean = 1;
  p.root_idx = root_idx; p.root = root; p.ld_root = ld_root;
  p.keep_l2 = g_tune_keep; p.hot_rows = -1;
  return run_agg(p, as_stream(stream));
}

int32_t ngnn_gcn_agg_fwd(const int32_t* rowptr, const int32_t* col, const float* z, int64_t ld_z, int64_t n_dst, int64_t O,
                         const float* bias, float* out, int64_t ld_out, ngnn_stream_t stream) {
  NGNN_REQUIRE(n_dst >= 0 && O >= 0, NGNN_E_INVALID, "gcn_agg_fwd: negative size");
  if (n_dst == 0 || O == 0) return NGNN_OK;
  NGNN_REQUIRE(rowptr && z && out, NGNN_E_INVALID, "gcn_agg_fwd: null pointer");
  NGNN_REQUIRE(ld_z >= O && ld_out >= O, NGNN_E_INVALID, "gcn_agg_fwd: leading dimension < O");
  AggParams p{};
  p.ptr = rowptr; p.idx = col; p.x = z; p.ld_x = ld_z; p.n_rows = n_dst; p.F = O;
  p.out = out; p.ld_out = ld_out; p.mean = 0; p.bias = bias;
  return run_agg(p, as_stream(stream));
}

int32_t ngnn_sage_agg_bwd(const int32_t* colptr_t, const int32_t* row_t, const float* dmean_scaled, int64_t ld_dmean,
                          int64_t n_src, int64_t F, const float* dx_root, int64_t ld_root, int64_t n_root,
                          const float* act_ref, int64_t ld_act, float act_scale, float* dx, int64_t ld_dx,
                          ngnn_stream_t stream) {
  NGNN_REQUIRE(n_src >= 0 && F >= 0 && n_root >= 0, NGNN_E_INVALID, "agg_bwd: negative size");
  if (n_src == 0 || F == 0) return NGNN_OK;
  return agg_bwd_impl(colptr_t, row_t, dmean_scaled, ld_dmean, ext_host(n_src), F, dx_root, ld_root, ext_host(n_root), act_ref,
                      ld_act, act_scale, dx, ld_dx, as_stream(stream));
}

}  // extern "C"
