// agg_bulk.cuh — K-AGG forward with the feature rows staged through shared memory by the async copy engines.
//
// Same arithmetic as k_agg_fwd_pipe (agg.cu): mean[i] = 1/max(deg_i,1) * sum_p x[col[p]], optional fused root gather.
// The register-pipelined kernel keeps U neighbour rows per lane in flight and its bytes-in-flight are bounded by the
// register file; here every warp owns a ring of S shared-memory slots and the gathers are asynchronous copies
// global -> shared that complete on an mbarrier, so a warp has (S-1) units of up to CH+1 rows in flight while it
// sums an earlier unit out of shared memory — bytes in flight are bounded by shared memory (227 KB / SM), not registers.
//   MODE 0: one `cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes` per neighbour row, issued by one
//           lane each (SASS UBLKCP — the TMA engine's 1-D path; rows are 16-byte multiples, 16-byte aligned);
//   MODE 1: `cp.async.cg.shared.global` 16 B per lane (LDGSTS) + `cp.async.mbarrier.arrive.noinc`.
// A unit is a chunk of <= CH neighbours of one row (long rows span several units, the partial sum stays in
// registers), plus the row's own feature row on its first unit when ROOT.  Extents are prefetched three rows ahead
// and index windows two rows ahead in registers, exactly like the register-pipelined kernel.  No atomics; the
// summation order inside a row is the stored order => bitwise equal to the other K-AGG kernels.
#pragma once
#include "common.cuh"

namespace ngnn {

__device__ __forceinline__ uint32_t agg_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void agg_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(agg_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void agg_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(agg_smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void agg_mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = agg_smem_u32(bar);
  uint32_t done = 0;
  // bounded: a protocol bug traps instead of hanging the GPU
  for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
    if (done) return;
  }
  __trap();
}
__device__ __forceinline__ void agg_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(agg_smem_u32(dst)), "l"(src), "r"(bytes), "r"(agg_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void agg_ldgsts16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(agg_smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void agg_ldgsts_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(agg_smem_u32(bar)) : "memory");
}

constexpr int kBulkWarps = 8;

__host__ __device__ inline size_t agg_bulk_smem_bytes(int CH, int S, int64_t F) {
  return (size_t)kBulkWarps * S * (8 + 16) + (size_t)kBulkWarps * S * (CH + 1) * (size_t)F * 4;
}

template <int CH, int S, int MODE, bool ROOT>
__global__ void __launch_bounds__(kBulkWarps * 32, 1) k_agg_fwd_bulk(AggParams p) {
  static_assert(32 % CH == 0, "a chunk never straddles a 32-index window");
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int RB = (int)p.F * 4;                          // row bytes (multiple of 16)
  const int F4 = (int)(p.F >> 2);                       // <= 32: one float4 per lane
  const int slot_bytes = (CH + 1) * RB;
  uint64_t* wbar = reinterpret_cast<uint64_t*>(smem) + wib * S;
  int4* wmeta = reinterpret_cast<int4*>(smem + kBulkWarps * S * 8) + wib * S;
  unsigned char* wdata = smem + kBulkWarps * S * 24 + (size_t)wib * S * slot_bytes;
  if (lane < S) agg_mbar_init(&wbar[lane], MODE == 0 ? 1u : 32u);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();

  const int64_t n = p.n_rows;
  const int64_t W = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const unsigned full = 0xffffffffu;

  // ---- issue-side cursor: row ri (extents ib..ie, root id rid, index window iwin covering [iwb, iwb+32)) ----
  int64_t ri = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int64_t r1 = ri + W, r2 = r1 + W, r3 = r2 + W;
  int ib = 0, ie = 0, rid = 0, b1 = 0, e1 = 0, rid1 = 0, b2 = 0, e2 = 0, rid2 = 0, b3 = 0, e3 = 0, rid3 = 0;
  if (ri < n) { ib = __ldg(p.ptr + ri); ie = __ldg(p.ptr + ri + 1); if (ROOT) rid = __ldg(p.root_idx + ri); }
  if (r1 < n) { b1 = __ldg(p.ptr + r1); e1 = __ldg(p.ptr + r1 + 1); if (ROOT) rid1 = __ldg(p.root_idx + r1); }
  if (r2 < n) { b2 = __ldg(p.ptr + r2); e2 = __ldg(p.ptr + r2 + 1); if (ROOT) rid2 = __ldg(p.root_idx + r2); }
  if (r3 < n) { b3 = __ldg(p.ptr + r3); e3 = __ldg(p.ptr + r3 + 1); if (ROOT) rid3 = __ldg(p.root_idx + r3); }
  int iwin = (ri < n && ib + lane < ie) ? __ldg(p.idx + ib + lane) : 0;
  int win1 = (r1 < n && b1 + lane < e1) ? __ldg(p.idx + b1 + lane) : 0;
  int win2 = (r2 < n && b2 + lane < e2) ? __ldg(p.idx + b2 + lane) : 0;
  int ipos = ib, iwb = ib;
  uint32_t iu = 0, cu = 0;                              // units issued / consumed by this warp

  auto issue = [&]() -> bool {
    if (ri >= n) return false;
    const int slot = (int)(iu % S);
    if (ipos - iwb >= 32) {                             // rows longer than a window (not the sampled-block case)
      iwb += 32;
      iwin = (iwb + lane < ie) ? __ldg(p.idx + iwb + lane) : 0;
    }
    const int cnt = min(CH, ie - ipos);
    const bool first = ipos == ib, last = ipos + cnt >= ie;
    const int nb = __shfl_sync(full, iwin, (ipos - iwb + lane) & 31);
    unsigned char* sdst = wdata + (size_t)slot * slot_bytes;
    const bool want_root = ROOT && first;
    if (MODE == 0) {
      if (lane == 0) {
        wmeta[slot] = make_int4((int)ri, cnt, (first ? 1 : 0) | (last ? 2 : 0), ie - ib);
        agg_mbar_expect_tx(&wbar[slot], (uint32_t)((cnt + (want_root ? 1 : 0)) * RB));
      }
      __syncwarp();
      if (lane < cnt) agg_bulk_g2s(sdst + lane * RB, p.x + (int64_t)nb * p.ld_x, (uint32_t)RB, &wbar[slot]);
      else if (want_root && lane == CH) agg_bulk_g2s(sdst + CH * RB, p.x + (int64_t)rid * p.ld_x, (uint32_t)RB, &wbar[slot]);
    } else {
      if (lane == 0) wmeta[slot] = make_int4((int)ri, cnt, (first ? 1 : 0) | (last ? 2 : 0), ie - ib);
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        const int sj = __shfl_sync(full, nb, j);
        if (j < cnt && lane < F4) agg_ldgsts16(sdst + j * RB + lane * 16, reinterpret_cast<const float4*>(p.x + (int64_t)sj * p.ld_x) + lane);
      }
      if (want_root && lane < F4) agg_ldgsts16(sdst + CH * RB + lane * 16, reinterpret_cast<const float4*>(p.x + (int64_t)rid * p.ld_x) + lane);
      agg_ldgsts_arrive(&wbar[slot]);
    }
    ipos += cnt;
    ++iu;
    if (last) {                                         // rotate the prefetch registers to the next row of this warp
      ri = r1; ib = b1; ie = e1; rid = rid1; iwin = win1;
      r1 = r2; b1 = b2; e1 = e2; rid1 = rid2; win1 = win2;
      r2 = r3; b2 = b3; e2 = e3; rid2 = rid3;
      win2 = (r2 < n && b2 + lane < e2) ? __ldg(p.idx + b2 + lane) : 0;
      r3 = r2 + W; b3 = 0; e3 = 0; rid3 = 0;
      if (r3 < n) { b3 = __ldg(p.ptr + r3); e3 = __ldg(p.ptr + r3 + 1); if (ROOT) rid3 = __ldg(p.root_idx + r3); }
      ipos = ib; iwb = ib;
    }
    return true;
  };

  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  auto consume = [&]() {
    const int slot = (int)(cu % S);
    agg_mbar_wait(&wbar[slot], (cu / S) & 1u);
    const int4 m = wmeta[slot];
    const unsigned char* s = wdata + (size_t)slot * slot_bytes + lane * 16;
    if (m.z & 1) acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane < F4) {
      float4 v[CH];
#pragma unroll
      for (int j = 0; j < CH; ++j) v[j] = (j < m.y) ? *reinterpret_cast<const float4*>(s + j * RB) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < CH; ++j) if (j < m.y) f4_add(acc, v[j]);
      if (ROOT && (m.z & 1))
        reinterpret_cast<float4*>(p.root + (int64_t)m.x * p.ld_root)[lane] = *reinterpret_cast<const float4*>(s + CH * RB);
      if (m.z & 2) {
        const float scale = p.mean ? 1.0f / (float)max(m.w, 1) : 1.0f;
        reinterpret_cast<float4*>(p.out + (int64_t)m.x * p.ld_out)[lane] =
            make_float4(acc.x * scale, acc.y * scale, acc.z * scale, acc.w * scale);
      }
    }
    __syncwarp();
    ++cu;
  };

#pragma unroll 1
  for (int k = 0; k < S - 1; ++k) if (!issue()) break;
#pragma unroll 1
  while (cu < iu) {
    issue();
    consume();
  }
}

}  // namespace ngnn
