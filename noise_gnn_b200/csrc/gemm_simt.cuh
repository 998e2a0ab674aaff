// gemm_simt.cuh — fp32 SIMT GEMM used for shapes the tcgen05 path does not cover
// (rows that are not 16-byte addressable: F = 1433, 767; tiny n) and as the in-library
// cross-check of the tensor-core path.  C[M,N] = sum over up to two operand pairs
//   C[m,n] = sum_k A1(m,k)*B1(n,k) + sum_k A2(m,k)*B2(n,k)
// with arbitrary (row, k) strides, so the same kernel serves the forward projection
// (K-contiguous operands), dgrad (B = W read N-contiguous) and wgrad (both operands
// M/N-contiguous, split over the long k = n dimension).
#pragma once
#include "common.cuh"

namespace ngnn {

struct GemmOperand {
  const float* p;   // element (i,k) at p[i*s_i + k*s_k]; p == nullptr => pair absent
  int64_t s_i;
  int64_t s_k;
};

struct SimtGemmParams {
  GemmOperand A1, B1, A2, B2;
  int64_t K1, K2;
  int64_t M, N;
  float* C;                  // C[z*split_stride + m*ldc + n]
  int64_t ldc;
  int64_t split_stride;      // elements between split-K partials (0 when gridDim.z == 1)
  int32_t tiles_per_split;   // k-tiles (of BK) per z slice
  const float* bias;         // [N] or null
  int32_t act;               // NGNN_ACT_*
  float drop_p;              // 0 => no dropout
  uint32_t seed_lo, seed_hi, off_lo, off_hi;
  const int32_t* rowptr_scale;  // optional: row m scaled by 1/max(rowptr[m+1]-rowptr[m],1)
  const int32_t* M_dev;         // optional device-side row count (clamped to M: the grid is sized for M)
  const int32_t* K1_dev;        // optional device-side length of the first contraction (weight gradients: the block's rows)
  const StepCtl* ctl;           // optional device-side step control: dropout offset = ctl->drop_off + ctl_layer
  uint32_t ctl_layer;
};

// Keep-mask of the fused inverted dropout.  One Philox4x32-10 call (128 random bits) serves the 8 columns
// [8*cg, 8*cg+8) of row m, 16 bits each: column 8*cg + h uses halfword h (word h/2, low half for even h) and is kept
// iff halfword >= floor(p * 2^16).  (32 bits per element made the tensor-core epilogue spend more instructions on
// Philox than on everything else; 16 bits resolve p to 1.5e-5.)  oracle/philox.py::dropout_keep_mask restates it.
__device__ __forceinline__ uint32_t dropout_keep8(uint32_t m, uint32_t cg, uint32_t seed_lo, uint32_t seed_hi,
                                                  uint32_t off_lo, uint32_t off_hi, uint32_t thr16) {
  const Philox4 r = philox4x32_10(m, cg, off_lo, off_hi, seed_lo, seed_hi);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
  uint32_t keep = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    keep |= ((w[i] & 0xFFFFu) >= thr16 ? 1u : 0u) << (2 * i);
    keep |= ((w[i] >> 16) >= thr16 ? 1u : 0u) << (2 * i + 1);
  }
  return keep;
}
// the 4 columns [4*cq, 4*cq+4): low or high nibble of the group's 8 bits
__device__ __forceinline__ uint32_t dropout_keep4(uint32_t m, uint32_t cq, uint32_t seed_lo, uint32_t seed_hi,
                                                  uint32_t off_lo, uint32_t off_hi, uint32_t thr16) {
  return (dropout_keep8(m, cq >> 1, seed_lo, seed_hi, off_lo, off_hi, thr16) >> ((cq & 1u) * 4u)) & 0xFu;
}
__host__ __device__ __forceinline__ uint32_t dropout_threshold(float p) {
  double t = (double)p * 65536.0;
  if (t < 0.0) t = 0.0;
  if (t > 65535.0) t = 65535.0;
  return (uint32_t)t;
}

constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 16, SG_PAD = 4, SG_T = 256;

__device__ __forceinline__ void simt_load_tile(const GemmOperand& op, int64_t i0, int64_t i_max, int64_t k0,
                                               int64_t k_max, float (&r)[4]) {
  const int t = threadIdx.x;
  const bool kc = (op.s_k == 1);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int idx = t + SG_T * q;
    const int k = kc ? (idx & (SG_BK - 1)) : (idx >> 6);
    const int i = kc ? (idx >> 4) : (idx & (SG_BM - 1));
    const int64_t gi = i0 + i, gk = k0 + k;
    r[q] = (gi < i_max && gk < k_max) ? __ldg(op.p + gi * op.s_i + gk * op.s_k) : 0.f;
  }
}

__device__ __forceinline__ void simt_store_tile(const GemmOperand& op, float (*S)[SG_BM + SG_PAD], const float (&r)[4]) {
  const int t = threadIdx.x;
  const bool kc = (op.s_k == 1);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int idx = t + SG_T * q;
    const int k = kc ? (idx & (SG_BK - 1)) : (idx >> 6);
    const int i = kc ? (idx >> 4) : (idx & (SG_BM - 1));
    S[k][i] = r[q];
  }
}

__global__ void __launch_bounds__(SG_T) k_gemm_simt(SimtGemmParams p) {
  __shared__ __align__(16) float As[SG_BK][SG_BM + SG_PAD];
  __shared__ __align__(16) float Bs[SG_BK][SG_BN + SG_PAD];
  const int t = threadIdx.x;
  const int ty = t >> 4, tx = t & 15;
  const int64_t m0 = (int64_t)blockIdx.x * SG_BM, n0 = (int64_t)blockIdx.y * SG_BN;

  int64_t M = p.M, K1 = p.K1;
  if (p.M_dev != nullptr) { const int64_t v = __ldg(p.M_dev); M = v < M ? (v < 0 ? 0 : v) : M; }
  if (p.K1_dev != nullptr) { const int64_t v = __ldg(p.K1_dev); K1 = v < K1 ? (v < 0 ? 0 : v) : K1; }
  const int32_t T1 = p.A1.p ? (int32_t)((K1 + SG_BK - 1) / SG_BK) : 0;
  const int32_t T2 = p.A2.p ? (int32_t)((p.K2 + SG_BK - 1) / SG_BK) : 0;
  const int32_t t_beg = blockIdx.z * p.tiles_per_split;
  const int32_t t_end = min(T1 + T2, t_beg + p.tiles_per_split);

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float ra[4], rb[4];
  auto fetch = [&](int32_t tile) {
    if (tile < T1) {
      simt_load_tile(p.A1, m0, M, (int64_t)tile * SG_BK, K1, ra);
      simt_load_tile(p.B1, n0, p.N, (int64_t)tile * SG_BK, K1, rb);
    } else {
      simt_load_tile(p.A2, m0, M, (int64_t)(tile - T1) * SG_BK, p.K2, ra);
      simt_load_tile(p.B2, n0, p.N, (int64_t)(tile - T1) * SG_BK, p.K2, rb);
    }
  };

  if (t_beg < t_end) fetch(t_beg);
  for (int32_t tile = t_beg; tile < t_end; ++tile) {
    const GemmOperand& opa = tile < T1 ? p.A1 : p.A2;
    const GemmOperand& opb = tile < T1 ? p.B1 : p.B2;
    simt_store_tile(opa, As, ra);
    simt_store_tile(opb, Bs, rb);
    __syncthreads();
    if (tile + 1 < t_end) fetch(tile + 1);
#pragma unroll
    for (int k = 0; k < SG_BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue ----
  float* C = p.C + (int64_t)blockIdx.z * p.split_stride;
  const uint32_t thr = dropout_threshold(p.drop_p);
  const float keep_scale = p.drop_p > 0.f ? 1.0f / (1.0f - p.drop_p) : 1.0f;
  uint32_t off_lo = p.off_lo, off_hi = p.off_hi;
  if (p.ctl != nullptr) {
    const uint64_t o = (((uint64_t)p.ctl->drop_off_hi << 32) | p.ctl->drop_off_lo) + p.ctl_layer;
    off_lo = (uint32_t)o; off_hi = (uint32_t)(o >> 32);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
    float rs = 1.0f;
    if (p.rowptr_scale) rs = 1.0f / (float)max(__ldg(p.rowptr_scale + m + 1) - __ldg(p.rowptr_scale + m), 1);
    const int64_t nb = n0 + tx * 4;
    uint32_t keep = 0xFu;
    if (p.drop_p > 0.f && nb < p.N)
      keep = dropout_keep4((uint32_t)m, (uint32_t)(nb >> 2), p.seed_lo, p.seed_hi, off_lo, off_hi, thr);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t n = nb + j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (p.bias) v += __ldg(p.bias + n);
      v *= rs;
      if (p.act == NGNN_ACT_RELU) v = fmaxf(v, 0.f);
      if (p.drop_p > 0.f) v = ((keep >> j) & 1u) ? v * keep_scale : 0.f;
      C[m * p.ldc + n] = v;
    }
  }
}

// out[i] (+)= sum_z part[z*stride + i].  Fixed association order ((z0+z4+...) + (z1+z5+...) + ...) => deterministic;
// four independent accumulators keep four loads in flight per thread.
__global__ void k_reduce_partials(const float* __restrict__ part, int64_t stride, int32_t splits, int64_t n,
                                  float* __restrict__ out, int32_t accumulate) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int32_t z = 0;
  for (; z + 3 < splits; z += 4) {
    s0 += part[(int64_t)z * stride + i];
    s1 += part[(int64_t)(z + 1) * stride + i];
    s2 += part[(int64_t)(z + 2) * stride + i];
    s3 += part[(int64_t)(z + 3) * stride + i];
  }
  for (; z < splits; ++z) s0 += part[(int64_t)z * stride + i];
  const float s = (s0 + s1) + (s2 + s3);
  out[i] = accumulate ? out[i] + s : s;
}

// Up to three independent partial sets reduced by one launch (blockIdx.y selects): dW_l, dW_r and db of one K-WGRAD call.
struct ReduceJob {
  const float* part;     // [splits][stride]
  float* out;            // [n]
  int64_t n, stride;
};
struct ReduceJobs {
  ReduceJob job[3];
  int32_t n;
};
__global__ void k_reduce_partials_multi(const ReduceJobs jobs, int32_t splits, int32_t accumulate) {
  pdl_trigger();
  pdl_wait();
  const ReduceJob jb = jobs.job[blockIdx.y];
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= jb.n) return;
  // eight independent chains: the loads of a slice round are all in flight together (the kernel is pure latency)
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  int32_t z = 0;
  for (; z + 7 < splits; z += 8) {
#pragma unroll
    for (int q = 0; q < 8; ++q) s[q] += __ldg(jb.part + (int64_t)(z + q) * jb.stride + i);
  }
#pragma unroll
  for (int q = 0; q < 8; ++q)
    if (z + q < splits) s[q] += __ldg(jb.part + (int64_t)(z + q) * jb.stride + i);
  const float t = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
  jb.out[i] = accumulate ? jb.out[i] + t : t;
}

// column sums of dy over a row slice: part[z*O + o] = sum_{i in slice z} dy[i*ld + o].
// Block = 32 columns x 8 row lanes: every warp reads 128 contiguous bytes of a row; the 8 row lanes are combined
// through shared memory in a fixed order (deterministic).
__global__ void __launch_bounds__(256) k_colsum_partial(const float* __restrict__ dy, int64_t ld, Ext n_ext, int64_t O,
                                                        float* __restrict__ part) {
  __shared__ float sm[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t o = (int64_t)blockIdx.x * 32 + tx;
  const int64_t n = ext_get(n_ext);
  const int64_t rows_per_slice = (n + gridDim.y - 1) / gridDim.y;
  const int64_t i0 = min(n, (int64_t)blockIdx.y * rows_per_slice);
  const int64_t i1 = min(n, i0 + rows_per_slice);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (o < O) {
    int64_t i = i0 + ty;
    for (; i + 56 < i1; i += 64) {      // 8 independent loads in flight per thread
      const float a0 = __ldg(dy + i * ld + o), a1 = __ldg(dy + (i + 8) * ld + o);
      const float a2 = __ldg(dy + (i + 16) * ld + o), a3 = __ldg(dy + (i + 24) * ld + o);
      const float a4 = __ldg(dy + (i + 32) * ld + o), a5 = __ldg(dy + (i + 40) * ld + o);
      const float a6 = __ldg(dy + (i + 48) * ld + o), a7 = __ldg(dy + (i + 56) * ld + o);
      s0 += a0; s1 += a1; s2 += a2; s3 += a3;
      s0 += a4; s1 += a5; s2 += a6; s3 += a7;
    }
    for (; i + 24 < i1; i += 32) {
      s0 += __ldg(dy + i * ld + o);
      s1 += __ldg(dy + (i + 8) * ld + o);
      s2 += __ldg(dy + (i + 16) * ld + o);
      s3 += __ldg(dy + (i + 24) * ld + o);
    }
    for (; i < i1; i += 8) s0 += __ldg(dy + i * ld + o);
  }
  sm[ty][tx] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (ty == 0 && o < O) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) s += sm[r][tx];
    part[(int64_t)blockIdx.y * O + o] = s;
  }
}

__global__ void k_act_bwd(const float* __restrict__ dh, int64_t ld_dh, const float* __restrict__ h, int64_t ld_h,
                          int64_t n, int64_t O, float scale, float* __restrict__ dz, int64_t ld_dz) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= n * O) return;
  const int64_t i = idx / O, o = idx - i * O;
  dz[i * ld_dz + o] = h[i * ld_h + o] > 0.f ? dh[i * ld_dh + o] * scale : 0.f;
}

static inline int32_t launch_simt_gemm(SimtGemmParams& p, int32_t splits, cudaStream_t st) {
  const int32_t T1 = p.A1.p ? (int32_t)ceil_div(p.K1, SG_BK) : 0;
  const int32_t T2 = p.A2.p ? (int32_t)ceil_div(p.K2, SG_BK) : 0;
  if (splits < 1) splits = 1;
  p.tiles_per_split = (int32_t)ceil_div((int64_t)(T1 + T2 > 0 ? T1 + T2 : 1), splits);
  dim3 grid((unsigned)ceil_div(p.M, SG_BM), (unsigned)ceil_div(p.N, SG_BN), (unsigned)splits);
  NGNN_REQUIRE(grid.y <= 65535u && grid.z <= 65535u, NGNN_E_UNSUPPORTED, "simt gemm: grid too large (N=%lld)", (long long)p.N);
  k_gemm_simt<<<grid, SG_T, 0, st>>>(p);
  NGNN_LAUNCH_CHECK();
  return NGNN_OK;
}

}  // namespace ngnn
