// sagepl.cu — the SAGEPL extras of the reference on the device:
//
//   ngnn_noise_add_fwd / _bwd   SAGEPL.adding_noise (reference src/models/layers/sagePL.py:41-49): the learnable per-node
//                               noise rows are gathered by the block's n_id, L2-normalised (F.normalize, eps 1e-12),
//                               scaled by noise_rate and added to the features — gather + normalize + scale + add in ONE
//                               pass (the reference issues clone, index_select, norm, clamp, div, mul, add), and its
//                               backward (the projection rate/|v| (g - u u.g) scattered back to the gathered rows).
//   ngnn_shuffle_rows           shuffle_pos (reference src/utils/augmentation.py:88-102): per row, k = int(F * prob) distinct
//                               random positions get their values permuted among themselves.  The reference loops over
//                               the rows in Python (two torch.randperm calls per row: minutes for a products block);
//                               here one warp per row draws the subset (Robert Floyd) and the permutation (Fisher-Yates)
//                               from a counter-based Philox stream keyed (seed, call offset, row).
// One warp per row, coalesced 128-bit accesses where the rows allow, no atomics for distinct ids.
#include "common.cuh"

namespace ngnn {

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr float kNormEps = 1e-12f;     // torch.nn.functional.normalize default

// out[i,:] = x[i,:] + s ⊙ rate * v / max(|v|, eps),  v = noise[idx[i],:],  s = sign(x[i,:]) or 1
__global__ void __launch_bounds__(256) k_noise_add_fwd(const float* __restrict__ x, int64_t ld_x, const float* __restrict__ noise,
                                                       int64_t ld_noise, const int32_t* __restrict__ idx, int64_t n, int64_t F,
                                                       float rate, int32_t use_sign, float* __restrict__ out, int64_t ld_out) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (i >= n) return;
  const float* v = noise + (idx != nullptr ? (int64_t)__ldg(idx + i) : i) * ld_noise;
  float ss = 0.f;
  for (int64_t c = lane; c < F; c += 32) { const float t = __ldg(v + c); ss = fmaf(t, t, ss); }
  ss = warp_sum_f(ss);
  const float scale = rate / fmaxf(sqrtf(ss), kNormEps);
  const float* xr = x + i * ld_x;
  float* o = out + i * ld_out;
  for (int64_t c = lane; c < F; c += 32) {
    const float xv = xr[c];
    const float s = use_sign ? (xv > 0.f ? 1.f : (xv < 0.f ? -1.f : 0.f)) : 1.f;
    o[c] = fmaf(s * scale, __ldg(v + c), xv);
  }
}

// dnoise[idx[i],:] (+)= rate/|v| * (g - u (u.g)),  g = s ⊙ dout[i,:], u = v/|v|   (|v| < eps: rate/eps * g)
__global__ void __launch_bounds__(256) k_noise_add_bwd(const float* __restrict__ dout, int64_t ld_d, const float* __restrict__ x,
                                                       int64_t ld_x, const float* __restrict__ noise, int64_t ld_noise,
                                                       const int32_t* __restrict__ idx, int64_t n, int64_t F, float rate,
                                                       int32_t use_sign, float* __restrict__ dnoise, int64_t ld_dn) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (i >= n) return;
  const int64_t r = idx != nullptr ? (int64_t)__ldg(idx + i) : i;
  const float* v = noise + r * ld_noise;
  const float* g = dout + i * ld_d;
  const float* xr = x + i * ld_x;
  float ss = 0.f, vg = 0.f;
  for (int64_t c = lane; c < F; c += 32) {
    const float t = __ldg(v + c);
    float gv = g[c];
    if (use_sign) { const float xv = xr[c]; gv *= xv > 0.f ? 1.f : (xv < 0.f ? -1.f : 0.f); }
    ss = fmaf(t, t, ss);
    vg = fmaf(t, gv, vg);
  }
  ss = warp_sum_f(ss);
  vg = warp_sum_f(vg);
  const float nrm = sqrtf(ss);
  const bool tiny = nrm < kNormEps;
  const float a = rate / fmaxf(nrm, kNormEps);
  const float b = tiny ? 0.f : vg / ss;                // (u.g)/|v| = (v.g)/|v|^2
  float* dn = dnoise + r * ld_dn;
  for (int64_t c = lane; c < F; c += 32) {
    float gv = g[c];
    if (use_sign) { const float xv = xr[c]; gv *= xv > 0.f ? 1.f : (xv < 0.f ? -1.f : 0.f); }
    atomicAdd(dn + c, a * (gv - b * __ldg(v + c)));     // distinct ids (a block's n_id): every address is touched once
  }
}

// bounded uniform integer in [0, m) from a 32-bit word (same multiply-high law as the sampler)
__device__ __forceinline__ uint32_t bounded(uint32_t w, uint32_t m) { return mulhi32(w, m); }

constexpr int kShuffleMaxF = 2048;     // positions fit int16 lists in shared memory
constexpr int kShuffleWarps = 4;

// One warp per row.  lane 0 runs the (sequential, k-step) subset + permutation; the warp copies the row.
//   positions: Robert Floyd's algorithm over [0, F): for j = F-k .. F-1: t = U[0, j]; take t unless taken, else j   (insertion order kept)
//   permutation: Fisher-Yates over the k selected slots: for j = k-1 .. 1: swap(sel[j], sel[U[0, j]])
//   out[row, pos[j]] = x[row, sel[j]]    (pos = insertion order, sel = shuffled copy)
// Philox4x32-10 key (seed), counter (row, word index / 4, offset): word q of the row's stream = draw q; the subset uses words
// 0..k-1, the permutation words k..2k-2.
__global__ void __launch_bounds__(32 * kShuffleWarps) k_shuffle_rows(const float* __restrict__ x, int64_t ld_x, int64_t n, int32_t F,
                                                                     int32_t k, uint32_t seed_lo, uint32_t seed_hi, uint32_t off_lo,
                                                                     uint32_t off_hi, float* __restrict__ out, int64_t ld_out) {
  __shared__ uint32_t s_taken[kShuffleWarps][kShuffleMaxF / 32];
  __shared__ int16_t s_pos[kShuffleWarps][kShuffleMaxF];
  __shared__ int16_t s_sel[kShuffleWarps][kShuffleMaxF];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t row = (int64_t)blockIdx.x * kShuffleWarps + w;
  if (row >= n) return;
  const float* xr = x + row * ld_x;
  float* o = out + row * ld_out;
  for (int32_t c = lane; c < F; c += 32) o[c] = xr[c];
  for (int32_t c = lane; c < (F + 31) / 32; c += 32) s_taken[w][c] = 0u;
  __syncwarp();
  if (lane == 0 && k > 1) {
    Philox4 r{0, 0, 0, 0};
    auto word = [&](uint32_t q) {
      if ((q & 3u) == 0u) r = philox4x32_10((uint32_t)row, q >> 2, off_lo, off_hi ^ (uint32_t)(row >> 32), seed_lo, seed_hi);
      return (q & 3u) == 0u ? r.x : (q & 3u) == 1u ? r.y : (q & 3u) == 2u ? r.z : r.w;
    };
    uint32_t q = 0;
    for (int32_t j = 0; j < k; ++j, ++q) {
      const int32_t jj = F - k + j;
      int32_t t = (int32_t)bounded(word(q), (uint32_t)(jj + 1));
      if (s_taken[w][t >> 5] & (1u << (t & 31))) t = jj;
      s_taken[w][t >> 5] |= 1u << (t & 31);
      s_pos[w][j] = (int16_t)t;
      s_sel[w][j] = (int16_t)t;
    }
    if ((q & 3u) != 0u) q = (q + 3u) & ~3u;            // the permutation starts on a fresh Philox block
    for (int32_t j = k - 1; j >= 1; --j, ++q) {
      const int32_t u = (int32_t)bounded(word(q), (uint32_t)(j + 1));
      const int16_t tmp = s_sel[w][j]; s_sel[w][j] = s_sel[w][u]; s_sel[w][u] = tmp;
    }
  }
  __syncwarp();
  if (k > 1)
    for (int32_t j = lane; j < k; j += 32) o[s_pos[w][j]] = xr[s_sel[w][j]];
}

}  // namespace ngnn

using namespace ngnn;

extern "C" {

int32_t ngnn_noise_add_fwd(const float* x, int64_t ld_x, const float* noise, int64_t ld_noise, const int32_t* idx, int64_t n,
                           int64_t F, float rate, int32_t use_sign, float* out, int64_t ld_out, ngnn_stream_t stream) {
  NGNN_REQUIRE(n >= 0 && F >= 0, NGNN_E_INVALID, "noise_add_fwd: negative size");
  if (n == 0 || F == 0) return NGNN_OK;
  NGNN_REQUIRE(x && noise && out, NGNN_E_INVALID, "noise_add_fwd: null pointer");
  NGNN_REQUIRE(ld_x >= F && ld_noise >= F && ld_out >= F, NGNN_E_INVALID, "noise_add_fwd: leading dimension < F");
  k_noise_add_fwd<<<(unsigned)ceil_div(n * 32, 256), 256, 0, as_stream(stream)>>>(x, ld_x, noise, ld_noise, idx, n, F, rate, use_sign,
                                                                               out, ld_out);
  NGNN_LAUNCH_CHECK();
  return NGNN_OK;
}

int32_t ngnn_noise_add_bwd(const float* dout, int64_t ld_d, const float* x, int64_t ld_x, const float* noise, int64_t ld_noise,
                           const int32_t* idx, int64_t n, int64_t F, float rate, int32_t use_sign, float* dnoise, int64_t ld_dn,
                           ngnn_stream_t stream) {
  NGNN_REQUIRE(n >= 0 && F >= 0, NGNN_E_INVALID, "noise_add_bwd: negative size");
  if (n == 0 || F == 0) return NGNN_OK;
  NGNN_REQUIRE(dout && noise && dnoise && (x || !use_sign), NGNN_E_INVALID, "noise_add_bwd: null pointer");
  NGNN_REQUIRE(ld_d >= F && ld_noise >= F && ld_dn >= F && (!use_sign || ld_x >= F), NGNN_E_INVALID,
               "noise_add_bwd: leading dimension < F");
  k_noise_add_bwd<<<(unsigned)ceil_div(n * 32, 256), 256, 0, as_stream(stream)>>>(dout, ld_d, x ? x : dout, x ? ld_x : ld_d, noise,
                                                                               ld_noise, idx, n, F, rate, use_sign, dnoise, ld_dn);
  NGNN_LAUNCH_CHECK();
  return NGNN_OK;
}

int32_t ngnn_shuffle_rows(const float* x, int64_t ld_x, int64_t n, int64_t F, int32_t k, uint64_t seed, uint64_t offset, float* out,
                          int64_t ld_out, ngnn_stream_t stream) {
  NGNN_REQUIRE(n >= 0 && F >= 0 && k >= 0 && k <= F, NGNN_E_INVALID, "shuffle_rows: bad sizes");
  if (n == 0 || F == 0) return NGNN_OK;
  NGNN_REQUIRE(x && out && x != out, NGNN_E_INVALID, "shuffle_rows: null pointer / in-place call");
  NGNN_REQUIRE(ld_x >= F && ld_out >= F, NGNN_E_INVALID, "shuffle_rows: leading dimension < F");
  NGNN_REQUIRE(F <= kShuffleMaxF, NGNN_E_UNSUPPORTED, "shuffle_rows: F = %lld > %d", (long long)F, kShuffleMaxF);
  k_shuffle_rows<<<(unsigned)ceil_div(n, kShuffleWarps), 32 * kShuffleWarps, 0, as_stream(stream)>>>(
      x, ld_x, n, (int32_t)F, k, (uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)offset, (uint32_t)(offset >> 32), out, ld_out);
  NGNN_LAUNCH_CHECK();
  return NGNN_OK;
}

}  // extern "C"
