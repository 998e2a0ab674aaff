// train_ops.cu — the two small step-tail kernels of the hot loop: softmax cross-entropy on the seed
// rows (reference src/pipeline.py:155-165: F.cross_entropy + argmax accuracy count) and Adam
// (reference src/models/model.py:67-69: torch.optim.Adam(lr), stepped at src/pipeline.py:169).
// Both are tiny (bs x C, ~2e5 parameters); they exist so the step has no host synchronisation and
// runs as a handful of launches over one flat parameter bucket.
#include "common.cuh"
#include <math.h>

namespace ngnn {

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One warp per row: log-sum-exp, gradient row, per-row loss / correct flag into `rows` ([2*bs]).
__global__ void __launch_bounds__(256) k_ce_rows(const float* __restrict__ logits, int64_t ld,
                                                 const int64_t* __restrict__ target, const int64_t* __restrict__ y_true,
                                                 const int32_t* __restrict__ row_ids, int64_t bs, int64_t C,
                                                 float grad_scale, float* __restrict__ rows,
                                                 float* __restrict__ dlogits, int64_t ld_d) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (i >= bs) return;
  const float* row = logits + i * ld;
  float mx = -INFINITY;
  int64_t arg = 0;
  for (int64_t c = lane; c < C; c += 32) {
    const float v = row[c];
    if (v > mx) { mx = v; arg = c; }
  }
  // argmax with first-index tie-break (torch.argmax returns the first maximal index)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float omx = __shfl_xor_sync(0xffffffffu, mx, o);
    const int64_t oarg = __shfl_xor_sync(0xffffffffu, arg, o);
    if (omx > mx || (omx == mx && oarg < arg)) { mx = omx; arg = oarg; }
  }
  float se = 0.f;
  for (int64_t c = lane; c < C; c += 32) se += expf(row[c] - mx);
  se = warp_sum(se);
  const float lse = mx + logf(se);
  const int64_t li = row_ids != nullptr ? (int64_t)row_ids[i] : i;   // label index (global id when gathering)
  const int64_t tgt = target[li];
  if (dlogits != nullptr) {
    const float g = grad_scale / (float)bs;
    for (int64_t c = lane; c < C; c += 32) {
      const float pr = expf(row[c] - lse);
      dlogits[i * ld_d + c] = (pr - (c == tgt ? 1.f : 0.f)) * g;
    }
  }
  if (lane == 0) {
    rows[i] = lse - row[tgt];
    rows[bs + i] = (y_true != nullptr && arg == y_true[li]) ? 1.f : 0.f;
  }
}

// Single CTA: fixed-order tree sum of the per-row values => bitwise reproducible loss.
__global__ void __launch_bounds__(1024) k_ce_reduce(const float* __restrict__ rows, int64_t bs, float* __restrict__ stats) {
  __shared__ float s_l[1024];
  __shared__ float s_c[1024];
  float l = 0.f, c = 0.f;
  for (int64_t i = threadIdx.x; i < bs; i += 1024) { l += rows[i]; c += rows[bs + i]; }
  s_l[threadIdx.x] = l; s_c[threadIdx.x] = c;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) { s_l[threadIdx.x] += s_l[threadIdx.x + o]; s_c[threadIdx.x] += s_c[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { stats[0] += s_l[0] / (float)bs; stats[1] += s_c[0]; }
}

__global__ void k_adam(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ m,
                       float* __restrict__ v, int64_t n, float lr, float beta1, float beta2, float eps,
                       float weight_decay, float grad_scale, const int64_t* __restrict__ step_dev) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double t = (double)(*step_dev + 1);
  const float bc1 = (float)(1.0 - pow((double)beta1, t));
  const float bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, t));
  float g = grad[i] * grad_scale;
  const float p = param[i];
  if (weight_decay != 0.f) g += weight_decay * p;
  const float mi = m[i] + (g - m[i]) * (1.0f - beta1);
  const float vi = beta2 * v[i] + (1.0f - beta2) * g * g;
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  param[i] = p - (lr / bc1) * (mi / denom);
}

__global__ void k_inc_step(int64_t* step_dev) { *step_dev += 1; }

}  // namespace ngnn

using namespace ngnn;

extern "C" {

int32_t ngnn_ce_fwd_bwd_gather(const float* logits, int64_t ld, const int64_t* target, const int64_t* y_true,
                               const int32_t* row_ids, int64_t bs, int64_t C, float grad_scale, float* stats, float* dlogits,
                               int64_t ld_d, float* row_scratch, ngnn_stream_t stream) {
  NGNN_REQUIRE(bs >= 0 && C >= 0, NGNN_E_INVALID, "ce: negative size");
  if (bs == 0 || C == 0) return NGNN_OK;
  NGNN_REQUIRE(logits && target && stats && row_scratch, NGNN_E_INVALID, "ce: null pointer");
  NGNN_REQUIRE(ld >= C && (dlogits == nullptr || ld_d >= C), NGNN_E_INVALID, "ce: leading dimension < C");
  k_ce_rows<<<(unsigned)ceil_div(bs * 32, 256), 256, 0, as_stream(stream)>>>(logits, ld, target, y_true, row_ids, bs, C,
                                                                               grad_scale, row_scratch, dlogits, ld_d);
  NGNN_LAUNCH_CHECK();
  k_ce_reduce<<<1, 1024, 0, as_stream(stream)>>>(row_scratch, bs, stats);
  NGNN_LAUNCH_CHECK();
  return NGNN_OK;
}

int32_t ngnn_ce_fwd_bwd(const float* logits, int64_t ld, const int64_t* target, const int64_t* y_true, int64_t bs,
                        int64_t C, float grad_scale, float* stats, float* dlogits, int64_t ld_d, float* row_scratch,
                        ngnn_stream_t stream) {
  return ngnn_ce_fwd_bwd_gather(logits, ld, target, y_true, nullptr, bs, C, grad_scale, stats, dlogits, ld_d, row_scratch, stream);
}

int32_t ngnn_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                       float beta1, float beta2, float eps, float weight_decay, float grad_scale, int64_t* step_dev,
                       int32_t advance_step, ngnn_stream_t stream) {
  NGNN_REQUIRE(n >= 0, NGNN_E_INVALID, "adam: negative size");
  NGNN_REQUIRE(step_dev, NGNN_E_INVALID, "adam: step_dev is null");
  cudaStream_t st = as_stream(stream);
  if (n > 0) {
    NGNN_REQUIRE(param && grad && exp_avg && exp_avg_sq, NGNN_E_INVALID, "adam: null pointer");
    k_adam<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                       weight_decay, grad_scale, step_dev);
    NGNN_LAUNCH_CHECK();
  }
  if (advance_step) {
    k_inc_step<<<1, 1, 0, st>>>(step_dev);
    NGNN_LAUNCH_CHECK();
  }
  return NGNN_OK;
}

}  // extern "C"
