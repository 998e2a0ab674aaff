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

// Single CTA, 32 warps; warp w handles rows w, w+32, ...; partial sums are combined in a fixed
// order, so the loss is bitwise reproducible.
__global__ void __launch_bounds__(1024) k_ce_fwd_bwd(const float* __restrict__ logits, int64_t ld,
                                                     const int64_t* __restrict__ target,
                                                     const int64_t* __restrict__ y_true, int64_t bs, int64_t C,
                                                     float grad_scale, float* __restrict__ stats,
                                                     float* __restrict__ dlogits, int64_t ld_d) {
  __shared__ float s_loss[32];
  __shared__ float s_corr[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float inv_bs = 1.0f / (float)bs;
  float loss_acc = 0.f, corr_acc = 0.f;
  for (int64_t i = warp; i < bs; i += 32) {
    const float* row = logits + i * ld;
    float mx = -INFINITY;
    int64_t arg = 0;
    for (int64_t c = lane; c < C; c += 32) {
      const float v = row[c];
      if (v > mx) { mx = v; arg = c; }
    }
    // argmax with first-index tie-break (torch.argmax semantics on ties are first occurrence)
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
    const int64_t tgt = target[i];
    if (dlogits != nullptr) {
      const float g = grad_scale * inv_bs;
      for (int64_t c = lane; c < C; c += 32) {
        const float pr = expf(row[c] - lse);
        dlogits[i * ld_d + c] = (pr - (c == tgt ? 1.f : 0.f)) * g;
      }
    }
    if (lane == 0) {
      loss_acc += (lse - row[tgt]);
      if (y_true != nullptr && arg == y_true[i]) corr_acc += 1.f;
    }
  }
  if (lane == 0) { s_loss[warp] = loss_acc; s_corr[warp] = corr_acc; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float l = 0.f, c = 0.f;
    for (int w = 0; w < 32; ++w) { l += s_loss[w]; c += s_corr[w]; }
    stats[0] += l * inv_bs;
    stats[1] += c;
  }
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

int32_t ngnn_ce_fwd_bwd(const float* logits, int64_t ld, const int64_t* target, const int64_t* y_true, int64_t bs,
                        int64_t C, float grad_scale, float* stats, float* dlogits, int64_t ld_d,
                        ngnn_stream_t stream) {
  NGNN_REQUIRE(bs >= 0 && C >= 0, NGNN_E_INVALID, "ce: negative size");
  if (bs == 0 || C == 0) return NGNN_OK;
  NGNN_REQUIRE(logits && target && stats, NGNN_E_INVALID, "ce: null pointer");
  NGNN_REQUIRE(ld >= C && (dlogits == nullptr || ld_d >= C), NGNN_E_INVALID, "ce: leading dimension < C");
  k_ce_fwd_bwd<<<1, 1024, 0, as_stream(stream)>>>(logits, ld, target, y_true, bs, C, grad_scale, stats, dlogits, ld_d);
  NGNN_LAUNCH_CHECK();
  return NGNN_OK;
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
