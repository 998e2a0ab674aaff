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

__device__ __forceinline__ void ct_row_stats(const float* row, int64_t C, int lane, int64_t tgt, float& loss, float& lse, int64_t& arg);

// One warp per row: log-sum-exp, gradient row, per-row loss / correct flag into `rows` ([2*bs]).
__global__ void __launch_bounds__(256) k_ce_rows(const float* __restrict__ logits, int64_t ld,
                                                 const int64_t* __restrict__ target, const int64_t* __restrict__ y_true,
                                                 const int32_t* __restrict__ row_ids, int64_t bs, int64_t C,
                                                 float grad_scale, float* __restrict__ rows,
                                                 float* __restrict__ dlogits, int64_t ld_d, const StepCtl* ctl) {
  if (ctl != nullptr) grad_scale *= __uint_as_float(ctl->loss_scale_bits);     // data-parallel weight of this rank's batch
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

// fixed-order tree sum of the per-row values by one CTA (any block size that is a power of two <= 1024)
__device__ __forceinline__ void ce_reduce_cta(const float* rows, int64_t bs, float* stats, float* s_l, float* s_c) {
  const int T = blockDim.x;
  float l = 0.f, c = 0.f;
  for (int64_t i = threadIdx.x; i < bs; i += T) { l += rows[i]; c += rows[bs + i]; }
  s_l[threadIdx.x] = l; s_c[threadIdx.x] = c;
  __syncthreads();
  for (int o = T >> 1; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) { s_l[threadIdx.x] += s_l[threadIdx.x + o]; s_c[threadIdx.x] += s_c[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { stats[0] += s_l[0] / (float)bs; stats[1] += s_c[0]; }
}

// k_ce_rows + the reduction in ONE launch (replayed step): the CTA that finishes last (a ticket in the step's control words,
// reset by ngnn_step_ctl_set before every step) sums the rows in the same fixed order as k_ce_reduce with 256 threads.
__global__ void __launch_bounds__(256) k_ce_rows_reduce(const float* __restrict__ logits, int64_t ld,
                                                        const int64_t* __restrict__ target, const int64_t* __restrict__ y_true,
                                                        const int32_t* __restrict__ row_ids, int64_t bs, int64_t C,
                                                        float grad_scale, float* rows, float* __restrict__ dlogits,
                                                        int64_t ld_d, StepCtl* ctl, float* stats) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s_l[256], s_c[256];
  __shared__ int s_last;
  const float w = __uint_as_float(ctl->loss_scale_bits);
  grad_scale *= w;
  const int lane = threadIdx.x & 31;
  const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (i < bs) {
    const float* row = logits + i * ld;
    const int64_t li = row_ids != nullptr ? (int64_t)row_ids[i] : i;
    const int64_t tgt = target[li];
    float loss, lse;
    int64_t arg;
    ct_row_stats(row, C, lane, tgt, loss, lse, arg);
    if (dlogits != nullptr) {
      const float g = grad_scale / (float)bs;
      for (int64_t c = lane; c < C; c += 32) {
        const float pr = expf(row[c] - lse);
        dlogits[i * ld_d + c] = (pr - (c == tgt ? 1.f : 0.f)) * g;
      }
    }
    if (lane == 0) {
      rows[i] = loss;
      rows[bs + i] = (y_true != nullptr && arg == y_true[li]) ? 1.f : 0.f;
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&ctl->ticket, 1u) == gridDim.x - 1 ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (w != 0.f) ce_reduce_cta(rows, bs, stats, s_l, s_c);        // a padded (masked-out) batch is not logged
}

// Single CTA: fixed-order tree sum of the per-row values => bitwise reproducible loss.
__global__ void __launch_bounds__(1024) k_ce_reduce(const float* __restrict__ rows, int64_t bs, float* __restrict__ stats,
                                                    const StepCtl* ctl) {
  if (ctl != nullptr && __uint_as_float(ctl->loss_scale_bits) == 0.f) return;   // a padded (masked-out) batch is not logged
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

// ---------------------------------------------------------------------------------------------------
// Co-teaching loss (reference src/utils/losses.py:10-49, CTLoss; Han et al. 2018): each of two peer networks is
// trained on the (1 - forget_rate) fraction of the batch its PEER finds easiest.  The reference sorts the per-sample
// losses on the host (np.argsort(loss.cpu()) twice per step: two device->host round trips); here the ranks are
// counted on the device (bs <= a few thousand: bs^2 comparisons), the step stays free of host synchronisation.
//   k_ct_rows : warp per row, both models: per-sample CE, lse, argmax == label
//   k_ct_rank : rank_m[i] = #{ j : (loss_m[j], j) < (loss_m[i], i) }  (ties by index: a stable argsort), order_m[rank] = i
//   k_ct_grad : model 1 learns from rows with rank_2 < R, model 2 from rows with rank_1 < R:
//               dlogits_m[i] = keep_peer ? (softmax_m - onehot) / R : 0 ; selected losses / pure flags per row
//   k_ct_reduce: single CTA, fixed-order sums into stats[6]
// scratch layout (floats, n = bs): loss1 | loss2 | lse1 | lse2 | corr1 | corr2 | sel1 | sel2 | pure1 | pure2 | rank1 | rank2 (as int)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ct_row_stats(const float* row, int64_t C, int lane, int64_t tgt, float& loss, float& lse, int64_t& arg) {
  float mx = -INFINITY;
  arg = 0;
  for (int64_t c = lane; c < C; c += 32) {
    const float v = row[c];
    if (v > mx) { mx = v; arg = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float omx = __shfl_xor_sync(0xffffffffu, mx, o);
    const int64_t oarg = __shfl_xor_sync(0xffffffffu, arg, o);
    if (omx > mx || (omx == mx && oarg < arg)) { mx = omx; arg = oarg; }
  }
  float se = 0.f;
  for (int64_t c = lane; c < C; c += 32) se += expf(row[c] - mx);
  se = warp_sum(se);
  lse = mx + logf(se);
  loss = lse - row[tgt];
}

__global__ void __launch_bounds__(256) k_ct_rows(const float* __restrict__ l1, int64_t ld1, const float* __restrict__ l2, int64_t ld2,
                                                 const int64_t* __restrict__ target, const int64_t* __restrict__ y_true,
                                                 const int32_t* __restrict__ row_ids, int64_t bs, int64_t C,
                                                 float* __restrict__ sc) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (i >= bs) return;
  const int64_t li = row_ids != nullptr ? (int64_t)row_ids[i] : i;
  const int64_t tgt = target[li];
  float loss, lse;
  int64_t arg;
  ct_row_stats(l1 + i * ld1, C, lane, tgt, loss, lse, arg);
  if (lane == 0) { sc[i] = loss; sc[2 * bs + i] = lse; sc[4 * bs + i] = (y_true != nullptr && arg == y_true[li]) ? 1.f : 0.f; }
  ct_row_stats(l2 + i * ld2, C, lane, tgt, loss, lse, arg);
  if (lane == 0) { sc[bs + i] = loss; sc[3 * bs + i] = lse; sc[5 * bs + i] = (y_true != nullptr && arg == y_true[li]) ? 1.f : 0.f; }
}

__global__ void __launch_bounds__(256) k_ct_rank(float* __restrict__ sc, int64_t bs, int32_t* __restrict__ order1,
                                                 int32_t* __restrict__ order2) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= 2 * bs) return;
  const int m = t >= bs ? 1 : 0;
  const int64_t i = t - m * bs;
  const float* loss = sc + m * bs;
  const float mine = loss[i];
  int32_t r = 0;
  for (int64_t j = 0; j < bs; ++j) {
    const float o = loss[j];
    r += (o < mine || (o == mine && j < i)) ? 1 : 0;
  }
  reinterpret_cast<int32_t*>(sc + (10 + m) * bs)[i] = r;
  int32_t* order = m ? order2 : order1;
  if (order != nullptr) order[r] = (int32_t)i;
}

__global__ void __launch_bounds__(256) k_ct_grad(const float* __restrict__ l1, int64_t ld1, const float* __restrict__ l2, int64_t ld2,
                                                 const int64_t* __restrict__ target, const int32_t* __restrict__ row_ids,
                                                 const uint8_t* __restrict__ clean, int64_t bs, int64_t C, int64_t R,
                                                 float* __restrict__ sc, float* __restrict__ d1, int64_t ldd1,
                                                 float* __restrict__ d2, int64_t ldd2) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (i >= bs) return;
  const int64_t li = row_ids != nullptr ? (int64_t)row_ids[i] : i;
  const int64_t tgt = target[li];
  const bool keep1 = reinterpret_cast<const int32_t*>(sc + 10 * bs)[i] < R;   // among model 1's R smallest losses
  const bool keep2 = reinterpret_cast<const int32_t*>(sc + 11 * bs)[i] < R;
  const float g = R > 0 ? 1.0f / (float)R : 0.f;
  if (d1 != nullptr) {          // model 1 is updated on model 2's selection
    const float lse = sc[2 * bs + i];
    for (int64_t c = lane; c < C; c += 32)
      d1[i * ldd1 + c] = keep2 ? (expf(l1[i * ld1 + c] - lse) - (c == tgt ? 1.f : 0.f)) * g : 0.f;
  }
  if (d2 != nullptr) {
    const float lse = sc[3 * bs + i];
    for (int64_t c = lane; c < C; c += 32)
      d2[i * ldd2 + c] = keep1 ? (expf(l2[i * ld2 + c] - lse) - (c == tgt ? 1.f : 0.f)) * g : 0.f;
  }
  if (lane == 0) {
    sc[6 * bs + i] = keep2 ? sc[i] : 0.f;
    sc[7 * bs + i] = keep1 ? sc[bs + i] : 0.f;
    const float pure = (clean != nullptr && clean[li]) ? 1.f : 0.f;
    sc[8 * bs + i] = keep1 ? pure : 0.f;
    sc[9 * bs + i] = keep2 ? pure : 0.f;
  }
}

// stats[0..5] += loss_1, loss_2 (means over the R selected rows), correct_1, correct_2, pure_ratio_1, pure_ratio_2
__global__ void __launch_bounds__(1024) k_ct_reduce(const float* __restrict__ sc, int64_t bs, int64_t R, float* __restrict__ stats) {
  __shared__ float sm[6][1024];
  const int src[6] = {6, 7, 4, 5, 8, 9};
  float a[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int64_t i = threadIdx.x; i < bs; i += 1024)
#pragma unroll
    for (int k = 0; k < 6; ++k) a[k] += sc[src[k] * bs + i];
#pragma unroll
  for (int k = 0; k < 6; ++k) sm[k][threadIdx.x] = a[k];
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o)
#pragma unroll
      for (int k = 0; k < 6; ++k) sm[k][threadIdx.x] += sm[k][threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float inv = R > 0 ? 1.0f / (float)R : 0.f;
    stats[0] += sm[0][0] * inv; stats[1] += sm[1][0] * inv;
    stats[2] += sm[2][0]; stats[3] += sm[3][0];
    stats[4] += sm[4][0] * inv; stats[5] += sm[5][0] * inv;
  }
}

// ticket != nullptr: the CTA that finishes last advances *step_dev (every CTA has read it by then) and clears the ticket —
// the counter moves without a second launch.
__global__ void k_adam(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ m,
                       float* __restrict__ v, int64_t n, float lr, float beta1, float beta2, float eps,
                       float weight_decay, float grad_scale, int64_t* step_dev, unsigned int* ticket) {
  pdl_trigger();
  pdl_wait();
  // the two bias corrections (double-precision pow) once per CTA, not once per parameter
  __shared__ float s_bc[2];
  if (threadIdx.x == 0) {
    const double t = (double)(*step_dev + 1);
    s_bc[0] = (float)(1.0 - pow((double)beta1, t));
    s_bc[1] = (float)sqrt(1.0 - pow((double)beta2, t));
  }
  __syncthreads();
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) {
    const float bc1 = s_bc[0], bc2_sqrt = s_bc[1];
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
  if (ticket != nullptr) {
    __syncthreads();                               // thread 0's read of *step_dev above is behind this CTA
    if (threadIdx.x == 0 && atomicAdd(ticket, 1u) == gridDim.x - 1) { *step_dev += 1; *ticket = 0u; }
  }
}

__global__ void k_inc_step(int64_t* step_dev) { *step_dev += 1; }

}  // namespace ngnn

using namespace ngnn;

extern "C" {

}  // extern "C"

namespace ngnn {
int32_t ce_impl(const float* logits, int64_t ld, const int64_t* target, const int64_t* y_true, const int32_t* row_ids,
                int64_t bs, int64_t C, float grad_scale, float* stats, float* dlogits, int64_t ld_d, float* row_scratch,
                const StepCtl* ctl, cudaStream_t st) {
  NGNN_REQUIRE(bs >= 0 && C >= 0, NGNN_E_INVALID, "ce: negative size");
  if (bs == 0 || C == 0) return NGNN_OK;
  NGNN_REQUIRE(logits && target && stats && row_scratch, NGNN_E_INVALID, "ce: null pointer");
  NGNN_REQUIRE(ld >= C && (dlogits == nullptr || ld_d >= C), NGNN_E_INVALID, "ce: leading dimension < C");
  if (ctl != nullptr) {      // replayed step: one launch, the last CTA reduces (ticket in the control words)
    launch_chain(k_ce_rows_reduce, dim3((unsigned)ceil_div(bs * 32, 256)), dim3(256), 0, st, logits, ld, target, y_true, row_ids, bs, C,
                 grad_scale, row_scratch, dlogits, ld_d, const_cast<StepCtl*>(ctl), stats);
    NGNN_LAUNCH_CHECK();
    return NGNN_OK;
  }
  k_ce_rows<<<(unsigned)ceil_div(bs * 32, 256), 256, 0, st>>>(logits, ld, target, y_true, row_ids, bs, C, grad_scale, row_scratch,
                                                                dlogits, ld_d, ctl);
  NGNN_LAUNCH_CHECK();
  k_ce_reduce<<<1, 1024, 0, st>>>(row_scratch, bs, stats, ctl);
  NGNN_LAUNCH_CHECK();
  return NGNN_OK;
}
}  // namespace ngnn

extern "C" {

int32_t ngnn_ce_fwd_bwd_gather(const float* logits, int64_t ld, const int64_t* target, const int64_t* y_true,
                               const int32_t* row_ids, int64_t bs, int64_t C, float grad_scale, float* stats, float* dlogits,
                               int64_t ld_d, float* row_scratch, ngnn_stream_t stream) {
  return ce_impl(logits, ld, target, y_true, row_ids, bs, C, grad_scale, stats, dlogits, ld_d, row_scratch, nullptr,
                 as_stream(stream));
}

int32_t ngnn_ce_fwd_bwd(const float* logits, int64_t ld, const int64_t* target, const int64_t* y_true, int64_t bs,
                        int64_t C, float grad_scale, float* stats, float* dlogits, int64_t ld_d, float* row_scratch,
                        ngnn_stream_t stream) {
  return ngnn_ce_fwd_bwd_gather(logits, ld, target, y_true, nullptr, bs, C, grad_scale, stats, dlogits, ld_d, row_scratch, stream);
}

int32_t ngnn_ct_loss(const float* logits1, int64_t ld1, const float* logits2, int64_t ld2, const int64_t* target,
                     const int64_t* y_true, const int32_t* row_ids, const uint8_t* clean_mask, int64_t bs, int64_t C,
                     int64_t num_remember, float* stats, float* dlogits1, int64_t ldd1, float* dlogits2, int64_t ldd2,
                     int32_t* order1, int32_t* order2, float* scratch, ngnn_stream_t stream) {
  NGNN_REQUIRE(bs >= 0 && C >= 0 && num_remember >= 0 && num_remember <= bs, NGNN_E_INVALID, "ct_loss: bad sizes");
  if (bs == 0 || C == 0) return NGNN_OK;
  NGNN_REQUIRE(logits1 && logits2 && target && stats && scratch, NGNN_E_INVALID, "ct_loss: null pointer");
  NGNN_REQUIRE(ld1 >= C && ld2 >= C && (dlogits1 == nullptr || ldd1 >= C) && (dlogits2 == nullptr || ldd2 >= C), NGNN_E_INVALID,
               "ct_loss: leading dimension < C");
  cudaStream_t st = as_stream(stream);
  k_ct_rows<<<(unsigned)ceil_div(bs * 32, 256), 256, 0, st>>>(logits1, ld1, logits2, ld2, target, y_true, row_ids, bs, C, scratch);
  NGNN_LAUNCH_CHECK();
  k_ct_rank<<<(unsigned)ceil_div(2 * bs, 256), 256, 0, st>>>(scratch, bs, order1, order2);
  NGNN_LAUNCH_CHECK();
  k_ct_grad<<<(unsigned)ceil_div(bs * 32, 256), 256, 0, st>>>(logits1, ld1, logits2, ld2, target, row_ids, clean_mask, bs, C,
                                                              num_remember, scratch, dlogits1, ldd1, dlogits2, ldd2);
  NGNN_LAUNCH_CHECK();
  k_ct_reduce<<<1, 1024, 0, st>>>(scratch, bs, num_remember, stats);
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
    // advance_step == 2: step_dev[1] is a zero-initialised ticket word owned by this call sequence (no second launch)
    unsigned int* ticket = advance_step == 2 ? reinterpret_cast<unsigned int*>(step_dev + 1) : nullptr;
    launch_chain(k_adam, dim3((unsigned)ceil_div(n, 256)), dim3(256), 0, st, param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                 weight_decay, grad_scale, step_dev, ticket);
    NGNN_LAUNCH_CHECK();
    if (ticket != nullptr) return NGNN_OK;
  }
  if (advance_step) {
    k_inc_step<<<1, 1, 0, st>>>(step_dev);
    NGNN_LAUNCH_CHECK();
  }
  return NGNN_OK;
}

}  // extern "C"
