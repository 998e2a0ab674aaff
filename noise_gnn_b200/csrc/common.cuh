// common.cuh — shared helpers for libngnn_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/ngnn_b200.h"

namespace ngnn {

// thread-local error message (read through ngnn_last_error)
int32_t set_error(int32_t code, const char* fmt, ...);

#define NGNN_REQUIRE(cond, code, ...)                                        \
  do {                                                                       \
    if (!(cond)) return ::ngnn::set_error((code), __VA_ARGS__);              \
  } while (0)

#define NGNN_CUDA(call)                                                      \
  do {                                                                       \
    cudaError_t _e = (call);                                                 \
    if (_e != cudaSuccess)                                                   \
      return ::ngnn::set_error(NGNN_E_CUDA, "%s failed: %s (%s:%d)", #call,  \
                               cudaGetErrorString(_e), __FILE__, __LINE__);  \
  } while (0)

// every kernel launch in the library is followed by NGNN_LAUNCH_CHECK(), which also counts it
// (ngnn_launch_count); library primitives (cub) add their launches with count_launches().
void count_launches(int n);

#define NGNN_LAUNCH_CHECK()                                                  \
  do {                                                                       \
    ::ngnn::count_launches(1);                                               \
    cudaError_t _e = cudaPeekAtLastError();                                  \
    if (_e != cudaSuccess)                                                   \
      return ::ngnn::set_error(NGNN_E_CUDA, "kernel launch failed: %s (%s:%d)", \
                               cudaGetErrorString(_e), __FILE__, __LINE__);  \
  } while (0)

static inline cudaStream_t as_stream(ngnn_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// Programmatic dependent launch.  The step is a chain of ~15 dependent kernels and each boundary costs a launch latency of a
// few microseconds on top of the predecessor's tail.  A kernel launched with launch_chain() may be scheduled as soon as
// every CTA of its predecessor in the stream has started (pdl_trigger at the top of each kernel) and then blocks in
// pdl_wait() until the predecessor has completed and flushed — its CTAs are already resident when that happens.
// Both instructions are no-ops for a kernel launched the ordinary way.  The attribute is OFF by default (measured slower on the
// full step, see abi.cu); ngnn_set_tuning(10, 1) turns it on.
#ifdef NGNN_PDL_EARLY
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#else
// Without the early trigger a dependent launch is released when every CTA of its predecessor has exited (the implicit
// trigger): no early residency, only the completion / launch hand-off is shortened.
__device__ __forceinline__ void pdl_trigger() {}
#endif
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
extern int g_use_pdl;
extern int g_draw_group;        // sampler.cu: fewest lanes per frontier node in k_hop_draw (ngnn_set_tuning(16, g))
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = g_use_pdl ? 1 : 0;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline bool is_aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

constexpr int kNumSMs = 148;   // B200: 2 dies x 74 SMs

// An extent (row / edge count) that may live in device memory.  A sampled block's sizes are only known on the device
// (the sampler writes them to `counts`); kernels launched with worst-case grids read them there, so a whole step is a
// FIXED launch sequence with no host round trip and can be captured in a CUDA graph.  `cap` is the host-side bound: the
// exact value when dev == nullptr, the capacity of the buffers otherwise (the device value is clamped to it).
struct Ext {
  const int32_t* dev;
  int64_t cap;
};
static inline Ext ext_host(int64_t n) { return Ext{nullptr, n}; }
static inline Ext ext_dev(const int32_t* p, int64_t cap) { return Ext{p, cap}; }
__device__ __forceinline__ int64_t ext_get(const Ext& e) {
  if (e.dev == nullptr) return e.cap;
  const int64_t v = (int64_t)__ldg(e.dev);
  return v < e.cap ? (v < 0 ? 0 : v) : e.cap;
}

// Per-step control words in device memory (ngnn_step_ctl_t): what changes from one replay of a captured step to the
// next.  Written by ngnn_step_ctl_set (a one-thread kernel whose arguments travel by value, so no host buffer has to
// outlive the launch) and read by the sampler (RNG key) and the K-GEMM epilogues (dropout stream offset).
struct StepCtl {
  uint32_t epoch, batch_idx;
  uint32_t drop_off_lo, drop_off_hi;
  uint32_t loss_scale_bits;      // float: weight of this batch's loss gradient (data parallel: bs_r * R / sum bs; 0 = padded batch)
  uint32_t ticket;               // arrival counter of the loss kernel's CTAs (the last one reduces); zeroed by ngnn_step_ctl_set
  uint32_t reserved[2];
};

// Library-internal entry points of the dense kernels (gemm.cu) for the fused step (step.cu): the split weight planes
// are prepared once per step, off the critical path, instead of inside every GEMM call.
//   mode 0: forward pack [W_l | W_r] (K-major) -> ws of ngnn_sage_gemm_workspace_bytes(F, O)
//   mode 2: the same without padding between the halves (the two activations are the halves of one [n, 2F] matrix)
//   mode 1: data-gradient pack [W_l^T ; W_r^T]  -> ws of ngnn_sage_dgrad_workspace_bytes(F, O)
// Returns NGNN_E_UNSUPPORTED when the shape takes the SIMT kernels (which read the weights directly).
// K-AGG forward from the resident feature table with a hot-row split (agg.cu): table rows < hot_rows are gathered with
// L2 evict_last priority, the rest with evict_first (hot_rows < 0: every row evict_last).
int32_t agg_fwd_table_impl(const int32_t* rowptr, const int32_t* col_table, const float* table, int64_t ld_table, Ext n_dst,
                           int64_t F, float* mean, int64_t ld_mean, const int32_t* root_table, float* root, int64_t ld_root,
                           int64_t hot_rows, cudaStream_t st, unsigned long long* clock = nullptr);
int32_t agg_fwd_impl(const int32_t* rowptr, const int32_t* col, const float* x, int64_t ld_x, Ext n_dst, int64_t F, float* mean,
                     int64_t ld_mean, cudaStream_t st);
int32_t agg_bwd_impl(const int32_t* colptr_t, const int32_t* row_t, const float* dmean_scaled, int64_t ld_dmean, Ext n_src,
                     int64_t F, const float* dx_root, int64_t ld_root, Ext n_root, const float* act_ref, int64_t ld_act,
                     float act_scale, float* dx, int64_t ld_dx, cudaStream_t st);
int32_t prep_weights_impl(int32_t mode, const float* w_l, const float* w_r, int64_t F, int64_t O, void* ws, size_t ws_bytes,
                          cudaStream_t st);
// Batched form: prep_batch_add collects jobs (same arguments / return codes), prep_batch_launch runs them in one launch.
void prep_batch_begin();
int32_t prep_batch_add(int32_t mode, const float* w_l, const float* w_r, int64_t F, int64_t O, void* ws, size_t ws_bytes);
int32_t prep_batch_launch(cudaStream_t st);
int32_t gemm_fwd_impl(const float* a_l, int64_t ld_al, const float* a_r, int64_t ld_ar, const float* w_l, const float* w_r,
                      const float* bias, int64_t n, int64_t F, int64_t O, int32_t act, float drop_p, uint64_t seed,
                      uint64_t offset, float* out, int64_t ld_out, int32_t* path, void* ws, size_t ws_bytes, cudaStream_t st,
                      bool prepped, const int32_t* n_dev, const struct StepCtl* ctl, uint32_t ctl_layer, bool concat_k);
int32_t dgrad_impl(const float* dy, int64_t ld_dy, const float* w_l, const float* w_r, const int32_t* rowptr, int64_t n,
                   int64_t F, int64_t O, float* dmean_scaled, int64_t ld_dmean, float* dx_root, int64_t ld_root, void* ws,
                   size_t ws_bytes, cudaStream_t st, bool prepped, const int32_t* n_dev);
int32_t ce_impl(const float* logits, int64_t ld, const int64_t* target, const int64_t* y_true, const int32_t* row_ids,
                int64_t bs, int64_t C, float grad_scale, float* stats, float* dlogits, int64_t ld_d, float* row_scratch,
                const struct StepCtl* ctl, cudaStream_t st);
int32_t wgrad_impl(const float* dy, int64_t ld_dy, const float* a_l, int64_t ld_al, const float* a_r, int64_t ld_ar, int64_t n,
                   const int32_t* n_dev, int64_t F, int64_t O, float* dw_l, float* dw_r, float* db, int32_t accumulate,
                   void* ws, size_t ws_bytes, cudaStream_t st);

// ---------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11).  Counter-based: the same (key, counter)
// gives the same 4 words on host and device.  oracle/sampler_oracle.c carries an
// independent restatement used to pin the sampler bit-exactly.
// ---------------------------------------------------------------------------
struct Philox4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = mulhi32(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = mulhi32(M1, c2), lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  return Philox4{c0, c1, c2, c3};
}

// 128-bit read-only global load that does not pollute L1 (streamed gathers)
__device__ __forceinline__ float4 ldg_nc_f4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// Same with an L2 evict_last priority: rows of the resident feature table.  Hub nodes of a power-law graph are gathered
// again and again (within a block and from step to step); at default priority the step's streaming activations
// (~0.5 GB per step through a 126 MB L2) push them out between uses.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ float4 ldg_nc_f4_hint(const float4* p, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
  return v;
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

}  // namespace ngnn
