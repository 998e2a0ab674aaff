// abi.cu — library-level entry points of libngnn_b200 (version, errors, device check).
#include "common.cuh"
#include <string.h>
#include <atomic>

namespace ngnn {

static thread_local char g_last_error[512] = "";

int32_t set_error(int32_t code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
  return code;
}

// ngnn_set_tuning(10, 0|1): programmatic dependent launch along the step's kernel chain.  OFF: measured on the products step
// (profiles/r02_notes.md) it is SLOWER, 0.494 ms against 0.413 ms per step — the early-scheduled CTAs of the next kernel
// (200 KB of shared memory, 57 k registers for a GEMM) sit on the SMs while they wait and keep the auxiliary stream's
// weight-gradient kernels and the sampler's kernels from being scheduled beside the running kernel.
int g_use_pdl = 0;

static std::atomic<uint64_t> g_launches{0};
void count_launches(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

}  // namespace ngnn

extern "C" {

uint64_t ngnn_launch_count(void) { return ngnn::g_launches.load(std::memory_order_relaxed); }

int32_t ngnn_version(void) { return 10000 * 0 + 100 * 1 + 0; }

int32_t ngnn_last_error(char* buf, size_t cap) {
  if (!buf || cap == 0) return NGNN_E_INVALID;
  strncpy(buf, ngnn::g_last_error, cap - 1);
  buf[cap - 1] = '\0';
  return NGNN_OK;
}

int32_t ngnn_device_supported(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return major == 10 ? 1 : 0;
}

}  // extern "C"
