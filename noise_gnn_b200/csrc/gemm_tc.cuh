// gemm_tc.cuh — tcgen05 / TMA tensor-core path of K-GEMM and K-DGRAD (sm_100a).
//
//   D[m, n] = sum_k A1[m,k] * B1[n,k] + sum_k A2[m,k] * B2[n,k]        (both operands K-major)
//
// fp32-grade accuracy on the TF32 tensor pipe by the 3xTF32 split (SURVEY §7.3 hard part 1):
//   x = hi + lo, hi = round-to-nearest of x to 10 mantissa bits, lo = x - hi (exact in fp32);
//   D += A_hi*B_hi + A_lo*B_hi + A_hi*B_lo   (the dropped lo*lo term is ~2^-22 relative).
// Weights (B) are split once per step (or per call) by a tiny prep kernel into a packed K-major workspace
// [rows, Kpack] (hi and lo planes, K zero-padded to whole 32-float blocks); activations (A) are split by the
// converter warps right after TMA lands them — into TENSOR MEMORY in the default TS form, into shared memory in the
// SS form (kept for tiles wider than 128 columns and for A/B comparison, ngnn_set_tuning(6, 0)).
//
// Persistent, warp-specialised, one CTA per SM, 128 x BN output tiles (BN = N rounded up to 16, <= 128 for TS), 14 warps:
//   warp 0      TMA producer: per K-block (32 floats = one 128-byte swizzle atom) loads the raw A tile [128 x 32] and
//               the B_hi / B_lo tiles [BN x 32] (SWIZZLE_128B) into a stage of the ring (4 stages of 48 KB at BN = 128)
//   warp 1      allocates TMEM, issues tcgen05.mma.kind::tf32 (12 per K-block: 4 k-steps x 3 terms) with elect.sync and
//               warp-uniform operands, waits on ONE mbarrier per K-block; tcgen05.commit frees the stage / signals the epilogue
//   warps 2-5   converters: thread = tile row, raw A row -> hi / lo -> 64 TMEM columns of its lane (tcgen05.st)
//   warps 6-13  epilogue (two per TMEM lane quarter): tcgen05.ld the fp32 accumulator, + bias, row scale, ReLU, Philox
//               dropout (16 random bits per element), transposed through a shared-memory patch into full-line stores;
//               two TMEM accumulator buffers let it run under the next tile's main loop.
// Why TS: with A and B both in shared memory the kernel was bound by the shared-memory port (210 KB per K-block at
// 128 B/cycle = 1,640 cycles against 830 cycles of tensor work); see DESIGN.md §3.2 and profiles/r01_gemm_trace_notes.txt.
// Partial tiles rely on TMA's zero fill for out-of-bounds rows / columns; stores are bounds-checked.
#pragma once
#include "common.cuh"
#include "gemm_simt.cuh"   // dropout_keep8 / dropout_threshold (the mask definition is shared)
#include <cuda.h>

namespace ngnn {

constexpr int TC_BM = 128;            // rows per CTA tile (UMMA M)
constexpr int TC_BK = 32;             // floats per K-block: 128 bytes = one SWIZZLE_128B atom
constexpr int TC_THREADS = 192;
constexpr int TC_MAX_STAGES = 4;
constexpr uint32_t TC_SMEM_LIMIT = 227u * 1024u;

// ----------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// Bounded spin: a protocol bug traps (CUDA error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  for (uint32_t it = 0; it < (1u << 26); ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (done) return;
  }
  __trap();
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], tf32 inputs, fp32 accumulate; issued by ONE thread
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same with the A operand read from tensor memory (lane = row, one 32-bit column per tf32 k element): only B crosses the
// shared-memory port, which the SS form saturates (A 4 KB + B 4 KB per 64-cycle 128x128x8 tf32 MMA = 128 B/cycle).
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Warp-convergent forms: every lane executes the block, `elect.sync` picks the issuing lane.  Issued from inside an
// `if (lane == 0)` region the operands are per-thread values and ptxas wraps every UTCHMMA in an ELECT / R2UR / BRA
// loop (~65 cycles per MMA: as long as the MMA itself, so the tensor pipe never has a backlog and every wait of the
// issuing thread is a pipe bubble); with warp-uniform operands they go straight to uniform registers.
__device__ __forceinline__ void umma_tf32_elect(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_ts_elect(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(smem_u32(bar)) : "memory");
}
// 16 consecutive 32-bit columns of this thread's TMEM lane <- registers
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// arrives on the mbarrier once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor, K-major, SWIZZLE_128B: rows of 128 bytes, 8-row groups 1024 B apart
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout [61,64)).
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)0 << 16;                       // LBO unused: one swizzle atom along K
  d |= (uint64_t)((1024u >> 4) & 0x3FFF) << 32; // SBO: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                       // LayoutType::SWIZZLE_128B
  return d;
}
// Instruction descriptor, kind::tf32: D fp32, A/B tf32, both K-major, shape M x N (x8).
__host__ __device__ __forceinline__ uint32_t umma_idesc_tf32(uint32_t M, uint32_t N, uint32_t a_mn_major = 0, uint32_t b_mn_major = 0) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (a_mn_major << 15) | (b_mn_major << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// hi = x rounded to nearest at mantissa bit 13 (10 explicit bits kept), lo = x - hi (exact)
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  const uint32_t u = __float_as_uint(x);
  hi = __uint_as_float((u + 0x1000u) & 0xFFFFE000u);
  lo = x - hi;
}

// Coalesced tile store for the epilogue warps.  After tcgen05.ld each lane holds consecutive columns of ITS row, so
// a direct store makes every warp-level instruction touch 32 different lines with 16 bytes each.  Instead the warp
// transposes through a private 32 x 32 staging patch in shared memory (row stride 36 floats: conflict-free for
// 128-bit accesses) and writes 4 rows x 128 contiguous bytes (full lines) per instruction.
// `r` = this lane's finished values for columns [nb, nb+32) of row (row0 + lane); columns >= n_cols are dropped.
constexpr int EPI_STRIDE = 36;
constexpr int EPI_PATCH = 32 * EPI_STRIDE;      // floats per warp
__device__ __forceinline__ void epilogue_store32(float* stage /*warp-private [32][36]*/, const float (&r)[32], float* out,
                                                 int64_t ld_out, int64_t row0, int64_t n_rows, int32_t nb, int32_t n_cols,
                                                 bool vec_ok, int lane) {
#pragma unroll
  for (int g = 0; g < 8; ++g)
    *reinterpret_cast<float4*>(stage + lane * EPI_STRIDE + 4 * g) = make_float4(r[4 * g], r[4 * g + 1], r[4 * g + 2], r[4 * g + 3]);
  __syncwarp();
  const int sub = lane >> 3, part = lane & 7;          // 4 rows per instruction, 8 lanes (128 B) per row
#pragma unroll
  for (int rr = 0; rr < 32; rr += 4) {
    const int lr = rr + sub;
    const float4 v = *reinterpret_cast<const float4*>(stage + lr * EPI_STRIDE + 4 * part);
    const int64_t m = row0 + lr;
    const int32_t n4 = nb + 4 * part;
    if (m < n_rows && n4 < n_cols) {
      float* dst = out + m * ld_out + n4;
      if (vec_ok && n4 + 3 < n_cols) {
        *reinterpret_cast<float4*>(dst) = v;
      } else {
        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) if (n4 + jj < n_cols) dst[jj] = e[jj];
      }
    }
  }
  __syncwarp();
}

// Same for a 16-column unit, through a [32][20] patch (2.5 KB per warp instead of 4.5 KB: eight epilogue warps fit
// beside four 48 KB stages); 8 rows x 64 contiguous bytes per store instruction.
constexpr int EPI16_STRIDE = 20;
constexpr int EPI16_PATCH = 32 * EPI16_STRIDE;  // floats per warp
__device__ __forceinline__ void epilogue_store16(float* stage /*warp-private [32][20]*/, const float (&r)[16], float* out,
                                                 int64_t ld_out, int64_t row0, int64_t n_rows, int32_t nb, int32_t n_cols,
                                                 bool vec_ok, int lane) {
  const int sub = lane >> 2, part = lane & 3;          // 8 rows per instruction, 4 lanes (64 B) per row
#pragma unroll
  for (int g = 0; g < 4; ++g)
    *reinterpret_cast<float4*>(stage + lane * EPI16_STRIDE + 4 * g) = make_float4(r[4 * g], r[4 * g + 1], r[4 * g + 2], r[4 * g + 3]);
  __syncwarp();
#pragma unroll
  for (int rr = 0; rr < 32; rr += 8) {
    const int lr = rr + sub;
    const float4 v = *reinterpret_cast<const float4*>(stage + lr * EPI16_STRIDE + 4 * part);
    const int64_t m = row0 + lr;
    const int32_t n4 = nb + 4 * part;
    if (m < n_rows && n4 < n_cols) {
      float* dst = out + m * ld_out + n4;
      if (vec_ok && n4 + 3 < n_cols) {
        *reinterpret_cast<float4*>(dst) = v;
      } else {
        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) if (n4 + jj < n_cols) dst[jj] = e[jj];
      }
    }
  }
  __syncwarp();
}

// ----------------------------------------------------------------------------- weight prep
// Packs up to two [R, C] fp32 weight matrices side by side along K into hi / lo planes [R_out, Kpack]:
//   transpose == 0: out[r, seg*Cpad + c] = W_seg[r, c]           (forward:  B = [W_l | W_r], rows = O, K = F)
//   transpose == 1: out[seg*Rpad + c, r] = W_seg[r, c]           (dgrad:    B = [W_l^T ; W_r^T], rows = F, K = O)
// K padding is zero-filled.
struct PrepParams {
  const float* w[2];
  int32_t R, C;          // source rows / cols
  int32_t transpose;
  int32_t seg_stride;    // forward: Cpad (k offset of segment 1); dgrad: row offset of segment 1
  int32_t Kpack;         // leading dimension of the packed planes
  int32_t rows_out;
  float *hi, *lo;
};
__device__ __forceinline__ void prep_weights_element(const PrepParams& p, int64_t idx) {
  const int64_t total = (int64_t)p.rows_out * p.Kpack;
  if (idx >= total) return;
  const int32_t ro = (int32_t)(idx / p.Kpack), ko = (int32_t)(idx - (int64_t)ro * p.Kpack);
  float v = 0.f;
  if (!p.transpose) {
    const int32_t seg = ko / p.seg_stride, c = ko - seg * p.seg_stride;
    if (seg < 2 && p.w[seg] != nullptr && c < p.C) v = p.w[seg][(int64_t)ro * p.C + c];
  } else {
    const int32_t seg = ro / p.seg_stride, c = ro - seg * p.seg_stride;
    if (seg < 2 && p.w[seg] != nullptr && c < p.C && ko < p.R) v = p.w[seg][(int64_t)ko * p.C + c];
  }
  float hi, lo;
  split_tf32(v, hi, lo);
  p.hi[idx] = hi;
  p.lo[idx] = lo;
}
__global__ void k_prep_weights(PrepParams p) { prep_weights_element(p, blockIdx.x * (int64_t)blockDim.x + threadIdx.x); }

// Every pack of a whole network in ONE launch (the fused step prepares the forward and data-gradient packs of all
// layers before its first kernel): blocks [first_block[j], first_block[j+1]) work on job j.
constexpr int PREP_MAX_JOBS = 32;
struct PrepBatch {
  PrepParams job[PREP_MAX_JOBS];
  int32_t first_block[PREP_MAX_JOBS + 1];
  int32_t n_jobs;
};
__global__ void k_prep_weights_batch(const __grid_constant__ PrepBatch b) {
  pdl_trigger();
  pdl_wait();
  int j = 0;
  while (j + 1 < b.n_jobs && (int32_t)blockIdx.x >= b.first_block[j + 1]) ++j;
  prep_weights_element(b.job[j], (int64_t)(blockIdx.x - b.first_block[j]) * blockDim.x + threadIdx.x);
}

// ----------------------------------------------------------------------------- main kernel
struct TcSegment {            // one N range of the packed B matrix and where its output goes
  float* out;
  int64_t ld_out;
  int32_t n_cols;             // valid output columns of this segment
  int32_t b_row0;             // first row of the segment in the packed B planes
  int32_t scale_rows;         // 1: scale row m by 1/max(rowptr[m+1]-rowptr[m],1)
};
struct TcGemmParams {
  int32_t M;                  // rows (capacity when M_dev != nullptr)
  const int32_t* M_dev;       // optional device-side row count (clamped to M): the tensor maps and the grid are sized for M
  const StepCtl* ctl;         // optional device-side step control: dropout offset = ctl->drop_off + ctl_layer
  uint32_t ctl_layer;
  int32_t BN;                 // UMMA N / columns per CTA tile
  int32_t kblocks1, kblocks2; // K-blocks of operand pair 1 / 2 covered by THIS launch
  int32_t kb_begin1, kb_begin2; // first K-block of each pair this launch covers (a long contraction is cut into several launches)
  int32_t ktail1, ktail2;     // 8-float k-steps that hold data in the LAST K-block of pair 1 / 2 of this launch (1..4): the rest is TMA zero fill
  int32_t acc_in, final;      // acc_in: add the output the previous launch left; final: apply bias / scale / activation / dropout
  int32_t b_koff2;            // k offset (floats) of pair 2 in the packed B planes
  int32_t stages;
  int32_t tiles_per_seg;      // N tiles per segment
  int32_t num_segs;
  TcSegment seg[2];
  const float* bias;
  const int32_t* rowptr;
  int32_t act;
  float drop_p;
  uint32_t seed_lo, seed_hi, off_lo, off_hi;
  long long* trace;           // debug: CTA (0,0) records clock64 per K-block: [kb][0..5] (see profiles/trace_gemm.py)
};

static long long* g_tc_trace = nullptr;   // ngnn_debug_set_trace

// Persistent, warp-specialised:  warp 0 = TMA producer, warp 1 = MMA issuer (owns TMEM), warps 2-5 = converters
// (hi/lo split of the A tile in shared memory), warps 6-9 = epilogue.  Two TMEM accumulator buffers let the epilogue
// of tile j (TMEM -> registers -> smem transpose -> full-line global stores) overlap the main loop of tile j+1; the
// smem ring runs continuously across tile boundaries.
constexpr int TG_EPI_WARPS = 8;                  // two per TMEM lane quarter, alternating 32-column chunks
constexpr int TG_THREADS = 192 + 32 * TG_EPI_WARPS;

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// TS = true: the converter warps write A_hi / A_lo into tensor memory (tcgen05.st) and the MMAs take A from there, so a
// stage holds raw A + B_hi + B_lo only (48 KB at BN = 128 -> 4 stages) and the shared-memory port carries 130 KB per
// K-block instead of 210 KB (TMA 48 + converter 16 r [+ 32 w] + MMA operands 48 [+ 48] + epilogue 18): the SS form was
// bound by exactly that port (measured cadence 1550 cycles / K-block = 210 KB / 128 B per cycle).  Needs BN <= 128
// (TMEM: 2 accumulator buffers + 4 x 64 A columns = 512).
// ring position without a runtime modulo / division per K-block
struct StageIter {
  int s; uint32_t ph; int n;
  __device__ __forceinline__ void next() { if (++s == n) { s = 0; ph ^= 1u; } }
};
constexpr uint32_t TS_A_COLS = 64;      // TMEM columns per stage: A_hi [0,32) | A_lo [32,64)

template <bool TS>
__global__ void __launch_bounds__(TG_THREADS, 1)
k_tc_gemm(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
          const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo, const TcGemmParams p) {
  pdl_trigger();       // the next kernel of the chain may be scheduled; it blocks in its own pdl_wait until this grid is done
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const uint32_t a_bytes = TC_BM * TC_BK * 4;                 // 16 KB
  const uint32_t b_bytes = (uint32_t)p.BN * TC_BK * 4;
  const uint32_t a_span = TS ? a_bytes : 2 * a_bytes;          // SS keeps a second (lo) plane of A in the stage
  const uint32_t stage_bytes = a_span + 2 * b_bytes;
  float* epi_stage = reinterpret_cast<float*>(smem + (size_t)p.stages * stage_bytes);     // 8 warps x [32][20] floats
  float* bias_s = epi_stage + TG_EPI_WARPS * EPI16_PATCH;                                  // [288]
  uint64_t* full_raw = reinterpret_cast<uint64_t*>(bias_s + 288);
  uint64_t* full_conv = full_raw + TC_MAX_STAGES;
  uint64_t* empty = full_conv + TC_MAX_STAGES;
  uint64_t* a_free = empty + TC_MAX_STAGES;         // TS: the converters have read the raw A tile out of the stage
  uint64_t* tmem_full = a_free + TC_MAX_STAGES;     // [2]
  uint64_t* tmem_empty = tmem_full + 2;             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int32_t KB = p.kblocks1 + p.kblocks2;
  int32_t M = p.M;                                     // (the device-side row count is read after pdl_wait below)
  uint32_t buf_cols = 32;
  while (buf_cols < (uint32_t)p.BN) buf_cols <<= 1;
  const uint32_t tmem_cols = TS ? 512u : 2 * buf_cols;
  const uint32_t tmem_a0 = 2 * buf_cols;                       // TS: first A column
  long long* tr = (p.trace != nullptr && blockIdx.x == 0) ? p.trace : nullptr;
  if (tr && threadIdx.x == 0) tr[0] = clock64();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA1); tma_prefetch_desc(&tmA2); tma_prefetch_desc(&tmBhi); tma_prefetch_desc(&tmBlo);
    // full_raw: the raw A tile landed (converters wait); full_conv: 128 converter arrivals + the producer's arrival
    // carrying the B tiles' bytes (the MMA thread waits on this one only)
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_raw[s], 1); mbar_init(&full_conv[s], 129); mbar_init(&empty[s], 1); mbar_init(&a_free[s], 128);
    }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 32 * TG_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above (barriers, tensor-memory allocation, descriptor prefetch) touched no data of the preceding kernel and
  // ran under its tail; nothing below may start before it has completed
  pdl_wait();
  if (p.M_dev != nullptr) { const int32_t v = __ldg(p.M_dev); M = v < p.M ? (v < 0 ? 0 : v) : p.M; }
  const int32_t m_tiles = (M + TC_BM - 1) / TC_BM;
  const int32_t total_tiles = m_tiles * p.tiles_per_seg * p.num_segs;

  // tile -> (m tile, segment, n tile): consecutive tiles walk M so that co-running CTAs share the same B tile in L2
  auto decode = [&](int32_t tile, int32_t& m0, int32_t& seg_id, int32_t& n0) {
    const int32_t n_idx = tile / m_tiles;
    m0 = (tile - n_idx * m_tiles) * TC_BM;
    seg_id = n_idx / p.tiles_per_seg;
    n0 = (n_idx - seg_id * p.tiles_per_seg) * p.BN;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0;
      StageIter si{0, 0u, p.stages};
      for (int32_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int32_t m0, seg_id, n0;
        decode(tile, m0, seg_id, n0);
        const int32_t b_row = p.seg[seg_id].b_row0 + n0;
        for (int32_t kb = 0; kb < KB; ++kb, ++it, si.next()) {
          const int s = si.s;
          const uint32_t ph = si.ph;
          uint8_t* st = smem + (size_t)s * stage_bytes;
          const bool first = kb < p.kblocks1;
          const int32_t ka = (first ? kb + p.kb_begin1 : kb - p.kblocks1 + p.kb_begin2) * TC_BK;
          const int32_t kbk = first ? ka : p.b_koff2 + ka;
          // TS: the raw A slot is free as soon as the converters have copied it to registers — about one MMA period before
          // the stage's MMAs retire — so the A tile (the one with a conversion step behind it) is requested that much earlier:
          // the loop TMA latency + conversion + MMA no longer has to fit into `stages` MMA periods.
          if (TS) mbar_wait(&a_free[s], ph ^ 1u); else mbar_wait(&empty[s], ph ^ 1u);
          if (tr && it < 32) tr[8 + it * 8 + 0] = clock64();
          mbar_arrive_expect_tx(&full_raw[s], a_bytes);
          tma_load_2d(st, first ? &tmA1 : &tmA2, &full_raw[s], ka, m0);
          if (TS) mbar_wait(&empty[s], ph ^ 1u);       // B tiles (and the stage's TMEM A columns) are in use until the MMAs retire
          mbar_arrive_expect_tx(&full_conv[s], 2 * b_bytes);
          tma_load_2d(st + a_span, &tmBhi, &full_conv[s], kbk, b_row);
          tma_load_2d(st + a_span + b_bytes, &tmBlo, &full_conv[s], kbk, b_row);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs the loop in convergence (warp-uniform operands, see umma_tf32_elect); ONE barrier wait per
    // K-block: the B tiles complete_tx on full_conv too, so its phase means "B landed and A converted".
    const uint32_t idesc = umma_idesc_tf32(TC_BM, (uint32_t)p.BN);
    uint32_t it = 0, j = 0;
    StageIter si{0, 0u, p.stages};
    for (int32_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++j) {
      const uint32_t ab = j & 1u;
      mbar_wait(&tmem_empty[ab], ((j >> 1) & 1u) ^ 1u);        // epilogue has drained this accumulator buffer
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + ab * buf_cols;
      for (int32_t kb = 0; kb < KB; ++kb, ++it, si.next()) {
        const int s = si.s;
        const uint32_t ph = si.ph;
        mbar_wait(&full_conv[s], ph);
        tc_fence_after();
        if (tr && lane == 0 && it < 32) tr[8 + it * 8 + 3] = clock64();
        const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
        const uint32_t a_hi = sa, a_lo = sa + a_bytes, b_hi = sa + a_span, b_lo = b_hi + b_bytes;
        const uint32_t ta = tmem_base + tmem_a0 + (uint32_t)s * TS_A_COLS;
        // a ragged K (products layer 1: K = 200 = 6 K-blocks + 8 floats) leaves all-zero k-steps in the last K-block: not issued
        const int nk = (kb == p.kblocks1 - 1) ? p.ktail1 : (kb == KB - 1 ? p.ktail2 : TC_BK / 8);
#pragma unroll
        for (int k = 0; k < TC_BK / 8; ++k) {
          if (k >= nk) break;
          const uint32_t koff = k * 32;   // 8 tf32 = 32 bytes inside the 128-byte swizzle atom
          const uint64_t dbh = umma_desc_k_sw128(b_hi + koff), dbl = umma_desc_k_sw128(b_lo + koff);
          if (TS) {
            umma_tf32_ts_elect(tmem_d, ta + k * 8, dbh, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma_tf32_ts_elect(tmem_d, ta + 32 + k * 8, dbh, idesc, 1u);
            umma_tf32_ts_elect(tmem_d, ta + k * 8, dbl, idesc, 1u);
          } else {
            const uint64_t dah = umma_desc_k_sw128(a_hi + koff), dal = umma_desc_k_sw128(a_lo + koff);
            umma_tf32_elect(tmem_d, dah, dbh, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma_tf32_elect(tmem_d, dal, dbh, idesc, 1u);
            umma_tf32_elect(tmem_d, dah, dbl, idesc, 1u);
          }
        }
        if (tr && lane == 0 && it < 32) tr[8 + it * 8 + 4] = clock64();
        umma_commit_elect(&empty[s]);                          // stage reusable once these MMAs have read it
        if (kb == KB - 1) umma_commit_elect(&tmem_full[ab]);   // accumulator complete
      }
    }
  } else if (warp < 6) {
    // ===================== converter warps (2..5) =====================
    const int t = threadIdx.x - 64;                  // 0..127
    uint32_t it = 0;
    StageIter si{0, 0u, p.stages};
    for (int32_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      for (int32_t kb = 0; kb < KB; ++kb, ++it, si.next()) {
        const int s = si.s;
        const uint32_t ph = si.ph;
        mbar_wait(&full_raw[s], ph);
        if (tr && t == 0 && it < 32) tr[8 + it * 8 + 1] = clock64();
        if (TS) {
          // thread = one tile row (TMEM lane 32*(warp%4) + lane): its 128 swizzled bytes -> registers (the shared-memory slot
          // is handed back at once) -> hi / lo -> tensor memory (once the MMAs of the stage's previous use have retired)
          const int r = (warp & 3) * 32 + lane;
          const uint8_t* rowp = smem + (size_t)s * stage_bytes + (size_t)(r >> 3) * 1024 + (size_t)(r & 7) * 128;
          const uint32_t ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + tmem_a0 + (uint32_t)s * TS_A_COLS;
          float4 raw[8];
#pragma unroll
          for (int j = 0; j < 8; ++j)                      // 16-byte chunk j of the row sits at chunk position j ^ (r % 8)
            raw[j] = *reinterpret_cast<const float4*>(rowp + ((j ^ (r & 7)) << 4));
          mbar_arrive(&a_free[s]);
          mbar_wait(&empty[s], ph ^ 1u);
          tc_fence_after();
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t h[16], l[16];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const float4 x = raw[half * 4 + jj];
              float hx, lx;
              split_tf32(x.x, hx, lx); h[4 * jj + 0] = __float_as_uint(hx); l[4 * jj + 0] = __float_as_uint(lx);
              split_tf32(x.y, hx, lx); h[4 * jj + 1] = __float_as_uint(hx); l[4 * jj + 1] = __float_as_uint(lx);
              split_tf32(x.z, hx, lx); h[4 * jj + 2] = __float_as_uint(hx); l[4 * jj + 2] = __float_as_uint(lx);
              split_tf32(x.w, hx, lx); h[4 * jj + 3] = __float_as_uint(hx); l[4 * jj + 3] = __float_as_uint(lx);
            }
            tmem_st16(ta + half * 16, h);
            tmem_st16(ta + 32 + half * 16, l);
          }
          tmem_st_wait();
          tc_fence_before();
        } else {
          float4* hi4 = reinterpret_cast<float4*>(smem + (size_t)s * stage_bytes);
          float4* lo4 = reinterpret_cast<float4*>(smem + (size_t)s * stage_bytes + a_bytes);
#pragma unroll
          for (int q4 = 0; q4 < (TC_BM * TC_BK / 4) / 128; ++q4) {
            const int i = t + 128 * q4;
            const float4 x = hi4[i];
            float4 h, l;
            split_tf32(x.x, h.x, l.x); split_tf32(x.y, h.y, l.y); split_tf32(x.z, h.z, l.z); split_tf32(x.w, h.w, l.w);
            hi4[i] = h;
            lo4[i] = l;
          }
          fence_proxy_async_smem();                  // generic-proxy writes -> visible to the tensor core (async proxy)
        }
        mbar_arrive(&full_conv[s]);
        if (tr && t == 0 && it < 32) tr[8 + it * 8 + 2] = clock64();
      }
    }
  } else {
    // ===================== epilogue warps (6..13) =====================
    // Two warps per TMEM lane quarter (a warp may only touch lanes 32*(warp%4)..+31); the pair alternates the tile's
    // 32-column chunks.  One warp per scheduler could not hide its own latencies (Philox chains, tcgen05.ld, smem
    // transpose): ncu showed the epilogue as long as the main loop it is supposed to hide under.
    const int et = threadIdx.x - 192;                // 0..255
    const int q = warp & 3;                          // TMEM lane quarter this warp may access
    const int half = (warp - 6) >> 2;                // which chunks of the tile: (c / 32) % 2 == half
    float* stage = epi_stage + (warp - 6) * EPI16_PATCH;
    const uint32_t thr = dropout_threshold(p.drop_p);
    const float keep_scale = p.drop_p > 0.f ? 1.0f / (1.0f - p.drop_p) : 1.0f;
    uint32_t off_lo = p.off_lo, off_hi = p.off_hi;
    if (p.ctl != nullptr) {                          // replayed step: the dropout stream offset lives on the device
      const uint64_t o = (((uint64_t)p.ctl->drop_off_hi << 32) | p.ctl->drop_off_lo) + p.ctl_layer;
      off_lo = (uint32_t)o; off_hi = (uint32_t)(o >> 32);
    }
    // last chunk this warp reads from TMEM (-1: none — it hands the buffer back right away)
    int32_t last_c = -1;
    for (int32_t c = 32 * half; c < p.BN; c += 64) last_c = c;
    uint32_t j = 0;
    for (int32_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++j) {
      int32_t m0, seg_id, n0;
      decode(tile, m0, seg_id, n0);
      const TcSegment sg = p.seg[seg_id];
      const uint32_t ab = j & 1u;
      // bias of this tile's columns -> smem (the previous tile's readers are past their last use: barrier first)
      epi_bar_sync();
      for (int i = et; i < 288; i += 32 * TG_EPI_WARPS) bias_s[i] = (p.bias != nullptr && n0 + i < sg.n_cols) ? __ldg(p.bias + n0 + i) : 0.f;
      epi_bar_sync();
      const int64_t row0 = (int64_t)m0 + q * 32;
      const int64_t m = row0 + lane;
      const bool row_ok = m < M;
      float rs = 1.0f;
      if (sg.scale_rows && p.rowptr != nullptr && row_ok)
        rs = 1.0f / (float)max(__ldg(p.rowptr + m + 1) - __ldg(p.rowptr + m), 1);
      const bool vec_ok = ((sg.ld_out & 3) == 0) && ((reinterpret_cast<uintptr_t>(sg.out) & 15) == 0);
      mbar_wait(&tmem_full[ab], (j >> 1) & 1u);
      tc_fence_after();
      if (tr && et == 0 && j == 0) tr[1] = clock64();
      if (last_c < 0) { tc_fence_before(); mbar_arrive(&tmem_empty[ab]); }
      const uint32_t tmem_d = tmem_base + ab * buf_cols + ((uint32_t)(q * 32) << 16);
      for (int32_t c = 32 * half; c < p.BN; c += 64) {
        // a 32-column chunk as two 16-column units (16 accumulator + 16 result registers live at a time: the kernel
        // runs 448 threads, 128 registers each)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int32_t cc = c + 16 * h;
          const int32_t nb = n0 + cc;
          const bool live = cc < p.BN && nb < sg.n_cols;         // warp-uniform
          uint32_t v[16];
          if (live) tmem_ld16(tmem_d + (uint32_t)cc, v);
          if (c == last_c && h == 1) {                            // this warp's last TMEM read of the tile: hand the buffer back
            tc_fence_before();
            mbar_arrive(&tmem_empty[ab]);
          }
          if (!live) continue;
          uint32_t keep16 = 0xFFFFu;                              // nb is a multiple of 16: two 8-column Philox groups
          if (p.drop_p > 0.f && row_ok && p.final) {
            keep16 = dropout_keep8((uint32_t)m, (uint32_t)(nb >> 3), p.seed_lo, p.seed_hi, off_lo, off_hi, thr);
            if (nb + 8 < sg.n_cols)
              keep16 |= dropout_keep8((uint32_t)m, (uint32_t)(nb >> 3) + 1u, p.seed_lo, p.seed_hi, off_lo, off_hi, thr) << 8;
          }
          float r[16];
          if (p.acc_in) {
            // a contraction longer than one tensor-memory accumulation chain: this launch continues the partial result the
            // previous one stored (each lane owns a row: 16 consecutive floats)
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
              const float prev = (row_ok && nb + jj < sg.n_cols) ? sg.out[m * sg.ld_out + nb + jj] : 0.f;
              v[jj] = __float_as_uint(__uint_as_float(v[jj]) + prev);
            }
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const float4 b4 = *reinterpret_cast<const float4*>(bias_s + cc + 4 * g);     // smem broadcast
            const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              float x = __uint_as_float(v[4 * g + jj]);
              if (p.final) {
                x = (x + bb[jj]) * rs;
                if (p.act == NGNN_ACT_RELU) x = fmaxf(x, 0.f);
                if (p.drop_p > 0.f) x = ((keep16 >> (4 * g + jj)) & 1u) ? x * keep_scale : 0.f;
              }
              r[4 * g + jj] = x;
            }
          }
          epilogue_store16(stage, r, sg.out, sg.ld_out, row0, M, nb, sg.n_cols, vec_ok, lane);
        }
      }
      if (tr && et == 0 && j == 0) tr[3] = clock64();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tr && threadIdx.x == 0) tr[2] = clock64();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, tmem_cols); }
}

// ----------------------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(ptr);
    else
      cudaGetLastError();
  }
  return fn;
}

// 2-D fp32 row-major [rows, cols] with leading dimension ld (elements); box = [box_rows x 32 floats], SWIZZLE_128B
static inline bool make_tmap_2d(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, uint32_t box_rows,
                                uint32_t box_cols = TC_BK, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

static inline bool tma_addressable(const float* p, int64_t ld) { return p != nullptr && is_aligned(p, 16) && (ld % 4 == 0) && ld > 0; }

static inline int32_t round_up_i(int64_t x, int64_t a) { return (int32_t)((x + a - 1) / a * a); }

static int g_tc_bn_max = 128;   // ngnn_set_tuning(4, 128|256): widest N tile.  128 -> 3 smem stages (measured 2x faster
                                // than 256 -> 2 stages on the products layer-1 shape: TMA latency is the limiter)

static int g_tc_ts = 1;         // ngnn_set_tuning(6, 0|1): A operand from tensor memory (TS form) when BN <= 128

struct TcPlan {
  int32_t BN, tiles_per_seg, stages;
  uint32_t smem_bytes;
  bool ts;
};
static inline TcPlan tc_plan(int64_t n_cols) {
  TcPlan pl;
  pl.BN = n_cols >= g_tc_bn_max ? g_tc_bn_max : round_up_i(n_cols, 16);
  pl.tiles_per_seg = (int32_t)ceil_div(n_cols, pl.BN);
  pl.ts = g_tc_ts != 0 && pl.BN <= 128;
  const uint32_t stage = (pl.ts ? 1u : 2u) * TC_BM * TC_BK * 4u + 2u * (uint32_t)pl.BN * TC_BK * 4u;
  const uint32_t fixed = 1024u /*align*/ + (uint32_t)TG_EPI_WARPS * EPI16_PATCH * 4u /*epilogue staging*/ + 288u * 4u /*bias*/ + 256u /*barriers*/;
  int st = (int)((TC_SMEM_LIMIT - fixed) / stage);
  pl.stages = st > TC_MAX_STAGES ? TC_MAX_STAGES : st;
  pl.smem_bytes = (uint32_t)pl.stages * stage + fixed;
  return pl;
}

// workspace: hi + lo planes of the packed weights
static inline size_t tc_fwd_ws_bytes(int64_t F, int64_t O) {
  const int64_t Kpack = 2 * (int64_t)round_up_i(F, TC_BK);
  return 2 * align_up((size_t)O * Kpack * sizeof(float), 256) + 256;
}
static inline size_t tc_dgrad_ws_bytes(int64_t F, int64_t O) {
  const int64_t Kpack = round_up_i(O, TC_BK);
  const int64_t rows = 2 * (int64_t)round_up_i(F, 16);
  return 2 * align_up((size_t)rows * Kpack * sizeof(float), 256) + 256;
}

static inline int32_t tc_launch(const CUtensorMap& a1, const CUtensorMap& a2, const CUtensorMap& bh, const CUtensorMap& bl,
                                const TcGemmParams& p, const TcPlan& pl, int num_segs, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    NGNN_CUDA(cudaFuncSetAttribute(k_tc_gemm<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_LIMIT));
    NGNN_CUDA(cudaFuncSetAttribute(k_tc_gemm<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_LIMIT));
    attr_set = true;
  }
  TcGemmParams pp = p;
  pp.trace = g_tc_trace;
  pp.num_segs = num_segs;
  const int64_t tiles = ceil_div(p.M, TC_BM) * pl.tiles_per_seg * num_segs;
  const unsigned grid = (unsigned)(tiles < kNumSMs ? tiles : kNumSMs);      // persistent: one CTA per SM
  if (pl.ts) launch_chain(k_tc_gemm<true>, dim3(grid), dim3(TG_THREADS), pl.smem_bytes, st, a1, a2, bh, bl, pp);
  else launch_chain(k_tc_gemm<false>, dim3(grid), dim3(TG_THREADS), pl.smem_bytes, st, a1, a2, bh, bl, pp);
  NGNN_LAUNCH_CHECK();
  return NGNN_OK;
}

// out = drop(act(a_l W_l^T + a_r W_r^T + b)).  Returns NGNN_E_UNSUPPORTED (no error text) when the
// operands are not TMA-addressable; the caller then uses the SIMT kernel.
// Packed hi / lo weight planes of the forward projection ([W_l | W_r], K-major) into ws.  `use_l` / `use_r`: which operand
// pairs the GEMM will contract (a missing one is zero-filled).
// A tensor-memory accumulation chain longer than this many K-blocks (x 12 MMAs, each rounding the accumulator toward zero)
// drifts past the 1e-5 bar (see the K-WGRAD notes below): a wider forward contraction is cut into several launches of at most
// TC_CHAIN_KBLOCKS, each adding the output of the one before (fp32, round to nearest) — F = 767 / 1433 stay on the tensor cores.
constexpr int TC_MAX_CHAIN_KBLOCKS = 40;
constexpr int TC_CHAIN_KBLOCKS = 32;

// concat_k: the two operands are the halves of ONE [n, 2F] matrix ([mean | root] side by side, layer 1 of the fused step), so the
// weights are packed [W_l | W_r] without padding in between and the contraction runs over ceil(2F / 32) K-blocks instead of
// 2 * ceil(F / 32) (F = 100: 7 instead of 8).
static inline int32_t tc_prep_fwd(const float* w_l, const float* w_r, bool use_l, bool use_r, int64_t F, int64_t O, void* ws,
                                  size_t ws_bytes, cudaStream_t st, PrepParams* collect = nullptr, bool concat_k = false) {
  if (F < 1 || O < 1) return NGNN_E_UNSUPPORTED;
  if (ws == nullptr || ws_bytes < tc_fwd_ws_bytes(F, O)) return NGNN_E_UNSUPPORTED;
  const int32_t Fpad = concat_k ? (int32_t)F : round_up_i(F, TC_BK), Kpack = concat_k ? round_up_i(2 * F, TC_BK) : 2 * Fpad;
  float* hi = reinterpret_cast<float*>(align_up(reinterpret_cast<size_t>(ws), 256));
  float* lo = hi + align_up((size_t)O * Kpack * sizeof(float), 256) / sizeof(float);
  PrepParams pp{};
  pp.w[0] = use_l ? w_l : nullptr; pp.w[1] = use_r ? w_r : nullptr;
  pp.R = (int32_t)O; pp.C = (int32_t)F; pp.transpose = 0; pp.seg_stride = Fpad; pp.Kpack = Kpack; pp.rows_out = (int32_t)O;
  pp.hi = hi; pp.lo = lo;
  if (collect) { *collect = pp; return NGNN_OK; }
  k_prep_weights<<<(unsigned)ceil_div((int64_t)O * Kpack, 256), 256, 0, st>>>(pp);
  NGNN_LAUNCH_CHECK();
  return NGNN_OK;
}

static inline int32_t tc_gemm_fwd(const float* a_l, int64_t ld_al, const float* a_r, int64_t ld_ar, const float* w_l,
                                  const float* w_r, const float* bias, int64_t n, int64_t F, int64_t O, int32_t act,
                                  float drop_p, uint64_t seed, uint64_t offset, float* out, int64_t ld_out, void* ws,
                                  size_t ws_bytes, cudaStream_t st, bool prepped = false, const int32_t* n_dev = nullptr,
                                  const StepCtl* ctl = nullptr, uint32_t ctl_layer = 0, bool concat_k = false) {
  if (F < 1 || n < 1 || O < 1 || n >= (1LL << 31) - 256) return NGNN_E_UNSUPPORTED;
  if (concat_k && !(a_l && a_r && a_r == a_l + F && ld_al == ld_ar && ld_al >= 2 * F)) return NGNN_E_UNSUPPORTED;
  if (a_l && !tma_addressable(a_l, ld_al)) return NGNN_E_UNSUPPORTED;
  if (a_r && !tma_addressable(a_r, ld_ar)) return NGNN_E_UNSUPPORTED;
  if (!a_l && !a_r) return NGNN_E_UNSUPPORTED;
  if (ws == nullptr || ws_bytes < tc_fwd_ws_bytes(F, O) || get_encode_fn() == nullptr) return NGNN_E_UNSUPPORTED;

  const int32_t Fpad = concat_k ? (int32_t)F : round_up_i(F, TC_BK), Kpack = concat_k ? round_up_i(2 * F, TC_BK) : 2 * Fpad;
  float* hi = reinterpret_cast<float*>(align_up(reinterpret_cast<size_t>(ws), 256));
  float* lo = hi + align_up((size_t)O * Kpack * sizeof(float), 256) / sizeof(float);
  if (!prepped) {
    const int32_t rc = tc_prep_fwd(w_l, w_r, a_l != nullptr, a_r != nullptr, F, O, ws, ws_bytes, st, nullptr, concat_k);
    if (rc != NGNN_OK) return rc;
  }

  const TcPlan pl = tc_plan(O);
  CUtensorMap tA1, tA2, tBh, tBl;
  const float* a1 = a_l ? a_l : a_r;
  const int64_t ld1 = a_l ? ld_al : ld_ar;
  const bool two = a_l && a_r && !concat_k;
  bool ok = make_tmap_2d(&tA1, a1, n, concat_k ? 2 * F : F, ld1, TC_BM);
  ok = ok && make_tmap_2d(&tA2, two ? a_r : a1, n, F, two ? ld_ar : ld1, TC_BM);
  ok = ok && make_tmap_2d(&tBh, hi, O, Kpack, Kpack, (uint32_t)pl.BN);
  ok = ok && make_tmap_2d(&tBl, lo, O, Kpack, Kpack, (uint32_t)pl.BN);
  NGNN_REQUIRE(ok, NGNN_E_CUDA, "gemm_fwd: cuTensorMapEncodeTiled failed");

  TcGemmParams p{};
  p.M = (int32_t)n; p.M_dev = n_dev; p.ctl = ctl; p.ctl_layer = ctl_layer;
  p.BN = pl.BN; p.stages = pl.stages; p.tiles_per_seg = pl.tiles_per_seg;
  p.kblocks1 = concat_k ? Kpack / TC_BK : Fpad / TC_BK; p.kblocks2 = two ? Fpad / TC_BK : 0;
  p.b_koff2 = Fpad;
  if (!a_l) { p.b_koff2 = 0; /* single operand is a_r: its weights sit in segment 1 of the pack */ }
  p.seg[0] = TcSegment{out, ld_out, (int32_t)O, 0, 0};
  p.bias = bias; p.act = act; p.drop_p = drop_p;
  p.seed_lo = (uint32_t)seed; p.seed_hi = (uint32_t)(seed >> 32); p.off_lo = (uint32_t)offset; p.off_hi = (uint32_t)(offset >> 32);
  CUtensorMap shifted, shifted_lo;
  const CUtensorMap* mBh = &tBh;
  const CUtensorMap* mBl = &tBl;
  if (!a_l) {
    // only the root term: read its weights from k offset Fpad of the pack
    p.b_koff2 = 0;
    ok = make_tmap_2d(&shifted, hi + Fpad, O, Fpad, Kpack, (uint32_t)pl.BN);
    ok = ok && make_tmap_2d(&shifted_lo, lo + Fpad, O, Fpad, Kpack, (uint32_t)pl.BN);
    NGNN_REQUIRE(ok, NGNN_E_CUDA, "gemm_fwd: cuTensorMapEncodeTiled failed");
    mBh = &shifted; mBl = &shifted_lo;
  }
  // one launch per accumulation chain of at most TC_CHAIN_KBLOCKS K-blocks (a single launch for every F <= 512)
  const int32_t kb1 = p.kblocks1, kb2 = p.kblocks2, total = kb1 + kb2;
  const int64_t k_op = concat_k ? 2 * F : F;                         // floats of data along K per operand
  const int32_t tail = (int32_t)((k_op - 1) % TC_BK) / 8 + 1, tail1 = tail, tail2 = tail;
  const int32_t chains = (total + TC_MAX_CHAIN_KBLOCKS - 1) / TC_MAX_CHAIN_KBLOCKS <= 1 ? 1 : (total + TC_CHAIN_KBLOCKS - 1) / TC_CHAIN_KBLOCKS;
  const int32_t per = (total + chains - 1) / chains;
  for (int32_t c = 0; c < chains; ++c) {
    const int32_t g0 = c * per, g1 = g0 + per < total ? g0 + per : total;       // global K-block range of this launch
    const int32_t o1b = g0 < kb1 ? g0 : kb1, o1e = g1 < kb1 ? g1 : kb1;         // part in operand 1
    const int32_t b0 = g0 > kb1 ? g0 - kb1 : 0, b1 = g1 > kb1 ? g1 - kb1 : 0;   // part in operand 2
    TcGemmParams pc = p;
    pc.kblocks1 = o1e - o1b; pc.kb_begin1 = o1b; pc.kblocks2 = b1 - b0; pc.kb_begin2 = b0;
    pc.ktail1 = (o1e == kb1) ? tail1 : 4; pc.ktail2 = (b1 == kb2) ? tail2 : 4;
    pc.acc_in = c > 0; pc.final = c == chains - 1;
    const int32_t rc = tc_launch(tA1, tA2, *mBh, *mBl, pc, pl, 1, st);
    if (rc != NGNN_OK) return rc;
  }
  return NGNN_OK;
}

// dmean_scaled = rowscale * (dy W_l), dx_root = dy W_r, both in one launch (two N segments).
// Packed hi / lo planes of [W_l^T ; W_r^T] (rows = F per segment, K = O) for the data gradient.
static inline int32_t tc_prep_dgrad(const float* w_l, const float* w_r, bool use_l, bool use_r, int64_t F, int64_t O, void* ws,
                                    size_t ws_bytes, cudaStream_t st, PrepParams* collect = nullptr) {
  if (F < 1 || O < 1 || ceil_div(O, TC_BK) > TC_MAX_CHAIN_KBLOCKS) return NGNN_E_UNSUPPORTED;
  if (ws == nullptr || ws_bytes < tc_dgrad_ws_bytes(F, O)) return NGNN_E_UNSUPPORTED;
  const int32_t Kpack = round_up_i(O, TC_BK), Rpad = round_up_i(F, 16);
  const int32_t rows = 2 * Rpad;
  float* hi = reinterpret_cast<float*>(align_up(reinterpret_cast<size_t>(ws), 256));
  float* lo = hi + align_up((size_t)rows * Kpack * sizeof(float), 256) / sizeof(float);
  PrepParams pp{};
  pp.w[0] = use_l ? w_l : nullptr; pp.w[1] = use_r ? w_r : nullptr;
  pp.R = (int32_t)O; pp.C = (int32_t)F; pp.transpose = 1; pp.seg_stride = Rpad; pp.Kpack = Kpack; pp.rows_out = rows;
  pp.hi = hi; pp.lo = lo;
  if (collect) { *collect = pp; return NGNN_OK; }
  k_prep_weights<<<(unsigned)ceil_div((int64_t)rows * Kpack, 256), 256, 0, st>>>(pp);
  NGNN_LAUNCH_CHECK();
  return NGNN_OK;
}

static inline int32_t tc_gemm_dgrad(const float* dy, int64_t ld_dy, const float* w_l, const float* w_r, const int32_t* rowptr,
                                    int64_t n, int64_t F, int64_t O, float* dmean, int64_t ld_dmean, float* droot,
                                    int64_t ld_root, void* ws, size_t ws_bytes, cudaStream_t st, bool prepped = false,
                                    const int32_t* n_dev = nullptr) {
  if (n < 1 || F < 1 || O < 1 || n >= (1LL << 31) - 256) return NGNN_E_UNSUPPORTED;
  if (ceil_div(O, TC_BK) > TC_MAX_CHAIN_KBLOCKS) return NGNN_E_UNSUPPORTED;
  if (!tma_addressable(dy, ld_dy)) return NGNN_E_UNSUPPORTED;
  if (ws == nullptr || ws_bytes < tc_dgrad_ws_bytes(F, O) || get_encode_fn() == nullptr) return NGNN_E_UNSUPPORTED;
  if (!dmean && !droot) return NGNN_OK;

  const int32_t Kpack = round_up_i(O, TC_BK), Rpad = round_up_i(F, 16);
  const int32_t rows = 2 * Rpad;
  float* hi = reinterpret_cast<float*>(align_up(reinterpret_cast<size_t>(ws), 256));
  float* lo = hi + align_up((size_t)rows * Kpack * sizeof(float), 256) / sizeof(float);
  if (!prepped) {
    const int32_t rc = tc_prep_dgrad(w_l, w_r, dmean != nullptr, droot != nullptr, F, O, ws, ws_bytes, st);
    if (rc != NGNN_OK) return rc;
  }

  const TcPlan pl = tc_plan(F);
  CUtensorMap tA, tBh, tBl;
  bool ok = make_tmap_2d(&tA, dy, n, O, ld_dy, TC_BM);
  ok = ok && make_tmap_2d(&tBh, hi, rows, Kpack, Kpack, (uint32_t)pl.BN);
  ok = ok && make_tmap_2d(&tBl, lo, rows, Kpack, Kpack, (uint32_t)pl.BN);
  NGNN_REQUIRE(ok, NGNN_E_CUDA, "dgrad: cuTensorMapEncodeTiled failed");

  TcGemmParams p{};
  p.M = (int32_t)n; p.M_dev = n_dev; p.BN = pl.BN; p.stages = pl.stages; p.tiles_per_seg = pl.tiles_per_seg;
  p.kblocks1 = Kpack / TC_BK; p.kblocks2 = 0; p.b_koff2 = 0; p.final = 1;
  p.ktail1 = (int32_t)((O - 1) % TC_BK) / 8 + 1; p.ktail2 = 4;
  p.rowptr = rowptr;
  int ns = 0;
  if (dmean) p.seg[ns++] = TcSegment{dmean, ld_dmean, (int32_t)F, 0, rowptr != nullptr ? 1 : 0};
  if (droot) p.seg[ns++] = TcSegment{droot, ld_root, (int32_t)F, Rpad, 0};
  return tc_launch(tA, tA, tBh, tBl, p, pl, ns, st);
}

// ----------------------------------------------------------------------------- K-WGRAD on tensor cores
//   dW_seg[o, f] = sum_i dY[i, o] * X_seg[i, f]          (seg 0: X = mean -> dW_l, seg 1: X = root rows -> dW_r)
// Both operands are MN-major for the UMMA (the reduction index i is the slow dimension of dY [n,O] and X [n,F]).
// For 32-bit MN-major operands the only UMMA shared-memory layout is SWIZZLE_128B_BASE32B: atoms of 4 k-rows x 128
// bytes (32 MN elements) with the 32-byte chunks of a row XOR-ed by (row % 4) — what TMA writes for a
// [32 rows(i) x 32 floats] box with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  The 32-wide MN groups of a tile are
// LBO = 4096 B (one box) apart, the 4-row k groups SBO = 512 B apart.  The long i dimension is split across CTAs (grid.z) into
// slices; each CTA writes its partial tile and k_reduce_partials sums them in a fixed order (deterministic).
//
// Accumulation accuracy.  The tensor core adds into its fp32 accumulator with round-toward-zero, so a long chain of MMAs
// drifts by up to one ulp PER MMA, always toward zero — a bias that grows linearly with the chain.  A weight gradient
// reduces over ~77 k rows (products layer 1: 2,400 K-blocks, 66 per slice, 12 MMAs each) and its terms largely cancel:
// measured on the full-scale products step, the 800-MMA chains of the first version put dW_l 8e-4 (relative) away from the
// fp64 oracle, where the fp32 CPU oracle is 8e-7 away.  So the chain is cut: the MMA warp accumulates TW_CHUNK K-blocks
// (96 MMAs) into one of two TMEM accumulator buffers, then eight "promoter" warps tcgen05.ld that chunk and add it to
// fp32 running sums held in registers (round-to-nearest FADD on the CUDA cores) while the next chunk's MMAs run into
// the other buffer.  The chunking depends only on (n, grid), so results stay bitwise reproducible.
constexpr int TW_KB = 32;                      // reduction rows per K-block
constexpr uint32_t TW_BOX_BYTES = TW_KB * 128; // one [32 x 32 floats] box
constexpr int TW_CHUNK = 8;                    // K-blocks (256 reduction rows) per tensor-memory accumulation chain
constexpr int TW_EPI_WARPS = 8;                // promoter / epilogue warps: two per TMEM lane quarter
constexpr int TW_THREADS = 192 + 32 * TW_EPI_WARPS;
constexpr int TW_BN_MAX = 128;                 // 2 accumulator buffers x 128 + 4 x 64 A columns = 512 TMEM columns

__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                       // LayoutType::SWIZZLE_128B_BASE32B
  return d;
}

struct TcWgradSeg {
  float* out;                 // partials [splits][O][n_cols] (or the final dW when splits == 1)
  int64_t split_stride;       // elements between split partials
  int32_t n_cols;             // F
  int32_t use_x2;             // 0: operand X1, 1: operand X2
};
struct TcWgradParams {
  int32_t O;                  // rows of dW
  int32_t BN;                 // UMMA N (multiple of 16, <= TW_BN_MAX)
  int32_t nbox;               // ceil(BN / 32) boxes of X per K-block
  int32_t stages;
  int32_t tiles_per_seg;
  int32_t n;                  // reduction rows (capacity when n_dev != nullptr)
  const int32_t* n_dev;       // optional device-side reduction length (clamped to n)
  float* db_part;             // optional [splits][O]: column sums of dY over this CTA's slice (bias gradient), TS form only
  TcWgradSeg seg[2];
};

// TS = true: dY^T (the M x K operand) is transposed + split by the converter warps straight into tensor memory
// (lane = output row o, column = reduction row i), X stays in shared memory (MN-major, split in place): the SS form
// moved 218 KB per K-block through the 128 B/cycle shared-memory port for 711 cycles of tensor work.
// Rows >= n of the last K-block are zeroed by the converters in both operands (the tensor maps may cover the buffers'
// capacity, so TMA's out-of-bounds zero fill cannot be relied on when n lives on the device).
template <bool TS>
__global__ void __launch_bounds__(TW_THREADS, 1)
k_tc_wgrad(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX1,
           const __grid_constant__ CUtensorMap tmX2, const TcWgradParams p) {
  pdl_trigger();
  pdl_wait();          // (n is read from device memory right below: nothing of the prologue can run ahead of the predecessor)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const uint32_t a_bytes = 4 * TW_BOX_BYTES;                   // 128 o-columns: 16 KB
  const uint32_t b_bytes = (uint32_t)p.nbox * TW_BOX_BYTES;
  const uint32_t a_span = TS ? a_bytes : 2 * a_bytes;
  const uint32_t stage_bytes = a_span + 2 * b_bytes;
  uint8_t* bar_base = smem + (size_t)p.stages * stage_bytes;
  uint64_t* full_raw = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* full_conv = full_raw + TC_MAX_STAGES;
  uint64_t* empty = full_conv + TC_MAX_STAGES;
  uint64_t* acc_full = empty + TC_MAX_STAGES;       // [2]
  uint64_t* acc_empty = acc_full + 2;               // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int32_t o0 = blockIdx.x * TC_BM;
  const int32_t seg_id = blockIdx.y / p.tiles_per_seg;
  const int32_t n_tile = blockIdx.y - seg_id * p.tiles_per_seg;
  const TcWgradSeg sg = p.seg[seg_id];
  const int32_t f0 = n_tile * p.BN;
  int32_t n = p.n;
  if (p.n_dev != nullptr) { const int32_t v = __ldg(p.n_dev); n = v < p.n ? (v < 0 ? 0 : v) : p.n; }
  const int32_t kb_total = (n + TW_KB - 1) / TW_KB;
  const int32_t kps = (kb_total + (int32_t)gridDim.z - 1) / (int32_t)gridDim.z;      // K-blocks per slice
  const int32_t kb_beg = min(kb_total, (int32_t)blockIdx.z * kps);
  const int32_t kb_end = min(kb_total, kb_beg + kps);
  const int32_t KB = kb_end - kb_beg;
  const int32_t n_chunks = (KB + TW_CHUNK - 1) / TW_CHUNK;
  uint32_t buf_cols = 32;
  while (buf_cols < (uint32_t)p.BN) buf_cols <<= 1;
  const uint32_t tmem_cols = TS ? 512u : 2 * buf_cols;         // TS: two accumulator buffers (<= 2 x 128) + 4 x 64 A columns
  const uint32_t tmem_a0 = 2 * buf_cols;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDY); tma_prefetch_desc(&tmX1); tma_prefetch_desc(&tmX2);
    // full_conv: TS = 128 A-converter + 256 B-converter (promoter warps) arrivals; SS = the 128 converter threads do both
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_raw[s], 1); mbar_init(&full_conv[s], TS ? 128 + 32 * TW_EPI_WARPS : 128); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 32 * TW_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const CUtensorMap* tmX = sg.use_x2 ? &tmX2 : &tmX1;
      StageIter si{0, 0u, p.stages};
      for (int32_t it = 0; it < KB; ++it, si.next()) {
        const int s = si.s;
        const uint32_t ph = si.ph;
        mbar_wait(&empty[s], ph ^ 1u);
        uint8_t* st = smem + (size_t)s * stage_bytes;
        mbar_arrive_expect_tx(&full_raw[s], a_bytes + b_bytes);
        const int32_t i0 = (kb_beg + it) * TW_KB;
#pragma unroll
        for (int g = 0; g < 4; ++g) tma_load_2d(st + g * TW_BOX_BYTES, &tmDY, &full_raw[s], o0 + 32 * g, i0);
        for (int g = 0; g < p.nbox; ++g) tma_load_2d(st + a_span + g * TW_BOX_BYTES, tmX, &full_raw[s], f0 + 32 * g, i0);
      }
    }
  } else if (warp == 1) {
    // whole warp in convergence, elect-issued MMAs (warp-uniform operands: see umma_tf32_elect)
    const uint32_t idesc = umma_idesc_tf32(TC_BM, (uint32_t)p.BN, TS ? 0 : 1, 1);   // SS: both operands MN-major; TS: A K-major in TMEM
    StageIter si{0, 0u, p.stages};
    for (int32_t c = 0; c < n_chunks; ++c) {
      const uint32_t ab = (uint32_t)c & 1u;
      mbar_wait(&acc_empty[ab], (((uint32_t)c >> 1) & 1u) ^ 1u);    // the promoters have drained this accumulator buffer
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + ab * buf_cols;
      const int32_t kbc = min(TW_CHUNK, KB - c * TW_CHUNK);
      for (int32_t kk = 0; kk < kbc; ++kk, si.next()) {
        const int s = si.s;
        mbar_wait(&full_conv[s], si.ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
        const uint32_t a_hi = sa, a_lo = sa + a_bytes, b_hi = sa + a_span, b_lo = b_hi + b_bytes;
        const uint32_t ta = tmem_base + tmem_a0 + (uint32_t)s * TS_A_COLS;
#pragma unroll
        for (int k = 0; k < TW_KB / 8; ++k) {
          const uint32_t koff = k * 1024;   // 8 reduction rows x 128 B
          const uint64_t dbh = umma_desc_mn_sw128(b_hi + koff, TW_BOX_BYTES, 512), dbl = umma_desc_mn_sw128(b_lo + koff, TW_BOX_BYTES, 512);
          if (TS) {
            umma_tf32_ts_elect(tmem_d, ta + k * 8, dbh, idesc, (kk > 0 || k > 0) ? 1u : 0u);
            umma_tf32_ts_elect(tmem_d, ta + 32 + k * 8, dbh, idesc, 1u);
            umma_tf32_ts_elect(tmem_d, ta + k * 8, dbl, idesc, 1u);
          } else {
            const uint64_t dah = umma_desc_mn_sw128(a_hi + koff, TW_BOX_BYTES, 512), dal = umma_desc_mn_sw128(a_lo + koff, TW_BOX_BYTES, 512);
            umma_tf32_elect(tmem_d, dah, dbh, idesc, (kk > 0 || k > 0) ? 1u : 0u);
            umma_tf32_elect(tmem_d, dal, dbh, idesc, 1u);
            umma_tf32_elect(tmem_d, dah, dbl, idesc, 1u);
          }
        }
        umma_commit_elect(&empty[s]);
        if (kk == kbc - 1) umma_commit_elect(&acc_full[ab]);       // this chunk's chain is complete
      }
    }
  } else if (warp < 6) {
    // ===================== converter warps (2..5): hi / lo split of both operands, rows >= n zeroed =====================
    const int t = threadIdx.x - 64;
    float db_acc = 0.f;                                         // TS: this thread's output row o, summed over the slice (bias gradient)
    StageIter si{0, 0u, p.stages};
    for (int32_t it = 0; it < KB; ++it, si.next()) {
      const int s = si.s;
      const uint32_t ph = si.ph;
      mbar_wait(&full_raw[s], ph);
      uint8_t* st = smem + (size_t)s * stage_bytes;
      const int32_t rows_ok = n - (kb_beg + it) * TW_KB;       // reduction rows of this K-block that exist (>= 32: all)
      if (TS) {
        // A: thread = output row o = 32*(warp%4) + lane = column `lane` of box (warp%4); reduction row i of that box is one
        // 128-byte line whose 32-byte chunks are XOR-ed with (i % 4) (SWIZZLE_128B_ATOM_32B) -> a warp reads one full
        // line per i, conflict-free.  32 values -> hi / lo -> 64 TMEM columns of this thread's lane.
        const int qa = warp & 3;
        const float* box = reinterpret_cast<const float*>(st + (size_t)qa * TW_BOX_BYTES);
        const uint32_t ta = tmem_base + ((uint32_t)(qa * 32) << 16) + tmem_a0 + (uint32_t)s * TS_A_COLS;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t h[16], l[16];
#pragma unroll
          for (int ii = 0; ii < 16; ++ii) {
            const int i = half * 16 + ii;
            float x = box[i * 32 + ((((lane >> 3) ^ (i & 3)) << 3) | (lane & 7))];
            if (i >= rows_ok) x = 0.f;
            db_acc += x;
            float hx, lx;
            split_tf32(x, hx, lx);
            h[ii] = __float_as_uint(hx); l[ii] = __float_as_uint(lx);
          }
          tmem_st16(ta + half * 16, h);
          tmem_st16(ta + 32 + half * 16, l);
        }
        tmem_st_wait();
        tc_fence_before();
      } else {
        const int a4 = (int)(a_bytes / 16), n4 = (int)((a_bytes + b_bytes) / 16);
        for (int i = t; i < n4; i += 128) {
          float4* hp; float4* lp;
          int q;
          if (i < a4) { q = i; hp = reinterpret_cast<float4*>(st) + i; lp = reinterpret_cast<float4*>(st + a_bytes) + i; }
          else { q = i - a4; hp = reinterpret_cast<float4*>(st + 2 * a_bytes) + q; lp = reinterpret_cast<float4*>(st + 2 * a_bytes + b_bytes) + q; }
          float4 x = *hp;
          if (((q & 255) >> 3) >= rows_ok) x = make_float4(0.f, 0.f, 0.f, 0.f);
          float4 h, l;
          split_tf32(x.x, h.x, l.x); split_tf32(x.y, h.y, l.y); split_tf32(x.z, h.z, l.z); split_tf32(x.w, h.w, l.w);
          *hp = h;
          *lp = l;
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(&full_conv[s]);
    }
    // bias gradient: db_part[slice][o] = sum over this slice's rows of dY[i, o]  (one CTA per (o tile, slice) writes it)
    if (TS && p.db_part != nullptr && blockIdx.y == 0) {
      const int32_t o = o0 + (warp & 3) * 32 + lane;
      if (o < p.O) p.db_part[(int64_t)blockIdx.z * p.O + o] = db_acc;
    }
  } else {
    // ===================== promoter / epilogue warps (6..13) =====================
    // Two warps per TMEM lane quarter; the pair alternates the tile's 16-column units.  Running sums of up to 4 units
    // (64 columns) per thread in registers; every finished chain is added in with round-to-nearest.
    // TS: these 256 threads also split the X tile (B operand) of EVERY K-block in place — the four converter warps alone,
    // doing dY^T and X one after the other, ran at 2,170 cycles per K-block against 1,125 of shared-memory-port time and
    // 672 of tensor time.  A chunk is promoted 4 K-blocks after its last one was converted: the MMAs trail the converters
    // by at most the ring depth, so that wait never stalls the conversion.
    const int q = warp & 3;
    const int half = (warp - 6) >> 2;
    const int pt = threadIdx.x - 192;                // 0..255
    float sum[4][16];
#pragma unroll
    for (int uu = 0; uu < 4; ++uu)
#pragma unroll
      for (int jj = 0; jj < 16; ++jj) sum[uu][jj] = 0.f;
    auto promote = [&](int32_t c) {
      const uint32_t ab = (uint32_t)c & 1u;
      mbar_wait(&acc_full[ab], ((uint32_t)c >> 1) & 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + ab * buf_cols + ((uint32_t)(q * 32) << 16);
#pragma unroll
      for (int uu = 0; uu < 4; ++uu) {
        const int32_t cc = 16 * (2 * uu + half);
        if (cc < p.BN) {                                           // warp-uniform
          uint32_t v[16];
          tmem_ld16(tmem_d + (uint32_t)cc, v);
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) sum[uu][jj] += __uint_as_float(v[jj]);
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[ab]);
    };
    int32_t next_chunk = 0;
    if (TS) {
      StageIter si{0, 0u, p.stages};
      const int b4 = (int)(b_bytes / 16);
      for (int32_t it = 0; it < KB; ++it, si.next()) {
        const int s = si.s;
        mbar_wait(&full_raw[s], si.ph);
        uint8_t* st = smem + (size_t)s * stage_bytes;
        const int32_t rows_ok = n - (kb_beg + it) * TW_KB;
        // split in place (hi) + lo plane; float4 i of a box belongs to reduction row (i % 256) / 8; rows >= n are zeroed
        float4* bh = reinterpret_cast<float4*>(st + a_span);
        float4* bl = reinterpret_cast<float4*>(st + a_span + b_bytes);
        for (int i = pt; i < b4; i += 32 * TW_EPI_WARPS) {
          float4 x = bh[i];
          if (((i & 255) >> 3) >= rows_ok) x = make_float4(0.f, 0.f, 0.f, 0.f);
          float4 h, l;
          split_tf32(x.x, h.x, l.x); split_tf32(x.y, h.y, l.y); split_tf32(x.z, h.z, l.z); split_tf32(x.w, h.w, l.w);
          bh[i] = h;
          bl[i] = l;
        }
        fence_proxy_async_smem();
        mbar_arrive(&full_conv[s]);
        if (it >= TW_CHUNK + 4 && ((it - 4) % TW_CHUNK) == 0) promote(next_chunk++);
      }
    }
    while (next_chunk < n_chunks) promote(next_chunk++);
    // partial tile -> global, coalesced through a shared-memory staging patch (the ring is idle by now: every K-block of
    // this CTA has been consumed by MMAs that completed before the last acc_full)
    const int64_t row0 = (int64_t)o0 + q * 32;
    float* obase = sg.out + (int64_t)blockIdx.z * sg.split_stride;
    const bool vec_ok = ((sg.n_cols & 3) == 0) && ((reinterpret_cast<uintptr_t>(obase) & 15) == 0);
    float* stage = reinterpret_cast<float*>(smem) + (warp - 6) * EPI16_PATCH;
#pragma unroll
    for (int uu = 0; uu < 4; ++uu) {
      const int32_t cc = 16 * (2 * uu + half);
      if (cc < p.BN && f0 + cc < sg.n_cols)
        epilogue_store16(stage, sum[uu], obase, sg.n_cols, row0, p.O, f0 + cc, sg.n_cols, vec_ok, lane);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, tmem_cols); }
}

struct TcWgradPlan {
  int32_t BN, nbox, tiles_per_seg, stages, splits;
  uint32_t smem_bytes;
  bool ts;
};
static int g_tc_wgrad_splits = 0;   // ngnn_set_tuning(8, s): force the number of reduction slices (0 = automatic)
static inline TcWgradPlan tc_wgrad_plan(int64_t n, int64_t F, int64_t O, int num_segs) {
  TcWgradPlan pl;
  // equal N tiles: F = 200 (products layer 1, [mean | root] side by side) is 2 x 112 columns, not 128 + 72 padded to 128
  pl.tiles_per_seg = (int32_t)ceil_div(F, (int64_t)TW_BN_MAX);
  pl.BN = round_up_i(ceil_div(F, (int64_t)pl.tiles_per_seg), 16);
  pl.nbox = (pl.BN + 31) / 32;
  pl.ts = g_tc_ts != 0;
  const uint32_t stage = (pl.ts ? 1u : 2u) * 4u * TW_BOX_BYTES + 2u * (uint32_t)pl.nbox * TW_BOX_BYTES;
  int st = (int)((TC_SMEM_LIMIT - 2048u) / stage);
  pl.stages = st > TC_MAX_STAGES ? TC_MAX_STAGES : st;
  pl.smem_bytes = (uint32_t)pl.stages * stage + 1024u + 256u;
  const uint32_t epi = (uint32_t)TW_EPI_WARPS * EPI16_PATCH * 4u + 1024u + 256u;   // the epilogue's staging reuses the ring
  if (pl.smem_bytes < epi) pl.smem_bytes = epi;
  const int64_t kblocks = ceil_div(n, TW_KB);
  const int64_t tiles = ceil_div(O, TC_BM) * pl.tiles_per_seg * num_segs;
  int64_t s = kNumSMs / tiles;                                   // one wave of CTAs (1 CTA / SM: smem-bound)
  const int64_t max_s = ceil_div(kblocks, 4);                    // at least 4 K-blocks per slice
  if (s > max_s) s = max_s;
  if (g_tc_wgrad_splits > 0) s = g_tc_wgrad_splits;
  if (s < 1) s = 1;
  pl.splits = (int32_t)s;
  return pl;
}
// partial planes for both segments
static inline size_t tc_wgrad_ws_bytes(int64_t n, int64_t F, int64_t O) {
  const TcWgradPlan pl = tc_wgrad_plan(n, F, O, 2);
  const TcWgradPlan pl1 = tc_wgrad_plan(n, F, O, 1);
  const int64_t s = pl.splits > pl1.splits ? pl.splits : pl1.splits;
  return 2 * align_up((size_t)s * O * F * sizeof(float), 256) + align_up((size_t)s * O * sizeof(float), 256) + 256;
}

// dw_l (+)= dy^T a_l ; dw_r (+)= dy^T a_r.  UNSUPPORTED when an operand is not TMA-addressable.
// n_dev (optional): device-side reduction length, n is then the capacity the launch is sized for.
// db (optional): the bias gradient, column sums of dy, folded into the same pass over dy (TS form); *db_done says whether it was.
static inline int32_t tc_gemm_wgrad(const float* dy, int64_t ld_dy, const float* a_l, int64_t ld_al, const float* a_r,
                                    int64_t ld_ar, int64_t n, const int32_t* n_dev, int64_t F, int64_t O, float* dw_l,
                                    float* dw_r, float* db, bool* db_done, int32_t accumulate, void* ws, size_t ws_bytes,
                                    cudaStream_t st) {
  if (db_done) *db_done = false;
  if (n < 1 || F < 1 || O < 1 || n >= (1LL << 31) - 256) return NGNN_E_UNSUPPORTED;
  if (!tma_addressable(dy, ld_dy)) return NGNN_E_UNSUPPORTED;
  if (dw_l && !tma_addressable(a_l, ld_al)) return NGNN_E_UNSUPPORTED;
  if (dw_r && !tma_addressable(a_r, ld_ar)) return NGNN_E_UNSUPPORTED;
  if (!dw_l && !dw_r) return NGNN_OK;
  if (ws == nullptr || ws_bytes < tc_wgrad_ws_bytes(n, F, O) || get_encode_fn() == nullptr) return NGNN_E_UNSUPPORTED;
  const int num_segs = (dw_l ? 1 : 0) + (dw_r ? 1 : 0);
  const TcWgradPlan pl = tc_wgrad_plan(n, F, O, num_segs);

  CUtensorMap tDY, tX1, tX2;
  const float* x1 = dw_l ? a_l : a_r;
  const int64_t ldx1 = dw_l ? ld_al : ld_ar;
  const CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  bool ok = make_tmap_2d(&tDY, dy, n, O, ld_dy, TW_KB, 32, sw);
  ok = ok && make_tmap_2d(&tX1, x1, n, F, ldx1, TW_KB, 32, sw);
  ok = ok && make_tmap_2d(&tX2, dw_r ? a_r : x1, n, F, dw_r ? ld_ar : ldx1, TW_KB, 32, sw);
  NGNN_REQUIRE(ok, NGNN_E_CUDA, "wgrad: cuTensorMapEncodeTiled failed");

  float* part0 = reinterpret_cast<float*>(align_up(reinterpret_cast<size_t>(ws), 256));
  float* part1 = part0 + align_up((size_t)pl.splits * O * F * sizeof(float), 256) / sizeof(float);
  float* part_db = part1 + align_up((size_t)pl.splits * O * F * sizeof(float), 256) / sizeof(float);
  const bool fuse_db = db != nullptr && pl.ts;
  const bool direct = pl.splits == 1 && !accumulate && !fuse_db;
  TcWgradParams p{};
  p.O = (int32_t)O; p.BN = pl.BN; p.nbox = pl.nbox; p.stages = pl.stages; p.tiles_per_seg = pl.tiles_per_seg;
  p.n = (int32_t)n; p.n_dev = n_dev; p.db_part = fuse_db ? part_db : nullptr;
  int ns = 0;
  float* outs[2] = {nullptr, nullptr};
  float* parts[2] = {nullptr, nullptr};
  if (dw_l) { p.seg[ns] = TcWgradSeg{direct ? dw_l : part0, O * F, (int32_t)F, 0}; outs[ns] = dw_l; parts[ns] = part0; ++ns; }
  if (dw_r) { p.seg[ns] = TcWgradSeg{direct ? dw_r : part1, O * F, (int32_t)F, dw_l ? 1 : 0}; outs[ns] = dw_r; parts[ns] = part1; ++ns; }

  static bool attr_set = false;
  if (!attr_set) {
    NGNN_CUDA(cudaFuncSetAttribute(k_tc_wgrad<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_LIMIT));
    NGNN_CUDA(cudaFuncSetAttribute(k_tc_wgrad<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_LIMIT));
    attr_set = true;
  }
  dim3 grid((unsigned)ceil_div(O, TC_BM), (unsigned)(pl.tiles_per_seg * ns), (unsigned)pl.splits);
  if (pl.ts) launch_chain(k_tc_wgrad<true>, grid, dim3(TW_THREADS), pl.smem_bytes, st, tDY, tX1, tX2, p);
  else launch_chain(k_tc_wgrad<false>, grid, dim3(TW_THREADS), pl.smem_bytes, st, tDY, tX1, tX2, p);
  NGNN_LAUNCH_CHECK();
  if (!direct) {
    // same per-element summation order as k_reduce_partials; both weight gradients and the bias gradient in one launch
    ReduceJobs jobs{};
    for (int j = 0; j < ns; ++j) jobs.job[jobs.n++] = ReduceJob{parts[j], outs[j], O * F, O * F};
    if (fuse_db) jobs.job[jobs.n++] = ReduceJob{part_db, db, O, O};
    dim3 rgrid((unsigned)ceil_div(O * F, 256), (unsigned)jobs.n);
    launch_chain(k_reduce_partials_multi, rgrid, dim3(256), 0, st, jobs, pl.splits, accumulate);
    NGNN_LAUNCH_CHECK();
    if (fuse_db && db_done) *db_done = true;
  }
  return NGNN_OK;
}

}  // namespace ngnn
