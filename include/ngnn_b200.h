/*
 * ngnn_b200.h — C ABI of libngnn_b200.so (sm_100a only).
 *
 * This is the drop-in boundary for the GraphSAGE mini-batch hot path of
 * hhilsber/noise-GNN.  The reference has no native code: its hot path is two
 * third-party Python entry points,
 *     torch_geometric.nn.SAGEConv            (reference src/models/layers/sage.py:4,16-19,34,52)
 *     torch_geometric.loader.NeighborLoader  (reference src/pipeline.py:6,75-92,152)
 * and each function below names the piece of those two it replaces.
 *
 * Conventions (all functions):
 *   - extern "C", plain pointers and sizes; no C++/torch types cross the boundary.
 *   - every pointer is a DEVICE pointer into caller-owned memory unless the
 *     parameter is documented "(host)".  Nothing is retained after return.
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*), asynchronously;
 *     no internal synchronisation and no hidden allocation, so every call is
 *     CUDA-graph capturable.  Scratch memory comes from the caller through
 *     (ws, ws_bytes); the matching *_workspace_bytes() gives the size.
 *   - return value: NGNN_OK (0) or a negative NGNN_E_* code; the message of the
 *     last failure on the calling thread is read with ngnn_last_error().
 *   - index dtype is int32 at this ABI (the COO import/export take PyG's int64).
 *   - matrices are row-major fp32 with an explicit leading dimension (in elements).
 */
#ifndef NGNN_B200_H
#define NGNN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NGNN_OK             0
#define NGNN_E_INVALID     -1   /* bad shape / null pointer / bad flag            */
#define NGNN_E_ALIGN       -2   /* pointer or leading dimension not aligned       */
#define NGNN_E_CUDA        -3   /* a CUDA runtime call failed                     */
#define NGNN_E_WORKSPACE   -4   /* workspace too small                            */
#define NGNN_E_UNSUPPORTED -5   /* valid request outside what the kernels cover   */

#define NGNN_ACT_NONE 0
#define NGNN_ACT_RELU 1

typedef void* ngnn_stream_t;     /* cudaStream_t */

/* ---- library ---------------------------------------------------------- */
int32_t ngnn_version(void);                              /* 10000*major+100*minor+patch */
int32_t ngnn_last_error(char* buf, size_t cap);          /* copies a NUL-terminated message */
/* 1 if the current device is compute capability 10.x (the only one the cubin runs on). */
int32_t ngnn_device_supported(void);
/* Total number of device kernels this library has launched in the process (monotonic); bench.py
 * differences it around the timed region to report gpu_launches.                               */
uint64_t ngnn_launch_count(void);

/* Per-step control words in DEVICE memory: what changes between two replays of a captured step.  ngnn_step_ctl_set is a
 * one-thread kernel whose values travel as launch arguments (no host buffer has to outlive the call); the sampler reads
 * (epoch, batch_idx) from it, the K-GEMM epilogues the dropout stream offset (+ layer index), the loss its weight. */
typedef struct {
  uint32_t epoch, batch_idx;
  uint32_t drop_offset_lo, drop_offset_hi;
  float    loss_scale;    /* weight of this batch's loss gradient: 1 on one GPU; data parallel: bs_r * R / sum_r bs_r, so that the
                             all-reduced mean is the GLOBAL-batch mean; 0 masks a padded (wrapped-around) batch out entirely */
  uint32_t ticket;        /* library-internal arrival counter (zeroed by ngnn_step_ctl_set) */
  uint32_t reserved[2];
} ngnn_step_ctl_t;
int32_t ngnn_step_ctl_set(ngnn_step_ctl_t* ctl /*device*/, uint32_t epoch, uint32_t batch_idx, uint64_t drop_offset,
                          float loss_scale, ngnn_stream_t stream);

/* ---- block structure (replaces PyG's COO->CSC conversion) ---- */
/* Stable sort of a COO edge list by destination:
 *   perm = argsort_stable(dst); col = src[perm]; rowptr[i] = #edges with dst < i.
 * Bit-exact target: oracle/structure.py::coo_to_csr.                                  */
size_t  ngnn_coo_to_csr_workspace_bytes(int64_t e, int64_t n_rows);
int32_t ngnn_coo_to_csr(const int64_t* src, const int64_t* dst, int64_t e, int64_t n_rows,
                        int32_t* rowptr /*[n_rows+1]*/, int32_t* col /*[e]*/, int32_t* perm /*[e]*/,
                        void* ws, size_t ws_bytes, ngnn_stream_t stream);

/* Transpose of a CSR block (CSC by source) for the atomic-free backward:
 *   perm_t = argsort_stable(col); row_t = dst_of_edge[perm_t]; colptr_t over n_cols.
 * Only the first e_limit edges (a hop prefix) take part.                               */
size_t  ngnn_csr_transpose_workspace_bytes(int64_t e, int64_t n_cols);
int32_t ngnn_csr_transpose(const int32_t* rowptr, const int32_t* col, int64_t n_rows, int64_t e_limit,
                           int64_t n_cols, int32_t* colptr_t /*[n_cols+1]*/, int32_t* row_t /*[e_limit]*/,
                           int32_t* perm_t /*[e_limit]*/, void* ws, size_t ws_bytes, ngnn_stream_t stream);

/* CSR -> PyG edge_index int64 [2,e] (row 0 = source, row 1 = destination). */
int32_t ngnn_csr_to_coo(const int32_t* rowptr, const int32_t* col, int64_t n_rows, int64_t e,
                        int64_t* edge_index /*[2*e]*/, ngnn_stream_t stream);

/* out[i,:] = table[idx[i],:]  (NeighborLoader's x[n_id] slice). */
int32_t ngnn_gather_rows(const float* table, int64_t ld_table, const int32_t* idx, int64_t n, int64_t F,
                         float* out, int64_t ld_out, ngnn_stream_t stream);

/* ---- K-AGG: CSR segment mean (SAGEConv's MeanAggregation) ----
 *   mean[i,:] = (1/max(deg_i,1)) * sum_{p in [rowptr[i],rowptr[i+1])} x[col[p],:]   for i < n_dst
 * Optional fused root gather: if root_idx != NULL, root[i,:] = x[root_idx[i],:].
 * x is any row-major table (a block's x[n,F] or the resident feature table with
 * global ids in col).  No atomics: one warp (or sub-warp) owns one destination row. */
int32_t ngnn_sage_agg_fwd(const int32_t* rowptr, const int32_t* col, const float* x, int64_t ld_x,
                          int64_t n_dst, int64_t F, float* mean, int64_t ld_mean,
                          const int32_t* root_idx, float* root, int64_t ld_root,
                          ngnn_stream_t stream);

/* ---- GCNConv(normalize=False) aggregation (reference src/models/layers/convolution.py:19-23 -> PyG GCNConv: sum over
 * in-neighbours of the already projected rows, then + bias; no self loops, duplicates counted) ----
 *   out[i,:] = sum_{p in [rowptr[i],rowptr[i+1])} z[col[p],:] + bias      (bias may be NULL)
 * Same kernel family as K-AGG without the 1/deg scale.  The backward is ngnn_sage_agg_bwd on the transposed block. */
int32_t ngnn_gcn_agg_fwd(const int32_t* rowptr, const int32_t* col, const float* z, int64_t ld_z,
                         int64_t n_dst, int64_t O, const float* bias, float* out, int64_t ld_out,
                         ngnn_stream_t stream);

/* Development knobs for kernel sweeps and A/B measurements (profiles/prof_agg.py, profiles/prof_gemm.py, NGNN_TUNING in
 * bench.py).  key 0: neighbour rows in flight per lane of K-AGG for F <= 128 (0 = default); 1: CTA size of the generic
 * aggregation kernel (128/256/512); 2: lanes per row for 64 < F <= 128 (32/16/8); 3: software-pipelined persistent K-AGG
 * (1 = default); 4: widest N tile of the tcgen05 GEMM (128 = default, 256); 6: A operand of the tcgen05 kernels in
 * tensor memory (1 = default) or shared memory (0); 7: L2 evict_last priority on the layer-1 table gathers (1 = default);
 * 8: number of reduction slices of the tcgen05 K-WGRAD (0 = automatic: one wave of CTAs); 9: K-AGG-T variant for
 * 128 < F <= 256 (0 = half-warp per row, default; 1 / 2 = warp per row, unroll 2 / 4); 10: programmatic dependent launch
 * along the step's kernel chain (0 = off, default: measured slower). */
int32_t ngnn_set_tuning(int32_t key, int32_t value);

/* ---- K-AGG-T: transpose (CSC) segment sum, backward of K-AGG ----
 *   dx[j,:] = gate_j * ( sum_{q in [colptr_t[j],colptr_t[j+1])} dmean_scaled[row_t[q],:]
 *                        + (j < n_root ? dx_root[j,:] : 0) )                         for j < n_src
 * dmean_scaled already carries the 1/max(deg,1) factor (applied by ngnn_sage_dgrad).
 * gate: if act_ref != NULL, gate_j[f] = act_ref[j,f] > 0 ? act_scale : 0  (ReLU+dropout
 * backward of the producing layer, folded in); else 1.                                   */
int32_t ngnn_sage_agg_bwd(const int32_t* colptr_t, const int32_t* row_t, const float* dmean_scaled,
                          int64_t ld_dmean, int64_t n_src, int64_t F,
                          const float* dx_root, int64_t ld_root, int64_t n_root,
                          const float* act_ref, int64_t ld_act, float act_scale,
                          float* dx, int64_t ld_dx, ngnn_stream_t stream);

/* ---- K-GEMM: fused projection (lin_l + lin_r + bias [+ReLU +dropout]) ----
 *   out[i,o] = drop( act( sum_f a_l[i,f]*w_l[o,f] + sum_f a_r[i,f]*w_r[o,f] + bias[o] ) )
 * a_l (the mean) or a_r (the root rows) may be NULL to drop that term; bias may be NULL.
 * w_l, w_r are PyG-layout weights [O,F] (ld = F).  Dropout: inverted, keep-mask from a
 * counter-based Philox stream keyed (seed, offset, element index); drop_p = 0 disables.
 * Tensor-core path (tcgen05, 3xTF32 split => fp32-grade accuracy) when F%4==0 and the
 * operands are 16-byte aligned, SIMT fp32 path otherwise; `path` (host, may be NULL)
 * receives 1 for the tcgen05 path, 0 for SIMT.                                          */
size_t  ngnn_sage_gemm_workspace_bytes(int64_t F, int64_t O);   /* split-weight planes of the tcgen05 path */
int32_t ngnn_sage_gemm_fwd(const float* a_l, int64_t ld_al, const float* a_r, int64_t ld_ar,
                           const float* w_l, const float* w_r, const float* bias,
                           int64_t n, int64_t F, int64_t O, int32_t act,
                           float drop_p, uint64_t seed, uint64_t offset,
                           float* out, int64_t ld_out, int32_t* path,
                           void* ws, size_t ws_bytes, ngnn_stream_t stream);
/* Development aid: when set to a device buffer of >= 8 + 8*K-blocks int64, CTA (0,0) of the tcgen05 GEMM records
 * clock64() per pipeline event (profiles/trace_gemm.py).  NULL disables.                                      */
int32_t ngnn_debug_set_trace(void* device_buffer);
/* 0 = automatic dispatch (default), 1 = force the SIMT fp32 kernels (tests / A-B timing). */
int32_t ngnn_set_gemm_path(int32_t mode);

/* ---- K-DGRAD: data gradients of the projection ----
 *   dmean_scaled[i,f] = (1/max(deg_i,1)) * sum_o dy[i,o]*w_l[o,f]     (deg from rowptr; rowptr NULL => 1)
 *   dx_root[i,f]      =                   sum_o dy[i,o]*w_r[o,f]
 * Either output may be NULL.                                                              */
size_t  ngnn_sage_dgrad_workspace_bytes(int64_t F, int64_t O);
int32_t ngnn_sage_dgrad(const float* dy, int64_t ld_dy, const float* w_l, const float* w_r,
                        const int32_t* rowptr, int64_t n, int64_t F, int64_t O,
                        float* dmean_scaled, int64_t ld_dmean, float* dx_root, int64_t ld_root,
                        void* ws, size_t ws_bytes, ngnn_stream_t stream);

/* ---- K-WGRAD: weight / bias gradients ----
 *   dw_l[o,f] (+)= sum_i dy[i,o]*a_l[i,f];  dw_r[o,f] (+)= sum_i dy[i,o]*a_r[i,f];  db[o] (+)= sum_i dy[i,o]
 * Split over the long n dimension into fixed slices reduced in a fixed order
 * (deterministic, atomic-free).  accumulate != 0 adds into the outputs.                  */
size_t  ngnn_sage_wgrad_workspace_bytes(int64_t n, int64_t F, int64_t O);
int32_t ngnn_sage_wgrad(const float* dy, int64_t ld_dy, const float* a_l, int64_t ld_al,
                        const float* a_r, int64_t ld_ar, int64_t n, int64_t F, int64_t O,
                        float* dw_l, float* dw_r, float* db, int32_t accumulate,
                        void* ws, size_t ws_bytes, ngnn_stream_t stream);

/* Elementwise backward of ReLU + inverted dropout given the saved post-activation output:
 *   dz = dh * (h > 0 ? scale : 0)          (scale = 1/(1-p); dropped and negative entries have h == 0) */
int32_t ngnn_act_bwd(const float* dh, int64_t ld_dh, const float* h, int64_t ld_h, int64_t n, int64_t O,
                     float scale, float* dz, int64_t ld_dz, ngnn_stream_t stream);

/* ---- loss (F.cross_entropy on the seed rows, reference src/pipeline.py:155-165) ----
 * Mean softmax cross-entropy over bs rows against `target`; a warp-per-row launch plus a fixed-order reduce produce
 *   stats[0] += mean loss, stats[1] += #(argmax == y_true)  (y_true NULL => skipped)
 *   dlogits[i,c] = (softmax_ic - [c==target_i]) * grad_scale / bs                          */
int32_t ngnn_ce_fwd_bwd(const float* logits, int64_t ld, const int64_t* target, const int64_t* y_true,
                        int64_t bs, int64_t C, float grad_scale, float* stats /*[2]*/,
                        float* dlogits, int64_t ld_d, float* row_scratch /*[2*bs]*/, ngnn_stream_t stream);

/* Same, with the labels gathered by id: target[row_ids[i]], y_true[row_ids[i]] (row_ids = the block's n_id, so the
 * loader does not have to materialise y[n_id] / yhn[n_id]).  row_ids NULL = identity.                          */
int32_t ngnn_ce_fwd_bwd_gather(const float* logits, int64_t ld, const int64_t* target, const int64_t* y_true,
                               const int32_t* row_ids, int64_t bs, int64_t C, float grad_scale, float* stats /*[2]*/,
                               float* dlogits, int64_t ld_d, float* row_scratch /*[2*bs]*/, ngnn_stream_t stream);

/* ---- optimizer (torch.optim.Adam(lr), reference src/models/model.py:67-69) ----
 * One fused pass over a flat parameter bucket; step_count is the 1-based step t held in a
 * device int64 (so a captured graph can be replayed): the kernel reads *step_dev, and the
 * a follow-up one-thread launch increments it when advance_step == 1; advance_step == 2: step_dev
 * points at TWO int64 words, the second a zero-initialised ticket, and the last CTA of the Adam
 * kernel itself advances the counter (one launch).                                              */
int32_t ngnn_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                       float lr, float beta1, float beta2, float eps, float weight_decay,
                       float grad_scale, int64_t* step_dev, int32_t advance_step, ngnn_stream_t stream);

/* ---- K-SAMPLE / K-RELABEL / K-CSR: fan-out neighbour sampler ----
 * Graph = CSC by destination: colptr[N+1], row[E] (in-neighbours in stored order).
 * For hop h = 0..H-1 every node first discovered at hop h-1 (hop 0: the seeds), in
 * discovery order, draws min(deg,fanout[h]) distinct in-neighbour positions (all of them,
 * in stored order, when deg <= fanout[h]; fanout[h] draws with replacement when replace!=0
 * and deg>0).  Draws come from Philox4x32-10 keyed (seed, epoch) with counter
 * (node global id, hop, batch_idx, draw/4), so a block depends only on
 * (seed, epoch, batch_idx, seeds) — not on the GPU count or launch geometry.
 * Newly seen neighbours get local ids in first-seen order (seeds are 0..bs-1).
 * Outputs (capacity = the worst case given by ngnn_sample_capacity):
 *   n_id[n]      global id of local node i
 *   rowptr[n+1]  CSR by destination over ALL n local nodes (last-hop nodes have empty rows)
 *   col[e]       local source id per edge;  col_global[e] the same in global ids
 *   e_pos[e]     position of the sampled edge in row[] (maps to PyG's e_id through the CSC perm)
 *   counts[2*(H+1)] device int32: counts[h] = #nodes after hop h-1 (counts[0]=bs, counts[H]=n),
 *                counts[H+1+h] = #edges after hop h-1 (counts[H+1]=0, counts[2H+1]=e)
 * The workspace holds two N-sized maps that must be initialised ONCE with
 * ngnn_sample_workspace_init and are restored by every call.  Seeds must be distinct.
 * Bit-exact target: oracle/sampler_oracle.c (same Philox, sequential).                      */
int32_t ngnn_sample_capacity(int32_t bs, const int32_t* fanouts /*(host)[H]*/, int32_t H, int64_t N,
                             int64_t* max_nodes, int64_t* max_edges);
size_t  ngnn_sample_workspace_bytes(int64_t N, int32_t bs, const int32_t* fanouts /*(host)*/, int32_t H);
int32_t ngnn_sample_workspace_init(void* ws, size_t ws_bytes, int64_t N, ngnn_stream_t stream);
int32_t ngnn_sample_block(const int32_t* colptr, const int32_t* row, int64_t N,
                          const int64_t* seeds, int32_t bs, const int32_t* fanouts /*(host)[H]*/, int32_t H,
                          int32_t replace, uint64_t seed, uint32_t epoch, uint32_t batch_idx,
                          int32_t* n_id, int32_t* rowptr, int32_t* col, int32_t* col_global, int32_t* e_pos,
                          int32_t* counts, void* ws, size_t ws_bytes, ngnn_stream_t stream);

/* The same, plus what a replayed (CUDA-graph) step needs:
 *   ctl (device, optional): when non-NULL the RNG key (epoch, batch_idx) is read from it on the device instead of from the
 *       arguments, so one captured launch sequence samples a different block at every replay;
 *   num_transposes = T (0..min(H,4)), colptr_t / row_t (host arrays of T device pointers): the CSC transposes of the hop
 *       prefixes b = 1..T — the first counts[H+1+b] edges over the counts[b] local nodes they touch — which the backward's
 *       atomic-free transpose segment-sum reads (layer l of L uses b = min(L-l+1, H)).  colptr_t[b-1] has capacity
 *       (cumulative nodes after hop b-1) + 1, row_t[b-1] capacity (cumulative edges after hop b-1); every row lists its
 *       destinations in ascending order (= ascending edge position: the stable order of ngnn_csr_transpose).
 * 5 + 2H kernel launches, none of them a library primitive, no host synchronisation.                              */
int32_t ngnn_sample_block_ex(const int32_t* colptr, const int32_t* row, int64_t N,
                             const int64_t* seeds, int32_t bs, const int32_t* fanouts /*(host)[H]*/, int32_t H,
                             int32_t replace, uint64_t seed, uint32_t epoch, uint32_t batch_idx,
                             const ngnn_step_ctl_t* ctl /*device, optional*/,
                             int32_t* n_id, int32_t* rowptr, int32_t* col, int32_t* col_global, int32_t* e_pos,
                             int32_t* edge_dst /*[e] local destination id per edge; may be NULL when T == 0*/,
                             int32_t* counts, int32_t num_transposes, int32_t* const* colptr_t /*(host)[T]*/,
                             int32_t* const* row_t /*(host)[T]*/, void* ws, size_t ws_bytes, ngnn_stream_t stream);

/* Table rows of a sampled block for a feature table stored in another row order (remap[global id] = table row, int32 [N]):
 * col_table[p] = remap[col_global[p]] for p < e, n_table[i] = remap[n_id[i]] for i < n, with n / e read from the
 * sampler's device-side `counts` (no host round trip; grids sized by max_nodes / max_edges).                     */
int32_t ngnn_block_table_index(const int32_t* remap, const int32_t* col_global, const int32_t* n_id,
                               const int32_t* counts, int32_t H, int64_t max_nodes, int64_t max_edges,
                               int32_t* col_table, int32_t* n_table, ngnn_stream_t stream);

/* ---- the whole step behind one call (loop body of PipelineCO.train, reference src/pipeline.py:152-168) ----
 * SAGE network of reference src/models/layers/sage.py:7-40: num_layers SAGEConv layers
 * (in_dim -> hidden_dim -> ... -> out_dim), ReLU + dropout between layers.  Parameters and gradients live in flat
 * fp32 buckets laid out per layer as [lin_l.weight (O*F) | lin_l.bias (O) | lin_r.weight (O*F)] — the order of
 * torch's module.parameters() for the reference module, so state_dicts map one to one.                             */
typedef struct {
  int32_t num_layers;
  int32_t in_dim, hidden_dim, out_dim;
  float   dropout;      /* applied between layers when training != 0 */
  int32_t training;
} ngnn_sage_model_t;

/* A sampled block as produced by ngnn_sample_block (device arrays) with its per-hop extents on the host. */
typedef struct {
  const int32_t* rowptr;      /* [n+1]  */
  const int32_t* col;         /* [e] local source ids  */
  const int32_t* col_global;  /* [e] global source ids */
  const int32_t* n_id;        /* [n] global id of local node */
  int32_t        num_hops;    /* H */
  const int32_t* hop_nodes;   /* (host) [H+1] cumulative nodes, hop_nodes[0] = batch size */
  const int32_t* hop_edges;   /* (host) [H+1] cumulative edges, hop_edges[0] = 0 */
  /* Optional precomputed transposes of hop prefixes (ngnn_csr_transpose of the first hop_edges[b] edges over
   * hop_nodes[b] columns), indexed by b = 1..H; NULL entries are computed inside ngnn_sage_step.  The loader
   * builds them on its side stream so the sort is off the step's critical path.                           */
  const int32_t* colptr_t[8];
  const int32_t* row_t[8];
  /* Optional feature-table addressing (ngnn_block_table_index): when the resident table is stored hot rows first,
   * col_table[p] / n_table[i] are the TABLE rows of col_global[p] / n_id[i] and rows < hot_rows are the hot set
   * (L2 evict_last priority; the others stream through with evict_first).  NULL / 0: the table is indexed by node id. */
  const int32_t* col_table;
  const int32_t* n_table;
  int64_t        hot_rows;
  /* Device-side extents (optional).  counts = the sampler's device array [2*(H+1)] (ngnn_sample_block): when non-NULL,
   * hop_nodes / hop_edges may be NULL, the step reads every extent on the device and sizes its launches for the
   * declared capacities (max_hop_nodes / max_hop_edges), so the launch sequence does not depend on the block and the
   * step can be captured in a CUDA graph and replayed on the next block written into the same buffers.  Requires
   * batch_size (host: number of seeds = counts[0]) and, for training, the transposes colptr_t / row_t built by
   * ngnn_sample_block_ex.  ctl (device, optional): per-replay control words; the dropout stream offset is then
   * ctl->drop_offset + layer index instead of the drop_offset argument.                                         */
  const int32_t* counts;
  int32_t        batch_size;
  const ngnn_step_ctl_t* ctl;
  /* 0: the step runs layer 1's aggregation itself; 1 / 2: the caller already ran ngnn_sage_agg1 for this block into copy
   * 0 / 1 of the arena's layer-1 buffers (it depends on the block and the table only, so it can run for the NEXT block
   * beside the current block's step: an HBM-bound gather under tensor-bound GEMMs).                              */
  int32_t        agg1_buffer;
  /* != 0: ngnn_sage_prep_weights already split this step's parameters into the arena (it only depends on the parameters, so
   * a caller can run it on another stream beside ngnn_sage_agg1); the step then skips its own weight-pack launch.   */
  int32_t        weights_prepared;
} ngnn_block_t;

int64_t ngnn_sage_num_params(const ngnn_sage_model_t* model);
/* Arena size for blocks up to the given cumulative worst-case extents (host arrays [H+1], e.g. from
 * ngnn_sample_capacity per hop).                                                                      */
size_t  ngnn_sage_step_workspace_bytes(const ngnn_sage_model_t* model, int32_t num_hops,
                                       const int64_t* max_hop_nodes, const int64_t* max_hop_edges);
/* Forward (+ loss + backward) of one block.  Layer l of L computes only the rows within L-l hops of the seeds
 * (exact for the seed rows); layer 1 reads the resident feature table `table` by global id.
 *   grads  != NULL : training step — gradients of the mean CE over the seed rows are WRITTEN (not accumulated) to
 *                    `grads`; dropout is active iff model->training.  Requires target_global.
 *   grads  == NULL : forward (+ loss when target_global != NULL) only.
 *   target_global / label_global: int64 [N] label arrays indexed by GLOBAL node id (yhn / y of the reference);
 *                    stats[0] += mean loss, stats[1] += #(argmax == label) (label_global NULL => skipped).
 *   logits_out (optional): [bs, out_dim] seed-row logits.
 * No host synchronisation, no allocation; scratch = (ws, ws_bytes) from ngnn_sage_step_workspace_bytes.          */
int32_t ngnn_sage_step(const ngnn_sage_model_t* model, const float* params, float* grads, const ngnn_block_t* block,
                       const int64_t* max_hop_nodes, const int64_t* max_hop_edges,
                       const float* table, int64_t ld_table,
                       const int64_t* target_global, const int64_t* label_global,
                       uint64_t drop_seed, uint64_t drop_offset, float* stats /*[2]*/,
                       float* logits_out, int64_t ld_logits, void* ws, size_t ws_bytes, ngnn_stream_t stream);

/* Layer 1's aggregation alone (mean of the sampled in-neighbours + root gather from the resident table) into copy `buffer`
 * (0 / 1) of the arena's layer-1 buffers; consumed by a later ngnn_sage_step / _forward / _backward on the same arena whose
 * block carries agg1_buffer = buffer + 1.                                                                          */
int32_t ngnn_sage_agg1(const ngnn_sage_model_t* model, const ngnn_block_t* block, const int64_t* max_hop_nodes,
                       const int64_t* max_hop_edges, const float* table, int64_t ld_table, int32_t buffer,
                       void* ws, size_t ws_bytes, ngnn_stream_t stream);

/* The step's weight pack alone: hi / lo (3xTF32) K-major planes of every layer's [W_l | W_r] and, for the data gradients,
 * [W_l^T ; W_r^T], written into the arena.  One launch; depends on the parameters only.                              */
int32_t ngnn_sage_prep_weights(const ngnn_sage_model_t* model, const float* params, int32_t num_hops,
                               const int64_t* max_hop_nodes, const int64_t* max_hop_edges, void* ws, size_t ws_bytes,
                               ngnn_stream_t stream);

/* The step in two calls, for losses that couple several networks (co-teaching, reference src/pipeline.py:95-142: two
 * forwards, one joint loss, two backwards).  ngnn_sage_forward = the training-mode forward of ngnn_sage_step (dropout
 * iff model->training), seed-row logits to logits_out, activations kept in ws; ngnn_sage_backward = its backward from
 * a caller-supplied top-layer gradient dlogits [bs, out_dim], gradients WRITTEN to grads.  Same ws, block, params and
 * model for the pair; nothing else may use that ws in between.                                                     */
int32_t ngnn_sage_forward(const ngnn_sage_model_t* model, const float* params, const ngnn_block_t* block,
                          const int64_t* max_hop_nodes, const int64_t* max_hop_edges,
                          const float* table, int64_t ld_table, uint64_t drop_seed, uint64_t drop_offset,
                          float* logits_out, int64_t ld_logits, void* ws, size_t ws_bytes, ngnn_stream_t stream);
int32_t ngnn_sage_backward(const ngnn_sage_model_t* model, const float* params, float* grads, const ngnn_block_t* block,
                           const int64_t* max_hop_nodes, const int64_t* max_hop_edges,
                           const float* table, int64_t ld_table, const float* dlogits, int64_t ld_dlogits,
                           void* ws, size_t ws_bytes, ngnn_stream_t stream);

/* ---- co-teaching loss (reference src/utils/losses.py:10-49, CTLoss.forward) without its two host argsorts ----
 * Per-sample CE of two networks on the same seed rows; network 1 is trained on the num_remember rows with the smallest
 * loss under network 2 and vice versa (ties broken by row index = a stable argsort).
 *   target / y_true / clean_mask: label arrays indexed by row_ids[i] when row_ids != NULL (global node ids), else by i;
 *       clean_mask (uint8, optional) = the reference's noise_or_not, for the pure ratios.
 *   stats[0..5] += loss_1, loss_2 (means over the selected rows), #correct_1, #correct_2, pure_ratio_1, pure_ratio_2
 *   dlogits1/2 (optional): gradient of loss_1 / loss_2 w.r.t. logits1 / logits2;   order1/2 (optional, [bs]): row
 *   indices in ascending loss order (order[:num_remember] = the reference's ind_update, the rest = ind_noisy).
 *   scratch: 12*bs floats.  num_remember = int((1 - forget_rate) * bs), computed by the caller like the reference;
 *   num_remember == 0 gives zero losses and gradients (the reference divides by zero there).                       */
int32_t ngnn_ct_loss(const float* logits1, int64_t ld1, const float* logits2, int64_t ld2, const int64_t* target,
                     const int64_t* y_true, const int32_t* row_ids, const uint8_t* clean_mask, int64_t bs, int64_t C,
                     int64_t num_remember, float* stats /*[6]*/, float* dlogits1, int64_t ldd1, float* dlogits2,
                     int64_t ldd2, int32_t* order1, int32_t* order2, float* scratch, ngnn_stream_t stream);

/* ---- SAGEPL extras (reference src/models/layers/sagePL.py:41-49, src/utils/augmentation.py:88-102) ----
 * adding_noise, fused: out[i,:] = x[i,:] + s_i * rate * v_i / max(|v_i|_2, 1e-12),  v_i = noise[idx[i],:]
 *   idx = the block's n_id (reference line 47: s_i = 1) or NULL = identity with use_sign (line 44: s_i = sign(x[i,:])).
 * Backward w.r.t. the noise table: dnoise[idx[i],:] += rate/|v_i| * (g - u_i (u_i . g)),  g = s_i * dout[i,:], u_i = v_i/|v_i|
 *   (added with atomics: a block's n_id are distinct, so every address is touched once and the result is deterministic;
 *   the caller zeroes dnoise).  The gradient w.r.t. x is dout itself.                                               */
int32_t ngnn_noise_add_fwd(const float* x, int64_t ld_x, const float* noise, int64_t ld_noise, const int32_t* idx,
                           int64_t n, int64_t F, float rate, int32_t use_sign, float* out, int64_t ld_out,
                           ngnn_stream_t stream);
int32_t ngnn_noise_add_bwd(const float* dout, int64_t ld_d, const float* x, int64_t ld_x, const float* noise,
                           int64_t ld_noise, const int32_t* idx, int64_t n, int64_t F, float rate, int32_t use_sign,
                           float* dnoise, int64_t ld_dn, ngnn_stream_t stream);
/* shuffle_pos: out = x with, per row, k distinct random positions' values permuted among themselves
 *   pos = Robert Floyd subset of [0,F) (insertion order), sel = Fisher-Yates shuffle of pos, out[row,pos[j]] = x[row,sel[j]].
 * Philox4x32-10 keyed (seed) with counter (row, draw/4, offset): bit-exact target oracle/sagepl_oracle.py::shuffle_rows.
 * The reference draws from torch.randperm, so only the LAW is comparable (validity: each row is a permutation of itself
 * that moves at most k positions).  F <= 2048; x and out must not alias.                                            */
int32_t ngnn_shuffle_rows(const float* x, int64_t ld_x, int64_t n, int64_t F, int32_t k, uint64_t seed, uint64_t offset,
                          float* out, int64_t ld_out, ngnn_stream_t stream);

/* 1 (default): inside ngnn_sage_step the weight gradients of layers >= 2 run on an internal auxiliary stream, forked
 * from / joined back into the caller's stream with events (they are off the backward's critical path); 0: strictly
 * one stream.                                                                                                  */
int32_t ngnn_set_step_overlap(int32_t on);

/* In-situ kernel timing for the roofline report: after ngnn_probe_enable(k), the next k ngnn_sage_step calls record
 * a CUDA-event pair on their stream around the layer-1 K-AGG launch; ngnn_probe_read waits for them and returns the
 * per-launch durations in milliseconds.  ngnn_probe_enable(0) disables and frees the events.                         */
int32_t ngnn_probe_enable(int32_t max_samples);
int32_t ngnn_probe_read(float* ms /*(host)[cap]*/, int32_t cap, int32_t* n /*(host)*/);
/* The same steps' second event pair: around the K-AGG-T launch that writes the gradient of layer 1's output rows (the widest
 * transpose-sum of the step; networks with one layer have none). */
int32_t ngnn_probe_read_agg_t(float* ms /*(host)[cap]*/, int32_t cap, int32_t* n /*(host)*/);
/* The same launches by the device's own clock: last CTA end - first CTA start (%globaltimer), i.e. without the two event
 * records and the launch latency an event pair includes.  Synchronises the device.                               */
int32_t ngnn_probe_read_device_clock(float* ms /*(host)[cap]*/, int32_t cap, int32_t* n /*(host)*/);

#ifdef __cplusplus
}
#endif
#endif /* NGNN_B200_H */
