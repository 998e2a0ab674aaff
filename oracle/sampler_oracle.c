/*
 * sampler_oracle.c — TEST INFRASTRUCTURE ONLY (never linked or called by the product path).
 *
 * Sequential CPU restatement of the fan-out neighbour sampling that the reference obtains from
 * torch_geometric.loader.NeighborLoader -> pyg-lib `neighbor_sample` (third-party, un-vendored:
 * torch-geometric==2.5.1, reference docs/requirements.txt:11; pyg-lib wheel index torch-2.2.2+cu118,
 * reference docs/commands.txt:14).  Reference call sites: src/pipeline.py:75-83 (train loader:
 * input_nodes, num_neighbors, batch_size, shuffle; defaults replace=False, directed, disjoint=False)
 * and src/pipeline.py:152 (iteration).
 *
 * Published algorithm restated here:
 *   - the graph is CSC by destination (colptr, row);
 *   - hop h expands, in discovery order, every node first discovered at hop h-1 (hop 0: the seeds);
 *   - a node with deg <= fanout[h] takes all its in-neighbours in stored order, otherwise fanout[h]
 *     distinct positions drawn uniformly without replacement (Robert Floyd's subset algorithm, as
 *     in pyg-lib); with replace!=0 it takes fanout[h] independent uniform positions when deg>0;
 *   - every sampled neighbour gets a local id on FIRST SIGHT (seeds are 0..bs-1), and the edge
 *     (local_src = neighbour, local_dst = expanding node) is appended, so edges are grouped by
 *     destination in non-decreasing local-id order (CSR by construction);
 *   - nodes discovered in the last hop are not expanded.
 *
 * PARITY UNPINNED: pyg-lib is not installable here and the reference ships no golden vectors, so the
 * random stream cannot be matched to pyg-lib's (it draws from torch's global generator).  What IS
 * pinned: this file and the CUDA sampler share the SAME counter-based stream (Philox4x32-10, key =
 * seed, counter = (node, hop<<16 | draw/4, batch_idx, epoch), position = mulhi32(word, range)), so
 * the GPU block must equal this one bit for bit; validity (true neighbours, fan-out caps, no
 * duplicate positions) is asserted separately in tests/.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32); }

/* Philox4x32 with 10 rounds (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11). */
static void philox4x32_10(uint32_t ctr[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * ctr[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * ctr[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ ctr[1] ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ ctr[3] ^ k1;
    uint32_t n3 = (uint32_t)p0;
    ctr[0] = n0; ctr[1] = n1; ctr[2] = n2; ctr[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

/* Exposed so tests can pin the generator against published known-answer vectors. */
void ngnn_oracle_philox(const uint32_t ctr_in[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c[4] = {ctr_in[0], ctr_in[1], ctr_in[2], ctr_in[3]};
  philox4x32_10(c, key[0], key[1]);
  memcpy(out, c, sizeof(c));
}

static uint32_t draw_word(uint32_t v, uint32_t h, uint32_t j, uint32_t batch_idx, uint32_t epoch, uint64_t seed) {
  uint32_t c[4] = {v, (h << 16) | (j >> 2), batch_idx, epoch};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  return c[j & 3];
}

/*
 * local_of: caller-owned int32[N] scratch, all -1 on entry; restored to -1 on exit.
 * Returns 0, or -1 if a capacity would be exceeded, -2 on bad arguments.
 * Outputs as documented for ngnn_sample_block in include/ngnn_b200.h.
 */
int ngnn_oracle_sample_block(const int32_t* colptr, const int32_t* row, int64_t N, const int64_t* seeds, int32_t bs,
                             const int32_t* fanouts, int32_t H, int32_t replace, uint64_t seed, uint32_t epoch,
                             uint32_t batch_idx, int32_t* n_id, int32_t* rowptr, int32_t* col, int32_t* col_global,
                             int32_t* e_pos, int32_t* counts, int64_t cap_nodes, int64_t cap_edges,
                             int32_t* local_of) {
  if (!colptr || !row || !seeds || !fanouts || H < 1 || bs < 0 || N <= 0) return -2;
  int64_t n = 0, e = 0;
  int rc = 0;
  for (int32_t i = 0; i < bs; ++i) {
    if (n >= cap_nodes) { rc = -1; goto done; }
    n_id[n] = (int32_t)seeds[i];
    local_of[seeds[i]] = (int32_t)n;
    ++n;
  }
  counts[0] = bs;
  counts[H + 1] = 0;
  rowptr[0] = 0;
  {
    int64_t lo = 0, hi = bs;
    for (int32_t h = 0; h < H; ++h) {
      const int32_t fanout = fanouts[h];
      for (int64_t i = lo; i < hi; ++i) {
        const int32_t v = n_id[i];
        const int32_t beg = colptr[v], d = colptr[v + 1] - beg;
        int32_t k = replace ? (d > 0 ? fanout : 0) : (d < fanout ? d : fanout);
        int32_t* taken = (int32_t*)malloc(sizeof(int32_t) * (size_t)(k > 0 ? k : 1));
        if (e + k > cap_edges) { free(taken); rc = -1; goto done; }
        for (int32_t j = 0; j < k; ++j) {
          int32_t pos;
          if (!replace && d <= fanout) {
            pos = j;                                   /* take all, stored order */
          } else if (replace) {
            pos = (int32_t)mulhi32(draw_word((uint32_t)v, (uint32_t)h, (uint32_t)j, batch_idx, epoch, seed), (uint32_t)d);
          } else {                                     /* Floyd: jj = d-k+j, t = U[0,jj] */
            const int32_t jj = d - k + j;
            pos = (int32_t)mulhi32(draw_word((uint32_t)v, (uint32_t)h, (uint32_t)j, batch_idx, epoch, seed), (uint32_t)(jj + 1));
            for (int32_t q = 0; q < j; ++q) if (taken[q] == pos) { pos = jj; break; }
          }
          taken[j] = pos;
          const int32_t g = row[beg + pos];
          if (local_of[g] < 0) {                       /* first sight: next local id */
            if (n >= cap_nodes) { free(taken); rc = -1; goto done; }
            local_of[g] = (int32_t)n;
            n_id[n] = g;
            ++n;
          }
          col[e] = local_of[g];
          col_global[e] = g;
          if (e_pos) e_pos[e] = beg + pos;
          ++e;
        }
        free(taken);
        rowptr[i + 1] = (int32_t)e;
      }
      lo = hi;
      hi = n;
      counts[h + 1] = (int32_t)n;
      counts[H + 2 + h] = (int32_t)e;
    }
    for (int64_t i = lo; i < n; ++i) rowptr[i + 1] = (int32_t)e;   /* last-hop nodes: empty rows */
  }
done:
  for (int64_t i = 0; i < n; ++i) local_of[n_id[i]] = -1;
  return rc;
}
