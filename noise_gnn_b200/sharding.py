"""Host-side seed sharding for data-parallel mini-batch training (pure CPU logic, no CUDA).

The reference is single-process (src/main.py:76-83); its loader draws ``ceil(len(input_nodes)/batch_size)``
shuffled batches per epoch (src/pipeline.py:75-83, 152).  For N GPUs the same global batch sequence is dealt
round-robin: rank r of R trains on global batches r, r+R, r+2R, ...  The epoch order is a pure function of
(seed, epoch) — identical on every rank with no communication — and the last round wraps around so every rank
issues the same number of steps (collectives stay matched); the wrapped batches are weighted 0 (``loss_scale``), and a
short last batch is weighted by its size, so the all-reduced gradient is the mean over the round's real seeds.
"""
from __future__ import annotations

import math

import torch


class SeedSharder:
    def __init__(self, input_nodes: torch.Tensor, batch_size: int, shuffle: bool, seed: int = 1232, rank: int = 0,
                 world_size: int = 1, drop_last: bool = False):
        if not (0 <= rank < world_size):
            raise ValueError(f"rank {rank} outside world of {world_size}")
        self.input_nodes = input_nodes.to(torch.int64).view(-1).cpu()
        self.batch_size, self.shuffle, self.seed = int(batch_size), bool(shuffle), int(seed)
        self.rank, self.world_size, self.drop_last = int(rank), int(world_size), bool(drop_last)

    @property
    def num_batches_global(self) -> int:
        n = len(self.input_nodes)
        return n // self.batch_size if self.drop_last else math.ceil(n / self.batch_size)

    def __len__(self) -> int:
        """Steps per epoch on this rank (equal on all ranks)."""
        return math.ceil(self.num_batches_global / self.world_size)

    def epoch_permutation(self, epoch: int) -> torch.Tensor:
        """Rank-agnostic seed order for an epoch: a pure function of (seed, epoch)."""
        if not self.shuffle:
            return self.input_nodes
        g = torch.Generator(device="cpu")
        g.manual_seed((self.seed * 1000003 + epoch) & 0x7FFFFFFFFFFFFFFF)
        return self.input_nodes[torch.randperm(len(self.input_nodes), generator=g)]

    def global_batch_index(self, step: int) -> int:
        """Global batch trained by this rank at local step `step` (wraps around in the last, padded round)."""
        return (step * self.world_size + self.rank) % max(self.num_batches_global, 1)

    @property
    def full_len(self) -> int:
        """Seeds of a full batch: ``batch_size``, or every seed when there are fewer (the reference's full-batch configs:
        pubmed's 60 train seeds, cora's 140, with batch_size 512)."""
        return min(self.batch_size, len(self.input_nodes))

    def batch_len(self, global_batch_idx: int) -> int:
        """Number of seeds of a global batch (only the epoch's last batch can be short)."""
        n = len(self.input_nodes)
        return max(min(self.batch_size, n - global_batch_idx * self.batch_size), 0)

    def is_padded(self, step: int) -> bool:
        """True when this rank's batch at local step `step` is a wrap-around filler of the epoch's last, incomplete round: it
        keeps the collectives matched but must not contribute to the gradient or the logged loss."""
        return step * self.world_size + self.rank >= self.num_batches_global

    def loss_scale(self, step: int) -> float:
        """Weight of this rank's mean-loss gradient at local step `step` such that the all-reduced MEAN over the ranks is
        the mean over the seeds of the whole round (the global batch): bs_r * R / sum_r bs_r over the ranks that hold a real
        batch; 0 for a padded one.  1.0 on a single rank.  Pure host arithmetic, identical on every rank."""
        if self.is_padded(step):
            return 0.0
        nb = self.num_batches_global
        lens = [self.batch_len(step * self.world_size + r) for r in range(self.world_size) if step * self.world_size + r < nb]
        total = sum(lens)
        mine = self.batch_len(step * self.world_size + self.rank)
        return float(mine) * self.world_size / float(total) if total > 0 else 0.0

    def round_is_full(self, step: int) -> bool:
        """True when EVERY rank holds a real, full-size batch at local step `step`.  The captured (replayed) step is only used
        for such rounds, so that all ranks take the same decision: a collective captured in a CUDA graph on one rank must not
        meet an eagerly issued one on another."""
        last = step * self.world_size + self.world_size - 1
        return last < self.num_batches_global and self.batch_len(last) == self.full_len

    def batch_seeds(self, order: torch.Tensor, global_batch_idx: int) -> torch.Tensor:
        b = global_batch_idx % max(self.num_batches_global, 1)
        return order[b * self.batch_size:(b + 1) * self.batch_size]
