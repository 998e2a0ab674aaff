"""SAGEPL extras oracle — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

* ``adding_noise``  restates reference src/models/layers/sagePL.py:41-49 with the same torch ops (clone, sign,
                    F.normalize, index, multiply, add), so autograd gives the reference's gradients.
* ``shuffle_rows``  sequential numpy twin of ngnn_shuffle_rows: the LAW of reference src/utils/augmentation.py:88-102
                    (per row, k = int(F * prob) distinct positions, their values permuted among themselves) driven by the
                    documented Philox stream instead of torch.randperm (RNG streams are not comparable; PARITY of the
                    random choices with the reference is therefore UNPINNED, the structure is checked by validity).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import philox


def adding_noise(x: torch.Tensor, noise: torch.Tensor, noise_rate: float, n_id=None) -> torch.Tensor:
    noisy_x = x.clone()
    if n_id is None:
        noisy_x = noisy_x + torch.sign(noisy_x) * F.normalize(noise) * noise_rate
    else:
        noisy_x = noisy_x + F.normalize(noise[n_id]) * noise_rate
    return noisy_x


def _row_words(row: int, n_words: int, seed: int, offset: int) -> np.ndarray:
    blocks = (n_words + 3) // 4
    q = np.arange(blocks, dtype=np.uint64)
    w = philox.philox4x32_10(np.uint64(row & 0xFFFFFFFF), q, offset & 0xFFFFFFFF, ((offset >> 32) ^ (row >> 32)) & 0xFFFFFFFF,
                             seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return np.stack(w, axis=-1).reshape(-1)


def shuffle_rows(x: np.ndarray, k: int, seed: int, offset: int) -> np.ndarray:
    out = x.copy()
    n, F_ = x.shape
    if k <= 1:
        return out
    k_al = (k + 3) // 4 * 4
    for row in range(n):
        words = _row_words(row, k_al + k, seed, offset)
        taken, pos = set(), []
        for j in range(k):
            jj = F_ - k + j
            t = (int(words[j]) * (jj + 1)) >> 32
            if t in taken:
                t = jj
            taken.add(t)
            pos.append(t)
        sel = list(pos)
        q = k_al
        for j in range(k - 1, 0, -1):
            u = (int(words[q]) * (j + 1)) >> 32
            sel[j], sel[u] = sel[u], sel[j]
            q += 1
        out[row, pos] = x[row, sel]
    return out


class SAGEPLRef(torch.nn.Module):
    """Structure of reference src/models/layers/sagePL.py:6-86 on the oracle SAGEConv (use_bn is dead in the reference)."""

    def __init__(self, in_size, hidden_size, out_size, num_layers, nbr_nodes, dropout=0.5, dtype=torch.float32):
        super().__init__()
        from .sage_oracle import SAGEConvRef
        self.num_layers, self.dropout = num_layers, dropout
        self.convs = torch.nn.ModuleList([SAGEConvRef(in_size, hidden_size, dtype=dtype)])
        for _ in range(num_layers - 2):
            self.convs.append(SAGEConvRef(hidden_size, hidden_size, dtype=dtype))
        self.convs.append(SAGEConvRef(hidden_size, out_size, dtype=dtype))
        self.noise = torch.nn.Parameter(torch.randn(nbr_nodes, in_size, dtype=dtype))

    def _stack(self, x, edge_index):
        h = None
        for i, conv in enumerate(self.convs):
            x = conv(x, edge_index)
            if i != self.num_layers - 1:
                x = x.relu()
                h = x
                x = F.dropout(x, p=self.dropout, training=self.training)
        return h, torch.log_softmax(x, dim=1), x

    def forward(self, x, edge_index, noise_rate=0.1, n_id=None):
        x_pure, y_pure, z_pure = self._stack(x, edge_index)
        noisy_x = adding_noise(x, self.noise, noise_rate, n_id)
        x_noisy, y_noisy, z_noisy = self._stack(noisy_x, edge_index)
        return x_pure, y_pure, z_pure, x_noisy, y_noisy, z_noisy
