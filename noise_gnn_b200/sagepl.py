"""SAGEPL + shuffle_pos — the reference's noise-injecting GraphSAGE variant (src/models/layers/sagePL.py:6-104) and its
feature-shuffling augmentation (src/utils/augmentation.py:88-102) on the B200-native kernels.

``SAGEPL`` keeps the reference module's constructor, attributes (``convs``, ``noise``), methods and the six return
values of ``forward``; its convolutions are the drop-in ``SAGEConv`` (every layer on the whole block, ``x`` gradients
included — the learnable noise sits in front of layer 1, so layer 1's data gradient IS needed here), and
``adding_noise`` is one fused kernel each way (gather the noise rows by ``n_id``, L2-normalise, scale, add) instead of
seven elementwise passes.  ``shuffle_pos`` replaces a Python loop over every row of the block (two ``torch.randperm``
per row: minutes per products-sized block) by one launch with a counter-based Philox stream; the random streams differ
from torch's, the law is the same (per row, ``int(F * prob)`` distinct positions permuted among themselves).
The unmodified reference ``sagePL.py`` also runs on the drop-in ``SAGEConv`` through ``compat/torch_geometric``.
"""
from __future__ import annotations

import itertools

import torch
import torch.nn.functional as F

from . import ops
from .conv import SAGEConv

_calls = itertools.count()


class SAGEPL(torch.nn.Module):
    def __init__(self, in_size, hidden_size, out_size, num_layers, nbr_nodes, dropout=0.5, use_bn=False):
        super().__init__()
        if num_layers < 2:
            raise NotImplementedError("SAGEPL needs num_layers >= 2 (every reference config uses 2 or 3)")
        self.num_layers, self.dropout, self.use_bn = num_layers, dropout, use_bn
        dims = [in_size] + [hidden_size] * (num_layers - 1) + [out_size]
        self.convs = torch.nn.ModuleList(SAGEConv(dims[i], dims[i + 1]) for i in range(num_layers))
        self.noise = torch.nn.Parameter(torch.randn(nbr_nodes, in_size))       # adaptive noise (reference sagePL.py:22)
        if use_bn:
            self.bn1 = torch.nn.BatchNorm1d(in_size)
            self.bn2 = torch.nn.BatchNorm1d(hidden_size)

    def reset_parameters(self):
        for conv in self.convs:
            conv.reset_parameters()

    def forward(self, x, edge_index, noise_rate=0.1, n_id=None):
        x_pure, y_pure, z_pure = self.forward_pure(x, edge_index)
        noisy_x = self.adding_noise(x, noise_rate=noise_rate, n_id=n_id)
        x_noisy, y_noisy, z_noisy = self.forward_noisy(noisy_x, edge_index)
        return x_pure, y_pure, z_pure, x_noisy, y_noisy, z_noisy

    def adding_noise(self, x, noise_rate, n_id=None):
        """x + F.normalize(noise[n_id]) * rate   (n_id given), or x + sign(x) * F.normalize(noise) * rate (whole graph)."""
        if not x.is_cuda:
            raise RuntimeError("noise_gnn_b200.SAGEPL runs on CUDA tensors only (no CPU fallback)")
        idx = None if n_id is None else n_id.to(torch.int32).contiguous()
        if idx is None and x.size(0) != self.noise.size(0):
            raise ValueError("adding_noise without n_id needs one feature row per noise row")
        return ops.NoiseAddFunction.apply(x.float(), self.noise, idx, float(noise_rate), n_id is None)

    def _stack(self, x, edge_index):
        if self.use_bn:
            x = self.bn1(x)
        h = None
        for i, conv in enumerate(self.convs):
            x = conv(x, edge_index)
            if i != self.num_layers - 1:
                x = x.relu()
                if self.use_bn:
                    x = self.bn2(x)
                h = x
                x = F.dropout(x, p=self.dropout, training=self.training)
        return h, torch.log_softmax(x, dim=1), x

    def forward_pure(self, x, edge_index):
        return self._stack(x, edge_index)

    def forward_noisy(self, x, edge_index):
        return self._stack(x, edge_index)

    @torch.no_grad()
    def inference(self, x_all, subgraph_loader, device=None):
        """Layer-wise inference (reference sagePL.py:88-104): identical to SAGE.inference."""
        from .sage import SAGE
        return SAGE.inference(self, x_all, subgraph_loader, device)


def shuffle_pos(features: torch.Tensor, device="cuda", prob: float = 0.1, seed: int = 1232) -> torch.Tensor:
    """Drop-in for reference ``utils.augmentation.shuffle_pos(features, device, prob)``: a detached copy of ``features`` in
    which, per row, ``int(F * prob)`` distinct random positions have their values shuffled among themselves."""
    x = features.detach()
    if not x.is_cuda:
        x = x.to(device)
    if not x.is_cuda:
        raise RuntimeError("noise_gnn_b200.shuffle_pos runs on the GPU (no CPU fallback)")
    k = int(x.shape[1] * prob)
    return ops.shuffle_rows(x.float(), k, seed=seed, offset=next(_calls))
