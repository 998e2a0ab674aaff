"""SAGEConv / SAGE / train-step oracle (pure torch, CPU) — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

PARITY UNPINNED (torch-geometric is not installable here; the reference holds no golden vectors).
Restated from the published PyG 2.5.1 implementation the reference calls:

* ``SAGEConvRef``  = ``torch_geometric.nn.SAGEConv(in, out)`` with its defaults (aggr='mean',
  root_weight=True, bias=True, normalize=False, project=False), constructed at reference
  src/models/layers/sage.py:16-19 and called at :34.  The forward issues the same torch op sequence
  PyG lowers to: ``x.index_select(0, src)`` -> ``scatter_add_`` by dst -> degree count by
  ``scatter_add_`` of ones -> ``clamp(min=1)`` -> divide -> ``lin_l`` (with bias) + ``lin_r`` (no bias).
  Rows without in-edges aggregate to 0; duplicate edges count each time; no self loops are added.
  Parameter names/shapes match PyG (``lin_l.weight [O,F]``, ``lin_l.bias [O]``, ``lin_r.weight [O,F]``)
  so state_dicts interchange.  Init = PyG ``Linear.reset_parameters`` (kaiming-uniform a=sqrt(5) for
  weights => U(+-1/sqrt(F)); bias U(+-1/sqrt(F))), the same law as ``torch.nn.Linear``.
* ``SAGERef``      = reference src/models/layers/sage.py:6-40 verbatim in structure (conv -> relu ->
  dropout between layers, nothing after the last, every layer on the whole block).
* ``train_step``   = loop body of ``PipelineCO.train`` (reference src/pipeline.py:152-169) with
  ``compare_loss == 'normal'``.

Because it uses exactly the ops the PyG CPU path runs, the same code is the timed CPU baseline
(``bench.py`` ``cpu_baseline`` / ``--impl reference``).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def mean_aggregate(x: torch.Tensor, edge_index: torch.Tensor, n_dst: int | None = None) -> torch.Tensor:
    """PyG MeanAggregation via utils.scatter(reduce='mean'): sum / clamp(count, 1)."""
    src, dst = edge_index[0], edge_index[1]
    n = x.size(0) if n_dst is None else n_dst
    msg = x.index_select(0, src)
    out = x.new_zeros((n, x.size(1))).scatter_add_(0, dst.view(-1, 1).expand_as(msg), msg)
    cnt = x.new_zeros(n).scatter_add_(0, dst, x.new_ones(dst.numel()))
    return out / cnt.clamp(min=1).view(-1, 1)


class SAGEConvRef(torch.nn.Module):
    def __init__(self, in_channels: int, out_channels: int, normalize: bool = False, dtype=torch.float32):
        super().__init__()
        self.in_channels, self.out_channels, self.normalize = in_channels, out_channels, normalize
        self.lin_l = torch.nn.Linear(in_channels, out_channels, bias=True, dtype=dtype)
        self.lin_r = torch.nn.Linear(in_channels, out_channels, bias=False, dtype=dtype)

    def reset_parameters(self):
        self.lin_l.reset_parameters()
        self.lin_r.reset_parameters()

    def forward(self, x, edge_index):
        out = self.lin_l(mean_aggregate(x, edge_index)) + self.lin_r(x)
        if self.normalize:
            out = F.normalize(out, p=2.0, dim=-1)
        return out


class GCNConvRef(torch.nn.Module):
    """``torch_geometric.nn.GCNConv(in, out, normalize=False)`` as the reference builds it
    (src/models/layers/convolution.py:19-23): ``x = lin(x)`` (no bias), messages ``x.index_select(0, src)`` summed by
    ``scatter_add_`` over dst (aggr='add'; no self loops and no symmetric norm because normalize=False), then ``+ bias``.
    Parameter names as in PyG: ``lin.weight`` (glorot), ``bias`` (zeros)."""

    def __init__(self, in_channels: int, out_channels: int, dtype=torch.float32):
        super().__init__()
        self.lin = torch.nn.Linear(in_channels, out_channels, bias=False, dtype=dtype)
        self.bias = torch.nn.Parameter(torch.zeros(out_channels, dtype=dtype))
        torch.nn.init.xavier_uniform_(self.lin.weight)

    def forward(self, x, edge_index):
        z = self.lin(x)
        src, dst = edge_index[0], edge_index[1]
        msg = z.index_select(0, src)
        out = z.new_zeros(z.shape).scatter_add_(0, dst.view(-1, 1).expand_as(msg), msg)
        return out + self.bias


class SimpleGCNRef(torch.nn.Module):
    """Structure of reference src/models/layers/convolution.py:7-35."""

    def __init__(self, in_size, hidden_size, out_size, num_layers, dropout=0.5, dtype=torch.float32):
        super().__init__()
        self.num_layers, self.dropout = num_layers, dropout
        self.convs = torch.nn.ModuleList([GCNConvRef(in_size, hidden_size, dtype=dtype)])
        for _ in range(num_layers - 2):
            self.convs.append(GCNConvRef(hidden_size, hidden_size, dtype=dtype))
        self.convs.append(GCNConvRef(hidden_size, out_size, dtype=dtype))

    def forward(self, x, edge_index):
        for i, conv in enumerate(self.convs):
            x = conv(x, edge_index)
            if i != self.num_layers - 1:
                x = x.relu()
                x = F.dropout(x, p=self.dropout, training=self.training)
        return x


class SAGERef(torch.nn.Module):
    """Structure of reference src/models/layers/sage.py:6-40 (use_bn is dead in the reference: no caller sets it)."""

    def __init__(self, in_size, hidden_size, out_size, num_layers, dropout=0.5, dtype=torch.float32):
        super().__init__()
        self.num_layers, self.dropout = num_layers, dropout
        self.convs = torch.nn.ModuleList()
        self.convs.append(SAGEConvRef(in_size, hidden_size, dtype=dtype))
        for _ in range(num_layers - 2):
            self.convs.append(SAGEConvRef(hidden_size, hidden_size, dtype=dtype))
        self.convs.append(SAGEConvRef(hidden_size, out_size, dtype=dtype))

    def reset_parameters(self):
        for conv in self.convs:
            conv.reset_parameters()

    def forward(self, x, edge_index, dropout_masks=None, relu_masks=None):
        """dropout_masks: optional list of keep-masks (one per hidden layer) to replace torch's RNG, so a
        fused-dropout kernel can be compared on the identical mask.
        relu_masks: optional list of boolean "pre-activation counted as positive" masks (one per hidden layer; rows beyond a
        mask's length use the oracle's own sign).  ReLU is discontinuous in its gradient: an fp32 implementation and this
        fp64 oracle legitimately disagree on the sign of a pre-activation that is zero to rounding, and ONE such element
        moves a weight gradient of a 77 k-row block by ~1e-3 (its terms cancel to ~1/sqrt(n) of their magnitude).  Gradient
        parity at scale is therefore taken on identical gates, like dropout parity on identical masks."""
        for i, conv in enumerate(self.convs):
            x = conv(x, edge_index)
            if i != self.num_layers - 1:
                if relu_masks is not None:
                    m = x > 0
                    r = relu_masks[i]
                    m[: r.size(0)] = r
                    x = torch.where(m, x, torch.zeros_like(x))
                else:
                    x = x.relu()
                if dropout_masks is not None:
                    m = dropout_masks[i]
                    x = x * m[: x.size(0)].to(x.dtype) / (1.0 - self.dropout)
                else:
                    x = F.dropout(x, p=self.dropout, training=self.training)
        return x

    @torch.no_grad()
    def inference(self, x_all, subgraph_loader, device="cpu"):
        """Layer-wise inference, reference src/models/layers/sage.py:42-58."""
        for i in range(self.num_layers):
            xs = []
            for batch in subgraph_loader:
                x = x_all[batch.n_id].to(device)
                x = self.convs[i](x, batch.edge_index.to(device))[: batch.batch_size]
                if i != self.num_layers - 1:
                    x = x.relu()
                xs.append(x.cpu())
            x_all = torch.cat(xs, dim=0)
        return x_all


def train_step(model: SAGERef, optimizer, x, edge_index, y, yhn, batch_size: int):
    """Loop body of PipelineCO.train (reference src/pipeline.py:152-169, compare_loss 'normal')."""
    out = model(x, edge_index)[:batch_size]
    y = y[:batch_size].squeeze()
    yhn = yhn[:batch_size].squeeze()
    loss = F.cross_entropy(out, yhn)
    total_loss = float(loss.detach())
    total_correct = int(out.argmax(dim=-1).eq(y).sum())
    optimizer.zero_grad()
    loss.backward()
    optimizer.step()
    return total_loss, total_correct
