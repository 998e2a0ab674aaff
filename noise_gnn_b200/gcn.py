"""SimpleGCN — the reference's GCN network (src/models/layers/convolution.py:7-53, ``module: 'gcn'`` in
config_cora5-8.yml / config_arxiv3-4.yml) on the B200 kernels: a stack of ``GCNConv(normalize=False)`` with ReLU +
dropout between layers.  Same constructor, attributes (``convs``) and methods as the reference class."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .conv import GCNConv


class SimpleGCN(torch.nn.Module):
    def __init__(self, in_size, hidden_size, out_size, num_layers, dropout=0.5, use_bn=False):
        super().__init__()
        self.num_layers, self.dropout = num_layers, dropout
        self.convs = torch.nn.ModuleList()
        self.convs.append(GCNConv(in_size, hidden_size, normalize=False))
        for _ in range(num_layers - 2):
            self.convs.append(GCNConv(hidden_size, hidden_size, normalize=False))
        self.convs.append(GCNConv(hidden_size, out_size, normalize=False))

    def reset_parameters(self):
        for conv in self.convs:
            conv.reset_parameters()

    def forward(self, x, edge_index):
        for i, conv in enumerate(self.convs):
            x = conv(x, edge_index)
            if i != self.num_layers - 1:
                x = x.relu()
                x = F.dropout(x, p=self.dropout, training=self.training)
        return x

    @torch.no_grad()
    def inference(self, x_all, subgraph_loader, device=None):
        """Layer-wise inference (reference convolution.py:37-53) with the activations kept on the GPU: every batch's
        rows are gathered from / written back to a device-resident [N, hidden] table by global id."""
        dev = next(self.parameters()).device
        x_all = x_all.to(dev, dtype=torch.float32)
        for i in range(self.num_layers):
            xs = []
            for batch in subgraph_loader:
                x = x_all.index_select(0, batch.n_id)
                x = self.convs[i](x, batch.edge_index)[: batch.batch_size]
                if i != self.num_layers - 1:
                    x = x.relu()
                xs.append(x)
            x_all = torch.cat(xs, dim=0)
        return x_all.cpu()
