"""SAGEConv — drop-in for ``torch_geometric.nn.SAGEConv`` as the reference uses it.

Reference call sites: constructed at src/models/layers/sage.py:16-19 (and sagePL.py:16-19, sageH.py,
sageFC.py, gcn.py:12 with ``normalize=False``), called as ``conv(x, edge_index)`` at sage.py:34,52.
Same constructor arguments, parameter names and shapes (``lin_l.weight [O,F]``, ``lin_l.bias [O]``,
``lin_r.weight [O,F]``), initialisation law and error behaviour as PyG's module with its defaults
(aggr='mean', root_weight=True, bias=True, project=False), so state_dicts interchange.

forward(x: fp32 [n,F], edge_index: int64 [2,e] (row 0 = source, row 1 = destination; any order,
duplicates counted)) -> fp32 [n,O]; differentiable w.r.t. the parameters and, when required, x.
All arithmetic runs in libngnn_b200.so (CSR segment mean -> fused lin_l+lin_r+bias GEMM; backward:
weight-gradient GEMMs, data-gradient GEMM, transpose segment sum).  No CPU fallback.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

from . import ops


class _Linear(torch.nn.Module):
    """Parameter holder named like PyG's ``Linear`` (weight [out,in], optional bias)."""

    def __init__(self, in_channels: int, out_channels: int, bias: bool):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = torch.nn.Parameter(torch.empty(out_channels, in_channels))
        if bias:
            self.bias = torch.nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        # PyG Linear.reset_parameters == torch.nn.Linear's law: kaiming_uniform(a=sqrt(5)) => U(+-1/sqrt(fan_in))
        torch.nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            bound = 1.0 / math.sqrt(self.in_channels) if self.in_channels > 0 else 0.0
            torch.nn.init.uniform_(self.bias, -bound, bound)


class _BlockCache:
    """edge_index tensor -> CSR Block.  Holds a reference to the tensor so its storage cannot be
    recycled under a stale key; keyed on identity + in-place version."""

    def __init__(self, capacity: int = 8):
        self.capacity = capacity
        self.entries: OrderedDict = OrderedDict()

    def get(self, edge_index: torch.Tensor, n_rows: int) -> ops.Block:
        blk = getattr(edge_index, "_ngnn_block", None)
        if blk is not None and blk.n_rows == n_rows:
            return blk
        key = (id(edge_index), edge_index._version, n_rows)
        hit = self.entries.get(key)
        if hit is not None and hit[0] is edge_index:
            self.entries.move_to_end(key)
            return hit[1]
        blk = ops.coo_to_csr(edge_index, n_rows)
        self.entries[key] = (edge_index, blk)
        while len(self.entries) > self.capacity:
            self.entries.popitem(last=False)
        return blk


_block_cache = _BlockCache()


class SAGEConv(torch.nn.Module):
    def __init__(self, in_channels: int, out_channels: int, aggr: str = "mean", normalize: bool = False,
                 root_weight: bool = True, project: bool = False, bias: bool = True, **kwargs):
        super().__init__()
        if isinstance(in_channels, (tuple, list)):
            raise NotImplementedError("bipartite (in_src, in_dst) channels are not part of the reference's use of SAGEConv")
        if aggr != "mean":
            raise NotImplementedError(f"aggr={aggr!r}: the reference only uses the default mean aggregation")
        if project:
            raise NotImplementedError("project=True is not used by the reference")
        if kwargs:
            raise TypeError(f"unsupported SAGEConv arguments: {sorted(kwargs)}")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.normalize, self.root_weight = normalize, root_weight
        self.lin_l = _Linear(in_channels, out_channels, bias=bias)
        if root_weight:
            self.lin_r = _Linear(in_channels, out_channels, bias=False)

    def reset_parameters(self):
        self.lin_l.reset_parameters()
        if self.root_weight:
            self.lin_r.reset_parameters()

    def forward(self, x: torch.Tensor, edge_index: torch.Tensor, size=None) -> torch.Tensor:
        if size is not None:
            raise NotImplementedError("size= (bipartite propagation) is not used by the reference")
        if not x.is_cuda:
            raise RuntimeError("noise_gnn_b200.SAGEConv runs on CUDA tensors only (no CPU fallback); move the "
                               "module and inputs to a B200 device")
        if x.dim() != 2 or x.size(1) != self.in_channels:
            raise ValueError(f"x must be [n, {self.in_channels}], got {tuple(x.shape)}")
        n = x.size(0)
        block = _block_cache.get(edge_index, n)
        w_r = self.lin_r.weight if self.root_weight else None
        out = ops.SAGEConvFunction.apply(x.float(), self.lin_l.weight, self.lin_l.bias, w_r, block, n, block.e)
        if self.normalize:
            out = F.normalize(out, p=2.0, dim=-1)
        return out

    def __repr__(self):
        return f"{self.__class__.__name__}({self.in_channels}, {self.out_channels}, aggr=mean)"


class GCNConv(torch.nn.Module):
    """Drop-in for ``torch_geometric.nn.GCNConv(in, out, normalize=False)`` — the only way the reference builds it
    (src/models/layers/convolution.py:19-23): ``out = A_sum (x W^T) + bias`` with sum aggregation over in-neighbours,
    duplicate edges counted, no self loops, no symmetric normalisation.  Parameter names follow PyG
    (``lin.weight [O, F]`` glorot-initialised, ``bias [O]`` zeros) so state_dicts interchange."""

    def __init__(self, in_channels: int, out_channels: int, improved: bool = False, cached: bool = False,
                 add_self_loops=None, normalize: bool = True, bias: bool = True, **kwargs):
        super().__init__()
        if normalize:
            raise NotImplementedError("GCNConv(normalize=True) (symmetric normalisation + self loops) is not used by the "
                                      "reference, which always passes normalize=False")
        if add_self_loops:
            raise NotImplementedError("add_self_loops=True is not used by the reference")
        if improved or cached or kwargs:
            raise NotImplementedError("unsupported GCNConv arguments for the reference's use (normalize=False only)")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin = _Linear(in_channels, out_channels, bias=False)
        self.bias = torch.nn.Parameter(torch.empty(out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        torch.nn.init.xavier_uniform_(self.lin.weight)     # PyG Linear(weight_initializer='glorot')
        if self.bias is not None:
            torch.nn.init.zeros_(self.bias)

    def forward(self, x: torch.Tensor, edge_index: torch.Tensor, edge_weight=None) -> torch.Tensor:
        if edge_weight is not None:
            raise NotImplementedError("edge_weight is not used by the reference")
        if not x.is_cuda:
            raise RuntimeError("noise_gnn_b200.GCNConv runs on CUDA tensors only (no CPU fallback)")
        if x.dim() != 2 or x.size(1) != self.in_channels:
            raise ValueError(f"x must be [n, {self.in_channels}], got {tuple(x.shape)}")
        block = _block_cache.get(edge_index, x.size(0))
        return ops.GCNConvFunction.apply(x.float(), self.lin.weight, self.bias, block)

    def __repr__(self):
        return f"{self.__class__.__name__}({self.in_channels}, {self.out_channels}, normalize=False)"
