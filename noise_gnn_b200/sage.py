"""SAGE — the reference's GraphSAGE network (src/models/layers/sage.py:6-78) on B200-native kernels.

Same constructor, attributes and methods as the reference module
(``SAGE(in_size, hidden_size, out_size, num_layers, dropout=0.5, use_bn=False)``, ``.convs``,
``.reset_parameters()``, ``.forward(x, edge_index)``, ``.inference(x_all, subgraph_loader, device)``) and
the same state_dict keys (``convs.{i}.lin_l.weight|bias``, ``convs.{i}.lin_r.weight``), so it can be
swapped in where ``NGNN.init_network`` builds the reference module (src/models/model.py:44-50).
(The unmodified reference ``sage.py`` also runs on these kernels through ``compat/torch_geometric``.)

Two execution modes:

* ``forward(x, edge_index)`` — reference-exact: every layer on the whole block, ReLU then dropout
  between layers (torch's generator supplies the dropout mask, as in the reference).
* ``forward_batch(batch)`` — the B200-first path for blocks that come from our NeighborLoader: layer
  ``l`` of ``L`` only computes the rows within ``L-l`` hops of the seeds (prefixes of the block, exact
  for the seed rows the caller keeps — SURVEY §8 trimming note), layer 1 aggregates straight from the
  resident feature table (no ``x[n_id]`` materialisation), and ReLU + dropout are fused into the GEMM
  epilogue with a counter-based Philox mask.  Returns ``[batch_size, out_size]``.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import ops
from .conv import SAGEConv, _block_cache

import itertools

NGNN_ACT_NONE, NGNN_ACT_RELU = 0, 1
_instances = itertools.count()


class _SAGELayerFunction(torch.autograd.Function):
    """One trimmed SAGE layer with fused epilogue.

    x_src: rows the edges read (a previous layer's output, or the resident feature table when
    `table_mode`), block prefix (n_dst rows, e_limit edges).  act/dropout fused in the GEMM; the saved
    post-activation output gates the backward (h > 0 <=> pre-activation > 0 and kept)."""

    @staticmethod
    def forward(ctx, x_src, w_l, b_l, w_r, block, n_dst, e_limit, n_src, table_mode, act, drop_p, seed, offset, tag,
                in_gate_scale=0.0, grad_pregated=False):
        if table_mode:   # aggregate from the resident table by global ids, gather the root rows in the same launch
            mean, root = ops.agg_fwd(block.rowptr, block.col_global, x_src, n_dst, root_idx=block.n_id, tag="agg_" + tag)
        else:
            mean, root = ops.agg_fwd(block.rowptr, block.col, x_src, n_dst, tag="agg_" + tag), x_src
        out = ops.gemm_fwd(mean, root, w_l, w_r, b_l, n_dst, act=act, drop_p=drop_p, seed=seed, offset=offset,
                           tag="gemm_" + tag)
        ctx.save_for_backward(mean, root, w_l, w_r, out if act or drop_p > 0 else None)
        ctx.block, ctx.n_dst, ctx.e_limit, ctx.n_src = block, n_dst, e_limit, n_src
        ctx.table_mode, ctx.act, ctx.drop_p, ctx.tag = table_mode, act, drop_p, tag
        ctx.in_gate_scale, ctx.grad_pregated = in_gate_scale, grad_pregated
        return out

    @staticmethod
    def backward(ctx, dy):
        mean, root, w_l, w_r, out = ctx.saved_tensors
        x_src_saved = root   # in layer mode the root operand IS x_src (all n_src rows)
        block, n_dst = ctx.block, ctx.n_dst
        if out is not None and not ctx.grad_pregated:   # else the consumer layer already applied this layer's ReLU/dropout gate
            dy = ops.act_bwd(dy, out, 1.0 / (1.0 - ctx.drop_p))
        F_ = mean.size(1)
        dw_l, dw_r, db = ops.wgrad(dy, mean, root, n_dst, F_, tag="wgrad_" + ctx.tag)
        dx = None
        if ctx.needs_input_grad[0] and not ctx.table_mode:
            dmean, droot = ops.dgrad(dy, w_l, w_r, block.rowptr, n_dst, tag="dgrad_" + ctx.tag)
            colptr_t, row_t = block.transpose(ctx.e_limit, ctx.n_src)
            # x_src is the producing layer's post-activation output: fold its ReLU+dropout backward into this pass
            gate = x_src_saved if ctx.in_gate_scale > 0 else None
            dx = ops.agg_bwd(colptr_t, row_t, dmean, ctx.n_src, dx_root=droot, n_root=n_dst, act_ref=gate,
                             act_scale=ctx.in_gate_scale, tag="aggT_" + ctx.tag)
        return (dx, dw_l, db, dw_r) + (None,) * 12


class SAGE(torch.nn.Module):
    def __init__(self, in_size, hidden_size, out_size, num_layers, dropout=0.5, use_bn=False):
        super().__init__()
        if num_layers < 2:
            # the reference constructor always appends a first and a last conv (sage.py:16-19), so num_layers=1 there
            # means two convs with the activation after the LAST one — a quirk no reference config uses
            raise NotImplementedError("SAGE needs num_layers >= 2 (every reference config uses 2 or 3)")
        self.num_layers, self.dropout, self.use_bn = num_layers, dropout, use_bn
        dims = [in_size] + [hidden_size] * (num_layers - 1) + [out_size]
        self.convs = torch.nn.ModuleList(SAGEConv(dims[i], dims[i + 1]) for i in range(num_layers))
        if use_bn:   # dead in the reference (no caller passes use_bn=True); kept for API parity, torch BN
            self.bn1 = torch.nn.BatchNorm1d(in_size)
            self.bn2 = torch.nn.BatchNorm1d(hidden_size)
        self._drop_calls = 0
        # key of the fused (Philox) dropout stream: follows torch's seed, distinct per network instance — two peer networks
        # of a co-teaching pair must not share their masks (the reference draws them independently from torch's generator)
        self.drop_seed = (torch.initial_seed() * 1000003 + next(_instances)) & (2**63 - 1)

    def reset_parameters(self):
        for conv in self.convs:
            conv.reset_parameters()

    # ---- reference-exact mode ------------------------------------------------------------
    def forward(self, x, edge_index):
        if self.use_bn:
            x = self.bn1(x)
        last = self.num_layers - 1
        for i, conv in enumerate(self.convs):
            x = conv(x, edge_index)
            if i != last:
                x = x.relu()
                if self.use_bn:
                    x = self.bn2(x)
                x = F.dropout(x, p=self.dropout, training=self.training)
        return x

    # ---- trimmed, fused mode ---------------------------------------------------------------
    @staticmethod
    def layer_extents(block: ops.Block, num_layers: int):
        """Per layer (n_dst, e_limit, n_src): rows within L-l hops of the seeds, the edges into them, and
        the rows those edges read — all prefixes of the block (SURVEY §8 trimming note)."""
        hn, he = block.hop_nodes, block.hop_edges
        H = len(hn) - 1
        ext = []
        for layer in range(1, num_layers + 1):
            d = num_layers - layer
            ext.append((hn[min(d, H)], he[min(d + 1, H)], hn[min(d + 1, H)]))
        return ext

    def forward_batch(self, batch, x_table=None, return_hidden: bool = False):
        """Trimmed fused forward on a Batch of our NeighborLoader; returns logits of the seed rows (and, with
        return_hidden, the list of every layer's output rows)."""
        if self.use_bn:
            raise NotImplementedError("use_bn is dead code in the reference; the fused path does not cover it")
        block = batch.block
        table = x_table if x_table is not None else batch._loader.x
        p = float(self.dropout) if self.training else 0.0
        self._drop_calls += 1
        ext = self.layer_extents(block, self.num_layers)
        h = table
        hidden = []
        last = self.num_layers - 1
        for i, conv in enumerate(self.convs):
            n_dst, e_limit, n_src = ext[i]
            act = NGNN_ACT_RELU if i != last else NGNN_ACT_NONE
            h = _SAGELayerFunction.apply(h, conv.lin_l.weight, conv.lin_l.bias, conv.lin_r.weight, block, n_dst, e_limit,
                                         n_src, i == 0, act, p if i != last else 0.0, self.drop_seed,
                                         self._drop_calls * self.num_layers + i, f"l{i + 1}",
                                         (1.0 / (1.0 - p)) if i > 0 else 0.0,   # input = previous layer's relu/dropout output
                                         i != last)                              # the next layer gates this layer's gradient
            hidden.append(h)
        return (h[: batch.batch_size], hidden) if return_hidden else h[: batch.batch_size]

    # ---- layer-wise inference (reference sage.py:42-58) -----------------------------------------
    @torch.no_grad()
    def inference(self, x_all, subgraph_loader, device=None, return_cpu: bool = True):
        """Layer by layer over all input nodes of `subgraph_loader`, with the loader's sampled fan-outs (as the reference
        does: its eval loader keeps `num_neighbors`, pipeline.py:85-92).  Activations stay on the GPU between layers.
        Per batch the reference runs one conv on the whole multi-hop block and keeps `[:batch_size]`; only the hop-1
        edges reach those rows, and a node's hop-1 draws do not depend on the deeper hops (the sampler's RNG is keyed by
        node and hop), so the loader is asked for hop 1 only and K-AGG gathers straight from the [N, F] activation table
        by global id (no x[n_id] copy).  Seed-row outputs are the same function of the same sampled edges."""
        dev = subgraph_loader.device
        nodes = subgraph_loader.input_nodes
        if nodes.numel() != subgraph_loader.num_nodes or not torch.equal(nodes, torch.arange(nodes.numel())):
            # every layer's output table is indexed by GLOBAL node id by the next layer (x_all[batch.n_id], reference
            # sage.py:50): that is only meaningful when the loader walks all nodes in id order, as the reference's
            # subgraph loaders do (input_nodes=None).  The reference would raise an IndexError or silently misalign here.
            raise ValueError("SAGE.inference needs a loader over all nodes in id order (input_nodes=None, shuffle=False)")
        if subgraph_loader.shuffle:
            raise ValueError("SAGE.inference needs an unshuffled loader (rows of the layer outputs are node ids)")
        x_all = x_all.to(dev, dtype=torch.float32)
        last = self.num_layers - 1
        hops = list(subgraph_loader.num_neighbors)
        subgraph_loader.set_num_neighbors(hops[:1])
        try:
            for i, conv in enumerate(self.convs):
                xs = []
                for batch in subgraph_loader:
                    blk = batch.block
                    bs = batch.batch_size
                    mean, root = ops.agg_fwd(blk.rowptr, blk.col_global, x_all, bs, root_idx=blk.n_id)
                    out = ops.gemm_fwd(mean, root, conv.lin_l.weight, conv.lin_r.weight, conv.lin_l.bias, bs,
                                       act=NGNN_ACT_RELU if i != last else NGNN_ACT_NONE)
                    xs.append(out)
                x_all = torch.cat(xs, dim=0)
        finally:
            subgraph_loader.set_num_neighbors(hops)
        return x_all.cpu() if return_cpu else x_all
