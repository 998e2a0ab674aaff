"""Train step of the hot loop — the body of ``PipelineCO.train`` (reference src/pipeline.py:152-169) on the
B200-native path, host-sync-free after sampling:

    batch -> SAGE.forward_batch (trimmed, fused) -> softmax-CE + accuracy count on the seed rows
          -> backward -> [DP: one NCCL all-reduce of the flat gradient bucket] -> fused Adam.

Parameters, gradients and Adam state live in flat fp32 buckets (the model's ``Parameter``s are views into
them, so ``state_dict`` keeps PyG's names); loss / accuracy accumulate in a device tensor and are read by
the caller when it wants them (the reference reads them every step, src/pipeline.py:164-165).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops


class FlatBuckets:
    """Re-homes a module's parameters (and their .grad) into contiguous flat fp32 buffers."""

    def __init__(self, module: torch.nn.Module):
        params = [p for p in module.parameters() if p.requires_grad]
        if not params:
            raise ValueError("module has no trainable parameters")
        dev = params[0].device
        total = sum(p.numel() for p in params)
        self.param = torch.empty(total, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        for p in params:
            n = p.numel()
            self.param[off:off + n].copy_(p.data.view(-1))
            p.data = self.param[off:off + n].view_as(p)
            p.grad = self.grad[off:off + n].view_as(p)
            off += n
        self.params = params
        self.numel = total

    def rebind_grads(self):
        """autograd may replace p.grad when it was None; keep them pointing into the bucket."""
        off = 0
        for p in self.params:
            n = p.numel()
            view = self.grad[off:off + n].view_as(p)
            if p.grad is None or p.grad.data_ptr() != view.data_ptr():
                if p.grad is not None:
                    view.copy_(p.grad)
                p.grad = view
            off += n


class Trainer:
    """Owns the flat buckets and the fused optimizer for one SAGE network."""

    def __init__(self, model, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 process_group=None, world_size: int = 1):
        self.model = model
        self.buckets = FlatBuckets(model)
        dev = self.buckets.param.device
        self.exp_avg = torch.zeros_like(self.buckets.param)
        self.exp_avg_sq = torch.zeros_like(self.buckets.param)
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self.stats = torch.zeros(2, dtype=torch.float32, device=dev)     # [sum of step losses, #correct]
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.process_group, self.world_size = process_group, world_size

    def reset_stats(self):
        self.stats.zero_()

    def read_stats(self):
        """(sum of per-step mean losses, number of correct seed predictions) — one device->host read."""
        s = self.stats.tolist()
        return s[0], int(s[1])

    def train_step(self, batch, target_attr: str = "yhn", label_attr: Optional[str] = "y"):
        model, bk = self.model, self.buckets
        bs = batch.batch_size
        logits = model.forward_batch(batch)
        target = getattr(batch, target_attr)[:bs].view(-1)
        y_true = getattr(batch, label_attr)[:bs].view(-1) if label_attr else None
        _, dlogits = ops.ce_fwd_bwd(logits.detach(), target, y_true, stats=self.stats)
        bk.grad.zero_()                       # optimizer.zero_grad()
        logits.backward(dlogits)              # loss.backward(): wgrad / dgrad / transpose-sum kernels
        bk.rebind_grads()
        scale = 1.0
        if self.world_size > 1:               # DP: one all-reduce of the flat bucket, then average
            torch.distributed.all_reduce(bk.grad, group=self.process_group)
            scale = 1.0 / self.world_size
        ops.adam_step(bk.param, bk.grad, self.exp_avg, self.exp_avg_sq, self.step_dev, lr=self.lr, betas=self.betas,
                      eps=self.eps, weight_decay=self.weight_decay, grad_scale=scale)
        return logits
