"""CTLoss — drop-in for the reference's co-teaching loss (src/utils/losses.py:10-49; Han et al., NeurIPS 2018) that
keeps the whole computation on the GPU.  The reference moves the per-sample losses to the host twice per step for
``np.argsort`` (two synchronising round trips); ``ngnn_ct_loss`` ranks them on the device instead.

Same call signature and the same 8 return values as the reference module:

    loss_1, loss_2, pure_ratio_1, pure_ratio_2, ind_1_update, ind_2_update, ind_noisy_1, ind_noisy_2
        = CTLoss(device)(y_1, y_2, y_noise, forget_rate, ind, noise_or_not)

``loss_1`` / ``loss_2`` are differentiable w.r.t. ``y_1`` / ``y_2`` (network 1 learns from the rows network 2 finds
easiest and vice versa).  Ties in the per-sample loss are broken by row index (a stable argsort; numpy's default
quicksort leaves them unspecified)."""
from __future__ import annotations

import torch

from . import ops


class _ScaledGrad(torch.autograd.Function):
    """loss value with a precomputed gradient: d(loss)/d(logits) = d.  One node PER loss, each saving only its own
    gradient, so the reference's pattern ``loss_1.backward(); optimizer1.step(); loss_2.backward()``
    (src/pipeline.py:127-133) works: the first backward frees nothing the second one needs."""

    @staticmethod
    def forward(ctx, logits, value, d):
        ctx.save_for_backward(d)
        return value.clone()

    @staticmethod
    def backward(ctx, g):
        (d,) = ctx.saved_tensors
        return d * g, None, None


class CTLoss(torch.nn.Module):
    def __init__(self, device=None):
        super().__init__()
        self.device = device

    def forward(self, y_1, y_2, y_noise, forget_rate, ind=None, noise_or_not=None):
        if not y_1.is_cuda:
            raise RuntimeError("noise_gnn_b200.CTLoss runs on CUDA tensors only (no CPU fallback)")
        n = y_1.size(0)
        num_remember = int((1 - forget_rate) * n)                      # reference losses.py:28-29
        clean_rows = None
        if noise_or_not is not None and ind is not None:               # reference: noise_or_not[ind.cpu()[...]] on the host
            clean_rows = torch.as_tensor(noise_or_not).to(y_1.device).view(-1)[ind[:n].long()].to(torch.uint8)
        y_1, y_2 = y_1.float(), y_2.float()
        with torch.no_grad():      # both per-sample losses, the device-side ranking / exchange and both gradients: one call
            stats, d1, d2, o1, o2 = ops.ct_loss(y_1, y_2, y_noise.view(-1), num_remember, clean_mask=clean_rows,
                                                want_grad=True, want_order=True)
        loss_1 = _ScaledGrad.apply(y_1, stats[0], d1)
        loss_2 = _ScaledGrad.apply(y_2, stats[1], d2)
        pure_1, pure_2 = stats[4], stats[5]
        o1, o2 = o1.long(), o2.long()
        return (loss_1, loss_2, pure_1, pure_2, o1[:num_remember], o2[:num_remember], o1[num_remember:], o2[num_remember:])
