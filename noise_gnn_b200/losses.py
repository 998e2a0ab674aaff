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


class _CTLossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y_1, y_2, y_noise, num_remember, clean_rows):
        stats, d1, d2, o1, o2 = ops.ct_loss(y_1.detach(), y_2.detach(), y_noise, num_remember, clean_mask=clean_rows,
                                            want_grad=True, want_order=True)
        ctx.save_for_backward(d1, d2)
        ctx.mark_non_differentiable(o1, o2)
        return stats[0], stats[1], stats[4], stats[5], o1, o2

    @staticmethod
    def backward(ctx, g1, g2, *_):
        d1, d2 = ctx.saved_tensors
        return d1 * g1, d2 * g2, None, None, None


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
        loss_1, loss_2, pure_1, pure_2, o1, o2 = _CTLossFunction.apply(y_1.float(), y_2.float(), y_noise.view(-1),
                                                                       num_remember, clean_rows)
        o1, o2 = o1.long(), o2.long()
        return (loss_1, loss_2, pure_1, pure_2, o1[:num_remember], o2[:num_remember], o1[num_remember:], o2[num_remember:])
