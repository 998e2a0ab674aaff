"""Co-teaching loss oracle — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates ``CTLoss.forward`` of reference src/utils/losses.py:19-49 line by line in torch on the CPU (fp64-capable).
The one deliberate difference: ``np.argsort`` (quicksort, ties unspecified) is replaced by a stable argsort, the rule
the device kernel implements (ties by row index); with distinct losses the two agree."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def ct_loss(y_1, y_2, y_noise, forget_rate, ind=None, noise_or_not=None):
    loss_1 = F.cross_entropy(y_1, y_noise, reduction="none")            # losses.py:20
    ind_1_sorted = torch.argsort(loss_1.detach(), stable=True)           # losses.py:21 (np.argsort on the host)
    loss_2 = F.cross_entropy(y_2, y_noise, reduction="none")            # losses.py:24
    ind_2_sorted = torch.argsort(loss_2.detach(), stable=True)           # losses.py:25
    remember_rate = 1 - forget_rate                                      # losses.py:28
    num_remember = int(remember_rate * len(loss_1))                      # losses.py:29
    pure_1 = pure_2 = None
    if noise_or_not is not None and ind is not None:                     # losses.py:31-32
        pure_1 = noise_or_not[ind[ind_1_sorted[:num_remember]]].sum() / float(num_remember)
        pure_2 = noise_or_not[ind[ind_2_sorted[:num_remember]]].sum() / float(num_remember)
    ind_1_update, ind_2_update = ind_1_sorted[:num_remember], ind_2_sorted[:num_remember]     # losses.py:34-35
    ind_noisy_1, ind_noisy_2 = ind_1_sorted[num_remember:], ind_2_sorted[num_remember:]       # losses.py:42-43
    loss_1_update = F.cross_entropy(y_1[ind_2_update], y_noise[ind_2_update])                # losses.py:45 (exchange)
    loss_2_update = F.cross_entropy(y_2[ind_1_update], y_noise[ind_1_update])                # losses.py:46
    return loss_1_update, loss_2_update, pure_1, pure_2, ind_1_update, ind_2_update, ind_noisy_1, ind_noisy_2
