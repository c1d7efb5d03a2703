"""Losses of the ForwardTransformer training step, mirroring the reference's ``utils/losses.py`` for the hot path:
``masked_mean_absolute_error`` exactly as it is CALLED there (mask=None: the mean runs over all elements, padding
included -- utils/losses.py:41-49) and ``weighted_sum_losses`` (:63-70).  Values come from the CUDA
kernel ``ttsb_mae_loss``; the cross-entropy variants belong to the Aligner and are out of scope."""
from __future__ import annotations

from typing import Sequence

import torch

from .. import lib


def masked_mean_absolute_error(targets: torch.Tensor, logits: torch.Tensor, mask_value=0, mask=None) -> torch.Tensor:
    """mean |targets - logits| over every element.  As in the reference, the mask branch only runs when ``mask`` is given,
    which the training step never does; passing one is rejected instead of silently changing the loss."""
    if mask is not None:
        raise NotImplementedError('the reference never passes a mask on this path (utils/losses.py:63-70)')
    pred = logits.to(dtype=torch.float32).contiguous()
    if not pred.is_cuda:
        pred = pred.cuda()
    tgt = targets.to(pred.device)
    tgt = tgt.to(torch.int32).contiguous() if not tgt.is_floating_point() else tgt.to(torch.float32).contiguous()
    if pred.dim() == 2:
        pred, tgt = pred[..., None], tgt.reshape(*pred.shape, 1)
    B, Tp, C = pred.shape
    Tt = tgt.shape[1]
    loss = torch.zeros(1, dtype=torch.float32, device=pred.device)
    lib.mae_loss(pred, B, Tp, Tt, C, tgt, 1.0, loss, None)
    return loss[0]


def weighted_sum_losses(targets: Sequence[torch.Tensor], pred: Sequence[torch.Tensor], loss_functions, coeffs):
    """reference: utils/losses.py:63-70."""
    total, vals = 0, []
    for i in range(len(loss_functions)):
        v = loss_functions[i](targets[i], pred[i])
        vals.append(v)
        total = total + coeffs[i] * v
    return total, vals
