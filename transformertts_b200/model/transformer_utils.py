"""Host-side mirror of model/transformer_utils.py of the reference (torch tensors instead of tf tensors)."""
from __future__ import annotations

import numpy as np
import torch


def positional_encoding(position: int, model_dim: int) -> torch.Tensor:
    """Sinusoidal table (1, position, model_dim), built in float64 and cast to float32
    (reference: model/transformer_utils.py:5-21)."""
    pos = np.arange(position, dtype=np.float64)[:, None]
    i = np.arange(model_dim)[None, :]
    rates = 1.0 / np.power(10000.0, (2 * (i // 2)) / np.float32(model_dim))
    ang = pos * rates
    table = np.empty_like(ang)
    table[:, 0::2] = np.sin(ang[:, 0::2])
    table[:, 1::2] = np.cos(ang[:, 1::2])
    return torch.from_numpy(table[None].astype(np.float32))


def create_encoder_padding_mask(seq: torch.Tensor) -> torch.Tensor:
    """(B,1,1,T) float mask, 1.0 where the token id is 0 (reference: model/transformer_utils.py:24-26)."""
    return (seq == 0).to(torch.float32)[:, None, None, :]


def create_mel_padding_mask(seq: torch.Tensor) -> torch.Tensor:
    """(B,1,1,T) float mask, 1.0 where a frame is all zeros (reference: model/transformer_utils.py:29-32)."""
    return (seq.abs().sum(dim=-1) == 0).to(torch.float32)[:, None, None, :]


def create_look_ahead_mask(size: int) -> torch.Tensor:
    """Strictly-upper-triangular ones (reference: model/transformer_utils.py:35-37)."""
    return 1 - torch.tril(torch.ones(size, size))


def mask_from_lengths(lengths: torch.Tensor, T: int) -> torch.Tensor:
    """(B,1,1,T) float padding mask from per-row valid lengths (equals the value-derived masks above on batches
    padded at the end)."""
    ar = torch.arange(T, device=lengths.device)
    return (ar[None, :] >= lengths[:, None]).to(torch.float32)[:, None, None, :]
