"""Length helpers on padded batches, mirroring the reference's ``utils/spectrogram_ops.py`` (same names and arguments);
the integer reductions run in the CUDA kernels ``ttsb_mel_lengths`` / ``ttsb_phoneme_lengths`` (bit-exact)."""
from __future__ import annotations

import torch

from .. import lib


def mel_padding_mask(mel_batch: torch.Tensor, padding_value=0) -> torch.Tensor:
    """1.0 where an element differs from the padding value (reference: utils/spectrogram_ops.py:4-5)."""
    return 1.0 - (mel_batch == padding_value).to(torch.float32)


def mel_lengths(mel_batch: torch.Tensor, padding_value=0) -> torch.Tensor:
    """Frames per row whose channel-sum of the padding mask differs from C*padding (reference: :8-13) -> int32 (B,)."""
    mel = mel_batch.to(dtype=torch.float32).contiguous()
    if not mel.is_cuda:
        mel = mel.cuda()
    out = torch.empty((mel.shape[0],), dtype=torch.int32, device=mel.device)
    lib.mel_lengths(mel, float(padding_value), out)
    return out


def phoneme_lengths(phonemes: torch.Tensor, phoneme_padding=0) -> torch.Tensor:
    """Tokens per row different from the padding id (reference: :16-17) -> int32 (B,)."""
    ph = phonemes.to(dtype=torch.int32).contiguous()
    if not ph.is_cuda:
        ph = ph.cuda()
    out = torch.empty((ph.shape[0],), dtype=torch.int32, device=ph.device)
    lib.phoneme_lengths(ph, int(phoneme_padding), out)
    return out
