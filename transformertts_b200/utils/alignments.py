"""Duration extraction from the Aligner's attention maps, mirroring the reference's ``utils/alignments.py`` (same function
names and argument meaning).  The attention scores (utils/metrics.py:5-44) and the shortest-monotonic-path search
(utils/alignments.py:58-91, scipy Dijkstra in the reference) run on the GPU: ``ttsb_attention_scores`` and
``ttsb_durations_from_attention`` (anti-diagonal dynamic programme in float64, csrc/alignment.cu)."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch

from .. import lib
from .spectrogram_ops import mel_lengths, phoneme_lengths


def duration_to_alignment_matrix(durations) -> np.ndarray:
    """utils/alignments.py:94-100: (phonemes, frames) 0/1 matrix with durations[i] ones in row i, one after the other."""
    durations = np.asarray(durations).astype(int)
    starts = np.cumsum(np.append([0], durations[:-1]))
    tot = int(np.sum(durations))
    out = np.zeros((len(durations), tot))
    for i, (s, d) in enumerate(zip(starts, durations)):
        out[i, s:s + d] = 1.0
    return out


def attention_score(att: torch.Tensor, mel_len: torch.Tensor, phon_len: torch.Tensor, r: int = 1):
    """utils/metrics.py:5-24 -> (loc_score, peak_score, 3 / diag_score), each (N, heads) float32 on the GPU."""
    att = att.to(dtype=torch.float32).contiguous()
    B, H = att.shape[:2]
    scores = torch.empty((B, H, 3), dtype=torch.float32, device=att.device)
    lib.attention_scores(att, mel_len.to(device=att.device, dtype=torch.int32).contiguous(),
                         phon_len.to(device=att.device, dtype=torch.int32).contiguous(), r, scores)
    return scores[..., 0], scores[..., 1], scores[..., 2]


def get_durations_from_alignment(batch_alignments, mels, phonemes, weighted: bool = False) -> Tuple[List[np.ndarray], None, torch.Tensor,
                                                                                                 torch.Tensor, torch.Tensor]:
    """utils/alignments.py:103-143.  batch_alignments: (N, heads, mel, phonemes) attention weights of the last decoder block
    (``Decoder_LastBlock_CrossAttention``); mels with start/end vectors, phonemes with start/end tokens.
    Returns (durations [list of int32 arrays of length phon_len - 1], None, jumpiness, peakiness, diag_measure); the second
    element is the reference's plotting matrix (best attention + binary alignment), which is not produced here.

    Ties: the reference runs scipy's Dijkstra on an explicit graph; the CUDA kernel runs the equivalent dynamic programme over
    anti-diagonals and breaks EXACT cost ties in a fixed order (left, up, diagonal), scipy by heap-pop order.  On generic
    attention maps the shortest path is unique and the durations are bit-identical (tests); on plateaus of exactly equal cost
    (saturated / all-zero attention regions) the two may pick different, equally short paths."""
    att = torch.as_tensor(batch_alignments)
    if not att.is_cuda:
        att = att.cuda()
    att = att.to(torch.float32).contiguous()
    dev = att.device
    B, H, Tq, Tk = att.shape
    mel_len = (mel_lengths(torch.as_tensor(mels).to(dev), padding_value=0.) - 1).to(torch.int32).contiguous()
    phon_len = (phoneme_lengths(torch.as_tensor(phonemes).to(dev)) - 1).to(torch.int32).contiguous()
    scores = torch.empty((B, H, 3), dtype=torch.float32, device=dev)
    lib.attention_scores(att, mel_len, phon_len, 1, scores)
    durations = torch.empty((B, Tk), dtype=torch.int32, device=dev)
    scratch = torch.empty((B, Tq * Tk), dtype=torch.uint8, device=dev)
    lib.durations_from_attention(att, mel_len, phon_len, scores, weighted, scratch, durations)
    d_host = durations.cpu().numpy()
    ml, pl = mel_len.cpu().numpy(), phon_len.cpu().numpy()
    out = []
    for b in range(B):
        d = d_host[b, :max(int(pl[b]) - 1, 0)].copy()
        if int(d.sum()) != int(ml[b]) - 1:   # same assertion as the reference (alignments.py:136)
            raise AssertionError(f'{int(d.sum())} vs {int(ml[b]) - 1}')
        out.append(d)
    return out, None, scores[..., 0], scores[..., 1], scores[..., 2]
