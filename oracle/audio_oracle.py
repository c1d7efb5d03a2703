"""CPU oracle for the STFT -> 80-bin mel -> log front-end (data/audio.py:72-92, 209-231).

TEST INFRASTRUCTURE ONLY (see oracle/forward_oracle.py header for the import rule).

PARITY UNPINNED: the arithmetic lives in librosa==0.7.1 (requirements.txt:2, not vendored, not installed
here).  This file restates the published librosa 0.7.1 algorithms used at the reference call sites
(``librosa.stft`` at data/audio.py:82-86; ``librosa.feature.melspectrogram(S=...)`` -> ``librosa.filters.mel``
at data/audio.py:73-79).  It is cross-checked in tests/test_oracle.py against two structurally different
implementations available in this image: ``torch.stft`` and ``torchaudio.functional.melscale_fbanks``.

librosa 0.7.1 semantics restated:
  stft(y, n_fft, hop_length, win_length): window='hann' (scipy get_window(..., fftbins=True) = periodic),
    center=True, pad_mode='reflect', dtype=complex64; frames = 1 + len(y)//hop.
  filters.mel(sr, n_fft, n_mels, fmin, fmax, htk=False, norm=1): Slaney mel scale, triangular filters,
    each row scaled by 2/(f[i+2]-f[i]); float32.
  feature.melspectrogram(S=S, ...): mel_basis @ S  (S used as given: the reference passes |D|, power is NOT applied).
"""
from __future__ import annotations

import numpy as np


def hann_periodic(n: int) -> np.ndarray:
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n))


def stft(y: np.ndarray, n_fft: int = 1024, hop_length: int = 256, win_length: int = 1024) -> np.ndarray:
    """librosa.stft as called at data/audio.py:82-86 -> complex64 (1+n_fft/2, 1+len(y)//hop)."""
    y = np.asarray(y, dtype=np.float32)
    win = hann_periodic(win_length)
    if win_length < n_fft:  # librosa pad_center
        lpad = (n_fft - win_length) // 2
        win = np.pad(win, (lpad, n_fft - win_length - lpad))
    win = win.astype(np.float32)
    ypad = np.pad(y, n_fft // 2, mode='reflect')
    n_frames = 1 + (len(ypad) - n_fft) // hop_length
    idx = np.arange(n_fft)[:, None] + hop_length * np.arange(n_frames)[None, :]
    frames = ypad[idx] * win[:, None]
    # numpy 1.17 (pinned by librosa 0.7.1) transforms in double precision and the result is stored as complex64
    return np.fft.rfft(frames.astype(np.float32).astype(np.float64), axis=0).astype(np.complex64)


def hz_to_mel(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mels)


def mel_to_hz(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def mel_filterbank(sr: int = 22050, n_fft: int = 1024, n_mels: int = 80, fmin: float = 0.0,
                   fmax: float = 8000.0) -> np.ndarray:
    """librosa.filters.mel(htk=False, norm=1) -> float32 (n_mels, 1+n_fft/2)."""
    n_bins = 1 + n_fft // 2
    fftfreqs = np.linspace(0, float(sr) / 2, n_bins, endpoint=True)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    weights = np.zeros((n_mels, n_bins))
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights.astype(np.float32)


def melgan_normalize(S: np.ndarray) -> np.ndarray:
    """data/audio.py:209-216 -- log(clip(S, 1e-5, None))."""
    return np.log(np.clip(S, a_min=1.0e-5, a_max=None))


def wavernn_normalize(S: np.ndarray) -> np.ndarray:
    """data/audio.py:222-231 -- clip((20*log10(max(1e-5,S)) + 100)/100, 0, 1)*8 - 4."""
    db = 20 * np.log10(np.maximum(1e-5, S))
    return np.clip((db - (-100)) / 100, 0, 1) * 2 * 4 - 4


def mel_spectrogram(wav: np.ndarray, sr: int = 22050, n_fft: int = 1024, hop_length: int = 256,
                    win_length: int = 1024, n_mels: int = 80, fmin: float = 0.0, fmax: float = 8000.0,
                    normalizer: str = 'MelGAN', _basis_cache: dict = {}) -> np.ndarray:
    """data/audio.py:88-92 -- (T, n_mels) float32."""
    key = (sr, n_fft, n_mels, fmin, fmax)
    if key not in _basis_cache:
        _basis_cache[key] = mel_filterbank(sr, n_fft, n_mels, fmin, fmax)
    D = stft(wav, n_fft, hop_length, win_length)
    S = np.dot(_basis_cache[key], np.abs(D))
    S = melgan_normalize(S) if normalizer == 'MelGAN' else wavernn_normalize(S)
    return S.T.astype(np.float32)


def make_clips(n_clips: int, n_samples: int, seed: int = 400) -> np.ndarray:
    """SURVEY 8d C4 inputs: 0.1*N(0,1) + three sinusoids (220/880/3520 Hz)."""
    rng = np.random.default_rng(seed)
    t = np.arange(n_samples, dtype=np.float64) / 22050.0
    base = 0.2 * np.sin(2 * np.pi * 220 * t) + 0.1 * np.sin(2 * np.pi * 880 * t) + 0.05 * np.sin(2 * np.pi * 3520 * t)
    x = 0.1 * rng.standard_normal((n_clips, n_samples)) + base[None, :]
    return x.astype(np.float32)


# utils/spectrogram_ops.py:4-17 restated in numpy -------------------------------------------------
def mel_lengths(mel_batch: np.ndarray, padding_value: float = 0) -> np.ndarray:
    mask = 1.0 - (mel_batch == padding_value).astype(np.float32)
    sum_tot = np.float32(mel_batch.shape[-1]) * padding_value
    idxs = (mask.sum(axis=-1) != sum_tot).astype(np.int32)
    return idxs.sum(axis=-1).astype(np.int32)


def phoneme_lengths(phonemes: np.ndarray, phoneme_padding: int = 0) -> np.ndarray:
    return (phonemes != phoneme_padding).astype(np.int32).sum(axis=-1).astype(np.int32)
