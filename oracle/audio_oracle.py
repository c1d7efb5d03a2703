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


# -------------------------------------------------------------------------------------------------------------------
# mel -> waveform (data/audio.py:94-110: librosa.feature.inverse.mel_to_stft + librosa.core.griffinlim, librosa 0.7.1)
# -------------------------------------------------------------------------------------------------------------------
def melgan_denormalize(S: np.ndarray) -> np.ndarray:
    """data/audio.py:218-219."""
    return np.exp(S)


def wavernn_denormalize(S: np.ndarray) -> np.ndarray:
    """data/audio.py:233-236 -- S = (S + 4) / 8; 10 ** (((clip(S, 0, 1) * 100) - 100) / 20)."""
    S = (S + 4) / 8
    return np.power(10.0, ((np.clip(S, 0, 1) * 100) - 100) * 0.05)


def window_sumsquare(n_frames: int, n_fft: int = 1024, hop_length: int = 256, win_length: int = 1024) -> np.ndarray:
    """librosa.filters.window_sumsquare(window='hann', norm=None): sum of the squared window over the overlapping frames."""
    n = n_fft + hop_length * (n_frames - 1)
    x = np.zeros(n, dtype=np.float32)
    win_sq = hann_periodic(win_length).astype(np.float64) ** 2
    if win_length < n_fft:
        lpad = (n_fft - win_length) // 2
        win_sq = np.pad(win_sq, (lpad, n_fft - win_length - lpad))
    for i in range(n_frames):
        s = i * hop_length
        x[s:min(n, s + n_fft)] += win_sq[:max(0, min(n_fft, n - s))]
    return x


def istft(D: np.ndarray, hop_length: int = 256, win_length: int = 1024) -> np.ndarray:
    """librosa.istft(center=True, window='hann', dtype=float32, length=None) for a (1+n_fft/2, n_frames) complex matrix:
    irfft (double precision, as numpy does) * window, overlap-add, division by the window sum-square where it exceeds
    tiny(float32), centre trimming."""
    n_fft = 2 * (D.shape[0] - 1)
    n_frames = D.shape[1]
    win = hann_periodic(win_length)
    if win_length < n_fft:
        lpad = (n_fft - win_length) // 2
        win = np.pad(win, (lpad, n_fft - win_length - lpad))
    y = np.zeros(n_fft + hop_length * (n_frames - 1), dtype=np.float32)
    ytmp = win[:, None] * np.fft.irfft(D.astype(np.complex128), axis=0)
    for f in range(n_frames):
        s = f * hop_length
        y[s:s + n_fft] += ytmp[:, f]
    wss = window_sumsquare(n_frames, n_fft, hop_length, win_length)
    nz = wss > np.finfo(np.float32).tiny
    y[nz] /= wss[nz]
    return y[n_fft // 2:-(n_fft // 2)]


def griffinlim(S: np.ndarray, n_iter: int = 32, hop_length: int = 256, win_length: int = 1024, momentum: float = 0.99,
               init_angles: np.ndarray = None, seed: int = 0) -> np.ndarray:
    """librosa.griffinlim (0.7.1: "fast" Griffin-Lim with momentum 0.99, random initial phase).  The reference draws the
    initial phase from numpy's global RNG (not reproducible across runs); here it is an argument (unit-modulus complex
    (bins, frames)) or drawn from default_rng(seed), so that the CUDA path can be compared on identical input."""
    n_fft = 2 * (S.shape[0] - 1)
    if init_angles is None:
        init_angles = np.exp(2j * np.pi * np.random.default_rng(seed).random(S.shape))
    angles = init_angles.astype(np.complex64)
    rebuilt = 0.0
    for _ in range(n_iter):
        tprev = rebuilt
        inverse = istft(S * angles, hop_length, win_length)
        rebuilt = stft(inverse, n_fft, hop_length, win_length)
        angles = (rebuilt - (momentum / (1 + momentum)) * tprev).astype(np.complex64)
        angles = (angles / (np.abs(angles) + 1e-16)).astype(np.complex64)
    return istft(S * angles, hop_length, win_length)


def nnls_lbfgsb(A: np.ndarray, B: np.ndarray) -> np.ndarray:
    """librosa.util.nnls for a matrix right-hand side (0.7.1: _nnls_lbfgs_block): minimise 0.5 * ||A x - B||^2 over x >= 0
    with scipy's L-BFGS-B (m = A.shape[1] corrections), started from the clipped least-squares solution."""
    import scipy.optimize
    A64, B64 = A.astype(np.float64), B.astype(np.float64)
    x_init = np.linalg.lstsq(A64, B64, rcond=None)[0]
    np.clip(x_init, 0, None, out=x_init)
    shape = x_init.shape

    def obj(x):
        x = x.reshape(shape)
        diff = A64 @ x - B64
        return 0.5 * np.sum(diff ** 2), (A64.T @ diff).ravel()

    x, _, _ = scipy.optimize.fmin_l_bfgs_b(obj, x_init.ravel(), bounds=[(0, None)] * x_init.size, m=A.shape[1])
    return x.reshape(shape)


def nnls_projected_gradient(A: np.ndarray, B: np.ndarray, n_iter: int = 64) -> np.ndarray:
    """The solver the CUDA path uses (transformertts_b200/csrc/griffin_lim.cu): same objective and the same start (clipped
    pseudo-inverse solution), FISTA-accelerated projected gradient with the fixed step 1 / ||A||_2^2 and a fixed iteration
    count -- deterministic and data-parallel per frame.  It reaches the NNLS optimum like L-BFGS-B; the minimiser itself is
    not unique (80 equations, 513 unknowns per frame), so the two solvers agree in the objective and in A x, not bit-wise in x."""
    A32 = A.astype(np.float32)
    pinv = np.linalg.pinv(A.astype(np.float64)).astype(np.float32)
    step = np.float32(1.0 / np.linalg.norm(A.astype(np.float64), 2) ** 2)
    B32 = B.astype(np.float32)
    x = np.maximum(pinv @ B32, 0).astype(np.float32)
    yk, tk = x.copy(), np.float32(1.0)
    for _ in range(n_iter):
        g = A32.T @ (A32 @ yk - B32)
        x_new = np.maximum(yk - step * g, 0).astype(np.float32)
        t_new = np.float32((1 + np.sqrt(1 + 4 * tk * tk)) / 2)
        yk = (x_new + ((tk - 1) / t_new) * (x_new - x)).astype(np.float32)
        x, tk = x_new, t_new
    return x


def mel_to_stft(M: np.ndarray, sr: int = 22050, n_fft: int = 1024, fmin: float = 0.0, fmax: float = 8000.0, solver: str = 'lbfgsb',
                n_iter: int = 64) -> np.ndarray:
    """librosa.feature.inverse.mel_to_stft(M (n_mels, T), power=1): non-negative least squares against the mel basis."""
    basis = mel_filterbank(sr, n_fft, M.shape[0], fmin, fmax)
    if solver == 'lbfgsb':
        return nnls_lbfgsb(basis, M).astype(np.float32)
    return nnls_projected_gradient(basis, M, n_iter)


def reconstruct_waveform(mel: np.ndarray, n_iter: int = 32, normalizer: str = 'MelGAN', solver: str = 'lbfgsb', nnls_iter: int = 64,
                         init_angles: np.ndarray = None, seed: int = 0) -> np.ndarray:
    """data/audio.py:94-110 -- mel (n_mels, T) normalised -> waveform."""
    amp = melgan_denormalize(mel) if normalizer == 'MelGAN' else wavernn_denormalize(mel)
    S = mel_to_stft(amp.astype(np.float32), solver=solver, n_iter=nnls_iter)
    return griffinlim(S, n_iter=n_iter, init_angles=init_angles, seed=seed)
