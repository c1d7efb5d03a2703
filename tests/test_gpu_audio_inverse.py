"""SURVEY 8(f) row 4 on the GPU: Audio.reconstruct_waveform (data/audio.py:94-110) -- mel inversion, complex STFT, iSTFT and the
Griffin-Lim loop against the numpy restatement of librosa 0.7.1 (oracle/audio_oracle.py), with a SHARED initial phase (the
reference draws it from numpy's global RNG, so a reference run is not reproducible either)."""
import numpy as np
import pytest
import torch

from oracle import audio_oracle as ao

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _audio():
    from transformertts_b200.data.audio import Audio
    return Audio(sampling_rate=22050, n_fft=1024, mel_channels=80, hop_length=256, win_length=1024, f_min=0, f_max=8000, normalizer='MelGAN')


@pytest.mark.parametrize('n', [11008, 5000, 22050])
def test_complex_stft_and_istft(n):
    from transformertts_b200 import lib
    y = ao.make_clips(1, n, seed=7)[0]
    T = 1 + n // 256
    spec = torch.full((T, 513, 2), float('nan'), device=DEV)
    lib.stft_complex(torch.from_numpy(y).to(DEV), spec)
    D = ao.stft(y)                                              # (513, T) complex64
    got = torch.view_as_complex(spec).cpu().numpy().T
    scale = np.abs(D).max()
    assert np.abs(got - D).max() < 2e-6 * scale + 1e-5
    # iSTFT of the oracle spectrum == oracle iSTFT; and the round trip reconstructs the signal
    ws = torch.empty(lib.istft_workspace_bytes(T) // 4, device=DEV)
    wav = torch.full((256 * (T - 1),), float('nan'), device=DEV)
    lib.istft(torch.view_as_real(torch.from_numpy(np.ascontiguousarray(D.T)).to(DEV)).contiguous(), ws, wav)
    want = ao.istft(D)
    assert np.abs(wav.cpu().numpy() - want).max() < 2e-6
    assert np.abs(wav.cpu().numpy() - y[:len(want)]).max() < 5e-6
    # a spectrum with imaginary DC / Nyquist parts: irfft ignores them
    D2 = D.copy()
    D2[0] += 0.3j
    D2[-1] -= 0.7j
    lib.istft(torch.view_as_real(torch.from_numpy(np.ascontiguousarray(D2.T)).to(DEV)).contiguous(), ws, wav)
    assert np.abs(wav.cpu().numpy() - ao.istft(D2)).max() < 2e-6


def test_mel_inversion_matches_oracle_solver_and_librosa_solver():
    a = _audio()
    y = ao.make_clips(1, 22050, seed=8)[0]
    mel = ao.mel_spectrogram(y)                                   # (T, 80)
    amp = np.exp(mel).astype(np.float32)
    got = a.mel_to_linear_device(torch.from_numpy(amp).to(DEV), n_iter=64).cpu().numpy().T     # (513, T)
    want = ao.mel_to_stft(amp.T, solver='pg', n_iter=64)
    assert got.min() >= 0
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-4
    lib_x = ao.mel_to_stft(amp.T, solver='lbfgsb')                # what librosa's L-BFGS-B returns from the same start
    assert np.linalg.norm(got - lib_x) / np.linalg.norm(lib_x) < 1e-2
    A = ao.mel_filterbank()
    assert 0.5 * np.sum((A @ got - amp.T) ** 2) <= 0.5 * np.sum((A @ lib_x - amp.T) ** 2) + 1e-6


@pytest.mark.parametrize('n_iter', [0, 4, 32])
def test_griffinlim_with_shared_initial_phase(n_iter):
    a = _audio()
    y = ao.make_clips(1, 16000, seed=9)[0]
    mel = ao.mel_spectrogram(y)
    S = ao.mel_to_stft(np.exp(mel).T.astype(np.float32), solver='pg', n_iter=64)               # (513, T)
    init = np.exp(2j * np.pi * np.random.default_rng(11).random(S.shape)).astype(np.complex64)
    want = ao.griffinlim(S, n_iter=n_iter, init_angles=init)
    got = a.griffinlim_device(torch.from_numpy(np.ascontiguousarray(S.T)).to(DEV), n_iter=n_iter,
                              init_angles=torch.from_numpy(np.ascontiguousarray(init.T))).cpu().numpy()
    assert got.shape == want.shape
    # the iteration is a fixed-point map: fp32 FFT rounding (the reference transforms in double) grows slowly with n_iter
    tol = {0: 5e-6, 4: 2e-4, 32: 5e-3}[n_iter]
    assert np.abs(got - want).max() < tol * max(1.0, np.abs(want).max()), np.abs(got - want).max()
    # and the result is as consistent with the target magnitudes as the oracle's
    e_got = np.linalg.norm(np.abs(ao.stft(got)) - S) / np.linalg.norm(S)
    e_want = np.linalg.norm(np.abs(ao.stft(want)) - S) / np.linalg.norm(S)
    assert e_got < 1.02 * e_want + 1e-4


def test_reconstruct_waveform_api():
    """reference API: mel (n_mels, T) normalised in, waveform out (predict_tts.py:57)."""
    a = _audio()
    y = ao.make_clips(1, 16000, seed=10)[0]
    mel = a.mel_spectrogram(y)                                    # (T, 80) through the CUDA front end
    wav = a.reconstruct_waveform(mel.T, n_iter=32, seed=5)
    assert wav.dtype == np.float32 and wav.shape == (256 * (mel.shape[0] - 1),) and np.isfinite(wav).all()
    mel2 = a.mel_spectrogram(wav)                                 # re-analysis is close to the input mel in the log domain
    assert np.abs(mel2[2:-2] - mel[2:mel2.shape[0] - 2]).mean() < 0.35
    assert np.array_equal(a.reconstruct_waveform(mel.T, n_iter=4, seed=5), a.reconstruct_waveform(mel.T, n_iter=4, seed=5))
