"""CPU tests that pin the oracle (oracle/) -- the checker every GPU parity test relies on.

The reference (TensorFlow / librosa) cannot run here and its own tests hold no vectors for this path, so the oracle is
pinned by (a) the one known-answer example the reference contains (Expand docstring), (b) an independent
implementation of every layer from stock torch.nn.functional / torch.stft / torchaudio ops, (c) the quirks listed in
SURVEY.md App. A, and (d) the committed golden vectors.
"""
import math
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import audio_oracle as ao
from oracle import forward_oracle as fo

GOLD = Path(__file__).resolve().parent / 'golden'
torch.set_num_threads(4)


# ----------------------------------------------------------------------------------------------------------
# (a) known-answer material inside the reference
# ----------------------------------------------------------------------------------------------------------
def test_expand_docstring_example():
    """model/layers.py:532-542."""
    x = torch.tensor([[[0.54710746, 0.8943467], [0.7140938, 0.97968304], [0.5347662, 0.15213418]]])
    dims = torch.tensor([[[1.], [3.], [2.]]])
    out = fo.expand(x, dims)
    want = torch.tensor([[[0.54710746, 0.8943467], [0.7140938, 0.97968304], [0.7140938, 0.97968304],
                          [0.7140938, 0.97968304], [0.5347662, 0.15213418], [0.5347662, 0.15213418]]])
    assert out.shape == (1, 6, 2)
    assert torch.equal(out, want)


def test_expand_matches_repeat_interleave_and_pads_rows():
    g = torch.Generator().manual_seed(3)
    x = torch.randn(4, 9, 6, generator=g)
    d = torch.randint(0, 5, (4, 9), generator=g)
    out = fo.expand(x, d[..., None].float())
    lengths, idx = fo.expand_indices(d.to(torch.int32))
    assert out.shape[1] == int(lengths.max())
    for b in range(4):
        ref = torch.repeat_interleave(x[b], d[b], dim=0)
        assert torch.equal(out[b, :ref.shape[0]], ref)
        assert torch.count_nonzero(out[b, ref.shape[0]:]) == 0
        assert torch.equal(idx[b, :ref.shape[0]].long(), torch.repeat_interleave(torch.arange(9), d[b]))
        assert (idx[b, ref.shape[0]:] == -1).all()


def test_round_half_to_even():
    """tf.math.round (model/layers.py:551) is banker's rounding."""
    v = torch.tensor([[[0.5], [1.5], [2.5], [3.5], [2.4999], [2.5001], [0.0]]])
    assert fo.round_durations(v).tolist() == [[0, 2, 2, 4, 2, 3, 0]]


# ----------------------------------------------------------------------------------------------------------
# (b) independent implementation from stock torch ops
# ----------------------------------------------------------------------------------------------------------
def _indep_block(p, pre, x, key_pad, kind, nh):
    """Same block from F.scaled_dot_product_attention / F.conv1d / F.layer_norm."""
    B, T, d = x.shape
    q = F.linear(x, p[pre + 'wq.w'].T, p[pre + 'wq.b']).view(B, T, nh, d // nh).transpose(1, 2)
    k = F.linear(x, p[pre + 'wk.w'].T, p[pre + 'wk.b']).view(B, T, nh, d // nh).transpose(1, 2)
    v = F.linear(x, p[pre + 'wv.w'].T, p[pre + 'wv.b']).view(B, T, nh, d // nh).transpose(1, 2)
    attn_mask = torch.zeros(B, 1, 1, T).masked_fill(key_pad[:, None, None, :], float('-inf'))
    a = F.scaled_dot_product_attention(q, k, v, attn_mask=attn_mask).transpose(1, 2).reshape(B, T, d)
    o = F.linear(torch.cat([x, a], -1), p[pre + 'wo.w'].T, p[pre + 'wo.b'])
    keep = (~key_pad)[..., None].float()
    y = F.layer_norm(o + x, (d,), p[pre + 'ln1.gamma'], p[pre + 'ln1.beta'], eps=1e-6) * keep
    if kind == 'dense':
        h = F.linear(torch.relu(F.linear(y, p[pre + 'ffn1.w'].T, p[pre + 'ffn1.b'])), p[pre + 'ffn2.w'].T, p[pre + 'ffn2.b'])
    else:
        h = y.transpose(1, 2)
        n = sum(1 for kk in p if kk.startswith(pre + 'conv') and kk.endswith('.w'))
        for j in range(n):
            w = p[pre + f'conv{j}.w'].permute(2, 1, 0)  # (k,in,out) -> (out,in,k)
            h = F.conv1d(h, w, p[pre + f'conv{j}.b'], padding=w.shape[-1] // 2)
            if j < n - 1:
                h = torch.relu(h)
        h = h.transpose(1, 2)
    return F.layer_norm(h + y, (d,), p[pre + 'ln2.gamma'], p[pre + 'ln2.beta'], eps=1e-6) * keep


def _indep_forward(p, cfg, tokens, durs, pitch):
    B, Tp = tokens.shape
    d = cfg['encoder_model_dimension']
    pad = tokens == 0
    x = F.embedding(tokens.long(), p['embedding'])
    x = F.layer_norm(x, (d,), p['encoder.ln.gamma'], p['encoder.ln.beta'], eps=1e-6)
    x = x + p['encoder.pos_scalar'] * fo.positional_encoding(cfg['encoder_max_position_encoding'], d)[:, :Tp]
    for i, nh in enumerate(cfg['encoder_num_heads']):
        x = _indep_block(p, f'encoder.b{i}.', x, pad, 'dense' if i < cfg['encoder_dense_blocks'] else 'conv', nh)

    def predictor(name, relu_out):
        h = (x * (~pad)[..., None]).transpose(1, 2)
        for j in range(2):
            w = p[f'{name}.conv{j}.w'].permute(2, 1, 0)
            h = torch.relu(F.conv1d(h, w, p[f'{name}.conv{j}.b'], padding=1))
            h = F.layer_norm(h.transpose(1, 2), (w.shape[0],), p[f'{name}.ln{j}.gamma'], p[f'{name}.ln{j}.beta'], eps=1e-6).transpose(1, 2)
        o = F.linear(h.transpose(1, 2), p[f'{name}.out.w'].T, p[f'{name}.out.b'])
        return (torch.relu(o) if relu_out else o) * (~pad)[..., None]

    dur_pred, pitch_pred = predictor('dur_pred', True), predictor('pitch_pred', False)
    x = x + torch.relu(pitch[..., None] * p['pitch_embed.w'][0] + p['pitch_embed.b'])
    lens = durs.sum(1)
    Tm = int(lens.max())
    m = torch.zeros(B, Tm, d)
    for b in range(B):
        m[b, :lens[b]] = torch.repeat_interleave(x[b], durs[b].long(), dim=0)
    mpad = torch.arange(Tm)[None] >= lens[:, None]
    m = F.layer_norm(m, (d,), p['decoder.ln.gamma'], p['decoder.ln.beta'], eps=1e-6)
    m = m + p['decoder.pos_scalar'] * fo.positional_encoding(cfg['decoder_max_position_encoding'], d)[:, :Tm]
    for i, nh in enumerate(cfg['decoder_num_heads']):
        m = _indep_block(p, f'decoder.b{i}.', m, mpad, 'dense' if i < cfg['decoder_dense_blocks'] else 'conv', nh)
    return F.linear(m, p['out.w'].T, p['out.b']), dur_pred, pitch_pred


@pytest.mark.parametrize('kind,B,Tp,Tm', [('full', 1, 32, 250), ('ragged', 3, 24, 120)])
def test_oracle_matches_independent_torch_implementation(kind, B, Tp, Tm):
    cfg = fo.CONFIGS['C1']
    p = fo.init_params(cfg, seed=7)
    tok, dur, pit = fo.make_inputs(kind, B, Tp, Tm, seed=11)
    out = fo.forward_transformer_call(p, cfg, tok, dur[..., None], pit[..., None])
    mel2, d2, p2 = _indep_forward(p, cfg, tok, dur, pit)
    assert out['mel'].shape == mel2.shape
    assert (out['mel'] - mel2).abs().max() < 2e-4
    assert (out['duration'] - d2).abs().max() < 1e-4
    assert (out['pitch'] - p2).abs().max() < 1e-4


def test_float64_oracle_agrees_with_float32():
    cfg = fo.CONFIGS['C1']
    p = fo.init_params(cfg, seed=7)
    p64 = {k: v.double() for k, v in p.items()}
    tok, dur, pit = fo.make_inputs('ragged', 2, 24, 100, seed=12)
    a = fo.forward_transformer_call(p, cfg, tok, dur[..., None], pit[..., None])['mel']
    b = fo.forward_transformer_call(p64, cfg, tok, dur[..., None], pit[..., None].double())['mel']
    assert (a.double() - b).abs().max() < 2e-4


# ----------------------------------------------------------------------------------------------------------
# (c) quirks of the reference the oracle must reproduce
# ----------------------------------------------------------------------------------------------------------
def test_concat_projection_is_2d_to_d():
    """MHA output Dense takes concat([q_in, attention]) (model/layers.py:148-149)."""
    p = fo.init_params(fo.CONFIGS['C1'])
    assert p['encoder.b0.wo.w'].shape == (256, 128)


def test_padded_frames_equal_output_bias():
    """Inputs are zeroed at padded frames before the final Dense (layers.py:264, models.py:543)."""
    cfg = fo.CONFIGS['C1']
    p = fo.init_params(cfg, seed=7)
    tok, dur, pit = fo.make_inputs('ragged', 3, 24, 100, seed=5)
    out = fo.forward_transformer_call(p, cfg, tok, dur[..., None], pit[..., None])
    lens = dur.sum(1)
    b = int(torch.argmin(lens))
    assert lens[b] < out['mel'].shape[1]
    assert torch.allclose(out['mel'][b, lens[b]:], p['out.b'].expand(out['mel'].shape[1] - int(lens[b]), -1), atol=1e-6)
    # value-derived mask == length-derived mask
    want = (torch.arange(out['mel'].shape[1])[None] >= lens[:, None]).float()
    assert torch.equal(out['expanded_mask'][:, 0, 0], want)


def test_conv_halo_leak_on_padded_batches():
    """Stacked 'same' convs: the last valid frame of a padded sample sees relu(conv1) of the first padded frame
    (SURVEY App. A.4), so a sample alone and the same sample inside a longer batch differ at that frame only."""
    g = torch.Generator().manual_seed(0)
    w1, b1 = torch.randn(3, 4, 8, generator=g), torch.randn(8, generator=g)
    w2, b2 = torch.randn(3, 8, 4, generator=g), torch.randn(4, generator=g)
    x = torch.randn(1, 10, 4, generator=g)
    alone = fo.conv1d_same(torch.relu(fo.conv1d_same(x, w1, b1)), w2, b2)
    padded = torch.cat([x, torch.zeros(1, 5, 4)], 1)
    inside = fo.conv1d_same(torch.relu(fo.conv1d_same(padded, w1, b1)), w2, b2)[:, :10]
    assert torch.allclose(alone[:, :9], inside[:, :9], atol=1e-6)
    assert (alone[:, 9] - inside[:, 9]).abs().max() > 1e-3


def test_mae_is_unmasked_and_weighted():
    """utils/losses.py:41-49 with mask=None: plain mean over all elements; weights [1,1,3] (models.py:485)."""
    t = torch.tensor([[[1.0, 0.0], [0.0, 0.0]]])
    q = torch.tensor([[[0.0, 1.0], [2.0, 0.0]]])
    assert fo.masked_mean_absolute_error(t, q).item() == pytest.approx(1.0)
    total, vals = fo.weighted_sum_losses((t, t, t), (q, q, q))
    assert total.item() == pytest.approx(5.0)


def test_adam_keras_epsilon_placement():
    """theta -= lr*sqrt(1-b2^t)/(1-b1^t) * m/(sqrt(v)+eps): differs from torch.optim.Adam's eps placement."""
    p, g = torch.tensor([1.0]), torch.tensor([1e-6])
    m, v = torch.zeros(1), torch.zeros(1)
    fo.adam_tf_step(p, g, m, v, step=1, lr=1e-3)
    lr_t = 1e-3 * math.sqrt(1 - 0.98) / (1 - 0.9)
    want = 1.0 - lr_t * (0.1 * 1e-6) / (math.sqrt(0.02 * 1e-12) + 1e-9)
    assert p.item() == pytest.approx(want, rel=1e-6)


def test_model_sizes_and_flops_match_survey():
    n = sum(v.numel() for v in fo.init_params(fo.CONFIGS['LJ256']).values())
    assert abs(n / 1e6 - 23.64) < 0.01
    fl = fo.forward_flops(fo.CONFIGS['LJ256'], [128] * 64, [1000] * 64)
    assert abs(fl / 1e9 - 2060.9) < 0.5


def test_attention_dict_keys_follow_reference_names():
    cfg = fo.CONFIGS['C1']
    p = fo.init_params(cfg)
    tok, dur, pit = fo.make_inputs('full', 1, 8, 20, seed=1)
    out = fo.forward_transformer_call(p, cfg, tok, dur[..., None], pit[..., None])
    assert list(out['encoder_attention']) == ['Encoder_DenseBlock1_SelfAttention', 'Encoder_ConvBlock1_SelfAttention']
    assert list(out['decoder_attention']) == ['Decoder_DenseBlock1_SelfAttention', 'Decoder_ConvBlock1_SelfAttention']
    assert out['decoder_attention']['Decoder_ConvBlock1_SelfAttention'].shape == (1, 2, 20, 20)


# ----------------------------------------------------------------------------------------------------------
# audio front-end
# ----------------------------------------------------------------------------------------------------------
def test_stft_matches_torch_stft():
    y = ao.make_clips(1, 8000, seed=1)[0]
    D = ao.stft(y)
    T = torch.stft(torch.from_numpy(y), n_fft=1024, hop_length=256, win_length=1024, window=torch.hann_window(1024, periodic=True),
                   center=True, pad_mode='reflect', return_complex=True).numpy()
    assert D.shape == T.shape == (513, 1 + 8000 // 256)
    assert np.abs(np.abs(D) - np.abs(T)).max() < 1e-4


def test_mel_filterbank_matches_torchaudio_slaney():
    torchaudio = pytest.importorskip('torchaudio')
    fb = ao.mel_filterbank()
    ta = torchaudio.functional.melscale_fbanks(513, 0.0, 8000.0, 80, 22050, norm='slaney', mel_scale='slaney').T.numpy()
    assert fb.shape == (80, 513)
    assert np.abs(fb - ta).max() < 1e-6
    assert int((fb != 0).sum()) == 727 and int((fb != 0).sum(1).max()) == 27


def test_mel_spectrogram_shape_and_frame_count():
    y = ao.make_clips(1, 22050, seed=2)[0]
    m = ao.mel_spectrogram(y)
    assert m.shape == (1 + 22050 // 256, 80) and m.dtype == np.float32
    assert m.min() >= np.log(1e-5) - 1e-6
    w = ao.mel_spectrogram(y, normalizer='WaveRNN')
    assert w.min() >= -4 and w.max() <= 4


def test_length_helpers():
    mel = np.zeros((2, 5, 4), np.float32)
    mel[0, :3] = 1.0
    mel[1, :5, 2] = -2.0
    assert ao.mel_lengths(mel).tolist() == [3, 5]
    assert ao.phoneme_lengths(np.array([[3, 4, 0, 0], [1, 1, 1, 1]])).tolist() == [2, 4]


# ----------------------------------------------------------------------------------------------------------
# (d) golden vectors
# ----------------------------------------------------------------------------------------------------------
def test_golden_forward_vectors():
    g = np.load(GOLD / 'c1_forward.npz')
    cfg = fo.CONFIGS['C1']
    p = fo.init_params(cfg, seed=7)
    tok, dur, pit = fo.make_inputs('full', 1, 32, 250, seed=100)
    assert np.array_equal(tok.numpy(), g['tokens']) and np.array_equal(dur.numpy(), g['durations'])
    out = fo.forward_transformer_call(p, cfg, tok, dur[..., None], pit[..., None])
    assert np.abs(out['mel'].numpy() - g['mel']).max() < 1e-4
    pred = fo.predict(p, cfg, tok)
    assert np.array_equal(pred['int_durations'].numpy(), g['pred_int_durations'])
    assert np.abs(pred['mel'].numpy() - g['pred_mel']).max() < 1e-4


def test_golden_audio_vectors():
    g = np.load(GOLD / 'audio_mel.npz')
    clips = ao.make_clips(2, int(g['n_samples']), seed=int(g['clips_seed']))
    mels = np.stack([ao.mel_spectrogram(c) for c in clips])
    assert mels.shape == g['mel'].shape == (2, 1 + 11008 // 256, 80)
    assert np.abs(mels - g['mel']).max() < 1e-5


# ----------------------------------------------------------------------------------------------------------
# (e) mel -> waveform oracle (data/audio.py:94-110): istft against torch.istft, NNLS solvers against each other
# ----------------------------------------------------------------------------------------------------------
def test_istft_restatement_matches_torch_istft_and_inverts_stft():
    y = ao.make_clips(1, 11008, seed=5)[0]
    D = ao.stft(y)
    yi = ao.istft(D)
    assert yi.shape == (256 * (D.shape[1] - 1),)
    assert np.abs(yi - y[:len(yi)]).max() < 1e-6                      # perfect reconstruction (Hann, hop = n_fft / 4)
    yt = torch.istft(torch.from_numpy(D), 1024, 256, 1024, torch.hann_window(1024, periodic=True), center=True).numpy()
    assert np.abs(yt - yi[:len(yt)]).max() < 1e-6
    wss = ao.window_sumsquare(D.shape[1])
    assert abs(wss[2048] - 1.5) < 1e-6                                # sum of four shifted squared Hann windows


def test_nnls_solvers_agree_and_griffinlim_converges():
    """librosa's nnls (L-BFGS-B from the clipped least-squares start) vs the fixed-count FISTA projected gradient the CUDA path
    runs: same objective value (both essentially exact), solutions within 1 % of each other; Griffin-Lim with a shared
    initial phase reproduces a spectrogram consistent with the target magnitudes."""
    y = ao.make_clips(1, 11008, seed=6)[0]
    mel = ao.mel_spectrogram(y).T
    amp = np.exp(mel).astype(np.float32)
    A = ao.mel_filterbank()
    x1 = ao.mel_to_stft(amp, solver='lbfgsb')
    x2 = ao.mel_to_stft(amp, solver='pg', n_iter=64)
    assert x1.min() >= 0 and x2.min() >= 0
    o1, o2 = 0.5 * np.sum((A @ x1 - amp) ** 2), 0.5 * np.sum((A @ x2 - amp) ** 2)
    assert o2 <= o1 + 1e-6 and o2 < 1e-6 * 0.5 * np.sum(amp ** 2)
    assert np.linalg.norm(x1 - x2) / np.linalg.norm(x1) < 1e-2
    w = ao.griffinlim(x2, n_iter=16, seed=3)
    assert w.shape == (256 * (x2.shape[1] - 1),) and np.isfinite(w).all()
    err = np.linalg.norm(np.abs(ao.stft(w)) - x2) / np.linalg.norm(x2)   # spectral convergence after 16 iterations
    err0 = np.linalg.norm(np.abs(ao.stft(ao.griffinlim(x2, n_iter=0, seed=3))) - x2) / np.linalg.norm(x2)
    assert err < 0.5 * err0
