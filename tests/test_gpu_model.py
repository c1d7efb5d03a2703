"""GPU parity of the whole text->mel path through the reference-facing API (ForwardTransformer.call / predict)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import forward_oracle as fo

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / 'golden'
MEL_TOL = 1e-3  # BASELINE.json north_star: "mel fp32 within 1e-3 abs"


def _model(cfg_name, params, **kw):
    from transformertts_b200.model.models import ForwardTransformer
    m = ForwardTransformer(**fo.CONFIGS[cfg_name], **kw)
    m.set_weights(params)
    return m


@pytest.mark.parametrize('impl', ['simt', 'tcgen05'])
def test_c1_golden_forced_durations(impl):
    g = np.load(GOLD / 'c1_forward.npz')
    p = fo.init_params(fo.CONFIGS['C1'], seed=7)
    m = _model('C1', p, impl=impl)
    out = m.call(torch.from_numpy(g['tokens']), target_durations=torch.from_numpy(g['durations'])[..., None],
                 target_pitch=torch.from_numpy(g['pitch'])[..., None])
    assert out['mel'].shape == (1, 250, 80)
    assert np.abs(out['mel'].cpu().numpy() - g['mel']).max() < MEL_TOL
    assert np.abs(out['duration'].cpu().numpy() - g['duration_pred']).max() < 1e-3
    assert np.abs(out['pitch'].cpu().numpy() - g['pitch_pred']).max() < 1e-3
    assert np.array_equal(out['int_durations'].cpu().numpy(), g['durations'])


def test_c1_golden_ragged_batch_and_padding_quirks():
    """c1_forward.npz is written by the reference's own code (tests/golden/make_golden_ref.py)."""
    g = np.load(GOLD / 'c1_forward.npz')
    p = fo.init_params(fo.CONFIGS['C1'], seed=7)
    m = _model('C1', p)
    out = m.call(torch.from_numpy(g['tokens3']), target_durations=torch.from_numpy(g['durations3']),
                 target_pitch=torch.from_numpy(g['pitch3']))
    mel = out['mel'].cpu().numpy()
    assert mel.shape == g['mel3'].shape
    assert np.abs(mel - g['mel3']).max() < MEL_TOL  # includes the conv halo frame and padded frames (= output bias)
    lens = g['durations3'].sum(1)
    b = int(np.argmin(lens))
    assert np.abs(mel[b, lens[b]:] - p['out.b'].numpy()).max() < 1e-6
    want_mask = (np.arange(mel.shape[1])[None] >= lens[:, None]).astype(np.float32)
    assert np.array_equal(out['expanded_mask'].cpu().numpy()[:, 0, 0], want_mask)
    assert np.array_equal(out['expanded_mask'].cpu().numpy(), g['expanded_mask3'])


def test_c1_predict_api_integer_durations():
    g = np.load(GOLD / 'c1_forward.npz')
    p = fo.init_params(fo.CONFIGS['C1'], seed=7)
    m = _model('C1', p)
    out = m.predict(g['tokens'][0], encode=False)
    got = out['int_durations'].cpu().numpy()
    want = g['pred_int_durations']
    # integer durations are bit-exact wherever the float duration is not within 1e-3 of a rounding boundary
    ref_float = fo.predict(p, fo.CONFIGS['C1'], torch.from_numpy(g['tokens']))['duration'][..., 0].numpy()
    safe = np.abs(ref_float - np.floor(ref_float) - 0.5) > 1e-3
    assert np.array_equal(got[safe], want[safe])
    if np.array_equal(got, want):
        assert out['mel'].shape == g['pred_mel'].shape
        assert np.abs(out['mel'].cpu().numpy() - g['pred_mel']).max() < MEL_TOL
    for speed in (0.9, 1.2):
        o = m.predict(g['tokens'][0], encode=False, speed_regulator=speed)
        r = fo.predict(p, fo.CONFIGS['C1'], torch.from_numpy(g['tokens']), speed_regulator=speed)
        rf = (r['duration'][..., 0] * np.float32(1.0 / speed)).numpy()
        safe = np.abs(rf - np.floor(rf) - 0.5) > 1e-3
        assert np.array_equal(o['int_durations'].cpu().numpy()[safe], r['int_durations'].numpy()[safe])


@pytest.mark.parametrize('cfg_name', ['LJ256', 'LJ256-dense', 'REF384'])
def test_lj256_ragged_parity(cfg_name):
    """BASELINE configs[1] model (6+6 layers, d=256) on a small ragged batch the oracle finishes in seconds."""
    torch.set_num_threads(8)
    cfg = fo.CONFIGS[cfg_name]
    p = fo.init_params(cfg, seed=7)
    tok, dur, pit = fo.make_inputs('ragged', 3, 40, 300, seed=201)
    ref = fo.forward_transformer_call(p, cfg, tok, dur[..., None], pit[..., None])
    m = _model(cfg_name, p)
    out = m.call(tok, target_durations=dur, target_pitch=pit)
    err = (out['mel'].cpu() - ref['mel']).abs().max().item()
    assert err < MEL_TOL, err
    assert (out['duration'].cpu() - ref['duration']).abs().max() < 1e-3
    assert (out['pitch'].cpu() - ref['pitch']).abs().max() < 1e-3
    # fast mode: report (not gate) the single-pass bf16 error
    mf = _model(cfg_name, p, precision='bf16')
    ef = (mf.call(tok, target_durations=dur, target_pitch=pit)['mel'].cpu() - ref['mel']).abs().max().item()
    print(f'{cfg_name}: mel max-abs-err bf16x3 {err:.2e}, bf16 {ef:.2e}')
    assert ef < 0.25


def test_lj256_full_size_properties():
    """BASELINE configs[1] at full size (B=64, 128 phonemes -> 1000 frames): size-independent properties plus
    oracle parity on a 2-row sub-batch (rows are independent)."""
    torch.set_num_threads(8)
    cfg = fo.CONFIGS['LJ256']
    p = fo.init_params(cfg, seed=7)
    tok, dur, pit = fo.make_inputs('full', 64, 128, 1000, seed=200)
    m = _model('LJ256', p)
    out = m.call(tok, target_durations=dur, target_pitch=pit)
    mel = out['mel']
    assert mel.shape == (64, 1000, 80) and torch.isfinite(mel).all()
    assert int(out['mel_lengths'].min()) == 1000
    # batch-permutation equivariance (rows independent)
    perm = torch.randperm(64, generator=torch.Generator().manual_seed(0))
    out_p = m.call(tok[perm], target_durations=dur[perm], target_pitch=pit[perm])
    # rows are independent; a row that lands in the pair-mode tail of a LayerNorm GEMM (csrc/gemm_tc.cu, hybrid schedule)
    # has its row statistics combined in a different (mathematically equivalent) order, so equality is to rounding noise
    assert (out_p['mel'] - mel[perm.to(mel.device)]).abs().max() < 2e-4
    # oracle parity on two rows
    sel = [3, 41]
    ref = fo.forward_transformer_call(p, cfg, tok[sel], dur[sel][..., None], pit[sel][..., None])
    assert (mel[sel].cpu() - ref['mel']).abs().max() < MEL_TOL


def test_spectrogram_ops_and_losses_mirror_reference_api():
    """utils/spectrogram_ops.py and utils/losses.py of the reference, same names, on the CUDA kernels."""
    from oracle import audio_oracle as ao
    from transformertts_b200.utils import losses, spectrogram_ops
    g = torch.Generator().manual_seed(5)
    mel = torch.randn(4, 60, 80, generator=g)
    mel[0, 40:] = 0
    mel[2, 10:] = 0
    ph = torch.randint(1, 127, (4, 30), generator=g)
    ph[1, 12:] = 0
    assert spectrogram_ops.mel_lengths(mel).cpu().tolist() == ao.mel_lengths(mel.numpy()).tolist()
    assert spectrogram_ops.phoneme_lengths(ph).cpu().tolist() == ao.phoneme_lengths(ph.numpy()).tolist()
    assert torch.equal(spectrogram_ops.mel_padding_mask(mel), 1.0 - (mel == 0).float())
    pred = torch.randn(4, 60, 80, generator=g)
    ref = fo.masked_mean_absolute_error(mel, pred)
    got = losses.masked_mean_absolute_error(mel, pred)
    assert abs(got.item() - ref.item()) < 1e-5
    dur_t = torch.randint(0, 9, (4, 30), generator=g, dtype=torch.int32)
    dur_p = torch.rand(4, 30, 1, generator=g) * 8
    tot, vals = losses.weighted_sum_losses((mel, dur_t[..., None], dur_t[..., None]), (pred, dur_p, dur_p),
                                           [losses.masked_mean_absolute_error] * 3, [1., 1., 3.])
    rt, _ = fo.weighted_sum_losses((mel, dur_t[..., None], dur_t[..., None]), (pred, dur_p, dur_p))
    assert abs(float(tot) - float(rt)) < 1e-4


def test_save_load_roundtrip_and_factory(tmp_path):
    """save_model / load_model / factory.tts_custom keep the reference's two-file layout (config.yaml + weights)."""
    from transformertts_b200.model import factory
    from transformertts_b200.model.models import ForwardTransformer
    p = fo.init_params(fo.CONFIGS['C1'], seed=7)
    m = _model('C1', p)
    tok, dur, pit = fo.make_inputs('ragged', 2, 16, 60, seed=9)
    want = m.call(tok, target_durations=dur, target_pitch=pit)['mel']
    m.save_model(str(tmp_path / 'step_1'))
    assert (tmp_path / 'step_1' / 'config.yaml').exists() and (tmp_path / 'step_1' / 'model_weights.pt').exists()
    m2 = ForwardTransformer.load_model(str(tmp_path / 'step_1'))
    assert torch.equal(m2.call(tok, target_durations=dur, target_pitch=pit)['mel'], want)
    m3, cfg = factory.tts_custom(str(tmp_path / 'step_1' / 'config.yaml'), str(tmp_path / 'step_1'))
    assert cfg['encoder_model_dimension'] == 128
    assert torch.equal(m3.call(tok, target_durations=dur, target_pitch=pit)['mel'], want)
    # the published weights are a directory (zip) of config.yaml + Keras model_weights.hdf5: same layout, read by the
    # pure-python HDF5 reader; without the file on disk (no network here) the factory says where it looked
    pub = tmp_path / 'bdf06b9_ljspeech_step_95000'
    pub.mkdir()
    for f in ('config.yaml', 'model_weights.hdf5'):
        (pub / f).write_bytes((tmp_path / 'step_1' / f).read_bytes())
    m4 = factory.tts_ljspeech(path=str(pub))
    assert torch.equal(m4.call(tok, target_durations=dur, target_pitch=pit)['mel'], want)
    import shutil
    import zipfile
    with zipfile.ZipFile(tmp_path / 'w' / 'bdf06b9_ljspeech_step_90000.zip' if (tmp_path / 'w').mkdir() is None else None, 'w') as z:
        for f in ('config.yaml', 'model_weights.hdf5'):
            z.write(pub / f, f'bdf06b9_ljspeech_step_90000/{f}')
    shutil.rmtree(pub)
    import os
    os.environ['TTSB_WEIGHTS_DIR'] = str(tmp_path / 'w')
    try:
        m5 = factory.tts_ljspeech(step='90000')
        assert torch.equal(m5.call(tok, target_durations=dur, target_pitch=pit)['mel'], want)
        with pytest.raises(FileNotFoundError):
            factory.tts_ljspeech(step='12345')
    finally:
        del os.environ['TTSB_WEIGHTS_DIR']


# ----------------------------------------------------------------------------------------------------------
# edge cases: smallest inputs, rows that expand to nothing, long sequences, duration clamps
# ----------------------------------------------------------------------------------------------------------
def test_edge_single_token_and_single_frame():
    cfg = fo.CONFIGS['C1']
    p = fo.init_params(cfg, seed=7)
    m = _model('C1', p)
    tok = torch.tensor([[5]], dtype=torch.int32)
    dur = torch.tensor([[1]], dtype=torch.int32)
    pit = torch.tensor([[0.3]])
    out = m.call(tok, target_durations=dur, target_pitch=pit)
    ref = fo.forward_transformer_call(p, cfg, tok, dur[..., None].float(), pit[..., None])
    assert out['mel'].shape == (1, 1, 80)
    assert (out['mel'].cpu() - ref['mel']).abs().max() < MEL_TOL
    assert (out['duration'].cpu() - ref['duration']).abs().max() < 1e-3


def test_edge_rows_that_expand_to_zero_frames():
    """One row of the batch has all-zero durations: its decoder rows are all padding (mel = output bias everywhere,
    models.py:541-543); a batch where every row is empty returns a (B, 0, 80) mel."""
    cfg = fo.CONFIGS['C1']
    p = fo.init_params(cfg, seed=7)
    m = _model('C1', p)
    tok, dur, pit = fo.make_inputs('ragged', 3, 20, 60, seed=77)
    dur = dur.clone()
    dur[1] = 0
    out = m.call(tok, target_durations=dur, target_pitch=pit)
    ref = fo.forward_transformer_call(p, cfg, tok, dur[..., None].float(), pit[..., None])
    assert out['mel'].shape == ref['mel'].shape
    assert (out['mel'].cpu() - ref['mel']).abs().max() < MEL_TOL
    assert int(out['mel_lengths'][1]) == 0
    assert (out['mel'][1].cpu() - p['out.b']).abs().max() < 1e-6
    out0 = m.call(tok, target_durations=torch.zeros_like(dur), target_pitch=pit)
    assert out0['mel'].shape == (3, 0, 80)


def test_edge_long_sequences_up_to_the_position_table():
    """encoder_max_position_encoding = 2000 tokens (C1) and a 3000-frame decoder row: many key tiles, ragged tail tiles."""
    torch.set_num_threads(16)
    cfg = fo.CONFIGS['C1']
    p = fo.init_params(cfg, seed=7)
    m = _model('C1', p)
    g = torch.Generator().manual_seed(5)
    Tp = 2000
    tok = torch.randint(1, 127, (1, Tp), generator=g, dtype=torch.int32)
    dur = torch.zeros((1, Tp), dtype=torch.int32)
    dur[0, :1500] = 2                      # 3000 frames; the last 500 tokens expand to nothing
    pit = torch.randn(1, Tp, generator=g)
    out = m.call(tok, target_durations=dur, target_pitch=pit)
    ref = fo.forward_transformer_call(p, cfg, tok, dur[..., None].float(), pit[..., None])
    assert out['mel'].shape == (1, 3000, 80)
    assert (out['mel'].cpu() - ref['mel']).abs().max() < MEL_TOL


def test_edge_predict_speed_regulator_and_duration_clamps():
    """predict(): durations * (1/speed) -> min(max_mask) -> max(min_mask) -> round half to even (models.py:532-539,566)."""
    cfg = fo.CONFIGS['C1']
    p = fo.init_params(cfg, seed=7)
    m = _model('C1', p)
    tok, _, _ = fo.make_inputs('ragged', 2, 24, 100, seed=78)
    for speed in (0.5, 1.0, 1.7):
        out = m.predict(tok, encode=False, speed_regulator=speed)
        ref = fo.predict(p, cfg, tok, speed_regulator=speed)
        fl = ref['duration'][..., 0].numpy() * np.float32(1.0 / speed)
        safe = np.abs(fl - np.floor(fl) - 0.5) > 2e-3          # away from a rounding boundary
        got, want = out['int_durations'].cpu().numpy(), ref['int_durations'].numpy()
        assert np.array_equal(got[safe], want[safe])
        if np.array_equal(got, want):
            assert (out['mel'].cpu() - ref['mel']).abs().max() < MEL_TOL


def test_lj256_against_reference_code_golden():
    """tests/golden/ref_lj256.npz: outputs of the UNMODIFIED reference ForwardTransformer (run on tests/tf_shim by
    tests/golden/make_golden_ref.py) -- the CUDA path against numbers the reference's own code produced."""
    g = np.load(GOLD / 'ref_lj256.npz')
    p = fo.init_params(fo.CONFIGS['LJ256'], seed=7)
    m = _model('LJ256', p)
    out = m.call(torch.from_numpy(g['tokens']), target_durations=torch.from_numpy(g['durations']),
                 target_pitch=torch.from_numpy(g['pitch']))
    assert np.array_equal(out['int_durations'].cpu().numpy(), g['durations'])
    assert np.abs(out['mel'].cpu().numpy() - g['mel']).max() < MEL_TOL
    assert np.abs(out['duration'].cpu().numpy() - g['duration_pred']).max() < 1e-3
    assert np.abs(out['pitch'].cpu().numpy() - g['pitch_pred']).max() < 1e-3


def test_cuda_graph_replay_equals_eager():
    """cuda_graphs=True: the two halves of call() are captured per shape and replayed; results must equal the eager path bit
    for bit, stay valid after later calls (outputs are copied out of the static buffers), follow new input VALUES on replay,
    and a new output length must get its own decoder graph."""
    from transformertts_b200 import lib
    cfg = fo.CONFIGS['C1']
    p = fo.init_params(cfg, seed=7)
    eager = _model('C1', p)
    graphed = _model('C1', p, cuda_graphs=True)
    outs = []
    for seed, Tm in ((11, 180), (12, 180), (13, 210), (11, 180)):
        tok, dur, pit = fo.make_inputs('ragged', 3, 32, Tm, seed=seed)
        a = eager.call(tok, target_durations=dur, target_pitch=pit)
        n0 = lib.launch_count()
        b = graphed.call(tok, target_durations=dur, target_pitch=pit)
        assert lib.launch_count() - n0 > 20          # replays are counted as the launches they contain
        for k in ('mel', 'duration', 'pitch', 'int_durations', 'mel_lengths', 'expanded_mask'):
            assert torch.equal(a[k], b[k]), (seed, k)
        outs.append((a['mel'].clone(), b['mel']))
    for want, got in outs:                            # earlier results were not overwritten by later replays
        assert torch.equal(want, got)
    assert len(graphed._enc_graphs) == 1 and len(next(iter(graphed._enc_graphs.values()))['dec']) == 2
    # predicted durations (no targets) through predict(), graph path == eager path
    tok, _, _ = fo.make_inputs('ragged', 2, 32, 100, seed=14)
    a = eager.predict(tok, encode=False, speed_regulator=0.9)
    b = graphed.predict(tok, encode=False, speed_regulator=0.9)
    assert torch.equal(a['mel'], b['mel']) and torch.equal(a['int_durations'], b['int_durations'])
    with pytest.raises(ValueError):
        graphed.call(tok, target_durations=-torch.ones(2, 32), target_pitch=torch.zeros(2, 32))
