"""GPU parity of the Aligner teacher-forced forward + validation losses through the
reference-facing API (Aligner.call / _val_step), against oracle/aligner_oracle.py and the committed golden vectors."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import aligner_oracle as alo

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / 'golden'
TOL = 1e-3   # same bar as the text->mel path: fp32 outputs within 1e-3 abs
ATT_TOL = 2e-3


def _model(cfg_name, params, **kw):
    from transformertts_b200.model.aligner import Aligner
    cfg = alo.ALIGNER_CONFIGS[cfg_name]
    m = Aligner.from_config(dict(cfg, **kw), max_r=cfg['max_r'])
    m.set_weights(params)
    return m


def _cmp_outputs(out, ref, lens_q=None):
    assert (out['mel'].cpu() - ref['mel']).abs().max() < TOL
    assert (out['stop_prob'].cpu() - ref['stop_prob']).abs().max() < TOL
    assert (out['linear'].cpu() - ref['linear']).abs().max() < TOL
    assert set(out['decoder_attention']) == set(ref['decoder_attention'])
    assert set(out['encoder_attention']) == set(ref['encoder_attention'])
    for k, w in ref['decoder_attention'].items():
        got = out['decoder_attention'][k].cpu()
        assert got.shape == w.shape
        assert (got - w).abs().max() < ATT_TOL, k
    for k, w in ref['encoder_attention'].items():
        assert (out['encoder_attention'][k].cpu() - w).abs().max() < ATT_TOL, k
    assert torch.equal(out['mel_mask'].cpu(), ref['mel_mask'])
    assert torch.equal(out['text_mask'].cpu(), ref['text_mask'])


@pytest.mark.parametrize('impl', ['tcgen05', 'simt'])
def test_aligner_small_golden_and_losses(impl):
    g = np.load(GOLD / 'aligner_small.npz')
    cfg = alo.ALIGNER_CONFIGS['A-small']
    p = alo.init_aligner_params(cfg, seed=7)
    tokens, mel, stop = alo.make_aligner_inputs(cfg, int(g['B']), int(g['Tp']), int(g['Tm']), seed=int(g['seed']))
    m = _model('A-small', p, impl=impl)
    m.set_constants(reduction_factor=1, force_decoder_diagonal=True)
    out = m._val_step(tokens, mel, stop)
    assert np.abs(out['mel'].cpu().numpy() - g['mel']).max() < TOL
    assert np.abs(out['stop_prob'].cpu().numpy() - g['stop_prob']).max() < TOL
    assert np.abs(out['decoder_attention']['Decoder_LastBlock_CrossAttention'].cpu().numpy() - g['last_attention']).max() < ATT_TOL
    assert abs(float(out['losses']['mel']) - float(g['mel_loss'])) < 1e-3
    assert abs(float(out['losses']['stop_prob']) - float(g['stop_loss'])) < 1e-3
    assert abs(float(out['losses']['diag_loss']) - float(g['diag_loss'])) < 1e-3
    assert abs(float(out['loss']) - float(g['loss'])) < 2e-3


@pytest.mark.parametrize('r', [1, 2])
def test_aligner_small_all_outputs_vs_oracle(r):
    torch.set_num_threads(8)
    cfg = alo.ALIGNER_CONFIGS['A-small']
    p = alo.init_aligner_params(cfg, seed=7)
    tokens, mel, _ = alo.make_aligner_inputs(cfg, 3, 30, 150, seed=504)
    tgt = mel[:, 0::r].contiguous()
    m = _model('A-small', p)
    m.set_constants(reduction_factor=r)
    out = m.call(tokens, tgt, training=False)
    ref = alo.aligner_call(p, cfg, tokens, tgt, r=r)
    assert out['mel'].shape == (3, tgt.shape[1] * r, 80)
    _cmp_outputs(out, ref)


def test_aligner_shipped_config_ragged_batch():
    """aligner_settings as shipped (config/training_config.yaml:58-69): 4 encoder blocks of 4 heads, decoder heads
    [4,4,4,4,1] -> the last block attends with ONE head of dimension 256."""
    torch.set_num_threads(16)
    cfg = alo.ALIGNER_CONFIGS['A5']
    p = alo.init_aligner_params(cfg, seed=7)
    tokens, mel, stop = alo.make_aligner_inputs(cfg, 3, 70, 333, seed=500)
    m = _model('A5', p)
    m.set_constants(reduction_factor=1, force_decoder_diagonal=True, force_encoder_diagonal=True)
    out = m._val_step(tokens, mel, stop)
    ref = alo.gta_forward(p, cfg, tokens, mel, stop, r=1, stop_scaling=8.0, force_decoder_diagonal=True, force_encoder_diagonal=True)
    _cmp_outputs(out, ref)
    for k in ('mel', 'stop_prob', 'diag_loss'):
        assert abs(float(out['losses'][k]) - float(ref['losses'][k])) < 1e-3, k
    assert out['decoder_attention']['Decoder_LastBlock_CrossAttention'].shape == (3, 1, 332, 70)


def test_scaled_ce_kernel_matches_reference_known_answers():
    """reference tests/test_loss.py:12-24 -- the same known answers, through the CUDA kernel."""
    from transformertts_b200 import lib
    dev = torch.device('cuda:0')
    targets = torch.tensor([[0, 1, 2]], dtype=torch.int32, device=dev)
    logits = torch.tensor([[[.3, .2, .1], [.3, .2, .1], [.3, .2, .1]]], dtype=torch.float32, device=dev)
    for scaling, want in ((5.0, 2.3705523014068604), (1.0, 0.7679619193077087)):
        out = torch.zeros(1, device=dev)
        lib.scaled_ce_loss(logits, 3, 3, targets, 2, scaling, out)
        assert abs(float(out) - want) < 1e-6


def test_diag_loss_kernel_vs_oracle():
    from transformertts_b200 import lib
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(3)
    att = torch.softmax(torch.randn(3, 2, 50, 17, generator=g), -1)
    q_len = torch.tensor([50, 31, 7], dtype=torch.int32)
    k_len = torch.tensor([17, 9, 3], dtype=torch.int32)
    want = (att * alo.batch_diagonal_mask(att, q_len, k_len)).sum((-2, -1)).mean() / 10.0
    out = torch.zeros(1, device=dev)
    lib.diag_loss(att.to(dev).contiguous(), q_len.to(dev), k_len.to(dev), out)
    assert abs(float(out) - float(want)) < 1e-5


def test_aligner_unbuilt_paths_fail_loudly():
    from transformertts_b200 import lib
    cfg = alo.ALIGNER_CONFIGS['A-small']
    m = _model('A-small', alo.init_aligner_params(cfg, seed=7))
    tokens, mel, stop = alo.make_aligner_inputs(cfg, 2, 12, 20, seed=1)
    with pytest.raises(NotImplementedError):
        m.predict('some text')               # encode=True needs the external phonemizer (no text_pipeline attached)
    with pytest.raises(lib.TtsbError):
        m.call(tokens, mel, training=True)   # training goes through _train_step (dropout + backward)


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.mark.parametrize('cfg_name,B,Tp,Tm,r,diag', [('A-small', 3, 20, 97, 1, True), ('A-small', 2, 16, 81, 2, False),
                                                     ('A5', 2, 40, 161, 1, True)])
def test_aligner_train_step_loss_grads_and_adam(cfg_name, B, Tp, Tm, r, diag):
    """One deterministic Aligner training step (dropout off): losses, every parameter gradient and the Adam update vs
    torch autograd on the restated fp32 graph.  The GPU forward/backward is single-pass bf16 (as the ForwardTransformer's
    training step), hence the same tolerances as tests/test_gpu_train.py."""
    import math
    torch.set_num_threads(8)
    from oracle import forward_oracle as fo
    from transformertts_b200.model.training import Adam
    cfg = alo.ALIGNER_CONFIGS[cfg_name]
    p = alo.init_aligner_params(cfg, seed=7)
    tokens, mel, stop = alo.make_aligner_inputs(cfg, B, Tp, Tm, seed=510)
    ref_out, ref_g = alo.loss_and_grads(p, cfg, tokens, mel, stop, r=r, stop_scaling=8.0, force_decoder_diagonal=diag,
                                        force_encoder_diagonal=diag)
    m = _model(cfg_name, p, train_dropout=False)
    m._compile(8.0, Adam(1e-4))
    m.set_constants(reduction_factor=r, force_decoder_diagonal=diag, force_encoder_diagonal=diag)
    eng = m._get_engine()
    out = eng.forward_backward(tokens, mel, stop, training=True)
    torch.cuda.synchronize()
    for k in ('mel', 'stop_prob', 'diag_loss'):
        assert abs(out['losses'][k].item() - float(ref_out['losses'][k])) < 2e-2 * abs(float(ref_out['losses'][k])) + 1e-4, k
    assert abs(out['loss'].item() - float(ref_out['loss'])) < 2e-2 * abs(float(ref_out['loss']))
    worst = []
    gscale = max(float(g.norm()) for g in ref_g.values())
    unused_fp = cfg['mel_channels'] * r
    for name, gref in ref_g.items():
        got = eng.g[name].detach().double().cpu()
        if name.startswith('final_proj'):
            # only the first r*mel output columns take part (models.py:146): the rest has exactly zero gradient
            assert float(got[..., unused_fp:].abs().max() if got[..., unused_fp:].numel() else 0.0) == 0.0
        if float(gref.norm()) < 1e-6 * gscale:
            assert float(got.norm()) < 2e-3 * gscale, name   # analytically zero (key biases)
            continue
        if gref.dim() == 0:
            assert abs(float(got) - float(gref)) < 0.5 * abs(float(gref)) + 5e-3 * gscale, name
            continue
        cos = float((got * gref.double()).sum() / (got.norm() * gref.double().norm()))
        worst.append((_rel(got, gref), cos, name))
    worst.sort(reverse=True)
    print('worst gradient relative errors:', worst[:8])
    assert worst[0][0] < 0.3 and min(w[1] for w in worst) > 0.95, worst[:8]
    w0, g0 = eng.flat_w.clone(), eng.flat_g.clone()
    eng.apply_adam(m.optimizer)
    torch.cuda.synchronize()
    m_ref, v_ref, w_ref = torch.zeros_like(w0).cpu(), torch.zeros_like(w0).cpu(), w0.cpu().clone()
    fo.adam_tf_step(w_ref, g0.cpu(), m_ref, v_ref, 1, 1e-4)
    assert (eng.flat_w.cpu() - w_ref).abs().max() < 3e-7
    assert m.step == 1
    out2 = m.train_step(tokens, mel, stop)
    assert math.isfinite(out2['loss'].item())
    # the high-precision validation path still works after training (weights re-packed from the flat buffer)
    out3 = m._val_step(tokens, mel, stop)
    assert math.isfinite(float(out3['loss']))


def test_aligner_training_with_dropout_runs_and_differs():
    cfg = alo.ALIGNER_CONFIGS['A-small']
    p = alo.init_aligner_params(cfg, seed=7)
    tokens, mel, stop = alo.make_aligner_inputs(cfg, 3, 18, 64, seed=511)
    m = _model('A-small', p, train_dropout=True)
    m.set_constants(reduction_factor=1, force_decoder_diagonal=True)
    eng = m._get_engine()
    a = eng.forward_backward(tokens, mel, stop, training=True)['loss'].item()
    b = eng.forward_backward(tokens, mel, stop, training=True)['loss'].item()
    e = eng.forward_backward(tokens, mel, stop, training=False)['loss'].item()
    assert abs(a - b) < 1e-5 * abs(a)          # same step counter -> same dropout masks
    assert abs(a - e) > 1e-4                    # dropout really was active
    assert torch.isfinite(eng.flat_g).all()


def test_durations_from_aligner_attention_match_reference_dijkstra():
    """Aligner forward -> last-block cross-attention -> durations, GPU kernels vs the reference's own procedure
    (oracle/alignment_oracle.py: numpy scores + scipy.sparse.csgraph.dijkstra exactly as utils/alignments.py calls it)."""
    from oracle import alignment_oracle as ao
    from transformertts_b200.utils.alignments import get_durations_from_alignment
    cfg = alo.ALIGNER_CONFIGS['A-small']
    p = alo.init_aligner_params(cfg, seed=7)
    tokens, mel, stop = alo.make_aligner_inputs(cfg, 4, 28, 120, seed=520)
    m = _model('A-small', p)
    m.set_constants(reduction_factor=1)
    out = m.call(tokens, mel[:, :-1].contiguous(), training=False)   # teacher-forced input as in extract_durations.py
    # the last block of A-small has one head: also exercise a 4-head map with a clear diagonal
    g = torch.Generator().manual_seed(9)
    B, H, Tq, Tk = 4, 4, mel.shape[1] - 1, tokens.shape[1]
    lens_q = (mel.abs().sum(-1) != 0).sum(-1) - 1
    lens_k = (tokens != 0).sum(-1) - 1
    logits = torch.randn(B, H, Tq, Tk, generator=g)
    for b in range(B):
        qi = torch.arange(Tq)[:, None] / max(int(lens_q[b]), 1)
        ki = torch.arange(Tk)[None, :] / max(int(lens_k[b]), 1)
        for h in range(H):
            logits[b, h] -= (2.0 + 3.0 * h) * Tk * 0.2 * (qi - ki).abs()   # head 3 is the sharpest diagonal
        logits[b, :, :, int(lens_k[b]) + 1:] = -1e9
    synth = torch.softmax(logits, -1)
    for att in (out['decoder_attention']['Decoder_LastBlock_CrossAttention'].cpu(), synth):
        for weighted in (False, True):
            d_ref, jump_r, peak_r, diag_r = ao.get_durations_from_alignment(att.numpy(), mel.numpy(), tokens.numpy(), weighted=weighted)
            d_gpu, _, jump, peak, diag = get_durations_from_alignment(att, mel, tokens, weighted=weighted)
            assert np.allclose(jump.cpu().numpy(), jump_r, atol=1e-6)
            assert np.allclose(peak.cpu().numpy(), peak_r, rtol=1e-5, atol=1e-7)
            assert np.allclose(diag.cpu().numpy(), diag_r, rtol=1e-4)
            for a, b_ in zip(d_gpu, d_ref):
                assert a.dtype == np.int32 and np.array_equal(a, b_)      # integer durations: bit-exact


def test_aligner_cuda_graph_replay_equals_eager():
    """cuda_graphs / train_graphs: the teacher-forced validation step and the training step replayed as CUDA graphs must give
    the eager results (same kernels, same arguments; weight-gradient sums use fp32 atomics -> training equal to rounding noise)."""
    from transformertts_b200.model.aligner import Aligner
    from transformertts_b200.model.training import Adam
    cfg = alo.ALIGNER_CONFIGS['A-small']
    p = alo.init_aligner_params(cfg, seed=7)
    tok, mel, stop = alo.make_aligner_inputs(cfg, 3, 20, 49, seed=505)
    eager = Aligner.from_config(dict(cfg), max_r=cfg['max_r'])
    graphed = Aligner.from_config(dict(cfg, cuda_graphs=True), max_r=cfg['max_r'])
    for m in (eager, graphed):
        m.set_weights(p)
        m.set_constants(reduction_factor=1, force_decoder_diagonal=True)
    for _ in range(2):
        a, b = eager._val_step(tok, mel, stop), graphed._val_step(tok, mel, stop)
        assert torch.equal(a['mel'], b['mel']) and torch.equal(a['stop_prob'], b['stop_prob'])
        # the loss kernels fold block partials with fp32 atomics: equal up to the order of those additions
        assert abs(float(a['loss']) - float(b['loss'])) <= 2e-6 * abs(float(a['loss']))
        for k in a['decoder_attention']:
            assert torch.equal(a['decoder_attention'][k], b['decoder_attention'][k])
    losses = []
    for graphs in (False, True):
        m = Aligner.from_config(dict(cfg, train_dropout=False, train_graphs=graphs), max_r=cfg['max_r'])
        m.set_weights(p)
        m._compile(cfg['stop_loss_scaling'], Adam(1e-4))
        m.set_constants(reduction_factor=1, force_decoder_diagonal=True)
        losses.append([float(m.train_step(tok, mel, stop)['loss']) for _ in range(4)])
    assert all(abs(x - y) < 5e-4 * abs(x) for x, y in zip(*losses)), losses
    assert losses[0][3] < losses[0][0]


@pytest.mark.parametrize('r,stop_bias', [(1, (6.0, 0.0, -6.0)), (2, (6.0, 0.0, -6.0)), (2, (-6.0, 0.0, 6.0))])
def test_aligner_autoregressive_predict(r, stop_bias):
    """Aligner.predict (models.py:271-292) on the GPU against the oracle loop (itself pinned to the reference's predict in
    tests/test_reference_shim.py).  The stop head is biased so the stop decision does not hang on a near-tie: either the loop
    runs to max_length (the prefix is fed back 13 / 7 times) or it stops after the first iteration."""
    cfg = alo.ALIGNER_CONFIGS['A-small']
    p = dict(alo.init_aligner_params(cfg, seed=7))
    p['postnet.stop.b'] = torch.tensor(stop_bias)
    tok, _, _ = alo.make_aligner_inputs(cfg, 2, 12, 21, seed=3)
    m = _model('A-small', p)
    m.set_constants(reduction_factor=r)
    out = m.predict(tok[0], max_length=12, encode=False, verbose=False)
    ref = alo.aligner_predict(p, dict(cfg, dropout_rate=0.0, decoder_prenet_dropout=0.0), tok[0], float(m.start_vec[0, 0]), max_length=12, r=r,
                              stop_prob_index=m.stop_prob_index)
    a, b = out['mel'].float().cpu(), ref['mel']
    assert a.shape == b.shape, (a.shape, b.shape)
    assert a.shape[0] == ((12 // r + 1) * r if stop_bias[0] > 0 else r)
    assert float((a - b).abs().max()) < 5e-3 * max(1.0, float(b.abs().max())), float((a - b).abs().max())
    with pytest.raises(NotImplementedError):
        m.predict('text', max_length=4)      # encode=True needs the external phonemizer
