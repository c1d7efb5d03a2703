"""CPU tests of the host-side logic (no GPU): tile selection, parameter inventory, config handling, tables."""
import numpy as np
import pytest
import torch

from oracle import forward_oracle as fo


def test_block_n_selection():
    from transformertts_b200.model.models import _pick_block_n
    assert _pick_block_n(768, False) == 256
    assert _pick_block_n(1024, False) == 256
    assert _pick_block_n(1536, False) == 256
    assert _pick_block_n(226, True) == 240
    assert _pick_block_n(80, False) == 80
    assert _pick_block_n(128, True) == 128
    assert _pick_block_n(384, True) == 384    # d=384 LayerNorm GEMM: one logical tile, run as a CTA pair of 192 columns each
    assert _pick_block_n(400, True) == 80     # not pairable: several N tiles + separate LayerNorm row kernel
    assert _pick_block_n(1152, False) == 192


def test_positional_encoding_matches_oracle_bitwise():
    from transformertts_b200.model.transformer_utils import positional_encoding
    for n, d in ((2000, 256), (300, 128)):
        assert torch.equal(positional_encoding(n, d), fo.positional_encoding(n, d))


def test_masks_mirror_reference_semantics():
    from transformertts_b200.model import transformer_utils as tu
    seq = torch.tensor([[3, 4, 0, 0], [1, 2, 3, 4]])
    assert torch.equal(tu.create_encoder_padding_mask(seq), fo.create_encoder_padding_mask(seq))
    mel = torch.zeros(2, 5, 3)
    mel[0, :2] = 1
    mel[1, :5] = -1
    assert torch.equal(tu.create_mel_padding_mask(mel), fo.create_mel_padding_mask(mel))
    assert torch.equal(tu.mask_from_lengths(torch.tensor([2, 5]), 5), fo.create_mel_padding_mask(mel))
    assert tu.create_look_ahead_mask(3).tolist() == [[0, 1, 1], [0, 0, 1], [0, 0, 0]]


def test_scheduling_and_losses_mirror_reference():
    from transformertts_b200.utils.scheduling import piecewise_linear_schedule, reduction_schedule
    sched = [[0, 1.0e-4], [100, 5.0e-5], [200, 1.0e-5]]
    assert piecewise_linear_schedule(0, sched) == pytest.approx(1e-4)
    assert piecewise_linear_schedule(50, sched) == pytest.approx(7.5e-5)
    assert piecewise_linear_schedule(1000, sched) == pytest.approx(1e-5)
    assert reduction_schedule(90000, [[0, 10], [80000, 5], [100000, 2]]) == 5


def test_aligner_host_mirror_parameters_and_persistence(tmp_path):
    """Constructor / parameter dictionary / save-load of the Aligner mirror (no kernels involved: device='cpu')."""
    import torch
    from oracle import aligner_oracle as alo
    from transformertts_b200 import lib
    from transformertts_b200.model.aligner import Aligner
    cfg = alo.ALIGNER_CONFIGS['A5']
    m = Aligner.from_config(dict(cfg, device='cpu'), max_r=cfg['max_r'])
    want = alo.init_aligner_params(cfg, seed=7)
    shapes = m._param_shapes()
    assert set(shapes) == set(want)
    assert all(tuple(want[k].shape) == tuple(shapes[k]) for k in want)
    m.set_weights(want)
    m.save_model(tmp_path / 'aligner')
    m2 = Aligner.load_model(tmp_path / 'aligner', device='cpu')
    assert m2.max_r == cfg['max_r'] and m2.r == cfg['max_r']
    assert all(torch.equal(m2.weights[k].cpu(), want[k]) for k in want)
    m2.set_constants(reduction_factor=2, force_decoder_diagonal=True)
    assert m2.r == 2 and m2.force_decoder_diagonal and not m2.force_encoder_diagonal
    import pytest
    with pytest.raises(NotImplementedError):
        m2.predict('text')          # encode=True needs the external phonemizer (no text_pipeline attached)


def test_duration_to_alignment_matrix():
    import numpy as np
    from transformertts_b200.utils.alignments import duration_to_alignment_matrix
    m = duration_to_alignment_matrix([2, 0, 3, 1])
    assert m.shape == (4, 6)
    assert np.array_equal(m, np.array([[1, 1, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0], [0, 0, 1, 1, 1, 0], [0, 0, 0, 0, 0, 1]], dtype=float))
    assert np.array_equal(m.sum(1), [2, 0, 3, 1]) and np.array_equal(m.sum(0), np.ones(6))
