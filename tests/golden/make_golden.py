"""Generate the committed golden vectors from the CPU oracle (run in the build container):

    python tests/golden/make_golden.py

The reference itself (TensorFlow / librosa) cannot be imported here, so these are outputs of the restatement in
oracle/ with fixed seeds -- see the "PARITY UNPINNED" note in oracle/forward_oracle.py.  Weights are regenerated from
the seed at test time (oracle.forward_oracle.init_params), only inputs' seeds and outputs are stored.
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import audio_oracle as ao  # noqa: E402
from oracle import aligner_oracle as alo  # noqa: E402
from oracle import forward_oracle as fo  # noqa: E402

OUT = Path(__file__).resolve().parent


def main():
    torch.set_num_threads(4)
    cfg = fo.CONFIGS['C1']
    p = fo.init_params(cfg, seed=7)
    # C1: B=1, 32 phonemes -> 250 frames, forced durations/pitch
    tok, dur, pit = fo.make_inputs('full', 1, 32, 250, seed=100)
    out = fo.forward_transformer_call(p, cfg, tok, dur[..., None], pit[..., None])
    # predicted durations through predict()
    pred = fo.predict(p, cfg, tok, speed_regulator=1.0)
    # ragged batch of 3
    tok3, dur3, pit3 = fo.make_inputs('ragged', 3, 40, 200, seed=101)
    out3 = fo.forward_transformer_call(p, cfg, tok3, dur3[..., None], pit3[..., None])
    np.savez_compressed(
        OUT / 'c1_forward.npz',
        tokens=tok.numpy(), durations=dur.numpy(), pitch=pit.numpy(),
        mel=out['mel'].numpy(), duration_pred=out['duration'].numpy(), pitch_pred=out['pitch'].numpy(),
        pred_mel=pred['mel'].numpy(), pred_int_durations=pred['int_durations'].numpy(),
        tokens3=tok3.numpy(), durations3=dur3.numpy(), pitch3=pit3.numpy(), mel3=out3['mel'].numpy(),
        duration_pred3=out3['duration'].numpy(), pitch_pred3=out3['pitch'].numpy())
    # audio: 2 clips of 0.5 s (one with len % 256 == 0)
    clips = ao.make_clips(2, 11008, seed=400)
    mels = np.stack([ao.mel_spectrogram(c) for c in clips])
    np.savez_compressed(OUT / 'audio_mel.npz', clips_seed=400, n_samples=11008, mel=mels,
                        mel_wavernn=ao.mel_spectrogram(clips[0], normalizer='WaveRNN'))
    # Aligner: plumbing-size config, ragged batch of 3, teacher-forced forward + losses
    acfg = alo.ALIGNER_CONFIGS['A-small']
    ap = alo.init_aligner_params(acfg, seed=7)
    tokens, amel, stop = alo.make_aligner_inputs(acfg, 3, 24, 61, seed=503)
    aout = alo.gta_forward(ap, acfg, tokens, amel, stop, r=1, force_decoder_diagonal=True)
    np.savez_compressed(OUT / 'aligner_small.npz', B=3, Tp=24, Tm=61, seed=503, mel=aout['mel'].numpy(),
                        stop_prob=aout['stop_prob'].numpy(),
                        last_attention=aout['decoder_attention']['Decoder_LastBlock_CrossAttention'].numpy(),
                        loss=float(aout['loss']), mel_loss=float(aout['losses']['mel']),
                        stop_loss=float(aout['losses']['stop_prob']), diag_loss=float(aout['losses']['diag_loss']))
    print('wrote', [f.name for f in OUT.glob('*.npz')])


if __name__ == '__main__':
    main()
