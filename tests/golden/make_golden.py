"""Generate the AUDIO golden vectors from the CPU oracle (run in the build container):

    python tests/golden/make_golden.py

librosa cannot be imported here, so audio_mel.npz holds outputs of the numpy restatement in oracle/audio_oracle.py
(cross-checked against torch.stft / torchaudio's Slaney filterbank in tests/test_oracle.py) -- this one golden is
"parity unpinned".  The model goldens (c1_forward.npz, ref_lj256.npz, ref_train_c1.npz, aligner_small.npz) are written by
make_golden_ref.py from the reference's own code.
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import audio_oracle as ao  # noqa: E402

OUT = Path(__file__).resolve().parent


def main():
    torch.set_num_threads(4)
    # audio: 2 clips of 0.5 s (one with len % 256 == 0)
    clips = ao.make_clips(2, 11008, seed=400)
    mels = np.stack([ao.mel_spectrogram(c) for c in clips])
    np.savez_compressed(OUT / 'audio_mel.npz', clips_seed=400, n_samples=11008, mel=mels,
                        mel_wavernn=ao.mel_spectrogram(clips[0], normalizer='WaveRNN'))
    print('wrote', [f.name for f in OUT.glob('*.npz')])


if __name__ == '__main__':
    main()
