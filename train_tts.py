#!/usr/bin/env python
"""ForwardTransformer training driver with the loop contract of the reference's ``train_tts.py`` (:149-209):

    batch -> learning_rate = piecewise_linear_schedule(step) -> model.set_constants -> model.train_step -> loss / checkpoints

    python train_tts.py --config config/training_config.yaml --synthetic [--max_steps N] [--batch_size B]
    torchrun --nproc-per-node 8 train_tts.py --config ... --synthetic          # data parallel, one process per GPU

The dataset readers / bucketing (`data/datasets.py`) are a later row of the scope table, so batches
come from the seeded synthetic generator (`--synthetic`, LJSpeech-shaped) unless a directory of pre-batched ``.npz``
files (keys: phonemes, mel, durations, pitch) is given with ``--batches``.
"""
from __future__ import annotations

import argparse
import glob
import os
from pathlib import Path

import numpy as np
import torch

from transformertts_b200.utils.data_parallel import init_from_env
from transformertts_b200.utils.scheduling import piecewise_linear_schedule
from transformertts_b200.utils.training_config_manager import TrainingConfigManager


def synthetic_batches(B, Tp, Tm, mel_channels, seed):
    g = torch.Generator().manual_seed(seed)
    while True:
        tok = torch.randint(1, 127, (B, Tp), generator=g, dtype=torch.int32)
        extra = torch.multinomial(torch.ones(B, Tp), Tm - Tp, replacement=True, generator=g)
        dur = torch.ones(B, Tp, dtype=torch.int32)
        dur.scatter_add_(1, extra, torch.ones_like(extra, dtype=torch.int32))
        pitch = torch.randn(B, Tp, generator=g)
        mel = (torch.randn(B, Tm, mel_channels, generator=g) * 2 - 5).clamp(-11.5, 2.0)
        yield mel, tok, dur, pitch


def file_batches(pattern):
    files = sorted(glob.glob(pattern))
    if not files:
        raise FileNotFoundError(pattern)
    while True:
        for f in files:
            z = np.load(f)
            yield torch.from_numpy(z['mel']), torch.from_numpy(z['phonemes']), torch.from_numpy(z['durations']), torch.from_numpy(z['pitch'])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--config', required=True)
    ap.add_argument('--reset_dir', action='store_true')
    ap.add_argument('--reset_logs', action='store_true')
    ap.add_argument('--reset_weights', action='store_true')
    ap.add_argument('--synthetic', action='store_true')
    ap.add_argument('--batches', default=None, help='glob of pre-batched .npz files')
    ap.add_argument('--max_steps', type=int, default=None)
    ap.add_argument('--batch_size', type=int, default=32)
    ap.add_argument('--weights_dir', default=None)
    args = ap.parse_args()

    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    rank, world = init_from_env(device=torch.device('cuda', local_rank))
    np.random.seed(42)
    torch.manual_seed(42)

    cm = TrainingConfigManager(args.config)
    cfg = cm.config
    model = cm.get_model(device=f'cuda:{local_rank}')
    cm.compile_model(model)
    weights_dir = Path(args.weights_dir) if args.weights_dir else cm.weights_dir
    if args.batches:
        data = file_batches(args.batches)
    elif args.synthetic:
        data = synthetic_batches(args.batch_size, 128, 1000, int(cfg.get('mel_channels', 80)), seed=1000 + rank)
    else:
        raise SystemExit('dataset readers are outside the hot path: pass --synthetic or --batches "<glob of .npz>"')
    max_steps = args.max_steps or int(cfg['max_steps'])
    losses = []
    while model.step < max_steps:
        mel, phonemes, durations, pitch = next(data)
        lr = piecewise_linear_schedule(model.step, cfg['learning_rate_schedule'])
        model.set_constants(learning_rate=lr)
        out = model.train_step(input_sequence=phonemes, target_sequence=mel, target_durations=durations, target_pitch=pitch,
                               data_parallel=world > 1)
        losses.append(float(out['loss']))
        if rank == 0 and (model.step % 10 == 0 or model.step == 1):
            print(f'step {model.step}  loss {losses[-1]:.4f}  mel {float(out["losses"]["mel"]):.4f}  '
                  f'duration {float(out["losses"]["duration"]):.4f}  pitch {float(out["losses"]["pitch"]):.4f}  lr {lr:.2e}', flush=True)
        if rank == 0 and model.step % int(cfg.get('weights_save_frequency', 5000)) == 0 and \
                model.step >= int(cfg.get('weights_save_starting_step', 0)):
            model.save_model(weights_dir / f'step_{model.step}')
    if rank == 0:
        model.save_model(weights_dir / f'step_{model.step}')
        print('Done.')


if __name__ == '__main__':
    main()
