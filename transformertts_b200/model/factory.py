"""API entry points mirroring the reference's ``model/factory.py`` (names and signatures kept)."""
from __future__ import annotations

from pathlib import Path
from typing import Tuple

import torch
import yaml

from .aligner import Aligner
from .models import ForwardTransformer


def _flatten(config: dict) -> dict:
    """The reference's training yaml nests sections; its config manager flattens them into one dict
    (utils/training_config_manager.py:49-56).  Exported model configs (save_model) are flat already."""
    if 'tts_settings' not in config:
        return dict(config)
    flat = {}
    for key in ('paths', 'naming', 'training_data_settings', 'audio_settings', 'text_settings', 'tts_settings'):
        flat.update(config.get(key, {}))
    return flat


def tts_custom(config_path: str, weights_path: str) -> Tuple[ForwardTransformer, dict]:
    """reference: model/factory.py:22-29.  weights_path: a ``model_weights.pt`` file (or the directory holding it)
    with the flat parameter dictionary documented in transformertts_b200/model/models.py."""
    with open(config_path, 'rb') as f:
        config = _flatten(yaml.safe_load(f))
    model = ForwardTransformer.from_config(config)
    model.build_model_weights()
    wp = Path(weights_path)
    if wp.is_dir():
        wp = wp / 'model_weights.pt'
    model.set_weights(torch.load(wp, map_location='cpu'))
    return model, config


def aligner_custom(config_path: str, weights_path: str) -> Tuple[Aligner, dict]:
    """reference: model/factory.py:32-39.  weights_path: ``model_weights.pt`` with the flat parameter dictionary
    documented in transformertts_b200/model/aligner.py (or the directory holding it)."""
    with open(config_path, 'rb') as f:
        raw = yaml.safe_load(f)
    config = {}
    if 'aligner_settings' in raw:
        for key in ('paths', 'naming', 'training_data_settings', 'audio_settings', 'text_settings', 'aligner_settings'):
            config.update(raw.get(key, {}))
    else:
        config = dict(raw)
    model = Aligner.from_config(config, max_r=int(config.get('max_r', 10)))
    model.build_model_weights()
    wp = Path(weights_path)
    if wp.is_dir():
        wp = wp / 'model_weights.pt'
    model.set_weights(torch.load(wp, map_location='cpu'))
    return model, config


def tts_ljspeech(step='95000') -> ForwardTransformer:
    """reference: model/factory.py:10-19 downloads ``bdf06b9_ljspeech_step_{step}.zip`` (Keras HDF5 weights).
    There is no network here and HDF5 import is a later row: load a converted directory instead."""
    raise NotImplementedError('downloading the published Keras weights needs network access and an HDF5 reader; '
                              'convert them offline and use ForwardTransformer.load_model(directory)')
