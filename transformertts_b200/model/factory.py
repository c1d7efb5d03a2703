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
        wp = wp / ('model_weights.pt' if (wp / 'model_weights.pt').exists() else 'model_weights.hdf5')
    if wp.suffix in ('.hdf5', '.h5'):       # a Keras weight file written by the reference's save_weights
        from .hdf5_weights import load_keras_hdf5
        model.set_weights(load_keras_hdf5(model, wp))
    else:
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


def tts_ljspeech(step='95000', path=None) -> ForwardTransformer:
    """reference: model/factory.py:10-19 downloads ``bdf06b9_ljspeech_step_{step}.zip`` (config.yaml + Keras
    ``model_weights.hdf5``) into the Keras cache and calls ``ForwardTransformer.load_model`` on the extracted directory.

    There is no network here, so the archive (or its extracted directory) must already be on disk: ``path`` names it, or it
    is looked up as ``$TTSB_WEIGHTS_DIR/bdf06b9_ljspeech_step_{step}[.zip]`` and in the Keras cache location the reference
    uses (``~/.keras/TransformerTTS_models``).  The HDF5 file is read by the pure-python reader
    (transformertts_b200/utils/hdf5_lite.py); no h5py / TensorFlow is needed."""
    import os
    import zipfile
    name = f'bdf06b9_ljspeech_step_{step}'
    cands = [Path(path)] if path is not None else []
    for root in (os.environ.get('TTSB_WEIGHTS_DIR'), Path.home() / '.keras' / 'TransformerTTS_models'):
        if root:
            cands += [Path(root) / name, Path(root) / (name + '.zip')]
    for c in cands:
        if c.is_dir() and (c / 'config.yaml').exists():
            return ForwardTransformer.load_model(c)
        if c.is_file() and c.suffix == '.zip':
            target = c.with_suffix('')
            with zipfile.ZipFile(c) as z:
                z.extractall(target.parent)
            inner = target if (target / 'config.yaml').exists() else next(p.parent for p in target.parent.rglob('config.yaml'))
            return ForwardTransformer.load_model(inner)
    raise FileNotFoundError(
        f'{name}: not found in {[str(c) for c in cands]}.  The reference fetches it from https://public-asai-dl-models.s3.'
        f'eu-central-1.amazonaws.com/TransformerTTS/api_weights/bdf06b9_ljspeech/{name}.zip; place the archive (or its extracted '
        f'directory) there or pass path=...')
