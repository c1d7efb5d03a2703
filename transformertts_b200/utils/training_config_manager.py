"""Config handling mirroring the reference's ``utils/training_config_manager.py`` for the ForwardTransformer path:
the yaml sections are flattened into one dict (:49-56) that is splatted into the model constructor (:94-100) and the
optimizer is Adam(lr, beta_1 0.9, beta_2 0.98, epsilon 1e-9) (:102-106).  Directory bookkeeping, git-hash checks and
the Aligner branch are outside the hot path."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import yaml

from ..model.models import ForwardTransformer
from ..model.training import Adam

SECTIONS = ('paths', 'naming', 'training_data_settings', 'audio_settings', 'text_settings', 'tts_settings')


class TrainingConfigManager:
    def __init__(self, config_path: str, aligner: bool = False):
        if aligner:
            raise NotImplementedError('the Aligner is a later row of the scope table (SURVEY.md 8f)')
        self.config_path = Path(config_path)
        self.model_kind = 'tts'
        self.config = self._load_config()
        self.learning_rate = float(np.array(self.config['learning_rate_schedule'])[0, 1])
        log_dir = Path(self.config.get('log_directory', '.')) / str(self.config.get('data_name', 'data'))
        self.base_dir = log_dir / f"{self.config.get('tts_settings_name', 'tts')}.{self.config.get('aligner_settings_name', 'aligner')}"
        self.weights_dir = self.base_dir / 'weights'

    def _load_config(self) -> dict:
        with open(self.config_path, 'rb') as f:
            raw = yaml.safe_load(f)
        flat = {}
        for key in SECTIONS:
            flat.update(raw.get(key, {}) or {})
        return flat

    def get_model(self, ignore_hash: bool = True, **overrides) -> ForwardTransformer:
        cfg = dict(self.config)
        cfg.update(overrides)
        return ForwardTransformer.from_config(cfg)

    def compile_model(self, model: ForwardTransformer, beta_1: float = 0.9, beta_2: float = 0.98):
        model._compile(optimizer=Adam(self.learning_rate, beta_1=beta_1, beta_2=beta_2, epsilon=1e-9))
