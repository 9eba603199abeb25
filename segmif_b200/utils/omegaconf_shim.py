"""Minimal stand-in for the two OmegaConf calls the reference's entry scripts make (train.py:25,43-ish,
test_fusion.py:30,44, test_segmentation.py:29,43: `cfg = OmegaConf.load(args.config)` followed by attribute access such
as cfg.dataset.crop_size, cfg.optimizer.betas).  segmif_b200.dropin registers it as `omegaconf` ONLY when the real package
is not installed, so that `configs/voc*.yaml` load in an environment without it.  Not a general OmegaConf replacement:
no interpolation, no merge, no structured configs."""
import re

import yaml

_FLOAT = re.compile(r"^[-+]?(\d+\.?\d*|\.\d+)([eE][-+]?\d+)$")      # '1e-6', '6e-5': OmegaConf reads these as floats, PyYAML as str


class DictConfig(dict):
    """dict with attribute access; nested mappings are wrapped recursively, lists keep their type."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(f"Missing key {name}") from e

    def __setattr__(self, name, value):
        self[name] = _wrap(value)


def _wrap(v):
    if isinstance(v, dict):
        return DictConfig({k: _wrap(x) for k, x in v.items()})
    if isinstance(v, (list, tuple)):
        return [_wrap(x) for x in v]
    if isinstance(v, str) and _FLOAT.match(v):
        return float(v)
    return v


class OmegaConf:
    @staticmethod
    def load(path):
        with open(path) as f:
            return _wrap(yaml.safe_load(f) or {})

    @staticmethod
    def create(obj=None):
        if isinstance(obj, str):
            return _wrap(yaml.safe_load(obj) or {})
        return _wrap(obj or {})

    @staticmethod
    def to_container(cfg, resolve=True):
        if isinstance(cfg, dict):
            return {k: OmegaConf.to_container(v) for k, v in cfg.items()}
        if isinstance(cfg, list):
            return [OmegaConf.to_container(v) for v in cfg]
        return cfg
