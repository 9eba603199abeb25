"""Makes the reference's own entry scripts (train.py, test_fusion.py, test_segmentation.py) resolve their imports
-- `core`, `core.model_fusion`, `pytorch_ssim`, `lap_loss`, `utils.optimizer` -- to the segmif_b200 mirror.

    python -m segmif_b200.dropin path/to/test_fusion.py [script args...]
or, from Python, before importing the script:
    import segmif_b200.dropin as d; d.install()

Nothing in the scripts has to change; dataset / checkpoint paths inside them are of course still
the reference's own.  The scripts read configs/voc*.yaml through `OmegaConf.load`; when omegaconf is not installed a
minimal stand-in (segmif_b200/utils/omegaconf_shim.py: load + attribute access) is registered under that name."""
import importlib
import runpy
import sys

_ALIASES = {
    "core": "segmif_b200.core",
    "core.mix_transformer": "segmif_b200.core.mix_transformer",
    "core.segformer_head": "segmif_b200.core.segformer_head",
    "core.model_fusion": "segmif_b200.core.model_fusion",
    "core.model": "segmif_b200.core.model",
    "core.loss": "segmif_b200.core.loss",
    "core.Entropy": "segmif_b200.core.Entropy",
    "pytorch_ssim": "segmif_b200.pytorch_ssim",
    "lap_loss": "segmif_b200.lap_loss",
    "utils.optimizer": "segmif_b200.utils.optimizer",
}


def install(force=False):
    """Registers the aliases in sys.modules.  Existing unrelated modules of the same name are kept unless force."""
    for alias, target in _ALIASES.items():
        if alias in sys.modules and not force:
            continue
        mod = importlib.import_module(target)
        sys.modules[alias] = mod
        if "." in alias:                                   # make `utils.optimizer` importable even if `utils` is not ours
            parent, child = alias.rsplit(".", 1)
            if parent not in sys.modules:
                pmod = importlib.import_module(_ALIASES.get(parent, target.rsplit(".", 1)[0]))
                sys.modules[parent] = pmod
            if not hasattr(sys.modules[parent], child):
                setattr(sys.modules[parent], child, mod)
    if "omegaconf" not in sys.modules:
        try:
            importlib.import_module("omegaconf")
        except ImportError:
            sys.modules["omegaconf"] = importlib.import_module("segmif_b200.utils.omegaconf_shim")
    return sorted(_ALIASES)


if __name__ == "__main__":
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    import os
    install()
    script = sys.argv[1]
    sys.argv = sys.argv[1:]
    sys.path.insert(0, os.path.dirname(os.path.abspath(script)))     # `python script.py` semantics: sibling modules importable
    runpy.run_path(script, run_name="__main__")
