"""core/model.py of the reference is an older duplicate of WeTr; re-export the one implementation."""
from .model_fusion import WeTr  # noqa: F401
