"""Mirror of the reference's `core` package surface (core/__init__.py:1-5 of SegMiF)."""
from .segformer_head import SegFormerHead  # noqa: F401
from .mix_transformer import *  # noqa: F401,F403
from .model import WeTr  # noqa: F401
from .model_fusion import Network  # noqa: F401  (alias of Network3; the reference's import of it is broken)
from .loss import *  # noqa: F401,F403
