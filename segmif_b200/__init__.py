"""segmif_b200: the SegMiF hot path as hand-written sm_100a kernels behind a C ABI (include/segmif_b200.h)."""
from .strict import get_precision, precision, set_precision  # noqa: F401
