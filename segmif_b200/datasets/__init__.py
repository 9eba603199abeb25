"""Device-side training data path: mirrors of the reference's datasets/imutils.py and
datasets/voc_fusion3.py whose pixel work runs in csrc/datapath.cu."""
from . import imutils, voc_fusion2, voc_fusion3                            # noqa: F401
from .imutils import DeviceTransforms, PhotoMetricDistortion, Rng, to_chw_float64          # noqa: F401
from .loader import DeviceLoader                                            # noqa: F401
from .voc_fusion3 import VOC12Dataset, VOC12SegDataset                      # noqa: F401
