"""Mirror of /root/reference/datasets/voc_fusion2.py -- the dataset `train_seg` reads (train.py:21,131-160): identical to
voc_fusion3.py except that the third image comes from `Mask/` (voc_fusion2.py:27: the fused RGB images `train_fusion` wrote,
train.py:409-411) and is used as the H x W x 3 image it is instead of being replicated from one plane (voc_fusion2.py:47-48 are
commented out).  The device transforms take either kind (csrc/datapath.cu `mask_c`)."""
from . import voc_fusion3
from .voc_fusion3 import load_img_name_list      # noqa: F401


class VOC12Dataset(voc_fusion3.VOC12Dataset):
    MASK_DIR = "Mask"


class VOC12SegDataset(voc_fusion3.VOC12SegDataset):
    MASK_DIR = "Mask"
