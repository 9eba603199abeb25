"""Mirror of /root/reference/datasets/voc_fusion3.py (`VOC12Dataset` :13-62, `VOC12SegDataset` :142-216): same constructor
arguments, same directory layout (`Infrared/ Visible/ Mask2/ Label/` + `<name_list_dir>/<split>.txt`), same return tuple
`(name, image, image_vis, image_mask, label)` -- but the PNGs are decoded once on the host (PNG inflate is sequential work and
stays there), uploaded as uint8, and every transform runs on the device (datasets/imutils.py, csrc/datapath.cu); tensors are
returned on the device.  `batch(indices)` is the entry point that matters at B200 rates: one call per batch."""
import os

import numpy as np
import torch

from . import imutils


def load_img_name_list(img_name_list_path):
    return np.loadtxt(img_name_list_path, dtype=str)                       # voc_fusion3.py:8-10


def _imread(path):
    try:
        import imageio
        return np.asarray(imageio.imread(path))
    except ImportError:                                                    # the reader is plumbing; Pillow decodes the same PNG bytes
        from PIL import Image
        return np.asarray(Image.open(path))


class VOC12Dataset:
    MASK_DIR = "Mask2"             # voc_fusion3.py:27; datasets/voc_fusion2.py overrides it with "Mask" (voc_fusion2.py:27)

    def __init__(self, root_dir=None, name_list_dir=None, split="train", stage="train", device="cuda"):
        self.root_dir, self.stage, self.device = root_dir, stage, torch.device(device)
        self.img_dir = os.path.join(root_dir, "Infrared")
        self.img_dir_vis = os.path.join(root_dir, "Visible")
        self.img_dir_mask = os.path.join(root_dir, self.MASK_DIR)
        self.label_dir = os.path.join(root_dir, "Label")
        self.name_list_dir = os.path.join(name_list_dir, split + ".txt")
        self.name_list = np.atleast_1d(load_img_name_list(self.name_list_dir))

    def __len__(self):
        return len(self.name_list)

    def decode_host(self, idx):
        """voc_fusion3.py:34-60 without the three-fold replication of the single-channel planes: (name, ir, vis, mask, label)
        as uint8 numpy arrays.  Thread-safe (DeviceLoader decodes a batch ahead in a thread pool)."""
        name = str(self.name_list[idx])
        rd = lambda d: np.array(_imread(os.path.join(d, name + ".png")), copy=True)
        return name, rd(self.img_dir), rd(self.img_dir_vis), rd(self.img_dir_mask), rd(self.label_dir)

    def upload(self, decoded):
        """numpy planes of decode_host -> uint8 tensors on the device (on the CALLER's current stream)."""
        return (decoded[0],) + tuple(torch.from_numpy(a).to(self.device, non_blocking=True) if isinstance(a, np.ndarray) else a for a in decoded[1:])

    def decode(self, idx):
        return self.upload(self.decode_host(idx))


class VOC12SegDataset(VOC12Dataset):
    def __init__(self, root_dir=None, name_list_dir=None, split="train", stage="train", resize_range=[512, 640],
                 rescale_range=[0.5, 2.0], crop_size=512, img_fliplr=True, ignore_index=255, aug=False, device="cuda", rng=None, **kwargs):
        super().__init__(root_dir, name_list_dir, split, stage, device)
        self.aug, self.ignore_index = aug, ignore_index
        self.resize_range, self.rescale_range, self.crop_size, self.img_fliplr = resize_range, rescale_range, crop_size, img_fliplr
        self.color_jittor = imutils.PhotoMetricDistortion()
        self.rng = rng if rng is not None else imutils.Rng()
        self.transforms = imutils.DeviceTransforms(crop_size, rescale_range, resize_range, img_fliplr, ignore_index,
                                                   color_jittor=self.color_jittor) if aug else None

    def batch(self, indices, rng=None, label_int64=True):
        """Decoded samples -> one device call.  Returns (names, image, image_vis, image_mask, label[, label_int64])."""
        if not self.aug:
            raise ValueError("segmif_b200.datasets: batch() needs aug=True (a common crop size); use __getitem__ for validation")
        return self.transform_decoded([self.decode(i) for i in indices], rng, label_int64)

    def transform_decoded(self, decoded, rng=None, label_int64=False):
        """`decoded`: list of `decode()` results -> (names, image, image_vis, image_mask, label[, label_int64])."""
        if not self.aug:
            raise ValueError("segmif_b200.datasets: batching needs aug=True (a common crop size); use __getitem__ for validation")
        decoded = [self.upload(d) for d in decoded]
        out = self.transforms([d[1:] for d in decoded], rng if rng is not None else self.rng, label_int64=label_int64)
        return ([d[0] for d in decoded],) + tuple(out)

    def __getitem__(self, idx):
        name, ir, vis, mask, label = self.decode(idx)
        if self.aug:
            a, b, c, d = self.transforms([(ir, vis, mask, label)], self.rng)
            return name, a[0], b[0], c[0], d[0]
        return name, imutils.to_chw_float64(ir), imutils.to_chw_float64(vis), imutils.to_chw_float64(mask), label
