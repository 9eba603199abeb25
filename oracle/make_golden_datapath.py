"""Generates tests/golden/datapath.npz by running the UNMODIFIED reference data path (/root/reference/datasets/imutils.py and
voc_fusion3.py `VOC12SegDataset.__transforms`) in this container, with Pillow and OpenCV as installed here.

mmcv (requirements.txt:71) is not installed: `mmcv.bgr2hsv / hsv2bgr` are `cv2.cvtColor(img, cv2.COLOR_BGR2HSV / HSV2BGR)`
(mmcv/image/colorspace.py convert_color_factory), so a two-function stand-in module is registered under that name; imageio
(only used to read PNGs) is stubbed.  Inputs come from oracle.data_oracle.synth_sample(seed, h, w); the global `random` and
`np.random` generators are seeded with the case's seed right before the call, and one extra draw from each is stored after it so
a test also sees that the restatement consumed exactly as many draws as the reference.

    python oracle/make_golden_datapath.py            # needs /root/reference, PIL, cv2
"""
import os
import random
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("SEGMIF_REFERENCE", "/root/reference")

MASK3 = {16, 17, 18, 19}       # cases whose mask is an H x W x 3 image (train_seg's dataset, voc_fusion2.py:44-48)
CASES = [  # (seed, h, w, crop, rescale_range)
    (0, 48, 64, 40, (0.5, 2.0)), (1, 48, 64, 40, (0.5, 2.0)), (2, 48, 64, 40, (0.5, 2.0)), (3, 48, 64, 40, (0.5, 2.0)),
    (4, 60, 80, 64, (0.5, 2.0)), (5, 60, 80, 64, (0.5, 2.0)), (6, 60, 80, 32, (0.5, 2.0)), (7, 60, 80, 32, (0.5, 2.0)),
    (8, 37, 53, 48, (0.5, 2.0)), (9, 37, 53, 48, (0.5, 2.0)), (10, 96, 72, 56, (0.75, 1.25)), (11, 96, 72, 56, (0.75, 1.25)),
    (12, 48, 64, 40, None), (13, 48, 64, 40, None), (14, 48, 64, 40, None), (15, 48, 64, 40, None),
    (16, 48, 64, 40, (0.5, 2.0)), (17, 60, 80, 64, (0.5, 2.0)), (18, 37, 53, 48, (0.5, 2.0)), (19, 48, 64, 40, None),
]


def load_reference():
    import cv2
    mm = types.ModuleType("mmcv")
    mm.bgr2hsv = lambda img: cv2.cvtColor(img, cv2.COLOR_BGR2HSV)
    mm.hsv2bgr = lambda img: cv2.cvtColor(img, cv2.COLOR_HSV2BGR)
    sys.modules.setdefault("mmcv", mm)
    sys.modules.setdefault("imageio", types.ModuleType("imageio"))
    sys.path.insert(0, REF)
    from datasets import imutils, voc_fusion3          # the reference's own modules
    return imutils, voc_fusion3


def main():
    import cv2
    from PIL import Image
    from oracle import data_oracle as do
    imutils, voc = load_reference()
    out = {"cases": np.array([(s, h, w, c, -1 if r is None else r[0], -1 if r is None else r[1], 3 if s in MASK3 else 1) for s, h, w, c, r in CASES], np.float64)}
    for seed, h, w, crop, rr in CASES:
        ir, vis, mask, label = do.synth_sample(seed, h, w, mask_channels=3 if seed in MASK3 else 1)
        image, image_vis, image_mask = do.dataset_views(ir, vis, mask)
        ds = object.__new__(voc.VOC12SegDataset)
        ds.aug, ds.ignore_index, ds.resize_range, ds.rescale_range, ds.crop_size, ds.img_fliplr = True, 255, [512, 640], rr, crop, True
        ds.color_jittor = imutils.PhotoMetricDistortion()
        random.seed(seed)
        np.random.seed(seed)
        a, b, c, d = ds._VOC12SegDataset__transforms(image, image_vis, image_mask, label)
        out[f"c{seed}_image"], out[f"c{seed}_vis"], out[f"c{seed}_mask"], out[f"c{seed}_label"] = (np.ascontiguousarray(x) for x in (a, b, c, d))
        out[f"c{seed}_after"] = np.array([random.random(), np.random.randint(1 << 30)], np.float64)
    # unit vectors of the third-party pieces
    rs = np.random.RandomState(123)
    for i, (h, w, nh, nw) in enumerate([(48, 64, 60, 80), (48, 64, 30, 41), (37, 53, 91, 64), (64, 48, 33, 129), (50, 50, 25, 25), (48, 64, 95, 127)]):
        img = rs.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
        lab = rs.randint(0, 9, size=(h, w)).astype(np.uint8)
        out[f"rs{i}_shape"] = np.array([h, w, nh, nw])
        out[f"rs{i}_img"], out[f"rs{i}_lab"] = img, lab
        out[f"rs{i}_bilinear"] = np.asarray(Image.fromarray(img).resize((nw, nh), resample=Image.BILINEAR))
        out[f"rs{i}_nearest"] = np.asarray(Image.fromarray(lab).resize((nw, nh), resample=Image.NEAREST))
    u8 = rs.randint(0, 256, size=(64, 96, 3)).astype(np.uint8)
    u8[:8] = u8[:8, :, :1]                                  # greys
    hsv8 = cv2.cvtColor(u8, cv2.COLOR_BGR2HSV)
    fl = (rs.rand(64, 96, 3) * 255).astype(np.float32)
    fl[:8] = np.floor(fl[:8])
    fl[8:16] = fl[8:16, :, :1]
    hsvf = cv2.cvtColor(fl, cv2.COLOR_BGR2HSV)
    hsvf_j = hsvf.copy()
    hsvf_j[..., 1] = np.clip(hsvf_j[..., 1] * np.float32(1.3), 0, 255).astype(np.uint8)
    hsvf_j[..., 0] = (hsvf_j[..., 0].astype(int) + 11) % 180
    out.update(hsv_u8_in=u8, hsv_u8_fwd=hsv8, hsv_u8_back=cv2.cvtColor(hsv8, cv2.COLOR_HSV2BGR), hsv_f_in=fl, hsv_f_fwd=hsvf,
               hsv_f_back=cv2.cvtColor(hsvf, cv2.COLOR_HSV2BGR), hsv_f_jit=hsvf_j, hsv_f_jit_back=cv2.cvtColor(hsvf_j, cv2.COLOR_HSV2BGR))
    out["versions"] = np.array([f"pillow {Image.__version__ if hasattr(Image, '__version__') else __import__('PIL').__version__}", f"opencv {cv2.__version__}"])
    dst = os.path.join(ROOT, "tests", "golden", "datapath.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
