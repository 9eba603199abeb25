"""Generates tests/golden/validation.npz from the UNMODIFIED reference: util/util.py::compute_results is imported from
/root/reference and run on a seeded confusion matrix (built with sklearn exactly as test_segmentation.py:173-176 does),
and the uint8 post-processing lines of val_performance.py:447-460 are executed verbatim on a seeded image batch.
Run in the build container only:  python -m oracle.make_golden_validation"""
import importlib.util
import os
import sys

import numpy as np
import torch
from sklearn.metrics import confusion_matrix

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import segmif_oracle as O  # noqa: E402


def main():
    spec = importlib.util.spec_from_file_location("ref_util", "/root/reference/util/util.py")
    ref_util = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_util)
    rs = np.random.RandomState(7)
    label = rs.randint(0, 9, size=(2, 48, 64)).astype(np.int64)
    label[rs.uniform(size=label.shape) < 0.07] = 255
    label[label == 6] = 5                                            # class 6 never occurs: NaN rows
    pred = np.where(rs.uniform(size=label.shape) < 0.7, np.clip(label, 0, 8), rs.randint(0, 9, size=label.shape)).astype(np.int64)
    pred[pred == 6] = 4
    conf = confusion_matrix(y_true=label.flatten(), y_pred=pred.flatten(), labels=[0, 1, 2, 3, 4, 5, 6, 7, 8])   # test_segmentation.py:176
    prec, rec, iou = ref_util.compute_results(conf)                                                              # util/util.py:31-55
    fusion_image = torch.from_numpy(rs.uniform(-0.15, 0.9, size=(2, 3, 20, 28)).astype(np.float32))
    # ---- val_performance.py:447-460, verbatim (minus .cuda())
    ones = torch.ones_like(fusion_image)
    zeros = torch.zeros_like(fusion_image)
    fi = torch.where(fusion_image > ones, ones, fusion_image)
    fi = torch.where(fi < zeros, zeros, fi)
    fused_image = fi.cpu().numpy()
    fused_image = np.uint8(255.0 * fused_image)
    fused_image = fused_image.transpose((0, 2, 3, 1))
    fused_image = (fused_image - np.min(fused_image)) / (np.max(fused_image) - np.min(fused_image))
    fused_image = np.uint8(255.0 * fused_image)
    # ----
    out = os.path.join(ROOT, "tests", "golden", "validation.npz")
    np.savez_compressed(out, label=label, pred=pred, conf=conf, precision=prec, recall=rec, iou=iou,
                        fusion_image=fusion_image.numpy(), fused_uint8=fused_image)
    oc = O.confusion_matrix(torch.from_numpy(label), torch.from_numpy(pred)).numpy()
    op, orr, oi = O.compute_results(oc)
    print("oracle vs reference: conf equal", bool((oc == conf).all()), "| metrics equal",
          all(np.allclose(a, b, equal_nan=True, rtol=0, atol=0) for a, b in ((op, prec), (orr, rec), (oi, iou))),
          "| uint8 equal", bool((O.fused_to_uint8(fusion_image) == fused_image).all()))
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
