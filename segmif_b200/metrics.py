"""Validation on the device: the confusion matrix / per-class precision, recall and IoU of
test_segmentation.py:169-180 + util/util.py:31-55, and the fused-image post-processing of val_performance.py:447-460,
without the per-image device->host copies, numpy and sklearn of the reference.  Only the final nc x nc matrix (or the
uint8 images to be written to disk) ever leaves the GPU."""
import ctypes

import torch

from . import _lib
from .ops import _prep, _ptr


def confusion_matrix(labels, prediction, num_classes=9, out=None):
    """int64 [nc, nc] on the device; rows = ground truth, columns = prediction; accumulates into `out` when given
    (`conf_total += conf`).  Values outside [0, nc) -- the ignore label 255 -- are dropped, as sklearn's
    confusion_matrix(..., labels=range(nc)) does."""
    t = labels.reshape(-1).long().contiguous()
    p = prediction.reshape(-1).long().contiguous()
    if t.numel() != p.numel():
        raise ValueError("confusion_matrix: labels and prediction differ in size")
    if out is None:
        out = torch.zeros((num_classes, num_classes), dtype=torch.int64, device=t.device)
    st = _prep(t, p, out)
    _lib.call("segmif_confusion_matrix", _ptr(t), _ptr(p), t.numel(), num_classes, _ptr(out), st)
    return out


def compute_results(conf_total):
    """util/util.py:31-55 of the reference (consider_unlabeled = True): per-class precision, recall and IoU from the
    accumulated confusion matrix; NaN where a class never occurs.  Runs on the 81 numbers of the matrix on the host."""
    conf = conf_total.detach().to("cpu", torch.float64)
    tp = conf.diag()
    pred_total, true_total = conf.sum(0), conf.sum(1)
    nan = torch.full_like(tp, float("nan"))
    precision = torch.where(pred_total == 0, nan, tp / pred_total)
    recall = torch.where(true_total == 0, nan, tp / true_total)
    union = true_total + pred_total - tp
    iou = torch.where(union == 0, nan, tp / union)
    return precision.numpy(), recall.numpy(), iou.numpy()


class SegmentationMetrics:
    """Accumulates test_segmentation.py's `conf_total` on the device: update(model, images, labels) runs Network3 ->
    fused bilinear upsample + argmax -> confusion kernel; results() = compute_results(conf_total)."""

    def __init__(self, num_classes=9, device="cuda"):
        self.num_classes = num_classes
        self.conf_total = torch.zeros((num_classes, num_classes), dtype=torch.int64, device=device)

    @torch.no_grad()
    def update(self, model, images, labels):
        pred = model.predict_labels(images, size=labels.shape[-2:])
        confusion_matrix(labels.to(pred.device), pred, self.num_classes, out=self.conf_total)
        return pred

    def results(self):
        return compute_results(self.conf_total)


def fused_to_uint8(rgb):
    """val_performance.py:447-460 for a batch: fp32 NCHW RGB (e.g. ops.recompose_rgb(fused_y, vis, clamp=True)) ->
    uint8 NHWC, renormalised to the batch's own [min, max] exactly as the reference's numpy code does."""
    rgb = rgb.float().contiguous()
    B, C, H, W = rgb.shape
    if C != 3:
        raise ValueError("fused_to_uint8 expects [B, 3, H, W]")
    st = _prep(rgb)
    out = torch.empty((B, H, W, 3), dtype=torch.uint8, device=rgb.device)
    mm = torch.empty((2,), dtype=torch.int32, device=rgb.device)
    _lib.call("segmif_fused_to_uint8", _ptr(rgb), ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(mm.data_ptr()), B, H * W, st)
    return out
