"""LapLoss / LapLoss2 with the reference's surface (lap_loss.py of SegMiF): three same-resolution
difference-of-Gaussian residuals (k = 3, 5, 7; sigma = 2; zero padding) and L1 terms weighted 10, 10, 1.
One kernel produces all three residuals of all images from a single 3-pixel-halo shared-memory tile."""
import torch

from . import ops  # noqa: F401
from .autograd import LapLoss2Fn, LapLossFn


def _planes(t, channels):
    if t.shape[1] != channels:
        raise ValueError(f"expected {channels} channel(s), got {t.shape[1]}")
    B, C, H, W = t.shape
    return t.float().reshape(B * C, 1, H, W)


class LapLoss(torch.nn.Module):
    def __init__(self, max_levels=3, channels=1, device=None):
        super().__init__()
        if max_levels != 3:
            raise NotImplementedError("segmif_b200: LapLoss kernels implement the reference's three levels")
        self.max_levels, self.channels = max_levels, channels

    def forward(self, input, target):
        return LapLossFn.apply(_planes(input, self.channels), _planes(target, self.channels))


class LapLoss2(torch.nn.Module):
    def __init__(self, max_levels=3, channels=1, device=None):
        super().__init__()
        if max_levels != 3:
            raise NotImplementedError("segmif_b200: LapLoss kernels implement the reference's three levels")
        self.max_levels, self.channels = max_levels, channels

    def forward(self, input, ir, vis):
        return LapLoss2Fn.apply(_planes(input, self.channels), _planes(ir, self.channels), _planes(vis, self.channels))
