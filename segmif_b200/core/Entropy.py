"""Soft-histogram patch entropy with the reference's surface (core/Entropy.py of SegMiF): non-overlapping
p x p patches, 32 bins on [0,1], Gaussian kernel sigma=0.01, -sum p log p summed over patches and batch.
The reference materialises a [B*L, p*p, 32] fp32 tensor (8.6 GB at batch 64 of 1024x1024); the kernel keeps one
histogram per warp in registers (lane == bin) and reads each pixel once."""
import torch
from torch import nn, Tensor

from ..autograd import EntropyFn


class Entropy(nn.Sequential):
    def __init__(self, patch_size):
        super().__init__()
        self.psize = patch_size

    def forward(self, inputs: Tensor) -> torch.Tensor:
        self.width, self.height = inputs.shape[3], inputs.shape[2]
        self.patch_num = int(self.width * self.height / self.psize ** 2)
        return EntropyFn.apply(inputs, self.psize)
