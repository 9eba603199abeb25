"""Runs each cfg-5 loss kernel (batch 64 of 1x1024x1024 fp32) twice -- the target of `ncu --set full -k regex:...` captures."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from segmif_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
g = torch.Generator(device="cuda").manual_seed(0)
a, b, c = (torch.rand((B, 1, 1024, 1024), generator=g, device="cuda") for _ in range(3))
for _ in range(2):
    ops.ssim(a, b)
    ops.laploss2(a, b, c)
    ops.entropy(a, 4)
    ops.sobel_l1(a, b)
torch.cuda.synchronize()
