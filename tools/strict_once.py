"""One strict-precision pass of the pipeline (batch 2, 480x640, MiT-B2) for ncu captures of the split-bf16 kernels."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import segmif_b200  # noqa: E402
from segmif_b200 import synth  # noqa: E402
from segmif_b200.core.model_fusion import Fusion_Network3_ac, Network3  # noqa: E402
from segmif_b200.pipeline import FusionSegPipeline  # noqa: E402

dev = torch.device("cuda", 0)
seg = synth.load_synthetic(Network3("mit_b2", 9, 256, None), 0).eval().to(dev)
fus = synth.load_synthetic(Fusion_Network3_ac(), 0).eval().to(dev)
pipe = FusionSegPipeline(seg, fus)
inp = {k: v.to(dev) for k, v in synth.synth_inputs(2, 480, 640, seed=0).items()}
with segmif_b200.precision("strict"), torch.no_grad():
    for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
        pipe(inp["ir"], inp["vis"], inp["mask"])
torch.cuda.synchronize()
print("ok")
