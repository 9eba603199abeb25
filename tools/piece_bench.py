"""CUDA-event timing of individual inference-side kernels at BASELINE configs[1] sizes (batch 8, 480x640) and the
configs[4] loss kernels: achieved GB/s over each kernel's algorithmic bytes.  Writes gpurun_out/piece_bench.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from segmif_b200 import ops, synth  # noqa: E402
from segmif_b200.core.model_fusion import Fusion_Network3_ac  # noqa: E402

DEV = "cuda"


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3        # us


def main():
    out = {}
    B, H, W = 8, 480, 640
    M = B * H * W
    fus = synth.load_synthetic(Fusion_Network3_ac(), 0).eval().to(DEV)
    alpha = fus.relu.weight.detach()
    g = torch.Generator(device=DEV).manual_seed(0)
    ir = torch.rand((B, 1, H, W), generator=g, device=DEV)
    buf = torch.empty((B, H, W, 224), dtype=torch.bfloat16, device=DEV)
    w9 = fus._packs.taps_f32(fus.conv1_ir.weight)
    us = timeit(lambda: ops.conv3x3_in1(ir, w9, fus.conv1_ir.bias.detach(), alpha, buf, 224, 0, 64))
    out["conv3x3_in1"] = {"us": us, "gbs": (M * 4 + M * 128) / us / 1e3}
    f32 = torch.rand((M, 32), generator=g, device=DEV).bfloat16()
    w22 = fus._packs.taps_f32(fus.conv22.weight)
    us = timeit(lambda: ops.conv3x3_out1(f32, w22, fus.conv22.bias.detach(), alpha, B, H, W, 32))
    out["conv3x3_out1"] = {"us": us, "gbs": (M * 64 + M * 4) / us / 1e3}
    x, y, z = (torch.rand((64, 1, 1024, 1024), generator=g, device=DEV) for _ in range(3))
    n = 64 * 1024 * 1024
    for name, fn, planes in (("ssim", lambda: ops.ssim(x, y), 2), ("laploss2", lambda: ops.laploss2(x, y, z), 3),
                             ("entropy4", lambda: ops.entropy(x, 4), 1), ("entropy8", lambda: ops.entropy(x, 8), 1),
                             ("entropy16", lambda: ops.entropy(x, 16), 1), ("sobel_l1", lambda: ops.sobel_l1(x, y), 2)):
        us = timeit(fn)
        out[name] = {"us": us, "gbs": planes * n * 4 / us / 1e3}
    for k, v in out.items():
        print(f"{k:<16} {v['us']:9.1f} us  {v['gbs']:8.0f} GB/s", flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "piece_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
