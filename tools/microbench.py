"""Secondary measurements named by BASELINE.json (not the bench line):
  cfg 5  fusion-loss kernels, batch 64 of 1x1024x1024 fp32: achieved HBM GB/s against the measured copy peak
  cfg 4  MiT-B4 backbone (forward_features), batch 4, 1024x1024: images/s and achieved TFLOP/s (631.5 GF/image)
Writes gpurun_out/microbench.json.  CUDA events, >= 3 warm-ups, inputs far larger than the 126 MB L2."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from segmif_b200 import ops, synth  # noqa: E402


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    dev = torch.device("cuda", 0)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
    out = {"peaks": {"hbm_gbs": peaks["hbm_gbs"], "bf16_tflops": peaks["bf16_tflops"]}}
    B, H, W = 64, 1024, 1024
    g = torch.Generator(device=dev).manual_seed(0)
    a, b, c = (torch.rand((B, 1, H, W), generator=g, device=dev) for _ in range(3))
    n = B * H * W
    losses = {}
    for name, fn, planes in (("ssim", lambda: ops.ssim(a, b), 2), ("laploss2", lambda: ops.laploss2(a, b, c), 3),
                             ("entropy4", lambda: ops.entropy(a, 4), 1), ("sobel_l1", lambda: ops.sobel_l1(a, b), 2),
                             ("mse_l1", lambda: ops.mse_l1(a, b), 2)):
        ms = timeit(fn)
        gbs = planes * n * 4 / (ms * 1e-3) / 1e9
        losses[name] = {"ms": ms, "algorithmic_bytes": planes * n * 4, "achieved_gbs": gbs, "frac_of_measured_hbm": gbs / peaks["hbm_gbs"]}
        print(name, f"{ms:.3f} ms  {gbs:.0f} GB/s  ({100 * gbs / peaks['hbm_gbs']:.1f}% of measured copy peak)", flush=True)
    out["cfg5_losses_b64_1024x1024"] = losses
    del a, b, c
    from segmif_b200.core import mix_transformer as MT
    enc = synth.load_synthetic(MT.mit_b4(), 0).eval().to(dev)
    x = torch.rand((4, 3, 1024, 1024), generator=g, device=dev)
    with torch.no_grad():
        ms = timeit(lambda: enc.forward_stages(x), iters=5, warm=3)
    out["cfg4_mit_b4_b4_1024x1024"] = {"ms_per_batch": ms, "images_per_s": 4 / (ms * 1e-3), "achieved_tflops": 4 * 631.5e9 / (ms * 1e-3) / 1e12}
    print("mit_b4 1024x1024 B=4:", f"{ms:.2f} ms/batch  {4 / (ms * 1e-3):.1f} img/s  {4 * 631.5e9 / (ms * 1e-3) / 1e12:.1f} TFLOP/s", flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "microbench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
