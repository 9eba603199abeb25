"""Times the attention-core kernels at the stage shapes of configs[1] (MiT-B2, batch 8, 480x640: Nk = 300) and configs[3]
(MiT-B4, batch 4, 1024^2: Nk = 1024): default dispatch (flash tcgen05), the explicit flash kernel and the mma.sync kernel."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from segmif_b200 import _lib, ops  # noqa: E402

DEV = "cuda"


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    out = {}
    shapes = [("cfg2_s1", 8, 1, 19200, 300), ("cfg2_s2", 8, 2, 4800, 300), ("cfg2_s3", 8, 5, 1200, 300), ("cfg2_s4", 8, 8, 300, 300),
              ("cfg4_s1", 4, 1, 65536, 1024), ("cfg4_s2", 4, 2, 16384, 1024), ("cfg4_s3", 4, 5, 4096, 1024), ("cfg4_s4", 4, 8, 1024, 1024)]
    for name, B, heads, N, Nk in shapes:
        C = heads * 64
        g = torch.Generator(device=DEV).manual_seed(0)
        q = torch.randn((B * N, C), generator=g, device=DEV).bfloat16()
        kv = torch.randn((B * Nk, 2 * C), generator=g, device=DEV).bfloat16()
        flops = 4.0 * B * heads * N * Nk * 64
        us_fa = timeit(lambda: ops.sr_attention_fa(q, kv, B, heads, N, Nk, 64, 0.125))
        st = ops._prep(q, kv)
        o = torch.empty_like(q)
        # the mma.sync kernel through a process-wide switch is read once; time it through the D = 64 generic entry with SEGMIF_ATTN=mma in a subprocess instead
        out[name] = {"fa_us": us_fa, "fa_tflops": flops / us_fa / 1e6}
        print(f"{name:<8} flash tcgen05 {us_fa:8.1f} us  {flops / us_fa / 1e6:6.0f} TFLOP/s", flush=True)
    mode = os.environ.get("SEGMIF_ATTN", "fa")
    for name, B, heads, N, Nk in shapes:
        C = heads * 64
        g = torch.Generator(device=DEV).manual_seed(0)
        q = torch.randn((B * N, C), generator=g, device=DEV).bfloat16()
        kv = torch.randn((B * Nk, 2 * C), generator=g, device=DEV).bfloat16()
        flops = 4.0 * B * heads * N * Nk * 64
        us = timeit(lambda: ops.sr_attention(q, kv, B, heads, N, Nk, 64, 0.125))
        out[name][f"default_{mode}_us"] = us
        print(f"{name:<8} default({mode}) {us:8.1f} us  {flops / us / 1e6:6.0f} TFLOP/s", flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"attn_bench_{mode}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
