"""CUDA-event timing of individual training-side kernels at the train_seg shapes (MiT-B2, batch 4, 480x640): a quick
way to compare kernel variants on the GPU box without a profiler.  Prints one JSON line per kernel and shape with the
achieved GB/s over the kernel's algorithmic bytes (unique input + output bytes)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from segmif_b200 import ops  # noqa: E402

DEV = "cuda"


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=DEV)      # 256 MB > 126 MB L2
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters * 1e3                                                       # us


def main(which):
    B = 4
    stages = [(120, 160, 64), (60, 80, 128), (30, 40, 320), (15, 20, 512)]
    out = []
    for (H, W, C) in stages:
        hid, M = 4 * C, B * H * W
        g = torch.Generator(device=DEV).manual_seed(0)
        if "dwconv" in which:
            x = torch.randn(M, hid, device=DEV, generator=g).bfloat16()
            dy = torch.randn(M, hid, device=DEV, generator=g).bfloat16()
            w9c = torch.randn(9, hid, device=DEV, generator=g) * 0.3
            b = torch.randn(hid, device=DEV, generator=g) * 0.1
            dw, db = torch.zeros(9, hid, device=DEV), torch.zeros(hid, device=DEV)
            for name, fn, nbytes in (("dwconv3x3_gelu_bwd", lambda: ops.dwconv3x3_gelu_bwd(x, w9c, b, dy, B, H, W, dw, db), 3 * M * hid * 2),
                                     ("dwconv3x3_flip", lambda: ops.dwconv3x3(dy, w9c, None, B, H, W, flip=True), 2 * M * hid * 2),
                                     ("dwconv3x3_gelu_fwd", lambda: ops.dwconv3x3_gelu(x.view(B, H * W, hid), w9c, b, B, H, W), 2 * M * hid * 2)):
                us = timeit(fn)
                out.append(dict(kernel=name, H=H, W=W, C=hid, us=round(us, 1), gbps=round(nbytes / us * 1e-3, 1)))
        if "colsum" in which:
            for N in (C, hid):
                dy = torch.randn(M, N, device=DEV, generator=g).bfloat16()
                o = torch.zeros(N, device=DEV)
                us = timeit(lambda: ops.colsum(dy, N, 0, M, N, o))
                out.append(dict(kernel="colsum", rows=M, C=N, us=round(us, 1), gbps=round(M * N * 2 / us * 1e-3, 1)))
        if "wgrad" in which:
            for (N, K) in ((C, C), (hid, C), (C, hid)):
                dy = torch.randn(M, N, device=DEV, generator=g).bfloat16()
                x = torch.randn(M, K, device=DEV, generator=g).bfloat16()
                gr = torch.zeros(N, K, device=DEV)
                us = timeit(lambda: ops.wgrad(dy, N, 0, x, K, 0, B=1, H=1, W=1, P=M, Cin=K, Cout=N, taps=1, dil=1, grad=gr, s_co=K, s_tap=1, s_ci=1))
                out.append(dict(kernel="wgrad<1>", rows=M, N=N, K=K, us=round(us, 1), tflops=round(2.0 * M * N * K / us * 1e-6, 1),
                                gbps=round(M * (N + K) * 2 / us * 1e-3, 1)))
    for r in out:
        print(json.dumps(r))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "kernel_bench.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main(sys.argv[1:] or ["dwconv", "colsum", "wgrad"])
