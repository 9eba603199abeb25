"""Per-layer time of the linear weight + bias gradient at train_seg sizes (MiT-B2, batch 4, 480x640): the tcgen05 one-pass
kernel (wgrad_lin_tc.cu) against the mma.sync kernel + separate column sums (SEGMIF_WGRAD_LIN_TC=0 in a second process)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from segmif_b200 import ops  # noqa: E402

LAYERS = [  # name, tokens, Cin, Cout, count per step
    ("s1 q/proj 64x64", 76800, 64, 64, 6), ("s1 fc1 64->256", 76800, 64, 256, 3), ("s1 fc2 256->64", 76800, 256, 64, 3),
    ("s1 kv 64->128 (reduced)", 1200, 64, 128, 3), ("s1 sr 4096->64", 1200, 4096, 64, 3),
    ("s2 q/proj 128x128", 19200, 128, 128, 8), ("s2 fc1 128->512", 19200, 128, 512, 4), ("s2 fc2 512->128", 19200, 512, 128, 4),
    ("s2 kv 128->256", 1200, 128, 256, 4), ("s2 sr 2048->128", 1200, 2048, 128, 4),
    ("s3 q/proj 320x320", 4800, 320, 320, 12), ("s3 fc1 320->1280", 4800, 320, 1280, 6), ("s3 fc2 1280->320", 4800, 1280, 320, 6),
    ("s3 kv 320->640", 1200, 320, 640, 6), ("s3 sr 1280->320", 1200, 1280, 320, 6),
    ("s4 q/proj 512x512", 1200, 512, 512, 6), ("s4 kv 512->1024", 1200, 512, 1024, 3), ("s4 fc1 512->2048", 1200, 512, 2048, 3),
    ("s4 fc2 2048->512", 1200, 2048, 512, 3),
    ("pe2 576->128", 19200, 576, 128, 1), ("pe3 1152->320", 4800, 1152, 320, 1), ("pe4 2880->512", 1200, 2880, 512, 1),
    ("head c1 64->256", 76800, 64, 256, 1), ("head fuse 1024->256", 76800, 1024, 256, 1), ("head pred 256->32", 76800, 256, 32, 1),
]


def main():
    dev = torch.device("cuda", 0)
    out, total = [], 0.0
    for name, P, Cin, Cout, cnt in LAYERS:
        x = torch.randn(P, Cin, device=dev).bfloat16()
        dy = torch.randn(P, Cout, device=dev).bfloat16()
        g = torch.zeros(Cout, Cin, device=dev)
        b = torch.zeros(Cout, device=dev)
        fn = lambda: ops.wgrad_lin(dy, Cout, 0, x, Cin, 0, P=P, Cin=Cin, Cout=Cout, grad=g, s_co=Cin, dbias=b)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        # the training step replays a CUDA graph: time the kernels the same way (eager calls are bound by ~25 us of host work each)
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            with torch.cuda.graph(graph, stream=side):
                for _ in range(10):
                    fn()
        torch.cuda.synchronize()
        graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 40 * 1e3
        byt = P * (Cin + Cout) * 2
        out.append({"layer": name, "us": us, "count": cnt, "GBps": byt / us / 1e3, "TFLOPs": 2.0 * P * Cin * Cout / us / 1e6})
        total += us * cnt
        print(f"{name:28s} {us:8.1f} us x{cnt:2d}  {byt / us / 1e3:7.0f} GB/s  {2.0 * P * Cin * Cout / us / 1e6:7.1f} TFLOP/s")
    print(f"total per step: {total / 1e3:.3f} ms  (mode: {'tcgen05' if os.environ.get('SEGMIF_WGRAD_LIN_TC', '1') != '0' else 'mma.sync + colsum'})")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    tag = "tc" if os.environ.get("SEGMIF_WGRAD_LIN_TC", "1") != "0" else "mma"
    json.dump({"layers": out, "total_ms_per_step": total / 1e3}, open(os.path.join(ROOT, "gpurun_out", f"wgrad_lin_bench_{tag}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
