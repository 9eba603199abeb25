"""Times one DRDB (core/model_fusion.py:134-157) at BASELINE configs[1] size (batch 8, 480x640) in its sequential
('hybrid') and concurrent ('dataflow') forms, with CUDA events over 10 calls after 3 warm-ups, checks that both give the
same bits, and sweeps the SM split of the dataflow form.  The growth + partial buffers (1.7 GB) exceed the L2 many times
over, so consecutive calls do not reuse each other's data.

    python tools/drdb_bench.py [--splits "25,17,12,16,20,26,32;24,16,10,14,20,28,36"]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from segmif_b200 import ops, synth  # noqa: E402
from segmif_b200.core.model_fusion import DRDB, Fusion_Network3_ac  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--splits", default="")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--dataflow", type=int, default=0, help="1: also time the concurrent (dataflow) form")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    fus = synth.load_synthetic(Fusion_Network3_ac(), 0).eval().to(dev)
    d = fus.DRDB1
    B, H, W = a.batch, a.height, a.width
    g = torch.Generator(device=dev).manual_seed(0)
    buf = torch.zeros((B, H, W, 224), dtype=torch.bfloat16, device=dev)
    buf[..., :64] = (torch.rand((B, H, W, 64), generator=g, device=dev) - 0.3).bfloat16()
    part = torch.empty((B, H, W, 128), dtype=torch.bfloat16, device=dev)
    out = torch.empty((B * H * W, 64), dtype=torch.bfloat16, device=dev)
    flops = 2.0 * B * H * W * (9 * 640 * 32 + 224 * 64)

    def run(mode, ctas=None):
        DRDB.MODE = mode
        if ctas is not None:
            ops._DF_CTAS[:] = ctas
        else:
            ops._DF_CTAS[:] = []
        with torch.no_grad():
            for _ in range(3):
                d.forward_buffer(buf, B, H, W, out=out, ld_dst=64, partials=part)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.iters):
                d.forward_buffer(buf, B, H, W, out=out, ld_dst=64, partials=part)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        return ms, out.clone()

    res = {}

    def ev(fn):
        with torch.no_grad():
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.iters
    DRDB.MODE = "hybrid"
    ms_g = ev(lambda: d._growth_hybrid(buf, part, B, H, W))
    gf = 2.0 * B * H * W * 9 * 640 * 32
    print("growth only (hybrid)", f"{ms_g:.3f} ms  {gf / ms_g / 1e9:.0f} TFLOP/s", flush=True)
    ms_p = ev(lambda: d._growth_push(buf, part, B, H, W))
    print("growth only (push-all)", f"{ms_p:.3f} ms  {gf / ms_p / 1e9:.0f} TFLOP/s", flush=True)
    res["growth_hybrid_ms"], res["growth_push_ms"] = ms_g, ms_p
    ms_q = ev(lambda: d._growth_pair(buf, part, B, H, W))
    print("growth only (pair)", f"{ms_q:.3f} ms  {gf / ms_q / 1e9:.0f} TFLOP/s", flush=True)
    res["growth_pair_ms"] = ms_q
    with torch.no_grad():
        d._growth_hybrid(buf, part, B, H, W)
        ref_g = buf[..., 64:].float().clone()
        d._growth_pair(buf, part, B, H, W)
        err = float((buf[..., 64:].float() - ref_g).abs().max() / ref_g.abs().max())
    print("pair vs hybrid growth slabs: max rel err", f"{err:.3e}", flush=True)
    res["pair_vs_hybrid_rel_err"] = err
    ms_h, ref = run("hybrid")
    res["hybrid"] = {"ms": ms_h, "tflops": flops / ms_h / 1e9}
    print("hybrid  ", f"{ms_h:.3f} ms  {flops / ms_h / 1e9:.0f} TFLOP/s", flush=True)
    splits = ([None] if a.dataflow else []) + [[int(v) for v in s.split(",")] for s in a.splits.split(";") if s.strip()]
    for sp in splits:
        ms, got = run("dataflow", sp)
        same = bool((got == ref).all())
        key = "dataflow_" + ("default" if sp is None else "-".join(map(str, sp)))
        res[key] = {"ms": ms, "tflops": flops / ms / 1e9, "equals_hybrid": same, "timed_out": d.dataflow_timed_out()}
        print(key, f"{ms:.3f} ms  {flops / ms / 1e9:.0f} TFLOP/s  equal={same} timeout={d.dataflow_timed_out()}", flush=True)
        print("   stage [begin, end] us:", " ".join("-" if t is None else f"[{t[0]:.0f},{t[1]:.0f}]" for t in d.dataflow_stage_times()), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "drdb_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
