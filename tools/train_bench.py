"""Times the data-parallel training steps of train.py (BASELINE configs[2]; not the headline bench -- bench.py measures
configs[1]), per-GPU batch 4 at 480x640:
  --mode fusion     train_fusion round 1 (train.py:343-386, Fusionloss3; --loss grad3 = rounds >= 2 without the CE term):
                    frozen-encoder forward_fusion (no_grad) -> Fusion_Network3_ac forward -> loss -> hand-written
                    backward -> ONE gradient all-reduce -> fused AdamW
  --mode fusion_ce  train_fusion rounds >= 2 (train.py:361-380): + YCrCb2RGB -> Network3._loss (CE) differentiated
                    through the frozen segmentation network (train mode: batch-statistics BN, DropPath, Dropout2d)
  --mode seg        train_seg (train.py:207-226): Network3 forward -> upsample + CE -> backward -> all-reduce of the
                    24.7 M-parameter flat buffer -> AdamW over three param groups
    python tools/train_bench.py [--batch 4] [--steps 10]            # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/train_bench.py"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--backbone", default="mit_b2")
    ap.add_argument("--loss", default="loss3", choices=["loss3", "grad3", "grad2"])
    ap.add_argument("--mode", default="fusion", choices=["fusion", "fusion_ce", "seg"])
    ap.add_argument("--graph", type=int, default=1, help="1: forward+backward replayed from one CUDA graph; 0: eager launches")
    ap.add_argument("--profile", type=int, default=0, help="N > 0: after the timed steps, run N more under torch.profiler (CUPTI) and "
                    "write the per-kernel totals of those steps to gpurun_out/train_profile_<mode>.json")
    a = ap.parse_args()
    from segmif_b200 import _lib, synth
    from segmif_b200.core.loss import Fusionloss3, Fusionloss_grad2, Fusionloss_grad3
    from segmif_b200.core.model_fusion import Fusion_Network3_ac, Network3
    from segmif_b200.ddp import FusionTrainer, SegTrainer
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(1234 + rank)                                                   # DropPath / Dropout2d draws
    seg = synth.load_synthetic(Network3(a.backbone, 9, 256, None), 0).to(dev)
    seg = seg.train() if a.mode != "fusion" else seg.eval()       # train.py never calls .eval() on it; "fusion" keeps v1's setup
    inp = {k: v.to(dev) for k, v in synth.synth_inputs(a.batch, a.height, a.width, seed=rank).items()}
    if a.mode == "seg":
        tr = SegTrainer(seg, lr=6e-5, weight_decay=0.01, betas=(0.9, 0.999), warmup_iter=1500, max_iter=160000,
                        warmup_ratio=1e-6, power=1.0)

        def step():
            return (tr.step(inp["mask"], inp["labels"]),)

        capture = lambda: tr.capture(inp["mask"], inp["labels"])
    else:
        fus = synth.load_synthetic(Fusion_Network3_ac(), 0).train().to(dev)
        if a.mode == "fusion_ce":
            tr = FusionTrainer(fus, Fusionloss_grad3(), lr=1.5e-4, weight_decay=0.01, betas=(0.9, 0.999), warmup_iter=1.5e-5,
                               max_iter=6000, warmup_ratio=1e-6, power=1.0, seg_net=seg, iter_=2)
        else:
            crit = {"loss3": Fusionloss3, "grad3": Fusionloss_grad3, "grad2": Fusionloss_grad2}[a.loss]()
            tr = FusionTrainer(fus, crit, lr=3e-4, weight_decay=0.01, betas=(0.9, 0.999), warmup_iter=3e-5, max_iter=6000,
                               warmup_ratio=1e-6, power=1.0, seg_net=seg, with_ce=False)
        lab = inp["labels"] if a.mode == "fusion_ce" else None

        def step():
            return tr.step_images(inp["ir"], inp["vis"], inp["mask"], lab)

        capture = lambda: tr.capture_images(inp["ir"], inp["vis"], inp["mask"], lab)

    losses = []
    losses.append(step()[0])                      # eager: lazy per-kernel initialisation happens here
    if a.graph:
        capture()
    for _ in range(max(a.warmup, 1)):
        losses.append(step()[0])
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = _lib.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        losses.append(step()[0])
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        # replicas must stay bit-identical: same all-reduced gradients, same update
        chk = tr.flat.param.double().sum().reshape(1)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        in_sync = bool((hi - lo).abs().item() == 0.0)
    else:
        in_sync = True
    if rank == 0:
        ms = float(ms.item())
        print(json.dumps({"metric": {"fusion": "train_fusion_pairs_per_sec", "fusion_ce": "train_fusion_ce_pairs_per_sec",
                                     "seg": "train_seg_images_per_sec"}[a.mode], "value": a.batch * world * a.steps / (ms * 1e-3), "unit": "pairs/s",
                          "n_gpus": world, "steps": a.steps, "ms_per_step": ms / a.steps, "batch_per_gpu": a.batch,
                          "height": a.height, "width": a.width, "loss": a.loss if a.mode == "fusion" else "ce", "mode": a.mode, "cuda_graph": bool(a.graph), "backbone": a.backbone, "scaling": "weak",
                          "allreduce_mb": tr.flat.numel * 4 / 2 ** 20,
                          "eager_launches_per_step": (_lib.launch_count - l0) / a.steps, "replicas_in_sync": in_sync,
                          "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30,
                          "loss_first_last": [float(losses[0]), float(losses[-1])]}), flush=True)
    if a.profile and rank == 0:
        from torch.profiler import ProfilerActivity, profile
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(a.profile):
                step()
            torch.cuda.synchronize()
        rows = sorted(((e.key, e.count, getattr(e, "device_time_total", None) or getattr(e, "cuda_time_total", 0.0)) for e in prof.key_averages()),
                      key=lambda r: -r[2])
        tot = sum(r[2] for r in rows)
        out = {"mode": a.mode, "steps": a.profile, "kernel_time_ms_per_step": tot / 1e3 / a.profile,
               "kernels": [{"name": k[:110], "launches_per_step": c / a.profile, "ms_per_step": t / 1e3 / a.profile, "share": t / tot} for k, c, t in rows[:45]]}
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"train_profile_{a.mode}.json"), "w"), indent=1)
        print(f"kernel time per step {out['kernel_time_ms_per_step']:.3f} ms (sum over streams)")
        for k in out["kernels"][:32]:
            print(f"  {k['name'][:84]:84s} n={k['launches_per_step']:6.1f} {k['ms_per_step']:7.3f} ms {100 * k['share']:5.1f}%")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
