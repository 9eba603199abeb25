"""Generates tests/golden/train_grads.npz: gradients computed by the UNMODIFIED reference's own autograd
(oracle/ref_shim.py imports it from /root/reference) for the two training losses of train.py, on seeded synthetic weights
and inputs -- the pin of the oracle's BACKWARD (tests/test_oracle_golden.py compares torch.autograd over the oracle with
these; the GPU training-parity checks then compare the CUDA backward with the oracle).
  * train_seg (train.py:222-226): Network3('mit_b1') in train mode, DropPath / Dropout2d rates set to 0 (the only
    stochastic parts), batch-statistics BatchNorm: loss = model._loss(mask, labels, CrossEntropyLoss(ignore_index=255))
  * train_fusion rounds >= 2 (train.py:361-368): Fusion_Network3_ac + Fusionloss_grad3 on given encoder features
Stored: the loss values, every parameter gradient's max |.| and sum, and a strided sample of the larger tensors.
Run in the build container only:  python -m oracle.make_golden_train"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shim, segmif_oracle as O            # noqa: E402
from segmif_b200 import synth                              # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
B, H, W = 2, 64, 96


def _summ(prefix, named_grads, gold):
    names = []
    for k, g in named_grads:
        names.append(k)
        g = g.detach().double()
        gold[f"{prefix}|{k}|stat"] = np.array([float(g.abs().max()), float(g.sum())])
        flat = g.reshape(-1)
        gold[f"{prefix}|{k}|sample"] = flat[:: max(1, flat.numel() // 64)][:64].numpy().astype(np.float32)
    gold[f"{prefix}|names"] = np.array(names)


def seg_case(ns, gold):
    with contextlib.redirect_stdout(io.StringIO()):
        seg = ns.model_fusion.Network3("mit_b1", 9, 256, None)
    synth.load_synthetic(seg, 0)
    seg.train()
    seg.denoise_net.decoder.dropout.p = 0.0
    for m in seg.modules():
        if hasattr(m, "drop_prob"):
            m.drop_prob = 0.0
    inp = synth.synth_inputs(B, H, W, seed=7)
    crit = torch.nn.CrossEntropyLoss(ignore_index=255)
    sd = {k: v.clone() for k, v in seg.state_dict().items()}
    with ref_shim.cuda_is_identity():
        loss = seg._loss(inp["mask"], inp["labels"], crit)                      # core/model_fusion.py:1090-1097
        loss.backward()
    gold["seg|loss"] = np.array(float(loss))
    named = [(k, p.grad) for k, p in seg.named_parameters() if p.grad is not None]
    _summ("seg", named, gold)
    # oracle autograd on the same case
    names = [k for k, _ in named]
    leaves = {k: sd[k].clone().requires_grad_(True) for k in names}
    full = dict(sd)
    full.update(leaves)
    lo = O.seg_cross_entropy(O.network3_forward(inp["mask"], full, "mit_b1", train_bn=True), inp["labels"])
    lo.backward()
    worst = max(float((leaves[k].grad - g).abs().max()) / (float(g.abs().max()) + 1e-12) for k, g in named
                if float(g.abs().max()) > 1e-6)
    print(f"  train_seg: loss ref {float(loss):.6f} oracle {float(lo):.6f}; worst relative gradient difference {worst:.2e} "
          f"over {len(names)} tensors")


def fusion_case(ns, gold):
    lossmod = ref_shim.load_reference_losses()
    with contextlib.redirect_stdout(io.StringIO()):
        fus = ns.model_fusion.Fusion_Network3_ac()
    synth.load_synthetic(fus, 0)
    fus.train()
    inp = synth.synth_inputs(B, 40, 56, seed=2)
    g = torch.Generator().manual_seed(52)
    out1 = (torch.randn(B, 64, 40, 56, generator=g) * 0.5)
    out2 = (torch.randn(B, 128, 40, 56, generator=g) * 0.5)
    sd = {k: v.clone() for k, v in fus.state_dict().items()}
    cpu = torch.device("cpu")
    orig_defaults = ns.lap_loss.LapLoss2.__init__.__defaults__
    ns.lap_loss.LapLoss2.__init__.__defaults__ = (3, 1, cpu)                    # Fusionloss_grad3 builds an unused LapLoss2(cuda)
    try:
        with ref_shim.cuda_is_identity():
            vis = ns.model_fusion.RGB2YCrCb(inp["vis"])                         # train.py:356
            fused = fus(inp["ir"], vis, out1, out2)                             # train.py:360
            loss = lossmod.Fusionloss_grad3()(inp["ir"], vis, fused, inp["mask"])   # train.py:362,367
            loss.backward()
    finally:
        ns.lap_loss.LapLoss2.__init__.__defaults__ = orig_defaults
    gold["fusion|loss"] = np.array(float(loss))
    named = [(k, p.grad) for k, p in fus.named_parameters() if p.grad is not None]
    _summ("fusion", named, gold)
    names = [k for k, _ in named]
    leaves = {k: sd[k].clone().requires_grad_(True) for k in names}
    full = dict(sd)
    full.update(leaves)
    visr = O.rgb2ycrcb(inp["vis"])
    lo = O.fusionloss_grad3(inp["ir"], visr, O.fusion_network3_ac(inp["ir"], visr, out1, out2, full), inp["mask"])
    lo.backward()
    worst = max(float((leaves[k].grad - g).abs().max()) / (float(g.abs().max()) + 1e-12) for k, g in named
                if float(g.abs().max()) > 1e-9)
    print(f"  train_fusion: loss ref {float(loss):.6f} oracle {float(lo):.6f}; worst relative gradient difference {worst:.2e} "
          f"over {len(names)} tensors (ffm2.* receive none: {all(not k.startswith('ffm2.') for k in names)})")


def main():
    torch.set_num_threads(os.cpu_count())
    ns = ref_shim.load_reference()
    gold = {}
    seg_case(ns, gold)
    fusion_case(ns, gold)
    out = os.path.join(GOLDEN_DIR, "train_grads.npz")
    np.savez_compressed(out, **gold)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
