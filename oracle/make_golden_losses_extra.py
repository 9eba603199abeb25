"""Generates tests/golden/losses_extra.npz: the fusion-loss composites that train.py imports but never calls
(core/loss.py:386-399 new_loss_sobel, :423-439 Fusionloss, :545-577 Fusionloss4 / Fusionloss_add, :578-599
Total_fusion_loss[2], :605-633 IQALoss), evaluated by the UNMODIFIED reference classes (values and the gradient with
respect to the fused image / the mask from the reference's own autograd) on seeded synthetic inputs.

    python -m oracle.make_golden_losses_extra          (build container only: needs /root/reference)
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shim, segmif_oracle as O            # noqa: E402
from segmif_b200 import synth                              # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "losses_extra.npz")


def inputs():
    inp = synth.synth_inputs(2, 48, 80, seed=3)
    ir, vis, mask = inp["ir"], inp["vis"], inp["mask"]
    fused = (0.6 * ir + 0.4 * vis[:, :1]).clamp(0, 1)
    return ir, vis, mask[:, :1].contiguous(), fused


def _err(a, b):
    return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30))


def main():
    ref_shim.load_reference()
    L = ref_shim.load_reference_losses()
    ir, vis, mask, fused = inputs()
    gold = {}
    cases = {
        "fusionloss": (lambda f: L.Fusionloss()(ir, vis, f), lambda f: O.fusionloss(ir, vis, f)),
        "fusionloss4": (lambda f: L.Fusionloss4()(ir, vis, f, mask), lambda f: O.fusionloss4(ir, vis, f, mask)),
        "fusionloss_add": (lambda f: L.Fusionloss_add()(ir, vis, f), lambda f: O.fusionloss_add(ir, vis, f)),
        "total_fusion_loss": (lambda f: L.Total_fusion_loss()(ir, vis, mask, f), lambda f: O.total_fusion_loss(ir, vis, mask, f)),
        "total_fusion_loss2": (lambda f: L.Total_fusion_loss2()(ir, vis, mask, f), lambda f: O.total_fusion_loss2(ir, vis, mask, f)),
        # IQALoss(lr, vis, mask): the differentiable argument is the mask (here: the fused plane plays the network output)
        "iqa_loss": (lambda f: L.IQALoss()(ir, vis, f), lambda f: O.iqa_loss(ir, vis, f)),
    }
    with ref_shim.cuda_is_identity():
        for name, (ref_fn, orc_fn) in cases.items():
            f = fused.clone().requires_grad_(True)
            val = ref_fn(f)
            (g,) = torch.autograd.grad(val, f)
            gold[name] = val.detach().numpy()
            gold[name + "_grad_max"] = g.abs().max().numpy()
            gold[name + "_grad_s"] = g[:, :, ::3, ::5].numpy()
            f2 = fused.clone().requires_grad_(True)
            v2 = orc_fn(f2)
            (g2,) = torch.autograd.grad(v2, f2)
            print(f"  {name:<20} value {float(val):.6f}  oracle-vs-reference: value {_err(v2, val):.1e}  grad {_err(g2, g):.1e}")
        sob = L.Sobelxy()(fused)
        gold["sobel_map_s"] = sob[:, :, ::2, ::2].detach().numpy()
    np.savez_compressed(GOLDEN, **gold)
    print(GOLDEN, os.path.getsize(GOLDEN), "bytes")


if __name__ == "__main__":
    main()
