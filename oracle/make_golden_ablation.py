"""Generates tests/golden/ablation.npz: outputs of the UNMODIFIED reference ablation networks (core/model_fusion.py:626-1025:
Fusion_Network3, _S, _M, _Con, _Add, _Average, Fusion_Network_rmseg, AttentionModule, CrossPath_M / CrossPath_S) on seeded
synthetic weights and inputs, and prints how far the oracle restatement is from each.

    python -m oracle.make_golden_ablation        (build container only: needs /root/reference)"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shim, segmif_oracle as O            # noqa: E402
from segmif_b200 import synth                              # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ablation.npz")
VARIANTS = {"Fusion_Network3": "base", "Fusion_Network3_S": "S", "Fusion_Network3_M": "M", "Fusion_Network3_Con": "Con",
            "Fusion_Network3_Add": "Add", "Fusion_Network3_Average": "Average"}


def inputs(B=1, H=32, W=48):
    inp = synth.synth_inputs(B, H, W, seed=21)
    g = torch.Generator().manual_seed(77)
    out1 = torch.randn(B, 64, H, W, generator=g) * 0.5
    out2 = torch.randn(B, 128, H, W, generator=g) * 0.5
    return inp["ir"], inp["vis"], out1, out2


def _err(a, b):
    return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30))


def main():
    ns = ref_shim.load_reference()
    ir, vis, out1, out2 = inputs()
    gold = {}
    with torch.no_grad():
        for name, variant in VARIANTS.items():
            with contextlib.redirect_stdout(io.StringIO()):
                net = synth.load_synthetic(getattr(ns.model_fusion, name)(), 3).eval()
            ref = net(ir, vis, out1, out2)
            gold[name] = ref.numpy()
            got = O.fusion_network3_variant(ir, vis, out1, out2, dict(net.state_dict()), variant)
            print(f"  {name:<26} oracle-vs-reference {_err(got, ref):.1e}")
        with contextlib.redirect_stdout(io.StringIO()):
            net = synth.load_synthetic(ns.model_fusion.Fusion_Network_rmseg(), 3).eval()
        ref = net(ir, vis)
        gold["Fusion_Network_rmseg"] = ref.numpy()
        print(f"  {'Fusion_Network_rmseg':<26} oracle-vs-reference {_err(O.fusion_network_rmseg(ir, vis, dict(net.state_dict()))[0], ref):.1e}")
        g = torch.Generator().manual_seed(5)
        t1, t2, t3 = (torch.randn(2, 150, 32, generator=g) for _ in range(3))
        for name, mode in (("CrossPath_M", "M"), ("CrossPath_S", "S")):
            cp = synth.load_synthetic(getattr(ns.model_fusion, name)(32), 3).eval()
            sd = {"cross." + k: v * (10.0 if ".kv" in k else 1.0) for k, v in cp.state_dict().items()}      # livelier contexts
            cp.load_state_dict({k[6:]: v for k, v in sd.items()})
            r1, r2 = cp(t1, t2, t3)
            gold[name + "_1"], gold[name + "_2"] = r1.numpy(), r2.numpy()
            o1, o2 = O.cross_path_variant(t1, t2, t3, sd, "cross", mode)
            print(f"  {name:<26} oracle-vs-reference {max(_err(o1, r1), _err(o2, r2)):.1e}")
        am = synth.load_synthetic(ns.model_fusion.AttentionModule(), 3).eval()
        x = torch.randn(1, 32, 20, 28, generator=g)
        gold["AttentionModule"] = am(x).numpy()
        print(f"  {'AttentionModule':<26} oracle-vs-reference {_err(O.attention_module(x, {'att.' + k: v for k, v in am.state_dict().items()}, 'att'), am(x)):.1e}")
    np.savez_compressed(GOLDEN, **gold)
    print(GOLDEN, os.path.getsize(GOLDEN), "bytes")


if __name__ == "__main__":
    main()
