"""Generates tests/golden/*.npz by running the UNMODIFIED SegMiF reference (imported from
/root/reference through oracle/ref_shim.py) on seeded synthetic weights and inputs, and
prints how far the oracle restatement is from each stored result.

Run in the build container only (the reference is not present on the GPU box):
    python -m oracle.make_golden
The fixtures are committed; tests/test_oracle_golden.py pins the oracle to them.
"""
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
BACKBONE = "mit_b1"      # smallest backbone Fusion_Network3_ac accepts unmodified (SURVEY.md finding 6)
H, W = 64, 96


def _np(t):
    return t.detach().cpu().numpy()


def _err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def build_reference_models(ns, seed=0):
    with contextlib.redirect_stdout(io.StringIO()):       # DRDB.__init__ prints (model_fusion.py:131)
        seg = ns.model_fusion.Network3(BACKBONE, 9, 256, None)
        fus = ns.model_fusion.Fusion_Network3_ac()
    synth.load_synthetic(seg, seed)
    synth.load_synthetic(fus, seed)
    return seg.eval(), fus.eval()


def pipeline_case(ns):
    seg, fus = build_reference_models(ns)
    inp = synth.synth_inputs(1, H, W, seed=0)
    ir, vis, mask = inp["ir"], inp["vis"], inp["mask"]
    with torch.no_grad(), ref_shim.cuda_is_identity():
        out0, out1 = seg.denoise_net.encoder.forward_fusion(mask)                 # train.py:358-359
        feats_mask = seg.denoise_net.encoder.forward_features(mask)
        vis_ycc = ns.model_fusion.RGB2YCrCb(vis)                                  # train.py:356
        fused = fus(ir, vis_ycc, out0, out1)                                      # train.py:360
        ycc = vis_ycc.clone()
        ycc[:, 0:1] = fused                                                        # train.py:364-366
        rgb = ns.model_fusion.YCrCb2RGB(ycc).clamp(0, 1)                           # test_fusion.py:104-111
        logits = seg(rgb.clone())[2]                                               # test_segmentation.py:169
        up = torch.nn.functional.interpolate(logits, size=(H, W), mode="bilinear", align_corners=False)
        labels = up.argmax(1)
        ce = seg._loss(rgb.clone(), inp["labels"], torch.nn.CrossEntropyLoss(ignore_index=255))
        # sub-module results (same weights) for per-op parity tests
        x64 = out0                                                                 # any [1,64,H,W] tensor
        drdb1 = fus.DRDB1(x64)
        seg3 = fus.conv3(out0)
        ffm1, ffm2 = fus.ffm(x64, drdb1, seg3)
        head_in = seg.denoise_net.encoder.forward_features((rgb * 255 - 110.0) / 58.0)
        head_out = seg.denoise_net.decoder(head_in)
    gold = dict(
        out0_s=_np(out0[:, ::4, ::4, ::4]), out1_s=_np(out1[:, ::8, ::4, ::4]),
        feat0=_np(feats_mask[0]), feat1=_np(feats_mask[1]), feat2=_np(feats_mask[2]), feat3=_np(feats_mask[3]),
        vis_ycc_s=_np(vis_ycc[:, :, ::4, ::4]), fused=_np(fused), rgb_s=_np(rgb[:, :, ::2, ::2]),
        logits=_np(logits), labels=_np(labels).astype(np.int16), ce=_np(ce),
        drdb1_s=_np(drdb1[:, ::4, ::4, ::4]), ffm1_s=_np(ffm1[:, ::4, ::4, ::4]), ffm2_s=_np(ffm2[:, ::4, ::4, ::4]),
        head_out=_np(head_out),
    )
    # oracle vs reference, live
    seg_sd = {k: v for k, v in seg.state_dict().items()}
    fus_sd = {k: v for k, v in fus.state_dict().items()}
    with torch.no_grad():
        o = O.inference_pipeline(ir, vis, mask, seg_sd, fus_sd, BACKBONE)
        print("  out0   ", _err(o["out0"], out0), " out1 ", _err(o["out1"], out1))
        print("  fused  ", _err(o["fused"], fused), " rgb ", _err(o["rgb"], rgb))
        print("  logits ", _err(o["logits"], logits), " labels equal:", bool((o["labels"] == labels).all()))
        print("  ce     ", _err(O.seg_cross_entropy(o["logits"], inp["labels"]), ce))
        print("  drdb   ", _err(O.drdb(x64, fus_sd, "DRDB1"), drdb1))
        f1, f2 = O.feature_fusion_module(x64, drdb1, seg3, fus_sd, "ffm")
        print("  ffm    ", _err(f1, ffm1), _err(f2, ffm2))
        print("  head   ", _err(O.segformer_head(head_in, O._sub(seg_sd, "denoise_net.decoder")), head_out))
    return gold


def loss_case(ns):
    lossmod = ref_shim.load_reference_losses()
    a, b, c = synth.analytic_images()
    inp = synth.synth_inputs(2, 48, 80, seed=3)
    r_ir, r_vis, r_mask = inp["ir"], inp["vis"], inp["mask"]
    r_fused = (0.6 * r_ir + 0.4 * r_vis[:, :1]).clamp(0, 1)
    cpu = torch.device("cpu")
    gold = {}
    with torch.no_grad(), ref_shim.cuda_is_identity():
        gold["kat_ssim"] = _np(ns.pytorch_ssim.ssim(a, b))
        gold["kat_ssim_per_image"] = _np(ns.pytorch_ssim.ssim(a, b, size_average=False))
        gold["kat_lap2"] = _np(ns.lap_loss.LapLoss2(device=cpu)(a, b, c))
        gold["kat_lap"] = _np(ns.lap_loss.LapLoss(device=cpu)(a, b))
        gold["kat_entropy4"] = _np(ns.Entropy.Entropy(4)(a))
        gold["kat_entropy8"] = _np(ns.Entropy.Entropy(8)(a))
        gold["rnd_ssim"] = _np(ns.pytorch_ssim.ssim(r_fused, r_mask[:, :1]))
        gold["rnd_lap2"] = _np(ns.lap_loss.LapLoss2(device=cpu)(r_fused, r_ir, r_vis[:, :1]))
        gold["rnd_entropy4"] = _np(ns.Entropy.Entropy(4)(r_fused))
        sob = lossmod.Sobelxy()
        gold["rnd_sobel_s"] = _np(sob(r_fused)[:, :, ::2, ::2])
        gold["rnd_fusionloss3"] = _np(lossmod.Fusionloss3()(r_ir, r_vis, r_fused, r_mask))
        # Fusionloss_grad3/grad2 construct LapLoss2() with the default device=cuda (lap_loss.py:101);
        # torch.device('cuda') objects are fine on CPU until a tensor is moved, and .to(device) of the
        # frozen conv would fail -- so build them with the default device patched to cpu.
        orig_defaults = ns.lap_loss.LapLoss2.__init__.__defaults__
        ns.lap_loss.LapLoss2.__init__.__defaults__ = (3, 1, cpu)
        try:
            gold["rnd_fusionloss_grad3"] = _np(lossmod.Fusionloss_grad3()(r_ir, r_vis, r_fused, r_mask))
            gold["rnd_fusionloss_grad2"] = _np(lossmod.Fusionloss_grad2()(r_ir, r_vis, r_fused, r_mask))
        finally:
            ns.lap_loss.LapLoss2.__init__.__defaults__ = orig_defaults
        print("  ssim   ", _err(O.ssim(a, b), torch.tensor(gold["kat_ssim"])))
        print("  lap2   ", _err(O.lap_loss2(a, b, c), torch.tensor(gold["kat_lap2"])))
        print("  lap    ", _err(O.lap_loss(a, b), torch.tensor(gold["kat_lap"])))
        print("  entr4  ", _err(O.entropy(a, 4), torch.tensor(gold["kat_entropy4"])))
        print("  entr8  ", _err(O.entropy(a, 8), torch.tensor(gold["kat_entropy8"])))
        print("  fl3    ", _err(O.fusionloss3(r_ir, r_vis, r_fused, r_mask), torch.tensor(gold["rnd_fusionloss3"])))
        print("  flg3   ", _err(O.fusionloss_grad3(r_ir, r_vis, r_fused, r_mask), torch.tensor(gold["rnd_fusionloss_grad3"])))
        print("  flg2   ", _err(O.fusionloss_grad2(r_ir, r_vis, r_fused, r_mask), torch.tensor(gold["rnd_fusionloss_grad2"])))
    return gold


def keys_case(ns):
    """state-dict key lists + shapes of the reference modules (drop-in contract)."""
    out = {}
    with contextlib.redirect_stdout(io.StringIO()):
        fus = ns.model_fusion.Fusion_Network3_ac()
        out["fusion_keys"] = np.array([f"{k}|{','.join(map(str, v.shape))}" for k, v in fus.state_dict().items()])
        for bb in ("mit_b0", "mit_b1", "mit_b2"):
            seg = ns.model_fusion.Network3(bb, 9, 256, None)
            out[f"network3_{bb}_keys"] = np.array([f"{k}|{','.join(map(str, v.shape))}" for k, v in seg.state_dict().items()])
    return out


def main():
    torch.set_num_threads(os.cpu_count())
    ns = ref_shim.load_reference()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    print("pipeline case (", BACKBONE, H, "x", W, ")")
    np.savez_compressed(os.path.join(GOLDEN_DIR, "pipeline_mit_b1_64x96.npz"), **pipeline_case(ns))
    print("loss case")
    np.savez_compressed(os.path.join(GOLDEN_DIR, "losses.npz"), **loss_case(ns))
    np.savez_compressed(os.path.join(GOLDEN_DIR, "state_dict_keys.npz"), **keys_case(ns))
    for f in sorted(os.listdir(GOLDEN_DIR)):
        print(f, os.path.getsize(os.path.join(GOLDEN_DIR, f)), "bytes")


if __name__ == "__main__":
    main()
