"""Parity checks of the strict-precision (fp32-parity) mode against the CPU oracle and the reference-generated fixture.

north_star's contract: outputs within 1e-3 relative of the reference's fp32 path, argmax labels bit-exact.  The bounds
below are far inside that (the split-bf16 contraction is fp32-grade: its error is the rounding-order noise any two fp32
implementations show), and are written next to each check:
  * split_gemm (row / patch mode), all six partial products, against an fp64 contraction of the SAME fp32 operands:
    <= 2e-6 of max |out| (an fp32 GEMM on the CPU measures 1e-6 .. 3e-6 on the same cases);
  * fp32 companions (attention, depthwise, im2col, Gram, edge convs): <= 1e-5;
  * modules and the whole pipeline against the oracle / the fixture: <= 1e-4 (contract 1e-3); labels: every pixel equal to
    the reference's except numerical ties -- pixels whose reference top-2 logit margin is below TIE_REL * max |logit|.
"""
import numpy as np
import torch
import torch.nn.functional as F

import segmif_b200
from gpu_checks import DEV, check, models, rel_err, result, rnd, _golden
from oracle import segmif_oracle as O
from segmif_b200 import ops, strict, synth
from segmif_b200.ops import ACT_NONE, ACT_PRELU, ACT_RELU

TIE_REL = 1e-5      # a label may differ only where the reference's own top-2 margin is below this fraction of max |logit|


def _act(v, act, alpha=0.25):
    if act == ACT_RELU:
        return v.clamp_min(0)
    if act == ACT_PRELU:
        return torch.where(v >= 0, v, alpha * v)
    return v


def _row_case(name, M, K, N, act=ACT_NONE, bias=True, residual=False, ld_a=None, a_coff=0, ld_dst=None, dst_coff=0, nterms=6,
              tol=2e-6, seed=0, planes_out=True):
    a = rnd(M, K, seed=seed, bf16=False)
    w = rnd(N, K, seed=seed + 1, scale=K ** -0.5, bf16=False)
    b = rnd(N, seed=seed + 2, bf16=False) if bias else None
    r = rnd(M, N, seed=seed + 3, bf16=False) if residual else None
    ref = a.double() @ w.double().t()
    if b is not None:
        ref = ref + b.double()
    ref = _act(ref, act)
    if r is not None:
        ref = ref + r.double()
    ld_a = K if ld_a is None else ld_a
    full = torch.zeros(M, ld_a)
    full[:, a_coff:a_coff + K] = a
    full[:, :a_coff] = 7.0                      # neighbours of the slice must not leak in
    full[:, a_coff + K:] = -5.0
    A = strict.split(full.to(DEV))
    alpha = torch.tensor([0.25], device=DEV)
    ld_dst = N if ld_dst is None else ld_dst
    dst = torch.full((M, ld_dst), 3.0, device=DEV)
    dp = strict.Planes(M, ld_dst, DEV) if planes_out else None
    if dp is not None:
        dp.t.zero_()
    strict.gemm(A, K, strict.pack_rows(w.to(DEV)), N, a_coff=a_coff, bias=b.to(DEV) if bias else None, act=act, alpha=alpha,
                residual=r.to(DEV) if residual else None, dst=dst, ld_dst=ld_dst, dst_coff=dst_coff, dst_planes=dp, dp_coff=dst_coff,
                nterms=nterms)
    got = dst[:, dst_coff:dst_coff + N]
    err = rel_err(got, ref)
    res = result(name, err, tol)
    untouched = bool((dst[:, :dst_coff] == 3.0).all()) and bool((dst[:, dst_coff + N:] == 3.0).all())
    if not untouched:
        res["ok"] = False
        res["note"] = "wrote outside the destination slice"
    if dp is not None:
        s = dp.t.float().sum(0)[:, dst_coff:dst_coff + N]
        perr = rel_err(s, got)
        if perr > 1e-6:
            res["ok"] = False
            res["note"] = f"split planes do not reproduce the fp32 result ({perr:.2e})"
    return res


@check
def strict_split_gemm_rows():
    rs = [
        _row_case("split_row_64x64", 300, 64, 64),
        _row_case("split_row_bias_relu_res", 1000, 320, 640, act=ACT_RELU, residual=True),
        _row_case("split_row_k96_partial_block", 257, 96, 32, act=ACT_PRELU),
        _row_case("split_row_n9_tail", 517, 256, 9, planes_out=False),
        _row_case("split_row_n160", 300, 160, 160, residual=True),
        _row_case("split_row_slices", 400, 64, 128, ld_a=224, a_coff=64, ld_dst=1024, dst_coff=256, act=ACT_RELU),
        # K = 4096 (Attention.sr of stage 1): 256 sequential fp32 accumulations in TMEM; the CPU's blocked fp32 GEMM is ~2e-6 here
        _row_case("split_row_k4096", 300, 4096, 64, bias=False, tol=1e-5),
        _row_case("split_row_big", 19200, 64, 256),
        _row_case("split_row_2planes", 300, 256, 64, nterms=3, tol=5e-5),
        _row_case("split_row_1plane_is_bf16", 300, 256, 64, nterms=1, tol=1e-2),
    ]
    # the plain-fp32 yardstick: what an fp32 GEMM on the CPU loses against fp64 on one of these cases
    a, w = rnd(1000, 320, seed=0, bf16=False), rnd(640, 320, seed=1, scale=320 ** -0.5, bf16=False)
    yard = rel_err(a @ w.t(), a.double() @ w.double().t())
    rs[1]["note"] = rs[1].get("note", "") + f" (CPU fp32 GEMM vs fp64 on the same operands: {yard:.1e})"
    return rs


def _patch_case(name, B, H, W, Cin, Cout, dil, act=ACT_RELU, ld_a=None, a_coff=0, seed=0, tol=2e-6):
    x = rnd(B, Cin, H, W, seed=seed, bf16=False)
    w = rnd(Cout, Cin, 3, 3, seed=seed + 1, scale=(9 * Cin) ** -0.5, bf16=False)
    b = rnd(Cout, seed=seed + 2, bf16=False)
    ref = _act(F.conv2d(x.double(), w.double(), b.double(), padding=dil, dilation=dil), act).permute(0, 2, 3, 1).reshape(-1, Cout)
    ld_a = Cin if ld_a is None else ld_a
    rows = torch.full((B * H * W, ld_a), 9.0)
    rows[:, a_coff:a_coff + Cin] = x.permute(0, 2, 3, 1).reshape(-1, Cin)
    A = strict.split(rows.to(DEV))
    wp = strict.pack_rows(w.permute(0, 2, 3, 1).reshape(-1, Cin).to(DEV))
    dst, dp = strict.gemm(A, Cin, wp, Cout, a_coff=a_coff, bias=b.to(DEV), act=act, alpha=torch.tensor([0.25], device=DEV),
                          patch=(B, H, W, dil), want_planes=True)
    res = result(name, rel_err(dst, ref), tol)
    perr = rel_err(dp.t.float().sum(0), dst)
    if perr > 1e-6:
        res["ok"], res["note"] = False, f"split planes do not reproduce the fp32 result ({perr:.2e})"
    return res


@check
def strict_split_gemm_patches():
    return [
        _patch_case("split_conv_dil2_64", 2, 37, 53, 64, 32, 2),
        _patch_case("split_conv_dil2_192_slice", 1, 48, 40, 192, 32, 2, ld_a=224),
        _patch_case("split_conv_dil2_96", 2, 16, 8, 96, 32, 2, ld_a=224),
        _patch_case("split_conv_dil1_128_prelu", 2, 33, 41, 128, 64, 1, act=ACT_PRELU),
        _patch_case("split_conv_dil1_64_32", 1, 64, 96, 64, 32, 1, act=ACT_PRELU),
    ]


@check
def strict_fp32_companions():
    rs = []
    # attention core (core/mix_transformer.py:107-111)
    for name, B, heads, N, Nk, D in (("attn_f32_d64", 2, 2, 300, 77, 64), ("attn_f32_d32", 1, 5, 130, 33, 32),
                                     ("attn_f32_nk1024", 1, 1, 256, 1024, 64)):
        C = heads * D
        q, kv = rnd(B, N, C, seed=1, bf16=False), rnd(B, Nk, 2 * C, seed=2, bf16=False)
        qh = q.view(B, N, heads, D).permute(0, 2, 1, 3).double()
        k = kv[..., :C].reshape(B, Nk, heads, D).permute(0, 2, 1, 3).double()
        v = kv[..., C:].reshape(B, Nk, heads, D).permute(0, 2, 1, 3).double()
        ref = (((qh @ k.transpose(-2, -1)) * D ** -0.5).softmax(-1) @ v).permute(0, 2, 1, 3).reshape(B * N, C)
        got = strict.attention_f32(q.to(DEV).view(-1, C), kv.to(DEV).view(-1, 2 * C), B, heads, N, Nk, D, D ** -0.5)
        rs.append(result(name, rel_err(got, ref), 1e-5))
    # depthwise 3x3 + GELU
    B, H, W, C = 2, 13, 17, 64
    x, w, b = rnd(B, C, H, W, seed=3, bf16=False), rnd(C, 1, 3, 3, seed=4, scale=0.3, bf16=False), rnd(C, seed=5, bf16=False)
    ref = F.gelu(F.conv2d(x.double(), w.double(), b.double(), padding=1, groups=C)).permute(0, 2, 3, 1)
    got = strict.dwconv_f32(x.permute(0, 2, 3, 1).contiguous().to(DEV), w.reshape(C, 9).t().contiguous().to(DEV), b.to(DEV), B, H, W)
    rs.append(result("dwconv_gelu_f32", rel_err(got, ref), 1e-5))
    # im2col + split: reproduces F.unfold's patches exactly (sum of the three planes == the fp32 value)
    for name, k, s, p in (("im2col_k3s2p1", 3, 2, 1), ("im2col_k4s4", 4, 4, 0)):
        B, H, W, C = 2, 16, 24, 32
        x = rnd(B, H, W, C, seed=6, bf16=False)
        P, Ho, Wo = strict.im2col_split(x.to(DEV), B, H, W, C, k, s, p)
        un = F.unfold(x.permute(0, 3, 1, 2), k, padding=p, stride=s)                        # [B, C*k*k, L], column c*k*k + tap
        ref = un.view(B, C, k * k, -1).permute(0, 3, 2, 1).reshape(B * Ho * Wo, k * k * C)
        rs.append(result(name, rel_err(P.t.float().sum(0), ref), 1e-7))
    # Gram (fp64 accumulation)
    B, HW = 2, 1000
    p = rnd(B, HW, 128, seed=7, bf16=False)
    part = torch.empty((B, 3, 64, 64), dtype=torch.float64, device=DEV)
    from segmif_b200 import _lib
    st = ops._prep(part)
    _lib.call("segmif_gram64_f64", strict._p(p.to(DEV)), 128, 64, B, HW, 1, strict._p(part), 3, st)
    pr = p[..., 64:].clamp_min(0).double()
    rs.append(result("gram64_f64", rel_err(part.sum(1), pr.transpose(1, 2) @ pr), 1e-7))
    # edge convolutions
    B, H, W = 2, 19, 23
    img, w, b = rnd(B, 3, H, W, seed=8, bf16=False), rnd(64, 1, 3, 3, seed=9, scale=0.3, bf16=False), rnd(64, seed=10, bf16=False)
    alpha = torch.tensor([0.25])
    ref = F.prelu(F.conv2d(img[:, :1].double(), w.double(), b.double(), padding=1), alpha.double()).permute(0, 2, 3, 1).reshape(-1, 64)
    got = strict.conv3x3_in1_f32(img.to(DEV), w.reshape(64, 9).t().contiguous().to(DEV), b.to(DEV), alpha.to(DEV), 64)
    rs.append(result("conv3x3_in1_f32", rel_err(got, ref), 1e-5))
    x, w, b = rnd(B, 32, H, W, seed=11, bf16=False), rnd(1, 32, 3, 3, seed=12, scale=0.1, bf16=False), rnd(1, seed=13, bf16=False)
    ref = F.prelu(F.conv2d(x.double(), w.double(), b.double(), padding=1), alpha.double())
    got = strict.conv3x3_out1_f32(x.permute(0, 2, 3, 1).reshape(-1, 32).contiguous().to(DEV), w.reshape(32, 9).t().contiguous().to(DEV),
                                  b.to(DEV), alpha.to(DEV), B, H, W, 32)
    rs.append(result("conv3x3_out1_f32", rel_err(got, ref), 1e-5))
    return rs


@check
def strict_modules():
    """DRDB, FeatureFusionModule, one MiT block and the decode head in strict mode against the oracle (fp32, CPU)."""
    seg, fus, (seg_sd, fus_sd) = models()
    rs = []
    with segmif_b200.precision("strict"), torch.no_grad():
        x = rnd(1, 64, 48, 40, seed=1, bf16=False) * 0.5
        rs.append(result("strict_drdb", rel_err(fus.DRDB1(x.to(DEV)), O.drdb(x, fus_sd, "DRDB1")), 1e-5))
        x1, x2, x3 = (rnd(2, 64, 24, 40, seed=s, bf16=False) for s in (2, 3, 4))
        r1, r2 = O.feature_fusion_module(x1, x2, x3, fus_sd, "ffm")
        g1, g2 = fus.ffm(x1.to(DEV), x2.to(DEV), x3.to(DEV))
        rs.append(result("strict_ffm", max(rel_err(g1, r1), rel_err(g2, r2)), 1e-5))
        enc_sd = O._sub(seg_sd, "denoise_net.encoder")
        inp = synth.synth_inputs(1, 64, 96, seed=0)
        feats = seg.denoise_net.encoder.forward_features(inp["mask"].to(DEV))
        ref = O.mit_forward_features(inp["mask"], enc_sd, "mit_b1")
        for i, (f, r) in enumerate(zip(feats, ref)):
            rs.append(result(f"strict_encoder_feat{i}", rel_err(f, r), 1e-4))
        o0, o1 = seg.denoise_net.encoder.forward_fusion(inp["mask"].to(DEV))
        q0, q1 = O.mit_forward_fusion(inp["mask"], enc_sd, "mit_b1")
        rs.append(result("strict_forward_fusion", max(rel_err(o0, q0), rel_err(o1, q1)), 1e-4))
        rgb = torch.rand(1, 3, 64, 96, generator=torch.Generator().manual_seed(5))
        _, _, lg = seg(rgb.to(DEV))
        rs.append(result("strict_network3_logits", rel_err(lg, O.network3_forward(rgb, seg_sd, "mit_b1")), 1e-4))
    return rs


def label_report(name, got_labels, ref_logits, size):
    """All labels must equal the reference's except numerical ties of the reference's own upsampled logits."""
    up = F.interpolate(ref_logits, size=size, mode="bilinear", align_corners=False)
    ref_lab = up.argmax(1)
    mism = got_labels.cpu() != ref_lab
    top2 = up.topk(2, dim=1).values
    margin = top2[:, 0] - top2[:, 1]
    thr = TIE_REL * float(up.abs().max())
    bad = int((margin[mism] > thr).sum()) if mism.any() else 0
    note = (f"{int(mism.sum())} of {mism.numel()} pixels differ, largest reference margin among them "
            f"{float(margin[mism].max()):.2e} (tie threshold {thr:.2e})") if mism.any() else "bit-exact"
    return result(name, float(bad), 0.0, note=note)


@check
def strict_pipeline_golden():
    """The whole unit of work in strict mode against the fixture generated by the UNMODIFIED reference (mit_b1, 64x96):
    both the materialised path (the reference's interface tensors) and the low-resolution path bench.py times."""
    from segmif_b200.pipeline import FusionSegPipeline
    seg, fus, _ = models()
    g = _golden()
    inp = synth.synth_inputs(1, 64, 96, seed=0)
    pipe = FusionSegPipeline(seg, fus)
    rs = []
    with segmif_b200.precision("strict"), torch.no_grad():
        out = pipe(inp["ir"].to(DEV), inp["vis"].to(DEV), inp["mask"].to(DEV), return_intermediates=True)
        fused_lr, labels_lr = pipe(inp["ir"].to(DEV), inp["vis"].to(DEV), inp["mask"].to(DEV))
    rs.append(result("strict_out0_vs_reference", rel_err(out["out0"].float()[:, ::4, ::4, ::4], torch.from_numpy(g["out0_s"])), 1e-4))
    rs.append(result("strict_out1_vs_reference", rel_err(out["out1"].float()[:, ::8, ::4, ::4], torch.from_numpy(g["out1_s"])), 1e-4))
    rs.append(result("strict_fused_vs_reference", rel_err(out["fused"], torch.from_numpy(g["fused"])), 1e-4))
    rs.append(result("strict_fused_lowres_path_vs_reference", rel_err(fused_lr, torch.from_numpy(g["fused"])), 1e-4))
    rs.append(result("strict_rgb_vs_reference", rel_err(out["rgb"][:, :, ::2, ::2], torch.from_numpy(g["rgb_s"])), 1e-4))
    rs.append(result("strict_logits_vs_reference", rel_err(out["logits"], torch.from_numpy(g["logits"])), 1e-4))
    ref_logits = torch.from_numpy(g["logits"])
    rs.append(label_report("strict_labels_vs_reference", out["labels"], ref_logits, (64, 96)))
    rs.append(label_report("strict_labels_lowres_path_vs_reference", labels_lr, ref_logits, (64, 96)))
    return rs


@check
def strict_pipeline_480x640_mit_b2():
    """BASELINE configs[1] geometry: one 480x640 pair through MiT-B2 + the fusion network in strict mode against the
    oracle on the same inputs and weights (the oracle takes a few seconds on the host)."""
    from segmif_b200.pipeline import FusionSegPipeline
    seg, fus, (seg_sd, fus_sd) = models("mit_b2")
    inp = synth.synth_inputs(1, 480, 640, seed=2)
    with torch.no_grad():
        ref = O.inference_pipeline(inp["ir"], inp["vis"], inp["mask"], seg_sd, fus_sd, "mit_b2")
    pipe = FusionSegPipeline(seg, fus)
    with segmif_b200.precision("strict"), torch.no_grad():
        out = pipe(inp["ir"].to(DEV), inp["vis"].to(DEV), inp["mask"].to(DEV), return_intermediates=True)
        fused_lr, labels_lr = pipe(inp["ir"].to(DEV), inp["vis"].to(DEV), inp["mask"].to(DEV))
    rs = [result("strict480_fused", rel_err(out["fused"], ref["fused"]), 1e-3),
          result("strict480_fused_lowres_path", rel_err(fused_lr, ref["fused"]), 1e-3),
          result("strict480_logits", rel_err(out["logits"], ref["logits"]), 1e-3),
          label_report("strict480_labels", out["labels"], ref["logits"], (480, 640)),
          label_report("strict480_labels_lowres_path", labels_lr, ref["logits"], (480, 640))]
    # the default bf16 path on the same pair, with its own stated gate (bf16 operands: 2^-8 per rounding)
    with torch.no_grad():
        fb, lb = pipe(inp["ir"].to(DEV), inp["vis"].to(DEV), inp["mask"].to(DEV))
    agree = float((lb.cpu() == ref["labels"]).float().mean())
    rs.append(result("bf16_480_fused", rel_err(fb, ref["fused"]), 3e-2))
    rs.append(result("bf16_480_label_disagreement", 1.0 - agree, 0.02, note=f"{agree * 100:.3f}% equal (bf16 operands; strict mode is the parity path)"))
    return rs


@check
def ablation_networks():
    """the paper's ablation networks (core/model_fusion.py:363-425, :626-1025) as compositions of the
    strict-precision kernels, against the fixture generated by the UNMODIFIED reference (oracle/make_golden_ablation.py;
    the oracle restatements are pinned to the same fixture on the CPU)."""
    import os
    from conftest import GOLDEN
    from segmif_b200.core import model_fusion as MF
    g = np.load(os.path.join(GOLDEN, "ablation.npz"))
    inp = synth.synth_inputs(1, 32, 48, seed=21)
    gen = torch.Generator().manual_seed(77)
    out1 = torch.randn(1, 64, 32, 48, generator=gen) * 0.5
    out2 = torch.randn(1, 128, 32, 48, generator=gen) * 0.5
    d = lambda t: t.to(DEV)
    rs = []
    with torch.no_grad():
        for name in ("Fusion_Network3", "Fusion_Network3_S", "Fusion_Network3_M", "Fusion_Network3_Con", "Fusion_Network3_Add",
                     "Fusion_Network3_Average"):
            net = synth.load_synthetic(getattr(MF, name)(), 3).eval().to(DEV)
            got = net(d(inp["ir"]), d(inp["vis"]), d(out1), d(out2))
            rs.append(result(f"ablation_{name}", rel_err(got, torch.from_numpy(g[name])), 1e-4))
        net = synth.load_synthetic(MF.Fusion_Network_rmseg(), 3).eval().to(DEV)
        rs.append(result("ablation_Fusion_Network_rmseg", rel_err(net(d(inp["ir"]), d(inp["vis"])), torch.from_numpy(g["Fusion_Network_rmseg"])), 1e-4))
        net = synth.load_synthetic(MF.Fusion_Network_rmseg_att(), 3).eval().to(DEV)
        o, feats = net(d(inp["ir"]), d(inp["vis"]))
        rs.append(result("ablation_Fusion_Network_rmseg_att", rel_err(o, torch.from_numpy(g["Fusion_Network_rmseg"])), 1e-4,
                         note=f"features {tuple(feats[0].shape)}"))
        gen = torch.Generator().manual_seed(5)
        t1, t2, t3 = (torch.randn(2, 150, 32, generator=gen) for _ in range(3))
        for name in ("CrossPath_M", "CrossPath_S"):
            cp = synth.load_synthetic(getattr(MF, name)(32), 3).eval()
            cp.load_state_dict({k: v * (10.0 if ".kv" in k else 1.0) for k, v in cp.state_dict().items()})
            cp = cp.to(DEV)
            r1, r2 = cp(d(t1), d(t2), d(t3))
            rs.append(result(f"ablation_{name}", max(rel_err(r1, torch.from_numpy(g[name + "_1"])), rel_err(r2, torch.from_numpy(g[name + "_2"]))), 1e-4))
        am = synth.load_synthetic(MF.AttentionModule(), 3).eval().to(DEV)
        x = torch.randn(1, 32, 20, 28, generator=gen)
        rs.append(result("ablation_AttentionModule", rel_err(am(d(x)), torch.from_numpy(g["AttentionModule"])), 1e-4))
    return rs
