"""Parity checks of the TRAINING side (backward kernels, fused AdamW, the fusion network's hand-written backward)
against torch.autograd run over the CPU oracle (oracle/segmif_oracle.py is plain differentiable torch code, so its
autograd gradients are the reference's gradients).  Same conventions as tests/gpu_checks.py.

Tolerances: fp32 loss gradients <= 1e-4 of the gradient's max |value| (SSIM 2e-3: the sigma^2 = E[x^2]-mu^2
cancellation amplifies fp32 reordering, as in the forward); tensor-core weight/data gradients on bf16-representable
operands <= 2e-3 (bf16 output rounding 2^-8 where the output is bf16); module-level bf16 backward chains 3e-2 .. 6e-2
of each gradient tensor's max |value| (bounds set from measured error with ~3x headroom)."""
import copy

import torch
import torch.nn.functional as F

from gpu_checks import CHECKS, DEV, check, rel_err, result, rnd  # noqa: F401
from oracle import segmif_oracle as O
from segmif_b200 import ops, synth
from segmif_b200.ops import ACT_PRELU, ACT_RELU


def _grad_of(fn, x, *rest):
    x = x.clone().requires_grad_(True)
    out = fn(x, *rest)
    out.backward()
    return x.grad


# ----------------------------------------------------------------------------------------- loss gradients
@check
def loss_gradients():
    res = []
    g = torch.Generator().manual_seed(3)
    for (B, H, W) in ((2, 64, 96), (1, 72, 104), (3, 40, 56)):
        a, b, c = (torch.rand(B, 1, H, W, generator=g) for _ in range(3))
        one = torch.ones((), device=DEV)
        tag = f"{B}x{H}x{W}"
        da = ops.ssim_bwd(a.to(DEV), b.to(DEV), one * 0.7, True)
        res.append(result(f"ssim_bwd_{tag}", rel_err(da, _grad_of(lambda x: 0.7 * O.ssim(x, b), a)), 2e-3))
        gv = torch.linspace(0.5, 1.5, B)
        da = ops.ssim_bwd(a.to(DEV), b.to(DEV), gv.to(DEV), False)
        res.append(result(f"ssim_bwd_per_image_{tag}", rel_err(da, _grad_of(lambda x: (gv * O.ssim(x, b, size_average=False)).sum(), a)), 2e-3))
        da = ops.laploss2_bwd(a.to(DEV), b.to(DEV), c.to(DEV), one * 1.3)
        res.append(result(f"laploss2_bwd_{tag}", rel_err(da, _grad_of(lambda x: 1.3 * O.lap_loss2(x, b, c), a)), 1e-4))
        da = ops.laploss_bwd(a.to(DEV), b.to(DEV), one)
        res.append(result(f"laploss_bwd_{tag}", rel_err(da, _grad_of(lambda x: O.lap_loss(x, b), a)), 1e-4))
        da = ops.sobel_l1_bwd(a.to(DEV), b.to(DEV), one * 0.5, one * 2.0)
        ref = _grad_of(lambda x: 0.5 * F.l1_loss(x, b) + 2.0 * F.l1_loss(O.sobelxy(x), O.sobelxy(b)), a)
        res.append(result(f"sobel_l1_bwd_{tag}", rel_err(da, ref), 1e-4))
        da = ops.mse_l1_bwd(a.to(DEV), b.to(DEV), one * 0.3, one * 1.1)
        ref = _grad_of(lambda x: 0.3 * F.mse_loss(x, b) + 1.1 * F.l1_loss(x, b), a)
        res.append(result(f"mse_l1_bwd_{tag}", rel_err(da, ref), 1e-5))
        for p in (4, 8):
            da = ops.entropy_bwd(a.to(DEV), p, one)
            res.append(result(f"entropy{p}_bwd_{tag}", rel_err(da, _grad_of(lambda x: O.entropy(x, p), a)), 2e-3))
        # accumulate flag: two terms into one plane
        acc = ops.mse_l1_bwd(a.to(DEV), b.to(DEV), one, None)
        ops.laploss_bwd(a.to(DEV), b.to(DEV), one, out=acc)
        ref = _grad_of(lambda x: F.mse_loss(x, b) + O.lap_loss(x, b), a)
        res.append(result(f"accumulate_{tag}", rel_err(acc, ref), 1e-4))
    return res


@check
def loss_modules_autograd():
    """The reference-named loss modules with gradients flowing through autograd.Function registration."""
    from segmif_b200.core.loss import Fusionloss3, Fusionloss_grad2, Fusionloss_grad3
    inp = synth.synth_inputs(2, 64, 96, seed=5)
    fused = torch.rand(2, 1, 64, 96, generator=torch.Generator().manual_seed(9))
    res = []
    for cls, ofn in ((Fusionloss3, O.fusionloss3), (Fusionloss_grad3, O.fusionloss_grad3), (Fusionloss_grad2, O.fusionloss_grad2)):
        f = fused.to(DEV).requires_grad_(True)
        loss = cls()(inp["ir"].to(DEV), inp["vis"].to(DEV), f, inp["mask"].to(DEV))
        loss.backward()
        fr = fused.clone().requires_grad_(True)
        lr = ofn(inp["ir"], inp["vis"], fr, inp["mask"])
        lr.backward()
        res.append(result(f"{cls.__name__}_value", rel_err(loss, lr), 5e-5))
        res.append(result(f"{cls.__name__}_grad", rel_err(f.grad, fr.grad), 2e-3))
    return res


# ----------------------------------------------------------------------------------------- elementwise / norm
@check
def act_and_layernorm_bwd():
    res = []
    N = 1000
    for C, ld, coff in ((32, 224, 64), (64, 64, 0), (224, 224, 0), (128, 128, 0)):
        z = rnd(N, ld, seed=C, scale=1.0)
        dy = rnd(N, ld, seed=C + 1, scale=1.0)
        for act, alpha in ((ACT_RELU, None), (ACT_PRELU, 0.25)):
            y = F.relu(z) if act == ACT_RELU else F.prelu(z, torch.tensor([alpha]))
            y16 = y.bfloat16()
            ys = y16.float()[:, coff:coff + C]
            slope = torch.where(ys > 0, torch.ones_like(ys), torch.full_like(ys, alpha or 0.0))
            ref = dy[:, coff:coff + C] * slope
            dz = torch.zeros((N, C), dtype=torch.bfloat16, device=DEV)
            dbias = torch.zeros((C,), device=DEV)
            dalpha = torch.zeros((1,), device=DEV)
            a_t = torch.tensor([alpha], device=DEV) if alpha else None
            ops.act_bwd(y16.to(DEV), ld, coff, dy.bfloat16().to(DEV), ld, coff, dz, C, 0, N, C, act, alpha=a_t, dbias=dbias,
                        dalpha=dalpha if alpha else None)
            tag = f"{'prelu' if alpha else 'relu'}_C{C}"
            res.append(result(f"act_bwd_{tag}", rel_err(dz.float(), ref), 4e-3))
            res.append(result(f"act_bwd_dbias_{tag}", rel_err(dbias, ref.sum(0)), 1e-4))
            if alpha:
                refa = (dy[:, coff:coff + C] * torch.where(ys > 0, torch.zeros_like(ys), ys / alpha)).sum()
                res.append(result(f"act_bwd_dalpha_{tag}", rel_err(dalpha[0], refa), 1e-4))
        cs = torch.zeros((C,), device=DEV)
        ops.colsum(dy.bfloat16().to(DEV), ld, coff, N, C, cs)
        res.append(result(f"colsum_C{C}", rel_err(cs, dy[:, coff:coff + C].sum(0)), 1e-5))
    a, b = rnd(N, 224, seed=1), rnd(N, 64, seed=2)
    o = torch.empty((N, 64), dtype=torch.bfloat16, device=DEV)
    ops.add_bf16(a.bfloat16().to(DEV), 224, 0, b.bfloat16().to(DEV), 64, 0, o, 64, 0, N, 64)
    res.append(result("add_bf16", rel_err(o.float(), a[:, :64] + b), 4e-3))
    # conv22's plane
    zf = rnd(2, 1, 20, 30, seed=4, bf16=False)
    of = F.prelu(zf, torch.tensor([0.25]))
    df = rnd(2, 1, 20, 30, seed=5, bf16=False)
    dz = torch.zeros((1200, 32), dtype=torch.bfloat16, device=DEV)
    db, da = torch.zeros((1,), device=DEV), torch.zeros((1,), device=DEV)
    ops.prelu_plane_bwd(of.to(DEV), df.to(DEV), torch.tensor([0.25], device=DEV), dz, 32, 0, dbias=db, dalpha=da)
    refz = df * torch.where(zf > 0, torch.ones_like(zf), torch.full_like(zf, 0.25))
    res.append(result("prelu_plane_bwd", rel_err(dz[:, 0].float(), refz.reshape(-1)), 4e-3))
    res.append(result("prelu_plane_dbias", rel_err(db[0], refz.sum()), 1e-4))
    res.append(result("prelu_plane_dalpha", rel_err(da[0], (df * torch.where(zf > 0, torch.zeros_like(zf), zf)).sum()), 1e-4))
    res.append(result("prelu_plane_untouched", float(dz[:, 1:].abs().max()), 0.0))
    # LayerNorm backward
    for C in (64, 128, 320, 512):
        x = (rnd(333, C, seed=C, scale=2.0) + 0.3).bfloat16().float()
        dy = rnd(333, C, seed=C + 7)
        gam = 1 + 0.1 * rnd(C, seed=C + 8, bf16=False)
        xr = x.clone().requires_grad_(True)
        gr = gam.clone().requires_grad_(True)
        br = torch.zeros(C, requires_grad=True)
        F.layer_norm(xr, (C,), gr, br, 1e-5).backward(dy)
        for dt, tol in ((torch.bfloat16, 6e-3), (torch.float32, 2e-5)):
            dx = torch.empty((333, C), dtype=dt, device=DEV)
            dg, dbt, dxs = (torch.zeros((C,), device=DEV) for _ in range(3))
            ops.layernorm_bwd(x.to(dt).to(DEV), dy.to(dt).to(DEV), C, 0, gam.to(DEV), 1e-5, dx, C, 0, 333, C, dgamma=dg, dbeta=dbt, dxsum=dxs)
            tag = f"C{C}_{'bf16' if dt == torch.bfloat16 else 'f32'}"
            res.append(result(f"layernorm_bwd_dx_{tag}", rel_err(dx.float(), xr.grad), tol))
            res.append(result(f"layernorm_bwd_dgamma_{tag}", rel_err(dg, gr.grad), 1e-4))
            res.append(result(f"layernorm_bwd_dbeta_{tag}", rel_err(dbt, br.grad), 1e-4))
            res.append(result(f"layernorm_bwd_dxsum_{tag}", float((dxs.cpu() - dx.float().sum(0).cpu()).abs().max()) / (float(xr.grad.abs().max()) * 333), 1e-3))
    return res


# ----------------------------------------------------------------------------------------- wgrad / dgrad / AdamW
@check
def conv_weight_and_data_gradients():
    from segmif_b200.core.fusion_train import _conv_dgrad, _dgrad_pack, _zeros_bias  # noqa: F401
    res = []
    for (B, H, W, Cin, Cout, dil) in ((2, 24, 40, 64, 32, 2), (1, 19, 37, 96, 32, 2), (2, 16, 32, 224, 32, 2), (1, 24, 24, 128, 64, 1),
                                      (1, 20, 28, 32, 32, 1), (1, 17, 23, 8, 64, 1)):
        x = rnd(B, Cin, H, W, seed=Cin + dil)
        dy = rnd(B, Cout, H, W, seed=Cin + 11)
        w = (rnd(Cout, Cin, 3, 3, seed=Cin + 3, scale=0.1)).requires_grad_(True)
        xr = x.clone().requires_grad_(True)
        F.conv2d(xr, w, None, padding=dil, dilation=dil).backward(dy)
        xp = x.permute(0, 2, 3, 1).contiguous().bfloat16().to(DEV)
        dyp = dy.permute(0, 2, 3, 1).reshape(-1, Cout).contiguous().bfloat16().to(DEV)
        grad = torch.zeros((Cout, Cin, 3, 3), device=DEV)
        ops.wgrad(dyp, Cout, 0, xp, Cin, 0, B=B, H=H, W=W, Cin=Cin, Cout=Cout, taps=9, dil=dil, grad=grad, s_co=Cin * 9, s_tap=1, s_ci=9)
        res.append(result(f"wgrad3x3_{Cin}to{Cout}_d{dil}_{H}x{W}", rel_err(grad, w.grad), 1e-3))
        ops.wgrad(dyp, Cout, 0, xp, Cin, 0, B=B, H=H, W=W, Cin=Cin, Cout=Cout, taps=9, dil=dil, grad=grad, s_co=Cin * 9, s_tap=1, s_ci=9)
        res.append(result(f"wgrad3x3_accumulates_{Cin}to{Cout}", rel_err(grad, 2 * w.grad), 1e-3))
        if Cin % 32 == 0:
            dx = torch.empty((B * H * W, Cin), dtype=torch.bfloat16, device=DEV)
            _conv_dgrad(dyp, Cout, _dgrad_pack(w).to(DEV), B, H, W, dil, dx, Cin, 0, accumulate=False)
            ref = xr.grad.permute(0, 2, 3, 1).reshape(-1, Cin)
            res.append(result(f"dgrad3x3_{Cin}to{Cout}_d{dil}", rel_err(dx.float(), ref), 5e-3))
            base = rnd(B * H * W, Cin, seed=77)
            dx2 = base.bfloat16().to(DEV)
            _conv_dgrad(dyp, Cout, _dgrad_pack(w).to(DEV), B, H, W, dil, dx2, Cin, 0, accumulate=True)
            res.append(result(f"dgrad3x3_inplace_accumulate_{Cin}to{Cout}", rel_err(dx2.float(), ref + base), 6e-3))
    for (P, K, N) in ((1000, 64, 128), (777, 224, 64), (4096, 128, 64), (50, 64, 64)):
        x = rnd(P, K, seed=K)
        dy = rnd(P, N, seed=N + 1)
        grad = torch.zeros((N, K), device=DEV)
        ops.wgrad(dy.bfloat16().to(DEV), N, 0, x.bfloat16().to(DEV), K, 0, B=1, H=1, W=1, P=P, Cin=K, Cout=N, taps=1, dil=1,
                  grad=grad, s_co=K, s_tap=1, s_ci=1)
        res.append(result(f"wgrad_linear_{P}x{K}x{N}", rel_err(grad, dy.t() @ x), 1e-3))
    # data gradient of a linear layer straight from the forward pack: dX = dY W with W [out, in] read as a [K, N] operand
    for (P, out_f, in_f) in ((1000, 128, 64), (333, 64, 256), (4096, 2048, 512), (77, 320, 1280), (500, 256, 1024)):
        W_ = rnd(out_f, in_f, seed=out_f, scale=0.1)
        dy = rnd(P, out_f, seed=in_f + 1)
        base = rnd(P, in_f, seed=3, bf16=False)
        pack = W_.bfloat16().reshape(out_f, 1, in_f).contiguous().to(DEV)
        dx = ops.linear_tc(dy.bfloat16().to(DEV), pack, None, weight_kn=True, out_dtype=torch.float32)
        res.append(result(f"dgrad_linear_weight_kn_{P}x{out_f}x{in_f}", rel_err(dx, dy @ W_), 1e-3))
        dx2 = ops.linear_tc(dy.bfloat16().to(DEV), pack, None, weight_kn=True, residual=base.to(DEV), out_dtype=torch.float32)
        res.append(result(f"dgrad_linear_weight_kn_residual_{out_f}x{in_f}", rel_err(dx2, dy @ W_ + base), 1e-3))
    # slices: dy / x inside wider pitches, co_take / ci_take
    x = rnd(1, 16, 24, 40, seed=5)
    dy = rnd(1, 16, 24, 64, seed=6)
    grad = torch.zeros((1, 8, 3, 3), device=DEV)
    ops.wgrad(dy.bfloat16().to(DEV).reshape(-1, 64), 64, 32, x.bfloat16().to(DEV), 40, 8, B=1, H=16, W=24, Cin=8, Cout=32, taps=9,
              dil=1, grad=grad, s_co=72, s_tap=1, s_ci=9, co_take=1)
    xs = x[..., 8:16].permute(0, 3, 1, 2).clone()
    ws = torch.zeros(1, 8, 3, 3, requires_grad=True)
    F.conv2d(xs, ws, None, padding=1).backward(dy[..., 32:33].permute(0, 3, 1, 2))
    res.append(result("wgrad3x3_slices_co_take", rel_err(grad, ws.grad), 1e-3))
    return res


@check
def adamw_kernel():
    n = 10007
    p0 = rnd(n, seed=1, bf16=False)
    p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([p], lr=3e-4, betas=(0.9, 0.999), weight_decay=0.01, eps=1e-8)
    pd = p0.clone().to(DEV)
    m, v = torch.zeros_like(pd), torch.zeros_like(pd)
    for t in range(1, 6):
        g = rnd(n, seed=10 + t, bf16=False)
        p.grad = g.clone() * 0.5
        opt.step()
        ops.adamw_step(pd, g.to(DEV), m, v, lr=3e-4, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.01, step=t, grad_scale=0.5)
    return [result("adamw_5_steps_param", rel_err(pd, p.detach()), 1e-6),
            result("adamw_5_steps_update", rel_err(pd.cpu() - p0, p.detach() - p0), 1e-4)]


# ----------------------------------------------------------------------------------------- module-level backward
def _param_grads_oracle(loss_fn, sd, names):
    leaves = {k: sd[k].clone().requires_grad_(True) for k in names}
    full = dict(sd)
    full.update(leaves)
    loss_fn(full).backward()
    return {k: v.grad for k, v in leaves.items()}


def _fusion_case(B=1, H=64, W=96, seed=0):
    from segmif_b200.core.model_fusion import Fusion_Network3_ac
    fus = synth.load_synthetic(Fusion_Network3_ac(), seed)
    sd = {k: v.clone() for k, v in fus.state_dict().items()}
    inp = synth.synth_inputs(B, H, W, seed=seed)
    g = torch.Generator().manual_seed(seed + 50)
    out1 = (torch.randn(B, 64, H, W, generator=g) * 0.5).bfloat16().float()
    out2 = (torch.randn(B, 128, H, W, generator=g) * 0.5).bfloat16().float()
    vis = O.rgb2ycrcb(inp["vis"])
    return fus, sd, inp, vis, out1, out2


@check
def fusion_network_backward():
    """Every parameter gradient of Fusion_Network3_ac against autograd over the oracle.
    (a) a fixed smooth cotangent d(fused) fed to both sides isolates the network's backward kernels (6e-2 of each
        tensor's max |grad|: bf16 activations and gradients through 4 DRDBs and 2 attention modules);
    (b) end to end through Fusionloss_grad3 (MSE + SSIM, train.py rounds >= 2) and Fusionloss3 (L1 + Sobel-L1, round 1).
        The L1 terms have sign() gradients, so the ~1e-2 bf16 difference of the fused image itself flips some of them:
        their bound is 0.2 and the check mainly guards against gross errors."""
    from segmif_b200.core.loss import Fusionloss3, Fusionloss_grad3
    res = []
    for B, H, W, cls, ofn, tol in ((1, 64, 96, None, None, 6e-2), (2, 40, 56, Fusionloss_grad3, O.fusionloss_grad3, 8e-2),
                                   (1, 64, 96, Fusionloss3, O.fusionloss3, 0.2)):
        fus, sd, inp, vis, out1, out2 = _fusion_case(B, H, W, seed=B)
        names = [k for k, _ in fus.named_parameters() if not k.startswith("ffm2.")]
        tag = cls.__name__ if cls else "cotangent"
        yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
        cot = (torch.sin(0.21 * xx + 0.13 * yy) + 0.3 * torch.cos(0.05 * xx * yy / W)).expand(B, 1, H, W).contiguous() / (B * H * W)
        if cls is None:
            loss_ref = lambda s: (O.fusion_network3_ac(inp["ir"], vis, out1, out2, s) * cot).sum()
        else:
            loss_ref = lambda s: ofn(inp["ir"], vis, O.fusion_network3_ac(inp["ir"], vis, out1, out2, s), inp["mask"])
        ref = _param_grads_oracle(loss_ref, sd, names)
        net = copy.deepcopy(fus).to(DEV).train()
        fused = net(inp["ir"].to(DEV), vis.to(DEV), out1.to(DEV), out2.to(DEV))
        with torch.no_grad():
            fref = O.fusion_network3_ac(inp["ir"], vis, out1, out2, sd)
        res.append(result(f"train_forward_fused_{tag}", rel_err(fused, fref), 3e-2))
        if cls is None:
            fused.backward(cot.to(DEV))
        else:
            cls()(inp["ir"].to(DEV), vis.to(DEV), fused, inp["mask"].to(DEV)).backward()
        worst, worst_name = 0.0, ""
        got = dict(net.named_parameters())
        groups = {}
        for k in names:
            if got[k].grad is None:
                res.append(result(f"grad_missing_{k}", float("nan"), 0.0))
                continue
            e = rel_err(got[k].grad, ref[k])
            grp = k.split(".")[0] + ("." + k.split(".")[2] if k.startswith("ffm.") else "")
            groups[grp] = max(groups.get(grp, 0.0), e)
            if e > worst:
                worst, worst_name = e, k
        for grp, e in sorted(groups.items()):
            res.append(result(f"grad_{tag}_{grp}", e, tol))
        res.append(result(f"grad_worst_{tag}", worst, tol, note=worst_name))
        res.append(result(f"ffm2_untouched_{tag}", 0.0 if all(p.grad is None for k, p in got.items() if k.startswith("ffm2.")) else 1.0, 0.0))
    return res


@check
def fusion_trainer_steps():
    """Three optimisation steps of FusionTrainer (flat gradient buffer + fused AdamW) against torch.optim.AdamW over
    the oracle's autograd gradients: loss trajectory and parameter updates."""
    from segmif_b200.core.loss import Fusionloss3
    from segmif_b200.ddp import FusionTrainer
    fus, sd, inp, vis, out1, out2 = _fusion_case(1, 48, 64, seed=3)
    names = [k for k, _ in fus.named_parameters() if not k.startswith("ffm2.")]
    leaves = {k: sd[k].clone().requires_grad_(True) for k in names}
    opt = torch.optim.AdamW(list(leaves.values()), lr=3e-4, betas=(0.9, 0.999), weight_decay=0.01, eps=1e-8)
    ref_losses = []
    for it in range(3):
        full = dict(sd)
        full.update(leaves)
        opt.param_groups[0]["lr"] = 3e-4 * (1e-6 if it == 0 else (1 - it / 100.0))     # utils/optimizer.py:16-27
        opt.zero_grad()
        l = O.fusionloss3(inp["ir"], vis, O.fusion_network3_ac(inp["ir"], vis, out1, out2, full), inp["mask"])
        l.backward()
        opt.step()
        ref_losses.append(float(l.detach()))
    net = copy.deepcopy(fus).to(DEV).train()
    tr = FusionTrainer(net, Fusionloss3(), lr=3e-4, weight_decay=0.01, betas=(0.9, 0.999), warmup_iter=3e-5, max_iter=100,
                       warmup_ratio=1e-6, power=1.0)
    got_losses = []
    args = [t.to(DEV) for t in (inp["ir"], vis, out1, out2, inp["mask"])]
    for it in range(3):
        l, _ = tr.step(*args)
        got_losses.append(float(l))
    res = [result("trainer_loss_trajectory", max(abs(a - b) / abs(b) for a, b in zip(got_losses, ref_losses)), 2e-2,
                  note=f"{got_losses} vs {ref_losses}")]
    # parameter updates: sign-SGD-like first steps of Adam make the update direction the sensitive quantity
    num = den = 0.0
    agree = tot = 0
    for k, p in net.named_parameters():
        if k.startswith("ffm2."):
            continue
        du, dr = (p.detach().cpu() - sd[k]), (leaves[k].detach() - sd[k])
        num += float((du - dr).pow(2).sum())
        den += float(dr.pow(2).sum())
        agree += int((torch.sign(du) == torch.sign(dr)).sum())
        tot += du.numel()
    res.append(result("trainer_update_rel_l2", (num / den) ** 0.5, 0.35, note=f"sign agreement {agree / tot:.3f}"))
    res.append(result("trainer_ffm2_frozen", max(float((p.detach().cpu() - sd[k]).abs().max()) for k, p in net.named_parameters() if k.startswith("ffm2.")), 0.0))
    res.append(result("trainer_lr_schedule", abs(tr.opt.lr - 3e-4 * (1 - 2 / 100.0)) / 3e-4, 1e-6))
    return res


# ----------------------------------------------------------------------------------------- segmentation-net training kernels
@check
def attention_backward():
    res = []
    for (B, heads, N, Nk) in ((2, 1, 384, 6), (1, 2, 200, 200), (2, 5, 150, 70), (1, 8, 300, 300), (1, 1, 1000, 130)):
        D, C = 64, heads * 64
        scale = D ** -0.5
        q = rnd(B, N, C, seed=N, scale=1.0)
        kv = rnd(B, Nk, 2 * C, seed=N + 1, scale=1.0)
        do = rnd(B, N, C, seed=N + 2, scale=1.0)
        qr, kvr = q.clone().requires_grad_(True), kv.clone().requires_grad_(True)
        qh = qr.reshape(B, N, heads, D).permute(0, 2, 1, 3)
        kh = kvr[..., :C].reshape(B, Nk, heads, D).permute(0, 2, 1, 3)
        vh = kvr[..., C:].reshape(B, Nk, heads, D).permute(0, 2, 1, 3)
        ref = ((qh @ kh.transpose(-2, -1)) * scale).softmax(-1) @ vh
        ref = ref.transpose(1, 2).reshape(B, N, C)
        ref.backward(do)
        qd, kvd = q.bfloat16().to(DEV).reshape(B * N, C), kv.bfloat16().to(DEV).reshape(B * Nk, 2 * C)
        out, lse = ops.sr_attention_train(qd, kvd, B, heads, N, Nk, D, scale)
        tag = f"B{B}h{heads}N{N}Nk{Nk}"
        res.append(result(f"attn_train_fwd_{tag}", rel_err(out.float().reshape(B, N, C), ref), 1e-2))
        dq, dkv = ops.sr_attention_bwd(qd, kvd, out, do.bfloat16().to(DEV).reshape(B * N, C), lse, B, heads, N, Nk, D, scale)
        res.append(result(f"attn_bwd_dq_{tag}", rel_err(dq.float().reshape(B, N, C), qr.grad), 2e-2))
        res.append(result(f"attn_bwd_dkv_{tag}", rel_err(dkv.reshape(B, Nk, 2 * C), kvr.grad), 2e-2))
    return res


@check
def seg_head_training_kernels():
    res = []
    g = torch.Generator().manual_seed(11)
    # upsample + CE backward
    for (B, h, w, H, W) in ((2, 16, 24, 64, 96), (1, 15, 20, 60, 77), (1, 12, 12, 12, 12)):
        lg = torch.randn(B, 9, h, w, generator=g) * 2
        lab = torch.randint(0, 9, (B, H, W), generator=g)
        lab[torch.rand(B, H, W, generator=g) < 0.1] = 255
        lr = lg.clone().requires_grad_(True)
        O.seg_cross_entropy(lr, lab).backward()
        lgp = lg.permute(0, 2, 3, 1).contiguous().to(DEV)
        loss, cnt = ops.upsample_ce(lgp, B, h, w, 9, lab.to(DEV), 255, return_count=True)
        dl = ops.upsample_ce_bwd(lgp, B, h, w, 9, lab.to(DEV), torch.ones((), device=DEV), cnt, 255)
        res.append(result(f"upsample_ce_bwd_{h}x{w}to{H}x{W}", rel_err(dl, lr.grad.permute(0, 2, 3, 1)), 1e-4))
    # bilinear adjoint
    for (B, h, w, H, W, C, ld, coff) in ((2, 8, 12, 16, 24, 64, 128, 32), (1, 5, 7, 40, 56, 256, 1024, 512), (1, 10, 10, 10, 10, 32, 32, 0)):
        dd = rnd(B, H, W, ld, seed=h)
        xr = torch.zeros(B, C, h, w, requires_grad=True)
        F.interpolate(xr, size=(H, W), mode="bilinear", align_corners=False).backward(dd[..., coff:coff + C].permute(0, 3, 1, 2))
        ds = ops.bilinear_nhwc_bwd(dd.bfloat16().to(DEV), ld, coff, B, H, W, h, w, C)
        res.append(result(f"bilinear_bwd_{h}x{w}to{H}x{W}_C{C}", rel_err(ds.float(), xr.grad.permute(0, 2, 3, 1)), 5e-3))
    # train-mode BatchNorm + ReLU
    for (rows, C) in ((1000, 256), (333, 64)):
        z = (rnd(rows, C, seed=C, scale=1.5) + 0.2).bfloat16().float()
        gam, bet = 1 + 0.1 * rnd(C, seed=1, bf16=False), 0.1 * rnd(C, seed=2, bf16=False)
        rm, rv = 0.1 * rnd(C, seed=3, bf16=False), 1 + 0.1 * rnd(C, seed=4, bf16=False).abs()
        dy = rnd(rows, C, seed=5)
        zr, gr, br = z.clone().requires_grad_(True), gam.clone().requires_grad_(True), bet.clone().requires_grad_(True)
        rm_ref, rv_ref = rm.clone(), rv.clone()
        yr = F.relu(F.batch_norm(zr, rm_ref, rv_ref, gr, br, True, 0.1, 1e-5))
        yr.backward(dy)
        rmd, rvd = rm.clone().to(DEV), rv.clone().to(DEV)
        y, stats = ops.bn_train_fwd(z.bfloat16().to(DEV), gam.to(DEV), bet.to(DEV), 1e-5, 0.1, rmd, rvd)
        res.append(result(f"bn_train_fwd_C{C}", rel_err(y.float(), yr), 5e-3))
        res.append(result(f"bn_running_mean_C{C}", rel_err(rmd, rm_ref), 1e-5))
        res.append(result(f"bn_running_var_C{C}", rel_err(rvd, rv_ref), 1e-5))
        dg, db = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
        dz = ops.bn_train_bwd(z.bfloat16().to(DEV), y, dy.bfloat16().to(DEV), stats, gam.to(DEV), dg, db)
        res.append(result(f"bn_train_bwd_dz_C{C}", rel_err(dz.float(), zr.grad), 8e-3))
        res.append(result(f"bn_train_bwd_dgamma_C{C}", rel_err(dg, gr.grad), 5e-3))
        res.append(result(f"bn_train_bwd_dbeta_C{C}", rel_err(db, br.grad), 5e-3))
    x = rnd(2, 50, 64, seed=8)
    sc = torch.tensor([[0.0, 1 / 0.9] * 32, [1 / 0.9, 0.0] * 32])
    y = ops.channel_scale(x.bfloat16().to(DEV), sc.to(DEV), 2, 50, 64)
    res.append(result("channel_scale", rel_err(y.float(), x * sc[:, None, :]), 4e-3))
    return res


@check
def encoder_training_kernels():
    res = []
    # depthwise conv (+ flipped) and dwconv + GELU backward
    for (B, H, W, C) in ((2, 12, 20, 256), (1, 7, 9, 2048), (1, 16, 16, 64), (2, 33, 47, 128), (1, 9, 11, 40)):   # C % 64 != 0: strip-walking fallback
        x = rnd(B, C, H, W, seed=C)
        w = rnd(C, 1, 3, 3, seed=C + 1, scale=0.3, bf16=False)
        b = 0.1 * rnd(C, seed=C + 2, bf16=False)
        dy = rnd(B, C, H, W, seed=C + 3)
        xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
        zr = F.conv2d(xr, wr, br, padding=1, groups=C)
        zr.retain_grad()
        F.gelu(zr).backward(dy)
        pm = lambda t: t.permute(0, 2, 3, 1).contiguous()
        w9c = w.reshape(C, 9).t().contiguous().to(DEV)
        xd, dyd = pm(x).bfloat16().to(DEV), pm(dy).bfloat16().to(DEV)
        z = ops.dwconv3x3(xd, w9c, b.to(DEV), B, H, W)
        res.append(result(f"dwconv3x3_C{C}", rel_err(z.float(), pm(zr.detach())), 5e-3))
        dw, db = torch.zeros(9, C, device=DEV), torch.zeros(C, device=DEV)
        dz = ops.dwconv3x3_gelu_bwd(xd, w9c, b.to(DEV), dyd, B, H, W, dw, db)
        res.append(result(f"dwconv_gelu_bwd_dz_C{C}", rel_err(dz.float(), pm(zr.grad)), 5e-3))
        res.append(result(f"dwconv_gelu_bwd_dw_C{C}", rel_err(dw, wr.grad.reshape(C, 9).t()), 1e-3))
        res.append(result(f"dwconv_gelu_bwd_db_C{C}", rel_err(db, br.grad), 1e-3))
        dx = ops.dwconv3x3(dz, w9c, None, B, H, W, flip=True)
        res.append(result(f"dwconv_dgrad_C{C}", rel_err(dx.float(), pm(xr.grad)), 1e-2))
    # col2im against F.fold
    for (B, H, W, C, k, s, p) in ((2, 12, 16, 64, 3, 2, 1), (1, 16, 24, 3, 7, 4, 3), (1, 8, 8, 32, 2, 2, 0), (1, 9, 11, 16, 3, 2, 1)):
        Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
        ld = (C * k * k + 7) // 8 * 8
        dcol = rnd(B * Ho * Wo, ld, seed=k)
        cols = dcol[:, :C * k * k].reshape(B, Ho * Wo, C * k * k).transpose(1, 2)
        ref = F.fold(cols, (H, W), kernel_size=k, stride=s, padding=p)
        for dt in (torch.float32, torch.bfloat16):
            dx = ops.col2im(dcol.bfloat16().to(DEV), B, H, W, C, k, s, p, out_dtype=dt)
            res.append(result(f"col2im_k{k}s{s}_C{C}_{'f32' if dt == torch.float32 else 'bf16'}", rel_err(dx.float(), ref.permute(0, 2, 3, 1)),
                              1e-6 if dt == torch.float32 else 4e-3))
    # small pieces
    img = torch.rand(2, 3, 10, 14, generator=torch.Generator().manual_seed(2))
    sc, sh = torch.tensor([4.3, 4.5, 4.4]), torch.tensor([-2.1, -2.0, -1.8])
    y = ops.channel_affine_nchw(img.to(DEV), sc.to(DEV), sh.to(DEV))
    res.append(result("channel_affine_nchw", rel_err(y, img * sc.view(1, 3, 1, 1) + sh.view(1, 3, 1, 1)), 1e-6))
    fused = torch.rand(2, 1, 10, 14, generator=torch.Generator().manual_seed(3)) * 1.4 - 0.2
    ycc = O.rgb2ycrcb(img)
    fr = fused.clone().requires_grad_(True)
    rgb = O.recompose_rgb(fr, ycc)
    drgb = torch.randn(2, 3, 10, 14, generator=torch.Generator().manual_seed(4))
    rgb.backward(drgb)
    d = ops.recompose_rgb_bwd(rgb.detach().to(DEV), drgb.to(DEV), True)
    res.append(result("recompose_rgb_bwd", rel_err(d, fr.grad), 1e-6))
    xf = rnd(300, 64, seed=9, bf16=False)
    res.append(result("cast_f32_bf16", rel_err(ops.cast(xf.to(DEV), torch.bfloat16).float(), xf.bfloat16().float()), 0.0))
    yb = rnd(300, 64, seed=10)
    s2 = torch.tensor([0.0, 1.0 / 0.9, 1.0])
    out = ops.scale_add_rows(xf.to(DEV), yb.bfloat16().to(DEV), s2.to(DEV), 100)
    res.append(result("scale_add_rows", rel_err(out, xf + s2.repeat_interleave(100)[:, None] * yb), 1e-6))
    sc16 = ops.scale_cast_rows(xf.to(DEV), s2.to(DEV), 100)
    res.append(result("scale_cast_rows", rel_err(sc16.float(), (s2.repeat_interleave(100)[:, None] * xf).bfloat16().float()), 0.0))
    # DropPath factor in the GEMM epilogue: y = (x W^T + b) * s[row // rps] + residual
    Wl, bl = rnd(128, 64, seed=21, scale=0.2), 0.1 * rnd(128, seed=22, bf16=False)
    xin, resid = rnd(300, 64, seed=23), rnd(300, 128, seed=24, bf16=False)
    yl = ops.linear_tc(xin.bfloat16().to(DEV), Wl.bfloat16().reshape(128, 1, 64).contiguous().to(DEV), bl.to(DEV), residual=resid.to(DEV),
                       out_dtype=torch.float32, row_scale=s2.to(DEV), rows_per_scale=100)
    res.append(result("linear_row_scale_residual", rel_err(yl, (xin @ Wl.t() + bl) * s2.repeat_interleave(100)[:, None] + resid), 1e-3))
    # LayerNorm backward accumulating into the residual-stream gradient
    x = rnd(100, 128, seed=12, bf16=False)
    dy = rnd(100, 128, seed=13)
    base = rnd(100, 128, seed=14, bf16=False)
    xr = x.clone().requires_grad_(True)
    F.layer_norm(xr, (128,), torch.ones(128), torch.zeros(128), 1e-6).backward(dy)
    dx = base.clone().to(DEV)
    ops.layernorm_bwd(x.to(DEV), dy.bfloat16().to(DEV), 128, 0, torch.ones(128, device=DEV), 1e-6, dx, 128, 0, 100, 128, accumulate=True)
    res.append(result("layernorm_bwd_accumulate", rel_err(dx, base + xr.grad), 1e-5))
    return res


def _seg_case():
    from segmif_b200.core.model_fusion import Network3
    B, H, W = 2, 64, 96
    net0 = synth.load_synthetic(Network3("mit_b1", 9, 256, None), 0)
    sd = {k: v.clone() for k, v in net0.state_dict().items()}
    names = [k for k, _ in net0.named_parameters() if not k.endswith("classifier.weight")]
    gen = torch.Generator().manual_seed(21)
    x = torch.rand(B, 3, H, W, generator=gen)
    drop = ((torch.rand(B, 256, generator=gen) >= 0.1).float() / 0.9)
    labels = synth.synth_inputs(B, H, W, seed=4)["labels"]
    h, w = H // 4, W // 4
    cot = torch.randn(B, 9, h, w, generator=gen) / (B * h * w)
    dps = [((torch.rand(B, generator=gen) < 0.8).float() / 0.8, (torch.rand(B, generator=gen) < 0.8).float() / 0.8) for _ in range(8)]
    return net0, sd, names, x, drop, labels, cot, dps


def _floored_err(got, ref, floor):
    """max |got - ref| over max(max |ref|, floor): gradients that are analytically zero (a per-channel constant in front
    of the train-mode BatchNorm: decoder.linear_c*.proj.bias, encoder.norm4.bias -- the reference's own values are 1e-9
    rounding noise) are judged on the scale of the network's largest gradient instead of their own."""
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    return float((got - ref).abs().max()) / max(float(ref.abs().max()), floor)


@check
def seg_modules_isolated():
    """Every MiT block and the decode head of Network3 run ALONE through the training tape, fed with the oracle's exact
    input and the oracle's exact output gradient, so an error is attributable to one module instead of being noise
    accumulated over the chain.  Bounds: one block 5e-2 of each tensor's max |grad| (measured <= 2e-2); the head 0.12:
    a CPU emulation of nothing but bf16 rounding of the GEMM operands (train-mode BatchNorm + ReLU in front of a random
    cotangent) already gives 6.7e-2 on linear_fuse.conv.weight and 0.12 on the stage-1 feature gradient."""
    from segmif_b200.core import seg_train as T
    net0, sd, names, x, drop, labels, cot, dps = _seg_case()
    B = x.shape[0]
    res = []
    leaves = {k: sd[k].clone().requires_grad_(True) for k in names}
    full = dict(sd)
    full.update(leaves)
    esd = O._sub(full, "denoise_net.encoder")
    mean = torch.tensor(O.IMAGENET_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(O.IMAGENET_STD).view(1, 3, 1, 1)
    cur = (x * 255 - mean) / std
    cfg = O.MIT_CONFIGS["mit_b1"]
    toks, feats, bi = [], [], 0
    for s in range(4):
        patch, stride = (7, 4) if s == 0 else (3, 2)
        tok, hh, ww = O.overlap_patch_embed(cur, esd, f"patch_embed{s + 1}", patch, stride)
        tok.retain_grad()
        lst = [tok]
        for i in range(cfg["depths"][s]):
            tok = O.mit_block(tok, hh, ww, esd, f"block{s + 1}.{i}", O.MIT_HEADS[s], O.MIT_SR[s], dps[bi])
            tok.retain_grad()
            lst.append(tok)
            bi += 1
        toks.append((lst, hh, ww))
        cur = O.layer_norm(tok, esd, f"norm{s + 1}", O.BLOCK_LN_EPS).reshape(B, hh, ww, -1).permute(0, 3, 1, 2).contiguous()
        cur.retain_grad()
        feats.append(cur)
    lg_ref = O.segformer_head(feats, O._sub(full, "denoise_net.decoder"), train_bn=True, dropout_scale=drop)
    (lg_ref * cot).sum().backward()
    gmax = max(float(leaves[n].grad.abs().max()) for n in names)

    net = copy.deepcopy(net0).to(DEV).train()
    enc, head = net.denoise_net.encoder, net.denoise_net.decoder
    lookup = dict(net.named_parameters())
    newg = lambda: T._Grads((n, torch.zeros(lookup[n].shape, dtype=torch.float32, device=DEV)) for n in names)

    def worst(g, prefix):
        w, wn = 0.0, ""
        for n in names:
            if n.startswith(prefix):
                e = _floored_err(g[n], leaves[n].grad, 1e-2 * gmax)
                if e > w:
                    w, wn = e, n
        return w, wn

    bi = 0
    for s in range(4):
        lst, hh, ww = toks[s]
        N = hh * ww
        for i, blk in enumerate(getattr(enc, f"block{s + 1}")):
            xin = lst[i].detach().reshape(B * N, -1).contiguous().to(DEV)
            x3, sv = T._block_forward(blk, xin, B, N, hh, ww, (dps[bi][0].to(DEV), dps[bi][1].to(DEV)))
            res.append(result(f"seg_block{s + 1}.{i}_alone_fwd", rel_err(x3.reshape(B, N, -1), lst[i + 1]), 5e-3))
            dx = lst[i + 1].grad.reshape(B * N, -1).contiguous().to(DEV).clone()
            g = newg()
            pre = f"denoise_net.encoder.block{s + 1}.{i}."
            T._block_backward(blk, sv, dx, B, N, g, pre)
            res.append(result(f"seg_block{s + 1}.{i}_alone_dx", rel_err(dx.reshape(B, N, -1), lst[i].grad), 1e-2))
            w, wn = worst(g, pre)
            res.append(result(f"seg_block{s + 1}.{i}_alone_param_grads", w, 5e-2, note=wn))
            bi += 1
    stages = [(f.detach().permute(0, 2, 3, 1).reshape(-1, f.shape[1]).contiguous().bfloat16().to(DEV), f.shape[2], f.shape[3]) for f in feats]
    logits, htape = T.head_forward(head, stages, B, {"dropout2d": drop.to(DEV)})
    res.append(result("seg_head_alone_logits", rel_err(logits.permute(0, 3, 1, 2), lg_ref), 2e-2))
    g = newg()
    douts = T.head_backward(head, htape, cot.permute(0, 2, 3, 1).contiguous().to(DEV), B, g, "denoise_net.decoder.")
    f4 = feats[3]           # the only stage output whose oracle gradient has no share from a later patch embedding
    res.append(result("seg_head_alone_dfeat4", rel_err(douts[3].float().reshape(B, f4.shape[2], f4.shape[3], -1).permute(0, 3, 1, 2), f4.grad), 0.12))
    w, wn = worst(g, "denoise_net.decoder.")
    res.append(result("seg_head_alone_param_grads", w, 0.12, note=wn))
    return res


@check
def seg_network_backward():
    """Every parameter gradient of Network3 (MiT-B1 encoder + SegFormer head, train mode: batch-statistics BatchNorm,
    injected Dropout2d / DropPath masks) and the gradient w.r.t. the input image against autograd over the oracle:
    (a) fixed cotangent on the logits, DropPath masks all ones; (b) the same with DropPath masks; (c) end to end through
    Network3._loss (upsample + CrossEntropy(ignore_index=255)).  Bound 0.2 of each tensor's max |grad| over the whole
    chain (8 blocks, bf16 activations AND gradients): a CPU emulation that only rounds the GEMM operands of the oracle
    to bf16 already differs from the fp32 oracle by 6e-2 (median tensor) to 0.15 (worst tensor) on this case, while
    every module fed alone with exact inputs is within 2e-2 (seg_modules_isolated) -- the chain is noise-, not
    bug-limited."""
    from segmif_b200.core.seg_train import logits_with_grad
    res = []
    net0, sd, names, x, drop, labels, cot, dps = _seg_case()
    B, H, W = x.shape[0], x.shape[2], x.shape[3]
    ones = [(torch.ones(B), torch.ones(B)) for _ in range(8)]
    for tag, droppath, use_ce in (("cotangent", None, False), ("droppath", dps, False), ("ce", None, True)):
        leaves = {k: sd[k].clone().requires_grad_(True) for k in names}
        full = dict(sd)
        full.update(leaves)
        xr = x.clone().requires_grad_(True)
        lg_ref = O.network3_forward(xr, full, "mit_b1", train_bn=True, dropout_scale=drop, droppath=droppath)
        (O.seg_cross_entropy(lg_ref, labels) if use_ce else (lg_ref * cot).sum()).backward()
        gmax = max(float(leaves[n].grad.abs().max()) for n in names)
        net = copy.deepcopy(net0).to(DEV).train()
        masks = {"dropout2d": drop.to(DEV)}
        # the modules are in train mode: without injected masks every block would draw its own DropPath mask
        for i, (a, b) in enumerate(droppath if droppath is not None else ones):
            masks[("droppath", i)] = (a.to(DEV), b.to(DEV))
        xd = x.to(DEV).requires_grad_(True)
        lg = logits_with_grad(net, xd, masks)
        res.append(result(f"seg_train_logits_{tag}", rel_err(lg.permute(0, 3, 1, 2), lg_ref), 3e-2))
        if use_ce:
            from segmif_b200.core.seg_train import CeFn
            CeFn.apply(lg, labels.to(DEV), 255).backward()
        else:
            lg.backward(cot.permute(0, 2, 3, 1).contiguous().to(DEV))
        got = dict(net.named_parameters())
        groups, worst, worst_name = {}, 0.0, ""
        for k in names:
            if got[k].grad is None:
                res.append(result(f"seg_grad_missing_{k}", float("nan"), 0.0))
                continue
            e = _floored_err(got[k].grad, leaves[k].grad, 1e-2 * gmax)
            parts = k.split(".")
            grp = ".".join(parts[1:3]) if parts[1] == "decoder" else parts[2]
            groups[grp] = max(groups.get(grp, 0.0), e)
            if e > worst:
                worst, worst_name = e, k
        for grp, e in sorted(groups.items()):
            res.append(result(f"seg_grad_{tag}_{grp}", e, 0.2))
        res.append(result(f"seg_grad_worst_{tag}", worst, 0.2, note=worst_name))
        res.append(result(f"seg_grad_input_{tag}", rel_err(xd.grad, xr.grad), 0.2))
        res.append(result(f"seg_classifier_untouched_{tag}", 0.0 if got["denoise_net.classifier.weight"].grad is None else 1.0, 0.0))
    # running statistics follow nn.BatchNorm2d's update rule
    bn = net.denoise_net.decoder.linear_fuse.bn
    res.append(result("seg_bn_batches_tracked", abs(int(bn.num_batches_tracked) - 1), 0.0))
    return res


@check
def seg_trainer_steps():
    """Three optimisation steps of SegTrainer (train.py:207-226: three AdamW param groups over the flat buffer) against
    torch.optim.AdamW with the same groups over the oracle's autograd gradients.  DropPath / Dropout2d masks are pinned
    to the identity on both sides (drop rates set to 0), BatchNorm uses batch statistics."""
    from segmif_b200.ddp import SegTrainer
    net0, sd, names, x, drop, labels, cot, dps = _seg_case()
    res = []
    gid = lambda k: 2 if ".decoder." in k else (1 if "norm" in k.split("encoder.", 1)[-1] else 0)
    leaves = {k: sd[k].clone().requires_grad_(True) for k in names}
    lr, wd = 6e-5, 0.01
    groups = [dict(params=[leaves[k] for k in names if gid(k) == g], lr=lr * (10 if g == 2 else 1), weight_decay=0.0 if g == 1 else wd)
              for g in range(3)]
    opt = torch.optim.AdamW(groups, lr=lr, betas=(0.9, 0.999), weight_decay=wd, eps=1e-8)
    base = [g["lr"] for g in opt.param_groups]
    ref_losses = []
    for it in range(3):
        full = dict(sd)
        full.update(leaves)
        mult = 1 - (1 - it / 2) * (1 - 1e-6) if it < 2 else (1 - it / 100.0)          # warmup_iter = 2 (utils/optimizer.py:49-60)
        for g, b in zip(opt.param_groups, base):
            g["lr"] = b * mult
        opt.zero_grad()
        l = O.seg_cross_entropy(O.network3_forward(x, full, "mit_b1", train_bn=True), labels)
        l.backward()
        opt.step()
        ref_losses.append(float(l.detach()))
    net = copy.deepcopy(net0).to(DEV).train()
    net.denoise_net.decoder.dropout.p = 0.0
    for m in net.modules():
        if hasattr(m, "drop_prob"):
            m.drop_prob = 0.0
    tr = SegTrainer(net, lr=lr, weight_decay=wd, betas=(0.9, 0.999), warmup_iter=2, max_iter=100, warmup_ratio=1e-6, power=1.0)
    got = [float(tr.step(x.to(DEV), labels.to(DEV))) for _ in range(3)]
    res.append(result("seg_trainer_loss_trajectory", max(abs(a - b) / abs(b) for a, b in zip(got, ref_losses)), 2e-2, note=f"{got} vs {ref_losses}"))
    for g in range(3):
        num = den = 0.0
        for k, p in net.named_parameters():
            if k in leaves and gid(k) == g:
                du, dr = (p.detach().cpu() - sd[k]), (leaves[k].detach() - sd[k])
                num += float((du - dr).pow(2).sum())
                den += float(dr.pow(2).sum())
        res.append(result(f"seg_trainer_update_rel_l2_group{g}", (num / den) ** 0.5, 0.35))
    # the same three steps with forward + backward replayed from a CUDA graph (captured after one eager step)
    net2 = copy.deepcopy(net0).to(DEV).train()
    net2.denoise_net.decoder.dropout.p = 0.0
    for m in net2.modules():
        if hasattr(m, "drop_prob"):
            m.drop_prob = 0.0
    tr2 = SegTrainer(net2, lr=lr, weight_decay=wd, betas=(0.9, 0.999), warmup_iter=2, max_iter=100, warmup_ratio=1e-6, power=1.0)
    xd, ld = x.to(DEV), labels.to(DEV)
    got2 = [float(tr2.step(xd, ld))]
    tr2.capture(xd, ld)
    got2 += [float(tr2.step(xd, ld)) for _ in range(2)]
    res.append(result("seg_trainer_graph_equals_eager_losses", max(abs(a - b) / abs(b) for a, b in zip(got2, got)), 1e-3, note=f"{got2}"))
    res.append(result("seg_trainer_graph_equals_eager_params", rel_err(tr2.flat.param, tr.flat.param), 1e-2))
    cw = net.denoise_net.classifier.weight
    res.append(result("seg_trainer_classifier_frozen", float((cw.detach().cpu() - sd["denoise_net.classifier.weight"]).abs().max()), 0.0))
    res.append(result("seg_trainer_lr_groups", max(abs(tr.opt.groups[i]["lr"] - b * (1 - 2 / 100.0)) / b for i, b in enumerate(base)), 1e-6))
    return res


@check
def fusion_composite_step():
    """train_fusion rounds >= 2 (train.py:361-380): Fusionloss_grad3 on the fused plane + CE through YCrCb2RGB and the
    frozen segmentation network.  Checks d(loss)/d(fused) of the CE branch (recompose + Network3._loss with frozen
    weights: no weight gradients are formed) against the oracle, and one composite FusionTrainer step's loss value."""
    from segmif_b200.autograd import recompose_rgb
    from segmif_b200.core.model_fusion import Network3
    res = []
    B, H, W = 2, 64, 96
    seg0 = synth.load_synthetic(Network3("mit_b1", 9, 256, None), 0)
    ssd = {k: v.clone() for k, v in seg0.state_dict().items()}
    inp = synth.synth_inputs(B, H, W, seed=5)
    fused = (inp["ir"] * 0.6 + 0.3 * inp["vis"][:, :1]).clone()
    ycc = O.rgb2ycrcb(inp["vis"])
    fr = fused.clone().requires_grad_(True)
    ce_ref = O.seg_cross_entropy(O.network3_forward(O.recompose_rgb(fr, ycc, clamp=False), ssd, "mit_b1", train_bn=True), inp["labels"])
    ce_ref.backward()
    seg = copy.deepcopy(seg0).to(DEV).train()
    seg.denoise_net.decoder.dropout.p = 0.0
    for m in seg.modules():
        if hasattr(m, "drop_prob"):
            m.drop_prob = 0.0
    for p in seg.parameters():
        p.requires_grad_(False)
    fd = fused.to(DEV).requires_grad_(True)
    ce = seg._loss(recompose_rgb(fd, inp["vis"].to(DEV), False), inp["labels"].to(DEV), torch.nn.CrossEntropyLoss(ignore_index=255))
    ce.backward()
    res.append(result("composite_ce_value", abs(float(ce) - float(ce_ref)) / abs(float(ce_ref)), 2e-2))
    res.append(result("composite_ce_dfused", rel_err(fd.grad, fr.grad), 0.2))
    res.append(result("composite_seg_params_no_grad", 0.0 if all(p.grad is None for p in seg.parameters()) else 1.0, 0.0))
    # one composite trainer step: loss = 0.4/iter_ * Fusionloss_grad3 + 0.8 * CE (train.py:379-380, n_iter <= 10)
    from segmif_b200.core.loss import Fusionloss_grad3
    from segmif_b200.ddp import FusionTrainer
    fus, sd, finp, vis, out1, out2 = _fusion_case(B, H, W, seed=5)
    with torch.no_grad():
        f_ref = O.fusion_network3_ac(finp["ir"], vis, out1, out2, sd)
        l_ref = 0.2 * O.fusionloss_grad3(finp["ir"], vis, f_ref, finp["mask"]) + 0.8 * O.seg_cross_entropy(
            O.network3_forward(O.recompose_rgb(f_ref, vis, clamp=False), ssd, "mit_b1", train_bn=True), finp["labels"])
    net = copy.deepcopy(fus).to(DEV).train()
    tr = FusionTrainer(net, Fusionloss_grad3(), lr=1.5e-4, max_iter=100, seg_net=seg, iter_=2)
    before = tr.flat.param.clone()
    l, _ = tr.step(finp["ir"].to(DEV), vis.to(DEV), out1.to(DEV), out2.to(DEV), finp["mask"].to(DEV), vis_rgb=finp["vis"].to(DEV),
                   labels=finp["labels"].to(DEV))
    res.append(result("composite_step_loss", abs(float(l) - float(l_ref)) / abs(float(l_ref)), 3e-2, note=f"{float(l):.5f} vs {float(l_ref):.5f}"))
    res.append(result("composite_step_updates_params", 0.0 if float((tr.flat.param - before).abs().max()) > 0 else 1.0, 0.0))
    return res


@check
def dropin_train_loop_bodies():
    """The reference's own loop bodies, statement for statement, on the mirror modules -- what `python -m
    segmif_b200.dropin train.py` executes: train_seg (train.py:207-226: model(mask) -> F.interpolate -> CrossEntropyLoss ->
    backward -> PolyWarmupAdamW_seg over WeTr.get_param_groups()) and train_fusion rounds >= 2 (train.py:350-381:
    forward_fusion under no_grad, model2(...), Fusionloss_grad3, torch-op YCrCb2RGB, model._loss, .item()-weighted sum,
    backward, PolyWarmupAdamW).  torch's own ops (interpolate, CE, the colour matrix product, the optimizers) run beside
    the segmif_b200 autograd nodes.  Losses of the first step are compared with the oracle; two more steps must run and
    move the parameters."""
    from segmif_b200.core import Fusionloss_grad3
    from segmif_b200.core.model_fusion import Fusion_Network3_ac, Network3, RGB2YCrCb
    from segmif_b200.utils.optimizer import PolyWarmupAdamW, PolyWarmupAdamW_seg
    res = []
    B, H, W = 2, 64, 96
    inp = synth.synth_inputs(B, H, W, seed=7)
    dev = torch.device(DEV)

    def no_drop(net):
        net.denoise_net.decoder.dropout.p = 0.0
        for m in net.modules():
            if hasattr(m, "drop_prob"):
                m.drop_prob = 0.0
        return net

    # ---------------- train_seg
    seg0 = synth.load_synthetic(Network3("mit_b1", 9, 256, None), 0)
    ssd = {k: v.clone() for k, v in seg0.state_dict().items()}
    with torch.no_grad():
        ref_seg_loss = float(O.seg_cross_entropy(O.network3_forward(inp["mask"], ssd, "mit_b1", train_bn=True), inp["labels"]))
    criterion_seg = torch.nn.CrossEntropyLoss(ignore_index=255).to(dev)
    model = no_drop(copy.deepcopy(seg0)).cuda()
    param_groups = model.denoise_net.get_param_groups()
    model = model.train()
    model.to(dev)
    optimizer = PolyWarmupAdamW_seg(
        params=[{"params": param_groups[0], "lr": 6e-5, "weight_decay": 0.01}, {"params": param_groups[1], "lr": 6e-5, "weight_decay": 0.0},
                {"params": param_groups[2], "lr": 6e-4, "weight_decay": 0.01}],
        lr=6e-5, weight_decay=0.01, betas=[0.9, 0.999], iter_curr=0, warmup_iter=2, max_iter=100, warmup_ratio=1e-6, power=1.0)
    inputs_mask, labels = inp["mask"].to(dev, non_blocking=True), inp["labels"].to(dev, non_blocking=True)
    before = model.denoise_net.decoder.linear_pred.weight.detach().clone()
    losses = []
    for n_iter in range(3):
        _, __, segmap = model(inputs_mask)
        outputs = F.interpolate(segmap, size=labels.shape[1:], mode='bilinear', align_corners=False)
        seg_loss = criterion_seg(outputs, labels.type(torch.long))
        optimizer.zero_grad()
        seg_loss.backward()
        optimizer.step()
        losses.append(seg_loss.item())
    res.append(result("dropin_train_seg_first_loss", abs(losses[0] - ref_seg_loss) / abs(ref_seg_loss), 2e-2, note=f"{losses}"))
    res.append(result("dropin_train_seg_finite_and_moving", 0.0 if all(map(lambda v: v == v and abs(v) < 1e4, losses))
                      and float((model.denoise_net.decoder.linear_pred.weight.detach() - before).abs().max()) > 0 else 1.0, 0.0))
    res.append(result("dropin_train_seg_classifier_grad_none", 0.0 if model.denoise_net.classifier.weight.grad is None else 1.0, 0.0))

    # ---------------- train_fusion, iter_ = 2
    def YCrCb2RGB(input_im):                                   # train.py:243-264 with .cuda() -> the input's device
        im_flat = input_im.transpose(1, 3).transpose(1, 2).reshape(-1, 3)
        mat = torch.tensor([[1.0, 1.0, 1.0], [1.403, -0.714, 0.0], [0.0, -0.344, 1.773]], device=input_im.device)
        bias = torch.tensor([0.0 / 255, -0.5, -0.5], device=input_im.device)
        temp = (im_flat + bias).mm(mat)
        return temp.reshape(list(input_im.size())[0], list(input_im.size())[2], list(input_im.size())[3], 3).transpose(1, 3).transpose(2, 3)

    iter_ = 2
    fus0 = synth.load_synthetic(Fusion_Network3_ac(), 0)
    fsd = {k: v.clone() for k, v in fus0.state_dict().items()}
    with torch.no_grad():
        vis_ref = O.rgb2ycrcb(inp["vis"])
        o0, o1 = O.mit_forward_fusion(inp["mask"], O._sub(ssd, "denoise_net.encoder"), "mit_b1")
        f_ref = O.fusion_network3_ac(inp["ir"][:, 0:1], vis_ref, o0, o1, fsd)
        l1_ref = float(O.fusionloss_grad3(inp["ir"], vis_ref, f_ref, inp["mask"]))
        l2_ref = float(O.seg_cross_entropy(O.network3_forward(O.recompose_rgb(f_ref, vis_ref, clamp=False), ssd, "mit_b1", train_bn=True),
                                           inp["labels"]))
    model = no_drop(copy.deepcopy(seg0)).cuda()                 # train mode, as train.py leaves it
    model2 = copy.deepcopy(fus0)
    model.to(dev)
    model2.to(dev)
    optimizer = PolyWarmupAdamW(params=[{"params": model2.parameters(), "lr": 3e-4 / iter_, "weight_decay": 0.01}], lr=(3e-4) / iter_,
                                weight_decay=0.01, betas=[0.9, 0.999], warmup_iter=(3e-5) / iter_, max_iter=100, warmup_ratio=1e-6, power=1.0)
    criterion_seg = torch.nn.CrossEntropyLoss(ignore_index=255)
    w_before = model2.conv2.weight.detach().clone()
    got = []
    for n_iter in range(3):
        inputs_ir = inp["ir"].to(dev, non_blocking=True)
        inputs_vis = inp["vis"].to(dev, non_blocking=True)
        inputs_mask = inp["mask"].to(dev, non_blocking=True)
        labels = inp["labels"].to(dev, non_blocking=True)
        inputs_ir = inputs_ir[:, 0:1, :, :]
        inputs_vis = RGB2YCrCb(inputs_vis)
        with torch.no_grad():
            out0, out1 = model.denoise_net.encoder.forward_fusion(inputs_mask)
        fusion = model2(inputs_ir, inputs_vis, out0, out1)
        optimizer.zero_grad()
        fusion_loss = Fusionloss_grad3()
        fused_ycbcr = inputs_vis.clone()
        fused_ycbcr[:, 0:1, :, :] = fusion
        fused_rgb = YCrCb2RGB(fused_ycbcr)
        loss1 = fusion_loss(inputs_ir, inputs_vis, fusion, inputs_mask)
        loss2 = model._loss(fused_rgb, labels, criterion_seg)
        seg_loss = (0.4 / iter_) * loss1 + 0.8 * loss2
        seg_loss.backward()
        optimizer.step()
        got.append((loss1.item(), loss2.item()))
    res.append(result("dropin_train_fusion_loss1", abs(got[0][0] - l1_ref) / abs(l1_ref), 3e-2, note=f"{got[0][0]:.5f} vs {l1_ref:.5f}"))
    res.append(result("dropin_train_fusion_loss2_ce", abs(got[0][1] - l2_ref) / abs(l2_ref), 3e-2, note=f"{got[0][1]:.5f} vs {l2_ref:.5f}"))
    res.append(result("dropin_train_fusion_finite_and_moving", 0.0 if all(v == v for pair in got for v in pair)
                      and float((model2.conv2.weight.detach() - w_before).abs().max()) > 0 else 1.0, 0.0))
    res.append(result("dropin_train_fusion_ffm2_grad_none", 0.0 if all(p.grad is None for k, p in model2.named_parameters() if k.startswith("ffm2.")) else 1.0, 0.0))
    return res


@check
def seg_training_after_eval():
    """train.py:232-236: val_segformer() leaves the model in eval() and train_seg keeps training -- running-statistics
    BatchNorm, no Dropout2d, no DropPath.  Network3._loss in that state: every parameter gradient against autograd over
    the oracle with train_bn=False (bound as in seg_network_backward), and the running statistics must not move."""
    from segmif_b200.core.seg_train import CeFn, logits_with_grad
    res = []
    net0, sd, names, x, drop, labels, cot, dps = _seg_case()
    leaves = {k: sd[k].clone().requires_grad_(True) for k in names}
    full = dict(sd)
    full.update(leaves)
    lg_ref = O.network3_forward(x, full, "mit_b1", train_bn=False)
    O.seg_cross_entropy(lg_ref, labels).backward()
    gmax = max(float(leaves[n].grad.abs().max()) for n in names)
    net = copy.deepcopy(net0).to(DEV).eval()
    rm0 = net.denoise_net.decoder.linear_fuse.bn.running_mean.clone()
    loss = net._loss(x.to(DEV), labels.to(DEV), torch.nn.CrossEntropyLoss(ignore_index=255))
    loss.backward()
    res.append(result("seg_eval_train_loss", rel_err(loss.detach(), O.seg_cross_entropy(lg_ref, labels).detach()), 2e-2))
    got = dict(net.named_parameters())
    worst, worst_name = 0.0, ""
    for k in names:
        if got[k].grad is None:
            res.append(result(f"seg_eval_grad_missing_{k}", float("nan"), 0.0))
            continue
        e = _floored_err(got[k].grad, leaves[k].grad, 1e-2 * gmax)
        if e > worst:
            worst, worst_name = e, k
    res.append(result("seg_eval_train_grad_worst", worst, 0.2, note=worst_name))
    for k in ("denoise_net.decoder.linear_fuse.bn.weight", "denoise_net.decoder.linear_fuse.bn.bias", "denoise_net.decoder.linear_fuse.conv.weight"):
        res.append(result(f"seg_eval_train_grad_{k.split('decoder.')[1]}", _floored_err(got[k].grad, leaves[k].grad, 1e-2 * gmax), 0.1))
    res.append(result("seg_eval_running_stats_untouched", float((net.denoise_net.decoder.linear_fuse.bn.running_mean - rm0).abs().max()), 0.0))
    return res


@check
def wgrad_lin_tcgen05():
    """Linear weight + bias gradient in one pass (csrc/wgrad_lin_tc.cu: both operands MN-major on tcgen05, ones-MMA for the
    column sums) against an fp64 contraction of the SAME bf16 operands.  fp32 accumulation over P tokens: <= 2e-5 of max |dW|
    (measured ~1e-6); every MiT-B2 layer shape class, token counts that are not multiples of the 64-token stage, padded
    operands (co_take / ci_take), channel slices of wider buffers, accumulation into a non-zero gradient."""
    res = []
    cases = [  # name, P, Cin, Cout, ldx, coffx, ldy, coffy, co_take, ci_take, bias
        ("s1_q_64x64", 5000, 64, 64, 64, 0, 64, 0, None, None, True),
        ("s1_fc1_64x256", 4100, 64, 256, 64, 0, 256, 0, None, None, True),
        ("s1_fc2_256x64", 4100, 256, 64, 256, 0, 64, 0, None, None, True),
        ("s3_fc1_320x1280", 1234, 320, 1280, 320, 0, 1280, 0, None, None, True),
        ("s4_fc2_2048x512", 300, 2048, 512, 2048, 0, 512, 0, None, None, True),
        ("pred_pad32_take9", 3001, 256, 32, 256, 0, 32, 0, 9, None, True),
        ("patch_embed_kpad", 2000, 152, 64, 152, 0, 64, 0, None, 147, True),
        ("drdb_1x1_slice", 2500, 224, 64, 224, 0, 96, 32, None, None, False),
        ("x_slice_of_wider", 1500, 64, 128, 192, 64, 128, 0, None, None, True),
        ("tiny_P_63", 63, 128, 320, 128, 0, 320, 0, None, None, True),
        ("single_chunk_P_64", 64, 512, 512, 512, 0, 512, 0, None, None, True),
    ]
    for name, P, Cin, Cout, ldx, coffx, ldy, coffy, co_take, ci_take, bias in cases:
        x = rnd(P, ldx, seed=11).to(torch.bfloat16)
        dy = (rnd(P, ldy, seed=12) * 0.05 + 0.01).to(torch.bfloat16)
        ct, it = co_take or Cout, ci_take or Cin
        g0 = rnd(ct, it, seed=13, bf16=False)
        b0 = rnd(Cout, seed=14, bf16=False)
        xs, ys = x[:, coffx:coffx + Cin].double(), dy[:, coffy:coffy + Cout].double()
        ref_w = g0.double() + (ys.t() @ xs)[:ct, :it]
        ref_b = b0.double() + ys.sum(0)
        grad, dbias = g0.clone().to(DEV), b0.clone().to(DEV)
        ops.wgrad_lin(dy.to(DEV), ldy, coffy, x.to(DEV), ldx, coffx, P=P, Cin=Cin, Cout=Cout, grad=grad, s_co=it, co_take=co_take,
                      ci_take=ci_take, dbias=dbias if bias else None)
        res.append(result(f"wgrad_lin_{name}", rel_err(grad, ref_w), 2e-5))
        if bias:
            res.append(result(f"wgrad_lin_{name}_bias", rel_err(dbias, ref_b), 2e-5))
    return res
