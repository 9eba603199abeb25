"""Training path of Fusion_Network3_ac (core/model_fusion.py:1047-1067 forward, and the backward torch.autograd derives
for it in train.py:381-386): forward with every tensor the backward needs kept in pixel-major bf16, and a hand-written
backward that walks the network once in reverse using only segmif_b200 kernels:

  conv22 / conv21 / conv2 / conv1   act_bwd (PReLU from the stored output) -> wgrad (tensor-core contraction over all
                                    pixels) -> dgrad = the forward tcgen05 3x3 kernel on flipped, transposed weights
  FeatureFusionModule (x2, shared)  layernorm_bwd -> ffm_bwd_gram -> ffm_bwd_ctx -> ffm_bwd_apply -> wgrad / GEMM
  DRDB x4                           1x1: act_bwd, wgrad, GEMM with W^T; growth layers 5..1: act_bwd on the slab of the
                                    224-channel gradient buffer, wgrad, and the dilated dgrad accumulated IN PLACE into
                                    the lower channels of the same buffer (pre_add == dst), mirroring the forward's
                                    in-place growth buffer
  conv3 / conv4                     explicit 1x1 GEMMs in training (not folded into channel_proj3) so their weight
                                    gradients are plain wgrad calls.

Parameter gradients are accumulated in fp32 by the kernels into zero-initialised tensors and handed to autograd.
Restrictions (documented in DESIGN.md): the shared PReLU slope must stay > 0 (its pre-activations are recovered from
the stored outputs); the encoder features out1 / out2 and the images receive no gradient (train.py:358-359 computes
them under no_grad)."""
import torch

from .. import ops
from ..ops import ACT_PRELU, ACT_RELU

G = 224          # DRDB growth-buffer pitch


def _pm(x):
    """Logical NCHW -> pixel-major bf16 [B*HW, C]."""
    from .model_fusion import _pixel_major_bf16
    B, C, H, W = x.shape
    return _pixel_major_bf16(x).reshape(B * H * W, C)


# ------------------------------------------------------------------------------------------------ weight packs
def _dgrad_pack(w, pad_cout=None):
    """nn.Conv2d weight [Cout, Cin, 3, 3] -> the weight of the transposed convolution as a forward conv:
    bf16 [Cin, 9, Cout] with the taps flipped (dX = corr(dY, flip(W)^T))."""
    w = w.detach().float().flip(2, 3).permute(1, 2, 3, 0).reshape(w.shape[1], 9, w.shape[0])
    if pad_cout is not None and pad_cout > w.shape[2]:
        w = torch.cat([w, w.new_zeros(w.shape[0], 9, pad_cout - w.shape[2])], 2)
    return w.to(torch.bfloat16).contiguous()


def _t_pack(w2d):
    """[N, K] linear / 1x1 weight -> bf16 [K, 1, N]: the operand of dX = dY @ W."""
    return w2d.detach().float().t().to(torch.bfloat16).reshape(w2d.shape[1], 1, w2d.shape[0]).contiguous()


def _drdb_bwd_packs(m):
    convs = [getattr(m, f"Dcov{i}") for i in range(1, 6)]

    def build(*ws):
        return dict(wt1x1=_t_pack(ws[5].reshape(ws[5].shape[0], -1)), wd=[_dgrad_pack(w) for w in ws[:5]])
    return m._packs.get_multi([c.weight for c in convs] + [m.conv.weight], build, "bwd")


def _ffm_bwd_packs(cp):
    plist = [cp.channel_proj1.weight, cp.channel_proj1.bias, cp.channel_proj2.weight, cp.channel_proj2.bias,
             cp.channel_proj3.weight, cp.channel_proj3.bias]

    def build(w1, b1, w2, b2, w3, b3):
        f = lambda t: t.detach().float()
        return dict(wfull=torch.stack([f(w1), f(w2), f(w3)]).to(torch.bfloat16).contiguous(),
                    bfull=torch.stack([f(b1), f(b2), f(b3)]).contiguous(),
                    wt=[_t_pack(w1), _t_pack(w2), _t_pack(w3)])
    return cp._packs.get_multi(plist, build, "ffm_bwd")


def _zeros_bias(n, dev, cache={}):
    key = (n, dev)
    if key not in cache:
        cache[key] = torch.zeros((n,), dtype=torch.float32, device=dev)
    return cache[key]


# ------------------------------------------------------------------------------------------------ forward
def train_forward(net, ir, vis, out1, out2):
    B, _, H, W = ir.shape
    N, HW, dev = B * H * W, H * W, ir.device
    alpha = net.relu.weight.detach()
    ir0 = ir[:, :1].float().contiguous()
    vis0 = vis[:, :1].float().contiguous()
    bufs = [torch.empty((B, H, W, G), dtype=torch.bfloat16, device=dev) for _ in range(4)]
    ops.conv3x3_in1(ir0, net._packs.taps_f32(net.conv1_ir.weight), net.conv1_ir.bias.detach(), alpha, bufs[0], G, 0, 64)
    ops.conv3x3_in1(vis0, net._packs.taps_f32(net.conv1_vis.weight), net.conv1_vis.bias.detach(), alpha, bufs[1], G, 0, 64)
    part = torch.empty((B, H, W, 128), dtype=torch.bfloat16, device=dev)

    def drdb(m, buf):
        m._growth_hybrid(buf, part, B, H, W)
        r = ops.conv(buf, m._packs.conv(m.conv.weight), m.conv.bias.detach(), B=B, H=H, W=W, Cin=G, ld_src=G, Cout=64,
                     act=ACT_RELU)                                                    # relu(conv1x1), [N, 64]
        x = torch.empty((N, 64), dtype=torch.bfloat16, device=dev)
        ops.add_bf16(buf, G, 0, r, 64, 0, x, 64, 0, N, 64)                            # + block input (residual)
        return x, r

    def seg_proj(conv, feat):
        t = _pm(feat)
        return t, ops.conv(t, net._packs.conv(conv.weight), conv.bias.detach(), B=1, H=1, W=N, Cin=t.shape[1], Cout=64)

    pk = net.ffm.cross.packs(None)
    x1, r1 = drdb(net.DRDB1, bufs[0])
    x2, r2 = drdb(net.DRDB2, bufs[1])
    f1_in, s3a = seg_proj(net.conv3, out1)
    fa = ops.ffm_train_fwd(x1, 64, 0, x2, 64, 0, s3a, 64, pk, bufs[2], G, 0, bufs[3], G, 0, B, HW)
    x3, r3 = drdb(net.DRDB3, bufs[2])
    x4, r4 = drdb(net.DRDB4, bufs[3])
    f2_in, s3b = seg_proj(net.conv4, out2)
    cat = torch.empty((N, 128), dtype=torch.bfloat16, device=dev)
    fb = ops.ffm_train_fwd(x3, 64, 0, x4, 64, 0, s3b, 64, pk, cat, 128, 0, cat, 128, 64, B, HW)
    o2 = ops.conv(cat, net._packs.conv(net.conv2.weight), net.conv2.bias.detach(), B=B, H=H, W=W, Cin=128, KH=3, KW=3,
                  pad=1, Cout=64, act=ACT_PRELU, prelu_alpha=alpha)
    o21 = ops.conv(o2, net._packs.conv(net.conv21.weight), net.conv21.bias.detach(), B=B, H=H, W=W, Cin=64, KH=3, KW=3,
                   pad=1, Cout=32, act=ACT_PRELU, prelu_alpha=alpha)
    fused = ops.conv3x3_out1(o21, net._packs.taps_f32(net.conv22.weight), net.conv22.bias.detach(), alpha, B, H, W, 32)
    saved = dict(B=B, H=H, W=W, ir0=ir0, vis0=vis0, bufs=bufs, r=[r1, r2, r3, r4], x=[x1, x2, x3, x4], s3=[s3a, s3b],
                 seg_in=[f1_in, f2_in], ffm=[fa, fb], cat=cat, o2=o2, o21=o21, fused=fused)
    return fused, saved


# ------------------------------------------------------------------------------------------------ backward
def _conv_dgrad(dy, cin_dy, wd, B, H, W, dil, out, ld_out, coff_out, accumulate):
    """out[:, coff : coff + wd.shape[0]] (+)= transposed conv of dy; launched in slices of 64 / 32 output channels."""
    n_out = wd.shape[0]
    c0 = 0
    while c0 < n_out:
        w = 64 if n_out - c0 >= 64 else 32
        ops.conv(dy, wd[c0:c0 + w], _zeros_bias(w, dy.device), B=B, H=H, W=W, Cin=cin_dy, KH=3, KW=3, pad=dil, dil=dil,
                 Cout=w, out=out, ld_dst=ld_out, dst_coff=coff_out + c0, pre_add=out if accumulate else None,
                 pre_coff=coff_out + c0, entry="segmif_conv3x3_tc_fwd")
        c0 += w
    return out


WGRAD_SIDE = None           # a core.seg_train.SideWgrad installed by ddp.FusionTrainer around loss.backward()


def _side(fn, *operands):
    """Weight-gradient launches are off the critical path of the reverse pass: with WGRAD_SIDE installed they go to a second
    stream (a parallel branch of the step's CUDA graph) and overlap the HBM-bound activation / data-gradient kernels; their
    operands stay referenced until the join, and no operand may be overwritten later in the pass (fresh buffers below)."""
    if WGRAD_SIDE is not None:
        WGRAD_SIDE.run(fn, *operands)
    else:
        fn()


def _drdb_backward(m, buf, r, dout, ld_do, coff_do, g, prefix, B, H, W):
    """dout: gradient of the block output [N, .] (slice ld_do / coff_do).  Returns dbuf [N, 224] whose first 64
    channels are the gradient of the block input."""
    N, dev = B * H * W, buf.device
    pk = _drdb_bwd_packs(m)
    buf2 = buf.view(N, G)
    dz = torch.empty((N, 64), dtype=torch.bfloat16, device=dev)
    ops.act_bwd(r, 64, 0, dout, ld_do, coff_do, dz, 64, 0, N, 64, ACT_RELU, dbias=g[prefix + "conv.bias"])
    _side(lambda: ops.wgrad_lin(dz, 64, 0, buf2, G, 0, P=N, Cin=G, Cout=64, grad=g[prefix + "conv.weight"], s_co=G), dz, buf2)
    dbuf = torch.empty((N, G), dtype=torch.bfloat16, device=dev)
    wt = pk["wt1x1"]                                                                   # [224, 1, 64]
    ops.linear_tc(dz, wt[:64], None, residual=dout, ld_res=ld_do, res_coff=coff_do, out=dbuf, ld_dst=G, dst_coff=0)
    ops.linear_tc(dz, wt[64:], None, out=dbuf, ld_dst=G, dst_coff=64)
    for j in range(5, 0, -1):
        cin = 64 + 32 * (j - 1)
        dg = torch.empty((N, 32), dtype=torch.bfloat16, device=dev)      # fresh per layer: the weight gradient may still be reading the last one
        ops.act_bwd(buf2, G, cin, dbuf, G, cin, dg, 32, 0, N, 32, ACT_RELU, dbias=g[f"{prefix}Dcov{j}.bias"])
        _side(lambda: ops.wgrad(dg, 32, 0, buf2, G, 0, B=B, H=H, W=W, Cin=cin, Cout=32, taps=9, dil=2,
                                grad=g[f"{prefix}Dcov{j}.weight"], s_co=cin * 9, s_tap=1, s_ci=9), dg, buf2)
        _conv_dgrad(dg, 32, pk["wd"][j - 1], B, H, W, 2, dbuf, G, 0, accumulate=True)
    return dbuf


def _ffm_backward(net, pk, bp, x1, x2, s3, fw, do1, ld1, coff1, do2, ld2, coff2, g, B, HW):
    """Returns (dx1, dx2, ds3), each [N, 64] bf16."""
    N, dev = B * HW, x1.device
    cp = "ffm.cross."
    dr = []
    for i, (pre, do, ld, coff) in enumerate(((fw["pre1"], do1, ld1, coff1), (fw["pre2"], do2, ld2, coff2)), 1):
        d = torch.empty((N, 64), dtype=torch.bfloat16, device=dev)
        norm = getattr(net.ffm.cross, f"norm{i}")
        ops.layernorm_bwd(pre, do, ld, coff, norm.weight.detach(), norm.eps, d, 64, 0, N, 64, dgamma=g[f"{cp}norm{i}.weight"],
                          dbeta=g[f"{cp}norm{i}.bias"], dxsum=g[f"{cp}end_proj{i}.bias"])
        dr.append(d)
    rpart, nchunk_r = ops.ffm_bwd_gram(x1, 64, 0, x2, 64, 0, s3, 64, 0, dr[0], dr[1], bp["wfull"], bp["bfull"], B, HW)
    mats = ops.ffm_bwd_ctx(rpart, nchunk_r, fw["partials"], fw["nchunk"], fw["ctx"], pk["wkv"], pk["wend"], fw["folded"],
                           g["_dwkv"], g["_dwend"], B)
    dP = ops.ffm_bwd_apply(x1, 64, 0, x2, 64, 0, s3, 64, 0, dr[0], dr[1], bp["wfull"], bp["bfull"], mats, B, HW)
    outs = []
    for i, (dp, x, res) in enumerate(((dP[0], x1, dr[0]), (dP[1], x2, dr[1]), (dP[2], s3, None)), 1):
        _side(lambda: ops.wgrad_lin(dp, 128, 0, x, 64, 0, P=N, Cin=64, Cout=128, grad=g[f"{cp}channel_proj{i}.weight"], s_co=64,
                                    dbias=g[f"{cp}channel_proj{i}.bias"]), dp, x)
        outs.append(ops.linear_tc(dp, bp["wt"][i - 1], None, residual=res))
    return outs


def train_backward(net, sv, dfused):
    B, H, W = sv["B"], sv["H"], sv["W"]
    N, HW = B * H * W, H * W
    dev = dfused.device
    alpha = net.relu.weight.detach()
    used = [(k, p) for k, p in net.named_parameters() if not k.startswith("ffm2.")]
    g = {k: torch.zeros(p.shape, dtype=torch.float32, device=dev) for k, p in used}
    g["_dwkv"] = torch.zeros((3, 128, 64), dtype=torch.float32, device=dev)
    g["_dwend"] = torch.zeros((2, 64, 128), dtype=torch.float32, device=dev)
    dalpha = g["relu.weight"]
    dfused = dfused.float().contiguous()

    # conv22 (32 -> 1): the single gradient plane is padded to a 32-channel pixel-major tensor (channel 0 live)
    dz22 = torch.zeros((N, 32), dtype=torch.bfloat16, device=dev)
    ops.prelu_plane_bwd(sv["fused"], dfused, alpha, dz22, 32, 0, dbias=g["conv22.bias"], dalpha=dalpha)
    _side(lambda: ops.wgrad(dz22, 32, 0, sv["o21"], 32, 0, B=B, H=H, W=W, Cin=32, Cout=32, taps=9, dil=1, grad=g["conv22.weight"],
                            s_co=288, s_tap=1, s_ci=9, co_take=1), dz22, sv["o21"])
    wd22 = net._packs.get(net.conv22.weight, lambda w: _dgrad_pack(w, 32), "dgrad")
    do21 = torch.empty((N, 32), dtype=torch.bfloat16, device=dev)
    _conv_dgrad(dz22, 32, wd22, B, H, W, 1, do21, 32, 0, accumulate=False)
    # conv21 (64 -> 32)
    ops.act_bwd(sv["o21"], 32, 0, do21, 32, 0, do21, 32, 0, N, 32, ACT_PRELU, alpha=alpha, dbias=g["conv21.bias"], dalpha=dalpha)
    _side(lambda: ops.wgrad(do21, 32, 0, sv["o2"], 64, 0, B=B, H=H, W=W, Cin=64, Cout=32, taps=9, dil=1, grad=g["conv21.weight"],
                            s_co=576, s_tap=1, s_ci=9), do21, sv["o2"])
    do2 = torch.empty((N, 64), dtype=torch.bfloat16, device=dev)
    _conv_dgrad(do21, 32, net._packs.get(net.conv21.weight, _dgrad_pack, "dgrad"), B, H, W, 1, do2, 64, 0, accumulate=False)
    # conv2 (128 -> 64)
    ops.act_bwd(sv["o2"], 64, 0, do2, 64, 0, do2, 64, 0, N, 64, ACT_PRELU, alpha=alpha, dbias=g["conv2.bias"], dalpha=dalpha)
    _side(lambda: ops.wgrad(do2, 64, 0, sv["cat"], 128, 0, B=B, H=H, W=W, Cin=128, Cout=64, taps=9, dil=1, grad=g["conv2.weight"],
                            s_co=1152, s_tap=1, s_ci=9), do2, sv["cat"])
    dcat = torch.empty((N, 128), dtype=torch.bfloat16, device=dev)
    _conv_dgrad(do2, 64, net._packs.get(net.conv2.weight, _dgrad_pack, "dgrad"), B, H, W, 1, dcat, 128, 0, accumulate=False)

    pk = net.ffm.cross.packs(None)
    bp = _ffm_bwd_packs(net.ffm.cross)
    x1, x2, x3, x4 = sv["x"]
    # second application of ffm (inputs x3, x4, conv4(out2))
    dx3, dx4, ds3b = _ffm_backward(net, pk, bp, x3, x4, sv["s3"][1], sv["ffm"][1], dcat, 128, 0, dcat, 128, 64, g, B, HW)
    _side(lambda: ops.wgrad_lin(ds3b, 64, 0, sv["seg_in"][1], 128, 0, P=N, Cin=128, Cout=64, grad=g["conv4.weight"], s_co=128,
                                dbias=g["conv4.bias"]), ds3b, sv["seg_in"][1])
    db3 = _drdb_backward(net.DRDB3, sv["bufs"][2], sv["r"][2], dx3, 64, 0, g, "DRDB3.", B, H, W)
    db4 = _drdb_backward(net.DRDB4, sv["bufs"][3], sv["r"][3], dx4, 64, 0, g, "DRDB4.", B, H, W)
    # first application of ffm (inputs x1, x2, conv3(out1)); its outputs were the inputs of DRDB3 / DRDB4
    dx1, dx2, ds3a = _ffm_backward(net, pk, bp, x1, x2, sv["s3"][0], sv["ffm"][0], db3, G, 0, db4, G, 0, g, B, HW)
    _side(lambda: ops.wgrad_lin(ds3a, 64, 0, sv["seg_in"][0], 64, 0, P=N, Cin=64, Cout=64, grad=g["conv3.weight"], s_co=64,
                                dbias=g["conv3.bias"]), ds3a, sv["seg_in"][0])
    del db3, db4
    db1 = _drdb_backward(net.DRDB1, sv["bufs"][0], sv["r"][0], dx1, 64, 0, g, "DRDB1.", B, H, W)
    db2 = _drdb_backward(net.DRDB2, sv["bufs"][1], sv["r"][1], dx2, 64, 0, g, "DRDB2.", B, H, W)
    # conv1_ir / conv1_vis (1 -> 64): the image plane is padded to 8 channels so the generic wgrad applies
    for name, plane, buf, db in (("conv1_ir", sv["ir0"], sv["bufs"][0], db1), ("conv1_vis", sv["vis0"], sv["bufs"][1], db2)):
        dz1 = torch.empty((N, 64), dtype=torch.bfloat16, device=dev)
        ops.act_bwd(buf.view(N, G), G, 0, db, G, 0, dz1, 64, 0, N, 64, ACT_PRELU, alpha=alpha, dbias=g[name + ".bias"], dalpha=dalpha)
        x8 = torch.zeros((B, HW, 8), dtype=torch.bfloat16, device=dev)
        ops.nchw_to_nhwc(plane, out=x8, ld_dst=8, dst_coff=0)
        _side(lambda: ops.wgrad(dz1, 64, 0, x8, 8, 0, B=B, H=H, W=W, Cin=8, Cout=64, taps=9, dil=1, grad=g[name + ".weight"],
                                s_co=9, s_tap=1, s_ci=9, ci_take=1), dz1, x8)
    # scatter the stacked kv / end_proj gradients to their parameters
    ca = "ffm.cross."
    g[ca + "cross_attn2.kv1.weight"] = g["_dwkv"][0]
    g[ca + "cross_attn2.kv2.weight"] = g["_dwkv"][1]
    g[ca + "cross_attn.kv3.weight"] = g["_dwkv"][2]
    g[ca + "end_proj1.weight"] = g["_dwend"][0]
    g[ca + "end_proj2.weight"] = g["_dwend"][1]
    if WGRAD_SIDE is not None:          # autograd accumulates the returned tensors on this stream as soon as we return
        WGRAD_SIDE.join()
    return g


class FusionNetFn(torch.autograd.Function):
    """autograd node of the whole fusion network: forward = train_forward, backward = train_backward."""

    @staticmethod
    def forward(ctx, net, ir, vis, out1, out2, names, *params):
        fused, saved = train_forward(net, ir, vis, out1, out2)
        ctx.net, ctx.saved_state, ctx.names = net, saved, names
        return fused

    @staticmethod
    def backward(ctx, dfused):
        g = train_backward(ctx.net, ctx.saved_state, dfused)
        ctx.saved_state = None
        grads = tuple(g.get(n) if ctx.needs_input_grad[6 + i] else None for i, n in enumerate(ctx.names))
        return (None, None, None, None, None, None) + grads


def forward_with_grad(net, ir, vis, out1, out2):
    for t in (ir, vis, out1, out2):
        if t.requires_grad:
            raise NotImplementedError("segmif_b200: Fusion_Network3_ac's backward produces parameter gradients only; the "
                                      "images and encoder features must not require grad (train.py:358-360 detaches them)")
    named = [(k, p) for k, p in net.named_parameters() if not k.startswith("ffm2.")]
    names = tuple(k for k, _ in named)
    return FusionNetFn.apply(net, ir, vis, out1, out2, names, *[p for _, p in named])
