"""Training path of the segmentation network Network3 / WeTr (core/model_fusion.py:62-68,1081-1097: MiT encoder +
SegFormer head) -- what torch.autograd derives for `seg_loss.backward()` in train.py:222-226 (train_seg) and for the CE
term of train_fusion (train.py:368), written out as one explicit reverse pass over segmif_b200 kernels.

Forward (training mode): the inference kernels, except that
  * the strided convolutions (OverlapPatchEmbed.proj, Attention.sr) run as im2col (torch pad/unfold/reshape: data
    movement only) + the tcgen05 GEMM, so that their weight / data gradients are the generic wgrad / GEMM / col2im kernels;
  * the attention core also stores its row log-sum-exp; BatchNorm uses batch statistics and updates the running ones;
    Dropout2d and DropPath draw their masks with torch's device RNG (or take injected masks, for parity tests);
  * every tensor the backward needs stays resident (bf16 activations, fp32 residual stream).
Backward per op:  nn.Linear -> wgrad_kernel<1> + colsum + gemm_tc with W^T;  nn.LayerNorm -> layernorm_bwd (accumulating
into the fp32 residual-stream gradient);  attention core -> sr_attention_bwd;  DWConv+GELU -> dwconv3x3_gelu_bwd +
dwconv3x3(flip);  BN+ReLU -> bn_train_bwd;  bilinear -> bilinear_nhwc_bwd;  upsample+CE -> upsample_ce_bwd;
im2col convs -> wgrad / GEMM / col2im.  WeTr.classifier never runs (core/model_fusion.py:66 discards it) and gets no
gradient, exactly as in the reference."""
import torch
import torch.nn.functional as F

from .. import ops
from ..ops import ACT_NONE

BF16, F32 = torch.bfloat16, torch.float32
STAGE_DONE_HOOK = None      # set by ddp.SegTrainer: called with the 0-based stage index when the reverse pass has finished a stage


# ------------------------------------------------------------------------------------------------ helpers
def _t_pack(cache, w2d_param, tag):
    """[N, K] weight -> bf16 [K, 1, N] (operand of dX = dY @ W); cached on the owning module's PackCache."""
    return cache.get(w2d_param, lambda w: w.detach().reshape(w.shape[0], -1).float().t().to(BF16)
                     .reshape(-1, 1, w.shape[0]).contiguous(), tag)


def _flat_pack(cache, w_param, kpad, tag):
    """conv weight [Cout, Cin, k, k] -> bf16 [Cout, 1, Kp] in (c, ky, kx) order, zero-padded to Kp columns."""
    def f(w):
        w2 = w.detach().reshape(w.shape[0], -1).float()
        if kpad > w2.shape[1]:
            w2 = torch.cat([w2, w2.new_zeros(w2.shape[0], kpad - w2.shape[1])], 1)
        return w2.to(BF16).reshape(w2.shape[0], 1, kpad).contiguous()
    return cache.get(w_param, f, tag)


def _flat_t_pack(cache, w_param, kpad, tag):
    def f(w):
        w2 = w.detach().reshape(w.shape[0], -1).float()
        if kpad > w2.shape[1]:
            w2 = torch.cat([w2, w2.new_zeros(w2.shape[0], kpad - w2.shape[1])], 1)
        return w2.t().to(BF16).reshape(kpad, 1, w2.shape[0]).contiguous()
    return cache.get(w_param, f, tag)


class _Grads(dict):
    """name -> fp32 gradient tensor.  `frozen`: the parameters take no gradient (train_fusion differentiates through the
    frozen segmentation network, train.py:368): the weight-gradient GEMMs and bias column sums are skipped."""
    frozen = False


class SideWgrad:
    """Weight-gradient launches are off the critical path of the reverse pass (nothing reads them before the all-reduce), and
    most of them are small (19-114 blocks): with this object installed in WGRAD_SIDE they are queued on a second stream -- a
    parallel branch of the step's CUDA graph -- and fill the SMs the dgrad chain leaves idle.  The operands are kept referenced
    until `join`, so the caching allocator cannot hand their memory to a later kernel of the main stream."""

    def __init__(self, device=None):
        self.stream = torch.cuda.Stream(device=device)
        self.keep, self.used = [], False

    def run(self, fn, *operands):
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            fn()
        self.keep.append(operands)
        self.used = True

    def join(self):
        if self.used:
            torch.cuda.current_stream().wait_stream(self.stream)
        self.keep.clear()
        self.used = False


WGRAD_SIDE = None           # set by ddp.SegTrainer around loss.backward()


def _wgrad_lin(g, dy, x, **kw):
    """Parameter-side backward of a linear layer (ops.wgrad_lin) unless the parameters are frozen."""
    if g.frozen:
        return
    fn = lambda: ops.wgrad_lin(dy, kw.pop("ldy"), 0, x, kw.pop("ldx"), 0, **kw)
    if WGRAD_SIDE is not None:
        WGRAD_SIDE.run(fn, dy, x)
    else:
        fn()


def _wgrad(g, *args, **kw):
    if not g.frozen:
        ops.wgrad(*args, **kw)


def _colsum(dy, N, rows, out, g=None):
    if g is not None and g.frozen:
        return
    ops.colsum(dy, N, 0, rows, N, out)


def _lin_bwd(dy, x, weight_param, cache, g, wname, bname, need_dx=True, residual=None):
    """dy bf16 [M, N], x bf16 [M, K] -> weight / bias gradients accumulated; returns dx bf16 [M, K] (+ residual)."""
    M, N = dy.shape
    K = x.shape[1]
    # weight and bias gradient from one pass over dy and x (tcgen05, csrc/wgrad_lin_tc.cu)
    _wgrad_lin(g, dy, x, ldy=N, ldx=K, P=M, Cin=K, Cout=N, grad=g[wname], s_co=K, dbias=g[bname] if bname is not None else None)
    if not need_dx:
        return None
    return _dgrad(dy, cache.linear(weight_param), lambda: _t_pack(cache, weight_param, "lin_t"), K, residual)


def _dgrad(dy, fwd_pack, t_pack_fn, n_in, residual=None):
    """dX = dY W.  The forward pack [out, 1, in] is read as a [K, N] operand (MN-major tcgen05 B descriptor) -- no
    transposed copy of the weights -- whenever the input width is a multiple of 64; otherwise the transposed pack."""
    if n_in % 64 == 0:
        return ops.linear_tc(dy, fwd_pack, None, residual=residual, weight_kn=True)
    return ops.linear_tc(dy, t_pack_fn(), None, residual=residual)


def _im2col_nhwc(x, B, H, W, C, k, s, p):
    """bf16 pixel-major [B, H, W, C] -> [B*Ho*Wo, C*k*k] with column (c*k + ky)*k + kx (conv weight's own flattening)."""
    xp = F.pad(x.view(B, H, W, C), (0, 0, p, p, p, p)) if p else x.view(B, H, W, C)
    pt = xp.unfold(1, k, s).unfold(2, k, s)                       # [B, Ho, Wo, C, k, k]
    Ho, Wo = pt.shape[1], pt.shape[2]
    return pt.reshape(B * Ho * Wo, C * k * k), Ho, Wo


def _patches_to_map(dP, B, Hk, Wk, C, r, H, W):
    """non-overlapping k = s = r patches [B*Hk*Wk, C*r*r] (c, ky, kx) -> pixel-major [B*H*W, C] (zero where uncovered)."""
    t = dP.view(B, Hk, Wk, C, r, r).permute(0, 1, 4, 2, 5, 3).reshape(B, Hk * r, Wk * r, C)
    if Hk * r == H and Wk * r == W:
        return t.reshape(B * H * W, C)
    full = torch.zeros((B, H, W, C), dtype=dP.dtype, device=dP.device)
    full[:, :Hk * r, :Wk * r] = t
    return full.view(B * H * W, C)


def _droppath_scale(blk, B, dev, injected):
    dp = getattr(blk.drop_path, "drop_prob", 0.0)
    if injected is not None:
        return injected
    if dp == 0.0 or not blk.training:
        return None
    keep = 1.0 - dp
    return (torch.rand((B,), device=dev) < keep).float() / keep        # timm.DropPath: bernoulli(keep) / keep per sample


# ------------------------------------------------------------------------------------------------ forward
def _block_forward(blk, x, B, N, H, W, dp_scales):
    """x fp32 [B*N, C] -> fp32 [B*N, C] and the saved tensors."""
    C = x.shape[1]
    at, ml = blk.attn, blk.mlp
    heads, D = at.num_heads, C // at.num_heads
    sv = dict(x=x, H=H, W=W)
    n1 = ops.layernorm(x, blk.norm1.weight.detach(), blk.norm1.bias.detach(), blk.norm1.eps)
    q = ops.linear(n1, at._packs.linear(at.q.weight), at._b(at.q))
    if at.sr_ratio > 1:
        r = at.sr_ratio
        P, Hk, Wk = _im2col_nhwc(n1, B, H, W, C, r, r, 0)
        red = ops.linear(P, _flat_pack(at._packs, at.sr.weight, C * r * r, "sr_flat"), at.sr.bias.detach(), out_dtype=F32)
        src = ops.layernorm(red, at.norm.weight.detach(), at.norm.bias.detach(), at.norm.eps)
        Nk = Hk * Wk
        sv.update(P_sr=P, red=red, Hk=Hk, Wk=Wk)
    else:
        src, Nk = n1, N
    kv = ops.linear(src, at._packs.linear(at.kv.weight), at._b(at.kv))
    att, lse = ops.sr_attention_train(q, kv, B, heads, N, Nk, D, at.scale)
    s1, s2 = dp_scales
    # x2 = x + s1 * proj(att): the DropPath factor rides in the GEMM epilogue (row_scale), like the residual add
    x2 = ops.linear_tc(att, at._packs.linear(at.proj.weight), at.proj.bias.detach(), residual=x, out_dtype=F32,
                       row_scale=s1, rows_per_scale=N)
    n2 = ops.layernorm(x2, blk.norm2.weight.detach(), blk.norm2.bias.detach(), blk.norm2.eps)
    h1 = ops.linear(n2, ml._packs.linear(ml.fc1.weight), ml.fc1.bias.detach())
    h2 = ml.dwconv.forward_gelu(h1.view(B, N, -1), H, W).view(B * N, -1)
    x3 = ops.linear_tc(h2, ml._packs.linear(ml.fc2.weight), ml.fc2.bias.detach(), residual=x2, out_dtype=F32,
                       row_scale=s2, rows_per_scale=N)
    sv.update(n1=n1, q=q, src=src, kv=kv, att=att, lse=lse, Nk=Nk, x2=x2, n2=n2, h1=h1, h2=h2, s1=s1, s2=s2)
    return x3, sv


def encoder_forward(enc, x, in_scale, in_shift, masks):
    """x fp32 NCHW [B, 3, H, W] -> four (tokens bf16 [B*N, C], H_s, W_s) and the tape."""
    B, _, H0, W0 = x.shape
    tape = dict(B=B, H0=H0, W0=W0, stages=[], in_scale=in_scale)
    xn = ops.channel_affine_nchw(x.float().contiguous(), in_scale, in_shift) if in_scale is not None else x.float().contiguous()
    outs = []
    prev, H, W = None, H0, W0
    bi = 0
    for s in range(4):
        pe = getattr(enc, f"patch_embed{s + 1}")
        k, st, p = pe.patch_size[0], pe.stride, pe.patch_size[0] // 2
        Cout = pe.proj.out_channels
        if s == 0:
            pt = F.pad(xn, (p, p, p, p)).unfold(2, k, st).unfold(3, k, st)                  # [B, 3, Ho, Wo, k, k]
            Ho, Wo = pt.shape[2], pt.shape[3]
            Kp = (3 * k * k + 31) // 32 * 32             # the data-gradient GEMM has N = Kp (multiple of 32)
            P = torch.zeros((B * Ho * Wo, Kp), dtype=BF16, device=x.device)
            P[:, :3 * k * k] = pt.permute(0, 2, 3, 1, 4, 5).reshape(B * Ho * Wo, 3 * k * k)
            Cin = 3
        else:
            Cin = prev.shape[1]
            P, Ho, Wo = _im2col_nhwc(prev, B, H, W, Cin, k, st, p)
            Kp = Cin * k * k
        y = ops.linear(P, _flat_pack(pe._packs, pe.proj.weight, Kp, "pe_flat"), pe.proj.bias.detach(), out_dtype=F32)
        tok = ops.layernorm(y, pe.norm.weight.detach(), pe.norm.bias.detach(), pe.norm.eps, out_dtype=F32)
        Hin, Win, H, W = H, W, Ho, Wo
        N = H * W
        blocks = []
        for blk in getattr(enc, f"block{s + 1}"):
            inj = masks.get(("droppath", bi)) if masks else None
            s1 = _droppath_scale(blk, B, x.device, inj[0] if inj is not None else None)
            s2 = _droppath_scale(blk, B, x.device, inj[1] if inj is not None else None)
            tok, sv = _block_forward(blk, tok, B, N, H, W, (s1, s2))
            blocks.append(sv)
            bi += 1
        norm = getattr(enc, f"norm{s + 1}")
        out = ops.layernorm(tok, norm.weight.detach(), norm.bias.detach(), norm.eps)
        tape["stages"].append(dict(P=P, Kp=Kp, y=y, tok_final=tok, blocks=blocks, H=H, W=W, Hin=Hin, Win=Win, Cin=Cin, k=k, s=st, p=p))
        outs.append((out, H, W))
        prev = out
    return outs, tape


def head_forward(head, stages, B, masks):
    """stages: four (tokens bf16 [B*N_i, C_i], H_i, W_i) -> logits fp32 [B, h1, w1, nc] pixel-major, and the tape."""
    (t1, h1, w1), (t2, h2, w2), (t3, h3, w3), (t4, h4, w4) = stages
    E, dev = head.embedding_dim, t1.device
    M = B * h1 * w1
    cat = torch.empty((B, h1, w1, 4 * E), dtype=BF16, device=dev)
    for slot, (mlp, t, h, w) in enumerate(((head.linear_c4, t4, h4, w4), (head.linear_c3, t3, h3, w3), (head.linear_c2, t2, h2, w2))):
        y = mlp.forward_tokens(t)
        ops.bilinear_nhwc(y, B, h, w, E, h1, w1, out=cat, ld_dst=4 * E, dst_coff=slot * E)
    head.linear_c1.forward_tokens(t1, out=cat.view(-1, 4 * E), ld_dst=4 * E, dst_coff=3 * E)
    conv, bn = head.linear_fuse.conv, head.linear_fuse.bn
    z = ops.linear(cat.view(M, 4 * E), head._packs.conv(conv.weight), None)                       # 1x1 conv, bias=False
    if head.training:
        mom = 0.1 if bn.momentum is None else bn.momentum
        y, stats = ops.bn_train_fwd(z, bn.weight.detach(), bn.bias.detach(), bn.eps, mom, bn.running_mean, bn.running_var)
        bn.num_batches_tracked += 1
        scale = masks.get("dropout2d") if masks else None
        if scale is None:
            pdrop = head.dropout.p
            scale = (torch.rand((B, E), device=dev) >= pdrop).float() / (1.0 - pdrop)
    else:
        # The reference's train_seg calls val_segformer() every 1000 iterations, which leaves the model in eval() and never
        # restores train() (train.py:232-236): from then on it trains with running-statistics BatchNorm, no Dropout2d and no
        # DropPath.  Same here: BN uses the running statistics (constants for the backward), scale = None.
        y, stats = ops.bn_eval_fwd(z, bn.weight.detach(), bn.bias.detach(), bn.eps, bn.running_mean, bn.running_var)
        scale = None
    yd = ops.channel_scale(y, scale, B, h1 * w1, E) if scale is not None else y
    logits = ops.linear(yd, head._packs.conv(head.linear_pred.weight), head.linear_pred.bias.detach(), act=ACT_NONE, out_dtype=F32)
    tape = dict(cat=cat, z=z, y=y, yd=yd, stats=stats, scale=scale, dims=((h1, w1), (h2, w2), (h3, w3), (h4, w4)),
                toks=(t1, t2, t3, t4), train_bn=head.training)
    return logits.view(B, h1, w1, head.num_classes), tape


# ------------------------------------------------------------------------------------------------ backward
def _block_backward(blk, sv, dx, B, N, g, pre):
    """dx: fp32 [B*N, C] running gradient of the residual stream (updated in place)."""
    at, ml = blk.attn, blk.mlp
    C = dx.shape[1]
    H, W = sv["H"], sv["W"]
    heads, D = at.num_heads, C // at.num_heads
    M = B * N
    dev = dx.device

    def branch_grad(scale):
        if scale is None:
            return ops.cast(dx, BF16)
        return ops.scale_cast_rows(dx, scale, N)

    # ---- Mix-FFN branch: x3 = x2 + s2 * fc2(gelu(dwconv(fc1(LN2(x2)))))
    dy = branch_grad(sv["s2"])
    dh2 = _lin_bwd(dy, sv["h2"], ml.fc2.weight, ml._packs, g, pre + "mlp.fc2.weight", pre + "mlp.fc2.bias")
    hid = dh2.shape[1]
    dw9c = torch.zeros((9, hid), dtype=F32, device=dev)
    dz = ops.dwconv3x3_gelu_bwd(sv["h1"], ml.dwconv._w(), ml.dwconv.dwconv.bias.detach(), dh2, B, H, W, dw9c,
                                g[pre + "mlp.dwconv.dwconv.bias"])
    g[pre + "mlp.dwconv.dwconv.weight"].add_(dw9c.t().reshape(hid, 1, 3, 3))
    dh1 = ops.dwconv3x3(dz, ml.dwconv._w(), None, B, H, W, flip=True)
    dn2 = _lin_bwd(dh1, sv["n2"], ml.fc1.weight, ml._packs, g, pre + "mlp.fc1.weight", pre + "mlp.fc1.bias")
    ops.layernorm_bwd(sv["x2"], dn2, C, 0, blk.norm2.weight.detach(), blk.norm2.eps, dx, C, 0, M, C,
                      dgamma=g[pre + "norm2.weight"], dbeta=g[pre + "norm2.bias"], accumulate=True)
    # ---- attention branch: x2 = x + s1 * proj(attn(LN1(x)))
    dy = branch_grad(sv["s1"])
    datt = _lin_bwd(dy, sv["att"], at.proj.weight, at._packs, g, pre + "attn.proj.weight", pre + "attn.proj.bias")
    Nk = sv["Nk"]
    dq, dkv32 = ops.sr_attention_bwd(sv["q"], sv["kv"], sv["att"], datt, sv["lse"], B, heads, N, Nk, D, at.scale)
    dkv = ops.cast(dkv32, BF16)
    dsrc = _lin_bwd(dkv, sv["src"], at.kv.weight, at._packs, g, pre + "attn.kv.weight", pre + "attn.kv.bias" if at.kv.bias is not None else None)
    if at.sr_ratio > 1:
        r = at.sr_ratio
        Mk = B * Nk
        dred = torch.empty((Mk, C), dtype=F32, device=dev)
        ops.layernorm_bwd(sv["red"], dsrc, C, 0, at.norm.weight.detach(), at.norm.eps, dred, C, 0, Mk, C,
                          dgamma=g[pre + "attn.norm.weight"], dbeta=g[pre + "attn.norm.bias"])
        dred16 = ops.cast(dred, BF16)
        K = C * r * r
        _wgrad_lin(g, dred16, sv["P_sr"], ldy=C, ldx=K, P=Mk, Cin=K, Cout=C, grad=g[pre + "attn.sr.weight"], s_co=K,
                   dbias=g[pre + "attn.sr.bias"])
        dP = _dgrad(dred16, _flat_pack(at._packs, at.sr.weight, K, "sr_flat"), lambda: _flat_t_pack(at._packs, at.sr.weight, K, "sr_flat_t"), K)
        extra = _patches_to_map(dP, B, sv["Hk"], sv["Wk"], C, r, H, W)
        extra = extra if extra.is_contiguous() else extra.contiguous()
    else:
        extra = dsrc
    dn1 = _lin_bwd(dq, sv["n1"], at.q.weight, at._packs, g, pre + "attn.q.weight", pre + "attn.q.bias" if at.q.bias is not None else None,
                   residual=extra)
    ops.layernorm_bwd(sv["x"], dn1, C, 0, blk.norm1.weight.detach(), blk.norm1.eps, dx, C, 0, M, C,
                      dgamma=g[pre + "norm1.weight"], dbeta=g[pre + "norm1.bias"], accumulate=True)


def encoder_backward(enc, tape, douts, g, prefix, want_input_grad):
    """douts: four bf16 [B*N_s, C_s] gradients of the stage outputs (None = no gradient).  Returns d image (fp32 NCHW) or None."""
    B = tape["B"]
    carry = None                                   # gradient flowing from stage s+1's patch embedding into stage s's output
    dimg = None
    for s in range(3, -1, -1):
        st = tape["stages"][s]
        H, W = st["H"], st["W"]
        N = H * W
        M = B * N
        norm = getattr(enc, f"norm{s + 1}")
        C = st["tok_final"].shape[1]
        dout = douts[s]
        if carry is not None:
            if dout is None:
                dout = carry
            else:
                ops.add_bf16(dout, C, 0, carry, C, 0, dout, C, 0, M, C)
        dx = torch.empty((M, C), dtype=F32, device=dout.device)
        ops.layernorm_bwd(st["tok_final"], dout, C, 0, norm.weight.detach(), norm.eps, dx, C, 0, M, C,
                          dgamma=g[f"{prefix}norm{s + 1}.weight"], dbeta=g[f"{prefix}norm{s + 1}.bias"])
        blocks = getattr(enc, f"block{s + 1}")
        for i in range(len(blocks) - 1, -1, -1):
            _block_backward(blocks[i], st["blocks"][i], dx, B, N, g, f"{prefix}block{s + 1}.{i}.")
        pe = getattr(enc, f"patch_embed{s + 1}")
        pp = f"{prefix}patch_embed{s + 1}."
        dy = torch.empty((M, C), dtype=F32, device=dx.device)
        ops.layernorm_bwd(st["y"], dx, C, 0, pe.norm.weight.detach(), pe.norm.eps, dy, C, 0, M, C,
                          dgamma=g[pp + "norm.weight"], dbeta=g[pp + "norm.bias"])
        dy16 = ops.cast(dy, BF16)
        Kp, k, Cin = st["Kp"], st["k"], st["Cin"]
        _wgrad_lin(g, dy16, st["P"], ldy=C, ldx=Kp, P=M, Cin=Kp, Cout=C, grad=g[pp + "proj.weight"], s_co=Cin * k * k,
                   ci_take=Cin * k * k, dbias=g[pp + "proj.bias"])
        if s > 0 or want_input_grad:
            dP = _dgrad(dy16, _flat_pack(pe._packs, pe.proj.weight, Kp, "pe_flat"), lambda: _flat_t_pack(pe._packs, pe.proj.weight, Kp, "pe_flat_t"), Kp)
            if s > 0:
                carry = ops.col2im(dP, B, st["Hin"], st["Win"], Cin, k, st["s"], st["p"], out_dtype=BF16).view(-1, Cin)
            else:
                d_nhwc = ops.col2im(dP, B, st["Hin"], st["Win"], 3, k, st["s"], st["p"], out_dtype=F32)
                dxn = ops.nhwc_to_nchw(d_nhwc, B, st["Hin"] * st["Win"], 3).view(B, 3, st["Hin"], st["Win"])
                dimg = ops.channel_affine_nchw(dxn, tape["in_scale"], None) if tape["in_scale"] is not None else dxn
        if STAGE_DONE_HOOK is not None:
            STAGE_DONE_HOOK(s)
    return dimg


def head_backward(head, tape, dlogits, B, g, prefix):
    """dlogits fp32 [B, h1, w1, nc] -> four bf16 gradients of the stage outputs."""
    E, nc, dev = head.embedding_dim, head.num_classes, dlogits.device
    (h1, w1), (h2, w2), (h3, w3), (h4, w4) = tape["dims"]
    M = B * h1 * w1
    dl = torch.zeros((M, 32), dtype=BF16, device=dev)                       # class gradients padded to 32 channels
    dl[:, :nc] = dlogits.reshape(M, nc)
    db = torch.zeros((32,), dtype=F32, device=dev)
    if g.frozen:
        ops.colsum(dl, 32, 0, M, 32, db)
    else:
        ops.wgrad_lin(dl, 32, 0, tape["yd"], E, 0, P=M, Cin=E, Cout=32, grad=g[prefix + "linear_pred.weight"], s_co=E, co_take=nc, dbias=db)
    g[prefix + "linear_pred.bias"].add_(db[:nc])

    def wt_pred(w):
        w2 = w.detach().reshape(nc, E).float()
        w2 = torch.cat([w2, w2.new_zeros(32 - nc, E)], 0)                   # [32, E]
        return w2.t().to(BF16).reshape(E, 1, 32).contiguous()
    dyd = ops.linear_tc(dl, head._packs.get(head.linear_pred.weight, wt_pred, "pred_t"), None)
    dy = ops.channel_scale(dyd, tape["scale"], B, h1 * w1, E) if tape["scale"] is not None else dyd
    bn, conv = head.linear_fuse.bn, head.linear_fuse.conv
    dz = ops.bn_train_bwd(tape["z"], tape["y"], dy, tape["stats"], bn.weight.detach(), g[prefix + "linear_fuse.bn.weight"],
                          g[prefix + "linear_fuse.bn.bias"], eval_mode=not tape["train_bn"])
    cat2 = tape["cat"].view(M, 4 * E)
    _wgrad_lin(g, dz, cat2, ldy=E, ldx=4 * E, P=M, Cin=4 * E, Cout=E, grad=g[prefix + "linear_fuse.conv.weight"], s_co=4 * E)
    dcat = _dgrad(dz, head._packs.conv(conv.weight), lambda: _t_pack(head._packs, conv.weight, "fuse_t"), 4 * E)   # [M, 4E]
    t1, t2, t3, t4 = tape["toks"]
    douts = [None] * 4
    for slot, (name, mlp, t, h, w, idx) in enumerate((("linear_c4", head.linear_c4, t4, h4, w4, 3), ("linear_c3", head.linear_c3, t3, h3, w3, 2),
                                                      ("linear_c2", head.linear_c2, t2, h2, w2, 1))):
        dyi = ops.bilinear_nhwc_bwd(dcat, 4 * E, slot * E, B, h1, w1, h, w, E).view(B * h * w, E)
        douts[idx] = _lin_bwd(dyi, t, mlp.proj.weight, mlp._packs, g, f"{prefix}{name}.proj.weight", f"{prefix}{name}.proj.bias")
    dy1 = dcat[:, 3 * E:].contiguous()                                     # linear_c1's slice (no resize)
    douts[0] = _lin_bwd(dy1, t1, head.linear_c1.proj.weight, head.linear_c1._packs, g, prefix + "linear_c1.proj.weight",
                        prefix + "linear_c1.proj.bias")
    return douts


# ------------------------------------------------------------------------------------------------ autograd nodes
class SegNetFn(torch.autograd.Function):
    """Network3's encoder + decode head as one autograd node: returns pixel-major fp32 logits [B, h, w, nc]."""

    @staticmethod
    def forward(ctx, net3, x, masks, names, *params):
        wetr = net3.denoise_net
        sc, sh = net3._input_affine(x.device)
        stages, etape = encoder_forward(wetr.encoder, x, sc, sh, masks)
        logits, htape = head_forward(wetr.decoder, stages, x.shape[0], masks)
        ctx.net3, ctx.etape, ctx.htape, ctx.names = net3, etape, htape, names
        ctx.want_x = x.requires_grad
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        net3 = ctx.net3
        wetr = net3.denoise_net
        dev = dlogits.device
        B = ctx.etape["B"]
        lookup = dict(net3.named_parameters())
        # A parameter owned by ddp.FlatParams (which pre-allocates its .grad as a view of one flat buffer and marks it with
        # _segmif_epoch) gets its gradient accumulated IN PLACE by the kernels (they all add into their output) and None
        # is returned for it: saves a zero-fill, an AccumulateGrad add and 2x the parameter bytes of traffic per tensor.
        # Everything else goes through autograd's own accumulation (so hooks, e.g. of torch's DistributedDataParallel,
        # still fire) out of one zero-filled scratch buffer (a single fill instead of one per tensor).
        g = _Grads()
        g.frozen = not any(ctx.needs_input_grad[4:])
        direct, pending = {}, []
        for i, n in enumerate(ctx.names):
            p = lookup[n]
            pg = p.grad
            if ctx.needs_input_grad[4 + i] and getattr(p, "_segmif_epoch", None) is not None and pg is not None \
                    and pg.dtype == F32 and pg.is_contiguous() and pg.device == dev and not torch.is_grad_enabled() \
                    and pg.shape == p.shape:
                g[n] = pg
                direct[n] = True
            else:
                pending.append((n, p))
        if pending:
            pad4 = lambda k: (k + 3) & ~3
            flat = torch.zeros((sum(pad4(p.numel()) for _, p in pending),), dtype=F32, device=dev)
            off = 0
            for n, p in pending:
                g[n] = flat[off:off + p.numel()].view(p.shape)
                off += pad4(p.numel())
        douts = head_backward(wetr.decoder, ctx.htape, dlogits.float().contiguous(), B, g, "denoise_net.decoder.")
        dimg = encoder_backward(wetr.encoder, ctx.etape, douts, g, "denoise_net.encoder.", ctx.want_x)
        ctx.etape = ctx.htape = None
        if WGRAD_SIDE is not None and pending:      # tensors handed back to autograd are consumed on this stream right away
            WGRAD_SIDE.join()
        grads = tuple(g[n] if (ctx.needs_input_grad[4 + i] and n not in direct) else None for i, n in enumerate(ctx.names))
        return (None, dimg if ctx.want_x else None, None, None) + grads


class CeFn(torch.autograd.Function):
    """CrossEntropyLoss(ignore_index) over bilinearly upsampled pixel-major logits (core/model_fusion.py:1095-1096)."""

    @staticmethod
    def forward(ctx, logits_pm, labels, ignore_index):
        B, h, w, nc = logits_pm.shape
        lg = logits_pm.float().contiguous()
        labels = labels.long().contiguous()
        loss, cnt = ops.upsample_ce(lg, B, h, w, nc, labels, ignore_index, return_count=True)
        ctx.save_for_backward(lg, labels, cnt)
        ctx.ignore_index = ignore_index
        return loss

    @staticmethod
    def backward(ctx, gout):
        lg, labels, cnt = ctx.saved_tensors
        B, h, w, nc = lg.shape
        return ops.upsample_ce_bwd(lg, B, h, w, nc, labels, gout, cnt, ctx.ignore_index), None, None


def logits_with_grad(net3, x, masks=None):
    named = [(k, p) for k, p in net3.named_parameters() if not k.endswith("classifier.weight")]
    names = tuple(k for k, _ in named)
    return SegNetFn.apply(net3, x, masks, names, *[p for _, p in named])
