"""Strict-precision (fp32-parity) inference path: the same modules, parameters and call graph as the default bf16 path,
with fp32 activations in HBM and every dense contraction evaluated by `segmif_split_gemm_fwd` on 3-way bf16-split
operands (six partial products on tcgen05, fp32-grade result) -- see csrc/gemm_split_tc.cu and csrc/strict_ops.cu.

north_star: "outputs match the reference PyTorch path on identical inputs within 1e-3 rel fp32 (bit-exact for argmax
segmentation labels)".  The default path (bf16 operands) sits at the bf16 floor (~8e-3 on the fused image, ~99.5 % equal
labels); this path meets the contract and is what `segmif_b200.set_precision("strict")` / SEGMIF_PRECISION=strict
select for `FusionSegPipeline`, `Network3.forward/_loss/predict_labels`, `MixVisionTransformer.forward_*`,
`Fusion_Network3_ac.forward` and the module-level forwards of DRDB / FeatureFusionModule (inference only).

Reference call sites mirrored here (SegMiF repository):
  core/mix_transformer.py:94-115 (Attention), :46-53 (Mlp), :151-155 (Block), :192-198 (OverlapPatchEmbed),
  :312-375 (forward_features / forward_fusion); core/segformer_head.py:59-82; core/model_fusion.py:134-157 (DRDB),
  :350-361 + :263-328 (CrossPath / CrossAttention[2]), :1047-1067 (Fusion_Network3_ac.forward), :1081-1097 (Network3).
"""
import ctypes
import os

import torch

from . import _lib, ops
from ._lib import ACT_NONE, ACT_PRELU, ACT_RELU

_PRECISION = [os.environ.get("SEGMIF_PRECISION", "bf16")]
NTERMS = int(os.environ.get("SEGMIF_STRICT_TERMS", "6"))      # 6 = fp32-grade; 3 = two planes (diagnostic)


def set_precision(mode):
    """'bf16' (default: bf16 operands, fp32 accumulation) or 'strict' (fp32 parity, ~5x the tensor-core work)."""
    if mode not in ("bf16", "strict"):
        raise ValueError("segmif_b200.set_precision: mode must be 'bf16' or 'strict'")
    _PRECISION[0] = mode


def get_precision():
    return _PRECISION[0]


def is_strict():
    return _PRECISION[0] == "strict"


class precision:
    """Context manager: `with segmif_b200.precision('strict'): ...`"""

    def __init__(self, mode):
        self.mode = mode

    def __enter__(self):
        self.prev = get_precision()
        set_precision(self.mode)

    def __exit__(self, *exc):
        set_precision(self.prev)


# ------------------------------------------------------------------------------------------------ low-level wrappers
def _p(t, elem_off=0):
    return ctypes.c_void_p(t.data_ptr() + elem_off * t.element_size()) if t is not None else None


class Planes:
    """The three bf16 planes of an fp32 matrix [M, ld]: storage [3, M, ld]."""
    __slots__ = ("t", "M", "ld")

    def __init__(self, M, ld, device):
        self.t = torch.empty((3, M, ld), dtype=torch.bfloat16, device=device)
        self.M, self.ld = M, ld

    @property
    def plane_stride(self):
        return self.M * self.ld


def split(x, C=None, *, ld_x=None, coff_x=0, relu=False, out_f32=None, ld_y=None, coff_y=0, planes=None, coff_p=0):
    """x fp32 [rows, ld_x] -> Planes (allocated [rows, C] unless given); optional ReLU, optional fp32 copy of the result."""
    ld_x = x.shape[-1] if ld_x is None else ld_x
    rows = x.numel() // ld_x
    C = ld_x if C is None else C
    st = ops._prep(x, out_f32, planes.t if planes is not None else None)
    if planes is None:
        planes = Planes(rows, C, x.device)
    if out_f32 is not None and ld_y is None:
        ld_y = out_f32.shape[-1]
    _lib.call("segmif_split3", _p(x), ld_x, coff_x, rows, C, 1 if relu else 0, _p(out_f32), ld_y or 0, coff_y,
              _p(planes.t), planes.ld, coff_p, planes.plane_stride, st)
    return planes


def pack_rows(w2d):
    """fp32 [rows, K] -> bf16 [rows, 3, Kp] (Kp = K rounded up to 64, zero padded): the W operand of split_gemm."""
    w2d = w2d.detach().float().contiguous()
    rows, K = w2d.shape
    Kp = (K + 63) // 64 * 64
    out = torch.zeros((rows, 3, Kp), dtype=torch.bfloat16, device=w2d.device)
    st = ops._prep(w2d, out)
    _lib.call("segmif_split3", _p(w2d), K, 0, rows, K, 0, None, 0, 0, _p(out), 3 * Kp, 0, Kp, st)
    return out


def pack_linear(cache, weight):
    return cache.get(weight, lambda w: pack_rows(w.reshape(w.shape[0], -1)), "s3lin")


def pack_conv_taps(cache, weight):
    """[Cout, Cin, KH, KW] -> [Cout * KH*KW, 3, Kp]: one K block per tap (patch mode)."""
    return cache.get(weight, lambda w: pack_rows(w.detach().permute(0, 2, 3, 1).reshape(-1, w.shape[1])), "s3taps")


def pack_conv_im2col(cache, weight):
    """[Cout, Cin, KH, KW] -> [Cout, 3, Kp] over K = KH*KW*Cin, column (ky*KW + kx)*Cin + c (row mode on im2col rows)."""
    return cache.get(weight, lambda w: pack_rows(w.detach().permute(0, 2, 3, 1).reshape(w.shape[0], -1)), "s3col")


def gemm(A, K, w, N, *, a_coff=0, row0=0, rows=None, bias=None, act=ACT_NONE, alpha=None, residual=None, ld_res=None,
         res_coff=0, dst=None, ld_dst=None, dst_coff=0, want_f32=True, dst_planes=None, dp_coff=0, want_planes=False,
         patch=None, nterms=None):
    """dst[m, n] = residual + act(bias + sum A[m(+tap), k] W[n, (tap,) k]).  `A`: Planes; `patch` = (B, H, W, dil) selects
    the 3x3 'same' convolution over A viewed as [B, H, W, ld]; row0/rows select a row range (all row-indexed tensors)."""
    dev = A.t.device
    M = (A.M - row0) if rows is None else rows
    if patch is not None:
        B, H, W, dil = patch
        M = B * H * W
    st = ops._prep(A.t, w, bias, alpha, residual, dst, dst_planes.t if dst_planes is not None else None)
    if dst is None and want_f32:
        ld_dst = N if ld_dst is None else ld_dst
        dst = torch.empty((M, ld_dst), dtype=torch.float32, device=dev)
    elif dst is not None and ld_dst is None:
        ld_dst = dst.shape[-1]
    if dst_planes is None and want_planes:
        dst_planes = Planes(M, N, dev)
    p = _lib.SplitGemmParams()
    p.a_planes = A.t.data_ptr() + 2 * row0 * A.ld
    p.a_plane_stride = A.plane_stride
    p.w_planes = w.data_ptr()
    p.bias = bias.data_ptr() if bias is not None else None
    p.prelu_alpha = alpha.data_ptr() if alpha is not None else None
    if residual is not None:
        ld_res = residual.shape[-1] if ld_res is None else ld_res
        p.residual = residual.data_ptr() + 4 * row0 * ld_res
    p.ld_res, p.res_coff = ld_res or 0, res_coff
    if dst is not None:
        p.dst = dst.data_ptr() + 4 * row0 * ld_dst
    p.ld_dst, p.dst_coff = ld_dst or 0, dst_coff
    if dst_planes is not None:
        p.dst_planes = dst_planes.t.data_ptr() + 2 * row0 * dst_planes.ld
        p.dst_plane_stride = dst_planes.plane_stride
        p.ld_dp, p.dp_coff = dst_planes.ld, dp_coff
    p.M, p.N, p.K, p.ld_a, p.a_coff = M, N, K, A.ld, a_coff
    p.act = act
    p.nterms = NTERMS if nterms is None else nterms
    if patch is not None:
        p.B, p.H, p.W, p.ntaps = B, H, W, 9
        for t in range(9):
            p.tap_dx[t] = (t % 3 - 1) * dil
            p.tap_dy[t] = (t // 3 - 1) * dil
    _lib.call("segmif_split_gemm_fwd", ctypes.byref(p), st)
    return dst, dst_planes


def im2col_split(x, B, H, W, C, k, stride, pad):
    """x fp32 [B, H, W, C] -> Planes [B*Ho*Wo, k*k*C]."""
    st = ops._prep(x)
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    P = Planes(B * Ho * Wo, k * k * C, x.device)
    _lib.call("segmif_im2col_split3", _p(x), C, 0, B, H, W, C, k, stride, pad, _p(P.t), P.plane_stride, st)
    return P, Ho, Wo


def attention_f32(q, kv, B, heads, N, Nk, D, scale):
    st = ops._prep(q, kv)
    C = heads * D
    out = torch.empty((B * N, C), dtype=torch.float32, device=q.device)
    _lib.call("segmif_sr_attention_f32_fwd", _p(q), C, _p(kv), _p(kv, C), 2 * C, _p(out), C, B, heads, N, Nk, D,
              float(scale), st)
    return out


def dwconv_f32(x, w9c, bias, B, H, W, gelu=True):
    st = ops._prep(x, w9c, bias)
    y = torch.empty_like(x)
    _lib.call("segmif_dwconv3x3_f32_fwd", _p(x), _p(w9c), _p(bias), _p(y), B, H, W, x.shape[-1], 1 if gelu else 0, st)
    return y


def conv3x3_in1_f32(img_nchw, w9c, bias, alpha, Cout):
    st = ops._prep(w9c, bias, alpha)
    B, C, H, W = img_nchw.shape
    if not img_nchw.is_cuda or img_nchw.dtype != torch.float32 or img_nchw.stride(3) != 1 or img_nchw.stride(2) != W:
        raise RuntimeError("segmif_b200.strict.conv3x3_in1_f32: need a CUDA fp32 NCHW tensor with dense rows")
    out = torch.empty((B * H * W, Cout), dtype=torch.float32, device=img_nchw.device)
    _lib.call("segmif_conv3x3_in1_f32_fwd", _p(img_nchw), img_nchw.stride(0), _p(w9c), _p(bias), _p(alpha), _p(out), Cout, 0,
              B, H, W, Cout, st)
    return out


def conv3x3_out1_f32(src, w9c, bias, alpha, B, H, W, Cin):
    st = ops._prep(src, w9c, bias, alpha)
    out = torch.empty((B, 1, H, W), dtype=torch.float32, device=src.device)
    _lib.call("segmif_conv3x3_out1_f32_fwd", _p(src), src.shape[-1], _p(w9c), _p(bias), _p(alpha), _p(out), B, H, W, Cin, st)
    return out


def _ln(x, norm, eps=None):
    return ops.layernorm(x, norm.weight.detach(), norm.bias.detach(), norm.eps if eps is None else eps, out_dtype=torch.float32)


def _b(lin):
    return lin.bias.detach() if lin.bias is not None else None


# ------------------------------------------------------------------------------------------------ MiT encoder
def _attention(att, n1, x_res, B, N, C, H, W):
    """core/mix_transformer.py:94-115 on fp32 tokens; returns x_res + proj(attn(...)) (Block's first residual add)."""
    M = B * N
    D = C // att.num_heads
    P1 = split(n1.view(M, C))
    q, _ = gemm(P1, C, pack_linear(att._packs, att.q.weight), C, bias=_b(att.q))
    if att.sr_ratio > 1:
        r = att.sr_ratio
        col, Ho, Wo = im2col_split(n1, B, H, W, C, r, r, 0)
        red, _ = gemm(col, r * r * C, pack_conv_im2col(att._packs, att.sr.weight), C, bias=att.sr.bias.detach())
        src = _ln(red, att.norm)
        Ps, Nk = split(src), Ho * Wo
    else:
        Ps, Nk = P1, N
    kv, _ = gemm(Ps, C, pack_linear(att._packs, att.kv.weight), 2 * C, bias=_b(att.kv))
    o = attention_f32(q, kv, B, att.num_heads, N, Nk, D, att.scale)
    y, _ = gemm(split(o), C, pack_linear(att._packs, att.proj.weight), C, bias=att.proj.bias.detach(), residual=x_res)
    return y


def _mlp(mlp, n2, x_res, B, N, C, H, W):
    """core/mix_transformer.py:46-53: fc1 -> depthwise 3x3 -> GELU(erf) -> fc2 (+ residual)."""
    M = B * N
    hid = mlp.fc1.out_features
    h, _ = gemm(split(n2.view(M, C)), C, pack_linear(mlp._packs, mlp.fc1.weight), hid, bias=mlp.fc1.bias.detach())
    g = dwconv_f32(h, mlp.dwconv._w(), mlp.dwconv.dwconv.bias.detach(), B, H, W, gelu=True)
    y, _ = gemm(split(g), hid, pack_linear(mlp._packs, mlp.fc2.weight), C, bias=mlp.fc2.bias.detach(), residual=x_res)
    return y


def block(blk, x, H, W):
    """core/mix_transformer.py:151-155 (eval: DropPath is the identity)."""
    B, N, C = x.shape
    x2 = x.reshape(B * N, C)
    x2 = _attention(blk.attn, _ln(x2, blk.norm1).view(B, N, C), x2, B, N, C, H, W)
    x2 = _mlp(blk.mlp, _ln(x2, blk.norm2).view(B, N, C), x2, B, N, C, H, W)
    return x2.view(B, N, C)


def patch_embed_tokens(pe, tok, B, H, W):
    """core/mix_transformer.py:192-198 for stages 2-4: k3 s2 p1 conv (im2col + split GEMM) -> LayerNorm."""
    k, s = pe.patch_size[0], pe.stride
    Cin, Cout = tok.shape[-1], pe.proj.out_channels
    col, Ho, Wo = im2col_split(tok, B, H, W, Cin, k, s, k // 2)
    y, _ = gemm(col, k * k * Cin, pack_conv_im2col(pe._packs, pe.proj.weight), Cout, bias=pe.proj.bias.detach())
    return _ln(y, pe.norm).view(B, Ho * Wo, Cout), Ho, Wo


def encoder_stages(enc, x, in_scale=None, in_shift=None, n_stages=4):
    """Per stage (tokens fp32 [B, N, C] after the stage LayerNorm, H, W) -- core/mix_transformer.py:312-348."""
    if enc.training:
        raise NotImplementedError("segmif_b200: the strict-precision path is inference only (call .eval())")
    B = x.shape[0]
    outs, tok, H, W = [], None, None, None
    for s in range(n_stages):
        pe = getattr(enc, f"patch_embed{s + 1}")
        if s == 0:
            tok, H, W = pe.forward_image(x.float().contiguous(), in_scale, in_shift)     # fp32 direct conv + LN
        else:
            tok, H, W = patch_embed_tokens(pe, tok, B, H, W)
        for blk in getattr(enc, f"block{s + 1}"):
            tok = block(blk, tok, H, W)
        tok = _ln(tok, getattr(enc, f"norm{s + 1}"))
        outs.append((tok, H, W))
    return outs


def forward_features(enc, x):
    B = x.shape[0]
    return [ops.nhwc_to_nchw(t, B, H * W, t.shape[-1]).view(B, t.shape[-1], H, W) for t, H, W in encoder_stages(enc, x)]


def forward_fusion(enc, x):
    """core/mix_transformer.py:358-375: stage-1/2 maps bilinearly upsampled to the input size, fp32 NCHW-shaped
    channels_last views."""
    B, _, H, W = x.shape
    outs = []
    for t, h, w in encoder_stages(enc, x, n_stages=2):
        C = t.shape[-1]
        up = ops.bilinear_nhwc(t, B, h, w, C, H, W, out_dtype=torch.float32)
        outs.append(up.permute(0, 3, 1, 2))
    return outs[0], outs[1]


# ------------------------------------------------------------------------------------------------ decode head
def head_logits(head, stages):
    """core/segformer_head.py:59-82 (eval) -> fp32 logits [B, h1, w1, nc] pixel-major."""
    if head.training:
        raise NotImplementedError("segmif_b200: the strict-precision path is inference only (call .eval())")
    (t1, h1, w1), (t2, h2, w2), (t3, h3, w3), (t4, h4, w4) = stages
    B, E = t1.shape[0], head.embedding_dim
    cat = torch.empty((B, h1, w1, 4 * E), dtype=torch.float32, device=t1.device)
    for slot, (mlp, t, h, w) in enumerate(((head.linear_c4, t4, h4, w4), (head.linear_c3, t3, h3, w3),
                                           (head.linear_c2, t2, h2, w2))):
        C = t.shape[-1]
        y, _ = gemm(split(t.reshape(-1, C)), C, pack_linear(mlp._packs, mlp.proj.weight), E, bias=mlp.proj.bias.detach())
        ops.bilinear_nhwc(y, B, h, w, E, h1, w1, out=cat, ld_dst=4 * E, dst_coff=slot * E)
    C1 = t1.shape[-1]
    gemm(split(t1.reshape(-1, C1)), C1, pack_linear(head.linear_c1._packs, head.linear_c1.proj.weight), E,
         bias=head.linear_c1.proj.bias.detach(), dst=cat.view(-1, 4 * E), ld_dst=4 * E, dst_coff=3 * E)
    bn, conv = head.linear_fuse.bn, head.linear_fuse.conv

    def fold(w, g, b, mean, var):
        scale = g.double() / torch.sqrt(var.double() + bn.eps)
        wf = (w.double().reshape(w.shape[0], -1) * scale[:, None]).float()
        return pack_rows(wf), (b.double() - mean.double() * scale).float().contiguous()
    wf, bf = head._packs.get_multi([conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var], fold, "s3fuse")
    _, fused = gemm(split(cat.view(-1, 4 * E)), 4 * E, wf, E, bias=bf, act=ACT_RELU, want_f32=False, want_planes=True)
    nc = head.num_classes
    logits, _ = gemm(fused, E, pack_linear(head._packs, head.linear_pred.weight), nc, bias=head.linear_pred.bias.detach())
    return logits.view(B, h1, w1, nc)


def wetr_logits(wetr, x, in_scale=None, in_shift=None):
    return head_logits(wetr.decoder, encoder_stages(wetr.encoder, x, in_scale, in_shift))


# ------------------------------------------------------------------------------------------------ fusion network
def drdb(d, g, x_f32, B, H, W, want_planes=True):
    """core/model_fusion.py:134-157.  g: Planes [M, 224] with the block input in channels 0..in_ch; x_f32 [M, in_ch] is
    the same input in fp32 (the residual).  Returns (out fp32 [M, in_ch], its Planes)."""
    cin = d.in_ch
    for i in range(1, 6):
        cv = getattr(d, f"Dcov{i}")
        gemm(g, cin, pack_conv_taps(d._packs, cv.weight), d.growth, bias=cv.bias.detach(), act=ACT_RELU,
             patch=(B, H, W, 2), want_f32=False, dst_planes=g, dp_coff=cin)
        cin += d.growth
    return gemm(g, cin, pack_linear(d._packs, d.conv.weight), d.in_ch, bias=d.conv.bias.detach(), act=ACT_RELU,
                residual=x_f32, want_planes=want_planes)


def cross_path(cp, x1, p1, x2, p2, p3_f32, p3, B, HW):
    """core/model_fusion.py:350-361 with CrossAttention (:263-288) and CrossAttention2 (:303-328).
    x_i fp32 [B*HW, 64] with Planes p_i; p3_f32 / p3 = relu(channel_proj3(seg)) [B*HW, 128] (y3 | u3).
    Returns the two LayerNorm outputs fp32 [B*HW, 64]."""
    dev = x1.device
    M = B * HW
    P = []
    for x, pl, proj in ((x1, p1, cp.channel_proj1), (x2, p2, cp.channel_proj2)):
        P.append(gemm(pl, 64, pack_linear(cp._packs, proj.weight), 128, bias=proj.bias.detach(), act=ACT_RELU, want_planes=True))
    # contexts: k^T v = Wk (P^T P) Wv^T (the kv Linears have no bias), Gram matrices accumulated in fp64
    nchunk = max(1, min(296 // max(B, 1), (HW + 511) // 512))
    partials = torch.empty((3, B, nchunk, 64, 64), dtype=torch.float64, device=dev)
    st = ops._prep(x1)
    for s, (t, coff) in enumerate(((P[0][0], 0), (P[1][0], 0), (p3_f32, 64))):
        _lib.call("segmif_gram64_f64", _p(t), 128, coff, B, HW, 0, _p(partials[s]), nchunk, st)
    wkv = cp._packs.get_multi([cp.cross_attn2.kv1.weight, cp.cross_attn2.kv2.weight, cp.cross_attn.kv3.weight],
                              lambda a, b, c: torch.stack([a.detach().float(), b.detach().float(), c.detach().float()]).contiguous(), "s3wkv")
    wend = cp._packs.get_multi([cp.end_proj1.weight, cp.end_proj2.weight],
                               lambda a, b: torch.stack([a.detach().float(), b.detach().float()]).contiguous(), "s3wend")
    folded = torch.empty((B, 4, 64, 64), dtype=torch.float32, device=dev)
    ctx = torch.empty((B, 3, 8, 8, 8), dtype=torch.float32, device=dev)
    _lib.call("segmif_ffm_ctx_f64_fwd", _p(partials), nchunk, _p(wkv), _p(wend), _p(folded), _p(ctx), B, st)
    wf = pack_rows(folded.view(B * 4 * 64, 64)).view(B, 4, 64, 3, 64)
    outs = []
    for i, (x, end, norm) in enumerate(((x1, cp.end_proj1, cp.norm1), (x2, cp.end_proj2, cp.norm2))):
        pre = torch.empty((M, 64), dtype=torch.float32, device=dev)
        for b in range(B):
            # x_i + b_end + y3 Mz_i^T, then + u_i Mv_i^T
            gemm(p3, 64, wf[b, 2 * i], 64, a_coff=0, row0=b * HW, rows=HW, bias=end.bias.detach(), residual=x, dst=pre)
            gemm(P[i][1], 64, wf[b, 2 * i + 1], 64, a_coff=64, row0=b * HW, rows=HW, residual=pre, dst=pre)
        outs.append(_ln(pre, norm))
    return outs[0], outs[1], ctx


def ffm_lowres(ffm, x1, p1, x2, p2, seg, pre_conv, B, H, W):
    """FeatureFusionModule on (x1, x2, pre_conv(upsample(seg))): the 1x1 conv and channel_proj3 run at the encoder's
    resolution (they commute with the bilinear resize: all linear, interpolation weights sum to one)."""
    cp = ffm.cross
    tok, h, w = seg
    C3 = tok.shape[-1]
    _, t = gemm(split(tok.reshape(-1, C3)), C3, pack_conv_im2col(cp._packs, pre_conv.weight), 64, bias=pre_conv.bias.detach(),
                want_f32=False, want_planes=True)
    q, _ = gemm(t, 64, pack_linear(cp._packs, cp.channel_proj3.weight), 128, bias=cp.channel_proj3.bias.detach())
    up = ops.bilinear_nhwc(q, B, h, w, 128, H, W, out_dtype=torch.float32).view(-1, 128)
    p3 = split(up, relu=True, out_f32=up)
    return cross_path(cp, x1, p1, x2, p2, up, p3, B, H * W)


def ffm_full(ffm, x1, p1, x2, p2, seg_f32, pre_conv, B, H, W):
    """Reference order: pre_conv (1x1) at full resolution on the given feature map [B*HW, C3] fp32, then channel_proj3."""
    cp = ffm.cross
    C3 = seg_f32.shape[-1]
    if pre_conv is not None:
        _, t = gemm(split(seg_f32), C3, pack_conv_im2col(cp._packs, pre_conv.weight), 64, bias=pre_conv.bias.detach(),
                    want_f32=False, want_planes=True)
    else:
        t = split(seg_f32)
    up, p3 = gemm(t, 64, pack_linear(cp._packs, cp.channel_proj3.weight), 128, bias=cp.channel_proj3.bias.detach(),
                  act=ACT_RELU, want_planes=True)
    return cross_path(cp, x1, p1, x2, p2, up, p3, B, H * W)


def fusion_network(fus, ir, vis, seg1, seg2):
    """core/model_fusion.py:1047-1067.  seg_i = ('lowres', tokens fp32 [B, h*w, C], h, w) or ('full', fp32 [B*HW, C])."""
    if fus.training:
        raise NotImplementedError("segmif_b200: the strict-precision path is inference only (call .eval())")
    B, _, H, W = ir.shape
    M = B * H * W
    dev = ir.device
    alpha = fus.relu.weight.detach()
    ir, vis = ir.float(), vis.float()
    G = 224
    g1, g2 = Planes(M, G, dev), Planes(M, G, dev)
    x1 = conv3x3_in1_f32(ir, fus._packs.taps_f32(fus.conv1_ir.weight), fus.conv1_ir.bias.detach(), alpha, 64)
    x2 = conv3x3_in1_f32(vis, fus._packs.taps_f32(fus.conv1_vis.weight), fus.conv1_vis.bias.detach(), alpha, 64)
    split(x1, planes=g1)
    split(x2, planes=g2)
    x1, p1 = drdb(fus.DRDB1, g1, x1, B, H, W)
    x2, p2 = drdb(fus.DRDB2, g2, x2, B, H, W)

    def ffm(x1, p1, x2, p2, seg, pre_conv):
        if seg[0] == "lowres":
            return ffm_lowres(fus.ffm, x1, p1, x2, p2, seg[1:], pre_conv, B, H, W)
        return ffm_full(fus.ffm, x1, p1, x2, p2, seg[1], pre_conv, B, H, W)
    o1, o2, _ = ffm(x1, p1, x2, p2, seg1, fus.conv3)
    split(o1, planes=g1)
    split(o2, planes=g2)
    x1, p1 = drdb(fus.DRDB3, g1, o1, B, H, W)
    x2, p2 = drdb(fus.DRDB4, g2, o2, B, H, W)
    o1, o2, _ = ffm(x1, p1, x2, p2, seg2, fus.conv4)
    cat = Planes(M, 128, dev)
    split(o1, planes=cat, coff_p=0)
    split(o2, planes=cat, coff_p=64)
    _, f = gemm(cat, 128, pack_conv_taps(fus._packs, fus.conv2.weight), 64, bias=fus.conv2.bias.detach(), act=ACT_PRELU,
                alpha=alpha, patch=(B, H, W, 1), want_f32=False, want_planes=True)
    f2, _ = gemm(f, 64, pack_conv_taps(fus._packs, fus.conv21.weight), 32, bias=fus.conv21.bias.detach(), act=ACT_PRELU,
                 alpha=alpha, patch=(B, H, W, 1))
    return conv3x3_out1_f32(f2, fus._packs.taps_f32(fus.conv22.weight), fus.conv22.bias.detach(), alpha, B, H, W, 32)


def nchw_to_rows_f32(x):
    """Logical NCHW tensor -> fp32 pixel-major rows [B*HW, C] (zero-copy for channels_last fp32 views)."""
    B, C, H, W = x.shape
    if x.dtype == torch.float32 and x.permute(0, 2, 3, 1).is_contiguous():
        return x.permute(0, 2, 3, 1).reshape(B * H * W, C)
    return ops.nchw_to_nhwc(x.float().contiguous(), out_dtype=torch.float32).view(B * H * W, C)
