"""MiT (Mix Transformer) encoder with the reference's class / attribute / state_dict surface
(core/mix_transformer.py of SegMiF) on top of the segmif_b200 sm_100a kernels.

Data flow inside a stage (no torch compute ops; `ops.*` are C-ABI kernel launches):
  residual stream  x      fp32 [B, N, C]   (kept fp32 so bf16 rounding never accumulates)
  LayerNorm                fp32 -> bf16     (ops.layernorm)
  q / kv / proj / fc1 / fc2 / sr / patch_embed2-4   bf16 tensor-core implicit GEMMs (ops.conv / ops.linear),
                           bias + residual fused in the epilogue
  attention core           flash-style kernel, scores never materialised (ops.sr_attention)
  DWConv + GELU            one fused stencil kernel on pixel-major tokens (ops.dwconv3x3_gelu)
The nn.Linear / nn.Conv2d / nn.LayerNorm children only hold parameters (fp32 masters, reference key
names); bf16 packed copies are cached per parameter version.
"""
import math
from functools import partial

import torch
import torch.nn as nn

from .. import ops
from .. import strict as _strict
from ..packing import PackCache
from ..ops import ACT_NONE


def to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


class DropPath(nn.Module):
    """Per-sample stochastic depth (timm semantics).  Block applies it itself (a per-sample scale fused into the residual
    add, Block._droppath_scale / core/seg_train.py); this module only carries drop_prob and is the identity in eval."""

    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        raise NotImplementedError("segmif_b200: train-mode DropPath is applied inside Block (fused residual scale), not as a module call")


def _reference_init(m):
    # core/mix_transformer.py:31-44 (identical rule in every class of the reference)
    if isinstance(m, nn.Linear):
        nn.init.trunc_normal_(m.weight, std=.02)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
    elif isinstance(m, nn.LayerNorm):
        nn.init.constant_(m.bias, 0)
        nn.init.constant_(m.weight, 1.0)
    elif isinstance(m, nn.Conv2d):
        fan_out = m.kernel_size[0] * m.kernel_size[1] * m.out_channels // m.groups
        m.weight.data.normal_(0, math.sqrt(2.0 / fan_out))
        if m.bias is not None:
            m.bias.data.zero_()


def _as_tokens_bf16(x):
    """Public entry points accept fp32 or bf16 [B, N, C]; kernels take bf16 operands."""
    if x.dtype == torch.bfloat16:
        return x.contiguous()
    return ops.nchw_to_nhwc(x.contiguous().view(1, 1, -1), out_dtype=torch.bfloat16).view(x.shape)


class DWConv(nn.Module):
    """core/mix_transformer.py:376-387; the GELU that follows it in Mlp is fused into the same kernel."""

    def __init__(self, dim=768):
        super().__init__()
        self.dwconv = nn.Conv2d(dim, dim, 3, 1, 1, bias=True, groups=dim)
        self._packs = PackCache()

    def _w(self):
        return self._packs.get(self.dwconv.weight, lambda w: w.detach().reshape(w.shape[0], 9).t().contiguous().float())

    def forward_gelu(self, x, H, W):
        B = x.shape[0]
        return ops.dwconv3x3_gelu(x, self._w(), self.dwconv.bias.detach(), B, H, W)

    def forward(self, x, H, W):
        """core/mix_transformer.py:381-387: tokens [B, N, C] -> depthwise 3x3 (pad 1, bias) -> tokens, no activation."""
        B, N, C = x.shape
        if x.dtype == torch.float32:
            return _strict.dwconv_f32(x.contiguous().view(B * N, C), self._w(), self.dwconv.bias.detach(), B, H, W, gelu=False).view(B, N, C)
        return ops.dwconv3x3(_as_tokens_bf16(x).view(B * N, C), self._w(), self.dwconv.bias.detach(), B, H, W).view(B, N, C)


class Mlp(nn.Module):
    """Mix-FFN: fc1 -> depthwise 3x3 -> GELU -> fc2 (core/mix_transformer.py:18-53)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.dwconv = DWConv(hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)
        self._packs = PackCache()
        self.apply(_reference_init)

    def _forward(self, x_bf16, H, W, residual=None, out_dtype=torch.float32):
        B, N, C = x_bf16.shape
        h = ops.linear(x_bf16, self._packs.linear(self.fc1.weight), self.fc1.bias.detach())
        h = self.dwconv.forward_gelu(h.view(B, N, -1), H, W)
        y = ops.linear(h, self._packs.linear(self.fc2.weight), self.fc2.bias.detach(), residual=residual,
                       out_dtype=out_dtype)
        return y.view(B, N, -1)

    def forward(self, x, H, W):
        return self._forward(_as_tokens_bf16(x), H, W, out_dtype=x.dtype)


class Attention(nn.Module):
    """Spatial-reduction self attention (core/mix_transformer.py:56-115)."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0., sr_ratio=1):
        super().__init__()
        assert dim % num_heads == 0, f"dim {dim} should be divided by num_heads {num_heads}."
        self.dim = dim
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = qk_scale or head_dim ** -0.5
        self.q = nn.Linear(dim, dim, bias=qkv_bias)
        self.kv = nn.Linear(dim, dim * 2, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self.sr_ratio = sr_ratio
        if sr_ratio > 1:
            self.sr = nn.Conv2d(dim, dim, kernel_size=sr_ratio, stride=sr_ratio)
            self.norm = nn.LayerNorm(dim)
        self._packs = PackCache()
        self.apply(_reference_init)

    @staticmethod
    def _b(lin):
        return lin.bias.detach() if lin.bias is not None else None

    def _forward(self, x_bf16, H, W, residual=None, out_dtype=torch.float32):
        B, N, C = x_bf16.shape
        D = C // self.num_heads
        q = ops.linear(x_bf16, self._packs.linear(self.q.weight), self._b(self.q))
        if self.sr_ratio > 1:
            r = self.sr_ratio
            red = ops.conv(x_bf16, self._packs.conv(self.sr.weight), self.sr.bias.detach(), B=B, H=H, W=W, Cin=C,
                           KH=r, KW=r, stride=r, pad=0, Cout=C, out_dtype=torch.float32)
            src = ops.layernorm(red, self.norm.weight.detach(), self.norm.bias.detach(), self.norm.eps)
            Nk = red.shape[0] // B
        else:
            src, Nk = x_bf16, N
        kv = ops.linear(src, self._packs.linear(self.kv.weight), self._b(self.kv))
        att = ops.sr_attention(q, kv, B, self.num_heads, N, Nk, D, self.scale)
        y = ops.linear(att, self._packs.linear(self.proj.weight), self.proj.bias.detach(), residual=residual,
                       out_dtype=out_dtype)
        return y.view(B, N, C)

    def forward(self, x, H, W):
        return self._forward(_as_tokens_bf16(x), H, W, out_dtype=x.dtype)


class Block(nn.Module):
    """core/mix_transformer.py:118-155.  x + attn(LN(x)); x + mlp(LN(x)) with both residual adds fused into
    the proj / fc2 GEMM epilogues."""

    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop=0., attn_drop=0.,
                 drop_path=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm, sr_ratio=1):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop,
                              proj_drop=drop, sr_ratio=sr_ratio)
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
        self.apply(_reference_init)

    def _droppath_scale(self, B, device):
        """timm.DropPath in train mode (core/mix_transformer.py:129,152-153): per-sample bernoulli(keep) / keep, drawn
        with torch's device RNG; None = identity (eval mode or drop_prob 0)."""
        dp = getattr(self.drop_path, "drop_prob", 0.0) or 0.0
        if not self.training or dp == 0.0:
            return None
        keep = 1.0 - dp
        return (torch.rand((B,), device=device) < keep).float() / keep

    def forward(self, x, H, W):
        """Gradient-free forward (inference, and train.py:358-359's no_grad feature pass, which the reference runs with
        the module still in train mode: DropPath active).  Gradients go through core/seg_train.py's tape instead."""
        if self.training and torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            raise NotImplementedError("segmif_b200: Block.forward is the gradient-free path; training runs through "
                                      "Network3.forward/_loss (core/seg_train.py) or under torch.no_grad()")
        x = x.contiguous() if x.dtype == torch.float32 else x.float().contiguous()
        B, N, C = x.shape
        if _strict.is_strict() and not self.training:
            return _strict.block(self, x, H, W)
        s1, s2 = self._droppath_scale(B, x.device), self._droppath_scale(B, x.device)
        n1 = ops.layernorm(x, self.norm1.weight.detach(), self.norm1.bias.detach(), self.norm1.eps)
        if s1 is None:
            x = self.attn._forward(n1, H, W, residual=x.view(-1, C))
        else:
            x = ops.scale_add_rows(x.view(-1, C), self.attn._forward(n1, H, W).view(-1, C), s1, N).view(B, N, C)
        n2 = ops.layernorm(x, self.norm2.weight.detach(), self.norm2.bias.detach(), self.norm2.eps)
        if s2 is None:
            x = self.mlp._forward(n2, H, W, residual=x.view(-1, C))
        else:
            x = ops.scale_add_rows(x.view(-1, C), self.mlp._forward(n2, H, W).view(-1, C), s2, N).view(B, N, C)
        return x


class OverlapPatchEmbed(nn.Module):
    """core/mix_transformer.py:158-198.  Stage 1 (7x7 s4, 3 input channels, NCHW fp32 image) is one fused
    direct-conv + LayerNorm kernel; stages 2-4 (3x3 s2 on pixel-major bf16 features) are implicit GEMMs
    followed by the LayerNorm kernel."""

    def __init__(self, img_size=224, patch_size=7, stride=4, in_chans=3, embed_dim=768):
        super().__init__()
        img_size = to_2tuple(img_size)
        patch_size = to_2tuple(patch_size)
        self.img_size = img_size
        self.patch_size = patch_size
        self.stride = stride
        self.H, self.W = img_size[0] // patch_size[0], img_size[1] // patch_size[1]
        self.num_patches = self.H * self.W
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=stride,
                              padding=(patch_size[0] // 2, patch_size[1] // 2))
        self.norm = nn.LayerNorm(embed_dim)
        self._packs = PackCache()
        self.apply(_reference_init)

    def forward_image(self, img, in_scale=None, in_shift=None):
        """img fp32 NCHW [B,3,H,W] -> (tokens fp32 [B,N,C], H/4, W/4); optional fused per-channel input affine."""
        w = self._packs.get(self.proj.weight, lambda t: t.detach().reshape(t.shape[0], -1).t().contiguous().float())
        return ops.patch_embed7_ln(img.contiguous(), w, self.proj.bias.detach(), self.norm.weight.detach(),
                                   self.norm.bias.detach(), self.norm.eps, in_scale, in_shift)

    def forward_tokens(self, x_bf16, B, H, W):
        """x bf16 pixel-major [B, H*W, Cin] -> (tokens fp32 [B, N', C], H', W')."""
        k, s = self.patch_size[0], self.stride
        Cin = x_bf16.shape[-1]
        Cout = self.proj.out_channels
        y = ops.conv(x_bf16, self._packs.conv(self.proj.weight), self.proj.bias.detach(), B=B, H=H, W=W, Cin=Cin,
                     KH=k, KW=k, stride=s, pad=k // 2, Cout=Cout, out_dtype=torch.float32)
        Ho, Wo = (H + 2 * (k // 2) - k) // s + 1, (W + 2 * (k // 2) - k) // s + 1
        t = ops.layernorm(y, self.norm.weight.detach(), self.norm.bias.detach(), self.norm.eps, out_dtype=torch.float32)
        return t.view(B, Ho * Wo, Cout), Ho, Wo

    def forward(self, x):
        if x.shape[1] == 3 and self.patch_size[0] == 7 and self.stride == 4:
            return self.forward_image(x.float())
        B, C, H, W = x.shape
        tok = ops.nchw_to_nhwc(x.float().contiguous(), out_dtype=torch.bfloat16)
        return self.forward_tokens(tok, B, H, W)


class MixVisionTransformer(nn.Module):
    """core/mix_transformer.py:201-375."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dims=[64, 128, 256, 512],
                 num_heads=[1, 2, 4, 8], mlp_ratios=[4, 4, 4, 4], qkv_bias=False, qk_scale=None, drop_rate=0.,
                 attn_drop_rate=0., drop_path_rate=0., norm_layer=nn.LayerNorm, depths=[3, 4, 6, 3],
                 sr_ratios=[8, 4, 2, 1]):
        super().__init__()
        self.num_classes = num_classes
        self.depths = depths
        self.embed_dims = embed_dims
        self.patch_embed1 = OverlapPatchEmbed(img_size=img_size, patch_size=7, stride=4, in_chans=in_chans,
                                              embed_dim=embed_dims[0])
        self.patch_embed2 = OverlapPatchEmbed(img_size=img_size // 4, patch_size=3, stride=2, in_chans=embed_dims[0],
                                              embed_dim=embed_dims[1])
        self.patch_embed3 = OverlapPatchEmbed(img_size=img_size // 8, patch_size=3, stride=2, in_chans=embed_dims[1],
                                              embed_dim=embed_dims[2])
        self.patch_embed4 = OverlapPatchEmbed(img_size=img_size // 16, patch_size=3, stride=2, in_chans=embed_dims[2],
                                              embed_dim=embed_dims[3])
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, sum(depths))]
        cur = 0
        for s in range(4):
            blocks = nn.ModuleList([
                Block(dim=embed_dims[s], num_heads=num_heads[s], mlp_ratio=mlp_ratios[s], qkv_bias=qkv_bias,
                      qk_scale=qk_scale, drop=drop_rate, attn_drop=attn_drop_rate, drop_path=dpr[cur + i],
                      norm_layer=norm_layer, sr_ratio=sr_ratios[s]) for i in range(depths[s])])
            setattr(self, f"block{s + 1}", blocks)
            setattr(self, f"norm{s + 1}", norm_layer(embed_dims[s]))
            cur += depths[s]
        self.apply(_reference_init)

    def reset_drop_path(self, drop_path_rate):
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, sum(self.depths))]
        cur = 0
        for s in range(4):
            for i in range(self.depths[s]):
                getattr(self, f"block{s + 1}")[i].drop_path.drop_prob = dpr[cur + i]
            cur += self.depths[s]

    def freeze_patch_emb(self):
        self.patch_embed1.requires_grad = False

    @torch.jit.ignore
    def no_weight_decay(self):
        return {'pos_embed1', 'pos_embed2', 'pos_embed3', 'pos_embed4', 'cls_token'}

    # ---- the stage pipeline, pixel-major throughout -----------------------------------------------------
    def forward_stages(self, x, in_scale=None, in_shift=None, n_stages=4):
        """Returns per stage (tokens bf16 [B, N, C] after the stage LayerNorm, H, W); fp32 tokens in strict precision."""
        if _strict.is_strict() and not self.training:
            return _strict.encoder_stages(self, x, in_scale, in_shift, n_stages)
        B = x.shape[0]
        outs = []
        tok_bf16, H, W = None, None, None
        for s in range(n_stages):
            pe = getattr(self, f"patch_embed{s + 1}")
            if s == 0:
                tok, H, W = pe.forward_image(x.float() if x.dtype != torch.float32 else x, in_scale, in_shift)
            else:
                tok, H, W = pe.forward_tokens(tok_bf16, B, H, W)
            for blk in getattr(self, f"block{s + 1}"):
                tok = blk(tok, H, W)
            norm = getattr(self, f"norm{s + 1}")
            tok_bf16 = ops.layernorm(tok, norm.weight.detach(), norm.bias.detach(), norm.eps)
            outs.append((tok_bf16, H, W))
        return outs

    def forward_features(self, x):
        """core/mix_transformer.py:312-348 -- four NCHW fp32 maps (C = 64,128,320,512 at /4,/8,/16,/32)."""
        B = x.shape[0]
        feats = []
        for tok, H, W in self.forward_stages(x):
            C = tok.shape[-1]
            feats.append(ops.nhwc_to_nchw(tok, B, H * W, C).view(B, C, H, W))
        return feats

    def forward(self, x):
        return self.forward_features(x)

    def forward_fusion(self, x):
        """core/mix_transformer.py:358-375 -- stage-1/2 maps bilinearly upsampled to the input size.
        Returned tensors are logically NCHW [B,C,H,W] (the reference's interface) but are bf16 views of
        pixel-major storage (channels_last strides), which Fusion_Network3_ac consumes without a copy.
        Stages 3-4, which the reference computes and discards here, are skipped."""
        if _strict.is_strict() and not self.training:
            return _strict.forward_fusion(self, x)
        B, _, H, W = x.shape
        outs = []
        for tok, h, w in self.forward_stages(x, n_stages=2):
            C = tok.shape[-1]
            up = ops.bilinear_nhwc(tok, B, h, w, C, H, W)            # [B, H, W, C] bf16
            outs.append(up.permute(0, 3, 1, 2))
        return outs[0], outs[1]


def _mit(embed_dims, depths):
    class _M(MixVisionTransformer):
        def __init__(self, **kwargs):     # the reference ignores **kwargs too (mix_transformer.py:389-434)
            super().__init__(patch_size=4, embed_dims=embed_dims, num_heads=[1, 2, 5, 8], mlp_ratios=[4, 4, 4, 4],
                             qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6), depths=depths,
                             sr_ratios=[8, 4, 2, 1], drop_rate=0.0, drop_path_rate=0.1)
    return _M


class mit_b0(_mit([32, 64, 160, 256], [2, 2, 2, 2])):
    pass


class mit_b1(_mit([64, 128, 320, 512], [2, 2, 2, 2])):
    pass


class mit_b2(_mit([64, 128, 320, 512], [3, 4, 6, 3])):
    pass


class mit_b3(_mit([64, 128, 320, 512], [3, 4, 18, 3])):
    pass


class mit_b4(_mit([64, 128, 320, 512], [3, 8, 27, 3])):
    pass


class mit_b5(_mit([64, 128, 320, 512], [3, 6, 40, 3])):
    pass
