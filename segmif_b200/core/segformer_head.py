"""SegFormer all-MLP decode head with the reference's surface (core/segformer_head.py of SegMiF).

Pixel-major pipeline: the four per-stage Linear embeddings are tensor-core GEMMs; c2..c4 are bilinearly
resized straight into their channel slice of one [B, H/4, W/4, 4*E] bf16 buffer and c1's GEMM writes its
slice in place (no torch.cat); linear_fuse (1x1 conv, no bias) + BatchNorm (eval: folded into the packed
weights) + ReLU is one GEMM; Dropout2d is the identity in eval; linear_pred is a GEMM with fp32 output.
"""
import torch
import torch.nn as nn

from .. import ops
from .. import strict as _strict
from ..ops import ACT_NONE, ACT_RELU
from ..packing import PackCache


class ConvModule(nn.Module):
    """Stand-in for mmcv.cnn.ConvModule as instantiated at core/segformer_head.py:50-55: 1x1 conv without bias,
    BatchNorm2d as `.bn`, ReLU as `.activate` (same attribute and state_dict key names)."""

    def __init__(self, in_channels, out_channels, kernel_size, norm_cfg=None, **kwargs):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, bias=norm_cfg is None)
        self.bn = nn.BatchNorm2d(out_channels) if norm_cfg is not None else None
        self.activate = nn.ReLU(inplace=True)


class DepthwiseSeparableConvModule(nn.Module):
    """Imported by the reference (segformer_head.py:11) but never instantiated."""


class MLP(nn.Module):
    """Linear Embedding (core/segformer_head.py:13-24)."""

    def __init__(self, input_dim=2048, embed_dim=768):
        super().__init__()
        self.proj = nn.Linear(input_dim, embed_dim)
        self._packs = PackCache()

    def forward_tokens(self, tok_bf16, **kw):
        return ops.linear(tok_bf16, self._packs.linear(self.proj.weight), self.proj.bias.detach(), **kw)

    def forward(self, x):
        B, C = x.shape[0], x.shape[1]
        tok = ops.nchw_to_nhwc(x.float().contiguous(), out_dtype=torch.bfloat16)
        return self.forward_tokens(tok, out_dtype=torch.float32).view(B, -1, self.proj.out_features)


class SegFormerHead(nn.Module):
    """core/segformer_head.py:27-82."""

    def __init__(self, feature_strides=None, in_channels=128, embedding_dim=256, num_classes=20, **kwargs):
        super().__init__()
        self.in_channels = in_channels
        self.num_classes = num_classes
        assert len(feature_strides) == len(self.in_channels)
        assert min(feature_strides) == feature_strides[0]
        self.feature_strides = feature_strides
        c1_in, c2_in, c3_in, c4_in = self.in_channels
        self.embedding_dim = embedding_dim
        self.linear_c4 = MLP(input_dim=c4_in, embed_dim=embedding_dim)
        self.linear_c3 = MLP(input_dim=c3_in, embed_dim=embedding_dim)
        self.linear_c2 = MLP(input_dim=c2_in, embed_dim=embedding_dim)
        self.linear_c1 = MLP(input_dim=c1_in, embed_dim=embedding_dim)
        self.dropout = nn.Dropout2d(0.1)
        self.linear_fuse = ConvModule(in_channels=embedding_dim * 4, out_channels=embedding_dim, kernel_size=1,
                                      norm_cfg=dict(type='BN', requires_grad=True))
        self.linear_pred = nn.Conv2d(embedding_dim, self.num_classes, kernel_size=1)
        self._packs = PackCache()

    def _fuse_pack(self):
        bn, conv = self.linear_fuse.bn, self.linear_fuse.conv

        def fold(w, g, b, mean, var):
            scale = g.float() / torch.sqrt(var.float() + bn.eps)
            wf = (w.float().reshape(w.shape[0], -1) * scale[:, None]).to(torch.bfloat16)
            return wf.reshape(w.shape[0], 1, -1).contiguous(), (b.float() - mean.float() * scale).contiguous()
        return self._packs.get_multi([conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var], fold, "fuse")

    def forward_tokens(self, stages):
        """stages: four (tokens bf16 [B, N_i, C_i], H_i, W_i); returns pixel-major fp32 logits [B, H1, W1, nc]."""
        if self.training:
            raise NotImplementedError("segmif_b200: train-mode decode head (batch-stat BN, Dropout2d, autograd) is not "
                                      "built yet; call .eval()")
        if _strict.is_strict():
            return _strict.head_logits(self, [(t.float(), h, w) for t, h, w in stages])
        (t1, h1, w1), (t2, h2, w2), (t3, h3, w3), (t4, h4, w4) = stages
        B, E = t1.shape[0], self.embedding_dim
        cat = torch.empty((B, h1, w1, 4 * E), dtype=torch.bfloat16, device=t1.device)
        for slot, (mlp, t, h, w) in enumerate(((self.linear_c4, t4, h4, w4), (self.linear_c3, t3, h3, w3),
                                               (self.linear_c2, t2, h2, w2))):
            y = mlp.forward_tokens(t)                                              # [B*h*w, E] bf16
            ops.bilinear_nhwc(y, B, h, w, E, h1, w1, out=cat, ld_dst=4 * E, dst_coff=slot * E)
        self.linear_c1.forward_tokens(t1, out=cat.view(-1, 4 * E), ld_dst=4 * E, dst_coff=3 * E)
        wf, bf = self._fuse_pack()
        fused = ops.linear(cat, wf, bf, act=ACT_RELU)                               # conv(no bias)+BN+ReLU
        logits = ops.linear(fused, self._packs.conv(self.linear_pred.weight), self.linear_pred.bias.detach(),
                            act=ACT_NONE, out_dtype=torch.float32)
        return logits.view(B, h1, w1, self.num_classes)

    def forward(self, x):
        c1 = x[0]
        B = c1.shape[0]
        stages = []
        for c in x:
            tok = ops.nchw_to_nhwc(c.float().contiguous(), out_dtype=torch.float32 if _strict.is_strict() else torch.bfloat16)
            stages.append((tok, c.shape[2], c.shape[3]))
        logits = self.forward_tokens(stages)
        h1, w1 = c1.shape[2], c1.shape[3]
        return ops.nhwc_to_nchw(logits, B, h1 * w1, self.num_classes).view(B, self.num_classes, h1, w1)
