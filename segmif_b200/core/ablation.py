"""The paper's ablation networks (core/model_fusion.py:363-425 CrossPath_M / CrossPath_S, :465-523 FeatureFusionModule_SoAM /
_MoAM, :626-1025 Fusion_Network3, _Con, _Add, _Average, _S, _M, AttentionModule, Fusion_Network_rmseg[_att]) -- SURVEY.md 8(f)
row 3: compositions of the kernels the hot path already has, with the reference's class / attribute / state_dict surface.

They use 32-channel streams (8 heads of dimension 4), which the fused bf16 kernels of the main network (specialised to
64 channels) do not cover, so they run on the width-generic fp32-parity kernels of segmif_b200.strict: split-bf16 tcgen05
contractions for every conv / Linear, fp32 LayerNorm / element-wise pieces, fp64 k^T v.  Inference only (no_grad)."""
import ctypes

import torch
import torch.nn as nn

from .. import _lib, ops
from .. import strict as S
from ..ops import ACT_NONE, ACT_PRELU, ACT_RELU
from ..packing import PackCache
from .mix_transformer import _reference_init
from .model_fusion import DRDB, CrossAttention, CrossAttention2, FeatureFusionModule, _no_autograd

_p = S._p


# ------------------------------------------------------------------------------------------------ generic pieces (fp32 rows [M, C])
def _linear(cache, x_planes, lin, K, act=ACT_NONE, a_coff=0, want_planes=False, **kw):
    return S.gemm(x_planes, K, S.pack_linear(cache, lin.weight), lin.out_features, a_coff=a_coff,
                  bias=lin.bias.detach() if lin.bias is not None else None, act=act, want_planes=want_planes, **kw)


def _conv3x3(cache, x_planes, conv, B, H, W, act=ACT_NONE, alpha=None, a_coff=0, want_f32=True, want_planes=False):
    """nn.Conv2d(k=3, padding=1) on pixel-major planes; Cout % 4 == 0 (Cout == 1 goes through conv3x3_out1_f32)."""
    return S.gemm(x_planes, conv.in_channels, S.pack_conv_taps(cache, conv.weight), conv.out_channels, a_coff=a_coff,
                  bias=conv.bias.detach(), act=act, alpha=alpha, patch=(B, H, W, 1), want_f32=want_f32, want_planes=want_planes)


def _conv1x1(cache, x_planes, conv, want_planes=False, want_f32=True):
    return S.gemm(x_planes, conv.in_channels, S.pack_conv_im2col(cache, conv.weight), conv.out_channels, bias=conv.bias.detach(),
                  want_planes=want_planes, want_f32=want_f32)


def _add(a, b):
    return ops.ew2(a, b, ops.EW_LINCOMB, 1.0, 1.0)


def _ctx_weights(k_src, k_coff, v_src, v_coff, ld, C, heads, B, HW, scale):
    """Block-diagonal per-image matrices that apply softmax_{dim=-2}(k^T v * scale) per head; k, v = column slices of fp32 rows."""
    dev = k_src.device
    st = ops._prep(k_src, v_src)
    nchunk = max(1, min(296 // max(B, 1), (HW + 511) // 512))
    partials = torch.empty((B, nchunk, C, C), dtype=torch.float64, device=dev)
    _lib.call("segmif_xty_f64", _p(k_src), ld, k_coff, C, _p(v_src), ld, v_coff, C, B, HW, _p(partials), nchunk, st)
    ctx = torch.empty((B, heads, C // heads, C // heads), dtype=torch.float32, device=dev)
    w = torch.empty((B, C, C), dtype=torch.float32, device=dev)
    _lib.call("segmif_ctx_blockdiag", _p(partials), nchunk, C, heads, float(scale), _p(ctx), _p(w), B, st)
    return S.pack_rows(w.view(B * C, C)).view(B, C, 3, -1), ctx


def _apply_ctx(q_planes, q_coff, wctx, C, B, HW, dst, dst_coff):
    for b in range(B):
        S.gemm(q_planes, C, wctx[b], C, a_coff=q_coff, row0=b * HW, rows=HW, dst=dst, dst_coff=dst_coff)


def cross_path_generic(cp, x1, x2, seg, B, HW, mode):
    """core/model_fusion.py:350-361 ('full'), :384-395 ('M': MoAM only) and :417-428 ('S': SoAM only) for any dim that is
    a multiple of 32 up to 64.  x1, x2, seg: fp32 rows [B*HW, dim].  Returns the two LayerNorm outputs."""
    C = cp.channel_proj1.in_features
    M = B * HW
    dev = x1.device
    P = [_linear(cp._packs, S.split(x), proj, C, act=ACT_RELU, want_planes=True)
         for x, proj in ((x1, cp.channel_proj1), (x2, cp.channel_proj2), (seg, cp.channel_proj3))]     # (fp32 [M, 2C], planes): y | u
    width = 2 * C if mode == "full" else C
    cat = [torch.empty((M, width), dtype=torch.float32, device=dev) for _ in range(2)]
    if mode in ("full", "S"):                               # SoAM: ctx_i from kv_i(y_i); z_i = y3 @ ctx_i
        att = cp.cross_attn2
        for i, kvl in enumerate((att.kv1, att.kv2)):
            kv, _ = _linear(cp._packs, P[i][1], kvl, C)                                                  # [M, 2C]: k | v
            wctx, _ = _ctx_weights(kv, 0, kv, C, 2 * C, C, att.num_heads, B, HW, att.scale)
            _apply_ctx(P[2][1], 0, wctx, C, B, HW, cat[i], 0)
    if mode in ("full", "M"):                               # MoAM: ctx3 from kv3(u3); v_i = u_i @ ctx3
        att = cp.cross_attn
        kv, _ = _linear(cp._packs, P[2][1], att.kv3, C, a_coff=C)
        wctx, _ = _ctx_weights(kv, 0, kv, C, 2 * C, C, att.num_heads, B, HW, att.scale)
        for i in range(2):
            _apply_ctx(P[i][1], C, wctx, C, B, HW, cat[i], C if mode == "full" else 0)
    outs = []
    for i, (x, end, norm) in enumerate(((x1, cp.end_proj1, cp.norm1), (x2, cp.end_proj2, cp.norm2))):
        pre, _ = _linear(cp._packs, S.split(cat[i]), end, width, residual=x)
        outs.append(S._ln(pre, norm))
    return outs[0], outs[1]


class _CrossPathVariant(nn.Module):
    MODE = "full"

    def __init__(self, dim, reduction=1, num_heads=8, norm_layer=nn.LayerNorm):
        super().__init__()
        if reduction != 1 or dim % 4 or dim > 64 or dim % num_heads or dim // num_heads > 8:
            raise NotImplementedError("segmif_b200: cross paths support reduction=1, dim <= 64, head dim <= 8")
        self.channel_proj1 = nn.Linear(dim, dim * 2)
        self.channel_proj2 = nn.Linear(dim, dim * 2)
        self.channel_proj3 = nn.Linear(dim, dim * 2)
        self.act1, self.act2, self.act3 = nn.ReLU(inplace=True), nn.ReLU(inplace=True), nn.ReLU(inplace=True)
        if self.MODE in ("full", "M"):
            self.cross_attn = CrossAttention(dim, num_heads=num_heads)
        if self.MODE in ("full", "S"):
            self.cross_attn2 = CrossAttention2(dim, num_heads=num_heads)
        width = dim * 2 if self.MODE == "full" else dim
        self.end_proj1 = nn.Linear(width, dim)
        self.end_proj2 = nn.Linear(width, dim)
        self.norm1 = norm_layer(dim)
        self.norm2 = norm_layer(dim)
        self._packs = PackCache()

    def forward(self, x1, x2, segfeature):
        _no_autograd(self, x1, x2, segfeature)
        B, N, C = x1.shape
        rows = lambda t: t.float().contiguous().view(B * N, C)
        o1, o2 = cross_path_generic(self, rows(x1), rows(x2), rows(segfeature), B, N, self.MODE)
        return o1.view(B, N, C), o2.view(B, N, C)


class CrossPath_M(_CrossPathVariant):
    """core/model_fusion.py:363-395 (MoAM only)."""
    MODE = "M"


class CrossPath_S(_CrossPathVariant):
    """core/model_fusion.py:397-428 (SoAM only)."""
    MODE = "S"


class CrossPath32(_CrossPathVariant):
    """core/model_fusion.py:329-361 at a width other than 64 (FeatureFusionModule(32) of Fusion_Network3)."""
    MODE = "full"


class _FfmVariant(nn.Module):
    CROSS = None

    def __init__(self, dim, reduction=1, num_heads=8, norm_layer=nn.BatchNorm2d):
        super().__init__()
        self.cross = self.CROSS(dim=dim, reduction=reduction, num_heads=num_heads)
        self.apply(_reference_init)

    def forward_rows(self, x1, x2, seg, B, HW):
        return cross_path_generic(self.cross, x1, x2, seg, B, HW, self.cross.MODE)

    def forward(self, x1, x2, segfeature):
        _no_autograd(self, x1, x2, segfeature)
        B, C, H, W = x1.shape
        o1, o2 = self.forward_rows(S.nchw_to_rows_f32(x1), S.nchw_to_rows_f32(x2), S.nchw_to_rows_f32(segfeature), B, H * W)
        return tuple(ops.nhwc_to_nchw(o, B, H * W, C).view(B, C, H, W) for o in (o1, o2))


class FeatureFusionModule_SoAM(_FfmVariant):
    """core/model_fusion.py:465-494."""
    CROSS = CrossPath_S


class FeatureFusionModule_MoAM(_FfmVariant):
    """core/model_fusion.py:495-523."""
    CROSS = CrossPath_M


class FeatureFusionModule32(_FfmVariant):
    """FeatureFusionModule (core/model_fusion.py:430-463) for widths the fused 64-channel kernels do not cover."""
    CROSS = CrossPath32


def _ffm(dim):
    return FeatureFusionModule(dim) if dim == 64 else FeatureFusionModule32(dim)


class AttentionModule(nn.Module):
    """core/model_fusion.py:759-771: conv3x3 -> ReLU -> conv3x3, then out * sigmoid(out)."""

    def __init__(self):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv2d(32, 32, 3, padding=1), nn.ReLU(inplace=True), nn.Conv2d(32, 32, 3, padding=1))
        self._packs = PackCache()

    def forward_rows(self, x_planes, B, H, W):
        _, t = _conv3x3(self._packs, x_planes, self.conv[0], B, H, W, act=ACT_RELU, want_f32=False, want_planes=True)
        y, _ = _conv3x3(self._packs, t, self.conv[2], B, H, W)
        st = ops._prep(y)
        out = torch.empty_like(y)
        _lib.call("segmif_sigmoid_gate", _p(y), _p(out), y.numel(), st)
        return out

    def forward(self, x1):
        _no_autograd(self, x1)
        B, C, H, W = x1.shape
        out = self.forward_rows(S.split(S.nchw_to_rows_f32(x1)), B, H, W)
        return ops.nhwc_to_nchw(out, B, H * W, C).view(B, C, H, W)


# ------------------------------------------------------------------------------------------------ networks
class _AblationNet(nn.Module):
    """Shared scaffolding: conv1_ir / conv1_vis (1 -> C) + shared PReLU, four DRDBs, conv2 (+ conv21 [+ conv22]) tail."""
    C = 32

    def _stem(self):
        C = self.C
        self.conv1_ir = nn.Conv2d(1, C, 3, padding=1)
        self.conv1_vis = nn.Conv2d(1, C, 3, padding=1)
        for i in range(1, 5):
            setattr(self, f"DRDB{i}", DRDB(in_ch=C))
        self.conv2 = nn.Conv2d(2 * C, C, 3, padding=1)
        self._packs = PackCache()

    def _head32(self):
        self.conv21 = nn.Conv2d(32, 1, 3, padding=1)
        self.relu = nn.PReLU()
        self.conv3 = nn.Conv2d(64, 32, 1, padding=0)
        self.conv4 = nn.Conv2d(128, 32, 1, padding=0)

    # -- pieces on fp32 rows -------------------------------------------------------------------------------------
    def _in(self, img, conv):
        return S.conv3x3_in1_f32(img.float(), self._packs.taps_f32(conv.weight), conv.bias.detach(), self.relu.weight.detach(), self.C)

    def _drdb(self, d, x, B, H, W):
        g = S.Planes(B * H * W, d.total, x.device)
        S.split(x, planes=g)
        return S.drdb(d, g, x, B, H, W, want_planes=False)[0]

    def _tail(self, x1, x2, B, H, W):
        """relu(conv2(cat(x1, x2))) -> relu(conv21(.)) [-> relu(conv22(.))], `relu` being the shared PReLU; returns [B,1,H,W]."""
        C, alpha = self.C, self.relu.weight.detach()
        cat = S.Planes(B * H * W, 2 * C, x1.device)
        S.split(x1, planes=cat, coff_p=0)
        S.split(x2, planes=cat, coff_p=C)
        if hasattr(self, "conv22"):
            _, f = _conv3x3(self._packs, cat, self.conv2, B, H, W, act=ACT_PRELU, alpha=alpha, want_f32=False, want_planes=True)
            f3, _ = _conv3x3(self._packs, f, self.conv21, B, H, W, act=ACT_PRELU, alpha=alpha)
            return S.conv3x3_out1_f32(f3, self._packs.taps_f32(self.conv22.weight), self.conv22.bias.detach(), alpha, B, H, W, 32)
        fa, _ = _conv3x3(self._packs, cat, self.conv2, B, H, W, act=ACT_PRELU, alpha=alpha)
        return S.conv3x3_out1_f32(fa, self._packs.taps_f32(self.conv21.weight), self.conv21.bias.detach(), alpha, B, H, W, C)

    def _seg(self, out, conv):
        return _conv1x1(self._packs, S.split(S.nchw_to_rows_f32(out)), conv)[0]

    def _streams(self, ir, vis, B, H, W):
        x1 = self._drdb(self.DRDB1, self._in(ir, self.conv1_ir), B, H, W)
        x2 = self._drdb(self.DRDB2, self._in(vis, self.conv1_vis), B, H, W)
        return x1, x2

    def _check(self, *t):
        _no_autograd(self, *t)
        if self.training and torch.is_grad_enabled():
            raise NotImplementedError("segmif_b200: the ablation networks are inference only (wrap in torch.no_grad())")


class Fusion_Network3(_AblationNet):
    """core/model_fusion.py:626-660."""
    FFM = staticmethod(_ffm)

    def __init__(self):
        super().__init__()
        self._stem()
        self._head32()
        self.ffm = self.FFM(32)
        if type(self) is Fusion_Network3:
            self.ffm2 = _ffm(32)                      # allocated and saved by the reference, never used (:639)
        # module order fixes the state_dict key order; values are loaded by name

    def forward(self, ir, vis, out1, out2):
        self._check(ir, vis, out1, out2)
        B, _, H, W = ir.shape
        x1, x2 = self._streams(ir, vis, B, H, W)
        x1, x2 = self.ffm.forward_rows(x1, x2, self._seg(out1, self.conv3), B, H * W)
        x1, x2 = self._drdb(self.DRDB3, x1, B, H, W), self._drdb(self.DRDB4, x2, B, H, W)
        x1, x2 = self.ffm.forward_rows(x1, x2, self._seg(out2, self.conv4), B, H * W)
        return self._tail(x1, x2, B, H, W)


class Fusion_Network3_S(Fusion_Network3):
    """core/model_fusion.py:821-855 (SoAM only)."""
    FFM = staticmethod(lambda dim: FeatureFusionModule_SoAM(dim))


class Fusion_Network3_M(Fusion_Network3):
    """core/model_fusion.py:856-890 (MoAM only)."""
    FFM = staticmethod(lambda dim: FeatureFusionModule_MoAM(dim))


class Fusion_Network3_Con(_AblationNet):
    """core/model_fusion.py:661-709: segmentation features concatenated and mixed by plain 3x3 convs (no attention)."""
    MIX_IN = 64

    def __init__(self):
        super().__init__()
        self._stem()
        for n in ("conv211", "conv221", "conv411", "conv421"):
            setattr(self, n, nn.Conv2d(self.MIX_IN, 32, 3, padding=1))
        self._head32()

    def _mix(self, x, seg, conv, B, H, W):
        cat = S.Planes(B * H * W, 64, x.device)
        S.split(x, planes=cat, coff_p=0)
        S.split(seg, planes=cat, coff_p=32)
        return _conv3x3(self._packs, cat, conv, B, H, W)[0]

    def forward(self, ir, vis, out1, out2):
        self._check(ir, vis, out1, out2)
        B, _, H, W = ir.shape
        x1, x2 = self._streams(ir, vis, B, H, W)
        s1, s2 = self._seg(out1, self.conv3), self._seg(out2, self.conv4)
        x1, x2 = self._mix(x1, s1, self.conv211, B, H, W), self._mix(x2, s1, self.conv221, B, H, W)
        x1, x2 = self._drdb(self.DRDB3, x1, B, H, W), self._drdb(self.DRDB4, x2, B, H, W)
        x1, x2 = self._mix(x1, s2, self.conv411, B, H, W), self._mix(x2, s2, self.conv421, B, H, W)
        return self._tail(x1, x2, B, H, W)


class Fusion_Network3_Add(Fusion_Network3_Con):
    """core/model_fusion.py:710-758: segmentation features ADDED, then a 3x3 conv."""
    MIX_IN = 32

    def _mix(self, x, seg, conv, B, H, W):
        return _conv3x3(self._packs, S.split(_add(x, seg)), conv, B, H, W)[0]


class Fusion_Network3_Average(_AblationNet):
    """core/model_fusion.py:772-820: sigmoid-gated AttentionModules on both streams, summed."""

    def __init__(self):
        super().__init__()
        self._stem()
        for i in range(1, 9):
            setattr(self, f"att{i}", AttentionModule())
        self._head32()

    def _att(self, i, x, B, H, W):
        return getattr(self, f"att{i}").forward_rows(S.split(x), B, H, W)

    def forward(self, ir, vis, out1, out2):
        self._check(ir, vis, out1, out2)
        B, _, H, W = ir.shape
        x1, x2 = self._streams(ir, vis, B, H, W)
        s1, s2 = self._seg(out1, self.conv3), self._seg(out2, self.conv4)
        x1 = _add(self._att(1, x1, B, H, W), self._att(2, s1, B, H, W))
        x2 = _add(self._att(3, x2, B, H, W), self._att(4, s1, B, H, W))
        x1, x2 = self._drdb(self.DRDB3, x1, B, H, W), self._drdb(self.DRDB4, x2, B, H, W)
        x1 = _add(self._att(5, x1, B, H, W), self._att(6, s2, B, H, W))
        x2 = _add(self._att(7, x2, B, H, W), self._att(8, s2, B, H, W))
        return self._tail(x1, x2, B, H, W)


class Fusion_Network_rmseg(_AblationNet):
    """core/model_fusion.py:938-973: the fusion network without any segmentation input (64-channel streams)."""
    C = 64

    def __init__(self):
        super().__init__()
        self._stem()
        self.conv21 = nn.Conv2d(64, 32, 3, padding=1)
        self.conv22 = nn.Conv2d(32, 1, 3, padding=1)
        self.relu = nn.PReLU()

    def _forward(self, ir, vis):
        self._check(ir, vis)
        B, _, H, W = ir.shape
        x1, x2 = self._streams(ir, vis, B, H, W)
        x1, x2 = self._drdb(self.DRDB3, x1, B, H, W), self._drdb(self.DRDB4, x2, B, H, W)
        return self._tail(x1, x2, B, H, W), x1, x2, (B, H, W)

    def forward(self, ir, vis):
        return self._forward(ir, vis)[0]


class Fusion_Network_rmseg_att(Fusion_Network_rmseg):
    """core/model_fusion.py:974-1025: additionally returns the two stream features [x1, x2] (NCHW)."""

    def forward(self, ir, vis):
        out, x1, x2, (B, H, W) = self._forward(ir, vis)
        nchw = lambda t: ops.nhwc_to_nchw(t, B, H * W, 64).view(B, 64, H, W)
        return out, [nchw(x1), nchw(x2)]


__all__ = ["CrossPath_M", "CrossPath_S", "FeatureFusionModule_SoAM", "FeatureFusionModule_MoAM", "AttentionModule",
           "Fusion_Network3", "Fusion_Network3_S", "Fusion_Network3_M", "Fusion_Network3_Con", "Fusion_Network3_Add",
           "Fusion_Network3_Average", "Fusion_Network_rmseg", "Fusion_Network_rmseg_att"]
