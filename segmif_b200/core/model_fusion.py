"""Segmentation wrapper (WeTr / Network3), fusion network (Fusion_Network3_ac, DRDB) and hierarchical
interactive attention (FeatureFusionModule / CrossPath / CrossAttention[2]) with the reference's class,
attribute and state_dict surface (core/model_fusion.py of SegMiF), computed by the segmif_b200 kernels.

Full-resolution feature maps live pixel-major in bf16.  A DRDB owns one [B, H, W, 224] growth buffer: the
five dilated convs append their 32 channels in place (torch.cat never runs) and the 1x1 conv reads all 224.
conv3 / conv4 on the segmentation features are folded into channel_proj3 on the host (both are linear and
nothing sits between them), so the FFM kernels read the upsampled encoder features directly.
"""
import os

import torch
import torch.nn as nn

from .. import ops
from .. import strict as _strict
from ..ops import ACT_PRELU, ACT_RELU
from ..packing import PackCache
from . import mix_transformer
from .mix_transformer import _reference_init
from .segformer_head import SegFormerHead

IMAGENET_MEAN = [123.675, 116.28, 103.53]
IMAGENET_STD = [58.395, 57.12, 57.375]


def _no_autograd(module, *tensors):
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors):
        raise NotImplementedError(
            f"segmif_b200: {type(module).__name__} was asked for gradients, but only the forward kernels exist so "
            "far (backward is the next row of the build plan). Wrap the call in torch.no_grad().")


def _pixel_major_bf16(x):
    """Logical NCHW tensor -> (pixel-major bf16 storage [B, HW, C], C).  Zero-copy when `x` already is a
    channels_last bf16 view (what MixVisionTransformer.forward_fusion returns); otherwise one convert kernel."""
    B, C, H, W = x.shape
    if x.dtype == torch.bfloat16 and x.permute(0, 2, 3, 1).is_contiguous():
        return x.permute(0, 2, 3, 1).reshape(B, H * W, C)
    return ops.nchw_to_nhwc(x.float().contiguous(), out_dtype=torch.bfloat16)


# ------------------------------------------------------------------------------------------------ colour
def RGB2YCrCb(input_im):
    """core/model_fusion.py:69-92 (device-agnostic: runs on the input's device)."""
    return ops.rgb2ycrcb(input_im.float().contiguous())


def YCrCb2RGB(input_im):
    """core/model_fusion.py:94-111."""
    return ops.ycrcb2rgb(input_im.float().contiguous())


# ------------------------------------------------------------------------------------------------ segmentation
class WeTr(nn.Module):
    """core/model_fusion.py:9-68."""

    def __init__(self, backbone, num_classes=20, embedding_dim=256, pretrained=None):
        super().__init__()
        self.num_classes = num_classes
        self.embedding_dim = embedding_dim
        self.backbone = backbone
        self.feature_strides = [4, 8, 16, 32]
        self.encoder = getattr(mix_transformer, backbone)()
        self.in_channels = self.encoder.embed_dims
        if pretrained:
            self.initialize()
        self.decoder = SegFormerHead(feature_strides=self.feature_strides, in_channels=self.in_channels,
                                     embedding_dim=self.embedding_dim, num_classes=self.num_classes)
        # kept for state_dict / optimizer-group parity; its output is discarded by the reference (:66) so it never runs
        self.classifier = nn.Conv2d(in_channels=self.in_channels[-1], out_channels=self.num_classes, kernel_size=1,
                                    bias=False)

    def initialize(self):
        state_dict = torch.load('pretrained/' + self.backbone + '.pth')
        state_dict.pop('head.weight')
        state_dict.pop('head.bias')
        self.encoder.load_state_dict(state_dict)

    def get_param_groups(self):
        groups = [[], [], []]
        for name, param in list(self.encoder.named_parameters()):
            groups[1 if "norm" in name else 0].append(param)
        for param in list(self.decoder.parameters()):
            groups[2].append(param)
        groups[2].append(self.classifier.weight)
        return groups

    def forward_pixel_major(self, x, in_scale=None, in_shift=None):
        """fp32 logits [B, H/4, W/4, nc], pixel-major (what the fused upsample+argmax / CE kernels consume)."""
        if _strict.is_strict() and not self.training:
            return _strict.wetr_logits(self, x, in_scale, in_shift)
        stages = self.encoder.forward_stages(x, in_scale, in_shift)
        return self.decoder.forward_tokens(stages)

    def forward(self, x):
        _no_autograd(self, x)
        lg = self.forward_pixel_major(x)
        B, h, w, nc = lg.shape
        return ops.nhwc_to_nchw(lg, B, h * w, nc).view(B, nc, h, w)


class Network3(nn.Module):
    """core/model_fusion.py:1068-1104.  The x*255 / ImageNet mean-std normalisation is fused into patch_embed1."""

    def __init__(self, backbone, num_classes=20, embedding_dim=256, pretrained=True):
        super().__init__()
        self.fusion_nums = 2
        self.seg_nums = 2
        self.fusion_channel = 48
        self.seg_channel = 64
        self.denoise_net = WeTr(backbone, num_classes, embedding_dim, pretrained)
        self.mean = IMAGENET_MEAN
        self.std = IMAGENET_STD
        self._affine = {}

    def _input_affine(self, device):
        key = (device, tuple(self.mean), tuple(self.std))
        if key not in self._affine:
            std = torch.tensor(self.std, dtype=torch.float64)
            mean = torch.tensor(self.mean, dtype=torch.float64)
            self._affine = {key: ((255.0 / std).float().to(device), (-mean / std).float().to(device))}
        return self._affine[key]

    def logits_pixel_major(self, fused_seg1):
        sc, sh = self._input_affine(fused_seg1.device)
        return self.denoise_net.forward_pixel_major(fused_seg1, sc, sh)

    def _wants_grad(self, x):
        return torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))

    def forward(self, fused_seg1):
        if self._wants_grad(fused_seg1):
            # training (train.py:222): one autograd node for encoder + head; the NCHW view is what the caller upsamples
            from .seg_train import logits_with_grad
            lg = logits_with_grad(self, fused_seg1)
            return fused_seg1, fused_seg1, lg.permute(0, 3, 1, 2)
        lg = self.logits_pixel_major(fused_seg1)
        B, h, w, nc = lg.shape
        seg_map = ops.nhwc_to_nchw(lg, B, h * w, nc).view(B, nc, h, w)
        return fused_seg1, fused_seg1, seg_map

    def predict_labels(self, fused_seg1, size=None):
        """Extension: test_segmentation.py:169-175 in one call -- logits, bilinear upsample and argmax fused;
        the [B, nc, H, W] upsampled logits are never materialised."""
        lg = self.logits_pixel_major(fused_seg1)
        B, h, w, nc = lg.shape
        H, W = size if size is not None else fused_seg1.shape[2:]
        return ops.upsample_argmax(lg, B, h, w, nc, H, W)

    def _loss(self, fused_seg1, label, criterion):
        """core/model_fusion.py:1090-1097 for criterion = CrossEntropyLoss(ignore_index=...): upsample + CE fused."""
        if not isinstance(criterion, nn.CrossEntropyLoss) or criterion.weight is not None \
                or criterion.reduction != "mean" or getattr(criterion, "label_smoothing", 0.0) != 0.0:
            raise NotImplementedError("segmif_b200: _loss supports plain mean CrossEntropyLoss(ignore_index=...) only")
        if self._wants_grad(fused_seg1):
            from .seg_train import CeFn, logits_with_grad
            return CeFn.apply(logits_with_grad(self, fused_seg1), label, criterion.ignore_index)
        lg = self.logits_pixel_major(fused_seg1)
        B, h, w, nc = lg.shape
        return ops.upsample_ce(lg, B, h, w, nc, label.long().contiguous(), criterion.ignore_index)

    def enhance_net_parameters(self):
        return self.enhance_net.parameters()

    def denoise_net_parameters(self):
        return self.denoise_net.parameters()


Network = Network3     # core/__init__.py:4 of the reference imports a `Network` that does not exist there


# ------------------------------------------------------------------------------------------------ DRDB
class DRDB(nn.Module):
    """Dilated residual dense block (core/model_fusion.py:117-157)."""
    GROWTH_LD = 224
    # growth-layer formulation: "hybrid" (default: x0 slab pushed with N = 96/64, g-slabs pulled with N = 32 + partial add,
    # one launch after the other), "dataflow" (the same stages plus the 1x1 as seven CONCURRENT kernels chained through L2 by
    # per-tile-row counters, csrc/drdb_dataflow.cu: bit-identical to hybrid, measured SLOWER on B200 -- 3.5 ms vs 1.6 ms per DRDB
    # at batch 8, 480x640 -- because every stage keeps its per-SM rate and the SMs are merely divided), "push" (every
    # slab pushed, bf16 partials read-modify-written), "pair" (g-slabs read with N = 64 for two consumer layers at a time + two
    # single-slab pulls: 12 instead of 20 slab-k-steps per tap, but 512 B/px more partial traffic -- measured 1.72 ms against
    # 1.26 ms for the hybrid growth layers, tools/drdb_bench.py), "pull" (one N = 32 conv per layer).
    MODE = os.environ.get("SEGMIF_DRDB_MODE", "hybrid")

    def __init__(self, in_ch=64, growth_rate=32):
        super().__init__()
        c = in_ch
        for i in range(1, 6):
            setattr(self, f"Dcov{i}", nn.Conv2d(c, growth_rate, 3, padding=2, dilation=2))
            c += growth_rate
        self.conv = nn.Conv2d(c, in_ch, 1, padding=0)
        self.in_ch, self.growth, self.total = in_ch, growth_rate, c
        self._packs = PackCache()

    # ---- push form of the five growth layers (see csrc/drdb_tc.cu) -------------------------------------------------
    def _push_packs(self):
        """Per step: bf16 weights of every later layer restricted to one input slab, [n_out][taps*slab] tap-major."""
        convs = [getattr(self, f"Dcov{i}") for i in range(1, 6)]

        def build(*ws):
            def slab(layers, c0, c1):
                w = torch.cat([ws[j - 1].detach().float()[:, c0:c1] for j in layers], 0)          # [n_out, slab, 3, 3]
                w = w.permute(0, 2, 3, 1).reshape(w.shape[0], 9, c1 - c0)                          # [n_out, tap, slab]
                if c1 - c0 == 32:                                                                   # two taps per 128-byte row
                    w = torch.cat([w, torch.zeros_like(w[:, :1])], 1)
                return w.reshape(w.shape[0], -1).to(torch.bfloat16).contiguous()
            g = self.growth
            c = self.in_ch
            return [slab((1, 2, 3), 0, c), slab((4, 5), 0, c), slab((2, 3, 4, 5), c, c + g),
                    slab((3, 4, 5), c + g, c + 2 * g), slab((4, 5), c + 2 * g, c + 3 * g), slab((5,), c + 3 * g, c + 4 * g)]
        return self._packs.get_multi([cv.weight for cv in convs], build, "push")

    def _growth_push(self, buf, part, B, H, W):
        """buf [B,H,W,224] holds x0 in channels 0..63; part [B,H,W,128] receives the partial pre-activations P2..P5."""
        w = self._push_packs()
        b = [getattr(self, f"Dcov{i}").bias.detach() for i in range(1, 6)]
        c, g = self.in_ch, self.growth
        G = lambda coff, bias, pin_off: dict(bias=bias, partial_in=part if pin_off is not None else None,
                                             coff_partial_in=pin_off or 0, dst=buf, coff_dst=coff, relu=True)
        Pn = lambda off, fresh: dict(bias=None, partial_in=None if fresh else part, coff_partial_in=off, dst=part,
                                     coff_dst=off, relu=False)
        ops.drdb_push(buf, w[0], B, H, W, 0, c, [G(c, b[0], None), Pn(0, True), Pn(g, True)])
        ops.drdb_push(buf, w[1], B, H, W, 0, c, [Pn(2 * g, True), Pn(3 * g, True)])
        ops.drdb_push(buf, w[2], B, H, W, c, g, [G(c + g, b[1], 0), Pn(g, False), Pn(2 * g, False), Pn(3 * g, False)])
        ops.drdb_push(buf, w[3], B, H, W, c + g, g, [G(c + 2 * g, b[2], g), Pn(2 * g, False), Pn(3 * g, False)])
        ops.drdb_push(buf, w[4], B, H, W, c + 2 * g, g, [G(c + 3 * g, b[3], 2 * g), Pn(3 * g, False)])
        ops.drdb_push(buf, w[5], B, H, W, c + 3 * g, g, [G(c + 4 * g, b[4], 3 * g)])

    def _hybrid_packs(self):
        """Layer j >= 2 restricted to the g-slabs (input channels in_ch..Cin_j): bf16 [32][9][Cin_j - in_ch]."""
        convs = [getattr(self, f"Dcov{i}") for i in range(2, 6)]

        def build(*ws):
            c = self.in_ch
            return [w.detach().float()[:, c:].permute(0, 2, 3, 1).reshape(w.shape[0], 9, w.shape[1] - c)
                    .to(torch.bfloat16).contiguous() for w in ws]
        return self._packs.get_multi([cv.weight for cv in convs], build, "hybrid")

    def _growth_hybrid(self, buf, part, B, H, W):
        """x0 slab in push form (two launches, N = 96 and 64: layer 1 finished, P2..P5 = x0's share of layers 2..5),
        then layers 2..5 as N = 32 pull convolutions over the g-slabs only, with P_j added before the ReLU."""
        w = self._push_packs()
        wh = self._hybrid_packs()
        b = [getattr(self, f"Dcov{i}").bias.detach() for i in range(1, 6)]
        c, g, ld = self.in_ch, self.growth, buf.shape[-1]
        Pn = lambda off: dict(bias=None, partial_in=None, dst=part, coff_dst=off, relu=False)
        ops.drdb_push(buf, w[0], B, H, W, 0, c, [dict(bias=b[0], partial_in=None, dst=buf, coff_dst=c, relu=True), Pn(0), Pn(g)])
        ops.drdb_push(buf, w[1], B, H, W, 0, c, [Pn(2 * g), Pn(3 * g)])
        for j in range(2, 6):
            cin = g * (j - 1)
            ops.conv(buf, wh[j - 2], b[j - 1], B=B, H=H, W=W, Cin=cin, ld_src=ld, src_coff=c, KH=3, KW=3, pad=2, dil=2,
                     Cout=g, act=ACT_RELU, out=buf.view(-1, ld), ld_dst=ld, dst_coff=c + cin, pre_add=part.view(-1, part.shape[-1]),
                     pre_coff=g * (j - 2), entry="segmif_conv3x3_tc_fwd")

    def _pair_packs(self):
        """Weights of the "pair" formulation: slab pushes with N = 64 (two consumer layers at a time) + two N = 32 pulls."""
        convs = [getattr(self, f"Dcov{i}") for i in range(1, 6)]

        def build(*ws):
            c, g = self.in_ch, self.growth

            def slab(layers, c0, c1):
                w = torch.cat([ws[j - 1].detach().float()[:, c0:c1] for j in layers], 0)
                w = w.permute(0, 2, 3, 1).reshape(w.shape[0], 9, c1 - c0)
                if c1 - c0 == 32:
                    w = torch.cat([w, torch.zeros_like(w[:, :1])], 1)
                return w.reshape(w.shape[0], -1).to(torch.bfloat16).contiguous()

            def pull(j, c0, c1):
                w = ws[j - 1].detach().float()[:, c0:c1]
                return w.permute(0, 2, 3, 1).reshape(w.shape[0], 9, c1 - c0).to(torch.bfloat16).contiguous()
            return [slab((2, 3), c, c + g), slab((4, 5), c, c + 2 * g), slab((4, 5), c + 2 * g, c + 3 * g),
                    pull(3, c + g, c + 2 * g), pull(5, c + 3 * g, c + 4 * g)]
        return self._packs.get_multi([cv.weight for cv in convs], build, "pair")

    def _growth_pair(self, buf, part, B, H, W):
        """x0 pushed as in the hybrid form; then every g-slab is read with N = 64 where two layers consume it:
        g1 -> (finish g2, add into P3); L3 pulls g2 only; g1g2 -> (add into P4, P5); g3 -> (finish g4, add into P5); L5 pulls g4
        only.  12 slab-k-steps per tap on the shared-memory-bound N <= 64 path instead of 20, for 512 B/px more partial traffic."""
        w, pw = self._push_packs(), self._pair_packs()
        b = [getattr(self, f"Dcov{i}").bias.detach() for i in range(1, 6)]
        c, g, ld = self.in_ch, self.growth, buf.shape[-1]
        Pn = lambda off: dict(bias=None, partial_in=None, dst=part, coff_dst=off, relu=False)
        Pa = lambda off: dict(bias=None, partial_in=part, coff_partial_in=off, dst=part, coff_dst=off, relu=False)
        Fin = lambda bias, poff, coff: dict(bias=bias, partial_in=part, coff_partial_in=poff, dst=buf, coff_dst=coff, relu=True)
        pull = lambda wj, bj, soff, doff, poff: ops.conv(buf, wj, bj, B=B, H=H, W=W, Cin=g, ld_src=ld, src_coff=soff, KH=3, KW=3, pad=2,
                                                         dil=2, Cout=g, act=ACT_RELU, out=buf.view(-1, ld), ld_dst=ld, dst_coff=doff,
                                                         pre_add=part.view(-1, part.shape[-1]), pre_coff=poff, entry="segmif_conv3x3_tc_fwd")
        ops.drdb_push(buf, w[0], B, H, W, 0, c, [dict(bias=b[0], partial_in=None, dst=buf, coff_dst=c, relu=True), Pn(0), Pn(g)])
        ops.drdb_push(buf, w[1], B, H, W, 0, c, [Pn(2 * g), Pn(3 * g)])
        ops.drdb_push(buf, pw[0], B, H, W, c, g, [Fin(b[1], 0, c + g), Pa(g)])                     # g1 -> g2, P3
        ops.drdb_push(buf, pw[1], B, H, W, c, 2 * g, [Pa(2 * g), Pa(3 * g)])                       # g1 g2 -> P4, P5
        pull(pw[3], b[2], c + g, c + 2 * g, g)                                                    # L3 over g2 -> g3
        ops.drdb_push(buf, pw[2], B, H, W, c + 2 * g, g, [Fin(b[3], 2 * g, c + 3 * g), Pa(3 * g)])  # g3 -> g4, P5
        pull(pw[4], b[4], c + 3 * g, c + 4 * g, 3 * g)                                            # L5 over g4 -> g5

    @staticmethod
    def _df_words(B, H):
        from .. import _lib
        return _lib.load().segmif_drdb_dataflow_workspace_bytes(B, H) // 4

    def dataflow_timed_out(self):
        """True if a dependency wait of the last dataflow forward hit its timeout (never expected; checked by the tests)."""
        f = getattr(self, "_df_flags", None)
        return bool(f is not None and int(f[-36].item()) != 0)

    def dataflow_stage_times(self):
        """Diagnostics: per stage (push a, push b, L2..L5, 1x1) the [begin, end] of the last dataflow forward in microseconds
        relative to the earliest begin (globaltimer stamps written by the kernels)."""
        f = getattr(self, "_df_flags", None)
        if f is None:
            return None
        t = f[-32:].view(torch.int64)[:14].view(7, 2).cpu().tolist()
        t0 = min(b for b, e in t if e > 0)
        return [((b - t0) / 1e3, (e - t0) / 1e3) if e > 0 else None for b, e in t]

    def forward_buffer(self, buf, B, H, W, out=None, ld_dst=None, dst_coff=0, partials=None):
        """`buf` bf16 [B, H, W, total] with the block input in channels 0..in_ch; appends the five growth slices
        in place, then writes x + relu(conv1x1(all)) to `out` (pixel-major bf16).  `partials` (bf16 [B,H,W,128]
        scratch) selects the push form of the growth layers; without it the per-layer (pull) kernels run."""
        ld = buf.shape[-1]
        cin = self.in_ch
        if partials is not None and DRDB.MODE == "dataflow" and self.in_ch == 64 and self.growth == 32 and ld >= self.total:
            w, wh = self._push_packs(), self._hybrid_packs()
            if out is None:
                ld_dst = self.in_ch
                out = torch.empty((B * H * W, ld_dst), dtype=torch.bfloat16, device=buf.device)
            elif ld_dst is None:
                ld_dst = out.shape[-1]
            self._df_flags = ops.drdb_dataflow(buf, partials, w[0], w[1], wh, [getattr(self, f"Dcov{i}").bias.detach() for i in range(1, 6)],
                                               self._packs.conv(self.conv.weight), self.conv.bias.detach(), out, ld_dst, dst_coff, B, H, W,
                                               flags=getattr(self, "_df_flags", None) if getattr(self, "_df_flags", None) is not None
                                               and self._df_flags.device == buf.device and self._df_flags.numel() == self._df_words(B, H) else None)
            return out
        if partials is not None and DRDB.MODE != "pull" and self.in_ch == 64 and self.growth == 32:
            {"push": self._growth_push, "pair": self._growth_pair}.get(DRDB.MODE, self._growth_hybrid)(buf, partials, B, H, W)
            cin = self.total
        for i in range(1, 6) if cin == self.in_ch else ():
            cv = getattr(self, f"Dcov{i}")
            ops.conv(buf, self._packs.conv(cv.weight), cv.bias.detach(), B=B, H=H, W=W, Cin=cin, ld_src=ld, KH=3, KW=3,
                     pad=2, dil=2, Cout=self.growth, act=ACT_RELU, out=buf.view(-1, ld), ld_dst=ld, dst_coff=cin)
            cin += self.growth
        return ops.conv(buf, self._packs.conv(self.conv.weight), self.conv.bias.detach(), B=B, H=H, W=W, Cin=cin,
                        ld_src=ld, Cout=self.in_ch, act=ACT_RELU, residual=buf.view(-1, ld), ld_res=ld, res_coff=0,
                        out=out, ld_dst=ld_dst, dst_coff=dst_coff)

    def forward(self, x):
        _no_autograd(self, x)
        B, C, H, W = x.shape
        if _strict.is_strict() and self.in_ch == 64 and self.growth == 32:
            rows = _strict.nchw_to_rows_f32(x)
            g = _strict.Planes(B * H * W, self.total, x.device)
            _strict.split(rows, planes=g)
            y, _ = _strict.drdb(self, g, rows, B, H, W, want_planes=False)
            return ops.nhwc_to_nchw(y, B, H * W, C).view(B, C, H, W)
        buf = torch.empty((B, H, W, self.total), dtype=torch.bfloat16, device=x.device)
        ops.nchw_to_nhwc(x.float().contiguous(), out=buf.view(B, H * W, self.total), ld_dst=self.total)
        part = torch.empty((B, H, W, 128), dtype=torch.bfloat16, device=x.device)
        y = self.forward_buffer(buf, B, H, W, partials=part)
        return ops.nhwc_to_nchw(y, B, H * W, C).view(B, C, H, W)


# ------------------------------------------------------------------------------------------------ HIA
def _cross_ctx(mod, feats, kv_weights):
    """Per-head 8x8 contexts of up to three token streams (None entries are skipped): k^T v = Wk (P^T P) Wv^T because the kv
    Linears carry no bias (core/model_fusion.py:251,291).  Returns ctx fp32 [B, 3, 8, 8, 8]."""
    from .. import _lib
    if mod.dim != 64 or mod.num_heads != 8 or any(w.shape != (128, 64) for w in kv_weights) or mod.kv_has_bias():
        raise NotImplementedError("segmif_b200: the cross-attention kernels are specialised for dim=64, 8 heads, qkv_bias=False")
    live = [f for f in feats if f is not None]
    B, N, C = live[0].shape
    dev = live[0].device
    nchunk = max(1, min(296 // max(B, 1), (N + 511) // 512))
    partials = torch.zeros((3, B, nchunk, 64, 64), dtype=torch.float64, device=dev)
    st = ops._prep(live[0])
    for s, f in enumerate(feats):
        if f is not None:
            f = f.float().contiguous()
            _lib.call("segmif_gram64_f64", _strict._p(f), C, 0, B, N, 0, _strict._p(partials[s]), nchunk, st)
    wkv = torch.stack([w.detach().float() for w in kv_weights]).contiguous()
    wend = torch.zeros((2, 64, 128), dtype=torch.float32, device=dev)
    folded = torch.empty((B, 4, 64, 64), dtype=torch.float32, device=dev)
    ctx = torch.empty((B, 3, 8, 8, 8), dtype=torch.float32, device=dev)
    _lib.call("segmif_ffm_ctx_f64_fwd", _strict._p(partials), nchunk, _strict._p(wkv), _strict._p(wend), _strict._p(folded),
              _strict._p(ctx), B, st)
    return ctx


def _apply_ctx(q, ctx):
    """(q per head) @ ctx: q fp32 [B, N, 64], ctx [B, 8, 8, 8] (head, i, j) -> [B, N, 64]; one tensor-core product per image
    with the block-diagonal 64x64 matrix W[h8+j, h8+i] = ctx[h, i, j] (index shuffling only on the host side)."""
    B, N, C = q.shape
    W = torch.zeros((B, 8, 8, 8, 8), dtype=torch.float32, device=q.device)            # [b, h_out, j, h_in, i]
    for h in range(8):
        W[:, h, :, h, :] = ctx[:, h].transpose(-1, -2)
    wp = _strict.pack_rows(W.view(B * 64, 64)).view(B, 64, 3, 64)
    q2 = q.float().contiguous().view(B * N, C)
    planes = _strict.split(q2)
    out = torch.empty((B * N, C), dtype=torch.float32, device=q.device)
    for b in range(B):
        _strict.gemm(planes, 64, wp[b], 64, row0=b * N, rows=N, dst=out)
    return out.view(B, N, C)


class _KvBiasMixin:
    def kv_has_bias(self):
        return any(getattr(self, n).bias is not None for n in ("kv1", "kv2", "kv3") if hasattr(self, n))


class CrossAttention(_KvBiasMixin, nn.Module):
    """MoAM parameters (core/model_fusion.py:250-262); the computation lives in CrossPath's fused kernels."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None):
        super().__init__()
        assert dim % num_heads == 0, f"dim {dim} should be divided by num_heads {num_heads}."
        self.dim = dim
        self.num_heads = num_heads
        self.scale = qk_scale or (dim // num_heads) ** -0.5
        self.kv3 = nn.Linear(dim, dim * 2, bias=qkv_bias)

    def forward(self, x1, x2, segfeature):
        """core/model_fusion.py:263-288 on fp32 tokens [B, N, 64]: ctx3 = softmax_{dim=-2}(k3^T v3 * scale) from
        kv3(segfeature); returns (q1 @ ctx3, q2 @ ctx3).  Built from the strict-precision kernels (fp64 Gram / context,
        split-bf16 tensor-core product); CrossPath uses the fused kernels instead."""
        ctx = _cross_ctx(self, [None, None, segfeature], [self.kv3.weight] * 3)
        return _apply_ctx(x1, ctx[:, 2]), _apply_ctx(x2, ctx[:, 2])


class CrossAttention2(_KvBiasMixin, nn.Module):
    """SoAM parameters (core/model_fusion.py:290-302)."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None):
        super().__init__()
        assert dim % num_heads == 0, f"dim {dim} should be divided by num_heads {num_heads}."
        self.dim = dim
        self.num_heads = num_heads
        self.scale = qk_scale or (dim // num_heads) ** -0.5
        self.kv1 = nn.Linear(dim, dim * 2, bias=qkv_bias)
        self.kv2 = nn.Linear(dim, dim * 2, bias=qkv_bias)

    def forward(self, x1, x2, segfeature):
        """core/model_fusion.py:303-328: ctx_i = softmax_{dim=-2}(k_i^T v_i * scale) from kv_i(x_i); returns
        (q3 @ ctx1, q3 @ ctx2) with q3 = segfeature."""
        ctx = _cross_ctx(self, [x1, x2, None], [self.kv1.weight, self.kv2.weight, self.kv1.weight])
        return _apply_ctx(segfeature, ctx[:, 0]), _apply_ctx(segfeature, ctx[:, 1])


class CrossPath(nn.Module):
    """core/model_fusion.py:329-361."""

    def __init__(self, dim, reduction=1, num_heads=8, norm_layer=nn.LayerNorm):
        super().__init__()
        if dim != 64 or reduction != 1 or num_heads != 8:
            raise NotImplementedError("segmif_b200: the HIA kernels are specialised for dim=64, 8 heads (the only "
                                      "configuration Fusion_Network3_ac instantiates)")
        self.channel_proj1 = nn.Linear(dim, dim // reduction * 2)
        self.channel_proj2 = nn.Linear(dim, dim // reduction * 2)
        self.channel_proj3 = nn.Linear(dim, dim // reduction * 2)
        self.act1 = nn.ReLU(inplace=True)
        self.act2 = nn.ReLU(inplace=True)
        self.act3 = nn.ReLU(inplace=True)
        self.cross_attn = CrossAttention(dim // reduction, num_heads=num_heads)
        self.cross_attn2 = CrossAttention2(dim // reduction, num_heads=num_heads)
        self.end_proj1 = nn.Linear(dim // reduction * 2, dim)
        self.end_proj2 = nn.Linear(dim // reduction * 2, dim)
        self.norm1 = norm_layer(dim)
        self.norm2 = norm_layer(dim)
        self._packs = PackCache()

    def packs(self, pre_conv=None):
        """Kernel operand packs; `pre_conv` (a 1x1 nn.Conv2d applied to the segmentation features before this
        module, i.e. conv3 / conv4) is folded into channel_proj3."""
        plist = [self.channel_proj1.weight, self.channel_proj1.bias, self.channel_proj2.weight,
                 self.channel_proj2.bias, self.channel_proj3.weight, self.channel_proj3.bias,
                 self.cross_attn2.kv1.weight, self.cross_attn2.kv2.weight, self.cross_attn.kv3.weight,
                 self.end_proj1.weight, self.end_proj1.bias, self.end_proj2.weight, self.end_proj2.bias,
                 self.norm1.weight, self.norm1.bias, self.norm2.weight, self.norm2.bias]
        if pre_conv is not None:
            plist += [pre_conv.weight, pre_conv.bias]

        def build(w1, b1, w2, b2, w3, b3, kv1, kv2, kv3, we1, be1, we2, be2, g1, n1, g2, n2, wc=None, bc=None):
            f = lambda t: t.detach().float()
            w3e, b3e = f(w3), f(b3)
            if wc is not None:
                wcm = f(wc).reshape(wc.shape[0], -1)                      # [64, Cin]
                b3e = w3e @ f(bc) + b3e
                w3e = w3e @ wcm                                           # [128, Cin]
            bf = lambda *ts: torch.cat([t.reshape(-1) for t in ts]).to(torch.bfloat16).contiguous()
            fl = lambda *ts: torch.cat([t.reshape(-1) for t in ts]).float().contiguous()
            return dict(
                C3=w3e.shape[1],
                w_gram=bf(f(w1)[:64], f(w2)[:64], w3e[64:]), b_gram=fl(f(b1)[:64], f(b2)[:64], b3e[64:]),
                w_apply=bf(w3e[:64], f(w1)[64:], f(w2)[64:]), b_apply=fl(b3e[:64], f(b1)[64:], f(b2)[64:]),
                wkv=torch.stack([f(kv1), f(kv2), f(kv3)]).contiguous(),
                wend=torch.stack([f(we1), f(we2)]).contiguous(), bend=fl(f(be1), f(be2)),
                ln_g=fl(f(g1), f(g2)), ln_b=fl(f(n1), f(n2)))
        return self._packs.get_multi(plist, build, "ffm" if pre_conv is None else f"ffm+{id(pre_conv)}")

    def packs_lr(self, pre_conv):
        """Operand packs for the low-resolution seg path: Q = (channel_proj3 o pre_conv)(f) is evaluated at the encoder's
        resolution with q_w / q_b; the FFM kernels keep only the projections of the two image streams."""
        pk = self.packs(pre_conv)

        def build(wg, bg, wa, ba):
            c3 = pk["C3"]
            w3u, w3y = wg[2 * 4096:].reshape(64, c3), wa[:64 * c3].reshape(64, c3)
            return dict(q_w=torch.cat([w3y, w3u], 0).reshape(128, 1, c3).contiguous(),
                        q_b=torch.cat([ba[:64], bg[128:]]).contiguous(),
                        w_gram=wg[:2 * 4096].contiguous(), b_gram=bg[:128].contiguous(),
                        w_apply=wa[64 * c3:].contiguous(), b_apply=ba[64:].contiguous())
        lr = self._packs.get_multi([pk["w_gram"], pk["b_gram"], pk["w_apply"], pk["b_apply"]], build, f"lr+{id(pre_conv)}")
        out = dict(pk)
        out.update(lr)
        return out

    def forward(self, x1, x2, segfeature):
        """Token interface of the reference: three [B, N, 64] tensors -> two [B, N, 64] tensors (fp32)."""
        _no_autograd(self, x1, x2, segfeature)
        B, N, C = x1.shape
        cvt = lambda t: ops.nchw_to_nhwc(t.float().contiguous().view(1, 1, -1), out_dtype=torch.bfloat16).view(B, N, -1)
        t1, t2, t3 = cvt(x1), cvt(x2), cvt(segfeature)
        o1 = torch.empty((B, N, C), dtype=torch.bfloat16, device=x1.device)
        o2 = torch.empty_like(o1)
        pk = self.packs()
        ops.ffm(t1, C, 0, t2, C, 0, t3, t3.shape[-1], pk["C3"], pk, o1, C, 0, o2, C, 0, B, N)
        return o1.float(), o2.float()


class FeatureFusionModule(nn.Module):
    """core/model_fusion.py:430-463."""

    def __init__(self, dim, reduction=1, num_heads=8, norm_layer=nn.BatchNorm2d):
        super().__init__()
        self.cross = CrossPath(dim=dim, reduction=reduction, num_heads=num_heads)
        self.apply(_reference_init)

    def forward_pixel_major(self, x1, ld1, x2, ld2, seg, ld3, out1, ldo1, coffo1, out2, ldo2, coffo2, B, HW,
                            pre_conv=None):
        pk = self.cross.packs(pre_conv)
        return ops.ffm(x1, ld1, 0, x2, ld2, 0, seg, ld3, pk["C3"], pk, out1, ldo1, coffo1, out2, ldo2, coffo2, B, HW)

    def forward_pixel_major_lr(self, x1, ld1, x2, ld2, seg_tokens, qh, qw, out1, ldo1, coffo1, out2, ldo2, coffo2, B,
                               H, W, pre_conv):
        """Same as forward_pixel_major with the segmentation features given at the encoder's resolution
        (bf16 tokens [B, qh*qw, Cin]); `pre_conv` is conv3 / conv4."""
        pk = self.cross.packs_lr(pre_conv)
        q = ops.linear(seg_tokens, pk["q_w"], pk["q_b"])                      # [B*qh*qw, 128] bf16, L2 resident
        return ops.ffm_lr(x1, ld1, 0, x2, ld2, 0, q, qh, qw, H, W, pk, out1, ldo1, coffo1, out2, ldo2, coffo2, B)

    def forward(self, x1, x2, segfeature):
        _no_autograd(self, x1, x2, segfeature)
        B, C, H, W = x1.shape
        if _strict.is_strict():
            r1, r2, r3 = (_strict.nchw_to_rows_f32(t) for t in (x1, x2, segfeature))
            o1, o2, _ = _strict.ffm_full(self, r1, _strict.split(r1), r2, _strict.split(r2), r3, None, B, H, W)
            return tuple(ops.nhwc_to_nchw(o, B, H * W, C).view(B, C, H, W) for o in (o1, o2))
        t1, t2, t3 = _pixel_major_bf16(x1), _pixel_major_bf16(x2), _pixel_major_bf16(segfeature)
        o1 = torch.empty((B, H * W, C), dtype=torch.bfloat16, device=x1.device)
        o2 = torch.empty_like(o1)
        self.forward_pixel_major(t1, C, t2, C, t3, t3.shape[-1], o1, C, 0, o2, C, 0, B, H * W)
        nchw = lambda t: ops.nhwc_to_nchw(t, B, H * W, C).view(B, C, H, W)
        return nchw(o1), nchw(o2)


# ------------------------------------------------------------------------------------------------ fusion net
class Fusion_Network3_ac(nn.Module):
    """core/model_fusion.py:1026-1067."""

    def __init__(self, in_ch1=64, in_ch2=128):
        """`in_ch1` / `in_ch2`: channels of the two encoder feature maps fed to conv3 / conv4.  The reference hard-codes
        64 / 128 (core/model_fusion.py:1041-1042), which fits every backbone except mit_b0 (32 / 64 channels,
        core/mix_transformer.py:392) -- the backbone BASELINE configs[0] names; SURVEY.md 8(c) asks for the kwargs."""
        super().__init__()
        self.in_ch1, self.in_ch2 = in_ch1, in_ch2
        self.conv1_ir = nn.Conv2d(1, 64, 3, padding=1)
        self.conv1_vis = nn.Conv2d(1, 64, 3, padding=1)
        self.DRDB1 = DRDB(in_ch=64)
        self.DRDB2 = DRDB(in_ch=64)
        self.DRDB3 = DRDB(in_ch=64)
        self.DRDB4 = DRDB(in_ch=64)
        self.conv2 = nn.Conv2d(128, 64, 3, padding=1)
        self.relu = nn.PReLU()                 # ONE shared scalar slope for all five call sites (:1038)
        self.ffm = FeatureFusionModule(64)
        self.ffm2 = FeatureFusionModule(64)    # allocated and saved by the reference, never used (:1040)
        self.conv3 = nn.Conv2d(in_ch1, 64, 1, padding=0)
        self.conv4 = nn.Conv2d(in_ch2, 64, 1, padding=0)
        self.conv21 = nn.Conv2d(64, 32, 3, padding=1)
        self.conv22 = nn.Conv2d(32, 1, 3, padding=1)
        self._packs = PackCache()

    def forward(self, ir, vis, out1, out2):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            # training (train.py:360,381-386): forward that keeps what the hand-written backward needs
            from .fusion_train import forward_with_grad
            return forward_with_grad(self, ir, vis, out1, out2)
        _no_autograd(self, ir, vis, out1, out2)
        if _strict.is_strict():
            return _strict.fusion_network(self, ir, vis, ("full", _strict.nchw_to_rows_f32(out1)),
                                          ("full", _strict.nchw_to_rows_f32(out2)))
        return self._run(ir, vis, ("full", _pixel_major_bf16(out1)), ("full", _pixel_major_bf16(out2)))

    def forward_lowres(self, ir, vis, stage1, stage2):
        """Extension: same result as forward(ir, vis, upsample(f1), upsample(f2)) with the two encoder maps passed at
        their own resolution as (tokens bf16 [B, h*w, C], h, w) -- what MixVisionTransformer.forward_stages returns.
        The 1x1 convs, channel_proj3 and the bilinear resize commute (all linear), so the full-resolution feature
        maps are never written; see csrc/ffm.cu."""
        _no_autograd(self, ir, vis)
        if _strict.is_strict():
            f32 = lambda st: ("lowres", st[0].float(), st[1], st[2])
            return _strict.fusion_network(self, ir, vis, f32(stage1), f32(stage2))
        return self._run(ir, vis, ("lowres",) + tuple(stage1), ("lowres",) + tuple(stage2))

    def _ffm(self, x1, x2, seg, out1, ldo1, coffo1, out2, ldo2, coffo2, B, H, W, pre_conv):
        if seg[0] == "lowres":
            _, tok, qh, qw = seg
            self.ffm.forward_pixel_major_lr(x1, 64, x2, 64, tok, qh, qw, out1, ldo1, coffo1, out2, ldo2, coffo2, B, H, W, pre_conv)
        else:
            t = seg[1]
            if t.shape[-1] not in (64, 128):      # mit_b0's 32-channel map: the 1x1 conv as its own GEMM, then the 64-channel kernels
                t = ops.linear(t, self._packs.conv(pre_conv.weight), pre_conv.bias.detach()).view(B, H * W, 64)
                pre_conv = None
            self.ffm.forward_pixel_major(x1, 64, x2, 64, t, t.shape[-1], out1, ldo1, coffo1, out2, ldo2, coffo2, B, H * W,
                                         pre_conv=pre_conv)

    def _run(self, ir, vis, seg1, seg2):
        B, _, H, W = ir.shape
        dev = ir.device
        alpha = self.relu.weight.detach()
        ir, vis = ir.float(), vis.float()      # no-ops for fp32 inputs; channel 0 is read through the batch stride
        G = DRDB.GROWTH_LD
        buf1 = torch.empty((B, H, W, G), dtype=torch.bfloat16, device=dev)
        buf2 = torch.empty((B, H, W, G), dtype=torch.bfloat16, device=dev)
        # x = PReLU(conv1(channel 0)) written straight into the DRDB growth buffers
        ops.conv3x3_in1(ir, self._packs.taps_f32(self.conv1_ir.weight), self.conv1_ir.bias.detach(), alpha, buf1, G, 0, 64)
        ops.conv3x3_in1(vis, self._packs.taps_f32(self.conv1_vis.weight), self.conv1_vis.bias.detach(), alpha, buf2, G, 0, 64)
        part = torch.empty((B, H, W, 128), dtype=torch.bfloat16, device=dev)   # P2..P5 scratch, reused by all four DRDBs
        x1 = self.DRDB1.forward_buffer(buf1, B, H, W, partials=part)    # [B*HW, 64] bf16
        x2 = self.DRDB2.forward_buffer(buf2, B, H, W, partials=part)
        # ffm(x1, x2, conv3(out1)) -> inputs of DRDB3 / DRDB4 (channels 0..63 of the growth buffers)
        self._ffm(x1, x2, seg1, buf1, G, 0, buf2, G, 0, B, H, W, self.conv3)
        x1 = self.DRDB3.forward_buffer(buf1, B, H, W, out=x1, ld_dst=64, partials=part)
        x2 = self.DRDB4.forward_buffer(buf2, B, H, W, out=x2, ld_dst=64, partials=part)
        # second pass of the SAME ffm with conv4(out2); outputs land side by side = torch.cat([x1, x2], 1)
        cat = torch.empty((B, H, W, 128), dtype=torch.bfloat16, device=dev)
        self._ffm(x1, x2, seg2, cat, 128, 0, cat, 128, 64, B, H, W, self.conv4)
        f = ops.conv(cat, self._packs.conv(self.conv2.weight), self.conv2.bias.detach(), B=B, H=H, W=W, Cin=128, KH=3,
                     KW=3, pad=1, Cout=64, act=ACT_PRELU, prelu_alpha=alpha)
        f = ops.conv(f, self._packs.conv(self.conv21.weight), self.conv21.bias.detach(), B=B, H=H, W=W, Cin=64, KH=3,
                     KW=3, pad=1, Cout=32, act=ACT_PRELU, prelu_alpha=alpha)
        return ops.conv3x3_out1(f, self._packs.taps_f32(self.conv22.weight), self.conv22.bias.detach(), alpha, B, H, W, 32)


class Mean(nn.Module):
    """core/model_fusion.py:184-214 (imported by train.py:18): recompose RGB from a given Y plane, clamp,
    global min-max renormalisation.  Not on the hot path; the renormalisation uses torch reductions."""

    def __init__(self):
        super().__init__()
        self.fusion_nums = 2
        self.seg_nums = 2
        self.fusion_channel = 48
        self.seg_channel = 64
        self.mean = IMAGENET_MEAN
        self.std = IMAGENET_STD

    def forward(self, mask, vis):
        rgb = ops.recompose_rgb(mask[:, 0:1].float().contiguous(), vis.float().contiguous(), clamp=True)
        lo, hi = torch.aminmax(rgb)
        return (rgb - lo) / (hi - lo)


# the paper's ablation networks (core/model_fusion.py:363-425, :465-523, :626-1025) live in ablation.py; imported last because
# they build on the classes above
from .ablation import (AttentionModule, CrossPath_M, CrossPath_S, FeatureFusionModule_MoAM, FeatureFusionModule_SoAM,  # noqa: E402,F401
                       Fusion_Network3, Fusion_Network3_Add, Fusion_Network3_Average, Fusion_Network3_Con, Fusion_Network3_M,
                       Fusion_Network3_S, Fusion_Network_rmseg, Fusion_Network_rmseg_att)
