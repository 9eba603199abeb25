"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the SegMiF hot path (the oracle).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this file; the product package `segmif_b200` never does (it fails
loudly when its CUDA library is missing instead of falling back to anything here).

Every function below restates, in functional form over a plain ``{name: tensor}``
state dict, what one reference module computes, and cites the reference file:line
it follows (paths relative to the SegMiF repository root).  Arithmetic is done with
torch CPU tensor ops in the dtype of the inputs (fp32 by default, fp64 for tighter
checks) -- the reference itself is PyTorch, so this is the "port" flavour of oracle.

Pinning: the reference ships no golden vectors, KATs or fixtures,
so the oracle is pinned against *outputs of the reference itself run in the build
container*: `oracle/make_golden.py` imports the unmodified reference modules
(`oracle/ref_shim.py`), runs them on seeded synthetic weights/inputs and stores the
results under tests/golden/; `tests/test_oracle_golden.py` checks this file against
those fixtures on every CPU test run, and against the loss known-answer values
recorded in SURVEY.md Appendix C.
"""
import math

import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------- helpers


def _sub(sd, prefix):
    """View of a state dict below `prefix` (keys with the prefix stripped)."""
    p = prefix + "."
    return {k[len(p):]: v for k, v in sd.items() if k.startswith(p)}


def _linear(x, sd, name):
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def layer_norm(x, sd, name, eps):
    # nn.LayerNorm over the last dim, biased variance.
    c = x.shape[-1]
    return F.layer_norm(x, (c,), sd[name + ".weight"], sd[name + ".bias"], eps)


# ----------------------------------------------------------------------------- MiT encoder

MIT_CONFIGS = {
    # core/mix_transformer.py:389-434
    "mit_b0": dict(embed_dims=[32, 64, 160, 256], depths=[2, 2, 2, 2]),
    "mit_b1": dict(embed_dims=[64, 128, 320, 512], depths=[2, 2, 2, 2]),
    "mit_b2": dict(embed_dims=[64, 128, 320, 512], depths=[3, 4, 6, 3]),
    "mit_b3": dict(embed_dims=[64, 128, 320, 512], depths=[3, 4, 18, 3]),
    "mit_b4": dict(embed_dims=[64, 128, 320, 512], depths=[3, 8, 27, 3]),
    "mit_b5": dict(embed_dims=[64, 128, 320, 512], depths=[3, 6, 40, 3]),
}
MIT_HEADS = [1, 2, 5, 8]
MIT_SR = [8, 4, 2, 1]
BLOCK_LN_EPS = 1e-6    # partial(nn.LayerNorm, eps=1e-6), mix_transformer.py:393
DEFAULT_LN_EPS = 1e-5  # nn.LayerNorm default: patch-embed norm, sr norm, CrossPath norms


def overlap_patch_embed(x, sd, name, patch, stride):
    """core/mix_transformer.py:192-198 -- strided conv, flatten to tokens, LayerNorm(1e-5)."""
    y = F.conv2d(x, sd[name + ".proj.weight"], sd[name + ".proj.bias"], stride=stride, padding=patch // 2)
    _, _, h, w = y.shape
    tok = y.flatten(2).transpose(1, 2)
    return layer_norm(tok, sd, name + ".norm", DEFAULT_LN_EPS), h, w


def sr_attention(x, h, w, sd, name, heads, sr):
    """core/mix_transformer.py:94-115 -- spatial-reduction self attention."""
    b, n, c = x.shape
    d = c // heads
    q = _linear(x, sd, name + ".q").reshape(b, n, heads, d).permute(0, 2, 1, 3)
    src = x
    if sr > 1:
        img = x.permute(0, 2, 1).reshape(b, c, h, w)
        red = F.conv2d(img, sd[name + ".sr.weight"], sd[name + ".sr.bias"], stride=sr)
        src = layer_norm(red.reshape(b, c, -1).permute(0, 2, 1), sd, name + ".norm", DEFAULT_LN_EPS)
    kv = _linear(src, sd, name + ".kv").reshape(b, -1, 2, heads, d).permute(2, 0, 3, 1, 4)
    k, v = kv[0], kv[1]
    att = (q @ k.transpose(-2, -1)) * (d ** -0.5)
    att = att.softmax(dim=-1)
    out = (att @ v).transpose(1, 2).reshape(b, n, c)
    return _linear(out, sd, name + ".proj")


def mix_ffn(x, h, w, sd, name):
    """core/mix_transformer.py:46-53,381-387 -- fc1 -> depthwise 3x3 -> GELU(erf) -> fc2."""
    b, n, _ = x.shape
    y = _linear(x, sd, name + ".fc1")
    hid = y.shape[-1]
    img = y.transpose(1, 2).reshape(b, hid, h, w)
    img = F.conv2d(img, sd[name + ".dwconv.dwconv.weight"], sd[name + ".dwconv.dwconv.bias"], padding=1, groups=hid)
    y = F.gelu(img.flatten(2).transpose(1, 2))
    return _linear(y, sd, name + ".fc2")


def mit_block(x, h, w, sd, name, heads, sr, droppath=None):
    """core/mix_transformer.py:151-155.  Eval: DropPath is the identity; `droppath` = (s1, s2), two per-sample scale
    vectors [B] (0 or 1/keep_prob), reproduces a train-mode draw of timm's DropPath (:129,152-153)."""
    a = sr_attention(layer_norm(x, sd, name + ".norm1", BLOCK_LN_EPS), h, w, sd, name + ".attn", heads, sr)
    x = x + (a if droppath is None else a * droppath[0].view(-1, 1, 1))
    m = mix_ffn(layer_norm(x, sd, name + ".norm2", BLOCK_LN_EPS), h, w, sd, name + ".mlp")
    x = x + (m if droppath is None else m * droppath[1].view(-1, 1, 1))
    return x


def mit_forward_features(x, sd, backbone, droppath=None):
    """core/mix_transformer.py:312-348 -- returns the four NCHW stage outputs.  `droppath`: optional list with one
    (s1, s2) pair per block (train-mode stochastic depth with the masks given)."""
    cfg = MIT_CONFIGS[backbone]
    outs = []
    b = x.shape[0]
    bi = 0
    for s in range(4):
        patch, stride = (7, 4) if s == 0 else (3, 2)
        tok, h, w = overlap_patch_embed(x, sd, f"patch_embed{s + 1}", patch, stride)
        for i in range(cfg["depths"][s]):
            tok = mit_block(tok, h, w, sd, f"block{s + 1}.{i}", MIT_HEADS[s], MIT_SR[s], None if droppath is None else droppath[bi])
            bi += 1
        tok = layer_norm(tok, sd, f"norm{s + 1}", BLOCK_LN_EPS)
        x = tok.reshape(b, h, w, -1).permute(0, 3, 1, 2).contiguous()
        outs.append(x)
    return outs


def mit_forward_fusion(x, sd, backbone):
    """core/mix_transformer.py:358-375 -- stage-1/2 maps bilinearly upsampled to the input size."""
    hh, ww = x.shape[2:]
    outs = mit_forward_features(x, sd, backbone)
    up = lambda t: F.interpolate(t, size=[hh, ww], mode="bilinear", align_corners=False)
    return up(outs[0]), up(outs[1])


# ----------------------------------------------------------------------------- SegFormer head


def segformer_head(feats, sd, bn_eps=1e-5, train_bn=False, dropout_scale=None):
    """core/segformer_head.py:59-82.  Default = eval mode: BN uses running stats, Dropout2d is identity.  train_bn=True
    uses batch statistics (nn.BatchNorm2d in training mode; running stats are not touched here) and `dropout_scale`
    [B, E] (0 or 1/(1-p) per sample and channel) reproduces a train-mode draw of Dropout2d (:57,79)."""
    c1, c2, c3, c4 = feats
    n = c1.shape[0]
    size = c1.shape[2:]
    parts = []
    for name, c in (("linear_c4", c4), ("linear_c3", c3), ("linear_c2", c2), ("linear_c1", c1)):
        y = _linear(c.flatten(2).transpose(1, 2), sd, name + ".proj")          # segformer_head.py:21-24
        y = y.permute(0, 2, 1).reshape(n, -1, c.shape[2], c.shape[3])
        if c is not c1:
            y = F.interpolate(y, size=size, mode="bilinear", align_corners=False)
        parts.append(y)
    cat = torch.cat(parts, dim=1)
    y = F.conv2d(cat, sd["linear_fuse.conv.weight"])                           # bias=False: BN follows
    if train_bn:
        y = F.batch_norm(y, None, None, sd["linear_fuse.bn.weight"], sd["linear_fuse.bn.bias"], True, 0.0, bn_eps)
    else:
        y = F.batch_norm(y, sd["linear_fuse.bn.running_mean"], sd["linear_fuse.bn.running_var"],
                         sd["linear_fuse.bn.weight"], sd["linear_fuse.bn.bias"], False, 0.0, bn_eps)
    y = F.relu(y)
    if dropout_scale is not None:
        y = y * dropout_scale.view(y.shape[0], -1, 1, 1)
    return F.conv2d(y, sd["linear_pred.weight"], sd["linear_pred.bias"])


IMAGENET_MEAN = [123.675, 116.28, 103.53]
IMAGENET_STD = [58.395, 57.12, 57.375]


def network3_forward(x, sd, backbone, train_bn=False, dropout_scale=None, droppath=None):
    """core/model_fusion.py:1081-1088 + WeTr.forward :62-68 -- x*255, ImageNet mean/std, encoder, head.
    `sd` is the Network3 state dict (keys start with denoise_net.).  The keyword arguments select train-mode behaviour
    with injected masks (see segformer_head / mit_block)."""
    mean = torch.tensor(IMAGENET_MEAN, dtype=x.dtype).view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD, dtype=x.dtype).view(1, 3, 1, 1)
    xn = (x * 255 - mean) / std
    feats = mit_forward_features(xn, _sub(sd, "denoise_net.encoder"), backbone, droppath)
    return segformer_head(feats, _sub(sd, "denoise_net.decoder"), train_bn=train_bn, dropout_scale=dropout_scale)


def seg_labels(logits, size):
    """test_segmentation.py:170-175 -- bilinear upsample to label size, argmax over classes."""
    up = F.interpolate(logits, size=size, mode="bilinear", align_corners=False)
    return up.argmax(dim=1)


def seg_cross_entropy(logits, labels, ignore_index=255):
    """model_fusion.py:1095-1096 + train.py:156 -- upsample then CE(ignore_index=255, mean)."""
    up = F.interpolate(logits, size=labels.shape[1:], mode="bilinear", align_corners=False)
    return F.cross_entropy(up, labels.long(), ignore_index=ignore_index)


# ----------------------------------------------------------------------------- fusion network


def drdb(x, sd, name):
    """core/model_fusion.py:134-157 -- five dilated (pad 2, dil 2) growth convs + 1x1 + residual."""
    cat = x
    for i in range(1, 6):
        g = F.relu(F.conv2d(cat, sd[f"{name}.Dcov{i}.weight"], sd[f"{name}.Dcov{i}.bias"], padding=2, dilation=2))
        cat = torch.cat([cat, g], dim=1)
    return x + F.relu(F.conv2d(cat, sd[name + ".conv.weight"], sd[name + ".conv.bias"]))


def _ctx(k_tokens, v_tokens, heads):
    """softmax over dim=-2 of (k^T v) * scale, per head; model_fusion.py:281-282,316-319."""
    b, n, c = k_tokens.shape
    d = c // heads
    k = k_tokens.reshape(b, n, heads, d).permute(0, 2, 1, 3)
    v = v_tokens.reshape(b, n, heads, d).permute(0, 2, 1, 3)
    return ((k.transpose(-2, -1) @ v) * (d ** -0.5)).softmax(dim=-2)


def _apply_ctx(q_tokens, ctx):
    b, n, c = q_tokens.shape
    heads = ctx.shape[1]
    q = q_tokens.reshape(b, n, heads, c // heads).permute(0, 2, 1, 3)
    return (q @ ctx).permute(0, 2, 1, 3).reshape(b, n, c)


def cross_path(x1, x2, seg, sd, name, heads=8):
    """core/model_fusion.py:350-361 (CrossPath) with CrossAttention :263-288 (MoAM, kv from the
    segmentation stream) and CrossAttention2 :303-328 (SoAM, query from the segmentation stream)."""
    y1, u1 = F.relu(_linear(x1, sd, name + ".channel_proj1")).chunk(2, dim=-1)
    y2, u2 = F.relu(_linear(x2, sd, name + ".channel_proj2")).chunk(2, dim=-1)
    y3, u3 = F.relu(_linear(seg, sd, name + ".channel_proj3")).chunk(2, dim=-1)
    c = u1.shape[-1]
    kv3 = F.linear(u3, sd[name + ".cross_attn.kv3.weight"])
    ctx3 = _ctx(kv3[..., :c], kv3[..., c:], heads)
    v1, v2 = _apply_ctx(u1, ctx3), _apply_ctx(u2, ctx3)
    kv1 = F.linear(y1, sd[name + ".cross_attn2.kv1.weight"])
    kv2 = F.linear(y2, sd[name + ".cross_attn2.kv2.weight"])
    z1 = _apply_ctx(y3, _ctx(kv1[..., :c], kv1[..., c:], heads))
    z2 = _apply_ctx(y3, _ctx(kv2[..., :c], kv2[..., c:], heads))
    o1 = layer_norm(x1 + _linear(torch.cat((z1, v1), -1), sd, name + ".end_proj1"), sd, name + ".norm1", DEFAULT_LN_EPS)
    o2 = layer_norm(x2 + _linear(torch.cat((z2, v2), -1), sd, name + ".end_proj2"), sd, name + ".norm2", DEFAULT_LN_EPS)
    return o1, o2


def feature_fusion_module(x1, x2, seg, sd, name):
    """core/model_fusion.py:453-463 -- NCHW -> tokens -> CrossPath -> NCHW."""
    b, c, h, w = x1.shape
    tok = lambda t: t.flatten(2).transpose(1, 2)
    o1, o2 = cross_path(tok(x1), tok(x2), tok(seg), sd, name + ".cross")
    img = lambda t: t.reshape(b, h, w, -1).permute(0, 3, 1, 2).contiguous()
    return img(o1), img(o2)


def fusion_network3_ac(ir, vis, out1, out2, sd):
    """core/model_fusion.py:1047-1067 -- one shared scalar PReLU, `ffm` applied twice, `ffm2` unused."""
    a = sd["relu.weight"]
    x1 = drdb(F.prelu(F.conv2d(ir[:, 0:1], sd["conv1_ir.weight"], sd["conv1_ir.bias"], padding=1), a), sd, "DRDB1")
    x2 = drdb(F.prelu(F.conv2d(vis[:, 0:1], sd["conv1_vis.weight"], sd["conv1_vis.bias"], padding=1), a), sd, "DRDB2")
    x1, x2 = feature_fusion_module(x1, x2, F.conv2d(out1, sd["conv3.weight"], sd["conv3.bias"]), sd, "ffm")
    x1, x2 = drdb(x1, sd, "DRDB3"), drdb(x2, sd, "DRDB4")
    x1, x2 = feature_fusion_module(x1, x2, F.conv2d(out2, sd["conv4.weight"], sd["conv4.bias"]), sd, "ffm")
    f = F.prelu(F.conv2d(torch.cat([x1, x2], 1), sd["conv2.weight"], sd["conv2.bias"], padding=1), a)
    f = F.prelu(F.conv2d(f, sd["conv21.weight"], sd["conv21.bias"], padding=1), a)
    return F.prelu(F.conv2d(f, sd["conv22.weight"], sd["conv22.bias"], padding=1), a)


# ----------------------------------------------------------------------------- ablation networks


def cross_path_variant(x1, x2, seg, sd, name, mode, heads=8):
    """core/model_fusion.py:350-361 (mode 'full'), :384-395 ('M': MoAM only), :417-428 ('S': SoAM only), any dim."""
    y1, u1 = F.relu(_linear(x1, sd, name + ".channel_proj1")).chunk(2, dim=-1)
    y2, u2 = F.relu(_linear(x2, sd, name + ".channel_proj2")).chunk(2, dim=-1)
    y3, u3 = F.relu(_linear(seg, sd, name + ".channel_proj3")).chunk(2, dim=-1)
    c = u1.shape[-1]
    parts1, parts2 = [], []
    if mode in ("full", "S"):
        kv1 = F.linear(y1, sd[name + ".cross_attn2.kv1.weight"])
        kv2 = F.linear(y2, sd[name + ".cross_attn2.kv2.weight"])
        parts1.append(_apply_ctx(y3, _ctx(kv1[..., :c], kv1[..., c:], heads)))
        parts2.append(_apply_ctx(y3, _ctx(kv2[..., :c], kv2[..., c:], heads)))
    if mode in ("full", "M"):
        kv3 = F.linear(u3, sd[name + ".cross_attn.kv3.weight"])
        ctx3 = _ctx(kv3[..., :c], kv3[..., c:], heads)
        parts1.append(_apply_ctx(u1, ctx3))
        parts2.append(_apply_ctx(u2, ctx3))
    o1 = layer_norm(x1 + _linear(torch.cat(parts1, -1), sd, name + ".end_proj1"), sd, name + ".norm1", DEFAULT_LN_EPS)
    o2 = layer_norm(x2 + _linear(torch.cat(parts2, -1), sd, name + ".end_proj2"), sd, name + ".norm2", DEFAULT_LN_EPS)
    return o1, o2


def ffm_variant(x1, x2, seg, sd, name, mode):
    """core/model_fusion.py:453-463 / :486-494 / :516-523 -- NCHW -> tokens -> cross path variant -> NCHW."""
    b, c, h, w = x1.shape
    tok = lambda t: t.flatten(2).transpose(1, 2)
    o1, o2 = cross_path_variant(tok(x1), tok(x2), tok(seg), sd, name + ".cross", mode)
    img = lambda t: t.reshape(b, h, w, -1).permute(0, 3, 1, 2).contiguous()
    return img(o1), img(o2)


def attention_module(x, sd, name):
    """core/model_fusion.py:759-771 -- conv3x3, ReLU, conv3x3, then x * sigmoid(x)."""
    t = F.conv2d(F.relu(F.conv2d(x, sd[name + ".conv.0.weight"], sd[name + ".conv.0.bias"], padding=1)),
                 sd[name + ".conv.2.weight"], sd[name + ".conv.2.bias"], padding=1)
    return torch.sigmoid(t) * t


def fusion_network3_variant(ir, vis, out1, out2, sd, variant):
    """core/model_fusion.py:626-660 ('base'), :821-890 ('S', 'M'), :661-709 ('Con'), :710-758 ('Add'), :772-820 ('Average'):
    32-channel streams, conv21 is the 32 -> 1 output layer, one shared scalar PReLU."""
    a = sd["relu.weight"]
    cv = lambda x, n, pad: F.conv2d(x, sd[n + ".weight"], sd[n + ".bias"], padding=pad)
    x1 = drdb(F.prelu(cv(ir[:, 0:1], "conv1_ir", 1), a), sd, "DRDB1")
    x2 = drdb(F.prelu(cv(vis[:, 0:1], "conv1_vis", 1), a), sd, "DRDB2")
    s1, s2 = cv(out1, "conv3", 0), cv(out2, "conv4", 0)
    if variant in ("base", "S", "M"):
        mode = {"base": "full", "S": "S", "M": "M"}[variant]
        x1, x2 = ffm_variant(x1, x2, s1, sd, "ffm", mode)
        x1, x2 = drdb(x1, sd, "DRDB3"), drdb(x2, sd, "DRDB4")
        x1, x2 = ffm_variant(x1, x2, s2, sd, "ffm", mode)
    elif variant in ("Con", "Add"):
        mix = (lambda x, s: torch.cat([x, s], 1)) if variant == "Con" else (lambda x, s: x + s)
        x1, x2 = cv(mix(x1, s1), "conv211", 1), cv(mix(x2, s1), "conv221", 1)
        x1, x2 = drdb(x1, sd, "DRDB3"), drdb(x2, sd, "DRDB4")
        x1, x2 = cv(mix(x1, s2), "conv411", 1), cv(mix(x2, s2), "conv421", 1)
    elif variant == "Average":
        x1 = attention_module(x1, sd, "att1") + attention_module(s1, sd, "att2")
        x2 = attention_module(x2, sd, "att3") + attention_module(s1, sd, "att4")
        x1, x2 = drdb(x1, sd, "DRDB3"), drdb(x2, sd, "DRDB4")
        x1 = attention_module(x1, sd, "att5") + attention_module(s2, sd, "att6")
        x2 = attention_module(x2, sd, "att7") + attention_module(s2, sd, "att8")
    else:
        raise ValueError(variant)
    f = F.prelu(cv(torch.cat([x1, x2], 1), "conv2", 1), a)
    return F.prelu(cv(f, "conv21", 1), a)


def fusion_network_rmseg(ir, vis, sd):
    """core/model_fusion.py:938-973 -- the fusion network without segmentation features; returns (fused, [x1, x2])."""
    a = sd["relu.weight"]
    cv = lambda x, n: F.conv2d(x, sd[n + ".weight"], sd[n + ".bias"], padding=1)
    x1 = drdb(drdb(F.prelu(cv(ir[:, 0:1], "conv1_ir"), a), sd, "DRDB1"), sd, "DRDB3")
    x2 = drdb(drdb(F.prelu(cv(vis[:, 0:1], "conv1_vis"), a), sd, "DRDB2"), sd, "DRDB4")
    f = F.prelu(cv(torch.cat([x1, x2], 1), "conv2"), a)
    f = F.prelu(cv(f, "conv21"), a)
    return F.prelu(cv(f, "conv22"), a), [x1, x2]


# ----------------------------------------------------------------------------- colour transforms


def rgb2ycrcb(x):
    """core/model_fusion.py:69-92 (device-agnostic restatement, NCHW in / NCHW out)."""
    r, g, b = x[:, 0:1], x[:, 1:2], x[:, 2:3]
    y = 0.299 * r + 0.587 * g + 0.114 * b
    cr = (r - y) * 0.713 + 0.5
    cb = (b - y) * 0.564 + 0.5
    return torch.cat([y, cr, cb], dim=1)


def ycrcb2rgb(x):
    """core/model_fusion.py:94-111 -- (x + [0,-.5,-.5]) @ [[1,1,1],[1.403,-.714,0],[0,-.344,1.773]]."""
    mat = torch.tensor([[1.0, 1.0, 1.0], [1.403, -0.714, 0.0], [0.0, -0.344, 1.773]], dtype=x.dtype)
    bias = torch.tensor([0.0, -0.5, -0.5], dtype=x.dtype)
    flat = x.permute(0, 2, 3, 1).reshape(-1, 3)
    out = (flat + bias).mm(mat)
    return out.reshape(x.shape[0], x.shape[2], x.shape[3], 3).permute(0, 3, 1, 2).contiguous()


def recompose_rgb(fused_y, vis_ycrcb, clamp=True):
    """test_fusion.py:102-111 / train.py:364-366 -- replace Y by the fused image, back to RGB, clamp."""
    ycc = vis_ycrcb.clone()
    ycc[:, 0:1] = fused_y
    rgb = ycrcb2rgb(ycc)
    return rgb.clamp(0.0, 1.0) if clamp else rgb


# ----------------------------------------------------------------------------- losses


def gaussian_1d(size, sigma, dtype=torch.float32):
    """pytorch_ssim/__init__.py:8-10."""
    g = torch.tensor([math.exp(-(i - size // 2) ** 2 / float(2 * sigma ** 2)) for i in range(size)], dtype=torch.float32)
    return (g / g.sum()).to(dtype)


def ssim(img1, img2, window_size=11, size_average=True):
    """pytorch_ssim/__init__.py:19-43,70-78 -- 11x11 sigma=1.5 window, zero pad, C1=.01^2, C2=.03^2."""
    ch = img1.shape[1]
    g = gaussian_1d(window_size, 1.5).unsqueeze(1)
    win = g.mm(g.t()).float().to(img1.dtype).expand(ch, 1, window_size, window_size).contiguous()
    pad = window_size // 2
    conv = lambda t: F.conv2d(t, win, padding=pad, groups=ch)
    mu1, mu2 = conv(img1), conv(img2)
    mu1_sq, mu2_sq, mu12 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    s1 = conv(img1 * img1) - mu1_sq
    s2 = conv(img2 * img2) - mu2_sq
    s12 = conv(img1 * img2) - mu12
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    m = ((2 * mu12 + c1) * (2 * s12 + c2)) / ((mu1_sq + mu2_sq + c1) * (s1 + s2 + c2))
    return m.mean() if size_average else m.mean(1).mean(1).mean(1)


def lap_gaussian_2d(k, sigma=2.0, dtype=torch.float32):
    """lap_loss.py:39-60 -- normalised 2-D Gaussian on an integer grid, computed in fp32."""
    ax = torch.arange(k)
    xg = ax.repeat(k).view(k, k)
    yg = xg.t()
    mean = (k - 1) / 2.0
    var = sigma ** 2.0
    e = torch.exp((-((xg - mean) ** 2.0 + (yg - mean) ** 2.0) / (2 * var)).float())
    ker = (1.0 / (2.0 * math.pi * var)) * e
    return (ker / ker.sum()).to(dtype)


def lap_residuals(img):
    """lap_loss.py:74-80 with the three kernels of :88-90 -- same-resolution DoG residuals."""
    ch = img.shape[1]
    res = []
    for k in (3, 5, 7):
        ker = lap_gaussian_2d(k, 2.0, img.dtype).view(1, 1, k, k).repeat(ch, 1, 1, 1)
        res.append(img - F.conv2d(img, ker, padding=k // 2, groups=ch))
    return res


def lap_loss(inp, target):
    """lap_loss.py:93-98."""
    a, b = lap_residuals(inp), lap_residuals(target)
    return 10.0 * (F.l1_loss(a[0], b[0]) + F.l1_loss(a[1], b[1])) + F.l1_loss(a[2], b[2])


def lap_loss2(inp, ir, vis):
    """lap_loss.py:112-118 -- target is max(residual(ir), residual(vis)) per level."""
    a, b, c = lap_residuals(inp), lap_residuals(ir), lap_residuals(vis)
    l = [F.l1_loss(a[i], torch.maximum(b[i], c[i])) for i in range(3)]
    return 10.0 * (l[0] + l[1]) + l[2]


def entropy(x, patch):
    """core/Entropy.py:15-56 -- soft 32-bin histogram entropy of non-overlapping patch x patch tiles."""
    b = x.shape[0]
    cols = F.unfold(x, kernel_size=(patch, patch), stride=patch).transpose(1, 2)   # [B, L, p*p]
    vals = cols.reshape(-1, cols.shape[2])
    bins = torch.linspace(0, 1, 32).to(x.dtype)
    sigma = torch.tensor(0.01).to(x.dtype)
    kv = torch.exp(-0.5 * ((vals.unsqueeze(2) - bins.view(1, 1, -1)) / sigma).pow(2))
    pdf = kv.mean(dim=1)
    pdf = pdf / (pdf.sum(dim=1, keepdim=True) + 1e-40) + 1e-40
    return (-(pdf * torch.log(pdf)).sum(dim=1)).reshape(b, -1).sum()


def sobelxy(x):
    """core/loss.py:634-650 -- |Gx| + |Gy|, zero pad 1 (cross-correlation, as F.conv2d)."""
    kx = torch.tensor([[-1.0, 0.0, 1.0], [-2.0, 0.0, 2.0], [-1.0, 0.0, 1.0]], dtype=x.dtype).view(1, 1, 3, 3)
    ky = torch.tensor([[1.0, 2.0, 1.0], [0.0, 0.0, 0.0], [-1.0, -2.0, -1.0]], dtype=x.dtype).view(1, 1, 3, 3)
    return F.conv2d(x, kx, padding=1).abs() + F.conv2d(x, ky, padding=1).abs()


def fusionloss3(ir, vis, fused, mask):
    """core/loss.py:464-476 -- L1(mask, fused) + L1(sobel(mask), sobel(fused))."""
    m = mask[:, :1]
    return F.l1_loss(m, fused) + F.l1_loss(sobelxy(m), sobelxy(fused))


def fusionloss_grad3(ir, vis, fused, mask):
    """core/loss.py:511-517 -- MSE(mask, fused) + 1.1 * (1 - ssim(fused, mask))."""
    m = mask[:, :1]
    return F.mse_loss(m, fused) + 1.1 * (1 - ssim(fused, m))


def fusionloss_grad2(ir, vis, fused, mask):
    """core/loss.py:497-505 -- L1 + 0.1 * LapLoss2(fused, vis_y, ir) + 1.1 * (1 - ssim)."""
    m = mask[:, :1]
    return F.l1_loss(m, fused) + 0.1 * lap_loss2(fused, vis[:, :1], ir[:, :1]) + 1.1 * (1 - ssim(fused, m))


def fusionloss(ir, vis, fused):
    """core/loss.py:423-439 -- L1(max(vis_y, ir), fused) + 8 * L1(max(sobel(vis_y), sobel(ir)), sobel(fused))."""
    y, i = vis[:, :1], ir[:, :1]
    return F.l1_loss(torch.max(y, i), fused) + 8 * F.l1_loss(torch.max(sobelxy(y), sobelxy(i)), sobelxy(fused))


def fusionloss4(ir, vis, fused, mask):
    """core/loss.py:545-559 -- L1((vis_y + ir) / 2, fused) + 4 * L1(sobel((vis_y + ir) / 2), sobel(fused))."""
    syn = (vis[:, :1] + ir[:, :1]) / 2
    return F.l1_loss(syn, fused) + 4 * F.l1_loss(sobelxy(syn), sobelxy(fused))


def fusionloss_add(ir, vis, fused):
    """core/loss.py:561-577 -- 1.5 * L1(0.4 vis_y + 0.6 ir, fused) + 5 * L1(max(sobel(vis_y), sobel(ir)), sobel(fused))."""
    y, i = vis[:, :1], ir[:, :1]
    return 1.5 * F.l1_loss(y * 0.4 + i * 0.6, fused) + 5 * F.l1_loss(torch.max(sobelxy(y), sobelxy(i)), sobelxy(fused))


def new_loss_sobel(ir, vis, mask_ir, fused):
    """core/loss.py:386-399, INCLUDING the rebinding of mask_ir / mask_vis to the scalar MSE terms (:394-395) that
    then scale the Sobel maps (:396-397)."""
    mask_vis = torch.abs(1 - mask_ir)
    a = F.mse_loss(mask_ir * fused, mask_ir * ir)
    v = F.mse_loss(mask_vis * fused, mask_vis * vis)
    a2 = F.mse_loss(a * sobelxy(fused), a * sobelxy(ir))
    v2 = F.mse_loss(v * sobelxy(fused), v * sobelxy(vis))
    return (v + v2) * 1.0 + (a + a2) * 0.85


def total_fusion_loss(ir, vis, mask, fused):
    """core/loss.py:578-588."""
    y, i = vis[:, :1], ir[:, :1]
    return fusionloss(i, y, fused) * 1.2 + new_loss_sobel(i, y, mask, fused) * 0.85


def total_fusion_loss2(ir, vis, mask, fused):
    """core/loss.py:591-599."""
    return new_loss_sobel(ir[:, :1], vis[:, :1], mask, fused)


def iqa_loss(lr, vis, mask):
    """core/loss.py:610-633 -- the entropy / std softmax weights (:616-626) are computed and never used."""
    lr, vis, mask = lr[:, 0:1], vis[:, 0:1], mask[:, 0:1]
    inv = torch.abs(1 - mask)
    mse = 0.5 * F.mse_loss(lr, mask) + 0.5 * F.mse_loss(vis, inv)
    grad = 0.5 * F.mse_loss(sobelxy(lr), sobelxy(mask)) + 0.5 * F.mse_loss(sobelxy(vis), sobelxy(inv))
    return mse + grad


# ----------------------------------------------------------------------------- whole pipeline


def inference_pipeline(ir, vis_rgb, mask, seg_sd, fusion_sd, backbone, ycrcb_input=True):
    """One unit of work of the headline metric: forward_fusion(mask) ->
    Fusion_Network3_ac -> colour recompose -> Network3.forward -> upsample -> argmax.
    train.py:356-366 feeds YCrCb (channel 0 = Y) to the fusion net; test_fusion.py:101 feeds RGB."""
    enc = _sub(seg_sd, "denoise_net.encoder")
    out0, out1 = mit_forward_fusion(mask, enc, backbone)
    vis_in = rgb2ycrcb(vis_rgb) if ycrcb_input else vis_rgb
    fused = fusion_network3_ac(ir, vis_in, out0, out1, fusion_sd)
    rgb = recompose_rgb(fused, rgb2ycrcb(vis_rgb))
    logits = network3_forward(rgb, seg_sd, backbone)
    labels = seg_labels(logits, ir.shape[2:])
    return dict(out0=out0, out1=out1, fused=fused, rgb=rgb, logits=logits, labels=labels)


# ----------------------------------------------------------------------------- validation


def confusion_matrix(labels, prediction, num_classes=9):
    """test_segmentation.py:173-176 -- sklearn.metrics.confusion_matrix(y_true, y_pred, labels=[0..nc-1]): rows = truth,
    columns = prediction; pairs with a value outside `labels` (the ignore index 255) are not counted."""
    t, p = labels.reshape(-1).long(), prediction.reshape(-1).long()
    ok = (t >= 0) & (t < num_classes) & (p >= 0) & (p < num_classes)
    idx = t[ok] * num_classes + p[ok]
    return torch.bincount(idx, minlength=num_classes * num_classes).reshape(num_classes, num_classes)


def compute_results(conf_total):
    """util/util.py:31-55 (consider_unlabeled = True): precision = TP / column sum, recall = TP / row sum,
    IoU = TP / (row + column - TP); NaN where the denominator is zero."""
    import numpy as np
    conf = np.asarray(conf_total, dtype=np.float64)
    n = conf.shape[0]
    prec, rec, iou = np.zeros(n), np.zeros(n), np.zeros(n)
    for c in range(n):
        col, row = conf[:, c].sum(), conf[c, :].sum()
        prec[c] = np.nan if col == 0 else conf[c, c] / col
        rec[c] = np.nan if row == 0 else conf[c, c] / row
        iou[c] = np.nan if row + col - conf[c, c] == 0 else conf[c, c] / (row + col - conf[c, c])
    return prec, rec, iou


def fused_to_uint8(rgb):
    """val_performance.py:447-460 -- clamp to [0,1], np.uint8(255.0 * x), NCHW -> NHWC, (u - min) / (max - min) over the
    whole batch, np.uint8(255.0 * y); numpy's own dtype rules (uint8 difference, float64 quotient, truncating casts)."""
    import numpy as np
    x = rgb.clamp(0.0, 1.0).cpu().numpy()
    u = np.uint8(255.0 * x).transpose((0, 2, 3, 1))
    y = (u - np.min(u)) / (np.max(u) - np.min(u))
    return np.uint8(255.0 * y)
