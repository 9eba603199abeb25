"""Thin tensor-level wrappers over the C ABI: allocate outputs, pass raw device pointers and the
current CUDA stream.  PyTorch is used here for device memory and streams only; every computation
is a kernel in libsegmif_b200.so.  All wrappers raise on non-CUDA tensors -- there is no fallback."""
import ctypes
import os

import torch

from . import _lib
from ._lib import ACT_GELU, ACT_NONE, ACT_PRELU, ACT_RELU, BF16, F32

_DT = {torch.float32: F32, torch.bfloat16: BF16}


def _dt(t):
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError(f"segmif_b200: unsupported dtype {t.dtype}")


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _prep(*tensors):
    """Validates tensors, initialises the library for their device and returns the stream handle."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("segmif_b200: tensors must live on a CUDA (sm_100) device; there is no CPU path")
        if not t.is_contiguous():
            raise RuntimeError("segmif_b200: tensors must be contiguous")
        dev = t.device if dev is None else dev
        if t.device != dev:
            raise RuntimeError("segmif_b200: tensors on different devices")
    _lib.ensure_init(dev.index if dev.index is not None else torch.cuda.current_device())
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def layernorm(x, gamma, beta, eps, out_dtype=torch.bfloat16, out=None):
    st = _prep(x, gamma, beta, out)
    C = x.shape[-1]
    rows = x.numel() // C
    if out is None:
        out = torch.empty(x.shape, dtype=out_dtype, device=x.device)
    _lib.call("segmif_layernorm_fwd", _ptr(x), _dt(x), _ptr(gamma), _ptr(beta), _ptr(out), _dt(out), rows, C,
              float(eps), st)
    return out


def conv(src, weight, bias, *, B, H, W, Cin, KH=1, KW=1, stride=1, pad=0, dil=1, Cout, ld_src=None, src_coff=0,
         act=ACT_NONE, prelu_alpha=None, residual=None, ld_res=None, res_coff=0, out=None, out_dtype=torch.bfloat16,
         ld_dst=None, dst_coff=0, entry=None, pre_add=None, pre_coff=0):
    """Dense contraction on the tensor cores.  `src` is pixel-major bf16 with channel pitch ld_src; `weight` is the
    packed bf16 [Cout, KH*KW, Cin] tensor.  Returns `out` ([B*Ho*Wo, ld_dst]).  Kernel selection (entry=None):
    1x1 -> segmif_linear_tc_fwd and 3x3 stride-1 'same' -> segmif_conv3x3_tc_fwd (tcgen05 + TMA) whenever the shape
    qualifies; strided / odd-width cases (patch_embed2-4, Attention.sr, linear_pred's 9 classes) ->
    segmif_conv_fwd (mma.sync implicit GEMM)."""
    st = _prep(src, weight, bias, prelu_alpha, residual, out, pre_add)
    if src.dtype != torch.bfloat16 or weight.dtype != torch.bfloat16:
        raise TypeError("segmif_b200.conv: src and weight must be bf16")
    Ho = (H + 2 * pad - dil * (KH - 1) - 1) // stride + 1
    Wo = (W + 2 * pad - dil * (KW - 1) - 1) // stride + 1
    M = B * Ho * Wo
    ld_src = Cin if ld_src is None else ld_src
    if out is None:
        ld_dst = Cout if ld_dst is None else ld_dst
        out = torch.empty((M, ld_dst), dtype=out_dtype, device=src.device)
    elif ld_dst is None:
        ld_dst = out.shape[-1]
    if residual is not None and ld_res is None:
        ld_res = residual.shape[-1]
    p = _lib.ConvParams()
    p.src, p.weight, p.bias = src.data_ptr(), weight.data_ptr(), (bias.data_ptr() if bias is not None else None)
    p.prelu_alpha = prelu_alpha.data_ptr() if prelu_alpha is not None else None
    p.residual = residual.data_ptr() if residual is not None else None
    p.dst = out.data_ptr()
    p.B, p.H, p.W, p.Cin, p.ld_src, p.src_coff = B, H, W, Cin, ld_src, src_coff
    p.KH, p.KW, p.stride, p.pad, p.dil, p.Ho, p.Wo, p.Cout = KH, KW, stride, pad, dil, Ho, Wo, Cout
    p.act = act
    p.res_dtype = _dt(residual) if residual is not None else F32
    p.ld_res, p.res_coff = (ld_res or 0), res_coff
    p.dst_dtype, p.ld_dst, p.dst_coff = _dt(out), ld_dst, dst_coff
    if pre_add is not None:
        p.pre_add, p.ld_pre, p.pre_coff = pre_add.data_ptr(), pre_add.shape[-1], pre_coff
    if entry is None:
        entry = "segmif_conv_fwd"
        if USE_TCGEN05:
            dal = 4 if out.dtype == torch.float32 else 8
            ral = 4 if (residual is not None and residual.dtype == torch.float32) else 8
            aligned = (ld_src % 8 == 0 and src_coff % 8 == 0 and ld_dst % dal == 0 and dst_coff % dal == 0
                       and (residual is None or ((ld_res or 0) % ral == 0 and res_coff % ral == 0)))
            if KH == 1 and KW == 1 and stride == 1 and pad == 0 and Cout % 32 == 0 and Cin % 8 == 0 and aligned and pre_add is None:
                lp = _linear_params(src, weight, bias, M, Cout, Cin, ld_src, src_coff, act, prelu_alpha, residual, ld_res,
                                    res_coff, out, ld_dst, dst_coff)
                _lib.call("segmif_linear_tc_fwd", ctypes.byref(lp), st)
                return out
            if (KH == 3 and KW == 3 and stride == 1 and pad == dil and dil in (1, 2) and Cout in (32, 64)
                    and out.dtype == torch.bfloat16 and residual is None and bias is not None and act != ACT_GELU
                    and aligned and _conv3x3_tc_fits(Cin, Cout, dil, pre_add is not None)):
                entry = "segmif_conv3x3_tc_fwd"
    _lib.call(entry, ctypes.byref(p), st)
    return out


USE_TCGEN05 = os.environ.get("SEGMIF_TCGEN05", "1") != "0"


def _conv3x3_tc_fits(Cin, Cout, dil, has_pre=False):
    """Resident weights + output / partial staging tiles + two halo-tile slots must fit the 227 KB of shared memory
    (mirrors conv_tc.cu for the smallest tile, NSUB = 1)."""
    wbytes = ((Cin + 63) // 64) * 9 * Cout * 128
    out_bytes = 128 * Cout * 2
    a_stride = ((16 + 2 * dil) * (8 + 2 * dil) * 128 + 1023) // 1024 * 1024
    return wbytes + out_bytes * (3 if has_pre else 1) + 17 * 8 + 16 + 2 * a_stride <= 227 * 1024 - 1024


def conv3x3_tc(src, weight, bias, **kw):
    """3x3 stride-1 'same' conv (dil 1|2, Cout 32|64, bf16 out) on tcgen05 tensor cores with TMA halo tiles."""
    return conv(src, weight, bias, KH=3, KW=3, entry="segmif_conv3x3_tc_fwd", **kw)


def drdb_push(buf, weight, B, H, W, slab_offset, slab_width, groups):
    """One DRDB growth step in push form (segmif_drdb_push_tc_fwd).  `groups`: list of dicts
    (bias, partial_in, coff_partial_in, dst, coff_dst, relu), one per 32 output channels."""
    tensors = [buf, weight] + [g.get("bias") for g in groups] + [g.get("partial_in") for g in groups] + [g["dst"] for g in groups]
    st = _prep(*tensors)
    p = _lib.DrdbPushParams()
    p.src, p.weight = buf.data_ptr(), weight.data_ptr()
    p.B, p.H, p.W, p.ld_src = B, H, W, buf.shape[-1]
    p.slab_offset, p.slab_width, p.n_out = slab_offset, slab_width, 32 * len(groups)
    for i, g in enumerate(groups):
        q = p.groups[i]
        q.bias = g["bias"].data_ptr() if g.get("bias") is not None else None
        pin = g.get("partial_in")
        q.partial_in = pin.data_ptr() if pin is not None else None
        q.ld_partial_in = pin.shape[-1] if pin is not None else 0
        q.coff_partial_in = g.get("coff_partial_in", 0)
        q.dst, q.ld_dst, q.coff_dst = g["dst"].data_ptr(), g["dst"].shape[-1], g["coff_dst"]
        q.relu = 1 if g.get("relu") else 0
    _lib.call("segmif_drdb_push_tc_fwd", ctypes.byref(p), st)


_DF_CTAS = [int(v) for v in os.environ.get("SEGMIF_DRDB_CTAS", "").split(",") if v.strip()]     # tuning: 7 SM counts


def drdb_dataflow(buf, part, w_push_a, w_push_b, w_pull, biases, w_1x1, bias_1x1, out, ld_out, out_coff, B, H, W, flags=None,
                  ctas=None):
    """One whole DRDB as seven concurrent kernels chained through L2 (segmif_drdb_dataflow_fwd).  Returns the flags
    workspace (its last word is non-zero if a dependency wait timed out)."""
    st = _prep(buf, part, w_push_a, w_push_b, *w_pull, *biases, w_1x1, bias_1x1, out, flags)
    dev = buf.device
    nwords = _lib.load().segmif_drdb_dataflow_workspace_bytes(B, H) // 4
    if flags is None:
        flags = torch.empty((nwords,), dtype=torch.int32, device=dev)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    if idx not in _DF_READY:
        _lib.check(_lib.load().segmif_drdb_dataflow_prepare(idx), "segmif_drdb_dataflow_prepare")
        _DF_READY.add(idx)
    p = _lib.DrdbDataflowParams()
    p.growth, p.ld, p.partial, p.ld_partial = buf.data_ptr(), buf.shape[-1], part.data_ptr(), part.shape[-1]
    p.w_push_a, p.w_push_b = w_push_a.data_ptr(), w_push_b.data_ptr()
    for i in range(4):
        p.w_pull[i] = w_pull[i].data_ptr()
    for i in range(5):
        p.bias[i] = biases[i].data_ptr()
    p.w_1x1, p.bias_1x1 = w_1x1.data_ptr(), bias_1x1.data_ptr()
    p.out, p.ld_out, p.out_coff = out.data_ptr(), ld_out, out_coff
    p.B, p.H, p.W = B, H, W
    p.flags = flags.data_ptr()
    ctas = ctas if ctas is not None else (_DF_CTAS if len(_DF_CTAS) == 7 else None)
    for i in range(7):
        p.ctas[i] = int(ctas[i]) if ctas else 0
    _lib.call("segmif_drdb_dataflow_fwd", ctypes.byref(p), st)
    return flags


_DF_READY = set()


def conv_mma(src, weight, bias, **kw):
    """Forces the mma.sync implicit-GEMM kernel (segmif_conv_fwd)."""
    return conv(src, weight, bias, entry="segmif_conv_fwd", **kw)


def linear(x, weight, bias, *, act=ACT_NONE, residual=None, out_dtype=torch.bfloat16, out=None, ld_dst=None,
           dst_coff=0):
    """y[rows, N] = act(x[rows, K] @ W^T + b) (+ residual); x bf16 [..., K], W packed bf16 [N, 1, K]."""
    K = x.shape[-1]
    rows = x.numel() // K
    N = weight.shape[0]
    return conv(x, weight, bias, B=1, H=1, W=rows, Cin=K, Cout=N, act=act, residual=residual, out=out,
                out_dtype=out_dtype, ld_dst=ld_dst, dst_coff=dst_coff)


def _linear_params(x, weight, bias, M, N, K, ld_src, src_coff, act, prelu_alpha, residual, ld_res, res_coff, out, ld_dst,
                   dst_coff, weight_kn=False, row_scale=None, rows_per_scale=0):
    p = _lib.LinearParams()
    p.weight_kn = 1 if weight_kn else 0
    p.row_scale = row_scale.data_ptr() if row_scale is not None else None
    p.rows_per_scale = int(rows_per_scale)
    p.src, p.weight = x.data_ptr(), weight.data_ptr()
    p.bias = bias.data_ptr() if bias is not None else None
    p.prelu_alpha = prelu_alpha.data_ptr() if prelu_alpha is not None else None
    p.residual = residual.data_ptr() if residual is not None else None
    p.dst = out.data_ptr()
    p.M, p.N, p.K, p.ld_src, p.src_coff = M, N, K, ld_src, src_coff
    p.act = act
    p.res_dtype = _dt(residual) if residual is not None else F32
    p.ld_res, p.res_coff = (ld_res or 0), res_coff
    p.dst_dtype, p.ld_dst, p.dst_coff = _dt(out), ld_dst, dst_coff
    return p


def linear_tc(x, weight, bias, *, M=None, K=None, ld_src=None, src_coff=0, act=ACT_NONE, prelu_alpha=None, residual=None,
              ld_res=None, res_coff=0, out=None, out_dtype=torch.bfloat16, ld_dst=None, dst_coff=0, weight_kn=False,
              row_scale=None, rows_per_scale=0):
    """tcgen05 + TMA linear layer (segmif_linear_tc_fwd).  x bf16 [..., ld_src]; weight packed bf16 [N, 1, K], or with
    weight_kn the [K, 1, N] buffer -- a forward pack reused as the operand of the data gradient dX = dY W."""
    st = _prep(x, weight, bias, prelu_alpha, residual, out, row_scale)
    if weight_kn:
        N = weight.shape[-1]
        K = weight.shape[0] if K is None else K
    else:
        N = weight.shape[0]
        K = weight.shape[-1] if K is None else K
    ld_src = x.shape[-1] if ld_src is None else ld_src
    M = x.numel() // ld_src if M is None else M
    if out is None:
        ld_dst = N if ld_dst is None else ld_dst
        out = torch.empty((M, ld_dst), dtype=out_dtype, device=x.device)
    elif ld_dst is None:
        ld_dst = out.shape[-1]
    if residual is not None and ld_res is None:
        ld_res = residual.shape[-1]
    p = _linear_params(x, weight, bias, M, N, K, ld_src, src_coff, act, prelu_alpha, residual, ld_res, res_coff, out,
                       ld_dst, dst_coff, weight_kn, row_scale, rows_per_scale)
    _lib.call("segmif_linear_tc_fwd", ctypes.byref(p), st)
    return out


def patch_embed7_ln(img, w147, bias, gamma, beta, eps, in_scale=None, in_shift=None):
    st = _prep(img, w147, bias, gamma, beta, in_scale, in_shift)
    B, C, H, W = img.shape
    assert C == 3 and img.dtype == torch.float32
    C0 = w147.shape[1]
    Ho, Wo = (H + 6 - 7) // 4 + 1, (W + 6 - 7) // 4 + 1
    tokens = torch.empty((B, Ho * Wo, C0), dtype=torch.float32, device=img.device)
    _lib.call("segmif_patch_embed7_ln_fwd", _ptr(img), _ptr(w147), _ptr(bias), _ptr(gamma), _ptr(beta), float(eps),
              _ptr(in_scale), _ptr(in_shift), _ptr(tokens), B, H, W, C0, st)
    return tokens, Ho, Wo


def sr_attention(q, kv, B, heads, N, Nk, D, scale):
    """q bf16 [B*N, C]; kv bf16 [B*Nk, 2C] (k = first C columns, v = last C); returns bf16 [B*N, C]."""
    st = _prep(q, kv)
    C = heads * D
    out = torch.empty((B * N, C), dtype=torch.bfloat16, device=q.device)
    kptr = kv.data_ptr()
    _lib.call("segmif_sr_attention_fwd", _ptr(q), C, ctypes.c_void_p(kptr), ctypes.c_void_p(kptr + 2 * C), 2 * C,
              _ptr(out), C, B, heads, N, Nk, D, float(scale), st)
    return out


def dwconv3x3_gelu(x, w9c, bias, B, H, W):
    st = _prep(x, w9c, bias)
    C = x.shape[-1]
    y = torch.empty_like(x)
    _lib.call("segmif_dwconv3x3_gelu_fwd", _ptr(x), _ptr(w9c), _ptr(bias), _ptr(y), B, H, W, C, st)
    return y


def bilinear_nhwc(src, B, h, w, C, H, W, *, ld_src=None, out=None, out_dtype=torch.bfloat16, ld_dst=None, dst_coff=0):
    st = _prep(src, out)
    ld_src = C if ld_src is None else ld_src
    if out is None:
        ld_dst = C if ld_dst is None else ld_dst
        out = torch.empty((B, H, W, ld_dst), dtype=out_dtype, device=src.device)
    elif ld_dst is None:
        ld_dst = out.shape[-1]
    _lib.call("segmif_bilinear_nhwc_fwd", _ptr(src), _dt(src), B, h, w, C, ld_src, _ptr(out), _dt(out), H, W, ld_dst,
              dst_coff, st)
    return out


def upsample_argmax(logits_nhwc, B, h, w, nc, H, W):
    st = _prep(logits_nhwc)
    assert logits_nhwc.dtype == torch.float32
    labels = torch.empty((B, H, W), dtype=torch.int64, device=logits_nhwc.device)
    _lib.call("segmif_upsample_argmax_fwd", _ptr(logits_nhwc), B, h, w, nc, _ptr(labels), H, W, st)
    return labels


def nhwc_to_nchw(src, B, HW, C, *, ld_src=None, src_coff=0, out=None):
    st = _prep(src, out)
    ld_src = C if ld_src is None else ld_src
    if out is None:
        out = torch.empty((B, C, HW), dtype=torch.float32, device=src.device)
    _lib.call("segmif_nhwc_to_nchw", _ptr(src), _dt(src), ld_src, src_coff, _ptr(out), B, HW, C, st)
    return out


def nchw_to_nhwc(src, *, out=None, out_dtype=torch.bfloat16, ld_dst=None, dst_coff=0):
    """src fp32 [B, C, ...spatial] -> pixel-major [B, HW, ld_dst]."""
    st = _prep(src, out)
    assert src.dtype == torch.float32
    B, C = src.shape[0], src.shape[1]
    HW = src.numel() // (B * C)
    if out is None:
        ld_dst = C if ld_dst is None else ld_dst
        out = torch.empty((B, HW, ld_dst), dtype=out_dtype, device=src.device)
    elif ld_dst is None:
        ld_dst = out.shape[-1]
    _lib.call("segmif_nchw_to_nhwc", _ptr(src), _ptr(out), _dt(out), ld_dst, dst_coff, B, HW, C, st)
    return out


def conv3x3_in1(img_nchw, w9c, bias, alpha, out, ld_dst, dst_coff, Cout):
    """Channel 0 of an fp32 NCHW tensor -> bf16 pixel-major `out` channels dst_coff..dst_coff+Cout."""
    st = _prep(w9c, bias, alpha, out)
    B, C, H, W = img_nchw.shape
    if not img_nchw.is_cuda or img_nchw.dtype != torch.float32 or img_nchw.stride(3) != 1 or img_nchw.stride(2) != W:
        raise RuntimeError("segmif_b200.conv3x3_in1: need a CUDA fp32 NCHW tensor with dense rows")
    _lib.call("segmif_conv3x3_in1_fwd", _ptr(img_nchw), img_nchw.stride(0), _ptr(w9c), _ptr(bias), _ptr(alpha), _ptr(out),
              ld_dst, dst_coff, B, H, W, Cout, st)
    return out


def conv3x3_out1(src, w9c, bias, alpha, B, H, W, Cin, ld_src=None):
    st = _prep(src, w9c, bias, alpha)
    out = torch.empty((B, 1, H, W), dtype=torch.float32, device=src.device)
    _lib.call("segmif_conv3x3_out1_fwd", _ptr(src), Cin if ld_src is None else ld_src, _ptr(w9c), _ptr(bias),
              _ptr(alpha), _ptr(out), B, H, W, Cin, st)
    return out


def ffm(x1, ld1, coff1, x2, ld2, coff2, x3, ld3, C3, packs, out1, ldo1, coffo1, out2, ldo2, coffo2, B, HW,
        want_ctx=False):
    """Three-pass hierarchical interactive attention; `packs` is the dict built by core.model_fusion._pack_ffm."""
    st = _prep(x1, x2, x3, out1, out2)
    dev = x1.device
    nchunk = max(1, min((148 * 4) // max(B, 1), (HW + 63) // 64))
    partials = torch.empty((B, nchunk, 3, 64, 64), dtype=torch.float32, device=dev)
    folded = torch.empty((B, 4, 64, 64), dtype=torch.bfloat16, device=dev)
    ctx = torch.empty((B, 3, 8, 8, 8), dtype=torch.float32, device=dev)
    _lib.call("segmif_ffm_gram_fwd", _ptr(x1), ld1, coff1, _ptr(x2), ld2, coff2, _ptr(x3), ld3, C3,
              _ptr(packs["w_gram"]), _ptr(packs["b_gram"]), _ptr(partials), nchunk, B, HW, st)
    _lib.call("segmif_ffm_ctx_fwd", _ptr(partials), nchunk, _ptr(packs["wkv"]), _ptr(packs["wend"]), _ptr(folded),
              _ptr(ctx), B, st)
    _lib.call("segmif_ffm_apply_fwd", _ptr(x1), ld1, coff1, _ptr(x2), ld2, coff2, _ptr(x3), ld3, C3,
              _ptr(packs["w_apply"]), _ptr(packs["b_apply"]), _ptr(folded), _ptr(packs["bend"]), _ptr(packs["ln_g"]),
              _ptr(packs["ln_b"]), 1e-5, _ptr(out1), ldo1, coffo1, _ptr(out2), ldo2, coffo2, B, HW, st)
    return ctx


_SM_COUNT = {}


def _sm_count(dev):
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    if idx not in _SM_COUNT:
        _SM_COUNT[idx] = torch.cuda.get_device_properties(idx).multi_processor_count
    return _SM_COUNT[idx]


def ffm_lr(x1, ld1, coff1, x2, ld2, coff2, q3, qh, qw, H, W, packs, out1, ldo1, coffo1, out2, ldo2, coffo2, B):
    """Hierarchical interactive attention with the seg stream given as the low-resolution pre-activation Q
    (bf16 [B, qh, qw, 128]); `packs` from CrossPath.packs_lr."""
    st = _prep(x1, x2, q3, out1, out2)
    dev = x1.device
    HW = H * W
    nchunk = max(1, min(_sm_count(dev) // max(B, 1), (HW + 127) // 128))     # one persistent CTA per SM, 128-pixel tiles
    partials = torch.empty((B, nchunk, 3, 64, 64), dtype=torch.float32, device=dev)
    folded = torch.empty((B, 4, 64, 64), dtype=torch.bfloat16, device=dev)
    ctx = torch.empty((B, 3, 8, 8, 8), dtype=torch.float32, device=dev)
    _lib.call("segmif_ffm_gram_lr_fwd", _ptr(x1), ld1, coff1, _ptr(x2), ld2, coff2, _ptr(q3), qh, qw, H, W,
              _ptr(packs["w_gram"]), _ptr(packs["b_gram"]), _ptr(partials), nchunk, B, st)
    _lib.call("segmif_ffm_ctx_fwd", _ptr(partials), nchunk, _ptr(packs["wkv"]), _ptr(packs["wend"]), _ptr(folded),
              _ptr(ctx), B, st)
    _lib.call("segmif_ffm_apply_lr_fwd", _ptr(x1), ld1, coff1, _ptr(x2), ld2, coff2, _ptr(q3), qh, qw, H, W,
              _ptr(packs["w_apply"]), _ptr(packs["b_apply"]), _ptr(folded), _ptr(packs["bend"]), _ptr(packs["ln_g"]),
              _ptr(packs["ln_b"]), 1e-5, _ptr(out1), ldo1, coffo1, _ptr(out2), ldo2, coffo2, B, st)
    return ctx


def rgb2ycrcb(x):
    st = _prep(x)
    assert x.dtype == torch.float32 and x.shape[1] == 3
    out = torch.empty_like(x)
    _lib.call("segmif_rgb2ycrcb", _ptr(x), _ptr(out), x.shape[0], x.shape[2] * x.shape[3], st)
    return out


def ycrcb2rgb(x):
    st = _prep(x)
    assert x.dtype == torch.float32 and x.shape[1] == 3
    out = torch.empty_like(x)
    _lib.call("segmif_ycrcb2rgb", _ptr(x), _ptr(out), x.shape[0], x.shape[2] * x.shape[3], st)
    return out


def recompose_rgb(fused_y, vis_rgb, clamp=True):
    st = _prep(fused_y, vis_rgb)
    out = torch.empty_like(vis_rgb)
    _lib.call("segmif_recompose_rgb", _ptr(fused_y), _ptr(vis_rgb), _ptr(out), 1 if clamp else 0, vis_rgb.shape[0],
              vis_rgb.shape[2] * vis_rgb.shape[3], st)
    return out


# ---------------------------------------------------------------------------------------------- losses
def _loss_ws(ref, B, H, W):
    nbytes = _lib.load().segmif_loss_workspace_bytes(B, H, W)
    return torch.empty(((nbytes + 3) // 4,), dtype=torch.float32, device=ref.device)


def _plane(t):
    if t.dim() != 4 or t.shape[1] != 1 or t.dtype != torch.float32:
        raise ValueError("segmif_b200 losses take fp32 [B,1,H,W] planes")
    return t if t.is_contiguous() else t.contiguous()


def ssim(a, b, size_average=True):
    a, b = _plane(a), _plane(b)
    st = _prep(a, b)
    B, _, H, W = a.shape
    ws = _loss_ws(a, B, H, W)
    out = torch.empty((1 if size_average else B,), dtype=torch.float32, device=a.device)
    _lib.call("segmif_ssim_fwd", _ptr(a), _ptr(b), B, H, W, 0 if size_average else 1, _ptr(ws), _ptr(out), st)
    return out[0] if size_average else out


def laploss2(inp, ir, vis):
    inp, ir, vis = _plane(inp), _plane(ir), _plane(vis)
    st = _prep(inp, ir, vis)
    B, _, H, W = inp.shape
    ws = _loss_ws(inp, B, H, W)
    out = torch.empty((1,), dtype=torch.float32, device=inp.device)
    _lib.call("segmif_laploss2_fwd", _ptr(inp), _ptr(ir), _ptr(vis), B, H, W, _ptr(ws), _ptr(out), st)
    return out[0]


def laploss(inp, target):
    inp, target = _plane(inp), _plane(target)
    st = _prep(inp, target)
    B, _, H, W = inp.shape
    ws = _loss_ws(inp, B, H, W)
    out = torch.empty((1,), dtype=torch.float32, device=inp.device)
    _lib.call("segmif_laploss_fwd", _ptr(inp), _ptr(target), B, H, W, _ptr(ws), _ptr(out), st)
    return out[0]


def entropy(img, patch):
    img = _plane(img)
    st = _prep(img)
    B, _, H, W = img.shape
    ws = _loss_ws(img, B, H, W)
    out = torch.empty((1,), dtype=torch.float32, device=img.device)
    _lib.call("segmif_entropy_fwd", _ptr(img), B, H, W, int(patch), _ptr(ws), _ptr(out), st)
    return out[0]


def sobel_l1(x, y):
    """returns (mean |x-y|, mean |sobel(x)-sobel(y)|)"""
    x, y = _plane(x), _plane(y)
    st = _prep(x, y)
    B, _, H, W = x.shape
    ws = _loss_ws(x, B, H, W)
    out = torch.empty((2,), dtype=torch.float32, device=x.device)
    _lib.call("segmif_sobel_l1_fwd", _ptr(x), _ptr(y), B, H, W, _ptr(ws), _ptr(out), st)
    return out[0], out[1]


def mse_l1(x, y):
    """returns (mean (x-y)^2, mean |x-y|) over all elements"""
    x = x if x.is_contiguous() else x.contiguous()
    y = y if y.is_contiguous() else y.contiguous()
    st = _prep(x, y)
    ws = _loss_ws(x, 1, 32, 32)
    out = torch.empty((2,), dtype=torch.float32, device=x.device)
    _lib.call("segmif_mse_l1_fwd", _ptr(x), _ptr(y), x.numel(), _ptr(ws), _ptr(out), st)
    return out[0], out[1]


def sobel_map(x):
    """core/loss.py:647-650: |sobel_x(x)| + |sobel_y(x)| as a map."""
    x = _plane(x)
    st = _prep(x)
    B, _, H, W = x.shape
    out = torch.empty_like(x)
    _lib.call("segmif_sobel_map_fwd", _ptr(x), _ptr(out), B, H, W, st)
    return out


def sobel_map_bwd(x, dout):
    x, dout = _plane(x), _plane(dout)
    st = _prep(x, dout)
    B, _, H, W = x.shape
    dx = torch.empty_like(x)
    _lib.call("segmif_sobel_map_bwd", _ptr(x), _ptr(dout), _ptr(dx), B, H, W, 0, st)
    return dx


EW_LINCOMB, EW_MAX, EW_ABS_AFFINE, EW_MUL, EW_ABS_AFFINE_BWD = 0, 1, 2, 3, 4


def ew2(x, y, mode, a=0.0, b=0.0):
    x = x.float().contiguous()
    y = y.float().contiguous() if y is not None else None
    st = _prep(x, y)
    out = torch.empty_like(x)
    _lib.call("segmif_ew2", _ptr(x), _ptr(y), float(a), float(b), int(mode), _ptr(out), x.numel(), st)
    return out


def upsample_ce(logits_nhwc, B, h, w, nc, labels, ignore_index=255, return_count=False):
    st = _prep(logits_nhwc, labels)
    H, W = labels.shape[1], labels.shape[2]
    ws = _loss_ws(logits_nhwc, 1, 32, 32)
    out = torch.empty((2,), dtype=torch.float32, device=labels.device)       # {mean loss, number of valid labels}
    _lib.call("segmif_upsample_ce_fwd", _ptr(logits_nhwc), B, h, w, nc, _ptr(labels), H, W, int(ignore_index),
              _ptr(ws), _ptr(out), st)
    return (out[0], out[1:2]) if return_count else out[0]


# ---------------------------------------------------------------------------------------------- training side
def _gout(g, n, like):
    """Upstream gradient(s) as a contiguous fp32 device vector of n entries (None -> 0)."""
    if isinstance(g, (list, tuple)):
        parts = [torch.zeros((), dtype=torch.float32, device=like.device) if x is None else x.reshape(()).float() for x in g]
        g = torch.stack(parts)
    g = g.reshape(-1).float().contiguous()
    assert g.numel() == n, (g.shape, n)
    return g


def mse_l1_bwd(x, y, g_mse, g_l1, out=None):
    """d(g_mse * mse + g_l1 * l1)/dx for the two outputs of mse_l1; accumulates into `out` when given."""
    x = x if x.is_contiguous() else x.contiguous()
    y = y if y.is_contiguous() else y.contiguous()
    g = _gout([g_mse, g_l1], 2, x)
    st = _prep(x, y, g, out)
    acc = out is not None
    out = torch.empty_like(x) if out is None else out
    _lib.call("segmif_mse_l1_bwd", _ptr(x), _ptr(y), x.numel(), _ptr(g), _ptr(out), 1 if acc else 0, st)
    return out


def sobel_l1_bwd(x, y, g_l1, g_sobel, out=None):
    x, y = _plane(x), _plane(y)
    g = _gout([g_l1, g_sobel], 2, x)
    st = _prep(x, y, g, out)
    B, _, H, W = x.shape
    acc = out is not None
    out = torch.empty_like(x) if out is None else out
    _lib.call("segmif_sobel_l1_bwd", _ptr(x), _ptr(y), B, H, W, _ptr(g), _ptr(out), 1 if acc else 0, st)
    return out


def ssim_bwd(a, b, gout, size_average=True, out=None):
    a, b = _plane(a), _plane(b)
    B, _, H, W = a.shape
    g = _gout(gout, 1 if size_average else B, a)
    st = _prep(a, b, g, out)
    acc = out is not None
    out = torch.empty_like(a) if out is None else out
    _lib.call("segmif_ssim_bwd", _ptr(a), _ptr(b), B, H, W, 0 if size_average else 1, _ptr(g), _ptr(out), 1 if acc else 0, st)
    return out


def laploss2_bwd(inp, ir, vis, gout, out=None):
    inp, ir, vis = _plane(inp), _plane(ir), _plane(vis)
    g = _gout(gout, 1, inp)
    st = _prep(inp, ir, vis, g, out)
    B, _, H, W = inp.shape
    acc = out is not None
    out = torch.empty_like(inp) if out is None else out
    _lib.call("segmif_laploss2_bwd", _ptr(inp), _ptr(ir), _ptr(vis), B, H, W, _ptr(g), _ptr(out), 1 if acc else 0, st)
    return out


def laploss_bwd(inp, target, gout, out=None):
    inp, target = _plane(inp), _plane(target)
    g = _gout(gout, 1, inp)
    st = _prep(inp, target, g, out)
    B, _, H, W = inp.shape
    acc = out is not None
    out = torch.empty_like(inp) if out is None else out
    _lib.call("segmif_laploss_bwd", _ptr(inp), _ptr(target), B, H, W, _ptr(g), _ptr(out), 1 if acc else 0, st)
    return out


def entropy_bwd(img, patch, gout, out=None):
    img = _plane(img)
    g = _gout(gout, 1, img)
    st = _prep(img, g, out)
    B, _, H, W = img.shape
    acc = out is not None
    out = torch.empty_like(img) if out is None else out
    _lib.call("segmif_entropy_bwd", _ptr(img), B, H, W, int(patch), _ptr(g), _ptr(out), 1 if acc else 0, st)
    return out


def act_bwd(y, ldy, coffy, dy, lddy, coffdy, dz, lddz, coffdz, rows, C, act, alpha=None, dbias=None, dalpha=None):
    st = _prep(y, dy, dz, alpha, dbias, dalpha)
    _lib.call("segmif_act_bwd", _ptr(y), ldy, coffy, _ptr(dy), lddy, coffdy, _ptr(dz), lddz, coffdz, rows, C, act,
              _ptr(alpha), _ptr(dbias), _ptr(dalpha), st)
    return dz


def prelu_plane_bwd(out, dout, alpha, dz, lddz, coffdz, dbias=None, dalpha=None):
    st = _prep(out, dout, alpha, dz, dbias, dalpha)
    _lib.call("segmif_prelu_plane_bwd", _ptr(out), _ptr(dout), out.numel(), _ptr(alpha), _ptr(dz), lddz, coffdz,
              _ptr(dbias), _ptr(dalpha), st)
    return dz


def colsum(x, ld, coff, rows, C, out):
    st = _prep(x, out)
    _lib.call("segmif_colsum", _ptr(x), ld, coff, rows, C, _ptr(out), st)
    return out


def add_bf16(a, lda, coffa, b, ldb, coffb, out, ldo, coffo, rows, C):
    st = _prep(a, b, out)
    _lib.call("segmif_add_bf16", _ptr(a), lda, coffa, _ptr(b), ldb, coffb, _ptr(out), ldo, coffo, rows, C, st)
    return out


def layernorm_bwd(x, dy, lddy, coffdy, gamma, eps, dx, lddx, coffdx, rows, C, dgamma=None, dbeta=None, dxsum=None,
                  accumulate=False):
    st = _prep(x, dy, gamma, dx, dgamma, dbeta, dxsum)
    _lib.call("segmif_layernorm_bwd", _ptr(x), _dt(x), _ptr(dy), _dt(dy), lddy, coffdy, _ptr(gamma), float(eps), _ptr(dx),
              _dt(dx), lddx, coffdx, rows, C, _ptr(dgamma), _ptr(dbeta), _ptr(dxsum), 1 if accumulate else 0, st)
    return dx


def wgrad(dy, ldy, coffy, x, ldx, coffx, *, B, H, W, Cin, Cout, taps, dil, grad, s_co, s_tap, s_ci, co_take=None,
          ci_take=None, P=None):
    """grad[co*s_co + tap*s_tap + ci*s_ci] += sum_p dy[p][co] * x[p + tap][ci]  (fp32 `grad`, accumulated).
    taps = 1: linear layer over P rows (B, H, W ignored);  taps = 9: 3x3 'same' conv with dilation `dil`."""
    st = _prep(dy, x, grad)
    if taps == 1:
        P = B * H * W if P is None else P
        B, W = 1, 16
        H = (P + 15) // 16
    else:
        P = B * H * W
    nchunk = int(_lib.load().segmif_wgrad_chunks(B, H, W, P, Cin, Cout, taps, dil))
    ws = torch.empty((nchunk * Cout * taps * Cin,), dtype=torch.float32, device=dy.device)
    _lib.call("segmif_wgrad", _ptr(dy), ldy, coffy, _ptr(x), ldx, coffx, B, H, W, P, Cin, Cout, taps, dil, _ptr(ws), nchunk,
              _ptr(grad), s_co, s_tap, s_ci, Cout if co_take is None else co_take, Cin if ci_take is None else ci_take, st)
    return grad


def wgrad_lin(dy, ldy, coffy, x, ldx, coffx, *, P, Cin, Cout, grad, s_co, s_ci=1, dbias=None, co_take=None, ci_take=None):
    """Linear layer: grad[co*s_co + ci*s_ci] += sum_t dy[t][co] x[t][ci] and (dbias given) dbias[co] += sum_t dy[t][co] in
    one pass on tcgen05 (csrc/wgrad_lin_tc.cu)."""
    st = _prep(dy, x, grad) if dbias is None else _prep(dy, x, grad, dbias)
    nchunk = int(_lib.load().segmif_wgrad_lin_chunks(P, Cin, Cout))
    ws = torch.empty((nchunk * Cout * Cin,), dtype=torch.float32, device=dy.device)
    _lib.call("segmif_wgrad_lin", _ptr(dy), ldy, coffy, _ptr(x), ldx, coffx, P, Cin, Cout, _ptr(ws), nchunk, _ptr(grad), s_co, s_ci,
              Cout if co_take is None else co_take, Cin if ci_take is None else ci_take, _ptr(dbias) if dbias is not None else None, st)
    return grad


def ffm_train_fwd(x1, ld1, coff1, x2, ld2, coff2, x3, ld3, packs, out1, ldo1, coffo1, out2, ldo2, coffo2, B, HW):
    """Training forward of the HIA module (C3 = 64): returns the tensors the backward needs."""
    st = _prep(x1, x2, x3, out1, out2)
    dev = x1.device
    nchunk = max(1, min((_sm_count(dev) * 4) // max(B, 1), (HW + 63) // 64))
    partials = torch.empty((B, nchunk, 3, 64, 64), dtype=torch.float32, device=dev)
    folded = torch.empty((B, 4, 64, 64), dtype=torch.bfloat16, device=dev)
    ctx = torch.empty((B, 3, 8, 8, 8), dtype=torch.float32, device=dev)
    pre1 = torch.empty((B * HW, 64), dtype=torch.bfloat16, device=dev)
    pre2 = torch.empty_like(pre1)
    _lib.call("segmif_ffm_gram_fwd", _ptr(x1), ld1, coff1, _ptr(x2), ld2, coff2, _ptr(x3), ld3, 64,
              _ptr(packs["w_gram"]), _ptr(packs["b_gram"]), _ptr(partials), nchunk, B, HW, st)
    _lib.call("segmif_ffm_ctx_fwd", _ptr(partials), nchunk, _ptr(packs["wkv"]), _ptr(packs["wend"]), _ptr(folded),
              _ptr(ctx), B, st)
    _lib.call("segmif_ffm_apply_train_fwd", _ptr(x1), ld1, coff1, _ptr(x2), ld2, coff2, _ptr(x3), ld3, 64,
              _ptr(packs["w_apply"]), _ptr(packs["b_apply"]), _ptr(folded), _ptr(packs["bend"]), _ptr(packs["ln_g"]),
              _ptr(packs["ln_b"]), 1e-5, _ptr(out1), ldo1, coffo1, _ptr(out2), ldo2, coffo2, B, HW, _ptr(pre1), _ptr(pre2), st)
    return dict(partials=partials, nchunk=nchunk, folded=folded, ctx=ctx, pre1=pre1, pre2=pre2)


def ffm_bwd_gram(x1, ld1, coff1, x2, ld2, coff2, x3, ld3, coff3, dr1, dr2, wfull, bfull, B, HW):
    st = _prep(x1, x2, x3, dr1, dr2, wfull, bfull)
    nchunk = max(1, min((_sm_count(x1.device) * 2) // max(B, 1), (HW + 63) // 64))
    partials = torch.empty((B, nchunk, 4, 64, 64), dtype=torch.float32, device=x1.device)
    _lib.call("segmif_ffm_bwd_gram", _ptr(x1), ld1, coff1, _ptr(x2), ld2, coff2, _ptr(x3), ld3, coff3, _ptr(dr1), _ptr(dr2),
              _ptr(wfull), _ptr(bfull), _ptr(partials), nchunk, B, HW, st)
    return partials, nchunk


def ffm_bwd_ctx(rpart, nchunk_r, gpart, nchunk_g, ctx, wkv, wend, folded, dwkv, dwend, B):
    st = _prep(rpart, gpart, ctx, wkv, wend, folded, dwkv, dwend)
    mats = torch.empty((B, 7, 64, 64), dtype=torch.bfloat16, device=rpart.device)
    _lib.call("segmif_ffm_bwd_ctx", _ptr(rpart), nchunk_r, _ptr(gpart), nchunk_g, _ptr(ctx), _ptr(wkv), _ptr(wend),
              _ptr(folded), _ptr(mats), _ptr(dwkv), _ptr(dwend), B, st)
    return mats


def ffm_bwd_apply(x1, ld1, coff1, x2, ld2, coff2, x3, ld3, coff3, dr1, dr2, wfull, bfull, mats, B, HW):
    st = _prep(x1, x2, x3, dr1, dr2, wfull, bfull, mats)
    dP = [torch.empty((B * HW, 128), dtype=torch.bfloat16, device=x1.device) for _ in range(3)]
    _lib.call("segmif_ffm_bwd_apply", _ptr(x1), ld1, coff1, _ptr(x2), ld2, coff2, _ptr(x3), ld3, coff3, _ptr(dr1), _ptr(dr2),
              _ptr(wfull), _ptr(bfull), _ptr(mats), _ptr(dP[0]), _ptr(dP[1]), _ptr(dP[2]), B, HW, st)
    return dP


def adamw_step(param, grad, exp_avg, exp_avg_sq, *, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0):
    st = _prep(param, grad, exp_avg, exp_avg_sq)
    _lib.call("segmif_adamw_step", _ptr(param), _ptr(grad), _ptr(exp_avg), _ptr(exp_avg_sq), param.numel(), float(lr),
              float(beta1), float(beta2), float(eps), float(weight_decay), int(step), float(grad_scale), st)


# ---------------------------------------------------------------------------------------------- segmentation-net training
def sr_attention_train(q, kv, B, heads, N, Nk, D, scale):
    """Forward as sr_attention, plus the per-row log-sum-exp (exp2 domain) the backward recomputes P from."""
    st = _prep(q, kv)
    C = heads * D
    out = torch.empty((B * N, C), dtype=torch.bfloat16, device=q.device)
    lse = torch.empty((B * heads, N), dtype=torch.float32, device=q.device)
    kptr = kv.data_ptr()
    _lib.call("segmif_sr_attention_train_fwd", _ptr(q), C, ctypes.c_void_p(kptr), ctypes.c_void_p(kptr + 2 * C), 2 * C,
              _ptr(out), C, B, heads, N, Nk, D, float(scale), _ptr(lse), st)
    return out, lse


def sr_attention_tc(q, kv, B, heads, N, Nk, D, scale, want_lse=False):
    """The tcgen05 attention kernel explicitly (head dim 64, Nk <= 320); same arguments as sr_attention."""
    st = _prep(q, kv)
    C = heads * D
    out = torch.empty((B * N, C), dtype=torch.bfloat16, device=q.device)
    lse = torch.empty((B * heads, N), dtype=torch.float32, device=q.device) if want_lse else None
    kptr = kv.data_ptr()
    _lib.call("segmif_sr_attention_tc_fwd", _ptr(q), C, ctypes.c_void_p(kptr), ctypes.c_void_p(kptr + 2 * C), 2 * C,
              _ptr(out), C, B, heads, N, Nk, D, float(scale), _ptr(lse), st)
    return (out, lse) if want_lse else out


def sr_attention_fa(q, kv, B, heads, N, Nk, D, scale, want_lse=False):
    """The flash-style tcgen05 attention kernel explicitly (head dim 64, any Nk); same arguments as sr_attention."""
    st = _prep(q, kv)
    C = heads * D
    out = torch.empty((B * N, C), dtype=torch.bfloat16, device=q.device)
    lse = torch.empty((B * heads, N), dtype=torch.float32, device=q.device) if want_lse else None
    kptr = kv.data_ptr()
    _lib.call("segmif_sr_attention_fa_fwd", _ptr(q), C, ctypes.c_void_p(kptr), ctypes.c_void_p(kptr + 2 * C), 2 * C,
              _ptr(out), C, B, heads, N, Nk, D, float(scale), _ptr(lse), st)
    return (out, lse) if want_lse else out


def sr_attention_bwd(q, kv, out, dout, lse, B, heads, N, Nk, D, scale):
    """Returns (dq bf16 [B*N, C], dkv fp32 [B*Nk, 2C])."""
    st = _prep(q, kv, out, dout, lse)
    C = heads * D
    dq = torch.empty((B * N, C), dtype=torch.bfloat16, device=q.device)
    dkv = torch.zeros((B * Nk, 2 * C), dtype=torch.float32, device=q.device)
    kptr = kv.data_ptr()
    _lib.call("segmif_sr_attention_bwd", _ptr(q), C, ctypes.c_void_p(kptr), ctypes.c_void_p(kptr + 2 * C), 2 * C, _ptr(out),
              _ptr(dout), C, _ptr(lse), _ptr(dq), C, _ptr(dkv), 2 * C, C, B, heads, N, Nk, D, float(scale), st)
    return dq, dkv


def upsample_ce_bwd(logits_nhwc, B, h, w, nc, labels, gout, count, ignore_index=255):
    g = _gout(gout, 1, logits_nhwc)
    st = _prep(logits_nhwc, labels, g, count)
    H, W = labels.shape[1], labels.shape[2]
    dl = torch.empty_like(logits_nhwc)
    _lib.call("segmif_upsample_ce_bwd", _ptr(logits_nhwc), B, h, w, nc, _ptr(labels), H, W, int(ignore_index), _ptr(g),
              _ptr(count), _ptr(dl), st)
    return dl


def bilinear_nhwc_bwd(ddst, ld_dst, dst_coff, B, H, W, h, w, C):
    """Adjoint of bilinear_nhwc: ddst bf16 [B, H, W, ld_dst] slice -> dsrc bf16 [B, h, w, C]."""
    st = _prep(ddst)
    dsrc = torch.empty((B, h, w, C), dtype=torch.bfloat16, device=ddst.device)
    _lib.call("segmif_bilinear_nhwc_bwd", _ptr(ddst), ld_dst, dst_coff, H, W, _ptr(dsrc), B, h, w, C, st)
    return dsrc


def bn_train_fwd(z, gamma, beta, eps, momentum, running_mean, running_var):
    """Train-mode BatchNorm + ReLU over rows of z bf16 [rows, C]; returns (y bf16, stats fp32 [2, C] = mean, rstd)."""
    st = _prep(z, gamma, beta, running_mean, running_var)
    rows, C = z.shape
    ws = torch.zeros((2 * C,), dtype=torch.float64, device=z.device)
    stats = torch.empty((2, C), dtype=torch.float32, device=z.device)
    y = torch.empty_like(z)
    _lib.call("segmif_bn_train_fwd", _ptr(z), rows, C, _ptr(gamma), _ptr(beta), float(eps), float(momentum),
              _ptr(running_mean), _ptr(running_var), _ptr(ws), _ptr(stats), _ptr(y), st)
    for buf in (running_mean, running_var):        # the kernel wrote through raw pointers: bump torch's version counters so
        if buf is not None:                        # version-stamped caches (SegFormerHead._fuse_pack) see the update
            torch.autograd.graph.increment_version(buf)
    return y, stats


def bn_eval_fwd(z, gamma, beta, eps, running_mean, running_var):
    """Eval-mode BatchNorm + ReLU that keeps what its backward needs; returns (y bf16, stats = running mean, rstd)."""
    st = _prep(z, gamma, beta, running_mean, running_var)
    rows, C = z.shape
    stats = torch.empty((2, C), dtype=torch.float32, device=z.device)
    y = torch.empty_like(z)
    _lib.call("segmif_bn_eval_fwd", _ptr(z), rows, C, _ptr(gamma), _ptr(beta), float(eps), _ptr(running_mean),
              _ptr(running_var), _ptr(stats), _ptr(y), st)
    return y, stats


def bn_train_bwd(z, y, dy, stats, gamma, dgamma, dbeta, eval_mode=False):
    st = _prep(z, y, dy, stats, gamma, dgamma, dbeta)
    rows, C = z.shape
    ws = torch.zeros((2 * C,), dtype=torch.float64, device=z.device)
    dz = torch.empty_like(z)
    _lib.call("segmif_bn_eval_bwd" if eval_mode else "segmif_bn_train_bwd", _ptr(z), _ptr(y), _ptr(dy), _ptr(stats), _ptr(gamma), rows, C, _ptr(ws), _ptr(dz),
              _ptr(dgamma), _ptr(dbeta), st)
    return dz


def channel_scale(x, scale, B, HW, C, out=None):
    st = _prep(x, scale, out)
    out = torch.empty_like(x) if out is None else out
    _lib.call("segmif_channel_scale", _ptr(x), _ptr(scale), _ptr(out), B, HW, C, st)
    return out


def dwconv3x3(x, w9c, bias, B, H, W, flip=False):
    st = _prep(x, w9c, bias)
    y = torch.empty_like(x)
    _lib.call("segmif_dwconv3x3", _ptr(x), _ptr(w9c), _ptr(bias), _ptr(y), B, H, W, x.shape[-1], 1 if flip else 0, st)
    return y


def dwconv3x3_gelu_bwd(x, w9c, bias, dy, B, H, W, dw9c, dbias):
    st = _prep(x, w9c, bias, dy, dw9c, dbias)
    dz = torch.empty_like(x)
    C = x.shape[-1]
    ws = torch.empty((int(_lib.load().segmif_dwconv3x3_gelu_bwd_workspace(B, H, W, C)),), dtype=torch.float32, device=x.device)
    _lib.call("segmif_dwconv3x3_gelu_bwd", _ptr(x), _ptr(w9c), _ptr(bias), _ptr(dy), _ptr(dz), B, H, W, C,
              _ptr(dw9c), _ptr(dbias), _ptr(ws), st)
    return dz


def col2im(dcol, B, H, W, C, k, stride, pad, out_dtype=torch.bfloat16):
    st = _prep(dcol)
    dx = torch.empty((B, H, W, C), dtype=out_dtype, device=dcol.device)
    _lib.call("segmif_col2im", _ptr(dcol), dcol.shape[-1], _ptr(dx), _dt(dx), B, H, W, C, k, stride, pad, st)
    return dx


def channel_affine_nchw(x, scale, shift):
    st = _prep(x, scale, shift)
    y = torch.empty_like(x)
    B, C = x.shape[0], x.shape[1]
    _lib.call("segmif_channel_affine_nchw", _ptr(x), _ptr(scale), _ptr(shift), _ptr(y), B, C, x.numel() // (B * C), st)
    return y


def recompose_rgb_bwd(rgb, drgb, clamp=True):
    st = _prep(rgb, drgb)
    B, _, H, W = rgb.shape
    d = torch.empty((B, 1, H, W), dtype=torch.float32, device=rgb.device)
    _lib.call("segmif_recompose_rgb_bwd", _ptr(rgb), _ptr(drgb), _ptr(d), 1 if clamp else 0, B, H * W, st)
    return d


def cast(x, dtype):
    if x.dtype == dtype:
        return x
    st = _prep(x)
    y = torch.empty(x.shape, dtype=dtype, device=x.device)
    _lib.call("segmif_cast", _ptr(x), _dt(x), _ptr(y), _dt(y), x.numel(), st)
    return y


def scale_cast_rows(x, scale, rows_per_sample):
    """bf16(scale[row // rows_per_sample] * x) for fp32 x [rows, C]."""
    st = _prep(x, scale)
    C = x.shape[-1]
    y = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    _lib.call("segmif_scale_cast_rows", _ptr(x), _ptr(scale), _ptr(y), x.numel() // C, rows_per_sample, C, st)
    return y


def scale_add_rows(x, y, scale, rows_per_sample):
    """x fp32 [rows, C] + scale[row // rows_per_sample] * y  (scale None -> 1)."""
    st = _prep(x, y, scale)
    C = x.shape[-1]
    out = torch.empty_like(x)
    _lib.call("segmif_scale_add_rows", _ptr(x), _ptr(y), _dt(y), _ptr(scale), _ptr(out), x.numel() // C, rows_per_sample, C, st)
    return out
