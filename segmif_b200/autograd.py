"""torch.autograd registration of the loss kernels: each Function's forward is the fused forward kernel and its
backward the hand-written backward kernel of csrc/losses_bwd.cu (gradient with respect to the FIRST image argument,
the fused image in every loss train.py builds; the other arguments are data).  Autograd is used for graph
bookkeeping only -- no torch op computes anything here."""
import torch

from . import ops


def _need(ctx, idx=0):
    return ctx.needs_input_grad[idx]


def _no_second(ctx, *idx):
    for i in idx:
        if ctx.needs_input_grad[i]:
            raise NotImplementedError("segmif_b200: loss gradients are implemented for the first image argument only "
                                      "(the fused image); detach the other arguments")


class MseL1Fn(torch.autograd.Function):
    """(mean (x-y)^2, mean |x-y|)"""

    @staticmethod
    def forward(ctx, x, y):
        x, y = x.float().contiguous(), y.float().contiguous()
        ctx.save_for_backward(x, y)
        return ops.mse_l1(x, y)

    @staticmethod
    def backward(ctx, g_mse, g_l1):
        _no_second(ctx, 1)
        x, y = ctx.saved_tensors
        return ops.mse_l1_bwd(x, y, g_mse, g_l1), None


class SobelL1Fn(torch.autograd.Function):
    """(mean |x-y|, mean |sobel(x)-sobel(y)|)"""

    @staticmethod
    def forward(ctx, x, y):
        x, y = x.float().contiguous(), y.float().contiguous()
        ctx.save_for_backward(x, y)
        return ops.sobel_l1(x, y)

    @staticmethod
    def backward(ctx, g_l1, g_sobel):
        _no_second(ctx, 1)
        x, y = ctx.saved_tensors
        return ops.sobel_l1_bwd(x, y, g_l1, g_sobel), None


class SsimFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, size_average):
        a, b = a.float().contiguous(), b.float().contiguous()
        ctx.save_for_backward(a, b)
        ctx.size_average = size_average
        return ops.ssim(a, b, size_average)

    @staticmethod
    def backward(ctx, g):
        _no_second(ctx, 1)
        a, b = ctx.saved_tensors
        return ops.ssim_bwd(a, b, g, ctx.size_average), None, None


class LapLoss2Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inp, ir, vis):
        inp, ir, vis = inp.float().contiguous(), ir.float().contiguous(), vis.float().contiguous()
        ctx.save_for_backward(inp, ir, vis)
        return ops.laploss2(inp, ir, vis)

    @staticmethod
    def backward(ctx, g):
        _no_second(ctx, 1, 2)
        inp, ir, vis = ctx.saved_tensors
        return ops.laploss2_bwd(inp, ir, vis, g), None, None


class LapLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inp, target):
        inp, target = inp.float().contiguous(), target.float().contiguous()
        ctx.save_for_backward(inp, target)
        return ops.laploss(inp, target)

    @staticmethod
    def backward(ctx, g):
        _no_second(ctx, 1)
        inp, target = ctx.saved_tensors
        return ops.laploss_bwd(inp, target, g), None


class EntropyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, patch):
        img = img.float().contiguous()
        ctx.save_for_backward(img)
        ctx.patch = patch
        return ops.entropy(img, patch)

    @staticmethod
    def backward(ctx, g):
        (img,) = ctx.saved_tensors
        return ops.entropy_bwd(img, ctx.patch, g), None


class RecomposeRgbFn(torch.autograd.Function):
    """train.py:363-366: the visible image's YCrCb with Y replaced by the fused plane, back to RGB (optionally clamped
    to [0, 1] as test_fusion.py:108-111 does).  Gradient w.r.t. the fused plane only; `vis_rgb` is data."""

    @staticmethod
    def forward(ctx, fused_y, vis_rgb, clamp):
        rgb = ops.recompose_rgb(fused_y.float().contiguous(), vis_rgb.float().contiguous(), clamp)
        ctx.save_for_backward(rgb)
        ctx.clamp = clamp
        return rgb

    @staticmethod
    def backward(ctx, drgb):
        _no_second(ctx, 1)
        (rgb,) = ctx.saved_tensors
        return ops.recompose_rgb_bwd(rgb, drgb.float().contiguous(), ctx.clamp), None, None


def recompose_rgb(fused_y, vis_rgb, clamp=False):
    return RecomposeRgbFn.apply(fused_y, vis_rgb, clamp)


def mse_l1(x, y):
    return MseL1Fn.apply(x, y)


def sobel_l1(x, y):
    return SobelL1Fn.apply(x, y)


class SobelMapFn(torch.autograd.Function):
    """core/loss.py:647-650 Sobelxy.forward as a map, with its backward (sign maps re-derived from x)."""

    @staticmethod
    def forward(ctx, x):
        x = x.float().contiguous()
        ctx.save_for_backward(x)
        return ops.sobel_map(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return ops.sobel_map_bwd(x, g.float().contiguous())


class MulDataFn(torch.autograd.Function):
    """x * data (gradient to x only): the mask products of new_loss_sobel (core/loss.py:394-397)."""

    @staticmethod
    def forward(ctx, x, data):
        data = data.float().contiguous()
        ctx.save_for_backward(data)
        return ops.ew2(x, data, ops.EW_MUL)

    @staticmethod
    def backward(ctx, g):
        _no_second(ctx, 1)
        (data,) = ctx.saved_tensors
        return ops.ew2(g, data, ops.EW_MUL), None


class AbsAffineFn(torch.autograd.Function):
    """|a + b*x| (torch.abs(1 - mask), core/loss.py:392,615)."""

    @staticmethod
    def forward(ctx, x, a, b):
        x = x.float().contiguous()
        ctx.save_for_backward(x)
        ctx.ab = (a, b)
        return ops.ew2(x, None, ops.EW_ABS_AFFINE, a, b)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        a, b = ctx.ab
        return ops.ew2(x, g, ops.EW_ABS_AFFINE_BWD, a, b), None, None


def sobel_map(x):
    return SobelMapFn.apply(x)


def mul_data(x, data):
    return MulDataFn.apply(x, data)


def abs_affine(x, a, b):
    return AbsAffineFn.apply(x, a, b)
