"""Fusion-loss composites with the reference's names and argument order (core/loss.py:423-650 of SegMiF),
as thin compositions of the fused loss kernels: every term is ONE kernel launch that reduces to a device
scalar (the reference chains five cuDNN convolutions and ~10 elementwise kernels per SSIM).  The detection
leftovers of the reference file (FCOS / focal / OHEM, core/loss.py:18-397) are dead code there and are not
reproduced.  Gradients flow to the fused image (`generate_img`) through the hand-written backward kernels registered in
segmif_b200/autograd.py; the other arguments are data."""
import torch
import torch.nn as nn

from .. import ops  # noqa: F401
from ..autograd import abs_affine, mse_l1, mul_data, sobel_l1, sobel_map
from ..lap_loss import LapLoss, LapLoss2
from ..pytorch_ssim import ssim
from .Entropy import Entropy
from .model_fusion import RGB2YCrCb


def _y(t):
    return t[:, :1].float().contiguous()


class Sobelxy(nn.Module):
    """core/loss.py:634-650: |Gx| + |Gy| with zero padding, as a map (one kernel; backward registered).  The live
    composites (Fusionloss3) use the fused Sobel+L1 reduction instead of materialising gradient maps."""

    def __init__(self):
        super().__init__()
        kernelx = torch.tensor([[-1., 0., 1.], [-2., 0., 2.], [-1., 0., 1.]]).view(1, 1, 3, 3)
        kernely = torch.tensor([[1., 2., 1.], [0., 0., 0.], [-1., -2., -1.]]).view(1, 1, 3, 3)
        # plain attributes, as in the reference (nn.Parameter(...).cuda() returns a non-registered tensor there: no state_dict keys)
        self.weightx, self.weighty = kernelx, kernely

    def forward(self, x):
        return sobel_map(x)


class Fusionloss3(nn.Module):
    """core/loss.py:459-476: L1(mask, fused) + L1(sobel(mask), sobel(fused)) -- one fused kernel."""

    def __init__(self):
        super().__init__()
        self.sobelconv = Sobelxy()

    def forward(self, image_ir, image_vis, generate_img, mask):
        l1, lgrad = sobel_l1(generate_img, _y(mask))
        return l1 + lgrad


class Fusionloss2(nn.Module):
    """core/loss.py:441-457: L1(mask, fused)."""

    def __init__(self):
        super().__init__()
        self.sobelconv = Sobelxy()

    def forward(self, image_ir, image_vis, generate_img, mask):
        return mse_l1(generate_img, _y(mask))[1]


class Fusionloss_grad(nn.Module):
    """core/loss.py:479-490: L1(mask, fused) + 0.8 * LapLoss2(fused, ir, vis_y)."""

    def __init__(self):
        super().__init__()
        self.lap = LapLoss2()

    def forward(self, image_ir, image_vis, generate_img, mask):
        g = generate_img
        return mse_l1(g, _y(mask))[1] + 0.8 * self.lap(g, _y(image_ir), _y(image_vis))


class Fusionloss_grad2(nn.Module):
    """core/loss.py:492-505: L1 + 0.1 * LapLoss2(fused, vis_y, ir) + 1.1 * (1 - ssim(fused, mask))."""

    def __init__(self):
        super().__init__()
        self.lap = LapLoss2()

    def forward(self, image_ir, image_vis, generate_img, mask):
        g, m = generate_img, _y(mask)
        return mse_l1(g, m)[1] + 0.1 * self.lap(g, _y(image_vis), _y(image_ir)) + 1.1 * (1 - ssim(g, m))


class Fusionloss_grad3(nn.Module):
    """core/loss.py:506-517: MSE(mask, fused) + 1.1 * (1 - ssim(fused, mask)) -- the live loss of train.py:363-367."""

    def __init__(self):
        super().__init__()
        self.lap = LapLoss2()

    def forward(self, image_ir, image_vis, generate_img, mask):
        g, m = generate_img, _y(mask)
        return mse_l1(g, m)[0] + 1.1 * (1 - ssim(g, m))


class Fusionloss(nn.Module):
    """core/loss.py:423-439: L1(max(vis_y, ir), fused) + 8 * L1(max(sobel(vis_y), sobel(ir)), sobel(fused))."""

    def __init__(self):
        super().__init__()
        self.sobelconv = Sobelxy()

    def forward(self, image_ir, image_vis, generate_img):
        y, ir = _y(image_vis), _y(image_ir)
        loss_in = mse_l1(generate_img, ops.ew2(y, ir, ops.EW_MAX))[1]
        joint = ops.ew2(ops.sobel_map(y), ops.sobel_map(ir), ops.EW_MAX)
        loss_grad = mse_l1(self.sobelconv(generate_img), joint)[1]
        return loss_in + 8 * loss_grad


class Fusionloss4(nn.Module):
    """core/loss.py:545-559: L1((vis_y + ir) / 2, fused) + 4 * L1(sobel((vis_y + ir) / 2), sobel(fused))."""

    def __init__(self):
        super().__init__()
        self.sobelconv = Sobelxy()

    def forward(self, image_ir, image_vis, generate_img, mask):
        syn = ops.ew2(_y(image_vis), _y(image_ir), ops.EW_LINCOMB, 0.5, 0.5)
        l1, lgrad = sobel_l1(generate_img, syn)
        return l1 + 4 * lgrad


class Fusionloss_add(nn.Module):
    """core/loss.py:561-577: 1.5 * L1(0.4 vis_y + 0.6 ir, fused) + 5 * L1(max(sobel(vis_y), sobel(ir)), sobel(fused))."""

    def __init__(self):
        super().__init__()
        self.sobelconv = Sobelxy()

    def forward(self, image_ir, image_vis, generate_img):
        y, ir = _y(image_vis), _y(image_ir)
        loss_in = mse_l1(generate_img, ops.ew2(y, ir, ops.EW_LINCOMB, 0.4, 0.6))[1]
        joint = ops.ew2(ops.sobel_map(y), ops.sobel_map(ir), ops.EW_MAX)
        loss_grad = mse_l1(self.sobelconv(generate_img), joint)[1]
        return loss_in * 1.5 + 5 * loss_grad


class new_loss_sobel(nn.Module):
    """core/loss.py:386-399.  The reference REBINDS `mask_ir` / `mask_vis` to the two scalar MSE terms before they
    multiply the Sobel maps (:394-397), so the gradient terms are a^2 * MSE(sobel(fused), sobel(ir)) and
    v^2 * MSE(sobel(fused), sobel(vis)) with a, v the (differentiable) intensity terms; reproduced as written."""

    def __init__(self):
        super().__init__()
        self.sobel = Sobelxy()

    def forward(self, ir, vis, mask_ir, fused_img):
        ir, vis, m = _y(ir), _y(vis), _y(mask_ir)
        mv = ops.ew2(m, None, ops.EW_ABS_AFFINE, 1.0, -1.0)                       # torch.abs(1 - mask_ir)
        a = mse_l1(mul_data(fused_img, m), ops.ew2(m, ir, ops.EW_MUL))[0]
        v = mse_l1(mul_data(fused_img, mv), ops.ew2(mv, vis, ops.EW_MUL))[0]
        sf = self.sobel(fused_img)
        a2 = a * a * mse_l1(sf, ops.sobel_map(ir))[0]                              # MSE(a * S(f), a * S(ir)), a scalar
        v2 = v * v * mse_l1(sf, ops.sobel_map(vis))[0]
        return (v + v2) * 1.0 + (a + a2) * 0.85


class Total_fusion_loss(nn.Module):
    """core/loss.py:578-588 (note the argument order: mask BEFORE generate_img)."""

    def __init__(self):
        super().__init__()
        self.nls = new_loss_sobel()
        self.fl = Fusionloss()

    def forward(self, image_ir, image_vis, mask, generate_img):
        return self.fl(image_ir, image_vis, generate_img) * 1.2 + self.nls(image_ir, image_vis, mask, generate_img) * 0.85


class Total_fusion_loss2(nn.Module):
    """core/loss.py:591-599."""

    def __init__(self):
        super().__init__()
        self.nls = new_loss_sobel()

    def forward(self, image_ir, image_vis, mask, generate_img):
        return self.nls(image_ir, image_vis, mask, generate_img)


class Total_fusion_loss3(nn.Module):
    """core/loss.py:600-608."""

    def __init__(self):
        super().__init__()
        self.fl = Fusionloss()

    def forward(self, image_ir, image_vis, mask, generate_img):
        return self.fl(image_ir, image_vis, generate_img) * 3


class IQALoss(nn.Module):
    """core/loss.py:605-633: 0.5 MSE(lr, mask) + 0.5 MSE(vis, |1 - mask|) + the same two terms on Sobel maps.
    The entropy / std softmax weights the reference computes (:616-626) never enter the result; the two entropies are
    still evaluated (the only caller of core/Entropy.py) and kept in `last_entropy` as device scalars -- without the
    reference's host round trip (`torch.tensor([e1, e2])`).  Gradients flow to `mask`."""

    def __init__(self):
        super().__init__()
        self.entropy_ = Entropy(4)
        self.sobel = Sobelxy()
        self.last_entropy = None

    def forward(self, lr, vis, mask):
        lr, vis = _y(lr), _y(vis)
        m = mask[:, 0:1].float().contiguous()
        inv = abs_affine(m, 1.0, -1.0)                                             # torch.abs(1 - mask)
        with torch.no_grad():
            self.last_entropy = (self.entropy_(m.detach()), self.entropy_(inv.detach()))
        mse_loss = 0.5 * mse_l1(m, lr)[0] + 0.5 * mse_l1(inv, vis)[0]
        grad_loss = 0.5 * mse_l1(self.sobel(m), ops.sobel_map(lr))[0] + 0.5 * mse_l1(self.sobel(inv), ops.sobel_map(vis))[0]
        return mse_loss + grad_loss
