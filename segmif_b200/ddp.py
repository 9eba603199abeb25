"""Data-parallel training plumbing for train.py's loops: one process per GPU, the image-pair batch
sharded over ranks, replicated weights, and exactly ONE gradient all-reduce per optimizer step over a flat fp32
buffer (NCCL over NVLink on the B200 box; gloo in the CPU tests), followed by the fused AdamW kernel that also applies
the 1/world_size scale.  The reference has no distributed code at all (train.py:119,271 are commented out), so this
is the additive wrapper its `--local_rank` flag was meant for.

Parameters that never receive a gradient on the live path (Fusion_Network3_ac.ffm2.*, WeTr.classifier) are kept OUT of
the flat buffer: torch.optim.AdamW skips parameters whose grad is None (no weight decay either), and a zero-filled
slot would silently decay them ("unused parameters under DDP")."""
import torch
import torch.distributed as dist

from . import packing


class FlatParams:
    """Re-homes the selected parameters of `module` as views of one flat fp32 buffer and gives them gradient views
    of a second flat buffer, so the all-reduce and the optimizer each touch a single contiguous tensor."""

    def __init__(self, module, used=lambda name: True):
        self.named = [(k, p) for k, p in module.named_parameters() if p.requires_grad and used(k)]
        self.skipped = [k for k, p in module.named_parameters() if p.requires_grad and not used(k)]
        if not self.named:
            raise ValueError("FlatParams: no parameters selected")
        dev = self.named[0][1].device
        pad4 = lambda n: (n + 3) & ~3              # every view starts 16-byte aligned (the kernels read biases / gains as float4)
        self.numel = sum(pad4(p.numel()) for _, p in self.named)
        self.param = torch.zeros((self.numel,), dtype=torch.float32, device=dev)
        self.grad = torch.zeros((self.numel,), dtype=torch.float32, device=dev)
        self.offsets = {}
        off = 0
        with torch.no_grad():
            for k, p in self.named:
                n = p.numel()
                self.param[off:off + n].copy_(p.detach().reshape(-1))
                p.data = self.param[off:off + n].view(p.shape)
                p.grad = self.grad[off:off + n].view(p.shape)
                self.offsets[k] = (off, n)
                off += pad4(n)

    def zero_grad(self):
        self.grad.zero_()
        for k, p in self.named:                       # autograd may have replaced a view; re-attach if so
            off, n = self.offsets[k]
            if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + off * 4:
                p.grad = self.grad[off:off + n].view(p.shape)

    def all_reduce(self, group=None):
        """The step's single collective: SUM over ranks (the mean's 1/world is applied by the optimizer kernel)."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=group)
            return dist.get_world_size(group)
        return 1


def poly_warmup_lr(base_lr, global_step, warmup_iter, max_iter, warmup_ratio, power, current):
    """utils/optimizer.py:16-27 of the reference: linear warm-up, then polynomial decay; unchanged past max_iter."""
    if global_step < warmup_iter:
        return base_lr * (1 - (1 - global_step / warmup_iter) * (1 - warmup_ratio))
    if global_step < max_iter:
        return base_lr * (1 - global_step / max_iter) ** power
    return current


class FusedPolyWarmupAdamW:
    """PolyWarmupAdamW (utils/optimizer.py:3-33) over a FlatParams buffer: one segmif_adamw_step launch per step."""

    def __init__(self, flat, lr, weight_decay, betas, warmup_iter, max_iter, warmup_ratio, power, eps=1e-8):
        self.flat = flat
        self.base_lr = self.lr = float(lr)
        self.weight_decay, self.betas, self.eps = float(weight_decay), tuple(betas), float(eps)
        self.warmup_iter, self.max_iter, self.warmup_ratio, self.power = warmup_iter, max_iter, warmup_ratio, power
        self.global_step = 0
        self.exp_avg = torch.zeros_like(flat.param)
        self.exp_avg_sq = torch.zeros_like(flat.param)

    def step(self, grad_scale=1.0):
        from . import ops
        self.lr = poly_warmup_lr(self.base_lr, self.global_step, self.warmup_iter, self.max_iter, self.warmup_ratio,
                                 self.power, self.lr)
        ops.adamw_step(self.flat.param, self.flat.grad, self.exp_avg, self.exp_avg_sq, lr=self.lr, beta1=self.betas[0],
                       beta2=self.betas[1], eps=self.eps, weight_decay=self.weight_decay, step=self.global_step + 1,
                       grad_scale=grad_scale)
        packing.invalidate_all()
        self.global_step += 1


class FusionTrainer:
    """One optimisation step of train_fusion (train.py:343-386) for the fusion network, data parallel:
        fused = model2(ir, vis_ycrcb, out0, out1);  loss = criterion(ir, vis, fused, mask);  loss.backward();
        all-reduce;  AdamW.
    The encoder features out0 / out1 come from the frozen segmentation network (train.py:358-359, no_grad)."""

    def __init__(self, fusion_net, criterion, lr=3e-4, weight_decay=0.01, betas=(0.9, 0.999), warmup_iter=3e-5,
                 max_iter=40000, warmup_ratio=1e-6, power=1.0, group=None):
        self.net, self.criterion, self.group = fusion_net, criterion, group
        self.flat = FlatParams(fusion_net, used=lambda k: not k.startswith("ffm2."))
        self.opt = FusedPolyWarmupAdamW(self.flat, lr, weight_decay, betas, warmup_iter, max_iter, warmup_ratio, power)

    def step(self, ir, vis_ycrcb, out0, out1, mask):
        self.flat.zero_grad()
        fused = self.net(ir, vis_ycrcb, out0, out1)
        loss = self.criterion(ir, vis_ycrcb, fused, mask)
        loss.backward()
        world = self.flat.all_reduce(self.group)
        self.opt.step(grad_scale=1.0 / world)
        return loss.detach(), fused.detach()
