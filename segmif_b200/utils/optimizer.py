"""AdamW with linear warm-up then polynomial decay applied inside step() -- the interface of the reference's
utils/optimizer.py (PolyWarmupAdamW :3-33, PolyWarmupAdamW_seg :36-66).  Stock torch.optim.AdamW underneath;
a fused multi-tensor AdamW merged with the gradient allreduce is listed under SURVEY.md 8(f)."""
import torch


class _PolyWarmup(torch.optim.AdamW):
    def __init__(self, params, lr, weight_decay, betas, start_step, warmup_iter, max_iter, warmup_ratio, power):
        super().__init__(params, lr=lr, betas=betas, weight_decay=weight_decay, eps=1e-8)
        self.global_step = start_step
        self.warmup_iter = warmup_iter
        self.warmup_ratio = warmup_ratio
        self.max_iter = max_iter
        self.power = power
        self._base_lr = [g['lr'] for g in self.param_groups]

    def _lr_mult(self):
        if self.global_step < self.warmup_iter:
            return 1 - (1 - self.global_step / self.warmup_iter) * (1 - self.warmup_ratio)
        if self.global_step < self.max_iter:
            return (1 - self.global_step / self.max_iter) ** self.power
        return None

    def step(self, closure=None):
        mult = self._lr_mult()
        if mult is not None:
            for g, base in zip(self.param_groups, self._base_lr):
                g['lr'] = base * mult
        out = super().step(closure)
        self.global_step += 1
        return out


class PolyWarmupAdamW(_PolyWarmup):
    def __init__(self, params, lr, weight_decay, betas, warmup_iter=None, max_iter=None, warmup_ratio=None, power=None):
        super().__init__(params, lr, weight_decay, betas, 0, warmup_iter, max_iter, warmup_ratio, power)


class PolyWarmupAdamW_seg(_PolyWarmup):
    def __init__(self, params, lr, weight_decay, betas, iter_curr, warmup_iter=None, max_iter=None, warmup_ratio=None,
                 power=None):
        super().__init__(params, lr, weight_decay, betas, iter_curr, warmup_iter, max_iter, warmup_ratio, power)
