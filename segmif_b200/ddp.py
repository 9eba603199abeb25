"""Data-parallel training plumbing for train.py's loops: one process per GPU, the image-pair batch
sharded over ranks, replicated weights, and exactly ONE gradient all-reduce per optimizer step over a flat fp32
buffer (NCCL over NVLink on the B200 box; gloo in the CPU tests), followed by the fused AdamW kernel that also applies
the 1/world_size scale.  The reference has no distributed code at all (train.py:119,271 are commented out), so this
is the additive wrapper its `--local_rank` flag was meant for.

Parameters that never receive a gradient on the live path (Fusion_Network3_ac.ffm2.*, WeTr.classifier) are kept OUT of
the flat buffer: torch.optim.AdamW skips parameters whose grad is None (no weight decay either), and a zero-filled
slot would silently decay them ("unused parameters under DDP")."""
import os

import torch
import torch.distributed as dist



class FlatParams:
    """Re-homes the selected parameters of `module` as views of one flat fp32 buffer and gives them gradient views
    of a second flat buffer, so the all-reduce and the optimizer each touch a single contiguous tensor."""

    def __init__(self, module, used=lambda name: True, group_of=None, early_of=None):
        """`group_of(name)`: optimizer param group id (groups become contiguous ranges).  `early_of(name)`: True for parameters
        whose gradient is complete EARLY in the backward pass; inside every group they are placed first, so that the early
        part of each group is one contiguous range that can be all-reduced while the rest of the backward still runs
        (`early_ranges` / `late_ranges`)."""
        self.named = [(k, p) for k, p in module.named_parameters() if p.requires_grad and used(k)]
        self.skipped = [k for k, p in module.named_parameters() if p.requires_grad and not used(k)]
        if not self.named:
            raise ValueError("FlatParams: no parameters selected")
        if group_of is not None or early_of is not None:      # optimizer param groups become contiguous ranges of the buffer
            gkey = group_of if group_of is not None else (lambda k: 0)
            ekey = (lambda k: 0 if early_of(k) else 1) if early_of is not None else (lambda k: 0)
            self.named.sort(key=lambda kp: (gkey(kp[0]), ekey(kp[0])))
        dev = self.named[0][1].device
        pad4 = lambda n: (n + 3) & ~3              # every view starts 16-byte aligned (the kernels read biases / gains as float4)
        self.numel = sum(pad4(p.numel()) for _, p in self.named)
        self.param = torch.zeros((self.numel,), dtype=torch.float32, device=dev)
        self.grad = torch.zeros((self.numel,), dtype=torch.float32, device=dev)
        self.offsets = {}
        self.epoch = [0]                              # bumped by the optimizer; PackCache reads it through the parameter
        off = 0
        with torch.no_grad():
            for k, p in self.named:
                p._segmif_epoch = self.epoch
                n = p.numel()
                self.param[off:off + n].copy_(p.detach().reshape(-1))
                p.data = self.param[off:off + n].view(p.shape)
                p.grad = self.grad[off:off + n].view(p.shape)
                self.offsets[k] = (off, n)
                off += pad4(n)
        self.group_ranges = None
        if group_of is not None:
            self.group_ranges = {}
            for k, p in self.named:
                o, n = self.offsets[k]
                lo, hi = self.group_ranges.get(group_of(k), (o, o))
                self.group_ranges[group_of(k)] = (min(lo, o), max(hi, o + pad4(n)))
        # contiguous [lo, hi) ranges of the early / late parameters (at most one of each per group)
        self.early_ranges, self.late_ranges = [], [(0, self.numel)]
        if early_of is not None:
            spans = {}
            for k, p in self.named:
                o, n = self.offsets[k]
                key = ((group_of(k) if group_of is not None else 0), bool(early_of(k)))
                lo, hi = spans.get(key, (o, o))
                spans[key] = (min(lo, o), max(hi, o + pad4(n)))
            self.early_ranges = sorted(v for (g, e), v in spans.items() if e)
            self.late_ranges = sorted(v for (g, e), v in spans.items() if not e)

    def broadcast(self, module=None, group=None, src=0):
        """Makes every replica start from rank `src`'s parameters (and `module`'s buffers, e.g. BatchNorm running
        statistics): random initialisation or construction order may differ between ranks, and replicas that start apart
        diverge silently because only gradients are exchanged afterwards."""
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
            return False
        dist.broadcast(self.param, src=src, group=group)
        if module is not None:
            owned = {p.data_ptr() for _, p in self.named}
            for t in list(module.parameters()) + list(module.buffers()):
                if t.data_ptr() not in owned and t.numel() > 0:
                    dist.broadcast(t.data, src=src, group=group)
        self.epoch[0] += 1                            # packed copies of the old values are stale
        return True

    def zero_grad(self):
        self.grad.zero_()
        for k, p in self.named:                       # autograd may have replaced a view; re-attach if so
            off, n = self.offsets[k]
            if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + off * 4:
                p.grad = self.grad[off:off + n].view(p.shape)

    def all_reduce(self, group=None, ranges=None):
        """The step's collective: SUM over ranks (the mean's 1/world is applied by the optimizer kernel).  ranges=None: the
        whole flat buffer in ONE call; otherwise only the given [lo, hi) ranges (the overlapped two-bucket schedule)."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            if ranges is None:
                dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=group)
            else:
                for lo, hi in ranges:
                    if hi > lo:
                        dist.all_reduce(self.grad[lo:hi], op=dist.ReduceOp.SUM, group=group)
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
    """PolyWarmupAdamW / PolyWarmupAdamW_seg (utils/optimizer.py:3-33,36-66) over a FlatParams buffer: one
    segmif_adamw_step launch per param group and step.  `groups`: {group id: dict(lr=, weight_decay=)} matching the
    FlatParams' group_of (train.py:170-189: encoder weights / encoder norms with weight_decay 0 / decoder at 10x lr);
    None = one group.  `iter_curr` is PolyWarmupAdamW_seg's starting global step (the schedule position); AdamW's own
    bias-correction step count starts at 1 either way, as torch.optim.AdamW's per-parameter state does."""

    def __init__(self, flat, lr, weight_decay, betas, warmup_iter, max_iter, warmup_ratio, power, eps=1e-8, groups=None,
                 iter_curr=0):
        self.flat = flat
        self.betas, self.eps = tuple(betas), float(eps)
        self.warmup_iter, self.max_iter, self.warmup_ratio, self.power = warmup_iter, max_iter, warmup_ratio, power
        self.global_step = int(iter_curr)
        self.opt_steps = 0
        if groups is None:
            self.groups = [dict(range=(0, flat.numel), base_lr=float(lr), lr=float(lr), weight_decay=float(weight_decay))]
        else:
            self.groups = [dict(range=flat.group_ranges[gid], base_lr=float(g.get("lr", lr)), lr=float(g.get("lr", lr)),
                                weight_decay=float(g.get("weight_decay", weight_decay)))
                           for gid, g in sorted(groups.items()) if gid in flat.group_ranges]
        self.exp_avg = torch.zeros_like(flat.param)
        self.exp_avg_sq = torch.zeros_like(flat.param)

    @property
    def lr(self):
        return self.groups[0]["lr"]

    @property
    def base_lr(self):
        return self.groups[0]["base_lr"]

    @property
    def weight_decay(self):
        return self.groups[0]["weight_decay"]

    def step(self, grad_scale=1.0):
        from . import ops
        for g in self.groups:
            g["lr"] = poly_warmup_lr(g["base_lr"], self.global_step, self.warmup_iter, self.max_iter, self.warmup_ratio,
                                     self.power, g["lr"])
            lo, hi = g["range"]
            ops.adamw_step(self.flat.param[lo:hi], self.flat.grad[lo:hi], self.exp_avg[lo:hi], self.exp_avg_sq[lo:hi],
                           lr=g["lr"], beta1=self.betas[0], beta2=self.betas[1], eps=self.eps,
                           weight_decay=g["weight_decay"], step=self.opt_steps + 1, grad_scale=grad_scale)
        self.flat.epoch[0] += 1                       # the kernel wrote through raw pointers: packed copies are stale
        self.global_step += 1
        self.opt_steps += 1


class _GraphedStep:
    """Optional CUDA-graph form of a trainer's forward + backward: `capture(*example_inputs)` records zero_grad, the
    forward, the loss and the whole reverse pass (about 850 kernel launches for train_seg, each a Python -> ctypes ->
    cudaLaunchKernel round trip in eager mode) into one graph over static input buffers; `step` then copies the batch
    into those buffers and replays.  The gradient all-reduce and the AdamW launches stay outside the graph (the
    learning rate and step count are host scalars that change every step).  Run at least one eager step first
    (lazy per-kernel initialisation is not capturable)."""

    _graph = None

    def _capture(self, body, **static_inputs):
        from . import packing
        self._static = {k: v.clone() for k, v in static_inputs.items() if v is not None}
        torch.cuda.synchronize()
        # Every packed (bf16) weight copy must be REBUILT INSIDE the graph, otherwise replays keep reading the copies of
        # capture time while AdamW updates the fp32 masters: declare all of them stale before capturing, and again afterwards
        # so that eager code never aliases tensors living in the graph's private memory pool.
        packing.invalidate_all()
        self.flat.epoch[0] += 1
        from . import _lib
        graph = torch.cuda.CUDAGraph()
        l0 = _lib.launch_count
        with torch.cuda.graph(graph):
            outs = body(**self._static)
        self.launches_per_replay = _lib.launch_count - l0       # segmif_b200 kernel-launching calls recorded in the graph
        packing.invalidate_all()
        self.flat.epoch[0] += 1
        self._graph, self._graph_outs = graph, outs
        return self

    def _replay(self, **inputs):
        for k, v in inputs.items():
            if v is not None:
                self._static[k].copy_(v, non_blocking=True)
        self._graph.replay()
        return tuple(o.clone() for o in self._graph_outs)

    def _finish(self):
        world = self.flat.all_reduce(self.group)
        self.opt.step(grad_scale=1.0 / world)
        self._post_step()

    def _post_step(self):
        pass


class FusionTrainer(_GraphedStep):
    """One optimisation step of train_fusion (train.py:343-386) for the fusion network, data parallel:
        fused = model2(ir, vis_ycrcb, out0, out1);  loss = criterion(ir, vis, fused, mask);  loss.backward();
        all-reduce;  AdamW.
    The encoder features out0 / out1 come from the frozen segmentation network (train.py:358-359, no_grad):
    `step` takes them as arguments, `step_images` computes them itself from `seg_net` (and can be graph-captured).

    With `seg_net` given and `with_ce`, the step is the rounds >= 2 composite (train.py:361-380, iter_ > 1):
        loss1 = criterion(...);  loss2 = seg_net._loss(YCrCb2RGB([fused, Cr, Cb]), labels, CE);
        loss = w0 * loss1 * (0.4 / iter_) + w1 * loss2 * 0.8,   (w0, w1) = 2 softmax(ratio of the two previous losses / 1000)
    The reference reads the loss history on the host with .item() every step; here the two previous loss pairs and the
    step count are device-resident (no synchronisation, capturable), weights all ones for the first 11 steps as in
    train.py:377-380.  The segmentation network is differentiated through but frozen: its weight gradients (which the
    reference computes and never uses) are skipped."""

    def __init__(self, fusion_net, criterion, lr=3e-4, weight_decay=0.01, betas=(0.9, 0.999), warmup_iter=3e-5,
                 max_iter=40000, warmup_ratio=1e-6, power=1.0, group=None, seg_net=None, iter_=1, ignore_index=255,
                 with_ce=None):
        self.net, self.criterion, self.group = fusion_net, criterion, group
        self.flat = FlatParams(fusion_net, used=lambda k: not k.startswith("ffm2."))
        self.opt = FusedPolyWarmupAdamW(self.flat, lr, weight_decay, betas, warmup_iter, max_iter, warmup_ratio, power)
        self.seg_net, self.iter_ = seg_net, iter_
        self.with_ce = (seg_net is not None) if with_ce is None else with_ce
        self.flat.broadcast(fusion_net, group)
        if seg_net is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            for t in list(seg_net.parameters()) + list(seg_net.buffers()):      # the frozen network must agree across ranks too
                if t.numel() > 0:
                    dist.broadcast(t.data, src=0, group=group)
        if self.with_ce:
            # train.py never steps the segmentation network in train_fusion; its weight gradients are skipped while these
            # flags are off.  `release_seg_net()` restores them (a later SegTrainer on the same network needs them).
            self._seg_requires_grad = [(p, p.requires_grad) for p in seg_net.parameters()]
            for p in seg_net.parameters():
                p.requires_grad_(False)
            self.ce = torch.nn.CrossEntropyLoss(ignore_index=ignore_index)
            dev = self.flat.param.device
            self.prev1 = torch.ones((2,), dtype=torch.float32, device=dev)        # losses of step n-1
            self.prev2 = torch.ones((2,), dtype=torch.float32, device=dev)        # losses of step n-2
            self.count = torch.zeros((), dtype=torch.float32, device=dev)         # n_iter

    def _post_step(self):
        # The activation backward recovers pre-activations from stored outputs, which needs the shared PReLU slope > 0
        # (csrc/train_ops.cu).  AdamW + weight decay can in principle drive it to <= 0; the kernels then emit NaN gradients,
        # and this periodic host check (one 4-byte read every 50 steps) names the cause instead of a bare NaN loss.
        if self.opt.opt_steps % 50 == 0 and not float(self.net.relu.weight.detach().reshape(-1)[0]) > 0.0:
            raise RuntimeError("segmif_b200: Fusion_Network3_ac.relu (the shared PReLU slope) reached %g <= 0; the hand-written "
                               "activation backward is only valid for a positive slope" % float(self.net.relu.weight.detach().reshape(-1)[0]))

    def _forward_backward(self, ir, vis_ycrcb, out0, out1, mask, vis_rgb=None, labels=None):
        self.flat.zero_grad()
        fused = self.net(ir, vis_ycrcb, out0, out1)
        loss = self.criterion(ir, vis_ycrcb, fused, mask)
        if self.with_ce:
            from .autograd import recompose_rgb
            loss2 = self.seg_net._loss(recompose_rgb(fused, vis_rgb, False), labels, self.ce)
            bw = 2 * torch.softmax((self.prev1 / self.prev2) / 1000.0, dim=-1)                  # train.py:373-374
            bw = torch.where(self.count > 10, bw, torch.ones_like(bw))                          # train.py:370,377-380
            cur = torch.stack([loss.detach(), loss2.detach()])
            loss = bw[0] * loss * (0.4 / self.iter_) + bw[1] * loss2 * 0.8
            self.prev2.copy_(self.prev1)
            self.prev1.copy_(cur)
            self.count += 1
        from .core import fusion_train, seg_train
        # Weight gradients on a parallel branch of the step graph: opt-in here (SEGMIF_WGRAD_SIDE_STREAM=all).  The fusion
        # network's weight-gradient kernels fill the machine by themselves, and measured on B200 the overlap gives nothing
        # (37.9 ms vs 37.6 ms in line); the segmentation network's 103 small ones gain 10 % (SegTrainer, on by default).
        if getattr(self, "_wgrad_side", "lazy") == "lazy":
            self._wgrad_side = (seg_train.SideWgrad(ir.device) if ir.is_cuda and os.environ.get("SEGMIF_WGRAD_SIDE_STREAM", "1") == "all" else None)
        fusion_train.WGRAD_SIDE = self._wgrad_side
        try:
            loss.backward()
        finally:
            fusion_train.WGRAD_SIDE = None
            if self._wgrad_side is not None:
                self._wgrad_side.join()
        return loss.detach(), fused.detach()

    def release_seg_net(self):
        """Restores the requires_grad flags this trainer cleared on the frozen segmentation network."""
        for p, flag in getattr(self, "_seg_requires_grad", []):
            p.requires_grad_(flag)
        self._seg_requires_grad = []

    def _images_body(self, ir, vis_rgb, mask, labels=None):
        from .core.model_fusion import RGB2YCrCb
        if self.seg_net is None:
            raise ValueError("FusionTrainer.step_images needs seg_net (the frozen Network3 whose encoder provides out0 / out1)")
        with torch.no_grad():
            vis = RGB2YCrCb(vis_rgb)                                                            # train.py:356
            out0, out1 = self.seg_net.denoise_net.encoder.forward_fusion(mask)                  # train.py:358-359
        return self._forward_backward(ir, vis, out0, out1, mask, vis_rgb, labels)

    def step(self, ir, vis_ycrcb, out0, out1, mask, vis_rgb=None, labels=None):
        out = self._forward_backward(ir, vis_ycrcb, out0, out1, mask, vis_rgb, labels)
        self._finish()
        return out

    def capture_images(self, ir, vis_rgb, mask, labels=None):
        return self._capture(self._images_body, ir=ir, vis_rgb=vis_rgb, mask=mask, labels=labels)

    def step_images(self, ir, vis_rgb, mask, labels=None):
        """train.py:350-381 from the loader's tensors: RGB->YCrCb, frozen-encoder features, fusion forward, loss(es),
        backward; then the all-reduce and AdamW."""
        if self._graph is not None:
            out = self._replay(ir=ir, vis_rgb=vis_rgb, mask=mask, labels=labels)
        else:
            out = self._images_body(ir, vis_rgb, mask, labels)
        self._finish()
        return out


class SegTrainer(_GraphedStep):
    """One optimisation step of train_seg (train.py:207-226), data parallel:
        _, _, segmap = model(mask);  loss = CE(bilinear(segmap -> label size), labels);  backward;  all-reduce;  AdamW
    with train.py:170-189's three param groups (WeTr.get_param_groups: encoder weights, encoder norms with weight_decay 0,
    decode head at 10x the learning rate).  WeTr.classifier.weight never receives a gradient (its output is discarded,
    core/model_fusion.py:66), so torch.optim.AdamW never touches it: it stays out of the flat buffer."""

    def __init__(self, seg_net, lr=6e-5, weight_decay=0.01, betas=(0.9, 0.999), warmup_iter=1500, max_iter=40000,
                 warmup_ratio=1e-6, power=1.0, iter_curr=0, ignore_index=255, group=None):
        self.net, self.group = seg_net, group
        gid = lambda k: 2 if ".decoder." in k else (1 if "norm" in k.split("encoder.", 1)[-1] else 0)
        # gradients of the decode head and of encoder stages 3-4 (93 % of the 94 MB) are complete when the reverse pass
        # leaves stage 3; they are all-reduced on a side stream while stages 2 and 1 (most of the backward's TIME: they hold
        # 94 % of the tokens) are still being differentiated.  Opt-in (SEGMIF_OVERLAP_ALLREDUCE=1): measured on B200 x2 and x8
        # (profiles/r2_train_seg_n8_*_v1.json: 12.33 ms overlapped vs 12.24 ms single) the NCCL channels' CTAs take SMs from
        # the stage-2/1 backward kernels for as long as they hide the 0.3 ms reduction, so the single call stays the default.
        early = lambda k: ".decoder." in k or any(t in k for t in ("patch_embed3.", "patch_embed4.", "block3.", "block4.", "norm3.", "norm4."))
        self.overlap = os.environ.get("SEGMIF_OVERLAP_ALLREDUCE", "0") == "1"
        self.flat = FlatParams(seg_net, used=lambda k: not k.endswith("classifier.weight"), group_of=gid,
                               early_of=early if self.overlap else None)
        self._side = None
        # small weight-gradient kernels on a parallel branch of the step graph (core/seg_train.py SideWgrad); 0 = in line
        self._wgrad_side = None if os.environ.get("SEGMIF_WGRAD_SIDE_STREAM", "1") == "0" else "lazy"
        self.opt = FusedPolyWarmupAdamW(self.flat, lr, weight_decay, betas, warmup_iter, max_iter, warmup_ratio, power,
                                        groups={0: dict(lr=lr, weight_decay=weight_decay), 1: dict(lr=lr, weight_decay=0.0),
                                                2: dict(lr=lr * 10, weight_decay=weight_decay)}, iter_curr=iter_curr)
        self.flat.broadcast(seg_net, group)
        self.ce = torch.nn.CrossEntropyLoss(ignore_index=ignore_index)

    def _world(self):
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def _early_hook(self, stage):
        """Called by the reverse pass (core/seg_train.py) when encoder stage `stage` (0-based) is done: after stage index 2
        every gradient of `early_ranges` is final."""
        if stage != 2 or self._world() == 1 or not self.flat.early_ranges:
            return
        cur = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream()
        self._side.wait_stream(cur)
        if self._wgrad_side is not None and self._wgrad_side != "lazy":
            self._side.wait_stream(self._wgrad_side.stream)       # weight gradients of the early ranges queued on their own stream
        with torch.cuda.stream(self._side):
            self.flat.all_reduce(self.group, ranges=self.flat.early_ranges)
        self._early_pending = True

    def _forward_backward(self, mask, labels):
        from .core import seg_train
        self.flat.zero_grad()
        loss = self.net._loss(mask, labels, self.ce)
        self._early_pending = False
        seg_train.STAGE_DONE_HOOK = self._early_hook if self.overlap else None
        if self._wgrad_side == "lazy":
            self._wgrad_side = seg_train.SideWgrad(mask.device) if mask.is_cuda else None
        seg_train.WGRAD_SIDE = self._wgrad_side
        try:
            loss.backward()
        finally:
            seg_train.STAGE_DONE_HOOK = None
            seg_train.WGRAD_SIDE = None
            if self._wgrad_side is not None:
                self._wgrad_side.join()
        if self._early_pending:                    # join (inside the captured graph when capturing)
            torch.cuda.current_stream().wait_stream(self._side)
        return (loss.detach(),)

    def _finish(self):
        if self.overlap and self.flat.early_ranges and self._world() > 1:
            world = self.flat.all_reduce(self.group, ranges=self.flat.late_ranges)       # stages 1-2: ~7 % of the bytes
        else:
            world = self.flat.all_reduce(self.group)
        self.opt.step(grad_scale=1.0 / world)
        self._post_step()

    def capture(self, mask, labels):
        return self._capture(self._forward_backward, mask=mask, labels=labels)

    def step(self, mask, labels):
        if self._graph is not None:
            (loss,) = self._replay(mask=mask, labels=labels)
        else:
            (loss,) = self._forward_backward(mask, labels)
        self._finish()
        return loss
