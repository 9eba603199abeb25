"""The unit of work of the headline metric as one public call:

    forward_fusion(mask) -> Fusion_Network3_ac(ir, YCrCb(vis), out0, out1) -> colour recompose + clamp
    -> Network3 -> bilinear upsample -> argmax

i.e. what train.py:356-366 + test_fusion.py:100-111 + test_segmentation.py:169-175 of the reference do for
one batch of IR / visible image pairs, built only from the reference-named modules of segmif_b200.core.

`capture()` records the ~380 kernel launches of one step into a CUDA graph (static shapes, static buffers), so a
step costs one graph launch instead of ~380 Python->ctypes->cudaLaunch round trips."""
import torch

from . import ops
from .core.model_fusion import RGB2YCrCb


class FusionSegPipeline:
    def __init__(self, seg_net, fusion_net):
        self.seg = seg_net.eval()
        self.fus = fusion_net.eval()
        self._pinned = {}
        self._graph = None
        self._static_in = None
        self._static_out = None

    @torch.no_grad()
    def __call__(self, ir, vis_rgb, mask, return_intermediates=False):
        out0, out1 = self.seg.denoise_net.encoder.forward_fusion(mask)
        vis_ycc = RGB2YCrCb(vis_rgb)                                 # train.py:356 (the fusion net reads Y)
        fused = self.fus(ir, vis_ycc, out0, out1)                     # [B,1,H,W] fp32
        rgb = ops.recompose_rgb(fused, vis_rgb, clamp=True)           # train.py:364-366 + clamp test_fusion.py:108-111
        lg = self.seg.logits_pixel_major(rgb)                         # [B,h,w,nc] fp32
        B, h, w, nc = lg.shape
        labels = ops.upsample_argmax(lg, B, h, w, nc, ir.shape[2], ir.shape[3])
        if not return_intermediates:
            return fused, labels
        logits = ops.nhwc_to_nchw(lg, B, h * w, nc).view(B, nc, h, w)
        return dict(out0=out0, out1=out1, fused=fused, rgb=rgb, logits=logits, labels=labels)

    # ---- CUDA-graph path -------------------------------------------------------------------------------
    @torch.no_grad()
    def capture(self, batch, height, width, device):
        """Records one step for inputs of this shape.  Afterwards `static_inputs` are the buffers to fill and
        `replay()` runs the step; outputs live in static buffers that the next replay overwrites."""
        dev = torch.device(device)
        mk = lambda c: torch.zeros((batch, c, height, width), dtype=torch.float32, device=dev)
        self._static_in = dict(ir=mk(1), vis=mk(3), mask=mk(3))
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                                  # warm-up: weight packing, smem opt-ins, allocator
            for _ in range(2):
                self(self._static_in["ir"], self._static_in["vis"], self._static_in["mask"])
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self._static_out = self(self._static_in["ir"], self._static_in["vis"], self._static_in["mask"])
        self._graph = graph
        return self._static_in

    @property
    def static_inputs(self):
        return self._static_in

    def replay(self):
        if self._graph is None:
            raise RuntimeError("FusionSegPipeline.replay: call capture() first")
        self._graph.replay()
        return self._static_out

    # ---- host-buffer entry point (what bench.py's e2e leg times) --------------------------------------
    def _pin(self, name, like):
        t = self._pinned.get(name)
        if t is None or t.shape != like.shape or t.dtype != like.dtype:
            t = torch.empty(like.shape, dtype=like.dtype, pin_memory=True)
            self._pinned[name] = t
        return t

    @torch.no_grad()
    def run_host(self, ir_host, vis_host, mask_host, device):
        """Inputs: pinned (or pageable) CPU tensors.  Copies them to `device`, runs the pipeline (through the
        captured graph when capture() was called for this shape) and copies the fused image and the label map back
        into pinned host buffers.  Returns (fused_host, labels_host); the caller must synchronise the current stream
        before reading them."""
        si = self._static_in
        if self._graph is not None and si["ir"].shape == ir_host.shape and si["vis"].shape == vis_host.shape:
            si["ir"].copy_(ir_host, non_blocking=True)
            si["vis"].copy_(vis_host, non_blocking=True)
            si["mask"].copy_(mask_host, non_blocking=True)
            fused, labels = self.replay()
        else:
            ir = ir_host.to(device, non_blocking=True)
            vis = vis_host.to(device, non_blocking=True)
            mask = mask_host.to(device, non_blocking=True)
            fused, labels = self(ir, vis, mask)
        fh = self._pin("fused", fused)
        lh = self._pin("labels", labels)
        fh.copy_(fused, non_blocking=True)
        lh.copy_(labels, non_blocking=True)
        return fh, lh
