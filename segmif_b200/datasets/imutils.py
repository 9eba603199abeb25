"""Training data path on the device -- the host half.

Mirror of /root/reference/datasets/imutils.py as used by datasets/voc_fusion3.py:169-209: the random draws are made HERE, from
Python's `random` and `np.random`, in exactly the order and with exactly the calls the reference makes (so a seeded run picks the
same scale, flip, distortion, canvas offset and crop window, and leaves both generators in the same state), and everything that
touches pixels runs in segmif_b200/csrc/datapath.cu.  Intermediate arrays of the reference (the resized float32 images, the HSV
planes, the padded canvases) are never materialised: the unit of work is the whole `__transforms` of one batch.

    draws (reference order, per sample)                                                          reference
      ratio     = random.uniform(*scale_range)                                                   imutils.py:40
      flip      = random.random() > 0.5                                                          :123
      distortion: np.random.randint(2) [brightness] -> random.uniform(-32, 32)                   :316-320
                  np.random.randint(2) [mode]                                                    :364
                  contrast  (mode 1 here / mode 0 last): np.random.randint(2) -> random.uniform  :325-329
                  saturation: np.random.randint(2) -> random.uniform(0.5, 1.5)                   :334-339
                  hue:        np.random.randint(2) -> np.random.randint(-18, 18)                 :345-349
      H_pad, W_pad = np.random.randint(H - h + 1), np.random.randint(W - w + 1)                  :211-212
      up to 10 x (random.randrange(0, H - crop + 1, 1), random.randrange(0, W - crop + 1, 1))    :225-228
    The number of candidate windows the reference draws depends on the label (it stops at the first acceptable one): all ten
    are drawn and evaluated on the device in one launch, and `random` is then rewound to the state after the accepted draw.
"""
import ctypes
import math
import random as _py_random

import numpy as np
import torch

from .. import _lib

OP_CONVERT, OP_SATURATION, OP_HUE = 0, 1, 2
MEAN_RGB = (123.675, 116.28, 103.53)          # voc_fusion3.py:188


class Rng:
    """The two generators the reference draws from.  Default: the process-global `random` and `np.random` modules, like the
    reference; pass `random.Random(seed)` / `np.random.RandomState(seed)` for an independent stream per sample or worker."""

    def __init__(self, py=None, npr=None):
        self.py = py if py is not None else _py_random
        self.np = npr if npr is not None else np.random

    @staticmethod
    def seeded(seed):
        return Rng(_py_random.Random(seed), np.random.RandomState(seed))


class PhotoMetricDistortion:
    """imutils.py:295-391 -- holds the ranges and draws one distortion PROGRAM (list of (kind, alpha, beta, delta, on_uint8));
    the pixels are transformed by the `finish` kernel."""

    def __init__(self, brightness_delta=32, contrast_range=(0.5, 1.5), saturation_range=(0.5, 1.5), hue_delta=18):
        self.brightness_delta = brightness_delta
        self.contrast_lower, self.contrast_upper = contrast_range
        self.saturation_lower, self.saturation_upper = saturation_range
        self.hue_delta = hue_delta

    def draw(self, rng, is_uint8):
        """`is_uint8`: dtype of the image entering __call__ (float32 after random_scaling2, uint8 otherwise).  A convert()
        turns the image into uint8 for every later op (:308-312)."""
        ops = []
        state = {"u8": bool(is_uint8)}

        def convert(alpha=1.0, beta=0.0):
            ops.append((OP_CONVERT, float(np.float32(alpha)), float(np.float32(beta)), 0, state["u8"]))
            state["u8"] = True

        def contrast():
            if rng.np.randint(2):
                convert(alpha=rng.py.uniform(self.contrast_lower, self.contrast_upper))

        if rng.np.randint(2):
            convert(beta=rng.py.uniform(-self.brightness_delta, self.brightness_delta))
        mode = rng.np.randint(2)
        if mode == 1:
            contrast()
        if rng.np.randint(2):
            ops.append((OP_SATURATION, float(np.float32(rng.py.uniform(self.saturation_lower, self.saturation_upper))), 0.0, 0, state["u8"]))
        if rng.np.randint(2):
            ops.append((OP_HUE, 1.0, 0.0, int(rng.np.randint(-self.hue_delta, self.hue_delta)), state["u8"]))
        if mode == 0:
            contrast()
        return ops

    def __repr__(self):
        return (f"{self.__class__.__name__}(brightness_delta={self.brightness_delta}, contrast_range=({self.contrast_lower}, "
                f"{self.contrast_upper}), saturation_range=({self.saturation_lower}, {self.saturation_upper}), hue_delta={self.hue_delta})")


def _ksize(in_size, out_size):
    return int(math.ceil(max(in_size / out_size, 1.0))) * 2 + 1          # Pillow precompute_coeffs


def _p16(w):
    return (w + 15) & ~15          # row pitch of the uint8 intermediates (datapath.cu pitch16)


def _src_rows(in_size, out_size, y0, y1):
    """Source rows [lo, hi) that output rows [y0, y1) of the bilinear pass read (same double arithmetic as the kernel)."""
    scale = in_size / out_size
    support = max(scale, 1.0)
    lo = max(int((y0 + 0.5) * scale - support + 0.5), 0)
    hi = min(int((y1 - 1 + 0.5) * scale + support + 0.5), in_size)
    return lo, hi


def _check_plane(t, shape_len, what):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.uint8 and t.dim() == shape_len and t.is_contiguous()):
        raise ValueError(f"segmif_b200.datasets: {what} must be a contiguous uint8 CUDA tensor with {shape_len} dimensions")


class DeviceTransforms:
    """`VOC12SegDataset.__transforms` (voc_fusion3.py:169-209, aug=True) for a BATCH of decoded samples resident in HBM.

    __call__(samples, rng) -> image [n,3,c,c], image_vis [n,3,c,c], image_mask [n,3,c,c] (fp32, CHW, / 255.0) and label
    [n,c,c] (fp32 like the reference's pad_label; `label_int64=True` adds the int64 tensor the loss consumes).
      samples: list of (ir uint8 [H,W], vis uint8 [H,W,3], mask uint8 [H,W] or [H,W,3], label uint8 [H,W]) CUDA tensors -- the
               single-channel planes are NOT replicated to three channels on the way in (voc_fusion3.py:40-48 does; the result
               is the same three identical planes, except over the canvas where each channel takes its mean_rgb value).  A
               three-channel mask is the fused RGB image train_seg reads back (voc_fusion2.py:44-48).
      rng:     one Rng shared by all samples (the reference's worker semantics: sample k+1 continues where sample k stopped;
               costs one device->host read of 30 integers per sample) or a list with one Rng per sample (one read per batch).
    """

    def __init__(self, crop_size=512, rescale_range=(0.5, 2.0), resize_range=(512, 640), img_fliplr=True, ignore_index=255,
                 mean_rgb=MEAN_RGB, color_jittor=None):
        if not crop_size:
            raise ValueError("segmif_b200.datasets: crop_size is required (a batch needs one output size)")
        self.crop_size, self.rescale_range, self.resize_range = int(crop_size), rescale_range, resize_range
        self.img_fliplr, self.ignore_index = img_fliplr, int(ignore_index)
        self.mean = (ctypes.c_float * 3)(*[float(np.float32(m)) for m in mean_rgb])
        self.color_jittor = color_jittor if color_jittor is not None else PhotoMetricDistortion()
        self.last = None                                  # host view of the last batch's sample descriptors (tests, debugging)

    # ------------------------------------------------------------------------------------------------ draws
    def _draw(self, s, sample, rng):
        ir, vis, mask, label = sample
        h, w = label.shape
        s.ir, s.vis, s.mask, s.label = ir.data_ptr(), vis.data_ptr(), mask.data_ptr(), label.data_ptr()
        s.H, s.W = h, w
        s.mask_c = 1 if mask.dim() == 2 else 3
        if self.rescale_range:
            ratio = rng.py.uniform(self.rescale_range[0], self.rescale_range[1])      # imutils.py:40
            s.nw, s.nh, s.resized = int(ratio * w), int(ratio * h), 1                 # :73
            if s.nw <= 0 or s.nh <= 0:
                raise ValueError("segmif_b200.datasets: scaled size is empty (Pillow raises here as well)")
            s.ks_x, s.ks_y = _ksize(w, s.nw), _ksize(h, s.nh)
        else:
            s.nw, s.nh, s.resized, s.ks_x, s.ks_y = w, h, 0, 0, 0
        s.flip = int(rng.py.random() > 0.5) if self.img_fliplr else 0                 # :123
        ops = self.color_jittor.draw(rng, is_uint8=not s.resized)
        s.n_ops = len(ops)
        for i, (kind, alpha, beta, delta, u8) in enumerate(ops):
            s.op_kind[i], s.op_alpha[i], s.op_beta[i], s.op_delta[i], s.op_u8[i] = kind, alpha, beta, delta, int(u8)
        c = self.crop_size
        s.PH, s.PW = max(c, s.nh), max(c, s.nw)                                       # :202-203
        s.pad_h = int(rng.np.randint(s.PH - s.nh + 1))                                # :211-212
        s.pad_w = int(rng.np.randint(s.PW - s.nw + 1))
        state = rng.py.getstate()                                                     # rewound in _decide
        for i in range(10):                                                           # :223-228
            s.cand_hs[i] = rng.py.randrange(0, s.PH - c + 1, 1)
            s.cand_ws[i] = rng.py.randrange(0, s.PW - c + 1, 1)
        return state

    def _decide(self, s, stats, states, rng):
        """:229-233: the first candidate holding a non-ignored class whose largest class covers < 75 % of the non-ignored
        pixels, else the tenth.  max / sum < 0.75 is evaluated as 4 max < 3 sum (exact for counts < 2^29)."""
        pick = 9
        for i in range(10):
            n_val, mx, sm = (int(v) for v in stats[i])
            if n_val > 0 and 4 * mx < 3 * sm:
                pick = i
                break
        c = self.crop_size
        if pick < 9:                               # leave `random` where the reference would: after pick + 1 candidate draws
            rng.py.setstate(states)
            for _ in range(pick + 1):
                rng.py.randrange(0, s.PH - c + 1, 1)
                rng.py.randrange(0, s.PW - c + 1, 1)
        s.hs, s.ws = s.cand_hs[pick], s.cand_ws[pick]
        if s.resized:
            ys0, ys1 = max(s.hs, s.pad_h) - s.pad_h, min(s.hs + c, s.pad_h + s.nh) - s.pad_h
            xs0, xs1 = max(s.ws, s.pad_w) - s.pad_w, min(s.ws + c, s.pad_w + s.nw) - s.pad_w
            s.roi_y0, s.roi_y1 = ys0, ys1
            s.roi_x0, s.roi_x1 = (s.nw - xs1, s.nw - xs0) if s.flip else (xs0, xs1)
            s.src_y0, s.src_y1 = _src_rows(s.H, s.nh, ys0, ys1)

    # ------------------------------------------------------------------------------------------------ batch
    def __call__(self, samples, rng=None, label_int64=False):
        n = len(samples)
        if n == 0:
            raise ValueError("segmif_b200.datasets: empty batch")
        for ir, vis, mask, label in samples:
            _check_plane(label, 2, "label")
            h, w = label.shape
            for t, nd, what in ((ir, 2, "infrared"), (vis, 3, "visible"), (mask, mask.dim() if isinstance(mask, torch.Tensor) and mask.dim() == 3 else 2, "mask")):
                _check_plane(t, nd, what)
                if tuple(t.shape[:2]) != (h, w) or (nd == 3 and t.shape[2] != 3):
                    raise ValueError(f"segmif_b200.datasets: {what} has shape {tuple(t.shape)}, label is {(h, w)}")
        dev = samples[0][3].device
        _lib.ensure_init(dev.index or 0)
        stream = torch.cuda.current_stream(dev).cuda_stream
        shared = not isinstance(rng, (list, tuple))
        rngs = [rng if rng is not None else Rng()] * n if shared else list(rng)
        if len(rngs) != n:
            raise ValueError("segmif_b200.datasets: one Rng per sample expected")
        c = self.crop_size
        host = (_lib.DpSample * n)()
        nbytes = ctypes.sizeof(host)
        host_t = torch.frombuffer(host, dtype=torch.uint8)
        dev_t = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        stats = torch.empty((n, 10, 3), dtype=torch.int32, device=dev)
        hist_ws = torch.empty((n, 10, 256), dtype=torch.int32, device=dev)
        # label stage: arenas grow with the draws, so in shared-rng mode they are sized per sample
        tab_parts, lab_parts = [], []
        groups = [[k] for k in range(n)] if shared else [list(range(n))]
        for group in groups:
            states = {}
            tab_elems = lab_bytes = 0
            for k in group:
                s = host[k]
                states[k] = self._draw(s, samples[k], rngs[k])
                s.tab_off, s.lab_off = tab_elems, lab_bytes
                if s.resized:
                    tab_elems += 3 * s.nw + s.nw * s.ks_x + 3 * s.nh + s.nh * s.ks_y
                lab_bytes += s.PH * _p16(s.PW)
            tab = torch.empty(max(tab_elems, 1), dtype=torch.int32, device=dev)
            lab = torch.empty(lab_bytes, dtype=torch.uint8, device=dev)
            tab_parts.append(tab)
            lab_parts.append(lab)
            g0, gn = group[0], len(group)
            off = g0 * ctypes.sizeof(_lib.DpSample)
            dev_t[off:off + gn * ctypes.sizeof(_lib.DpSample)].copy_(host_t[off:off + gn * ctypes.sizeof(_lib.DpSample)], non_blocking=False)
            _lib.call("segmif_dp_label_stage", dev_t.data_ptr() + off, ctypes.addressof(host) + off, gn, c, self.ignore_index,
                      tab.data_ptr(), lab.data_ptr(), hist_ws[g0:g0 + gn].data_ptr(), stats[g0:g0 + gn].data_ptr(), stream)
            st = stats[g0:g0 + gn].cpu().numpy()                       # the one device->host read of the stage
            for j, k in enumerate(group):
                self._decide(host[k], st[j], states[k], rngs[k])
        # image stage: one batch; arenas of a shared-rng batch are stitched by rebasing the offsets onto absolute addresses
        if len(groups) > 1:
            tab = torch.cat(tab_parts)
            lab = torch.cat(lab_parts)
            t_acc = l_acc = 0
            for k, (tp, lp) in enumerate(zip(tab_parts, lab_parts)):
                host[k].tab_off, host[k].lab_off = t_acc, l_acc
                t_acc, l_acc = t_acc + tp.numel(), l_acc + lp.numel()
        tmp_bytes = rs_bytes = 0
        for k in range(n):
            s = host[k]
            if s.resized:
                s.tmp_off, s.rs_off = tmp_bytes, rs_bytes
                tmp_bytes += (4 + s.mask_c) * (s.src_y1 - s.src_y0) * _p16(s.roi_x1 - s.roi_x0)
                rs_bytes += (4 + s.mask_c) * (s.roi_y1 - s.roi_y0) * _p16(s.roi_x1 - s.roi_x0)
        tmp = torch.empty(max(tmp_bytes, 1), dtype=torch.uint8, device=dev)
        rs = torch.empty(max(rs_bytes, 1), dtype=torch.uint8, device=dev)
        dev_t.copy_(host_t, non_blocking=False)
        out_ir = torch.empty((n, 3, c, c), dtype=torch.float32, device=dev)
        out_vis, out_mask = torch.empty_like(out_ir), torch.empty_like(out_ir)
        out_label = torch.empty((n, c, c), dtype=torch.float32, device=dev)
        out_i64 = torch.empty((n, c, c), dtype=torch.int64, device=dev) if label_int64 else None
        _lib.call("segmif_dp_image_stage", dev_t.data_ptr(), ctypes.addressof(host), n, c, ctypes.addressof(self.mean), tab.data_ptr(),
                  tmp.data_ptr(), rs.data_ptr(), lab.data_ptr(), out_ir.data_ptr(), out_vis.data_ptr(), out_mask.data_ptr(),
                  out_label.data_ptr(), out_i64.data_ptr() if label_int64 else None, stream)
        self.last = host
        # the workspaces are consumed by kernels queued on `stream`; torch's caching allocator reuses them stream-ordered
        if label_int64:
            return out_ir, out_vis, out_mask, out_label, out_i64
        return out_ir, out_vis, out_mask, out_label


def to_chw_float64(img):
    """The aug=False tail of `__transforms` (voc_fusion3.py:193-205) on a uint8 image: `img / 255.0` is a float64 division in
    numpy; [H,W] planes come out as three identical channels (voc_fusion3.py:40-48)."""
    if img.dim() == 2:
        _check_plane(img, 2, "image")
        C = 1
    else:
        _check_plane(img, 3, "image")
        C = img.shape[2]
    h, w = img.shape[:2]
    _lib.ensure_init(img.device.index or 0)
    out = torch.empty((3, h, w), dtype=torch.float64, device=img.device)
    _lib.call("segmif_dp_u8_to_chw_f64", img.data_ptr(), h, w, C, out.data_ptr(), torch.cuda.current_stream(img.device).cuda_stream)
    return out
