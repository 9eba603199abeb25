"""CPU restatement (numpy) of the reference's TRAINING DATA PATH -- SURVEY.md 8(f1).

TEST INFRASTRUCTURE ONLY: imported by tests/, by __graft_entry__.smoke() and by bench.py's cpu_baseline leg as the checker of the
CUDA data path (segmif_b200/csrc/datapath.cu, segmif_b200/datasets/).  The product never imports this module.

What is restated, and from where:
  * the reference's own glue                    /root/reference/datasets/imutils.py:34-49,69-91,121-129,199-249,295-391
                                                /root/reference/datasets/voc_fusion3.py:169-209
  * Pillow `Image.resize(BILINEAR | NEAREST)`   third-party (requirements.txt:85 pins pillow 8.4.0; this image has 12.2.0):
                                                src/libImaging/Resample.c (precompute_coeffs, normalize_coeffs_8bpc,
                                                ImagingResampleHorizontal_8bpc / Vertical_8bpc, PRECISION_BITS = 32-8-2) and
                                                src/libImaging/Geometry.c (nearest: affine walk with an accumulated coordinate)
  * OpenCV `cvtColor(BGR2HSV | HSV2BGR)`        third-party, reached through mmcv.bgr2hsv / mmcv.hsv2bgr (requirements.txt:71,81
                                                pin mmcv 1.7.1 and opencv-python 4.5.4.58; this image has OpenCV 4.13.0, mmcv is
                                                absent and is exactly `cv2.cvtColor(img, cv2.COLOR_BGR2HSV / HSV2BGR)`):
                                                modules/imgproc/src/color_hsv.simd.hpp (RGB2HSV_b with the sdiv/hdiv tables and
                                                hsv_shift 12; RGB2HSV_f; HSV2RGB_f / HSV2RGB_b through the float path)

PINNING (tests/test_datapath_oracle.py): every function below is bit-exact against Pillow and OpenCV as installed in this image
(exhaustively for the 8-bit colour conversions), against the reference's own imutils functions imported from /root/reference
with the same RNG seeds (oracle/make_golden_datapath.py -> tests/golden/datapath.npz), and the fixture travels to the GPU box.
"""
import random as _py_random

import numpy as np

f32 = np.float32
PRECISION_BITS = 32 - 8 - 2          # Resample.c


# --------------------------------------------------------------------------------------------------- Pillow resize restated
def bilinear_coeffs(in_size, out_size):
    """precompute_coeffs + normalize_coeffs_8bpc for the triangle filter (support 1.0) over the whole axis.
    Returns (xmin[out], xcnt[out], k[out, ksize] int32): out[xx] = clip8((2^21 + sum_j k[xx,j] * in[xmin[xx]+j]) >> 22)."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    ss = 1.0 / filterscale
    xmin = np.zeros(out_size, np.int32)
    xcnt = np.zeros(out_size, np.int32)
    kk = np.zeros((out_size, ksize), np.float64)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        lo = int(center - support + 0.5)
        lo = max(lo, 0)
        hi = int(center + support + 0.5)
        hi = min(hi, in_size)
        n = hi - lo
        x = np.arange(n)
        w = np.maximum(0.0, 1.0 - np.abs((x + lo - center + 0.5) * ss))
        ww = 0.0
        for v in w:                                   # the C loop accumulates in this order
            ww += v
        if ww != 0.0:
            w = w / ww
        kk[xx, :n] = w
        xmin[xx], xcnt[xx] = lo, n
    ik = np.where(kk < 0, (-0.5 + kk * (1 << PRECISION_BITS)).astype(np.int64), (0.5 + kk * (1 << PRECISION_BITS)).astype(np.int64))
    return xmin, xcnt, ik.astype(np.int32)


def _resample_axis(img, out_size, axis):
    xmin, xcnt, ik = bilinear_coeffs(img.shape[axis], out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], np.uint8)
    for xx in range(out_size):
        n = xcnt[xx]
        acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(ik[xx, :n].astype(np.int64), src[xmin[xx]:xmin[xx] + n], axes=(0, 0))
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_bilinear_u8(img, new_w, new_h):
    """`Image.fromarray(img).resize((new_w, new_h), resample=Image.BILINEAR)` for uint8 HW or HWC: the horizontal pass runs
    first and is rounded to uint8 before the vertical pass (ImagingResample)."""
    h, w = img.shape[:2]
    t = _resample_axis(img, new_w, 1) if new_w != w else img
    return _resample_axis(t, new_h, 0) if new_h != h else t


def nearest_index(in_size, out_size):
    """Geometry.c affine_transform + nearest filter: xo starts at a0*0.5 and is ACCUMULATED (`xo += a0`) in double, index =
    (int) xo (COORD truncation), columns outside the source are left untouched (cannot happen for a pure scale)."""
    a0 = in_size / out_size
    xo = a0 * 0.5
    idx = np.empty(out_size, np.int32)
    for x in range(out_size):
        idx[x] = int(xo)
        xo += a0
    return np.clip(idx, 0, in_size - 1)


def resize_nearest(lab, new_w, new_h):
    """`Image.fromarray(lab).resize((new_w, new_h), resample=Image.NEAREST)`."""
    h, w = lab.shape[:2]
    return lab[nearest_index(h, new_h)][:, nearest_index(w, new_w)]


# --------------------------------------------------------------------------------------------------- OpenCV HSV restated
HSV_SHIFT = 12


def _hsv_tables():
    sdiv = np.zeros(256, np.int64)
    hdiv = np.zeros(256, np.int64)
    for i in range(1, 256):
        sdiv[i] = int(np.rint((255 << HSV_SHIFT) / (1.0 * i)))
        hdiv[i] = int(np.rint((180 << HSV_SHIFT) / (6.0 * i)))
    return sdiv, hdiv


SDIV, HDIV = _hsv_tables()
_SECTOR = np.array([[1, 3, 0], [1, 0, 2], [3, 0, 1], [0, 2, 1], [0, 1, 3], [2, 1, 0]])


def _fma(a, b, c):
    """float32 fused multiply-add (exact in double for float32 operands, one rounding)."""
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(np.float32)


def bgr2hsv_u8(img):
    """RGB2HSV_b, hrange 180: integer tables, round-to-nearest shifts."""
    b, g, r = (img[..., i].astype(np.int64) for i in range(3))
    v = np.maximum(np.maximum(b, g), r)
    vmin = np.minimum(np.minimum(b, g), r)
    diff = v - vmin
    vr = np.where(v == r, -1, 0)
    vg = np.where(v == g, -1, 0)
    s = (diff * SDIV[v] + (1 << (HSV_SHIFT - 1))) >> HSV_SHIFT
    h = (vr & (g - b)) + (~vr & ((vg & (b - r + 2 * diff)) + ((~vg) & (r - g + 4 * diff))))
    h = (h * HDIV[diff] + (1 << (HSV_SHIFT - 1))) >> HSV_SHIFT
    h = h + np.where(h < 0, 180, 0)
    return np.stack([h, s, v], -1).astype(np.uint8)


def _hsv2bgr_core(hh, s, v):
    """HSV2RGB_native on hh = h * 6/hrange (>= 0 here): sector = floor(hh) mod 6, f = hh - floor(hh); the vector path's
    products are t1 = v(1-s), t2 = v*fma(-s,f,1), t3 = v*fma(-s,1-f,1)."""
    pre = np.floor(hh)
    f = (hh - pre).astype(np.float32)
    sector = pre.astype(np.int64) % 6
    one = np.ones_like(f)
    tabs = np.stack([v, v * (one - s), v * _fma(-s, f, one), v * _fma(-s, one - f, one)], -1)
    idx = _SECTOR[sector]
    return np.stack([np.take_along_axis(tabs, idx[..., k:k + 1], -1)[..., 0] for k in range(3)], -1)


def _tail_mask(shape, lanes):
    """True for the columns OpenCV's row loop leaves to its scalar tail: x >= (W // lanes) * lanes.  The AVX2 dispatch this
    image's OpenCV runs processes 32 uint8 pixels (8 float pixels) per vector step; body and tail round differently, so a
    pixel's result depends on its COLUMN -- measured behaviour of the library the reference calls, pinned exhaustively in
    tests/test_datapath_oracle.py."""
    w = shape[-2]
    return (np.arange(w) >= (w // lanes) * lanes).reshape((1,) * (len(shape) - 2) + (w,))


def hsv2bgr_u8(img):
    """HSV2RGB_b: h*(6/180), s/255, v/255 through the float kernel, then *255; the 32-pixel vector body TRUNCATES, the scalar
    tail rounds to nearest even (saturate_cast)."""
    h = img[..., 0].astype(f32) * f32(6.0 / 180.0)
    s = img[..., 1].astype(f32) * f32(1 / 255.0)
    v = img[..., 2].astype(f32) * f32(1 / 255.0)
    out = _hsv2bgr_core(h, s, v) * f32(255.0)
    out = np.where(_tail_mask(img.shape, 32)[..., None], np.rint(out), np.floor(out))
    return np.clip(out, 0, 255).astype(np.uint8)


def bgr2hsv_f32(img):
    """RGB2HSV_f, hrange 360, on float32 BGR of any scale (the reference feeds 0..255 floats)."""
    b, g, r = img[..., 0], img[..., 1], img[..., 2]
    v = np.maximum(np.maximum(b, g), r)
    vmin = np.minimum(np.minimum(b, g), r)
    diff = v - vmin
    eps = f32(np.finfo(np.float32).eps)
    s = diff / (np.abs(v) + eps)
    d = f32(60.0) / (diff + eps)
    hr = (g - b) * d
    hg = _fma(b - r, d, f32(120))
    hb = _fma(r - g, d, f32(240))
    h = np.where(v == r, hr, np.where(v == g, hg, hb))
    # negative hue (only the v == r branch can be): the 8-pixel vector body adds 360 inside the fma, the scalar tail after it
    wrapped = np.where((v == r) & ~_tail_mask(img.shape, 8), _fma(g - b, d, f32(360)), h + f32(360))
    h = np.where(h < 0, wrapped, h)
    return np.stack([h, s, v], -1).astype(np.float32)


def hsv2bgr_f32(hsv):
    """HSV2RGB_f, hrange 360; s == 0 returns (v, v, v)."""
    h, s, v = hsv[..., 0], hsv[..., 1], hsv[..., 2]
    out = _hsv2bgr_core(h * f32(6.0 / 360.0), s, v)
    return np.where((s == 0)[..., None], v[..., None], out).astype(np.float32)


def bgr2hsv(img):
    return bgr2hsv_u8(img) if img.dtype == np.uint8 else bgr2hsv_f32(img)


def hsv2bgr(img):
    return hsv2bgr_u8(img) if img.dtype == np.uint8 else hsv2bgr_f32(img)


# --------------------------------------------------------------------------------------------------- imutils.py restated
class Rng:
    """The two global generators the reference draws from (`random` and `np.random`), passed explicitly so a test can hand
    the oracle and the CUDA path identical, independent streams."""

    def __init__(self, py=None, npr=None):
        self.py = py if py is not None else _py_random
        self.np = npr if npr is not None else np.random

    @staticmethod
    def seeded(seed):
        return Rng(_py_random.Random(seed), np.random.RandomState(seed))


def convert(img, alpha=1, beta=0):
    """imutils.py:308-312 -- float32 multiply, float32 add (two roundings), clip, truncate to uint8."""
    out = img.astype(np.float32) * alpha + beta
    return np.clip(out, 0, 255).astype(np.uint8)


def photometric_distortion(img, rng, brightness_delta=32, contrast_range=(0.5, 1.5), saturation_range=(0.5, 1.5), hue_delta=18):
    """imutils.py:295-380 (PhotoMetricDistortion.__call__).  NOTE the dtype state machine the reference really has: a float32
    image (after random_scaling2) becomes uint8 only when a `convert` fires; on a float32 image the saturation branch writes
    convert()'s uint8 (0 or 1) into the float S plane and the hue branch writes (int(H) + delta) % 180 into a 0..360 plane."""
    def contrast(x):
        if rng.np.randint(2):
            return convert(x, alpha=rng.py.uniform(*contrast_range))
        return x
    if rng.np.randint(2):                                                       # brightness :314-321
        img = convert(img, beta=rng.py.uniform(-brightness_delta, brightness_delta))
    mode = rng.np.randint(2)                                                    # :364
    if mode == 1:
        img = contrast(img)
    if rng.np.randint(2):                                                       # saturation :332-341
        hsv = bgr2hsv(img)
        hsv[:, :, 1] = convert(hsv[:, :, 1], alpha=rng.py.uniform(*saturation_range))
        img = hsv2bgr(hsv)
    if rng.np.randint(2):                                                       # hue :343-351
        hsv = bgr2hsv(img)
        hsv[:, :, 0] = (hsv[:, :, 0].astype(int) + rng.np.randint(-hue_delta, hue_delta)) % 180
        img = hsv2bgr(hsv)
    if mode == 0:
        img = contrast(img)
    return img


def random_scaling2(image, image_vis, image_mask, label, size_range, scale_range, rng):
    """imutils.py:34-49 + :69-91."""
    h, w = label.shape
    ratio = rng.py.uniform(scale_range[0], scale_range[1])
    new_w, new_h = int(ratio * w), int(ratio * h)
    outs = [resize_bilinear_u8(x.astype(np.uint8), new_w, new_h).astype(np.float32) for x in (image, image_vis, image_mask)]
    return outs[0], outs[1], outs[2], resize_nearest(label, new_w, new_h)


def random_fliplr2(image, image_vis, image_mask, label, rng):
    """imutils.py:121-129."""
    if rng.py.random() > 0.5:
        label, image, image_vis, image_mask = (np.fliplr(x) for x in (label, image, image_vis, image_mask))
    return image, image_vis, image_mask, label


def random_crop2(image, image_vis, image_mask, label, crop_size, mean_rgb, ignore_index, rng):
    """imutils.py:199-249: pad to >= crop_size with mean_rgb / ignore_index at a random offset, then up to ten candidate
    windows; a window is accepted when it holds a non-ignored class and the largest class fills < 75 % of its non-ignored
    pixels; the LAST candidate is kept when none is accepted."""
    h, w = label.shape
    H, W = max(crop_size, h), max(crop_size, w)
    pads = []
    for src in (image, image_vis, image_mask):
        p = np.zeros((H, W, 3), np.float32)
        p[:, :, 0], p[:, :, 1], p[:, :, 2] = mean_rgb
        pads.append(p)
    pad_label = np.ones((H, W), np.float32) * ignore_index
    H_pad = int(rng.np.randint(H - h + 1))
    W_pad = int(rng.np.randint(W - w + 1))
    for p, src in zip(pads, (image, image_vis, image_mask)):
        p[H_pad:H_pad + h, W_pad:W_pad + w, :] = src
    pad_label[H_pad:H_pad + h, W_pad:W_pad + w] = label
    for _ in range(10):
        hs = rng.py.randrange(0, H - crop_size + 1, 1)
        ws = rng.py.randrange(0, W - crop_size + 1, 1)
        index, cnt = np.unique(pad_label[hs:hs + crop_size, ws:ws + crop_size], return_counts=True)
        cnt = cnt[index != ignore_index]
        if len(cnt) and np.max(cnt) / np.sum(cnt) < 0.75:
            break
    sl = (slice(hs, hs + crop_size), slice(ws, ws + crop_size))
    return pads[0][sl], pads[1][sl], pads[2][sl], pad_label[sl]


def transforms(image, image_vis, image_mask, label, rng, aug=True, rescale_range=(0.5, 2.0), resize_range=(512, 640), crop_size=512,
               img_fliplr=True, ignore_index=255):
    """voc_fusion3.py:169-209 (`VOC12SegDataset.__transforms`): image / image_mask are the single-channel planes replicated to
    three channels (:40-48), image_vis is H x W x 3; returns CHW arrays and the label."""
    if aug:
        if rescale_range:
            image, image_vis, image_mask, label = random_scaling2(image, image_vis, image_mask, label, resize_range, rescale_range, rng)
        if img_fliplr:
            image, image_vis, image_mask, label = random_fliplr2(image, image_vis, image_mask, label, rng)
        image_vis = photometric_distortion(image_vis, rng)
        if crop_size:
            image, image_vis, image_mask, label = random_crop2(image, image_vis, image_mask, label, crop_size,
                                                               [123.675, 116.28, 103.53], ignore_index, rng)
    image, image_vis, image_mask = image / 255.0, image_vis / 255.0, image_mask / 255.0
    return (np.transpose(image, (2, 0, 1)), np.transpose(image_vis, (2, 0, 1)), np.transpose(image_mask, (2, 0, 1)), label)


# --------------------------------------------------------------------------------------------------- seeded synthetic samples
def synth_sample(seed, h, w, n_class=9, ignore_frac=0.03, mask_channels=1):
    """One decoded training sample as the dataset class holds it after imread (voc_fusion3.py:36-55): uint8 infrared H x W,
    visible H x W x 3, mask H x W, label H x W.  The label is made of rectangles so that crop windows dominated by one class
    (the retry branch of random_crop2) and windows of ignore_index both occur."""
    rs = np.random.RandomState(seed)
    ir = rs.randint(0, 256, size=(h, w)).astype(np.uint8)
    vis = rs.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
    # smooth-ish content in half of the visible image so HSV sectors / grey pixels (s == 0) are hit as well
    vis[: h // 2] = (vis[: h // 2] // 64) * 64
    vis[:, : w // 4, 1] = vis[:, : w // 4, 0]
    vis[:, : w // 8, 2] = vis[:, : w // 8, 0]
    mask = (rs.rand(h, w) > 0.7).astype(np.uint8) * 255
    label = np.zeros((h, w), np.uint8)
    for _ in range(6):
        y0, x0 = rs.randint(0, h), rs.randint(0, w)
        y1, x1 = min(h, y0 + rs.randint(4, h)), min(w, x0 + rs.randint(4, w))
        label[y0:y1, x0:x1] = rs.randint(0, n_class)
    label[rs.rand(h, w) < ignore_frac] = 255
    if mask_channels == 3:          # the fused RGB image train_seg reads back as its "mask" (voc_fusion2.py:44-48); own stream
        mask = np.random.RandomState(seed + 7919).randint(0, 256, size=(h, w, 3)).astype(np.uint8)
    return ir, vis, mask, label


def dataset_views(ir, vis, mask):
    """voc_fusion3.py:39-48: single-channel planes replicated to three channels; voc_fusion2.py:44-48 keeps a three-channel
    mask image as it is."""
    return np.repeat(ir[:, :, None], 3, 2), vis, (np.repeat(mask[:, :, None], 3, 2) if mask.ndim == 2 else mask)
