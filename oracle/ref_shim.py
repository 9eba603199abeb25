"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* SegMiF reference modules from
/root/reference so that (a) the oracle restatement in `segmif_oracle.py` can be
validated against them and (b) golden fixtures under tests/golden/ can be generated
(`make_golden.py`).  Never imported by the product package `segmif_b200`.

/root/reference does not exist on the GPU box, so nothing that runs there may call
`load_reference()`; `available()` is the guard.

The reference needs three third-party helpers that are not installed here
(timm.models.layers.{DropPath,to_2tuple,trunc_normal_}, mmcv.cnn.ConvModule); we
provide minimal stand-ins with the documented semantics of timm 0.6.12 / mmcv 1.7.1
(requirements.txt:71,122 of the reference).  The reference's own core/__init__.py
imports a symbol that does not exist (core/__init__.py:4), so the modules are loaded
by file path under a synthetic `core` package instead.
"""
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

def _default_root():
    """/root/reference in the build container; on the GPU box the byte-for-byte copies oracle/build_ref.py vendored."""
    if os.path.isfile("/root/reference/core/model_fusion.py"):
        return "/root/reference"
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


REF_ROOT = os.environ.get("SEGMIF_REFERENCE_ROOT") or _default_root()
_PREFIX = "_segmif_ref"


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "core", "model_fusion.py"))


class _DropPath(nn.Module):
    """timm DropPath: per-sample stochastic depth, identity in eval."""

    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        mask = x.new_empty(shape).bernoulli_(keep)
        return x * mask / keep


class _ConvModule(nn.Module):
    """mmcv ConvModule as SegFormerHead instantiates it (segformer_head.py:50-55):
    1x1 conv without bias (a norm follows) -> BatchNorm2d -> ReLU(inplace)."""

    def __init__(self, in_channels, out_channels, kernel_size, norm_cfg=None, **kw):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, bias=norm_cfg is None)
        self.bn = nn.BatchNorm2d(out_channels) if norm_cfg is not None else None
        self.activate = nn.ReLU(inplace=True)

    def forward(self, x):
        x = self.conv(x)
        if self.bn is not None:
            x = self.bn(x)
        return self.activate(x)


def _install_stubs():
    if "timm.models.layers" not in sys.modules:
        timm = types.ModuleType("timm")
        models = types.ModuleType("timm.models")
        layers = types.ModuleType("timm.models.layers")
        layers.DropPath = _DropPath
        layers.to_2tuple = lambda x: tuple(x) if isinstance(x, (tuple, list)) else (x, x)
        layers.trunc_normal_ = torch.nn.init.trunc_normal_
        timm.models = models
        models.layers = layers
        sys.modules.setdefault("timm", timm)
        sys.modules.setdefault("timm.models", models)
        sys.modules.setdefault("timm.models.layers", layers)
    if "mmcv.cnn" not in sys.modules:
        mmcv = types.ModuleType("mmcv")
        cnn = types.ModuleType("mmcv.cnn")
        cnn.ConvModule = _ConvModule
        cnn.DepthwiseSeparableConvModule = type("DepthwiseSeparableConvModule", (nn.Module,), {})
        mmcv.cnn = cnn
        sys.modules.setdefault("mmcv", mmcv)
        sys.modules.setdefault("mmcv.cnn", cnn)


def _load(modname, relpath, package=None):
    full = f"{_PREFIX}.{modname}"
    if full in sys.modules:
        return sys.modules[full]
    spec = importlib.util.spec_from_file_location(full, os.path.join(REF_ROOT, relpath))
    mod = importlib.util.module_from_spec(spec)
    if package:
        mod.__package__ = package
    sys.modules[full] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference():
    """Returns a namespace with the reference modules:
    .mix_transformer .segformer_head .model_fusion .Entropy .pytorch_ssim .lap_loss
    (core/loss.py is not loaded: its composites hard-code .cuda(); the oracle restates them.)"""
    if not available():
        raise RuntimeError(f"reference not mounted at {REF_ROOT}")
    _install_stubs()
    pkg = f"{_PREFIX}.core"
    if pkg not in sys.modules:
        root = types.ModuleType(_PREFIX)
        root.__path__ = []
        sys.modules[_PREFIX] = root
        core = types.ModuleType(pkg)
        core.__path__ = [os.path.join(REF_ROOT, "core")]
        sys.modules[pkg] = core
    ns = types.SimpleNamespace()
    ns.mix_transformer = _load("core.mix_transformer", "core/mix_transformer.py", pkg)
    ns.segformer_head = _load("core.segformer_head", "core/segformer_head.py", pkg)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        ns.model_fusion = _load("core.model_fusion", "core/model_fusion.py", pkg)
    ns.Entropy = _load("core.Entropy", "core/Entropy.py", pkg)
    ns.pytorch_ssim = _load("pytorch_ssim", "pytorch_ssim/__init__.py")
    ns.lap_loss = _load("lap_loss", "lap_loss.py")
    return ns


class cuda_is_identity:
    """Context manager: makes Tensor.cuda()/Module.cuda() no-ops so the reference's hard-coded
    `.cuda()` call sites (core/loss.py:345-646, core/model_fusion.py:81,98-100) run on the CPU
    unmodified.  Used only while generating golden fixtures."""

    def __enter__(self):
        self._t, self._m = torch.Tensor.cuda, nn.Module.cuda
        torch.Tensor.cuda = lambda self, *a, **k: self
        nn.Module.cuda = lambda self, *a, **k: self
        return self

    def __exit__(self, *exc):
        torch.Tensor.cuda, nn.Module.cuda = self._t, self._m
        return False


def load_reference_losses():
    """core/loss.py (imports lap_loss / pytorch_ssim by top-level name, so REF_ROOT joins sys.path
    for the duration of the import).  Instantiate its classes under `cuda_is_identity()`."""
    ns = load_reference()
    full = f"{_PREFIX}.core.loss"
    if full in sys.modules:
        return sys.modules[full]
    saved = {k: sys.modules.get(k) for k in ("lap_loss", "pytorch_ssim")}
    sys.modules["lap_loss"], sys.modules["pytorch_ssim"] = ns.lap_loss, ns.pytorch_ssim
    try:
        mod = _load("core.loss", "core/loss.py", f"{_PREFIX}.core")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def reference_inference_pipeline(ns, seg, fus, ir, vis_rgb, mask):
    """The unit of work of the headline metric executed by the UNMODIFIED reference modules on the CPU, statement for
    statement as train.py:356-366 + test_fusion.py:100-111 + test_segmentation.py:169-175 do (the hard-coded .cuda()
    calls neutralised by cuda_is_identity).  `seg` / `fus`: reference Network3 / Fusion_Network3_ac instances."""
    import torch.nn.functional as F
    with torch.no_grad(), cuda_is_identity():
        out0, out1 = seg.denoise_net.encoder.forward_fusion(mask)
        vis_ycc = ns.model_fusion.RGB2YCrCb(vis_rgb)
        fused = fus(ir, vis_ycc, out0, out1)
        ycc = vis_ycc.clone()
        ycc[:, 0:1] = fused
        rgb = ns.model_fusion.YCrCb2RGB(ycc).clamp(0, 1)
        logits = seg(rgb.clone())[2]
        labels = F.interpolate(logits, size=ir.shape[2:], mode="bilinear", align_corners=False).argmax(1)
    return dict(fused=fused, rgb=rgb, logits=logits, labels=labels)


def build_reference_models(ns, backbone, seed=0):
    """Reference Network3(backbone) + Fusion_Network3_ac with the deterministic synthetic weights of segmif_b200.synth."""
    import contextlib
    import io
    from segmif_b200 import synth
    with contextlib.redirect_stdout(io.StringIO()):           # DRDB.__init__ prints (model_fusion.py:131)
        seg = ns.model_fusion.Network3(backbone, 9, 256, None)
        fus = ns.model_fusion.Fusion_Network3_ac()
    synth.load_synthetic(seg, seed)
    synth.load_synthetic(fus, seed)
    return seg.eval(), fus.eval()


def load_reference_datapath():
    """The reference's data path (datasets/imutils.py + datasets/voc_fusion3.py), unmodified, as a callable
    `transforms(image, image_vis, image_mask, label, crop_size, rescale_range) -> (image, image_vis, image_mask, label)` that runs
    `VOC12SegDataset.__transforms` (voc_fusion3.py:169-209) with the global `random` / `np.random` generators.  mmcv is not in
    this image: its two colour helpers are `cv2.cvtColor(img, cv2.COLOR_BGR2HSV / HSV2BGR)` (mmcv/image/colorspace.py) and are
    attached to the stand-in module; imageio (PNG reading only) is stubbed.  Needs Pillow and OpenCV."""
    import cv2
    _install_stubs()
    mm = sys.modules["mmcv"]
    mm.bgr2hsv = lambda img: cv2.cvtColor(img, cv2.COLOR_BGR2HSV)
    mm.hsv2bgr = lambda img: cv2.cvtColor(img, cv2.COLOR_HSV2BGR)
    sys.modules.setdefault("imageio", types.ModuleType("imageio"))
    try:
        import torchvision  # noqa: F401  (imutils.py:6)
    except ImportError:
        sys.modules.setdefault("torchvision", types.ModuleType("torchvision"))
    pkg_name = f"{_PREFIX}.datasets"
    if pkg_name not in sys.modules:
        pkg = types.ModuleType(pkg_name)
        pkg.__path__ = [os.path.join(REF_ROOT, "datasets")]
        sys.modules[pkg_name] = pkg
    imutils = _load("datasets.imutils", os.path.join("datasets", "imutils.py"), package=pkg_name)
    sys.modules[pkg_name].imutils = imutils
    voc = _load("datasets.voc_fusion3", os.path.join("datasets", "voc_fusion3.py"), package=pkg_name)

    def transforms(image, image_vis, image_mask, label, crop_size=512, rescale_range=(0.5, 2.0)):
        ds = object.__new__(voc.VOC12SegDataset)
        ds.aug, ds.ignore_index, ds.resize_range, ds.rescale_range, ds.crop_size, ds.img_fliplr = True, 255, [512, 640], rescale_range, crop_size, True
        ds.color_jittor = imutils.PhotoMetricDistortion()
        return ds._VOC12SegDataset__transforms(image, image_vis, image_mask, label)

    return transforms
