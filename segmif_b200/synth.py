"""Deterministic synthetic weights and image pairs (no datasets or checkpoints exist offline).

Weights are a pure function of (parameter name, shape, seed) through numpy's legacy
RandomState (bit-stable across numpy versions), so the reference modules, the oracle
and the CUDA modules can be given *identical* parameters from nothing but their
state-dict key lists -- which also checks that the key lists agree.

Inputs follow SURVEY.md 8(d): ir ~ U[0,1) [B,1,H,W]; vis ~ U[0,1) [B,3,H,W]; mask = one
U[0,1) plane replicated to 3 channels (datasets/voc_fusion3.py:46-48 of the reference);
labels ~ randint(0, classes) with 5 % set to the ignore index 255.
"""
import zlib

import numpy as np
import torch


def _rs(name, seed):
    return np.random.RandomState((zlib.crc32(name.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)


def synth_tensor(name, shape, seed=0):
    shape = tuple(shape)
    rs = _rs(name, seed)
    leaf = name.rsplit(".", 1)[-1]
    if leaf == "num_batches_tracked":
        return torch.zeros(shape, dtype=torch.int64)
    if leaf == "running_var":
        return torch.from_numpy(rs.uniform(0.5, 1.5, size=shape).astype(np.float32))
    if leaf == "running_mean":
        return torch.from_numpy((0.1 * rs.standard_normal(shape)).astype(np.float32))
    if name.endswith("relu.weight") and int(np.prod(shape)) == 1:      # the shared scalar PReLU
        return torch.full(shape, 0.25, dtype=torch.float32)
    if len(shape) == 1:
        is_norm_w = leaf == "weight"                                    # 1-D weights are LN / BN gains
        v = rs.standard_normal(shape)
        return torch.from_numpy(((1.0 + 0.1 * v) if is_norm_w else 0.05 * v).astype(np.float32))
    fan_in = int(np.prod(shape[1:]))
    if ".cross." in name:                                               # FFM linears: reference init scale
        std = 0.02                                                       # (trunc_normal_(std=.02), model_fusion.py:440)
    else:
        std = 0.8 / np.sqrt(fan_in)
    return torch.from_numpy((std * rs.standard_normal(shape)).astype(np.float32))


def synth_state_dict(shapes, seed=0):
    """shapes: mapping name -> shape (e.g. {k: v.shape for k, v in module.state_dict().items()})."""
    return {k: synth_tensor(k, s, seed) for k, s in shapes.items()}


def load_synthetic(module, seed=0):
    """Fill `module`'s parameters and buffers in place with the synthetic values."""
    sd = module.state_dict()
    new = synth_state_dict({k: v.shape for k, v in sd.items()}, seed)
    module.load_state_dict({k: new[k].to(sd[k].dtype) for k in sd})
    return module


def synth_inputs(batch, height, width, seed=0, num_classes=9, ignore_index=255):
    rs = np.random.RandomState(1234 + seed)
    f32 = lambda a: torch.from_numpy(a.astype(np.float32))
    ir = f32(rs.uniform(0, 1, size=(batch, 1, height, width)))
    vis = f32(rs.uniform(0, 1, size=(batch, 3, height, width)))
    mask1 = rs.uniform(0, 1, size=(batch, 1, height, width))
    mask = f32(np.repeat(mask1, 3, axis=1))
    labels = rs.randint(0, num_classes, size=(batch, height, width)).astype(np.int64)
    labels[rs.uniform(size=labels.shape) < 0.05] = ignore_index
    return dict(ir=ir, vis=vis, mask=mask, labels=torch.from_numpy(labels))


def analytic_images(batch=2, height=64, width=96, phases=(0.0, 0.9, 2.1)):
    """RNG-free test images of SURVEY.md Appendix C: 0.5+0.5*sin(0.13x + 0.07y(b+1) + phi), fp64->fp32."""
    y, x = np.meshgrid(np.arange(height, dtype=np.float64), np.arange(width, dtype=np.float64), indexing="ij")
    out = []
    for phi in phases:
        img = np.stack([0.5 + 0.5 * np.sin(0.13 * x + 0.07 * y * (b + 1) + phi) for b in range(batch)])
        out.append(torch.from_numpy(img[:, None].astype(np.float32)))
    return out


def synth_decoded_sample(seed, h, w, n_class=9, ignore_frac=0.03, mask_channels=1):
    """One decoded training sample as the reference's dataset class holds it after imread (datasets/voc_fusion3.py:36-55):
    uint8 infrared [h, w], visible [h, w, 3], mask [h, w], label [h, w] (numpy).  The label is made of rectangles, so crop
    windows dominated by one class (the retry branch of random_crop2) and windows of ignore_index both occur; half of the
    visible image is quantised / grey so every HSV sector and the s == 0 path are hit."""
    rs = np.random.RandomState(seed)
    ir = rs.randint(0, 256, size=(h, w)).astype(np.uint8)
    vis = rs.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
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
