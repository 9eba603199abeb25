"""Host half of the device data path (segmif_b200/datasets/imutils.py) on the CPU: the random draws must be the reference's,
in the reference's order, and the crop-window decision must leave `random` where the reference leaves it.  The pixel work is
replaced here by numpy (oracle pieces) fed from the descriptor the host code filled in, so the test covers exactly the logic
that does NOT run on the GPU."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import data_oracle as do
from segmif_b200 import _lib
from segmif_b200.datasets import imutils
from segmif_b200.datasets.imutils import DeviceTransforms, Rng


def _canvas_from_descriptor(s, label):
    lab = do.resize_nearest(label, s.nw, s.nh) if s.resized else label
    if s.flip:
        lab = np.fliplr(lab)
    canvas = np.full((s.PH, s.PW), 255, np.uint8)
    canvas[s.pad_h:s.pad_h + s.nh, s.pad_w:s.pad_w + s.nw] = lab
    return canvas


def _stats(canvas, s, crop):
    out = np.zeros((10, 3), np.int32)
    for i in range(10):
        win = canvas[s.cand_hs[i]:s.cand_hs[i] + crop, s.cand_ws[i]:s.cand_ws[i] + crop]
        idx, cnt = np.unique(win, return_counts=True)
        cnt = cnt[idx != 255]
        out[i] = (len(cnt), cnt.max() if len(cnt) else 0, cnt.sum())
    return out


@pytest.mark.parametrize("seed", range(12))
@pytest.mark.parametrize("rescale", [(0.5, 2.0), None])
def test_draws_and_window_decision_follow_the_reference(seed, rescale):
    h, w, crop = 60, 80, 48
    ir, vis, mask, label = do.synth_sample(seed, h, w)
    if seed % 3 == 0:
        label = np.full_like(label, 4)                       # single class: never accepted, all ten candidates consumed
    r_ref, r_dev = do.Rng.seeded(seed), Rng.seeded(seed)
    ref = do.transforms(*do.dataset_views(ir, vis, mask), label, r_ref, rescale_range=rescale, crop_size=crop)
    tf = DeviceTransforms(crop_size=crop, rescale_range=rescale)
    s = _lib.DpSample()
    sample = tuple(torch.from_numpy(a) for a in (ir, vis, mask, label))
    state = tf._draw(s, sample, r_dev)
    canvas = _canvas_from_descriptor(s, label)
    tf._decide(s, _stats(canvas, s, crop), state, r_dev)
    # same window (the label crop is the reference's), same generator states afterwards
    assert np.array_equal(canvas[s.hs:s.hs + crop, s.ws:s.ws + crop].astype(np.float32), ref[3])
    assert r_ref.py.random() == r_dev.py.random()
    assert r_ref.np.randint(1 << 30) == r_dev.np.randint(1 << 30)
    # region bookkeeping: the region covers exactly the image pixels inside the window
    if s.resized:
        assert 0 <= s.roi_y0 < s.roi_y1 <= s.nh and 0 <= s.roi_x0 < s.roi_x1 <= s.nw
        xmin, xcnt, _ = do.bilinear_coeffs(h, s.nh)
        assert s.src_y0 == xmin[s.roi_y0] and s.src_y1 == xmin[s.roi_y1 - 1] + xcnt[s.roi_y1 - 1]
        assert s.ks_y == do.bilinear_coeffs(h, s.nh)[2].shape[1] and s.ks_x == do.bilinear_coeffs(w, s.nw)[2].shape[1]


def test_distortion_program_matches_oracle_dtype_state_machine():
    """The op list (kind, on-uint8 flag) the host hands to the kernel describes what the oracle's PhotoMetricDistortion does."""
    for seed in range(40):
        for start_u8 in (False, True):
            ops = imutils.PhotoMetricDistortion().draw(Rng.seeded(seed), is_uint8=start_u8)
            img = np.random.RandomState(seed).randint(0, 256, size=(8, 40, 3)).astype(np.uint8 if start_u8 else np.float32)
            want = do.photometric_distortion(img.copy(), do.Rng.seeded(seed))
            u8 = start_u8
            for kind, alpha, beta, delta, flag in ops:
                assert flag == u8
                if kind == imutils.OP_CONVERT:
                    u8 = True
            assert (want.dtype == np.uint8) == u8
            assert len(ops) <= _lib.DP_MAX_OPS


def test_product_and_oracle_synthesise_the_same_samples():
    from segmif_b200 import synth
    for seed, h, w in ((0, 48, 64), (7, 60, 80), (1003, 480, 640)):
        for mc in (1, 3):
            for a, b in zip(synth.synth_decoded_sample(seed, h, w, mask_channels=mc), do.synth_sample(seed, h, w, mask_channels=mc)):
                assert a.dtype == b.dtype and np.array_equal(a, b)


def test_descriptor_layout_matches_header():
    """ctypes mirror vs include/segmif_b200.h (compiled with the host compiler)."""
    import os
    import subprocess
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = ('#include <stdio.h>\n#include <stddef.h>\n#include "%s/include/segmif_b200.h"\n'
           'int main(){printf("%%zu %%zu %%zu %%zu %%zu", sizeof(segmif_dp_sample), offsetof(segmif_dp_sample, cand_hs), '
           'offsetof(segmif_dp_sample, op_alpha), offsetof(segmif_dp_sample, ks_x), offsetof(segmif_dp_sample, tab_off));return 0;}\n') % root
    with tempfile.TemporaryDirectory() as d:
        with open(os.path.join(d, "t.c"), "w") as f:
            f.write(src)
        subprocess.check_call(["gcc", os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
        got = [int(x) for x in subprocess.check_output([os.path.join(d, "t")]).split()]
    S = _lib.DpSample
    assert got == [ctypes.sizeof(S), S.cand_hs.offset, S.op_alpha.offset, S.ks_x.offset, S.tab_off.offset]


def test_device_loader_batching_order():
    """DeviceLoader's index batching (sequential like train.py's loaders, drop_last, seeded shuffle) without touching a GPU."""
    from segmif_b200.datasets import DeviceLoader

    class Fake:
        def __len__(self):
            return 7
    ld = DeviceLoader(Fake(), batch_size=3, drop_last=True)
    assert len(ld) == 2 and list(ld._batches()) == [[0, 1, 2], [3, 4, 5]]
    ld = DeviceLoader(Fake(), batch_size=3, drop_last=False)
    assert len(ld) == 3 and list(ld._batches())[-1] == [6]
    g = torch.Generator().manual_seed(3)
    a = list(DeviceLoader(Fake(), batch_size=2, shuffle=True, generator=g)._batches())
    g = torch.Generator().manual_seed(3)
    assert a == list(DeviceLoader(Fake(), batch_size=2, shuffle=True, generator=g)._batches())
    assert sorted(i for b in a for i in b) != [] and len({i for b in a for i in b}) == 6
