"""Throughput of the device data path on its dataset-default workload: 480 x 640 decoded samples resident
in HBM -> random scale (0.5..2), flip, PhotoMetricDistortion, 512 x 512 crop, / 255, CHW -- one call per batch of 32.  Reports
device time (CUDA events, includes the label-stage read-back the host waits on), wall time (host draws + launches included), and
the reference's own CPU `__transforms` (oracle/_ref, one core) on a few samples of the same workload.
Algorithmic bytes per sample: inputs 6 B/px x 480 x 640 + outputs (9 fp32 channels + fp32 label) x 512 x 512 = 12.3 MB."""
import json
import os
import random
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def measure(batch=32, crop=512, h=480, w=640, iters=20, warm=3, cpu_samples=6, shared_rng=False):
    from segmif_b200 import _lib, synth
    from segmif_b200.datasets import DeviceTransforms, Rng
    dev = torch.device("cuda", 0)
    host_samples = [synth.synth_decoded_sample(1000 + k, h, w) for k in range(batch)]
    samples = [tuple(torch.from_numpy(a).to(dev) for a in s) for s in host_samples]
    tf = DeviceTransforms(crop_size=crop)
    rngs = Rng.seeded(1) if shared_rng else [Rng.seeded(k) for k in range(batch)]
    for _ in range(warm):
        tf(samples, rngs, label_int64=True)
    torch.cuda.synchronize()
    l0 = _lib.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(iters):
        out = tf(samples, rngs, label_int64=True)
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / iters
    dev_ms = e0.elapsed_time(e1) / iters
    bytes_per_sample = 6 * h * w + (9 * 4 + 4 + 8) * crop * crop
    res = {"workload": f"{h}x{w} uint8 samples in HBM -> scale 0.5..2 + flip + distortion + crop {crop} + /255 + CHW (+ int64 labels), batch {batch}",
           "rng": "one shared generator (per-sample read-back)" if shared_rng else "one generator per sample (one read-back per batch)",
           "samples_per_sec_wall": batch / wall, "ms_per_batch_wall": wall * 1e3, "ms_per_batch_device_events": dev_ms,
           "entry_point_calls_per_batch": (_lib.launch_count - l0) // iters,
           "algorithmic_bytes_per_sample": bytes_per_sample, "achieved_GBps_wall": batch * bytes_per_sample / wall / 1e9}
    if cpu_samples:
        from oracle import data_oracle as do                  # cpu_baseline leg only: the reference / its restatement as the thing timed
        try:
            from oracle import ref_shim
            ref = ref_shim.load_reference_datapath()
            kind = "reference"
            run = lambda s: ref(*do.dataset_views(*s[:3]), s[3], crop_size=crop)
        except Exception as e:  # noqa: BLE001
            kind = f"port ({type(e).__name__}: {e})"
            run = lambda s: do.transforms(*do.dataset_views(*s[:3]), s[3], do.Rng(), crop_size=crop)
        random.seed(0)
        np.random.seed(0)
        run(host_samples[0])
        t0 = time.perf_counter()
        for k in range(cpu_samples):
            run(host_samples[k % batch])
        cpu = (time.perf_counter() - t0) / cpu_samples
        res["cpu_baseline"] = {"value": 1.0 / cpu, "unit": "samples/s", "cores": 1, "kind": kind,
                               "sample": f"{cpu_samples} samples of the same workload through VOC12SegDataset.__transforms (PNG decode excluded on both sides)",
                               "note": "train.py:116 runs 4 DataLoader workers: x4 at best"}
    return res


if __name__ == "__main__":
    out = {"own_rng": measure(), "shared_rng": measure(shared_rng=True, cpu_samples=0), "batch8": measure(batch=8, cpu_samples=0)}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "datapath_bench.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))
