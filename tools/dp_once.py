import sys, torch
sys.path.insert(0, '/root/repo')
from segmif_b200 import synth
from segmif_b200.datasets import DeviceTransforms, Rng
dev = torch.device("cuda", 0)
hs = [synth.synth_decoded_sample(1000 + k, 480, 640) for k in range(32)]
samples = [tuple(torch.from_numpy(a).to(dev) for a in s) for s in hs]
tf = DeviceTransforms(crop_size=512)
rngs = [Rng.seeded(k) for k in range(32)]
for _ in range(2):
    tf(samples, rngs, label_int64=True)
torch.cuda.synchronize()
import time
# host-only cost of the draws
import ctypes
from segmif_b200 import _lib
host = (_lib.DpSample * 32)()
t = time.perf_counter()
for k in range(32):
    tf._draw(host[k], samples[k], rngs[k])
print("draw host ms per batch", (time.perf_counter() - t) * 1e3)
