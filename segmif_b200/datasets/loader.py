"""Batch iterator over a device dataset: stands where train.py builds its `DataLoader` (train.py:137-142, :289-294) and yields
the tuple that loader yields -- `(names, inputs_ir, inputs_vis, inputs_mask, labels)` (train.py:212, :346) -- with the tensors
already on the device.  PNG decoding (sequential inflate, the one step left on the host) runs in a thread pool one batch ahead;
the transforms of a whole batch are one call of `VOC12SegDataset.batch` (csrc/datapath.cu).

    train_loader = DeviceLoader(train_dataset, batch_size=cfg.train.samples_per_gpu // 2, drop_last=True)   # instead of DataLoader(...)

Like train.py's loaders (their `shuffle=True` is commented out) the default order is sequential; `shuffle=True` draws a
permutation per epoch from `generator` (torch.randperm, what DataLoader's RandomSampler does)."""
from concurrent.futures import ThreadPoolExecutor

import torch


class DeviceLoader:
    def __init__(self, dataset, batch_size, shuffle=False, drop_last=True, decode_threads=4, generator=None, label_int64=False):
        if batch_size <= 0:
            raise ValueError("segmif_b200.datasets: batch_size must be positive")
        self.dataset, self.batch_size, self.shuffle, self.drop_last = dataset, int(batch_size), shuffle, drop_last
        self.generator, self.label_int64 = generator, label_int64
        self.decode_threads = max(1, int(decode_threads))

    def __len__(self):
        n = len(self.dataset)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def _batches(self):
        n = len(self.dataset)
        order = torch.randperm(n, generator=self.generator).tolist() if self.shuffle else list(range(n))
        for i in range(0, n, self.batch_size):
            idx = order[i:i + self.batch_size]
            if len(idx) < self.batch_size and self.drop_last:
                return
            yield idx

    def __iter__(self):
        ds = self.dataset
        with ThreadPoolExecutor(self.decode_threads) as pool:
            ahead = None
            for idx in self._batches():
                futs = [pool.submit(ds.decode_host, i) for i in idx]          # next batch decodes while this one is transformed / consumed
                if ahead is not None:
                    yield ds.transform_decoded([f.result() for f in ahead], label_int64=self.label_int64)
                ahead = futs
            if ahead is not None:
                yield ds.transform_decoded([f.result() for f in ahead], label_int64=self.label_int64)
