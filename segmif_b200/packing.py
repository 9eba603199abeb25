"""Host-side weight packing (done once per parameter version, never inside the per-step hot path):
fp32 master parameters -> the bf16 / tap-major layouts the kernels consume."""
import torch


_EPOCH = [0]


def invalidate_all():
    """Declares every cached pack stale (for code that writes parameters through raw pointers without going through
    ddp.FlatParams, whose optimizer bumps a per-buffer epoch instead -- see _epoch)."""
    _EPOCH[0] += 1


def _epoch(param):
    """The fused optimizer updates parameters through raw pointers, which does not bump torch's per-tensor version
    counters; ddp.FlatParams therefore hangs a shared mutable epoch on each parameter it owns and the optimizer bumps
    it.  Parameters outside any flat buffer (a frozen network next to a trained one) are not invalidated by it."""
    e = getattr(param, "_segmif_epoch", None)
    return (_EPOCH[0], e[0] if e is not None else -1)


class PackCache:
    """Caches packed copies of parameters; an entry is rebuilt when the parameter is modified in place
    (optimizer step, load_state_dict), moved, or replaced."""

    def __init__(self):
        self._c = {}

    def __deepcopy__(self, memo):
        return PackCache()

    def __getstate__(self):
        return {}

    def __setstate__(self, state):
        self._c = {}

    def get(self, param, fn, tag=""):
        key = (id(param), tag)
        stamp = (param.data_ptr(), param._version, param.device, _epoch(param))
        hit = self._c.get(key)
        if hit is not None and hit[0] == stamp:
            return hit[1]
        with torch.no_grad():
            packed = fn(param)
        self._c[key] = (stamp, packed)
        return packed

    def get_multi(self, params, fn, tag):
        """One packed object derived from several parameters."""
        key = ("multi", tag)
        stamp = tuple((p.data_ptr(), p._version, p.device, _epoch(p)) for p in params)
        hit = self._c.get(key)
        if hit is not None and hit[0] == stamp:
            return hit[1]
        with torch.no_grad():
            packed = fn(*params)
        self._c[key] = (stamp, packed)
        return packed

    def linear(self, weight):
        """nn.Linear weight [N, K] -> bf16 [N, 1, K] (K-major rows, the B operand layout)."""
        return self.get(weight, lambda w: w.detach().to(torch.bfloat16).reshape(w.shape[0], 1, w.shape[1]).contiguous(), "lin")

    def conv(self, weight):
        """nn.Conv2d weight [Cout, Cin, KH, KW] -> bf16 [Cout, KH*KW, Cin] (tap-major, channels innermost)."""
        return self.get(weight, lambda w: w.detach().permute(0, 2, 3, 1).reshape(w.shape[0], -1, w.shape[1])
                        .to(torch.bfloat16).contiguous(), "conv")

    def taps_f32(self, weight):
        """[Cout, Cin, 3, 3] with Cin == 1 or Cout == 1 -> fp32 [9, C] for the edge-layer stencils."""
        def f(w):
            w = w.detach().float()
            if w.shape[1] == 1:                                   # 1 -> Cout
                return w.reshape(w.shape[0], 9).t().contiguous()
            return w.reshape(w.shape[1], 9).t().contiguous()      # Cin -> 1
        return self.get(weight, f, "taps")
