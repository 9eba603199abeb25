"""GPU parity tests proper: every check in tests/gpu_checks.py calls the CUDA path through the C ABI
(ctypes -> libsegmif_b200.so) and compares it with the CPU oracle / the reference-generated fixtures."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _checks():
    import gpu_checks
    import gpu_checks_train  # noqa: F401  (registers the training-side checks in the same list)
    import gpu_checks_strict  # noqa: F401  (strict-precision mode)
    import gpu_checks_data  # noqa: F401  (device data path)
    return gpu_checks.CHECKS


def _names():
    try:
        return [c.__name__ for c in _checks()]
    except Exception:  # noqa: BLE001 -- collection must not fail on machines without the built library
        return []


@pytest.mark.parametrize("name", _names())
def test_parity(name):
    fn = {c.__name__: c for c in _checks()}[name]
    res = fn()
    res = res if isinstance(res, list) else [res]
    torch.cuda.synchronize()
    bad = [r for r in res if not r["ok"]]
    assert not bad, "\n".join(f"{r['name']}: err {r['err']:.3e} > tol {r['tol']:.1e} {r.get('note', '')}" for r in bad)


def test_native_library_is_what_ran():
    """The extension must be the thing that ran: it is loaded from inside the repo and served kernel launches."""
    from segmif_b200 import _lib
    import gpu_checks
    gpu_checks.layernorm()
    assert _lib.LIB_PATH.endswith("segmif_b200/libsegmif_b200.so")
    assert _lib.launch_count > 0
    with open("/proc/self/maps") as f:
        assert "libsegmif_b200.so" in f.read()
