"""CPU-side checks: the C-ABI library loads and exports every symbol include/segmif_b200.h declares (no compute
calls without a GPU), the Python mirror has the reference's state_dict surface, and the product path refuses to
run without CUDA instead of falling back."""
import os
import re

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT


def test_library_exports_every_declared_symbol():
    from segmif_b200 import _lib
    header = open(os.path.join(ROOT, "include", "segmif_b200.h")).read()
    declared = set(re.findall(r"\b(segmif_[a-z0-9_]+)\s*\(", header))
    declared -= {"segmif_conv_params"}
    lib = _lib.load()
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.segmif_abi_version() == 1


def test_init_reports_missing_device_loudly():
    from segmif_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    rc = _lib.load().segmif_init(0)
    assert rc != 0 and "CUDA" in _lib.last_error()


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from segmif_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU path|CUDA"):
        ops.layernorm(torch.zeros(4, 64), torch.ones(64), torch.zeros(64), 1e-5)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "segmif_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f"{f} imports the oracle"


def _keys(mod):
    return [f"{k}|{','.join(map(str, v.shape))}" for k, v in mod.state_dict().items()]


def test_state_dict_surface_matches_reference():
    from segmif_b200.core.model_fusion import Fusion_Network3_ac, Network3
    g = np.load(os.path.join(GOLDEN, "state_dict_keys.npz"))
    assert _keys(Fusion_Network3_ac()) == [str(s) for s in g["fusion_keys"]]
    for bb in ("mit_b0", "mit_b1", "mit_b2"):
        assert _keys(Network3(bb, 9, 256, None)) == [str(s) for s in g[f"network3_{bb}_keys"]]


def test_core_package_surface():
    import segmif_b200.core as core
    for name in ("Network", "WeTr", "SegFormerHead", "mit_b0", "mit_b5", "MixVisionTransformer", "OverlapPatchEmbed",
                 "Attention", "Mlp", "DWConv", "Block", "Fusionloss3", "Fusionloss_grad3", "Fusionloss_grad2",
                 "Total_fusion_loss", "Total_fusion_loss2", "Fusionloss", "Fusionloss_add", "Fusionloss2",
                 "Fusionloss4", "RGB2YCrCb"):
        assert hasattr(core, name), name
    from segmif_b200.core.model_fusion import Mean, Network3, Fusion_Network3_ac  # noqa: F401  (train.py:18)
    n = Network3("mit_b1", 9, 256, None)
    groups = n.denoise_net.get_param_groups()
    assert len(groups) == 3 and groups[2][-1] is n.denoise_net.classifier.weight
    assert all("norm" in k for k, p in n.denoise_net.encoder.named_parameters() if any(p is q for q in groups[1]))


def test_ffm_operand_packing_matches_oracle_algebra():
    """Host-side folding of conv3 into channel_proj3 and the y/u halves: pure tensor algebra, checked on CPU."""
    from oracle import segmif_oracle as O
    from segmif_b200 import synth
    from segmif_b200.core.model_fusion import Fusion_Network3_ac
    fus = synth.load_synthetic(Fusion_Network3_ac(), 0)
    pk = fus.ffm.cross.packs(fus.conv4)
    assert pk["C3"] == 128
    sd = fus.state_dict()
    x = torch.randn(5, 128)
    seg = torch.nn.functional.linear(x, sd["conv4.weight"].reshape(64, 128), sd["conv4.bias"])
    full = torch.nn.functional.linear(seg, sd["ffm.cross.channel_proj3.weight"], sd["ffm.cross.channel_proj3.bias"])
    w3y = pk["w_apply"][:64 * 128].float().reshape(64, 128)
    w3u = pk["w_gram"][2 * 4096:].float().reshape(64, 128)
    y3 = x @ w3y.t() + pk["b_apply"][:64]
    u3 = x @ w3u.t() + pk["b_gram"][128:]
    assert torch.allclose(y3, full[:, :64], atol=2e-2)      # bf16-rounded folded weights
    assert torch.allclose(u3, full[:, 64:], atol=2e-2)


def test_optimizer_schedule_matches_reference_formula():
    from segmif_b200.utils.optimizer import PolyWarmupAdamW
    p = torch.nn.Parameter(torch.zeros(3))
    opt = PolyWarmupAdamW([{"params": [p], "lr": 1e-3}], lr=1e-3, weight_decay=0.01, betas=(0.9, 0.999), warmup_iter=10,
                          max_iter=100, warmup_ratio=1e-6, power=1.0)
    lrs = []
    for _ in range(30):
        p.grad = torch.ones(3)
        opt.step()
        lrs.append(opt.param_groups[0]["lr"])
    assert abs(lrs[0] - 1e-3 * 1e-6) < 1e-12
    assert abs(lrs[5] - 1e-3 * (1 - (1 - 5 / 10) * (1 - 1e-6))) < 1e-12
    assert abs(lrs[20] - 1e-3 * (1 - 20 / 100)) < 1e-12


def test_host_only_sizing_entry_points():
    """Workspace / chunk sizing functions are pure host code: callable without a GPU, and their answers must be
    consistent with what the kernels index (partials[chunk][Cout][taps][Cin], per-block depthwise partials)."""
    from segmif_b200 import _lib
    lib = _lib.load()
    # wgrad: at least one chunk, never more chunks than 128-pixel tiles (3x3) / 64-pixel steps (linear)
    for (B, H, W, Cin, Cout, taps, dil) in ((4, 480, 640, 64, 32, 9, 2), (1, 17, 23, 8, 64, 9, 1), (4, 480, 640, 224, 32, 9, 2),
                                            (1, 8, 16, 192, 32, 9, 2)):
        n = lib.segmif_wgrad_chunks(B, H, W, B * H * W, Cin, Cout, taps, dil)
        tiles = B * ((H + 7) // 8) * ((W + 15) // 16)
        assert 1 <= n <= tiles
        assert lib.segmif_wgrad_workspace_bytes(n, Cout, taps, Cin) == n * Cout * taps * Cin * 4
    for (P, K, N) in ((76800, 64, 64), (50, 64, 64), (1200, 512, 2048), (4800, 1280, 320)):
        n = lib.segmif_wgrad_chunks(1, (P + 15) // 16, 16, P, K, N, 1, 1)
        assert 1 <= n <= (P + 63) // 64
    # depthwise backward: one partial row of 10*C floats per block
    for (B, H, W, C) in ((4, 120, 160, 256), (1, 7, 9, 2048), (1, 9, 11, 40), (8, 15, 20, 2048)):
        ws = lib.segmif_dwconv3x3_gelu_bwd_workspace(B, H, W, C)
        assert ws > 0 and ws % (10 * C) == 0
    assert lib.segmif_dwconv3x3_gelu_bwd_workspace(1, 8, 8, 12) == 0          # C % 8 != 0: rejected
    assert lib.segmif_loss_workspace_bytes(2, 64, 96) >= 256 + 4096 * 3 * 4


def test_argument_validation_fails_before_any_launch():
    """Bad arguments are refused with a message by the entry points themselves (checked before the first CUDA call)."""
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import ctypes
    from segmif_b200 import _lib
    lib = _lib.load()
    buf = (ctypes.c_float * 64)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert lib.segmif_confusion_matrix(p, p, 10, 100, p, None) != 0 and "num_classes" in _lib.last_error()
    assert lib.segmif_colsum(p, 12, 0, 10, 12, p, None) != 0 and "multiples of 8" in _lib.last_error()
    assert lib.segmif_sr_attention_tc_fwd(p, 64, p, p, 128, p, 64, 1, 1, 10, 400, 64, 0.125, None, None) != 0 \
        and "Nk <= 320" in _lib.last_error()


def test_configs_load_through_the_dropin_omegaconf():
    """configs/voc*.yaml (the reference's schema) load through `OmegaConf.load` exactly as train.py:25 / test_fusion.py:44
    do, with or without the real omegaconf: attribute access, list values, and the exponent-only floats ('1e-6') that
    OmegaConf reads as numbers."""
    import glob
    import sys
    import segmif_b200.dropin as d
    had = sys.modules.get("omegaconf")
    try:
        d.install()
        from omegaconf import OmegaConf
        paths = sorted(glob.glob(os.path.join(ROOT, "configs", "voc*.yaml")))
        assert paths
        for path in paths:
            cfg = OmegaConf.load(path)
            assert cfg.exp.backbone.startswith("mit_b")
            assert isinstance(cfg.dataset.num_classes, int) and cfg.dataset.ignore_index == 255
            assert len(cfg.optimizer.betas) == 2 and abs(cfg.optimizer.betas[0] - 0.9) < 1e-12
            assert isinstance(cfg.optimizer.learning_rate, float) and 0 < cfg.optimizer.learning_rate < 1e-2
            assert isinstance(cfg.scheduler.warmup_ratio, float) and cfg.scheduler.warmup_ratio in (1e-6, 1e-4)
            assert cfg.train.samples_per_gpu // 2 >= 1 and cfg.train.max_iters > 0            # train.py:138,197
    finally:
        if had is None:
            sys.modules.pop("omegaconf", None)


def test_dropin_runs_a_script_with_the_reference_import_lines(tmp_path):
    """`python -m segmif_b200.dropin script.py` with the import statements of the reference's entry scripts
    (train.py:18,25,111-112, test_fusion.py:9,16,30,44): sibling modules of the script resolve (python script.py semantics),
    `core.*` / `pytorch_ssim` / `lap_loss` / `utils.optimizer` / `omegaconf` resolve to the mirror, the config loads."""
    import subprocess
    import sys
    (tmp_path / "TaskFusion_dataset2.py").write_text("class Fusion_dataset:\n    pass\n")
    (tmp_path / "configs").mkdir()
    (tmp_path / "configs" / "voc.yaml").write_text(open(os.path.join(ROOT, "configs", "voc.yaml")).read())
    (tmp_path / "entry.py").write_text(
        "from core.model_fusion import Mean, Network3, Fusion_Network3_ac\n"
        "from core import Total_fusion_loss, Total_fusion_loss2, RGB2YCrCb, Fusionloss, Fusionloss_add, Fusionloss2, Fusionloss3, Fusionloss4, Fusionloss_grad3\n"
        "from core.model_fusion import RGB2YCrCb as R2, YCrCb2RGB\n"
        "from TaskFusion_dataset2 import Fusion_dataset\n"
        "from utils.optimizer import PolyWarmupAdamW, PolyWarmupAdamW_seg\n"
        "import pytorch_ssim, lap_loss\n"
        "from omegaconf import OmegaConf\n"
        "cfg = OmegaConf.load('configs/voc.yaml')\n"
        "m = Network3(cfg.exp.backbone, cfg.dataset.num_classes, 256, None)\n"
        "print('DROPIN_OK', cfg.exp.backbone, sum(p.numel() for p in m.parameters()))\n")
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-m", "segmif_b200.dropin", "entry.py"], cwd=str(tmp_path), env=env, capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "DROPIN_OK mit_b3 44604" in r.stdout.replace(",", ""), r.stdout        # 44.605 M parameters
