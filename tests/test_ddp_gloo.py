"""Host-side logic of the data-parallel training path, exercised with world_size 2 over gloo on CPU: the flat
parameter / gradient buffers, the single all-reduce, exclusion of parameters that never receive gradients, and the
PolyWarmup schedule.  (The fused AdamW kernel itself is CUDA-only and is checked on the GPU.)"""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


class _Toy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.a = torch.nn.Linear(4, 3)
        self.ffm2 = torch.nn.Linear(4, 3)          # never used, like Fusion_Network3_ac.ffm2
        self.b = torch.nn.Conv2d(2, 2, 3)

    def forward(self, x, img):
        return self.a(x).sum() + self.b(img).sum()


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from segmif_b200.ddp import FlatParams
    m = _Toy()
    before = {k: p.detach().clone() for k, p in m.named_parameters()}
    flat = FlatParams(m, used=lambda k: not k.startswith("ffm2."))
    assert flat.skipped == ["ffm2.weight", "ffm2.bias"]
    for k, p in m.named_parameters():                       # values preserved, storage re-homed
        assert torch.equal(p.detach(), before[k])
        if not k.startswith("ffm2."):
            off, n = flat.offsets[k]
            assert p.data_ptr() == flat.param.data_ptr() + 4 * off
    flat.zero_grad()
    g = torch.Generator().manual_seed(100 + rank)          # each rank sees its own shard of the batch
    x, img = torch.randn(5, 4, generator=g), torch.randn(2, 2, 6, 6, generator=g)
    m(x, img).backward()
    local = flat.grad.clone()
    assert m.ffm2.weight.grad is None
    world_seen = flat.all_reduce()
    assert world_seen == world
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    assert torch.allclose(flat.grad, sum(gathered), atol=1e-6)
    assert m.a.weight.grad.data_ptr() == flat.grad.data_ptr() + 4 * flat.offsets["a.weight"][0]   # still views
    # a second step accumulates into the same (re-zeroed) buffer
    flat.zero_grad()
    assert float(flat.grad.abs().max()) == 0.0
    m(x, img).backward()
    assert torch.allclose(flat.grad, local, atol=1e-6)
    if rank == 0:
        torch.save(dict(ok=True, numel=flat.numel), out)
    dist.destroy_process_group()


def test_flat_gradient_allreduce_world2(tmp_path):
    out = str(tmp_path / "r0.pt")
    port = 29500 + (os.getpid() % 500)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["ok"] and r["numel"] == 12 + 4 + 36 + 4          # a.weight, a.bias (3 -> 4), b.weight, b.bias (2 -> 4)


def test_poly_warmup_schedule_matches_reference_optimizer():
    from segmif_b200.ddp import poly_warmup_lr
    from segmif_b200.utils.optimizer import PolyWarmupAdamW
    p = torch.nn.Parameter(torch.zeros(3))
    for warm in (10, 3e-5):
        opt = PolyWarmupAdamW([{"params": [p], "lr": 3e-4}], lr=3e-4, weight_decay=0.01, betas=(0.9, 0.999),
                              warmup_iter=warm, max_iter=40, warmup_ratio=1e-6, power=1.0)
        lr = 3e-4
        for step in range(50):
            p.grad = torch.ones(3)
            opt.step()
            lr = poly_warmup_lr(3e-4, step, warm, 40, 1e-6, 1.0, lr)
            assert abs(lr - opt.param_groups[0]["lr"]) < 1e-15, (warm, step)


def test_fusion_net_flat_params_exclude_ffm2():
    from segmif_b200.core.model_fusion import Fusion_Network3_ac
    from segmif_b200.ddp import FlatParams
    net = Fusion_Network3_ac()
    keys_before = list(net.state_dict().keys())
    flat = FlatParams(net, used=lambda k: not k.startswith("ffm2."))
    assert all(k.startswith("ffm2.") for k in flat.skipped) and flat.skipped
    assert flat.numel == sum((p.numel() + 3) // 4 * 4 for k, p in net.named_parameters() if not k.startswith("ffm2."))
    assert all(off % 4 == 0 for off, _ in flat.offsets.values())
    assert list(net.state_dict().keys()) == keys_before          # checkpoint surface unchanged


def test_training_forward_refuses_cpu():
    from segmif_b200.core.model_fusion import Fusion_Network3_ac
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    net = Fusion_Network3_ac().train()
    with pytest.raises(RuntimeError, match="no CPU path|CUDA"):
        net(torch.zeros(1, 1, 16, 16), torch.zeros(1, 3, 16, 16), torch.zeros(1, 64, 16, 16), torch.zeros(1, 128, 16, 16))


def test_seg_param_groups_are_contiguous_ranges_matching_get_param_groups():
    """train.py:170-189: WeTr.get_param_groups() -> [encoder weights, encoder norms, decoder (+ classifier)].  The flat
    buffer must hold each group as one contiguous range (one AdamW launch per group) with exactly those members."""
    from segmif_b200.core.model_fusion import Network3
    from segmif_b200.ddp import FlatParams
    net = Network3("mit_b0", 9, 256, None)
    ref_groups = net.denoise_net.get_param_groups()
    ids = [{id(p) for p in g} for g in ref_groups]
    gid = lambda k: 2 if ".decoder." in k else (1 if "norm" in k.split("encoder.", 1)[-1] else 0)
    keys_before = list(net.state_dict().keys())
    flat = FlatParams(net, used=lambda k: not k.endswith("classifier.weight"), group_of=gid)
    assert flat.skipped == ["denoise_net.classifier.weight"]
    assert list(net.state_dict().keys()) == keys_before
    lookup = dict(net.named_parameters())
    covered = 0
    for g in (0, 1, 2):
        lo, hi = flat.group_ranges[g]
        members = [k for k, (off, n) in flat.offsets.items() if lo <= off < hi]
        assert all(gid(k) == g for k in members)
        assert {id(lookup[k]) for k in members} == ids[g] - {id(net.denoise_net.classifier.weight)}
        covered += hi - lo
    assert covered == flat.numel                                  # the three ranges tile the buffer
    r = sorted(flat.group_ranges.values())
    assert r[0][0] == 0 and r[0][1] == r[1][0] and r[1][1] == r[2][0] and r[2][1] == flat.numel


def test_grouped_schedule_matches_reference_seg_optimizer():
    """PolyWarmupAdamW_seg (utils/optimizer.py:36-66): every group's lr follows its own base lr times the common
    multiplier, starting from iter_curr."""
    from segmif_b200.ddp import poly_warmup_lr
    from segmif_b200.utils.optimizer import PolyWarmupAdamW_seg
    ps = [torch.nn.Parameter(torch.zeros(3)) for _ in range(3)]
    opt = PolyWarmupAdamW_seg([{"params": [ps[0]], "lr": 6e-5, "weight_decay": 0.01}, {"params": [ps[1]], "lr": 6e-5, "weight_decay": 0.0},
                               {"params": [ps[2]], "lr": 6e-4, "weight_decay": 0.01}], lr=6e-5, weight_decay=0.01, betas=(0.9, 0.999),
                              iter_curr=5, warmup_iter=10, max_iter=40, warmup_ratio=1e-6, power=1.0)
    lrs = [6e-5, 6e-5, 6e-4]
    for step in range(5, 50):
        for p in ps:
            p.grad = torch.ones(3)
        opt.step()
        for i, base in enumerate((6e-5, 6e-5, 6e-4)):
            lrs[i] = poly_warmup_lr(base, step, 10, 40, 1e-6, 1.0, lrs[i])
            assert abs(lrs[i] - opt.param_groups[i]["lr"]) < 1e-15, (i, step)


def test_fused_optimizer_host_logic_with_a_stub_kernel(monkeypatch):
    """FusedPolyWarmupAdamW's host side (per-group schedule, contiguous ranges, bias-correction step count, pack-epoch
    bump) with the CUDA kernel replaced by a recorder -- the kernel itself is checked on the GPU."""
    from segmif_b200 import ops
    from segmif_b200.core.model_fusion import Network3
    from segmif_b200.ddp import FlatParams, FusedPolyWarmupAdamW, poly_warmup_lr
    calls = []
    monkeypatch.setattr(ops, "adamw_step", lambda param, grad, m, v, **kw: calls.append((param.data_ptr(), param.numel(), kw)))
    net = Network3("mit_b0", 9, 256, None)
    gid = lambda k: 2 if ".decoder." in k else (1 if "norm" in k.split("encoder.", 1)[-1] else 0)
    flat = FlatParams(net, used=lambda k: not k.endswith("classifier.weight"), group_of=gid)
    opt = FusedPolyWarmupAdamW(flat, 6e-5, 0.01, (0.9, 0.999), warmup_iter=4, max_iter=50, warmup_ratio=1e-6, power=1.0,
                               groups={0: dict(lr=6e-5, weight_decay=0.01), 1: dict(lr=6e-5, weight_decay=0.0),
                                       2: dict(lr=6e-4, weight_decay=0.01)}, iter_curr=3)
    epoch0 = flat.epoch[0]
    lrs = [6e-5, 6e-5, 6e-4]
    for it in range(3):
        calls.clear()
        opt.step(grad_scale=0.5)
        assert len(calls) == 3                                              # one launch per param group
        covered = 0
        for gi, (ptr, n, kw) in enumerate(calls):
            lo, hi = flat.group_ranges[gi]
            assert ptr == flat.param.data_ptr() + 4 * lo and n == hi - lo
            lrs[gi] = poly_warmup_lr((6e-5, 6e-5, 6e-4)[gi], 3 + it, 4, 50, 1e-6, 1.0, lrs[gi])
            assert abs(kw["lr"] - lrs[gi]) < 1e-18
            assert kw["weight_decay"] == (0.01, 0.0, 0.01)[gi] and kw["grad_scale"] == 0.5
            assert kw["step"] == it + 1                                     # AdamW's own step count starts at 1 (not iter_curr)
            covered += n
        assert covered == flat.numel
    assert flat.epoch[0] == epoch0 + 3 and opt.global_step == 6
    # a parameter owned by the flat buffer reports a new pack epoch after every optimizer step
    from segmif_b200.packing import _epoch
    p0 = flat.named[0][1]
    e = _epoch(p0)
    opt.step()
    assert _epoch(p0) != e


class _ToyBn(torch.nn.Module):
    def __init__(self, seed):
        super().__init__()
        torch.manual_seed(seed)
        self.a = torch.nn.Linear(4, 3)
        self.bn = torch.nn.BatchNorm1d(3)
        with torch.no_grad():
            self.bn.running_mean.normal_()


def _broadcast_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from segmif_b200.ddp import FlatParams
    m = _ToyBn(seed=10 + rank)                              # replicas constructed with DIFFERENT seeds
    flat = FlatParams(m)
    e0 = flat.epoch[0]
    assert flat.broadcast(m) is True
    assert flat.epoch[0] == e0 + 1                          # cached weight packs of the old values are declared stale
    mine = torch.cat([flat.param, m.bn.running_mean])
    gathered = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    assert all(torch.equal(gathered[0], t) for t in gathered)          # parameters AND buffers equal rank 0's
    ref = _ToyBn(seed=10)
    assert torch.equal(m.a.weight.detach(), ref.a.weight.detach()) and torch.equal(m.bn.running_mean, ref.bn.running_mean)
    if rank == 0:
        torch.save(dict(ok=True), out)
    dist.destroy_process_group()


def test_replicas_start_from_rank0_world2(tmp_path):
    """only gradients are exchanged per step, so replicas must be made identical once at construction."""
    out = str(tmp_path / "b0.pt")
    port = 29500 + ((os.getpid() + 137) % 500)
    mp.spawn(_broadcast_worker, args=(2, port, out), nprocs=2, join=True)
    assert torch.load(out)["ok"]
