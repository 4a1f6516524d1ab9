"""Trainer plumbing on CPU: the optimizer groups / scheduler / checkpoint format of tools/builder.py, the
synthetic dataset, and the bucketed gradient all-reduce against plain averaging with two gloo ranks."""

import math
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from si_mamba_b200 import train
from si_mamba_b200.config import Config


class _Tiny(nn.Module):
    def __init__(self):
        super().__init__()
        self.fc = nn.Linear(4, 8)
        self.norm = nn.LayerNorm(8)
        self.mask_token = nn.Parameter(torch.zeros(1, 1, 8))
        self.A_log = nn.Parameter(torch.zeros(8, 2))
        self.unused = nn.Linear(3, 3)
        self.out = nn.Linear(8, 2, bias=False)

    def forward(self, x):
        return self.out(self.norm(self.fc(x)) + self.mask_token[0] + self.A_log.sum(-1))


def _ocfg(lr=1e-3):
    return Config(optimizer=Config(type="AdamW", kwargs=Config(lr=lr, weight_decay=0.05)),
                  scheduler=Config(type="CosLR", kwargs=Config(epochs=300, initial_epochs=10)))


def test_no_decay_rule_matches_builder():
    """tools/builder.py:60-73: 1-D tensors, *.bias and names containing 'token' do not decay; 2-D weights (A_log too) do."""
    m = _Tiny()
    m.fc.bias.requires_grad_(False)  # frozen weights are skipped
    groups = train.add_weight_decay(m, 0.05)
    names = {id(p): n for n, p in m.named_parameters()}
    no_decay = sorted(names[id(p)] for p in groups[0]["params"])
    decay = sorted(names[id(p)] for p in groups[1]["params"])
    assert groups[0]["weight_decay"] == 0.0 and groups[1]["weight_decay"] == 0.05
    assert no_decay == ["mask_token", "norm.bias", "norm.weight", "unused.bias"]
    assert decay == ["A_log", "fc.weight", "out.weight", "unused.weight"]
    # a DDP-style wrapper exposes the model as .module (builder.py:63 walks model.module)
    wrapped = nn.Module()
    wrapped.module = m
    assert [len(g["params"]) for g in train.add_weight_decay(wrapped, 0.05)] == [4, 4]


def test_cosine_scheduler_values():
    """CosineLRScheduler(t_initial=300, lr_min=1e-6, warmup_lr_init=1e-6, warmup_t=10, cycle_limit=1, cycle_decay=0.1)."""
    m = _Tiny()
    opt, sch = train.build_opti_sche(m, _ocfg(1e-3))
    assert all(abs(g["lr"] - 1e-6) < 1e-12 for g in opt.param_groups)      # construction sets the warm-up start
    sch.step(0)
    assert abs(opt.param_groups[0]["lr"] - 1e-6) < 1e-12
    sch.step(5)
    assert abs(opt.param_groups[1]["lr"] - (1e-6 + 5 * (1e-3 - 1e-6) / 10)) < 1e-12
    sch.step(10)   # no warm-up prefix: the cosine is evaluated at t = 10 of 300
    want = 1e-6 + 0.5 * (1e-3 - 1e-6) * (1 + math.cos(math.pi * 10 / 300))
    assert abs(opt.param_groups[0]["lr"] - want) < 1e-12
    sch.step(150)
    assert abs(opt.param_groups[0]["lr"] - (1e-6 + 0.5 * (1e-3 - 1e-6))) < 1e-9
    sch.step(300)  # past cycle_limit
    assert opt.param_groups[0]["lr"] == 1e-6


def test_checkpoint_round_trip_in_reference_layout(tmp_path):
    """save_checkpoint writes {base_model, optimizer, epoch, metrics, best_metrics}; resume_* read it back, also when the
    keys carry DDP's 'module.' prefix (tools/builder.py:112-161)."""
    torch.manual_seed(0)
    m = _Tiny()
    opt, _ = train.build_opti_sche(m, _ocfg())
    m(torch.randn(5, 4)).sum().backward()
    opt.step()
    path = train.save_checkpoint(m, opt, 7, train.Acc_Metric(91.5), train.Acc_Metric(92.0), "ckpt-last", str(tmp_path))
    sd = torch.load(path, weights_only=False)
    assert sorted(sd) == ["base_model", "best_metrics", "epoch", "metrics", "optimizer"]
    assert sd["epoch"] == 7 and sd["best_metrics"] == {"acc": 92.0} and sd["metrics"] == {"acc": 91.5}
    assert all(not k.startswith("module.") for k in sd["base_model"])
    # a checkpoint written by the reference under DDP has module.-prefixed keys
    sd["base_model"] = {"module." + k: v for k, v in sd["base_model"].items()}
    torch.save(sd, path)
    m2 = _Tiny()
    opt2, _ = train.build_opti_sche(m2, _ocfg())
    start, best = train.resume_model(m2, str(tmp_path))
    train.resume_optimizer(opt2, str(tmp_path))
    assert start == 8 and best == {"acc": 92.0}
    for a, b in zip(m.state_dict().values(), m2.state_dict().values()):
        assert torch.equal(a, b)
    s1, s2 = opt.state_dict()["state"], opt2.state_dict()["state"]
    assert s1.keys() == s2.keys() and all(torch.equal(s1[k]["exp_avg"], s2[k]["exp_avg"]) for k in s1)
    assert train.resume_model(m2, str(tmp_path / "nowhere")) == (0, 0)
    ep, _ = train.load_model(m2, path)
    assert ep == 7
    assert train.save_checkpoint(m, opt, 7, None, None, "x", str(tmp_path), rank=1) is None


def test_synthetic_dataset_is_deterministic_and_normalised():
    ds = train.SyntheticClouds(10, 256, task="seg", seed=3)
    p0, lab, parts = ds[4]
    p1, _, _ = train.SyntheticClouds(10, 256, task="seg", seed=3)[4]
    assert torch.equal(p0, p1) and p0.shape == (256, 3) and parts.shape == (256,) and 0 <= lab < 16
    assert p0.mean(0).abs().max() < 1e-5 and abs(float(p0.norm(dim=-1).max()) - 1) < 1e-5
    assert not torch.equal(p0, ds[5][0])
    s0 = list(train.shard_sampler(ds, 0, 2, shuffle=False))
    s1 = list(train.shard_sampler(ds, 1, 2, shuffle=False))
    assert sorted(s0 + s1) == list(range(10))


def test_gradsync_single_process_flat_views_and_clip():
    torch.manual_seed(0)
    m = _Tiny()
    sync = train.GradSync(m, bucket_mb=1e-4)  # tiny buckets: several of them
    assert len(sync.buckets) > 2
    x = torch.randn(6, 4)
    m(x).pow(2).sum().backward()
    sync.finish()
    assert sync.prune_unused() == 2  # unused.weight / unused.bias never receive a gradient
    assert m.unused.weight.grad is None and m.fc.weight.grad is not None
    assert m.fc.weight.grad.data_ptr() == sync.flat[sync._offs[m.fc.weight]:].data_ptr()
    ref = _Tiny()
    ref.load_state_dict(m.state_dict())
    ref(x).pow(2).sum().backward()
    for (n, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
        if q.grad is not None:
            assert torch.allclose(p.grad, q.grad), n
    total = torch.nn.utils.clip_grad_norm_([p for p in ref.parameters() if p.grad is not None], 0.5)
    got = sync.clip_grad_norm_(0.5)
    assert torch.allclose(total, got)
    for p, q in zip(m.parameters(), ref.parameters()):
        if q.grad is not None:
            assert torch.allclose(p.grad, q.grad, atol=1e-7)
    sync.begin_step()
    assert m.fc.weight.grad is None
    m(x).pow(2).sum().backward()   # second step: fresh gradients, packed into the same flat slices by finish()
    sync.finish()
    assert m.fc.weight.grad.data_ptr() == sync.flat[sync._offs[m.fc.weight]:].data_ptr()
    assert m.unused.weight.grad is None


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    torch.manual_seed(0)
    m = _Tiny()
    sync = train.GradSync(m, bucket_mb=1e-4)
    xs = torch.randn(2, 6, 4, generator=torch.Generator().manual_seed(5))
    for step in range(2):  # the second step runs with the unused parameters pruned
        sync.begin_step()
        m(xs[rank]).pow(2).sum().backward()
        sync.finish()
        if step == 0:
            sync.prune_unused()
    grads = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
    if rank == 0:
        q.put((grads, sync.allreduce_launches))
    dist.barrier()
    dist.destroy_process_group()


def test_gradsync_two_ranks_average_gradients():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    grads, launches = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.manual_seed(0)
    ref = _Tiny()
    xs = torch.randn(2, 6, 4, generator=torch.Generator().manual_seed(5))
    (0.5 * (ref(xs[0]).pow(2).sum() + ref(xs[1]).pow(2).sum())).backward()
    assert "unused.weight" not in grads
    for n, p in ref.named_parameters():
        if p.grad is not None:
            assert torch.allclose(grads[n], p.grad, atol=1e-6), n
    assert launches >= 2 * 3  # every bucket, every step
