"""Trainer plumbing on the GPU: the single-kernel FlatAdamW against torch.optim.AdamW (same groups, same gradients, a
parameter without gradient, the clip coefficient applied in-kernel), its checkpoint layout, and a captured TrainStep."""

import copy

import pytest
import torch
import torch.nn as nn

from si_mamba_b200 import train
from si_mamba_b200.config import Config

pytestmark = pytest.mark.gpu


class _Tiny(nn.Module):
    def __init__(self):
        super().__init__()
        self.fc = nn.Linear(16, 32)
        self.norm = nn.LayerNorm(32)
        self.mask_token = nn.Parameter(torch.zeros(1, 1, 32))
        self.unused = nn.Linear(3, 5)
        self.out = nn.Linear(32, 7, bias=False)

    def forward(self, x):
        return self.out(self.norm(self.fc(x)) + self.mask_token[0])


def _ocfg(lr=1e-2):
    return Config(optimizer=Config(type="AdamW", kwargs=Config(lr=lr, weight_decay=0.05)),
                  scheduler=Config(type="CosLR", kwargs=Config(epochs=300, initial_epochs=10)))


def test_flat_adamw_matches_torch_adamw(lib):
    torch.manual_seed(0)
    m1 = _Tiny().cuda()
    m2 = copy.deepcopy(m1)
    sync = train.GradSync(m1, bucket_mb=1e-4)
    opt1, sch1 = train.build_opti_sche(m1, _ocfg(), sync=sync)
    assert isinstance(opt1, train.FlatAdamW)
    opt2 = torch.optim.AdamW(train.add_weight_decay(m2, 0.05), lr=1e-2, weight_decay=0.05)
    sch2 = train.CosineLRScheduler(opt2, t_initial=300, lr_min=1e-6, cycle_decay=0.1, warmup_lr_init=1e-6, warmup_t=10)
    unused_before = m1.unused.weight.detach().clone()
    g = torch.Generator().manual_seed(1)
    for step in range(6):
        x = torch.randn(9, 16, generator=g).cuda()
        for sch in (sch1, sch2):
            sch.step(step + 3)  # some point of the warm-up ramp
        sync.begin_step()
        m1(x).pow(2).sum().backward()
        sync.finish()
        total1, opt1.grad_scale = sync.clip_coef(0.5)
        opt1.step()
        opt2.zero_grad(set_to_none=True)
        m2(x).pow(2).sum().backward()
        total2 = torch.nn.utils.clip_grad_norm_(m2.parameters(), 0.5)
        opt2.step()
        assert torch.allclose(total1, total2, rtol=1e-5)
    for (n, p), (_, q) in zip(m1.named_parameters(), m2.named_parameters()):
        assert torch.allclose(p, q, rtol=2e-5, atol=2e-6), n
    assert torch.equal(m1.unused.weight, unused_before)  # no gradient -> untouched (no weight decay either), as torch
    # checkpoint layout = torch.optim.AdamW's: the state dict loads into a torch AdamW and back
    sd = opt1.state_dict()
    assert set(next(iter(sd["state"].values()))) == {"step", "exp_avg", "exp_avg_sq"}
    opt3 = torch.optim.AdamW(train.add_weight_decay(copy.deepcopy(m1), 0.05), lr=1e-2, weight_decay=0.05)
    opt3.load_state_dict(sd)
    s2 = opt2.state_dict()["state"]
    for k, st in sd["state"].items():
        assert float(st["step"]) == 6.0
        assert torch.allclose(st["exp_avg"], s2[k]["exp_avg"], rtol=2e-5, atol=1e-7)
    m4 = copy.deepcopy(m2)
    sync4 = train.GradSync(m4)
    opt4, _ = train.build_opti_sche(m4, _ocfg(), sync=sync4)
    opt4.load_state_dict(opt2.state_dict())
    assert float(opt4.step_t) == 6.0
    assert torch.allclose(opt4.state[m4.fc.weight]["exp_avg_sq"], opt2.state[m2.fc.weight]["exp_avg_sq"])


def test_train_step_graph_matches_eager(lib):
    """TrainStep: the captured step (forward + backward + packed gradients + clip coefficient + FlatAdamW in one CUDA graph)
    follows the same trajectory as the eager step."""
    torch.manual_seed(0)
    base = _Tiny().cuda()
    xs = [torch.randn(9, 16, generator=torch.Generator().manual_seed(i)).cuda() for i in range(5)]
    finals = []
    for graph in (False, True):
        m = copy.deepcopy(base)
        x = xs[0].clone()
        sync = train.GradSync(m)
        opt, _ = train.build_opti_sche(m, _ocfg(), sync=sync)
        step = train.TrainStep(m, opt, lambda: m(x).pow(2).mean(), grad_clip=10.0, graph=graph, sync=sync, warmup=1)
        for i in range(4):
            x.copy_(xs[i + 1])
            step()
        finals.append((m.fc.weight.detach().clone(), float(opt.step_t)))
    # both variants: one eager step on xs[0] (the graphed one as its warm-up; the capture itself executes nothing), then
    # four steps on xs[1..4]
    assert finals[0][1] == 5.0 and finals[1][1] == 5.0
    assert torch.allclose(finals[0][0], finals[1][0], rtol=1e-4, atol=1e-6)
