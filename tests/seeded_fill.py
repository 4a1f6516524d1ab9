"""Deterministic weights from (name, shape) alone, shared by tools/make_reference_golden.py and the tests so that
large state dicts need not be stored in fixtures: only the ordered (name, shape) list and the seed are."""

import math

import torch


def seeded_state_dict(spec, seed):
    """spec: ordered list of (name, shape) of the floating-point entries -> {name: fp32 tensor}."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shape in spec:
        shape = tuple(shape)
        if name.endswith("running_var"):
            t = 0.5 + torch.rand(shape, generator=g)
        elif name.endswith("running_mean") or name.endswith("bias"):
            t = 0.1 * torch.randn(shape, generator=g)
        elif len(shape) == 1:  # norm scales
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            t = torch.randn(shape, generator=g) / math.sqrt(fan_in)
        out[name] = t
    return out
