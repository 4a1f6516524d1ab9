"""``Mamba`` mixer with the constructor / parameters / forward(hidden_states) of
mamba_ssm.modules.mamba_simple.Mamba as SI-Mamba instantiates it
(models/point_mamba.py:162 via partial(Mamba, layer_idx=..., **ssm_cfg); called at models/block.py:72).

State-dict keys and shapes are the reference's (logs/finetuned_modelnet40.log parameter table):
A_log (768,16), D (768), in_proj.weight (1536,384), conv1d.weight (768,1,4), conv1d.bias,
x_proj.weight (56,768), dt_proj.weight (768,24), dt_proj.bias, out_proj.weight (384,768).

B200 data path: activations stay TOKEN-major (B, L, channels) end to end, so the projection
GEMMs (cuBLAS, tensor cores) read and write row-major matrices without a transpose and the conv /
scan kernels consume x, z, B, C in place as column slices of the GEMM outputs.
"""

from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .autograd import mamba_inner_tm


class Mamba(nn.Module):
    def __init__(self, d_model, d_state=16, d_conv=4, expand=2, dt_rank="auto", dt_min=0.001, dt_max=0.1,
                 dt_init="random", dt_scale=1.0, dt_init_floor=1e-4, conv_bias=True, bias=False,
                 use_fast_path=True, layer_idx=None, device=None, dtype=None):
        factory_kwargs = {"device": device, "dtype": dtype}
        super().__init__()
        self.d_model = d_model
        self.d_state = d_state
        self.d_conv = d_conv
        self.expand = expand
        self.d_inner = int(self.expand * self.d_model)
        self.dt_rank = math.ceil(self.d_model / 16) if dt_rank == "auto" else dt_rank
        self.use_fast_path = use_fast_path
        self.layer_idx = layer_idx

        self.in_proj = nn.Linear(self.d_model, self.d_inner * 2, bias=bias, **factory_kwargs)
        self.conv1d = nn.Conv1d(self.d_inner, self.d_inner, bias=conv_bias, kernel_size=d_conv,
                                groups=self.d_inner, padding=d_conv - 1, **factory_kwargs)
        self.activation = "silu"
        self.act = nn.SiLU()
        self.x_proj = nn.Linear(self.d_inner, self.dt_rank + self.d_state * 2, bias=False, **factory_kwargs)
        self.dt_proj = nn.Linear(self.dt_rank, self.d_inner, bias=True, **factory_kwargs)

        # upstream initialisation
        dt_init_std = self.dt_rank ** -0.5 * dt_scale
        if dt_init == "constant":
            nn.init.constant_(self.dt_proj.weight, dt_init_std)
        elif dt_init == "random":
            nn.init.uniform_(self.dt_proj.weight, -dt_init_std, dt_init_std)
        else:
            raise NotImplementedError
        dt = torch.exp(torch.rand(self.d_inner, **factory_kwargs) * (math.log(dt_max) - math.log(dt_min))
                       + math.log(dt_min)).clamp(min=dt_init_floor)
        inv_dt = dt + torch.log(-torch.expm1(-dt))
        with torch.no_grad():
            self.dt_proj.bias.copy_(inv_dt)
        self.dt_proj.bias._no_reinit = True

        A = torch.arange(1, self.d_state + 1, dtype=torch.float32, device=device).repeat(self.d_inner, 1).contiguous()
        self.A_log = nn.Parameter(torch.log(A))
        self.A_log._no_weight_decay = True
        self.D = nn.Parameter(torch.ones(self.d_inner, device=device))
        self.D._no_weight_decay = True
        self.out_proj = nn.Linear(self.d_inner, self.d_model, bias=bias, **factory_kwargs)

    def forward(self, hidden_states, inference_params=None):
        """hidden_states (B, L, d_model) -> (B, L, d_model)."""
        if inference_params is not None:
            raise NotImplementedError("step-wise decoding is not on SI-Mamba's path (inference_params is always None)")
        if self.in_proj.bias is not None or self.out_proj.bias is not None:
            raise NotImplementedError("SI-Mamba builds Mamba with bias=False")
        return mamba_inner_tm(hidden_states, self.in_proj.weight, self.conv1d.weight, self.conv1d.bias,
                              self.x_proj.weight, self.dt_proj.weight, self.dt_proj.bias, self.A_log, self.D,
                              self.out_proj.weight, self.dt_rank, self.d_state)

    def wants_split3(self, hidden_states) -> bool:
        """Whether forward() takes its input as an ops.Split3 (written by the Block's fused add + LayerNorm)."""
        from .autograd import wants_split3
        return (hidden_states.is_cuda and not hidden_states.requires_grad
                and wants_split3(hidden_states.dtype if not torch.is_autocast_enabled("cuda") else
                                 torch.get_autocast_dtype("cuda"), self.in_proj.weight, self.d_model,
                                 tuple(self.parameters()))
                and self.d_inner % 64 == 0)

    def allocate_inference_cache(self, batch_size, max_seqlen, dtype=None, **kwargs):
        raise NotImplementedError("step-wise decoding is not on SI-Mamba's path")
