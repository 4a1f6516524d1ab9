"""Autograd wiring of the mixer's CUDA kernels (forward + backward through the C ABI).

``mamba_inner_tm`` is the token-major equivalent of mamba-ssm's ``mamba_inner_fn``:
in_proj -> causal conv1d + SiLU -> x_proj -> dt_proj -> selective scan -> out_proj, with the
GEMMs on cuBLAS (tensor cores) and conv / scan on the hand-written kernels.
"""

from __future__ import annotations

import torch
import torch.nn.functional as F

from . import ops


def _amp_dtype(x: torch.Tensor) -> torch.dtype:
    """Activation dtype of the mixer: the autocast dtype when autocast is on (runner_pretrain.py:243), else x's."""
    if torch.is_autocast_enabled("cuda"):
        return torch.get_autocast_dtype("cuda")
    return x.dtype


def mamba_inner_tm(hidden, in_proj_w, conv_w, conv_b, x_proj_w, dt_proj_w, dt_proj_b, A_log, D, out_proj_w,
                   dt_rank: int, d_state: int) -> torch.Tensor:
    """Token-major Mamba mixer body.  hidden (B, L, d_model) -> (B, L, d_model)."""
    d_inner = conv_w.shape[0]
    act = _amp_dtype(hidden)
    need_grad = torch.is_grad_enabled() and any(
        t.requires_grad for t in (hidden, in_proj_w, conv_w, conv_b, x_proj_w, dt_proj_w, dt_proj_b, A_log, D,
                                  out_proj_w))
    xz = F.linear(hidden.to(act), in_proj_w.to(act))  # (B, L, 2*d_inner)
    x, z = xz[..., :d_inner], xz[..., d_inner:]
    A = -torch.exp(A_log.float())
    if need_grad:
        u = ops.CausalConv1dTM.apply(x, conv_w, conv_b, True)
    else:
        u = ops.causal_conv1d_tm(x, conv_w, conv_b, silu=True)
    x_dbl = F.linear(u, x_proj_w.to(act))  # (B, L, dt_rank + 2*d_state)
    dt = F.linear(x_dbl[..., :dt_rank], dt_proj_w.to(act))  # bias is applied inside the scan
    Bm = x_dbl[..., dt_rank:dt_rank + d_state]
    Cm = x_dbl[..., dt_rank + d_state:]
    if need_grad:
        y = ops.SelectiveScanTM.apply(u, dt, A, Bm, Cm, D, z, dt_proj_b, True)
    else:
        y = ops.selective_scan_tm(u, dt, A, Bm, Cm, D, z, dt_proj_b, delta_softplus=True)
    return F.linear(y, out_proj_w.to(act))
