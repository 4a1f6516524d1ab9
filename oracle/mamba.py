"""Oracle: causal depthwise conv1d, selective scan, Mamba mixer, Block, MixerModel.

Test infrastructure (see oracle/__init__.py).  The arithmetic of these rows lives
in third-party wheels that are NOT in the reference tree (README.md:55-56:
causal-conv1d==1.1.1, mamba-ssm==1.1.1); this file restates their published
reference semantics (``selective_scan_ref``, ``causal_conv1d_ref``, the
non-fused ``Mamba.forward`` path) and the reference's own glue:

  * Block.forward       models/block.py:47-73
  * MixerModel.forward  models/point_mamba.py:247-272
  * Mamba ctor call     models/point_mamba.py:162 (d_state 16, d_conv 4, expand 2)

Layout: like mamba-ssm the scan takes channel-major (B, D, L) tensors.
"""

from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def causal_conv1d_ref(x, weight, bias=None, activation="silu"):
    """x (B,D,L), weight (D,W), bias (D,) -> (B,D,L); zero left-pad; optional SiLU.  fp32 math."""
    dtype_in = x.dtype
    x = x.float()
    D, W = weight.shape
    L = x.shape[-1]
    out = F.conv1d(x, weight.float().unsqueeze(1), None if bias is None else bias.float(), padding=W - 1, groups=D)
    out = out[..., :L]
    if activation in ("silu", "swish"):
        out = F.silu(out)
    return out.to(dtype_in)


def selective_scan_ref(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                       return_last_state=False):
    """mamba-ssm ``selective_scan_ref`` semantics (real A, input-dependent B/C).

    u, delta, z (B,D,L); A (D,N); B, C (B,N,L); D, delta_bias (D,).
    h_t = exp(delta_t A) h_{t-1} + delta_t B_t u_t ; y_t = <h_t, C_t> + D u_t ; out = y * silu(z).
    All math fp32, cast to the input dtype at the end.
    """
    dtype_in = u.dtype
    u = u.float()
    delta = delta.float()
    if delta_bias is not None:
        delta = delta + delta_bias[..., None].float()
    if delta_softplus:
        delta = F.softplus(delta)
    batch, dim, L = u.shape
    N = A.shape[1]
    B = B.float()
    C = C.float()
    A = A.float()
    x = A.new_zeros((batch, dim, N))
    ys = []
    deltaA = torch.exp(torch.einsum("bdl,dn->bdln", delta, A))
    deltaB_u = torch.einsum("bdl,bnl,bdl->bdln", delta, B, u)
    for i in range(L):
        x = deltaA[:, :, i] * x + deltaB_u[:, :, i]
        y = torch.einsum("bdn,bn->bd", x, C[:, :, i])
        ys.append(y)
    y = torch.stack(ys, dim=2)
    out = y if D is None else y + u * D.float()[None, :, None]
    if z is not None:
        out = out * F.silu(z.float())
    out = out.to(dtype_in)
    return (out, x) if return_last_state else out


def selective_scan_fp64(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False):
    """Same recurrence in fp64 - the 'truth' tolerance tests measure both fp32 paths against."""
    args = [t.double() if t is not None else None for t in (u, delta, A, B, C, D, z, delta_bias)]
    u, delta, A, B, C, D, z, delta_bias = args
    if delta_bias is not None:
        delta = delta + delta_bias[..., None]
    if delta_softplus:
        delta = F.softplus(delta)
    batch, dim, L = u.shape
    x = torch.zeros(batch, dim, A.shape[1], dtype=torch.float64)
    out = torch.empty_like(u)
    for i in range(L):
        dA = torch.exp(delta[:, :, i, None] * A[None])
        x = dA * x + (delta[:, :, i] * u[:, :, i])[..., None] * B[:, None, :, i]
        out[:, :, i] = (x * C[:, None, :, i]).sum(-1)
    if D is not None:
        out = out + u * D[None, :, None]
    if z is not None:
        out = out * F.silu(z)
    return out


def mamba_mixer(sd: dict, prefix: str, hidden: torch.Tensor) -> torch.Tensor:
    """``Mamba.forward`` (mamba-ssm mamba_simple, non-fused path) from a state dict.

    hidden (B,L,d_model) -> (B,L,d_model).  Parameter names/shapes are the ones
    in logs/finetuned_modelnet40.log (in_proj.weight (1536,384), conv1d.weight
    (768,1,4), x_proj.weight (56,768), dt_proj.weight (768,24), A_log (768,16), D,
    out_proj.weight (384,768)); no in/out-proj bias.
    """
    W_in = sd[prefix + "in_proj.weight"]
    conv_w = sd[prefix + "conv1d.weight"]
    conv_b = sd[prefix + "conv1d.bias"]
    W_x = sd[prefix + "x_proj.weight"]
    W_dt = sd[prefix + "dt_proj.weight"]
    b_dt = sd[prefix + "dt_proj.bias"]
    A = -torch.exp(sd[prefix + "A_log"].float())
    Dp = sd[prefix + "D"].float()
    W_out = sd[prefix + "out_proj.weight"]
    d_inner, dt_rank = W_dt.shape
    d_state = A.shape[1]
    Bsz, L, _ = hidden.shape
    xz = (hidden @ W_in.t()).transpose(1, 2)  # (B, 2*d_inner, L)
    if (prefix + "in_proj.bias") in sd:
        xz = xz + sd[prefix + "in_proj.bias"][None, :, None]
    x, z = xz[:, :d_inner], xz[:, d_inner:]
    x = causal_conv1d_ref(x, conv_w[:, 0, :], conv_b, "silu")
    x_dbl = x.transpose(1, 2).reshape(Bsz * L, d_inner) @ W_x.t()
    dt, Bm, Cm = torch.split(x_dbl, [dt_rank, d_state, d_state], dim=-1)
    dt = (dt @ W_dt.t()).reshape(Bsz, L, d_inner).transpose(1, 2)
    Bm = Bm.reshape(Bsz, L, d_state).transpose(1, 2)
    Cm = Cm.reshape(Bsz, L, d_state).transpose(1, 2)
    y = selective_scan_ref(x, dt, A, Bm, Cm, Dp, z=z, delta_bias=b_dt.float(), delta_softplus=True)
    out = y.transpose(1, 2) @ W_out.t()
    if (prefix + "out_proj.bias") in sd:
        out = out + sd[prefix + "out_proj.bias"]
    return out


def layer_norm(sd, prefix, x, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + "weight"], sd[prefix + "bias"], eps)


def mixer_model(sd: dict, prefix: str, tokens, pos, n_layer: int, eps: float = 1e-5, fetch_idx=None,
                mixer=None, trace=None, drop_scale=None):
    """MixerModel.forward (point_mamba.py:247-258); eval mode (DropPath/Dropout identity) unless ``drop_scale`` is
    given: a list with one (B,) tensor per layer i >= 1 holding timm DropPath's per-sample factor mask / keep that
    multiplies h in ``drop_path(h) + residual`` (block.py:59; layer 0 has no residual yet, so its input is never dropped).

    Block.forward (block.py:56-72): residual = h (+ residual); h = LN(residual); h = mixer(h).
    With ``fetch_idx`` returns the seg variant's list of norm_f(h + residual) taps
    (pt_mamba.py:390-416).  ``mixer`` (default: mamba_mixer) and ``trace`` (list that receives each layer's
    (h, residual)) exist so tests/test_oracle_vs_reference.py can pin this loop against the reference's own Block
    driven with a stand-in mixer."""
    mixer = mixer or mamba_mixer
    h = tokens + pos
    residual = None
    taps = []
    for i in range(n_layer):
        if residual is None:
            residual = h
        elif drop_scale is not None:
            residual = h * drop_scale[i - 1].view(-1, 1, 1) + residual
        else:
            residual = h + residual
        h = layer_norm(sd, f"{prefix}layers.{i}.norm.", residual, eps)
        h = mixer(sd, f"{prefix}layers.{i}.mixer.", h)
        if trace is not None:
            trace.append((h, residual))
        if fetch_idx is not None and i in fetch_idx:
            taps.append(layer_norm(sd, prefix + "norm_f.", h + residual, eps))
    if fetch_idx is not None:
        return taps
    return layer_norm(sd, prefix + "norm_f.", h + residual, eps)


def init_mamba_params(d_model=384, d_state=16, d_conv=4, expand=2, n_layer=12, seed=0,
                      dt_min=1e-3, dt_max=0.1, dt_init_floor=1e-4):
    """Upstream Mamba init for ONE mixer, as a plain dict of tensors."""
    g = torch.Generator().manual_seed(seed)
    d_inner = expand * d_model
    dt_rank = math.ceil(d_model / 16)
    sd = {}

    def kaiming_uniform(shape, fan_in):
        bound = 1.0 / math.sqrt(fan_in)
        return (torch.rand(shape, generator=g) * 2 - 1) * bound

    sd["in_proj.weight"] = kaiming_uniform((2 * d_inner, d_model), d_model)
    sd["conv1d.weight"] = kaiming_uniform((d_inner, 1, d_conv), d_conv)
    sd["conv1d.bias"] = kaiming_uniform((d_inner,), d_conv)
    sd["x_proj.weight"] = kaiming_uniform((dt_rank + 2 * d_state, d_inner), d_inner)
    std = dt_rank ** -0.5
    sd["dt_proj.weight"] = (torch.rand((d_inner, dt_rank), generator=g) * 2 - 1) * std
    dt = torch.exp(torch.rand(d_inner, generator=g) * (math.log(dt_max) - math.log(dt_min)) + math.log(dt_min))
    dt = dt.clamp(min=dt_init_floor)
    sd["dt_proj.bias"] = dt + torch.log(-torch.expm1(-dt))
    sd["A_log"] = torch.log(torch.arange(1, d_state + 1, dtype=torch.float32))[None].repeat(d_inner, 1)
    sd["D"] = torch.ones(d_inner)
    sd["out_proj.weight"] = kaiming_uniform((d_model, d_inner), d_inner) / math.sqrt(n_layer)
    return sd
