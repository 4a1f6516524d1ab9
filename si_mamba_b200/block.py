"""``Block``: Add -> LayerNorm -> mixer, the reference's models/block.py:17-76 interface
(ctor arguments, forward(hidden_states, residual, inference_params) -> (hidden_states, residual),
attributes .mixer / .norm / .drop_path / .layer_idx).

B200 path: the residual add and the LayerNorm run as ONE kernel (sim_add_layernorm) that reads
hidden + residual once and writes the fp32 residual stream and the normalised activations; the
reference issues them as separate ATen kernels (its Triton fused_add_norm is never enabled).
"""

from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
from torch import Tensor

from . import ops
from .autograd import _amp_dtype


class DropPath(nn.Module):
    """timm 0.4.5 DropPath semantics (per-sample Bernoulli, scale 1/keep), train mode only."""

    def __init__(self, drop_prob: float = 0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        mask = (keep + torch.rand(shape, dtype=x.dtype, device=x.device)).floor_()
        return x.div(keep) * mask

    def sample_scale(self, x):
        """The per-sample factor mask / keep of forward(), (B,) fp32, drawn exactly like forward() draws its mask (same
        shape, dtype and generator state), for kernels that fold DropPath in."""
        keep = 1.0 - self.drop_prob
        shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        mask = (keep + torch.rand(shape, dtype=x.dtype, device=x.device)).floor_()
        return (mask.float() / keep).reshape(-1)


def fused_add_norm(norm: nn.LayerNorm, hidden: Tensor, residual: Optional[Tensor], want_residual: bool = True,
                   split: bool = False, row_scale: Optional[Tensor] = None):
    """(LayerNorm(hidden + residual), hidden + residual) through the CUDA kernel when no autograd graph is
    needed; plain torch ops (so autograd works) when training."""
    needs_grad = torch.is_grad_enabled() and (hidden.requires_grad or (residual is not None and residual.requires_grad)
                                              or norm.weight.requires_grad)
    if not hidden.is_cuda:
        raise RuntimeError("si-mamba Block runs on CUDA tensors only (there is no CPU fallback)")
    if not isinstance(norm, nn.LayerNorm) or norm.weight is None or norm.bias is None:
        # `rms_norm: True` / affine-free norms: not selected by any shipped config, no kernel - plain torch ops on the device
        res = hidden + residual if residual is not None else hidden
        res = res.float() if res.dtype != torch.float32 else res
        return norm(res.to(dtype=norm.weight.dtype)), res
    if needs_grad:  # training: same forward kernel, backward through sim_add_layernorm_bwd
        return ops.AddLayerNorm.apply(hidden, residual, norm.weight, norm.bias, norm.eps, _amp_dtype(hidden), row_scale)
    if row_scale is not None:  # DropPath without a graph (train mode under no_grad): plain scaling, then the kernel
        hidden = hidden * row_scale.view(-1, *([1] * (hidden.dim() - 1))).to(hidden.dtype)
    out_dtype = _amp_dtype(hidden)
    return ops.add_layernorm(hidden, residual, norm.weight, norm.bias, norm.eps, out_dtype=out_dtype,
                             want_residual=want_residual, split=split if out_dtype == torch.float32 else False)


class Block(nn.Module):
    def __init__(self, dim, mixer_cls, norm_cls=nn.LayerNorm, fused_add_norm=False, residual_in_fp32=False,
                 drop_path=0.):
        super().__init__()
        self.residual_in_fp32 = residual_in_fp32
        self.fused_add_norm = fused_add_norm  # kept for signature parity; the CUDA add+LN is always used in eval
        self.mixer = mixer_cls(dim)
        self.norm = norm_cls(dim)
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()

    def forward(self, hidden_states: Tensor, residual: Optional[Tensor] = None, inference_params=None):
        """residual = drop_path(hidden) + residual; hidden = mixer(LN(residual)).  block.py:47-73."""
        # fp32 inference: LayerNorm writes the in_proj operand (three bf16 planes) directly, see autograd.wants_split3
        want = getattr(self.mixer, "wants_split3", None)
        split = bool(want and want(hidden_states))
        if split and isinstance(self.norm, nn.LayerNorm) and self.norm.bias is not None:
            from .autograd import inproj_f16_ok
            if inproj_f16_ok(self.norm.weight, self.norm.bias, self.mixer.in_proj.weight):
                split = "f16x2"  # bounded operand: two fp16 planes, half the tensor-core work of in_proj
        # block.py:59: `drop_path(h) + residual if residual is not None else h` - the first block's input is never dropped.
        # The per-sample factor mask / keep goes into the add + LayerNorm kernels (forward and backward) as row_scale.
        scale = None
        h = hidden_states
        if residual is not None and isinstance(self.drop_path, DropPath) and self.drop_path.drop_prob > 0. and self.training:
            if hidden_states.is_cuda and hidden_states.dim() == 3 and isinstance(self.norm, nn.LayerNorm):
                scale = self.drop_path.sample_scale(hidden_states)
            else:
                h = self.drop_path(hidden_states)
        hidden_states, residual = fused_add_norm(self.norm, h, residual, split=split, row_scale=scale)
        hidden_states = self.mixer(hidden_states, inference_params=inference_params)
        return hidden_states, residual

    def allocate_inference_cache(self, batch_size, max_seqlen, dtype=None, **kwargs):
        return self.mixer.allocate_inference_cache(batch_size, max_seqlen, dtype=dtype, **kwargs)
