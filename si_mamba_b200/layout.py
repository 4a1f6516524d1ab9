"""Index maps of the token layouts built around the spectral permutation (tiny integer glue on the device; the row
movement itself is done by the sim_gather_rows / sim_order_gather kernels).

  * MAE visible-token compaction and token restore   models/point_mamba.py:2734-2796, 3147-3197
  * HLT chunked forward / reverse layout              part_segmentation/models/pt_mamba.py:670-723
"""

from __future__ import annotations

import torch

from . import ops


def mae_index_maps(perm: torch.Tensor, mask: torch.Tensor, n_vis: int = None, check: bool = False):
    """Index maps of the MAE layout from ONE kernel (ops.mae_index_maps / sim_mae_index_maps):
    perm (B,k,G) int, mask (B,G) bool with the same number of masked patches in every cloud ->
    dict(src_vis (B, 2k*n_vis) patch feeding each encoder token,
         restore_src (B, 2kG) row of the encoder output for each decoder position, -1 = mask token,
         mask_full (B, 2kG) bool,
         rec_src (B, 2k*m) decoder positions that are reconstructed, ascending,
         perm_full (B, 2kG) patch behind every decoder position, vis_pos / inv_vis the inverse maps).
    The torch restatement the kernel is tested against lives in oracle/mae.py (mae_index_maps_torch).
    ``n_vis`` = visible patches per cloud (G - int(mask_ratio * G)); None counts them from the mask (host sync)."""
    if not perm.is_cuda:
        raise RuntimeError("mae_index_maps runs on CUDA tensors only (there is no CPU fallback)")
    if n_vis is None:
        n_vis = int((~mask[0]).sum())
        check = True
    return ops.mae_index_maps(perm, mask, n_vis, check=check)


_HLT_SLOTS = {}


def hlt_slots(G: int, k: int, reverse: bool, device) -> torch.Tensor:
    """Cached per (G, k, reverse, device): a constant of the layout (also keeps the forward CUDA-graph capturable)."""
    key = (G, k, bool(reverse), str(device))
    if key not in _HLT_SLOTS:
        _HLT_SLOTS[key] = _hlt_slots(G, k, reverse, device)
    return _HLT_SLOTS[key]


def _hlt_slots(G: int, k: int, reverse: bool, device) -> torch.Tensor:
    """Rank (in the bucket-sorted order) shown at each of the 2G output slots, -1 = zero token.  The reference loop
    writes chunk i (c = 2^k tokens) at [(i+1)c, (i+2)c) for i >= 1 - over the previous chunk's reverse - and its
    reverse right after, so the net layout is [F0, R0, F1, F2, ..., F_last, R_last, zeros]."""
    c = 2 ** k
    slots = torch.full((2 * G,), -1, dtype=torch.int64)
    if reverse:
        for i in range(G // c):
            fwd = torch.arange(i * c, (i + 1) * c)
            base = 0 if i == 0 else (i + 1) * c
            slots[base:base + c] = fwd
            slots[base + c:base + 2 * c] = fwd.flip(0)
    return slots.to(device)


def hlt_src_index(order: torch.Tensor, k: int, reverse: bool = True) -> torch.Tensor:
    """order (B,G): argsort of the bucket keys -> src (B, 2G) int32 patch index per output slot (-1 = zero token)."""
    B, G = order.shape
    slots = hlt_slots(G, k, reverse, order.device)
    src = torch.where(slots >= 0, order.long()[:, slots.clamp(min=0)], torch.full((1,), -1, device=order.device))
    return src.int()


class _GatherRows(torch.autograd.Function):
    """Differentiable sim_gather_rows: the backward is a deterministic gather over the inverse map
    (sim_invert_row_map + sim_gather_sum_rows), no atomics."""

    @staticmethod
    def forward(ctx, x, src_idx, fanout):
        ctx.save_for_backward(src_idx)
        ctx.r_in, ctx.fanout = x.shape[1], fanout
        return ops.gather_rows(x, src_idx, None)

    @staticmethod
    def backward(ctx, dout):
        (src_idx,) = ctx.saved_tensors
        inv = ops.invert_row_map(src_idx, ctx.r_in, ctx.fanout)
        return ops.gather_sum_rows(dout.contiguous(), inv), None, None


def gather_rows(x: torch.Tensor, src_idx: torch.Tensor, fill: torch.Tensor = None, fanout: int = None) -> torch.Tensor:
    """out[b,t] = x[b, src_idx[b,t]] (src >= 0) else fill / zeros; differentiable w.r.t. x.  ``fanout`` = the largest
    number of output rows that read one source row (2 for the HLT layout, 2k for the MAE position gather); None counts
    it from the map (a host sync - pass it when the step is captured in a CUDA graph)."""
    src_idx = src_idx.to(torch.int32).contiguous()
    if torch.is_grad_enabled() and fill is not None and fill.requires_grad:
        raise NotImplementedError("gather_rows: a trainable fill row is MaeRestore's job (ops.MaeRestore)")
    if torch.is_grad_enabled() and x.requires_grad:
        if fill is not None:
            raise NotImplementedError("gather_rows: differentiable gather with a fill row is ops.MaeRestore")
        if fanout is None:
            flat = (src_idx.long() + torch.arange(src_idx.shape[0], device=src_idx.device)[:, None] * x.shape[1])
            fanout = max(1, int(torch.bincount(flat[src_idx >= 0].reshape(-1), minlength=1).max()))
        return _GatherRows.apply(x.contiguous(), src_idx, int(fanout))
    return ops.gather_rows(x, src_idx, fill)
