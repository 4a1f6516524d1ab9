"""Index maps of the token layouts built around the spectral permutation (tiny integer glue on the device; the row
movement itself is done by the sim_gather_rows / sim_order_gather kernels).

  * MAE visible-token compaction and token restore   models/point_mamba.py:2734-2796, 3147-3197
  * HLT chunked forward / reverse layout              part_segmentation/models/pt_mamba.py:670-723
"""

from __future__ import annotations

import torch

from . import ops


def mae_index_maps(perm: torch.Tensor, mask: torch.Tensor, n_vis: int = None, check: bool = False):
    """Index maps of the MAE layout from ONE kernel (ops.mae_index_maps / sim_mae_index_maps); see
    mae_index_maps_torch for the meaning of every map (kept as the in-tree description and for CPU use).
    ``n_vis`` = visible patches per cloud (G - int(mask_ratio * G)); None counts them from the mask (host sync)."""
    if not perm.is_cuda:
        return mae_index_maps_torch(perm, mask)
    if n_vis is None:
        n_vis = int((~mask[0]).sum())
        check = True
    return ops.mae_index_maps(perm, mask, n_vis, check=check)


def mae_index_maps_torch(perm: torch.Tensor, mask: torch.Tensor):
    """perm (B,k,G) int, mask (B,G) bool with the same number of masked patches in every cloud ->
    dict(src_vis (B, 2k*n_vis) patch feeding each encoder token,
         restore_src (B, 2kG) row of the encoder output for each decoder position, -1 = mask token,
         mask_full (B, 2kG) bool,
         rec_src (B, 2k*m) decoder positions that are reconstructed, ascending,
         perm_full (B, 2kG) patch behind every decoder position)."""
    B, k, G = perm.shape
    flat = perm.reshape(B, k * G).long()
    perm_full = torch.cat((flat, flat.flip(1)), dim=1)
    m_sorted = torch.gather(mask, 1, flat)
    mask_full = torch.cat((m_sorted, m_sorted.flip(1)), dim=1)
    n_vis_total = int((~mask_full[0]).sum())
    order = torch.sort(mask_full.to(torch.int8), dim=1, stable=True).indices  # visible positions first, in order
    vis_pos, msk_pos = order[:, :n_vis_total], order[:, n_vis_total:]
    src_vis = torch.gather(perm_full, 1, vis_pos)
    rank = torch.cumsum((~mask_full).to(torch.int32), dim=1) - 1
    restore_src = torch.where(mask_full, torch.full_like(rank, -1), rank)
    return dict(src_vis=src_vis.int(), restore_src=restore_src.int(), mask_full=mask_full, rec_src=msk_pos.int(),
                perm_full=perm_full.int())


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
    """Differentiable sim_gather_rows: backward is the scatter-add of the same map (plus the fill-row reduction)."""

    @staticmethod
    def forward(ctx, x, src_idx, fill):
        ctx.save_for_backward(src_idx)
        ctx.r_in = x.shape[1]
        ctx.has_fill = fill is not None
        return ops.gather_rows(x, src_idx, fill)

    @staticmethod
    def backward(ctx, dout):
        (src_idx,) = ctx.saved_tensors
        B, R_out, C = dout.shape
        valid = (src_idx >= 0)
        idx = src_idx.clamp(min=0).long()[..., None].expand(-1, -1, C)
        dx = torch.zeros(B, ctx.r_in, C, dtype=dout.dtype, device=dout.device)
        dx.scatter_add_(1, idx, dout * valid[..., None].to(dout.dtype))
        dfill = (dout * (~valid)[..., None].to(dout.dtype)).sum(dim=(0, 1)) if ctx.has_fill else None
        return dx, None, dfill


def gather_rows(x: torch.Tensor, src_idx: torch.Tensor, fill: torch.Tensor = None) -> torch.Tensor:
    """out[b,t] = x[b, src_idx[b,t]] (src >= 0) else fill / zeros; differentiable w.r.t. x and fill."""
    src_idx = src_idx.to(torch.int32).contiguous()
    if torch.is_grad_enabled() and (x.requires_grad or (fill is not None and fill.requires_grad)):
        return _GatherRows.apply(x.contiguous(), src_idx, fill)
    return ops.gather_rows(x, src_idx, fill)
