"""Oracle: MAE mask / masked spectral sort / token restore.

Test infrastructure (see oracle/__init__.py).  Follows models/point_mamba.py:
  * _mask_center_rand           :2232-2255  (m = int(mask_ratio * G) ones, shuffled per cloud)
  * masked sort (MaskMamba_3)   :2734-2796  (per order: sort tokens+mask, keep ~mask rows in order;
                                             cat k orders; cat(seq, flip(seq)))
  * token restore               :3147-3197  (x_full[mask] = mask_token, x_full[~mask] = x_vis rows in order)
"""

from __future__ import annotations

import torch


def mask_full(mask: torch.Tensor, perm: torch.Tensor) -> torch.Tensor:
    """mask (B,G) bool, perm (B,k,G) -> mask_full (B,2kG): cat_s mask[perm[s]] then cat(m, flip(m)) (:2795-2796)."""
    B, k, G = perm.shape
    m = torch.gather(mask, 1, perm.reshape(B, k * G))
    return torch.cat((m, m.flip(1)), dim=1)


def compact_visible(x: torch.Tensor, perm: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """x (B,G,C) -> x_vis (B, 2k*n_vis, C): visible rows of each sorted copy, order preserved, then flip+cat."""
    B, k, G = perm.shape
    C = x.shape[-1]
    seq = torch.gather(x, 1, perm.reshape(B, k * G)[..., None].expand(-1, -1, C))
    m = torch.gather(mask, 1, perm.reshape(B, k * G))
    vis = seq[~m].reshape(B, -1, C)
    return torch.cat((vis, vis.flip(1)), dim=1)


def restore(x_vis: torch.Tensor, mfull: torch.Tensor, mask_token: torch.Tensor) -> torch.Tensor:
    """Token restore (:3147-3197): x_full[b,t] = mask_token if mfull[b,t] else x_vis[b, rank_vis(b,t)]."""
    B, T = mfull.shape
    C = x_vis.shape[-1]
    rank = torch.cumsum((~mfull).to(torch.int64), dim=1) - 1
    rank = rank.clamp(min=0)
    gathered = torch.gather(x_vis, 1, rank[..., None].expand(-1, -1, C))
    return torch.where(mfull[..., None], mask_token.reshape(1, 1, C).to(x_vis.dtype), gathered)


def gather_masked(x_full: torch.Tensor, mfull: torch.Tensor) -> torch.Tensor:
    """x_rec = x_full[mask_full] reshaped (B, 2k*m, C) (:3194-3197)."""
    B = x_full.shape[0]
    return x_full[mfull].reshape(B, -1, x_full.shape[-1])


def rand_mask(B: int, G: int, mask_ratio: float, seed: int) -> torch.Tensor:
    """_mask_center_rand (:2232-2255) with a torch generator instead of numpy's global RNG."""
    g = torch.Generator().manual_seed(seed)
    m = int(mask_ratio * G)
    out = torch.zeros(B, G, dtype=torch.bool)
    for b in range(B):
        p = torch.randperm(G, generator=g)
        out[b, p[:m]] = True
    return out


def chamfer_l2(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """pytorch3d chamfer_distance(x, y, batch_reduction=None)[0], squared L2, mean over points (:2950, :3203)."""
    d = (x[:, :, None, :] - y[:, None, :, :]).pow(2).sum(-1)
    return d.min(dim=2).values.mean(dim=1) + d.min(dim=1).values.mean(dim=1)


def point_mae_forward(sd: dict, cfg: dict, pts: torch.Tensor, mask: torch.Tensor, perm_override=None):
    """Point_MAE_Mamba.forward (models/point_mamba.py:3053-3213), spectral method, eval mode, given the (B,G) mask.

    Spectral ordering uses the batched variant (:3001-3050, deg.clamp(1e-12)).  Returns (loss, intermediates)."""
    from . import mamba, model, spectral, tokenizer
    tc = cfg["transformer_config"]
    k = tc["k_top_eigenvectors"]
    nbr, center, org, fidx, kidx = tokenizer.group(pts, cfg["num_group"], cfg["group_size"])
    vals, vecs, allv, S = spectral.spectral_eig(center, tc["knn_graph"], tc["alpha"], tc["symmetric"], tc["self_loop"],
                                                tc["binary"], k, tc["smallest"], "laplacian", "clamp1e-12")
    perm = spectral.sast_perm(vecs) if perm_override is None else perm_override
    tok = model.encoder(sd, "MAE_encoder.encoder.", nbr)
    pos = model.pos_embed(sd, "MAE_encoder.pos_embed.", center)
    x_vis = compact_visible(tok, perm, mask)
    pos_vis = compact_visible(pos, perm, mask)
    x_vis = mamba.mixer_model(sd, "MAE_encoder.blocks.", x_vis, pos_vis, tc["depth"])
    x_vis = mamba.layer_norm(sd, "MAE_encoder.norm.", x_vis)
    mfull = mask_full(mask, perm)
    x_full = restore(x_vis, mfull, sd["mask_token"].reshape(-1))
    pos_full = spectral.order_gather(pos, perm, True)
    x_rec = mamba.mixer_model(sd, "MAE_decoder.blocks.", x_full, pos_full, tc["decoder_depth"])
    x_rec = mamba.layer_norm(sd, "MAE_decoder.norm.", x_rec)
    x_rec = gather_masked(x_rec, mfull)
    B, M, C = x_rec.shape
    import torch.nn.functional as F
    reb = F.conv1d(x_rec.transpose(1, 2), sd["increase_dim.0.weight"], sd["increase_dim.0.bias"]).transpose(1, 2)
    reb = reb.reshape(B * M, -1, 3)
    Gs = nbr.shape[2]
    nbr_full = spectral.order_gather(nbr.reshape(nbr.shape[0], nbr.shape[1], -1), perm, True)
    gt = nbr_full[mfull].reshape(B * M, Gs, 3)
    loss = chamfer_l2(reb, gt).mean()
    return loss, dict(perm=perm, x_vis=x_vis, x_full=x_full, mask_full=mfull, eigvecs=vecs)


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
