"""Oracle: PointMamba classification forward on CPU from a state dict.

Test infrastructure (see oracle/__init__.py).  Follows models/point_mamba.py:
Encoder.forward :59-73, PointMamba.forward :843-898 + :982-989 + :1115-1130
(method == "SAST", default arguments, eval mode: BatchNorm running stats,
Dropout/DropPath identity).
"""

from __future__ import annotations

import torch
import torch.nn.functional as F

from . import mamba, spectral, tokenizer


def _bn_eval(sd, prefix, x, eps=1e-5):
    return F.batch_norm(x, sd[prefix + "running_mean"], sd[prefix + "running_var"],
                        sd[prefix + "weight"], sd[prefix + "bias"], False, 0.0, eps)


def encoder(sd: dict, prefix: str, point_groups: torch.Tensor) -> torch.Tensor:
    """Encoder.forward (point_mamba.py:59-73): (B,G,M,3) -> (B,G,C)."""
    bs, g, n, _ = point_groups.shape
    x = point_groups.reshape(bs * g, n, 3).transpose(2, 1)
    f = F.conv1d(x, sd[prefix + "first_conv.0.weight"], sd[prefix + "first_conv.0.bias"])
    f = F.relu(_bn_eval(sd, prefix + "first_conv.1.", f))
    f = F.conv1d(f, sd[prefix + "first_conv.3.weight"], sd[prefix + "first_conv.3.bias"])
    fg = torch.max(f, dim=2, keepdim=True)[0]
    f = torch.cat([fg.expand(-1, -1, n), f], dim=1)
    f = F.conv1d(f, sd[prefix + "second_conv.0.weight"], sd[prefix + "second_conv.0.bias"])
    f = F.relu(_bn_eval(sd, prefix + "second_conv.1.", f))
    f = F.conv1d(f, sd[prefix + "second_conv.3.weight"], sd[prefix + "second_conv.3.bias"])
    fg = torch.max(f, dim=2, keepdim=False)[0]
    return fg.reshape(bs, g, -1)


def pos_embed(sd, prefix, center):
    h = F.gelu(F.linear(center, sd[prefix + "0.weight"], sd[prefix + "0.bias"]))
    return F.linear(h, sd[prefix + "2.weight"], sd[prefix + "2.bias"])


def cls_head(sd, prefix, x):
    h = F.linear(x, sd[prefix + "0.weight"], sd[prefix + "0.bias"])
    h = F.relu(_bn_eval(sd, prefix + "1.", h))
    h = F.linear(h, sd[prefix + "4.weight"], sd[prefix + "4.bias"])
    h = F.relu(_bn_eval(sd, prefix + "5.", h))
    return F.linear(h, sd[prefix + "8.weight"], sd[prefix + "8.bias"])


def point_mamba_forward(sd: dict, cfg: dict, pts: torch.Tensor, return_intermediates: bool = False,
                        perm_override: torch.Tensor = None):
    """PointMamba.forward(pts) -> logits (B, cls_dim), SAST path.

    ``cfg`` uses the reference's config keys (cfgs/finetune_modelnet.yaml:23-50).
    """
    sd = {k: v.float() if v.is_floating_point() else v for k, v in sd.items()}
    nbr, center, org, fidx, kidx = tokenizer.group(pts, cfg["num_group"], cfg["group_size"])
    tok = encoder(sd, "encoder.", nbr)
    pos = pos_embed(sd, "pos_embed.", center)
    assert cfg["method"] == "SAST"
    vals, vecs, allv, S = spectral.spectral_eig(
        center, cfg["knn_graph"], cfg["alpha"], cfg["symmetric"], cfg["self_loop"], cfg["binary"],
        cfg["k_top_eigenvectors"], cfg["smallest"], cfg.get("matrix", "laplacian"))
    perm = spectral.sast_perm(vecs)
    perm_oracle = perm
    if perm_override is not None:
        # near-tied eigenvector entries admit several valid orderings; a caller that has verified
        # its permutation against `vecs` can ask for the downstream result under that ordering
        perm = perm_override
    x = spectral.order_gather(tok, perm, cfg["reverse"])
    p = spectral.order_gather(pos, perm, cfg["reverse"])
    h = mamba.mixer_model(sd, "blocks.", x, p, cfg["depth"])
    h = mamba.layer_norm(sd, "norm.", h)
    feat = h.mean(1)
    logits = cls_head(sd, "cls_head_finetune.", feat)
    if return_intermediates:
        return logits, dict(fps_idx=fidx, knn_idx=kidx, center=center, tokens=tok, pos=pos,
                            eigvals=vals, eigvecs=vecs, perm=perm_oracle, hidden=h, feat=feat)
    return logits
