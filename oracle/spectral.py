"""Oracle: patch graph -> Laplacian -> eigenvectors -> spectral orderings.

Test infrastructure (see oracle/__init__.py).  Follows models/point_mamba.py:
  * create_graph_from_feature_space_gpu_weighted_adjacency  :664-715
  * create_graph_from_centers                               :620-661, :2958-2999
  * calc_top_k_eigenvalues_eigenvectors (per-cloud loop)    :717-761
  * calc_top_k_eigenvalues_eigenvectors (batched, MAE)      :3001-3050
  * calc_top_k_eigenvalues_eigenvectors_symmetric           :764-814
  * sort_points_by_fiedler                                  :817-826
  * multilevel_travers                                      :829-841
  * SAST order assembly + reverse                           :889-898, :982-989
  * HLT layout   part_segmentation/models/pt_mamba.py:670-723
  * sign rule    work_order.py:360-365
"""

from __future__ import annotations

import torch

from .tokenizer import sqdist


# --------------------------------------------------------------------------- graph
def pairwise_dist(points: torch.Tensor) -> torch.Tensor:
    """sqrt of the fixed-order squared distance (point_mamba.py:682): (B,G,3)->(B,G,G)."""
    return torch.sqrt(sqdist(points[:, :, None, :], points[:, None, :, :]))


def knn_adjacency(points, k, alpha, symmetric, self_loop, binary, sigma_mode=False):
    """Adjacency of the patch graph (point_mamba.py:664-715 / :620-661).

    topk(-dist, k+1) with (distance, index) lexicographic ties; column 0 (self)
    dropped unless ``self_loop``; weight 1 or exp(-alpha * d^2) with d the
    sqrt-ed distance squared again (:702).  ``sigma_mode`` is the alpha == 0
    branch of create_graph_from_centers (:647-648): exp(-d^2 / (2 sigma^2)),
    sigma = mean of the whole batch distance tensor.
    """
    B, G, _ = points.shape
    dist = pairwise_dist(points)
    order = torch.sort(dist, dim=-1, stable=True).indices[..., : k + 1]
    dsel = torch.gather(dist, -1, order)
    if not self_loop:
        order, dsel = order[..., 1:], dsel[..., 1:]
    if sigma_mode:
        sigma = dist.mean()
        w = torch.exp(-dsel ** 2 / (2 * sigma ** 2))
    else:
        w = torch.exp((-1) * alpha * dsel ** 2)
    if binary:
        w = torch.ones_like(w)
    A = torch.zeros(B, G, G, dtype=torch.float32)
    b_idx = torch.arange(B)[:, None, None]
    n_idx = torch.arange(G)[None, :, None]
    A[b_idx, n_idx, order] = w
    if symmetric:
        A[b_idx, order, n_idx] = w
    return A


# ----------------------------------------------------------------------- laplacian
def laplacian_operator(A: torch.Tensor, matrix: str = "laplacian", eps_mode: str = "add1e-6") -> torch.Tensor:
    """The symmetric operator ``torch.linalg.eigh`` actually sees.

    fp32 arithmetic exactly as the reference: A=(A+A^T)/2; deg=sum_j A;
      "add1e-6"    : L = I - diag(1/(deg+1e-6)) A      (point_mamba.py:731-740)
      "clamp1e-12" : L = I - A / clamp(deg, 1e-12)     (point_mamba.py:3023-3031)
      matrix != "laplacian": L = I - D^-1/2 A D^-1/2   (:778-792)
    eigh reads the LOWER triangle only (UPLO='L'), so the operator is
    S = tril(L) + tril(L,-1)^T  (SURVEY.md section 7-2).
    """
    A = (A + A.transpose(-1, -2)) / 2
    deg = A.sum(dim=-1)
    G = A.shape[-1]
    eye = torch.eye(G, dtype=A.dtype)
    if matrix == "laplacian":
        if eps_mode == "add1e-6":
            r = 1.0 / (deg + 1e-6)
            L = eye - r[..., :, None] * A
        elif eps_mode == "clamp1e-12":
            L = eye - A / deg.clamp(min=1e-12)[..., :, None]
        else:
            raise ValueError(eps_mode)
    else:
        dis = torch.pow(deg, -0.5)
        L = eye - (dis[..., :, None] * A) * dis[..., None, :]
    low = torch.tril(L)
    return low + torch.tril(L, -1).transpose(-1, -2)


def canonical_sign(vecs: torch.Tensor) -> torch.Tensor:
    """Sign rule of work_order.py:360-365 (first entry non-negative), with the
    SURVEY A.5 fallback: when |v[0]| < 1e-6 use the sign of the entry of largest
    magnitude (lowest index on ties).  vecs (B,G,k)."""
    v0 = vecs[:, 0, :]
    big = vecs.abs().argmax(dim=1)  # first max index
    vbig = torch.gather(vecs, 1, big[:, None, :])[:, 0, :]
    ref = torch.where(v0.abs() < 1e-6, vbig, v0)
    s = torch.where(ref < 0, -torch.ones_like(ref), torch.ones_like(ref))
    return vecs * s[:, None, :]


def topk_eigen(S: torch.Tensor, k: int, smallest: bool, drop_first: bool = False, dtype=torch.float64):
    """k extremal eigenpairs of the symmetric operator S (B,G,G) via LAPACK in
    ``dtype`` (fp64 = the eigen-oracle of SURVEY 7-1).  Returns (vals (B,k),
    vecs (B,G,k) sign-canonicalised, all_vals (B,G))."""
    w, V = torch.linalg.eigh(S.to(dtype))
    kk = k + 1 if drop_first else k
    if smallest:
        sel = torch.arange(kk)
    else:
        sel = torch.arange(S.shape[-1] - 1, S.shape[-1] - 1 - kk, -1)
    vals, vecs = w[:, sel], V[:, :, sel]
    if drop_first:
        vals, vecs = vals[:, 1:], vecs[:, :, 1:]
    return vals, canonical_sign(vecs), w


def spectral_eig(center, k_nn, alpha, symmetric, self_loop, binary, k, smallest,
                 matrix="laplacian", eps_mode="add1e-6", sigma_mode=False, dtype=torch.float64):
    """centres -> (vals, vecs, all_vals, S): the whole a-3 + a-4 chain."""
    A = knn_adjacency(center, k_nn, alpha, symmetric, self_loop, binary, sigma_mode)
    S = laplacian_operator(A, matrix, eps_mode)
    vals, vecs, allv = topk_eigen(S, k, smallest, drop_first=(matrix != "laplacian"), dtype=dtype)
    return vals, vecs, allv, S


# ------------------------------------------------------------------------ ordering
def argsort_stable(keys: torch.Tensor) -> torch.Tensor:
    """Ascending argsort along dim 1 with lower-index-first ties (torch.sort at
    point_mamba.py:820 is not stable-flagged; the contract fixes the tie rule)."""
    return torch.sort(keys, dim=1, stable=True).indices


def sast_perm(vecs: torch.Tensor) -> torch.Tensor:
    """vecs (B,G,k) -> perm (B,k,G): perm[b,s,r] = index of the r-th smallest entry of eigenvector s."""
    return argsort_stable(vecs).transpose(1, 2).contiguous()


def order_gather(x: torch.Tensor, perm: torch.Tensor, reverse: bool = True) -> torch.Tensor:
    """SAST assembly (point_mamba.py:889-898, 982-989): cat_s x[perm[s]] then cat(seq, flip(seq)).

    x (B,G,C), perm (B,k,G) -> (B, 2kG or kG, C)."""
    B, k, G = perm.shape
    flat = perm.reshape(B, k * G)
    seq = torch.gather(x, 1, flat[..., None].expand(-1, -1, x.shape[-1]))
    if reverse:
        seq = torch.cat((seq, seq.flip(1)), dim=1)
    return seq


def multilevel_travers(vecs: torch.Tensor, level: int) -> torch.Tensor:
    """point_mamba.py:829-841: sign bits vs per-vector mean -> bucket id (B,G) int64."""
    means = vecs.mean(dim=1, keepdim=True)
    bits = (vecs >= means)[:, :, :level]
    pw = 2 ** torch.arange(level - 1, -1, -1)
    return (bits * pw[None, None, :]).sum(dim=-1)


def hlt_slots(G: int, k: int, reverse: bool = True) -> torch.Tensor:
    """Source rank for each of the 2G output slots of the HLT layout
    (pt_mamba.py:696-723), -1 = zero token.  Chunk c = 2^k; for i = 0 the chunk
    and its reverse go to [0,c) and [c,2c); for i >= 1 chunk i is written at
    [(i+1)c,(i+2)c) (overwriting the previous reverse) and its reverse at
    [(i+2)c,(i+3)c)."""
    c = 2 ** k
    nd = G // c
    slots = torch.full((2 * G,), -1, dtype=torch.int64)
    if not reverse:
        return slots  # reference leaves the buffer all-zero when reverse is False
    for i in range(nd):
        fwd = torch.arange(i * c, (i + 1) * c)
        base = 0 if i == 0 else (i + 1) * c
        slots[base:base + c] = fwd
        slots[base + c:base + 2 * c] = fwd.flip(0)
    return slots


def hlt_order(vecs: torch.Tensor, k: int, noise: torch.Tensor) -> torch.Tensor:
    """ids + U[0,1) tie-break noise (CPU RNG in the reference) -> argsort (B,G)."""
    ids = multilevel_travers(vecs, k).to(torch.float32)
    return argsort_stable(ids + noise)


def hlt_layout(x: torch.Tensor, order: torch.Tensor, k: int, reverse: bool = True) -> torch.Tensor:
    """x (B,G,C) -> (B,2G,C) in the HLT layout, zero tokens where slot == -1."""
    B, G, C = x.shape
    slots = hlt_slots(G, k, reverse)
    out = torch.zeros(B, 2 * G, C, dtype=x.dtype)
    live = slots >= 0
    src = order[:, slots[live]]  # (B, n_live)
    out[:, live] = torch.gather(x, 1, src[..., None].expand(-1, -1, C))
    return out
