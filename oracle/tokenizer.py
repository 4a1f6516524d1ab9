"""Oracle: FPS + kNN patch grouping (test infrastructure, see oracle/__init__.py).

Restates pytorch3d ``sample_farthest_points`` / ``knn_points`` as the reference
calls them in ``Group.forward`` (models/point_mamba.py:83-111; the seg twin is
part_segmentation/models/pt_mamba.py:165-191).  pytorch3d is not vendored in
the reference tree; its published semantics are:

  FPS : first index 0; running min of squared L2 to the selected set;
        next = first arg-max of the running min.
  kNN : the K smallest squared-L2 distances per query, order unspecified.

Arithmetic contract (shared with the CUDA kernels): squared distance is
``((dx*dx) + (dy*dy)) + (dz*dz)`` in fp32 with every product and sum rounded
separately (no FMA contraction); ties break towards the lower point index.
"""

from __future__ import annotations

import torch


def sqdist(a: torch.Tensor, b: torch.Tensor, fma: bool = False) -> torch.Tensor:
    """Squared L2 between broadcastable (...,3) fp32 tensors, fixed op order.  ``fma=False``: no FMA (the contract).
    ``fma=True``: fma(dz, dz, fma(dy, dy, dx*dx)) - what nvcc's default contraction makes of pytorch3d's
    ``dist2 += diff * diff`` device loops; each fused step is evaluated in fp64 (the product of two fp32 values is exact
    there) and rounded to fp32 once."""
    dx = a[..., 0] - b[..., 0]
    dy = a[..., 1] - b[..., 1]
    dz = a[..., 2] - b[..., 2]
    if fma:
        t = (dx * dx).double()
        t = (dy.double() * dy.double() + t).float().double()
        return (dz.double() * dz.double() + t).float()
    # separate torch ops => each product / sum is rounded to fp32 on its own
    return (dx * dx + dy * dy) + dz * dz


def fps(xyz: torch.Tensor, num_group: int, fma: bool = False) -> torch.Tensor:
    """Farthest point sampling, pytorch3d semantics (call site point_mamba.py:93).

    xyz (B,N,3) fp32 -> idx (B,G) int64.  Start index 0, min-dist init +inf.
    """
    assert xyz.dtype == torch.float32 and xyz.dim() == 3
    B, N, _ = xyz.shape
    idx = torch.zeros(B, num_group, dtype=torch.int64)
    min_d = torch.full((B, N), float("inf"), dtype=torch.float32)
    last = torch.zeros(B, dtype=torch.int64)
    ar = torch.arange(B)
    for j in range(1, num_group):
        p = xyz[ar, last]  # (B,3)
        d = sqdist(xyz, p[:, None, :], fma)
        min_d = torch.minimum(min_d, d)
        # first arg-max: torch.max over dim returns the first maximal index on CPU
        mx = min_d.max(dim=1, keepdim=True).values
        is_max = min_d == mx
        last = torch.argmax(is_max.to(torch.uint8), dim=1)  # first True
        idx[:, j] = last
    return idx


def knn_group(xyz: torch.Tensor, center: torch.Tensor, group_size: int, fma: bool = False):
    """kNN grouping (point_mamba.py:96-110).

    Returns (idx sorted ascending by point index (B,G,M) int64,
             neighborhood centred (B,G,M,3), neighborhood_org (B,G,M,3)).
    The reference's neighbour order is unspecified (return_sorted=False); the
    contract here is the SET of the M smallest (distance, index) pairs, emitted
    in ascending point-index order.
    """
    B, N, _ = xyz.shape
    G = center.shape[1]
    d = sqdist(center[:, :, None, :], xyz[:, None, :, :], fma)  # (B,G,N)
    # lexicographic (distance, index): stable sort on distance keeps index order
    order = torch.sort(d, dim=-1, stable=True).indices[..., :group_size]
    idx = torch.sort(order, dim=-1).values
    flat = (idx + (torch.arange(B).view(B, 1, 1) * N)).reshape(-1)
    org = xyz.reshape(B * N, 3)[flat].reshape(B, G, group_size, 3)
    nbr = org - center[:, :, None, :]
    return idx, nbr, org


def group(xyz: torch.Tensor, num_group: int, group_size: int):
    """``Group.forward`` (point_mamba.py:83-111): (neighborhood, center, neighborhood_org, fps_idx, knn_idx)."""
    fidx = fps(xyz, num_group)
    B = xyz.shape[0]
    center = xyz[torch.arange(B)[:, None], fidx]
    kidx, nbr, org = knn_group(xyz, center, group_size)
    return nbr, center, org, fidx, kidx


def synthetic_clouds(B: int, N: int, seed: int, kind: str = "ball") -> torch.Tensor:
    """Synthetic clouds of SURVEY.md section 8(d), normalised like the datasets
    (datasets/ModelNetDataset.py:52-57: centroid 0, max-norm 1)."""
    g = torch.Generator().manual_seed(seed)
    if kind == "ball":
        v = torch.randn(B, N, 3, generator=g)
        v = v / v.norm(dim=-1, keepdim=True)
        r = torch.rand(B, N, 1, generator=g) ** (1.0 / 3.0)
        pts = v * r
    elif kind in ("surface", "duplicates"):
        pts = torch.empty(B, N, 3)
        for b in range(B):
            n_patch = int(torch.randint(3, 7, (1,), generator=g))
            which = torch.randint(0, n_patch, (N,), generator=g)
            ctr = torch.randn(n_patch, 3, generator=g) * 0.5
            axes = torch.rand(n_patch, 3, generator=g) * 0.6 + 0.1
            v = torch.randn(N, 3, generator=g)
            v = v / v.norm(dim=-1, keepdim=True)
            p = ctr[which] + v * axes[which]
            pts[b] = p + 0.01 * torch.randn(N, 3, generator=g)
        if kind == "duplicates":
            # part_segmentation/dataset.py:155 samples WITH replacement
            sel = torch.randint(0, N, (B, N), generator=g)
            pts = torch.gather(pts, 1, sel[..., None].expand(-1, -1, 3))
    else:
        raise ValueError(kind)
    pts = pts - pts.mean(dim=1, keepdim=True)
    m = pts.norm(dim=-1).max(dim=1).values
    pts = pts / m[:, None, None]
    return pts.contiguous().float()


def fps_pointnet2(xyz: torch.Tensor, npoint: int) -> torch.Tensor:
    """pointnet2_ops ``furthest_point_sample`` (un-vendored git dependency, README.md:49; call sites utils/misc.py:19,
    tools/runner_finetune.py:191-193) restated from its published CUDA kernel: idx[0] = 0; temp = 1e10; per step, for
    every point k with |p_k|^2 > 1e-3: d = dx*dx + dy*dy + dz*dz (FMA-contracted, emulated in fp64 -> fp32),
    temp[k] = min(temp[k], d); the next index is the arg-max of temp over the visited points, where each of the
    bs = min(512, 2^floor(log2 N)) upstream threads keeps its first strict maximum over k = tid, tid + bs, ... and the
    tree reduction keeps the lower thread on ties.  xyz (B,N,3) fp32 -> (B,npoint) int64.  Pure loops: small cases only."""
    import numpy as np
    B, N, _ = xyz.shape
    bs = 1
    while bs * 2 <= N and bs < 512:
        bs *= 2
    out = torch.zeros(B, npoint, dtype=torch.int64)
    f32 = np.float32
    for b in range(B):
        p = xyz[b].numpy().astype(np.float64)
        fma = lambda a, c: (a * a + c).astype(f32).astype(np.float64)  # round(a*a + c) to fp32
        mag = fma(p[:, 2], fma(p[:, 1], (p[:, 0] * p[:, 0]).astype(f32).astype(np.float64)))
        ok = mag > 1e-3
        temp = np.full(N, 1e10, dtype=np.float64)
        key = (np.arange(N) % bs) * 64 + np.arange(N) // bs
        old = 0
        for j in range(1, npoint):
            d = p - p[old]
            dist = fma(d[:, 2], fma(d[:, 1], (d[:, 0] * d[:, 0]).astype(f32).astype(np.float64)))
            temp = np.where(ok, np.minimum(temp, dist), temp)
            cand = np.where(ok, temp, -1.0)
            best = cand.max()
            if best < 0:
                old = 0
            else:
                ties = np.nonzero(cand == best)[0]
                old = int(ties[np.argmin(key[ties])])
            out[b, j] = old
    return out
