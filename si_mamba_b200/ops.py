"""Torch-tensor front end of the C ABI (include/simamba.h).

PyTorch is plumbing here: it owns device memory and streams; every op below
enqueues hand-written sm_100a kernels from libsimamba_b200.so on the current
CUDA stream.  No CPU path, no eager-PyTorch fallback: a CPU tensor or a missing
library raises.

Function names follow the packages they replace on the reference's hot path:
``selective_scan_fn`` (mamba-ssm), ``causal_conv1d_fn`` (causal-conv1d),
``sample_farthest_points`` / ``knn_points`` (pytorch3d).
"""

from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from ._lib import SIM_BF16, SIM_F32


# ----------------------------------------------------------------------------- helpers
def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return SIM_F32
    if t.dtype == torch.bfloat16:
        return SIM_BF16
    raise TypeError(f"si-mamba kernels take float32 or bfloat16 tensors, got {t.dtype}")


def _cuda(*ts):
    """Every tensor must live on the CURRENT CUDA device: kernels launch on torch.cuda.current_stream() of that device
    (a model on cuda:1 under current device cuda:0 would otherwise run foreign pointers on the wrong context; under
    nn.DataParallel / DDP each replica thread already runs with its own device current)."""
    cur = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("si-mamba ops run on CUDA tensors only (there is no CPU fallback)")
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            raise RuntimeError(f"si-mamba ops launch on the current CUDA device (cuda:{cur}) but got a tensor on {t.device}: "
                               "wrap the call in torch.cuda.device(tensor.device)")


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _tm(t: torch.Tensor) -> int:
    """Row stride (elements) of a token-major (B, L, D) tensor whose rows are uniformly strided."""
    assert t.dim() == 3 and (t.stride(2) == 1 or t.shape[2] == 1), "token-major tensors need unit channel stride"
    B, L, D = t.shape
    if L == 1:  # size-1 dims carry arbitrary strides: rows are then indexed by the batch stride alone
        return t.stride(0) if B > 1 else D
    assert B == 1 or t.stride(0) == L * t.stride(1), "batch stride must equal L * row stride"
    return t.stride(1)


def _f32c(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    return None if t is None else t.detach().float().contiguous()


# ----------------------------------------------------------------------------- tokenizer (a-1)
def sample_farthest_points(points: torch.Tensor, K: int):
    """pytorch3d-compatible subset: (B,N,3) fp32 -> (centres (B,K,3), idx (B,K) int64)."""
    center, idx = fps(points, K)
    return center, idx.long()


# Distance arithmetic of FPS / kNN (include/simamba.h, sim_distance_flags): False = separately rounded products and sums
# (the documented contract), True = the FMA contraction of pytorch3d's device loops.  A process-wide default so a
# maintainer can flip the whole model with one line; every call can override it.
DIST_FMA = __import__("os").environ.get("SIM_DIST_FMA", "0") == "1"


def fps(xyz: torch.Tensor, num_group: int, fma: Optional[bool] = None):
    """-> (center (B,G,3) fp32, idx (B,G) int32).  models/point_mamba.py:93."""
    _cuda(xyz)
    xyz = xyz.float().contiguous()
    B, N, _ = xyz.shape
    idx = torch.empty(B, num_group, dtype=torch.int32, device=xyz.device)
    center = torch.empty(B, num_group, 3, dtype=torch.float32, device=xyz.device)
    if DIST_FMA if fma is None else fma:
        _lib.call("sim_fps_ex", _p(xyz), B, N, num_group, _p(idx), _p(center), _lib.DIST_FMA, _stream())
    else:
        _lib.call("sim_fps", _p(xyz), B, N, num_group, _p(idx), _p(center), _stream())
    return center, idx


def fps_pointnet2(data: torch.Tensor, number: int, return_idx: bool = False):
    """utils/misc.py:14-21 ``fps(data, number)``: pointnet2_ops furthest_point_sample + gather_operation, data (B,N,3)
    -> (B,number,3) (and the int32 indices).  One kernel (sim_fps_pointnet2)."""
    _cuda(data)
    x = data.float().contiguous()
    B, N, _ = x.shape
    idx = torch.empty(B, number, dtype=torch.int32, device=x.device)
    out = torch.empty(B, number, 3, dtype=torch.float32, device=x.device)
    _lib.call("sim_fps_pointnet2", _p(x), B, N, number, _p(idx), _p(out), _stream())
    return (out, idx) if return_idx else out


def knn_group(xyz: torch.Tensor, center: torch.Tensor, group_size: int, want_org: bool = True,
              fma: Optional[bool] = None):
    """-> (idx (B,G,M) int32 ascending, neighborhood centred, neighborhood_org).  point_mamba.py:96-110."""
    _cuda(xyz, center)
    xyz = xyz.float().contiguous()
    center = center.float().contiguous()
    B, N, _ = xyz.shape
    G = center.shape[1]
    idx = torch.empty(B, G, group_size, dtype=torch.int32, device=xyz.device)
    nbr = torch.empty(B, G, group_size, 3, dtype=torch.float32, device=xyz.device)
    org = torch.empty_like(nbr) if want_org else None
    if DIST_FMA if fma is None else fma:
        _lib.call("sim_knn_group_ex", _p(xyz), _p(center), B, N, G, group_size, _p(idx), _p(nbr), _p(org), _lib.DIST_FMA,
                  _stream())
    else:
        _lib.call("sim_knn_group", _p(xyz), _p(center), B, N, G, group_size, _p(idx), _p(nbr), _p(org), _stream())
    return idx, nbr, org


# ----------------------------------------------------------------------------- spectral (a-3..a-5)
def spectral_flags(symmetric, self_loop, binary, smallest, matrix="laplacian", eps_mode="add1e-6",
                   canonical_sign=True) -> int:
    f = 0
    f |= _lib.GRAPH_SYMMETRIC if symmetric else 0
    f |= _lib.GRAPH_SELF_LOOP if self_loop else 0
    f |= _lib.GRAPH_BINARY if binary else 0
    f |= _lib.EIG_SMALLEST if smallest else 0
    f |= _lib.LAP_SYMMETRIC if matrix != "laplacian" else 0
    f |= _lib.LAP_EPS_CLAMP if eps_mode == "clamp1e-12" else 0
    f |= _lib.EIG_CANONICAL_SIGN if canonical_sign else 0
    return f


def pairwise_dist_mean(center: torch.Tensor) -> torch.Tensor:
    """torch.mean of all pairwise centre distances of the batch (create_graph_from_centers, point_mamba.py:626-628):
    centres (B,G,3) -> device scalar (1,) fp32."""
    _cuda(center)
    center = center.detach().float().contiguous()
    B, G, _ = center.shape
    partial = torch.empty(B, dtype=torch.float64, device=center.device)
    sigma = torch.empty(1, dtype=torch.float32, device=center.device)
    _lib.call("sim_pairwise_dist_mean", _p(center), B, G, _p(partial), _p(sigma), _stream())
    return sigma


def spectral_eig(center: Optional[torch.Tensor], k_nn: int, alpha: float, symmetric: bool, self_loop: bool, binary: bool,
                 k: int, smallest: bool, matrix: str = "laplacian", eps_mode: str = "add1e-6",
                 canonical_sign: bool = True, want_adjacency: bool = False, sigma_mode: bool = False,
                 adjacency_in: Optional[torch.Tensor] = None, first: int = 0):
    """centres (B,G,3) -> dict(vals (B,k), vecs (B,G,k), perm (B,k,G) i32, inv_perm, [adjacency (B,G,G)]).
    ``sigma_mode``: the alpha == 0 weights of create_graph_from_centers (exp(-d^2 / (2 sigma^2)), sigma = batch-wide mean
    distance); ``adjacency_in`` (B,G,G): decompose the Laplacian of this adjacency instead of building the graph;
    ``first``: index of the first wanted eigenpair (k <= 8 per call)."""
    src = adjacency_in if adjacency_in is not None else center
    _cuda(center, adjacency_in)
    if center is not None:
        center = center.detach().float().contiguous()
    if adjacency_in is not None:
        adjacency_in = adjacency_in.detach().float().contiguous()
    B, G = src.shape[0], src.shape[1]
    dev = src.device
    vals = torch.empty(B, k, dtype=torch.float32, device=dev)
    vecs = torch.empty(B, G, k, dtype=torch.float32, device=dev)
    perm = torch.empty(B, k, G, dtype=torch.int32, device=dev)
    inv = torch.empty(B, k, G, dtype=torch.int32, device=dev)
    adj = torch.empty(B, G, G, dtype=torch.float32, device=dev) if want_adjacency else None
    ws_bytes = _lib.load().sim_spectral_eig_workspace_bytes(B, G, k)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev) if ws_bytes else None
    flags = spectral_flags(symmetric, self_loop, binary, smallest, matrix, eps_mode, canonical_sign)
    if not sigma_mode and adjacency_in is None and first == 0:
        _lib.call("sim_spectral_eig", _p(center), B, G, int(k_nn), float(alpha), flags, int(k), _p(vals), _p(vecs),
                  _p(perm), _p(inv), _p(adj), _p(ws), ws_bytes, _stream())
    else:
        sigma = pairwise_dist_mean(center) if (sigma_mode and adjacency_in is None) else None
        _lib.call("sim_spectral_eig_ex", _p(center), _p(adjacency_in), _p(sigma), B, G, int(k_nn), float(alpha), flags,
                  int(first), int(k), _p(vals), _p(vecs), _p(perm), _p(inv), _p(adj), _p(ws), ws_bytes, _stream())
    out = dict(vals=vals, vecs=vecs, perm=perm, inv_perm=inv)
    if want_adjacency:
        out["adjacency"] = adj
    return out


def eig_from_adjacency(adj: torch.Tensor, k: int, smallest: bool, matrix: str = "laplacian", eps_mode: str = "add1e-6",
                       full: bool = True, canonical_sign: bool = False):
    """PointMamba.calc_top_k_eigenvalues_eigenvectors(adj_matrices, k, smallest) -> the reference's 4-tuple
    (top_k_eigenvalues (B,k), top_k_eigenvectors (B,G,k), eigenvalues (B,G) ascending, eigenvectors (B,G,G)); the full
    decomposition (only the wavelet code of the reference reads it) costs ceil(G / 8) extra calls - ``full=False`` returns
    None for the last two.  Signs are LAPACK's in the reference, i.e. unspecified: here the un-normalised kernel output
    unless ``canonical_sign``."""
    G = adj.shape[1]
    sym = matrix != "laplacian"
    top = spectral_eig(None, 1, 0.0, False, False, True, k, smallest, matrix, eps_mode, canonical_sign, adjacency_in=adj)
    if not full:
        return top["vals"], top["vecs"], None, None
    vals, vecs = [], []
    for first in range(0, G, 8):
        kk = min(8, G - first)
        o = spectral_eig(None, 1, 0.0, False, False, True, kk, True, matrix, eps_mode, canonical_sign, adjacency_in=adj,
                         first=first - (1 if sym else 0))
        vals.append(o["vals"])
        vecs.append(o["vecs"])
    return top["vals"], top["vecs"], torch.cat(vals, dim=1), torch.cat(vecs, dim=2)


def argsort_rows(keys: torch.Tensor):
    """Stable ascending argsort along the last dim of a 2-D fp32 tensor (any strides) -> (perm, inv_perm) int32."""
    _cuda(keys)
    assert keys.dim() == 2 and keys.dtype == torch.float32
    rows, n = keys.shape
    perm = torch.empty(rows, n, dtype=torch.int32, device=keys.device)
    inv = torch.empty_like(perm)
    _lib.call("sim_argsort_rows", _p(keys), keys.stride(0), keys.stride(1), rows, n, _p(perm), _p(inv), _stream())
    return perm, inv


# ----------------------------------------------------------------------------- order gather (a-6)
class _OrderGather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, perm, inv_perm, reverse):
        B, G, Cc = x.shape
        k = perm.shape[1]
        T = (2 if reverse else 1) * k * G
        out = torch.empty(B, T, Cc, dtype=x.dtype, device=x.device)
        _lib.call("sim_order_gather_fwd", _p(x), None, _p(perm), _p(out), None, B, G, k, Cc, int(reverse), _dt(x),
                  _stream())
        ctx.save_for_backward(inv_perm)
        ctx.meta = (B, G, k, Cc, reverse)
        return out

    @staticmethod
    def backward(ctx, dout):
        (inv_perm,) = ctx.saved_tensors
        B, G, k, Cc, reverse = ctx.meta
        dout = dout.contiguous()
        dx = torch.empty(B, G, Cc, dtype=dout.dtype, device=dout.device)
        _lib.call("sim_order_gather_bwd", _p(dout), _p(inv_perm), _p(dx), B, G, k, Cc, int(reverse), _dt(dout),
                  _stream())
        return dx, None, None, None


def order_gather(x: torch.Tensor, perm: torch.Tensor, reverse: bool = True,
                 inv_perm: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x (B,G,C), perm (B,k,G) int32 -> (B, (2|1)*k*G, C): cat of the k sorted copies (+ flipped copy)."""
    _cuda(x, perm)
    x = x.contiguous()
    perm = perm.contiguous()
    if inv_perm is None:
        if x.requires_grad and torch.is_grad_enabled():
            inv_perm = torch.empty_like(perm)
            ar = torch.arange(perm.shape[-1], dtype=torch.int32, device=perm.device).expand_as(perm)
            inv_perm.scatter_(2, perm.long(), ar)
        else:
            inv_perm = perm  # unused without grad
    return _OrderGather.apply(x, perm, inv_perm.contiguous(), bool(reverse))


def order_gather_add(x: torch.Tensor, x2: torch.Tensor, perm: torch.Tensor, reverse: bool = True) -> torch.Tensor:
    """Inference fast path: gather(x)+gather(x2) in one pass (tokens + pos, point_mamba.py:250)."""
    _cuda(x, x2, perm)
    x, x2, perm = x.contiguous(), x2.to(x.dtype).contiguous(), perm.contiguous()
    B, G, Cc = x.shape
    k = perm.shape[1]
    T = (2 if reverse else 1) * k * G
    out = torch.empty(B, T, Cc, dtype=x.dtype, device=x.device)
    _lib.call("sim_order_gather_fwd", _p(x), _p(x2), _p(perm), _p(out), None, B, G, k, Cc, int(reverse), _dt(x),
              _stream())
    return out


def gather_rows(x: torch.Tensor, src_idx: torch.Tensor, fill: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[b,t] = x[b, src_idx[b,t]] if src_idx >= 0 else fill (zeros when fill is None).  Forward only."""
    _cuda(x, src_idx, fill)
    x = x.contiguous()
    src_idx = src_idx.to(torch.int32).contiguous()
    B, R_in, Cc = x.shape
    R_out = src_idx.shape[1]
    out = torch.empty(B, R_out, Cc, dtype=x.dtype, device=x.device)
    f = None if fill is None else fill.to(x.dtype).contiguous()
    _lib.call("sim_gather_rows", _p(x), _p(src_idx), _p(f), _p(out), B, R_in, R_out, Cc, _dt(x), _stream())
    return out


def invert_row_map(src_idx: torch.Tensor, R_in: int, fanout: int, check: bool = False) -> torch.Tensor:
    """src_idx (B, R_out) int32 -> inv (B, R_in, fanout) int32: the output rows that read each source row, ascending,
    -1 padded (sim_invert_row_map).  ``check`` syncs and raises if a row has more than ``fanout`` readers."""
    _cuda(src_idx)
    B, R_out = src_idx.shape
    inv = torch.empty(B, R_in, fanout, dtype=torch.int32, device=src_idx.device)
    err = torch.zeros(1, dtype=torch.int32, device=src_idx.device)
    _lib.call("sim_invert_row_map", _p(src_idx), B, R_in, R_out, fanout, _p(inv), _p(err), _stream())
    if check and int(err.item()) != 0:
        raise ValueError(f"invert_row_map: a source row of cloud {int(err.item()) - 1} has more than {fanout} readers")
    return inv


def gather_sum_rows(x: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """out[b,r] = sum_j x[b, idx[b,r,j]] over idx >= 0 (sim_gather_sum_rows); x (B,R_in,C), idx (B,R_out,J) int32."""
    _cuda(x, idx)
    x = x.contiguous()
    B, R_in, Cc = x.shape
    R_out, J = idx.shape[1], idx.shape[2]
    out = torch.empty(B, R_out, Cc, dtype=x.dtype, device=x.device)
    _lib.call("sim_gather_sum_rows", _p(x), _p(idx), _p(out), B, R_in, R_out, J, Cc, _dt(x), _stream())
    return out


# ----------------------------------------------------------------------------- Encoder / head glue (a-2, a-14)
def group_max(x: torch.Tensor, M: int) -> torch.Tensor:
    """x (groups*M, C) -> (groups, C): max over each group's M rows (forward only)."""
    _cuda(x)
    x = x.contiguous()
    P, C = x.shape
    out = torch.empty(P // M, C, dtype=x.dtype, device=x.device)
    _lib.call("sim_group_max", _p(x), _p(out), P // M, M, C, _dt(x), _stream())
    return out


def group_bias_relu_(x: torch.Tensor, gvec: torch.Tensor, M: int) -> torch.Tensor:
    """In place x[p] = relu(x[p] + gvec[p // M]) for x (groups*M, C), gvec (groups, C) (forward only)."""
    _cuda(x, gvec)
    assert x.is_contiguous() and gvec.is_contiguous() and x.dtype == gvec.dtype
    P, C = x.shape
    _lib.call("sim_group_bias_relu", _p(x), _p(gvec), P, M, C, _dt(x), _stream())
    return x


def layernorm_mean(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """mean over tokens of LayerNorm(x): x (B, L, C) fp32 -> (B, C) (forward only)."""
    _cuda(x, weight, bias)
    x = x.float().contiguous()
    B, L, C = x.shape
    out = torch.zeros(B, C, dtype=torch.float32, device=x.device)
    _lib.call("sim_layernorm_mean", _p(x), _p(_f32c(weight)), _p(_f32c(bias)), _p(out), B, L, C, float(eps), _stream())
    return out


def mlp3_relu_rows(x: torch.Tensor, w1t, b1, w2t, b2, w3t, b3) -> torch.Tensor:
    """y = W3 relu(W2 relu(W1 x + b1) + b2) + b3 for a few rows (sim_mlp3_relu_rows): x (rows, d0) fp32, weights
    transposed (in, out) fp32 contiguous with widths <= 256 (the eval-mode classifier head, BatchNorm folded)."""
    _cuda(x, w1t, w2t, w3t)
    assert x.dim() == 2 and x.dtype == torch.float32 and x.stride(1) == 1
    rows, d0 = x.shape
    d1, d2, d3 = w1t.shape[1], w2t.shape[1], w3t.shape[1]
    assert w1t.shape[0] == d0 and w2t.shape[0] == d1 and w3t.shape[0] == d2
    y = torch.empty(rows, d3, dtype=torch.float32, device=x.device)
    _lib.call("sim_mlp3_relu_rows", _p(x), x.stride(0), rows, d0, _p(w1t), _p(b1), d1, _p(w2t), _p(b2), d2, _p(w3t), _p(b3),
              d3, _p(y), y.stride(0), _stream())
    return y


# ----------------------------------------------------------------------------- Chamfer-L2 (a-18)
class ChamferL2(torch.autograd.Function):
    """(R,P,3), (R,Q,3) fp32 -> (R,) pytorch3d chamfer_distance(..., batch_reduction=None)[0] (squared L2, mean over
    points); backward through the recorded arg-mins (sim_chamfer_l2_fwd / _bwd)."""

    @staticmethod
    def forward(ctx, x, y):
        _cuda(x, y)
        x, y = x.float().contiguous(), y.float().contiguous()
        R, P, _ = x.shape
        Q = y.shape[1]
        loss = torch.empty(R, dtype=torch.float32, device=x.device)
        ix = torch.empty(R, P, dtype=torch.int32, device=x.device)
        iy = torch.empty(R, Q, dtype=torch.int32, device=x.device)
        _lib.call("sim_chamfer_l2_fwd", _p(x), _p(y), R, P, Q, _p(loss), _p(ix), _p(iy), _stream())
        ctx.save_for_backward(x, y, ix, iy)
        return loss

    @staticmethod
    def backward(ctx, gloss):
        x, y, ix, iy = ctx.saved_tensors
        R, P, _ = x.shape
        Q = y.shape[1]
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dy = torch.empty_like(y) if ctx.needs_input_grad[1] else None
        if dx is None and dy is None:
            return None, None
        _lib.call("sim_chamfer_l2_bwd", _p(x), _p(y), _p(ix), _p(iy), _p(gloss.float().contiguous()), R, P, Q, _p(dx),
                  _p(dy), _stream())
        return dx, dy


def chamfer_l2(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    return ChamferL2.apply(x, y)


# ----------------------------------------------------------------------------- 3-NN interpolation (a-19)
class ThreeNNInterpolate(torch.autograd.Function):
    """xyz1 (B,N,3), xyz2 (B,S,3), points2 (B,S,C) -> (B,N,C): inverse-squared-distance interpolation from the three
    nearest centres (PointNetFeaturePropagation.forward, pointnet2_utils.py:273-311); gradient to points2 only, as in
    the reference (the coordinates are inputs).  sim_three_nn_interp_fwd / sim_three_interp_bwd."""

    @staticmethod
    def forward(ctx, xyz1, xyz2, points2):
        _cuda(xyz1, xyz2, points2)
        xyz1, xyz2, p2 = _f32c(xyz1), _f32c(xyz2), points2.detach().float().contiguous()
        B, N, _ = xyz1.shape
        S, C = p2.shape[1], p2.shape[2]
        out = torch.empty(B, N, C, dtype=torch.float32, device=p2.device)
        idx = torch.empty(B, N, 3, dtype=torch.int32, device=p2.device)
        w = torch.empty(B, N, 3, dtype=torch.float32, device=p2.device)
        _lib.call("sim_three_nn_interp_fwd", _p(xyz1), _p(xyz2), _p(p2), B, N, S, C, _p(out), _p(idx), _p(w), _stream())
        ctx.save_for_backward(idx, w)
        ctx.shape = (B, N, S, C)
        ctx.in_dtype = points2.dtype
        return out.to(points2.dtype)

    @staticmethod
    def backward(ctx, dout):
        idx, w = ctx.saved_tensors
        B, N, S, C = ctx.shape
        dp2 = torch.empty(B, S, C, dtype=torch.float32, device=dout.device)
        _lib.call("sim_three_interp_bwd", _p(dout.float().contiguous()), _p(idx), _p(w), B, N, S, C, _p(dp2), _stream())
        return None, None, dp2.to(ctx.in_dtype)


def three_nn_interpolate(xyz1: torch.Tensor, xyz2: torch.Tensor, points2: torch.Tensor) -> torch.Tensor:
    return ThreeNNInterpolate.apply(xyz1, xyz2, points2)


def three_nn(xyz1: torch.Tensor, xyz2: torch.Tensor):
    """Indices (B,N,3) int32 and normalised weights (B,N,3) of the three nearest centres (no interpolation)."""
    _cuda(xyz1, xyz2)
    xyz1, xyz2 = _f32c(xyz1), _f32c(xyz2)
    B, N, _ = xyz1.shape
    idx = torch.empty(B, N, 3, dtype=torch.int32, device=xyz1.device)
    w = torch.empty(B, N, 3, dtype=torch.float32, device=xyz1.device)
    _lib.call("sim_three_nn_interp_fwd", _p(xyz1), _p(xyz2), None, B, N, xyz2.shape[1], 0, None, _p(idx), _p(w), _stream())
    return idx, w


# ----------------------------------------------------------------------------- MAE layout (a-16 / a-17)
def mae_index_maps(perm: torch.Tensor, mask: torch.Tensor, n_vis: int, check: bool = False) -> dict:
    """perm (B,k,G) int32, mask (B,G) bool with G - n_vis masked patches per cloud -> the index maps of the masked
    spectral sort and the token restore (include/simamba.h, sim_mae_index_maps), one kernel, no host sync unless
    ``check`` (which raises if a cloud's visible count differs from n_vis)."""
    _cuda(perm, mask)
    B, k, G = perm.shape
    T, RV = 2 * k * G, 2 * k * n_vis
    dev, i32 = perm.device, torch.int32
    perm = perm.to(i32).contiguous()
    m8 = mask.to(torch.uint8).contiguous()
    out = dict(perm_full=torch.empty(B, T, dtype=i32, device=dev), mask_full=torch.empty(B, T, dtype=torch.uint8, device=dev),
               restore_src=torch.empty(B, T, dtype=i32, device=dev), inv_vis=torch.empty(B, G, 2 * k, dtype=i32, device=dev))
    base = {}
    for name, n in (("src_vis", RV), ("vis_pos", RV), ("rec_src", T - RV)):  # never hand the ABI a null (empty) buffer
        base[name] = torch.empty(B * n + 1, dtype=i32, device=dev)
        out[name] = base[name][:B * n].view(B, n)
    err = torch.zeros(1, dtype=i32, device=dev)
    _lib.call("sim_mae_index_maps", _p(perm), _p(m8), B, k, G, n_vis, _p(out["perm_full"]), _p(out["mask_full"]),
              _p(out["restore_src"]), _p(base["src_vis"]), _p(base["vis_pos"]), _p(base["rec_src"]), _p(out["inv_vis"]),
              _p(err), _stream())
    if check and int(err.item()) != 0:
        raise ValueError(f"mae_index_maps: cloud {int(err.item()) - 1} does not have {n_vis} visible patches")
    out["mask_full"] = out["mask_full"].bool()
    return out


class MaeCompact(torch.autograd.Function):
    """x_vis[b,r] = tokens[b, src_vis[b,r]] (sim_mae_compact_fwd / _bwd: the backward is a gather over inv_vis)."""

    @staticmethod
    def forward(ctx, tokens, src_vis, inv_vis):
        tokens = tokens.contiguous()
        B, G, C = tokens.shape
        R = src_vis.shape[1]
        out = torch.empty(B, R, C, dtype=tokens.dtype, device=tokens.device)
        _lib.call("sim_mae_compact_fwd", _p(tokens), _p(src_vis), _p(out), B, G, R, C, _dt(tokens), _stream())
        ctx.save_for_backward(inv_vis)
        ctx.G = G
        return out

    @staticmethod
    def backward(ctx, dout):
        (inv_vis,) = ctx.saved_tensors
        dout = dout.contiguous()
        B, R, C = dout.shape
        dx = torch.empty(B, ctx.G, C, dtype=dout.dtype, device=dout.device)
        _lib.call("sim_mae_compact_bwd", _p(dout), _p(inv_vis), _p(dx), B, ctx.G, R, inv_vis.shape[2], C, _dt(dout),
                  _stream())
        return dx, None, None


class MaeRestore(torch.autograd.Function):
    """x_full[b,t] = restore_src >= 0 ? x_vis[b, restore_src[b,t]] : mask_token (sim_mae_restore_fwd / _bwd)."""

    @staticmethod
    def forward(ctx, x_vis, mask_token, restore_src, vis_pos):
        x_vis = x_vis.contiguous()
        B, R, C = x_vis.shape
        T = restore_src.shape[1]
        fill = mask_token.reshape(-1).to(x_vis.dtype).contiguous()
        out = torch.empty(B, T, C, dtype=x_vis.dtype, device=x_vis.device)
        _lib.call("sim_mae_restore_fwd", _p(x_vis), _p(restore_src), _p(fill), _p(out), B, R, T, C, _dt(x_vis), _stream())
        ctx.save_for_backward(restore_src, vis_pos)
        ctx.token_shape, ctx.token_dtype = mask_token.shape, mask_token.dtype
        return out

    @staticmethod
    def backward(ctx, dout):
        restore_src, vis_pos = ctx.saved_tensors
        dout = dout.contiguous()
        B, T, C = dout.shape
        R = vis_pos.shape[1]
        dx = torch.empty(B, R, C, dtype=dout.dtype, device=dout.device)
        dtok = torch.zeros(C, dtype=torch.float32, device=dout.device)
        _lib.call("sim_mae_restore_bwd", _p(dout), _p(vis_pos), _p(restore_src), _p(dx), _p(dtok), B, R, T, C, _dt(dout),
                  _stream())
        return dx, dtok.to(ctx.token_dtype).reshape(ctx.token_shape), None, None


# ----------------------------------------------------------------------------- add + LayerNorm (a-9)
def add_layernorm(x: torch.Tensor, residual: Optional[torch.Tensor], weight: torch.Tensor, bias: torch.Tensor,
                  eps: float = 1e-5, x2: Optional[torch.Tensor] = None, out_dtype: Optional[torch.dtype] = None,
                  want_residual: bool = True, split: bool = False):
    """res = x (+ x2) (+ residual) in fp32;  y = LayerNorm(res).  -> (y, res or None).  Forward only.
    ``split``: y is returned as a Split3 for linear_split3 - True: three bf16 planes of the fp32 result, "f16x2": two
    fp16 planes (see split2h)."""
    _cuda(x, residual, weight, bias, x2)
    x = x.contiguous()
    Cc = x.shape[-1]
    rows = x.numel() // Cc
    out_dtype = out_dtype or x.dtype
    res_out = torch.empty(x.shape, dtype=torch.float32, device=x.device) if want_residual else None
    if residual is not None:
        residual = residual.float().contiguous()
    if x2 is not None:
        x2 = x2.to(x.dtype).contiguous()
    if split:
        assert Cc % 8 == 0
        if split == "f16x2":  # two fp16 planes (hi, 2^11 lo): the caller has checked that |y| stays inside the fp16 range
            planes = torch.empty(2, rows, Cc, dtype=torch.float16, device=x.device)
            entry = "sim_add_layernorm_split2h"
        else:
            planes = torch.empty(3, rows, Cc, dtype=torch.bfloat16, device=x.device)
            entry = "sim_add_layernorm_split3"
        _lib.call(entry, _p(x), _p(x2), _p(residual), _p(_f32c(weight)), _p(_f32c(bias)),
                  _p(res_out), _p(planes), planes.stride(0), rows, Cc, float(eps), _dt(x), _stream())
        return Split3(planes, x.shape), res_out
    y = torch.empty(x.shape, dtype=out_dtype, device=x.device)
    _lib.call("sim_add_layernorm", _p(x), _p(x2), _p(residual), _p(_f32c(weight)), _p(_f32c(bias)), _p(res_out),
              _p(y), rows, Cc, float(eps), _dt(x), _dt(y), _stream())
    return y, res_out


# ----------------------------------------------------------------------------- fp32 GEMM on tensor cores (a-10)
class AddLayerNorm(torch.autograd.Function):
    """(y, res) = (LayerNorm(s * x + residual), s * x + residual) with the residual stream in fp32: sim_add_layernorm
    forward, sim_add_layernorm_bwd backward (statistics recomputed from ``res``, which autograd keeps alive anyway).
    ``row_scale`` (B,) fp32 or None is the DropPath factor of each sample (mask / keep, block.py:59), applied to x inside
    the kernels - forward and backward - instead of two elementwise passes each way."""

    @staticmethod
    def forward(ctx, x, residual, weight, bias, eps, out_dtype, row_scale=None):
        if row_scale is None:
            y, res = add_layernorm(x, residual, weight, bias, eps, out_dtype=out_dtype, want_residual=True)
        else:
            _cuda(x, residual, weight, bias, row_scale)
            x = x.contiguous()
            C = x.shape[-1]
            rows = x.numel() // C
            assert x.dim() == 3 and row_scale.numel() == x.shape[0] and row_scale.dtype == torch.float32
            res = torch.empty(x.shape, dtype=torch.float32, device=x.device)
            y = torch.empty(x.shape, dtype=out_dtype, device=x.device)
            r_in = None if residual is None else residual.float().contiguous()
            _lib.call("sim_add_layernorm_droppath", _p(x), _p(row_scale.contiguous()), x.shape[1], _p(r_in),
                      _p(_f32c(weight)), _p(_f32c(bias)), _p(res), _p(y), rows, C, float(eps), _dt(x), _dt(y), _stream())
        ctx.save_for_backward(res, weight, row_scale)
        ctx.eps = eps
        ctx.x_dtype = x.dtype
        ctx.rows_per_sample = x.shape[1] if x.dim() == 3 else 0
        ctx.res_dtype = None if residual is None else residual.dtype
        ctx.wdtype = weight.dtype
        ctx.mark_non_differentiable()
        return y, res

    @staticmethod
    def backward(ctx, dy, dres_out):
        res, weight, row_scale = ctx.saved_tensors
        C = res.shape[-1]
        rows = res.numel() // C
        dy = dy.contiguous()
        if dres_out is not None:
            dres_out = dres_out.float().contiguous()
        dres = torch.empty_like(res)
        stats = torch.zeros(2, C, dtype=torch.float32, device=res.device)  # dgamma | dbeta, one fill
        dg, db = stats[0], stats[1]
        need_dx = row_scale is not None or ctx.x_dtype != torch.float32
        dx = torch.empty(res.shape, dtype=ctx.x_dtype, device=res.device) if need_dx else None
        _lib.call("sim_add_layernorm_bwd_dx", _p(res), _p(dy), _p(dres_out), _p(_f32c(weight)), _p(row_scale),
                  int(ctx.rows_per_sample), _p(dres), _p(dx), 0 if dx is None else _dt(dx), _p(dg), _p(db), rows, C,
                  float(ctx.eps), _dt(dy), _stream())
        if dx is None:
            dx = dres
        dr = None if ctx.res_dtype is None else (dres if ctx.res_dtype == torch.float32 else dres.to(ctx.res_dtype))
        return dx, dr, dg.to(ctx.wdtype), db.to(ctx.wdtype), None, None, None


class Split3:
    """An fp32 activation (..., K) carried as three bf16 planes (3, rows, K) - or, where its magnitude is provably
    inside the fp16 range, two fp16 planes (2, rows, K) -: the operand formats of the tcgen05 projection GEMM
    (csrc/gemm_split3.cu).  Produced by add_layernorm / causal_conv1d_tm / selective_scan_tm with ``split`` or by
    split3() / split2h(); consumed by linear_split3."""

    __slots__ = ("planes", "shape")

    def __init__(self, planes: torch.Tensor, shape):
        self.planes, self.shape = planes, tuple(shape)

    @property
    def dtype(self):
        return torch.float32

    @property
    def is_cuda(self):
        return True

    @property
    def requires_grad(self):
        return False


def split3(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 (..., K) -> (3, rows, Kp) bf16 planes with x = p0 + p1 + p2 (exact residuals), Kp = K rounded up to 8."""
    _cuda(x)
    assert x.dtype == torch.float32
    K = x.shape[-1]
    x2 = x if x.dim() == 2 else x.reshape(-1, K) if x.is_contiguous() else _as_rows(x)
    rows, Kp = x2.shape[0], (K + 7) // 8 * 8
    if out is None:
        out = (torch.zeros if Kp != K else torch.empty)(3, rows, Kp, dtype=torch.bfloat16, device=x.device)
    _lib.call("sim_split3_bf16", _p(x2), x2.stride(0), rows, K, _p(out), out.stride(1), out.stride(0), _stream())
    return out


def split2h(x: torch.Tensor) -> torch.Tensor:
    """fp32 (..., K) -> (2, rows, Kp) fp16 planes with x = p0 + 2^-11 p1 to 2^-22 relative (p0 = fp16(x), p1 =
    fp16(2^11 (x - p0))), Kp = K rounded up to 8.  Only for operands with |x| < 65504 (callers check)."""
    _cuda(x)
    assert x.dtype == torch.float32
    K = x.shape[-1]
    x2 = x if x.dim() == 2 else x.reshape(-1, K) if x.is_contiguous() else _as_rows(x)
    rows, Kp = x2.shape[0], (K + 7) // 8 * 8
    out = (torch.zeros if Kp != K else torch.empty)(2, rows, Kp, dtype=torch.float16, device=x.device)
    _lib.call("sim_split2_f16", _p(x2), x2.stride(0), rows, K, _p(out), out.stride(1), out.stride(0), _stream())
    return out


def split3_t(x: torch.Tensor) -> torch.Tensor:
    """fp32 (rows, K) -> planes of the transpose (3, K, rows_p) bf16, rows_p = rows rounded up to 8 (zero padded): one
    kernel instead of ``split3(x.t().contiguous())``."""
    _cuda(x)
    assert x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1
    rows, K = x.shape
    rp = (rows + 7) // 8 * 8
    out = (torch.zeros if rp != rows else torch.empty)(3, K, rp, dtype=torch.bfloat16, device=x.device)
    _lib.call("sim_split3_bf16_t", _p(x), x.stride(0), rows, K, _p(out), out.stride(1), out.stride(0), _stream())
    return out


_EPI_ACT = {None: 0, "silu_from": 1, "softplus_bias": 2}


def linear_split3(xs: torch.Tensor, ws: torch.Tensor, K: int, out: Optional[torch.Tensor] = None, act: Optional[str] = None,
                  act_col0: int = 0, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y (rows, N) fp32 = x @ w.T from pre-split planes (tcgen05 kernel): xs (3, rows, >=K), ws (3, N, >=K) bf16, or
    xs (2, rows, >=K), ws (2, N, >=K) fp16 (split2h).  ``act`` (inference): "silu_from" - columns >= act_col0 leave as
    silu(y); "softplus_bias" - every column leaves as softplus(y + bias) (sim_gemm_planes)."""
    _cuda(xs, ws, bias)
    assert xs.dtype == ws.dtype and xs.dtype in (torch.bfloat16, torch.float16) and xs.stride(2) == 1 and ws.stride(2) == 1
    np_ = 3 if xs.dtype == torch.bfloat16 else 2
    assert xs.shape[0] == np_ and ws.shape[0] == np_
    M, N = xs.shape[1], ws.shape[1]
    if out is None:
        out = torch.empty(M, N, dtype=torch.float32, device=xs.device)
    if np_ == 3 and act is None:
        _lib.call("sim_gemm_bf16x3", _p(xs), xs.stride(1), xs.stride(0), _p(ws), ws.stride(1), ws.stride(0), _p(out),
                  out.stride(0), M, N, K, _stream())
    else:
        _lib.call("sim_gemm_planes", np_, _p(xs), xs.stride(1), xs.stride(0), _p(ws), ws.stride(1), ws.stride(0), _p(out),
                  out.stride(0), M, N, K, _EPI_ACT[act], int(act_col0), _p(_f32c(bias)), _stream())
    return out


def _gemm_operand(t: torch.Tensor) -> torch.Tensor:
    """2-D bf16 operand the TMA unit can read in place: unit inner stride, 16-byte aligned base and row stride."""
    assert t.dim() == 2 and t.dtype == torch.bfloat16
    if t.stride(1) != 1 or t.stride(0) % 8 != 0 or t.data_ptr() % 16 != 0:
        t = t.contiguous()
        if t.stride(0) % 8 != 0:  # odd widths (never on the SI-Mamba shapes): pad the row pitch
            p = torch.zeros(t.shape[0], (t.shape[1] + 7) // 8 * 8, dtype=t.dtype, device=t.device)
            p[:, :t.shape[1]] = t
            t = p[:, :t.shape[1]]
    return t


def gemm_bf16(a: torch.Tensor, b: torch.Tensor, a_mn: bool = False, b_mn: bool = False,
              out_dtype: torch.dtype = torch.bfloat16, out: Optional[torch.Tensor] = None, splits: int = 1,
              silu_col0: Optional[int] = None) -> torch.Tensor:
    """Y = op(a) @ op(b).T on the tcgen05 bf16 kernel (sim_gemm_bf16).  a: (M,K), or (K,M) when ``a_mn``; b: (N,K), or (K,N)
    when ``b_mn`` - row-major views with a uniform row stride are read in place.  ``splits`` = 0: automatic split-K into an
    fp32 result (weight gradients)."""
    _cuda(a, b, out)
    a, b = _gemm_operand(a), _gemm_operand(b)
    M, K = (a.shape[1], a.shape[0]) if a_mn else a.shape
    N, Kb = (b.shape[1], b.shape[0]) if b_mn else b.shape
    assert K == Kb, f"contraction sizes differ: {K} vs {Kb}"
    if splits != 1:
        out_dtype = torch.float32
    if out is None:
        out = (torch.zeros if splits != 1 else torch.empty)(M, N, dtype=out_dtype, device=a.device)
    assert out.shape == (M, N) and out.stride(1) == 1 and out.dtype in (torch.bfloat16, torch.float32)
    if silu_col0 is not None:  # inference in_proj: columns >= silu_col0 (the z half) leave as silu(z)
        assert not a_mn and not b_mn and splits == 1
        _lib.call("sim_gemm_bf16_silu", _p(a), a.stride(0), _p(b), b.stride(0), _p(out), out.stride(0),
                  int(out.dtype == torch.bfloat16), M, N, K, int(silu_col0), _stream())
        return out
    _lib.call("sim_gemm_bf16", _p(a), a.stride(0), int(a_mn), _p(b), b.stride(0), int(b_mn), _p(out), out.stride(0),
              int(out.dtype == torch.bfloat16), M, N, K, int(splits), _stream())
    return out


def gemm_tf32(a: torch.Tensor, b: torch.Tensor, a_mn: bool = False, b_mn: bool = False, bias: Optional[torch.Tensor] = None,
              relu: bool = False, out: Optional[torch.Tensor] = None, splits: int = 1) -> torch.Tensor:
    """fp32 Y = op(a) @ op(b).T (+ bias, ReLU) with the operands consumed as TF32 on the tcgen05 kernel (sim_gemm_tf32): the
    Encoder's 1x1 convolutions at the precision torch's default cudnn.allow_tf32 = True gives the reference's Conv1d."""
    _cuda(a, b, bias, out)
    assert a.dtype == torch.float32 and b.dtype == torch.float32 and a.dim() == 2 and b.dim() == 2

    def operand(t):
        return t if (t.stride(1) == 1 and t.stride(0) % 4 == 0 and t.data_ptr() % 16 == 0) else t.contiguous()
    a, b = operand(a), operand(b)
    M, K = (a.shape[1], a.shape[0]) if a_mn else a.shape
    N, Kb = (b.shape[1], b.shape[0]) if b_mn else b.shape
    assert K == Kb and (a.stride(0) % 4 == 0) and (b.stride(0) % 4 == 0)
    if out is None:
        out = (torch.zeros if splits != 1 else torch.empty)(M, N, dtype=torch.float32, device=a.device)
    _lib.call("sim_gemm_tf32", _p(a), a.stride(0), int(a_mn), _p(b), b.stride(0), int(b_mn), _p(out), out.stride(0), 0, M, N, K,
              int(splits), _p(_f32c(bias)), int(relu), _stream())
    return out


def gemm_tf32_group(a: torch.Tensor, b: torch.Tensor, bias: Optional[torch.Tensor] = None,
                    gbias: Optional[torch.Tensor] = None, relu: bool = False, want_y: bool = True, want_gmax: bool = False):
    """gemm_tf32 with the per-patch glue of Encoder.forward in its epilogue (sim_gemm_tf32_group); rows of ``a`` are points,
    every 32 consecutive rows one patch.  y = relu?(a @ b.T + bias + gbias[row // 32]) (returned if ``want_y``), gmax =
    y.view(-1, 32, N).max(1) (returned if ``want_gmax``): the `torch.max(feature, dim=2)` / `cat([global.expand, local])`
    passes of models/point_mamba.py:66-72 without a separate kernel or, for the last conv, the (points, C) tensor itself.
    -> (y or None, gmax or None)."""
    _cuda(a, b, bias, gbias)
    assert a.dtype == torch.float32 and b.dtype == torch.float32 and a.dim() == 2 and b.dim() == 2
    assert gbias is not None or want_gmax

    def operand(t):
        return t if (t.stride(1) == 1 and t.stride(0) % 4 == 0 and t.data_ptr() % 16 == 0) else t.contiguous()
    a, b = operand(a), operand(b)
    (M, K), (N, Kb) = a.shape, b.shape
    assert K == Kb and M % 32 == 0
    if gbias is not None:
        gbias = operand(gbias.float())
        assert gbias.shape == (M // 32, N)
    y = torch.empty(M, N, dtype=torch.float32, device=a.device) if want_y else None
    gmax = torch.empty(M // 32, N, dtype=torch.float32, device=a.device) if want_gmax else None
    _lib.call("sim_gemm_tf32_group", _p(a), a.stride(0), _p(b), b.stride(0), _p(y), 0 if y is None else y.stride(0), M, N, K,
              _p(_f32c(bias)), _p(gbias), 0 if gbias is None else gbias.stride(0), int(relu), _p(gmax),
              0 if gmax is None else gmax.stride(0), _stream())
    return y, gmax


def point_linear3(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], act: str = "none") -> torch.Tensor:
    """y = act(x @ w.T + b) for 3-D points: x (rows, 3), w (C, 3) -> (rows, C) fp32; act in none | relu | gelu (erf)."""
    _cuda(x, w, b)
    x, w = _f32c(x), _f32c(w)
    rows, C = x.shape[0], w.shape[0]
    assert x.shape[1] == 3 and w.shape[1] == 3
    y = torch.empty(rows, C, dtype=torch.float32, device=x.device)
    _lib.call("sim_point_linear3", _p(x), _p(w), _p(_f32c(b)), _p(y), rows, C, {"none": 0, "relu": 1, "gelu": 2}[act], _stream())
    return y


class LinearBF16(torch.autograd.Function):
    """y = x @ w.T for the bf16 (autocast) mixer on the hand-written tcgen05 kernel: x (..., K) bf16, w (N, K) - an fp32
    parameter is cast once per call - -> (..., N) bf16.  Backward: dX = dY @ W and dW = dY^T @ X read dY, W and X in place
    (MN-major operands), dW with split-K in fp32 - the dtype of the master weight's gradient."""

    @staticmethod
    def forward(ctx, x, w):
        wb = w if w.dtype == torch.bfloat16 else w.detach().to(torch.bfloat16)
        K = w.shape[1]
        x2 = x.reshape(-1, K) if x.is_contiguous() or x.dim() == 2 else _as_rows(x)
        ctx.save_for_backward(x2, wb)
        ctx.x_shape, ctx.w_dtype = x.shape, w.dtype
        return gemm_bf16(x2, wb).view(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, wb = ctx.saved_tensors
        N, K = wb.shape
        dy2 = dy.reshape(-1, N)
        if dy2.dtype != torch.bfloat16:
            dy2 = dy2.to(torch.bfloat16)
        dy2 = _gemm_operand(dy2)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = gemm_bf16(dy2, wb, b_mn=True).view(ctx.x_shape)
        if ctx.needs_input_grad[1]:
            dw = gemm_bf16(dy2, x2, a_mn=True, b_mn=True, splits=0)
            if ctx.w_dtype != torch.float32:
                dw = dw.to(ctx.w_dtype)
        return dx, dw


def linear_bf16(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """Differentiable y = x @ w.T on the tcgen05 bf16 kernel (see LinearBF16); plain forward when no graph is needed."""
    if torch.is_grad_enabled() and (x.requires_grad or w.requires_grad):
        return LinearBF16.apply(x, w)
    wb = w if w.dtype == torch.bfloat16 else w.to(torch.bfloat16)
    K = w.shape[1]
    x2 = x.reshape(-1, K) if x.is_contiguous() or x.dim() == 2 else _as_rows(x)
    return gemm_bf16(x2, wb).view(*x.shape[:-1], w.shape[0])


class LinearX3(torch.autograd.Function):
    """y = x @ w.T for fp32 TRAINING on the tcgen05 split-plane GEMM (the reference's fp32 runs do all three GEMMs of a
    Linear as SIMT SGEMMs): forward, dX = dY @ W and dW = dY^T @ X each split their two operands into bf16 planes
    (sim_split3_bf16) and run sim_gemm_bf16x3; accuracy is that of an fp32 GEMM."""

    @staticmethod
    def forward(ctx, x, w):
        K = w.shape[1]
        ctx.save_for_backward(x, w)
        y = linear_split3(split3(x), split3(w), K)
        return y.view(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        N, K = w.shape
        dy2 = dy.reshape(-1, N)
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = linear_split3(split3(dy2), split3_t(w if w.stride(1) == 1 else w.contiguous()), N).view(x.shape)
        if ctx.needs_input_grad[1]:
            x2 = x.reshape(-1, K)
            if x2.stride(1) != 1:
                x2 = x2.contiguous()
            M = x2.shape[0]
            dw = linear_split3(split3_t(dy2), split3_t(x2), M)  # transposed operands split in one pass each
        return dx, dw


def linear_x3_train(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    return LinearX3.apply(x, w)


def linear_split3_planes_out(xs: torch.Tensor, ws: torch.Tensor, K: int, planes_cols: int):
    """linear_split3 for N <= 64 that also returns the first ``planes_cols`` output columns as split planes
    (3, rows, planes_cols): y, planes."""
    _cuda(xs, ws)
    M, N = xs.shape[1], ws.shape[1]
    out = torch.empty(M, N, dtype=torch.float32, device=xs.device)
    planes = torch.empty(3, M, planes_cols, dtype=torch.bfloat16, device=xs.device)
    _lib.call("sim_gemm_bf16x3_split_out", _p(xs), xs.stride(1), xs.stride(0), _p(ws), ws.stride(1), ws.stride(0), _p(out),
              out.stride(0), M, N, K, _p(planes), planes_cols, planes.stride(1), planes.stride(0), _stream())
    return out, planes


def linear_f32a_planes_out(x: torch.Tensor, ws: torch.Tensor, K: int, planes_cols: int = 0):
    """Narrow projection (N <= 64) from the fp32 activation x (..., K) itself: the split happens in-kernel
    (sim_gemm_f32a_bf16x3).  Returns (y (rows, N), planes (3, rows, planes_cols) or None)."""
    _cuda(x, ws)
    assert x.dtype == torch.float32
    x2 = x if x.dim() == 2 else x.reshape(-1, K) if x.is_contiguous() else _as_rows(x)
    M, N = x2.shape[0], ws.shape[1]
    out = torch.empty(M, N, dtype=torch.float32, device=x.device)
    planes = torch.empty(3, M, planes_cols, dtype=torch.bfloat16, device=x.device) if planes_cols else None
    _lib.call("sim_gemm_f32a_bf16x3", _p(x2), x2.stride(0), _p(ws), ws.stride(1), ws.stride(0), _p(out), out.stride(0), M, N, K,
              _p(planes), planes_cols, 0 if planes is None else planes.stride(1), 0 if planes is None else planes.stride(0),
              _stream())
    return out, planes


def conv_xproj_f32(x: torch.Tensor, conv_w: torch.Tensor, conv_b: Optional[torch.Tensor], ws: torch.Tensor,
                   planes_cols: int = 0):
    """Fused causal conv1d(width 4) + SiLU + x_proj (fp32 inference, sim_conv_xproj_f32): x (B, L, D) token-major fp32
    (any uniform row stride) -> u (B, L, D), x_dbl (B, L, N), planes (3, B*L, planes_cols) of x_dbl's first columns."""
    _cuda(x, conv_w, conv_b, ws)
    B, L, D = x.shape
    N = ws.shape[1]
    w = _f32c(conv_w.reshape(D, -1))
    assert x.dtype == torch.float32 and w.shape[1] == 4
    u = torch.empty(B, L, D, dtype=torch.float32, device=x.device)
    x_dbl = torch.empty(B, L, N, dtype=torch.float32, device=x.device)
    planes = torch.empty(3, B * L, planes_cols, dtype=torch.bfloat16, device=x.device) if planes_cols else None
    _lib.call("sim_conv_xproj_f32", _p(x), _tm(x), _p(w), _p(_f32c(conv_b)), _p(u), D, _p(ws), ws.stride(1), ws.stride(0),
              _p(x_dbl), N, B, L, D, N, _p(planes), planes_cols, 0 if planes is None else planes.stride(1),
              0 if planes is None else planes.stride(0), _stream())
    return u, x_dbl, planes


def linear_f32_x3(x, weight_planes: torch.Tensor, K: int, act: Optional[str] = None, act_col0: int = 0,
                  bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = x @ W.T (fp32-accurate) with W given as split planes (see split3 / split2h); x is a Split3 from its producer,
    or an fp32 tensor that is split here (in the weight planes' format)."""
    if isinstance(x, Split3):
        xs = x.planes
    else:
        xs = split3(x) if weight_planes.dtype == torch.bfloat16 else split2h(x)
    y = linear_split3(xs, weight_planes, K, act=act, act_col0=act_col0, bias=bias)
    return y.view(*x.shape[:-1], weight_planes.shape[1])


def _as_rows(x: torch.Tensor) -> torch.Tensor:
    """(B, L, K) view with unit inner stride and batch stride == L * row stride -> (B*L, K) strided 2-D view."""
    assert x.dim() == 3 and x.stride(2) == 1 and (x.shape[0] == 1 or x.stride(0) == x.shape[1] * x.stride(1))
    return x.as_strided((x.shape[0] * x.shape[1], x.shape[2]), (x.stride(1), 1))


# ----------------------------------------------------------------------------- causal conv1d (a-12)
def causal_conv1d_tm(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], silu: bool = True,
                     out: Optional[torch.Tensor] = None, split: bool = False):
    """Token-major causal depthwise conv (forward only): x (B,L,D) (any uniform row stride), weight (D,W) -> (B,L,D).
    ``split`` (fp32): also returns the result as a Split3 for linear_split3 -> (out, Split3)."""
    _cuda(x, weight, bias)
    B, L, D = x.shape
    ld_x = _tm(x)
    if out is None:
        out = torch.empty(B, L, D, dtype=x.dtype, device=x.device)
    w = _f32c(weight.reshape(D, -1))
    if split:
        assert x.dtype == torch.float32 and D % 8 == 0
        planes = torch.empty(3, B * L, D, dtype=torch.bfloat16, device=x.device)
        _lib.call("sim_causal_conv1d_fwd_split3", _p(x), ld_x, _p(w), _p(_f32c(bias)), _p(out), _tm(out), _p(planes),
                  planes.stride(1), planes.stride(0), B, L, D, w.shape[1], int(silu), _stream())
        return out, Split3(planes, (B, L, D))
    _lib.call("sim_causal_conv1d_fwd", _p(x), ld_x, _p(w), _p(_f32c(bias)), _p(out), _tm(out), B, L, D,
              w.shape[1], int(silu), _dt(x), _stream())
    return out


def causal_conv1d_bwd_tm(x, weight, bias, dy, silu: bool = True):
    """-> (dx (B,L,D) in x's dtype, dw (D,W) fp32, dbias (D) fp32)."""
    _cuda(x, weight, bias, dy)
    B, L, D = x.shape
    dy = dy.contiguous()
    dx = torch.empty(B, L, D, dtype=x.dtype, device=x.device)
    w = _f32c(weight.reshape(D, -1))
    dw = torch.zeros_like(w)
    db = torch.zeros(D, dtype=torch.float32, device=x.device) if bias is not None else None
    _lib.call("sim_causal_conv1d_bwd", _p(x), _tm(x), _p(w), _p(_f32c(bias)), _p(dy), _tm(dy), _p(dx), _tm(dx),
              _p(dw), _p(db), B, L, D, w.shape[1], int(silu), _dt(x), _stream())
    return dx, dw, db


class CausalConv1dTM(torch.autograd.Function):
    """Token-major causal conv1d + SiLU with the CUDA backward (recomputes the pre-activation from x)."""

    @staticmethod
    def forward(ctx, x, weight, bias, silu):
        ctx.save_for_backward(x, weight, bias)
        ctx.silu = silu
        return causal_conv1d_tm(x, weight, bias, silu=silu)

    @staticmethod
    def backward(ctx, dy):
        x, weight, bias = ctx.saved_tensors
        dx, dw, db = causal_conv1d_bwd_tm(x, weight, bias, dy, ctx.silu)
        return dx, dw.reshape(weight.shape).to(weight.dtype), None if bias is None else db.to(bias.dtype), None


def causal_conv1d_fn(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None,
                     activation: Optional[str] = None) -> torch.Tensor:
    """causal-conv1d compatible signature: x (B,D,L) channel-major -> (B,D,L); differentiable."""
    assert activation in (None, "silu", "swish")
    xt = x.transpose(1, 2)
    if xt.stride(2) != 1 or (xt.shape[0] > 1 and xt.stride(0) != xt.shape[1] * xt.stride(1)):
        xt = xt.contiguous()
    y = CausalConv1dTM.apply(xt, weight, bias, activation is not None)
    return y.transpose(1, 2)


# ----------------------------------------------------------------------------- selective scan (a-11)
def selective_scan_tm(u, delta, A, Bm, Cm, D=None, z=None, delta_bias=None, delta_softplus=False,
                      out: Optional[torch.Tensor] = None, variant: int = 0,
                      checkpoints: Optional[torch.Tensor] = None, split: bool = False, z_gate: bool = False):
    """Token-major selective scan (forward only).  u, delta, z (B,L,D); Bm, Cm (B,L,N) - all may be column slices of
    wider row-major buffers; A (D,N) fp32.  Returns out (B,L,D) in u's dtype.  ``checkpoints`` (fp32,
    scan_checkpoint_shape) receives the tile-start states the backward kernel needs.  ``z_gate`` (inference): z already
    holds silu(z) (in_proj epilogue, linear_split3(act="silu_from")), the kernel multiplies by it as is."""
    delta_softplus = int(bool(delta_softplus)) | (2 if z_gate else 0)
    _cuda(u, delta, A, Bm, Cm, D, z, delta_bias)
    B, L, Dm = u.shape
    N = A.shape[1]
    u, delta, Bm, Cm, z = (_bulk_ok(t) for t in (u, delta, Bm, Cm, z))
    assert delta.dtype == u.dtype and Bm.dtype == u.dtype and Cm.dtype == u.dtype and (z is None or z.dtype == u.dtype)
    if split:  # fp32 only: the result leaves as the three bf16 planes out_proj consumes (returns a Split3)
        assert u.dtype == torch.float32 and Dm % 64 == 0 and checkpoints is None
        planes = torch.empty(3, B * L, Dm, dtype=torch.bfloat16, device=u.device)
        _lib.call("sim_selective_scan_fwd_split3", _p(u), _tm(u), _p(delta), _tm(delta), _p(_f32c(A)), _p(Bm),
                  _tm(Bm), _p(Cm), _tm(Cm), _p(_f32c(D)), _p(z), 0 if z is None else _tm(z), _p(_f32c(delta_bias)),
                  _p(planes), planes.stride(1), planes.stride(0), B, L, Dm, N, int(delta_softplus), _stream())
        return Split3(planes, (B, L, Dm))
    if out is None:
        out = torch.empty(B, L, Dm, dtype=u.dtype, device=u.device)
    _lib.call("sim_selective_scan_fwd", _p(u), _tm(u), _p(delta), _tm(delta), _p(_f32c(A)), _p(Bm), _tm(Bm), _p(Cm),
              _tm(Cm), _p(_f32c(D)), _p(z), 0 if z is None else _tm(z), _p(_f32c(delta_bias)), _p(out), _tm(out),
              _p(checkpoints), B, L, Dm, N, int(delta_softplus), _dt(u), int(variant), _stream())
    return out


def dt_proj_planes(w_dt: torch.Tensor, act_dtype: torch.dtype) -> torch.Tensor:
    """dt_proj.weight (D, 24) -> the bf16 planes the fused scan consumes: (3, D, 32) for fp32 activations (3 x bf16 split),
    (1, D, 32) for bf16; K zero-padded to 32."""
    D, R = w_dt.shape
    assert R <= 32
    w32 = torch.zeros(D, 32, dtype=torch.float32, device=w_dt.device)
    w32[:, :R] = w_dt.float()
    if act_dtype == torch.float32:
        return split3(w32)
    return w32.to(torch.bfloat16).unsqueeze(0).contiguous()


def selective_scan_fused_dt_tm(u, x_dbl, dt_rank: int, wdt_planes, A, D=None, z=None, delta_bias=None,
                               delta_softplus=True, split: bool = False):
    """Token-major selective scan with dt_proj fused in (inference): x_dbl (B,L,dt_rank+32) = x_proj output rows
    (dt_low | B | C), wdt_planes from dt_proj_planes().  delta never touches HBM.  Returns out (B,L,D) or, with
    ``split`` (fp32), the result as a Split3."""
    _cuda(u, x_dbl, wdt_planes, A, D, z, delta_bias)
    B, L, Dm = u.shape
    N = A.shape[1]
    assert x_dbl.shape[-1] == dt_rank + 2 * N and x_dbl.dtype == u.dtype and (z is None or z.dtype == u.dtype)
    u, x_dbl, z = (_bulk_ok(t) for t in (u, x_dbl, z))
    out = planes = None
    if split:
        assert u.dtype == torch.float32
        planes = torch.empty(3, B * L, Dm, dtype=torch.bfloat16, device=u.device)
    else:
        out = torch.empty(B, L, Dm, dtype=u.dtype, device=u.device)
    _lib.call("sim_selective_scan_fwd_fused_dt", _p(u), _tm(u), _p(x_dbl), _tm(x_dbl), int(dt_rank), _p(wdt_planes),
              _p(_f32c(A)), _p(_f32c(D)), _p(z), 0 if z is None else _tm(z), _p(_f32c(delta_bias)), _p(out),
              0 if out is None else _tm(out), _p(planes), 0 if planes is None else planes.stride(1),
              0 if planes is None else planes.stride(0), B, L, Dm, N, int(delta_softplus), _dt(u), _stream())
    return Split3(planes, (B, L, Dm)) if split else out


def scan_checkpoint_shape(B: int, L: int, D: int):
    return (B, (L + 7) // 8, D, 16)  # kScanCkpt = 8 steps between saved states


def selective_scan_bwd_tm(u, delta, A, Bm, Cm, D, z, delta_bias, dout, checkpoints, delta_softplus=True):
    """-> (du, ddelta, dA, dB, dC, dD, dz, ddelta_bias); du/ddelta/dz/dB/dC in u's dtype, the rest fp32."""
    _cuda(u, delta, A, Bm, Cm, D, z, delta_bias, dout, checkpoints)
    B, L, Dm = u.shape
    N = A.shape[1]
    dev = u.device
    u, delta, Bm, Cm, z, dout = (_bulk_ok(t) for t in (u, delta, Bm, Cm, z, dout))
    du = torch.empty(B, L, Dm, dtype=u.dtype, device=dev)
    ddelta = torch.empty_like(du)
    dz = torch.empty_like(du) if z is not None else None
    # the accumulated outputs share one zero-filled buffer (one fill instead of five)
    nbc, na = B * L * N, Dm * N
    acc = torch.zeros(2 * nbc + na + 2 * Dm, dtype=torch.float32, device=dev)
    dB, dC = acc[:nbc].view(B, L, N), acc[nbc:2 * nbc].view(B, L, N)
    dA = acc[2 * nbc:2 * nbc + na].view(Dm, N)
    dD = acc[2 * nbc + na:2 * nbc + na + Dm] if D is not None else None
    dbias = acc[2 * nbc + na + Dm:] if delta_bias is not None else None
    _lib.call("sim_selective_scan_bwd", _p(u), _tm(u), _p(delta), _tm(delta), _p(_f32c(A)), _p(Bm), _tm(Bm), _p(Cm),
              _tm(Cm), _p(_f32c(D)), _p(z), 0 if z is None else _tm(z), _p(_f32c(delta_bias)), _p(dout), _tm(dout),
              _p(checkpoints), _p(du), _tm(du), _p(ddelta), _tm(ddelta), _p(dz), 0 if dz is None else _tm(dz),
              _p(dB), _p(dC), _p(dA), _p(dD), _p(dbias), B, L, Dm, N, int(delta_softplus), _dt(u), _stream())
    return du, ddelta, dA, dB.to(u.dtype), dC.to(u.dtype), dD, dz, dbias


class SelectiveScanTM(torch.autograd.Function):
    """Token-major selective scan with the CUDA backward; saves the inputs and one checkpoint tensor."""

    @staticmethod
    def forward(ctx, u, delta, A, Bm, Cm, D, z, delta_bias, delta_softplus):
        B, L, Dm = u.shape
        ckpt = torch.empty(scan_checkpoint_shape(B, L, Dm), dtype=torch.float32, device=u.device)
        out = selective_scan_tm(u, delta, A, Bm, Cm, D, z, delta_bias, delta_softplus, checkpoints=ckpt)
        ctx.save_for_backward(u, delta, A, Bm, Cm, D, z, delta_bias, ckpt)
        ctx.delta_softplus = delta_softplus
        return out

    @staticmethod
    def backward(ctx, dout):
        u, delta, A, Bm, Cm, D, z, delta_bias, ckpt = ctx.saved_tensors
        du, ddelta, dA, dB, dC, dD, dz, dbias = selective_scan_bwd_tm(
            u, delta, A, Bm, Cm, D, z, delta_bias, dout.contiguous(), ckpt, ctx.delta_softplus)
        cast = lambda g, ref: None if (g is None or ref is None) else g.to(ref.dtype)
        return du, ddelta, cast(dA, A), dB, dC, cast(dD, D), dz, cast(dbias, delta_bias), None


def _bulk_ok(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """TMA tensor maps need 16-byte aligned row starts; copy the (rare) slices that are not."""
    if t is None:
        return None
    ok = t.stride(2) == 1 and (t.shape[0] == 1 or t.shape[1] == 1 or t.stride(0) == t.shape[1] * t.stride(1)) \
        and (_tm(t) * t.element_size()) % 16 == 0 and t.data_ptr() % 16 == 0
    return t if ok else t.contiguous()


def _to_tm(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if t is None:
        return None
    tt = t.transpose(1, 2)
    ok = tt.stride(2) == 1 and (tt.shape[0] == 1 or tt.stride(0) == tt.shape[1] * tt.stride(1)) \
        and (tt.stride(1) * tt.element_size()) % 16 == 0 and tt.data_ptr() % 16 == 0
    return tt if ok else tt.contiguous()


def selective_scan_fn(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                      return_last_state=False):
    """mamba-ssm compatible signature (channel-major): u, delta, z (B,D,L); B, C (B,N,L); A (D,N); differentiable.

    Channel-major arguments that are transposed views of token-major memory are used in place;
    anything else is transposed once."""
    if return_last_state:
        raise NotImplementedError("return_last_state is not on SI-Mamba's path (inference_params is always None)")
    args = (_to_tm(u), _to_tm(delta), A, _to_tm(B), _to_tm(C), D, _to_tm(z), delta_bias)
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in args):
        out = SelectiveScanTM.apply(*args, bool(delta_softplus))
    else:
        out = selective_scan_tm(*args, delta_softplus)
    return out.transpose(1, 2)
