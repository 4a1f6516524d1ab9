"""Launch ONE kernel a few times (eagerly, no CUDA graph) so ncu can capture it:

    python tools/prof_one.py scan  [--variant 4] [--dtype fp32|bf16] [--B 32] [--L 512]
    python tools/prof_one.py fps | knn | spectral [--G 64] | conv | addln | gather
"""

from __future__ import annotations

import argparse
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from si_mamba_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("kernel")
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--dtype", default="fp32")
    ap.add_argument("--B", type=int, default=32)
    ap.add_argument("--L", type=int, default=512)
    ap.add_argument("--G", type=int, default=64)
    ap.add_argument("--iters", type=int, default=6)
    a = ap.parse_args()
    dt = torch.float32 if a.dtype == "fp32" else torch.bfloat16
    B, L, D = a.B, a.L, 768
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    r = lambda *s: torch.randn(*s, generator=g, device=dev)
    if a.kernel == "scan":
        sets = []
        for _ in range(3):
            xz, u, dl, xd = r(B, L, 2 * D).to(dt), r(B, L, D).to(dt), (0.5 * r(B, L, D)).to(dt), r(B, L, 56).to(dt)
            sets.append((u, dl, xd[..., 24:40], xd[..., 40:], xz[..., D:], torch.empty(B, L, D, dtype=dt, device=dev)))
        A = -torch.arange(1, 17, device=dev, dtype=torch.float32).repeat(D, 1)
        Dv, bias = torch.ones(D, device=dev), torch.full((D,), -4.0, device=dev)
        for i in range(a.iters):
            s = sets[i % 3]
            ops.selective_scan_tm(s[0], s[1], A, s[2], s[3], Dv, s[4], bias, True, out=s[5], variant=a.variant)
    elif a.kernel == "gemm":
        M, N, K = B * L, 1536, 384
        xs = [ops.split3(r(M, K)) for _ in range(3)]
        ws = ops.split3(r(N, K) * K ** -0.5)
        outs = [torch.empty(M, N, device=dev) for _ in range(3)]
        for i in range(a.iters):
            ops.linear_split3(xs[i % 3], ws, K, out=outs[i % 3])
    elif a.kernel == "gemm_f16":  # in_proj on two fp16 planes with the silu epilogue on the z half (the C1 inference path)
        M, N, K = B * L, 1536, 384
        xs = [ops.split2h(r(M, K)) for _ in range(3)]
        ws = ops.split2h(r(N, K) * K ** -0.5)
        outs = [torch.empty(M, N, device=dev) for _ in range(3)]
        for i in range(a.iters):
            ops.linear_split3(xs[i % 3], ws, K, out=outs[i % 3], act="silu_from", act_col0=N // 2)
    elif a.kernel == "scan_gate":  # the scan as the C1 inference path calls it: z arrives as the gate silu(z)
        sets = []
        for _ in range(3):
            xz, u, dl, xd = r(B, L, 2 * D).to(dt), r(B, L, D).to(dt), (0.5 * r(B, L, D)).to(dt), r(B, L, 56).to(dt)
            sets.append((u, dl, xd[..., 24:40], xd[..., 40:], xz[..., D:], torch.empty(B, L, D, dtype=dt, device=dev)))
        A = -torch.arange(1, 17, device=dev, dtype=torch.float32).repeat(D, 1)
        Dv, bias = torch.ones(D, device=dev), torch.full((D,), -4.0, device=dev)
        for i in range(a.iters):
            s = sets[i % 3]
            ops.selective_scan_tm(s[0], s[1], A, s[2], s[3], Dv, s[4], bias, True, out=s[5], variant=a.variant, z_gate=True)
    elif a.kernel == "scanfused":
        sets = []
        for _ in range(3):
            sets.append((r(B, L, D).to(dt), r(B, L, 2 * D).to(dt)[..., D:], (0.3 * r(B, L, 56)).to(dt)))
        A = -torch.arange(1, 17, device=dev, dtype=torch.float32).repeat(D, 1)
        Dv, bias = torch.ones(D, device=dev), torch.full((D,), -4.0, device=dev)
        planes = ops.dt_proj_planes(torch.randn(D, 24, device=dev) * 0.2, dt)
        for i in range(a.iters):
            s = sets[i % 3]
            ops.selective_scan_fused_dt_tm(s[0], s[2], 24, planes, A, Dv, s[1], bias, True)
    elif a.kernel == "scanbwd":
        xz, u, dl, xd = r(B, L, 2 * D).to(dt), r(B, L, D).to(dt), (0.5 * r(B, L, D)).to(dt), r(B, L, 56).to(dt)
        dout = r(B, L, D).to(dt)
        A = -torch.arange(1, 17, device=dev, dtype=torch.float32).repeat(D, 1)
        Dv, bias = torch.ones(D, device=dev), torch.full((D,), -4.0, device=dev)
        ck = torch.empty(ops.scan_checkpoint_shape(B, L, D), dtype=torch.float32, device=dev)
        ops.selective_scan_tm(u, dl, A, xd[..., 24:40], xd[..., 40:], Dv, xz[..., D:], bias, True, checkpoints=ck)
        for i in range(a.iters):
            ops.selective_scan_bwd_tm(u, dl, A, xd[..., 24:40], xd[..., 40:], Dv, xz[..., D:], bias, dout, ck, True)
    elif a.kernel in ("fps", "knn", "spectral"):
        N = 1024 if a.G <= 64 else 2048
        xyz = torch.rand(B, N, 3, generator=g, device=dev)
        center, _ = ops.fps(xyz, a.G)
        for _ in range(a.iters):
            if a.kernel == "fps":
                ops.fps(xyz, a.G)
            elif a.kernel == "knn":
                ops.knn_group(xyz, center, 32)
            else:
                ops.spectral_eig(center, 20, 100.0, True, False, True, 4, True)
    elif a.kernel == "conv":
        xs = [r(B, L, 2 * D).to(dt) for _ in range(3)]
        w, b = r(D, 4), r(D)
        for i in range(a.iters):
            ops.causal_conv1d_tm(xs[i % 3][..., :D], w, b, True)
    elif a.kernel == "addln":
        xs = [(r(B, L, 384), r(B, L, 384)) for _ in range(4)]
        w, b = torch.ones(384, device=dev), torch.zeros(384, device=dev)
        for i in range(a.iters):
            ops.add_layernorm(xs[i % 4][0], xs[i % 4][1], w, b)
    elif a.kernel == "gather":
        tok = [(r(B, 64, 384), r(B, 64, 384)) for _ in range(4)]
        perm = torch.stack([torch.stack([torch.randperm(64, device=dev) for _ in range(4)]) for _ in range(B)]).int()
        for i in range(a.iters):
            ops.order_gather_add(tok[i % 4][0], tok[i % 4][1], perm, True)
    else:
        raise SystemExit(f"unknown kernel {a.kernel}")
    torch.cuda.synchronize()
    print("ok")


if __name__ == "__main__":
    main()
