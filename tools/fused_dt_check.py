"""Timing of the scan with dt_proj fused in vs dt_proj GEMM + scan (C1 layer shape), CUDA graph, rotated inputs."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from si_mamba_b200 import ops  # noqa: E402
from tools.kernel_bench import time_fn  # noqa: E402

B, L, D = 32, 512, 768
for dtype in (torch.float32, torch.bfloat16):
    sets = []
    for i in range(4):
        g = torch.Generator(device="cuda").manual_seed(i)
        r = lambda *s: torch.randn(*s, generator=g, device="cuda")
        sets.append((r(B, L, D).to(dtype), r(B, L, 2 * D).to(dtype)[..., D:], (0.3 * r(B, L, 56)).to(dtype)))
    w_dt = torch.randn(D, 24, device="cuda") * 24 ** -0.5
    A = -torch.arange(1, 17, device="cuda", dtype=torch.float32).repeat(D, 1)
    Dv, bias = torch.ones(D, device="cuda"), torch.full((D,), -4.0, device="cuda")
    planes = ops.dt_proj_planes(w_dt, dtype)
    wsp = ops.split3(w_dt) if dtype == torch.float32 else w_dt.to(dtype)

    def separate(s):
        u, z, x = s
        dt = ops.linear_f32_x3(x[..., :24], wsp, 24) if dtype == torch.float32 else torch.nn.functional.linear(x[..., :24], wsp)
        return ops.selective_scan_tm(u, dt, A, x[..., 24:40], x[..., 40:], Dv, z, bias, True)

    def fused(s):
        u, z, x = s
        return ops.selective_scan_fused_dt_tm(u, x, 24, planes, A, Dv, z, bias, True)

    def scan_only(s, dts={}):
        u, z, x = s
        if id(u) not in dts:
            dts[id(u)] = torch.randn(B, L, D, device="cuda").to(dtype) * 0.5
        return ops.selective_scan_tm(u, dts[id(u)], A, x[..., 24:40], x[..., 40:], Dv, z, bias, True)

    for s in sets:
        scan_only(s)
    for name, fn in (("dt_proj GEMM + scan", separate), ("fused", fused), ("scan only", scan_only)):
        t = time_fn([(lambda s=s: fn(s)) for s in sets])
        print(f"{str(dtype):16s} {name:22s} {t * 1e6:7.1f} us", flush=True)
