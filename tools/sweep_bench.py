#!/usr/bin/env python
"""BASELINE.json config 5: kernel sweep - spectral ordering (graph + eigensolver + argsort) at 64-512 patches and the
selective scan at L = 64-4096, batch 1-4096 (capped at ~8 GB of activations), fp32 and bf16.  One JSON line per point:
time per launch (CUDA events around a CUDA graph, inputs rotated over > L2), achieved algorithmic GB/s for the scan and
its fraction of the measured HBM peak; clouds/s for the ordering.  Under torchrun every rank runs the same sweep on its
own GPU (the kernels shard by batch with no communication) and rank 0 reports the per-GPU numbers.

    python tools/sweep_bench.py [--quick]
"""
from __future__ import annotations

import json
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from si_mamba_b200 import ops  # noqa: E402
from tools.kernel_bench import HBM, time_fn  # noqa: E402


def main():
    quick = "--quick" in sys.argv
    local = int(os.environ.get("LOCAL_RANK", 0))
    rank = int(os.environ.get("RANK", 0))
    torch.cuda.set_device(local)
    out = (lambda d: print(json.dumps(d), flush=True)) if rank == 0 else (lambda d: None)
    out(dict(device=torch.cuda.get_device_name(local), hbm_peak_gbs=HBM, world=int(os.environ.get("WORLD_SIZE", 1))))
    only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None
    # ---- spectral ordering
    for G in ((64, 128, 256, 512) if only in (None, "spectral") else ()):
        for B in ((1, 32, 256) if quick else (1, 32, 256, 4096)):
            if G >= 256 and B > 256:
                continue  # > 1 s of eigensolver work and GBs of workspace: outside the sweep's budget
            center = torch.rand(B, G, 3, device="cuda")
            fn = lambda: ops.spectral_eig(center, 20, 10.0, True, False, True, 4, True)
            t = time_fn([fn], iters=5 if B * G > 8192 else 20, warmup=2)
            out(dict(kernel="spectral_eig", G=G, B=B, us=round(t * 1e6, 1), clouds_per_s=round(B / t, 1)))
    # ---- selective scan
    D = 768
    for dtype in ((torch.float32, torch.bfloat16) if only in (None, "scan") else ()):
        es = 4 if dtype == torch.float32 else 2
        for L in (64, 256, 1024, 4096):
            for B in ((1, 32, 256) if quick else (1, 8, 32, 256, 1024, 4096)):
                E = B * L * D
                if 4 * E * es > 8e9:
                    continue
                nsets = max(1, min(4, int(400e6 // (4 * E * es)) + 1))
                sets = []
                for i in range(nsets):
                    g = torch.Generator(device="cuda").manual_seed(i)
                    r = lambda *s: torch.randn(*s, generator=g, device="cuda")
                    xz, u, dl, xd = r(B, L, 2 * D).to(dtype), r(B, L, D).to(dtype), (0.5 * r(B, L, D)).to(dtype), r(B, L, 56).to(dtype)
                    sets.append((u, dl, xd[..., 24:40], xd[..., 40:], xz[..., D:], torch.empty(B, L, D, dtype=dtype, device="cuda")))
                A = -torch.arange(1, 17, device="cuda", dtype=torch.float32).repeat(D, 1) * (1 + 0.1 * torch.rand(D, 16, device="cuda"))
                Dv, bias = torch.ones(D, device="cuda"), torch.full((D,), -4.0, device="cuda")
                fns = [(lambda s=s: ops.selective_scan_tm(s[0], s[1], A, s[2], s[3], Dv, s[4], bias, True, out=s[5])) for s in sets]
                t = time_fn(fns, iters=20 if E < 2e8 else 4, warmup=2)
                alg = 4 * E * es + 2 * B * L * 16 * es
                out(dict(kernel="selective_scan_fwd", dtype=str(dtype).split(".")[-1], L=L, B=B, us=round(t * 1e6, 1),
                         GBps=round(alg / t / 1e9, 1), frac_of_measured_hbm=round(alg / t / 1e9 / HBM, 3)))
                del sets, fns
                torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
