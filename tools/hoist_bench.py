"""Timing of the r02 inference choices of the fp32 mixer at the C1 layer shape (B = 32, L = 512, D = 768, d_model = 384):
in_proj on three bf16 planes vs two fp16 planes (with / without the silu epilogue), dt_proj with / without the softplus
epilogue, and the scan with its activations hoisted into those epilogues, per scan variant.  One JSON line per case.

    python tools/hoist_bench.py [--vlist 5008,5108,5208]
"""

from __future__ import annotations

import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))

from kernel_bench import HBM, time_fn  # noqa: E402
from si_mamba_b200 import ops  # noqa: E402


def main():
    torch.cuda.set_device(0)
    B, L, D, C = 32, 512, 768, 384
    M = B * L
    g = torch.Generator(device="cuda").manual_seed(0)
    r = lambda *s: torch.randn(*s, generator=g, device="cuda")
    nset = 4
    hs = [r(M, C) for _ in range(nset)]
    w_in = r(2 * D, C) * C ** -0.5
    outs = [torch.empty(M, 2 * D, device="cuda") for _ in range(nset)]
    for fmt, split in (("bf16x3", ops.split3), ("f16x2", ops.split2h)):
        xs, ws = [split(h) for h in hs], split(w_in)
        for act in (None, "silu_from"):
            fns = [(lambda x=x, o=o: ops.linear_split3(x, ws, C, out=o, act=act, act_col0=D)) for x, o in zip(xs, outs)]
            t = time_fn(fns)
            print(json.dumps(dict(kernel="in_proj", fmt=fmt, act=act, us=round(t * 1e6, 2),
                                  tflops_fp32_equiv=round(2 * M * C * 2 * D / t / 1e12, 1))), flush=True)
    # out_proj: scan output (unbounded: three bf16 planes) . W_out^T, K = 768, N = 384
    ys = [ops.split3(r(M, D)) for _ in range(nset)]
    w_out = ops.split3(r(C, D) * D ** -0.5)
    oouts = [torch.empty(M, C, device="cuda") for _ in range(nset)]
    t = time_fn([(lambda x=x, o=o: ops.linear_split3(x, w_out, D, out=o)) for x, o in zip(ys, oouts)])
    print(json.dumps(dict(kernel="out_proj", fmt="bf16x3", us=round(t * 1e6, 2),
                          tflops_fp32_equiv=round(2 * M * C * D / t / 1e12, 1))), flush=True)
    del ys, oouts
    # dt_proj: K = 32 (24 zero-padded), N = 768
    dl = [ops.split3(r(M, 32)) for _ in range(nset)]
    wdt = ops.split3(r(D, 32) * 0.2)
    bias = r(D) - 3.0
    douts = [torch.empty(M, D, device="cuda") for _ in range(nset)]
    for act in (None, "softplus_bias"):
        fns = [(lambda x=x, o=o: ops.linear_split3(x, wdt, 32, out=o, act=act, bias=bias)) for x, o in zip(dl, douts)]
        t = time_fn(fns)
        print(json.dumps(dict(kernel="dt_proj", act=act, us=round(t * 1e6, 2), GBps=round(M * D * 4 / t / 1e9, 1))), flush=True)
    # scan: who evaluates softplus / silu
    variants = (5008,)
    if "--vlist" in sys.argv:
        variants = tuple(int(v) for v in sys.argv[sys.argv.index("--vlist") + 1].split(","))
    sets = []
    for i in range(nset):
        xz, u, delta, xdbl = r(B, L, 2 * D), r(B, L, D), 0.5 * r(B, L, D), r(B, L, 56)
        sets.append((u, delta, xdbl[..., 24:40], xdbl[..., 40:], xz[..., D:], torch.empty(B, L, D, device="cuda")))
    A = -torch.arange(1, 17, device="cuda", dtype=torch.float32).repeat(D, 1) * (1 + 0.1 * torch.rand(D, 16, device="cuda"))
    Dv = torch.ones(D, device="cuda")
    sb = torch.full((D,), -4.0, device="cuda")
    alg = 4 * B * L * D * 4 + 2 * B * L * 16 * 4
    for variant in variants:
        for mode in ("scan", "z", "zdt"):
            zg = mode != "scan"
            fin = mode == "zdt"
            fns = [(lambda s=s: ops.selective_scan_tm(s[0], s[1], A, s[2], s[3], Dv, s[4], None if fin else sb, not fin,
                                                      out=s[5], variant=variant, z_gate=zg)) for s in sets]
            t = time_fn(fns)
            print(json.dumps(dict(kernel="selective_scan_fwd", activations_in=mode, variant=variant, us=round(t * 1e6, 2),
                                  frac_of_measured_hbm=round(alg / t / 1e9 / HBM, 3))), flush=True)


if __name__ == "__main__":
    main()
