"""Per-kernel timing on the GPU box (CUDA events on the launching stream, warm-up, L2 defeated by
rotating over input sets whose total size exceeds the 126 MB L2).  Prints one JSON line per kernel:

    python tools/kernel_bench.py [--quick]

Algorithmic bytes follow SURVEY.md section 8(d): scan fwd = 4*E*s + 2*S*s, conv fwd = 2*E*s,
order-gather = (G + 2kG)*C*B*s, add+LN = (2 reads + 2 writes)*rows*C*4.
"""

from __future__ import annotations

import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from si_mamba_b200 import ops  # noqa: E402

PEAKS = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
HBM = PEAKS.get("hbm_gbs", 6650.0)


def time_fn(fns, iters=20, warmup=5):
    """fns: list of callables rotated round-robin (different buffers => cold L2)."""
    for i in range(warmup):
        fns[i % len(fns)]()
    torch.cuda.synchronize()
    # capture the launch sequence in a CUDA graph so host-side launch overhead (ctypes, allocator) is not timed
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for i in range(iters):
            fns[i % len(fns)]()
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (iters * reps) * 1e-3


def report(name, secs, bytes_, **extra):
    gbs = bytes_ / secs / 1e9
    print(json.dumps(dict(kernel=name, us=round(secs * 1e6, 2), alg_MB=round(bytes_ / 1e6, 2), GBps=round(gbs, 1),
                          frac_of_measured_hbm=round(gbs / HBM, 3), **extra)), flush=True)


def scan_case(B, L, D, dtype, variant, nsets):
    sets = []
    for i in range(nsets):
        g = torch.Generator(device="cuda").manual_seed(i)
        r = lambda *s: torch.randn(*s, generator=g, device="cuda")
        xz = r(B, L, 2 * D).to(dtype)
        u = r(B, L, D).to(dtype)
        delta = (0.5 * r(B, L, D)).to(dtype)
        xdbl = r(B, L, 56).to(dtype)
        out = torch.empty(B, L, D, dtype=dtype, device="cuda")
        sets.append((u, delta, xdbl[..., 24:40], xdbl[..., 40:], xz[..., D:], out))
    A = -torch.arange(1, 17, device="cuda", dtype=torch.float32).repeat(D, 1) * (1 + 0.1 * torch.rand(D, 16, device="cuda"))
    Dv = torch.ones(D, device="cuda")
    bias = torch.full((D,), -4.0, device="cuda")
    fns = [(lambda s=s: ops.selective_scan_tm(s[0], s[1], A, s[2], s[3], Dv, s[4], bias, True, out=s[5],
                                              variant=variant)) for s in sets]
    es = 4 if dtype == torch.float32 else 2
    alg = 4 * B * L * D * es + 2 * B * L * 16 * es
    return fns, alg


def main():
    quick = "--quick" in sys.argv
    only_f32 = "--f32" in sys.argv
    only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None
    want = lambda name: only is None or only == name
    torch.cuda.set_device(0)
    print(json.dumps(dict(device=torch.cuda.get_device_name(0), hbm_peak_gbs=HBM)), flush=True)
    shapes = [(32, 512, 768), (256, 512, 768)] if not quick else [(32, 512, 768)]
    for (B, L, D) in (shapes if want("scan") else []):
        for dtype in ((torch.float32,) if only_f32 else (torch.float32, torch.bfloat16)):
            variants = (108, 5008) if "--variants" in sys.argv else (0,)
            if "--vlist" in sys.argv:
                variants = tuple(int(v) for v in sys.argv[sys.argv.index("--vlist") + 1].split(","))
            for variant in variants:
                nsets = max(2, int(300e6 // (4 * B * L * D * (4 if dtype == torch.float32 else 2))) + 1)
                fns, alg = scan_case(B, L, D, dtype, variant, min(nsets, 4))
                t = time_fn(fns)
                report("selective_scan_fwd", t, alg, B=B, L=L, D=D, dtype=str(dtype).split(".")[-1], variant=variant)
                del fns
                torch.cuda.empty_cache()
    # scan backward (training configs): C1 layer shape fp32, C2 layer shape (L = 1024) bf16
    for (B, L, D, dtype) in (((32, 512, 768, torch.float32), (32, 1024, 768, torch.bfloat16)) if want("scanbwd") else []):
        fns, alg = scan_case(B, L, D, dtype, 0, 2)
        es = 4 if dtype == torch.float32 else 2
        sets = []
        for i in range(2):
            g = torch.Generator(device="cuda").manual_seed(10 + i)
            r = lambda *s: torch.randn(*s, generator=g, device="cuda")
            xz, u, delta, xdbl = r(B, L, 2 * D).to(dtype), r(B, L, D).to(dtype), (0.5 * r(B, L, D)).to(dtype), r(B, L, 56).to(dtype)
            dout = r(B, L, D).to(dtype)
            ck = torch.empty(ops.scan_checkpoint_shape(B, L, D), dtype=torch.float32, device="cuda")
            sets.append((u, delta, xdbl[..., 24:40], xdbl[..., 40:], xz[..., D:], dout, ck))
        A = -torch.arange(1, 17, device="cuda", dtype=torch.float32).repeat(D, 1)
        Dv, bias = torch.ones(D, device="cuda"), torch.full((D,), -4.0, device="cuda")
        for s_ in sets:
            ops.selective_scan_tm(s_[0], s_[1], A, s_[2], s_[3], Dv, s_[4], bias, True, checkpoints=s_[6])
        bfn = [(lambda s_=s_: ops.selective_scan_bwd_tm(s_[0], s_[1], A, s_[2], s_[3], Dv, s_[4], bias, s_[5], s_[6], True))
               for s_ in sets]
        E, S = B * L * D, B * L * 16
        alg_b = 7 * E * es + 2 * S * es + 2 * S * 4
        report("selective_scan_bwd", time_fn(bfn), alg_b, B=B, L=L, D=D, dtype=str(dtype).split(".")[-1],
               note="includes the zero-fills of dB / dC / dA and the output allocations")
        ffn = [(lambda s_=s_: ops.selective_scan_tm(s_[0], s_[1], A, s_[2], s_[3], Dv, s_[4], bias, True, checkpoints=s_[6]))
               for s_ in sets]
        report("selective_scan_fwd+ckpt", time_fn(ffn), 4 * E * es + 2 * S * es + E * 4, B=B, L=L, D=D,
               dtype=str(dtype).split(".")[-1])
        del sets, bfn, ffn
        torch.cuda.empty_cache()
    # conv
    for (B, L, D) in (shapes[:2] if want("conv") else []):
        for dtype in (torch.float32, torch.bfloat16):
            xs = [torch.randn(B, L, 2 * D, device="cuda").to(dtype) for _ in range(4)]
            outs = [torch.empty(B, L, D, device="cuda", dtype=dtype) for _ in range(4)]
            w, b = torch.randn(D, 4, device="cuda"), torch.randn(D, device="cuda")
            fns = [(lambda x=x, o=o: ops.causal_conv1d_tm(x[..., :D], w, b, True, out=o)) for x, o in zip(xs, outs)]
            es = 4 if dtype == torch.float32 else 2
            report("causal_conv1d_fwd", time_fn(fns), 2 * B * L * D * es, B=B, L=L, D=D, dtype=str(dtype).split(".")[-1])
    if not want("rows") and not want("tok"):
        return
    # add + layernorm
    B, L, C = 32, 512, 384
    xs = [(torch.randn(B, L, C, device="cuda"), torch.randn(B, L, C, device="cuda")) for _ in range(6)]
    w, b = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    fns = [(lambda x=x, r=r: ops.add_layernorm(x, r, w, b)) for x, r in xs]
    report("add_layernorm", time_fn(fns), 4 * B * L * C * 4, rows=B * L, C=C, note="includes 2 torch.empty allocations")
    # order gather (tokens + pos fused)
    G, k = 64, 4
    tok = [(torch.randn(B, G, C, device="cuda"), torch.randn(B, G, C, device="cuda")) for _ in range(4)]
    perm = torch.stack([torch.stack([torch.randperm(G, device="cuda") for _ in range(k)]) for _ in range(B)]).int()
    fns = [(lambda t=t, p=p: ops.order_gather_add(t, p, perm, True)) for t, p in tok]
    report("order_gather_add", time_fn(fns), (2 * G + 2 * k * G) * C * B * 4, B=B, G=G, k=k)
    # tokenizer + spectral: latency-bound, report time only
    for (N, G) in ((1024, 64), (2048, 128)):
        xyz = torch.rand(B, N, 3, device="cuda")
        t = time_fn([lambda: ops.fps(xyz, G)])
        print(json.dumps(dict(kernel="fps", us=round(t * 1e6, 2), B=B, N=N, G=G)), flush=True)
        center, _ = ops.fps(xyz, G)
        t = time_fn([lambda: ops.knn_group(xyz, center, 32)])
        print(json.dumps(dict(kernel="knn_group", us=round(t * 1e6, 2), B=B, N=N, G=G, M=32)), flush=True)
        t = time_fn([lambda: ops.spectral_eig(center, 20, 100.0, True, False, True, 4, True)])
        print(json.dumps(dict(kernel="spectral_eig", us=round(t * 1e6, 2), B=B, G=G, k=4)), flush=True)


if __name__ == "__main__":
    main()
