"""Accuracy (vs fp64) and speed (vs cuBLAS fp32 / TF32) of the tcgen05 fp32 GEMM on the four Mamba projection shapes."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from si_mamba_b200 import ops  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
M = 32 * 512
g = torch.Generator(device="cuda").manual_seed(0)


def timeit(f, n=20):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for name, N, K, lda in (("in_proj", 1536, 384, 384), ("x_proj", 56, 768, 768), ("dt_proj", 768, 24, 56), ("out_proj", 384, 768, 768)):
    xb = torch.randn(M, lda, generator=g, device="cuda")
    x = xb[:, :K]
    w = torch.randn(N, K, generator=g, device="cuda") * K ** -0.5
    ref = (x.double() @ w.double().t())
    y32 = torch.nn.functional.linear(x, w)
    torch.backends.cuda.matmul.allow_tf32 = True
    ytf = torch.nn.functional.linear(x, w)
    t_tf = timeit(lambda: torch.nn.functional.linear(x, w))
    torch.backends.cuda.matmul.allow_tf32 = False
    err = lambda a: ((a.double() - ref).abs().max() / ref.abs().max()).item()
    xs, ws = ops.split3(x), ops.split3(w)
    y3 = ops.linear_split3(xs, ws, K)
    t_x3 = timeit(lambda: ops.linear_split3(xs, ws, K))
    t_sp = timeit(lambda: ops.split3(x, out=xs))
    t_32 = timeit(lambda: torch.nn.functional.linear(x, w))
    fl = 2.0 * M * N * K
    print(f"{name:9s} N={N:5d} K={K:4d}  err cublas-fp32 {err(y32):.2e}  tf32 {err(ytf):.2e} | "
          f"us fp32 {t_32:7.1f} ({fl / t_32 * 1e-6:6.1f} TF)  tf32 {t_tf:7.1f} | "
          f"x3 err {err(y3):.2e} us {t_x3:7.1f} ({6 * fl / t_x3 * 1e-6:6.1f} bf16-TF)  split {t_sp:6.1f}", flush=True)
