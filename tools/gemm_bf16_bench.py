"""bf16 projection GEMMs of one Mamba layer (forward, dgrad, wgrad) on the hand-written tcgen05 kernel vs torch (cuBLAS):
CUDA events around a CUDA graph of 20 launches on rotating operands.

    python tools/gemm_bf16_bench.py [--rows 16384]
"""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from si_mamba_b200 import ops  # noqa: E402
from tools.kernel_bench import time_fn  # noqa: E402


def main():
    M = int(sys.argv[sys.argv.index("--rows") + 1]) if "--rows" in sys.argv else 16384
    dev = "cuda"
    for name, N, K in (("in_proj", 1536, 384), ("x_proj", 56, 768), ("dt_proj", 768, 24), ("out_proj", 384, 768)):
        nset = 4
        xs = [torch.randn(M, K, device=dev).to(torch.bfloat16) for _ in range(nset)]
        dys = [torch.randn(M, N, device=dev).to(torch.bfloat16) for _ in range(nset)]
        w = (torch.randn(N, K, device=dev) * K ** -0.5).to(torch.bfloat16)
        flops = 2.0 * M * N * K
        rows = {}
        rows["fwd"] = (time_fn([(lambda x=x: ops.gemm_bf16(x, w)) for x in xs]),
                       time_fn([(lambda x=x: torch.nn.functional.linear(x, w)) for x in xs]))
        rows["dgrad"] = (time_fn([(lambda d=d: ops.gemm_bf16(d, w, b_mn=True)) for d in dys]),
                         time_fn([(lambda d=d: d @ w) for d in dys]))
        rows["wgrad"] = (time_fn([(lambda d=d, x=x: ops.gemm_bf16(d, x, a_mn=True, b_mn=True, splits=0)) for d, x in zip(dys, xs)]),
                         time_fn([(lambda d=d, x=x: (d.t() @ x).float()) for d, x in zip(dys, xs)]))
        for k, (own, lib) in rows.items():
            print(json.dumps(dict(gemm=name, op=k, M=M, N=N, K=K, own_us=round(own * 1e6, 2), cublas_us=round(lib * 1e6, 2),
                                  own_tflops=round(flops / own / 1e12, 1))), flush=True)


if __name__ == "__main__":
    main()
