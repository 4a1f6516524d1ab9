"""x_proj at the C1 layer shape: pre-split planes (conv writes them) vs in-kernel split of the fp32 activation."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from si_mamba_b200 import ops  # noqa: E402
from tools.kernel_bench import time_fn  # noqa: E402
M, N, K = 16384, 56, 768
us = [torch.randn(M, K, device="cuda") for _ in range(4)]
ws = ops.split3(torch.randn(N, K, device="cuda") * K ** -0.5)
ps = [ops.split3(u) for u in us]
print(f"pre-split planes + planes out : {time_fn([(lambda p=p: ops.linear_split3_planes_out(p, ws, K, 32)) for p in ps]) * 1e6:6.1f} us")
print(f"fp32 A, in-kernel split       : {time_fn([(lambda u=u: ops.linear_f32a_planes_out(u, ws, K, 32)) for u in us]) * 1e6:6.1f} us")
