"""Timing of causal_conv1d forward at the C1 layer shape: fp32, fp32 + split planes, bf16 (SIM_CONV_TC sweeps the chunk)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from si_mamba_b200 import ops  # noqa: E402
from tools.kernel_bench import time_fn  # noqa: E402
B, L, D = 32, 512, 768
w, b = torch.randn(D, 4, device="cuda"), torch.randn(D, device="cuda")
for name, dt, split in (("fp32", torch.float32, False), ("fp32+planes", torch.float32, True), ("bf16", torch.bfloat16, False)):
    xs = [torch.randn(B, L, 2 * D, device="cuda").to(dt) for _ in range(4)]
    t = time_fn([(lambda x=x: ops.causal_conv1d_tm(x[..., :D], w, b, True, split=split)) for x in xs])
    print(f"{name:12s} {t * 1e6:6.1f} us", flush=True)
