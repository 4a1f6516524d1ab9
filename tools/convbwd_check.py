"""Times the causal conv1d backward at the C2 layer shape (B=32, L=1024, D=768, bf16) and the C1 fp32 shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from si_mamba_b200 import ops

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for dtype, L in ((torch.bfloat16, 1024), (torch.float32, 512)):
    B, D = 32, 768
    xz = torch.randn(B, L, 2 * D, device="cuda").to(dtype)
    x = xz[..., :D].requires_grad_(True)
    w = (torch.randn(D, 4, device="cuda") * 0.5).requires_grad_(True)
    b = (torch.randn(D, device="cuda") * 0.1).requires_grad_(True)
    y = ops.CausalConv1dTM.apply(x, w, b, True)
    dy = torch.randn_like(y)
    ts = []
    for i in range(13):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.autograd.grad(y, (x, w, b), dy, retain_graph=True)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts = sorted(ts[3:])
    byts = 3 * B * L * D * xz.element_size()
    print(f"conv1d bwd {dtype} L={L}: {ts[len(ts)//2]:.1f} us (incl. autograd launch overhead), {byts/ts[len(ts)//2]/1e3:.0f} GB/s algorithmic")
