"""Times the fused conv + x_proj kernel against the unfused pair on the C1 shape (B=32, L=512, d_inner=768)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from si_mamba_b200 import ops

B, L, D, N = 32, 512, 768, 56
xz = torch.randn(B, L, 2 * D, device="cuda")
x = xz[..., :D]
cw = torch.randn(D, 1, 4, device="cuda") * 0.5
cb = torch.randn(D, device="cuda") * 0.1
ws = ops.split3(torch.randn(N, D, device="cuda") * D ** -0.5)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def unfused():
    u = ops.causal_conv1d_tm(x, cw, cb, silu=True)
    return ops.linear_f32a_planes_out(u, ws, D, 32)


print("unfused conv + x_proj: %.1f us" % timeit(unfused))
print("fused conv_xproj:      %.1f us" % timeit(lambda: ops.conv_xproj_f32(x, cw, cb, ws, 32)))
