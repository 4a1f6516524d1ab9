import sys, torch
sys.path.insert(0, '/root/repo')
from si_mamba_b200 import ops
from tools.kernel_bench import time_fn
M = 16384
for K in (24, 32, 64, 128):
    x = torch.randn(M, K, device="cuda"); w = torch.randn(768, K, device="cuda")
    xs, ws = ops.split3(x), ops.split3(w)
    outs = [torch.empty(M, 768, device="cuda") for _ in range(4)]
    t = time_fn([(lambda o=o: ops.linear_split3(xs, ws, K, out=o)) for o in outs])
    print(f"N=768 K={K}: {t*1e6:.1f} us", flush=True)
for N in (192, 384, 768, 1536):
    K = 32
    x = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda")
    xs, ws = ops.split3(x), ops.split3(w)
    outs = [torch.empty(M, N, device="cuda") for _ in range(4)]
    t = time_fn([(lambda o=o: ops.linear_split3(xs, ws, K, out=o)) for o in outs])
    print(f"N={N} K=32: {t*1e6:.1f} us  ({M*N*4/t/1e9:.0f} GB/s written)", flush=True)
# plain copy of the same output size for reference
a = [torch.randn(M, 768, device="cuda") for _ in range(4)]
b = [torch.empty(M, 768, device="cuda") for _ in range(4)]
t = time_fn([(lambda i=i: b[i].copy_(a[i])) for i in range(4)])
print(f"copy 50 MB: {t*1e6:.1f} us")
