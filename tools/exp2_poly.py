"""Coefficients of the degree-N polynomial of 2^f on [-0.5, 0.5] used by ex2_poly2 (selective_scan_fwd_ws.cu),
minimising the maximum RELATIVE error (iteratively re-weighted least squares on a dense grid = discrete Remez),
then checked in simulated fp32 Horner arithmetic exactly as the kernel evaluates it."""
import sys
import numpy as np

deg = int(sys.argv[1]) if len(sys.argv) > 1 else 5
x = np.linspace(-0.5, 0.5, 20001)
y = np.exp2(x)
w = np.ones_like(x)
V = np.vander(x, deg + 1, increasing=True)
for _ in range(200):
    c, *_ = np.linalg.lstsq(V * (w / y)[:, None], w, rcond=None)
    err = np.abs(V @ c / y - 1)
    w = w * (1 + 5 * err / err.max())
    w /= w.max()
print("float64 max rel err", err.max())
c32 = c.astype(np.float32)
xs = np.random.default_rng(0).uniform(-0.5, 0.5, 2_000_000).astype(np.float32)
p = np.full_like(xs, c32[deg])
for k in range(deg - 1, -1, -1):
    p = (p * xs + c32[k]).astype(np.float32)  # numpy has no fma: this is an upper bound on the kernel's error
ref = np.exp2(xs.astype(np.float64))
print("fp32 Horner max rel err", np.max(np.abs(p.astype(np.float64) / ref - 1)))
print(", ".join(f"c{k} = {float(v)!r}f" for k, v in enumerate(c32)))
