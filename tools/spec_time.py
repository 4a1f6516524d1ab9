"""Time sim_spectral_eig at the C1 / C2 patch counts (B = 32 clouds, k = 4 eigenvectors, 20-NN binary graph)."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
from kernel_bench import time_fn  # noqa: E402
from si_mamba_b200 import ops  # noqa: E402

B = 32
for (N, G) in ((1024, 64), (2048, 128)):
    xyz = torch.rand(B, N, 3, device="cuda")
    center, _ = ops.fps(xyz, G)
    t = time_fn([lambda: ops.spectral_eig(center, 20, 100.0, True, False, True, 4, True)])
    print(json.dumps(dict(kernel="spectral_eig", B=B, G=G, us=round(t * 1e6, 1))))
