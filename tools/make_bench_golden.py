#!/usr/bin/env python
"""Oracle outputs for the inputs bench.py itself times: BASELINE.json config C1 at its own batch
(32 clouds of 1024 points, the first rotating input set of rank 0) and the C2 shape (2048 points, 128 patches) at batch 32,
both with the seeded random-init weights bench.py builds.  Run here on the CPU (tens of seconds); commits
tests/golden/bench_c1.pt and bench_c2.pt, which tests/test_gpu_model.py::test_bench_inputs_match_oracle checks the CUDA
forward against at full batch.

    python tools/make_bench_golden.py
"""
from __future__ import annotations

import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import bench  # noqa: E402
import si_mamba_b200 as sm  # noqa: E402
from oracle import model as omodel  # noqa: E402


def one(name, cfg, n_points, batch):
    torch.manual_seed(0)
    model = sm.PointMamba(cfg).eval()
    sd = model.state_dict()
    pts = bench.make_clouds(batch, 0, 1, n_points=n_points)[0]
    t = time.perf_counter()
    logits, inter = omodel.point_mamba_forward(sd, dict(cfg), pts, return_intermediates=True)
    print(f"{name}: oracle forward of {batch} clouds in {time.perf_counter() - t:.1f} s")
    out = dict(logits=logits.float(), perm=inter["perm"].to(torch.int16), eigvecs=inter["eigvecs"].double(),
               center=inter["center"].float(), n_points=n_points, batch=batch, seed="bench.make_clouds(batch, 0, 1)[0]")
    torch.save(out, ROOT / "tests" / "golden" / f"bench_{name}.pt")


def main():
    torch.set_num_threads(max(1, torch.get_num_threads()))
    one("c1", sm.finetune_modelnet(), 1024, 32)
    one("c2", sm.finetune_scan_hardest(), 2048, 32)


if __name__ == "__main__":
    main()
