#!/usr/bin/env python
"""Headline benchmark: point clouds/s of the SI-Mamba-cls forward (BASELINE.json config[0] shape: 1024 points,
64 patches x 32, depth 12, d=384, batch 32 per GPU) + achieved HBM GB/s of the selective-scan kernel.

    python bench.py --gpus N --steps K --warmup W                 # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K --warmup W  # CPU oracle port of the reference path

One process per GPU (torchrun for N > 1, NCCL only for the timing barrier / max-over-ranks: the forward
path shards clouds across GPUs with no data-path collective).  Prints ONE JSON line on rank 0.
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

METRIC = "point clouds/sec SI-Mamba-cls fwd (1024pts)"
UNIT = "clouds/s"
N_POINTS, N_SETS = 1024, 8


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="clouds per GPU per step (config[0]: 32)")
    ap.add_argument("--precision", default=os.environ.get("SIM_PRECISION", "fp32"), choices=["fp32", "tf32", "bf16"],
                    help="fp32 = the reference finetune/test precision (no autocast); bf16 = autocast as in runner_pretrain")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--ref-sample", type=int, default=1, help="clouds per step of the CPU reference arm")
    return ap.parse_args()


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def synthetic_clouds(B, N, seed):
    """Synthetic "surface" clouds of SURVEY.md section 8(d): a few anisotropic blobs with noise, normalised like the
    datasets (datasets/ModelNetDataset.py:52-57: centroid 0, max-norm 1).  bench.py's own generator - the timed arm
    imports nothing from oracle/; tests/test_abi_host.py checks it draws the same clouds as the tests' generator."""
    g = torch.Generator().manual_seed(seed)
    pts = torch.empty(B, N, 3)
    for b in range(B):
        n_patch = int(torch.randint(3, 7, (1,), generator=g))
        which = torch.randint(0, n_patch, (N,), generator=g)
        ctr = torch.randn(n_patch, 3, generator=g) * 0.5
        axes = torch.rand(n_patch, 3, generator=g) * 0.6 + 0.1
        v = torch.randn(N, 3, generator=g)
        v = v / v.norm(dim=-1, keepdim=True)
        pts[b] = ctr[which] + v * axes[which] + 0.01 * torch.randn(N, 3, generator=g)
    pts = pts - pts.mean(dim=1, keepdim=True)
    pts = pts / pts.norm(dim=-1).max(dim=1).values[:, None, None]
    return pts.contiguous().float()


def make_clouds(batch, rank, sets=N_SETS):
    return [synthetic_clouds(batch, N_POINTS, 1234 + 1000 * 1 + 17 * rank + i) for i in range(sets)]


_JSON_OUT = None


def claim_stdout():
    """One JSON line on stdout, nothing else: from here on anything a library writes to fd 1 (NCCL's version banner,
    torchrun children ...) lands on stderr, and emit() writes to the real stdout."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(obj):
    claim_stdout()
    _JSON_OUT.write(json.dumps(obj) + "\n")
    _JSON_OUT.flush()


def build_model():
    import si_mamba_b200 as sm
    torch.manual_seed(0)
    cfg = sm.finetune_modelnet()
    return sm.PointMamba(cfg).eval(), cfg


# ----------------------------------------------------------------------------- clock sampler (NVML)
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.stop_flag, self.ok = [], set(), False, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:  # noqa: BLE001
            self.max_mhz = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {getattr(nv, k): k for k in dir(nv) if k.startswith("nvmlClocksThrottleReason") and
                 isinstance(getattr(nv, k), int) and getattr(nv, k) not in (0,)}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit and "GpuIdle" not in name and "None" not in name and "All" not in name:
                        self.reasons.add(name.replace("nvmlClocksThrottleReason", ""))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.02)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ----------------------------------------------------------------------------- reference (CPU) arm
def cpu_reference_rate(sample_clouds, repeats=3):
    """Oracle port of the reference forward (oracle/model.py) on the host cores: clouds/s, median of `repeats`."""
    from oracle import model as omodel
    model, cfg = build_model()
    sd = model.state_dict()
    torch.set_num_threads(os.cpu_count() or 1)
    pts = make_clouds(sample_clouds, 0, 1)[0]
    omodel.point_mamba_forward(sd, dict(cfg), pts[:1])  # warm-up
    ts = []
    for _ in range(repeats):
        t = time.perf_counter()
        omodel.point_mamba_forward(sd, dict(cfg), pts)
        ts.append(time.perf_counter() - t)
    ts.sort()
    return sample_clouds / ts[len(ts) // 2], torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import model as omodel
    model, cfg = build_model()
    sd = model.state_dict()
    torch.set_num_threads(os.cpu_count() or 1)
    sets = make_clouds(args.ref_sample, 0, 4)
    for i in range(args.warmup):
        omodel.point_mamba_forward(sd, dict(cfg), sets[i % 4])
    t0 = time.perf_counter()
    for i in range(args.steps):
        omodel.point_mamba_forward(sd, dict(cfg), sets[i % 4])
    dt = time.perf_counter() - t0
    val = args.ref_sample * args.steps / dt
    cores = torch.get_num_threads()
    sample = f"{args.ref_sample} cloud(s) of the same workload per step, {args.steps} steps"
    emit({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": f"C1 SI-Mamba-cls forward: 1024 pts, 64 patches x 32, L=512, depth 12, d=384, batch "
                               f"{args.batch} per GPU, eval mode, random-init weights",
                   "arm": "CPU oracle port of the reference path (its CUDA-only wheels cannot run on the host cores); "
                          f"each step is a bounded sample of {args.ref_sample} cloud(s) of that workload"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


# ----------------------------------------------------------------------------- our arm
def run_ours(args):
    from si_mamba_b200 import _lib, ops

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert torch.cuda.is_available(), "bench.py (impl=ours) needs a CUDA device: there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    # fp32 mirrors the reference's runtime defaults (no autocast in runner_finetune): fp32 matmuls, while
    # cuDNN convolutions (the Encoder's 1x1 convs) keep torch's default allow_tf32=True
    torch.backends.cuda.matmul.allow_tf32 = args.precision == "tf32"
    torch.backends.cudnn.allow_tf32 = True
    model, cfg = build_model()
    model = model.to(dev)
    B = args.batch
    host_sets = [t.pin_memory() for t in make_clouds(B, rank)]
    dev_sets = [t.to(dev) for t in host_sets]
    static_in = torch.empty_like(dev_sets[0])
    out_host = [torch.empty(B, cfg.cls_dim).pin_memory() for _ in range(2)]

    def fwd(x):
        with torch.no_grad():
            if args.precision == "bf16":
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    return model(x).float()
            return model(x)

    # eager warm-up (also sets kernel attributes, cuBLAS handles, allocator pools)
    static_in.copy_(dev_sets[0])
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            static_out = fwd(static_in)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    n0 = _lib.launches()
    fwd(static_in)
    launches_per_step = _lib.launches() - n0

    graph = None
    if not args.no_graph:
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            static_out = fwd(static_in)

    def step_resident(i):
        static_in.copy_(dev_sets[i % N_SETS])
        if graph is not None:
            graph.replay()
            return static_out
        return fwd(static_in)

    def step_e2e(i):
        static_in.copy_(host_sets[i % N_SETS], non_blocking=True)
        o = step_resident_nocopy()
        out_host[i % 2].copy_(o, non_blocking=True)

    def step_resident_nocopy():
        if graph is not None:
            graph.replay()
            return static_out
        return fwd(static_in)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, steps, warmup):
        for i in range(warmup):
            step_fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            step_fn(warmup + i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_res = timed(step_resident, args.steps, max(args.warmup, 3))
    ms_e2e = timed(step_e2e, args.steps, max(args.warmup, 3))
    sampler.stop_flag = True

    # ---- dominant own kernel: the selective scan, on the real layer shapes, inputs rotated so they miss L2
    L = 2 * cfg.k_top_eigenvectors * cfg.num_group
    Dm = 2 * cfg.trans_dim
    adt = torch.bfloat16 if args.precision == "bf16" else torch.float32
    es = 2 if adt == torch.bfloat16 else 4
    nsets = max(2, int(300e6 // (4 * B * L * Dm * es)) + 1)
    g = torch.Generator(device=dev).manual_seed(7)
    scan_sets = []
    for _ in range(nsets):
        xz = torch.randn(B, L, 2 * Dm, generator=g, device=dev).to(adt)
        u = torch.randn(B, L, Dm, generator=g, device=dev).to(adt)
        dl = (0.5 * torch.randn(B, L, Dm, generator=g, device=dev)).to(adt)
        xd = torch.randn(B, L, 56, generator=g, device=dev).to(adt)
        scan_sets.append((u, dl, xd[..., 24:40], xd[..., 40:], xz[..., Dm:], torch.empty(B, L, Dm, dtype=adt, device=dev)))
    mix = model.blocks.layers[0].mixer
    A = -torch.exp(mix.A_log.float())
    def scan_call(s):
        ops.selective_scan_tm(s[0], s[1], A, s[2], s[3], mix.D, s[4], mix.dt_proj.bias, True, out=s[5])
    # CUDA events around a captured graph of back-to-back launches on this stream: the average is the kernel's launch
    # duration, not the Python / tensor-map-encode time of an eager loop (which is longer than the kernel itself)
    scan_iters, reps = 20, 3
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(nsets):
            scan_call(scan_sets[i])
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    sgraph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(sgraph):
        for i in range(scan_iters):
            scan_call(scan_sets[i % nsets])
    sgraph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        sgraph.replay()
    e1.record()
    torch.cuda.synchronize()
    scan_s = e0.elapsed_time(e1) / (scan_iters * reps) * 1e-3
    alg_bytes = 4 * B * L * Dm * es + 2 * B * L * 16 * es
    peak, peak_src = peaks()
    achieved = alg_bytes / scan_s / 1e9
    traffic = None
    tf = ROOT / "profiles" / "scan_traffic.json"
    if tf.exists():
        traffic = json.loads(tf.read_text()).get(f"{args.precision}_B{B}")

    if rank != 0:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
        return
    # reported baseline, rank 0 at N=1 only (the scaling runs would only repeat it)
    cpu_base = None
    if world == 1:
        cpu_val, cores = cpu_reference_rate(2)
        cpu_base = {"value": cpu_val, "unit": UNIT, "cores": cores, "kind": "port",
                    "sample": "2 clouds of the same workload, median of 3 runs (oracle/model.py)"}
    clouds = B * world * args.steps
    line = {
        "metric": METRIC, "value": clouds / (ms_res * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": {"workload": f"C1 SI-Mamba-cls forward: 1024 pts, 64 patches x 32, L=512, depth 12, d=384, batch {B} "
                               f"per GPU, eval mode, random-init weights",
                   "parallelism": f"dp{world} (clouds sharded, no data-path collective)",
                   "l2": f"{N_SETS} rotating input batches; per-layer activations (3 x 50 MB fp32) exceed the 126 MB L2",
                   "cuda_graph": graph is not None},
        "e2e": {"value": clouds / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": B * N_POINTS * 3 * 4,
                "d2h_bytes_per_step": B * cfg.cls_dim * 4, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches_per_step * args.steps,
        "gpu_launches_per_step": launches_per_step,
        "clocks": sampler.summary(),
        "roofline": {"kernel": "selective_scan_fwd", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "alg_bytes_per_launch": alg_bytes, "us_per_launch": scan_s * 1e6,
                     "shape": {"B": B, "L": L, "D": Dm, "N": 16, "dtype": str(adt).split(".")[-1]},
                     "note": "general-A fp32 scan needs 20 MUFU ops per channel-step; at B200's 16 MUFU/clk/SM that alone is 54 us "
                             "= 0.58 of this roofline (DESIGN.md 4.1); timed inside a CUDA graph of 20 launches"},
        "cpu_baseline": cpu_base,
    }
    emit(line)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def main():
    args = parse()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
