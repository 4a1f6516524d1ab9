#!/usr/bin/env python
"""Headline benchmark: point clouds/s of the SI-Mamba-cls forward (BASELINE.json config[0] shape: 1024 points,
64 patches x 32, depth 12, d=384, batch 32 per GPU) + achieved HBM GB/s of the selective-scan kernel.

    python bench.py --gpus N --steps K --warmup W                 # this repo's sm_100a path, headline workload c1
    python bench.py --impl reference --gpus N --steps K --warmup W  # CPU oracle port of the reference path
    python bench.py --workload c2|c3|c4 ...                       # BASELINE.json's training configs (whole step captured
                                                                  # in a CUDA graph; c3 / c4 all-reduce gradients over NCCL)

One process per GPU (torchrun for N > 1).  The forward path shards clouds across GPUs with no data-path collective (NCCL
only for the timing barrier / max-over-ranks); the training workloads add the gradient all-reduce, overlapped with the
backward and captured in the step's graph (si_mamba_b200/train.py).  Prints ONE JSON line on rank 0.
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

# BASELINE.json configs.  c1 is the headline (the driver's default command); c2-c4 are the training configs, measured by
# the same harness: `bench.py --workload c3 --gpus 8` under torchrun.
WORKLOADS = {
    "c1": dict(metric=METRIC, batch=32, points=1024, dtype="fp32",
               text="C1 SI-Mamba-cls forward: 1024 pts, 64 patches x 32, L=512, depth 12, d=384, batch {B} per GPU, eval mode, "
                    "random-init weights"),
    "c2": dict(metric="point clouds/sec SI-Mamba-cls training step (2048pts, bf16)", batch=32, points=2048, dtype="bf16",
               text="C2 ScanObjectNN-hardest finetune step: 2048 pts, 128 patches x 32, L=1024, depth 12, d=384, batch {B} per GPU, "
                    "bf16 autocast, forward + backward + clip_grad_norm(10) + AdamW, train mode, random-init weights"),
    "c3": dict(metric="point clouds/sec SI-Mamba MAE pre-training step (1024pts, bf16)", batch=16, points=1024, dtype="bf16",
               text="C3 MAE pre-training step (pretrain.yaml): 1024 pts, 64 patches x 32, mask 0.6, 12 + 4 layers, token restore, "
                    "Chamfer-L2, batch {B} per GPU (total_bs 128 on 8 GPUs), bf16 autocast, forward + backward + gradient "
                    "all-reduce + clip_grad_norm(10) + AdamW, random-init weights"),
    "c4": dict(metric="point clouds/sec SI-Mamba part-segmentation training step (2048pts)", batch=16, points=2048, dtype="fp32",
               text="C4 part segmentation step: 2048 pts, 128 patches x 32, HLT (recursive spectral partition) layout L=256, depth "
                    "12, batch {B} per GPU, fp32, forward + backward + gradient all-reduce + clip_grad_norm(10) + AdamW, "
                    "random-init weights"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c1", choices=sorted(WORKLOADS), help="BASELINE.json config (default: the headline c1)")
    ap.add_argument("--batch", type=int, default=None, help="clouds per GPU per step (default: the config's own, c1: 32)")
    ap.add_argument("--precision", default=os.environ.get("SIM_PRECISION", "fp32"), choices=["fp32", "tf32", "bf16"],
                    help="fp32 = the reference finetune/test precision (no autocast); bf16 = autocast as in runner_pretrain")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--ref-sample", type=int, default=None, help="clouds per step of the CPU reference arm (default: the batch)")
    ap.add_argument("--bucket-mb", type=float, default=16.0, help="gradient all-reduce bucket size (training workloads)")
    ap.add_argument("--no-overlap", action="store_true", help="training: all-reduce after the backward instead of overlapped")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.batch is None:
        args.batch = w["batch"]
    if args.ref_sample is None:
        args.ref_sample = args.batch
    if args.workload != "c1":
        args.precision = w["dtype"]
    return args


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


def make_clouds(batch, rank, sets=N_SETS, n_points=N_POINTS):
    return [synthetic_clouds(batch, n_points, 1234 + 1000 * 1 + 17 * rank + i) for i in range(sets)]


def config_block(args, world):
    """`config` of the JSON line - the same dict on both arms (ours / --impl reference): it names the workload."""
    w = WORKLOADS[args.workload]
    return {"workload": w["text"].format(B=args.batch), "batch_per_gpu": args.batch, "precision": args.precision,
            "parallelism": f"dp{world} (clouds sharded across ranks" + (", no data-path collective)" if args.workload in ("c1", "c2")
                                                                        else ", one gradient all-reduce per step)"),
            "l2": f"{N_SETS} rotating input batches; the per-layer activations (3 x 50 MB fp32 at c1) exceed the 126 MB L2"}


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
    """The reference's own CPU implementation of the path = the oracle port (its CUDA-only wheels cannot run on the host
    cores), all host threads, same workload / config / metric as our arm.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload != "c1":
        emit({"impl": "reference", "unavailable": f"the CPU oracle port covers the forward path only; workload {args.workload} is a "
                                                  "training step (forward + backward + optimizer)"})
        return
    from oracle import model as omodel
    model, cfg = build_model()
    sd = model.state_dict()
    torch.set_num_threads(os.cpu_count() or 1)
    sets = make_clouds(args.ref_sample, 0, 4)
    for i in range(args.warmup):
        omodel.point_mamba_forward(sd, dict(cfg), sets[i % 4][: max(1, args.ref_sample // 8)])
    t0 = time.perf_counter()
    for i in range(args.steps):
        omodel.point_mamba_forward(sd, dict(cfg), sets[i % 4])
    dt = time.perf_counter() - t0
    val = args.ref_sample * args.steps / dt
    cores = torch.get_num_threads()
    sample = (f"{args.ref_sample} cloud(s) of the same workload per step, {args.steps} steps (warm-up steps run "
              f"{max(1, args.ref_sample // 8)} cloud(s) each); CPU oracle port of the reference path (oracle/model.py)")
    emit({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": config_block(args, args.gpus),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


# ----------------------------------------------------------------------------- shared harness
class Harness:
    def __init__(self, args):
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        assert torch.cuda.is_available(), "bench.py (impl=ours) needs a CUDA device: there is no CPU fallback"
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
        self.copy_stream = torch.cuda.Stream()

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(self, step_fn, steps, warmup):
        """W warm-up steps, then exactly `steps` steps between barrier + synchronize, CUDA events on the launching stream,
        max over ranks."""
        for i in range(warmup):
            step_fn(i)
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            step_fn(warmup + i)
        e1.record()
        self.barrier()
        ms = e0.elapsed_time(e1)
        if self.world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    def finish(self, hard=False):
        """Leave the process group.  ``hard``: the process holds CUDA graphs with captured NCCL kernels; tearing the
        communicator down under them can block for ever (seen with NCCL 2.28), so after a final barrier every rank flushes
        and exits without running destructors."""
        if self.world > 1:
            import torch.distributed as dist
            if hard:
                self.barrier()
                sys.stdout.flush()
                sys.stderr.flush()
                os._exit(0)
            dist.destroy_process_group()


class Feeder:
    """End-to-end input path: pinned host batch -> staging buffer on a copy stream (double-buffered) -> static graph input
    by a device copy on the compute stream.  The H2D transfer of step i+1 overlaps the compute of step i."""

    def __init__(self, h, host_sets, static_ins):
        self.h, self.host_sets, self.static_ins = h, host_sets, static_ins  # host_sets[i] = tuple of pinned tensors
        self.stage = [[torch.empty_like(t, device=h.dev) for t in static_ins] for _ in range(2)]
        self.ready = [torch.cuda.Event() for _ in range(2)]
        self.free = [torch.cuda.Event() for _ in range(2)]
        for e in self.free:
            e.record()
        self.bytes = sum(t.numel() * t.element_size() for t in static_ins)

    def feed(self, i):
        s = i % 2
        cur = torch.cuda.current_stream()
        cs = self.h.copy_stream
        cs.wait_event(self.free[s])
        with torch.cuda.stream(cs):
            for d, src in zip(self.stage[s], self.host_sets[i % len(self.host_sets)]):
                d.copy_(src, non_blocking=True)
            self.ready[s].record(cs)
        cur.wait_event(self.ready[s])
        for d, src in zip(self.static_ins, self.stage[s]):
            d.copy_(src)
        self.free[s].record(cur)


def scan_roofline(args, dev, mix, B, L, Dm, backward=False):
    """Dominant own kernel on the layer's real shapes, inputs rotated so they miss L2, CUDA events around a captured graph
    of back-to-back launches on the launching stream (an eager loop would time the host's tensor-map encodes).  The
    forward is timed the way the timed model calls it - fp32 inference hands it z as the gate silu(z), written by the
    in_proj GEMM's epilogue (si_mamba_b200/autograd.py, SIM_HOIST_ACT) - and, for reference, under the full mamba-ssm
    contract (softplus and silu evaluated in the kernel): `roofline.general_contract`."""
    from si_mamba_b200 import autograd as sim_autograd, ops
    z_gate = (not backward) and args.precision != "bf16" and sim_autograd._HOIST_ACT in ("z", "zdt")
    adt = torch.bfloat16 if args.precision == "bf16" else torch.float32
    es = 2 if adt == torch.bfloat16 else 4
    per_set = (8 if backward else 4) * B * L * Dm * es
    nsets = max(2, int(300e6 // per_set) + 1)
    g = torch.Generator(device=dev).manual_seed(7)
    sets = []
    for _ in range(nsets):
        xz = torch.randn(B, L, 2 * Dm, generator=g, device=dev).to(adt)
        u = torch.randn(B, L, Dm, generator=g, device=dev).to(adt)
        dl = (0.5 * torch.randn(B, L, Dm, generator=g, device=dev)).to(adt)
        xd = torch.randn(B, L, 56, generator=g, device=dev).to(adt)
        s = dict(u=u, dl=dl, B=xd[..., 24:40], C=xd[..., 40:], z=xz[..., Dm:], out=torch.empty(B, L, Dm, dtype=adt, device=dev))
        if backward:
            s["dout"] = torch.randn(B, L, Dm, generator=g, device=dev).to(adt)
            s["ckpt"] = torch.empty(ops.scan_checkpoint_shape(B, L, Dm), dtype=torch.float32, device=dev)
        sets.append(s)
    A = -torch.exp(mix.A_log.detach().float())
    Dp, bias = mix.D.detach(), mix.dt_proj.bias.detach()

    def fwd_call(s, gate=None):
        ops.selective_scan_tm(s["u"], s["dl"], A, s["B"], s["C"], Dp, s["z"], bias, True, out=s["out"], checkpoints=s.get("ckpt"),
                              z_gate=z_gate if gate is None else gate)

    def bwd_call(s):
        ops.selective_scan_bwd_tm(s["u"], s["dl"], A, s["B"], s["C"], Dp, s["z"], bias, s["dout"], s["ckpt"], True)

    call = bwd_call if backward else fwd_call
    iters, reps = 20, 3
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for s in sets:
            fwd_call(s, gate=False)
            if backward:
                bwd_call(s)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()

    def timed(fn):
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for i in range(iters):
                fn(sets[i % nsets])
        graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (iters * reps) * 1e-3

    sec = timed(call)
    sec_general = timed(lambda s: fwd_call(s, gate=False)) if z_gate else None
    E, S = B * L * Dm, B * L * 16
    if backward:  # reads u, delta, z, dout + writes du, ddelta, dz (7E) + B, C reads (2S) + fp32 dB, dC (2S*4) -- DESIGN.md 4.2
        alg = 7 * E * es + 2 * S * es + 2 * S * 4
        # the graph above also holds the zero-fills of dB / dC / dA / dD / dbias the op issues per launch
    else:
        alg = 4 * E * es + 2 * S * es
    peak, peak_src = peaks()
    key = f"{'bwd_' if backward else ''}{args.precision}_B{B}_L{L}{'_gate' if z_gate else ''}"
    traffic, tsrc = None, None
    tf = ROOT / "profiles" / "scan_traffic.json"
    if tf.exists():
        d = json.loads(tf.read_text())
        traffic = d.get(key)
        tsrc = d.get("_source")
    achieved = alg / sec / 1e9
    out = {"kernel": "selective_scan_bwd" if backward else "selective_scan_fwd", "bound": "hbm", "achieved": achieved, "peak": peak,
           "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
           "traffic_source": tsrc if traffic is not None else None, "peak_source": peak_src,
           "alg_bytes_per_launch": alg, "us_per_launch": sec * 1e6,
           "shape": {"B": B, "L": L, "D": Dm, "N": 16, "dtype": str(adt).split(".")[-1]},
           "timing": "CUDA events around a CUDA graph of 20 back-to-back launches on rotating inputs (> L2), 3 replays"}
    if not backward:
        # MUFU lane-ops per channel-step: 16 (one exp per state update of a general A) + 2 (softplus) + 2 (silu, unless the
        # gate arrives precomputed), at 16 /clk/SM and 1.965 GHz
        mufu = 18 if z_gate else 20
        t_mufu = mufu * E / (148 * 16 * 1.965e9)
        out["called_as"] = ("z = silu(z) precomputed by the in_proj GEMM epilogue (fp32 inference path of the timed model)"
                            if z_gate else "mamba-ssm contract: softplus(delta + bias) and silu(z) evaluated in the kernel")
        out["ceiling"] = {"bound": "mufu", "frac": (alg / t_mufu / 1e9) / peak,
                          "why": f"a general A needs one MUFU.EX2 per state update: {mufu} MUFU lane-ops per channel-step at "
                                 f"16 /clk/SM = {t_mufu * 1e6:.0f} us at this shape, before the 0.86-wave imbalance of 384 CTAs "
                                 "on 148 x 3 slots (DESIGN.md 4.1)"}
        if sec_general is not None:
            out["general_contract"] = {"us_per_launch": sec_general * 1e6, "achieved": alg / sec_general / 1e9,
                                       "frac": alg / sec_general / 1e9 / peak}
    return out


def clock_sampler(h):
    s = ClockSampler(h.local_rank)
    s.start()
    return s


# ----------------------------------------------------------------------------- our arm, C1 (forward)
def run_c1(args):
    from si_mamba_b200 import _lib

    h = Harness(args)
    dev, rank, world = h.dev, h.rank, h.world
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

    def compute():
        if graph is not None:
            graph.replay()
            return static_out
        return fwd(static_in)

    def step_resident(i):
        static_in.copy_(dev_sets[i % N_SETS])
        return compute()

    feeder = Feeder(h, [(t,) for t in host_sets], [static_in])

    def step_e2e(i):
        feeder.feed(i)
        out_host[i % 2].copy_(compute(), non_blocking=True)

    sampler = clock_sampler(h)
    ms_res = h.timed(step_resident, args.steps, max(args.warmup, 3))
    ms_e2e = h.timed(step_e2e, args.steps, max(args.warmup, 3))
    sampler.stop_flag = True

    L = 2 * cfg.k_top_eigenvectors * cfg.num_group
    roof = scan_roofline(args, dev, model.blocks.layers[0].mixer, B, L, 2 * cfg.trans_dim)

    if rank != 0:
        h.finish()
        return
    # reported baseline, rank 0 at N=1 only (the scaling runs would only repeat it)
    cpu_base = None
    if world == 1:
        cpu_val, cores = cpu_reference_rate(2)
        cpu_base = {"value": cpu_val, "unit": UNIT, "cores": cores, "kind": "port",
                    "sample": "2 clouds of the same workload, median of 3 runs (oracle/model.py)"}
    clouds = B * world * args.steps
    emit({
        "metric": METRIC, "value": clouds / (ms_res * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": config_block(args, world),
        "cuda_graph": graph is not None,
        "e2e": {"value": clouds / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": feeder.bytes,
                "d2h_bytes_per_step": B * cfg.cls_dim * 4, "ms_per_step": ms_e2e / args.steps,
                "how": "pinned host clouds -> device on a copy stream (double-buffered), graph replay, logits -> pinned host"},
        "gpu_launches": launches_per_step * args.steps,
        "gpu_launches_per_step": launches_per_step,
        "clocks": sampler.summary(),
        "roofline": roof,
        "cpu_baseline": cpu_base,
    })
    h.finish()


# ----------------------------------------------------------------------------- our arm, C2-C4 (training steps)
def run_train(args):
    import si_mamba_b200 as sm
    from si_mamba_b200 import _lib, train
    from si_mamba_b200.config import Config

    h = Harness(args)
    dev, rank, world = h.dev, h.rank, h.world
    name = args.workload
    w = WORKLOADS[name]
    B, N = args.batch, w["points"]
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = True
    torch.manual_seed(0)  # same initial weights on every rank (DDP broadcasts rank 0's; same seed = same effect)
    host_clouds = make_clouds(B, rank, n_points=N)
    gen = torch.Generator().manual_seed(99 + rank)
    autocast = torch.bfloat16 if w["dtype"] == "bf16" else None
    ocfg = Config(optimizer=Config(type="AdamW", kwargs=Config(lr=1e-3 if name == "c3" else 5e-4, weight_decay=0.05)),
                  scheduler=Config(type="CosLR", kwargs=Config(epochs=300, initial_epochs=10)))
    if name == "c2":
        cfg = sm.finetune_scan_hardest()
        model = sm.PointMamba(cfg).to(dev).train()
        host_sets = [(c.pin_memory(), torch.randint(0, cfg.cls_dim, (B,), generator=gen).pin_memory()) for c in host_clouds]
        pts = torch.empty(B, N, 3, device=dev)
        label = torch.zeros(B, dtype=torch.long, device=dev)
        statics = [pts, label]

        def loss_fn():
            ret = model(pts)
            loss, _acc = model.get_loss_acc(ret.float(), label)  # runner_finetune.py:203
            return loss.mean()
        L, G = 2 * cfg.k_top_eigenvectors * cfg.num_group, cfg.num_group
        mix = model.blocks.layers[0].mixer
        extra_host = None
    elif name == "c3":
        from si_mamba_b200.mae import rand_mask_host
        cfg = sm.pretrain()
        model = sm.Point_MAE_Mamba(cfg).to(dev).train()
        G, ratio = cfg.num_group, cfg.transformer_config.mask_ratio
        n_vis = G - int(ratio * G)
        # the per-cloud random mask is drawn on the host exactly as the reference does (numpy shuffle, :2232-2255) and
        # enters the captured step through a static tensor, like the clouds
        host_sets = [(c.pin_memory(), rand_mask_host(B, G, ratio).pin_memory()) for c in host_clouds]
        pts = torch.empty(B, N, 3, device=dev)
        mask = torch.zeros(B, G, dtype=torch.bool, device=dev)
        statics = [pts, mask]

        def loss_fn():
            return model(pts, bool_masked_pos=mask, n_vis=n_vis)
        L = 2 * cfg.transformer_config.k_top_eigenvectors * n_vis
        mix = model.MAE_encoder.blocks.layers[0].mixer
    else:
        cfg = sm.part_seg_config()
        model = sm.get_model(50, cfg).to(dev).train()
        G = 128
        host_sets = [(c.transpose(1, 2).contiguous().pin_memory(),
                      torch.nn.functional.one_hot(torch.randint(0, 16, (B,), generator=gen), 16).float().pin_memory(),
                      torch.randint(0, 50, (B, N), generator=gen).pin_memory(),
                      torch.rand(B, G, generator=gen).pin_memory())  # HLT tie-break noise, drawn on the CPU as the reference (:673)
                     for c in host_clouds]
        pts = torch.empty(B, 3, N, device=dev)
        cls = torch.zeros(B, 16, device=dev)
        target = torch.zeros(B, N, dtype=torch.long, device=dev)
        noise = torch.zeros(B, G, device=dev)
        statics = [pts, cls, target, noise]

        def loss_fn():
            out = model(pts, cls, hlt_noise=noise)
            return torch.nn.functional.nll_loss(out.reshape(-1, 50), target.reshape(-1))  # main.py:236-242
        L = 2 * G
        mix = model.blocks.layers[0].mixer

    for d, s in zip(statics, host_sets[0]):
        d.copy_(s)
    sync = train.GradSync(model, bucket_mb=1e9 if args.no_overlap else args.bucket_mb)
    optimizer, scheduler = train.build_opti_sche(model, ocfg, capturable=True, sync=sync)  # FlatAdamW: one kernel per step
    n0 = _lib.launches()
    step = train.TrainStep(model, optimizer, loss_fn, grad_clip=10.0, autocast_dtype=autocast, graph=not args.no_graph,
                           sync=sync)
    torch.cuda.synchronize()
    n1 = _lib.launches()
    step_launches = (n1 - n0) // (4 if not args.no_graph else 1)  # 3 eager warm-up steps + 1 capture
    dev_sets = [tuple(t.to(dev) for t in s) for s in host_sets]
    loss_host = [torch.empty(()).pin_memory() for _ in range(2)]

    def step_resident(i):
        for d, s in zip(statics, dev_sets[i % N_SETS]):
            d.copy_(s)
        return step()

    feeder = Feeder(h, host_sets, statics)

    def step_e2e(i):
        feeder.feed(i)
        loss_host[i % 2].copy_(step(), non_blocking=True)

    sampler = clock_sampler(h)
    ms_res = h.timed(step_resident, args.steps, max(args.warmup, 3))
    ms_e2e = h.timed(step_e2e, args.steps, max(args.warmup, 3))
    sampler.stop_flag = True
    final_loss = float(loss_host[(args.steps + max(args.warmup, 3) - 1) % 2])

    # exposed all-reduce time: the same step with the collectives removed (world forced to 1) on the same weights
    exposed = None
    if world > 1:
        sync.world = 1
        step1 = train.TrainStep(model, optimizer, loss_fn, grad_clip=10.0, autocast_dtype=autocast, graph=not args.no_graph,
                                sync=sync)
        def step_nocomm(i):
            for d, s in zip(statics, dev_sets[i % N_SETS]):
                d.copy_(s)
            return step1()
        ms_nocomm = h.timed(step_nocomm, args.steps, 3)
        sync.world = world
        exposed = {"ms_per_step_with_allreduce": ms_res / args.steps, "ms_per_step_without": ms_nocomm / args.steps,
                   "exposed_ms": (ms_res - ms_nocomm) / args.steps, "grad_bytes": sync.flat.numel() * 4,
                   "buckets": len(sync.buckets), "how": "same captured step with the all-reduces left out, timed the same way"}

    roof = scan_roofline(args, dev, mix, B, L, 768, backward=True)
    if rank == 0:
        clouds = B * world * args.steps
        emit({
            "metric": w["metric"], "value": clouds / (ms_res * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": config_block(args, world),
            "cuda_graph": not args.no_graph,
            "e2e": {"value": clouds / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": feeder.bytes, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps,
                    "how": "pinned host inputs -> device on a copy stream (double-buffered), captured step replay, loss -> pinned host"},
            "gpu_launches": step_launches * args.steps, "gpu_launches_per_step": step_launches,
            "allreduce": exposed, "loss": final_loss,
            "clocks": sampler.summary(), "roofline": roof, "cpu_baseline": None,
        })
    h.finish(hard=not args.no_graph)


def main():
    args = parse()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "c1":
        run_c1(args)
    else:
        run_train(args)


if __name__ == "__main__":
    main()
