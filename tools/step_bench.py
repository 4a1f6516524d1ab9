#!/usr/bin/env python
"""Training-step timings of BASELINE.json configs 2-4 on the kernels of this repo (synthetic clouds, random init):

  C2  ScanObjectNN-hardest finetune shape: 2048 pts, 128 patches x 32, L=1024, bf16 autocast, forward + backward
      + grad-clip (tools/runner_finetune.py), batch 32 per GPU
  C3  MAE pre-training step, pretrain.yaml shape: 1024 pts, 64 patches, mask 0.6, 12 + 4 layers, bf16 autocast,
      forward + backward + AdamW; batch 16 per GPU (128 over 8 GPUs), DDP over NCCL when launched by torchrun
  C4  part segmentation: 2048 pts, 128 patches, HLT (L=256) or SAST (L=1024), forward + backward, batch 16 per GPU

    python tools/step_bench.py [--configs C2,C3,C4] [--steps 10] [--warmup 3]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/step_bench.py ...

One JSON line per config on rank 0: device time per step (CUDA events, max over ranks) and whole-job clouds/s.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import si_mamba_b200 as sm  # noqa: E402
from oracle import tokenizer  # noqa: E402  (synthetic-input generator only)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="C2,C3,C4")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--graph", action="store_true", help="capture forward + backward + optimizer of C3 in one CUDA graph "
                    "(single GPU; the host-side random mask becomes a graph input)")
    args = ap.parse_args()
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ddp = world > 1
    if ddp:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)

    def wrap(m):
        if not ddp:
            return m
        from torch.nn.parallel import DistributedDataParallel as DDP
        return DDP(m, device_ids=[local], find_unused_parameters=True)  # fork-only heads get no gradient

    def timed(step):
        for _ in range(max(args.warmup, 3)):
            step()
        torch.cuda.synchronize()
        if ddp:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        if ddp:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    def report(name, ms, batch, **extra):
        if rank == 0:
            print(json.dumps(dict(config=name, n_gpus=world, ms_per_step=round(ms, 3), batch_per_gpu=batch,
                                  clouds_per_s=round(batch * world / ms * 1e3, 1), **extra)), flush=True)

    for name in args.configs.split(","):
        torch.manual_seed(0)
        if name == "C2":
            cfg = sm.finetune_scan_hardest()
            B, N = 32, 2048
            model = wrap(sm.PointMamba(cfg).to(dev).train())
            pts = tokenizer.synthetic_clouds(B, N, 2000 + rank, "surface").to(dev)
            label = torch.randint(0, cfg.cls_dim, (B,), device=dev)

            def step():
                model.zero_grad(set_to_none=True)
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    logits = model(pts)
                loss = torch.nn.functional.cross_entropy(logits.float(), label)
                loss.backward()
                torch.nn.utils.clip_grad_norm_(model.parameters(), 10.0)

            tag = ""
            if args.graph and not ddp:
                from si_mamba_b200.train import GraphedStep

                def train_step():
                    model.zero_grad(set_to_none=False)
                    with torch.autocast("cuda", dtype=torch.bfloat16):
                        logits = model(pts)
                    loss = torch.nn.functional.cross_entropy(logits.float(), label)
                    loss.backward()
                    torch.nn.utils.clip_grad_norm_(model.parameters(), 10.0)
                    return loss

                gs = GraphedStep(train_step)
                step = gs.replay  # noqa: F811
                tag = ", CUDA graph"
            report("C2 scan-hardest finetune fwd+bwd bf16 (2048 pts, L=1024" + tag + ")", timed(step), B, dtype="bf16 autocast")
        elif name == "C3":
            cfg = sm.pretrain()
            B, N = 16, 1024
            model = wrap(sm.Point_MAE_Mamba(cfg).to(dev).train())
            opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=0.05)
            pts = tokenizer.synthetic_clouds(B, N, 3000 + rank, "surface").to(dev)

            def step():
                opt.zero_grad(set_to_none=True)
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    loss = model(pts)
                loss.backward()
                opt.step()

            tag = ""
            if args.graph and not ddp:
                # the step at batch 16 is launch-bound in eager mode (11 ms of GPU work in a 23 ms step): capture it.
                # The per-cloud random mask is drawn on the host exactly as the reference does (numpy shuffle) and
                # copied into a static tensor before every replay.
                from si_mamba_b200.mae import rand_mask_host
                from si_mamba_b200.train import GraphedStep
                opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=0.05, capturable=True)
                G, ratio = cfg.num_group, cfg.transformer_config.mask_ratio
                static_mask = rand_mask_host(B, G, ratio).to(dev)

                def train_step():
                    opt.zero_grad(set_to_none=False)
                    with torch.autocast("cuda", dtype=torch.bfloat16):
                        loss = model(pts, bool_masked_pos=static_mask, n_vis=G - int(ratio * G))
                    loss.backward()
                    opt.step()
                    return loss

                gs = GraphedStep(train_step)

                def step():  # noqa: F811
                    static_mask.copy_(rand_mask_host(B, G, ratio), non_blocking=True)
                    gs.replay()

                tag = ", CUDA graph"
            report("C3 MAE pretrain step bf16 (1024 pts, mask 0.6, 12+4 layers, AdamW" + (", DDP" if ddp else "") + tag + ")",
                   timed(step), B, dtype="bf16 autocast")
        elif name == "C4":
            for method in ("HLT", "SAST"):
                cfg = sm.part_seg_config()
                cfg.update(method=method)
                B, N = 16, 2048
                model = wrap(sm.get_model(50, cfg).to(dev).train())
                pts = tokenizer.synthetic_clouds(B, N, 4000 + rank, "surface").to(dev).transpose(1, 2).contiguous()
                cls = torch.nn.functional.one_hot(torch.randint(0, 16, (B,), device=dev), 16).float()
                target = torch.randint(0, 50, (B, N), device=dev)

                def step():
                    model.zero_grad(set_to_none=True)
                    out = model(pts, cls)
                    loss = torch.nn.functional.nll_loss(out.reshape(-1, 50), target.reshape(-1))
                    loss.backward()

                tag = ""
                if args.graph and not ddp:
                    from si_mamba_b200.train import GraphedStep
                    noise = torch.rand(B, 128, device=dev)  # HLT tie-break noise: a graph input

                    def train_step(model=model, pts=pts, cls=cls, target=target, noise=noise):
                        model.zero_grad(set_to_none=False)
                        out = model(pts, cls, hlt_noise=noise)
                        loss = torch.nn.functional.nll_loss(out.reshape(-1, 50), target.reshape(-1))
                        loss.backward()
                        return loss

                    gs = GraphedStep(train_step)

                    def step(gs=gs, noise=noise):  # noqa: F811
                        noise.copy_(torch.rand(noise.shape), non_blocking=True)
                        gs.replay()

                    tag = ", CUDA graph"
                report(f"C4 part segmentation fwd+bwd fp32 (2048 pts, 128 patches, {method}{tag})", timed(step), B,
                       dtype="fp32")
        else:
            raise SystemExit(f"unknown config {name}")
    if ddp:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
