"""Trainer plumbing on B200: what the reference's runners put around the path in a training step,
kept to the pieces the step needs and in the reference's own formats.

  * ``add_weight_decay`` / ``build_opti_sche``     tools/builder.py:57-95 (AdamW with the no-decay rule, CosLR)
  * ``CosineLRScheduler``                            timm's scheduler as builder.py:77-86 configures it (un-vendored
                                                     dependency: restated from its published formula)
  * ``save_checkpoint`` / ``resume_model`` / ``resume_optimizer`` / ``load_model``
                                                     tools/builder.py:112-190: the ``{base_model, optimizer, epoch,
                                                     metrics, best_metrics}`` dict with ``module.`` stripping
  * ``Acc_Metric``                                   tools/runner_pretrain.py:24-44
  * ``SyntheticClouds``                              a seeded stand-in for datasets/*.py (no dataset on the GPU box)
  * ``GradSync``                                     the DDP gradient all-reduce of runner_pretrain.py:109-120 as
                                                     bucketed NCCL all-reduces of one flat gradient buffer, launched
                                                     from autograd hooks on a communication stream so they overlap the
                                                     rest of the backward - and, unlike DistributedDataParallel with
                                                     ``find_unused_parameters=True`` (a host-side graph walk per step),
                                                     capturable in a CUDA graph
  * ``GraphedStep`` / ``TrainStep``                  forward + backward + all-reduce + clip + optimizer captured once
                                                     and replayed (the step at 16 clouds per GPU is launch-bound eager)
"""

from __future__ import annotations

import math
import os
from typing import Callable, Optional

import torch
import torch.distributed as dist

from .autograd import invalidate_param_cache


# ----------------------------------------------------------------------------- optimizer / scheduler
def add_weight_decay(model: torch.nn.Module, weight_decay: float = 1e-5, skip_list=()):
    """tools/builder.py:60-73: biases, 1-D parameters (norm scales, A_log is 2-D and decays, D does not) and anything with
    'token' in its name go to the weight_decay = 0 group.  ``model`` may be wrapped (``.module``) or bare."""
    base = model.module if hasattr(model, "module") else model
    decay, no_decay = [], []
    for name, param in base.named_parameters():
        if not param.requires_grad:
            continue
        if len(param.shape) == 1 or name.endswith(".bias") or "token" in name or name in skip_list:
            no_decay.append(param)
        else:
            decay.append(param)
    return [{"params": no_decay, "weight_decay": 0.0}, {"params": decay, "weight_decay": weight_decay}]


class CosineLRScheduler:
    """timm ``CosineLRScheduler(optimizer, t_initial, lr_min, warmup_t, warmup_lr_init, cycle_mul=1, cycle_decay,
    cycle_limit, t_in_epochs=True)`` as tools/builder.py:77-86 and part_segmentation/main.py:205-213 build it: linear
    warm-up from ``warmup_lr_init`` over ``warmup_t`` epochs, then ``lr_min + (lr - lr_min)/2 * (1 + cos(pi t / T))`` with
    the cycle index decaying both ends by ``cycle_decay``; past ``cycle_limit`` cycles the rate stays at ``lr_min``.
    ``step(epoch)`` is called once per epoch AFTER the epoch (runner_pretrain.py:282-286).  Learning rates held as device
    tensors (capturable optimizers inside a CUDA graph) are updated in place."""

    def __init__(self, optimizer, t_initial: int, lr_min: float = 0.0, cycle_mul: float = 1.0, cycle_decay: float = 1.0,
                 cycle_limit: int = 1, warmup_t: int = 0, warmup_lr_init: float = 0.0, warmup_prefix: bool = False,
                 t_in_epochs: bool = True):
        assert t_initial > 0 and cycle_mul == 1.0, "the reference only uses cycle_mul = 1"
        self.optimizer = optimizer
        self.t_initial, self.lr_min, self.cycle_decay, self.cycle_limit = t_initial, lr_min, cycle_decay, cycle_limit
        self.warmup_t, self.warmup_lr_init, self.warmup_prefix, self.t_in_epochs = warmup_t, warmup_lr_init, warmup_prefix, t_in_epochs
        for g in optimizer.param_groups:
            g.setdefault("initial_lr", float(g["lr"]))
        self.base_values = [float(g["initial_lr"]) for g in optimizer.param_groups]
        if warmup_t:
            self.warmup_steps = [(v - warmup_lr_init) / warmup_t for v in self.base_values]
            self._set([warmup_lr_init for _ in self.base_values])
        else:
            self.warmup_steps = [1.0 for _ in self.base_values]

    def _get_lr(self, t: int):
        if t < self.warmup_t:
            return [self.warmup_lr_init + t * s for s in self.warmup_steps]
        if self.warmup_prefix:
            t = t - self.warmup_t
        i = t // self.t_initial
        t_curr = t - self.t_initial * i
        gamma = self.cycle_decay ** i
        if i < self.cycle_limit:
            return [self.lr_min * gamma + 0.5 * (v * gamma - self.lr_min * gamma) * (1 + math.cos(math.pi * t_curr / self.t_initial))
                    for v in self.base_values]
        return [self.lr_min for _ in self.base_values]

    def _set(self, values):
        for g, v in zip(self.optimizer.param_groups, values):
            if isinstance(g["lr"], torch.Tensor):
                g["lr"].fill_(v)
            else:
                g["lr"] = v

    def step(self, epoch: int, metric=None):
        if self.t_in_epochs:
            self._set(self._get_lr(epoch))

    def step_update(self, num_updates: int, metric=None):
        if not self.t_in_epochs:
            self._set(self._get_lr(num_updates))

    def state_dict(self):
        return {k: v for k, v in self.__dict__.items() if k != "optimizer"}

    def load_state_dict(self, sd):
        self.__dict__.update(sd)


def build_opti_sche(base_model, config, capturable: bool = False, sync: Optional["GradSync"] = None):
    """tools/builder.py:57-106 for the optimizer / scheduler types the shipped configs select (AdamW + CosLR; Adam / SGD /
    StepLR kept, LambdaLR and the BN-momentum scheduler are not used by any hot-path config).  ``capturable`` puts the
    learning rate and step counters on the device so ``optimizer.step()`` can live inside a CUDA graph.  With a GradSync
    (``sync``) on a CUDA model, AdamW is the single-kernel FlatAdamW over its flat buffers."""
    oc = config.optimizer
    kw = dict(oc.kwargs)
    dev = next(base_model.parameters()).device
    if oc.type == "AdamW" and sync is not None and dev.type == "cuda":
        optimizer = FlatAdamW(add_weight_decay(base_model, weight_decay=oc.kwargs.weight_decay), sync, **kw)
        return optimizer, _build_scheduler(optimizer, config)
    if capturable:
        kw["lr"] = torch.tensor(float(kw["lr"]), device=dev)
        kw["capturable"] = True
    if oc.type == "AdamW":
        optimizer = torch.optim.AdamW(add_weight_decay(base_model, weight_decay=oc.kwargs.weight_decay), **kw)
    elif oc.type == "Adam":
        optimizer = torch.optim.Adam(base_model.parameters(), **kw)
    elif oc.type == "SGD":
        kw.pop("capturable", None)
        optimizer = torch.optim.SGD(base_model.parameters(), nesterov=True, **kw)
    else:
        raise NotImplementedError(oc.type)
    if capturable:  # every group gets its own lr tensor (the scheduler writes them in place)
        for g in optimizer.param_groups:
            g["initial_lr"] = float(oc.kwargs.lr)
            g["lr"] = torch.tensor(float(oc.kwargs.lr), device=dev)
    return optimizer, _build_scheduler(optimizer, config)


def _build_scheduler(optimizer, config):
    sc = config.scheduler
    if sc.type == "CosLR":
        scheduler = CosineLRScheduler(optimizer, t_initial=sc.kwargs.epochs, cycle_mul=1, lr_min=1e-6, cycle_decay=0.1,
                                      warmup_lr_init=1e-6, warmup_t=sc.kwargs.initial_epochs, cycle_limit=1, t_in_epochs=True)
    elif sc.type == "StepLR":
        scheduler = torch.optim.lr_scheduler.StepLR(optimizer, **dict(sc.kwargs))
    elif sc.type == "function":
        scheduler = None
    else:
        raise NotImplementedError(sc.type)
    return scheduler


# ----------------------------------------------------------------------------- checkpoints (tools/builder.py:112-190)
class Acc_Metric:
    """tools/runner_pretrain.py:24-44."""

    def __init__(self, acc=0.):
        self.acc = acc["acc"] if isinstance(acc, dict) else acc

    def better_than(self, other):
        return self.acc > other.acc

    def state_dict(self):
        return {"acc": self.acc}


def _bare(model):
    return model.module if hasattr(model, "module") else model


def save_checkpoint(base_model, optimizer, epoch, metrics, best_metrics, prefix, experiment_path, rank: int = 0):
    """tools/builder.py:151-161: rank 0 writes ``<experiment_path>/<prefix>.pth``.  Keys carry no ``module.`` prefix."""
    if rank != 0:
        return None
    path = os.path.join(experiment_path, prefix + ".pth")
    torch.save({
        "base_model": _bare(base_model).state_dict(),
        "optimizer": optimizer.state_dict(),
        "epoch": epoch,
        "metrics": metrics.state_dict() if metrics is not None else dict(),
        "best_metrics": best_metrics.state_dict() if best_metrics is not None else dict(),
    }, path)
    return path


def resume_model(base_model, experiment_path, map_location="cpu"):
    """tools/builder.py:112-134 -> (start_epoch, best_metrics dict); (0, 0) when there is no ``ckpt-last.pth``."""
    path = os.path.join(experiment_path, "ckpt-last.pth")
    if not os.path.exists(path):
        return 0, 0
    sd = torch.load(path, map_location=map_location, weights_only=False)
    _bare(base_model).load_state_dict({k.replace("module.", ""): v for k, v in sd["base_model"].items()}, strict=True)
    invalidate_param_cache()
    best = sd["best_metrics"]
    if not isinstance(best, dict):
        best = best.state_dict()
    return sd["epoch"] + 1, best


def resume_optimizer(optimizer, experiment_path):
    """tools/builder.py:137-148."""
    path = os.path.join(experiment_path, "ckpt-last.pth")
    if not os.path.exists(path):
        return 0, 0, 0
    optimizer.load_state_dict(torch.load(path, map_location="cpu", weights_only=False)["optimizer"])
    return None


def load_model(base_model, ckpt_path):
    """tools/builder.py:164-190: weights under ``model`` or ``base_model``, ``module.`` stripped, strict."""
    if not os.path.exists(ckpt_path):
        raise NotImplementedError("no checkpoint file from path %s..." % ckpt_path)
    sd = torch.load(ckpt_path, map_location="cpu", weights_only=False)
    if sd.get("model") is not None:
        base = {k.replace("module.", ""): v for k, v in sd["model"].items()}
    elif sd.get("base_model") is not None:
        base = {k.replace("module.", ""): v for k, v in sd["base_model"].items()}
    else:
        raise RuntimeError("mismatch of ckpt weight")
    _bare(base_model).load_state_dict(base, strict=True)
    invalidate_param_cache()
    return sd.get("epoch", -1), sd.get("metrics", "No Metrics")


# ----------------------------------------------------------------------------- synthetic dataset
class SyntheticClouds(torch.utils.data.Dataset):
    """Seeded stand-in for the reference's datasets (datasets/ShapeNet55Dataset.py, ModelNetDataset.py,
    part_segmentation/dataset.py): item i is a deterministic function of (seed, i).  ``task``:
    'pretrain' -> points (N,3);  'cls' -> (points, label);  'seg' -> (points, object class, per-point part label).
    Clouds are anisotropic blob mixtures normalised like the datasets (centroid 0, max-norm 1)."""

    def __init__(self, n_items: int, n_points: int, task: str = "pretrain", n_classes: int = 40, n_parts: int = 50,
                 seed: int = 0):
        assert task in ("pretrain", "cls", "seg")
        self.n_items, self.n_points, self.task, self.n_classes, self.n_parts, self.seed = n_items, n_points, task, n_classes, n_parts, seed

    def __len__(self):
        return self.n_items

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed * 1000003 + i)
        N = self.n_points
        n_blob = int(torch.randint(3, 7, (1,), generator=g))
        which = torch.randint(0, n_blob, (N,), generator=g)
        ctr = torch.randn(n_blob, 3, generator=g) * 0.5
        axes = torch.rand(n_blob, 3, generator=g) * 0.6 + 0.1
        v = torch.randn(N, 3, generator=g)
        v = v / v.norm(dim=-1, keepdim=True)
        pts = ctr[which] + v * axes[which] + 0.01 * torch.randn(N, 3, generator=g)
        pts = pts - pts.mean(dim=0, keepdim=True)
        pts = (pts / pts.norm(dim=-1).max()).float()
        if self.task == "pretrain":
            return pts
        label = int(torch.randint(0, self.n_classes, (1,), generator=g))
        if self.task == "cls":
            return pts, label
        return pts, label % 16, (which % self.n_parts).long()


def shard_sampler(dataset, rank: int, world: int, shuffle: bool = True, seed: int = 0):
    """tools/builder.py:23-24: DistributedSampler(dataset, shuffle=...) - every rank sees len/world items per epoch."""
    return torch.utils.data.distributed.DistributedSampler(dataset, num_replicas=world, rank=rank, shuffle=shuffle, seed=seed)


# ----------------------------------------------------------------------------- gradient all-reduce
class GradSync:
    """Data-parallel gradient averaging for one process per GPU (runner_pretrain.py:109-120 wraps the model in
    DistributedDataParallel(find_unused_parameters=True); main.py:72-79 gives each rank total_bs // world_size clouds).

    Gradients live in ONE flat buffer cut into buckets in reverse registration order (the order gradients become ready).
    Every step starts with ``grad = None`` on all parameters, so autograd hands each parameter its gradient tensor as is
    (a pre-set ``.grad`` would cost one ``grad += new`` kernel per parameter: ~490 launches and 1.7 ms of the 25 ms C2
    step).  A post-accumulate hook counts arrivals; when a bucket is complete its gradients are packed into the flat buffer
    by one multi-tensor copy and the bucket's all-reduce (average) is enqueued on a communication stream that waits on the
    backward stream at that point only, so the transfer overlaps the remaining backward kernels; ``finish()`` flushes
    buckets that are still incomplete, joins the streams and re-points every ``.grad`` at its (now averaged) slice of the
    flat buffer for clipping and the optimizer.  No per-step host-side graph traversal and no host sync, so the whole step
    - collectives included - can be captured in a CUDA graph.  Parameters that never receive a gradient (fork-only heads,
    the MAE model's ``decoder_pos_embed``) are detected in the first step by ``prune_unused()`` and keep ``grad = None``,
    so the optimizer skips them exactly as it does under DDP."""

    def __init__(self, model: torch.nn.Module, process_group=None, bucket_mb: float = 16.0, compress: Optional[str] = None):
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.params = [p for p in model.parameters() if p.requires_grad]
        assert self.params, "no trainable parameters"
        p0 = self.params[0]
        assert all(p.dtype == p0.dtype and p.device == p0.device for p in self.params), "one dtype / device"
        self.compress = compress
        assert compress in (None, "bf16")
        order = list(reversed(self.params))
        offs, off = {}, 0
        self.buckets = []  # (start, end, [params])
        cur, cur_start = [], 0
        limit = int(bucket_mb * 2 ** 20 // p0.element_size())
        for p in order:
            offs[p] = off
            off += (p.numel() + 3) // 4 * 4  # 16-byte aligned views
            cur.append(p)
            if off - cur_start >= limit:
                self.buckets.append((cur_start, off, cur))
                cur, cur_start = [], off
        if cur:
            self.buckets.append((cur_start, off, cur))
        self.flat = torch.zeros(off, dtype=p0.dtype, device=p0.device)
        self.bucket_of = {}
        for bi, (_, _, ps) in enumerate(self.buckets):
            for p in ps:
                self.bucket_of[p] = bi
        self._offs = offs
        self._views = {p: self.flat[offs[p]:offs[p] + p.numel()].view_as(p) for p in self.params}
        self.comm = torch.cuda.Stream(device=p0.device) if p0.is_cuda else None
        self._arrived = [0] * len(self.buckets)
        self._launched = [False] * len(self.buckets)
        self._expected = [len(ps) for _, _, ps in self.buckets]
        self._fired = set()
        self._handles = [p.register_post_accumulate_grad_hook(self._hook) for p in self.params]
        self.enabled = True
        self.allreduce_launches = 0

    # -- per step
    def begin_step(self):
        """Drop last step's gradients (host only - no kernel): autograd then stores each new gradient as is."""
        for p in self.params:
            p.grad = None

    zero_grad = begin_step

    def prune_unused(self):
        """After the first backward: parameters whose hook never fired get no gradient in this model."""
        n = 0
        for bi, (_, _, ps) in enumerate(self.buckets):
            for p in ps:
                if p not in self._fired:
                    p._gradsync_unused = True
                    n += 1
            self._expected[bi] = sum(1 for p in ps if p in self._fired)
        return n

    def _reduce(self, bi: int):
        s, e, ps = self.buckets[bi]
        self._launched[bi] = True
        have = [p for p in ps if p.grad is not None]
        if have:  # pack the bucket: one multi-tensor copy
            torch._foreach_copy_([self._views[p] for p in have], [p.grad for p in have])
        for p in ps:  # a parameter that usually has a gradient but got none this step contributes zeros
            if p.grad is None and p in self._fired:
                self._views[p].zero_()
        if self.world == 1:
            return
        view = self.flat[s:e]
        self.allreduce_launches += 1
        if self.comm is None:  # CPU tensors (gloo tests): synchronous
            dist.all_reduce(view, group=self.group)
            view.div_(self.world)
            return
        cur = torch.cuda.current_stream()
        self.comm.wait_stream(cur)
        with torch.cuda.stream(self.comm):
            if self.compress == "bf16":
                low = view.to(torch.bfloat16)
                low.div_(self.world)
                dist.all_reduce(low, group=self.group)
                view.copy_(low)
            elif dist.get_backend(self.group) == "nccl":
                dist.all_reduce(view, op=dist.ReduceOp.AVG, group=self.group)
            else:
                dist.all_reduce(view, group=self.group)
                view.div_(self.world)

    def _hook(self, p):
        if not self.enabled:
            return
        self._fired.add(p)
        bi = self.bucket_of[p]
        self._arrived[bi] += 1
        if self._arrived[bi] == self._expected[bi] and not self._launched[bi]:
            self._reduce(bi)

    def finish(self):
        """Call after ``backward()``: reduce what is left, make the compute stream wait for the communication stream and
        point every ``.grad`` at its averaged slice of the flat buffer."""
        for bi in range(len(self.buckets)):
            if not self._launched[bi]:
                self._reduce(bi)
        if self.comm is not None and self.world > 1:
            torch.cuda.current_stream().wait_stream(self.comm)
        for p in self.params:
            if p.grad is not None:
                p.grad = self._views[p]
        self._arrived = [0] * len(self.buckets)
        self._launched = [False] * len(self.buckets)

    def flatten_parameters(self):
        """Move every trainable parameter into ONE flat buffer laid out like the gradient buffer (same offsets), so an
        optimizer can update all of them with a single kernel (FlatAdamW).  Values, shapes, names and state-dict keys are
        unchanged; only ``p.data`` now points into ``flat_param``.  Idempotent."""
        if getattr(self, "flat_param", None) is not None:
            return self.flat_param
        self.flat_param = torch.zeros_like(self.flat)
        with torch.no_grad():
            for p in self.params:
                o = self._offs[p]
                view = self.flat_param[o:o + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
        invalidate_param_cache()
        return self.flat_param

    def clip_coef(self, max_norm: float):
        """(total_norm, coefficient) of torch.nn.utils.clip_grad_norm_(parameters, max_norm, 2) as device scalars, for an
        optimizer that applies the coefficient itself (FlatAdamW.grad_scale) instead of a pass over the gradients."""
        total = self.grad_norm()
        return total, torch.clamp(max_norm / (total + 1e-6), max=1.0).reshape(1)

    def grad_norm(self) -> torch.Tensor:
        """2-norm over every gradient = norm of the flat buffer (padding and unused slots are zero)."""
        return torch.linalg.vector_norm(self.flat, 2)

    def clip_grad_norm_(self, max_norm: float) -> torch.Tensor:
        """torch.nn.utils.clip_grad_norm_(parameters, max_norm, 2) on the flat buffer: two kernels, no host sync."""
        total = self.grad_norm()
        self.flat.mul_(torch.clamp(max_norm / (total + 1e-6), max=1.0))
        return total


class FlatAdamW(torch.optim.Optimizer):
    """torch.optim.AdamW (the optimizer tools/builder.py:74 and part_segmentation/main.py:201 build) as ONE kernel over the
    flat parameter / gradient buffers of a GradSync (sim_adamw_flat): same update rule, same hyper-parameters per group,
    parameters without a gradient are left untouched like torch does, learning rate / step count / clip coefficient are
    device scalars so the step can be captured in a CUDA graph.  (torch's capturable multi-tensor AdamW spends ~700 scalar
    kernels per step on bias corrections: 1.5 - 2.5 ms of a 9 - 22 ms training step.)  ``state_dict()`` /
    ``load_state_dict()`` use torch.optim.AdamW's layout - per-parameter ``step`` / ``exp_avg`` / ``exp_avg_sq`` - so
    checkpoints written by builder.save_checkpoint move between this optimizer and the reference's."""

    def __init__(self, params, sync: GradSync, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False, foreach=None,
                        capturable=True, differentiable=False, fused=None)
        super().__init__(params, defaults)
        self.sync = sync
        flat_p = sync.flatten_parameters()
        assert flat_p.is_cuda and flat_p.dtype == torch.float32, "FlatAdamW runs on fp32 CUDA parameters"
        n, dev = flat_p.numel(), flat_p.device
        self.m, self.v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
        self.wd = torch.full((n,), -1.0, device=dev)
        self.step_t = torch.zeros(1, device=dev)
        self.lr_t = torch.tensor([float(lr)], device=dev)
        self.grad_scale = None  # optional device scalar (GradSync.clip_coef): gradients are scaled inside the kernel
        g0 = self.param_groups[0]
        assert all(g["betas"] == g0["betas"] and g["eps"] == g0["eps"] and float(g["lr"]) == float(g0["lr"])
                   for g in self.param_groups), \
            "one learning rate / betas / eps for all groups (only weight_decay differs in the reference)"
        for g in self.param_groups:
            g["initial_lr"] = float(g["lr"])
            g["lr"] = self.lr_t[0]  # a 0-dim view: CosineLRScheduler fills it in place
            for p in g["params"]:
                assert p in sync._offs, "every optimised parameter must belong to the GradSync"
                o = sync._offs[p]
                self.state[p] = {"step": self.step_t[0], "exp_avg": self.m[o:o + p.numel()].view_as(p),
                                 "exp_avg_sq": self.v[o:o + p.numel()].view_as(p)}
        self._active = None

    def _rebuild_wd(self, active):
        self.wd.fill_(-1.0)
        for g in self.param_groups:
            for p in g["params"]:
                if active[p]:
                    o = self.sync._offs[p]
                    self.wd[o:o + p.numel()].fill_(float(g["weight_decay"]))
        self._active = dict(active)

    @torch.no_grad()
    def step(self, closure=None):
        from . import _lib
        assert closure is None
        active = {p: p.grad is not None for g in self.param_groups for p in g["params"]}
        if active != self._active:  # first step, or the set of parameters with a gradient changed
            self._rebuild_wd(active)
        g0 = self.param_groups[0]
        flat_p = self.sync.flat_param
        gs = None if self.grad_scale is None else self.grad_scale.float().reshape(1).contiguous()
        _lib.call("sim_adamw_flat", flat_p.data_ptr(), self.sync.flat.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                  self.wd.data_ptr(), flat_p.numel(), self.lr_t.data_ptr(), self.step_t.data_ptr(),
                  None if gs is None else gs.data_ptr(), float(g0["betas"][0]), float(g0["betas"][1]), float(g0["eps"]),
                  torch.cuda.current_stream().cuda_stream)
        return None

    def state_dict(self):
        sd = super().state_dict()
        mine = [p for g in self.param_groups for p in g["params"]]
        for i, p in enumerate(mine):  # torch keeps no state for a parameter that never had a gradient
            if self._active is not None and not self._active.get(p, False):
                sd["state"].pop(i, None)
        for st in sd["state"].values():  # independent tensors, torch.optim.AdamW's layout
            st["step"] = st["step"].detach().clone().reshape(())
            st["exp_avg"] = st["exp_avg"].detach().clone()
            st["exp_avg_sq"] = st["exp_avg_sq"].detach().clone()
        for g in sd["param_groups"]:
            g["lr"] = float(self.lr_t)
        return sd

    def load_state_dict(self, state_dict):
        groups = state_dict["param_groups"]
        assert [len(g["params"]) for g in groups] == [len(g["params"]) for g in self.param_groups], "parameter groups differ"
        ids = [i for g in groups for i in g["params"]]
        mine = [p for g in self.param_groups for p in g["params"]]
        step = None
        with torch.no_grad():
            for i, p in zip(ids, mine):
                st = state_dict["state"].get(i)
                if st is None:
                    continue
                self.state[p]["exp_avg"].copy_(st["exp_avg"])
                self.state[p]["exp_avg_sq"].copy_(st["exp_avg_sq"])
                step = float(st["step"])
            if step is not None:
                self.step_t.fill_(step)
            self.lr_t.fill_(float(groups[0]["lr"]))
        for g, src in zip(self.param_groups, groups):
            g["weight_decay"] = src["weight_decay"]
            g["initial_lr"] = src.get("initial_lr", g.get("initial_lr"))
        self._active = None


# ----------------------------------------------------------------------------- whole-step CUDA graph
class GraphedStep:
    """Capture ``step_fn()`` (forward + backward [+ optimizer.step()]) into one CUDA graph after ``warmup`` eager runs on
    a side stream; ``replay()`` runs it.  ``step_fn`` must read its inputs from tensors that stay alive and are updated
    in place between replays (``.copy_``), must not sync with the host, and must zero gradients with
    ``set_to_none=True`` BEFORE capture only (gradients are then static graph outputs).  Optimizers need
    ``capturable=True``.  ``after_first`` runs once after the first eager run (GradSync.prune_unused)."""

    def __init__(self, step_fn: Callable[[], torch.Tensor], warmup: int = 3, after_first: Optional[Callable[[], None]] = None):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for i in range(warmup):
                step_fn()
                if i == 0 and after_first is not None:
                    after_first()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.output = step_fn()

    def replay(self):
        self.graph.replay()
        # the replay updated parameters / BatchNorm statistics in place without bumping their version counters
        invalidate_param_cache()
        return self.output


class TrainStep:
    """One optimisation step of the reference's loops on static inputs:

        zero_grad -> autocast(bf16)? forward -> loss.backward() -> [gradient all-reduce, overlapped] ->
        clip_grad_norm_(10) -> optimizer.step()          (runner_pretrain.py:243-260, runner_finetune.py, main.py:239-251)

    ``loss_fn()`` computes the loss from tensors the caller keeps alive and refreshes in place (``inputs``); with
    ``graph=True`` the whole step - NCCL all-reduces included - is captured once and replayed.  ``double_step`` mirrors the
    part-segmentation loop, which calls ``optimizer.step()`` once before and once after the clip (main.py:244-251)."""

    def __init__(self, model, optimizer, loss_fn: Callable[[], torch.Tensor], grad_clip: Optional[float] = 10.0,
                 autocast_dtype: Optional[torch.dtype] = None, graph: bool = True, sync: Optional[GradSync] = None,
                 double_step: bool = False, bucket_mb: float = 16.0, warmup: int = 3):
        self.model, self.optimizer, self.loss_fn = model, optimizer, loss_fn
        self.sync = sync if sync is not None else GradSync(model, bucket_mb=bucket_mb)
        self.grad_clip, self.autocast_dtype, self.double_step = grad_clip, autocast_dtype, double_step
        self.graphed = None
        self.grad_norm = None
        if graph:
            self.graphed = GraphedStep(self._step, warmup=warmup, after_first=self._after_first)
        else:
            self._step()
            self._after_first()

    def _after_first(self):
        self.sync.prune_unused()

    def _step(self):
        self.sync.begin_step()
        if self.autocast_dtype is not None:
            with torch.autocast("cuda", dtype=self.autocast_dtype):
                loss = self.loss_fn()
        else:
            loss = self.loss_fn()
        loss.backward()
        self.sync.finish()
        if self.double_step:
            if isinstance(self.optimizer, FlatAdamW):
                self.optimizer.grad_scale = None  # main.py:244 steps once on the unclipped gradients
            self.optimizer.step()
        if self.grad_clip is not None:
            if isinstance(self.optimizer, FlatAdamW):  # the coefficient is applied inside the optimizer kernel
                self.grad_norm, self.optimizer.grad_scale = self.sync.clip_coef(self.grad_clip)
            else:
                self.grad_norm = self.sync.clip_grad_norm_(self.grad_clip)
        self.optimizer.step()
        return loss.detach()

    def __call__(self) -> torch.Tensor:
        if self.graphed is not None:
            return self.graphed.replay()
        out = self._step()
        invalidate_param_cache()
        return out


def reduce_tensor(t: torch.Tensor, world: Optional[int] = None) -> torch.Tensor:
    """utils/dist_utils.py:41-48: all-reduce (sum) / world of a logged scalar."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return t
    rt = t.clone()
    dist.all_reduce(rt)
    rt /= world or dist.get_world_size()
    return rt
