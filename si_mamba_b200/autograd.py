"""Wiring of the mixer body: ``mamba_inner_tm`` is the token-major equivalent of mamba-ssm's ``mamba_inner_fn``
(in_proj -> causal conv1d + SiLU -> x_proj -> dt_proj -> selective scan -> out_proj), every stage on the hand-written
kernels - the projections on the tcgen05 GEMMs (fp32: split planes, csrc/gemm_split3.cu; bf16 autocast: csrc/gemm_bf16.cu),
conv and scan on their TMA kernels - through autograd nodes when a graph is needed.  ``SIM_FP32_GEMM=cublas`` /
``SIM_BF16_GEMM=cublas`` keep F.linear as ablation switches.
"""

from __future__ import annotations

import torch
import torch.nn.functional as F

from . import ops


def _amp_dtype(x: torch.Tensor) -> torch.dtype:
    """Activation dtype of the mixer: the autocast dtype when autocast is on (runner_pretrain.py:243), else x's."""
    if torch.is_autocast_enabled("cuda"):
        return torch.get_autocast_dtype("cuda")
    return x.dtype


class _ParamCache:
    """Inference-only cache of derived parameter tensors (low-precision copies of the projection weights,
    A = -exp(A_log)).  ncu showed 111 cast / elementwise launches per bf16 forward, most of them re-casting the
    same weights; entries are keyed on the parameter's storage and version counter, so an eager optimizer step or a
    load_state_dict invalidates them.  Updates that bump neither key - a CUDA-graph replay that contains
    optimizer.step() (train.GraphedStep), writes through ``p.data`` (EMA swaps) - must call
    ``invalidate_param_cache()``: GraphedStep.replay() does, and every torch optimizer step does through a global
    post-step hook.  Never used while autograd is recording."""

    def __init__(self):
        from torch.utils.weak import WeakIdKeyDictionary  # identity-keyed: tensors do not compare with ==
        self._d = WeakIdKeyDictionary()  # parameter object -> {tag: (version, data_ptr, value)}

    def get(self, p: torch.Tensor, tag, fn):
        per = self._d.get(p)
        if per is None:
            per = self._d[p] = {}
        hit = per.get(tag)
        if hit is not None and hit[0] == p._version and hit[1] == p.data_ptr():
            return hit[2]
        val = fn(p.detach())
        per[tag] = (p._version, p.data_ptr(), val)
        return val


    def get_multi(self, anchor: torch.Tensor, tag, deps, fn):
        """Like get(), for a value derived from several tensors (e.g. conv + BatchNorm folding): valid while the
        versions and addresses of all ``deps`` are unchanged."""
        per = self._d.get(anchor)
        if per is None:
            per = self._d[anchor] = {}
        sig = tuple((t._version, t.data_ptr()) for t in deps)
        hit = per.get(tag)
        if hit is not None and hit[0] == sig:
            return hit[2]
        val = fn()
        per[tag] = (sig, None, val)
        return val


    def clear(self):
        self._d.clear()


_CACHE = _ParamCache()


def invalidate_param_cache(*_args, **_kwargs) -> None:
    """Drop every derived-parameter entry (cast / split weights, A, BatchNorm folds).  Call after parameters or BN
    running statistics changed without their version counters noticing (graph replay, ``.data`` writes)."""
    _CACHE.clear()


try:  # any eager optimizer.step() invalidates too (cheap: the cache is rebuilt lazily by the next eval forward)
    from torch.optim.optimizer import register_optimizer_step_post_hook as _reg_hook
    _reg_hook(invalidate_param_cache)
except ImportError:  # pragma: no cover
    pass
# SIM_FUSE_DT=1: dt_proj inside the scan kernel (inference).  Measured neutral on B200 (106 us fused vs 84 us scan + 22 us
# dt_proj GEMM at the C1 layer shape: the scan is issue-bound, the elementwise warps' extra 140 instructions per tile cost
# what the GEMM launch did), so the separate, simpler path stays the default; the fused kernel saves 100 MB of HBM traffic
# per layer and one launch.
_FUSE_DT = __import__("os").environ.get("SIM_FUSE_DT", "0") == "1"
# SIM_OVERLAP_Z=1: z half of in_proj on a side stream (fp32 inference).  Measured +0.7 % on the C1 forward (the two halves'
# CTAs interleave, so the x half - and the conv behind it - is not done any earlier): off by default.
_OVERLAP_Z = __import__("os").environ.get("SIM_OVERLAP_Z", "0") == "1"
_SIDE = {}


def _side_stream(device):
    """One auxiliary stream per device for the mixer's fork / join (re-entrant: keyed on the device)."""
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=key)
    return _SIDE[key]


_XPROJ_F32A = __import__("os").environ.get("SIM_XPROJ_F32A", "1") != "0"  # x_proj reads fp32 u and splits it in-kernel
_CONV_XPROJ = __import__("os").environ.get("SIM_CONV_XPROJ", "0") != "0"  # causal conv fused into that x_proj kernel (measured slower: 60 us vs 17 + 19)
# fp32 TRAINING projections (forward, dX, dW) on the split GEMM.  Measured on C4 (fp32, 16 clouds, whole step in a CUDA
# graph): HLT 21.8 -> 19.4 ms, SAST 48.0 -> 37.0 ms.  dW contracts over all B*L rows with few output tiles: the kernel's
# split-K both fills the SMs and keeps each tensor-core accumulation chain short (6e-7 against fp64, cuBLAS SGEMM 8e-7).
_TRAIN_X3 = __import__("os").environ.get("SIM_TRAIN_X3", "1") != "0"
_FP32_GEMM = __import__("os").environ.get("SIM_FP32_GEMM", "x3")  # x3 (pre-split tcgen05 kernel) | cublas (ablation)
# bf16 (autocast) projections - inference, training forward, dgrad and wgrad - on the hand-written tcgen05 bf16 kernel
# (csrc/gemm_bf16.cu, operands read in place, MN-major for the backward); "cublas" keeps F.linear as the ablation switch
_BF16_GEMM = __import__("os").environ.get("SIM_BF16_GEMM", "own")


# fp32 inference, in_proj on two fp16 planes (three tensor-core products instead of six, csrc/gemm_split3.cu NP = 2): its
# operand is a LayerNorm output, bounded by sqrt(C) max|gamma| + max|beta|, so the fp16 range (65504) can be checked once
# per parameter version on the host (inproj_f16_ok).  SIM_INPROJ_F16X2=0 keeps the three bf16 planes (ablation).
_INPROJ_F16 = __import__("os").environ.get("SIM_INPROJ_F16X2", "1") != "0"
# fp32 inference: activations the scan would evaluate on its XU pipe (the unit that bounds it) are applied by
# the GEMM epilogue that produces the operand instead.  "z": in_proj writes silu(z) (free inside a tensor-bound kernel);
# "zdt": dt_proj also writes softplus(delta + bias); "0": the scan does both itself (ablation).  Bit-identical results.
_HOIST_ACT = __import__("os").environ.get("SIM_HOIST_ACT", "z")


def inproj_f16_ok(norm_w: torch.Tensor, norm_b: torch.Tensor, in_proj_w: torch.Tensor) -> bool:
    """True when LayerNorm(gamma = norm_w, beta = norm_b) outputs and in_proj_w are provably inside the fp16 range (with a
    2x margin), so in_proj may take two fp16 planes.  One host read per parameter version; while a CUDA graph is being
    captured an uncached answer is "no" (bf16 planes are always valid)."""
    if not _INPROJ_F16 or _OVERLAP_Z:
        return False
    C = norm_w.numel()

    def check():
        if torch.cuda.is_current_stream_capturing():
            return None
        bound = float(C) ** 0.5 * norm_w.detach().abs().max() + norm_b.detach().abs().max()
        return bool((bound < 3.0e4).item() and (in_proj_w.detach().abs().max() < 3.0e4).item())

    ok = _CACHE.get_multi(norm_w, "f16ok", (norm_w, norm_b, in_proj_w), check)
    if ok is None:  # asked during capture before any eager forward: do not cache the refusal
        per = _CACHE._d.get(norm_w)
        if per is not None:
            per.pop("f16ok", None)
        return False
    return ok


class _SplitCols(torch.autograd.Function):
    """x (..., sum(sizes)) -> views x[..., a:b] per size; the backward writes the pieces' gradients side by side with
    one torch.cat (autograd's own SliceBackward allocates a zero tensor of the full shape per slice and then adds them:
    10 % of the C2 training step for the in_proj output alone)."""

    @staticmethod
    def forward(ctx, x, *sizes):
        ctx.sizes = sizes
        ctx.meta = (x.shape, x.dtype, x.device)
        outs, a = [], 0
        for n in sizes:
            outs.append(x[..., a:a + n])
            a += n
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        shape, dtype, device = ctx.meta
        parts = [g if g is not None else torch.zeros(*shape[:-1], n, dtype=dtype, device=device)
                 for g, n in zip(grads, ctx.sizes)]
        return (torch.cat(parts, dim=-1),) + (None,) * len(ctx.sizes)


def wants_split3(hidden_dtype: torch.dtype, in_proj_w: torch.Tensor, d_model: int, params=()) -> bool:
    """True when the mixer would consume its input as a Split3 (fp32 inference on the x3 GEMM path): the Block's
    fused add + LayerNorm then writes the split planes directly instead of an fp32 tensor.  ``params`` = every mixer
    parameter: the predicate is mamba_inner_tm's own ``need_grad`` test, so a partially frozen mixer never gets one."""
    return (_FP32_GEMM != "cublas" and hidden_dtype == torch.float32 and in_proj_w.dtype == torch.float32
            and in_proj_w.is_cuda and d_model % 8 == 0
            and not (torch.is_grad_enabled() and any(t.requires_grad for t in (in_proj_w, *params))))


def mamba_inner_tm(hidden, in_proj_w, conv_w, conv_b, x_proj_w, dt_proj_w, dt_proj_b, A_log, D, out_proj_w,
                   dt_rank: int, d_state: int) -> torch.Tensor:
    """Token-major Mamba mixer body.  hidden (B, L, d_model) -> (B, L, d_model)."""
    d_inner = conv_w.shape[0]
    act = _amp_dtype(hidden) if not isinstance(hidden, ops.Split3) else torch.float32
    need_grad = torch.is_grad_enabled() and any(
        t.requires_grad for t in (hidden, in_proj_w, conv_w, conv_b, x_proj_w, dt_proj_w, dt_proj_b, A_log, D,
                                  out_proj_w))
    own_bf16 = (act == torch.bfloat16 and _BF16_GEMM != "cublas" and not isinstance(hidden, ops.Split3) and hidden.is_cuda
                and all(w.shape[1] % 8 == 0 and w.shape[0] % 4 == 0 for w in (in_proj_w, x_proj_w, dt_proj_w, out_proj_w)))
    if need_grad and own_bf16:
        # LinearBF16 takes the fp32 master weights themselves (one cast per call) and returns their gradients in fp32
        w_in, w_x, w_dt, w_out = in_proj_w, x_proj_w, dt_proj_w, out_proj_w
        A = -torch.exp(A_log.float())
    elif need_grad:
        w_in, w_x, w_dt, w_out = (w.to(act) for w in (in_proj_w, x_proj_w, dt_proj_w, out_proj_w))
        A = -torch.exp(A_log.float())
    else:
        cast = (lambda w: w) if act == in_proj_w.dtype else (lambda w: _CACHE.get(w, act, lambda t: t.to(act)))
        w_in, w_x, w_dt, w_out = (cast(w) for w in (in_proj_w, x_proj_w, dt_proj_w, out_proj_w))
        A = _CACHE.get(A_log, "A", lambda t: -torch.exp(t.float()))
    # fp32 inference: the four projections run on the tensor cores with fp32-accurate 3 x bf16 operand splitting.
    # x3: hand-written TMA + tcgen05 kernel on pre-split planes (weights split once and cached, 3.7x cuBLAS SGEMM on
    # in_proj); cublas: F.linear (SIMT SGEMM), kept as the ablation switch.
    linear = ops.linear_bf16 if own_bf16 else F.linear
    x3 = False
    if not need_grad and act == torch.float32 and hidden.is_cuda and _FP32_GEMM != "cublas":
        x3 = d_inner % 64 == 0  # every producer below then emits the split operand of the next projection itself

        def linear(x, w):
            return ops.linear_f32_x3(x, _CACHE.get(w, "x3", ops.split3), w.shape[1])
    if (need_grad and _TRAIN_X3 and act == torch.float32 and hidden.is_cuda and in_proj_w.dtype == torch.float32
            and _FP32_GEMM != "cublas" and d_inner % 8 == 0 and hidden.shape[-1] % 8 == 0 and dt_rank % 4 == 0):
        linear = ops.linear_x3_train  # fp32 training: forward, dX and dW GEMMs on the split-plane tensor-core kernel
    join_z = None
    hoist_z = False
    if x3 and _OVERLAP_Z:
        # in_proj as two GEMMs: the x half on this stream, the z half (only needed by the scan) on a side stream where it
        # runs next to the HBM-bound conv and the small x_proj / dt_proj GEMMs; a fork / join that CUDA graphs capture
        hs = hidden if isinstance(hidden, ops.Split3) else ops.Split3(ops.split3(hidden.to(act)), hidden.shape)
        wp = _CACHE.get(in_proj_w, "x3", ops.split3)  # (3, 2*d_inner, d_model)
        K = in_proj_w.shape[1]
        cur, side = torch.cuda.current_stream(), _side_stream(hidden.planes.device if isinstance(hidden, ops.Split3) else hidden.device)
        fork = torch.cuda.Event()
        fork.record(cur)
        x = ops.linear_split3(hs.planes, wp[:, :d_inner], K).view(*hs.shape[:-1], d_inner)
        side.wait_event(fork)
        with torch.cuda.stream(side):
            z = ops.linear_split3(hs.planes, wp[:, d_inner:], K).view(*hs.shape[:-1], d_inner)
        join_z = (cur, side, z)
    elif x3 and isinstance(hidden, ops.Split3):
        # in_proj from the planes the Block's add + LayerNorm wrote (two fp16 planes where inproj_f16_ok, else three bf16);
        # with the hoist its epilogue turns the z half into the gate silu(z) the scan multiplies by
        hoist_z = _HOIST_ACT in ("z", "zdt") and not _FUSE_DT
        f16 = hidden.planes.dtype == torch.float16
        wp = _CACHE.get(in_proj_w, "x2h", ops.split2h) if f16 else _CACHE.get(in_proj_w, "x3", ops.split3)
        xz = ops.linear_f32_x3(hidden, wp, in_proj_w.shape[1], act="silu_from" if hoist_z else None, act_col0=d_inner)
        x, z = xz[..., :d_inner], xz[..., d_inner:]
    elif own_bf16 and not need_grad and _HOIST_ACT in ("z", "zdt") and not _FUSE_DT:
        # bf16 inference: the same hoist - in_proj's epilogue turns the z half into the gate (silu on the fp32 accumulator,
        # then the bf16 rounding), the scan multiplies by it as is
        hb = hidden.to(act)
        h2 = hb.reshape(-1, hb.shape[-1]) if hb.is_contiguous() else ops._as_rows(hb)
        xz = ops.gemm_bf16(h2, w_in, silu_col0=d_inner).view(*hb.shape[:-1], 2 * d_inner)
        x, z = xz[..., :d_inner], xz[..., d_inner:]
        hoist_z = True
    else:
        xz = linear(hidden if isinstance(hidden, ops.Split3) else hidden.to(act), w_in)  # (B, L, 2*d_inner)
        if need_grad:
            x, z = _SplitCols.apply(xz, d_inner, d_inner)
        else:
            x, z = xz[..., :d_inner], xz[..., d_inner:]
    if need_grad:
        u = u_op = ops.CausalConv1dTM.apply(x, conv_w, conv_b, True)
    elif (x3 and _XPROJ_F32A and _CONV_XPROJ and dt_rank <= 32 and 32 <= x_proj_w.shape[0] <= 64 and conv_w.shape[-1] == 4
          and d_inner % 8 == 0 and d_inner <= 1024 and x.dtype == torch.float32):
        # one kernel: conv + SiLU in the x_proj GEMM's transform warps (u written for the scan on the way)
        wxp = _CACHE.get(x_proj_w, "x3", ops.split3)
        u, x_dbl_f, dt_planes_f = ops.conv_xproj_f32(x, conv_w, conv_b, wxp, 32)
        u_op = None
    elif x3 and _XPROJ_F32A and dt_rank <= 32 and 32 <= x_proj_w.shape[0] <= 64:
        u = u_op = ops.causal_conv1d_tm(x, conv_w, conv_b, silu=True)  # x_proj splits u itself (gemm_f32a kernel)
    elif x3:
        u, u_op = ops.causal_conv1d_tm(x, conv_w, conv_b, silu=True, split=True)
    else:
        u = u_op = ops.causal_conv1d_tm(x, conv_w, conv_b, silu=True)
    dt_planes = None
    dt_final = False
    if u_op is None:
        x_dbl, dt_planes = x_dbl_f, dt_planes_f
    elif x3 and _XPROJ_F32A and not isinstance(u_op, ops.Split3) and dt_rank <= 32 and 32 <= x_proj_w.shape[0] <= 64:
        wxp = _CACHE.get(x_proj_w, "x3", ops.split3)
        x_dbl, dt_planes = ops.linear_f32a_planes_out(u_op, wxp, x_proj_w.shape[1], 32)
        x_dbl = x_dbl.view(*u_op.shape[:-1], x_proj_w.shape[0])
    elif x3 and isinstance(u_op, ops.Split3) and dt_rank <= 32 and x_proj_w.shape[0] >= 32 and x_proj_w.shape[0] <= 64:
        # x_proj whose epilogue also writes the dt_proj operand (split planes of the first 32 output columns)
        wxp = _CACHE.get(x_proj_w, "x3", ops.split3)
        x_dbl, dt_planes = ops.linear_split3_planes_out(u_op.planes, wxp, x_proj_w.shape[1], 32)
        x_dbl = x_dbl.view(*u_op.shape[:-1], x_proj_w.shape[0])
    else:
        x_dbl = linear(u_op, w_x)  # (B, L, dt_rank + 2*d_state)
    if join_z is None and (not need_grad and _FUSE_DT and hidden.is_cuda and dt_rank == 24 and d_state == 16 and d_inner % 64 == 0
            and u.dtype in (torch.float32, torch.bfloat16) and x_dbl.dtype == u.dtype):
        # inference: dt_proj runs inside the scan kernel (mma.sync in its elementwise warps); delta never touches HBM
        planes = _CACHE.get(dt_proj_w, ("dtp", u.dtype), lambda t: ops.dt_proj_planes(t, u.dtype))
        y = ops.selective_scan_fused_dt_tm(u, x_dbl, dt_rank, planes, A, D, z, dt_proj_b, True, split=x3)
        return linear(y, w_out)
    if x3 and dt_rank <= 32 and x_dbl.shape[-1] >= 32:
        # dt_proj with K padded to 32: a 64-byte row pitch is what the TMA tensor copies like (19.5 -> 13.6 us).  The
        # padding columns of the weight are zero, so the operand is simply the first 32 columns of x_dbl (the extra 8 are
        # B values that meet zero weights) - no padded copy of the activations is needed.
        wdt32 = _CACHE.get(dt_proj_w, "x3_pad32",
                           lambda t: ops.split3(F.pad(t.float(), (0, 32 - t.shape[1])).contiguous()))
        if dt_planes is None:
            dt_planes = ops.split3(x_dbl[..., :32])
        if _HOIST_ACT == "zdt" and dt_proj_b is not None:  # dt = softplus(delta + bias) leaves the GEMM's epilogue
            dt = ops.linear_split3(dt_planes, wdt32, 32, act="softplus_bias", bias=dt_proj_b).view(*x_dbl.shape[:-1], d_inner)
            dt_final = True
        else:
            dt = ops.linear_split3(dt_planes, wdt32, 32).view(*x_dbl.shape[:-1], d_inner)
        Bm = x_dbl[..., dt_rank:dt_rank + d_state]
        Cm = x_dbl[..., dt_rank + d_state:]
    elif need_grad and x_dbl.shape[-1] == dt_rank + 2 * d_state:
        dt_in, Bm, Cm = _SplitCols.apply(x_dbl, dt_rank, d_state, d_state)
        dt = linear(dt_in, w_dt)  # bias is applied inside the scan
    else:
        dt = linear(x_dbl[..., :dt_rank], w_dt)  # bias is applied inside the scan
        Bm = x_dbl[..., dt_rank:dt_rank + d_state]
        Cm = x_dbl[..., dt_rank + d_state:]
    if join_z is not None:
        cur, side, z = join_z
        cur.wait_stream(side)
        z.record_stream(cur)
        hs.planes.record_stream(side)
    if need_grad:
        y = ops.SelectiveScanTM.apply(u, dt, A, Bm, Cm, D, z, dt_proj_b, True)
    else:
        y = ops.selective_scan_tm(u, dt, A, Bm, Cm, D, z, None if dt_final else dt_proj_b, delta_softplus=not dt_final,
                                  split=x3, z_gate=hoist_z)
    return linear(y, w_out)
