"""ctypes binding of libsimamba_b200.so (the C ABI in include/simamba.h).

There is NO fallback: if the shared library is missing, or a call returns a
non-zero status, a RuntimeError is raised.  The library is built in-tree by
``python -m si_mamba_b200.build`` (``__graft_entry__.build()``).
"""

from __future__ import annotations

import ctypes as C
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "libsimamba_b200.so"

SIM_F32, SIM_BF16 = 0, 1

# sim_spectral_flags
GRAPH_SYMMETRIC = 1 << 0
GRAPH_SELF_LOOP = 1 << 1
GRAPH_BINARY = 1 << 2
EIG_SMALLEST = 1 << 3
LAP_SYMMETRIC = 1 << 4
LAP_EPS_CLAMP = 1 << 5
EIG_CANONICAL_SIGN = 1 << 6
# sim_distance_flags
DIST_FMA = 1

_p, _i, _l, _f, _sz = C.c_void_p, C.c_int, C.c_long, C.c_float, C.c_size_t

# name -> (restype, argtypes); must list every symbol include/simamba.h declares
SIGNATURES = {
    "sim_version": (_i, []),
    "sim_last_error_string": (C.c_char_p, []),
    "sim_fps": (_i, [_p, _i, _i, _i, _p, _p, _p]),
    "sim_knn_group": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _p, _p]),
    "sim_fps_ex": (_i, [_p, _i, _i, _i, _p, _p, _i, _p]),
    "sim_knn_group_ex": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _p, _i, _p]),
    "sim_spectral_eig_workspace_bytes": (_sz, [_i, _i, _i]),
    "sim_spectral_eig": (_i, [_p, _i, _i, _i, _f, _i, _i, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "sim_spectral_eig_ex": (_i, [_p, _p, _p, _i, _i, _i, _f, _i, _i, _i, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "sim_pairwise_dist_mean": (_i, [_p, _i, _i, _p, _p, _p]),
    "sim_split3_bf16_t": (_i, [_p, _l, _i, _i, _p, _l, _l, _p]),
    "sim_gemm_bf16": (_i, [_p, _l, _i, _p, _l, _i, _p, _l, _i, _i, _i, _i, _i, _p]),
    "sim_add_layernorm_droppath": (_i, [_p, _p, _i, _p, _p, _p, _p, _p, _l, _i, _f, _i, _i, _p]),
    "sim_add_layernorm_bwd_dx": (_i, [_p, _p, _p, _p, _p, _i, _p, _p, _i, _p, _p, _l, _i, _f, _i, _p]),
    "sim_adamw_flat": (_i, [_p, _p, _p, _p, _p, _l, _p, _p, _p, _f, _f, _f, _p]),
    "sim_point_linear3": (_i, [_p, _p, _p, _p, _l, _i, _i, _p]),
    "sim_gemm_tf32": (_i, [_p, _l, _i, _p, _l, _i, _p, _l, _i, _i, _i, _i, _i, _p, _i, _p]),
    "sim_gemm_tf32_group": (_i, [_p, _l, _p, _l, _p, _l, _i, _i, _i, _p, _p, _l, _i, _p, _l, _p]),
    "sim_gemm_bf16_silu": (_i, [_p, _l, _p, _l, _p, _l, _i, _i, _i, _i, _i, _p]),
    "sim_argsort_rows": (_i, [_p, _l, _l, _i, _i, _p, _p, _p]),
    "sim_order_gather_fwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "sim_order_gather_bwd": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "sim_gather_rows": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "sim_add_layernorm": (_i, [_p, _p, _p, _p, _p, _p, _p, _l, _i, _f, _i, _i, _p]),
    "sim_causal_conv1d_fwd": (_i, [_p, _l, _p, _p, _p, _l, _i, _i, _i, _i, _i, _i, _p]),
    "sim_selective_scan_fwd": (_i, [_p, _l, _p, _l, _p, _p, _l, _p, _l, _p, _p, _l, _p, _p, _l, _p,
                                    _i, _i, _i, _i, _i, _i, _i, _p]),
    "sim_selective_scan_checkpoint_bytes": (_sz, [_i, _i, _i]),
    "sim_selective_scan_bwd": (_i, [_p, _l, _p, _l, _p, _p, _l, _p, _l, _p, _p, _l, _p, _p, _l, _p,
                                    _p, _l, _p, _l, _p, _l, _p, _p, _p, _p, _p,
                                    _i, _i, _i, _i, _i, _i, _p]),
    "sim_add_layernorm_split3": (_i, [_p, _p, _p, _p, _p, _p, _p, _l, _l, _i, _f, _i, _p]),
    "sim_causal_conv1d_fwd_split3": (_i, [_p, _l, _p, _p, _p, _l, _p, _l, _l, _i, _i, _i, _i, _i, _p]),
    "sim_selective_scan_fwd_split3": (_i, [_p, _l, _p, _l, _p, _p, _l, _p, _l, _p, _p, _l, _p, _p, _l, _l,
                                           _i, _i, _i, _i, _i, _p]),
    "sim_mae_index_maps": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "sim_mae_compact_fwd": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "sim_mae_compact_bwd": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "sim_mae_restore_fwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "sim_mae_restore_bwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "sim_gather_sum_rows": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "sim_invert_row_map": (_i, [_p, _i, _i, _i, _i, _p, _p, _p]),
    "sim_spectral_perm": (_i, [_p, _l, _l, _i, _i, _p, _p, _p]),
    "sim_three_nn_interp_fwd": (_i, [_p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p]),
    "sim_three_interp_bwd": (_i, [_p, _p, _p, _i, _i, _i, _i, _p, _p]),
    "sim_chamfer_l2_fwd": (_i, [_p, _p, _l, _i, _i, _p, _p, _p, _p]),
    "sim_chamfer_l2_bwd": (_i, [_p, _p, _p, _p, _p, _l, _i, _i, _p, _p, _p]),
    "sim_fps_pointnet2": (_i, [_p, _i, _i, _i, _p, _p, _p]),
    "sim_add_layernorm_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _l, _i, _f, _i, _p]),
    "sim_selective_scan_fwd_fused_dt": (_i, [_p, _l, _p, _l, _i, _p, _p, _p, _p, _l, _p, _p, _l, _p, _l, _l,
                                             _i, _i, _i, _i, _i, _i, _p]),
    "sim_gemm_bf16x3_split_out": (_i, [_p, _l, _l, _p, _l, _l, _p, _l, _i, _i, _i, _p, _i, _l, _l, _p]),
    "sim_gemm_f32a_bf16x3": (_i, [_p, _l, _p, _l, _l, _p, _l, _i, _i, _i, _p, _i, _l, _l, _p]),
    "sim_conv_xproj_f32": (_i, [_p, _l, _p, _p, _p, _l, _p, _l, _l, _p, _l, _i, _i, _i, _i, _p, _i, _l, _l, _p]),
    "sim_group_max": (_i, [_p, _p, _l, _i, _i, _i, _p]),
    "sim_group_bias_relu": (_i, [_p, _p, _l, _i, _i, _i, _p]),
    "sim_mlp3_relu_rows": (_i, [_p, _l, _l, _i, _p, _p, _i, _p, _p, _i, _p, _p, _i, _p, _l, _p]),
    "sim_layernorm_mean": (_i, [_p, _p, _p, _p, _i, _i, _i, _f, _p]),
    "sim_split3_bf16": (_i, [_p, _l, _i, _i, _p, _l, _l, _p]),
    "sim_split2_f16": (_i, [_p, _l, _i, _i, _p, _l, _l, _p]),
    "sim_gemm_planes": (_i, [_i, _p, _l, _l, _p, _l, _l, _p, _l, _i, _i, _i, _i, _i, _p, _p]),
    "sim_add_layernorm_split2h": (_i, [_p, _p, _p, _p, _p, _p, _p, _l, _l, _i, _f, _i, _p]),
    "sim_gemm_bf16x3": (_i, [_p, _l, _l, _p, _l, _l, _p, _l, _i, _i, _i, _p]),
    "sim_causal_conv1d_bwd": (_i, [_p, _l, _p, _p, _p, _l, _p, _l, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
}

_lib = None


def lib_path() -> Path:
    return _LIB_PATH


def load():
    """Load the shared library (once) and bind every entry point.  Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise RuntimeError(
            f"{_LIB_PATH} is missing: build it with `python -m si_mamba_b200.build` "
            "(there is no CPU or PyTorch fallback for the si-mamba hot path)")
    lib = C.CDLL(str(_LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class SimError(RuntimeError):
    pass


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().sim_last_error_string().decode(errors="replace")
        raise SimError(f"{what} failed with status {status}: {msg}")


def launches() -> int:
    """Number of kernels this process has launched through the C ABI (bench.py's gpu_launches claim)."""
    return _launch_count[0]


_launch_count = [0]


def call(name: str, *args) -> None:
    _launch_count[0] += 1
    check(getattr(load(), name)(*args), name)
