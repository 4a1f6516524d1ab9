"""si-mamba B200: the spectrally-ordered token encoder path of denix56/SI-Mamba, rebuilt for sm_100a.

Python modules here mirror the reference's module API (same class names, constructor arguments,
forward signatures, config keys and state-dict keys) and call hand-written CUDA kernels through
the C ABI declared in include/simamba.h.  See DESIGN.md.
"""

from .config import Config, finetune_modelnet, finetune_scan_hardest, pretrain  # noqa: F401
from .block import Block, DropPath  # noqa: F401
from .mamba import Mamba  # noqa: F401
from .point_mamba import Encoder, Group, MixerModel, PointMamba, create_block  # noqa: F401
from .mae import MaskMamba_2, MambaDecoder_SST, Point_MAE_Mamba  # noqa: F401
from .seg import get_model, part_seg_config  # noqa: F401
from . import ops  # noqa: F401

__all__ = ["Config", "Block", "DropPath", "Mamba", "Encoder", "Group", "MixerModel", "PointMamba", "create_block",
           "ops", "finetune_modelnet", "finetune_scan_hardest", "pretrain", "MaskMamba_2", "MambaDecoder_SST",
           "Point_MAE_Mamba", "get_model", "part_seg_config"]
