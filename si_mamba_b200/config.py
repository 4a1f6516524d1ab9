"""Minimal attribute-style config (the reference passes EasyDict objects built from cfgs/*.yaml).

Only the config *keys* of the hot path are kept; the YAML / registry
machinery of the reference (utils/config.py, utils/registry.py) is out of scope.
"""

from __future__ import annotations


class Config(dict):
    """dict with attribute access, enough of EasyDict for ``config.key`` and ``"key" in config``."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return Config(v) if isinstance(v, dict) and not isinstance(v, Config) else v

    def __setattr__(self, k, v):
        self[k] = v


def finetune_modelnet() -> Config:
    """model section of cfgs/finetune_modelnet.yaml:23-50 (config C1 of BASELINE.json)."""
    return Config(
        NAME="PointMamba", trans_dim=384, depth=12, cls_dim=40, num_heads=6, group_size=32, num_group=64,
        encoder_dims=384, rms_norm=False, drop_path=0.3, drop_out=0.0, method="SAST", reverse=True,
        reverse_2=False, reverse_3=False, knn_graph=20, k_top_eigenvectors=4, alpha=100.0, smallest=True,
        symmetric=True, self_loop=False, binary=True, matrix="laplacian", add_after_layer=False, rotation=False)


def finetune_scan_hardest() -> Config:
    """model section of cfgs/finetune_scan_hardest.yaml:22-49 (config C2)."""
    c = finetune_modelnet()
    c.update(cls_dim=15, num_group=128, drop_path=0.1, alpha=10.0, rotation=True)
    return c


def pretrain() -> Config:
    """model section of cfgs/pretrain.yaml:35-65 (config C3: MAE on ShapeNet55)."""
    return Config(
        NAME="Point_MAE_Mamba", group_size=32, num_group=64, loss="cdl2", rms_norm=False, use_cls_token=False,
        drop_path=0.1, drop_out=0.1,
        transformer_config=Config(
            mask_ratio=0.6, mask_type="rand", trans_dim=384, encoder_dims=384, depth=12, drop_path_rate=0.1,
            num_heads=6, decoder_depth=4, decoder_num_heads=6,
            method="smallest_eigenvectors_seperate_learnable_tokens", reverse=True, knn_graph=20,
            k_top_eigenvectors=4, smallest=True, alpha=10, symmetric=True, self_loop=False, binary=True))
