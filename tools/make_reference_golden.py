"""Golden vectors produced by the UNMODIFIED reference, run in the build container.

Imports /root/reference/models/point_mamba.py as it lies on disk and calls its own pure-PyTorch methods of the
spectral-ordering path on seeded inputs; the results are committed as tests/golden/reference_*.pt so that
tests/test_oracle.py can pin oracle/ against the reference itself (the GPU box has no /root/reference).

How the reference is made to run here without touching it:
  * its third-party imports that are CUDA-only wheels (mamba_ssm, pytorch3d, knn_cuda, pointnet2_ops, timm ...) are
    replaced by empty stub modules - none of them is called by the methods exercised below;
  * the methods hard-code `.cuda()` / `device='cuda'`; a TorchFunctionMode rewrites those requests to the CPU.
    The arithmetic (torch ops, LAPACK syevd behind torch.linalg.eigh) is the reference's own.

Not reachable this way (their arithmetic lives in absent wheels): Group.forward (pytorch3d), Mamba (mamba-ssm),
misc.fps (pointnet2_ops).  Those stay "parity unpinned" (DESIGN.md section 2).

    python tools/make_reference_golden.py        # writes tests/golden/reference_spectral.pt, reference_mae.pt
"""

from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import os
import sys
import types
from unittest import mock

import torch
from torch.overrides import TorchFunctionMode

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
KEEP = 8

STUB_ROOTS = ["mamba_ssm", "pytorch3d", "knn_cuda", "pointnet2_ops", "timm", "causal_conv1d", "easydict",
              "tensorboardX", "termcolor", "open3d", "matplotlib", "cv2", "h5py", "sklearn_extra", "pywt", "ot",
              "torch_geometric", "torch_scatter", "torch_cluster", "geomloss", "pykeops"]


class _StubLoader(importlib.abc.Loader):
    def create_module(self, spec):
        m = types.ModuleType(spec.name)
        m.__path__ = []
        m.__getattr__ = lambda name: mock.MagicMock(name=f"{spec.name}.{name}")  # type: ignore[assignment]
        return m

    def exec_module(self, module):
        pass


class _StubFinder(importlib.abc.MetaPathFinder):
    """Serve empty modules for the absent third-party roots (and only for those)."""

    def find_spec(self, name, path=None, target=None):
        if name.split(".")[0] in STUB_ROOTS:
            try:  # prefer the real thing if the image has it
                sys.meta_path.remove(self)
                try:
                    real = importlib.util.find_spec(name)
                except (ImportError, ValueError, ModuleNotFoundError):
                    real = None
                finally:
                    sys.meta_path.insert(0, self)
                if real is not None:
                    return None
            except ValueError:
                pass
            return importlib.machinery.ModuleSpec(name, _StubLoader(), is_package=True)
        return None


class CudaToCpu(TorchFunctionMode):
    """Run code that hard-codes the 'cuda' device on the CPU, arithmetic untouched."""

    @staticmethod
    def _fix(v):
        if isinstance(v, str) and v.startswith("cuda"):
            return "cpu"
        if isinstance(v, torch.device) and v.type == "cuda":
            return torch.device("cpu")
        return v

    def __torch_function__(self, func, types_, args=(), kwargs=None):
        kwargs = dict(kwargs or {})
        if func is torch.Tensor.cuda:
            return args[0]
        if "device" in kwargs:
            kwargs["device"] = self._fix(kwargs["device"])
        if func is torch.Tensor.to:
            args = tuple(self._fix(a) for a in args)
        return func(*args, **kwargs)


def load_reference():
    sys.meta_path.insert(0, _StubFinder())
    sys.path.insert(0, REF)
    import importlib.util  # noqa: F401
    for _ in range(40):  # any further absent third-party root is stubbed too (printed, so the list is auditable)
        try:
            return importlib.import_module("models.point_mamba")
        except ModuleNotFoundError as e:
            root = e.name.split(".")[0]
            if root in STUB_ROOTS or root in ("models", "utils"):
                raise
            print("stubbing absent module:", root)
            STUB_ROOTS.append(root)
            for k in [k for k in sys.modules if k.split(".")[0] in ("models", "utils")]:
                del sys.modules[k]
    raise RuntimeError("reference import did not converge")


def seeded_centres(B, G, seed):
    """Patch centres the way the product sees them: FPS-like well-spread points of a synthetic cloud (plain torch)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, G, 3, generator=g)
    x = x / x.norm(dim=-1, keepdim=True).clamp_min(1e-6) * torch.rand(B, G, 1, generator=g) ** (1 / 3)
    return x.contiguous()


def main():
    pm = load_reference()
    torch.manual_seed(0)
    out_spec = {"cases": []}
    me = types.SimpleNamespace()  # the methods below read `self.alpha` at most

    # (graph builder, knn, alpha, symmetric, self_loop, binary): finetune_modelnet.yaml / pretrain.yaml / seg settings
    graph_cfgs = [
        ("feature_space", 20, 1.0, True, False, True),
        ("feature_space", 20, 100.0, True, False, False),
        ("feature_space", 10, 1.0, False, False, True),
        ("feature_space", 8, 10.0, True, True, False),
        ("centers", 20, 1.0, True, False, True),
        ("centers", 20, 100.0, True, False, False),
        ("centers", 10, 0.0, True, False, False),  # alpha == 0: sigma branch
    ]
    with CudaToCpu():
        for gi, (builder, knn, alpha, sym, loop, binary) in enumerate(graph_cfgs):
            for G, B in ((64, 4), (128, 2)):
                centre = seeded_centres(B, G, 100 + gi * 7 + G)
                me.alpha = alpha  # create_graph_from_centers branches on self.alpha == 0 (:647)
                if builder == "feature_space":
                    adj = pm.PointMamba.create_graph_from_feature_space_gpu_weighted_adjacency(
                        me, centre, knn, alpha, sym, loop, binary)
                else:
                    adj = pm.PointMamba.create_graph_from_centers(me, centre, knn, alpha, sym, loop, binary)
                case = {"builder": builder, "knn": knn, "alpha": alpha, "symmetric": sym, "self_loop": loop,
                        "binary": binary, "centre": centre, "adj": adj.clone()}
                for smallest in (True, False):
                    k = 4
                    vals, vecs, vals_all, vecs_all = pm.PointMamba.calc_top_k_eigenvalues_eigenvectors(
                        me, adj, k, smallest)
                    case[f"loop_vals_{int(smallest)}"] = vals.clone()
                    case[f"loop_vecs_{int(smallest)}"] = vecs.clone()
                    bvals, bvecs, _, _ = pm.Point_MAE_Mamba.calc_top_k_eigenvalues_eigenvectors(
                        me, adj, k, smallest)
                    case[f"batched_vals_{int(smallest)}"] = bvals.clone()
                    case[f"batched_vecs_{int(smallest)}"] = bvecs.clone()
                    if sym:
                        svals, svecs, _, _ = pm.PointMamba.calc_top_k_eigenvalues_eigenvectors_symmetric(
                            me, adj, k, smallest)
                        case[f"sym_vals_{int(smallest)}"] = svals.clone()
                        case[f"sym_vecs_{int(smallest)}"] = svecs.clone()
                case["vals_all"] = vals_all.clone()
                # ordering of 384-wide tokens by each eigenvector (:817-826) and the multilevel code (:829-841)
                g = torch.Generator().manual_seed(gi * 13 + G)
                tokens = torch.randn(B, G, 384, generator=g)
                vecs = case["loop_vecs_1"]
                # (row gathers: the first KEEP channels are stored, enough to identify every row, to keep fixtures small)
                case["tokens"] = tokens[..., :KEEP].clone()
                case["sorted_tokens"] = torch.stack(
                    [pm.PointMamba.sort_points_by_fiedler(me, tokens, vecs[:, :, i])[..., :KEEP].clone()
                     for i in range(vecs.shape[-1])])
                case["multilevel"] = torch.stack(
                    [pm.PointMamba.multilevel_travers(me, vecs, lvl).reshape(B, G) for lvl in (1, 2, 3, 4)])
                out_spec["cases"].append(case)

        # MAE masked sort / token restore helpers (:2639-2670, :2697-2715): G=64, 38 masked (mask_ratio 0.6)
        out_mae = {"cases": []}
        for seed in (1, 2, 3):
            B, G = 4, 64
            g = torch.Generator().manual_seed(seed)
            tokens = torch.randn(B, G, 384, generator=g)
            fied = torch.randn(B, G, generator=g)
            mask = torch.zeros(B, G, dtype=torch.bool)
            for b in range(B):
                mask[b, torch.randperm(G, generator=g)[:38]] = True
            s_tok, s_mask, s_learn, s_idx = pm.MaskMamba_3.sort_points_by_fiedler(me, tokens, mask, fied)
            # find_indices_vectorized: position of each of a[b, :] inside the row of learnable-token indices
            a = torch.stack([s_learn[b][torch.randperm(38, generator=g)] for b in range(B)])
            pos = pm.MaskMamba_3.find_indices_vectorized(me, a, s_learn)
            nb = torch.randn(B, G, 32, 3, generator=g)
            s_nb = pm.MaskMamba_3.sort_points_by_fiedler_for_neighberhood(me, nb, fied)
            out_mae["cases"].append({"tokens": tokens[..., :KEEP].clone(), "fiedler": fied, "mask": mask,
                                     "sorted_tokens": s_tok[..., :KEEP].clone(),
                                     "sorted_mask": s_mask, "sorted_learnable": s_learn, "sorted_indices": s_idx,
                                     "a": a, "found": pos, "neighborhood": nb, "sorted_neighborhood": s_nb})

    os.makedirs(OUT, exist_ok=True)
    torch.save(out_spec, os.path.join(OUT, "reference_spectral.pt"))
    torch.save(out_mae, os.path.join(OUT, "reference_mae.pt"))
    for f in ("reference_spectral.pt", "reference_mae.pt"):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
