"""CPU oracle for the SI-Mamba spectrally-ordered token encoder path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``si_mamba_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or
as the timed CPU reference arm - never as the product path.

What it is: a plain PyTorch-CPU (fp32, with fp64 where noted) restatement of
the reference functions listed in SURVEY.md section 8(a).  All ``file:line``
citations are relative to the upstream repository denix56/SI-Mamba.

PIN STATUS (DESIGN.md section 2 has the table).  The reference ships no tests,
golden vectors or fixtures, so the oracle is pinned against OUTPUTS OF THE
REFERENCE ITSELF where the reference is plain PyTorch: ``tools/
make_reference_golden.py`` imports /root/reference unmodified in the build
container and commits what its own methods return on seeded inputs
(tests/golden/reference_*.pt; checked by tests/test_oracle_vs_reference.py):

  * pinned: patch graph, Laplacian + eigenpairs (loop / batched / symmetric),
    spectral sort + gather, multilevel codes, MAE masked sort / restore helpers,
    Encoder, Block residual plumbing, seg feature propagation
    -> oracle/spectral.py, oracle/mae.py, oracle/model.py (encoder),
       oracle/mamba.py (mixer_model loop), oracle/seg.py (feature_propagation)

PARITY UNPINNED for the rows whose arithmetic lives in third-party CUDA-only
wheels that are absent from the reference tree and from this image:

  * mamba-ssm  (README.md:56 pins ==1.1.1)   -> oracle/mamba.py (mixer, scan)
  * causal-conv1d (README.md:55, ==1.1.1)    -> oracle/mamba.py (conv1d)
  * pytorch3d  (README.md:29, unpinned)      -> oracle/tokenizer.py (FPS, kNN),
                                                oracle/mae.py (chamfer_l2)
  * pointnet2_ops (README.md:49)             -> oracle/tokenizer.py (fps_pointnet2)

Those are restated from the published algorithms; their self-made pins are
committed golden vectors from ``tools/make_golden.py``, brute-force scalar
restatements in tests/test_oracle.py, and HuggingFace ``transformers``
``MambaMixer.slow_forward`` as an independent witness for the mixer.
"""

from . import tokenizer, spectral, mamba, model, mae, seg  # noqa: F401
