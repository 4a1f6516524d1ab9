"""CPU oracle for the SI-Mamba spectrally-ordered token encoder path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``si_mamba_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or
as the timed CPU reference arm - never as the product path.

What it is: a plain PyTorch-CPU (fp32, with fp64 where noted) restatement of
the reference functions listed in SURVEY.md section 8(a).  All ``file:line``
citations are relative to the upstream repository denix56/SI-Mamba.

PARITY UNPINNED: the reference ships no tests, golden vectors
or fixtures, and its arithmetic lives in third-party wheels that are absent
from the reference tree and from this image:

  * mamba-ssm  (README.md:56 pins ==1.1.1)   -> oracle/mamba.py
  * causal-conv1d (README.md:55, ==1.1.1)    -> oracle/mamba.py
  * pytorch3d  (README.md:29, unpinned)      -> oracle/tokenizer.py
  * torch.linalg.eigh (LAPACK, in image)     -> oracle/spectral.py (fp64)

The pins this build creates for itself are (a) committed golden vectors in
``tests/golden`` produced by ``tools/make_golden.py`` from this restatement,
(b) ``torch.linalg.eigh`` in fp64 as the eigen-oracle and (c) HuggingFace
``transformers`` ``MambaMixer.slow_forward`` as an independent second witness
for the mixer (tests/test_oracle_mamba.py).
"""

from . import tokenizer, spectral, mamba, model, mae, seg  # noqa: F401
