"""In-tree nvcc build of libsimamba_b200.so for sm_100a (cross-compiles without a GPU).

    python -m si_mamba_b200.build [--force]

One object per csrc/*.cu, compiled in parallel, linked with the static CUDA runtime
into si_mamba_b200/libsimamba_b200.so.  The .so is git-ignored but travels to the
GPU box with the repo snapshot.  ptxas -v output (registers / spills / smem per
kernel) is kept in build/ptxas.log.
"""

from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
BUILD = ROOT / "build"
LIB = PKG / "libsimamba_b200.so"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr"]


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(ARCH + CFLAGS).encode())
    return h.hexdigest()


def sources():
    return sorted(CSRC.glob("*.cu"))


def build(force: bool = False, verbose: bool = False) -> Path:
    srcs = sources()
    deps = srcs + sorted(CSRC.glob("*.cuh")) + [ROOT / "include" / "simamba.h"]
    stamp = BUILD / "stamp.txt"
    dig = _digest(deps)
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == dig:
        return LIB
    BUILD.mkdir(exist_ok=True)

    def compile_one(src: Path):
        obj = BUILD / (src.stem + ".o")
        cmd = [NVCC, *ARCH, *CFLAGS, "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, r

    logs = []
    objs = []
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        for src, obj, r in ex.map(compile_one, srcs):
            logs.append(f"==== {src.name}\n{r.stdout}{r.stderr}")
            if r.returncode != 0:
                sys.stderr.write(logs[-1])
                raise RuntimeError(f"nvcc failed on {src.name}")
            objs.append(obj)
    (BUILD / "ptxas.log").write_text("\n".join(logs))
    cmd = [NVCC, *ARCH, "-shared", "-o", str(LIB), *map(str, objs), "-Xcompiler", "-fPIC", "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc link failed")
    stamp.write_text(dig)
    if verbose:
        print("\n".join(logs))
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
