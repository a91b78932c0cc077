"""In-tree build of libb200q4.so (CUDA kernels + C ABI + the imax hook) for sm_100a.

`python -m vsim_b200.build` or `vsim_b200.build.build_library()`.  nvcc
cross-compiles without a GPU; the resulting .so is git-ignored but travels
to the GPU box with the repository snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libb200q4.so"
SOURCES = [CSRC / "q4_api.cu", CSRC / "q4_prog_api.cu", CSRC / "neox_graph.cu", CSRC / "imax_b200.c"]
DEPS = SOURCES + [CSRC / "q4_kernels.cuh", CSRC / "q4_tile.cuh", CSRC / "q4_prog.cuh", CSRC / "q4_prog_host.h", CSRC / "q4_internal.h", PKG.parent / "include" / "b200q4.h",
                  PKG.parent / "include" / "b200q4_ggml_abi.h", PKG.parent / "include" / "b200neox.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-t", "0",
              "-cudart", "shared",            # the runtime stays a shared library: no copy of its symbol table inside the artefact
              "-Xcompiler", "-fPIC", "-shared"]


def nvcc_path() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(cand).exists():
        raise RuntimeError("nvcc not found: the B200 back-end cannot be built (no CPU fallback exists)")
    return cand


def is_stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any(d.stat().st_mtime > t for d in DEPS)


def build_library(force: bool = False, verbose: bool = False) -> Path:
    if not force and not is_stale():
        return LIB
    out = Path(os.environ["B200Q4_LIB_OUT"]) if os.environ.get("B200Q4_LIB_OUT") else LIB     # development: a second build beside the product
    cmd = [nvcc_path(), *NVCC_FLAGS, "-o", str(out), *map(str, SOURCES)]
    if os.environ.get("B200Q4_TIMELINE"):          # development: per-CTA %globaltimer stamps (tools/timeline.py)
        cmd.insert(1, "-DB200Q4_TIMELINE")
    if os.environ.get("B200Q4_DEFS"):              # development: extra -D switches
        cmd[1:1] = ["-D" + d for d in os.environ["B200Q4_DEFS"].split(",")]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libb200q4.so")
    if verbose:
        sys.stderr.write(r.stderr)
    return out


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
