"""In-tree build of libunitex_b200.so (nvcc, sm_100a only) and of the oracle's C pieces.

`python -m unitex_b200.build` or `__graft_entry__.build()`.  The .so is git-ignored but travels to the GPU box with
the gpurun snapshot; there is deliberately no JIT / torch.utils.cpp_extension cache involved.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
OBJ = ROOT / "_build"
LIB = ROOT / "libunitex_b200.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _sources():
    return sorted(CSRC.glob("*.cu"))


def _digest(src: Path) -> str:
    h = hashlib.sha256()
    h.update(src.read_bytes())
    for hdr in sorted(list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")) + list((ROOT.parent / "include").glob("*.h"))):
        h.update(hdr.read_bytes())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _compile(src: Path) -> Path:
    OBJ.mkdir(exist_ok=True)
    obj = OBJ / (src.stem + ".o")
    stamp = OBJ / (src.stem + ".sha")
    dig = _digest(src)
    if obj.exists() and stamp.exists() and stamp.read_text() == dig:
        return obj
    # bake_* kernels decide triangle ids: no FMA contraction, so the C oracle (-ffp-contract=off) matches bit for bit
    extra = ["-fmad=false"] if src.name.startswith("bake_") else []
    cmd = [NVCC, *FLAGS, *extra, "-c", str(src), "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    (OBJ / (src.stem + ".log")).write_text(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
    stamp.write_text(dig)
    return obj


def build(verbose: bool = False) -> Path:
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(_compile, srcs))
    newest = max(o.stat().st_mtime for o in objs)
    if not LIB.exists() or LIB.stat().st_mtime < newest:
        cmd = [NVCC, "-shared", "-o", str(LIB), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a",
               "-Xcompiler", "-fPIC", "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        for s in srcs:
            log = (OBJ / (s.stem + ".log")).read_text() if (OBJ / (s.stem + ".log")).exists() else ""
            for line in log.splitlines():
                if "registers" in line or "spill" in line or "error" in line:
                    print(f"[{s.name}] {line.strip()}")
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
