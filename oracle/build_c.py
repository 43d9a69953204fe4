"""Builds the oracle's C restatement (oracle/bake_ref.c -> oracle/_build/libbake_oracle.so) with gcc.
TEST INFRASTRUCTURE: building the checker is not using it.  -ffp-contract=off keeps every fp32 op separately rounded."""
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
OUT = HERE / "_build" / "libbake_oracle.so"


def build() -> Path:
    src = HERE / "bake_ref.c"
    OUT.parent.mkdir(exist_ok=True)
    if OUT.exists() and OUT.stat().st_mtime >= src.stat().st_mtime:
        return OUT
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", str(OUT), str(src), "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(r.stderr)
    return OUT


if __name__ == "__main__":
    print(build())
