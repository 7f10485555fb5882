"""Compile the oracle's C restatement (oracle/isokann_oracle.c) into oracle/liboracle.so with gcc.

TEST INFRASTRUCTURE ONLY.  The reference is Julia: there is nothing to compile into oracle/_ref."""
from __future__ import annotations

import shutil
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
SRC = HERE / "isokann_oracle.c"
LIB = HERE / "liboracle.so"


def build(force: bool = False) -> Path:
    if not force and LIB.exists() and LIB.stat().st_mtime >= SRC.stat().st_mtime:
        return LIB
    gcc = shutil.which("gcc") or "gcc"
    cmd = [gcc, "-O2", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", str(SRC), "-o", str(LIB), "-lm"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("gcc failed on the oracle:\n" + r.stdout)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
