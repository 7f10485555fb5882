"""Build libisokann_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libisokann_b200.so"
STAMP = HERE / "build" / "stamp.txt"

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def sources():
    return sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cpp")))


def _digest() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*")) + [HERE.parent / "include" / "isokann_b200.h", Path(__file__)]):
        if p.is_file():
            h.update(p.name.encode())
            h.update(p.read_bytes())
    return h.hexdigest()


def nccl_include_dir() -> Path:
    """nccl.h: $NCCL_INCLUDE_DIR, $NCCL_HOME/include, the nvidia-nccl wheel next to the running interpreter, or the
    system include directory (the library itself is dlopen'ed at run time, csrc/nccl_dyn.cpp)"""
    cands = []
    if os.environ.get("NCCL_INCLUDE_DIR"):
        cands.append(Path(os.environ["NCCL_INCLUDE_DIR"]))
    if os.environ.get("NCCL_HOME"):
        cands.append(Path(os.environ["NCCL_HOME"]) / "include")
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        if spec and spec.submodule_search_locations:
            cands += [Path(p) / "include" for p in spec.submodule_search_locations]
    except Exception:
        pass
    cands += [Path("/usr/include"), Path("/usr/local/cuda/include")]
    for c in cands:
        if (c / "nccl.h").exists():
            return c
    raise RuntimeError("nccl.h not found: set NCCL_INCLUDE_DIR or NCCL_HOME")


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source into one shared library.  Objects are cached per source."""
    dig = _digest()
    if not force and LIB.exists() and STAMP.exists() and STAMP.read_text().strip() == dig:
        return LIB
    nvcc = _nvcc()
    objdir = HERE / "build"
    objdir.mkdir(exist_ok=True)
    incs = ["-I", str(HERE.parent / "include"), "-I", str(CSRC)]
    incs += ["-I", str(nccl_include_dir())]
    objs = []
    procs = []
    for src in sources():
        obj = objdir / (src.stem + ".o")
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, *incs, "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"--- nvcc failed on {src.name} ---\n{out}\n")
        elif verbose and out:
            sys.stderr.write(f"--- {src.name} ---\n{out}\n")
    if failed:
        raise RuntimeError("nvcc compilation failed")
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB), *map(str, objs),
            "-ldl"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link failed")
    STAMP.write_text(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
