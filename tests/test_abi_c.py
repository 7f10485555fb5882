"""The drop-in boundary from plain C: tests/abi_c/roundtrip.c includes include/isokann_b200.h (C99, gcc) and runs
create -> upload -> set_data -> iterate -> download against the shared library, with no Python in the data path."""
import os
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
LIBDIR = ROOT / "isokann.jl_b200"


def build_c(tmp_path, pkg):
    exe = tmp_path / "roundtrip"
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", str(ROOT / "include"),
           str(ROOT / "tests" / "abi_c" / "roundtrip.c"), "-o", str(exe), "-L", str(LIBDIR), "-lisokann_b200", "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_header_compiles_as_c99_and_links(tmp_path, pkg):
    exe = build_c(tmp_path, pkg)
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the run is checked by the gpu test")
    r = subprocess.run([str(exe)], capture_output=True, text=True, env={**os.environ, "LD_LIBRARY_PATH": str(LIBDIR)})
    assert r.returncode == 1 and "no CUDA device available" in r.stderr       # fails loudly, no fallback


@pytest.mark.gpu
def test_c_program_runs_the_hot_path(tmp_path, pkg):
    exe = build_c(tmp_path, pkg)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300,
                       env={**os.environ, "LD_LIBRARY_PATH": str(LIBDIR)})
    assert r.returncode == 0 and "ABI_C OK" in r.stdout, r.stdout + r.stderr
