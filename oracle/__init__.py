"""CPU oracle for the ISOKANN per-iteration hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is shipped or measured as
the product: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and there only as
the checker or as the CPU baseline.

PARITY UNPINNED: the reference (axsk/ISOKANN.jl, Julia) ships no golden
vectors or known-answer tests for this path (``test/runtests.jl:18,53,62,77``
are ``@test true`` smoke tests) and neither Julia nor the reference's
third-party dependencies (Flux 0.16.9, Optimisers 0.4.7, MLUtils 0.4.8,
PCCAPlus 1.1.2, Combinatorics 1.1.0, LinearAlgebra/OpenBLAS) are present in
this environment, so the reference cannot be executed here.  The oracle is a
restatement of the cited reference lines plus the *published* algorithms of
those packages; its pins are the known-answer tests derived from the cited
code (tests/test_oracle_kat.py) and the committed fixtures under
tests/golden/ that the oracle itself generated (tests/golden/make_golden.py).
"""
from .isokann_oracle import *  # noqa: F401,F403
