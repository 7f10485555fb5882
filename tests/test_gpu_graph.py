"""The optimiser steps of a training epoch replayed as one captured CUDA graph (reference HOT LOOP 2,
src/iso.jl:184-192: floor(N/B) dependent steps) must be bit-identical to the eager launch sequence."""
import copy

import numpy as np
import pytest

from tests.helpers import make_iso, oracle_model

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,widths,gemm,N,B,target", [
    ("c3", None, "auto", 3000, 250, "shiftscale"),          # fused narrow step, 12 steps per epoch
    ("c1", [231, 256, 256, 1], "tc", 1500, 300, "shiftscale"),   # tcgen05 path with the fused thin head
    ("c4", [231, 38, 6, 3], "fp32", 1200, 128, "isa"),      # FP32 CUDA-core GEMMs, N-D target, ragged epoch (9 steps)
    ("c2", None, "auto", 5000, 512, "shiftscale"),          # smallnet, identity featurizer
])
def test_epoch_graph_is_bit_identical_to_eager(pkg, oracle, monkeypatch, name, widths, gemm, N, B, target):
    w = copy.deepcopy(pkg.synthetic.WORKLOADS[name])
    if widths:
        w.widths = list(widths)
    K = 2
    xs, ys = pkg.synthetic.make_data(w, N, K)
    flat = oracle.flatten_params(oracle_model(oracle, w.widths, w.layernorm, 5))
    perms = pkg.synthetic.make_perms(w, N, 5)
    graph = make_iso(pkg, w, xs, ys, flat, opt="adam", target=target, minibatch=B, gemm=gemm)
    monkeypatch.setenv("ISOKANN_GRAPH", "0")
    eager = make_iso(pkg, w, xs, ys, flat, opt="adam", target=target, minibatch=B, gemm=gemm)
    pkg.run_(graph, 5, perms=perms)
    pkg.run_(eager, 5, perms=perms)
    sg, se = graph.engine.stats(), eager.engine.stats()
    assert se["graph_launches"] == 0
    assert sg["graph_launches"] == 4                       # the first epoch allocates and runs eagerly
    assert sg["kernel_launches"] == se["kernel_launches"]  # replayed kernels are counted
    assert np.array_equal(graph.losses, eager.losses)
    assert np.array_equal(graph.engine.download_params(), eager.engine.download_params())
    m_g, v_g, bt_g = graph.engine.download_opt_state()
    m_e, v_e, bt_e = eager.engine.download_opt_state()
    assert np.array_equal(m_g, m_e) and np.array_equal(v_g, v_e) and np.array_equal(bt_g, bt_e)
    # a different minibatch size re-captures; going back to per-epoch calls keeps working
    graph.minibatch = eager.minibatch = B // 2
    for it in range(2):
        pkg.isotarget(graph), pkg.isotarget(eager)
        lg = pkg.train_batch_(graph, perms[it])
        le = pkg.train_batch_(eager, perms[it])
        assert lg == le
    assert np.array_equal(graph.engine.download_params(), eager.engine.download_params())


def test_epoch_graph_survives_data_growth(pkg, oracle):
    """addcoords! between iterations reallocates the resident data: the captured epoch must not be replayed on
    stale pointers"""
    w = pkg.synthetic.WORKLOADS["c1"]
    N, K = 600, 2
    xs, ys = pkg.synthetic.make_data(w, N + 200, K)
    flat = oracle.flatten_params(oracle_model(oracle, w.widths, True, 5))
    iso = make_iso(pkg, w, xs[:, :N], ys[:, :, :N], flat, opt="adam", minibatch=100)
    ref = make_iso(pkg, w, xs, ys, flat, opt="adam", minibatch=100)
    pkg.run_(iso, 3, perms=pkg.synthetic.make_perms(w, N, 3))
    pkg.addcoords_(iso, xs[:, N:], ys[:, :, N:])
    ref.engine.upload_params(iso.engine.download_params())
    m, v, bt = iso.engine.download_opt_state()
    ref.engine.upload_opt_state(m, v, bt)
    perms = pkg.synthetic.make_perms(w, N + 200, 3)
    pkg.run_(iso, 3, perms=perms)
    pkg.run_(ref, 3, perms=perms)
    assert np.array_equal(iso.losses[3:], ref.losses)
    assert np.array_equal(iso.engine.download_params(), ref.engine.download_params())
