"""CPU: the oracle reproduces the committed golden fixtures (guards against oracle drift)."""
from pathlib import Path

import numpy as np
import pytest

GOLD = Path(__file__).resolve().parent / "golden"
CASES = sorted(p.stem for p in GOLD.glob("*.npz") if p.stem != "adp_geometry")


def rec(a):
    return np.ascontiguousarray(np.asarray(a).T)


def test_adp_geometry_fixture(oracle):
    z = np.load(GOLD / "adp_geometry.npz")
    f = oracle.flatpairdists(z["coords_nm"].reshape(1, -1), out_dtype=np.float64)[0]
    assert np.allclose(f, z["pairdists"], rtol=1e-14, atol=0)
    assert np.allclose(oracle.flatpairdists_gram(z["coords_nm"].reshape(1, -1))[0], z["pairdists"], atol=1e-6)


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_golden(oracle, pkg, case):
    z = np.load(GOLD / f"{case}.npz")
    N, K, B, n_iter = (int(v) for v in z["meta"])
    widths = [int(v) for v in z["widths"]]
    ident = widths[0] == z["xs"].shape[0]
    xsf = rec(z["xs"]).astype(np.float32) if ident else oracle.flatpairdists(rec(z["xs"]))
    ysf = rec(z["ys"]).astype(np.float32) if ident else oracle.flatpairdists(rec(z["ys"]))
    assert np.array_equal(xsf[:8], z["features_x"])
    m = oracle.unflatten_params(oracle.Model(widths, not ident), z["flat0"])
    assert np.allclose(oracle.forward(m, xsf), z["chi0"], rtol=1e-5, atol=1e-6)
    assert np.allclose(oracle.expectation(m, ysf), z["kchi0"], rtol=1e-5, atol=1e-6)
    tk = str(z["target"])
    topts = {str(k): False for k in z["topts"]}
    assert np.allclose(oracle.isotarget(tk, m, xsf, ysf, **topts), z["target0"], rtol=1e-4, atol=1e-4)
    cfg = oracle.OptConfig(kind=str(z["opt"]))
    st = oracle.opt_init(cfg, z["flat0"].size)
    losses = oracle.run(m, xsf, ysf, cfg, st, n_iter, B, list(z["perms"]), tk, **topts)
    assert np.allclose(losses, z["losses"], rtol=1e-4)
    assert np.allclose(oracle.flatten_params(m), z["flat_final"], rtol=1e-4, atol=1e-6)
