"""CPU: the oracle reproduces the committed golden fixtures (guards against oracle drift)."""
from pathlib import Path

import numpy as np
import pytest

GOLD = Path(__file__).resolve().parent / "golden"
CASES = sorted(p.stem for p in GOLD.glob("*.npz") if p.stem not in ("adp_geometry", "diagnostics"))


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


@pytest.mark.parametrize("key", ["adp_shiftscale_adam", "adp_isa_2d"])
def test_diagnostics_golden(oracle, pkg, key):
    """rates / residual_subspace / residual_ritz: the oracle reproduces the committed numbers, and so
    does the library's host-side algebra when it is fed the second moments the device reduction would deliver"""
    z = np.load(GOLD / "diagnostics.npz")
    chi, kchi = z[f"{key}__chi"], z[f"{key}__kchi"]
    d = chi.shape[1]
    assert np.allclose(np.real(oracle.rates(chi, kchi)), z[f"{key}__rates"], atol=1e-9)
    res, relres = oracle.residual_subspace(chi, kchi)
    assert np.allclose(res, z[f"{key}__res"], atol=1e-12) and np.allclose(relres, z[f"{key}__relres"], rtol=1e-9)
    _, rr, vals, _, _ = oracle.residual_ritz(chi, kchi)
    assert np.allclose(rr, z[f"{key}__ritz_relres"], rtol=1e-8) and np.allclose(vals, z[f"{key}__ritz_vals"], atol=1e-10)
    # the library's d x d algebra on numpy-computed moments
    lib, ptr = pkg.lib.load(), pkg.lib.ptr
    n = chi.shape[0]
    u = np.concatenate([chi.astype(np.float64), np.ones((n, 1))], axis=1)
    v = np.concatenate([kchi.astype(np.float64), np.ones((n, 1))], axis=1)
    uu, vu = np.ascontiguousarray(u.T @ u), np.ascontiguousarray(v.T @ u)
    out = np.zeros(82)
    assert lib.isokann_host_diag(0, ptr(uu), ptr(vu), d, ptr(out)) == 0
    m = int(out[0])
    assert np.allclose(out[1:1 + m * m].reshape(m, m), z[f"{key}__rates"], atol=1e-8)
    out = np.zeros(2 * d + 6 * d * d + 1)
    assert lib.isokann_host_diag(2, ptr(uu), ptr(vu), d, ptr(out)) == 0
    assert np.allclose(out[:2 * d].view(np.complex128), z[f"{key}__ritz_vals"], atol=1e-9)
