"""GPU: the CUDA library (through the C ABI) against the committed golden fixtures."""
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"
CASES = sorted(p.stem for p in GOLD.glob("*.npz") if p.stem not in ("adp_geometry", "diagnostics"))


def rec(a):
    return np.ascontiguousarray(np.asarray(a).T)


def test_adp_geometry(pkg):
    z = np.load(GOLD / "adp_geometry.npz")
    f = pkg.flatpairdists(np.asfortranarray(z["coords_nm"].reshape(-1, 1).astype(np.float32)))
    assert np.allclose(f[:, 0], z["pairdists"], rtol=1e-5)     # 1e-5 relative on distances


@pytest.mark.parametrize("case", CASES)
def test_library_matches_golden(pkg, case):
    z = np.load(GOLD / f"{case}.npz")
    N, K, B, n_iter = (int(v) for v in z["meta"])
    widths = [int(v) for v in z["widths"]]
    ident = widths[0] == z["xs"].shape[0]
    feat = pkg.FeaturesCoords() if ident else pkg.FeaturesAll()
    data = pkg.SimulationData(z["xs"], z["ys"], featurizer=feat)
    if not ident:
        assert np.allclose(rec(data.features())[:8], z["features_x"], rtol=1e-5)
    model = pkg.Chain(widths, not ident).load_flat(z["flat0"])
    opt = pkg.AdamRegularized() if str(z["opt"]) == "adam" else pkg.NesterovRegularized()
    tk = str(z["target"])
    topts = {str(k): False for k in z["topts"]}
    tobj = {"shiftscale": pkg.TransformShiftscale, "isa": pkg.TransformISA, "pinv": pkg.TransformPseudoInv}[tk](**topts)
    iso = pkg.Iso(data, opt=opt, model=model, target=tobj, minibatch=B)
    assert np.allclose(rec(pkg.chis(iso)), z["chi0"], rtol=1e-4, atol=5e-5)       # 1e-4 on chi
    assert np.allclose(rec(pkg.koopman(iso)), z["kchi0"], rtol=1e-4, atol=5e-5)
    tol_t = 2e-4 if tk == "shiftscale" else 2e-3 * np.abs(z["target0"]).max()
    assert np.allclose(rec(pkg.isotarget(iso)), z["target0"], atol=tol_t)
    pkg.run_(iso, n_iter, perms=z["perms"])
    assert np.allclose(iso.losses, z["losses"], rtol=1e-3 if tk == "shiftscale" else 1e-2)
    assert np.allclose(rec(pkg.chis(iso)), z["chi_final"], rtol=1e-3, atol=1e-3)
