"""Generates tests/golden/*.npz from the ORACLE (oracle/isokann_oracle.py).

The reference (Julia) cannot run in this environment and ships no golden vectors for this path
(test/runtests.jl only asserts `@test true`), so these fixtures are NOT reference outputs: they
freeze the oracle's numbers on small seeded cases so that (a) the oracle cannot drift silently
between rounds (tests/test_golden.py, CPU) and (b) the CUDA library is compared against committed
numbers, not only against a live oracle run (tests/test_gpu_golden.py).

    python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))

import __graft_entry__ as g  # noqa: E402
import oracle  # noqa: E402

OUT = Path(__file__).resolve().parent


def case(pkg, name, widths, N, K, B, n_iter, opt, target, seed_model, topts=None):
    topts = dict(topts or {})
    import copy
    w = copy.deepcopy(pkg.synthetic.WORKLOADS[name])
    w.widths = list(widths)
    xs, ys = pkg.synthetic.make_data(w, N, K)
    rec = lambda a: np.ascontiguousarray(np.asarray(a).T)
    if w.featurizer == "identity":
        xsf, ysf = rec(xs).astype(np.float32), rec(ys).astype(np.float32)
    else:
        xsf, ysf = oracle.flatpairdists(rec(xs)), oracle.flatpairdists(rec(ys))
    m = oracle.init_params(oracle.Model(list(w.widths), w.layernorm), np.random.default_rng(seed_model))
    flat0 = oracle.flatten_params(m)
    perms = pkg.synthetic.make_perms(w, N, n_iter)
    chi0 = oracle.forward(m, xsf)
    kchi0 = oracle.expectation(m, ysf)
    target0 = oracle.isotarget(target, m, xsf, ysf, **topts)
    cfg = oracle.OptConfig(kind=opt)
    st = oracle.opt_init(cfg, flat0.size)
    losses = oracle.run(m, xsf, ysf, cfg, st, n_iter, B, list(perms), target, **topts)
    return dict(xs=np.asarray(xs), ys=np.asarray(ys), features_x=xsf[:8], flat0=flat0, perms=perms, chi0=chi0,
                kchi0=kchi0, target0=target0, losses=np.array(losses), flat_final=oracle.flatten_params(m),
                chi_final=oracle.forward(m, xsf), widths=np.array(w.widths), meta=np.array([N, K, B, n_iter]),
                name=name, opt=opt, target=target, topts=np.array(sorted(k for k, v in topts.items() if v is False)))


def diagnostics_fixture():
    """chi / Kchi of the final models of two committed cases and the oracle's rates / residual_subspace /
    residual_ritz on them.  Written to diagnostics.npz only; the other fixtures are not touched."""
    out = {}
    for key in ("adp_shiftscale_adam", "adp_isa_2d"):
        z = np.load(OUT / f"{key}.npz")
        widths = [int(v) for v in z["widths"]]
        rec = lambda a: np.ascontiguousarray(np.asarray(a).T)
        xsf, ysf = oracle.flatpairdists(rec(z["xs"])), oracle.flatpairdists(rec(z["ys"]))
        m = oracle.unflatten_params(oracle.Model(widths, True), z["flat_final"])
        chi, kchi = oracle.forward(m, xsf), oracle.expectation(m, ysf)
        res, relres = oracle.residual_subspace(chi, kchi)
        _, ritz_relres, vals, _, _ = oracle.residual_ritz(chi, kchi)
        out.update({f"{key}__chi": chi, f"{key}__kchi": kchi, f"{key}__rates": np.real(oracle.rates(chi, kchi)),
                    f"{key}__res": res, f"{key}__relres": relres, f"{key}__ritz_relres": ritz_relres,
                    f"{key}__ritz_vals": np.asarray(vals, dtype=np.complex128)})
    np.savez_compressed(OUT / "diagnostics.npz", **out)
    print("wrote diagnostics")


def main():
    if "--diagnostics-only" in sys.argv:
        return diagnostics_fixture()
    pkg = g.load_package()
    cases = {
        "adp_shiftscale_nesterov": ("c1", [231, 38, 6, 1], 48, 3, 16, 3, "nesterov", "shiftscale", 11),
        "adp_shiftscale_adam": ("c1", [231, 38, 6, 1], 48, 3, 16, 3, "adam", "shiftscale", 11),
        "triplewell_smallnet": ("c2", [2, 8, 8, 8, 1], 64, 4, 32, 3, "nesterov", "shiftscale", 12),
        "villin_shiftscale": ("c3", [595, 71, 8, 1], 40, 2, 20, 2, "nesterov", "shiftscale", 13),
        # N-D targets without the rounding-sensitive choices (Schur vector signs, fixperm ties at random init);
        # those are covered by tests/test_gpu_parity.py::test_nd_targets on seeds with a clear margin
        "adp_pinv_3d": ("c4", [231, 38, 6, 3], 60, 3, 20, 2, "adam", "pinv", 14, {"eigenvecs": False, "permute": False}),
        "adp_isa_2d": ("c4", [231, 38, 6, 2], 60, 3, 20, 2, "adam", "isa", 15, {"permute": False}),
    }
    for key, args in cases.items():
        np.savez_compressed(OUT / f"{key}.npz", **case(pkg, *args))
        print("wrote", key)
    # ADP reference geometry: analytic pair distances in float64 (KAT 2 of SURVEY section 8c)
    x = pkg.synthetic.ADP_NM
    d = np.array([np.linalg.norm(x[i] - x[j]) for j in range(1, 22) for i in range(j)])
    np.savez_compressed(OUT / "adp_geometry.npz", coords_nm=x, pairdists=d)
    diagnostics_fixture()


if __name__ == "__main__":
    main()
