"""Wall time of the device-side diagnostics (isokann_rates / _residual_subspace / _residual_ritz) on c4-shaped data,
next to what the reference's host path has to move (chi and Kchi to the host).  Builder measurement, not a bench value.

    python profiles/scripts/diag_time.py [N]
"""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import __graft_entry__ as g  # noqa: E402


def main():
    pkg = g.load_package()
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
    w = pkg.synthetic.WORKLOADS["c4"]
    t0 = time.perf_counter()
    xs, ys = pkg.synthetic.make_data(w, N, 8)
    t_gen = time.perf_counter() - t0
    model = pkg.densenet(w.widths, layernorm=True, rng=np.random.default_rng(1))
    iso = pkg.Iso(pkg.SimulationData(xs, ys, featurizer=pkg.FeaturesAll()), opt=pkg.AdamRegularized(), model=model,
                  target=pkg.TransformISA(), minibatch=65536)
    pkg.run_(iso, 2)

    def timed(f, n=3):
        f()
        iso.engine.synchronize()
        t = time.perf_counter()
        for _ in range(n):
            f()
        iso.engine.synchronize()
        return (time.perf_counter() - t) / n * 1e3

    out = {"N": N, "K": 8, "d": 3, "datagen_s": t_gen,
           "ms_chis_plus_koopman_to_host": timed(lambda: (pkg.chis(iso), pkg.koopman(iso))),
           "ms_rates": timed(lambda: pkg.rates(iso)),
           "ms_residual_subspace_relres_only": timed(lambda: pkg.residual_subspace(iso)),
           "ms_residual_subspace_with_res": timed(lambda: pkg.residual_subspace(iso, want_res=True)),
           "ms_residual_ritz_relres_only": timed(lambda: pkg.residual_ritz(iso)),
           "relres": pkg.residual_subspace(iso)[1].tolist(), "ritz_vals": [complex(v).real for v in pkg.residual_ritz(iso)["vals"]]}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
