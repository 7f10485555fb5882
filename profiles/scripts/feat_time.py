"""featurizer timing through the library's own per-kernel CUDA events (scratch script, not a bench)"""
import os, sys, json
sys.path.insert(0, os.getcwd())
import numpy as np
import __graft_entry__ as g
pkg = g.load_package()
import copy
cfgname = sys.argv[1] if len(sys.argv) > 1 else "c5"
w = copy.deepcopy(pkg.synthetic.WORKLOADS[cfgname])
N, K = 65536 * 2, 4
xs, ys = pkg.synthetic.make_data(w, N, K)
data = pkg.SimulationData(xs, ys, featurizer=pkg.FeaturesAll())
model = pkg.densenet(list(w.widths), layernorm=True)
iso = pkg.Iso(data, opt=pkg.AdamRegularized(), model=model, minibatch=65536)
for _ in range(2): pkg.koopman(iso)
iso.engine.reset_stats(); iso.engine.enable_timing(True)
for _ in range(3): pkg.koopman(iso)
st = iso.engine.stats()
n = st["n_featurize_launches"]; ms = st["ms_featurize"]
rows = 3 * N * K
F = w.widths[0]; D = xs.shape[0]
ld = (F + 64) // 64 * 64
print(json.dumps({"cfg": cfgname, "env": {k: v for k, v in os.environ.items() if k.startswith("ISOKANN")}, "launches": n, "us_per_launch": 1e3 * ms / n,
                  "ns_per_record": 1e6 * ms / rows, "GBps_split": (4 * D + 4 * ld) * rows / (ms * 1e-3) / 1e9, "GBps_algorithmic": 4 * (D + F) * rows / (ms * 1e-3) / 1e9}))
