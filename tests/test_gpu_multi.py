"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): see tests/multi_gpu_check.py"""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_two_rank_run_matches_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29631", str(ROOT / "tests" / "multi_gpu_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert "MULTI_GPU_CHECK OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_two_devices_in_one_process():
    """a single host process driving two GPUs through two contexts (kernel attributes are per device)"""
    import copy
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, str(ROOT))
    import __graft_entry__ as g
    pkg = g.load_package()
    for name, widths, gemm in [("c3", None, "auto"), ("c1", [231, 256, 256, 1], "tc")]:
        w = copy.deepcopy(pkg.synthetic.WORKLOADS[name])
        if widths:
            w.widths = widths
        N, K = 700, 3
        xs, ys = pkg.synthetic.make_data(w, N, K)
        perms = pkg.synthetic.make_perms(w, N, 2)
        flat0 = pkg.densenet(w.widths, layernorm=True, rng=np.random.default_rng(3)).flat()
        out = []
        for dev in (1, 0):                                  # device 1 first: nothing was set up on it before
            m = pkg.Chain(list(w.widths), True).load_flat(flat0)
            data = pkg.SimulationData(xs, ys, featurizer=pkg.FeaturesAll())
            iso = pkg.Iso(data, opt=pkg.AdamRegularized(), model=m, minibatch=200, device=dev, gemm=gemm)
            pkg.run_(iso, 2, perms=perms)
            out.append((np.array(iso.losses), pkg.chis(iso), iso.engine.download_params()))
        for a, b in zip(out[0], out[1]):
            assert np.array_equal(a, b)
