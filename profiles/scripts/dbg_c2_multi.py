"""debug: c2 (triple well, smallnet) on 2 ranks vs 1 rank, iteration by iteration"""
import os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import torch, torch.distributed as dist
import __graft_entry__ as g
pkg = g.load_package()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
w = pkg.synthetic.WORKLOADS["c2"]
N, K, B = int(os.environ.get("DBG_N", 100000)), 8, int(os.environ.get("DBG_B", 4096))
xs, ys = pkg.synthetic.make_data(w, N, K)
perms = pkg.synthetic.make_perms(w, N, 4)
flat0 = pkg.densenet(w.widths, layernorm=False, rng=np.random.default_rng(int(os.environ.get("DBG_SEED", 2)))).flat()
print("data", float(np.abs(xs).max()), float(np.abs(ys).max()), np.isnan(ys).any(), "flat", float(np.abs(flat0).max()), flush=True)
def make(comm):
    m = pkg.Chain(list(w.widths), False).load_flat(flat0)
    data = pkg.SimulationData(xs, ys, featurizer=pkg.FeaturesCoords())
    return pkg.Iso(data, opt=pkg.NesterovRegularized(), model=m, minibatch=B, device=local, comm=comm)
mode = os.environ.get("DBG_MODE", "both")
isos = []
if world > 1 and mode in ("both", "multi"):
    isos.append(("multi", make((world, rank, pkg.parallel.broadcast_unique_id(rank)))))
if mode in ("both", "single"):
    isos.append(("single", make(None)))
for it in range(3):
    for name, iso in isos:
        try:
            c = pkg.chis(iso)
            k = pkg.koopman(iso)
            if rank == 0:
                print(name, it, "chi", float(c.min()), float(c.max()), "kchi", float(k.min()), float(k.max()), "nan" if np.isnan(k).any() else "", flush=True)
            t = pkg.isotarget(iso)
            if rank == 0:
                print(name, it, "target", float(t.min()), float(t.max()), flush=True)
            l = pkg.train_batch_(iso, perms[it])
            p = iso.engine.download_params()
            if rank == 0:
                print(name, it, "loss", l, "|p|", float(np.linalg.norm(p)), flush=True)
        except Exception as e:
            if rank == 0:
                print(name, it, "EXC", repr(e)[:200], flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
