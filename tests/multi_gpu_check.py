"""torchrun worker: N-rank run of the library (start points sharded, NCCL all-gather of K-chi and
all-reduce of gradients) must reproduce the single-GPU run of the same workload.

    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_check.py
"""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def incremental_data_check(pkg, world, rank, local):
    """addcoords! and the cutoff window of run_kde! (src/iso.jl:238,288-290) on several ranks: isokann_append_data /
    isokann_keep_last rebuild every rank's shard of ys on the device; afterwards the context must behave bit for bit
    like a fresh multi-rank upload of the same data (same shards -> same kernels, same summation order)."""
    failures = []
    w = pkg.synthetic.WORKLOADS["c1"]
    N0, n_new, K, keep = 501, 203, 3, 397             # odd sizes: every shard boundary moves
    xs, ys = pkg.synthetic.make_data(w, N0 + n_new, K)
    flat0 = pkg.densenet(w.widths, layernorm=True, rng=np.random.default_rng(11)).flat()

    def make(lo, hi):
        m = pkg.Chain(list(w.widths), True).load_flat(flat0)
        data = pkg.SimulationData(xs[:, lo:hi], ys[:, :, lo:hi], featurizer=pkg.FeaturesAll())
        return pkg.Iso(data, opt=pkg.AdamRegularized(), model=m, minibatch=100, device=local,
                       comm=(world, rank, pkg.parallel.broadcast_unique_id(rank)))

    inc = make(0, N0)
    k0 = pkg.koopman(inc)                               # touch the data before it grows
    pkg.addcoords_(inc, xs[:, N0:], ys[:, :, N0:])
    fresh = make(0, N0 + n_new)
    if not (len(inc.data) == N0 + n_new and np.array_equal(pkg.koopman(inc), pkg.koopman(fresh))
            and np.array_equal(pkg.koopman(inc)[:, :N0], k0)):
        failures.append(("append", "Koopman vector differs from a fresh upload"))
    perms = pkg.synthetic.make_perms(w, N0 + n_new, 2)
    pkg.run_(inc, 2, perms=perms)
    pkg.run_(fresh, 2, perms=perms)
    if not (np.array_equal(inc.losses, fresh.losses)
            and np.array_equal(inc.engine.download_params(), fresh.engine.download_params())):
        failures.append(("append", "training after the append differs", inc.losses, fresh.losses))
    pkg.cutoff_(inc, keep)
    last = make(N0 + n_new - keep, N0 + n_new)
    last.engine.upload_params(inc.engine.download_params())
    last.engine.upload_opt_state(*inc.engine.download_opt_state())
    if not (len(inc.data) == keep and np.array_equal(pkg.koopman(inc), pkg.koopman(last))
            and np.array_equal(pkg.chis(inc), pkg.chis(last))):
        failures.append(("keep_last", "chi / Koopman vector differ from a fresh upload of the window"))
    perms = pkg.synthetic.make_perms(w, keep, 2)
    pkg.run_(inc, 2, perms=perms)
    pkg.run_(last, 2, perms=perms)
    if not (np.array_equal(inc.losses[-2:], last.losses[-2:])
            and np.array_equal(inc.engine.download_params(), last.engine.download_params())):
        failures.append(("keep_last", "training after the cutoff differs", inc.losses[-2:], last.losses[-2:]))
    for iso_ in (inc, fresh, last):
        iso_.engine.close()
    return failures


def main():
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    pkg = g.load_package()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import copy
    failures = []
    # ISOKANN_MULTI_CASES="0,6" restricts the run to some of the cases below, "none" to the incremental-data part only
    sel = os.environ.get("ISOKANN_MULTI_CASES", "")
    cases = [("c1", None, "shiftscale", 1003, 3, 250, "auto"),
                                                ("c4", [231, 38, 6, 3], "pinv", 777, 2, 128, "auto"),
                                                ("c4", [231, 38, 6, 2], "isa", 640, 2, 0, "auto"),
                                                ("c1", [231, 256, 256, 1], "shiftscale", 900, 2, 300, "tc"),
                                                ("c4", [231, 256, 264, 3], "pinv_default", 1200, 2, 400, "tc"),
                                                ("c1", [231, 512, 1], "shiftscale", 700, 2, 0, "tc"),
                                                # triple well: identity featurizer, no LayerNorm, F = 2 (the fused
                                                # narrow step once read uninitialised shared memory here)
                                                ("c2", None, "shiftscale", 1000, 3, 250, "auto")]
    if sel:
        cases = [] if sel == "none" else [cases[int(i)] for i in sel.split(",")]
    for name, widths, target, N, K, B, gemm in cases:
        w = copy.deepcopy(pkg.synthetic.WORKLOADS[name])
        if widths:
            w.widths = widths
        xs, ys = pkg.synthetic.make_data(w, N, K)
        perms = pkg.synthetic.make_perms(w, N, 3)
        model = pkg.densenet(w.widths, layernorm=w.layernorm, rng=np.random.default_rng(7))
        flat0 = model.flat()
        tobj = {"shiftscale": pkg.TransformShiftscale, "isa": pkg.TransformISA, "pinv": pkg.TransformPseudoInv,
                "pinv_default": pkg.TransformPseudoInv}[target]()
        if target == "pinv":
            tobj = pkg.TransformPseudoInv(eigenvecs=False)   # Schur vectors are rounding-sensitive

        def make(comm):
            m = pkg.Chain(list(w.widths), w.layernorm).load_flat(flat0)
            feat = pkg.FeaturesAll() if w.featurizer == "allpairs" else pkg.FeaturesCoords()
            data = pkg.SimulationData(xs, ys, featurizer=feat)
            return pkg.Iso(data, opt=pkg.AdamRegularized(), model=m, target=tobj, minibatch=B, device=local, gemm=gemm,
                           comm=comm)
        uid = pkg.parallel.broadcast_unique_id(rank)      # a communicator needs its own fresh id
        multi = make((world, rank, uid))
        pkg.run_(multi, 3, perms=perms)
        chi_m = pkg.chis(multi)
        flat_m = multi.engine.download_params()
        single = make(None)
        pkg.run_(single, 3, perms=perms)
        chi_s = pkg.chis(single)
        flat_s = single.engine.download_params()
        tol = 2e-3 if gemm == "tc" or target != "shiftscale" else 2e-4
        ok = (np.allclose(multi.losses, single.losses, rtol=tol) and np.allclose(chi_m, chi_s, rtol=tol, atol=tol)
              and np.abs(flat_m - flat_s).max() < tol * np.abs(flat_s).max())
        st = multi.engine.stats()
        if not ok or st["nccl_calls"] == 0:
            failures.append((name, target, multi.losses, single.losses, float(np.abs(chi_m - chi_s).max())))
        # the gradient exchange over NVLink peer memory (csrc/p2p.cu) against ncclAllReduce: with two ranks both add
        # the same two numbers, so the runs must agree bit for bit
        if st["p2p_exchanges"] == 0:
            failures.append((name, "peer-memory exchange not in use", st))
        os.environ["ISOKANN_NO_P2P"] = "1"
        viaN = make((world, rank, pkg.parallel.broadcast_unique_id(rank)))
        del os.environ["ISOKANN_NO_P2P"]
        pkg.run_(viaN, 3, perms=perms)
        if viaN.engine.stats()["p2p_exchanges"] != 0 or not (
                np.array_equal(viaN.losses, multi.losses) and np.array_equal(viaN.engine.download_params(), flat_m)):
            failures.append((name, "NCCL and peer-memory exchange differ", viaN.losses, multi.losses))
        viaN.engine.close()
        # asynchronous upload: every rank sends only its own rows of xs over PCIe and fetches the rest from its peers
        # (in place for equal shards, padded otherwise); must reproduce the blocking upload bit for bit
        uid2 = pkg.parallel.broadcast_unique_id(rank)
        again = make((world, rank, uid2))
        off, n = pkg.parallel.shard_range(N, world, rank)
        again.engine.set_data_async(xs, np.asfortranarray(ys[:, :, off:off + n]), n_offset=off, n_local=n)
        pkg.run_(again, 3, perms=perms)
        if not (np.array_equal(again.losses, multi.losses)
                and np.array_equal(again.engine.download_params(), flat_m) and np.array_equal(pkg.chis(again), chi_m)):
            failures.append((name, "async upload differs", again.losses, multi.losses))
        if gemm == "tc":
            # the bucketed exchange beside the backward pass against the single-stream step with one all-reduce;
            # plus a ragged last minibatch (partial=true) that leaves the last rank without rows
            os.environ["ISOKANN_NO_COMM_OVERLAP"] = "1"
            plain = make((world, rank, pkg.parallel.broadcast_unique_id(rank)))
            del os.environ["ISOKANN_NO_COMM_OVERLAP"]
            pkg.run_(plain, 3, perms=perms)
            if not (np.allclose(plain.losses, multi.losses, rtol=1e-5)
                    and np.abs(plain.engine.download_params() - flat_m).max() < 1e-4 * np.abs(flat_m).max()):
                failures.append((name, "overlapped step differs from the single-stream step", plain.losses, multi.losses))
            Br = N - 1 if B == 0 else (N - 1) // 2
            for iso_ in (multi, single):
                iso_.minibatch = Br
                pkg.isotarget(iso_)
            l_m = pkg.train_batch_(multi, perms[0], partial=True)
            l_s = pkg.train_batch_(single, perms[0], partial=True)
            if not (np.isclose(l_m, l_s, rtol=tol)
                    and np.abs(multi.engine.download_params() - single.engine.download_params()).max()
                    < tol * np.abs(flat_s).max()):
                failures.append((name, "ragged last minibatch", l_m, l_s))
            flat_m = multi.engine.download_params()
        # every rank must hold identical parameters
        t = torch.from_numpy(flat_m.copy()).cuda()
        ref = t.clone()
        dist.broadcast(ref, 0)
        if not torch.equal(t, ref):
            failures.append((name, "replicas diverged"))
    failures += incremental_data_check(pkg, world, rank, local)
    flag = torch.tensor([len(failures)], device="cuda")
    dist.all_reduce(flag)
    if rank == 0:
        print("MULTI_GPU_CHECK", "OK" if flag.item() == 0 else f"FAILED {failures}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 0 else 1)


if __name__ == "__main__":
    main()
