"""World-size-2 (gloo, CPU) test of the host-side sharding plan used for N > 1 GPUs: start points
are split contiguously for the Koopman pass, every minibatch is split contiguously for the
training step, gradients (scaled by the GLOBAL batch size) are all-reduced.  The oracle stands in
for the device compute; what is under test is isokann.jl_b200/parallel.py, the plan the library
follows (csrc/api.cu: split_range, train_step)."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = g.load_package()
    par = pkg.parallel
    w = pkg.synthetic.WORKLOADS["c1"]
    N, K, B = 101, 3, 40
    xs, ys = pkg.synthetic.make_data(w, N, K)
    xsf = oracle.flatpairdists(np.ascontiguousarray(xs.T))
    ysf = oracle.flatpairdists(np.ascontiguousarray(ys.T))
    m = oracle.init_params(oracle.Model(list(w.widths), True), np.random.default_rng(1))
    # --- sharded Koopman expectation + all-gather (padded shards, like allgather_rows) ---
    off, n = par.shard_range(N, world, rank)
    local = oracle.expectation(m, ysf[off:off + n])
    nmax = -(-N // world)
    pad = np.zeros((nmax, 1), np.float32)
    pad[:n] = local
    bufs = [torch.zeros(nmax, 1) for _ in range(world)]
    dist.all_gather(bufs, torch.from_numpy(pad))
    full = np.concatenate([bufs[r].numpy()[:par.shard_range(N, world, r)[1]] for r in range(world)])
    target = oracle.shiftscale(full)
    # --- one epoch: every rank takes its slice of every minibatch, gradients are summed ---
    perm = pkg.synthetic.make_perms(w, N, 1)[0]
    cfg = oracle.OptConfig(kind="adam")
    st = oracle.opt_init(cfg, oracle.num_params(m))
    ls = 0.0
    for start, length in par.batch_bounds(N, B):
        idx = par.rank_batch_slice(perm, start, length, world, rank) - 1
        l, grad = oracle.batch_loss_and_grad(m, xsf[idx], target[idx], None)
        grad = grad * (len(idx) / length)                 # gradient of l/B_global, not l/B_local
        t = torch.from_numpy(np.concatenate([grad.astype(np.float64), [l]]))
        dist.all_reduce(t)
        ls += float(t[-1])
        oracle.unflatten_params(m, oracle.opt_update(cfg, st, oracle.flatten_params(m), t[:-1].numpy()))
    np.savez(Path(out_dir) / f"rank{rank}.npz", flat=oracle.flatten_params(m), loss=ls / N, target=target)
    dist.destroy_process_group()


def test_world2_matches_single_process(tmp_path, pkg, oracle):
    import torch.multiprocessing as mp
    import socket
    with socket.socket() as sk:                         # a port that is free right now
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    assert np.array_equal(r0["flat"], r1["flat"])          # replicas stay bit-identical
    # single-process reference
    w = pkg.synthetic.WORKLOADS["c1"]
    N, K, B = 101, 3, 40
    xs, ys = pkg.synthetic.make_data(w, N, K)
    xsf = oracle.flatpairdists(np.ascontiguousarray(xs.T))
    ysf = oracle.flatpairdists(np.ascontiguousarray(ys.T))
    m = oracle.init_params(oracle.Model(list(w.widths), True), np.random.default_rng(1))
    target = oracle.isotarget_shiftscale(m, xsf, ysf)
    assert np.allclose(target, r0["target"], atol=2e-6)    # BLAS rounding depends on the shard shape
    cfg = oracle.OptConfig(kind="adam")
    perm = pkg.synthetic.make_perms(w, N, 1)[0]
    loss = oracle.train_batch(m, xsf, target, cfg, oracle.opt_init(cfg, oracle.num_params(m)), B, perm)
    assert np.isclose(loss, float(r0["loss"]), rtol=1e-6)
    assert np.abs(oracle.flatten_params(m) - r0["flat"]).max() < 2e-5


def _reshard_worker(rank, world, port, out_dir):
    """the data movement of isokann_append_data / isokann_keep_last with several ranks (reshard_ys, csrc/api.cu):
    all-gather the padded old shards, cut the new range out of the gathered copy, take the rest from the appended
    block -- with gloo tensors standing in for device buffers"""
    sys.path.insert(0, str(ROOT))
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    par = g.load_package().parallel
    rng = np.random.default_rng(0)
    data = rng.standard_normal((1000, 6)).astype(np.float32)     # global rows; every rank can index the "host" copy
    ok = True
    for n_old, appended, keep in [(501, 203, 397), (7, 1, 3), (64, 64, 1), (5, 0, 5), (3, 8, 2)]:
        off, n = par.shard_range(n_old, world, rank)
        local = data[off:off + n]
        for shift, n_new, block in [(0, n_old + appended, data[n_old:n_old + appended]), (n_old - keep, keep, None)]:
            nmax = -(-n_old // world)
            pad = np.zeros((nmax, 6), np.float32)
            pad[:n] = local
            bufs = [torch.zeros(nmax, 6) for _ in range(world)]
            dist.all_gather(bufs, torch.from_numpy(pad))
            full = np.concatenate([bufs[r].numpy()[:par.shard_range(n_old, world, r)[1]] for r in range(world)])
            (o1, l1), (a, b), (a2, b2) = par.reshard_plan(n_old, shift, n_new, world, rank)
            new = np.full((l1, 6), np.nan, np.float32)
            new[:b - a] = full[a:b]
            if b2 > a2:
                new[a2 - shift - o1:b2 - shift - o1] = block[a2 - n_old:b2 - n_old]
            expect = (data[:n_new] if shift == 0 else data[shift:shift + n_new])[o1:o1 + l1]
            ok = ok and np.array_equal(new, expect)
    Path(out_dir, f"reshard{rank}.txt").write_text("ok" if ok else "mismatch")
    dist.destroy_process_group()


def test_world2_reshard_plan(tmp_path, pkg):
    import torch.multiprocessing as mp
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mp.spawn(_reshard_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "reshard0.txt").read_text() == "ok" and (tmp_path / "reshard1.txt").read_text() == "ok"
    # the same rule for other world sizes, without processes
    par = pkg.parallel
    for world in (3, 8):
        for n_old, shift, n_new in [(501, 0, 704), (704, 307, 397), (5, 0, 6), (9, 8, 1)]:
            rows = []
            for r in range(world):
                (o1, l1), (a, b), (a2, b2) = par.reshard_plan(n_old, shift, n_new, world, r)
                got = list(range(a, b)) + list(range(a2, b2))
                assert got == list(range(o1 + shift, o1 + shift + l1))
                rows += got
            assert rows == list(range(shift, shift + n_new))
