"""Host-side mirror of the reference interface, checked without a GPU: data containers, featurizer specs, model
constructors, epoch bookkeeping helpers, the synthetic workloads of BASELINE.json.  (Everything that computes goes
through libisokann_b200.so and is covered by the ``-m gpu`` tests.)"""
import numpy as np
import pytest


def test_simulationdata_shapes_accessors_and_slicing(pkg):
    rng = np.random.default_rng(0)
    xs = rng.normal(size=(6, 10)).astype(np.float32)
    ys = rng.normal(size=(6, 3, 10)).astype(np.float32)
    w = rng.uniform(size=(3, 10))
    d = pkg.SimulationData(xs, ys, featurizer=pkg.FeaturesAll(), weights=w)       # SimulationData(xs, ys)
    assert len(d) == 10 and d.nk() == 3 and d.featuredim() == 1                  # 2 atoms -> 1 distance
    assert pkg.coords(d) is d.getcoords() and pkg.propcoords(d).shape == (6, 3, 10)
    sub = d[slice(4, None)]                                                      # data[5:end] in Julia
    assert len(sub) == 6 and np.array_equal(sub.coords[0], xs[:, 4:]) and np.array_equal(sub.coords[1], ys[:, :, 4:])
    assert np.array_equal(sub.weights, w[:, 4:]) and sub.featurizer is d.featurizer
    picked = d[[7, 1, 1]]                                                        # getobs with an index vector
    assert np.array_equal(picked.coords[0], xs[:, [7, 1, 1]])
    d2 = pkg.SimulationData(pkg.ExternalSimulation(), (xs, ys))                  # SimulationData(sim, (xs, ys))
    assert isinstance(d2.featurizer, pkg.FeaturesCoords) and d2.featuredim() == 6
    with pytest.raises(AssertionError):
        pkg.SimulationData(xs, ys[:, :, :9])


def test_featurizer_specs(pkg):
    # (kind, n_atoms, 1-based index list, feature dimension) handed to isokann_create
    assert pkg.FeaturesCoords().spec(66)[0] == "identity" and pkg.FeaturesCoords().spec(66)[3] == 66
    kind, na, idx, F = pkg.FeaturesAll().spec(66)
    assert (kind, na, F) == ("allpairs", 22, 231)
    kind, na, idx, F = pkg.FeaturesAtoms([2, 5, 7]).spec(66)
    assert (kind, na, idx, F) == ("atoms", 22, [2, 5, 7], 3)
    kind, na, idx, F = pkg.FeaturesPairs([(1, 22), (5, 7)]).spec(66)
    assert (kind, na, idx, F) == ("pairs", 22, [1, 22, 5, 7], 2)


def test_model_constructors_and_optimiser_rules(pkg):
    m = pkg.pairnet(n=595)                                       # src/models.jl:65-69
    assert m.widths == [595, 71, 8, 1] and m.layernorm
    assert pkg.inputdim(m) == 595 and pkg.outputdim(m) == 1
    assert m.num_params() == 2 * 595 + 595 * 71 + 71 + 71 * 8 + 8 + 8 + 1 == 44091
    s = pkg.smallnet(2)                                          # src/models.jl:102-108
    assert s.widths == [2, 8, 8, 8, 1] and not s.layernorm
    dn = pkg.densenet([10, 4, 3], layernorm=True, rng=np.random.default_rng(1))
    lim = np.sqrt(6.0 / (10 + 4))                                # glorot uniform, zero bias, LayerNorm scale 1 / bias 0
    flat = dn.flat()
    assert flat.dtype == np.float32 and flat.size == dn.num_params()
    assert np.all(flat[:10] == 1) and np.all(flat[10:20] == 0)
    assert np.abs(flat[20:20 + 40]).max() <= lim and np.all(flat[60:64] == 0)
    a, n = pkg.AdamRegularized(), pkg.NesterovRegularized()
    assert (a.kind, a.eta, a.reg) == ("adam", 1e-3, 1e-4) and (n.kind, n.eta, n.reg, n.rho) == ("nesterov", 1e-3, 1e-4, 0.9)


def test_epoch_bookkeeping_helpers(pkg):
    bb = pkg.parallel.batch_bounds
    assert bb(1000, 100) == [(i * 100, 100) for i in range(10)]
    assert bb(1050, 100) == [(i * 100, 100) for i in range(10)]                   # tail of 50 dropped (partial=false)
    assert bb(1050, 100, partial=True)[-1] == (1000, 50)
    assert bb(50, 100) == [(0, 50)] and bb(50, 0) == [(0, 50)]                    # N < B and B = 0: one full batch
    perm = np.arange(1, 21)
    parts = [pkg.parallel.rank_batch_slice(perm, 5, 10, 3, r) for r in range(3)]
    assert np.array_equal(np.concatenate(parts), perm[5:15]) and [len(p) for p in parts] == [4, 3, 3]
    offs = [pkg.parallel.shard_range(10, 4, r) for r in range(4)]
    assert offs == [(0, 3), (3, 3), (6, 2), (8, 2)]


def test_synthetic_workloads_follow_baseline_configs(pkg):
    W = pkg.synthetic.WORKLOADS
    assert W["c1"].widths == [231, 38, 6, 1] and (W["c1"].N, W["c1"].K) == (100, 5)
    assert W["c2"].widths == [2, 8, 8, 8, 1] and (W["c2"].N, W["c2"].K) == (100_000, 8)
    assert W["c3"].widths == [595, 71, 8, 1] and (W["c3"].N, W["c3"].K) == (100_000, 8)
    assert W["c4"].widths[-1] == 3 and (W["c4"].N, W["c4"].K) == (1_000_000, 8)
    assert W["c5"].widths == [595, 2048, 2048, 1] and (W["c5"].N, W["c5"].K) == (1_000_000, 16)
    for name in ("c1", "c2", "c3"):
        w = W[name]
        xs, ys = pkg.synthetic.make_data(w, 64, 3)
        xs2, ys2 = pkg.synthetic.make_data(w, 64, 3)
        assert xs.shape == (w.D, 64) and ys.shape == (w.D, 3, 64) and xs.dtype == np.float32
        assert xs.flags["F_CONTIGUOUS"] and ys.flags["F_CONTIGUOUS"]               # Julia memory order
        assert np.array_equal(xs, xs2) and np.array_equal(ys, ys2)                 # seeded: oracle and library agree
        p = pkg.synthetic.make_perms(w, 64, 2)
        assert p.shape == (2, 64) and sorted(p[0]) == list(range(1, 65))           # 1-based like Julia's randperm


def test_iso_rejects_cpu_and_mismatched_model(pkg):
    xs = np.zeros((6, 4), np.float32)
    ys = np.zeros((6, 2, 4), np.float32)
    data = pkg.SimulationData(xs, ys, featurizer=pkg.FeaturesAll())
    with pytest.raises(RuntimeError):
        pkg.Iso(data, gpu=False)
    with pytest.raises(AssertionError):
        pkg.Iso(data, model=pkg.pairnet(n=7))                    # featuredim is 1


def test_exchange_slices_cover_every_element_once(pkg):
    """the ownership rule of the peer-memory gradient exchange (csrc/p2p.cu mirrors parallel.exchange_slices):
    every element of [lo, hi) is summed by exactly one rank, vector runs start on 16-byte boundaries"""
    par = pkg.parallel
    for world in (2, 3, 4, 8):
        for lo, hi in ((0, 179), (1221798, 5420201), (0, 1221798), (5, 7), (8, 20), (3, 4), (0, 44093), (7, 7)):
            owner = {}
            for r, runs in enumerate(par.exchange_slices(lo, hi, world)):
                for a, b in runs:
                    assert lo <= a < b <= hi
                    for i in range(a, b) if b - a < 4096 else (a, (a + b) // 2, b - 1):
                        assert i not in owner, (lo, hi, world, i)
                        owner[i] = r
                    if b - a >= 4 and (a % 4 == 0):
                        assert (b - a) % 4 == 0 or r == 0
            covered = sum(b - a for runs in par.exchange_slices(lo, hi, world) for a, b in runs)
            assert covered == hi - lo, (lo, hi, world, covered)
