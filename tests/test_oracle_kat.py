"""Known-answer tests that pin the oracle to the cited reference code (KATs 1-10).
CPU only."""
import itertools

import numpy as np
import pytest


def test_halfinds_order(oracle):
    # reference src/utils/pairdists.jl:50-56: findall over a column-major strict upper triangle
    assert oracle.halfinds(3) == [(1, 2), (1, 3), (2, 3)]
    assert oracle.halfinds(4) == [(1, 2), (1, 3), (2, 3), (1, 4), (2, 4), (3, 4)]
    # feature index f(i,j) = (j-1)(j-2)/2 + i
    for n in (5, 22, 35):
        for f, (i, j) in enumerate(oracle.halfinds(n), start=1):
            assert f == (j - 1) * (j - 2) // 2 + i


def test_flatpairdists_adp_geometry(oracle, pkg):
    x = pkg.synthetic.ADP_NM
    f = oracle.flatpairdists(x.reshape(1, -1))[0]
    assert f.shape == (231,) and f.dtype == np.float32
    k = 0
    for j in range(1, 22):
        for i in range(j):
            assert f[k] == np.float32(np.linalg.norm(x[i] - x[j]))
            k += 1
    # C-N peptide bond (atoms 5-7) ~ 0.133 nm, CA-C (9-15) ~ 0.15 nm: sanity of the units
    idx = {p: n for n, p in enumerate(oracle.halfinds(22))}
    assert 0.12 < f[idx[(5, 7)]] < 0.15


def test_flatpairdists_rigid_invariance_and_gram_form(oracle):
    rng = np.random.default_rng(0)
    x = rng.normal(size=(7, 30))
    Q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    y = (x.reshape(7, 10, 3) @ Q.T + rng.normal(size=3)).reshape(7, 30)
    a = oracle.flatpairdists(x, out_dtype=np.float64)
    b = oracle.flatpairdists(y, out_dtype=np.float64)
    assert np.allclose(a, b, atol=1e-12)
    assert (a >= 0).all()
    # the reference's literal Gram formulation (pairdists.jl:32-35) agrees in float64
    g = oracle.flatpairdists_gram(x)
    assert np.allclose(a, g, atol=1e-7)


def test_flatpairdists_cols_and_pdists(oracle):
    rng = np.random.default_rng(1)
    x = rng.normal(size=(4, 18))
    cols = [2, 5, 6]
    a = oracle.flatpairdists(x, cols, out_dtype=np.float64)
    pairs = [(2, 5), (2, 6), (5, 6)]
    b = oracle.pdists(x, pairs, out_dtype=np.float64)
    assert np.allclose(a, b)


def test_shiftscale(oracle):
    ks = np.array([[0.3], [0.5], [0.9], [0.4]], dtype=np.float32)
    t = oracle.shiftscale(ks)
    assert t.min() == 0 and t.max() == 1
    assert np.allclose(t[:, 0], (ks[:, 0] - 0.3) / 0.6, rtol=1e-6)
    with pytest.raises(oracle.DomainError) as e:
        oracle.shiftscale(np.full((5, 1), 0.25, dtype=np.float32))
    assert e.value.code == 1
    with pytest.raises(AssertionError):
        oracle.shiftscale(np.zeros((5, 2), dtype=np.float32))


def test_expectation_of_linear_model(oracle):
    rng = np.random.default_rng(2)
    m = oracle.densenet([6, 2], layernorm=False, act=oracle.isokann_oracle.ACT_IDENTITY, rng=rng)
    ys = rng.normal(size=(9, 4, 6)).astype(np.float32)
    e = oracle.expectation(m, ys)
    assert np.allclose(e, oracle.forward(m, ys.mean(axis=1)), atol=1e-5)


def test_isa_selected_rows_map_to_unit_vectors(oracle):
    rng = np.random.default_rng(3)
    ks = rng.normal(size=(200, 3)).astype(np.float32)
    ind = oracle.indexmap(ks.astype(np.float64))
    assert ind[0] == int(np.argmax(np.linalg.norm(ks.astype(np.float64), axis=1)))
    assert len(set(ind)) == 3
    A = oracle.myisa(ks)
    target = ks.astype(np.float64) @ A
    assert np.allclose(target[ind], np.eye(3), atol=1e-9)


def test_fixperm_recovers_every_permutation(oracle):
    rng = np.random.default_rng(4)
    old = rng.normal(size=(50, 3))
    for p in itertools.permutations(range(3)):
        shuffled = old[:, list(p)]
        assert np.array_equal(oracle.fixperm(shuffled, old), old)


def test_pinv_target_is_projection(oracle):
    rng = np.random.default_rng(5)
    m = oracle.pairnet(10, nout=2, rng=rng)
    xsf = rng.normal(size=(40, 10)).astype(np.float32)
    ysf = rng.normal(size=(40, 3, 10)).astype(np.float32)
    t = oracle.isotarget_pinv(m, xsf, ysf, normalize=False, eigenvecs=False, permute=False)
    chi = oracle.forward(m, xsf).T.astype(np.float64)
    kchi = oracle.expectation(m, ysf).T.astype(np.float64)
    proj = chi @ np.linalg.pinv(kchi) @ kchi
    assert np.allclose(t.T, proj, rtol=1e-3, atol=1e-4)


def test_optimiser_first_step(oracle):
    theta = np.array([0.5, -0.25, 2.0], dtype=np.float32)
    g = np.array([0.1, -0.3, 0.0], dtype=np.float32)
    gp = g + np.float32(1e-4) * theta
    cfg = oracle.OptConfig(kind="adam")
    st = oracle.opt_init(cfg, 3)
    new = oracle.opt_update(cfg, st, theta, g)
    assert np.allclose(new, theta - 1e-3 * gp / (np.abs(gp) + 1e-8), rtol=1e-5)
    assert np.allclose(st.beta_t, [0.9 ** 2, 0.999 ** 2], rtol=1e-6)
    cfg = oracle.OptConfig(kind="nesterov")
    st = oracle.opt_init(cfg, 3)
    new = oracle.opt_update(cfg, st, theta, g)
    assert np.allclose(new, theta - 1.9 * 1e-3 * gp, rtol=1e-5)
    assert np.allclose(st.m, -1e-3 * gp, rtol=1e-6)


def test_train_batch_bookkeeping(oracle):
    rng = np.random.default_rng(6)
    m = oracle.pairnet(12, rng=rng)
    xsf = rng.normal(size=(25, 12)).astype(np.float32)
    t = rng.uniform(size=(25, 1)).astype(np.float32)
    cfg = oracle.OptConfig()
    calls = []
    orig = oracle.isokann_oracle.opt_update

    def spy(c, s, th, g):
        calls.append(1)
        return orig(c, s, th, g)
    oracle.isokann_oracle.opt_update = spy
    try:
        perm = rng.permutation(25) + 1
        oracle.train_batch(m, xsf, t, cfg, oracle.opt_init(cfg, oracle.num_params(m)), 10, perm)
        assert len(calls) == 2                      # floor(25/10), tail of 5 dropped
        calls.clear()
        oracle.train_batch(m, xsf, t, cfg, oracle.opt_init(cfg, oracle.num_params(m)), 100, perm)
        assert len(calls) == 1                      # N < minibatch -> one full batch
        calls.clear()
        oracle.train_batch(m, xsf, t, cfg, oracle.opt_init(cfg, oracle.num_params(m)), 0, perm)
        assert len(calls) == 1
    finally:
        oracle.isokann_oracle.opt_update = orig


def test_pairnet_layer_rule(oracle, pkg):
    assert oracle.pairnet_layers(595) == [595, 71, 8, 1]
    assert oracle.pairnet_layers(231) == [231, 38, 6, 1]
    assert oracle.pairnet_layers(66) == [66, 16, 4, 1]
    assert pkg.pairnet_layers(595) == [595, 71, 8, 1]
    assert pkg.pairnet_layers(231, nout=3) == [231, 38, 6, 3]


def test_gradient_matches_finite_differences(oracle):
    rng = np.random.default_rng(7)
    m = oracle.pairnet(9, nout=2, rng=rng)
    m.ln_scale = rng.uniform(0.5, 1.5, 9).astype(np.float32)
    m.ln_bias = rng.normal(size=9).astype(np.float32) * 0.1
    x = rng.normal(size=(6, 9)).astype(np.float32)
    y = rng.normal(size=(6, 2)).astype(np.float32)
    w = np.array([0.7, 1.3], dtype=np.float32)
    _, g = oracle.batch_loss_and_grad(m, x, y, w)
    flat = oracle.flatten_params(m).astype(np.float64)

    def loss64(f):
        mm = oracle.unflatten_params(m.copy(), f.astype(np.float32))
        # evaluate in float64 for a clean finite difference
        z = x.astype(np.float64)
        z = oracle.layernorm(z) * mm.ln_scale.astype(np.float64) + mm.ln_bias.astype(np.float64)
        for i in range(mm.nlayers):
            a = z @ mm.W[i].astype(np.float64) + mm.b[i].astype(np.float64)
            z = oracle.sigmoid(a) if i < mm.nlayers - 1 else a
        return np.sum(((z - y) * w) ** 2) / x.shape[0]
    for k in rng.choice(flat.size, 25, replace=False):
        h = 1e-3
        fp, fm = flat.copy(), flat.copy()
        fp[k] += h
        fm[k] -= h
        fd = (loss64(fp) - loss64(fm)) / (2 * h)
        assert abs(fd - g[k]) < 2e-3 * max(1.0, abs(fd)), (k, fd, g[k])


def test_chi_vjp_matches_finite_differences(oracle):
    rng = np.random.default_rng(8)
    m = oracle.pairnet(15, nout=2, rng=rng)                 # 6 atoms -> 15 pair distances
    m.ln_scale = rng.uniform(0.5, 1.5, 15).astype(np.float32)
    m.ln_bias = (0.1 * rng.normal(size=15)).astype(np.float32)
    pairs0 = oracle.pair_table(6)
    x = rng.normal(size=(3, 18))
    cot = rng.normal(size=(3, 2))

    def scalar(xx):
        f = oracle.flatpairdists(xx, out_dtype=np.float64)
        mm = m
        z = oracle.layernorm(f) * mm.ln_scale.astype(np.float64) + mm.ln_bias.astype(np.float64)
        for i in range(mm.nlayers):
            a = z @ mm.W[i].astype(np.float64) + mm.b[i].astype(np.float64)
            z = oracle.sigmoid(a) if i < mm.nlayers - 1 else a
        return float((z * cot).sum())
    g = oracle.chi_vjp(m, x, cot, pairs0)
    for _ in range(20):
        i, k = rng.integers(0, 3), rng.integers(0, 18)
        h = 1e-6
        xp, xm = x.copy(), x.copy()
        xp[i, k] += h
        xm[i, k] -= h
        fd = (scalar(xp) - scalar(xm)) / (2 * h)
        assert abs(fd - g[i, k]) < 1e-6 * max(1.0, abs(fd)), (fd, g[i, k])
    # rigid translation does not change chi: the gradient sums to zero over atoms
    assert np.abs(g.reshape(3, 6, 3).sum(axis=1)).max() < 1e-12


# ---------------------------------------------------------------------------------------------
# independent implementations (torch autograd / torch.optim): not the reference, but a second,
# widely used implementation of the same published arithmetic -- a tighter pin than finite differences
# ---------------------------------------------------------------------------------------------
def _torch_model(oracle, om, x64):
    """the oracle Model evaluated with torch ops in float64; returns (chi, leaf parameter tensors in flat order)"""
    import torch
    t = lambda a: torch.tensor(np.asarray(a, dtype=np.float64), requires_grad=True)
    leaves = []
    h = torch.tensor(x64)
    if om.layernorm:
        g, b = t(om.ln_scale), t(om.ln_bias)
        leaves += [g, b]
        h = torch.nn.functional.layer_norm(h, (h.shape[1],), eps=om.ln_eps ** 2) * g + b   # Flux: sqrt(var + eps^2)
    for i, (W, bb) in enumerate(zip(om.W, om.b)):
        Wt, bt = t(W), t(bb)
        leaves += [Wt, bt]
        h = h @ Wt + bt
        if i < om.nlayers - 1:
            h = torch.sigmoid(h)
    return h, leaves


@pytest.mark.parametrize("d", [1, 3])
def test_gradient_matches_torch_autograd(oracle, d):
    import torch
    rng = np.random.default_rng(0)
    om = oracle.densenet([12, 7, 5, d], layernorm=True, rng=rng)
    om.ln_scale = rng.uniform(0.5, 1.5, 12).astype(np.float32)
    om.ln_bias = (0.1 * rng.normal(size=12)).astype(np.float32)
    om.b = [(0.1 * rng.normal(size=b.shape)).astype(np.float32) for b in om.b]
    x = rng.normal(size=(9, 12)).astype(np.float32)
    y = rng.uniform(size=(9, d)).astype(np.float32)
    w = None if d == 1 else rng.uniform(0.5, 2.0, d).astype(np.float32)
    l, g = oracle.batch_loss_and_grad(om, x, y, w)
    chi, leaves = _torch_model(oracle, om, x.astype(np.float64))
    wt = torch.ones(d, dtype=torch.float64) if w is None else torch.tensor(w.astype(np.float64))
    lt = (((chi - torch.tensor(y.astype(np.float64))) * wt) ** 2).sum()
    (lt / x.shape[0]).backward()
    # flat order: [gamma, beta,] W1 (in Julia's memory order = the oracle's (in, out) row-major), b1, ...
    gt = np.concatenate([p.grad.numpy().ravel() for p in leaves])
    assert np.isclose(l, lt.item(), rtol=1e-5)
    assert np.allclose(g, gt, rtol=2e-4, atol=2e-6), np.abs(g - gt).max()


@pytest.mark.parametrize("kind", ["adam", "nesterov"])
def test_optimiser_matches_torch_optim(oracle, kind):
    # Optimisers.jl OptimiserChain(WeightDecay(lam), Adam/Nesterov) == torch Adam / SGD(nesterov) with L2 weight decay
    import torch
    rng = np.random.default_rng(1)
    theta = rng.normal(size=50).astype(np.float32)
    cfg = oracle.OptConfig(kind=kind, eta=1e-2, lam=1e-3)
    st = oracle.opt_init(cfg, theta.size)
    p = torch.tensor(theta.astype(np.float64), requires_grad=True)
    if kind == "adam":
        opt = torch.optim.Adam([p], lr=cfg.eta, betas=(cfg.beta1, cfg.beta2), eps=cfg.eps, weight_decay=cfg.lam)
    else:
        opt = torch.optim.SGD([p], lr=cfg.eta, momentum=cfg.rho, nesterov=True, weight_decay=cfg.lam)
    th = theta.copy()
    for step in range(25):
        g = rng.normal(size=50).astype(np.float32)
        th = oracle.opt_update(cfg, st, th, g)
        p.grad = torch.tensor(g.astype(np.float64))
        opt.step()
        assert np.allclose(th, p.detach().numpy(), rtol=2e-4, atol=2e-6), (step, np.abs(th - p.detach().numpy()).max())
