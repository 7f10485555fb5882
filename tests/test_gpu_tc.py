"""Parity of the tcgen05 (3xBF16 split) path for wide Dense layers against the oracle and
against the library's own FP32 CUDA-core path, through the C ABI."""
import copy

import numpy as np
import pytest

from tests.helpers import make_iso, oracle_features, oracle_model, records, run_pair

pytestmark = pytest.mark.gpu

TOL_CHI = 1e-4


def wide(pkg, widths, name="c1"):
    w = copy.deepcopy(pkg.synthetic.WORKLOADS[name])
    w.widths = list(widths)
    return w


@pytest.mark.parametrize("widths,N,K", [([231, 256, 256, 1], 300, 4), ([231, 264, 512, 2], 77, 3),
                                        ([231, 512, 1], 129, 2)])
def test_tc_forward_matches_oracle_and_fp32_path(pkg, oracle, widths, N, K):
    w = wide(pkg, widths)
    xs, ys = pkg.synthetic.make_data(w, N, K)
    om = oracle_model(oracle, w.widths, True, 5)
    rng = np.random.default_rng(9)
    om.ln_scale = rng.uniform(0.5, 1.5, w.F).astype(np.float32)
    om.ln_bias = (0.1 * rng.normal(size=w.F)).astype(np.float32)
    om.b = [(0.1 * rng.normal(size=b.shape)).astype(np.float32) for b in om.b]
    flat = oracle.flatten_params(om)
    tgt = "isa" if widths[-1] > 1 else "shiftscale"
    tc = make_iso(pkg, w, xs, ys, flat, target=tgt, gemm="tc")
    fp = make_iso(pkg, w, xs, ys, flat, target=tgt, gemm="fp32")
    xsf, ysf = oracle_features(oracle, w, xs, ys)
    chi_ref = oracle.forward(om, xsf)
    chi_tc, chi_fp = records(pkg.chis(tc)), records(pkg.chis(fp))
    assert np.allclose(chi_fp, chi_ref, rtol=TOL_CHI, atol=1e-5)
    assert np.allclose(chi_tc, chi_ref, rtol=TOL_CHI, atol=5e-5), np.abs(chi_tc - chi_ref).max()
    k_ref = oracle.expectation(om, ysf)
    assert np.allclose(records(pkg.koopman(tc)), k_ref, rtol=TOL_CHI, atol=5e-5)


@pytest.mark.parametrize("opt", ["adam", "nesterov"])
def test_tc_one_iteration_parity(pkg, oracle, opt):
    r = run_pair(pkg, oracle, "c1", N=512, K=3, minibatch=128, n_iter=1, opt=opt, gemm="tc",
                 widths=[231, 256, 320, 1])
    assert np.allclose(r["target_lib"], r["target_ref"], atol=5e-4)
    assert np.allclose(r["loss_lib"], r["loss_ref"], rtol=2e-3)
    assert np.allclose(r["chi_lib"], r["chi_ref"], rtol=1e-3, atol=1e-3)
    scale = np.abs(r["flat_ref"]).max()
    assert np.abs(r["flat_lib"] - r["flat_ref"]).max() < 2e-3 * scale


def test_tc_gradient_first_step(pkg, oracle):
    # recover the gradient from one Nesterov step (see test_first_step_gradient_through_params)
    w = wide(pkg, [231, 256, 256, 1])
    N = 200
    xs, ys = pkg.synthetic.make_data(w, N, 2)
    om = oracle_model(oracle, w.widths, True, 5)
    rng = np.random.default_rng(2)
    om.ln_scale = rng.uniform(0.5, 1.5, w.F).astype(np.float32)
    om.ln_bias = (0.1 * rng.normal(size=w.F)).astype(np.float32)
    flat0 = oracle.flatten_params(om)
    iso = make_iso(pkg, w, xs, ys, flat0, minibatch=0, gemm="tc")
    pkg.isotarget(iso)
    pkg.train_batch_(iso, np.arange(1, N + 1))
    g_lib = (flat0 - iso.engine.download_params()) / (1.9e-3) - 1e-4 * flat0
    xsf, ysf = oracle_features(oracle, w, xs, ys)
    t = oracle.isotarget_shiftscale(om, xsf, ysf)
    _, g_ref = oracle.batch_loss_and_grad(om, xsf, t, None)
    err = np.abs(g_lib - g_ref).max() / np.abs(g_ref).max()
    assert err < 5e-3, err


def test_tc_large_batch_split_k(pkg, oracle):
    r = run_pair(pkg, oracle, "c1", N=8192, K=1, minibatch=0, n_iter=1, opt="adam", gemm="tc",
                 widths=[231, 256, 256, 1])
    assert np.allclose(r["loss_lib"], r["loss_ref"], rtol=2e-3)
    assert np.allclose(r["chi_lib"], r["chi_ref"], rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("target,topts", [("pinv", {"eigenvecs": False, "permute": False}), ("isa", {})])
def test_tc_nd_target_iteration(pkg, oracle, target, topts):
    # (fixperm at a random initialisation is a near-tie and the Schur vectors of eigenvecs=True are rounding-sensitive, see DESIGN.md section 2: a 1e-7 change of Kinv
    #  may flip them, so multi-iteration comparisons across different arithmetic use the stable options)
    r = run_pair(pkg, oracle, "c4", N=400, K=3, minibatch=100, n_iter=2, opt="adam", target=target, gemm="tc",
                 widths=[231, 256, 256, 3], target_opts=topts)
    assert np.allclose(r["loss_lib"], r["loss_ref"], rtol=1e-2)
    assert np.allclose(r["chi_lib"], r["chi_ref"], rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize("widths,target", [([231, 256, 320, 1], "shiftscale"), ([231, 512, 264, 3], "isa"),
                                           ([231, 300, 1], "shiftscale")])
def test_fp16x2_inference_forward(pkg, oracle, monkeypatch, widths, target):
    """ISOKANN_TC_FWD=fp16x2 (opt-in): the inference forward (chis, Koopman pass) with fp16 (hi, lo) activations,
    weights rounded once to fp16 and two MMAs per product, against the oracle and the default bf16 x 3 path; then
    three iterations (the training steps stay bf16 x 3, only their targets come from the fp16 forward).
    Rounding every weight once to 11 bits costs ~2e-4 relative on chi (measured: profiles/r02_split_accuracy_gpu.md),
    which is why this mode is not the default: the tolerances below are 3-4x the default path's."""
    w = wide(pkg, widths)
    N, K = 1300, 3
    xs, ys = pkg.synthetic.make_data(w, N, K)
    om = oracle_model(oracle, w.widths, True, 5)
    rng = np.random.default_rng(9)
    om.ln_scale = rng.uniform(0.5, 1.5, w.F).astype(np.float32)
    om.ln_bias = (0.1 * rng.normal(size=w.F)).astype(np.float32)
    om.b = [(0.1 * rng.normal(size=b.shape)).astype(np.float32) for b in om.b]
    flat = oracle.flatten_params(om)
    ref = make_iso(pkg, w, xs, ys, flat, opt="adam", target=target, minibatch=400, gemm="tc")
    monkeypatch.setenv("ISOKANN_TC_FWD", "fp16x2")
    h2 = make_iso(pkg, w, xs, ys, flat, opt="adam", target=target, minibatch=400, gemm="tc")
    xsf, ysf = oracle_features(oracle, w, xs, ys)
    chi_ref, k_ref = oracle.forward(om, xsf), oracle.expectation(om, ysf)
    c2, k2 = records(pkg.chis(h2)), records(pkg.koopman(h2))
    assert np.allclose(c2, chi_ref, rtol=3e-4, atol=1e-4), np.abs(c2 - chi_ref).max()
    assert np.allclose(k2, k_ref, rtol=3e-4, atol=1e-4), np.abs(k2 - k_ref).max()
    assert np.abs(c2 - records(pkg.chis(ref))).max() < 3e-4 * max(1.0, np.abs(chi_ref).max())
    perms = pkg.synthetic.make_perms(w, N, 3)
    pkg.run_(h2, 3, perms=perms)
    pkg.run_(ref, 3, perms=perms)
    assert np.allclose(h2.losses, ref.losses, rtol=2e-2), (h2.losses, ref.losses)
    assert np.abs(pkg.chis(h2) - pkg.chis(ref)).max() < 5e-3


def test_tc_mode_rejects_narrow_nets(pkg):
    model = pkg.pairnet(n=231)
    with pytest.raises(pkg.IsokannError):
        pkg.Engine(model, pkg.NesterovRegularized(), "allpairs", 22, gemm="tc")


def test_tc_two_cta_kernel(pkg, oracle):
    # large enough for the cta_group::2 kernel (>= 74 tile pairs): forward (split + fused-dot epilogues) and a
    # full-batch training step (data-gradient GEMM) against the oracle
    r = run_pair(pkg, oracle, "c1", N=40000, K=1, minibatch=0, n_iter=1, opt="adam", gemm="tc",
                 widths=[231, 264, 512, 1])
    assert np.allclose(r["chi0_lib"], r["chi0_ref"], rtol=TOL_CHI, atol=5e-5)
    assert np.allclose(r["loss_lib"], r["loss_ref"], rtol=2e-3)
    assert np.allclose(r["chi_lib"], r["chi_ref"], rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("name,widths,atoms,N", [("c1", [231, 256, 1], None, 77), ("c3", [595, 256, 256, 1], None, 1000),
                                                ("c1", [91, 256, 1], [1, 2, 4, 5, 7, 9, 10, 12, 15, 17, 19, 20, 21, 22], 333)])
def test_record_featurizer_matches_feature_featurizer(pkg, oracle, monkeypatch, name, widths, atoms, N):
    # lane = record kernel (featurize_rec.cu, default for upper-triangle featurizers) against the lane = feature
    # kernel (ISOKANN_FEAT_REC=0) and the oracle: ragged record blocks, atom subsets, gathered minibatches
    w = wide(pkg, widths, name)
    K = 3
    xs, ys = pkg.synthetic.make_data(w, N, K)
    om = oracle_model(oracle, w.widths, True, 5)
    rng = np.random.default_rng(3)
    om.ln_scale = rng.uniform(0.5, 1.5, widths[0]).astype(np.float32)
    om.ln_bias = (0.1 * rng.normal(size=widths[0])).astype(np.float32)
    flat = oracle.flatten_params(om)
    feat = pkg.FeaturesAll() if atoms is None else pkg.FeaturesAtoms(atoms)

    def build():
        data = pkg.SimulationData(xs, ys, featurizer=feat)
        model = pkg.Chain(list(w.widths), True).load_flat(flat)
        return pkg.Iso(data, opt=pkg.AdamRegularized(), model=model, minibatch=max(32, N // 3), gemm="tc")
    new = build()
    monkeypatch.setenv("ISOKANN_FEAT_REC", "0")
    old = build()
    xf = oracle.flatpairdists(records(xs), atoms)
    chi_ref = oracle.forward(om, xf)
    c_new, c_old = records(pkg.chis(new)), records(pkg.chis(old))
    assert np.allclose(c_new, chi_ref, rtol=TOL_CHI, atol=5e-5), np.abs(c_new - chi_ref).max()
    assert np.abs(c_new - c_old).max() < 2e-5
    assert np.abs(pkg.koopman(new) - pkg.koopman(old)).max() < 2e-5
    perms = pkg.synthetic.make_perms(w, N, 2)
    pkg.run_(new, 2, perms=perms)                       # gathered minibatches through the same kernel
    pkg.run_(old, 2, perms=perms)
    assert np.allclose(new.losses, old.losses, rtol=1e-4)
    assert np.abs(pkg.chis(new) - pkg.chis(old)).max() < 2e-4


@pytest.mark.parametrize("A", [12, 50])
def test_record_featurizer_split_output_other_atom_counts(pkg, oracle, A):
    # split (hi, lo) rows of the lane = record kernel for atom counts outside the BASELINE shapes: F = 66 fills
    # exactly one 64-column octet plus a tail, F = 1225 needs 1280-column rows (one block per SM)
    F = A * (A - 1) // 2
    widths = [F, 256, 1]
    rng = np.random.default_rng(A)
    N, K = 45, 2
    xs = np.asfortranarray(rng.normal(scale=0.5, size=(3 * A, N)).astype(np.float32))
    ys = np.asfortranarray((xs[:, None, :] + 0.03 * rng.normal(size=(3 * A, K, N))).astype(np.float32))
    om = oracle_model(oracle, widths, True, 5)
    om.ln_scale = rng.uniform(0.5, 1.5, F).astype(np.float32)
    om.ln_bias = (0.1 * rng.normal(size=F)).astype(np.float32)
    data = pkg.SimulationData(xs, ys, featurizer=pkg.FeaturesAll())
    model = pkg.Chain(widths, True).load_flat(oracle.flatten_params(om))
    iso = pkg.Iso(data, opt=pkg.AdamRegularized(), model=model, minibatch=N, gemm="tc")
    chi_ref = oracle.forward(om, oracle.flatpairdists(records(xs)))
    assert np.allclose(records(pkg.chis(iso)), chi_ref, rtol=TOL_CHI, atol=5e-5)
    k_ref = oracle.expectation(om, oracle.flatpairdists(records(ys)))
    assert np.allclose(records(pkg.koopman(iso)), k_ref, rtol=TOL_CHI, atol=5e-5)


@pytest.mark.parametrize("widths,target", [([231, 256, 320, 1], "shiftscale"), ([231, 512, 264, 3], "isa"),
                                            ([231, 256, 1032, 2], "isa")])
def test_fused_thin_head_matches_separate_kernels(pkg, oracle, monkeypatch, widths, target):
    # training step: thin_head_kernel (last layer + loss + both backward products in one pass) against the
    # thin_forward / loss_delta / f32_to_split / thin_dgrad sequence (ISOKANN_TC_NO_HEAD=1)
    w = wide(pkg, widths)
    N, K = 700, 2
    xs, ys = pkg.synthetic.make_data(w, N, K)
    flat = oracle.flatten_params(oracle_model(oracle, w.widths, True, 5))
    perms = pkg.synthetic.make_perms(w, N, 3)
    fused = make_iso(pkg, w, xs, ys, flat, opt="adam", target=target, minibatch=250, gemm="tc")
    monkeypatch.setenv("ISOKANN_TC_NO_HEAD", "1")
    plain = make_iso(pkg, w, xs, ys, flat, opt="adam", target=target, minibatch=250, gemm="tc")
    pkg.run_(fused, 3, perms=perms)
    pkg.run_(plain, 3, perms=perms)
    assert np.allclose(fused.losses, plain.losses, rtol=1e-10)       # same terms, another summation order
    assert np.array_equal(fused.engine.download_params(), plain.engine.download_params())
    assert fused.engine.stats()["kernel_launches"] < plain.engine.stats()["kernel_launches"]


def test_featurizer_gemm_overlap_is_bit_identical(pkg, oracle, monkeypatch):
    # ISOKANN_OVERLAP=1: the featurizer of chunk i+1 runs on a second stream beside the GEMMs of chunk i
    w = wide(pkg, [231, 256, 256, 1])
    N, K = 3000, 4
    xs, ys = pkg.synthetic.make_data(w, N, K)
    flat = oracle.flatten_params(oracle_model(oracle, w.widths, True, 5))
    plain = make_iso(pkg, w, xs, ys, flat, gemm="tc", chunk=1024)     # 12 chunks of 256 start points
    k0 = pkg.koopman(plain)
    monkeypatch.setenv("ISOKANN_OVERLAP", "1")
    over = make_iso(pkg, w, xs, ys, flat, gemm="tc", chunk=1024)
    for _ in range(3):
        assert np.array_equal(pkg.koopman(over), k0)
    pkg.run_(over, 2, perms=pkg.synthetic.make_perms(w, N, 2))
    pkg.run_(plain, 2, perms=pkg.synthetic.make_perms(w, N, 2))
    assert np.array_equal(pkg.chis(over), pkg.chis(plain))


def test_full_size_c5_properties(pkg, oracle):
    """BASELINE config 5 at full size (N = 10^6, K = 16, pairnet [595, 2048, 2048, 1]) through size-independent
    properties; the data is generated on the device like bench.py does."""
    import torch
    w = pkg.synthetic.WORKLOADS["c5"]
    N, K = w.N, w.K
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(w.seed)
    states = pkg.synthetic.villin_states(rng, w.n_atoms)
    base = torch.tensor(np.stack([s.reshape(-1) for s in states]), dtype=torch.float32, device=dev)
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    xs = (base[torch.randint(0, 2, (N,), generator=g, device=dev)] + 0.05 * torch.randn((N, w.D), generator=g, device=dev))
    ys = torch.empty((N, K, w.D), dtype=torch.float32, device=dev)
    for s in range(0, N, 1 << 16):
        e = min(N, s + (1 << 16))
        ys[s:e] = xs[s:e, None, :] + 0.03 * torch.randn((e - s, K, w.D), generator=g, device=dev)
    torch.cuda.synchronize()
    om = oracle_model(oracle, w.widths, True, w.seed + 1)
    flat = oracle.flatten_params(om)
    eng = pkg.Engine(pkg.Chain(list(w.widths), True).load_flat(flat), pkg.AdamRegularized(), "allpairs", w.n_atoms)
    eng.set_data_dev(xs, ys, w.D, K, N)
    k = eng.koopman()
    t = eng.target("shiftscale")
    assert t.shape == (1, N) and t.min() == 0.0 and t.max() == 1.0
    assert np.allclose(t, (k - k.min()) / (k.max() - k.min()), atol=1e-6)
    # rows spread over the whole data set (different chunks, different CTA pairs) against the oracle
    rows = np.array([0, 1, 65535, 65536, 65537, 500_000, 999_998, 999_999])
    ysub = ys[torch.from_numpy(rows).to(dev)].cpu().numpy()               # (8, K, D)
    ref = oracle.expectation(om, oracle.flatpairdists(ysub))
    assert np.allclose(records(k)[rows], ref, rtol=TOL_CHI, atol=5e-5)
    xsub = xs[torch.from_numpy(rows).to(dev)].cpu().numpy()
    chi = eng.chis()
    assert np.allclose(records(chi)[rows], oracle.forward(om, oracle.flatpairdists(xsub)), rtol=TOL_CHI, atol=5e-5)
    # the Koopman expectation of a constant-in-k sample equals chi of that sample (linearity of the K-mean)
    perms = pkg.synthetic.make_perms(w, N, 2)
    losses = eng.iterate("shiftscale", 2, 1, w.minibatch, perms)
    assert np.isfinite(losses).all() and losses[1] < losses[0]
    # the permutation is what defines the minibatches: the same permutation twice gives bit-identical parameters
    eng2 = pkg.Engine(pkg.Chain(list(w.widths), True).load_flat(flat), pkg.AdamRegularized(), "allpairs", w.n_atoms)
    eng2.set_data_dev(xs, ys, w.D, K, N)
    losses2 = eng2.iterate("shiftscale", 2, 1, w.minibatch, perms)
    assert np.array_equal(losses, losses2)
    assert np.array_equal(eng.download_params(), eng2.download_params())
    eng.close()
    eng2.close()
