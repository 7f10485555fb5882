"""Shared test plumbing: run the library (through the C ABI) and the oracle on identical inputs."""
from __future__ import annotations

import numpy as np


def oracle_model(oracle, widths, layernorm, seed):
    m = oracle.Model(list(widths), layernorm)
    return oracle.init_params(m, np.random.default_rng(seed))


def records(a):
    """Julia-shaped (D, N) / (D, K, N) -> records layout (N, D) / (N, K, D), same memory"""
    return np.ascontiguousarray(np.asarray(a).T)


def oracle_features(oracle, w, xs, ys):
    if w.featurizer == "identity":
        return records(xs).astype(np.float32), records(ys).astype(np.float32)
    return oracle.flatpairdists(records(xs)), oracle.flatpairdists(records(ys))


def make_iso(pkg, w, xs, ys, flat, opt="nesterov", target=None, minibatch=None, target_opts=None, **kw):
    feat = pkg.FeaturesAll() if w.featurizer == "allpairs" else pkg.FeaturesCoords()
    data = pkg.SimulationData(xs, ys, featurizer=feat)
    model = pkg.Chain(list(w.widths), w.layernorm).load_flat(flat)
    rule = pkg.AdamRegularized() if opt == "adam" else pkg.NesterovRegularized()
    tk = target or w.target
    tobj = {"shiftscale": pkg.TransformShiftscale, "isa": pkg.TransformISA,
            "pinv": pkg.TransformPseudoInv}[tk](**(target_opts or {}))
    return pkg.Iso(data, opt=rule, model=model, target=tobj, minibatch=w.minibatch if minibatch is None else minibatch,
                   **kw)


def run_pair(pkg, oracle, name, N, K, minibatch, n_iter, opt="nesterov", target=None, epochs=1, gemm="auto",
             widths=None, target_opts=None):
    """one workload through both implementations; returns losses, final chi, first target, stats"""
    import copy
    w = copy.deepcopy(pkg.synthetic.WORKLOADS[name])
    if widths is not None:
        w.widths = list(widths)
    tk = target or w.target
    xs, ys = pkg.synthetic.make_data(w, N, K)
    perms = pkg.synthetic.make_perms(w, N, n_iter * epochs)
    om = oracle_model(oracle, w.widths, w.layernorm, w.seed + 1)
    flat0 = oracle.flatten_params(om)

    topts = dict(target_opts or {})
    iso = make_iso(pkg, w, xs, ys, flat0, opt, tk, minibatch, target_opts=topts, gemm=gemm)
    t_lib = pkg.isotarget(iso)                       # target of the first iteration (before any training)
    chi0_lib = pkg.chis(iso)
    iso.engine.reset_stats()
    pkg.run_(iso, n_iter, epochs, perms=perms)
    stats = iso.engine.stats()
    chi_lib = pkg.chis(iso)
    flat_lib = iso.engine.download_params()

    xsf, ysf = oracle_features(oracle, w, xs, ys)
    cfg = oracle.OptConfig(kind=opt)
    st = oracle.opt_init(cfg, flat0.size)
    t_ref = oracle.isotarget(tk, om, xsf, ysf, **topts)
    chi0_ref = oracle.forward(om, xsf)
    losses_ref = oracle.run(om, xsf, ysf, cfg, st, n_iter, minibatch, list(perms), tk, epochs, **topts)
    chi_ref = oracle.forward(om, xsf)
    return {
        "loss_lib": np.array(iso.losses), "loss_ref": np.array(losses_ref),
        "chi_lib": records(chi_lib), "chi_ref": chi_ref,
        "chi0_lib": records(chi0_lib), "chi0_ref": chi0_ref,
        "target_lib": records(t_lib), "target_ref": t_ref,
        "flat_lib": flat_lib, "flat_ref": oracle.flatten_params(om),
        "stats": stats, "iso": iso,
    }
