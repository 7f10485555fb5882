"""Host-side mirror of reference src/iso.jl: ``Iso``, ``run!``, ``train_batch!``, ``chis``,
``chicoords``, ``isotarget``, ``cpu``/``save``/``load`` -- same names (``!`` -> ``_``), argument
meaning and error behaviour, so that ``iso = Iso(data); run_(iso, n)`` is the drop-in for
``iso = Iso(data); run!(iso, n)``.  All arithmetic goes through libisokann_b200.so.
"""
from __future__ import annotations

from typing import Any, List, Optional

import numpy as np

from .data import SimulationData
from .engine import DomainError, Engine
from .isotarget import TransformISA, TransformPseudoInv, TransformShiftscale
from .models import Chain, NesterovRegularized, OptimiserRule, outputdim, pairnet


def defaultmodel(data: SimulationData, **kw) -> Chain:
    """defaultmodel(d) = pairnet(n=featuredim(d)) (src/simulation.jl:126-127)"""
    return pairnet(n=data.featuredim(), **kw)


class Iso:
    """Iso(data; opt=NesterovRegularized(), model=defaultmodel(data), gpu, target, minibatch=100, loggers)
    (src/iso.jl:3-11,17-43).  ``gpu=False`` is rejected: this implementation has no CPU path."""

    def __init__(self, data: SimulationData, opt: Optional[OptimiserRule] = None, model: Optional[Chain] = None,
                 gpu: bool = True, target=None, minibatch: int = 100, loggers: Optional[List[Any]] = None,
                 device: int = 0, gemm: str = "auto", chunk: int = 0, seed: Optional[int] = None,
                 comm=None):
        if not gpu:
            raise RuntimeError("isokann.jl_b200 is GPU-only: there is no CPU fallback for the ISOKANN hot path")
        self.data = data
        self.opt = opt if opt is not None else NesterovRegularized()
        self.model = model if model is not None else defaultmodel(data)
        if target is None:  # src/iso.jl:28-34
            target = TransformShiftscale() if outputdim(self.model) == 1 else TransformISA()
        self.target = target
        self.losses: List[float] = []
        self.loggers = list(loggers or [])
        self.minibatch = minibatch
        self.rng = np.random.default_rng(seed)
        kind, n_atoms, index, F = data.featurizer.spec(data.coords[0].shape[0])
        assert F == self.model.widths[0], f"model input width {self.model.widths[0]} != feature dim {F}"
        self.engine = Engine(self.model, self.opt, kind, n_atoms, index, device=device, gemm=gemm, chunk=chunk)
        xs, ys = data.coords
        if comm is not None:  # (world, rank, unique_id): one process per GPU, N sharded over ranks
            from .parallel import shard_range
            world, rank, uid = comm
            self.engine.comm_init(world, rank, uid)
            off, n = shard_range(xs.shape[1], world, rank)
            self.engine.set_data(xs, ys[:, :, off:off + n], n_offset=off, n_local=n)
            if data.weights is not None:
                self.engine.set_koopman_weights(np.asarray(data.weights)[:, off:off + n])
        else:
            self.engine.set_data(xs, ys)
            if data.weights is not None:
                self.engine.set_koopman_weights(data.weights)

    def sync_model(self) -> Chain:
        """pull the parameters back into the host-side Chain (what ``cpu(iso).model`` holds)"""
        return self.model.load_flat(self.engine.download_params())


def draw_perm(iso: Iso) -> np.ndarray:
    """the one ``randperm(N)`` per epoch of Flux.DataLoader(shuffle=true) (src/iso.jl:181), 1-based.
    (numpy's generator, not Julia's Xoshiro: pass ``perm`` explicitly for bit-exact replays.)"""
    return iso.rng.permutation(len(iso.data)).astype(np.int64) + 1


def isotarget(iso: Iso, target=None) -> np.ndarray:
    """isotarget(iso) (src/isotarget.jl:10-12) -> (d, N); stays resident for train_batch_"""
    t = target if target is not None else iso.target
    if hasattr(t, "name"):
        return iso.engine.target(t.name, **t.opts())
    # user-defined transform: any callable (iso) -> (d, N) array, as scripts adding isotarget methods do
    out = np.asarray(t(iso))
    iso.engine.set_target(out)
    return out


def train_batch_(iso: Iso, perm: Optional[np.ndarray] = None, partial: bool = False) -> float:
    """train_batch!(model, xs, target, opt, minibatch) (src/iso.jl:179-194) on the resident target"""
    p = draw_perm(iso) if perm is None else np.asarray(perm, dtype=np.int64)
    return iso.engine.train_epoch(p, iso.minibatch, partial)


def run_(iso: Iso, n: int = 1, epochs: int = 1, perms: Optional[np.ndarray] = None, showprogress: bool = False) -> Iso:
    """run!(iso, n, epochs) (src/iso.jl:72-94).  ``perms``: optional (n*epochs, N) 1-based permutations."""
    fused = hasattr(iso.target, "name") and not iso.loggers
    N = len(iso.data)
    if perms is None:
        perms = np.stack([draw_perm(iso) for _ in range(n * epochs)]) if n * epochs > 0 else np.zeros((0, N), np.int64)
    perms = np.ascontiguousarray(perms, dtype=np.int64).reshape(n * epochs, N)
    if fused:  # no host round trips between iterations
        losses = iso.engine.iterate(iso.target.name, n, epochs, iso.minibatch, perms, **iso.target.opts())
        iso.losses.extend(float(x) for x in losses)
        return iso
    k = 0
    for _ in range(n):
        isotarget(iso)
        for _e in range(epochs):
            iso.losses.append(train_batch_(iso, perms[k]))
            k += 1
        for logger in iso.loggers:
            logger(iso)
    return iso


def chis(iso: Iso) -> np.ndarray:
    """chis(iso) (src/iso.jl:203) -> (d, N)"""
    return iso.engine.chis()


def chicoords(iso: Iso, xs) -> np.ndarray:
    """chicoords(iso, xs) (src/iso.jl:211): featurizer + model at raw coordinates"""
    return iso.engine.forward(xs, is_features=False)


def dchidx(iso: Iso, x, cot=None) -> np.ndarray:
    """dchidx(iso, x) (src/utils/minimumpath.jl:3-7): gradient of chi(x) w.r.t. the raw coordinates x (D,) or (D, M);
    for multi-dimensional chi pass the cotangent ``cot`` (d, M) (vector-Jacobian product)"""
    x = np.asarray(x)
    g = iso.engine.chi_vjp(x.reshape(x.shape[0], -1), cot, is_features=False)
    return g.reshape(x.shape)


def dchidfeat(iso: Iso, feat, cot=None) -> np.ndarray:
    """dchidfeat(iso, feat) (src/utils/minimumpath.jl:9-13): gradient w.r.t. the features"""
    feat = np.asarray(feat)
    g = iso.engine.chi_vjp(feat.reshape(feat.shape[0], -1), cot, is_features=True)
    return g.reshape(feat.shape)


def addcoords_(iso: Iso, xs_new, ys_new):
    """addcoords!(iso, coords) (src/iso.jl:238, src/simulation.jl:183-185).  The reference propagates ``coords`` with
    its simulation to obtain ys; simulators stay on the host here, so the caller passes the propagated samples.
    Only the new block is uploaded."""
    xs_new, ys_new = np.asarray(xs_new), np.asarray(ys_new)
    iso.engine.append_data(xs_new, ys_new)
    xs, ys = iso.data.coords
    iso.data = SimulationData(iso.data.sim, (np.concatenate([xs, xs_new], axis=1), np.concatenate([ys, ys_new], axis=2)),
                              featurizer=iso.data.featurizer)


def cutoff_(iso: Iso, cutoff: int):
    """iso.data = iso.data[end-cutoff+1:end] (run_kde!, src/iso.jl:288-290)"""
    if len(iso.data) > cutoff:
        iso.engine.keep_last(cutoff)
        iso.data = iso.data[slice(len(iso.data) - cutoff, None)]


def propchis(iso: Iso) -> np.ndarray:
    """chi of every Koopman sample, (d, K, N): model(propfeatures(data)) as used by resample_kde / chistratcoords
    (src/simulation.jl:199-207,227-228)"""
    return iso.engine.chis_prop()


def validationloss(iso: Iso, valdata: SimulationData) -> float:
    """validationloss(iso, valdata) (src/iso.jl:160-168): mean squared difference between chi on the validation
    start points and the shift-scaled Koopman expectation, the shift-scale being estimated on validation and
    training Koopman values together (isokann_validationloss)"""
    vx, vy = valdata.coords
    return iso.engine.validationloss(vx, vy)     # one library call, only the scalar comes back


def rates(iso: Iso, lagtime: float = 1.0) -> np.ndarray:
    """rates(iso) (src/iso.jl:339-343): the coarse-grained rate matrix Q with K chi = exp(tau Q) chi; for one dimensional
    chi the rates of chi and 1 - chi.  lagtime = lagtime(iso.data.sim) (the host simulation object knows it).
    One library call (isokann_rates); chi and K chi never leave the device."""
    return iso.engine.rates() / float(lagtime)


def residual_subspace(iso: Iso, v_norms: bool = False, want_res: bool = False):
    """residual_subspace(iso) (src/isotarget.jl:805-821): how badly each K chi_j is represented in span(chi).
    Returns (res, relres) like the reference's named tuple; res only on request (it is N x d)."""
    return iso.engine.residual_subspace(v_norms, want_res)


def residual_ritz(iso: Iso, want_residues: bool = False) -> dict:
    """residual_ritz(iso) (src/isotarget.jl:787-802): Ritz values/vectors of the Koopman operator on span(chi) and
    their relative residuals (isokann_residual_ritz)"""
    return iso.engine.residual_ritz(want_residues)


def koopman(iso: Iso) -> np.ndarray:
    """koopman(iso) = expectation(model, propfeatures(data)) (src/isotarget.jl:20)"""
    return iso.engine.koopman()


def chi_kchi(iso: Iso):
    """src/isotarget.jl:22-23"""
    return chis(iso), koopman(iso)


def cpu(iso: Iso) -> Iso:
    """cpu(iso) (src/iso.jl:257): refresh the host-side parameter copy; the object stays usable"""
    iso.sync_model()
    return iso


def save(path: str, iso: Iso):
    """save(path, iso) (src/iso.jl:405-408): parameters, optimiser state, losses, coordinates"""
    m, v, bt = iso.engine.download_opt_state()
    xs, ys = iso.data.coords
    np.savez(path, flat=iso.engine.download_params(), widths=np.array(iso.model.widths),
             layernorm=iso.model.layernorm, opt_m=m, opt_v=v if v is not None else np.zeros(0),
             beta_t=bt if bt is not None else np.zeros(0), losses=np.array(iso.losses), xs=xs, ys=ys)


def load_state(path: str, iso: Iso) -> Iso:
    """restore parameters / optimiser state / losses saved by ``save`` into a compatible Iso"""
    z = np.load(path if path.endswith(".npz") else path + ".npz")
    iso.engine.upload_params(z["flat"])
    iso.model.load_flat(z["flat"])
    if iso.engine.kind == "adam":
        iso.engine.upload_opt_state(z["opt_m"], z["opt_v"], z["beta_t"])
    else:
        iso.engine.upload_opt_state(z["opt_m"])
    iso.losses = [float(x) for x in z["losses"]]
    return iso


__all__ = ["Iso", "validationloss", "rates", "residual_subspace", "residual_ritz", "dchidx", "dchidfeat", "addcoords_", "cutoff_", "propchis", "run_", "train_batch_", "isotarget", "chis", "chicoords", "koopman", "chi_kchi", "cpu", "save",
           "load_state", "defaultmodel", "draw_perm", "DomainError", "TransformShiftscale", "TransformISA",
           "TransformPseudoInv"]
