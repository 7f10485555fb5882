"""numpy restatement of ISOKANN.jl's per-iteration hot path (test oracle / CPU baseline).

TEST INFRASTRUCTURE, PARITY UNPINNED -- see oracle/__init__.py.

Array convention ("records" layout).  Julia arrays are column-major; every
array here is the C-ordered numpy array with the *same memory image*:

    Julia xs[D,N]        <->  numpy (N, D)
    Julia ys[D,K,N]      <->  numpy (N, K, D)
    Julia feats[F,N]     <->  numpy (N, F)         feats[F,K,N] <-> (N, K, F)
    Julia chi[d,N]       <->  numpy (N, d)
    Julia W[out,in]      <->  numpy (in, out)      (Flux.Dense weight)

so ``W_julia * x_julia`` is ``x @ W`` here.  All citations are file:line under
the reference checkout (/root/reference/).
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

__all__ = [
    "halfinds", "pair_table", "flatpairdists", "flatpairdists_gram", "pdists",
    "Model", "pairnet_layers", "pairnet", "densenet", "smallnet", "init_params",
    "flatten_params", "unflatten_params", "num_params", "layernorm", "sigmoid",
    "forward", "forward_cache", "expectation", "shiftscale", "DomainError",
    "indexmap", "myisa", "fixperm", "isotarget_shiftscale", "isotarget_isa",
    "isotarget_pinv", "isotarget", "isa_from_chi", "pinv_from_chi", "OptConfig", "OptState", "opt_init",
    "opt_update", "loss_weights", "batch_loss_and_grad", "train_batch",
    "run", "weighted_expectation", "chi_vjp", "xoshiro256pp_next", "julia_randperm",
    "rates", "residual_subspace", "residual_ritz",
]

F32 = np.float32
F64 = np.float64


class DomainError(ValueError):
    """Mirror of Julia's DomainError as thrown on the hot path
    (src/iso.jl:186-189, src/isotarget.jl:39,94-97,159-163)."""

    def __init__(self, code: int, msg: str):
        super().__init__(msg)
        self.code = code


# ----------------------------------------------------------------------------------------------
# featurizer  (src/utils/pairdists.jl:6-24, 32-35, 50-56, 109-127; cast at src/simulation.jl:112)
# ----------------------------------------------------------------------------------------------

def halfinds(n: int) -> List[Tuple[int, int]]:
    """1-based (i, j), i<j, in the order of ``findall`` over a column-major
    strict upper triangle (src/utils/pairdists.jl:50-56): (1,2),(1,3),(2,3),(1,4)..."""
    return [(i, j) for j in range(2, n + 1) for i in range(1, j)]


def pair_table(n_atoms: int, cols: Optional[Sequence[int]] = None) -> np.ndarray:
    """0-based atom index pairs, shape (F, 2), for flatpairdists(x, cols)
    (src/utils/pairdists.jl:13-18).  ``cols`` are 1-based atom indices
    (FeaturesAtoms, src/utils/features.jl:26-29)."""
    if cols is None:
        cols = list(range(1, n_atoms + 1))
    cols = list(cols)
    h = halfinds(len(cols))
    return np.array([(cols[i - 1] - 1, cols[j - 1] - 1) for i, j in h], dtype=np.int32).reshape(-1, 2)


_CLIB = None


def _clib():
    """oracle/liboracle.so (oracle/isokann_oracle.c, built by oracle/build.py): the same featurizer
    arithmetic in C, used for inputs too large for the numpy gathers; None if it has not been built."""
    global _CLIB
    if _CLIB is None:
        import ctypes
        from pathlib import Path
        p = Path(__file__).resolve().parent / "liboracle.so"
        _CLIB = ctypes.CDLL(str(p)) if p.exists() else False
    return _CLIB or None


def _dists_from_pairs_c(x: np.ndarray, pairs0: np.ndarray, out_dtype) -> Optional[np.ndarray]:
    import ctypes
    lib = _clib()
    if lib is None or x.dtype not in (F32, F64) or np.dtype(out_dtype) not in (np.dtype(F32), np.dtype(F64)):
        return None
    d = x.shape[-1]
    xc = np.ascontiguousarray(x).reshape(-1, d)
    p0 = np.ascontiguousarray(pairs0, dtype=np.int32).reshape(-1, 2)
    out = np.empty((xc.shape[0], p0.shape[0]), dtype=out_dtype)
    fn = getattr(lib, "oracle_pdists_%s_%s" % ("f32" if xc.dtype == F32 else "f64",
                                               "f32" if out.dtype == F32 else "f64"))
    fn.restype = None
    fn(ctypes.c_void_p(xc.ctypes.data), ctypes.c_int64(xc.shape[0]), ctypes.c_int64(d),
       ctypes.c_void_p(p0.ctypes.data), ctypes.c_int(p0.shape[0]), ctypes.c_void_p(out.ctypes.data))
    return out.reshape(*x.shape[:-1], p0.shape[0])


def _dists_from_pairs(x: np.ndarray, pairs0: np.ndarray, out_dtype, use_c: Optional[bool] = None) -> np.ndarray:
    d = x.shape[-1]
    lead = x.shape[:-1]
    x = np.asarray(x)
    if use_c is None:
        use_c = x.size >= (1 << 16)
    if use_c:
        r = _dists_from_pairs_c(x, pairs0, out_dtype)
        if r is not None:
            return r
    c = np.asarray(x, dtype=F64).reshape(-1, d // 3, 3)
    out = np.empty((c.shape[0], len(pairs0)), dtype=out_dtype)
    step = max(1, (1 << 22) // max(1, len(pairs0)))
    for s in range(0, c.shape[0], step):
        blk = c[s:s + step]
        diff = blk[:, pairs0[:, 0], :] - blk[:, pairs0[:, 1], :]
        sq = np.einsum("mfc,mfc->mf", diff, diff)
        out[s:s + step] = np.sqrt(np.maximum(sq, 0.0)).astype(out_dtype)
    return out.reshape(*lead, len(pairs0))


def flatpairdists(x: np.ndarray, cols: Optional[Sequence[int]] = None, out_dtype=F32) -> np.ndarray:
    """flatpairdists (src/utils/pairdists.jl:6-24) in the oracle rule of SURVEY
    section 8(a1): direct differences in float64, ``max(.,0)``, ``sqrt``, then the
    ``Float32.`` cast of src/simulation.jl:112.  x: (..., D) -> (..., F)."""
    return _dists_from_pairs(x, pair_table(x.shape[-1] // 3, cols), out_dtype)


def flatpairdists_gram(x: np.ndarray, cols: Optional[Sequence[int]] = None) -> np.ndarray:
    """The reference's literal CPU formulation in the *input* eltype: Gram trick
    ``-2 x'x + |xi|^2 + |xj|^2`` (src/utils/pairdists.jl:32-35), gather of the
    strict upper triangle (:19-20), clamp (:21), sqrt (:22)."""
    d = x.shape[-1]
    lead = x.shape[:-1]
    c = x.reshape(-1, d // 3, 3)
    if cols is not None:
        c = c[:, np.asarray(cols) - 1, :]
    n = c.shape[1]
    g = np.einsum("mic,mjc->mij", c, c)
    sq = np.einsum("mic,mic->mi", c, c)
    p = -2 * g + sq[:, :, None] + sq[:, None, :]
    h = halfinds(n)
    ii = np.array([a - 1 for a, _ in h]); jj = np.array([b - 1 for _, b in h])
    p = p[:, ii, jj]
    p = np.sqrt(np.maximum(p, 0))
    return p.reshape(*lead, len(h))


def pdists(x: np.ndarray, pairs: Sequence[Tuple[int, int]], out_dtype=F32) -> np.ndarray:
    """pdists(coords, pairs) (src/utils/pairdists.jl:109-127): explicit 1-based
    pair list, direct differences, feature order = order of ``pairs``."""
    p0 = np.asarray(pairs, dtype=np.int64).reshape(-1, 2) - 1
    return _dists_from_pairs(x, p0, out_dtype)


# ----------------------------------------------------------------------------------------------
# model  (src/models.jl:65-69, 87-92, 102-108; Flux 0.16.9 Dense / LayerNorm semantics)
# ----------------------------------------------------------------------------------------------

ACT_IDENTITY, ACT_SIGMOID, ACT_TANH, ACT_RELU = 0, 1, 2, 3


@dataclass
class Model:
    """Flux.Chain([LayerNorm(F)]; Dense(.., act)...; Dense(.., lastact)) (src/models.jl:87-92)."""
    widths: List[int]                 # [F, h1, ..., d]
    layernorm: bool = True
    act: int = ACT_SIGMOID
    lastact: int = ACT_IDENTITY
    ln_eps: float = 1e-5
    ln_scale: Optional[np.ndarray] = None   # (F,)
    ln_bias: Optional[np.ndarray] = None    # (F,)
    W: List[np.ndarray] = field(default_factory=list)   # (in, out) each
    b: List[np.ndarray] = field(default_factory=list)

    @property
    def nlayers(self) -> int:
        return len(self.widths) - 1

    @property
    def nout(self) -> int:
        return self.widths[-1]

    def copy(self) -> "Model":
        return Model(list(self.widths), self.layernorm, self.act, self.lastact, self.ln_eps,
                     None if self.ln_scale is None else self.ln_scale.copy(),
                     None if self.ln_bias is None else self.ln_bias.copy(),
                     [w.copy() for w in self.W], [b.copy() for b in self.b])


def _julia_round(x: float) -> int:
    """Julia ``round(Int, x)`` rounds half to even, like Python's round()."""
    return int(round(x))


def pairnet_layers(n: int, layers: int = 3, nout: int = 1) -> List[int]:
    """Layer-size rule of pairnet (src/models.jl:66-67)."""
    ws = [_julia_round(n ** (l / layers)) for l in range(layers, 0, -1)]
    return ws + [nout]


def init_params(m: Model, rng: np.random.Generator) -> Model:
    """Flux defaults: glorot_uniform weights U(+-sqrt(6/(in+out))), zero bias,
    LayerNorm scale 1 / bias 0.  (The draw itself is numpy's, not Julia's.)"""
    m.W, m.b = [], []
    for i in range(m.nlayers):
        fin, fout = m.widths[i], m.widths[i + 1]
        lim = np.sqrt(6.0 / (fin + fout))
        m.W.append(rng.uniform(-lim, lim, size=(fin, fout)).astype(F32))
        m.b.append(np.zeros(fout, dtype=F32))
    if m.layernorm:
        m.ln_scale = np.ones(m.widths[0], dtype=F32)
        m.ln_bias = np.zeros(m.widths[0], dtype=F32)
    return m


def densenet(layers: Sequence[int], layernorm: bool = False, act=ACT_SIGMOID, lastact=ACT_IDENTITY,
             rng: Optional[np.random.Generator] = None) -> Model:
    m = Model(list(layers), layernorm, act, lastact)
    return init_params(m, rng if rng is not None else np.random.default_rng(0))


def pairnet(n: int, layers: int = 3, nout: int = 1, layernorm: bool = True, act=ACT_SIGMOID,
            lastact=ACT_IDENTITY, rng: Optional[np.random.Generator] = None) -> Model:
    return densenet(pairnet_layers(n, layers, nout), layernorm, act, lastact, rng)


def smallnet(nin: int, nout: int = 1, rng: Optional[np.random.Generator] = None) -> Model:
    """src/models.jl:102-108 (no LayerNorm)."""
    return densenet([nin, 8, 8, 8, nout], False, rng=rng)


def num_params(m: Model) -> int:
    p = 2 * m.widths[0] if m.layernorm else 0
    for i in range(m.nlayers):
        p += m.widths[i] * m.widths[i + 1] + m.widths[i + 1]
    return p


def flatten_params(m: Model) -> np.ndarray:
    """Flat order of the C ABI = Functors traversal of the Chain: [gamma, beta,] W1, b1, W2, b2...
    with W in Julia memory order (column-major out x in == C-order (in, out))."""
    parts = []
    if m.layernorm:
        parts += [m.ln_scale, m.ln_bias]
    for w, b in zip(m.W, m.b):
        parts += [w.ravel(), b]
    return np.concatenate([np.asarray(p, dtype=F32).ravel() for p in parts])


def unflatten_params(m: Model, flat: np.ndarray) -> Model:
    flat = np.asarray(flat, dtype=F32)
    o = 0
    if m.layernorm:
        n = m.widths[0]
        m.ln_scale = flat[o:o + n].copy(); o += n
        m.ln_bias = flat[o:o + n].copy(); o += n
    m.W, m.b = [], []
    for i in range(m.nlayers):
        fin, fout = m.widths[i], m.widths[i + 1]
        m.W.append(flat[o:o + fin * fout].reshape(fin, fout).copy()); o += fin * fout
        m.b.append(flat[o:o + fout].copy()); o += fout
    assert o == flat.size
    return m


def sigmoid(x: np.ndarray) -> np.ndarray:
    """NNlib sigmoid_fast: t=exp(-|x|); x>=0 ? 1/(1+t) : t/(1+t)  (stable form)."""
    t = np.exp(-np.abs(x))
    return np.where(x >= 0, 1 / (1 + t), t / (1 + t)).astype(x.dtype)


def _act(a: np.ndarray, kind: int) -> np.ndarray:
    if kind == ACT_IDENTITY:
        return a
    if kind == ACT_SIGMOID:
        return sigmoid(a)
    if kind == ACT_TANH:
        return np.tanh(a)
    if kind == ACT_RELU:
        return np.maximum(a, 0)
    raise ValueError(kind)


def _dact_from_out(z: np.ndarray, kind: int) -> np.ndarray:
    if kind == ACT_IDENTITY:
        return np.ones_like(z)
    if kind == ACT_SIGMOID:
        return z * (1 - z)
    if kind == ACT_TANH:
        return 1 - z * z
    if kind == ACT_RELU:
        return (z > 0).astype(z.dtype)
    raise ValueError(kind)


def layernorm(x: np.ndarray, eps: float = 1e-5) -> np.ndarray:
    """Flux.normalise(x; dims=1, eps): (x-mu)/sqrt(var_biased + eps^2), per sample
    (Flux 0.16 LayerNorm, used at src/models.jl:90)."""
    mu = x.mean(axis=-1, keepdims=True, dtype=x.dtype)
    xc = x - mu
    var = (xc * xc).mean(axis=-1, keepdims=True, dtype=x.dtype)
    return xc / np.sqrt(var + x.dtype.type(eps) ** 2)


def forward_cache(m: Model, feats: np.ndarray):
    """model(x) keeping what the backward pass needs.  feats: (M, F) -> chi (M, d)."""
    dt = feats.dtype
    cache = {}
    z = feats
    if m.layernorm:
        xh = layernorm(z, m.ln_eps)
        cache["xhat"] = xh
        z = xh * m.ln_scale.astype(dt) + m.ln_bias.astype(dt)
    zs = [z]
    for i in range(m.nlayers):
        a = z @ m.W[i].astype(dt) + m.b[i].astype(dt)
        z = _act(a, m.act if i < m.nlayers - 1 else m.lastact)
        zs.append(z)
    cache["zs"] = zs
    return z, cache


def forward(m: Model, feats: np.ndarray, chunk: int = 1 << 16) -> np.ndarray:
    """model(feats) for feats (..., F) -> (..., d); Dense flattens leading dims
    (Flux Dense on N-D input), LayerNorm normalises the first Julia dim (= last here)."""
    lead = feats.shape[:-1]
    x = feats.reshape(-1, feats.shape[-1])
    out = np.empty((x.shape[0], m.nout), dtype=feats.dtype)
    for s in range(0, x.shape[0], chunk):
        out[s:s + chunk] = forward_cache(m, x[s:s + chunk])[0]
    return out.reshape(*lead, m.nout)


# ----------------------------------------------------------------------------------------------
# Koopman expectation and targets  (src/isotarget.jl:18, 32-42, 74-127, 145-179)
# ----------------------------------------------------------------------------------------------

def expectation(m: Model, ysf: np.ndarray) -> np.ndarray:
    """expectation(f, ys) = dropdims(sum(f(ys); dims=2); dims=2) ./ K  (src/isotarget.jl:18).
    ysf: (N, K, F) -> (N, d).  The K-sum runs in index order in the array eltype."""
    chi = forward(m, ysf)                      # (N, K, d)
    k = chi.shape[1]
    acc = chi[:, 0, :].copy()
    for j in range(1, k):
        acc += chi[:, j, :]
    return acc / chi.dtype.type(k)


def weighted_expectation(m: Model, ysf: np.ndarray, weights: np.ndarray) -> np.ndarray:
    """expectation(f, gs::WeightedSamples) (src/data.jl:215): sum(f(values).*weights; dims=2)/K.
    weights: (N, K) or (N, K, 1)."""
    chi = forward(m, ysf)
    w = weights.reshape(chi.shape[0], chi.shape[1], 1).astype(chi.dtype)
    prod = chi * w
    acc = prod[:, 0, :].copy()
    for j in range(1, prod.shape[1]):
        acc += prod[:, j, :]
    return acc / chi.dtype.type(prod.shape[1])


def shiftscale(ks: np.ndarray) -> np.ndarray:
    """src/isotarget.jl:36-42."""
    if not (ks.ndim == 1 or ks.shape[-1] == 1):
        raise AssertionError("TransformShiftscale only works with one dimensional chi functions")
    lo, hi = ks.min(), ks.max()
    if not hi > lo:
        raise DomainError(1, "Could not compute the shift-scale. chi function is constant")
    return (ks - lo) / (hi - lo)


def indexmap(X: np.ndarray) -> List[int]:
    """PCCAPlus.jl 1.1.2 ``indexmap`` (inner simplex algorithm), restated from the
    published algorithm (source not vendored -> assumption).
    X: (N, d) float64.  Returns 0-based row indices."""
    X = np.array(X, dtype=F64)
    d = X.shape[1]
    ind = []
    for j in range(d):
        rownorm = np.sqrt(np.sum(X * X, axis=1))
        i = int(np.argmax(rownorm))           # first maximum, like Julia argmax
        ind.append(i)
        if j == 0:
            X = X - X[i, :][None, :]
        else:
            X = X / rownorm[i]
            v = X[i, :].copy()
            X = X - np.outer(X @ v, v)
    return ind


def _inv_sqrt_sym(C: np.ndarray) -> np.ndarray:
    w, V = np.linalg.eigh(C)
    return (V * (w ** -0.5)) @ V.T


def myisa(X: np.ndarray, whitening: bool = False) -> np.ndarray:
    """src/isotarget.jl:81-98.  X = ks' (N, d) -> inv(X[i,:]) (d, d), float64."""
    X = np.asarray(X, dtype=F64)
    try:
        if whitening:
            C = (X.T @ X) / X.shape[0]
            i = indexmap(X @ _inv_sqrt_sym(C))
        else:
            i = indexmap(X)
        S = X[i, :]
        if not np.all(np.isfinite(S)):
            raise np.linalg.LinAlgError("non-finite")
        return np.linalg.inv(S)
    except np.linalg.LinAlgError:
        raise DomainError(3, "Could not compute the simplex transformation. The subspace might be singular/collapsed")


def fixperm(new: np.ndarray, old: np.ndarray) -> np.ndarray:
    """src/isotarget.jl:120-127 in records layout: new, old are (N, d); permute the d
    components of ``new`` to minimise the entry-wise 1-norm to ``old``; first minimum over
    lexicographic ``Combinatorics.permutations(1:d)``."""
    d = new.shape[1]
    best, bestp = None, None
    for p in itertools.permutations(range(d)):
        c = np.abs(new[:, list(p)].astype(F64) - old.astype(F64)).sum()
        if best is None or c < best:
            best, bestp = c, p
    return new[:, list(bestp)]


def isotarget_shiftscale(m: Model, xsf: np.ndarray, ysf: np.ndarray) -> np.ndarray:
    """src/isotarget.jl:34."""
    return shiftscale(expectation(m, ysf))


def isa_from_chi(chi: np.ndarray, ks: np.ndarray, permute: bool = True, whitening: bool = False) -> np.ndarray:
    """body of src/isotarget.jl:100-107 given chi = model(xs) and ks = expectation(model, ys), both (N, d)."""
    assert chi.shape[1] > 1, "TransformISA does not work with one dimensional chi functions"
    A = myisa(ks, whitening)                              # (d, d) float64 = inv(X[i,:])
    # Julia: target = A' * ks  ([d,d] x [d,N]); records layout: ks_rec @ A
    target = ks.astype(F64) @ A
    if permute:
        target = fixperm(target, chi)
    return target.astype(chi.dtype)


def isotarget_isa(m: Model, xsf: np.ndarray, ysf: np.ndarray, permute: bool = True,
                  whitening: bool = False) -> np.ndarray:
    """src/isotarget.jl:100-107.  Returns (N, d) float32."""
    return isa_from_chi(forward(m, xsf), expectation(m, ysf), permute, whitening).astype(xsf.dtype)


def pinv_from_chi(chi_r: np.ndarray, kchi_r: np.ndarray, normalize: bool = True, direct: bool = True,
                  eigenvecs: bool = True, permute: bool = True, details: Optional[dict] = None) -> np.ndarray:
    """body of src/isotarget.jl:152-179 given chi = model(xs) and kchi = expectation(model, ys), both (N, d)
    (float32 LAPACK via scipy: pinv = gesdd SVD with rtol = eps*min(d,N); schur = sgees, unsorted).
    ``details`` receives Kinv (or K) and the Schur vectors T, Julia-shaped (d, d)."""
    import scipy.linalg as sla
    assert chi_r.shape[1] > 1, "TransformPseudoInv does not work with one dimensional chi functions"
    chi = np.ascontiguousarray(chi_r.T)                   # Julia-shaped [d, N]
    kchi = np.ascontiguousarray(kchi_r.T)
    d, n = kchi.shape
    try:
        if not np.all(np.isfinite(kchi)):
            raise np.linalg.LinAlgError("non-finite")
        kchi_inv = sla.pinv(kchi, rtol=float(np.finfo(kchi.dtype).eps) * min(d, n))
    except (np.linalg.LinAlgError, ValueError):
        raise DomainError(4, "Could not compute the pseudoinverse. The subspace might be singular/collapsed")
    if direct:
        Kinv = chi @ kchi_inv
        T = sla.schur(Kinv, output="real")[1] if eigenvecs else np.eye(d, dtype=kchi.dtype)
        target = (T @ Kinv) @ kchi
    else:
        Kinv = kchi @ kchi_inv
        T = sla.schur(Kinv, output="real")[1] if eigenvecs else np.eye(d, dtype=kchi.dtype)
        target = (T @ np.linalg.inv(Kinv)) @ kchi
    if details is not None:
        details["Kinv"] = Kinv
        details["T"] = T
    target = target.astype(kchi.dtype)
    if normalize:
        l1 = np.abs(target).sum(axis=1, keepdims=True, dtype=kchi.dtype)
        target = target / l1 * kchi.dtype.type(n)
    t_r = np.ascontiguousarray(target.T)
    if permute:
        t_r = fixperm(t_r, chi_r)
    return t_r.astype(chi_r.dtype)


def isotarget_pinv(m: Model, xsf: np.ndarray, ysf: np.ndarray, normalize: bool = True,
                   direct: bool = True, eigenvecs: bool = True, permute: bool = True) -> np.ndarray:
    """src/isotarget.jl:152-179.  Returns (N, d) float32."""
    return pinv_from_chi(forward(m, xsf), expectation(m, ysf), normalize, direct, eigenvecs, permute).astype(xsf.dtype)


def isotarget(kind: str, m: Model, xsf: np.ndarray, ysf: np.ndarray, **kw) -> np.ndarray:
    """isotarget(target, model, xs, ys) dispatch (src/isotarget.jl:10-12)."""
    if kind == "shiftscale":
        return isotarget_shiftscale(m, xsf, ysf)
    if kind == "isa":
        return isotarget_isa(m, xsf, ysf, **kw)
    if kind == "pinv":
        return isotarget_pinv(m, xsf, ysf, **kw)
    raise ValueError(kind)


# ----------------------------------------------------------------------------------------------
# optimiser  (src/models.jl:4,12,20; Optimisers.jl 0.4.7 WeightDecay / Adam / Nesterov / OptimiserChain)
# ----------------------------------------------------------------------------------------------

@dataclass
class OptConfig:
    kind: str = "nesterov"        # default NesterovRegularized() (src/iso.jl:18)
    eta: float = 1e-3
    lam: float = 1e-4
    beta1: float = 0.9
    beta2: float = 0.999
    eps: float = 1e-8
    rho: float = 0.9


@dataclass
class OptState:
    m: Optional[np.ndarray] = None       # Adam first moment / Nesterov velocity
    v: Optional[np.ndarray] = None       # Adam second moment
    beta_t: Optional[np.ndarray] = None  # Adam running (beta1^t, beta2^t), starts at (beta1, beta2)


def opt_init(cfg: OptConfig, nparams: int) -> OptState:
    if cfg.kind == "adam":
        return OptState(np.zeros(nparams, F32), np.zeros(nparams, F32),
                        np.array([cfg.beta1, cfg.beta2], dtype=F32))
    return OptState(np.zeros(nparams, F32))


def opt_update(cfg: OptConfig, st: OptState, theta: np.ndarray, g: np.ndarray) -> np.ndarray:
    """One ``Optimisers.update!`` on the flat float32 parameter vector; returns new theta.
    All arithmetic in float32 with hyper-parameters cast to float32 (``T(o.eta)`` ...)."""
    theta = theta.astype(F32); g = g.astype(F32)
    lam = F32(cfg.lam); eta = F32(cfg.eta)
    g = g + lam * theta                                         # WeightDecay
    if cfg.kind == "adam":
        b1, b2, eps = F32(cfg.beta1), F32(cfg.beta2), F32(cfg.eps)
        st.m = b1 * st.m + (F32(1) - b1) * g
        st.v = b2 * st.v + (F32(1) - b2) * (g * g)
        bt1, bt2 = st.beta_t
        dx = st.m / (F32(1) - bt1) / (np.sqrt(st.v / (F32(1) - bt2)) + eps) * eta
        st.beta_t = (st.beta_t * np.array([b1, b2], dtype=F32)).astype(F32)
    elif cfg.kind == "nesterov":
        rho = F32(cfg.rho)
        dx = -(rho * rho) * st.m + (F32(1) + rho) * eta * g
        st.m = rho * st.m - eta * g
    else:
        raise ValueError(cfg.kind)
    return (theta - dx).astype(F32)


# ----------------------------------------------------------------------------------------------
# training epoch  (src/iso.jl:179-194; MLUtils DataLoader(shuffle, partial=false); Zygote backward)
# ----------------------------------------------------------------------------------------------

def loss_weights(target: np.ndarray):
    """w of src/iso.jl:183: 1 ./ std(target, dims=2) (corrected) for d>1, Float64 1.0 for d==1."""
    if target.shape[1] > 1:
        return (F32(1) / target.std(axis=0, ddof=1, dtype=F64).astype(F32)).astype(F32)
    return None


def batch_loss_and_grad(m: Model, x: np.ndarray, y: np.ndarray, w):
    """loss l = sum(abs2, (m(x)-y).*w) and the gradient of l/B w.r.t. the flat params
    (src/iso.jl:184-191)."""
    B = x.shape[0]
    chi, cache = forward_cache(m, x)
    r = chi - y                                                   # float32
    if w is None:                                                 # d==1: Float64 promotion (w = 1.0)
        l = float(np.sum(r.astype(F64) ** 2))
        delta = (2.0 * r.astype(F64) / B).astype(F32)
    else:
        z = r * w[None, :]
        l = float(np.sum(z * z, dtype=F32))
        delta = ((F32(2) * z * F32(1.0 / B)) * w[None, :]).astype(F32)
    zs = cache["zs"]
    gW, gb = [None] * m.nlayers, [None] * m.nlayers
    for i in range(m.nlayers - 1, -1, -1):
        kind = m.act if i < m.nlayers - 1 else m.lastact
        delta = delta * _dact_from_out(zs[i + 1], kind)
        gW[i] = zs[i].T @ delta
        gb[i] = delta.sum(axis=0)
        delta = delta @ m.W[i].T
    parts = []
    if m.layernorm:
        parts += [(delta * cache["xhat"]).sum(axis=0), delta.sum(axis=0)]
    for a, b in zip(gW, gb):
        parts += [a.ravel(), b]
    return l, np.concatenate([p.astype(F32).ravel() for p in parts])


def train_batch(m: Model, xsf: np.ndarray, target: np.ndarray, cfg: OptConfig, st: OptState,
                minibatch: int, perm1: np.ndarray, partial: bool = False) -> float:
    """train_batch! (src/iso.jl:179-194).  ``perm1`` is the 1-based permutation the
    DataLoader would draw (one ``randperm`` per epoch).  Mutates m and st; returns ls/N."""
    n = xsf.shape[0]
    bs = n if (minibatch == 0 or n < minibatch) else minibatch
    w = loss_weights(target)
    ls = 0.0
    nb = n // bs if not partial else -(-n // bs)
    for i in range(nb):
        idx = np.asarray(perm1[i * bs:(i + 1) * bs], dtype=np.int64) - 1
        l, g = batch_loss_and_grad(m, xsf[idx], target[idx], w)
        if not np.isfinite(l):
            raise DomainError(2, "The ISOKANN model collapsed under training. Try reducing the learning rate or increasing regularization")
        ls += l
        # the loss handed to Flux.train! is l / numobs(x): the gradient above is of l/len(idx)
        unflatten_params(m, opt_update(cfg, st, flatten_params(m), g))
    return ls / n


def run(m: Model, xsf: np.ndarray, ysf: np.ndarray, cfg: OptConfig, st: OptState, n_iter: int,
        minibatch: int, perms1: Sequence[np.ndarray], target_kind: str = "shiftscale",
        epochs: int = 1, **target_kw) -> List[float]:
    """run!(iso, n, epochs) (src/iso.jl:72-94) on cached float32 features."""
    losses = []
    p = 0
    for _ in range(n_iter):
        t = isotarget(target_kind, m, xsf, ysf, **target_kw)
        for _e in range(epochs):
            losses.append(train_batch(m, xsf, t, cfg, st, minibatch, perms1[p]))
            p += 1
    return losses


# ----------------------------------------------------------------------------------------------
# minibatch order: Julia's randperm(rng::Xoshiro, n)  (Random stdlib of Julia 1.12; consumed by Flux.DataLoader at
# src/iso.jl:181 through MLUtils.shuffleobs).  UNPINNED: restated from the published algorithm,
# no golden vector from a Julia session is available here.  Pure-Python loops: small n only.
# ----------------------------------------------------------------------------------------------

_M64 = (1 << 64) - 1


def xoshiro256pp_next(st: List[int]) -> int:
    """one draw of Xoshiro256++ (Blackman & Vigna); st = [s0, s1, s2, s3] is advanced in place"""
    def rotl(x, k):
        return ((x << k) | (x >> (64 - k))) & _M64
    s0, s1, s2, s3 = st
    res = (rotl((s0 + s3) & _M64, 23) + s0) & _M64
    t = (s1 << 17) & _M64
    s2 ^= s0
    s3 ^= s1
    s1 ^= s2
    s0 ^= s3
    s2 ^= t
    s3 = rotl(s3, 45)
    st[:] = [s0, s1, s2, s3]
    return res


def julia_randperm(state4: Sequence[int], n: int):
    """randperm!(rng, a) of Random: a[1] = 1; mask = 3; for i in 2:n: j = 1 + rand(rng, ltm52(i, mask));
    a[i] = a[j]; a[j] = i; i == 1 + mask && (mask = 2mask + 1), where ltm52 draws (rand(UInt64) >>> 12) & mask until
    the value is <= i - 1.  Returns (1-based permutation, advanced state)."""
    st = [int(x) & _M64 for x in state4]
    a = [0] * n
    if n > 0:
        a[0] = 1
    mask = 3
    for i in range(2, n + 1):
        while True:
            x = (xoshiro256pp_next(st) >> 12) & mask
            if x <= i - 1:
                break
        j = 1 + x
        if i != j:
            a[i - 1] = a[j - 1]
        a[j - 1] = i
        if i == 1 + mask:
            mask = 2 * mask + 1
    return np.array(a, dtype=np.int64), st


# ----------------------------------------------------------------------------------------------
# gradient of chi w.r.t. coordinates  (dchidx / dchidfeat, src/utils/minimumpath.jl:3-13; the pullback of
# the featurizer is src/utils/pairdists.jl:153-167,179-196) -- SURVEY section 8f, "next" row 1
# ----------------------------------------------------------------------------------------------

def chi_vjp(m: Model, x: np.ndarray, cot: Optional[np.ndarray] = None, pairs0: Optional[np.ndarray] = None):
    """Vector-Jacobian product d(sum(cot .* chi(x)))/dx in float64.
    x: (M, D) coordinate records (pairs0 = (F, 2) 0-based atom pairs) or (M, F) features (pairs0 None);
    cot: (M, d) cotangent (default ones, i.e. Zygote.gradient of `chicoords(iso, x) |> only` for d = 1)."""
    x = np.asarray(x, dtype=F64)
    M = x.shape[0]
    if pairs0 is not None:
        c = x.reshape(M, -1, 3)
        diff = c[:, pairs0[:, 0], :] - c[:, pairs0[:, 1], :]          # (M, F, 3)
        f = np.sqrt(np.maximum((diff * diff).sum(-1), 0.0))
    else:
        f = x
    z = f
    if m.layernorm:
        mu = f.mean(-1, keepdims=True)
        xc = f - mu
        r = 1.0 / np.sqrt((xc * xc).mean(-1, keepdims=True) + float(m.ln_eps) ** 2)
        xh = xc * r
        z = xh * m.ln_scale.astype(F64) + m.ln_bias.astype(F64)
    zs = [z]
    for i in range(m.nlayers):
        a = z @ m.W[i].astype(F64) + m.b[i].astype(F64)
        z = _act(a, m.act if i < m.nlayers - 1 else m.lastact)
        zs.append(z)
    g = np.ones_like(z) if cot is None else np.asarray(cot, dtype=F64)
    for i in range(m.nlayers - 1, -1, -1):
        g = g * _dact_from_out(zs[i + 1], m.act if i < m.nlayers - 1 else m.lastact)
        g = g @ m.W[i].astype(F64).T
    if m.layernorm:
        g = g * m.ln_scale.astype(F64)
        g = r * (g - g.mean(-1, keepdims=True) - xh * (g * xh).mean(-1, keepdims=True))
    if pairs0 is None:
        return g
    w = np.where(f > 0, g / np.where(f > 0, f, 1.0), 0.0)                 # dL/df / f
    contrib = w[:, :, None] * diff                                       # (M, F, 3): gradient w.r.t. atom a
    out = np.zeros_like(c)
    np.add.at(out, (np.arange(M)[:, None], pairs0[None, :, 0]), contrib)
    np.add.at(out, (np.arange(M)[:, None], pairs0[None, :, 1]), -contrib)
    return out.reshape(M, -1)


# ----------------------------------------------------------------------------------------------
# diagnostics ("next" row 4): rates (src/iso.jl:339-351), residual_subspace
# (src/isotarget.jl:805-821), residual_ritz (src/isotarget.jl:787-802).  chi, kchi: (N, d) records,
# i.e. already the V = chis(iso)' and KV = koopman(iso)' of the reference.
# ----------------------------------------------------------------------------------------------

def rates(chi: np.ndarray, kchi: np.ndarray, dtype=F64) -> np.ndarray:
    """rates(x, y) = log(y / x) with x = chi (d x N), y = Kchi, src/iso.jl:345-351 (no division by the lag time).
    `y / x` is Julia's right division = the least-squares solution M of M x = y (LAPACK QR), here numpy lstsq;
    `log` is the principal matrix logarithm (scipy.linalg.logm).  The reference runs this in Float32
    (dtype=np.float32 reproduces that up to LAPACK's rounding); returns the Julia-shaped (dim, dim) matrix."""
    import scipy.linalg
    x = np.asarray(chi, dtype=dtype).T                      # d x N
    y = np.asarray(kchi, dtype=dtype).T
    if x.shape[0] == 1:                                     # src/iso.jl:346-349
        x = np.concatenate([x, 1 - x], axis=0)
        y = np.concatenate([y, 1 - y], axis=0)
    m = np.linalg.lstsq(x.T, y.T, rcond=None)[0].T          # y / x
    return scipy.linalg.logm(m)


def _qr_thin(V: np.ndarray):
    return np.linalg.qr(V, mode="reduced")                  # qr_thin, src/isotarget.jl:823


def residual_subspace(chi: np.ndarray, kchi: np.ndarray, v_norms: bool = False):
    """src/isotarget.jl:810-821: res = KV - Q Q' KV, relres = column norms relative to KV (or V).  Float64."""
    V = np.asarray(chi, dtype=F64)
    KV = np.asarray(kchi, dtype=F64)
    Q, _ = _qr_thin(V)
    res = KV - Q @ (Q.T @ KV)
    den = np.linalg.norm(V, axis=0) if v_norms else np.linalg.norm(KV, axis=0)
    return res, np.linalg.norm(res, axis=0) / den


def residual_ritz(chi: np.ndarray, kchi: np.ndarray):
    """src/isotarget.jl:787-802: Q, R = qr(V); KQ = KV inv(R); Kr = Q' KQ; eigen(Kr, sortby = x -> abs(1 - x));
    residues = KQ vecs - vals' .* (Q vecs); relres = column norms relative to KQ vecs.
    Returns (residues, relres, vals, vecs, Q); vals/vecs complex when Kr has complex eigenvalues."""
    V = np.asarray(chi, dtype=F64)
    KV = np.asarray(kchi, dtype=F64)
    Q, R = _qr_thin(V)
    KQ = KV @ np.linalg.inv(R)
    Kr = Q.T @ KQ
    vals, vecs = np.linalg.eig(Kr)                          # LAPACK dgeev, like Julia's eigen
    order = np.argsort(np.abs(1 - vals), kind="stable")     # sortby (insertion sort for d <= 20: stable)
    vals, vecs = vals[order], vecs[:, order]
    residues = KQ @ vecs - vals[None, :] * (Q @ vecs)
    relres = np.linalg.norm(residues, axis=0) / np.linalg.norm(KQ @ vecs, axis=0)
    return residues, relres, vals, vecs, Q
