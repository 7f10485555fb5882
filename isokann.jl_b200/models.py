"""Host-side mirror of reference src/models.jl: network and optimiser *descriptions*.

The arithmetic lives in libisokann_b200.so; these classes only carry the shapes, the
hyper-parameters and a host copy of the parameters in Flux's layout (``W[out, in]``,
column-major), exactly what ``cpu(iso)`` / JLD2 hold in the reference.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np


@dataclass
class OptimiserRule:
    """OptimiserChain(WeightDecay(reg), rule) -- ``Regularized`` (src/models.jl:4)."""
    kind: str          # "adam" | "nesterov"
    eta: float = 1e-3
    reg: float = 1e-4
    beta1: float = 0.9
    beta2: float = 0.999
    eps: float = 1e-8
    rho: float = 0.9


def AdamRegularized(adam: float = 1e-3, reg: float = 1e-4) -> OptimiserRule:
    """src/models.jl:12"""
    return OptimiserRule("adam", eta=adam, reg=reg)


def NesterovRegularized(lr: float = 1e-3, reg: float = 1e-4) -> OptimiserRule:
    """src/models.jl:20 (the default optimiser of Iso, src/iso.jl:18)"""
    return OptimiserRule("nesterov", eta=lr, reg=reg)


@dataclass
class Chain:
    """Flux.Chain([LayerNorm(F)], Dense(F=>h1, act), ..., Dense(h=>d, lastact)) (src/models.jl:87-92).

    ``weights[l]`` is Julia-shaped ``(out, in)`` Fortran-ordered float32, ``biases[l]`` is ``(out,)``.
    """
    widths: List[int]
    layernorm: bool = False
    activation: str = "sigmoid"
    lastactivation: str = "identity"
    ln_eps: float = 1e-5
    ln_scale: Optional[np.ndarray] = None
    ln_bias: Optional[np.ndarray] = None
    weights: List[np.ndarray] = field(default_factory=list)
    biases: List[np.ndarray] = field(default_factory=list)

    @property
    def layers(self):  # reference host code pokes at model.layers (src/iso.jl:261, src/models.jl:26-31)
        out = []
        if self.layernorm:
            out.append(("LayerNorm", self.widths[0]))
        for i in range(len(self.widths) - 1):
            last = i == len(self.widths) - 2
            out.append(("Dense", self.widths[i], self.widths[i + 1], self.lastactivation if last else self.activation))
        return out

    def num_params(self) -> int:
        p = 2 * self.widths[0] if self.layernorm else 0
        for i in range(len(self.widths) - 1):
            p += self.widths[i] * self.widths[i + 1] + self.widths[i + 1]
        return p

    def flat(self) -> np.ndarray:
        """flat parameter vector in the C-ABI order (Functors traversal; W column-major)."""
        parts = []
        if self.layernorm:
            parts += [self.ln_scale, self.ln_bias]
        for w, b in zip(self.weights, self.biases):
            parts += [np.asarray(w, dtype=np.float32).ravel(order="F"), b]
        return np.ascontiguousarray(np.concatenate([np.asarray(p, np.float32).ravel() for p in parts]))

    def load_flat(self, flat: np.ndarray) -> "Chain":
        flat = np.asarray(flat, dtype=np.float32)
        o = 0
        if self.layernorm:
            n = self.widths[0]
            self.ln_scale = flat[o:o + n].copy(); o += n
            self.ln_bias = flat[o:o + n].copy(); o += n
        self.weights, self.biases = [], []
        for i in range(len(self.widths) - 1):
            fin, fout = self.widths[i], self.widths[i + 1]
            self.weights.append(flat[o:o + fin * fout].reshape((fout, fin), order="F").copy(order="F")); o += fin * fout
            self.biases.append(flat[o:o + fout].copy()); o += fout
        assert o == flat.size
        return self


def _glorot_init(m: Chain, rng: np.random.Generator) -> Chain:
    """Flux defaults: glorot_uniform weights, zero bias, LayerNorm scale 1 / bias 0."""
    m.weights, m.biases = [], []
    for i in range(len(m.widths) - 1):
        fin, fout = m.widths[i], m.widths[i + 1]
        lim = np.sqrt(6.0 / (fin + fout))
        m.weights.append(np.asfortranarray(rng.uniform(-lim, lim, size=(fout, fin)).astype(np.float32)))
        m.biases.append(np.zeros(fout, dtype=np.float32))
    if m.layernorm:
        m.ln_scale = np.ones(m.widths[0], dtype=np.float32)
        m.ln_bias = np.zeros(m.widths[0], dtype=np.float32)
    return m


def densenet(layers: Sequence[int], activation: str = "sigmoid", lastactivation: str = "identity",
             layernorm: bool = False, rng: Optional[np.random.Generator] = None) -> Chain:
    """src/models.jl:87-92"""
    m = Chain(list(layers), layernorm, activation, lastactivation)
    return _glorot_init(m, rng if rng is not None else np.random.default_rng())


def pairnet_layers(n: int, layers: int = 3, nout: int = 1) -> List[int]:
    """[round(Int, n^(l/layers)) for l in layers:-1:1]; push nout (src/models.jl:66-67);
    Julia rounds half to even, as Python's round() does."""
    return [int(round(n ** (l / layers))) for l in range(layers, 0, -1)] + [nout]


def pairnet(data=None, *, n: Optional[int] = None, layers: int = 3, activation: str = "sigmoid",
            lastactivation: str = "identity", nout: int = 1, layernorm: bool = True,
            rng: Optional[np.random.Generator] = None) -> Chain:
    """src/models.jl:62,65-69.  ``pairnet(data)`` takes n = featuredim(data)."""
    if n is None:
        n = data.featuredim() if hasattr(data, "featuredim") else int(np.shape(data[0])[0])
    return densenet(pairnet_layers(n, layers, nout), activation, lastactivation, layernorm, rng)


def smallnet(nin: int, nout: int = 1, activation: str = "sigmoid", lastactivation: str = "identity",
             rng: Optional[np.random.Generator] = None) -> Chain:
    """src/models.jl:102-108"""
    return densenet([nin, 8, 8, 8, nout], activation, lastactivation, False, rng)


def inputdim(model: Chain) -> int:
    """src/models.jl:26-27"""
    return model.widths[0]


def outputdim(model: Chain) -> int:
    """src/models.jl:30-31"""
    return model.widths[-1]
