"""Host-side mirror of reference src/simulation.jl:71-143 (SimulationData and accessors),
src/utils/features.jl:18-35 (featurizer functors) and the layout rules of src/data.jl:5-12.

Arrays are Julia-shaped: xs (D, N), ys (D, K, N); they are kept Fortran-ordered so the buffer
handed to the C ABI is byte-identical to what Julia would pass.  Featurizers are *descriptions*
(the distances are computed on the GPU by libisokann_b200.so); calling one evaluates it through
the library, as ``featurizer(coords)`` does in the reference.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np


class _Featurizer:
    kind = "identity"

    def spec(self, D: int):
        """-> (kind, n_atoms, index list or None, feature dim)"""
        raise NotImplementedError

    def __call__(self, coords):
        from .engine import Engine
        from .models import Chain, NesterovRegularized
        kind, n_atoms, index, F = self.spec(int(np.shape(coords)[0]))
        probe = Chain([F, 1], False, "identity", "identity", weights=[np.zeros((1, F), np.float32, order="F")],
                      biases=[np.zeros(1, np.float32)])
        eng = Engine(probe, NesterovRegularized(), kind, n_atoms, index)
        try:
            return eng.featurize(coords)
        finally:
            eng.close()


class FeaturesCoords(_Featurizer):
    """src/utils/features.jl:18-19 (and the ``identity`` default of src/simulation.jl:15)"""

    def spec(self, D):
        return "identity", 0, None, D


class FeaturesAll(_Featurizer):
    """pairwise distances between all atoms -> flatpairdists(coords) (src/utils/features.jl:22-23)"""

    def spec(self, D):
        a = D // 3
        return "allpairs", a, None, a * (a - 1) // 2


@dataclass
class FeaturesAtoms(_Featurizer):
    """flatpairdists(coords, atominds), 1-based (src/utils/features.jl:26-29)"""
    atominds: Sequence[int]

    def spec(self, D):
        n = len(self.atominds)
        return "atoms", D // 3, list(self.atominds), n * (n - 1) // 2


@dataclass
class FeaturesPairs(_Featurizer):
    """pdists(coords, pairs), 1-based tuples (src/utils/features.jl:31-34)"""
    pairs: Sequence[Tuple[int, int]]

    def spec(self, D):
        flat = [int(v) for p in self.pairs for v in p]
        return "pairs", D // 3, flat, len(self.pairs)


def flatpairdists(x, cols=None):
    """src/utils/pairdists.jl:6-24"""
    return (FeaturesAll() if cols is None else FeaturesAtoms(list(cols)))(x)


def pdists(x, pairs):
    """src/utils/pairdists.jl:109-127"""
    return FeaturesPairs(list(pairs))(x)


@dataclass
class ExternalSimulation:
    """src/simulation.jl:41-50"""
    dict: dict = field(default_factory=dict)


class SimulationData:
    """SimulationData(sim, (xs, ys); featurizer) (src/simulation.jl:71-76,100-114).

    The reference featurizes once at construction and caches Float32 features (:112).  Here the
    coordinates are what is uploaded and the featurizer is fused into every pass on the GPU
    (same numbers, 4*D instead of 4*F bytes per sample); ``features(d)`` / ``propfeatures(d)``
    materialise the cache on demand through the library.
    """

    def __init__(self, *args, featurizer: Optional[_Featurizer] = None, weights=None):
        if len(args) == 2 and not isinstance(args[1], tuple):
            sim, (xs, ys) = ExternalSimulation(), (args[0], args[1])      # SimulationData(xs, ys)
        else:
            sim, (xs, ys) = args[0], args[1]                              # SimulationData(sim, (xs, ys))
        xs = np.asarray(xs)
        ys = np.asarray(ys)
        assert xs.ndim == 2 and ys.ndim == 3 and xs.shape[0] == ys.shape[0] and xs.shape[1] == ys.shape[2], \
            "xs must be (D, N) and ys (D, K, N)"
        self.sim = sim
        self.coords = (xs, ys)
        self.featurizer = featurizer if featurizer is not None else FeaturesCoords()
        self.weights = weights            # WeightedSamples weights (K, N), src/data.jl:187-215
        self._features = None

    # accessors, src/simulation.jl:126-143
    def getcoords(self):
        return self.coords[0]

    def propcoords(self):
        return self.coords[1]

    def features(self):
        if self._features is None:
            self._features = (self.featurizer(self.coords[0]), self.featurizer(self.coords[1]))
        return self._features[0]

    def propfeatures(self):
        self.features()
        return self._features[1]

    def featuredim(self) -> int:
        return self.featurizer.spec(self.coords[0].shape[0])[3]

    def nk(self) -> int:
        return self.coords[1].shape[1]

    def __len__(self):
        return self.coords[0].shape[1]

    def __getitem__(self, i):
        """src/simulation.jl:135 -- slicing the observations returns a new data object"""
        xs, ys = self.coords
        w = None if self.weights is None else self.weights[:, i]
        return SimulationData(self.sim, (xs[:, i], ys[:, :, i]), featurizer=self.featurizer, weights=w)


def coords(d: SimulationData):
    return d.getcoords()


def propcoords(d: SimulationData):
    return d.propcoords()


def features(d: SimulationData, coords_=None):
    """features(d) / features(d, coords) (src/simulation.jl:121-124,141)"""
    return d.features() if coords_ is None else d.featurizer(coords_)


def propfeatures(d: SimulationData):
    return d.propfeatures()
