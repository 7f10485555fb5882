"""Synthetic inputs of the benchmark / parity configurations (BASELINE.json configs).

Everything is seeded numpy so that the oracle and the library consume byte-identical arrays.
Arrays are Julia-shaped and Fortran-ordered: xs (D, N), ys (D, K, N).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

# 22 atom positions of alanine dipeptide in nm: reference data/systems/"alanine dipeptide.pdb":2-23 (A / 10)
ADP_NM = np.array([
    [3.225, 27.427, 2.566], [3.720, 26.570, 2.110], [4.088, 25.905, 2.891], [4.557, 26.914, 1.502],
    [2.770, 25.800, 1.230], [1.600, 26.150, 1.090], [3.270, 24.640, 0.690], [4.259, 24.471, 0.810],
    [2.480, 23.690, -0.190], [1.733, 24.315, -0.679], [3.470, 23.160, -1.270], [4.219, 22.525, -0.797],
    [2.922, 22.582, -2.014], [3.963, 24.002, -1.756], [1.730, 22.590, 0.490], [2.340, 21.880, 1.280],
    [0.400, 22.430, 0.210], [-0.008, 23.118, -0.407], [-0.470, 21.350, 0.730], [0.112, 20.693, 1.376],
    [-1.290, 21.786, 1.300], [-0.873, 20.775, -0.103],
], dtype=np.float64) / 10.0


def _rotate_about_axis(pts: np.ndarray, p0: np.ndarray, p1: np.ndarray, angle: float) -> np.ndarray:
    k = (p1 - p0) / np.linalg.norm(p1 - p0)
    v = pts - p0
    c, s = np.cos(angle), np.sin(angle)
    return p0 + v * c + np.cross(k, v) * s + np.outer(v @ k, k) * (1 - c)


def adp_states(n_states: int) -> List[np.ndarray]:
    """conformers of the ADP geometry: the C-terminal part (atoms 10..22) rotated about the
    N(7)-CA(9) bond (the phi dihedral), so pair distances -- and chi -- differ between states"""
    out = [ADP_NM.copy()]
    for s in range(1, n_states):
        x = ADP_NM.copy()
        x[9:] = _rotate_about_axis(x[9:], x[6], x[8], s * 2 * np.pi / 3)
        out.append(x)
    return out


def villin_states(rng: np.random.Generator, n_atoms: int = 35) -> List[np.ndarray]:
    """two C-alpha-chain basins with 0.38 nm bonds: a compact and an extended random walk"""
    def walk(persistence):
        pts = [np.zeros(3)]
        d = np.array([1.0, 0.0, 0.0])
        for _ in range(n_atoms - 1):
            d = persistence * d + (1 - persistence) * rng.normal(size=3)
            d /= np.linalg.norm(d)
            pts.append(pts[-1] + 0.38 * d)
        return np.array(pts)
    return [walk(0.2), walk(0.9)]


def mixture_data(states: List[np.ndarray], N: int, K: int, rng: np.random.Generator, sx: float = 0.05,
                 sy: float = 0.03, dtype=np.float32) -> Tuple[np.ndarray, np.ndarray]:
    """xs = state(n) + N(0, sx^2); ys[:, k, n] = xs[:, n] + N(0, sy^2)"""
    D = states[0].size
    base = np.stack([s.reshape(-1) for s in states])                  # (S, D)
    which = rng.integers(0, len(states), size=N)
    xs = np.empty((N, D), dtype=dtype)                                # records layout == Julia (D, N) memory
    ys = np.empty((N, K, D), dtype=dtype)
    step = max(1, (1 << 22) // max(1, D * K))
    for s in range(0, N, step):
        e = min(N, s + step)
        x = base[which[s:e]] + sx * rng.standard_normal((e - s, D))
        xs[s:e] = x
        ys[s:e] = x[:, None, :] + sy * rng.standard_normal((e - s, K, D))
    return xs.T, ys.T                                                 # Fortran-ordered (D, N), (D, K, N) views


def triplewell_grad(x: np.ndarray, y: np.ndarray):
    """gradient of the triple-well potential of reference src/simulators/langevin.jl:112-118"""
    e1 = np.exp(-x ** 2 - (y - 1 / 3) ** 2)
    e2 = np.exp(-x ** 2 - (y - 5 / 3) ** 2)
    e3 = np.exp(-(x - 1) ** 2 - y ** 2)
    e4 = np.exp(-(x + 1) ** 2 - y ** 2)
    gx = 3 * e1 * (-2 * x) - 3 * e2 * (-2 * x) - 5 * e3 * (-2 * (x - 1)) - 5 * e4 * (-2 * (x + 1)) + 0.8 * x ** 3
    gy = (3 * e1 * (-2 * (y - 1 / 3)) - 3 * e2 * (-2 * (y - 5 / 3)) - 5 * e3 * (-2 * y) - 5 * e4 * (-2 * y)
          + 0.8 * (y - 1 / 3) ** 3)
    return gx, gy


def triplewell_data(N: int, K: int, rng: np.random.Generator, sigma: float = 1.0, dt: float = 0.01,
                    T: float = 1.0, dtype=np.float32):
    """xs ~ U([-2,2] x [-1.5,2.5]) (langevin.jl:47-51,107); ys = Euler-Maruyama of
    dX = -grad V dt + sigma dW to time T (langevin.jl:63-70), K independent replicas"""
    xs = np.stack([rng.uniform(-2, 2, N), rng.uniform(-1.5, 2.5, N)], axis=1)      # (N, 2)
    x = np.repeat(xs[:, None, 0], K, axis=1)
    y = np.repeat(xs[:, None, 1], K, axis=1)
    sq = sigma * np.sqrt(dt)
    for _ in range(int(round(T / dt))):
        gx, gy = triplewell_grad(x, y)
        x = x - gx * dt + sq * rng.standard_normal(x.shape)
        y = y - gy * dt + sq * rng.standard_normal(y.shape)
    ys = np.stack([x, y], axis=2)                                                   # (N, K, 2)
    return xs.astype(dtype).T, ys.astype(dtype).T


@dataclass
class Workload:
    name: str
    featurizer: str               # "allpairs" | "identity"
    n_atoms: int
    widths: List[int]
    layernorm: bool
    N: int
    K: int
    target: str                   # "shiftscale" | "isa" | "pinv"
    opt: str                      # "nesterov" | "adam"
    minibatch: int
    seed: int
    states: int = 2

    @property
    def D(self) -> int:
        return 3 * self.n_atoms if self.featurizer == "allpairs" else self.widths[0]

    @property
    def F(self) -> int:
        return self.widths[0]

    def macs(self) -> int:
        return sum(a * b for a, b in zip(self.widths[:-1], self.widths[1:]))


# BASELINE.json configs (c1..c5); N/K/minibatch can be overridden for parity-sized runs
WORKLOADS = {
    "c1": Workload("c1_adp_small", "allpairs", 22, [231, 38, 6, 1], True, 100, 5, "shiftscale", "nesterov", 100, 0),
    "c2": Workload("c2_triplewell", "identity", 0, [2, 8, 8, 8, 1], False, 100_000, 8, "shiftscale", "nesterov", 4096, 1),
    "c3": Workload("c3_villin", "allpairs", 35, [595, 71, 8, 1], True, 100_000, 8, "shiftscale", "nesterov", 1000, 2),
    # (TransformPseudoInv with the default Nesterov rule diverges on this data under the reference's own
    #  semantics -- the oracle raises the same "model collapsed" DomainError -- so the bench default is ISA + Adam)
    "c4": Workload("c4_adp_nd", "allpairs", 22, [231, 38, 6, 3], True, 1_000_000, 8, "isa", "adam", 65536, 3, 3),
    "c5": Workload("c5_villin_wide", "allpairs", 35, [595, 2048, 2048, 1], True, 1_000_000, 16, "shiftscale", "adam",
                   65536, 4),
}


def make_data(w: Workload, N: Optional[int] = None, K: Optional[int] = None, dtype=np.float32):
    N = w.N if N is None else N
    K = w.K if K is None else K
    rng = np.random.default_rng(w.seed)
    if w.featurizer == "identity":
        return triplewell_data(N, K, rng, dtype=dtype)
    if w.n_atoms == 22:
        return mixture_data(adp_states(w.states), N, K, rng, dtype=dtype)
    return mixture_data(villin_states(rng, w.n_atoms), N, K, rng, dtype=dtype)


def make_perms(w: Workload, N: int, count: int) -> np.ndarray:
    """(count, N) 1-based permutations from default_rng(seed+2)"""
    rng = np.random.default_rng(w.seed + 2)
    return np.stack([rng.permutation(N) for _ in range(count)]).astype(np.int64) + 1
