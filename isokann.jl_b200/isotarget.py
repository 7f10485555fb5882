"""Host-side mirror of the target transforms of reference src/isotarget.jl (structs only; the
computation is ``isokann_target`` in the library)."""
from __future__ import annotations

from dataclasses import dataclass


@dataclass
class TransformShiftscale:
    """Classical 1-D shift-scale (src/isotarget.jl:32-42)"""
    name = "shiftscale"

    def opts(self):
        return {}


@dataclass
class TransformISA:
    """Inner simplex algorithm target (src/isotarget.jl:74-107)"""
    permute: bool = True
    whitening: bool = False
    name = "isa"

    def opts(self):
        return {"permute": self.permute, "whitening": self.whitening}


@dataclass
class TransformPseudoInv:
    """Pseudo-inverse target (src/isotarget.jl:145-179)"""
    normalize: bool = True
    direct: bool = True
    eigenvecs: bool = True
    permute: bool = True
    name = "pinv"

    def opts(self):
        return {"normalize": self.normalize, "direct": self.direct, "eigenvecs": self.eigenvecs,
                "permute": self.permute}
