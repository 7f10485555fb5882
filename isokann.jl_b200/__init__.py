"""isokann.jl_b200 -- B200-native drop-in for the per-iteration hot path of axsk/ISOKANN.jl.

The directory name contains a dot, so it is not importable by name; load it with
``__graft_entry__.load_package()`` (importlib, module alias ``isokann_jl_b200``).

Layout: ``csrc/`` holds the CUDA kernels and the C ABI (libisokann_b200.so, declared in
include/isokann_b200.h); the Python modules are the host-side mirror of the reference's
``Iso / run! / SimulationData / isotarget`` interface on top of that ABI.  There is no CPU or
PyTorch fallback: without the built library and a CUDA device every compute call raises.
"""
from .data import (ExternalSimulation, FeaturesAll, FeaturesAtoms, FeaturesCoords, FeaturesPairs, SimulationData,
                   coords, features, flatpairdists, pdists, propcoords, propfeatures)
from .engine import DomainError, Engine, IsokannError
from .iso import (Iso, addcoords_, cutoff_, propchis, chi_kchi, chicoords, chis, cpu, dchidfeat, dchidx, defaultmodel, draw_perm, isotarget, koopman, load_state, run_, validationloss, rates, residual_subspace, residual_ritz,
                  save, train_batch_)
from .isotarget import TransformISA, TransformPseudoInv, TransformShiftscale
from .models import (AdamRegularized, Chain, NesterovRegularized, OptimiserRule, densenet, inputdim, outputdim, pairnet,
                     pairnet_layers, smallnet)
from . import lib, parallel, synthetic  # noqa: F401
