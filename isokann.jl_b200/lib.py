"""ctypes binding of libisokann_b200.so -- the executable stand-in for the Julia ``ccall`` shim
(julia/ISOKANNB200.jl).  Every symbol declared in include/isokann_b200.h is bound here with
the same plain-pointer signature Julia would use."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "libisokann_b200.so"

MAX_LAYERS = 8

OK = 0
DOMAIN_CONSTANT_CHI, DOMAIN_NONFINITE_LOSS, DOMAIN_SINGULAR_SIMPLEX, DOMAIN_PINV = 1, 2, 3, 4
BAD_ARGUMENT, ERR_CUDA, ERR_NCCL, ERR_STATE = 5, 10, 11, 12
ACT = {"identity": 0, "sigmoid": 1, "tanh": 2, "relu": 3}
OPT = {"nesterov": 0, "adam": 1}
FEAT = {"identity": 0, "allpairs": 1, "atoms": 2, "pairs": 3}
TARGET = {"shiftscale": 0, "isa": 1, "pinv": 2}
GEMM = {"auto": 0, "fp32": 1, "tc": 2}


class Config(C.Structure):
    _fields_ = [
        ("n_layers", C.c_int32),
        ("widths", C.c_int32 * (MAX_LAYERS + 1)),
        ("layernorm", C.c_int32),
        ("ln_eps", C.c_float),
        ("activation", C.c_int32),
        ("last_activation", C.c_int32),
        ("optimiser", C.c_int32),
        ("eta", C.c_float), ("lam", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float),
        ("eps", C.c_float), ("rho", C.c_float),
        ("featurizer", C.c_int32),
        ("n_atoms", C.c_int32),
        ("n_index", C.c_int32),
        ("index", C.POINTER(C.c_int32)),
        ("device", C.c_int32),
        ("gemm_mode", C.c_int32),
        ("chunk", C.c_int64),
    ]


class TargetOpts(C.Structure):
    _fields_ = [("permute", C.c_int32), ("whitening", C.c_int32), ("normalize", C.c_int32),
                ("direct", C.c_int32), ("eigenvecs", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [
        ("kernel_launches", C.c_int64), ("nccl_calls", C.c_int64),
        ("ms_featurize", C.c_double), ("ms_gemm", C.c_double), ("ms_reduce", C.c_double),
        ("ms_train_elementwise", C.c_double), ("ms_optimiser", C.c_double),
        ("ms_koopman_total", C.c_double), ("ms_target_total", C.c_double), ("ms_train_total", C.c_double),
        ("n_gemm_launches", C.c_int64), ("n_featurize_launches", C.c_int64),
        ("gemm_flops", C.c_double), ("featurize_bytes", C.c_double), ("ms_nccl", C.c_double),
        ("graph_launches", C.c_int64), ("gemm_mma_flops", C.c_double), ("p2p_exchanges", C.c_int64),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_p = C.c_void_p
_f = C.POINTER(C.c_float)
_d = C.POINTER(C.c_double)
_i64 = C.POINTER(C.c_int64)

# name -> (restype, argtypes); mirrors include/isokann_b200.h one to one
SIGNATURES = {
    "isokann_abi_version": (C.c_int32, []),
    "isokann_create": (C.c_int32, [C.POINTER(Config), C.POINTER(_p)]),
    "isokann_destroy": (C.c_int32, [_p]),
    "isokann_last_error": (C.c_char_p, [_p]),
    "isokann_num_params": (C.c_int64, [_p]),
    "isokann_feature_dim": (C.c_int32, [_p]),
    "isokann_coord_dim": (C.c_int32, [_p]),
    "isokann_comm_get_unique_id": (C.c_int32, [_p]),
    "isokann_comm_init": (C.c_int32, [_p, C.c_int32, C.c_int32, _p]),
    "isokann_set_data": (C.c_int32, [_p, _p, _p, C.c_int64, C.c_int64, C.c_int64]),
    "isokann_set_data_f64": (C.c_int32, [_p, _p, _p, C.c_int64, C.c_int64, C.c_int64]),
    "isokann_set_data_sharded": (C.c_int32, [_p, _p, _p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64]),
    "isokann_set_data_async": (C.c_int32, [_p, _p, _p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64]),
    "isokann_release_host_buffers": (C.c_int32, [_p]),
    "isokann_append_data": (C.c_int32, [_p, _p, _p, C.c_int64, C.c_int64, C.c_int64]),
    "isokann_keep_last": (C.c_int32, [_p, C.c_int64]),
    "isokann_chis_prop": (C.c_int32, [_p, _p]),
    "isokann_set_data_dev": (C.c_int32, [_p, _p, _p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64]),
    "isokann_set_koopman_weights": (C.c_int32, [_p, _p]),
    "isokann_upload_params": (C.c_int32, [_p, _p, C.c_int64]),
    "isokann_download_params": (C.c_int32, [_p, _p, C.c_int64]),
    "isokann_upload_opt_state": (C.c_int32, [_p, _p, _p, _p, C.c_int64]),
    "isokann_download_opt_state": (C.c_int32, [_p, _p, _p, _p, C.c_int64]),
    "isokann_featurize": (C.c_int32, [_p, _p, C.c_int64, C.c_int64, _p]),
    "isokann_forward": (C.c_int32, [_p, _p, C.c_int64, C.c_int64, C.c_int32, _p]),
    "isokann_chi_vjp": (C.c_int32, [_p, _p, C.c_int64, C.c_int64, C.c_int32, _p, _p]),
    "isokann_chis": (C.c_int32, [_p, _p]),
    "isokann_koopman": (C.c_int32, [_p, _p]),
    "isokann_target": (C.c_int32, [_p, C.c_int32, C.POINTER(TargetOpts), _p]),
    "isokann_download_target": (C.c_int32, [_p, _p]),
    "isokann_validationloss": (C.c_int32, [_p, _p, _p, C.c_int64, C.c_int64, C.c_int64, _d]),
    "isokann_rates": (C.c_int32, [_p, _p, C.POINTER(C.c_int32)]),
    "isokann_residual_subspace": (C.c_int32, [_p, C.c_int32, _p, _p]),
    "isokann_residual_ritz": (C.c_int32, [_p, _p, _p, _p, _p]),
    "isokann_randperm": (C.c_int32, [_p, C.c_int64, _p]),
    "isokann_set_target": (C.c_int32, [_p, _p, C.c_int64, C.c_int64]),
    "isokann_train_epoch": (C.c_int32, [_p, _p, C.c_int64, C.c_int32, _d]),
    "isokann_iterate": (C.c_int32, [_p, C.c_int32, C.POINTER(TargetOpts), C.c_int64, C.c_int64, C.c_int64, _p, _d]),
    "isokann_download_grads": (C.c_int32, [_p, _p, C.c_int64]),
    "isokann_target_matrices": (C.c_int32, [_p, _p, _p, _p]),
    "isokann_enable_timing": (C.c_int32, [_p, C.c_int32]),
    "isokann_get_stats": (C.c_int32, [_p, C.POINTER(Stats)]),
    "isokann_reset_stats": (C.c_int32, [_p]),
    "isokann_synchronize": (C.c_int32, [_p]),
    "isokann_stream": (_p, [_p]),
    "isokann_host_schur": (C.c_int32, [_p, C.c_int32, _p, _p]),
    "isokann_host_logm": (C.c_int32, [_p, C.c_int32, _p]),
    "isokann_host_diag": (C.c_int32, [C.c_int32, _p, _p, C.c_int32, _p]),
    "isokann_host_eig": (C.c_int32, [_p, C.c_int32, _p, _p]),
}

_lib = None


def load():
    """dlopen the library (building is __graft_entry__.build()'s job).  Raises if it is missing:
    there is no fallback implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU or PyTorch fallback for the ISOKANN hot path)")
    lib = C.CDLL(str(LIB_PATH), mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def ptr(a):
    """host pointer of a contiguous numpy array / raw address of a torch tensor / int / None"""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return C.c_void_p(int(a))
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"] or a.flags["F_CONTIGUOUS"]
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    raise TypeError(type(a))
