"""Handle wrapper around one ``isokann_ctx`` -- what the Julia shim's ``B200Model`` struct is
(julia/ISOKANNB200.jl).  Converts status codes 1-4 into the reference's ``DomainError``s."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import lib as L
from .models import Chain, OptimiserRule

DOMAIN_MESSAGES = {
    L.DOMAIN_CONSTANT_CHI: "Could not compute the shift-scale. chi function is constant",
    L.DOMAIN_NONFINITE_LOSS: "The ISOKANN model collapsed under training. Try reducing the learning rate or "
                             "increasing regularization",
    L.DOMAIN_SINGULAR_SIMPLEX: "Could not compute the simplex transformation. The subspace might be "
                               "singular/collapsed",
    L.DOMAIN_PINV: "Could not compute the pseudoinverse. The subspace might be singular/collapsed",
}


class DomainError(ValueError):
    """Julia's DomainError as thrown by the reference hot path (src/iso.jl:188, src/isotarget.jl:39,96,162)."""

    def __init__(self, code: int, msg: str):
        super().__init__(msg)
        self.code = code


class IsokannError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[isokann status {code}] {msg}")
        self.code = code


def julia_f32(a, ndim: Optional[int] = None) -> np.ndarray:
    """Julia-shaped array -> float32 Fortran-ordered (column-major) buffer, copying only if needed."""
    a = np.asarray(a)
    if a.dtype != np.float32 or not a.flags["F_CONTIGUOUS"]:
        a = np.asfortranarray(a, dtype=np.float32)
    if ndim is not None:
        assert a.ndim == ndim, f"expected {ndim}-D array, got shape {a.shape}"
    return a


class Engine:
    def __init__(self, model: Chain, opt: OptimiserRule, featurizer: str = "identity", n_atoms: int = 0,
                 index: Optional[Sequence[int]] = None, device: int = 0, gemm: str = "auto", chunk: int = 0):
        self.lib = L.load()
        cfg = L.Config()
        cfg.n_layers = len(model.widths) - 1
        for i, w in enumerate(model.widths):
            cfg.widths[i] = int(w)
        cfg.layernorm = int(model.layernorm)
        cfg.ln_eps = float(model.ln_eps)
        cfg.activation = L.ACT[model.activation]
        cfg.last_activation = L.ACT[model.lastactivation]
        cfg.optimiser = L.OPT[opt.kind]
        cfg.eta, cfg.lam = float(opt.eta), float(opt.reg)
        cfg.beta1, cfg.beta2, cfg.eps, cfg.rho = float(opt.beta1), float(opt.beta2), float(opt.eps), float(opt.rho)
        cfg.featurizer = L.FEAT[featurizer]
        cfg.n_atoms = int(n_atoms)
        self._index = None
        if index is not None and len(index) > 0:
            self._index = np.ascontiguousarray(np.asarray(index, dtype=np.int32).ravel())
            cfg.index = self._index.ctypes.data_as(C.POINTER(C.c_int32))
            cfg.n_index = len(self._index) // (2 if featurizer == "pairs" else 1)
        cfg.device = int(device)
        cfg.gemm_mode = L.GEMM[gemm]
        cfg.chunk = int(chunk)
        h = C.c_void_p()
        rc = self.lib.isokann_create(C.byref(cfg), C.byref(h))
        if rc != L.OK:
            raise IsokannError(rc, (self.lib.isokann_last_error(None) or b"").decode())
        self.h = h
        self.P = int(self.lib.isokann_num_params(h))
        self.F = int(self.lib.isokann_feature_dim(h))
        self.D = int(self.lib.isokann_coord_dim(h))
        self.d = int(model.widths[-1])
        self.N = 0
        self.K = 0
        self.kind = opt.kind
        self._keep = []
        self.upload_params(model.flat())

    # -- plumbing --
    def _check(self, rc: int):
        if rc == L.OK:
            return
        msg = (self.lib.isokann_last_error(self.h) or b"").decode()
        if rc in DOMAIN_MESSAGES:
            raise DomainError(rc, msg or DOMAIN_MESSAGES[rc])
        raise IsokannError(rc, msg)

    def close(self):
        if getattr(self, "h", None):
            self.lib.isokann_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- multi GPU --
    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        rc = L.load().isokann_comm_get_unique_id(buf)
        if rc != L.OK:
            raise IsokannError(rc, "ncclGetUniqueId failed")
        return buf.raw

    def comm_init(self, world: int, rank: int, uid: Optional[bytes]):
        buf = C.create_string_buffer(uid, 128) if uid is not None else None
        self._check(self.lib.isokann_comm_init(self.h, world, rank, buf))
        self.world, self.rank = world, rank

    # -- data --
    def set_data(self, xs, ys=None, n_offset: Optional[int] = None, n_local: Optional[int] = None):
        """xs: Julia-shaped (D, N); ys: (D, K, N_local) or None."""
        if np.asarray(xs).dtype == np.float64 and n_offset is None:
            xs = np.asfortranarray(xs)
            ys = None if ys is None else np.asfortranarray(ys)
            D, N = xs.shape
            K = 0 if ys is None else ys.shape[1]
            self._check(self.lib.isokann_set_data_f64(self.h, L.ptr(xs), L.ptr(ys), D, K, N))
        else:
            xs = julia_f32(xs, 2)
            ys = None if ys is None else julia_f32(ys, 3)
            D, N = xs.shape
            K = 0 if ys is None else ys.shape[1]
            if n_offset is None:
                self._check(self.lib.isokann_set_data(self.h, L.ptr(xs), L.ptr(ys), D, K, N))
            else:
                self._check(self.lib.isokann_set_data_sharded(self.h, L.ptr(xs), L.ptr(ys), D, K, N, n_offset, n_local))
        self.N, self.K = N, K

    def set_data_async(self, xs, ys, n_offset: int = 0, n_local: Optional[int] = None):
        """like set_data, but ys and then xs (float32, Fortran-ordered; page-locked by the library) are streamed in
        behind the call: the next Koopman pass consumes ys chunk-wise, the first reader of xs waits for it on
        the device; the caller keeps both alive until results computed from them have come back"""
        xs = julia_f32(xs, 2)
        ys = julia_f32(ys, 3)
        D, N = xs.shape
        K = ys.shape[1]
        n_local = ys.shape[2] if n_local is None else n_local
        self._keep = [xs, ys]
        self._check(self.lib.isokann_set_data_async(self.h, L.ptr(xs), L.ptr(ys), D, K, N, n_offset, n_local))
        self.N, self.K = N, K

    def release_host_buffers(self):
        """un-pin the arrays handed to set_data_async (call before freeing them)"""
        self._check(self.lib.isokann_release_host_buffers(self.h))
        self._keep = []

    def append_data(self, xs_new, ys_new):
        xs_new = julia_f32(xs_new, 2)
        ys_new = None if ys_new is None else julia_f32(ys_new, 3)
        D, n = xs_new.shape
        K = 0 if ys_new is None else ys_new.shape[1]
        self._check(self.lib.isokann_append_data(self.h, L.ptr(xs_new), L.ptr(ys_new), D, K, n))
        self.N += n

    def keep_last(self, n_keep: int):
        self._check(self.lib.isokann_keep_last(self.h, int(n_keep)))
        self.N = min(self.N, int(n_keep))

    def chis_prop(self) -> np.ndarray:
        out = np.empty((self.d, self.K, self.N), dtype=np.float32, order="F")
        self._check(self.lib.isokann_chis_prop(self.h, L.ptr(out)))
        return out

    def set_data_dev(self, dev_xs, dev_ys, D: int, K: int, N: int, n_offset: int = 0, n_local: Optional[int] = None):
        """device-resident float32 buffers (torch tensors or raw addresses), records layout."""
        n_local = N if n_local is None else n_local
        self._keep = [dev_xs, dev_ys]
        self._check(self.lib.isokann_set_data_dev(self.h, L.ptr(dev_xs), L.ptr(dev_ys), D, K, N, n_offset, n_local))
        self.N, self.K = N, K

    def set_koopman_weights(self, w):
        """w: Julia-shaped (K, N_local) weights of WeightedSamples, or None"""
        buf = None if w is None else julia_f32(w)      # keep the converted buffer alive across the call
        self._check(self.lib.isokann_set_koopman_weights(self.h, L.ptr(buf)))

    # -- parameters --
    def upload_params(self, flat: np.ndarray):
        flat = np.ascontiguousarray(flat, dtype=np.float32)
        self._check(self.lib.isokann_upload_params(self.h, L.ptr(flat), flat.size))

    def download_params(self) -> np.ndarray:
        out = np.empty(self.P, dtype=np.float32)
        self._check(self.lib.isokann_download_params(self.h, L.ptr(out), out.size))
        return out

    def download_grads(self) -> np.ndarray:
        """flat gradient of the last optimiser step (what Zygote.withgradient returns at src/iso.jl:185)"""
        out = np.empty(self.P, dtype=np.float32)
        self._check(self.lib.isokann_download_grads(self.h, L.ptr(out), out.size))
        return out

    def target_matrices(self):
        """(Kinv, schur(Kinv).vectors, A) of the last N-D target; Julia-shaped (d, d) arrays, target = A @ Kchi"""
        d = self.d
        kinv = np.zeros((d, d), dtype=np.float32, order="F")
        z = np.zeros((d, d), dtype=np.float32, order="F")
        a = np.zeros((d, d), dtype=np.float64)
        self._check(self.lib.isokann_target_matrices(self.h, L.ptr(kinv), L.ptr(z), L.ptr(a)))
        return kinv, z, a

    def upload_opt_state(self, m, v=None, beta_t=None):
        m = np.ascontiguousarray(m, dtype=np.float32)
        v = None if v is None else np.ascontiguousarray(v, dtype=np.float32)
        bt = None if beta_t is None else np.ascontiguousarray(beta_t, dtype=np.float32)
        self._check(self.lib.isokann_upload_opt_state(self.h, L.ptr(m), L.ptr(v), L.ptr(bt), m.size))

    def download_opt_state(self):
        m = np.empty(self.P, dtype=np.float32)
        v = np.empty(self.P, dtype=np.float32)
        bt = np.zeros(2, dtype=np.float32)
        self._check(self.lib.isokann_download_opt_state(self.h, L.ptr(m), L.ptr(v), L.ptr(bt), self.P))
        return (m, v, bt) if self.kind == "adam" else (m, None, None)

    # -- compute --
    def featurize(self, coords) -> np.ndarray:
        """coords Julia-shaped (D, ...) -> features (F, ...)"""
        c = julia_f32(coords)
        lead = c.shape[1:]
        M = int(np.prod(lead)) if lead else 1
        out = np.empty((self.F,) + tuple(lead), dtype=np.float32, order="F")
        self._check(self.lib.isokann_featurize(self.h, L.ptr(c), c.shape[0], M, L.ptr(out)))
        return out

    def forward(self, x, is_features: bool = False) -> np.ndarray:
        """model(x): x Julia-shaped (rows, ...) -> (d, ...)"""
        x = julia_f32(x)
        lead = x.shape[1:]
        M = int(np.prod(lead)) if lead else 1
        out = np.empty((self.d,) + tuple(lead), dtype=np.float32, order="F")
        self._check(self.lib.isokann_forward(self.h, L.ptr(x), x.shape[0], M, int(is_features), L.ptr(out)))
        return out

    def chi_vjp(self, x, cot=None, is_features: bool = False) -> np.ndarray:
        """d(sum(cot .* model(x)))/dx for Julia-shaped x (rows, ...) -> same shape; cot (d, ...) defaults to ones"""
        x = julia_f32(x)
        lead = x.shape[1:]
        M = int(np.prod(lead)) if lead else 1
        ct = None if cot is None else julia_f32(cot)
        out = np.empty(x.shape, dtype=np.float32, order="F")
        self._check(self.lib.isokann_chi_vjp(self.h, L.ptr(x), x.shape[0], M, int(is_features), L.ptr(ct), L.ptr(out)))
        return out

    def chis(self) -> np.ndarray:
        out = np.empty((self.d, self.N), dtype=np.float32, order="F")
        self._check(self.lib.isokann_chis(self.h, L.ptr(out)))
        return out

    def koopman(self) -> np.ndarray:
        out = np.empty((self.d, self.N), dtype=np.float32, order="F")
        self._check(self.lib.isokann_koopman(self.h, L.ptr(out)))
        return out

    @staticmethod
    def _opts(kw) -> L.TargetOpts:
        return L.TargetOpts(int(kw.get("permute", True)), int(kw.get("whitening", False)),
                            int(kw.get("normalize", True)), int(kw.get("direct", True)),
                            int(kw.get("eigenvecs", True)))

    def target(self, transform: str, fetch: bool = True, **kw) -> Optional[np.ndarray]:
        out = np.empty((self.d, self.N), dtype=np.float32, order="F") if fetch else None
        o = self._opts(kw)
        self._check(self.lib.isokann_target(self.h, L.TARGET[transform], C.byref(o), L.ptr(out)))
        return out

    def download_target(self) -> np.ndarray:
        out = np.empty((self.d, self.N), dtype=np.float32, order="F")
        self._check(self.lib.isokann_download_target(self.h, L.ptr(out)))
        return out

    def validationloss(self, vxs, vys) -> float:
        """validationloss(iso, valdata) (src/iso.jl:160-168) on the device; vxs (D, Nv), vys (D, K, Nv)"""
        vxs, vys = julia_f32(vxs, 2), julia_f32(vys, 3)
        D, K, Nv = vys.shape
        out = C.c_double()
        self._check(self.lib.isokann_validationloss(self.h, L.ptr(vxs), L.ptr(vys), D, K, Nv, C.byref(out)))
        return out.value

    def rates(self) -> np.ndarray:
        """rates(iso) * lagtime (src/iso.jl:339-351): log(Kchi / chi) on the resident data, (dim, dim) float64"""
        q = np.zeros(64, dtype=np.float64)
        dim = C.c_int32()
        self._check(self.lib.isokann_rates(self.h, L.ptr(q), C.byref(dim)))
        n = dim.value
        return q[:n * n].reshape(n, n, order="F").copy()

    def residual_subspace(self, v_norms: bool = False, want_res: bool = False):
        """residual_subspace(iso) (src/isotarget.jl:805-821) -> (res (N, d) float64 or None, relres (d,))"""
        d, N = self.d, self.N
        relres = np.zeros(d, dtype=np.float64)
        res = np.zeros((N, d), dtype=np.float64, order="F") if want_res else None
        self._check(self.lib.isokann_residual_subspace(self.h, 1 if v_norms else 0, L.ptr(relres), L.ptr(res)))
        return res, relres

    def residual_ritz(self, want_residues: bool = False):
        """residual_ritz(iso) (src/isotarget.jl:787-802) -> dict(residues, relres, vals, vecs); complex128 arrays"""
        d, N = self.d, self.N
        vals = np.zeros(d, dtype=np.complex128)
        vecs = np.zeros((d, d), dtype=np.complex128, order="F")
        relres = np.zeros(d, dtype=np.float64)
        residues = np.zeros((N, d), dtype=np.complex128, order="F") if want_residues else None
        self._check(self.lib.isokann_residual_ritz(self.h, L.ptr(vals), L.ptr(vecs), L.ptr(relres), L.ptr(residues)))
        return {"residues": residues, "relres": relres, "vals": vals, "vecs": vecs}

    @staticmethod
    def randperm(state4, n: int):
        """Julia's randperm(Xoshiro(s0, s1, s2, s3), n): returns (1-based permutation, advanced state)"""
        st = np.ascontiguousarray(state4, dtype=np.uint64).copy()
        out = np.empty(n, dtype=np.int64)
        rc = L.load().isokann_randperm(L.ptr(st), int(n), L.ptr(out))
        if rc != L.OK:
            raise IsokannError(rc, "isokann_randperm: bad argument")
        return out, st

    def set_target(self, target):
        t = julia_f32(target, 2)
        self._check(self.lib.isokann_set_target(self.h, L.ptr(t), t.shape[0], t.shape[1]))

    def train_epoch(self, perm1, minibatch: int, partial: bool = False) -> float:
        p = np.ascontiguousarray(perm1, dtype=np.int64)
        assert p.size == self.N
        loss = C.c_double()
        self._check(self.lib.isokann_train_epoch(self.h, L.ptr(p), int(minibatch), int(partial), C.byref(loss)))
        return loss.value

    def iterate(self, transform: str, n_iter: int, epochs: int, minibatch: int, perms1, **kw) -> np.ndarray:
        p = np.ascontiguousarray(perms1, dtype=np.int64)
        assert p.size == n_iter * epochs * self.N
        losses = np.zeros(n_iter * epochs, dtype=np.float64)
        o = self._opts(kw)
        self._check(self.lib.isokann_iterate(self.h, L.TARGET[transform], C.byref(o), n_iter, epochs, int(minibatch),
                                             L.ptr(p), losses.ctypes.data_as(C.POINTER(C.c_double))))
        return losses

    # -- accounting --
    def enable_timing(self, on: bool = True):
        self._check(self.lib.isokann_enable_timing(self.h, int(on)))

    def stats(self) -> dict:
        s = L.Stats()
        self._check(self.lib.isokann_get_stats(self.h, C.byref(s)))
        return s.as_dict()

    def reset_stats(self):
        self._check(self.lib.isokann_reset_stats(self.h))

    def synchronize(self):
        self._check(self.lib.isokann_synchronize(self.h))

    def stream(self) -> int:
        return int(self.lib.isokann_stream(self.h) or 0)
