"""CPU tests of the diagnostics (rates src/iso.jl:339-351, residual_subspace / residual_ritz
src/isotarget.jl:787-821): known-answer tests of the oracle's restatement, and the library's host-side algebra
(matrix logarithm, general eigenproblem, the moment-based formulas the device reductions feed) against scipy/numpy and
against the oracle.  The two CUDA reductions themselves are covered by tests/test_gpu_parity.py."""
import warnings

import numpy as np
import pytest
import scipy.linalg as sl


def _moments(chi, kchi):
    """what moments_kernel accumulates: sums of u u', v u', v v' with u = [chi, 1], v = [Kchi, 1] (fp64)"""
    n = chi.shape[0]
    u = np.concatenate([chi.astype(np.float64), np.ones((n, 1))], axis=1)
    v = np.concatenate([kchi.astype(np.float64), np.ones((n, 1))], axis=1)
    return np.ascontiguousarray(u.T @ u), np.ascontiguousarray(v.T @ u), np.ascontiguousarray(v.T @ v)


def _chi_pair(rng, n, d, noise=0.05, rotate=False):
    """chi (n, d) float32 and Kchi = chi M' + noise: a Koopman-like linear map plus an out-of-subspace part"""
    chi = rng.random((n, d)).astype(np.float32)
    m = 0.8 * np.eye(d) + 0.2 * rng.random((d, d)) / d
    if rotate and d >= 2:                          # a rotation block gives Kr complex eigenvalues
        th = 0.6
        m[:2, :2] = 0.9 * np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    kchi = (chi @ m.T + noise * rng.standard_normal((n, d))).astype(np.float32)
    return chi, kchi


# ---------------------------------------------------------------------------------------------
# oracle known-answer tests (derived from the cited reference lines)
# ---------------------------------------------------------------------------------------------
def test_oracle_rates_recovers_the_generator(oracle):
    rng = np.random.default_rng(0)
    d, n, tau = 3, 500, 0.7
    q = rng.random((d, d))
    q -= np.diag(q.sum(1))                          # a rate matrix: rows sum to zero
    chi = rng.random((n, d))
    kchi = chi @ sl.expm(tau * q).T                 # K chi = exp(tau Q) chi   (src/iso.jl:336)
    assert np.allclose(oracle.rates(chi, kchi) / tau, q, atol=1e-10)
    # one dimensional chi: rates of chi and 1 - chi (src/iso.jl:346-349)
    q2 = np.array([[-0.3, 0.5], [0.3, -0.5]])   # columns sum to zero: chi + (1 - chi) stays 1
    c1 = rng.random((n, 1))
    x = np.concatenate([c1, 1 - c1], axis=1)
    y = x @ sl.expm(q2).T
    assert np.allclose(y.sum(1), 1.0)
    assert np.allclose(oracle.rates(c1, y[:, :1]), q2, atol=1e-10)


def test_oracle_residuals_known_answers(oracle):
    rng = np.random.default_rng(1)
    n, d = 400, 3
    V = rng.standard_normal((n, d))
    A = 0.5 * np.eye(d) + 0.1 * rng.standard_normal((d, d))
    # K V inside span(V): both residuals vanish and the Ritz values are the eigenvalues of A
    res, relres = oracle.residual_subspace(V, V @ A)
    assert np.abs(relres).max() < 1e-12 and np.abs(res).max() < 1e-12
    residues, rr, vals, vecs, Q = oracle.residual_ritz(V, V @ A)
    assert np.abs(rr).max() < 1e-10
    assert np.allclose(np.sort_complex(vals), np.sort_complex(np.linalg.eigvals(A)), atol=1e-10)
    assert np.all(np.diff(np.abs(1 - vals)) >= -1e-12)          # sortby = x -> abs(1 - x)
    # a component orthogonal to span(V) is exactly the residual
    Qf, _ = np.linalg.qr(np.concatenate([V, rng.standard_normal((n, d))], axis=1))
    W = Qf[:, d:]                                                # orthonormal, orthogonal to span(V)
    res, relres = oracle.residual_subspace(V, V @ A + 0.25 * W)
    assert np.allclose(res, 0.25 * W, atol=1e-12)
    assert np.allclose(relres, 0.25 / np.linalg.norm(V @ A + 0.25 * W, axis=0), atol=1e-12)
    _, relres_v = oracle.residual_subspace(V, V @ A + 0.25 * W, v_norms=True)
    assert np.allclose(relres_v, 0.25 / np.linalg.norm(V, axis=0), atol=1e-12)


# ---------------------------------------------------------------------------------------------
# host-side algebra of the library (no device needed)
# ---------------------------------------------------------------------------------------------
def test_host_logm_matches_scipy(pkg):
    lib, ptr = pkg.lib.load(), pkg.lib.ptr
    rng = np.random.default_rng(5)
    worst = 0.0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for t in range(400):
            n = int(rng.integers(1, 10))
            kind = t % 4
            if kind == 0:
                a = sl.expm(rng.standard_normal((n, n)) * 0.7)
            elif kind == 1:
                a = np.eye(n) + 0.3 * rng.standard_normal((n, n)) / np.sqrt(n)
            elif kind == 2:                                      # lazy stochastic matrix (what Kchi / chi looks like)
                p = rng.random((n, n)) + 0.05
                a = 0.5 * np.eye(n) + 0.5 * p / p.sum(1, keepdims=True)
            else:
                q = rng.standard_normal((n, n))
                a = q @ q.T + 0.1 * np.eye(n)
            ev = np.linalg.eigvals(a)
            if np.any((np.abs(ev.imag) < 1e-9) & (ev.real <= 1e-6)):
                continue
            af = np.asfortranarray(a)
            out = np.zeros((n, n), order="F")
            assert lib.isokann_host_logm(ptr(af), n, ptr(out)) == 0
            ref = sl.logm(a)
            worst = max(worst, np.abs(out - ref.real).max() / max(1.0, np.abs(ref).max()))
    assert worst < 1e-11, worst
    # an eigenvalue on the negative real axis has no real logarithm: rejected, not garbage
    a = np.asfortranarray(np.diag([-1.0, 2.0]))
    assert lib.isokann_host_logm(ptr(a), 2, ptr(np.zeros((2, 2), order="F"))) == pkg.lib.BAD_ARGUMENT


def test_host_eig_matches_numpy_and_lapack_conventions(pkg):
    lib, ptr = pkg.lib.load(), pkg.lib.ptr
    rng = np.random.default_rng(6)
    for t in range(600):
        n = int(rng.integers(1, 9))
        a = rng.standard_normal((n, n))
        if t % 3 == 1:
            a = np.eye(n) * 0.8 + 0.2 * a
        if t % 3 == 2:
            a = a + a.T
        vals = np.zeros(n, dtype=np.complex128)
        vecs = np.zeros((n, n), dtype=np.complex128, order="F")
        assert lib.isokann_host_eig(ptr(np.asfortranarray(a)), n, ptr(vals), ptr(vecs)) == 0
        assert np.abs(a @ vecs - vecs * vals[None, :]).max() < 1e-12 * max(1.0, np.abs(a).max()) * 50
        wn = np.linalg.eigvals(a)
        assert np.allclose(np.sort_complex(np.round(vals, 8)), np.sort_complex(np.round(wn, 8)), atol=1e-7)
        assert np.allclose(np.linalg.norm(vecs, axis=0), 1.0, atol=1e-12)       # dgeev: unit 2-norm ...
        j = 0
        while j < n:
            ib = int(np.argmax(np.abs(vecs[:, j])))
            assert abs(vecs[ib, j].imag) < 1e-14 and vecs[ib, j].real > 0       # ... largest component real
            if vals[j].imag != 0:                                               # pairs: positive imaginary part first
                assert vals[j].imag > 0 and np.isclose(vals[j + 1], np.conj(vals[j]))
                assert np.allclose(vecs[:, j + 1], np.conj(vecs[:, j]), atol=1e-10)
                j += 2
            else:
                assert np.all(vecs[:, j].imag == 0)
                j += 1


@pytest.mark.parametrize("d", [1, 2, 3, 8])
def test_moment_formulas_match_the_oracle(pkg, oracle, d):
    """the d x d algebra behind isokann_rates / isokann_residual_subspace / isokann_residual_ritz, fed with moments
    computed in numpy, reproduces the oracle's QR-based restatement of the reference"""
    lib, ptr = pkg.lib.load(), pkg.lib.ptr
    rng = np.random.default_rng(10 + d)
    n = 3000
    for rotate in (False, True):
        chi, kchi = _chi_pair(rng, n, d, rotate=rotate)
        uu, vu, vv = _moments(chi, kchi)
        dd = d * d
        # rates
        out = np.zeros(1 + 81)
        assert lib.isokann_host_diag(0, ptr(uu), ptr(vu), d, ptr(out)) == 0
        m = int(out[0])
        assert m == max(d, 2)
        q = out[1:1 + m * m].reshape(m, m)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            q_ref = oracle.rates(chi, kchi)
        assert np.allclose(q, q_ref.real, atol=1e-9), np.abs(q - q_ref.real).max()
        # residual_subspace: res = A Kchi - B chi per record
        out = np.zeros(2 * dd)
        assert lib.isokann_host_diag(1, ptr(uu), ptr(vu), d, ptr(out)) == 0
        A, B = out[:dd].reshape(d, d), out[dd:].reshape(d, d)
        res = kchi.astype(np.float64) @ A.T - chi.astype(np.float64) @ B.T
        res_ref, relres_ref = oracle.residual_subspace(chi, kchi)
        assert np.allclose(res, res_ref, atol=1e-9)
        assert np.allclose(np.linalg.norm(res, axis=0) / np.sqrt(np.diag(vv)[:d]), relres_ref, rtol=1e-9)
        # residual_ritz
        out = np.zeros(2 * d + 2 * dd + 4 * dd + 1)
        assert lib.isokann_host_diag(2, ptr(uu), ptr(vu), d, ptr(out)) == 0
        vals = out[:2 * d].view(np.complex128)
        vecs = out[2 * d:2 * d + 2 * dd].view(np.complex128).reshape(d, d, order="F")
        mats = out[2 * d + 2 * dd:2 * d + 6 * dd].reshape(4, d, d)
        any_complex = out[-1] != 0
        residues_ref, rr_ref, vals_ref, vecs_ref, Q = oracle.residual_ritz(chi, kchi)
        assert any_complex == bool(np.iscomplexobj(vals_ref) and np.abs(vals_ref.imag).max() > 0)
        if rotate and d >= 2:
            assert any_complex
        assert np.allclose(vals, vals_ref, atol=1e-9)
        k64, c64 = kchi.astype(np.float64), chi.astype(np.float64)
        residues = (k64 @ mats[0].T - c64 @ mats[1].T) + 1j * (k64 @ mats[2].T - c64 @ mats[3].T)
        kq = k64 @ mats[0].T + 1j * (k64 @ mats[2].T)
        relres = np.linalg.norm(residues, axis=0) / np.linalg.norm(kq, axis=0)
        assert np.allclose(relres, rr_ref, rtol=1e-8)
        # eigenvectors and residues agree up to the sign of R's diagonal (rows) and a unit phase per column
        _, R = np.linalg.qr(chi.astype(np.float64))
        sgn = np.sign(np.diag(R))
        for j in range(d):
            v_ref = sgn * vecs_ref[:, j]
            ph = np.vdot(v_ref, vecs[:, j])
            ph /= abs(ph)
            assert np.allclose(vecs[:, j], ph * v_ref, atol=1e-7)
            assert np.allclose(residues[:, j], ph * residues_ref[:, j], atol=1e-7)


def test_diag_rejects_collapsed_chi(pkg):
    lib, ptr = pkg.lib.load(), pkg.lib.ptr
    rng = np.random.default_rng(3)
    chi = rng.random((100, 1)).astype(np.float32) @ np.ones((1, 2), np.float32)      # two identical chi components
    uu, vu, _ = _moments(chi, chi)
    out = np.zeros(200)
    assert lib.isokann_host_diag(2, ptr(uu), ptr(vu), 2, ptr(out)) == pkg.lib.DOMAIN_PINV
