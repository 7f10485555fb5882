// Host-side d x d algebra (d <= 8, double precision) of the device-side diagnostics:
//   rates(iso)             src/iso.jl:339-351          log(Kchi / chi)            -> host_logm
//   residual_ritz(iso)     src/isotarget.jl:787-802    eigen(Q' K Q, sortby=...)  -> host_eig_general
//   residual_subspace(iso) src/isotarget.jl:805-821    Q Q' KV                    -> host_cholesky_upper
// The O(N) work (second moments of chi and Kchi, residual matrices and their column norms) are CUDA reductions
// (csrc/reduce.cu); what is left is O(d^3).  The reference calls LAPACK through Julia's LinearAlgebra here
// (`/` = least squares via QR, `log` = Schur based matrix logarithm, `eigen` = dgeev); the results are defined by the
// mathematics up to rounding, so other numerically sound algorithms are used:
//   eigen: real Schur (the dlahqr pipeline of hostlinalg.cpp in double), rotated to a complex triangular form
//          (the rsf2csf step), eigenvectors by back substitution, LAPACK's normalisation (unit 2-norm, largest
//          component real -- here also positive, which fixes the sign LAPACK leaves to its Schur vectors);
//   log:   inverse scaling and squaring -- product-form Denman-Beavers square roots until ||A - I||_1 <= 1/4, then the
//          Gregory series log A = 2 atanh((A - I)(A + I)^-1).
#include <cmath>
#include <complex>
#include <cstring>
#include <algorithm>

#include "common.cuh"

namespace ik {

namespace {
using cd = std::complex<double>;
constexpr int M = kMaxD + 1;  // rates of a one dimensional chi work on [chi; 1 - chi]: allow d + 1

void matmul(const double *a, const double *b, int n, double *c) {
  double t[M * M];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      double s = 0.0;
      for (int k = 0; k < n; ++k) s += a[i * n + k] * b[k * n + j];
      t[i * n + j] = s;
    }
  std::memcpy(c, t, sizeof(double) * n * n);
}

double norm1_minus_identity(const double *a, int n) {
  double best = 0.0;
  for (int j = 0; j < n; ++j) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += std::fabs(a[i * n + j] - (i == j ? 1.0 : 0.0));
    best = std::max(best, s);
  }
  return best;
}

bool inverse_n(const double *a, int n, double *inv) {  // Gauss-Jordan with partial pivoting, n <= M
  double m[M][2 * M];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      m[i][j] = a[i * n + j];
      m[i][n + j] = (i == j) ? 1.0 : 0.0;
      if (!std::isfinite(m[i][j])) return false;
    }
  for (int col = 0; col < n; ++col) {
    int piv = col;
    double best = std::fabs(m[col][col]);
    for (int r = col + 1; r < n; ++r)
      if (std::fabs(m[r][col]) > best) {
        best = std::fabs(m[r][col]);
        piv = r;
      }
    if (best == 0.0) return false;
    if (piv != col)
      for (int j = 0; j < 2 * n; ++j) std::swap(m[piv][j], m[col][j]);
    const double p = m[col][col];
    for (int j = 0; j < 2 * n; ++j) m[col][j] /= p;
    for (int r = 0; r < n; ++r) {
      if (r == col) continue;
      const double f = m[r][col];
      if (f == 0.0) continue;
      for (int j = 0; j < 2 * n; ++j) m[r][j] -= f * m[col][j];
    }
  }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      inv[i * n + j] = m[i][n + j];
      if (!std::isfinite(inv[i * n + j])) return false;
    }
  return true;
}

// principal square root by the product form of the Denman-Beavers iteration:
//   M <- (I + (M + M^-1)/2)/2,  Y <- Y (I + M^-1)/2,   M -> I, Y -> A^(1/2)
// converges for matrices without eigenvalues on the closed negative real axis
bool sqrtm_db(const double *a, int n, double *out) {
  double Mk[M * M], Y[M * M], Mi[M * M], t[M * M];
  std::memcpy(Mk, a, sizeof(double) * n * n);
  std::memcpy(Y, a, sizeof(double) * n * n);
  for (int it = 0; it < 60; ++it) {
    if (!inverse_n(Mk, n, Mi)) return false;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) t[i * n + j] = 0.5 * ((i == j ? 1.0 : 0.0) + Mi[i * n + j]);
    matmul(Y, t, n, Y);
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j)
        Mk[i * n + j] = 0.5 * ((i == j ? 1.0 : 0.0) + 0.5 * (Mk[i * n + j] + Mi[i * n + j]));
    for (int i = 0; i < n * n; ++i)
      if (!std::isfinite(Y[i]) || !std::isfinite(Mk[i])) return false;
    if (norm1_minus_identity(Mk, n) <= 1e-15 * n) {
      std::memcpy(out, Y, sizeof(double) * n * n);
      return true;
    }
  }
  return false;
}
}  // namespace

// principal logarithm of a real n x n matrix (row-major), n <= kMaxD + 1.  false: not finite, singular, or an
// eigenvalue on the closed negative real axis (the reference's `log` would return a complex matrix there)
bool host_logm(const double *a, int n, double *out) {
  if (n < 1 || n > M) return false;
  double A[M * M];
  for (int i = 0; i < n * n; ++i) {
    if (!std::isfinite(a[i])) return false;
    A[i] = a[i];
  }
  int k = 0;
  while (norm1_minus_identity(A, n) > 0.25) {
    if (k >= 60 || !sqrtm_db(A, n, A)) return false;
    ++k;
  }
  // Z = (A - I)(A + I)^-1, log A = 2 (Z + Z^3/3 + Z^5/5 + ...), ||Z|| <= 1/7
  double P[M * M], Q[M * M], Qi[M * M], Z[M * M], Z2[M * M], term[M * M], sum[M * M];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      P[i * n + j] = A[i * n + j] - (i == j ? 1.0 : 0.0);
      Q[i * n + j] = A[i * n + j] + (i == j ? 1.0 : 0.0);
    }
  if (!inverse_n(Q, n, Qi)) return false;
  matmul(P, Qi, n, Z);
  matmul(Z, Z, n, Z2);
  std::memcpy(term, Z, sizeof(double) * n * n);
  std::memcpy(sum, Z, sizeof(double) * n * n);
  for (int j = 1; j < 14; ++j) {
    matmul(term, Z2, n, term);
    for (int i = 0; i < n * n; ++i) sum[i] += term[i] / (double)(2 * j + 1);
  }
  const double scale = 2.0 * std::ldexp(1.0, k);
  for (int i = 0; i < n * n; ++i) {
    out[i] = scale * sum[i];
    if (!std::isfinite(out[i])) return false;
  }
  return true;
}

// upper Cholesky factor R (row-major) of a symmetric positive definite matrix: R' R = G.  This is the R of the thin
// QR factorisation of V with G = V' V, taken with a positive diagonal.
bool host_cholesky_upper(const double *g, int n, double *r) {
  for (int i = 0; i < n * n; ++i) r[i] = 0.0;
  for (int i = 0; i < n; ++i)
    for (int j = i; j < n; ++j) {
      double s = g[i * n + j];
      for (int k = 0; k < i; ++k) s -= r[k * n + i] * r[k * n + j];
      if (i == j) {
        if (!(s > 0.0) || !std::isfinite(s)) return false;
        r[i * n + i] = std::sqrt(s);
      } else {
        r[i * n + j] = s / r[i * n + i];
      }
    }
  return true;
}

// eigenvalues (wr + i wi) and right eigenvectors (column k of vre + i vim, row-major n x n) of a real matrix, in
// LAPACK's order (the order of the real Schur form, of a conjugate pair the member with positive imaginary part
// first).  Every vector has unit 2-norm and its largest component is real and positive.
bool host_eig_general(const double *a_rowmajor, int n, double *wr, double *wi, double *vre, double *vim) {
  if (n < 1 || n > kMaxD) return false;
  double acol[kMaxD * kMaxD], zcol[kMaxD * kMaxD], tcol[kMaxD * kMaxD];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) acol[i + j * n] = a_rowmajor[i * n + j];
  if (!host_schur_f64(acol, n, zcol, tcol)) return false;
  cd T[kMaxD][kMaxD], Z[kMaxD][kMaxD];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      T[i][j] = tcol[i + j * n];
      Z[i][j] = zcol[i + j * n];
    }
  // 2 x 2 blocks of the real Schur form (standardised: equal diagonal, off-diagonals of opposite sign) are
  // triangularised by a complex rotation that puts the eigenvalue with positive imaginary part first
  for (int m = n - 1; m >= 1; --m) {
    const double sub = tcol[m + (m - 1) * n];
    if (sub == 0.0) continue;
    const double aa = tcol[(m - 1) + (m - 1) * n], bb = tcol[(m - 1) + m * n], dd = tcol[m + m * n];
    // eigenvalues of [[aa, bb], [sub, dd]]
    const double tr = 0.5 * (aa + dd), det = (aa - tr) * (dd - tr) - bb * sub;
    cd lam;
    if (det >= 0.0) lam = cd(tr, std::sqrt(det));  // complex pair, positive imaginary part
    else lam = cd(tr + std::sqrt(-det), 0.0);
    // eigenvector of the block for lam: (lam - dd, sub); rotate it onto e1
    cd x0 = lam - cd(dd, 0.0), x1 = cd(sub, 0.0);
    const double r = std::sqrt(std::norm(x0) + std::norm(x1));
    if (r == 0.0) continue;
    const cd c = x0 / r, s = x1 / r;
    // G = [[conj(c), conj(s)], [-s, c]] is unitary with G (x0, x1)' = (r, 0)'; T <- G T G^H, Z <- Z G^H
    for (int j = 0; j < n; ++j) {
      const cd t0 = T[m - 1][j], t1 = T[m][j];
      T[m - 1][j] = std::conj(c) * t0 + std::conj(s) * t1;
      T[m][j] = -s * t0 + c * t1;
    }
    for (int i = 0; i < n; ++i) {
      const cd t0 = T[i][m - 1], t1 = T[i][m];
      T[i][m - 1] = t0 * c + t1 * s;
      T[i][m] = -t0 * std::conj(s) + t1 * std::conj(c);
      const cd z0 = Z[i][m - 1], z1 = Z[i][m];
      Z[i][m - 1] = z0 * c + z1 * s;
      Z[i][m] = -z0 * std::conj(s) + z1 * std::conj(c);
    }
    T[m][m - 1] = 0.0;
  }
  double tnorm = 0.0;
  for (int i = 0; i < n; ++i)
    for (int j = i; j < n; ++j) tnorm = std::max(tnorm, std::abs(T[i][j]));
  const double smin = std::max(tnorm * 2.220446049250313e-16, 2.2250738585072014e-308);
  for (int k = 0; k < n; ++k) {
    const cd lam = T[k][k];
    const bool real_ev = (k + 1 >= n || tcol[(k + 1) + k * n] == 0.0) && (k == 0 || tcol[k + (k - 1) * n] == 0.0);
    wr[k] = lam.real();
    wi[k] = real_ev ? 0.0 : lam.imag();
    cd y[kMaxD];
    for (int i = 0; i < n; ++i) y[i] = 0.0;
    y[k] = 1.0;
    for (int i = k - 1; i >= 0; --i) {
      cd s = 0.0;
      for (int j = i + 1; j <= k; ++j) s += T[i][j] * y[j];
      cd den = T[i][i] - lam;
      if (std::abs(den) < smin) den = smin;  // LAPACK's perturbation of a (numerically) repeated eigenvalue
      y[i] = -s / den;
    }
    cd x[kMaxD];
    double nrm = 0.0, big = -1.0;
    int ib = 0;
    for (int i = 0; i < n; ++i) {
      cd s = 0.0;
      for (int j = 0; j <= k; ++j) s += Z[i][j] * y[j];
      x[i] = s;
      nrm += std::norm(s);
    }
    nrm = std::sqrt(nrm);
    if (!(nrm > 0.0) || !std::isfinite(nrm)) return false;
    for (int i = 0; i < n; ++i)
      if (std::abs(x[i]) > big * (1.0 + 1e-12)) {  // first of (numerically) equal maxima
        big = std::abs(x[i]);
        ib = i;
      }
    const cd phase = std::conj(x[ib]) / std::abs(x[ib]);
    for (int i = 0; i < n; ++i) {
      const cd v = x[i] * phase / nrm;
      vre[i * n + k] = v.real();
      vim[i * n + k] = real_ev ? 0.0 : v.imag();
    }
    vim[ib * n + k] = 0.0;
  }
  return true;
}

// ---- the diagnostics as functions of the second moments (row-major e x e, e = d + 1) of u = [chi, 1], v = [Kchi, 1]:
//      uu = sum_n u u', vu = sum_n v u'.  Pure host code, so the CPU test-suite exercises it without a device. ----

// rates (src/iso.jl:345-351): Q = log((y x')(x x')^-1), x = chi (d > 1) or [chi; 1 - chi] (d == 1).
// Returns 0, or 1: x x' singular, 2: no real logarithm.  q is n x n row-major, n = max(d, 2).
int diag_rates(const double *uu, const double *vu, int d, double *q, int *n_out) {
  const int e = d + 1, n = d == 1 ? 2 : d;
  double Gxx[kMaxD * kMaxD], Gyx[kMaxD * kMaxD], Gi[kMaxD * kMaxD], Mq[kMaxD * kMaxD];
  if (d == 1) {  // [chi; 1 - chi] = P [chi; 1]
    const double P[4] = {1.0, 0.0, -1.0, 1.0};
    auto sandwich = [&](const double *m, double *out) {  // P m P'
      double t[4];
      for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) t[i * 2 + j] = P[i * 2] * m[j] + P[i * 2 + 1] * m[2 + j];
      for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) out[i * 2 + j] = t[i * 2] * P[j * 2] + t[i * 2 + 1] * P[j * 2 + 1];
    };
    sandwich(uu, Gxx);
    sandwich(vu, Gyx);
  } else {
    for (int a = 0; a < d; ++a)
      for (int b = 0; b < d; ++b) {
        Gxx[a * d + b] = uu[a * e + b];
        Gyx[a * d + b] = vu[a * e + b];
      }
  }
  // y / x = (y x')(x x')^-1: the least-squares right division of a full-rank d x N system
  if (!host_inverse(Gxx, n, Gi)) return 1;
  for (int a = 0; a < n; ++a)
    for (int b = 0; b < n; ++b) {
      double s = 0.0;
      for (int k = 0; k < n; ++k) s += Gyx[a * n + k] * Gi[k * n + b];
      Mq[a * n + b] = s;
    }
  if (!host_logm(Mq, n, q)) return 2;
  *n_out = n;
  return 0;
}

// residual_subspace (src/isotarget.jl:810-813): Q Q' KV = V (V'V)^-1 V'KV = V C, so res = A Kchi - B chi per record
// with A = I and B[j][a] = C[a][j]
bool diag_subspace(const double *uu, const double *vu, int d, Mat8 &A, Mat8 &B) {
  const int e = d + 1;
  double Gvv[kMaxD * kMaxD], Gvk[kMaxD * kMaxD], Gi[kMaxD * kMaxD];
  for (int a = 0; a < d; ++a)
    for (int b = 0; b < d; ++b) {
      Gvv[a * d + b] = uu[a * e + b];
      Gvk[a * d + b] = vu[b * e + a];  // sum_n chi[n,a] Kchi[n,b]
    }
  if (!host_inverse(Gvv, d, Gi)) return false;
  for (int j = 0; j < d; ++j)
    for (int a = 0; a < d; ++a) {
      double s = 0.0;
      for (int k = 0; k < d; ++k) s += Gi[a * d + k] * Gvk[k * d + j];
      A.m[j * d + a] = (a == j) ? 1.0 : 0.0;
      B.m[j * d + a] = s;
    }
  return true;
}

// residual_ritz (src/isotarget.jl:787-799): V = Q R with R'R = V'V; Kr = Q'(KV R^-1) = R^-T (V'KV) R^-1;
// eigen(Kr, sortby = x -> abs(1 - x)); with a = R^-1 vecs[:, j]: residues[:, j] = KV a - lambda_j V a, i.e. per record
// Re = Are Kchi - Bre chi, Im = Aim Kchi - Bim chi.  vals: d x (re, im); vecs: d x d complex column-major interleaved.
bool diag_ritz(const double *uu, const double *vu, int d, double *vals, double *vecs, Mat8 &Are, Mat8 &Bre, Mat8 &Aim,
               Mat8 &Bim, bool *any_complex) {
  const int e = d + 1;
  double Gvv[kMaxD * kMaxD], Gvk[kMaxD * kMaxD], R[kMaxD * kMaxD], Ri[kMaxD * kMaxD], Kr[kMaxD * kMaxD];
  for (int a = 0; a < d; ++a)
    for (int b = 0; b < d; ++b) {
      Gvv[a * d + b] = uu[a * e + b];
      Gvk[a * d + b] = vu[b * e + a];
    }
  if (!host_cholesky_upper(Gvv, d, R) || !host_inverse(R, d, Ri)) return false;
  for (int a = 0; a < d; ++a)
    for (int b = 0; b < d; ++b) {
      double s = 0.0;
      for (int k = 0; k < d; ++k)
        for (int l = 0; l < d; ++l) s += Ri[k * d + a] * Gvk[k * d + l] * Ri[l * d + b];
      Kr[a * d + b] = s;
    }
  double wr[kMaxD], wi[kMaxD], vre[kMaxD * kMaxD], vim[kMaxD * kMaxD];
  if (!host_eig_general(Kr, d, wr, wi, vre, vim)) return false;
  int ord[kMaxD];
  for (int j = 0; j < d; ++j) ord[j] = j;
  std::stable_sort(ord, ord + d, [&](int p, int q) {
    return std::hypot(1.0 - wr[p], wi[p]) < std::hypot(1.0 - wr[q], wi[q]);
  });
  *any_complex = false;
  for (int j = 0; j < d; ++j) {
    const int k = ord[j];
    vals[2 * j] = wr[k];
    vals[2 * j + 1] = wi[k];
    *any_complex = *any_complex || wi[k] != 0.0;
    for (int i = 0; i < d; ++i) {
      if (vecs) {
        vecs[2 * (i + j * d)] = vre[i * d + k];
        vecs[2 * (i + j * d) + 1] = vim[i * d + k];
      }
      double are = 0.0, aim = 0.0;
      for (int l = 0; l < d; ++l) {
        are += Ri[i * d + l] * vre[l * d + k];
        aim += Ri[i * d + l] * vim[l * d + k];
      }
      Are.m[j * d + i] = are;
      Aim.m[j * d + i] = aim;
      Bre.m[j * d + i] = wr[k] * are - wi[k] * aim;
      Bim.m[j * d + i] = wr[k] * aim + wi[k] * are;
    }
  }
  return true;
}

}  // namespace ik
