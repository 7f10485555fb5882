// Host-side dense algebra on d x d matrices (d <= 8) for the N-D isotarget transforms.
//
// The reference calls LAPACK through Julia's LinearAlgebra for these tiny problems
// (inv: src/isotarget.jl:93, schur: :167, C^(-1/2): :86).  The O(N) work lives in the CUDA
// reductions; what is left is O(d^3) and runs here.  The real Schur factorisation follows
// LAPACK's sgees pipeline for small matrices (sgehd2 -> sorg2r -> slahqr with slanv2
// standardisation, no balancing permutation for dense input, no eigenvalue sorting) in
// single precision so that the Schur-vector sign/order conventions agree with the
// reference's `schur(Kinv).vectors`.
#include <cmath>
#include <cstring>
#include <algorithm>

#include "common.cuh"

namespace ik {

bool host_inverse(const double *a, int d, double *inv) {
  double m[kMaxD][2 * kMaxD];
  for (int i = 0; i < d; ++i) {
    for (int j = 0; j < d; ++j) {
      m[i][j] = a[i * d + j];
      m[i][d + j] = (i == j) ? 1.0 : 0.0;
      if (!std::isfinite(m[i][j])) return false;
    }
  }
  for (int col = 0; col < d; ++col) {
    int piv = col;
    double best = std::fabs(m[col][col]);
    for (int r = col + 1; r < d; ++r)
      if (std::fabs(m[r][col]) > best) {
        best = std::fabs(m[r][col]);
        piv = r;
      }
    if (best == 0.0) return false;  // exactly singular -> SingularException in the reference
    if (piv != col)
      for (int j = 0; j < 2 * d; ++j) std::swap(m[piv][j], m[col][j]);
    const double p = m[col][col];
    for (int j = 0; j < 2 * d; ++j) m[col][j] /= p;
    for (int r = 0; r < d; ++r) {
      if (r == col) continue;
      const double f = m[r][col];
      if (f == 0.0) continue;
      for (int j = 0; j < 2 * d; ++j) m[r][j] -= f * m[col][j];
    }
  }
  for (int i = 0; i < d; ++i)
    for (int j = 0; j < d; ++j) {
      inv[i * d + j] = m[i][d + j];
      if (!std::isfinite(inv[i * d + j])) return false;
    }
  return true;
}

// cyclic Jacobi for symmetric matrices; evecs columns are eigenvectors
void host_sym_eig(const double *a_in, int d, double *evals, double *evecs) {
  double a[kMaxD][kMaxD], v[kMaxD][kMaxD];
  for (int i = 0; i < d; ++i)
    for (int j = 0; j < d; ++j) {
      a[i][j] = a_in[i * d + j];
      v[i][j] = (i == j) ? 1.0 : 0.0;
    }
  for (int sweep = 0; sweep < 100; ++sweep) {
    double off = 0.0;
    for (int i = 0; i < d; ++i)
      for (int j = i + 1; j < d; ++j) off += a[i][j] * a[i][j];
    if (off < 1e-300) break;
    for (int p = 0; p < d; ++p)
      for (int q = p + 1; q < d; ++q) {
        if (a[p][q] == 0.0) continue;
        const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double cs = 1.0 / std::sqrt(t * t + 1.0), sn = t * cs;
        for (int k = 0; k < d; ++k) {
          const double akp = a[k][p], akq = a[k][q];
          a[k][p] = cs * akp - sn * akq;
          a[k][q] = sn * akp + cs * akq;
        }
        for (int k = 0; k < d; ++k) {
          const double apk = a[p][k], aqk = a[q][k];
          a[p][k] = cs * apk - sn * aqk;
          a[q][k] = sn * apk + cs * aqk;
        }
        for (int k = 0; k < d; ++k) {
          const double vkp = v[k][p], vkq = v[k][q];
          v[k][p] = cs * vkp - sn * vkq;
          v[k][q] = sn * vkp + cs * vkq;
        }
      }
  }
  for (int i = 0; i < d; ++i) {
    evals[i] = a[i][i];
    for (int j = 0; j < d; ++j) evecs[i * d + j] = v[i][j];
  }
}

// ------------------------------------------------------------------------------------------
// real Schur, LAPACK conventions, templated on the scalar type (float: sgees parity with the reference's
// schur(Kinv); double: eigenvalues for the diagnostics).  Matrices are column-major with leading
// dimension n, indices in the helpers are 0-based.
// ------------------------------------------------------------------------------------------
namespace {

template <typename T>
struct MatT {
  T *p;
  int n;
  T &operator()(int i, int j) { return p[i + (size_t)j * n]; }
};

// machine constants as LAPACK's xLAMCH returns them: 'P' (eps * base) and 'S' (safe minimum)
template <typename T> struct Lim;
template <> struct Lim<float> {
  static constexpr float ulp = 1.1920929e-07f, safmin = 1.17549435e-38f;
};
template <> struct Lim<double> {
  static constexpr double ulp = 2.220446049250313e-16, safmin = 2.2250738585072014e-308;
};

template <typename T>
inline T sgn(T a, T b) { return b >= T(0.) ? std::fabs(a) : -std::fabs(a); }  // Fortran SIGN(a,b)
template <typename T>
inline T slapy2(T x, T y) {
  const T xa = std::fabs(x), ya = std::fabs(y);
  const T w = std::max(xa, ya), z = std::min(xa, ya);
  if (z == T(0.)) return w;
  const T q = z / w;
  return w * std::sqrt(T(1.) + q * q);
}

// slarfg: on entry alpha, x[0..n-2]; on exit alpha=beta, x=v(2:n), returns tau
template <typename T>
T slarfg(int n, T &alpha, T *x, int incx) {
  if (n <= 1) return T(0.);
  double ss = 0.0;
  for (int i = 0; i < n - 1; ++i) ss += (double)x[i * incx] * (double)x[i * incx];
  const T xnorm = (T)std::sqrt(ss);
  if (xnorm == T(0.)) return T(0.);
  const T beta = -sgn(slapy2(alpha, xnorm), alpha);
  const T tau = (beta - alpha) / beta;
  const T sc = T(1.) / (alpha - beta);
  for (int i = 0; i < n - 1; ++i) x[i * incx] *= sc;
  alpha = beta;
  return tau;
}

template <typename T>
void slanv2(T &a, T &b, T &c, T &d, T &cs, T &sn) {
  const T eps = Lim<T>::ulp;  // xLAMCH('P')
  const T multpl = T(4.);
  if (c == T(0.)) {
    cs = T(1.); sn = T(0.);
  } else if (b == T(0.)) {
    cs = T(0.); sn = T(1.);
    const T temp = d;
    d = a; a = temp; b = -c; c = T(0.);
  } else if ((a - d) == T(0.) && sgn(T(1.), b) != sgn(T(1.), c)) {
    cs = T(1.); sn = T(0.);
  } else {
    T temp = a - d;
    T p = T(0.5) * temp;
    const T bcmax = std::max(std::fabs(b), std::fabs(c));
    const T bcmis = std::min(std::fabs(b), std::fabs(c)) * sgn(T(1.), b) * sgn(T(1.), c);
    const T scale = std::max(std::fabs(p), bcmax);
    T z = (p / scale) * p + (bcmax / scale) * bcmis;
    if (z >= multpl * eps) {
      z = p + sgn(std::sqrt(scale) * std::sqrt(z), p);
      a = d + z;
      d = d - (bcmax / z) * bcmis;
      const T tau = slapy2(c, z);
      cs = z / tau;
      sn = c / tau;
      b = b - c;
      c = T(0.);
    } else {
      const T sigma = b + c;
      const T tau = slapy2(sigma, temp);
      cs = std::sqrt(T(0.5) * (T(1.) + std::fabs(sigma) / tau));
      sn = -(p / (tau * cs)) * sgn(T(1.), sigma);
      const T aa = a * cs + b * sn, bb = -a * sn + b * cs;
      const T cc = c * cs + d * sn, dd = -c * sn + d * cs;
      a = aa * cs + cc * sn;
      b = bb * cs + dd * sn;
      c = -aa * sn + cc * cs;
      d = -bb * sn + dd * cs;
      temp = T(0.5) * (a + d);
      a = temp;
      d = temp;
      if (c != T(0.)) {
        if (b != T(0.)) {
          if (sgn(T(1.), b) == sgn(T(1.), c)) {
            const T sab = std::sqrt(std::fabs(b)), sac = std::sqrt(std::fabs(c));
            p = sgn(sab * sac, c);
            const T tau2 = T(1.) / std::sqrt(std::fabs(b + c));
            a = temp + p;
            d = temp - p;
            b = b - c;
            c = T(0.);
            const T cs1 = sab * tau2, sn1 = sac * tau2;
            temp = cs * cs1 - sn * sn1;
            sn = cs * sn1 + sn * cs1;
            cs = temp;
          }
        } else {
          b = -c;
          c = T(0.);
          temp = cs;
          cs = -sn;
          sn = temp;
        }
      }
    }
  }
}

template <typename T>
inline void srot(int n, T *x, int incx, T *y, int incy, T c, T s) {
  for (int i = 0; i < n; ++i) {
    const T t = c * x[i * incx] + s * y[i * incy];
    y[i * incy] = c * y[i * incy] - s * x[i * incx];
    x[i * incx] = t;
  }
}

// slahqr with wantt = wantz = true on the full matrix (ilo = 1, ihi = n).  Returns 0 on success.
template <typename T>
int slahqr(MatT<T> H, MatT<T> Z) {
  const int n = H.n;
  if (n == 0) return 0;
  if (n == 1) return 0;
  for (int j = 0; j < n - 3; ++j) {
    H(j + 2, j) = T(0.);
    H(j + 3, j) = T(0.);
  }
  if (n >= 3) H(n - 1, n - 3) = T(0.);
  const T safmin = Lim<T>::safmin;
  const T ulp = Lim<T>::ulp;
  const T smlnum = safmin * ((T)n / ulp);
  const T dat1 = T(0.75), dat2 = T(-0.4375);
  const int kexsh = 10;
  const int itmax = 30 * std::max(10, n);
  int kdefl = 0;
  const int i1 = 0, i2 = n - 1;
  int i = n - 1;
  while (i >= 0) {
    int l = 0;
    bool converged = false;
    for (int its = 0; its <= itmax; ++its) {
      int k;
      for (k = i; k > l; --k) {
        if (std::fabs(H(k, k - 1)) <= smlnum) break;
        T tst = std::fabs(H(k - 1, k - 1)) + std::fabs(H(k, k));
        if (tst == T(0.)) {
          if (k - 2 >= 0) tst += std::fabs(H(k - 1, k - 2));
          if (k + 1 <= n - 1) tst += std::fabs(H(k + 1, k));
        }
        if (std::fabs(H(k, k - 1)) <= ulp * tst) {
          const T ab = std::max(std::fabs(H(k, k - 1)), std::fabs(H(k - 1, k)));
          const T ba = std::min(std::fabs(H(k, k - 1)), std::fabs(H(k - 1, k)));
          const T aa = std::max(std::fabs(H(k, k)), std::fabs(H(k - 1, k - 1) - H(k, k)));
          const T bb = std::min(std::fabs(H(k, k)), std::fabs(H(k - 1, k - 1) - H(k, k)));
          const T s = aa + ab;
          if (ba * (ab / s) <= std::max(smlnum, ulp * (bb * (aa / s)))) break;
        }
      }
      l = k;
      if (l > 0) H(l, l - 1) = T(0.);
      if (l >= i - 1) {
        converged = true;
        break;
      }
      kdefl++;
      T h11, h21, h12, h22;
      if (kdefl % (2 * kexsh) == 0) {
        const T s = std::fabs(H(i, i - 1)) + std::fabs(H(i - 1, i - 2));
        h11 = dat1 * s + H(i, i);
        h12 = dat2 * s;
        h21 = s;
        h22 = h11;
      } else if (kdefl % kexsh == 0) {
        const T s = std::fabs(H(l + 1, l)) + std::fabs(H(l + 2, l + 1));
        h11 = dat1 * s + H(l, l);
        h12 = dat2 * s;
        h21 = s;
        h22 = h11;
      } else {
        h11 = H(i - 1, i - 1);
        h21 = H(i, i - 1);
        h12 = H(i - 1, i);
        h22 = H(i, i);
      }
      T rt1r, rt1i, rt2r, rt2i;
      T s = std::fabs(h11) + std::fabs(h12) + std::fabs(h21) + std::fabs(h22);
      if (s == T(0.)) {
        rt1r = rt1i = rt2r = rt2i = T(0.);
      } else {
        h11 /= s; h21 /= s; h12 /= s; h22 /= s;
        const T tr = (h11 + h22) / T(2.);
        const T det = (h11 - tr) * (h22 - tr) - h12 * h21;
        const T rtdisc = std::sqrt(std::fabs(det));
        if (det >= T(0.)) {
          rt1r = tr * s; rt2r = rt1r; rt1i = rtdisc * s; rt2i = -rt1i;
        } else {
          rt1r = tr + rtdisc;
          rt2r = tr - rtdisc;
          if (std::fabs(rt1r - h22) <= std::fabs(rt2r - h22)) {
            rt1r = rt1r * s; rt2r = rt1r;
          } else {
            rt2r = rt2r * s; rt1r = rt2r;
          }
          rt1i = rt2i = T(0.);
        }
      }
      T v[3];
      int m;
      for (m = i - 2; m >= l; --m) {
        T h21s = std::fabs(H(m + 1, m));
        s = std::fabs(H(m, m) - rt2r) + std::fabs(rt2i) + h21s;
        h21s = H(m + 1, m) / s;
        v[0] = h21s * H(m, m + 1) + (H(m, m) - rt1r) * ((H(m, m) - rt2r) / s) - rt1i * (rt2i / s);
        v[1] = h21s * (H(m, m) + H(m + 1, m + 1) - rt1r - rt2r);
        v[2] = h21s * H(m + 2, m + 1);
        s = std::fabs(v[0]) + std::fabs(v[1]) + std::fabs(v[2]);
        v[0] /= s; v[1] /= s; v[2] /= s;
        if (m == l) break;
        const T h00 = std::fabs(H(m - 1, m - 1));
        const T h10 = std::fabs(H(m, m - 1));
        const T h11a = std::fabs(H(m, m));
        const T h22a = std::fabs(H(m + 1, m + 1));
        if (h10 * (std::fabs(v[1]) + std::fabs(v[2])) <= ulp * std::fabs(v[0]) * (h00 + h11a + h22a)) break;
      }
      for (int k2 = m; k2 <= i - 1; ++k2) {
        const int nr = std::min(3, i - k2 + 1);
        if (k2 > m)
          for (int q = 0; q < nr; ++q) v[q] = H(k2 + q, k2 - 1);
        const T t1 = slarfg(nr, v[0], &v[1], 1);
        if (k2 > m) {
          H(k2, k2 - 1) = v[0];
          H(k2 + 1, k2 - 1) = T(0.);
          if (k2 < i - 1) H(k2 + 2, k2 - 1) = T(0.);
        } else if (m > l) {
          H(k2, k2 - 1) = H(k2, k2 - 1) * (T(1.) - t1);
        }
        const T v2 = v[1];
        const T t2 = t1 * v2;
        if (nr == 3) {
          const T v3 = v[2];
          const T t3 = t1 * v3;
          for (int j = k2; j <= i2; ++j) {
            const T sum = H(k2, j) + v2 * H(k2 + 1, j) + v3 * H(k2 + 2, j);
            H(k2, j) -= sum * t1;
            H(k2 + 1, j) -= sum * t2;
            H(k2 + 2, j) -= sum * t3;
          }
          for (int j = i1; j <= std::min(k2 + 3, i); ++j) {
            const T sum = H(j, k2) + v2 * H(j, k2 + 1) + v3 * H(j, k2 + 2);
            H(j, k2) -= sum * t1;
            H(j, k2 + 1) -= sum * t2;
            H(j, k2 + 2) -= sum * t3;
          }
          for (int j = 0; j < n; ++j) {
            const T sum = Z(j, k2) + v2 * Z(j, k2 + 1) + v3 * Z(j, k2 + 2);
            Z(j, k2) -= sum * t1;
            Z(j, k2 + 1) -= sum * t2;
            Z(j, k2 + 2) -= sum * t3;
          }
        } else if (nr == 2) {
          for (int j = k2; j <= i2; ++j) {
            const T sum = H(k2, j) + v2 * H(k2 + 1, j);
            H(k2, j) -= sum * t1;
            H(k2 + 1, j) -= sum * t2;
          }
          for (int j = i1; j <= i; ++j) {
            const T sum = H(j, k2) + v2 * H(j, k2 + 1);
            H(j, k2) -= sum * t1;
            H(j, k2 + 1) -= sum * t2;
          }
          for (int j = 0; j < n; ++j) {
            const T sum = Z(j, k2) + v2 * Z(j, k2 + 1);
            Z(j, k2) -= sum * t1;
            Z(j, k2 + 1) -= sum * t2;
          }
        }
      }
    }
    if (!converged) return i + 1;
    if (l == i - 1) {
      T cs, sn;
      slanv2(H(i - 1, i - 1), H(i - 1, i), H(i, i - 1), H(i, i), cs, sn);
      if (i2 > i) srot(i2 - i, &H(i - 1, i + 1), n, &H(i, i + 1), n, cs, sn);
      srot(i - i1 - 1, &H(i1, i - 1), 1, &H(i1, i), 1, cs, sn);
      srot(n, &Z(0, i - 1), 1, &Z(0, i), 1, cs, sn);
    }
    kdefl = 0;
    i = l - 1;
  }
  return 0;
}

template <typename T>
bool host_schur_t(const T *a_colmajor, int n, T *z_out, T *t_out) {
  if (n < 1 || n > kMaxD) return false;
  T hbuf[kMaxD * kMaxD], zbuf[kMaxD * kMaxD], tau[kMaxD];
  for (int i = 0; i < n * n; ++i) {
    if (!std::isfinite(a_colmajor[i])) return false;
    hbuf[i] = a_colmajor[i];
  }
  MatT<T> A{hbuf, n}, Q{zbuf, n};
  // ---- sgehd2: A <- Q^T A Q upper Hessenberg, reflectors stored below the subdiagonal ----
  for (int i = 0; i < n - 1; ++i) {
    // reflector H(i) annihilates A(i+2:n-1, i)
    const int len = n - 1 - i;  // ihi - i in 1-based terms
    T alpha = A(i + 1, i);
    tau[i] = slarfg(len, alpha, &A(std::min(i + 2, n - 1), i), 1);
    A(i + 1, i) = T(1.);
    T *vv = &A(i + 1, i);
    // apply H(i) from the right to A(0:n-1, i+1:n-1)
    for (int r = 0; r < n; ++r) {
      T w = T(0.);
      for (int q = 0; q < len; ++q) w += A(r, i + 1 + q) * vv[q];
      const T tw = tau[i] * w;
      if (tw != T(0.))
        for (int q = 0; q < len; ++q) A(r, i + 1 + q) -= tw * vv[q];
    }
    // apply H(i)^T from the left to A(i+1:n-1, i+1:n-1)
    for (int cidx = i + 1; cidx < n; ++cidx) {
      T w = T(0.);
      for (int q = 0; q < len; ++q) w += A(i + 1 + q, cidx) * vv[q];
      const T tw = tau[i] * w;
      if (tw != T(0.))
        for (int q = 0; q < len; ++q) A(i + 1 + q, cidx) -= tw * vv[q];
    }
    A(i + 1, i) = alpha;
  }
  // ---- sorghr: Q = H(0) H(1) ... H(n-2) via sorg2r on the trailing (n-1)x(n-1) block ----
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) Q(i, j) = (i == j) ? T(1.) : T(0.);
  {
    const int m = n - 1;  // order of the trailing block; reflector k has v = [1; A(k+2:n-1, k)]
    // Qs(r, c) = Q(r+1, c+1); column c initially holds reflector vector c below its diagonal
    for (int c2 = 0; c2 < m; ++c2)
      for (int r = c2 + 1; r < m; ++r) Q(r + 1, c2 + 1) = A(r + 1, c2);
    for (int k = m - 1; k >= 0; --k) {
      if (k < m - 1) {
        Q(k + 1, k + 1) = T(1.);
        // apply H(k) from the left to Qs(k:m-1, k+1:m-1); v = Qs(k:m-1, k)
        for (int cidx = k + 1; cidx < m; ++cidx) {
          T w = T(0.);
          for (int r = k; r < m; ++r) w += Q(r + 1, cidx + 1) * Q(r + 1, k + 1);
          const T tw = tau[k] * w;
          if (tw != T(0.))
            for (int r = k; r < m; ++r) Q(r + 1, cidx + 1) -= tw * Q(r + 1, k + 1);
        }
      }
      if (k < m - 1)
        for (int r = k + 1; r < m; ++r) Q(r + 1, k + 1) *= -tau[k];
      Q(k + 1, k + 1) = T(1.) - tau[k];
      for (int r = 0; r < k; ++r) Q(r + 1, k + 1) = T(0.);
    }
  }
  // zero the reflector storage below the subdiagonal, then QR-iterate
  for (int j = 0; j < n; ++j)
    for (int i = j + 2; i < n; ++i) A(i, j) = T(0.);
  if (slahqr(A, Q) != 0) return false;
  for (int j = 0; j < n; ++j)
    for (int i = j + 2; i < n; ++i) A(i, j) = T(0.);
  std::memcpy(z_out, zbuf, sizeof(T) * n * n);
  if (t_out) std::memcpy(t_out, hbuf, sizeof(T) * n * n);
  return true;
}

}  // namespace

bool host_schur_f32(const float *a_colmajor, int n, float *z_out, float *t_out) {
  return host_schur_t<float>(a_colmajor, n, z_out, t_out);
}
// the same pipeline in double precision (dgehd2 -> dorg2r -> dlahqr): eigenvalues of the diagnostics' d x d matrices
bool host_schur_f64(const double *a_colmajor, int n, double *z_out, double *t_out) {
  return host_schur_t<double>(a_colmajor, n, z_out, t_out);
}

}  // namespace ik
