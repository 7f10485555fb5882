/* C restatement of the ISOKANN featurizer for the CPU oracle (TEST INFRASTRUCTURE ONLY, PARITY UNPINNED --
 * see oracle/__init__.py).  Same arithmetic as oracle/isokann_oracle.py:_dists_from_pairs, which it replaces
 * for large inputs: direct differences in double, max(.,0), sqrt, then the Float32 cast.
 *
 * Reference lines followed (under /root/reference/):
 *   src/utils/pairdists.jl:6-24    flatpairdists: squared distances, upper triangle, max(.,0), sqrt
 *   src/utils/pairdists.jl:109-118 pdists: sqrt(sum((a-b)^2)) over an explicit pair list
 *   src/simulation.jl:112          Float32.(featurizer(coords))
 *
 * Build: gcc -O2 -fPIC -shared -fopenmp oracle/isokann_oracle.c -o oracle/liboracle.so -lm  (oracle/build.py)
 */
#include <math.h>
#include <stdint.h>

/* x: M records of D coordinates (float or double), pairs0: F pairs of 0-based atom indices,
 * out: M x F (float or double) */
#define PDISTS_BODY(TIN, TOUT)                                                  \
  _Pragma("omp parallel for schedule(static)")                                  \
  for (int64_t m = 0; m < M; ++m) {                                             \
    const TIN *r = x + m * D;                                                   \
    TOUT *o = out + m * F;                                                      \
    for (int f = 0; f < F; ++f) {                                               \
      const TIN *a = r + 3 * pairs0[2 * f], *b = r + 3 * pairs0[2 * f + 1];     \
      const double dx = (double)a[0] - (double)b[0];                            \
      const double dy = (double)a[1] - (double)b[1];                            \
      const double dz = (double)a[2] - (double)b[2];                            \
      double sq = dx * dx + dy * dy + dz * dz;                                  \
      if (!(sq > 0.0)) sq = sq != sq ? sq : 0.0; /* max(sq, 0), NaN kept */     \
      o[f] = (TOUT)sqrt(sq);                                                    \
    }                                                                           \
  }

void oracle_pdists_f32_f32(const float *x, int64_t M, int64_t D, const int32_t *pairs0, int F, float *out) {
  PDISTS_BODY(float, float)
}
void oracle_pdists_f64_f32(const double *x, int64_t M, int64_t D, const int32_t *pairs0, int F, float *out) {
  PDISTS_BODY(double, float)
}
void oracle_pdists_f32_f64(const float *x, int64_t M, int64_t D, const int32_t *pairs0, int F, double *out) {
  PDISTS_BODY(float, double)
}
void oracle_pdists_f64_f64(const double *x, int64_t M, int64_t D, const int32_t *pairs0, int F, double *out) {
  PDISTS_BODY(double, double)
}

int32_t oracle_c_version(void) { return 1; }
