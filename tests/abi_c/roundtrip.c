/* C99 caller of libisokann_b200.so through include/isokann_b200.h only (no Python, no torch): what a foreign
 * host -- the reference's Julia ccall shim, julia/ISOKANNB200.jl -- does:
 *   create -> upload_params -> set_data -> iterate (run!(iso, 3)) -> download_params / chis -> destroy
 * Build:  gcc -std=c99 -I include tests/abi_c/roundtrip.c -o roundtrip -L isokann.jl_b200 -lisokann_b200 -lm
 * Prints "ABI_C OK ..." and exits 0 on success. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "isokann_b200.h"

static unsigned long long rng_state = 0x853c49e6748fea9bULL;
static double uniform(void) { /* xorshift64*: deterministic synthetic inputs */
  rng_state ^= rng_state >> 12;
  rng_state ^= rng_state << 25;
  rng_state ^= rng_state >> 27;
  return (double)((rng_state * 2685821657736338717ULL) >> 11) / 9007199254740992.0;
}
static double normal(void) { return sqrt(-2.0 * log(uniform() + 1e-300)) * cos(6.283185307179586 * uniform()); }

#define CHECK(call)                                                                        \
  do {                                                                                     \
    int32_t rc_ = (call);                                                                  \
    if (rc_ != ISOKANN_OK) {                                                               \
      fprintf(stderr, "%s -> status %d: %s\n", #call, (int)rc_, isokann_last_error(ctx)); \
      return 1;                                                                            \
    }                                                                                      \
  } while (0)

int main(void) {
  enum { A = 22, D = 3 * A, F = A * (A - 1) / 2, N = 640, K = 3, H1 = 38, H2 = 6, ITER = 3 };
  isokann_ctx *ctx = NULL;
  isokann_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.n_layers = 3;
  cfg.widths[0] = F; cfg.widths[1] = H1; cfg.widths[2] = H2; cfg.widths[3] = 1;
  cfg.layernorm = 1; cfg.ln_eps = 1e-5f;
  cfg.activation = ISOKANN_ACT_SIGMOID; cfg.last_activation = ISOKANN_ACT_IDENTITY;
  cfg.optimiser = ISOKANN_OPT_ADAM;
  cfg.eta = 1e-3f; cfg.lambda = 1e-4f; cfg.beta1 = 0.9f; cfg.beta2 = 0.999f; cfg.eps = 1e-8f; cfg.rho = 0.9f;
  cfg.featurizer = ISOKANN_FEAT_ALLPAIRS; cfg.n_atoms = A;
  cfg.device = 0; cfg.gemm_mode = ISOKANN_GEMM_AUTO; cfg.chunk = 0;
  if (isokann_abi_version() != ISOKANN_ABI_VERSION) { fprintf(stderr, "ABI version mismatch\n"); return 1; }
  if (isokann_create(&cfg, &ctx) != ISOKANN_OK) { fprintf(stderr, "create: %s\n", isokann_last_error(NULL)); return 1; }

  const int64_t P = isokann_num_params(ctx);
  if (P != 2 * F + (F + 1) * H1 + (H1 + 1) * H2 + (H2 + 1) || isokann_feature_dim(ctx) != F || isokann_coord_dim(ctx) != D) {
    fprintf(stderr, "unexpected sizes P=%lld\n", (long long)P);
    return 1;
  }
  /* parameters in the flat Functors order: LayerNorm scale (1), bias (0), then glorot-uniform W, zero b per layer */
  float *flat = (float *)calloc((size_t)P, sizeof(float)), *flat1 = (float *)calloc((size_t)P, sizeof(float));
  int64_t o = 0;
  for (int i = 0; i < F; ++i) flat[o++] = 1.f;
  o += F;
  const int w[4] = {F, H1, H2, 1};
  for (int l = 0; l < 3; ++l) {
    const double lim = sqrt(6.0 / (w[l] + w[l + 1]));
    for (int i = 0; i < w[l] * w[l + 1]; ++i) flat[o++] = (float)((2.0 * uniform() - 1.0) * lim);
    o += w[l + 1];
  }
  CHECK(isokann_upload_params(ctx, flat, P));

  /* two conformations of a random chain + noise; ys = xs + noise (column-major D x N and D x K x N) */
  float *xs = (float *)malloc(sizeof(float) * D * N), *ys = (float *)malloc(sizeof(float) * D * K * N);
  float base[2][D];
  for (int s = 0; s < 2; ++s)
    for (int i = 0; i < D; ++i) base[s][i] = (float)(0.4 * normal() * (s + 1));
  for (int n = 0; n < N; ++n) {
    const int s = uniform() < 0.5;
    for (int i = 0; i < D; ++i) xs[(size_t)n * D + i] = base[s][i] + 0.05f * (float)normal();
    for (int k = 0; k < K; ++k)
      for (int i = 0; i < D; ++i) ys[((size_t)n * K + k) * D + i] = xs[(size_t)n * D + i] + 0.03f * (float)normal();
  }
  CHECK(isokann_set_data(ctx, xs, ys, D, K, N));

  /* ITER permutations (1-based), Fisher-Yates */
  int64_t *perms = (int64_t *)malloc(sizeof(int64_t) * ITER * N);
  for (int it = 0; it < ITER; ++it) {
    int64_t *p = perms + (size_t)it * N;
    for (int i = 0; i < N; ++i) p[i] = i + 1;
    for (int i = N - 1; i > 0; --i) {
      const int j = (int)(uniform() * (i + 1));
      const int64_t t = p[i]; p[i] = p[j]; p[j] = t;
    }
  }
  double losses[ITER];
  CHECK(isokann_iterate(ctx, ISOKANN_TARGET_SHIFTSCALE, NULL, ITER, 1, 128, perms, losses));
  CHECK(isokann_download_params(ctx, flat1, P));
  float *chi = (float *)malloc(sizeof(float) * N), *tgt = (float *)malloc(sizeof(float) * N);
  CHECK(isokann_chis(ctx, chi));
  CHECK(isokann_download_target(ctx, tgt));
  isokann_stats st;
  CHECK(isokann_get_stats(ctx, &st));

  double moved = 0.0, tmin = 1e30, tmax = -1e30;
  for (int64_t i = 0; i < P; ++i) moved = fmax(moved, fabs((double)flat1[i] - (double)flat[i]));
  for (int n = 0; n < N; ++n) {
    if (!isfinite(chi[n])) { fprintf(stderr, "non-finite chi\n"); return 1; }
    tmin = fmin(tmin, tgt[n]); tmax = fmax(tmax, tgt[n]);
  }
  for (int it = 0; it < ITER; ++it)
    if (!isfinite(losses[it]) || losses[it] <= 0.0) { fprintf(stderr, "bad loss %g\n", losses[it]); return 1; }
  if (!(moved > 1e-4 && moved < 1e-1)) { fprintf(stderr, "parameters moved by %g\n", moved); return 1; }
  if (tmin != 0.0 || tmax != 1.0) { fprintf(stderr, "shiftscale target range [%g, %g]\n", tmin, tmax); return 1; }
  if (st.kernel_launches <= 0) { fprintf(stderr, "no kernels were launched\n"); return 1; }
  /* an out-of-range permutation is rejected with BAD_ARGUMENT and leaves the context usable */
  perms[0] = N + 7;
  if (isokann_train_epoch(ctx, perms, 128, 0, NULL) != ISOKANN_BAD_ARGUMENT) { fprintf(stderr, "bad perm accepted\n"); return 1; }
  perms[0] = 1;
  printf("ABI_C OK losses %.6g %.6g %.6g launches %lld moved %.3g\n", losses[0], losses[1], losses[2],
         (long long)st.kernel_launches, moved);
  CHECK(isokann_destroy(ctx));
  free(flat); free(flat1); free(xs); free(ys); free(perms); free(chi); free(tgt);
  return 0;
}
