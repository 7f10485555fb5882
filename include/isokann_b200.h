/* libisokann_b200.so -- C ABI of the B200-native ISOKANN per-iteration hot path.
 *
 * Drop-in boundary for axsk/ISOKANN.jl.  The reference has no FFI: its plugin API is
 * Julia multiple dispatch.  Each entry point below names the reference interface it
 * replaces (file:line under the reference checkout).  The Julia-side `ccall` shim is
 * julia/ISOKANNB200.jl; in this repo the same ABI is driven from
 * Python ctypes (isokann.jl_b200/lib.py).
 *
 * Conventions
 *  - every function returns an int32 status (ISOKANN_OK == 0); no C++ exception crosses
 *    the boundary; isokann_last_error(ctx) gives the message of the last failure.
 *  - pointers are HOST pointers to caller-owned, contiguous, COLUMN-MAJOR (Julia) arrays
 *    unless the parameter name starts with `dev_`; the library copies in/out and owns all
 *    device memory behind the opaque handle.
 *      xs[D,N]   : N records of D floats          ys[D,K,N] : for each n, K records of D floats
 *      chi[d,N]  : N records of d floats          W[out,in] : Flux.Dense weight, column-major
 *  - flat parameter order = Functors traversal of the Flux.Chain:
 *      [LayerNorm.scale(F), LayerNorm.bias(F),]  W1(out x in col-major), b1, W2, b2, ...
 *  - index arguments (perm, atom indices, pairs) are 1-based, as Julia produces them.
 *  - single caller: one host thread drives one context; the library is not re-entrant.
 *  - there is NO CPU fallback: without a CUDA device isokann_create fails with
 *    ISOKANN_ERR_CUDA.
 */
#ifndef ISOKANN_B200_H
#define ISOKANN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ISOKANN_ABI_VERSION 3
#define ISOKANN_MAX_LAYERS 8

/* status codes; 1..4 are the reference's DomainErrors and must be re-thrown as such */
enum {
  ISOKANN_OK = 0,
  ISOKANN_DOMAIN_CONSTANT_CHI = 1,     /* src/isotarget.jl:39  "chi function is constant"            */
  ISOKANN_DOMAIN_NONFINITE_LOSS = 2,   /* src/iso.jl:186-189   "model collapsed under training"      */
  ISOKANN_DOMAIN_SINGULAR_SIMPLEX = 3, /* src/isotarget.jl:94-97 "simplex transformation"            */
  ISOKANN_DOMAIN_PINV = 4,             /* src/isotarget.jl:159-163 "pseudoinverse"                   */
  ISOKANN_BAD_ARGUMENT = 5,            /* e.g. shiftscale with d>1 (src/isotarget.jl:37)             */
  ISOKANN_ERR_CUDA = 10,
  ISOKANN_ERR_NCCL = 11,
  ISOKANN_ERR_STATE = 12               /* call order (no data / no target yet)                       */
};

enum { ISOKANN_ACT_IDENTITY = 0, ISOKANN_ACT_SIGMOID = 1, ISOKANN_ACT_TANH = 2, ISOKANN_ACT_RELU = 3 };
enum { ISOKANN_OPT_NESTEROV = 0, ISOKANN_OPT_ADAM = 1 };
/* featurizer functors, src/utils/features.jl:18-35 */
enum {
  ISOKANN_FEAT_IDENTITY = 0, /* FeaturesCoords / `identity` (ExternalSimulation, Langevin) */
  ISOKANN_FEAT_ALLPAIRS = 1, /* FeaturesAll   -> flatpairdists(x)        src/utils/pairdists.jl:6-24  */
  ISOKANN_FEAT_ATOMS = 2,    /* FeaturesAtoms -> flatpairdists(x, inds)  src/utils/pairdists.jl:13-16 */
  ISOKANN_FEAT_PAIRS = 3     /* FeaturesPairs -> pdists(x, pairs)        src/utils/pairdists.jl:109-127 */
};
/* isotarget transforms, src/isotarget.jl:32,74,145 */
enum { ISOKANN_TARGET_SHIFTSCALE = 0, ISOKANN_TARGET_ISA = 1, ISOKANN_TARGET_PINV = 2 };
/* GEMM engine for the Dense layers */
enum {
  ISOKANN_GEMM_AUTO = 0,   /* tcgen05 for wide layers (in,out >= 256), FP32 CUDA cores otherwise */
  ISOKANN_GEMM_FP32 = 1,   /* FP32 FFMA everywhere                                               */
  ISOKANN_GEMM_TC = 2      /* tcgen05 3xBF16 split (fp32-accurate) wherever shapes allow          */
};

typedef struct isokann_ctx isokann_ctx;

/* Model / optimiser / featurizer description: replaces pairnet/densenet/smallnet
 * (src/models.jl:65-69,87-92,102-108), AdamRegularized/NesterovRegularized (src/models.jl:12,20)
 * + Flux.setup (src/iso.jl:27) and the featurizer functor (src/utils/features.jl:18-35). */
typedef struct {
  int32_t n_layers;                        /* number of Dense layers L                          */
  int32_t widths[ISOKANN_MAX_LAYERS + 1];  /* F, h1, ..., d                                      */
  int32_t layernorm;                       /* 1: Flux.LayerNorm(F) in front (src/models.jl:90)   */
  float ln_eps;                            /* Flux default 1f-5: (x-mu)/sqrt(var+eps^2)          */
  int32_t activation;                      /* hidden layers (Flux.sigmoid)                       */
  int32_t last_activation;                 /* last layer (identity)                              */
  int32_t optimiser;                       /* ISOKANN_OPT_*                                      */
  float eta, lambda, beta1, beta2, eps, rho;
  int32_t featurizer;                      /* ISOKANN_FEAT_*                                     */
  int32_t n_atoms;                         /* A; a coordinate record has D = 3A floats           */
  int32_t n_index;                         /* ATOMS: #atoms; PAIRS: #pairs                       */
  const int32_t *index;                    /* ATOMS: 1-based atom ids; PAIRS: (a1,b1,a2,b2,...)  */
  int32_t device;                          /* CUDA device ordinal                                */
  int32_t gemm_mode;                       /* ISOKANN_GEMM_*                                     */
  int64_t chunk;                           /* samples per forward chunk (0 = library default)    */
} isokann_config;

/* options of TransformISA (src/isotarget.jl:74-77) and TransformPseudoInv (:145-150) */
typedef struct {
  int32_t permute;    /* both, default 1 */
  int32_t whitening;  /* ISA, default 0 */
  int32_t normalize;  /* PseudoInv, default 1 */
  int32_t direct;     /* PseudoInv, default 1 */
  int32_t eigenvecs;  /* PseudoInv, default 1 */
} isokann_target_opts;

typedef struct {
  int64_t kernel_launches;   /* kernels launched by this library since create / reset         */
  int64_t nccl_calls;
  double ms_featurize;       /* accumulated CUDA-event times per kernel class (only while      */
  double ms_gemm;            /* timing is enabled with isokann_enable_timing)                  */
  double ms_reduce;
  double ms_train_elementwise;
  double ms_optimiser;
  double ms_koopman_total;   /* whole expectation(model, ys) pass                              */
  double ms_target_total;
  double ms_train_total;
  int64_t n_gemm_launches;
  int64_t n_featurize_launches;
  double gemm_flops;         /* algorithmic 2*M*N*K of the GEMMs timed in ms_gemm             */
  double featurize_bytes;    /* algorithmic 4*(D+F)*M of the launches timed in ms_featurize   */
  double ms_nccl;            /* gradient all-reduces (includes waiting for the slowest rank)  */
  int64_t graph_launches;    /* training epochs replayed as one captured CUDA graph; their kernels are
                                counted in kernel_launches as well                                */
  double gemm_mma_flops;     /* flops the tensor pipe executed for gemm_flops (2 or 3 MMAs per product)   */
  int64_t p2p_exchanges;     /* gradient all-reduces done by the library's own NVLink peer-memory kernel
                                (single node, CUDA IPC) instead of ncclAllReduce                     */
} isokann_stats;

int32_t isokann_abi_version(void);

/* Iso(data; model, opt) -> handle (src/iso.jl:17-43).  Parameters are zero until uploaded. */
int32_t isokann_create(const isokann_config *cfg, isokann_ctx **out);
int32_t isokann_destroy(isokann_ctx *ctx);                /* the Julia finalizer of the model wrapper */
const char *isokann_last_error(const isokann_ctx *ctx);   /* message of the DomainError / error to throw */

/* length of the flat parameter vector (sum(length, Flux.trainables(model))); inputdim(model) (src/models.jl:26-27)
 * as implied by the featurizer; size(coords, 1) the featurizer expects */
int64_t isokann_num_params(const isokann_ctx *ctx);
int32_t isokann_feature_dim(const isokann_ctx *ctx);  /* F implied by the featurizer          */
int32_t isokann_coord_dim(const isokann_ctx *ctx);    /* D expected in coordinate records      */

/* Multi-GPU: one process per GPU.  Rank 0 obtains the 128-byte NCCL id, the host broadcasts
 * it (torch.distributed / MPI / Distributed.jl) and every rank joins.  Not in the reference
 * (no multi-GPU there); numbers must equal the 1-GPU run up to fp32 summation order. */
int32_t isokann_comm_get_unique_id(void *id128);
int32_t isokann_comm_init(isokann_ctx *ctx, int32_t world, int32_t rank, const void *id128);

/* SimulationData(xs, ys; featurizer) (src/simulation.jl:100-114): uploads coordinates; the
 * Float32 feature cast of :112 happens on device.  ys may be NULL (inference only). */
int32_t isokann_set_data(isokann_ctx *ctx, const float *xs, const float *ys, int64_t D, int64_t K, int64_t N);
int32_t isokann_set_data_f64(isokann_ctx *ctx, const double *xs, const double *ys, int64_t D, int64_t K, int64_t N);
/* Sharded form: xs is the full D x N on every rank, ys_local holds the start points
 * [n_offset, n_offset+n_local) of this rank only. */
int32_t isokann_set_data_sharded(isokann_ctx *ctx, const float *xs, const float *ys_local, int64_t D, int64_t K,
                                 int64_t N, int64_t n_offset, int64_t n_local);
/* Asynchronous form of isokann_set_data_sharded: ys_local, then xs, are streamed to the device on a second
 * CUDA stream.  The next Koopman pass (isokann_koopman / isokann_target / isokann_iterate) consumes ys chunk
 * by chunk as it arrives and the first reader of xs (isokann_chis, training, an N-D target) waits for xs on
 * the device, so the PCIe transfer overlaps the compute.  The library page-locks both buffers itself
 * (cudaHostRegister; cached, so handing over the same arrays again costs nothing; buffers that are already
 * page-locked are used as they are).  They must stay valid and unmodified until isokann_synchronize (or a call
 * that returns results computed from them, such as isokann_iterate) has returned, and must not be freed while
 * registered: call isokann_release_host_buffers first (isokann_destroy and a later isokann_set_data_async with
 * other buffers release them too). */
int32_t isokann_set_data_async(isokann_ctx *ctx, const float *xs, const float *ys_local, int64_t D, int64_t K,
                               int64_t N, int64_t n_offset, int64_t n_local);
int32_t isokann_release_host_buffers(isokann_ctx *ctx);
/* addcoords!(iso, coords) / mergedata (src/simulation.jl:162-185, src/iso.jl:238): append n_new start points and
 * their K Koopman samples to the resident data without re-uploading what is already there (data must be
 * library-owned).  The resident target becomes invalid, as in the reference where run! recomputes it.
 * Several ranks: every rank passes the same complete block; the contiguous split of the grown N moves all shard
 * boundaries, so the shards of ys are rebuilt on the devices (one all-gather over NVLink, then every rank cuts out
 * its new range); results equal a fresh isokann_set_data_sharded of the grown data bit for bit. */
int32_t isokann_append_data(isokann_ctx *ctx, const float *xs_new, const float *ys_new, int64_t D, int64_t K,
                            int64_t n_new);
/* iso.data = iso.data[end-cutoff+1:end] of run_kde! (src/iso.jl:288-290): keep the newest n_keep start points
 * (several ranks: same n_keep everywhere, shards rebuilt like isokann_append_data) */
int32_t isokann_keep_last(isokann_ctx *ctx, int64_t n_keep);
/* model(flattenlast(propfeatures(data))) of resample_kde / chistratcoords (src/simulation.jl:199-207,227-228):
 * chi of every resident Koopman sample, d x K x N, no K-mean */
int32_t isokann_chis_prop(isokann_ctx *ctx, float *chi_out);
/* Same with buffers already resident on this context's device (no copy of ys: it is adopted by
 * reference and must stay alive until the next set_data / destroy). */
int32_t isokann_set_data_dev(isokann_ctx *ctx, const float *dev_xs, const float *dev_ys_local, int64_t D, int64_t K,
                             int64_t N, int64_t n_offset, int64_t n_local);
/* WeightedSamples (Girsanov) weights for the resident ys, K x n_local (src/data.jl:187-215);
 * NULL clears them. */
int32_t isokann_set_koopman_weights(isokann_ctx *ctx, const float *weights_local);

/* gpu(iso)/cpu(iso), save/load (src/iso.jl:256-257,405-420): parameter and optimiser-state sync */
int32_t isokann_upload_params(isokann_ctx *ctx, const float *flat, int64_t P);
int32_t isokann_download_params(isokann_ctx *ctx, float *flat, int64_t P);
/* Adam: m, v (P each) + beta_t[2]; Nesterov: m = velocity, v and beta_t may be NULL */
int32_t isokann_upload_opt_state(isokann_ctx *ctx, const float *m, const float *v, const float *beta_t, int64_t P);
int32_t isokann_download_opt_state(isokann_ctx *ctx, float *m, float *v, float *beta_t, int64_t P);

/* featurizer(coords) (src/simulation.jl:112,121-124; src/utils/features.jl:22-35) -> F x M */
int32_t isokann_featurize(isokann_ctx *ctx, const float *coords, int64_t D, int64_t M, float *features_out);
/* model(x) (src/iso.jl:203 chis, :211 chicoords; src/isotarget.jl:18,101,154): rows == D
 * (coordinates, is_features == 0) or rows == F (features, is_features == 1) -> d x M */
int32_t isokann_forward(isokann_ctx *ctx, const float *in, int64_t rows, int64_t M, int32_t is_features,
                        float *chi_out);
/* Vector-Jacobian product of model(featurizer(x)) w.r.t. x: grad_out[rows x M] = d(sum(cot .* chi(x)))/dx.
 * dchidx(iso, x) / dchidfeat(iso, feat) (src/utils/minimumpath.jl:3-13) with cot == NULL (ones); the pullback
 * of the featurizer replaces sqpairdist_bwd_kernel! (src/utils/pairdists.jl:153-167,179-196), which the
 * metadynamics bias (src/simulators/metadynamics.jl:40-49) and optimal control differentiate through. */
int32_t isokann_chi_vjp(isokann_ctx *ctx, const float *in, int64_t rows, int64_t M, int32_t is_features,
                        const float *cot, float *grad_out);
/* chis(iso) = model(features(data)) (src/iso.jl:203) on the resident xs -> d x N */
int32_t isokann_chis(isokann_ctx *ctx, float *chi_out);
/* expectation(model, propfeatures(data)) (src/isotarget.jl:18,20) on the resident ys -> d x N */
int32_t isokann_koopman(isokann_ctx *ctx, float *kchi_out);
/* isotarget(target, model, xs, ys) (src/isotarget.jl:12,34,100-107,152-179); the target stays
 * resident for train_epoch; target_out (d x N) may be NULL */
int32_t isokann_target(isokann_ctx *ctx, int32_t transform, const isokann_target_opts *opts, float *target_out);
/* the resident target (d x N) of the last isokann_target / isokann_set_target, for loggers that look at it
 * (log!(logger; iso) in run!, src/iso.jl:85-89) */
int32_t isokann_download_target(isokann_ctx *ctx, float *target_out);
/* validationloss(iso, valdata) (src/iso.jl:160-168): mean((chi(vx) - shiftscale([K chi(vy); K chi(ys)])[1:Nv])^2)
 * in one call; vxs is D x Nv, vys is D x K x Nv (host, column-major).  Only the scalar leaves the device.
 * One dimensional chi only (it shift-scales); ISOKANN_DOMAIN_CONSTANT_CHI as src/isotarget.jl:39. */
int32_t isokann_validationloss(isokann_ctx *ctx, const float *vxs, const float *vys, int64_t D, int64_t K, int64_t Nv,
                               double *loss_out);
/* Diagnostics on the resident data (loggers call them every `logevery` iterations; in the reference each of them pulls
 * chi and Kchi to the host).  chi = chis(iso), Kchi = koopman(iso) are evaluated on the device, their second moments
 * and the residual column norms are fp64 device reductions, the d x d algebra runs on the host in double precision
 * (the reference: Float32 `/` and `log` for rates, Float64 for the residuals).  With several ranks the calls are
 * collective (chi and Kchi are all-gathered: every rank must make the call) and every rank returns the same result.
 * Matrices are column-major as in Julia.
 *
 * rates(iso) (src/iso.jl:339-351) WITHOUT the division by lagtime(sim): Q = log(Kchi / chi), `/` the least-squares
 * right division.  One dimensional chi: rows [chi; 1 - chi] and [Kchi; 1 - Kchi] (:346-349), Q is 2 x 2.
 * q_colmajor: dim x dim doubles, dim = max(d, 2) returned in *dim_out (may be NULL).  ISOKANN_BAD_ARGUMENT if Kchi / chi
 * has an eigenvalue on the closed negative real axis (the reference's `log` turns complex there). */
int32_t isokann_rates(isokann_ctx *ctx, double *q_colmajor, int32_t *dim_out);
/* residual_subspace(iso) (src/isotarget.jl:805-821): V = chis(iso)', KV = koopman(iso)' (N x d, Float64),
 * res = KV - Q Q' KV with Q of the thin QR of V, relres[j] = |res[:, j]| / |KV[:, j]| (v_norms != 0: / |V[:, j]|).
 * relres_out: d doubles.  res_out: N x d doubles or NULL (then only d numbers leave the device). */
int32_t isokann_residual_subspace(isokann_ctx *ctx, int32_t v_norms, double *relres_out, double *res_out);
/* residual_ritz(iso) (src/isotarget.jl:787-802): Kr = Q' KV R^-1 (V = Q R), eigen(Kr, sortby = x -> abs(1 - x)),
 * residues = KQ vecs - vals' .* (Q vecs), relres[j] = |residues[:, j]| / |(KQ vecs)[:, j]|.
 * vals_out: d complex numbers (re, im interleaved); vecs_out: d x d complex, column-major, interleaved, or NULL --
 * eigenvectors of Kr in the basis Q whose R has a positive diagonal, unit 2-norm, largest component real and positive
 * (LAPACK, which the reference calls, fixes neither the signs of R's diagonal nor the sign of a real eigenvector;
 * vals, relres and the residues up to that sign/phase do not depend on either); relres_out: d doubles;
 * residues_out: N x d complex, column-major, interleaved, or NULL. */
int32_t isokann_residual_ritz(isokann_ctx *ctx, double *vals_out, double *vecs_out, double *relres_out,
                              double *residues_out);
/* randperm(rng::Xoshiro, n) of Julia's Random stdlib (the draw Flux.DataLoader(shuffle=true) makes once per epoch,
 * src/iso.jl:181): Xoshiro256++ state in/out (s0..s3 of the Xoshiro struct / task-local RNG), 1-based permutation
 * out.  Host-side (the algorithm is inherently sequential).  UNPINNED against a real Julia session. */
int32_t isokann_randperm(uint64_t *state4, int64_t n, int64_t *perm_out);
/* user-defined isotarget methods (any isotarget(iso, target) dispatched at src/isotarget.jl:10-12, e.g.
 * scripts/251126_carsten/main.jl:132, and the experimental transforms of src/isotarget.jl:190-824): upload a d x N
 * target computed on the host */
int32_t isokann_set_target(isokann_ctx *ctx, const float *target, int64_t d, int64_t N);
/* train_batch!(model, xs, target, opt, minibatch; shuffle, partial) (src/iso.jl:179-194).  perm
 * is the 1-based randperm(N) the DataLoader would draw; returns sum(l)/N in *loss_out */
int32_t isokann_train_epoch(isokann_ctx *ctx, const int64_t *perm, int64_t minibatch, int32_t partial,
                            double *loss_out);
/* run!(iso, n, epochs) (src/iso.jl:72-94) without host round trips: perms holds
 * n_iter*epochs permutations of length N; losses_out receives n_iter*epochs values */
int32_t isokann_iterate(isokann_ctx *ctx, int32_t transform, const isokann_target_opts *opts, int64_t n_iter,
                        int64_t epochs, int64_t minibatch, const int64_t *perms, double *losses_out);

/* Diagnostics of the path's intermediates (loggers / parity tests):
 *  - the flat gradient of the last optimiser step of train_epoch, i.e. what Zygote.withgradient returns at
 *    src/iso.jl:185 (gradient of l/B, summed over ranks), in the flat parameter order;
 *  - the d x d matrices of the last N-D isotarget: TransformPseudoInv's Kinv = chi*pinv(Kchi) (or K when
 *    direct == 0) and T = schur(Kinv).vectors (src/isotarget.jl:165-168), both column-major Float32 as in Julia
 *    (identity for ISA), and the final matrix A with target = A * Kchi after normalisation and fixperm
 *    (row-major double; for TransformISA this is inv(X[i,:])' of src/isotarget.jl:93,104).  Any pointer may be NULL. */
int32_t isokann_download_grads(isokann_ctx *ctx, float *flat, int64_t P);
int32_t isokann_target_matrices(isokann_ctx *ctx, float *kinv_colmajor, float *schur_colmajor,
                                double *applied_rowmajor);

/* on = 1: CUDA events around every kernel launch (per-class times in isokann_stats; training epochs run eagerly);
 * on = 2: one event pair per phase only (ms_*_total), cheap enough to leave on; 0: off */
int32_t isokann_enable_timing(isokann_ctx *ctx, int32_t on);
int32_t isokann_get_stats(isokann_ctx *ctx, isokann_stats *out);
int32_t isokann_reset_stats(isokann_ctx *ctx);
int32_t isokann_synchronize(isokann_ctx *ctx);
/* the CUDA stream all kernels of this context are launched on (for external event timing) */
void *isokann_stream(isokann_ctx *ctx);

/* host-side small dense algebra used by the N-D targets, exposed for CPU tests:
 * real Schur vectors of a d x d float matrix (column-major in/out), LAPACK sgees conventions */
int32_t isokann_host_schur(const float *a_colmajor, int32_t d, float *z_colmajor, float *t_colmajor);
/* ... and by the diagnostics: principal logarithm of a real n x n matrix (n <= 9, column-major doubles), and
 * eigenvalues / right eigenvectors of a real d x d matrix (d <= 8) in LAPACK's dgeev order and normalisation
 * (vals: d x (re, im); vecs: d x d complex column-major interleaved, largest component real and positive) */
int32_t isokann_host_logm(const double *a_colmajor, int32_t n, double *out_colmajor);
/* the d x d part of the three diagnostics as a function of the second moments uu = sum u u', vu = sum v u' of
 * u = [chi, 1], v = [Kchi, 1] (row-major (d+1) x (d+1) doubles), so that CPU tests reach it without a device.
 * what = 0 (rates): out = [n, Q (n x n row-major)];  what = 1 (residual_subspace): out = [A, B] (d x d row-major
 * each; res = A Kchi - B chi per record);  what = 2 (residual_ritz): out = [vals (d x (re, im)), vecs (d x d complex
 * column-major interleaved), Are, Bre, Aim, Bim (d x d row-major each), any_complex]. */
int32_t isokann_host_diag(int32_t what, const double *uu, const double *vu, int32_t d, double *out);
int32_t isokann_host_eig(const double *a_colmajor, int32_t d, double *vals_reim, double *vecs_reim_colmajor);

#ifdef __cplusplus
}
#endif
#endif /* ISOKANN_B200_H */
