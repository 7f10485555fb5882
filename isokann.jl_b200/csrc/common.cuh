// Shared declarations of libisokann_b200: context, error handling, launch accounting.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/isokann_b200.h"

namespace ik {

struct Error {
  int32_t code;
  std::string msg;
};

#define IK_CUDA(expr)                                                                              \
  do {                                                                                             \
    cudaError_t e__ = (expr);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      throw ik::Error{ISOKANN_ERR_CUDA, std::string(#expr) + " -> " + cudaGetErrorString(e__)};    \
  } while (0)

#define IK_REQUIRE(cond, code, text)            \
  do {                                          \
    if (!(cond)) throw ik::Error{(code), (text)}; \
  } while (0)

constexpr int kMaxD = 8;  // largest chi dimension handled by the N-D target kernels
#define ISOKANN_MAX_RANKS 16  // ranks of one NVSwitch node reachable through peer memory (csrc/p2p.cu)

// device-side error flags (sticky until read by the host)
enum : int { FLAG_CONSTANT_CHI = 1, FLAG_NONFINITE_LOSS = 2 };

// kernel classes for the event timers
enum KClass {
  KC_FEATURIZE = 0, KC_GEMM = 1, KC_REDUCE = 2, KC_TRAIN_EW = 3, KC_OPT = 4,
  KC_PHASE_KOOPMAN = 5, KC_PHASE_TARGET = 6, KC_PHASE_TRAIN = 7, KC_NCCL = 8, KC_COUNT = 9
};

struct EventTimer {
  struct Pair {
    cudaEvent_t a, b;
    int cls;
  };
  bool enabled = false;
  bool phases_only = false;  // record the three phase pairs only (isokann_enable_timing(ctx, 2))
  std::vector<Pair> pool;
  size_t used = 0;
  std::vector<size_t> open;
  double ms[KC_COUNT] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  void begin(int cls, cudaStream_t s);  // pairs may nest (phase around kernels)
  void end(cudaStream_t s);
  void flush(cudaStream_t s);  // synchronises the stream and accumulates elapsed times
  void destroy();
};

// bumped by every device (re)allocation: a captured epoch graph holds raw pointers and is only replayed while
// this is unchanged since its capture
inline unsigned long long &alloc_generation() {
  static unsigned long long g = 0;
  return g;
}

template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  void ensure(size_t count) {
    if (count <= n) return;
    ++alloc_generation();
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
    IK_CUDA(cudaMalloc(&p, count * sizeof(T)));
    n = count;
  }
  void release() {
    if (p) {
      cudaFree(p);
      ++alloc_generation();
    }
    p = nullptr;
    n = 0;
  }
};

struct Mat8 {
  double m[kMaxD * kMaxD];  // row-major d x d
};

// replay record of the inner-simplex row transformation (PCCAPlus indexmap)
struct IsaReplay {
  double pre[kMaxD * kMaxD];   // whitening matrix W (row-major), identity if none: x <- x * W
  double x0[kMaxD];            // translation (row selected in round 1, after pre-multiplication)
  double r[kMaxD];             // r[j]  : divisor of round j (j >= 1)
  double v[kMaxD][kMaxD];      // v[j]  : unit row projected out in round j (j >= 1)
  int rounds;                  // number of completed rounds
  int d;
};

struct ArgmaxPartial {
  double val;
  long long idx;
};

// ---- kernel launch wrappers (defined in the .cu files) ----
struct GemmP {
  const float *A;
  int64_t lda;
  const float *B;
  int64_t ldb;
  float *C;
  int64_t ldc;
  const float *Z;  // EPI_MULDACT: activation outputs the derivative is taken at
  int64_t ldz;
  int M, N, K;     // logical extents (including an augmented ones index if used)
  int ones_k;      // A(i, ones_k) := 1 (bias row of the forward pass), -1 = none
  int ones_i;      // A(ones_i, k) := 1 (bias row of the weight gradient), -1 = none
  int kchunk;      // K range per blockIdx.z (split-K); == K rounded up when gridDim.z == 1
  int act;         // activation id for the epilogue
  int epi;         // 0: C = act(acc); 1: C = acc * dact(Z); 2: raw split-K partial at C + z*M*N
};
enum { EPI_ACT = 0, EPI_MULDACT = 1, EPI_PARTIAL = 2 };

struct Ctx;

void launch_featurize(Ctx &c, const float *coords, const int64_t *gather, int64_t gather_off, int64_t M, bool pairs,
                      bool do_ln, float *out, int64_t ldo);
int launch_gemm(Ctx &c, const GemmP &p, bool a_kcontig, bool b_jcontig, int splits);  // returns splits used
// lane = record featurizer (featurize_rec.cu): upper-triangle featurizers (all pairs / atom subset)
bool featurize_rec_applicable(const Ctx &c, bool split, int64_t ld);
void launch_featurize_rec(Ctx &c, const float *coords, const int64_t *gather, int64_t M, bool do_ln, float *out,
                          __nv_bfloat16 *out_hi, __nv_bfloat16 *out_lo, int64_t ld);
void launch_featurize_backward(Ctx &c, const float *in, int64_t M, bool pairs, bool do_ln, const float *gxhat,
                               float *out);
void launch_vjp_seed(Ctx &c, const float *chi, const float *cot, int64_t M, int d, int lastact, float *delta);
void launch_splitk_reduce(Ctx &c, const float *partials, int splits, int64_t count, float *out);
void launch_kmean(Ctx &c, const float *chi, const float *weights, int64_t n, int K, int d, float *out);
void launch_minmax(Ctx &c, const float *x, int64_t n, float *partials, int *nblocks_out);
void launch_shiftscale(Ctx &c, const float *x, int64_t n, const float *partials, int nblocks, float *out, int *flags);
void launch_fill(Ctx &c, float *x, int64_t n, float v);
void launch_f64_to_f32(Ctx &c, const double *in, int64_t n, float *out);
void launch_valloss(Ctx &c, const float *chi, const float *k1, int64_t n, float mn, float mx, double *partials,
                    int *nblocks_out);
void launch_gram(Ctx &c, const float *chi, const float *kchi, int64_t n, int d, double *partials, int *nblocks_out);
void launch_moments(Ctx &c, const float *chi, const float *kchi, int64_t n, int d, double *partials, int *nblocks_out);
void launch_resid(Ctx &c, const float *kchi, const float *chi, int64_t n, int d, const Mat8 &A, const Mat8 &B,
                  double *out_colmajor, double *partials, int *nblocks_out);
void launch_apply(Ctx &c, int mode, const float *kchi, const float *chi, int64_t n, int d, const Mat8 &mat,
                  float *target_out, double *partials, int *nblocks_out);
void launch_isa_argmax(Ctx &c, const float *kchi, int64_t n, const IsaReplay &rp, ArgmaxPartial *partials,
                       int *nblocks_out);
void launch_loss_delta(Ctx &c, const float *chi, const float *target, const int64_t *idx, int64_t idx_off,
                       const float *w, int64_t Bloc, int d, double Bglobal, int lastact, float *delta,
                       double *partials, unsigned int *ticket, float *packed_tail);
void launch_fold_ln(Ctx &c, const float *gamma, const float *beta, const float *W1, const float *b1, int F, int h1,
                    float *folded);
void launch_unfold_ln(Ctx &c, const float *gamma, const float *beta, const float *W1, const float *gfold, int F,
                      int h1, float *g_gamma, float *g_beta, float *g_W1, float *g_b1);
void launch_optimiser(Ctx &c, int64_t P);  // whole vector + beta^t advance
void launch_optimiser_range(Ctx &c, int64_t lo, int64_t hi, bool accumulate_loss);
void launch_advance_beta(Ctx &c);
bool narrow_train_eligible(const isokann_config &g);
bool tiny_forward_eligible(const isokann_config &g);
void launch_tiny_forward(Ctx &c, const float *in, int64_t M, float *out);
void launch_narrow_train(Ctx &c, const float *xhat, int64_t Bloc, const int64_t *idx, double Bglobal,
                         const float *seg0, float *g0);
void launch_perm_to_zero_based(Ctx &c, const int64_t *perm1, int64_t n, int64_t *out0);
void launch_p2p_allreduce(Ctx &c, int64_t lo, int64_t hi);  // in-place sum of grads[lo, hi) over all ranks
void launch_compact_gather(Ctx &c, const float *padded, int world, int64_t nmax, int64_t N, int d, float *out);

double isa_row_norm_host(const float *row, const IsaReplay &rp, double *xout);

// host-side small dense algebra (hostlinalg.cpp)
bool host_inverse(const double *a_rowmajor, int d, double *inv_rowmajor);
void host_sym_eig(const double *a_rowmajor, int d, double *evals, double *evecs_rowmajor);
bool host_schur_f32(const float *a_colmajor, int d, float *z_colmajor, float *t_colmajor);
bool host_schur_f64(const double *a_colmajor, int d, double *z_colmajor, double *t_colmajor);
// diagnostics (hostdiag.cpp): principal matrix logarithm (n <= kMaxD + 1), upper Cholesky factor, general eigenproblem
bool host_logm(const double *a_rowmajor, int n, double *out_rowmajor);
bool host_cholesky_upper(const double *g_rowmajor, int n, double *r_rowmajor);
bool host_eig_general(const double *a_rowmajor, int n, double *wr, double *wi, double *vre_rowmajor,
                      double *vim_rowmajor);
int diag_rates(const double *uu, const double *vu, int d, double *q_rowmajor, int *n_out);
bool diag_subspace(const double *uu, const double *vu, int d, Mat8 &A, Mat8 &B);
bool diag_ritz(const double *uu, const double *vu, int d, double *vals, double *vecs, Mat8 &Are, Mat8 &Bre, Mat8 &Aim,
               Mat8 &Bim, bool *any_complex);

// dynamically loaded NCCL (nccl_dyn.cpp)
struct Nccl;
Nccl *nccl_load(std::string &err);
int nccl_get_unique_id(Nccl *n, void *id128, std::string &err);
void *nccl_comm_init(Nccl *n, int world, int rank, const void *id128, std::string &err);
void nccl_comm_destroy(Nccl *n, void *comm);
void *nccl_comm_split_limited(Nccl *n, void *comm, int rank, int max_ctas);
int nccl_allreduce_sum_f32(Nccl *n, void *comm, float *buf, size_t count, cudaStream_t s, std::string &err);
int nccl_allgather_f32(Nccl *n, void *comm, const float *send, float *recv, size_t count_per_rank, cudaStream_t s,
                       std::string &err);

struct TcState;  // tensor-core path buffers (tc.cuh)

struct Ctx {
  isokann_config cfg{};
  std::vector<int32_t> index;
  int dev = 0;
  cudaStream_t stream = nullptr;
  int num_sms = 148;
  int L = 0, F = 0, D = 0, d = 0, maxw = 0;
  int64_t P = 0;
  bool ln = false;
  int64_t off_gamma = -1, off_beta = -1;
  std::vector<int64_t> off_w, off_b;  // offsets into the flat parameter vector

  // parameters, gradients (+4 tail floats: packed step loss), optimiser state
  DevBuf<float> params, grads, opt_m, opt_v, folded1, gfold;
  DevBuf<float> beta_dev;     // Adam: running (beta1^t, beta2^t), device-resident (replayable steps)
  bool folded_valid = false;  // folded1 matches the current parameters
  bool tc = false;            // wide Dense layers run on tcgen05 (3xBF16 split)
  bool tiny = false;          // every width <= 16, no featurizer/LayerNorm: thread-per-sample forward
  bool fused_train = false;   // narrow net + small minibatch: one fused fwd/loss/bwd kernel per step
  bool tcn = false;           // narrow net: inference forward = one tcgen05 GEMM with the MLP tail in its epilogue
  bool tc_weights_valid = false;
  int split_fmt = 0;          // format the split featurizer kernels write (0 bf16 pairs, 1 fp16 pairs); set per launch
  bool fwd_fp16x2 = false;    // inference forward of the wide layers: fp16 operands, 2 MMAs per product
  bool wf16_valid = false;    // tcs->wF16 matches the current parameters
  int tri_n = 0;            // atoms of an upper-triangle featurizer (all pairs / atom subset), else 0
  DevBuf<int> tri_cmap;     // atom subset: coordinate c of the selection -> coordinate of the record (else null)
  DevBuf<short2> koop_start16;  // fused narrow forward: (i, j) of every 16th feature of the upper triangle
  bool koop_fused_off = false;  // ISOKANN_KOOP_FUSED=0: materialise x_hat and run featurizer + GEMM separately (A/B)
  bool feat_rec_off = false;  // ISOKANN_FEAT_REC=0: keep the lane = feature kernel (A/B comparison)
  bool tc_no_overlap = false; // unless ISOKANN_OVERLAP=1: featurizer and GEMMs on one stream
  bool tc_no_head = false;     // ISOKANN_TC_NO_HEAD=1: separate thin_forward / loss_delta / thin_dgrad kernels (A/B)
  bool head_fused_now = false; // set around the forward pass of a training step whose thin head is fused
  bool tc_no_pair = false;    // ISOKANN_TC_NO_PAIR=1: keep the 1-CTA GEMM (A/B comparison of the 2-CTA kernel)
  TcState *tcs = nullptr;
  DevBuf<int2> pairs;  // coordinate offsets (3a, 3b) per feature
  int n_pairs = 0;
  // atom -> incident features (CSR), for the featurizer pullback: adj[e] = (feature, coordinate offset of the other atom)
  DevBuf<int> adj_off;
  DevBuf<int2> adj;

  // resident data
  DevBuf<float> xs_own, ys_own, kweights;
  const float *xs = nullptr, *ys = nullptr;
  int64_t N = 0, K = 0, n_off = 0, n_loc = 0;
  bool has_weights = false;
  // asynchronous upload of ys (isokann_set_data_async): chunk i of ys is complete when ys_events[i] fires
  cudaStream_t copy_stream = nullptr;
  std::vector<cudaEvent_t> ys_events;
  int64_t ys_chunk_pts = 0;   // start points per upload chunk; 0 = no pending upload
  int64_t ys_chunks_pending = 0;
  cudaEvent_t perm_event = nullptr;            // the coming epoch's permutation is on its way (preload_perm)
  const int64_t *perm_preloaded = nullptr;     // caller pointer whose contents perm_raw holds / will hold
  int64_t perm_preloaded_n = 0;
  cudaEvent_t perm0_event = nullptr;           // perm_raw has been converted (it may be overwritten)
  bool perm0_recorded = false, perm_staged = false;
  void *perm_pinned = nullptr;                 // page-locked staging buffer of the permutation upload
  size_t perm_pinned_bytes = 0;
  cudaEvent_t xs_event = nullptr;  // xs rides the copy stream behind ys (only the training side reads it)
  // cudaFuncSetAttribute is per device: remembered per context, not per process (one process may hold contexts
  // on several devices)
  enum { ATTR_FEAT_BWD = 0, ATTR_FEAT_LN, ATTR_FEAT_BLK0, ATTR_NARROW = 6, ATTR_TC1, ATTR_TC2, ATTR_P2P, ATTR_KOOPF };
  uint32_t func_attr_done = 0;
  bool attr_needed(int bit) {
    if (func_attr_done >> bit & 1u) return false;
    func_attr_done |= 1u << bit;
    return true;
  }
  bool xs_pending = false;
  // caller buffers page-locked by isokann_set_data_async (slot 0: ys, 1: xs); ours = registered here
  struct HostReg {
    void *ptr = nullptr;
    size_t bytes = 0;
    bool ours = false;
  };
  std::vector<HostReg> host_regs;
  DevBuf<float> xs_stage;          // send buffer of the padded xs all-gather (unequal shards)
  bool xs_gather_pending = false;  // multi-rank async upload: only this rank's rows of xs came from the host
  DevBuf<float> chi_x, kchi, kchi_loc, gather_pad, target, w_loss;
  DevBuf<float> val_chi, val_k1;   // validationloss: chi and Koopman expectation on the validation points
  bool has_target = false;
  // d x d matrices of the last N-D target (isokann_target_matrices): Kinv / K and its real Schur vectors
  // (TransformPseudoInv, column-major like the reference's Julia matrices) and the final matrix applied to Kchi
  float last_kinv[kMaxD * kMaxD] = {0}, last_schur[kMaxD * kMaxD] = {0};
  double last_mat[kMaxD * kMaxD] = {0};

  // workspaces
  std::vector<DevBuf<float>> act;  // act[l]: rows x widths[l]
  DevBuf<float> delta_a, delta_b, splitk, staging_in, staging_out, red_f;
  DevBuf<double> red_d, epoch_loss, staging_f64;
  DevBuf<double> diag_part, diag_out;  // diagnostics (rates, residual_*): block partials, N x d residual matrix
  DevBuf<ArgmaxPartial> red_am;
  DevBuf<int64_t> perm_dev, perm_raw;
  DevBuf<int> flags;
  DevBuf<unsigned int> ticket;
  int64_t act_rows = 0;
  void *pinned = nullptr;  // small pinned host scratch for reductions read back
  size_t pinned_bytes = 0;

  // multi-GPU
  Nccl *nccl = nullptr;
  void *comm = nullptr;
  void *comm_ov = nullptr;   // same ranks, at most comm_sms CTAs per collective: the bucket that overlaps the GEMMs
  int world = 1, rank = 0;
  // bucketed gradient exchange beside the backward pass (train_step_overlapped in api.cu)
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_upper = nullptr, ev_lower = nullptr, ev_weights = nullptr;
  bool comm_overlap = false;      // set while train_epoch runs the overlapped step
  bool no_comm_overlap = false;   // ISOKANN_NO_COMM_OVERLAP=1: single-stream step with one all-reduce (A/B)
  bool weights_in_flight = false; // the communication stream still owes the refreshed parameters (ev_weights)
  // ISOKANN_STEP_TRACE=1: CUDA events at the joints of the overlapped step, averaged per epoch and printed to
  // stderr by rank 0 (diagnostic; forces eager epochs)
  bool step_trace = false;
  std::vector<cudaEvent_t> trace_ev;  // [step][12]
  int trace_steps = 0;
  // gradient exchange over CUDA-IPC peer memory (csrc/p2p.cu); falls back to NCCL when the mapping is not possible
  struct P2P {
    bool on = false;
    int ctas = 16;                                  // CTAs of the exchange kernel = SMs the overlapped GEMMs leave free
    float *grads[ISOKANN_MAX_RANKS] = {nullptr};    // every rank's gradient buffer mapped into this process
    uint32_t *flags[ISOKANN_MAX_RANKS] = {nullptr}; // every rank's flag block
    DevBuf<uint32_t> flag_block, seq;
    DevBuf<unsigned int> ticket;
    std::vector<void *> opened;                     // cudaIpcOpenMemHandle results to close
  } p2p;
  int comm_sms = 8;               // SMs the training-step GEMMs leave to NCCL (= NCCL_MAX_CTAS set at comm init)
  int sm_reserve = 0;             // currently reserved (comm_sms during an overlapped epoch)

  // the steps of one training epoch captured as a CUDA graph (train_epoch in api.cu): the offsets into the
  // permutation are static for given (N, minibatch), so an epoch replays with one launch
  struct EpochGraph {
    cudaGraphExec_t exec = nullptr;
    int64_t N = -1, bs = -1, nb = -1;
    const void *xs = nullptr, *target = nullptr, *perm = nullptr;
    unsigned long long alloc_gen = 0;
    bool overlap = false;
    int64_t launches = 0, nccl_calls = 0, gemm_launches = 0, feat_launches = 0;
  } egraph;
  int graph_mode = 1;          // ISOKANN_GRAPH=0 disables capture; epochs with kernel timers on run eagerly anyway
  int64_t eager_epochs = 0;    // epochs run eagerly with the current shapes (the first one allocates: never captured)
  bool weights_swapped = false;  // the operand sets wF/wD are exchanged relative to their canonical order

  // accounting
  isokann_stats stats{};
  EventTimer timer;
  std::string err;

  void count_launch(int cls, double flops_or_bytes = 0.0) {
    stats.kernel_launches++;
    if (cls == KC_GEMM) {
      stats.n_gemm_launches++;
      if (timer.enabled) stats.gemm_flops += flops_or_bytes;
    }
    if (cls == KC_FEATURIZE) {
      stats.n_featurize_launches++;
      if (timer.enabled) stats.featurize_bytes += flops_or_bytes;
    }
  }
};

inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace ik
