// C ABI of libisokann_b200 (include/isokann_b200.h): host orchestration of the ISOKANN
// iteration  featurize -> chi forward over K*N -> K-mean -> isotarget -> minibatched
// fwd/bwd -> optimiser  (reference src/iso.jl:72-94,179-194; src/isotarget.jl:10-42,74-179).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "tc.cuh"

using namespace ik;

struct isokann_ctx : public ik::Ctx {};

// ------------------------------------------------------------------------------------------
// event timers
// ------------------------------------------------------------------------------------------
namespace ik {

void EventTimer::begin(int cls, cudaStream_t s) {
  if (!enabled) return;
  if (phases_only && !(cls >= KC_PHASE_KOOPMAN && cls <= KC_PHASE_TRAIN)) {
    open.push_back((size_t)-1);  // keeps begin/end paired
    return;
  }
  if (used == pool.size()) {
    Pair p;
    cudaEventCreate(&p.a);
    cudaEventCreate(&p.b);
    pool.push_back(p);
  }
  pool[used].cls = cls;
  cudaEventRecord(pool[used].a, s);
  open.push_back(used);
  used++;
}

void EventTimer::end(cudaStream_t s) {
  if (!enabled || open.empty()) return;
  if (open.back() != (size_t)-1) cudaEventRecord(pool[open.back()].b, s);
  open.pop_back();
}

void EventTimer::flush(cudaStream_t s) {
  if (!enabled || used == 0) return;
  cudaStreamSynchronize(s);
  for (size_t i = 0; i < used; ++i) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, pool[i].a, pool[i].b) == cudaSuccess) ms[pool[i].cls] += t;
  }
  used = 0;
  open.clear();
}

void EventTimer::destroy() {
  for (auto &p : pool) {
    cudaEventDestroy(p.a);
    cudaEventDestroy(p.b);
  }
  pool.clear();
  used = 0;
}

}  // namespace ik

// ------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------
namespace {

thread_local std::string g_create_err;

template <typename Fn>
int32_t guarded(isokann_ctx *ctx, Fn &&fn) {
  if (!ctx) return ISOKANN_BAD_ARGUMENT;
  try {
    cudaSetDevice(ctx->dev);
    fn();
    return ISOKANN_OK;
  } catch (const ik::Error &e) {
    ctx->err = e.msg;
    // leave the device and the timer in a clean state for the next call
    if (e.code == ISOKANN_ERR_CUDA) cudaGetLastError();
    ctx->timer.open.clear();
    return e.code;
  } catch (const std::exception &e) {
    ctx->err = e.what();
    ctx->timer.open.clear();
    return ISOKANN_ERR_STATE;
  }
}

// contiguous split of [0, n) over `world` ranks
void split_range(int64_t n, int world, int rank, int64_t *off, int64_t *len) {
  const int64_t base = n / world, rem = n % world;
  *off = rank * base + std::min<int64_t>(rank, rem);
  *len = base + (rank < rem ? 1 : 0);
}

void sync_stream(Ctx &c) { IK_CUDA(cudaStreamSynchronize(c.stream)); }

void ensure_pinned(Ctx &c, size_t bytes) {
  if (bytes <= c.pinned_bytes) return;
  if (c.pinned) cudaFreeHost(c.pinned);
  c.pinned = nullptr;
  c.pinned_bytes = 0;
  IK_CUDA(cudaMallocHost(&c.pinned, bytes));
  c.pinned_bytes = bytes;
}

// device -> pinned host, synchronous
template <typename T>
T *read_back(Ctx &c, const T *dev, size_t count) {
  ensure_pinned(c, count * sizeof(T));
  IK_CUDA(cudaMemcpyAsync(c.pinned, dev, count * sizeof(T), cudaMemcpyDeviceToHost, c.stream));
  sync_stream(c);
  return reinterpret_cast<T *>(c.pinned);
}

int check_flags(Ctx &c) {
  int *f = read_back(c, c.flags.p, 1);
  const int v = *f;
  if (v) IK_CUDA(cudaMemsetAsync(c.flags.p, 0, sizeof(int), c.stream));
  return v;
}

void ensure_act(Ctx &c, int64_t rows) {
  if (rows <= c.act_rows) return;
  for (int l = 0; l <= c.L; ++l) c.act[l].ensure((size_t)rows * c.cfg.widths[l]);
  c.act_rows = rows;
}

const float *layer_segment(Ctx &c, int l) {
  if (l == 0 && c.ln) return c.folded1.p;
  return c.params.p + c.off_w[l];
}

void ensure_folded(Ctx &c) {
  if (!c.ln || c.folded_valid) return;
  launch_fold_ln(c, c.params.p + c.off_gamma, c.params.p + c.off_beta, c.params.p + c.off_w[0],
                 c.params.p + c.off_b[0], c.F, c.cfg.widths[1], c.folded1.p);
  c.folded_valid = true;
}

// featurize (+ parameter-free LayerNorm) M records and run all Dense layers; results stay in
// c.act[0..L] (row-major M x width).  `in` holds coordinate records (in_is_coords) or features.
void forward_rows_tc(Ctx &c, const float *in, const int64_t *gather, int64_t goff, int64_t M, bool in_is_coords,
                     bool keep, const SplitBuf *x_pre = nullptr);
void forward_rows_tcn(Ctx &c, const float *in, const int64_t *gather, int64_t goff, int64_t M, bool in_is_coords);

// keep: the backward pass will need every layer's activations (training step)
void forward_rows(Ctx &c, const float *in, const int64_t *gather, int64_t goff, int64_t M, bool in_is_coords,
                  bool keep = false, bool force_fp32 = false) {
  if (M <= 0) return;
  if (!force_fp32 && c.tcn && !keep) {  // narrow net inference: one GEMM with the MLP tail in its epilogue
    forward_rows_tcn(c, in, gather, goff, M, in_is_coords);
    return;
  }
  if (!force_fp32 && c.tc) {
    forward_rows_tc(c, in, gather, goff, M, in_is_coords, keep);
    return;
  }
  if (!force_fp32 && c.tiny && !keep && gather == nullptr) {
    c.act[c.L].ensure((size_t)M * c.d);
    launch_tiny_forward(c, in, M, c.act[c.L].p);
    return;
  }
  ensure_act(c, M);
  ensure_folded(c);
  const bool pairs = in_is_coords && c.cfg.featurizer != ISOKANN_FEAT_IDENTITY;
  launch_featurize(c, in, gather, goff, M, pairs, c.ln, c.act[0].p, c.F);
  for (int l = 0; l < c.L; ++l) {
    const int fin = c.cfg.widths[l], fout = c.cfg.widths[l + 1];
    GemmP p{};
    p.A = c.act[l].p; p.lda = fin;
    p.B = layer_segment(c, l); p.ldb = fout;
    p.C = c.act[l + 1].p; p.ldc = fout;
    p.M = (int)M; p.N = fout; p.K = fin + 1;
    p.ones_k = fin; p.ones_i = -1;
    p.act = (l < c.L - 1) ? c.cfg.activation : c.cfg.last_activation;
    p.epi = EPI_ACT;
    launch_gemm(c, p, true, true, 1);
  }
}


// ------------------------------------------------------------------------------------------
// tensor-core path (wide nets): split-bf16 activations, tcgen05 GEMMs for layers 0..L-2, thin
// kernels for the last layer
// ------------------------------------------------------------------------------------------
// wide: every hidden layer is a real tensor-core GEMM (AUTO uses tcgen05 for everything);
// any:  the tensor-core kernels can run the net at all (TMA zero-fills the padded tiles) -- used by AUTO for the
//       large-minibatch training steps of narrow nets, where the FP32 CUDA-core GEMMs are latency-bound
bool tc_eligible(const isokann_config &g, bool wide) {
  if (g.n_layers < 2) return false;
  if (g.widths[0] < 64 || g.widths[g.n_layers] > kMaxD) return false;
  for (int l = 1; l < g.n_layers; ++l)
    if (g.widths[l] < (wide ? 256 : 1)) return false;
  return true;
}

// narrow nets (default pairnets): first layer on tensor cores, the rest (widths <= 16) in its epilogue
bool tcn_eligible(const isokann_config &g) {
  if (g.n_layers < 2 || g.n_layers > 4) return false;
  if (g.widths[0] < 64 || g.widths[1] > 128 || g.widths[1] < 1) return false;
  for (int l = 2; l <= g.n_layers; ++l)
    if (g.widths[l] > 16) return false;
  return g.widths[g.n_layers] <= kMaxD;
}

void tc_ensure_rows(Ctx &c, int64_t rows) {
  TcState &t = *c.tcs;
  if (rows > t.rows) {
    for (int l = 0; l < (c.tc ? c.L : 1); ++l) {
      t.act[l].ensure(rows, t.wp[l]);
      // column w_l := 1 once; the GEMM epilogues never touch it when w_l is a multiple of 32 and rewrite it otherwise
      if (l > 0) launch_set_ones_col(c, t.act[l].hi.p, t.act[l].lo.p, rows, t.wp[l], c.cfg.widths[l]);
    }
    c.act[c.L].ensure((size_t)rows * c.d);
    t.rows = rows;
  }
}

void ensure_tc_weights(Ctx &c) {
  if (c.tc_weights_valid) return;
  ensure_folded(c);
  TcState &t = *c.tcs;
  for (int l = 0; l + 1 < (c.tc ? c.L : 2); ++l) {
    const int fin = c.cfg.widths[l], fout = c.cfg.widths[l + 1];
    t.wF[l].ensure(fout, t.wp[l]);
    if (l > 0) t.wD[l].ensure(fin, t.wp[l + 1]);
    launch_prep_weights(c, layer_segment(c, l), fin, fout, t.wF[l].hi.p, t.wF[l].lo.p, t.wp[l],
                        l > 0 ? t.wD[l].hi.p : nullptr, l > 0 ? t.wD[l].lo.p : nullptr, l > 0 ? t.wp[l + 1] : 0);
  }
  c.tc_weights_valid = true;
}

// fp16 copies of the forward operands for the 2-MMA inference forward (regenerated lazily after an update)
void ensure_wf16(Ctx &c) {
  if (c.wf16_valid) return;
  ensure_folded(c);
  TcState &t = *c.tcs;
  t.wF16.resize(c.L);
  for (int l = 0; l + 1 < c.L; ++l) {
    const int fin = c.cfg.widths[l], fout = c.cfg.widths[l + 1];
    t.wF16[l].ensure((size_t)fout * t.wp[l]);
    launch_prep_weights_f16(c, layer_segment(c, l), fin, fout, t.wF16[l].p, t.wp[l]);
  }
  c.wf16_valid = true;
}

// x_pre: x_hat of these rows already featurized (on another stream) into that buffer
// Inference (keep == false) may run in the fp16 / 2-MMA mode (Ctx::fwd_fp16x2): activations as fp16 (hi, lo) pairs,
// weights rounded once to fp16, hi*hi + lo*hi per product.  The training forward always keeps bf16 x 3: its
// activations are operands of the weight-gradient GEMMs together with bf16 deltas.
void forward_rows_tc(Ctx &c, const float *in, const int64_t *gather, int64_t goff, int64_t M, bool in_is_coords,
                     bool keep, const SplitBuf *x_pre) {
  TcState &t = *c.tcs;
  tc_ensure_rows(c, M);
  const bool h2 = c.fwd_fp16x2 && !keep;
  if (h2) ensure_wf16(c);
  else ensure_tc_weights(c);
  const bool pairs = in_is_coords && c.cfg.featurizer != ISOKANN_FEAT_IDENTITY;
  if (!x_pre)
    launch_featurize_split(c, in, gather, goff, M, pairs, c.ln, t.act[0].hi.p, t.act[0].lo.p, t.wp[0], h2 ? 1 : 0);
  const SplitBuf &x0 = x_pre ? *x_pre : t.act[0];
  const int last = c.L - 1;
  const float *seg_last = c.params.p + c.off_w[last];
  for (int l = 0; l + 1 < c.L; ++l) {
    const int fin = c.cfg.widths[l], fout = c.cfg.widths[l + 1];
    TcGemm g{};
    g.a_hi = l == 0 ? x0.hi.p : t.act[l].hi.p; g.a_lo = l == 0 ? x0.lo.p : t.act[l].lo.p; g.lda = t.wp[l];
    if (h2) {
      g.b_hi = g.b_lo = t.wF16[l].p;
      g.fmt = 1;
      g.nmma = 2;
    } else {
      g.b_hi = t.wF[l].hi.p; g.b_lo = t.wF[l].lo.p;
    }
    g.ldb = t.wp[l];
    g.M = (int)M; g.N = fout; g.K = fin;
    g.act = c.cfg.activation;
    g.bias = layer_segment(c, l) + (int64_t)fin * fout;
    g.splits = 1;
    if (l + 2 == c.L && !keep) {
      // inference: the thin last layer is folded into this GEMM's epilogue, z_{L-1} never hits HBM
      const int slots = 2 * cdiv(fout, 256);
      t.dot_partial.ensure((size_t)M * slots * c.d);
      g.epi = TC_EPI_BIAS_ACT_DOT;
      g.w_last = seg_last; g.dot_out = t.dot_partial.p; g.d = c.d;
      launch_tc_gemm(c, g);
      launch_dot_finish(c, t.dot_partial.p, M, slots, c.d, seg_last + (int64_t)fout * c.d, c.cfg.last_activation,
                        c.act[c.L].p);
      return;
    }
    g.epi = TC_EPI_BIAS_ACT_SPLIT;
    g.ones_col = h2 ? 0 : 1;  // the fp16 pass never feeds a weight gradient; the training forward rewrites the column
    g.out_hi = t.act[l + 1].hi.p; g.out_lo = t.act[l + 1].lo.p; g.ldo = t.wp[l + 1];
    launch_tc_gemm(c, g);
  }
  if (c.head_fused_now) return;  // training step: launch_thin_head does the last layer together with its backward
  launch_thin_forward(c, t.act[last].hi.p, t.act[last].lo.p, M, c.cfg.widths[last], t.wp[last], seg_last, c.d,
                      c.cfg.last_activation, c.act[c.L].p);
}

// can forward_rows_tcn run these rows through the fused coordinates -> chi kernel (no x_hat buffers needed)?
bool koop_fused_applicable(const Ctx &c) {
  return c.tcn && !c.koop_fused_off && c.tri_n >= 2 && c.tri_n <= 64 && c.cfg.featurizer != ISOKANN_FEAT_IDENTITY &&
         c.cfg.widths[0] >= 64 && c.cfg.widths[0] == c.tri_n * (c.tri_n - 1) / 2;
}

void forward_rows_tcn(Ctx &c, const float *in, const int64_t *gather, int64_t goff, int64_t M, bool in_is_coords) {
  TcState &t = *c.tcs;
  ensure_tc_weights(c);
  const bool pairs = in_is_coords && c.cfg.featurizer != ISOKANN_FEAT_IDENTITY;
  const int fin = c.cfg.widths[0], fout = c.cfg.widths[1];
  c.act[c.L].ensure((size_t)M * c.d);
  TcGemm g{};
  g.b_hi = t.wF[0].hi.p; g.b_lo = t.wF[0].lo.p; g.ldb = t.wp[0];
  g.M = (int)M; g.N = fout; g.K = fin;
  g.act = c.cfg.activation;
  g.bias = layer_segment(c, 0) + (int64_t)fin * fout;
  g.epi = TC_EPI_TAIL;
  g.tail.nl = c.L - 1;
  g.tail.act = c.cfg.activation;
  g.tail.last_act = c.cfg.last_activation;
  for (int l = 1; l <= c.L; ++l) g.tail.w[l - 1] = c.cfg.widths[l];
  for (int l = 1; l < c.L; ++l) g.tail.seg[l - 1] = c.params.p + c.off_w[l];
  g.chi_out = c.act[c.L].p;
  g.splits = 1;
  // coordinates straight into the tensor-core kernel: x_hat never reaches HBM (and needs no buffer, so the caller
  // may pass millions of rows at once)
  if (pairs && gather == nullptr && koop_fused_applicable(c) && launch_koop_fused(c, in, M, c.ln, g)) return;
  tc_ensure_rows(c, M);
  g.a_hi = t.act[0].hi.p; g.a_lo = t.act[0].lo.p; g.lda = t.wp[0];
  g.chi_out = c.act[c.L].p;
  launch_featurize_split(c, in, gather, goff, M, pairs, c.ln, t.act[0].hi.p, t.act[0].lo.p, t.wp[0]);
  launch_tc_gemm(c, g);
}

// backward + gradient assembly of one minibatch slice on the tensor-core path; forward_rows_tc and
// launch_loss_delta (delta of the last layer in c.delta_a, B x d fp32) have already run
void tc_ensure_delta(Ctx &c, int64_t Bloc) {
  TcState &t = *c.tcs;
  int wpmax = 0;
  for (int l = 0; l < c.L; ++l) wpmax = std::max(wpmax, t.wp[l]);
  t.delta[0].ensure(Bloc, wpmax);
  t.delta[1].ensure(Bloc, wpmax);
  t.dlast.ensure(Bloc, 64);
}

void comm_bucket_upper(Ctx &c);

// The tensor core adds every MMA (K = 16) into the fp32 TMEM accumulator with truncation, so an accumulator that runs
// over the whole minibatch picks up a bias of ~ (#MMAs) * 2^-24 relative to its magnitude: 5e-4 at 65 536 rows
// (measured against the fp32 oracle, tests/test_gpu_fullsize.py).  Weight gradients therefore contract at most 4 096
// rows per accumulator; the slices are summed in fp32 with round-to-nearest by splitk_reduce (fixed order).
int wgrad_accuracy_splits(int64_t Bloc) { return (int)((Bloc + 4095) / 4096); }

// head_done: launch_thin_head already produced delta_L (split, t.dlast) and delta_{L-1} (t.delta[0])
// overlap: multi-rank step -- the gradients of layers >= 2 go to the communication stream as soon as the weight
//          gradient of layer 2 has been launched (comm_bucket_upper)
void backward_tc(Ctx &c, int64_t Bloc, bool head_done, bool overlap = false) {
  TcState &t = *c.tcs;
  const int L = c.L, d = c.d;
  auto sms_now = [&] { return c.num_sms - c.sm_reserve; };  // the reserve changes once the upper bucket is under way
  tc_ensure_delta(c, Bloc);
  int cur = 0;
  {  // last (thin) layer
    const int l = L - 1, fin = c.cfg.widths[l];
    // thin weight gradient [(fin+1) x d] = [z, 1]^T * delta_L on the same MN-major GEMM (N padded by TMA zero fill)
    if (!head_done) launch_f32_to_split(c, c.delta_a.p, Bloc, d, t.dlast.hi.p, t.dlast.lo.p, 64);
    TcGemm w{};
    w.mn_major = 1;
    w.a_hi = t.act[l].hi.p; w.a_lo = t.act[l].lo.p; w.lda = t.wp[l];
    w.b_hi = t.dlast.hi.p; w.b_lo = t.dlast.lo.p; w.ldb = 64;
    w.M = fin + 1; w.N = d; w.K = (int)Bloc;
    w.epi = TC_EPI_F32; w.act = ISOKANN_ACT_IDENTITY;
    w.ldc = d;
    const int tiles = cdiv(w.M, 128);
    const int splits = std::max(wgrad_accuracy_splits(Bloc), std::max(1, std::min(sms_now() / tiles, (int)(Bloc / 1024))));
    if (splits > 1) {
      c.splitk.ensure((size_t)splits * w.M * w.N);
      w.out_f32 = c.splitk.p;
      w.splits = splits;
      const int used = launch_tc_gemm(c, w);
      launch_splitk_reduce(c, c.splitk.p, used, (int64_t)w.M * w.N, c.grads.p + c.off_w[l]);
    } else {
      w.out_f32 = c.grads.p + c.off_w[l];
      w.splits = 1;
      launch_tc_gemm(c, w);
    }
    if (overlap && l == 1) comm_bucket_upper(c);
    if (!head_done)
      launch_thin_dgrad(c, c.delta_a.p, Bloc, d, c.params.p + c.off_w[l], fin, t.act[l].hi.p, t.act[l].lo.p, t.wp[l],
                        c.cfg.activation, t.delta[cur].hi.p, t.delta[cur].lo.p, t.wp[l]);
  }
  for (int l = L - 2; l >= 0; --l) {
    const int fin = c.cfg.widths[l], fout = c.cfg.widths[l + 1];
    // weight + bias gradient: [(fin+1) x fout] = [act_l, 1]^T * delta_{l+1}.  Both operands are consumed as stored
    // (batch-major rows) through MN-major UMMA descriptors; column fin of act_l is the constant 1.
    float *dest = (l == 0 && c.ln) ? c.gfold.p : c.grads.p + c.off_w[l];
    TcGemm w{};
    w.mn_major = 1;
    w.a_hi = t.act[l].hi.p; w.a_lo = t.act[l].lo.p; w.lda = t.wp[l];
    w.b_hi = t.delta[cur].hi.p; w.b_lo = t.delta[cur].lo.p; w.ldb = t.wp[l + 1];
    w.M = fin + 1; w.N = fout; w.K = (int)Bloc;
    w.epi = TC_EPI_F32; w.act = ISOKANN_ACT_IDENTITY;
    w.ldc = fout;
    const int tiles = cdiv(w.M, 128) * cdiv(w.N, 256);
    // as many split-K slices as fit in ONE wave of CTAs (a partial second wave would double the kernel time)
    int splits = std::max(wgrad_accuracy_splits(Bloc), std::max(1, std::min(sms_now() / tiles, (int)(Bloc / 2048))));
    if (splits > 1) {
      c.splitk.ensure((size_t)splits * w.M * w.N);
      w.out_f32 = c.splitk.p;
      w.splits = splits;
      const int used = launch_tc_gemm(c, w);
      launch_splitk_reduce(c, c.splitk.p, used, (int64_t)w.M * w.N, dest);
    } else {
      w.out_f32 = dest;
      w.splits = 1;
      launch_tc_gemm(c, w);
    }
    if (overlap && l == 1 && l < L - 1) comm_bucket_upper(c);
    if (l > 0) {
      // delta_l = (delta_{l+1} * W_l^T) .* act'(z_l)
      TcGemm g{};
      g.a_hi = t.delta[cur].hi.p; g.a_lo = t.delta[cur].lo.p; g.lda = t.wp[l + 1];
      g.b_hi = t.wD[l].hi.p; g.b_lo = t.wD[l].lo.p; g.ldb = t.wp[l + 1];
      g.M = (int)Bloc; g.N = fin; g.K = fout;
      g.epi = TC_EPI_MULDACT_SPLIT; g.act = c.cfg.activation;
      g.z_hi = t.act[l].hi.p; g.z_lo = t.act[l].lo.p; g.ldz = t.wp[l];
      g.out_hi = t.delta[cur ^ 1].hi.p; g.out_lo = t.delta[cur ^ 1].lo.p; g.ldo = t.wp[l];
      g.splits = 1;
      launch_tc_gemm(c, g);
      cur ^= 1;
    }
  }
}

int64_t chunk_rows(const Ctx &c, int64_t multiple) {
  int64_t ch = c.cfg.chunk > 0 ? c.cfg.chunk : 65536;
  // keep the widest activation buffer below ~1 GiB
  const int64_t cap = std::max<int64_t>(1024, (int64_t)(1ull << 28) / std::max(1, c.maxw));
  ch = std::min(ch, cap);
  if (multiple > 1) ch = std::max<int64_t>(multiple, ch / multiple * multiple);
  return ch;
}

// chi for M device-resident records -> dev_out (M x d)
void forward_to(Ctx &c, const float *dev_in, int64_t M, bool in_is_coords, float *dev_out) {
  const int64_t rowlen = in_is_coords ? c.D : c.F;
  const int64_t ch = chunk_rows(c, 1);
  for (int64_t m0 = 0; m0 < M; m0 += ch) {
    const int64_t m = std::min(ch, M - m0);
    forward_rows(c, dev_in + m0 * rowlen, nullptr, 0, m, in_is_coords);
    IK_CUDA(cudaMemcpyAsync(dev_out + m0 * c.d, c.act[c.L].p, (size_t)m * c.d * sizeof(float),
                            cudaMemcpyDeviceToDevice, c.stream));
  }
}

void allgather_rows(Ctx &c, const float *local, int64_t n_local, float *full, int width = 0) {
  // shards follow split_range(N); pad every shard to nmax rows for the fixed-size collective
  const int64_t wd = width > 0 ? width : c.d;
  const int64_t nmax = (c.N + c.world - 1) / c.world;
  c.gather_pad.ensure((size_t)(c.world + 1) * nmax * wd);
  float *send = c.gather_pad.p + (size_t)c.world * nmax * wd;
  IK_CUDA(cudaMemsetAsync(send, 0, (size_t)nmax * wd * sizeof(float), c.stream));
  IK_CUDA(cudaMemcpyAsync(send, local, (size_t)n_local * wd * sizeof(float), cudaMemcpyDeviceToDevice, c.stream));
  std::string err;
  int rc = nccl_allgather_f32(c.nccl, c.comm, send, c.gather_pad.p, (size_t)nmax * wd, c.stream, err);
  IK_REQUIRE(rc == ISOKANN_OK, ISOKANN_ERR_NCCL, err);
  c.stats.nccl_calls++;
  launch_compact_gather(c, c.gather_pad.p, c.world, nmax, c.N, (int)wd, full);
}

// chis(iso): model(features(xs)) on the resident start points -> c.chi_x (N x d)
// xs uploaded by isokann_set_data_async is still in flight on the copy stream: order this stream behind it
void wait_xs(Ctx &c) {
  if (!c.xs_pending) return;
  IK_CUDA(cudaStreamWaitEvent(c.stream, c.xs_event, 0));
  c.xs_pending = false;
}

// After a multi-rank isokann_set_data_async every rank holds only its own rows of xs (they came over PCIe); the
// training side gathers by the global permutation, so the other ranks' rows are fetched over NVLink here: in place
// when the shards are equal, through the padded buffer otherwise.
void gather_xs(Ctx &c) {
  wait_xs(c);
  if (!c.xs_gather_pending) return;
  c.xs_gather_pending = false;
  float *xs = c.xs_own.p;
  if (c.N % c.world == 0) {
    std::string err;
    int rc = nccl_allgather_f32(c.nccl, c.comm, xs + c.n_off * c.D, xs, (size_t)c.n_loc * c.D, c.stream, err);
    IK_REQUIRE(rc == ISOKANN_OK, ISOKANN_ERR_NCCL, err);
    c.stats.nccl_calls++;
  } else {
    c.xs_stage.ensure((size_t)std::max<int64_t>(1, c.n_loc) * c.D);
    IK_CUDA(cudaMemcpyAsync(c.xs_stage.p, xs + c.n_off * c.D, (size_t)c.n_loc * c.D * sizeof(float),
                            cudaMemcpyDeviceToDevice, c.stream));
    allgather_rows(c, c.xs_stage.p, c.n_loc, xs, (int)c.D);
  }
}

void compute_chis(Ctx &c) {
  IK_REQUIRE(c.xs != nullptr, ISOKANN_ERR_STATE, "no data: call isokann_set_data first");
  wait_xs(c);
  c.chi_x.ensure((size_t)c.N * c.d);
  if (c.world == 1) {
    forward_to(c, c.xs, c.N, true, c.chi_x.p);
  } else {
    c.kchi_loc.ensure((size_t)std::max<int64_t>(1, c.n_loc) * c.d);
    forward_to(c, c.xs + c.n_off * c.D, c.n_loc, true, c.kchi_loc.p);
    allgather_rows(c, c.kchi_loc.p, c.n_loc, c.chi_x.p);
  }
}

// expectation(model, ys) -> c.kchi (N x d)
void compute_koopman(Ctx &c) {
  IK_REQUIRE(c.ys != nullptr, ISOKANN_ERR_STATE, "no Koopman samples: call isokann_set_data with ys");
  c.timer.begin(KC_PHASE_KOOPMAN, c.stream);
  c.kchi.ensure((size_t)c.N * c.d);
  float *dst = c.kchi.p;
  if (c.world > 1) {
    c.kchi_loc.ensure((size_t)std::max<int64_t>(1, c.n_loc) * c.d);
    dst = c.kchi_loc.p;
  }
  // the fused narrow-net kernel has no intermediate buffers: it takes up to 4M rows per launch (no wave
  // quantisation between 65 536-row chunks, one K-mean launch), unless ys is still streaming in chunk by chunk
  const bool big = koop_fused_applicable(c) && c.ys_chunk_pts == 0;
  const int64_t ch = big ? std::max<int64_t>(c.K, ((int64_t)1 << 22) / c.K * c.K) : chunk_rows(c, c.K);
  const int64_t nsp = ch / c.K;
  const bool overlap = c.tc && !c.tcn && c.n_loc > nsp && !c.tc_no_overlap;
  if (overlap) {
    // Two streams: the featurizer (CUDA cores / LSU) of chunk i+1 runs next to the tensor-core GEMMs of chunk i.
    // One featurizer block fits beside the resident GEMM CTA on every SM (28 KiB smem and ~27k registers are
    // free), x_hat is double-buffered, events order producer and consumer.
    TcState &t = *c.tcs;
    const bool pairs = c.cfg.featurizer != ISOKANN_FEAT_IDENTITY;
    if (!t.feat_stream) {
      IK_CUDA(cudaStreamCreateWithFlags(&t.feat_stream, cudaStreamNonBlocking));
      for (int b = 0; b < 2; ++b) {
        IK_CUDA(cudaEventCreateWithFlags(&t.feat_done[b], cudaEventDisableTiming));
        IK_CUDA(cudaEventCreateWithFlags(&t.x_free[b], cudaEventDisableTiming));
      }
      IK_CUDA(cudaEventCreateWithFlags(&t.koop_start, cudaEventDisableTiming));
    }
    tc_ensure_rows(c, ch);
    if (c.fwd_fp16x2) ensure_wf16(c);
    else ensure_tc_weights(c);
    t.x_alt.ensure(ch, t.wp[0]);
    SplitBuf *bufs[2] = {&t.act[0], &t.x_alt};
    IK_CUDA(cudaEventRecord(t.koop_start, c.stream));
    IK_CUDA(cudaStreamWaitEvent(t.feat_stream, t.koop_start, 0));  // earlier work on the main stream owns the buffers
    int i = 0;
    for (int64_t n0 = 0; n0 < c.n_loc; n0 += nsp, ++i) {
      const int64_t ns = std::min(nsp, c.n_loc - n0);
      const int b = i & 1;
      if (i >= 2) IK_CUDA(cudaStreamWaitEvent(t.feat_stream, t.x_free[b], 0));
      if (c.ys_chunk_pts > 0) {
        const int64_t last = std::min<int64_t>((n0 + ns - 1) / c.ys_chunk_pts, c.ys_chunks_pending - 1);
        IK_CUDA(cudaStreamWaitEvent(t.feat_stream, c.ys_events[(size_t)last], 0));
      }
      {
        struct SwapBack {  // launch on the featurizer stream; restore the context's stream even if the launch throws
          cudaStream_t &a, &b;
          ~SwapBack() { std::swap(a, b); }
        } back{c.stream, t.feat_stream};
        std::swap(c.stream, t.feat_stream);
        launch_featurize_split(c, c.ys + n0 * c.K * c.D, nullptr, 0, ns * c.K, pairs, c.ln, bufs[b]->hi.p,
                               bufs[b]->lo.p, t.wp[0], c.fwd_fp16x2 ? 1 : 0);
      }
      IK_CUDA(cudaEventRecord(t.feat_done[b], t.feat_stream));
      IK_CUDA(cudaStreamWaitEvent(c.stream, t.feat_done[b], 0));
      forward_rows_tc(c, nullptr, nullptr, 0, ns * c.K, true, false, bufs[b]);
      IK_CUDA(cudaEventRecord(t.x_free[b], c.stream));
      launch_kmean(c, c.act[c.L].p, c.has_weights ? c.kweights.p + n0 * c.K : nullptr, ns, (int)c.K, c.d,
                   dst + n0 * c.d);
    }
  }
  for (int64_t n0 = 0; n0 < c.n_loc && !overlap; n0 += nsp) {
    const int64_t ns = std::min(nsp, c.n_loc - n0);
    if (c.ys_chunk_pts > 0) {  // ys is still streaming in (isokann_set_data_async): wait for the covering chunk
      const int64_t last = std::min<int64_t>((n0 + ns - 1) / c.ys_chunk_pts, c.ys_chunks_pending - 1);
      IK_CUDA(cudaStreamWaitEvent(c.stream, c.ys_events[(size_t)last], 0));
    }
    forward_rows(c, c.ys + n0 * c.K * c.D, nullptr, 0, ns * c.K, true);
    launch_kmean(c, c.act[c.L].p, c.has_weights ? c.kweights.p + n0 * c.K : nullptr, ns, (int)c.K, c.d,
                 dst + n0 * c.d);
  }
  if (c.world > 1) allgather_rows(c, c.kchi_loc.p, c.n_loc, c.kchi.p);
  c.ys_chunk_pts = 0;  // the whole upload has been waited for on this stream
  c.timer.end(c.stream);
}

void set_unit_weights(Ctx &c) {
  c.w_loss.ensure(kMaxD);
  launch_fill(c, c.w_loss.p, kMaxD, 1.0f);
}

// first minimiser over lexicographic permutations of sum_b C[p_b][b]  (fixperm, src/isotarget.jl:120-127)
void best_perm(const double *C, int d, int *perm_out) {
  int p[kMaxD];
  for (int i = 0; i < d; ++i) p[i] = i;
  double best = INFINITY;
  bool have = false;
  do {
    double s = 0.0;
    for (int b = 0; b < d; ++b) s += C[p[b] * d + b];
    if (!have || s < best) {
      best = s;
      have = true;
      for (int i = 0; i < d; ++i) perm_out[i] = p[i];
    }
  } while (std::next_permutation(p, p + d));
}

void sum_partials(const double *part, int nblocks, int per_block, double *out) {
  for (int i = 0; i < per_block; ++i) out[i] = 0.0;
  for (int b = 0; b < nblocks; ++b)
    for (int i = 0; i < per_block; ++i) out[i] += part[(size_t)b * per_block + i];
}

// shared tail of the N-D targets: optional L1 normalisation, optional fixperm, final write + loss weights
void finish_nd_target(Ctx &c, Mat8 mat, bool normalize, bool permute) {
  const int d = c.d;
  int nb = 0;
  c.target.ensure((size_t)c.N * d);
  if (normalize) {  // target ./ norm.(eachrow(target), 1) .* N   (src/isotarget.jl:175)
    launch_apply(c, 0, c.kchi.p, nullptr, c.N, d, mat, nullptr, c.red_d.p, &nb);
    double *part = read_back(c, c.red_d.p, (size_t)nb * d);
    double l1[kMaxD];
    sum_partials(part, nb, d, l1);
    for (int a = 0; a < d; ++a) {
      const double s = (double)c.N / l1[a];
      for (int b = 0; b < d; ++b) mat.m[a * d + b] *= s;
    }
  }
  if (permute) {
    launch_apply(c, 1, c.kchi.p, c.chi_x.p, c.N, d, mat, nullptr, c.red_d.p, &nb);
    double *part = read_back(c, c.red_d.p, (size_t)nb * d * d);
    double C[kMaxD * kMaxD];
    sum_partials(part, nb, d * d, C);
    int p[kMaxD];
    best_perm(C, d, p);
    Mat8 m2;
    for (int b = 0; b < d; ++b)
      for (int k = 0; k < d; ++k) m2.m[b * d + k] = mat.m[p[b] * d + k];
    mat = m2;
  }
  for (int i = 0; i < d * d; ++i) c.last_mat[i] = mat.m[i];
  launch_apply(c, 2, c.kchi.p, nullptr, c.N, d, mat, c.target.p, c.red_d.p, &nb);
  double *part = read_back(c, c.red_d.p, (size_t)nb * d * 2);
  double mom[2 * kMaxD];
  sum_partials(part, nb, 2 * d, mom);
  float w[kMaxD];
  for (int a = 0; a < kMaxD; ++a) w[a] = 1.0f;
  for (int a = 0; a < d; ++a) {  // w = 1 ./ std(target, dims=2), corrected  (src/iso.jl:183)
    const double s = mom[2 * a], q = mom[2 * a + 1];
    const double var = (q - s * s / (double)c.N) / (double)(c.N - 1);
    w[a] = (float)(1.0 / std::sqrt(var));
  }
  c.w_loss.ensure(kMaxD);
  IK_CUDA(cudaMemcpyAsync(c.w_loss.p, w, sizeof(w), cudaMemcpyHostToDevice, c.stream));
  sync_stream(c);
}

void fetch_row(Ctx &c, const float *dev_rows, int64_t idx, int d, float *out) {
  float *r = read_back(c, dev_rows + idx * d, (size_t)d);
  for (int i = 0; i < d; ++i) out[i] = r[i];
}

void target_isa(Ctx &c, const isokann_target_opts &o) {
  const int d = c.d;
  IK_REQUIRE(d > 1, ISOKANN_BAD_ARGUMENT, "TransformISA does not work with one dimensional chi functions");
  IK_REQUIRE(d <= kMaxD, ISOKANN_BAD_ARGUMENT, "chi dimension exceeds ISOKANN_MAX_D");
  if (o.permute) compute_chis(c);
  IsaReplay rp{};
  rp.d = d;
  rp.rounds = 0;
  for (int i = 0; i < d; ++i)
    for (int j = 0; j < d; ++j) rp.pre[i * d + j] = (i == j) ? 1.0 : 0.0;
  int nb = 0;
  if (o.whitening) {  // C = X'X/N, W = C^(-1/2)   (src/isotarget.jl:85-88)
    launch_gram(c, nullptr, c.kchi.p, c.N, d, c.red_d.p, &nb);
    double *part = read_back(c, c.red_d.p, (size_t)nb * d * 2 * d);
    double g[2 * kMaxD * kMaxD];
    sum_partials(part, nb, d * 2 * d, g);
    double C[kMaxD * kMaxD], ev[kMaxD], V[kMaxD * kMaxD];
    for (int a = 0; a < d; ++a)
      for (int b = 0; b < d; ++b) C[a * d + b] = g[a * 2 * d + d + b] / (double)c.N;
    host_sym_eig(C, d, ev, V);
    for (int a = 0; a < d; ++a)
      for (int b = 0; b < d; ++b) {
        double s = 0.0;
        for (int k = 0; k < d; ++k) s += V[a * d + k] * (1.0 / std::sqrt(ev[k])) * V[b * d + k];
        rp.pre[a * d + b] = s;
      }
    for (int i = 0; i < d * d; ++i)
      IK_REQUIRE(std::isfinite(rp.pre[i]), ISOKANN_DOMAIN_SINGULAR_SIMPLEX,
                 "Could not compute the simplex transformation. The subspace might be singular/collapsed");
  }
  long long ind[kMaxD];
  double S[kMaxD * kMaxD];
  for (int j = 0; j < d; ++j) {
    launch_isa_argmax(c, c.kchi.p, c.N, rp, c.red_am.p, &nb);
    ArgmaxPartial *part = read_back(c, c.red_am.p, (size_t)nb);
    double best = -1.0;
    long long bi = -1;
    bool bestnan = false;
    for (int b = 0; b < nb; ++b) {
      const double v = part[b].val;
      const long long i = part[b].idx;
      if (i < 0) continue;
      const bool vnan = v != v;
      bool take;
      if (bi < 0) take = true;
      else if (bestnan) take = vnan && i < bi;
      else if (vnan) take = true;
      else take = v > best || (v == best && i < bi);
      if (take) {
        best = v;
        bi = i;
        bestnan = vnan;
      }
    }
    IK_REQUIRE(bi >= 0 && !bestnan, ISOKANN_DOMAIN_SINGULAR_SIMPLEX,
               "Could not compute the simplex transformation. The subspace might be singular/collapsed");
    ind[j] = bi;
    float row[kMaxD];
    fetch_row(c, c.kchi.p, bi, d, row);
    for (int b = 0; b < d; ++b) S[j * d + b] = (double)row[b];
    double y[kMaxD];
    const double nrm = isa_row_norm_host(row, rp, y);
    if (j == 0) {
      for (int b = 0; b < d; ++b) rp.x0[b] = y[b];
    } else {
      rp.r[j] = nrm;
      for (int b = 0; b < d; ++b) rp.v[j][b] = y[b] / nrm;
    }
    rp.rounds = j + 1;
  }
  double A[kMaxD * kMaxD];
  IK_REQUIRE(host_inverse(S, d, A), ISOKANN_DOMAIN_SINGULAR_SIMPLEX,
             "Could not compute the simplex transformation. The subspace might be singular/collapsed");
  Mat8 mat{};
  for (int a = 0; a < d; ++a)
    for (int b = 0; b < d; ++b) mat.m[a * d + b] = A[b * d + a];  // target = A' * ks
  finish_nd_target(c, mat, false, o.permute != 0);
}

void target_pinv(Ctx &c, const isokann_target_opts &o) {
  const int d = c.d;
  IK_REQUIRE(d > 1, ISOKANN_BAD_ARGUMENT, "TransformPseudoInv does not work with one dimensional chi functions");
  IK_REQUIRE(d <= kMaxD, ISOKANN_BAD_ARGUMENT, "chi dimension exceeds ISOKANN_MAX_D");
  compute_chis(c);
  int nb = 0;
  launch_gram(c, c.chi_x.p, c.kchi.p, c.N, d, c.red_d.p, &nb);
  double *part = read_back(c, c.red_d.p, (size_t)nb * d * 2 * d);
  double g[2 * kMaxD * kMaxD];
  sum_partials(part, nb, d * 2 * d, g);
  double G1[kMaxD * kMaxD], G2[kMaxD * kMaxD];
  const char *msg = "Could not compute the pseudoinverse. The subspace might be singular/collapsed";
  for (int a = 0; a < d; ++a)
    for (int b = 0; b < d; ++b) {
      G1[a * d + b] = g[a * 2 * d + b];       // chi  * kchi'
      G2[a * d + b] = g[a * 2 * d + d + b];   // kchi * kchi'
      IK_REQUIRE(std::isfinite(G1[a * d + b]) && std::isfinite(G2[a * d + b]), ISOKANN_DOMAIN_PINV, msg);
    }
  // pinv(kchi) = kchi' V L^+ V' with kchi kchi' = V L V'; singular values sqrt(L) are cut at
  // rtol*max with rtol = eps(Float32)*min(d,N), as LinearAlgebra.pinv does
  double ev[kMaxD], V[kMaxD * kMaxD], G2p[kMaxD * kMaxD];
  host_sym_eig(G2, d, ev, V);
  double smax = 0.0;
  for (int k = 0; k < d; ++k) smax = std::max(smax, std::sqrt(std::max(ev[k], 0.0)));
  const double tol = 1.1920929e-07 * (double)std::min<int64_t>(d, c.N) * smax;
  for (int a = 0; a < d; ++a)
    for (int b = 0; b < d; ++b) {
      double s = 0.0;
      for (int k = 0; k < d; ++k) {
        const double sv = std::sqrt(std::max(ev[k], 0.0));
        if (sv > tol) s += V[a * d + k] * (1.0 / ev[k]) * V[b * d + k];
      }
      G2p[a * d + b] = s;
    }
  double Kmat[kMaxD * kMaxD];  // direct: Kinv = chi*pinv(kchi); else K = kchi*pinv(kchi)
  const double *left = o.direct ? G1 : G2;
  for (int a = 0; a < d; ++a)
    for (int b = 0; b < d; ++b) {
      double s = 0.0;
      for (int k = 0; k < d; ++k) s += left[a * d + k] * G2p[k * d + b];
      Kmat[a * d + b] = s;
    }
  float Kf_col[kMaxD * kMaxD], Zf_col[kMaxD * kMaxD];
  for (int a = 0; a < d; ++a)
    for (int b = 0; b < d; ++b) Kf_col[a + b * d] = (float)Kmat[a * d + b];
  double T[kMaxD * kMaxD];
  for (int i = 0; i < d * d; ++i) {
    c.last_kinv[i] = Kf_col[i];
    c.last_schur[i] = (i % (d + 1) == 0) ? 1.f : 0.f;
  }
  if (o.eigenvecs) {
    IK_REQUIRE(host_schur_f32(Kf_col, d, Zf_col, nullptr), ISOKANN_DOMAIN_PINV, msg);
    for (int i = 0; i < d * d; ++i) c.last_schur[i] = Zf_col[i];
    for (int a = 0; a < d; ++a)
      for (int b = 0; b < d; ++b) T[a * d + b] = (double)Zf_col[a + b * d];
  } else {
    for (int a = 0; a < d; ++a)
      for (int b = 0; b < d; ++b) T[a * d + b] = (a == b) ? 1.0 : 0.0;
  }
  double R[kMaxD * kMaxD];  // direct: Kinv; else inv(K)
  if (o.direct) {
    for (int i = 0; i < d * d; ++i) R[i] = (double)(float)Kmat[i];
  } else {
    double Kr[kMaxD * kMaxD];
    for (int i = 0; i < d * d; ++i) Kr[i] = (double)(float)Kmat[i];
    IK_REQUIRE(host_inverse(Kr, d, R), ISOKANN_DOMAIN_PINV, msg);
  }
  Mat8 mat{};
  for (int a = 0; a < d; ++a)
    for (int b = 0; b < d; ++b) {
      double s = 0.0;
      for (int k = 0; k < d; ++k) s += T[a * d + k] * R[k * d + b];
      mat.m[a * d + b] = s;
    }
  finish_nd_target(c, mat, o.normalize != 0, o.permute != 0);
}

void compute_target(Ctx &c, int transform, const isokann_target_opts *opts) {
  isokann_target_opts o{1, 0, 1, 1, 1};
  if (opts) o = *opts;
  compute_koopman(c);
  c.timer.begin(KC_PHASE_TARGET, c.stream);
  if (transform == ISOKANN_TARGET_SHIFTSCALE) {
    IK_REQUIRE(c.d == 1, ISOKANN_BAD_ARGUMENT, "TransformShiftscale only works with one dimensional chi functions");
    c.target.ensure((size_t)c.N);
    int nb = 0;
    launch_minmax(c, c.kchi.p, c.N, c.red_f.p, &nb);
    launch_shiftscale(c, c.kchi.p, c.N, c.red_f.p, nb, c.target.p, c.flags.p);
    set_unit_weights(c);
    const int f = check_flags(c);
    if (f & FLAG_CONSTANT_CHI)
      throw ik::Error{ISOKANN_DOMAIN_CONSTANT_CHI, "Could not compute the shift-scale. chi function is constant"};
  } else if (transform == ISOKANN_TARGET_ISA) {
    target_isa(c, o);
  } else if (transform == ISOKANN_TARGET_PINV) {
    target_pinv(c, o);
  } else {
    throw ik::Error{ISOKANN_BAD_ARGUMENT, "unknown target transform"};
  }
  c.has_target = true;
  c.timer.end(c.stream);
}

int pick_splits(const Ctx &c, int Mout, int Nout, int Kred) {
  const int tiles = cdiv(Mout, Mout > 64 ? 128 : 64) * cdiv(Nout, Nout > 64 ? 128 : 64);
  int s = (2 * c.num_sms + tiles - 1) / tiles;
  const int maxs = std::max(1, Kred / 256);
  s = std::max(1, std::min(s, maxs));
  return std::min(s, 64);
}

// Map every rank's gradient buffer and flag block into this process (CUDA IPC; the 64-byte handles travel through
// one NCCL all-gather).  All ranks agree on the outcome with an all-reduce, so either every rank uses the peer-memory
// exchange or every rank stays on NCCL.
void setup_p2p(Ctx &c) {
  Ctx::P2P &q = c.p2p;
  q.on = false;
  if (c.world < 2 || c.world > ISOKANN_MAX_RANKS) return;
  if (const char *e = getenv("ISOKANN_P2P_CTAS")) q.ctas = std::max(1, std::min(64, atoi(e)));
  q.flag_block.ensure(2 * ISOKANN_MAX_RANKS);
  q.seq.ensure(1);
  q.ticket.ensure(1);
  IK_CUDA(cudaMemset(q.flag_block.p, 0, 2 * ISOKANN_MAX_RANKS * sizeof(uint32_t)));
  IK_CUDA(cudaMemset(q.seq.p, 0, sizeof(uint32_t)));
  IK_CUDA(cudaMemset(q.ticket.p, 0, sizeof(unsigned int)));
  constexpr int HF = (int)(sizeof(cudaIpcMemHandle_t) / sizeof(float));  // 16 floats per handle
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is expected to be 64 bytes");
  float send_h[2 * HF + 2];
  float ok = 1.f;
  cudaIpcMemHandle_t hg, hf;
  if (cudaIpcGetMemHandle(&hg, c.grads.p) != cudaSuccess || cudaIpcGetMemHandle(&hf, q.flag_block.p) != cudaSuccess) {
    cudaGetLastError();
    ok = 0.f;
    memset(&hg, 0, sizeof hg);
    memset(&hf, 0, sizeof hf);
  }
  memcpy(send_h, &hg, 64);
  memcpy(send_h + HF, &hf, 64);
  send_h[2 * HF] = ok;
  send_h[2 * HF + 1] = 0.f;
  const int per = 2 * HF + 2;
  DevBuf<float> xch;
  xch.ensure((size_t)per * (c.world + 1));
  IK_CUDA(cudaMemcpyAsync(xch.p, send_h, per * sizeof(float), cudaMemcpyHostToDevice, c.stream));
  std::string err;
  int rc = nccl_allgather_f32(c.nccl, c.comm, xch.p, xch.p + per, (size_t)per, c.stream, err);
  IK_REQUIRE(rc == ISOKANN_OK, ISOKANN_ERR_NCCL, err);
  std::vector<float> all((size_t)per * c.world);
  IK_CUDA(cudaMemcpyAsync(all.data(), xch.p + per, all.size() * sizeof(float), cudaMemcpyDeviceToHost, c.stream));
  sync_stream(c);
  for (int p = 0; p < c.world && ok != 0.f; ++p) {
    if (all[(size_t)p * per + 2 * HF] == 0.f) ok = 0.f;
  }
  for (int p = 0; p < c.world && ok != 0.f; ++p) {
    if (p == c.rank) {
      q.grads[p] = c.grads.p;
      q.flags[p] = q.flag_block.p;
      continue;
    }
    cudaIpcMemHandle_t a, b;
    memcpy(&a, &all[(size_t)p * per], 64);
    memcpy(&b, &all[(size_t)p * per + HF], 64);
    void *pg = nullptr, *pf = nullptr;
    if (cudaIpcOpenMemHandle(&pg, a, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      ok = 0.f;
      break;
    }
    q.opened.push_back(pg);
    if (cudaIpcOpenMemHandle(&pf, b, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      ok = 0.f;
      break;
    }
    q.opened.push_back(pf);
    q.grads[p] = (float *)pg;
    q.flags[p] = (uint32_t *)pf;
  }
  // agreement: the exchange is used only if every rank mapped every peer
  IK_CUDA(cudaMemcpyAsync(xch.p, &ok, sizeof(float), cudaMemcpyHostToDevice, c.stream));
  rc = nccl_allgather_f32(c.nccl, c.comm, xch.p, xch.p + per, 1, c.stream, err);
  IK_REQUIRE(rc == ISOKANN_OK, ISOKANN_ERR_NCCL, err);
  std::vector<float> oks((size_t)c.world);
  IK_CUDA(cudaMemcpyAsync(oks.data(), xch.p + per, oks.size() * sizeof(float), cudaMemcpyDeviceToHost, c.stream));
  sync_stream(c);
  bool all_ok = true;
  for (float v : oks) all_ok = all_ok && v != 0.f;
  xch.release();
  if (!all_ok) {
    for (void *h : q.opened) cudaIpcCloseMemHandle(h);
    q.opened.clear();
    cudaGetLastError();
    return;
  }
  q.on = true;
  c.comm_sms = q.ctas;  // the overlapped GEMMs leave exactly the exchange kernel's CTAs free
}

// launch on another stream through the same wrappers; the context's stream is restored even if a launch throws
struct StreamScope {
  Ctx &c;
  cudaStream_t saved;
  StreamScope(Ctx &ctx, cudaStream_t s) : c(ctx), saved(ctx.stream) { c.stream = s; }
  ~StreamScope() { c.stream = saved; }
};

// in-place all-reduce(SUM) of grads[lo, hi) on the context's current stream
void allreduce_grads(Ctx &c, int64_t lo, int64_t hi, bool limited = false) {
  if (c.p2p.on) {  // the library's own two-shot exchange over NVLink peer memory (csrc/p2p.cu)
    launch_p2p_allreduce(c, lo, hi);
    return;
  }
  std::string err;
  c.timer.begin(KC_NCCL, c.stream);
  void *comm = limited && c.comm_ov ? c.comm_ov : c.comm;
  int rc = nccl_allreduce_sum_f32(c.nccl, comm, c.grads.p + lo, (size_t)(hi - lo), c.stream, err);
  c.timer.end(c.stream);
  IK_REQUIRE(rc == ISOKANN_OK, ISOKANN_ERR_NCCL, err);
  c.stats.nccl_calls++;
}

void train_step_overlapped(Ctx &c, int64_t s0, int64_t Bloc, int64_t len);

// one optimiser step on the minibatch perm[start, start+len): forward, loss, backward, update
void train_step(Ctx &c, int64_t start, int64_t len) {
  int64_t loff, Bloc;
  split_range(len, c.world, c.rank, &loff, &Bloc);
  const int64_t s0 = start + loff;
  const int L = c.L, d = c.d;
  if (c.comm_overlap) {  // multi-rank tensor-core path: bucketed gradient exchange next to the backward pass
    train_step_overlapped(c, s0, Bloc, len);
    return;
  }
  if (Bloc == 0) {
    IK_CUDA(cudaMemsetAsync(c.grads.p, 0, (size_t)(c.P + 4) * sizeof(float), c.stream));
  } else if (c.fused_train && Bloc <= 8192) {
    // narrow net, small minibatch: featurize -> one fused forward/loss/backward kernel -> ordered reduce
    ensure_act(c, Bloc);
    ensure_folded(c);
    const bool pairs = c.cfg.featurizer != ISOKANN_FEAT_IDENTITY;
    launch_featurize(c, c.xs, c.perm_dev.p, s0, Bloc, pairs, c.ln, c.act[0].p, c.F);
    launch_narrow_train(c, c.act[0].p, Bloc, c.perm_dev.p + s0, (double)len, layer_segment(c, 0),
                        c.ln ? c.gfold.p : c.grads.p + c.off_w[0]);
    if (c.ln)
      launch_unfold_ln(c, c.params.p + c.off_gamma, c.params.p + c.off_beta, c.params.p + c.off_w[0], c.gfold.p, c.F,
                       c.cfg.widths[1], c.grads.p + c.off_gamma, c.grads.p + c.off_beta, c.grads.p + c.off_w[0],
                       c.grads.p + c.off_b[0]);
  } else {
    const bool head = c.tc && !c.tc_no_head && thin_head_eligible(c);
    {
      struct Reset {  // the flag must not survive an error thrown inside the forward pass
        bool &f;
        ~Reset() { f = false; }
      } reset{c.head_fused_now};
      c.head_fused_now = head;
      forward_rows(c, c.xs, c.perm_dev.p, s0, Bloc, true, true);
    }
    c.delta_a.ensure((size_t)Bloc * c.maxw);
    c.delta_b.ensure((size_t)Bloc * c.maxw);
    float *cur = c.delta_a.p, *other = c.delta_b.p;
    if (head) {
      TcState &t = *c.tcs;
      const int last = L - 1;
      tc_ensure_delta(c, Bloc);
      c.red_d.ensure(1024);
      launch_thin_head(c, t.act[last].hi.p, t.act[last].lo.p, Bloc, c.cfg.widths[last], t.wp[last],
                       c.params.p + c.off_w[last], c.target.p, c.perm_dev.p + s0, c.w_loss.p, (double)len, c.act[L].p,
                       cur, t.dlast.hi.p, t.dlast.lo.p, 64, t.delta[0].hi.p, t.delta[0].lo.p, t.wp[last], c.red_d.p,
                       c.ticket.p, c.grads.p + c.P);
    } else {
      launch_loss_delta(c, c.act[L].p, c.target.p, c.perm_dev.p, s0, c.w_loss.p, Bloc, d, (double)len,
                        c.cfg.last_activation, cur, c.red_d.p, c.ticket.p, c.grads.p + c.P);
    }
    if (c.tc) backward_tc(c, Bloc, head);
    for (int l = L - 1; l >= 0 && !c.tc; --l) {
      const int fin = c.cfg.widths[l], fout = c.cfg.widths[l + 1];
      // weight + bias gradient: [(fin+1) x fout] = [act[l], 1]^T * delta
      float *dest = (l == 0 && c.ln) ? c.gfold.p : c.grads.p + c.off_w[l];
      GemmP w{};
      w.A = c.act[l].p; w.lda = fin;
      w.B = cur; w.ldb = fout;
      w.M = fin + 1; w.N = fout; w.K = (int)Bloc;
      w.ones_k = -1; w.ones_i = fin;
      w.act = ISOKANN_ACT_IDENTITY;
      const int splits = pick_splits(c, w.M, w.N, w.K);
      if (splits > 1) {
        c.splitk.ensure((size_t)splits * w.M * w.N);
        w.C = c.splitk.p; w.ldc = fout; w.epi = EPI_PARTIAL;
        const int used = launch_gemm(c, w, false, true, splits);
        launch_splitk_reduce(c, c.splitk.p, used, (int64_t)w.M * w.N, dest);
      } else {
        w.C = dest; w.ldc = fout; w.epi = EPI_ACT;
        launch_gemm(c, w, false, true, 1);
      }
      if (l > 0) {
        // delta_{l-1} = (delta_l * W_l^T) .* act'(z_{l-1})
        GemmP g{};
        g.A = cur; g.lda = fout;
        g.B = c.params.p + c.off_w[l]; g.ldb = fout;
        g.C = other; g.ldc = fin;
        g.Z = c.act[l].p; g.ldz = fin;
        g.M = (int)Bloc; g.N = fin; g.K = fout;
        g.ones_k = -1; g.ones_i = -1;
        g.act = c.cfg.activation;
        g.epi = EPI_MULDACT;
        launch_gemm(c, g, true, false, 1);
        std::swap(cur, other);
      }
    }
    if (c.ln)
      launch_unfold_ln(c, c.params.p + c.off_gamma, c.params.p + c.off_beta, c.params.p + c.off_w[0], c.gfold.p, c.F,
                       c.cfg.widths[1], c.grads.p + c.off_gamma, c.grads.p + c.off_beta, c.grads.p + c.off_w[0],
                       c.grads.p + c.off_b[0]);
  }
  if (c.world > 1) allreduce_grads(c, 0, c.P + 2);
  launch_optimiser(c, c.P);
  c.folded_valid = false;
  c.tc_weights_valid = false;
  c.wf16_valid = false;
}

// ------------------------------------------------------------------------------------------
// multi-rank training step on the tensor-core path.  The flat gradient is exchanged in two buckets on a second
// stream while the backward pass is still running:
//   upper bucket  [off_w[1], P+2): layers >= 2 and the packed step loss -- complete once the weight gradient of
//                 layer 2 has run; all-reduce -> optimiser -> split-bf16 operands of those layers, all next to the
//                 data-gradient GEMM of layer 2 and the weight-gradient GEMM of layer 1 on the main stream
//   lower bucket  [0, off_w[1]): LayerNorm affine + layer 1 -- after unfold_ln; all-reduce -> optimiser -> fold ->
//                 operands of layer 1, next to the gather + featurizer of the NEXT minibatch
// The refreshed operands go to a second set of buffers (the data-gradient GEMM still reads the current set) and the
// two sets are swapped at the end of the step.  The GEMMs of the step leave `sm_reserve` SMs to the NCCL kernels: a
// persistent GEMM CTA fills an SM's shared memory, so without the reserve the collective would only start when a
// GEMM ends.  Same arithmetic as the single-stream step (same global minibatch, gradient of l/B, src/iso.jl:184-192).
// ------------------------------------------------------------------------------------------
void join_comm_stream(Ctx &c);

// trace slots: main stream 0 step begin, 1 after featurizer, 2 weights ready, 3 after forward + head, 4 upper bucket
// handed over, 5 backward done (lower bucket handed over); communication stream 6 upper all-reduce starts, 7 ends,
// 8 upper optimiser + operands done, 9 lower all-reduce starts, 10 ends, 11 weights published
constexpr int kTraceSlots = 12;
void trace_mark(Ctx &c, int slot) {
  if (!c.step_trace) return;
  const size_t need = (size_t)(c.trace_steps + 1) * kTraceSlots;
  while (c.trace_ev.size() < need) {
    cudaEvent_t e;
    IK_CUDA(cudaEventCreate(&e));
    c.trace_ev.push_back(e);
  }
  IK_CUDA(cudaEventRecord(c.trace_ev[(size_t)c.trace_steps * kTraceSlots + slot], c.stream));
}

void trace_report(Ctx &c) {
  if (!c.step_trace || c.trace_steps == 0) return;
  IK_CUDA(cudaStreamSynchronize(c.stream));
  if (c.comm_stream) IK_CUDA(cudaStreamSynchronize(c.comm_stream));
  double acc[kTraceSlots] = {0};
  double span = 0.0;
  for (int st = 0; st < c.trace_steps; ++st) {
    cudaEvent_t *e = &c.trace_ev[(size_t)st * kTraceSlots];
    for (int k = 1; k < kTraceSlots; ++k) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, e[0], e[k]) == cudaSuccess) acc[k] += ms;
    }
    if (st + 1 < c.trace_steps) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, e[0], e[kTraceSlots]) == cudaSuccess) span += ms;
    }
  }
  cudaGetLastError();
  if (c.rank == 0) {
    fprintf(stderr, "[isokann step trace] %d steps, us from step begin: featurized %.0f | weights ready %.0f | forward+head %.0f | "
                    "upper handed over %.0f | backward done %.0f || comm: upper AR %.0f..%.0f, upper opt+prep done %.0f, "
                    "lower AR %.0f..%.0f, weights published %.0f | step period %.0f\n",
            c.trace_steps, 1e3 * acc[1] / c.trace_steps, 1e3 * acc[2] / c.trace_steps, 1e3 * acc[3] / c.trace_steps,
            1e3 * acc[4] / c.trace_steps, 1e3 * acc[5] / c.trace_steps, 1e3 * acc[6] / c.trace_steps,
            1e3 * acc[7] / c.trace_steps, 1e3 * acc[8] / c.trace_steps, 1e3 * acc[9] / c.trace_steps,
            1e3 * acc[10] / c.trace_steps, 1e3 * acc[11] / c.trace_steps,
            c.trace_steps > 1 ? 1e3 * span / (c.trace_steps - 1) : 0.0);
  }
  c.trace_steps = 0;
}

void ensure_comm_stream(Ctx &c) {
  if (c.comm_stream) return;
  IK_CUDA(cudaStreamCreateWithFlags(&c.comm_stream, cudaStreamNonBlocking));
  IK_CUDA(cudaEventCreateWithFlags(&c.ev_upper, cudaEventDisableTiming));
  IK_CUDA(cudaEventCreateWithFlags(&c.ev_lower, cudaEventDisableTiming));
  IK_CUDA(cudaEventCreateWithFlags(&c.ev_weights, cudaEventDisableTiming));
}

void prep_layer_alt(Ctx &c, int l) {
  TcState &t = *c.tcs;
  const int fin = c.cfg.widths[l], fout = c.cfg.widths[l + 1];
  launch_prep_weights(c, layer_segment(c, l), fin, fout, t.wF_alt[l].hi.p, t.wF_alt[l].lo.p, t.wp[l],
                      l > 0 ? t.wD_alt[l].hi.p : nullptr, l > 0 ? t.wD_alt[l].lo.p : nullptr, l > 0 ? t.wp[l + 1] : 0);
}

void comm_bucket_upper(Ctx &c) {
  trace_mark(c, 4);
  IK_CUDA(cudaEventRecord(c.ev_upper, c.stream));
  StreamScope on(c, c.comm_stream);
  IK_CUDA(cudaStreamWaitEvent(c.stream, c.ev_upper, 0));
  const int64_t lo = c.off_w[1];
  trace_mark(c, 6);
  allreduce_grads(c, lo, c.P + 2, true);
  trace_mark(c, 7);
  launch_optimiser_range(c, lo, c.P, true);
  for (int l = 1; l + 1 < c.L; ++l) prep_layer_alt(c, l);
  trace_mark(c, 8);
  c.sm_reserve = c.comm_sms;  // the GEMMs launched from here to the end of the backward pass share the GPU with NCCL
}

void comm_bucket_lower(Ctx &c) {
  c.sm_reserve = 0;  // nothing on the main stream overlaps the lower bucket but the next featurizer
  IK_CUDA(cudaEventRecord(c.ev_lower, c.stream));
  StreamScope on(c, c.comm_stream);
  IK_CUDA(cudaStreamWaitEvent(c.stream, c.ev_lower, 0));
  trace_mark(c, 9);
  allreduce_grads(c, 0, c.off_w[1]);
  trace_mark(c, 10);
  launch_optimiser_range(c, 0, c.off_w[1], false);
  launch_advance_beta(c);
  if (c.ln)
    launch_fold_ln(c, c.params.p + c.off_gamma, c.params.p + c.off_beta, c.params.p + c.off_w[0],
                   c.params.p + c.off_b[0], c.F, c.cfg.widths[1], c.folded1.p);
  prep_layer_alt(c, 0);
  trace_mark(c, 11);
  IK_CUDA(cudaEventRecord(c.ev_weights, c.stream));
}

void train_step_overlapped(Ctx &c, int64_t s0, int64_t Bloc, int64_t len) {
  TcState &t = *c.tcs;
  const int L = c.L;
  ensure_comm_stream(c);
  tc_ensure_rows(c, Bloc);
  if (!c.weights_in_flight) ensure_tc_weights(c);
  for (int l = 0; l + 1 < L; ++l) {  // second operand set
    t.wF_alt[l].ensure(c.cfg.widths[l + 1], t.wp[l]);
    if (l > 0) t.wD_alt[l].ensure(c.cfg.widths[l], t.wp[l + 1]);
  }
  if (Bloc == 0) {  // a ragged last minibatch can leave this rank without rows: it contributes zeros
    join_comm_stream(c);
    IK_CUDA(cudaMemsetAsync(c.grads.p, 0, (size_t)(c.P + 4) * sizeof(float), c.stream));
    comm_bucket_upper(c);
    comm_bucket_lower(c);
    std::swap(t.wF, t.wF_alt);
    std::swap(t.wD, t.wD_alt);
    c.weights_swapped = !c.weights_swapped;
    c.weights_in_flight = c.folded_valid = c.tc_weights_valid = true;
    c.wf16_valid = false;
    return;
  }
  const bool pairs = c.cfg.featurizer != ISOKANN_FEAT_IDENTITY;
  trace_mark(c, 0);
  // gather + featurizer of this minibatch: independent of the weights, so it runs beside the tail of the previous step
  launch_featurize_split(c, c.xs, c.perm_dev.p, s0, Bloc, pairs, c.ln, t.act[0].hi.p, t.act[0].lo.p, t.wp[0]);
  trace_mark(c, 1);
  if (c.weights_in_flight) {
    IK_CUDA(cudaStreamWaitEvent(c.stream, c.ev_weights, 0));
    c.weights_in_flight = false;
  }
  trace_mark(c, 2);
  const bool head = !c.tc_no_head && thin_head_eligible(c);
  {
    struct Reset {
      bool &f;
      ~Reset() { f = false; }
    } reset{c.head_fused_now};
    c.head_fused_now = head;
    forward_rows_tc(c, nullptr, nullptr, 0, Bloc, true, true, &t.act[0]);
  }
  c.delta_a.ensure((size_t)Bloc * c.maxw);
  const int last = L - 1;
  tc_ensure_delta(c, Bloc);
  if (head) {
    c.red_d.ensure(1024);
    launch_thin_head(c, t.act[last].hi.p, t.act[last].lo.p, Bloc, c.cfg.widths[last], t.wp[last],
                     c.params.p + c.off_w[last], c.target.p, c.perm_dev.p + s0, c.w_loss.p, (double)len, c.act[L].p,
                     c.delta_a.p, t.dlast.hi.p, t.dlast.lo.p, 64, t.delta[0].hi.p, t.delta[0].lo.p, t.wp[last],
                     c.red_d.p, c.ticket.p, c.grads.p + c.P);
  } else {
    launch_loss_delta(c, c.act[L].p, c.target.p, c.perm_dev.p, s0, c.w_loss.p, Bloc, c.d, (double)len,
                      c.cfg.last_activation, c.delta_a.p, c.red_d.p, c.ticket.p, c.grads.p + c.P);
  }
  trace_mark(c, 3);
  backward_tc(c, Bloc, head, true);
  if (c.ln)
    launch_unfold_ln(c, c.params.p + c.off_gamma, c.params.p + c.off_beta, c.params.p + c.off_w[0], c.gfold.p, c.F,
                     c.cfg.widths[1], c.grads.p + c.off_gamma, c.grads.p + c.off_beta, c.grads.p + c.off_w[0],
                     c.grads.p + c.off_b[0]);
  trace_mark(c, 5);
  comm_bucket_lower(c);
  if (c.step_trace) c.trace_steps++;
  std::swap(t.wF, t.wF_alt);
  std::swap(t.wD, t.wD_alt);
  c.weights_swapped = !c.weights_swapped;
  c.weights_in_flight = true;   // the next reader of the parameters / operands waits for ev_weights
  c.folded_valid = true;
  c.tc_weights_valid = true;
  c.wf16_valid = false;
}

// order the main stream behind the communication stream (end of an epoch, or before anything else reads the model)
void join_comm_stream(Ctx &c) {
  if (!c.weights_in_flight) return;
  IK_CUDA(cudaStreamWaitEvent(c.stream, c.ev_weights, 0));
  c.weights_in_flight = false;
}

// Capture the optimiser steps of one epoch into a CUDA graph.  Nothing is executed here; the caller launches the
// instantiated graph.  Any failure leaves the context without a graph (the epoch then runs eagerly) and switches
// capture off for this context.
template <typename Steps>
void capture_epoch(Ctx &c, Steps &&steps, int64_t N, int64_t bs, int64_t nb, bool overlap) {
  Ctx::EpochGraph &g = c.egraph;
  if (g.exec) {
    cudaGraphExecDestroy(g.exec);
    g.exec = nullptr;
  }
  const isokann_stats before = c.stats;
  const bool fv = c.folded_valid, tv = c.tc_weights_valid, sw = c.weights_swapped;
  const unsigned long long gen = alloc_generation();
  cudaGraph_t graph = nullptr;
  bool ok = cudaStreamBeginCapture(c.stream, cudaStreamCaptureModeRelaxed) == cudaSuccess;
  if (ok) {
    try {
      steps();
    } catch (const ik::Error &) {
      ok = false;
    }
    if (cudaStreamEndCapture(c.stream, &graph) != cudaSuccess || !graph) ok = false;
  }
  const bool realloc = gen != alloc_generation();  // a buffer grew during the capture: the graph holds stale pointers
  if (ok && !realloc) ok = cudaGraphInstantiate(&g.exec, graph, 0) == cudaSuccess;
  if (graph) cudaGraphDestroy(graph);
  // the capture ran the host-side bookkeeping of the steps without running the steps: restore it
  g.launches = c.stats.kernel_launches - before.kernel_launches;
  g.nccl_calls = c.stats.nccl_calls - before.nccl_calls;
  g.gemm_launches = c.stats.n_gemm_launches - before.n_gemm_launches;
  g.feat_launches = c.stats.n_featurize_launches - before.n_featurize_launches;
  c.stats = before;
  c.folded_valid = fv;
  c.tc_weights_valid = tv;
  c.weights_in_flight = false;
  if (c.tcs && c.weights_swapped != sw) {
    std::swap(c.tcs->wF, c.tcs->wF_alt);
    std::swap(c.tcs->wD, c.tcs->wD_alt);
    c.weights_swapped = sw;
  }
  if (!ok || realloc) {
    cudaGetLastError();
    if (g.exec) cudaGraphExecDestroy(g.exec);
    g.exec = nullptr;
    if (!ok) c.graph_mode = 0;  // capture itself failed: stay eager; after a reallocation the next epoch retries
    return;
  }
  g.N = N; g.bs = bs; g.nb = nb;
  g.xs = c.xs; g.target = c.target.p; g.perm = c.perm_dev.p;
  g.overlap = overlap;
  g.alloc_gen = gen;
}

// the indices feed device gathers (coords + idx*D, target[idx]): reject anything outside 1..N before the upload
void validate_perm(const int64_t *perm_host, int64_t N) {
  IK_REQUIRE(perm_host != nullptr, ISOKANN_BAD_ARGUMENT, "perm must not be NULL");
  int64_t lo = INT64_MAX, hi = INT64_MIN;
  for (int64_t i = 0; i < N; ++i) {
    lo = std::min(lo, perm_host[i]);
    hi = std::max(hi, perm_host[i]);
  }
  IK_REQUIRE(lo >= 1 && hi <= N, ISOKANN_BAD_ARGUMENT,
             "perm must hold 1-based indices in 1..N (got " + std::to_string(lo) + ".." + std::to_string(hi) + ")");
}

// isokann_iterate knows the permutation of the coming epoch before the Koopman pass starts: its upload (8 bytes per
// start point, from pageable caller memory) runs on the copy stream beside that pass instead of between the target
// and the first optimiser step
void preload_perm(Ctx &c, const int64_t *perm_host) {
  validate_perm(perm_host, c.N);
  if (!c.copy_stream) IK_CUDA(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
  if (!c.perm_event) {
    IK_CUDA(cudaEventCreateWithFlags(&c.perm_event, cudaEventDisableTiming));
    IK_CUDA(cudaEventCreateWithFlags(&c.perm0_event, cudaEventDisableTiming));
  }
  c.perm_raw.ensure((size_t)c.N);
  c.perm_dev.ensure((size_t)c.N);
  // through a page-locked staging buffer: a copy from pageable memory would first drain the copy stream (which may
  // hold gigabytes of isokann_set_data_async uploads) and block this thread until then
  const size_t bytes = (size_t)c.N * sizeof(int64_t);
  if (bytes > c.perm_pinned_bytes) {
    if (c.perm_pinned) cudaFreeHost(c.perm_pinned);
    c.perm_pinned = nullptr;
    c.perm_pinned_bytes = 0;
    IK_CUDA(cudaMallocHost(&c.perm_pinned, bytes));
    c.perm_pinned_bytes = bytes;
  } else if (c.perm_staged) {
    IK_CUDA(cudaEventSynchronize(c.perm_event));  // the previous permutation has left the staging buffer
  }
  memcpy(c.perm_pinned, perm_host, bytes);
  if (c.perm0_recorded) IK_CUDA(cudaStreamWaitEvent(c.copy_stream, c.perm0_event, 0));  // perm_raw is free again
  IK_CUDA(cudaMemcpyAsync(c.perm_raw.p, c.perm_pinned, bytes, cudaMemcpyHostToDevice, c.copy_stream));
  IK_CUDA(cudaEventRecord(c.perm_event, c.copy_stream));
  c.perm_staged = true;
  c.perm_preloaded = perm_host;
  c.perm_preloaded_n = c.N;
}

// next_perm: the permutation of the epoch after this one, if the caller already knows it (isokann_iterate): its
// upload is started while this epoch's steps are still running
double train_epoch(Ctx &c, const int64_t *perm_host, int64_t minibatch, bool partial, const int64_t *next_perm = nullptr) {
  IK_REQUIRE(c.xs != nullptr, ISOKANN_ERR_STATE, "no data: call isokann_set_data first");
  IK_REQUIRE(c.has_target, ISOKANN_ERR_STATE, "no target: call isokann_target / isokann_set_target first");
  gather_xs(c);
  IK_REQUIRE(perm_host != nullptr, ISOKANN_BAD_ARGUMENT, "perm must not be NULL");
  IK_REQUIRE(minibatch >= 0, ISOKANN_BAD_ARGUMENT, "minibatch must be >= 0");
  const int64_t N = c.N;
  const bool preloaded = c.perm_preloaded == perm_host && c.perm_preloaded_n == N;
  c.perm_preloaded = nullptr;
  if (!preloaded) validate_perm(perm_host, N);
  const int64_t bs = (minibatch == 0 || N < minibatch) ? N : minibatch;  // src/iso.jl:180
  const int64_t nb = partial ? (N + bs - 1) / bs : N / bs;               // partial=false drops the tail
  c.timer.begin(KC_PHASE_TRAIN, c.stream);
  c.perm_raw.ensure((size_t)N);
  c.perm_dev.ensure((size_t)N);
  if (preloaded)
    IK_CUDA(cudaStreamWaitEvent(c.stream, c.perm_event, 0));
  else
    IK_CUDA(cudaMemcpyAsync(c.perm_raw.p, perm_host, (size_t)N * sizeof(int64_t), cudaMemcpyHostToDevice, c.stream));
  launch_perm_to_zero_based(c, c.perm_raw.p, N, c.perm_dev.p);
  if (c.perm0_event) {
    IK_CUDA(cudaEventRecord(c.perm0_event, c.stream));
    c.perm0_recorded = true;
  }
  IK_CUDA(cudaMemsetAsync(c.epoch_loss.p, 0, sizeof(double), c.stream));
  {
    // every rank gets rows in every step (bs >= world), so all ranks take the same code path
    const bool overlap = c.world > 1 && c.tc && !c.tcn && c.L >= 2 && !c.no_comm_overlap && bs >= c.world;
    struct Scope {
      Ctx &c;
      ~Scope() {
        c.comm_overlap = false;
        c.sm_reserve = 0;
      }
    } scope{c};
    c.comm_overlap = overlap;
    auto steps = [&] {
      for (int64_t i = 0; i < nb; ++i) {
        const int64_t start = i * bs;
        train_step(c, start, std::min(bs, N - start));
      }
      join_comm_stream(c);
    };
    // canonical starting state of an epoch (what a captured graph assumes): operand sets in their original
    // order and the folded / split weights current
    if (c.tcs && c.weights_swapped) {
      std::swap(c.tcs->wF, c.tcs->wF_alt);
      std::swap(c.tcs->wD, c.tcs->wD_alt);
      c.weights_swapped = false;
      c.tc_weights_valid = false;
    }
    Ctx::EpochGraph &g = c.egraph;
    const bool want_graph = c.graph_mode && (!c.timer.enabled || c.timer.phases_only) && nb >= 2 && !c.step_trace;
    const bool same = g.exec && g.N == N && g.bs == bs && g.nb == nb && g.xs == c.xs && g.target == c.target.p &&
                      g.perm == c.perm_dev.p && g.overlap == overlap && g.alloc_gen == alloc_generation();
    if (want_graph && (same || c.eager_epochs >= 1)) {
      ensure_folded(c);
      if (c.tc || c.tcn) ensure_tc_weights(c);
      if (!same) capture_epoch(c, steps, N, bs, nb, overlap);
    }
    if (want_graph && g.exec && g.N == N && g.bs == bs && g.nb == nb && g.overlap == overlap &&
        g.alloc_gen == alloc_generation() && g.xs == c.xs && g.target == c.target.p && g.perm == c.perm_dev.p) {
      IK_CUDA(cudaGraphLaunch(g.exec, c.stream));
      c.stats.kernel_launches += g.launches;
      c.stats.nccl_calls += g.nccl_calls;
      c.stats.n_gemm_launches += g.gemm_launches;
      c.stats.n_featurize_launches += g.feat_launches;
      c.stats.graph_launches++;
      // host-side mirrors of what the replayed steps did
      c.folded_valid = false;
      c.tc_weights_valid = false;
      c.wf16_valid = false;
      if (overlap) {
        c.folded_valid = c.tc_weights_valid = true;
        if (nb & 1) {
          std::swap(c.tcs->wF, c.tcs->wF_alt);
          std::swap(c.tcs->wD, c.tcs->wD_alt);
          c.weights_swapped = true;
        }
      }
    } else {
      steps();
      c.eager_epochs++;
      trace_report(c);
    }
  }
  c.timer.end(c.stream);
  if (next_perm) preload_perm(c, next_perm);  // host-side staging and the DMA run beside this epoch's steps
  double *lp = read_back(c, c.epoch_loss.p, 1);
  const double ls = *lp;
  const int f = check_flags(c);
  if (f & FLAG_NONFINITE_LOSS)
    throw ik::Error{ISOKANN_DOMAIN_NONFINITE_LOSS,
                    "The ISOKANN model collapsed under training. Try reducing the learning rate or increasing "
                    "regularization"};
  return ls / (double)N;  // src/iso.jl:193
}

void upload_rows(Ctx &c, const void *host, bool f64, int64_t count, float *dev) {
  if (count <= 0) return;
  if (!f64) {
    IK_CUDA(cudaMemcpyAsync(dev, host, (size_t)count * sizeof(float), cudaMemcpyHostToDevice, c.stream));
    sync_stream(c);
    return;
  }
  // Float64 coordinates: the reference featurizes in the input eltype and casts the features to
  // Float32 (src/simulation.jl:112); here coordinates are rounded to Float32 once, on the device (with direct
  // differences that keeps distances within 2e-6 relative of the Float64 result).  The doubles go up in 32 MiB
  // pieces through a device staging buffer, so the host does no conversion work and never waits per piece.
  const int64_t blk = 1 << 22;
  c.staging_f64.ensure((size_t)std::min(blk, count));
  const double *src = (const double *)host;
  for (int64_t o = 0; o < count; o += blk) {
    const int64_t n = std::min(blk, count - o);
    IK_CUDA(cudaMemcpyAsync(c.staging_f64.p, src + o, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    launch_f64_to_f32(c, c.staging_f64.p, n, dev + o);
  }
  sync_stream(c);
}

// Page-lock caller memory for the asynchronous upload (cudaHostRegister) so that the DMA engine reads it directly
// and the copy really overlaps the Koopman pass.  Registrations are cached per (pointer, size): a caller that hands
// over the same arrays every iteration (Julia's iso.data) pays the pinning once.  Memory that is already page-locked
// (cudaMallocHost, torch pin_memory) is left alone; if pinning fails the copy still works, staged by the driver.
void release_host_registrations(Ctx &c) {
  for (auto &r : c.host_regs)
    if (r.ours) cudaHostUnregister(r.ptr);
  c.host_regs.clear();
  cudaGetLastError();
}

void ensure_host_registered(Ctx &c, int slot, const void *ptr, size_t bytes) {
  if ((int)c.host_regs.size() <= slot) c.host_regs.resize(slot + 1);
  Ctx::HostReg &r = c.host_regs[slot];
  if (r.ptr == ptr && r.bytes >= bytes) return;
  if (r.ours) cudaHostUnregister(r.ptr);
  r = Ctx::HostReg{};
  cudaGetLastError();
  if (!ptr || bytes == 0) return;
  cudaPointerAttributes at{};
  if (cudaPointerGetAttributes(&at, ptr) == cudaSuccess && at.type != cudaMemoryTypeUnregistered) {
    r.ptr = const_cast<void *>(ptr);  // already page-locked (or managed) by its owner
    r.bytes = bytes;
    return;
  }
  cudaGetLastError();
  const cudaError_t e = cudaHostRegister(const_cast<void *>(ptr), bytes, cudaHostRegisterDefault);
  r.ptr = const_cast<void *>(ptr);
  r.bytes = bytes;
  r.ours = e == cudaSuccess;
  if (e != cudaSuccess) cudaGetLastError();
}

void set_data_impl(Ctx &c, const void *xs, const void *ys, bool f64, bool dev_ptrs, int64_t D, int64_t K, int64_t N,
                   int64_t n_off, int64_t n_loc, bool async_ys = false) {
  if (c.ys_chunk_pts > 0 || c.xs_pending) {  // a previous asynchronous upload may still be running
    IK_CUDA(cudaStreamSynchronize(c.copy_stream));
    c.ys_chunk_pts = 0;
    c.xs_pending = false;
  }
  c.xs_gather_pending = false;
  IK_REQUIRE(xs != nullptr, ISOKANN_BAD_ARGUMENT, "xs must not be NULL");
  IK_REQUIRE(D == c.D, ISOKANN_BAD_ARGUMENT, "coordinate dimension does not match the featurizer/model");
  IK_REQUIRE(N > 0 && K >= 0, ISOKANN_BAD_ARGUMENT, "N must be positive");
  int64_t eo, el;
  split_range(N, c.world, c.rank, &eo, &el);
  IK_REQUIRE(n_off == eo && n_loc == el, ISOKANN_BAD_ARGUMENT,
             "shard does not follow the contiguous split rule (see isokann_set_data_sharded)");
  // while the upload is under way the context holds no data: if anything below throws, later calls report
  // "no data" instead of running on a half-updated shard description
  c.has_target = false;
  c.has_weights = false;
  c.xs = nullptr;
  c.ys = nullptr;
  if (dev_ptrs) {
    c.N = N; c.K = K; c.n_off = n_off; c.n_loc = n_loc;
    c.xs = (const float *)xs;
    c.ys = (const float *)ys;
    return;
  }
  c.xs_own.ensure((size_t)N * D);
  const bool have_ys = K > 0 && (ys != nullptr || n_loc == 0);
  // the asynchronous path (and with it the later all-gather of xs) must be taken by every rank or by none: decide on
  // values that are the same everywhere (K, element type, the entry point), never on this rank's shard size
  const bool async_all = async_ys && !f64 && K > 0;
  IK_REQUIRE(!async_all || have_ys, ISOKANN_BAD_ARGUMENT, "isokann_set_data_async needs the Koopman samples of this shard");
  if (!async_all) upload_rows(c, xs, f64, N * D, c.xs_own.p);
  // a rank whose shard is empty (N < world) still owns a (1-element) sample buffer, so that it enters the Koopman
  // pass and its collectives like everybody else
  if (have_ys) c.ys_own.ensure((size_t)std::max<int64_t>(1, n_loc * K * D));
  if (async_all) {
    // stream ys in on a second stream, ~64 MiB per chunk, one event per chunk; the Koopman pass waits per
    // chunk, so the PCIe transfer overlaps the forward pass over the chunks that already arrived
    if (!c.copy_stream) IK_CUDA(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
    ensure_host_registered(c, 0, n_loc > 0 ? ys : nullptr, (size_t)n_loc * K * D * sizeof(float));
    ensure_host_registered(c, 1, xs, (size_t)N * D * sizeof(float));
    const int64_t pts = std::max<int64_t>(1, (64ll << 20) / (K * D * 4));
    const int64_t nchunks = (n_loc + pts - 1) / pts;
    while ((int64_t)c.ys_events.size() < nchunks) {
      cudaEvent_t e;
      IK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      c.ys_events.push_back(e);
    }
    const float *src = (const float *)ys;
    for (int64_t i = 0; i < nchunks; ++i) {
      const int64_t p0 = i * pts, np = std::min(pts, n_loc - p0);
      IK_CUDA(cudaMemcpyAsync(c.ys_own.p + p0 * K * D, src + p0 * K * D, (size_t)np * K * D * sizeof(float),
                              cudaMemcpyHostToDevice, c.copy_stream));
      IK_CUDA(cudaEventRecord(c.ys_events[(size_t)i], c.copy_stream));
    }
    // xs last: the Koopman pass reads ys only, so xs arrives while that pass is already running
    if (!c.xs_event) IK_CUDA(cudaEventCreateWithFlags(&c.xs_event, cudaEventDisableTiming));
    if (c.world > 1) {  // only this rank's rows cross PCIe; gather_xs() fetches the rest from the peers
      if (n_loc > 0)
        IK_CUDA(cudaMemcpyAsync(c.xs_own.p + n_off * D, (const float *)xs + n_off * D,
                                (size_t)n_loc * D * sizeof(float), cudaMemcpyHostToDevice, c.copy_stream));
    } else {
      IK_CUDA(cudaMemcpyAsync(c.xs_own.p, xs, (size_t)N * D * sizeof(float), cudaMemcpyHostToDevice, c.copy_stream));
    }
    IK_CUDA(cudaEventRecord(c.xs_event, c.copy_stream));
    c.ys_chunk_pts = nchunks > 0 ? pts : 0;
    c.ys_chunks_pending = nchunks;
    c.xs_gather_pending = c.world > 1;
    c.xs_pending = true;
  } else if (K > 0 && n_loc > 0 && ys != nullptr) {
    upload_rows(c, ys, f64, n_loc * K * D, c.ys_own.p);
  }
  c.N = N; c.K = K; c.n_off = n_off; c.n_loc = n_loc;
  c.xs = c.xs_own.p;
  c.ys = have_ys ? c.ys_own.p : nullptr;
}

// Incremental data on a multi-rank context (isokann_append_data / isokann_keep_last): rebuild this rank's shard of
// ys for the new global range.  New start point i (0 <= i < N_new) is the old start point i + shift while
// i + shift < N_old, else row i + shift - N_old of the appended block (host, the same on every rank).  The contiguous
// split moves every shard boundary, so rows change owner; the old shards are all-gathered once over NVLink (the same
// padded collective as the Koopman vector) and every rank cuts its new range out of the gathered copy.  xs is
// replicated and handled by the caller.
void reshard_ys(Ctx &c, int64_t N_old, int64_t shift, int64_t N_new, const float *ys_new_host) {
  const int64_t KD = c.K * c.D;
  int64_t o1, l1;
  split_range(N_new, c.world, c.rank, &o1, &l1);
  if (KD > 0) {
    IK_REQUIRE(KD <= INT32_MAX, ISOKANN_BAD_ARGUMENT, "K * D too large");
    DevBuf<float> full, nb;
    full.ensure((size_t)std::max<int64_t>(1, N_old * KD));
    allgather_rows(c, c.ys_own.p, c.n_loc, full.p, (int)KD);  // c.N is still N_old here
    nb.ensure((size_t)std::max<int64_t>(1, l1 * KD));
    const int64_t a = o1 + shift, b = std::min(o1 + l1 + shift, N_old);  // rows taken from the old data
    if (b > a)
      IK_CUDA(cudaMemcpyAsync(nb.p, full.p + a * KD, (size_t)(b - a) * KD * sizeof(float), cudaMemcpyDeviceToDevice,
                              c.stream));
    const int64_t a2 = std::max(o1 + shift, N_old), b2 = o1 + l1 + shift;  // rows taken from the appended block
    if (b2 > a2) {
      IK_REQUIRE(ys_new_host != nullptr, ISOKANN_BAD_ARGUMENT, "appended Koopman samples missing");
      IK_CUDA(cudaMemcpyAsync(nb.p + (a2 - shift - o1) * KD, ys_new_host + (a2 - N_old) * KD,
                              (size_t)(b2 - a2) * KD * sizeof(float), cudaMemcpyHostToDevice, c.stream));
    }
    sync_stream(c);
    full.release();
    c.gather_pad.release();  // it held (world + 1) padded shards of ys; the per-iteration gathers need far less
    c.ys_own.release();
    c.ys_own = nb;
    c.ys = c.ys_own.p;
  }
  c.N = N_new;
  c.n_off = o1;
  c.n_loc = l1;
}

void build_pair_table(Ctx &c) {
  const isokann_config &g = c.cfg;
  std::vector<int2> tab;
  auto upper = [&](const std::vector<int> &atoms) {  // column-major strict upper triangle (halfinds)
    const int n = (int)atoms.size();
    c.tri_n = n;
    if (g.featurizer == ISOKANN_FEAT_ATOMS && n > 0) {
      std::vector<int> cmap(3 * (size_t)n);
      for (int a = 0; a < n; ++a)
        for (int k = 0; k < 3; ++k) cmap[3 * a + k] = 3 * atoms[a] + k;
      c.tri_cmap.ensure(cmap.size());
      IK_CUDA(cudaMemcpy(c.tri_cmap.p, cmap.data(), cmap.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    for (int j = 1; j < n; ++j)
      for (int i = 0; i < j; ++i) tab.push_back(make_int2(3 * atoms[i], 3 * atoms[j]));
  };
  if (g.featurizer == ISOKANN_FEAT_ALLPAIRS) {
    std::vector<int> atoms(g.n_atoms);
    for (int i = 0; i < g.n_atoms; ++i) atoms[i] = i;
    upper(atoms);
  } else if (g.featurizer == ISOKANN_FEAT_ATOMS) {
    std::vector<int> atoms;
    for (int i = 0; i < g.n_index; ++i) {
      IK_REQUIRE(c.index[i] >= 1 && c.index[i] <= g.n_atoms, ISOKANN_BAD_ARGUMENT, "atom index out of range");
      atoms.push_back(c.index[i] - 1);
    }
    upper(atoms);
  } else if (g.featurizer == ISOKANN_FEAT_PAIRS) {
    for (int i = 0; i < g.n_index; ++i) {
      const int a = c.index[2 * i], b = c.index[2 * i + 1];
      IK_REQUIRE(a >= 1 && a <= g.n_atoms && b >= 1 && b <= g.n_atoms, ISOKANN_BAD_ARGUMENT,
                 "pair index out of range");
      tab.push_back(make_int2(3 * (a - 1), 3 * (b - 1)));
    }
  }
  c.n_pairs = (int)tab.size();
  if (!tab.empty()) {
    c.pairs.ensure(tab.size());
    IK_CUDA(cudaMemcpy(c.pairs.p, tab.data(), tab.size() * sizeof(int2), cudaMemcpyHostToDevice));
    // atom -> incident features, for the featurizer pullback
    const int A = g.n_atoms;
    std::vector<int> off(A + 1, 0);
    for (const int2 &p : tab) {
      off[p.x / 3 + 1]++;
      off[p.y / 3 + 1]++;
    }
    for (int a = 0; a < A; ++a) off[a + 1] += off[a];
    std::vector<int2> adj(2 * tab.size());
    std::vector<int> fill(off.begin(), off.end() - 1);
    for (int f = 0; f < (int)tab.size(); ++f) {
      adj[fill[tab[f].x / 3]++] = make_int2(f, tab[f].y);
      adj[fill[tab[f].y / 3]++] = make_int2(f, tab[f].x);
    }
    c.adj_off.ensure(off.size());
    c.adj.ensure(adj.size());
    IK_CUDA(cudaMemcpy(c.adj_off.p, off.data(), off.size() * sizeof(int), cudaMemcpyHostToDevice));
    IK_CUDA(cudaMemcpy(c.adj.p, adj.data(), adj.size() * sizeof(int2), cudaMemcpyHostToDevice));
  }
}

// ---- diagnostics on the resident data (loggers) --------------------------------------------------
// chi = chis(iso) and Kchi = koopman(iso) on the resident data (both N x d on every rank), then the second moments
// of u = [chi, 1] and v = [Kchi, 1]: uu = sum u u', vu = sum v u', vv = sum v v' (row-major (d+1) x (d+1), fp64)
struct Moments {
  double uu[(kMaxD + 1) * (kMaxD + 1)], vu[(kMaxD + 1) * (kMaxD + 1)], vv[(kMaxD + 1) * (kMaxD + 1)];
};

void diag_moments(Ctx &c, Moments &mo) {
  IK_REQUIRE(c.d >= 1 && c.d <= kMaxD, ISOKANN_BAD_ARGUMENT, "chi dimension exceeds ISOKANN_MAX_D");
  compute_chis(c);
  compute_koopman(c);
  const int d = c.d, e = d + 1;
  c.diag_part.ensure((size_t)256 * 3 * (kMaxD + 1) * (kMaxD + 1));
  int nb = 0;
  launch_moments(c, c.chi_x.p, c.kchi.p, c.N, d, c.diag_part.p, &nb);
  double *part = read_back(c, c.diag_part.p, (size_t)nb * e * 3 * e);
  for (int i = 0; i < e * e; ++i) mo.uu[i] = mo.vu[i] = mo.vv[i] = 0.0;
  for (int b = 0; b < nb; ++b)
    for (int a = 0; a < e; ++a)
      for (int k = 0; k < e; ++k) {
        const double *q = part + (((size_t)b * e + a) * 3) * e + k;
        mo.uu[a * e + k] += q[0];
        mo.vu[a * e + k] += q[e];
        mo.vv[a * e + k] += q[2 * e];
      }
  for (int i = 0; i < e * e; ++i)
    IK_REQUIRE(std::isfinite(mo.uu[i]) && std::isfinite(mo.vu[i]) && std::isfinite(mo.vv[i]), ISOKANN_BAD_ARGUMENT,
               "chi or Kchi is not finite");
}

// one pass r = A Kchi - B chi over the resident data: column sums of r^2 and of (A Kchi)^2, optional N x d output
void diag_resid_pass(Ctx &c, const Mat8 &A, const Mat8 &B, double *dev_out, double *sum_r2, double *sum_k2) {
  const int d = c.d;
  int nb = 0;
  launch_resid(c, c.kchi.p, c.chi_x.p, c.N, d, A, B, dev_out, c.diag_part.p, &nb);
  double *part = read_back(c, c.diag_part.p, (size_t)nb * d * 2);
  for (int j = 0; j < d; ++j) sum_r2[j] = sum_k2[j] = 0.0;
  for (int b = 0; b < nb; ++b)
    for (int j = 0; j < d; ++j) {
      sum_r2[j] += part[((size_t)b * d + j) * 2];
      sum_k2[j] += part[((size_t)b * d + j) * 2 + 1];
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------
// exported entry points
// ------------------------------------------------------------------------------------------
extern "C" {

int32_t isokann_abi_version(void) { return ISOKANN_ABI_VERSION; }

const char *isokann_last_error(const isokann_ctx *ctx) {
  if (!ctx) return g_create_err.c_str();
  return ctx->err.c_str();
}

int32_t isokann_create(const isokann_config *cfg, isokann_ctx **out) {
  if (!cfg || !out) return ISOKANN_BAD_ARGUMENT;
  *out = nullptr;
  isokann_ctx *c = nullptr;
  try {
    IK_REQUIRE(cfg->n_layers >= 1 && cfg->n_layers <= ISOKANN_MAX_LAYERS, ISOKANN_BAD_ARGUMENT, "n_layers out of range");
    for (int l = 0; l <= cfg->n_layers; ++l)
      IK_REQUIRE(cfg->widths[l] >= 1, ISOKANN_BAD_ARGUMENT, "layer widths must be positive");
    IK_REQUIRE(cfg->optimiser == ISOKANN_OPT_ADAM || cfg->optimiser == ISOKANN_OPT_NESTEROV, ISOKANN_BAD_ARGUMENT,
               "unknown optimiser");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
      throw ik::Error{ISOKANN_ERR_CUDA, std::string("no CUDA device available (there is no CPU fallback): ") +
                                            cudaGetErrorString(e)};
    IK_REQUIRE(cfg->device >= 0 && cfg->device < ndev, ISOKANN_BAD_ARGUMENT, "device ordinal out of range");
    c = new isokann_ctx;
    c->cfg = *cfg;
    c->dev = cfg->device;
    IK_CUDA(cudaSetDevice(c->dev));
    cudaDeviceProp prop;
    IK_CUDA(cudaGetDeviceProperties(&prop, c->dev));
    c->num_sms = prop.multiProcessorCount;
    IK_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    const int nidx = cfg->featurizer == ISOKANN_FEAT_PAIRS ? 2 * cfg->n_index
                                                           : (cfg->featurizer == ISOKANN_FEAT_ATOMS ? cfg->n_index : 0);
    if (nidx > 0) {
      IK_REQUIRE(cfg->index != nullptr, ISOKANN_BAD_ARGUMENT, "featurizer index list missing");
      c->index.assign(cfg->index, cfg->index + nidx);
    }
    c->cfg.index = nullptr;
    c->L = cfg->n_layers;
    c->F = cfg->widths[0];
    c->d = cfg->widths[c->L];
    c->ln = cfg->layernorm != 0;
    if (c->cfg.ln_eps <= 0.f) c->cfg.ln_eps = 1e-5f;
    if (cfg->featurizer == ISOKANN_FEAT_IDENTITY) {
      c->D = c->F;
    } else {
      IK_REQUIRE(cfg->n_atoms >= 2, ISOKANN_BAD_ARGUMENT, "n_atoms must be >= 2 for distance featurizers");
      c->D = 3 * cfg->n_atoms;
      build_pair_table(*c);
      IK_REQUIRE(c->n_pairs == c->F, ISOKANN_BAD_ARGUMENT,
                 "model input width does not match the number of pair-distance features");
    }
    int64_t off = 0;
    if (c->ln) {
      c->off_gamma = off; off += c->F;
      c->off_beta = off; off += c->F;
    }
    c->maxw = 0;
    for (int l = 0; l < c->L; ++l) {
      const int64_t fin = cfg->widths[l], fout = cfg->widths[l + 1];
      c->off_w.push_back(off); off += fin * fout;
      c->off_b.push_back(off); off += fout;
    }
    for (int l = 0; l <= c->L; ++l) c->maxw = std::max(c->maxw, cfg->widths[l]);
    c->P = off;
    c->params.ensure((size_t)c->P);
    c->grads.ensure((size_t)c->P + 4);
    c->opt_m.ensure((size_t)c->P);
    c->opt_v.ensure((size_t)c->P);
    IK_CUDA(cudaMemset(c->params.p, 0, (size_t)c->P * sizeof(float)));
    IK_CUDA(cudaMemset(c->grads.p, 0, ((size_t)c->P + 4) * sizeof(float)));
    IK_CUDA(cudaMemset(c->opt_m.p, 0, (size_t)c->P * sizeof(float)));
    IK_CUDA(cudaMemset(c->opt_v.p, 0, (size_t)c->P * sizeof(float)));
    if (c->ln) {
      const size_t seg = (size_t)(c->F + 1) * cfg->widths[1];
      c->folded1.ensure(seg);
      c->gfold.ensure(seg);
    }
    if (cfg->gemm_mode == ISOKANN_GEMM_TC)
      IK_REQUIRE(tc_eligible(*cfg, true), ISOKANN_BAD_ARGUMENT,
                 "ISOKANN_GEMM_TC needs >= 2 layers, hidden widths >= 256, input width >= 64, output <= 8");
    c->tc = cfg->gemm_mode != ISOKANN_GEMM_FP32 && tc_eligible(*cfg, cfg->gemm_mode == ISOKANN_GEMM_TC);
    c->tcn = cfg->gemm_mode == ISOKANN_GEMM_AUTO && !tc_eligible(*cfg, true) && tcn_eligible(*cfg);
    c->fused_train = cfg->gemm_mode == ISOKANN_GEMM_AUTO && !tc_eligible(*cfg, true) && narrow_train_eligible(*cfg);
    c->tiny = cfg->gemm_mode == ISOKANN_GEMM_AUTO && tiny_forward_eligible(*cfg);
    { const char *e = getenv("ISOKANN_GRAPH"); c->graph_mode = !(e && e[0] == '0'); }
    { const char *e = getenv("ISOKANN_TC_FWD"); c->fwd_fp16x2 = e && strcmp(e, "fp16x2") == 0; }
    c->tc_no_pair = getenv("ISOKANN_TC_NO_PAIR") != nullptr;
    c->tc_no_head = getenv("ISOKANN_TC_NO_HEAD") != nullptr;
    { const char *e = getenv("ISOKANN_FEAT_REC"); c->feat_rec_off = e && e[0] == '0'; }
    { const char *e = getenv("ISOKANN_KOOP_FUSED"); c->koop_fused_off = e && e[0] == '0'; }
    c->tc_no_overlap = getenv("ISOKANN_OVERLAP") == nullptr;  // opt-in: measured +0.6 % under the 1 kW cap
    if (c->tc || c->tcn) {
      c->tcs = new TcState;
      c->tcs->act.resize(c->L);
      c->tcs->wF.resize(c->L);
      c->tcs->wD.resize(c->L);
      c->tcs->wF_alt.resize(c->L);
      c->tcs->wD_alt.resize(c->L);
      // padded row length: room for a constant-1 column behind the activations (bias column of the wgrad GEMM)
      for (int l = 0; l <= c->L; ++l) c->tcs->wp.push_back((cfg->widths[l] + 1 + 63) & ~63);
    }
    c->beta_dev.ensure(2);
    {
      const float bt[2] = {cfg->beta1, cfg->beta2};
      IK_CUDA(cudaMemcpy(c->beta_dev.p, bt, sizeof(bt), cudaMemcpyHostToDevice));
    }
    c->act.resize(c->L + 1);
    c->flags.ensure(1);
    c->ticket.ensure(1);
    c->epoch_loss.ensure(1);
    IK_CUDA(cudaMemset(c->flags.p, 0, sizeof(int)));
    IK_CUDA(cudaMemset(c->ticket.p, 0, sizeof(unsigned int)));
    IK_CUDA(cudaMemset(c->epoch_loss.p, 0, sizeof(double)));
    c->red_f.ensure(2 * 1024);
    c->red_d.ensure(256 * kMaxD * 2 * kMaxD + 1024);
    c->red_am.ensure(512);
    set_unit_weights(*c);
    IK_CUDA(cudaStreamSynchronize(c->stream));
    *out = c;
    return ISOKANN_OK;
  } catch (const ik::Error &e) {
    g_create_err = e.msg;
    if (c) {
      delete c;
    }
    return e.code;
  }
}

int32_t isokann_destroy(isokann_ctx *c) {
  if (!c) return ISOKANN_OK;
  cudaSetDevice(c->dev);
  cudaStreamSynchronize(c->stream);
  if (c->egraph.exec) cudaGraphExecDestroy(c->egraph.exec);
  for (auto e : c->trace_ev) cudaEventDestroy(e);
  for (void *h : c->p2p.opened) cudaIpcCloseMemHandle(h);
  c->p2p.flag_block.release();
  c->p2p.seq.release();
  c->p2p.ticket.release();
  if (c->comm_ov) nccl_comm_destroy(c->nccl, c->comm_ov);
  if (c->comm) nccl_comm_destroy(c->nccl, c->comm);
  c->timer.destroy();
  for (auto &a : c->act) a.release();
  DevBuf<float> *fb[] = {&c->params, &c->grads, &c->opt_m, &c->opt_v, &c->folded1, &c->gfold, &c->xs_own, &c->ys_own,
                         &c->kweights, &c->chi_x, &c->kchi, &c->kchi_loc, &c->gather_pad, &c->target, &c->w_loss,
                         &c->delta_a, &c->delta_b, &c->splitk, &c->staging_in, &c->staging_out, &c->red_f,
                         &c->xs_stage, &c->val_chi, &c->val_k1, &c->beta_dev};
  c->staging_f64.release();
  c->diag_part.release();
  c->diag_out.release();
  for (auto *b : fb) b->release();
  if (c->tcs) {
    for (auto &b : c->tcs->act) b.release();
    for (auto &b : c->tcs->wF) b.release();
    for (auto &b : c->tcs->wD) b.release();
    for (auto &b : c->tcs->wF_alt) b.release();
    for (auto &b : c->tcs->wD_alt) b.release();
    for (auto &b : c->tcs->wF16) b.release();
    c->tcs->delta[0].release();
    c->tcs->delta[1].release();
    c->tcs->dot_partial.release();
    c->tcs->dlast.release();
    c->tcs->x_alt.release();
    if (c->tcs->feat_stream) {
      cudaStreamSynchronize(c->tcs->feat_stream);
      cudaStreamDestroy(c->tcs->feat_stream);
      for (int b = 0; b < 2; ++b) {
        cudaEventDestroy(c->tcs->feat_done[b]);
        cudaEventDestroy(c->tcs->x_free[b]);
      }
      cudaEventDestroy(c->tcs->koop_start);
    }
    delete c->tcs;
  }
  c->tri_cmap.release();
  c->koop_start16.release();
  c->pairs.release();
  c->adj_off.release();
  c->adj.release();
  c->red_d.release();
  c->epoch_loss.release();
  c->red_am.release();
  c->perm_dev.release();
  c->perm_raw.release();
  c->flags.release();
  c->ticket.release();
  if (c->copy_stream) {
    cudaStreamSynchronize(c->copy_stream);
    cudaStreamDestroy(c->copy_stream);
  }
  if (c->comm_stream) {
    cudaStreamSynchronize(c->comm_stream);
    cudaStreamDestroy(c->comm_stream);
    cudaEventDestroy(c->ev_upper);
    cudaEventDestroy(c->ev_lower);
    cudaEventDestroy(c->ev_weights);
  }
  for (auto e : c->ys_events) cudaEventDestroy(e);
  if (c->xs_event) cudaEventDestroy(c->xs_event);
  if (c->perm_event) cudaEventDestroy(c->perm_event);
  if (c->perm0_event) cudaEventDestroy(c->perm0_event);
  if (c->perm_pinned) cudaFreeHost(c->perm_pinned);
  release_host_registrations(*c);
  if (c->pinned) cudaFreeHost(c->pinned);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return ISOKANN_OK;
}

int64_t isokann_num_params(const isokann_ctx *c) { return c ? c->P : -1; }
int32_t isokann_feature_dim(const isokann_ctx *c) { return c ? c->F : -1; }
int32_t isokann_coord_dim(const isokann_ctx *c) { return c ? c->D : -1; }

int32_t isokann_comm_get_unique_id(void *id128) {
  if (!id128) return ISOKANN_BAD_ARGUMENT;
  std::string err;
  Nccl *n = nccl_load(err);
  if (!n) {
    g_create_err = err;
    return ISOKANN_ERR_NCCL;
  }
  int rc = nccl_get_unique_id(n, id128, err);
  if (rc != ISOKANN_OK) g_create_err = err;
  return rc;
}

int32_t isokann_comm_init(isokann_ctx *c, int32_t world, int32_t rank, const void *id128) {
  return guarded(c, [&] {
    IK_REQUIRE(world >= 1 && rank >= 0 && rank < world, ISOKANN_BAD_ARGUMENT, "bad world/rank");
    IK_REQUIRE(c->xs == nullptr, ISOKANN_ERR_STATE, "isokann_comm_init must precede isokann_set_data");
    if (world == 1) {
      c->world = 1;
      c->rank = 0;
      return;
    }
    IK_REQUIRE(id128 != nullptr, ISOKANN_BAD_ARGUMENT, "id128 must not be NULL");
    // the gradient all-reduces run beside the GEMMs of the backward pass: bound the SMs NCCL may take and leave
    // exactly that many free (train_step_overlapped); a user setting of NCCL_MAX_CTAS is respected
    c->no_comm_overlap = getenv("ISOKANN_NO_COMM_OVERLAP") != nullptr;
    c->step_trace = getenv("ISOKANN_STEP_TRACE") != nullptr;
    if (const char *e = getenv("ISOKANN_COMM_SMS")) c->comm_sms = std::max(0, std::min(64, atoi(e)));
    std::string err;
    c->nccl = nccl_load(err);
    IK_REQUIRE(c->nccl != nullptr, ISOKANN_ERR_NCCL, err);
    c->comm = nccl_comm_init(c->nccl, world, rank, id128, err);
    IK_REQUIRE(c->comm != nullptr, ISOKANN_ERR_NCCL, err);
    // the bucket that is reduced beside the GEMMs of the backward pass uses a communicator limited to comm_sms
    // CTAs (exactly the SMs those GEMMs leave free); everything else keeps NCCL's own choice
    if (!c->no_comm_overlap && c->comm_sms > 0 && getenv("ISOKANN_COMM_SPLIT_OFF") == nullptr)
      c->comm_ov = nccl_comm_split_limited(c->nccl, c->comm, rank, c->comm_sms);
    c->world = world;
    c->rank = rank;
    if (getenv("ISOKANN_NO_P2P") == nullptr) setup_p2p(*c);
    c->world = world;
    c->rank = rank;
  });
}

int32_t isokann_set_data(isokann_ctx *c, const float *xs, const float *ys, int64_t D, int64_t K, int64_t N) {
  return guarded(c, [&] {
    IK_REQUIRE(c->world == 1, ISOKANN_ERR_STATE, "use isokann_set_data_sharded on a multi-rank context");
    set_data_impl(*c, xs, ys, false, false, D, K, N, 0, N);
  });
}

int32_t isokann_set_data_f64(isokann_ctx *c, const double *xs, const double *ys, int64_t D, int64_t K, int64_t N) {
  return guarded(c, [&] {
    IK_REQUIRE(c->world == 1, ISOKANN_ERR_STATE, "use isokann_set_data_sharded on a multi-rank context");
    set_data_impl(*c, xs, ys, true, false, D, K, N, 0, N);
  });
}

int32_t isokann_set_data_sharded(isokann_ctx *c, const float *xs, const float *ys_local, int64_t D, int64_t K,
                                 int64_t N, int64_t n_offset, int64_t n_local) {
  return guarded(c, [&] { set_data_impl(*c, xs, ys_local, false, false, D, K, N, n_offset, n_local); });
}

int32_t isokann_set_data_async(isokann_ctx *c, const float *xs, const float *ys_local, int64_t D, int64_t K, int64_t N,
                               int64_t n_offset, int64_t n_local) {
  return guarded(c, [&] { set_data_impl(*c, xs, ys_local, false, false, D, K, N, n_offset, n_local, true); });
}

int32_t isokann_append_data(isokann_ctx *c, const float *xs_new, const float *ys_new, int64_t D, int64_t K,
                            int64_t n_new) {
  return guarded(c, [&] {
    IK_REQUIRE(c->xs != nullptr && c->xs == c->xs_own.p && (c->K == 0 || c->ys == c->ys_own.p), ISOKANN_ERR_STATE,
               "append needs library-owned data (isokann_set_data)");
    IK_REQUIRE(xs_new && D == c->D && n_new >= 0 && (c->K == 0 || (ys_new && K == c->K)), ISOKANN_BAD_ARGUMENT,
               "appended block must match D and K of the resident data");
    if (c->ys_chunk_pts > 0 || c->xs_pending) {
      IK_CUDA(cudaStreamSynchronize(c->copy_stream));
      c->ys_chunk_pts = 0;
      c->xs_pending = false;
    }
    if (c->world > 1) gather_xs(*c);  // an asynchronous upload may have left other ranks' rows of xs to be fetched
    if (n_new == 0) return;
    const int64_t N0 = c->N, N1 = N0 + n_new;
    auto grow = [&](DevBuf<float> &buf, int64_t per_point) {
      if ((size_t)(N1 * per_point) <= buf.n) return;
      DevBuf<float> nb;
      nb.ensure((size_t)(std::max<int64_t>(N1, N0 + N0 / 2) * per_point));  // amortised growth
      IK_CUDA(cudaMemcpyAsync(nb.p, buf.p, (size_t)N0 * per_point * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
      sync_stream(*c);
      buf.release();
      buf = nb;
    };
    grow(c->xs_own, D);
    IK_CUDA(cudaMemcpyAsync(c->xs_own.p + N0 * D, xs_new, (size_t)n_new * D * sizeof(float), cudaMemcpyHostToDevice,
                            c->stream));
    c->xs = c->xs_own.p;
    if (c->world > 1) {
      // every rank passes the same appended block; the contiguous split of N0 + n_new moves all shard boundaries
      sync_stream(*c);
      reshard_ys(*c, N0, 0, N1, ys_new);
    } else {
      if (c->K > 0) {
        grow(c->ys_own, c->K * D);
        IK_CUDA(cudaMemcpyAsync(c->ys_own.p + N0 * c->K * D, ys_new, (size_t)n_new * c->K * D * sizeof(float),
                                cudaMemcpyHostToDevice, c->stream));
        c->ys = c->ys_own.p;
      }
      sync_stream(*c);
      c->N = N1;
      c->n_loc = N1;
    }
    c->has_target = false;
    c->has_weights = false;
  });
}

int32_t isokann_keep_last(isokann_ctx *c, int64_t n_keep) {
  return guarded(c, [&] {
    IK_REQUIRE(c->xs != nullptr && c->xs == c->xs_own.p && (c->K == 0 || c->ys == c->ys_own.p), ISOKANN_ERR_STATE,
               "keep_last needs library-owned data (isokann_set_data)");
    IK_REQUIRE(n_keep >= 1, ISOKANN_BAD_ARGUMENT, "n_keep must be positive");
    if (n_keep >= c->N) return;
    if (c->ys_chunk_pts > 0 || c->xs_pending) {
      IK_CUDA(cudaStreamSynchronize(c->copy_stream));
      c->ys_chunk_pts = 0;
      c->xs_pending = false;
    }
    if (c->world > 1) gather_xs(*c);
    const int64_t drop = c->N - n_keep;
    auto shift = [&](DevBuf<float> &buf, int64_t per_point) {
      DevBuf<float> nb;
      nb.ensure(buf.n);
      IK_CUDA(cudaMemcpyAsync(nb.p, buf.p + drop * per_point, (size_t)n_keep * per_point * sizeof(float),
                              cudaMemcpyDeviceToDevice, c->stream));
      sync_stream(*c);
      buf.release();
      buf = nb;
    };
    shift(c->xs_own, c->D);
    c->xs = c->xs_own.p;
    if (c->world > 1) {
      reshard_ys(*c, c->N, drop, n_keep, nullptr);
    } else {
      if (c->K > 0) {
        shift(c->ys_own, c->K * c->D);
        c->ys = c->ys_own.p;
      }
      c->N = n_keep;
      c->n_loc = n_keep;
    }
    c->has_target = false;
    c->has_weights = false;
  });
}

int32_t isokann_chis_prop(isokann_ctx *c, float *chi_out) {
  return guarded(c, [&] {
    IK_REQUIRE(c->ys != nullptr && chi_out, ISOKANN_ERR_STATE, "no Koopman samples resident / NULL output");
    const int64_t ch = chunk_rows(*c, 1);
    const int64_t M = c->n_loc * c->K;
    for (int64_t m0 = 0; m0 < M; m0 += ch) {
      const int64_t m = std::min(ch, M - m0);
      if (c->ys_chunk_pts > 0) {
        const int64_t last = std::min<int64_t>((m0 + m - 1) / c->K / c->ys_chunk_pts, c->ys_chunks_pending - 1);
        IK_CUDA(cudaStreamWaitEvent(c->stream, c->ys_events[(size_t)last], 0));
      }
      forward_rows(*c, c->ys + m0 * c->D, nullptr, 0, m, true);
      IK_CUDA(cudaMemcpyAsync(chi_out + m0 * c->d, c->act[c->L].p, (size_t)m * c->d * sizeof(float),
                              cudaMemcpyDeviceToHost, c->stream));
      sync_stream(*c);
    }
    c->timer.flush(c->stream);
  });
}

int32_t isokann_set_data_dev(isokann_ctx *c, const float *dev_xs, const float *dev_ys_local, int64_t D, int64_t K,
                             int64_t N, int64_t n_offset, int64_t n_local) {
  return guarded(c, [&] { set_data_impl(*c, dev_xs, dev_ys_local, false, true, D, K, N, n_offset, n_local); });
}

int32_t isokann_set_koopman_weights(isokann_ctx *c, const float *w) {
  return guarded(c, [&] {
    if (!w) {
      c->has_weights = false;
      return;
    }
    IK_REQUIRE(c->ys != nullptr, ISOKANN_ERR_STATE, "no Koopman samples resident");
    c->kweights.ensure((size_t)c->n_loc * c->K);
    IK_CUDA(cudaMemcpyAsync(c->kweights.p, w, (size_t)c->n_loc * c->K * sizeof(float), cudaMemcpyHostToDevice,
                            c->stream));
    sync_stream(*c);
    c->has_weights = true;
  });
}

int32_t isokann_upload_params(isokann_ctx *c, const float *flat, int64_t P) {
  return guarded(c, [&] {
    IK_REQUIRE(flat && P == c->P, ISOKANN_BAD_ARGUMENT, "parameter count mismatch");
    IK_CUDA(cudaMemcpyAsync(c->params.p, flat, (size_t)P * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    sync_stream(*c);
    c->folded_valid = false;
    c->tc_weights_valid = false;
    c->wf16_valid = false;
  });
}

int32_t isokann_download_params(isokann_ctx *c, float *flat, int64_t P) {
  return guarded(c, [&] {
    IK_REQUIRE(flat && P == c->P, ISOKANN_BAD_ARGUMENT, "parameter count mismatch");
    IK_CUDA(cudaMemcpyAsync(flat, c->params.p, (size_t)P * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    sync_stream(*c);
  });
}

int32_t isokann_upload_opt_state(isokann_ctx *c, const float *m, const float *v, const float *beta_t, int64_t P) {
  return guarded(c, [&] {
    IK_REQUIRE(m && P == c->P, ISOKANN_BAD_ARGUMENT, "optimiser state size mismatch");
    IK_CUDA(cudaMemcpyAsync(c->opt_m.p, m, (size_t)P * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    if (c->cfg.optimiser == ISOKANN_OPT_ADAM) {
      IK_REQUIRE(v && beta_t, ISOKANN_BAD_ARGUMENT, "Adam state needs v and beta_t");
      IK_CUDA(cudaMemcpyAsync(c->opt_v.p, v, (size_t)P * sizeof(float), cudaMemcpyHostToDevice, c->stream));
      IK_CUDA(cudaMemcpyAsync(c->beta_dev.p, beta_t, 2 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    }
    sync_stream(*c);
  });
}

int32_t isokann_download_opt_state(isokann_ctx *c, float *m, float *v, float *beta_t, int64_t P) {
  return guarded(c, [&] {
    IK_REQUIRE(m && P == c->P, ISOKANN_BAD_ARGUMENT, "optimiser state size mismatch");
    IK_CUDA(cudaMemcpyAsync(m, c->opt_m.p, (size_t)P * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    if (c->cfg.optimiser == ISOKANN_OPT_ADAM) {
      if (v) IK_CUDA(cudaMemcpyAsync(v, c->opt_v.p, (size_t)P * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
      if (beta_t)
        IK_CUDA(cudaMemcpyAsync(beta_t, c->beta_dev.p, 2 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    }
    sync_stream(*c);
  });
}

int32_t isokann_featurize(isokann_ctx *c, const float *coords, int64_t D, int64_t M, float *features_out) {
  return guarded(c, [&] {
    IK_REQUIRE(coords && features_out, ISOKANN_BAD_ARGUMENT, "NULL buffer");
    IK_REQUIRE(D == c->D, ISOKANN_BAD_ARGUMENT, "coordinate dimension does not match the featurizer");
    const int64_t ch = chunk_rows(*c, 1);
    c->staging_in.ensure((size_t)std::min(ch, std::max<int64_t>(M, 1)) * D);
    c->staging_out.ensure((size_t)std::min(ch, std::max<int64_t>(M, 1)) * c->F);
    const bool pairs = c->cfg.featurizer != ISOKANN_FEAT_IDENTITY;
    for (int64_t m0 = 0; m0 < M; m0 += ch) {
      const int64_t m = std::min(ch, M - m0);
      IK_CUDA(cudaMemcpyAsync(c->staging_in.p, coords + m0 * D, (size_t)m * D * sizeof(float), cudaMemcpyHostToDevice,
                              c->stream));
      launch_featurize(*c, c->staging_in.p, nullptr, 0, m, pairs, false, c->staging_out.p, c->F);
      IK_CUDA(cudaMemcpyAsync(features_out + m0 * c->F, c->staging_out.p, (size_t)m * c->F * sizeof(float),
                              cudaMemcpyDeviceToHost, c->stream));
      sync_stream(*c);
    }
  });
}

int32_t isokann_forward(isokann_ctx *c, const float *in, int64_t rows, int64_t M, int32_t is_features,
                        float *chi_out) {
  return guarded(c, [&] {
    IK_REQUIRE(in && chi_out, ISOKANN_BAD_ARGUMENT, "NULL buffer");
    const bool coords = !is_features;
    IK_REQUIRE(rows == (coords ? c->D : c->F), ISOKANN_BAD_ARGUMENT, "row count does not match the model input");
    const int64_t ch = chunk_rows(*c, 1);
    c->staging_in.ensure((size_t)std::min(ch, std::max<int64_t>(M, 1)) * rows);
    for (int64_t m0 = 0; m0 < M; m0 += ch) {
      const int64_t m = std::min(ch, M - m0);
      IK_CUDA(cudaMemcpyAsync(c->staging_in.p, in + m0 * rows, (size_t)m * rows * sizeof(float),
                              cudaMemcpyHostToDevice, c->stream));
      forward_rows(*c, c->staging_in.p, nullptr, 0, m, coords);
      IK_CUDA(cudaMemcpyAsync(chi_out + m0 * c->d, c->act[c->L].p, (size_t)m * c->d * sizeof(float),
                              cudaMemcpyDeviceToHost, c->stream));
      sync_stream(*c);
    }
    c->timer.flush(c->stream);
  });
}

int32_t isokann_chi_vjp(isokann_ctx *c, const float *in, int64_t rows, int64_t M, int32_t is_features,
                        const float *cot, float *grad_out) {
  return guarded(c, [&] {
    IK_REQUIRE(in && grad_out, ISOKANN_BAD_ARGUMENT, "NULL buffer");
    const bool coords = !is_features;
    IK_REQUIRE(rows == (coords ? c->D : c->F), ISOKANN_BAD_ARGUMENT, "row count does not match the model input");
    const bool pairs = coords && c->cfg.featurizer != ISOKANN_FEAT_IDENTITY;
    const int L = c->L, d = c->d, F = c->F;
    const int64_t ch = std::min<int64_t>(chunk_rows(*c, 1), 16384);
    const int64_t mc = std::min(ch, std::max<int64_t>(M, 1));
    c->staging_in.ensure((size_t)mc * rows);
    c->staging_out.ensure((size_t)mc * std::max<int64_t>(rows, F));
    c->delta_a.ensure((size_t)mc * std::max(c->maxw, d));
    c->delta_b.ensure((size_t)mc * std::max(c->maxw, d));
    DevBuf<float> &cotbuf = c->red_f;
    cotbuf.ensure((size_t)std::max<int64_t>(2048, mc * d));
    for (int64_t m0 = 0; m0 < M; m0 += ch) {
      const int64_t m = std::min(ch, M - m0);
      IK_CUDA(cudaMemcpyAsync(c->staging_in.p, in + m0 * rows, (size_t)m * rows * sizeof(float),
                              cudaMemcpyHostToDevice, c->stream));
      if (cot)
        IK_CUDA(cudaMemcpyAsync(cotbuf.p, cot + m0 * d, (size_t)m * d * sizeof(float), cudaMemcpyHostToDevice,
                                c->stream));
      forward_rows(*c, c->staging_in.p, nullptr, 0, m, coords, true, true);
      float *cur = c->delta_a.p, *other = c->delta_b.p;
      launch_vjp_seed(*c, c->act[L].p, cot ? cotbuf.p : nullptr, m, d, c->cfg.last_activation, cur);
      for (int l = L - 1; l >= 1; --l) {  // delta_{l-1} = (delta_l * W_l^T) .* act'(z_{l-1})
        const int fin = c->cfg.widths[l], fout = c->cfg.widths[l + 1];
        GemmP g{};
        g.A = cur; g.lda = fout;
        g.B = c->params.p + c->off_w[l]; g.ldb = fout;
        g.C = other; g.ldc = fin;
        g.Z = c->act[l].p; g.ldz = fin;
        g.M = (int)m; g.N = fin; g.K = fout;
        g.ones_k = -1; g.ones_i = -1;
        g.act = c->cfg.activation; g.epi = EPI_MULDACT;
        launch_gemm(*c, g, true, false, 1);
        std::swap(cur, other);
      }
      {  // dL/dx_hat = delta_1 * W1'^T (folded first layer), no activation in front of it
        const int fout = c->cfg.widths[1];
        GemmP g{};
        g.A = cur; g.lda = fout;
        g.B = layer_segment(*c, 0); g.ldb = fout;
        g.C = other; g.ldc = F;
        g.M = (int)m; g.N = F; g.K = fout;
        g.ones_k = -1; g.ones_i = -1;
        g.act = ISOKANN_ACT_IDENTITY; g.epi = EPI_ACT;
        launch_gemm(*c, g, true, false, 1);
      }
      launch_featurize_backward(*c, c->staging_in.p, m, pairs, c->ln, other, c->staging_out.p);
      IK_CUDA(cudaMemcpyAsync(grad_out + m0 * rows, c->staging_out.p, (size_t)m * rows * sizeof(float),
                              cudaMemcpyDeviceToHost, c->stream));
      sync_stream(*c);
    }
    c->timer.flush(c->stream);
  });
}

int32_t isokann_chis(isokann_ctx *c, float *chi_out) {
  return guarded(c, [&] {
    compute_chis(*c);
    if (chi_out)
      IK_CUDA(cudaMemcpyAsync(chi_out, c->chi_x.p, (size_t)c->N * c->d * sizeof(float), cudaMemcpyDeviceToHost,
                              c->stream));
    sync_stream(*c);
    c->timer.flush(c->stream);
  });
}

int32_t isokann_koopman(isokann_ctx *c, float *kchi_out) {
  return guarded(c, [&] {
    compute_koopman(*c);
    if (kchi_out)
      IK_CUDA(cudaMemcpyAsync(kchi_out, c->kchi.p, (size_t)c->N * c->d * sizeof(float), cudaMemcpyDeviceToHost,
                              c->stream));
    sync_stream(*c);
    c->timer.flush(c->stream);
  });
}

int32_t isokann_target(isokann_ctx *c, int32_t transform, const isokann_target_opts *opts, float *target_out) {
  return guarded(c, [&] {
    compute_target(*c, transform, opts);
    if (target_out)
      IK_CUDA(cudaMemcpyAsync(target_out, c->target.p, (size_t)c->N * c->d * sizeof(float), cudaMemcpyDeviceToHost,
                              c->stream));
    sync_stream(*c);
    c->timer.flush(c->stream);
  });
}

int32_t isokann_download_target(isokann_ctx *c, float *target_out) {
  return guarded(c, [&] {
    IK_REQUIRE(c->has_target && target_out, ISOKANN_ERR_STATE, "no resident target / NULL output");
    IK_CUDA(cudaMemcpyAsync(target_out, c->target.p, (size_t)c->N * c->d * sizeof(float), cudaMemcpyDeviceToHost,
                            c->stream));
    sync_stream(*c);
  });
}

// validationloss(iso, valdata) (src/iso.jl:160-168) with only scalars leaving the device:
//   c = model(vx); k1 = expectation(model, vy); k2 = expectation(model, ys) on the resident data;
//   SKc = shiftscale([k1; k2])[1:length(c)];  mean(abs2, c - SKc)
int32_t isokann_validationloss(isokann_ctx *c, const float *vxs, const float *vys, int64_t D, int64_t K, int64_t Nv,
                               double *loss_out) {
  return guarded(c, [&] {
    IK_REQUIRE(vxs && vys && loss_out && Nv > 0 && K > 0, ISOKANN_BAD_ARGUMENT, "NULL buffer or empty validation set");
    IK_REQUIRE(D == c->D, ISOKANN_BAD_ARGUMENT, "coordinate dimension does not match the featurizer/model");
    IK_REQUIRE(c->d == 1, ISOKANN_BAD_ARGUMENT, "validationloss shift-scales chi: one dimensional chi functions only");
    Ctx &x = *c;
    compute_koopman(x);  // k2 -> x.kchi (N)
    x.val_chi.ensure((size_t)Nv);
    x.val_k1.ensure((size_t)Nv);
    const int64_t ch = chunk_rows(x, 1);
    // chi on the validation start points
    for (int64_t m0 = 0; m0 < Nv; m0 += ch) {
      const int64_t m = std::min(ch, Nv - m0);
      x.staging_in.ensure((size_t)m * D);
      IK_CUDA(cudaMemcpyAsync(x.staging_in.p, vxs + m0 * D, (size_t)m * D * sizeof(float), cudaMemcpyHostToDevice,
                              x.stream));
      forward_rows(x, x.staging_in.p, nullptr, 0, m, true);
      IK_CUDA(cudaMemcpyAsync(x.val_chi.p + m0, x.act[x.L].p, (size_t)m * sizeof(float), cudaMemcpyDeviceToDevice,
                              x.stream));
      sync_stream(x);  // the staging buffer is reused
    }
    // Koopman expectation on the validation samples, whole start points per chunk
    const int64_t nsp = std::max<int64_t>(1, chunk_rows(x, K) / K);
    for (int64_t n0 = 0; n0 < Nv; n0 += nsp) {
      const int64_t ns = std::min(nsp, Nv - n0);
      x.staging_in.ensure((size_t)ns * K * D);
      IK_CUDA(cudaMemcpyAsync(x.staging_in.p, vys + n0 * K * D, (size_t)ns * K * D * sizeof(float),
                              cudaMemcpyHostToDevice, x.stream));
      forward_rows(x, x.staging_in.p, nullptr, 0, ns * K, true);
      launch_kmean(x, x.act[x.L].p, nullptr, ns, (int)K, 1, x.val_k1.p + n0);
      sync_stream(x);
    }
    // joint extrema of [k1; k2]
    int nb1 = 0, nb2 = 0;
    x.red_f.ensure(4 * 1024);
    launch_minmax(x, x.val_k1.p, Nv, x.red_f.p, &nb1);
    launch_minmax(x, x.kchi.p, x.N, x.red_f.p + 2 * nb1, &nb2);
    float *mm = read_back(x, x.red_f.p, (size_t)2 * (nb1 + nb2));
    float mn = INFINITY, mx = -INFINITY;
    bool nan = false;
    for (int b = 0; b < nb1 + nb2; ++b) {
      nan = nan || mm[2 * b] != mm[2 * b] || mm[2 * b + 1] != mm[2 * b + 1];
      mn = std::min(mn, mm[2 * b]);
      mx = std::max(mx, mm[2 * b + 1]);
    }
    if (nan || !(mx > mn))
      throw ik::Error{ISOKANN_DOMAIN_CONSTANT_CHI, "Could not compute the shift-scale. chi function is constant"};
    int nb = 0;
    launch_valloss(x, x.val_chi.p, x.val_k1.p, Nv, mn, mx, x.red_d.p, &nb);
    double *part = read_back(x, x.red_d.p, (size_t)nb);
    double s = 0.0;
    for (int b = 0; b < nb; ++b) s += part[b];
    *loss_out = s / (double)Nv;
    x.timer.flush(x.stream);
  });
}

int32_t isokann_rates(isokann_ctx *c, double *q_colmajor, int32_t *dim_out) {
  return guarded(c, [&] {
    IK_REQUIRE(q_colmajor != nullptr, ISOKANN_BAD_ARGUMENT, "NULL output");
    Ctx &x = *c;
    Moments mo;
    diag_moments(x, mo);
    double L[kMaxD * kMaxD];
    int n = 0;
    const int rc = diag_rates(mo.uu, mo.vu, x.d, L, &n);
    IK_REQUIRE(rc != 1, ISOKANN_DOMAIN_PINV, "rates: chi chi' is singular (collapsed chi)");
    IK_REQUIRE(rc == 0, ISOKANN_BAD_ARGUMENT,
               "rates: Kchi/chi has no real logarithm (an eigenvalue on the closed negative real axis)");
    for (int a = 0; a < n; ++a)
      for (int b = 0; b < n; ++b) q_colmajor[a + b * n] = L[a * n + b];
    if (dim_out) *dim_out = n;
  });
}

int32_t isokann_residual_subspace(isokann_ctx *c, int32_t v_norms, double *relres_out, double *res_out) {
  return guarded(c, [&] {
    IK_REQUIRE(relres_out != nullptr, ISOKANN_BAD_ARGUMENT, "NULL output");
    Ctx &x = *c;
    Moments mo;
    diag_moments(x, mo);
    const int d = x.d, e = d + 1;
    Mat8 A{}, B{};
    IK_REQUIRE(diag_subspace(mo.uu, mo.vu, d, A, B), ISOKANN_DOMAIN_PINV,
               "residual_subspace: chi has linearly dependent rows");
    double *dev_out = nullptr;
    if (res_out) {
      x.diag_out.ensure((size_t)x.N * d);
      dev_out = x.diag_out.p;
    }
    double r2[kMaxD], k2[kMaxD];
    diag_resid_pass(x, A, B, dev_out, r2, k2);
    for (int j = 0; j < d; ++j) {
      const double den = v_norms ? mo.uu[j * e + j] : mo.vv[j * e + j];
      relres_out[j] = std::sqrt(r2[j]) / std::sqrt(den);
    }
    if (res_out) {
      IK_CUDA(cudaMemcpyAsync(res_out, dev_out, (size_t)x.N * d * sizeof(double), cudaMemcpyDeviceToHost, x.stream));
      sync_stream(x);
    }
  });
}

int32_t isokann_residual_ritz(isokann_ctx *c, double *vals_out, double *vecs_out, double *relres_out,
                              double *residues_out) {
  return guarded(c, [&] {
    IK_REQUIRE(vals_out != nullptr && relres_out != nullptr, ISOKANN_BAD_ARGUMENT, "NULL output");
    Ctx &x = *c;
    Moments mo;
    diag_moments(x, mo);
    const int d = x.d;
    bool any_complex = false;
    Mat8 Are{}, Bre{}, Aim{}, Bim{};
    IK_REQUIRE(diag_ritz(mo.uu, mo.vu, d, vals_out, vecs_out, Are, Bre, Aim, Bim, &any_complex), ISOKANN_DOMAIN_PINV,
               "residual_ritz: chi has linearly dependent rows or eigen(Kr) failed");
    double *dev_re = nullptr, *dev_im = nullptr;
    if (residues_out) {
      x.diag_out.ensure((size_t)x.N * d * 2);
      dev_re = x.diag_out.p;
      dev_im = x.diag_out.p + (size_t)x.N * d;
    }
    double r2[kMaxD], k2[kMaxD], r2i[kMaxD], k2i[kMaxD];
    diag_resid_pass(x, Are, Bre, dev_re, r2, k2);
    for (int j = 0; j < d; ++j) r2i[j] = k2i[j] = 0.0;
    if (any_complex) diag_resid_pass(x, Aim, Bim, dev_im, r2i, k2i);
    for (int j = 0; j < d; ++j) relres_out[j] = std::sqrt(r2[j] + r2i[j]) / std::sqrt(k2[j] + k2i[j]);
    if (residues_out) {  // interleave (re, im) like a Julia Matrix{ComplexF64}, in pieces through the pinned buffer
      const size_t total = (size_t)x.N * d, piece = (size_t)1 << 20;
      for (size_t o = 0; o < total; o += piece) {
        const size_t m = std::min(piece, total - o);
        ensure_pinned(x, 2 * m * sizeof(double));
        double *h = reinterpret_cast<double *>(x.pinned);
        IK_CUDA(cudaMemcpyAsync(h, dev_re + o, m * sizeof(double), cudaMemcpyDeviceToHost, x.stream));
        if (any_complex)
          IK_CUDA(cudaMemcpyAsync(h + m, dev_im + o, m * sizeof(double), cudaMemcpyDeviceToHost, x.stream));
        sync_stream(x);
        for (size_t i = 0; i < m; ++i) {
          residues_out[2 * (o + i)] = h[i];
          residues_out[2 * (o + i) + 1] = any_complex ? h[m + i] : 0.0;
        }
      }
    }
  });
}

// Julia's randperm(rng::Xoshiro, n) replayed on the host (Random stdlib, Julia 1.12): Xoshiro256++ draws, the 52-bit
// raw sample `rand(UInt64) >>> 12`, and randperm!'s inside-out shuffle with ltm52's masked rejection sampling
// ("minibatch order").  UNPINNED: no golden vector from a Julia session is available in this environment;
// tests compare it with an independent Python restatement of the same published algorithm.
int32_t isokann_randperm(uint64_t *state4, int64_t n, int64_t *perm_out) {
  if (!state4 || n < 0 || (n > 0 && !perm_out)) return ISOKANN_BAD_ARGUMENT;
  uint64_t s0 = state4[0], s1 = state4[1], s2 = state4[2], s3 = state4[3];
  auto rotl = [](uint64_t x, int k) { return (x << k) | (x >> (64 - k)); };
  auto next = [&]() {
    const uint64_t res = rotl(s0 + s3, 23) + s0;
    const uint64_t t = s1 << 17;
    s2 ^= s0;
    s3 ^= s1;
    s1 ^= s2;
    s0 ^= s3;
    s2 ^= t;
    s3 = rotl(s3, 45);
    return res;
  };
  if (n > 0) perm_out[0] = 1;
  uint64_t mask = 3;
  for (int64_t i = 2; i <= n; ++i) {
    uint64_t x;
    do {
      x = (next() >> 12) & mask;           // ltm52(i, mask): masked 52-bit raw draw, rejected until <= i - 1
    } while (x > (uint64_t)(i - 1));
    const int64_t j = 1 + (int64_t)x;
    if (i != j) perm_out[i - 1] = perm_out[j - 1];
    perm_out[j - 1] = i;
    if ((uint64_t)i == 1 + mask) mask = 2 * mask + 1;
  }
  state4[0] = s0; state4[1] = s1; state4[2] = s2; state4[3] = s3;
  return ISOKANN_OK;
}

int32_t isokann_set_target(isokann_ctx *c, const float *target, int64_t d, int64_t N) {
  return guarded(c, [&] {
    IK_REQUIRE(target && d == c->d && N == c->N, ISOKANN_BAD_ARGUMENT, "target shape mismatch");
    c->target.ensure((size_t)N * d);
    IK_CUDA(cudaMemcpyAsync(c->target.p, target, (size_t)N * d * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    if (d == 1) {
      set_unit_weights(*c);
    } else {
      IK_REQUIRE(d <= kMaxD, ISOKANN_BAD_ARGUMENT, "chi dimension exceeds ISOKANN_MAX_D");
      Mat8 id{};
      for (int a = 0; a < d; ++a) id.m[a * d + a] = 1.0;
      // identity apply in place: recomputes the loss weights 1 ./ std(target, dims=2)
      c->kchi.ensure((size_t)N * d);
      IK_CUDA(cudaMemcpyAsync(c->kchi.p, c->target.p, (size_t)N * d * sizeof(float), cudaMemcpyDeviceToDevice,
                              c->stream));
      finish_nd_target(*c, id, false, false);
    }
    sync_stream(*c);
    c->has_target = true;
  });
}

int32_t isokann_train_epoch(isokann_ctx *c, const int64_t *perm, int64_t minibatch, int32_t partial,
                            double *loss_out) {
  return guarded(c, [&] {
    const double l = train_epoch(*c, perm, minibatch, partial != 0);
    if (loss_out) *loss_out = l;
    c->timer.flush(c->stream);
  });
}

int32_t isokann_iterate(isokann_ctx *c, int32_t transform, const isokann_target_opts *opts, int64_t n_iter,
                        int64_t epochs, int64_t minibatch, const int64_t *perms, double *losses_out) {
  return guarded(c, [&] {
    IK_REQUIRE(perms != nullptr, ISOKANN_BAD_ARGUMENT, "perms must not be NULL");
    int64_t k = 0;
    const int64_t total = n_iter * epochs;
    for (int64_t it = 0; it < n_iter; ++it) {
      IK_REQUIRE(c->xs != nullptr, ISOKANN_ERR_STATE, "no data: call isokann_set_data first");
      if (k == 0 && total > 0) preload_perm(*c, perms);
      compute_target(*c, transform, opts);
      for (int64_t e = 0; e < epochs; ++e, ++k) {
        const double l = train_epoch(*c, perms + k * c->N, minibatch, false,
                                     k + 1 < total ? perms + (k + 1) * c->N : nullptr);
        if (losses_out) losses_out[k] = l;
      }
    }
    c->timer.flush(c->stream);
  });
}

int32_t isokann_download_grads(isokann_ctx *c, float *flat, int64_t P) {
  return guarded(c, [&] {
    IK_REQUIRE(flat && P == c->P, ISOKANN_BAD_ARGUMENT, "parameter count mismatch");
    IK_CUDA(cudaMemcpyAsync(flat, c->grads.p, (size_t)P * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    sync_stream(*c);
  });
}

int32_t isokann_target_matrices(isokann_ctx *c, float *kinv_colmajor, float *schur_colmajor, double *applied_rowmajor) {
  return guarded(c, [&] {
    IK_REQUIRE(c->d > 1 && c->has_target, ISOKANN_ERR_STATE, "no N-D target has been computed");
    const int n = c->d * c->d;
    for (int i = 0; i < n; ++i) {
      if (kinv_colmajor) kinv_colmajor[i] = c->last_kinv[i];
      if (schur_colmajor) schur_colmajor[i] = c->last_schur[i];
      if (applied_rowmajor) applied_rowmajor[i] = c->last_mat[i];
    }
  });
}

int32_t isokann_enable_timing(isokann_ctx *c, int32_t on) {
  return guarded(c, [&] {
    c->timer.flush(c->stream);
    c->timer.enabled = on != 0;
    c->timer.phases_only = on == 2;  // 2: one event pair per phase only -- cheap enough for the timed pass
  });
}

int32_t isokann_get_stats(isokann_ctx *c, isokann_stats *out) {
  return guarded(c, [&] {
    IK_REQUIRE(out != nullptr, ISOKANN_BAD_ARGUMENT, "NULL stats");
    c->timer.flush(c->stream);
    c->stats.ms_featurize = c->timer.ms[KC_FEATURIZE];
    c->stats.ms_gemm = c->timer.ms[KC_GEMM];
    c->stats.ms_reduce = c->timer.ms[KC_REDUCE];
    c->stats.ms_train_elementwise = c->timer.ms[KC_TRAIN_EW];
    c->stats.ms_optimiser = c->timer.ms[KC_OPT];
    c->stats.ms_koopman_total = c->timer.ms[KC_PHASE_KOOPMAN];
    c->stats.ms_target_total = c->timer.ms[KC_PHASE_TARGET];
    c->stats.ms_train_total = c->timer.ms[KC_PHASE_TRAIN];
    c->stats.ms_nccl = c->timer.ms[KC_NCCL];
    *out = c->stats;
  });
}

int32_t isokann_reset_stats(isokann_ctx *c) {
  return guarded(c, [&] {
    c->timer.flush(c->stream);
    for (double &m : c->timer.ms) m = 0.0;
    c->stats = isokann_stats{};
  });
}

int32_t isokann_synchronize(isokann_ctx *c) {
  return guarded(c, [&] {
    if (c->copy_stream) IK_CUDA(cudaStreamSynchronize(c->copy_stream));
    c->xs_pending = false;
    sync_stream(*c);
  });
}

void *isokann_stream(isokann_ctx *c) { return c ? (void *)c->stream : nullptr; }

int32_t isokann_release_host_buffers(isokann_ctx *c) {
  return guarded(c, [&] {
    if (c->copy_stream) IK_CUDA(cudaStreamSynchronize(c->copy_stream));
    c->xs_pending = false;
    sync_stream(*c);
    release_host_registrations(*c);
  });
}

int32_t isokann_host_diag(int32_t what, const double *uu, const double *vu, int32_t d, double *out) {
  if (!uu || !vu || !out || d < 1 || d > ik::kMaxD) return ISOKANN_BAD_ARGUMENT;
  const int dd = d * d;
  if (what == 0) {
    int n = 0;
    const int rc = ik::diag_rates(uu, vu, d, out + 1, &n);
    out[0] = (double)n;
    return rc == 0 ? ISOKANN_OK : (rc == 1 ? ISOKANN_DOMAIN_PINV : ISOKANN_BAD_ARGUMENT);
  }
  if (what == 1) {
    ik::Mat8 A{}, B{};
    if (!ik::diag_subspace(uu, vu, d, A, B)) return ISOKANN_DOMAIN_PINV;
    for (int i = 0; i < dd; ++i) {
      out[i] = A.m[i];
      out[dd + i] = B.m[i];
    }
    return ISOKANN_OK;
  }
  if (what == 2) {
    ik::Mat8 Are{}, Bre{}, Aim{}, Bim{};
    bool any_complex = false;
    if (!ik::diag_ritz(uu, vu, d, out, out + 2 * d, Are, Bre, Aim, Bim, &any_complex)) return ISOKANN_DOMAIN_PINV;
    double *q = out + 2 * d + 2 * dd;
    for (int i = 0; i < dd; ++i) {
      q[i] = Are.m[i];
      q[dd + i] = Bre.m[i];
      q[2 * dd + i] = Aim.m[i];
      q[3 * dd + i] = Bim.m[i];
    }
    q[4 * dd] = any_complex ? 1.0 : 0.0;
    return ISOKANN_OK;
  }
  return ISOKANN_BAD_ARGUMENT;
}

int32_t isokann_host_logm(const double *a_colmajor, int32_t n, double *out_colmajor) {
  if (!a_colmajor || !out_colmajor || n < 1 || n > ik::kMaxD + 1) return ISOKANN_BAD_ARGUMENT;
  double a[(ik::kMaxD + 1) * (ik::kMaxD + 1)], l[(ik::kMaxD + 1) * (ik::kMaxD + 1)];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) a[i * n + j] = a_colmajor[i + j * n];
  if (!ik::host_logm(a, n, l)) return ISOKANN_BAD_ARGUMENT;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) out_colmajor[i + j * n] = l[i * n + j];
  return ISOKANN_OK;
}

int32_t isokann_host_eig(const double *a_colmajor, int32_t n, double *vals_reim, double *vecs_reim_colmajor) {
  if (!a_colmajor || !vals_reim || !vecs_reim_colmajor || n < 1 || n > ik::kMaxD) return ISOKANN_BAD_ARGUMENT;
  double a[ik::kMaxD * ik::kMaxD], wr[ik::kMaxD], wi[ik::kMaxD], vre[ik::kMaxD * ik::kMaxD], vim[ik::kMaxD * ik::kMaxD];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) a[i * n + j] = a_colmajor[i + j * n];
  if (!ik::host_eig_general(a, n, wr, wi, vre, vim)) return ISOKANN_BAD_ARGUMENT;
  for (int j = 0; j < n; ++j) {
    vals_reim[2 * j] = wr[j];
    vals_reim[2 * j + 1] = wi[j];
    for (int i = 0; i < n; ++i) {
      vecs_reim_colmajor[2 * (i + j * n)] = vre[i * n + j];
      vecs_reim_colmajor[2 * (i + j * n) + 1] = vim[i * n + j];
    }
  }
  return ISOKANN_OK;
}

int32_t isokann_host_schur(const float *a_colmajor, int32_t d, float *z_colmajor, float *t_colmajor) {
  if (!a_colmajor || !z_colmajor || d < 1 || d > kMaxD) return ISOKANN_BAD_ARGUMENT;
  return host_schur_f32(a_colmajor, d, z_colmajor, t_colmajor) ? ISOKANN_OK : ISOKANN_DOMAIN_PINV;
}

}  // extern "C"
