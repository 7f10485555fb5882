// Element-wise kernels of the training step (reference src/iso.jl:179-194, src/models.jl:4-20):
// loss + output delta, LayerNorm-affine folding / unfolding, and the fused
// WeightDecay -> {Adam | Nesterov} -> subtract update over the flat parameter vector
// (Optimisers.jl 0.4.7 semantics; one pass instead of ~4 broadcast kernels per parameter array).
#include "common.cuh"

namespace ik {

__device__ __forceinline__ float dact_out(float z, int kind) {
  switch (kind) {
    case ISOKANN_ACT_SIGMOID: return z * (1.0f - z);
    case ISOKANN_ACT_TANH: return 1.0f - z * z;
    case ISOKANN_ACT_RELU: return z > 0.f ? 1.0f : 0.f;
    default: return 1.0f;
  }
}

__device__ __forceinline__ double warp_sum_dd(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// l = sum(abs2, (m(x) .- y) .* w); delta = d(l/B)/d(chi) (times the last activation's derivative).
// d == 1 follows the reference's Float64 promotion (w = 1.0, src/iso.jl:183).  The last block to
// finish adds the per-block partials in block order and packs the step loss as (hi, lo) floats
// behind the gradient vector, so it rides the gradient all-reduce.
__global__ void __launch_bounds__(256) loss_delta_kernel(const float *__restrict__ chi, const float *__restrict__ target,
                                                         const int64_t *__restrict__ idx, const float *__restrict__ w,
                                                         int64_t Bloc, int d, double Bglobal, int lastact,
                                                         float *__restrict__ delta, double *__restrict__ partials,
                                                         unsigned int *__restrict__ ticket,
                                                         float *__restrict__ packed_tail) {
  __shared__ double sh[8];
  __shared__ bool is_last;
  double l = 0.0;
  const int64_t total = Bloc * d;
  const float invB = (float)(1.0 / Bglobal);
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = t / d;
    const int a = (int)(t - m * d);
    const float c = chi[t];
    const float y = target[idx[m] * d + a];
    const float r = c - y;
    float dl;
    if (d == 1) {
      l += (double)r * (double)r;
      dl = (float)(2.0 * (double)r / Bglobal);
    } else {
      const float wa = w[a];
      const float z = r * wa;
      l += (double)(z * z);
      dl = ((2.0f * z) * invB) * wa;
    }
    delta[t] = dl * dact_out(c, lastact);
  }
  l = warp_sum_dd(l);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = l;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) s += sh[k];
    partials[blockIdx.x] = s;
    __threadfence();
    const unsigned int prev = atomicAdd(ticket, 1u);
    is_last = (prev == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    __threadfence();
    double s = 0.0;
    for (unsigned int b = 0; b < gridDim.x; ++b) s += ((volatile double *)partials)[b];
    const float hi = (float)s;
    const float lo = (hi == hi && fabsf(hi) != INFINITY) ? (float)(s - (double)hi) : 0.f;
    packed_tail[0] = hi;
    packed_tail[1] = lo;
    *ticket = 0u;
  }
}

void launch_loss_delta(Ctx &c, const float *chi, const float *target, const int64_t *idx, int64_t idx_off,
                       const float *w, int64_t Bloc, int d, double Bglobal, int lastact, float *delta,
                       double *partials, unsigned int *ticket, float *packed_tail) {
  int64_t total = Bloc * d;
  int grid = (int)((total + 255) / 256);
  if (grid > 128) grid = 128;
  if (grid < 1) grid = 1;
  c.timer.begin(KC_TRAIN_EW, c.stream);
  loss_delta_kernel<<<grid, 256, 0, c.stream>>>(chi, target, idx + idx_off, w, Bloc, d, Bglobal, lastact, delta,
                                                partials, ticket, packed_tail);
  c.timer.end(c.stream);
  IK_CUDA(cudaGetLastError());
  c.count_launch(KC_TRAIN_EW);
}

// seed of a vector-Jacobian product: delta = cot .* act'(chi)   (cot == nullptr: ones)
__global__ void vjp_seed_kernel(const float *__restrict__ chi, const float *__restrict__ cot, int64_t total,
                                int lastact, float *__restrict__ delta) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x)
    delta[t] = (cot ? cot[t] : 1.0f) * dact_out(chi[t], lastact);
}

void launch_vjp_seed(Ctx &c, const float *chi, const float *cot, int64_t M, int d, int lastact, float *delta) {
  const int64_t total = M * d;
  if (total <= 0) return;
  int grid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)c.num_sms * 8);
  vjp_seed_kernel<<<grid, 256, 0, c.stream>>>(chi, cot, total, lastact, delta);
  IK_CUDA(cudaGetLastError());
  c.count_launch(KC_TRAIN_EW);
}

// LayerNorm affine folded into the first Dense layer: z0 = gamma .* xhat .+ beta, so
//   W1 z0 + b1 = (W1 diag(gamma)) xhat + (b1 + W1 beta).
// folded is the row-major (F+1) x h1 segment [W1' ; b1'] consumed by the GEMM.
__global__ void fold_ln_kernel(const float *__restrict__ gamma, const float *__restrict__ beta,
                               const float *__restrict__ W1, const float *__restrict__ b1, int F, int h1,
                               float *__restrict__ folded) {
  const int64_t total = (int64_t)F * h1;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x)
    folded[t] = gamma[t / h1] * W1[t];
  // bias row: one warp per output column j, lanes stride over the input features
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int j = blockIdx.x * wpb + (threadIdx.x >> 5); j < h1; j += gridDim.x * wpb) {
    float s = 0.f;
    for (int g = lane; g < F; g += 32) s = fmaf(beta[g], W1[(int64_t)g * h1 + j], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) folded[total + j] = b1[j] + s;
  }
}

void launch_fold_ln(Ctx &c, const float *gamma, const float *beta, const float *W1, const float *b1, int F, int h1,
                    float *folded) {
  int64_t total = (int64_t)F * h1;
  int grid = (int)((total + 255) / 256);
  if (grid > c.num_sms * 8) grid = c.num_sms * 8;
  c.timer.begin(KC_TRAIN_EW, c.stream);
  fold_ln_kernel<<<grid, 256, 0, c.stream>>>(gamma, beta, W1, b1, F, h1, folded);
  c.timer.end(c.stream);
  IK_CUDA(cudaGetLastError());
  c.count_launch(KC_TRAIN_EW);
}

// Chain rule of the folding.  gfold = [G ; gb] is the gradient w.r.t. the folded segment:
//   dW1[f,j] = gamma[f] G[f,j] + beta[f] gb[j],  db1 = gb,
//   dgamma[f] = sum_j W1[f,j] G[f,j],            dbeta[f] = sum_j W1[f,j] gb[j].
// One warp per input feature f.
__global__ void unfold_ln_kernel(const float *__restrict__ gamma, const float *__restrict__ beta,
                                 const float *__restrict__ W1, const float *__restrict__ gfold, int F, int h1,
                                 float *__restrict__ g_gamma, float *__restrict__ g_beta, float *__restrict__ g_W1,
                                 float *__restrict__ g_b1) {
  // one BLOCK per input feature f (F is a few hundred: one warp per feature left most of the GPU idle, 29 us for
  // the c5 layer); the warps split the h1 columns with 16-byte accesses and their partial sums are added in warp order
  __shared__ float sg_s[8], sb_s[8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float *gb = gfold + (int64_t)F * h1;
  const bool vec = (h1 & 3) == 0 && ((reinterpret_cast<uintptr_t>(W1) | reinterpret_cast<uintptr_t>(gfold) |
                                      reinterpret_cast<uintptr_t>(g_W1)) & 15) == 0;
  for (int f = blockIdx.x; f < F; f += gridDim.x) {
    const float ga = gamma[f], be = beta[f];
    const float *wr = W1 + (int64_t)f * h1, *gr = gfold + (int64_t)f * h1;
    float *orow = g_W1 + (int64_t)f * h1;
    float sg = 0.f, sb = 0.f;
    if (vec) {
      for (int j = 4 * threadIdx.x; j < h1; j += 4 * blockDim.x) {
        const float4 wv = *reinterpret_cast<const float4 *>(wr + j), G = *reinterpret_cast<const float4 *>(gr + j);
        const float4 b = *reinterpret_cast<const float4 *>(gb + j);
        *reinterpret_cast<float4 *>(orow + j) =
            make_float4(fmaf(ga, G.x, be * b.x), fmaf(ga, G.y, be * b.y), fmaf(ga, G.z, be * b.z), fmaf(ga, G.w, be * b.w));
        sg = fmaf(wv.x, G.x, sg); sg = fmaf(wv.y, G.y, sg); sg = fmaf(wv.z, G.z, sg); sg = fmaf(wv.w, G.w, sg);
        sb = fmaf(wv.x, b.x, sb); sb = fmaf(wv.y, b.y, sb); sb = fmaf(wv.z, b.z, sb); sb = fmaf(wv.w, b.w, sb);
      }
    } else {
      for (int j = threadIdx.x; j < h1; j += blockDim.x) {
        const float wv = wr[j], G = gr[j], b = gb[j];
        orow[j] = fmaf(ga, G, be * b);
        sg = fmaf(wv, G, sg);
        sb = fmaf(wv, b, sb);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sg += __shfl_xor_sync(0xffffffffu, sg, o);
      sb += __shfl_xor_sync(0xffffffffu, sb, o);
    }
    __syncthreads();  // the previous feature's partials have been consumed
    if (lane == 0) {
      sg_s[w] = sg;
      sb_s[w] = sb;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      float a = 0.f, b = 0.f;
      for (int k = 0; k < nw; ++k) {
        a += sg_s[k];
        b += sb_s[k];
      }
      g_gamma[f] = a;
      g_beta[f] = b;
    }
  }
  if (blockIdx.x == 0)
    for (int j = threadIdx.x; j < h1; j += blockDim.x) g_b1[j] = gb[j];
}

void launch_unfold_ln(Ctx &c, const float *gamma, const float *beta, const float *W1, const float *gfold, int F,
                      int h1, float *g_gamma, float *g_beta, float *g_W1, float *g_b1) {
  // narrow first layers keep one warp per block; wide ones use up to 8 warps on a row
  int threads = 32;
  while (threads < 256 && threads * 4 < h1) threads *= 2;
  int grid = F;
  if (grid > c.num_sms * 8) grid = c.num_sms * 8;
  c.timer.begin(KC_TRAIN_EW, c.stream);
  unfold_ln_kernel<<<grid, threads, 0, c.stream>>>(gamma, beta, W1, gfold, F, h1, g_gamma, g_beta, g_W1, g_b1);
  c.timer.end(c.stream);
  IK_CUDA(cudaGetLastError());
  c.count_launch(KC_TRAIN_EW);
}

// Fused optimiser.  g' = g + lambda*theta (WeightDecay), then
//   Adam     : m = b1 m + (1-b1) g';  v = b2 v + (1-b2) g'^2;
//              theta -= m/(1-bt1) / (sqrt(v/(1-bt2)) + eps) * eta          (bt = running beta^t)
//   Nesterov : dx = -(rho^2) vel + (1+rho) eta g';  vel = rho vel - eta g';  theta -= dx
// The step loss (packed behind the gradient) gates the update: a non-finite loss raises the
// sticky flag and leaves the parameters untouched, as the reference throws before update!
// (src/iso.jl:186-189).  Block 0 accumulates the epoch loss.
struct OptHyper {
  float eta, lambda, beta1, beta2, eps, rho, bt1, bt2;
  int kind;
};

// The update runs over the index range [lo, hi) of the flat vector: the multi-GPU step reduces and updates the
// gradient in two buckets (layers >= 2 first, while the backward pass of layer 1 is still running).
// beta_dev (Adam): running (beta1^t, beta2^t) kept on the device so that a captured step can be replayed.
__global__ void __launch_bounds__(256) optimiser_kernel(float *__restrict__ theta, const float *__restrict__ g,
                                                        float *__restrict__ m, float *__restrict__ v, int64_t P,
                                                        int64_t lo, int64_t hi, OptHyper h,
                                                        const float *__restrict__ beta_dev,
                                                        double *__restrict__ epoch_loss, int *__restrict__ flags) {
  const double loss = (double)g[P] + (double)g[P + 1];
  const bool finite = (loss == loss) && (fabs(loss) != INFINITY);
  const bool poisoned = (*(volatile int *)flags & FLAG_NONFINITE_LOSS) != 0;
  if (!finite || poisoned) {
    if (blockIdx.x == 0 && threadIdx.x == 0 && !finite) atomicOr(flags, FLAG_NONFINITE_LOSS);
    return;
  }
  if (epoch_loss && blockIdx.x == 0 && threadIdx.x == 0) *epoch_loss += loss;
  const float bt1 = beta_dev[0], bt2 = beta_dev[1];
  for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x) {
    const float th = theta[i];
    const float gg = g[i] + h.lambda * th;
    float dx;
    if (h.kind == ISOKANN_OPT_ADAM) {
      const float mt = h.beta1 * m[i] + (1.0f - h.beta1) * gg;
      const float vt = h.beta2 * v[i] + (1.0f - h.beta2) * (gg * gg);
      m[i] = mt;
      v[i] = vt;
      dx = mt / (1.0f - bt1) / (sqrtf(vt / (1.0f - bt2)) + h.eps) * h.eta;
    } else {
      const float vel = m[i];
      dx = -(h.rho * h.rho) * vel + (1.0f + h.rho) * h.eta * gg;
      m[i] = h.rho * vel - h.eta * gg;
    }
    theta[i] = th - dx;
  }
}

// beta^t <- beta^t * beta after a step whose loss was finite (the reference throws before update!, so a rejected
// step leaves the optimiser state untouched); its own launch so that every block of the update saw the old value
__global__ void advance_beta_kernel(float *__restrict__ beta_dev, float beta1, float beta2, const float *__restrict__ g,
                                    int64_t P, const int *__restrict__ flags) {
  const double loss = (double)g[P] + (double)g[P + 1];
  const bool finite = (loss == loss) && (fabs(loss) != INFINITY);
  if (!finite || (*flags & FLAG_NONFINITE_LOSS)) return;
  beta_dev[0] *= beta1;
  beta_dev[1] *= beta2;
}

void launch_optimiser_range(Ctx &c, int64_t lo, int64_t hi, bool accumulate_loss) {
  if (hi <= lo) return;
  OptHyper h{c.cfg.eta, c.cfg.lambda, c.cfg.beta1, c.cfg.beta2, c.cfg.eps, c.cfg.rho, 0.f, 0.f, c.cfg.optimiser};
  int grid = (int)((hi - lo + 255) / 256);
  if (grid > c.num_sms * 8) grid = c.num_sms * 8;
  c.timer.begin(KC_OPT, c.stream);
  optimiser_kernel<<<grid, 256, 0, c.stream>>>(c.params.p, c.grads.p, c.opt_m.p, c.opt_v.p, c.P, lo, hi, h,
                                               c.beta_dev.p, accumulate_loss ? c.epoch_loss.p : nullptr, c.flags.p);
  c.timer.end(c.stream);
  IK_CUDA(cudaGetLastError());
  c.count_launch(KC_OPT);
}

void launch_advance_beta(Ctx &c) {
  if (c.cfg.optimiser != ISOKANN_OPT_ADAM) return;
  advance_beta_kernel<<<1, 1, 0, c.stream>>>(c.beta_dev.p, c.cfg.beta1, c.cfg.beta2, c.grads.p, c.P, c.flags.p);
  IK_CUDA(cudaGetLastError());
  c.count_launch(KC_OPT);
}

void launch_optimiser(Ctx &c, int64_t P) {
  launch_optimiser_range(c, 0, P, true);
  launch_advance_beta(c);
}

}  // namespace ik
