// Fused forward + loss + backward for narrow nets (default pairnets: F -> h1 <= 128 -> <= 16 ... -> d)
// on small minibatches -- the strictly sequential regime of train_batch! (reference
// src/iso.jl:179-194: floor(N/B) dependent steps, default B = 100, scripts use 1000).
//
// The reference pays ~4 broadcast kernels x 8 parameter arrays + 2 gathers + fwd/bwd GEMM launches
// per step; the generic path of this library still needed ~17 launches of latency-bound GEMMs
// (~0.6 ms per step at B = 1000).  Here one kernel does everything between the featurizer and the
// optimiser for a slice of 16 minibatch rows per CTA:
//   x_hat rows -> shared memory (transposed, 16 rows per feature) -> layer 1 (thread = output
//   column, 8 rows in registers, weights streamed coalesced from L2) -> tail layers -> loss and
//   delta -> backward through the tail -> layer-1 weight gradient as 16-row outer products.
// Every CTA writes its gradient partial (folded layout [W1'; b1'] | [W2; b2] | ...) to a private
// slice; narrow_reduce_kernel adds the slices in CTA order, so the result is deterministic.
#include "common.cuh"

namespace ik {

namespace {

constexpr int RMAX = 16;       // minibatch rows per group: 16, or 8 when 16-row groups would leave SMs idle
constexpr int H1P = 128;       // padded first hidden width
constexpr int TW = 16;         // padded tail width
constexpr int NT = 512;        // 128 columns x 4 k-quarters
constexpr int KQ = NT / H1P;

struct NarrowP {
  const float *xhat;
  int64_t ldx;
  int B, F, L;
  int w[ISOKANN_MAX_LAYERS + 1];
  const float *seg[ISOKANN_MAX_LAYERS];
  int64_t off[ISOKANN_MAX_LAYERS + 1];  // offsets of the layers inside one partial
  const float *target;
  const int64_t *idx;
  const float *wloss;
  double Bglobal;
  int act, last_act;
  float *part;
  int64_t part_stride;
  double *part_loss;
  int groups;
};

__device__ __forceinline__ float actf(float a, int kind) {
  switch (kind) {
    case ISOKANN_ACT_SIGMOID: return 1.0f / (1.0f + __expf(-a));
    case ISOKANN_ACT_TANH: return tanhf(a);
    case ISOKANN_ACT_RELU: return fmaxf(a, 0.f);
    default: return a;
  }
}
__device__ __forceinline__ float dactf(float z, int kind) {
  switch (kind) {
    case ISOKANN_ACT_SIGMOID: return z * (1.0f - z);
    case ISOKANN_ACT_TANH: return 1.0f - z * z;
    case ISOKANN_ACT_RELU: return z > 0.f ? 1.0f : 0.f;
    default: return 1.0f;
  }
}

template <int R>
__global__ void __launch_bounds__(NT) narrow_fwd_bwd_kernel(NarrowP p) {
  extern __shared__ __align__(16) float sm[];
  float *xT = sm;                                 // [F][R]
  float *z1 = xT + (size_t)p.F * R;               // [R][H1P]
  float *d1 = z1 + R * H1P;                       // [R][H1P]
  float *zt = d1 + R * H1P;                       // [L-1][R][TW]  activations of layers 2..L
  float *dt = zt + (ISOKANN_MAX_LAYERS)*R * TW;   // [L-1][R][TW]  deltas of layers 2..L
  float *scr = dt + (ISOKANN_MAX_LAYERS)*R * TW;  // [KQ-1][R][H1P] partial sums of the k quarters
  __shared__ double red[NT / 32];
  const int tid = threadIdx.x;
  const int h1 = p.w[1];
  const int L = p.L;
  float *part = p.part + (int64_t)blockIdx.x * p.part_stride;
  double loss_acc = 0.0;
  bool first = true;

  for (int g = blockIdx.x; g < p.groups; g += gridDim.x) {
    const int row0 = g * R;
    __syncthreads();
    // (1) x_hat rows of this group, transposed: xT[k][r]
    for (int e = tid; e < R * p.F; e += NT) {
      const int r = e & (R - 1), k = e / R;
      const int row = row0 + r;
      xT[e] = row < p.B ? __ldg(p.xhat + (int64_t)row * p.ldx + k) : 0.f;
    }
    __syncthreads();
    // (2) layer 1: thread = (column j, k half); 16 rows in registers, the two k halves are combined
    //     through shared memory (d1 is free during the forward pass)
    {
      const int j = tid & (H1P - 1), kh = tid >> 7;  // kh: k quarter
      float acc[R];
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = 0.f;
      if (j < h1) {
        const int kq = (p.F + KQ - 1) / KQ;
        const int k0 = min(p.F, kh * kq), k1 = min(p.F, k0 + kq);
        const float *wcol = p.seg[0] + j;
        const float4 *x4 = reinterpret_cast<const float4 *>(xT);
        // weights are streamed from L2: issue PF loads before any dependent FMA (the first ncu capture showed
        // the compiler serialising load -> use, 8 exposed L2 round trips per unrolled body)
        // Every CTA streams the same weight rows: start each CTA at a different row (rotation) so that
        // they do not all hit the same L2 lines at the same time (second capture: ~5k cycles per batch of 16
        // rows with 63 CTAs in lockstep).
        constexpr int PF = 16;
        const int len = k1 - k0;
        const int rot = len > 0 ? (int)((blockIdx.x * 29u) % (unsigned)len) : 0;
        for (int kb = 0; kb < len; kb += PF) {
          float wv[PF];
          int kk[PF];
#pragma unroll
          for (int u = 0; u < PF; ++u) {
            int k = kb + u + rot;
            if (k >= len) k -= len;
            // slots past the end of this thread's k range multiply a zero weight, but the x they read must still be
            // finite: row k0 of the staged tile (rows beyond F are uninitialised shared memory -- a NaN pattern left
            // there by another kernel turned every gradient of the F = 2 triple-well net into NaN on multi-rank runs)
            const bool live = kb + u < len;
            kk[u] = live ? k0 + k : k0;
            wv[u] = live ? __ldg(wcol + (int64_t)kk[u] * h1) : 0.f;
          }
#pragma unroll
          for (int u = 0; u < PF; ++u) {
            const int k = kk[u];
            const float w = wv[u];
#pragma unroll
            for (int q = 0; q < R / 4; ++q) {
              const float4 a = x4[k * (R / 4) + q];
              acc[4 * q] = fmaf(a.x, w, acc[4 * q]);
              acc[4 * q + 1] = fmaf(a.y, w, acc[4 * q + 1]);
              acc[4 * q + 2] = fmaf(a.z, w, acc[4 * q + 2]);
              acc[4 * q + 3] = fmaf(a.w, w, acc[4 * q + 3]);
            }
          }
        }
        if (kh) {
#pragma unroll
          for (int r = 0; r < R; ++r) scr[((kh - 1) * R + r) * H1P + j] = acc[r];
        }
      }
      __syncthreads();
      if (j < h1 && !kh) {
        const float b = __ldg(p.seg[0] + (int64_t)p.F * h1 + j);
        const int kind = L == 1 ? p.last_act : p.act;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          float a = acc[r];
#pragma unroll
          for (int q = 0; q < KQ - 1; ++q) a += scr[(q * R + r) * H1P + j];
          z1[r * H1P + j] = actf(a + b, kind);
        }
      }
    }
    __syncthreads();
    // (3) tail layers l = 2..L (layer l: w[l-1] -> w[l])
    for (int l = 2; l <= L; ++l) {
      const int win = p.w[l - 1], wout = p.w[l];
      const float *in = l == 2 ? z1 : zt + (l - 3) * R * TW;
      const int ldin = l == 2 ? H1P : TW;
      float *out = zt + (l - 2) * R * TW;
      const float *sg = p.seg[l - 1];
      if (tid < R * wout) {
        const int r = tid / wout, j = tid - r * wout;
        float acc = __ldg(sg + win * wout + j);
        for (int k = 0; k < win; ++k) acc = fmaf(in[r * ldin + k], __ldg(sg + k * wout + j), acc);
        out[r * TW + j] = actf(acc, l == L ? p.last_act : p.act);
      }
      __syncthreads();
    }
    // (4) loss and delta of the last layer
    {
      const int d = p.w[L];
      const float *chi = L == 1 ? z1 : zt + (L - 2) * R * TW;
      const int ldc = L == 1 ? H1P : TW;
      float *dl = L == 1 ? d1 : dt + (L - 2) * R * TW;
      double l = 0.0;
      if (tid < R * d) {
        const int r = tid / d, a = tid - r * d;
        const int row = row0 + r;
        float delta = 0.f;
        if (row < p.B) {
          const float c = chi[r * ldc + a];
          const float y = __ldg(p.target + p.idx[row] * d + a);
          const float res = c - y;
          if (d == 1) {
            l = (double)res * (double)res;
            delta = (float)(2.0 * (double)res / p.Bglobal);
          } else {
            const float wa = __ldg(p.wloss + a);
            const float z = res * wa;
            l = (double)(z * z);
            delta = ((2.0f * z) * (float)(1.0 / p.Bglobal)) * wa;
          }
          delta *= dactf(c, p.last_act);
        }
        dl[r * ldc + a] = delta;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
      if ((tid & 31) == 0) red[tid >> 5] = l;
      __syncthreads();
      if (tid == 0) {
        double s = 0.0;
        for (int k = 0; k < NT / 32; ++k) s += red[k];
        loss_acc += s;
      }
    }
    // (5) backward through the tail layers l = L..2
    for (int l = L; l >= 2; --l) {
      const int win = p.w[l - 1], wout = p.w[l];
      const float *in = l == 2 ? z1 : zt + (l - 3) * R * TW;
      const int ldin = l == 2 ? H1P : TW;
      const float *dl = dt + (l - 2) * R * TW;
      float *dprev = l == 2 ? d1 : dt + (l - 3) * R * TW;
      const float *sg = p.seg[l - 1];
      float *gp = part + p.off[l - 1];
      for (int o = tid; o < (win + 1) * wout; o += NT) {
        const int k = o / wout, j = o - k * wout;
        float s = 0.f;
        if (k < win) {
#pragma unroll
          for (int r = 0; r < R; ++r) s = fmaf(in[r * ldin + k], dl[r * TW + j], s);
        } else {
#pragma unroll
          for (int r = 0; r < R; ++r) s += dl[r * TW + j];
        }
        gp[o] = first ? s : gp[o] + s;
      }
      for (int o = tid; o < R * win; o += NT) {
        const int r = o / win, k = o - r * win;
        float s = 0.f;
        for (int j = 0; j < wout; ++j) s = fmaf(dl[r * TW + j], __ldg(sg + k * wout + j), s);
        dprev[r * ldin + k] = s * dactf(in[r * ldin + k], p.act);
      }
      __syncthreads();
    }
    // (6) layer-1 gradient [(F+1) x h1] = [x_hat, 1]^T * delta_1 over this group's 16 rows
    {
      const int j = tid & (H1P - 1), half = tid >> 7;
      if (j < h1) {
        float dc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) dc[r] = d1[r * H1P + j];
        const int kh = (p.F + KQ) / KQ;
        const int k0 = min(p.F + 1, half * kh), k1 = min(p.F + 1, k0 + kh);
        const float4 *x4 = reinterpret_cast<const float4 *>(xT);
        float *gp = part + j;
#pragma unroll 4
        for (int k = k0; k < k1; ++k) {
          float s = 0.f;
          if (k < p.F) {
            // independent partial sums per float4: a single R-deep FMA chain exposed its full latency
            float ps[R / 4];
#pragma unroll
            for (int q = 0; q < R / 4; ++q) {
              const float4 a = x4[k * (R / 4) + q];
              float t = a.x * dc[4 * q];
              t = fmaf(a.y, dc[4 * q + 1], t);
              t = fmaf(a.z, dc[4 * q + 2], t);
              ps[q] = fmaf(a.w, dc[4 * q + 3], t);
            }
            if (R == 16) s = (ps[0] + ps[1]) + (ps[2 % (R / 4)] + ps[3 % (R / 4)]);
            else s = ps[0] + ps[1 % (R / 4)];
          } else {
#pragma unroll
            for (int r = 0; r < R; ++r) s += dc[r];
          }
          const int64_t o = (int64_t)k * h1;
          gp[o] = first ? s : gp[o] + s;
        }
      }
    }
    first = false;
  }
  if (tid == 0) p.part_loss[blockIdx.x] = loss_acc;
}

struct ReduceP {
  const float *part;
  int64_t part_stride;
  int nparts;
  int nseg;
  int64_t off[ISOKANN_MAX_LAYERS + 1];
  float *dest[ISOKANN_MAX_LAYERS];
  const double *part_loss;
  float *packed_tail;
};

__global__ void narrow_reduce_kernel(ReduceP p) {
  const int64_t total = p.off[p.nseg];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    float s = p.part[i];
    for (int c = 1; c < p.nparts; ++c) s += p.part[(int64_t)c * p.part_stride + i];
    int sgi = 0;
    while (i >= p.off[sgi + 1]) ++sgi;
    p.dest[sgi][i - p.off[sgi]] = s;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double l = 0.0;
    for (int c = 0; c < p.nparts; ++c) l += p.part_loss[c];
    const float hi = (float)l;
    p.packed_tail[0] = hi;
    p.packed_tail[1] = (hi == hi && fabsf(hi) != INFINITY) ? (float)(l - (double)hi) : 0.f;
  }
}

// Forward pass of a tiny net (every width <= 16, e.g. smallnet [2,8,8,8,1] of the Langevin toy systems,
// reference src/models.jl:102-108): one thread per sample, all [W; b] segments in shared memory,
// activations in registers.  HBM traffic: 4*F bytes in, 4*d bytes out per sample.
struct TinyP {
  const float *in;      // [M x F] records (identity featurizer), optionally gathered
  int64_t M;
  int L;
  int w[ISOKANN_MAX_LAYERS + 1];
  const float *seg[ISOKANN_MAX_LAYERS];
  int act, last_act;
  float *out;           // [M x d]
};

__global__ void __launch_bounds__(256) tiny_forward_kernel(TinyP p) {
  __shared__ float sw[ISOKANN_MAX_LAYERS * 17 * 16];
  __shared__ int soff[ISOKANN_MAX_LAYERS + 1];
  if (threadIdx.x == 0) {
    int o = 0;
    for (int l = 0; l < p.L; ++l) {
      soff[l] = o;
      o += (p.w[l] + 1) * p.w[l + 1];
    }
    soff[p.L] = o;
  }
  __syncthreads();
  for (int l = 0; l < p.L; ++l) {
    const int n = (p.w[l] + 1) * p.w[l + 1];
    for (int e = threadIdx.x; e < n; e += blockDim.x) sw[soff[l] + e] = __ldg(p.seg[l] + e);
  }
  __syncthreads();
  const int F = p.w[0];
  for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < p.M; m += (int64_t)gridDim.x * blockDim.x) {
    float h[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) h[k] = k < F ? __ldg(p.in + m * F + k) : 0.f;
    for (int l = 0; l < p.L; ++l) {
      const int win = p.w[l], wout = p.w[l + 1];
      const float *W = sw + soff[l];
      float a[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) a[j] = j < wout ? W[win * wout + j] : 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        if (k < win) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (j < wout) a[j] = fmaf(h[k], W[k * wout + j], a[j]);
        }
      }
      const int kind = l == p.L - 1 ? p.last_act : p.act;
#pragma unroll
      for (int j = 0; j < 16; ++j) h[j] = j < wout ? actf(a[j], kind) : 0.f;
    }
    const int d = p.w[p.L];
#pragma unroll
    for (int j = 0; j < kMaxD; ++j)
      if (j < d) p.out[m * d + j] = h[j];
  }
}

}  // namespace

bool tiny_forward_eligible(const isokann_config &g) {
  if (g.layernorm || g.featurizer != ISOKANN_FEAT_IDENTITY) return false;
  for (int l = 0; l <= g.n_layers; ++l)
    if (g.widths[l] > 16) return false;
  return g.widths[g.n_layers] <= kMaxD;
}

void launch_tiny_forward(Ctx &c, const float *in, int64_t M, float *out) {
  if (M <= 0) return;
  TinyP p{};
  p.in = in; p.M = M; p.L = c.L;
  for (int l = 0; l <= c.L; ++l) p.w[l] = c.cfg.widths[l];
  for (int l = 0; l < c.L; ++l) p.seg[l] = c.params.p + c.off_w[l];
  p.act = c.cfg.activation; p.last_act = c.cfg.last_activation;
  p.out = out;
  int grid = (int)std::min<int64_t>((M + 255) / 256, (int64_t)c.num_sms * 8);
  c.timer.begin(KC_GEMM, c.stream);
  tiny_forward_kernel<<<grid, 256, 0, c.stream>>>(p);
  c.timer.end(c.stream);
  IK_CUDA(cudaGetLastError());
  double macs = 0;
  for (int l = 0; l < c.L; ++l) macs += (double)c.cfg.widths[l] * c.cfg.widths[l + 1];
  c.count_launch(KC_GEMM, 2.0 * macs * (double)M);
}

bool narrow_train_eligible(const isokann_config &g) {
  if (g.n_layers < 1 || g.n_layers > 4) return false;
  if (g.widths[1] > H1P) return false;
  for (int l = 2; l <= g.n_layers; ++l)
    if (g.widths[l] > TW) return false;
  if (g.widths[g.n_layers] > kMaxD) return false;
  // x_hat tile + activations must fit shared memory
  const size_t smem = ((size_t)g.widths[0] * RMAX + (2 + KQ - 1) * RMAX * H1P + 2 * ISOKANN_MAX_LAYERS * RMAX * TW) * sizeof(float);
  return smem <= 200 * 1024;
}

// xhat: [Bloc x F] normalised features of the minibatch slice (row-major); writes the gradient of the folded
// first segment to g0 ((F+1) x h1) and of the tail layers to gtail[l-1], and the packed step loss
void launch_narrow_train(Ctx &c, const float *xhat, int64_t Bloc, const int64_t *idx, double Bglobal,
                         const float *seg0, float *g0) {
  const int F = c.F, L = c.L;
  NarrowP p{};
  p.xhat = xhat; p.ldx = F; p.B = (int)Bloc; p.F = F; p.L = L;
  for (int l = 0; l <= L; ++l) p.w[l] = c.cfg.widths[l];
  p.seg[0] = seg0;
  for (int l = 1; l < L; ++l) p.seg[l] = c.params.p + c.off_w[l];
  int64_t off = 0;
  for (int l = 0; l < L; ++l) {
    p.off[l] = off;
    off += (int64_t)(c.cfg.widths[l] + 1) * c.cfg.widths[l + 1];
  }
  p.off[L] = off;
  p.target = c.target.p; p.idx = idx; p.wloss = c.w_loss.p; p.Bglobal = Bglobal;
  p.act = c.cfg.activation; p.last_act = c.cfg.last_activation;
  // 16-row groups unless they would leave a third of the SMs without work (B = 1000: 63 groups on 148 SMs)
  const int R = cdiv(Bloc, RMAX) * 3 < c.num_sms * 2 ? 8 : RMAX;
  p.groups = cdiv(Bloc, R);
  const int nparts = std::min(p.groups, 2 * c.num_sms);
  const int64_t stride = (off + 3) & ~(int64_t)3;
  c.splitk.ensure((size_t)nparts * stride);
  c.red_d.ensure((size_t)std::max(nparts, 1024));
  p.part = c.splitk.p; p.part_stride = stride; p.part_loss = c.red_d.p;
  const size_t smem = ((size_t)F * R + (2 + KQ - 1) * R * H1P + 2 * ISOKANN_MAX_LAYERS * R * TW) * sizeof(float);
  if (c.attr_needed(Ctx::ATTR_NARROW)) {
    IK_CUDA(cudaFuncSetAttribute(narrow_fwd_bwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    IK_CUDA(cudaFuncSetAttribute(narrow_fwd_bwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  c.timer.begin(KC_GEMM, c.stream);
  if (R == 8) narrow_fwd_bwd_kernel<8><<<nparts, NT, smem, c.stream>>>(p);
  else narrow_fwd_bwd_kernel<16><<<nparts, NT, smem, c.stream>>>(p);
  c.timer.end(c.stream);
  IK_CUDA(cudaGetLastError());
  double macs = 0;
  for (int l = 0; l < L; ++l) macs += (double)c.cfg.widths[l] * c.cfg.widths[l + 1];
  c.count_launch(KC_GEMM, 6.0 * macs * (double)Bloc);

  ReduceP r{};
  r.part = c.splitk.p; r.part_stride = stride; r.nparts = nparts; r.nseg = L;
  for (int l = 0; l <= L; ++l) r.off[l] = p.off[l];
  r.dest[0] = g0;
  for (int l = 1; l < L; ++l) r.dest[l] = c.grads.p + c.off_w[l];
  r.part_loss = c.red_d.p;
  r.packed_tail = c.grads.p + c.P;
  int grid = (int)std::min<int64_t>((off + 255) / 256, (int64_t)c.num_sms * 4);
  c.timer.begin(KC_REDUCE, c.stream);
  narrow_reduce_kernel<<<grid, 256, 0, c.stream>>>(r);
  c.timer.end(c.stream);
  IK_CUDA(cudaGetLastError());
  c.count_launch(KC_REDUCE);
}

}  // namespace ik
