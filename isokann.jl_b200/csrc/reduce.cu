// Reductions of the Koopman / target stage: K-mean (expectation, reference
// src/isotarget.jl:18 and the weighted form src/data.jl:215), global min/max + affine map
// (shiftscale, src/isotarget.jl:36-42), and the small-matrix reductions of the N-D targets
// (TransformISA src/isotarget.jl:81-107, fixperm :120-127, TransformPseudoInv :152-179).
// All are HBM-bound passes over d*N floats; block-level warp-shuffle reductions write one
// partial per block and the (tiny) final stage runs where the result is consumed.
#include "common.cuh"

namespace ik {

constexpr int kRedThreads = 256;

static int red_grid(const Ctx &c, int64_t n) {
  int64_t g = (n + kRedThreads - 1) / kRedThreads;
  int64_t cap = (int64_t)c.num_sms * 4;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block sum of a double; result valid in thread 0
__device__ double block_sum_d(double v) {
  __shared__ double sh[kRedThreads / 32];
  v = warp_sum_d(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x < 32) {
    r = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
    r = warp_sum_d(r);
  }
  return r;
}

// ---- Koopman expectation: out[n,:] = (sum_k chi[n*K+k,:] (* w[n*K+k])) / K, sequential in k ----
__global__ void kmean_kernel(const float *__restrict__ chi, const float *__restrict__ w, int64_t n, int K, int d,
                             float *__restrict__ out) {
  const int64_t total = n * d;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / d;
    const int a = (int)(t - i * d);
    const float *p = chi + i * K * d + a;
    float s;
    if (w) {
      const float *q = w + i * K;
      s = p[0] * q[0];
      for (int k = 1; k < K; ++k) s += p[(int64_t)k * d] * q[k];
    } else {
      s = p[0];
      for (int k = 1; k < K; ++k) s += p[(int64_t)k * d];
    }
    out[t] = s / (float)K;
  }
}

void launch_kmean(Ctx &c, const float *chi, const float *weights, int64_t n, int K, int d, float *out) {
  if (n <= 0) return;
  c.timer.begin(KC_REDUCE, c.stream);
  kmean_kernel<<<red_grid(c, n * d), kRedThreads, 0, c.stream>>>(chi, weights, n, K, d, out);
  c.timer.end(c.stream);
  IK_CUDA(cudaGetLastError());
  c.count_launch(KC_REDUCE);
}

// ---- shiftscale ----
// NaN-propagating min/max like Julia's extrema
__device__ __forceinline__ float nanmin(float a, float b) { return (b < a || b != b) ? b : a; }
__device__ __forceinline__ float nanmax(float a, float b) { return (b > a || b != b) ? b : a; }

__global__ void minmax_kernel(const float *__restrict__ x, int64_t n, float *__restrict__ partials) {
  __shared__ float smin[kRedThreads / 32], smax[kRedThreads / 32];
  float mn = INFINITY, mx = -INFINITY;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    mn = nanmin(mn, v);
    mx = nanmax(mx, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = nanmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = nanmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) {
    smin[threadIdx.x >> 5] = mn;
    smax[threadIdx.x >> 5] = mx;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kRedThreads / 32; ++w) {
      mn = nanmin(mn, smin[w]);
      mx = nanmax(mx, smax[w]);
    }
    partials[2 * blockIdx.x] = mn;
    partials[2 * blockIdx.x + 1] = mx;
  }
}

void launch_minmax(Ctx &c, const float *x, int64_t n, float *partials, int *nblocks_out) {
  int grid = red_grid(c, n);
  if (grid > 512) grid = 512;
  c.timer.begin(KC_REDUCE, c.stream);
  minmax_kernel<<<grid, kRedThreads, 0, c.stream>>>(x, n, partials);
  c.timer.end(c.stream);
  IK_CUDA(cudaGetLastError());
  c.count_launch(KC_REDUCE);
  *nblocks_out = grid;
}

__global__ void shiftscale_kernel(const float *__restrict__ x, int64_t n, const float *__restrict__ partials,
                                  int nblocks, float *__restrict__ out, int *__restrict__ flags) {
  __shared__ float s_mn, s_mx;
  if (threadIdx.x < 32) {
    float mn = INFINITY, mx = -INFINITY;
    for (int b = threadIdx.x; b < nblocks; b += 32) {
      mn = nanmin(mn, partials[2 * b]);
      mx = nanmax(mx, partials[2 * b + 1]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn = nanmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = nanmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (threadIdx.x == 0) {
      s_mn = mn;
      s_mx = mx;
    }
  }
  __syncthreads();
  const float mn = s_mn, mx = s_mx;
  if (!(mx > mn)) {  // src/isotarget.jl:39
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(flags, FLAG_CONSTANT_CHI);
    return;
  }
  const float range = mx - mn;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (x[i] - mn) / range;
}

void launch_shiftscale(Ctx &c, const float *x, int64_t n, const float *partials, int nblocks, float *out,
                       int *flags) {
  c.timer.begin(KC_REDUCE, c.stream);
  shiftscale_kernel<<<red_grid(c, n), kRedThreads, 0, c.stream>>>(x, n, partials, nblocks, out, flags);
  c.timer.end(c.stream);
  IK_CUDA(cudaGetLastError());
  c.count_launch(KC_REDUCE);
}

__global__ void fill_kernel(float *x, int64_t n, float v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    x[i] = v;
}
void launch_fill(Ctx &c, float *x, int64_t n, float v) {
  if (n <= 0) return;
  fill_kernel<<<red_grid(c, n), kRedThreads, 0, c.stream>>>(x, n, v);
  IK_CUDA(cudaGetLastError());
  c.count_launch(KC_REDUCE);
}

// ---- Gram matrices for TransformPseudoInv / whitening: per block and row a,
//      partial[(blk*d + a)*2d + b]     = sum_n chi[n,a]  * kchi[n,b]
//      partial[(blk*d + a)*2d + d + b] = sum_n kchi[n,a] * kchi[n,b]            (fp64) ----
__global__ void gram_kernel(const float *__restrict__ chi, const float *__restrict__ kchi, int64_t n, int d,
                            double *__restrict__ partials) {
  const int a = blockIdx.y;
  double acc[2 * kMaxD];
#pragma unroll
  for (int b = 0; b < 2 * kMaxD; ++b) acc[b] = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double ca = chi ? (double)chi[i * d + a] : 0.0;
    const double ka = (double)kchi[i * d + a];
#pragma unroll
    for (int b = 0; b < kMaxD; ++b) {
      if (b < d) {
        const double kb = (double)kchi[i * d + b];
        acc[b] += ca * kb;
        acc[kMaxD + b] += ka * kb;
      }
    }
  }
  for (int b = 0; b < d; ++b) {
    const double s1 = block_sum_d(acc[b]);
    const double s2 = block_sum_d(acc[kMaxD + b]);
    if (threadIdx.x == 0) {
      partials[((int64_t)blockIdx.x * d + a) * 2 * d + b] = s1;
      partials[((int64_t)blockIdx.x * d + a) * 2 * d + d + b] = s2;
    }
  }
}

void launch_gram(Ctx &c, const float *chi, const float *kchi, int64_t n, int d, double *partials, int *nblocks_out) {
  int gx = red_grid(c, n);
  if (gx > 256) gx = 256;
  dim3 grid(gx, d);
  c.timer.begin(KC_REDUCE, c.stream);
  gram_kernel<<<grid, kRedThreads, 0, c.stream>>>(chi, kchi, n, d, partials);
  c.timer.end(c.stream);
  IK_CUDA(cudaGetLastError());
  c.count_launch(KC_REDUCE);
  *nblocks_out = gx;
}

// ---- second moments of the augmented rows u = [chi[n,:], 1], v = [kchi[n,:], 1] for the diagnostics (rates:
//      src/iso.jl:339-351, residual_ritz/subspace: src/isotarget.jl:787-821).  With e = d + 1, per block and row a:
//      partial[((blk*e + a)*3 + 0)*e + b] = sum_n u_a u_b     (chi chi', column sums of chi, N)
//      partial[((blk*e + a)*3 + 1)*e + b] = sum_n v_a u_b     (Kchi chi')
//      partial[((blk*e + a)*3 + 2)*e + b] = sum_n v_a v_b     (Kchi Kchi')                         (fp64) ----
__global__ void moments_kernel(const float *__restrict__ chi, const float *__restrict__ kchi, int64_t n, int d,
                               double *__restrict__ partials) {
  constexpr int E = kMaxD + 1;
  const int a = blockIdx.y, e = d + 1;
  double acc[3 * E];
#pragma unroll
  for (int b = 0; b < 3 * E; ++b) acc[b] = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double ua = a < d ? (double)chi[i * d + a] : 1.0;
    const double va = a < d ? (double)kchi[i * d + a] : 1.0;
#pragma unroll
    for (int b = 0; b < E; ++b) {
      if (b < e) {
        const double ub = b < d ? (double)chi[i * d + b] : 1.0;
        const double vb = b < d ? (double)kchi[i * d + b] : 1.0;
        acc[b] += ua * ub;
        acc[E + b] += va * ub;
        acc[2 * E + b] += va * vb;
      }
    }
  }
  for (int k = 0; k < 3; ++k)
    for (int b = 0; b < e; ++b) {
      const double s = block_sum_d(acc[k * E + b]);
      if (threadIdx.x == 0) partials[(((int64_t)blockIdx.x * e + a) * 3 + k) * e + b] = s;
    }
}

void launch_moments(Ctx &c, const float *chi, const float *kchi, int64_t n, int d, double *partials, int *nblocks_out) {
  int gx = red_grid(c, n);
  if (gx > 256) gx = 256;
  dim3 grid(gx, d + 1);
  moments_kernel<<<grid, kRedThreads, 0, c.stream>>>(chi, kchi, n, d, partials);
  IK_CUDA(cudaGetLastError());
  c.count_launch(KC_REDUCE);
  *nblocks_out = gx;
}

// ---- residual matrices of the diagnostics: r[n,j] = sum_b A[j,b] kchi[n,b] - sum_b B[j,b] chi[n,b]   (fp64)
//      out (optional): N x d column-major doubles, as Julia holds `res` / `residues`
//      partial[(blk*d + j)*2 + {0,1}] = sum_n r[n,j]^2, sum_n (sum_b A[j,b] kchi[n,b])^2
__global__ void resid_kernel(const float *__restrict__ kchi, const float *__restrict__ chi, int64_t n, int d, Mat8 A,
                             Mat8 B, double *__restrict__ out, double *__restrict__ partials) {
  const int j = blockIdx.y;
  double s_r = 0.0, s_k = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double tk = 0.0, tv = 0.0;
#pragma unroll
    for (int b = 0; b < kMaxD; ++b)
      if (b < d) {
        tk += A.m[j * d + b] * (double)kchi[i * d + b];
        tv += B.m[j * d + b] * (double)chi[i * d + b];
      }
    const double r = tk - tv;
    if (out) out[(int64_t)j * n + i] = r;
    s_r += r * r;
    s_k += tk * tk;
  }
  const double t0 = block_sum_d(s_r);
  const double t1 = block_sum_d(s_k);
  if (threadIdx.x == 0) {
    partials[((int64_t)blockIdx.x * d + j) * 2] = t0;
    partials[((int64_t)blockIdx.x * d + j) * 2 + 1] = t1;
  }
}

void launch_resid(Ctx &c, const float *kchi, const float *chi, int64_t n, int d, const Mat8 &A, const Mat8 &B,
                  double *out_colmajor, double *partials, int *nblocks_out) {
  int gx = red_grid(c, n);
  if (gx > 256) gx = 256;
  dim3 grid(gx, d);
  resid_kernel<<<grid, kRedThreads, 0, c.stream>>>(kchi, chi, n, d, A, B, out_colmajor, partials);
  IK_CUDA(cudaGetLastError());
  c.count_launch(KC_REDUCE);
  *nblocks_out = gx;
}

// ---- t[n,a] = sum_b mat[a,b] * kchi[n,b] with three consumers:
//   mode 0 (L1)   : partial[blk*d + a]       = sum_n |t[n,a]|
//   mode 1 (COST) : partial[(blk*d + a)*d+b] = sum_n |t[n,a] - chi[n,b]|        (fixperm cost matrix)
//   mode 2 (WRITE): target[n,a] = float(t);  partial[(blk*d + a)*2 + {0,1}] = sum t, sum t^2 (of the rounded value)
__global__ void apply_kernel(int mode, const float *__restrict__ kchi, const float *__restrict__ chi, int64_t n, int d,
                             Mat8 mat, float *__restrict__ target, double *__restrict__ partials) {
  const int a = blockIdx.y;
  double acc[kMaxD];
#pragma unroll
  for (int b = 0; b < kMaxD; ++b) acc[b] = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double t = 0.0;
#pragma unroll
    for (int b = 0; b < kMaxD; ++b)
      if (b < d) t += mat.m[a * d + b] * (double)kchi[i * d + b];
    if (mode == 0) {
      acc[0] += fabs(t);
    } else if (mode == 1) {
#pragma unroll
      for (int b = 0; b < kMaxD; ++b)
        if (b < d) acc[b] += fabs(t - (double)chi[i * d + b]);
    } else {
      const float tf = (float)t;
      target[i * d + a] = tf;
      acc[0] += (double)tf;
      acc[1] += (double)tf * (double)tf;
    }
  }
  const int nacc = mode == 0 ? 1 : (mode == 1 ? d : 2);
  for (int b = 0; b < nacc; ++b) {
    const double s = block_sum_d(acc[b]);
    if (threadIdx.x == 0) partials[((int64_t)blockIdx.x * d + a) * nacc + b] = s;
  }
}

void launch_apply(Ctx &c, int mode, const float *kchi, const float *chi, int64_t n, int d, const Mat8 &mat,
                  float *target_out, double *partials, int *nblocks_out) {
  int gx = red_grid(c, n);
  if (gx > 256) gx = 256;
  dim3 grid(gx, d);
  c.timer.begin(KC_REDUCE, c.stream);
  apply_kernel<<<grid, kRedThreads, 0, c.stream>>>(mode, kchi, chi, n, d, mat, target_out, partials);
  c.timer.end(c.stream);
  IK_CUDA(cudaGetLastError());
  c.count_launch(KC_REDUCE);
  *nblocks_out = gx;
}

// ---- one round of the inner simplex algorithm (PCCAPlus.indexmap): argmax_n of the 2-norm of
//      the row after replaying the affine/projection steps of the previous rounds (fp64) ----
__device__ __host__ inline double isa_row_norm(const float *row, const IsaReplay &rp, double *xout) {
  double x[kMaxD], y[kMaxD];
  const int d = rp.d;
  for (int b = 0; b < d; ++b) x[b] = (double)row[b];
  for (int b = 0; b < d; ++b) {  // x <- x * W
    double s = 0.0;
    for (int a = 0; a < d; ++a) s += x[a] * rp.pre[a * d + b];
    y[b] = s;
  }
  if (rp.rounds >= 1)
    for (int b = 0; b < d; ++b) y[b] -= rp.x0[b];
  for (int j = 1; j < rp.rounds; ++j) {
    double dot = 0.0;
    for (int b = 0; b < d; ++b) {
      y[b] /= rp.r[j];
      dot += y[b] * rp.v[j][b];
    }
    for (int b = 0; b < d; ++b) y[b] -= dot * rp.v[j][b];
  }
  double q = 0.0;
  for (int b = 0; b < d; ++b) {
    q += y[b] * y[b];
    if (xout) xout[b] = y[b];
  }
  return sqrt(q);
}

__global__ void isa_argmax_kernel(const float *__restrict__ kchi, int64_t n, IsaReplay rp,
                                  ArgmaxPartial *__restrict__ partials) {
  __shared__ double sval[kRedThreads];
  __shared__ long long sidx[kRedThreads];
  double best = -1.0;
  long long bi = -1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = isa_row_norm(kchi + i * rp.d, rp, nullptr);
    if (v > best || (v != v && best == best)) {  // first maximum; NaN wins like Julia's argmax
      best = v;
      bi = i;
    }
  }
  sval[threadIdx.x] = best;
  sidx[threadIdx.x] = bi;
  __syncthreads();
  for (int s = kRedThreads / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      const double v2 = sval[threadIdx.x + s];
      const long long i2 = sidx[threadIdx.x + s];
      const double v1 = sval[threadIdx.x];
      const long long i1 = sidx[threadIdx.x];
      bool take = false;
      if (i2 >= 0) {
        if (i1 < 0) take = true;
        else if (v1 != v1) take = (v2 != v2) && i2 < i1;
        else if (v2 != v2) take = true;
        else take = (v2 > v1) || (v2 == v1 && i2 < i1);
      }
      if (take) {
        sval[threadIdx.x] = v2;
        sidx[threadIdx.x] = i2;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    partials[blockIdx.x].val = sval[0];
    partials[blockIdx.x].idx = sidx[0];
  }
}

void launch_isa_argmax(Ctx &c, const float *kchi, int64_t n, const IsaReplay &rp, ArgmaxPartial *partials,
                       int *nblocks_out) {
  int gx = red_grid(c, n);
  if (gx > 256) gx = 256;
  c.timer.begin(KC_REDUCE, c.stream);
  isa_argmax_kernel<<<gx, kRedThreads, 0, c.stream>>>(kchi, n, rp, partials);
  c.timer.end(c.stream);
  IK_CUDA(cudaGetLastError());
  c.count_launch(KC_REDUCE);
  *nblocks_out = gx;
}

// host twin of the replay, used to rebuild the selected row exactly as the kernel saw it
double isa_row_norm_host(const float *row, const IsaReplay &rp, double *xout) { return isa_row_norm(row, rp, xout); }

// ---- validationloss (src/iso.jl:160-168): sum over the validation points of (c - (k1 - mn) / (mx - mn))^2 ----
__global__ void valloss_kernel(const float *__restrict__ c, const float *__restrict__ k1, int64_t n, float mn, float mx,
                               double *__restrict__ partials) {
  double s = 0.0;
  const float span = mx - mn;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float r = c[i] - (k1[i] - mn) / span;   // Float32 arithmetic like the reference's broadcast
    s += (double)(r * r);
  }
  s = block_sum_d(s);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

void launch_valloss(Ctx &c, const float *chi, const float *k1, int64_t n, float mn, float mx, double *partials,
                    int *nblocks_out) {
  int grid = red_grid(c, n);
  if (grid > 256) grid = 256;
  valloss_kernel<<<grid, kRedThreads, 0, c.stream>>>(chi, k1, n, mn, mx, partials);
  IK_CUDA(cudaGetLastError());
  c.count_launch(KC_REDUCE);
  *nblocks_out = grid;
}

// ---- Float64 coordinates -> Float32 (round to nearest), isokann_set_data_f64 ----
__global__ void f64_to_f32_kernel(const double *__restrict__ in, int64_t n, float *__restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (float)in[i];
}
void launch_f64_to_f32(Ctx &c, const double *in, int64_t n, float *out) {
  if (n <= 0) return;
  f64_to_f32_kernel<<<red_grid(c, n), kRedThreads, 0, c.stream>>>(in, n, out);
  IK_CUDA(cudaGetLastError());
  c.count_launch(KC_REDUCE);
}

// ---- perm (1-based, Julia) -> 0-based device indices ----
__global__ void perm0_kernel(const int64_t *__restrict__ p1, int64_t n, int64_t *__restrict__ p0) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    p0[i] = p1[i] - 1;
}
void launch_perm_to_zero_based(Ctx &c, const int64_t *perm1, int64_t n, int64_t *out0) {
  perm0_kernel<<<red_grid(c, n), kRedThreads, 0, c.stream>>>(perm1, n, out0);
  IK_CUDA(cudaGetLastError());
  c.count_launch(KC_REDUCE);
}

// ---- all-gathered, padded shards [world][nmax][d] -> dense [N][d] (contiguous-split rule) ----
__global__ void compact_kernel(const float *__restrict__ padded, int world, int64_t nmax, int64_t N, int d,
                               float *__restrict__ out) {
  const int64_t total = N * d;
  const int64_t base = N / world, rem = N % world;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / d;
    const int a = (int)(t - i * d);
    // rank r owns [r*base + min(r,rem), ...) of size base + (r<rem)
    int64_t r, off;
    const int64_t cut = rem * (base + 1);
    if (i < cut) {
      r = i / (base + 1);
      off = i - r * (base + 1);
    } else {
      r = rem + (base > 0 ? (i - cut) / base : 0);
      off = i - cut - (r - rem) * base;
    }
    out[t] = padded[(r * nmax + off) * d + a];
  }
}
void launch_compact_gather(Ctx &c, const float *padded, int world, int64_t nmax, int64_t N, int d, float *out) {
  compact_kernel<<<red_grid(c, N * d), kRedThreads, 0, c.stream>>>(padded, world, nmax, N, d, out);
  IK_CUDA(cudaGetLastError());
  c.count_launch(KC_REDUCE);
}

}  // namespace ik
