// Kernels around the tcgen05 GEMM: operand preparation (fp32 -> bf16 hi/lo split, weight layouts)
// and the thin last Dense layer (out = d <= 8), which is a memory-bound row reduction rather
// than a GEMM.  All of them are HBM-bound passes with 16-byte vector accesses.
#include <cuda_fp16.h>

#include "tc.cuh"

namespace ik {

namespace {

__device__ __forceinline__ void split1(float x, __nv_bfloat16 &h, __nv_bfloat16 &l) {
  h = __float2bfloat16_rn(x);
  l = __float2bfloat16_rn(x - __bfloat162float(h));
}
__device__ __forceinline__ float bflo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bfhi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }
__device__ __forceinline__ float act_f(float a, int kind) {
  switch (kind) {
    case ISOKANN_ACT_SIGMOID: return __fdividef(1.0f, 1.0f + __expf(-a));
    case ISOKANN_ACT_TANH: return tanhf(a);
    case ISOKANN_ACT_RELU: return fmaxf(a, 0.f);
    default: return a;
  }
}
__device__ __forceinline__ float dact_f(float z, int kind) {
  switch (kind) {
    case ISOKANN_ACT_SIGMOID: return z * (1.0f - z);
    case ISOKANN_ACT_TANH: return 1.0f - z * z;
    case ISOKANN_ACT_RELU: return z > 0.f ? 1.0f : 0.f;
    default: return 1.0f;
  }
}

// ---- weights: fp32 [fin x fout] -> Wd split (same orientation) and Wf split (transposed) ----
__global__ void __launch_bounds__(256) prep_weights_kernel(const float *__restrict__ seg, int fin, int fout,
                                                           __nv_bfloat16 *__restrict__ wf_hi,
                                                           __nv_bfloat16 *__restrict__ wf_lo, int64_t ld_f,
                                                           __nv_bfloat16 *__restrict__ wd_hi,
                                                           __nv_bfloat16 *__restrict__ wd_lo, int64_t ld_d) {
  __shared__ float t[64][65];
  const int i0 = blockIdx.x * 64;  // fin index
  const int j0 = blockIdx.y * 64;  // fout index
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  for (int r = ty; r < 64; r += 4) {
    const int i = i0 + r, j = j0 + tx;
    float v = 0.f;
    if (i < fin && j < fout) {
      v = seg[(int64_t)i * fout + j];
      if (wd_hi) {
        __nv_bfloat16 h, l;
        split1(v, h, l);
        wd_hi[(int64_t)i * ld_d + j] = h;
        wd_lo[(int64_t)i * ld_d + j] = l;
      }
    }
    t[r][tx] = v;
  }
  __syncthreads();
  for (int r = ty; r < 64; r += 4) {
    const int j = j0 + r, i = i0 + tx;
    if (j < fout && i < fin) {
      __nv_bfloat16 h, l;
      split1(t[tx][r], h, l);
      wf_hi[(int64_t)j * ld_f + i] = h;
      wf_lo[(int64_t)j * ld_f + i] = l;
    }
  }
}

// ---- weights: fp32 [fin x fout] -> Wf16 [fout x ld_f] (transposed), each weight rounded once to fp16 ----
__global__ void __launch_bounds__(256) prep_weights_f16_kernel(const float *__restrict__ seg, int fin, int fout,
                                                               __half *__restrict__ wf, int64_t ld_f) {
  __shared__ float t[64][65];
  const int i0 = blockIdx.x * 64, j0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  for (int r = ty; r < 64; r += 4) {
    const int i = i0 + r, j = j0 + tx;
    t[r][tx] = (i < fin && j < fout) ? seg[(int64_t)i * fout + j] : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 64; r += 4) {
    const int j = j0 + r, i = i0 + tx;
    if (j < fout && i < fin) wf[(int64_t)j * ld_f + i] = __float2half_rn(t[tx][r]);
  }
}

// ---- thin forward: one warp per row ----
__global__ void __launch_bounds__(256) thin_forward_kernel(const __nv_bfloat16 *__restrict__ z_hi,
                                                           const __nv_bfloat16 *__restrict__ z_lo, int64_t M, int fin,
                                                           int64_t ldz, const float *__restrict__ seg, int d, int act,
                                                           float *__restrict__ chi) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t m = wid; m < M; m += nw) {
    float acc[kMaxD];
#pragma unroll
    for (int a = 0; a < kMaxD; ++a) acc[a] = 0.f;
    const uint4 *ph = reinterpret_cast<const uint4 *>(z_hi + m * ldz);
    const uint4 *pl = reinterpret_cast<const uint4 *>(z_lo + m * ldz);
    for (int k8 = lane; k8 * 8 < fin; k8 += 32) {
      const uint4 h = __ldg(ph + k8), l = __ldg(pl + k8);
      const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = k8 * 8 + 2 * e;
        const float z0 = bflo(hw[e]) + bflo(lw[e]);
        const float z1 = bfhi(hw[e]) + bfhi(lw[e]);
#pragma unroll
        for (int a = 0; a < kMaxD; ++a) {
          if (a < d) {
            if (k < fin) acc[a] = fmaf(z0, __ldg(seg + (int64_t)k * d + a), acc[a]);
            if (k + 1 < fin) acc[a] = fmaf(z1, __ldg(seg + (int64_t)(k + 1) * d + a), acc[a]);
          }
        }
      }
    }
#pragma unroll
    for (int a = 0; a < kMaxD; ++a) {
      if (a < d) {
        float s = acc[a];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) chi[m * d + a] = act_f(s + __ldg(seg + (int64_t)fin * d + a), act);
      }
    }
  }
}

// ---- thin dgrad: 8 consecutive k per thread ----
__global__ void __launch_bounds__(256) thin_dgrad_kernel(const float *__restrict__ delta, int64_t M, int d,
                                                         const float *__restrict__ seg, int fin,
                                                         const __nv_bfloat16 *__restrict__ z_hi,
                                                         const __nv_bfloat16 *__restrict__ z_lo, int64_t ldz, int act,
                                                         __nv_bfloat16 *__restrict__ out_hi,
                                                         __nv_bfloat16 *__restrict__ out_lo, int64_t ldo) {
  const int k8n = (int)(ldo / 8);
  const int64_t total = M * k8n;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = t / k8n;
    const int k0 = (int)(t - m * k8n) * 8;
    float dl[kMaxD];
#pragma unroll
    for (int a = 0; a < kMaxD; ++a) dl[a] = a < d ? __ldg(delta + m * d + a) : 0.f;
    uint4 h = make_uint4(0, 0, 0, 0), l = h;
    if (k0 < ldz) {
      h = __ldg(reinterpret_cast<const uint4 *>(z_hi + m * ldz + k0));
      l = __ldg(reinterpret_cast<const uint4 *>(z_lo + m * ldz + k0));
    }
    const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
    uint32_t oh[4], ol[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float g[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int k = k0 + 2 * e + q;
        float s = 0.f;
        if (k < fin) {
#pragma unroll
          for (int a = 0; a < kMaxD; ++a)
            if (a < d) s = fmaf(dl[a], __ldg(seg + (int64_t)k * d + a), s);
          const float z = q == 0 ? bflo(hw[e]) + bflo(lw[e]) : bfhi(hw[e]) + bfhi(lw[e]);
          s *= dact_f(z, act);
        }
        g[q] = s;
      }
      __nv_bfloat16 h0, l0, h1, l1;
      split1(g[0], h0, l0);
      split1(g[1], h1, l1);
      oh[e] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
      ol[e] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
    *reinterpret_cast<uint4 *>(out_hi + m * ldo + k0) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
    *reinterpret_cast<uint4 *>(out_lo + m * ldo + k0) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
  }
}

__global__ void f32_to_split_kernel(const float *__restrict__ in, int64_t rows, int cols, __nv_bfloat16 *hi,
                                    __nv_bfloat16 *lo, int64_t ld) {
  const int64_t total = rows * ld;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / ld;
    const int cc = (int)(t - r * ld);
    const float v = cc < cols ? in[r * cols + cc] : 0.f;
    __nv_bfloat16 h, l;
    split1(v, h, l);
    hi[t] = h;
    lo[t] = l;
  }
}

__global__ void set_ones_col_kernel(__nv_bfloat16 *hi, __nv_bfloat16 *lo, int64_t rows, int64_t ld, int col) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    hi[r * ld + col] = __float2bfloat16_rn(1.0f);
    lo[r * ld + col] = __float2bfloat16_rn(0.0f);
  }
}

__global__ void dot_finish_kernel(const float *__restrict__ partial, int64_t M, int slots, int d,
                                  const float *__restrict__ bias, int act, float *__restrict__ chi) {
  const int64_t total = M * d;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = t / d;
    const int a = (int)(t - m * d);
    float s = 0.f;
    for (int k = 0; k < slots; ++k) s += partial[(m * slots + k) * d + a];
    chi[t] = act_f(s + __ldg(bias + a), act);
  }
}

// ---- thin head of a training step: last Dense layer + loss + both of its backward products, one warp per row ----
// Replaces thin_forward -> loss_delta -> f32_to_split -> thin_dgrad (4 launches, z read twice) on the training path:
//   chi = lastact(z W_L + b_L);  l += sum(((chi - y) .* w)^2)  (Float64 promotion for d == 1, src/iso.jl:183-185)
//   delta_L = d(l/B)/d(pre-activation)           -> fp32 (B x d) and split bf16 (B operand of the thin weight gradient)
//   delta_{L-1} = (delta_L W_L^T) .* act'(z)      -> split bf16 rows (operand of the next two GEMMs)
// The row stays in registers between the dot product and the backward product.  The loss is reduced in a fixed
// order (lane 0 of each warp over its rows, warps in order, last block over the block partials) and packed as
// (hi, lo) floats behind the gradient vector exactly like loss_delta_kernel does.
template <int NCH, int DT>  // NCH 16-byte chunks per lane (rows of up to 256 * NCH columns); DT = d at compile time
__global__ void __launch_bounds__(256, 3)
    thin_head_kernel(const __nv_bfloat16 *__restrict__ z_hi, const __nv_bfloat16 *__restrict__ z_lo, int64_t M, int fin,
                     int64_t ldz, const float *__restrict__ seg, int d, int act, int lastact,
                     const float *__restrict__ target, const int64_t *__restrict__ idx, const float *__restrict__ wl,
                     double Bglobal, float *__restrict__ chi, float *__restrict__ delta,
                     __nv_bfloat16 *__restrict__ dl_hi, __nv_bfloat16 *__restrict__ dl_lo, int ld_dl,
                     __nv_bfloat16 *__restrict__ out_hi, __nv_bfloat16 *__restrict__ out_lo, int64_t ldo,
                     double *__restrict__ partials, unsigned int *__restrict__ ticket, float *__restrict__ packed_tail) {
  __shared__ double sh[8];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31;
  const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
  const float invB = (float)(1.0 / Bglobal);
  double lsum = 0.0;
  for (int64_t m = wid; m < M; m += nw) {
    const uint4 *ph = reinterpret_cast<const uint4 *>(z_hi + m * ldz);
    const uint4 *pl = reinterpret_cast<const uint4 *>(z_lo + m * ldz);
    // ---- chi = lastact(z W + b): the row is streamed once here and once more (from L1) for the backward product
    float acc[DT];
#pragma unroll
    for (int a = 0; a < DT; ++a) acc[a] = 0.f;
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      const int k8 = lane + 32 * j;
      if (k8 * 8 < fin) {
        const uint4 h = __ldg(ph + k8), l = __ldg(pl + k8);
        const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int k = k8 * 8 + 2 * e;
          const float z0 = bflo(hw[e]) + bflo(lw[e]);
          const float z1 = bfhi(hw[e]) + bfhi(lw[e]);
#pragma unroll
          for (int a = 0; a < DT; ++a) {
            if (k < fin) acc[a] = fmaf(z0, __ldg(seg + k * DT + a), acc[a]);
            if (k + 1 < fin) acc[a] = fmaf(z1, __ldg(seg + (k + 1) * DT + a), acc[a]);
          }
        }
      }
    }
    float dla[DT];
    const int64_t row = idx[m];
#pragma unroll
    for (int a = 0; a < DT; ++a) {
      float s = acc[a];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float c = act_f(s + __ldg(seg + fin * DT + a), lastact);
      const float r = c - __ldg(target + row * DT + a);
      float dl;
      if (DT == 1) {
        lsum += (double)r * (double)r;
        dl = (float)(2.0 * (double)r / Bglobal);
      } else {
        const float wa = __ldg(wl + a);
        const float zz = r * wa;
        lsum += (double)(zz * zz);
        dl = ((2.0f * zz) * invB) * wa;
      }
      dl *= dact_f(c, lastact);
      dla[a] = dl;
      if (lane == 0) {
        chi[m * DT + a] = c;
        delta[m * DT + a] = dl;
        __nv_bfloat16 h0, l0;
        split1(dl, h0, l0);
        dl_hi[m * ld_dl + a] = h0;
        dl_lo[m * ld_dl + a] = l0;
      }
    }
    // ---- delta_{L-1} = (delta_L W^T) .* act'(z); columns >= fin (ones column, padding) are written as 0
    uint4 *qh = reinterpret_cast<uint4 *>(out_hi + m * ldo);
    uint4 *ql = reinterpret_cast<uint4 *>(out_lo + m * ldo);
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      const int k8 = lane + 32 * j;
      if ((int64_t)k8 * 8 < ldo) {
        uint4 h = make_uint4(0, 0, 0, 0), l = h;
        if (k8 * 8 < fin) {
          h = __ldg(ph + k8);
          l = __ldg(pl + k8);
        }
        const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
        uint32_t oh[4], ol[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float g[2];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int k = k8 * 8 + 2 * e + q;
            float s = 0.f;
            if (k < fin) {
#pragma unroll
              for (int a = 0; a < DT; ++a) s = fmaf(dla[a], __ldg(seg + k * DT + a), s);
              const float z = q == 0 ? bflo(hw[e]) + bflo(lw[e]) : bfhi(hw[e]) + bfhi(lw[e]);
              s *= dact_f(z, act);
            }
            g[q] = s;
          }
          __nv_bfloat16 h0, l0, h1, l1;
          split1(g[0], h0, l0);
          split1(g[1], h1, l1);
          oh[e] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
          ol[e] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
        }
        qh[k8] = make_uint4(oh[0], oh[1], oh[2], oh[3]);
        ql[k8] = make_uint4(ol[0], ol[1], ol[2], ol[3]);
      }
    }
    for (int k8 = lane + 32 * NCH; (int64_t)k8 * 8 < ldo; k8 += 32) {  // padding past 256 * NCH columns
      qh[k8] = make_uint4(0, 0, 0, 0);
      ql[k8] = make_uint4(0, 0, 0, 0);
    }
  }
  // every lane of a warp carries the same lsum; lane 0 speaks for the warp
  if (lane == 0) sh[threadIdx.x >> 5] = lsum;
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

}  // namespace


void launch_f32_to_split(Ctx &c, const float *in, int64_t rows, int cols, __nv_bfloat16 *hi, __nv_bfloat16 *lo,
                         int64_t ld) {
  if (rows <= 0) return;
  int grid = (int)std::min<int64_t>((rows * ld + 255) / 256, (int64_t)c.num_sms * 8);
  c.timer.begin(KC_TRAIN_EW, c.stream);
  f32_to_split_kernel<<<grid, 256, 0, c.stream>>>(in, rows, cols, hi, lo, ld);
  c.timer.end(c.stream);
  IK_CUDA(cudaGetLastError());
  c.count_launch(KC_TRAIN_EW);
}

void launch_set_ones_col(Ctx &c, __nv_bfloat16 *hi, __nv_bfloat16 *lo, int64_t rows, int64_t ld, int col) {
  if (rows <= 0) return;
  int grid = (int)std::min<int64_t>((rows + 255) / 256, (int64_t)c.num_sms * 4);
  set_ones_col_kernel<<<grid, 256, 0, c.stream>>>(hi, lo, rows, ld, col);
  IK_CUDA(cudaGetLastError());
  c.count_launch(KC_TRAIN_EW);
}

void launch_dot_finish(Ctx &c, const float *partial, int64_t M, int slots, int d, const float *bias, int act,
                       float *chi) {
  if (M <= 0) return;
  int grid = (int)std::min<int64_t>((M * d + 255) / 256, (int64_t)c.num_sms * 8);
  c.timer.begin(KC_REDUCE, c.stream);
  dot_finish_kernel<<<grid, 256, 0, c.stream>>>(partial, M, slots, d, bias, act, chi);
  c.timer.end(c.stream);
  IK_CUDA(cudaGetLastError());
  c.count_launch(KC_REDUCE);
}

void launch_prep_weights(Ctx &c, const float *seg, int fin, int fout, __nv_bfloat16 *wf_hi, __nv_bfloat16 *wf_lo,
                         int64_t ld_f, __nv_bfloat16 *wd_hi, __nv_bfloat16 *wd_lo, int64_t ld_d) {
  dim3 grid(cdiv(fin, 64), cdiv(fout, 64));
  c.timer.begin(KC_TRAIN_EW, c.stream);
  prep_weights_kernel<<<grid, 256, 0, c.stream>>>(seg, fin, fout, wf_hi, wf_lo, ld_f, wd_hi, wd_lo, ld_d);
  c.timer.end(c.stream);
  IK_CUDA(cudaGetLastError());
  c.count_launch(KC_TRAIN_EW);
}

void launch_prep_weights_f16(Ctx &c, const float *seg, int fin, int fout, __nv_bfloat16 *wf16, int64_t ld_f) {
  dim3 grid(cdiv(fin, 64), cdiv(fout, 64));
  c.timer.begin(KC_TRAIN_EW, c.stream);
  prep_weights_f16_kernel<<<grid, 256, 0, c.stream>>>(seg, fin, fout, reinterpret_cast<__half *>(wf16), ld_f);
  c.timer.end(c.stream);
  IK_CUDA(cudaGetLastError());
  c.count_launch(KC_TRAIN_EW);
}

void launch_thin_forward(Ctx &c, const __nv_bfloat16 *z_hi, const __nv_bfloat16 *z_lo, int64_t M, int fin, int64_t ldz,
                         const float *seg, int d, int act, float *chi) {
  if (M <= 0) return;
  int64_t want = (M + 7) / 8;
  int grid = (int)std::min<int64_t>(want, (int64_t)c.num_sms * 8);
  c.timer.begin(KC_REDUCE, c.stream);
  thin_forward_kernel<<<grid, 256, 0, c.stream>>>(z_hi, z_lo, M, fin, ldz, seg, d, act, chi);
  c.timer.end(c.stream);
  IK_CUDA(cudaGetLastError());
  c.count_launch(KC_REDUCE);
}

void launch_thin_dgrad(Ctx &c, const float *delta, int64_t M, int d, const float *seg, int fin,
                       const __nv_bfloat16 *z_hi, const __nv_bfloat16 *z_lo, int64_t ldz, int act,
                       __nv_bfloat16 *out_hi, __nv_bfloat16 *out_lo, int64_t ldo) {
  if (M <= 0) return;
  const int64_t total = M * (ldo / 8);
  int grid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)c.num_sms * 16);
  c.timer.begin(KC_TRAIN_EW, c.stream);
  thin_dgrad_kernel<<<grid, 256, 0, c.stream>>>(delta, M, d, seg, fin, z_hi, z_lo, ldz, act, out_hi, out_lo, ldo);
  c.timer.end(c.stream);
  IK_CUDA(cudaGetLastError());
  c.count_launch(KC_TRAIN_EW);
}

bool thin_head_eligible(const Ctx &c) {
  const int fin = c.cfg.widths[c.L - 1];
  return c.d >= 1 && c.d <= 4 && fin <= 2048;
}

template <int NCH>
static void thin_head_go(Ctx &c, int grid, const __nv_bfloat16 *z_hi, const __nv_bfloat16 *z_lo, int64_t M, int fin,
                         int64_t ldz, const float *seg, const float *target, const int64_t *idx, const float *wl,
                         double Bglobal, float *chi, float *delta, __nv_bfloat16 *dl_hi, __nv_bfloat16 *dl_lo, int ld_dl,
                         __nv_bfloat16 *out_hi, __nv_bfloat16 *out_lo, int64_t ldo, double *partials,
                         unsigned int *ticket, float *packed_tail) {
  const int d = c.d, act = c.cfg.activation, lact = c.cfg.last_activation;
#define IK_HEAD(DT)                                                                                                   \
  thin_head_kernel<NCH, DT><<<grid, 256, 0, c.stream>>>(z_hi, z_lo, M, fin, ldz, seg, d, act, lact, target, idx, wl,   \
                                                        Bglobal, chi, delta, dl_hi, dl_lo, ld_dl, out_hi, out_lo, ldo, \
                                                        partials, ticket, packed_tail)
  switch (d) {
    case 1: IK_HEAD(1); break;
    case 2: IK_HEAD(2); break;
    case 3: IK_HEAD(3); break;
    default: IK_HEAD(4); break;
  }
#undef IK_HEAD
}

void launch_thin_head(Ctx &c, const __nv_bfloat16 *z_hi, const __nv_bfloat16 *z_lo, int64_t M, int fin, int64_t ldz,
                      const float *seg, const float *target, const int64_t *idx, const float *wl, double Bglobal,
                      float *chi, float *delta, __nv_bfloat16 *dl_hi, __nv_bfloat16 *dl_lo, int ld_dl,
                      __nv_bfloat16 *out_hi, __nv_bfloat16 *out_lo, int64_t ldo, double *partials,
                      unsigned int *ticket, float *packed_tail) {
  if (M <= 0) return;
  // the grid is a function of M only, so the order of the loss partials is reproducible
  const int grid = (int)std::min<int64_t>((M + 7) / 8, 1024);
  c.timer.begin(KC_TRAIN_EW, c.stream);
  if (fin <= 512)
    thin_head_go<2>(c, grid, z_hi, z_lo, M, fin, ldz, seg, target, idx, wl, Bglobal, chi, delta, dl_hi, dl_lo, ld_dl,
                    out_hi, out_lo, ldo, partials, ticket, packed_tail);
  else if (fin <= 1024)
    thin_head_go<4>(c, grid, z_hi, z_lo, M, fin, ldz, seg, target, idx, wl, Bglobal, chi, delta, dl_hi, dl_lo, ld_dl,
                    out_hi, out_lo, ldo, partials, ticket, packed_tail);
  else
    thin_head_go<8>(c, grid, z_hi, z_lo, M, fin, ldz, seg, target, idx, wl, Bglobal, chi, delta, dl_hi, dl_lo, ld_dl,
                    out_hi, out_lo, ldo, partials, ticket, packed_tail);
  c.timer.end(c.stream);
  IK_CUDA(cudaGetLastError());
  c.count_launch(KC_TRAIN_EW);
}

}  // namespace ik
