// Pairwise-distance featurizer fused with the parameter-free part of Flux.LayerNorm.
//
// Replaces flatpairdists / pdists (reference src/utils/pairdists.jl:6-24,109-127; the CUDA.jl
// kernel :137-150 plus the gather/max/sqrt broadcasts :19-22) and normalise(x; dims=1)
// of Flux.LayerNorm (src/models.jl:90).  One warp owns one coordinate record: the record is
// staged in shared memory with coalesced loads, every lane evaluates a strided subset of the
// feature list (a table of coordinate-offset pairs, so FeaturesAll / FeaturesAtoms /
// FeaturesPairs are one code path), the features are staged in shared memory for the two-pass
// mean/variance, and the normalised row is written back coalesced.
//
// Algorithmic HBM traffic per record: 4*D read + 4*F written.
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc.cuh"

namespace ik {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// gather: optional 0-based record ids (minibatch slice of the permutation), nullptr = identity
__global__ void __launch_bounds__(256) featurize_ln_kernel(const float *__restrict__ coords,
                                                           const int64_t *__restrict__ gather, int64_t M, int D,
                                                           int F, const int2 *__restrict__ pairs, int do_ln,
                                                           float eps2, float *__restrict__ out, int64_t ldo,
                                                           int Dp, int Fp, __nv_bfloat16 *__restrict__ out_hi,
                                                           __nv_bfloat16 *__restrict__ out_lo, int fmt) {
  extern __shared__ float sm[];
  const int warps = blockDim.x >> 5;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float *sc = sm + (size_t)w * (Dp + Fp);
  float *sf = sc + Dp;
  for (int64_t m = (int64_t)blockIdx.x * warps + w; m < M; m += (int64_t)gridDim.x * warps) {
    const int64_t src = gather ? gather[m] : m;
    const float *c = coords + src * D;
    for (int i = lane; i < D; i += 32) sc[i] = __ldg(c + i);
    __syncwarp();
    float s = 0.f;
    if (pairs) {
      for (int f = lane; f < F; f += 32) {
        const int2 p = __ldg(pairs + f);
        const float dx = sc[p.x] - sc[p.y];
        const float dy = sc[p.x + 1] - sc[p.y + 1];
        const float dz = sc[p.x + 2] - sc[p.y + 2];
        const float v = sqrtf(fmaxf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)), 0.f));
        sf[f] = v;
        s += v;
      }
    } else {
      for (int f = lane; f < F; f += 32) {
        const float v = sc[f];
        sf[f] = v;
        s += v;
      }
    }
    float mu = 0.f, rstd = 1.f;
    if (do_ln) {
      mu = warp_sum(s) / (float)F;
      float q = 0.f;
      for (int f = lane; f < F; f += 32) {
        const float t = sf[f] - mu;
        q = fmaf(t, t, q);
      }
      const float var = warp_sum(q) / (float)F;
      rstd = 1.0f / sqrtf(var + eps2);
    }
    if (out_hi) {  // bf16 (hi, lo) split for the tensor-core path; pad columns [F, ldo) are zeroed
      __nv_bfloat16 *oh = out_hi + m * ldo, *ol = out_lo + m * ldo;
      for (int f = lane; f < (int)ldo; f += 32) {
        const float v = f < F ? (sf[f] - mu) * rstd : (f == F ? 1.f : 0.f);  // column F = 1: bias column
        if (fmt) {  // fp16 pairs (2-MMA inference forward)
          const __half h = __float2half_rn(v);
          reinterpret_cast<__half *>(oh)[f] = h;
          reinterpret_cast<__half *>(ol)[f] = __float2half_rn(v - __half2float(h));
        } else {
          const __nv_bfloat16 h = __float2bfloat16_rn(v);
          oh[f] = h;
          ol[f] = __float2bfloat16_rn(v - __bfloat162float(h));
        }
      }
    } else {
      float *o = out + m * ldo;
      for (int f = lane; f < F; f += 32) o[f] = (sf[f] - mu) * rstd;
    }
    __syncwarp();
  }
}


// Register-resident variant for F <= 32*NF.  Per lane the NF features AND their (packed) pair-table
// entries live in registers; a coordinate record is staged in shared memory as one float4 per atom
// so a distance needs two LDS.128, not six LDS.32; distances use rsqrt (2 ulp; the tolerance on
// distances is 1e-5 relative); stores use immediate offsets from one row pointer.  The first ncu
// capture of the shared-memory-staged kernel showed it issue-bound (1090 warp instructions per
// record, 70-78 % issue-slot utilisation): this layout needs ~2.5x fewer.  SPLIT writes the bf16
// (hi, lo) operand of the tensor-core path, two adjacent features per lane so every store
// instruction covers 128 contiguous bytes.
template <int NF, bool SPLIT>
__global__ void __launch_bounds__(256, (NF <= 20 ? 3 : 2)) featurize_reg_kernel(const float *__restrict__ coords,
                                                            const int64_t *__restrict__ gather, int64_t M, int D,
                                                            int F, const int2 *__restrict__ pairs, int do_ln,
                                                            float eps2, float *__restrict__ out, int64_t ldo,
                                                            int atoms_pad, __nv_bfloat16 *__restrict__ out_hi,
                                                            __nv_bfloat16 *__restrict__ out_lo, int fmt) {
  extern __shared__ float4 sm4[];
  const int warps = blockDim.x >> 5;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4 *sc = sm4 + (size_t)w * atoms_pad;
  float *scf = reinterpret_cast<float *>(sc);
  const bool have_pairs = pairs != nullptr;
  // packed absolute shared-memory addresses of the two atoms (16 B per atom) of each of this lane's
  // features; slots beyond F point at atom 0 twice (distance 0)
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sc);
  uint32_t tab[NF];
  int ninv = 0;
#pragma unroll
  for (int t = 0; t < NF; ++t) {
    const int f = SPLIT ? (2 * lane + 64 * (t >> 1) + (t & 1)) : (lane + 32 * t);
    uint32_t a = sbase, b = sbase;
    if (f >= F) ++ninv;
    if (have_pairs && f < F) {
      const int2 p = __ldg(pairs + f);  // coordinate offsets 3a, 3b
      a = sbase + (uint32_t)(p.x / 3) * 16u;
      b = sbase + (uint32_t)(p.y / 3) * 16u;
    }
    tab[t] = a | (b << 16);
  }
  const float finv = (float)ninv;
  const float invF = 1.0f / (float)F;
  // staging slots of this lane's coordinates (x,y,z of atom a -> float4 slot a), loop-invariant
  constexpr int KC = 8;  // covers D <= 256 without a runtime loop
  int dst[KC];
#pragma unroll
  for (int k = 0; k < KC; ++k) {
    const int i = lane + 32 * k;
    dst[k] = 4 * (i / 3) + (i % 3);
  }
  for (int64_t m = (int64_t)blockIdx.x * warps + w; m < M; m += (int64_t)gridDim.x * warps) {
    const int64_t src = gather ? gather[m] : m;
    const float *c = coords + src * D;
    float v[NF];
    float s = 0.f;
    if (have_pairs) {
#pragma unroll
      for (int k = 0; k < KC; ++k)
        if (lane + 32 * k < D) scf[dst[k]] = __ldg(c + lane + 32 * k);
      for (int i = lane + 32 * KC; i < D; i += 32) {
        const int a = i / 3;
        scf[4 * a + (i - 3 * a)] = __ldg(c + i);
      }
      __syncwarp();
#pragma unroll
      for (int t = 0; t < NF; ++t) {
        float4 pa, pb;
        uint32_t aa, ab;
        // volatile: keep the 2-instruction decode inside the record loop instead of 2*NF hoisted
        // address registers (register pressure decides the occupancy of this kernel)
        asm volatile("and.b32 %0, %1, 0xFFFF;" : "=r"(aa) : "r"(tab[t]));
        asm volatile("shr.u32 %0, %1, 16;" : "=r"(ab) : "r"(tab[t]));
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(pa.x), "=f"(pa.y), "=f"(pa.z), "=f"(pa.w)
                     : "r"(aa));
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(pb.x), "=f"(pb.y), "=f"(pb.z), "=f"(pb.w)
                     : "r"(ab));
        const float dx = pa.x - pb.x, dy = pa.y - pb.y, dz = pa.z - pb.z;
        const float sq = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
        const float val = sq > 0.f ? sq * rsqrtf(sq) : 0.f;  // features beyond F read atom 0 twice -> 0
        v[t] = val;
        s += val;
      }
    } else {
#pragma unroll
      for (int t = 0; t < NF; ++t) {
        const int f = SPLIT ? (2 * lane + 64 * (t >> 1) + (t & 1)) : (lane + 32 * t);
        const float val = f < F ? __ldg(c + f) : 0.f;
        v[t] = val;
        s += val;
      }
    }
    float scale = 1.f, shift = 0.f;  // xhat = v * scale + shift
    if (do_ln) {
      const float mu = warp_sum(s) * invF;
      float q = 0.f;
#pragma unroll
      for (int t = 0; t < NF; ++t) {
        const float d = v[t] - mu;
        q = fmaf(d, d, q);
      }
      q = fmaf(-finv * mu, mu, q);  // slots beyond F hold 0 and contributed mu^2 each
      scale = rsqrtf(fmaxf(warp_sum(q), 0.f) * invF + eps2);
      shift = -mu * scale;
    }
    if (SPLIT) {
      __nv_bfloat16 *oh = out_hi + m * ldo + 2 * lane, *ol = out_lo + m * ldo + 2 * lane;
#pragma unroll
      for (int t = 0; t < NF; t += 2) {
        const int f = 2 * lane + 64 * (t >> 1);
        if (f < (int)ldo) {
          // column F carries the constant 1 that turns the weight-gradient GEMM into [x_hat, 1]^T * delta
          const float x0 = f < F ? fmaf(v[t], scale, shift) : (f == F ? 1.f : 0.f);
          const float x1 = f + 1 < F ? fmaf(v[t + 1], scale, shift) : (f + 1 == F ? 1.f : 0.f);
          if (fmt) {  // fp16 pairs
            const __half2 h2 = __floats2half2_rn(x0, x1);
            const float2 hf = __half22float2(h2);
            *reinterpret_cast<__half2 *>(oh + 64 * (t >> 1)) = h2;
            *reinterpret_cast<__half2 *>(ol + 64 * (t >> 1)) = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
          } else {
            const __nv_bfloat162 h2 = __floats2bfloat162_rn(x0, x1);
            const uint32_t hw = *reinterpret_cast<const uint32_t *>(&h2);
            const __nv_bfloat162 l2 =
                __floats2bfloat162_rn(x0 - __uint_as_float(hw << 16), x1 - __uint_as_float(hw & 0xFFFF0000u));
            *reinterpret_cast<__nv_bfloat162 *>(oh + 64 * (t >> 1)) = h2;
            *reinterpret_cast<__nv_bfloat162 *>(ol + 64 * (t >> 1)) = l2;
          }
        }
      }
    } else {
      float *o = out + m * ldo + lane;
#pragma unroll
      for (int t = 0; t < NF; ++t)
        if (lane + 32 * t < F) o[32 * t] = fmaf(v[t], scale, shift);
    }
    __syncwarp();
  }
}

template <int NF, bool SPLIT>
static void launch_reg(Ctx &c, const float *coords, const int64_t *gather, int64_t M, int D, int F, bool pairs,
                       bool do_ln, float *out, int64_t ldo, __nv_bfloat16 *out_hi, __nv_bfloat16 *out_lo) {
  const int atoms_pad = pairs ? (D + 2) / 3 : 0;
  const int warps = 8;
  const size_t smem = (size_t)warps * atoms_pad * sizeof(float4);
  int64_t want = (M + warps - 1) / warps;
  int64_t cap = (int64_t)c.num_sms * 8;
  int grid = (int)(want < cap ? want : cap);
  const float eps = c.cfg.ln_eps;
  featurize_reg_kernel<NF, SPLIT><<<grid, warps * 32, smem, c.stream>>>(
      coords, gather, M, D, F, pairs ? c.pairs.p : nullptr, do_ln ? 1 : 0, eps * eps, out, ldo, atoms_pad, out_hi,
      out_lo, c.split_fmt);
}

// Pullback of featurizer + LayerNorm (reference: sqpairdist_bwd_kernel! and its rrule,
// src/utils/pairdists.jl:153-167,179-196, used by dchidx src/utils/minimumpath.jl:3-7 and the metadynamics
// bias src/simulators/metadynamics.jl:40-49).  One warp per record: recompute the distances and the
// LayerNorm statistics, turn dL/dx_hat into dL/df (LayerNorm backward), then every lane owns atoms and
// walks their incident features (CSR built once per context) -- no atomics, deterministic.
__global__ void __launch_bounds__(256) featurize_backward_kernel(const float *__restrict__ in, int64_t M, int D, int F,
                                                                 const int2 *__restrict__ pairs,
                                                                 const int *__restrict__ adj_off,
                                                                 const int2 *__restrict__ adj, int do_ln, float eps2,
                                                                 const float *__restrict__ gxhat,
                                                                 float *__restrict__ out, int Dp, int Fp) {
  extern __shared__ float sm[];
  const int warps = blockDim.x >> 5;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float *sc = sm + (size_t)w * (Dp + 2 * Fp);
  float *sf = sc + Dp;   // features, then x_hat
  float *sg = sf + Fp;   // dL/dx_hat, then dL/df (/ f for distance features)
  for (int64_t m = (int64_t)blockIdx.x * warps + w; m < M; m += (int64_t)gridDim.x * warps) {
    const float *c = in + m * D;
    for (int i = lane; i < D; i += 32) sc[i] = __ldg(c + i);
    __syncwarp();
    float s = 0.f;
    for (int f = lane; f < F; f += 32) {
      float v;
      if (pairs) {
        const int2 p = __ldg(pairs + f);
        const float dx = sc[p.x] - sc[p.y], dy = sc[p.x + 1] - sc[p.y + 1], dz = sc[p.x + 2] - sc[p.y + 2];
        v = sqrtf(fmaxf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)), 0.f));
      } else {
        v = sc[f];
      }
      sf[f] = v;
      sg[f] = __ldg(gxhat + m * F + f);
      s += v;
    }
    if (do_ln) {
      const float mu = warp_sum(s) / (float)F;
      float q = 0.f;
      for (int f = lane; f < F; f += 32) {
        const float t = sf[f] - mu;
        q = fmaf(t, t, q);
      }
      const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)F + eps2);
      float s1 = 0.f, s2 = 0.f;
      for (int f = lane; f < F; f += 32) {
        const float xh = (sf[f] - mu) * rstd;
        s1 += sg[f];
        s2 = fmaf(sg[f], xh, s2);
      }
      s1 = warp_sum(s1) / (float)F;
      s2 = warp_sum(s2) / (float)F;
      for (int f = lane; f < F; f += 32) {
        const float xh = (sf[f] - mu) * rstd;
        sg[f] = rstd * (sg[f] - s1 - xh * s2);   // dL/df
      }
    }
    if (!pairs) {
      for (int f = lane; f < F; f += 32) out[m * D + f] = sg[f];
    } else {
      for (int f = lane; f < F; f += 32) sg[f] = sf[f] > 0.f ? sg[f] / sf[f] : 0.f;
      __syncwarp();
      const int A = D / 3;
      for (int a = lane; a < A; a += 32) {
        float gx = 0.f, gy = 0.f, gz = 0.f;
        const float ax = sc[3 * a], ay = sc[3 * a + 1], az = sc[3 * a + 2];
        for (int e = adj_off[a]; e < adj_off[a + 1]; ++e) {
          const int2 q = __ldg(adj + e);
          const float wv = sg[q.x];
          gx = fmaf(wv, ax - sc[q.y], gx);
          gy = fmaf(wv, ay - sc[q.y + 1], gy);
          gz = fmaf(wv, az - sc[q.y + 2], gz);
        }
        out[m * D + 3 * a] = gx;
        out[m * D + 3 * a + 1] = gy;
        out[m * D + 3 * a + 2] = gz;
      }
    }
    __syncwarp();
  }
}

void launch_featurize_backward(Ctx &c, const float *in, int64_t M, bool pairs, bool do_ln, const float *gxhat,
                               float *out) {
  if (M <= 0) return;
  const int D = pairs ? c.D : c.F;
  const int F = c.F;
  const int Dp = (D + 3) & ~3, Fp = (F + 3) & ~3;
  int warps = 8;
  size_t smem = (size_t)warps * (Dp + 2 * Fp) * sizeof(float);
  while (smem > 200 * 1024 && warps > 1) {
    warps >>= 1;
    smem = (size_t)warps * (Dp + 2 * Fp) * sizeof(float);
  }
  IK_REQUIRE(smem <= 200 * 1024, ISOKANN_BAD_ARGUMENT, "feature dimension too large for the featurizer pullback");
  if (c.attr_needed(Ctx::ATTR_FEAT_BWD))
    IK_CUDA(cudaFuncSetAttribute(featurize_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  int64_t want = (M + warps - 1) / warps;
  int grid = (int)std::min<int64_t>(want, (int64_t)c.num_sms * 8);
  const float eps = c.cfg.ln_eps;
  c.timer.begin(KC_FEATURIZE, c.stream);
  featurize_backward_kernel<<<grid, warps * 32, smem, c.stream>>>(in, M, D, F, pairs ? c.pairs.p : nullptr,
                                                                  pairs ? c.adj_off.p : nullptr,
                                                                  pairs ? c.adj.p : nullptr, do_ln ? 1 : 0, eps * eps,
                                                                  gxhat, out, Dp, Fp);
  c.timer.end(c.stream);
  IK_CUDA(cudaGetLastError());
  c.count_launch(KC_FEATURIZE);
}

static void launch_featurize_impl(Ctx &c, const float *coords, const int64_t *gather, int64_t gather_off, int64_t M,
                                  bool pairs, bool do_ln, float *out, int64_t ldo, __nv_bfloat16 *out_hi,
                                  __nv_bfloat16 *out_lo) {
  if (M <= 0) return;
  const int D = pairs ? c.D : c.F;  // identity featurizer: records are already features
  const int F = c.F;
  if (pairs && featurize_rec_applicable(c, out_hi != nullptr, ldo)) {
    c.timer.begin(KC_FEATURIZE, c.stream);
    launch_featurize_rec(c, coords, gather ? gather + gather_off : nullptr, M, do_ln, out, out_hi, out_lo, ldo);
    c.timer.end(c.stream);
    IK_CUDA(cudaGetLastError());
    c.count_launch(KC_FEATURIZE, 4.0 * (double)(D + F) * (double)M);
    return;
  }
  {
    // register-resident kernels: F (fp32 output) or the padded row (split output) must fit 32*NF, coordinate
    // offsets must fit the packed 16-bit table, 8 coordinate records must fit shared memory
    const int64_t need = out_hi ? ldo : F;
    const bool small = D < 3 * 4096 && ((D + 2) / 3) * 16 * 8 <= 48 * 1024;
    const int64_t *gp = gather ? gather + gather_off : nullptr;
    if (small && need <= 1024) {
      c.timer.begin(KC_FEATURIZE, c.stream);
      if (out_hi) {
        if (need <= 64) launch_reg<2, true>(c, coords, gp, M, D, F, pairs, do_ln, out, ldo, out_hi, out_lo);
        else if (need <= 256) launch_reg<8, true>(c, coords, gp, M, D, F, pairs, do_ln, out, ldo, out_hi, out_lo);
        else if (need <= 640) launch_reg<20, true>(c, coords, gp, M, D, F, pairs, do_ln, out, ldo, out_hi, out_lo);
        else launch_reg<32, true>(c, coords, gp, M, D, F, pairs, do_ln, out, ldo, out_hi, out_lo);
      } else {
        if (need <= 64) launch_reg<2, false>(c, coords, gp, M, D, F, pairs, do_ln, out, ldo, out_hi, out_lo);
        else if (need <= 256) launch_reg<8, false>(c, coords, gp, M, D, F, pairs, do_ln, out, ldo, out_hi, out_lo);
        else if (need <= 640) launch_reg<20, false>(c, coords, gp, M, D, F, pairs, do_ln, out, ldo, out_hi, out_lo);
        else launch_reg<32, false>(c, coords, gp, M, D, F, pairs, do_ln, out, ldo, out_hi, out_lo);
      }
      c.timer.end(c.stream);
      IK_CUDA(cudaGetLastError());
      c.count_launch(KC_FEATURIZE, 4.0 * (double)(D + F) * (double)M);
      return;
    }
  }
  const int Dp = (D + 3) & ~3, Fp = (F + 3) & ~3;
  int warps = 8;
  size_t smem = (size_t)warps * (Dp + Fp) * sizeof(float);
  while (smem > 200 * 1024 && warps > 1) {
    warps >>= 1;
    smem = (size_t)warps * (Dp + Fp) * sizeof(float);
  }
  IK_REQUIRE(smem <= 200 * 1024, ISOKANN_BAD_ARGUMENT, "feature dimension too large for the featurizer kernel");
  if (c.attr_needed(Ctx::ATTR_FEAT_LN))
    IK_CUDA(cudaFuncSetAttribute(featurize_ln_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  int64_t want = (M + warps - 1) / warps;
  int64_t cap = (int64_t)c.num_sms * 8;
  int grid = (int)(want < cap ? want : cap);
  const float eps = c.cfg.ln_eps;
  c.timer.begin(KC_FEATURIZE, c.stream);
  featurize_ln_kernel<<<grid, warps * 32, smem, c.stream>>>(coords, gather ? gather + gather_off : nullptr, M, D, F,
                                                            pairs ? c.pairs.p : nullptr, do_ln ? 1 : 0, eps * eps,
                                                            out, ldo, Dp, Fp, out_hi, out_lo, c.split_fmt);
  c.timer.end(c.stream);
  IK_CUDA(cudaGetLastError());
  c.count_launch(KC_FEATURIZE, 4.0 * (double)(D + F) * (double)M);
}

void launch_featurize(Ctx &c, const float *coords, const int64_t *gather, int64_t gather_off, int64_t M, bool pairs,
                      bool do_ln, float *out, int64_t ldo) {
  launch_featurize_impl(c, coords, gather, gather_off, M, pairs, do_ln, out, ldo, nullptr, nullptr);
}

void launch_featurize_split(Ctx &c, const float *coords, const int64_t *gather, int64_t gather_off, int64_t M,
                            bool pairs, bool do_ln, __nv_bfloat16 *out_hi, __nv_bfloat16 *out_lo, int64_t ld, int fmt) {
  struct Reset {
    int &f;
    ~Reset() { f = 0; }
  } reset{c.split_fmt};
  c.split_fmt = fmt;  // read by the kernel launchers below
  launch_featurize_impl(c, coords, gather, gather_off, M, pairs, do_ln, nullptr, ld, out_hi, out_lo);
}

}  // namespace ik
