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
                                                           __nv_bfloat16 *__restrict__ out_lo) {
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
        const float v = f < F ? (sf[f] - mu) * rstd : 0.f;
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        oh[f] = h;
        ol[f] = __float2bfloat16_rn(v - __bfloat162float(h));
      }
    } else {
      float *o = out + m * ldo;
      for (int f = lane; f < F; f += 32) o[f] = (sf[f] - mu) * rstd;
    }
    __syncwarp();
  }
}

static void launch_featurize_impl(Ctx &c, const float *coords, const int64_t *gather, int64_t gather_off, int64_t M,
                                  bool pairs, bool do_ln, float *out, int64_t ldo, __nv_bfloat16 *out_hi,
                                  __nv_bfloat16 *out_lo) {
  if (M <= 0) return;
  const int D = pairs ? c.D : c.F;  // identity featurizer: records are already features
  const int F = c.F;
  const int Dp = (D + 3) & ~3, Fp = (F + 3) & ~3;
  int warps = 8;
  size_t smem = (size_t)warps * (Dp + Fp) * sizeof(float);
  while (smem > 200 * 1024 && warps > 1) {
    warps >>= 1;
    smem = (size_t)warps * (Dp + Fp) * sizeof(float);
  }
  IK_REQUIRE(smem <= 200 * 1024, ISOKANN_BAD_ARGUMENT, "feature dimension too large for the featurizer kernel");
  static bool attr_set = false;
  if (!attr_set) {
    IK_CUDA(cudaFuncSetAttribute(featurize_ln_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  int64_t want = (M + warps - 1) / warps;
  int64_t cap = (int64_t)c.num_sms * 8;
  int grid = (int)(want < cap ? want : cap);
  const float eps = c.cfg.ln_eps;
  c.timer.begin(KC_FEATURIZE, c.stream);
  featurize_ln_kernel<<<grid, warps * 32, smem, c.stream>>>(coords, gather ? gather + gather_off : nullptr, M, D, F,
                                                            pairs ? c.pairs.p : nullptr, do_ln ? 1 : 0, eps * eps,
                                                            out, ldo, Dp, Fp, out_hi, out_lo);
  c.timer.end(c.stream);
  IK_CUDA(cudaGetLastError());
  c.count_launch(KC_FEATURIZE, 4.0 * (double)(D + F) * (double)M);
}

void launch_featurize(Ctx &c, const float *coords, const int64_t *gather, int64_t gather_off, int64_t M, bool pairs,
                      bool do_ln, float *out, int64_t ldo) {
  launch_featurize_impl(c, coords, gather, gather_off, M, pairs, do_ln, out, ldo, nullptr, nullptr);
}

void launch_featurize_split(Ctx &c, const float *coords, const int64_t *gather, int64_t gather_off, int64_t M,
                            bool pairs, bool do_ln, __nv_bfloat16 *out_hi, __nv_bfloat16 *out_lo, int64_t ld) {
  launch_featurize_impl(c, coords, gather, gather_off, M, pairs, do_ln, nullptr, ld, out_hi, out_lo);
}

}  // namespace ik
