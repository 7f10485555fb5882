// Pairwise-distance featurizer + LayerNorm, "lane = record" form.  Reference semantics: flatpairdists
// (src/utils/pairdists.jl:6-24, halfinds :50-56) over the coordinates of SimulationData, followed by the
// LayerNorm statistics of the pairnet's first layer (src/models.jl:65-69; Flux normalise:
// (x - mean) / sqrt(var_biased + eps^2)).
//
// Why a second kernel: featurize_reg_kernel (featurize.cu) maps lanes to FEATURES of one record, so every
// feature costs two 16-byte shared-memory gathers (8 wavefronts per 32 features) and ~34 issued instructions;
// ncu shows it bound by LSU wavefronts and issue slots at 25-36 % of the HBM roofline.  Here a block owns 32
// RECORDS and every lane of every warp is one record:
//   load    the coordinates of the 32 records are staged transposed ([coordinate][record], pitch 33 words, so
//           both the transposing stores and the per-lane reads are conflict-free);
//   pass 1  the warps split the strict upper triangle in tiles of two columns (j, j+1); the two column atoms
//           live in registers and one LDS.32 per coordinate of atom i serves both columns: ~12 instructions
//           per pair for 32 records.  The raw distances are parked in shared memory ([feature][record]) and
//           each warp keeps a partial sum / sum of squares per record;
//   pass 2  after one block barrier every warp folds the partials into (scale, shift) per record, then the
//           warps read the parked distances back transposed (8 features of one record per lane, record
//           rotation -> conflict-free), normalise, split into bf16 (hi, lo) and write 16-byte pieces that
//           cover whole 128-byte lines of the row-major x_hat rows (or plain fp32 rows for isokann_featurize).
// No distance is computed twice and no value leaves the SM before it is final.
#include <cuda_fp16.h>

#include "common.cuh"

namespace ik {
namespace {

constexpr int RP = 33;  // words per staged row: 32 records + 1 pad word

__device__ __forceinline__ float dist3(float ax, float ay, float az, float bx, float by, float bz, float &sq) {
  const float dx = ax - bx, dy = ay - by, dz = az - bz;
  sq = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
  float d;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(sq));  // one MUFU; sqrt(0) = 0, no special case
  return d;
}

// Columns j0 .. j0+JT-1 of the strict upper triangle against all atoms i < j.  at = staged atoms + lane,
// dt = parked-distance row of feature (0, j0) + lane.  Feature (i, j0+k) sits k*j0 + k(k-1)/2 + i rows after dt.
// s, q accumulate the LayerNorm sums about the pivot piv (the record's first distance): sum(d - piv) and
// sum((d - piv)^2).  The plain single-pass form E[d^2] - mu^2 cancels for large offsets / nearly uniform distances.
template <int JT, int SC>  // SC: words between consecutive coordinates of one record in the staged block
__device__ __forceinline__ void sweep_tile(const float *at, int j0, float piv, float &s, float &q, float *dt) {
  float cx[JT], cy[JT], cz[JT];
#pragma unroll
  for (int k = 0; k < JT; ++k) {
    const float *p = at + 3 * (j0 + k) * SC;
    cx[k] = p[0];
    cy[k] = p[SC];
    cz[k] = p[2 * SC];
  }
  float s2[JT], q2[JT];
#pragma unroll
  for (int k = 0; k < JT; ++k) s2[k] = q2[k] = 0.f;
  // atoms i+1 and i+2 are loaded while atom i is being used (reads one or two atoms past j0 - 1 stay inside the
  // staged block: atom j0 exists, and the row after the last atom is the start of the parked distances)
  const float *ap = at;
  float *dp = dt;
  float ax = ap[0], ay = ap[SC], az = ap[2 * SC];
  float bx = ap[3 * SC], by = ap[4 * SC], bz = ap[5 * SC];
#pragma unroll 4
  for (int i = 0; i < j0; ++i) {
    ap += 3 * SC;
    const float nx = ap[3 * SC], ny = ap[4 * SC], nz = ap[5 * SC];
#pragma unroll
    for (int k = 0; k < JT; ++k) {
      float sq;
      const float d = dist3(ax, ay, az, cx[k], cy[k], cz[k], sq);
      const float e = d - piv;
      s2[k] += e;
      q2[k] = fmaf(e, e, q2[k]);
      dp[(k * j0 + (k * (k - 1)) / 2) * RP] = d;
    }
    ax = bx, ay = by, az = bz;
    bx = nx, by = ny, bz = nz;
    dp += RP;
  }
  // the triangle inside the tile: atom i = j0 + t is column atom t, already in registers
#pragma unroll
  for (int t = 0; t + 1 < JT; ++t) {
#pragma unroll
    for (int k = t + 1; k < JT; ++k) {
      float sq;
      const float d = dist3(cx[t], cy[t], cz[t], cx[k], cy[k], cz[k], sq);
      const float e = d - piv;
      s2[k] += e;
      q2[k] = fmaf(e, e, q2[k]);
      dt[(k * j0 + (k * (k - 1)) / 2 + j0 + t) * RP] = d;
    }
  }
#pragma unroll
  for (int k = 0; k < JT; ++k) {
    s += s2[k];
    q += q2[k];
  }
}

// two values -> (bf16x2 hi, bf16x2 lo) with hi = rn(x), lo = rn(x - hi): the split of featurize_reg_kernel
__device__ __forceinline__ void split2(float x0, float x1, uint32_t &h, uint32_t &l) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(x1), "f"(x0));
  const float r0 = x0 - __uint_as_float(h << 16), r1 = x1 - __uint_as_float(h & 0xFFFF0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(r1), "f"(r0));
}

// fp16 pairs (2-MMA inference forward): hi = rn_f16(x), lo = rn_f16(x - hi)
__device__ __forceinline__ void split2h(float x0, float x1, uint32_t &h, uint32_t &l) {
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(x1), "f"(x0));
  const float2 hf = __half22float2(*reinterpret_cast<const __half2 *>(&h));
  const float r0 = x0 - hf.x, r1 = x1 - hf.y;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(r1), "f"(r0));
}

template <bool SPLIT, bool HASMAP>
__global__ void __launch_bounds__(256, 2)
    featurize_blk_kernel(const float *__restrict__ coords, const int64_t *__restrict__ gather, int64_t M, int D, int A,
                         const int *__restrict__ cmap, int F, int do_ln, float eps2, float *__restrict__ out,
                         __nv_bfloat16 *__restrict__ out_hi, __nv_bfloat16 *__restrict__ out_lo, int64_t ldo, int fmt) {
  extern __shared__ __align__(16) float smem_f[];
  const int W = blockDim.x >> 5;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = 3 * A;
  float *atoms = smem_f;                                       // [C][RP]
  float *dist = atoms + C * RP;                                // [F][RP]
  const int Fp = (F + 64) & ~63;                               // rows incl. the padding pass 2 reads through
  float2 *part = reinterpret_cast<float2 *>(smem_f + ((((C + Fp) * RP) + 1) & ~1));  // [W][32] partial (sum, sum sq)
  float2 *stat = part + W * 32;                                // [W][32] (scale, shift), one private copy per warp
  const float *at = atoms + lane;
  const int64_t nblk = (M + 31) >> 5;
  const float invF = 1.0f / (float)F;
  const int T = A >> 1;  // tiles of two columns: j0 = 1, 3, ...; the last one has a single column if A is even
  // Staging of one block of 32 records: 4-byte cp.async copies straight into the transposed layout (no registers,
  // one wait).  Two alternatives were measured and dropped (profiles/r01_ncu_featurize_blk_v6.md): issuing the
  // copies of block i+1 under pass 2 of block i (slower: they lengthen the store-bound pass 2, and the second
  // resident block already covers the latency), and one cp.async.bulk per block into an untransposed layout
  // (10 instead of 86 instructions per record, same time: the kernel is bound by MIO/LSU latency, not by issue).
  const uint32_t atoms_s = (uint32_t)__cvta_generic_to_shared(atoms);
  const int nfull = C >> 5, ctail = lane + (nfull << 5);
  auto issue = [&](int64_t blk_) {
    const int64_t m0_ = blk_ << 5;
    const int nv = (int)(M - m0_ < 32 ? M - m0_ : 32);
    for (int rec = w; rec < 32; rec += W) {
      const int64_t m = m0_ + (rec < nv ? rec : nv - 1);
      const float *r = coords + (gather ? __ldg(gather + m) : m) * D;
      uint32_t sa = atoms_s + 4u * (uint32_t)(lane * RP + rec);
      if (HASMAP) {
        for (int cc = lane; cc < C; cc += 32) {
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(r + __ldg(cmap + cc)) : "memory");
          sa += 4u * 32u * RP;
        }
      } else {
        const float *g = r + lane;
#pragma unroll 4
        for (int t = 0; t < nfull; ++t) {
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(g) : "memory");
          g += 32;
          sa += 4u * 32u * RP;
        }
        if (ctail < C) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(g) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  for (int64_t blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    const int64_t m0 = blk << 5;
    const int nvalid = (int)(M - m0 < 32 ? M - m0 : 32);
    issue(blk);
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();  // atoms complete; every warp has left pass 2 of the previous block
    // pivot of the LayerNorm sums: the record's first distance (atoms 0 and 1); the same in every warp
    float piv;
    {
      float sq;
      piv = dist3(at[0], at[RP], at[2 * RP], at[3 * RP], at[4 * RP], at[5 * RP], sq);
    }
    // ---- pass 1: distances -> shared memory, partial LayerNorm sums per warp
    {
      float s = 0.f, q = 0.f;
      // tiles in descending size, dealt boustrophedon (0..W-1, W-1..0, ...) so the warps carry equal pair counts
      for (int base = 0; base < T; base += 2 * W) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int u = base + (h == 0 ? w : 2 * W - 1 - w);
          if (u < T) {
            const int j0 = 1 + 2 * (T - 1 - u);
            float *dt = dist + ((j0 * (j0 - 1)) >> 1) * RP + lane;
            if (j0 + 1 < A) sweep_tile<2, RP>(at, j0, piv, s, q, dt);
            else sweep_tile<1, RP>(at, j0, piv, s, q, dt);
          }
        }
      }
      part[w * 32 + lane] = make_float2(s, q);
    }
    __syncthreads();  // distances and partial sums complete
    // ---- LayerNorm scale / shift per record (every warp folds the partials itself: no second barrier)
    {
      float scale = 1.f, shift = 0.f;
      if (do_ln) {
        float s = 0.f, q = 0.f;
        for (int v = 0; v < W; ++v) {
          const float2 p = part[v * 32 + lane];
          s += p.x;
          q += p.y;
        }
        const float me = s * invF;  // mean of (d - pivot)
        scale = rsqrtf(fmaxf(fmaf(-me, me, q * invF), 0.f) + eps2);
        shift = -(piv + me) * scale;
      }
      stat[w * 32 + lane] = make_float2(scale, shift);
      __syncwarp();
    }
    const float2 *st = stat + w * 32;
    // ---- pass 2: read back transposed, normalise, (split,) store
    if (SPLIT) {
      // lane = (8-feature group q8 of an octet, record rsub of a rotation): per instruction 4 records x 64
      // features, i.e. 4 x 128 contiguous bytes of hi and of lo.  A warp keeps its rotation(s) `it`, so the
      // record, its (scale, shift) and the row pointers are loop-invariant; only the octet advances.
      const int q8 = lane & 7, rsub = lane >> 3;
      const int noct = (int)(ldo >> 6), full = F >> 6;
      for (int it = w; it < 8; it += W) {
        const int rec = (4 * it - 4 * q8 + rsub) & 31;  // bank of dist[(f0+k)*RP + rec] = 4 q8 + rsub + const
        const float2 ss = st[rec];
        // rows past M hold copies of the last valid record (see the staging), so they are written, unpredicated,
        // over that record's row with identical bytes
        const int orec = rec < nvalid ? rec : nvalid - 1;
        const float *p = dist + (q8 << 3) * RP + rec;
        __nv_bfloat16 *oh = out_hi + (m0 + orec) * ldo + (q8 << 3), *ol = out_lo + (m0 + orec) * ldo + (q8 << 3);
        // Two register sets alternate: the raw distances of octet i+1 are loaded before octet i is converted and
        // stored, so the loads never wait for registers that a 16-byte store still has to read, and their latency
        // hides behind the conversion.
        auto load8 = [&](float (&r)[8], const float *pp) {
#pragma unroll
          for (int k = 0; k < 8; ++k) r[k] = pp[k * RP];
        };
        auto emit = [&](const float (&r)[8], int oct, __nv_bfloat16 *ph, __nv_bfloat16 *pl) {
          float v[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = fmaf(r[k], ss.x, ss.y);
          if (oct >= full) {
            // column F carries the constant 1 of the augmented weight-gradient GEMM, the rest of the padded row is
            // 0 (the rows read past F belong to the padding of the parked distances and are discarded here)
            const int f0 = (oct << 6) + (q8 << 3);
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = f0 + k < F ? v[k] : (f0 + k == F ? 1.f : 0.f);
          }
          uint4 h, l;
          if (fmt) {
            split2h(v[0], v[1], h.x, l.x);
            split2h(v[2], v[3], h.y, l.y);
            split2h(v[4], v[5], h.z, l.z);
            split2h(v[6], v[7], h.w, l.w);
          } else {
            split2(v[0], v[1], h.x, l.x);
            split2(v[2], v[3], h.y, l.y);
            split2(v[4], v[5], h.z, l.z);
            split2(v[6], v[7], h.w, l.w);
          }
          *reinterpret_cast<uint4 *>(ph) = h;
          *reinterpret_cast<uint4 *>(pl) = l;
        };
        float ra[8], rb[8];
        load8(ra, p);
        for (int oct = 0; oct < noct; oct += 2) {
          if (oct + 1 < noct) load8(rb, p + 64 * RP);
          emit(ra, oct, oh, ol);
          if (oct + 2 < noct) load8(ra, p + 128 * RP);
          if (oct + 1 < noct) emit(rb, oct + 1, oh + 64, ol + 64);
          p += 128 * RP;
          oh += 128;
          ol += 128;
        }
      }
    } else {
      // one record row per warp-iteration: 32 consecutive features per instruction (128 contiguous bytes)
      for (int rec = w; rec < nvalid; rec += W) {
        const float2 ss = st[rec];
        const float *p = dist + lane * RP + rec;
        float *o = out + (m0 + rec) * ldo + lane;
#pragma unroll 4
        for (int f = lane; f < F; f += 32) {
          *o = fmaf(*p, ss.x, ss.y);
          p += 32 * RP;
          o += 32;
        }
      }
    }
  }
}

struct RecPlan {
  int warps;
  size_t smem;
  int blocks_per_sm;
};

RecPlan rec_plan(int A) {
  const int T = A / 2, F = A * (A - 1) / 2;
  int W = 1;
  while (W < 8 && 4 * W <= T) W *= 2;  // at least two tiles per warp
  RecPlan p;
  p.warps = W;
  p.smem = (size_t)(3 * A + ((F + 64) & ~63)) * RP * sizeof(float) + 8 + (size_t)2 * W * 32 * sizeof(float2);
  const size_t cap = 227 * 1024;
  int bps = (int)(cap / (p.smem + 1024));
  const int by_threads = 2048 / (W * 32);
  if (bps > by_threads) bps = by_threads;
  if (bps > 16) bps = 16;
  p.blocks_per_sm = bps;
  return p;
}

}  // namespace

bool featurize_rec_applicable(const Ctx &c, bool split, int64_t ld) {
  const int A = c.tri_n;
  if (c.feat_rec_off || A < 2) return false;
  const int F = A * (A - 1) / 2;
  if (F != c.F) return false;
  if (rec_plan(A).blocks_per_sm < 1) return false;
  return split ? (ld % 64 == 0 && ld > F && ld - F <= 64) : ld >= F;
}

void launch_featurize_rec(Ctx &c, const float *coords, const int64_t *gather, int64_t M, bool do_ln, float *out,
                          __nv_bfloat16 *out_hi, __nv_bfloat16 *out_lo, int64_t ld) {
  const int A = c.tri_n;
  const RecPlan p = rec_plan(A);
  const int64_t nblk = (M + 31) / 32;
  const int grid = (int)std::min<int64_t>(nblk, (int64_t)c.num_sms * p.blocks_per_sm);
  const float eps = c.cfg.ln_eps;
  auto go = [&](auto kernel, int id) {  // id: one attribute flag per kernel instantiation
    if (c.attr_needed(Ctx::ATTR_FEAT_BLK0 + id))
      IK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    kernel<<<grid, p.warps * 32, p.smem, c.stream>>>(coords, gather, M, c.D, A, c.tri_cmap.p, A * (A - 1) / 2,
                                                     do_ln ? 1 : 0, eps * eps, out, out_hi, out_lo, ld, c.split_fmt);
  };
  const bool map = c.tri_cmap.p != nullptr;
  if (out_hi) {
    if (map) go(featurize_blk_kernel<true, true>, 0);
    else go(featurize_blk_kernel<true, false>, 1);
  } else {
    if (map) go(featurize_blk_kernel<false, true>, 2);
    else go(featurize_blk_kernel<false, false>, 3);
  }
}

}  // namespace ik
