// tcgen05 / TMA GEMM for the wide Dense layers (fp32-accurate 3xBF16 split).
//
// Replaces the cuBLAS SGEMM + broadcast kernels behind Flux.Dense forward/backward for wide
// layers (reference call sites src/isotarget.jl:18, src/iso.jl:185).  The reference computes in
// FP32; tensor cores are reached without giving that up by splitting every operand into two
// bf16 terms  x = hi + lo  (hi = bf16(x), lo = bf16(x - hi), 16 mantissa bits kept) and issuing
// three MMAs per k-slice into one fp32 TMEM accumulator:  hi*hi + hi*lo + lo*hi  (the dropped
// lo*lo term is ~2^-16 relative).
//
//   D[M x N] = A[M x K] * B[N x K]^T            A, B: K-major bf16 (hi, lo), fp32 accumulate
//
// Structure (one CTA per SM, persistent over output tiles of 128 x 256):
//   warp 0   : TMA producer  - 4 tiles per stage (A_hi, A_lo, B_hi, B_lo), 128B swizzle, 2 stages
//   warp 1   : MMA issuer    - one elected lane issues tcgen05.mma (M=128, N=256, K=16) x 3 x 4 per
//              stage; tcgen05.commit releases the smem stage / publishes the accumulator
//   warps 2-5: epilogue      - tcgen05.ld the fp32 accumulator (double-buffered in TMEM: 2 x 256
//              columns, so the epilogue of tile i overlaps the MMAs of tile i+1), fuse
//              bias + activation (or the activation derivative of the backward pass), split to
//              bf16 hi/lo for the next layer's A operand, or store fp32 (weight gradients).
// Out-of-range rows/columns/k are zero-filled by TMA; stores are predicated.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc.cuh"

namespace ik {

namespace {

constexpr int BM = 128, BN = 256, BK = 64, STAGES = 2;
constexpr int A_TILE = BM * BK * 2;  // bytes of one bf16 A tile
constexpr int B_TILE = BN * BK * 2;
constexpr int STAGE_BYTES = 2 * A_TILE + 2 * B_TILE;  // 96 KiB
constexpr int TAIL_FLOATS = 3072;   // staged tail weights: (128+1)*16 + 17*16 + 17*16 floats at most
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + TAIL_FLOATS * 4;
constexpr int EPI_WARPS = 8;
constexpr int NTHREADS = 64 + 32 * EPI_WARPS;  // TMA warp + MMA warp + epilogue warps
constexpr uint32_t TMEM_COLS = 512;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// bounded spin: a protocol bug must surface as a trap, never as a hung GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  uint32_t spins = 0;
  long long t0 = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if ((++spins & 1023u) == 0) {  // try_wait suspends in hardware between probes; look at the clock rarely
      if (t0 == 0) t0 = clock64();
      else if (clock64() - t0 > 6000000000LL) __trap();  // ~3 s: protocol bug, not a slow tile
    }
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, 128B-swizzled operand tile: rows of 128 B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);  // start address
  d |= (uint64_t)1 << 16;                       // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset
  d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
  return d;
}
// MN-major, 128B-swizzled operand tile built from 64(MN) x 64(K) TMA boxes: 64 MN-elements (128 B) per
// k row, 8-row k groups 1024 B apart (stride byte offset), next 64 MN-elements 8192 B apart (leading byte offset)
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(8192 >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 256-bit global store (sm_100, PTX ISA 8.8): one full 32-byte sector per lane and half the LSU instructions
__device__ __forceinline__ void st_global_v8(void *ptr, const uint32_t *v) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

__device__ __forceinline__ float act_fwd(float a, int kind) {
  switch (kind) {
    case ISOKANN_ACT_SIGMOID: return __fdividef(1.0f, 1.0f + __expf(-a));
    case ISOKANN_ACT_TANH: return tanhf(a);
    case ISOKANN_ACT_RELU: return fmaxf(a, 0.f);
    default: return a;
  }
}
__device__ __forceinline__ float dact_o(float z, int kind) {
  switch (kind) {
    case ISOKANN_ACT_SIGMOID: return z * (1.0f - z);
    case ISOKANN_ACT_TANH: return 1.0f - z * z;
    case ISOKANN_ACT_RELU: return z > 0.f ? 1.0f : 0.f;
    default: return 1.0f;
  }
}

__device__ __forceinline__ void split_pair(float x0, float x1, uint32_t &hi, uint32_t &lo) {
  const __nv_bfloat162 h2 = __floats2bfloat162_rn(x0, x1);  // .x (low half) = x0
  hi = *reinterpret_cast<const uint32_t *>(&h2);
  const float r0 = x0 - __uint_as_float(hi << 16);
  const float r1 = x1 - __uint_as_float(hi & 0xFFFF0000u);
  const __nv_bfloat162 l2 = __floats2bfloat162_rn(r0, r1);
  lo = *reinterpret_cast<const uint32_t *>(&l2);
}
// fp16 variant (inference forward, TcGemm::fmt == 1): 11 significant bits per term, 22 kept
__device__ __forceinline__ void split_pair_h(float x0, float x1, uint32_t &hi, uint32_t &lo) {
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));  // low half = x0
  const float2 hf = __half22float2(*reinterpret_cast<const __half2 *>(&hi));
  const float r0 = x0 - hf.x, r1 = x1 - hf.y;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r1), "f"(r0));
}
__device__ __forceinline__ float bf16lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

struct TcParams {
  int M, N, K;          // D is M x N, reduction length K (elements)
  int m_tiles, n_tiles, splits, kb_per_split, num_kb;
  int epi, act, mn_major, ones_col;
  int fmt;              // operand format: 0 bf16 (hi, lo), 1 fp16 (hi, lo)
  int nmma;             // MMAs per k-slice: 3 = hi*hi + hi*lo + lo*hi; 2 = hi*hi + lo*hi (B rounded once, no B_lo tile)
  int st_v8;            // split outputs are 32-byte aligned with a 32-byte multiple pitch: 256-bit stores
  const float *bias;    // [N] or nullptr
  __nv_bfloat16 *out_hi, *out_lo;  // split outputs, row-major, leading dimension ldo
  int64_t ldo;
  float *out_f32;       // fp32 output (+ split-K slices of M*ldc)
  int64_t ldc;
  TcTail tail;          // TC_EPI_TAIL
  float *chi_out;
  const float *w_last;  // TC_EPI_BIAS_ACT_DOT: last-layer weights [N x d] row-major
  float *dot_out;       // partial chi [M x dot_slots x d]
  int d, dot_slots;
  int f32_vec;          // fp32 output is 16-byte aligned with a 16-byte multiple pitch: float4 stores
  const __nv_bfloat16 *z_hi, *z_lo;  // EPI_MULDACT: activation outputs (split), leading dimension ldz
  int64_t ldz;
};

// TC_EPI_TAIL: the whole rest of a narrow net per accumulator row -- bias + activation of the tensor-core layer
// (N <= 128 columns), then the tiny Dense layers (widths <= 16) with their [W; b] rows broadcast from shared memory.
// One warp = 32 rows of the accumulator.  Written for instruction count: the first ncu capture of the generic
// version showed ~3000 issued instructions per row (fully unrolled 16 x 16 predicated loops, 64 sigmoids for 38
// columns) and made the epilogue, not the MMAs or the operand stream, the bound of the kernel (27 % tensor-pipe
// activity).  Here every loop stops at the real extent with a warp-uniform branch, the first tail layer's padded
// width CP is a compile-time constant and the first layer's bias comes from shared memory (tailw + kTailBias).
constexpr int kTailBias = TAIL_FLOATS - 128;   // 128 staged bias values of the tensor-core layer

template <int CP>
__device__ __forceinline__ void tail_rows(const TcParams &p, uint32_t taddr0, int64_t row, bool row_ok,
                                          const float *tailw) {
  float tl[CP];
#pragma unroll
  for (int k = 0; k < CP; ++k) tl[k] = 0.f;
  const float *bias_s = tailw + kTailBias;
  const bool sig = p.act == ISOKANN_ACT_SIGMOID;
#pragma unroll 1
  for (int c = 0; c < p.N; c += 32) {
    uint32_t r[32];
    tmem_ld32(taddr0 + c, r);
    const int nv = p.N - c < 32 ? p.N - c : 32;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (j >= nv) break;  // warp-uniform
      float v = __uint_as_float(r[j]) + bias_s[c + j];
      v = sig ? __fdividef(1.0f, 1.0f + __expf(-v)) : act_fwd(v, p.act);
      const float4 *wr = reinterpret_cast<const float4 *>(tailw + (c + j) * CP);
#pragma unroll
      for (int q = 0; q < CP / 4; ++q) {
        const float4 w4 = wr[q];
        tl[4 * q] = fmaf(v, w4.x, tl[4 * q]);
        tl[4 * q + 1] = fmaf(v, w4.y, tl[4 * q + 1]);
        tl[4 * q + 2] = fmaf(v, w4.z, tl[4 * q + 2]);
        tl[4 * q + 3] = fmaf(v, w4.w, tl[4 * q + 3]);
      }
    }
  }
  if (!row_ok) return;
  float h[16];
  int off;
  {
    const int cols = p.tail.w[1];
    const float *brow = tailw + p.tail.w[0] * CP;
    const int kind = p.tail.nl == 1 ? p.tail.last_act : p.tail.act;
#pragma unroll
    for (int k = 0; k < 16; ++k) h[k] = (k < CP && k < cols) ? act_fwd(tl[k < CP ? k : 0] + brow[k < CP ? k : 0], kind) : 0.f;
    off = (p.tail.w[0] + 1) * CP;
  }
#pragma unroll 1
  for (int i = 1; i < p.tail.nl; ++i) {
    const int rows = p.tail.w[i], cols = p.tail.w[i + 1], cp = (cols + 3) & ~3;
    float a2[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (4 * q < cp) b4 = *reinterpret_cast<const float4 *>(tailw + off + rows * cp + 4 * q);
      a2[4 * q] = b4.x; a2[4 * q + 1] = b4.y; a2[4 * q + 2] = b4.z; a2[4 * q + 3] = b4.w;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (j >= rows) break;  // warp-uniform
      const float4 *wr = reinterpret_cast<const float4 *>(tailw + off + j * cp);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (4 * q >= cp) break;
        const float4 w4 = wr[q];
        a2[4 * q] = fmaf(h[j], w4.x, a2[4 * q]);
        a2[4 * q + 1] = fmaf(h[j], w4.y, a2[4 * q + 1]);
        a2[4 * q + 2] = fmaf(h[j], w4.z, a2[4 * q + 2]);
        a2[4 * q + 3] = fmaf(h[j], w4.w, a2[4 * q + 3]);
      }
    }
    const int kind = i == p.tail.nl - 1 ? p.tail.last_act : p.tail.act;
#pragma unroll
    for (int k = 0; k < 16; ++k) h[k] = k < cols ? act_fwd(a2[k], kind) : 0.f;
    off += (rows + 1) * cp;
  }
  const int dd = p.tail.w[p.tail.nl];
#pragma unroll
  for (int k = 0; k < kMaxD; ++k)
    if (k < dd) p.chi_out[row * dd + k] = h[k];
}

__device__ __forceinline__ void tail_rows_any(const TcParams &p, uint32_t taddr0, int64_t row, bool row_ok,
                                              const float *tailw) {
  switch ((p.tail.w[1] + 3) >> 2) {
    case 1: tail_rows<4>(p, taddr0, row, row_ok, tailw); break;
    case 2: tail_rows<8>(p, taddr0, row, row_ok, tailw); break;
    case 3: tail_rows<12>(p, taddr0, row, row_ok, tailw); break;
    default: tail_rows<16>(p, taddr0, row, row_ok, tailw); break;
  }
}

// stage the tail layers' [W; b] (rows padded to a multiple of 4 floats) and the tensor-core layer's bias
__device__ __forceinline__ void stage_tail(const TcParams &p, float *tailw) {
  int off = 0;
  for (int i = 0; i < p.tail.nl; ++i) {
    const int rows = p.tail.w[i] + 1, cols = p.tail.w[i + 1], cp = (cols + 3) & ~3;
    for (int e = threadIdx.x; e < rows * cp; e += blockDim.x) {
      const int r = e / cp, cc = e - r * cp;
      tailw[off + e] = cc < cols ? __ldg(p.tail.seg[i] + (int64_t)r * cols + cc) : 0.f;
    }
    off += rows * cp;
  }
  for (int e = threadIdx.x; e < 128; e += blockDim.x) tailw[kTailBias + e] = (e < p.N && p.bias) ? __ldg(p.bias + e) : 0.f;
}

// Epilogue of one 128 x 256 accumulator for one warp (its 32 rows, one 128-column half): shared by the 1-CTA and
// the 2-CTA kernels.  row0 is the first output row of this CTA's accumulator.
template <int EPI>
__device__ __forceinline__ void epilogue_tile(const TcParams &p, uint32_t tmem_acc, int64_t row0, int nb, int split,
                                              int quarter, int half, int lane, const float *tailw) {
  const int64_t row = row0 + quarter * 32 + lane;
  const bool row_ok = row < p.M;
  const uint32_t taddr0 = tmem_acc + ((uint32_t)(quarter * 32) << 16);
  if (EPI == TC_EPI_TAIL) {  // N <= 128: the whole row belongs to the warps of half 0
    if (half == 0) tail_rows_any(p, taddr0, row, row_ok, tailw);
    return;
  }
  float dot[kMaxD];
#pragma unroll
  for (int a = 0; a < kMaxD; ++a) dot[a] = 0.f;
  float tl[16];  // TC_EPI_TAIL: pre-activations of the first tail layer
#pragma unroll
  for (int a = 0; a < 16; ++a) tl[a] = 0.f;
#pragma unroll 1
  for (int c = half * (BN / 2); c < (half + 1) * (BN / 2); c += 32) {
    const int col0 = nb * BN + c;
    if (col0 >= p.N) break;  // warp-uniform
    uint32_t r[32];
    tmem_ld32(taddr0 + c, r);
    if (!row_ok) continue;
    const bool full = col0 + 32 <= p.N;
    if (EPI == TC_EPI_F32) {
      float *dst = p.out_f32 + (int64_t)split * p.M * p.ldc + row * p.ldc + col0;
      if (full && p.f32_vec) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4 *>(dst + j) =
              make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                          __uint_as_float(r[j + 3]));
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (col0 + j < p.N) dst[j] = __uint_as_float(r[j]);
      }
      continue;
    }
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
    if (EPI == TC_EPI_MULDACT_SPLIT) {
      const uint4 *zh = reinterpret_cast<const uint4 *>(p.z_hi + row * p.ldz + col0);
      const uint4 *zl = reinterpret_cast<const uint4 *>(p.z_lo + row * p.ldz + col0);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 h = __ldg(zh + q), l = __ldg(zl + q);
        const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float z0 = bf16lo(hw[e]) + bf16lo(lw[e]);
          const float z1 = bf16hi(hw[e]) + bf16hi(lw[e]);
          v[q * 8 + 2 * e] *= dact_o(z0, p.act);
          v[q * 8 + 2 * e + 1] *= dact_o(z1, p.act);
        }
      }
    } else {  // bias + activation
      if (p.bias) {
        if (full) {
          const float4 *b4 = reinterpret_cast<const float4 *>(p.bias + col0);
          if ((reinterpret_cast<uintptr_t>(b4) & 15) == 0) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 b = __ldg(b4 + q);
              v[4 * q] += b.x; v[4 * q + 1] += b.y; v[4 * q + 2] += b.z; v[4 * q + 3] += b.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += __ldg(p.bias + col0 + j);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < p.N) v[j] += __ldg(p.bias + col0 + j);
        }
      }
      if (p.act == ISOKANN_ACT_SIGMOID) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __fdividef(1.0f, 1.0f + __expf(-v[j]));
      } else if (p.act != ISOKANN_ACT_IDENTITY) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = act_fwd(v[j], p.act);
      }
    }
    if (EPI == TC_EPI_TAIL) {
      // first tail layer: a[k] += z[col] * W[col, k] with W rows broadcast from shared memory
      const int cp = (p.tail.w[1] + 3) & ~3;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (col0 + j < p.N) {
          const float4 *wr = reinterpret_cast<const float4 *>(tailw + (col0 + j) * cp);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (4 * q < cp) {
              const float4 w4 = wr[q];
              tl[4 * q] = fmaf(v[j], w4.x, tl[4 * q]);
              tl[4 * q + 1] = fmaf(v[j], w4.y, tl[4 * q + 1]);
              tl[4 * q + 2] = fmaf(v[j], w4.z, tl[4 * q + 2]);
              tl[4 * q + 3] = fmaf(v[j], w4.w, tl[4 * q + 3]);
            }
          }
        }
      }
      continue;
    }
    if (EPI == TC_EPI_BIAS_ACT_DOT) {
      // chi partial: sum over this warp's columns of z[col] * W_last[col, a]
      if (p.d == 1 && full && ((reinterpret_cast<uintptr_t>(p.w_last + col0) & 15) == 0)) {
        const float4 *w4 = reinterpret_cast<const float4 *>(p.w_last + col0);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 w = __ldg(w4 + q);
          dot[0] = fmaf(v[4 * q], w.x, dot[0]);
          dot[0] = fmaf(v[4 * q + 1], w.y, dot[0]);
          dot[0] = fmaf(v[4 * q + 2], w.z, dot[0]);
          dot[0] = fmaf(v[4 * q + 3], w.w, dot[0]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (col0 + j < p.N) {
            const float *wl = p.w_last + (int64_t)(col0 + j) * p.d;
#pragma unroll
            for (int a = 0; a < kMaxD; ++a)
              if (a < p.d) dot[a] = fmaf(v[j], __ldg(wl + a), dot[a]);
          }
        }
      }
      continue;
    }
    if (!full) {  // columns >= N inside the padded leading dimension: zeros, except the bias column N
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j >= p.N) v[j] = (p.ones_col && col0 + j == p.N) ? 1.f : 0.f;
    }
    uint32_t ph[16], pl[16];
    if (p.fmt) {
#pragma unroll
      for (int j = 0; j < 16; ++j) split_pair_h(v[2 * j], v[2 * j + 1], ph[j], pl[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) split_pair(v[2 * j], v[2 * j + 1], ph[j], pl[j]);
    }
    __nv_bfloat16 *dh = p.out_hi + row * p.ldo + col0;
    __nv_bfloat16 *dl = p.out_lo + row * p.ldo + col0;
    if (p.st_v8) {
      st_global_v8(dh, ph);
      st_global_v8(dh + 16, ph + 8);
      st_global_v8(dl, pl);
      st_global_v8(dl + 16, pl + 8);
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        reinterpret_cast<uint4 *>(dh)[q] = make_uint4(ph[4 * q], ph[4 * q + 1], ph[4 * q + 2], ph[4 * q + 3]);
        reinterpret_cast<uint4 *>(dl)[q] = make_uint4(pl[4 * q], pl[4 * q + 1], pl[4 * q + 2], pl[4 * q + 3]);
      }
    }
  }
  if (EPI == TC_EPI_TAIL && row_ok && half == 0) {
    // finish the tail: bias + activation of its first layer, then the remaining (tiny) layers
    int off = 0;
    float h[16];
    {
      const int cols = p.tail.w[1], cp = (cols + 3) & ~3;
      const float *brow = tailw + p.tail.w[0] * cp;
      const int kind = p.tail.nl == 1 ? p.tail.last_act : p.tail.act;
#pragma unroll
      for (int k = 0; k < 16; ++k) h[k] = k < cols ? act_fwd(tl[k] + brow[k], kind) : 0.f;
      off += (p.tail.w[0] + 1) * cp;
    }
    for (int i = 1; i < p.tail.nl; ++i) {
      const int rows = p.tail.w[i], cols = p.tail.w[i + 1], cp = (cols + 3) & ~3;
      float a2[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) a2[k] = k < cols ? tailw[off + rows * cp + k] : 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (j < rows) {
#pragma unroll
          for (int k = 0; k < 16; ++k)
            if (k < cp) a2[k] = fmaf(h[j], tailw[off + j * cp + k], a2[k]);
        }
      }
      const int kind = i == p.tail.nl - 1 ? p.tail.last_act : p.tail.act;
#pragma unroll
      for (int k = 0; k < 16; ++k) h[k] = k < cols ? act_fwd(a2[k], kind) : 0.f;
      off += (rows + 1) * cp;
    }
    const int dd = p.tail.w[p.tail.nl];
#pragma unroll
    for (int k = 0; k < kMaxD; ++k)
      if (k < dd) p.chi_out[row * dd + k] = h[k];
  }
  if (EPI == TC_EPI_BIAS_ACT_DOT && row_ok) {
    float *dst = p.dot_out + (row * p.dot_slots + (nb * 2 + half)) * p.d;
    for (int a = 0; a < p.d; ++a) dst[a] = dot[a];
  }
}

template <int EPI>
__global__ void __launch_bounds__(NTHREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap map_ah, const __grid_constant__ CUtensorMap map_al,
               const __grid_constant__ CUtensorMap map_bh, const __grid_constant__ CUtensorMap map_bl, TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;            // 1024-B alignment for SWIZZLE_128B
  uint8_t *base_ptr = smem_raw + (base - raw);
  uint64_t *bars = reinterpret_cast<uint64_t *>(base_ptr + STAGES * STAGE_BYTES);
  // barrier slots: full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2]
  const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * STAGES;
  const uint32_t bar_tfull = bar_empty + 8 * STAGES, bar_tempty = bar_tfull + 16;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.n_tiles * p.splits;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM allocation is a warp-wide operation
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // TC_EPI_TAIL: stage the tail layers' [W; b] in shared memory, rows padded to a multiple of 4 floats
  float *tailw = reinterpret_cast<float *>(base_ptr + STAGES * STAGE_BYTES + 256);
  if (EPI == TC_EPI_TAIL) {
    stage_tail(p, tailw);
    __syncthreads();
  }

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int split = t / (p.m_tiles * p.n_tiles);
        const int rem = t - split * (p.m_tiles * p.n_tiles);
        const int mb = rem / p.n_tiles, nb = rem - mb * p.n_tiles;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.num_kb, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t sa = base + stage * STAGE_BYTES;
          const uint32_t full = bar_full + 8 * stage;
          mbar_arrive_expect_tx(full, p.nmma == 2 ? STAGE_BYTES - B_TILE : STAGE_BYTES);
          if (!p.mn_major) {
            tma_load_2d(sa, &map_ah, full, kb * BK, mb * BM);
            tma_load_2d(sa + A_TILE, &map_al, full, kb * BK, mb * BM);
            tma_load_2d(sa + 2 * A_TILE, &map_bh, full, kb * BK, nb * BN);
            if (p.nmma != 2) tma_load_2d(sa + 2 * A_TILE + B_TILE, &map_bl, full, kb * BK, nb * BN);
          } else {  // boxes of 64 (M or N, contiguous) x 64 (k)
#pragma unroll
            for (int b = 0; b < BM / 64; ++b) {
              tma_load_2d(sa + b * 8192, &map_ah, full, mb * BM + 64 * b, kb * BK);
              tma_load_2d(sa + A_TILE + b * 8192, &map_al, full, mb * BM + 64 * b, kb * BK);
            }
#pragma unroll
            for (int b = 0; b < BN / 64; ++b) {
              tma_load_2d(sa + 2 * A_TILE + b * 8192, &map_bh, full, nb * BN + 64 * b, kb * BK);
              tma_load_2d(sa + 2 * A_TILE + B_TILE + b * 8192, &map_bl, full, nb * BN + 64 * b, kb * BK);
            }
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // instruction descriptor: D=f32, A=B=bf16, both K-major, N=256, M=128
      // (format fields: 1 = bf16, 0 = fp16)
      const uint32_t fbits = p.fmt ? 0u : ((1u << 7) | (1u << 10));
      uint32_t idesc = (1u << 4) | fbits | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      if (p.mn_major) idesc |= (1u << 15) | (1u << 16);  // A and B are MN-major
      uint32_t stage = 0, phase = 0;
      int it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
        const int split = t / (p.m_tiles * p.n_tiles);
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.num_kb, kb0 + p.kb_per_split);
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sa = base + stage * STAGE_BYTES;
          const bool mn = p.mn_major != 0;
          const uint64_t dah = mn ? make_desc_mn_sw128(sa) : make_desc_sw128(sa);
          const uint64_t dal = mn ? make_desc_mn_sw128(sa + A_TILE) : make_desc_sw128(sa + A_TILE);
          const uint64_t dbh = mn ? make_desc_mn_sw128(sa + 2 * A_TILE) : make_desc_sw128(sa + 2 * A_TILE);
          const uint64_t dbl =
              mn ? make_desc_mn_sw128(sa + 2 * A_TILE + B_TILE) : make_desc_sw128(sa + 2 * A_TILE + B_TILE);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // K-major: 16 bf16 = 32 B inside the swizzle atom; MN-major: 16 k rows = two 1024 B groups
            const uint64_t adv = mn ? (uint64_t)((k * 2048) >> 4) : (uint64_t)((k * 32) >> 4);
            umma_f16(tmem_d, dah + adv, dbh + adv, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            if (p.nmma != 2) umma_f16(tmem_d, dah + adv, dbl + adv, idesc, 1u);
            umma_f16(tmem_d, dal + adv, dbh + adv, idesc, 1u);
          }
          umma_commit(bar_empty + 8 * stage);  // smem stage reusable once these MMAs retire
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(bar_tfull + 8 * acc);  // accumulator complete
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    // a warp may only touch TMEM lanes [32*(warp%4), +32): two warps share each lane quarter and
    // take one 128-column half of the accumulator each
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const int split = t / (p.m_tiles * p.n_tiles);
      const int rem = t - split * (p.m_tiles * p.n_tiles);
      const int mb = rem / p.n_tiles, nb = rem - mb * p.n_tiles;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(bar_tfull + 8 * acc, acc_phase);
      tc_fence_after();
      epilogue_tile<EPI>(p, tmem_base + acc * BN, (int64_t)mb * BM, nb, split, quarter, half, lane, tailw);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}


// ------------------------------------------------------------------------------------------
// 2-CTA variant (cta_group::2): a pair of CTAs on neighbouring SMs computes one 256 x 256 tile.  Each CTA loads its
// own 128 rows of A and HALF of the B tile (128 of the 256 N rows); the leader CTA issues tcgen05.mma with M = 256
// and the tensor cores of both SMs read both halves, so the B operand crosses L2 -> SMEM once per pair: 64 KiB
// instead of 96 KiB per CTA and k-block (3 stages fit).  Barrier protocol:
//   full[s]    leader only: its producer arms 2 x 64 KiB; both CTAs' TMA loads complete_tx on the LEADER's barrier
//   empty[s]   both CTAs: tcgen05.commit ... multicast::cluster arrives on both when the MMAs have read stage s
//   tfull[a]   both CTAs: multicast commit when the accumulator is complete (each CTA drains its own TMEM)
//   tempty[a]  leader only: 2 x 8 epilogue warps arrive (the peer's through mapa / shared::cluster)
// ------------------------------------------------------------------------------------------
constexpr int STAGES2 = 3;
constexpr int B2_TILE = (BN / 2) * BK * 2;
constexpr int STAGE2_BYTES = 2 * A_TILE + 2 * B2_TILE;  // 64 KiB per CTA
constexpr int SMEM2_BYTES = STAGES2 * STAGE2_BYTES + 1024 + 256;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_cg2(uint32_t dst, const CUtensorMap *map, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t local_bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
      :
      : "r"(local_bar), "r"(cta)
      : "memory");
}

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
tc_gemm2_kernel(const __grid_constant__ CUtensorMap map_ah, const __grid_constant__ CUtensorMap map_al,
                const __grid_constant__ CUtensorMap map_bh, const __grid_constant__ CUtensorMap map_bl, TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *base_ptr = smem_raw + (base - raw);
  uint64_t *bars = reinterpret_cast<uint64_t *>(base_ptr + STAGES2 * STAGE2_BYTES);
  const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * STAGES2;
  const uint32_t bar_tfull = bar_empty + 8 * STAGES2, bar_tempty = bar_tfull + 16;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * STAGES2 + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int m_tiles2 = (p.M + 2 * BM - 1) / (2 * BM);
  const int total_tiles = m_tiles2 * p.n_tiles;
  const int cl = blockIdx.x >> 1, ncl = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES2; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, 2 * EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int t = cl; t < total_tiles; t += ncl) {
        const int mb2 = t / p.n_tiles, nb = t - mb2 * p.n_tiles;
        const int m0 = mb2 * 2 * BM + (int)rank * BM;
        const int n0 = nb * BN + (int)rank * (BN / 2);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t sa = base + stage * STAGE2_BYTES;
          const uint32_t full_local = bar_full + 8 * stage;
          const uint32_t full_leader = full_local & 0xFEFFFFFFu;  // peer bit cleared: the even CTA of the pair
          if (leader) mbar_arrive_expect_tx(full_local, 2 * (p.nmma == 2 ? STAGE2_BYTES - B2_TILE : STAGE2_BYTES));
          tma_load_2d_cg2(sa, &map_ah, full_leader, kb * BK, m0);
          tma_load_2d_cg2(sa + A_TILE, &map_al, full_leader, kb * BK, m0);
          tma_load_2d_cg2(sa + 2 * A_TILE, &map_bh, full_leader, kb * BK, n0);
          if (p.nmma != 2) tma_load_2d_cg2(sa + 2 * A_TILE + B2_TILE, &map_bl, full_leader, kb * BK, n0);
          if (++stage == STAGES2) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && lane == 0) {
      // D=f32, A=B=bf16, K-major, N=256, M=256 (2 x 128 across the CTA pair)
      const uint32_t fbits = p.fmt ? 0u : ((1u << 7) | (1u << 10));  // 1 = bf16, 0 = fp16
      const uint32_t idesc = (1u << 4) | fbits | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
      uint32_t stage = 0, phase = 0;
      int it = 0;
      for (int t = cl; t < total_tiles; t += ncl, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sa = base + stage * STAGE2_BYTES;
          const uint64_t dah = make_desc_sw128(sa), dal = make_desc_sw128(sa + A_TILE);
          const uint64_t dbh = make_desc_sw128(sa + 2 * A_TILE), dbl = make_desc_sw128(sa + 2 * A_TILE + B2_TILE);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t adv = (uint64_t)((k * 32) >> 4);
            umma2_f16(tmem_d, dah + adv, dbh + adv, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            if (p.nmma != 2) umma2_f16(tmem_d, dah + adv, dbl + adv, idesc, 1u);
            umma2_f16(tmem_d, dal + adv, dbh + adv, idesc, 1u);
          }
          umma_commit_mc(bar_empty + 8 * stage);
          if (++stage == STAGES2) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_mc(bar_tfull + 8 * acc);
      }
    }
  } else {
    // ===================== epilogue (both CTAs, own 128 rows) =====================
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    int it = 0;
    for (int t = cl; t < total_tiles; t += ncl, ++it) {
      const int mb2 = t / p.n_tiles, nb = t - mb2 * p.n_tiles;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(bar_tfull + 8 * acc, acc_phase);
      tc_fence_after();
      epilogue_tile<EPI>(p, tmem_base + acc * BN, (int64_t)mb2 * 2 * BM + (int64_t)rank * BM, nb, 0, quarter, half, lane,
                         nullptr);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(bar_tempty + 8 * acc);
        else mbar_arrive_cluster(bar_tempty + 8 * acc, 0);
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}


// ------------------------------------------------------------------------------------------
// Fused narrow-net forward: coordinates -> chi in ONE kernel (pairnet F -> h1 <= 128 -> <= 16 ... -> d).
// Replaces flatpairdists + cached features + model(x) of the reference for the K*N-sample Koopman pass
// (src/utils/pairdists.jl:6-24, src/simulation.jl:110-114, src/isotarget.jl:18): x_hat never exists in HBM, the
// kernel reads 4*D bytes per sample and writes 4*d.
//
//   warps 0..15  producers.  A tile is 128 records; lane = record, the 4 warps of a record group split every k-block
//                of 64 features in 16-feature quarters.  Coordinates are staged transposed ([coordinate][record]) by
//                cp.async.  Pass 1 walks the strict upper triangle (column-major, halfinds order) once for the
//                LayerNorm statistics (pivoted sums: sum(d - d0), sum((d - d0)^2), so no cancellation); pass 2 walks
//                it again k-block by k-block, normalises, splits into bf16 (hi, lo) and stores 16-byte pieces
//                straight into the 128B-swizzled K-major A stage that tcgen05.mma reads.  (Recomputing the distances
//                costs ~12 instructions per feature; parking them would cost 84 KB per 32 villin records and cap the
//                tile at 32 rows.)
//   warp 16      MMA issuer: 3 MMAs per k-slice (hi*hi, hi*lo, lo*hi), M = 128, N = h1 rounded up to 16
//   warp 17      TMA loads of the first layer's weights (hi, lo), one [N x 64] box pair per k-block
//   warps 18..21 epilogue (warp % 4 = TMEM lane quarter): bias + activation + the remaining tiny layers per row
//                (the TC_EPI_TAIL epilogue of the GEMM above), chi to global memory
// A stages: 2 x 32 KiB, B stages: 2 x (2 * N * 128 B), accumulator double-buffered in TMEM.
// ------------------------------------------------------------------------------------------
constexpr int KF_PW = 16;                      // producer warps
constexpr int KF_THREADS = 32 * (KF_PW + 6);   // + MMA, TMA, 4 epilogue warps
constexpr int KF_ROWS = 128;
constexpr int KF_CP = 129;                     // words per staged coordinate row (128 records + 1 pad)
constexpr int KF_ASTAGE = 2 * KF_ROWS * 128;   // bytes: hi + lo tile of one k-block

struct KoopFusedP {
  const float *coords;       // [M][D] records
  const int *cmap;           // atom subset: coordinate c of the selection -> coordinate of the record, or nullptr
  int64_t M;
  int D, A, F, C;            // record length, atoms in the triangle, features, staged coordinates (3A)
  int nkb, nmma_n;           // k-blocks of 64 features, MMA N (multiple of 16, <= 128)
  int do_ln;
  float eps2;
  const short2 *start16;     // (i, j) of feature 16*t for every t (column-major strict upper triangle)
  TcParams ep;               // epilogue parameters (TC_EPI_TAIL)
};

// 16 consecutive features of the column-major strict upper triangle for one record (lane = record), starting at
// (i, j).  at: shared-memory address of the record's staged coordinates ([coordinate][record], pitch KF_CP).
// EMIT == false: accumulate the pivoted LayerNorm sums; EMIT == true: v[k] = d * scale + shift.  i and j are
// warp-uniform, so the column switch is a non-divergent branch; FULL: all 16 features exist.
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}

// (32-bit shared-window addresses: with generic pointers a third of the issued instructions were 64-bit pointer
//  arithmetic and generic->shared conversions.)
template <bool EMIT, bool FULL>
__device__ __forceinline__ void walk16(uint32_t at, int i, int j, int nv, float piv, float scale, float shift,
                                       float &s1, float &s2, float (&v)[16]) {
  constexpr uint32_t ROW = 4u * KF_CP;          // bytes between consecutive coordinates of a record
  uint32_t ap = at + 3u * ROW * (uint32_t)i;    // atom i (row of the pair)
  uint32_t cq = at + 3u * ROW * (uint32_t)j;    // atom j (column of the pair), kept in registers
  float cx = lds_f32(cq), cy = lds_f32(cq + ROW), cz = lds_f32(cq + 2 * ROW);
  int sw = j - i;                               // index k at which the walk moves on to the next column
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    if (!FULL && k >= nv) {
      if (EMIT) v[k] = 0.f;
      continue;
    }
    if (k == sw) {  // next column (warp-uniform).  The __syncwarp keeps this a real branch: if-converted, the eight
      __syncwarp();  // instructions of the switch were issued, predicated off, for every single feature
      cq += 3 * ROW;
      cx = lds_f32(cq); cy = lds_f32(cq + ROW); cz = lds_f32(cq + 2 * ROW);
      ap = at;
      sw = k + (++j);
    }
    const float dx = lds_f32(ap) - cx, dy = lds_f32(ap + ROW) - cy, dz = lds_f32(ap + 2 * ROW) - cz;
    const float sq = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
    float d;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(sq));
    ap += 3 * ROW;
    if (EMIT) {
      v[k] = fmaf(d, scale, shift);
    } else {
      const float e = d - piv;
      s1 += e;
      s2 = fmaf(e, e, s2);
    }
  }
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(KF_THREADS, 1)
koop_fused_kernel(const __grid_constant__ CUtensorMap map_bh, const __grid_constant__ CUtensorMap map_bl, KoopFusedP p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *base_ptr = smem_raw + (base - raw);
  const int bstage = 2 * p.nmma_n * 128;                          // bytes of one B stage (hi + lo)
  const int bstage_al = (bstage + 1023) & ~1023;
  // layout: A stages | B stages | barriers | tail weights | stats | coordinates
  const uint32_t a_off = 0, b_off = 2 * KF_ASTAGE, bar_off = b_off + 2 * bstage_al;
  uint64_t *bars = reinterpret_cast<uint64_t *>(base_ptr + bar_off);
  const uint32_t bar_afull = smem_u32(bars), bar_aempty = bar_afull + 16, bar_bfull = bar_afull + 32,
                 bar_bempty = bar_afull + 48, bar_tfull = bar_afull + 64, bar_tempty = bar_afull + 80;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 12);
  float *tailw = reinterpret_cast<float *>(base_ptr + bar_off + 128);
  float2 *part = reinterpret_cast<float2 *>(tailw + TAIL_FLOATS);       // [4][128] partial (s1, s2) per quarter
  float2 *stat = part + 4 * KF_ROWS;                                     // [128] (scale, shift)
  float *coord = reinterpret_cast<float *>(stat + KF_ROWS);              // [C][KF_CP]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ntiles = (p.M + KF_ROWS - 1) / KF_ROWS;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_afull + 8 * s, KF_PW);
      mbar_init(bar_aempty + 8 * s, 1);
      mbar_init(bar_bfull + 8 * s, 1);
      mbar_init(bar_bempty + 8 * s, 1);
      mbar_init(bar_tfull + 8 * s, 1);
      mbar_init(bar_tempty + 8 * s, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == KF_PW) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(256u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  stage_tail(p.ep, tailw);  // tail layers' [W; b] and the first layer's bias
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < KF_PW) {
    // ===================== producers =====================
    const int grp = warp & 3, qtr = warp >> 2;        // record group (32 records), 16-feature quarter of a k-block
    const int rec = grp * 32 + lane;                  // row of the tile
    const float *atp = coord + rec;
    const uint32_t at = smem_u32(atp);
    const float invF = 1.0f / (float)p.F;
    const uint32_t coord_s = smem_u32(coord);
    uint32_t it = 0;                                  // k-blocks produced so far (stage = it & 1)
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
      const int64_t m0 = t * KF_ROWS;
      const int nvalid = (int)(p.M - m0 < KF_ROWS ? p.M - m0 : KF_ROWS);
      // ---- stage the coordinates of the tile, transposed; rows past M repeat the last valid record
      named_bar_sync(1, KF_PW * 32);  // everybody has finished reading the previous tile's coordinates
      for (int r = warp; r < KF_ROWS; r += KF_PW) {
        const int64_t m = m0 + (r < nvalid ? r : nvalid - 1);
        const float *g = p.coords + m * p.D;
        for (int c = lane; c < p.C; c += 32) {
          const float *src = g + (p.cmap ? __ldg(p.cmap + c) : c);
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(coord_s + 4u * (uint32_t)(c * KF_CP + r)), "l"(src)
                       : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_all;" ::: "memory");
      named_bar_sync(1, KF_PW * 32);
      float scale = 1.f, shift = 0.f;
      // pivot of the variance sums: the record's first distance
      float piv;
      {
        const float dx = atp[0] - atp[3 * KF_CP], dy = atp[KF_CP] - atp[4 * KF_CP], dz = atp[2 * KF_CP] - atp[5 * KF_CP];
        piv = sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
      }
      // ---- pass 1: LayerNorm statistics (pivoted sums over this thread's quarter of every k-block)
      if (p.do_ln) {
        float s1 = 0.f, s2 = 0.f, dummy[16];
        for (int kb = 0; kb < p.nkb; ++kb) {
          const int f0 = kb * 64 + qtr * 16;
          if (f0 >= p.F) break;
          const short2 st = __ldg(p.start16 + (f0 >> 4));
          if (f0 + 16 <= p.F) walk16<false, true>(at, st.x, st.y, 16, piv, 1.f, 0.f, s1, s2, dummy);
          else walk16<false, false>(at, st.x, st.y, p.F - f0, piv, 1.f, 0.f, s1, s2, dummy);
        }
        part[qtr * KF_ROWS + rec] = make_float2(s1, s2);
        named_bar_sync(1, KF_PW * 32);
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 x = part[q * KF_ROWS + rec];
          a += x.x;
          b += x.y;
        }
        const float me = a * invF;                                  // mean of (d - pivot)
        scale = rsqrtf(fmaxf(fmaf(-me, me, b * invF), 0.f) + p.eps2);
        shift = -(piv + me) * scale;
      }
      // ---- pass 2: k-block by k-block into the A stages
      for (int kb = 0; kb < p.nkb; ++kb, ++it) {
        const int f0 = kb * 64 + qtr * 16;
        const uint32_t stage = it & 1u;
        mbar_wait(bar_aempty + 8 * stage, ((it >> 1) & 1u) ^ 1u);   // the MMAs that read this stage have retired
        const uint32_t a_hi_s = base + a_off + stage * KF_ASTAGE;
        float v[16], u1 = 0.f, u2 = 0.f;
        if (f0 + 16 <= p.F) {
          const short2 st = __ldg(p.start16 + (f0 >> 4));
          walk16<true, true>(at, st.x, st.y, 16, piv, scale, shift, u1, u2, v);
        } else if (f0 < p.F) {
          const short2 st = __ldg(p.start16 + (f0 >> 4));
          walk16<true, false>(at, st.x, st.y, p.F - f0, piv, scale, shift, u1, u2, v);
        } else {
#pragma unroll
          for (int k = 0; k < 16; ++k) v[k] = 0.f;
        }
        // row `rec` of the K-major, 128B-swizzled tile: 16-byte chunk c of the row sits at chunk c ^ (rec & 7)
        uint32_t h[8], l[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) split_pair(v[2 * k], v[2 * k + 1], h[k], l[k]);
        const uint32_t row = a_hi_s + (uint32_t)rec * 128u;
        const uint32_t c0 = (uint32_t)((2 * qtr) ^ (rec & 7)) << 4, c1 = (uint32_t)((2 * qtr + 1) ^ (rec & 7)) << 4;
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + c0), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + c1), "r"(h[4]), "r"(h[5]), "r"(h[6]), "r"(h[7]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + KF_ROWS * 128 + c0), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + KF_ROWS * 128 + c1), "r"(l[4]), "r"(l[5]), "r"(l[6]), "r"(l[7]) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> tensor-core reads
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_afull + 8 * stage);
      }
    }
  } else if (warp == KF_PW) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.nmma_n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      uint32_t it = 0;
      int tl = 0;
      for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++tl) {
        const int acc = tl & 1;
        mbar_wait(bar_tempty + 8 * acc, ((tl >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * 128;
        for (int kb = 0; kb < p.nkb; ++kb, ++it) {
          const uint32_t stage = it & 1u, ph = (it >> 1) & 1u;
          mbar_wait(bar_afull + 8 * stage, ph);
          mbar_wait(bar_bfull + 8 * stage, ph);
          tc_fence_after();
          const uint32_t sa = base + a_off + stage * KF_ASTAGE, sb = base + b_off + stage * bstage_al;
          const uint64_t dah = make_desc_sw128(sa), dal = make_desc_sw128(sa + KF_ROWS * 128);
          const uint64_t dbh = make_desc_sw128(sb), dbl = make_desc_sw128(sb + p.nmma_n * 128);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t adv = (uint64_t)((k * 32) >> 4);
            umma_f16(tmem_d, dah + adv, dbh + adv, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma_f16(tmem_d, dah + adv, dbl + adv, idesc, 1u);
            umma_f16(tmem_d, dal + adv, dbh + adv, idesc, 1u);
          }
          umma_commit(bar_aempty + 8 * stage);
          umma_commit(bar_bempty + 8 * stage);
        }
        umma_commit(bar_tfull + 8 * acc);
      }
    }
  } else if (warp == KF_PW + 1) {
    // ===================== TMA: first-layer weights =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        for (int kb = 0; kb < p.nkb; ++kb, ++it) {
          const uint32_t stage = it & 1u, ph = (it >> 1) & 1u;
          mbar_wait(bar_bempty + 8 * stage, ph ^ 1u);
          const uint32_t sb = base + b_off + stage * bstage_al, full = bar_bfull + 8 * stage;
          mbar_arrive_expect_tx(full, (uint32_t)bstage);
          tma_load_2d(sb, &map_bh, full, kb * BK, 0);
          tma_load_2d(sb + p.nmma_n * 128, &map_bl, full, kb * BK, 0);
        }
      }
    }
  } else {
    // ===================== epilogue (warps 18..21) =====================
    const int quarter = warp & 3;
    int tl = 0;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++tl) {
      const int acc = tl & 1;
      mbar_wait(bar_tfull + 8 * acc, (tl >> 1) & 1);
      tc_fence_after();
      epilogue_tile<TC_EPI_TAIL>(p.ep, tmem_base + acc * 128, t * KF_ROWS, 0, 0, quarter, 0, lane, tailw);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == KF_PW) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void *sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  IK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres));
  IK_REQUIRE(sym != nullptr && qres == cudaDriverEntryPointSuccess, ISOKANN_ERR_CUDA,
             "cuTensorMapEncodeTiled is not available from this driver");
  fn = (EncodeTiledFn)sym;
  return fn;
}

// rows x cols bf16, row-major with leading dimension ld (elements); box = box_rows x 64 columns
void make_map(CUtensorMap *m, const void *ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  IK_REQUIRE(((uintptr_t)ptr & 15) == 0 && (ld * 2) % 16 == 0, ISOKANN_BAD_ARGUMENT,
             "tensor-core operand must be 16-byte aligned with a 16-byte multiple row pitch");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  IK_REQUIRE(r == CUDA_SUCCESS, ISOKANN_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
}

// MN-major operand: memory is [k_extent][ld] with mn_extent contiguous elements per row; boxes 64 x 64
void make_map_mn(CUtensorMap *m, const void *ptr, int64_t mn_extent, int64_t k_extent, int64_t ld) {
  IK_REQUIRE(((uintptr_t)ptr & 15) == 0 && (ld * 2) % 16 == 0, ISOKANN_BAD_ARGUMENT,
             "tensor-core operand must be 16-byte aligned with a 16-byte multiple row pitch");
  cuuint64_t dims[2] = {(cuuint64_t)mn_extent, (cuuint64_t)k_extent};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)BK};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  IK_REQUIRE(r == CUDA_SUCCESS, ISOKANN_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
}

}  // namespace

int launch_tc_gemm(Ctx &c, const TcGemm &g) {
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return 0;
  if (c.attr_needed(Ctx::ATTR_TC1)) {
    IK_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<TC_EPI_BIAS_ACT_SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    IK_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<TC_EPI_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    IK_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<TC_EPI_MULDACT_SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    IK_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<TC_EPI_BIAS_ACT_DOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    IK_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<TC_EPI_TAIL>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  }
  alignas(64) CUtensorMap mah, mal, mbh, mbl;
  if (g.mn_major) {
    make_map_mn(&mah, g.a_hi, g.M, g.K, g.lda);
    make_map_mn(&mal, g.a_lo, g.M, g.K, g.lda);
    make_map_mn(&mbh, g.b_hi, g.N, g.K, g.ldb);
    make_map_mn(&mbl, g.b_lo, g.N, g.K, g.ldb);
  } else {
    make_map(&mah, g.a_hi, g.M, g.K, g.lda, BM);
    make_map(&mal, g.a_lo, g.M, g.K, g.lda, BM);
    make_map(&mbh, g.b_hi, g.N, g.K, g.ldb, BN);
    make_map(&mbl, g.b_lo, g.N, g.K, g.ldb, BN);
  }
  TcParams p{};
  p.M = g.M; p.N = g.N; p.K = g.K;
  p.m_tiles = cdiv(g.M, BM);
  p.n_tiles = cdiv(g.N, BN);
  p.num_kb = cdiv(g.K, BK);
  int splits = g.splits < 1 ? 1 : g.splits;
  if (g.epi != TC_EPI_F32) splits = 1;
  splits = std::min(splits, p.num_kb);
  p.kb_per_split = cdiv(p.num_kb, splits);
  p.splits = cdiv(p.num_kb, p.kb_per_split);
  p.epi = g.epi; p.act = g.act; p.mn_major = g.mn_major; p.ones_col = g.ones_col;
  p.fmt = g.fmt; p.nmma = g.nmma == 2 ? 2 : 3;
  p.bias = g.bias;
  p.out_hi = g.out_hi; p.out_lo = g.out_lo; p.ldo = g.ldo;
  p.out_f32 = g.out_f32; p.ldc = g.ldc;
  p.f32_vec = (((uintptr_t)g.out_f32 & 15) == 0 && (g.ldc & 3) == 0 && (((int64_t)g.M * g.ldc) & 3) == 0) ? 1 : 0;
  p.z_hi = g.z_hi; p.z_lo = g.z_lo; p.ldz = g.ldz;
  p.tail = g.tail; p.chi_out = g.chi_out;
  p.st_v8 = (((uintptr_t)g.out_hi & 31) == 0 && ((uintptr_t)g.out_lo & 31) == 0 && (g.ldo * 2) % 32 == 0) ? 1 : 0;
  p.w_last = g.w_last; p.dot_out = g.dot_out; p.d = g.d; p.dot_slots = 2 * p.n_tiles;
  if (g.epi == TC_EPI_BIAS_ACT_SPLIT || g.epi == TC_EPI_MULDACT_SPLIT)
    IK_REQUIRE(g.ldo % 8 == 0 && g.ldo >= (int64_t)p.n_tiles * 0 + ((g.N + 31) / 32) * 32, ISOKANN_BAD_ARGUMENT,
               "split output needs a leading dimension padded to 32 columns");
  if (g.epi == TC_EPI_MULDACT_SPLIT)
    IK_REQUIRE(g.ldz % 8 == 0 && g.ldz >= ((g.N + 31) / 32) * 32, ISOKANN_BAD_ARGUMENT, "z leading dimension");
  if (g.epi == TC_EPI_TAIL) {
    IK_REQUIRE(g.N <= 128 && g.tail.nl >= 1 && g.tail.nl <= 3 && g.tail.w[0] == g.N, ISOKANN_BAD_ARGUMENT,
               "tail epilogue needs N <= 128 and 1..3 tail layers");
    int fl = 0;
    for (int i = 0; i < g.tail.nl; ++i) {
      IK_REQUIRE(g.tail.w[i + 1] >= 1 && g.tail.w[i + 1] <= 16, ISOKANN_BAD_ARGUMENT, "tail widths must be <= 16");
      fl += (g.tail.w[i] + 1) * ((g.tail.w[i + 1] + 3) & ~3);
    }
    IK_REQUIRE(fl <= TAIL_FLOATS - 128 && g.tail.w[g.tail.nl] <= kMaxD, ISOKANN_BAD_ARGUMENT, "tail too large");
  }
  // 2-CTA pairs for the large K-major GEMMs (forward / data gradient): B crosses L2 -> SMEM once per pair
  const int m_tiles2 = cdiv(g.M, 2 * BM);
  const int sms = c.num_sms - c.sm_reserve;  // an overlapped multi-rank step leaves SMs to the NCCL kernels
  const bool pair = !g.mn_major && g.epi != TC_EPI_F32 && g.epi != TC_EPI_TAIL && g.N > BN / 2 &&
                    m_tiles2 * p.n_tiles >= sms / 2 && !c.tc_no_pair;
  if (pair) {
    if (c.attr_needed(Ctx::ATTR_TC2)) {
      IK_CUDA(cudaFuncSetAttribute(tc_gemm2_kernel<TC_EPI_BIAS_ACT_SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES));
      IK_CUDA(cudaFuncSetAttribute(tc_gemm2_kernel<TC_EPI_MULDACT_SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES));
      IK_CUDA(cudaFuncSetAttribute(tc_gemm2_kernel<TC_EPI_BIAS_ACT_DOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES));
    }
    make_map(&mbh, g.b_hi, g.N, g.K, g.ldb, BN / 2);  // each CTA of the pair loads half of the B tile
    make_map(&mbl, g.b_lo, g.N, g.K, g.ldb, BN / 2);
    const int grid2 = 2 * std::min(m_tiles2 * p.n_tiles, sms / 2);
    c.timer.begin(KC_GEMM, c.stream);
    switch (g.epi) {
      case TC_EPI_BIAS_ACT_SPLIT:
        tc_gemm2_kernel<TC_EPI_BIAS_ACT_SPLIT><<<grid2, NTHREADS, SMEM2_BYTES, c.stream>>>(mah, mal, mbh, mbl, p);
        break;
      case TC_EPI_MULDACT_SPLIT:
        tc_gemm2_kernel<TC_EPI_MULDACT_SPLIT><<<grid2, NTHREADS, SMEM2_BYTES, c.stream>>>(mah, mal, mbh, mbl, p);
        break;
      default:
        tc_gemm2_kernel<TC_EPI_BIAS_ACT_DOT><<<grid2, NTHREADS, SMEM2_BYTES, c.stream>>>(mah, mal, mbh, mbl, p);
        break;
    }
    c.timer.end(c.stream);
    IK_CUDA(cudaGetLastError());
    c.count_launch(KC_GEMM, 2.0 * (double)g.M * (double)g.N * (double)g.K);
    if (c.timer.enabled) c.stats.gemm_mma_flops += 2.0 * p.nmma * (double)g.M * (double)g.N * (double)g.K;
    return 1;
  }
  const int total = p.m_tiles * p.n_tiles * p.splits;
  const int grid = std::min(total, sms);
  c.timer.begin(KC_GEMM, c.stream);
  switch (g.epi) {
    case TC_EPI_BIAS_ACT_SPLIT:
      tc_gemm_kernel<TC_EPI_BIAS_ACT_SPLIT><<<grid, NTHREADS, SMEM_BYTES, c.stream>>>(mah, mal, mbh, mbl, p);
      break;
    case TC_EPI_F32:
      tc_gemm_kernel<TC_EPI_F32><<<grid, NTHREADS, SMEM_BYTES, c.stream>>>(mah, mal, mbh, mbl, p);
      break;
    case TC_EPI_MULDACT_SPLIT:
      tc_gemm_kernel<TC_EPI_MULDACT_SPLIT><<<grid, NTHREADS, SMEM_BYTES, c.stream>>>(mah, mal, mbh, mbl, p);
      break;
    case TC_EPI_TAIL:
      tc_gemm_kernel<TC_EPI_TAIL><<<grid, NTHREADS, SMEM_BYTES, c.stream>>>(mah, mal, mbh, mbl, p);
      break;
    default:
      tc_gemm_kernel<TC_EPI_BIAS_ACT_DOT><<<grid, NTHREADS, SMEM_BYTES, c.stream>>>(mah, mal, mbh, mbl, p);
      break;
  }
  c.timer.end(c.stream);
  IK_CUDA(cudaGetLastError());
  c.count_launch(KC_GEMM, 2.0 * (double)g.M * (double)g.N * (double)g.K);
  if (c.timer.enabled) c.stats.gemm_mma_flops += 2.0 * p.nmma * (double)g.M * (double)g.N * (double)g.K;
  return p.splits;
}

// coordinates -> chi for a narrow net in one kernel (see koop_fused_kernel).  g describes the first-layer GEMM exactly
// like the TC_EPI_TAIL launch it replaces (b_hi/b_lo, ldb, N, K, bias, act, tail, chi_out); returns false if the
// shapes do not fit, in which case the caller materialises x_hat and uses launch_tc_gemm.
bool launch_koop_fused(Ctx &c, const float *coords, int64_t M, bool do_ln, const TcGemm &g) {
  const int A = c.tri_n;
  if (A < 2 || c.koop_fused_off || g.N > 128 || g.K != A * (A - 1) / 2 || g.K < 64 || M <= 0) return false;
  const int F = g.K, C = 3 * A;
  const int nmma_n = (g.N + 15) & ~15;
  const int bstage_al = (2 * nmma_n * 128 + 1023) & ~1023;
  const size_t smem = 1024 + 2 * KF_ASTAGE + 2 * (size_t)bstage_al + 128 + TAIL_FLOATS * 4 +
                      (4 * KF_ROWS + KF_ROWS) * sizeof(float2) + (size_t)(C + 3) * KF_CP * sizeof(float) + 16;
  if (smem > 227 * 1024) return false;
  // (i, j) of every 16th feature of the column-major strict upper triangle
  if (c.koop_start16.n == 0) {
    std::vector<short2> st;
    int f = 0;
    for (int j = 1; j < A; ++j)
      for (int i = 0; i < j; ++i, ++f)
        if ((f & 15) == 0) st.push_back(make_short2((short)i, (short)j));
    st.push_back(make_short2(0, (short)A));
    c.koop_start16.ensure(st.size());
    IK_CUDA(cudaMemcpy(c.koop_start16.p, st.data(), st.size() * sizeof(short2), cudaMemcpyHostToDevice));
  }
  if (c.attr_needed(Ctx::ATTR_KOOPF))
    IK_CUDA(cudaFuncSetAttribute(koop_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  alignas(64) CUtensorMap mbh, mbl;
  make_map(&mbh, g.b_hi, g.N, g.K, g.ldb, nmma_n);
  make_map(&mbl, g.b_lo, g.N, g.K, g.ldb, nmma_n);
  KoopFusedP p{};
  p.coords = coords; p.cmap = c.tri_cmap.p; p.M = M;
  p.D = c.D; p.A = A; p.F = F; p.C = C;
  p.nkb = cdiv(F, BK); p.nmma_n = nmma_n;
  p.do_ln = do_ln ? 1 : 0;
  p.eps2 = c.cfg.ln_eps * c.cfg.ln_eps;
  p.start16 = c.koop_start16.p;
  p.ep.M = (int)M; p.ep.N = g.N; p.ep.K = g.K;
  p.ep.epi = TC_EPI_TAIL; p.ep.act = g.act; p.ep.bias = g.bias;
  p.ep.tail = g.tail; p.ep.chi_out = g.chi_out;
  const int64_t ntiles = (M + KF_ROWS - 1) / KF_ROWS;
  const int grid = (int)std::min<int64_t>(ntiles, c.num_sms);
  c.timer.begin(KC_GEMM, c.stream);
  koop_fused_kernel<<<grid, KF_THREADS, smem, c.stream>>>(mbh, mbl, p);
  c.timer.end(c.stream);
  IK_CUDA(cudaGetLastError());
  c.count_launch(KC_GEMM, 2.0 * (double)M * (double)g.N * (double)g.K);
  c.stats.n_featurize_launches++;
  if (c.timer.enabled) {
    c.stats.gemm_mma_flops += 6.0 * (double)M * (double)g.N * (double)g.K;
    c.stats.featurize_bytes += 4.0 * (double)c.D * (double)M;
  }
  return true;
}

}  // namespace ik
