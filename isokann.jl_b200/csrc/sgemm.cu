// FP32 CUDA-core GEMM with fused epilogues for the Dense layers of narrow nets (default
// pairnets, smallnet) and as the exact-fp32 engine of every layer when gemm_mode == FP32.
//
// Replaces Flux.Dense forward `sigma.(W*x .+ b)` and its Zygote pullbacks (call sites:
// reference src/isotarget.jl:18, src/iso.jl:185,203), which the reference runs as cuBLAS/
// OpenBLAS SGEMM + one broadcast kernel per layer.  Here bias and activation are part of
// the GEMM:
//   * the bias is an augmented row: the flat parameter segment [W(in x out); b(out)] is a
//     row-major (in+1) x out matrix, and the A operand supplies a 1 at k == in;
//   * the weight gradient uses the same trick on the output rows, so dW and db land in the
//     flat gradient vector in one pass;
//   * the backward data GEMM multiplies by the activation derivative in its epilogue.
// C(i,j) = sum_k A(i,k) * B(k,j), 16x16 threads, (BM/16)x(BN/16) micro-tiles laid out as 4x4
// sub-blocks 64 apart so shared-memory reads are conflict-free float4s.
#include "common.cuh"

namespace ik {

__device__ __forceinline__ float act_apply(float a, int kind) {
  switch (kind) {
    case ISOKANN_ACT_SIGMOID: return 1.0f / (1.0f + __expf(-a));
    case ISOKANN_ACT_TANH: return tanhf(a);
    case ISOKANN_ACT_RELU: return fmaxf(a, 0.f);
    default: return a;
  }
}
__device__ __forceinline__ float dact_from_out(float z, int kind) {
  switch (kind) {
    case ISOKANN_ACT_SIGMOID: return z * (1.0f - z);
    case ISOKANN_ACT_TANH: return 1.0f - z * z;
    case ISOKANN_ACT_RELU: return z > 0.f ? 1.0f : 0.f;
    default: return 1.0f;
  }
}

template <int BM, int BN, bool A_KC, bool B_JC>
__global__ void __launch_bounds__(256) sgemm_kernel(GemmP p) {
  constexpr int BK = 16;
  constexpr int RM = BM / 64, RN = BN / 64;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t i0 = (int64_t)blockIdx.x * BM;
  const int j0 = blockIdx.y * BN;
  const int kb = blockIdx.z * p.kchunk;
  const int ke = min(p.K, kb + p.kchunk);

  float acc[RM * 4][RN * 4];
#pragma unroll
  for (int a = 0; a < RM * 4; ++a)
#pragma unroll
    for (int b = 0; b < RN * 4; ++b) acc[a][b] = 0.f;

  for (int k0 = kb; k0 < ke; k0 += BK) {
    // ---- A tile (BM x BK) ----
    if (A_KC) {  // A(i,k) = A[i*lda + k]
      const int kk = tid & 15, r = tid >> 4;
      const int k = k0 + kk;
#pragma unroll
      for (int it = 0; it < BM / 16; ++it) {
        const int64_t i = i0 + r + 16 * it;
        float v = 0.f;
        if (i < p.M && k < ke) v = (k == p.ones_k || i == p.ones_i) ? 1.0f : __ldg(p.A + i * p.lda + k);
        As[kk][r + 16 * it] = v;
      }
    } else {  // A(i,k) = A[k*lda + i]
      const int ii = tid % BM, kq = tid / BM;
      constexpr int KSTEP = 256 / BM;
      const int64_t i = i0 + ii;
#pragma unroll
      for (int it = 0; it < BK / KSTEP; ++it) {
        const int kk = kq + KSTEP * it;
        const int k = k0 + kk;
        float v = 0.f;
        if (i < p.M && k < ke) v = (k == p.ones_k || i == p.ones_i) ? 1.0f : __ldg(p.A + (int64_t)k * p.lda + i);
        As[kk][ii] = v;
      }
    }
    // ---- B tile (BK x BN) ----
    if (B_JC) {  // B(k,j) = B[k*ldb + j]
      const int jj = tid % BN, kq = tid / BN;
      constexpr int KSTEP = 256 / BN;
      const int j = j0 + jj;
#pragma unroll
      for (int it = 0; it < BK / KSTEP; ++it) {
        const int kk = kq + KSTEP * it;
        const int k = k0 + kk;
        float v = 0.f;
        if (j < p.N && k < ke) v = __ldg(p.B + (int64_t)k * p.ldb + j);
        Bs[kk][jj] = v;
      }
    } else {  // B(k,j) = B[j*ldb + k]
      const int kk = tid & 15, r = tid >> 4;
      const int k = k0 + kk;
#pragma unroll
      for (int it = 0; it < BN / 16; ++it) {
        const int j = j0 + r + 16 * it;
        float v = 0.f;
        if (j < p.N && k < ke) v = __ldg(p.B + (int64_t)j * p.ldb + k);
        Bs[kk][r + 16 * it] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[RM * 4], b[RN * 4];
#pragma unroll
      for (int rb = 0; rb < RM; ++rb) {
        const float4 t = *reinterpret_cast<const float4 *>(&As[kk][rb * 64 + ty * 4]);
        a[rb * 4 + 0] = t.x; a[rb * 4 + 1] = t.y; a[rb * 4 + 2] = t.z; a[rb * 4 + 3] = t.w;
      }
#pragma unroll
      for (int cb = 0; cb < RN; ++cb) {
        const float4 t = *reinterpret_cast<const float4 *>(&Bs[kk][cb * 64 + tx * 4]);
        b[cb * 4 + 0] = t.x; b[cb * 4 + 1] = t.y; b[cb * 4 + 2] = t.z; b[cb * 4 + 3] = t.w;
      }
#pragma unroll
      for (int x = 0; x < RM * 4; ++x)
#pragma unroll
        for (int y = 0; y < RN * 4; ++y) acc[x][y] = fmaf(a[x], b[y], acc[x][y]);
    }
    __syncthreads();
  }

  float *Cz = p.C;
  if (p.epi == EPI_PARTIAL) Cz += (int64_t)blockIdx.z * (int64_t)p.M * (int64_t)p.ldc;
#pragma unroll
  for (int x = 0; x < RM * 4; ++x) {
    const int64_t i = i0 + (x >> 2) * 64 + ty * 4 + (x & 3);
    if (i >= p.M) continue;
#pragma unroll
    for (int cb = 0; cb < RN; ++cb) {
      const int j = j0 + cb * 64 + tx * 4;
#pragma unroll
      for (int y = 0; y < 4; ++y) {
        if (j + y >= p.N) continue;
        float v = acc[x][cb * 4 + y];
        if (p.epi == EPI_ACT) v = act_apply(v, p.act);
        else if (p.epi == EPI_MULDACT) v *= dact_from_out(__ldg(p.Z + i * p.ldz + j + y), p.act);
        Cz[i * p.ldc + j + y] = v;
      }
    }
  }
}

template <int BM, int BN>
static void launch_variant(const GemmP &p, bool akc, bool bjc, dim3 grid, cudaStream_t s) {
  if (akc && bjc) sgemm_kernel<BM, BN, true, true><<<grid, 256, 0, s>>>(p);
  else if (akc && !bjc) sgemm_kernel<BM, BN, true, false><<<grid, 256, 0, s>>>(p);
  else if (!akc && bjc) sgemm_kernel<BM, BN, false, true><<<grid, 256, 0, s>>>(p);
  else sgemm_kernel<BM, BN, false, false><<<grid, 256, 0, s>>>(p);
}

int launch_gemm(Ctx &c, const GemmP &p_in, bool a_kcontig, bool b_jcontig, int splits) {
  GemmP p = p_in;
  if (p.M <= 0 || p.N <= 0) return 0;
  if (splits < 1) splits = 1;
  int kchunk = ((p.K + splits - 1) / splits + 15) & ~15;
  if (kchunk < 16) kchunk = 16;
  splits = (p.K + kchunk - 1) / kchunk;
  p.kchunk = kchunk;
  if (p.epi == EPI_PARTIAL) IK_REQUIRE(p.ldc == p.N, ISOKANN_BAD_ARGUMENT, "split-K partials must be dense");
  else IK_REQUIRE(splits == 1, ISOKANN_BAD_ARGUMENT, "split-K requires the partial epilogue");
  const bool wideN = p.N > 64;
  const bool tallM = p.M > 64;
  c.timer.begin(KC_GEMM, c.stream);
  if (wideN && tallM) {
    dim3 grid(cdiv(p.M, 128), cdiv(p.N, 128), splits);
    launch_variant<128, 128>(p, a_kcontig, b_jcontig, grid, c.stream);
  } else if (tallM) {
    dim3 grid(cdiv(p.M, 128), cdiv(p.N, 64), splits);
    launch_variant<128, 64>(p, a_kcontig, b_jcontig, grid, c.stream);
  } else if (wideN) {
    dim3 grid(cdiv(p.M, 64), cdiv(p.N, 128), splits);
    launch_variant<64, 128>(p, a_kcontig, b_jcontig, grid, c.stream);
  } else {
    dim3 grid(cdiv(p.M, 64), cdiv(p.N, 64), splits);
    launch_variant<64, 64>(p, a_kcontig, b_jcontig, grid, c.stream);
  }
  c.timer.end(c.stream);
  IK_CUDA(cudaGetLastError());
  c.count_launch(KC_GEMM, 2.0 * (double)p.M * (double)p.N * (double)p.K);
  return splits;
}

__global__ void splitk_reduce_kernel(const float *__restrict__ partials, int splits, int64_t count,
                                     float *__restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    float s = partials[i];
    for (int z = 1; z < splits; ++z) s += partials[(int64_t)z * count + i];
    out[i] = s;
  }
}

void launch_splitk_reduce(Ctx &c, const float *partials, int splits, int64_t count, float *out) {
  int grid = (int)((count + 255) / 256);
  if (grid > c.num_sms * 8) grid = c.num_sms * 8;
  c.timer.begin(KC_REDUCE, c.stream);
  splitk_reduce_kernel<<<grid, 256, 0, c.stream>>>(partials, splits, count, out);
  c.timer.end(c.stream);
  IK_CUDA(cudaGetLastError());
  c.count_launch(KC_REDUCE);
}

}  // namespace ik
