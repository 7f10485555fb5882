// Gradient all-reduce over NVLink peer memory (one process per GPU, CUDA IPC), replacing ncclAllReduce on the
// training step of a single NVSwitch node.  Not in the reference (single device); it is the exchange step
// of SURVEY 8(e)(2).
//
// Two-shot, in place, on the flat gradient range [lo, hi):
//   ready   every rank tells every peer that its gradient is complete (one flag store per peer, release.sys)
//   reduce  rank r owns the r-th slice: it reads that slice from all W ranks over NVLink, adds them in rank order
//           0..W-1 (every element is summed exactly once, by its owner, in a fixed order: all ranks end up with
//           bit-identical values, whatever the timing), and
//   push    stores the sum into all W gradient buffers (its own included)
//   done    the last CTA of every rank publishes "my pushes have landed"; a rank leaves the kernel only when all W
//           ranks have, so the kernels that follow in stream order (optimiser) see the complete result.
// Peers never touch the same words in the same phase (a rank only writes its own slice everywhere and only reads
// other ranks' copies of its own slice), so the exchange needs no staging buffer.  Flags are monotonically increasing
// sequence numbers kept in device memory, which makes the kernel replayable from a captured CUDA graph.
// At 16.8 MB over 8 GPUs NCCL 2.28 needed 210-550 us per call (SM-limited to share the GPU with the GEMMs of the
// backward pass); each rank here moves 2 x 7/8 x 1/8 of the range over its links.
// Every wait is a bounded spin that traps instead of hanging the GPU.
#include "common.cuh"

namespace ik {

namespace {

constexpr int kP2PThreads = 512;

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer4(const float *p) {  // straight from the owner's L2, never a stale L1 line
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_peer1(const float *p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void wait_flag(const uint32_t *p, uint32_t target) {
  long long t0 = 0;
  uint32_t spins = 0;
  // (int32_t)(v - target) >= 0: robust against wrap-around of the sequence number
  while ((int32_t)(ld_acquire_sys(p) - target) < 0) {
    if ((++spins & 255u) == 0) {
      if (t0 == 0) t0 = clock64();
      else if (clock64() - t0 > 20000000000LL) __trap();  // ~10 s: a peer died or the ranks diverged
    }
    __nanosleep(64);
  }
}

struct P2PArgs {
  float *grads[ISOKANN_MAX_RANKS];      // gradient buffer of every rank (own entry = local pointer)
  uint32_t *flags[ISOKANN_MAX_RANKS];   // flag block of every rank: [0..W) ready, [W..2W) done, indexed by SOURCE rank
  uint32_t *seq;                        // local: sequence number of the last completed exchange
  unsigned int *ticket;                 // local: CTAs that finished their slice
  int64_t lo, hi;
  int rank, world;
};

// The peer loads are cp.async copies into shared memory (16 bytes each, no destination register, so nothing limits how
// many are in flight): every thread fetches the W copies of its element, two elements deep (double-buffered), and
// consumes only what it fetched itself, so a cp.async.wait_group is all the synchronisation the pipeline needs.  An
// earlier version loaded into registers: ptxas kept 2-3 loads in flight per thread and 16.8 MB on 8 GPUs took 355 us.
// WT: world size at compile time (2, 4, 8) or 0 for any world size.
template <int WT>
__global__ void __launch_bounds__(kP2PThreads) p2p_allreduce_kernel(P2PArgs a) {
  extern __shared__ __align__(16) float4 stage[];  // [2][W][kP2PThreads]
  __shared__ bool is_last;
  const int W = WT ? WT : a.world;
  const uint32_t seq = *((volatile uint32_t *)a.seq) + 1u;
  // ---- ready: the kernels before this one in stream order produced the local gradient
  if (blockIdx.x == 0 && threadIdx.x < W) {
    __threadfence_system();
    st_release_sys(a.flags[threadIdx.x] + a.rank, seq);
  }
  if (threadIdx.x < W) wait_flag(a.flags[a.rank] + threadIdx.x, seq);
  __syncthreads();
  // ---- reduce + push the slice this rank owns.  Slices are cut on 4-element boundaries of the global index so
  // that the body is float4; the unaligned head and tail of the range go to rank 0, element by element.
  const int64_t lo4 = (a.lo + 3) & ~(int64_t)3, hi4 = a.hi & ~(int64_t)3;
  if (hi4 > lo4) {
    const int64_t nvec = (hi4 - lo4) >> 2;
    const int64_t per = (nvec + W - 1) / W;
    const int64_t v0 = min(nvec, per * a.rank), v1 = min(nvec, v0 + per);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(stage) + 16u * threadIdx.x;
    auto fetch = [&](int64_t v, int buf) {
      if (v < v1) {
        const int64_t i = lo4 + (v << 2);
#pragma unroll
        for (int p = 0; p < (WT ? WT : ISOKANN_MAX_RANKS); ++p)
          if (p < W)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + 16u * kP2PThreads * (buf * W + p)),
                         "l"(a.grads[p] + i)
                         : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int64_t v = v0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int buf = 0;
    fetch(v, 0);
    for (; v < v1; v += stride, buf ^= 1) {
      fetch(v + stride, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
      const float4 *x = stage + (size_t)kP2PThreads * (buf * W) + threadIdx.x;
      float4 s = x[0];
#pragma unroll
      for (int p = 1; p < (WT ? WT : ISOKANN_MAX_RANKS); ++p)
        if (p < W) {  // summed in rank order
          const float4 y = x[(size_t)kP2PThreads * p];
          s.x += y.x; s.y += y.y; s.z += y.z; s.w += y.w;
        }
      const int64_t i = lo4 + (v << 2);
#pragma unroll
      for (int p = 0; p < (WT ? WT : ISOKANN_MAX_RANKS); ++p)
        if (p < W) *reinterpret_cast<float4 *>(a.grads[p] + i) = s;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  }
  if (a.rank == 0 && blockIdx.x == 0) {
    const int64_t head = min(lo4, a.hi) - a.lo, tail = a.hi - max(hi4, min(lo4, a.hi));
    for (int64_t t = threadIdx.x; t < head + tail; t += blockDim.x) {
      const int64_t i = t < head ? a.lo + t : max(hi4, min(lo4, a.hi)) + (t - head);
      float s = ld_peer1(a.grads[0] + i);
      for (int p = 1; p < W; ++p) s += ld_peer1(a.grads[p] + i);
      for (int p = 0; p < W; ++p) a.grads[p][i] = s;
    }
  }
  // ---- done: the last CTA of this rank publishes, then waits for every rank
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(a.ticket, 1u);
    is_last = prev == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  if (threadIdx.x < W) {
    __threadfence_system();
    st_release_sys(a.flags[threadIdx.x] + W + a.rank, seq);
    wait_flag(a.flags[a.rank] + W + threadIdx.x, seq);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    *a.ticket = 0u;
    *a.seq = seq;
    __threadfence();
  }
}

}  // namespace

void launch_p2p_allreduce(Ctx &c, int64_t lo, int64_t hi) {
  if (hi <= lo) return;
  P2PArgs a{};
  for (int p = 0; p < c.world; ++p) {
    a.grads[p] = c.p2p.grads[p];
    a.flags[p] = c.p2p.flags[p];
  }
  a.seq = c.p2p.seq.p;
  a.ticket = c.p2p.ticket.p;
  a.lo = lo; a.hi = hi;
  a.rank = c.rank; a.world = c.world;
  // every CTA spins on peer flags, so all of them must be resident together with whatever else runs: as many CTAs as
  // the overlapped GEMMs leave SMs free
  const int64_t nvec = (hi - lo) / 4;
  int grid = (int)std::max<int64_t>(1, std::min<int64_t>(c.p2p.ctas, (nvec / c.world + kP2PThreads - 1) / kP2PThreads));
  const size_t smem = (size_t)2 * c.world * kP2PThreads * sizeof(float4);
  if (c.attr_needed(Ctx::ATTR_P2P)) {
    IK_CUDA(cudaFuncSetAttribute(p2p_allreduce_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    IK_CUDA(cudaFuncSetAttribute(p2p_allreduce_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    IK_CUDA(cudaFuncSetAttribute(p2p_allreduce_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    IK_CUDA(cudaFuncSetAttribute(p2p_allreduce_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
  }
  IK_REQUIRE(smem <= 208 * 1024, ISOKANN_ERR_STATE, "peer-memory exchange: world size too large for the staging buffer");
  c.timer.begin(KC_NCCL, c.stream);
  switch (c.world) {
    case 2: p2p_allreduce_kernel<2><<<grid, kP2PThreads, smem, c.stream>>>(a); break;
    case 4: p2p_allreduce_kernel<4><<<grid, kP2PThreads, smem, c.stream>>>(a); break;
    case 8: p2p_allreduce_kernel<8><<<grid, kP2PThreads, smem, c.stream>>>(a); break;
    default: p2p_allreduce_kernel<0><<<grid, kP2PThreads, smem, c.stream>>>(a); break;
  }
  c.timer.end(c.stream);
  IK_CUDA(cudaGetLastError());
  c.count_launch(KC_REDUCE);
  c.stats.p2p_exchanges++;
}

}  // namespace ik
