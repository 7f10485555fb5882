// Tensor-core (tcgen05) path: split-bf16 operand buffers and the kernels around the GEMM.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace ik {

enum { TC_EPI_BIAS_ACT_SPLIT = 0, TC_EPI_F32 = 1, TC_EPI_MULDACT_SPLIT = 2, TC_EPI_BIAS_ACT_DOT = 3, TC_EPI_TAIL = 4 };

// the Dense layers after the tensor-core layer of a narrow net (e.g. 71 -> 8 -> 1), evaluated per row in
// the GEMM epilogue: w[0] = GEMM N (<= 128), w[1..nl] <= 16; seg[i] = row-major (w[i]+1) x w[i+1] [W; b]
struct TcTail {
  int nl;
  int w[4];
  const float *seg[3];
  int act, last_act;
};

// D[M x N] = A[M x K] * B[N x K]^T with A, B given as bf16 (hi, lo) pairs, K contiguous
struct TcGemm {
  const __nv_bfloat16 *a_hi, *a_lo;
  int64_t lda;
  const __nv_bfloat16 *b_hi, *b_lo;
  int64_t ldb;
  int M, N, K;
  int mn_major;                       // 1: A is [K x M] and B is [K x N] in memory (M / N contiguous): weight gradients
  int ones_col;                       // split epilogues: write 1.0 into column N (bias column of the next wgrad)
  int fmt;                            // 0: operands are bf16 (hi, lo); 1: fp16 (hi, lo) -- split outputs use the same format
  int nmma;                           // 3 (default): hi*hi + hi*lo + lo*hi; 2: hi*hi + lo*hi (B rounded once; b_lo unused)
  int epi, act;
  const float *bias;                  // TC_EPI_BIAS_ACT_SPLIT
  __nv_bfloat16 *out_hi, *out_lo;     // split outputs
  int64_t ldo;
  float *out_f32;                     // TC_EPI_F32 (split-K slice s at out_f32 + s*M*ldc)
  int64_t ldc;
  const __nv_bfloat16 *z_hi, *z_lo;   // TC_EPI_MULDACT_SPLIT
  int64_t ldz;
  const float *w_last;                // TC_EPI_BIAS_ACT_DOT: fused last (thin) layer, weights [N x d]
  float *dot_out;                     // partial chi [M x 2*ceil(N/256) x d]
  int d;
  TcTail tail;                        // TC_EPI_TAIL
  float *chi_out;                     // TC_EPI_TAIL: [M x tail.w[tail.nl]]
  int splits;
};

int launch_tc_gemm(Ctx &c, const TcGemm &g);  // returns the number of split-K slices written
// narrow nets: featurizer + LayerNorm + first layer + tail in one kernel, from coordinate records (tc_gemm.cu)
bool launch_koop_fused(Ctx &c, const float *coords, int64_t M, bool do_ln, const TcGemm &g);

// a split-bf16 matrix: hi/lo planes, rows x ld
struct SplitBuf {
  DevBuf<__nv_bfloat16> hi, lo;
  int64_t rows = 0, ld = 0;
  void ensure(int64_t r, int64_t l) {
    if (r * l > (int64_t)hi.n) {
      hi.ensure((size_t)(r * l));
      lo.ensure((size_t)(r * l));
    }
    rows = r;
    ld = l;
  }
  void release() {
    hi.release();
    lo.release();
  }
};

struct TcState {
  std::vector<SplitBuf> act;   // act[l], l = 0..L-1: row-major rows x wp_l
  std::vector<SplitBuf> wF;    // forward operand of layer l (l = 0..L-2): [w_{l+1} x wp_l]
  std::vector<SplitBuf> wD;    // dgrad operand of layer l   (l = 1..L-2): [w_l x wp_{l+1}]
  std::vector<SplitBuf> wF_alt, wD_alt;  // second operand set: refreshed on the communication stream, then swapped in
  std::vector<DevBuf<__nv_bfloat16>> wF16;  // fp16 copy of the forward operand (inference forward with 2 MMAs)
  SplitBuf delta[2];           // row-major delta ping-pong
  SplitBuf x_alt;              // second x_hat buffer: the featurizer of chunk i+1 overlaps the GEMMs of chunk i
  cudaStream_t feat_stream = nullptr;
  cudaEvent_t feat_done[2] = {nullptr, nullptr}, x_free[2] = {nullptr, nullptr}, koop_start = nullptr;
  SplitBuf dlast;              // split copy of the last layer's delta (B x d, padded to 64 columns)
  DevBuf<float> dot_partial;   // fused last layer: partial chi per 128-column slot
  std::vector<int> wp;         // padded widths
  int64_t rows = 0, train_rows = 0;
};

// featurizer + LayerNorm writing the split-bf16 A operand [M x ld] (pad columns zeroed)
// fmt: 0 bf16 pairs, 1 fp16 pairs (same 16-bit buffers)
void launch_featurize_split(Ctx &c, const float *coords, const int64_t *gather, int64_t gather_off, int64_t M,
                            bool pairs, bool do_ln, __nv_bfloat16 *out_hi, __nv_bfloat16 *out_lo, int64_t ld,
                            int fmt = 0);  // column F := 1
// forward operand of a Dense layer rounded once to fp16: Wf16[out x ld_f] (K = in contiguous), for the 2-MMA
// inference forward
void launch_prep_weights_f16(Ctx &c, const float *seg, int fin, int fout, __nv_bfloat16 *wf16, int64_t ld_f);
// fp32 [rows x cols] (dense) -> split bf16 [rows x ld], pad columns zeroed
void launch_f32_to_split(Ctx &c, const float *in, int64_t rows, int cols, __nv_bfloat16 *hi, __nv_bfloat16 *lo,
                         int64_t ld);
// set column `col` of a split matrix to (hi, lo) = (1, 0) for every row
void launch_set_ones_col(Ctx &c, __nv_bfloat16 *hi, __nv_bfloat16 *lo, int64_t rows, int64_t ld, int col);
// weights of one Dense layer: flat fp32 segment seg[(in) x out] (row-major) ->
//   fwd operand  Wf[out x ld_f] (K = in contiguous) and dgrad operand Wd[in x ld_d] (K = out contiguous)
void launch_prep_weights(Ctx &c, const float *seg, int fin, int fout, __nv_bfloat16 *wf_hi, __nv_bfloat16 *wf_lo,
                         int64_t ld_f, __nv_bfloat16 *wd_hi, __nv_bfloat16 *wd_lo, int64_t ld_d);
// last (thin) Dense layer from a split activation: chi[m, a] = act(sum_k z[m,k] W[k,a] + b[a])
void launch_thin_forward(Ctx &c, const __nv_bfloat16 *z_hi, const __nv_bfloat16 *z_lo, int64_t M, int fin, int64_t ldz,
                         const float *seg, int d, int act, float *chi);
// delta_prev[m,k] = (sum_a delta[m,a] W[k,a]) * act'(z[m,k]) written split (row-major)
bool thin_head_eligible(const Ctx &c);
void launch_thin_head(Ctx &c, const __nv_bfloat16 *z_hi, const __nv_bfloat16 *z_lo, int64_t M, int fin, int64_t ldz,
                      const float *seg, const float *target, const int64_t *idx, const float *wl, double Bglobal,
                      float *chi, float *delta, __nv_bfloat16 *dl_hi, __nv_bfloat16 *dl_lo, int ld_dl,
                      __nv_bfloat16 *out_hi, __nv_bfloat16 *out_lo, int64_t ldo, double *partials,
                      unsigned int *ticket, float *packed_tail);
void launch_thin_dgrad(Ctx &c, const float *delta, int64_t M, int d, const float *seg, int fin,
                       const __nv_bfloat16 *z_hi, const __nv_bfloat16 *z_lo, int64_t ldz, int act,
                       __nv_bfloat16 *out_hi, __nv_bfloat16 *out_lo, int64_t ldo);
// chi[m, a] = act(sum_s partial[m, s, a] + b[a]) -- finishes the fused last layer
void launch_dot_finish(Ctx &c, const float *partial, int64_t M, int slots, int d, const float *bias, int act,
                       float *chi);

}  // namespace ik
