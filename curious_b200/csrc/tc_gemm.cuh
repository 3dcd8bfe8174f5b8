// Host interface of the tensor-core layer GEMM (tc_gemm.cu): a batch of independent problems -> one launch.
#pragma once
#include "mlp_kernels.cuh"

namespace cur {

constexpr int TC_MAX_PROBS = 16;
constexpr int TC_COLSUM_ROWS = 256;      // rows per partial row reduction

struct TcReduce {
  const float* part; float* out;
  int64_t count, stride;                 // floats per slice (multiple of 4), distance between slices
  int splits, block_begin, vec;
};
struct TcReduceBatch {
  TcReduce p[2 * TC_MAX_PROBS];
  int n;
};

bool tc_supported(const GemmProb& p);     // shape / alignment fit the tcgen05 kernel
int tc_pick_splits(const GemmProb& p);
int64_t tc_partial_floats(const GemmProb& p);   // floats of split-K workspace the problem needs (0: none)
int64_t tc_rowred_partial_floats(int64_t rows, int M, int NJ);
bool tc_skinny_supported(const GemmProb& p);

struct TcLauncher {
  struct { int n, total_tiles; } G;
  TcReduceBatch R;
  struct RowRedDesc { const float* X; int64_t ldx; const float* Y; int64_t ldy; int64_t rows; int M, NJ, chunks; float* part; }
      rowred[TC_MAX_PROBS];
  int n_rowred;
  struct SkinnyDesc { GemmProb p; } skinny[TC_MAX_PROBS];
  int n_skinny;
  bool pair;                 // launch the CTA-pair (cta_group::2) kernel: every problem of the batch needs M % 256 == 0
  int level;                 // flushes so far; partial workspaces alternate between two sets (level & 1)
  bool red_pending[2];       // deferred reductions reading partial set 0 / 1 are still in flight on the side stream
  alignas(64) unsigned char storage[TC_MAX_PROBS * (4 * 128 + 128) + 64];   // TcBatch (tensor maps + problems)
  TcLauncher();
  int add(const GemmProb& p, float* partial);
  // out[m][j] = sum_r X[r][m] * Y[r][j] (j < NJ <= 4), or the column sums of X when Y == NULL;
  // partial: tc_rowred_partial_floats(rows, M, NJ) floats
  int add_rowred(const float* X, int64_t ldx, int M, const float* Y, int64_t ldy, int NJ, int64_t rows, float* out,
                 float* partial);
  // bandwidth-bound problems with N <= 4 (output layers) or K <= 4 (backward through them), see tc_skinny_kernel
  int add_skinny(const GemmProb& p);
  int flush(cudaStream_t s);
  int finish(cudaStream_t s);        // join the deferred reductions (call once after the last flush)
  bool empty() const { return G.n == 0 && R.n == 0 && n_rowred == 0 && n_skinny == 0; }
};

}  // namespace cur
