// Host interface of the tensor-core layer GEMM (tc_gemm.cu): a batch of independent problems -> one launch.
#pragma once
#include "mlp_kernels.cuh"

namespace cur {

constexpr int TC_MAX_PROBS = 16;
constexpr int TC_COLSUM_ROWS = 512;      // rows per partial column sum

struct TcReduce {
  const float* part; float* out;
  int64_t count, stride;                 // floats per slice (multiple of 4), distance between slices
  int splits, block_begin;
};
struct TcReduceBatch {
  TcReduce p[2 * TC_MAX_PROBS];
  int n;
};

bool tc_supported(const GemmProb& p);     // shape / alignment fit the tcgen05 kernel
int tc_pick_splits(const GemmProb& p);
int64_t tc_partial_floats(const GemmProb& p);   // floats of split-K workspace the problem needs (0: none)

struct TcLauncher {
  struct { int n, total_tiles; } G;
  TcReduceBatch R;
  struct ColSum { const float* B; int64_t ldb, rows; int N, chunks; float* part; } colsum[TC_MAX_PROBS];
  int n_colsum;
  alignas(64) unsigned char storage[TC_MAX_PROBS * (2 * 128 + 128) + 64];   // TcBatch (tensor maps + problems)
  TcLauncher();
  int add(const GemmProb& p, float* partial);
  // out[n] = sum over rows of B[rows][N] (bias gradient); partial: ceil(rows / TC_COLSUM_ROWS) * N floats
  int add_colsum(const float* B, int64_t ldb, int64_t rows, int N, float* out, float* partial);
  int flush(cudaStream_t s);
  bool empty() const { return G.n == 0 && R.n == 0 && n_colsum == 0; }
};

}  // namespace cur
