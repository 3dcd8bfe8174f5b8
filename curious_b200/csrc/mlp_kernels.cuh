// Grouped small-GEMM kernel for the DDPG actor/critic MLPs (sm_100a, fp32 CUDA cores).
//
// At the reference batch (256 rows, 256 hidden units) one layer is a 256x256x256 GEMM: 17 MFLOP.
// That is far below what fills tensor-core tiles on 148 SMs and the parity target (rel 1e-5 vs an
// fp32 restatement of the TF graph) rules out single-pass TF32/BF16, so the layer GEMMs
// run on FFMA.  What matters is latency: every launch of this kernel executes ALL independent GEMMs
// of one dependency level of the DDPG graph (e.g. main.pi / target.pi / main.Q layer k) as one grid
// ("grouped GEMM"), 32x32 output tiles, K split 4 ways inside the CTA (8 warps) and reduced
// deterministically through shared memory.
//
// One problem:  C[M,N] = epi( opA(A)[M,K(+1)] * opB(B)[K(+1),N]  (+ A2[M,K2] * B2[K2,N]) )
//   bias   : added per output column before the activation (forward layers)
//   a_trans: A is stored [K][M] (dW = X^T dY); with a_ones the LAST output row (m == M-1) is the
//            column sum of B, i.e. the db row of the contiguous [dW;db] block of the flat gradient
//   b_trans: B is stored [N][K] (dX = dY W^T)
#pragma once
#include "common.cuh"

namespace cur {

enum { EPI_NONE = 0, EPI_RELU = 1, EPI_RELU_MASK = 2, EPI_TANH = 3, EPI_ACTOR_DY = 4 };

struct GemmProb {
  const float* A;  int lda;
  const float* B;  int ldb;
  const float* A2; int lda2;   // optional second K segment (plain NN), K2 == 0 if unused
  const float* B2; int ldb2;
  float* C;  int ldc;
  float* C2; int ldc2;         // optional second destination: C2 = scale2 * value
  const float* bias;           // optional [N]
  const float* aux; int ldaux; // RELU_MASK: activation whose sign gates the gradient; ACTOR_DY: tanh output
  int M, N, K, K2;
  int a_trans, b_trans, a_ones;
  int epi;
  float scale2, coef;
  int tiles_n, tile_begin;     // filled by the launcher
};

constexpr int GEMM_MAX_PROBS = 8;
constexpr int GT = 32;          // tile edge
constexpr int GK = 64;          // K chunk per stage (4 k-groups x 16)
constexpr int GEMM_THREADS = 256;

struct GemmBatch {
  GemmProb p[GEMM_MAX_PROBS];
  int n;
  int total_tiles;
};

__global__ void __launch_bounds__(GEMM_THREADS) grouped_gemm_kernel(const __grid_constant__ GemmBatch G) {
  __shared__ __align__(16) float As[GK][GT + 4];
  __shared__ __align__(16) float Bs[GK][GT + 4];

  // which problem / tile
  int pi = 0;
#pragma unroll 1
  while (pi + 1 < G.n && (int)blockIdx.x >= G.p[pi + 1].tile_begin) ++pi;
  const GemmProb& P = G.p[pi];
  const int tile = blockIdx.x - P.tile_begin;
  const int tm = tile / P.tiles_n, tn = tile - tm * P.tiles_n;
  const int m0 = tm * GT, n0 = tn * GT;

  const int tid = threadIdx.x;
  const int kg = tid >> 6;          // k-group 0..3
  const int lt = tid & 63;
  const int ty = lt >> 3, tx = lt & 7;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int Ktot1 = P.K;

  for (int seg = 0; seg < 2; ++seg) {
    const int Kseg = (seg == 0) ? Ktot1 : P.K2;
    if (Kseg <= 0) continue;
    for (int k0 = 0; k0 < Kseg; k0 += GK) {
      // ---- stage tiles (coalesced along the contiguous global dimension)
      if (seg == 0) {
        if (!P.a_trans) {
          // A[m][k], k contiguous: thread -> (m = i / 64, k = i % 64)
          for (int i = tid; i < GT * GK; i += GEMM_THREADS) {
            int m = i >> 6, k = i & 63;
            int gm = m0 + m, gk = k0 + k;
            float v = 0.f;
            if (gm < P.M && gk < P.K) v = P.A[(int64_t)gm * P.lda + gk];
            As[k][m] = v;
          }
        } else {
          // A stored [k][m] (+ implicit ones ROW at m == M-1 when a_ones): m contiguous
          for (int i = tid; i < GT * GK; i += GEMM_THREADS) {
            int k = i >> 5, m = i & 31;
            int gm = m0 + m, gk = k0 + k;
            float v = 0.f;
            if (gk < P.K && gm < P.M) {
              if (P.a_ones && gm == P.M - 1) v = 1.f;
              else v = P.A[(int64_t)gk * P.lda + gm];
            }
            As[k][m] = v;
          }
        }
        if (!P.b_trans) {
          for (int i = tid; i < GT * GK; i += GEMM_THREADS) {
            int k = i >> 5, n = i & 31;
            int gn = n0 + n, gk = k0 + k;
            float v = 0.f;
            if (gn < P.N && gk < Kseg) v = P.B[(int64_t)gk * P.ldb + gn];
            Bs[k][n] = v;
          }
        } else {
          // B stored [n][k]: k contiguous
          for (int i = tid; i < GT * GK; i += GEMM_THREADS) {
            int n = i >> 6, k = i & 63;
            int gn = n0 + n, gk = k0 + k;
            float v = 0.f;
            if (gn < P.N && gk < Kseg) v = P.B[(int64_t)gn * P.ldb + gk];
            Bs[k][n] = v;
          }
        }
      } else {
        for (int i = tid; i < GT * GK; i += GEMM_THREADS) {
          int m = i >> 6, k = i & 63;
          int gm = m0 + m, gk = k0 + k;
          As[k][m] = (gm < P.M && gk < Kseg) ? P.A2[(int64_t)gm * P.lda2 + gk] : 0.f;
        }
        for (int i = tid; i < GT * GK; i += GEMM_THREADS) {
          int k = i >> 5, n = i & 31;
          int gn = n0 + n, gk = k0 + k;
          Bs[k][n] = (gn < P.N && gk < Kseg) ? P.B2[(int64_t)gk * P.ldb2 + gn] : 0.f;
        }
      }
      __syncthreads();
      // ---- each k-group handles 16 of the 64 staged k
      const int kb = kg * 16;
#pragma unroll
      for (int kk = 0; kk < 16; ++kk) {
        const float4 a = *reinterpret_cast<const float4*>(&As[kb + kk][ty * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&Bs[kb + kk][tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w};
        const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  // ---- deterministic in-CTA split-K reduction through shared memory (reuse As/Bs)
  // k-groups 0,1 park their partial tiles in As, groups 2,3 in Bs (2 * 1024 floats fit in each)
  float* mine = (kg < 2) ? (&As[0][0] + kg * (GT * GT)) : (&Bs[0][0] + (kg - 2) * (GT * GT));
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) mine[(ty * 4 + i) * GT + tx * 4 + j] = acc[i][j];
  __syncthreads();
  const float* r0 = &As[0][0];
  const float* r1 = &As[0][0] + GT * GT;
  const float* r2 = &Bs[0][0];
  const float* r3 = &Bs[0][0] + GT * GT;
  for (int i = tid; i < GT * GT; i += GEMM_THREADS) {
    const int m = i >> 5, n = i & 31;
    const int gm = m0 + m, gn = n0 + n;
    if (gm >= P.M || gn >= P.N) continue;
    float v = (r0[i] + r1[i]) + (r2[i] + r3[i]);
    if (P.bias) v += P.bias[gn];
    switch (P.epi) {
      case EPI_RELU: v = fmaxf(v, 0.f); break;
      case EPI_RELU_MASK: v = (P.aux[(int64_t)gm * P.ldaux + gn] > 0.f) ? v : 0.f; break;
      case EPI_TANH: v = tanhf(v); break;
      case EPI_ACTOR_DY: {
        float th = P.aux[(int64_t)gm * P.ldaux + gn];
        v = (v + P.coef * th) * (1.f - th * th);
        break;
      }
      default: break;
    }
    P.C[(int64_t)gm * P.ldc + gn] = v;
    if (P.C2) P.C2[(int64_t)gm * P.ldc2 + gn] = v * P.scale2;
  }
}

inline int launch_gemm_batch(GemmBatch& G, cudaStream_t s) {
  int t = 0;
  for (int i = 0; i < G.n; ++i) {
    GemmProb& P = G.p[i];
    P.tiles_n = (P.N + GT - 1) / GT;
    P.tile_begin = t;
    t += ((P.M + GT - 1) / GT) * P.tiles_n;
  }
  G.total_tiles = t;
  if (t == 0) return CUR_OK;
  grouped_gemm_kernel<<<t, GEMM_THREADS, 0, s>>>(G);
  CUR_CHECK_LAUNCH();
  return CUR_OK;
}

}  // namespace cur
