// Grouped small-GEMM kernel for the DDPG actor/critic MLPs (sm_100a, fp32 CUDA cores).
//
// At the reference batch (256 rows, 256 hidden units) one layer is a 256x256x256 GEMM: 17 MFLOP.
// That is far below what fills tensor-core tiles on 148 SMs, and the parity target (rel 1e-5 vs an
// fp32 restatement of the TF graph) rules out single-pass TF32/BF16, so the layer GEMMs
// run on FFMA.  What matters is latency: every launch of this kernel executes ALL independent GEMMs
// of one dependency level of the DDPG graph (e.g. main.pi / target.pi / main.Q layer k) as one grid
// ("grouped GEMM"): 32x32 output tiles, K split 4 ways inside the CTA (8 warps, 4x4 micro-tiles) and
// reduced deterministically through shared memory.
//
// One problem:  C[M,N] = epi( opA(A)[M,K] * opB(B)[K,N]  (+ A2[M,K2] * B2[K2,N])  + bias )
//   a_trans: A is stored [K][M] (dW = X^T dY)        b_trans: B is stored [N][K] (dX = dY W^T)
//   ones_a : A is the all-ones row vector (M == 1): C = column sums of B = the bias gradient, written
//            right behind dW because [W;b] is one contiguous block of the flat GetFlat vector
//
// Operand tiles are staged global -> shared with 16-byte cp.async (LDGSTS, zero-fill for the edges)
// into a 2-stage ring, in the operand's NATIVE orientation (no transposing stores): the micro-tile
// row/column assignment is chosen per orientation so that every shared-memory read is a conflict-free
// LDS.128.  Problems whose leading dimensions / extents are not multiples of 4 (the N == 1 critic
// output layer) take the scalar generic path.
#pragma once
#include "common.cuh"

namespace cur {

enum { EPI_NONE = 0, EPI_RELU = 1, EPI_RELU_MASK = 2, EPI_TANH = 3, EPI_ACTOR_DY = 4 };
enum { VAR_GENERIC = 0, VAR_MK_KN = 1, VAR_MK_NK = 2, VAR_KM_KN = 3 };

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
  int a_trans, b_trans, ones_a;
  int split_k;                 // tensor-core path: split K over CTAs although A is not transposed (K = batch, K-major
                               // operands); 2: coarse splits (512 rows above batch 2048), see tc_pick_splits
  int accumulate;              // weight-gradient launch of the rows schedule: C += result (micro-batch j > 0)
  int epi;
  float scale2, coef;
  int variant;                 // filled by the launcher
  int tiles_n, tile_begin;     // filled by the launcher
};

// Optional Adam epilogue (fused dW + optimiser launch of the "rows" schedule): C is a block of the flat
// gradient arena; the parameter / moment at the same arena offset is stepped right after the gradient
// element is final.  All scalars are resolved by the calling kernel (see ddpg_rows.cu).
struct AdamCtx {
  const float* grads;          // arena base of C
  float *theta, *m, *v;        // arenas with the same layout
  float neg_a, b1, omb1, b2, omb2, eps;
};

constexpr int GEMM_MAX_PROBS = 24;
constexpr int GT = 32;          // tile edge
constexpr int GK = 64;          // K chunk per stage (4 k-groups x 16)
constexpr int GEMM_THREADS = 256;
constexpr int LD_MAJ = GK + 4;  // row stride of a [32][64] tile (k contiguous)
constexpr int LD_MIN = GT + 4;  // row stride of a [64][32] tile (m or n contiguous)
constexpr int TILE_FLOATS = (GT * LD_MAJ > GK * LD_MIN) ? GT * LD_MAJ : GK * LD_MIN;   // 2304

struct GemmBatch {
  GemmProb p[GEMM_MAX_PROBS];
  int n;
  int total_tiles;
};

__device__ __forceinline__ void cp16_zfill(float* dst_smem, const float* src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_wait0() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ float gemm_epilogue(const GemmProb& P, float v, int gm, int gn) {
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
  return v;
}

// Split-K reduction over the 4 k-groups + epilogue.  `red` holds 4 partial 32x32 tiles.
__device__ __forceinline__ void reduce_and_store(const GemmProb& P, const float* red, int m0, int n0, int tid,
                                                 const AdamCtx* ax = nullptr) {
  for (int i = tid; i < GT * GT; i += GEMM_THREADS) {
    const int m = i >> 5, n = i & 31;
    const int gm = m0 + m, gn = n0 + n;
    if (gm >= P.M || gn >= P.N) continue;
    float v = (red[i] + red[GT * GT + i]) + (red[2 * GT * GT + i] + red[3 * GT * GT + i]);
    v = gemm_epilogue(P, v, gm, gn);
    float* c = P.C + (int64_t)gm * P.ldc + gn;
    *c = v;
    if (P.C2) P.C2[(int64_t)gm * P.ldc2 + gn] = v * P.scale2;
    if (ax) {
      const int64_t off = c - ax->grads;
      float th = ax->theta[off], mm = ax->m[off], vv = ax->v[off];
      adam_elem(th, v, mm, vv, ax->neg_a, ax->b1, ax->omb1, ax->b2, ax->omb2, ax->eps);
      ax->theta[off] = th; ax->m[off] = mm; ax->v[off] = vv;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// fast path: cp.async staging in native orientation
//   A_KM == false: A tile [32 m][64 k] (k contiguous), micro-tile rows  ty + 8 i
//   A_KM == true : A tile [64 k][32 m] (m contiguous), micro-tile rows  4 ty + i
//   B_NK == false: B tile [64 k][32 n] (n contiguous), micro-tile cols  4 tx + j
//   B_NK == true : B tile [32 n][64 k] (k contiguous), micro-tile cols  tx + 8 j
// The optional second K segment (A2,B2) exists only for the <false,false> instantiation.
// ------------------------------------------------------------------------------------------------
template <bool A_KM, bool B_NK>
__device__ __forceinline__ void gemm_tile_fast(const GemmProb& P, float* As, float* Bs, int m0, int n0,
                                               const AdamCtx* ax = nullptr) {
  const int tid = threadIdx.x;
  const int kg = tid >> 6;
  const int lt = tid & 63;
  const int ty = lt >> 3, tx = lt & 7;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int nc0 = (P.K + GK - 1) / GK;
  const int nc1 = (!A_KM && !B_NK) ? (P.K2 + GK - 1) / GK : 0;
  const int nchunks = nc0 + nc1;
  const bool ones_a = A_KM && (P.ones_a != 0);

  if (ones_a) {
    // A = ones row vector (M == 1): [k][m] tile whose column 0 is 1, written once, never restaged
    for (int i = tid; i < 2 * TILE_FLOATS; i += GEMM_THREADS) As[i] = 0.f;
    __syncthreads();
    for (int i = tid; i < 2 * GK; i += GEMM_THREADS) As[(i >> 6) * TILE_FLOATS + (i & 63) * LD_MIN] = 1.f;
  }

  // stage chunk c into ring slot s
  auto stage = [&](int c, int s) {
    float* as = As + s * TILE_FLOATS;
    float* bs = Bs + s * TILE_FLOATS;
    if (!A_KM && !B_NK && c >= nc0) {   // second K segment: plain A2 [m][k], B2 [k][n]
      const int k0 = (c - nc0) * GK;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int f = tid + j * GEMM_THREADS;
        {
          const int m = f >> 4, k = (f & 15) << 2;
          const bool ok = (m0 + m < P.M) && (k0 + k < P.K2);
          cp16_zfill(as + m * LD_MAJ + k, P.A2 + (ok ? (int64_t)(m0 + m) * P.lda2 + (k0 + k) : 0), ok);
        }
        {
          const int k = f >> 3, n = (f & 7) << 2;
          const bool ok = (k0 + k < P.K2) && (n0 + n < P.N);
          cp16_zfill(bs + k * LD_MIN + n, P.B2 + (ok ? (int64_t)(k0 + k) * P.ldb2 + (n0 + n) : 0), ok);
        }
      }
      return;
    }
    const int k0 = c * GK;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int f = tid + j * GEMM_THREADS;
      if (!ones_a) {
        if (!A_KM) {
          const int m = f >> 4, k = (f & 15) << 2;
          const bool ok = (m0 + m < P.M) && (k0 + k < P.K);
          cp16_zfill(as + m * LD_MAJ + k, P.A + (ok ? (int64_t)(m0 + m) * P.lda + (k0 + k) : 0), ok);
        } else {
          const int k = f >> 3, m = (f & 7) << 2;
          const bool ok = (k0 + k < P.K) && (m0 + m < P.M);
          cp16_zfill(as + k * LD_MIN + m, P.A + (ok ? (int64_t)(k0 + k) * P.lda + (m0 + m) : 0), ok);
        }
      }
      if (!B_NK) {
        const int k = f >> 3, n = (f & 7) << 2;
        const bool ok = (k0 + k < P.K) && (n0 + n < P.N);
        cp16_zfill(bs + k * LD_MIN + n, P.B + (ok ? (int64_t)(k0 + k) * P.ldb + (n0 + n) : 0), ok);
      } else {
        const int n = f >> 4, k = (f & 15) << 2;
        const bool ok = (n0 + n < P.N) && (k0 + k < P.K);
        cp16_zfill(bs + n * LD_MAJ + k, P.B + (ok ? (int64_t)(n0 + n) * P.ldb + (k0 + k) : 0), ok);
      }
    }
  };

  if (nchunks > 0) stage(0, 0);
  cp_commit();
  for (int c = 0; c < nchunks; ++c) {
    const int s = c & 1;
    cp_wait0();
    __syncthreads();                       // chunk c landed; everybody is done computing chunk c-1
    if (c + 1 < nchunks) stage(c + 1, s ^ 1);
    cp_commit();
    const float* as = As + s * TILE_FLOATS;
    const float* bs = Bs + s * TILE_FLOATS;
    const int kb = kg * 16;
#pragma unroll
    for (int k4 = 0; k4 < 16; k4 += 4) {
      float a[4][4], b[4][4];   // [kk][i], [kk][j]
      if (!A_KM) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 v = *reinterpret_cast<const float4*>(as + (ty + 8 * i) * LD_MAJ + kb + k4);
          a[0][i] = v.x; a[1][i] = v.y; a[2][i] = v.z; a[3][i] = v.w;
        }
      } else {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const float4 v = *reinterpret_cast<const float4*>(as + (kb + k4 + kk) * LD_MIN + 4 * ty);
          a[kk][0] = v.x; a[kk][1] = v.y; a[kk][2] = v.z; a[kk][3] = v.w;
        }
      }
      if (!B_NK) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const float4 v = *reinterpret_cast<const float4*>(bs + (kb + k4 + kk) * LD_MIN + 4 * tx);
          b[kk][0] = v.x; b[kk][1] = v.y; b[kk][2] = v.z; b[kk][3] = v.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 v = *reinterpret_cast<const float4*>(bs + (tx + 8 * j) * LD_MAJ + kb + k4);
          b[0][j] = v.x; b[1][j] = v.y; b[2][j] = v.z; b[3][j] = v.w;
        }
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[kk][i], b[kk][j], acc[i][j]);
    }
  }
  __syncthreads();   // all reads of the ring are done before it is reused for the reduction

  float* red = As;   // 4 * 1024 floats <= 2 * TILE_FLOATS
  float* mine = red + kg * (GT * GT);
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = A_KM ? 4 * ty + i : ty + 8 * i;
      const int cc = B_NK ? tx + 8 * j : 4 * tx + j;
      mine[r * GT + cc] = acc[i][j];
    }
  __syncthreads();
  reduce_and_store(P, red, m0, n0, tid, ax);
}

// ------------------------------------------------------------------------------------------------
// generic path: scalar loads with full bounds checks, register double buffering, transposing stores
// ------------------------------------------------------------------------------------------------
struct ChunkRegs {
  float a[8], b[8];
};

__device__ __forceinline__ void chunk_gload(const GemmProb& P, int seg, int k0, int m0, int n0, int tid,
                                            ChunkRegs& R) {
  if (seg == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int i = tid + j * GEMM_THREADS;
      float v = 0.f;
      if (P.ones_a) {
        const int gk = k0 + (i >> 5), gm = m0 + (i & 31);
        if (gk < P.K && gm < P.M) v = 1.f;
      } else if (!P.a_trans) {
        const int gm = m0 + (i >> 6), gk = k0 + (i & 63);
        if (gm < P.M && gk < P.K) v = P.A[(int64_t)gm * P.lda + gk];
      } else {
        const int gk = k0 + (i >> 5), gm = m0 + (i & 31);
        if (gk < P.K && gm < P.M) v = P.A[(int64_t)gk * P.lda + gm];
      }
      R.a[j] = v;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int i = tid + j * GEMM_THREADS;
      float v = 0.f;
      if (!P.b_trans) {
        const int gk = k0 + (i >> 5), gn = n0 + (i & 31);
        if (gn < P.N && gk < P.K) v = P.B[(int64_t)gk * P.ldb + gn];
      } else {
        const int gn = n0 + (i >> 6), gk = k0 + (i & 63);
        if (gn < P.N && gk < P.K) v = P.B[(int64_t)gn * P.ldb + gk];
      }
      R.b[j] = v;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int i = tid + j * GEMM_THREADS;
      const int gm = m0 + (i >> 6), gk = k0 + (i & 63);
      R.a[j] = (gm < P.M && gk < P.K2) ? P.A2[(int64_t)gm * P.lda2 + gk] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int i = tid + j * GEMM_THREADS;
      const int gk = k0 + (i >> 5), gn = n0 + (i & 31);
      R.b[j] = (gn < P.N && gk < P.K2) ? P.B2[(int64_t)gk * P.ldb2 + gn] : 0.f;
    }
  }
}

__device__ __forceinline__ void chunk_sstore(const GemmProb& P, int seg, int tid, const ChunkRegs& R, float* As,
                                             float* Bs) {
  const bool at = (seg == 0) && (P.a_trans || P.ones_a);
  const bool bt = (seg == 0) && P.b_trans;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int i = tid + j * GEMM_THREADS;
    if (!at) As[(i & 63) * LD_MIN + (i >> 6)] = R.a[j];
    else As[(i >> 5) * LD_MIN + (i & 31)] = R.a[j];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int i = tid + j * GEMM_THREADS;
    if (!bt) Bs[(i >> 5) * LD_MIN + (i & 31)] = R.b[j];
    else Bs[(i & 63) * LD_MIN + (i >> 6)] = R.b[j];
  }
}

static __device__ __noinline__ void gemm_tile_generic(const GemmProb& P, float* As, float* Bs, int m0, int n0,
                                               const AdamCtx* ax = nullptr) {
  const int tid = threadIdx.x;
  const int kg = tid >> 6;
  const int lt = tid & 63;
  const int ty = lt >> 3, tx = lt & 7;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int nc0 = (P.K + GK - 1) / GK;
  const int nc1 = (P.K2 + GK - 1) / GK;
  const int nchunks = nc0 + nc1;
  ChunkRegs R;
  if (nchunks > 0) {
    chunk_gload(P, nc0 > 0 ? 0 : 1, 0, m0, n0, tid, R);
    chunk_sstore(P, nc0 > 0 ? 0 : 1, tid, R, As, Bs);
  }
  __syncthreads();
  for (int c = 0; c < nchunks; ++c) {
    const int cur = c & 1;
    const bool more = (c + 1 < nchunks);
    const int nseg = (c + 1 < nc0) ? 0 : 1;
    if (more) chunk_gload(P, nseg, (nseg == 0 ? (c + 1) : (c + 1 - nc0)) * GK, m0, n0, tid, R);
    const float* as = As + cur * TILE_FLOATS;
    const float* bs = Bs + cur * TILE_FLOATS;
    const int kb = kg * 16;
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(as + (kb + kk) * LD_MIN + ty * 4);
      const float4 b = *reinterpret_cast<const float4*>(bs + (kb + kk) * LD_MIN + tx * 4);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) chunk_sstore(P, nseg, tid, R, As + (cur ^ 1) * TILE_FLOATS, Bs + (cur ^ 1) * TILE_FLOATS);
    __syncthreads();
  }
  float* red = As;
  float* mine = red + kg * (GT * GT);
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) mine[(ty * 4 + i) * GT + tx * 4 + j] = acc[i][j];
  __syncthreads();
  reduce_and_store(P, red, m0, n0, tid, ax);
}

// problem / tile lookup shared by the kernels built on the tile routines; the descriptor is copied to
// shared memory once (dynamic indexing of the kernel-parameter array would otherwise turn every field
// access into a constant-bank load)
__device__ __forceinline__ void gemm_run_tile(const GemmBatch& G, GemmProb& Ps, float* As, float* Bs,
                                              const AdamCtx* ax) {
  int pi = 0;
#pragma unroll 1
  while (pi + 1 < G.n && (int)blockIdx.x >= G.p[pi + 1].tile_begin) ++pi;
  {
    const int* src = reinterpret_cast<const int*>(&G.p[pi]);
    int* dst = reinterpret_cast<int*>(&Ps);
    for (int i = threadIdx.x; i < (int)(sizeof(GemmProb) / 4); i += GEMM_THREADS) dst[i] = src[i];
  }
  __syncthreads();
  const GemmProb& P = Ps;
  const int tile = blockIdx.x - P.tile_begin;
  const int tm = tile / P.tiles_n, tn = tile - tm * P.tiles_n;
  const int m0 = tm * GT, n0 = tn * GT;
  switch (P.variant) {
    case VAR_MK_KN: gemm_tile_fast<false, false>(P, As, Bs, m0, n0, ax); break;
    case VAR_MK_NK: gemm_tile_fast<false, true>(P, As, Bs, m0, n0, ax); break;
    case VAR_KM_KN: gemm_tile_fast<true, false>(P, As, Bs, m0, n0, ax); break;
    default: gemm_tile_generic(P, As, Bs, m0, n0, ax); break;
  }
}

static __global__ void __launch_bounds__(GEMM_THREADS, 2) grouped_gemm_kernel(const __grid_constant__ GemmBatch G) {
  __shared__ __align__(16) float As[2 * TILE_FLOATS];
  __shared__ __align__(16) float Bs[2 * TILE_FLOATS];
  __shared__ GemmProb Ps;
  gemm_run_tile(G, Ps, As, Bs, nullptr);
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

inline int pick_variant(const GemmProb& P) {
  // fast variants need 16-byte aligned, multiple-of-4 contiguous extents for every staged operand
  if (P.ones_a) {
    if (!P.b_trans && al16(P.B) && P.ldb % 4 == 0 && P.N % 4 == 0 && P.K2 == 0) return VAR_KM_KN;
    return VAR_GENERIC;
  }
  if (!P.a_trans && !P.b_trans) {
    const bool seg2_ok = (P.K2 == 0) || (al16(P.A2) && al16(P.B2) && P.lda2 % 4 == 0 && P.ldb2 % 4 == 0 &&
                                         P.K2 % 4 == 0);
    if (al16(P.A) && al16(P.B) && P.lda % 4 == 0 && P.ldb % 4 == 0 && P.K % 4 == 0 && P.N % 4 == 0 && seg2_ok)
      return VAR_MK_KN;
  } else if (!P.a_trans && P.b_trans) {
    if (al16(P.A) && al16(P.B) && P.lda % 4 == 0 && P.ldb % 4 == 0 && P.K % 4 == 0 && P.K2 == 0) return VAR_MK_NK;
  } else if (P.a_trans && !P.b_trans) {
    if (al16(P.A) && al16(P.B) && P.lda % 4 == 0 && P.ldb % 4 == 0 && P.M % 4 == 0 && P.N % 4 == 0 && P.K2 == 0)
      return VAR_KM_KN;
  }
  return VAR_GENERIC;
}

inline int plan_gemm_batch(GemmBatch& G) {
  int t = 0;
  for (int i = 0; i < G.n; ++i) {
    GemmProb& P = G.p[i];
    P.variant = pick_variant(P);
    P.tiles_n = (P.N + GT - 1) / GT;
    P.tile_begin = t;
    t += ((P.M + GT - 1) / GT) * P.tiles_n;
  }
  G.total_tiles = t;
  return t;
}

inline int launch_gemm_batch(GemmBatch& G, cudaStream_t s) {
  const int t = plan_gemm_batch(G);
  if (t == 0) return CUR_OK;
  grouped_gemm_kernel<<<t, GEMM_THREADS, 0, s>>>(G);
  CUR_CHECK_LAUNCH();
  return CUR_OK;
}

}  // namespace cur
