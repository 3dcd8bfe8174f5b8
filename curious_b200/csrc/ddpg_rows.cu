// DDPG / UVFA update, "rows" schedule (sm_100a): the whole actor-critic graph of one update in TWO launches.
//
// Replaces (reference flowersteam/curious), like ddpg.cu but with a latency-oriented schedule:
//   baselines/her/actor_critic.py:5-98, util.py:56-107   networks
//   baselines/her/ddpg.py:412-449                        losses + tf.gradients + flatten_grads
//   baselines/common/mpi_adam.py:30-35                   Adam (fused into the weight-gradient launch)
//
// Why: at the reference batch (256 rows, 3x256 MLPs) one update is 0.73 GFLOP - ~10 us of FFMA - but the
// dependency-level schedule of ddpg.cu needs 17 dependent launches (140 us).  Rows of the batch are
// independent until the weight gradient, so:
//
//   launch 1  ddpg_rows_kernel   one thread-block CLUSTER of 8 CTAs per 16 batch rows.  CTA c owns hidden
//             columns [32c, 32c+32) of every layer of every net.  The cluster runs the whole chain for its
//             rows - input assembly (normalise / concat), 5 forward nets, loss seeds, the critic,
//             actor-through-critic and actor data-gradient chains.
//               * weights never touch shared memory: every thread loads the 16 k-rows x 2 columns it
//                 multiplies straight from L2 into registers (each element is fetched once per CTA), and
//                 the loads for the NEXT net/layer are issued as soon as the current FFMA loop ends, so
//                 their latency hides behind the reduction, the epilogue and the exchange;
//               * activations live in a transposed shared tile [k][16 rows]: a k-slice (half warp) reads
//                 its 16 rows with 4 broadcast LDS.128 per k for 32 FFMA;
//               * K is split 16 ways (2 k-slices per warp), combined by one shuffle + a shared-memory
//                 reduction in fixed order (deterministic);
//               * layer outputs are exchanged through DISTRIBUTED SHARED MEMORY: the thread that owns 4 rows
//                 of one output column stores the float4 straight into the activation tile of all 8 CTAs of
//                 the cluster (st.shared::cluster), bracketed by split cluster barriers (arrive after the
//                 last tile read / wait before the remote stores; arrive.release after them / wait.acquire
//                 before the next layer).  No L2 round trip and no global fence between layers; the
//                 row-major global copies that launch 2 needs are plain posted stores off the critical path.
//             <= 128 registers and ~112 KB of shared memory per CTA, so two CTAs fit on an SM and all 16
//             clusters of a 256-row batch are co-resident (the first version - 1 CTA/SM - could only place
//             15 clusters and ran in two waves, see profiles/).
//   launch 2  rows_dw_kernel     every dW = X^T dY and db = 1^T dY of both nets as one grouped GEMM with
//             the full batch as K (deterministic, no atomics), written into the flat GetFlat-ordered
//             gradient arena; optionally Adam is applied to the element in the same epilogue (world
//             size 1: no all-reduce between gradient and step).  The last CTA to finish folds the
//             per-cluster loss partials and bumps the device step counter.
//
// All arithmetic is FP32 FFMA (IEEE); see DESIGN.md section 4 for why tensor cores do not apply here.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "net_layout.cuh"

namespace cg = cooperative_groups;

namespace cur {

constexpr int R_ROWS = 16;               // batch rows per cluster
constexpr int R_CS = 8;                  // CTAs per cluster
constexpr int R_CW = 32;                 // hidden columns per CTA
constexpr int R_H = R_CS * R_CW;         // 256 hidden units
constexpr int R_THREADS = 256;
constexpr int R_NSLICE = 16;             // k-slices (half warps)
constexpr int R_TILE = R_H * R_ROWS;     // floats of one transposed activation tile [k][16]
constexpr int R_MAXA = 3;                // activation tiles (nets per step)
constexpr int R_LDP = R_ROWS + 4;        // column stride of a split-K partial [32 cols][16 rows (+4)]
constexpr int R_PART = R_CW * R_LDP;     // floats of one warp's partial
constexpr int R_RED = (R_THREADS / 32) * R_PART;          // floats of the split-K partials of one net
constexpr int R_MAXL = 4;                // hidden layers supported by this schedule
constexpr int R_MAXW = 7 * R_MAXL;       // net-layer weight descriptors
constexpr int R_MAXS = 4 * R_MAXL;       // steps
constexpr int R_DU = 8;                  // max action dim
constexpr int R_MISC = 1024;
constexpr size_t R_SMEM_FLOATS = (size_t)R_MAXA * R_TILE + R_MAXA * R_RED + R_MISC;

struct WDesc {
  const float* w;     // fwd: rows = k ([K][H]);  bwd: rows = output columns ([H][H])
  const float* w2;    // first layer: rows >= split come from here (the goal block W0g)
  int split;          // rows < split from w, the rest from w2
  int kvalid;         // rows >= kvalid are zero
  int kper;           // k-rows per slice (K padded / 16)
  int kind;           // 0: forward hidden layer (fast path), 1: backward, 2: forward first layer (generic)
};

enum { POST_NONE = 0, POST_FOUT = 1, POST_GOUT = 2, POST_BIN = 3, POST_GOUT_BIN = 4 };

struct RStep {
  const float* aux[R_MAXA];   // forward: bias [H];  backward: row-major activation whose sign masks the gradient
  float* out_rm[R_MAXA];      // row-major [n][H] copy of the output (NULL: only the cluster needs it)
  int nA, bwd, shared_w, kper, post;
  int _pad;
};

struct RowsParams {
  cur_net_desc d;
  int in_sp, in_sq, in_g, KP, L, nw, nsteps;
  int64_t n;
  const float *o, *g, *u, *td, *o_2, *g_2, *r;
  const float *o_mean, *o_std, *g_mean, *g_std;
  float gamma, clip_return, action_l2;
  int clip_pos;
  const float *WoutP, *boutP, *WoutPT, *boutPT, *WoutQ, *boutQ, *WoutQT, *boutQT;
  const float* W0Q_act;      // main Q first-layer rows of the action inputs: [dimu][H]
  float *Xp, *Xq;            // [n][KP] first-layer inputs of main.pi / main.Q(u)   (weight-gradient operands)
  const float *hq_last, *hqp_last, *hp_last;    // row-major last hidden activations (ReLU masks of the seeds)
  float *dc_last, *dp_last;  // row-major gradients at the last hidden layer
  float *dQ, *dy;            // [n], [n][lddy]
  int lddy;
  float* loss_part;          // [n / 16][4]
  float* q_pi;               // [n]
  long long* tl;            // optional debug timeline (clock64 stamps of CTA 0), CUR_ROWS_TIMELINE=1
  WDesc wd[R_MAXW];
  RStep steps[R_MAXS];
};

#define R_TL(i)                                                                   \
  do {                                                                            \
    if (P.tl != nullptr && threadIdx.x == 0 && blockIdx.x == 0) P.tl[i] = clock64(); \
  } while (0)

struct Lane {
  int cp, s;      // column pair (0..15) and k-slice (0..15) of this thread (GEMM role)
  int col, rq;    // column (0..31) and row quad (0..3) this thread finishes (epilogue role), group = tid >> 7
};

// Load this thread's weights of one net-layer into registers: wr[kk] = W[k = s*kper + kk][cols 2cp, 2cp+1]
// (forward) or the transposed equivalent (backward: rows 32*rank + 2cp + {0,1}, k contiguous).
__device__ __forceinline__ void load_w(const RowsParams& P, int idx, int rank, const Lane& ln, float2 (&wr)[16]) {
  if (idx >= P.nw) return;
  const float* w = P.wd[idx].w;
  const int kind = P.wd[idx].kind;
  if (kind == 0) {
    const float2* p = reinterpret_cast<const float2*>(w + (int64_t)(ln.s * 16) * R_H + rank * R_CW + 2 * ln.cp);
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) wr[kk] = __ldg(p + kk * (R_H / 2));
  } else if (kind == 1) {
    const float4* r0 = reinterpret_cast<const float4*>(w + (int64_t)(rank * R_CW + 2 * ln.cp) * R_H + ln.s * 16);
    const float4* r1 = r0 + R_H / 4;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 a = __ldg(r0 + q);
      const float4 b = __ldg(r1 + q);
      wr[4 * q + 0] = make_float2(a.x, b.x);
      wr[4 * q + 1] = make_float2(a.y, b.y);
      wr[4 * q + 2] = make_float2(a.z, b.z);
      wr[4 * q + 3] = make_float2(a.w, b.w);
    }
  } else {
    const float* w2 = P.wd[idx].w2;
    const int split = P.wd[idx].split, kvalid = P.wd[idx].kvalid, kper = P.wd[idx].kper;
    const int col = rank * R_CW + 2 * ln.cp;
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float2 v = make_float2(0.f, 0.f);
      const int gk = ln.s * kper + kk;
      if (kk < kper && gk < kvalid) {
        const float* src = (gk < split) ? w + (int64_t)gk * R_H : w2 + (int64_t)(gk - split) * R_H;
        v = __ldg(reinterpret_cast<const float2*>(src + col));
      }
      wr[kk] = v;
    }
  }
}

// acc[r][c] = sum over this thread's k-slice of tile[k][r] * w[k][c];  then the two k-slices of the warp are
// combined and lanes 0..15 park the warp's partial, transposed [32 cols][16 rows], in `red`.
__device__ __forceinline__ void slice_gemm(const float* __restrict__ tile, const float2 (&wr)[16], int kper,
                                           const Lane& ln, float* __restrict__ red) {
  float acc[R_ROWS][2];
#pragma unroll
  for (int r = 0; r < R_ROWS; ++r) acc[r][0] = acc[r][1] = 0.f;
  const float* base = tile + ln.s * kper * R_ROWS;
#pragma unroll
  for (int kk = 0; kk < 16; ++kk) {
    if (kk < kper) {
      float a[R_ROWS];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(base + kk * R_ROWS + 4 * q);
        a[4 * q] = v.x; a[4 * q + 1] = v.y; a[4 * q + 2] = v.z; a[4 * q + 3] = v.w;
      }
#pragma unroll
      for (int r = 0; r < R_ROWS; ++r) {
        acc[r][0] = fmaf(a[r], wr[kk].x, acc[r][0]);
        acc[r][1] = fmaf(a[r], wr[kk].y, acc[r][1]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < R_ROWS; ++r) {
    acc[r][0] += __shfl_xor_sync(0xffffffffu, acc[r][0], 16);
    acc[r][1] += __shfl_xor_sync(0xffffffffu, acc[r][1], 16);
  }
  if ((threadIdx.x & 31) < 16) {
    float* mine = red + (threadIdx.x >> 5) * R_PART + (2 * ln.cp) * R_LDP;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      *reinterpret_cast<float4*>(mine + 4 * q) =
          make_float4(acc[4 * q][0], acc[4 * q + 1][0], acc[4 * q + 2][0], acc[4 * q + 3][0]);
      *reinterpret_cast<float4*>(mine + R_LDP + 4 * q) =
          make_float4(acc[4 * q][1], acc[4 * q + 1][1], acc[4 * q + 2][1], acc[4 * q + 3][1]);
    }
  }
}

// rows 4rq..4rq+3 of column `col` of one net: the 8 warp partials summed in fixed order
__device__ __forceinline__ float4 reduce_partials(const float* __restrict__ red, const Lane& ln) {
  const float* p = red + ln.col * R_LDP + 4 * ln.rq;
  float4 v[8];
#pragma unroll
  for (int w = 0; w < 8; ++w) v[w] = *reinterpret_cast<const float4*>(p + w * R_PART);
  float4 o;
  o.x = ((v[0].x + v[1].x) + (v[2].x + v[3].x)) + ((v[4].x + v[5].x) + (v[6].x + v[7].x));
  o.y = ((v[0].y + v[1].y) + (v[2].y + v[3].y)) + ((v[4].y + v[5].y) + (v[6].y + v[7].y));
  o.z = ((v[0].z + v[1].z) + (v[2].z + v[3].z)) + ((v[4].z + v[5].z) + (v[6].z + v[7].z));
  o.w = ((v[0].w + v[1].w) + (v[2].w + v[3].w)) + ((v[4].w + v[5].w) + (v[6].w + v[7].w));
  return o;
}

// store rows 4rq..4rq+3 of global column gcol into activation tile `tile_off` of every CTA of the cluster
__device__ __forceinline__ void publish(cg::cluster_group& cluster, float* smem_base, int tile_off, int gcol, int rq,
                                        float4 v) {
  float* local = smem_base + tile_off + gcol * R_ROWS + 4 * rq;
#pragma unroll
  for (int p = 0; p < R_CS; ++p) *reinterpret_cast<float4*>(cluster.map_shared_rank(local, p)) = v;
}

__device__ __forceinline__ void store_rm(float* out_rm, int64_t row0, int gcol, int rq, float4 v) {
  float* o = out_rm + (row0 + 4 * rq) * R_H + gcol;
  o[0] = v.x; o[R_H] = v.y; o[2 * R_H] = v.z; o[3 * R_H] = v.w;
}

__device__ __forceinline__ float4 load_mask(const float* rm, int64_t row0, int gcol, int rq) {
  const float* m = rm + (row0 + 4 * rq) * R_H + gcol;
  return make_float4(__ldcg(m), __ldcg(m + R_H), __ldcg(m + 2 * R_H), __ldcg(m + 3 * R_H));
}

__device__ __forceinline__ float4 mask4(float4 v, float4 m) {
  return make_float4(m.x > 0.f ? v.x : 0.f, m.y > 0.f ? v.y : 0.f, m.z > 0.f ? v.z : 0.f, m.w > 0.f ? v.w : 0.f);
}

__device__ __forceinline__ float norm1r(float x, const float* mean, const float* std, int k, float clip) {
  float v = __fdiv_rn(__fsub_rn(x, mean[k]), std[k]);          // normalizer.py:72-77
  return fminf(fmaxf(v, -clip), clip);
}

// Element (row r, column k) of a first-layer input [o | task_descr | action | g] (modular) or [o | g | action]
// (flat), zero beyond the fan-in (actor_critic.py:76-91).  act_kind: 0 none (pi net), 1 u / max_u, 2 `ths`.
__device__ __forceinline__ float x_elem(const RowsParams& P, int64_t row, int r, int k, bool target, int act_kind,
                                        const float* ths) {
  const cur_net_desc& d = P.d;
  const bool nrm = d.normalize_obs != 0;
  const float* o = target ? P.o_2 : P.o;
  const float* g = target ? P.g_2 : P.g;
  const int in_s = P.in_sp + (act_kind ? d.dimu : 0);
  float v = 0.f;
  int gj = -1, aj = -1;
  if (k < d.dimo) {
    v = o[row * d.dimo + k];
    if (nrm) v = norm1r(v, P.o_mean, P.o_std, k, d.norm_clip);
  } else if (d.modular) {
    if (k < d.dimo + d.dimtd) v = P.td[row * d.dimtd + (k - d.dimo)];     // never normalised
    else if (k < in_s) aj = k - d.dimo - d.dimtd;
    else if (k < in_s + d.dimg) gj = k - in_s;
  } else {
    if (k < d.dimo + d.dimg) gj = k - d.dimo;
    else if (k < in_s) aj = k - d.dimo - d.dimg;
  }
  if (gj >= 0) {
    v = g[row * d.dimg + gj];
    if (nrm) v = norm1r(v, P.g_mean, P.g_std, gj, d.norm_clip);
  }
  if (aj >= 0) v = (act_kind == 1) ? __fdiv_rn(P.u[row * d.dimu + aj], d.max_u) : ths[r * R_DU + aj];
  return v;
}

// transposed first-layer tile [KP][16]; optionally also the row-major [n][KP] copy for the weight gradient
__device__ __noinline__ void build_x(const RowsParams& P, float* tile, int64_t row0, bool target, int act_kind,
                                     const float* ths, float* gout) {
  for (int idx = threadIdx.x; idx < R_ROWS * P.KP; idx += R_THREADS) {
    const int k = idx >> 4, r = idx & 15;
    tile[idx] = x_elem(P, row0 + r, r, k, target, act_kind, ths);
  }
  if (gout) {
    for (int idx = threadIdx.x; idx < R_ROWS * P.KP; idx += R_THREADS) {
      const int r = idx / P.KP, k = idx - r * P.KP;
      gout[(row0 + r) * P.KP + k] = x_elem(P, row0 + r, r, k, target, act_kind, ths);
    }
  }
}

// partial[part][r][j] = sum over k in [16 part, 16 part + 16) of tile[k][r] * W[k * ldk + j * ldj], j < NOUT.
// 16 rows x 16 k-parts = 256 threads; the 16 lanes of a half warp read the same weight (broadcast).
template <int NOUT>
__device__ __forceinline__ void rowdot_partial(const float* __restrict__ tile, const float* __restrict__ W, int ldk,
                                               int ldj, float* __restrict__ scratch) {
  const int r = threadIdx.x & 15, part = threadIdx.x >> 4;
  float acc[NOUT];
#pragma unroll
  for (int j = 0; j < NOUT; ++j) acc[j] = 0.f;
#pragma unroll
  for (int kk = 0; kk < 16; ++kk) {
    const int k = part * 16 + kk;
    const float x = tile[k * R_ROWS + r];
#pragma unroll
    for (int j = 0; j < NOUT; ++j) acc[j] = fmaf(x, __ldg(W + (int64_t)k * ldk + (int64_t)j * ldj), acc[j]);
  }
#pragma unroll
  for (int j = 0; j < NOUT; ++j) scratch[(part * R_ROWS + r) * R_DU + j] = acc[j];
}

__device__ __noinline__ void rowdot(const float* tile, const float* W, int ldk, int ldj, int nout, float* scratch) {
  if (nout == 1) rowdot_partial<1>(tile, W, ldk, ldj, scratch);
  else if (nout == 4) rowdot_partial<4>(tile, W, ldk, ldj, scratch);
  else
    for (int j = 0; j < nout; ++j) rowdot_partial<1>(tile, W + (int64_t)j * ldj, ldk, ldj, scratch + j);
}

// out[r][j] = sum over the 16 k-parts (fixed order) + bias[j];  call with all threads after a __syncthreads
__device__ __forceinline__ void rowdot_finish(const float* scratch, int nout, const float* bias, float* out) {
  if (threadIdx.x < R_ROWS * R_DU) {
    const int r = threadIdx.x >> 3, j = threadIdx.x & (R_DU - 1);
    if (j < nout) {
      float v = 0.f;
#pragma unroll
      for (int p = 0; p < 16; ++p) v += scratch[(p * R_ROWS + r) * R_DU + j];
      out[threadIdx.x] = v + (bias ? bias[j] : 0.f);
    }
  }
}

__global__ void __cluster_dims__(R_CS, 1, 1) __launch_bounds__(R_THREADS, 2)
ddpg_rows_kernel(const __grid_constant__ RowsParams P) {
  extern __shared__ __align__(16) float smem[];
  float* tiles = smem;
  float* red = tiles + R_MAXA * R_TILE;
  float* misc = red + R_MAXA * R_RED;
  float* s_th = misc;                 // [16][8] tanh output of main.pi (= pi / max_u)
  float* s_tht = misc + 128;          // target.pi
  float* s_q = misc + 256;            // [16][8] col 0: main.Q(o,g,u)
  float* s_qpi = misc + 384;          // main.Q(o,g,pi)
  float* s_qt = misc + 512;           // target.Q
  float* s_dq = misc + 640;           // [16] dQ, [16..32) dQpi, [32..48) squared TD error
  float* s_dy = misc + 768;           // [16][8]

  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int64_t row0 = (int64_t)(blockIdx.x / R_CS) * R_ROWS;
  const int tid = threadIdx.x;
  const cur_net_desc& d = P.d;
  const int L = P.L;
  const float inv_n = 1.0f / (float)P.n;
  Lane ln;
  ln.cp = tid & 15;
  ln.s = 2 * (tid >> 5) + ((tid >> 4) & 1);
  ln.col = (tid & 127) >> 2;
  ln.rq = tid & 3;
  const int group = tid >> 7;                   // which nets of a step this thread finishes: a with (a & 1) == group
  const int gcol = rank * R_CW + ln.col;

  float2 wr[16];
  int wi = 0;
  R_TL(0);
  load_w(P, wi, rank, ln, wr);

  // first-layer inputs of main.pi | target.pi | main.Q(o,g,u)
  build_x(P, tiles + 0 * R_TILE, row0, false, 0, nullptr, rank == 0 ? P.Xp : nullptr);
  build_x(P, tiles + 1 * R_TILE, row0, true, 0, nullptr, nullptr);
  build_x(P, tiles + 2 * R_TILE, row0, false, 1, nullptr, rank == 0 ? P.Xq : nullptr);
  R_TL(1);
  cluster.sync();        // every CTA of the cluster is running before the first remote store (and tiles are built)
  R_TL(2);

#pragma unroll 1
  for (int st = 0; st < P.nsteps; ++st) {
    const int nA = P.steps[st].nA, bwd = P.steps[st].bwd, shared_w = P.steps[st].shared_w;
    const int kper = P.steps[st].kper, post = P.steps[st].post;
    // ---- prefetch what the epilogue of this step needs (bias / ReLU mask), for the nets this thread finishes
    const int a0 = group, a1 = group + 2;       // group 0: nets 0 and 2;  group 1: net 1
    float4 aux0 = make_float4(0.f, 0.f, 0.f, 0.f), aux1 = aux0;
    if (a0 < nA) {
      const float* ax = P.steps[st].aux[a0];
      if (!bwd) aux0.x = __ldg(ax + gcol); else aux0 = load_mask(ax, row0, gcol, ln.rq);
    }
    if (a1 < nA) {
      const float* ax = P.steps[st].aux[a1];
      if (!bwd) aux1.x = __ldg(ax + gcol); else aux1 = load_mask(ax, row0, gcol, ln.rq);
    }
    // ---- the layer GEMMs of this step
#pragma unroll 1
    for (int a = 0; a < nA; ++a) {
      slice_gemm(tiles + a * R_TILE, wr, kper, ln, red + a * R_RED);
      if (!shared_w || a == nA - 1) load_w(P, ++wi, rank, ln, wr);
    }
    R_TL(8 + 8 * st + 0);
    __syncthreads();
    R_TL(8 + 8 * st + 1);
    cluster.barrier_arrive();                   // E1: this thread will not read the tiles of this step again
    float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
    if (a0 < nA) {
      v0 = reduce_partials(red + a0 * R_RED, ln);
      if (!bwd) {
        v0.x = fmaxf(v0.x + aux0.x, 0.f); v0.y = fmaxf(v0.y + aux0.x, 0.f);
        v0.z = fmaxf(v0.z + aux0.x, 0.f); v0.w = fmaxf(v0.w + aux0.x, 0.f);
      } else v0 = mask4(v0, aux0);
      if (P.steps[st].out_rm[a0]) store_rm(P.steps[st].out_rm[a0], row0, gcol, ln.rq, v0);
    }
    if (a1 < nA) {
      v1 = reduce_partials(red + a1 * R_RED, ln);
      if (!bwd) {
        v1.x = fmaxf(v1.x + aux1.x, 0.f); v1.y = fmaxf(v1.y + aux1.x, 0.f);
        v1.z = fmaxf(v1.z + aux1.x, 0.f); v1.w = fmaxf(v1.w + aux1.x, 0.f);
      } else v1 = mask4(v1, aux1);
      if (P.steps[st].out_rm[a1]) store_rm(P.steps[st].out_rm[a1], row0, gcol, ln.rq, v1);
    }
    R_TL(8 + 8 * st + 2);
    cluster.barrier_wait();                     // E1: every CTA of the cluster is done reading its tiles
    R_TL(8 + 8 * st + 3);
    if (a0 < nA) publish(cluster, smem, a0 * R_TILE, gcol, ln.rq, v0);
    if (a1 < nA) publish(cluster, smem, a1 * R_TILE, gcol, ln.rq, v1);
    R_TL(8 + 8 * st + 4);
    cluster.barrier_arrive();                   // E2 (release): remote stores issued
    cluster.barrier_wait();                     // E2 (acquire): all slices of the next layer's input have landed
    R_TL(8 + 8 * st + 5);

    if (post == POST_FOUT) {
      // output layers of main.pi | target.pi | main.Q(u), redundantly in every CTA, then the first-layer inputs
      // of main.Q(o,g,pi) | target.Q(o2,g2,pi_t)
      rowdot(tiles + 0 * R_TILE, P.WoutP, d.dimu, 1, d.dimu, red);
      rowdot(tiles + 1 * R_TILE, P.WoutPT, d.dimu, 1, d.dimu, red + 2048);
      rowdot(tiles + 2 * R_TILE, P.WoutQ, 1, 1, 1, red + 4096);
      __syncthreads();
      rowdot_finish(red, d.dimu, P.boutP, s_th);
      rowdot_finish(red + 2048, d.dimu, P.boutPT, s_tht);
      rowdot_finish(red + 4096, 1, P.boutQ, s_q);
      __syncthreads();
      if (tid < R_ROWS * R_DU && (tid & (R_DU - 1)) < d.dimu) {
        s_th[tid] = tanhf(s_th[tid]);          // actor_critic.py:89: pi = max_u * tanh(.)
        s_tht[tid] = tanhf(s_tht[tid]);
      }
      __syncthreads();
      build_x(P, tiles + 0 * R_TILE, row0, false, 2, s_th, nullptr);
      build_x(P, tiles + 1 * R_TILE, row0, true, 2, s_tht, nullptr);   // same u-slot and td for the target (ddpg.py:427-431)
      __syncthreads();
    }
    if (post == POST_GOUT || post == POST_GOUT_BIN) {
      rowdot(tiles + 0 * R_TILE, P.WoutQ, 1, 1, 1, red);
      rowdot(tiles + 1 * R_TILE, P.WoutQT, 1, 1, 1, red + 2048);
      __syncthreads();
      rowdot_finish(red, 1, P.boutQ, s_qpi);
      rowdot_finish(red + 2048, 1, P.boutQT, s_qt);
      __syncthreads();
      // losses (ddpg.py:436-441) and backward seeds
      if (tid < R_ROWS) {
        const int64_t row = row0 + tid;
        const float hi = P.clip_pos ? 0.f : INFINITY;
        const float tgt = fminf(fmaxf(P.r[row] + P.gamma * s_qt[tid * R_DU], -P.clip_return), hi);
        const float diff = tgt - s_q[tid * R_DU];
        s_dq[tid] = -2.0f * inv_n * diff;          // d mean((tgt - Q)^2) / dQ
        s_dq[R_ROWS + tid] = -inv_n;               // d (-mean(Q_pi)) / dQ_pi
        s_dq[2 * R_ROWS + tid] = diff * diff;
        if (rank == 0) {
          P.dQ[row] = s_dq[tid];
          P.q_pi[row] = s_qpi[tid * R_DU];
        }
      }
      __syncthreads();
      if (rank == 0 && tid == 0) {
        float ssq = 0.f, sq = 0.f, sth = 0.f;
        for (int r = 0; r < R_ROWS; ++r) {
          ssq += s_dq[2 * R_ROWS + r];
          sq += s_qpi[r * R_DU];
          for (int j = 0; j < d.dimu; ++j) sth += s_th[r * R_DU + j] * s_th[r * R_DU + j];
        }
        float* lp = P.loss_part + (blockIdx.x / R_CS) * 4;
        lp[0] = ssq; lp[1] = sq; lp[2] = sth; lp[3] = 0.f;
      }
      // critic / actor-through-critic gradients at the last hidden layer: dY * Wout^T (Wout is [H][1]),
      // group 0 -> critic chain into tile 0, group 1 -> actor-through-critic chain into tile 1
      cluster.barrier_arrive();                 // E1 (the rowdots above were the last readers of the tiles)
      {
        const float wq = __ldg(P.WoutQ + gcol);
        const float4 m = load_mask(group == 0 ? P.hq_last : P.hqp_last, row0, gcol, ln.rq);
        const float* sd = s_dq + group * R_ROWS + 4 * ln.rq;
        float4 v = mask4(make_float4(sd[0] * wq, sd[1] * wq, sd[2] * wq, sd[3] * wq), m);
        if (group == 0) store_rm(P.dc_last, row0, gcol, ln.rq, v);
        cluster.barrier_wait();
        publish(cluster, smem, group * R_TILE, gcol, ln.rq, v);
      }
      cluster.barrier_arrive();
      cluster.barrier_wait();
    }
    if (post == POST_BIN || post == POST_GOUT_BIN) {
      // gradient wrt the action inputs of main.Q (tile 1 holds the full actor-through-critic gradient at layer 0),
      // then through tanh and the action penalty (ddpg.py:440-441):
      // d pi_loss / d(pre-tanh) = (dL/d(pi/max_u) + action_l2 * 2/(B*dimu) * th) * (1 - th^2)
      rowdot(tiles + 1 * R_TILE, P.W0Q_act, 1, R_H, d.dimu, red);
      __syncthreads();
      rowdot_finish(red, d.dimu, nullptr, s_dy);
      __syncthreads();
      if (tid < R_ROWS * R_DU) {
        const int r = tid >> 3, j = tid & (R_DU - 1);
        float v = 0.f;
        if (j < d.dimu) {
          const float coef = P.action_l2 * 2.0f / (float)(P.n * d.dimu);
          const float th = s_th[tid];
          v = (s_dy[tid] + coef * th) * (1.f - th * th);
        }
        s_dy[tid] = v;
        if (rank == 0 && j < P.lddy) P.dy[(row0 + r) * P.lddy + j] = v;
      }
      __syncthreads();
      // actor gradient at the last hidden layer: dy * Wout_pi^T, into tile 0
      cluster.barrier_arrive();
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (group == 0) {
        const float4 m = load_mask(P.hp_last, row0, gcol, ln.rq);
        float o4[4] = {0.f, 0.f, 0.f, 0.f};
        for (int j = 0; j < d.dimu; ++j) {
          const float wj = __ldg(P.WoutP + (int64_t)gcol * d.dimu + j);
#pragma unroll
          for (int i = 0; i < 4; ++i) o4[i] = fmaf(s_dy[(4 * ln.rq + i) * R_DU + j], wj, o4[i]);
        }
        v = mask4(make_float4(o4[0], o4[1], o4[2], o4[3]), m);
        store_rm(P.dp_last, row0, gcol, ln.rq, v);
      }
      cluster.barrier_wait();
      if (group == 0) publish(cluster, smem, 0, gcol, ln.rq, v);
      cluster.barrier_arrive();
      cluster.barrier_wait();
    }
    R_TL(8 + 8 * st + 6);
  }
  R_TL(3);
}

// ------------------------------------------------------------------------------------------------
// launch 2: all weight gradients (+ optional Adam), loss fold, step counter
// ------------------------------------------------------------------------------------------------
struct DwTail {
  AdamCtx ax;                      // ax.theta == NULL: gradients only
  const float* neg_a_table;
  int table_len;
  int64_t* step_counter;           // may be NULL
  int ring;
  unsigned int* ticket;
  const float* loss_part;
  int n_clusters;
  int64_t n;
  int dimu;
  float action_l2;
  float *q_loss, *pi_loss;
};

__global__ void __launch_bounds__(GEMM_THREADS, 2)
rows_dw_kernel(const __grid_constant__ GemmBatch G, const __grid_constant__ DwTail T) {
  __shared__ __align__(16) float As[2 * TILE_FLOATS];
  __shared__ __align__(16) float Bs[2 * TILE_FLOATS];
  __shared__ GemmProb Ps;
  __shared__ AdamCtx ax;
  __shared__ unsigned int s_last;
  const long long st = T.step_counter ? *T.step_counter : 0;     // value BEFORE this update's bump
  if (threadIdx.x == 0) {
    ax = T.ax;
    if (T.ax.theta != nullptr && T.neg_a_table != nullptr) {
      long long t = st + 1;                                       // Adam's 1-based step of this update
      ax.neg_a = T.neg_a_table[(t <= T.table_len ? t : (long long)T.table_len) - 1];
    }
  }
  // (gemm_run_tile synchronises before the first use of `ax`)
  gemm_run_tile(G, Ps, As, Bs, T.ax.theta != nullptr ? &ax : nullptr);
  // ---- the last CTA to finish folds the loss partials and bumps the step counter
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int t = atomicAdd(T.ticket, 1u);
    s_last = (t == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    float ssq = 0.f, sq = 0.f, sth = 0.f;
    for (int c = 0; c < T.n_clusters; ++c) {
      ssq += T.loss_part[4 * c + 0];
      sq += T.loss_part[4 * c + 1];
      sth += T.loss_part[4 * c + 2];
    }
    const float inv_n = 1.0f / (float)T.n;
    const long long slot = (T.step_counter && T.ring > 0) ? st % T.ring : 0;
    if (T.q_loss) T.q_loss[slot] = ssq * inv_n;                                                   // ddpg.py:439
    if (T.pi_loss) T.pi_loss[slot] = -sq * inv_n + T.action_l2 * sth / (float)(T.n * T.dimu);     // ddpg.py:440-441
    if (T.step_counter) *T.step_counter = st + 1;
    *T.ticket = 0u;
    __threadfence();
  }
}

// ------------------------------------------------------------------------------------------------
struct RowsWorkspace {
  unsigned int* ticket;
  float* loss_part;
  float *Xp, *Xq;
  float *hp[R_MAXL], *hq[R_MAXL], *hqp[R_MAXL], *dc[R_MAXL], *dp[R_MAXL];   // row-major [n][256]
  float *dQ, *dy;
  int KP, lddy;
  int64_t total;
};

static RowsWorkspace carve_rows(const cur_net_desc& d, int64_t n, float* base) {
  RowsWorkspace w;
  const NetLayout q = net_layout(d, 0);
  const int K0q = q.in_s + q.in_g;
  w.KP = ((K0q + 63) / 64) * 64;
  w.lddy = (int)r4(d.dimu);
  int64_t o = 0;
  auto take = [&](int64_t floats) {
    float* ptr = base ? base + o : nullptr;
    o += r4(floats);
    return ptr;
  };
  w.ticket = reinterpret_cast<unsigned int*>(take(4));
  w.loss_part = take((n / R_ROWS) * 4);
  w.Xp = take(n * w.KP);
  w.Xq = take(n * w.KP);
  for (int l = 0; l < d.layers; ++l) {
    w.hp[l] = take(n * R_H); w.hq[l] = take(n * R_H); w.hqp[l] = take(n * R_H);
    w.dc[l] = take(n * R_H); w.dp[l] = take(n * R_H);
  }
  w.dQ = take(n);
  w.dy = take(n * w.lddy);
  w.total = o;
  return w;
}

static bool rows_supported(const cur_net_desc* d, int64_t n) {
  if (check_desc(d) != CUR_OK) return false;
  if (d->hidden != R_H || d->layers < 1 || d->layers > R_MAXL) return false;
  if (d->dimu > R_DU || n <= 0 || (n % R_ROWS) != 0 || n >= (1 << 24)) return false;
  const NetLayout q = net_layout(*d, 0);
  if (q.in_s + q.in_g > R_H) return false;
  return true;
}

}  // namespace cur

using namespace cur;

extern "C" int cur_ddpg_rows_supported(const cur_net_desc* d, int64_t batch) { return rows_supported(d, batch) ? 1 : 0; }

extern "C" int64_t cur_ddpg_rows_workspace_floats(const cur_net_desc* d, int64_t batch) {
  if (!rows_supported(d, batch)) return -1;
  return carve_rows(*d, batch, nullptr).total;
}

extern "C" int cur_ddpg_rows_step(void* stream, const cur_net_desc* d, float* theta_main, const float* theta_target,
                                  const cur_norm_stats* stats, const cur_batch* batch, const cur_ddpg_hyper* h,
                                  float* workspace, float* grads, float* q_loss, float* pi_loss, float* q_pi,
                                  const cur_adam_fused* adam) {
  CUR_TRY(check_desc(d));
  CUR_REQUIRE(theta_main && theta_target && batch && h && workspace && grads && q_pi, "NULL argument");
  CUR_REQUIRE(batch->o && batch->g && batch->u && batch->o_2 && batch->g_2 && batch->r, "NULL batch array");
  CUR_REQUIRE(!d->modular || batch->td, "task_descr required for a modular net");
  CUR_REQUIRE(rows_supported(d, batch->n), "shape not supported by the rows schedule (see cur_ddpg_rows_supported)");
  if (d->normalize_obs)
    CUR_REQUIRE(stats && stats->o_mean && stats->o_std && stats->g_mean && stats->g_std, "normalizer stats required");
  if (adam) {
    CUR_REQUIRE(adam->m && adam->v && adam->neg_a_table && adam->table_len > 0, "incomplete Adam block");
    CUR_REQUIRE(h->step_counter != nullptr, "fused Adam needs the device step counter");
  }
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t n = batch->n;
  const NetLayout LQ = net_layout(*d, 0), LP = net_layout(*d, 1);
  const int64_t offP = r4(LQ.total);
  const float *mQ = theta_main, *mP = theta_main + offP, *tQ = theta_target, *tP = theta_target + offP;
  float *gQ = grads, *gP = grads + offP;
  const RowsWorkspace w = carve_rows(*d, n, workspace);
  const int L = d->layers, H = d->hidden;

  static bool configured = false;
  const size_t smem = R_SMEM_FLOATS * sizeof(float);
  if (!configured) {
    CUR_CUDA_TRY(cudaFuncSetAttribute(ddpg_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }

  RowsParams P;
  memset(&P, 0, sizeof(P));
  P.d = *d;
  P.in_sp = LP.in_s; P.in_sq = LQ.in_s; P.in_g = LQ.in_g; P.KP = w.KP; P.L = L;
  P.n = n;
  P.o = batch->o; P.g = batch->g; P.u = batch->u; P.td = batch->td; P.o_2 = batch->o_2; P.g_2 = batch->g_2;
  P.r = batch->r;
  if (stats) { P.o_mean = stats->o_mean; P.o_std = stats->o_std; P.g_mean = stats->g_mean; P.g_std = stats->g_std; }
  P.gamma = h->gamma; P.clip_return = h->clip_return; P.action_l2 = h->action_l2; P.clip_pos = h->clip_pos_returns;
  P.WoutP = mP + LP.off_Wout; P.boutP = mP + LP.off_bout; P.WoutPT = tP + LP.off_Wout; P.boutPT = tP + LP.off_bout;
  P.WoutQ = mQ + LQ.off_Wout; P.boutQ = mQ + LQ.off_bout; P.WoutQT = tQ + LQ.off_Wout; P.boutQT = tQ + LQ.off_bout;
  P.W0Q_act = mQ + LQ.off_W0 + (int64_t)LP.in_s * H;
  P.Xp = w.Xp; P.Xq = w.Xq; P.dQ = w.dQ; P.dy = w.dy; P.lddy = w.lddy;
  P.hq_last = w.hq[L - 1]; P.hqp_last = w.hqp[L - 1]; P.hp_last = w.hp[L - 1];
  P.dc_last = w.dc[L - 1]; P.dp_last = w.dp[L - 1];
  P.loss_part = w.loss_part; P.q_pi = q_pi;

  // ---- steps and net-layer weight descriptors, in consumption order
  int nw = 0, ns = 0;
  auto first_layer = [&](const float* th, const NetLayout& NL) {
    WDesc& D = P.wd[nw++];
    D.w = th + NL.off_W0; D.w2 = th + NL.off_W0g; D.split = NL.in_s; D.kvalid = NL.in_s + NL.in_g;
    D.kper = w.KP / R_NSLICE; D.kind = 2;
  };
  auto hidden_layer = [&](const float* th, const NetLayout& NL, int l, int bwd) {
    WDesc& D = P.wd[nw++];
    D.w = th + NL.off_W[l]; D.w2 = nullptr; D.split = H; D.kvalid = H; D.kper = H / R_NSLICE; D.kind = bwd ? 1 : 0;
  };
  auto bias = [&](const float* th, const NetLayout& NL, int l) { return th + (l == 0 ? NL.off_b0 : NL.off_b[l]); };
  for (int l = 0; l < L; ++l) {                       // forward 1: main.pi | target.pi | main.Q(u)
    RStep& S = P.steps[ns++];
    S.nA = 3; S.bwd = 0; S.shared_w = 0; S.kper = (l == 0 ? w.KP : H) / R_NSLICE;
    S.post = (l == L - 1) ? POST_FOUT : POST_NONE;
    S.aux[0] = bias(mP, LP, l); S.aux[1] = bias(tP, LP, l); S.aux[2] = bias(mQ, LQ, l);
    S.out_rm[0] = w.hp[l]; S.out_rm[1] = nullptr; S.out_rm[2] = w.hq[l];
    if (l == 0) { first_layer(mP, LP); first_layer(tP, LP); first_layer(mQ, LQ); }
    else { hidden_layer(mP, LP, l, 0); hidden_layer(tP, LP, l, 0); hidden_layer(mQ, LQ, l, 0); }
  }
  for (int l = 0; l < L; ++l) {                       // forward 2: main.Q(pi) | target.Q
    RStep& S = P.steps[ns++];
    S.nA = 2; S.bwd = 0; S.shared_w = 0; S.kper = (l == 0 ? w.KP : H) / R_NSLICE;
    S.post = (l == L - 1) ? (L == 1 ? POST_GOUT_BIN : POST_GOUT) : POST_NONE;
    S.aux[0] = bias(mQ, LQ, l); S.aux[1] = bias(tQ, LQ, l);
    S.out_rm[0] = w.hqp[l]; S.out_rm[1] = nullptr;
    if (l == 0) { first_layer(mQ, LQ); first_layer(tQ, LQ); }
    else { hidden_layer(mQ, LQ, l, 0); hidden_layer(tQ, LQ, l, 0); }
  }
  for (int l = L - 1; l >= 1; --l) {                  // backward 1: critic | actor-through-critic share main.Q's W_l
    RStep& S = P.steps[ns++];
    S.nA = 2; S.bwd = 1; S.shared_w = 1; S.kper = H / R_NSLICE;
    S.post = (l == 1) ? POST_BIN : POST_NONE;
    S.aux[0] = w.hq[l - 1]; S.aux[1] = w.hqp[l - 1];
    S.out_rm[0] = w.dc[l - 1]; S.out_rm[1] = nullptr;
    hidden_layer(mQ, LQ, l, 1);
  }
  for (int l = L - 1; l >= 1; --l) {                  // backward 2: actor
    RStep& S = P.steps[ns++];
    S.nA = 1; S.bwd = 1; S.shared_w = 1; S.kper = H / R_NSLICE; S.post = POST_NONE;
    S.aux[0] = w.hp[l - 1];
    S.out_rm[0] = w.dp[l - 1];
    hidden_layer(mP, LP, l, 1);
  }
  P.nw = nw; P.nsteps = ns;

  static long long* tl_dev = nullptr;
  static int tl_calls = 0;
  const bool tl_on = getenv("CUR_ROWS_TIMELINE") != nullptr;
  if (tl_on && tl_dev == nullptr) CUR_CUDA_TRY(cudaMalloc(&tl_dev, 256 * sizeof(long long)));
  P.tl = tl_on ? tl_dev : nullptr;

  const unsigned int n_clusters = (unsigned int)(n / R_ROWS);
  ddpg_rows_kernel<<<n_clusters * R_CS, R_THREADS, smem, s>>>(P);
  CUR_CHECK_LAUNCH();
  if (tl_on && ++tl_calls == 40) {          // debug only: print one warmed-up timeline of CTA 0
    long long h_tl[256];
    CUR_CUDA_TRY(cudaStreamSynchronize(s));
    CUR_CUDA_TRY(cudaMemcpy(h_tl, tl_dev, sizeof(h_tl), cudaMemcpyDeviceToHost));
    fprintf(stderr, "[rows timeline] prologue load_w+build_x %lld, first sync %lld, total %lld cycles\n",
            h_tl[1] - h_tl[0], h_tl[2] - h_tl[1], h_tl[3] - h_tl[0]);
    long long prev = h_tl[2];
    for (int st = 0; st < ns; ++st) {
      const long long* t = h_tl + 8 + 8 * st;
      fprintf(stderr, "[rows timeline] step %2d nA=%d bwd=%d post=%d: gemm %6lld sync %5lld reduce %5lld E1wait %5lld publish %5lld E2 %5lld post %6lld\n",
              st, P.steps[st].nA, P.steps[st].bwd, P.steps[st].post, t[0] - prev, t[1] - t[0], t[2] - t[1], t[3] - t[2],
              t[4] - t[3], t[5] - t[4], t[6] - t[5]);
      prev = t[6];
    }
  }

  // ---- launch 2: weight gradients
  GemmBatch G;
  G.n = 0; G.total_tiles = 0;
  auto add = [&](const GemmProb& p) { G.p[G.n++] = p; };
  auto net_grads = [&](const NetLayout& NL, float* gN, const float* X0, float* const* hN, float* const* dN,
                       const float* dOut, int lddo) {
    add(bwd_dw(hN[L - 1], H, H, dOut, lddo, NL.out, gN + NL.off_Wout, n));
    add(bwd_db(dOut, lddo, NL.out, gN + NL.off_bout, n));
    for (int l = L - 1; l >= 1; --l) {
      add(bwd_dw(hN[l - 1], H, H, dN[l], H, H, gN + NL.off_W[l], n));
      add(bwd_db(dN[l], H, H, gN + NL.off_b[l], n));
    }
    add(bwd_dw(X0, w.KP, NL.in_s, dN[0], H, H, gN + NL.off_W0, n));
    add(bwd_db(dN[0], H, H, gN + NL.off_b0, n));
    if (NL.in_g > 0) add(bwd_dw(X0 + NL.in_s, w.KP, NL.in_g, dN[0], H, H, gN + NL.off_W0g, n));
  };
  net_grads(LQ, gQ, w.Xq, w.hq, w.dc, w.dQ, 1);
  net_grads(LP, gP, w.Xp, w.hp, w.dp, w.dy, w.lddy);
  const int tiles = plan_gemm_batch(G);

  DwTail T;
  memset(&T, 0, sizeof(T));
  if (adam) {
    T.ax.grads = grads; T.ax.theta = theta_main; T.ax.m = adam->m; T.ax.v = adam->v;
    T.ax.b1 = (float)adam->beta1; T.ax.omb1 = (float)(1.0 - adam->beta1);
    T.ax.b2 = (float)adam->beta2; T.ax.omb2 = (float)(1.0 - adam->beta2);
    T.ax.eps = (float)adam->eps;
    T.neg_a_table = adam->neg_a_table; T.table_len = adam->table_len;
  }
  T.step_counter = h->step_counter; T.ring = h->loss_ring;
  T.ticket = w.ticket; T.loss_part = w.loss_part; T.n_clusters = (int)n_clusters; T.n = n; T.dimu = d->dimu;
  T.action_l2 = h->action_l2; T.q_loss = q_loss; T.pi_loss = pi_loss;
  rows_dw_kernel<<<tiles, GEMM_THREADS, 0, s>>>(G, T);
  CUR_CHECK_LAUNCH();
  return CUR_OK;
}
