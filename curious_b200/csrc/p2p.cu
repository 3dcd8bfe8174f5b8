// Data-parallel gradient exchange over NVLink peer memory, fused with the optimiser step (sm_100a).
//
// Replaces (reference flowersteam/curious):
//   baselines/common/mpi_adam.py:24-28   comm.Allreduce(localg, globalg, op=MPI.SUM)   (scale_grad_by_procs=False)
//   baselines/common/mpi_adam.py:30-35   Adam on the summed gradient
//
// One process per GPU.  Every rank owns a "symmetric" region (cudaMalloc + CUDA IPC, mapped into every peer):
//     [ flags: CUR_MAX_RANKS x u64 | gradient arena, buffer 0 | gradient arena, buffer 1 ]
// The weight-gradient launch of update s writes its flat gradient into buffer (s & 1) of its own region.  Then ONE
// kernel per rank (capturable in the update's CUDA graph, no host involvement, no NCCL):
//   1. tells every peer "my gradient of update s is complete" (st.release.sys of s into the peer's flag slot),
//   2. waits until all peers said the same (ld.acquire.sys on its own flags),
//   3. sums the world's gradients straight out of the peers' memory (ld.global over NVLink, 16-byte vectors)
//      in rank order 0..W-1 - every rank gets bit-identical sums - and applies Adam to its own parameter copy.
// Two gradient buffers make a second barrier unnecessary: buffer (s & 1) is next overwritten by update s + 2,
// which a rank can only reach after every peer signalled s + 1, i.e. after it finished reading update s.
//
// Sharded variant (p2p_sharded_adam_kernel, opt-in): reading W - 1 full gradients per rank is 8.3 MB at 8 GPUs.
// Here rank r owns slice r of the arena: it sums ONLY its slice out of the
// peers' gradient buffers (reduce-scatter by loads), steps Adam for that slice (m / v of other slices are never touched
// on this rank) and stores the new parameters of the slice into a staging arena in EVERY peer's region (all-gather by
// stores); a second flag round tells the peers the slice has landed and every rank copies the staged slices into its
// own parameter vector.  Per rank: ~(W-1)/W x 1.18 MB of peer loads + the same of peer stores, two flag rounds.
// Parameters are bit-identical on all ranks by construction (one owner computes each element).  Measured: no faster
// than the one-round kernel (2 GPUs 91 vs 83 us / update, 8 GPUs 98.2 vs 97.3): at 1.18 MB the exchange is bound by
// flag latency and by waiting for the slowest rank, not by NVLink bytes - hence opt-in.
// Traffic per rank and update: (W - 1) x 1.18 MB of peer loads (8 GPUs: 8.3 MB ~ 11 us at the measured
// 770 GB/s per direction) - at this size the exchange is latency-, not bandwidth-bound, which is why it is one
// kernel with one flag round instead of a ring.
#include <string.h>

#include "common.cuh"

namespace cur {

constexpr int P2P_FLAG_BYTES = 256;     // u64 grad flags [8] | u64 theta flags [8] | u32 ticket | pad

struct P2PParams {
  int rank, world;
  const unsigned char* region[CUR_MAX_RANKS];   // region[r] as mapped in THIS process (region[rank] is local)
  int64_t arena;                                // floats per gradient buffer (multiple of 4)
  float *theta, *m, *v;
  const float* neg_a_table;
  int table_len;
  const int64_t* step_counter;                  // already bumped for this update (Adam's 1-based t = counter / step_div)
  int step_div;
  float b1, omb1, b2, omb2, eps;
  int* error_flag;                              // set to 1 if the peers did not show up in time
  // optional: H x H blocks of the arena whose stepped values are also written TRANSPOSED to dst (the backward
  // operands W^T of the rows schedule, see cur_p2p_transposes)
  int n_t, H;
  int64_t t_begin4[CUR_P2P_MAX_TRANSPOSES];     // first float4 of the block inside the arena
  float* t_dst[CUR_P2P_MAX_TRANSPOSES];
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer_f4(const float4* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256) p2p_allreduce_adam_kernel(const __grid_constant__ P2PParams P) {
  __shared__ int s_ok;
  const long long t = *P.step_counter / P.step_div;            // update number, 1-based
  const unsigned long long want = (unsigned long long)t;
  if (blockIdx.x == 0 && threadIdx.x < P.world && (int)threadIdx.x != P.rank) {
    // my gradient of update t was written by the previous kernel on this stream: publish
    unsigned long long* peer_flags =
        reinterpret_cast<unsigned long long*>(const_cast<unsigned char*>(P.region[threadIdx.x]));
    __threadfence_system();
    st_release_sys(peer_flags + P.rank, want);
  }
  if (threadIdx.x == 0) s_ok = 1;
  __syncthreads();
  if (threadIdx.x < P.world && (int)threadIdx.x != P.rank) {
    const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(P.region[P.rank]);
    long long spins = 0;
    while (ld_acquire_sys(mine + threadIdx.x) < want) {
      if (++spins > (1ll << 25)) {           // a peer that never shows up: fail the launch (no partial update of theta / m / v)
        if (P.error_flag) *P.error_flag = 1;
        __threadfence_system();
        __trap();
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
  if (!s_ok) return;
  float neg_a = P.neg_a_table[(t <= P.table_len ? (t < 1 ? 1 : t) : (long long)P.table_len) - 1];
  const int64_t buf_off = P2P_FLAG_BYTES + (int64_t)(t & 1) * P.arena * 4;
  const int64_t n4 = P.arena >> 2;
  float4* th4 = reinterpret_cast<float4*>(P.theta);
  float4* m4 = reinterpret_cast<float4*>(P.m);
  float4* v4 = reinterpret_cast<float4*>(P.v);
  // one element group: sum over ranks in fixed order (bit-identical on every rank), Adam, returns the new parameters
  auto step4 = [&](int64_t i) {
    // all ranks' loads are issued before the first add: a runtime-length loop would serialise W - 1 NVLink round
    // trips per element group (~1.5 us each)
    float4 x[CUR_MAX_RANKS];
#pragma unroll
    for (int r = 0; r < CUR_MAX_RANKS; ++r)
      if (r < P.world) x[r] = ld_peer_f4(reinterpret_cast<const float4*>(P.region[r] + buf_off) + i);
    float4 T = th4[i], M = m4[i], V = v4[i];
    float4 g = x[0];
#pragma unroll
    for (int r = 1; r < CUR_MAX_RANKS; ++r)
      if (r < P.world) { g.x = __fadd_rn(g.x, x[r].x); g.y = __fadd_rn(g.y, x[r].y); g.z = __fadd_rn(g.z, x[r].z); g.w = __fadd_rn(g.w, x[r].w); }
    adam_elem(T.x, g.x, M.x, V.x, neg_a, P.b1, P.omb1, P.b2, P.omb2, P.eps);
    adam_elem(T.y, g.y, M.y, V.y, neg_a, P.b1, P.omb1, P.b2, P.omb2, P.eps);
    adam_elem(T.z, g.z, M.z, V.z, neg_a, P.b1, P.omb1, P.b2, P.omb2, P.eps);
    adam_elem(T.w, g.w, M.w, V.w, neg_a, P.b1, P.omb1, P.b2, P.omb2, P.eps);
    th4[i] = T; m4[i] = M; v4[i] = V;
    return T;
  };
  const int64_t blk4 = (int64_t)P.H * P.H / 4;                 // float4s of one H x H block
  // ---- everything outside the transposed blocks: linear
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    bool in_block = false;
    for (int b = 0; b < P.n_t; ++b) in_block |= (i >= P.t_begin4[b] && i < P.t_begin4[b] + blk4);
    if (!in_block) step4(i);
  }
  // ---- the H x H blocks by 32 x 32 tiles: stepped values go to theta and, turned in shared memory, to W^T
  if (P.n_t > 0) {
    __shared__ float tile[32][33];
    const int tiles_per_row = P.H / 32, tiles = tiles_per_row * tiles_per_row;
    const int ty = threadIdx.x >> 3, tx = threadIdx.x & 7;
    for (int w = blockIdx.x; w < P.n_t * tiles; w += gridDim.x) {
      const int b = w / tiles, t = w - b * tiles;
      const int tr = t / tiles_per_row, tc = t - tr * tiles_per_row;
      const float4 T = step4(P.t_begin4[b] + (int64_t)(tr * 32 + ty) * (P.H / 4) + tc * 8 + tx);
      tile[4 * tx + 0][ty] = T.x; tile[4 * tx + 1][ty] = T.y; tile[4 * tx + 2][ty] = T.z; tile[4 * tx + 3][ty] = T.w;
      __syncthreads();
      // row ty of the turned tile = column tc * 32 + ty of W; its 32 values are rows tr * 32 .. + 31 of W
      *reinterpret_cast<float4*>(P.t_dst[b] + (int64_t)(tc * 32 + ty) * P.H + tr * 32 + 4 * tx) =
          make_float4(tile[ty][4 * tx], tile[ty][4 * tx + 1], tile[ty][4 * tx + 2], tile[ty][4 * tx + 3]);
      __syncthreads();
    }
  }
}


// ---------------------------------------------------------------------------------------------- sharded variant
__device__ __forceinline__ void st_peer_f4(float4* p, float4 v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ bool wait_flags(const unsigned long long* flags, int world, int rank, unsigned long long want,
                                           int* error_flag) {
  __shared__ int s_ok2;
  if (threadIdx.x == 0) s_ok2 = 1;
  __syncthreads();
  if ((int)threadIdx.x < world && (int)threadIdx.x != rank) {
    long long spins = 0;
    while (ld_acquire_sys(flags + threadIdx.x) < want) {
      if (++spins > (1ll << 25)) {
        if (error_flag) *error_flag = 1;
        __threadfence_system();
        __trap();
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
  const bool ok = s_ok2 != 0;
  __syncthreads();
  return ok;
}

__global__ void __launch_bounds__(256) p2p_sharded_adam_kernel(const __grid_constant__ P2PParams P) {
  const long long t = *P.step_counter / P.step_div;            // update number, 1-based
  const unsigned long long want = (unsigned long long)t;
  unsigned char* own = const_cast<unsigned char*>(P.region[P.rank]);
  unsigned long long* own_flags = reinterpret_cast<unsigned long long*>(own);
  unsigned int* ticket = reinterpret_cast<unsigned int*>(own + 128);
  // ---- round 1: gradients of update t are complete everywhere
  if (blockIdx.x == 0 && (int)threadIdx.x < P.world && (int)threadIdx.x != P.rank) {
    __threadfence_system();
    st_release_sys(reinterpret_cast<unsigned long long*>(const_cast<unsigned char*>(P.region[threadIdx.x])) + P.rank, want);
  }
  if (!wait_flags(own_flags, P.world, P.rank, want, P.error_flag)) return;
  // ---- my slice: sum over ranks (fixed order), Adam, publish the new parameters to every rank's staging arena
  const float neg_a = P.neg_a_table[(t <= P.table_len ? (t < 1 ? 1 : t) : (long long)P.table_len) - 1];
  const int64_t n4 = P.arena >> 2;
  const int64_t per = (n4 + P.world - 1) / P.world;
  const int64_t lo = per * P.rank, hi = (lo + per < n4) ? lo + per : n4;
  const int64_t gbuf = P2P_FLAG_BYTES + (int64_t)(t & 1) * P.arena * 4;
  const int64_t stage = P2P_FLAG_BYTES + 2 * P.arena * 4;
  float4* th4 = reinterpret_cast<float4*>(P.theta);
  float4* m4 = reinterpret_cast<float4*>(P.m);
  float4* v4 = reinterpret_cast<float4*>(P.v);
  for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x) {
    float4 x[CUR_MAX_RANKS];
#pragma unroll
    for (int r = 0; r < CUR_MAX_RANKS; ++r)
      if (r < P.world) x[r] = ld_peer_f4(reinterpret_cast<const float4*>(P.region[r] + gbuf) + i);
    float4 T = th4[i], M = m4[i], V = v4[i];
    float4 g = x[0];
#pragma unroll
    for (int r = 1; r < CUR_MAX_RANKS; ++r)
      if (r < P.world) { g.x = __fadd_rn(g.x, x[r].x); g.y = __fadd_rn(g.y, x[r].y); g.z = __fadd_rn(g.z, x[r].z); g.w = __fadd_rn(g.w, x[r].w); }
    adam_elem(T.x, g.x, M.x, V.x, neg_a, P.b1, P.omb1, P.b2, P.omb2, P.eps);
    adam_elem(T.y, g.y, M.y, V.y, neg_a, P.b1, P.omb1, P.b2, P.omb2, P.eps);
    adam_elem(T.z, g.z, M.z, V.z, neg_a, P.b1, P.omb1, P.b2, P.omb2, P.eps);
    adam_elem(T.w, g.w, M.w, V.w, neg_a, P.b1, P.omb1, P.b2, P.omb2, P.eps);
    th4[i] = T; m4[i] = M; v4[i] = V;
    for (int r = 0; r < P.world; ++r)
      if (r != P.rank) st_peer_f4(reinterpret_cast<float4*>(const_cast<unsigned char*>(P.region[r]) + stage) + i, T);
  }
  // ---- round 2: the last block of this rank to finish its stores tells the peers "slice `rank` of update t has landed"
  __shared__ unsigned int s_last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (s_last) {
    if ((int)threadIdx.x < P.world && (int)threadIdx.x != P.rank) {
      __threadfence_system();
      st_release_sys(reinterpret_cast<unsigned long long*>(const_cast<unsigned char*>(P.region[threadIdx.x])) + 8 + P.rank, want);
    }
    if (threadIdx.x == 0) *ticket = 0u;
  }
  if (!wait_flags(own_flags + 8, P.world, P.rank, want, P.error_flag)) return;
  // ---- copy the other ranks' slices from my staging arena into my parameter vector
  const float4* st4 = reinterpret_cast<const float4*>(own + stage);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    if (i >= lo && i < hi) continue;
    th4[i] = ld_peer_f4(st4 + i);
  }
}

}  // namespace cur

using namespace cur;

extern "C" int64_t cur_p2p_region_bytes(int64_t arena_floats) {
  if (arena_floats <= 0 || (arena_floats & 3)) return -1;
  return P2P_FLAG_BYTES + 3 * arena_floats * 4;      // flags | grads 0 | grads 1 | parameter staging
}

// region of the tile-level exchange of ddpg_rows.cu: [ partial slots: world x arena x 8 B | result slots: arena x 8 B ]
extern "C" int64_t cur_xchg_region_bytes(int64_t arena_floats, int world) {
  if (arena_floats <= 0 || world < 1 || world > CUR_MAX_RANKS) return -1;
  return ((int64_t)world + 1) * arena_floats * 8;
}

extern "C" int cur_p2p_alloc(int64_t bytes, void** ptr, unsigned char* handle64) {
  CUR_REQUIRE(ptr && handle64 && bytes > 0, "bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  CUR_CUDA_TRY(cudaMalloc(ptr, (size_t)bytes));
  CUR_CUDA_TRY(cudaMemset(*ptr, 0, (size_t)bytes));
  cudaIpcMemHandle_t h;
  CUR_CUDA_TRY(cudaIpcGetMemHandle(&h, *ptr));
  memcpy(handle64, &h, 64);
  return CUR_OK;
}

extern "C" int cur_p2p_open(const unsigned char* handle64, void** ptr) {
  CUR_REQUIRE(ptr && handle64, "bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  CUR_CUDA_TRY(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return CUR_OK;
}

extern "C" int cur_p2p_close(void* ptr) {
  CUR_CUDA_TRY(cudaIpcCloseMemHandle(ptr));
  return CUR_OK;
}

extern "C" int cur_p2p_zero(void* stream, void* ptr, int64_t bytes) {
  CUR_REQUIRE(ptr && bytes > 0, "bad argument");
  CUR_CUDA_TRY(cudaMemsetAsync(ptr, 0, (size_t)bytes, (cudaStream_t)stream));
  return CUR_OK;
}

extern "C" int cur_p2p_free(void* ptr) {
  CUR_CUDA_TRY(cudaFree(ptr));
  return CUR_OK;
}

static int p2p_launch(void* stream, const cur_p2p_ctx* ctx, float* theta, float* m, float* v, const float* neg_a_table,
                      int table_len, const int64_t* step_counter, double beta1, double beta2, double eps,
                      int32_t* error_flag, bool sharded, const cur_p2p_transposes* tr = nullptr) {
  CUR_REQUIRE(ctx && theta && m && v && neg_a_table && step_counter && table_len > 0, "NULL argument");
  CUR_REQUIRE(ctx->world >= 1 && ctx->world <= CUR_MAX_RANKS && ctx->rank >= 0 && ctx->rank < ctx->world, "bad rank/world");
  CUR_REQUIRE(ctx->arena > 0 && (ctx->arena & 3) == 0, "arena must be a positive multiple of 4 floats");
  P2PParams P;
  memset(&P, 0, sizeof(P));
  P.rank = ctx->rank; P.world = ctx->world; P.arena = ctx->arena;
  for (int r = 0; r < ctx->world; ++r) {
    CUR_REQUIRE(ctx->region[r] != nullptr, "peer region not mapped");
    P.region[r] = reinterpret_cast<const unsigned char*>(ctx->region[r]);
  }
  P.theta = theta; P.m = m; P.v = v; P.neg_a_table = neg_a_table; P.table_len = table_len;
  P.step_counter = step_counter; P.step_div = ctx->step_div > 1 ? ctx->step_div : 1;
  P.b1 = (float)beta1; P.omb1 = (float)(1.0 - beta1); P.b2 = (float)beta2; P.omb2 = (float)(1.0 - beta2);
  P.eps = (float)eps; P.error_flag = error_flag;
  P.H = 32;
  if (tr != nullptr && tr->n > 0) {
    CUR_REQUIRE(!sharded, "the sharded exchange does not maintain transposes");
    CUR_REQUIRE(tr->n <= CUR_P2P_MAX_TRANSPOSES && tr->H >= 32 && (tr->H % 32) == 0, "bad transposes table");
    P.n_t = tr->n; P.H = tr->H;
    for (int b = 0; b < tr->n; ++b) {
      CUR_REQUIRE(tr->begin[b] >= 0 && (tr->begin[b] & 3) == 0 && tr->begin[b] + (int64_t)tr->H * tr->H <= ctx->arena &&
                  tr->dst[b] != nullptr, "bad transposes entry");
      P.t_begin4[b] = tr->begin[b] >> 2; P.t_dst[b] = tr->dst[b];
    }
  }
  const int64_t n4 = ctx->arena >> 2;
  int blocks = (int)((n4 + 255) / 256);
  const int cap = 2 * sm_count();          // co-resident (256 threads, no shared memory): all blocks spin on the flags
  if (blocks > cap) blocks = cap;
  if (sharded && ctx->world > 1) {
    // all blocks spin on flags: they must be co-resident (blocks <= SM count holds) and few enough that the slice loop
    // still has work for each of them
    int64_t per = (n4 + ctx->world - 1) / ctx->world;
    int b2 = (int)((per + 255) / 256);
    if (b2 > cap) b2 = cap;
    if (b2 < 1) b2 = 1;
    p2p_sharded_adam_kernel<<<b2, 256, 0, (cudaStream_t)stream>>>(P);
  } else {
    p2p_allreduce_adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(P);
  }
  CUR_CHECK_LAUNCH();
  return CUR_OK;
}

extern "C" int cur_p2p_allreduce_adam(void* stream, const cur_p2p_ctx* ctx, float* theta, float* m, float* v,
                                      const float* neg_a_table, int table_len, const int64_t* step_counter,
                                      double beta1, double beta2, double eps, int32_t* error_flag) {
  return p2p_launch(stream, ctx, theta, m, v, neg_a_table, table_len, step_counter, beta1, beta2, eps, error_flag, false);
}

extern "C" int cur_p2p_allreduce_adam_t(void* stream, const cur_p2p_ctx* ctx, float* theta, float* m, float* v,
                                        const float* neg_a_table, int table_len, const int64_t* step_counter,
                                        double beta1, double beta2, double eps, int32_t* error_flag,
                                        const cur_p2p_transposes* transposes) {
  return p2p_launch(stream, ctx, theta, m, v, neg_a_table, table_len, step_counter, beta1, beta2, eps, error_flag, false,
                    transposes);
}

extern "C" int cur_p2p_sharded_adam(void* stream, const cur_p2p_ctx* ctx, float* theta, float* m, float* v,
                                    const float* neg_a_table, int table_len, const int64_t* step_counter,
                                    double beta1, double beta2, double eps, int32_t* error_flag) {
  return p2p_launch(stream, ctx, theta, m, v, neg_a_table, table_len, step_counter, beta1, beta2, eps, error_flag, true);
}
