// Replay storage + fused HER relabel / gather / reward / preprocess kernel (sm_100a).
//
// Replaces (reference flowersteam/curious, paths relative to its root):
//   baselines/her/replay_buffer.py:57-72      store_episode           -> cur_store_episodes
//   baselines/her/her.py:20-66, 99-183        _sample_her_transitions -> cur_her_sample
//   baselines/her/ddpg.py:325-345, 350-353    per-buffer loop, concat, shuffle, _preprocess_og
//   config.py:158-159 reward_fun              restated module-distance reward
//
// Kernel design (HBM bound, one ~0.5 KB gather per transition; ncu history in profiles/):
//   * one CTA = TILE consecutive OUTPUT rows, 4 warps;
//   * warp 0 decodes the per-row draws (injected stream or Philox4x32-10), one lane per row, and
//     publishes three source addresses per row (main span, future achieved goal, cold row);
//   * all warps then copy global -> shared with per-lane 16-byte cp.async (SASS LDGSTS, L2-only
//     `.cg`): thanks to the shifted hot-row layout the whole transition is ONE contiguous span, so a
//     single warp-wide LDGSTS moves an Arm4 transition (28 lanes span + 3 lanes future goal);
//     no register staging, ~16 KB in flight per CTA, up to 14 CTAs per SM;
//   * relabel (g, task_descr), the float64 reward and the clip are evaluated straight out of shared
//     memory and every output array is written as contiguous, fully coalesced 16-byte stores.
#include "her_device.cuh"

namespace cur {

thread_local char g_last_error[256] = {0};

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return 148;
  }
  return cached;
}

constexpr int TILE = 32;         // transitions per CTA
constexpr int HER_THREADS = 128;
constexpr int HER_WARPS = HER_THREADS / 32;

struct HerKernelParams {
  cur_her_args a;
  HerPlan p;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src_gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem)
               : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory");
}

__device__ __forceinline__ float clipf(float x, float c) { return fminf(fmaxf(x, -c), c); }

// Cooperative, coalesced write of `nrows` rows of `dim` floats starting at output row j0.
// src(tr, k) returns element k of tile-row tr.  The output span is contiguous and 128-byte aligned
// (j0 is a multiple of TILE), so dim % 4 == 0 takes the 16-byte path.
template <typename Src4, typename Src1>
__device__ __forceinline__ void emit(float* __restrict__ out, int dim, int64_t j0, int nrows,
                                     Src4 src4, Src1 src1) {
  if (out == nullptr || dim <= 0) return;
  float* dst = out + j0 * (int64_t)dim;
  if ((dim & 3) == 0) {
    const int d4 = dim >> 2;
    const float inv = 1.0f / (float)d4;
    const int n4 = nrows * d4;
    for (int i = threadIdx.x; i < n4; i += HER_THREADS) {
      int tr = __float2int_rz(((float)i + 0.5f) * inv);
      int k4 = i - tr * d4;
      reinterpret_cast<float4*>(dst)[i] = src4(tr, k4 << 2);
    }
  } else {
    const float inv = 1.0f / (float)dim;
    const int n = nrows * dim;
    for (int i = threadIdx.x; i < n; i += HER_THREADS) {
      int tr = __float2int_rz(((float)i + 0.5f) * inv);
      int k = i - tr * dim;
      dst[i] = src1(tr, k);
    }
  }
}

// Two outputs of the same width sharing the index arithmetic (o/o_2, g/g_2, ag/ag_2).
template <typename A4, typename B4, typename A1, typename B1>
__device__ __forceinline__ void emit2(float* __restrict__ outa, float* __restrict__ outb, int dim, int64_t j0,
                                      int nrows, A4 a4, B4 b4, A1 a1, B1 b1) {
  if (outa == nullptr) { emit(outb, dim, j0, nrows, b4, b1); return; }
  if (outb == nullptr) { emit(outa, dim, j0, nrows, a4, a1); return; }
  if (dim <= 0) return;
  float* da = outa + j0 * (int64_t)dim;
  float* db = outb + j0 * (int64_t)dim;
  if ((dim & 3) == 0) {
    const int d4 = dim >> 2;
    const float inv = 1.0f / (float)d4;
    const int n4 = nrows * d4;
    for (int i = threadIdx.x; i < n4; i += HER_THREADS) {
      int tr = __float2int_rz(((float)i + 0.5f) * inv);
      int k = (i - tr * d4) << 2;
      reinterpret_cast<float4*>(da)[i] = a4(tr, k);
      reinterpret_cast<float4*>(db)[i] = b4(tr, k);
    }
  } else {
    const float inv = 1.0f / (float)dim;
    const int n = nrows * dim;
    for (int i = threadIdx.x; i < n; i += HER_THREADS) {
      int tr = __float2int_rz(((float)i + 0.5f) * inv);
      int k = i - tr * dim;
      da[i] = a1(tr, k);
      db[i] = b1(tr, k);
    }
  }
}

__global__ void __launch_bounds__(HER_THREADS)
her_sample_kernel(const __grid_constant__ HerKernelParams P) {
  const cur_her_args& a = P.a;
  const cur_layout& L = a.L;
  const HerPlan& pl = P.p;
  extern __shared__ __align__(128) unsigned char smem_raw[];

  float* stage = reinterpret_cast<float*>(smem_raw);                              // TILE * stage_stride
  // per row: source address of [0] the main span, [1] the future achieved goal (NULL: not a HER
  // row), [2] the cold row (NULL: not requested)
  const float** m_src = reinterpret_cast<const float**>(stage + TILE * pl.stage_stride);    // TILE * 3
  float* rew = reinterpret_cast<float*>(m_src + 3 * TILE);                        // TILE
  int* m_task = reinterpret_cast<int*>(rew + TILE);        // module written to task_descr (-1: keep)
  int* m_relab = m_task + TILE;                            // module whose goal slice is relabelled (-1: none)

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int64_t j0 = (int64_t)blockIdx.x * TILE;
  const int nrows = (int)min((int64_t)TILE, a.batch - j0);
  const int ss = pl.stage_stride;

  // ---------------------------------------------------------------- phase A: draws (warp 0)
  HerRow row;
  row.ft = -1; row.choice = -1; row.ep = 0; row.t = 0; row.ttr = -1; row.her = false;
  if (warp == 0 && lane < nrows) her_draw_row(a, pl, j0 + lane, row, m_src + 3 * lane);
  __syncthreads();

  // ---------------------------------------------------------------- copies: global -> shared
  // Each lane owns a fixed 16-byte chunk slot of the per-transition stage (region, index in region,
  // destination offset are loop invariant); per row it only fetches the region's source address.
  {
    const int per_row = pl.img4 + pl.fut4 + pl.cold4;
    const uint32_t stage_s = smem_u32(stage);
    for (int c0 = 0; c0 < per_row; c0 += 32) {
      const int c = c0 + lane;
      int sel = 0, q = c, doff = 4 * c;
      if (c >= pl.img4 + pl.fut4) { sel = 2; q = c - pl.img4 - pl.fut4; doff = pl.cold_off + 4 * q; }
      else if (c >= pl.img4) { sel = 1; q = c - pl.img4; doff = pl.fut_off + 4 * q; }
      const bool active = c < per_row;
      uint32_t dst = stage_s + 4u * (uint32_t)(warp * ss + doff);
      const uint32_t dstep = 4u * (uint32_t)(HER_WARPS * ss);
      for (int r = warp; r < nrows; r += HER_WARPS, dst += dstep) {
        const float* src = m_src[3 * r + sel];
        if (active && src != nullptr)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + 4 * q) : "memory");
      }
    }
    cp_async_wait_all();
  }
  __syncthreads();

  // ---------------------------------------------------------------- module decisions + reward
  // image indices: row t sections at (off - img_off); row t+1 sections at (i0 + off);
  // g/u/td of step t live in row t+1 (shifted layout)
  const int iG = pl.i0 + L.off_g, iU = pl.i0 + L.off_u, iTD = pl.i0 + L.off_td;
  const int iAG2 = pl.i0 + L.off_ag, iO2 = pl.i0 + L.off_o;
  const int iO = L.off_o - pl.img_off, iAG = L.off_ag - pl.img_off;   // iAG valid only if img_off <= off_ag
  const bool wipe = (a.mode == CUR_MODE_BUFFER || a.mode == CUR_MODE_RANDOM_TASK || a.mode == CUR_MODE_CP_TASK);

  if (warp == 0 && lane < nrows) {
    int relab_out = -1;
    const float rew1 = her_relabel_row(a, pl, stage + lane * ss, row, &relab_out);
    if (a.idx_out) {
      int32_t* io = a.idx_out + (j0 + lane) * 4;
      io[0] = row.ep; io[1] = row.t; io[2] = row.ft;
      io[3] = row.her ? ((a.mode == CUR_MODE_FLAT) ? 0 : relab_out) : -1;
    }
    if (a.r != nullptr) a.r[j0 + lane] = rew1;
  }
  __syncthreads();

  // ---------------------------------------------------------------- coalesced outputs
  const float c = a.clip_obs;
  const bool do_clip = c > 0.0f;
  auto ld4 = [&](int tr, int off) {
    return *reinterpret_cast<const float4*>(stage + tr * ss + off);
  };
  auto clip4 = [&](float4 v) {
    if (do_clip) { v.x = clipf(v.x, c); v.y = clipf(v.y, c); v.z = clipf(v.z, c); v.w = clipf(v.w, c); }
    return v;
  };
  auto clip1 = [&](float v) { return do_clip ? clipf(v, c) : v; };
  auto sub4 = [&](float4 x, float4 y) { return make_float4(x.x - y.x, x.y - y.y, x.z - y.z, x.w - y.w); };

  emit2(a.o, a.o_2, L.dimo, j0, nrows,
        [&](int tr, int k) { return clip4(ld4(tr, iO + k)); },
        [&](int tr, int k) { return clip4(ld4(tr, iO2 + k)); },
        [&](int tr, int k) { return clip1(stage[tr * ss + iO + k]); },
        [&](int tr, int k) { return clip1(stage[tr * ss + iO2 + k]); });
  emit(a.u, L.dimu, j0, nrows,
       [&](int tr, int k) { return ld4(tr, iU + k); },
       [&](int tr, int k) { return stage[tr * ss + iU + k]; });
  emit(a.td, L.dimtd, j0, nrows,
       [&](int tr, int k) { return ld4(tr, iTD + k); },
       [&](int tr, int k) { return stage[tr * ss + iTD + k]; });
  // g / g_2 (ddpg.py:350-353): optional relative goals, then clip
  if (a.relative_goals) {
    emit2(a.g, a.g_2, L.dimg, j0, nrows,
          [&](int tr, int k) { return clip4(sub4(ld4(tr, iG + k), ld4(tr, iAG + k))); },
          [&](int tr, int k) { return clip4(sub4(ld4(tr, iG + k), ld4(tr, iAG2 + k))); },
          [&](int tr, int k) { return clip1(stage[tr * ss + iG + k] - stage[tr * ss + iAG + k]); },
          [&](int tr, int k) { return clip1(stage[tr * ss + iG + k] - stage[tr * ss + iAG2 + k]); });
  } else {
    emit2(a.g, a.g_2, L.dimg, j0, nrows,
          [&](int tr, int k) { return clip4(ld4(tr, iG + k)); },
          [&](int tr, int k) { return clip4(ld4(tr, iG + k)); },
          [&](int tr, int k) { return clip1(stage[tr * ss + iG + k]); },
          [&](int tr, int k) { return clip1(stage[tr * ss + iG + k]); });
  }
  emit2(a.ag, a.ag_2, L.dimag, j0, nrows,
        [&](int tr, int k) { return ld4(tr, iAG + k); },
        [&](int tr, int k) { return ld4(tr, iAG2 + k); },
        [&](int tr, int k) { return stage[tr * ss + iAG + k]; },
        [&](int tr, int k) { return stage[tr * ss + iAG2 + k]; });
  emit(a.change, L.dimchange, j0, nrows,
       [&](int tr, int k) { return ld4(tr, pl.cold_off + L.off_change + k); },
       [&](int tr, int k) { return stage[tr * ss + pl.cold_off + L.off_change + k]; });
  emit(a.info, L.diminfo, j0, nrows,
       [&](int tr, int k) { return ld4(tr, pl.cold_off + L.off_info + k); },
       [&](int tr, int k) { return stage[tr * ss + pl.cold_off + L.off_info + k]; });
}

// ------------------------------------------------------------------------------------------------
// store: pack key-major episodes into hot/cold rows, one CTA per (row, copy)
// ------------------------------------------------------------------------------------------------
struct StoreParams {
  cur_layout L;
  cur_episode_src src;
  int n_copies;
  int32_t copy_src[CUR_MAX_COPIES];
  float* copy_hot[CUR_MAX_COPIES];
  float* copy_cold[CUR_MAX_COPIES];
  int64_t copy_slot[CUR_MAX_COPIES];
};

__global__ void __launch_bounds__(128) store_episodes_kernel(const __grid_constant__ StoreParams S) {
  const cur_layout& L = S.L;
  const int r = blockIdx.x;        // hot row 0..T
  const int cpy = blockIdx.y;
  const int e = S.copy_src[cpy];
  float* dst = S.copy_hot[cpy] + (S.copy_slot[cpy] * (L.T + 1) + r) * (int64_t)L.row_stride;
  const int64_t rprev = (int64_t)e * L.T + (r - 1);        // step whose g/u/td live in this row
  const int64_t rcur = (int64_t)e * (L.T + 1) + r;
  for (int k = threadIdx.x; k < L.row_stride; k += blockDim.x) {
    float v = 0.0f;
    if (k < L.off_ag) {
      if (r > 0) {
        if (k < L.off_u) {
          int j = k - L.off_g;
          if (j < L.dimg) v = S.src.g[rprev * L.dimg + j];
        } else if (k < L.off_td) {
          int j = k - L.off_u;
          if (j < L.dimu) v = S.src.u[rprev * L.dimu + j];
        } else {
          int j = k - L.off_td;
          if (j < L.dimtd && S.src.td) v = S.src.td[rprev * L.dimtd + j];
        }
      }
    } else if (k < L.off_o) {
      int j = k - L.off_ag;
      if (j < L.dimag) v = S.src.ag[rcur * L.dimag + j];
    } else {
      int j = k - L.off_o;
      if (j < L.dimo) v = S.src.o[rcur * L.dimo + j];
    }
    dst[k] = v;
  }
  if (r < L.T && L.cold_stride > 0 && S.copy_cold[cpy] != nullptr) {
    float* cd = S.copy_cold[cpy] + (S.copy_slot[cpy] * L.T + r) * (int64_t)L.cold_stride;
    const int64_t rt = (int64_t)e * L.T + r;
    for (int k = threadIdx.x; k < L.cold_stride; k += blockDim.x) {
      float v = 0.0f;
      if (k < L.off_info) {
        int j = k - L.off_change;
        if (j < L.dimchange && S.src.change) v = S.src.change[rt * L.dimchange + j];
      } else {
        int j = k - L.off_info;
        if (j < L.diminfo && S.src.info) v = S.src.info[rt * L.diminfo + j];
      }
      cd[k] = v;
    }
  }
}

}  // namespace cur

using namespace cur;

extern "C" int cur_abi_version(void) { return CUR_ABI_VERSION; }
extern "C" const char* cur_last_error(void) { return g_last_error; }

extern "C" int cur_device_info(int* sms, int* major, int* minor) {
  int dev = 0;
  CUR_CUDA_TRY(cudaGetDevice(&dev));
  if (sms) CUR_CUDA_TRY(cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, dev));
  if (major) CUR_CUDA_TRY(cudaDeviceGetAttribute(major, cudaDevAttrComputeCapabilityMajor, dev));
  if (minor) CUR_CUDA_TRY(cudaDeviceGetAttribute(minor, cudaDevAttrComputeCapabilityMinor, dev));
  return CUR_OK;
}

extern "C" int cur_layout_init(cur_layout* L, int T, int dimo, int dimag, int dimg, int dimu, int dimtd,
                               int dimchange, int diminfo) {
  CUR_REQUIRE(L != nullptr, "layout is NULL");
  CUR_REQUIRE(T > 0 && dimo > 0 && dimag > 0 && dimg > 0 && dimu > 0, "T and o/ag/g/u dims must be > 0");
  CUR_REQUIRE(dimtd >= 0 && dimchange >= 0 && diminfo >= 0, "negative dim");
  CUR_REQUIRE(dimtd <= CUR_MAX_TASKS, "too many modules");
  L->T = T;
  L->dimo = dimo; L->dimag = dimag; L->dimg = dimg; L->dimu = dimu;
  L->dimtd = dimtd; L->dimchange = dimchange; L->diminfo = diminfo;
  int off = 0;
  L->off_g = off; off += round_up4(dimg);
  L->off_u = off; off += round_up4(dimu);
  L->off_td = off; off += round_up4(dimtd);
  L->off_ag = off; off += round_up4(dimag);
  L->off_o = off; off += round_up4(dimo);
  L->row_stride = off;
  off = 0;
  L->off_change = off; off += round_up4(dimchange);
  L->off_info = off; off += round_up4(diminfo);
  L->cold_stride = off;
  return CUR_OK;
}

extern "C" int cur_store_episodes(void* stream, const cur_layout* L, const cur_episode_src* src, int n_ep,
                                  int n_copies, const int32_t* copy_src, float* const* copy_hot,
                                  float* const* copy_cold, const int64_t* copy_slot) {
  CUR_REQUIRE(L && src && copy_src && copy_hot && copy_slot, "NULL argument");
  CUR_REQUIRE(src->o && src->ag && src->g && src->u, "o/ag/g/u sources are required");
  CUR_REQUIRE(n_copies >= 0 && n_ep > 0, "bad counts");
  CUR_REQUIRE(L->cold_stride == 0 || copy_cold != nullptr, "cold destinations required");
  cudaStream_t s = (cudaStream_t)stream;
  for (int done = 0; done < n_copies; done += CUR_MAX_COPIES) {
    StoreParams S;
    S.L = *L;
    S.src = *src;
    S.n_copies = (n_copies - done < CUR_MAX_COPIES) ? n_copies - done : CUR_MAX_COPIES;
    for (int i = 0; i < S.n_copies; ++i) {
      CUR_REQUIRE(copy_src[done + i] >= 0 && copy_src[done + i] < n_ep, "copy_src out of range");
      CUR_REQUIRE(copy_hot[done + i] != nullptr && copy_slot[done + i] >= 0, "bad destination");
      S.copy_src[i] = copy_src[done + i];
      S.copy_hot[i] = copy_hot[done + i];
      S.copy_cold[i] = copy_cold ? copy_cold[done + i] : nullptr;
      S.copy_slot[i] = copy_slot[done + i];
    }
    dim3 grid(L->T + 1, S.n_copies);
    store_episodes_kernel<<<grid, 128, 0, s>>>(S);
    CUR_CHECK_LAUNCH();
  }
  return CUR_OK;
}

extern "C" int cur_her_sample(void* stream, const cur_her_args* args) {
  CUR_REQUIRE(args != nullptr, "args is NULL");
  const cur_her_args& a = *args;
  CUR_REQUIRE(a.batch >= 0, "negative batch");
  if (a.batch == 0) return CUR_OK;
  CUR_REQUIRE(a.n_segments >= 1 && a.n_segments <= CUR_MAX_SEGMENTS, "n_segments out of range");
  CUR_REQUIRE(a.mode >= CUR_MODE_BUFFER && a.mode <= CUR_MODE_FLAT, "unknown mode");
  CUR_REQUIRE(a.tasks.n_tasks >= 0 && a.tasks.n_tasks <= CUR_MAX_TASKS, "n_tasks out of range");
  CUR_REQUIRE(a.mode == CUR_MODE_FLAT || a.tasks.n_tasks == a.L.dimtd, "n_tasks must equal dimtd");
  CUR_REQUIRE(a.L.row_stride > 0 && (a.L.row_stride & 3) == 0, "layout not initialised");
  int64_t total = 0;
  for (int i = 0; i < a.n_segments; ++i) {
    CUR_REQUIRE(a.seg[i].count >= 0, "negative segment count");
    if (a.dyn != nullptr) {
      CUR_REQUIRE(a.seg[i].base != nullptr, "segment base is NULL");
    } else if (a.seg[i].count > 0) {
      CUR_REQUIRE(a.seg[i].base != nullptr, "segment base is NULL");
      CUR_REQUIRE(a.seg[i].n_episodes > 0, "sampling from an empty buffer (replay_buffer.py:43)");
      CUR_REQUIRE(a.seg[i].task_to_replay < a.tasks.n_tasks, "task_to_replay out of range");
    }
    total += a.seg[i].count;
  }
  CUR_REQUIRE(a.dyn != nullptr || total == a.batch, "segment counts must sum to batch (ddpg.py:323)");
  CUR_REQUIRE(a.dyn == nullptr || a.inj_ep == nullptr, "a device control block implies Philox draws");
  for (int m = 0; m < a.tasks.n_tasks; ++m) {
    CUR_REQUIRE(a.tasks.len[m] >= 0 && a.tasks.len[m] <= CUR_MAX_SLICE, "module slice too long");
    for (int k = 0; k < a.tasks.len[m]; ++k) {
      CUR_REQUIRE(a.tasks.g_idx[m][k] >= 0 && a.tasks.g_idx[m][k] < a.L.dimg, "g index out of range");
      CUR_REQUIRE(a.tasks.ag_idx[m][k] >= 0 && a.tasks.ag_idx[m][k] < a.L.dimag, "ag index out of range");
      if (a.tasks.kind[m] == CUR_REWARD_PAIR)
        CUR_REQUIRE(a.tasks.ref_idx[m][k] >= 0 && a.tasks.ref_idx[m][k] < a.L.dimag, "reference ag index out of range");
    }
    CUR_REQUIRE(a.tasks.kind[m] >= CUR_REWARD_DISTANCE && a.tasks.kind[m] <= CUR_REWARD_INFO, "unknown reward kind");
    if (a.tasks.kind[m] == CUR_REWARD_INFO)
      CUR_REQUIRE(a.tasks.info_col[m] >= 0 && a.tasks.info_col[m] < a.L.diminfo, "info column out of range");
  }
  if (a.inj_ep != nullptr)
    CUR_REQUIRE(a.inj_t && a.inj_u_her && a.inj_u_off, "incomplete injected stream");
  if (a.relative_goals) CUR_REQUIRE(a.L.dimg == a.L.dimag, "relative goals need dimg == dimag");

  HerKernelParams P;
  P.a = a;
  make_plan(a, &P.p);
  if (P.p.cold4 > 0)
    for (int i = 0; i < a.n_segments; ++i)
      CUR_REQUIRE(a.seg[i].count == 0 || a.seg[i].cold != nullptr, "change/info requested but segment has no cold rows");
  size_t smem = (size_t)TILE * P.p.stage_stride * 4 + 3 * TILE * sizeof(void*) + 3 * TILE * 4 + 16;
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    CUR_REQUIRE(smem <= 227 * 1024, "row too large for the shared-memory stage");
    CUR_CUDA_TRY(cudaFuncSetAttribute(her_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
    configured = smem;
  }
  const int64_t blocks = (a.batch + TILE - 1) / TILE;
  CUR_REQUIRE(blocks <= 0x7fffffff, "batch too large for one launch");
  her_sample_kernel<<<(unsigned)blocks, HER_THREADS, smem, (cudaStream_t)stream>>>(P);
  CUR_CHECK_LAUNCH();
  return CUR_OK;
}

// ---------------------------------------------------------------- known-answer access to the generator
__global__ void philox_kat_kernel(const uint32_t* __restrict__ in, int64_t n, uint32_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t* c = in + 6 * i;
  const cur::Philox x = cur::philox4x32_10(c[0], c[1], c[2], c[3], c[4], c[5]);
  for (int k = 0; k < 4; ++k) out[4 * i + k] = x.x[k];
}

extern "C" int cur_philox4x32_10(void* stream, const uint32_t* in, int64_t n, uint32_t* out) {
  CUR_REQUIRE(n >= 0 && (n == 0 || (in != nullptr && out != nullptr)), "bad arguments");
  if (n == 0) return CUR_OK;
  philox_kat_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(in, n, out);
  CUR_CHECK_LAUNCH();
  return CUR_OK;
}
