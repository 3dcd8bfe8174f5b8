// Replay storage + fused HER relabel / gather / reward / preprocess kernel (sm_100a).
//
// Replaces (reference flowersteam/curious, paths relative to its root):
//   baselines/her/replay_buffer.py:57-72      store_episode           -> cur_store_episodes
//   baselines/her/her.py:20-66, 99-183        _sample_her_transitions -> cur_her_sample
//   baselines/her/ddpg.py:325-345, 350-353    per-buffer loop, concat, shuffle, _preprocess_og
//   config.py:158-159 reward_fun              restated module-distance reward
//
// Kernel design (HBM bound, gather of ~0.5 KB per transition):
//   * one CTA handles a tile of TILE consecutive OUTPUT rows;
//   * the per-row draws are decoded by TILE threads (injected stream or Philox4x32-10);
//   * each of those threads issues cp.async.bulk (TMA engine, SASS UBLKCP) copies of
//       row t [+ head of row t+1]  and, for HER rows, the future achieved goal
//     into shared memory, completion tracked by one mbarrier with expect_tx byte counts - no
//     register staging, ~20 KB in flight per CTA, several CTAs per SM;
//   * relabel (g, task_descr), float64 reward and clip are done out of shared memory and every
//     output array is written as contiguous, fully coalesced 16-byte stores.
#include "common.cuh"

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

struct HerPlan {
  // shared-memory image of one transition: [row t | head of row t+1] at the SAME float offsets as in
  // global memory, followed by the future achieved goal.
  int img_floats;   // row_stride + next_prefix
  int fut_off;      // float offset of the future-ag copy inside the per-transition stage
  int stage_stride; // floats per transition in shared memory
  int c1_off, c1_len;  // first bulk copy  (floats, relative to row t)
  int c2_off, c2_len;  // second bulk copy (0 length if merged into the first)
  int fut_len;         // floats of the future-ag copy (dimag padded to 4)
  int dimg_pad;
};

struct HerKernelParams {
  cur_her_args a;
  HerPlan p;
};

// ------------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier + bulk async copy (global -> shared::cta of this CTA)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
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

__global__ void __launch_bounds__(HER_THREADS)
her_sample_kernel(const __grid_constant__ HerKernelParams P) {
  const cur_her_args& a = P.a;
  const cur_layout& L = a.L;
  const HerPlan& pl = P.p;
  extern __shared__ __align__(128) unsigned char smem_raw[];

  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);               // 16 bytes reserved
  float* stage = reinterpret_cast<float*>(smem_raw + 16);             // TILE * stage_stride
  float* gfin = stage + TILE * pl.stage_stride;                        // TILE * dimg_pad
  float* rew = gfin + TILE * pl.dimg_pad;                              // TILE
  int* m_her = reinterpret_cast<int*>(rew + TILE);                     // TILE: 1 if HER row
  int* m_task = m_her + TILE;                                          // TILE: module written to td (-1: keep)
  int* m_relab = m_task + TILE;                                        // TILE: module whose slice is relabelled
  int16_t* gmap = reinterpret_cast<int16_t*>(m_relab + TILE);          // n_maps * dimg_pad
  const int n_maps = (a.mode == CUR_MODE_FLAT) ? 1 : a.tasks.n_tasks;

  const int tid = threadIdx.x;
  const int64_t j0 = (int64_t)blockIdx.x * TILE;
  const int nrows = (int)min((int64_t)TILE, a.batch - j0);

  // goal-column -> achieved-goal-column map per module (her.py:145-155); -1 = not in the slice
  for (int i = tid; i < n_maps * pl.dimg_pad; i += HER_THREADS) gmap[i] = -1;
  if (tid == 0) {
    mbar_init(bar, TILE);
    fence_mbar_init();
  }
  __syncthreads();
  if (a.mode == CUR_MODE_FLAT) {
    for (int m = tid; m < a.tasks.n_tasks; m += HER_THREADS)
      for (int k = 0; k < a.tasks.len[m]; ++k) gmap[a.tasks.g_idx[m][k]] = a.tasks.ag_idx[m][k];
  } else {
    for (int m = tid; m < a.tasks.n_tasks; m += HER_THREADS)
      for (int k = 0; k < a.tasks.len[m]; ++k)
        gmap[m * pl.dimg_pad + a.tasks.g_idx[m][k]] = a.tasks.ag_idx[m][k];
  }

  // ---------------------------------------------------------------- phase A: draws + bulk copies
  int my_ft = -1, my_choice = -1, my_ep = 0, my_t = 0, my_ttr = -1;
  if (tid < TILE) {
    if (tid < nrows) {
      const int64_t j = j0 + tid;
      const int64_t c = a.perm ? (int64_t)a.perm[j] : j;
      // segment lookup: concat rows are the segments' counts laid end to end (ddpg.py:326-345)
      int s = 0;
      int64_t acc = 0;
      while (s + 1 < a.n_segments && c >= acc + a.seg[s].count) {
        acc += a.seg[s].count;
        ++s;
      }
      const float* base = a.seg[s].base;
      const int E = a.seg[s].n_episodes;
      my_ttr = a.seg[s].task_to_replay;
      double u_her, u_off;
      if (a.inj_ep != nullptr) {
        my_ep = a.inj_ep[c];
        my_t = a.inj_t[c];
        u_her = a.inj_u_her[c];
        u_off = a.inj_u_off[c];
        if (a.inj_choice) my_choice = a.inj_choice[c];
      } else {
        Philox x = philox4x32_10((uint32_t)c, (uint32_t)(c >> 32), (uint32_t)a.call_offset,
                                 (uint32_t)(a.call_offset >> 32), (uint32_t)a.seed,
                                 (uint32_t)(a.seed >> 32));
        my_ep = (int)mulhi32(x.x[0], (uint32_t)E);
        my_t = (int)mulhi32(x.x[1], (uint32_t)L.T);
        u_her = u01_from_u32(x.x[2]);
        u_off = u01_from_u32(x.x[3]);
        if (a.mode == CUR_MODE_RANDOM_TASK || a.mode == CUR_MODE_CP_TASK) {
          Philox y = philox4x32_10((uint32_t)c, (uint32_t)(c >> 32), (uint32_t)a.call_offset,
                                   (uint32_t)(a.call_offset >> 32) ^ 0x80000000u, (uint32_t)a.seed,
                                   (uint32_t)(a.seed >> 32));
          if (a.mode == CUR_MODE_RANDOM_TASK) {
            my_choice = (int)mulhi32(y.x[0], (uint32_t)a.tasks.n_tasks);
          } else {
            // np.random.choice(p=): cdf.searchsorted(u, side='right')
            double u = u01_from_u32(y.x[0]);
            int k = 0;
            while (k < a.tasks.n_tasks - 1 && a.tasks.cdf[k] <= u) ++k;
            my_choice = k;
          }
        }
      }
      const bool her = u_her < a.future_p;                         // her.py:115
      if (her) my_ft = my_t + 1 + (int)(u_off * (double)(L.T - my_t));   // her.py:116-118
      m_her[tid] = her ? 1 : 0;

      const float* row = base + ((int64_t)my_ep * (L.T + 1) + my_t) * (int64_t)L.row_stride;
      float* st = stage + tid * pl.stage_stride;
      uint32_t bytes = (uint32_t)(pl.c1_len + pl.c2_len + (her ? pl.fut_len : 0)) * 4u;
      mbar_arrive_expect_tx(bar, bytes);
      bulk_g2s(st + pl.c1_off, row + pl.c1_off, (uint32_t)pl.c1_len * 4u, bar);
      if (pl.c2_len > 0) bulk_g2s(st + pl.c2_off, row + pl.c2_off, (uint32_t)pl.c2_len * 4u, bar);
      if (her) {
        const float* frow = base + ((int64_t)my_ep * (L.T + 1) + my_ft) * (int64_t)L.row_stride;
        bulk_g2s(st + pl.fut_off, frow + L.off_ag, (uint32_t)pl.fut_len * 4u, bar);
      }
    } else {
      m_her[tid] = 0;
      mbar_arrive(bar);
    }
  }
  __syncthreads();   // gmap + m_her visible; (copies still in flight)
  mbar_wait(bar, 0);

  // ---------------------------------------------------------------- phase A': module decisions
  if (tid < nrows) {
    const float* st = stage + tid * pl.stage_stride;
    int own = -1;
    for (int k = 0; k < L.dimtd; ++k)
      if (st[L.off_td + k] == 1.0f) { own = k; break; }            // argwhere(td == 1) (her.py:134)
    int relab = -1, newtd = -1;
    if (m_her[tid]) {
      switch (a.mode) {
        case CUR_MODE_BUFFER: relab = (my_ttr >= 0) ? my_ttr : own; newtd = relab; break;
        case CUR_MODE_RANDOM_TASK:
        case CUR_MODE_CP_TASK: relab = my_choice; newtd = relab; break;
        case CUR_MODE_CURRENT_TASK: relab = own; newtd = -1; break;
        default: relab = 0; newtd = -1; break;   // FLAT: single map
      }
    }
    m_relab[tid] = relab;
    m_task[tid] = newtd;
    if (a.idx_out) {
      int32_t* io = a.idx_out + (j0 + tid) * 4;
      io[0] = my_ep; io[1] = my_t; io[2] = my_ft; io[3] = m_her[tid] ? relab : -1;
    }
    // module used by the reward: relabelled module for HER rows that rewrite td, else the row's own
    my_choice = (newtd >= 0) ? newtd : own;
  }
  __syncthreads();

  // ---------------------------------------------------------------- phase B: relabelled goal
  {
    const int n = nrows * pl.dimg_pad;
    const float inv = 1.0f / (float)pl.dimg_pad;
    const bool wipe = (a.mode == CUR_MODE_BUFFER || a.mode == CUR_MODE_RANDOM_TASK ||
                       a.mode == CUR_MODE_CP_TASK);
    for (int i = tid; i < n; i += HER_THREADS) {
      int tr = __float2int_rz(((float)i + 0.5f) * inv);
      int k = i - tr * pl.dimg_pad;
      const float* st = stage + tr * pl.stage_stride;
      float v = (k < L.dimg) ? st[L.off_g + k] : 0.0f;
      if (m_her[tr] && k < L.dimg) {
        int relab = m_relab[tr];
        int src = (relab >= 0) ? gmap[relab * pl.dimg_pad + k] : -1;
        if (src >= 0) v = st[pl.fut_off + src];          // her.py:154 / 163 / 47
        else if (wipe) v = 0.0f;                          // her.py:151
      }
      gfin[i] = v;
    }
  }
  __syncthreads();

  // ---------------------------------------------------------------- phase C: reward (float64)
  if (tid < nrows && a.r != nullptr) {
    const float* st = stage + tid * pl.stage_stride;
    const float* ag2 = st + L.row_stride + L.off_ag;
    const float* gf = gfin + tid * pl.dimg_pad;
    double d2 = 0.0;
    if (a.mode == CUR_MODE_FLAT) {
      for (int m = 0; m < a.tasks.n_tasks; ++m)
        for (int k = 0; k < a.tasks.len[m]; ++k) {
          double diff = (double)ag2[a.tasks.ag_idx[m][k]] - (double)gf[a.tasks.g_idx[m][k]];
          d2 = __dadd_rn(d2, __dmul_rn(diff, diff));
        }
    } else {
      int m = my_choice;
      if (m >= 0)
        for (int k = 0; k < a.tasks.len[m]; ++k) {
          double diff = (double)ag2[a.tasks.ag_idx[m][k]] - (double)gf[a.tasks.g_idx[m][k]];
          d2 = __dadd_rn(d2, __dmul_rn(diff, diff));
        }
    }
    rew[tid] = (sqrt(d2) > a.tasks.threshold) ? -1.0f : 0.0f;
  }
  __syncthreads();

  // ---------------------------------------------------------------- phase D: coalesced outputs
  const float c = a.clip_obs;
  const bool do_clip = c > 0.0f;
  const int ss = pl.stage_stride;
  auto ld4 = [&](int tr, int off) {
    return *reinterpret_cast<const float4*>(stage + tr * ss + off);
  };
  auto clip4 = [&](float4 v) {
    if (do_clip) { v.x = clipf(v.x, c); v.y = clipf(v.y, c); v.z = clipf(v.z, c); v.w = clipf(v.w, c); }
    return v;
  };
  auto clip1 = [&](float v) { return do_clip ? clipf(v, c) : v; };

  emit(a.o, L.dimo, j0, nrows,
       [&](int tr, int k) { return clip4(ld4(tr, L.off_o + k)); },
       [&](int tr, int k) { return clip1(stage[tr * ss + L.off_o + k]); });
  emit(a.o_2, L.dimo, j0, nrows,
       [&](int tr, int k) { return clip4(ld4(tr, L.row_stride + L.off_o + k)); },
       [&](int tr, int k) { return clip1(stage[tr * ss + L.row_stride + L.off_o + k]); });
  emit(a.u, L.dimu, j0, nrows,
       [&](int tr, int k) { return ld4(tr, L.off_u + k); },
       [&](int tr, int k) { return stage[tr * ss + L.off_u + k]; });
  emit(a.ag, L.dimag, j0, nrows,
       [&](int tr, int k) { return ld4(tr, L.off_ag + k); },
       [&](int tr, int k) { return stage[tr * ss + L.off_ag + k]; });
  emit(a.ag_2, L.dimag, j0, nrows,
       [&](int tr, int k) { return ld4(tr, L.row_stride + L.off_ag + k); },
       [&](int tr, int k) { return stage[tr * ss + L.row_stride + L.off_ag + k]; });
  emit(a.change, L.dimchange, j0, nrows,
       [&](int tr, int k) { return ld4(tr, L.off_change + k); },
       [&](int tr, int k) { return stage[tr * ss + L.off_change + k]; });
  emit(a.info, L.diminfo, j0, nrows,
       [&](int tr, int k) { return ld4(tr, L.off_info + k); },
       [&](int tr, int k) { return stage[tr * ss + L.off_info + k]; });
  // task_descr: one-hot of the replayed module on HER rows that rewrite it (her.py:152,155)
  auto td1 = [&](int tr, int k) {
    int nt = m_task[tr];
    return (nt >= 0) ? ((k == nt) ? 1.0f : 0.0f) : stage[tr * ss + L.off_td + k];
  };
  emit(a.td, L.dimtd, j0, nrows,
       [&](int tr, int k) { return make_float4(td1(tr, k), td1(tr, k + 1), td1(tr, k + 2), td1(tr, k + 3)); },
       td1);
  // g / g_2 (ddpg.py:350-353): optional relative goals, then clip
  auto g1 = [&](int tr, int k) {
    float v = gfin[tr * pl.dimg_pad + k];
    if (a.relative_goals) v = v - stage[tr * ss + L.off_ag + k];
    return clip1(v);
  };
  auto g2_1 = [&](int tr, int k) {
    float v = gfin[tr * pl.dimg_pad + k];
    if (a.relative_goals) v = v - stage[tr * ss + L.row_stride + L.off_ag + k];
    return clip1(v);
  };
  emit(a.g, L.dimg, j0, nrows,
       [&](int tr, int k) { return make_float4(g1(tr, k), g1(tr, k + 1), g1(tr, k + 2), g1(tr, k + 3)); }, g1);
  emit(a.g_2, L.dimg, j0, nrows,
       [&](int tr, int k) { return make_float4(g2_1(tr, k), g2_1(tr, k + 1), g2_1(tr, k + 2), g2_1(tr, k + 3)); },
       g2_1);
  if (a.r != nullptr && tid < nrows) a.r[j0 + tid] = rew[tid];
}

// ------------------------------------------------------------------------------------------------
// store: pack key-major episodes into rows, one CTA per (row, copy)
// ------------------------------------------------------------------------------------------------
struct StoreParams {
  cur_layout L;
  cur_episode_src src;
  int n_copies;
  int32_t copy_src[CUR_MAX_COPIES];
  float* copy_base[CUR_MAX_COPIES];
  int64_t copy_slot[CUR_MAX_COPIES];
};

__global__ void __launch_bounds__(128) store_episodes_kernel(const __grid_constant__ StoreParams S) {
  const cur_layout& L = S.L;
  const int t = blockIdx.x;        // 0..T
  const int cpy = blockIdx.y;
  const int e = S.copy_src[cpy];
  float* dst = S.copy_base[cpy] + (S.copy_slot[cpy] * (L.T + 1) + t) * (int64_t)L.row_stride;
  const bool last = (t == L.T);
  for (int k = threadIdx.x; k < L.row_stride; k += blockDim.x) {
    float v = 0.0f;
    if (k < L.off_o) {
      int j = k - L.off_ag;
      if (j < L.dimag) v = S.src.ag[((int64_t)e * (L.T + 1) + t) * L.dimag + j];
    } else if (k < L.off_g) {
      int j = k - L.off_o;
      if (j < L.dimo) v = S.src.o[((int64_t)e * (L.T + 1) + t) * L.dimo + j];
    } else if (!last) {
      const int64_t rt = (int64_t)e * L.T + t;
      if (k < L.off_u) {
        int j = k - L.off_g;
        if (j < L.dimg) v = S.src.g[rt * L.dimg + j];
      } else if (k < L.off_td) {
        int j = k - L.off_u;
        if (j < L.dimu) v = S.src.u[rt * L.dimu + j];
      } else if (k < L.off_change) {
        int j = k - L.off_td;
        if (j < L.dimtd && S.src.td) v = S.src.td[rt * L.dimtd + j];
      } else if (k < L.off_info) {
        int j = k - L.off_change;
        if (j < L.dimchange && S.src.change) v = S.src.change[rt * L.dimchange + j];
      } else {
        int j = k - L.off_info;
        if (j < L.diminfo && S.src.info) v = S.src.info[rt * L.diminfo + j];
      }
    }
    dst[k] = v;
  }
}

static int make_plan(const cur_her_args& a, HerPlan* p) {
  const cur_layout& L = a.L;
  p->img_floats = L.row_stride + L.next_prefix;
  p->fut_off = p->img_floats;
  p->fut_len = round_up4(L.dimag);
  p->stage_stride = p->img_floats + p->fut_len;
  p->dimg_pad = round_up4(L.dimg);
  const bool need_ag_t = (a.ag != nullptr) || a.relative_goals;
  const bool need_tail = (a.change != nullptr && L.dimchange > 0) || (a.info != nullptr && L.diminfo > 0);
  const int start = need_ag_t ? 0 : L.off_o;
  if (need_tail) {
    p->c1_off = start;
    p->c1_len = p->img_floats - start;
    p->c2_off = 0;
    p->c2_len = 0;
  } else {
    p->c1_off = start;
    p->c1_len = L.off_change - start;
    p->c2_off = L.row_stride;
    p->c2_len = L.next_prefix;
    if (p->c1_off + p->c1_len == p->c2_off) {  // nothing to skip: merge
      p->c1_len += p->c2_len;
      p->c2_len = 0;
    }
  }
  return CUR_OK;
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
  L->off_ag = off; off += round_up4(dimag);
  L->off_o = off; off += round_up4(dimo);
  L->next_prefix = off;
  L->off_g = off; off += round_up4(dimg);
  L->off_u = off; off += round_up4(dimu);
  L->off_td = off; off += round_up4(dimtd);
  L->off_change = off; off += round_up4(dimchange);
  L->off_info = off; off += round_up4(diminfo);
  L->row_stride = off;
  return CUR_OK;
}

extern "C" int cur_store_episodes(void* stream, const cur_layout* L, const cur_episode_src* src, int n_ep,
                                  int n_copies, const int32_t* copy_src, float* const* copy_base,
                                  const int64_t* copy_slot) {
  CUR_REQUIRE(L && src && copy_src && copy_base && copy_slot, "NULL argument");
  CUR_REQUIRE(src->o && src->ag && src->g && src->u, "o/ag/g/u sources are required");
  CUR_REQUIRE(n_copies >= 0 && n_ep > 0, "bad counts");
  cudaStream_t s = (cudaStream_t)stream;
  for (int done = 0; done < n_copies; done += CUR_MAX_COPIES) {
    StoreParams S;
    S.L = *L;
    S.src = *src;
    S.n_copies = (n_copies - done < CUR_MAX_COPIES) ? n_copies - done : CUR_MAX_COPIES;
    for (int i = 0; i < S.n_copies; ++i) {
      CUR_REQUIRE(copy_src[done + i] >= 0 && copy_src[done + i] < n_ep, "copy_src out of range");
      CUR_REQUIRE(copy_base[done + i] != nullptr && copy_slot[done + i] >= 0, "bad destination");
      S.copy_src[i] = copy_src[done + i];
      S.copy_base[i] = copy_base[done + i];
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
    if (a.seg[i].count > 0) {
      CUR_REQUIRE(a.seg[i].base != nullptr, "segment base is NULL");
      CUR_REQUIRE(a.seg[i].n_episodes > 0, "sampling from an empty buffer (replay_buffer.py:43)");
      CUR_REQUIRE(a.seg[i].task_to_replay < a.tasks.n_tasks, "task_to_replay out of range");
    }
    total += a.seg[i].count;
  }
  CUR_REQUIRE(total == a.batch, "segment counts must sum to batch (ddpg.py:323)");
  for (int m = 0; m < a.tasks.n_tasks; ++m) {
    CUR_REQUIRE(a.tasks.len[m] >= 0 && a.tasks.len[m] <= CUR_MAX_SLICE, "module slice too long");
    for (int k = 0; k < a.tasks.len[m]; ++k) {
      CUR_REQUIRE(a.tasks.g_idx[m][k] >= 0 && a.tasks.g_idx[m][k] < a.L.dimg, "g index out of range");
      CUR_REQUIRE(a.tasks.ag_idx[m][k] >= 0 && a.tasks.ag_idx[m][k] < a.L.dimag, "ag index out of range");
    }
  }
  if (a.inj_ep != nullptr)
    CUR_REQUIRE(a.inj_t && a.inj_u_her && a.inj_u_off, "incomplete injected stream");
  if (a.relative_goals) CUR_REQUIRE(a.L.dimg == a.L.dimag, "relative goals need dimg == dimag");

  HerKernelParams P;
  P.a = a;
  make_plan(a, &P.p);
  const int n_maps = (a.mode == CUR_MODE_FLAT) ? 1 : a.tasks.n_tasks;
  size_t smem = (size_t)TILE * P.p.stage_stride * 4 + (size_t)TILE * P.p.dimg_pad * 4 + TILE * 4 +
                3 * TILE * 4 + (size_t)n_maps * P.p.dimg_pad * 2 + 16 + 16;
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
