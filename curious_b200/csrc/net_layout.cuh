// Flat-parameter layout of one actor/critic net (GetFlat order) and the GEMM problem builders shared by
// the two DDPG schedules (ddpg.cu: dependency levels; ddpg_rows.cu: row clusters).
#pragma once
#include <string.h>

#include "mlp_kernels.cuh"

namespace cur {

struct NetLayout {
  int modular, in_s, in_g, H, L, out;
  int64_t off_W0, off_b0, off_W0g;
  int64_t off_W[CUR_MAX_LAYERS], off_b[CUR_MAX_LAYERS];   // hidden layers 1..L-1
  int64_t off_Wout, off_bout, total;
};

inline NetLayout net_layout(const cur_net_desc& d, int which /*0 Q, 1 pi*/) {
  NetLayout n;
  n.modular = d.modular;
  n.H = d.hidden;
  n.L = d.layers;
  const int act_in = (which == 0) ? d.dimu : 0;
  if (d.modular) {
    n.in_s = d.dimo + d.dimtd + act_in;
    n.in_g = d.dimg;
  } else {
    n.in_s = d.dimo + d.dimg + act_in;
    n.in_g = 0;
  }
  n.out = (which == 0) ? 1 : d.dimu;
  int64_t o = 0;
  n.off_W0 = o; o += (int64_t)n.in_s * n.H;
  n.off_b0 = o; o += n.H;
  n.off_W0g = o; o += (int64_t)n.in_g * n.H;
  for (int l = 1; l < n.L; ++l) {
    n.off_W[l] = o; o += (int64_t)n.H * n.H;
    n.off_b[l] = o; o += n.H;
  }
  n.off_Wout = o; o += (int64_t)n.H * n.out;
  n.off_bout = o; o += n.out;
  n.total = o;
  return n;
}

inline int64_t r4(int64_t x) { return (x + 3) & ~(int64_t)3; }

// ------------------------------------------------------------------------------------------------
// problem builders
// ------------------------------------------------------------------------------------------------
inline GemmProb zero_prob() {
  GemmProb p;
  memset(&p, 0, sizeof(p));
  p.scale2 = 1.f;
  return p;
}

// forward layer 0
inline GemmProb fwd0(const NetLayout& L, const float* th, const float* Xs, int ld_s, const float* Xg, int ld_g,
                     float* out, int64_t n) {
  GemmProb p = zero_prob();
  p.A = Xs; p.lda = ld_s; p.B = th + L.off_W0; p.ldb = L.H; p.K = L.in_s;
  if (L.in_g > 0) { p.A2 = Xg; p.lda2 = ld_g; p.B2 = th + L.off_W0g; p.ldb2 = L.H; p.K2 = L.in_g; }
  p.bias = th + L.off_b0;
  p.C = out; p.ldc = L.H; p.M = (int)n; p.N = L.H; p.epi = EPI_RELU;
  return p;
}
inline GemmProb fwdl(const NetLayout& L, const float* th, int l, const float* in, float* out, int64_t n) {
  GemmProb p = zero_prob();
  p.A = in; p.lda = L.H; p.B = th + L.off_W[l]; p.ldb = L.H; p.K = L.H;
  p.bias = th + L.off_b[l];
  p.C = out; p.ldc = L.H; p.M = (int)n; p.N = L.H; p.epi = EPI_RELU;
  return p;
}
inline GemmProb fwdout(const NetLayout& L, const float* th, const float* in, float* out, int ldc, int epi, int64_t n) {
  GemmProb p = zero_prob();
  p.A = in; p.lda = L.H; p.B = th + L.off_Wout; p.ldb = L.out; p.K = L.H;
  p.bias = th + L.off_bout;
  p.C = out; p.ldc = ldc; p.M = (int)n; p.N = L.out; p.epi = epi;
  return p;
}
// dX = dY * W^T, gated by the sign of the activation feeding W
inline GemmProb bwd_dx(const float* dY, int lddy, const float* W, int n_in, int n_out, const float* act, int ldact,
                       float* dX, int lddx, int64_t n) {
  GemmProb p = zero_prob();
  p.A = dY; p.lda = lddy; p.B = W; p.ldb = n_out; p.b_trans = 1; p.K = n_out;
  p.C = dX; p.ldc = lddx; p.M = (int)n; p.N = n_in;
  p.epi = act ? EPI_RELU_MASK : EPI_NONE; p.aux = act; p.ldaux = ldact;
  return p;
}
// dW = X^T * dY  -> n_in x n_out block of the flat gradient
inline GemmProb bwd_dw(const float* X, int ldx, int n_in, const float* dY, int lddy, int n_out, float* dW,
                       int64_t n) {
  GemmProb p = zero_prob();
  p.A = X; p.lda = ldx; p.a_trans = 1; p.K = (int)n;
  p.B = dY; p.ldb = lddy;
  p.C = dW; p.ldc = n_out; p.M = n_in; p.N = n_out;
  return p;
}
// db = 1^T * dY  -> the n_out floats right behind dW in the flat gradient ([W;b] is contiguous)
inline GemmProb bwd_db(const float* dY, int lddy, int n_out, float* db, int64_t n) {
  GemmProb p = zero_prob();
  p.ones_a = 1; p.a_trans = 1; p.K = (int)n;
  p.B = dY; p.ldb = lddy;
  p.C = db; p.ldc = n_out; p.M = 1; p.N = n_out;
  return p;
}

#define CUR_TRY(expr)            \
  do {                           \
    int _rc = (expr);            \
    if (_rc != CUR_OK) return _rc; \
  } while (0)

inline int check_desc(const cur_net_desc* d) {
  CUR_REQUIRE(d != nullptr, "net desc is NULL");
  CUR_REQUIRE(d->dimo > 0 && d->dimg > 0 && d->dimu > 0, "bad input dims");
  CUR_REQUIRE(d->hidden > 0 && (d->hidden % 4) == 0, "hidden must be a positive multiple of 4");
  CUR_REQUIRE(d->layers >= 1 && d->layers <= CUR_MAX_LAYERS, "layers out of range");
  CUR_REQUIRE(!d->modular || d->dimtd > 0, "modular net needs dimtd > 0");
  CUR_REQUIRE(d->max_u > 0.f, "max_u must be > 0");
  return CUR_OK;
}


}  // namespace cur
