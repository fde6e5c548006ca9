// L3-L6, engine MC_GEMM_SIMT_FP32: the contrastive soft-target loss with true-fp32 FMA GEMMs and
// materialised (b x B) strips.  This is the bring-up / cross-check engine: it restates
// /root/reference CLIP.py:34-43 and its autograd (closed form) with the
// same row partitioning as the tcgen05 engine, so both can be compared rank for rank.  It is NOT
// the performance path (it writes three b x B strips to HBM).
//
// Strips (row i local, global row gi = row_offset + i; column j global):
//   S  = T_loc I_all^T / tau           (logits rows owned by this rank)
//   St = I_loc T_all^T / tau           (St[i][j] = S[j][gi]: the transposed strip -> column stats)
//   Z  = (I_loc I_all^T + T_loc T_all^T) tau / 2
// The workspace keeps the strips from stats() to rowloss() to bwd(): call the phases in order
// with the same `ws`.
#include "clip_loss.cuh"

namespace mc {
namespace simt {

size_t workspace_bytes(int b, int B, int /*D*/) { return 3 * round_up((size_t)b * B * 4, 256); }

struct Strips {
  float *S, *St, *Z;
};
static Strips carve(void* ws, int b, int B) {
  size_t n = round_up((size_t)b * B * 4, 256);
  char* p = static_cast<char*>(ws);
  return {reinterpret_cast<float*>(p), reinterpret_cast<float*>(p + n),
          reinterpret_cast<float*>(p + 2 * n)};
}

// one warp per local row: three row log-sum-exps
__global__ void __launch_bounds__(256) row_lse3_kernel(const float* S, const float* St,
                                                       const float* Z, int b, int B, float* r,
                                                       float* c, float* rz, float* ps) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= b) return;
  Lse a, d, e;
  a.init(); d.init(); e.init();
  const size_t off = (size_t)row * B;
  for (int j = lane; j < B; j += 32) {
    a.add(S[off + j]);
    d.add(St[off + j]);
    e.add(Z[off + j]);
  }
  warp_merge_lse(a); warp_merge_lse(d); warp_merge_lse(e);
  const float rzv = e.value();
  float acc = 0.f;  // sum_j P_ij S_ij
  for (int j = lane; j < B; j += 32) acc = fmaf(__expf(Z[off + j] - rzv), S[off + j], acc);
  acc = warp_sum(acc);
  if (lane == 0) { r[row] = a.value(); c[row] = d.value(); rz[row] = rzv; ps[row] = acc; }
}

// one warp per local row: g_i = sum_j P_ij G_ij ; q_i = sum_k exp(Z_ik - rz_k) (Z symmetric)
__global__ void __launch_bounds__(256) rowloss_kernel(const float* S, const float* Z, int b, int B,
                                                      int row_offset, ClipStatsAll s, float* g_loc,
                                                      float* q_loc) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= b) return;
  const int gi = row_offset + row;
  const float ri = s.r[gi], rzi = s.rz[gi];
  const float inv2B = 0.5f / (float)B;
  float g = 0.f, q = 0.f;
  const size_t off = (size_t)row * B;
  for (int j = lane; j < B; j += 32) {
    float sv = S[off + j], zv = Z[off + j];
    float P = __expf(zv - rzi);
    float G = -(2.f * sv - ri - s.c[j]) * inv2B;
    g = fmaf(P, G, g);
    q += __expf(zv - s.rz[j]);
  }
  g = warp_sum(g);
  q = warp_sum(q);
  if (lane == 0) { g_loc[row] = g; q_loc[row] = q; }
}

__global__ void __launch_bounds__(1024) sum_kernel(const float* v, int n, float* out) {
  __shared__ double sm[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += (double)v[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double t = (threadIdx.x < (blockDim.x >> 5)) ? sm[threadIdx.x] : 0.0;
    t = warp_sum(t);
    if (threadIdx.x == 0) *out = (float)t;
  }
}

// elementwise: overwrite the strips with the three gradient-weight strips
//   S  <- MS  = gl * dS_ij / tau          (dT_loc += MS  I_all)
//   St <- MSt = gl * dS_ji / tau          (dI_loc += MSt T_all)
//   Z  <- MZ  = gl * tau/2 * (dZ_ij + dZ_ji)   (dI_loc += MZ I_all ; dT_loc += MZ T_all)
__global__ void __launch_bounds__(256) grad_weights_kernel(float* S, float* St, float* Z, int b,
                                                           int B, int row_offset, float tau,
                                                           ClipStatsAll s, const float* grad_loss) {
  const float gl = grad_loss ? *grad_loss : 1.f;
  const float inv2B = 0.5f / (float)B;
  const size_t total = (size_t)b * B;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int i = idx / B, j = idx % B, gi = row_offset + i;
    const float sv = S[idx], stv = St[idx], zv = Z[idx];
    const float ri = s.r[gi], ci = s.c[gi], rzi = s.rz[gi], gi_ = s.g[gi], qi = s.q[gi];
    const float rj = s.r[j], cj = s.c[j], rzj = s.rz[j], gj = s.g[j], qj = s.q[j];
    const float P = __expf(zv - rzi), Pt = __expf(zv - rzj);
    const float dS = (__expf(sv - ri) + __expf(sv - cj) * qj - 2.f * P) * inv2B;
    const float dSt = (__expf(stv - rj) + __expf(stv - ci) * qi - 2.f * Pt) * inv2B;
    const float G = -(2.f * sv - ri - cj) * inv2B;
    const float Gt = -(2.f * stv - rj - ci) * inv2B;
    const float dZs = P * (G - gi_) + Pt * (Gt - gj);
    S[idx] = gl * dS / tau;
    St[idx] = gl * dSt / tau;
    Z[idx] = gl * 0.5f * tau * dZs;
  }
}

int stats(const ClipProblem& p, float* r_loc, float* c_loc, float* rz_loc, float* ps_loc, void* ws,
          size_t ws_bytes, cudaStream_t st) {
  MC_REQUIRE(ws_bytes >= workspace_bytes(p.b, p.B, p.D), MC_ERR_WORKSPACE,
             "clip_stats(simt): workspace %zu < %zu", ws_bytes, workspace_bytes(p.b, p.B, p.D));
  Strips w = carve(ws, p.b, p.B);
  const int D = p.D;
  const float* I_loc = p.I_all + (size_t)p.row_offset * D;
  const float* T_loc = p.T_all + (size_t)p.row_offset * D;
  int rc;
  // S = T_loc I_all^T / tau
  SgemmArgs a{T_loc, D, 1, p.I_all, 1, D, w.S, p.B, p.b, p.B, D, 1.f / p.tau, nullptr, nullptr, 0};
  if ((rc = sgemm(a, st))) return rc;
  // St = I_loc T_all^T / tau
  SgemmArgs a2{I_loc, D, 1, p.T_all, 1, D, w.St, p.B, p.b, p.B, D, 1.f / p.tau, nullptr, nullptr, 0};
  if ((rc = sgemm(a2, st))) return rc;
  // Z = tau/2 (I_loc I_all^T + T_loc T_all^T)
  SgemmArgs a3{I_loc, D, 1, p.I_all, 1, D, w.Z, p.B, p.b, p.B, D, 0.5f * p.tau, nullptr, nullptr, 0};
  if ((rc = sgemm(a3, st))) return rc;
  SgemmArgs a4{T_loc, D, 1, p.T_all, 1, D, w.Z, p.B, p.b, p.B, D, 0.5f * p.tau, nullptr, nullptr, 1};
  if ((rc = sgemm(a4, st))) return rc;
  row_lse3_kernel<<<(p.b + 7) / 8, 256, 0, st>>>(w.S, w.St, w.Z, p.b, p.B, r_loc, c_loc, rz_loc, ps_loc);
  MC_LAUNCH_CHECK();
  return MC_OK;
}

int rowloss(const ClipProblem& p, const ClipStatsAll& s, const float* /*ps_loc: this engine keeps S*/,
            float* g_loc, float* q_loc, float* loss_part, void* ws, size_t ws_bytes, cudaStream_t st) {
  MC_REQUIRE(ws_bytes >= workspace_bytes(p.b, p.B, p.D), MC_ERR_WORKSPACE,
             "clip_rowloss(simt): workspace too small");
  Strips w = carve(ws, p.b, p.B);
  rowloss_kernel<<<(p.b + 7) / 8, 256, 0, st>>>(w.S, w.Z, p.b, p.B, p.row_offset, s, g_loc, q_loc);
  MC_LAUNCH_CHECK();
  sum_kernel<<<1, 1024, 0, st>>>(g_loc, p.b, loss_part);
  MC_LAUNCH_CHECK();
  return MC_OK;
}

int bwd(const ClipProblem& p, const ClipStatsAll& s, const float* grad_loss, float* dI, float* dT,
        void* ws, size_t ws_bytes, cudaStream_t st) {
  MC_REQUIRE(ws_bytes >= workspace_bytes(p.b, p.B, p.D), MC_ERR_WORKSPACE,
             "clip_bwd(simt): workspace too small");
  Strips w = carve(ws, p.b, p.B);
  const int D = p.D;
  size_t total = (size_t)p.b * p.B;
  int blocks = (int)((total + 255) / 256);
  int cap = num_sms() * 16;
  if (blocks > cap) blocks = cap;
  grad_weights_kernel<<<blocks, 256, 0, st>>>(w.S, w.St, w.Z, p.b, p.B, p.row_offset, p.tau, s,
                                              grad_loss);
  MC_LAUNCH_CHECK();
  int rc;
  // dT_loc = MS I_all + MZ T_all
  SgemmArgs a{w.S, p.B, 1, p.I_all, D, 1, dT, D, p.b, D, p.B, 1.f, nullptr, nullptr, 0};
  if ((rc = sgemm(a, st))) return rc;
  SgemmArgs a2{w.Z, p.B, 1, p.T_all, D, 1, dT, D, p.b, D, p.B, 1.f, nullptr, nullptr, 1};
  if ((rc = sgemm(a2, st))) return rc;
  // dI_loc = MSt T_all + MZ I_all
  SgemmArgs a3{w.St, p.B, 1, p.T_all, D, 1, dI, D, p.b, D, p.B, 1.f, nullptr, nullptr, 0};
  if ((rc = sgemm(a3, st))) return rc;
  SgemmArgs a4{w.Z, p.B, 1, p.I_all, D, 1, dI, D, p.b, D, p.B, 1.f, nullptr, nullptr, 1};
  if ((rc = sgemm(a4, st))) return rc;
  return MC_OK;
}

}  // namespace simt
}  // namespace mc
