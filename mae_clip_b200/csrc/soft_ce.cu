// L5: standalone soft-label cross entropy on materialised (rows, cols) inputs.
// Replaces /root/reference CLIP.py:46-52 (LogSoftmax + multiply + row sum) and its autograd.
// HBM-bound: forward reads preds and targets exactly once (online log-sum-exp), backward reads
// them once more and writes the two gradients.  Element strides are honoured so the `.T` views
// passed at CLIP.py:41 run coalesced without the transposed copies eager PyTorch makes.
#include "common.cuh"

namespace mc {

struct Strided2D {
  const float* p;
  int64_t rs, cs;
  __device__ __forceinline__ float at(int r, int c) const { return p[r * rs + c * cs]; }
};

// ---- forward, reduction dim contiguous: one warp per row -------------------------------------
__global__ void __launch_bounds__(256) soft_ce_fwd_rowmajor(Strided2D preds, Strided2D tg, int rows,
                                                           int cols, float* loss_rows,
                                                           float* row_lse, float* row_tsum,
                                                           int vec_ok) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* pr = preds.p + (int64_t)warp * preds.rs;
  const float* tr = tg.p + (int64_t)warp * tg.rs;
  Lse l;
  l.init();
  float tsum = 0.f, tp = 0.f;
  int c = 0;
  if (vec_ok) {
    const int nv = cols >> 2;
    for (int v = lane; v < nv; v += 32) {
      float4 a = ld_stream(reinterpret_cast<const float4*>(pr) + v);
      float4 b = ld_stream(reinterpret_cast<const float4*>(tr) + v);
      float m4 = fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w));
      if (m4 > l.m) { l.s *= __expf(l.m - m4); l.m = m4; }
      l.s += __expf(a.x - l.m) + __expf(a.y - l.m) + __expf(a.z - l.m) + __expf(a.w - l.m);
      tsum += (b.x + b.y) + (b.z + b.w);
      tp = fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, tp))));
    }
    c = nv << 2;
  }
  for (int j = c + lane; j < cols; j += 32) {
    float a = pr[j], b = tr[j];
    l.add(a);
    tsum += b;
    tp = fmaf(a, b, tp);
  }
  warp_merge_lse(l);
  tsum = warp_sum(tsum);
  tp = warp_sum(tp);
  if (lane == 0) {
    float lse = l.value();
    loss_rows[warp] = lse * tsum - tp;
    if (row_lse) row_lse[warp] = lse;
    if (row_tsum) row_tsum[warp] = tsum;
  }
}

// ---- forward, generic strides (coalesced when the ROW stride is 1: the transposed view) -------
// block = 32 rows (x) x 32 column groups (y); partials merged through shared memory.
__global__ void __launch_bounds__(1024) soft_ce_fwd_strided(Strided2D preds, Strided2D tg, int rows,
                                                           int cols, float* loss_rows,
                                                           float* row_lse, float* row_tsum) {
  __shared__ float sm_m[32][33], sm_s[32][33], sm_t[32][33], sm_tp[32][33];
  const int r = blockIdx.x * 32 + threadIdx.x;
  Lse l;
  l.init();
  float tsum = 0.f, tp = 0.f;
  if (r < rows) {
    for (int c = threadIdx.y; c < cols; c += 32) {
      float a = preds.at(r, c), b = tg.at(r, c);
      l.add(a);
      tsum += b;
      tp = fmaf(a, b, tp);
    }
  }
  sm_m[threadIdx.y][threadIdx.x] = l.m;
  sm_s[threadIdx.y][threadIdx.x] = l.s;
  sm_t[threadIdx.y][threadIdx.x] = tsum;
  sm_tp[threadIdx.y][threadIdx.x] = tp;
  __syncthreads();
  // warp y=k reduces row k of the block: lanes read column k of the partial arrays
  const int rr = threadIdx.y, g = threadIdx.x;
  Lse t;
  t.m = sm_m[g][rr];
  t.s = sm_s[g][rr];
  float ts = sm_t[g][rr], tps = sm_tp[g][rr];
  warp_merge_lse(t);
  ts = warp_sum(ts);
  tps = warp_sum(tps);
  const int row = blockIdx.x * 32 + rr;
  if (g == 0 && row < rows) {
    float lse = t.value();
    loss_rows[row] = lse * ts - tps;
    if (row_lse) row_lse[row] = lse;
    if (row_tsum) row_tsum[row] = ts;
  }
}

// ---- backward: pure elementwise given the saved row statistics ----------------------------------
// fast_is_col: 1 -> threads run along columns (col stride 1), 0 -> along rows (row stride 1)
__global__ void __launch_bounds__(256) soft_ce_bwd_kernel(Strided2D preds, Strided2D tg, int rows,
                                                         int cols, const float* row_lse,
                                                         const float* row_tsum, const float* grad,
                                                         float* dp, int64_t dp_rs, int64_t dp_cs,
                                                         float* dt, int64_t dt_rs, int64_t dt_cs,
                                                         int fast_is_col) {
  const int64_t total = (int64_t)rows * cols;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int r, c;
    if (fast_is_col) { r = idx / cols; c = idx % cols; } else { c = idx / rows; r = idx % rows; }
    float g = grad[r];
    float lp = preds.at(r, c) - row_lse[r];
    if (dp) dp[r * dp_rs + c * dp_cs] = g * (__expf(lp) * row_tsum[r] - tg.at(r, c));
    if (dt) dt[r * dt_rs + c * dt_cs] = -g * lp;
  }
}

// ---- backward, everything contiguous along the reduction dim: one warp per row, 16-byte vectors ----
__global__ void __launch_bounds__(256) soft_ce_bwd_rowmajor(const float* __restrict__ preds, int64_t p_rs,
                                                           const float* __restrict__ tg, int64_t t_rs, int rows,
                                                           int cols, const float* __restrict__ row_lse,
                                                           const float* __restrict__ row_tsum,
                                                           const float* __restrict__ grad, float* __restrict__ dp,
                                                           int64_t dp_rs, float* __restrict__ dt, int64_t dt_rs) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int nv = cols >> 2;
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
    const float g = grad[r], lse = row_lse[r], ts = row_tsum[r];
    const float4* pr = reinterpret_cast<const float4*>(preds + (int64_t)r * p_rs);
    const float4* tr = reinterpret_cast<const float4*>(tg + (int64_t)r * t_rs);
    float4* dpr = dp ? reinterpret_cast<float4*>(dp + (int64_t)r * dp_rs) : nullptr;
    float4* dtr = dt ? reinterpret_cast<float4*>(dt + (int64_t)r * dt_rs) : nullptr;
    for (int v = lane; v < nv; v += 32) {
      const float4 a = ld_stream(pr + v);
      const float lx = a.x - lse, ly = a.y - lse, lz = a.z - lse, lw = a.w - lse;
      if (dpr) {
        const float4 b = ld_stream(tr + v);
        st_stream(dpr + v, make_float4(g * (__expf(lx) * ts - b.x), g * (__expf(ly) * ts - b.y),
                                       g * (__expf(lz) * ts - b.z), g * (__expf(lw) * ts - b.w)));
      }
      if (dtr) st_stream(dtr + v, make_float4(-g * lx, -g * ly, -g * lz, -g * lw));
    }
  }
}

}  // namespace mc

extern "C" {

int mc_soft_ce_fwd(const float* preds, int64_t p_rs, int64_t p_cs, const float* targets,
                   int64_t t_rs, int64_t t_cs, int rows, int cols, float* loss_rows,
                   float* row_lse, float* row_tsum, void* stream) {
  MC_ARCH_GUARD();
  MC_REQUIRE(preds && targets && loss_rows, MC_ERR_BAD_ARG, "soft_ce_fwd: null pointer");
  MC_REQUIRE(rows >= 0 && cols > 0, MC_ERR_BAD_ARG, "soft_ce_fwd: bad shape (%d, %d)", rows, cols);
  if (rows == 0) return MC_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  mc::Strided2D P{preds, p_rs, p_cs}, T{targets, t_rs, t_cs};
  if (p_cs == 1 && t_cs == 1) {
    int vec_ok = mc::aligned(preds, 16) && mc::aligned(targets, 16) && (p_rs % 4 == 0) &&
                 (t_rs % 4 == 0);
    int blocks = (rows + 7) / 8;
    mc::soft_ce_fwd_rowmajor<<<blocks, 256, 0, st>>>(P, T, rows, cols, loss_rows, row_lse, row_tsum,
                                                     vec_ok);
  } else {
    dim3 block(32, 32);
    mc::soft_ce_fwd_strided<<<(rows + 31) / 32, block, 0, st>>>(P, T, rows, cols, loss_rows,
                                                                row_lse, row_tsum);
  }
  MC_LAUNCH_CHECK();
  return MC_OK;
}

int mc_soft_ce_bwd(const float* preds, int64_t p_rs, int64_t p_cs, const float* targets,
                   int64_t t_rs, int64_t t_cs, int rows, int cols, const float* row_lse,
                   const float* row_tsum, const float* grad_rows, float* dpreds, int64_t dp_rs,
                   int64_t dp_cs, float* dtargets, int64_t dt_rs, int64_t dt_cs, void* stream) {
  MC_ARCH_GUARD();
  MC_REQUIRE(preds && targets && row_lse && row_tsum && grad_rows, MC_ERR_BAD_ARG,
             "soft_ce_bwd: null pointer");
  MC_REQUIRE(rows >= 0 && cols > 0, MC_ERR_BAD_ARG, "soft_ce_bwd: bad shape (%d, %d)", rows, cols);
  if (rows == 0 || (!dpreds && !dtargets)) return MC_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  mc::Strided2D P{preds, p_rs, p_cs}, T{targets, t_rs, t_cs};
  const bool vec = p_cs == 1 && t_cs == 1 && (!dpreds || dp_cs == 1) && (!dtargets || dt_cs == 1) && cols % 4 == 0 &&
                   p_rs % 4 == 0 && t_rs % 4 == 0 && (!dpreds || dp_rs % 4 == 0) && (!dtargets || dt_rs % 4 == 0) &&
                   mc::aligned(preds, 16) && mc::aligned(targets, 16) && (!dpreds || mc::aligned(dpreds, 16)) &&
                   (!dtargets || mc::aligned(dtargets, 16));
  if (vec) {
    int nb = (rows + 7) / 8;
    const int capv = mc::num_sms() * 8;
    if (nb > capv) nb = capv;
    mc::soft_ce_bwd_rowmajor<<<nb, 256, 0, st>>>(preds, p_rs, targets, t_rs, rows, cols, row_lse, row_tsum, grad_rows,
                                                 dpreds, dp_rs, dtargets, dt_rs);
    MC_LAUNCH_CHECK();
    return MC_OK;
  }
  int fast_is_col = (p_cs == 1) ? 1 : 0;
  int64_t total = (int64_t)rows * cols;
  int blocks = (int)((total + 255) / 256);
  int cap = mc::num_sms() * 16;
  if (blocks > cap) blocks = cap;
  mc::soft_ce_bwd_kernel<<<blocks, 256, 0, st>>>(P, T, rows, cols, row_lse, row_tsum, grad_rows,
                                                 dpreds, dp_rs, dp_cs, dtargets, dt_rs, dt_cs,
                                                 fast_is_col);
  MC_LAUNCH_CHECK();
  return MC_OK;
}

}  // extern "C"
