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

// ---- forward / backward for the TRANSPOSED views of CLIP.py:41 (row stride 1): 16-byte vectors ----
// Memory is [c][r] with r contiguous, so a thread owns FOUR consecutive rows (one float4 per
// column) and walks the columns of its group with four columns in flight.  block = 8 row-quads (x)
// x 32 column groups (y): a warp touches four full 128-byte lines per load instruction.
// Online log-sum-exp over four columns in flight for the four rows a lane owns: one running-max update (and at
// most one rescale) per row per group of four columns, then four exponentials - about a third of the
// instructions of an element-by-element update, which made the transposed forward issue-bound.
template <typename Live>
__device__ __forceinline__ void online_lse4(const float4 (&a)[4], const float4 (&b)[4], Live live, float (&m)[4],
                                            float (&sacc)[4], float (&tsum)[4], float (&tp)[4]) {
  float av[4][4], bv[4][4];
  bool okv[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const bool ok = okv[u] = live(u);
    av[u][0] = ok ? a[u].x : -INFINITY; av[u][1] = ok ? a[u].y : -INFINITY;
    av[u][2] = ok ? a[u].z : -INFINITY; av[u][3] = ok ? a[u].w : -INFINITY;
    bv[u][0] = ok ? b[u].x : 0.f; bv[u][1] = ok ? b[u].y : 0.f; bv[u][2] = ok ? b[u].z : 0.f; bv[u][3] = ok ? b[u].w : 0.f;
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float m4 = fmaxf(fmaxf(av[0][e], av[1][e]), fmaxf(av[2][e], av[3][e]));
    if (m4 > m[e]) {
      sacc[e] *= __expf(m[e] - m4);  // exp(-inf) = 0 covers the first group
      m[e] = m4;
    }
    if (m[e] > -INFINITY) {  // an all-dead group before any live one leaves the accumulators untouched
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        sacc[e] += __expf(av[u][e] - m[e]);  // dead columns: exp(-inf) = 0
        tsum[e] += bv[u][e];
        if (okv[u]) tp[e] = fmaf(av[u][e], bv[u][e], tp[e]);  // a genuine -inf logit still propagates like torch
      }
    }
  }
}

template <int QUADS>  // row-quads per block: 4 * QUADS rows, 256 / QUADS column groups
__global__ void __launch_bounds__(256) soft_ce_fwd_tvec(const float* __restrict__ preds, int64_t p_cs,
                                                        const float* __restrict__ tg, int64_t t_cs, int rows, int cols,
                                                        float* __restrict__ loss_rows, float* __restrict__ row_lse,
                                                        float* __restrict__ row_tsum) {
  constexpr int G = 256 / QUADS, R = 4 * QUADS;
  __shared__ float sm_m[G][R + 1], sm_s[G][R + 1], sm_t[G][R + 1], sm_tp[G][R + 1];
  const int tx = threadIdx.x % QUADS, ty = threadIdx.x / QUADS;
  const int r0 = blockIdx.x * R + 4 * tx;
  float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY}, sacc[4] = {0.f, 0.f, 0.f, 0.f};
  float tsum[4] = {0.f, 0.f, 0.f, 0.f}, tp[4] = {0.f, 0.f, 0.f, 0.f};
  if (r0 < rows) {  // rows % 4 == 0: a quad is entirely inside or outside
    for (int c0 = ty; c0 < cols; c0 += 4 * G) {
      float4 a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = c0 + G * u;
        if (c < cols) {
          a[u] = ld_stream(reinterpret_cast<const float4*>(preds + (int64_t)c * p_cs + r0));
          b[u] = ld_stream(reinterpret_cast<const float4*>(tg + (int64_t)c * t_cs + r0));
        }
      }
      online_lse4(a, b, [&](int u) { return c0 + G * u < cols; }, m, sacc, tsum, tp);
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    sm_m[ty][4 * tx + e] = m[e]; sm_s[ty][4 * tx + e] = sacc[e];
    sm_t[ty][4 * tx + e] = tsum[e]; sm_tp[ty][4 * tx + e] = tp[e];
  }
  __syncthreads();
  // each warp reduces R / 8 rows of the block: lanes stride over the G column-group partials of a row
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int e = 0; e < R / 8; ++e) {
    const int rr = (R / 8) * warp + e;
    Lse t;
    t.init();
    float ts = 0.f, tps = 0.f;
    for (int gi = lane; gi < G; gi += 32) {
      t.merge(sm_m[gi][rr], sm_s[gi][rr]);
      ts += sm_t[gi][rr];
      tps += sm_tp[gi][rr];
    }
    warp_merge_lse(t);
    ts = warp_sum(ts);
    tps = warp_sum(tps);
    const int row = blockIdx.x * R + rr;
    if (lane == 0 && row < rows) {
      const float lse = t.value();
      loss_rows[row] = lse * ts - tps;
      if (row_lse) row_lse[row] = lse;
      if (row_tsum) row_tsum[row] = ts;
    }
  }
}

// Column-split form for tall-enough problems: block = 64 row-quads (256 rows: a warp reads 512 contiguous
// bytes of a column) x 4 column phases, grid.y = column slices; every block writes the (max, sum, sum t,
// sum t*p) partials of its 256 rows for its slice, a second kernel merges the slices in a fixed order.
__global__ void __launch_bounds__(256) soft_ce_fwd_tsplit(const float* __restrict__ preds, int64_t p_cs,
                                                          const float* __restrict__ tg, int64_t t_cs, int rows, int cols,
                                                          int cols_per_slice, float4* __restrict__ part /* [slices][rows] */) {
  __shared__ float4 sm[4][256];
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const int r0 = (blockIdx.x * 64 + tx) * 4;
  const int cbeg = blockIdx.y * cols_per_slice, cend = min(cols, cbeg + cols_per_slice);
  float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY}, sacc[4] = {0.f, 0.f, 0.f, 0.f};
  float tsum[4] = {0.f, 0.f, 0.f, 0.f}, tp[4] = {0.f, 0.f, 0.f, 0.f};
  if (r0 < rows) {
    for (int c0 = cbeg + ty; c0 < cend; c0 += 16) {
      float4 a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = c0 + 4 * u;
        if (c < cend) {
          a[u] = ld_stream(reinterpret_cast<const float4*>(preds + (int64_t)c * p_cs + r0));
          b[u] = ld_stream(reinterpret_cast<const float4*>(tg + (int64_t)c * t_cs + r0));
        }
      }
      online_lse4(a, b, [&](int u) { return c0 + 4 * u < cend; }, m, sacc, tsum, tp);
    }
  }
  // merge the four column phases of a row quad through shared memory: phase p holds row element e at [p][4*tx+e]
#pragma unroll
  for (int e = 0; e < 4; ++e) sm[ty][4 * tx + e] = make_float4(m[e], sacc[e], tsum[e], tp[e]);
  __syncthreads();
  const int rl = threadIdx.x;  // one of the block's 256 rows
  const int row = blockIdx.x * 256 + rl;
  if (row < rows) {
    Lse t;
    t.init();
    float ts = 0.f, tps = 0.f;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const float4 v = sm[p][rl];
      t.merge(v.x, v.y);
      ts += v.z;
      tps += v.w;
    }
    part[(size_t)blockIdx.y * rows + row] = make_float4(t.m, t.s, ts, tps);
  }
}

__global__ void __launch_bounds__(256) soft_ce_fwd_tmerge(const float4* __restrict__ part, int slices, int rows,
                                                          float* __restrict__ loss_rows, float* __restrict__ row_lse,
                                                          float* __restrict__ row_tsum) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  Lse t;
  t.init();
  float ts = 0.f, tps = 0.f;
  for (int s = 0; s < slices; ++s) {
    const float4 v = part[(size_t)s * rows + row];
    t.merge(v.x, v.y);
    ts += v.z;
    tps += v.w;
  }
  const float lse = t.value();
  loss_rows[row] = lse * ts - tps;
  if (row_lse) row_lse[row] = lse;
  if (row_tsum) row_tsum[row] = ts;
}

// column slices so that ~4 blocks per SM exist; 0 = use the single-kernel form
static int tsplit_slices(int rows, int cols) {
  if (rows < 1024 || cols < 256) return 0;
  const int gx = (rows / 4 + 63) / 64;
  int slices = (mc::num_sms() * 4 + gx - 1) / gx;
  if (slices > cols / 64) slices = cols / 64;
  return slices < 2 ? 0 : slices;
}

// grid.x: blocks of 64 row-quads (256 rows), grid.y: column slices; thread = (row-quad, 4 column phases)
__global__ void __launch_bounds__(256) soft_ce_bwd_tvec(const float* __restrict__ preds, int64_t p_cs,
                                                        const float* __restrict__ tg, int64_t t_cs, int rows, int cols,
                                                        const float* __restrict__ row_lse,
                                                        const float* __restrict__ row_tsum,
                                                        const float* __restrict__ grad, float* __restrict__ dp,
                                                        int64_t dp_cs, float* __restrict__ dt, int64_t dt_cs,
                                                        int cols_per_slice) {
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const int r0 = (blockIdx.x * 64 + tx) * 4;
  if (r0 >= rows) return;
  const float4 g4 = *reinterpret_cast<const float4*>(grad + r0);
  const float4 l4 = *reinterpret_cast<const float4*>(row_lse + r0);
  const float4 s4 = *reinterpret_cast<const float4*>(row_tsum + r0);
  const int cbeg = blockIdx.y * cols_per_slice;
  const int cend = min(cols, cbeg + cols_per_slice);
  for (int c0 = cbeg + ty; c0 < cend; c0 += 16) {
    float4 a[4], b[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = c0 + 4 * u;
      if (c < cend) {
        a[u] = ld_stream(reinterpret_cast<const float4*>(preds + (int64_t)c * p_cs + r0));
        if (dp) b[u] = ld_stream(reinterpret_cast<const float4*>(tg + (int64_t)c * t_cs + r0));
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = c0 + 4 * u;
      if (c >= cend) continue;
      const float lx = a[u].x - l4.x, ly = a[u].y - l4.y, lz = a[u].z - l4.z, lw = a[u].w - l4.w;
      if (dp)
        st_stream(reinterpret_cast<float4*>(dp + (int64_t)c * dp_cs + r0),
                  make_float4(g4.x * (__expf(lx) * s4.x - b[u].x), g4.y * (__expf(ly) * s4.y - b[u].y),
                              g4.z * (__expf(lz) * s4.z - b[u].z), g4.w * (__expf(lw) * s4.w - b[u].w)));
      if (dt)
        st_stream(reinterpret_cast<float4*>(dt + (int64_t)c * dt_cs + r0),
                  make_float4(-g4.x * lx, -g4.y * ly, -g4.z * lz, -g4.w * lw));
    }
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
    // four 16-byte loads per operand in flight per lane before the first store
    for (int v0 = 0; v0 < nv; v0 += 128) {
      float4 a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int v = v0 + lane + 32 * u;
        if (v < nv) {
          a[u] = ld_stream(pr + v);
          if (dpr) b[u] = ld_stream(tr + v);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int v = v0 + lane + 32 * u;
        if (v >= nv) continue;
        const float lx = a[u].x - lse, ly = a[u].y - lse, lz = a[u].z - lse, lw = a[u].w - lse;
        if (dpr)
          st_stream(dpr + v, make_float4(g * (__expf(lx) * ts - b[u].x), g * (__expf(ly) * ts - b[u].y),
                                         g * (__expf(lz) * ts - b[u].z), g * (__expf(lw) * ts - b[u].w)));
        if (dtr) st_stream(dtr + v, make_float4(-g * lx, -g * ly, -g * lz, -g * lw));
      }
    }
  }
}

}  // namespace mc

extern "C" {

size_t mc_soft_ce_workspace_bytes(int rows, int cols) {
  if (rows <= 0 || cols <= 0) return 0;
  const int slices = mc::tsplit_slices(rows, cols);
  return slices ? (size_t)slices * rows * sizeof(float4) : 0;
}

int mc_soft_ce_fwd(const float* preds, int64_t p_rs, int64_t p_cs, const float* targets,
                   int64_t t_rs, int64_t t_cs, int rows, int cols, float* loss_rows,
                   float* row_lse, float* row_tsum, void* ws, size_t ws_bytes, void* stream) {
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
  } else if (p_rs == 1 && t_rs == 1 && rows % 4 == 0 && p_cs % 4 == 0 && t_cs % 4 == 0 && mc::aligned(preds, 16) &&
             mc::aligned(targets, 16) && ws && mc::aligned(ws, 16) && mc::tsplit_slices(rows, cols) > 0 &&
             ws_bytes >= mc_soft_ce_workspace_bytes(rows, cols)) {
    const int slices = mc::tsplit_slices(rows, cols);
    const int cps = ((cols + slices - 1) / slices + 15) / 16 * 16;
    const int ny = (cols + cps - 1) / cps;
    dim3 grid((rows / 4 + 63) / 64, ny);
    mc::soft_ce_fwd_tsplit<<<grid, 256, 0, st>>>(preds, p_cs, targets, t_cs, rows, cols, cps, static_cast<float4*>(ws));
    MC_LAUNCH_CHECK();
    mc::soft_ce_fwd_tmerge<<<(rows + 255) / 256, 256, 0, st>>>(static_cast<const float4*>(ws), ny, rows, loss_rows, row_lse,
                                                              row_tsum);
  } else if (p_rs == 1 && t_rs == 1 && rows % 4 == 0 && p_cs % 4 == 0 && t_cs % 4 == 0 && mc::aligned(preds, 16) &&
             mc::aligned(targets, 16)) {
    // measured on 8192 x 8192: 32-row blocks 150 us, 16-row blocks 156 us, 128-row blocks (64 blocks only) 425 us
    mc::soft_ce_fwd_tvec<8><<<(rows + 31) / 32, 256, 0, st>>>(preds, p_cs, targets, t_cs, rows, cols, loss_rows,
                                                              row_lse, row_tsum);
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
  const bool tvec = p_rs == 1 && t_rs == 1 && (!dpreds || dp_rs == 1) && (!dtargets || dt_rs == 1) && rows % 4 == 0 &&
                    p_cs % 4 == 0 && t_cs % 4 == 0 && (!dpreds || dp_cs % 4 == 0) && (!dtargets || dt_cs % 4 == 0) &&
                    mc::aligned(preds, 16) && mc::aligned(targets, 16) && (!dpreds || mc::aligned(dpreds, 16)) &&
                    (!dtargets || mc::aligned(dtargets, 16)) && mc::aligned(row_lse, 16) && mc::aligned(row_tsum, 16) &&
                    mc::aligned(grad_rows, 16);
  if (tvec) {
    const int gx = (rows / 4 + 63) / 64;
    int slices = (mc::num_sms() * 4 + gx - 1) / gx;           // ~4 blocks per SM in total
    if (slices > (cols + 15) / 16) slices = (cols + 15) / 16;
    if (slices < 1) slices = 1;
    const int cps = ((cols + slices - 1) / slices + 15) / 16 * 16;
    dim3 grid(gx, (cols + cps - 1) / cps);
    mc::soft_ce_bwd_tvec<<<grid, 256, 0, st>>>(preds, p_cs, targets, t_cs, rows, cols, row_lse, row_tsum, grad_rows,
                                               dpreds, dp_cs, dtargets, dt_cs, cps);
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
