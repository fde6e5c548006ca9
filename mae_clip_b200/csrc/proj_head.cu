// L1-L2: ProjectionHead forward/backward (/root/reference modules.py:55-76).
//   projected = x Wp^T + bp ; hidden = gelu(projected) ; y = hidden Wf^T + bf
//   z = keep * y / (1-p) + projected ; out = LayerNorm(z) * gamma + beta
// GEMMs: bias and exact-erf GELU are fused into the GEMM epilogue; dropout + residual +
// LayerNorm are one warp-per-row pass (K2), its backward another (K2b) that also emits the
// gamma/beta column partials.  The dropout keep-mask is an INPUT (SURVEY.md section 7 f).
#include <stdlib.h>

#include "common.cuh"
#include "gemm_tc.cuh"
#include "head_tc.cuh"

namespace mc {

// ---------------- K2: dropout + residual + LayerNorm forward (warp per row) ----------------
template <int NV>  // float4 slots per lane: P <= 128*NV
__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ y,
                                                     const float* __restrict__ projected,
                                                     const uint8_t* __restrict__ keep, float scale,
                                                     const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, float eps, int B,
                                                     int P, float* __restrict__ z_out,
                                                     float* __restrict__ mean_out,
                                                     float* __restrict__ rstd_out,
                                                     float* __restrict__ out) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= B) return;
  const int nvec = P >> 2;
  const size_t off = (size_t)row * P;
  float4 zv[NV];
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int v = lane + 32 * k;
    zv[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (v < nvec) {
      float4 a = *reinterpret_cast<const float4*>(y + off + 4 * v);
      float4 pr = *reinterpret_cast<const float4*>(projected + off + 4 * v);
      if (keep) {
        uchar4 m = *reinterpret_cast<const uchar4*>(keep + off + 4 * v);
        a.x = m.x ? a.x * scale : 0.f;
        a.y = m.y ? a.y * scale : 0.f;
        a.z = m.z ? a.z * scale : 0.f;
        a.w = m.w ? a.w * scale : 0.f;
      }
      zv[k] = make_float4(a.x + pr.x, a.y + pr.y, a.z + pr.z, a.w + pr.w);
      sum += (zv[k].x + zv[k].y) + (zv[k].z + zv[k].w);
    }
  }
  const float mean = warp_sum(sum) / (float)P;
  float var = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    if (lane + 32 * k < nvec) {
      float a = zv[k].x - mean, b = zv[k].y - mean, c = zv[k].z - mean, d = zv[k].w - mean;
      var += (a * a + b * b) + (c * c + d * d);
    }
  }
  const float rstd = rsqrtf(warp_sum(var) / (float)P + eps);
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int v = lane + 32 * k;
    if (v < nvec) {
      float4 g = *reinterpret_cast<const float4*>(gamma + 4 * v);
      float4 bt = *reinterpret_cast<const float4*>(beta + 4 * v);
      float4 o;
      o.x = (zv[k].x - mean) * rstd * g.x + bt.x;
      o.y = (zv[k].y - mean) * rstd * g.y + bt.y;
      o.z = (zv[k].z - mean) * rstd * g.z + bt.z;
      o.w = (zv[k].w - mean) * rstd * g.w + bt.w;
      *reinterpret_cast<float4*>(out + off + 4 * v) = o;
      if (z_out) *reinterpret_cast<float4*>(z_out + off + 4 * v) = zv[k];
    }
  }
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
}

// ---------------- K2b: LayerNorm + dropout backward (warp per row, grid-stride rows) ---------
// dz = rstd * (dxhat - mean(dxhat) - xhat * mean(dxhat * xhat)), dxhat = dO * gamma
// dy = dz * keep / (1-p);  per-block partials of dgamma = sum dO*xhat, dbeta = sum dO, db_fc = sum dy
template <int NV>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* __restrict__ grad_out,
                                                     const float* __restrict__ z,
                                                     const float* __restrict__ mean,
                                                     const float* __restrict__ rstd,
                                                     const float* __restrict__ gamma,
                                                     const uint8_t* __restrict__ keep, float scale,
                                                     int B, int P, float* __restrict__ dz_out,
                                                     float* __restrict__ dy_out,
                                                     float* __restrict__ partials /*[grid][3][P]*/,
                                                     unsigned int* __restrict__ dy_amax /* or null */,
                                                     unsigned int* __restrict__ zero_words /* 64 words to clear, or null */) {
  extern __shared__ float sm[];  // [8 warps][3][P]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = P >> 2;
  if (zero_words && blockIdx.x == 0 && threadIdx.x < 64) zero_words[threadIdx.x] = 0u;   // counters of colsum_tall_kernel
  float amx = 0.f;  // max |dy| this thread wrote: saves the staging of dy a pass over it
  float4 dg[NV], db[NV], dbf[NV], gm[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    dg[k] = db[k] = dbf[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int v = lane + 32 * k;
    gm[k] = (v < nvec) ? *reinterpret_cast<const float4*>(gamma + 4 * v) : dg[k];
  }
  for (int row = blockIdx.x * 8 + warp; row < B; row += gridDim.x * 8) {
    const size_t off = (size_t)row * P;
    const float mu = mean[row], rs = rstd[row];
    float4 xh[NV], dxh[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int v = lane + 32 * k;
      xh[k] = dxh[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (v < nvec) {
        float4 go = *reinterpret_cast<const float4*>(grad_out + off + 4 * v);
        float4 zz = *reinterpret_cast<const float4*>(z + off + 4 * v);
        xh[k] = make_float4((zz.x - mu) * rs, (zz.y - mu) * rs, (zz.z - mu) * rs, (zz.w - mu) * rs);
        dxh[k] = make_float4(go.x * gm[k].x, go.y * gm[k].y, go.z * gm[k].z, go.w * gm[k].w);
        dg[k].x += go.x * xh[k].x; dg[k].y += go.y * xh[k].y;
        dg[k].z += go.z * xh[k].z; dg[k].w += go.w * xh[k].w;
        db[k].x += go.x; db[k].y += go.y; db[k].z += go.z; db[k].w += go.w;
        s1 += (dxh[k].x + dxh[k].y) + (dxh[k].z + dxh[k].w);
        s2 += (dxh[k].x * xh[k].x + dxh[k].y * xh[k].y) + (dxh[k].z * xh[k].z + dxh[k].w * xh[k].w);
      }
    }
    const float m1 = warp_sum(s1) / (float)P, m2 = warp_sum(s2) / (float)P;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int v = lane + 32 * k;
      if (v < nvec) {
        float4 d;
        d.x = rs * (dxh[k].x - m1 - xh[k].x * m2);
        d.y = rs * (dxh[k].y - m1 - xh[k].y * m2);
        d.z = rs * (dxh[k].z - m1 - xh[k].z * m2);
        d.w = rs * (dxh[k].w - m1 - xh[k].w * m2);
        *reinterpret_cast<float4*>(dz_out + off + 4 * v) = d;
        if (keep) {
          uchar4 m = *reinterpret_cast<const uchar4*>(keep + off + 4 * v);
          d.x = m.x ? d.x * scale : 0.f;
          d.y = m.y ? d.y * scale : 0.f;
          d.z = m.z ? d.z * scale : 0.f;
          d.w = m.w ? d.w * scale : 0.f;
        }
        *reinterpret_cast<float4*>(dy_out + off + 4 * v) = d;
        dbf[k].x += d.x; dbf[k].y += d.y; dbf[k].z += d.z; dbf[k].w += d.w;
        amx = fmaxf(amx, fmaxf(fmaxf(fabsf(d.x), fabsf(d.y)), fmaxf(fabsf(d.z), fabsf(d.w))));
      }
    }
  }
  // block-level column partials
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int v = lane + 32 * k;
    if (v < nvec) {
      *reinterpret_cast<float4*>(sm + (warp * 3 + 0) * P + 4 * v) = dg[k];
      *reinterpret_cast<float4*>(sm + (warp * 3 + 1) * P + 4 * v) = db[k];
      *reinterpret_cast<float4*>(sm + (warp * 3 + 2) * P + 4 * v) = dbf[k];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * P; i += blockDim.x) {
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) acc += sm[w * 3 * P + i];
    partials[(size_t)blockIdx.x * 3 * P + i] = acc;
  }
  if (dy_amax) {   // one atomic per block, not per warp: same-address atomics serialise in L2
    amx = warp_max(amx);
    __syncthreads();
    if (lane == 0) sm[warp] = amx;
    __syncthreads();
    if (threadIdx.x == 0) {
      float a = sm[0];
#pragma unroll
      for (int w = 1; w < 8; ++w) a = fmaxf(a, sm[w]);
      if (a > 0.f) atomicMax(dy_amax, __float_as_uint(a));
    }
  }
}

// out_k[c] = sum_r in[r][k * P + c] for k = 0, 1, 2 (rows x 3P partials of ln_bwd_kernel -> dgamma, dbeta, db_fc)
__global__ void __launch_bounds__(1024) colsum3_kernel(const float* __restrict__ in, int rows, int P,
                                                       float* __restrict__ out0, float* __restrict__ out1,
                                                       float* __restrict__ out2) {
  __shared__ float sm[32][33];
  const int cols = 3 * P;
  const int c = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (c < cols)
    for (int r = threadIdx.y; r < rows; r += 32) acc += in[(size_t)r * cols + c];
  sm[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  float v = sm[threadIdx.x][threadIdx.y];
  v = warp_sum(v);
  const int oc = blockIdx.x * 32 + threadIdx.y;
  if (threadIdx.x == 0 && oc < cols) {
    float* o = oc < P ? out0 : (oc < 2 * P ? out1 : out2);
    o[oc % P] = v;
  }
}

// Tall column sums in ONE launch, deterministic: grid (cols / 32, kColSlices).  Block (bx, by) sums its slice of the rows
// for 32 columns into part2[by][c]; the last slice to arrive for a column group (per-group counter, zeroed by an earlier
// kernel of the same call and re-armed here) folds the kColSlices partial rows in a fixed order.  Output column c goes
// to out_k[c % P] with k = c / P (the three LayerNorm-backward sums share one partial buffer).
constexpr int kColSlices = 8;
__global__ void __launch_bounds__(1024) colsum_tall_kernel(const float* __restrict__ in, int rows, int cols, int P,
                                                           float* __restrict__ part2, unsigned int* __restrict__ counters,
                                                           float* __restrict__ out0, float* __restrict__ out1,
                                                           float* __restrict__ out2) {
  __shared__ float sm[32][33];
  __shared__ bool last;
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int per = (rows + kColSlices - 1) / kColSlices;
  const int r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
  float acc = 0.f;
  if (c < cols)
    for (int r = r0 + threadIdx.y; r < r1; r += 32) acc += in[(size_t)r * cols + c];
  sm[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  float v = sm[threadIdx.x][threadIdx.y];
  v = warp_sum(v);
  const int oc = blockIdx.x * 32 + threadIdx.y;
  if (threadIdx.x == 0 && oc < cols) part2[(size_t)blockIdx.y * cols + oc] = v;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    const unsigned int n = atomicAdd(counters + blockIdx.x, 1u);
    last = n == kColSlices - 1;
    if (last) counters[blockIdx.x] = 0u;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  if (threadIdx.y == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < kColSlices; ++k) t += *reinterpret_cast<volatile const float*>(part2 + (size_t)k * cols + c);
    float* o = c < P ? out0 : (c < 2 * P ? out1 : out2);
    o[c % P] = t;
  }
}

// out[c] = sum_r in[r][c]   (rows x cols, ld = cols); one block per 32 columns
__global__ void __launch_bounds__(1024) colsum_kernel(const float* __restrict__ in, int rows,
                                                      int cols, float* __restrict__ out) {
  __shared__ float sm[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (c < cols)
    for (int r = threadIdx.y; r < rows; r += 32) acc += in[(size_t)r * cols + c];
  sm[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  float v = sm[threadIdx.x][threadIdx.y];
  v = warp_sum(v);
  const int oc = blockIdx.x * 32 + threadIdx.y;
  if (threadIdx.x == 0 && oc < cols) out[oc] = v;
}

// first stage of a tall column sum: block b sums its contiguous slice of rows with 16-byte loads
// (four in flight per thread) and writes part[b][cols]; colsum_kernel then folds the few partial rows.
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ in, int rows, int cols,
                                                             float* __restrict__ part) {
  __shared__ float4 sm[256];
  const int cols4 = cols >> 2, rsubs = 256 / cols4;
  const int c4 = threadIdx.x % cols4, rsub = threadIdx.x / cols4;
  const int per = (rows + gridDim.x - 1) / gridDim.x;
  const int rbeg = blockIdx.x * per, rend = min(rows, rbeg + per);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (rsub < rsubs) {
    const float4* p = reinterpret_cast<const float4*>(in) + c4;
    int r = rbeg + rsub;
    for (; r + 3 * rsubs < rend; r += 4 * rsubs) {
      const float4 a = ld_stream(p + (size_t)r * cols4), b = ld_stream(p + (size_t)(r + rsubs) * cols4);
      const float4 c = ld_stream(p + (size_t)(r + 2 * rsubs) * cols4), d = ld_stream(p + (size_t)(r + 3 * rsubs) * cols4);
      acc.x += (a.x + b.x) + (c.x + d.x); acc.y += (a.y + b.y) + (c.y + d.y);
      acc.z += (a.z + b.z) + (c.z + d.z); acc.w += (a.w + b.w) + (c.w + d.w);
    }
    for (; r < rend; r += rsubs) {
      const float4 a = ld_stream(p + (size_t)r * cols4);
      acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
    }
  }
  sm[threadIdx.x] = acc;
  __syncthreads();
  if (rsub == 0) {
    for (int k = 1; k < rsubs; ++k) {
      const float4 o = sm[k * cols4 + c4];
      acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
    }
    reinterpret_cast<float4*>(part + (size_t)blockIdx.x * cols)[c4] = acc;
  }
}

// dp = dh * gelu'(projected) + dz   (in place on dh)
__global__ void __launch_bounds__(256) gelu_bwd_add_kernel(float* __restrict__ dh,
                                                           const float* __restrict__ projected,
                                                           const float* __restrict__ dz, size_t n4,
                                                           unsigned int* __restrict__ dp_amax /* or null */) {
  float amx = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (size_t)gridDim.x * blockDim.x) {
    float4 h = reinterpret_cast<float4*>(dh)[i];
    float4 p = reinterpret_cast<const float4*>(projected)[i];
    float4 d = reinterpret_cast<const float4*>(dz)[i];
    h.x = fmaf(h.x, gelu_erf_grad(p.x), d.x);
    h.y = fmaf(h.y, gelu_erf_grad(p.y), d.y);
    h.z = fmaf(h.z, gelu_erf_grad(p.z), d.z);
    h.w = fmaf(h.w, gelu_erf_grad(p.w), d.w);
    reinterpret_cast<float4*>(dh)[i] = h;
    amx = fmaxf(amx, fmaxf(fmaxf(fabsf(h.x), fabsf(h.y)), fmaxf(fabsf(h.z), fabsf(h.w))));
  }
  if (dp_amax) {
    amx = warp_max(amx);
    if ((threadIdx.x & 31) == 0 && amx > 0.f) atomicMax(dp_amax, __float_as_uint(amx));
  }
}

static int ln_blocks(int B) {
  int nb = (B + 7) / 8;
  // ln_bwd_kernel keeps 77 registers per thread: three 256-thread blocks fit an SM.  One resident wave of grid-stride
  // blocks (a larger grid ran 2.7 waves with a partly filled last one) - and fewer partial rows to fold afterwards.
  int cap = num_sms() * 3;
  return nb < cap ? nb : cap;
}

// out[c] = sum_r in[r][c]; tall inputs go through per-block partials in `scratch` (>= scratch_rows x cols floats)
static int colsum_rows(const float* in, int rows, int cols, float* out, float* scratch, int scratch_rows,
                       cudaStream_t st) {
  dim3 blk(32, 32);
  const bool two_stage = rows >= 4096 && cols % 4 == 0 && cols <= 1024 && aligned(in, 16) && scratch && scratch_rows >= 8;
  if (!two_stage) {
    colsum_kernel<<<(cols + 31) / 32, blk, 0, st>>>(in, rows, cols, out);
    MC_LAUNCH_CHECK();
    return MC_OK;
  }
  int nb = (rows + 63) / 64;
  if (nb > scratch_rows) nb = scratch_rows;
  if (nb > num_sms() * 4) nb = num_sms() * 4;
  colsum_partial_kernel<<<nb, 256, 0, st>>>(in, rows, cols, scratch);
  MC_LAUNCH_CHECK();
  colsum_kernel<<<(cols + 31) / 32, blk, 0, st>>>(scratch, nb, cols, out);
  MC_LAUNCH_CHECK();
  return MC_OK;
}

template <int NV>
static int launch_ln_fwd(const float* y, const float* projected, const uint8_t* keep, float scale,
                         const float* gamma, const float* beta, float eps, int B, int P, float* z,
                         float* mean, float* rstd, float* out, cudaStream_t st) {
  ln_fwd_kernel<NV><<<(B + 7) / 8, 256, 0, st>>>(y, projected, keep, scale, gamma, beta, eps, B, P, z,
                                                 mean, rstd, out);
  MC_LAUNCH_CHECK();
  return MC_OK;
}

template <int NV>
static int launch_ln_bwd(const float* go, const float* z, const float* mean, const float* rstd,
                         const float* gamma, const uint8_t* keep, float scale, int B, int P,
                         float* dz, float* dy, float* partials, int blocks, unsigned int* dy_amax,
                         cudaStream_t st, unsigned int* zero_words = nullptr) {
  size_t smem = (size_t)8 * 3 * P * sizeof(float);
  ln_bwd_kernel<NV><<<blocks, 256, smem, st>>>(go, z, mean, rstd, gamma, keep, scale, B, P, dz, dy,
                                               partials, dy_amax, zero_words);
  MC_LAUNCH_CHECK();
  return MC_OK;
}

#define MC_DISPATCH_NV(P, CALL)                                \
  do {                                                         \
    if ((P) <= 128) { constexpr int NV = 1; rc = CALL; }       \
    else if ((P) <= 256) { constexpr int NV = 2; rc = CALL; }  \
    else if ((P) <= 512) { constexpr int NV = 4; rc = CALL; }  \
    else { constexpr int NV = 8; rc = CALL; }                  \
  } while (0)

struct HeadWs {
  float *y, *hidden_tmp, *dz, *dy, *dh, *partials;
  size_t total;
};
static HeadWs head_ws(void* ws, int B, int /*E*/, int P) {
  size_t bp = round_up((size_t)B * P * 4, 256);
  size_t part = round_up((size_t)ln_blocks(B) * 3 * P * 4, 256);
  char* p = static_cast<char*>(ws);
  HeadWs w;
  w.y = reinterpret_cast<float*>(p);
  w.hidden_tmp = reinterpret_cast<float*>(p + bp);
  w.dz = reinterpret_cast<float*>(p);           // backward reuses the same region
  w.dy = reinterpret_cast<float*>(p + bp);
  w.dh = reinterpret_cast<float*>(p + 2 * bp);
  w.partials = reinterpret_cast<float*>(p + 3 * bp);
  w.total = 3 * bp + part;
  return w;
}


// ---------------- tcgen05 path: workspace map (bump allocation, every piece 256-byte aligned) ----------------
struct TcHeadWs {
  // fp32 temporaries (same roles as HeadWs)
  float *y, *hidden_tmp, *dz, *dy, *dh, *partials;
  unsigned int* amax;                                // producer-side max |.|: [0] dy, [1] dp (zeroed per backward)
  // staged operands
  tcg::Planes x, wp, h, wf;                          // forward
  tcg::Planes dyT, hT, dy_p, wfT, dpT, xT, dp_p, wpT;  // backward
  void* gemm_ws;
  size_t gemm_ws_bytes;
  size_t total;
};
static TcHeadWs tc_head_ws(void* ws, int B, int E, int P) {
  TcHeadWs w;
  char* base = static_cast<char*>(ws);
  size_t off = 0;
  auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += round_up(bytes, 256); return p; };
  auto planes = [&](int rows, int cols) {
    void* mem = take(tcg::planes_bytes(rows, cols));
    return mem ? tcg::carve_planes(mem, rows, cols) : tcg::Planes{nullptr, nullptr, nullptr, rows, cols, 0};
  };
  const size_t bp = (size_t)B * P * 4;
  w.y = reinterpret_cast<float*>(take(bp));
  w.hidden_tmp = reinterpret_cast<float*>(take(bp));
  w.dh = reinterpret_cast<float*>(take(bp));
  w.dz = w.y;            // backward reuses the forward temporaries
  w.dy = w.hidden_tmp;
  w.partials = reinterpret_cast<float*>(take((size_t)ln_blocks(B) * 3 * P * 4));
  w.amax = reinterpret_cast<unsigned int*>(take(256));
  w.x = planes(B, E);  w.wp = planes(P, E);  w.h = planes(B, P);  w.wf = planes(P, P);
  w.dyT = planes(P, B); w.hT = planes(P, B); w.dy_p = planes(B, P); w.wfT = planes(P, P);
  w.dpT = planes(P, B); w.xT = planes(E, B); w.dp_p = planes(B, P); w.wpT = planes(E, P);
  size_t g = tcg::gemm_workspace_bytes(B, P, E);
  size_t cands[5] = {tcg::gemm_workspace_bytes(B, P, P), tcg::gemm_workspace_bytes(P, P, B),
                     tcg::gemm_workspace_bytes(P, E, B), tcg::gemm_workspace_bytes(B, E, P), 256};
  for (size_t c : cands) g = c > g ? c : g;
  w.gemm_ws_bytes = g;
  w.gemm_ws = take(g);
  w.total = off;
  return w;
}

// The tensor-core path stages every GEMM operand (about 40 launches per backward): below ~2048 rows
// a head is launch-bound and the true-fp32 FMA path, with a third of the launches, is faster
// (measured: B = 256, E = 2048 fwd+bwd 0.37 ms vs 0.53 ms; equal at B = 2048; 2x slower at 4096).
// The switch depends on the shape only, so forward and backward always agree.
constexpr int kTcMinRows = 2048;
static bool use_tc(int mode, int B) {
  return (mode == MC_GEMM_TC_F16X3 || mode == MC_GEMM_TC_F16) && B >= kTcMinRows;
}


// ---------------- CTA-pair path (csrc/head_tc.cu): workspace map ----------------
// Only WEIGHTS are staged as planes here (a few MB); activations are converted inside the GEMM kernels.
struct PairHeadWs {
  float *hidden_tmp, *dz, *dy, *dp, *partials, *colpart, *part2;
  unsigned int* counters;             // 64 words for colsum_tall_kernel (cleared by ln_bwd_kernel)
  unsigned int* amax;                 // [0] x, [1] hidden (when the caller passes no fwd_amax), [2] dy, [3] dp; [8..11] staging sync
  tcg::Planes wp, wf, wfT, wpT;       // (P, E), (P, P), (P, P)^T, (E, P)
  void* tt_ws;
  size_t tt_ws_bytes;
  size_t total;
};
static PairHeadWs pair_head_ws(void* ws, int B, int E, int P) {
  PairHeadWs w;
  char* base = static_cast<char*>(ws);
  size_t off = 0;
  auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += round_up(bytes, 256); return p; };
  auto planes = [&](int rows, int cols) {
    void* mem = take(tcg::planes_bytes(rows, cols));
    return mem ? tcg::carve_planes(mem, rows, cols) : tcg::Planes{nullptr, nullptr, nullptr, rows, cols, 0};
  };
  const size_t bp = (size_t)B * P * 4;
  w.hidden_tmp = reinterpret_cast<float*>(take(bp));
  w.dz = reinterpret_cast<float*>(take(bp));
  w.dy = w.hidden_tmp;   // backward reuses the forward temporary
  w.dp = reinterpret_cast<float*>(take(bp));
  w.partials = reinterpret_cast<float*>(take((size_t)ln_blocks(B) * 3 * P * 4));
  w.colpart = reinterpret_cast<float*>(take((size_t)hg::colpart_rows(B) * P * 4));
  w.part2 = reinterpret_cast<float*>(take((size_t)kColSlices * 3 * P * 4));
  w.counters = reinterpret_cast<unsigned int*>(take(256));
  w.amax = reinterpret_cast<unsigned int*>(take(256));
  w.wp = planes(P, E); w.wf = planes(P, P); w.wfT = planes(P, P); w.wpT = planes(E, P);
  size_t t1 = hg::tt_workspace_bytes(P, B), t2 = hg::tt_workspace_bytes(E, B);
  w.tt_ws_bytes = t1 > t2 ? t1 : t2;
  w.tt_ws = take(w.tt_ws_bytes);
  w.total = off;
  return w;
}

// The pair kernels cover the reference's shape family: projection_dim 256 (config.py:23), any embedding width that
// is a multiple of 4.  MAE_CLIP_HEAD_PAIR=0 keeps the round-1 single-CTA GEMMs (A/B switch).
static bool use_pair(int mode, int B, int E, int P) {
  static const bool off = getenv("MAE_CLIP_HEAD_PAIR") != nullptr && getenv("MAE_CLIP_HEAD_PAIR")[0] == '0';
  return !off && use_tc(mode, B) && hg::supported(P) && E % 4 == 0 && (E + 255) / 256 <= num_sms() / 2;
}

}  // namespace mc

using namespace mc;

extern "C" {

size_t mc_tc_gemm_workspace_bytes(int M, int N, int K) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  return round_up(tcg::planes_bytes(M, K), 256) + round_up(tcg::planes_bytes(N, K), 256) +
         tcg::gemm_workspace_bytes(M, N, K);
}

int mc_tc_gemm(const float* A, const float* B, int M, int N, int K, const float* bias, float* C,
               float* gelu_out, void* ws, size_t ws_bytes, void* stream) {
  MC_ARCH_GUARD();
  MC_REQUIRE(A && B && C && ws, MC_ERR_BAD_ARG, "tc_gemm: null pointer");
  MC_REQUIRE(M > 0 && N > 0 && K > 0, MC_ERR_BAD_ARG, "tc_gemm: bad sizes %d %d %d", M, N, K);
  MC_REQUIRE(aligned(ws, 256), MC_ERR_ALIGN, "tc_gemm: workspace must be 256-byte aligned");
  MC_REQUIRE(ws_bytes >= mc_tc_gemm_workspace_bytes(M, N, K), MC_ERR_WORKSPACE, "tc_gemm: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* base = static_cast<char*>(ws);
  const size_t a_bytes = round_up(tcg::planes_bytes(M, K), 256), b_bytes = round_up(tcg::planes_bytes(N, K), 256);
  tcg::Planes pa = tcg::carve_planes(base, M, K), pb = tcg::carve_planes(base + a_bytes, N, K);
  int rc;
  if ((rc = tcg::stage(A, M, K, K, 0, pa, st))) return rc;
  if ((rc = tcg::stage(B, N, K, K, 0, pb, st))) return rc;
  tcg::GemmOut o{C, N, bias, gelu_out};
  return tcg::gemm(pa, pb, M, N, K, o, gelu_out ? tcg::kEpiGelu : tcg::kEpiPlain, base + a_bytes + b_bytes,
                   ws_bytes - a_bytes - b_bytes, st);
}

size_t mc_proj_head_workspace_bytes(int B, int E, int P, int mode) {
  if (B <= 0 || E <= 0 || P <= 0) return 0;
  if (use_pair(mode, B, E, P)) return pair_head_ws(nullptr, B, E, P).total;
  return use_tc(mode, B) ? tc_head_ws(nullptr, B, E, P).total : head_ws(nullptr, B, E, P).total;
}

int mc_proj_head_fwd(const float* x, int B, int E, int P, const float* w_proj, const float* b_proj,
                     const float* w_fc, const float* b_fc, const float* gamma, const float* beta,
                     const uint8_t* keep_mask, float p_drop, float eps, int mode, float* projected,
                     float* hidden, float* z, float* mean, float* rstd, float* out, float* fwd_amax,
                     const float* prev_amax, void* ws, size_t ws_bytes, void* stream) {
  MC_ARCH_GUARD();
  MC_REQUIRE(x && w_proj && b_proj && w_fc && b_fc && gamma && beta && projected && out && ws,
             MC_ERR_BAD_ARG, "proj_head_fwd: null pointer");
  MC_REQUIRE(B > 0 && E > 0 && P > 0, MC_ERR_BAD_ARG, "proj_head_fwd: bad sizes B=%d E=%d P=%d", B, E,
             P);
  MC_REQUIRE(P % 4 == 0 && P <= 1024, MC_ERR_UNSUPPORTED,
             "proj_head_fwd: projection_dim %d must be a multiple of 4 and <= 1024", P);
  MC_REQUIRE(p_drop >= 0.f && p_drop < 1.f, MC_ERR_BAD_ARG, "proj_head_fwd: dropout p=%g", p_drop);
  MC_REQUIRE(mode >= MC_GEMM_SIMT_FP32 && mode <= MC_GEMM_TC_F16, MC_ERR_BAD_ARG, "proj_head_fwd: bad mode %d", mode);
  MC_REQUIRE(aligned(projected, 16) && aligned(out, 16) && aligned(gamma, 16) && aligned(beta, 16) &&
                 (!keep_mask || aligned(keep_mask, 4)) && (!z || aligned(z, 16)),
             MC_ERR_ALIGN, "proj_head_fwd: pointers must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float scale = 1.f / (1.f - p_drop);
  int rc;
  if (use_pair(mode, B, E, P)) {
    // CTA-pair path: x and hidden are converted fp32 -> fp16 hi/lo inside the GEMMs; bias + GELU ride on the first,
    // dropout + residual + LayerNorm on the second (N = 256 is one tile row).  Only the weights are staged.
    MC_REQUIRE(aligned(ws, 256), MC_ERR_ALIGN, "proj_head_fwd: workspace must be 256-byte aligned");
    MC_REQUIRE(aligned(x, 16), MC_ERR_ALIGN, "proj_head_fwd: x must be 16-byte aligned");
    PairHeadWs t = pair_head_ws(ws, B, E, P);
    MC_REQUIRE(ws_bytes >= t.total, MC_ERR_WORKSPACE, "proj_head_fwd: workspace %zu < %zu", ws_bytes, t.total);
    const int passes = mode == MC_GEMM_TC_F16X3 ? 3 : 1;
    // four words: [0] max |x|, [1] max |hidden| (what the backward reads), [2] / [3] max |hidden| of the first attempt /
    // of the redo (scale protocol, head_tc.cuh)
    unsigned int* am = fwd_amax ? reinterpret_cast<unsigned int*>(fwd_amax) : t.amax;
    const unsigned int* stale = reinterpret_cast<const unsigned int*>(prev_amax);
    static const bool no_stale = getenv("MAE_CLIP_HEAD_STALE_SCALE") != nullptr && getenv("MAE_CLIP_HEAD_STALE_SCALE")[0] == '0';
    if (no_stale) stale = nullptr;
    if (stale) {
      MC_CUDA(cudaMemsetAsync(am, 0, 16, st));
    } else {
      if ((rc = tcg::amax(x, B, E, E, am, st))) return rc;   // no earlier call to take the scale from: reduce max |x| first
      MC_CUDA(cudaMemsetAsync(am + 1, 0, 12, st));
    }
    // words [4], [5] of fwd_amax: max |Wp|, max |Wf| for the backward's (transposed) staging of the same weights
    if ((rc = tcg::stage_pair(tcg::StageJob{w_proj, P, E, 0, t.wp}, tcg::StageJob{w_fc, P, P, 0, t.wf}, t.amax + 8, st, nullptr,
                              am + 4)))
      return rc;
    float* hid = hidden ? hidden : t.hidden_tmp;
    hg::RowArgs f1 = {};
    f1.A = x; f1.lda = E; f1.a_amax = am; f1.W = t.wp; f1.M = B; f1.K = E; f1.passes = passes;
    f1.epilogue = hg::kEpiBiasGelu; f1.bias = b_proj; f1.out0 = projected; f1.out1 = hid;
    hg::RowArgs f2 = {};
    f2.A = hid; f2.lda = P; f2.W = t.wf; f2.M = B; f2.K = P; f2.passes = passes;
    f2.epilogue = hg::kEpiLN; f2.bias = b_fc; f2.in0 = projected; f2.keep = keep_mask; f2.drop_scale = scale; f2.eps = eps;
    f2.gamma = gamma; f2.beta = beta; f2.out0 = out; f2.out1 = z; f2.mean = mean; f2.rstd = rstd;
    if (stale) {
      f1.vmode = 1; f1.v_stale = stale; f1.v_live = am; f1.a_live = am; f1.out_amax = am + 2;
      if ((rc = hg::rows_gemm(f1, st))) return rc;
      f1.vmode = 2; f1.a_live = nullptr; f1.out_amax = am + 3;          // leaves at once unless the stale scale was unsafe
      if ((rc = hg::rows_gemm(f1, st))) return rc;
      f2.vmode = 3; f2.v_stale = stale; f2.v_live = am; f2.a_amax = am + 2; f2.a_alt = am + 3; f2.amax_publish = am + 1;
    } else {
      f1.out_amax = am + 1;
      if ((rc = hg::rows_gemm(f1, st))) return rc;
      f2.a_amax = am + 1;
    }
    return hg::rows_gemm(f2, st);
  }
  if (use_tc(mode, B)) {
    // tensor-core path: stage fp16 hi/lo planes, two tcgen05 GEMMs (bias + exact GELU fused into the first)
    MC_REQUIRE(aligned(ws, 256), MC_ERR_ALIGN, "proj_head_fwd: workspace must be 256-byte aligned");
    TcHeadWs t = tc_head_ws(ws, B, E, P);
    MC_REQUIRE(ws_bytes >= t.total, MC_ERR_WORKSPACE, "proj_head_fwd: workspace %zu < %zu", ws_bytes, t.total);
    float* hid_tc = hidden ? hidden : t.hidden_tmp;
    if ((rc = tcg::stage(x, B, E, E, 0, t.x, st))) return rc;
    if ((rc = tcg::stage(w_proj, P, E, E, 0, t.wp, st))) return rc;
    tcg::GemmOut o1{projected, P, b_proj, hid_tc};
    if ((rc = tcg::gemm(t.x, t.wp, B, P, E, o1, tcg::kEpiGelu, t.gemm_ws, t.gemm_ws_bytes, st))) return rc;
    if ((rc = tcg::stage(hid_tc, B, P, P, 0, t.h, st))) return rc;
    if (fwd_amax) {  // max |x|, max |hidden|: backward stages both again (transposed) and skips the reductions
      MC_CUDA(cudaMemcpyAsync(fwd_amax, tcg::amax_slot(t.x), 4, cudaMemcpyDeviceToDevice, st));
      MC_CUDA(cudaMemcpyAsync(fwd_amax + 1, tcg::amax_slot(t.h), 4, cudaMemcpyDeviceToDevice, st));
    }
    if ((rc = tcg::stage(w_fc, P, P, P, 0, t.wf, st))) return rc;
    tcg::GemmOut o2{t.y, P, b_fc, nullptr};
    if ((rc = tcg::gemm(t.h, t.wf, B, P, P, o2, tcg::kEpiPlain, t.gemm_ws, t.gemm_ws_bytes, st))) return rc;
    MC_DISPATCH_NV(P, (launch_ln_fwd<NV>(t.y, projected, keep_mask, scale, gamma, beta, eps, B, P, z, mean, rstd,
                                         out, st)));
    return rc;
  }
  HeadWs w = head_ws(ws, B, E, P);
  MC_REQUIRE(ws_bytes >= w.total, MC_ERR_WORKSPACE, "proj_head_fwd: workspace %zu < %zu", ws_bytes,
             w.total);
  float* hid = hidden ? hidden : w.hidden_tmp;
  SgemmArgs g1{x, E, 1, w_proj, 1, E, projected, P, B, P, E, 1.f, b_proj, hid, 0};
  if ((rc = sgemm(g1, st))) return rc;
  SgemmArgs g2{hid, P, 1, w_fc, 1, P, w.y, P, B, P, P, 1.f, b_fc, nullptr, 0};
  if ((rc = sgemm(g2, st))) return rc;
  MC_DISPATCH_NV(P, (launch_ln_fwd<NV>(w.y, projected, keep_mask, scale, gamma, beta, eps, B, P, z,
                                       mean, rstd, out, st)));
  return rc;
}

int mc_proj_head_bwd(const float* grad_out, const float* x, int B, int E, int P,
                     const float* w_proj, const float* w_fc, const float* gamma,
                     const uint8_t* keep_mask, float p_drop, int mode, const float* projected,
                     const float* hidden, const float* z, const float* mean, const float* rstd,
                     float* dx, float* dw_proj, float* db_proj, float* dw_fc, float* db_fc,
                     float* dgamma, float* dbeta, const float* fwd_amax, void* ws, size_t ws_bytes, void* stream) {
  MC_ARCH_GUARD();
  MC_REQUIRE(grad_out && x && w_proj && w_fc && gamma && projected && hidden && z && mean && rstd &&
                 dw_proj && db_proj && dw_fc && db_fc && dgamma && dbeta && ws,
             MC_ERR_BAD_ARG, "proj_head_bwd: null pointer");
  MC_REQUIRE(B > 0 && E > 0 && P > 0, MC_ERR_BAD_ARG, "proj_head_bwd: bad sizes");
  MC_REQUIRE(P % 4 == 0 && P <= 1024, MC_ERR_UNSUPPORTED, "proj_head_bwd: projection_dim %d", P);
  MC_REQUIRE(mode >= MC_GEMM_SIMT_FP32 && mode <= MC_GEMM_TC_F16, MC_ERR_BAD_ARG, "proj_head_bwd: bad mode %d", mode);
  MC_REQUIRE(aligned(grad_out, 16) && aligned(z, 16) && aligned(projected, 16) && aligned(gamma, 16) &&
                 (!keep_mask || aligned(keep_mask, 4)),
             MC_ERR_ALIGN, "proj_head_bwd: pointers must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float scale = 1.f / (1.f - p_drop);
  const int blocks = ln_blocks(B);
  int rc;
  if (use_pair(mode, B, E, P)) {
    MC_REQUIRE(aligned(ws, 256), MC_ERR_ALIGN, "proj_head_bwd: workspace must be 256-byte aligned");
    MC_REQUIRE(aligned(x, 16) && aligned(hidden, 16) && aligned(dw_proj, 16) && aligned(dw_fc, 16) && (!dx || aligned(dx, 16)),
               MC_ERR_ALIGN, "proj_head_bwd: tensors must be 16-byte aligned");
    PairHeadWs t = pair_head_ws(ws, B, E, P);
    MC_REQUIRE(ws_bytes >= t.total, MC_ERR_WORKSPACE, "proj_head_bwd: workspace %zu < %zu", ws_bytes, t.total);
    const int passes = mode == MC_GEMM_TC_F16X3 ? 3 : 1;
    MC_CUDA(cudaMemsetAsync(t.amax, 0, 16, st));
    const unsigned int* amax_x = fwd_amax ? reinterpret_cast<const unsigned int*>(fwd_amax) : t.amax;
    const unsigned int* amax_h = amax_x + 1;
    unsigned int *amax_dy = t.amax + 2, *amax_dp = t.amax + 3;
    if (!fwd_amax) {  // the forward did not hand its reductions over: redo them
      if ((rc = tcg::amax(x, B, E, E, t.amax, st))) return rc;
      if ((rc = tcg::amax(hidden, B, P, P, t.amax + 1, st))) return rc;
    }
    // LayerNorm + dropout backward: dz, dy, max |dy|, column partials of dgamma / dbeta / db_fc
    MC_DISPATCH_NV(P, (launch_ln_bwd<NV>(grad_out, z, mean, rstd, gamma, keep_mask, scale, B, P, t.dz, t.dy, t.partials,
                                         blocks, amax_dy, st, t.counters)));
    if (rc) return rc;
    dim3 blk(32, 32);
    colsum_tall_kernel<<<dim3((3 * P + 31) / 32, kColSlices), blk, 0, st>>>(t.partials, blocks, 3 * P, P, t.part2, t.counters,
                                                                          dgamma, dbeta, db_fc);
    MC_LAUNCH_CHECK();
    // dp = (dy Wf) * gelu'(projected) + dz : the GELU backward is the epilogue of the GEMM; it also reduces max |dp|
    // and the per-warp column sums of dp (db_proj)
    // both weights, transposed, in one launch (W_p^T is only needed for dx, but staging it costs nothing extra)
    // (the forward left max |Wp|, max |Wf| in words [4], [5] of fwd_amax: no reduction, no grid barrier here)
    if ((rc = tcg::stage_pair(tcg::StageJob{w_proj, P, E, 1, t.wpT}, tcg::StageJob{w_fc, P, P, 1, t.wfT}, t.amax + 8, st,
                              fwd_amax ? reinterpret_cast<const unsigned int*>(fwd_amax) + 4 : nullptr)))
      return rc;
    hg::RowArgs b1 = {};
    b1.A = t.dy; b1.lda = P; b1.a_amax = amax_dy; b1.W = t.wfT; b1.M = B; b1.K = P; b1.passes = passes;
    b1.epilogue = hg::kEpiGeluBwd; b1.in0 = projected; b1.in1 = t.dz; b1.out0 = t.dp; b1.out_amax = amax_dp;
    b1.colpart = t.colpart;
    if ((rc = hg::rows_gemm(b1, st))) return rc;
    colsum_tall_kernel<<<dim3((P + 31) / 32, kColSlices), blk, 0, st>>>(t.colpart, hg::colpart_rows(B), P, P, t.part2,
                                                                      t.counters + 32, db_proj, nullptr, nullptr);
    MC_LAUNCH_CHECK();
    // dWf = dy^T hidden, dWp = dp^T x : K is the batch; both operands are transposed inside the kernel
    static const int big_passes_env = getenv("MAE_CLIP_HEAD_BWD_PASSES") ? atoi(getenv("MAE_CLIP_HEAD_BWD_PASSES")) : 0;
    const int big_passes = (big_passes_env == 1 || big_passes_env == 3) ? big_passes_env : passes;
    hg::TtArgs g1{t.dy, P, amax_dy, hidden, P, amax_h, B, P, passes, dw_fc, P};
    if ((rc = hg::tt_gemm(g1, t.tt_ws, t.tt_ws_bytes, st))) return rc;
    hg::TtArgs g2{t.dp, P, amax_dp, x, E, amax_x, B, E, big_passes, dw_proj, E};
    if ((rc = hg::tt_gemm(g2, t.tt_ws, t.tt_ws_bytes, st))) return rc;
    if (dx) {
      // dx = dp Wp : short K, the converted row block of dp stays resident in tensor memory
      hg::AresArgs g3{t.dp, P, amax_dp, t.wpT, B, E, P, big_passes, dx, E};
      if ((rc = hg::ares_gemm(g3, st))) return rc;
    }
    return MC_OK;
  }
  if (use_tc(mode, B)) {
    MC_REQUIRE(aligned(ws, 256), MC_ERR_ALIGN, "proj_head_bwd: workspace must be 256-byte aligned");
    TcHeadWs t = tc_head_ws(ws, B, E, P);
    MC_REQUIRE(ws_bytes >= t.total, MC_ERR_WORKSPACE, "proj_head_bwd: workspace %zu < %zu", ws_bytes, t.total);
    MC_CUDA(cudaMemsetAsync(t.amax, 0, 16, st));
    const unsigned int* x_amax = fwd_amax ? reinterpret_cast<const unsigned int*>(fwd_amax) : nullptr;
    const unsigned int* h_amax = fwd_amax ? reinterpret_cast<const unsigned int*>(fwd_amax) + 1 : nullptr;
    MC_DISPATCH_NV(P, (launch_ln_bwd<NV>(grad_out, z, mean, rstd, gamma, keep_mask, scale, B, P, t.dz, t.dy,
                                         t.partials, blocks, t.amax, st)));
    if (rc) return rc;
    dim3 blk(32, 32);
    colsum3_kernel<<<(3 * P + 31) / 32, blk, 0, st>>>(t.partials, blocks, P, dgamma, dbeta, db_fc);
    MC_LAUNCH_CHECK();
    // dWf[n,k] = sum_m dy[m,n] hidden[m,k]  ->  (dy^T) . (hidden^T)^T, K = B (split-K)
    if ((rc = tcg::stage(t.dy, B, P, P, 1, t.dyT, st, t.amax))) return rc;
    if ((rc = tcg::stage(hidden, B, P, P, 1, t.hT, st, h_amax))) return rc;
    tcg::GemmOut o1{dw_fc, P, nullptr, nullptr};
    if ((rc = tcg::gemm(t.dyT, t.hT, P, P, B, o1, tcg::kEpiPlain, t.gemm_ws, t.gemm_ws_bytes, st))) return rc;
    // dh[m,k] = sum_n dy[m,n] Wf[n,k]  ->  dy . (Wf^T)^T
    if ((rc = tcg::stage(t.dy, B, P, P, 0, t.dy_p, st, t.amax))) return rc;
    if ((rc = tcg::stage(w_fc, P, P, P, 1, t.wfT, st))) return rc;
    tcg::GemmOut o2{t.dh, P, nullptr, nullptr};
    if ((rc = tcg::gemm(t.dy_p, t.wfT, B, P, P, o2, tcg::kEpiPlain, t.gemm_ws, t.gemm_ws_bytes, st))) return rc;
    {
      size_t n4 = (size_t)B * P / 4;
      int nb = (int)((n4 + 255) / 256);
      int cap = num_sms() * 8;
      if (nb > cap) nb = cap;
      gelu_bwd_add_kernel<<<nb, 256, 0, st>>>(t.dh, projected, t.dz, n4, t.amax + 1);  // dp = dh * gelu'(p) + dz
      MC_LAUNCH_CHECK();
      if ((rc = colsum_rows(t.dh, B, P, db_proj, t.partials, 2 * blocks, st))) return rc;
    }
    // dWp[n,e] = sum_m dp[m,n] x[m,e]  ->  (dp^T) . (x^T)^T, K = B (split-K)
    if ((rc = tcg::stage(t.dh, B, P, P, 1, t.dpT, st, t.amax + 1))) return rc;
    if ((rc = tcg::stage(x, B, E, E, 1, t.xT, st, x_amax))) return rc;
    tcg::GemmOut o3{dw_proj, E, nullptr, nullptr};
    if ((rc = tcg::gemm(t.dpT, t.xT, P, E, B, o3, tcg::kEpiPlain, t.gemm_ws, t.gemm_ws_bytes, st))) return rc;
    if (dx) {
      // dx[m,e] = sum_n dp[m,n] Wp[n,e]  ->  dp . (Wp^T)^T
      if ((rc = tcg::stage(t.dh, B, P, P, 0, t.dp_p, st, t.amax + 1))) return rc;
      if ((rc = tcg::stage(w_proj, P, E, E, 1, t.wpT, st))) return rc;
      tcg::GemmOut o4{dx, E, nullptr, nullptr};
      if ((rc = tcg::gemm(t.dp_p, t.wpT, B, E, P, o4, tcg::kEpiPlain, t.gemm_ws, t.gemm_ws_bytes, st))) return rc;
    }
    return MC_OK;
  }
  HeadWs w = head_ws(ws, B, E, P);
  MC_REQUIRE(ws_bytes >= w.total, MC_ERR_WORKSPACE, "proj_head_bwd: workspace %zu < %zu", ws_bytes,
             w.total);
  MC_DISPATCH_NV(P, (launch_ln_bwd<NV>(grad_out, z, mean, rstd, gamma, keep_mask, scale, B, P, w.dz,
                                       w.dy, w.partials, blocks, nullptr, st)));
  if (rc) return rc;
  // dgamma / dbeta: column sums of the block partials ([blocks][2P] viewed as rows x 2P)
  {
    dim3 blk(32, 32);
    // partial rows are [dgamma(P) | dbeta(P) | db_fc(P)]
    colsum3_kernel<<<(3 * P + 31) / 32, blk, 0, st>>>(w.partials, blocks, P, dgamma, dbeta, db_fc);
    MC_LAUNCH_CHECK();
  }
  // dWf[n,k] = sum_m dy[m,n] hidden[m,k]
  SgemmArgs g1{w.dy, 1, P, hidden, P, 1, dw_fc, P, P, P, B, 1.f, nullptr, nullptr, 0};
  if ((rc = sgemm(g1, st))) return rc;
  // dh[m,k] = sum_n dy[m,n] Wf[n,k]
  SgemmArgs g2{w.dy, P, 1, w_fc, P, 1, w.dh, P, B, P, P, 1.f, nullptr, nullptr, 0};
  if ((rc = sgemm(g2, st))) return rc;
  {
    size_t n4 = (size_t)B * P / 4;
    int nb = (int)((n4 + 255) / 256);
    int cap = num_sms() * 8;
    if (nb > cap) nb = cap;
    gelu_bwd_add_kernel<<<nb, 256, 0, st>>>(w.dh, projected, w.dz, n4, nullptr);
    MC_LAUNCH_CHECK();
    if ((rc = colsum_rows(w.dh, B, P, db_proj, w.partials, 2 * blocks, st))) return rc;
  }
  // dWp[n,e] = sum_m dp[m,n] x[m,e]
  SgemmArgs g3{w.dh, 1, P, x, E, 1, dw_proj, E, P, E, B, 1.f, nullptr, nullptr, 0};
  if ((rc = sgemm(g3, st))) return rc;
  if (dx) {
    // dx[m,e] = sum_n dp[m,n] Wp[n,e]
    SgemmArgs g4{w.dh, P, 1, w_proj, E, 1, dx, E, B, E, P, 1.f, nullptr, nullptr, 0};
    if ((rc = sgemm(g4, st))) return rc;
  }
  return MC_OK;
}

}  // extern "C"
