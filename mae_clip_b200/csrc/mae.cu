// M1-M3: MAE per-sample random masking and normalised-pixel masked-patch MSE.
// NOT IN THE REFERENCE (/root/reference holds no MAE code - SURVEY.md section 0.2); the contract
// is BASELINE.json's north_star as restated by oracle/mae_ref.py.  All kernels are HBM-bound
// byte movers: coalesced 16-byte rows, no tensor cores.
#include <cuda_fp16.h>

#include "common.cuh"

namespace mc {

// ------------------------------------------------------------------------------------------
// M1a: per-row stable ascending argsort of the noise -> ids_restore, mask, ids_keep.
// One block per sample; bitonic sort of 64-bit keys (order-preserving float bits << 32 | index)
// in shared memory.  The index in the low word makes ties resolve to the lower index, which is
// exactly torch.argsort(stable=True) (SURVEY.md section 7 hard part e).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t sortable_bits(float f) {
  if (f != f) return 0xFFFFFFFFu;          // NaN sorts last, like torch
  if (f == 0.f) return 0x80000000u;        // -0.0 == +0.0
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void argsort_mask_kernel(const float* __restrict__ noise, int L, int LP, int len_keep,
                                    float* __restrict__ mask, int64_t* __restrict__ ids_restore,
                                    int64_t* __restrict__ ids_keep) {
  extern __shared__ unsigned long long keys[];
  const int n = blockIdx.x;
  const float* row = noise + (size_t)n * L;
  for (int i = threadIdx.x; i < LP; i += blockDim.x)
    keys[i] = (i < L) ? (((unsigned long long)sortable_bits(row[i]) << 32) | (unsigned)i)
                      : 0xFFFFFFFFFFFFFFFFull;
  __syncthreads();
  for (int k = 2; k <= LP; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (LP >> 1); t += blockDim.x) {
        const int i = ((t / j) * 2 * j) + (t % j);
        const int l = i + j;
        const bool asc = ((i & k) == 0);
        unsigned long long a = keys[i], b = keys[l];
        if ((a > b) == asc) { keys[i] = b; keys[l] = a; }
      }
      __syncthreads();
    }
  }
  for (int pos = threadIdx.x; pos < L; pos += blockDim.x) {
    const int idx = (int)(keys[pos] & 0xFFFFFFFFull);
    ids_restore[(size_t)n * L + idx] = pos;
    mask[(size_t)n * L + idx] = pos < len_keep ? 0.f : 1.f;
    if (ids_keep && pos < len_keep) ids_keep[(size_t)n * len_keep + pos] = idx;
  }
}

// ------------------------------------------------------------------------------------------
// M1 fused (L <= 1024): one block per sample does the whole op.  The stable argsort is a RANK
// sort - position of element i = number of 64-bit keys smaller than key i (keys are unique: the
// index sits in the low word) - i.e. one pass of broadcast shared-memory reads per element and a
// single barrier instead of the 36 barrier-separated stages of a 256-key bitonic network; the rank
// IS ids_restore[i], so nothing is scattered.  The block then copies its len_keep kept rows with
// 16-byte vectors, all of a row's loads in flight at once; blocks finish sorting at different
// times, so the latency-bound sort of one sample overlaps the HBM-bound copy of another.
// ------------------------------------------------------------------------------------------
template <typename V, int kMaxVecPerLane>
__global__ void __launch_bounds__(256) random_masking_fused_kernel(const float* __restrict__ noise, int L, int len_keep,
                                                                   float* __restrict__ mask,
                                                                   int64_t* __restrict__ ids_restore,
                                                                   int64_t* __restrict__ ids_keep,
                                                                   const V* __restrict__ src, int row_vecs,
                                                                   V* __restrict__ out) {
  extern __shared__ unsigned long long keys[];           // [L] keys, then [len_keep] kept indices
  int* keep_idx = reinterpret_cast<int*>(keys + L);
  const int n = blockIdx.x;
  const float* row = noise + (size_t)n * L;
  for (int i = threadIdx.x; i < L; i += blockDim.x)
    keys[i] = ((unsigned long long)sortable_bits(row[i]) << 32) | (unsigned)i;
  __syncthreads();
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    const unsigned long long mine = keys[i];
    int rank = 0;
    int j = 0;
    for (; j + 4 <= L; j += 4)
      rank += (keys[j] < mine) + (keys[j + 1] < mine) + (keys[j + 2] < mine) + (keys[j + 3] < mine);
    for (; j < L; ++j) rank += keys[j] < mine;
    ids_restore[(size_t)n * L + i] = rank;
    mask[(size_t)n * L + i] = rank < len_keep ? 0.f : 1.f;
    if (rank < len_keep) {
      keep_idx[rank] = i;
      if (ids_keep) ids_keep[(size_t)n * len_keep + rank] = i;
    }
  }
  if (src == nullptr) return;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < len_keep; r += 8) {
    const V* sp = src + ((size_t)n * L + keep_idx[r]) * row_vecs;
    V* dp = out + ((size_t)n * len_keep + r) * row_vecs;
    if (row_vecs <= 32 * kMaxVecPerLane) {
      V v[kMaxVecPerLane];
#pragma unroll
      for (int u = 0; u < kMaxVecPerLane; ++u)
        if (lane + 32 * u < row_vecs) v[u] = sp[lane + 32 * u];
#pragma unroll
      for (int u = 0; u < kMaxVecPerLane; ++u)
        if (lane + 32 * u < row_vecs) dp[lane + 32 * u] = v[u];
    } else {
      for (int c = lane; c < row_vecs; c += 32) dp[c] = sp[c];
    }
  }
}

// ------------------------------------------------------------------------------------------
// M1b: row gather with a default.  Output row (n, r): idx = index[n*out_rows + r]; when
// idx < limit copy src row (n, idx), else fill (zeros or a broadcast token).  One warp per row,
// VEC-byte vectors.  Used for the keep-gather, its backward (scatter expressed as a gather through
// ids_restore) and the decoder-side un-shuffle.
// ------------------------------------------------------------------------------------------
template <typename V>
__global__ void __launch_bounds__(256) gather_rows_kernel(const V* __restrict__ src,
                                                          int src_rows, const int64_t* __restrict__ index,
                                                          int out_rows, long long total_rows,
                                                          int row_vecs, int limit,
                                                          const V* __restrict__ fill,
                                                          V* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < total_rows;
       row += warps) {
    const long long n = row / out_rows;
    const long long idx = index[row];
    V* dst = out + row * row_vecs;
    if (idx < limit) {
      const V* s = src + (n * src_rows + idx) * row_vecs;
      for (int v = lane; v < row_vecs; v += 32) dst[v] = s[v];
    } else if (fill) {
      for (int v = lane; v < row_vecs; v += 32) dst[v] = fill[v];
    } else {
      V zero;
      memset(&zero, 0, sizeof(V));
      for (int v = lane; v < row_vecs; v += 32) dst[v] = zero;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Backward of the decoder-side un-shuffle (restore_tokens): every row (n, l) of grad_out goes either
// to d x_kept[n, ids_restore[n, l]] (a copy: targets are unique) or into the column sum that is the
// mask token's gradient.  A block walks its rows four at a time (four 16-byte loads in flight per
// thread), each thread owning fixed columns; block partials are merged in a fixed order by a
// second kernel (deterministic, no atomics).
// ------------------------------------------------------------------------------------------
template <typename E> struct Vec16;  // 16 bytes of E <-> floats
template <> struct Vec16<float> {
  static constexpr int kN = 4;
  __device__ static void unpack(const uint4& u, float* f) {
    f[0] = __uint_as_float(u.x); f[1] = __uint_as_float(u.y); f[2] = __uint_as_float(u.z); f[3] = __uint_as_float(u.w);
  }
};
template <> struct Vec16<__nv_bfloat16> {
  static constexpr int kN = 8;
  __device__ static void unpack(const uint4& u, float* f) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
  }
};
template <> struct Vec16<__half> {
  static constexpr int kN = 8;
  __device__ static void unpack(const uint4& u, float* f) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
      f[2 * i] = __low2float(h);
      f[2 * i + 1] = __high2float(h);
    }
  }
};

constexpr int kRestoreVecsPerThread = 4;  // rows up to 256 * 4 * 16 bytes = 16 KB

template <typename E>
__global__ void __launch_bounds__(256) restore_bwd_kernel(const uint4* __restrict__ gout,
                                                          const int64_t* __restrict__ ids_restore, long long rows, int L,
                                                          int row_vecs, int len_keep, uint4* __restrict__ dx_kept,
                                                          float* __restrict__ partial /* [gridDim.x][row_vecs * kN] */) {
  constexpr int kN = Vec16<E>::kN;
  float acc[kRestoreVecsPerThread][kN];
#pragma unroll
  for (int a = 0; a < kRestoreVecsPerThread; ++a)
#pragma unroll
    for (int e = 0; e < kN; ++e) acc[a][e] = 0.f;
  const long long per = (rows + gridDim.x - 1) / gridDim.x;
  const long long beg = (long long)blockIdx.x * per, end = min(rows, beg + per);
  for (long long r0 = beg; r0 < end; r0 += 4) {
    long long dstrow[4];
    uint4 v[4][kRestoreVecsPerThread];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long row = r0 + u;
      dstrow[u] = -2;  // out of range
      if (row < end) {
        const long long idx = ids_restore[row];
        dstrow[u] = idx < len_keep ? (row / L) * len_keep + idx : -1;
#pragma unroll
        for (int a = 0; a < kRestoreVecsPerThread; ++a) {
          const int c = threadIdx.x + 256 * a;
          if (c < row_vecs) v[u][a] = ld_stream(gout + row * row_vecs + c);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (dstrow[u] == -2) continue;
#pragma unroll
      for (int a = 0; a < kRestoreVecsPerThread; ++a) {
        const int c = threadIdx.x + 256 * a;
        if (c >= row_vecs) continue;
        if (dstrow[u] >= 0) {
          dx_kept[dstrow[u] * row_vecs + c] = v[u][a];
        } else {
          float f[kN];
          Vec16<E>::unpack(v[u][a], f);
#pragma unroll
          for (int e = 0; e < kN; ++e) acc[a][e] += f[e];
        }
      }
    }
  }
#pragma unroll
  for (int a = 0; a < kRestoreVecsPerThread; ++a) {
    const int c = threadIdx.x + 256 * a;
    if (c >= row_vecs) continue;
#pragma unroll
    for (int e = 0; e < kN; ++e) partial[(size_t)blockIdx.x * row_vecs * kN + (size_t)c * kN + e] = acc[a][e];
  }
}

template <typename E>
__global__ void __launch_bounds__(256) restore_bwd_finalize_kernel(const float* __restrict__ partial, int nblocks, int Dm,
                                                                   E* __restrict__ dtoken) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= Dm) return;
  float s = 0.f;
  for (int b = 0; b < nblocks; ++b) s += partial[(size_t)b * Dm + d];
  if constexpr (sizeof(E) == 4) dtoken[d] = s;
  else dtoken[d] = static_cast<E>(s);
}

static int restore_bwd_grid(long long rows) {
  long long nb = (rows + 15) / 16;
  const long long cap = (long long)num_sms() * 4;
  return (int)(nb < cap ? nb : cap);
}

static int launch_gather(const void* src, int src_rows, const int64_t* index, int out_rows, int N,
                         size_t row_bytes, int limit, const void* fill, void* out, cudaStream_t st) {
  const long long total = (long long)N * out_rows;
  if (total == 0 || row_bytes == 0) return MC_OK;
  long long nb = (total + 7) / 8;
  const long long cap = (long long)num_sms() * 16;
  if (nb > cap) nb = cap;
  const bool a16 = row_bytes % 16 == 0 && aligned(src, 16) && aligned(out, 16) && (!fill || aligned(fill, 16));
  const bool a4 = row_bytes % 4 == 0 && aligned(src, 4) && aligned(out, 4) && (!fill || aligned(fill, 4));
  if (a16)
    gather_rows_kernel<uint4><<<(int)nb, 256, 0, st>>>((const uint4*)src, src_rows, index, out_rows,
                                                       total, (int)(row_bytes / 16), limit,
                                                       (const uint4*)fill, (uint4*)out);
  else if (a4)
    gather_rows_kernel<uint32_t><<<(int)nb, 256, 0, st>>>((const uint32_t*)src, src_rows, index,
                                                          out_rows, total, (int)(row_bytes / 4), limit,
                                                          (const uint32_t*)fill, (uint32_t*)out);
  else
    gather_rows_kernel<uint16_t><<<(int)nb, 256, 0, st>>>((const uint16_t*)src, src_rows, index,
                                                          out_rows, total, (int)(row_bytes / 2), limit,
                                                          (const uint16_t*)fill, (uint16_t*)out);
  MC_LAUNCH_CHECK();
  return MC_OK;
}

// ------------------------------------------------------------------------------------------
// M2+M3: fused patchify + norm-pix target + masked MSE.  One warp per patch; unmasked patches are
// skipped entirely (their bytes are never read).  The patch's pixels are read three times
// (mean, variance, difference) but only the first pass reaches HBM: 3 KB per warp stays in L1.
// Element order inside a patch is (ph, pw, c): e = (ph*p + pw)*3 + c  <->  imgs[n][c][h*p+ph][w*p+pw].
// ------------------------------------------------------------------------------------------
struct PatchGeom {
  int H, W, p, gw, L, PE;  // gw = W/p patches per row, L = patches per image, PE = p*p*3
};

__device__ __forceinline__ float ld_pred(const float* p, size_t i) { return p[i]; }
__device__ __forceinline__ float ld_pred(const __nv_bfloat16* p, size_t i) {
  return __bfloat162float(p[i]);
}
__device__ __forceinline__ void st_pred(float* p, size_t i, float v) { p[i] = v; }
__device__ __forceinline__ void st_pred(__nv_bfloat16* p, size_t i, float v) {
  p[i] = __float2bfloat16_rn(v);
}

__device__ __forceinline__ const float* patch_base(const float* imgs, const PatchGeom& g, long long n,
                                                   int l) {
  const int h = l / g.gw, w = l % g.gw;
  return imgs + ((size_t)n * 3 * g.H + (size_t)h * g.p) * g.W + (size_t)w * g.p;
}
// offset of element e (patch order) from patch_base
__device__ __forceinline__ size_t elem_off(const PatchGeom& g, int e) {
  const int c = e % 3, pix = e / 3, ph = pix / g.p, pw = pix % g.p;
  return ((size_t)c * g.H + ph) * g.W + pw;
}

__device__ __forceinline__ void patch_stats(const float* base, const PatchGeom& g, int lane,
                                            int norm_pix, float& mean, float& rstd) {
  mean = 0.f;
  rstd = 1.f;
  if (!norm_pix) return;
  float s = 0.f;
  for (int e = lane; e < g.PE; e += 32) s += base[elem_off(g, e)];
  mean = warp_sum(s) / (float)g.PE;
  float v = 0.f;
  for (int e = lane; e < g.PE; e += 32) {
    float d = base[elem_off(g, e)] - mean;
    v = fmaf(d, d, v);
  }
  rstd = rsqrtf(warp_sum(v) / (float)(g.PE - 1) + 1e-6f);
}


// ---- p = 16 fast path -------------------------------------------------------------------------
// A 16 x 16 x 3 patch is 48 image rows of 64 contiguous bytes.  The warp reads them with 16-byte
// loads (4 lanes per row, 8 rows per instruction), keeps them in registers for the mean / variance,
// and parks the target in a per-warp shared-memory tile [c][pixel] (plane stride 268 floats: the
// (e % 3, e / 3) read pattern of `pred` order is then bank-conflict free).  `pred` / `dpred` move as
// 16-byte (fp32) or 8-byte (bf16) vectors, four consecutive elements per lane.
constexpr int kPlane16 = 268;
constexpr int kTile16 = 3 * kPlane16;

__device__ __forceinline__ void stage_patch16(const float* __restrict__ base, int H, int W, int lane,
                                              int norm_pix, float* __restrict__ tile) {
  float4 v[6];
  float s = 0.f;
  const int quad = lane & 3;
#pragma unroll
  for (int it = 0; it < 6; ++it) {
    const int row = it * 8 + (lane >> 2), c = row >> 4, ph = row & 15;
    v[it] = ld_stream(reinterpret_cast<const float4*>(base + ((size_t)c * H + ph) * W + quad * 4));
    s += (v[it].x + v[it].y) + (v[it].z + v[it].w);
  }
  float mean = 0.f, rstd = 1.f;
  if (norm_pix) {
    mean = warp_sum(s) * (1.f / 768.f);
    float q = 0.f;
#pragma unroll
    for (int it = 0; it < 6; ++it) {
      const float a = v[it].x - mean, b = v[it].y - mean, c2 = v[it].z - mean, d = v[it].w - mean;
      q += (a * a + b * b) + (c2 * c2 + d * d);
    }
    rstd = rsqrtf(warp_sum(q) * (1.f / 767.f) + 1e-6f);
  }
  __syncwarp();  // the previous patch's readers are done with the tile
#pragma unroll
  for (int it = 0; it < 6; ++it) {
    const int row = it * 8 + (lane >> 2), c = row >> 4, ph = row & 15;
    float4 t;
    t.x = (v[it].x - mean) * rstd; t.y = (v[it].y - mean) * rstd;
    t.z = (v[it].z - mean) * rstd; t.w = (v[it].w - mean) * rstd;
    *reinterpret_cast<float4*>(tile + c * kPlane16 + ph * 16 + quad * 4) = t;
  }
  __syncwarp();
}
__device__ __forceinline__ float tile16_at(const float* tile, int e) {  // e = (ph*16 + pw)*3 + c
  const int pix = e / 3, c = e - 3 * pix;
  return tile[c * kPlane16 + pix];
}
__device__ __forceinline__ float4 ld_pred4(const float* p, size_t i) {
  return ld_stream(reinterpret_cast<const float4*>(p + i));
}
__device__ __forceinline__ float4 ld_pred4(const __nv_bfloat16* p, size_t i) {
  const uint2 u = *reinterpret_cast<const uint2*>(p + i);
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
  const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
  return make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
}
__device__ __forceinline__ void st_pred4(float* p, size_t i, float4 v) {
  st_stream(reinterpret_cast<float4*>(p + i), v);
}
__device__ __forceinline__ void st_pred4(__nv_bfloat16* p, size_t i, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p + i) = u;
}

template <typename PT, bool FAST16>
__global__ void __launch_bounds__(256, 4) masked_mse_fwd_kernel(const PT* __restrict__ pred,
                                                             const float* __restrict__ imgs,
                                                             const float* __restrict__ mask,
                                                             long long patches, PatchGeom g,
                                                             int norm_pix, float* __restrict__ part,
                                                             unsigned int* __restrict__ counter,
                                                             float* __restrict__ loss_out,
                                                             float* __restrict__ mask_sum_out) {
  __shared__ float sm_l[8], sm_m[8];
  __shared__ bool is_last;
  __shared__ __align__(16) float tiles[FAST16 ? 8 * kTile16 : 4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  float acc = 0.f, macc = 0.f;
  for (long long pt = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; pt < patches;
       pt += warps) {
    const float m = mask[pt];
    if (m == 0.f) continue;
    const long long n = pt / g.L;
    const int l = (int)(pt % g.L);
    const float* base = patch_base(imgs, g, n, l);
    float se = 0.f;
    const size_t poff = (size_t)pt * g.PE;
    if (FAST16) {
      float* tile = tiles + warp * kTile16;
      stage_patch16(base, g.H, g.W, lane, norm_pix, tile);
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const int e0 = 4 * lane + 128 * k;
        const float4 pv = ld_pred4(pred, poff + e0);
        const float d0 = pv.x - tile16_at(tile, e0), d1 = pv.y - tile16_at(tile, e0 + 1);
        const float d2 = pv.z - tile16_at(tile, e0 + 2), d3 = pv.w - tile16_at(tile, e0 + 3);
        se += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
      }
    } else {
      float mean, rstd;
      patch_stats(base, g, lane, norm_pix, mean, rstd);
      for (int e = lane; e < g.PE; e += 32) {
        float t = (base[elem_off(g, e)] - mean) * rstd;
        float d = ld_pred(pred, poff + e) - t;
        se = fmaf(d, d, se);
      }
    }
    se = warp_sum(se);
    acc += m * se / (float)g.PE;
    macc += m;
  }
  if (lane == 0) { sm_l[warp] = acc; sm_m[warp] = macc; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < 8; ++w) { a += sm_l[w]; b += sm_m[w]; }
    part[blockIdx.x] = a;
    part[gridDim.x + blockIdx.x] = b;
    __threadfence();
    unsigned int done = atomicAdd(counter, 1u);
    is_last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {  // deterministic final reduction by the last block to finish
    __threadfence();
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
      a += (double)((volatile float*)part)[i];
      b += (double)((volatile float*)part)[gridDim.x + i];
    }
    a = warp_sum(a);
    b = warp_sum(b);
    __shared__ double da[8], db[8];
    if (lane == 0) { da[warp] = a; db[warp] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double sa = 0.0, sb = 0.0;
      for (int w = 0; w < 8; ++w) { sa += da[w]; sb += db[w]; }
      *loss_out = (float)(sa / sb);
      if (mask_sum_out) *mask_sum_out = (float)sb;
      *counter = 0;
    }
  }
}

template <typename PT, bool FAST16>
__global__ void __launch_bounds__(256, 4) masked_mse_bwd_kernel(const PT* __restrict__ pred,
                                                             const float* __restrict__ imgs,
                                                             const float* __restrict__ mask,
                                                             long long patches, PatchGeom g,
                                                             int norm_pix,
                                                             const float* __restrict__ mask_sum,
                                                             const float* __restrict__ grad_loss,
                                                             PT* __restrict__ dpred) {
  __shared__ __align__(16) float tiles[FAST16 ? 8 * kTile16 : 4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const float gl = grad_loss ? *grad_loss : 1.f;
  const float coef = 2.f * gl / ((float)g.PE * (*mask_sum));
  for (long long pt = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; pt < patches;
       pt += warps) {
    const float m = mask[pt];
    const size_t poff = (size_t)pt * g.PE;
    if (m == 0.f) {
      if (FAST16) {
#pragma unroll
        for (int k = 0; k < 6; ++k) st_pred4(dpred, poff + 4 * lane + 128 * k, make_float4(0.f, 0.f, 0.f, 0.f));
      } else {
        for (int e = lane; e < g.PE; e += 32) st_pred(dpred, poff + e, 0.f);
      }
      continue;
    }
    const long long n = pt / g.L;
    const int l = (int)(pt % g.L);
    const float* base = patch_base(imgs, g, n, l);
    const float cm = coef * m;
    if (FAST16) {
      float* tile = tiles + warp * kTile16;
      stage_patch16(base, g.H, g.W, lane, norm_pix, tile);
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const int e0 = 4 * lane + 128 * k;
        const float4 pv = ld_pred4(pred, poff + e0);
        float4 o;
        o.x = cm * (pv.x - tile16_at(tile, e0)); o.y = cm * (pv.y - tile16_at(tile, e0 + 1));
        o.z = cm * (pv.z - tile16_at(tile, e0 + 2)); o.w = cm * (pv.w - tile16_at(tile, e0 + 3));
        st_pred4(dpred, poff + e0, o);
      }
    } else {
      float mean, rstd;
      patch_stats(base, g, lane, norm_pix, mean, rstd);
      for (int e = lane; e < g.PE; e += 32) {
        float t = (base[elem_off(g, e)] - mean) * rstd;
        st_pred(dpred, poff + e, cm * (ld_pred(pred, poff + e) - t));
      }
    }
  }
}

__global__ void __launch_bounds__(256) patchify_kernel(const float* __restrict__ imgs,
                                                       long long patches, PatchGeom g, int norm_pix,
                                                       float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long pt = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; pt < patches;
       pt += warps) {
    const long long n = pt / g.L;
    const int l = (int)(pt % g.L);
    const float* base = patch_base(imgs, g, n, l);
    float mean, rstd;
    patch_stats(base, g, lane, norm_pix, mean, rstd);
    for (int e = lane; e < g.PE; e += 32)
      out[(size_t)pt * g.PE + e] = (base[elem_off(g, e)] - mean) * rstd;
  }
}

static int make_geom(const char* who, int N, int H, int W, int p, PatchGeom& g) {
  MC_REQUIRE(N > 0 && H > 0 && W > 0 && p > 0, MC_ERR_BAD_ARG, "%s: bad sizes N=%d H=%d W=%d p=%d", who,
             N, H, W, p);
  MC_REQUIRE(H % p == 0 && W % p == 0, MC_ERR_BAD_ARG, "%s: image %dx%d not divisible by patch %d",
             who, H, W, p);
  g.H = H; g.W = W; g.p = p; g.gw = W / p; g.L = (H / p) * (W / p); g.PE = p * p * 3;
  MC_REQUIRE(g.PE > 1, MC_ERR_BAD_ARG, "%s: patch too small for an unbiased variance", who);
  return MC_OK;
}

static int mse_grid(long long patches) {
  long long nb = (patches + 7) / 8;
  const long long cap = (long long)num_sms() * 8;
  return (int)(nb < cap ? nb : cap);
}

}  // namespace mc

using namespace mc;

extern "C" {

int mc_random_masking(const void* x, int elem_size, const float* noise, int N, int L, int Dm,
                      int len_keep, void* x_masked, float* mask, int64_t* ids_restore,
                      int64_t* ids_keep, void* stream) {
  MC_ARCH_GUARD();
  MC_REQUIRE(noise && mask && ids_restore, MC_ERR_BAD_ARG, "random_masking: null pointer");
  MC_REQUIRE(N >= 0 && L > 0 && len_keep >= 0 && len_keep <= L, MC_ERR_BAD_ARG,
             "random_masking: bad sizes N=%d L=%d len_keep=%d", N, L, len_keep);
  MC_REQUIRE(L <= 4096, MC_ERR_UNSUPPORTED, "random_masking: L=%d > 4096 patches", L);
  MC_REQUIRE((x == nullptr) == (x_masked == nullptr), MC_ERR_BAD_ARG,
             "random_masking: pass both x and x_masked or neither");
  if (N == 0) return MC_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int LP = 2;
  while (LP < L) LP <<= 1;
  int threads = LP / 2;
  if (threads < 32) threads = 32;
  if (threads > 512) threads = 512;
  const bool need_keep = (x != nullptr) || ids_keep;
  MC_REQUIRE(!x || ids_keep, MC_ERR_BAD_ARG, "random_masking: ids_keep buffer is required with x");
  if (x) {
    MC_REQUIRE(elem_size == 2 || elem_size == 4, MC_ERR_BAD_ARG, "random_masking: elem_size %d", elem_size);
    MC_REQUIRE(Dm > 0, MC_ERR_BAD_ARG, "random_masking: Dm=%d", Dm);
  }
  const size_t row_bytes = x ? (size_t)Dm * elem_size : 0;
  if (L <= 1024 && (!x || (row_bytes % 16 == 0 && aligned(x, 16) && aligned(x_masked, 16)))) {
    // fused rank-sort + keep-gather: one launch, one block per sample
    const size_t smem = (size_t)L * 8 + (size_t)(len_keep > 0 ? len_keep : 1) * 4;
    random_masking_fused_kernel<uint4, 8><<<N, 256, smem, st>>>(
        noise, L, len_keep, mask, ids_restore, need_keep ? ids_keep : nullptr,
        (x && len_keep > 0) ? static_cast<const uint4*>(x) : nullptr, (int)(row_bytes / 16), static_cast<uint4*>(x_masked));
    MC_LAUNCH_CHECK();
    return MC_OK;
  }
  argsort_mask_kernel<<<N, threads, (size_t)LP * 8, st>>>(noise, L, LP, len_keep, mask, ids_restore,
                                                          need_keep ? ids_keep : nullptr);
  MC_LAUNCH_CHECK();
  if (x && len_keep > 0) {
    MC_REQUIRE(elem_size == 2 || elem_size == 4, MC_ERR_BAD_ARG, "random_masking: elem_size %d",
               elem_size);
    MC_REQUIRE(Dm > 0, MC_ERR_BAD_ARG, "random_masking: Dm=%d", Dm);
    return launch_gather(x, L, ids_keep, len_keep, N, (size_t)Dm * elem_size, L, nullptr, x_masked, st);
  }
  return MC_OK;
}

int mc_random_masking_bwd(const void* grad_x_masked, int elem_size, const float* /*mask*/,
                          const int64_t* ids_restore, int N, int L, int Dm, int len_keep,
                          void* grad_x, void* stream) {
  MC_ARCH_GUARD();
  MC_REQUIRE(grad_x_masked && ids_restore && grad_x, MC_ERR_BAD_ARG, "random_masking_bwd: null pointer");
  MC_REQUIRE(N >= 0 && L > 0 && Dm > 0 && len_keep >= 0 && len_keep <= L, MC_ERR_BAD_ARG,
             "random_masking_bwd: bad sizes");
  MC_REQUIRE(elem_size == 2 || elem_size == 4, MC_ERR_BAD_ARG, "random_masking_bwd: elem_size %d",
             elem_size);
  return launch_gather(grad_x_masked, len_keep, ids_restore, L, N, (size_t)Dm * elem_size, len_keep,
                       nullptr, grad_x, static_cast<cudaStream_t>(stream));
}

int mc_restore_tokens(const void* x_kept, int elem_size, const void* mask_token,
                      const int64_t* ids_restore, int N, int L, int Dm, int len_keep, void* out,
                      void* stream) {
  MC_ARCH_GUARD();
  MC_REQUIRE(x_kept && mask_token && ids_restore && out, MC_ERR_BAD_ARG, "restore_tokens: null pointer");
  MC_REQUIRE(N >= 0 && L > 0 && Dm > 0 && len_keep >= 0 && len_keep <= L, MC_ERR_BAD_ARG,
             "restore_tokens: bad sizes");
  MC_REQUIRE(elem_size == 2 || elem_size == 4, MC_ERR_BAD_ARG, "restore_tokens: elem_size %d", elem_size);
  return launch_gather(x_kept, len_keep, ids_restore, L, N, (size_t)Dm * elem_size, len_keep,
                       mask_token, out, static_cast<cudaStream_t>(stream));
}

size_t mc_restore_tokens_bwd_workspace_bytes(int N, int L, int Dm) {
  if (N <= 0 || L <= 0 || Dm <= 0) return 0;
  return (size_t)restore_bwd_grid((long long)N * L) * Dm * sizeof(float);
}

int mc_restore_tokens_bwd(const void* grad_out, int elem_size, int is_bf16, const int64_t* ids_restore, int N, int L,
                          int Dm, int len_keep, void* dx_kept, void* dmask_token, void* ws, size_t ws_bytes,
                          void* stream) {
  MC_ARCH_GUARD();
  MC_REQUIRE(grad_out && ids_restore && dmask_token && ws && (dx_kept || len_keep == 0), MC_ERR_BAD_ARG,
             "restore_tokens_bwd: null pointer");
  MC_REQUIRE(N > 0 && L > 0 && Dm > 0 && len_keep >= 0 && len_keep <= L, MC_ERR_BAD_ARG, "restore_tokens_bwd: bad sizes");
  MC_REQUIRE(elem_size == 2 || elem_size == 4, MC_ERR_BAD_ARG, "restore_tokens_bwd: elem_size %d", elem_size);
  const size_t row_bytes = (size_t)Dm * elem_size;
  MC_REQUIRE(row_bytes % 16 == 0 && row_bytes <= (size_t)256 * kRestoreVecsPerThread * 16 && aligned(grad_out, 16) &&
                 (!dx_kept || aligned(dx_kept, 16)),
             MC_ERR_UNSUPPORTED, "restore_tokens_bwd: rows must be 16-byte multiples of at most 16 KB, 16-byte aligned");
  MC_REQUIRE(ws_bytes >= mc_restore_tokens_bwd_workspace_bytes(N, L, Dm), MC_ERR_WORKSPACE,
             "restore_tokens_bwd: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long rows = (long long)N * L;
  const int grid = restore_bwd_grid(rows), row_vecs = (int)(row_bytes / 16);
  float* partial = static_cast<float*>(ws);
  const uint4* go = static_cast<const uint4*>(grad_out);
  uint4* dx = static_cast<uint4*>(dx_kept);
  const int fb = (Dm + 255) / 256;
  if (elem_size == 4) {
    restore_bwd_kernel<float><<<grid, 256, 0, st>>>(go, ids_restore, rows, L, row_vecs, len_keep, dx, partial);
    MC_LAUNCH_CHECK();
    restore_bwd_finalize_kernel<float><<<fb, 256, 0, st>>>(partial, grid, Dm, static_cast<float*>(dmask_token));
  } else if (is_bf16) {
    restore_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(go, ids_restore, rows, L, row_vecs, len_keep, dx, partial);
    MC_LAUNCH_CHECK();
    restore_bwd_finalize_kernel<__nv_bfloat16><<<fb, 256, 0, st>>>(partial, grid, Dm, static_cast<__nv_bfloat16*>(dmask_token));
  } else {
    restore_bwd_kernel<__half><<<grid, 256, 0, st>>>(go, ids_restore, rows, L, row_vecs, len_keep, dx, partial);
    MC_LAUNCH_CHECK();
    restore_bwd_finalize_kernel<__half><<<fb, 256, 0, st>>>(partial, grid, Dm, static_cast<__half*>(dmask_token));
  }
  MC_LAUNCH_CHECK();
  return MC_OK;
}

size_t mc_masked_mse_workspace_bytes(int N, int L) {
  if (N <= 0 || L <= 0) return 0;
  return 256 + (size_t)2 * mse_grid((long long)N * L) * sizeof(float);
}

int mc_masked_mse_fwd(const void* pred, int pred_elem_size, const float* imgs, const float* mask,
                      int N, int H, int W, int p, int norm_pix, float* loss_out,
                      float* mask_sum_out, void* ws, size_t ws_bytes, void* stream) {
  MC_ARCH_GUARD();
  MC_REQUIRE(pred && imgs && mask && loss_out && ws, MC_ERR_BAD_ARG, "masked_mse_fwd: null pointer");
  PatchGeom g;
  int rc = make_geom("masked_mse_fwd", N, H, W, p, g);
  if (rc) return rc;
  MC_REQUIRE(pred_elem_size == 2 || pred_elem_size == 4, MC_ERR_BAD_ARG,
             "masked_mse_fwd: pred_elem_size %d", pred_elem_size);
  MC_REQUIRE(ws_bytes >= mc_masked_mse_workspace_bytes(N, g.L), MC_ERR_WORKSPACE,
             "masked_mse_fwd: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long patches = (long long)N * g.L;
  const int grid = mse_grid(patches);
  unsigned int* counter = static_cast<unsigned int*>(ws);
  float* part = reinterpret_cast<float*>(static_cast<char*>(ws) + 256);
  MC_CUDA(cudaMemsetAsync(counter, 0, 4, st));
  const bool fast = p == 16 && W % 4 == 0 && aligned(imgs, 16) && aligned(pred, 16);
#define MC_MSE_FWD(PT, F)                                                                              \
  masked_mse_fwd_kernel<PT, F><<<grid, 256, 0, st>>>((const PT*)pred, imgs, mask, patches, g, norm_pix, \
                                                     part, counter, loss_out, mask_sum_out)
  if (pred_elem_size == 4) { if (fast) MC_MSE_FWD(float, true); else MC_MSE_FWD(float, false); }
  else { if (fast) MC_MSE_FWD(__nv_bfloat16, true); else MC_MSE_FWD(__nv_bfloat16, false); }
#undef MC_MSE_FWD
  MC_LAUNCH_CHECK();
  return MC_OK;
}

int mc_masked_mse_bwd(const void* pred, int pred_elem_size, const float* imgs, const float* mask,
                      int N, int H, int W, int p, int norm_pix, const float* mask_sum,
                      const float* grad_loss, void* dpred, void* stream) {
  MC_ARCH_GUARD();
  MC_REQUIRE(pred && imgs && mask && mask_sum && dpred, MC_ERR_BAD_ARG, "masked_mse_bwd: null pointer");
  PatchGeom g;
  int rc = make_geom("masked_mse_bwd", N, H, W, p, g);
  if (rc) return rc;
  MC_REQUIRE(pred_elem_size == 2 || pred_elem_size == 4, MC_ERR_BAD_ARG,
             "masked_mse_bwd: pred_elem_size %d", pred_elem_size);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long patches = (long long)N * g.L;
  const int grid = mse_grid(patches);
  const bool fast = p == 16 && W % 4 == 0 && aligned(imgs, 16) && aligned(pred, 16) && aligned(dpred, 16);
#define MC_MSE_BWD(PT, F)                                                                              \
  masked_mse_bwd_kernel<PT, F><<<grid, 256, 0, st>>>((const PT*)pred, imgs, mask, patches, g, norm_pix, \
                                                     mask_sum, grad_loss, (PT*)dpred)
  if (pred_elem_size == 4) { if (fast) MC_MSE_BWD(float, true); else MC_MSE_BWD(float, false); }
  else { if (fast) MC_MSE_BWD(__nv_bfloat16, true); else MC_MSE_BWD(__nv_bfloat16, false); }
#undef MC_MSE_BWD
  MC_LAUNCH_CHECK();
  return MC_OK;
}

int mc_patchify(const float* imgs, int N, int H, int W, int p, int norm_pix, float* out,
                void* stream) {
  MC_ARCH_GUARD();
  MC_REQUIRE(imgs && out, MC_ERR_BAD_ARG, "patchify: null pointer");
  PatchGeom g;
  int rc = make_geom("patchify", N, H, W, p, g);
  if (rc) return rc;
  const long long patches = (long long)N * g.L;
  patchify_kernel<<<mse_grid(patches), 256, 0, static_cast<cudaStream_t>(stream)>>>(imgs, patches, g,
                                                                                     norm_pix, out);
  MC_LAUNCH_CHECK();
  return MC_OK;
}

}  // extern "C"
