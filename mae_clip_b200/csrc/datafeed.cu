// "Next" row 4 (SURVEY.md section 8 f): the image side of the data feed, on the GPU.
// Replaces /root/reference dataset.py:33-34 + dataset.py:49 (albumentations Normalize with
// max_pixel_value = 255 followed by `torch.tensor(image).permute(2, 0, 1).float()`): uint8 HWC
// pixels -> normalised fp32 CHW planes.  albumentations (pinned 1.3.1, requirements.txt:4) is not
// under /root/reference nor installed here; its published algorithm is
//     img = float32(img); img -= mean * max_pixel_value; img *= float32(1 / (std * max_pixel_value))
// which this kernel follows operation for operation.  HBM-bound: 3 bytes read, 12 written per
// pixel; a thread converts four consecutive pixels (12 contiguous bytes in, one float4 per plane).
#include "common.cuh"

namespace mc {

struct NormConst {
  float sub[3];  // mean * max_pixel_value
  float mul[3];  // 1 / (std * max_pixel_value)
};

__global__ void __launch_bounds__(256) normalize_hwc_kernel(const uint8_t* __restrict__ in, long long npix4,
                                                            long long plane /* H*W */, NormConst k,
                                                            float* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix4;
       i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i * 4;              // first of four pixels; plane % 4 == 0 keeps them in one image
    const long long n = pix / plane, off = pix - n * plane;
    const uint32_t* p = reinterpret_cast<const uint32_t*>(in + pix * 3);
    const uint32_t w0 = p[0], w1 = p[1], w2 = p[2];
    const uint8_t b[12] = {(uint8_t)w0, (uint8_t)(w0 >> 8), (uint8_t)(w0 >> 16), (uint8_t)(w0 >> 24),
                           (uint8_t)w1, (uint8_t)(w1 >> 8), (uint8_t)(w1 >> 16), (uint8_t)(w1 >> 24),
                           (uint8_t)w2, (uint8_t)(w2 >> 8), (uint8_t)(w2 >> 16), (uint8_t)(w2 >> 24)};
    float* o = out + n * 3 * plane + off;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float4 v;
      v.x = ((float)b[c] - k.sub[c]) * k.mul[c];
      v.y = ((float)b[3 + c] - k.sub[c]) * k.mul[c];
      v.z = ((float)b[6 + c] - k.sub[c]) * k.mul[c];
      v.w = ((float)b[9 + c] - k.sub[c]) * k.mul[c];
      st_stream(reinterpret_cast<float4*>(o + c * plane), v);
    }
  }
}

// any H * W: one thread per pixel
__global__ void __launch_bounds__(256) normalize_hwc_scalar_kernel(const uint8_t* __restrict__ in, long long npix,
                                                                   long long plane, NormConst k, float* __restrict__ out) {
  for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < npix;
       pix += (long long)gridDim.x * blockDim.x) {
    const long long n = pix / plane, off = pix - n * plane;
#pragma unroll
    for (int c = 0; c < 3; ++c) out[(n * 3 + c) * plane + off] = ((float)in[pix * 3 + c] - k.sub[c]) * k.mul[c];
  }
}

// ---- token side of the feed (dataset.py:19-31).  The reference tokenises the whole caption list once in the dataset's
// constructor (`tokenizer(list(captions), padding=True, truncation=True, max_length=...)`: every row padded to ONE
// length L) and then, per sample, builds `torch.tensor(values[idx])` for input_ids / attention_mask, which the default
// collate stacks into (n, L) int64 tensors.  Here the tokenised corpus stays RESIDENT in HBM (N x L int64 per key: 128 MB
// for 40 k captions of 200 tokens) and a batch is one gather of rows by the sampler's indices - no per-sample host
// work, no per-step H2D besides the n indices.  One warp per (row, key); 16-byte vectors when L is even.
__global__ void __launch_bounds__(256) gather_token_rows_kernel(const long long* __restrict__ ids_all,
                                                                const long long* __restrict__ mask_all, long long N, int L,
                                                                const long long* __restrict__ idx, int n,
                                                                long long* __restrict__ ids_out,
                                                                long long* __restrict__ mask_out, int* __restrict__ bad) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= 2 * n) return;
  const int row = w >> 1, key = w & 1;
  long long src = idx[row];
  if (src < 0) src += N;                 // python-style negative indices, as values[idx] accepts them
  if (src < 0 || src >= N) {             // out of range: the reference raises IndexError; flag it, write zeros
    if (lane == 0) atomicExch(bad, 1);
    src = -1;
  }
  const long long* s = (key ? mask_all : ids_all) + (src < 0 ? 0 : src * L);
  long long* d = (key ? mask_out : ids_out) + (long long)row * L;
  if ((L & 1) == 0 && ((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(d)) & 15) == 0) {
    for (int i = lane; i < L / 2; i += 32) {
      uint4 v = src < 0 ? make_uint4(0, 0, 0, 0) : reinterpret_cast<const uint4*>(s)[i];
      reinterpret_cast<uint4*>(d)[i] = v;
    }
  } else {
    for (int i = lane; i < L; i += 32) d[i] = src < 0 ? 0 : s[i];
  }
}

}  // namespace mc

extern "C" int mc_gather_token_rows(const int64_t* ids_all, const int64_t* mask_all, int64_t N, int L,
                                    const int64_t* idx, int n, int64_t* ids_out, int64_t* mask_out,
                                    int* bad_index_flag, void* stream) {
  using namespace mc;
  MC_ARCH_GUARD();
  MC_REQUIRE(ids_all && mask_all && idx && ids_out && mask_out && bad_index_flag, MC_ERR_BAD_ARG, "gather_token_rows: null pointer");
  MC_REQUIRE(N > 0 && L > 0 && n >= 0, MC_ERR_BAD_ARG, "gather_token_rows: bad sizes N=%lld L=%d n=%d", (long long)N, L, n);
  if (n == 0) return MC_OK;
  const int warps = 2 * n, blocks = (warps + 7) / 8;
  gather_token_rows_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long*>(ids_all), reinterpret_cast<const long long*>(mask_all), (long long)N, L,
      reinterpret_cast<const long long*>(idx), n, reinterpret_cast<long long*>(ids_out),
      reinterpret_cast<long long*>(mask_out), bad_index_flag);
  MC_LAUNCH_CHECK();
  return MC_OK;
}

extern "C" int mc_normalize_images(const uint8_t* hwc, int N, int H, int W, const float* mean3_host,
                                   const float* std3_host, float max_pixel_value, float* out_nchw, void* stream) {
  using namespace mc;
  MC_ARCH_GUARD();
  MC_REQUIRE(hwc && out_nchw && mean3_host && std3_host, MC_ERR_BAD_ARG, "normalize_images: null pointer");
  MC_REQUIRE(N >= 0 && H > 0 && W > 0 && max_pixel_value > 0.f, MC_ERR_BAD_ARG, "normalize_images: bad sizes");
  if (N == 0) return MC_OK;
  NormConst k;
  for (int c = 0; c < 3; ++c) {
    MC_REQUIRE(std3_host[c] != 0.f, MC_ERR_BAD_ARG, "normalize_images: std[%d] is zero", c);
    k.sub[c] = mean3_host[c] * max_pixel_value;       // float32 products, as numpy does on float32 arrays
    k.mul[c] = 1.f / (std3_host[c] * max_pixel_value);
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long plane = (long long)H * W, npix = plane * N;
  const long long cap = (long long)num_sms() * 16;
  if (plane % 4 == 0 && aligned(hwc, 4) && aligned(out_nchw, 16)) {
    long long nb = (npix / 4 + 255) / 256;
    if (nb > cap) nb = cap;
    normalize_hwc_kernel<<<(int)nb, 256, 0, st>>>(hwc, npix / 4, plane, k, out_nchw);
  } else {
    long long nb = (npix + 255) / 256;
    if (nb > cap) nb = cap;
    normalize_hwc_scalar_kernel<<<(int)nb, 256, 0, st>>>(hwc, npix, plane, k, out_nchw);
  }
  MC_LAUNCH_CHECK();
  return MC_OK;
}
