// "Next" row 1 (SURVEY.md section 8 f): the optimiser step of the training driver.
// Replaces /root/reference main.py:103-105 (torch.optim.AdamW over model.parameters()) and the
// `optimizer.step()` of main.py:59: decoupled weight decay, bias-corrected first / second moments,
// no amsgrad - torch's defaults (lr, betas (0.9, 0.999), eps 1e-8 come from the caller).
//
// HBM-bound: 16 bytes read + 12 bytes written per parameter.  Eager torch runs ~10 elementwise
// launches per tensor (or one multi-tensor launch per op); here ONE launch updates up to 48
// tensors: the pointer table travels by value in the kernel parameters, blockIdx.y selects the
// tensor, 16-byte vectors, grid-stride over the tensor.  The arithmetic follows torch's operation
// order (lerp for the first moment, sqrt(v) / sqrt(bias_correction2) + eps) so results agree to
// rounding.
#include <math.h>

#include "common.cuh"

namespace mc {

constexpr int kAdamTensors = 48;
struct AdamTable {
  float* p[kAdamTensors];
  const float* g[kAdamTensors];
  float* m[kAdamTensors];
  float* v[kAdamTensors];
  long long n[kAdamTensors];
};

struct AdamScalars {
  float lr_wd;       // 1 - lr * weight_decay
  float one_m_b1;    // 1 - beta1
  float b2, one_m_b2;
  float sqrt_bc2, eps, step_size;      // sqrt(1 - beta2^t), eps, lr / (1 - beta1^t)
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamScalars& s, float gs) {
  g *= gs;
  p *= s.lr_wd;
  m = m + (g - m) * s.one_m_b1;                 // exp_avg.lerp_(grad, 1 - beta1)
  v = v * s.b2 + (g * g) * s.one_m_b2;          // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  const float denom = sqrtf(v) / s.sqrt_bc2 + s.eps;
  p -= s.step_size * (m / denom);
}

__global__ void __launch_bounds__(256) adamw_kernel(AdamTable t, AdamScalars s, const float* __restrict__ grad_scale) {
  const int ti = blockIdx.y;
  const long long n = t.n[ti];
  float* __restrict__ p = t.p[ti];
  const float* __restrict__ g = t.g[ti];
  float* __restrict__ m = t.m[ti];
  float* __restrict__ v = t.v[ti];
  const float gs = grad_scale ? *grad_scale : 1.f;
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  long long done = 0;
  if (vec) {
    const long long n4 = n >> 2;
    for (long long i = tid; i < n4; i += nthreads) {
      float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
      const float4 gg = ld_stream(reinterpret_cast<const float4*>(g) + i);
      adam_one(pp.x, gg.x, mm.x, vv.x, s, gs);
      adam_one(pp.y, gg.y, mm.y, vv.y, s, gs);
      adam_one(pp.z, gg.z, mm.z, vv.z, s, gs);
      adam_one(pp.w, gg.w, mm.w, vv.w, s, gs);
      reinterpret_cast<float4*>(p)[i] = pp;
      reinterpret_cast<float4*>(m)[i] = mm;
      reinterpret_cast<float4*>(v)[i] = vv;
    }
    done = n4 << 2;
  }
  for (long long i = done + tid; i < n; i += nthreads) {
    float pp = p[i], mm = m[i], vv = v[i];
    adam_one(pp, g[i], mm, vv, s, gs);
    p[i] = pp; m[i] = mm; v[i] = vv;
  }
}

}  // namespace mc

using namespace mc;

extern "C" int mc_adamw_step(int ntensors, float* const* params, const float* const* grads, float* const* exp_avg,
                             float* const* exp_avg_sq, const int64_t* numel, double lr, double beta1, double beta2,
                             double eps, double weight_decay, int step, const float* grad_scale, void* stream) {
  MC_ARCH_GUARD();
  MC_REQUIRE(ntensors >= 0 && (ntensors == 0 || (params && grads && exp_avg && exp_avg_sq && numel)), MC_ERR_BAD_ARG,
             "adamw_step: null table");
  MC_REQUIRE(step >= 1, MC_ERR_BAD_ARG, "adamw_step: step counts from 1 (got %d)", step);
  MC_REQUIRE(lr >= 0. && eps >= 0. && beta1 >= 0. && beta1 < 1. && beta2 >= 0. && beta2 < 1. && weight_decay >= 0.,
             MC_ERR_BAD_ARG, "adamw_step: bad hyper-parameter");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  AdamScalars s;
  // hyper-parameters arrive as doubles and every derived scalar is formed in double before it is rounded
  // once to fp32 - as torch does with python floats (1 - 0.999 is 1e-3 there, not 1 - fp32(0.999))
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  s.lr_wd = (float)(1.0 - lr * weight_decay);
  s.one_m_b1 = (float)(1.0 - beta1);
  s.b2 = (float)beta2;
  s.one_m_b2 = (float)(1.0 - beta2);
  s.sqrt_bc2 = (float)sqrt(bc2);
  s.eps = (float)eps;
  s.step_size = (float)(lr / bc1);
  for (int base = 0; base < ntensors; base += kAdamTensors) {
    const int cnt = ntensors - base < kAdamTensors ? ntensors - base : kAdamTensors;
    AdamTable t;
    long long nmax = 0;
    for (int i = 0; i < kAdamTensors; ++i) {
      const bool live = i < cnt;
      t.p[i] = live ? params[base + i] : nullptr;
      t.g[i] = live ? grads[base + i] : nullptr;
      t.m[i] = live ? exp_avg[base + i] : nullptr;
      t.v[i] = live ? exp_avg_sq[base + i] : nullptr;
      t.n[i] = live ? (long long)numel[base + i] : 0;
      if (live) {
        MC_REQUIRE(t.n[i] >= 0 && (t.n[i] == 0 || (t.p[i] && t.g[i] && t.m[i] && t.v[i])), MC_ERR_BAD_ARG,
                   "adamw_step: tensor %d has a null pointer", base + i);
        if (t.n[i] > nmax) nmax = t.n[i];
      }
    }
    if (nmax == 0) continue;
    long long bx = (nmax / 4 + 255) / 256;
    const long long cap = (long long)num_sms() * 8;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    dim3 grid((unsigned)bx, (unsigned)cnt);
    adamw_kernel<<<grid, 256, 0, st>>>(t, s, grad_scale);
    MC_LAUNCH_CHECK();
  }
  return MC_OK;
}
