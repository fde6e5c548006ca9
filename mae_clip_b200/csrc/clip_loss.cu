// C-ABI entry points of the contrastive soft-target loss (L3-L6): argument checks, engine
// dispatch (SIMT fp32 strips vs tcgen05 fused tiles) and the single-GPU fused / host-buffer
// conveniences.  Reference: /root/reference CLIP.py:34-43 (+ autograd at main.py:58).
#include <stdlib.h>

#include "clip_loss.cuh"

namespace mc {

static int check_problem(const char* who, const float* I_all, const float* T_all, int b, int B, int D,
                         int row_offset, float tau, int mode) {
  MC_REQUIRE(b > 0 && B > 0 && D > 0 && b <= B, MC_ERR_BAD_ARG, "%s: bad sizes b=%d B=%d D=%d", who,
             b, B, D);
  MC_REQUIRE(row_offset >= 0 && row_offset + b <= B, MC_ERR_BAD_ARG, "%s: row_offset %d out of range",
             who, row_offset);
  MC_REQUIRE(tau > 0.f, MC_ERR_BAD_ARG, "%s: temperature must be positive (got %g)", who, tau);
  MC_REQUIRE(mode >= MC_GEMM_SIMT_FP32 && mode <= MC_GEMM_TC_F16, MC_ERR_BAD_ARG, "%s: bad mode %d",
             who, mode);
  // the tcgen05 engine reads the staged planes only; the fp32 embeddings may then be absent
  // (peer-memory staging never assembles them)
  MC_REQUIRE((I_all && T_all) || (mode != MC_GEMM_SIMT_FP32 && tc::supported(D)), MC_ERR_BAD_ARG,
             "%s: null embedding pointer", who);
  return MC_OK;
}

// Shapes the tcgen05 engine does not cover (D other than 128 / 256) run on the fp32 SIMT
// engine: same device, same ABI, true-fp32 arithmetic - never a CPU path.
static int eff_mode(int mode, int D) {
  return (mode != MC_GEMM_SIMT_FP32 && !tc::supported(D)) ? (int)MC_GEMM_SIMT_FP32 : mode;
}

struct FusedLayout {
  size_t off_vec, off_flags, off_planes, off_phase, total;
  size_t vec_stride, flags_bytes;
};
static FusedLayout fused_layout(int B, int D, int mode) {
  mode = eff_mode(mode, D);
  FusedLayout l;
  l.vec_stride = round_up((size_t)B * 4, 256);
  l.off_vec = 0;
  l.flags_bytes = (mode == MC_GEMM_SIMT_FP32) ? 0 : round_up(tc::tile_flags_bytes(B, B), 256);
  l.off_flags = round_up(6 * l.vec_stride + 256, 256);    // raw flags of the statistics sweep, then the symmetrised ones
  l.off_planes = round_up(l.off_flags + 2 * l.flags_bytes, 1024);  // planes 1024-aligned
  size_t planes = (mode == MC_GEMM_SIMT_FP32) ? 0 : tc::planes_bytes(B, D, mode);
  l.off_phase = l.off_planes + round_up(planes, 256);
  size_t phase = (mode == MC_GEMM_SIMT_FP32) ? simt::workspace_bytes(B, B, D)
                                             : tc::workspace_bytes(B, B, D, mode);
  l.total = l.off_phase + round_up(phase, 256);
  return l;
}

}  // namespace mc

using namespace mc;

extern "C" {

size_t mc_clip_loss_workspace_bytes(int b, int B, int D, int mode) {
  if (b <= 0 || B <= 0 || D <= 0) return 0;
  mode = eff_mode(mode, D);
  return mode == MC_GEMM_SIMT_FP32 ? simt::workspace_bytes(b, B, D) : tc::workspace_bytes(b, B, D, mode);
}

size_t mc_clip_planes_bytes(int B, int D, int mode) {
  if (B <= 0 || D <= 0 || eff_mode(mode, D) == MC_GEMM_SIMT_FP32) return 0;
  return tc::planes_bytes(B, D, mode);
}

int mc_clip_prepare(const float* I_loc, const float* T_loc, int b, int B, int D, int row_offset,
                    int mode, void* planes_all, void* stream) {
  MC_ARCH_GUARD();
  MC_REQUIRE(I_loc && T_loc, MC_ERR_BAD_ARG, "clip_prepare: null pointer");
  MC_REQUIRE(b > 0 && B >= b && D > 0 && row_offset >= 0 && row_offset + b <= B, MC_ERR_BAD_ARG,
             "clip_prepare: bad sizes");
  MC_REQUIRE(mode >= MC_GEMM_SIMT_FP32 && mode <= MC_GEMM_TC_F16, MC_ERR_BAD_ARG, "clip_prepare: bad mode %d", mode);
  if (eff_mode(mode, D) == MC_GEMM_SIMT_FP32) return MC_OK;
  MC_REQUIRE(planes_all, MC_ERR_BAD_ARG, "clip_prepare: planes_all is null");
  return tc::prepare(I_loc, T_loc, b, B, D, row_offset, mode, planes_all,
                     static_cast<cudaStream_t>(stream));
}

size_t mc_clip_tile_flags_bytes(int b, int B, int D, int mode) {
  if (b <= 0 || B <= 0 || D <= 0 || eff_mode(mode, D) == MC_GEMM_SIMT_FP32) return 0;
  return tc::tile_flags_bytes(b, B);
}

int mc_clip_flags_finalize(const uint8_t* flags_all, int B, int b, int row_offset, uint8_t* flags_loc, void* stream) {
  MC_ARCH_GUARD();
  MC_REQUIRE(flags_all && flags_loc && b > 0 && B >= b && row_offset >= 0 && row_offset + b <= B, MC_ERR_BAD_ARG,
             "clip_flags_finalize: bad argument");
  return tc::flags_finalize(flags_all, B, b, row_offset, flags_loc, static_cast<cudaStream_t>(stream));
}

int mc_clip_amax(const float* I_loc, const float* T_loc, int b, int D, float* I_copy, float* T_copy,
                 unsigned int* amax_bits, void* stream) {
  MC_ARCH_GUARD();
  MC_REQUIRE(I_loc && T_loc && amax_bits && b > 0 && D > 0, MC_ERR_BAD_ARG, "clip_amax: bad argument");
  return tc::amax_copy(I_loc, T_loc, b, D, I_copy, T_copy, amax_bits, static_cast<cudaStream_t>(stream));
}

int mc_clip_push_shards(const float* I_loc, const float* T_loc, int b, int D, int rank, int world,
                        float* const* I_all_dst_host, float* const* T_all_dst_host,
                        unsigned int* const* amax_slots_host, unsigned int* scratch, void* stream) {
  MC_ARCH_GUARD();
  MC_REQUIRE(I_loc && T_loc && I_all_dst_host && T_all_dst_host && amax_slots_host && scratch && b > 0 && D > 0,
             MC_ERR_BAD_ARG, "clip_push_shards: bad argument");
  return tc::push_shards(I_loc, T_loc, b, D, rank, world, I_all_dst_host, T_all_dst_host, amax_slots_host, scratch,
                         static_cast<cudaStream_t>(stream));
}

int mc_clip_prepare_peers(const float* const* I_peers_host, const float* const* T_peers_host, int world, int b,
                          int D, int mode, const unsigned int* amax_slots, void* planes_all, void* stream) {
  MC_ARCH_GUARD();
  MC_REQUIRE(I_peers_host && T_peers_host && amax_slots && planes_all && b > 0 && D > 0, MC_ERR_BAD_ARG,
             "clip_prepare_peers: bad argument");
  MC_REQUIRE(mode == MC_GEMM_TC_F16X3 || mode == MC_GEMM_TC_F16, MC_ERR_UNSUPPORTED,
             "clip_prepare_peers: peer staging feeds the tcgen05 engine only (mode %d)", mode);
  return tc::prepare_peers(I_peers_host, T_peers_host, world, b, D, mode, amax_slots, planes_all,
                           static_cast<cudaStream_t>(stream));
}

int mc_clip_stats(const float* I_all, const float* T_all, const void* planes_all, int b, int B,
                  int D, int row_offset, float tau, int mode, float* r_loc, float* c_loc,
                  float* rz_loc, float* ps_loc, uint8_t* tile_flags_out, void* ws, size_t ws_bytes, void* stream) {
  MC_ARCH_GUARD();
  int rc = check_problem("clip_stats", I_all, T_all, b, B, D, row_offset, tau, mode);
  if (rc) return rc;
  MC_REQUIRE(r_loc && c_loc && rz_loc && ps_loc && ws, MC_ERR_BAD_ARG, "clip_stats: null output/workspace");
  ClipProblem p{I_all, T_all, planes_all, b, B, D, row_offset, tau};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  mode = eff_mode(mode, D);
  if (mode == MC_GEMM_SIMT_FP32) return simt::stats(p, r_loc, c_loc, rz_loc, ps_loc, ws, ws_bytes, st);
  p.tile_flags_out = tile_flags_out;
  return tc::stats(p, mode, r_loc, c_loc, rz_loc, ps_loc, ws, ws_bytes, st);
}

int mc_clip_rowloss(const float* I_all, const float* T_all, const void* planes_all, int b, int B,
                    int D, int row_offset, float tau, int mode, const float* r_all,
                    const float* c_all, const float* rz_all, const float* ps_loc, float* g_loc,
                    float* q_loc, float* loss_part, const uint8_t* tile_flags, void* ws, size_t ws_bytes,
                    void* stream) {
  MC_ARCH_GUARD();
  int rc = check_problem("clip_rowloss", I_all, T_all, b, B, D, row_offset, tau, mode);
  if (rc) return rc;
  MC_REQUIRE(r_all && c_all && rz_all && ps_loc && g_loc && q_loc && loss_part && ws, MC_ERR_BAD_ARG,
             "clip_rowloss: null pointer");
  ClipProblem p{I_all, T_all, planes_all, b, B, D, row_offset, tau};
  ClipStatsAll s{r_all, c_all, rz_all, nullptr, nullptr};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  mode = eff_mode(mode, D);
  if (mode == MC_GEMM_SIMT_FP32) return simt::rowloss(p, s, ps_loc, g_loc, q_loc, loss_part, ws, ws_bytes, st);
  p.tile_flags = tile_flags;
  return tc::rowloss(p, mode, s, ps_loc, g_loc, q_loc, loss_part, ws, ws_bytes, st);
}

int mc_clip_bwd(const float* I_all, const float* T_all, const void* planes_all, int b, int B, int D,
                int row_offset, float tau, int mode, const float* r_all, const float* c_all,
                const float* rz_all, const float* g_all, const float* q_all, const float* grad_loss,
                float* dI_loc, float* dT_loc, const uint8_t* tile_flags, void* ws, size_t ws_bytes, void* stream) {
  MC_ARCH_GUARD();
  int rc = check_problem("clip_bwd", I_all, T_all, b, B, D, row_offset, tau, mode);
  if (rc) return rc;
  MC_REQUIRE(r_all && c_all && rz_all && g_all && q_all && dI_loc && dT_loc && ws, MC_ERR_BAD_ARG,
             "clip_bwd: null pointer");
  ClipProblem p{I_all, T_all, planes_all, b, B, D, row_offset, tau};
  ClipStatsAll s{r_all, c_all, rz_all, g_all, q_all};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  mode = eff_mode(mode, D);
  if (mode == MC_GEMM_SIMT_FP32) return simt::bwd(p, s, grad_loss, dI_loc, dT_loc, ws, ws_bytes, st);
  p.tile_flags = tile_flags;
  return tc::bwd(p, mode, s, grad_loss, dI_loc, dT_loc, ws, ws_bytes, st);
}

size_t mc_clip_loss_fused_workspace_bytes(int B, int D, int mode) {
  if (B <= 0 || D <= 0) return 0;
  return fused_layout(B, D, mode).total;
}

int mc_clip_loss_fwd_bwd(const float* I, const float* T, int B, int D, float tau, int mode,
                         float* loss_out, float* dI, float* dT, void* ws, size_t ws_bytes,
                         void* stream) {
  MC_ARCH_GUARD();
  int rc = check_problem("clip_loss_fwd_bwd", I, T, B, B, D, 0, tau, mode);
  if (rc) return rc;
  MC_REQUIRE(loss_out && ws, MC_ERR_BAD_ARG, "clip_loss_fwd_bwd: null loss_out/workspace");
  MC_REQUIRE((dI == nullptr) == (dT == nullptr), MC_ERR_BAD_ARG,
             "clip_loss_fwd_bwd: pass both dI and dT or neither");
  FusedLayout l = fused_layout(B, D, mode);
  MC_REQUIRE(ws_bytes >= l.total, MC_ERR_WORKSPACE, "clip_loss_fwd_bwd: workspace %zu < %zu",
             ws_bytes, l.total);
  char* base = static_cast<char*>(ws);
  float* r = reinterpret_cast<float*>(base + l.off_vec + 0 * l.vec_stride);
  float* c = reinterpret_cast<float*>(base + l.off_vec + 1 * l.vec_stride);
  float* rz = reinterpret_cast<float*>(base + l.off_vec + 2 * l.vec_stride);
  float* g = reinterpret_cast<float*>(base + l.off_vec + 3 * l.vec_stride);
  float* q = reinterpret_cast<float*>(base + l.off_vec + 4 * l.vec_stride);
  float* ps = reinterpret_cast<float*>(base + l.off_vec + 5 * l.vec_stride);
  void* planes = base + l.off_planes;
  void* phase = base + l.off_phase;
  size_t phase_bytes = l.total - l.off_phase;
  // tile flags: the statistics sweep marks the tiles that carry soft-target mass; the later sweeps skip the rest
  // (MAE_CLIP_DENSE=1 keeps every tile: the A/B switch)
  static const bool dense = getenv("MAE_CLIP_DENSE") != nullptr && getenv("MAE_CLIP_DENSE")[0] == '1';
  uint8_t* flags_raw = (l.flags_bytes && !dense) ? reinterpret_cast<uint8_t*>(base + l.off_flags) : nullptr;
  uint8_t* flags = flags_raw ? flags_raw + l.flags_bytes : nullptr;
  if ((rc = mc_clip_prepare(I, T, B, B, D, 0, mode, planes, stream))) return rc;
  if ((rc = mc_clip_stats(I, T, planes, B, B, D, 0, tau, mode, r, c, rz, ps, flags_raw, phase, phase_bytes, stream)))
    return rc;
  if (flags_raw && (rc = mc_clip_flags_finalize(flags_raw, B, B, 0, flags, stream))) return rc;
  if ((rc = mc_clip_rowloss(I, T, planes, B, B, D, 0, tau, mode, r, c, rz, ps, g, q, loss_out, flags, phase,
                            phase_bytes, stream)))
    return rc;
  if (dI) {
    if ((rc = mc_clip_bwd(I, T, planes, B, B, D, 0, tau, mode, r, c, rz, g, q, nullptr, dI, dT, flags, phase,
                          phase_bytes, stream)))
      return rc;
  }
  return MC_OK;
}

size_t mc_clip_loss_host_workspace_bytes(int B, int D, int mode) {
  if (B <= 0 || D <= 0) return 0;
  size_t emb = round_up((size_t)B * D * 4, 256);
  return 4 * emb + 256 + fused_layout(B, D, mode).total;
}

int mc_clip_loss_fwd_bwd_host(const float* I_host, const float* T_host, int B, int D, float tau,
                              int mode, float* loss_host, float* dI_host, float* dT_host, void* dws,
                              size_t dws_bytes, void* stream) {
  MC_ARCH_GUARD();
  MC_REQUIRE(I_host && T_host && loss_host && dws, MC_ERR_BAD_ARG, "clip_loss_host: null pointer");
  MC_REQUIRE(B > 0 && D > 0, MC_ERR_BAD_ARG, "clip_loss_host: bad sizes");
  MC_REQUIRE(dws_bytes >= mc_clip_loss_host_workspace_bytes(B, D, mode), MC_ERR_WORKSPACE,
             "clip_loss_host: device workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  size_t emb = round_up((size_t)B * D * 4, 256);
  char* base = static_cast<char*>(dws);
  float* dIin = reinterpret_cast<float*>(base);
  float* dTin = reinterpret_cast<float*>(base + emb);
  float* gI = reinterpret_cast<float*>(base + 2 * emb);
  float* gT = reinterpret_cast<float*>(base + 3 * emb);
  float* loss = reinterpret_cast<float*>(base + 4 * emb);
  void* ws = base + 4 * emb + 256;
  size_t ws_bytes = dws_bytes - (4 * emb + 256);
  size_t bytes = (size_t)B * D * 4;
  MC_CUDA(cudaMemcpyAsync(dIin, I_host, bytes, cudaMemcpyHostToDevice, st));
  MC_CUDA(cudaMemcpyAsync(dTin, T_host, bytes, cudaMemcpyHostToDevice, st));
  bool want_grad = dI_host && dT_host;
  int rc = mc_clip_loss_fwd_bwd(dIin, dTin, B, D, tau, mode, loss, want_grad ? gI : nullptr,
                                want_grad ? gT : nullptr, ws, ws_bytes, stream);
  if (rc) return rc;
  MC_CUDA(cudaMemcpyAsync(loss_host, loss, 4, cudaMemcpyDeviceToHost, st));
  if (want_grad) {
    MC_CUDA(cudaMemcpyAsync(dI_host, gI, bytes, cudaMemcpyDeviceToHost, st));
    MC_CUDA(cudaMemcpyAsync(dT_host, gT, bytes, cudaMemcpyDeviceToHost, st));
  }
  MC_CUDA(cudaStreamSynchronize(st));
  return MC_OK;
}

}  // extern "C"
