// L3-L6, engines MC_GEMM_TC_BF16X3 / MC_GEMM_TC_BF16: tcgen05 fused tiles (placeholder until the
// kernels land; every entry reports MC_ERR_UNSUPPORTED so nothing silently falls back).
#include "clip_loss.cuh"

namespace mc {
namespace tc {

size_t workspace_bytes(int, int, int, int) { return 256; }
size_t planes_bytes(int, int, int) { return 256; }
int prepare(const float*, const float*, int, int, int, int, int, void*, cudaStream_t) {
  set_error("tcgen05 contrastive-loss engine not built yet");
  return MC_ERR_UNSUPPORTED;
}
int stats(const ClipProblem&, int, float*, float*, float*, void*, size_t, cudaStream_t) {
  set_error("tcgen05 contrastive-loss engine not built yet");
  return MC_ERR_UNSUPPORTED;
}
int rowloss(const ClipProblem&, int, const ClipStatsAll&, float*, float*, float*, void*, size_t,
            cudaStream_t) {
  set_error("tcgen05 contrastive-loss engine not built yet");
  return MC_ERR_UNSUPPORTED;
}
int bwd(const ClipProblem&, int, const ClipStatsAll&, const float*, float*, float*, void*, size_t,
        cudaStream_t) {
  set_error("tcgen05 contrastive-loss engine not built yet");
  return MC_ERR_UNSUPPORTED;
}

}  // namespace tc
}  // namespace mc
