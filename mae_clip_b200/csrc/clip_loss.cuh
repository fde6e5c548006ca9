// Internal interface between the C-ABI dispatcher (clip_loss.cu) and the two engines.
#pragma once
#include "common.cuh"

namespace mc {

struct ClipProblem {
  const float* I_all;
  const float* T_all;
  const void* planes_all;
  int b, B, D, row_offset;
  float tau;
  // tile relevance of the soft targets (tcgen05 engine; see mc_clip_tile_flags_bytes): written by the statistics
  // sweep, read (after mc_clip_flags_finalize) by the row-loss and gradient sweeps; null = dense
  uint8_t* tile_flags_out = nullptr;
  const uint8_t* tile_flags = nullptr;
  // Gradient form switch on the device (see tc::bwd_gate): 1 = the stored-weights / split kernels run, 0 = the own-rows
  // sweep runs; every gradient kernel is launched and the ones of the other form return at once.  Null = chosen on the host.
  const int* gate = nullptr;
};

struct ClipStatsAll {  // length-B vectors (device)
  const float* r;   // row LSE of S
  const float* c;   // col LSE of S
  const float* rz;  // row LSE of Z
  const float* g;   // row sums of G*P
  const float* q;   // col sums of P
};

namespace simt {
size_t workspace_bytes(int b, int B, int D);
int stats(const ClipProblem& p, float* r_loc, float* c_loc, float* rz_loc, float* ps_loc, void* ws,
          size_t ws_bytes, cudaStream_t st);
int rowloss(const ClipProblem& p, const ClipStatsAll& s, const float* ps_loc, float* g_loc, float* q_loc,
            float* loss_part, void* ws, size_t ws_bytes, cudaStream_t st);
int bwd(const ClipProblem& p, const ClipStatsAll& s, const float* grad_loss, float* dI, float* dT,
        void* ws, size_t ws_bytes, cudaStream_t st);
}  // namespace simt

namespace tc {
bool supported(int D);  // embedding widths the tcgen05 engine covers (others run on the fp32 SIMT engine)
size_t workspace_bytes(int b, int B, int D, int mode);       // incl. the stored-weights buffers when b == B
size_t core_workspace_bytes(int b, int B, int D, int mode);  // the sweeps' partials only
struct StoredLayout { size_t off_w, off_diz, off_cols, cols_bytes, total; };
StoredLayout stored_layout(int B, int D, int mode);
size_t planes_bytes(int B, int D, int mode);
int prepare(const float* I_loc, const float* T_loc, int b, int B, int D, int row_offset, int mode,
            void* planes_all, cudaStream_t st);
int amax_copy(const float* I_loc, const float* T_loc, int b, int D, float* I_copy, float* T_copy,
              unsigned int* amax_bits, cudaStream_t st);
int push_shards(const float* I_loc, const float* T_loc, int b, int D, int rank, int world, float* const* I_dst,
                float* const* T_dst, unsigned int* const* amax_slots, unsigned int* scratch, cudaStream_t st);
int prepare_peers(const float* const* I_peers, const float* const* T_peers, int world, int b, int D, int mode,
                  const unsigned int* amax_slots, void* planes_all, cudaStream_t st);
size_t tile_flags_bytes(int b, int B);   // [row blocks of b][column tiles of B] bytes
int flags_finalize(const uint8_t* flags_all, int B, int b, int row_offset, uint8_t* flags_loc, cudaStream_t st);
int stats(const ClipProblem& p, int mode, float* r_loc, float* c_loc, float* rz_loc, float* ps_loc, void* ws,
          size_t ws_bytes, cudaStream_t st, float* c_part_all = nullptr);
size_t stats_colpart_workspace_bytes(int b, int B, int D, int mode);
// the statistics sweep in pieces (host-buffer entry: one probe launch per arrived row chunk)
int stats_begin(const ClipProblem& p, cudaStream_t st);
int stats_chunk(const ClipProblem& p, int mode, int k, int chunks, void* ws, cudaStream_t st, float* c_part_all = nullptr);
int stats_end(const ClipProblem& p, int mode, int chunks, float* r_loc, float* c_loc, float* rz_loc, float* ps_loc, void* ws,
              cudaStream_t st, float* c_part_all = nullptr);
int prepare_chunk(const float* I, const float* T, int B, int D, int row0, int rows, void* planes_all, unsigned int* words,
                  cudaStream_t st);
int verify_scale(unsigned int* words, cudaStream_t st);
bool stats_chunkable(int B, int D, int chunks);
// stored-weights gradient: row half (S recompute -> dT, fp16 weight strip W) + column half (dI from W)
bool stored_form_enabled(int b, int B, int D);
size_t stored_weights_bytes(int b, int B);
size_t bwd_cols_workspace_bytes(int w_rows, int n_cols, int D);
int bwd_rows(const ClipProblem& p, int mode, const ClipStatsAll& s, const float* grad_loss, float* dT_loc, float* dIz_loc,
             void* W_rows, void* ws, size_t ws_bytes, cudaStream_t st, float* dI_loc_ownrows = nullptr);
int bwd_gate(const uint8_t* flags, size_t n_flags, int* gate_out, cudaStream_t st);
const float* bwd_rows_wscale(void* ws, int b, int B, int D);   // where bwd_rows(ws, strip of b rows) left the weight scale
int bwd_cols(const ClipProblem& p, int mode, const ClipStatsAll& s, const float* grad_loss, const void* W, int w_rows,
             int w_row_offset, int j0, int j1, const float* dIz, float* dI_out, void* ws, size_t ws_bytes, cudaStream_t st,
             const float* wscale_ready = nullptr);
int ranks_lse_merge(const float* parts, int n, int64_t stride, int B, float* c, cudaStream_t st);
int rowloss(const ClipProblem& p, int mode, const ClipStatsAll& s, const float* ps_loc, float* g_loc,
            float* q_loc, float* loss_part, void* ws, size_t ws_bytes, cudaStream_t st);
int bwd(const ClipProblem& p, int mode, const ClipStatsAll& s, const float* grad_loss, float* dI,
        float* dT, void* ws, size_t ws_bytes, cudaStream_t st);
}  // namespace tc

}  // namespace mc
