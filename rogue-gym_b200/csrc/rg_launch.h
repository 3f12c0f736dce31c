// rg_launch.h — host-callable launchers defined in rg_kernels.cu
#pragma once
#include <cuda_runtime.h>

#include "rg_types.h"

namespace rg {
cudaError_t configure_kernels(const DevBatch& b);
cudaError_t launch_reset(const DevBatch& b, cudaStream_t s);
// streams and events one env-step is enqueued on (all owned by the batch)
struct StepStreams {
  cudaStream_t main;   // the batch's stream
  cudaStream_t side;   // high priority: the full-path kernel
  cudaStream_t mon;    // high priority: monster kernels when the env range is stepped in pieces
  cudaEvent_t ev_fork, ev_join, ev_mon;
  cudaEvent_t ev_chunk[MAX_CHUNKS];
};
// one env-step = scan, full-path kernel beside {player, monster} kernels per piece, reset pass, end
cudaError_t launch_step(const DevBatch& b, const uint8_t* actions_dev, int auto_reset, const StepStreams& q);
// background generation of next-episode games into the sp_* buffers; serves window `slot` of refill_win
cudaError_t launch_prefetch(const DevBatch& b, int warps, int slot, cudaStream_t s);
cudaError_t launch_test_move_enemy(const DevBatch& b, int64_t env, int fx, int fy, int tx, int ty, int* out3_dev,
                                   cudaStream_t s);
cudaError_t launch_encode(const DevBatch& b, int mode, uint32_t flag, int with_hist, int channels, float* out_dev,
                          cudaStream_t s);
cudaError_t launch_complete_maps(const DevBatch& b, int64_t env_lo, int64_t env_hi, cudaStream_t s);
// delta write-back of the observation block into the mapped host mirror (k_mirror)
cudaError_t launch_mirror(const DevBatch& b, const rg_host_obs& host_dev_ptrs, uint8_t* host_hist_bits, uint8_t* s_screen,
                          uint8_t* s_hist, uint32_t* s_small, unsigned long long* bytes, int sm_count, cudaStream_t s);
cudaError_t launch_seed(const DevBatch& b, const uint64_t* lo_dev, const uint64_t* hi_dev, int seeded, cudaStream_t s);
cudaError_t launch_unpack_hist(const DevBatch& b, uint8_t* out_dev, cudaStream_t s);
cudaError_t launch_state_hash(const DevBatch& b, uint64_t* out_dev, cudaStream_t s);
}  // namespace rg
