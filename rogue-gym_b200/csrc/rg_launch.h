// rg_launch.h — host-callable launchers defined in rg_kernels.cu
#pragma once
#include <cuda_runtime.h>

#include "rg_types.h"

namespace rg {
cudaError_t configure_kernels(const DevBatch& b);
cudaError_t launch_reset(const DevBatch& b, cudaStream_t s);
// host mirror of the observation block (rg_mirror_get): where k_mirror reads shadows and writes the host
struct MirrorArgs {
  uint8_t* h_screen;     // host [N][C] dense
  uint8_t* h_hist;       // host [N][HB] bit-packed visited map
  uint32_t* h_status;    // host [N][10]
  int32_t* h_reward;     // host [N]
  uint8_t* h_done;       // host [N]
  uint32_t* h_message;   // host [N]
  uint8_t* h_error;      // host [N]
  uint8_t* s_screen;     // shadows, device: [N][CP]
  uint8_t* s_hist;       // [N][HB]
  uint8_t* s_flat;       // shadows of status, reward, message, done, error, each rounded up to whole 64-byte lines
  unsigned long long* bytes;  // [1] bytes stored to the host since the counter was last cleared
  // the pass that ends a host-facing step publishes its results itself (no copy-engine nodes in the step's graph):
  unsigned long long* h_bytes;  // mapped host: bytes stored by the step
  uint32_t* h_errflag;          // mapped host: OR of the errors raised
  uint32_t* ticket;             // device: blocks of the publishing pass that have finished
  unsigned long long* seq;      // device: publishing passes so far
  unsigned long long* h_seq;    // mapped host: the same, written last - the host may spin on it instead of waiting for the stream
  int wide;              // 1 = a changed screen piece is sent with its whole 64-byte line (RG_MIRROR_WIDE)
  int with_hist;         // 0 = the visited map is not mirrored (nobody asked for it: rg_mirror_get without history_bits)
  // mirror by lines (k_mirror_lines): mode 0 = k_mirror on every SM; 1 (default) = a few SMs (confined_sms) compare
  // whole 64-byte lines and store them to the host; only those SMs wait behind PCIe
  int mode, confined_sms;
  int fast_first;        // host-facing step: the active-monster branch starts after k_step_fast instead of beside it (RG_MIRROR_FAST_FIRST)
  uint8_t* h_base;       // device alias of the host block (every h_* pointer above lies inside it)
};

// streams and events one env-step is enqueued on (all owned by the batch)
struct StepStreams {
  cudaStream_t main;   // the batch's stream
  cudaStream_t side;   // high priority: the full-path kernel
  cudaStream_t mon;    // high priority: player + monster kernels of the envs with an active monster
  cudaStream_t mir;    // first host-mirror pass, beside the player / monster / full-path / reset kernels
  cudaEvent_t ev_fork, ev_join, ev_mon, ev_fast, ev_mir;
};
// one env-step = scan, then three branches side by side (full path | active-monster envs | thread-per-env kernel
// and its leftovers), reset pass, end
// `mirror` != nullptr adds the two host-mirror passes to the step (rg_step_mirror)
// actions_src: where k_step_scan reads the keys (device memory, or mapped host memory for the host-facing step);
// it leaves them in actions_dev, which every later kernel reads
cudaError_t launch_step(const DevBatch& b, const uint8_t* actions_src, uint8_t* actions_dev, int auto_reset, const StepStreams& q,
                        const MirrorArgs* mirror, int sm_count);
// background generation of next-episode games into the sp_* buffers; serves window `slot` of refill_win
cudaError_t launch_prefetch(const DevBatch& b, int warps, int slot, cudaStream_t s);
// background build of next-level skeletons requested up to the end of step slot (speculative descents)
cudaError_t launch_spec_build(const DevBatch& b, int slot, cudaStream_t s);
cudaError_t launch_test_move_enemy(const DevBatch& b, int64_t env, int fx, int fy, int tx, int ty, int* out3_dev,
                                   cudaStream_t s);
cudaError_t launch_encode(const DevBatch& b, int mode, uint32_t flag, int with_hist, int channels, float* out_dev,
                          cudaStream_t s);
cudaError_t launch_encode_compact(const DevBatch& b, uint8_t* sym_out_dev, int32_t* status_out_dev, uint8_t* hist_out_dev,
                                  int sm_count, cudaStream_t s);
cudaError_t launch_complete_maps(const DevBatch& b, int64_t env_lo, int64_t env_hi, cudaStream_t s);
// delta write-back of the whole observation block into the mapped host mirror (k_mirror, every env)
cudaError_t launch_mirror(const DevBatch& b, const MirrorArgs& m, int sm_count, cudaStream_t s);
// trainer-facing step: action index -> key before the step, float reward (+ stair bonus) after it
cudaError_t launch_keys_from_index(const DevBatch& b, const void* idx_dev, int index_bytes, uint8_t* keys_dev, cudaStream_t s);
cudaError_t launch_train_reward(const DevBatch& b, float stair_reward, int32_t* level_seen_dev, float* reward_out_dev,
                                cudaStream_t s);
cudaError_t launch_seed(const DevBatch& b, const uint64_t* lo_dev, const uint64_t* hi_dev, int seeded, int64_t count, cudaStream_t s);
cudaError_t launch_unpack_hist(const DevBatch& b, uint8_t* out_dev, cudaStream_t s);
cudaError_t launch_state_hash(const DevBatch& b, uint64_t* out_dev, cudaStream_t s);
cudaError_t launch_state_terminal(const DevBatch& b, uint8_t* out_dev, cudaStream_t s);
cudaError_t launch_export_rooms(const DevBatch& b, int16_t* rooms_dev, cudaStream_t s);
}  // namespace rg
