// Host-callable launchers for the replay-ring kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace spp {

struct RingView {
    const float* obs; const int32_t* oidx; const int32_t* nidx; const float* act; const float* rew;
    const uint8_t* done; const float* aacm;
    int64_t S; int ob, ac, ldo, lda;
};
struct GatherOut { float* obs; float* nobs; float* act; float* rew; int8_t* done; float* aacm; };

cudaError_t launch_ring_gather(const RingView& R, int agent, const int64_t* d_idx, int n, const GatherOut& o, cudaStream_t s);
cudaError_t launch_ring_gather_bench(const RingView& R, int P, int nb, int B, const int64_t* d_len, uint64_t seed,
                                     float* out_obs, float* out_nobs, float* out_aacm, float* out_rew, uint8_t* out_done,
                                     int grid, cudaStream_t s);
cudaError_t launch_ring_fill(const RingView& R, float* obs, int32_t* oidx, int32_t* nidx, float* act, float* rew,
                             uint8_t* done, uint8_t* end, float* aacm, int P, int64_t n, int T, uint64_t seed,
                             const float* norm, int norm_stride, int grid, cudaStream_t s);

// MetaReplayBuffer.update_obs_mean_std on the device (stats_kernels.cu): moments in d_moments [2][P][ob] (mean, std) and the four
// order statistics per column as sortable keys in d_state [P][ob][4][2] (key, 0) after the last pass
cudaError_t launch_ring_obs_stats(const RingView& R, int P, const int64_t* d_len, const int64_t* h_len, int nb, double* d_partial,
                                  double* d_moments, unsigned long long* d_state, unsigned int* d_hist, unsigned long long* h_state,
                                  cudaStream_t s);
float ring_stats_key_to_float(uint32_t key);

}  // namespace spp
