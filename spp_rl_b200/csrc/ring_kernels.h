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

}  // namespace spp
