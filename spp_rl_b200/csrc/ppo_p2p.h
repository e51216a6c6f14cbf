// Arguments of the fused reduce + NVLink peer-memory all-reduce kernel (ppo_p2p.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ppo_kernels.cuh"

namespace spp {

constexpr int kP2pMaxRanks = 8;

struct P2pArgs {
    const float* part; int part_stride; int n_part; int n_elems;      // per-CTA partial gradients (n_part == 0: gbuf is already reduced)
    const float* scal;                                                // per-CTA scalar partials [n_part][8]
    float* gbuf; int total;                                           // result [part_stride + 8]
    float* peer[kP2pMaxRanks];                                        // every rank's exchange buffer [2][total] (own one included)
    uint32_t* peer_flags[kP2pMaxRanks];                               // every rank's flag words [2][kP2pMaxRanks]
    uint32_t epoch; int rank, world;
    unsigned* ticket; unsigned ticket_target;                         // last-CTA detection (monotonic counter)
    int* err;                                                         // set when a peer's flag never arrived
    StepArgs step;                                                    // step.enabled: record + Adam ride behind the exchange (element-wise)
};

cudaError_t launch_ppo_reduce_p2p(const P2pArgs& a, cudaStream_t s);

}  // namespace spp
