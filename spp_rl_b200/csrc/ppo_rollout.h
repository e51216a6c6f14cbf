// Arguments of the device-resident SPP-PPO rollout (ppo_rollout.cu) and of the on-policy -> ACM replay ring hand-over.
#pragma once
#include "ppo_kernels.cuh"

namespace spp {

struct PpoRolloutArgs {
    PpoLayout L;
    const float* actor;         // policy parameter arena (PpoLayout::actor offsets)
    const float* norm;          // [NORM_COUNT][ldo]
    // ACM of the population the policy acts through (one agent)
    const float* acm;           // parameter arena of NET_ACM
    NetDesc acm_desc;
    int acm_kind, ac, lda, hm1, hm2, ldm1, ldm2;
    int stage_weights;          // set by the launcher: 1 = the layers' W^T and b are copied to shared memory once per launch (they fit)
    const float* acm_lim;       // [lda] environment action limit (AcM); BasicAcM uses its own t1
    // environments
    int E, T, max_ep_len;
    float done_prob;
    float* state;               // [E][ldo] raw observation of every environment (persists across launches)
    int* ep_len;                // [E] steps taken in the running episode
    // the [T][E] store (step-major rows, traj_stride = E)
    float *x, *xn, *act, *logp, *rew, *done, *end;
    float* aacm;                // [T * E][lda] ACM actions, or null
    float* raw_obs;             // [T * E][ldo] raw observations / next observations for the ACM replay ring, or null
    float* raw_next;
    // injected noise (tests): [T][E][ob] / [T][E]; null -> Philox(seed)
    const float* noise_act; const float* noise_env; const float* u_done; const float* noise_reset;
    uint64_t seed;
    int denorm_out;             // denormalize_actor_out
    int clamp;                  // mean-std normalisation clamps to +-10
    int rows_per_cta;           // environments per CTA: 8, 16, 24 or 32
};

size_t ppo_rollout_smem_bytes(const PpoRolloutArgs& a);
cudaError_t launch_ppo_rollout(const PpoRolloutArgs& a, cudaStream_t s);

// What ppo_abi.cu needs to know about a population (abi.cu owns the struct): its ACM, limits and replay ring.
struct PopulationAcmView {
    int device, ob, ac, lda, ldo, acm_kind, hm1, hm2, ldm1, ldm2;
    NetDesc acm_desc;
    const float* acm;           // agent's NET_ACM arena
    const float* acm_lim;
};

// What abi.cu needs to know about the [T][E] store a device rollout left in a policy (ppo_abi.cu owns the struct).
struct PpoStoreView {
    int device, E, T, ob, ldo, lda;
    const float* raw_obs;       // [T * E][ldo]
    const float* raw_next;      // [T * E][ldo]
    const float* aacm;          // [T * E][lda]
    const float* end;           // [T * E]
    cudaStream_t stream;
    cudaEvent_t ready;          // recorded behind the rollout kernel that filled the store
};

// device rows of the ACM replay ring written by ring_add_store_kernel: slot s of the obs ring takes obs_src[s] (store row, bit 62 set:
// the row's NEXT observation; -1: untouched), slot s of the timestep ring takes the ACM action of store row ts_src[s] with the
// observation indices ts_oidx / ts_nidx
cudaError_t launch_ring_add_store(float* r_obs, int32_t* r_oidx, int32_t* r_nidx, float* r_aacm, float* r_rew, uint8_t* r_done, uint8_t* r_end,
                                  int64_t S, int ob, int ac, int ldo, int lda, const int64_t* obs_src, const int64_t* ts_src,
                                  const int32_t* ts_oidx, const int32_t* ts_nidx, const PpoStoreView& st, cudaStream_t s);

cudaError_t launch_end_flags_env_major(const float* end, int E, int T, uint8_t* out, cudaStream_t s);

}  // namespace spp

struct spp_population;
struct spp_ppo;
int spp_ppo_store_view_(spp_ppo* p, spp::PpoStoreView* out);
int spp_population_acm_view_(spp_population* p, int agent, spp::PopulationAcmView* out);
