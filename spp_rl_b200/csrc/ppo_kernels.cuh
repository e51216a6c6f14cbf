// SPP-PPO device code: 64-wide tanh nets over very large on-policy batches.
// One policy, data-parallel over rows: every CTA owns a contiguous chunk of rows, runs the MLP chain on it with the
// same FFMA tile GEMMs as the off-policy kernels, and leaves its partial gradient in a private slot; a second
// kernel sums the slots in a fixed order (deterministic), after which the host may all-reduce the vector over
// NCCL (data-parallel SPP-PPO) before the Adam kernel applies it.
#pragma once
#include "gemm_tile.cuh"
#include "layout.h"

namespace spp {

constexpr int kPpoHidden = 64;                  // rltoolkit/basic_model.py:15-17,69-71
using MidTile = TileCfg<32, 8, 4, 8>;           // 128 x 64: one tile spans every hidden unit
using SmallTile = TileCfg<16, 16, 4, 4>;        // 64 x 64: a whole dW of the 64-wide layers

struct PpoLayout {
    int ob, ac, ldo, lda;
    NetDesc actor;      // fc1 [64 x ob], fc2 [64 x 64], fc3 [ob x 64], pseudo layer 3: log_scale [1 x ob] (vector)
    NetDesc critic;     // fc1 [64 x ob], fc2 [64 x 64], fc3 [1 x 64] (vector)
};

inline PpoLayout make_ppo_layout(int ob, int ac) {
    PpoLayout L{};
    L.ob = ob; L.ac = ac; L.ldo = pad4(ob); L.lda = pad4(ac);
    int off = 0;
    add_layer(L.actor, off, kPpoHidden, ob, 0);
    add_layer(L.actor, off, kPpoHidden, kPpoHidden, 0);
    add_layer(L.actor, off, ob, kPpoHidden, 0);
    add_layer(L.actor, off, 1, ob, 0, false);
    L.actor.size = off;
    off = 0;
    add_layer(L.critic, off, kPpoHidden, ob, 0);
    add_layer(L.critic, off, kPpoHidden, kPpoHidden, 0);
    add_layer(L.critic, off, 1, kPpoHidden, 0, false);
    L.critic.size = off;
    return L;
}

struct PpoHyper {
    float gamma, discount /* gamma * lambda */, epsilon, entropy_coef, custom_loss;
    double discount_d;
};

// rows of the current on-policy batch (any order in which each trajectory is contiguous and time-ascending)
struct PpoData {
    int64_t N;              // rows held by this rank
    int64_t Ntot;           // rows over all ranks (means are over the global batch)
    float* x;               // [N][ldo]  normalised obs          (Memory.norm_obs)
    float* xn;              // [N][ldo]  normalised next obs     (Memory.norm_next_obs)
    float* act;             // [N][ldo]  sampled state targets (normalised space, as stored by the rollout)
    float* logp;            // [N]       log-prob of the sampled target under the rollout policy
    float* rew; float* done; float* end;    // [N]
    float* v; float* nv;    // [N]       V(obs), V(next_obs)
    float* q;               // [N]       r + gamma (1 - done) V(next_obs)
    float* adv;             // [N]       GAE advantages (normalised in place by ppo_adv_normalize)
    int64_t* traj_start;    // [n_traj]  first row of every trajectory scanned by one thread (E entries for [E][T] data)
    int64_t* traj_len;      // [n_traj]
    int64_t traj_stride;    // row stride between consecutive time steps of one trajectory (1: rollout-major, E: step-major)
    int n_traj;
};

// minibatch staging (rows gathered by the shuffled permutation)
struct PpoBatch {
    int64_t n;              // rows in this minibatch (on this rank)
    int64_t n_mean;         // rows the loss mean is taken over (the global minibatch in data-parallel runs)
    float* x; float* act; float* xn;    // [n][ldo]
    float* logp; float* adv;            // [n]
};

struct PpoScratch {
    float* h1; float* h2; float* dz2; float* dz1;   // [rows][64]
    float* mean; float* t3; float* d3;              // [rows][ldo]  actor head: tanh*lim, tanh, grad of pre-activation
    float* newlogp;                                 // [rows]
};

struct PpoArgs {
    PpoLayout L;
    PpoHyper h;
    PpoData d;
    PpoBatch b;
    PpoScratch s;
    float* actor; float* actor_m; float* actor_v;       // parameter / moment arenas (NetDesc offsets)
    float* critic; float* critic_m; float* critic_v;
    const float* norm;          // [NORM_COUNT][ldo] as in the off-policy kernels (denormalise offsets / scales, limits)
    float* part;                // [grid][part_stride] per-CTA partial gradients
    int part_stride;
    float* gbuf;                // [part_stride] reduced gradient (what NCCL all-reduces in data-parallel runs)
    float* scal;                // [grid][8] per-CTA scalar partials; reduced into gscal[8]
    float* gscal;
    int rows_per_cta;           // multiple of 128
    int mode;                   // kernel specific
    int a2c;                    // actor gradient kernel: A2C policy gradient -mean(logp * adv) instead of the clipped ratio (a2c.py:279-283)
};

// sum of element i over the CTA slots of a gradient kernel in a FIXED order (deterministic): eight independent partial sums over
// interleaved slots (eight loads in flight instead of one serial chain: 21 -> ~5 us for 148 slots), combined as a fixed tree
__device__ __forceinline__ float slot_sum(const float* __restrict__ part, size_t stride, int n_part, int i) {
    float acc[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[u] = 0.f;
    int p = 0;
    for (; p + 8 <= n_part; p += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = part[(size_t)(p + u) * stride + i];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] += v[u];
    }
    for (int u = 0; p < n_part; ++p, ++u) acc[u] += part[(size_t)p * stride + i];
    return ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
}

// scalar slots
enum { PS_LOSS = 0, PS_ENTROPY = 1, PS_DIST = 2, PS_KL = 3, PS_COUNT = 8 };

// One optimiser step applied ELEMENT-WISE right where an element's reduced gradient becomes known (Adam is element-wise): the reduce /
// all-reduce kernels of the library's own loops call this instead of leaving the gradient for separate record + Adam launches.
// i = float offset inside the net's arena (NetDesc offsets); also writes the step's record slot (8 reduced scalars, then log_scale as
// it was BEFORE the step -- what ppo_record_kernel wrote).
struct StepArgs {
    NetDesc d; float* W; float* Mo; float* Vo; AdamScalars s; float extra_ls_grad; int ls_layer;
    float* slog; int ldo; int with_ls; int enabled;
};
__device__ __forceinline__ void apply_step_element(const StepArgs& t, int i, float g) {
    if (!t.with_ls && t.slog && i < t.ldo) t.slog[PS_COUNT + i] = 0.f;
    for (int li = 0; li < t.d.n_layers; ++li) {
        const LayerDesc& l = t.d.L[li];
        const int nw = l.rows * l.ld;
        if (i >= l.off_w && i < l.off_w + nw) {
            const int rel = i - l.off_w, row = rel / l.ld, col = rel % l.ld;
            if (li == t.ls_layer && t.with_ls && t.slog && col < t.ldo) t.slog[PS_COUNT + col] = t.W[i];      // log_scale before the step
            if (col >= l.cols) return;                                    // pad columns stay zero
            if (li == t.ls_layer) g += t.extra_ls_grad;                   // - entropy_coef * d entropy / d log_scale (= 1 per dim)
            float m = t.Mo[i], v = t.Vo[i];
            const float wn = adam_element(t.W[i], g, m, v, t.s);
            t.W[i] = wn; t.Mo[i] = m; t.Vo[i] = v;
            if (l.off_wt >= 0) t.W[l.off_wt + (size_t)col * l.ld_t + row] = wn;
            return;
        }
        if (i >= l.off_b && i < l.off_b + l.rows) {
            if (l.off_wt < 0 && li == t.ls_layer) return;                 // the log_scale pseudo layer has no bias
            float m = t.Mo[i], v = t.Vo[i];
            const float wn = adam_element(t.W[i], g, m, v, t.s);
            t.W[i] = wn; t.Mo[i] = m; t.Vo[i] = v;
            return;
        }
    }
}


cudaError_t launch_ppo_critic_values(const PpoArgs& a, int grid, cudaStream_t s);      // v, nv, q
cudaError_t launch_ppo_critic_grad(const PpoArgs& a, int grid, cudaStream_t s);        // partial grads of 0.5 mean((q - V)^2)
cudaError_t launch_ppo_critic_grad_tc(const PpoArgs& a, int grid, cudaStream_t s);     // the same on tcgen05 (ppo_critic_tc.cu); grid <= n_part slots
cudaError_t launch_ppo_critic_values_tc(const PpoArgs& a, int grid, cudaStream_t s);   // v / nv / q with fc2 on tcgen05, two CTAs per SM
cudaError_t launch_ppo_actor_grad_tc(const PpoArgs& a, int grid, cudaStream_t s);      // the actor minibatch gradient with the same tile machinery
bool ppo_critic_tc_supported(int ob, int ldo);
cudaError_t launch_ppo_actor_grad(const PpoArgs& a, int grid, cudaStream_t s);         // partial grads of the clipped loss
cudaError_t launch_ppo_act(const PpoArgs& a, int grid, cudaStream_t s);                // on-policy rollout step
cudaError_t launch_ppo_reduce(const PpoArgs& a, int n_part, int n_elems, cudaStream_t s);
cudaError_t launch_ppo_reduce_step(const PpoArgs& a, int n_part, int n_elems, const StepArgs& t, cudaStream_t s);      // reduce + record + Adam, one launch
cudaError_t launch_ppo_adam(const PpoArgs& a, int which_net, int step, double lr, cudaStream_t s);
cudaError_t launch_ppo_gae(const PpoArgs& a, cudaStream_t s);
cudaError_t launch_ppo_adv_stats(const PpoArgs& a, double* d_stats, int grid, cudaStream_t s);     // n, sum, sum of squares (fp64)
cudaError_t launch_ppo_adv_apply(const PpoArgs& a, float mean, float inv_std, int grid, cudaStream_t s);
cudaError_t launch_ppo_gather(const PpoArgs& a, const int64_t* d_perm, int denorm, int grid, cudaStream_t s);
cudaError_t launch_ppo_normalize_rows(const float* raw, float* out, int64_t n, int ob, int ldo, const float* norm, int clamp,
                                      int grid, cudaStream_t s);

}  // namespace spp
