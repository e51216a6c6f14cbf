/* spp_rl_b200 -- C ABI of the B200-native SPP-RL hot path.
 *
 * The reference (raznem/spp-rl, package rltoolkit) is pure Python and has no FFI: every seam on
 * the hot path is a Python method.  This header declares the entry points a rltoolkit maintainer
 * would bind (ctypes stub in INTEGRATION.md) to replace those method bodies.  Each entry point
 * cites the reference method it replaces.  Plain pointers and sizes only; no C++ or torch types.
 *
 * Conventions
 *   - A `spp_population` holds P independent agents (the reference runs them as OS processes,
 *     train/spp_sac_hopper.py:115); agent index `a` in [0, P).  `a = -1` means "all agents".
 *   - Host arrays are dense row-major float32 unless stated; the library owns all device memory.
 *   - Return value: SPP_OK (0) or a negative error class; text via spp_last_error() (thread local).
 *   - Calls on one population must come from one host thread at a time; populations are independent.
 *   - Entry points ending in `_device` take DEVICE pointers and are asynchronous on `stream`
 *     (a cudaStream_t passed as void*; NULL = default stream).  All others are synchronous.
 */
#ifndef SPP_RL_B200_H
#define SPP_RL_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPP_ABI_VERSION 1

enum { SPP_OK = 0, SPP_ERR_ARG = -1, SPP_ERR_CUDA = -2, SPP_ERR_STATE = -3, SPP_ERR_UNSUPPORTED = -4 };
enum { SPP_ALGO_SAC = 0, SPP_ALGO_DDPG = 1 };
enum { SPP_ACM_MLP = 0 /* rltoolkit/basic_model.py:108-132 */, SPP_ACM_BASIC = 1 /* rltoolkit/acm/models/basic_acm.py:11-32 */ };
/* nets of an agent (DDPG uses CRITIC_1 / CRITIC_1_TARG as critic / critic_targ) */
enum {
    SPP_NET_ACTOR = 0, SPP_NET_CRITIC_1 = 1, SPP_NET_CRITIC_2 = 2, SPP_NET_ACM = 3,
    SPP_NET_CRITIC_1_TARG = 4, SPP_NET_CRITIC_2_TARG = 5, SPP_NET_ACTOR_TARG = 6
};
/* slots of one row of the loss output (8 floats per agent per update step) */
enum {
    SPP_LOSS_CRITIC_1 = 0, SPP_LOSS_CRITIC_2 = 1, SPP_LOSS_ACTOR = 2, SPP_LOSS_PI = 3 /* "sac"/"ddpg" */,
    SPP_LOSS_DIST = 4, SPP_LOSS_ALPHA = 5, SPP_LOSS_ALPHA_VALUE = 6, SPP_LOSS_COUNT = 8
};

typedef struct spp_population spp_population;

/* Hyper-parameters of the update path; names follow the reference constructor kwargs
 * (rltoolkit/algorithms/ddpg/ddpg.py:19-33, sac/sac.py:17-26, acm/acm.py:16-37,
 *  acm/off_policy/off_policy.py:9-15, acm/off_policy/ddpg_acm.py:11-13). */
typedef struct spp_config {
    int32_t algo;                 /* SPP_ALGO_* */
    int32_t ob_dim, ac_dim;       /* observation / env-action dims; state-target dim == ob_dim */
    int32_t acm_kind;             /* SPP_ACM_* */
    int32_t acm_critic;           /* critic sees the ACM's env action (1) or the state target (0) */
    int32_t norm_closs;           /* custom loss in normalised space */
    int32_t min_max_denormalize;  /* min-max (1) or mean-std (0) (de)normalisation */
    int32_t update_batch_size;    /* B */
    int32_t acm_batch_size;
    int32_t store_actions;        /* keep the state-target action column of the ring (needed if !acm_critic) */
    int64_t buffer_size;          /* replay ring capacity per agent */
    double gamma, tau, actor_lr, critic_lr, alpha_lr, acm_lr, custom_loss, alpha, target_entropy;
} spp_config;

int spp_abi_version(void);
const char* spp_last_error(void);

int spp_population_create(const spp_config* cfg, int population, int device, spp_population** out);
int spp_population_destroy(spp_population* p);
int spp_sync(spp_population* p);

/* ---- limits and normalisation statistics ------------------------------------------------------
 * actor_lim[ob]: AcMTrainer.actor_ac_lim (rltoolkit/acm/acm.py:102-108) broadcast to ob entries;
 * acm_lim[ac]:   MetaLearner.ac_lim (rltoolkit/rl.py:53).
 * stats: replay_buffer.min_obs/max_obs/obs_mean/obs_std (rltoolkit/buffer/memory.py:76-127); any
 * pointer may be NULL = unset (identity, as the reference does for None). */
int spp_set_limits(spp_population* p, const float* actor_lim, const float* acm_lim);
int spp_set_norm_stats(spp_population* p, int a, const float* min_obs, const float* max_obs,
                       const float* obs_mean, const float* obs_std);

/* ---- parameters: state_dict() / load_state_dict() of the reference nn.Modules -------------------
 * Tensors are enumerated per net in the reference's state_dict order with the reference's names
 * (e.g. "fc1.weight", "fc_prob.bias", "t1"); shapes are the reference's [rows, cols]. */
int spp_net_tensor_count(spp_population* p, int net);
int spp_net_tensor_info(spp_population* p, int net, int t, char* name, int name_cap, int* rows, int* cols);
int spp_params_upload(spp_population* p, int a, int net, int t, const float* host);
int spp_params_download(spp_population* p, int a, int net, int t, float* host);
/* Adam moments / step of a trainable tensor (torch.optim.Adam state; rltoolkit/rl.py:62). */
int spp_adam_download(spp_population* p, int a, int net, int t, float* exp_avg, float* exp_avg_sq, int* step);
int spp_adam_reset(spp_population* p, int a, int net);
/* target <- online deep copy (DDPG.critic setter, rltoolkit/algorithms/ddpg/ddpg.py:150-157;
 * SAC.critic_1 setter, rltoolkit/algorithms/sac/sac.py:124-136). */
int spp_sync_targets(spp_population* p, int a);
/* SAC temperature: log_alpha (float64, rltoolkit/algorithms/sac/sac.py:107-110). */
int spp_alpha_get(spp_population* p, int a, double* log_alpha, double* alpha);
int spp_alpha_set(spp_population* p, int a, double log_alpha);

/* ---- replay ring: BufferAcMOffPolicy (rltoolkit/buffer/replay_buffer.py:303-401) ----------------
 * add_obs :56-60, add_timestep :65-75 + ReplayBuffer.addition :133-137, add_acm_action :332-333.
 * The cursor state machine runs on the host (bit-exact), rows are written to device memory. */
int spp_ring_add_obs(spp_population* p, int a, const float* obs, int64_t* out_idx);
int spp_ring_add_acm_action(spp_population* p, int a, const float* acm_action);
int spp_ring_add_timestep(spp_population* p, int a, int64_t obs_idx, int64_t next_obs_idx,
                          const float* action, float reward, int done, int end);
int spp_ring_reset(spp_population* p, int a);
/* cursors: out[0] = obs cursor, out[1] = timestep cursor, out[2] = current_len (== len(buffer)) */
int spp_ring_state(spp_population* p, int a, int64_t out[3]);
/* sample_batch :385-398 for given indices (np.random.randint stays on the caller's side so index
 * streams are the reference's): two-level gather on the device, results to host arrays.
 * action may be NULL.  done is int8 like the reference. */
int spp_ring_sample_batch(spp_population* p, int a, const int64_t* idx, int n, float* obs, float* next_obs,
                          float* action, float* reward, int8_t* done, float* acm_action);
/* Synthetic prefill for benchmarks: n transitions per agent in episodes of
 * `episode_len`, obs ~ U(min_obs,max_obs), next = obs + 0.02 N(0,1), reward ~ N(0,1), done ~ Bern(1e-3). */
/* MetaReplayBuffer.update_obs_mean_std (rltoolkit/buffer/replay_buffer.py:83-96) on the device for every agent of the population:
 * out [P][6][ob] (host, fp64) = mean, population std (numpy ddof 0), and the four order statistics of self.obs per column that
 * np.percentile(obs, 1) and np.percentile(obs, 99) interpolate between (ranks floor / ceil of q (n - 1)); exact elements of the
 * buffer, so the caller finishes the percentile with numpy's own lerp.  bytes_out (optional): ring bytes streamed. */
int spp_ring_obs_stats(spp_population* p, double* out, double* bytes_out);
int spp_ring_fill_synthetic(spp_population* p, uint64_t seed, int64_t n, int episode_len);
/* Device-resident gather of n_batches minibatches per agent (indices drawn on device): measures the
 * HBM path of sample_batch in isolation.  bytes_out = algorithmic bytes moved per call. */
int spp_ring_gather_bench_device(spp_population* p, int n_batches, uint64_t seed, double* bytes_out, void* stream);
/* test hook: rows [first, first + n) of the dense minibatches the last spp_ring_gather_bench_device call produced (host arrays
 * [n][ob], [n][ob], [n][ac], [n], [n]); the kernel's index of global row t is mulhi64(Philox4x32-10(seed, agent, t mod (n_batches B)).xy, len) */
int spp_ring_gather_bench_rows(spp_population* p, int n_batches, int64_t first, int n, float* obs, float* nobs, float* aacm, float* rew,
                               int8_t* done);

/* ---- the update step: SAC_AcM.update (rltoolkit/acm/off_policy/sac_acm.py:89-162) /
 *      DDPG_AcM.update (rltoolkit/acm/off_policy/ddpg_acm.py:147-201), `grad_steps` in a row as
 *      DDPG.make_update does (rltoolkit/algorithms/ddpg/ddpg.py:231-237). ---------------------------
 * Host-batch form = the reference's update(obs, next_obs, action, reward, done, acm_action):
 *   obs, next_obs [P][G][B][ob]; action [P][G][B][ob] (NULL allowed when acm_critic); reward [P][G][B];
 *   done int8 [P][G][B]; acm_action [P][G][B][ac];
 *   eps [P][G][2][B][ob]: the N(0,1) draws of Normal.rsample, target pass first then policy pass
 *   (rltoolkit/algorithms/sac/models.py:45); NULL -> drawn on device (Philox, `seed`).  Ignored for DDPG.
 *   losses [P][G][8] (may be NULL).  Copies H2D, runs one fused kernel, copies losses D2H. */
int spp_update_host(spp_population* p, int grad_steps, const float* obs, const float* next_obs,
                    const float* action, const float* reward, const int8_t* done, const float* acm_action,
                    const float* eps, uint64_t seed, float* losses);
/* Ring form = sample_batch + update: idx int64 [P][G][B] host indices (NULL -> device sampler). */
int spp_update_ring(spp_population* p, int grad_steps, const int64_t* idx, const float* eps, uint64_t seed,
                    float* losses);
/* Fully device-resident form (no host traffic): device sampler + device noise; losses_dev may be NULL. */
int spp_update_ring_device(spp_population* p, int grad_steps, uint64_t seed, float* losses_dev, void* stream);
/* profiling: the same burst with %globaltimer stamps at agent 0's 24 stage boundaries; out_us[0..23) = mean microseconds per
 * stage over the grad_steps updates (stage names: tools/stage_profile.py) */
int spp_update_stage_profile(spp_population* p, int grad_steps, uint64_t seed, double* out_us, int cap, int* n_out);

/* ---- ACM regression: AcMTrainer.update_acm_batches (rltoolkit/acm/acm.py:356-372), n_batches x
 *      [rbuffer_sample_acm (rltoolkit/buffer/replay_buffer.py:404-430) -> acm_cat (acm.py:260-264) ->
 *      batch_update (acm.py:246-258)] in one launch; Adam(acm_lr) on the ACM (AcM or BasicAcM incl. t, t1).
 * Host form = batch_update(x, y): x [P][n][acm_batch_size][2 ob] = cat[obs, next_obs], y [P][n][acm_batch_size][ac].
 * Ring form: idx int64 [P][n][acm_batch_size] host indices (np.random.randint stays with the caller), NULL ->
 * device sampler.  losses [P][n] = the MSE of each step (may be NULL).  Only acm_ob_idx = all observations.
 * last_rows: rows of the FINAL step when it is a partial minibatch (DataLoader's drop_last = False in
 * AcMTrainer.update_acm, rltoolkit/acm/acm.py:285-293); 0 = every step is a full acm_batch_size batch. */
int spp_acm_update_host(spp_population* p, int n_batches, const float* x, const float* y, int last_rows, float* losses);
int spp_acm_update_ring(spp_population* p, int n_batches, const int64_t* idx, int last_rows, uint64_t seed, float* losses);
/* AcMTrainer.calculate_validation_loss (rltoolkit/acm/acm.py:329-343): forward + MSE only, no optimiser step, on the same
 * host layout as spp_acm_update_host (the validation set cut into acm_batch_size chunks, last_rows in the final one);
 * losses [P][n_batches] = the MSE of each chunk -- the caller weights them by their row counts. */
int spp_acm_eval_host(spp_population* p, int n_batches, const float* x, const float* y, int last_rows, float* losses);
/* obs_norm=True of the off-policy classes: ReplayBuffer._sample_batch normalises obs / next_obs (rltoolkit/buffer/replay_buffer.py:246-248);
 * the fused ring updates (spp_update_ring*, device or host indices) then gather normalised rows.  spp_update_host takes its batch as given. */
int spp_set_obs_norm(spp_population* p, int on);
/* learning rates may change between calls (StepLR on the ACM optimiser, rltoolkit/acm/acm.py:181-183,299); negative = keep */
int spp_set_learning_rates(spp_population* p, double actor_lr, double critic_lr, double alpha_lr, double acm_lr);

/* ---- rollout step: body of DDPG.collect_batch_and_train (rltoolkit/algorithms/ddpg/ddpg.py:202-207) =
 *      replay_buffer.normalize (obs_norm gate) -> AcMOffPolicy.initial_act (random phase,
 *      rltoolkit/acm/off_policy/off_policy.py:50-54) | DDPG_AcM.noise_action (rltoolkit/acm/off_policy/ddpg_acm.py:40-50)
 *      -> AcMOffPolicy.process_action (off_policy.py:89-106), for E observations per agent at once.
 * obs, noise (torch.randn of the reference), eps (the SAC actor's rsample draw; NULL = deterministic mean, as test()
 * uses it) are [P][E][ob]; out_target [P][E][ob] is the state target stored in the ring, out_action [P][E][ac] the
 * ACM action for the environment.
 * random_phase: 0 = the actor acts (noise_action); 1 = frames < random_frames, target = actor_ac_lim * noise (initial_act);
 * 2 = `noise` IS the state target already (a bare process_action call: no actor, no limit scale, no clip). */
int spp_rollout_step_host(spp_population* p, int E, const float* obs, const float* noise, const float* eps, int random_phase,
                          double act_noise, int obs_norm, int denormalize_actor_out, float* out_target, float* out_action);
/* Device-resident rollout of a synthetic environment (MuJoCo is unavailable offline): `steps` consecutive vectorised
 * steps of E environments per agent, everything of the frame loop except the simulator -- actor, noise, clip,
 * denormalise, ACM, and the ring writes (obs row, timestep row, acm action, reward, done) -- in one launch.
 * random_phase 1: the frames before `random_frames` (ddpg.py:205-206): state target = actor_ac_lim * N(0,1), no actor pass. */
int spp_rollout_synthetic_device(spp_population* p, int E, int steps, uint64_t seed, double act_noise, int random_phase, void* stream);

/* ---- SPP-PPO (PPO_AcM): one policy, 64-wide tanh nets (rltoolkit/basic_model.py:7-77), data-parallel over rows.
 * A `spp_ppo` holds the actor (+ log_scale) and critic, one on-policy batch of up to max_rows transitions and the
 * minibatch staging.  Net ids: 0 = actor, 1 = critic; tensors in the reference's state_dict order. */
typedef struct spp_ppo spp_ppo;
typedef struct spp_ppo_config {
    int32_t ob_dim, ac_dim;
    int32_t min_max_denormalize, norm_closs;
    int64_t max_rows;               /* capacity of the on-policy batch held by this rank */
    int64_t max_batch_rows;         /* largest minibatch (ppo_batch_size) */
    double gamma, gae_lambda, ppo_epsilon, entropy_coef, custom_loss, actor_lr, critic_lr;
} spp_ppo_config;
int spp_ppo_create(const spp_ppo_config* cfg, int device, spp_ppo** out);
int spp_ppo_destroy(spp_ppo* p);
int spp_ppo_sync(spp_ppo* p);
int spp_ppo_stream(spp_ppo* p, void** stream);
int spp_ppo_set_limits(spp_ppo* p, const float* actor_lim);
int spp_ppo_set_norm_stats(spp_ppo* p, const float* min_obs, const float* max_obs, const float* obs_mean, const float* obs_std);
int spp_ppo_tensor_count(spp_ppo* p, int net);
int spp_ppo_tensor_info(spp_ppo* p, int net, int t, char* name, int name_cap, int* rows, int* cols);
int spp_ppo_params_upload(spp_ppo* p, int net, int t, const float* host);
int spp_ppo_params_download(spp_ppo* p, int net, int t, float* host);
int spp_ppo_adam_download(spp_ppo* p, int net, int t, float* exp_avg, float* exp_avg_sq, int* step);
/* The collected batch = the contents of rltoolkit's Memory after A2C.collect_batch (rltoolkit/algorithms/a2c/a2c.py:144-184):
 * raw obs / next_obs [N][ob] (Memory.obs / Memory.next_obs, rltoolkit/buffer/memory.py:146-168; normalised on the device
 * like Memory.norm_obs), sampled state targets [N][ob], their log-probs, rewards, done, end [N] (float 0/1).
 * Trajectories: traj_start[k], traj_len[k] rows apart by traj_stride (1 for the reference's rollout-major order, E for a
 * step-major [T][E] layout); every trajectory must end with end = 1 exactly as collect_batch guarantees.
 * global_rows = rows over all ranks (0 = N). */
int spp_ppo_load_rollout(spp_ppo* p, int64_t N, const float* obs, const float* next_obs, const float* actions, const float* logp,
                         const float* rew, const float* done, const float* end, const int64_t* traj_start, const int64_t* traj_len,
                         int n_traj, int64_t traj_stride, int64_t global_rows);
int spp_ppo_set_global_rows(spp_ppo* p, int64_t global_rows);
/* A2C.update_critic (a2c.py:186-225): n_target_updates x [q = r + gamma (1-done) V(s') ; n_updates_per_target x full-batch
 * Adam on 0.5 mean((q - V(s))^2)].  Step-wise forms for data-parallel runs: targets, grad (-> reduced gradient buffer and
 * scalar slot 0 = sum of squared errors), [all-reduce spp_ppo_grad_buffer over NCCL], apply. */
int spp_ppo_update_critic(spp_ppo* p, int n_target_updates, int n_updates_per_target, float* mean_loss);
int spp_ppo_critic_targets(spp_ppo* p);
int spp_ppo_critic_grad(spp_ppo* p);
int spp_ppo_critic_apply(spp_ppo* p);
/* critic gradient kernel: 1 (default) = the 64-wide contractions on tcgen05 with weights resident in shared memory
 * (csrc/ppo_critic_tc.cu; observations of up to 19 floats), 0 = the FFMA tile kernel (also the path for wider observations) */
int spp_ppo_set_critic_path(spp_ppo* p, int tensor_cores);
/* keep n SMs out of the grids of the policy's row-sharded kernels (default 0): for a concurrent kernel of another stream that occupies
 * whole SMs for long (the ACM regression burst running beside the policy update), so that no launch spills into a second wave */
int spp_ppo_set_reserved_sms(spp_ppo* p, int n);
/* PPO.calculate_advantage = calculate_q_val + calculate_gae (rltoolkit/algorithms/ppo/ppo.py:101-150) incl. the bootstrap
 * at non-terminal ends; adv_host [N] may be NULL.  AdvantageDataset normalisation (advantage_dataset.py:9-12): unbiased std,
 * eps 1.2e-7; global_stats = (n, sum, sum of squares) over all ranks or NULL for local. */
int spp_ppo_advantages(spp_ppo* p, float* adv_host);
int spp_ppo_adv_stats(spp_ppo* p, double out[3]);
int spp_ppo_normalize_adv(spp_ppo* p, const double* global_stats);
/* the same with the caller's epsilon: A2C normalises with (A - mean) / (std + 1e-8) (rltoolkit/algorithms/a2c/a2c.py:275-277,
 * rltoolkit/acm/on_policy.py:101-104); eps < 0 = the PPO datasets' 1.2e-7 */
int spp_ppo_normalize_adv_eps(spp_ppo* p, const double* global_stats, double eps);
/* PPO_AcM.update_actor_acm (rltoolkit/acm/on_policy.py:164-216): perms [max_epochs][N] replace DataLoader(shuffle=True);
 * keeps the partial last minibatch, takes KL on the last minibatch only, stops when KL >= kl_threshold, and divides the
 * summed losses (actor, entropy, policy, dist) by (i + 1) exactly like the reference.  Step-wise forms as for the critic;
 * scalar slots after a grad call: 0 = sum of clipped-loss terms, 2 = sum of squared custom-loss distances, 3 = sum of
 * (old - new) log-probs. */
int spp_ppo_update_actor(spp_ppo* p, const int64_t* perms, int max_epochs, int batch_size, double kl_threshold, float losses[4],
                         int* epochs_run, float* last_kl);
int spp_ppo_actor_minibatch_grad(spp_ppo* p, const int64_t* perm, int64_t n, int64_t n_global);
/* the same with the local row ids already on the DEVICE (produced on spp_ppo_stream()'s stream, e.g. by filtering the global
 * minibatch there): no host copy, no host-side range check -- the ids must lie in [0, local rows). */
int spp_ppo_actor_minibatch_grad_device(spp_ppo* p, const int64_t* perm_dev, int64_t n, int64_t n_global);
/* advantages [N] handed in by the caller instead of spp_ppo_advantages (PPO.update_actor(advantages, buffer), ppo.py:152) */
int spp_ppo_load_advantages(spp_ppo* p, const float* adv_host);
int spp_ppo_adam_reset(spp_ppo* p, int net);      /* fresh Adam state of one net (a module assigned to model.actor / model.critic) */
/* 1: the actor epochs are plain PPO.update_actor (rltoolkit/algorithms/ppo/ppo.py:152-192), what PPO_AcM falls back to when
 * custom_loss == 0 (rltoolkit/acm/on_policy.py:88-98): no distance term, the log-prob is taken of the stored actions as they are,
 * and losses[] = {actor, entropy, sum, 0} are the raw sums over minibatches (no division by the epoch count).  0 (default):
 * PPO_AcM.update_actor_acm (on_policy.py:164-216).
 * 2: A2C_AcM.update_actor_acm (on_policy.py:100-124): one full-batch policy-gradient step on mean(-logp * adv) -- no ratio, no
 * entropy term; the distance term carries no gradient (Actor.act samples without rsample, basic_model.py:47) and is reported only
 * (scalar slot 2).  3: plain A2C.update_actor (a2c.py:267-285; what A2C_AcM runs when custom_loss == 0).  In modes 2 / 3
 * spp_ppo_advantages returns q - V(s) (A2C.calculate_advantage, a2c.py:227-245) instead of the GAE scan. */
int spp_ppo_set_actor_mode(spp_ppo* p, int mode);
/* A2C_AcM.update_actor_acm never calls zero_grad (on_policy.py:117-123): after spp_ppo_actor_minibatch_grad,
 * accumulate != 0 adds the fresh gradient to the running sum kept since the last spp_ppo_adam_reset(actor) and hands the SUM to
 * spp_ppo_actor_apply; accumulate == 0 restarts the sum (A2C.update_actor zeroes the gradients, a2c.py:282). */
int spp_ppo_grad_accumulate(spp_ppo* p, int accumulate);
/* One epoch with the LOCAL row ids on the device: minibatch k = ids_dev[off[k], off[k+1]) (host offsets, nb + 1 of them; a rank may
 * own none of a minibatch), n_global[k] = rows of minibatch k over all ranks (NULL = local).  No host synchronisation per step;
 * log_host [nb][8 + pad4(ob)] = reduced scalars of every minibatch followed by log_scale as that minibatch saw it. */
int spp_ppo_actor_epoch_device(spp_ppo* p, const int64_t* ids_dev, const int64_t* off, const int64_t* n_global, int nb, float* log_host);
/* ---- data parallelism (config 4, one policy, environments shard over ranks): the reference has no distributed code;
 *      here every optimiser step of spp_ppo_update_critic / spp_ppo_actor_epoch_device all-reduces the gradient vector together
 *      with the 8 scalars (one NCCL collective over NVLink on the policy's stream), spp_ppo_normalize_adv all-reduces the fp64
 *      advantage statistics.  id: 128 bytes from spp_comm_unique_id on rank 0, distributed by the caller (e.g. torch.distributed). */
int spp_comm_unique_id(char out[128]);
int spp_ppo_comm_init(spp_ppo* p, const char id[128], int rank, int world);
int spp_ppo_comm_info(spp_ppo* p, int* world, int64_t* allreduces, int* nccl_version);
/* The per-step gradient all-reduce over NVLink PEER MEMORY instead of NCCL (csrc/ppo_p2p.cu): one kernel per optimiser step sums the
 * gradient kernel's per-CTA partials, publishes the rank's vector in an exchange buffer its peers have mapped (CUDA IPC), waits for
 * theirs and adds all ranks' vectors in rank order (bit-identical on every rank).  spp_ppo_p2p_handle returns this rank's 64-byte IPC
 * handle; the caller gathers the handles of all ranks (any transport) and passes them, rank-major, to spp_ppo_p2p_init on every
 * rank after spp_ppo_comm_init.  Unsupported set-ups (no peer access) return SPP_ERR_UNSUPPORTED and the NCCL path stays.
 * spp_ppo_p2p_info: whether the path is on, steps taken through it, and the device-side time-out flag (a peer never arrived). */
int spp_ppo_p2p_handle(spp_ppo* p, char out[64]);
int spp_ppo_p2p_init(spp_ppo* p, const char* handles, int rank, int world);
int spp_ppo_p2p_info(spp_ppo* p, int* on, int64_t* steps, int* err);
int spp_ppo_actor_apply(spp_ppo* p);
int spp_ppo_scalars(spp_ppo* p, float out[8]);
/* Rollout step of A2C.collect_batch (rltoolkit/algorithms/a2c/a2c.py:165-167) for E observations at once:
 * Memory.normalize -> Actor.act (rltoolkit/basic_model.py:32-51; noise [E][ob] = the N(0,1) of Normal.sample) -> the
 * denormalisation half of AcMOnPolicyTrainer.process_action (rltoolkit/acm/on_policy.py:46-47).  action [E][ob] is the
 * sampled target the Memory stores, logp [E] its log-prob, target [E][ob] what is concatenated with the normalised obs
 * for the ACM (spp_rollout_step_host with random_phase = 2, obs_norm = 1 and noise = action evaluates that ACM call). */
int spp_ppo_act(spp_ppo* p, int64_t E, const float* obs, const float* noise, int denormalize_actor_out, float* action, float* logp,
                float* target);
/* The whole collect_batch loop on the device (rltoolkit/algorithms/a2c/a2c.py:144-184 + rltoolkit/acm/on_policy.py:34-53), vectorised
 * over E synthetic environments (MuJoCo is unavailable offline) for T steps in ONE launch: normalise -> Actor.act -> denormalise ->
 * ACM of `pop`'s agent (it sees cat[normalised obs, target], quirk 18) -> env.step -> the policy's [T][E] store (step-major rows,
 * traj_stride = E: what spp_ppo_update_critic / spp_ppo_advantages / the actor epochs read), time-limit truncation at max_ep_len
 * (a2c.py:168-171), `end` set at episode ends and at the last step of the batch.  The environments persist across calls
 * (reset_envs = 1 restarts them).  noise_act / noise_env / noise_reset [T][E][ob] and u_done [T][E] are injected host tensors
 * (parity tests); NULL draws from Philox(seed).  Replaces spp_ppo_load_rollout; asynchronous on the policy's stream. */
int spp_ppo_rollout_synthetic(spp_ppo* p, spp_population* pop, int agent, int E, int T, int max_ep_len, double done_prob, uint64_t seed,
                              int denormalize_actor_out, int reset_envs, const float* noise_act, const float* noise_env,
                              const float* u_done, const float* noise_reset);
/* one column of the loaded rows as a dense host array: "x", "xn", "act", "raw_obs", "raw_next" [N][ob]; "aacm" [N][ac];
 * "logp", "rew", "done", "end", "adv", "v" [N] */
int spp_ppo_store_download(spp_ppo* p, const char* name, float* host);
/* ReplayBufferAcM.add_buffer (rltoolkit/buffer/replay_buffer.py:284-297) from the store a device rollout left in `store`: environment
 * by environment, every trajectory cut into rollouts at its `end` flags, with the reference's behaviour at rollout joints (quirk 19)
 * and the ring cursor state machine of add_obs / add_timestep (:56-75).  The host walks integers only; the rows move on the device. */
int spp_ring_add_rollout_store(spp_population* p, int agent, spp_ppo* store);
/* device pointers of the reduced gradient vector (n_floats) and the 8 scalar slots, for torch.distributed all_reduce */
int spp_ppo_grad_buffer(spp_ppo* p, void** dev_ptr, int* n_floats, void** scal_ptr);

/* ---- introspection for tests ------------------------------------------------------------------ */
/* copy a named scratch buffer of agent a to host ("xo","xn","xc","xcp","xm","ha1","ha2","ml","hc1_0",...);
 * rows/ld describe the returned dense [rows x ld] block. */
int spp_debug_scratch(spp_population* p, int a, const char* name, float* host, int cap, int* rows, int* ld);
/* tcgen05 building-block self-test: C[128x128] = A * B on one CTA through kind::tf32 tensor-core MMAs.
 * a_mn / b_mn: 0 = operand stored [128][K] (contraction contiguous), 1 = stored [K][128]; split: 0 = one tf32 pass,
 * 1 = three-pass hi/lo split (fp32-accurate).  K a multiple of 4.  Host arrays. */
int spp_umma_selftest(int a_mn, int b_mn, int K, int split, const float* A, const float* B, float* C);
/* the pipelined tensor-core GEMM of the fused kernels, C[M x 256] = A . B (three-pass split), on one CTA.
 * a_km: A stored [M][K] (1) or [K][M] (0); b_km: B stored [256][K] (1) or [K][256] (0); M <= 256, M and K multiples of 4;
 * reps repeats the product inside the kernel; ms_out (optional) receives the kernel time.  Host arrays. */
int spp_umma_gemm_selftest(int a_km, int b_km, int M, int K, int reps, const float* A, const float* B, float* C, float* ms_out);
/* hardware probes behind design decisions (tools/umma_probe.py): mode 1 = K-major A stored with the SWIZZLE_128B_BASE32B byte image,
 * mode 2 = M = 64 accumulator placement in TMEM; A, B [128][32] (tf32-exact values), C [128][128] = TMEM lanes x columns */
int spp_umma_probe(int mode, const float* A, const float* B, float* C);
/* process-wide switch for the 256-wide GEMMs of the fused kernels; all three are device paths:
 *   1 = tcgen05, 3-pass tf32 hi/lo split with per-chunk fp32 drain -- fp32-accurate, the default and the one every parity claim is made on;
 *   0 = FFMA tiles (A/B reference for the tensor-core path);
 *   2 = tcgen05, ONE tf32 pass, the whole K accumulated in TMEM -- the reduced-precision variant (stated tolerance 1e-2 relative on
 *       losses and post-step weights, tests/test_gpu_update_parity.py); never used unless selected here or by SPP_UMMA=2. */
int spp_set_gemm_path(int tensor_cores);
int spp_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, char* name, int name_cap);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
int64_t spp_kernel_launches(void);

#ifdef __cplusplus
}
#endif
#endif /* SPP_RL_B200_H */
