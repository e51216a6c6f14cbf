"""Host-side mirror of the reference's on-policy SPP class, PPO_AcM (rltoolkit/acm/on_policy.py:17-216 over
rltoolkit/algorithms/a2c/a2c.py, ppo/ppo.py, rltoolkit/acm/acm.py and rltoolkit/rl.py).

Same constructor keyword arguments, method names, loss keys and pickle layout as the reference, so its PPO_AcM script
(rltoolkit/acm/on_policy.py:219-247) runs against it unchanged apart from the import.  The arithmetic -- Actor.act,
process_action through the ACM, critic fit, q-values + GAE, advantage normalisation, clipped-ratio epochs, ACM regression,
buffer statistics -- runs on the device behind the C ABI (PpoPolicy for the 64-wide actor / critic, Population for the ACM and
its replay ring); what stays here is what the reference also keeps on the host: the environment loop, the rollout memory as
Python lists, the sampler's permutations and pickling.

Random streams follow the reference: the policy noise is drawn from torch's global generator with the call Normal.sample makes
(one standard-normal tensor of the action's shape per step), and the minibatch permutations are drawn the way
torch.utils.data.RandomSampler draws them (a seed from the global generator per epoch, then randperm on a private generator).
"""
import pickle
from collections import OrderedDict

import numpy as np
import torch

from . import envs as _envs
from .population import Population
from .ppo import PpoPolicy
from .rltoolkit_api import DEFAULTS, MAX_ABS_OBS_VALUE, NetView, ReplayRing, ValidationRing, _Frames, _OffPolicyAcM, add_rollouts_to_acm_ring, sampler_permutation

# rltoolkit/config.py: A2C / PPO blocks (constructor keyword defaults are part of the API contract)
PPO_DEFAULTS = dict(DEFAULTS)
PPO_DEFAULTS.update(actor_lr=3e-3, critic_lr=3e-4, critic_num_target_updates=10, num_critic_updates_per_target=10, normalize_adv=True,
                    obs_norm_alpha=None, epsilon=0.2, gae_lambda=0.95, kl_div_threshold=0.15, max_ppo_epochs=50, ppo_batch_size=1000,
                    entropy_coef=0.0)
for _k in ("tau", "update_batch_size", "buffer_size", "random_frames", "update_freq", "grad_steps", "act_noise", "alpha_lr", "alpha",
           "pi_update_freq", "unbiased_update", "acm_critic"):
    PPO_DEFAULTS.pop(_k)


class _PolicyNetView:
    """model.actor / model.critic: state_dict / load_state_dict with the reference's keys; parameters live on the device."""

    def __init__(self, pol, net):
        self._pol, self._net = pol, net

    def state_dict(self):
        return OrderedDict((k, torch.from_numpy(v)) for k, v in self._pol.state_dict(self._net).items())

    def load_state_dict(self, sd):
        self._pol.load_state_dict(self._net, sd)

    def eval(self):
        return self

    def train(self, mode=True):
        return self

    def act(self, obs, deterministic=False):
        """Actor.act (rltoolkit/basic_model.py:32-51) on already-normalised observations [E, ob] -> (action, log-prob); the noise is
        the one standard-normal tensor Normal.sample draws from torch's global generator."""
        if self._net != "actor":
            raise AttributeError("critic has no act()")
        o = torch.as_tensor(obs, dtype=torch.float32).reshape(-1, self._pol.ob_dim)
        noise = torch.zeros_like(o) if deterministic else torch.empty_like(o).normal_()
        action, logp, _ = self._pol.act_normalized(o.numpy(), noise.numpy())
        return torch.from_numpy(action), torch.from_numpy(logp)


class AcmReplayRing(ReplayRing):
    """ReplayBufferAcM surface (rltoolkit/buffer/replay_buffer.py:264-300) over the agent's device ring."""

    def add_timestep(self, obs_idx, next_obs_idx, acm_action):
        self.pop.ring_add_acm_action(self.agent, np.asarray(acm_action, np.float32).reshape(-1))
        self.pop.ring_add_timestep(self.agent, obs_idx, next_obs_idx, None, 0.0, False, False)

    def add_buffer(self, memory):
        chain = torch.cat(memory._obs).numpy()
        add_rollouts_to_acm_ring(self.pop, self.agent, chain, np.asarray(memory.actions_acm, np.float32), memory._new_rollout_idx)


class RolloutMemory:
    """MemoryAcM (rltoolkit/buffer/memory.py:130-315): one batch of full rollouts as a chain of observations (terminal
    observations included) plus per-timestep lists.  `_new_rollout_idx` holds the chain positions where a new rollout starts."""

    def __init__(self, min_obs=None, max_obs=None, obs_mean=None, obs_std=None, min_max_denormalize=False, device=None, alpha=None):
        # device / alpha: accepted for the notebook's MemoryAcM(obs_mean=..., device=..., alpha=..., ...) call (cell 44); the EMA
        # statistics update they belong to (memory.py:283-302) is not on the SPP path (the replay ring's statistics are used)
        self._obs, self._actions, self._action_logprobs, self._rewards, self._done, self._end = [], [], [], [], [], []
        self._new_rollout_idx, self.actions_acm = [], []
        self.current_len = 0
        self.min_obs, self.max_obs, self.obs_mean, self.obs_std = min_obs, max_obs, obs_mean, obs_std
        self.min_max_denormalize = min_max_denormalize

    def __len__(self):
        return len(self._actions)

    def add_obs(self, obs):
        self._obs.append(obs)
        self.current_len += 1
        return self.current_len - 1

    def add_timestep(self, obs_idx, next_obs_idx, action, action_logprobs, rew, done, end):
        self._actions.append(action); self._action_logprobs.append(action_logprobs)
        self._rewards.append(rew); self._done.append(done); self._end.append(end)

    def add_acm_action(self, acm_action):
        self.actions_acm.append(acm_action)

    def end_rollout(self):
        self._new_rollout_idx.append(self.current_len)

    def _rows(self, first):
        joints = set(self._new_rollout_idx)
        if first:       # obs: every chain entry that is not the last one of its rollout
            return [i for i in range(self.current_len) if (i + 1) not in joints]
        return [i for i in range(1, self.current_len) if i not in joints]

    @property
    def obs(self):
        return torch.cat([self._obs[i] for i in self._rows(True)])

    @property
    def next_obs(self):
        return torch.cat([self._obs[i] for i in self._rows(False)])

    actions = property(lambda self: self._actions)
    action_logprobs = property(lambda self: self._action_logprobs)
    rewards = property(lambda self: self._rewards)
    done = property(lambda self: self._done)
    end = property(lambda self: self._end)

    def normalize(self, obs, force=False):       # memory.py:76-87
        if self.min_max_denormalize:
            if self.min_obs is None and self.max_obs is None:
                return obs
            mean = (self.max_obs + self.min_obs) / 2
            return (obs - mean) / (self.max_obs - mean + 1e-8)
        if self.obs_std is None and self.obs_mean is None:
            return obs
        return torch.clamp((obs - self.obs_mean) / (self.obs_std + 1e-8), -MAX_ABS_OBS_VALUE, MAX_ABS_OBS_VALUE)


MemoryAcM = RolloutMemory      # the reference's name (rltoolkit.buffer.MemoryAcM), as the notebook imports it


class PPO_AcM:
    NET_ATTRS = ("actor", "critic", "acm")
    KW = PPO_DEFAULTS
    A2C = False      # A2C_AcM below: full-batch policy-gradient step instead of the clipped-ratio epochs

    def __init__(self, env=None, acm_model="acm", device=0, **kw):
        unknown = set(kw) - set(self.KW) - {"log_all", "evals", "max_frames", "vector_envs"}
        if unknown:
            raise TypeError("unexpected keyword arguments: %s" % sorted(unknown))
        c = dict(PPO_DEFAULTS)
        c.update(kw)
        self.__dict__.update({k: c[k] for k in PPO_DEFAULTS})
        # extension (not in the reference): vector_envs = E > 0 runs collect_batch as ONE device launch over E synthetic environments
        # (batch_size // E steps each) instead of one Python step per frame; only with the offline synthetic stand-in environment
        self.vector_envs = int(c.get("vector_envs", 0) or 0)
        assert self.iterations > 0, "Iteration has to be positive not %r" % (self.iterations,)
        if not (self.min_max_denormalize and self.denormalize_actor_out):
            raise NotImplementedError("the device path covers the published SPP-PPO setting: min_max_denormalize=True, denormalize_actor_out=True")
        if self.acm_ob_idx is not None:
            raise NotImplementedError("acm_ob_idx subsets are not supported by the device path")
        self.ppo_epsilon = self.epsilon
        self.env = env if env is not None else _envs.make(self.env_name)
        self.ob_dim = self.env.observation_space.shape[0]
        self.ac_dim = self.env.action_space.shape[0]
        self.ac_lim = torch.tensor(self.env.action_space.high)
        self.discrete = False
        self.device = torch.device("cpu")
        self.max_ep_len = self.env._max_episode_steps                 # rl.py:185
        lims = self.env.observation_space.high                        # acm.py:102-108
        if self.min_max_denormalize:
            lims = 1.0
        elif self.denormalize_actor_out or np.any(np.asarray(lims) == float("inf")):
            lims = MAX_ABS_OBS_VALUE
        self.actor_ac_lim = torch.tensor(lims, dtype=torch.float32)
        self.actor_output_dim = self.ob_dim
        self.normalized_buffer = self.denormalize_actor_out
        self.iteration = 0
        self.kl_div_updates_counter = 0
        self.stats_logger = _Frames()
        self.obs_mean, self.obs_std = torch.zeros(self.ob_dim), torch.ones(self.ob_dim)      # rl.py:389-405 with obs_norm
        self.min_obs = self.max_obs = None
        self.buffer = None
        self.buffer_size = int(self.acm_pre_train_samples * 1.1)      # acm.py:125-126
        self._device_index = device
        self._pop = self._make_population(acm_model)
        self._pol = PpoPolicy(self.ob_dim, self.ac_dim, max_rows=self.batch_size + self.max_ep_len,
                              max_batch_rows=(self.batch_size + self.max_ep_len) if self.A2C else self.ppo_batch_size,
                              device=device, min_max_denormalize=self.min_max_denormalize, norm_closs=self.norm_closs, gamma=self.gamma,
                              gae_lambda=self.gae_lambda, ppo_epsilon=self.ppo_epsilon, entropy_coef=self.entropy_coef,
                              custom_loss=float(self.custom_loss), actor_lr=self.actor_lr, critic_lr=self.critic_lr)
        self._pol.set_limits(self.actor_ac_lim.numpy())
        self._pol.set_actor_mode(plain_ppo=not self.custom_loss, a2c=self.A2C)      # custom_loss == 0: PPO / A2C.update_actor (on_policy.py:88-98)
        self.replay_buffer = AcmReplayRing(self._pop, 0, self.buffer_size, self.ob_dim, self.min_max_denormalize, self.obs_norm)
        self.acm_scheduler_epoch = 0
        self._init_weights()
        self.loss = {"actor": 0.0, "critic": 0.0, "acm": 0.0, "policy": 0.0, "dist": 0.0}
        if self.acm_val_buffer_size:                                  # acm.py:140-146
            self.acm_val_buffer_size = int(self.acm_val_buffer_size * 1.1)
            self.acm_val_buffer = ValidationRing(self.acm_val_buffer_size, self.ob_dim, self.ac_dim)
            self.loss["acm_val"] = 0.0
        self._push_stats()

    # ------------------------------------------------------------------ nets
    def _init_weights(self):
        """nn.Linear's default initialisation (uniform +-1/sqrt(fan_in) for weight and bias) from torch's global generator,
        Actor.log_scale = -1.34 (rltoolkit/basic_model.py:7-21,62-68); the ACM as in the off-policy classes."""
        from .init import init_state

        def linear(o, i):
            b = 1.0 / np.sqrt(i)
            return (torch.empty(o, i).uniform_(-b, b), torch.empty(o).uniform_(-b, b))
        for net, out in (("actor", self.ob_dim), ("critic", 1)):
            sd = {}
            for name, o, i in (("fc1", 64, self.ob_dim), ("fc2", 64, 64), ("fc3", out, 64)):
                sd[name + ".weight"], sd[name + ".bias"] = linear(o, i)
            if net == "actor":
                sd["log_scale"] = -1.34 * torch.ones(self.ob_dim)
            self._pol.load_state_dict(net, sd)
        seed = int(torch.randint(0, 2 ** 31 - 1, (1,)).item())
        kind = "basic" if self._pop.cfg.acm_kind == 1 else "acm"
        s0 = init_state("ddpg", self.ob_dim, self.ac_dim, seed, kind, True)
        self._pop.load_state_dict("acm", {k[4:]: v for k, v in s0.items() if k.startswith("acm.")})

    def _make_population(self, acm_kind):
        pop = Population(algo="ddpg", ob_dim=self.ob_dim, ac_dim=self.ac_dim, population=1, device=self._device_index, acm_kind=acm_kind,
                         acm_critic=True, norm_closs=self.norm_closs, min_max_denormalize=self.min_max_denormalize,
                         update_batch_size=64, acm_batch_size=self.acm_batch_size, buffer_size=self.buffer_size,
                         store_actions=False, acm_lr=self.acm_lr)
        pop.set_limits(np.broadcast_to(self.actor_ac_lim.numpy(), (self.ob_dim,)), self.ac_lim.numpy())
        return pop

    def __getattr__(self, name):
        if name in ("actor", "critic") and "_pol" in self.__dict__:
            return _PolicyNetView(self.__dict__["_pol"], name)
        if name == "acm" and "_pop" in self.__dict__:
            return NetView(self.__dict__["_pop"], "acm")
        raise AttributeError(name)

    def __setattr__(self, name, value):
        """`model.acm = BasicAcM(...)`, `model.actor = ...` as in the reference's property setters (acm.py:176-183, a2c.py): the module's
        parameters move to the device path with a fresh optimiser state; objects without state_dict() are refused."""
        if name in self.NET_ATTRS and "_pop" in self.__dict__ and "_pol" in self.__dict__:
            if not hasattr(value, "state_dict"):
                raise TypeError("model.%s needs a module with state_dict() (got %s)" % (name, type(value).__name__))
            sd = value.state_dict()
            if name == "acm":
                from .modules import acm_kind_of
                kind = acm_kind_of(sd, self.ob_dim, self.ac_dim)
                if (kind == "basic") != (self._pop.cfg.acm_kind == 1):
                    if len(self.replay_buffer) > 0:
                        raise RuntimeError("assign model.acm before collecting data: the ACM replay ring lives with the device population")
                    self._pop.close()
                    object.__setattr__(self, "_pop", self._make_population(kind))
                    self.replay_buffer.pop = self._pop
                    self._push_stats()
                self._pop.load_state_dict("acm", sd)
                self._pop.adam_reset("acm")
                self.acm_scheduler_epoch = 0
                self._pop.set_learning_rates(acm_lr=self.acm_lr)
            else:
                self._pol.load_state_dict(name, sd)
                self._pol.adam_reset(name)
        else:
            object.__setattr__(self, name, value)

    def _push_stats(self):
        mn = None if self.min_obs is None else self.min_obs.numpy()
        mx = None if self.max_obs is None else self.max_obs.numpy()
        self._pol.set_norm_stats(mn, mx, self.obs_mean.numpy(), self.obs_std.numpy())
        self._pop.set_norm_stats(mn, mx, self.obs_mean.numpy(), self.obs_std.numpy())

    # ------------------------------------------------------------------ ACM side: the same host methods as the off-policy classes
    update_acm = _OffPolicyAcM.update_acm
    update_acm_batches = _OffPolicyAcM.update_acm_batches
    get_val_x_y = _OffPolicyAcM.get_val_x_y
    calculate_validation_loss = _OffPolicyAcM.calculate_validation_loss
    collect_initial_batch = _OffPolicyAcM.collect_initial_batch

    def collect_samples(self):
        self.replay_buffer = self.collect_initial_batch(self.replay_buffer, self.acm_pre_train_samples)      # acm.py:196-202

    def update_obs_mean_std(self, buffer):
        buffer.update_obs_mean_std()                                  # rl.py:93-112
        self.obs_mean, self.obs_std, self.max_obs, self.min_obs = buffer.obs_mean, buffer.obs_std, buffer.max_obs, buffer.min_obs
        self._push_stats()
        return buffer

    def pre_train(self):
        if self.acm_val_buffer_size:                                  # acm.py:234-244
            self.acm_val_buffer = self.collect_initial_batch(self.acm_val_buffer, self.acm_val_buffer_size)
        self.collect_samples()
        self.update_acm(epochs=self.acm_pre_train_epochs, pretrain=True)
        self.update_obs_mean_std(self.replay_buffer)
        if not self.acm_keep_pretrain:
            self.replay_buffer.reset_idx()

    # ------------------------------------------------------------------ acting (P1)
    def process_obs(self, obs):
        return torch.tensor(obs, dtype=torch.float32).unsqueeze(0)

    def _act(self, obs, deterministic=False):
        """buffer.normalize -> Actor.act -> process_action (a2c.py:160-163, on_policy.py:34-53): the actor sees the normalised
        observation, and so does the ACM next to the denormalised state target (quirk 18)."""
        if deterministic:
            noise = np.zeros((1, self.ob_dim), np.float32)
        else:
            noise = torch.empty(1, self.ob_dim).normal_().numpy()     # Normal.sample(): one standard-normal tensor per step
        o = np.asarray(obs, np.float32).reshape(1, self.ob_dim)
        action, logp, target = self._pol.act(o, noise, self.denormalize_actor_out)
        _, acm_action = self._pop.rollout_step(o[None], action[None], None, random_phase=2, obs_norm=True,      # 2: the target is given
                                               denormalize_actor_out=self.denormalize_actor_out)
        return torch.from_numpy(action), torch.from_numpy(logp), acm_action[0, 0]

    def process_action(self, action, obs, *args, **kwargs):
        """AcMOnPolicyTrainer.process_action (on_policy.py:34-53) as the notebook calls it: `obs` is ALREADY normalised by the caller
        (buffer.normalize), `action` is the actor's output; the ACM sees cat[obs, denormalised action] (quirk 18)."""
        a = torch.as_tensor(action, dtype=torch.float32).reshape(1, -1, self.ob_dim).numpy()
        o = torch.as_tensor(obs, dtype=torch.float32).reshape(1, -1, self.ob_dim).numpy()
        _, acm_action = self._pop.rollout_step(o, a, None, random_phase=2, obs_norm=False, denormalize_actor_out=self.denormalize_actor_out)
        acm_action = acm_action[0, 0]
        if self.buffer is not None:
            self.buffer.add_acm_action(acm_action)
        return acm_action

    def collect_batch(self, buffer):
        start = len(buffer)
        while len(buffer) < self.batch_size:                          # a2c.py:141-180
            self.stats_logger.rollouts += 1
            obs = self.process_obs(self.env.reset())
            end = False
            prev_idx = buffer.add_obs(obs)
            ep_len = 0
            while not end:
                action, logp, acm_action = self._act(obs[0].numpy())
                buffer.add_acm_action(acm_action)
                obs, rew, done, _ = self.env.step(acm_action)
                ep_len += 1
                end = done
                done = False if ep_len == self.max_ep_len else done
                obs = self.process_obs(obs)
                next_idx = buffer.add_obs(obs)
                buffer.add_timestep(prev_idx, next_idx, action, logp, rew, done, end)
                prev_idx = next_idx
            buffer.end_rollout()
        self.stats_logger.frames += len(buffer) - start
        return buffer

    # ------------------------------------------------------------------ the update (P2-P6)
    def _load(self, buffer):
        end = np.asarray(buffer.end, np.float32)
        stops = np.nonzero(end)[0]
        starts = np.concatenate([[0], stops[:-1] + 1]).astype(np.int64)
        self._pol.load_rollout(buffer.obs.numpy(), buffer.next_obs.numpy(), torch.cat(buffer.actions).numpy(),
                               torch.cat(buffer.action_logprobs).numpy(), np.asarray(buffer.rewards, np.float32),
                               np.asarray(buffer.done, np.float32), end, starts, (stops + 1 - starts).astype(np.int64))

    def update_critic(self, buffer):
        """A2C.update_critic (a2c.py:182-221) + PPO's GAE advantages (ppo.py calculate_q_val); returns the advantages."""
        self._load(buffer)
        self.loss["critic"] = self._pol.update_critic(self.critic_num_target_updates, self.num_critic_updates_per_target)
        return torch.from_numpy(self._pol.advantages())

    def update_actor(self, advantages, buffer):
        """on_policy.py:88-98: PPO_AcM.update_actor_acm when custom_loss != 0, otherwise plain PPO.update_actor (ppo.py:152-192) --
        the same device epochs in the mode set at construction (spp_ppo_set_actor_mode); the plain form reports the loss sums
        `actor`, `entropy`, `sum` without the division by the epoch count."""
        self.update_actor_acm(advantages, buffer)

    def update_actor_acm(self, advantages, buffer):
        """PPO_AcM.update_actor_acm (on_policy.py:164-216); `advantages` are the ones update_critic left on the device."""
        if self.normalize_adv:
            self._pol.normalize_adv()
        n = len(buffer)
        state = torch.get_rng_state()
        perms = torch.stack([sampler_permutation(n) for _ in range(self.max_ppo_epochs)]).numpy()
        losses, epochs, kl = self._pol.update_actor(perms, self.ppo_batch_size, self.kl_div_threshold, self.max_ppo_epochs)
        torch.set_rng_state(state)                                    # the reference draws one sampler seed per epoch it actually runs
        for _ in range(2 * epochs):                                   # (DataLoader base seed + sampler seed) per epoch
            torch.empty((), dtype=torch.int64).random_()
        if self.custom_loss:
            self.loss.update({k: float(v) for k, v in losses.items()})
        else:
            self.loss.update({"actor": float(losses["actor"]), "entropy": float(losses["entropy"]), "sum": float(losses["policy"])})
        self.kl_div_updates_counter += min(epochs + 1, self.max_ppo_epochs)      # the reference adds i + 1, i = loop index at exit
        self.last_kl = kl

    def perform_iteration_device(self):
        """perform_iteration (on_policy.py:55-86) with every stage on the device: vectorised rollout into the [T][E] store
        (spp_ppo_rollout_synthetic) -> critic fit -> GAE -> advantage normalisation -> actor epochs -> add_buffer
        (spp_ring_add_rollout_store) -> ACM update -> statistics.  The permutations of the actor epochs still come from torch's
        generator as the reference's DataLoader draws them."""
        if not isinstance(self.env, _envs.SyntheticControl):
            raise RuntimeError("vector_envs needs the synthetic stand-in environment (the device loop has no MuJoCo)")
        E = self.vector_envs
        T = max(1, -(-self.batch_size // E))
        if E * T > self.batch_size + self.max_ep_len:
            raise ValueError("vector_envs * steps exceeds the policy's row capacity (batch_size + max_ep_len)")
        self._iter_seed = getattr(self, "_iter_seed", int(torch.randint(0, 2 ** 31 - 1, (1,)).item())) + 1
        self._pol.rollout_synthetic(self._pop, E, T, max_ep_len=self.max_ep_len, done_prob=0.004, seed=self._iter_seed,
                                    denormalize_actor_out=self.denormalize_actor_out, reset_envs=(self.iteration == 0))
        n = E * T
        self.stats_logger.frames += n
        self.loss["critic"] = self._pol.update_critic(self.critic_num_target_updates, self.num_critic_updates_per_target)
        self._pol.advantages(want_host=False)
        if self.normalize_adv:
            self._pol.normalize_adv()
        perms = torch.stack([sampler_permutation(n) for _ in range(self.max_ppo_epochs)]).numpy()
        losses, epochs, kl = self._pol.update_actor(perms, self.ppo_batch_size, self.kl_div_threshold, self.max_ppo_epochs)
        self.loss.update({k: float(v) for k, v in losses.items()})
        self.kl_div_updates_counter += min(epochs + 1, self.max_ppo_epochs)
        self.last_kl = kl
        self._pop.ring_add_rollout_store(0, self._pol)
        if self.acm_update_freq and self.iteration % self.acm_update_freq == 0:
            if self.acm_update_batches:
                self.update_acm_batches(self.acm_update_batches)
            else:
                self.update_acm(self.acm_epochs)
        if self.denormalize_actor_out:
            self.replay_buffer = self.update_obs_mean_std(self.replay_buffer)
        return None

    def perform_iteration(self):
        if self.vector_envs:
            return self.perform_iteration_device()
        self.buffer = RolloutMemory(self.min_obs, self.max_obs, self.obs_mean, self.obs_std, self.min_max_denormalize)
        self.collect_batch(self.buffer)                               # on_policy.py:55-86
        advantages = self.update_critic(self.buffer)
        self.update_actor(advantages, self.buffer)
        self.replay_buffer.add_buffer(self.buffer)
        if self.acm_update_freq and self.iteration % self.acm_update_freq == 0:
            if self.acm_update_batches:
                self.update_acm_batches(self.acm_update_batches)
            else:
                self.update_acm(self.acm_epochs)
        if self.denormalize_actor_out:
            self.replay_buffer = self.update_obs_mean_std(self.replay_buffer)
        return self.buffer

    def train(self, iterations=None):
        if iterations:
            self.iterations += iterations
        while self.iteration < self.iterations:                       # rl.py:197-235
            self.perform_iteration()
            self.iteration += 1
            if getattr(self, "max_frames", None) is not None and self.max_frames < self.stats_logger.frames:
                break

    def test(self, episodes=None):
        episodes = self.test_episodes if episodes is None else episodes
        ep_ret = 0.0
        for _ in range(episodes or 1):                                # a2c.py:325-350
            obs, done, ep_ret = self.env.reset(), False, 0.0
            while not done:
                _, _, acm_action = self._act(np.asarray(obs, np.float32), deterministic=True)
                obs, r, done, _ = self.env.step(acm_action)
                ep_ret += r
        return float(np.mean(ep_ret))                                 # the reference averages the LAST episode's return only

    # ------------------------------------------------------------------ persistence (rl.py:263-301, on_policy.py:144-151)
    def collect_params_dict(self):
        return {"actor": self.actor.state_dict(), "critic": self.critic.state_dict(), "obs_mean": self.obs_mean, "obs_std": self.obs_std,
                "min_obs": self.min_obs, "max_obs": self.max_obs, "acm": self.acm.state_dict()}

    def apply_params_dict(self, params_dict):
        self.actor.load_state_dict(params_dict["actor"])
        self.critic.load_state_dict(params_dict["critic"])
        self.obs_mean, self.obs_std = params_dict["obs_mean"], params_dict["obs_std"]
        self.min_obs, self.max_obs = params_dict["min_obs"], params_dict["max_obs"]
        self.acm.load_state_dict(params_dict["acm"])
        rb = self.replay_buffer
        rb.obs_mean, rb.obs_std, rb.min_obs, rb.max_obs = self.obs_mean, self.obs_std, self.min_obs, self.max_obs
        self._push_stats()

    def save(self, path):
        with open(path, "wb") as f:
            pickle.dump(self.collect_params_dict(), f)

    def load(self, path):
        with open(path, "rb") as f:
            self.apply_params_dict(pickle.load(f))

    def close(self):
        self._pol.close()
        self._pop.close()



A2C_DEFAULTS = {k: v for k, v in PPO_DEFAULTS.items() if k not in ("epsilon", "gae_lambda", "kl_div_threshold", "max_ppo_epochs", "ppo_batch_size",
                                                                     "entropy_coef")}


class A2C_AcM(PPO_AcM):
    """rltoolkit.acm.on_policy.A2C_AcM (on_policy.py:133-155 over AcMOnPolicyTrainer :17-131 and A2C, a2c.py): the same rollout, critic
    fit, ACM feed and statistics as PPO_AcM (which in the reference DERIVES from this class), with A2C's advantages q - V(s)
    (a2c.py:227-245) and ONE full-batch policy-gradient step per iteration:
      custom_loss != 0 -> update_actor_acm (on_policy.py:100-124): loss = mean(-logp * adv) + custom_loss * MSE(actions, next_obs), where the
                          distance term has no gradient path (Actor.act samples, basic_model.py:47) and the optimiser steps on gradients
                          that are NEVER zeroed -- they accumulate over the iterations;
      custom_loss == 0 -> A2C.update_actor (a2c.py:267-285), which does zero them."""
    KW = A2C_DEFAULTS
    A2C = True

    def update_actor(self, advantages, buffer):
        if self.custom_loss:
            self.update_actor_acm(advantages, buffer)
        else:
            loss, _ = self._pol.a2c_actor_step(accumulate=False, normalize_adv=self.normalize_adv)
            self.loss["actor"] = loss

    def update_actor_acm(self, advantages, buffer):
        loss, dist_sum = self._pol.a2c_actor_step(accumulate=True, normalize_adv=self.normalize_adv)
        dist = dist_sum / (len(buffer) * self.ob_dim)
        self.loss.update({"actor": loss, "dist": dist, "policy": loss + self.custom_loss * dist})

    def perform_iteration_device(self):
        raise NotImplementedError("vector_envs (device-resident iterations) is built for PPO_AcM")
