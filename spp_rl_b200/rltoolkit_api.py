"""Host-side mirror of the reference's algorithm classes for the off-policy SPP path (SAC_AcM, DDPG_AcM).

Same class names, constructor keyword arguments and method names as rltoolkit (rltoolkit/acm/off_policy/sac_acm.py,
ddpg_acm.py, off_policy.py; rltoolkit/acm/acm.py; rltoolkit/algorithms/ddpg/ddpg.py, sac/sac.py; rltoolkit/rl.py), so a
reference train script (train/spp_sac_hopper.py:76-112) or the notebook (notebooks/load_and_test.ipynb) runs against it
unchanged apart from the import.  Every hot method delegates to the C ABI (libspp_rl_b200.so); nothing here computes
on the CPU except host glue the reference also keeps on the host (index draws, episode bookkeeping, pickling).

Random streams are the reference's: replay indices come from numpy's global generator (np.random.randint, one draw per
update step, replay_buffer.py:234) and Gaussian noise from torch's global generator in the reference's draw order, so
seeding numpy and torch reproduces the reference's index and noise streams exactly.
"""
import pickle
from collections import OrderedDict

import numpy as np
import torch
from torch.distributions.utils import _standard_normal

from . import envs as _envs
from .population import Population

# defaults of rltoolkit/config.py (constructor keyword defaults are part of the API contract)
DEFAULTS = dict(
    env_name="CartPole-v0", iterations=2000, gamma=0.95, batch_size=200, stats_freq=20, test_episodes=None, return_done=None,
    log_dir=None, use_gpu=False, tensorboard_dir=None, tensorboard_comment="", verbose=1, render=False, debug_mode=True,
    acm_epochs=1, acm_batch_size=128, acm_update_freq=1, acm_ob_idx=None, acm_lr=3e-3, acm_pre_train_samples=1000,
    acm_pre_train_epochs=10, acm_scheduler_step=25, acm_scheduler_gamma=0.5, acm_val_buffer_size=10_000, acm_update_batches=False,
    denormalize_actor_out=False, acm_keep_pretrain=True, acm_critic=False, min_max_denormalize=False, norm_closs=True,
    actor_lr=1e-3, critic_lr=1e-3, tau=0.005, update_batch_size=100, buffer_size=int(1e6), random_frames=100, update_freq=50,
    grad_steps=50, act_noise=0.1, alpha_lr=1e-3, alpha=0.2, pi_update_freq=1, obs_norm=False, max_frames=None,
    unbiased_update=False, custom_loss=0.0,
)
MAX_ABS_OBS_VALUE = 10


def add_rollouts_to_acm_ring(pop, agent, chain_obs, actions_acm, new_rollout_idx):
    """ReplayBufferAcM.add_buffer (rltoolkit/buffer/replay_buffer.py:284-297) on the device ring, including its behaviour
    at rollout joints: the first transition of each new rollout is skipped and the last observation of the previous
    rollout is paired with the second observation of the new one."""
    joints = set(int(j) for j in new_rollout_idx)
    i = 0
    obs_idx = pop.ring_add_obs(agent, chain_obs[i])
    for acm_action in actions_acm:
        i += 1
        next_idx = pop.ring_add_obs(agent, chain_obs[i])
        if i in joints:
            i += 1
            continue
        pop.ring_add_acm_action(agent, acm_action)
        pop.ring_add_timestep(agent, obs_idx, next_idx, None, 0.0, False, False)
        obs_idx = next_idx


class NetView:
    """Stands where the reference keeps an nn.Module (model.actor, model.critic_1, model.acm ...): state_dict /
    load_state_dict with the reference's keys and CPU float32 tensors; parameters live on the device."""

    def __init__(self, pop, net, agent=0):
        self._pop, self._net, self._agent = pop, net, agent

    def state_dict(self):
        return OrderedDict((k, torch.from_numpy(v)) for k, v in self._pop.state_dict(self._net, self._agent).items())

    def load_state_dict(self, sd):
        self._pop.load_state_dict(self._net, sd, agent=self._agent)

    def eval(self):
        return self

    def train(self, mode=True):
        return self


def sampler_permutation(n, dataloader=True):
    """The row order of one pass over DataLoader(dataset, shuffle=True) (acm.py:285, on_policy.py:166) under torch's global
    generator: every new iterator of a DataLoader first draws its base seed from the global generator (_BaseDataLoaderIter), then
    RandomSampler.__iter__ draws the seed of a private generator and permutes with it.  dataloader=False: a bare RandomSampler."""
    if dataloader:
        torch.empty((), dtype=torch.int64).random_()
    seed = int(torch.empty((), dtype=torch.int64).random_().item())
    g = torch.Generator()
    g.manual_seed(seed)
    return torch.randperm(n, generator=g)


class ReplayRing:
    """BufferAcMOffPolicy surface (rltoolkit/buffer/replay_buffer.py:303-401) over one agent's device ring."""

    def __init__(self, pop, agent, size, ob_dim, min_max_denormalize, obs_norm):
        self.pop, self.agent, self.size, self.obs_shape = pop, agent, int(size), ob_dim
        self.min_max_denormalize, self.obs_norm = min_max_denormalize, obs_norm
        self.obs_mean, self.obs_std = torch.zeros(ob_dim), torch.ones(ob_dim)
        self.min_obs = self.max_obs = None
        self.dtype = torch.float32
        self._end = np.zeros(self.size, bool)          # host mirror of the end flags: last_rollout() walks them (replay_buffer.py:170-177)

    def last_end(self, idx):
        end = self._end[idx]
        while not end:
            idx -= 1
            if idx < 0:
                idx = self.current_len - 1
            end = self._end[idx]
        return idx

    def last_rollout(self):
        """BufferAcMOffPolicy.last_rollout (replay_buffer.py:335-383): walk back from the newest timestep to the last end flag, then
        to the end flag before it; the rows in between come from the device ring in one gather."""
        i = self.last_end(self.ts_idx - 1)
        rows = []
        next_end = False
        while not next_end:
            rows.insert(0, i)
            i -= 1
            if i < 0:
                i = self.current_len - 1
            next_end = self._end[i]
        obs, nobs, act, rew, _, aacm = self.pop.ring_sample_batch(self.agent, np.asarray(rows, np.int64))
        end = [bool(self._end[r]) for r in rows]
        end[-1] = True                                                  # Memory.add_rollout (memory.py:262-263)
        return LastRollout(np.concatenate([obs, nobs[-1:]]), act, list(rew), end, aacm)

    def __len__(self):
        return self.pop.ring_state(self.agent)[2]

    @property
    def ts_idx(self):
        return self.pop.ring_state(self.agent)[1]

    @property
    def current_len(self):
        return len(self)

    def reset_idx(self):
        self.pop.ring_reset(self.agent)

    def add_obs(self, obs):
        return self.pop.ring_add_obs(self.agent, torch.as_tensor(obs).cpu().numpy().reshape(-1))

    def add_acm_action(self, acm_action):
        self.pop.ring_add_acm_action(self.agent, np.asarray(acm_action, np.float32).reshape(-1))

    def add_timestep(self, obs_idx, next_obs_idx, action, rew, done, end):
        self._end[self.ts_idx] = bool(end)
        self.pop.ring_add_timestep(self.agent, obs_idx, next_obs_idx, torch.as_tensor(action).cpu().numpy().reshape(-1), rew, done, end)

    def sample_batch(self, batch_size=64, device=None):
        idxs = np.random.randint(0, len(self), batch_size)          # replay_buffer.py:234 -- same stream as the reference
        obs, nobs, act, rew, done, aacm = self.pop.ring_sample_batch(self.agent, idxs)
        batch = [torch.from_numpy(obs), torch.from_numpy(nobs), torch.from_numpy(act), torch.from_numpy(rew), torch.from_numpy(done)]
        if self.obs_norm:
            batch[0], batch[1] = self.normalize(batch[0]), self.normalize(batch[1])
        batch.append(torch.from_numpy(aacm))
        return batch

    def sample_acm_batch(self, batch_size=64):
        idxs = np.random.randint(0, len(self), batch_size)
        obs, nobs, _, _, _, aacm = self.pop.ring_sample_batch(self.agent, idxs)
        return [torch.from_numpy(obs), torch.from_numpy(nobs), torch.from_numpy(aacm)]

    # (de)normalisation of single observations on the host, as the notebook uses it (memory.py:76-127)
    def normalize(self, obs, force=False):
        if not (self.obs_norm or force):
            return obs
        if self.min_max_denormalize:
            if self.min_obs is None and self.max_obs is None:
                return obs
            mean = (self.max_obs + self.min_obs) / 2
            return (obs - mean) / (self.max_obs - mean + 1e-8)
        return torch.clamp((obs - self.obs_mean) / (self.obs_std + 1e-8), -MAX_ABS_OBS_VALUE, MAX_ABS_OBS_VALUE)

    def denormalize(self, obs):
        was_np = isinstance(obs, np.ndarray)
        obs = torch.from_numpy(obs) if was_np else obs
        if self.min_max_denormalize:
            out = (self.max_obs + self.min_obs) / 2 + obs * ((self.max_obs - self.min_obs) / 2)
        else:
            out = (self.obs_std + 1e-8) * obs + self.obs_mean
        return out.numpy() if was_np else out

    def all_obs(self):
        n = len(self)
        return self.pop.ring_sample_batch(self.agent, np.arange(n, dtype=np.int64))[0].astype(np.float64)

    def update_obs_mean_std(self):
        """MetaReplayBuffer.update_obs_mean_std (replay_buffer.py:83-96): whole-buffer mean / population std and the
        1st / 99th percentiles with running widening.  The statistics come from the device (spp_ring_obs_stats: fp64 moments and an
        exact radix select of the order statistics np.percentile interpolates between); only [ob]-sized vectors cross the ABI."""
        if len(self) > 10:
            st = self.pop.ring_obs_stats()
            a = self.agent
            self.obs_mean = torch.tensor(st["mean"][a], dtype=self.dtype)
            self.obs_std = torch.tensor(st["std"][a], dtype=self.dtype)
            cur_max = torch.from_numpy(st["p99"][a]).float()
            cur_min = torch.from_numpy(st["p1"][a]).float()
            if self.max_obs is None or self.min_obs is None:
                self.max_obs, self.min_obs = cur_max, cur_min
            else:
                self.max_obs, self.min_obs = torch.max(cur_max, self.max_obs), torch.min(cur_min, self.min_obs)


class ValidationRing:
    """The ACM validation buffer (ReplayBufferAcM, rltoolkit/buffer/replay_buffer.py:264-300 over MetaReplayBuffer :8-81):
    host memory in the reference as well -- it is filled once in pre_train and only read by calculate_validation_loss,
    whose arithmetic runs on the device (spp_acm_eval_host).  Same cursor state machine as the device ring (A10)."""

    def __init__(self, size, ob_dim, ac_dim):
        self.size = int(size)
        self._obs = np.zeros((self.size, ob_dim), np.float32)
        self._obs_idx = np.zeros(self.size, np.int64)
        self._next_obs_idx = np.zeros(self.size, np.int64)
        self._actions_acm = np.zeros((self.size, ac_dim), np.float32)
        self.reset_idx()

    def reset_idx(self):
        self.obs_idx = self.ts_idx = self.current_len = 0

    def __len__(self):
        return self.current_len

    def add_obs(self, obs):
        self._obs[self.obs_idx] = np.asarray(obs, np.float32).reshape(-1)
        i = self.obs_idx
        self.obs_idx = (self.obs_idx + 1) % self.size
        return i

    def add_timestep(self, obs_idx, next_obs_idx, acm_action):
        self._obs_idx[self.ts_idx], self._next_obs_idx[self.ts_idx] = obs_idx, next_obs_idx
        self._actions_acm[self.ts_idx] = np.asarray(acm_action, np.float32).reshape(-1)
        if next_obs_idx < self.ts_idx:
            self.current_len = self.ts_idx + 1
            self.ts_idx = 0
        else:
            self.ts_idx += 1
        self.current_len = max(self.ts_idx, self.current_len)

    @property
    def obs(self):
        return self._obs[self._obs_idx[: self.current_len]]

    @property
    def next_obs(self):
        return self._obs[self._next_obs_idx[: self.current_len]]

    @property
    def actions_acm(self):
        return self._actions_acm[: self.current_len]


class _Frames:
    """StatsLogger (rltoolkit/stats_logger.py:10-26): frame / rollout counters and the running return, an exponential moving
    average (0.9) of the last finished rollout's return."""

    def __init__(self):
        self.frames = 0
        self.rollouts = 0
        self.running_return = None
        self.test_return = None
        self.time_list = []
        self.stats = []
        self._alpha = 0.9

    def calc_running_return(self, buffer):
        new_mean_return = buffer.average_returns_per_rollout
        if self.running_return is None:
            self.running_return = new_mean_return
        else:
            self.running_return *= self._alpha
            self.running_return += (1 - self._alpha) * new_mean_return
        return self.running_return


class LastRollout:
    """What BufferAcMOffPolicy.last_rollout() hands to the stats logger (replay_buffer.py:335-383): the last FULL rollout of the
    ring as a MemoryAcM -- observation chain (T + 1 rows), state-target actions, rewards, end flags, ACM actions."""

    def __init__(self, obs, actions, rewards, end, actions_acm):
        self.obs_chain, self.actions, self.rewards, self.end, self.actions_acm = obs, actions, rewards, end, actions_acm

    def __len__(self):
        return len(self.rewards)

    @property
    def returns_rollouts(self):                  # memory.py:203-212 (float32 accumulation, as numpy does there)
        returns, ret = [], 0
        for r, e in zip(self.rewards, self.end):
            ret += r
            if e:
                returns.append(ret)
                ret = 0
        return np.array(returns)

    @property
    def rollouts_no(self):
        return sum(self.end)

    @property
    def average_returns_per_rollout(self):
        return sum(self.returns_rollouts) / self.rollouts_no


class _OffPolicyAcM:
    ALGO = "sac"

    def __init__(self, env=None, acm_model="acm", device=0, **kw):
        unknown = set(kw) - set(DEFAULTS) - {"log_all", "evals"}
        if unknown:
            raise TypeError("unexpected keyword arguments: %s" % sorted(unknown))
        c = dict(DEFAULTS)
        c.update(kw)
        self.__dict__.update({k: c[k] for k in DEFAULTS})
        assert self.iterations > 0, "Iteration has to be positive not %r" % (self.iterations,)
        if self.max_frames is not None:
            assert self.max_frames <= self.iterations * self.batch_size, "max_frames should be smaller or equal than iterations * batch_size"
        self.env = env if env is not None else _envs.make(self.env_name)
        self.ob_dim = self.env.observation_space.shape[0]
        self.ac_dim = self.env.action_space.shape[0]
        self.ac_lim = torch.tensor(self.env.action_space.high)
        self.discrete = False
        self.device = torch.device("cpu")      # host tensors; the arithmetic runs on cuda:<device>
        if self.acm_ob_idx is not None and list(self.acm_ob_idx) != list(range(self.ob_dim)):
            raise NotImplementedError("acm_ob_idx subsets are not supported by the device path")
        lims = self.env.observation_space.high                       # acm.py:102-108
        if self.min_max_denormalize:
            lims = 1.0
        elif self.denormalize_actor_out or np.any(np.asarray(lims) == float("inf")):
            lims = MAX_ABS_OBS_VALUE
        self.actor_ac_lim = torch.tensor(lims, dtype=torch.float32)
        self.actor_output_dim = self.ob_dim
        if self.ALGO == "sac":
            self.tau = DEFAULTS["tau"]      # SAC.__init__ swallows tau (quirk 1); DDPG.__init__ resets act_noise (quirk 2)
            self.act_noise = 0.1
        self.max_ep_len = None              # quirk 3: time-limit truncations count as terminals
        self.target_entropy = -float(self.ac_dim)
        self.iteration = 0
        self.stats_logger = _Frames()
        self.obs_mean, self.obs_std = torch.zeros(self.ob_dim), torch.ones(self.ob_dim)
        self.min_obs = self.max_obs = None
        self._device_index = device
        self._pop = self._make_population(acm_model)
        self.replay_buffer = ReplayRing(self._pop, 0, self.buffer_size, self.ob_dim, self.min_max_denormalize, self.obs_norm)
        self.acm_scheduler_epoch = 0
        self._init_weights()
        self.loss = {"actor": 0.0, "acm": 0.0}
        if self.acm_val_buffer_size:                                  # acm.py:140-146 (10 % head room for terminal observations)
            self.acm_val_buffer_size = int(self.acm_val_buffer_size * 1.1)
            self.acm_val_buffer = ValidationRing(self.acm_val_buffer_size, self.ob_dim, self.ac_dim)
            self.loss["acm_val"] = 0.0
        self._cached_acm_action = None

    def _make_population(self, acm_kind):
        pop = Population(
            algo=self.ALGO, ob_dim=self.ob_dim, ac_dim=self.ac_dim, population=1, device=self._device_index, acm_kind=acm_kind,
            acm_critic=self.acm_critic, norm_closs=self.norm_closs, min_max_denormalize=self.min_max_denormalize,
            update_batch_size=self.update_batch_size, acm_batch_size=self.acm_batch_size, buffer_size=self.buffer_size,
            store_actions=True, gamma=self.gamma, tau=self.tau, actor_lr=self.actor_lr, critic_lr=self.critic_lr,
            alpha_lr=self.alpha_lr, acm_lr=self.acm_lr, custom_loss=float(self.custom_loss), alpha=self.alpha,
            target_entropy=self.target_entropy)
        pop.set_limits(np.broadcast_to(self.actor_ac_lim.numpy(), (self.ob_dim,)), self.ac_lim.numpy())
        pop.set_obs_norm(self.obs_norm)          # the fused ring updates gather normalised rows (replay_buffer.py:246-248)
        return pop

    # ------------------------------------------------------------------ nets
    NET_ATTRS = ("actor", "critic", "critic_1", "critic_2", "acm", "actor_targ", "critic_targ", "critic_1_targ", "critic_2_targ")

    def __setattr__(self, name, value):
        """`model.acm = BasicAcM(...)` (notebook cell 24), `model.actor = ...`: the reference's property setters take an nn.Module,
        move it to the device and build a fresh optimiser (acm.py:176-183, ddpg.py:141-157, sac.py:116-136).  Here the module's
        parameters are uploaded to the device path and the optimiser state of that net is reset; an ACM of the other kind
        re-creates the device population.  Anything without a state_dict() is refused -- never silently detached."""
        if name in self.NET_ATTRS and "_pop" in self.__dict__:
            self._assign_net(name, value)
        else:
            object.__setattr__(self, name, value)

    def _assign_net(self, name, module):
        from .modules import acm_kind_of
        if not hasattr(module, "state_dict"):
            raise TypeError("model.%s needs a module with state_dict() (got %s); its parameters move to the device path" % (name, type(module).__name__))
        sd = module.state_dict()
        if name == "acm":
            kind = acm_kind_of(sd, self.ob_dim, self.ac_dim)
            if (kind == "basic") != (self._pop.cfg.acm_kind == 1):
                self._rebuild_population(kind)
            self._pop.load_state_dict("acm", sd)
            self._pop.adam_reset("acm")
            self.acm_scheduler_epoch = 0
            self._pop.set_learning_rates(acm_lr=self.acm_lr)
            return
        if name not in self._net_names() and not name.endswith("_targ"):
            raise AttributeError("%s has no net %r" % (type(self).__name__, name))
        self._pop.load_state_dict(name, sd)
        if not name.endswith("_targ"):
            self._pop.adam_reset(name)
            targ = name + "_targ"
            if name.startswith("critic") or self.ALGO == "ddpg":      # the setters deep-copy the target (SAC has no actor target)
                self._pop.load_state_dict(targ, sd)

    def _rebuild_population(self, acm_kind):
        """Same agent with the other kind of ACM: device memory is laid out per kind, so the population is created anew and every
        other net, the temperature and the statistics are carried over.  Only before any transition is stored."""
        if len(self.replay_buffer) > 0:
            raise RuntimeError("assign model.acm before collecting data: the replay ring lives with the device population")
        old = self._pop
        nets = self._net_names()[:-1] + (["critic_1_targ", "critic_2_targ"] if self.ALGO == "sac" else ["actor_targ", "critic_targ"])
        saved = {net: old.state_dict(net) for net in nets}
        log_alpha = old.alpha(0)[0] if self.ALGO == "sac" else None
        old.close()
        object.__setattr__(self, "_pop", self._make_population(acm_kind))
        for net, sd in saved.items():
            self._pop.load_state_dict(net, sd)
        if log_alpha is not None:
            self._pop.set_log_alpha(log_alpha)
        self.replay_buffer.pop = self._pop
        self._push_stats()

    def _net_names(self):
        return ["actor", "critic_1", "critic_2", "acm"] if self.ALGO == "sac" else ["actor", "critic", "acm"]

    def _init_weights(self):
        from .init import init_state
        seed = int(torch.randint(0, 2 ** 31 - 1, (1,)).item())       # follows torch.manual_seed like nn.Linear's init would
        kind = "basic" if self._pop.cfg.acm_kind == 1 else "acm"
        s0 = init_state(self.ALGO, self.ob_dim, self.ac_dim, seed, kind, self.acm_critic)
        for net in self._net_names():
            self._pop.load_state_dict(net, {k[len(net) + 1:]: v for k, v in s0.items() if k.startswith(net + ".")})
        self._pop.sync_targets()

    def __getattr__(self, name):
        if name in self.NET_ATTRS and "_pop" in self.__dict__:
            return NetView(self.__dict__["_pop"], name)
        raise AttributeError(name)

    @property
    def alpha_value(self):
        return self._pop.alpha(0)[1]

    def _push_stats(self):
        rb = self.replay_buffer
        self._pop.set_norm_stats(None if rb.min_obs is None else rb.min_obs.numpy(), None if rb.max_obs is None else rb.max_obs.numpy(),
                                 rb.obs_mean.numpy(), rb.obs_std.numpy())

    # ------------------------------------------------------------------ the update step
    def _draw_eps(self, G):
        if self.ALGO != "sac":
            return None
        B = self.update_batch_size
        eps = torch.empty(1, G, 2, B, self.ob_dim)
        for g in range(G):          # Normal.rsample order of SAC_AcM.update: target pass first, then the policy pass
            eps[0, g, 0] = _standard_normal((B, self.ob_dim), dtype=torch.float32, device=torch.device("cpu"))
            eps[0, g, 1] = _standard_normal((B, self.ob_dim), dtype=torch.float32, device=torch.device("cpu"))
        return eps.numpy()

    def _store_losses(self, row):
        if self.ALGO == "sac":
            self.loss.update({"critic_1": float(row[0]), "critic_2": float(row[1]), "actor": float(row[2])})
            self.alpha = float(row[6])
            if self.custom_loss:
                self.loss.update({"sac": float(row[3]), "dist": float(row[4])})
        else:
            self.loss.update({"critic": float(row[0]), "actor": float(row[2])})
            if self.custom_loss:
                self.loss.update({"ddpg": float(row[3]), "dist": float(row[4])})

    def update(self, obs, next_obs, action, reward, done, acm_action):
        """SAC_AcM.update / DDPG_AcM.update on one explicit minibatch (sac_acm.py:89-162, ddpg_acm.py:147-201)."""
        f = lambda t: np.ascontiguousarray(torch.as_tensor(t).cpu().numpy(), dtype=np.float32)[None, None]
        d = np.ascontiguousarray(torch.as_tensor(done).cpu().numpy(), dtype=np.int8)[None, None]
        losses = self._pop.update_host(1, f(obs), f(next_obs), f(action), f(reward), d, f(acm_action), eps=self._draw_eps(1))
        self._store_losses(losses[0, 0])

    def update_condition(self):
        return len(self.replay_buffer) > self.update_batch_size and self.stats_logger.frames % self.update_freq == 0

    def acm_update_condition(self):
        return self.iteration > 0 and self.acm_epochs > 0 and self.stats_logger.frames % self.acm_update_freq == 0

    def make_update(self):
        """DDPG_AcM.make_update (ddpg_acm.py:75-85): grad_steps x (sample_batch + update) in ONE fused launch, then the
        ACM regression when its condition holds."""
        if self.unbiased_update:
            self.make_unbiased_update()
        elif self.update_condition():
            G, B, n = self.grad_steps, self.update_batch_size, len(self.replay_buffer)
            idx = np.stack([np.random.randint(0, n, B) for _ in range(G)]).astype(np.int64)[None]
            losses = self._pop.update_ring(G, idx=np.ascontiguousarray(idx), eps=self._draw_eps(G))
            self._store_losses(losses[0, -1])
        if self.acm_update_condition():
            if self.acm_update_batches:
                self.update_acm_batches(self.acm_update_batches)
            else:
                self.update_acm(self.acm_epochs)

    def make_unbiased_update(self):
        """DDPG_AcM.make_unbiased_update (ddpg_acm.py:59-73): the stored next observation stands in for the state-target action.
        The grad_steps minibatches are gathered from the device ring and go through the explicit-batch form in ONE fused launch."""
        if self.update_condition():
            G, B, n = self.grad_steps, self.update_batch_size, len(self.replay_buffer)
            idx = np.concatenate([np.random.randint(0, n, B) for _ in range(G)]).astype(np.int64)
            obs, nobs, _, rew, done, aacm = self._pop.ring_sample_batch(0, idx)
            shp = lambda a: np.ascontiguousarray(a.reshape((1, G, B) + a.shape[1:]))
            losses = self._pop.update_host(G, shp(obs), shp(nobs), shp(nobs), shp(rew), shp(done), shp(aacm), eps=self._draw_eps(G))
            self._store_losses(losses[0, -1])

    def update_acm_batches(self, n_batches):
        n = len(self.replay_buffer)
        idx = np.stack([np.random.randint(0, n, self.acm_batch_size) for _ in range(n_batches)]).astype(np.int64)[None]
        losses = self._pop.acm_update_ring(n_batches, idx=np.ascontiguousarray(idx))
        self.loss["acm"] = float(losses[0].sum() / n_batches)
        if self.acm_val_buffer_size:
            self.loss["acm_val"] = self.calculate_validation_loss()

    def update_acm(self, epochs, pretrain=False):
        """AcMTrainer.update_acm (acm.py:266-303): shuffled epochs over the whole ring, StepLR stepped once per epoch."""
        n, B = len(self.replay_buffer), self.acm_batch_size
        nb = (n + B - 1) // B
        last = n - (nb - 1) * B
        for _ in range(epochs):
            perm = sampler_permutation(n).numpy().astype(np.int64)      # DataLoader(shuffle=True)
            idx = np.zeros((1, nb, B), np.int64)
            idx.reshape(-1)[:n] = perm
            lr = self.acm_lr * self.acm_scheduler_gamma ** (self.acm_scheduler_epoch // self.acm_scheduler_step)
            self._pop.set_learning_rates(acm_lr=lr)
            losses = self._pop.acm_update_ring(nb, idx=idx, last_rows=0 if last == B else last)
            self.loss["acm"] = float(losses[0].sum() / nb)
            self.acm_scheduler_epoch += 1
        if self.acm_val_buffer_size:
            self.loss["acm_val"] = self.calculate_validation_loss()

    def get_val_x_y(self):
        b = self.acm_val_buffer                                       # acm.py:313-327
        return np.concatenate([b.obs, b.next_obs], axis=1), b.actions_acm

    def calculate_validation_loss(self):
        """AcMTrainer.calculate_validation_loss (acm.py:329-343): MSE of the ACM over the whole validation buffer."""
        x, y = self.get_val_x_y()
        assert len(x) > 0, "No validation data. Were the pretrain ran?"
        return float(self._pop.acm_validation_loss(x[None], y[None])[0])

    def collect_initial_batch(self, buffer, samples_no):
        collected = 0
        while collected < samples_no:                                 # acm.py:204-232
            prev_idx = buffer.add_obs(self.env.reset())
            end = False
            while not end:
                action = self.env.action_space.sample()
                obs, _, end, _ = self.env.step(action)
                next_idx = buffer.add_obs(obs)
                buffer.add_timestep(prev_idx, next_idx, action)
                prev_idx = next_idx
                collected += 1
        return buffer

    # ------------------------------------------------------------------ acting
    def initial_act(self, obs):
        action = self.actor_ac_lim * torch.randn(1, self.ob_dim)      # off_policy.py:50-54
        if self.denormalize_actor_out:
            action = self.replay_buffer.denormalize(action)
        return action

    def noise_action(self, obs, act_noise, deterministic=False):
        """DDPG_AcM.noise_action (ddpg_acm.py:40-50); the ACM action computed in the same launch is cached for the
        process_action call that follows in the frame loop."""
        eps = None
        if self.ALGO == "sac" and not deterministic:
            eps = _standard_normal((1, self.ob_dim), dtype=torch.float32, device=torch.device("cpu")).numpy()[None]
        noise = torch.randn(self.actor_output_dim).numpy()[None, None]
        o = torch.as_tensor(obs, dtype=torch.float32).reshape(1, 1, self.ob_dim).numpy()
        tgt, act = self._pop.rollout_step(o, noise, eps, random_phase=0, act_noise=act_noise, obs_norm=False,
                                          denormalize_actor_out=self.denormalize_actor_out)
        self._cached_acm_action = (tgt[0].copy(), act[0, 0].copy())
        return torch.from_numpy(tgt[0])

    def process_action(self, action, obs, pre_train=False):
        """AcMOffPolicy.process_action (off_policy.py:89-106): ACM(cat[obs, action]) and the ring write of the ACM action."""
        a = torch.as_tensor(action, dtype=torch.float32).reshape(1, 1, self.ob_dim).numpy()
        if self._cached_acm_action is not None and np.array_equal(self._cached_acm_action[0], a[0]):
            acm_action = self._cached_acm_action[1]
        else:
            o = torch.as_tensor(obs, dtype=torch.float32).reshape(1, 1, self.ob_dim).numpy()
            _, act = self._pop.rollout_step(o, a, None, random_phase=2, act_noise=0.0, obs_norm=False, denormalize_actor_out=False)
            acm_action = act[0, 0]
        self._cached_acm_action = None
        self.replay_buffer.add_acm_action(acm_action)
        return acm_action

    def process_obs(self, obs):
        return torch.tensor(obs, dtype=torch.float32).unsqueeze(0)

    # ------------------------------------------------------------------ loops (host glue, as in the reference)
    def collect_batch_and_train(self, batch_size):
        collected = 0
        while collected < batch_size:                                 # ddpg.py:182-223
            self.stats_logger.rollouts += 1
            obs = self.process_obs(self.env.reset())
            end = False
            prev_idx = self.replay_buffer.add_obs(obs)
            ep_len = 0
            while not end:
                obs = self.replay_buffer.normalize(obs)
                if self.stats_logger.frames < self.random_frames:
                    action = self.initial_act(obs)
                else:
                    action = self.noise_action(obs, self.act_noise)
                action_proc = self.process_action(action, obs)
                obs, rew, done, _ = self.env.step(action_proc)
                ep_len += 1
                end = done
                done = False if ep_len == self.max_ep_len else done
                obs = self.process_obs(obs)
                next_idx = self.replay_buffer.add_obs(obs)
                self.replay_buffer.add_timestep(prev_idx, next_idx, action, rew, done, end)
                prev_idx = next_idx
                self.stats_logger.frames += 1
                collected += 1
                self.make_update()

    def update_obs_mean_std(self, buffer):
        buffer.update_obs_mean_std()                                  # rl.py:93-112
        self.obs_mean, self.obs_std, self.max_obs, self.min_obs = buffer.obs_mean, buffer.obs_std, buffer.max_obs, buffer.min_obs
        self._push_stats()
        return buffer

    def perform_iteration(self):
        self.collect_batch_and_train(self.batch_size)                 # ddpg.py:159-169
        self.replay_buffer = self.update_obs_mean_std(self.replay_buffer)
        return self.replay_buffer.last_rollout()

    def collect_samples(self):
        collected = 0
        while collected < self.acm_pre_train_samples:                 # off_policy.py:56-87
            obs = torch.tensor(self.env.reset(), dtype=torch.float32).unsqueeze(0)
            end = False
            prev_idx = self.replay_buffer.add_obs(obs)
            while not end:
                acm_action = self.env.action_space.sample()
                self.replay_buffer.add_acm_action(acm_action)
                obs, rew, done, _ = self.env.step(acm_action)
                obs = torch.tensor(obs, dtype=torch.float32).unsqueeze(0)
                end = done
                next_idx = self.replay_buffer.add_obs(obs)
                self.replay_buffer.add_timestep(prev_idx, next_idx, obs, rew, done, end)
                prev_idx = next_idx
                collected += 1

    def pre_train(self):
        if self.acm_val_buffer_size:                                  # acm.py:234-244
            self.acm_val_buffer = self.collect_initial_batch(self.acm_val_buffer, self.acm_val_buffer_size)
        self.collect_samples()
        self.update_acm(epochs=self.acm_pre_train_epochs, pretrain=True)
        self.update_obs_mean_std(self.replay_buffer)
        if not self.acm_keep_pretrain:
            self.replay_buffer.reset_idx()

    def train(self, iterations=None):
        if iterations:
            self.iterations += iterations
        import time
        buffer = None
        while self.iteration < self.iterations:                       # rl.py:197-235
            t0 = time.time()
            buffer = self.perform_iteration()
            self.stats_logger.time_list.append(time.time() - t0)
            running_return = self.stats_logger.calc_running_return(buffer)
            if self.return_done is not None and running_return >= self.return_done:
                break
            if self.iteration % self.stats_freq == 0:
                self.logs_after_iteration(buffer)
            self.iteration += 1
            if self.max_frames is not None and self.max_frames < self.stats_logger.frames:
                break
        if buffer is not None:
            self.logs_after_iteration(buffer, done=True)

    def logs_after_iteration(self, buffer, done=False):
        """rl.py:320-341 without the TensorBoard writer (observability is outside the hot path): the periodic test() and the
        (iteration, running return) history."""
        if self.test_episodes is not None:
            self.stats_logger.test_return = self.test()
        self.stats_logger.stats.append([self.iteration, self.stats_logger.running_return])
        self.stats_logger.time_list = []

    def test(self, episodes=None):
        episodes = self.test_episodes if episodes is None else episodes
        returns = []
        for _ in range(episodes or 1):                                # ddpg.py:385-410
            obs, done, ep_ret = self.env.reset(), False, 0.0
            while not done:
                obs = self.replay_buffer.normalize(self.process_obs(obs))
                action = self.noise_action(obs, act_noise=0, deterministic=True)
                obs, r, done, _ = self.env.step(self.process_action(action, obs))
                ep_ret += r
            returns.append(ep_ret)
        return float(np.mean(returns))

    # ------------------------------------------------------------------ persistence (pickle layout of rl.py:263-301)
    def collect_params_dict(self):
        d = {}
        for net in self._net_names():
            d[net] = getattr(self, net).state_dict()
        rb = self.replay_buffer
        d.update({"obs_mean": rb.obs_mean, "obs_std": rb.obs_std, "min_obs": rb.min_obs, "max_obs": rb.max_obs})
        acm = d.pop("acm")
        d["acm"] = acm
        return d

    def apply_params_dict(self, params_dict):
        for net in self._net_names():
            getattr(self, net).load_state_dict(params_dict[net])       # targets are NOT refreshed by load() (quirk 8)
        rb = self.replay_buffer
        rb.obs_mean, rb.obs_std = params_dict["obs_mean"], params_dict["obs_std"]
        rb.min_obs, rb.max_obs = params_dict["min_obs"], params_dict["max_obs"]
        self.obs_mean, self.obs_std, self.min_obs, self.max_obs = rb.obs_mean, rb.obs_std, rb.min_obs, rb.max_obs
        self._push_stats()

    def save(self, path):
        with open(path, "wb") as f:
            pickle.dump(self.collect_params_dict(), f)

    def load(self, path):
        with open(path, "rb") as f:
            self.apply_params_dict(pickle.load(f))

    def close(self):
        self._pop.close()


class SAC_AcM(_OffPolicyAcM):
    ALGO = "sac"


class DDPG_AcM(_OffPolicyAcM):
    ALGO = "ddpg"


def __getattr__(name):      # `from spp_rl_b200.rltoolkit_api import PPO_AcM` (the on-policy class lives in rltoolkit_ppo.py)
    if name == "PPO_AcM":
        from .rltoolkit_ppo import PPO_AcM
        return PPO_AcM
    raise AttributeError(name)
