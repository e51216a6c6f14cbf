"""Host-side handle on a device-resident population of SPP agents (thin wrapper over the C ABI).

One `Population` = P independent agents with identical shapes and hyper-parameters -- what the
reference runs as P OS processes (train/spp_sac_hopper.py:115, rltoolkit/evals.py:86-103).
All arithmetic happens in libspp_rl_b200.so; this file only marshals numpy / torch buffers.
"""
import ctypes as C
from collections import OrderedDict

import numpy as np

from . import _lib
from ._lib import Config, SppError, check


def _ptr(arr, ctype):
    """ctypes pointer to a numpy array or torch CPU tensor (must be C-contiguous), or None."""
    if arr is None:
        return None
    if hasattr(arr, "data_ptr"):            # torch tensor (possibly pinned)
        if not arr.is_contiguous():
            raise SppError("tensor must be contiguous")
        return C.cast(arr.data_ptr(), C.POINTER(ctype))
    if not arr.flags["C_CONTIGUOUS"]:
        raise SppError("array must be C-contiguous")
    return arr.ctypes.data_as(C.POINTER(ctype))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


NET_NAMES = {
    "actor": _lib.NET_ACTOR, "critic_1": _lib.NET_CRITIC_1, "critic_2": _lib.NET_CRITIC_2, "critic": _lib.NET_CRITIC_1,
    "acm": _lib.NET_ACM, "critic_1_targ": _lib.NET_CRITIC_1_TARG, "critic_2_targ": _lib.NET_CRITIC_2_TARG,
    "critic_targ": _lib.NET_CRITIC_1_TARG, "actor_targ": _lib.NET_ACTOR_TARG,
}
LOSS_KEYS = ("critic_1", "critic_2", "actor", "pi", "dist", "alpha_loss", "alpha")


class Population:
    def __init__(self, algo="sac", ob_dim=11, ac_dim=3, population=1, device=0, acm_kind="acm", acm_critic=True,
                 norm_closs=False, min_max_denormalize=True, update_batch_size=256, acm_batch_size=128,
                 buffer_size=0, store_actions=True, gamma=0.99, tau=0.005, actor_lr=1e-3, critic_lr=1e-3,
                 alpha_lr=1e-3, acm_lr=3e-3, custom_loss=0.0, alpha=0.2, target_entropy=None):
        self.lib = _lib.load_library()
        self.algo = algo
        self.ob_dim, self.ac_dim, self.P = int(ob_dim), int(ac_dim), int(population)
        self.B = int(update_batch_size)
        self.acm_critic = bool(acm_critic)
        cfg = Config()
        cfg.algo = _lib.ALGO_SAC if algo == "sac" else _lib.ALGO_DDPG
        cfg.ob_dim, cfg.ac_dim = self.ob_dim, self.ac_dim
        cfg.acm_kind = _lib.ACM_MLP if acm_kind in ("acm", "mlp") else _lib.ACM_BASIC
        cfg.acm_critic, cfg.norm_closs, cfg.min_max_denormalize = int(acm_critic), int(norm_closs), int(min_max_denormalize)
        cfg.update_batch_size, cfg.acm_batch_size = self.B, int(acm_batch_size)
        cfg.store_actions, cfg.buffer_size = int(store_actions), int(buffer_size)
        cfg.gamma, cfg.tau, cfg.actor_lr, cfg.critic_lr = gamma, tau, actor_lr, critic_lr
        cfg.alpha_lr, cfg.acm_lr, cfg.custom_loss, cfg.alpha = alpha_lr, acm_lr, float(custom_loss), alpha
        cfg.target_entropy = float(-ac_dim if target_entropy is None else target_entropy)
        self.cfg = cfg
        h = C.c_void_p()
        check(self.lib.spp_population_create(C.byref(cfg), self.P, int(device), C.byref(h)))
        self.h = h
        self._tensors = {}

    # ------------------------------------------------------------------ lifecycle
    def close(self):
        if getattr(self, "h", None):
            self.lib.spp_population_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        check(self.lib.spp_sync(self.h))

    # ------------------------------------------------------------------ limits / stats
    def set_limits(self, actor_lim, acm_lim):
        a = _f32(np.broadcast_to(np.asarray(actor_lim, dtype=np.float32), (self.ob_dim,)))
        m = _f32(np.broadcast_to(np.asarray(acm_lim, dtype=np.float32), (self.ac_dim,)))
        check(self.lib.spp_set_limits(self.h, _ptr(a, C.c_float), _ptr(m, C.c_float)))

    def set_norm_stats(self, min_obs=None, max_obs=None, obs_mean=None, obs_std=None, agent=-1):
        arrs = [None if v is None else _f32(v) for v in (min_obs, max_obs, obs_mean, obs_std)]
        check(self.lib.spp_set_norm_stats(self.h, agent, *[_ptr(v, C.c_float) for v in arrs]))

    # ------------------------------------------------------------------ parameters
    def tensor_list(self, net):
        nid = NET_NAMES[net] if isinstance(net, str) else net
        if nid not in self._tensors:
            out = []
            for t in range(self.lib.spp_net_tensor_count(self.h, nid)):
                name = C.create_string_buffer(64)
                r, c = C.c_int(), C.c_int()
                check(self.lib.spp_net_tensor_info(self.h, nid, t, name, 64, C.byref(r), C.byref(c)))
                out.append((name.value.decode(), r.value, c.value))
            self._tensors[nid] = out
        return nid, self._tensors[nid]

    @staticmethod
    def _ref_shape(name, rows, cols):
        """Shape of the reference's tensor: biases [rows], gains t [1] / t1 [ac], weights [rows, cols]."""
        return (rows,) if (name.endswith(".bias") or name in ("t", "t1")) else (rows, cols)

    def load_state_dict(self, net, sd, agent=-1):
        """`sd`: mapping name -> array in the reference's state_dict layout (rltoolkit nn.Module keys)."""
        nid, tl = self.tensor_list(net)
        for t, (name, rows, cols) in enumerate(tl):
            v = sd[name]
            v = v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)
            v = _f32(v).reshape(-1)
            if v.size != int(np.prod(self._ref_shape(name, rows, cols))):
                raise SppError("tensor %s: expected %s elements" % (name, self._ref_shape(name, rows, cols)))
            check(self.lib.spp_params_upload(self.h, agent, nid, t, _ptr(v, C.c_float)))

    def state_dict(self, net, agent=0):
        nid, tl = self.tensor_list(net)
        out = OrderedDict()
        for t, (name, rows, cols) in enumerate(tl):
            shape = self._ref_shape(name, rows, cols)
            v = np.empty(shape, np.float32)
            check(self.lib.spp_params_download(self.h, agent, nid, t, _ptr(v, C.c_float)))
            out[name] = v
        return out

    def adam_state(self, net, agent=0):
        """-> (OrderedDict name -> (exp_avg, exp_avg_sq), step)"""
        nid, tl = self.tensor_list(net)
        out = OrderedDict()
        step = C.c_int(0)
        for t, (name, rows, cols) in enumerate(tl):
            shape = self._ref_shape(name, rows, cols)
            m, v = np.empty(shape, np.float32), np.empty(shape, np.float32)
            check(self.lib.spp_adam_download(self.h, agent, nid, t, _ptr(m, C.c_float), _ptr(v, C.c_float), C.byref(step)))
            out[name] = (m, v)
        return out, step.value

    def adam_reset(self, net, agent=-1):
        """fresh optimiser state for one net (what assigning a new module to model.<net> does in the reference)"""
        nid = NET_NAMES[net] if isinstance(net, str) else net
        check(self.lib.spp_adam_reset(self.h, int(agent), nid))

    def set_learning_rates(self, actor_lr=-1.0, critic_lr=-1.0, alpha_lr=-1.0, acm_lr=-1.0):
        check(self.lib.spp_set_learning_rates(self.h, float(actor_lr), float(critic_lr), float(alpha_lr), float(acm_lr)))

    def set_obs_norm(self, on):
        """obs_norm=True: the fused ring updates gather normalised obs / next_obs (replay_buffer.py:246-248)."""
        check(self.lib.spp_set_obs_norm(self.h, int(bool(on))))

    def sync_targets(self, agent=-1):
        check(self.lib.spp_sync_targets(self.h, agent))

    def alpha(self, agent=0):
        la, al = C.c_double(), C.c_double()
        check(self.lib.spp_alpha_get(self.h, agent, C.byref(la), C.byref(al)))
        return la.value, al.value

    def set_log_alpha(self, log_alpha, agent=-1):
        check(self.lib.spp_alpha_set(self.h, agent, float(log_alpha)))

    # ------------------------------------------------------------------ replay ring
    def ring_add_obs(self, agent, obs):
        o = _f32(obs).reshape(-1)
        idx = C.c_int64()
        check(self.lib.spp_ring_add_obs(self.h, agent, _ptr(o, C.c_float), C.byref(idx)))
        return idx.value

    def ring_add_acm_action(self, agent, acm_action):
        a = _f32(acm_action).reshape(-1)
        check(self.lib.spp_ring_add_acm_action(self.h, agent, _ptr(a, C.c_float)))

    def ring_add_timestep(self, agent, obs_idx, next_obs_idx, action, rew, done, end):
        a = None if action is None else _f32(action).reshape(-1)
        check(self.lib.spp_ring_add_timestep(self.h, agent, int(obs_idx), int(next_obs_idx), _ptr(a, C.c_float),
                                             float(rew), int(bool(done)), int(bool(end))))

    def ring_add_rollout_store(self, agent, policy):
        """ReplayBufferAcM.add_buffer (replay_buffer.py:284-297) from the [T][E] store a device rollout left in `policy` (PpoPolicy)."""
        check(self.lib.spp_ring_add_rollout_store(self.h, int(agent), policy.h))

    def ring_reset(self, agent=-1):
        check(self.lib.spp_ring_reset(self.h, agent))

    def ring_state(self, agent=0):
        out = (C.c_int64 * 3)()
        check(self.lib.spp_ring_state(self.h, agent, out))
        return int(out[0]), int(out[1]), int(out[2])

    def ring_sample_batch(self, agent, idx):
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        n = idx.size
        obs, nobs = np.empty((n, self.ob_dim), np.float32), np.empty((n, self.ob_dim), np.float32)
        act = np.empty((n, self.ob_dim), np.float32)
        rew, done = np.empty(n, np.float32), np.empty(n, np.int8)
        aacm = np.empty((n, self.ac_dim), np.float32)
        check(self.lib.spp_ring_sample_batch(self.h, agent, _ptr(idx, C.c_int64), n, _ptr(obs, C.c_float), _ptr(nobs, C.c_float),
                                             _ptr(act, C.c_float), _ptr(rew, C.c_float), _ptr(done, C.c_int8), _ptr(aacm, C.c_float)))
        return obs, nobs, act, rew, done, aacm

    def ring_fill_synthetic(self, seed, n, episode_len=1000):
        check(self.lib.spp_ring_fill_synthetic(self.h, int(seed), int(n), int(episode_len)))

    def ring_gather_bench(self, n_batches, seed=0, stream=None):
        b = C.c_double()
        check(self.lib.spp_ring_gather_bench_device(self.h, int(n_batches), int(seed), C.byref(b), stream))
        return b.value

    def ring_gather_bench_rows(self, n_batches, first, n):
        """Rows [first, first + n) of the dense minibatches the last ring_gather_bench call produced (test hook)."""
        obs = np.empty((n, self.ob_dim), np.float32); nobs = np.empty_like(obs); aacm = np.empty((n, self.ac_dim), np.float32)
        rew = np.empty(n, np.float32); done = np.empty(n, np.int8)
        check(self.lib.spp_ring_gather_bench_rows(self.h, int(n_batches), int(first), int(n), _ptr(obs, C.c_float), _ptr(nobs, C.c_float),
                                                  _ptr(aacm, C.c_float), _ptr(rew, C.c_float), _ptr(done, C.c_int8)))
        return obs, nobs, aacm, rew, done

    # ------------------------------------------------------------------ updates
    def update_host(self, grad_steps, obs, next_obs, action, reward, done, acm_action, eps=None, seed=0, losses=None):
        """The reference's update(obs, next_obs, action, reward, done, acm_action), G steps for P agents.
        Arrays are [P, G, B, ...] host buffers (numpy or pinned torch).  Returns losses [P, G, 8]."""
        if losses is None:
            losses = np.empty((self.P, grad_steps, _lib.LOSS_COUNT), np.float32)
        check(self.lib.spp_update_host(self.h, int(grad_steps), _ptr(obs, C.c_float), _ptr(next_obs, C.c_float),
                                       _ptr(action, C.c_float), _ptr(reward, C.c_float), _ptr(done, C.c_int8),
                                       _ptr(acm_action, C.c_float), _ptr(eps, C.c_float), int(seed), _ptr(losses, C.c_float)))
        return losses

    def update_ring(self, grad_steps, idx=None, eps=None, seed=0, losses=None):
        """sample_batch + update from the device ring; idx int64 [P, G, B] host indices or None."""
        if losses is None:
            losses = np.empty((self.P, grad_steps, _lib.LOSS_COUNT), np.float32)
        check(self.lib.spp_update_ring(self.h, int(grad_steps), _ptr(idx, C.c_int64), _ptr(eps, C.c_float), int(seed),
                                       _ptr(losses, C.c_float)))
        return losses

    def update_ring_device(self, grad_steps, seed=0, losses_dev_ptr=None, stream=None):
        check(self.lib.spp_update_ring_device(self.h, int(grad_steps), int(seed), losses_dev_ptr, stream))

    # ------------------------------------------------------------------ ACM regression
    def acm_update_host(self, n_batches, x, y, losses=None, last_rows=0):
        """n x AcMTrainer.batch_update(x, y): x [P, n, acm_batch_size, 2*ob], y [P, n, acm_batch_size, ac] -> losses [P, n]"""
        if losses is None:
            losses = np.empty((self.P, n_batches), np.float32)
        check(self.lib.spp_acm_update_host(self.h, int(n_batches), _ptr(x, C.c_float), _ptr(y, C.c_float), int(last_rows), _ptr(losses, C.c_float)))
        return losses

    def acm_validation_loss(self, x, y):
        """AcMTrainer.calculate_validation_loss (acm.py:329-343) for one validation set per agent: x [P, N, 2*ob],
        y [P, N, ac] -> MSE [P].  The set is cut into acm_batch_size chunks on the device; chunk means are combined in fp64."""
        x = _f32(x); y = _f32(y)
        N, B = x.shape[1], int(self.cfg.acm_batch_size)
        if N < 1:
            raise ValueError("No validation data. Were the pretrain ran?")
        nb = (N + B - 1) // B
        xp = np.zeros((self.P, nb * B, x.shape[2]), np.float32); xp[:, :N] = x
        yp = np.zeros((self.P, nb * B, y.shape[2]), np.float32); yp[:, :N] = y
        last = N - (nb - 1) * B
        losses = np.empty((self.P, nb), np.float32)
        check(self.lib.spp_acm_eval_host(self.h, nb, _ptr(xp, C.c_float), _ptr(yp, C.c_float), 0 if last == B else last,
                                         _ptr(losses, C.c_float)))
        w = np.full(nb, B, np.float64); w[-1] = last
        return (losses.astype(np.float64) * w).sum(axis=1) / N

    def acm_update_ring(self, n_batches, idx=None, seed=0, losses=None, last_rows=0):
        """AcMTrainer.update_acm_batches(n) from the device ring; idx int64 [P, n, acm_batch_size] or None."""
        if losses is None:
            losses = np.empty((self.P, n_batches), np.float32)
        check(self.lib.spp_acm_update_ring(self.h, int(n_batches), _ptr(idx, C.c_int64), int(last_rows), int(seed), _ptr(losses, C.c_float)))
        return losses

    # ------------------------------------------------------------------ rollout
    def rollout_step(self, obs, noise, eps=None, random_phase=False, act_noise=0.1, obs_norm=False, denormalize_actor_out=True,
                     out=None):
        """noise_action / initial_act + process_action for E observations per agent: obs, noise, eps [P, E, ob]
        -> (state_target [P, E, ob], acm_action [P, E, ac]).  `out` = (target, action) float32 arrays to fill (e.g. views of
        pinned buffers, so that the device-to-host copies are true asynchronous DMA)."""
        obs = _f32(obs); noise = _f32(noise)
        E = obs.shape[1]
        eps = None if eps is None else _f32(eps)
        if out is None:
            tgt = np.empty((self.P, E, self.ob_dim), np.float32)
            act = np.empty((self.P, E, self.ac_dim), np.float32)
        else:
            tgt, act = out
            if tgt.shape != (self.P, E, self.ob_dim) or act.shape != (self.P, E, self.ac_dim) or tgt.dtype != np.float32 \
                    or act.dtype != np.float32 or not (tgt.flags.c_contiguous and act.flags.c_contiguous):
                raise ValueError("out = (target [P, E, ob], action [P, E, ac]) contiguous float32 arrays")
        check(self.lib.spp_rollout_step_host(self.h, int(E), _ptr(obs, C.c_float), _ptr(noise, C.c_float), _ptr(eps, C.c_float),
                                             int(random_phase), float(act_noise), int(bool(obs_norm)),
                                             int(bool(denormalize_actor_out)), _ptr(tgt, C.c_float), _ptr(act, C.c_float)))
        return tgt, act

    def ring_obs_stats(self, return_bytes=False):
        """MetaReplayBuffer.update_obs_mean_std (rltoolkit/buffer/replay_buffer.py:83-96) for every agent, on the device:
        -> dict of [P, ob] float64 arrays: mean, std (numpy ddof 0), p1, p99 (np.percentile 'linear').  The kernels return the
        exact order statistics; the interpolation below is numpy's own `_lerp` in fp64, so p1 / p99 are bit-exact."""
        out = np.empty((self.P, 6, self.ob_dim), np.float64)
        nbytes = C.c_double()
        check(self.lib.spp_ring_obs_stats(self.h, _ptr(out, C.c_double), C.byref(nbytes)))
        res = {"mean": out[:, 0].copy(), "std": out[:, 1].copy()}
        n = np.array([self.ring_state(a)[2] for a in range(self.P)], np.float64)
        for name, q, lo, hi in (("p1", 1, 2, 3), ("p99", 99, 4, 5)):
            qq = np.true_divide(q, 100)
            v = (n - 1) * qq                                           # numpy's 'linear' method: get_virtual_index
            t = (v - np.floor(v))[:, None]
            a, b = out[:, lo], out[:, hi]
            d = b - a
            lerp = a + d * t
            lerp = np.where(t >= 0.5, b - d * (1 - t), lerp)           # numpy.lib._function_base_impl._lerp
            res[name] = lerp
        return (res, nbytes.value) if return_bytes else res

    def rollout_synthetic(self, envs_per_agent, steps, seed=0, act_noise=0.1, stream=None, random_phase=False):
        check(self.lib.spp_rollout_synthetic_device(self.h, int(envs_per_agent), int(steps), int(seed), float(act_noise),
                                                    int(bool(random_phase)), stream))

    # ------------------------------------------------------------------ the whole train loop, population-batched and device-resident
    def train_synthetic(self, frames, envs_per_agent=1, update_freq=50, grad_steps=50, random_frames=100, act_noise=0.1,
                        steps_per_epoch=1000, acm_update_freq=0, acm_update_batches=0, update_stats=True, seed=0, stream=None,
                        state=None):
        """P agents' frame loop of DDPG.collect_batch_and_train + DDPG_AcM.make_update (rltoolkit/algorithms/ddpg/ddpg.py:191-237,
        rltoolkit/acm/off_policy/ddpg_acm.py:52-85) and the per-iteration statistics refresh (ddpg.py:159-169), with every
        stage on the device: rollout_synthetic (actor -> noise -> ACM -> synthetic env -> ring) -> update bursts -> ACM
        regression batches -> ring_obs_stats.  The launches are cut exactly where the reference's conditions fire (train_schedule):
        with envs_per_agent = 1 the frame counter, the update frames and the ACM-update frames are the reference's; with E
        environments per agent one vector step counts E frames.  `state`: the dict a previous call returned (frames, iteration,
        stats), to continue a run.  -> state with `launch_log` (the executed schedule)."""
        st = dict(frames=0, iteration=0, min_obs=None, max_obs=None, updates=0, acm_batches=0, stats_updates=0) if state is None else dict(state)
        E = int(envs_per_agent)
        log = []
        for ev in train_schedule(st["frames"], frames, E, self.B, update_freq, grad_steps, random_frames, steps_per_epoch,
                                 acm_update_freq, acm_update_batches, st["iteration"], lambda: self.ring_state(0)[2]):
            kind = ev[0]
            if kind == "rollout":
                _, steps, random_phase, f0 = ev
                self.rollout_synthetic(E, steps, seed=seed * 1000003 + f0, act_noise=act_noise, stream=stream, random_phase=random_phase)
                st["frames"] = f0 + steps * E
            elif kind == "update":
                self.update_ring_device(ev[1], seed=seed * 7919 + st["frames"], stream=stream)
                st["updates"] += ev[1] * self.P
            elif kind == "acm":
                if stream is not None:
                    self.sync_stream(stream)
                self.acm_update_ring(ev[1], idx=None, seed=seed * 31 + st["frames"])
                st["acm_batches"] += ev[1] * self.P
            elif kind == "stats":
                st["iteration"] += 1
                if update_stats:
                    if stream is not None:
                        self.sync_stream(stream)
                    r = self.ring_obs_stats()
                    mn, mx = r["p1"].astype(np.float32), r["p99"].astype(np.float32)      # MetaReplayBuffer.update_obs_mean_std: running widening
                    st["min_obs"] = mn if st["min_obs"] is None else np.minimum(mn, st["min_obs"])
                    st["max_obs"] = mx if st["max_obs"] is None else np.maximum(mx, st["max_obs"])
                    for a in range(self.P):
                        self.set_norm_stats(st["min_obs"][a], st["max_obs"][a], r["mean"][a].astype(np.float32), r["std"][a].astype(np.float32), agent=a)
                    st["stats_updates"] += 1
            log.append(ev)
        st["launch_log"] = log
        return st

    def sync_stream(self, stream):
        import torch
        torch.cuda.ExternalStream(stream).synchronize() if isinstance(stream, int) else stream.synchronize()

    # ------------------------------------------------------------------ introspection
    def debug_scratch(self, agent, name):
        cap = 4096 * 512
        buf = np.empty(cap, np.float32)
        r, ld = C.c_int(), C.c_int()
        check(self.lib.spp_debug_scratch(self.h, agent, name.encode(), _ptr(buf, C.c_float), cap, C.byref(r), C.byref(ld)))
        return buf[: r.value * ld.value].reshape(r.value, ld.value).copy()


def kernel_launches() -> int:
    return int(_lib.load_library().spp_kernel_launches())


def train_schedule(frame0, frames, E, batch, update_freq, grad_steps, random_frames, steps_per_epoch, acm_update_freq, acm_update_batches,
                   iteration0=0, ring_len=None):
    """The launches of Population.train_synthetic in order, as tuples: ("rollout", vector_steps, random_phase, first_frame),
    ("update", grad_steps), ("acm", n_batches), ("stats",).  Pure host logic (CPU-tested against a frame-by-frame restatement of
    the reference's conditions): after every frame f (frames counted AFTER the increment, ddpg.py:221-223)
        update      if len(ring) > batch and f % update_freq == 0                      (DDPG.update_condition, ddpg.py:225-229)
        acm update  if iteration > 0 and acm batches > 0 and f % acm_update_freq == 0  (DDPG_AcM.acm_update_condition, ddpg_acm.py:52-57)
        stats       when f reaches a multiple of steps_per_epoch                        (perform_iteration, ddpg.py:159-169)
    and the random phase ends at random_frames (ddpg.py:205-208).  Consecutive frames with no event in between are one rollout
    launch.  ring_len: callable -> current ring length (None: assume the ring holds every frame since frame 0, uncapped)."""
    E = int(E)
    for name, v in (("update_freq", update_freq), ("steps_per_epoch", steps_per_epoch), ("random_frames", random_frames)):
        if v % E:
            raise ValueError("%s (%d) must be a multiple of envs_per_agent (%d)" % (name, v, E))
    if acm_update_batches and acm_update_freq % E:
        raise ValueError("acm_update_freq must be a multiple of envs_per_agent")
    f, it = int(frame0), int(iteration0)
    end = f + int(frames)
    while f < end:
        # next frame count at which anything can happen
        nxt = end
        for period in (update_freq, steps_per_epoch) + ((acm_update_freq,) if acm_update_batches and acm_update_freq else ()):
            nxt = min(nxt, (f // period + 1) * period)
        if f < random_frames:
            nxt = min(nxt, random_frames)
        steps = (nxt - f) // E
        yield ("rollout", steps, f < random_frames, f)
        f = nxt
        n = ring_len() if ring_len is not None else f
        if n > batch and f % update_freq == 0:
            yield ("update", grad_steps)
        if it > 0 and acm_update_batches and acm_update_freq and f % acm_update_freq == 0:
            yield ("acm", acm_update_batches)
        if f % steps_per_epoch == 0:
            yield ("stats",)
            it += 1
