"""Host handle on the device-resident SPP-PPO policy (thin wrapper over the spp_ppo_* C ABI).

Mirrors the update part of rltoolkit's PPO_AcM.perform_iteration (rltoolkit/acm/on_policy.py:55-86):
update_critic -> calculate_advantage (q-values + GAE) -> advantage normalisation -> clipped-ratio actor epochs
with KL early stop.  In data-parallel runs (one process per GPU) the per-step gradient vector is all-reduced with
torch.distributed (NCCL) between the device-side grad and apply kernels.
"""
import ctypes as C
from collections import OrderedDict

import numpy as np

from . import _lib
from ._lib import PpoConfig, SppError, check
from .population import _f32, _ptr

NETS = {"actor": 0, "critic": 1}


class PpoPolicy:
    def __init__(self, ob_dim, ac_dim, max_rows, max_batch_rows, device=0, min_max_denormalize=True, norm_closs=False,
                 gamma=0.99, gae_lambda=0.95, ppo_epsilon=0.2, entropy_coef=0.0, custom_loss=0.0, actor_lr=3e-4, critic_lr=3e-4):
        self.lib = _lib.load_library()
        cfg = PpoConfig()
        cfg.ob_dim, cfg.ac_dim = int(ob_dim), int(ac_dim)
        cfg.min_max_denormalize, cfg.norm_closs = int(min_max_denormalize), int(norm_closs)
        cfg.max_rows, cfg.max_batch_rows = int(max_rows), int(max_batch_rows)
        cfg.gamma, cfg.gae_lambda, cfg.ppo_epsilon, cfg.entropy_coef = gamma, gae_lambda, ppo_epsilon, entropy_coef
        cfg.custom_loss, cfg.actor_lr, cfg.critic_lr = float(custom_loss), actor_lr, critic_lr
        self.cfg = cfg
        self.ob_dim, self.ac_dim = int(ob_dim), int(ac_dim)
        h = C.c_void_p()
        check(self.lib.spp_ppo_create(C.byref(cfg), int(device), C.byref(h)))
        self.h = h
        self.N = 0
        self._tensors = {}

    def close(self):
        if getattr(self, "h", None):
            self.lib.spp_ppo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        check(self.lib.spp_ppo_sync(self.h))

    # ------------------------------------------------------------------ parameters
    def set_limits(self, actor_lim):
        a = _f32(np.broadcast_to(np.asarray(actor_lim, np.float32), (self.ob_dim,)))
        check(self.lib.spp_ppo_set_limits(self.h, _ptr(a, C.c_float)))

    def set_norm_stats(self, min_obs=None, max_obs=None, obs_mean=None, obs_std=None):
        arrs = [None if v is None else _f32(v) for v in (min_obs, max_obs, obs_mean, obs_std)]
        self._stats = arrs
        check(self.lib.spp_ppo_set_norm_stats(self.h, *[_ptr(v, C.c_float) for v in arrs]))

    def tensor_list(self, net):
        nid = NETS[net]
        if nid not in self._tensors:
            out = []
            for t in range(self.lib.spp_ppo_tensor_count(self.h, nid)):
                name = C.create_string_buffer(64)
                r, c = C.c_int(), C.c_int()
                check(self.lib.spp_ppo_tensor_info(self.h, nid, t, name, 64, C.byref(r), C.byref(c)))
                out.append((name.value.decode(), r.value, c.value))
            self._tensors[nid] = out
        return nid, self._tensors[nid]

    @staticmethod
    def _shape(name, rows, cols):
        return (rows,) if (name.endswith(".bias") or name == "log_scale") else (rows, cols)

    def load_state_dict(self, net, sd):
        nid, tl = self.tensor_list(net)
        for t, (name, rows, cols) in enumerate(tl):
            v = sd[name]
            v = _f32(v.detach().cpu().numpy() if hasattr(v, "detach") else v).reshape(-1)
            if v.size != int(np.prod(self._shape(name, rows, cols))):
                raise SppError("tensor %s has the wrong size" % name)
            check(self.lib.spp_ppo_params_upload(self.h, nid, t, _ptr(v, C.c_float)))

    def state_dict(self, net):
        nid, tl = self.tensor_list(net)
        out = OrderedDict()
        for t, (name, rows, cols) in enumerate(tl):
            v = np.empty(self._shape(name, rows, cols), np.float32)
            check(self.lib.spp_ppo_params_download(self.h, nid, t, _ptr(v, C.c_float)))
            out[name] = v
        return out

    # ------------------------------------------------------------------ data
    def load_rollout(self, obs, next_obs, actions, logp, rew, done, end, traj_start, traj_len, traj_stride=1, global_rows=0):
        obs, next_obs, actions = _f32(obs), _f32(next_obs), _f32(actions)
        logp, rew, done, end = _f32(logp), _f32(rew), _f32(done), _f32(end)
        ts = np.ascontiguousarray(traj_start, np.int64); tl = np.ascontiguousarray(traj_len, np.int64)
        self.N = obs.shape[0]
        self.Ntot = int(global_rows) if global_rows else self.N
        check(self.lib.spp_ppo_load_rollout(self.h, self.N, _ptr(obs, C.c_float), _ptr(next_obs, C.c_float), _ptr(actions, C.c_float),
                                            _ptr(logp, C.c_float), _ptr(rew, C.c_float), _ptr(done, C.c_float), _ptr(end, C.c_float),
                                            _ptr(ts, C.c_int64), _ptr(tl, C.c_int64), int(ts.size), int(traj_stride), int(global_rows)))

    def rollout_synthetic(self, pop, envs, steps, agent=0, max_ep_len=1000, done_prob=0.001, seed=0, denormalize_actor_out=True,
                          reset_envs=False, noise_act=None, noise_env=None, u_done=None, noise_reset=None):
        """A2C.collect_batch on the device (spp_ppo_rollout_synthetic): `envs` vectorised synthetic environments x `steps` steps through
        this policy and the ACM of `pop`'s agent, into the policy's [T][E] store.  Injected noise tensors are for the parity tests."""
        arrs = [None if v is None else _f32(v) for v in (noise_act, noise_env, u_done, noise_reset)]
        check(self.lib.spp_ppo_rollout_synthetic(self.h, pop.h, int(agent), int(envs), int(steps), int(max_ep_len), float(done_prob), int(seed),
                                                 int(bool(denormalize_actor_out)), int(bool(reset_envs)), *[_ptr(v, C.c_float) for v in arrs]))
        self.N = self.Ntot = int(envs) * int(steps)

    def set_global_rows(self, n_global):
        """Data-parallel runs: the loss means are over the rows of ALL ranks (spp_ppo_set_global_rows)."""
        check(self.lib.spp_ppo_set_global_rows(self.h, int(n_global)))
        self.Ntot = int(n_global)

    def store(self, name):
        """One column of the loaded rows ("x", "xn", "act", "raw_obs", "raw_next", "aacm", "logp", "rew", "done", "end", "adv", "v")."""
        w = {"x": self.ob_dim, "xn": self.ob_dim, "act": self.ob_dim, "raw_obs": self.ob_dim, "raw_next": self.ob_dim, "aacm": self.ac_dim}.get(name, 0)
        out = np.empty((self.N, w) if w else (self.N,), np.float32)
        check(self.lib.spp_ppo_store_download(self.h, name.encode(), _ptr(out, C.c_float)))
        return out

    def act(self, obs, noise, denormalize_actor_out=True):
        """Actor.act + denormalise for E observations: -> (action [E, ob], logp [E], acm_target [E, ob])."""
        obs, noise = _f32(obs), _f32(noise)
        E = obs.shape[0]
        action = np.empty((E, self.ob_dim), np.float32); target = np.empty((E, self.ob_dim), np.float32)
        logp = np.empty(E, np.float32)
        check(self.lib.spp_ppo_act(self.h, E, _ptr(obs, C.c_float), _ptr(noise, C.c_float), int(bool(denormalize_actor_out)),
                                   _ptr(action, C.c_float), _ptr(logp, C.c_float), _ptr(target, C.c_float)))
        return action, logp, target

    def act_normalized(self, norm_obs, noise, denormalize_actor_out=True):
        """act() for observations that are normalised already (model.actor.act(buffer.normalize(obs)) in the notebook)."""
        saved = getattr(self, "_stats", None)
        try:
            check(self.lib.spp_ppo_set_norm_stats(self.h, None, None, None, None))      # identity normalisation for this call
            return self.act(norm_obs, noise, denormalize_actor_out)
        finally:
            if saved is not None:
                self.set_norm_stats(*saved)

    def adam_reset(self, net):
        check(self.lib.spp_ppo_adam_reset(self.h, NETS[net]))

    # ------------------------------------------------------------------ single-GPU forms
    def update_critic(self, n_target_updates=10, n_updates_per_target=10):
        loss = C.c_float()
        check(self.lib.spp_ppo_update_critic(self.h, int(n_target_updates), int(n_updates_per_target), C.byref(loss)))
        return loss.value

    def advantages(self, want_host=True):
        adv = np.empty(self.N, np.float32) if want_host else None
        check(self.lib.spp_ppo_advantages(self.h, _ptr(adv, C.c_float)))
        return adv

    def load_advantages(self, adv):
        a = _f32(adv).reshape(-1)
        if a.size != self.N:
            raise SppError("advantages must have one entry per loaded row")
        check(self.lib.spp_ppo_load_advantages(self.h, _ptr(a, C.c_float)))

    def normalize_adv(self, global_stats=None):
        g = None if global_stats is None else np.ascontiguousarray(global_stats, np.float64)
        check(self.lib.spp_ppo_normalize_adv(self.h, _ptr(g, C.c_double)))

    def adv_stats(self):
        out = (C.c_double * 3)()
        check(self.lib.spp_ppo_adv_stats(self.h, out))
        return np.array([out[0], out[1], out[2]])

    def set_actor_mode(self, plain_ppo=False, a2c=False):
        """plain_ppo=True: PPO.update_actor (custom_loss == 0, ppo.py:152-192) instead of PPO_AcM.update_actor_acm; a2c=True: the A2C
        policy-gradient step (A2C_AcM.update_actor_acm, on_policy.py:100-124; with plain_ppo also set: A2C.update_actor)."""
        check(self.lib.spp_ppo_set_actor_mode(self.h, (2 if a2c else 0) + int(bool(plain_ppo))))

    def set_critic_path(self, tensor_cores):
        """True (default): critic fit on tcgen05 (csrc/ppo_critic_tc.cu); False: the FFMA tile kernel."""
        check(self.lib.spp_ppo_set_critic_path(self.h, int(bool(tensor_cores))))

    def set_reserved_sms(self, n):
        """Leave n SMs out of the policy kernels' grids (a concurrent ACM burst on another stream holds one)."""
        check(self.lib.spp_ppo_set_reserved_sms(self.h, int(n)))

    def a2c_actor_step(self, accumulate, normalize_adv=True):
        """One A2C actor update on the loaded rollout and the advantages on the device: optional normalisation with A2C's epsilon
        (a2c.py:275-277), one full-batch gradient of mean(-logp * adv), gradient accumulation when the reference never zeroes
        (`accumulate`, on_policy.py:117-123), Adam.  Returns (actor loss, summed squared distance of the custom loss)."""
        if normalize_adv:
            check(self.lib.spp_ppo_normalize_adv_eps(self.h, None, 1e-8))
        perm = np.arange(self.N, dtype=np.int64)
        check(self.lib.spp_ppo_actor_minibatch_grad(self.h, _ptr(perm, C.c_int64), self.N, self.N))
        sc = self.scalars()
        check(self.lib.spp_ppo_grad_accumulate(self.h, int(bool(accumulate))))
        check(self.lib.spp_ppo_actor_apply(self.h))
        return float(sc[0]) / self.N, float(sc[2])

    def update_actor(self, perms, batch_size, kl_threshold, max_epochs=None):
        perms = np.ascontiguousarray(perms, np.int64)
        max_epochs = perms.shape[0] if max_epochs is None else int(max_epochs)
        losses = (C.c_float * 4)()
        epochs, kl = C.c_int(), C.c_float()
        check(self.lib.spp_ppo_update_actor(self.h, _ptr(perms, C.c_int64), max_epochs, int(batch_size), float(kl_threshold), losses,
                                            C.byref(epochs), C.byref(kl)))
        return {"actor": losses[0], "entropy": losses[1], "policy": losses[2], "dist": losses[3]}, epochs.value, kl.value

    # ------------------------------------------------------------------ data parallelism: NCCL inside the library
    def comm_init(self, dist):
        """One NCCL communicator for this policy over the ranks of the (already initialised) torch.distributed process group: rank 0
        draws the 128-byte id (spp_comm_unique_id), torch.distributed only carries it to the other ranks.  Afterwards update_critic,
        normalize_adv and actor_epoch_device all-reduce inside the library, on the policy's stream."""
        import torch

        rank, world = dist.get_rank(), dist.get_world_size()
        buf = C.create_string_buffer(128)
        if rank == 0:
            check(self.lib.spp_comm_unique_id(buf))
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone().to(dev)
        dist.broadcast(t, 0)
        check(self.lib.spp_ppo_comm_init(self.h, bytes(t.cpu().numpy().tobytes()), rank, world))
        self.world, self.rank = world, rank
        # the per-step gradient all-reduce over NVLink peer memory (csrc/ppo_p2p.cu): every rank exports its exchange buffer by CUDA
        # IPC, torch.distributed only carries the 64-byte handles.  SPP_PPO_P2P=0 keeps NCCL for that collective as well.
        import os
        import sys

        self.p2p = False
        if world > 1 and world <= 8 and dev == "cuda" and os.environ.get("SPP_PPO_P2P", "1") != "0":
            hb = C.create_string_buffer(64)
            check(self.lib.spp_ppo_p2p_handle(self.h, hb))
            mine = torch.frombuffer(bytearray(hb.raw), dtype=torch.uint8).clone().to(dev)
            allh = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allh, mine)
            blob = b"".join(bytes(x.cpu().numpy().tobytes()) for x in allh)
            rc = self.lib.spp_ppo_p2p_init(self.h, blob, rank, world)
            ok = torch.tensor([1 if rc == 0 else 0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)          # all ranks or none
            if int(ok.item()) == 1:
                self.p2p = True
            else:
                if rc == 0:
                    raise SppError("peer-memory all-reduce came up on this rank but not on every rank")
                print("spp_rl_b200: NVLink peer-memory all-reduce unavailable (%s); using NCCL" % self.lib.spp_last_error().decode(), file=sys.stderr)

    def comm_info(self):
        w, n, v = C.c_int(), C.c_int64(), C.c_int()
        check(self.lib.spp_ppo_comm_info(self.h, C.byref(w), C.byref(n), C.byref(v)))
        on, steps, err = C.c_int(), C.c_int64(), C.c_int()
        check(self.lib.spp_ppo_p2p_info(self.h, C.byref(on), C.byref(steps), C.byref(err)))
        if err.value:
            raise SppError("peer-memory all-reduce: a peer's vector never arrived (device-side time-out)")
        return {"world": w.value, "allreduces": n.value, "nccl_version": v.value, "p2p": bool(on.value), "p2p_steps": steps.value}

    def actor_epoch_device(self, ids_dev, off, n_global=None):
        """One epoch of minibatch steps (gather -> grad -> [all-reduce] -> Adam) with this rank's LOCAL row ids on the device
        (torch int64 CUDA tensor produced on the policy's stream); off: host offsets (len nb + 1).  -> log [nb, 8 + pad4(ob)]."""
        off = np.ascontiguousarray(off, np.int64)
        nb = off.size - 1
        ng = None if n_global is None else np.ascontiguousarray(n_global, np.int64)
        ldo = (self.ob_dim + 3) // 4 * 4
        log = np.empty((nb, 8 + ldo), np.float32)
        ptr = C.c_void_p(ids_dev.data_ptr()) if ids_dev is not None and ids_dev.numel() else None
        check(self.lib.spp_ppo_actor_epoch_device(self.h, ptr, _ptr(off, C.c_int64), _ptr(ng, C.c_int64), int(nb), _ptr(log, C.c_float)))
        return log

    def iteration_dp(self, perms, batch, E, kl_threshold=1e9, critic_targets=10, critic_steps=10, rank=0, world=1):
        """The update half of PPO_AcM.perform_iteration (on_policy.py:55-86) on this rank's environment shard: critic fit ->
        advantages -> advantage normalisation -> actor epochs with the KL early stop.  perms: [epochs, N_global] GLOBAL permutations
        (every rank holds the same); each rank filters its rows on the device (sharding.epoch_local_minibatches_device) and all
        collectives run inside the library (comm_init).  world == 1 runs the same code path without a communicator.
        -> dict(critic_loss, epochs, kl, phases_ms (host clock after a stream sync per phase), allreduces)."""
        import time

        import torch

        from .sharding import epoch_local_minibatches_device

        n0 = self.comm_info()["allreduces"]
        t0 = time.perf_counter()
        closs = self.update_critic(critic_targets, critic_steps)
        t1 = time.perf_counter()
        self.advantages(want_host=False)
        t1a = time.perf_counter()
        self.normalize_adv()
        t1b = time.perf_counter()
        self.sync()
        t2 = time.perf_counter()
        self._adv_split_ms = ((t1a - t1) * 1e3, (t1b - t1a) * 1e3, (t2 - t1b) * 1e3)
        on_device = isinstance(perms, (list, tuple)) and len(perms) > 0 and torch.is_tensor(perms[0])      # epoch permutations already on the device
        if not on_device:
            perms = np.ascontiguousarray(perms, np.int64)
        N = int(perms[0].numel()) if on_device else perms.shape[1]
        nb = (N + batch - 1) // batch
        ng = np.minimum(batch, N - batch * np.arange(nb)).astype(np.int64)
        kl, ran = 0.0, 0
        st = self._ext_stream()
        for ep in range(len(perms)):
            if kl >= kl_threshold:
                break
            with torch.cuda.stream(st):      # the epoch's permutation goes to the device once; the rank filters its rows there
                perm_dev = perms[ep] if on_device else torch.from_numpy(perms[ep]).to("cuda", non_blocking=False)
                if world > 1:
                    ids, off = epoch_local_minibatches_device(perm_dev, batch, E, rank, world)
                else:
                    ids, off = perm_dev, [min(k * batch, N) for k in range(nb + 1)]
                log = self.actor_epoch_device(ids, off, ng)
            kl = float(log[-1, 3]) / float(ng[-1])      # PS_KL of the LAST (possibly short) minibatch of the epoch (quirk 16)
            ran += 1
        t3 = time.perf_counter()
        return {"critic_loss": closs, "epochs": ran, "kl": kl, "allreduces": self.comm_info()["allreduces"] - n0,
                "phases_ms": {"critic_fit": (t1 - t0) * 1e3, "advantages_and_normalisation": (t2 - t1) * 1e3, "actor_epochs": (t3 - t2) * 1e3}}

    # ------------------------------------------------------------------ step-wise data-parallel forms (collective issued by torch.distributed)
    def grad_tensor(self):
        """torch views (no copy) of the reduced gradient vector and the 8 scalar slots, for dist.all_reduce."""
        g, sc, _ = self._grad_views()
        return g, sc

    def _grad_views(self):
        import torch

        if getattr(self, "_views", None) is None:
            ptr, sptr, n = C.c_void_p(), C.c_void_p(), C.c_int()
            check(self.lib.spp_ppo_grad_buffer(self.h, C.byref(ptr), C.byref(n), C.byref(sptr)))

            class _Arr:
                def __init__(self, p, k):
                    self.__cuda_array_interface__ = {"shape": (k,), "typestr": "<f4", "data": (p, False), "version": 2}
            g = torch.as_tensor(_Arr(ptr.value, n.value), device="cuda")
            sc = torch.as_tensor(_Arr(sptr.value, 8), device="cuda")
            # the library keeps the scalars right behind the gradient vector: one collective per optimiser step covers both
            both = torch.as_tensor(_Arr(ptr.value, n.value + 8), device="cuda") if sptr.value == ptr.value + 4 * n.value else None
            self._views = (g, sc, both)
        return self._views

    def _allreduce_grads(self, dist):
        g, sc, both = self._grad_views()
        if both is not None:
            dist.all_reduce(both)
        else:
            dist.all_reduce(g)
            dist.all_reduce(sc)

    def scalars(self):
        out = (C.c_float * 8)()
        check(self.lib.spp_ppo_scalars(self.h, out))
        return np.array(list(out), np.float32)

    def _ext_stream(self):
        """The library's stream as a torch stream: collectives issued under it are ordered between the kernels by the stream
        itself, so the data-parallel loops need no host synchronisation per optimiser step."""
        if getattr(self, "_torch_stream", None) is None:
            import torch

            ptr = C.c_void_p()
            check(self.lib.spp_ppo_stream(self.h, C.byref(ptr)))
            self._torch_stream = torch.cuda.ExternalStream(ptr.value) if ptr.value else torch.cuda.default_stream()
        return self._torch_stream

    def update_critic_dp(self, dist, n_target_updates=10, n_updates_per_target=10):
        """update_critic with one NCCL all-reduce of the gradient vector per optimiser step; everything is
        enqueued on the library's stream (grad kernel -> all-reduce -> Adam kernel), the loss sum is read once at the end."""
        import torch

        g, sc = self.grad_tensor()
        st = self._ext_stream()
        with torch.cuda.stream(st):
            tot = torch.zeros((), dtype=torch.float64, device=g.device)
            for _ in range(n_target_updates):
                check(self.lib.spp_ppo_critic_targets(self.h))
                for _ in range(n_updates_per_target):
                    check(self.lib.spp_ppo_critic_grad(self.h))
                    self._allreduce_grads(dist)
                    tot += sc[0].double()
                    check(self.lib.spp_ppo_critic_apply(self.h))
        st.synchronize()
        # the reference's loss["critic"]: mean over the optimiser steps of 0.5 * SSE / N (a2c.py:208-216), N = the GLOBAL row count
        return 0.5 * float(tot.item()) / (float(self.Ntot) * n_target_updates * n_updates_per_target)

    def actor_minibatch_dp(self, dist, perm_local, n_global, want_host=True):
        """One data-parallel actor minibatch; returns the 8 reduced scalars (host array, or a device tensor without any host
        synchronisation when want_host is False -- the caller then reads them once per epoch for the KL test)."""
        import torch

        g, sc = self.grad_tensor()
        st = self._ext_stream()
        with torch.cuda.stream(st):
            if torch.is_tensor(perm_local):      # local row ids already on the device (filtered there, under this stream)
                p = perm_local.contiguous()
                ptr = C.c_void_p(p.data_ptr()) if p.numel() else None      # a rank may own no row of this minibatch
                check(self.lib.spp_ppo_actor_minibatch_grad_device(self.h, ptr, int(p.numel()), int(n_global)))
            else:
                p = np.ascontiguousarray(perm_local, np.int64)
                check(self.lib.spp_ppo_actor_minibatch_grad(self.h, _ptr(p, C.c_int64), int(p.size), int(n_global)))
            self._allreduce_grads(dist)
            out = sc.clone()
            check(self.lib.spp_ppo_actor_apply(self.h))
        if want_host:
            st.synchronize()
            return out.cpu().numpy()
        return out
