#!/usr/bin/env python
"""Generate golden fixtures by running the UNMODIFIED reference (rltoolkit from /root/reference).

Run in the build container only (the reference tree does not exist on the GPU box):
    python tests/golden/make_golden.py
Reference arithmetic: torch 2.11.0 CPU (the reference pins torch 1.3.1, rltoolkit/requirements.txt:3),
numpy 2.3, single thread.  Inputs are produced by deterministic numpy generators shared with the tests
(tests/parity_util.py, spp_rl_b200/init.py), so a fixture stores only what the reference computed:
losses, temperature, post-update weights, target nets and Adam moments.

Fixtures written next to this file:
  sac_hopper_g3.npz      SAC_AcM.update x3, Hopper shapes, B=256, published flags (train/spp_sac_hopper.py:37-41)
  ddpg_hcheetah_g2.npz   DDPG_AcM.update x2, HalfCheetah shapes, B=100, BasicAcM (train/spp_ddpg_hcheetah.py)
  ring_ops.npz           BufferAcMOffPolicy add_obs/add_timestep/add_acm_action stream with wrap + sample_batch
  acm_add_buffer.npz     ReplayBufferAcM.add_buffer joint behaviour
  ppo_walker.npz         PPO_AcM pieces on Walker2d shapes: critic fit, GAE, advantage normalisation, actor epochs
  acm_regress.npz        AcMTrainer.batch_update x3 (AcM and BasicAcM)
  acm_epochs.npz         AcMTrainer.update_acm x3 epochs (shuffle, partial last minibatch, StepLR) + validation loss
  rollout_steps.npz      the frame-loop body under recorded noise: DDPG_AcM.noise_action / AcMOffPolicy.initial_act /
                         process_action for SAC_AcM and DDPG_AcM (three limit / normalisation settings), deterministic test()
                         actions, and Actor.act + AcMOnPolicyTrainer.process_action for PPO_AcM
  ppo_plain.npz          plain PPO.update_actor (what PPO_AcM runs when custom_loss == 0, on_policy.py:88-98) on ppo_walker's rollout
  a2c_acm.npz            A2C_AcM (on_policy.py:100-124): two iterations of A2C.update_critic -> q - V advantages -> update_actor_acm on
                         ppo_walker's rollout; the second one shows the never-zeroed gradients (post-step weights, Adam moments, losses)
  config1_pendulum.npz   BASELINE config 1: SAC_AcM on Pendulum-v0, pre_train() + train() for 400 frames from fixed weights and seeds:
                         replay indices drawn, ring cursors / index arrays, observation chain, final weights, statistics, returns
                         (its `ref_seconds` entry is the wall-clock time of the reference's pre_train / train here -- the frames/s
                         SURVEY 8d asks config 1 to report -- and therefore the ONE array that differs from run to run; every
                         other array of every fixture regenerates bit-identically)
  pkl_actions.npz        notebooks/load_and_test.ipynb flow on each of the 9 trained models/*.pkl (copied to tests/golden/models/):
                         construct with the notebook's flags, load(), deterministic action + ACM action on fixed observations
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.ref_import import import_reference  # noqa: E402

rl = import_reference()
from rltoolkit.acm.models.basic_acm import BasicAcM  # noqa: E402
from rltoolkit.buffer import BufferAcMOffPolicy, MemoryAcM, ReplayBufferAcM  # noqa: E402

from spp_rl_b200.init import init_state  # noqa: E402
from tests.parity_util import make_batches, make_stats  # noqa: E402

torch.set_num_threads(1)


def load_nets(model, s0, nets):
    for net in nets:
        sd = {k[len(net) + 1:]: torch.from_numpy(v.copy()) for k, v in s0.items() if k.startswith(net + ".")}
        getattr(model, net).load_state_dict(sd)


def dump_state(model, nets, opts):
    out = {}
    for net in nets:
        for k, v in getattr(model, net).state_dict().items():
            out[net + "." + k] = v.detach().numpy().copy()
    for net, opt in opts.items():
        mod = getattr(model, net)
        for (name, p) in mod.named_parameters():
            st = opt.state[p]
            out[net + "." + name + "#m"] = st["exp_avg"].numpy().copy()
            out[net + "." + name + "#v"] = st["exp_avg_sq"].numpy().copy()
            out[net + "#step"] = np.array(int(st["step"]))
    return out


def sac_fixture():
    ob, ac, B, G, seed = 11, 3, 256, 3, 0
    torch.manual_seed(0); np.random.seed(0)
    m = rl.SAC_AcM(env_name="Hopper-v2", update_batch_size=B, custom_loss=0.2, acm_critic=True, norm_closs=False,
                   denormalize_actor_out=True, min_max_denormalize=True, acm_pre_train_samples=100,
                   acm_val_buffer_size=None, buffer_size=1000, tensorboard_dir=None, log_dir=None, gamma=0.99,
                   actor_lr=1e-3, critic_lr=1e-3, alpha_lr=1e-3, alpha=0.2)
    mn, mx, mean, std = make_stats(ob, seed, True)
    for k, v in (("min_obs", mn), ("max_obs", mx), ("obs_mean", mean), ("obs_std", std)):
        setattr(m.replay_buffer, k, torch.from_numpy(v))
    s0 = init_state("sac", ob, ac, seed * 100)
    nets = ["actor", "critic_1", "critic_2", "acm"]
    load_nets(m, s0, nets)
    m.critic_1_targ.load_state_dict(m.critic_1.state_dict())
    m.critic_2_targ.load_state_dict(m.critic_2.state_dict())
    obs, nobs, act, rew, done, aacm, eps = make_batches(ob, ac, 1, G, B, seed, mn, mx)
    losses = []
    import torch.distributions.utils as du
    orig = du._standard_normal
    for g in range(G):
        draws = [torch.from_numpy(eps[0, g, 0]), torch.from_numpy(eps[0, g, 1])]
        it = iter(draws)
        import torch.distributions.normal as dn
        dn._standard_normal = lambda shape, dtype, device: next(it)
        t = lambda x: torch.from_numpy(x[0, g])
        m.update(t(obs), t(nobs), t(act), t(rew), t(done), t(aacm))
        dn._standard_normal = orig
        losses.append([m.loss["critic_1"], m.loss["critic_2"], m.loss["actor"], m.loss["sac"], m.loss["dist"], m.alpha])
    out = dump_state(m, nets[:3] + ["critic_1_targ", "critic_2_targ"],
                     {"actor": m.actor_optimizer, "critic_1": m.critic_1_optimizer, "critic_2": m.critic_2_optimizer})
    out["losses"] = np.array(losses, np.float64)
    out["log_alpha"] = np.array(float(m.log_alpha.item()))
    out["meta"] = np.array([ob, ac, B, G, seed])
    out["torch_version"] = np.array(torch.__version__)
    np.savez_compressed(os.path.join(HERE, "sac_hopper_g3.npz"), **out)
    print("sac_hopper_g3: losses", losses[-1])


def ddpg_fixture():
    ob, ac, B, G, seed = 17, 6, 100, 2, 1
    torch.manual_seed(0); np.random.seed(0)
    m = rl.DDPG_AcM(env_name="HalfCheetah-v2", update_batch_size=B, custom_loss=1.0, acm_critic=True, norm_closs=False,
                    denormalize_actor_out=True, min_max_denormalize=True, acm_pre_train_samples=100,
                    acm_val_buffer_size=None, buffer_size=1000, tensorboard_dir=None, log_dir=None, gamma=0.95,
                    actor_lr=5e-4, critic_lr=5e-4)
    m.acm = BasicAcM(2 * ob, ac, False)
    mn, mx, mean, std = make_stats(ob, seed, True)
    for k, v in (("min_obs", mn), ("max_obs", mx), ("obs_mean", mean), ("obs_std", std)):
        setattr(m.replay_buffer, k, torch.from_numpy(v))
    s0 = init_state("ddpg", ob, ac, seed * 100, "basic", True)
    s0["acm.t"][:] = 0.7
    s0["acm.t1"][:] = np.linspace(0.5, 1.5, ac)
    nets = ["actor", "critic", "acm"]
    load_nets(m, s0, nets)
    m.actor_targ.load_state_dict(m.actor.state_dict())
    m.critic_targ.load_state_dict(m.critic.state_dict())
    obs, nobs, act, rew, done, aacm, _ = make_batches(ob, ac, 1, G, B, seed, mn, mx)
    losses = []
    for g in range(G):
        t = lambda x: torch.from_numpy(x[0, g])
        m.update(t(obs), t(nobs), t(act), t(rew), t(done), t(aacm))
        losses.append([m.loss["critic"], m.loss["actor"], m.loss["ddpg"], m.loss["dist"]])
    out = dump_state(m, ["actor", "critic", "actor_targ", "critic_targ"], {"actor": m.actor_optimizer, "critic": m.critic_optimizer})
    out["losses"] = np.array(losses, np.float64)
    out["meta"] = np.array([ob, ac, B, G, seed])
    out["torch_version"] = np.array(torch.__version__)
    np.savez_compressed(os.path.join(HERE, "ddpg_hcheetah_g2.npz"), **out)
    print("ddpg_hcheetah_g2: losses", losses[-1])


def ring_fixture():
    rng = np.random.RandomState(0)
    size, ob, ac = 57, 5, 2
    ref = BufferAcMOffPolicy(size, ob, ob, acm_act_shape=ac, dtype=torch.float32, device=torch.device("cpu"))
    ops = []      # (kind, payload...) flattened into arrays below
    kinds, fvals, ivals, states = [], [], [], []
    for ep in range(30):
        T = rng.randint(1, 9)
        o = rng.randn(ob).astype(np.float32)
        prev = ref.add_obs(torch.from_numpy(o).unsqueeze(0))
        kinds.append(0); fvals.append(np.concatenate([o, np.zeros(ob + ac + 1, np.float32)])); ivals.append([prev, 0, 0, 0])
        states.append([ref.obs_idx, ref.ts_idx, ref.current_len])
        for t in range(T):
            a = rng.randn(ob).astype(np.float32); aa = rng.randn(ac).astype(np.float32)
            ref.add_acm_action(aa)
            o = rng.randn(ob).astype(np.float32)
            nxt = ref.add_obs(torch.from_numpy(o).unsqueeze(0))
            rew = np.float32(rng.randn()); done = bool(rng.rand() < 0.2); end = (t == T - 1)
            ref.add_timestep(prev, nxt, torch.from_numpy(a).unsqueeze(0), rew, done, end)
            kinds.append(1); fvals.append(np.concatenate([o, a, aa, [rew]]).astype(np.float32)); ivals.append([prev, nxt, int(done), int(end)])
            states.append([ref.obs_idx, ref.ts_idx, ref.current_len])
            prev = nxt
    L = ref.current_len
    np.random.seed(3)
    idx = np.random.randint(0, L, 16)
    np.random.seed(3)
    b = ref.sample_batch(16)
    np.savez_compressed(os.path.join(HERE, "ring_ops.npz"), size=size, ob=ob, ac=ac, kinds=np.array(kinds), fvals=np.array(fvals),
                        ivals=np.array(ivals), states=np.array(states), obs_idx=ref._obs_idx[:L].copy(),
                        next_obs_idx=ref._next_obs_idx[:L].copy(), idx=idx, s_obs=b[0].numpy(), s_next=b[1].numpy(),
                        s_act=b[2].numpy(), s_rew=b[3].numpy(), s_done=b[4].numpy(), s_aacm=b[5].numpy())
    print("ring_ops: len", L, "ops", len(kinds))

    # add_buffer joint behaviour
    mem = MemoryAcM()
    chain, acts = [], []
    for r in range(4):
        T = rng.randint(2, 6)
        o = rng.randn(ob).astype(np.float32); p = mem.add_obs(torch.from_numpy(o).unsqueeze(0)); chain.append(o)
        for t in range(T):
            aa = rng.randn(ac).astype(np.float32); mem.add_acm_action(aa); acts.append(aa)
            o = rng.randn(ob).astype(np.float32); n = mem.add_obs(torch.from_numpy(o).unsqueeze(0)); chain.append(o)
            mem.add_timestep(p, n, None, None, 0.0, False, t == T - 1); p = n
        mem.end_rollout()
    ref2 = ReplayBufferAcM(40, ob, ac, False)
    ref2.add_buffer(mem)
    np.savez_compressed(os.path.join(HERE, "acm_add_buffer.npz"), chain=np.array(chain), acts=np.array(acts),
                        joints=np.array(mem._new_rollout_idx), obs=ref2.obs.astype(np.float32), next_obs=ref2.next_obs.astype(np.float32),
                        actions_acm=ref2.actions_acm.astype(np.float32), length=ref2.current_len)
    print("acm_add_buffer: len", ref2.current_len, "of", len(acts))


def acm_regress_fixture():
    out = {}
    for kind in ("acm", "basic"):
        ob, ac = 17, 6
        torch.manual_seed(0)
        m = rl.DDPG_AcM(env_name="HalfCheetah-v2", acm_pre_train_samples=100, acm_val_buffer_size=None, buffer_size=1000,
                        tensorboard_dir=None, log_dir=None, acm_lr=1e-3)
        if kind == "basic":
            m.acm = BasicAcM(2 * ob, ac, False)
        s0 = init_state("ddpg", ob, ac, 7, kind, True)
        if kind == "basic":
            s0["acm.t"][:] = 0.7; s0["acm.t1"][:] = np.linspace(0.5, 1.5, ac)
        load_nets(m, s0, ["acm"])
        m.acm = m.acm      # rebuild optimiser on the loaded parameters
        rng = np.random.RandomState(11)
        losses = []
        for _ in range(3):
            x = rng.randn(100, 2 * ob).astype(np.float32); y = np.tanh(rng.randn(100, ac)).astype(np.float32)
            losses.append(m.batch_update(torch.from_numpy(x), torch.from_numpy(y)))
        d = dump_state(m, ["acm"], {"acm": m.acm_optimizer})
        out.update({kind + ":" + k: v for k, v in d.items()})
        out[kind + ":losses"] = np.array(losses)
    np.savez_compressed(os.path.join(HERE, "acm_regress.npz"), **out)
    print("acm_regress: done")


def acm_epochs_fixture():
    """AcMTrainer.update_acm (acm.py:266-303) x3 epochs over a 250-row buffer, batch 64 (partial last minibatch of 58),
    StepLR(step 1, gamma 0.5) stepped once per epoch, and calculate_validation_loss (acm.py:329-343) on a 90-row set."""
    ob, ac, n, nval, B = 17, 6, 250, 90, 64
    torch.manual_seed(0)
    m = rl.DDPG_AcM(env_name="HalfCheetah-v2", acm_pre_train_samples=n, acm_val_buffer_size=nval, buffer_size=1000,
                    tensorboard_dir=None, log_dir=None, acm_lr=2e-3, acm_batch_size=B, acm_scheduler_step=1,
                    acm_scheduler_gamma=0.5, debug_mode=False)
    s0 = init_state("ddpg", ob, ac, 9, "acm", True)
    load_nets(m, s0, ["acm"])
    m.acm = m.acm
    rng = np.random.RandomState(23)

    def fill(buf, rows, off_policy):
        chain = rng.randn(rows + 1, ob).astype(np.float32)
        acts = np.tanh(rng.randn(rows, ac)).astype(np.float32)
        prev = buf.add_obs(torch.from_numpy(chain[0:1]))
        for t in range(rows):
            if off_policy:
                buf.add_acm_action(acts[t])
            nxt = buf.add_obs(torch.from_numpy(chain[t + 1:t + 2]))
            if off_policy:
                buf.add_timestep(prev, nxt, torch.from_numpy(chain[t + 1:t + 2]), 0.0, False, False)
            else:
                buf.add_timestep(prev, nxt, torch.from_numpy(acts[t:t + 1]))
            prev = nxt
        return chain, acts

    chain, acts = fill(m.replay_buffer, n, True)
    vchain, vacts = fill(m.acm_val_buffer, nval, False)
    perms = [torch.randperm(n, generator=torch.Generator().manual_seed(31 + e)) for e in range(3)]
    calls = [0]
    orig = torch.randperm

    def fake(k, *a, **kw):     # RandomSampler calls randperm twice per epoch (the second for an empty remainder)
        calls[0] += 1
        return perms[(calls[0] - 1) // 2]
    torch.randperm = fake
    losses, vals, lrs = [], [], []
    for e in range(3):
        m.update_acm(epochs=1)
        losses.append(m.loss["acm"]); vals.append(m.loss["acm_val"]); lrs.append(m.acm_optimizer.param_groups[0]["lr"])
    torch.randperm = orig
    out = dump_state(m, ["acm"], {"acm": m.acm_optimizer})
    out.update(chain=chain, acts=acts, vchain=vchain, vacts=vacts, perms=torch.stack(perms).numpy(), losses=np.array(losses),
               val_losses=np.array(vals), lrs_after=np.array(lrs))
    np.savez_compressed(os.path.join(HERE, "acm_epochs.npz"), **out)
    print("acm_epochs: losses", losses, "val", vals, "lr after each epoch", lrs)


def ppo_fixture():
    from oracle import ppo as P
    from oracle.norm import NormStats, normalize

    torch.manual_seed(0); np.random.seed(0)
    m = rl.PPO_AcM(env_name="Walker2d-v2", gamma=0.99, acm_pre_train_samples=300, acm_pre_train_epochs=1, iterations=1,
                   batch_size=700, acm_update_freq=1, acm_epochs=1, acm_lr=1e-4, actor_lr=3e-4, critic_lr=3e-4,
                   kl_div_threshold=0.1, max_ppo_epochs=10, ppo_batch_size=256, acm_batch_size=64, denormalize_actor_out=True,
                   min_max_denormalize=True, custom_loss=0.1, obs_norm=True, tensorboard_dir=None, log_dir=None,
                   acm_val_buffer_size=None, norm_closs=False)
    m.pre_train()
    buf = MemoryAcM(obs_mean=m.obs_mean, obs_std=m.obs_std, device=m.device, alpha=m.obs_norm_alpha, max_obs=m.max_obs,
                    min_obs=m.min_obs, min_max_denormalize=m.min_max_denormalize)
    m.buffer = buf
    m.collect_batch(buf)
    N = len(buf)
    out = {"chain": torch.cat(buf._obs).numpy(), "joints": np.array(buf._new_rollout_idx), "rewards": np.array(buf.rewards, np.float32),
           "done": np.array(buf.done, np.float32), "end": np.array(buf.end, np.float32),
           "actions": torch.cat(buf.actions).numpy(), "logp": torch.cat(buf.action_logprobs).detach().numpy(),
           "actions_acm": np.array(buf.actions_acm, np.float32),
           "min_obs": m.min_obs.numpy(), "max_obs": m.max_obs.numpy(), "obs_mean": m.obs_mean.numpy(), "obs_std": m.obs_std.numpy(),
           "actor_lim": np.array(float(m.actor_ac_lim)), "acm_lim": m.ac_lim.numpy()}
    for net in ("actor", "critic", "acm"):
        for k, v in getattr(m, net).state_dict().items():
            out["pre:" + net + "." + k] = v.detach().numpy().copy()
    adv = m.update_critic(buf)
    out["critic_loss"] = np.array(m.loss["critic"])
    out["adv"] = adv.numpy().copy()
    for k, v in m.critic.state_dict().items():
        out["fit:critic." + k] = v.detach().numpy().copy()
    perms = [torch.randperm(N, generator=torch.Generator().manual_seed(7 + e)) for e in range(m.max_ppo_epochs)]
    out["perms"] = torch.stack(perms).numpy()
    calls = [0]
    orig = torch.randperm

    def fake(n, *a, **k):     # RandomSampler calls randperm twice per epoch (the second for an empty remainder)
        calls[0] += 1
        return perms[(calls[0] - 1) // 2]
    torch.randperm = fake
    m.update_actor_acm(adv, buf)
    torch.randperm = orig
    out["epochs_run"] = np.array(m.kl_div_updates_counter)
    out["actor_losses"] = np.array([m.loss["actor"], m.loss["entropy"], m.loss["policy"], m.loss["dist"]])
    for k, v in m.actor.state_dict().items():
        out["post:actor." + k] = v.detach().numpy().copy()
    out["hp"] = np.array([m.gamma, m.gae_lambda, m.ppo_epsilon, m.kl_div_threshold, m.max_ppo_epochs, m.ppo_batch_size, m.actor_lr,
                          m.critic_lr, m.entropy_coef, m.custom_loss, m.critic_num_target_updates, m.num_critic_updates_per_target])
    np.savez_compressed(os.path.join(HERE, "ppo_walker.npz"), **out)
    print("ppo_walker: N", N, "epochs", m.kl_div_updates_counter, "losses", out["actor_losses"])


class _Recorder:
    """Records the reference's random draws without changing them: rsample's _standard_normal, torch.randn and torch.normal
    (Normal.sample) are wrapped; the torch.normal wrapper draws z itself and returns z * scale + loc, which is what ATen's
    normal(Tensor, Tensor) computes (checked against the original under the same seed in rollout_fixture)."""

    def __init__(self):
        import torch.distributions.normal as dn
        self.dn, self.log = dn, []
        self.o_sn, self.o_randn, self.o_normal = dn._standard_normal, torch.randn, torch.normal

    def __enter__(self):
        def sn(shape, dtype, device):
            z = self.o_sn(shape, dtype, device); self.log.append(("eps", z.clone())); return z

        def randn(*a, **k):
            z = self.o_randn(*a, **k); self.log.append(("randn", z.clone())); return z

        def normal(loc, scale, *a, **k):
            z = torch.empty(loc.shape).normal_(); self.log.append(("normal", z.clone())); return z * scale + loc
        self.dn._standard_normal, torch.randn, torch.normal = sn, randn, normal
        return self

    def __exit__(self, *a):
        self.dn._standard_normal, torch.randn, torch.normal = self.o_sn, self.o_randn, self.o_normal

    def take(self, kind):
        for i, (k, z) in enumerate(self.log):
            if k == kind:
                return self.log.pop(i)[1]
        return None


def rollout_fixture():
    """Frame-loop body of the reference under recorded noise:
    off-policy: replay_buffer.normalize -> initial_act | noise_action -> process_action (ddpg.py:202-207, off_policy.py:50-54,
    89-106, ddpg_acm.py:40-50), one observation at a time as the reference does; on-policy: buffer.normalize -> Actor.act ->
    process_action (a2c.py:165-167, basic_model.py:32-51, on_policy.py:34-53)."""
    out = {}
    E = 12
    cases = [
        # published SPP-SAC flags: min-max denormalisation, lim = 1
        ("sac_hopper", rl.SAC_AcM, "Hopper-v2", dict(min_max_denormalize=True, denormalize_actor_out=True), "acm"),
        # bounded observation space, no denormalisation: lim = obs.high = [1, 1, 8] (acm.py:102-108)
        ("sac_pendulum", rl.SAC_AcM, "Pendulum-v0", dict(min_max_denormalize=False, denormalize_actor_out=False), "acm"),
        # mean-std denormalisation: lim = 10; BasicAcM as in train/spp_ddpg_hcheetah.py
        ("ddpg_hcheetah", rl.DDPG_AcM, "HalfCheetah-v2", dict(min_max_denormalize=False, denormalize_actor_out=True), "basic"),
    ]
    for ci, (name, cls, env, flags, kind) in enumerate(cases):
        torch.manual_seed(100 + ci); np.random.seed(100 + ci)
        m = cls(env_name=env, acm_pre_train_samples=100, acm_val_buffer_size=None, buffer_size=1000, tensorboard_dir=None,
                log_dir=None, **flags)
        ob, ac = m.ob_dim, m.ac_dim
        algo = "sac" if cls is rl.SAC_AcM else "ddpg"
        if kind == "basic":
            m.acm = BasicAcM(2 * ob, ac, False)
        s0 = init_state(algo, ob, ac, 40 + ci, kind, False)
        if kind == "basic":
            s0["acm.t"][:] = 0.7; s0["acm.t1"][:] = np.linspace(0.5, 1.5, ac)
        nets = ["actor", "acm"]
        load_nets(m, s0, nets)
        mn, mx, mean, std = make_stats(ob, 3 + ci, flags["min_max_denormalize"])
        for k, v in (("min_obs", mn), ("max_obs", mx), ("obs_mean", mean), ("obs_std", std)):
            setattr(m.replay_buffer, k, torch.from_numpy(v))
        rng = np.random.RandomState(5 + ci)
        obs = (rng.rand(E, ob) * (mx - mn) + mn).astype(np.float32)
        rec = {k: [] for k in ("eps", "noise", "target", "acm", "init_noise", "init_target", "init_acm", "det_target", "det_acm")}
        for e in range(E):
            o = m.replay_buffer.normalize(torch.from_numpy(obs[e:e + 1]))
            with _Recorder() as r:
                a = m.noise_action(o, m.act_noise)
                rec["noise"].append(r.take("randn").numpy().reshape(-1))
                z = r.take("eps")
                if z is not None:
                    rec["eps"].append(z.numpy().reshape(-1))
            rec["target"].append(a.numpy().reshape(-1).copy())
            rec["acm"].append(np.asarray(m.process_action(a, o)).reshape(-1).copy())
            with _Recorder() as r:
                a = m.initial_act(o)
                rec["init_noise"].append(r.take("randn").numpy().reshape(-1))
            rec["init_target"].append(a.numpy().reshape(-1).copy())
            rec["init_acm"].append(np.asarray(m.process_action(a, o)).reshape(-1).copy())
            a = m.noise_action(o, act_noise=0, deterministic=True)      # what test() does (ddpg.py:385-410)
            rec["det_target"].append(a.numpy().reshape(-1).copy())
            rec["det_acm"].append(np.asarray(m.process_action(a, o)).reshape(-1).copy())
        out[name + ":obs"] = obs
        for k, v in rec.items():
            if v:
                out[name + ":" + k] = np.array(v, np.float32)
        out[name + ":actor_lim"] = np.broadcast_to(np.asarray(m.actor_ac_lim, np.float32), (ob,)).copy()
        out[name + ":acm_lim"] = np.asarray(m.ac_lim, np.float32).reshape(-1)
        out[name + ":act_noise"] = np.array(m.act_noise)
        out[name + ":meta"] = np.array([ob, ac, 40 + ci, 3 + ci, int(flags["min_max_denormalize"]), int(flags["denormalize_actor_out"])])
        print("rollout", name, "lim", out[name + ":actor_lim"][:3], "act_noise", m.act_noise, "target[0]", rec["target"][0][:3])

    # torch.normal(loc, scale) == z * scale + loc with z = empty.normal_() under the same generator state (the recorder's claim)
    loc, sc = torch.randn(1, 17), torch.rand(1, 17) + 0.1
    torch.manual_seed(5); a = torch.normal(loc, sc)
    torch.manual_seed(5); b = torch.empty(1, 17).normal_() * sc + loc
    assert torch.equal(a, b)

    # on-policy
    torch.manual_seed(7); np.random.seed(7)
    m = rl.PPO_AcM(env_name="Walker2d-v2", acm_pre_train_samples=100, acm_pre_train_epochs=1, iterations=1, batch_size=100,
                   denormalize_actor_out=True, min_max_denormalize=True, custom_loss=0.1, obs_norm=True, tensorboard_dir=None,
                   log_dir=None, acm_val_buffer_size=None)
    ob, ac = m.ob_dim, m.ac_dim
    mn, mx, mean, std = make_stats(ob, 9, True)
    m.min_obs, m.max_obs = torch.from_numpy(mn), torch.from_numpy(mx)
    m.replay_buffer.min_obs, m.replay_buffer.max_obs = m.min_obs, m.max_obs
    buf = MemoryAcM(obs_mean=m.obs_mean, obs_std=m.obs_std, device=m.device, alpha=m.obs_norm_alpha, max_obs=m.max_obs,
                    min_obs=m.min_obs, min_max_denormalize=m.min_max_denormalize)
    m.buffer = buf
    rng = np.random.RandomState(19)
    obs = (rng.rand(E, ob) * (mx - mn) + mn).astype(np.float32)
    rec = {k: [] for k in ("noise", "action", "logp", "acm")}
    for e in range(E):
        o = buf.normalize(torch.from_numpy(obs[e:e + 1]))
        with _Recorder() as r:
            a, lp = m.actor.act(o)
            rec["noise"].append(r.take("normal").numpy().reshape(-1))
        rec["action"].append(a.numpy().reshape(-1).copy()); rec["logp"].append(float(lp))
        rec["acm"].append(np.asarray(m.process_action(a, o)).reshape(-1).copy())
    out["ppo_walker:obs"] = obs
    for k, v in rec.items():
        out["ppo_walker:" + k] = np.array(v, np.float32)
    for net in ("actor", "acm"):
        for k, v in getattr(m, net).state_dict().items():
            out["ppo_walker:" + net + "." + k] = v.detach().numpy().copy()
    out["ppo_walker:min_obs"], out["ppo_walker:max_obs"] = mn, mx
    out["ppo_walker:actor_lim"] = np.array(float(m.actor_ac_lim)); out["ppo_walker:acm_lim"] = m.ac_lim.numpy()
    out["torch_version"] = np.array(torch.__version__)
    np.savez_compressed(os.path.join(HERE, "rollout_steps.npz"), **out)
    print("rollout_steps: ppo logp[0]", rec["logp"][0])


def ppo_plain_fixture():
    """PPO_AcM(custom_loss=0).update_actor -> PPO.update_actor (ppo.py:152-192) on the rollout, advantages and pre-update actor of
    ppo_walker.npz, with the same injected shuffles: post-update actor, the raw loss sums and the epoch counter."""
    g = np.load(os.path.join(HERE, "ppo_walker.npz"))
    hp = g["hp"]
    torch.manual_seed(0); np.random.seed(0)
    m = rl.PPO_AcM(env_name="Walker2d-v2", gamma=float(hp[0]), acm_pre_train_samples=100, acm_pre_train_epochs=1, iterations=1, batch_size=700,
                   actor_lr=float(hp[6]), critic_lr=float(hp[7]), kl_div_threshold=float(hp[3]), max_ppo_epochs=int(hp[4]),
                   ppo_batch_size=int(hp[5]), denormalize_actor_out=True, min_max_denormalize=True, custom_loss=0.0, obs_norm=True,
                   tensorboard_dir=None, log_dir=None, acm_val_buffer_size=None, norm_closs=False)
    m.actor.load_state_dict({k[len("pre:actor."):]: torch.from_numpy(g[k].copy()) for k in g.files if k.startswith("pre:actor.")})
    m.actor = m.actor      # optimiser on the loaded parameters
    t = torch.from_numpy
    buf = MemoryAcM(obs_mean=t(g["obs_mean"]), obs_std=t(g["obs_std"]), device=m.device, alpha=m.obs_norm_alpha, max_obs=t(g["max_obs"]),
                    min_obs=t(g["min_obs"]), min_max_denormalize=True)
    chain, joints = g["chain"], set(int(j) for j in g["joints"])
    ts = 0
    prev = buf.add_obs(t(chain[0:1]))
    for i in range(1, len(chain)):
        if i in joints:
            buf.end_rollout()
            prev = buf.add_obs(t(chain[i:i + 1]))
            continue
        nxt = buf.add_obs(t(chain[i:i + 1]))
        buf.add_timestep(prev, nxt, t(g["actions"][ts:ts + 1]), t(g["logp"][ts:ts + 1]), float(g["rewards"][ts]), bool(g["done"][ts]), bool(g["end"][ts]))
        prev = nxt; ts += 1
    buf.end_rollout()
    assert ts == len(g["actions"]) == len(buf)
    perms = [t(p_) for p_ in g["perms"]]
    calls = [0]
    orig = torch.randperm

    def fake(n, *a, **k):
        calls[0] += 1
        return perms[(calls[0] - 1) // 2]
    torch.randperm = fake
    c0 = m.kl_div_updates_counter
    m.update_actor(t(g["adv"].copy()), buf)
    torch.randperm = orig
    out = {"epochs_counter": np.array(m.kl_div_updates_counter - c0),
           "losses": np.array([m.loss["actor"], m.loss["entropy"], m.loss["sum"]], np.float64)}
    for k, v in m.actor.state_dict().items():
        out["post:actor." + k] = v.detach().numpy().copy()
    np.savez_compressed(os.path.join(HERE, "ppo_plain.npz"), **out)
    print("ppo_plain: epochs counter", out["epochs_counter"], "losses", out["losses"])


def a2c_fixture():
    """A2C_AcM (on_policy.py:100-124,133-155) for two iterations on ppo_walker's rollout: A2C.update_critic -> q - V advantages ->
    update_actor_acm with the gradients that are never zeroed.  The second iteration re-records the log-probs under the updated actor
    (what a fresh rollout does) on the same observations / actions."""
    g = np.load(os.path.join(HERE, "ppo_walker.npz"))
    hp = g["hp"]
    torch.manual_seed(0); np.random.seed(0)
    m = rl.A2C_AcM(env_name="Walker2d-v2", gamma=float(hp[0]), acm_pre_train_samples=100, acm_pre_train_epochs=1, iterations=1, batch_size=700,
                   actor_lr=float(hp[6]), critic_lr=float(hp[7]), denormalize_actor_out=True, min_max_denormalize=True, custom_loss=0.1,
                   obs_norm=True, tensorboard_dir=None, log_dir=None, acm_val_buffer_size=None, norm_closs=False)
    t = torch.from_numpy
    for net in ("actor", "critic"):
        getattr(m, net).load_state_dict({k[len("pre:" + net + "."):]: t(g[k].copy()) for k in g.files if k.startswith("pre:" + net + ".")})
    m.actor = m.actor; m.critic = m.critic      # optimisers on the loaded parameters
    chain, joints = g["chain"], set(int(j) for j in g["joints"])
    out = {}

    def build():
        buf = MemoryAcM(obs_mean=t(g["obs_mean"]), obs_std=t(g["obs_std"]), device=m.device, alpha=m.obs_norm_alpha, max_obs=t(g["max_obs"]),
                        min_obs=t(g["min_obs"]), min_max_denormalize=True)
        ts = 0
        cur = t(chain[0:1].copy())
        prev = buf.add_obs(cur)
        for i in range(1, len(chain)):
            if i in joints:
                buf.end_rollout()
                cur = t(chain[i:i + 1].copy())
                prev = buf.add_obs(cur)
                continue
            a = t(g["actions"][ts:ts + 1].copy())
            logp = m.actor.get_actions_dist(buf.normalize(cur)).log_prob(torch.squeeze(a))      # Actor.act's expression, graph kept
            cur = t(chain[i:i + 1].copy())
            nxt = buf.add_obs(cur)
            buf.add_timestep(prev, nxt, a, logp, float(g["rewards"][ts]), bool(g["done"][ts]), bool(g["end"][ts]))
            prev = nxt; ts += 1
        buf.end_rollout()
        assert ts == len(g["actions"]) == len(buf)
        return buf

    for it in (1, 2):
        buf = build()
        out["logp%d" % it] = torch.cat(buf.action_logprobs).detach().numpy().copy()
        adv = m.update_critic(buf)
        out["adv%d" % it] = adv.numpy().copy()
        out["critic_loss%d" % it] = np.array(m.loss["critic"])
        for k, v in m.critic.state_dict().items():
            out["fit%d:critic.%s" % (it, k)] = v.detach().numpy().copy()
        m.update_actor(adv, buf)
        out["losses%d" % it] = np.array([m.loss["actor"], m.loss["dist"], m.loss["policy"]], np.float64)
        for k, v in m.actor.state_dict().items():
            out["post%d:actor.%s" % (it, k)] = v.detach().numpy().copy()
        st = m.actor_optimizer.state_dict()["state"]
        for i, (k, _) in enumerate(m.actor.named_parameters()):
            out["post%d:actor.%s#m" % (it, k)] = st[i]["exp_avg"].numpy().copy()
            out["post%d:actor.%s#v" % (it, k)] = st[i]["exp_avg_sq"].numpy().copy()
    np.savez_compressed(os.path.join(HERE, "a2c_acm.npz"), **out)
    print("a2c_acm: losses", out["losses1"], out["losses2"])


CONFIG1_KW = dict(env_name="Pendulum-v0", update_batch_size=256, acm_pre_train_samples=400, acm_pre_train_epochs=2, acm_val_buffer_size=None,
                  buffer_size=5000, iterations=2, batch_size=200, grad_steps=10, update_freq=50, random_frames=100, acm_update_freq=100,
                  acm_epochs=1, acm_batch_size=128, custom_loss=0.2, acm_critic=True, norm_closs=False, denormalize_actor_out=True,
                  min_max_denormalize=True, gamma=0.99, stats_freq=1, verbose=0)


def config1_fixture():
    """SURVEY 8d config 1 (reference plumbing on CPU, checks the boundary end to end): the reference's own SAC_AcM(...).pre_train();
    .train() on the Pendulum stub.  The mirror, started from the same weights under the same numpy / torch seeds, must draw the same
    replay indices, leave the same ring cursors and index arrays, and end with the same weights to 1e-5."""
    import time
    torch.manual_seed(11); np.random.seed(11)
    m = rl.SAC_AcM(tensorboard_dir=None, log_dir=None, **CONFIG1_KW)
    out = {}
    nets = ["actor", "critic_1", "critic_2", "acm"]
    for net in nets:
        for k, v in getattr(m, net).state_dict().items():
            out["init:" + net + "." + k] = v.detach().numpy().copy()
    # The reference's OWN sensitivity: the same run with ONE initial weight moved by one unit in the last place.  400 frames of closed
    # loop (the actions drive the pendulum that produces the next observations) and 80 Adam steps amplify that; the mirror cannot be
    # expected to track the reference more closely than the reference tracks itself, so the test bounds each tensor by
    # max(1e-5, 4 x this).
    torch.manual_seed(11); np.random.seed(11)
    m2 = rl.SAC_AcM(tensorboard_dir=None, log_dir=None, **CONFIG1_KW)
    for net in nets:
        getattr(m2, net).load_state_dict(getattr(m, net).state_dict())
    m2.critic_1_targ.load_state_dict(m.critic_1_targ.state_dict()); m2.critic_2_targ.load_state_dict(m.critic_2_targ.state_dict())
    with torch.no_grad():
        w = m2.actor.fc1.weight
        w[0, 0] = torch.nextafter(w[0, 0], torch.tensor(float("inf")))
    torch.manual_seed(12); np.random.seed(12)
    m2.pre_train(); m2.train()
    drawn = []
    orig = np.random.randint

    def rec(lo, hi=None, size=None, *a, **k):
        r = orig(lo, hi, size, *a, **k)
        drawn.append(np.concatenate([[hi, np.size(r)], np.ravel(r)]).astype(np.int64))
        return r
    torch.manual_seed(12); np.random.seed(12)
    np.random.randint = rec
    t0 = time.perf_counter()
    m.pre_train()
    t1 = time.perf_counter()
    m.train()
    t2 = time.perf_counter()
    np.random.randint = orig
    rb = m.replay_buffer
    L = rb.current_len
    out.update(obs_idx=np.array(rb.obs_idx), ts_idx=np.array(rb.ts_idx), current_len=np.array(L), ring_obs_idx=rb._obs_idx[:L].copy(),
               ring_next_obs_idx=rb._next_obs_idx[:L].copy(), ring_obs=rb._obs[: rb.obs_idx].astype(np.float32), ring_rew=rb._rewards[:L].copy(),
               ring_done=rb._done[:L].copy(), ring_aacm=rb._actions_acm[:L].astype(np.float32), ring_act=rb._actions[:L].astype(np.float32),
               drawn=np.concatenate(drawn), n_draws=np.array(len(drawn)), frames=np.array(m.stats_logger.frames),
               rollouts=np.array(m.stats_logger.rollouts), running_return=np.array(m.stats_logger.running_return),
               alpha=np.array(m.alpha), min_obs=m.min_obs.numpy(), max_obs=m.max_obs.numpy(), obs_mean=m.obs_mean.numpy(), obs_std=m.obs_std.numpy(),
               loss_keys=np.array(sorted(m.loss.keys())), loss_vals=np.array([m.loss[k] for k in sorted(m.loss.keys())], np.float64),
               ref_seconds=np.array([t1 - t0, t2 - t1]))
    from tests.parity_util import relnorm
    worst = 0.0
    for net in nets + ["critic_1_targ", "critic_2_targ"]:
        for k, v in getattr(m, net).state_dict().items():
            out["final:" + net + "." + k] = v.detach().numpy().copy()
            e = relnorm(getattr(m2, net).state_dict()[k].numpy(), v.detach().numpy())
            out["sens:" + net + "." + k] = np.array(e)
            worst = max(worst, e)
    rb2 = m2.replay_buffer
    out["sens_obs"] = np.array(np.abs(rb2._obs[: rb.obs_idx] - rb._obs[: rb.obs_idx]).max())
    out["sens_aacm"] = np.array(np.abs(rb2._actions_acm[:L] - rb._actions_acm[:L]).max())
    out["sens_act"] = np.array(np.abs(rb2._actions[:L] - rb._actions[:L]).max())
    print("config1: reference vs reference with one weight moved by 1 ulp: worst weight relnorm %.2e, obs %.2e, acm action %.2e, target %.2e"
          % (worst, out["sens_obs"], out["sens_aacm"], out["sens_act"]))
    np.savez_compressed(os.path.join(HERE, "config1_pendulum.npz"), **out)
    print("config1: frames", m.stats_logger.frames, "len", L, "draws", len(drawn), "running_return", m.stats_logger.running_return,
          "reference: pre_train %.2f s, train %.2f s = %.1f frames/s" % (t1 - t0, t2 - t1, m.stats_logger.frames / (t2 - t1)))


PKL_MODELS = [      # (file, class, env, BasicAcM?, norm_closs) -- the cells of notebooks/load_and_test.ipynb
    ("hopper_sac_acm_model.pkl", "SAC_AcM", "Hopper-v2", False, False),
    ("hcheetah_sac_acm_model.pkl", "SAC_AcM", "HalfCheetah-v2", False, False),
    ("ant3m_sac_acm_model.pkl", "SAC_AcM", "Ant-v2", False, False),
    ("hcheetah_ddpg_acm_model.pkl", "DDPG_AcM", "HalfCheetah-v2", True, False),
    ("hopper_ddpg_acm_model.pkl", "DDPG_AcM", "Hopper-v2", True, False),
    ("ant3m_ddpg_acm_model.pkl", "DDPG_AcM", "Ant-v2", True, False),
    ("hcheetah_ppo_acm.pkl", "PPO_AcM", "HalfCheetah-v2", False, True),
    ("hopper_ppo_acm.pkl", "PPO_AcM", "Hopper-v2", False, True),
    ("walker_ppo_acm.pkl", "PPO_AcM", "Walker2d-v2", False, True),
]


def pkl_fixture():
    """The notebook's load-and-act flow (notebooks/load_and_test.ipynb cells 2-8, 23-27, 42-46) on every shipped model, with
    fixed observations instead of a MuJoCo episode: deterministic actor output and the ACM action that would go to the env."""
    out = {}
    E = 6
    mdir = os.path.join(HERE, "models")
    for fi, (fname, cls, env, basic, norm_closs) in enumerate(PKL_MODELS):
        torch.manual_seed(0)
        kw = dict(env_name=env, custom_loss=1, denormalize_actor_out=True, min_max_denormalize=True, norm_closs=norm_closs,
                  tensorboard_dir=None, log_dir=None)
        if cls != "PPO_AcM":
            kw.update(acm_critic=True, buffer_size=1000)
        m = getattr(rl, cls)(**kw)
        if basic:
            m.acm = BasicAcM(m.ob_dim * 2, m.ac_dim, False)
        m.load(os.path.join(mdir, fname))
        rng = np.random.RandomState(77 + fi)
        mn, mx = m.min_obs.numpy(), m.max_obs.numpy()
        obs = (rng.rand(E, m.ob_dim) * (mx - mn) + mn).astype(np.float32)
        tg, aa = [], []
        if cls == "PPO_AcM":
            m.replay_buffer.min_obs, m.replay_buffer.max_obs = m.min_obs, m.max_obs
            m.buffer = MemoryAcM(obs_mean=m.obs_mean, obs_std=m.obs_std, device=m.device, alpha=m.obs_norm_alpha, max_obs=m.max_obs,
                                 min_obs=m.min_obs, min_max_denormalize=m.min_max_denormalize)
            for e in range(E):
                o = m.buffer.normalize(m.process_obs(obs[e]))
                a, _ = m.actor.act(o, deterministic=True)
                tg.append(a.numpy().reshape(-1).copy()); aa.append(np.asarray(m.process_action(a, o)).reshape(-1).copy())
        else:
            for e in range(E):
                o = m.replay_buffer.normalize(m.process_obs(obs[e]))
                a = m.noise_action(o, act_noise=0, deterministic=True)
                tg.append(a.numpy().reshape(-1).copy()); aa.append(np.asarray(m.process_action(a, o)).reshape(-1).copy())
        out[fname + ":obs"] = obs; out[fname + ":target"] = np.array(tg, np.float32); out[fname + ":acm"] = np.array(aa, np.float32)
        print("pkl", fname, "ob", m.ob_dim, "ac", m.ac_dim, "target[0][:3]", tg[0][:3], "acm[0]", aa[0][:3])
    out["torch_version"] = np.array(torch.__version__)
    np.savez_compressed(os.path.join(HERE, "pkl_actions.npz"), **out)


if __name__ == "__main__":
    if "--only-config1" in sys.argv:
        config1_fixture(); sys.exit(0)
    if "--only-ppo-plain" in sys.argv:
        ppo_plain_fixture(); sys.exit(0)
    if "--only-pkl" in sys.argv:
        pkl_fixture(); sys.exit(0)
    if "--only-rollout" in sys.argv:
        rollout_fixture(); sys.exit(0)
    if "--only-a2c" in sys.argv:
        a2c_fixture(); sys.exit(0)
    sac_fixture()
    ddpg_fixture()
    ring_fixture()
    acm_regress_fixture()
    acm_epochs_fixture()
    ppo_fixture()
    rollout_fixture()
    ppo_plain_fixture()
    a2c_fixture()
    config1_fixture()
    pkl_fixture()
