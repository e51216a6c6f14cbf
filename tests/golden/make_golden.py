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


if __name__ == "__main__":
    sac_fixture()
    ddpg_fixture()
    ring_fixture()
    acm_regress_fixture()
    acm_epochs_fixture()
    ppo_fixture()
