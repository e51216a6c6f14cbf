"""-m gpu: the SPP-PPO kernels against the golden fixture produced by the unmodified reference (PPO_AcM on Walker2d
shapes: critic fit, q-values + GAE incl. truncation bootstrap, advantage normalisation, clipped-ratio actor epochs with
KL early stop) and against the oracle on a larger synthetic batch laid out step-major ([T][E], stride E)."""
import os

import numpy as np
import pytest
import torch

from oracle import ppo as P
from oracle.norm import NormStats, denormalize, normalize
from spp_rl_b200.ppo import PpoPolicy
from tests.parity_util import relnorm

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _trajectories(end):
    starts, lens, s = [], [], 0
    for i, e in enumerate(end):
        if e:
            starts.append(s); lens.append(i + 1 - s); s = i + 1
    assert s == len(end), "every trajectory must finish with end = 1"
    return np.array(starts, np.int64), np.array(lens, np.int64)


@pytest.mark.parametrize("critic_tc", [True, False], ids=["critic-tcgen05", "critic-ffma"])
def test_ppo_update_matches_reference_fixture(critic_tc):
    g = np.load(os.path.join(G, "ppo_walker.npz"))
    gamma, lam, eps_clip, kl_thr, max_ep, bs, a_lr, c_lr, ent, closs, ntu, nupt = [float(x) for x in g["hp"]]
    ob, ac = g["chain"].shape[1], g["actions_acm"].shape[1]
    oi, ni = P.chain_views(len(g["chain"]), list(g["joints"]))
    obs, nobs = g["chain"][oi], g["chain"][ni]
    N = obs.shape[0]
    pol = PpoPolicy(ob, ac, max_rows=N, max_batch_rows=int(bs), min_max_denormalize=True, norm_closs=False, gamma=gamma, gae_lambda=lam,
                    ppo_epsilon=eps_clip, entropy_coef=ent, custom_loss=closs, actor_lr=a_lr, critic_lr=c_lr)
    pol.set_critic_path(critic_tc)
    pol.set_limits(float(g["actor_lim"]))
    pol.set_norm_stats(g["min_obs"], g["max_obs"], g["obs_mean"], g["obs_std"])
    for net in ("actor", "critic"):
        pol.load_state_dict(net, {k[len("pre:" + net) + 1:]: g[k] for k in g.files if k.startswith("pre:" + net + ".")})
    ts, tl = _trajectories(g["end"])
    pol.load_rollout(obs, nobs, g["actions"], g["logp"], g["rewards"], g["done"], g["end"], ts, tl)
    loss = pol.update_critic(int(ntu), int(nupt))
    assert loss == pytest.approx(float(g["critic_loss"]), rel=1e-5)
    sd = pol.state_dict("critic")
    for k, v in sd.items():
        assert relnorm(v, g["fit:critic." + k]) < 1e-5, (k, relnorm(v, g["fit:critic." + k]))
    adv = pol.advantages()
    assert np.abs(adv - g["adv"]).max() < 1e-5 * max(1.0, np.abs(g["adv"]).max())
    pol.normalize_adv()
    losses, epochs, kl = pol.update_actor(g["perms"], int(bs), kl_thr, int(max_ep))
    assert epochs + 1 == int(g["epochs_run"])          # the reference counts i + 1 after its early-stop break
    ref = g["actor_losses"]
    for key, r in zip(("actor", "entropy", "policy", "dist"), ref):
        assert losses[key] == pytest.approx(float(r), rel=1e-5), key
    sd = pol.state_dict("actor")
    for k, v in sd.items():
        assert relnorm(v, g["post:actor." + k]) < 1e-5, (k, relnorm(v, g["post:actor." + k]))
    pol.close()


def test_plain_ppo_actor_epochs_match_reference_fixture():
    """P6: custom_loss == 0 -> PPO.update_actor (ppo.py:152-192): no distance term, log-prob of the stored actions, raw loss sums."""
    g, gp = np.load(os.path.join(G, "ppo_walker.npz")), np.load(os.path.join(G, "ppo_plain.npz"))
    gamma, lam, eps_clip, kl_thr, max_ep, bs, a_lr, c_lr, ent, closs, ntu, nupt = [float(x) for x in g["hp"]]
    ob, ac = g["chain"].shape[1], g["actions_acm"].shape[1]
    oi, ni = P.chain_views(len(g["chain"]), list(g["joints"]))
    obs, nobs = g["chain"][oi], g["chain"][ni]
    N = obs.shape[0]
    pol = PpoPolicy(ob, ac, max_rows=N, max_batch_rows=int(bs), min_max_denormalize=True, norm_closs=False, gamma=gamma, gae_lambda=lam,
                    ppo_epsilon=eps_clip, entropy_coef=ent, custom_loss=0.0, actor_lr=a_lr, critic_lr=c_lr)
    pol.set_actor_mode(plain_ppo=True)
    pol.set_limits(float(g["actor_lim"]))
    pol.set_norm_stats(g["min_obs"], g["max_obs"], g["obs_mean"], g["obs_std"])
    for net in ("actor", "critic"):
        pol.load_state_dict(net, {k[len("pre:" + net) + 1:]: g[k] for k in g.files if k.startswith("pre:" + net + ".")})
    ts, tl = _trajectories(g["end"])
    pol.load_rollout(obs, nobs, g["actions"], g["logp"], g["rewards"], g["done"], g["end"], ts, tl)
    pol.load_advantages(g["adv"])
    pol.normalize_adv()
    losses, epochs, kl = pol.update_actor(g["perms"], int(bs), kl_thr, int(max_ep))
    assert min(epochs + 1, int(max_ep)) == int(gp["epochs_counter"])
    for key, r in zip(("actor", "entropy", "policy"), gp["losses"]):
        assert losses[key] == pytest.approx(float(r), rel=1e-5), key
    sd = pol.state_dict("actor")
    for k, v in sd.items():
        assert relnorm(v, gp["post:actor." + k]) < 1e-5, (k, relnorm(v, gp["post:actor." + k]))
    pol.close()


def test_a2c_acm_two_iterations_match_reference_fixture():
    """(f)4: A2C_AcM (on_policy.py:100-124): A2C.update_critic, advantages q - V(s), one full-batch policy-gradient step per iteration on
    gradients that are never zeroed -- two iterations of the unmodified reference (tests/golden/a2c_acm.npz)."""
    g, ga = np.load(os.path.join(G, "ppo_walker.npz")), np.load(os.path.join(G, "a2c_acm.npz"))
    gamma, lam, eps_clip, kl_thr, max_ep, bs, a_lr, c_lr, ent, closs, ntu, nupt = [float(x) for x in g["hp"]]
    ob, ac = g["chain"].shape[1], g["actions_acm"].shape[1]
    oi, ni = P.chain_views(len(g["chain"]), list(g["joints"]))
    obs, nobs = g["chain"][oi], g["chain"][ni]
    N = obs.shape[0]
    pol = PpoPolicy(ob, ac, max_rows=N, max_batch_rows=N, min_max_denormalize=True, norm_closs=False, gamma=gamma, gae_lambda=lam,
                    ppo_epsilon=eps_clip, entropy_coef=0.3, custom_loss=0.1, actor_lr=a_lr, critic_lr=c_lr)     # the entropy bonus must be ignored
    pol.set_actor_mode(a2c=True)
    pol.set_limits(float(g["actor_lim"]))
    pol.set_norm_stats(g["min_obs"], g["max_obs"], g["obs_mean"], g["obs_std"])
    for net in ("actor", "critic"):
        pol.load_state_dict(net, {k[len("pre:" + net) + 1:]: g[k] for k in g.files if k.startswith("pre:" + net + ".")})
    ts, tl = _trajectories(g["end"])
    for it in (1, 2):
        pol.load_rollout(obs, nobs, g["actions"], ga["logp%d" % it], g["rewards"], g["done"], g["end"], ts, tl)
        loss = pol.update_critic(int(ntu), int(nupt))
        assert loss == pytest.approx(float(ga["critic_loss%d" % it]), rel=1e-5)
        adv = pol.advantages()
        ref = ga["adv%d" % it]
        assert np.abs(adv - ref).max() < 1e-5 * max(1.0, np.abs(ref).max())
        pol.load_advantages(ref)
        actor_loss, dist_sum = pol.a2c_actor_step(accumulate=True, normalize_adv=True)
        dist = dist_sum / (N * ob)
        for v, r in zip((actor_loss, dist, actor_loss + 0.1 * dist), ga["losses%d" % it]):
            assert v == pytest.approx(float(r), rel=1e-5, abs=1e-7)
        sd = pol.state_dict("actor")
        for k, v in sd.items():
            assert relnorm(v, ga["post%d:actor.%s" % (it, k)]) < 1e-5, (it, k, relnorm(v, ga["post%d:actor.%s" % (it, k)]))
    pol.close()


@pytest.mark.parametrize("E,T,mb,n_epochs", [(96, 40, 1000, 3), (1024, 1024, 65536, 1)])
def test_ppo_step_major_large_batch_matches_oracle(E, T, mb, n_epochs):
    """E environments x T steps in step-major order (row = t * E + e): GAE strides by E; several CTAs per phase.  The second case is
    config 4's shape class: 1 048 576 rows, 65 536-row minibatches -- fp32 reductions over a million rows, every CTA's
    partial-gradient slot and the [T][E] stride against the oracle.  Tolerances as in the small case: north_star's 1e-5 (critic loss,
    advantages relative to their range, losses, post-epoch actor weights norm-relative)."""
    ob, ac = 17, 6
    N = E * T
    rng = np.random.RandomState(3)
    mn, mx = (-rng.rand(ob) * 2 - 0.5).astype(np.float32), (rng.rand(ob) * 2 + 0.5).astype(np.float32)
    obs = (rng.rand(N, ob) * (mx - mn) + mn).astype(np.float32)
    nobs = (obs + 0.05 * rng.randn(N, ob)).astype(np.float32)
    act = (0.5 * rng.randn(N, ob)).astype(np.float32)
    logp = (-10 + rng.randn(N)).astype(np.float32)
    rew = rng.randn(N).astype(np.float32)
    done = (rng.rand(N) < 0.02).astype(np.float32)
    end = done.copy()
    end[rng.rand(N) < 0.02] = 1
    end[(T - 1) * E:] = 1                                   # every environment's trajectory ends at the last step
    s = {}
    def lin(net, name, o, i):
        b = 1 / np.sqrt(i)
        s[net + "." + name + ".weight"] = torch.from_numpy(rng.uniform(-b, b, (o, i)).astype(np.float32))
        s[net + "." + name + ".bias"] = torch.from_numpy(rng.uniform(-b, b, (o,)).astype(np.float32))
    for net, out in (("actor", ob), ("critic", 1)):
        lin(net, "fc1", 64, ob); lin(net, "fc2", 64, 64); lin(net, "fc3", out, 64)
    s["actor.log_scale"] = torch.full((ob,), -1.34)
    pol = PpoPolicy(ob, ac, max_rows=N, max_batch_rows=max(1024, mb), min_max_denormalize=True, gamma=0.99, gae_lambda=0.95, custom_loss=0.1,
                    entropy_coef=0.01)
    pol.set_norm_stats(mn, mx)
    for net in ("actor", "critic"):
        pol.load_state_dict(net, {k[len(net) + 1:]: v for k, v in s.items() if k.startswith(net + ".")})
    pol.load_rollout(obs, nobs, act, logp, rew, done, end, np.arange(E), np.full(E, T), traj_stride=E)
    st = NormStats(True, torch.from_numpy(mn), torch.from_numpy(mx))
    x, xn = normalize(st, torch.from_numpy(obs), True), normalize(st, torch.from_numpy(nobs), True)
    tr, td, te = torch.from_numpy(rew), torch.from_numpy(done), torch.from_numpy(end)
    ref_loss = P.update_critic(s, x, xn, tr, td, 0.99, 3e-4, 2, 3)
    assert pol.update_critic(2, 3) == pytest.approx(ref_loss, rel=1e-5)
    # oracle GAE per environment (time-major view of the step-major rows)
    v = torch.zeros(N); nv = torch.zeros(N)
    from oracle import nets
    from oracle.offpolicy import sub
    v = nets.ppo_critic_fwd(sub(s, "critic"), x)[0].squeeze(-1); nv = nets.ppo_critic_fwd(sub(s, "critic"), xn)[0].squeeze(-1)
    q = P.q_values(tr, td, nv, 0.99)
    adv_ref = torch.empty(N)
    for e in range(min(E, 96)):
        rows = torch.arange(e, N, E)
        adv_ref[rows] = P.gae(q[rows], v[rows], nv[rows], td[rows], te[rows], 0.99, 0.95)
    if E > 96:      # the same recurrence (ppo.py:139-148), one vector op over the environments per time step; checked against P.gae above
        vec = torch.empty(N)
        carry = torch.zeros(E)
        delta = q - v
        boot = (nv.double() * (0.99 * 0.95)).float()
        for t in range(T - 1, -1, -1):
            r = slice(t * E, (t + 1) * E)
            carry = torch.where(td[r] != 0, delta[r], torch.where(te[r] != 0, boot[r] + delta[r], carry * np.float32(0.99 * 0.95) + delta[r]))
            vec[r] = carry
        first = torch.cat([torch.arange(e, N, E) for e in range(96)])
        assert torch.equal(vec[first], adv_ref[first])
        adv_ref = vec
    adv = pol.advantages()
    assert np.abs(adv - adv_ref.numpy()).max() < 1e-5 * max(1.0, float(adv_ref.abs().max()))
    pol.normalize_adv()
    advn = P.normalize_adv(adv_ref)
    perms = np.stack([rng.permutation(N) for _ in range(n_epochs)]).astype(np.int64)
    tot, epochs, kl = P.update_actor_acm(s, x, denormalize(st, torch.from_numpy(act)), denormalize(st, xn), torch.from_numpy(logp), advn,
                                         [torch.from_numpy(p) for p in perms], 1.0, 3e-4, 0.2, 1e9, n_epochs, mb, 0.01, 0.1)
    losses, epochs_c, kl_c = pol.update_actor(perms, mb, 1e9, n_epochs)
    assert epochs_c == epochs
    assert kl_c == pytest.approx(kl, rel=1e-3, abs=1e-5)
    for key in ("actor", "entropy", "policy", "dist"):
        assert losses[key] == pytest.approx(tot[key], rel=1e-5, abs=1e-6), key
    sd = pol.state_dict("actor")
    for k, val in sd.items():
        assert relnorm(val, s["actor." + k].numpy()) < 1e-5, (k, relnorm(val, s["actor." + k].numpy()))
    pol.close()


def test_on_policy_rollout_step_matches_oracle():
    """P1: Memory.normalize -> Actor.act -> process_action, incl. quirk 18 (normalised obs next to a denormalised target)."""
    from oracle import rollout as R
    from spp_rl_b200 import Population, init_state
    ob, ac, E = 17, 6, 300
    rng = np.random.RandomState(5)
    mn, mx = (-rng.rand(ob) * 2 - 0.5).astype(np.float32), (rng.rand(ob) * 2 + 0.5).astype(np.float32)
    obs = (rng.rand(E, ob) * (mx - mn) + mn).astype(np.float32)
    noise = rng.randn(E, ob).astype(np.float32)
    s = {}
    for name, o, i in (("fc1", 64, ob), ("fc2", 64, 64), ("fc3", ob, 64)):
        b = 1 / np.sqrt(i)
        s["actor." + name + ".weight"] = torch.from_numpy(rng.uniform(-b, b, (o, i)).astype(np.float32))
        s["actor." + name + ".bias"] = torch.from_numpy(rng.uniform(-b, b, (o,)).astype(np.float32))
    s["actor.log_scale"] = torch.full((ob,), -1.34)
    acm0 = {k: v for k, v in init_state("sac", ob, ac, 77).items() if k.startswith("acm.")}
    s.update({k: torch.from_numpy(v) for k, v in acm0.items()})
    st = NormStats(True, torch.from_numpy(mn), torch.from_numpy(mx))
    a_ref, lp_ref, acm_ref = R.on_policy_step(s, st, torch.from_numpy(obs), torch.from_numpy(noise), 1.0, torch.ones(ac))
    pol = PpoPolicy(ob, ac, max_rows=1024, max_batch_rows=1024, min_max_denormalize=True)
    pol.set_norm_stats(mn, mx)
    pol.load_state_dict("actor", {k[6:]: v for k, v in s.items() if k.startswith("actor.")})
    action, logp, target = pol.act(obs, noise)
    np.testing.assert_allclose(action, a_ref.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(logp, lp_ref.numpy(), rtol=1e-5, atol=1e-5)
    pop = Population(algo="sac", ob_dim=ob, ac_dim=ac, population=1, min_max_denormalize=True, update_batch_size=64)
    pop.set_norm_stats(mn, mx)
    pop.load_state_dict("acm", {k[4:]: v for k, v in acm0.items()})
    tgt2, acm_action = pop.rollout_step(obs[None], action[None], None, random_phase=2, obs_norm=True, denormalize_actor_out=True)
    np.testing.assert_allclose(tgt2[0], target, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(acm_action[0], acm_ref.numpy(), rtol=1e-5, atol=2e-6)
    pol.close(); pop.close()


def test_acm_ring_add_buffer_joint_quirk():
    """P7: ReplayBufferAcM.add_buffer (rltoolkit/buffer/replay_buffer.py:284-297) through the ring ABI, joint quirk included."""
    from spp_rl_b200 import Population
    from spp_rl_b200.rltoolkit_api import add_rollouts_to_acm_ring
    g = np.load(os.path.join(G, "acm_add_buffer.npz"))
    ob, ac = g["chain"].shape[1], g["acts"].shape[1]
    pop = Population(algo="sac", ob_dim=ob, ac_dim=ac, population=1, buffer_size=40, update_batch_size=8, acm_batch_size=8)
    add_rollouts_to_acm_ring(pop, 0, g["chain"], g["acts"], list(g["joints"]))
    L = int(g["length"])
    assert pop.ring_state(0)[2] == L
    obs, nobs, _, _, _, aacm = pop.ring_sample_batch(0, np.arange(L, dtype=np.int64))
    assert np.array_equal(obs, g["obs"]) and np.array_equal(nobs, g["next_obs"]) and np.array_equal(aacm, g["actions_acm"])
    pop.close()
