"""CPU: the oracle against golden fixtures produced by the unmodified reference (tests/golden/make_golden.py).

Pins every hot-path function the reference has no known-answer test for: SAC_AcM / DDPG_AcM update
(incl. Adam moments, Polyak targets, fp64 temperature), the replay-ring state machine and gather,
ReplayBufferAcM.add_buffer, the ACM regression step, and the PPO critic fit / GAE / actor epochs."""
import os

import numpy as np
import pytest
import torch

from oracle import nets, offpolicy as op, ppo as P
from oracle.norm import NormStats, denormalize, normalize
from oracle.ring import Ring, add_rollouts_to_acm_ring
from spp_rl_b200.init import init_state
from tests.parity_util import make_batches, make_stats, oracle_state, relnorm

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-5      # north_star bar; measured oracle-vs-reference: <= 9e-7 on moments and all but one weight tensor,
                # 5.8e-6 on actor.fc_scale.weight (Adam amplifies summation-order noise where |g| is tiny)


def _load(name):
    return np.load(os.path.join(G, name), allow_pickle=False)


def _check_state(s, gold, keys_prefixes):
    worst = 0.0
    for k in gold.files:
        if any(k.startswith(p) for p in keys_prefixes) and k in s:
            v = s[k]
            v = v.numpy() if hasattr(v, "numpy") else np.asarray(v)
            if k.endswith("#step"):
                assert int(v) == int(gold[k]), k
                continue
            e = relnorm(v, gold[k])
            if np.size(v) <= 16:     # scalars / tiny vectors (fc3.bias and its moments) are sums with heavy cancellation
                assert e < 1e-4, (k, e)
            else:
                worst = max(worst, e)
    return worst


def test_sac_update_matches_reference():
    gold = _load("sac_hopper_g3.npz")
    ob, ac, B, Gs, seed = [int(x) for x in gold["meta"]]
    mn, mx, mean, std = make_stats(ob, seed, True)
    obs, nobs, act, rew, done, aacm, eps = make_batches(ob, ac, 1, Gs, B, seed, mn, mx)
    st = NormStats(True, torch.from_numpy(mn), torch.from_numpy(mx), torch.from_numpy(mean), torch.from_numpy(std))
    hp = op.OffPolicyHP(gamma=0.99, custom_loss=0.2, norm_closs=False, acm_critic=True, target_entropy=-float(ac),
                        actor_lim=torch.ones(ob), acm_lim=torch.ones(ac))
    s = oracle_state(init_state("sac", ob, ac, seed * 100), "sac")
    alpha = None
    for g in range(Gs):
        t = lambda x: torch.from_numpy(x[0, g])
        l, alpha = op.sac_acm_update(s, hp, st, t(obs), t(nobs), t(act), t(rew), t(done), t(aacm),
                                     torch.from_numpy(eps[0, g, 0]), torch.from_numpy(eps[0, g, 1]), alpha)
        ref = gold["losses"][g]
        for v, r in zip([l["critic_1"], l["critic_2"], l["actor"], l["sac"], l["dist"], alpha], ref):
            assert v == pytest.approx(r, rel=2e-6, abs=1e-7)
    assert float(s["log_alpha"]) == pytest.approx(float(gold["log_alpha"]), rel=1e-9)
    assert _check_state(s, gold, ["actor", "critic_"]) < TOL


def test_ddpg_update_matches_reference():
    gold = _load("ddpg_hcheetah_g2.npz")
    ob, ac, B, Gs, seed = [int(x) for x in gold["meta"]]
    mn, mx, mean, std = make_stats(ob, seed, True)
    obs, nobs, act, rew, done, aacm, _ = make_batches(ob, ac, 1, Gs, B, seed, mn, mx)
    st = NormStats(True, torch.from_numpy(mn), torch.from_numpy(mx), torch.from_numpy(mean), torch.from_numpy(std))
    hp = op.OffPolicyHP(gamma=0.95, actor_lr=5e-4, critic_lr=5e-4, custom_loss=1.0, norm_closs=False, acm_critic=True,
                        actor_lim=torch.ones(ob), acm_lim=torch.ones(ac))
    s0 = init_state("ddpg", ob, ac, seed * 100, "basic", True)
    s0["acm.t"][:] = 0.7
    s0["acm.t1"][:] = np.linspace(0.5, 1.5, ac)
    s = oracle_state(s0, "ddpg")
    for g in range(Gs):
        t = lambda x: torch.from_numpy(x[0, g])
        l = op.ddpg_acm_update(s, hp, st, t(obs), t(nobs), t(act), t(rew), t(done), t(aacm))
        for v, r in zip([l["critic"], l["actor"], l["ddpg"], l["dist"]], gold["losses"][g]):
            assert v == pytest.approx(r, rel=2e-6, abs=1e-7)
    assert _check_state(s, gold, ["actor", "critic"]) < TOL


def test_ring_state_machine_and_gather_bit_exact():
    g = _load("ring_ops.npz")
    size, ob, ac = int(g["size"]), int(g["ob"]), int(g["ac"])
    ring = Ring(size, ob, ob, ac)
    for kind, f, iv, stt in zip(g["kinds"], g["fvals"], g["ivals"], g["states"]):
        if kind == 0:
            assert ring.add_obs(f[:ob]) == iv[0]
        else:
            ring.add_acm_action(f[2 * ob:2 * ob + ac])
            assert ring.add_obs(f[:ob]) == iv[1]
            ring.add_timestep(iv[0], iv[1], f[ob:2 * ob], f[-1], bool(iv[2]), bool(iv[3]))
        assert (ring.obs_cur, ring.ts_cur, ring.current_len) == tuple(int(x) for x in stt)
    L = ring.current_len
    assert np.array_equal(ring.obs_idx[:L], g["obs_idx"]) and np.array_equal(ring.next_obs_idx[:L], g["next_obs_idx"])
    out = ring.gather(g["idx"])
    for mine, ref in zip(out, (g["s_obs"], g["s_next"], g["s_act"], g["s_rew"], g["s_done"], g["s_aacm"])):
        assert np.array_equal(mine, ref)
    assert out[4].dtype == np.int8


def test_acm_add_buffer_joint_quirk():
    g = _load("acm_add_buffer.npz")
    ring = Ring(40, g["chain"].shape[1], 1, g["acts"].shape[1])
    add_rollouts_to_acm_ring(ring, g["chain"], g["acts"], g["joints"])
    L = int(g["length"])
    assert ring.current_len == L and L < len(g["acts"])          # transitions are dropped at joints
    assert np.array_equal(ring.obs[ring.obs_idx[:L]], g["obs"])
    assert np.array_equal(ring.obs[ring.next_obs_idx[:L]], g["next_obs"])
    assert np.array_equal(ring.actions_acm[:L], g["actions_acm"])


@pytest.mark.parametrize("kind", ["acm", "basic"])
def test_acm_regression_matches_reference(kind):
    g = _load("acm_regress.npz")
    ob, ac = 17, 6
    s0 = init_state("ddpg", ob, ac, 7, kind, True)
    if kind == "basic":
        s0["acm.t"][:] = 0.7
        s0["acm.t1"][:] = np.linspace(0.5, 1.5, ac)
    s = oracle_state(s0, "ddpg")
    rng = np.random.RandomState(11)
    for i in range(3):
        x = rng.randn(100, 2 * ob).astype(np.float32); y = np.tanh(rng.randn(100, ac)).astype(np.float32)
        l = op.acm_batch_update(s, torch.from_numpy(x), torch.from_numpy(y), torch.ones(ac), 1e-3)
        assert l == pytest.approx(float(g[kind + ":losses"][i]), rel=2e-6)
    for k in g.files:
        if k.startswith(kind + ":acm.") and "#" not in k:
            assert relnorm(s[k[len(kind) + 1:]].numpy(), g[k]) < 5e-6, k


def test_acm_epochs_and_validation_loss_match_reference():
    """update_acm x3 (shuffle, partial last minibatch, StepLR per epoch) + calculate_validation_loss, reference fixture."""
    g = _load("acm_epochs.npz")
    ob, ac = 17, 6
    s = oracle_state(init_state("ddpg", ob, ac, 9, "acm", True), "ddpg")
    chain, acts = torch.from_numpy(g["chain"]), torch.from_numpy(g["acts"])
    vchain, vacts = torch.from_numpy(g["vchain"]), torch.from_numpy(g["vacts"])
    for e in range(3):
        l = op.acm_update_epochs(s, chain[:-1], chain[1:], acts, [g["perms"][e]], 64, torch.ones(ac), 2e-3, 1, 0.5, epoch0=e)
        assert l[0] == pytest.approx(float(g["losses"][e]), rel=5e-6)
        v = op.acm_validation_loss(s, vchain[:-1], vchain[1:], vacts, torch.ones(ac))
        assert v == pytest.approx(float(g["val_losses"][e]), rel=5e-6)
    assert list(g["lrs_after"]) == [1e-3, 5e-4, 2.5e-4]
    for k in g.files:
        if k.startswith("acm.") and "#" not in k:
            assert relnorm(s[k].numpy(), g[k]) < 5e-6, k


def test_ppo_pieces_match_reference():
    g = _load("ppo_walker.npz")
    hp = g["hp"]
    gamma, lam, eps_clip, kl_thr, max_ep, bs, a_lr, c_lr, ent, closs, ntu, nupt = [float(x) for x in hp]
    st = NormStats(True, torch.from_numpy(g["min_obs"]), torch.from_numpy(g["max_obs"]), torch.from_numpy(g["obs_mean"]),
                   torch.from_numpy(g["obs_std"]))
    chain = torch.from_numpy(g["chain"])
    oi, ni = P.chain_views(len(chain), list(g["joints"]))
    nobs_, nnobs_ = normalize(st, chain[oi], True), normalize(st, chain[ni], True)
    s = {k[4:]: torch.from_numpy(g[k].copy()) for k in g.files if k.startswith("pre:")}
    rew, done, end = (torch.from_numpy(g[k]) for k in ("rewards", "done", "end"))
    loss = P.update_critic(s, nobs_, nnobs_, rew, done, gamma, c_lr, int(ntu), int(nupt))
    assert loss == pytest.approx(float(g["critic_loss"]), rel=1e-5)
    for k in g.files:
        if k.startswith("fit:"):
            assert relnorm(s[k[4:]].numpy(), g[k]) < 2e-6, k
    adv, _ = P.advantages(s, nobs_, nnobs_, rew, done, end, gamma, lam)
    assert np.abs(adv.numpy() - g["adv"]).max() < 1e-5 * max(1.0, np.abs(g["adv"]).max())
    advn = P.normalize_adv(torch.from_numpy(g["adv"]))
    acts, old = torch.from_numpy(g["actions"]), torch.from_numpy(g["logp"])
    perms = [torch.from_numpy(p) for p in g["perms"]]
    tot, epochs, _ = P.update_actor_acm(s, nobs_, denormalize(st, acts), denormalize(st, nnobs_), old, advn, perms,
                                        float(g["actor_lim"]), a_lr, eps_clip, kl_thr, int(max_ep), int(bs), ent, closs)
    ref = g["actor_losses"]
    for v, r in zip([tot["actor"], tot["entropy"], tot["policy"], tot["dist"]], ref):
        assert v == pytest.approx(float(r), rel=1e-5)
    for k in g.files:
        if k.startswith("post:"):
            assert relnorm(s[k[5:]].numpy(), g[k]) < 2e-6, k


# ---- frame-loop body (A11 / P1): the oracle against the reference's noise_action / initial_act / process_action / Actor.act
ROLLOUT_CASES = [("sac_hopper", "sac", "acm"), ("sac_pendulum", "sac", "acm"), ("ddpg_hcheetah", "ddpg", "basic")]


def rollout_case_inputs(g, name, algo, kind):
    """state, NormStats, limits and flags of one case of tests/golden/rollout_steps.npz (shared with the -m gpu tests)."""
    ob, ac, wseed, sseed, min_max, denorm = [int(x) for x in g[name + ":meta"]]
    s0 = init_state(algo, ob, ac, wseed, kind, False)
    if kind == "basic":
        s0["acm.t"][:] = 0.7
        s0["acm.t1"][:] = np.linspace(0.5, 1.5, ac)
    mn, mx, mean, std = make_stats(ob, sseed, bool(min_max))
    st = NormStats(bool(min_max), torch.from_numpy(mn), torch.from_numpy(mx), torch.from_numpy(mean), torch.from_numpy(std), obs_norm=False)
    return s0, st, (mn, mx, mean, std), g[name + ":actor_lim"], g[name + ":acm_lim"], float(g[name + ":act_noise"]), bool(denorm)


@pytest.mark.parametrize("name,algo,kind", ROLLOUT_CASES)
def test_rollout_step_oracle_matches_reference(name, algo, kind):
    from oracle import rollout as R
    g = _load("rollout_steps.npz")
    s0, st, _, lim, alim, act_noise, denorm = rollout_case_inputs(g, name, algo, kind)
    s = oracle_state(s0, algo)
    t = torch.from_numpy
    obs = t(g[name + ":obs"])
    eps = t(g[name + ":eps"]) if algo == "sac" else None
    tgt, acm = R.off_policy_step(s, st, obs, t(g[name + ":noise"]), t(lim), t(alim), act_noise, algo=algo, eps=eps,
                                 random_phase=False, denormalize_actor_out=denorm)
    np.testing.assert_allclose(tgt.numpy(), g[name + ":target"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(acm.numpy(), g[name + ":acm"], rtol=1e-5, atol=1e-6)
    tgt, acm = R.off_policy_step(s, st, obs, t(g[name + ":init_noise"]), t(lim), t(alim), act_noise, algo=algo, eps=eps,
                                 random_phase=True, denormalize_actor_out=denorm)
    np.testing.assert_allclose(tgt.numpy(), g[name + ":init_target"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(acm.numpy(), g[name + ":init_acm"], rtol=1e-5, atol=1e-6)
    # test(): deterministic actor, no exploration noise (ddpg.py:385-410)
    zero = torch.zeros_like(obs)
    tgt, acm = R.off_policy_step(s, st, obs, zero, t(lim), t(alim), 0.0, algo=algo, eps=zero if algo == "sac" else None,
                                 random_phase=False, denormalize_actor_out=denorm)
    np.testing.assert_allclose(tgt.numpy(), g[name + ":det_target"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(acm.numpy(), g[name + ":det_acm"], rtol=1e-5, atol=1e-6)


def test_on_policy_step_oracle_matches_reference():
    from oracle import rollout as R
    g = _load("rollout_steps.npz")
    t = torch.from_numpy
    s = {k[len("ppo_walker:"):]: t(g[k]) for k in g.files if k.startswith("ppo_walker:actor.") or k.startswith("ppo_walker:acm.")}
    st = NormStats(True, t(g["ppo_walker:min_obs"]), t(g["ppo_walker:max_obs"]))
    a, lp, acm = R.on_policy_step(s, st, t(g["ppo_walker:obs"]), t(g["ppo_walker:noise"]), float(g["ppo_walker:actor_lim"]),
                                  t(g["ppo_walker:acm_lim"]))
    np.testing.assert_allclose(a.numpy(), g["ppo_walker:action"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(lp.numpy(), g["ppo_walker:logp"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(acm.numpy(), g["ppo_walker:acm"], rtol=1e-5, atol=1e-6)


def test_plain_ppo_actor_epochs_match_reference():
    """P6: PPO_AcM(custom_loss=0).update_actor -> PPO.update_actor (ppo.py:152-192) on ppo_walker's rollout."""
    g, gp = _load("ppo_walker.npz"), _load("ppo_plain.npz")
    gamma, lam, eps_clip, kl_thr, max_ep, bs, a_lr, c_lr, ent, closs, ntu, nupt = [float(x) for x in g["hp"]]
    st = NormStats(True, torch.from_numpy(g["min_obs"]), torch.from_numpy(g["max_obs"]))
    chain = torch.from_numpy(g["chain"])
    oi, _ = P.chain_views(len(chain), list(g["joints"]))
    s = {k[4:]: torch.from_numpy(g[k].copy()) for k in g.files if k.startswith("pre:actor.")}
    advn = P.normalize_adv(torch.from_numpy(g["adv"]))
    tot, epochs, _ = P.update_actor_acm(s, normalize(st, chain[oi], True), torch.from_numpy(g["actions"]), None, torch.from_numpy(g["logp"]),
                                        advn, [torch.from_numpy(p) for p in g["perms"]], float(g["actor_lim"]), a_lr, eps_clip, kl_thr,
                                        int(max_ep), int(bs), ent, 0.0, plain=True)
    assert min(epochs + 1, int(max_ep)) == int(gp["epochs_counter"])
    for v, r in zip([tot["actor"], tot["entropy"], tot["policy"]], gp["losses"]):
        assert v == pytest.approx(float(r), rel=1e-5)
    for k in gp.files:
        if k.startswith("post:"):
            assert relnorm(s[k[5:]].numpy(), gp[k]) < 2e-6, k


def a2c_case(g):
    """Inputs of tests/golden/a2c_acm.npz (the rollout is ppo_walker's; shared with the -m gpu test)."""
    gamma, lam, eps_clip, kl_thr, max_ep, bs, a_lr, c_lr, ent, closs, ntu, nupt = [float(x) for x in g["hp"]]
    st = NormStats(True, torch.from_numpy(g["min_obs"]), torch.from_numpy(g["max_obs"]), torch.from_numpy(g["obs_mean"]),
                   torch.from_numpy(g["obs_std"]))
    chain = torch.from_numpy(g["chain"])
    oi, ni = P.chain_views(len(chain), list(g["joints"]))
    return dict(gamma=gamma, a_lr=a_lr, c_lr=c_lr, ntu=int(ntu), nupt=int(nupt), st=st, chain=chain, oi=oi, ni=ni,
                nobs=normalize(st, chain[oi], True), nnobs=normalize(st, chain[ni], True))


def test_a2c_acm_two_iterations_match_reference():
    """(f)4: A2C_AcM = A2C.update_critic -> q - V advantages -> update_actor_acm (on_policy.py:100-124) whose gradients are never
    zeroed: the second iteration steps on the sum of both iterations' gradients."""
    g, ga = _load("ppo_walker.npz"), _load("a2c_acm.npz")
    c = a2c_case(g)
    s = {k[4:]: torch.from_numpy(g[k].copy()) for k in g.files if k.startswith("pre:")}
    rew, done = torch.from_numpy(g["rewards"]), torch.from_numpy(g["done"])
    acts = torch.from_numpy(g["actions"])
    for it in (1, 2):
        loss = P.update_critic(s, c["nobs"], c["nnobs"], rew, done, c["gamma"], c["c_lr"], c["ntu"], c["nupt"])
        assert loss == pytest.approx(float(ga["critic_loss%d" % it]), rel=2e-5)
        adv = P.a2c_advantages(s, c["nobs"], c["nnobs"], rew, done, c["gamma"])
        assert np.abs(adv.numpy() - ga["adv%d" % it]).max() < 2e-5 * max(1.0, np.abs(ga["adv%d" % it]).max())
        losses = P.a2c_actor_step(s, c["nobs"], acts, torch.from_numpy(ga["logp%d" % it]), torch.from_numpy(ga["adv%d" % it]),
                                  float(g["actor_lim"]), c["a_lr"], accumulate=True, custom_loss=0.1, loss_actions=denormalize(c["st"], acts),
                                  next_obs=c["chain"][c["ni"]])
        for v, r in zip([losses["actor"], losses["dist"], losses["policy"]], ga["losses%d" % it]):
            assert v == pytest.approx(float(r), rel=1e-5)
        for k in ga.files:
            if k.startswith("post%d:actor." % it):
                assert relnorm(s[k[6:]].numpy(), ga[k]) < 5e-6, k
    # without the accumulation the second step would be a different one
    s2 = {k[4:]: torch.from_numpy(g[k].copy()) for k in g.files if k.startswith("pre:")}
    for it in (1, 2):
        P.a2c_actor_step(s2, c["nobs"], acts, torch.from_numpy(ga["logp%d" % it]), torch.from_numpy(ga["adv%d" % it]), float(g["actor_lim"]),
                         c["a_lr"], accumulate=False)
    assert relnorm(s2["actor.fc2.weight"].numpy(), ga["post2:actor.fc2.weight"]) > 1e-4
