"""-m gpu: the vectorised rollout step against the oracle restatement of the reference's frame loop body
(noise_action / initial_act + process_action), and the device-resident synthetic rollout's ring writes."""
import numpy as np
import pytest
import torch

from oracle import rollout as R
from oracle.norm import NormStats
from spp_rl_b200 import Population, init_state
from tests.parity_util import make_stats, oracle_state, upload_state

pytestmark = pytest.mark.gpu

CASES = [
    dict(algo="sac", ob=11, ac=3, E=300, min_max=True, random_phase=False, obs_norm=False),
    dict(algo="sac", ob=11, ac=3, E=64, min_max=True, random_phase=True, obs_norm=False),
    dict(algo="sac", ob=17, ac=6, E=130, min_max=False, random_phase=False, obs_norm=True),
    dict(algo="ddpg", ob=17, ac=6, E=257, min_max=True, random_phase=False, obs_norm=False, acm_kind="basic"),
    dict(algo="ddpg", ob=111, ac=8, E=40, min_max=True, random_phase=False, obs_norm=True),
    dict(algo="sac", ob=3, ac=1, E=5, min_max=False, random_phase=False, obs_norm=False, lim=[1.0, 1.0, 8.0], acm_lim=[2.0]),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join("%s%s" % (k, v) for k, v in c.items() if k not in ("lim", "acm_lim")))
def test_rollout_step_matches_oracle(case):
    algo, ob, ac, E = case["algo"], case["ob"], case["ac"], case["E"]
    kind = case.get("acm_kind", "acm")
    P = 2
    mn, mx, mean, std = make_stats(ob, 5, case["min_max"])
    rng = np.random.RandomState(77)
    obs = (rng.rand(P, E, ob) * (mx - mn) + mn).astype(np.float32)
    noise = rng.randn(P, E, ob).astype(np.float32)
    eps = rng.randn(P, E, ob).astype(np.float32)
    lim = np.broadcast_to(np.asarray(case.get("lim", 1.0), np.float32), (ob,)).copy()
    alim = np.broadcast_to(np.asarray(case.get("acm_lim", 1.0), np.float32), (ac,)).copy()
    pop = Population(algo=algo, ob_dim=ob, ac_dim=ac, population=P, acm_kind=kind, min_max_denormalize=case["min_max"],
                     update_batch_size=64)
    pop.set_limits(lim, alim)
    pop.set_norm_stats(mn, mx, mean, std)
    st = NormStats(case["min_max"], torch.from_numpy(mn), torch.from_numpy(mx), torch.from_numpy(mean), torch.from_numpy(std),
                   obs_norm=case["obs_norm"])
    states = []
    for a in range(P):
        s0 = init_state(algo, ob, ac, 300 + a, kind, True)
        if kind == "basic":
            s0["acm.t"][:] = 0.7
            s0["acm.t1"][:] = np.linspace(0.5, 1.5, ac)
        upload_state(pop, s0, a, algo)
        states.append(oracle_state(s0, algo))
    tgt, act = pop.rollout_step(obs, noise, eps if algo == "sac" else None, random_phase=case["random_phase"], act_noise=0.1,
                                obs_norm=case["obs_norm"], denormalize_actor_out=True)
    for a in range(P):
        t_ref, a_ref = R.off_policy_step(states[a], st, torch.from_numpy(obs[a]), torch.from_numpy(noise[a]), torch.from_numpy(lim),
                                         torch.from_numpy(alim), 0.1, algo=algo, eps=torch.from_numpy(eps[a]),
                                         random_phase=case["random_phase"], denormalize_actor_out=True)
        np.testing.assert_allclose(tgt[a], t_ref.numpy(), rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(act[a], a_ref.numpy(), rtol=1e-5, atol=2e-6)
    pop.close()


def test_synthetic_rollout_fills_ring_consistently():
    ob, ac, P, E, steps, S = 11, 3, 3, 128, 5, 4096
    pop = Population(algo="sac", ob_dim=ob, ac_dim=ac, population=P, update_batch_size=64, buffer_size=S, acm_critic=True,
                     custom_loss=0.2)
    pop.set_norm_stats(-np.ones(ob, np.float32), np.ones(ob, np.float32))
    for a in range(P):
        upload_state(pop, init_state("sac", ob, ac, 9 + a), a, "sac")
    pop.rollout_synthetic(E, steps, seed=3)
    pop.sync()
    for a in range(P):
        assert pop.ring_state(a) == (steps * E, steps * E, steps * E)
        idx = np.arange(steps * E, dtype=np.int64)
        obs, nobs, _, rew, done, aacm = pop.ring_sample_batch(a, idx)
        assert np.isfinite(obs).all() and np.isfinite(nobs).all() and np.isfinite(aacm).all()
        assert np.abs(aacm).max() <= 1.0 + 1e-6                      # tanh-bounded ACM action
        assert np.array_equal(nobs[:-E], obs[E:])                    # env e: next_obs at step t is obs at step t+1
        assert np.array_equal(rew, nobs[:, 0])                       # synthetic reward = first coordinate of the new obs
    pop.rollout_synthetic(E, steps, seed=4)                          # a second launch continues the chain
    pop.sync()
    obs2, nobs2, *_ = pop.ring_sample_batch(0, np.arange(2 * steps * E, dtype=np.int64))
    assert np.array_equal(nobs2[:-E], obs2[E:])
    losses = pop.update_ring(2)                                      # and the update kernel can train from it
    assert np.isfinite(losses[:, :, :3]).all()
    pop.close()
