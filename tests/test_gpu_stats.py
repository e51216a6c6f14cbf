"""-m gpu: MetaReplayBuffer.update_obs_mean_std on the device against numpy on the ring's contents:
percentiles bit-exact (the kernels return exact order statistics, the lerp is numpy's), mean / population std to fp64 round-off."""
import numpy as np
import pytest

from spp_rl_b200 import Population

pytestmark = pytest.mark.gpu


def _fill_by_adds(pop, agent, rng, n_steps, ep_len, ob, ac):
    """episodes through the ring state machine (add_obs / add_timestep), so obs rows and obs_idx are not the identity"""
    t = 0
    while t < n_steps:
        oi = pop.ring_add_obs(agent, rng.randn(ob).astype(np.float32) * (1 + agent))
        for s in range(ep_len):
            pop.ring_add_acm_action(agent, rng.randn(ac).astype(np.float32))
            ni = pop.ring_add_obs(agent, (rng.randn(ob) * (1 + 0.1 * s)).astype(np.float32))
            pop.ring_add_timestep(agent, oi, ni, rng.randn(ob).astype(np.float32), float(rng.randn()), s == ep_len - 1, s == ep_len - 1)
            oi = ni
            t += 1


@pytest.mark.parametrize("ob,ac,P,S,n", [(11, 3, 3, 4000, 2500), (17, 6, 2, 3000, 2999), (111, 8, 1, 2000, 1200), (3, 1, 2, 600, 101)])
def test_ring_obs_stats_match_numpy(ob, ac, P, S, n):
    pop = Population(algo="sac", ob_dim=ob, ac_dim=ac, population=P, buffer_size=S, update_batch_size=16)
    rng = np.random.RandomState(ob + P)
    for a in range(P):
        _fill_by_adds(pop, a, rng, n + 37 * a, 50 + 7 * a, ob, ac)
    st = pop.ring_obs_stats()
    for a in range(P):
        L = pop.ring_state(a)[2]
        obs = pop.ring_sample_batch(a, np.arange(L, dtype=np.int64))[0].astype(np.float64)      # self.obs of the reference
        assert np.array_equal(st["p1"][a], np.percentile(obs, 1, axis=0)), (a, "p1")
        assert np.array_equal(st["p99"][a], np.percentile(obs, 99, axis=0)), (a, "p99")
        np.testing.assert_allclose(st["mean"][a], obs.mean(axis=0), rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(st["std"][a], obs.std(axis=0), rtol=1e-12, atol=1e-14)
    pop.close()


def test_ring_obs_stats_full_size_synthetic_ring():
    """1 M-capacity rings (the bench's prefill): percentiles bit-exact against numpy for a sampled agent, and they bracket 98 % of the data"""
    P, ob, ac, S = 4, 11, 3, 1_000_000
    pop = Population(algo="sac", ob_dim=ob, ac_dim=ac, population=P, buffer_size=S, update_batch_size=16, store_actions=False)
    pop.set_norm_stats(-np.ones(ob, np.float32), np.ones(ob, np.float32))
    pop.ring_fill_synthetic(seed=3, n=S * 999 // 1000, episode_len=1000)
    st, nbytes = pop.ring_obs_stats(return_bytes=True)
    a = 2
    L = pop.ring_state(a)[2]
    obs = pop.ring_sample_batch(a, np.arange(L, dtype=np.int64))[0].astype(np.float64)
    assert np.array_equal(st["p1"][a], np.percentile(obs, 1, axis=0))
    assert np.array_equal(st["p99"][a], np.percentile(obs, 99, axis=0))
    np.testing.assert_allclose(st["mean"][a], obs.mean(axis=0), rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(st["std"][a], obs.std(axis=0), rtol=1e-11, atol=1e-13)
    inside = ((obs >= st["p1"][a]) & (obs <= st["p99"][a])).mean(axis=0)
    assert np.all(np.abs(inside - 0.98) < 1e-3)
    assert nbytes == pytest.approx(6.0 * P * L * (ob * 4 + 4))
    pop.close()
