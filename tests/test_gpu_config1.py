"""-m gpu: BASELINE config 1 end to end -- SPP-SAC (acm) on Pendulum-v0, hidden 256, batch 256, single seed.  The fixture
tests/golden/config1_pendulum.npz is the REFERENCE's own SAC_AcM(...).pre_train(); .train() for 400 frames on the Pendulum stub
(tests/golden/make_golden.py: config1_fixture); the mirror class, started from the same weights under the same numpy / torch
seeds, must draw the same replay indices and leave the same ring cursors and index arrays (bit-exact), and end with the same
weights, statistics and running return (tolerances stated below)."""
import os
import time

import numpy as np
import pytest
import torch

from tests.parity_util import relnorm

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config1_pendulum.npz")

CONFIG1_KW = dict(env_name="Pendulum-v0", update_batch_size=256, acm_pre_train_samples=400, acm_pre_train_epochs=2, acm_val_buffer_size=None,
                  buffer_size=5000, iterations=2, batch_size=200, grad_steps=10, update_freq=50, random_frames=100, acm_update_freq=100,
                  acm_epochs=1, acm_batch_size=128, custom_loss=0.2, acm_critic=True, norm_closs=False, denormalize_actor_out=True,
                  min_max_denormalize=True, gamma=0.99, stats_freq=1, verbose=0)


def test_sac_acm_pendulum_trajectory_matches_the_reference(monkeypatch):
    from spp_rl_b200.rltoolkit_api import SAC_AcM
    g = np.load(GOLD)
    torch.manual_seed(11); np.random.seed(11)
    m = SAC_AcM(**CONFIG1_KW)
    for net in ("actor", "critic_1", "critic_2", "acm"):
        getattr(m, net).load_state_dict({k[len("init:" + net) + 1:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("init:" + net + ".")})
    m._pop.sync_targets()
    drawn = []
    orig = np.random.randint

    def rec(lo, hi=None, size=None, *a, **k):
        r = orig(lo, hi, size, *a, **k)
        drawn.append(np.concatenate([[hi, np.size(r)], np.ravel(r)]).astype(np.int64))
        return r
    torch.manual_seed(12); np.random.seed(12)
    monkeypatch.setattr(np.random, "randint", rec)
    t0 = time.perf_counter()
    m.pre_train()
    t1 = time.perf_counter()
    m.train()
    t2 = time.perf_counter()
    monkeypatch.setattr(np.random, "randint", orig)
    frames = m.stats_logger.frames
    print("\nconfig 1: mirror %.1f frames/s (pre_train %.2f s, train %.2f s); reference on the fixture's host: %.1f frames/s"
          % (frames / (t2 - t1), t1 - t0, t2 - t1, float(g["frames"]) / float(g["ref_seconds"][1])))
    # ---- bit-exact: frame / rollout counters, every replay index drawn, ring cursors and index arrays
    assert frames == int(g["frames"]) and m.stats_logger.rollouts == int(g["rollouts"])
    assert len(drawn) == int(g["n_draws"]) and np.array_equal(np.concatenate(drawn), g["drawn"])
    obs_cur, ts_cur, cur_len = m._pop.ring_state(0)
    assert (obs_cur, ts_cur, cur_len) == (int(g["obs_idx"]), int(g["ts_idx"]), int(g["current_len"]))
    L = cur_len
    obs, nobs, act, rew, done, aacm = m._pop.ring_sample_batch(0, np.arange(L, dtype=np.int64))
    chain = g["ring_obs"]
    assert np.array_equal(chain[g["ring_obs_idx"]].shape, obs.shape)
    # ---- closed loop in fp32.  The pre-training rows (random actions) are bit-exact.  Afterwards rounding differences of the actor /
    #      ACM outputs feed back through the pendulum dynamics and 80 Adam steps; the fixture carries the reference's OWN sensitivity
    #      (the same run with one initial weight moved by one ulp: sens_*), and the mirror must stay within 4x of it.
    n_pre = CONFIG1_KW["acm_pre_train_samples"]
    assert np.array_equal(obs[:n_pre], chain[g["ring_obs_idx"]][:n_pre]) and np.array_equal(aacm[:n_pre], g["ring_aacm"][:n_pre])
    tol = lambda key: max(1e-5, 4.0 * float(g[key]))
    np.testing.assert_allclose(obs, chain[g["ring_obs_idx"]], rtol=0, atol=tol("sens_obs"))
    np.testing.assert_allclose(nobs, chain[g["ring_next_obs_idx"]], rtol=0, atol=tol("sens_obs"))
    np.testing.assert_allclose(aacm, g["ring_aacm"], rtol=0, atol=tol("sens_aacm"))
    np.testing.assert_allclose(act, g["ring_act"], rtol=0, atol=tol("sens_act"))
    np.testing.assert_allclose(rew, g["ring_rew"], rtol=1e-4, atol=tol("sens_obs") * 20)
    assert np.array_equal(done.astype(bool), g["ring_done"].astype(bool))
    first = slice(n_pre, n_pre + 300)      # the first 300 training frames: before the amplification sets in, 1e-5 absolute
    np.testing.assert_allclose(obs[first], chain[g["ring_obs_idx"]][first], rtol=0, atol=1e-5)
    np.testing.assert_allclose(aacm[first], g["ring_aacm"][first], rtol=0, atol=1e-5)
    # ---- statistics (whole-buffer percentiles / moments), temperature, running return
    np.testing.assert_allclose(m.min_obs.numpy(), g["min_obs"], rtol=0, atol=tol("sens_obs"))
    np.testing.assert_allclose(m.max_obs.numpy(), g["max_obs"], rtol=0, atol=tol("sens_obs"))
    np.testing.assert_allclose(m.obs_mean.numpy(), g["obs_mean"], rtol=0, atol=tol("sens_obs"))
    assert m.alpha == pytest.approx(float(g["alpha"]), rel=1e-5)
    assert float(m.stats_logger.running_return) == pytest.approx(float(g["running_return"]), rel=1e-5)
    # ---- weights after 80 SAC updates and 2 + 2 ACM epochs: norm-relative <= max(1e-5, 4 x the reference's own 1-ulp sensitivity)
    #      per tensor (1e-4 floor for tensors of <= 16 elements)
    worst = 0.0
    for k in g.files:
        if not k.startswith("final:"):
            continue
        net, name = k[len("final:"):].split(".", 1)
        v = getattr(m, net).state_dict()[name].numpy()
        e = relnorm(v, g[k])
        bound = max(1e-4 if v.size <= 16 else 1e-5, 4.0 * float(g["sens:" + net + "." + name]))
        worst = max(worst, e)
        assert e < bound, (k, e, bound)
    print("config 1: worst norm-relative weight difference vs the reference %.2e" % worst)
    m.close()
