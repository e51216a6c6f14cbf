"""-m gpu: the device-resident SPP-PPO rollout (P1: ppo_rollout_kernel, E vectorised synthetic environments x T steps in one launch)
against the oracle's restatement of A2C.collect_batch under injected noise, and the on-policy -> ACM replay ring hand-over (P7:
spp_ring_add_rollout_store) against the per-transition host path that the reference fixture acm_add_buffer.npz pins."""
import numpy as np
import pytest
import torch

from oracle import rollout as R
from oracle.norm import NormStats
from spp_rl_b200 import Population, init_state
from spp_rl_b200.ppo import PpoPolicy
from spp_rl_b200.rltoolkit_api import add_rollouts_to_acm_ring

pytestmark = pytest.mark.gpu


def _setup(ob, ac, acm_kind, rows, ring, seed=5):
    rng = np.random.RandomState(seed)
    mn, mx = (-rng.rand(ob) * 2 - 0.5).astype(np.float32), (rng.rand(ob) * 2 + 0.5).astype(np.float32)
    s = {}
    for name, o, i in (("fc1", 64, ob), ("fc2", 64, 64), ("fc3", ob, 64)):
        b = 1 / np.sqrt(i)
        s["actor." + name + ".weight"] = torch.from_numpy(rng.uniform(-b, b, (o, i)).astype(np.float32))
        s["actor." + name + ".bias"] = torch.from_numpy(rng.uniform(-b, b, (o,)).astype(np.float32))
    s["actor.log_scale"] = torch.full((ob,), -1.34)
    acm0 = {k: v for k, v in init_state("ddpg", ob, ac, 77, acm_kind, True).items() if k.startswith("acm.")}
    s.update({k: torch.from_numpy(v) for k, v in acm0.items()})
    pol = PpoPolicy(ob, ac, max_rows=rows, max_batch_rows=256, min_max_denormalize=True)
    pol.set_norm_stats(mn, mx)
    pol.load_state_dict("actor", {k[6:]: v for k, v in s.items() if k.startswith("actor.")})
    pop = Population(algo="ddpg", ob_dim=ob, ac_dim=ac, population=1, acm_kind=acm_kind, min_max_denormalize=True, update_batch_size=64,
                     buffer_size=ring, store_actions=False)
    pop.set_norm_stats(mn, mx)
    pop.set_limits(np.ones(ob, np.float32), np.full(ac, 0.7, np.float32))
    pop.load_state_dict("acm", {k[4:]: v for k, v in acm0.items()})
    return rng, s, NormStats(True, torch.from_numpy(mn), torch.from_numpy(mx)), pol, pop


@pytest.mark.parametrize("ob,ac,acm_kind,E,T", [(17, 6, "acm", 37, 50), (11, 3, "basic", 70, 24), (3, 1, "acm", 5, 40)])
def test_device_rollout_matches_oracle_collect_batch(ob, ac, acm_kind, E, T):
    rng, s, st, pol, pop = _setup(ob, ac, acm_kind, 2 * E * T, 4096)
    max_ep_len, done_prob = 16, 0.05
    obs0 = ep0 = None
    for call in range(2):      # the second launch continues the environments where the first one left them
        nz = [rng.randn(T, E, ob).astype(np.float32) for _ in range(2)]
        u = rng.rand(T, E).astype(np.float32)
        nr = rng.randn(T, E, ob).astype(np.float32)
        ref = R.on_policy_rollout_synthetic(s, st, E, T, torch.from_numpy(nz[0]), torch.from_numpy(nz[1]), torch.from_numpy(u), torch.from_numpy(nr),
                                            1.0, torch.full((ac,), 0.7), max_ep_len, done_prob, obs0, ep0)
        obs0, ep0 = ref["obs_final"], ref["ep_len_final"]
        pol.rollout_synthetic(pop, E, T, max_ep_len=max_ep_len, done_prob=done_prob, reset_envs=(call == 0), noise_act=nz[0], noise_env=nz[1],
                              u_done=u, noise_reset=nr)
        for k in ("done", "end"):
            assert np.array_equal(pol.store(k), ref[k].numpy()), k      # flags bit-exact
        for k, tol in (("x", 2e-5), ("xn", 2e-5), ("act", 2e-5), ("aacm", 2e-5), ("raw_obs", 2e-5), ("raw_next", 2e-5), ("rew", 2e-5)):
            np.testing.assert_allclose(pol.store(k), ref[k].numpy(), rtol=tol, atol=tol, err_msg=k)
        np.testing.assert_allclose(pol.store("logp"), ref["logp"].numpy(), rtol=2e-5, atol=2e-4)
        assert ref["end"].sum() > E      # episodes did end inside the batch (time limit 16 < T)
    pol.close(); pop.close()


@pytest.mark.parametrize("ring", [5000, 700, 250])
def test_store_to_acm_ring_equals_the_per_transition_add_buffer(ring):
    """P7 on device data: spp_ring_add_rollout_store == ReplayBufferAcM.add_buffer driven transition by transition through the ring
    ABI (add_rollouts_to_acm_ring, pinned by the reference fixture in test_gpu_ppo.py), with and without wrap-around; twice in a row."""
    ob, ac, E, T = 17, 6, 23, 40
    rng, s, st, pol, pop = _setup(ob, ac, "acm", E * T, ring)
    pop2 = Population(algo="ddpg", ob_dim=ob, ac_dim=ac, population=1, acm_kind="acm", min_max_denormalize=True, update_batch_size=64,
                      buffer_size=ring, store_actions=False)
    for call in range(2):
        pol.rollout_synthetic(pop, E, T, max_ep_len=9, done_prob=0.03, seed=3 + call, reset_envs=(call == 0))
        pop.ring_add_rollout_store(0, pol)
        raw, nxt, aacm, end = pol.store("raw_obs"), pol.store("raw_next"), pol.store("aacm"), pol.store("end")
        chain, acts, joints = [], [], []
        for e in range(E):                       # MemoryAcM as collect_batch would have filled it, environment by environment
            for t in range(T):
                row = t * E + e
                chain.append(raw[row]); acts.append(aacm[row])
                if end[row]:
                    chain.append(nxt[row]); joints.append(len(chain))
        add_rollouts_to_acm_ring(pop2, 0, np.stack(chain), np.stack(acts), joints)
        assert pop.ring_state(0) == pop2.ring_state(0)
        L = pop.ring_state(0)[2]
        idx = np.arange(L, dtype=np.int64)
        a, b = pop.ring_sample_batch(0, idx), pop2.ring_sample_batch(0, idx)
        for x, y, name in zip(a, b, ("obs", "next_obs", "act", "rew", "done", "aacm")):
            if name in ("act",):
                continue
            assert np.array_equal(x, y), (call, name)
    pol.close(); pop.close(); pop2.close()


def test_ppo_acm_class_runs_whole_iterations_on_the_device():
    """PPO_AcM(vector_envs=E): perform_iteration with the rollout, the update, add_buffer, the ACM update and the statistics refresh all
    device-resident (the component parities are the tests above and tests/test_gpu_ppo.py; this checks the class-level plumbing)."""
    import warnings

    from spp_rl_b200 import envs
    from spp_rl_b200.rltoolkit_ppo import PPO_AcM

    torch.manual_seed(0); np.random.seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env = envs.make("Walker2d-v2", synthetic=True)
    m = PPO_AcM(env=env, iterations=3, batch_size=2048, vector_envs=64, gamma=0.99, actor_lr=3e-4, critic_lr=3e-4, max_ppo_epochs=4,
                ppo_batch_size=512, custom_loss=0.5, norm_closs=True, min_max_denormalize=True, denormalize_actor_out=True,
                acm_pre_train_samples=1500, acm_pre_train_epochs=2, acm_val_buffer_size=200, acm_update_batches=5, acm_batch_size=64)
    m.pre_train()
    n0 = len(m.replay_buffer)
    w0 = {k: getattr(m, k).state_dict()["fc1.weight"].clone() for k in ("actor", "critic", "acm")}
    m.train()
    assert m.iteration == 3 and m.stats_logger.frames == 3 * 2048
    assert len(m.replay_buffer) > 1000 and n0 <= m.buffer_size
    assert all(np.isfinite(v) for v in m.loss.values()), m.loss
    for k in ("actor", "critic", "acm"):
        assert not torch.equal(w0[k], getattr(m, k).state_dict()["fc1.weight"]), k
    assert m._pol.store("end").sum() >= 64
    m.close()
