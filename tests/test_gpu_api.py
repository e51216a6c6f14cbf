"""-m gpu: the host-side mirror of the reference's classes (spp_rl_b200.rltoolkit_api): constructor kwargs of the train
scripts, update / make_update / pre_train / train / test, the reference's random streams, and the pickle layout."""
import os
import pickle

import numpy as np
import pytest
import torch

from spp_rl_b200.rltoolkit_api import DDPG_AcM, SAC_AcM

pytestmark = pytest.mark.gpu

SCRIPT_KW = dict(       # train/spp_sac_hopper.py:76-112 (sizes shrunk)
    env_name="Hopper-v2", iterations=2, max_frames=None, batch_size=120, test_episodes=1, stats_freq=5, gamma=0.99, actor_lr=1e-3,
    critic_lr=1e-3, alpha_lr=1e-3, alpha=0.2, update_batch_size=64, random_frames=50, update_freq=25, grad_steps=5,
    tensorboard_dir=None, tensorboard_comment="", log_dir=None, use_gpu=False, acm_epochs=1, acm_batch_size=50,
    acm_pre_train_samples=230, acm_pre_train_epochs=2, acm_update_freq=100, acm_lr=1e-3, acm_update_batches=4, custom_loss=0.2,
    denormalize_actor_out=True, min_max_denormalize=True, norm_closs=False, acm_critic=True, buffer_size=5000)


def test_sac_acm_script_kwargs_pretrain_train_test_save_load(tmp_path):
    torch.manual_seed(0); np.random.seed(0)
    m = SAC_AcM(**SCRIPT_KW)
    assert m.tau == 0.005 and m.act_noise == 0.1 and m.max_ep_len is None and m.target_entropy == -3.0     # quirks 1-4
    assert float(m.actor_ac_lim) == 1.0
    m.pre_train()
    assert len(m.replay_buffer) >= 230 and m.min_obs is not None and np.isfinite(m.loss["acm"])
    m.train()
    assert m.stats_logger.frames >= 240 and all(np.isfinite(v) for v in m.loss.values())
    assert set(m.loss) >= {"actor", "critic_1", "critic_2", "sac", "dist", "acm"}
    assert np.isfinite(m.test(1))
    path = os.path.join(tmp_path, "m.pkl")
    m.save(path)
    d = pickle.load(open(path, "rb"))
    assert list(d) == ["actor", "critic_1", "critic_2", "obs_mean", "obs_std", "min_obs", "max_obs", "acm"]
    assert tuple(d["critic_1"]["fc1.weight"].shape) == (256, 14) and d["actor"]["fc_scale.weight"].dtype == torch.float32
    m2 = SAC_AcM(**SCRIPT_KW)
    m2.load(path)
    for net in ("actor", "critic_1", "critic_2", "acm"):
        a, b = getattr(m, net).state_dict(), getattr(m2, net).state_dict()
        assert list(a) == list(b) and all(torch.equal(a[k], b[k]) for k in a)
    assert torch.equal(m2.replay_buffer.max_obs, m.replay_buffer.max_obs)
    with pytest.raises(TypeError):
        SAC_AcM(env_name="Hopper-v2", not_a_kwarg=1)
    m.close(); m2.close()


def test_fused_make_update_follows_the_reference_random_streams():
    """grad_steps x (sample_batch + update) in one launch == the reference's loop of single steps under the same seeds."""
    kw = dict(SCRIPT_KW, grad_steps=3, update_freq=1, random_frames=0, acm_update_freq=10 ** 9)
    models = []
    for _ in range(2):
        torch.manual_seed(5); np.random.seed(5)
        m = SAC_AcM(**kw)
        m.pre_train()
        models.append(m)
    a, b = models
    for net in ("actor", "critic_1", "critic_2", "acm"):
        b._pop.load_state_dict(net, a._pop.state_dict(net))
    b._pop.sync_targets(); a._pop.sync_targets()
    a.stats_logger.frames = b.stats_logger.frames = 7
    torch.manual_seed(11); np.random.seed(11)
    a.make_update()                                    # fused
    torch.manual_seed(11); np.random.seed(11)
    idx = [np.random.randint(0, len(b.replay_buffer), b.update_batch_size) for _ in range(3)]   # the fused path draws these first
    np.random.seed(11)
    eps_backup = torch.get_rng_state()
    for g in range(3):                                 # the reference's loop (ddpg.py:231-237)
        batch = b.replay_buffer.sample_batch(b.update_batch_size)
        b.update(*batch)
    # numpy stream is identical; torch's eps stream differs only in interleaving (fused draws all eps after all indices),
    # which is the same order because indices come from numpy and eps from torch
    for net in ("actor", "critic_1", "critic_2"):
        sa, sb = a._pop.state_dict(net), b._pop.state_dict(net)
        for k in sa:
            assert np.array_equal(sa[k], sb[k]), (net, k)
    assert a.alpha == b.alpha
    a.close(); b.close()


@pytest.mark.parametrize("min_max", [True, False])
def test_obs_norm_ring_updates_gather_normalised_rows(min_max):
    """obs_norm=True (replay_buffer.py:246-248): the fused ring update normalises obs / next_obs inside the gather; it must equal the
    reference's loop, where sample_batch normalises on the host (torch arithmetic) and update() takes the batch as given."""
    kw = dict(SCRIPT_KW, grad_steps=3, update_freq=1, random_frames=0, acm_update_freq=10 ** 9, obs_norm=True, min_max_denormalize=min_max)
    models = []
    for _ in range(2):
        torch.manual_seed(5); np.random.seed(5)
        m = DDPG_AcM(**kw)
        m.pre_train()
        models.append(m)
    a, b = models
    assert a.replay_buffer.obs_norm and (a.replay_buffer.max_obs is not None)
    o = a.replay_buffer.sample_batch(8)[0]
    raw = torch.from_numpy(a._pop.ring_sample_batch(0, np.arange(8))[0])
    assert not torch.equal(a.replay_buffer.normalize(raw), raw)            # the statistics are not the identity
    for net in ("actor", "critic", "acm"):
        b._pop.load_state_dict(net, a._pop.state_dict(net))
    b._pop.sync_targets(); a._pop.sync_targets()
    a.stats_logger.frames = b.stats_logger.frames = 7
    torch.manual_seed(11); np.random.seed(11)
    a.make_update()                                    # fused: device-side normalisation
    torch.manual_seed(11); np.random.seed(11)
    for g in range(3):                                 # the reference's loop (ddpg.py:231-237): host-side normalisation
        b.update(*b.replay_buffer.sample_batch(b.update_batch_size))
    for net in ("actor", "critic"):
        sa, sb = a._pop.state_dict(net), b._pop.state_dict(net)
        for k in sa:
            assert np.array_equal(sa[k], sb[k]), (net, k)
    a.close(); b.close()


def test_ddpg_acm_with_basic_acm_runs():
    torch.manual_seed(1); np.random.seed(1)
    m = DDPG_AcM(acm_model="basic", **dict(SCRIPT_KW, env_name="HalfCheetah-v2", custom_loss=1.0, act_noise=0.05))
    assert m.act_noise == 0.05
    m.pre_train(); m.train()
    assert set(m.acm.state_dict()) == {"t", "t1", "fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias", "fc21.weight", "fc21.bias", "fc3.weight", "fc3.bias"}
    assert all(np.isfinite(v) for v in m.loss.values())
    m.close()


def test_update_acm_epochs_steplr_and_validation_loss_match_reference_fixture():
    """AcMTrainer.update_acm (acm.py:266-303) x3 epochs -- shuffled pass with the partial last minibatch, StepLR(1, 0.5) stepped per
    epoch -- and calculate_validation_loss (acm.py:329-343), against the unmodified reference (tests/golden/acm_epochs.npz)."""
    from spp_rl_b200.init import init_state
    from tests.parity_util import relnorm

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "acm_epochs.npz"))
    ob, ac = 17, 6
    m = DDPG_AcM(env_name="HalfCheetah-v2", acm_pre_train_samples=250, acm_val_buffer_size=90, buffer_size=1000, acm_lr=2e-3,
                 acm_batch_size=64, acm_scheduler_step=1, acm_scheduler_gamma=0.5, update_batch_size=64)
    assert m.acm_val_buffer.size == 99 and m.loss["acm_val"] == 0.0
    s0 = init_state("ddpg", ob, ac, 9, "acm", True)
    m.acm.load_state_dict({k[4:]: torch.from_numpy(v.copy()) for k, v in s0.items() if k.startswith("acm.")})
    chain, acts = g["chain"], g["acts"]
    prev = m.replay_buffer.add_obs(chain[0:1])
    for t in range(len(acts)):
        m.replay_buffer.add_acm_action(acts[t])
        nxt = m.replay_buffer.add_obs(chain[t + 1:t + 2])
        m.replay_buffer.add_timestep(prev, nxt, torch.from_numpy(chain[t + 1:t + 2]), 0.0, False, False)
        prev = nxt
    vchain, vacts = g["vchain"], g["vacts"]
    prev = m.acm_val_buffer.add_obs(vchain[0])
    for t in range(len(vacts)):
        nxt = m.acm_val_buffer.add_obs(vchain[t + 1])
        m.acm_val_buffer.add_timestep(prev, nxt, vacts[t])
        prev = nxt
    assert len(m.acm_val_buffer) == 90 and len(m.replay_buffer) == 250
    perms = [torch.from_numpy(p) for p in g["perms"]]
    orig, calls = torch.randperm, [0]

    def fake(n, *a, **k):
        calls[0] += 1
        return perms[calls[0] - 1]
    torch.randperm = fake
    try:
        for e in range(3):
            m.update_acm(epochs=1)
            assert m.loss["acm"] == pytest.approx(float(g["losses"][e]), rel=1e-5)
            assert m.loss["acm_val"] == pytest.approx(float(g["val_losses"][e]), rel=1e-5)
    finally:
        torch.randperm = orig
    sd = m._pop.state_dict("acm")
    ad, step = m._pop.adam_state("acm")
    assert step == 12                                   # 3 epochs x ceil(250 / 64) minibatches
    for k, v in sd.items():
        lim = 1e-5
        assert relnorm(v, g["acm." + k]) < lim, (k, relnorm(v, g["acm." + k]))
        assert relnorm(ad[k][0], g["acm." + k + "#m"]) < lim, k
        assert relnorm(ad[k][1], g["acm." + k + "#v"]) < lim, k
    m.close()


PPO_KW = dict(      # rltoolkit/acm/on_policy.py:221-244 (sizes shrunk)
    env_name="HalfCheetah-v2", gamma=0.99, acm_pre_train_samples=400, acm_pre_train_epochs=2, iterations=2, batch_size=300, stats_freq=5,
    acm_update_freq=1, acm_epochs=1, acm_lr=1e-4, actor_lr=3e-4, critic_lr=3e-4, kl_div_threshold=0.1, max_ppo_epochs=4,
    ppo_batch_size=128, acm_batch_size=64, denormalize_actor_out=True, min_max_denormalize=True, custom_loss=0.5, tensorboard_dir=None,
    obs_norm=True, test_episodes=1, acm_val_buffer_size=200)


class _ShortEpisodes:
    """The synthetic stand-in with 60-step episodes, so that a 300-row batch holds several rollouts (joints, truncations)."""

    def __init__(self):
        from spp_rl_b200 import envs
        self.e = envs.make("HalfCheetah-v2", seed=3)
        self._max_episode_steps = 60
        self.observation_space, self.action_space = self.e.observation_space, self.e.action_space
        self.t = 0

    def reset(self):
        self.t = 0
        return self.e.reset()

    def step(self, a):
        o, r, d, i = self.e.step(a)
        self.t += 1
        return o, r, bool(d or self.t >= 60), i

    def close(self):
        pass


def test_ppo_acm_script_kwargs_pretrain_train_test_save_load(tmp_path):
    from spp_rl_b200.rltoolkit_ppo import PPO_AcM

    torch.manual_seed(0); np.random.seed(0)
    m = PPO_AcM(env=_ShortEpisodes(), **PPO_KW)
    assert float(m.actor_ac_lim) == 1.0 and m.max_ep_len == 60 and m.buffer_size == 440
    m.pre_train()
    assert len(m.replay_buffer) >= 400 and m.min_obs is not None and np.isfinite(m.loss["acm"]) and m.loss["acm_val"] > 0
    w0 = m.actor.state_dict()["fc1.weight"].clone()
    c0 = m.critic.state_dict()["fc3.weight"].clone()
    a0 = m.acm.state_dict()["fc1.weight"].clone()
    m.train()
    assert m.iteration == 2 and m.stats_logger.frames >= 600 and len(m.buffer) >= 300
    assert set(m.loss) >= {"actor", "critic", "acm", "acm_val", "policy", "dist", "entropy"}
    assert all(np.isfinite(v) for v in m.loss.values()), m.loss
    assert 2 <= m.kl_div_updates_counter <= 2 * 5
    assert not torch.equal(w0, m.actor.state_dict()["fc1.weight"]) and not torch.equal(c0, m.critic.state_dict()["fc3.weight"])
    assert not torch.equal(a0, m.acm.state_dict()["fc1.weight"])
    # memory views (memory.py:146-170): obs / next_obs skip the joints, one action per row
    b = m.buffer
    assert b.obs.shape == b.next_obs.shape == (len(b), 17) and len(b.actions_acm) == len(b)
    assert torch.equal(b.obs[1], b.next_obs[0]) and b.end[-1]
    assert np.isfinite(m.test(1))
    path = os.path.join(tmp_path, "ppo.pkl")
    m.save(path)
    d = pickle.load(open(path, "rb"))
    assert list(d) == ["actor", "critic", "obs_mean", "obs_std", "min_obs", "max_obs", "acm"]
    assert tuple(d["actor"]["fc3.weight"].shape) == (17, 64) and tuple(d["actor"]["log_scale"].shape) == (17,)
    m2 = PPO_AcM(env=_ShortEpisodes(), **PPO_KW)
    m2.load(path)
    for net in ("actor", "critic", "acm"):
        a, b2 = getattr(m, net).state_dict(), getattr(m2, net).state_dict()
        assert list(a) == list(b2) and all(torch.equal(a[k], b2[k]) for k in a)
    m3 = PPO_AcM(env=_ShortEpisodes(), **dict(PPO_KW, custom_loss=0.0))      # P6: falls back to plain PPO.update_actor (on_policy.py:88-98)
    m3.pre_train()
    m3.train()
    assert set(m3.loss) >= {"actor", "entropy", "sum", "critic"} and all(np.isfinite(v) for v in m3.loss.values())
    m.close(); m2.close(); m3.close()


def test_a2c_acm_script_kwargs_pretrain_train_save_load(tmp_path):
    """rltoolkit/acm/test/test_acm_on_policy.py:16-45 (A2C_AcM with and without the validation buffer), on the device path."""
    from spp_rl_b200.rltoolkit_ppo import A2C_AcM

    kw = {k: v for k, v in PPO_KW.items() if k not in ("kl_div_threshold", "max_ppo_epochs", "ppo_batch_size", "epsilon", "gae_lambda", "entropy_coef")}
    torch.manual_seed(0); np.random.seed(0)
    m = A2C_AcM(env=_ShortEpisodes(), **kw)
    m.pre_train()
    w0 = m.actor.state_dict()["fc1.weight"].clone()
    m.train()
    assert m.iteration == 2 and set(m.loss) >= {"actor", "critic", "acm", "policy", "dist"} and all(np.isfinite(v) for v in m.loss.values()), m.loss
    assert not torch.equal(w0, m.actor.state_dict()["fc1.weight"])
    assert m.loss["policy"] == pytest.approx(m.loss["actor"] + m.custom_loss * m.loss["dist"], rel=1e-6)
    path = os.path.join(tmp_path, "a2c.pkl")
    m.save(path)
    assert list(pickle.load(open(path, "rb"))) == ["actor", "critic", "obs_mean", "obs_std", "min_obs", "max_obs", "acm"]
    with pytest.raises(TypeError):
        A2C_AcM(env=_ShortEpisodes(), ppo_batch_size=64, **kw)          # not an A2C keyword
    m0 = A2C_AcM(env=_ShortEpisodes(), **dict(kw, custom_loss=0.0, acm_val_buffer_size=None))      # A2C.update_actor (zeroes its gradients)
    m0.pre_train()
    m0.train()
    assert np.isfinite(m0.loss["actor"])
    m.close(); m0.close()


def test_ppo_acm_iteration_equals_the_kernel_level_calls():
    """perform_iteration's update half == PpoPolicy calls on the same rollout (the fixture-pinned path of test_gpu_ppo.py)."""
    from spp_rl_b200.ppo import PpoPolicy
    from spp_rl_b200.rltoolkit_ppo import PPO_AcM, RolloutMemory

    torch.manual_seed(4); np.random.seed(4)
    m = PPO_AcM(env=_ShortEpisodes(), **PPO_KW)
    m.pre_train()
    m.buffer = RolloutMemory(m.min_obs, m.max_obs, m.obs_mean, m.obs_std, True)
    m.collect_batch(m.buffer)
    b = m.buffer
    N = len(b)
    pol = PpoPolicy(17, 6, max_rows=N, max_batch_rows=128, min_max_denormalize=True, norm_closs=True, gamma=0.99, gae_lambda=0.95,
                    ppo_epsilon=0.2, entropy_coef=0.0, custom_loss=0.5, actor_lr=3e-4, critic_lr=3e-4)
    pol.set_limits(1.0)
    pol.set_norm_stats(m.min_obs.numpy(), m.max_obs.numpy(), m.obs_mean.numpy(), m.obs_std.numpy())
    for net in ("actor", "critic"):
        pol.load_state_dict(net, getattr(m, net).state_dict())
    end = np.asarray(b.end, np.float32)
    stops = np.nonzero(end)[0]; starts = np.concatenate([[0], stops[:-1] + 1])
    pol.load_rollout(b.obs.numpy(), b.next_obs.numpy(), torch.cat(b.actions).numpy(), torch.cat(b.action_logprobs).numpy(),
                     np.asarray(b.rewards, np.float32), np.asarray(b.done, np.float32), end, starts, stops + 1 - starts)
    closs = pol.update_critic(10, 10)
    pol.advantages(); pol.normalize_adv()
    state = torch.get_rng_state()
    from spp_rl_b200.rltoolkit_api import sampler_permutation
    perms = torch.stack([sampler_permutation(N) for _ in range(4)]).numpy()
    torch.set_rng_state(state)
    losses, epochs, _ = pol.update_actor(perms, 128, 0.1, 4)
    adv = m.update_critic(b)
    m.update_actor(adv, b)
    assert m.loss["critic"] == closs and m.loss["policy"] == losses["policy"] and m.kl_div_updates_counter == min(epochs + 1, 4)
    for net in ("actor", "critic"):
        sa, sb = getattr(m, net).state_dict(), pol.state_dict(net)
        assert all(np.array_equal(sa[k].numpy(), sb[k]) for k in sb), net
    assert sum(e for e in b.end) >= 5 and any(e and not d for e, d in zip(b.end, b.done))      # truncated rollouts bootstrap
    pol.close(); m.close()


def test_unbiased_update_equals_the_reference_loop_of_single_steps():
    """DDPG_AcM.make_unbiased_update (ddpg_acm.py:59-73): action := next_obs; fused launch == loop of update() calls.  acm_critic=False
    makes the critic actually consume that action."""
    kw = dict(SCRIPT_KW, env_name="HalfCheetah-v2", grad_steps=3, update_freq=1, random_frames=0, acm_update_freq=10 ** 9, acm_critic=False,
              unbiased_update=True, custom_loss=1.0)
    models = []
    for _ in range(2):
        torch.manual_seed(6); np.random.seed(6)
        m = DDPG_AcM(**kw)
        m.pre_train()
        models.append(m)
    a, b = models
    for net in ("actor", "critic", "acm"):
        b._pop.load_state_dict(net, a._pop.state_dict(net))
    a._pop.sync_targets(); b._pop.sync_targets()
    a.stats_logger.frames = b.stats_logger.frames = 7
    torch.manual_seed(12); np.random.seed(12)
    a.make_update()
    torch.manual_seed(12); np.random.seed(12)
    for g in range(3):
        obs, next_obs, _, reward, done, acm_action = b.replay_buffer.sample_batch(b.update_batch_size)
        b.update(obs=obs, next_obs=next_obs, action=next_obs, reward=reward, done=done, acm_action=acm_action)
    for net in ("actor", "critic"):
        sa, sb = a._pop.state_dict(net), b._pop.state_dict(net)
        for k in sa:
            assert np.array_equal(sa[k], sb[k]), (net, k)
    assert a.loss["critic"] == b.loss["critic"]
    a.close(); b.close()
