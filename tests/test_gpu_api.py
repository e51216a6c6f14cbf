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


def test_ddpg_acm_with_basic_acm_runs():
    torch.manual_seed(1); np.random.seed(1)
    m = DDPG_AcM(acm_model="basic", **dict(SCRIPT_KW, env_name="HalfCheetah-v2", custom_loss=1.0, act_noise=0.05))
    assert m.act_noise == 0.05
    m.pre_train(); m.train()
    assert set(m.acm.state_dict()) == {"t", "t1", "fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias", "fc21.weight", "fc21.bias", "fc3.weight", "fc3.bias"}
    assert all(np.isfinite(v) for v in m.loss.values())
    m.close()
