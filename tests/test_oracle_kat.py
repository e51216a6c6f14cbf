"""CPU: the oracle against the reference's own known-answer tests (values restated with file:line).

These pin the oracle on every hot-path function for which the reference ships a golden vector:
  rltoolkit/algorithms/ppo/test/test_ppo.py:34-76   _clip_loss (5 cases, rel 1e-4)
  rltoolkit/algorithms/ppo/test/test_ppo.py:79-134  q-values and GAE on the 12-step / 4-rollout Memory fixture
  rltoolkit/algorithms/a2c/test/test_a2c.py:44-56   q-values with a constant critic
  rltoolkit/buffer/test/test_memory.py:43-92        Memory.obs / next_obs joint skipping, mean/std, normalize + clip
  rltoolkit/test/test_utils.py:16-43                kl_divergence, standardize_and_clip
"""
import numpy as np
import pytest
import torch

from oracle import ppo as P
from oracle.norm import NormStats, normalize

CLIP_CASES = [   # test_ppo.py:34-65
    ([-2.3, -5, -1.4, -1.5], [-2.3, -5, -1.4, -1.5], [1, 2.0, 3.0, 4.0], -2.5),
    ([-1.0], [-1.0], [-1.0], 1),
    ([-1.0], [-2.0], [-1.0], 0.8),
    ([-2.0], [-1.0], [1.0], -1.2),
    ([-1.0], [-2.0], [1.0], -0.3679),
]


@pytest.mark.parametrize("old, new, adv, expected", CLIP_CASES)
def test_clip_loss(old, new, adv, expected):
    r = P.clip_loss(torch.tensor(old), torch.tensor(new), torch.tensor(adv), 0.2)
    assert expected == pytest.approx(r.item(), 0.0001)


def _memory_fixture(obs_rows):
    """The 12-step, 4-rollout fixture of test_ppo.py:79-109 / test_memory.py:10-40 as plain arrays:
    dones at 3, 6, 11; a time-limit end at 9 -> rollouts of 4, 3, 3, 2 steps; chain of 13+3 observations."""
    rewards = np.array([i for i in range(10)] + [1, 2], np.float32)
    dones = np.zeros(12, np.float32); dones[[3, 6, 11]] = 1
    ends = dones.copy(); ends[9] = 1
    chain, joints, i, rollouts = [], [], 0, 0
    while rollouts < 4:
        rollouts += 1
        chain.append(obs_rows[i])
        end = False
        while not end:
            chain.append(obs_rows[i + 1])
            end = bool(ends[i])
            i += 1
        joints.append(len(chain))
    return np.array(chain, np.float32), joints, rewards, dones, ends


def test_obs_next_obs_joint_skipping():
    rows = [[i, 10 * i] for i in range(1, 14)]
    chain, joints, *_ = _memory_fixture(rows)
    oi, ni = P.chain_views(len(chain), joints)
    assert np.array_equal(chain[oi], np.array(rows[:12], np.float32))        # test_memory.py:43-46
    assert np.array_equal(chain[ni], np.array(rows[1:], np.float32))         # test_memory.py:49-52


def test_memory_mean_std_and_normalize():
    rows = [[i, 10 * i] for i in range(1, 14)]
    chain, joints, *_ = _memory_fixture(rows)
    obs = torch.from_numpy(chain)                       # Memory.update_obs_mean_std uses the whole chain (memory.py:284)
    mean, std = obs.mean(0), obs.std(0)
    assert torch.equal(mean, torch.tensor([7.1875, 71.8750]))                 # test_memory.py:56,62
    assert std[0].item() == pytest.approx(3.6737, abs=1e-4) and std[1].item() == pytest.approx(36.7367, abs=1e-4)
    st = NormStats(False, obs_mean=torch.tensor([2.5, 25.0]), obs_std=torch.tensor([2.0, 20.0]))
    ex = torch.tensor([[i, 10 * i] for i in range(6)]).float()
    expect = torch.tensor([[(i - 2.5) / 2, (i - 2.5) / 2] for i in range(6)])
    assert torch.equal(normalize(st, ex, force=True), expect)                 # test_memory.py:81-87
    ex[0, 0] = 1000
    expect[0, 0] = 10
    assert torch.equal(normalize(st, ex, force=True), expect)                 # clip at +-10, test_memory.py:89-92


def test_q_values_constant_critic():
    _, _, rewards, dones, _ = _memory_fixture([[1, 1]] * 13)
    q = P.q_values(torch.from_numpy(rewards), torch.from_numpy(dones), torch.full((12,), 10.0), 0.5)
    assert torch.equal(q, torch.tensor([5.0, 6.0, 7.0, 3.0, 9.0, 10.0, 6.0, 12.0, 13.0, 14.0, 6.0, 2.0]))   # test_a2c.py:45-56


def test_q_values_and_gae_truncation_bootstrap():
    chain, joints, rewards, dones, ends = _memory_fixture([[5, 5]] * 13)
    oi, ni = P.chain_views(len(chain), joints)
    obs, nobs = torch.from_numpy(chain[oi]), torch.from_numpy(chain[ni])
    critic = lambda x: x[:, 0] * 2                      # stub critic of test_ppo.py:120-124
    r, d, e = torch.from_numpy(rewards), torch.from_numpy(dones), torch.from_numpy(ends)
    q = P.q_values(r, d, critic(nobs), 0.5)
    assert torch.equal(q, torch.tensor([5.0, 6.0, 7.0, 3.0, 9.0, 10.0, 6.0, 12.0, 13.0, 14.0, 6.0, 2.0]))   # test_ppo.py:113-115,130-131
    adv = P.gae(q, critic(obs), critic(nobs), d, e, 0.5, 0.5).numpy()
    expected = np.array([-6.2969, -5.1875, -4.75, -7, -1.25, -1, -4, 3.1562, 4.6250, 6.5, -6, -8])           # test_ppo.py:116-118
    np.testing.assert_almost_equal(expected, adv, decimal=4)


def test_kl_divergence_and_standardize():
    lp = torch.tensor([-0.73, -0.72, -0.45], dtype=torch.float64)
    lq = torch.tensor([-0.57, -0.84, -0.13], dtype=torch.float64)
    assert -0.12 == pytest.approx(P.kl_divergence(lp, lq), 0.0001)            # test_utils.py:16-22
    obs = torch.tensor(np.arange(20).reshape(4, 5).T).float()
    st = NormStats(False, obs_mean=obs.mean(0), obs_std=obs.std(0))
    z = normalize(st, obs, force=True)
    assert z.mean().item() + 1 == pytest.approx(1.0, 0.0001)                  # test_utils.py:25-29
    assert z.std(0).mean().item() == pytest.approx(1.0, 0.0001)
