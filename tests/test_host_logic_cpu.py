"""CPU: host-side mirrors that carry no arithmetic but decide WHICH rows and WHICH random numbers the device path sees --
checked against torch's own DataLoader sampler and against the oracle's restatement of the reference's memory views."""
import numpy as np
import pytest
import torch
from torch.utils.data import DataLoader, RandomSampler, TensorDataset

from oracle import ppo as P
from spp_rl_b200.rltoolkit_api import ValidationRing, sampler_permutation
from spp_rl_b200.rltoolkit_ppo import PPO_DEFAULTS, RolloutMemory


def test_sampler_permutation_is_the_dataloaders_shuffle():
    """update_acm / update_actor_acm iterate DataLoader(shuffle=True) (acm.py:285, on_policy.py:166): same seed -> same order and
    the same state of torch's global generator afterwards (so later draws stay aligned with the reference as well)."""
    n = 257
    for seed in (0, 7):
        torch.manual_seed(seed)
        ref = [list(RandomSampler(range(n))) for _ in range(3)]
        after_ref = torch.rand(1).item()
        torch.manual_seed(seed)
        got = [sampler_permutation(n, dataloader=False).tolist() for _ in range(3)]
        after = torch.rand(1).item()
        assert got == ref and after == after_ref
    torch.manual_seed(3)
    x = torch.arange(n, dtype=torch.float32)[:, None]
    batches = [b[0][:, 0].long().tolist() for b in DataLoader(TensorDataset(x), batch_size=64, shuffle=True)]
    torch.manual_seed(3)
    perm = sampler_permutation(n).tolist()
    assert [perm[i:i + 64] for i in range(0, n, 64)] == batches and len(batches[-1]) == n - 4 * 64      # partial last minibatch kept


def test_rollout_memory_views_skip_the_joints_like_the_reference():
    """Memory.obs / next_obs (rltoolkit/buffer/memory.py:146-170) over a chain with three rollouts == the oracle's chain_views."""
    rng = np.random.RandomState(0)
    m = RolloutMemory()
    lens = [4, 1, 6]
    for L in lens:
        prev = m.add_obs(torch.from_numpy(rng.randn(1, 5).astype(np.float32)))
        for t in range(L):
            nxt = m.add_obs(torch.from_numpy(rng.randn(1, 5).astype(np.float32)))
            m.add_timestep(prev, nxt, torch.zeros(1, 5), torch.zeros(1), 0.0, False, t == L - 1)
            m.add_acm_action(np.zeros(2, np.float32))
            prev = nxt
        m.end_rollout()
    chain = torch.cat(m._obs)
    oi, ni = P.chain_views(len(chain), list(m._new_rollout_idx))
    assert len(m) == sum(lens) == len(oi) and m._new_rollout_idx == [5, 7, 14]
    assert torch.equal(m.obs, chain[oi]) and torch.equal(m.next_obs, chain[ni])
    assert [i for i, e in enumerate(m.end) if e] == [3, 4, 10]


def test_validation_ring_follows_the_replay_state_machine():
    """ValidationRing == MetaReplayBuffer's cursor rules (replay_buffer.py:56-75): wrap of the observation cursor shrinks current_len."""
    r = ValidationRing(6, 2, 1)
    prev = r.add_obs(np.zeros(2))
    for t in range(9):
        nxt = r.add_obs(np.full(2, t + 1.0))
        r.add_timestep(prev, nxt, np.array([float(t)]))
        prev = nxt
    # oracle: the same sequence through the reference-pinned ring restatement
    from oracle.ring import Ring
    o = Ring(6, 2, 2, 1)
    prev = o.add_obs(np.zeros(2))
    for t in range(9):
        nxt = o.add_obs(np.full(2, t + 1.0))
        o.add_acm_action(np.array([float(t)]))
        o.add_timestep(prev, nxt)
        prev = nxt
    assert (r.obs_idx, r.ts_idx, r.current_len) == (o.obs_cur, o.ts_cur, o.current_len)
    n = r.current_len
    assert np.array_equal(r._obs_idx[:n], o.obs_idx[:n]) and np.array_equal(r._next_obs_idx[:n], o.next_obs_idx[:n])
    assert np.array_equal(r.actions_acm, o.actions_acm[:n]) and np.array_equal(r.obs, o.obs[o.obs_idx[:n]])


def test_ppo_defaults_are_the_reference_config():
    assert PPO_DEFAULTS["actor_lr"] == 3e-3 and PPO_DEFAULTS["critic_lr"] == 3e-4 and PPO_DEFAULTS["ppo_batch_size"] == 1000
    assert PPO_DEFAULTS["kl_div_threshold"] == 0.15 and PPO_DEFAULTS["max_ppo_epochs"] == 50 and PPO_DEFAULTS["gae_lambda"] == 0.95
    assert "tau" not in PPO_DEFAULTS and "buffer_size" not in PPO_DEFAULTS


def test_synthetic_environment_fallback_is_loud(monkeypatch):
    """envs.make: without gym the MuJoCo names are shape-only stand-ins -- a RuntimeWarning unless the caller opted in."""
    import warnings

    from spp_rl_b200 import envs
    monkeypatch.setattr(envs, "_real_gym", lambda: None)
    monkeypatch.delenv("SPP_RL_SYNTHETIC_ENVS", raising=False)
    with pytest.warns(RuntimeWarning, match="SHAPE-ONLY"):
        e = envs.make("Hopper-v2")
    assert e.observation_space.shape == (11,)
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        envs.make("Hopper-v2", synthetic=True)
        envs.make("Pendulum-v0")                      # faithful classic-control dynamics: no warning
    with pytest.raises(KeyError):
        envs.make("NoSuchEnv-v0")

    class FakeGym:
        @staticmethod
        def make(name):
            return ("real", name)
    monkeypatch.setattr(envs, "_real_gym", lambda: FakeGym)
    assert envs.make("Hopper-v2") == ("real", "Hopper-v2")      # a real gym wins over the stand-in


def _frame_by_frame(frames, batch, update_freq, grad_steps, random_frames, steps_per_epoch, acm_freq, acm_batches):
    """The reference's frame loop, one frame at a time (ddpg.py:191-237, ddpg_acm.py:52-85, ddpg.py:159-169): -> event list with
    consecutive event-free frames of one phase merged into a single ("rollout", n, random, first_frame)."""
    ev, run, it, f = [], None, 0, 0
    while f < frames:
        random = f < random_frames                      # stats_logger.frames < random_frames, tested BEFORE the increment
        if run is not None and run[2] == random:
            run[1] += 1
        else:
            if run is not None:
                ev.append(tuple(run))
            run = ["rollout", 1, random, f]
        f += 1
        fired = []
        if f > batch and f % update_freq == 0:          # len(replay_buffer) > update_batch_size: one transition per frame
            fired.append(("update", grad_steps))
        if it > 0 and acm_batches > 0 and f % acm_freq == 0:
            fired.append(("acm", acm_batches))
        if f % steps_per_epoch == 0:
            fired.append(("stats",))
            it += 1
        if fired:
            ev.append(tuple(run)); run = None
            ev.extend(fired)
    if run is not None:
        ev.append(tuple(run))
    return ev


def test_population_train_schedule_fires_where_the_reference_loop_does():
    from spp_rl_b200.population import train_schedule

    for cfg in ((3000, 256, 50, 50, 100, 1000, 200, 10), (2500, 64, 7, 3, 21, 350, 35, 2), (1200, 256, 50, 50, 0, 400, 0, 0),
                (999, 10, 1, 1, 5, 100, 3, 1)):
        frames, batch, uf, gs, rf, spe, af, ab = cfg
        want = _frame_by_frame(frames, batch, uf, gs, rf, spe, af, ab)
        got = list(train_schedule(0, frames, 1, batch, uf, gs, rf, spe, af, ab))
        # the schedule cuts rollouts at every POSSIBLE event frame (the ring length is only known at run time); merge neighbours
        merged = []
        for e in got:
            if e[0] == "rollout" and merged and merged[-1][0] == "rollout" and merged[-1][2] == e[2]:
                merged[-1] = ("rollout", merged[-1][1] + e[1], e[2], merged[-1][3])
            else:
                merged.append(e)
        assert merged == want, cfg
    # E environments per agent: one vector step = E frames; the same frames fire
    got = [e for e in train_schedule(0, 2000, 10, 256, 50, 50, 100, 1000, 200, 10) if e[0] != "rollout"]
    want = [e for e in _frame_by_frame(2000, 256, 50, 50, 100, 1000, 200, 10) if e[0] != "rollout"]
    assert got == want
    assert sum(e[1] * 10 for e in train_schedule(0, 2000, 10, 256, 50, 50, 100, 1000, 200, 10) if e[0] == "rollout") == 2000


def test_interleaved_burst_schedule_runs_every_step_once_in_order_and_never_deadlocks():
    """update_burst_interleaved_kernel (populations larger than the grid; csrc/update_kernel.cu) restated on the host
    (sharding.burst_items): every (agent, step) is executed exactly once, an agent's steps in order, and an event-driven run with
    arbitrary item durations -- every CTA takes its items in order and waits for the agent's progress word -- always finishes
    (each item only waits for a smaller one); with equal step times 256 agents x 50 steps on 148 CTAs take 88 step times, not 100."""
    from spp_rl_b200.sharding import burst_items, burst_step_times
    rng = np.random.RandomState(0)
    for P_, G, grid in ((256, 50, 148), (188, 3, 148), (149, 7, 148), (300, 1, 148), (5, 4, 2), (7, 5, 3), (148, 50, 148), (3, 9, 8)):
        lists = [burst_items(P_, G, grid, c) for c in range(min(grid, P_))]
        seen = {}
        for items in lists:
            for agent, b, e in items:
                for g in range(b, e):
                    assert (agent, g) not in seen
                    seen[(agent, g)] = True
        assert len(seen) == P_ * G
        # event-driven run: CTA c is free at t[c]; its next item starts at max(t[c], time its agent reached `b` steps)
        dur = lambda: float(rng.rand() * 2 + 0.1)      # noqa: E731
        t = [0.0] * len(lists); pos = [0] * len(lists)
        progress = {a: (0, 0.0) for a in range(P_)}      # agent -> (steps complete, time of completion)
        remaining = sum(len(x) for x in lists)
        while remaining:
            ran = False
            for c, items in enumerate(lists):
                if pos[c] == len(items):
                    continue
                agent, b, e = items[pos[c]]
                done, when = progress[agent]
                if done < b:
                    continue      # the CTA spins on the progress word
                assert done == b, "steps of an agent out of order"
                start = max(t[c], when)
                t[c] = start + sum(dur() for _ in range(b, e))
                progress[agent] = (e, t[c])
                pos[c] += 1; remaining -= 1; ran = True
            assert ran, "deadlock: no CTA can take its next item"
        assert all(progress[a][0] == G for a in range(P_))
    assert burst_step_times(256, 50, 148) == 88 and burst_step_times(256, 50, 148, item_steps=50) == 100
    assert burst_step_times(148, 50, 148) == 50 and burst_step_times(128, 50, 148) == 50
