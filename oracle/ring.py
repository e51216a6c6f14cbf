"""Replay ring state machine and minibatch gather in numpy (ORACLE, test infra).

Follows rltoolkit/buffer/replay_buffer.py:
  MetaReplayBuffer.add_obs :56-60, add_timestep :65-75 (separate obs / timestep cursors, reset rule
  `next_obs_idx < ts_idx`, `current_len` may shrink -- SURVEY appendix B quirk 9),
  ReplayBuffer.addition :133-137, BufferAcMOffPolicy.add_acm_action :332-333,
  ReplayBuffer._sample_batch :233-261 and BufferAcMOffPolicy.sample_batch :385-398 (two-level
  gather obs = _obs[_obs_idx[idx]]), rbuffer_sample_acm :404-430,
  ReplayBufferAcM.add_buffer :284-297 (quirk 19: drops / corrupts one transition per rollout joint).
Storage is float32 here (the reference stores float64 copies of float32 values and casts back at
sample time -- a lossless round trip, quirk 10); indices are int64 and must match bit-exactly.
"""
import numpy as np


class Ring:
    def __init__(self, size, ob_dim, act_dim, acm_dim):
        self.size = int(size)
        self.obs = np.zeros((size, ob_dim), np.float32)
        self.obs_idx = np.zeros(size, np.int64)
        self.next_obs_idx = np.zeros(size, np.int64)
        self.actions = np.zeros((size, act_dim), np.float32)
        self.rewards = np.zeros(size, np.float32)
        self.done = np.zeros(size, np.bool_)
        self.end = np.zeros(size, np.bool_)
        self.actions_acm = np.zeros((size, acm_dim), np.float32)
        self.obs_cur = 0      # reference: self.obs_idx
        self.ts_cur = 0       # reference: self.ts_idx
        self.current_len = 0

    def add_obs(self, obs):
        self.obs[self.obs_cur] = obs
        i = self.obs_cur
        self.obs_cur = (self.obs_cur + 1) % self.size
        return i

    def add_acm_action(self, a):
        self.actions_acm[self.ts_cur] = a

    def add_timestep(self, obs_idx, next_obs_idx, action=None, rew=0.0, done=False, end=False):
        self.obs_idx[self.ts_cur] = obs_idx
        self.next_obs_idx[self.ts_cur] = next_obs_idx
        if action is not None:
            self.actions[self.ts_cur] = action
            self.rewards[self.ts_cur] = rew
            self.done[self.ts_cur] = done
            self.end[self.ts_cur] = end
        if next_obs_idx < self.ts_cur:
            self.current_len = self.ts_cur + 1
            self.ts_cur = 0
        else:
            self.ts_cur += 1
        self.current_len = max(self.ts_cur, self.current_len)

    def gather(self, idxs):
        """-> obs, next_obs, action, reward, done(int8), acm_action for int64 idxs [B]."""
        idxs = np.asarray(idxs, np.int64)
        return (self.obs[self.obs_idx[idxs]], self.obs[self.next_obs_idx[idxs]], self.actions[idxs],
                self.rewards[idxs], self.done[idxs].astype(np.int8), self.actions_acm[idxs])

    def gather_acm(self, idxs):
        """rbuffer_sample_acm: -> obs, next_obs, actions_acm."""
        idxs = np.asarray(idxs, np.int64)
        return self.obs[self.obs_idx[idxs]], self.obs[self.next_obs_idx[idxs]], self.actions_acm[idxs]


def add_rollouts_to_acm_ring(ring: Ring, chain_obs, actions_acm, new_rollout_idx):
    """ReplayBufferAcM.add_buffer, replay_buffer.py:284-297, INCLUDING its joint behaviour.

    `chain_obs` is Memory._obs (T_k+1 entries per rollout, concatenated), `new_rollout_idx` is
    Memory._new_rollout_idx (cumulative chain lengths).  At every joint the loop `continue`s after
    bumping i, which skips the first transition of the new rollout and then pairs the last obs of
    the previous rollout with the 2nd obs of the new one."""
    joints = set(int(j) for j in new_rollout_idx)
    i = 0
    obs_idx = ring.add_obs(chain_obs[i])
    for acm_action in actions_acm:
        i += 1
        next_idx = ring.add_obs(chain_obs[i])
        if i in joints:
            i += 1
            continue
        ring.actions_acm[ring.ts_cur] = acm_action
        ring.add_timestep(obs_idx, next_idx)
        obs_idx = next_idx
