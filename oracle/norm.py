"""Observation normalisation / state-target denormalisation (ORACLE, test infra).

Follows rltoolkit/buffer/memory.py:76-127 (MemoryMeta.normalize / denormalize),
rltoolkit/buffer/replay_buffer.py:77-81 (obs_norm / force gate) and
rltoolkit/utils.py:62-73 (standardize_and_clip).  Note quirk 13 of SURVEY appendix B: min-max
normalize divides by (max - mid + 1e-8) while denormalize multiplies by (max - min)/2.
"""
import torch

MAX_ABS_OBS_VALUE = 10  # rltoolkit/config.py:66


class NormStats:
    def __init__(self, min_max_denormalize=False, min_obs=None, max_obs=None,
                 obs_mean=None, obs_std=None, obs_norm=False):
        self.min_max_denormalize = min_max_denormalize
        self.min_obs, self.max_obs = min_obs, max_obs
        self.obs_mean, self.obs_std = obs_mean, obs_std
        self.obs_norm = obs_norm


def normalize(st: NormStats, obs: torch.Tensor, force: bool = False) -> torch.Tensor:
    if not (st.obs_norm or force):                       # replay_buffer.py:77-81
        return obs
    if st.min_max_denormalize:                           # memory.py:77-82
        if st.min_obs is None and st.max_obs is None:
            return obs
        mean = (st.max_obs + st.min_obs) / 2
        return (obs - mean) / (st.max_obs - mean + 1e-8)
    if st.obs_std is None and st.obs_mean is None:       # memory.py:84-86
        return obs
    z = (obs - st.obs_mean) / (st.obs_std + 1e-8)        # utils.py:70-71
    return torch.clamp(z, -MAX_ABS_OBS_VALUE, MAX_ABS_OBS_VALUE)


def denormalize(st: NormStats, x: torch.Tensor) -> torch.Tensor:
    if st.min_max_denormalize:                           # memory.py:106-121
        mean = (st.max_obs + st.min_obs) / 2
        max_delta = (st.max_obs - st.min_obs) / 2
        return mean + x * max_delta
    return (st.obs_std + 1e-8) * x + st.obs_mean         # memory.py:123


def denorm_affine(st: NormStats):
    """(offset, scale) with denormalize(x) == offset + x*scale evaluated as mul-then-add."""
    if st.min_max_denormalize:
        return (st.max_obs + st.min_obs) / 2, (st.max_obs - st.min_obs) / 2
    return st.obs_mean, st.obs_std + 1e-8
