"""Rollout inner step on plain tensors (ORACLE, test infra).

off_policy_step   body of DDPG.collect_batch_and_train, rltoolkit/algorithms/ddpg/ddpg.py:202-207
                  with AcMOffPolicy.initial_act (rltoolkit/acm/off_policy/off_policy.py:50-54),
                  DDPG_AcM.noise_action (rltoolkit/acm/off_policy/ddpg_acm.py:40-50) and
                  AcMOffPolicy.process_action (off_policy.py:89-106), vectorised over E envs that
                  share one agent (the reference runs E = 1).
on_policy_step    A2C.collect_batch body, rltoolkit/algorithms/a2c/a2c.py:165-167, with
                  basic_model.Actor.act (rltoolkit/basic_model.py:32-51) and
                  AcMOnPolicyTrainer.process_action (rltoolkit/acm/on_policy.py:34-53); note quirk 18:
                  the ACM sees the NORMALISED obs next to the DENORMALISED target.
Noise tensors are inputs (the reference draws torch.randn / Normal.sample there).
"""
import torch

from . import nets
from .norm import NormStats, denormalize, normalize
from .offpolicy import sub


def off_policy_step(state, st: NormStats, obs, noise, actor_lim, acm_lim, act_noise, algo="sac",
                    eps=None, random_phase=False, denormalize_actor_out=True):
    """-> (state_target [E,ob] as stored in the ring, acm_action [E,ac]).

    random_phase: frames < random_frames -> target = actor_lim * noise (initial_act)."""
    x = normalize(st, obs)                                     # ddpg.py:203 (obs_norm gate)
    if random_phase:
        target = actor_lim * noise
    else:
        if algo == "sac":
            z, _, _ = nets.sac_actor_fwd(sub(state, "actor"), x, eps, actor_lim)
        else:
            z, _ = nets.ddpg_actor_fwd(sub(state, "actor"), x, actor_lim)
        z = z + (act_noise * noise) * actor_lim                # ddpg_acm.py:42-43
        lim = torch.as_tensor(actor_lim, dtype=torch.float32)
        target = torch.max(torch.min(z, 1.1 * lim), -1.1 * lim)    # np.clip, ddpg_acm.py:44-46
    if denormalize_actor_out:
        target = denormalize(st, target)
    acm_action, _ = nets.acm_fwd(sub(state, "acm"), torch.cat([x, target], dim=1), acm_lim)
    return target, acm_action


def on_policy_step(state, st: NormStats, obs, noise, actor_lim, acm_lim, denormalize_actor_out=True):
    """-> (sampled target [E,ob] (stored, normalised space), logp [E], acm_action [E,ac])."""
    x = normalize(st, obs, force=True)                         # Memory.normalize has no gate
    p = sub(state, "actor")
    mean, _ = nets.ppo_actor_mean(p, x, actor_lim)
    std = torch.exp(p["log_scale"])
    action = mean + noise * std                                # Normal.sample(): normal(loc, scale)
    var = std ** 2
    logp = (-((action - mean) ** 2) / (2 * var) - std.log() - nets.LOG_SQRT_2PI).sum(-1)
    target = denormalize(st, action) if denormalize_actor_out else action
    acm_action, _ = nets.acm_fwd(sub(state, "acm"), torch.cat([x, target], dim=1), acm_lim)
    return action, logp, acm_action
