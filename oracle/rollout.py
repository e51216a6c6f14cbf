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


def on_policy_rollout_synthetic(state, st: NormStats, E, T, noise_act, noise_env, u_done, noise_reset, actor_lim, acm_lim,
                                max_ep_len, done_prob, obs0=None, ep_len0=None, denormalize_actor_out=True):
    """A2C.collect_batch (rltoolkit/algorithms/a2c/a2c.py:144-184) over E environments for T steps, every step through
    on_policy_step above (the part pinned by the reference fixture).  MuJoCo is unavailable, so the environment is THIS
    PROJECT'S synthetic one (not the reference's; restated here only so that the device loop has a checker):
        mix = tanh(sum_j a_j (0.3 + 0.1 j));  obs'_j = 0.98 obs_j + 0.1 mix (1 - 0.01 j) + 0.02 N(0,1);  reward = obs'_0;
        done = u < done_prob;  end = done or ep_len == max_ep_len;  done = False if ep_len == max_ep_len else done (a2c.py:168-171);
        reset -> 0.1 N(0,1).
    Noise: noise_act / noise_env / noise_reset [T, E, ob], u_done [T, E].  Rows are step-major (row = t * E + e); `end` is also set on
    the last step of the batch (every trajectory is cut there).  -> dict of [T*E, ...] tensors + final obs / ep_len."""
    ob = noise_act.shape[-1]
    obs = torch.zeros(E, ob) if obs0 is None else obs0.clone()
    ep_len = torch.zeros(E, dtype=torch.int64) if ep_len0 is None else ep_len0.clone()
    out = {k: [] for k in ("x", "xn", "act", "logp", "rew", "done", "end", "aacm", "raw_obs", "raw_next")}
    ac = None
    for t in range(T):
        action, logp, a = on_policy_step(state, st, obs, noise_act[t], actor_lim, acm_lim, denormalize_actor_out)
        ac = a.shape[1]
        w = 0.3 + 0.1 * torch.arange(ac, dtype=torch.float32)
        mix = torch.zeros(E)
        for j in range(ac):                                    # the device adds j ascending in fp32
            mix = mix + a[:, j] * w[j]
        mix = torch.tanh(mix)
        scale = 1.0 - 0.01 * torch.arange(ob, dtype=torch.float32)
        nx = (0.98 * obs + (0.1 * mix)[:, None] * scale) + 0.02 * noise_env[t]
        ep_len = ep_len + 1
        done = u_done[t] < done_prob
        limit = ep_len == max_ep_len
        end = done | limit
        out["x"].append(normalize(st, obs, force=True)); out["xn"].append(normalize(st, nx, force=True))
        out["act"].append(action); out["logp"].append(logp); out["rew"].append(nx[:, 0].clone())
        out["done"].append((done & ~limit).float()); out["end"].append((end | (t == T - 1)).float())
        out["aacm"].append(a); out["raw_obs"].append(obs.clone()); out["raw_next"].append(nx.clone())
        obs = torch.where(end[:, None], 0.1 * noise_reset[t], nx)
        ep_len = torch.where(end, torch.zeros_like(ep_len), ep_len)
    res = {k: torch.cat(v, dim=0) for k, v in out.items()}
    res["obs_final"], res["ep_len_final"] = obs, ep_len
    return res
