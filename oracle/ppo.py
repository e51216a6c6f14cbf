"""SPP-PPO pieces on plain tensors (ORACLE, test infra).

  chain_views        Memory.obs / Memory.next_obs joint skipping, rltoolkit/buffer/memory.py:146-168
  q_values           A2C.calculate_q_val, rltoolkit/algorithms/a2c/a2c.py:247-265
  gae                PPO.calculate_gae, rltoolkit/algorithms/ppo/ppo.py:117-150 (reverse scan; carry
                     reset at done, bootstrap with V(s') at non-terminal ends -- quirk 14)
  normalize_adv      AdvantageDataset.__init__, rltoolkit/algorithms/ppo/advantage_dataset.py:9-12
                     (unbiased std, eps 1.2e-7)
  clip_loss          PPO._clip_loss, ppo.py:194-204
  kl_divergence      rltoolkit/utils.py:48-59
  critic_fit_step    inner Adam step of A2C.update_critic, a2c.py:186-225
  actor_minibatch    one minibatch of PPO_AcM.update_actor_acm, rltoolkit/acm/on_policy.py:180-205
                     (Normal/Independent log_prob + entropy, basic_model.py:53-62)
"""
import math

import torch

from . import nets
from .adam import adam_step_net
from .offpolicy import sub

HALF_LOG_2PI = 0.5 * math.log(2 * math.pi)


def chain_views(n_obs, new_rollout_idx):
    """Index lists (into the obs chain) of Memory.obs and Memory.next_obs."""
    obs, nxt, r = [], [], 0
    for i in range(n_obs):
        if i != new_rollout_idx[r] - 1:
            obs.append(i)
        elif r < len(new_rollout_idx) - 1:
            r += 1
    r = 0
    for i in range(1, n_obs):
        if i != new_rollout_idx[r]:
            nxt.append(i)
        elif r < len(new_rollout_idx) - 1:
            r += 1
    return obs, nxt


def q_values(reward, done, next_value, gamma):
    """q = r + gamma * (1 - done) * V(s')  (float32 tensors)."""
    return reward + gamma * (1 - done) * next_value


def gae(q_val, value, next_value, done, end, gamma, lam):
    """Reverse scan of ppo.py:139-148; the carry is float32 like `gae * discount + delta`."""
    deltas = q_val - value
    adv = torch.empty_like(deltas)
    discount = lam * gamma
    carry = 0
    for i in range(len(deltas) - 1, -1, -1):
        if done[i]:
            carry = 0
        elif end[i]:
            carry = next_value[i].item()
        carry = carry * discount + deltas[i]
        adv[i] = carry
    return adv


def normalize_adv(adv, eps=1.2e-7):
    return (adv - torch.mean(adv)) / (torch.std(adv) + eps)


def clip_loss(old_logp, new_logp, adv, epsilon):
    ratio = torch.exp(new_logp - old_logp)
    clipped = torch.clamp(ratio, 1 - epsilon, 1 + epsilon)
    return -(torch.min(ratio * adv, clipped * adv)).mean()


def kl_divergence(log_p, log_q):
    return (log_p - log_q).mean().item()


def gauss_logp(actions, mean, log_scale):
    """Independent(Normal(mean, exp(log_scale)), 1).log_prob(actions)."""
    std = torch.exp(log_scale)
    var = std ** 2
    return (-((actions - mean) ** 2) / (2 * var) - std.log() - nets.LOG_SQRT_2PI).sum(-1)


def gauss_entropy(log_scale, n):
    """Independent(Normal).entropy() -> [n] (state independent): sum_j 0.5 + 0.5 log(2 pi) + log std."""
    std = torch.exp(log_scale)
    return (0.5 + HALF_LOG_2PI + torch.log(std)).sum().expand(n)


def critic_fit_step(state, obs, q_val, lr):
    """One Adam step on 0.5 * mean((q - V(obs))^2); returns the loss (a2c.py:209-216)."""
    p = sub(state, "critic")
    v, cache = nets.ppo_critic_fwd(p, obs)
    adv = q_val - v.squeeze(-1)
    loss = float((0.5 * adv.pow(2).mean()).item())
    dv = (-(adv) / adv.numel()).unsqueeze(-1)
    adam_step_net(state, "critic", nets.ppo_critic_bwd(p, cache, dv), lr)
    return loss


def actor_minibatch(state, norm_obs, actions, old_logp, adv, lim, lr, epsilon, entropy_coef=0.0,
                    custom_loss=0.0, next_obs=None):
    """One minibatch step of PPO_AcM.update_actor_acm; returns (losses, new_logp).

    `actions` / `next_obs` are already denormalised when norm_closs is False (on_policy.py:187-189).
    The custom-loss MSE term involves no actor parameter (quirk 17): it only shifts the loss."""
    p = sub(state, "actor")
    B = norm_obs.shape[0]
    mean, cache = nets.ppo_actor_mean(p, norm_obs, lim)
    ls = p["log_scale"]
    new_logp = gauss_logp(actions, mean, ls)
    ent = gauss_entropy(ls, B).mean()
    ratio = torch.exp(new_logp - old_logp)
    clipped = torch.clamp(ratio, 1 - epsilon, 1 + epsilon)
    s1, s2 = ratio * adv, clipped * adv
    actor_loss = -(torch.min(s1, s2)).mean()
    ppo_loss = actor_loss - entropy_coef * ent
    losses = {"actor": float(actor_loss.item()), "entropy": float(ent.item())}
    policy_loss = ppo_loss
    if custom_loss:
        dist = ((actions - next_obs) ** 2).mean()
        losses["dist"] = float(dist.item())
        policy_loss = ppo_loss + custom_loss * dist
    losses["policy"] = float(policy_loss.item())
    # backward: d(-mean(min(s1,s2)))
    g = torch.full_like(s1, -1.0 / B)
    tie = (s1 == s2).to(g.dtype)
    g1 = g * (s1 < s2).to(g.dtype) + 0.5 * g * tie
    g2 = g * (s2 < s1).to(g.dtype) + 0.5 * g * tie
    in_range = ((ratio >= 1 - epsilon) & (ratio <= 1 + epsilon)).to(g.dtype)
    dratio = g1 * adv + g2 * adv * in_range
    dlogp = (dratio * ratio).unsqueeze(-1)
    std = torch.exp(ls)
    var = std ** 2
    d = actions - mean
    dmean = dlogp * (d / var)
    # d logp / d log_scale = d^2/var - 1 ; entropy adds -entropy_coef * 1 per dim
    dls = (dlogp * ((d * d) / var - 1.0)).sum(0) - entropy_coef * torch.ones_like(ls)
    grads = nets.ppo_actor_mean_bwd(p, cache, dmean, lim)
    grads["log_scale"] = dls
    adam_step_net(state, "actor", grads, lr)
    return losses, new_logp


def update_critic(state, norm_obs, norm_next_obs, reward, done, gamma, lr, n_target_updates=10,
                  n_updates_per_target=10):
    """A2C.update_critic fitting loop, a2c.py:203-221 -> mean critic loss."""
    total = 0.0
    for _ in range(n_target_updates):
        nv, _ = nets.ppo_critic_fwd(sub(state, "critic"), norm_next_obs)
        q = q_values(reward, done, nv.squeeze(-1), gamma)
        for _ in range(n_updates_per_target):
            total += critic_fit_step(state, norm_obs, q, lr)
    return total / (n_target_updates * n_updates_per_target)


def advantages(state, norm_obs, norm_next_obs, reward, done, end, gamma, lam):
    """PPO.calculate_advantage, ppo.py:101-115: q-values then GAE with the fitted critic."""
    p = sub(state, "critic")
    nv = nets.ppo_critic_fwd(p, norm_next_obs)[0].squeeze(-1)
    v = nets.ppo_critic_fwd(p, norm_obs)[0].squeeze(-1)
    q = q_values(reward, done, nv, gamma)
    return gae(q, v, nv, done, end, gamma, lam), q


def update_actor_acm(state, norm_obs, actions, next_obs, old_logp, adv_norm, perms, lim, lr,
                     epsilon, kl_threshold, max_epochs, batch_size, entropy_coef=0.0,
                     custom_loss=0.0, plain=False):
    """PPO_AcM.update_actor_acm epoch loop, rltoolkit/acm/on_policy.py:164-216, with the shuffles
    injected (`perms[e]` replaces DataLoader(shuffle=True)'s torch.randperm of epoch e).

    Reproduces: the partial last minibatch is kept; KL is taken on the LAST minibatch only
    (quirk 16); the summed losses are divided by (i + 1) where i is the loop variable at exit --
    when the KL test breaks at the top of iteration i, that is one more than the epochs run.
    plain=True: PPO.update_actor (rltoolkit/algorithms/ppo/ppo.py:152-192), what PPO_AcM.update_actor runs when custom_loss == 0
    (on_policy.py:88-98): `actions` are the stored (normalised-space) actions, there is no distance term, and the loss sums
    `actor`, `entropy`, `sum` (returned under "policy") are NOT divided by the epoch count.
    -> (losses, epochs_run, kl)"""
    N = norm_obs.shape[0]
    tot = {"actor": 0.0, "entropy": 0.0, "policy": 0.0, "dist": 0.0}
    kl, i, epochs_run = 0.0, 0, 0
    for i in range(max_epochs):
        if kl >= kl_threshold:
            break
        perm = perms[i]
        for s in range(0, N, batch_size):
            idx = perm[s:s + batch_size]
            losses, new_logp = actor_minibatch(
                state, norm_obs[idx], actions[idx], old_logp[idx], adv_norm[idx], lim, lr, epsilon,
                entropy_coef, custom_loss, None if next_obs is None else next_obs[idx])
            for k in tot:
                tot[k] += losses.get(k, 0.0)
        kl = kl_divergence(old_logp[idx], new_logp)
        epochs_run += 1
    if not plain:
        for k in tot:
            tot[k] /= i + 1
    return tot, epochs_run, kl


def a2c_advantages(state, norm_obs, norm_next_obs, reward, done, gamma):
    """A2C.calculate_advantage, rltoolkit/algorithms/a2c/a2c.py:227-245: q - V(s) with the fitted critic (no GAE)."""
    p = sub(state, "critic")
    nv = nets.ppo_critic_fwd(p, norm_next_obs)[0].squeeze(-1)
    v = nets.ppo_critic_fwd(p, norm_obs)[0].squeeze(-1)
    return q_values(reward, done, nv, gamma) - v


def a2c_actor_step(state, norm_obs, actions, old_logp, adv, lim, lr, accumulate, normalize=True,
                   custom_loss=0.0, loss_actions=None, next_obs=None):
    """A2C_AcM.update_actor_acm (rltoolkit/acm/on_policy.py:100-124; accumulate=True) / A2C.update_actor
    (rltoolkit/algorithms/a2c/a2c.py:267-285; accumulate=False): ONE full-batch step on mean(-logp * adv).

    The log-probs are the ones the rollout recorded (their autograd graph is taken at the current weights: no update
    happened in between), so the gradient is that of the Gaussian log-density at the stored actions.  The distance term
    MSE(actions, next_obs) has no gradient path (Actor.act samples without rsample, basic_model.py:47).
    update_actor_acm never calls zero_grad (quirk 21): the optimiser steps on the SUM of this and all earlier
    iterations' gradients, kept in state["actor#gacc"].  -> losses dict"""
    if normalize:
        adv = (adv - adv.mean()) / (adv.std() + 1e-8)
    p = sub(state, "actor")
    n = norm_obs.shape[0]
    mean, cache = nets.ppo_actor_mean(p, norm_obs, lim)
    ls = p["log_scale"]
    actor_loss = (-old_logp * adv).mean()
    losses = {"actor": float(actor_loss.item())}
    if custom_loss:
        dist = ((loss_actions - next_obs) ** 2).mean()
        losses["dist"] = float(dist.item())
        losses["policy"] = float((actor_loss + custom_loss * dist).item())
    dlogp = (-adv / n).unsqueeze(-1)
    var = torch.exp(ls) ** 2
    d = actions - mean
    grads = nets.ppo_actor_mean_bwd(p, cache, dlogp * (d / var), lim)
    grads["log_scale"] = (dlogp * ((d * d) / var - 1.0)).sum(0)
    if accumulate:
        acc = state.setdefault("actor#gacc", {})
        for k, g in grads.items():
            acc[k] = acc[k] + g if k in acc else g.clone()
        grads = {k: v.clone() for k, v in acc.items()}
    adam_step_net(state, "actor", grads, lr)
    return losses
