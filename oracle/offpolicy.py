"""SPP-SAC and SPP-DDPG update steps on plain tensors (ORACLE, test infra).

`sac_acm_update`  restates SAC_AcM.update,  rltoolkit/acm/off_policy/sac_acm.py:89-162
                  (+ compute_qfunc_targ :30-58, compute_pi_loss :60-87,
                   SAC.update_target_q rltoolkit/algorithms/sac/sac.py:186-199,
                   SAC.compute_alpha_loss sac.py:201-216).
`ddpg_acm_update` restates DDPG_AcM.update, rltoolkit/acm/off_policy/ddpg_acm.py:147-201
                  (+ compute_qfunc_targ :100-123, compute_pi_loss :125-145,
                   DDPG.update_target_nets rltoolkit/algorithms/ddpg/ddpg.py:273-284).

State layout: dict of torch CPU tensors keyed '<net>.<param>' with the reference's state_dict
names; nets: actor, critic_1, critic_2, critic_1_targ, critic_2_targ, acm (SAC) and actor,
actor_targ, critic, critic_targ, acm (DDPG).  Adam moments at '<key>#m'/'#v', counters at
'<net>#step'.  SAC temperature: 'log_alpha' (float64 0-dim) + '#m'/'#v'/'#step'.
The random draws of Normal.rsample are INPUTS (eps_targ for the target pass on next_obs, eps_pi
for the policy pass on obs -- the order the reference draws them in).
"""
import math

import torch

from . import nets
from .adam import BETA1, BETA2, EPS, adam_step_net
from .norm import NormStats, denormalize, normalize


class OffPolicyHP:
    """Hyper-parameters read by the update step (names follow the reference attributes)."""

    def __init__(self, gamma=0.95, actor_lr=1e-3, critic_lr=1e-3, alpha_lr=1e-3, tau=0.005,
                 custom_loss=0.0, norm_closs=True, acm_critic=False, target_entropy=-1.0,
                 actor_lim=1.0, acm_lim=1.0):
        self.gamma, self.actor_lr, self.critic_lr, self.alpha_lr = gamma, actor_lr, critic_lr, alpha_lr
        self.tau, self.custom_loss, self.norm_closs, self.acm_critic = tau, custom_loss, norm_closs, acm_critic
        self.target_entropy = target_entropy
        self.actor_lim = actor_lim      # actor_ac_lim, rltoolkit/acm/acm.py:102-108 (scalar or [ob])
        self.acm_lim = acm_lim          # env action_space.high, rltoolkit/rl.py:53 ([ac])


def sub(state, net):
    pre = net + "."
    return {k[len(pre):]: v for k, v in state.items() if k.startswith(pre) and "#" not in k}


def _min_bwd(g, a, b):
    """Backward of torch.min(a, b) (== minimum): ties split the gradient in half."""
    tie = (a == b).to(g.dtype)
    ga = g * (a < b).to(g.dtype) + 0.5 * g * tie
    gb = g * (b < a).to(g.dtype) + 0.5 * g * tie
    return ga, gb


def _polyak(state, net, targ, tau):
    """targ.mul_(1 - tau); targ.add_(tau * param)  (sac.py:197-199, ddpg.py:282-284)."""
    for k, v in sub(state, net).items():
        t = state[targ + "." + k]
        t.mul_(1 - tau)
        t.add_(tau * v)


def sac_alpha(state) -> float:
    """self.alpha = self.log_alpha.exp().item()  (sac_acm.py:159); Python float (fp64)."""
    return float(state["log_alpha"].exp().item())


def sac_acm_update(state, hp: OffPolicyHP, st: NormStats, obs, next_obs, action, reward, done,
                   acm_action, eps_targ, eps_pi, alpha=None, capture=None):
    """One SAC_AcM.update.  Mutates `state`; returns (losses dict, new alpha).

    `alpha` is the Python-float temperature carried between updates (defaults to exp(log_alpha)).
    `capture`, if a dict, receives intermediates (y, q's, grads) for kernel bisection."""
    if alpha is None:
        alpha = sac_alpha(state)
    loss = {}
    actor, c1, c2 = sub(state, "actor"), sub(state, "critic_1"), sub(state, "critic_2")
    c1t, c2t, acm = sub(state, "critic_1_targ"), sub(state, "critic_2_targ"), sub(state, "acm")
    B = obs.shape[0]
    if hp.acm_critic:                                   # sac_acm.py:111-112
        action = acm_action

    # ---- Phase A: Q target (no grad), sac_acm.py:43-56
    z_n, logp_n, _ = nets.sac_actor_fwd(actor, next_obs, eps_targ, hp.actor_lim)
    zd_n = denormalize(st, z_n)
    if hp.acm_critic:
        a_n, _ = nets.acm_fwd(acm, torch.cat([next_obs, zd_n], dim=1), hp.acm_lim)
    else:
        a_n = zd_n
    q1t, _ = nets.critic_fwd(c1t, next_obs, a_n)
    q2t, _ = nets.critic_fwd(c2t, next_obs, a_n)
    q_t = torch.min(q1t, q2t)
    y = reward + hp.gamma * (1 - done) * (q_t - alpha * logp_n)

    # ---- Phase B: critics, sac_acm.py:117-131  (mse_loss mean; Adam per critic)
    for name, p in (("critic_1", c1), ("critic_2", c2)):
        q, cache = nets.critic_fwd(p, obs, action)
        diff = q - y
        loss[name] = float((diff * diff).mean().item())
        grads, _ = nets.critic_bwd(p, cache, 2.0 * diff / B)
        if capture is not None:
            capture[name + ".q"] = q.clone()
            capture.update({name + ".grad." + k: v.clone() for k, v in grads.items()})
        adam_step_net(state, name, grads, hp.critic_lr)

    # ---- Phase C: policy, sac_acm.py:137-145 + compute_pi_loss :60-87 (post-step critics)
    z, logp, a_cache = nets.sac_actor_fwd(actor, obs, eps_pi, hp.actor_lim)
    zd = denormalize(st, z)
    if hp.acm_critic:
        a_pi, m_cache = nets.acm_fwd(acm, torch.cat([obs, zd], dim=1), hp.acm_lim)
    else:
        a_pi = zd
    q1, c1_cache = nets.critic_fwd(c1, obs, a_pi)
    q2, c2_cache = nets.critic_fwd(c2, obs, a_pi)
    q = torch.min(q1, q2)
    pi_loss = (alpha * logp - q).mean()
    total = pi_loss
    d_scale = st_scale(st)
    dz = torch.zeros_like(z)
    dzd = torch.zeros_like(z)
    if hp.custom_loss:
        loss["sac"] = float(pi_loss.item())
        if hp.norm_closs:
            target = normalize(st, next_obs, force=True)
            pred = z
        else:
            target = next_obs
            pred = zd
        dist = ((pred - target) ** 2).mean()
        loss["dist"] = float(dist.item())
        total = pi_loss + hp.custom_loss * dist
        g_dist = hp.custom_loss * 2.0 * (pred - target) / pred.numel()
        if hp.norm_closs:
            dz = dz + g_dist
        else:
            dzd = dzd + g_dist
    loss["actor"] = float(total.item())
    dq = torch.full_like(q, -1.0 / B)
    dq1, dq2 = _min_bwd(dq, q1, q2)
    _, dx1 = nets.critic_bwd(c1, c1_cache, dq1, need_w=False, need_x=True)
    _, dx2 = nets.critic_bwd(c2, c2_cache, dq2, need_w=False, need_x=True)
    ob = obs.shape[1]
    da = dx1[:, ob:] + dx2[:, ob:]
    if hp.acm_critic:
        _, dxm = nets.acm_bwd(acm, m_cache, da, hp.acm_lim, need_w=False, need_x=True)
        dzd = dzd + dxm[:, ob:]
    else:
        dzd = dzd + da
    dz = dz + dzd * d_scale
    dlogp = torch.full_like(logp, alpha / B)
    a_grads, _ = nets.sac_actor_bwd(actor, a_cache, dz, dlogp, hp.actor_lim)
    if capture is not None:
        capture["y"] = y.clone(); capture["logp"] = logp.clone(); capture["z"] = z.clone()
        capture["q1_pi"] = q1.clone(); capture["q2_pi"] = q2.clone(); capture["a_pi"] = a_pi.clone()
        capture["dz"] = dz.clone()
        capture.update({"actor.grad." + k: v.clone() for k, v in a_grads.items()})
    adam_step_net(state, "actor", a_grads, hp.actor_lr)

    # ---- Phase D: Polyak on both critics, sac_acm.py:148 -> sac.py:186-199
    _polyak(state, "critic_1", "critic_1_targ", hp.tau)
    _polyak(state, "critic_2", "critic_2_targ", hp.tau)

    # ---- Phase E: temperature, sac_acm.py:154-159 -> sac.py:201-216.
    # alpha_loss = (exp(log_alpha) * (-logp - H)).mean(): log_alpha is a 0-dim float64 tensor, so the
    # product is evaluated in float32 with exp(log_alpha) rounded to float32; the gradient w.r.t.
    # log_alpha is sum_b[(1/B) * (-logp_b - H)] (float32) cast to float64, times exp(log_alpha).
    la = state["log_alpha"]
    ea64 = la.exp()
    term = (-logp - hp.target_entropy)
    loss["alpha"] = float((ea64.to(torch.float32) * term).mean().item())
    g_ea = (torch.full_like(term, 1.0 / B) * term).sum().to(torch.float64)
    g_la = g_ea * ea64
    _adam_scalar64(state, "log_alpha", g_la, hp.alpha_lr)
    return loss, sac_alpha(state)


def st_scale(st: NormStats):
    if st.min_max_denormalize:
        return (st.max_obs - st.min_obs) / 2
    return st.obs_std + 1e-8


def _adam_scalar64(state, key, grad, lr):
    """torch Adam on the float64 0-dim log_alpha (same recurrence, double arithmetic)."""
    step = int(state.get(key + "#step", 0)) + 1
    state[key + "#step"] = step
    if key + "#m" not in state:
        state[key + "#m"] = torch.zeros((), dtype=torch.float64)
        state[key + "#v"] = torch.zeros((), dtype=torch.float64)
    m, v, p = state[key + "#m"], state[key + "#v"], state[key]
    m.lerp_(grad, 1 - BETA1)
    v.mul_(BETA2).addcmul_(grad, grad, value=1 - BETA2)
    bc1, bc2 = 1 - BETA1 ** step, 1 - BETA2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(EPS)
    p.addcdiv_(m, denom, value=-(lr / bc1))


def ddpg_acm_update(state, hp: OffPolicyHP, st: NormStats, obs, next_obs, action, reward, done,
                    acm_action, capture=None):
    """One DDPG_AcM.update.  Mutates `state`; returns the losses dict."""
    loss = {}
    actor, actor_t = sub(state, "actor"), sub(state, "actor_targ")
    critic, critic_t, acm = sub(state, "critic"), sub(state, "critic_targ"), sub(state, "acm")
    B, ob = obs.shape
    if hp.acm_critic:                                   # ddpg_acm.py:169-170
        action = acm_action

    # ---- Q target (no grad), ddpg_acm.py:113-121
    z_n, _ = nets.ddpg_actor_fwd(actor_t, next_obs, hp.actor_lim)
    zd_n = denormalize(st, z_n)
    if hp.acm_critic:
        a_n, _ = nets.acm_fwd(acm, torch.cat([next_obs, zd_n], dim=1), hp.acm_lim)
    else:
        a_n = zd_n
    q_t, _ = nets.critic_fwd(critic_t, next_obs, a_n)
    y = reward + hp.gamma * (1 - done) * q_t

    # ---- critic, ddpg_acm.py:175-182
    q, cache = nets.critic_fwd(critic, obs, action)
    diff = q - y
    loss["critic"] = float((diff * diff).mean().item())
    grads, _ = nets.critic_bwd(critic, cache, 2.0 * diff / B)
    if capture is not None:
        capture["y"] = y.clone(); capture["critic.q"] = q.clone()
        capture.update({"critic.grad." + k: v.clone() for k, v in grads.items()})
    adam_step_net(state, "critic", grads, hp.critic_lr)

    # ---- policy, ddpg_acm.py:125-145,187-192 (post-step critic)
    z, a_cache = nets.ddpg_actor_fwd(actor, obs, hp.actor_lim)
    zd = denormalize(st, z)
    if hp.acm_critic:
        a_pi, m_cache = nets.acm_fwd(acm, torch.cat([obs, zd], dim=1), hp.acm_lim)
    else:
        a_pi = zd
    q_pi, c_cache = nets.critic_fwd(critic, obs, a_pi)
    pi_loss = -q_pi.mean()
    total = pi_loss
    dz = torch.zeros_like(z)
    dzd = torch.zeros_like(z)
    if hp.custom_loss:
        loss["ddpg"] = float(pi_loss.item())
        if hp.norm_closs:
            target, pred = normalize(st, next_obs, force=True), z
        else:
            target, pred = next_obs, zd
        dist = ((pred - target) ** 2).mean()
        loss["dist"] = float(dist.item())
        total = pi_loss + hp.custom_loss * dist
        g_dist = hp.custom_loss * 2.0 * (pred - target) / pred.numel()
        if hp.norm_closs:
            dz = dz + g_dist
        else:
            dzd = dzd + g_dist
    loss["actor"] = float(total.item())
    _, dx = nets.critic_bwd(critic, c_cache, torch.full_like(q_pi, -1.0 / B), need_w=False, need_x=True)
    da = dx[:, ob:]
    if hp.acm_critic:
        _, dxm = nets.acm_bwd(acm, m_cache, da, hp.acm_lim, need_w=False, need_x=True)
        dzd = dzd + dxm[:, ob:]
    else:
        dzd = dzd + da
    dz = dz + dzd * st_scale(st)
    a_grads, _ = nets.ddpg_actor_bwd(actor, a_cache, dz, hp.actor_lim)
    if capture is not None:
        capture["z"] = z.clone(); capture["q_pi"] = q_pi.clone(); capture["dz"] = dz.clone()
        capture.update({"actor.grad." + k: v.clone() for k, v in a_grads.items()})
    adam_step_net(state, "actor", a_grads, hp.actor_lr)

    # ---- Polyak: critic then actor, ddpg.py:273-284
    _polyak(state, "critic", "critic_targ", hp.tau)
    _polyak(state, "actor", "actor_targ", hp.tau)
    return loss


def acm_batch_update(state, x, y, acm_lim, lr):
    """AcMTrainer.batch_update (continuous), rltoolkit/acm/acm.py:246-258: MSE(acm(x), y), Adam."""
    acm = sub(state, "acm")
    if y.dim() < 2:
        y = y.reshape(len(y), -1)
    pred, cache = nets.acm_fwd(acm, x, acm_lim)
    diff = pred - y
    loss = float((diff * diff).mean().item())
    grads, _ = nets.acm_bwd(acm, cache, 2.0 * diff / diff.numel(), acm_lim, need_w=True, need_x=False)
    adam_step_net(state, "acm", grads, lr)
    return loss


def acm_update_epochs(state, obs, next_obs, actions_acm, perms, batch_size, acm_lim, lr0, sched_step, sched_gamma, epoch0=0):
    """AcMTrainer.update_acm, rltoolkit/acm/acm.py:266-303: per epoch one shuffled pass (DataLoader(shuffle=True), the partial
    last minibatch kept), then acm_scheduler.step() (StepLR, acm.py:181-183).  perms[e] stands for the sampler's permutation.
    Returns the per-epoch mean minibatch losses (self.loss["acm"] after each epoch)."""
    x = torch.cat([obs, next_obs], dim=1)
    out = []
    for e, perm in enumerate(perms):
        lr = lr0 * sched_gamma ** ((epoch0 + e) // sched_step)
        tot, nb = 0.0, 0
        for b0 in range(0, len(perm), batch_size):
            sel = torch.as_tensor(perm[b0:b0 + batch_size], dtype=torch.long)
            tot += acm_batch_update(state, x[sel], actions_acm[sel], acm_lim, lr)
            nb += 1
        out.append(tot / nb)
    return out


def acm_validation_loss(state, obs, next_obs, actions_acm, acm_lim):
    """AcMTrainer.calculate_validation_loss, rltoolkit/acm/acm.py:329-343: one no-grad MSE over the whole validation buffer."""
    pred, _ = nets.acm_fwd(sub(state, "acm"), torch.cat([obs, next_obs], dim=1), acm_lim)
    diff = pred - actions_acm
    return float((diff * diff).mean().item())
