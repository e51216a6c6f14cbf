"""Plain-tensor forward / backward of every network on the SPP-RL hot path (ORACLE, test infra).

All tensors are torch CPU float32.  Parameter dicts use the reference's state_dict key names
(`fc1.weight`, `fc1.bias`, ...).  Backward functions take the cache returned by the forward and
return `(grads: dict | None, dx)`; they reproduce autograd's arithmetic, not a simplification of it.
"""
import math

import torch

LOG_SQRT_2PI = math.log(math.sqrt(2 * math.pi))
LOG2 = math.log(2.0)


def linear(x, w, b):
    """torch.nn.functional.linear: x @ w.T + b."""
    return torch.addmm(b, x, w.t())


def relu_bwd(grad, out):
    """threshold_backward as used by torch.relu: grad where out > 0."""
    return grad * (out > 0).to(grad.dtype)


# --------------------------------------------------------------------------- critics
def critic_fwd(p, obs, act):
    """SAC_Critic.forward / Critic.forward, rltoolkit/algorithms/sac/models.py:81-91 and
    rltoolkit/algorithms/ddpg/models.py:39-44: fc3(relu(fc2(relu(fc1 cat[obs, a])))).squeeze(-1)."""
    x = torch.cat((obs, act), dim=-1)
    h1 = torch.relu(linear(x, p["fc1.weight"], p["fc1.bias"]))
    h2 = torch.relu(linear(h1, p["fc2.weight"], p["fc2.bias"]))
    q = linear(h2, p["fc3.weight"], p["fc3.bias"]).squeeze(-1)
    return q, (x, h1, h2)


def critic_bwd(p, cache, dq, need_w=True, need_x=False):
    x, h1, h2 = cache
    dq2 = dq.unsqueeze(-1)
    dz2 = relu_bwd(dq2 @ p["fc3.weight"], h2)
    dz1 = relu_bwd(dz2 @ p["fc2.weight"], h1)
    grads = None
    if need_w:
        grads = {
            "fc3.weight": dq2.t() @ h2, "fc3.bias": dq2.sum(0),
            "fc2.weight": dz2.t() @ h1, "fc2.bias": dz2.sum(0),
            "fc1.weight": dz1.t() @ x, "fc1.bias": dz1.sum(0),
        }
    dx = dz1 @ p["fc1.weight"] if need_x else None
    return grads, dx


# --------------------------------------------------------------------------- SAC actor
def sac_actor_fwd(p, x, eps, lim, log_scale_min=-20.0, log_scale_max=2.0):
    """SAC_Actor.forward (continuous branch), rltoolkit/algorithms/sac/models.py:24-54.

    `eps` is the standard-normal draw Normal.rsample makes (loc + eps * scale)."""
    h1 = torch.relu(linear(x, p["fc1.weight"], p["fc1.bias"]))
    h2 = torch.relu(linear(h1, p["fc2.weight"], p["fc2.bias"]))
    mu = linear(h2, p["fc_prob.weight"], p["fc_prob.bias"])
    ls_raw = linear(h2, p["fc_scale.weight"], p["fc_scale.bias"])
    ls = torch.clamp(ls_raw, log_scale_min, log_scale_max)
    std = torch.exp(ls)
    u = mu + eps * std
    # Normal.log_prob (torch/distributions/normal.py): var = scale**2, log_scale = scale.log()
    var = std ** 2
    logp_el = -((u - mu) ** 2) / (2 * var) - std.log() - LOG_SQRT_2PI
    logp = logp_el.sum(-1)
    corr = 2 * (LOG2 - u - torch.nn.functional.softplus(-2 * u)).sum(1)
    logp = logp - corr
    th = torch.tanh(u)
    z = th * lim
    return z, logp, (x, h1, h2, mu, ls_raw, std, u, th, eps)


def sac_actor_bwd(p, cache, dz, dlogp, lim, need_w=True, log_scale_min=-20.0, log_scale_max=2.0):
    """Backward of sac_actor_fwd for upstream grads dz [B,ob] (on z) and dlogp [B] (on logp).

    Mirrors autograd's accumulation: the rsample node u receives (tanh path) + (correction path)
    + (log_prob `u - mu` path); loc receives the negated log_prob term and then du; the scale
    receives var, log and eps paths.  The near-cancelling terms are kept, not simplified."""
    x, h1, h2, mu, ls_raw, std, u, th, eps = cache
    g = dlogp.unsqueeze(-1)
    var = std ** 2
    d = u - mu
    # ---- log_prob element: t = -(d**2)/(2 var) - log(std) - c
    ds = -g / (2 * var)                 # grad wrt d**2
    dd = ds * 2 * d                     # grad wrt d = u - mu
    d2var = g * (d ** 2) / ((2 * var) ** 2)   # grad wrt (2 var):  -(-s)/(2var)^2 * g
    dvar = 2 * d2var
    dstd = dvar * 2 * std + (-g / std)
    # ---- correction: logp -= 2*(log2 - u - softplus(-2u)).sum()
    #      d/du = -2 * (-1 - (-2) * sigmoid(-2u)) * g  -> autograd: softplus' = sigmoid(-2u)
    du = dz * lim * (1 - th * th)
    du = du + (-g) * 2 * (-1.0 + 2.0 * torch.sigmoid(-2 * u))
    du = du + dd
    dmu = -dd + du
    dstd = dstd + du * eps
    dls = dstd * std
    dls_raw = dls * ((ls_raw >= log_scale_min) & (ls_raw <= log_scale_max)).to(dls.dtype)
    dh2 = dmu @ p["fc_prob.weight"] + dls_raw @ p["fc_scale.weight"]
    dz2 = relu_bwd(dh2, h2)
    dz1 = relu_bwd(dz2 @ p["fc2.weight"], h1)
    grads = None
    if need_w:
        grads = {
            "fc_prob.weight": dmu.t() @ h2, "fc_prob.bias": dmu.sum(0),
            "fc_scale.weight": dls_raw.t() @ h2, "fc_scale.bias": dls_raw.sum(0),
            "fc2.weight": dz2.t() @ h1, "fc2.bias": dz2.sum(0),
            "fc1.weight": dz1.t() @ x, "fc1.bias": dz1.sum(0),
        }
    return grads, None


def sac_actor_det(p, x, lim):
    """SAC_Actor.forward(deterministic=True): tanh(mean) * lim (sac/models.py:42-43,52-53)."""
    h1 = torch.relu(linear(x, p["fc1.weight"], p["fc1.bias"]))
    h2 = torch.relu(linear(h1, p["fc2.weight"], p["fc2.bias"]))
    return torch.tanh(linear(h2, p["fc_prob.weight"], p["fc_prob.bias"])) * lim


# --------------------------------------------------------------------------- DDPG actor
def ddpg_actor_fwd(p, x, lim):
    """Actor.forward, rltoolkit/algorithms/ddpg/models.py:17-22: tanh(fc3(relu(fc2(relu(fc1 x)))))*lim."""
    h1 = torch.relu(linear(x, p["fc1.weight"], p["fc1.bias"]))
    h2 = torch.relu(linear(h1, p["fc2.weight"], p["fc2.bias"]))
    th = torch.tanh(linear(h2, p["fc3.weight"], p["fc3.bias"]))
    return th * lim, (x, h1, h2, th)


def ddpg_actor_bwd(p, cache, dz, lim):
    x, h1, h2, th = cache
    d3 = dz * lim * (1 - th * th)
    dz2 = relu_bwd(d3 @ p["fc3.weight"], h2)
    dz1 = relu_bwd(dz2 @ p["fc2.weight"], h1)
    return {
        "fc3.weight": d3.t() @ h2, "fc3.bias": d3.sum(0),
        "fc2.weight": dz2.t() @ h1, "fc2.bias": dz2.sum(0),
        "fc1.weight": dz1.t() @ x, "fc1.bias": dz1.sum(0),
    }, None


# --------------------------------------------------------------------------- ACM variants
def acm_kind(p):
    return "basic" if "fc21.weight" in p else "acm"


def acm_fwd(p, x, ac_lim):
    """AcM.forward (continuous), rltoolkit/basic_model.py:118-126, or BasicAcM.forward,
    rltoolkit/acm/models/basic_acm.py:24-28, selected by the parameter keys."""
    if acm_kind(p) == "acm":
        h1 = torch.tanh(linear(x, p["fc1.weight"], p["fc1.bias"]))
        h2 = torch.tanh(linear(h1, p["fc2.weight"], p["fc2.bias"]))
        t3 = torch.tanh(linear(h2, p["fc3.weight"], p["fc3.bias"]))
        return t3 * ac_lim, (x, h1, h2, t3)
    h = torch.tanh(linear(x, p["fc1.weight"], p["fc1.bias"]))
    s = linear(x, p["fc21.weight"], p["fc21.bias"])
    h1 = torch.tanh(linear(h, p["fc2.weight"], p["fc2.bias"]) + p["t"] * s)
    t3 = torch.tanh(linear(h1, p["fc3.weight"], p["fc3.bias"]))
    return t3 * p["t1"], (x, h, s, h1, t3)


def acm_bwd(p, cache, da, ac_lim, need_w=False, need_x=True):
    """Backward of acm_fwd.  need_w for the ACM regression (acm.py:246-258); need_x for the
    actor losses, where gradient flows THROUGH the frozen ACM (sac_acm.py:66-72)."""
    grads, dx = None, None
    if acm_kind(p) == "acm":
        x, h1, h2, t3 = cache
        d3 = da * ac_lim * (1 - t3 * t3)
        d2 = (d3 @ p["fc3.weight"]) * (1 - h2 * h2)
        d1 = (d2 @ p["fc2.weight"]) * (1 - h1 * h1)
        if need_w:
            grads = {
                "fc3.weight": d3.t() @ h2, "fc3.bias": d3.sum(0),
                "fc2.weight": d2.t() @ h1, "fc2.bias": d2.sum(0),
                "fc1.weight": d1.t() @ x, "fc1.bias": d1.sum(0),
            }
        if need_x:
            dx = d1 @ p["fc1.weight"]
        return grads, dx
    x, h, s, h1, t3 = cache
    d3 = da * p["t1"] * (1 - t3 * t3)
    d2 = (d3 @ p["fc3.weight"]) * (1 - h1 * h1)       # grad of pre-activation of h1
    ds = d2 * p["t"]
    d1 = (d2 @ p["fc2.weight"]) * (1 - h * h)
    if need_w:
        grads = {
            "t1": (da * t3).sum(0), "t": (d2 * s).sum().reshape(1),
            "fc3.weight": d3.t() @ h1, "fc3.bias": d3.sum(0),
            "fc2.weight": d2.t() @ h, "fc2.bias": d2.sum(0),
            "fc21.weight": ds.t() @ x, "fc21.bias": ds.sum(0),
            "fc1.weight": d1.t() @ x, "fc1.bias": d1.sum(0),
        }
    if need_x:
        dx = d1 @ p["fc1.weight"] + ds @ p["fc21.weight"]
    return grads, dx


# --------------------------------------------------------------------------- PPO nets (64-wide tanh)
def ppo_actor_mean(p, x, lim):
    """basic_model.Actor.forward (continuous) * ac_lim, rltoolkit/basic_model.py:23-30,41-42."""
    h1 = torch.tanh(linear(x, p["fc1.weight"], p["fc1.bias"]))
    h2 = torch.tanh(linear(h1, p["fc2.weight"], p["fc2.bias"]))
    t3 = torch.tanh(linear(h2, p["fc3.weight"], p["fc3.bias"]))
    return t3 * lim, (x, h1, h2, t3)


def ppo_actor_mean_bwd(p, cache, dmean, lim):
    x, h1, h2, t3 = cache
    d3 = dmean * lim * (1 - t3 * t3)
    d2 = (d3 @ p["fc3.weight"]) * (1 - h2 * h2)
    d1 = (d2 @ p["fc2.weight"]) * (1 - h1 * h1)
    return {
        "fc3.weight": d3.t() @ h2, "fc3.bias": d3.sum(0),
        "fc2.weight": d2.t() @ h1, "fc2.bias": d2.sum(0),
        "fc1.weight": d1.t() @ x, "fc1.bias": d1.sum(0),
    }


def ppo_critic_fwd(p, x):
    """basic_model.Critic.forward, rltoolkit/basic_model.py:73-77 -> [N,1]."""
    h1 = torch.tanh(linear(x, p["fc1.weight"], p["fc1.bias"]))
    h2 = torch.tanh(linear(h1, p["fc2.weight"], p["fc2.bias"]))
    return linear(h2, p["fc3.weight"], p["fc3.bias"]), (x, h1, h2)


def ppo_critic_bwd(p, cache, dv):
    x, h1, h2 = cache
    d2 = (dv @ p["fc3.weight"]) * (1 - h2 * h2)
    d1 = (d2 @ p["fc2.weight"]) * (1 - h1 * h1)
    return {
        "fc3.weight": dv.t() @ h2, "fc3.bias": dv.sum(0),
        "fc2.weight": d2.t() @ h1, "fc2.bias": d2.sum(0),
        "fc1.weight": d1.t() @ x, "fc1.bias": d1.sum(0),
    }
