"""Shared parity harness: drive the CUDA path through the C ABI and the CPU oracle on identical inputs.

Used by tests/ (-m gpu) and __graft_entry__.smoke().  The oracle is the checker only.
"""
import numpy as np
import torch

from oracle import offpolicy as op
from oracle.norm import NormStats
from spp_rl_b200 import Population, init_state


def relnorm(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def make_stats(ob, seed, min_max):
    rng = np.random.RandomState(1000 + seed)
    mn = (-rng.rand(ob) * 3 - 0.5).astype(np.float32)
    mx = (rng.rand(ob) * 3 + 0.5).astype(np.float32)
    mean = rng.randn(ob).astype(np.float32)
    std = (rng.rand(ob) + 0.5).astype(np.float32)
    return mn, mx, mean, std


def make_batches(ob, ac, P, G, B, seed, mn, mx):
    rng = np.random.RandomState(2000 + seed)
    obs = (rng.rand(P, G, B, ob) * (mx - mn) + mn).astype(np.float32)
    nobs = (obs + 0.05 * rng.randn(P, G, B, ob)).astype(np.float32)
    act = rng.randn(P, G, B, ob).astype(np.float32)
    rew = rng.randn(P, G, B).astype(np.float32)
    done = (rng.rand(P, G, B) < 0.1).astype(np.int8)
    aacm = np.tanh(rng.randn(P, G, B, ac)).astype(np.float32)
    eps = rng.randn(P, G, 2, B, ob).astype(np.float32)
    return obs, nobs, act, rew, done, aacm, eps


def oracle_state(np_state, algo, alpha=0.2):
    s = {k: torch.from_numpy(v.copy()) for k, v in np_state.items()}
    if algo == "sac":
        s["log_alpha"] = torch.tensor(np.log(alpha), dtype=torch.float64)
    return s


def upload_state(pop, np_state, agent, algo):
    nets = ["actor", "critic_1", "critic_2", "critic_1_targ", "critic_2_targ", "acm"] if algo == "sac" else \
        ["actor", "actor_targ", "critic", "critic_targ", "acm"]
    for net in nets:
        sd = {k[len(net) + 1:]: v for k, v in np_state.items() if k.startswith(net + ".")}
        pop.load_state_dict(net, sd, agent=agent)


def compare_states(pop, ostate, agent, algo, verbose=False, tag="", moment_weight=1.0, small_weight=1.0):
    """worst norm-relative error over all tensors and Adam moments of one agent"""
    nets = ["actor", "critic_1", "critic_2", "critic_1_targ", "critic_2_targ", "acm"] if algo == "sac" else \
        ["actor", "actor_targ", "critic", "critic_targ", "acm"]
    worst = 0.0
    for net in nets:
        sd = pop.state_dict(net, agent=agent)
        for k, v in sd.items():
            e = relnorm(v, ostate[net + "." + k].numpy())
            if v.size <= 16:
                e = e * small_weight      # (1.0 for the fp32-accurate path; only the reduced-precision variant relaxes tiny cancelling sums)
            worst = max(worst, e)
            if verbose and e > 1e-6:
                print("  %s agent %d %s.%s relnorm %.3e" % (tag, agent, net, k, e))
        if not net.endswith("_targ") and net != "acm":
            ad, step = pop.adam_state(net, agent=agent)
            for k, (m, v) in ad.items():
                key = net + "." + k
                if key + "#m" in ostate:
                    em, ev = relnorm(m, ostate[key + "#m"].numpy()), relnorm(v, ostate[key + "#v"].numpy())
                    if m.size <= 16:
                        em, ev = em * small_weight, ev * small_weight
                    worst = max(worst, em * moment_weight, ev * moment_weight)
                    if verbose and max(em, ev) > 1e-6:
                        print("  %s agent %d %s moments relnorm m %.3e v %.3e" % (tag, agent, key, em, ev))
            assert step == ostate[net + "#step"], (net, step, ostate[net + "#step"])
    return worst


def run_offpolicy_parity_case(algo="sac", ob=11, ac=3, batch=64, population=2, steps=2, seed=0, custom_loss=0.2,
                              norm_closs=False, acm_critic=True, min_max=True, acm_kind="acm", verbose=False,
                              gamma=0.99, lr=1e-3, small_std=False, actor_lim=1.0, acm_lim=1.0, alpha_tol=1e-6, moment_weight=1.0, small_weight=1.0,
                              per_agent=False, oracle_agents=None):
    """-> worst error over agents; per_agent=True: (list of per-agent worst errors, [(name, array)] dump of every loss / weight / moment
    of every agent for bitwise comparisons between two runs); oracle_agents: the agents the oracle is run for (default: all)."""
    P, G, B = population, steps, batch
    mn, mx, mean, std = make_stats(ob, seed, min_max)
    obs, nobs, act, rew, done, aacm, eps = make_batches(ob, ac, P, G, B, seed, mn, mx)
    pop = Population(algo=algo, ob_dim=ob, ac_dim=ac, population=P, acm_kind=acm_kind, acm_critic=acm_critic,
                     norm_closs=norm_closs, min_max_denormalize=min_max, update_batch_size=B, gamma=gamma,
                     actor_lr=lr, critic_lr=lr, alpha_lr=lr, custom_loss=custom_loss, alpha=0.2, target_entropy=-float(ac))
    alim = np.broadcast_to(np.asarray(actor_lim, np.float32), (ob,)).copy()
    mlim = np.broadcast_to(np.asarray(acm_lim, np.float32), (ac,)).copy()
    pop.set_limits(alim, mlim)
    pop.set_norm_stats(mn, mx, mean, std)
    st = NormStats(min_max, torch.from_numpy(mn), torch.from_numpy(mx), torch.from_numpy(mean), torch.from_numpy(std))
    hp = op.OffPolicyHP(gamma=gamma, actor_lr=lr, critic_lr=lr, alpha_lr=lr, tau=0.005, custom_loss=custom_loss,
                        norm_closs=norm_closs, acm_critic=acm_critic, target_entropy=-float(ac),
                        actor_lim=torch.from_numpy(alim), acm_lim=torch.from_numpy(mlim))
    ostates = []
    for a in range(P):
        s0 = init_state(algo, ob, ac, seed * 100 + a, acm_kind, acm_critic)
        if small_std and algo == "sac":
            s0["actor.fc_scale.bias"][:] = -4.0
        if acm_kind == "basic":
            s0["acm.t"][:] = 0.7
            s0["acm.t1"][:] = np.linspace(0.5, 1.5, ac)
        upload_state(pop, s0, a, algo)
        ostates.append(oracle_state(s0, algo))
    losses = pop.update_host(G, obs, nobs, act, rew, done, aacm, eps=eps if algo == "sac" else None)
    worst = 0.0
    per, dump = [], [("losses", np.array(losses, copy=True))]
    for a in range(P):
        if per_agent:
            per.append(worst)
            worst = 0.0
            for net in (["actor", "critic_1", "critic_2", "critic_1_targ", "critic_2_targ"] if algo == "sac" else ["actor", "actor_targ", "critic", "critic_targ"]):
                for k, v in pop.state_dict(net, agent=a).items():
                    dump.append(("%d.%s.%s" % (a, net, k), v))
                if not net.endswith("_targ"):
                    for k, (m, v) in pop.adam_state(net, agent=a)[0].items():
                        dump.append(("%d.%s.%s#m" % (a, net, k), m)); dump.append(("%d.%s.%s#v" % (a, net, k), v))
            if algo == "sac":
                dump.append(("%d.log_alpha" % a, np.float64(pop.alpha(a)[0])))
        if oracle_agents is not None and a not in oracle_agents:
            continue
        s = ostates[a]
        alpha = None
        for g in range(G):
            t = lambda x: torch.from_numpy(x[a, g])
            if algo == "sac":
                ol, alpha = op.sac_acm_update(s, hp, st, t(obs), t(nobs), t(act), t(rew), t(done), t(aacm),
                                              torch.from_numpy(eps[a, g, 0]), torch.from_numpy(eps[a, g, 1]), alpha)
                pairs = [("critic_1", 0), ("critic_2", 1), ("actor", 2)]
                if custom_loss:
                    pairs += [("sac", 3), ("dist", 4)]
                pairs += [("alpha", 5)]
                if abs(losses[a, g, 6] - alpha) > alpha_tol * abs(alpha):
                    if not per_agent:
                        raise AssertionError("alpha mismatch %r vs %r" % (losses[a, g, 6], alpha))
                    worst = max(worst, min(1.0, abs(losses[a, g, 6] - alpha) / abs(alpha)))
            else:
                ol = op.ddpg_acm_update(s, hp, st, t(obs), t(nobs), t(act), t(rew), t(done), t(aacm))
                pairs = [("critic", 0), ("actor", 2)]
                if custom_loss:
                    pairs += [("ddpg", 3), ("dist", 4)]
            for name, slot in pairs:
                e = abs(losses[a, g, slot] - ol[name]) / (abs(ol[name]) + 1e-3)
                if verbose and e > 1e-5:
                    print("  loss %s agent %d step %d: cuda %.8g oracle %.8g" % (name, a, g, losses[a, g, slot], ol[name]))
                worst = max(worst, min(e, 1.0) if e > 1e-5 else 0.0)
        worst = max(worst, compare_states(pop, s, a, algo, verbose=verbose, tag=algo, moment_weight=moment_weight, small_weight=small_weight))
        if algo == "sac":
            la, _ = pop.alpha(a)
            e = abs(la - float(s["log_alpha"])) / abs(float(s["log_alpha"]))
            worst = max(worst, e)
    pop.close()
    if per_agent:
        per.append(worst)
        return per[1:], dump
    return worst


def run_sac_parity_case(**kw):
    return run_offpolicy_parity_case(algo="sac", **kw)


def debug_first_step(algo="sac", ob=11, ac=3, batch=64, seed=0, custom_loss=0.2, norm_closs=False, acm_critic=True,
                     min_max=True, acm_kind="acm"):
    """One update of one agent; print the error of every intermediate the kernel leaves in scratch."""
    P, G, B = 1, 1, batch
    mn, mx, mean, std = make_stats(ob, seed, min_max)
    obs, nobs, act, rew, done, aacm, eps = make_batches(ob, ac, P, G, B, seed, mn, mx)
    pop = Population(algo=algo, ob_dim=ob, ac_dim=ac, population=P, acm_kind=acm_kind, acm_critic=acm_critic,
                     norm_closs=norm_closs, min_max_denormalize=min_max, update_batch_size=B, gamma=0.99,
                     custom_loss=custom_loss, alpha=0.2, target_entropy=-float(ac))
    pop.set_norm_stats(mn, mx, mean, std)
    st = NormStats(min_max, torch.from_numpy(mn), torch.from_numpy(mx), torch.from_numpy(mean), torch.from_numpy(std))
    hp = op.OffPolicyHP(gamma=0.99, custom_loss=custom_loss, norm_closs=norm_closs, acm_critic=acm_critic,
                        target_entropy=-float(ac), actor_lim=torch.ones(ob), acm_lim=torch.ones(ac))
    s0 = init_state(algo, ob, ac, seed * 100, acm_kind, acm_critic)
    upload_state(pop, s0, 0, algo)
    s = oracle_state(s0, algo)
    losses = pop.update_host(1, obs, nobs, act, rew, done, aacm, eps=eps if algo == "sac" else None)
    cap = {}
    t = lambda x: torch.from_numpy(x[0, 0])
    if algo == "sac":
        ol, _ = op.sac_acm_update(s, hp, st, t(obs), t(nobs), t(act), t(rew), t(done), t(aacm),
                                  torch.from_numpy(eps[0, 0, 0]), torch.from_numpy(eps[0, 0, 1]), None, capture=cap)
    else:
        ol = op.ddpg_acm_update(s, hp, st, t(obs), t(nobs), t(act), t(rew), t(done), t(aacm), capture=cap)
    print("losses cuda", losses[0, 0], "oracle", ol)
    vec = pop.debug_scratch(0, "vec")
    print("y     relnorm %.3e" % relnorm(vec[2, :B], cap["y"].numpy()))
    if algo == "sac":
        print("logp  relnorm %.3e" % relnorm(vec[4, :B], cap["logp"].numpy()))
    dml = pop.debug_scratch(0, "dml")
    xm = pop.debug_scratch(0, "xm")
    xcp = pop.debug_scratch(0, "xcp")
    ldo = (ob + 3) // 4 * 4
    if acm_critic:
        print("a_pi  relnorm %.3e" % relnorm(xcp[:, ldo:ldo + ac], cap["a_pi"].numpy()))
    print("worst state relnorm %.3e" % compare_states(pop, s, 0, algo, verbose=True))
    pop.close()
