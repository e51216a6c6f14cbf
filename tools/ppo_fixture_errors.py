import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from oracle import ppo as P
from spp_rl_b200.ppo import PpoPolicy
from tests.parity_util import relnorm
G = "/root/repo/tests/golden"
g = np.load(os.path.join(G, "ppo_walker.npz"))
gamma, lam, eps_clip, kl_thr, max_ep, bs, a_lr, c_lr, ent, closs, ntu, nupt = [float(x) for x in g["hp"]]
ob, ac = g["chain"].shape[1], g["actions_acm"].shape[1]
oi, ni = P.chain_views(len(g["chain"]), list(g["joints"]))
obs, nobs = g["chain"][oi], g["chain"][ni]
N = obs.shape[0]
for tc in (True, False):
    pol = PpoPolicy(ob, ac, max_rows=N, max_batch_rows=int(bs), min_max_denormalize=True, norm_closs=False, gamma=gamma, gae_lambda=lam,
                    ppo_epsilon=eps_clip, entropy_coef=ent, custom_loss=closs, actor_lr=a_lr, critic_lr=c_lr)
    pol.set_critic_path(tc)
    pol.set_limits(float(g["actor_lim"]))
    pol.set_norm_stats(g["min_obs"], g["max_obs"], g["obs_mean"], g["obs_std"])
    for net in ("actor", "critic"):
        pol.load_state_dict(net, {k[len("pre:" + net) + 1:]: g[k] for k in g.files if k.startswith("pre:" + net + ".")})
    end = g["end"]; starts, lens, s = [], [], 0
    for i, e in enumerate(end):
        if e: starts.append(s); lens.append(i + 1 - s); s = i + 1
    pol.load_rollout(obs, nobs, g["actions"], g["logp"], g["rewards"], g["done"], g["end"], np.array(starts, np.int64), np.array(lens, np.int64))
    loss = pol.update_critic(int(ntu), int(nupt))
    print("tc", tc, "critic loss rel", abs(loss - float(g["critic_loss"])) / float(g["critic_loss"]))
    for k, v in pol.state_dict("critic").items():
        print("   critic", k, relnorm(v, g["fit:critic." + k]))
    adv = pol.advantages()
    print("   adv", np.abs(adv - g["adv"]).max() / max(1.0, np.abs(g["adv"]).max()))
    pol.normalize_adv()
    losses, epochs, kl = pol.update_actor(g["perms"], int(bs), kl_thr, int(max_ep))
    for k, v in pol.state_dict("actor").items():
        print("   actor", k, relnorm(v, g["post:actor." + k]))
    for key, r in zip(("actor", "entropy", "policy", "dist"), g["actor_losses"]):
        print("   loss", key, abs(losses[key] - float(r)) / abs(float(r)))
    pol.close()
